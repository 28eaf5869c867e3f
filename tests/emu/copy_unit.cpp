// copy_unit.cpp -- TEST INFRASTRUCTURE: randomized host-side check of the resolve kernel's per-lane
// copy primitives (res_copy_piece: window ring -> staging ring, res_lit_piece: global -> staging)
// for every alignment, length 1..16 and ring wrap-around; neighbours of the destination range
// must stay untouched.
#include <vector>
#include <random>
#include "inflate_resolve.cu"
u32 *ldb_inflate_resolve_counter(const ldb_inflate_args &, const ldb_launch_cfg &) { return nullptr; }
int ldb_fail(int err, const char *, const char *, int) { return err; }
int main() {
    std::mt19937 rng(1);
    int bad = 0;
    std::vector<u8> winv((1u << 20) + 64), stgv(RES_STG + 64), litv(4096 + 64);
    u8 *win = (u8 *)(((uintptr_t)winv.data() + 15) & ~(uintptr_t)15);
    u8 *stg = (u8 *)(((uintptr_t)stgv.data() + 15) & ~(uintptr_t)15);
    u8 *lit = (u8 *)(((uintptr_t)litv.data() + 15) & ~(uintptr_t)15);
    for (u32 i = 0; i < (1u << 20) + 32; i++) win[i] = (u8)rng();
    for (u32 i = 0; i < 4096; i++) lit[i] = (u8)rng();
    for (int iter = 0; iter < 200000 && bad < 5; iter++) {
        for (u32 i = 0; i < RES_STG; i++) stg[i] = (u8)(i * 7 + iter);
        std::vector<u8> before(stg, stg + RES_STG);
        u32 qd = rng() % (1u << 20), m = 1 + rng() % 16;
        bool from_lit = rng() & 1;
        u32 qs = rng() % (1u << 20), ls = 4 + rng() % 4000;
        if (from_lit) res_lit_piece(lit + ls, stg, qd, m);
        else res_copy_piece(win, stg, qd, qs, m);
        for (u32 k = 0; k < m; k++) before[(qd + k) & RES_SMASK] = from_lit ? lit[ls + k] : win[qs + k];
        if (memcmp(before.data(), stg, RES_STG)) { bad++; printf("iter %d mismatch qd %u qs %u m %u lit %d\n", iter, qd, qs, m, (int)from_lit); }
    }
    printf("bad %d\n", bad);
    return bad != 0;
}
