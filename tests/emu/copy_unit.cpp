// copy_unit.cpp -- TEST INFRASTRUCTURE: randomized host-side check of the resolve kernel's ring
// copy primitives (res_copy_piece / res_copy_match) for every alignment, offset class
// (overlapping, short, far), ring wrap-around, and of the pending-bitmap range operations.
#include <vector>
#include <random>
#include "inflate_resolve.cu"
u32 *ldb_inflate_resolve_counter(const ldb_inflate_args &, const ldb_launch_cfg &) { return nullptr; }
int ldb_fail(int err, const char *, const char *, int) { return err; }
int main() {
    std::mt19937 rng(1);
    int bad = 0;
    std::vector<u8> ringv(RES_RING + 64);
    u8 *ring = (u8 *)(((uintptr_t)ringv.data() + 15) & ~(uintptr_t)15);
    for (int iter = 0; iter < 40000 && bad < 5; iter++) {
        // start anywhere in the ring (so that wrap-around is exercised), any alignment
        u32 q0 = rng() % (2 * RES_RING) + 70000;
        std::vector<u8> expect;
        u32 q = q0;
        int ntok = 1 + rng() % 60;
        for (int t = 0; t < ntok; t++) {
            if (expect.empty() || rng() % 3 == 0) {
                u8 b = rng(); expect.push_back(b); ring[q & RES_MASK] = b; q++;
            } else {
                u32 maxoff = (u32)std::min<size_t>(expect.size(), (rng() & 1) ? 40 : 3000);
                u32 off = 1 + rng() % maxoff;
                u32 len = 3 + rng() % ((rng() & 3) ? 30 : 256);
                for (u32 k = 0; k < len; k++) expect.push_back(expect[expect.size() - off]);
                res_copy_match(ring, q, off, len);
                q += len;
            }
        }
        for (size_t k = 0; k < expect.size(); k++)
            if (ring[(q0 + k) & RES_MASK] != expect[k]) {
                bad++;
                printf("iter %d mismatch at %zu of %zu (q0 %u)\n", iter, k, expect.size(), q0);
                break;
            }
    }
    // bitmap ranges
    std::vector<u32> bm(RES_SPAN / 32, 0), ref(RES_SPAN, 0);
    for (int iter = 0; iter < 20000 && bad < 5; iter++) {
        u32 lo = rng() % (RES_SPAN - 1), hi = lo + 1 + rng() % std::min<u32>(300, RES_SPAN - lo);
        int op = rng() % 3;
        if (op == 0) { res_bits_set(bm.data(), lo, hi); for (u32 k = lo; k < hi; k++) ref[k] = 1; }
        else if (op == 1) { res_bits_clear(bm.data(), lo, hi); for (u32 k = lo; k < hi; k++) ref[k] = 0; }
        else {
            bool any = false;
            for (u32 k = lo; k < hi; k++) any |= ref[k] != 0;
            if (any != res_bits_any(bm.data(), lo, hi)) { bad++; printf("bitmap any mismatch [%u,%u)\n", lo, hi); }
        }
    }
    printf("bad %d\n", bad);
    return bad != 0;
}
