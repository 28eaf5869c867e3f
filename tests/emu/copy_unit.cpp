// copy_unit.cpp -- TEST INFRASTRUCTURE: randomized host-side check of the inflate kernel output
// primitives (inf_put_byte / inf_put_word / inf_put_bytes / inf_copy_match) for every
// alignment, offset class (<8 pattern, 8..15 near, >=16 far) and overlap.
#include <vector>
#include <random>
#include "inflate_kernel.cu"
int main() {
    std::mt19937 rng(1);
    int bad = 0;
    for (int iter = 0; iter < 60000 && bad < 5; iter++) {
        std::vector<u8> buf(4096 + 64), ref(4096 + 64);
        int mis = rng() % 4;
        inf_lane s; memset(&s, 0, sizeof(s));
        u8 *base = (u8*)(((uintptr_t)buf.data() + 15) & ~15) + mis;
        s.out = base; s.out_avail = 3000; s.out_pos = 0; s.acc = 0; s.cnt = (u32)(uintptr_t)s.out & 3;
        std::vector<u8> expect;
        int ntok = 1 + rng() % 40;
        for (int t = 0; t < ntok; t++) {
            if (expect.empty() || rng() % 3 == 0) { u8 b = rng(); expect.push_back(b); inf_put_byte(s, b); }
            else {
                u32 off = 1 + rng() % std::min<size_t>(expect.size(), 40);
                u32 len = 3 + rng() % 60;
                for (u32 k = 0; k < len; k++) expect.push_back(expect[expect.size() - off]);
                inf_copy_match(s, len, off);
            }
        }
        inf_flush_pending(s);
        if (s.out_pos != expect.size() || memcmp(base, expect.data(), expect.size())) {
            bad++;
            size_t k = 0; while (k < expect.size() && base[k] == expect[k]) k++;
            printf("iter %d mismatch at %zu of %zu (mis %d)\n", iter, k, expect.size(), mis);
        }
    }
    printf("bad %d\n", bad);
    return bad != 0;
}
