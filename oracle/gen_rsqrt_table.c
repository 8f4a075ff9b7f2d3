/* TEST/BUILD TOOL. Dumps the x86 RSQRTSS approximation as a 2048-entry table.
 *
 * The reference's learning rate is eta * _mm_rsqrt_ps(G)   (mf/mf.cpp:1469-1470).
 * On Intel CPUs the result of RSQRTPS/RSQRTSS depends only on the exponent
 * parity and the top 10 mantissa bits of the input, and the exponent scales
 * exactly:  rsqrt(4x) == rsqrt(x)/2.  This tool (a) dumps the 2048 results for
 * x in [1,4) and (b) verifies the model exhaustively against the instruction.
 * The committed tables (oracle/rsqrt12_table.h and the product's copy under
 * question-recommendation-system_b200/csrc/) were produced by this tool on
 * an Intel Xeon (Sapphire Rapids class); `make -C oracle check-rsqrt` re-verifies.
 *
 *   gen_rsqrt_table dump   > table body
 *   gen_rsqrt_table check  -> exit 0 iff table model == hardware for every finite input probed
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <xmmintrin.h>

static uint32_t hw(uint32_t xb) {
    float x, y;
    uint32_t yb;
    memcpy(&x, &xb, 4);
    y = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x)));
    memcpy(&yb, &y, 4);
    return yb;
}

int main(int argc, char **argv) {
    static uint32_t tab[2048];
    int par, i;
    for (par = 0; par < 2; par++)
        for (i = 0; i < 1024; i++)
            tab[par * 1024 + i] = hw(((127u + (uint32_t)par) << 23) | ((uint32_t)i << 13));
    if (argc > 1 && strcmp(argv[1], "dump") == 0) {
        for (i = 0; i < 2048; i++) printf("0x%08xu,%s", tab[i], (i % 8 == 7) ? "\n" : " ");
        return 0;
    }
    long bad = 0;
    uint32_t e, m;
    for (e = 1; e < 255; e++)
        for (m = 0; m < (1u << 23); m += (e >= 120 && e <= 140) ? 1 : 97) {
            uint32_t p = (e & 1) ? 0 : 1;
            int32_t sh = ((int32_t)e - (127 + (int32_t)p)) / 2;
            uint32_t pred = tab[p * 1024 + (m >> 13)] - ((uint32_t)sh << 23);
            if (pred != hw((e << 23) | m)) bad++;
        }
    printf("rsqrt table model mismatches: %ld\n", bad);
    return bad != 0;
}
