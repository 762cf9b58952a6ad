// feasibility probe: can the host pack the upper three bytes of int32 DFSDM words (low byte constant) fast enough to
// make a 3-byte PCIe transport pay?  gcc -O3 -march=native -fopenmp pack24.c -o pack24
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
    const size_t n = (size_t) 155648 * 2048;                  // samples of one bench step
    int nth = argc > 1 ? atoi(argv[1]) : omp_get_max_threads();
    int32_t* src = aligned_alloc(64, n * 4);
    uint8_t* dst = aligned_alloc(64, n * 3 + 64);
#pragma omp parallel for num_threads(nth) schedule(static)
    for (size_t i = 0; i < n; ++i) { src[i] = (int32_t) ((i * 2654435761u) & 0xffffff00u); }
#pragma omp parallel for num_threads(nth) schedule(static)
    for (size_t i = 0; i < n * 3; i += 4096) dst[i] = 0;
    uint8_t idx[64];
    for (int i = 0; i < 16; ++i) { idx[3 * i] = 4 * i + 1; idx[3 * i + 1] = 4 * i + 2; idx[3 * i + 2] = 4 * i + 3; }
    for (int i = 48; i < 64; ++i) idx[i] = 0;
    const __m512i perm = _mm512_loadu_si512(idx);
    double best = 1e9;
    uint32_t bad_total = 0;
    for (int rep = 0; rep < 6; ++rep) {
        double t0 = now();
        uint32_t bad = 0;
#pragma omp parallel for num_threads(nth) schedule(static) reduction(| : bad)
        for (size_t blk = 0; blk < n / 4096; ++blk) {
            const int32_t* s = src + blk * 4096;
            uint8_t* d = dst + blk * 4096 * 3;
            __m512i lowacc = _mm512_setzero_si512();
            for (int j = 0; j < 4096; j += 16) {
                const __m512i v = _mm512_load_si512((const void*) (s + j));
                lowacc = _mm512_or_si512(lowacc, v);
                const __m512i p = _mm512_permutexvar_epi8(perm, v);
                _mm512_mask_storeu_epi8(d + 3 * j, 0x0000ffffffffffffull, p);
            }
            bad |= _mm512_reduce_or_epi32(_mm512_and_si512(lowacc, _mm512_set1_epi32(0xff)));
        }
        double dt = now() - t0;
        if (dt < best) best = dt;
        bad_total |= bad;
    }
    printf("threads %d: %.2f ms per step  %.1f GB/s of int32 input (%.1f GB/s written)  low bytes zero: %s\n", nth, best * 1e3,
           n * 4 / best / 1e9, n * 3 / best / 1e9, bad_total ? "no" : "yes");
    return 0;
}
