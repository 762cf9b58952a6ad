/*
 * ref_fast_dispatch.c — CPU ORACLE (test infrastructure, NOT product code): picks the widest SIMD build of
 * ref_fast.c the host supports.  Both builds produce the same bits (every lane is one frame).
 */
#include "ref_dsp.h"

void ref_fast_demod_frames_vw16(const ref_receiver *rx, const void *pcm, int is_float, size_t nframes, float *mag_up,
                                uint32_t *idx_up, float *mag_down, uint32_t *idx_down, int nthreads);
void ref_fast_demod_frames_vw8(const ref_receiver *rx, const void *pcm, int is_float, size_t nframes, float *mag_up,
                               uint32_t *idx_up, float *mag_down, uint32_t *idx_down, int nthreads);

int ref_fast_simd_width(void) {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f")) return 16;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return 8;
    return 0;
}

int ref_fast_demod_frames(const ref_receiver *rx, const void *pcm, int is_float, size_t nframes, float *mag_up,
                          uint32_t *idx_up, float *mag_down, uint32_t *idx_down, int nthreads) {
    if (!rx || rx->n != 2048) return -1;
    switch (ref_fast_simd_width()) {
    case 16: ref_fast_demod_frames_vw16(rx, pcm, is_float, nframes, mag_up, idx_up, mag_down, idx_down, nthreads); return 0;
    case 8: ref_fast_demod_frames_vw8(rx, pcm, is_float, nframes, mag_up, idx_up, mag_down, idx_down, nthreads); return 0;
    default: break;
    }
    /* no AVX2: the plain form */
    if (is_float) ref_demod_frames_f32(rx, (const float *) pcm, nframes, mag_up, idx_up, mag_down, idx_down, nthreads);
    else ref_demod_frames_i32(rx, (const int32_t *) pcm, nframes, mag_up, idx_up, mag_down, idx_down, nthreads);
    return 0;
}
