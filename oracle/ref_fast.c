/*
 * ref_fast.c — CPU ORACLE, tuned form (test infrastructure, NOT product code).
 *
 * The receiver chain of ref_dsp.c (ref_demod_frames_*: cast -> x up|down chirp -> x Hann -> 2048-point RFFT ->
 * magnitude -> arg-max over [0, bandwidth2), receiver/Src/main.c:163-215, receiver/Src/chirp.c:47-53) restated
 * for speed on the host: the honest CPU arm of bench.py (`cpu_baseline`, `--impl reference`).
 *
 *   - SAME canonical arithmetic, operation by operation: every lane of a SIMD register is
 *     one frame, so each vector instruction is VW independent scalar IEEE operations in the order ref_dsp.c
 *     performs them; results are bit-identical to ref_dsp.c (tests/test_oracle_fast.py) — the simple
 *     recursive form stays as the checker of this one;
 *   - iterative [32,32] plan, no allocation inside the frame loop (per-thread scratch), the PCM block is
 *     transposed once (frame-major -> sample-major) and shared by both hypotheses, the last pass only
 *     produces the bins the window reads;
 *   - OpenMP over blocks of VW frames.
 *
 * Compiled twice by oracle/Makefile: -DREF_FAST_VW=16 -mavx512f and -DREF_FAST_VW=8 -mavx2 -mfma;
 * ref_fast_dispatch.c picks at run time.  -ffp-contract=off as everywhere in oracle/.
 */
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ref_dsp.h"

#ifndef REF_FAST_VW
#define REF_FAST_VW 16
#endif

#if REF_FAST_VW == 16
typedef __m512 V;
#define VSET1(x) _mm512_set1_ps(x)
#define VADD(a, b) _mm512_add_ps(a, b)
#define VSUB(a, b) _mm512_sub_ps(a, b)
#define VMUL(a, b) _mm512_mul_ps(a, b)
#define VFMA(a, b, c) _mm512_fmadd_ps(a, b, c)    /*  a*b + c, one rounding */
#define VFNMA(a, b, c) _mm512_fnmadd_ps(a, b, c)  /* -a*b + c == fma(-a, b, c) */
#define VFMS(a, b, c) _mm512_fmsub_ps(a, b, c)    /*  a*b - c == fma(a, b, -c) */
#define VSQRT(a) _mm512_sqrt_ps(a)
#define VLOAD(p) _mm512_load_ps((const float *) (p))
#define VSTORE(p, v) _mm512_store_ps((float *) (p), v)
#define FN(name) name##_vw16
#else
typedef __m256 V;
#define VSET1(x) _mm256_set1_ps(x)
#define VADD(a, b) _mm256_add_ps(a, b)
#define VSUB(a, b) _mm256_sub_ps(a, b)
#define VMUL(a, b) _mm256_mul_ps(a, b)
#define VFMA(a, b, c) _mm256_fmadd_ps(a, b, c)
#define VFNMA(a, b, c) _mm256_fnmadd_ps(a, b, c)
#define VFMS(a, b, c) _mm256_fmsub_ps(a, b, c)
#define VSQRT(a) _mm256_sqrt_ps(a)
#define VLOAD(p) _mm256_load_ps((const float *) (p))
#define VSTORE(p, v) _mm256_store_ps((float *) (p), v)
#define FN(name) name##_vw8
#endif
#define VW REF_FAST_VW

/* ---- frame-major block -> sample-major vectors: xs[i] = { (float) frame_l[i] : l < VW } ------------------- */
#if REF_FAST_VW == 16
/* 16 x 16 transpose of 32-bit words: r[l] = 16 consecutive samples of frame l  ->  r[i] = sample i of the 16 frames */
static inline __attribute__((always_inline)) void transpose16(__m512 r[16]) {
    __m512 t[16], u[16];
    for (int i = 0; i < 16; i += 2) {
        t[i] = _mm512_unpacklo_ps(r[i], r[i + 1]);
        t[i + 1] = _mm512_unpackhi_ps(r[i], r[i + 1]);
    }
    for (int i = 0; i < 16; i += 4) {           /* u[i + q]: columns q, q+4, q+8, q+12 of rows i..i+3, one per 128-bit lane */
        u[i] = _mm512_castpd_ps(_mm512_unpacklo_pd(_mm512_castps_pd(t[i]), _mm512_castps_pd(t[i + 2])));
        u[i + 1] = _mm512_castpd_ps(_mm512_unpackhi_pd(_mm512_castps_pd(t[i]), _mm512_castps_pd(t[i + 2])));
        u[i + 2] = _mm512_castpd_ps(_mm512_unpacklo_pd(_mm512_castps_pd(t[i + 1]), _mm512_castps_pd(t[i + 3])));
        u[i + 3] = _mm512_castpd_ps(_mm512_unpackhi_pd(_mm512_castps_pd(t[i + 1]), _mm512_castps_pd(t[i + 3])));
    }
    for (int q = 0; q < 4; ++q) {
        const __m512 v0 = _mm512_shuffle_f32x4(u[q], u[q + 4], 0x88), v1 = _mm512_shuffle_f32x4(u[q], u[q + 4], 0xdd);
        const __m512 w0 = _mm512_shuffle_f32x4(u[q + 8], u[q + 12], 0x88), w1 = _mm512_shuffle_f32x4(u[q + 8], u[q + 12], 0xdd);
        r[q] = _mm512_shuffle_f32x4(v0, w0, 0x88);
        r[q + 8] = _mm512_shuffle_f32x4(v0, w0, 0xdd);
        r[q + 4] = _mm512_shuffle_f32x4(v1, w1, 0x88);
        r[q + 12] = _mm512_shuffle_f32x4(v1, w1, 0xdd);
    }
}
#endif

static void load_block(const void *pcm, int is_float, const size_t *frame_idx, uint32_t n, float *xs /* n x VW */) {
    /* exact (float) cast per sample (receiver/Src/main.c:663-665), lane l = frame frame_idx[l] */
#if REF_FAST_VW == 16
    for (uint32_t i0 = 0; i0 < n; i0 += 16) {
        __m512 r[16];
        if (is_float)
            for (int l = 0; l < 16; ++l) r[l] = _mm512_loadu_ps((const float *) pcm + frame_idx[l] * n + i0);
        else
            for (int l = 0; l < 16; ++l)
                r[l] = _mm512_cvtepi32_ps(_mm512_loadu_si512((const int32_t *) pcm + frame_idx[l] * n + i0));   /* round to nearest even, as the C cast */
        transpose16(r);
        for (int i = 0; i < 16; ++i) _mm512_store_ps(xs + (size_t) (i0 + i) * 16, r[i]);
    }
#else
    if (is_float) {
        const float *p = (const float *) pcm;
        for (int l = 0; l < VW; ++l) {
            const float *f = p + frame_idx[l] * n;
            for (uint32_t i = 0; i < n; ++i) xs[(size_t) i * VW + l] = f[i];
        }
    } else {
        const int32_t *p = (const int32_t *) pcm;
        for (int l = 0; l < VW; ++l) {
            const int32_t *f = p + frame_idx[l] * n;
            for (uint32_t i = 0; i < n; ++i) xs[(size_t) i * VW + l] = (float) f[i];
        }
    }
#endif
}

/* ---- 32-point base kernel: radix-2 DIT exactly as base_fft() in ref_dsp.c (and fft_base<32> on the device) ---- */
static const int kBrev5[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                               1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

static inline __attribute__((always_inline)) void base32(V *tr, V *ti, const float *w32 /* 16 x (re, im) */) {
#pragma GCC unroll 5
    for (int h = 1; h < 32; h <<= 1) {
#pragma GCC unroll 16
        for (int blk = 0; blk < 32; blk += 2 * h) {
#pragma GCC unroll 16
            for (int j = 0; j < h; ++j) {
                const int a = blk + j, b = blk + j + h;
                V er = tr[a], ei = ti[a], or_ = tr[b], oi = ti[b], sr, si, dr, di;
                if (j == 0) {
                    sr = VADD(er, or_); si = VADD(ei, oi);
                    dr = VSUB(er, or_); di = VSUB(ei, oi);
                } else if (2 * j == h) {
                    sr = VADD(er, oi); si = VSUB(ei, or_);
                    dr = VSUB(er, oi); di = VADD(ei, or_);
                } else {
                    const V wr = VSET1(w32[2 * (j * (16 / h))]), wi = VSET1(w32[2 * (j * (16 / h)) + 1]);
                    sr = VFMA(or_, wr, VFNMA(oi, wi, er));
                    si = VFMA(or_, wi, VFMA(oi, wr, ei));
                    const V two = VSET1(2.0f);
                    dr = VFMS(two, er, sr);
                    di = VFMS(two, ei, si);
                }
                tr[a] = sr; ti[a] = si; tr[b] = dr; ti[b] = di;
            }
        }
    }
}

typedef struct {
    float *xs;        /* n x VW: the block's samples, sample-major */
    float *vr, *vi;   /* 1024 x VW: pass-1 output, [d][a] */
    float *zr, *zi;   /* 1024 x VW: Z[k] for the bins the window needs */
} scratch_t;

/* one hypothesis of a block: front end + 1024-point complex FFT (plan [32,32]) + split + magnitude + arg-max */
static void hypothesis_block(const ref_receiver *rx, const float *chirp, const scratch_t *s, const float *w32,
                             float *mag_out /* VW */, uint32_t *idx_out /* VW */) {
    const float *tw = rx->S.cplx.tw;                  /* master table W_2048: (cos, -sin)(2 pi j / 2048) */
    const float *hann = rx->hann;
    const uint32_t bw2 = rx->bandwidth2;
    const int nb = (int) ((bw2 + 31) / 32);           /* pass-2 outputs c < nb and c > 31 - nb are needed (nb <= 16) */
    V tr[32], ti[32];
    /* pass 1: for every a, the 32-point FFT over b of z[a + 32 b], then x W_1024^(a d) for d != 0 */
    for (int a = 0; a < 32; ++a) {
        for (int i = 0; i < 32; ++i) {
            const int m = a + 32 * kBrev5[i];
            const V x0 = VLOAD(s->xs + (size_t) (2 * m) * VW), x1 = VLOAD(s->xs + (size_t) (2 * m + 1) * VW);
            /* (x * chirp) * hann: two roundings, the reference's order (chirp.c:47-53, main.c:171) */
            tr[i] = VMUL(VMUL(x0, VSET1(chirp[2 * m])), VSET1(hann[2 * m]));
            ti[i] = VMUL(VMUL(x1, VSET1(chirp[2 * m + 1])), VSET1(hann[2 * m + 1]));
        }
        base32(tr, ti, w32);
        for (int d = 0; d < 32; ++d) {
            V xr = tr[d], xi = ti[d];
            if (d != 0) {
                const size_t j = (size_t) a * d * 2;             /* W_1024^(a d) = W_2048^(2 a d) */
                const V br = VSET1(tw[2 * j]), bi = VSET1(tw[2 * j + 1]);
                const V t0 = VMUL(xi, bi), t1 = VMUL(xi, br);    /* cmul(): re = fma(ar,br,-(ai*bi)), im = fma(ar,bi,ai*br) */
                const V re = VFMS(xr, br, t0), im = VFMA(xr, bi, t1);
                xr = re; xi = im;
            }
            VSTORE(s->vr + ((size_t) d * 32 + a) * VW, xr);
            VSTORE(s->vi + ((size_t) d * 32 + a) * VW, xi);
        }
    }
    /* pass 2: for every d, the 32-point FFT over a; Z[d + 32 c] */
    for (int d = 0; d < 32; ++d) {
        for (int i = 0; i < 32; ++i) {
            tr[i] = VLOAD(s->vr + ((size_t) d * 32 + kBrev5[i]) * VW);
            ti[i] = VLOAD(s->vi + ((size_t) d * 32 + kBrev5[i]) * VW);
        }
        base32(tr, ti, w32);
        for (int c = 0; c < 32; ++c) {
            if (c >= nb && c < 32 - nb) continue;
            VSTORE(s->zr + ((size_t) d + 32 * (size_t) c) * VW, tr[c]);
            VSTORE(s->zi + ((size_t) d + 32 * (size_t) c) * VW, ti[c]);
        }
    }
    /* split (rfft_forward of ref_dsp.c), magnitude, arm_max_f32 over [0, bw2): ascending k, strict '<' update */
    V best;
    __attribute__((aligned(64))) float bestf[VW], magf[VW];
    uint32_t besti[VW];
    {
        const V z0r = VLOAD(s->zr), z0i = VLOAD(s->zi);
        const V x0 = VADD(z0r, z0i), xn = VSUB(z0r, z0i);        /* packed bin 0 = (X[0], X[N/2]) */
        best = VSQRT(VFMA(x0, x0, VMUL(xn, xn)));
        VSTORE(bestf, best);
        for (int l = 0; l < VW; ++l) besti[l] = 0;
    }
    const V half = VSET1(0.5f);
    for (uint32_t k = 1; k < bw2; ++k) {
        const V zkr = VLOAD(s->zr + (size_t) k * VW), zki = VLOAD(s->zi + (size_t) k * VW);
        const V zcr = VLOAD(s->zr + (size_t) (1024 - k) * VW), zci = VLOAD(s->zi + (size_t) (1024 - k) * VW);
        const V pr = VADD(zkr, zcr), pi = VSUB(zki, zci);
        const V qr = VADD(zki, zci), qi = VSUB(zcr, zkr);
        const V cr = VSET1(tw[2 * k]), si = VSET1(-tw[2 * k + 1]);
        const V xr = VMUL(half, VFMA(qr, cr, VFMA(qi, si, pr)));
        const V xi = VMUL(half, VFMA(qi, cr, VFNMA(qr, si, pi)));
        const V m = VSQRT(VFMA(xr, xr, VMUL(xi, xi)));
        VSTORE(magf, m);
        for (int l = 0; l < VW; ++l)
            if (bestf[l] < magf[l]) { bestf[l] = magf[l]; besti[l] = k; }
    }
    for (int l = 0; l < VW; ++l) { mag_out[l] = bestf[l]; idx_out[l] = besti[l]; }
}

/* receiver chain over nframes aligned frames, both hypotheses (same results as ref_demod_frames_i32/_f32) */
void FN(ref_fast_demod_frames)(const ref_receiver *rx, const void *pcm, int is_float, size_t nframes, float *mag_up,
                               uint32_t *idx_up, float *mag_down, uint32_t *idx_down, int nthreads) {
    if (rx->n != 2048 || !nframes) return;
    if (nthreads < 1) nthreads = 1;
    const uint32_t n = rx->n;
    float w32[32];
    for (int j = 0; j < 16; ++j) { w32[2 * j] = rx->S.cplx.tw[2 * (j * 64)]; w32[2 * j + 1] = rx->S.cplx.tw[2 * (j * 64) + 1]; }
    const long nblocks = (long) ((nframes + VW - 1) / VW);
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s;
        s.xs = (float *) aligned_alloc(64, sizeof(float) * (size_t) n * VW);
        s.vr = (float *) aligned_alloc(64, sizeof(float) * 1024 * VW);
        s.vi = (float *) aligned_alloc(64, sizeof(float) * 1024 * VW);
        s.zr = (float *) aligned_alloc(64, sizeof(float) * 1025 * VW);
        s.zi = (float *) aligned_alloc(64, sizeof(float) * 1025 * VW);
#pragma omp for schedule(dynamic, 4)
        for (long b = 0; b < nblocks; ++b) {
            size_t fi[VW];
            for (int l = 0; l < VW; ++l) {
                size_t f = (size_t) b * VW + l;
                fi[l] = f < nframes ? f : nframes - 1;           /* ragged tail: repeat the last frame, discard */
            }
            load_block(pcm, is_float, fi, n, s.xs);
            float m[VW];
            uint32_t ix[VW];
            for (int up = 1; up >= 0; --up) {
                hypothesis_block(rx, up ? rx->up_chirp : rx->down_chirp, &s, w32, m, ix);
                for (int l = 0; l < VW; ++l) {
                    size_t f = (size_t) b * VW + l;
                    if (f >= nframes) break;
                    if (up) { mag_up[f] = m[l]; idx_up[f] = ix[l]; }
                    else { mag_down[f] = m[l]; idx_down[f] = ix[l]; }
                }
            }
        }
        free(s.xs); free(s.vr); free(s.vi); free(s.zr); free(s.zi);
    }
}
