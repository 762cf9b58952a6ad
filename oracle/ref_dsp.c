/*
 * ref_dsp.c — CPU ORACLE (test infrastructure, NOT product code).  See ref_dsp.h.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -fPIC -shared  (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: the canonical arithmetic places every FMA explicitly.
 */
#include "ref_dsp.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FMA(a, b, c) __builtin_fmaf((a), (b), (c))

/* ------------------------------------------------------------------------------------------ */
/* arm_cos_f32 — arm_math.h:5685.  CMSIS-DSP V1.4.5 published algorithm: 512-entry sine table  */
/* (literals with 8 decimals in arm_common_tables.c) + linear interpolation.  Call sites:      */
/* receiver/Src/main.c:392 (Hann), experiments/chirp_compression_time_domain/Src/chirp.c:42,64 */
/* PINNED: reproduces all 24 device .flt captures (PCM x Hann) with 0 mismatches.              */
/* ------------------------------------------------------------------------------------------ */
#define FAST_MATH_TABLE_SIZE 512
static float g_sin_table[FAST_MATH_TABLE_SIZE + 1];
static int g_sin_table_ready = 0;

static void sin_table_init(void) {
    if (g_sin_table_ready) return;
    char buf[32];
    for (int i = 0; i <= FAST_MATH_TABLE_SIZE; ++i) {
        snprintf(buf, sizeof buf, "%.8f", sin(2.0 * M_PI * (double) i / (double) FAST_MATH_TABLE_SIZE));
        g_sin_table[i] = strtof(buf, NULL);
    }
    g_sin_table_ready = 1;
}

float32_t ref_arm_cos_f32(float32_t x) {
    sin_table_init();
    float in = x * 0.159154943092f + 0.25f;      /* x/(2*pi) + quarter turn */
    int32_t n = (int32_t) in;
    if (in < 0.0f) n--;
    in = in - (float) n;
    float findex = (float) FAST_MATH_TABLE_SIZE * in;
    uint16_t index = ((uint16_t) findex) & 0x1ff;
    float fract = findex - (float) index;
    float a = g_sin_table[index];
    float b = g_sin_table[index + 1];
    return (1.0f - fract) * a + fract * b;
}

/* arm_sin_cos_f32 — arm_math.h:4634-4637 (degrees in).  Call sites: receiver/Src/chirp.c:36,
 * experiments/synchronization/Src/chirp.c:37, experiments/iq_modulation/Src/iq_modem.c:43.
 * CMSIS-DSP V1.4.5's published algorithm: |theta|/360 -> fractional turns -> 512-entry sine table read at the
 * sine and the quarter-turn-shifted cosine index, cubic (Hermite) interpolation using the other function's table
 * values as derivatives (Dn = 2 pi/512), sine negated for negative theta.
 * PINNED to the reference's own binary: the operation sequence below was checked, instruction by instruction,
 * against the machine code of arm_sin_cos_f32 in receiver/Drivers/CMSIS/Lib/libarm_cortexM4lf_math.a (31 separate
 * VMUL/VADD/VSUB, no fused multiply-add; literals 0x3B360B61, 512.0f, 0x3C490FDB), and sinTable_f32 as stored in
 * that archive equals g_sin_table bit for bit (tools/cmsis_archive_probe.py, tests/golden/cmsis_archive.npz). */
void ref_arm_sin_cos_f32(float32_t theta_deg, float32_t *pSinVal, float32_t *pCosVal) {
    sin_table_init();
    float in = theta_deg * 0.00277777777778f;
    if (in < 0.0f) in = -in;
    in = in - (float) (int32_t) in;
    float findex = (float) FAST_MATH_TABLE_SIZE * in;
    uint16_t indexS = ((uint16_t) findex) & 0x1ff;
    uint16_t indexC = (uint16_t) ((indexS + (FAST_MATH_TABLE_SIZE / 4)) & 0x1ff);
    float fract = findex - (float) indexS;
    const float Dn = 0.0122718463030f;
    /* cosine: values from the cosine index, derivatives = -sine */
    float f1 = g_sin_table[indexC], f2 = g_sin_table[indexC + 1];
    float d1 = -g_sin_table[indexS], d2 = -g_sin_table[indexS + 1];
    float Df = f2 - f1;
    float temp = Dn * (d1 + d2) - 2 * Df;
    temp = fract * temp + (3 * Df - (d2 + 2 * d1) * Dn);
    temp = fract * temp + d1 * Dn;
    *pCosVal = fract * temp + f1;
    /* sine: values from the sine index, derivatives = cosine */
    f1 = g_sin_table[indexS]; f2 = g_sin_table[indexS + 1];
    d1 = g_sin_table[indexC]; d2 = g_sin_table[indexC + 1];
    Df = f2 - f1;
    temp = Dn * (d1 + d2) - 2 * Df;
    temp = fract * temp + (3 * Df - (d2 + 2 * d1) * Dn);
    temp = fract * temp + d1 * Dn;
    *pSinVal = fract * temp + f1;
    if (theta_deg < 0.0f) *pSinVal = -*pSinVal;
}

/* ------------------------------------------------------------------------------------------ */
/* element-wise primitives                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* arm_mult_f32 — arm_math.h:1938-1942 */
void ref_arm_mult_f32(const float32_t *a, const float32_t *b, float32_t *dst, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) dst[i] = a[i] * b[i];
}
/* arm_scale_f32 — arm_math.h:2508 */
void ref_arm_scale_f32(const float32_t *src, float32_t scale, float32_t *dst, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) dst[i] = src[i] * scale;
}
/* arm_copy_f32 — arm_math.h:2819 */
void ref_arm_copy_f32(const float32_t *src, float32_t *dst, uint32_t n) {
    memmove(dst, src, (size_t) n * sizeof(float));
}
/* arm_mean_f32 — arm_math.h:6192.  Sequential left-to-right sum, then one division. */
void ref_arm_mean_f32(const float32_t *src, uint32_t n, float32_t *result) {
    float sum = 0.0f;
    for (uint32_t i = 0; i < n; ++i) sum = sum + src[i];
    *result = sum / (float) n;
}
/* arm_max_f32 — arm_math.h:6537-6541.  First occurrence of the maximum (strict '<' update). */
void ref_arm_max_f32(const float32_t *src, uint32_t n, float32_t *result, uint32_t *index) {
    float out = src[0];
    uint32_t oi = 0;
    for (uint32_t i = 1; i < n; ++i)
        if (out < src[i]) { out = src[i]; oi = i; }
    *result = out;
    *index = oi;
}
/* arm_cmplx_mult_cmplx_f32 — arm_math.h:6579-6583.  (a+jb)(c+jd), no conjugate.
 * canonical: re = fma(a,c,-(b*d)); im = fma(a,d,b*c). */
static inline void cmul(float ar, float ai, float br, float bi, float *re, float *im) {
    float t0 = ai * bi;
    float t1 = ai * br;
    *re = FMA(ar, br, -t0);
    *im = FMA(ar, bi, t1);
}
void ref_arm_cmplx_mult_cmplx_f32(const float32_t *a, const float32_t *b, float32_t *dst, uint32_t ncplx) {
    for (uint32_t i = 0; i < ncplx; ++i) {
        float re, im;
        cmul(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], &re, &im);
        dst[2 * i] = re;
        dst[2 * i + 1] = im;
    }
}
/* arm_cmplx_mult_real_f32 — arm_math.h:6425-6429 */
void ref_arm_cmplx_mult_real_f32(const float32_t *cplx, const float32_t *real, float32_t *dst, uint32_t ncplx) {
    for (uint32_t i = 0; i < ncplx; ++i) {
        float r = real[i];
        float re = cplx[2 * i] * r, im = cplx[2 * i + 1] * r;
        dst[2 * i] = re;
        dst[2 * i + 1] = im;
    }
}
/* arm_cmplx_mag_f32 — arm_math.h:6312-6315.  canonical: sqrt(fma(re,re,im*im)), IEEE sqrt. */
static inline float cmag(float re, float im) { return sqrtf(FMA(re, re, im * im)); }
void ref_arm_cmplx_mag_f32(const float32_t *src, float32_t *dst, uint32_t ncplx) {
    for (uint32_t i = 0; i < ncplx; ++i) dst[i] = cmag(src[2 * i], src[2 * i + 1]);
}

/* ------------------------------------------------------------------------------------------ */
/* canonical FFT                                                              */
/* ------------------------------------------------------------------------------------------ */
void ref_twiddle(uint32_t j, uint32_t n, float *re, float *im) {
    double a = 2.0 * M_PI * (double) j / (double) n;
    *re = (float) cos(a);
    *im = (float) -sin(a);
}

/* Radix rule: peel 32s from the right while the remainder exceeds 32; the remainder goes first.
 * 1024 -> [32,32]; 2048 -> [2,32,32]; 512 -> [16,32]; 32768 -> [32,32,32]. */
uint32_t ref_fft_radices(uint32_t n, uint32_t *rad) {
    if (n < 2 || (n & (n - 1))) return 0;
    uint32_t tmp[REF_MAX_RADICES], cnt = 0;
    while (n > 32) { tmp[cnt++] = 32; n /= 32; if (cnt >= REF_MAX_RADICES - 1) return 0; }
    tmp[cnt++] = n;
    for (uint32_t i = 0; i < cnt; ++i) rad[i] = tmp[cnt - 1 - i];
    return cnt;
}

static int plan_init(ref_fft_plan *p, uint32_t n, uint32_t tw_n) {
    p->n = n;
    p->nrad = ref_fft_radices(n, p->rad);
    if (!p->nrad) return -1;
    p->tw_n = tw_n;
    p->tw = (float *) malloc(sizeof(float) * 2 * tw_n);
    if (!p->tw) return -1;
    for (uint32_t j = 0; j < tw_n; ++j) ref_twiddle(j, tw_n, &p->tw[2 * j], &p->tw[2 * j + 1]);
    return 0;
}

typedef struct { float r, i; } cf;

/* Base kernel: radix-2 decimation-in-time recursion on r <= 32 points, forward transform.
 *   j == 0   : s = E + O,           d = E - O
 *   j == r/4 : t = -j*O,            s = E + t, d = E - t
 *   otherwise: s = E + w*O via two chained FMAs per component, d = 2E - s (one FMA). */
static void base_fft(const ref_fft_plan *p, uint32_t r, const cf *in, uint32_t stride, cf *out) {
    if (r == 1) { out[0] = in[0]; return; }
    cf e[16], o[16];
    uint32_t h = r / 2;
    base_fft(p, h, in, stride * 2, e);
    base_fft(p, h, in + stride, stride * 2, o);
    uint32_t step = p->tw_n / r;
    for (uint32_t j = 0; j < h; ++j) {
        cf E = e[j], O = o[j], s, d;
        if (j == 0) {
            s.r = E.r + O.r; s.i = E.i + O.i;
            d.r = E.r - O.r; d.i = E.i - O.i;
        } else if (4 * j == r) {
            s.r = E.r + O.i; s.i = E.i - O.r;
            d.r = E.r - O.i; d.i = E.i + O.r;
        } else {
            float wr = p->tw[2 * j * step], wi = p->tw[2 * j * step + 1];
            s.r = FMA(O.r, wr, FMA(-O.i, wi, E.r));
            s.i = FMA(O.r, wi, FMA(O.i, wr, E.i));
            d.r = FMA(2.0f, E.r, -s.r);
            d.i = FMA(2.0f, E.i, -s.i);
        }
        out[j] = s;
        out[j + h] = d;
    }
}

/* Mixed-radix step.  n = A*B with B = rad[0]:  m = a + A*b,  k = B*c + d.
 *   1. V_a[d]  = base_fft_B over b of in[a + A*b]
 *   2. V_a[d] *= W_n^(a*d)  for d != 0  (d == 0 is skipped; a == 0 multiplies by W^0 = (1,-0))
 *   3. out[B*c + d] = fft_A over a of V_.[d]      (recursive with rad[1..]) */
static void fft_rec(const ref_fft_plan *p, uint32_t n, const uint32_t *rad, uint32_t nrad,
                    const cf *in, uint32_t stride, cf *out, cf *scratch) {
    if (nrad == 1) { base_fft(p, n, in, stride, out); return; }
    uint32_t B = rad[0], A = n / B;
    cf *tmp = scratch;                 /* B rows of A */
    cf *next = scratch + n;
    uint32_t step = p->tw_n / n;
    cf v[32];
    for (uint32_t a = 0; a < A; ++a) {
        base_fft(p, B, in + (size_t) a * stride, stride * A, v);
        for (uint32_t d = 0; d < B; ++d) {
            cf x = v[d];
            if (d != 0) {
                size_t j = (size_t) a * d * step;
                float re, im;
                cmul(x.r, x.i, p->tw[2 * j], p->tw[2 * j + 1], &re, &im);
                x.r = re; x.i = im;
            }
            tmp[(size_t) d * A + a] = x;
        }
    }
    cf *y = next;                      /* A outputs */
    for (uint32_t d = 0; d < B; ++d) {
        fft_rec(p, A, rad + 1, nrad - 1, tmp + (size_t) d * A, 1, y, next + A);
        for (uint32_t c = 0; c < A; ++c) out[(size_t) B * c + d] = y[c];
    }
}

/* forward, unscaled, natural order, out of place (in may not alias out) */
static void fft_forward(const ref_fft_plan *p, const cf *in, cf *out) {
    cf *scratch = (cf *) malloc(sizeof(cf) * (size_t) p->n * 4 + 64);
    fft_rec(p, p->n, p->rad, p->nrad, in, 1, out, scratch);
    free(scratch);
}

static inline void swap_ri(cf *x, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) { float t = x[i].r; x[i].r = x[i].i; x[i].i = t; }
}

/* arm_cfft_f32 — arm_math.h:2149-2153; consts arm_const_structs.h:49-57.  In place, interleaved.
 * Forward unscaled; inverse = swap(re,im) -> forward -> swap, scaled by 1/N (exact power of two).
 * bitReverseFlag must be 1 (natural-order output), the only mode the reference uses
 * (experiments/synchronization/Src/main.c:153, experiments/iq_modulation/Src/main.c:129). */
ref_status ref_arm_cfft_init_f32(ref_cfft_instance_f32 *S, uint32_t fftLen) {
    if (fftLen < 16 || fftLen > 65536 || (fftLen & (fftLen - 1))) return REF_MATH_ARGUMENT_ERROR;
    S->fftLen = (uint16_t) fftLen;
    return plan_init(&S->plan, fftLen, fftLen) ? REF_MATH_ARGUMENT_ERROR : REF_MATH_SUCCESS;
}
void ref_arm_cfft_free(ref_cfft_instance_f32 *S) { free(S->plan.tw); S->plan.tw = NULL; }

void ref_arm_cfft_f32(const ref_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag) {
    (void) bitReverseFlag;
    uint32_t n = S->plan.n;
    cf *in = (cf *) malloc(sizeof(cf) * n), *out = (cf *) malloc(sizeof(cf) * n);
    memcpy(in, p1, sizeof(cf) * n);
    if (ifftFlag) swap_ri(in, n);
    fft_forward(&S->plan, in, out);
    if (ifftFlag) {
        swap_ri(out, n);
        float sc = 1.0f / (float) n;
        for (uint32_t i = 0; i < n; ++i) { out[i].r *= sc; out[i].i *= sc; }
    }
    memcpy(p1, out, sizeof(cf) * n);
    free(in); free(out);
}

/* arm_rfft_fast_init_f32 / arm_rfft_fast_f32 — arm_math.h:2242-2249.
 * Forward output is packed [X0.re, X(N/2).re, X1.re, X1.im, ...] (N floats); inverse takes the
 * same packing and is normalised so irfft(rfft(x)) == x.  The reference's in-place aliasing
 * (hazard H2, experiments/chirp_compression_time_domain/Src/chirp.c:72-73,80,82) is DEFINED as the
 * mathematically correct result: we read everything before writing. */
ref_status ref_arm_rfft_fast_init_f32(ref_rfft_fast_instance_f32 *S, uint32_t fftLen) {
    if (fftLen < 32 || fftLen > 131072 || (fftLen & (fftLen - 1))) return REF_MATH_ARGUMENT_ERROR;
    S->fftLenRFFT = (uint16_t) fftLen;
    return plan_init(&S->cplx, fftLen / 2, fftLen) ? REF_MATH_ARGUMENT_ERROR : REF_MATH_SUCCESS;
}
void ref_arm_rfft_fast_free(ref_rfft_fast_instance_f32 *S) { free(S->cplx.tw); S->cplx.tw = NULL; }

static void rfft_forward(const ref_fft_plan *p, const float *a, float *out) {
    uint32_t h = p->n;                         /* N/2 */
    cf *Z = (cf *) malloc(sizeof(cf) * h);
    fft_forward(p, (const cf *) a, Z);         /* z[m] = a[2m] + j a[2m+1] */
    float x0 = Z[0].r + Z[0].i, xn = Z[0].r - Z[0].i;
    for (uint32_t k = 1; k < h; ++k) {
        cf Zk = Z[k], Zc = Z[h - k];
        float pr = Zk.r + Zc.r, pi = Zk.i - Zc.i;
        float qr = Zk.i + Zc.i, qi = Zc.r - Zk.r;
        float cr = p->tw[2 * k], si = -p->tw[2 * k + 1];       /* W_N^k = cr - j si */
        float xr = FMA(qr, cr, FMA(qi, si, pr));
        float xi = FMA(qi, cr, FMA(-qr, si, pi));
        out[2 * k] = 0.5f * xr;
        out[2 * k + 1] = 0.5f * xi;
    }
    out[0] = x0;
    out[1] = xn;
    free(Z);
}

static void rfft_inverse(const ref_fft_plan *p, const float *X, float *out) {
    uint32_t h = p->n;
    cf *Z = (cf *) malloc(sizeof(cf) * h), *z = (cf *) malloc(sizeof(cf) * h);
    /* 2Z[k] = (Xk + conj Xc) + j W_N^{-k} (Xk - conj Xc),  Xc = X[N/2-k] */
    Z[0].r = X[0] + X[1];
    Z[0].i = X[0] - X[1];
    for (uint32_t k = 1; k < h; ++k) {
        float xkr = X[2 * k], xki = X[2 * k + 1], xcr = X[2 * (h - k)], xci = X[2 * (h - k) + 1];
        float pr = xkr + xcr, pi = xki - xci;
        float qr = xkr - xcr, qi = xki + xci;
        float cr = p->tw[2 * k], si = -p->tw[2 * k + 1];
        Z[k].r = FMA(-qi, cr, FMA(-qr, si, pr));
        Z[k].i = FMA(qr, cr, FMA(-qi, si, pi));
    }
    swap_ri(Z, h);
    fft_forward(p, Z, z);
    float sc = 1.0f / (float) (2 * h);           /* 0.5 (from 2Z) * 1/(N/2) = 1/N, exact */
    for (uint32_t m = 0; m < h; ++m) {            /* swap back: re <- z.i, im <- z.r */
        out[2 * m] = z[m].i * sc;
        out[2 * m + 1] = z[m].r * sc;
    }
    free(Z); free(z);
}

void ref_arm_rfft_fast_f32(const ref_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag) {
    uint32_t n = S->fftLenRFFT ? S->fftLenRFFT : 2 * S->cplx.n;
    n = 2 * S->cplx.n;
    float *tmp = (float *) malloc(sizeof(float) * n);
    if (ifftFlag) rfft_inverse(&S->cplx, p, tmp);
    else rfft_forward(&S->cplx, p, tmp);
    memcpy(pOut, tmp, sizeof(float) * n);
    free(tmp);
}

/* arm_fir_init_f32 / arm_fir_f32 — arm_math.h:1194-1214.  pState holds numTaps+blockSize-1 floats
 * (zeroed by init); pCoeffs are in time-reversed order {b[numTaps-1] .. b[0]}.
 * y[n] = sum_i state[n+i]*coeffs[i], accumulated left to right with one FMA per tap from acc=0.
 * UNPINNED.  Call sites: experiments/iq_modulation/Src/iq_modem.c:48-49, 63-64. */
void ref_arm_fir_init_f32(ref_fir_instance_f32 *S, uint16_t numTaps, const float32_t *pCoeffs,
                          float32_t *pState, uint32_t blockSize) {
    S->numTaps = numTaps;
    S->pCoeffs = pCoeffs;
    memset(pState, 0, sizeof(float) * (numTaps + blockSize - 1));
    S->pState = pState;
}
void ref_arm_fir_f32(const ref_fir_instance_f32 *S, const float32_t *pSrc, float32_t *pDst, uint32_t blockSize) {
    uint32_t T = S->numTaps;
    float *st = S->pState;
    memcpy(st + (T - 1), pSrc, sizeof(float) * blockSize);   /* pSrc may alias pDst: copy first */
    for (uint32_t n = 0; n < blockSize; ++n) {
        float acc = 0.0f;
        for (uint32_t i = 0; i < T; ++i) acc = FMA(st[n + i], S->pCoeffs[i], acc);
        pDst[n] = acc;
    }
    memmove(st, st + blockSize, sizeof(float) * (T - 1));
}

/* ------------------------------------------------------------------------------------------ */
/* tables                                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* Periodic: receiver/Src/main.c:99,390-393  (WINDOW_SCALE = 2.0f * M_PI / (float) NN, double expr)
 * Symmetric: experiments/chirp_compression_time_domain/Src/chirp.c:13,63-65
 *            (WINDOW_SCALE = 2.0f * PI / (float)(PCM_SAMPLES - 1), PI = 3.14159265358979f: float expr) */
void ref_hann_window(float32_t *w, uint32_t n, ref_hann_kind kind) {
    float scale;
    if (kind == REF_HANN_PERIODIC) scale = (float) (2.0f * M_PI / (float) n);
    else scale = 2.0f * 3.14159265358979f / (float) (n - 1);
    for (uint32_t i = 0; i < n; ++i) w[i] = 0.5f - 0.5f * ref_arm_cos_f32((float) i * scale);
}

void ref_generate_ref_chirp(ref_chirp_variant v, const ref_chirp_params *p, int up, float32_t *out) {
    float t = 0.0f;
    if (v == REF_CHIRP_R || v == REF_CHIRP_S) {
        /* receiver/Src/chirp.c:16-40, experiments/synchronization/Src/chirp.c:16-44.
         * F0,F1 are ints in the reference; freq/theta expressions are evaluated in double
         * (the literals 2.0 and 360.0 are double) and stored to float. */
        float delta_f = (float) (p->f1 - p->f0) / p->sweep_T;
        float delta_t = p->sweep_T / (p->sweep_T * p->fs);
        for (uint32_t n = 0; n < p->n; ++n) {
            float freq;
            if (up) freq = (float) ((double) p->f0 + (double) (delta_f * t) / 2.0);
            else freq = (float) ((double) p->f1 - (double) (delta_f * t) / 2.0);
            float theta = (float) (360.0 * (double) freq * (double) t + (double) p->phase);
            t = t + delta_t;
            float s, c;
            ref_arm_sin_cos_f32(theta, &s, &c);
            if (v == REF_CHIRP_R) out[n] = s * 1.0f;          /* chirp.c:37-38: sin overwrites cos */
            else { out[2 * n] = c * 1.0f; out[2 * n + 1] = s * 1.0f; }
        }
    } else {
        /* experiments/chirp_compression_time_domain/Src/chirp.c:25-45 (T, F1/F2 float literals) and
         * experiments/chirp_compression_freq_domain/Src/chirp.c:15-35 (F, int F1/F2, phase ignored). */
        float time_frame = (float) p->n / p->fs;
        float delta_f = (p->f1 - p->f0) / time_frame;
        float delta_t = time_frame / (time_frame * p->fs);
        for (uint32_t i = 0; i < p->n; ++i) {
            float freq = up ? p->f0 + delta_f * t : p->f1 - delta_f * t;
            float arg;
            if (v == REF_CHIRP_T) arg = (float) (2.0 * (double) 3.14159265358979f * (double) freq * (double) t + (double) p->phase);
            else arg = (float) (2.0 * (double) 3.14159265358979f * (double) freq * (double) t);
            t = t + delta_t;
            out[i] = ref_arm_cos_f32(arg) * 1.0f;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* receiver chain                                                                              */
/* ------------------------------------------------------------------------------------------ */
int ref_receiver_init(ref_receiver *rx, uint32_t n, float fs, float f0, float f1, float sweep_T) {
    memset(rx, 0, sizeof *rx);
    rx->n = n;
    rx->fs = fs;
    /* main.c:372-374: bandwidth = (F1 - F0) * NN / fs  (int * unsigned long -> float division -> uint32) */
    rx->bandwidth = (uint32_t) ((float) ((unsigned long) (int) (f1 - f0) * (unsigned long) n) / fs);
    rx->bandwidth2 = rx->bandwidth * 2;
    rx->idx_left_zero = n - rx->bandwidth2;
    rx->hann = (float *) malloc(sizeof(float) * n);
    rx->up_chirp = (float *) malloc(sizeof(float) * n);
    rx->down_chirp = (float *) malloc(sizeof(float) * n);
    if (!rx->hann || !rx->up_chirp || !rx->down_chirp) return -1;
    if (ref_arm_rfft_fast_init_f32(&rx->S, n) != REF_MATH_SUCCESS) return -1;      /* main.c:377 */
    ref_chirp_params cp = { n, fs, f0, f1, sweep_T, -90.0f };
    ref_generate_ref_chirp(REF_CHIRP_R, &cp, 1, rx->up_chirp);                     /* chirp.c:43 */
    ref_generate_ref_chirp(REF_CHIRP_R, &cp, 0, rx->down_chirp);                   /* chirp.c:44 */
    ref_hann_window(rx->hann, n, REF_HANN_PERIODIC);                               /* main.c:390-393 */
    return 0;
}
void ref_receiver_free(ref_receiver *rx) {
    free(rx->hann); free(rx->up_chirp); free(rx->down_chirp);
    ref_arm_rfft_fast_free(&rx->S);
    memset(rx, 0, sizeof *rx);
}

/* main.c:154-160: integer arithmetic with fs truncated to int32 */
int32_t ref_idx2freq(const ref_receiver *rx, uint32_t idx) {
    uint32_t nn = rx->n;
    if (idx < nn / 2) return (int32_t) ((uint32_t) (int32_t) rx->fs * idx / nn);
    return (int32_t) ((uint32_t) (int32_t) rx->fs * (nn - idx) / nn) * -1;
}

/* main.c:163-180 */
void ref_pipeline(const ref_receiver *rx, float32_t *signal, int up) {
    uint32_t n = rx->n;
    float *buf = (float *) malloc(sizeof(float) * n);
    ref_arm_mult_f32(signal, up ? rx->up_chirp : rx->down_chirp, signal, n);       /* chirp.c:47-53 */
    ref_arm_mult_f32(signal, rx->hann, signal, n);                                 /* main.c:171 */
    ref_arm_rfft_fast_f32(&rx->S, signal, buf, 0);                                 /* main.c:174 */
    ref_arm_copy_f32(buf, signal, n);                                              /* main.c:175 */
    /* main.c:178 asks for NN complex magnitudes = 2*NN floats, reading NN floats of uninitialised
     * stack beyond the spectrum (hazard H1).  DEFINED: those read as zero -> mags[n/2..n) = 0. */
    ref_arm_cmplx_mag_f32(signal, buf, n / 2);
    memcpy(signal, buf, sizeof(float) * (n / 2));
    memset(signal + n / 2, 0, sizeof(float) * (n / 2));
    free(buf);
}

/* main.c:183-231 */
void ref_dsp(const ref_receiver *rx, const float32_t *fifo, uint32_t sync_position,
             ref_history *h, float mag_mean, int up) {
    uint32_t n = rx->n;
    float *tf = (float *) malloc(sizeof(float) * n);
    for (uint32_t i = 0; i < n; ++i) tf[i] = fifo[sync_position + i];               /* main.c:196-198 */
    ref_pipeline(rx, tf, up);
    float ml, mr, mm;
    uint32_t il, ir, im;
    ref_arm_max_f32(&tf[rx->idx_left_zero], rx->bandwidth2, &ml, &il);             /* main.c:206 */
    ref_arm_max_f32(&tf[0], rx->bandwidth2, &mr, &ir);                             /* main.c:208 */
    if (ml > mr) { mm = ml; im = rx->idx_left_zero + il; }                         /* main.c:209-215 */
    else { mm = mr; im = ir; }
    h->mag_max = mm; h->mag_max_left = ml; h->mag_max_right = mr;
    h->max_idx = im; h->max_idx_left = rx->idx_left_zero + il; h->max_idx_right = ir;
    h->max_freq = ref_idx2freq(rx, im);
    h->max_freq_left = ref_idx2freq(rx, rx->idx_left_zero + il);
    h->max_freq_right = ref_idx2freq(rx, ir);
    h->mag_mean = mag_mean;
    h->snr = (mm - mag_mean) / mag_mean;                                           /* main.c:229 */
    free(tf);
}

static void demod_one(const ref_receiver *rx, const float *frame, float *fifo,
                      float *mu, uint32_t *iu, float *md, uint32_t *id) {
    /* aligned frame = dsp() at sync_position 0 of a fifo that starts with the frame */
    (void) fifo;
    ref_history h;
    ref_dsp(rx, frame, 0, &h, 1.0f, 1);
    *mu = h.mag_max; *iu = h.max_idx;
    ref_dsp(rx, frame, 0, &h, 1.0f, 0);
    *md = h.mag_max; *id = h.max_idx;
}

void ref_demod_frames_f32(const ref_receiver *rx, const float *pcm, size_t nframes,
                          float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down,
                          int nthreads) {
    uint32_t n = rx->n;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long f = 0; f < (long) nframes; ++f)
        demod_one(rx, pcm + (size_t) f * n, NULL, &mag_up[f], &idx_up[f], &mag_down[f], &idx_down[f]);
}

void ref_demod_frames_i32(const ref_receiver *rx, const int32_t *pcm, size_t nframes,
                          float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down,
                          int nthreads) {
    uint32_t n = rx->n;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long f = 0; f < (long) nframes; ++f) {
        float *fr = (float *) malloc(sizeof(float) * n);
        /* main.c:663-665: fifo_queue[...] = (float) buf[i] — the whole PCM scaling */
        for (uint32_t i = 0; i < n; ++i) fr[i] = (float) pcm[(size_t) f * n + i];
        demod_one(rx, fr, NULL, &mag_up[f], &idx_up[f], &mag_down[f], &idx_down[f]);
        free(fr);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* complex-FFT variant — experiments/synchronization/Src/main.c:135-213, Src/chirp.c:16-57      */
/* ------------------------------------------------------------------------------------------ */
int ref_sync_receiver_init(ref_sync_receiver *rx, uint32_t n, float fs, float f0, float f1, float sweep_T) {
    memset(rx, 0, sizeof *rx);
    rx->n = n; rx->fs = fs;
    rx->bandwidth = (uint32_t) ((float) ((unsigned long) (int) (f1 - f0) * (unsigned long) n) / fs);   /* main.c:353 */
    rx->bandwidth2 = rx->bandwidth * 2;
    rx->idx_left_zero = n - rx->bandwidth2;
    rx->hann = (float *) malloc(sizeof(float) * n);
    rx->up_chirp = (float *) malloc(sizeof(float) * 2 * n);
    rx->down_chirp = (float *) malloc(sizeof(float) * 2 * n);
    if (!rx->hann || !rx->up_chirp || !rx->down_chirp) return -1;
    if (ref_arm_cfft_init_f32(&rx->C, n) != REF_MATH_SUCCESS) return -1;
    ref_chirp_params cp = { n, fs, f0, f1, sweep_T, -90.0f };
    ref_generate_ref_chirp(REF_CHIRP_S, &cp, 1, rx->up_chirp);                     /* chirp.c:47 */
    ref_generate_ref_chirp(REF_CHIRP_S, &cp, 0, rx->down_chirp);                   /* chirp.c:48 */
    ref_hann_window(rx->hann, n, REF_HANN_PERIODIC);
    return 0;
}
void ref_sync_receiver_free(ref_sync_receiver *rx) {
    free(rx->hann); free(rx->up_chirp); free(rx->down_chirp);
    ref_arm_cfft_free(&rx->C);
    memset(rx, 0, sizeof *rx);
}
/* main.c:144-158 */
void ref_sync_pipeline(const ref_sync_receiver *rx, float32_t *pframe, int up) {
    uint32_t n = rx->n;
    ref_arm_cmplx_mult_cmplx_f32(pframe, up ? rx->up_chirp : rx->down_chirp, pframe, n);   /* chirp.c:51-57 */
    ref_arm_cmplx_mult_real_f32(pframe, rx->hann, pframe, n);                      /* main.c:150 */
    ref_arm_cfft_f32(&rx->C, pframe, 0, 1);                                        /* main.c:153 */
    float *mag = (float *) malloc(sizeof(float) * n);
    ref_arm_cmplx_mag_f32(pframe, mag, n);                                         /* main.c:156 */
    memcpy(pframe, mag, sizeof(float) * n);
    free(mag);
}
/* main.c:161-213 */
void ref_sync_dsp(const ref_sync_receiver *rx, const float32_t *fifo, uint32_t sync_position, ref_history *h,
                  float mag_mean, int up) {
    uint32_t n = rx->n;
    float *tf = (float *) malloc(sizeof(float) * 2 * n);
    for (uint32_t i = 0; i < n; ++i) { tf[2 * i] = fifo[sync_position + i]; tf[2 * i + 1] = 0.0f; }   /* main.c:175-180 */
    ref_sync_pipeline(rx, tf, up);
    float ml, mr, mm;
    uint32_t il, ir, im;
    ref_arm_max_f32(&tf[rx->idx_left_zero], rx->bandwidth2, &ml, &il);             /* main.c:188 */
    ref_arm_max_f32(&tf[0], rx->bandwidth2, &mr, &ir);                             /* main.c:190 */
    if (ml > mr) { mm = ml; im = rx->idx_left_zero + il; } else { mm = mr; im = ir; }
    ref_receiver tmp; memset(&tmp, 0, sizeof tmp); tmp.n = n; tmp.fs = rx->fs;     /* idx2freq is identical (main.c:135-141) */
    h->mag_max = mm; h->mag_max_left = ml; h->mag_max_right = mr;
    h->max_idx = im; h->max_idx_left = rx->idx_left_zero + il; h->max_idx_right = ir;
    h->max_freq = ref_idx2freq(&tmp, im);
    h->max_freq_left = ref_idx2freq(&tmp, rx->idx_left_zero + il);
    h->max_freq_right = ref_idx2freq(&tmp, ir);
    h->mag_mean = mag_mean;
    h->snr = (mm - mag_mean) / mag_mean;
    free(tf);
}

/* ------------------------------------------------------------------------------------------ */
/* I/Q baseband path — experiments/iq_modulation/Src/iq_modem.c, IQ_modulation.ipynb           */
/* ------------------------------------------------------------------------------------------ */
int ref_iq_init(ref_iq *q, uint32_t n, float fs, float carrier, float bw, float sweep_T, const float *taps,
                uint32_t num_taps, uint32_t window_bins) {
    memset(q, 0, sizeof *q);
    if (num_taps < 1 || num_taps > 64) return -1;
    q->n = n; q->fs = fs; q->window_bins = window_bins; q->num_taps = num_taps;
    memcpy(q->taps, taps, sizeof(float) * num_taps);
    q->carrier_cos = (float *) malloc(sizeof(float) * n);
    q->carrier_sin = (float *) malloc(sizeof(float) * n);
    q->chirp = (float *) malloc(sizeof(float) * n);
    q->chirp_conj = (float *) malloc(sizeof(float) * n);
    q->hann = (float *) malloc(sizeof(float) * (n / 2));
    if (ref_arm_cfft_init_f32(&q->C, n / 2) != REF_MATH_SUCCESS) return -1;
    /* iq_modem.c:34-46 */
    float delta_t = sweep_T / (sweep_T * fs), t = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float theta = (float) (360.0 * (double) carrier * (double) t);
        ref_arm_sin_cos_f32(theta, &q->carrier_sin[i], &q->carrier_cos[i]);
        t = t + delta_t;
    }
    /* baseband chirp -bw/2 .. +bw/2 over one frame at the decimated rate (IQ_modulation.ipynb cells 2-3,10) */
    ref_chirp_params cp = { n / 2, fs / 2.0f, -bw / 2.0f, bw / 2.0f, (float) n / fs, 0.0f };
    ref_generate_ref_chirp(REF_CHIRP_S, &cp, 1, q->chirp);
    for (uint32_t m = 0; m < n / 2; ++m) { q->chirp_conj[2 * m] = q->chirp[2 * m]; q->chirp_conj[2 * m + 1] = -q->chirp[2 * m + 1]; }
    ref_hann_window(q->hann, n / 2, REF_HANN_PERIODIC);
    return 0;
}
void ref_iq_free(ref_iq *q) {
    free(q->carrier_cos); free(q->carrier_sin); free(q->chirp); free(q->chirp_conj); free(q->hann);
    ref_arm_cfft_free(&q->C);
    memset(q, 0, sizeof *q);
}
void ref_iq_demod_i32(const ref_iq *q, const int32_t *pcm, uint32_t nframes, float *mag_up, uint32_t *idx_up,
                      float *mag_down, uint32_t *idx_down) {
    const uint32_t n = q->n, h = n / 2, W = q->window_bins;
    float *x = (float *) malloc(sizeof(float) * n), *I = (float *) malloc(sizeof(float) * n), *Q = (float *) malloc(sizeof(float) * n);
    float *R = (float *) malloc(sizeof(float) * n), *P = (float *) malloc(sizeof(float) * n), *mag = (float *) malloc(sizeof(float) * h);
    float *si = (float *) malloc(sizeof(float) * (q->num_taps + n)), *sq = (float *) malloc(sizeof(float) * (q->num_taps + n));
    ref_fir_instance_f32 SI, SQ;
    ref_arm_fir_init_f32(&SI, (uint16_t) q->num_taps, q->taps, si, n);             /* iq_modem.c:48-49 */
    ref_arm_fir_init_f32(&SQ, (uint16_t) q->num_taps, q->taps, sq, n);
    for (uint32_t t = 0; t < nframes; ++t) {
        for (uint32_t i = 0; i < n; ++i) x[i] = (float) pcm[(size_t) t * n + i];
        ref_arm_mult_f32(x, q->carrier_sin, Q, n);                                 /* iq_modem.c:59 */
        ref_arm_mult_f32(x, q->carrier_cos, I, n);                                 /* iq_modem.c:60 */
        ref_arm_fir_f32(&SI, I, I, n);                                             /* iq_modem.c:63 */
        ref_arm_fir_f32(&SQ, Q, Q, n);                                             /* iq_modem.c:64 */
        for (uint32_t m = 0; m < h; ++m) { R[2 * m] = I[2 * m]; R[2 * m + 1] = Q[2 * m]; }   /* R = I + jQ, every 2nd sample */
        for (int hyp = 0; hyp < 2; ++hyp) {
            /* up: R x conj(chirp) (notebook cell 29); down: R x chirp (cell 30) */
            ref_arm_cmplx_mult_cmplx_f32(R, hyp == 0 ? q->chirp_conj : q->chirp, P, h);
            ref_arm_cmplx_mult_real_f32(P, q->hann, P, h);
            ref_arm_cfft_f32(&q->C, P, 0, 1);
            ref_arm_cmplx_mag_f32(P, mag, h);
            float ml, mr; uint32_t il, ir;
            ref_arm_max_f32(&mag[h - W], W, &ml, &il);
            ref_arm_max_f32(&mag[0], W, &mr, &ir);
            float mm = mr; uint32_t im = ir;
            if (ml > mr) { mm = ml; im = h - W + il; }
            if (hyp == 0) { mag_up[t] = mm; idx_up[t] = im; } else { mag_down[t] = mm; idx_down[t] = im; }
        }
    }
    free(x); free(I); free(Q); free(R); free(P); free(mag); free(si); free(sq);
}

/* ------------------------------------------------------------------------------------------ */
/* receiver state machine — receiver/Src/main.c:233-273, 311-339, 417-580                      */
/* ------------------------------------------------------------------------------------------ */
void ref_rx_state_init(ref_rx_state *st, const ref_receiver *rx) {
    memset(st, 0, sizeof *st);
    st->state = REF_IDLE;                                   /* main.c:111 */
    st->sync_position = rx->n / 2;                          /* main.c:330 */
    for (int i = 0; i < 12; ++i) st->mag_stat[i] = 1E37f;   /* main.c:321-322 */
    st->lock_frame = -1;
}

/* symbol_snr (main.c:233-236) with hazards H3/H5 defined: a probe outside [0, 2n] reads outside
 * fifo_queue in the reference; here it yields snr = -inf and leaves the history slot untouched. */
static float symbol_snr(const ref_receiver *rx, const float *fifo, int64_t pos, ref_history *h, int up) {
    if (pos < 0 || pos > (int64_t) 2 * rx->n) return -INFINITY;
    ref_dsp(rx, fifo, (uint32_t) pos, h, h->mag_mean, up);
    return h->snr;
}

/* resync (main.c:243-273) */
static void resync(const ref_receiver *rx, const float *fifo, float snr, ref_history *hist, uint32_t offset,
                   uint32_t *sync_position, int up) {
    int64_t l = (int64_t) *sync_position - offset, r = (int64_t) *sync_position + offset;
    float snr_l = symbol_snr(rx, fifo, l, &hist[2], up);
    float snr_r = symbol_snr(rx, fifo, r, &hist[3], up);
    if ((snr > snr_l) && (snr > snr_r)) {
        hist[2].rank = '-'; hist[3].rank = '-';
    } else if (snr_l >= snr_r) {
        if (l >= 0) *sync_position = (uint32_t) l;
    } else if (snr_l < snr_r) {
        if (r <= (int64_t) 2 * rx->n) *sync_position = (uint32_t) r;
    }
}

static void emit(uint8_t *out, uint32_t *nout, uint32_t cap, uint8_t c) {
    if (*nout < cap) out[*nout] = c;
    (*nout)++;
}

void ref_receiver_step(const ref_receiver *rx, ref_rx_state *st, float thr, const float *fifo,
                       uint8_t *out, uint32_t *nout, uint32_t cap) {
    const uint32_t n = rx->n, offset = n / 8, shift = n / 4;               /* main.c:406-407 */
    uint32_t prev = st->state;
    switch (st->state) {
    case REF_IDLE:
        st->sync_cnt = 0;
        ref_arm_mean_f32(&st->mag_stat[4], 8, &st->mag_mean);              /* main.c:431 */
        __attribute__((fallthrough));                                      /* as the firmware does, main.c:434 */
    case REF_SYNCHRONIZING:
        for (uint32_t i = 0; i < 4; ++i) {                                  /* main.c:447-451 */
            st->sync_position = n / 2 + st->turn * offset + shift * i;
            ref_dsp(rx, fifo, st->sync_position, &st->history[i * 2 + st->turn], st->mag_mean, 1);
        }
        st->turn = st->turn == 0 ? 1 : 0;
        if (st->turn == 1) {
            for (int i = 10; i >= 0; --i) st->mag_stat[i + 1] = st->mag_stat[i];   /* main.c:458-460 */
            float mag_max_max = 0.0f;
            for (uint32_t i = 0; i < 8; ++i) {                              /* main.c:463-471 */
                st->history[i].rank = '-';
                if (st->history[i].mag_max > mag_max_max) { mag_max_max = st->history[i].mag_max; st->max_idx = i; }
            }
            st->mag_stat[0] = mag_max_max;
            st->history[st->max_idx].rank = '+';
            float snr = (mag_max_max - st->mag_mean) / st->mag_mean;        /* main.c:477 */
            if (snr >= thr) {
                st->state = REF_SYNCHRONIZING;
                if (++st->sync_cnt >= 3) {
                    st->state = REF_SYNCHRONIZED;
                    st->sync_position = n / 2 + st->max_idx * offset;       /* main.c:483 */
                }
            } else {
                st->state = REF_IDLE;
            }
        }
        break;
    case REF_SYNCHRONIZED: {
        float up = symbol_snr(rx, fifo, st->sync_position, &st->history[0], 1);    /* main.c:493-494 */
        float down = symbol_snr(rx, fifo, st->sync_position, &st->history[1], 0);
        if (up >= thr || down >= thr) {
            if (down > up) {                                                /* the delimiter */
                resync(rx, fifo, down, st->history, offset, &st->sync_position, 0);
                st->state = REF_DATA_RECEIVING;
            } else {
                resync(rx, fifo, up, st->history, offset, &st->sync_position, 1);
            }
        } else {
            st->state = REF_IDLE;
        }
        break;
    }
    case REF_DATA_RECEIVING: {
        float up = symbol_snr(rx, fifo, st->sync_position, &st->history[0], 1);    /* main.c:518-519 */
        float down = symbol_snr(rx, fifo, st->sync_position, &st->history[1], 0);
        if (up >= thr || down >= thr) {
            if (down > up) {
                st->msg = ((st->msg << 1) + 0) & 0xffu;                     /* main.c:525 */
                resync(rx, fifo, down, st->history, offset, &st->sync_position, 0);
            } else {
                st->msg = ((st->msg << 1) + 1) & 0xffu;                     /* main.c:529 */
                resync(rx, fifo, up, st->history, offset, &st->sync_position, 1);
            }
            if (++st->msg_cnt >= 8) {                                       /* main.c:532-537 */
                emit(out, nout, cap, (uint8_t) st->msg);
                st->msg = 0; st->msg_cnt = 0;
            }
        } else {                                                            /* main.c:539-549 */
            emit(out, nout, cap, (uint8_t) '\n');
            st->state = REF_IDLE;
            st->msg = 0; st->msg_cnt = 0;
        }
        break;
    }
    }
    if (prev != REF_SYNCHRONIZED && st->state == REF_SYNCHRONIZED && st->lock_frame < 0) {
        st->lock_frame = (int32_t) st->frames_seen;
        st->lock_position = st->sync_position;
    }
    st->frames_seen++;
}

void ref_receiver_run_i32(const ref_receiver *rx, float thr, const int32_t *pcm, uint32_t nframes,
                          uint8_t *out, uint32_t *nout, uint32_t cap, ref_rx_state *final_state) {
    const uint32_t n = rx->n;
    float *fifo = (float *) calloc(3 * (size_t) n, sizeof(float));
    ref_rx_state st;
    ref_rx_state_init(&st, rx);
    *nout = 0;
    for (uint32_t t = 0; t < nframes; ++t) {
        memmove(fifo, fifo + n, sizeof(float) * 2 * n);                     /* main.c:662 */
        for (uint32_t i = 0; i < n; ++i) fifo[2 * n + i] = (float) pcm[(size_t) t * n + i];   /* main.c:663-665 */
        ref_receiver_step(rx, &st, thr, fifo, out, nout, cap);
    }
    if (final_state) *final_state = st;
    free(fifo);
}

void ref_sync_search_i32(const ref_receiver *rx, const int32_t *pcm, uint32_t nframes, uint32_t sync_add,
                         float *mag, uint32_t *idx) {
    const uint32_t n = rx->n, offset = n / 8, shift = n / 4;
    if (sync_add < 1) sync_add = 1;
    float *acc = (float *) malloc(sizeof(float) * 3 * n);
    for (uint32_t t = 0; t < nframes; ++t) {
        /* FIFO at frame t = frames t-2, t-1, t (zeros before the stream starts); synchronous
         * addition sums the FIFOs of frames t, t-1, ..., t-sync_add+1 sample by sample, oldest first */
        for (uint32_t i = 0; i < 3 * n; ++i) acc[i] = 0.0f;
        for (uint32_t j = sync_add; j-- > 0;) {
            for (uint32_t i = 0; i < 3 * n; ++i) {
                int64_t g = ((int64_t) t - (int64_t) j - 2) * n + i;
                float v = g >= 0 ? (float) pcm[g] : 0.0f;
                acc[i] = (j == sync_add - 1) ? v : acc[i] + v;
            }
        }
        ref_history h;
        for (uint32_t i = 0; i < 4; ++i) {
            uint32_t pos = n / 2 + (t & 1u) * offset + shift * i;
            ref_dsp(rx, acc, pos, &h, 1.0f, 1);
            mag[(size_t) t * 4 + i] = h.mag_max;
            idx[(size_t) t * 4 + i] = h.max_idx;
        }
    }
    free(acc);
}

/* ------------------------------------------------------------------------------------------ */
/* frequency-domain compression chain                                                          */
/* ------------------------------------------------------------------------------------------ */
/* experiments/chirp_compression_time_domain/Src/chirp.c:52-75 */
int ref_compressor_init(ref_compressor *c, uint32_t n, float fs, float f1, float f2) {
    memset(c, 0, sizeof *c);
    c->n = n; c->fs = fs;
    c->window = (float *) malloc(sizeof(float) * n);
    c->H_up = (float *) malloc(sizeof(float) * n);
    c->H_down = (float *) malloc(sizeof(float) * n);
    if (!c->window || !c->H_up || !c->H_down) return -1;
    if (ref_arm_rfft_fast_init_f32(&c->S, n) != REF_MATH_SUCCESS) return -1;       /* chirp.c:55 */
    ref_chirp_params cp = { n, fs, f1, f2, 0.0f, (float) (-3.14159265358979f / 2.0) };
    ref_generate_ref_chirp(REF_CHIRP_T, &cp, 1, c->H_up);                          /* chirp.c:58 */
    ref_generate_ref_chirp(REF_CHIRP_T, &cp, 0, c->H_down);                        /* chirp.c:59 */
    ref_hann_window(c->window, n, REF_HANN_SYMMETRIC);                             /* chirp.c:63-65 */
    ref_arm_mult_f32(c->H_up, c->window, c->H_up, n);                              /* chirp.c:68 */
    ref_arm_mult_f32(c->H_down, c->window, c->H_down, n);                          /* chirp.c:69 */
    ref_arm_rfft_fast_f32(&c->S, c->H_up, c->H_up, 0);                             /* chirp.c:72 */
    ref_arm_rfft_fast_f32(&c->S, c->H_down, c->H_down, 0);                         /* chirp.c:73 */
    return 0;
}
void ref_compressor_free(ref_compressor *c) {
    free(c->window); free(c->H_up); free(c->H_down);
    ref_arm_rfft_fast_free(&c->S);
    memset(c, 0, sizeof *c);
}
/* chirp.c:78-83 */
void ref_compress_chirp(const ref_compressor *c, float32_t *inout, int use_up) {
    uint32_t n = c->n;
    ref_arm_mult_f32(inout, c->window, inout, n);                                  /* windowing() */
    ref_arm_rfft_fast_f32(&c->S, inout, inout, 0);
    ref_arm_cmplx_mult_cmplx_f32(inout, use_up ? c->H_up : c->H_down, inout, n / 2);
    ref_arm_rfft_fast_f32(&c->S, inout, inout, 1);
}
/* experiments/chirp_compression_time_domain/Src/main.c:171-189 */
void ref_compress_frames_i32(const ref_compressor *c, const int32_t *pcm, size_t nframes, int use_up,
                             float *max_val, uint32_t *max_idx, int nthreads) {
    uint32_t n = c->n;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long f = 0; f < (long) nframes; ++f) {
        float *fr = (float *) malloc(sizeof(float) * n);
        for (uint32_t i = 0; i < n; ++i) fr[i] = (float) pcm[(size_t) f * n + i];
        ref_compress_chirp(c, fr, use_up);
        ref_arm_max_f32(fr, n, &max_val[f], &max_idx[f]);                          /* main.c:189 */
        free(fr);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* overlap-save frame synchroniser (twin of usc_correlate_os): compress_chirp's three steps        */
/* (experiments/chirp_compression_time_domain/Src/chirp.c:78-83) on 2n-sample windows that advance  */
/* by n, against G = rfft_2n(window * chirp zero-padded to 2n); lags [n, 2n) of each block are the  */
/* linear filter output.  No input window (overlap-save filters the raw stream).                    */
/* ------------------------------------------------------------------------------------------ */
void ref_correlate_os(const float *tmpl /* n: window * chirp */, uint32_t n, const int32_t *pcm_i32, const float *pcm_f32,
                      uint32_t nframes, float *out /* (nframes-1)*n or NULL */, float *max_val, uint32_t *max_idx) {
    ref_rfft_fast_instance_f32 S;
    if (nframes < 2 || ref_arm_rfft_fast_init_f32(&S, 2 * n) != REF_MATH_SUCCESS) return;
    float *G = (float *) calloc(2 * n, sizeof(float)), *blk = (float *) malloc(sizeof(float) * 2 * n);
    memcpy(G, tmpl, sizeof(float) * n);
    ref_arm_rfft_fast_f32(&S, G, G, 0);
    for (uint32_t b = 0; b + 1 < nframes; ++b) {
        for (uint32_t i = 0; i < 2 * n; ++i)
            blk[i] = pcm_i32 ? (float) pcm_i32[(size_t) b * n + i] : pcm_f32[(size_t) b * n + i];
        ref_arm_rfft_fast_f32(&S, blk, blk, 0);
        ref_arm_cmplx_mult_cmplx_f32(blk, G, blk, n);
        ref_arm_rfft_fast_f32(&S, blk, blk, 1);
        if (out) memcpy(out + (size_t) b * n, blk + n, sizeof(float) * n);
        float v; uint32_t i;
        ref_arm_max_f32(blk + n, n, &v, &i);
        if (max_val) max_val[b] = v;
        if (max_idx) max_idx[b] = i;
    }
    free(G); free(blk);
    ref_arm_rfft_fast_free(&S);
}

/* ------------------------------------------------------------------------------------------ */
/* 4-offset scan — experiments/chirp_compression_freq_domain/Src/main.c:113-160, 245-251         */
/* ------------------------------------------------------------------------------------------ */
void ref_scan4(const ref_rfft_fast_instance_f32 *S, const float *hann, const float *chirp, uint32_t n, uint32_t bandwidth,
               const float *pcm2n, ref_scan_entry *out) {
    float *buf = (float *) malloc(sizeof(float) * n), *mag = (float *) malloc(sizeof(float) * (n / 2));
    const uint32_t bw8 = bandwidth * 8;
    for (uint32_t i = 0; i < 4; ++i) {
        const uint32_t pos = (n / 4) * i;                                           /* main.c:246 */
        for (uint32_t j = 0; j < n; ++j) buf[j] = pcm2n[j + pos];                  /* main.c:247-249 */
        ref_arm_mult_f32(buf, chirp, buf, n);                                       /* chirp.c:42-44 */
        ref_arm_mult_f32(buf, hann, buf, n);                                        /* main.c:123 */
        ref_arm_rfft_fast_f32(S, buf, buf, 0);                                      /* main.c:126, in place */
        ref_arm_cmplx_mag_f32(buf, mag, n / 2);                                     /* main.c:129, in place: */
        memcpy(buf, mag, sizeof(float) * (n / 2));                                  /* lower half only */
        uint32_t ir, il;
        ref_arm_max_f32(&buf[0], bw8, &out[i].mag_max_right, &ir);                  /* main.c:146 */
        ref_arm_max_f32(&buf[n - bw8], bw8, &out[i].mag_max_left, &il);             /* main.c:147 */
        out[i].max_idx_right = ir;
        out[i].max_idx_left = bw8 - il;                                             /* main.c:157 */
    }
    free(buf); free(mag);
}

/* ------------------------------------------------------------------------------------------ */
/* twin of the device-side synthetic generator (Philox-4x32-10, integer arithmetic only)       */
/* ------------------------------------------------------------------------------------------ */
static void philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- the two earlier detectors ------------------------------------------------------------------------ */
void ref_legacy_magnitudes_i32(const int32_t *pcm, size_t nframes, uint32_t n, float *mag, int nthreads) {
    /* experiments/chirp/Src/main.c:200-216 == experiments/ultracom/Src/main.c:115-128 == experiments/basic fft() */
    float *hann = (float *) malloc(sizeof(float) * n);
    ref_hann_window(hann, n, REF_HANN_PERIODIC);
    const float inv = 1.0f / sqrtf((float) n);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        ref_rfft_fast_instance_f32 S;
        ref_arm_rfft_fast_init_f32(&S, (uint16_t) n);
        float *in = (float *) malloc(sizeof(float) * n), *out = (float *) malloc(sizeof(float) * n);
#pragma omp for schedule(static)
        for (long f = 0; f < (long) nframes; ++f) {
            for (uint32_t i = 0; i < n; ++i) in[i] = (float) pcm[(size_t) f * n + i];
            ref_arm_mult_f32(in, hann, in, n);
            ref_arm_rfft_fast_f32(&S, in, out, 0);
            ref_arm_cmplx_mag_f32(out, mag + (size_t) f * (n / 2), n / 2);
            ref_arm_scale_f32(mag + (size_t) f * (n / 2), inv, mag + (size_t) f * (n / 2), n / 2);
        }
        free(in); free(out);
        ref_arm_rfft_fast_free(&S);
    }
    free(hann);
}

void ref_onoff_band(uint32_t n, float fs, float f1, float f2, uint32_t *lo, uint32_t *hi) {
    const float freq1 = f1, freq2 = (float) ((double) f1 + 2.0 * ((double) f2 - (double) f1));   /* main.c:372-373 */
    uint32_t b1 = 0, b2 = 0;
    for (uint32_t i = 0; i < n / 2; ++i) {
        float f = (float) i * fs / (float) n;                                                   /* main.c:376 */
        if (b1 == 0 && f >= freq1) b1 = i;
        if (b2 == 0 && f >= freq2) b2 = i;
    }
    *lo = b1; *hi = b2;
}

void ref_onoff_levels(const float *mag, size_t nframes, uint32_t half, uint32_t lo, uint32_t hi, float mag_threshold,
                      float high_frac, float low_frac, uint16_t *strength, int8_t *level) {
    const uint16_t th = (uint16_t) ((float) (hi - lo + 1) * high_frac), tl = (uint16_t) ((float) (hi - lo + 1) * low_frac);
    for (size_t f = 0; f < nframes; ++f) {
        uint16_t c = 0;
        for (uint32_t i = lo; i <= hi; ++i)
            if (mag[f * half + i] > mag_threshold) c += 1;                                       /* main.c:238-242 */
        if (strength) strength[f] = c;
        if (level) level[f] = c >= th ? 1 : (c <= tl ? -1 : 0);                                  /* main.c:276-286 */
    }
}

uint32_t ref_onoff_decode(const int8_t *level, uint32_t nframes, uint32_t frame_start, uint32_t frame_bit, uint32_t sync_threshold,
                          uint32_t sampling_offset, uint8_t *chars, uint32_t cap, uint32_t *sync_errors) {
    /* decode(), main.c:119-198; the statics become locals of one stream */
    uint16_t count = 0, n = 0, high_count = 0;
    uint8_t bits = 0;
    uint32_t out = 0, errs = 0;
    const uint16_t offset = (uint16_t) (frame_start + sampling_offset), max_length = (uint16_t) (offset + frame_bit * 8);
    for (uint32_t t = 0; t < nframes; ++t) {
        const int lv = level[t];
        const int sampling_point = count == offset + frame_bit * n;
        if (lv > 0) {                                   /* CHIRP_HIGH */
            if (count < offset) high_count++;
            else if (sampling_point) { bits = (uint8_t) (bits | (n < 8 ? (0x80 >> n) : 0)); n++; }
            count++;
        } else if (lv == 0) {                           /* CHIRP_UNKNOWN */
            if (sampling_point) n++;
            count++;
        } else {                                        /* CHIRP_LOW */
            if (count > 0) {
                count++;
                if (sampling_point) n++;
            }
        }
        if (count >= frame_start && high_count < sync_threshold) {   /* Sync error: n and bits stay */
            errs++;
            count = 0;
            high_count = 0;
        }
        if (count >= max_length) {                      /* frame receiving completed */
            count = 0; n = 0; high_count = 0;
            if (chars && out < cap) chars[out] = bits;
            out++;
            bits = 0;
        }
    }
    if (sync_errors) *sync_errors = errs;
    return out;
}

void ref_fsk_codes(const float *mag, size_t nframes, uint32_t half, float fs, uint32_t n, uint32_t sof_bin, uint32_t eof_bin,
                   uint32_t hex0_bin, uint32_t hex_step, uint32_t tolerance, float mag_threshold, uint8_t *code, float *magnitude,
                   float *frequency) {
    for (size_t f = 0; f < nframes; ++f) {
        const float *m = mag + f * half;
        int found = 0;
        uint8_t data = 0xFF;
        uint32_t jj = 0;
        for (int c = 0; c < 18 && !found; ++c) {        /* start of frame, end of frame, then symbols[0..15] (main.c:130-166) */
            const uint32_t centre = c == 0 ? sof_bin : (c == 1 ? eof_bin : hex0_bin + (uint32_t) (c - 2) * hex_step);
            for (uint32_t j = centre - tolerance; j <= centre + tolerance; ++j)
                if (m[j] > mag_threshold) {
                    found = 1;
                    data = c == 0 ? 0xF0 : (c == 1 ? 0xF1 : (uint8_t) (c - 2));
                    jj = j;
                    break;
                }
        }
        code[f] = data;
        if (magnitude) magnitude[f] = found ? m[jj] : 0.0f;
        if (frequency) frequency[f] = found ? (float) (jj + 1) * fs / (float) n : 0.0f;          /* frequency[j + 1], main.c:137 */
    }
}

uint32_t ref_fsk_parse(const uint8_t *code, uint32_t nframes, uint32_t tq_n, uint8_t *chars, uint32_t cap, uint32_t *nsof,
                       uint32_t *neof) {
    /* parser(), ultracom/Src/main.c:175-236 */
    enum { IDLE, DATA_MSB, DATA_LSB } recv_state = IDLE;
    uint8_t hex_data_n = 0xFF, data_cnt = 0, data_msb = 0;
    uint32_t out = 0, sof = 0, eof = 0;
    for (uint32_t t = 0; t < nframes; ++t) {
        const uint8_t data = code[t];
        int output_result = 0;
        if (data != hex_data_n) { data_cnt = 0; hex_data_n = data; }
        else if (data_cnt == tq_n) { }
        else if (++data_cnt == tq_n && hex_data_n != 0xFF) output_result = 1;
        if (!output_result) continue;
        switch (hex_data_n) {
        case 0xF0: recv_state = DATA_MSB; sof++; break;
        case 0xF1: recv_state = IDLE; eof++; break;
        default:
            if (recv_state == DATA_MSB) { data_msb = (uint8_t) (hex_data_n << 4); recv_state = DATA_LSB; }
            else if (recv_state == DATA_LSB) {
                if (chars && out < cap) chars[out] = (uint8_t) (data_msb + hex_data_n);
                out++;
                data_msb = 0;
                recv_state = DATA_MSB;
            }
            break;
        }
    }
    if (nsof) *nsof = sof;
    if (neof) *neof = eof;
    return out;
}

void ref_resample_i16_to_pcm(const int16_t *in, size_t n_in, uint32_t up, uint32_t down, int32_t *out, size_t n_out) {
    const uint32_t K = 32;
    float *taps = (float *) malloc(sizeof(float) * (size_t) up * K);
    for (uint32_t p = 0; p < up; ++p)
        for (uint32_t i = 0; i < K; ++i) {
            double x = ((double) i - (double) (K / 2) + 1.0) - (double) p / (double) up;
            double sc = x == 0.0 ? 1.0 : sin(M_PI * x) / (M_PI * x);
            double w = 0.5 + 0.5 * cos(2.0 * M_PI * x / (double) K);
            taps[(size_t) p * K + i] = (float) (sc * w);
        }
    for (size_t j = 0; j < n_out; ++j) {
        unsigned long long pos = (unsigned long long) j * down;
        long long k0 = (long long) (pos / up);
        const float *h = taps + (size_t) (pos % up) * K;
        float acc = 0.0f;
        for (uint32_t i = 0; i < K; ++i) {
            long long k = k0 - (long long) (K / 2) + 1 + i;
            float x = (k >= 0 && (size_t) k < n_in) ? (float) in[k] : 0.0f;
            acc = FMA(x, h[i], acc);
        }
        out[j] = (int32_t) lrintf(acc) * 256;
    }
    free(taps);
}

static void synth_tables(uint32_t n, float fs, float f0, float f1, double amp, int32_t *tab) {
    const double T = (double) n / (double) fs, k = ((double) f1 - (double) f0) / T;
    for (int down = 0; down < 2; ++down)
        for (uint32_t i = 0; i < n; ++i) {          /* chirp_orth, simulation/signal.py:45-53 */
            double t = T * (double) i / (double) (n - 1);
            double f = down ? (double) f1 - k * t / 2.0 : (double) f0 + k * t / 2.0;
            double arg = 2.0 * M_PI * f * t - M_PI / 2.0;
            tab[(size_t) down * n + i] = (int32_t) llround(amp * (cos(arg) + sin(arg)));
        }
}

/* I/Q transmitter symbols: generator/ChirpGeneratorIQmodulation.ipynb cell 5 (sideband +1, phase -pi/2),
 * simulation/IQ_modulation.ipynb cell 4 (sideband -1, phase 0) */
static void synth_tables_iq(uint32_t n, float fs, double carrier, double bw, int sideband, double phase, double amp, int32_t *tab) {
    const double T = (double) n / (double) fs, k = bw / T;
    for (int down = 0; down < 2; ++down)
        for (uint32_t i = 0; i < n; ++i) {
            double t = T * (double) i / (double) (n - 1);
            double fb = down ? bw / 2.0 - k * t / 2.0 : -bw / 2.0 + k * t / 2.0;
            double arg = (2.0 * M_PI * (carrier + (sideband < 0 ? -fb : fb)) * t) + phase;
            tab[(size_t) down * n + i] = (int32_t) llround(amp * cos(arg));
        }
}

static uint32_t synth_msg_byte(uint32_t k0, uint32_t k1, uint64_t g, uint32_t m) {
    uint32_t r[4];
    philox(k0, k1, (uint32_t) g, (uint32_t) (g >> 32), 0x4D5347u, m >> 4, r);
    return 0x20u + ((r[(m >> 2) & 3u] >> (8u * (m & 3u))) & 0xffu) % 95u;
}

/* Twin of usc_synth_streams: whole streams in the transmitter's frame format (generator/ChirpGenerator.ipynb
 * cell 2): lead_in x G, 7 x H, L, message bits MSB first, guard x G, repeated; per-stream start offset. */
void ref_synth_streams(uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes, uint32_t n, float fs, float f0,
                       float f1, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard, double amp, double noise_sigma,
                       int32_t *pcm, uint32_t *offsets, uint8_t *messages) {
    int32_t *tab = (int32_t *) malloc(sizeof(int32_t) * 2 * n);
    synth_tables(n, fs, f0, f1, amp, tab);
    const int32_t gain = (int32_t) llround(noise_sigma / sqrt(4.0 * (65536.0 * 65536.0 - 1.0) / 12.0) * 65536.0);
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
    const uint32_t pattern = lead_in + 8u + 8u * msg_bytes + guard;
    uint8_t *msg = (uint8_t *) malloc(msg_bytes ? msg_bytes : 1);
    for (uint32_t s = 0; s < nstreams; ++s) {
        const uint64_t g = first_stream + s;
        uint32_t r[4];
        philox(k0, k1, (uint32_t) g, (uint32_t) (g >> 32), 0x0FF5E7u, 0u, r);
        const uint32_t off = r[0] % n;
        if (offsets) offsets[s] = off;
        for (uint32_t m = 0; m < msg_bytes; ++m) {
            msg[m] = (uint8_t) synth_msg_byte(k0, k1, g, m);
            if (messages) messages[(size_t) s * msg_bytes + m] = msg[m];
        }
        int32_t *dst = pcm + (size_t) s * nframes * n;
        const size_t total = (size_t) nframes * n;
        for (size_t blk = 0; blk < total / 2; ++blk) {
            philox(k0, k1, (uint32_t) g, (uint32_t) (g >> 32), (uint32_t) blk, 2u, r);
            int32_t sn[2];
            sn[0] = (int32_t) ((r[0] & 0xffffu) + (r[0] >> 16) + (r[1] & 0xffffu) + (r[1] >> 16)) - 131070;
            sn[1] = (int32_t) ((r[2] & 0xffffu) + (r[2] >> 16) + (r[3] & 0xffffu) + (r[3] >> 16)) - 131070;
            for (int e = 0; e < 2; ++e) {
                const size_t i = 2 * blk + e;
                int32_t v = (int32_t) (((int64_t) sn[e] * gain) / 65536);
                if (i >= off) {
                    const size_t t = i - off;
                    const uint32_t k = (uint32_t) ((t / n) % pattern), tau = (uint32_t) (t % n);
                    int kind = 0;
                    if (k >= lead_in && k < lead_in + 7u) kind = 1;
                    else if (k == lead_in + 7u) kind = 2;
                    else if (k > lead_in + 7u && k < lead_in + 8u + 8u * msg_bytes) {
                        const uint32_t b = k - (lead_in + 8u);
                        kind = ((msg[b >> 3] >> (7u - (b & 7u))) & 1u) ? 1 : 2;
                    }
                    if (kind) v += tab[(size_t) (kind - 1) * n + tau];
                }
                dst[i] = v * 256;
            }
        }
    }
    free(msg);
    free(tab);
}

static void synth_frames_from_table(const int32_t *tab, uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n,
                                    double noise_sigma, int32_t *pcm, uint8_t *bits) {
    const int32_t gain = (int32_t) llround(noise_sigma / sqrt(4.0 * (65536.0 * 65536.0 - 1.0) / 12.0) * 65536.0);
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
    for (size_t fl = 0; fl < nframes; ++fl) {
        uint64_t f = first_frame + fl;
        uint32_t r[4];
        philox(k0, k1, (uint32_t) f, (uint32_t) (f >> 32), 0xB175u, 0u, r);
        uint32_t bit = r[0] & 1u;
        if (bits) bits[fl] = (uint8_t) bit;
        for (uint32_t blk = 0; blk < n / 2; ++blk) {
            philox(k0, k1, (uint32_t) f, (uint32_t) (f >> 32), blk, 1u, r);
            const int32_t *t2 = tab + (size_t) (bit ? 0 : 1) * n + 2 * blk;
            int32_t s0 = (int32_t) ((r[0] & 0xffffu) + (r[0] >> 16) + (r[1] & 0xffffu) + (r[1] >> 16)) - 131070;
            int32_t s1 = (int32_t) ((r[2] & 0xffffu) + (r[2] >> 16) + (r[3] & 0xffffu) + (r[3] >> 16)) - 131070;
            pcm[fl * n + 2 * blk] = (t2[0] + (int32_t) (((int64_t) s0 * gain) / 65536)) * 256;
            pcm[fl * n + 2 * blk + 1] = (t2[1] + (int32_t) (((int64_t) s1 * gain) / 65536)) * 256;
        }
    }
}

void ref_synth_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, float fs, float f0, float f1,
                      double amp, double noise_sigma, int32_t *pcm, uint8_t *bits) {
    int32_t *tab = (int32_t *) malloc(sizeof(int32_t) * 2 * n);
    synth_tables(n, fs, f0, f1, amp, tab);
    synth_frames_from_table(tab, seed, first_frame, nframes, n, noise_sigma, pcm, bits);
    free(tab);
}

void ref_synth_iq_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, float fs, double carrier, double bw,
                         int sideband, double phase, double amp, double noise_sigma, int32_t *pcm, uint8_t *bits) {
    int32_t *tab = (int32_t *) malloc(sizeof(int32_t) * 2 * n);
    synth_tables_iq(n, fs, carrier, bw, sideband, phase, amp, tab);
    synth_frames_from_table(tab, seed, first_frame, nframes, n, noise_sigma, pcm, bits);
    free(tab);
}
