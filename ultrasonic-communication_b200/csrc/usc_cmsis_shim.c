/* usc_cmsis_shim.c — see include/usc_cmsis_shim.h.  Plain C over the libusc C-ABI: every call stages
 * its operands to the device, runs the batched operator with batch = 1, and copies the result back. */
#include "../../include/usc_cmsis_shim.h"

#include <string.h>

#include "../../include/usc.h"
#include "usc_tables.h"

const arm_cfft_instance_f32 arm_cfft_sR_f32_len1024 = {1024, 0, 0, 0};
const arm_cfft_instance_f32 arm_cfft_sR_f32_len2048 = {2048, 0, 0, 0};

static usc_handle *g_h;
static int g_status;
static void *g_dev[3];
static size_t g_cap[3];

int usc_cmsis_last_status(void) { return g_status; }

static int ready(void) {
    if (g_h) return 1;
    usc_config cfg;
    usc_default_config(&cfg);
    g_status = usc_create(&cfg, 0, &g_h);
    return g_status == USC_OK;
}

static void *scratch(int slot, size_t bytes) {
    if (g_cap[slot] < bytes) {
        if (g_dev[slot]) usc_free(g_dev[slot]);
        g_dev[slot] = 0;
        g_cap[slot] = 0;
        g_status = usc_malloc(&g_dev[slot], bytes);
        if (g_status != USC_OK) return 0;
        g_cap[slot] = bytes;
    }
    return g_dev[slot];
}

static int put(int slot, const void *host, size_t bytes) {
    void *d = scratch(slot, bytes);
    if (!d) return 0;
    g_status = usc_memcpy_h2d(g_h, d, host, bytes);
    return g_status == USC_OK;
}

static void get(void *host, int slot, size_t bytes) {
    if (g_status != USC_OK) return;
    g_status = usc_memcpy_d2h(g_h, host, g_dev[slot], bytes);
    if (g_status == USC_OK) g_status = usc_sync(g_h);
}

arm_status arm_rfft_fast_init_f32(arm_rfft_fast_instance_f32 *S, uint16_t fftLen) {
    if (!S || fftLen < 32 || fftLen > 4096 || (fftLen & (fftLen - 1))) return ARM_MATH_ARGUMENT_ERROR;
    memset(S, 0, sizeof *S);
    S->fftLenRFFT = fftLen;
    S->Sint.fftLen = fftLen / 2;
    return ARM_MATH_SUCCESS;
}

void arm_rfft_fast_f32(arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag) {
    const size_t b = (size_t) S->fftLenRFFT * 4;
    if (!ready() || !put(0, p, b) || !scratch(1, b)) return;
    g_status = usc_arm_rfft_fast_f32_batch(g_h, S->fftLenRFFT, (const float *) g_dev[0], (float *) g_dev[1], ifftFlag, 1);
    get(pOut, 1, b);
}

void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag) {
    (void) bitReverseFlag;                      /* natural-order output only (the reference passes 1) */
    const size_t b = (size_t) S->fftLen * 8;
    if (!ready() || !put(0, p1, b)) return;
    g_status = usc_arm_cfft_f32_batch(g_h, S->fftLen, (float *) g_dev[0], ifftFlag, 1);
    get(p1, 0, b);
}

void arm_mult_f32(float32_t *a, float32_t *bsrc, float32_t *dst, uint32_t n) {
    const size_t b = (size_t) n * 4;
    if (!ready() || !put(0, a, b) || !put(1, bsrc, b) || !scratch(2, b)) return;
    g_status = usc_arm_mult_f32_batch(g_h, (const float *) g_dev[0], n, (const float *) g_dev[1], n, (float *) g_dev[2], n, n, 1);
    get(dst, 2, b);
}

void arm_scale_f32(float32_t *src, float32_t scale, float32_t *dst, uint32_t n) {
    const size_t b = (size_t) n * 4;
    if (!ready() || !put(0, src, b) || !scratch(1, b)) return;
    g_status = usc_arm_scale_f32_batch(g_h, (const float *) g_dev[0], scale, (float *) g_dev[1], n, 1);
    get(dst, 1, b);
}

void arm_copy_f32(float32_t *src, float32_t *dst, uint32_t n) { memmove(dst, src, (size_t) n * 4); }

void arm_mean_f32(float32_t *src, uint32_t n, float32_t *result) {
    if (!ready() || !put(0, src, (size_t) n * 4) || !scratch(1, 4)) return;
    g_status = usc_arm_mean_f32_batch(g_h, (const float *) g_dev[0], n, n, (float *) g_dev[1], 1);
    get(result, 1, 4);
}

void arm_max_f32(float32_t *src, uint32_t n, float32_t *result, uint32_t *index) {
    if (!ready() || !put(0, src, (size_t) n * 4) || !scratch(1, 8)) return;
    g_status = usc_arm_max_f32_batch(g_h, (const float *) g_dev[0], n, n, (float *) g_dev[1], (uint32_t *) g_dev[1] + 1, 1);
    uint32_t out[2];
    get(out, 1, 8);
    memcpy(result, &out[0], 4);
    if (index) *index = out[1];
}

void arm_cmplx_mult_cmplx_f32(float32_t *a, float32_t *bsrc, float32_t *dst, uint32_t ns) {
    const size_t b = (size_t) ns * 8;
    if (!ready() || !put(0, a, b) || !put(1, bsrc, b) || !scratch(2, b)) return;
    g_status = usc_arm_cmplx_mult_cmplx_f32_batch(g_h, (const float *) g_dev[0], 2 * (size_t) ns, (const float *) g_dev[1],
                                                  2 * (size_t) ns, (float *) g_dev[2], 2 * (size_t) ns, ns, 1);
    get(dst, 2, b);
}

void arm_cmplx_mult_real_f32(float32_t *c, float32_t *r, float32_t *dst, uint32_t ns) {
    if (!ready() || !put(0, c, (size_t) ns * 8) || !put(1, r, (size_t) ns * 4) || !scratch(2, (size_t) ns * 8)) return;
    g_status = usc_arm_cmplx_mult_real_f32_batch(g_h, (const float *) g_dev[0], 2 * (size_t) ns, (const float *) g_dev[1], ns,
                                                 (float *) g_dev[2], 2 * (size_t) ns, ns, 1);
    get(dst, 2, (size_t) ns * 8);
}

void arm_cmplx_mag_f32(float32_t *src, float32_t *dst, uint32_t ns) {
    if (!ready() || !put(0, src, (size_t) ns * 8) || !scratch(1, (size_t) ns * 4)) return;
    g_status = usc_arm_cmplx_mag_f32_batch(g_h, (const float *) g_dev[0], 2 * (size_t) ns, (float *) g_dev[1], ns, ns, 1);
    get(dst, 1, (size_t) ns * 4);
}

void arm_fir_init_f32(arm_fir_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState, uint32_t blockSize) {
    S->numTaps = numTaps;
    S->pCoeffs = pCoeffs;
    S->pState = pState;
    memset(pState, 0, sizeof(float) * (numTaps + blockSize - 1));
}

void arm_fir_f32(const arm_fir_instance_f32 *S, float32_t *src, float32_t *dst, uint32_t n) {
    /* the caller-owned pState keeps the numTaps-1 history samples at its front between calls */
    const size_t b = (size_t) n * 4, sb = (size_t) (S->numTaps - 1) * 4;
    if (!ready() || !put(0, src, b) || !scratch(1, b)) return;
    void *dstate = 0;
    if ((g_status = usc_malloc(&dstate, sb ? sb : 4)) != USC_OK) return;
    g_status = usc_memcpy_h2d(g_h, dstate, S->pState, sb);
    if (g_status == USC_OK)
        g_status = usc_arm_fir_f32_batch(g_h, S->pCoeffs, S->numTaps, (float *) dstate, (const float *) g_dev[0], (float *) g_dev[1], n, 1);
    if (g_status == USC_OK) g_status = usc_memcpy_d2h(g_h, S->pState, dstate, sb);
    get(dst, 1, b);
    usc_free(dstate);
}

float32_t arm_cos_f32(float32_t x) { return usc_host_arm_cos_f32(x); }
void arm_sin_cos_f32(float32_t theta, float32_t *s, float32_t *c) { usc_host_arm_sin_cos_f32(theta, s, c); }
