/*
 * usc_cmsis_shim.h — the CMSIS-DSP V1.4.5b names and signatures the receiver calls, as a 1-frame
 * HOST-pointer compatibility layer over libusc (libusc_cmsis.so).  With it the reference's
 * pipeline()/dsp()/compress_chirp() (receiver/Src/main.c:163-231, experiments/<x>/Src/chirp.c) link
 * and run unchanged against the GPU library.  It exists to prove the drop-in boundary and for
 * tests; it moves one frame per call over PCIe, so it is NOT the fast path (use usc.h for that).
 *
 * Struct layouts mirror receiver/Drivers/CMSIS/Include/arm_math.h so that caller-allocated instances
 * have the right size; only the length fields are used.
 */
#ifndef USC_CMSIS_SHIM_H_
#define USC_CMSIS_SHIM_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef float float32_t;                                        /* arm_math.h:407 */
typedef enum { ARM_MATH_SUCCESS = 0, ARM_MATH_ARGUMENT_ERROR = -1 } arm_status;   /* arm_math.h:373-382 */

typedef struct {                                                /* arm_math.h:2141-2147 */
    uint16_t fftLen;
    const float32_t *pTwiddle;
    const uint16_t *pBitRevTable;
    uint16_t bitRevLength;
} arm_cfft_instance_f32;

typedef struct {                                                /* arm_math.h:2235-2240 */
    arm_cfft_instance_f32 Sint;
    uint16_t fftLenRFFT;
    float32_t *pTwiddleRFFT;
} arm_rfft_fast_instance_f32;

typedef struct {                                                /* arm_math.h:1059-1064 */
    uint16_t numTaps;
    float32_t *pState;
    float32_t *pCoeffs;
} arm_fir_instance_f32;

extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len1024;     /* arm_const_structs.h:55 */
extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len2048;     /* arm_const_structs.h:56 */

arm_status arm_rfft_fast_init_f32(arm_rfft_fast_instance_f32 *S, uint16_t fftLen);                 /* :2242-2244 */
void arm_rfft_fast_f32(arm_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag);   /* :2246-2249 */
void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag); /* :2149-2153 */
void arm_mult_f32(float32_t *pSrcA, float32_t *pSrcB, float32_t *pDst, uint32_t blockSize);        /* :1938-1942 */
void arm_scale_f32(float32_t *pSrc, float32_t scale, float32_t *pDst, uint32_t blockSize);          /* :2508 */
void arm_copy_f32(float32_t *pSrc, float32_t *pDst, uint32_t blockSize);                             /* :2819 */
void arm_mean_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult);                          /* :6192 */
void arm_max_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult, uint32_t *pIndex);        /* :6537-6541 */
void arm_cmplx_mult_cmplx_f32(float32_t *pSrcA, float32_t *pSrcB, float32_t *pDst, uint32_t numSamples);  /* :6579-6583 */
void arm_cmplx_mult_real_f32(float32_t *pSrcCmplx, float32_t *pSrcReal, float32_t *pCmplxDst, uint32_t numSamples); /* :6425-6429 */
void arm_cmplx_mag_f32(float32_t *pSrc, float32_t *pDst, uint32_t numSamples);                       /* :6312-6315 */
void arm_fir_init_f32(arm_fir_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState,
                      uint32_t blockSize);                                                             /* :1194-1214 */
void arm_fir_f32(const arm_fir_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
float32_t arm_cos_f32(float32_t x);                                                                    /* :5685 */
void arm_sin_cos_f32(float32_t theta, float32_t *pSinVal, float32_t *pCosVal);                        /* :4634-4637 */

/* last libusc status seen by the shim (the CMSIS processing functions return void) */
int usc_cmsis_last_status(void);

#ifdef __cplusplus
}
#endif
#endif
