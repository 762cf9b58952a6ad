/*
 * ref_dsp.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic on the reference's demodulation hot path:
 * the CMSIS-DSP V1.4.5b operator surface the receiver calls (contracts from
 * receiver/Drivers/CMSIS/Include/arm_math.h) plus the receiver / experiment DSP chains built on
 * it (receiver/Src/main.c, receiver/Src/chirp.c, experiments/<x>/Src/<y>.c).  Every function cites
 * the reference file:line it follows.  All paths are relative to /root/reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product library (libusc.so) never does.
 *
 * PARITY PINNING:
 *   pinned by device captures (agent/chirp_experiment, agent/vaccum_cleaner .raw/.flt/.fft):
 *     int32->float cast, ref_arm_cos_f32 (bit-exact: 0 mismatches in 24x2048 .flt samples),
 *     periodic Hann table, ref_arm_mult_f32, forward ref_arm_rfft_fast_f32 + ref_arm_cmplx_mag_f32
 *     + ref_arm_scale_f32 (to the 6-decimal print precision / fp32 rounding), arg-max bin.
 *   pinned by notebook known-answers: the de-chirp peak-location law (ChirpSynchronization.ipynb).
 *   UNPINNED by any reference artefact (no input/output pairs exist): arm_cmplx_mult_cmplx_f32,
 *     inverse rfft, arm_cfft_f32, arm_fir_f32, arm_mean_f32, arm_sin_cos_f32, the state machine.
 *     Those follow the arm_math.h contracts and are cross-checked against numpy float64.
 *
 * CANONICAL ARITHMETIC.  CMSIS-DSP is vendored only as an ARM-Thumb archive, so bit-equality with
 * the device FFT is impossible.  Instead this file fixes one fp32 operation order ("canonical
 * arithmetic") that the CUDA kernels reproduce operation for operation, so that
 * integer outputs (arg-max bins, sync offsets, symbols) are bit-identical by construction and
 * float outputs are bit-identical too.  Build with -ffp-contract=off: every fused multiply-add in
 * the spec is an explicit fmaf().
 */
#ifndef REF_DSP_H_
#define REF_DSP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float float32_t;                       /* arm_math.h:407 */

typedef enum {                                 /* arm_math.h:373-382 */
    REF_MATH_SUCCESS = 0,
    REF_MATH_ARGUMENT_ERROR = -1
} ref_status;

#define REF_MAX_RADICES 8

/* Canonical FFT plan: master twiddle table W_N^j and the radix list. */
typedef struct {
    uint32_t n;                 /* complex length */
    uint32_t nrad;
    uint32_t rad[REF_MAX_RADICES];
    float *tw;                  /* 2*tw_n floats: (cos, -sin)(2*pi*j/tw_n), j < tw_n */
    uint32_t tw_n;              /* master table length (n for cfft, 2n for the rfft that owns it) */
} ref_fft_plan;

typedef struct {                /* mirrors arm_rfft_fast_instance_f32, arm_math.h:2235-2240 */
    uint16_t fftLenRFFT;
    ref_fft_plan cplx;          /* N/2 complex plan sharing the N-entry master table */
} ref_rfft_fast_instance_f32;

typedef struct {                /* mirrors arm_cfft_instance_f32, arm_math.h:2141-2147 */
    uint16_t fftLen;
    ref_fft_plan plan;
} ref_cfft_instance_f32;

typedef struct {                /* mirrors arm_fir_instance_f32, arm_math.h:1059-1064 */
    uint16_t numTaps;
    float32_t *pState;
    const float32_t *pCoeffs;
} ref_fir_instance_f32;

/* ---- CMSIS-shaped primitives (arm_math.h line numbers in ref_dsp.c) ---- */
float32_t ref_arm_cos_f32(float32_t x);
void ref_arm_sin_cos_f32(float32_t theta_deg, float32_t *pSinVal, float32_t *pCosVal);
void ref_arm_mult_f32(const float32_t *a, const float32_t *b, float32_t *dst, uint32_t n);
void ref_arm_scale_f32(const float32_t *src, float32_t scale, float32_t *dst, uint32_t n);
void ref_arm_copy_f32(const float32_t *src, float32_t *dst, uint32_t n);
void ref_arm_mean_f32(const float32_t *src, uint32_t n, float32_t *result);
void ref_arm_max_f32(const float32_t *src, uint32_t n, float32_t *result, uint32_t *index);
void ref_arm_cmplx_mult_cmplx_f32(const float32_t *a, const float32_t *b, float32_t *dst, uint32_t ncplx);
void ref_arm_cmplx_mult_real_f32(const float32_t *cplx, const float32_t *real, float32_t *dst, uint32_t ncplx);
void ref_arm_cmplx_mag_f32(const float32_t *src, float32_t *dst, uint32_t ncplx);
ref_status ref_arm_rfft_fast_init_f32(ref_rfft_fast_instance_f32 *S, uint32_t fftLen);
void ref_arm_rfft_fast_free(ref_rfft_fast_instance_f32 *S);
void ref_arm_rfft_fast_f32(const ref_rfft_fast_instance_f32 *S, float32_t *p, float32_t *pOut, uint8_t ifftFlag);
ref_status ref_arm_cfft_init_f32(ref_cfft_instance_f32 *S, uint32_t fftLen);
void ref_arm_cfft_free(ref_cfft_instance_f32 *S);
void ref_arm_cfft_f32(const ref_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);
void ref_arm_fir_init_f32(ref_fir_instance_f32 *S, uint16_t numTaps, const float32_t *pCoeffs,
                          float32_t *pState, uint32_t blockSize);
void ref_arm_fir_f32(const ref_fir_instance_f32 *S, const float32_t *pSrc, float32_t *pDst, uint32_t blockSize);

/* Exposed for tests: canonical twiddle (cos, -sin)(2*pi*j/n) rounded once from double. */
void ref_twiddle(uint32_t j, uint32_t n, float *re, float *im);
/* Canonical radix list for a complex length. Returns count, 0 if unsupported. */
uint32_t ref_fft_radices(uint32_t n, uint32_t *rad);

/* ---- tables ---- */
typedef enum { REF_HANN_PERIODIC = 0, REF_HANN_SYMMETRIC = 1 } ref_hann_kind;
void ref_hann_window(float32_t *w, uint32_t n, ref_hann_kind kind);

typedef enum {
    REF_CHIRP_R = 0,   /* receiver/Src/chirp.c:16-40: real, sin(theta-90deg), degrees, /2 law */
    REF_CHIRP_S = 1,   /* experiments/synchronization/Src/chirp.c:16-44: complex (cos,sin) interleaved */
    REF_CHIRP_T = 2,   /* experiments/chirp_compression_time_domain/Src/chirp.c:25-45: real cos, rad, no /2 */
    REF_CHIRP_F = 3    /* experiments/chirp_compression_freq_domain/Src/chirp.c:15-35: as T, phase ignored */
} ref_chirp_variant;

typedef struct {
    uint32_t n;            /* samples per frame (NN / PCM_SAMPLES) */
    float fs;              /* sampling rate as the firmware holds it (float) */
    float f0, f1;          /* sweep range (F0,F1 or F1,F2 in the reference headers) */
    float sweep_T;         /* TIME_FRAME for R/S (0.0205f); ignored by T/F (they use n/fs) */
    float phase;           /* -90.0f (deg) for R/S; -PI/2 (rad) for T; ignored by F */
} ref_chirp_params;

/* out: n floats (R,T,F) or 2n floats (S). up != 0 -> up-chirp. */
void ref_generate_ref_chirp(ref_chirp_variant v, const ref_chirp_params *p, int up, float32_t *out);

/* ---- receiver chain (receiver/Src/main.c) ---- */
typedef struct {                         /* receiver/Src/main.c:124-136 (timing fields dropped) */
    float mag_max, mag_max_left, mag_max_right;
    int32_t max_freq, max_freq_left, max_freq_right;
    uint32_t max_idx, max_idx_left, max_idx_right;   /* raw bins (not in the reference struct) */
    float mag_mean, snr;
    char rank;
} ref_history;

typedef struct {
    uint32_t n;                 /* NN */
    float fs;
    uint32_t bandwidth, bandwidth2, idx_left_zero;   /* main.c:372-374 */
    float *hann;                /* periodic Hann, main.c:390-393 */
    float *up_chirp, *down_chirp;   /* variant R tables */
    ref_rfft_fast_instance_f32 S;
} ref_receiver;

int ref_receiver_init(ref_receiver *rx, uint32_t n, float fs, float f0, float f1, float sweep_T);
void ref_receiver_free(ref_receiver *rx);
int32_t ref_idx2freq(const ref_receiver *rx, uint32_t idx);                  /* main.c:154-160 */
/* main.c:163-180.  signal: n floats in, n floats out: mags of the n/2 packed bins in [0,n/2),
 * zeros in [n/2,n) (hazard H1 defined: the uninitialised upper half reads as zero). */
void ref_pipeline(const ref_receiver *rx, float32_t *signal, int up);
/* main.c:183-231.  fifo: 3n floats. */
void ref_dsp(const ref_receiver *rx, const float32_t *fifo, uint32_t sync_position,
             ref_history *h, float mag_mean, int up);

/* Batched aligned-frame demodulation = dsp() for UP and DOWN on every frame (config 2).
 * pcm: nframes*n int32 samples.  out arrays have nframes entries. */
void ref_demod_frames_i32(const ref_receiver *rx, const int32_t *pcm, size_t nframes,
                          float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down,
                          int nthreads);
void ref_demod_frames_f32(const ref_receiver *rx, const float *pcm, size_t nframes,
                          float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down,
                          int nthreads);
/* Tuned form of the two functions above (ref_fast.c: SIMD across frames, iterative plan, per-thread scratch;
 * bit-identical results).  n must be 2048.  Returns 0, or -1 for an unsupported receiver.  The CPU arm of bench.py. */
int ref_fast_demod_frames(const ref_receiver *rx, const void *pcm, int is_float, size_t nframes, float *mag_up,
                          uint32_t *idx_up, float *mag_down, uint32_t *idx_down, int nthreads);
int ref_fast_simd_width(void);          /* 16 (AVX-512), 8 (AVX2 + FMA) or 0 (plain form) */

/* ---- complex-FFT variant (experiments/synchronization/Src/main.c:135-213, chirp.c:16-57) ---- */
typedef struct {
    uint32_t n;
    float fs;
    uint32_t bandwidth, bandwidth2, idx_left_zero;
    float *hann;                    /* periodic Hann (main.c:77) */
    float *up_chirp, *down_chirp;   /* variant S: 2n floats, interleaved (cos, sin) */
    ref_cfft_instance_f32 C;        /* arm_cfft_sR_f32_len2048 */
} ref_sync_receiver;

int ref_sync_receiver_init(ref_sync_receiver *rx, uint32_t n, float fs, float f0, float f1, float sweep_T);
void ref_sync_receiver_free(ref_sync_receiver *rx);
/* main.c:144-158: pframe holds n complex (2n floats) in, n magnitudes out in pframe[0..n) */
void ref_sync_pipeline(const ref_sync_receiver *rx, float32_t *pframe, int up);
/* main.c:161-213: real PCM -> complex buffer, pipeline, left/right windowed arg-max, SNR */
void ref_sync_dsp(const ref_sync_receiver *rx, const float32_t *fifo, uint32_t sync_position, ref_history *h,
                  float mag_mean, int up);

/* ---- I/Q baseband path (experiments/iq_modulation/Src/iq_modem.c:16-66, main.c:117-134; semantics of
 * simulation/IQ_modulation.ipynb cells 16-31; BASELINE config 3: decimate by 2 -> 1024-pt complex FFT) ---- */
typedef struct {
    uint32_t n;                 /* real samples per frame (2048) */
    float fs;
    uint32_t window_bins;       /* search window [0,W) and [n/2-W, n/2) around DC */
    float *carrier_cos, *carrier_sin;     /* n each, iq_modem.c:34-46 */
    float *chirp, *chirp_conj;  /* n/2 complex each: baseband chirp at fs/2 and its conjugate */
    float *hann;                /* periodic Hann, n/2 */
    float taps[64];             /* CMSIS order (time-reversed) */
    uint32_t num_taps;
    ref_cfft_instance_f32 C;    /* arm_cfft_sR_f32_len1024 */
} ref_iq;

int ref_iq_init(ref_iq *q, uint32_t n, float fs, float carrier, float bw, float sweep_T, const float *taps,
                uint32_t num_taps, uint32_t window_bins);
void ref_iq_free(ref_iq *q);
/* One stream of nframes frames (FIR state carried frame to frame, zero at the start).
 * Outputs nframes each. */
void ref_iq_demod_i32(const ref_iq *q, const int32_t *pcm, uint32_t nframes, float *mag_up, uint32_t *idx_up,
                      float *mag_down, uint32_t *idx_down);

/* ---- 4-offset scan of experiments/chirp_compression_freq_domain/Src/main.c:113-160, 245-251 ---- */
typedef struct { float mag_max_right, mag_max_left; uint32_t max_idx_right, max_idx_left; } ref_scan_entry;
/* pcm2n: 2n floats (the 2-frame buffer, main.c:385-394).  hann: periodic; chirp: variant F down-chirp.
 * out: 4 entries (SYNC_RESOLUTION), offsets 0, n/4, n/2, 3n/4.  The "left" window reads the upper half of
 * the in-place buffer, which still holds packed-spectrum floats after the n/2 magnitudes were written
 * (main.c:126-147); max_idx_left = bandwidth*8 - index (main.c:157). */
void ref_scan4(const ref_rfft_fast_instance_f32 *S, const float *hann, const float *chirp, uint32_t n, uint32_t bandwidth,
               const float *pcm2n, ref_scan_entry *out);

/* ---- twin of the device-side synthetic generator (usc_synth_frames): regenerates any frame on the CPU ---- */
/* ---- the two earlier detectors (experiments/chirp, experiments/ultracom), n = 2048 ---- */
/* front half shared with the analyser: (float) pcm x Hann (periodic) -> RFFT -> magnitude * 1/sqrt(N); mag: nframes x n/2 */
void ref_legacy_magnitudes_i32(const int32_t *pcm, size_t nframes, uint32_t n, float *mag, int nthreads);
/* experiments/chirp/Src/main.c:237-286, 372-387: band indices from (fs, F1, F2), strength, level (1 / -1 / 0) */
void ref_onoff_band(uint32_t n, float fs, float f1, float f2, uint32_t *lo, uint32_t *hi);
void ref_onoff_levels(const float *mag, size_t nframes, uint32_t half, uint32_t lo, uint32_t hi, float mag_threshold,
                      float high_frac, float low_frac, uint16_t *strength, int8_t *level);
/* decode(), :119-198, over one stream's levels; returns the number of completed frames (bytes), writes min(that, cap) */
uint32_t ref_onoff_decode(const int8_t *level, uint32_t nframes, uint32_t frame_start, uint32_t frame_bit, uint32_t sync_threshold,
                          uint32_t sampling_offset, uint8_t *chars, uint32_t cap, uint32_t *sync_errors);
/* experiments/ultracom/Src/main.c:130-168: code, magnitude, frequency[j + 1] per frame */
void ref_fsk_codes(const float *mag, size_t nframes, uint32_t half, float fs, uint32_t n, uint32_t sof_bin, uint32_t eof_bin,
                   uint32_t hex0_bin, uint32_t hex_step, uint32_t tolerance, float mag_threshold, uint8_t *code, float *magnitude,
                   float *frequency);
/* parser(), :175-236, over one stream's codes */
uint32_t ref_fsk_parse(const uint8_t *code, uint32_t nframes, uint32_t tq_n, uint8_t *chars, uint32_t cap, uint32_t *nsof,
                       uint32_t *neof);
/* twin of usc_resample_i16_to_pcm: 32-tap Hann-windowed-sinc polyphase FIR at the exact ratio up/down */
void ref_resample_i16_to_pcm(const int16_t *in, size_t n_in, uint32_t up, uint32_t down, int32_t *out, size_t n_out);
void ref_synth_streams(uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes, uint32_t n, float fs, float f0,
                       float f1, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard, double amp, double noise_sigma,
                       int32_t *pcm, uint32_t *offsets, uint8_t *messages);
void ref_synth_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, float fs, float f0, float f1,
                      double amp, double noise_sigma, int32_t *pcm, uint8_t *bits);
/* twin of usc_synth_iq_frames: the I/Q transmitter's symbols (generator/ChirpGeneratorIQmodulation.ipynb cell 5,
 * simulation/IQ_modulation.ipynb cell 4) through the same generator */
void ref_synth_iq_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, float fs, double carrier, double bw,
                         int sideband, double phase, double amp, double noise_sigma, int32_t *pcm, uint8_t *bits);

/* twin of usc_correlate_os: overlap-save linear filtering of one stream (nframes frames of n samples, int32 or float)
 * with the n-sample template (window * chirp); per block b < nframes-1: the n valid lags, their arm_max_f32 */
void ref_correlate_os(const float *tmpl, uint32_t n, const int32_t *pcm_i32, const float *pcm_f32, uint32_t nframes,
                      float *out, float *max_val, uint32_t *max_idx);

/* ---- receiver state machine (receiver/Src/main.c:417-580) ---- */
enum { REF_IDLE = 0, REF_SYNCHRONIZING = 1, REF_SYNCHRONIZED = 2, REF_DATA_RECEIVING = 3 };   /* main.c:108-111 */

typedef struct {
    uint32_t state, turn, sync_cnt, sync_position, max_idx;
    float mag_stat[12];                      /* main.c:321-322: starts at 1e37 */
    float mag_mean;
    ref_history history[8];                  /* hazard H4 defined: zero-initialised */
    uint32_t msg, msg_cnt;
    int32_t lock_frame;                      /* frame index of the first IDLE/SYNCHRONIZING -> SYNCHRONIZED, -1 if none */
    uint32_t lock_position;                  /* sync_position chosen at that moment */
    uint32_t frames_seen;
} ref_rx_state;

void ref_rx_state_init(ref_rx_state *st, const ref_receiver *rx);
/* One pass of the while(1) body with new_pcm_data set (main.c:422-577).  fifo: 3n floats, already
 * shifted (main.c:659-668).  UART bytes (decoded chars and the '\n' that ends a message,
 * main.c:533,540) are appended to out[*nout], up to cap. */
void ref_receiver_step(const ref_receiver *rx, ref_rx_state *st, float snr_threshold, const float *fifo,
                       uint8_t *out, uint32_t *nout, uint32_t cap);
/* A whole stream: nframes frames of int32 PCM pushed through the 3-frame FIFO (zeros before the
 * first frame) and the state machine. */
void ref_receiver_run_i32(const ref_receiver *rx, float snr_threshold, const int32_t *pcm, uint32_t nframes,
                          uint8_t *out, uint32_t *nout, uint32_t cap, ref_rx_state *final_state);
/* The sliding-correlation search grid alone (main.c:447-451), optionally after synchronous addition
 * of `sync_add` frame-aligned windows (misc/Formula.ipynb cell 9, SynchronousAddition.ipynb cell 6):
 * for frame t and i < 4: dsp(UP) at N/2 + (t&1)*N/8 + i*N/4 on the sum over j < sync_add of the
 * FIFO as it stood at frame t-j.  mag/idx: nframes*4. */
void ref_sync_search_i32(const ref_receiver *rx, const int32_t *pcm, uint32_t nframes, uint32_t sync_add,
                         float *mag, uint32_t *idx);

/* ---- frequency-domain compression chain (experiments/chirp_compression_time_domain) ---- */
typedef struct {
    uint32_t n;
    float fs;
    float *window;              /* symmetric Hann, chirp.c:13,63-65 */
    float *H_up, *H_down;       /* packed spectra of the windowed reference chirps, chirp.c:58-73 */
    ref_rfft_fast_instance_f32 S;
} ref_compressor;

int ref_compressor_init(ref_compressor *c, uint32_t n, float fs, float f1, float f2);
void ref_compressor_free(ref_compressor *c);
/* chirp.c:78-83: window -> rfft -> x H (packed, n/2 complex, DC/Nyquist quirk) -> irfft. in place */
void ref_compress_chirp(const ref_compressor *c, float32_t *inout, int use_up);
/* main.c:171-189 batched: compress + signed arm_max over all n lags. */
void ref_compress_frames_i32(const ref_compressor *c, const int32_t *pcm, size_t nframes, int use_up,
                             float *max_val, uint32_t *max_idx, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* REF_DSP_H_ */
