/*
 * usc.h — C-ABI of libusc.so: the B200 (sm_100a) demodulation chain of the ultrasonic chirp receiver.
 *
 * This is the drop-in boundary for the reference's hot path: the CMSIS-DSP V1.4.5b
 * operator surface the receiver calls (receiver/Drivers/CMSIS/Include/arm_math.h) in batched,
 * device-pointer form, plus fused stage-level entry points that mirror the reference's own
 * functions (pipeline/dsp/compress_chirp/...).  Paths below are relative to the reference tree.
 *
 * Conventions
 *   - plain C, no CUDA or torch types; device buffers are raw device addresses (float*, int32_t*)
 *     obtained from usc_malloc() or from any CUDA allocator in the same process (e.g. torch).
 *   - every call is asynchronous on the handle's stream (usc_set_stream, default stream 0) unless
 *     stated; usc_sync() waits.  No hidden allocation on the hot calls.
 *   - return 0 on success; USC_ERR_ARGUMENT (-1, == ARM_MATH_ARGUMENT_ERROR, arm_math.h:373-382) for
 *     bad arguments; USC_ERR_CUDA_BASE - cudaError for CUDA failures.  Nothing throws.
 *   - there is NO CPU fallback: without a usable CUDA device usc_create() fails.
 *   - complex data is interleaved (re, im) (arm_math.h:181-184); real-FFT spectra use the CMSIS
 *     packed layout [X0.re, X(N/2).re, X1.re, X1.im, ...] (arm_math.h:2246-2249).
 *   - arithmetic is fp32 in the fixed operation order of DESIGN.md §3 ("canonical arithmetic"),
 *     bit-identical to the CPU oracle; integer results (bins, offsets, symbols) are exact.
 */
#ifndef USC_H_
#define USC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USC_OK 0
#define USC_ERR_ARGUMENT (-1)
#define USC_ERR_NOMEM (-2)
#define USC_ERR_CUDA_BASE (-1000)

/* reference chirp table variants */
#define USC_CHIRP_R 0u /* receiver/Src/chirp.c:16-40: real sin(theta-90deg), degrees, /2 law, sweep_T */
#define USC_CHIRP_S 1u /* experiments/synchronization/Src/chirp.c:16-44: complex (cos,sin) */
#define USC_CHIRP_T 2u /* experiments/chirp_compression_time_domain/Src/chirp.c:25-45: real cos, no /2 */
#define USC_CHIRP_F 3u /* experiments/chirp_compression_freq_domain/Src/chirp.c:15-35 */

#define USC_HANN_PERIODIC 0u  /* receiver/Src/main.c:99,390-393 */
#define USC_HANN_SYMMETRIC 1u /* experiments/chirp_compression_time_domain/Src/chirp.c:13,63-65 */

#define USC_PCM_F32 0u /* frames already cast to float (the reference's fifo_queue) */
#define USC_PCM_I32 1u /* raw DFSDM words; the cast of receiver/Src/main.c:663-665 is fused in */

#define USC_UP 1   /* receiver/Inc/chirp.h:12-14 */
#define USC_DOWN 0

/* Compile-time constants of the reference gathered in one POD (receiver/Inc/main.h:97-98,
 * receiver/Inc/chirp.h:16-19, receiver/Src/dfsdm.c:59-61,69). */
typedef struct usc_config {
    uint32_t n;             /* NN / PCM_SAMPLES: samples per frame, power of two, 32..65536       */
    float fs;               /* sampling rate as the firmware computes it (78125.0f)               */
    float f0, f1;           /* sweep range F0,F1 (receiver) or F1,F2 (experiments)                */
    float sweep_T;          /* TIME_FRAME (0.0205f) for variants R,S; T,F use n/fs                */
    uint32_t chirp_variant; /* USC_CHIRP_*                                                        */
    uint32_t window;        /* USC_HANN_*                                                         */
    float snr_threshold;    /* SNR_THRESHOLD (2.0f)                                               */
    uint32_t reserved[4];
} usc_config;

typedef struct usc_handle usc_handle;

/* One dsp() result — struct history of receiver/Src/main.c:124-136 without the tick stamps,
 * plus the raw bins the frequencies were derived from. 48 bytes. */
typedef struct usc_history {
    float mag_max, mag_max_left, mag_max_right;
    int32_t max_freq, max_freq_left, max_freq_right;
    uint32_t max_idx, max_idx_left, max_idx_right;
    float mag_mean, snr;
    uint32_t rank; /* the reference's char rank, widened */
} usc_history;

/* ---- lifecycle -------------------------------------------------------------------------------- */
/* Fills cfg with the final receiver's constants (n=2048, fs=78125, 16-19 kHz, variant R, periodic). */
void usc_default_config(usc_config *cfg);
/* Builds every table the config implies on the host in C (Hann via the arm_cos_f32 table algorithm,
 * reference chirps, twiddles), uploads them to `device`.  Replaces arm_rfft_fast_init_f32
 * (receiver/Src/main.c:377), init_ref_chirp (receiver/Src/chirp.c:42-45) and the Hann loop
 * (receiver/Src/main.c:390-393). */
int usc_create(const usc_config *cfg, int device, usc_handle **out);
void usc_destroy(usc_handle *h);
int usc_set_stream(usc_handle *h, void *cuda_stream);
int usc_sync(usc_handle *h);
const char *usc_error_string(int code);
/* geometry derived as in receiver/Src/main.c:372-374 */
int usc_get_geometry(const usc_handle *h, uint32_t *bandwidth, uint32_t *bandwidth2, uint32_t *idx_left_zero);
/* Copies a host-built table back to the caller (what == "hann", "up", "down", "twiddle", "H_up",
 * "H_down"); cap = capacity of dst in floats; returns the table length in floats or <0. */
int usc_get_table(const usc_handle *h, const char *what, float *dst, size_t cap);
/* number of kernel launches issued through this handle since creation (bench evidence) */
uint64_t usc_launch_count(const usc_handle *h);

/* device memory helpers so a plain-C host needs no CUDA headers */
int usc_malloc(void **dptr, size_t bytes);                    /* on the CURRENT device */
int usc_malloc_on(usc_handle *h, void **dptr, size_t bytes);  /* on the handle's device (processes holding several GPUs) */
int usc_free(void *dptr);
int usc_malloc_host(void **hptr, size_t bytes); /* pinned */
int usc_free_host(void *hptr);
int usc_memcpy_h2d(usc_handle *h, void *dst, const void *src, size_t bytes); /* async on the stream */
int usc_memcpy_d2h(usc_handle *h, void *dst, const void *src, size_t bytes); /* async on the stream */
int usc_memset(usc_handle *h, void *dst, int value, size_t bytes);

/* ---- batched CMSIS-shaped operators (device pointers) ------------------------------------------
 * `batch` independent vectors laid out back to back with the given strides (in floats).  A
 * broadcast operand has stride 0.  Each mirrors the arm_math.h function named in its comment.   */
/* (float) buf[i] — the ingest cast, receiver/Src/main.c:663-665 */
int usc_i32_to_f32(usc_handle *h, const int32_t *src, float *dst, size_t count);
/* arm_mult_f32, arm_math.h:1938-1942 */
int usc_arm_mult_f32_batch(usc_handle *h, const float *a, size_t stride_a, const float *b, size_t stride_b,
                           float *dst, size_t stride_dst, uint32_t block_size, uint32_t batch);
/* arm_scale_f32, arm_math.h:2508 */
int usc_arm_scale_f32_batch(usc_handle *h, const float *src, float scale, float *dst, uint32_t block_size,
                            uint32_t batch);
/* arm_cmplx_mult_cmplx_f32, arm_math.h:6579-6583 (num_samples complex; no conjugate) */
int usc_arm_cmplx_mult_cmplx_f32_batch(usc_handle *h, const float *a, size_t stride_a, const float *b,
                                       size_t stride_b, float *dst, size_t stride_dst, uint32_t num_samples,
                                       uint32_t batch);
/* arm_cmplx_mult_real_f32, arm_math.h:6425-6429 */
int usc_arm_cmplx_mult_real_f32_batch(usc_handle *h, const float *cplx, size_t stride_c, const float *real,
                                      size_t stride_r, float *dst, size_t stride_dst, uint32_t num_samples,
                                      uint32_t batch);
/* arm_cmplx_mag_f32, arm_math.h:6312-6315.  src and dst may be disjoint, or the reference's in-place form
 * (receiver/Src/main.c:178 `arm_cmplx_mag_f32(signal, signal, NN)`): dst == src with stride_dst == stride_src >=
 * 2*num_samples (any strides when batch == 1).  Any other overlap returns USC_ERR_ARGUMENT. */
int usc_arm_cmplx_mag_f32_batch(usc_handle *h, const float *src, size_t stride_src, float *dst,
                                size_t stride_dst, uint32_t num_samples, uint32_t batch);
/* arm_max_f32, arm_math.h:6537-6541: first occurrence of the maximum */
int usc_arm_max_f32_batch(usc_handle *h, const float *src, size_t stride_src, uint32_t block_size,
                          float *result, uint32_t *index, uint32_t batch);
/* arm_mean_f32, arm_math.h:6192: sequential sum / block_size */
int usc_arm_mean_f32_batch(usc_handle *h, const float *src, size_t stride_src, uint32_t block_size,
                           float *result, uint32_t batch);
/* arm_rfft_fast_f32, arm_math.h:2246-2249.  fft_len real points per vector (power of two; 32..65536 forward, ..16384 inverse);
 * in == out allowed (hazard H2 is defined away: the result is always the mathematically right one). */
int usc_arm_rfft_fast_f32_batch(usc_handle *h, uint32_t fft_len, const float *in, float *out, uint8_t ifft_flag,
                                uint32_t batch);
/* arm_cfft_f32, arm_math.h:2149-2153 with bitReverseFlag = 1: in place, fft_len complex points
 * (16..32768 forward, ..8192 inverse); forward unscaled, inverse scaled by 1/fft_len.  Lengths
 * beyond CMSIS's 4096 (arm_const_structs.h:49-57) exist for the long-frame sweep (BASELINE config 5). */
int usc_arm_cfft_f32_batch(usc_handle *h, uint32_t fft_len, float *data, uint8_t ifft_flag, uint32_t batch);
/* arm_fir_f32, arm_math.h:1194-1214.  coeffs (host pointer) in CMSIS time-reversed order; state:
 * batch x (num_taps-1) floats on the device carrying the filter history between calls (zero it to
 * start, like arm_fir_init_f32). */
int usc_arm_fir_f32_batch(usc_handle *h, const float *coeffs_host, uint32_t num_taps, float *state,
                          const float *src, float *dst, uint32_t block_size, uint32_t batch);

/* ---- fused stage-level operators ---------------------------------------------------------------- */
/* K1.  dsp() for UP and DOWN on `nframes` aligned frames (receiver/Src/main.c:163-215 twice per
 * frame): de-chirp x Hann x RFFT x |.| x arg-max over bins [0, bandwidth2) in ONE pass over the
 * PCM.  pcm: nframes*n samples of pcm_format.  Outputs (device, nframes each; any may be NULL):
 * peak magnitude and bin per hypothesis, and the symbol decision bit = !(mag_down > mag_up)
 * (receiver/Src/main.c:523: down wins only if strictly greater).  When only one hypothesis' outputs
 * are requested (the other pair and `bit` NULL) only that hypothesis is computed — dsp(.., UP) or
 * dsp(.., DOWN) alone — at half the cost per frame. */
int usc_demod_frames(usc_handle *h, const void *pcm, uint32_t pcm_format, size_t nframes, float *mag_up,
                     uint32_t *idx_up, float *mag_down, uint32_t *idx_down, uint8_t *bit);
/* ---- the same stages on HOST buffers (the call a firmware-style host makes once per capture) -------
 * The input is cut into chunks that flow through three stream lanes (H2D copy, kernel, D2H of the results), so
 * the copies overlap the kernels.  Blocking: results are in the host arrays on return.  pcm_host should be
 * pinned (usc_malloc_host) for the copies to overlap.  usc_host_workspace() sets the chunk size in units of
 * 8 KB of PCM (one 2048-sample frame) and allocates the device staging; 4096 is used otherwise.  The
 * staging grows on demand, never on a steady-state call.
 *   usc_demod_frames_host   K1 / K6: every frame length with a fused kernel (2048 ... 65536); chunks of frames
 *   usc_iq_demod_host       K5 (fused kernel configurations): chunks of whole streams (the FIR state runs along
 *                           a stream); stream s starts at pcm_host + s*stream_stride samples
 *   usc_receiver_run_host   K7: a chunk is a slice of time of ALL streams (one warp serves a stream), resumed
 *                           through the carried per-stream state of usc_receiver_run_chunk; the two frames of
 *                           FIFO history are re-sent with each chunk.  uart / results are host arrays. */
int usc_host_workspace(usc_handle *h, size_t chunk_frames);
int usc_demod_frames_host(usc_handle *h, const void *pcm_host, uint32_t pcm_format, size_t nframes,
                          float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down, uint8_t *bit);
int usc_iq_demod_host(usc_handle *h, const void *pcm_host, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                      size_t stream_stride, float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down,
                      uint8_t *bit);
/* pipeline() of receiver/Src/main.c:163-180 on `batch` frames, one hypothesis: frames (n floats
 * each, stride n) -> n floats each: magnitudes of the n/2 packed bins, zeros above (hazard H1). */
int usc_pipeline(usc_handle *h, const float *frames, float *mags, int updown, uint32_t batch);
/* dsp() of receiver/Src/main.c:183-231 for `batch` streams: stream s reads n samples at
 * fifo + s*fifo_stride + sync_position[s] (fifo_stride >= 3n in the reference), runs the pipeline
 * for hypothesis `updown`, the two windowed arg-max, side choice, idx2freq and SNR against
 * mag_mean[s].  hist: batch usc_history records on the device. */
int usc_dsp(usc_handle *h, const float *fifo, size_t fifo_stride, const uint32_t *sync_position,
            const float *mag_mean, int updown, usc_history *hist, uint32_t batch);
/* Result of the receiver state machine for one stream. 32 bytes. */
typedef struct usc_rx_result {
    uint32_t state;          /* final enum state: 0 IDLE, 1 SYNCHRONIZING, 2 SYNCHRONIZED, 3 DATA_RECEIVING
                                (receiver/Src/main.c:108-111) */
    uint32_t sync_position;  /* final sync_position */
    int32_t lock_frame;      /* frame index of the first transition to SYNCHRONIZED, -1 if none */
    uint32_t lock_position;  /* sync_position chosen then: N/2 + max_idx*N/8 (main.c:483) */
    uint32_t nbytes;         /* UART bytes produced (may exceed the capacity given; excess is dropped) */
    uint32_t frames_seen;
    uint32_t turn, sync_cnt;
} usc_rx_result;

/* K7.  The receiver's whole main loop (receiver/Src/main.c:417-580) over `nstreams` independent
 * streams of `nframes` frames each (stream s starts at pcm + s*stream_stride samples; stride even):
 * 3-frame FIFO, 8-offset sliding-correlation search with the mag_stat noise floor, 3-in-a-row lock,
 * up/down symbol decision, resync, MSB-first bit packing.  uart: nstreams*uart_cap bytes receiving
 * what the firmware would printf (decoded chars, '\n' at the end of each message).  Hazards
 * H1/H3/H4/H5 are defined as in DESIGN.md §3.4. */
int usc_receiver_run(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                     size_t stream_stride, uint8_t *uart, uint32_t uart_cap, usc_rx_result *results);
/* Everything the state machine carries from one frame to the next (state, turn, counters, sync_position,
 * the partial byte, mag_stat[12], the history magnitudes).  Opaque to the caller; all-zero = a receiver
 * that has just started.  160 bytes per stream, in device memory. */
typedef struct usc_rx_state { uint32_t opaque[40]; } usc_rx_state;
/* K7 on a stream that arrives in pieces: the same loop as usc_receiver_run, resumed.  `state` (nstreams
 * records, device memory, zeroed before the first chunk) is read at entry and written back at exit.
 * Each stream's buffer holds `carry_frames` = min(2, frames already processed) frames of history (the
 * receiver's FIFO spans three frames, main.c:659-668) followed by the `nframes` new frames; stream s starts
 * at pcm + s*stream_stride.  uart receives the bytes produced by THIS chunk (results[s].nbytes of them),
 * lock_frame and frames_seen count from the start of the stream.  Any split into chunks gives the same
 * bytes and the same final state as one call over the whole stream. */
int usc_receiver_run_chunk(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                           size_t stream_stride, uint32_t carry_frames, usc_rx_state *state, uint8_t *uart,
                           uint32_t uart_cap, usc_rx_result *results);
int usc_receiver_run_host(usc_handle *h, const void *pcm_host, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                          size_t stream_stride, uint8_t *uart, uint32_t uart_cap, usc_rx_result *results);
/* K4.  The sliding-correlation search grid alone (main.c:447-451) for every frame of every stream:
 * 4 x dsp(UP) at N/2 + (t&1)*N/8 + i*N/4 on the FIFO of frame t, after synchronous addition of
 * `sync_add` (>= 1) frame-aligned FIFOs (misc/Formula.ipynb cell 9).  mag/idx: nstreams*nframes*4. */
int usc_sync_search(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                    size_t stream_stride, uint32_t sync_add, float *mag, uint32_t *idx);
/* One entry of the 4-offset scan history (experiments/chirp_compression_freq_domain/Src/main.c:95-104). */
typedef struct usc_scan_entry {
    float mag_max_right, mag_max_left;
    uint32_t max_idx_right, max_idx_left;
} usc_scan_entry;
/* The 4-offset sliding scan of experiments/chirp_compression_freq_domain/Src/main.c:113-160, 245-251 on
 * `batch` 2n-sample buffers (device floats, stride 2n): for offsets 0, n/4, n/2, 3n/4: x down-chirp
 * (handle variant F or T) x Hann -> in-place RFFT -> n/2 magnitudes -> arm_max over [0, 8*bandwidth) and
 * over the last 8*bandwidth floats of the in-place buffer (which still hold packed-spectrum values, as
 * in the reference).  out: batch x 4 entries. */
int usc_scan4(usc_handle *h, const float *pcm2n, uint32_t batch, usc_scan_entry *out);
/* Synthetic receiver input generated on the device (the step before the path):
 * frame f (global index first_frame + i, so shards on different GPUs are disjoint and reproducible)
 * = the transmitter's up- or down-chirp symbol (chirp_orth, generator/ChirpGenerator.ipynb cell 1 +
 * simulation/signal.py:45-53, amplitude amp) chosen by a counter-based random bit, plus noise of
 * standard deviation noise_sigma, as int32 words with the 24-bit sample in bits 31:8 (x256,
 * receiver/Src/dfsdm.c:78).  Integer-only arithmetic on Philox-4x32-10 keyed by (seed, frame, sample):
 * the oracle regenerates any frame bit for bit on the CPU.  bits (nframes, may be NULL): 1 = up. */
int usc_synth_frames(usc_handle *h, uint64_t seed, uint64_t first_frame, size_t nframes, double amp,
                     double noise_sigma, int32_t *pcm, uint8_t *bits);
/* The same generator fed with the I/Q transmitter's symbols (generator/ChirpGeneratorIQmodulation.ipynb cells 2-9,
 * simulation/IQ_modulation.ipynb cell 4): symbol = amp * cos(2 pi (carrier + sideband * fb(t)) t + phase), fb the
 * baseband chirp -bw/2 .. +bw/2 over one frame (up) or the reverse (down), t = linspace(0, n/fs, n).  sideband +1 and
 * phase -pi/2 is the generator notebook's chirp_iq(); sideband -1 and phase 0 is the simulation's chirp_x_carrier(),
 * the sense usc_iq_demod decides as bit 1 = up.  Input of BASELINE config 3.  Same noise, layout and CPU twin. */
int usc_synth_iq_frames(usc_handle *h, uint64_t seed, uint64_t first_frame, size_t nframes, double carrier_hz,
                        double bw_hz, int sideband, double phase_rad, double amp, double noise_sigma, int32_t *pcm,
                        uint8_t *bits);
/* Whole synthetic receiver streams in the transmitter's frame format (generator/ChirpGenerator.ipynb
 * cell 2): stream g (global index first_stream + s) repeats the pattern
 * lead_in x G (silence), 7 x H (up symbol), 1 x L (down symbol), 8*msg_bytes data symbols (MSB first,
 * H = 1), guard x G; every symbol lasts n samples; the whole stream is delayed by a per-stream start
 * offset in [0, n); noise as in usc_synth_frames.  Offsets and message bytes (printable ASCII) come from
 * Philox keyed by (seed, g), so shards are disjoint and any stream can be regenerated on the CPU by the
 * oracle's twin.  Stream s starts at pcm + s*stream_stride samples (stride even, >= nframes*n).
 * offsets (nstreams) and messages (nstreams*msg_bytes) may be NULL. */
int usc_synth_streams(usc_handle *h, uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes,
                      size_t stream_stride, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard, double amp,
                      double noise_sigma, int32_t *pcm, uint32_t *offsets, uint8_t *messages);
/* Renders a 16-bit track recorded at one rate at another, band-limited, by the exact ratio up/down
 * (fs_out = fs_in * up / down; the transmitter's 44.1 kHz WAV at the receiver's 78.125 kHz is 3125/1764;
 * include/usc_tx.h builds that track).  32-tap Hann-windowed-sinc polyphase FIR, integer phase arithmetic,
 * one FMA per tap in ascending order: the oracle's twin gives the same words.  out[j] = round(y_j) * 256
 * (int32 words in the DFSDM layout, ready for every PCM_I32 entry point), n_out <= ceil(n_in * up / down).
 * in / out are device pointers; up <= 8192. */
int usc_resample_i16_to_pcm(usc_handle *h, const int16_t *in, size_t n_in, uint32_t up, uint32_t down, int32_t *out,
                            size_t n_out);
/* ---- the reference's two earlier detectors, n = 2048 only ------------------------
 * Both share the analyser's front half: (float) pcm x Hann -> RFFT -> magnitude * 1/sqrt(N)
 * (experiments/chirp/Src/main.c:200-216, experiments/ultracom/Src/main.c:115-128), computed for bins
 * [0, 512).  Streams are `nframes` consecutive frames each; stream_stride == nframes*n is required
 * (frames of all streams are contiguous).  Every output pointer may be NULL. */
typedef struct usc_onoff_config {
    float f1_hz, f2_hz;            /* CHIRP_F1 / CHIRP_F2 (17000 / 18000, Inc/main.h:79-80); the counted band is
                                      [F1, F1 + 2 (F2 - F1)] (Src/main.c:372-383) */
    float magnitude_threshold;     /* CHIRP_MAGNITUDE_THRESHOLD 3000 (Inc/main.h:83) */
    float high_frac, low_frac;     /* CHIRP_SIGNAL_THRESHOLD_HIGH / LOW 0.1 / 0.05 (Inc/main.h:86-87) */
    uint32_t frame_start, frame_bit, sync_threshold, sampling_offset;   /* 3, 2, 2, 1 (Inc/main.h:95-98) */
} usc_onoff_config;
void usc_onoff_default_config(usc_onoff_config *cfg);
/* On/off chirp detector: strength[f] = number of band bins above the threshold (Src/main.c:237-242),
 * level[f] = 1 HIGH / -1 LOW / 0 UNKNOWN (:276-286); decode() (:119-198) then runs per stream over the
 * levels: chars (nstreams*cap) receive the byte of every completed frame, nchars the count (may exceed
 * cap), sync_errors the number of "Sync error!" resets.  As in the firmware, a sync error does not
 * reset the bit counter. */
int usc_onoff_detect(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                     const usc_onoff_config *cfg, uint16_t *strength, int8_t *level, uint8_t *chars, uint32_t cap,
                     uint32_t *nchars, uint32_t *sync_errors);
typedef struct usc_fsk_config {
    uint32_t sof_bin, eof_bin, hex0_bin, hex_step;   /* 340, 344, 348, 4 (ultracom/Inc/main.h:84-101) */
    uint32_t tolerance;                              /* TOLERANCE 0 (:109) */
    uint32_t tq_n;                                   /* TQ_N 2 (:120) */
    float magnitude_threshold;                       /* MAGNITUDE_THRESHOLD 5000 (:111) */
} usc_fsk_config;
void usc_fsk_default_config(usc_fsk_config *cfg);
/* FSK 18-tone detector: code[f] = 0xF0 start of frame, 0xF1 end of frame, 0..15 hex digit, 0xFF nothing
 * (ultracom/Src/main.c:130-168; first hit in that order, tolerance window scanned upwards), with the
 * magnitude of the hit bin j and frequency[j + 1] (sic, :137); parser() (:175-236) then runs per stream:
 * a code must repeat tq_n times after its first sighting, start-of-frame arms, nibble pairs make chars.
 * nsof / neof count the accepted start / end markers. */
int usc_fsk_detect(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                   const usc_fsk_config *cfg, uint8_t *code, float *magnitude, float *frequency, uint8_t *chars,
                   uint32_t cap, uint32_t *nchars, uint32_t *nsof, uint32_t *neof);
/* The front half alone: nframes x 512 magnitudes (bins 0..511). */
int usc_band_magnitudes(usc_handle *h, const void *pcm, uint32_t pcm_format, size_t nframes, float *mag);
/* The audio spectrum analyser fft() of experiments/basic/Src/main.c:107-142 (the producer of the
 * reference's captured .raw/.flt/.fft files): PCM x Hann -> RFFT -> magnitude * 1/sqrt(N) -> bins below
 * ac_coupling_hz (FFT_AC_COUPLING_HZ = 1000) forced to 1.0 -> dB = 10*log10 -> arg-max.  Uses the
 * handle's window and fs.  mag/db: nframes*n/2 floats, peak/peak_idx: nframes (any may be NULL). */
int usc_spectrum_analyzer(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nframes, float ac_coupling_hz,
                          float *mag, float *db, float *peak, uint32_t *peak_idx);
/* K5.  I/Q baseband path (experiments/iq_modulation/Src/iq_modem.c:34-66, Src/main.c:117-134, with the
 * semantics of simulation/IQ_modulation.ipynb cells 16-31 and BASELINE config 3's decimation by 2).
 * usc_iq_init builds the carrier tables (init_iq_modem), the baseband reference chirp -bw/2..+bw/2
 * at fs/2 and the n/2-point Hann, and stores the FIR taps (CMSIS time-reversed order, e.g. the 27
 * taps of iq_modem.c:18).  usc_iq_demod, per frame of every stream: x*cos, x*sin -> FIR on each (state
 * carried from the previous frame of the stream, zero at its start) -> every 2nd sample -> R = I + jQ
 * -> up: R x conj(chirp), down: R x chirp -> Hann -> n/2-point complex FFT -> magnitude -> arg-max
 * over window_bins bins on either side of DC (left wins only if strictly greater).  Outputs have
 * nstreams*nframes entries (any may be NULL); bit = !(mag_down > mag_up). */
int usc_iq_init(usc_handle *h, float carrier_hz, float bw_hz, const float *fir_coeffs_host, uint32_t num_taps,
                uint32_t window_bins);
int usc_iq_demod(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                 size_t stream_stride, float *mag_up, uint32_t *idx_up, float *mag_down, uint32_t *idx_down,
                 uint8_t *bit);
/* K2.  compress_chirp() of experiments/chirp_compression_time_domain/Src/chirp.c:78-83 followed
 * by the signed arm_max_f32 over all n lags (.../Src/main.c:189) on `nframes` frames; the handle
 * must be created with chirp_variant T and the symmetric window.  out_frames (nframes*n floats)
 * may be NULL when only the peaks are wanted. */
int usc_compress_chirp(usc_handle *h, const void *pcm, uint32_t pcm_format, size_t nframes, int use_up,
                       float *out_frames, float *max_val, uint32_t *max_idx);
/* K8.  Overlap-save frame synchroniser (BASELINE.json north_star; no counterpart function in the reference, which
 * probes four offsets per frame — usc_sync_search keeps that arithmetic).  compress_chirp()'s three steps
 * (experiments/chirp_compression_time_domain/Src/chirp.c:78-83: RFFT, arm_cmplx_mult_cmplx_f32 by the spectrum of
 * window x reference chirp, inverse RFFT) run on 2n-sample windows advancing by n samples, the template zero-padded to
 * 2n, so that lags [n, 2n) of block b are the LINEAR filter output y[t] = sum_m g[m] x[t - m] at t = (b+1) n ... (b+2) n - 1:
 * every lag of the stream, sample resolution, each PCM sample read from HBM once.  Per stream of nframes frames
 * (stream_stride samples apart, 16-byte aligned) there are nframes-1 blocks; per block: out (n floats, may be NULL),
 * max_val / max_idx = arm_max_f32 over the block's n lags (signed, first occurrence).  use_up selects the up chirp's
 * spectrum as the filter (0 = the down chirp's, the reference's choice in compress_chirp).  n = 2048, real chirp
 * variants (R, T, F). */
int usc_correlate_os(usc_handle *h, const void *pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                     size_t stream_stride, int use_up, float *out, float *max_val, uint32_t *max_idx);

#ifdef __cplusplus
}
#endif
#endif /* USC_H_ */
