// usc_kernels.cuh — device kernels of libusc (sm_100a).  See DESIGN.md §4 for the data layout and
// the roofline of each kernel.  All arithmetic goes through usc_arith.cuh (canonical fp32 order).
#pragma once
#include "usc_arith.cuh"

namespace usc {

struct history_rec {   // == usc_history (include/usc.h)
    float mag_max, mag_max_left, mag_max_right;
    int32_t max_freq, max_freq_left, max_freq_right;
    uint32_t max_idx, max_idx_left, max_idx_right;
    float mag_mean, snr;
    uint32_t rank;
};

__device__ __forceinline__ float pcm_cast(int32_t v) { return __int2float_rn(v); }
__device__ __forceinline__ float pcm_cast(float v) { return v; }

struct rx_result_rec {   // == usc_rx_result (include/usc.h)
    uint32_t state, sync_position;
    int32_t lock_frame;
    uint32_t lock_position, nbytes, frames_seen, turn, sync_cnt;
};

constexpr uint32_t kRxStateMagic = 0x55534337u;      // "USC7"
struct rx_state_rec {    // == usc_rx_state (include/usc.h): everything the state machine carries between frames, 160 bytes
    uint32_t magic, state, turn, sync_cnt, pos, max_idx, msg, msg_cnt;
    int32_t lock_frame;
    uint32_t lock_pos, frames_seen;
    float mag_mean;
    float stat[24];      // mag_stat[12] | history mag_max[8] | history mag_mean[4]
    uint32_t reserved[4];
};
static_assert(sizeof(rx_state_rec) == 160, "usc_rx_state layout");

struct rx_launch {       // arguments of K4 / K7
    const void* pcm; uint32_t pcm_format; uint32_t nstreams; uint32_t nframes; size_t stream_stride;
    const float2* up; const float2* down; const float2* hann; const float2* tw_pass; const float2* tw_split;
    uint32_t bandwidth2; float snr_threshold;
    uint8_t* uart; uint32_t uart_cap; rx_result_rec* results;
    uint32_t sync_add; float* ss_mag; uint32_t* ss_idx;
    uint32_t carry; rx_state_rec* rx_state;              // chunked K7 (usc_receiver_run_chunk)
};

// ------------------------------------------------------------------------------------------------
// element-wise / reduction operators (one HBM pass each; bound: HBM)
// ------------------------------------------------------------------------------------------------
__global__ void k_i32_to_f32(const int32_t* __restrict__ src, float* __restrict__ dst, size_t count);
__global__ void k_mult(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                       uint32_t len, uint32_t batch);
__global__ void k_scale(const float* src, float scale, float* dst, size_t total);
__global__ void k_cmul(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                       uint32_t ncplx, uint32_t batch);
__global__ void k_cmul_real(const float* c, size_t sc, const float* r, size_t sr, float* dst, size_t sd,
                            uint32_t ncplx, uint32_t batch);
__global__ void k_cmag(const float* src, size_t ss, float* dst, size_t sd, uint32_t ncplx, uint32_t batch);
__global__ void k_max(const float* src, size_t ss, uint32_t len, float* result, uint32_t* index, uint32_t batch);
__global__ void k_mean(const float* src, size_t ss, uint32_t len, float* result, uint32_t batch);
__global__ void k_fir(const float* __restrict__ coeffs, uint32_t taps, float* state, const float* src,
                      float* dst, uint32_t len, uint32_t batch);

// ------------------------------------------------------------------------------------------------
// generic canonical FFT: one CTA per transform, data resident in shared memory
// ------------------------------------------------------------------------------------------------
struct fft_plan_dev {
    uint32_t n;          // complex length
    uint32_t nrad;
    uint32_t rad[8];
    uint32_t tw_n;       // master table length
    const float2* tw;    // device master table (cos, -sin)
};
enum fft_mode : int { FFT_C2C_FWD = 0, FFT_C2C_INV = 1, FFT_R2C = 2, FFT_C2R = 3 };
template <int MODE>
__global__ void k_fft_generic(fft_plan_dev plan, const float* in, float* out, uint32_t batch);

// ------------------------------------------------------------------------------------------------
// K1: fused receiver demodulator, N = 2048 (one warp per frame)
// ------------------------------------------------------------------------------------------------
// legacy detectors (k_legacy.cu): window + RFFT + magnitude/sqrt(N) over bins [0, 512), then band count and tone lookup
struct band_params {
    const void* pcm; size_t nframes;
    const float2* hann; const float2* tw_pass; const float2* tw_split;
    float inv_sqrt_n, fs;
    float* mag;                                             // optional: nframes x 512 magnitudes
    // on/off chirp detector
    uint32_t band_lo, band_hi; float onoff_threshold; uint32_t thr_high, thr_low;
    uint16_t* strength; int8_t* level;
    // FSK tone lookup
    uint32_t sof_bin, eof_bin, hex0_bin, hex_step, tolerance; float fsk_threshold;
    uint8_t* code; float* code_mag; float* code_freq;
};

struct demod_params {
    const void* pcm;            // nframes x 2048 samples (or fifo base when gather != 0)
    size_t nframes;
    const float2* chirp_up;     // 1024 float2 = 2048 floats
    const float2* chirp_down;
    const float2* chirp_ud;     // interleaved (up[i], down[i]) pairs, 2048 float2
    const float2* hann;
    const float2* tw_pass;      // [d][a] layout: W_1024^(a*d), 32x32 float2
    const float2* tw_split;     // (cos, sin)(2*pi*k/2048), k < 1024
    const float2* tw_master;    // (cos, -sin)(2*pi*j/2048), j < 2048
    uint32_t bandwidth2;        // arg-max window [0, bandwidth2)
    float* mag_up; uint32_t* idx_up; float* mag_down; uint32_t* idx_down; uint8_t* bit;
    // dsp() mode
    size_t fifo_stride; const uint32_t* sync_position; const float* mag_mean; history_rec* hist;
    uint32_t idx_left_zero; int32_t fs_int; int updown;
};
template <int NB>
__global__ void k_dsp2048(demod_params p);

}  // namespace usc
