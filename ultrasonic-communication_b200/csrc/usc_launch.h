// usc_launch.h — host-side launchers (one per kernel family), called by the C-ABI layer usc_api.cu.
#pragma once
#include <cuda_runtime.h>
#include "usc_kernels.cuh"

namespace usc {

// One-time per-kernel launch configuration (opt-in shared memory size, cluster occupancy) is a property of
// the (kernel, device) pair: a process that opens handles on several GPUs must configure each of them.
template <typename T> struct per_device {
    T v[64] = {};
    T& get() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
        return v[d];
    }
};

cudaError_t launch_i32_to_f32(const int32_t* src, float* dst, size_t count, cudaStream_t st);
cudaError_t launch_spectrum_tail(const float* spec, uint32_t n, float inv_sqrt_n, uint32_t ac_bins, float* mag, float* db,
                                 float* peak, uint32_t* peak_idx, uint32_t batch, cudaStream_t st);
cudaError_t launch_scan_pack(const float* mr, const uint32_t* ir, const float* ml, const uint32_t* il, uint32_t bw8,
                             float* out, uint32_t slot, uint32_t batch, cudaStream_t st);
cudaError_t launch_prep(const void* pcm, uint32_t pcm_format, const float* chirp, const float* hann, float* dst, uint32_t n,
                        size_t total, cudaStream_t st);
cudaError_t launch_mag_max(const float* spec, size_t stride, uint32_t window, float* result, uint32_t* index, uint32_t batch,
                           cudaStream_t st);
cudaError_t launch_decide(const float* mu, const float* md, uint8_t* bit, size_t n, cudaStream_t st);
cudaError_t launch_mult(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                        uint32_t len, uint32_t batch, cudaStream_t st);
cudaError_t launch_scale(const float* src, float scale, float* dst, size_t total, cudaStream_t st);
cudaError_t launch_cmul(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                        uint32_t ncplx, uint32_t batch, cudaStream_t st);
cudaError_t launch_cmul_real(const float* c, size_t sc, const float* r, size_t sr, float* dst, size_t sd,
                             uint32_t ncplx, uint32_t batch, cudaStream_t st);
cudaError_t launch_cmag(const float* src, size_t ss, float* dst, size_t sd, uint32_t ncplx, uint32_t batch,
                        cudaStream_t st);
cudaError_t launch_max(const float* src, size_t ss, uint32_t len, float* result, uint32_t* index,
                       uint32_t batch, cudaStream_t st);
cudaError_t launch_mean(const float* src, size_t ss, uint32_t len, float* result, uint32_t batch,
                        cudaStream_t st);
cudaError_t launch_fir(const float* coeffs_dev, uint32_t taps, float* state, const float* src, float* dst,
                       uint32_t len, uint32_t batch, cudaStream_t st);
// mode: fft_mode.  in/out: batch vectors of 2n floats (C2C) or 2n floats real (R2C/C2R, n = N/2).
cudaError_t launch_fft_generic(int mode, const fft_plan_dev& plan, const float* in, float* out, uint32_t batch,
                               cudaStream_t st);
cudaError_t launch_fft_large(int mode, const fft_plan_dev& plan, const float* in, float* out, float* work,
                             uint32_t batch, cudaStream_t st);
cudaError_t fft_generic_prepare();   // opt in to large dynamic shared memory (once)
cudaError_t launch_demod2048(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st);
cudaError_t launch_demod2048_single(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st);
cudaError_t launch_dsp2048(const demod_params& p, int num_sms, cudaStream_t st);
cudaError_t launch_demod_long32(const void* pcm, uint32_t pcm_format, size_t nframes, uint32_t n, const float2* chirp_ud,
                                const float2* hann, const float2* tw_master, const float2* tw_pass, const float2* tw_l0,
                                uint32_t bandwidth2, float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                                uint8_t* bit, int num_sms, cudaStream_t st);
cudaError_t launch_demod_long(const void* pcm, uint32_t pcm_format, size_t nframes, uint32_t n, const float2* chirp_ud,
                              const float2* hann, const float2* tw_master, const float2* tw_pass, uint32_t bandwidth2,
                              float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down, uint8_t* bit,
                              int num_sms, cudaStream_t st);
cudaError_t launch_dsp2048c(const demod_params& p, int num_sms, cudaStream_t st);
cudaError_t launch_compress2048(const void* pcm, uint32_t pcm_format, size_t nframes, const float2* window,
                                const float2* H, const float2* tw_pass, const float2* tw_split, float* out_frames,
                                float* max_val, uint32_t* max_idx, int num_sms, cudaStream_t st);
cudaError_t launch_correlate_os(const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                                const float2* G, const float2* tw_pass, const float2* tw0, const float2* tw_split, float* out,
                                float* max_val, uint32_t* max_idx, int num_sms, cudaStream_t st);
cudaError_t launch_receiver_run(const rx_launch& a, int num_sms, cudaStream_t st);
cudaError_t launch_sync_search(const rx_launch& a, int num_sms, cudaStream_t st);
cudaError_t launch_iq_frontend(const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                               size_t stream_stride, uint32_t n, const float* car_cos, const float* car_sin,
                               const float* taps, uint32_t ntaps, float* out, cudaStream_t st);
cudaError_t launch_iq_backend(const float* R, size_t nframes, const float* chirp, const float* hann, const float2* tw_pass,
                              uint32_t window, float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                              uint8_t* bit, int num_sms, cudaStream_t st);
cudaError_t launch_iq_fused(const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                            const float* car_cos, const float* car_sin, const float* taps, uint32_t ntaps, const float* chirp,
                            const float* hann, const float2* tw_pass, uint32_t window, float* mag_up, uint32_t* idx_up,
                            float* mag_down, uint32_t* idx_down, uint8_t* bit, int num_sms, cudaStream_t st);
cudaError_t launch_iq_pick(const float* mr, const uint32_t* ir, const float* ml, const uint32_t* il, uint32_t left0,
                           float* mag, uint32_t* idx, size_t count, cudaStream_t st);
cudaError_t launch_band2048(const band_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st);
cudaError_t launch_onoff_decode(const int8_t* level, uint32_t nstreams, uint32_t nframes, uint32_t frame_start, uint32_t frame_bit,
                                uint32_t sync_threshold, uint32_t sampling_offset, uint8_t* chars, uint32_t cap, uint32_t* nchars,
                                uint32_t* sync_errors, cudaStream_t st);
cudaError_t launch_fsk_parse(const uint8_t* code, uint32_t nstreams, uint32_t nframes, uint32_t tq_n, uint8_t* chars, uint32_t cap,
                             uint32_t* nchars, uint32_t* nsof, uint32_t* neof, cudaStream_t st);
cudaError_t launch_resample_i16(const int16_t* in, size_t n_in, uint32_t up, uint32_t down, uint32_t ktaps, const float* taps,
                                int32_t* out, size_t n_out, cudaStream_t st);
cudaError_t launch_cfft2048_warp(bool inverse, float* data, size_t batch, const float2* tw_pass, const float2* tw_master,
                                 int num_sms, cudaStream_t st);
cudaError_t launch_fft_warp(int mode, const float* in, float* out, size_t batch, const float2* tw_pass, const float2* tw_split,
                            int num_sms, cudaStream_t st);
cudaError_t launch_synth_streams(uint64_t seed, uint64_t first_stream, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                                 uint32_t n, uint32_t lead_in, uint32_t msg_bytes, uint32_t guard, const int32_t* table,
                                 int32_t gain, int32_t* pcm, uint32_t* offsets, uint8_t* messages, cudaStream_t st);
cudaError_t launch_synth_frames(uint64_t seed, uint64_t first_frame, size_t nframes, uint32_t n, const int32_t* table,
                                int32_t gain, int32_t* pcm, uint8_t* bits, cudaStream_t st);
cudaError_t launch_pipeline_tail(float* data, uint32_t n, uint32_t batch, int zero_upper, cudaStream_t st);
}  // namespace usc
