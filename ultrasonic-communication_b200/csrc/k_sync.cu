// k_sync.cu — K3: dsp() of the complex-FFT variant (experiments/synchronization/Src/main.c:135-213)
// for N = 2048, the variant whose LEFT window (negative frequencies) is meaningful.
//
//   time_frame[2i] = fifo[pos + i], time_frame[2i+1] = 0          main.c:175-180
//   arm_cmplx_mult_cmplx_f32(frame, up|down_chirp (cos,sin))      chirp.c:51-57
//   arm_cmplx_mult_real_f32(frame, hann_window)                   main.c:150
//   arm_cfft_f32(&arm_cfft_sR_f32_len2048, frame, 0, 1)           main.c:153
//   arm_cmplx_mag_f32                                             main.c:156
//   arm_max_f32 over [idx_left_zero, N) and [0, bandwidth2)       main.c:188-198
//
// One warp per call.  The canonical plan of a 2048-point complex FFT is [2,32,32]: one radix-2
// stage (z[a] +- z[a+1024], odd half x W_2048^a) feeding two 1024-point transforms that produce the
// even and odd output bins.  The two transforms ride in the halves of f32x2 registers through the
// same packed 32x32 core as K1.  Only c = k>>1 in [0, ceil(bw2/2)) and [1024 - bw2/2, 1024) are
// needed, so the last pass is pruned to 6 of 32 outputs per lane by dead-code elimination.
// The imaginary input is exactly zero, so (x + j0)(c + js) is evaluated as (x*c, x*s): it differs
// from the 4-multiply form only in the sign of exact zeros, which cannot change any magnitude.
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kSyWarps = 4;
constexpr int kSyNB = 3;                              // c < 96 (bandwidth2 <= 192)
constexpr int kSySmem = 16384 + 8192 + 8192 + 8192 + kSyWarps * 16384;

__global__ void __launch_bounds__(kSyWarps * 32, 2) k_dsp2048c(demod_params p) {
    extern __shared__ __align__(16) unsigned char s_sy[];
    float2* s_chirp = reinterpret_cast<float2*>(s_sy);                   // 2048 (cos, sin)
    float* s_hann = reinterpret_cast<float*>(s_sy + 16384);              // 2048
    float2* s_tw0 = reinterpret_cast<float2*>(s_sy + 16384 + 8192);      // W_2048^a, a < 1024
    float2* s_tw = reinterpret_cast<float2*>(s_sy + 16384 + 16384);      // W_1024^(a d) [d][a]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* tile = reinterpret_cast<float4*>(s_sy + 16384 + 24576) + warp * 1024;
    const float2* chirp = p.updown ? p.chirp_up : p.chirp_down;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
        s_chirp[i] = chirp[i];
        s_hann[i] = reinterpret_cast<const float*>(p.hann)[i];
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw0[i] = p.tw_master[i];
        s_tw[i] = p.tw_pass[i];
    }
    __syncthreads();
    const uint32_t bw2 = p.bandwidth2, left0 = p.idx_left_zero;
    const size_t nwarps = (size_t) gridDim.x * kSyWarps;
    for (size_t s = (size_t) blockIdx.x * kSyWarps + warp; s < p.nframes; s += nwarps) {
        const float* src = static_cast<const float*>(p.pcm) + s * p.fifo_stride + p.sync_position[s];
        float2 re[32], im[32];                        // (.x, .y) = (even-bin transform, odd-bin transform)
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int a = lane + 32 * b;
            const float xl = src[a], xh = src[a + 1024];
            const float2 cl = s_chirp[a], ch = s_chirp[a + 1024];
            const float wl = s_hann[a], wh = s_hann[a + 1024];
            const float lr = __fmul_rn(__fmul_rn(xl, cl.x), wl), li = __fmul_rn(__fmul_rn(xl, cl.y), wl);
            const float hr = __fmul_rn(__fmul_rn(xh, ch.x), wh), hi = __fmul_rn(__fmul_rn(xh, ch.y), wh);
            const float er = __fadd_rn(lr, hr), ei = __fadd_rn(li, hi);            // radix-2, d = 0
            const float dr = __fsub_rn(lr, hr), di = __fsub_rn(li, hi);            // d = 1, then x W_2048^a
            const float2 w = s_tw0[a];
            float orr, oii;
            cmul(dr, di, w.x, w.y, orr, oii);
            re[b] = make_float2(er, orr);
            im[b] = make_float2(ei, oii);
        }
        fft1024_warp2(re, im, tile, s_tw, lane);
        // candidates: right window k = 2c + d < bw2 with c = lane + 32 d1 (d1 < 3); left window
        // k >= left0 with c = lane + 32 (29 + j).  Ascending k within a lane in both lists.
        float pr[2 * kSyNB], pl[2 * kSyNB];
        uint32_t kr[2 * kSyNB], kl[2 * kSyNB];
        bool okr[2 * kSyNB], okl[2 * kSyNB];
#pragma unroll
        for (int j = 0; j < kSyNB; ++j) {
            const uint32_t c = (uint32_t) lane + 32u * j, cL = (uint32_t) lane + 32u * (32 - kSyNB + j);
            pr[2 * j] = __fmaf_rn(re[j].x, re[j].x, __fmul_rn(im[j].x, im[j].x));
            pr[2 * j + 1] = __fmaf_rn(re[j].y, re[j].y, __fmul_rn(im[j].y, im[j].y));
            kr[2 * j] = 2 * c; kr[2 * j + 1] = 2 * c + 1;
            okr[2 * j] = kr[2 * j] < bw2; okr[2 * j + 1] = kr[2 * j + 1] < bw2;
            const int jj = 32 - kSyNB + j;
            pl[2 * j] = __fmaf_rn(re[jj].x, re[jj].x, __fmul_rn(im[jj].x, im[jj].x));
            pl[2 * j + 1] = __fmaf_rn(re[jj].y, re[jj].y, __fmul_rn(im[jj].y, im[jj].y));
            kl[2 * j] = 2 * cL; kl[2 * j + 1] = 2 * cL + 1;
            okl[2 * j] = kl[2 * j] >= left0; okl[2 * j + 1] = kl[2 * j + 1] >= left0;
        }
        float mr, ml;
        uint32_t ir, il;
        argmax_exact<2 * kSyNB>(pr, kr, okr, mr, ir);
        argmax_exact<2 * kSyNB>(pl, kl, okl, ml, il);
        if (lane == 0) {
            float mm = mr;
            uint32_t im_ = ir;
            if (ml > mr) { mm = ml; im_ = il; }                                    // main.c:191-197
            auto idx2freq = [&](uint32_t idx) -> int32_t {                         // main.c:135-141
                if (idx < 1024u) return (int32_t) ((uint32_t) p.fs_int * idx / 2048u);
                return (int32_t) ((uint32_t) p.fs_int * (2048u - idx) / 2048u) * -1;
            };
            const float mean = p.mag_mean[s];
            history_rec h;
            h.mag_max = mm; h.mag_max_left = ml; h.mag_max_right = mr;
            h.max_idx = im_; h.max_idx_left = il; h.max_idx_right = ir;
            h.max_freq = idx2freq(im_); h.max_freq_left = idx2freq(il); h.max_freq_right = idx2freq(ir);
            h.mag_mean = mean;
            h.snr = __fdiv_rn(__fsub_rn(mm, mean), mean);
            h.rank = (uint32_t) '-';
            p.hist[s] = h;
        }
    }
}

cudaError_t launch_dsp2048c(const demod_params& p, int num_sms, cudaStream_t st) {
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_dsp2048c, cudaFuncAttributeMaxDynamicSharedMemorySize, kSySmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    size_t ctas = (p.nframes + kSyWarps - 1) / kSyWarps;
    const size_t cap = (size_t) num_sms * 2 * 2;
    if (ctas > cap) ctas = cap;
    k_dsp2048c<<<(int) ctas, kSyWarps * 32, kSySmem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace usc
