// k_compress.cu — K2: fused frequency-domain chirp compression for N = 2048.
//
// Reference path (experiments/chirp_compression_time_domain):
//   arm_copy_f32(pcm, fft_inout)                     Src/main.c:175   (+ the int32->float cast)
//   windowing()          x * symmetric Hann          Src/chirp.c:47-50, 79
//   arm_rfft_fast_f32    forward, packed             Src/chirp.c:80
//   arm_cmplx_mult_cmplx_f32(X, H_down, X, N/2)      Src/chirp.c:81   (packed DC/Nyquist quirk kept)
//   arm_rfft_fast_f32    inverse                     Src/chirp.c:82
//   arm_max_f32 over all N lags (signed)             Src/main.c:189
//
// One warp per frame, same register layout as K1: after the forward 1024-point FFT lane d0 holds
// Z[d0 + 32 d1]; split -> X[k] -> P = X*H -> merge -> 2Z'[k] happen bin by bin with two shuffle
// rounds for the partners (Z[1024-k], P[1024-k]); Z' is already in the layout the next
// fft1024 pass expects, so the inverse transform needs no re-shuffle.  Each frame crosses HBM once
// in (8 KB) and, unless the caller wants the compressed frames, 8 bytes out.
#include "usc_kernels.cuh"
#include "usc_launch.h"

namespace usc {

constexpr int kCWarps = 4;
constexpr int kCTileStride = 33;

__device__ __forceinline__ float c_to_float(int32_t v) { return __int2float_rn(v); }
__device__ __forceinline__ float c_to_float(float v) { return v; }
template <typename T> struct cvec2;
template <> struct cvec2<float> { using type = float2; };
template <> struct cvec2<int32_t> { using type = int2; };

__device__ __forceinline__ void c_fft1024(float (&re)[32], float (&im)[32], float2* tile,
                                          const float2* __restrict__ tw_pass, int lane) {
    fft_base<32>(re, im);
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        float xr = re[d], xi = im[d];
        if (d != 0) {
            float2 w = tw_pass[d * 32 + lane];
            cmul(re[d], im[d], w.x, w.y, xr, xi);
        }
        tile[d * kCTileStride + lane] = make_float2(xr, xi);
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) {
        float2 v = tile[lane * kCTileStride + a];
        re[a] = v.x;
        im[a] = v.y;
    }
    __syncwarp();
    fft_base<32>(re, im);
}

// value of register (lane == 0 ? 32 - d1 : 31 - d1) of lane (32 - lane) & 31: the partner bin 1024 - k
template <int D1>
__device__ __forceinline__ float partner(const float (&v)[32], int lane) {
    float mine = lane == 0 ? v[(32 - D1) & 31] : v[31 - D1];
    return __shfl_sync(0xffffffffu, mine, (32 - lane) & 31);
}

template <int D1>
__device__ __forceinline__ void spectral_step(float (&re)[32], float (&im)[32], float (&pr)[32], float (&pi)[32],
                                              const float2* __restrict__ H, const float2* __restrict__ tw_split,
                                              int lane) {
    // split this lane's bin k = lane + 32*D1 and multiply by H[k]
    float zcr = partner<D1>(re, lane), zci = partner<D1>(im, lane);
    const int k = lane + 32 * D1;
    float xr, xi;
    if (D1 == 0 && lane == 0) {
        xr = __fadd_rn(re[0], im[0]);
        xi = __fsub_rn(re[0], im[0]);
    } else {
        float2 w = tw_split[k];
        rfft_split(re[D1], im[D1], zcr, zci, w.x, w.y, xr, xi);
    }
    float2 h = __ldg(H + k);
    cmul(xr, xi, h.x, h.y, pr[D1], pi[D1]);
    if constexpr (D1 + 1 < 32) spectral_step<D1 + 1>(re, im, pr, pi, H, tw_split, lane);
}

template <int D1>
__device__ __forceinline__ void merge_step(const float (&pr)[32], const float (&pi)[32], float (&re)[32],
                                           float (&im)[32], const float2* __restrict__ tw_split, int lane) {
    float pcr = partner<D1>(pr, lane), pci = partner<D1>(pi, lane);
    const int k = lane + 32 * D1;
    float zr, zi;
    if (D1 == 0 && lane == 0) {
        zr = __fadd_rn(pr[0], pi[0]);
        zi = __fsub_rn(pr[0], pi[0]);
    } else {
        float2 w = tw_split[k];
        rfft_merge(pr[D1], pi[D1], pcr, pci, w.x, w.y, zr, zi);
    }
    re[D1] = zi;          // swap(re, im): the inverse transform is the forward one on swapped parts
    im[D1] = zr;
    if constexpr (D1 + 1 < 32) merge_step<D1 + 1>(pr, pi, re, im, tw_split, lane);
}

struct compress_params {
    const void* pcm; size_t nframes;
    const float2* window; const float2* H; const float2* tw_pass; const float2* tw_split;
    float* out_frames; float* max_val; uint32_t* max_idx;
};

template <typename PCM>
__global__ void __launch_bounds__(kCWarps * 32, 2) k_compress2048(compress_params p) {
    __shared__ float2 s_tw[32 * 32];
    __shared__ float2 s_tile[kCWarps][32 * kCTileStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_tw[i] = p.tw_pass[i];
    __syncthreads();
    using V2 = typename cvec2<PCM>::type;
    const size_t nwarps = (size_t) gridDim.x * kCWarps;
    for (size_t f = (size_t) blockIdx.x * kCWarps + warp; f < p.nframes; f += nwarps) {
        const V2* src = reinterpret_cast<const V2*>(static_cast<const PCM*>(p.pcm) + f * 2048);
        float re[32], im[32], pr[32], pi[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            V2 raw = src[m];
            float2 w = __ldg(p.window + m);
            re[b] = __fmul_rn(c_to_float(raw.x), w.x);
            im[b] = __fmul_rn(c_to_float(raw.y), w.y);
        }
        c_fft1024(re, im, s_tile[warp], s_tw, lane);
        spectral_step<0>(re, im, pr, pi, p.H, p.tw_split, lane);
        merge_step<0>(pr, pi, re, im, p.tw_split, lane);
        c_fft1024(re, im, s_tile[warp], s_tw, lane);
        // swap back and scale by 1/N: a[2m] = z.im/N, a[2m+1] = z.re/N
        const float sc = 1.0f / 2048.0f;
        float best = -INFINITY;
        uint32_t bi = 0xffffffffu;
#pragma unroll
        for (int d1 = 0; d1 < 32; ++d1) {
            const uint32_t m = (uint32_t) lane + 32u * d1;
            float a0 = __fmul_rn(im[d1], sc), a1 = __fmul_rn(re[d1], sc);
            if (p.out_frames) reinterpret_cast<float2*>(p.out_frames + f * 2048)[m] = make_float2(a0, a1);
            if (bi == 0xffffffffu || best < a0) { best = a0; bi = 2 * m; }
            if (best < a1) { best = a1; bi = 2 * m + 1; }
        }
        warp_argmax(best, bi);
        if (lane == 0) {
            if (p.max_val) p.max_val[f] = best;
            if (p.max_idx) p.max_idx[f] = bi;
        }
    }
}

cudaError_t launch_compress2048(const void* pcm, uint32_t pcm_format, size_t nframes, const float2* window,
                                const float2* H, const float2* tw_pass, const float2* tw_split, float* out_frames,
                                float* max_val, uint32_t* max_idx, int num_sms, cudaStream_t st) {
    compress_params p{pcm, nframes, window, H, tw_pass, tw_split, out_frames, max_val, max_idx};
    size_t ctas = (nframes + kCWarps - 1) / kCWarps;
    size_t cap = (size_t) num_sms * 2 * 4;
    if (ctas > cap) ctas = cap;
    if (pcm_format == 1u) k_compress2048<int32_t><<<(int) ctas, kCWarps * 32, 0, st>>>(p);
    else k_compress2048<float><<<(int) ctas, kCWarps * 32, 0, st>>>(p);
    return cudaGetLastError();
}

// tail of pipeline(): packed spectrum (n floats) -> n/2 magnitudes + n/2 zeros, in place, one CTA
// per vector staged through shared memory (receiver/Src/main.c:178 with hazard H1 defined).
__global__ void k_pipeline_tail(float* data, uint32_t n, uint32_t batch) {
    extern __shared__ float s_vec[];
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        float* p = data + (size_t) v * n;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_vec[i] = p[i];
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            p[i] = i < n / 2 ? cmag(s_vec[2 * i], s_vec[2 * i + 1]) : 0.0f;
        __syncthreads();
    }
}

cudaError_t launch_pipeline_tail(float* data, uint32_t n, uint32_t batch, cudaStream_t st) {
    int grid = batch < 148u * 16u ? (int) batch : 148 * 16;
    k_pipeline_tail<<<grid, 256, n * sizeof(float), st>>>(data, n, batch);
    return cudaGetLastError();
}

}  // namespace usc
