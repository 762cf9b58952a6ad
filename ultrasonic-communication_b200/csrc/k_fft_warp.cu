// k_fft_warp.cu — arm_rfft_fast_f32 (2048 points, forward and inverse) and arm_cfft_f32 (1024 points) as
// batched operators on the packed 32x32 register core (the FFT engine of K1/K2 without the receiver around
// it).  Two transforms per warp ride in the f32x2 halves; the pair's 16 KB of input arrive by one TMA bulk
// copy while the previous pair computes; outputs go straight from registers to global memory.  One HBM pass
// in, one out — the generic shared-memory kernel (k_fft_generic.cu) keeps every other length.
//   R2C   x[2m], x[2m+1] -> z[m] -> FFT -> split: lane d0 holds Z[d0 + 32 j]; for j < 16 it fetches the partner
//         Z[1024 - k] from lane 32 - d0 (element 31 - j) and produces BOTH X[k] and X[1024 - k] with the canonical
//         split, so every bin is computed once; packed CMSIS layout (X[0].re, X[N/2].re, X[1], ...)
//   C2R   merge from the staged spectrum (X[m], X[1024 - m] both in shared memory) -> FFT on swapped parts ->
//         swap back, x 1/2048
//   C2C   forward; inverse = swap, forward, swap, x 1/1024
// Arithmetic is the canonical one of DESIGN.md §3: results are bit-identical to the generic kernel and the oracle.
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kFwWarps = 8;
constexpr int kFwTabs = 2 * 8192;                      // pass twiddles | split twiddles
constexpr int kFwWarpBytes = 8192 + 16384;             // XOR-swizzled exchange tile + 2-transform input stage
constexpr int kFwBar = kFwTabs + kFwWarps * kFwWarpBytes;
constexpr int kFwSmem = kFwBar + kFwWarps * 8;

__device__ __forceinline__ float2 shfl2w(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

template <int MODE>
__global__ void __launch_bounds__(kFwWarps * 32, 1) k_fft_warp(const float* __restrict__ in, float* __restrict__ out, size_t batch,
                                                               const float2* __restrict__ tw_pass,
                                                               const float2* __restrict__ tw_split) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw);
    float2* s_ws = reinterpret_cast<float2*>(s_raw + 8192);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = s_raw + kFwTabs + warp * kFwWarpBytes;
    float2* stage = reinterpret_cast<float2*>(wbase);                 // 2 x 1024 float2
    float2* tile = reinterpret_cast<float2*>(wbase + 16384);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + kFwBar) + warp;

    const size_t npairs = (batch + 1) / 2;
    const size_t nwarps = (size_t) gridDim.x * kFwWarps;
    size_t q = (size_t) blockIdx.x * kFwWarps + warp;
    auto pair_bytes = [&](size_t pr) -> uint32_t { return 2 * pr + 1 < batch ? 16384u : 8192u; };
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (q < npairs) {
            mbar_expect_tx(bar, pair_bytes(q));
            bulk_g2s(stage, in + q * 4096, pair_bytes(q), bar);
        }
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = tw_pass[i];
        if (MODE == FFT_R2C || MODE == FFT_C2R) s_ws[i] = tw_split[i];
    }
    __syncthreads();

    uint32_t parity = 0;
    for (; q < npairs; q += nwarps) {
        const bool two = 2 * q + 1 < batch;
        mbar_wait(bar, parity);
        parity ^= 1u;
        const float2* sa = stage;
        const float2* sb = two ? stage + 1024 : stage;
        float2 re[32], im[32];                                        // (.x, .y) = (transform 2q, transform 2q+1)
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            const float2 xa = sa[m], xb = sb[m];
            if (MODE == FFT_R2C || MODE == FFT_C2C_FWD) {
                re[b] = make_float2(xa.x, xb.x);
                im[b] = make_float2(xa.y, xb.y);
            } else if (MODE == FFT_C2C_INV) {                         // swapped parts
                re[b] = make_float2(xa.y, xb.y);
                im[b] = make_float2(xa.x, xb.x);
            } else {                                                  // C2R: merge X[m] with X[1024 - m]
                const int mc = (1024 - m) & 1023;                     // m = 0 pairs with itself (packed DC / Nyquist)
                const float2 ca = sa[mc], cb = sb[mc];
                const float2 pkr = make_float2(xa.x, xb.x), pki = make_float2(xa.y, xb.y);
                float2 zr, zi;
                rfft_merge2(pkr, pki, make_float2(ca.x, cb.x), make_float2(ca.y, cb.y), s_ws[m].x, s_ws[m].y, zr, zi);
                if (b == 0) {
                    const float2 dr = __fadd2_rn(pkr, pki), di = __fadd2_rn(pkr, neg2(pki));
                    zr = lane == 0 ? dr : zr;
                    zi = lane == 0 ? di : zi;
                }
                re[b] = zi;                                           // swapped for the forward-on-swapped inverse
                im[b] = zr;
            }
        }
        __syncwarp();
        if (lane == 0 && q + nwarps < npairs) {
            mbar_expect_tx(bar, pair_bytes(q + nwarps));
            bulk_g2s(stage, in + (q + nwarps) * 4096, pair_bytes(q + nwarps), bar);
        }
        fft1024_pair(re, im, tile, s_tw, lane);
        float2* oa = reinterpret_cast<float2*>(out + (2 * q) * 2048);
        float2* ob = reinterpret_cast<float2*>(out + (2 * q + 1) * 2048);
        if (MODE == FFT_C2C_FWD) {
#pragma unroll
            for (int d1 = 0; d1 < 32; ++d1) {
                const int k = lane + 32 * d1;
                oa[k] = make_float2(re[d1].x, im[d1].x);
                if (two) ob[k] = make_float2(re[d1].y, im[d1].y);
            }
        } else if (MODE == FFT_C2C_INV || MODE == FFT_C2R) {
            const float sc = MODE == FFT_C2R ? 1.0f / 2048.0f : 1.0f / 1024.0f;
#pragma unroll
            for (int d1 = 0; d1 < 32; ++d1) {
                const int k = lane + 32 * d1;
                const float2 a0 = __fmul2_rn(im[d1], bc2(sc)), a1 = __fmul2_rn(re[d1], bc2(sc));    // swap back, scale
                oa[k] = make_float2(a0.x, a1.x);
                if (two) ob[k] = make_float2(a0.y, a1.y);
            }
        } else {                                                      // R2C: split, packed layout
            const int src = (32 - lane) & 31;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 sr = lane == 0 ? re[(32 - j) & 31] : re[31 - j];
                const float2 si = lane == 0 ? im[(32 - j) & 31] : im[31 - j];
                const float2 zcr = shfl2w(sr, src), zci = shfl2w(si, src);
                const int k = lane + 32 * j, kc = 1024 - k;           // kc == 1024 only for k == 0
                float2 xr, xi, yr, yi;
                rfft_split2(re[j], im[j], zcr, zci, s_ws[k].x, s_ws[k].y, xr, xi);
                rfft_split2(zcr, zci, re[j], im[j], s_ws[kc & 1023].x, s_ws[kc & 1023].y, yr, yi);
                if (j == 0) {                                         // packed bin 0 = (X[0], X[N/2]) on lane 0
                    const float2 dr = __fadd2_rn(re[0], im[0]), di = __fadd2_rn(re[0], neg2(im[0]));
                    xr = lane == 0 ? dr : xr;
                    xi = lane == 0 ? di : xi;
                }
                oa[k] = make_float2(xr.x, xi.x);
                if (two) ob[k] = make_float2(xr.y, xi.y);
                if (kc < 1024) {
                    oa[kc] = make_float2(yr.x, yi.x);
                    if (two) ob[kc] = make_float2(yr.y, yi.y);
                }
            }
            {   // bin 512 pairs with itself: lane 0, element 16
                float2 xr, xi;
                rfft_split2(re[16], im[16], re[16], im[16], s_ws[512].x, s_ws[512].y, xr, xi);
                if (lane == 0) {
                    oa[512] = make_float2(xr.x, xi.x);
                    if (two) ob[512] = make_float2(xr.y, xi.y);
                }
            }
        }
        __syncwarp();
    }
}

// arm_cfft_f32 at 2048 points (arm_cfft_sR_f32_len2048, the length of experiments/synchronization): canonical plan
// [2, 32, 32] — one radix-2 stage (z[a] +- z[a + 1024], odd half x W_2048^a) feeds two 1024-point transforms that
// produce the even and the odd bins; the two ride in the f32x2 halves, so a warp carries ONE transform per pass.
template <bool INVERSE>
__global__ void __launch_bounds__(kFwWarps * 32, 1) k_cfft2048_warp(const float* __restrict__ in, float* __restrict__ out, size_t batch,
                                                                    const float2* __restrict__ tw_pass,
                                                                    const float2* __restrict__ tw_master /* W_2048^a, a < 1024 */) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw);
    float2* s_tw0 = reinterpret_cast<float2*>(s_raw + 8192);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = s_raw + kFwTabs + warp * kFwWarpBytes;
    float2* stage = reinterpret_cast<float2*>(wbase);                 // 2048 float2
    float2* tile = reinterpret_cast<float2*>(wbase + 16384);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + kFwBar) + warp;
    const size_t nwarps = (size_t) gridDim.x * kFwWarps;
    size_t q = (size_t) blockIdx.x * kFwWarps + warp;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (q < batch) {
            mbar_expect_tx(bar, 16384u);
            bulk_g2s(stage, in + q * 4096, 16384u, bar);
        }
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = tw_pass[i];
        s_tw0[i] = tw_master[i];
    }
    __syncthreads();
    uint32_t parity = 0;
    for (; q < batch; q += nwarps) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 re[32], im[32];                                        // (.x, .y) = (even-bin transform, odd-bin transform)
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int a = lane + 32 * b;
            float2 lo = stage[a], hi = stage[a + 1024];
            if (INVERSE) { lo = make_float2(lo.y, lo.x); hi = make_float2(hi.y, hi.x); }
            const float er = __fadd_rn(lo.x, hi.x), ei = __fadd_rn(lo.y, hi.y);     // radix-2, d = 0
            const float dr = __fsub_rn(lo.x, hi.x), di = __fsub_rn(lo.y, hi.y);     // d = 1, then x W_2048^a
            const float2 w = s_tw0[a];
            float orr, oii;
            cmul(dr, di, w.x, w.y, orr, oii);
            re[b] = make_float2(er, orr);
            im[b] = make_float2(ei, oii);
        }
        __syncwarp();
        if (lane == 0 && q + nwarps < batch) {
            mbar_expect_tx(bar, 16384u);
            bulk_g2s(stage, in + (q + nwarps) * 4096, 16384u, bar);
        }
        fft1024_pair(re, im, tile, s_tw, lane);
        float4* o = reinterpret_cast<float4*>(out + q * 4096);
#pragma unroll
        for (int d1 = 0; d1 < 32; ++d1) {
            const int c = lane + 32 * d1;                             // X[2c] and X[2c + 1]
            if (INVERSE) {
                const float sc = 1.0f / 2048.0f;
                const float2 a0 = __fmul2_rn(im[d1], bc2(sc)), a1 = __fmul2_rn(re[d1], bc2(sc));   // swap back, scale
                o[c] = make_float4(a0.x, a1.x, a0.y, a1.y);
            } else {
                o[c] = make_float4(re[d1].x, im[d1].x, re[d1].y, im[d1].y);
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_cfft2048_warp(bool inverse, float* data, size_t batch, const float2* tw_pass, const float2* tw_master,
                                 int num_sms, cudaStream_t st) {
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_cfft2048_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cfft2048_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    size_t ctas = (batch + kFwWarps - 1) / kFwWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    if (inverse) k_cfft2048_warp<true><<<(int) ctas, kFwWarps * 32, kFwSmem, st>>>(data, data, batch, tw_pass, tw_master);
    else k_cfft2048_warp<false><<<(int) ctas, kFwWarps * 32, kFwSmem, st>>>(data, data, batch, tw_pass, tw_master);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_mode(const float* in, float* out, size_t batch, const float2* tw_pass, const float2* tw_split,
                               int num_sms, cudaStream_t st) {
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_fft_warp<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    size_t ctas = ((batch + 1) / 2 + kFwWarps - 1) / kFwWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    k_fft_warp<MODE><<<(int) ctas, kFwWarps * 32, kFwSmem, st>>>(in, out, batch, tw_pass, tw_split);
    return cudaGetLastError();
}

cudaError_t launch_fft_warp(int mode, const float* in, float* out, size_t batch, const float2* tw_pass, const float2* tw_split,
                            int num_sms, cudaStream_t st) {
    switch (mode) {
    case FFT_R2C: return launch_mode<FFT_R2C>(in, out, batch, tw_pass, tw_split, num_sms, st);
    case FFT_C2R: return launch_mode<FFT_C2R>(in, out, batch, tw_pass, tw_split, num_sms, st);
    case FFT_C2C_FWD: return launch_mode<FFT_C2C_FWD>(in, out, batch, tw_pass, tw_split, num_sms, st);
    case FFT_C2C_INV: return launch_mode<FFT_C2C_INV>(in, out, batch, tw_pass, tw_split, num_sms, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace usc
