// k_long.cu — K6: the receiver chain fused for long frames, N = 2048*R0 with R0 = 4 or 8 (8192- and
// 16384-point frames of BASELINE config 5).  Same reference path as K1 (cast, de-chirp, Hann, RFFT,
// magnitude, arg-max over [0, bandwidth2), both hypotheses; receiver/Src/main.c:163-215), one HBM pass.
//
// The N/2-point complex FFT has the canonical plan [R0, 32, 32]: one radix-R0 level over a
// (m = a + 1024 b), then R0 independent 1024-point transforms whose outputs interleave
// (k = R0 c + d).  One CTA of R0 warps owns a frame:
//   level 0   all threads: gather z[a + 1024 b] straight from global PCM (coalesced across a), apply
//             de-chirp x Hann for both hypotheses (f32x2 halves), radix-R0 butterfly, x W^(a d), park
//             sub-sequence d in shared memory
//   core      warp d runs the packed 32x32 register core of K1 on sub-sequence d; its parked region
//             doubles as the exchange tile once the data is in registers; only c < ceil(bw2/R0) and the
//             partner range near 1024 are produced (pruned last pass)
//   split     the needed sub-spectra meet in shared memory; threads take bins k < bw2, fetch
//             Z[k] = Y_{k mod R0}[k / R0] and Z[n-k], apply the real split, magnitudes, block arg-max
//             (first occurrence, exact: every root is taken — 2*bw2 roots are noise next to the FFT).
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kLongNB = 5;                            // c < 160 covers bandwidth2 / R0 <= 160
constexpr int kLongKeep = 32 * kLongNB;               // kept outputs per side of each sub-spectrum

struct long_params {
    const void* pcm; size_t nframes; uint32_t n;      // n real samples per frame (2048 * R0)
    const float4* chirp_ud; const float2* hann; const float2* tw_master;   // master: (cos, -sin)(2 pi j / n), j < n
    const float2* tw_pass;                            // W_1024^(a d), [d][a]
    uint32_t bandwidth2;
    float* mag_up; uint32_t* idx_up; float* mag_down; uint32_t* idx_down; uint8_t* bit;
};

template <int R0> struct long_smem {
    // per warp d: 16 KB region: sub-sequence d as (re pair[1024], im pair[1024]); later tile (8 KB) + kept outputs
    static constexpr int tw = 0, sub = 8192, region = 16384, red = sub + R0 * region, total = red + R0 * 64;
};

template <typename PCM, int R0>
__global__ void __launch_bounds__(R0 * 32, 1) k_demod_long(long_params p) {
    using L = long_smem<R0>;
    using V2 = typename vec2<PCM>::type;
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw + L::tw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int T = R0 * 32;
    const uint32_t nc = 1024u * R0;                    // complex length
    for (int i = tid; i < 1024; i += T) s_tw[i] = p.tw_pass[i];
    __syncthreads();
    const uint32_t bw2 = p.bandwidth2;

    for (size_t f = blockIdx.x; f < p.nframes; f += gridDim.x) {
        const V2* src = reinterpret_cast<const V2*>(static_cast<const PCM*>(p.pcm) + f * p.n);
        // ---- level 0: radix-R0 over b, twiddle, park sub-sequence d ----
        for (uint32_t a = tid; a < 1024u; a += T) {
            float2 re[R0], im[R0];
#pragma unroll
            for (int b = 0; b < R0; ++b) {
                const uint32_t m = a + 1024u * b;
                const V2 raw = src[m];
                const float x0 = pcm_to_float(raw.x), x1 = pcm_to_float(raw.y);
                const float4 c = __ldg(p.chirp_ud + m);
                const float2 w = __ldg(p.hann + m);
                // window multiply scalar: ptxas contracts packed mul.rn + add.rn into FFMA2 (see k_demod.cu)
                const float2 tr = __fmul2_rn(make_float2(c.x, c.y), bc2(x0)), ti = __fmul2_rn(make_float2(c.z, c.w), bc2(x1));
                re[b] = make_float2(__fmul_rn(tr.x, w.x), __fmul_rn(tr.y, w.x));
                im[b] = make_float2(__fmul_rn(ti.x, w.y), __fmul_rn(ti.y, w.y));
            }
            fft_base2<R0>(re, im);
#pragma unroll
            for (int d = 0; d < R0; ++d) {
                float2 xr = re[d], xi = im[d];
                if (d != 0) {
                    const float2 w = __ldg(p.tw_master + (size_t) a * d * 2u);      // W_nc^(a d) = W_n^(2 a d)
                    cmul2(re[d], im[d], w.x, w.y, xr, xi);
                }
                float2* reg = reinterpret_cast<float2*>(s_raw + L::sub + d * L::region);
                reg[a] = xr;
                reg[1024 + a] = xi;
            }
        }
        __syncthreads();
        // ---- 1024-point packed core on sub-sequence `warp` ----
        float2* reg = reinterpret_cast<float2*>(s_raw + L::sub + warp * L::region);
        float2 re[32], im[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            re[b] = reg[lane + 32 * b];
            im[b] = reg[1024 + lane + 32 * b];
        }
        __syncwarp();
        fft1024_pair(re, im, reg, s_tw, lane);           // first 8 KB of the region is now the exchange tile
        // keep Y_d[c] for c < 160 (elements 0..4) and c >= 864 (elements 27..31)
        float4* keep = reinterpret_cast<float4*>(s_raw + L::sub + warp * L::region + 8192);
#pragma unroll
        for (int j = 0; j < kLongNB; ++j) {
            keep[lane + 32 * j] = make_float4(re[j].x, re[j].y, im[j].x, im[j].y);
            keep[kLongKeep + lane + 32 * j] = make_float4(re[32 - kLongNB + j].x, re[32 - kLongNB + j].y,
                                                          im[32 - kLongNB + j].x, im[32 - kLongNB + j].y);
        }
        __syncthreads();
        // ---- split, magnitude, arg-max over k < bw2 ----
        auto Y = [&](uint32_t k) -> float4 {             // Z[k] = Y_{k mod R0}[k / R0], only kept ranges are asked for
            const uint32_t d = k & (R0 - 1u), c = k / R0;
            const float4* kp = reinterpret_cast<const float4*>(s_raw + L::sub + d * L::region + 8192);
            return c < (uint32_t) kLongKeep ? kp[c] : kp[kLongKeep + (c - (1024u - kLongKeep))];
        };
        float bu = -INFINITY, bd = -INFINITY;
        uint32_t iu = 0xffffffffu, id = 0xffffffffu;
        for (uint32_t k = tid; k < bw2; k += T) {
            const float4 zk = Y(k);
            float2 xr, xi;
            if (k == 0) {
                xr = __fadd2_rn(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w));
                xi = __fadd2_rn(make_float2(zk.x, zk.y), neg2(make_float2(zk.z, zk.w)));
            } else {
                const float4 zc = Y(nc - k);
                const float2 w = __ldg(p.tw_master + k);                          // W_n^k = (cos, -sin)
                rfft_split2(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w), make_float2(zc.x, zc.y),
                            make_float2(zc.z, zc.w), w.x, -w.y, xr, xi);
            }
            const float2 pw = __ffma2_rn(xr, xr, __fmul2_rn(xi, xi));
            const float mu = __fsqrt_rn(pw.x), md = __fsqrt_rn(pw.y);
            if (iu == 0xffffffffu || bu < mu) { bu = mu; iu = k; }
            if (id == 0xffffffffu || bd < md) { bd = md; id = k; }
        }
        warp_argmax(bu, iu);
        warp_argmax(bd, id);
        float* red = reinterpret_cast<float*>(s_raw + L::red);
        if (lane == 0) {
            red[warp * 4 + 0] = bu; reinterpret_cast<uint32_t*>(red)[warp * 4 + 1] = iu;
            red[warp * 4 + 2] = bd; reinterpret_cast<uint32_t*>(red)[warp * 4 + 3] = id;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w2 = 1; w2 < R0; ++w2) {
                argmax_combine(bu, iu, red[w2 * 4 + 0], reinterpret_cast<uint32_t*>(red)[w2 * 4 + 1]);
                argmax_combine(bd, id, red[w2 * 4 + 2], reinterpret_cast<uint32_t*>(red)[w2 * 4 + 3]);
            }
            if (p.mag_up) p.mag_up[f] = bu;
            if (p.idx_up) p.idx_up[f] = iu;
            if (p.mag_down) p.mag_down[f] = bd;
            if (p.idx_down) p.idx_down[f] = id;
            if (p.bit) p.bit[f] = bd > bu ? 0 : 1;
        }
        __syncthreads();
    }
}

template <typename PCM, int R0>
static cudaError_t launch_long_t(const long_params& p, int num_sms, cudaStream_t st) {
    static bool configured = false;
    const int smem = long_smem<R0>::total;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_demod_long<PCM, R0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int per_sm = R0 == 4 ? 2 : 1;
    size_t ctas = p.nframes < (size_t) num_sms * per_sm ? p.nframes : (size_t) num_sms * per_sm;
    k_demod_long<PCM, R0><<<(int) ctas, R0 * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_demod_long(const void* pcm, uint32_t pcm_format, size_t nframes, uint32_t n, const float2* chirp_ud,
                              const float2* hann, const float2* tw_master, const float2* tw_pass, uint32_t bandwidth2,
                              float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down, uint8_t* bit,
                              int num_sms, cudaStream_t st) {
    long_params p{pcm, nframes, n, reinterpret_cast<const float4*>(chirp_ud), hann, tw_master, tw_pass, bandwidth2,
                  mag_up, idx_up, mag_down, idx_down, bit};
    if (n == 8192) return pcm_format == 1u ? launch_long_t<int32_t, 4>(p, num_sms, st) : launch_long_t<float, 4>(p, num_sms, st);
    if (n == 16384) return pcm_format == 1u ? launch_long_t<int32_t, 8>(p, num_sms, st) : launch_long_t<float, 8>(p, num_sms, st);
    return cudaErrorInvalidValue;
}

}  // namespace usc
