// k_demod.cu — K1: the fused receiver demodulator for N = 2048.
//
// Reference path fused here (one HBM pass over the PCM, 8 or 16 bytes written per frame):
//   (float) buf[i]                                   receiver/Src/main.c:663-665
//   mult_ref_chirp  : x * up|down_chirp              receiver/Src/chirp.c:47-53
//   arm_mult_f32    : .. * hann_window               receiver/Src/main.c:171
//   arm_rfft_fast_f32 (2048 real)                    receiver/Src/main.c:174
//   arm_cmplx_mag_f32                                receiver/Src/main.c:178
//   arm_max_f32 over bins [0, bandwidth2)            receiver/Src/main.c:208   (+ :206,:209-215)
//
// Mapping: ONE WARP PER FRAME.  The 2048-point real FFT is a 1024-point complex FFT of
// z[m] = a[2m] + j a[2m+1] in the canonical [32,32] plan:
//   pass 1  lane a holds z[a + 32 b], b = 0..31, in registers -> 32-point FFT in registers
//           -> multiply element d by W_1024^(a d)
//   exchange through a padded 32x33 float2 tile in shared memory (conflict-free both ways)
//   pass 2  lane d0 holds V_a[d0], a = 0..31 -> 32-point FFT -> Z[d0 + 32 d1] in register d1
//   split   only bins k = d0 + 32 d1 < bandwidth2 (d1 < NB) are needed; their partners
//           Z[1024 - k] sit in lane (32 - d0) & 31 at register 31 - d1 (32 - d1 on lane 0):
//           one shuffle pair per bin.  The pass-2 outputs nobody reads (d1 in [NB, 32 - NB))
//           are dead code, so the compiler prunes the last FFT stages.
// Both hypotheses (up / down) share the PCM load and the Hann load.
#include "usc_kernels.cuh"
#include "usc_launch.h"

namespace usc {

constexpr int kWarpsPerCta = 4;
constexpr int kTileStride = 33;                       // float2 units, 32x33 padded tile
constexpr int kTileFloat2 = 32 * kTileStride;

__device__ __forceinline__ float pcm_to_float(int32_t v) { return __int2float_rn(v); }
__device__ __forceinline__ float pcm_to_float(float v) { return v; }

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<int32_t> { using type = int2; };

// pass 1 + twiddle + exchange + pass 2 for one hypothesis; on return lane d0 holds Z[d0 + 32*d1].
__device__ __forceinline__ void fft1024_warp(float (&re)[32], float (&im)[32], float2* tile,
                                             const float2* __restrict__ tw_pass, int lane) {
    fft_base<32>(re, im);
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        float xr = re[d], xi = im[d];
        if (d != 0) {
            float2 w = tw_pass[d * 32 + lane];
            cmul(re[d], im[d], w.x, w.y, xr, xi);
        }
        tile[d * kTileStride + lane] = make_float2(xr, xi);
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) {
        float2 v = tile[lane * kTileStride + a];
        re[a] = v.x;
        im[a] = v.y;
    }
    __syncwarp();
    fft_base<32>(re, im);
}

// split + magnitude + per-lane running arg-max over this lane's bins k = lane + 32*d1, d1 < NB.
template <int NB>
__device__ __forceinline__ void peak_right(const float (&re)[32], const float (&im)[32],
                                           const float2 (&ws)[NB], int lane, uint32_t bw2,
                                           float& best, uint32_t& best_idx) {
    best = -INFINITY;
    best_idx = 0xffffffffu;
    const int src = (32 - lane) & 31;
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        // partner Z[1024 - k]: provided by lane `src`; lane 0 serves itself from register 32 - d1
        float sr = lane == 0 ? re[(32 - d1) & 31] : re[31 - d1];
        float si = lane == 0 ? im[(32 - d1) & 31] : im[31 - d1];
        float zcr = __shfl_sync(0xffffffffu, sr, src);
        float zci = __shfl_sync(0xffffffffu, si, src);
        const uint32_t k = (uint32_t) lane + 32u * d1;
        float xr, xi;
        if (d1 == 0 && lane == 0) {          // packed bin 0 = (X[0], X[N/2])  [arm_math.h:2246-2249]
            xr = __fadd_rn(re[0], im[0]);
            xi = __fsub_rn(re[0], im[0]);
        } else {
            rfft_split(re[d1], im[d1], zcr, zci, ws[d1].x, ws[d1].y, xr, xi);
        }
        float m = cmag(xr, xi);
        if (k < bw2 && (best < m || best_idx == 0xffffffffu)) { best = m; best_idx = k; }
    }
}

template <typename PCM, int NB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_demod2048(demod_params p) {
    __shared__ float2 s_tw[32 * 32];
    __shared__ float2 s_tile[kWarpsPerCta][kTileFloat2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_tw[i] = p.tw_pass[i];
    float2 ws[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    __syncthreads();

    using V2 = typename vec2<PCM>::type;
    const size_t nwarps = (size_t) gridDim.x * kWarpsPerCta;
    for (size_t f = (size_t) blockIdx.x * kWarpsPerCta + warp; f < p.nframes; f += nwarps) {
        const V2* src = reinterpret_cast<const V2*>(static_cast<const PCM*>(p.pcm) + f * 2048);
        float ur[32], ui[32], dr[32], di[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            V2 raw = src[m];
            float x0 = pcm_to_float(raw.x), x1 = pcm_to_float(raw.y);
            float2 cu = __ldg(p.chirp_up + m), cd = __ldg(p.chirp_down + m), w = __ldg(p.hann + m);
            ur[b] = __fmul_rn(__fmul_rn(x0, cu.x), w.x);
            ui[b] = __fmul_rn(__fmul_rn(x1, cu.y), w.y);
            dr[b] = __fmul_rn(__fmul_rn(x0, cd.x), w.x);
            di[b] = __fmul_rn(__fmul_rn(x1, cd.y), w.y);
        }
        float mu, md;
        uint32_t iu, id;
        fft1024_warp(ur, ui, s_tile[warp], s_tw, lane);
        peak_right<NB>(ur, ui, ws, lane, p.bandwidth2, mu, iu);
        warp_argmax(mu, iu);
        fft1024_warp(dr, di, s_tile[warp], s_tw, lane);
        peak_right<NB>(dr, di, ws, lane, p.bandwidth2, md, id);
        warp_argmax(md, id);
        if (lane == 0) {
            if (p.mag_up) p.mag_up[f] = mu;
            if (p.idx_up) p.idx_up[f] = iu;
            if (p.mag_down) p.mag_down[f] = md;
            if (p.idx_down) p.idx_down[f] = id;
            if (p.bit) p.bit[f] = md > mu ? 0 : 1;     // receiver/Src/main.c:523: down only if strictly greater
        }
    }
}

// dsp() for one hypothesis with a per-stream gather offset (receiver/Src/main.c:183-231).
// The receiver variant's "left" window reads the zero upper half (hazard H1, defined): its maximum
// is 0 at relative index 0, so the right window wins unless its own maximum is negative (never).
template <int NB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_dsp2048(demod_params p) {
    __shared__ float2 s_tw[32 * 32];
    __shared__ float2 s_tile[kWarpsPerCta][kTileFloat2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_tw[i] = p.tw_pass[i];
    float2 ws[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    __syncthreads();
    const float2* chirp = p.updown ? p.chirp_up : p.chirp_down;
    const size_t nwarps = (size_t) gridDim.x * kWarpsPerCta;
    for (size_t s = (size_t) blockIdx.x * kWarpsPerCta + warp; s < p.nframes; s += nwarps) {
        const float* src = static_cast<const float*>(p.pcm) + s * p.fifo_stride + p.sync_position[s];
        float vr[32], vi[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            float x0 = src[2 * m], x1 = src[2 * m + 1];          // offset may be odd: scalar loads
            float2 c = __ldg(chirp + m), w = __ldg(p.hann + m);
            vr[b] = __fmul_rn(__fmul_rn(x0, c.x), w.x);
            vi[b] = __fmul_rn(__fmul_rn(x1, c.y), w.y);
        }
        float mr;
        uint32_t ir;
        fft1024_warp(vr, vi, s_tile[warp], s_tw, lane);
        peak_right<NB>(vr, vi, ws, lane, p.bandwidth2, mr, ir);
        warp_argmax(mr, ir);
        if (lane == 0) {
            const float ml = 0.0f;
            const uint32_t il = p.idx_left_zero;
            float mm = mr;
            uint32_t im = ir;
            if (ml > mr) { mm = ml; im = il; }
            auto idx2freq = [&](uint32_t idx) -> int32_t {       // receiver/Src/main.c:154-160
                if (idx < 1024u) return (int32_t) ((uint32_t) p.fs_int * idx / 2048u);
                return (int32_t) ((uint32_t) p.fs_int * (2048u - idx) / 2048u) * -1;
            };
            const float mean = p.mag_mean[s];
            history_rec h;
            h.mag_max = mm; h.mag_max_left = ml; h.mag_max_right = mr;
            h.max_idx = im; h.max_idx_left = il; h.max_idx_right = ir;
            h.max_freq = idx2freq(im); h.max_freq_left = idx2freq(il); h.max_freq_right = idx2freq(ir);
            h.mag_mean = mean;
            h.snr = __fdiv_rn(__fsub_rn(mm, mean), mean);         // receiver/Src/main.c:229
            h.rank = (uint32_t) '-';
            p.hist[s] = h;
        }
    }
}

static int grid_for(size_t nwork, int num_sms) {
    size_t ctas = (nwork + kWarpsPerCta - 1) / kWarpsPerCta;
    size_t cap = (size_t) num_sms * 2 * 4;             // persistent-ish: a few waves of resident CTAs
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int) ctas;
}

template <int NB>
static cudaError_t launch_demod_nb(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st) {
    int grid = grid_for(p.nframes, num_sms);
    if (pcm_format == 1u) k_demod2048<int32_t, NB><<<grid, kWarpsPerCta * 32, 0, st>>>(p);
    else k_demod2048<float, NB><<<grid, kWarpsPerCta * 32, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_demod2048(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st) {
    const uint32_t nb = (p.bandwidth2 + 31) / 32;
    if (nb <= 3) return launch_demod_nb<3>(p, pcm_format, num_sms, st);
    if (nb <= 5) return launch_demod_nb<5>(p, pcm_format, num_sms, st);
    if (nb <= 8) return launch_demod_nb<8>(p, pcm_format, num_sms, st);
    return launch_demod_nb<16>(p, pcm_format, num_sms, st);
}

cudaError_t launch_dsp2048(const demod_params& p, int num_sms, cudaStream_t st) {
    const uint32_t nb = (p.bandwidth2 + 31) / 32;
    int grid = grid_for(p.nframes, num_sms);
    if (nb <= 5) k_dsp2048<5><<<grid, kWarpsPerCta * 32, 0, st>>>(p);
    else k_dsp2048<16><<<grid, kWarpsPerCta * 32, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace usc
