// k_iq.cu — K5: the I/Q baseband path (experiments/iq_modulation/Src/iq_modem.c:34-75, Src/main.c:117-134, with the
// semantics of simulation/IQ_modulation.ipynb cells 16-31 and BASELINE config 3's decimation by 2).
//   k_iq_fused        the whole path in one kernel (the product path for ntaps <= 32, window <= 32 bins): carrier mix,
//                     FIR on I and Q with the state carried from the previous frame, /2, R = I + jQ, de-chirp both ways,
//                     Hann, 1024-point complex FFT, windowed arg-max around DC, decision
//   k_iq_frontend[_rb], k_iq_backend, k_iq_pick
//                     the same chain in pieces (front end -> R in HBM -> fused back end, or -> batched operators):
//                     fallback for longer filters / wider windows and the form USC_IQ_UNFUSED=1 selects for A/B runs
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

template <typename PCM>
__global__ void __launch_bounds__(256) k_iq_frontend(const PCM* __restrict__ pcm, uint32_t nstreams, uint32_t nframes,
                                                     size_t stream_stride, uint32_t n, const float* __restrict__ car_cos,
                                                     const float* __restrict__ car_sin, const float* __restrict__ taps,
                                                     uint32_t ntaps, float2* __restrict__ out) {
    extern __shared__ float s_iq[];
    const uint32_t H = ntaps - 1, L = n + H;
    float* sI = s_iq;
    float* sQ = s_iq + L;
    float* sT = s_iq + 2 * L;
    const size_t total = (size_t) nstreams * nframes;
    for (size_t w = blockIdx.x; w < total; w += gridDim.x) {
        const uint32_t s = (uint32_t) (w / nframes), t = (uint32_t) (w - (size_t) s * nframes);
        const PCM* cur = pcm + (size_t) s * stream_stride + (size_t) t * n;
        for (uint32_t i = threadIdx.x; i < ntaps; i += blockDim.x) sT[i] = taps[i];
        for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) {
            float vi = 0.0f, vq = 0.0f;
            if (j >= H) {
                const uint32_t k = j - H;
                const float x = pcm_cast(cur[k]);
                vi = __fmul_rn(x, car_cos[k]);
                vq = __fmul_rn(x, car_sin[k]);
            } else if (t > 0) {                              // history: tail of the previous frame
                const uint32_t k = n - H + j;
                const float x = pcm_cast(cur[(ptrdiff_t) k - (ptrdiff_t) n]);
                vi = __fmul_rn(x, car_cos[k]);
                vq = __fmul_rn(x, car_sin[k]);
            }
            sI[j] = vi;
            sQ[j] = vq;
        }
        __syncthreads();
        float2* dst = out + w * (n / 2);
        for (uint32_t m = threadIdx.x; m < n / 2; m += blockDim.x) {
            float ai = 0.0f, aq = 0.0f;
            for (uint32_t i = 0; i < ntaps; ++i) {
                ai = __fmaf_rn(sI[2 * m + i], sT[i], ai);
                aq = __fmaf_rn(sQ[2 * m + i], sT[i], aq);
            }
            dst[m] = make_float2(ai, aq);
        }
        __syncthreads();
    }
}

// Register-blocked variant for ntaps <= 32: one CTA per frame, each thread owns four consecutive outputs
// of both rails.  Mixed samples are staged as (even, odd) pairs; output j at tap pair q - j reads pair
// m0 + q, so one pair load feeds up to four outputs.  The accumulation order per output is the
// reference's (taps ascending, one FMA each).  Pair index p lives at p + (p >> 2) to spread the
// 4-pair lane stride over the banks.
constexpr int kFirTmax = 32, kFirOut = 4, kFirPairs = kFirTmax / 2 + kFirOut - 1;

template <typename PCM>
__global__ void __launch_bounds__(256) k_iq_frontend_rb(const PCM* __restrict__ pcm, uint32_t nstreams, uint32_t nframes,
                                                        size_t stream_stride, const float* __restrict__ car_cos,
                                                        const float* __restrict__ car_sin, const float* __restrict__ taps,
                                                        uint32_t ntaps, float2* __restrict__ out) {
    constexpr uint32_t n = 2048, P = n / 2 + kFirPairs + 4, PP = P + (P >> 2) + 1;      // pairs per rail (padded)
    __shared__ float2 sI[PP], sQ[PP];
    const uint32_t H = ntaps - 1;
    float t[kFirTmax];
#pragma unroll
    for (int i = 0; i < kFirTmax; ++i) t[i] = (uint32_t) i < ntaps ? taps[i] : 0.0f;
    const size_t total = (size_t) nstreams * nframes;
    for (size_t w = blockIdx.x; w < total; w += gridDim.x) {
        const uint32_t s = (uint32_t) (w / nframes), fr = (uint32_t) (w - (size_t) s * nframes);
        const PCM* cur = pcm + (size_t) s * stream_stride + (size_t) fr * n;
        for (uint32_t j = threadIdx.x; j < 2 * P; j += blockDim.x) {
            float vi = 0.0f, vq = 0.0f;
            if (j >= H && j < n + H) {
                const uint32_t k = j - H;
                const float x = pcm_cast(cur[k]);
                vi = __fmul_rn(x, car_cos[k]);
                vq = __fmul_rn(x, car_sin[k]);
            } else if (j < H && fr > 0) {                    // history: tail of the previous frame
                const uint32_t k = n - H + j;
                const float x = pcm_cast(cur[(ptrdiff_t) k - (ptrdiff_t) n]);
                vi = __fmul_rn(x, car_cos[k]);
                vq = __fmul_rn(x, car_sin[k]);
            }
            const uint32_t pp = j >> 1, slot = pp + (pp >> 2);
            reinterpret_cast<float*>(sI)[2 * slot + (j & 1)] = vi;
            reinterpret_cast<float*>(sQ)[2 * slot + (j & 1)] = vq;
        }
        __syncthreads();
        const uint32_t m0 = threadIdx.x * kFirOut;
        float ai[kFirOut], aq[kFirOut];
#pragma unroll
        for (int j = 0; j < kFirOut; ++j) ai[j] = aq[j] = 0.0f;
#pragma unroll
        for (int q = 0; q < kFirPairs; ++q) {
            if ((uint32_t) (2 * (q - (kFirOut - 1))) >= ntaps && q >= kFirOut - 1) break;     // past the last tap for every output
            const uint32_t pp = m0 + q, slot = pp + (pp >> 2);
            const float2 vi = sI[slot], vq = sQ[slot];
#pragma unroll
            for (int j = 0; j < kFirOut; ++j) {
                const int i0 = 2 * (q - j);
                if (i0 < 0 || i0 >= kFirTmax) continue;
                if ((uint32_t) i0 < ntaps) {
                    ai[j] = __fmaf_rn(vi.x, t[i0], ai[j]);
                    aq[j] = __fmaf_rn(vq.x, t[i0], aq[j]);
                }
                if ((uint32_t) (i0 + 1) < ntaps) {
                    ai[j] = __fmaf_rn(vi.y, t[i0 + 1], ai[j]);
                    aq[j] = __fmaf_rn(vq.y, t[i0 + 1], aq[j]);
                }
            }
        }
        float4* dst = reinterpret_cast<float4*>(out + w * (n / 2) + m0);
        dst[0] = make_float4(ai[0], aq[0], ai[1], aq[1]);
        dst[1] = make_float4(ai[2], aq[2], ai[3], aq[3]);
        __syncthreads();
    }
}

// left/right choice (strict '>' keeps the right window on ties) and absolute bin of the winner
__global__ void k_iq_pick(const float* __restrict__ mr, const uint32_t* __restrict__ ir, const float* __restrict__ ml,
                          const uint32_t* __restrict__ il, uint32_t left0, float* __restrict__ mag,
                          uint32_t* __restrict__ idx, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
        const bool left = ml[i] > mr[i];
        if (mag) mag[i] = left ? ml[i] : mr[i];
        if (idx) idx[i] = left ? left0 + il[i] : ir[i];
    }
}

cudaError_t launch_iq_frontend(const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes,
                               size_t stream_stride, uint32_t n, const float* car_cos, const float* car_sin,
                               const float* taps, uint32_t ntaps, float* out, cudaStream_t st) {
    const size_t total = (size_t) nstreams * nframes;
    const int grid = (int) (total < 148u * 16u ? total : 148u * 16u);
    const size_t smem = sizeof(float) * (2 * ((size_t) n + ntaps - 1) + ntaps);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    if (n == 2048 && ntaps <= (uint32_t) kFirTmax && ntaps >= 1) {
        const int g2 = (int) (total < 148u * 8u ? total : 148u * 8u);
        if (pcm_format == 1u)
            k_iq_frontend_rb<int32_t><<<g2, 256, 0, st>>>((const int32_t*) pcm, nstreams, nframes, stream_stride, car_cos,
                                                           car_sin, taps, ntaps, (float2*) out);
        else
            k_iq_frontend_rb<float><<<g2, 256, 0, st>>>((const float*) pcm, nstreams, nframes, stream_stride, car_cos, car_sin,
                                                         taps, ntaps, (float2*) out);
        return cudaGetLastError();
    }
    if (pcm_format == 1u)
        k_iq_frontend<int32_t><<<grid, 256, smem, st>>>((const int32_t*) pcm, nstreams, nframes, stream_stride, n, car_cos,
                                                         car_sin, taps, ntaps, (float2*) out);
    else
        k_iq_frontend<float><<<grid, 256, smem, st>>>((const float*) pcm, nstreams, nframes, stream_stride, n, car_cos,
                                                       car_sin, taps, ntaps, (float2*) out);
    return cudaGetLastError();
}

cudaError_t launch_iq_pick(const float* mr, const uint32_t* ir, const float* ml, const uint32_t* il, uint32_t left0,
                           float* mag, uint32_t* idx, size_t count, cudaStream_t st) {
    size_t b = (count + 255) / 256;
    if (b > 148u * 8u) b = 148u * 8u;
    k_iq_pick<<<(int) (b ? b : 1), 256, 0, st>>>(mr, ir, ml, il, left0, mag, idx, count);
    return cudaGetLastError();
}

// ---- fused back end of the I/Q path ------------------------------------------------------------------
// One warp per frame; the halves of the packed core carry the two hypotheses: .x = R x conj(chirp)
// (up), .y = R x chirp (down).  Hann, the 1024-point complex FFT ([32,32] plan), magnitudes and the two
// windowed arg-max searches around DC follow.  With window_bins <= 32 only outputs k = d0 (element 0)
// and k = 992 + d0 (element 31) of the last pass are needed, so 30 of its 32 outputs are pruned.

constexpr int kIqWarps = 8;
constexpr int kIqSmem = 8192 + 8192 + 4096 + kIqWarps * 8192;     // twiddles | chirp | Hann | per-warp tile

__global__ void __launch_bounds__(kIqWarps * 32, 1) k_iq_backend(const float2* __restrict__ R, size_t nframes,
                                                                 const float2* __restrict__ chirp,
                                                                 const float* __restrict__ hann,
                                                                 const float2* __restrict__ tw_pass, uint32_t W,
                                                                 float* mag_up, uint32_t* idx_up, float* mag_down,
                                                                 uint32_t* idx_down, uint8_t* bit) {
    extern __shared__ __align__(16) unsigned char s_iq2[];
    float2* s_tw = reinterpret_cast<float2*>(s_iq2);
    float2* s_c = reinterpret_cast<float2*>(s_iq2 + 8192);
    float* s_w = reinterpret_cast<float*>(s_iq2 + 16384);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* tile = reinterpret_cast<float2*>(s_iq2 + 20480) + warp * 1024;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = tw_pass[i];
        s_c[i] = chirp[i];
        s_w[i] = hann[i];
    }
    __syncthreads();
    const size_t nwarps = (size_t) gridDim.x * kIqWarps;
    for (size_t f = (size_t) blockIdx.x * kIqWarps + warp; f < nframes; f += nwarps) {
        const float2* src = R + f * 1024;
        float2 re[32], im[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            const float2 r = src[m], c = s_c[m];
            const float w = s_w[m];
            float ur, ui, dr, di;
            cmul(r.x, r.y, c.x, -c.y, ur, ui);           // arm_cmplx_mult_cmplx_f32(R, conj chirp)
            cmul(r.x, r.y, c.x, c.y, dr, di);            // arm_cmplx_mult_cmplx_f32(R, chirp)
            re[b] = __fmul2_rn(make_float2(ur, dr), bc2(w));              // arm_cmplx_mult_real_f32, packed (first stage: FMAs by 1.0)
            im[b] = __fmul2_rn(make_float2(ui, di), bc2(w));
        }
        fft1024_pair<true>(re, im, tile, s_tw, lane);
        // candidates of this lane: right window k = lane (element 0), left window k = 992 + lane (element 31)
        const float2 pr = __ffma2_rn(re[0], re[0], __fmul2_rn(im[0], im[0]));
        const float2 pl = __ffma2_rn(re[31], re[31], __fmul2_rn(im[31], im[31]));
        const bool in_r = (uint32_t) lane < W, in_l = 992u + lane >= 1024u - W;
        float out_m[2];
        uint32_t out_i[2];
#pragma unroll
        for (int hyp = 0; hyp < 2; ++hyp) {
            float mr = in_r ? __fsqrt_rn(hyp ? pr.y : pr.x) : -INFINITY, ml = in_l ? __fsqrt_rn(hyp ? pl.y : pl.x) : -INFINITY;
            uint32_t ir = in_r ? (uint32_t) lane : 0xffffffffu, il = in_l ? 992u + lane : 0xffffffffu;
            warp_argmax(mr, ir);                           // arm_max_f32 over [0, W)
            warp_argmax(ml, il);                           // arm_max_f32 over [1024 - W, 1024)
            const bool left = ml > mr;                     // strict: right wins ties
            out_m[hyp] = left ? ml : mr;
            out_i[hyp] = left ? il : ir;
        }
        if (lane == 0) {
            if (mag_up) mag_up[f] = out_m[0];
            if (idx_up) idx_up[f] = out_i[0];
            if (mag_down) mag_down[f] = out_m[1];
            if (idx_down) idx_down[f] = out_i[1];
            if (bit) bit[f] = out_m[1] > out_m[0] ? 0 : 1;
        }
    }
}

cudaError_t launch_iq_backend(const float* R, size_t nframes, const float* chirp, const float* hann, const float2* tw_pass,
                              uint32_t window, float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                              uint8_t* bit, int num_sms, cudaStream_t st) {
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_iq_backend, cudaFuncAttributeMaxDynamicSharedMemorySize, kIqSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    size_t ctas = (nframes + kIqWarps - 1) / kIqWarps;
    if (ctas > (size_t) num_sms * 2) ctas = (size_t) num_sms * 2;
    k_iq_backend<<<(int) ctas, kIqWarps * 32, kIqSmem, st>>>((const float2*) R, nframes, (const float2*) chirp, hann, tw_pass,
                                                             window, mag_up, idx_up, mag_down, idx_down, bit);
    return cudaGetLastError();
}

// ---- whole I/Q path in one kernel ---------------------------------------------------------------------
// One warp per frame, persistent CTAs of 8 warps.  Stage: PCM (+ ntaps-1 samples of history) is cast,
// mixed with the carrier and parked as (I, Q) pairs.  FIR: each lane owns four consecutive decimated
// outputs per pass (eight passes); both rails ride one FFMA2 with the tap broadcast, taps ascending as in
// arm_fir_f32.  The baseband frame R goes through the warp's region once (blocked -> strided), then the
// back end above runs unchanged.  R never touches HBM.

#ifndef USC_IQF_PAD
#define USC_IQF_PAD 1                                   // exchange tile with padded rows (0: XOR-swizzled)
#endif
constexpr int kIqfWarps = 8;
#ifndef USC_IQF_ROWS
#define USC_IQF_ROWS 64
#endif
constexpr int kIqfRows = USC_IQF_ROWS;                                       // 32-sample rows loaded per lane before any is used
// Tables in tensor memory, one row per lane (usc_tmem.cuh): carrier (cos, sin) of sample 32 r + lane at 2 r, r < 64 |
// (chirp.re, chirp.im, Hann, -) of m = lane + 32 b at 128 + 4 b | inter-pass twiddle W_1024^(lane d) at 256 + 2 d.
// Shared memory keeps only the carrier values of the last 32 samples (the history rows of the next frame read them
// at a lane offset) and the TMEM slot.
constexpr int kIqfTcs = 0, kIqfTcw = 128, kIqfTtw = 256, kIqfTcols = 512;
constexpr int kIqfTables = 256 + 16;                                          // carrier tail | TMEM slot
// NT = 0: the number of taps is a run-time value (<= 32): four outputs per lane and pass, tap pairs predicated.
// NT > 0: the filter length is a compile-time constant (27 = the reference's, iq_modem.c:16-18): EIGHT outputs per
// lane and pass — 21 instead of 2 x 19 shared-memory loads per eight outputs, exactly NT FMAs per output, no branches.
// *dst = (x * c.x, x * c.y) with the packed product handed to the store as one 64-bit register: written through the
// float2 intrinsics the compiler unpacks and repacks the pair (two MOVs per sample in the staging loop)
__device__ __forceinline__ void st_mul2(float2* dst, float x, float2 c) {
    asm volatile("{\n\t.reg .b64 a, b, r;\n\tmov.b64 a, {%1, %1};\n\tmov.b64 b, {%2, %3};\n\tmul.rn.f32x2 r, a, b;\n\t"
                 "st.shared.b64 [%0], r;\n\t}" ::"r"(smem_u32(dst)), "f"(x), "f"(c.x), "f"(c.y) : "memory");
}

template <int NT> struct iqf_geom {
    static_assert(NT == 0 || (NT & 1) == 1, "the specialised form pairs samples (2i, 2i+1) of the history-extended frame: odd filter lengths only");
    static constexpr int OUT = NT ? 8 : kFirOut;                              // consecutive decimated outputs per lane and pass
    static constexpr int SK = NT ? 3 : 2;                                     // pair p lives in float4 slot p + (p >> SK): conflict-free windows
    static constexpr int NP = NT ? (NT + 1) / 2 : kFirTmax / 2;               // tap pairs
    static constexpr int PS = NT ? (32 - (NT - 1)) / 2 : 0;                   // pairs of padding in front: PCM sample 0 lands on staged sample 32, so
                                                                              // every 32-sample row a warp stores is one aligned, gap-free 256-byte run
    static constexpr uint32_t pairs = 1024 + NP + OUT - 1 + 4 + PS;           // sample pairs per frame incl. history
    static constexpr uint32_t slots = pairs + (pairs >> SK) + 1;
    static constexpr int region = (int) slots * 16;
    static constexpr int smem = kIqfTables + kIqfWarps * region;
    // R (1024 decimated outputs, float2) is parked at the head of the region.  With OUT = 8 a lane stores 64 contiguous
    // bytes per pass: float4 i goes to i ^ ((i >> 3) & 7), which spreads the lanes over all banks and leaves the strided
    // reads of the back end (consecutive lanes, consecutive outputs) conflict-free as well.
    __host__ __device__ static constexpr uint32_t r_f4(uint32_t i) { return NT ? i ^ ((i >> 3) & 7u) : i; }
    __host__ __device__ static constexpr uint32_t r_index(uint32_t m) { return 2u * r_f4(m >> 1) + (m & 1u); }
};

template <typename PCM, int NT>
__global__ void __launch_bounds__(kIqfWarps * 32, 1) k_iq_fused(const PCM* __restrict__ pcm, uint32_t nstreams, uint32_t nframes,
                                                                size_t stream_stride, const float* __restrict__ car_cos,
                                                                const float* __restrict__ car_sin,
                                                                const float* __restrict__ taps, uint32_t ntaps,
                                                                const float2* __restrict__ chirp, const float* __restrict__ hann,
                                                                const float2* __restrict__ tw_pass, uint32_t W, float* mag_up,
                                                                uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                                                                uint8_t* bit) {
    constexpr uint32_t n = 2048;
    using G = iqf_geom<NT>;
    constexpr int OUT = G::OUT, SK = G::SK, NP = G::NP, PS = G::PS;
    constexpr int kIqfRegion = G::region;
    constexpr uint32_t ROWSTEP = 2u * (16u + (16u >> SK));                   // float2 per 32-sample row of the staged sequence
    extern __shared__ __align__(16) unsigned char s_f[];
    float2* s_cs_tail = reinterpret_cast<float2*>(s_f);      // carrier of samples 2016 .. 2047
    uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_f + 256);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* region = s_f + kIqfTables + warp * kIqfRegion;
    float2* smp = reinterpret_cast<float2*>(region);         // sample j at float2 index 2 * slot(j >> 1) + (j & 1)
    if (threadIdx.x < 32) s_cs_tail[threadIdx.x] = make_float2(car_cos[2016 + threadIdx.x], car_sin[2016 + threadIdx.x]);
    if (warp == 0) tmem_alloc<kIqfTcols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tq = tmem_quadrant(*s_tslot, warp);
    if (warp < 4) {                                           // warp q fills lane quadrant q; warps q and q + 4 read it
#pragma unroll 1
        for (int r0 = 0; r0 < 64; r0 += 4) {
            float2 c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = make_float2(car_cos[(r0 + j) * 32 + lane], car_sin[(r0 + j) * 32 + lane]);
            sttm_f2x4(tq + kIqfTcs + 2 * r0, c[0], c[1], c[2], c[3]);
        }
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += 4) {
            float2 c[4], w[4], z[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j] = chirp[lane + 32 * (b0 + j)];
                w[j] = make_float2(hann[lane + 32 * (b0 + j)], 0.0f);
                z[j] = tw_pass[(b0 + j) * 32 + lane];
            }
            sttm_f2x4(tq + kIqfTcw + 4 * b0, c[0], w[0], c[1], w[1]);
            sttm_f2x4(tq + kIqfTcw + 4 * b0 + 8, c[2], w[2], c[3], w[3]);
            sttm_f2x4(tq + kIqfTtw + 2 * b0, z[0], z[1], z[2], z[3]);
        }
        sttm_wait();
    }
    const float one = tw_pass[lane].x;                       // W^0 = 1.0f read from a table: opaque to the compiler
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t H = ntaps - 1;
    float t[kFirTmax];
#pragma unroll
    for (int i = 0; i < kFirTmax; ++i) t[i] = (uint32_t) i < ntaps ? taps[i] : 0.0f;
    const size_t total = (size_t) nstreams * nframes, nwarps = (size_t) gridDim.x * kIqfWarps;
    for (size_t f = (size_t) blockIdx.x * kIqfWarps + warp; f < total; f += nwarps) {
        const uint32_t s = (uint32_t) (f / nframes), fr = (uint32_t) (f - (size_t) s * nframes);
        const PCM* cur = pcm + (size_t) s * stream_stride + (size_t) fr * n;
        if (f + nwarps < total) {                            // pull the next frame of this warp towards L2
            const size_t g = f + nwarps;
            const uint32_t s2 = (uint32_t) (g / nframes), f2 = (uint32_t) (g - (size_t) s2 * nframes);
            const char* nxt = reinterpret_cast<const char*>(pcm + (size_t) s2 * stream_stride + (size_t) f2 * n);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + lane * 128));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 4096 + lane * 128));
        }
        // ---- stage: cast, mix, park as (I, Q) ----
        // Sample j = k + H of the staged sequence is PCM sample k: no bounds to test in the main part, and
        // sixteen independent loads are in flight per lane.
        {
            const uint32_t jl = lane + H + 2u * PS, pl = jl >> 1;
#pragma unroll 1
            for (uint32_t r0 = 0; r0 < 64; r0 += kIqfRows) {
                float x[kIqfRows];
#pragma unroll
                for (int u = 0; u < kIqfRows; ++u) x[u] = pcm_cast(cur[(r0 + u) * 32 + lane]);
                // pair pl + 16 r sits in slot pl + (pl >> SK) + (16 + (16 >> SK)) r: a constant number of float2 per row
                float2* dst = smp + 2 * (pl + (pl >> SK)) + (jl & 1u) + ROWSTEP * r0;
#pragma unroll
                for (int u0 = 0; u0 < kIqfRows; u0 += 8) {     // carrier values of eight rows per TMEM round trip
                    uint32_t c[16];
                    ldtm16(tq + kIqfTcs + 2 * (r0 + u0), c);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        st_mul2(dst + ROWSTEP * (u0 + u), x[u0 + u], make_float2(__uint_as_float(c[2 * u]), __uint_as_float(c[2 * u + 1])));
                }
            }
            if ((uint32_t) lane < H) {                         // history: tail of the previous frame (zeros before frame 0)
                const uint32_t k = n - H + lane, pp = ((uint32_t) lane >> 1) + PS;
                const float xh = fr > 0 ? pcm_cast(cur[(ptrdiff_t) lane - (ptrdiff_t) H]) : 0.0f;
                smp[2 * (pp + (pp >> SK)) + (lane & 1)] = __fmul2_rn(bc2(xh), s_cs_tail[k - 2016u]);
            }
        }
        __syncwarp();
        // ---- FIR + decimation by two; R parked blocked at the head of the region ----
        // Output m0 + j at tap pair ip reads sample pair m0 + j + ip: a window of four pairs slides one pair
        // per tap pair.  Per output the taps run in ascending order, one FMA each (arm_fir_f32).
#pragma unroll 1
        for (uint32_t pass = 0; pass < 32u / OUT; ++pass) {
            const uint32_t m0 = (pass * 32 + lane) * OUT;
            float2 acc[OUT];
#pragma unroll
            for (int j = 0; j < OUT; ++j) acc[j] = make_float2(0.0f, 0.0f);
            float4 w[NP + OUT];
#pragma unroll
            for (int q = 0; q < OUT - 1; ++q) {
                const uint32_t pp = m0 + q + PS;
                w[q] = reinterpret_cast<const float4*>(region)[pp + (pp >> SK)];
            }
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) {
                const uint32_t pp = m0 + ip + OUT - 1 + PS;
                w[ip + OUT - 1] = reinterpret_cast<const float4*>(region)[pp + (pp >> SK)];
                if (NT) {                                    // compile-time filter length: exactly NT FMAs per output
#pragma unroll
                    for (int j = 0; j < OUT; ++j) acc[j] = __ffma2_rn(make_float2(w[ip + j].x, w[ip + j].y), bc2(t[2 * ip]), acc[j]);
                    if (2 * ip + 1 < NT) {
#pragma unroll
                        for (int j = 0; j < OUT; ++j) acc[j] = __ffma2_rn(make_float2(w[ip + j].z, w[ip + j].w), bc2(t[2 * ip + 1]), acc[j]);
                    }
                } else {
                    // no early exit: predicated in place, so the accumulators keep their registers
                    const bool first = (uint32_t) (2 * ip) < ntaps, second = (uint32_t) (2 * ip + 1) < ntaps;
#pragma unroll
                    for (int j = 0; j < OUT; ++j) {
                        const float2 a0 = __ffma2_rn(make_float2(w[ip + j].x, w[ip + j].y), bc2(t[2 * ip]), acc[j]);
                        acc[j] = first ? a0 : acc[j];
                        const float2 a1 = __ffma2_rn(make_float2(w[ip + j].z, w[ip + j].w), bc2(t[2 * ip + 1]), acc[j]);
                        acc[j] = second ? a1 : acc[j];
                    }
                }
            }
            __syncwarp();                                    // this pass's samples overlap where R of pass 0 lands
            float4* dst = reinterpret_cast<float4*>(region);
#pragma unroll
            for (int j = 0; j < OUT; j += 2)
                dst[G::r_f4((m0 >> 1) + (j >> 1))] = make_float4(acc[j].x, acc[j].y, acc[j + 1].x, acc[j + 1].y);
        }
        __syncwarp();
        // ---- back end: de-chirp both ways, Hann, FFT, windowed peaks ----
        float2 re[32], im[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {                        // chirp and Hann values of four rows per TMEM round trip
            uint32_t tv[16];
            ldtm16(tq + kIqfTcw + 16 * g, tv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = 4 * g + j, m = lane + 32 * b;
                const float2 r = reinterpret_cast<const float2*>(region)[G::r_index((uint32_t) m)];
                const float2 c = make_float2(__uint_as_float(tv[4 * j]), __uint_as_float(tv[4 * j + 1]));
                const float w = __uint_as_float(tv[4 * j + 2]);
                // cmul(R, conj c) and cmul(R, c) share their two rounded products; the FMAs ride one FFMA2 each
                const float t0 = __fmul_rn(r.y, c.y), t1 = __fmul_rn(r.y, c.x);
                const float2 pr2 = __ffma2_rn(bc2(r.x), bc2(c.x), make_float2(t0, -t0));      // (up.re, down.re)
                const float2 pi2 = __ffma2_rn(bc2(r.x), make_float2(-c.y, c.y), bc2(t1));     // (up.im, down.im)
                re[b] = __fmul2_rn(pr2, bc2(w));                                              // packed (first stage: FMAs by 1.0)
                im[b] = __fmul2_rn(pi2, bc2(w));
            }
        }
        __syncwarp();
        fft1024_pair_tm<true, USC_IQF_PAD != 0>(re, im, reinterpret_cast<float2*>(region), tq + kIqfTtw, one, lane);
        const float2 pr = __ffma2_rn(re[0], re[0], __fmul2_rn(im[0], im[0]));
        const float2 pl = __ffma2_rn(re[31], re[31], __fmul2_rn(im[31], im[31]));
        const bool in_r = (uint32_t) lane < W, in_l = 992u + lane >= 1024u - W;
        float out_m[2];
        uint32_t out_i[2];
#pragma unroll
        for (int hyp = 0; hyp < 2; ++hyp) {
            // arm_max_f32 over [0, W) and [1024 - W, 1024): magnitudes are >= 0, so their bit patterns order like the
            // values and one redux.sync finds the maximum, a second one the first lane that holds it
            const uint32_t br = in_r ? __float_as_uint(__fsqrt_rn(hyp ? pr.y : pr.x)) : 0u;
            const uint32_t bl = in_l ? __float_as_uint(__fsqrt_rn(hyp ? pl.y : pl.x)) : 0u;
            const uint32_t mrb = __reduce_max_sync(0xffffffffu, br), mlb = __reduce_max_sync(0xffffffffu, bl);
            const uint32_t ir = __reduce_min_sync(0xffffffffu, in_r && br == mrb ? (uint32_t) lane : 0xffffffffu);
            const uint32_t il = __reduce_min_sync(0xffffffffu, in_l && bl == mlb ? 992u + lane : 0xffffffffu);
            const float mr = __uint_as_float(mrb), ml = __uint_as_float(mlb);
            const bool left = ml > mr;                       // strict: right wins ties
            out_m[hyp] = left ? ml : mr;
            out_i[hyp] = left ? il : ir;
        }
        if (lane == 0) {
            if (mag_up) mag_up[f] = out_m[0];
            if (idx_up) idx_up[f] = out_i[0];
            if (mag_down) mag_down[f] = out_m[1];
            if (idx_down) idx_down[f] = out_i[1];
            if (bit) bit[f] = out_m[1] > out_m[0] ? 0 : 1;
        }
        __syncwarp();
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kIqfTcols>(*s_tslot);
}

template <typename PCM, int NT>
static cudaError_t launch_iq_fused_t(const void* pcm, uint32_t nstreams, uint32_t nframes, size_t stream_stride, const float* car_cos,
                                     const float* car_sin, const float* taps, uint32_t ntaps, const float* chirp, const float* hann,
                                     const float2* tw_pass, uint32_t window, float* mag_up, uint32_t* idx_up, float* mag_down,
                                     uint32_t* idx_down, uint8_t* bit, int num_sms, cudaStream_t st) {
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    constexpr int smem = iqf_geom<NT>::smem;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_iq_fused<PCM, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const size_t total = (size_t) nstreams * nframes;
    size_t ctas = (total + kIqfWarps - 1) / kIqfWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    k_iq_fused<PCM, NT><<<(int) ctas, kIqfWarps * 32, smem, st>>>((const PCM*) pcm, nstreams, nframes, stream_stride, car_cos, car_sin,
        taps, ntaps, (const float2*) chirp, hann, tw_pass, window, mag_up, idx_up, mag_down, idx_down, bit);
    return cudaGetLastError();
}

cudaError_t launch_iq_fused(const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                            const float* car_cos, const float* car_sin, const float* taps, uint32_t ntaps, const float* chirp,
                            const float* hann, const float2* tw_pass, uint32_t window, float* mag_up, uint32_t* idx_up,
                            float* mag_down, uint32_t* idx_down, uint8_t* bit, int num_sms, cudaStream_t st) {
#define USC_IQF_ARGS pcm, nstreams, nframes, stream_stride, car_cos, car_sin, taps, ntaps, chirp, hann, tw_pass, window, mag_up, idx_up, \
                     mag_down, idx_down, bit, num_sms, st
    if (ntaps == 27u)                                        // the reference's filter (iq_modem.c:16-18): specialised form
        return pcm_format == 1u ? launch_iq_fused_t<int32_t, 27>(USC_IQF_ARGS) : launch_iq_fused_t<float, 27>(USC_IQF_ARGS);
    return pcm_format == 1u ? launch_iq_fused_t<int32_t, 0>(USC_IQF_ARGS) : launch_iq_fused_t<float, 0>(USC_IQF_ARGS);
#undef USC_IQF_ARGS
}

}  // namespace usc
