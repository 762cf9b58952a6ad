// k_long.cu — K6: the receiver chain fused for long frames, N = 2048*R0 with R0 = 2, 4 or 8 (4096 points — the
// longest CMSIS supports — and the 8192- and 16384-point frames of BASELINE config 5).  Same reference path as
// K1 (cast, de-chirp, Hann, RFFT, magnitude, arg-max over [0, bandwidth2), both hypotheses; receiver/Src/main.c:163-215), one HBM pass.
//
// The N/2-point complex FFT has the canonical plan [R0, 32, 32]: one radix-R0 level over a
// (m = a + 1024 b), then R0 independent 1024-point transforms whose outputs interleave
// (k = R0 c + d).  A group of R0 warps owns a frame (one group per CTA; two at 4096 points):
//   stage     the frame's PCM arrives in shared memory by TMA bulk copies issued one frame ahead
//   level 0   all threads: z[a + 1024 b] from the stage; chirp, Hann and level-0 twiddle values of the thread's a from
//             its own row of TENSOR MEMORY (one tcgen05.ld per round: the frame-sized tables, 60-240 KB, live there);
//             de-chirp x Hann for both hypotheses (f32x2 halves), radix-R0 butterfly, x W^(a d), park
//             sub-sequence d in shared memory
//   core      warp d runs the packed 32x32 register core of K1 on sub-sequence d; its parked region
//             doubles as the exchange tile once the data is in registers; only c < ceil(bw2/R0) and the
//             partner range near 1024 are produced (pruned last pass)
//   split     the needed sub-spectra meet in shared memory; threads take bins k < bw2, fetch
//             Z[k] = Y_{k mod R0}[k / R0] and Z[n-k], apply the real split, squared magnitudes; every warp finds the
//             exact (largest root, first index) of its bins with one square root (usc_warpfft.cuh), one thread combines
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"
#include "usc_tmem.cuh"

namespace usc {

#ifndef USC_LONG_UNROLL
#define USC_LONG_UNROLL 4
#endif
constexpr int kLongUnroll = USC_LONG_UNROLL;            // level-0 rounds whose loads are issued together
#ifndef USC_LONG_STAGE
#define USC_LONG_STAGE 1                                // PCM of the next frame staged in shared memory by TMA (0: gathered from global memory)
#endif
#ifndef USC_LONG_TMEM
#define USC_LONG_TMEM 1                                 // frame-sized tables (chirp, Hann, level-0 twiddles) in tensor memory
#endif
// Tensor-memory layout of the frame-sized tables.  Thread tid handles a = tid + T i in level-0 round i and
// always needs the same table entries: the (up, down) chirp and Hann values of m = a + 1024 b, b < R0, and the level-0
// twiddles W^(a d), d = 1..R0-1 — 8 R0 - 2 words, padded to 8 R0 columns per round, 32 / R0 rounds: 256 columns of the
// thread's own TMEM lane (warps 4..7 of the 8-warp form take columns 256..511).  The tables total 30 KB per 8 KB of
// PCM; read through L2/L1 for every frame they were the kernel's largest stream.
template <int R0> struct long_tmem {
    static constexpr int per_round = 8 * R0, rounds = 32 / R0, c_ud = 0, c_hann = 4 * R0, c_tw = 6 * R0,
                         cols = R0 == 8 ? 512 : 256;
};
template <int N> struct ldtm_n;
template <> struct ldtm_n<16> { static __device__ __forceinline__ void ld(uint32_t ta, uint32_t (&t)[16]) { ldtm16(ta, t); } };
template <> struct ldtm_n<32> { static __device__ __forceinline__ void ld(uint32_t ta, uint32_t (&t)[32]) { ldtm32(ta, t); } };
template <> struct ldtm_n<64> { static __device__ __forceinline__ void ld(uint32_t ta, uint32_t (&t)[64]) { ldtm64(ta, t); } };
constexpr int kLongNB = 5;                            // c < 160 covers bandwidth2 / R0 <= 160
constexpr int kLongKeep = 32 * kLongNB;               // kept outputs per side of each sub-spectrum
// The kept outputs of sub-sequence d sit in the upper half of region d, shifted by d * (128 / R) bytes (R regions are
// interleaved over consecutive lanes in the split): the regions are 16 KB apart, so without the shift the R lanes that
// read the same c from R different regions hit the same banks (R-way conflicts on every split load).
#ifndef USC_LONG_PAD
#define USC_LONG_PAD 1                                  // exchange tile of the core with padded rows (8448 bytes) instead of the XOR swizzle
#endif
template <int R, bool PAD = (USC_LONG_PAD != 0)> __device__ __forceinline__ constexpr uint32_t keep_off(uint32_t d) { return (PAD ? 8448u : 8192u) + d * (128u / R); }

struct long_params {
    const void* pcm; size_t nframes; uint32_t n;      // n real samples per frame (2048 * R0)
    const float4* chirp_ud; const float2* hann; const float2* tw_master;   // master: (cos, -sin)(2 pi j / n), j < n
    const float2* tw_pass;                            // W_1024^(a d), [d][a]
    uint32_t bandwidth2;
    float* mag_up; uint32_t* idx_up; float* mag_down; uint32_t* idx_down; uint8_t* bit;
};

// FR frames per CTA (one group of R0 warps each): the 4096-point form carries two, so that its CTA has four warps — one per
// TMEM lane quadrant — and two CTAs fill an SM like the 8192-point form.
template <int R0> struct long_geom {
    static constexpr int FR = R0 == 2 ? 2 : 1, T = 32 * R0, threads = FR * T;
    static constexpr bool staged = USC_LONG_STAGE != 0, tmem = USC_LONG_TMEM != 0;
};
template <int R0> struct long_smem {
    using G = long_geom<R0>;
    // pass twiddles | per group: { per warp d a 16 KB region: sub-sequence d as (re pair[1024], im pair[1024]); later tile
    // (8 KB) + kept outputs | PCM stage of one frame (filled by TMA one frame ahead) } | reduction slots | mbarriers | TMEM slot
    static constexpr int tw = 0, group0 = 8192, region = 16384, stage = R0 * region, group_bytes = stage + (G::staged ? R0 * 8192 : 0),
                         red = group0 + G::FR * group_bytes, bar = red + G::FR * R0 * 16, tslot = bar + G::FR * 8, total = tslot + 8;
};

__device__ __forceinline__ void group_sync(int group, int nthreads) {     // named barrier 1 + group: the warps of one frame
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(nthreads) : "memory");
}

template <typename PCM, int R0>
__device__ __forceinline__ void demod_long_body(const long_params& p) {
    using L = long_smem<R0>;
    using G = long_geom<R0>;
    using V2 = typename vec2<PCM>::type;
    constexpr int T = G::T, FR = G::FR;
    constexpr bool kStage = G::staged, kTmem = G::tmem;
    using TM = long_tmem<R0>;
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw + L::tw);
    const int lane = threadIdx.x & 31, cwarp = threadIdx.x >> 5;           // warp of the CTA (TMEM quadrant = cwarp & 3)
    const int group = FR == 1 ? 0 : cwarp / R0, warp = cwarp - group * R0, tid = threadIdx.x - group * T;   // within the frame's group
    unsigned char* gbase = s_raw + L::group0 + group * L::group_bytes;     // this group's regions and stage
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + L::bar) + group;
    constexpr uint32_t kFrameBytes = 2048u * R0 * 4u;
    const uint32_t nc = 1024u * R0;                    // complex length
    const PCM* pcm = static_cast<const PCM*>(p.pcm);
    const size_t f0 = (size_t) blockIdx.x * FR + group, fstep = (size_t) gridDim.x * FR;
    auto fetch = [&](size_t f) {                       // one frame of PCM into the stage: 1-D TMA bulk copies, 16 KB each
        mbar_expect_tx(bar, kFrameBytes);
#pragma unroll
        for (uint32_t o = 0; o < kFrameBytes; o += 16384u)
            bulk_g2s(gbase + L::stage + o, reinterpret_cast<const char*>(pcm + f * p.n) + o, kFrameBytes < 16384u ? kFrameBytes : 16384u, bar);
    };
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (kStage && f0 < p.nframes) fetch(f0);
    }
    for (int i = threadIdx.x; i < 1024; i += G::threads) s_tw[i] = p.tw_pass[i];
    uint32_t tq = 0;
    if (kTmem) {
        uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_raw + L::tslot);
        if (cwarp == 0) tmem_alloc<TM::cols>(s_tslot);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        tq = tmem_quadrant(*s_tslot, cwarp) + (cwarp >> 2) * 256u;
#pragma unroll 1
        for (int i = 0; i < TM::rounds; ++i) {            // this thread's table row, round by round
            const uint32_t a = tid + T * i;
            uint32_t v[TM::per_round];
#pragma unroll
            for (int b = 0; b < R0; ++b) {
                const uint32_t m = a + 1024u * b;
                const float4 c = __ldg(p.chirp_ud + m);
                const float2 w = __ldg(p.hann + m);
                v[TM::c_ud + 4 * b] = __float_as_uint(c.x); v[TM::c_ud + 4 * b + 1] = __float_as_uint(c.y);
                v[TM::c_ud + 4 * b + 2] = __float_as_uint(c.z); v[TM::c_ud + 4 * b + 3] = __float_as_uint(c.w);
                v[TM::c_hann + 2 * b] = __float_as_uint(w.x); v[TM::c_hann + 2 * b + 1] = __float_as_uint(w.y);
                const float2 z = __ldg(p.tw_master + (size_t) a * b * 2u);      // W_nc^(a b); b = 0 gives (1, -0), kept as padding
                v[TM::c_tw + 2 * b] = __float_as_uint(z.x); v[TM::c_tw + 2 * b + 1] = __float_as_uint(z.y);
            }
#pragma unroll
            for (int c0 = 0; c0 < TM::per_round; c0 += 8) {
                const uint32_t v8[8] = {v[c0], v[c0 + 1], v[c0 + 2], v[c0 + 3], v[c0 + 4], v[c0 + 5], v[c0 + 6], v[c0 + 7]};
                sttm8(tq + TM::per_round * i + c0, v8);
            }
        }
        sttm_wait();
        tmem_fence_before_sync();
    }
    __syncthreads();
    if (kTmem) tmem_fence_after_sync();
    const float one = s_tw[lane].x;                    // W^0 = 1.0f from the table: opaque to the compiler
    const uint32_t bw2 = p.bandwidth2;
    float2 w_split[kLongNB];                           // split twiddles W_n^k of this thread's bins k = tid + T j: frame-invariant
#pragma unroll
    for (int j = 0; j < kLongNB; ++j) w_split[j] = p.tw_master[min((uint32_t) (tid + T * j), bw2 - 1u)];

    uint32_t parity = 0;
    for (size_t f = f0; f < p.nframes; f += fstep) {
        const V2* stage = kStage ? reinterpret_cast<const V2*>(gbase + L::stage) : reinterpret_cast<const V2*>(pcm + f * p.n);
        if (kStage) {
            mbar_wait(bar, parity);                      // this frame's PCM has landed in the stage
            parity ^= 1u;
        } else if (f + fstep < p.nframes) {              // this group's next frame towards L2 while the current one computes
            const char* nxt = reinterpret_cast<const char*>(pcm + (f + fstep) * p.n);
            constexpr uint32_t per_thread = 2048u * R0 * 4u / T;                             // 256 bytes
#pragma unroll
            for (uint32_t o = 0; o < per_thread; o += 128u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t) tid * per_thread + o));
        }
        // ---- level 0: radix-R0 over b, twiddle, park sub-sequence d ----
        // The PCM loads of kLongUnroll rounds are issued together, ahead of the table loads (a TMEM round trip is an
        // asm statement the compiler does not move loads across).
        constexpr int kU = kLongUnroll < 32 / R0 ? kLongUnroll : 32 / R0;
#pragma unroll 1
        for (int i0 = 0; i0 < 32 / R0; i0 += kU) {
            V2 raw[kU][R0];
#pragma unroll
            for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int b = 0; b < R0; ++b) raw[u][b] = stage[tid + T * (i0 + u) + 1024u * b];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u;
                const uint32_t a = tid + T * i;
                float2 re[R0], im[R0];
                float2 twd[R0];
                if (kTmem) {                             // tables from this thread's TMEM row: one load per round
                    uint32_t t[TM::per_round];
                    ldtm_n<TM::per_round>::ld(tq + TM::per_round * i, t);
#pragma unroll
                    for (int b = 0; b < R0; ++b) {
                        const float x0 = pcm_to_float(raw[u][b].x), x1 = pcm_to_float(raw[u][b].y);
                        // packed window multiply; the first butterfly stage takes the products as FMAs by 1.0 (usc_arith.cuh)
                        const float2 tr = __fmul2_rn(make_float2(__uint_as_float(t[TM::c_ud + 4 * b]), __uint_as_float(t[TM::c_ud + 4 * b + 1])), bc2(x0));
                        const float2 ti = __fmul2_rn(make_float2(__uint_as_float(t[TM::c_ud + 4 * b + 2]), __uint_as_float(t[TM::c_ud + 4 * b + 3])), bc2(x1));
                        re[b] = __fmul2_rn(tr, bc2(__uint_as_float(t[TM::c_hann + 2 * b])));
                        im[b] = __fmul2_rn(ti, bc2(__uint_as_float(t[TM::c_hann + 2 * b + 1])));
                        twd[b] = make_float2(__uint_as_float(t[TM::c_tw + 2 * b]), __uint_as_float(t[TM::c_tw + 2 * b + 1]));
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < R0; ++b) {
                        const uint32_t m = a + 1024u * b;
                        const float x0 = pcm_to_float(raw[u][b].x), x1 = pcm_to_float(raw[u][b].y);
                        const float4 c = __ldg(p.chirp_ud + m);
                        const float2 w = __ldg(p.hann + m);
                        const float2 tr = __fmul2_rn(make_float2(c.x, c.y), bc2(x0)), ti = __fmul2_rn(make_float2(c.z, c.w), bc2(x1));
                        re[b] = __fmul2_rn(tr, bc2(w.x));
                        im[b] = __fmul2_rn(ti, bc2(w.y));
                        twd[b] = __ldg(p.tw_master + (size_t) a * b * 2u);          // W_nc^(a d) = W_n^(2 a d)
                    }
                }
                fft_base2_prod<R0>(re, im, one);
#pragma unroll
                for (int d = 0; d < R0; ++d) {
                    float2 xr = re[d], xi = im[d];
                    if (d != 0) cmul2(re[d], im[d], twd[d].x, twd[d].y, xr, xi);
                    float2* reg = reinterpret_cast<float2*>(gbase + d * L::region);
                    reg[a] = xr;
                    reg[1024 + a] = xi;
                }
            }
        }
        group_sync(group, T);                            // sub-sequences parked; every thread of the group is done with the stage
        if (kStage && tid == 0 && f + fstep < p.nframes) fetch(f + fstep);   // next frame arrives under the core and the split
        // ---- 1024-point packed core on sub-sequence `warp` ----
        float2* reg = reinterpret_cast<float2*>(gbase + warp * L::region);
        {
            float2 re[32], im[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                re[b] = reg[lane + 32 * b];
                im[b] = reg[1024 + lane + 32 * b];
            }
            __syncwarp();
            fft1024_pair<false, USC_LONG_PAD != 0>(re, im, reg, s_tw, lane);           // first 8 KB of the region is now the exchange tile
            // keep Y_d[c] for c < 160 (elements 0..4) and c >= 864 (elements 27..31)
            float4* keep = reinterpret_cast<float4*>(gbase + warp * L::region + keep_off<R0>(warp));
#pragma unroll
            for (int j = 0; j < kLongNB; ++j) {
                keep[lane + 32 * j] = make_float4(re[j].x, re[j].y, im[j].x, im[j].y);
                keep[kLongKeep + lane + 32 * j] = make_float4(re[32 - kLongNB + j].x, re[32 - kLongNB + j].y,
                                                              im[32 - kLongNB + j].x, im[32 - kLongNB + j].y);
            }
        }
        group_sync(group, T);
        // ---- split, squared magnitude, exact arg-max over k < bw2 ----
        auto Y = [&](uint32_t k) -> float4 {             // Z[k] = Y_{k mod R0}[k / R0], only kept ranges are asked for
            const uint32_t d = k & (R0 - 1u), c = k / R0;
            const float4* kp = reinterpret_cast<const float4*>(gbase + d * L::region + keep_off<R0>(d));
            return c < (uint32_t) kLongKeep ? kp[c] : kp[kLongKeep + (c - (1024u - kLongKeep))];
        };
        float pu[kLongNB], pd[kLongNB];
        uint32_t kk[kLongNB];
        bool ok[kLongNB];
#pragma unroll
        for (int j = 0; j < kLongNB; ++j) {
            const uint32_t k = tid + T * j;                                       // bw2 <= 160 R0 = kLongNB T
            kk[j] = k;
            ok[j] = k < bw2;
            const uint32_t kc = ok[j] ? k : 0u;
            const float4 zk = Y(kc);
            const float4 zc = Y(kc == 0u ? 0u : nc - kc);
            float2 xr, xi;
            const float2 w = w_split[j];                                          // W_n^k = (cos, -sin)
            rfft_split2(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w), make_float2(zc.x, zc.y),
                        make_float2(zc.z, zc.w), w.x, -w.y, xr, xi);
            if (kc == 0u) {                                                       // packed bin 0 = (X[0], X[N/2])
                xr = __fadd2_rn(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w));
                xi = __fadd2_rn(make_float2(zk.x, zk.y), neg2(make_float2(zk.z, zk.w)));
            }
            const float2 pw = __ffma2_rn(xr, xr, __fmul2_rn(xi, xi));
            pu[j] = pw.x;
            pd[j] = pw.y;
        }
        // every warp finds the exact (largest root, first index attaining it) of ITS bins with one square root per
        // hypothesis; combining the R0 results by value, then index, is exact for the frame
        float bu, bd;
        uint32_t iu, id;
        argmax_exact2<kLongNB>(pu, pd, kk, ok, bu, iu, bd, id);
        float4* red = reinterpret_cast<float4*>(s_raw + L::red) + group * R0;
        if (lane == 0) red[warp] = make_float4(bu, __uint_as_float(iu), bd, __uint_as_float(id));
        group_sync(group, T);
        if (tid == 0) {
            for (int w2 = 1; w2 < R0; ++w2) {
                const float4 v = red[w2];
                argmax_combine(bu, iu, v.x, __float_as_uint(v.y));
                argmax_combine(bd, id, v.z, __float_as_uint(v.w));
            }
            if (p.mag_up) p.mag_up[f] = bu;
            if (p.idx_up) p.idx_up[f] = iu;
            if (p.mag_down) p.mag_down[f] = bd;
            if (p.idx_down) p.idx_down[f] = id;
            if (p.bit) p.bit[f] = bd > bu ? 0 : 1;
        }
    }
    if (kTmem) {
        tmem_fence_before_sync();
        __syncthreads();
        if (cwarp == 0) tmem_dealloc<TM::cols>(*reinterpret_cast<uint32_t*>(s_raw + L::tslot));
    }
}

#ifndef USC_LONG_MAXREG4
#define USC_LONG_MAXREG4 224
#endif
template <typename PCM, int R0>
__global__ void __launch_bounds__(long_geom<R0>::threads, 1) k_demod_long(long_params p) { demod_long_body<PCM, R0>(p); }
// The 128-thread forms (8192 points; 4096 points with two frames per CTA) run two CTAs per SM: 2 x 107 KB of shared memory,
// 2 x 256 TMEM columns, and at most 224 registers per thread — an explicit cap, because launch bounds of 128 threads let
// ptxas take up to 255 (it took 238-244).
template <typename PCM, int R0>
__global__ void __maxnreg__(USC_LONG_MAXREG4) k_demod_long4(long_params p) { demod_long_body<PCM, R0>(p); }
template <typename PCM, int R0> struct long_kernel { static constexpr auto fn = k_demod_long<PCM, R0>; };
template <typename PCM> struct long_kernel<PCM, 4> { static constexpr auto fn = k_demod_long4<PCM, 4>; };
template <typename PCM> struct long_kernel<PCM, 2> { static constexpr auto fn = k_demod_long4<PCM, 2>; };

template <typename PCM, int R0>
static cudaError_t launch_long_t(const long_params& p, int num_sms, cudaStream_t st) {
    constexpr auto kernel = long_kernel<PCM, R0>::fn;
    using G = long_geom<R0>;
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    const int smem = long_smem<R0>::total;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    // Resident CTAs per SM from the kernel's own budget: the occupancy calculator reports ONE for any kernel that allocates
    // tensor memory (it cannot know how many columns a CTA takes), although two CTAs with 256 columns each do run side by
    // side (ncu: 8 warps per SM active).  128-thread forms: two CTAs; the 8-warp form (16384 points, 512 columns): one.
    const int per_sm = G::threads == 128 ? 2 : 1;
    const size_t want = (p.nframes + G::FR - 1) / G::FR;
    size_t ctas = want < (size_t) num_sms * per_sm ? want : (size_t) num_sms * per_sm;
    kernel<<<(int) ctas, G::threads, smem, st>>>(p);
    return cudaGetLastError();
}

// ---- N = 32768 / 65536: a thread-block cluster owns a frame -------------------------------------------
// Plan [R0, 32, 32] with R0 = 16 or 32.  The R0 sub-sequences (16 KB each) no longer fit one SM, so the cluster's
// distributed shared memory holds them: CL CTAs of eight warps, eight sub-sequences each (CL = R0 / 8).
//   level 0   CTA r, thread t takes a = (1024 / CL) r + 256 i + t: gathers z[a + 1024 b], b < R0, from global
//             memory, de-chirp x Hann for both hypotheses, radix-R0 butterfly, x W^(a d) (frame-invariant, kept
//             in shared memory), and stores element d into the owner CTA's copy of sub-sequence d —
//             16-byte remote stores over DSMEM
//   core      after a cluster barrier, each warp runs the packed 32x32 core on one local sub-sequence
//   split     after a second barrier each CTA takes the bins k = R0 c + d of its own d; the partner
//             Z[nc - k] sits in sub-sequence (R0 - d) mod R0, usually another CTA's: remote loads
//   result    the partial arg-max results meet in CTA 0 (remote stores), third barrier, one thread
//             writes the frame's outputs
// W = warps (sub-sequences) per CTA is a template parameter: the 4-warp forms with two CTAs per SM were measured
// slower (65536 points as 8 x 4: 8 %; 16384 points as a 2 x 4 cluster against the plain 8-warp CTA above: 9 %).
template <int R0, int W> struct cl_smem {                                         // W warps (= sub-sequences) per CTA
    static constexpr int T = 32 * W, CL = R0 / W, NR = 1024 / CL / T, SH = W == 8 ? 3 : 2;   // threads, cluster size, level-0 rounds per thread
    // pass twiddles | W sub-sequences | this CTA's share of the frame's PCM, [b][a] (filled by TMA one frame ahead) |
    // result slots of the cluster (CTA 0) + reduction scratch | mbarrier | TMEM slot
    static constexpr int tw = 0, sub = 8192, region = 16384, stage = sub + W * region, stage_bytes = R0 * (1024 / CL) * 8,
                         red = stage + stage_bytes, bar = red + 512, tslot = bar + 8, total = tslot + 8;
    // TMEM columns of a thread (its own lane; warps 4..7 take columns 256..511), per level-0 round 8 R0 columns:
    // (up, down) chirp of m = a + 1024 b at 4 b | Hann at 4 R0 + 2 b | level-0 twiddle W^(a d) at 6 R0 + 2 d
    static constexpr int per_round = 8 * R0, c_ud = 0, c_hann = 4 * R0, c_tw = 6 * R0, t_cols = 512;
    static_assert(NR * per_round == 256, "a thread's table row is 256 TMEM columns");
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(const void* local, uint32_t rank) {       // shared::cluster address of a peer's copy
    uint32_t l = (uint32_t) __cvta_generic_to_shared(local), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(l), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifndef USC_CL_SPLITBAR
#define USC_CL_SPLITBAR 1
#endif
template <typename PCM, int R0, int W>
__global__ void __launch_bounds__(32 * W, 1) k_demod_cluster(long_params p, const float2* __restrict__ tw_l0) {
    using L = cl_smem<R0, W>;
    constexpr int CL = L::CL, NR = L::NR, T = L::T, SH = L::SH;
    constexpr uint32_t n = 2048u * R0;                  // real samples per frame; the complex length is nc = n / 2
    constexpr uint32_t kShare = 1024u / CL;             // this CTA's a-range
    using V2 = typename vec2<PCM>::type;
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw + L::tw);
    const V2* stage = reinterpret_cast<const V2*>(s_raw + L::stage);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + L::bar);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_rank();
    const PCM* pcm = static_cast<const PCM*>(p.pcm);
    // this CTA's share of frame f: for every b the kShare pairs a = rank kShare .. of z[a + 1024 b], one bulk copy per b
    auto fetch = [&](size_t f) {                        // called by warp 0
        if (lane == 0) mbar_expect_tx(bar, L::stage_bytes);
        __syncwarp();
        for (int b = lane; b < R0; b += 32)
            bulk_g2s(s_raw + L::stage + b * (kShare * 8u), pcm + f * n + 2u * (1024u * b + rank * kShare), kShare * 8u, bar);
    };
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 1024; i += T) s_tw[i] = p.tw_pass[i];
    uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_raw + L::tslot);
    if (warp == 0) tmem_alloc<L::t_cols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    if (warp == 0 && cluster_id_x() < p.nframes) fetch(cluster_id_x());
    const uint32_t tq = tmem_quadrant(*s_tslot, warp) + (warp >> 2) * 256u;
#pragma unroll 1
    for (int i = 0; i < NR; ++i) {                      // this thread's table row
        const uint32_t a = rank * kShare + i * T + tid;
#pragma unroll 1
        for (int b0 = 0; b0 < R0; b0 += 4) {
            uint32_t c[16], w[8], z[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t m = a + 1024u * (b0 + j);
                const float4 cv = __ldg(p.chirp_ud + m);
                const float2 wv = __ldg(p.hann + m);
                const float2 zv = __ldg(tw_l0 + (size_t) (b0 + j) * 1024u + a);     // W_nc^(a d), d = b0 + j (d = 0: 1)
                c[4 * j] = __float_as_uint(cv.x); c[4 * j + 1] = __float_as_uint(cv.y); c[4 * j + 2] = __float_as_uint(cv.z); c[4 * j + 3] = __float_as_uint(cv.w);
                w[2 * j] = __float_as_uint(wv.x); w[2 * j + 1] = __float_as_uint(wv.y);
                z[2 * j] = __float_as_uint(zv.x); z[2 * j + 1] = __float_as_uint(zv.y);
            }
            const uint32_t c0[8] = {c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]}, c1[8] = {c[8], c[9], c[10], c[11], c[12], c[13], c[14], c[15]};
            sttm8(tq + L::per_round * i + L::c_ud + 4 * b0, c0);
            sttm8(tq + L::per_round * i + L::c_ud + 4 * b0 + 8, c1);
            sttm8(tq + L::per_round * i + L::c_hann + 2 * b0, w);
            sttm8(tq + L::per_round * i + L::c_tw + 2 * b0, z);
        }
    }
    sttm_wait();
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const float one = s_tw[lane].x;                      // W^0 = 1.0f from the table: opaque to the compiler
    const uint32_t bw2 = p.bandwidth2;
    float2 w_split[kLongNB];                             // split twiddles of this thread's bins, frame-invariant too
#pragma unroll
    for (int j = 0; j < kLongNB; ++j) {
        const uint32_t k = (uint32_t) R0 * ((tid >> SH) + 32u * j) + rank * W + (tid & (W - 1u));
        w_split[j] = p.tw_master[k];
    }
    // peer addresses of the sub-sequence area and of CTA 0's result slots
    uint32_t peer_sub[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) peer_sub[r] = map_to_rank(s_raw + L::sub, r);
    const uint32_t red0 = map_to_rank(s_raw + L::red, 0);
    cluster_sync_all();                                  // every CTA of the cluster is resident before remote traffic

    // CTA 0 combines the R0 per-warp results of a frame: one slot per lane, then the warp-wide exact arg-max
    auto emit = [&](size_t f) {
        const float4 v = reinterpret_cast<const float4*>(s_raw + L::red)[lane & (R0 - 1)];
        float bu = v.x, bd = v.z;
        uint32_t iu = __float_as_uint(v.y), id = __float_as_uint(v.w);
        warp_argmax(bu, iu);
        warp_argmax(bd, id);
        if (lane == 0) {
            if (p.mag_up) p.mag_up[f] = bu;
            if (p.idx_up) p.idx_up[f] = iu;
            if (p.mag_down) p.mag_down[f] = bd;
            if (p.idx_down) p.idx_down[f] = id;
            if (p.bit) p.bit[f] = bd > bu ? 0 : 1;
        }
    };
    // The barrier that closes a frame is split (USC_CL_SPLITBAR): a thread ARRIVES when its work on the frame is done (reads
    // of the sub-spectra, its warp's result stored in CTA 0) and WAITS only before its first remote store of the next frame,
    // so the wait for the next PCM share and the next level-0 front end and butterflies run while slower CTAs finish
    // their split.  CTA 0 combines the per-warp results between the next frame's first and second barrier — after every
    // arrive of the closing barrier, before any warp can store the following results (those come after the second).
    uint32_t parity = 0;
    bool pending = false;                                // a frame's results wait in CTA 0 / its closing barrier is open
    size_t fprev = 0;
    for (size_t f = cluster_id_x(); f < p.nframes; f += cluster_count_x()) {
        mbar_wait(bar, parity);                          // this CTA's share of the frame has landed in the stage
        parity ^= 1u;
        // ---- level 0: radix-R0 over b in registers, twiddle, 16-byte remote stores into the owner's sub-sequence ----
#pragma unroll 1
        for (int i = 0; i < NR; ++i) {
            const uint32_t al = i * T + tid, a = rank * kShare + al;
            float2 re[R0], im[R0];
#pragma unroll
            for (int g = 0; g < R0 / 4; ++g) {           // table values of four b per TMEM round trip
                uint32_t c[16], w[8];
                ldtm16_8(tq + L::per_round * i + L::c_ud + 16 * g, c, tq + L::per_round * i + L::c_hann + 8 * g, w);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = 4 * g + j;
                    const V2 raw = stage[b * kShare + al];
                    const float x0 = pcm_to_float(raw.x), x1 = pcm_to_float(raw.y);
                    const float2 tr = __fmul2_rn(make_float2(__uint_as_float(c[4 * j]), __uint_as_float(c[4 * j + 1])), bc2(x0));
                    const float2 ti = __fmul2_rn(make_float2(__uint_as_float(c[4 * j + 2]), __uint_as_float(c[4 * j + 3])), bc2(x1));
                    re[b] = __fmul2_rn(tr, bc2(__uint_as_float(w[2 * j])));              // packed; first stage: FMAs by 1.0
                    im[b] = __fmul2_rn(ti, bc2(__uint_as_float(w[2 * j + 1])));
                }
            }
            fft_base2_prod<R0>(re, im, one);
#if USC_CL_SPLITBAR
            if (i == 0 && pending) cluster_wait();       // every peer has read the previous frame's sub-spectra
#endif
#pragma unroll
            for (int g = 0; g < R0 / 8; ++g) {           // level-0 twiddles, eight per TMEM round trip
                uint32_t t[16];
                ldtm16(tq + L::per_round * i + L::c_tw + 16 * g, t);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int d = 8 * g + j;
                    float2 xr = re[d], xi = im[d];
                    if (d != 0) cmul2(re[d], im[d], __uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]), xr, xi);
                    const uint32_t base = peer_sub[d >> SH] + (uint32_t) (d & (W - 1)) * L::region;
                    st_cluster_f4(base + a * 16u, make_float4(xr.x, xr.y, xi.x, xi.y));   // one 16-byte remote store per element
                }
            }
        }
        cluster_sync_all();                              // sub-sequences complete; every thread is done with the stage
        if (warp == 0 && f + cluster_count_x() < p.nframes) fetch(f + cluster_count_x());   // next share arrives under the core and the split
#if USC_CL_SPLITBAR
        if (pending && rank == 0 && warp == W - 1) emit(fprev);
#endif
        // ---- 1024-point packed core on sub-sequence W rank + warp ----
        {
            float2* reg = reinterpret_cast<float2*>(s_raw + L::sub + warp * L::region);
            float2 re[32], im[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const float4 v = reinterpret_cast<const float4*>(reg)[lane + 32 * b];       // (re pair, im pair)
                re[b] = make_float2(v.x, v.y);
                im[b] = make_float2(v.z, v.w);
            }
            __syncwarp();
            fft1024_pair(re, im, reg, s_tw, lane);                           // XOR-swizzled tile: the padded one measured 3 % slower here
            float4* keep = reinterpret_cast<float4*>(s_raw + L::sub + warp * L::region + keep_off<W, false>(warp));
#pragma unroll
            for (int j = 0; j < kLongNB; ++j) {
                keep[lane + 32 * j] = make_float4(re[j].x, re[j].y, im[j].x, im[j].y);
                keep[kLongKeep + lane + 32 * j] = make_float4(re[32 - kLongNB + j].x, re[32 - kLongNB + j].y,
                                                              im[32 - kLongNB + j].x, im[32 - kLongNB + j].y);
            }
        }
        cluster_sync_all();
        // ---- split, squared magnitude, arg-max over the bins of this CTA's sub-sequences ----
        // thread -> (dl = tid & 7, c = tid >> 3 + 32 j): bins of one c are spread over 8 threads; ascending k per thread.
        // Each WARP finds the exact (largest root, first index attaining it) of ITS bins with one square root per
        // hypothesis and hands it to CTA 0, which combines the CL W results by value, then index — exact for the frame.
        float pu[kLongNB], pd[kLongNB];
        uint32_t kk[kLongNB];
        bool ok[kLongNB];
        float4 zcs[kLongNB];                                 // the remote partner loads first, all in flight together
#pragma unroll
        for (int j = 0; j < kLongNB; ++j) {
            const uint32_t c = (tid >> SH) + 32u * j, dl = tid & (W - 1u), d = rank * W + dl;
            // nc - k = R0 (1024 - c) for d = 0 (c >= 1), else R0 (1023 - c) + (R0 - d); k = 0 pairs with itself (unused)
            const uint32_t d2 = ((uint32_t) R0 - d) & (uint32_t) (R0 - 1), c2 = d == 0 ? (c == 0 ? 1023u : 1024u - c) : 1023u - c;
            const uint32_t addr = peer_sub[d2 >> SH] + (d2 & (W - 1u)) * L::region + keep_off<W, false>(d2 & (W - 1u)) + (kLongKeep + (c2 - (1024u - kLongKeep))) * 16u;
            zcs[j] = ld_cluster_f4(addr);
        }
#pragma unroll
        for (int j = 0; j < kLongNB; ++j) {
            const uint32_t c = (tid >> SH) + 32u * j, dl = tid & (W - 1u), d = rank * W + dl, k = (uint32_t) R0 * c + d;
            kk[j] = k;
            ok[j] = k < bw2;
            const float4 zk = reinterpret_cast<const float4*>(s_raw + L::sub + dl * L::region + keep_off<W, false>(dl))[c];
            const float4 zc = zcs[j];
            const float2 w = w_split[j];
            float2 xr, xi;
            rfft_split2(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w), make_float2(zc.x, zc.y),
                        make_float2(zc.z, zc.w), w.x, -w.y, xr, xi);
            if (k == 0) {
                xr = __fadd2_rn(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w));
                xi = __fadd2_rn(make_float2(zk.x, zk.y), neg2(make_float2(zk.z, zk.w)));
            }
            const float2 pw = __ffma2_rn(xr, xr, __fmul2_rn(xi, xi));
            pu[j] = pw.x;
            pd[j] = pw.y;
        }
        float bu, bd;
        uint32_t iu, id;
        argmax_exact2<kLongNB>(pu, pd, kk, ok, bu, iu, bd, id);
        if (lane == 0) st_cluster_f4(red0 + (rank * W + warp) * 16u, make_float4(bu, __uint_as_float(iu), bd, __uint_as_float(id)));
#if USC_CL_SPLITBAR
        cluster_arrive();                                // everything this thread does for frame f is done: reads of the sub-spectra, its result
        pending = true;
        fprev = f;
#else
        cluster_sync_all();                              // results are in CTA 0; all remote reads of this frame are done
        if (rank == 0 && warp == 0) emit(f);
#endif
    }
#if USC_CL_SPLITBAR
    if (pending) cluster_wait();
#endif
    tmem_fence_before_sync();
    cluster_sync_all();                                  // no CTA leaves while a peer may still address its memory
#if USC_CL_SPLITBAR
    if (pending && rank == 0 && warp == 0) emit(fprev);  // the last frame's results arrived with the barrier above
#endif
    if (warp == 0) tmem_dealloc<L::t_cols>(*s_tslot);
}

template <typename PCM, int R0, int W>
static cudaError_t launch_cluster_t(const long_params& p, const float2* tw_l0, int num_sms, cudaStream_t st) {
    // The loop inside the kernel strides by the number of clusters launched, so launch exactly as many as
    // can be resident at once (fewer than SMs / cluster size: clusters do not straddle GPCs) — a cluster
    // left for a second wave would run its whole share after everyone else has finished.
    using L = cl_smem<R0, W>;
    static per_device<int> max_clusters_pd;
    int& max_clusters = max_clusters_pd.get();
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = L::CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(L::T);
    cfg.dynamicSmemBytes = L::total;
    cfg.stream = st;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (!max_clusters) {
        cudaError_t e = cudaFuncSetAttribute(k_demod_cluster<PCM, R0, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
        if (e != cudaSuccess) return e;
        cfg.gridDim = dim3((unsigned) (num_sms * (W == 8 ? 1 : 2) / L::CL * L::CL));
        int nmax = 0;
        e = cudaOccupancyMaxActiveClusters(&nmax, k_demod_cluster<PCM, R0, W>, &cfg);
        if (e != cudaSuccess) return e;
        if (nmax < 1) return cudaErrorLaunchOutOfResources;
        max_clusters = nmax;
    }
    size_t clusters = (size_t) max_clusters;
    if (clusters > p.nframes) clusters = p.nframes;
    cfg.gridDim = dim3((unsigned) (clusters * L::CL));
    return cudaLaunchKernelEx(&cfg, k_demod_cluster<PCM, R0, W>, p, tw_l0);
}

cudaError_t launch_demod_long32(const void* pcm, uint32_t pcm_format, size_t nframes, uint32_t n, const float2* chirp_ud,
                                const float2* hann, const float2* tw_master, const float2* tw_pass, const float2* tw_l0,
                                uint32_t bandwidth2, float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down,
                                uint8_t* bit, int num_sms, cudaStream_t st) {
    long_params p{pcm, nframes, n, reinterpret_cast<const float4*>(chirp_ud), hann, tw_master, tw_pass, bandwidth2,
                  mag_up, idx_up, mag_down, idx_down, bit};
    if (n == 65536u) return pcm_format == 1u ? launch_cluster_t<int32_t, 32, 8>(p, tw_l0, num_sms, st) : launch_cluster_t<float, 32, 8>(p, tw_l0, num_sms, st);
    if (n == 32768u) return pcm_format == 1u ? launch_cluster_t<int32_t, 16, 8>(p, tw_l0, num_sms, st) : launch_cluster_t<float, 16, 8>(p, tw_l0, num_sms, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_demod_long(const void* pcm, uint32_t pcm_format, size_t nframes, uint32_t n, const float2* chirp_ud,
                              const float2* hann, const float2* tw_master, const float2* tw_pass, uint32_t bandwidth2,
                              float* mag_up, uint32_t* idx_up, float* mag_down, uint32_t* idx_down, uint8_t* bit,
                              int num_sms, cudaStream_t st) {
    long_params p{pcm, nframes, n, reinterpret_cast<const float4*>(chirp_ud), hann, tw_master, tw_pass, bandwidth2,
                  mag_up, idx_up, mag_down, idx_down, bit};
    if (n == 4096) return pcm_format == 1u ? launch_long_t<int32_t, 2>(p, num_sms, st) : launch_long_t<float, 2>(p, num_sms, st);
    if (n == 8192) return pcm_format == 1u ? launch_long_t<int32_t, 4>(p, num_sms, st) : launch_long_t<float, 4>(p, num_sms, st);
    if (n == 16384) return pcm_format == 1u ? launch_long_t<int32_t, 8>(p, num_sms, st) : launch_long_t<float, 8>(p, num_sms, st);
    return cudaErrorInvalidValue;
}

}  // namespace usc
