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
// Both hypotheses (up / down) share the PCM load, the Hann load and the twiddle loads (f32x2 halves).
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"
#include "usc_tmem.cuh"

namespace usc {

constexpr int kWarpsPerCta = 4;
#ifndef USC_K1_TMEM
#define USC_K1_TMEM 1                                   // tables in tensor memory (0: in shared memory, the round-1 form)
#endif
#ifndef USC_K1_WS_REGS
#define USC_K1_WS_REGS 0                                // split twiddles of the bins below bandwidth2 in shared memory (1: in registers)
#endif
#ifndef USC_K1_TW32
#define USC_K1_TW32 0                                   // eight inter-pass twiddles per TMEM round trip (1: sixteen, 2: four)
#endif
#ifndef USC_K1_PAD
#define USC_K1_PAD 1                                    // exchange tile with padded rows (0: XOR-swizzled 32 x 32)
#endif
#ifndef USC_K1_TG
#define USC_K1_TG 4                                     // front-end rows per TMEM load group (4 or 8)
#endif
// ---- dual-hypothesis kernel: both hypotheses ride in the halves of f32x2 registers ---------------
// (.x = up-chirp, .y = down-chirp).  The PCM, the Hann table and the inter-pass twiddles are loaded
// once for both; the two 32-point register FFTs issue as FADD2/FFMA2.
//
// Persistent: ONE CTA of 8 warps per SM, every warp loops over frames.  Shared memory (164 KB):
//   per warp: exchange tile 8 KB + PCM stage 8 KB | mbarriers
// The PCM of a warp's NEXT frame is fetched by a 1-D TMA bulk copy (cp.async.bulk, mbarrier
// completion) issued right after the current frame has been pulled into registers, so DRAM latency
// is hidden behind a whole frame of arithmetic.  The (up, down) chirp table, the Hann table and the inter-pass
// twiddles are read from TENSOR MEMORY (usc_tmem.cuh: one lane-private row of 256 columns per lane, filled once per
// CTA), which takes 254 of the 633 wavefronts a frame used to cost off the L1/shared data path (128 B per clock per
// SM): that path now carries PCM 8 KB + exchange 2 x 16 KB per frame, and the kernel is bound by the fp32 pipe.
constexpr int kDualWarps = 8;                         // warps per CTA of the pair kernel below

// Shared memory of the dual kernel with W warps: twiddles 8 KB | (up,down) table 16 KB | Hann 8 KB |
// per warp: XOR-swizzled 32x32 float2 tile 8 KB + PCM stage 8 KB | mbarriers.
template <int W> struct dual_smem {
#if USC_K1_TMEM
    static constexpr int tw = 0, ud = 0, hann = 0, warp = 0, warp_bytes = (USC_K1_PAD ? 8704 : 8192) + 8192,     // the three tables live in TMEM
#else
    static constexpr int tw = 0, ud = 8192, hann = ud + 16384, warp = hann + 8192, warp_bytes = 8192 + 8192,
#endif
                         bar = warp + W * warp_bytes, ws = bar + 128, total = ws + 16 * 32 * 8;   // ws: split twiddles of the bins below bandwidth2
    static constexpr int tslot = bar + 120;              // TMEM base address written by tcgen05.alloc
    // TMEM columns of lane a (one replica per lane quadrant): (up, down) chirp of m = a + 32 b at 4 b | Hann at 128 + 2 b |
    // inter-pass twiddle W_1024^(a d) at 192 + 2 d
    static constexpr int t_ud = 0, t_hann = 128, t_tw = 192, t_cols = 256;
};

template <typename PCM, int NB, int W>
__global__ void __launch_bounds__(W * 32, 1) k_demod2048(demod_params p) {
    using L = dual_smem<W>;
    extern __shared__ __align__(128) unsigned char s_raw[];
    float2* s_tw = reinterpret_cast<float2*>(s_raw + L::tw);
    float4* s_ud = reinterpret_cast<float4*>(s_raw + L::ud);
    float2* s_hann = reinterpret_cast<float2*>(s_raw + L::hann);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* tile = reinterpret_cast<float2*>(s_raw + L::warp + warp * L::warp_bytes);
    using V2 = typename vec2<PCM>::type;
    V2* xstage = reinterpret_cast<V2*>(s_raw + L::warp + warp * L::warp_bytes + (USC_K1_PAD ? 8704 : 8192));
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + L::bar) + warp;

    const size_t nwarps = (size_t) gridDim.x * W;
    size_t f = (size_t) blockIdx.x * W + warp;
    const PCM* pcm = static_cast<const PCM*>(p.pcm);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (f < p.nframes) {
            mbar_expect_tx(bar, 8192);
            bulk_g2s(xstage, pcm + f * 2048, 8192, bar);
        }
    }
#if USC_K1_TMEM
    uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_raw + L::tslot);
    if (warp == 0) tmem_alloc<L::t_cols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tq = tmem_quadrant(*s_tslot, warp);
    if (warp < 4) {                                                   // warp q fills lane quadrant q; warps q and q + 4 read it
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += 8) {                          // eight rows per round: 24 independent loads in flight
            float4 c[8];
            float2 w[8], z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = reinterpret_cast<const float4*>(p.chirp_ud)[lane + 32 * (b0 + j)];
                w[j] = p.hann[lane + 32 * (b0 + j)];
                z[j] = p.tw_pass[(b0 + j) * 32 + lane];
            }
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const uint32_t v[8] = {__float_as_uint(c[j].x), __float_as_uint(c[j].y), __float_as_uint(c[j].z), __float_as_uint(c[j].w),
                                       __float_as_uint(c[j + 1].x), __float_as_uint(c[j + 1].y), __float_as_uint(c[j + 1].z), __float_as_uint(c[j + 1].w)};
                sttm8(tq + L::t_ud + 4 * (b0 + j), v);
            }
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const uint32_t v[8] = {__float_as_uint(w[j].x), __float_as_uint(w[j].y), __float_as_uint(w[j + 1].x), __float_as_uint(w[j + 1].y),
                                       __float_as_uint(w[j + 2].x), __float_as_uint(w[j + 2].y), __float_as_uint(w[j + 3].x), __float_as_uint(w[j + 3].y)};
                const uint32_t t[8] = {__float_as_uint(z[j].x), __float_as_uint(z[j].y), __float_as_uint(z[j + 1].x), __float_as_uint(z[j + 1].y),
                                       __float_as_uint(z[j + 2].x), __float_as_uint(z[j + 2].y), __float_as_uint(z[j + 3].x), __float_as_uint(z[j + 3].y)};
                sttm8(tq + L::t_hann + 2 * (b0 + j), v);
                sttm8(tq + L::t_tw + 2 * (b0 + j), t);
            }
        }
        sttm_wait();
    }
    const float one = p.tw_pass[lane].x;                              // W^0 = 1.0f, read from the table: opaque to the compiler
#else
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = p.tw_pass[i];
        s_ud[i] = reinterpret_cast<const float4*>(p.chirp_ud)[i];
        s_hann[i] = p.hann[i];
    }
#endif
    float2* s_ws = reinterpret_cast<float2*>(s_raw + L::ws);         // split twiddles: fetched when the epilogue needs them
    for (int i = threadIdx.x; i < NB * 32; i += blockDim.x) s_ws[i] = p.tw_split[i];
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();

#if USC_K1_WS_REGS
    float2 ws_r[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) ws_r[d1] = p.tw_split[lane + 32 * d1];
#endif
    uint32_t parity = 0;
    for (; f < p.nframes; f += nwarps) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 re[32], im[32];                                        // (.x, .y) = (up, down)
        // ((x*c)*w) for both hypotheses, packed.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even
        // with -fmad=false), which would fuse this product into the first butterfly's additions and break bit-parity:
        // the first butterfly stage is therefore written as FMAs by 1.0 (fft_base2_prod), which cannot be contracted.
#if USC_K1_TMEM
#pragma unroll
        for (int g = 0; g < 32 / USC_K1_TG; ++g) {                    // table values of USC_K1_TG rows per TMEM round trip
#if USC_K1_TG == 8
            uint32_t c[32], w[16];
            ldtm32_16(tq + L::t_ud + 32 * g, c, tq + L::t_hann + 16 * g, w);
#elif USC_K1_TG == 2
            uint32_t c[8], w[4];
            ldtm8_4(tq + L::t_ud + 8 * g, c, tq + L::t_hann + 4 * g, w);
#else
            uint32_t c[16], w[8];
            ldtm16_8(tq + L::t_ud + 16 * g, c, tq + L::t_hann + 8 * g, w);
#endif
#pragma unroll
            for (int j = 0; j < USC_K1_TG; ++j) {
                const int b = USC_K1_TG * g + j;
                const V2 raw = xstage[lane + 32 * b];
                const float x0 = pcm_to_float(raw.x), x1 = pcm_to_float(raw.y);
                const float2 tr = __fmul2_rn(make_float2(__uint_as_float(c[4 * j]), __uint_as_float(c[4 * j + 1])), bc2(x0));
                const float2 ti = __fmul2_rn(make_float2(__uint_as_float(c[4 * j + 2]), __uint_as_float(c[4 * j + 3])), bc2(x1));
                re[b] = __fmul2_rn(tr, bc2(__uint_as_float(w[2 * j])));
                im[b] = __fmul2_rn(ti, bc2(__uint_as_float(w[2 * j + 1])));
            }
        }
#else
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            V2 raw = xstage[m];
            float x0 = pcm_to_float(raw.x), x1 = pcm_to_float(raw.y);
            float4 c = s_ud[m];                                       // (up[2m], down[2m], up[2m+1], down[2m+1])
            float2 w = s_hann[m];
            const float2 tr = __fmul2_rn(make_float2(c.x, c.y), bc2(x0)), ti = __fmul2_rn(make_float2(c.z, c.w), bc2(x1));
            re[b] = __fmul2_rn(tr, bc2(w.x));
            im[b] = __fmul2_rn(ti, bc2(w.y));
        }
#endif
        __syncwarp();                                                 // every lane has consumed the stage
        if (lane == 0 && f + nwarps < p.nframes) {                    // refill it with this warp's next frame
            mbar_expect_tx(bar, 8192);
            bulk_g2s(xstage, pcm + (f + nwarps) * 2048, 8192, bar);
        }
#if USC_K1_TMEM
        fft_base2_prod<32>(re, im, one);
#if USC_K1_TW32 == 2
#pragma unroll
        for (int g = 0; g < 8; ++g) {                                 // inter-pass twiddle, both hypotheses: 4 per TMEM round trip
            uint32_t t[8];
            ldtm8(tq + L::t_tw + 8 * g, t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = 4 * g + j;
                if (d == 0) continue;
                float2 tr, ti;
                cmul2(re[d], im[d], __uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]), tr, ti);
                re[d] = tr;
                im[d] = ti;
            }
        }
#elif USC_K1_TW32
#pragma unroll
        for (int g = 0; g < 2; ++g) {                                 // inter-pass twiddle, both hypotheses: 16 per TMEM round trip
            uint32_t t[32];
            ldtm32(tq + L::t_tw + 32 * g, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int d = 16 * g + j;
                if (d == 0) continue;
                float2 tr, ti;
                cmul2(re[d], im[d], __uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]), tr, ti);
                re[d] = tr;
                im[d] = ti;
            }
        }
#else
#pragma unroll
        for (int g = 0; g < 4; ++g) {                                 // inter-pass twiddle, both hypotheses: 8 per TMEM round trip
            uint32_t t[16];
            ldtm16(tq + L::t_tw + 16 * g, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = 8 * g + j;
                if (d == 0) continue;
                float2 tr, ti;
                cmul2(re[d], im[d], __uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]), tr, ti);
                re[d] = tr;
                im[d] = ti;
            }
        }
#endif
#else
        fft_base2_prod<32>(re, im, s_tw[lane].x);                     // W^0 = 1.0f, read from the table: opaque to the compiler
#pragma unroll
        for (int d = 1; d < 32; ++d) {                                // inter-pass twiddle, both hypotheses
            const float2 w = s_tw[d * 32 + lane];
            float2 tr, ti;
            cmul2(re[d], im[d], w.x, w.y, tr, ti);
            re[d] = tr;
            im[d] = ti;
        }
#endif
        // exchange in two rounds (real parts, then imaginary parts): each 64-bit word is an (up, down)
        // register pair, so values land in place; XOR swizzle keeps both directions conflict-free
#if USC_K1_PAD
        // padded rows (33 float2): every store and load of a round is one base register plus an immediate offset
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * kTileStride + lane] = re[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) re[a] = tile[lane * kTileStride + a];
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * kTileStride + lane] = im[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) im[a] = tile[lane * kTileStride + a];
        __syncwarp();
#else
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * 32 + (lane ^ d)] = re[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) re[a] = tile[lane * 32 + (a ^ lane)];
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * 32 + (lane ^ d)] = im[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) im[a] = tile[lane * 32 + (a ^ lane)];
        __syncwarp();
#endif
        fft_base2<32>(re, im);
        float mu, md;
        uint32_t iu, id;
#if USC_K1_WS_REGS
        peak_window_pair<NB>(re, im, ws_r, lane, p.bandwidth2, mu, iu, md, id);
#else
        peak_window_pair_s<NB>(re, im, s_ws, lane, p.bandwidth2, mu, iu, md, id);
#endif
        if (lane == 0) {
            if (p.mag_up) p.mag_up[f] = mu;
            if (p.idx_up) p.idx_up[f] = iu;
            if (p.mag_down) p.mag_down[f] = md;
            if (p.idx_down) p.idx_down[f] = id;
            if (p.bit) p.bit[f] = md > mu ? 0 : 1;     // receiver/Src/main.c:523: down only if strictly greater
        }
    }
#if USC_K1_TMEM
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<L::t_cols>(*s_tslot);
#endif
}

// ---- single-hypothesis kernel: TWO FRAMES ride in the halves of the f32x2 registers ----------------
// dsp() for one hypothesis (receiver/Src/main.c:183-215) on aligned frames.  Same packed core as the
// dual kernel, but .x / .y carry frames 2i and 2i+1, so one chirp table serves both and each frame
// costs half of a dual-hypothesis frame: this is the fused "window + FFT + compression + peak" kernel
// whose time per 8 KB frame is closest to the HBM roofline.  The 16 KB of PCM for a warp's next
// frame pair arrive by one TMA bulk copy; to stay inside 227 KB of shared memory with 8 warps the
// exchange goes through an 8 KB tile in two rounds (real parts, then imaginary parts).
#ifndef USC_PAIR_TMEM
#define USC_PAIR_TMEM 1                                 // tables in tensor memory, as in the dual kernel
#endif
#if USC_PAIR_TMEM
constexpr int kPairSmemTabs = 0;
#else
constexpr int kPairSmemTabs = 3 * 8192;               // twiddles | chirp | Hann
#endif
constexpr int kPairWarpBytes = kTileFloat2 * 8 + 16384;   // padded 32x33 float2 tile + 2-frame PCM stage
constexpr int kPairSmemBar = kPairSmemTabs + kDualWarps * kPairWarpBytes;
constexpr int kPairSmemTotal = kPairSmemBar + kDualWarps * 8 + 8;
// TMEM columns of lane a: (chirp pair, Hann pair) of m = a + 32 b at 4 b | inter-pass twiddle W_1024^(a d) at 128 + 2 d
constexpr int kPairTcw = 0, kPairTtw = 128, kPairTcols = 256;

template <typename PCM, int NB>
__global__ void __launch_bounds__(kDualWarps * 32, 1) k_demod2048_pair(demod_params p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = s_raw + kPairSmemTabs + warp * kPairWarpBytes;
    using V2 = typename vec2<PCM>::type;
    V2* xstage = reinterpret_cast<V2*>(wbase);                       // 2 x 1024 pairs (16-byte aligned)
    float2* tile = reinterpret_cast<float2*>(wbase + 16384);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + kPairSmemBar) + warp;

    const size_t npairs = (p.nframes + 1) / 2;
    const size_t nwarps = (size_t) gridDim.x * kDualWarps;
    size_t q = (size_t) blockIdx.x * kDualWarps + warp;
    const PCM* pcm = static_cast<const PCM*>(p.pcm);
    auto pair_bytes = [&](size_t pr) -> uint32_t { return 2 * pr + 1 < p.nframes ? 16384u : 8192u; };
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (q < npairs) {
            mbar_expect_tx(bar, pair_bytes(q));
            bulk_g2s(xstage, pcm + q * 4096, pair_bytes(q), bar);
        }
    }
    const float2* chirp = p.updown ? p.chirp_up : p.chirp_down;
#if USC_PAIR_TMEM
    uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_raw + kPairSmemBar + kDualWarps * 8);
    if (warp == 0) tmem_alloc<kPairTcols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tq = tmem_quadrant(*s_tslot, warp);
    if (warp < 4) {                                                   // warp q fills lane quadrant q; warps q and q + 4 read it
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += 8) {
            float2 c[8], w[8], z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = chirp[lane + 32 * (b0 + j)];
                w[j] = p.hann[lane + 32 * (b0 + j)];
                z[j] = p.tw_pass[(b0 + j) * 32 + lane];
            }
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const uint32_t v[8] = {__float_as_uint(c[j].x), __float_as_uint(c[j].y), __float_as_uint(w[j].x), __float_as_uint(w[j].y),
                                       __float_as_uint(c[j + 1].x), __float_as_uint(c[j + 1].y), __float_as_uint(w[j + 1].x), __float_as_uint(w[j + 1].y)};
                sttm8(tq + kPairTcw + 4 * (b0 + j), v);
            }
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const uint32_t t[8] = {__float_as_uint(z[j].x), __float_as_uint(z[j].y), __float_as_uint(z[j + 1].x), __float_as_uint(z[j + 1].y),
                                       __float_as_uint(z[j + 2].x), __float_as_uint(z[j + 2].y), __float_as_uint(z[j + 3].x), __float_as_uint(z[j + 3].y)};
                sttm8(tq + kPairTtw + 2 * (b0 + j), t);
            }
        }
        sttm_wait();
    }
    const float one = p.tw_pass[lane].x;                              // W^0 = 1.0f, read from the table: opaque to the compiler
#else
    float2* s_tw = reinterpret_cast<float2*>(s_raw);
    float2* s_chirp = reinterpret_cast<float2*>(s_raw + 8192);
    float2* s_hann = reinterpret_cast<float2*>(s_raw + 16384);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = p.tw_pass[i];
        s_chirp[i] = chirp[i];
        s_hann[i] = p.hann[i];
    }
#endif
    float2 ws[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) ws[d1] = p.tw_split[lane + 32 * d1];
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();

    uint32_t parity = 0;
    for (; q < npairs; q += nwarps) {
        const bool two = 2 * q + 1 < p.nframes;
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 re[32], im[32];                                        // (.x, .y) = (frame 2q, frame 2q+1)
        // ((x*c)*w) on both frames at once: the per-lane table values broadcast to the two halves
        // (first butterfly stage as FMAs by 1.0: see the note in k_demod2048 about ptxas contracting packed mul + add)
#if USC_PAIR_TMEM
#pragma unroll
        for (int g = 0; g < 8; ++g) {                                 // table values of four rows per TMEM round trip
            uint32_t t[16];
            ldtm16(tq + kPairTcw + 16 * g, t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = 4 * g + j, m = lane + 32 * b;
                const V2 ra = xstage[m];
                const V2 rb = two ? xstage[1024 + m] : ra;
                const float2 tr = __fmul2_rn(make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x)), bc2(__uint_as_float(t[4 * j])));
                const float2 ti = __fmul2_rn(make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y)), bc2(__uint_as_float(t[4 * j + 1])));
                re[b] = __fmul2_rn(tr, bc2(__uint_as_float(t[4 * j + 2])));
                im[b] = __fmul2_rn(ti, bc2(__uint_as_float(t[4 * j + 3])));
            }
        }
#else
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            const V2 ra = xstage[m];
            const V2 rb = two ? xstage[1024 + m] : ra;
            const float2 c = s_chirp[m], w = s_hann[m];
            const float2 tr = __fmul2_rn(make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x)), bc2(c.x));
            const float2 ti = __fmul2_rn(make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y)), bc2(c.y));
            re[b] = __fmul2_rn(tr, bc2(w.x));
            im[b] = __fmul2_rn(ti, bc2(w.y));
        }
#endif
        __syncwarp();
        if (lane == 0 && q + nwarps < npairs) {
            mbar_expect_tx(bar, pair_bytes(q + nwarps));
            bulk_g2s(xstage, pcm + (q + nwarps) * 4096, pair_bytes(q + nwarps), bar);
        }
#if USC_PAIR_TMEM
        fft_base2_prod<32>(re, im, one);
#pragma unroll
        for (int g = 0; g < 4; ++g) {                                 // inter-pass twiddle, both halves: 8 per TMEM round trip
            uint32_t t[16];
            ldtm16(tq + kPairTtw + 16 * g, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = 8 * g + j;
                if (d == 0) continue;
                float2 tr, ti;
                cmul2(re[d], im[d], __uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]), tr, ti);
                re[d] = tr;
                im[d] = ti;
            }
        }
#else
        fft_base2_prod<32>(re, im, s_tw[lane].x);
#pragma unroll
        for (int d = 1; d < 32; ++d) {                                // inter-pass twiddle, both halves
            const float2 w = s_tw[d * 32 + lane];
            float2 tr, ti;
            cmul2(re[d], im[d], w.x, w.y, tr, ti);
            re[d] = tr;
            im[d] = ti;
        }
#endif
        // exchange in two rounds through the 8 KB tile, one per component: each 64-bit word is a
        // (frame 2q, frame 2q+1) register pair, so values land in place with no repacking moves
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * kTileStride + lane] = re[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) re[a] = tile[lane * kTileStride + a];
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 32; ++d) tile[d * kTileStride + lane] = im[d];
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; ++a) im[a] = tile[lane * kTileStride + a];
        __syncwarp();
        fft_base2<32>(re, im);
        float ma, mb;
        uint32_t ia, ib;
        peak_window_pair<NB>(re, im, ws, lane, p.bandwidth2, ma, ia, mb, ib);
        if (lane == 0) {
            float* mag = p.updown ? p.mag_up : p.mag_down;
            uint32_t* idx = p.updown ? p.idx_up : p.idx_down;
            if (mag) { mag[2 * q] = ma; if (two) mag[2 * q + 1] = mb; }
            if (idx) { idx[2 * q] = ia; if (two) idx[2 * q + 1] = ib; }
        }
    }
#if USC_PAIR_TMEM
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kPairTcols>(*s_tslot);
#endif
}

// dsp() for one hypothesis with a per-stream gather offset (receiver/Src/main.c:183-231).
// The receiver variant's "left" window reads the zero upper half (hazard H1, defined): its maximum
// is 0 at relative index 0, so the right window wins unless its own maximum is negative (never).
template <int NB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) k_dsp2048(demod_params p) {
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
        peak_window<NB>(vr, vi, ws, lane, p.bandwidth2, mr, ir);
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
    size_t cap = (size_t) num_sms * 4 * 4;             // persistent-ish: a few waves of resident CTAs
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int) ctas;
}

#ifndef USC_DUAL_WARPS
#define USC_DUAL_WARPS 12                                  // three warps per scheduler at <= 168 registers (8: two at <= 255)
#endif
template <int NB>
static cudaError_t launch_demod_nb(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st) {
    constexpr int W = USC_DUAL_WARPS;
    size_t ctas = (p.nframes + W - 1) / W;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;            // persistent: one CTA per SM
    static per_device<bool> configured_pd[2];
    bool* configured[2] = {&configured_pd[0].get(), &configured_pd[1].get()};
    const int smem = dual_smem<W>::total;
    if (pcm_format == 1u) {
        if (!*configured[1]) {
            cudaError_t e = cudaFuncSetAttribute(k_demod2048<int32_t, NB, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            *configured[1] = true;
        }
        k_demod2048<int32_t, NB, W><<<(int) ctas, W * 32, smem, st>>>(p);
    } else {
        if (!*configured[0]) {
            cudaError_t e = cudaFuncSetAttribute(k_demod2048<float, NB, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            *configured[0] = true;
        }
        k_demod2048<float, NB, W><<<(int) ctas, W * 32, smem, st>>>(p);
    }
    return cudaGetLastError();
}

template <int NB>
static cudaError_t launch_pair_nb(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st) {
    size_t ctas = ((p.nframes + 1) / 2 + kDualWarps - 1) / kDualWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    static per_device<bool> configured_pd[2];
    bool* configured[2] = {&configured_pd[0].get(), &configured_pd[1].get()};
    if (pcm_format == 1u) {
        if (!*configured[1]) {
            cudaError_t e = cudaFuncSetAttribute(k_demod2048_pair<int32_t, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemTotal);
            if (e != cudaSuccess) return e;
            *configured[1] = true;
        }
        k_demod2048_pair<int32_t, NB><<<(int) ctas, kDualWarps * 32, kPairSmemTotal, st>>>(p);
    } else {
        if (!*configured[0]) {
            cudaError_t e = cudaFuncSetAttribute(k_demod2048_pair<float, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemTotal);
            if (e != cudaSuccess) return e;
            *configured[0] = true;
        }
        k_demod2048_pair<float, NB><<<(int) ctas, kDualWarps * 32, kPairSmemTotal, st>>>(p);
    }
    return cudaGetLastError();
}

// one hypothesis only (p.updown): frames are processed in pairs
cudaError_t launch_demod2048_single(const demod_params& p, uint32_t pcm_format, int num_sms, cudaStream_t st) {
    const uint32_t nb = (p.bandwidth2 + 31) / 32;
    if (nb <= 5) return launch_pair_nb<5>(p, pcm_format, num_sms, st);
    return launch_pair_nb<16>(p, pcm_format, num_sms, st);
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
