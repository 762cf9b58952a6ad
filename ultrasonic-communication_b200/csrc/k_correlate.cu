// k_correlate.cu — K8: overlap-save frame synchroniser (BASELINE north_star: "the frame synchronizer is an overlap-save
// correlation kernel").
//
// compress_chirp() of experiments/chirp_compression_time_domain/Src/chirp.c:78-83 filters ONE n-sample frame with the
// windowed reference chirp in the frequency domain (RFFT -> x H -> IRFFT): a circular result whose lags wrap.  The same
// three CMSIS-shaped steps on 2n-sample windows that advance by n samples give the LINEAR filter output at every lag of
// a stream (overlap-save): block b of a stream is samples [b n, b n + 2n); G = rfft_2n(template zero-padded to 2n);
//   y_b = irfft_2n( rfft_2n(block) x G )          (arm_cmplx_mult_cmplx_f32 on the packed spectrum, quirk kept)
// and y_b[l], l in [n, 2n), equals sum_m g[m] x[b n + l - m] with no wrap-around.  A stream of F frames has F - 1 blocks.
//
// Mapping: one warp walks consecutive blocks of one stream, so every PCM sample crosses HBM once: the upper half of
// block b is the lower half of block b + 1 and stays in shared memory (two 8 KB halves used as a ring); only the n new
// samples arrive per block, by a 1-D TMA bulk copy issued as soon as the old lower half is dead.
// Per block: 4096-point real FFT = 2048-point complex FFT in the canonical plan [2, 32, 32] (even- and odd-bin
// 1024-point transforms ride in the f32x2 halves of the packed register core) -> spectrum parked in natural order in
// shared memory as two planes, real parts | imaginary parts (the dead lower half + the idle exchange tile, 16 KB) ->
// split, x G, merge in place, one lane per (k, 2048 - k) pair, bin k in the .x and bin 2048 - k in the .y half of packed
// operations whose operands the scalar loads form directly (no register moves) -> 2048-point inverse by
// forward-on-swapped-parts -> only the upper half of the outputs is live,
// so half of the last pass is pruned -> signed first-occurrence arg-max over the n valid lags (arm_max_f32) and an
// optional coalesced store of the n filtered samples.
// Arithmetic: the canonical order of DESIGN.md §3 — bit-identical to the oracle's operator chain
// (ref_correlate_os = ref_arm_rfft_fast_f32(4096) . ref_arm_cmplx_mult_cmplx_f32 . ref_arm_rfft_fast_f32(4096, inverse)).
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kOsWarps = 8;
// Tables in tensor memory, one row per lane (usc_tmem.cuh): radix-2 level twiddle W_2048^a of a = lane + 32 q at 2 q |
// inter-pass twiddle W_1024^(lane d) at 64 + 2 d | spectral stage of the pair (k, 2048 - k), k = lane + 32 j, at 128 + 8 j:
// split twiddles and template spectrum as the packed operands of the stage — (re k, re kc), (im k, im kc) of W_4096 and of G,
// so every f32x2 operand is an aligned register pair of the load.  Shared memory keeps the TMEM slot only.
constexpr int kOsTtw0 = 0, kOsTtw = 64, kOsTspec = 128, kOsTcols = 512;
constexpr int kOsTabs = 128;                           // TMEM slot (+ padding to keep the per-warp areas 128-byte aligned)
constexpr int kOsTile = 8448;                          // exchange tile with padded rows (32 x 33 float2)
constexpr int kOsHalfB = 8192 + kOsTile;               // byte offset of half B
constexpr int kOsWarpBytes = kOsHalfB + 8192;          // half A | exchange tile | half B
constexpr int kOsBar = kOsTabs + kOsWarps * kOsWarpBytes;
constexpr int kOsSmem = kOsBar + kOsWarps * 8;

#ifndef USC_OS_SPECU
#define USC_OS_SPECU 4
#endif
constexpr int kOsSpecU = USC_OS_SPECU;                 // spectral pairs per lane and loop iteration (independent chains in flight)
template <int N> struct ldtm_os;
template <> struct ldtm_os<32> { static __device__ __forceinline__ void ld(uint32_t ta, uint32_t (&t)[32]) { ldtm32(ta, t); } };
template <> struct ldtm_os<64> { static __device__ __forceinline__ void ld(uint32_t ta, uint32_t (&t)[64]) { ldtm64(ta, t); } };

struct os_params {
    const void* pcm; uint32_t nstreams, nframes; size_t stream_stride;
    const float2* G;                                   // packed spectrum of the zero-padded template, 2048 float2
    const float2* tw_pass; const float2* tw0; const float2* tw_split;   // tw_split: (cos, sin)(2 pi k / 4096), k < 2048
    uint32_t seg_blocks, nseg;                         // a work item = seg_blocks consecutive blocks of one stream
    float* out; float* max_val; uint32_t* max_idx;
};

// OUT: the n filtered samples of every block are stored as well (otherwise only the peaks: the float4 transposition of the
// outputs and sixteen stores per lane leave the loop)
template <typename PCM, bool OUT>
__global__ void __launch_bounds__(kOsWarps * 32, 1) k_correlate_os(os_params p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = s_raw + kOsTabs + warp * kOsWarpBytes;
    float2* tile = reinterpret_cast<float2*>(wbase + 8192);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + kOsBar) + warp;
    using V2 = typename vec2<PCM>::type;

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<kOsTcols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tq = tmem_quadrant(*s_tslot, warp);
    if (warp < 4) {                                     // warp q fills lane quadrant q; warps q and q + 4 read it
#pragma unroll 1
        for (int q0 = 0; q0 < 32; q0 += 4) {
            float2 a[4], z[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[j] = p.tw0[lane + 32 * (q0 + j)];
                z[j] = p.tw_pass[(q0 + j) * 32 + lane];
            }
            sttm_f2x4(tq + kOsTtw0 + 2 * q0, a[0], a[1], a[2], a[3]);
            sttm_f2x4(tq + kOsTtw + 2 * q0, z[0], z[1], z[2], z[3]);
        }
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
            const int k = lane + 32 * j, kc = (2048 - k) & 2047;
            const float2 wk = p.tw_split[k], wc = p.tw_split[kc], gk = __ldg(p.G + k), gc = __ldg(p.G + kc);
            sttm_f2x4(tq + kOsTspec + 8 * j, make_float2(wk.x, wc.x), make_float2(wk.y, wc.y), make_float2(gk.x, gc.x), make_float2(gk.y, gc.y));
        }
        sttm_wait();
    }
    const float2 ws1024 = p.tw_split[1024], g1024 = __ldg(p.G + 1024);   // bin 1024 pairs with itself (lane 0)
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();

    const uint32_t nblocks = p.nframes - 1;
    const size_t nitems = (size_t) p.nstreams * p.nseg;
    const size_t nwarps = (size_t) gridDim.x * kOsWarps;
    const PCM* pcm = static_cast<const PCM*>(p.pcm);
    uint32_t parity = 0;
    for (size_t item = (size_t) blockIdx.x * kOsWarps + warp; item < nitems; item += nwarps) {
        const uint32_t s = (uint32_t) (item / p.nseg), seg = (uint32_t) (item % p.nseg);
        const uint32_t b0 = seg * p.seg_blocks;
        const uint32_t b1 = min(b0 + p.seg_blocks, nblocks);
        const PCM* stream = pcm + (size_t) s * p.stream_stride;
        if (lane == 0) {                               // first block of the item: both halves
            mbar_expect_tx(bar, 16384u);
            bulk_g2s(wbase, stream + (size_t) b0 * 2048, 8192u, bar);
            bulk_g2s(wbase + kOsHalfB, stream + (size_t) (b0 + 1) * 2048, 8192u, bar);
        }
        for (uint32_t b = b0; b < b1; ++b) {
            const uint32_t odd = (b - b0) & 1u;        // which physical half holds the older n samples
            const V2* lo_half = reinterpret_cast<const V2*>(wbase + (odd ? kOsHalfB : 0));
            const V2* hi_half = reinterpret_cast<const V2*>(wbase + (odd ? 0 : kOsHalfB));
            // parked spectrum, 2048 bins as two planes (real parts | imaginary parts): dead lower half + tile.  Planar, so
            // that the (k, 2048 - k) operands of the packed spectral stage are formed by the loads themselves (no moves)
            float* zre = reinterpret_cast<float*>(wbase + (odd ? 8192 : 0));
            float* zim = zre + 2048;
            mbar_wait(bar, parity);
            parity ^= 1u;
            float2 re[32], im[32];                     // (.x, .y) = (even-bin transform, odd-bin transform)
#pragma unroll
            for (int g = 0; g < 4; ++g) {              // radix-2 level of the [2,32,32] plan on z[m] = x[2m] + j x[2m+1]
                uint32_t t[16];                        // its twiddles, eight per TMEM round trip
                ldtm16(tq + kOsTtw0 + 16 * g, t);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = 8 * g + u, a = lane + 32 * q;
                    const V2 l = lo_half[a], h = hi_half[a];
                    const float lx = pcm_to_float(l.x), ly = pcm_to_float(l.y), hx = pcm_to_float(h.x), hy = pcm_to_float(h.y);
                    const float er = __fadd_rn(lx, hx), ei = __fadd_rn(ly, hy);
                    const float dr = __fsub_rn(lx, hx), di = __fsub_rn(ly, hy);
                    float orr, oii;
                    cmul(dr, di, __uint_as_float(t[2 * u]), __uint_as_float(t[2 * u + 1]), orr, oii);
                    re[q] = make_float2(er, orr);
                    im[q] = make_float2(ei, oii);
                }
            }
            __syncwarp();                              // the lower half has been consumed by every lane
            fft1024_pair_tm<false, true>(re, im, tile, tq + kOsTtw, 1.0f, lane);
            // spectrum to shared memory in natural order: lane d0, element d1 holds Z[2c], Z[2c+1], c = d0 + 32 d1
#pragma unroll
            for (int d1 = 0; d1 < 32; ++d1) {
                reinterpret_cast<float2*>(zre)[lane + 32 * d1] = re[d1];
                reinterpret_cast<float2*>(zim)[lane + 32 * d1] = im[d1];
            }
            __syncwarp();
            // split -> x G -> merge, in place; this lane owns the pairs (k, 2048 - k), k = lane + 32 j.  Bin k rides in
            // the .x half and bin 2048 - k in the .y half of packed operations (each half the scalar operator sequence).
#pragma unroll 1
            for (int j0 = 0; j0 < 32; j0 += kOsSpecU) {
                uint32_t t4[8 * kOsSpecU];             // packed (W re, W im, G re, G im) of kOsSpecU pairs per TMEM round trip
                ldtm_os<8 * kOsSpecU>::ld(tq + kOsTspec + 8 * j0, t4);
#pragma unroll
                for (int u = 0; u < kOsSpecU; ++u) {
                const int j = j0 + u;
                const uint32_t* t = t4 + 8 * u;
                const int k = lane + 32 * j, kc = (2048 - k) & 2047;
                const float2 zr2 = make_float2(zre[k], zre[kc]), zi2 = make_float2(zim[k], zim[kc]);
                const float2 wr2 = make_float2(__uint_as_float(t[0]), __uint_as_float(t[1])), wi2 = make_float2(__uint_as_float(t[2]), __uint_as_float(t[3]));
                const float2 gr2 = make_float2(__uint_as_float(t[4]), __uint_as_float(t[5])), gi2 = make_float2(__uint_as_float(t[6]), __uint_as_float(t[7]));
                float2 xr2, xi2;
                rfft_split2v(zr2, zi2, swap2(zr2), swap2(zi2), wr2, wi2, xr2, xi2);   // (X[k], X[2048 - k])
                if (k == 0) {                                                         // packed (X[0], X[2048]) in the .x half
                    xr2.x = __fadd_rn(zr2.x, zi2.x);
                    xi2.x = __fsub_rn(zr2.x, zi2.x);
                }
                float2 yr2, yi2;
                cmul2v(xr2, xi2, gr2, gi2, yr2, yi2);                                 // arm_cmplx_mult_cmplx_f32 (quirk at k = 0)
                float2 or2, oi2;
                rfft_merge2v(yr2, yi2, swap2(yr2), swap2(yi2), wr2, wi2, or2, oi2);   // (2 Z'[k], 2 Z'[2048 - k])
                if (k == 0) {
                    or2.x = __fadd_rn(yr2.x, yi2.x);
                    oi2.x = __fsub_rn(yr2.x, yi2.x);
                }
                zre[k] = or2.x;
                zim[k] = oi2.x;
                if (k != 0) {
                    zre[kc] = or2.y;
                    zim[kc] = oi2.y;
                }
                }
            }
            if (lane == 0) {                           // bin 1024 pairs with itself
                const float2 zk = make_float2(zre[1024], zim[1024]), wk = ws1024;
                float xr, xi, yr, yi, zr, zi;
                rfft_split(zk.x, zk.y, zk.x, zk.y, wk.x, wk.y, xr, xi);
                cmul(xr, xi, g1024.x, g1024.y, yr, yi);
                rfft_merge(yr, yi, yr, yi, wk.x, wk.y, zr, zi);
                zre[1024] = zr;
                zim[1024] = zi;
            }
            __syncwarp();
            // inverse = forward transform of the swapped parts (radix-2 level again)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t t[16];
                ldtm16(tq + kOsTtw0 + 16 * g, t);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = 8 * g + u, a = lane + 32 * q;
                    const float2 l = make_float2(zre[a], zim[a]), h = make_float2(zre[a + 1024], zim[a + 1024]);
                    const float er = __fadd_rn(l.y, h.y), ei = __fadd_rn(l.x, h.x);
                    const float dr = __fsub_rn(l.y, h.y), di = __fsub_rn(l.x, h.x);
                    float orr, oii;
                    cmul(dr, di, __uint_as_float(t[2 * u]), __uint_as_float(t[2 * u + 1]), orr, oii);
                    re[q] = make_float2(er, orr);
                    im[q] = make_float2(ei, oii);
                }
            }
            __syncwarp();                              // the parked spectrum is dead: its half can take the next n samples
            if (lane == 0 && b + 1 < b1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of the parked spectrum before the bulk copy
                mbar_expect_tx(bar, 8192u);
                bulk_g2s(wbase + (odd ? kOsHalfB : 0), stream + (size_t) (b + 2) * 2048, 8192u, bar);
            }
            fft1024_pair_tm<false, true>(re, im, tile, tq + kOsTtw, 1.0f, lane);
            // y[2m] = z[m].im / 4096, y[2m+1] = z[m].re / 4096; valid lags l = 2m (+1) in [2048, 4096): m = 2c (+1) with
            // c = lane + 32 d1 >= 512, i.e. d1 >= 16 (the other outputs of the last pass are never computed)
            const float sc = 1.0f / 4096.0f;
            float best = -INFINITY;
            uint32_t best_idx = 0xffffffffu;
            float4* o4 = OUT ? reinterpret_cast<float4*>(p.out + ((size_t) s * nblocks + b) * 2048) : nullptr;
#pragma unroll
            for (int d1 = 16; d1 < 32; ++d1) {
                const int c = lane + 32 * d1;
                const float2 a0 = __fmul2_rn(im[d1], bc2(sc)), a1 = __fmul2_rn(re[d1], bc2(sc));
                const float4 v = make_float4(a0.x, a1.x, a0.y, a1.y);                 // lags 4c-2048 .. 4c-2045
                const uint32_t l0 = 4u * (uint32_t) c - 2048u;
                if (v.x > best) { best = v.x; best_idx = l0; }
                if (v.y > best) { best = v.y; best_idx = l0 + 1; }
                if (v.z > best) { best = v.z; best_idx = l0 + 2; }
                if (v.w > best) { best = v.w; best_idx = l0 + 3; }
                if (OUT) o4[c - 512] = v;
            }
            if (best_idx == 0xffffffffu) best_idx = 4u * (uint32_t) (lane + 512) - 2048u;   // all NaN / -inf: first own lag
            warp_argmax(best, best_idx);
            if (lane == 0) {
                const size_t o = (size_t) s * nblocks + b;
                if (p.max_val) p.max_val[o] = best;
                if (p.max_idx) p.max_idx[o] = best_idx;
            }
            __syncwarp();
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kOsTcols>(*s_tslot);
}

cudaError_t launch_correlate_os(const void* pcm, uint32_t pcm_format, uint32_t nstreams, uint32_t nframes, size_t stream_stride,
                                const float2* G, const float2* tw_pass, const float2* tw0, const float2* tw_split, float* out,
                                float* max_val, uint32_t* max_idx, int num_sms, cudaStream_t st) {
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_correlate_os<int32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOsSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_correlate_os<int32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOsSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_correlate_os<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOsSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_correlate_os<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOsSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    os_params p;
    p.pcm = pcm; p.nstreams = nstreams; p.nframes = nframes; p.stream_stride = stream_stride;
    p.G = G; p.tw_pass = tw_pass; p.tw0 = tw0; p.tw_split = tw_split;
    p.out = out; p.max_val = max_val; p.max_idx = max_idx;
    // segments: keep every sample's single HBM crossing (a segment re-reads one frame at its start) but split long
    // streams when there are too few of them to fill the machine
    const uint32_t nblocks = nframes - 1;
    const size_t resident = (size_t) num_sms * kOsWarps;
    uint32_t seg = nblocks;
    if ((size_t) nstreams < 4 * resident) {
        const size_t want = (4 * resident + nstreams - 1) / nstreams;                 // segments per stream
        seg = (uint32_t) ((nblocks + want - 1) / want);
        if (seg < 8) seg = nblocks < 8 ? nblocks : 8;
    }
    p.seg_blocks = seg;
    p.nseg = (nblocks + seg - 1) / seg;
    size_t ctas = ((size_t) nstreams * p.nseg + kOsWarps - 1) / kOsWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    if (pcm_format == 1u) {
        if (out) k_correlate_os<int32_t, true><<<(int) ctas, kOsWarps * 32, kOsSmem, st>>>(p);
        else k_correlate_os<int32_t, false><<<(int) ctas, kOsWarps * 32, kOsSmem, st>>>(p);
    } else {
        if (out) k_correlate_os<float, true><<<(int) ctas, kOsWarps * 32, kOsSmem, st>>>(p);
        else k_correlate_os<float, false><<<(int) ctas, kOsWarps * 32, kOsSmem, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace usc
