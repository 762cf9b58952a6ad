// usc_warpfft.cuh — warp-per-frame building blocks shared by the fused kernels (K1, K4, K7):
// the 1024-point complex FFT of one warp ([32,32] plan, exchange through a padded shared tile) and
// the exact windowed peak search on the packed real-FFT bins.
#pragma once
#include "usc_arith.cuh"
#include "usc_tmem.cuh"

namespace usc {

constexpr int kTileStride = 33;                       // float2 units, 32x33 padded tile
constexpr int kTileFloat2 = 32 * kTileStride;

__device__ __forceinline__ float pcm_to_float(int32_t v) { return __int2float_rn(v); }
__device__ __forceinline__ float pcm_to_float(float v) { return v; }

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<int32_t> { using type = int2; };

// ---- mbarrier + 1-D TMA bulk copy (global -> shared) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// pass 1 + twiddle + exchange + pass 2 for one transform; lane a enters with z[a + 32 b] in
// register b and leaves (as lane d0) with Z[d0 + 32*d1] in register d1.
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

// Packed twin: two transforms side by side in the halves of f32x2 registers (K1: up/down hypothesis
// of one frame; K3: the even/odd halves of a 2048-point complex FFT).  tile: XOR-swizzled 32x32
// float4 (16 KB), conflict-free for the 128-bit stores and loads.
__device__ __forceinline__ void fft1024_warp2(float2 (&re)[32], float2 (&im)[32], float4* tile,
                                              const float2* __restrict__ tw_pass, int lane) {
    fft_base2<32>(re, im);
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        float4 v = make_float4(re[d].x, re[d].y, im[d].x, im[d].y);
        if (d != 0) {
            float2 w = tw_pass[d * 32 + lane];
            cmul(re[d].x, im[d].x, w.x, w.y, v.x, v.z);
            cmul(re[d].y, im[d].y, w.x, w.y, v.y, v.w);
        }
        tile[d * 32 + (lane ^ d)] = v;
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) {
        float4 v = tile[lane * 32 + (a ^ lane)];
        re[a] = make_float2(v.x, v.y);
        im[a] = make_float2(v.z, v.w);
    }
    __syncwarp();
    fft_base2<32>(re, im);
}

// pass 1, inter-pass twiddle, two-round exchange (real parts, imaginary parts), pass 2 — on pairs.
// PROD: the inputs are packed products (window multiplies): the first butterfly stage runs as FMAs by 1.0 (s_tw row 0
// holds W^0 = 1.0f), see bfly2_one_fma in usc_arith.cuh; same bits.
template <bool PROD = false, bool PAD = false>
__device__ __forceinline__ void fft1024_pair(float2 (&re)[32], float2 (&im)[32], float2* tile /* 32x32 float2 XOR-swizzled (8 KB), or 32x33 padded */,
                                             const float2* __restrict__ s_tw, int lane) {
    if (PROD) fft_base2_prod<32>(re, im, s_tw[lane].x);
    else fft_base2<32>(re, im);
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        const float2 w = s_tw[d * 32 + lane];
        float2 tr, ti;
        cmul2(re[d], im[d], w.x, w.y, tr, ti);
        re[d] = tr;
        im[d] = ti;
    }
#pragma unroll
    for (int d = 0; d < 32; ++d) tile[PAD ? d * kTileStride + lane : d * 32 + (lane ^ d)] = re[d];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) re[a] = tile[PAD ? lane * kTileStride + a : lane * 32 + (a ^ lane)];
    __syncwarp();
#pragma unroll
    for (int d = 0; d < 32; ++d) tile[PAD ? d * kTileStride + lane : d * 32 + (lane ^ d)] = im[d];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) im[a] = tile[PAD ? lane * kTileStride + a : lane * 32 + (a ^ lane)];
    __syncwarp();
    fft_base2<32>(re, im);
}

// fft1024_pair with the inter-pass twiddles W_1024^(a d) read from the lane's TMEM row (columns t_tw + 2 d, d < 32; see
// usc_tmem.cuh) instead of shared memory; `one` = 1.0f from a table (opaque to the compiler) for the PROD form.
// PAD: the tile has padded rows (32 x 33 float2, 8448 bytes) instead of the XOR swizzle: every store and load of a round is
// one base register plus an immediate offset, which is what lets a kernel hold three warps per scheduler at 168 registers.
template <bool PROD = false, bool PAD = false>
__device__ __forceinline__ void fft1024_pair_tm(float2 (&re)[32], float2 (&im)[32], float2* tile /* 32x32 float2, XOR-swizzled, 8 KB */,
                                                uint32_t t_tw, float one, int lane) {
    if (PROD) fft_base2_prod<32>(re, im, one);
    else fft_base2<32>(re, im);
#pragma unroll
    for (int g = 0; g < 4; ++g) {                                     // eight twiddles per TMEM round trip
        uint32_t t[16];
        ldtm16(t_tw + 16 * g, t);
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
#pragma unroll
    for (int d = 0; d < 32; ++d) tile[PAD ? d * kTileStride + lane : d * 32 + (lane ^ d)] = re[d];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) re[a] = tile[PAD ? lane * kTileStride + a : lane * 32 + (a ^ lane)];
    __syncwarp();
#pragma unroll
    for (int d = 0; d < 32; ++d) tile[PAD ? d * kTileStride + lane : d * 32 + (lane ^ d)] = im[d];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 32; ++a) im[a] = tile[PAD ? lane * kTileStride + a : lane * 32 + (a ^ lane)];
    __syncwarp();
    fft_base2<32>(re, im);
}

// table set-up helper: eight consecutive columns of the lane's row from four float2 values
__device__ __forceinline__ void sttm_f2x4(uint32_t taddr, float2 a, float2 b, float2 c, float2 d) {
    const uint32_t v[8] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(b.x), __float_as_uint(b.y),
                           __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(d.x), __float_as_uint(d.y)};
    sttm8(taddr, v);
}

// Exact arm_max_f32 over sqrt(p_k) with ONE square root and no rare path.  With r = sqrt_rn(pmax) and r- its predecessor,
// a candidate p <= pmax rounds to the same root exactly when sqrt(p) lies above the midpoint m = (r- + r) / 2, i.e. when
// p > m^2 (sqrt(p) = m is impossible: m has 25 significant bits, so m^2 is not a float).  m^2 = (r- + r)^2 / 4 is exact
// in double (a 25-bit sum, a 50-bit square), and the smallest float above it is its conversion rounded up: the first
// index attaining the maximum root is the first k with p_k >= that threshold.  pmax = 0: every root is 0, threshold 0.
__device__ __forceinline__ float same_root_threshold(float pmax, float& root) {
    root = __fsqrt_rn(pmax);
    const uint32_t rb = __float_as_uint(root);
    const double s = (double) root + (double) __uint_as_float(rb ? rb - 1u : 0u);
    return rb ? __double2float_ru(0.25 * (s * s)) : 0.0f;
}

// Each lane passes NC candidates (squared magnitude p >= 0, index k, validity); the warp gets the largest root and the
// first index attaining it (arm_max_f32 over sqrt(p_k)).
template <int NC>
__device__ __forceinline__ void argmax_exact(const float (&p)[NC], const uint32_t (&k)[NC], const bool (&ok)[NC],
                                             float& best, uint32_t& best_idx) {
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) q = ok[c] ? fmaxf(q, p[c]) : q;
    const float pmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(q)));   // p >= 0: bit patterns order like values
    const float thr = same_root_threshold(pmax, best);
    uint32_t kb = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < NC; ++c) kb = (ok[c] && p[c] >= thr) ? min(kb, k[c]) : kb;
    best_idx = __reduce_min_sync(0xffffffffu, kb);
}

// Real-FFT split of this lane's bins k = lane + 32*d1 (d1 < NB), magnitude and arg-max over
// [0, bw2): the receiver's "right" window (receiver/Src/main.c:208).  zr/zi: lane d0 holds
// Z[d0 + 32*d1] in element d1; only elements [0, NB) and [32-NB, 32) are read.
template <int NB>
__device__ __forceinline__ void peak_tail(const float (&pw)[NB], int lane, uint32_t bw2, float& best, uint32_t& best_idx) {
    uint32_t k[NB];
    bool ok[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        k[d1] = (uint32_t) lane + 32u * d1;
        ok[d1] = k[d1] < bw2;
    }
    argmax_exact<NB>(pw, k, ok, best, best_idx);
}

template <int NB>
__device__ __forceinline__ void peak_window(const float (&zr)[32], const float (&zi)[32],
                                            const float2 (&ws)[NB], int lane, uint32_t bw2,
                                            float& best, uint32_t& best_idx) {
    const int src = (32 - lane) & 31;
    float pw[NB];                                      // squared magnitudes of this lane's bins
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        // partner Z[1024 - k] lives in lane (32 - lane) & 31, register 31 - d1 (32 - d1 on lane 0)
        float sr = lane == 0 ? zr[(32 - d1) & 31] : zr[31 - d1];
        float si = lane == 0 ? zi[(32 - d1) & 31] : zi[31 - d1];
        float zcr = __shfl_sync(0xffffffffu, sr, src);
        float zci = __shfl_sync(0xffffffffu, si, src);
        float xr, xi;
        rfft_split(zr[d1], zi[d1], zcr, zci, ws[d1].x, ws[d1].y, xr, xi);
        if (d1 == 0) {                                 // packed bin 0 = (X[0], X[N/2]) on lane 0
            float dr = __fadd_rn(zr[0], zi[0]), di = __fsub_rn(zr[0], zi[0]);
            xr = lane == 0 ? dr : xr;
            xi = lane == 0 ? di : xi;
        }
        pw[d1] = __fmaf_rn(xr, xr, __fmul_rn(xi, xi));
    }
    peak_tail<NB>(pw, lane, bw2, best, best_idx);
}

// argmax_exact for two candidate sets that share their indices (the two hypotheses / frames of a packed pass), side by
// side so that the warp-wide reductions of the two searches overlap.
template <int NC>
__device__ __forceinline__ void argmax_exact2(const float (&pa)[NC], const float (&pb)[NC], const uint32_t (&k)[NC], const bool (&ok)[NC],
                                              float& bestA, uint32_t& idxA, float& bestB, uint32_t& idxB) {
    float qa = 0.0f, qb = 0.0f;                        // p >= 0: bit patterns order like values
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        qa = ok[c] ? fmaxf(qa, pa[c]) : qa;
        qb = ok[c] ? fmaxf(qb, pb[c]) : qb;
    }
    const float maxA = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(qa)));
    const float maxB = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(qb)));
    const float thrA = same_root_threshold(maxA, bestA), thrB = same_root_threshold(maxB, bestB);
    uint32_t ka = 0xffffffffu, kb = 0xffffffffu;
#pragma unroll
    for (int c = NC - 1; c >= 0; --c) {                // k ascends with c: descending, the lowest qualifying k of the lane survives
        ka = (ok[c] && pa[c] >= thrA) ? k[c] : ka;
        kb = (ok[c] && pb[c] >= thrB) ? k[c] : kb;
    }
    idxA = __reduce_min_sync(0xffffffffu, ka);
    idxB = __reduce_min_sync(0xffffffffu, kb);
}

// peak_tail for the two candidate sets of a packed pass
template <int NB>
__device__ __forceinline__ void peak_tail_pair(const float (&pa)[NB], const float (&pb)[NB], int lane, uint32_t bw2,
                                               float& bestA, uint32_t& idxA, float& bestB, uint32_t& idxB) {
    uint32_t k[NB];
    bool ok[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        k[d1] = (uint32_t) lane + 32u * d1;
        ok[d1] = k[d1] < bw2;
    }
    argmax_exact2<NB>(pa, pb, k, ok, bestA, idxA, bestB, idxB);
}

// squared magnitudes of the real-FFT bins k = lane + 32 d1, d1 < NB, of both halves (no peak search)
template <int NB>
__device__ __forceinline__ void mag2_window_pair(const float2 (&zr)[32], const float2 (&zi)[32], const float2 (&ws)[NB],
                                                 int lane, float (&pa)[NB], float (&pb)[NB]) {
    const int src = (32 - lane) & 31;
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        const float2 sr = lane == 0 ? zr[(32 - d1) & 31] : zr[31 - d1];
        const float2 si = lane == 0 ? zi[(32 - d1) & 31] : zi[31 - d1];
        const float2 zcr = make_float2(__shfl_sync(0xffffffffu, sr.x, src), __shfl_sync(0xffffffffu, sr.y, src));
        const float2 zci = make_float2(__shfl_sync(0xffffffffu, si.x, src), __shfl_sync(0xffffffffu, si.y, src));
        float2 xr, xi;
        rfft_split2(zr[d1], zi[d1], zcr, zci, ws[d1].x, ws[d1].y, xr, xi);
        if (d1 == 0) {                                 // packed bin 0 = (X[0], X[N/2]) on lane 0
            const float2 dr = __fadd2_rn(zr[0], zi[0]), di = __fadd2_rn(zr[0], neg2(zi[0]));
            xr = lane == 0 ? dr : xr;
            xi = lane == 0 ? di : xi;
        }
        const float2 p = __ffma2_rn(xr, xr, __fmul2_rn(xi, xi));
        pa[d1] = p.x;
        pb[d1] = p.y;
    }
}

// the same with the split twiddles fetched from shared memory when they are needed (ws_s[lane + 32 d1]) instead of
// living in NB register pairs across the whole frame loop
template <int NB>
__device__ __forceinline__ void peak_window_pair_s(const float2 (&zr)[32], const float2 (&zi)[32], const float2* ws_s,
                                                   int lane, uint32_t bw2, float& bestA, uint32_t& idxA, float& bestB,
                                                   uint32_t& idxB) {
    const int src = (32 - lane) & 31;
    float pa[NB], pb[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        const float2 sr = lane == 0 ? zr[(32 - d1) & 31] : zr[31 - d1];
        const float2 si = lane == 0 ? zi[(32 - d1) & 31] : zi[31 - d1];
        const float2 zcr = make_float2(__shfl_sync(0xffffffffu, sr.x, src), __shfl_sync(0xffffffffu, sr.y, src));
        const float2 zci = make_float2(__shfl_sync(0xffffffffu, si.x, src), __shfl_sync(0xffffffffu, si.y, src));
        const float2 w = ws_s[lane + 32 * d1];
        float2 xr, xi;
        rfft_split2(zr[d1], zi[d1], zcr, zci, w.x, w.y, xr, xi);
        if (d1 == 0) {                                 // packed bin 0 = (X[0], X[N/2]) on lane 0
            const float2 dr = __fadd2_rn(zr[0], zi[0]), di = __fadd2_rn(zr[0], neg2(zi[0]));
            xr = lane == 0 ? dr : xr;
            xi = lane == 0 ? di : xi;
        }
        const float2 p = __ffma2_rn(xr, xr, __fmul2_rn(xi, xi));
        pa[d1] = p.x;
        pb[d1] = p.y;
    }
    peak_tail_pair<NB>(pa, pb, lane, bw2, bestA, idxA, bestB, idxB);
}

template <int NB>
__device__ __forceinline__ void peak_window_pair(const float2 (&zr)[32], const float2 (&zi)[32], const float2 (&ws)[NB],
                                                 int lane, uint32_t bw2, float& bestA, uint32_t& idxA, float& bestB,
                                                 uint32_t& idxB) {
    const int src = (32 - lane) & 31;
    float pa[NB], pb[NB];
#pragma unroll
    for (int d1 = 0; d1 < NB; ++d1) {
        const float2 sr = lane == 0 ? zr[(32 - d1) & 31] : zr[31 - d1];
        const float2 si = lane == 0 ? zi[(32 - d1) & 31] : zi[31 - d1];
        const float2 zcr = make_float2(__shfl_sync(0xffffffffu, sr.x, src), __shfl_sync(0xffffffffu, sr.y, src));
        const float2 zci = make_float2(__shfl_sync(0xffffffffu, si.x, src), __shfl_sync(0xffffffffu, si.y, src));
        float2 xr, xi;
        rfft_split2(zr[d1], zi[d1], zcr, zci, ws[d1].x, ws[d1].y, xr, xi);
        if (d1 == 0) {                                 // packed bin 0 = (X[0], X[N/2]) on lane 0
            const float2 dr = __fadd2_rn(zr[0], zi[0]), di = __fadd2_rn(zr[0], neg2(zi[0]));
            xr = lane == 0 ? dr : xr;
            xi = lane == 0 ? di : xi;
        }
        const float2 p = __ffma2_rn(xr, xr, __fmul2_rn(xi, xi));
        pa[d1] = p.x;
        pb[d1] = p.y;
    }
    peak_tail_pair<NB>(pa, pb, lane, bw2, bestA, idxA, bestB, idxB);
}

}  // namespace usc
