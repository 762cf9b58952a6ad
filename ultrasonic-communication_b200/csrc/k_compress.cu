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
// Persistent, one CTA of 8 warps per SM; each warp carries TWO FRAMES in the halves of its f32x2
// registers through the packed 32x32 core (same layout as K1): after the forward 1024-point FFT
// lane d0 holds Z[d0 + 32 d1]; split -> X[k] -> P = X*H -> merge -> 2Z'[k] happen bin by bin with two
// shuffle rounds for the partners (Z[1024-k], P[1024-k]); Z' is already in the layout the next pass
// expects, so the inverse transform needs no re-shuffle.  The next pair's PCM (16 KB) arrives by one
// TMA bulk copy while the current pair computes.  Each frame crosses HBM once in (8 KB) and, unless
// the caller wants the compressed frames, 8 bytes go out.
#include "usc_kernels.cuh"
#include "usc_launch.h"
#include "usc_warpfft.cuh"

namespace usc {

constexpr int kCWarps = 8;
// Tables in tensor memory, one row per lane (usc_tmem.cuh): window pair of m = lane + 32 b at 2 b | inter-pass twiddle
// W_1024^(lane d) at 64 + 2 d | spectral stage, step j: (ws[ka], H[ka], ws[kb], H[kb]) with ka = lane + 32 j,
// kb = lane + 32 (31 - j), at 128 + 8 j | lane 0's column: (ws[32 lane], H[32 lane]) at 256.
constexpr int kCTwin = 0, kCTtw = 64, kCTspec = 128, kCTcol0 = 256, kCTcols = 512;
constexpr int kCSmemTabs = 128;                       // TMEM slot (+ padding: the per-warp areas stay 128-byte aligned)
constexpr int kCWarpBytes = 8448 + 16384;             // 2-frame PCM stage + float2 tile with padded rows (32 x 33)
constexpr int kCSmemBar = kCSmemTabs + kCWarps * kCWarpBytes;
constexpr int kCSmemTotal = kCSmemBar + kCWarps * 8;

__device__ __forceinline__ float2 shfl2(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// split -> x H -> merge for one bin k, given the partner spectra; returns swap(2Z'[k]) ready for the
// forward-on-swapped inverse.  `dc` marks the packed bin 0 = (X[0], X[N/2]).
__device__ __forceinline__ void split_mul(float2 zkr, float2 zki, float2 zcr, float2 zci, float2 w, float2 h, bool dc,
                                          float2& pr, float2& pi) {
    float2 xr, xi;
    rfft_split2(zkr, zki, zcr, zci, w.x, w.y, xr, xi);
    const float2 dr = __fadd2_rn(zkr, zki), di = __fadd2_rn(zkr, neg2(zki));
    xr = dc ? dr : xr;
    xi = dc ? di : xi;
    cmul2(xr, xi, h.x, h.y, pr, pi);
}
__device__ __forceinline__ void merge_swap(float2 pkr, float2 pki, float2 pcr, float2 pci, float2 w, bool dc,
                                           float2& out_re, float2& out_im) {
    float2 zr, zi;
    rfft_merge2(pkr, pki, pcr, pci, w.x, w.y, zr, zi);
    const float2 dr = __fadd2_rn(pkr, pki), di = __fadd2_rn(pkr, neg2(pki));
    zr = dc ? dr : zr;
    zi = dc ? di : zi;
    out_re = zi;          // swap(re, im): the inverse transform is the forward one on swapped parts
    out_im = zr;
}

// The whole spectral stage IN PLACE on the registers (lane d0 holds Z[d0 + 32 d1] in element d1).
// Bin k = d0 + 32 d1 pairs with 1024 - k = (32 - d0) + 32 (31 - d1): for d0 != 0 the partner lives in
// lane 32 - d0 at element 31 - d1, so processing elements (j, 31 - j) together keeps both ends of
// every pair inside one step and lets the results overwrite their inputs.  Lane 0's own column
// (k = 32 d1, partner 32 (32 - d1), same lane) does not fit that order; it is spread over the 32 lanes
// through the (idle) exchange tile, processed one bin per lane, and gathered back.
__device__ __forceinline__ void spectral_in_place(float2 (&re)[32], float2 (&im)[32], float4* tile4, uint32_t tq, int lane) {
    // ---- column of lane 0 ----
    if (lane == 0) {
#pragma unroll
        for (int d1 = 0; d1 < 32; ++d1) tile4[d1] = make_float4(re[d1].x, re[d1].y, im[d1].x, im[d1].y);
    }
    __syncwarp();
    {
        const int pl = (32 - lane) & 31;
        const float4 zk = tile4[lane], zc = tile4[pl];
        uint32_t t0[8];
        ldtm8(tq + kCTcol0, t0);
        const float2 w = make_float2(__uint_as_float(t0[0]), __uint_as_float(t0[1])), h = make_float2(__uint_as_float(t0[2]), __uint_as_float(t0[3]));
        float2 pr, pi, o_re, o_im;
        split_mul(make_float2(zk.x, zk.y), make_float2(zk.z, zk.w), make_float2(zc.x, zc.y), make_float2(zc.z, zc.w), w, h,
                  lane == 0, pr, pi);
        const float2 pcr = shfl2(pr, pl), pci = shfl2(pi, pl);
        merge_swap(pr, pi, pcr, pci, w, lane == 0, o_re, o_im);
        __syncwarp();
        tile4[lane] = make_float4(o_re.x, o_re.y, o_im.x, o_im.y);
    }
    __syncwarp();                                      // results of column 0 wait in the tile until the end
    // ---- all other lanes: elements (j, 31 - j) per step ----
    const int src = (32 - lane) & 31;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int ja = j, jb = 31 - j;
        uint32_t t[8];                                                          // table values of this step
        ldtm8(tq + kCTspec + 8 * j, t);
        const float2 zcar = shfl2(re[jb], src), zcai = shfl2(im[jb], src);     // partner of my element ja
        const float2 zcbr = shfl2(re[ja], src), zcbi = shfl2(im[ja], src);     // partner of my element jb
        const float2 wa = make_float2(__uint_as_float(t[0]), __uint_as_float(t[1])), ha = make_float2(__uint_as_float(t[2]), __uint_as_float(t[3]));
        const float2 wb = make_float2(__uint_as_float(t[4]), __uint_as_float(t[5])), hb = make_float2(__uint_as_float(t[6]), __uint_as_float(t[7]));
        float2 par, pai, pbr, pbi;
        split_mul(re[ja], im[ja], zcar, zcai, wa, ha, false, par, pai);
        split_mul(re[jb], im[jb], zcbr, zcbi, wb, hb, false, pbr, pbi);
        const float2 pcar = shfl2(pbr, src), pcai = shfl2(pbi, src);           // P[1024 - ka] = partner's P_b
        const float2 pcbr = shfl2(par, src), pcbi = shfl2(pai, src);
        float2 ar, ai, br, bi;
        merge_swap(par, pai, pcar, pcai, wa, false, ar, ai);
        merge_swap(pbr, pbi, pcbr, pcbi, wb, false, br, bi);
        re[ja] = ar; im[ja] = ai;
        re[jb] = br; im[jb] = bi;
    }
    if (lane == 0) {
#pragma unroll
        for (int d1 = 0; d1 < 32; ++d1) {
            const float4 c = tile4[d1];
            re[d1] = make_float2(c.x, c.y);
            im[d1] = make_float2(c.z, c.w);
        }
    }
    __syncwarp();
}

struct compress_params {
    const void* pcm; size_t nframes;
    const float2* window; const float2* H; const float2* tw_pass; const float2* tw_split;
    float* out_frames; float* max_val; uint32_t* max_idx;
};

template <typename PCM>
__global__ void __launch_bounds__(kCWarps * 32, 1) k_compress2048(compress_params p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    uint32_t* s_tslot = reinterpret_cast<uint32_t*>(s_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = s_raw + kCSmemTabs + warp * kCWarpBytes;
    using V2 = typename vec2<PCM>::type;
    V2* xstage = reinterpret_cast<V2*>(wbase);
    float2* tile = reinterpret_cast<float2*>(wbase + 16384);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_raw + kCSmemBar) + warp;

    const size_t npairs = (p.nframes + 1) / 2;
    const size_t nwarps = (size_t) gridDim.x * kCWarps;
    size_t q = (size_t) blockIdx.x * kCWarps + warp;
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
    if (warp == 0) tmem_alloc<kCTcols>(s_tslot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tq = tmem_quadrant(*s_tslot, warp);
    if (warp < 4) {                                     // warp q fills lane quadrant q; warps q and q + 4 read it
#pragma unroll 1
        for (int b0 = 0; b0 < 32; b0 += 4) {
            float2 w[4], z[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                w[j] = p.window[lane + 32 * (b0 + j)];
                z[j] = p.tw_pass[(b0 + j) * 32 + lane];
            }
            sttm_f2x4(tq + kCTwin + 2 * b0, w[0], w[1], w[2], w[3]);
            sttm_f2x4(tq + kCTtw + 2 * b0, z[0], z[1], z[2], z[3]);
        }
#pragma unroll 1
        for (int j = 0; j < 16; ++j) {
            const int ka = lane + 32 * j, kb = lane + 32 * (31 - j);
            sttm_f2x4(tq + kCTspec + 8 * j, p.tw_split[ka], p.H[ka], p.tw_split[kb], p.H[kb]);
        }
        sttm_f2x4(tq + kCTcol0, p.tw_split[32 * lane], p.H[32 * lane], make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f));
        sttm_wait();
    }
    const float one = p.tw_pass[lane].x;                // W^0 = 1.0f read from a table: opaque to the compiler
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();

    uint32_t parity = 0;
    for (; q < npairs; q += nwarps) {
        const bool two = 2 * q + 1 < p.nframes;
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 re[32], im[32];                         // (.x, .y) = (frame 2q, frame 2q+1)
#pragma unroll
        for (int g = 0; g < 4; ++g) {                   // window values of eight rows per TMEM round trip
            uint32_t t[16];
            ldtm16(tq + kCTwin + 16 * g, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = 8 * g + j, m = lane + 32 * b;
                const V2 ra = xstage[m];
                const V2 rb = two ? xstage[1024 + m] : ra;
                // packed window multiply; the first butterfly stage takes the products as FMAs by 1.0 (usc_arith.cuh)
                re[b] = __fmul2_rn(make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x)), bc2(__uint_as_float(t[2 * j])));
                im[b] = __fmul2_rn(make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y)), bc2(__uint_as_float(t[2 * j + 1])));
            }
        }
        __syncwarp();
        if (lane == 0 && q + nwarps < npairs) {
            mbar_expect_tx(bar, pair_bytes(q + nwarps));
            bulk_g2s(xstage, pcm + (q + nwarps) * 4096, pair_bytes(q + nwarps), bar);
        }
        fft1024_pair_tm<true, true>(re, im, tile, tq + kCTtw, one, lane);
        spectral_in_place(re, im, reinterpret_cast<float4*>(tile), tq, lane);
        fft1024_pair_tm<false, true>(re, im, tile, tq + kCTtw, one, lane);
        // swap back and scale by 1/N: a[2m] = z.im/N, a[2m+1] = z.re/N
        const float sc = 1.0f / 2048.0f;
        float ba = -INFINITY, bb = -INFINITY;
        uint32_t ia = 0xffffffffu, ib = 0xffffffffu;
#pragma unroll
        for (int d1 = 0; d1 < 32; ++d1) {
            const uint32_t m = (uint32_t) lane + 32u * d1;
            const float2 a0 = __fmul2_rn(im[d1], bc2(sc)), a1 = __fmul2_rn(re[d1], bc2(sc));
            if (p.out_frames) {
                reinterpret_cast<float2*>(p.out_frames + (2 * q) * 2048)[m] = make_float2(a0.x, a1.x);
                if (two) reinterpret_cast<float2*>(p.out_frames + (2 * q + 1) * 2048)[m] = make_float2(a0.y, a1.y);
            }
            if (ia == 0xffffffffu || ba < a0.x) { ba = a0.x; ia = 2 * m; }
            if (ba < a1.x) { ba = a1.x; ia = 2 * m + 1; }
            if (ib == 0xffffffffu || bb < a0.y) { bb = a0.y; ib = 2 * m; }
            if (bb < a1.y) { bb = a1.y; ib = 2 * m + 1; }
        }
        warp_argmax(ba, ia);
        warp_argmax(bb, ib);
        if (lane == 0) {
            if (p.max_val) { p.max_val[2 * q] = ba; if (two) p.max_val[2 * q + 1] = bb; }
            if (p.max_idx) { p.max_idx[2 * q] = ia; if (two) p.max_idx[2 * q + 1] = ib; }
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kCTcols>(*s_tslot);
}

cudaError_t launch_compress2048(const void* pcm, uint32_t pcm_format, size_t nframes, const float2* window,
                                const float2* H, const float2* tw_pass, const float2* tw_split, float* out_frames,
                                float* max_val, uint32_t* max_idx, int num_sms, cudaStream_t st) {
    compress_params p{pcm, nframes, window, H, tw_pass, tw_split, out_frames, max_val, max_idx};
    size_t ctas = ((nframes + 1) / 2 + kCWarps - 1) / kCWarps;
    if (ctas > (size_t) num_sms) ctas = (size_t) num_sms;
    static per_device<bool> configured_pd;
    bool& configured = configured_pd.get();
    if (!configured) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(k_compress2048<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCSmemTotal))) return e;
        if ((e = cudaFuncSetAttribute(k_compress2048<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCSmemTotal))) return e;
        configured = true;
    }
    if (pcm_format == 1u) k_compress2048<int32_t><<<(int) ctas, kCWarps * 32, kCSmemTotal, st>>>(p);
    else k_compress2048<float><<<(int) ctas, kCWarps * 32, kCSmemTotal, st>>>(p);
    return cudaGetLastError();
}

// tail of pipeline(): packed spectrum (n floats) -> n/2 magnitudes + n/2 zeros, in place, one CTA per vector
// (receiver/Src/main.c:178 with hazard H1 defined).  Ascending chunks, read -> barrier -> write: magnitude i lands
// on a float whose own consumer (magnitude i/2) has already been computed, so no staging buffer that grows with n.
__global__ void k_pipeline_tail(float* data, uint32_t n, uint32_t batch, int zero_upper) {
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        float* p = data + (size_t) v * n;
        for (uint32_t base = 0; base < n / 2; base += blockDim.x) {
            const uint32_t i = base + threadIdx.x;
            float m = 0.0f;
            if (i < n / 2) m = cmag(p[2 * i], p[2 * i + 1]);
            __syncthreads();
            if (i < n / 2) p[i] = m;
            __syncthreads();
        }
        if (zero_upper)                                     // otherwise the packed-spectrum floats stay (in-place semantics)
            for (uint32_t i = n / 2 + threadIdx.x; i < n; i += blockDim.x) p[i] = 0.0f;
    }
}

cudaError_t launch_pipeline_tail(float* data, uint32_t n, uint32_t batch, int zero_upper, cudaStream_t st) {
    int grid = batch < 148u * 16u ? (int) batch : 148 * 16;
    k_pipeline_tail<<<grid, 256, 0, st>>>(data, n, batch, zero_upper);
    return cudaGetLastError();
}

}  // namespace usc
