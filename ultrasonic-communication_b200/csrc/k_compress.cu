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
constexpr int kCSmemTabs = 4 * 8192;                  // pass twiddles | window | H | split twiddles
constexpr int kCWarpBytes = 8192 + 16384;             // XOR-swizzled float2 tile + 2-frame PCM stage
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
__device__ __forceinline__ void spectral_in_place(float2 (&re)[32], float2 (&im)[32], float4* tile4,
                                                  const float2* __restrict__ s_H, const float2* __restrict__ s_ws,
                                                  int lane) {
    // ---- column of lane 0 ----
    if (lane == 0) {
#pragma unroll
        for (int d1 = 0; d1 < 32; ++d1) tile4[d1] = make_float4(re[d1].x, re[d1].y, im[d1].x, im[d1].y);
    }
    __syncwarp();
    {
        const int pl = (32 - lane) & 31;
        const float4 zk = tile4[lane], zc = tile4[pl];
        const int k = 32 * lane;
        const float2 w = s_ws[k], h = s_H[k];
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
        const float2 zcar = shfl2(re[jb], src), zcai = shfl2(im[jb], src);     // partner of my element ja
        const float2 zcbr = shfl2(re[ja], src), zcbi = shfl2(im[ja], src);     // partner of my element jb
        const int ka = lane + 32 * ja, kb = lane + 32 * jb;
        const float2 wa = s_ws[ka], wb = s_ws[kb], ha = s_H[ka], hb = s_H[kb];
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
    float2* s_tw = reinterpret_cast<float2*>(s_raw);
    float2* s_win = reinterpret_cast<float2*>(s_raw + 8192);
    float2* s_H = reinterpret_cast<float2*>(s_raw + 16384);
    float2* s_ws = reinterpret_cast<float2*>(s_raw + 24576);
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
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s_tw[i] = p.tw_pass[i];
        s_win[i] = p.window[i];
        s_H[i] = p.H[i];
        s_ws[i] = p.tw_split[i];
    }
    __syncthreads();

    uint32_t parity = 0;
    for (; q < npairs; q += nwarps) {
        const bool two = 2 * q + 1 < p.nframes;
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 re[32], im[32];                         // (.x, .y) = (frame 2q, frame 2q+1)
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int m = lane + 32 * b;
            const V2 ra = xstage[m];
            const V2 rb = two ? xstage[1024 + m] : ra;
            const float2 w = s_win[m];
            // packed window multiply; the first butterfly stage takes the products as FMAs by 1.0 (usc_arith.cuh)
            re[b] = __fmul2_rn(make_float2(pcm_to_float(ra.x), pcm_to_float(rb.x)), bc2(w.x));
            im[b] = __fmul2_rn(make_float2(pcm_to_float(ra.y), pcm_to_float(rb.y)), bc2(w.y));
        }
        __syncwarp();
        if (lane == 0 && q + nwarps < npairs) {
            mbar_expect_tx(bar, pair_bytes(q + nwarps));
            bulk_g2s(xstage, pcm + (q + nwarps) * 4096, pair_bytes(q + nwarps), bar);
        }
        fft1024_pair<true>(re, im, tile, s_tw, lane);
        spectral_in_place(re, im, reinterpret_cast<float4*>(tile), s_H, s_ws, lane);
        fft1024_pair(re, im, tile, s_tw, lane);
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
