// k_elementwise.cu — batched CMSIS-shaped element-wise and reduction operators.
// Each is a single streaming pass (bound: HBM); they exist so a caller can run the reference's
// chain operator by operator.  The hot path uses the fused kernels instead.
#include "usc_kernels.cuh"
#include "usc_launch.h"

namespace usc {

static inline int blocks_for(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    if (b > 148u * 32u) b = 148u * 32u;         // grid-stride beyond 32 CTAs per SM
    return b ? (int) b : 1;
}

// (float) buf[i]  — receiver/Src/main.c:663-665
__global__ void k_i32_to_f32(const int32_t* __restrict__ src, float* __restrict__ dst, size_t count) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x, step = (size_t) gridDim.x * blockDim.x;
    const size_t n4 = count / 4;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    if (aligned) {
        for (size_t j = i; j < n4; j += step) {
            int4 v = s4[j];
            d4[j] = make_float4(__int2float_rn(v.x), __int2float_rn(v.y), __int2float_rn(v.z), __int2float_rn(v.w));
        }
        for (size_t j = n4 * 4 + i; j < count; j += step) dst[j] = __int2float_rn(src[j]);
    } else {
        for (size_t j = i; j < count; j += step) dst[j] = __int2float_rn(src[j]);
    }
}

// arm_mult_f32 — arm_math.h:1938-1942
__global__ void k_mult(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                       uint32_t len, uint32_t batch) {
    const size_t total = (size_t) len * batch;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        size_t v = i / len, e = i - v * len;
        dst[v * sd + e] = __fmul_rn(a[v * sa + e], b[v * sb + e]);
    }
}

// Row-wise float4 forms of the three streaming operators of the receiver chain (mult, cmplx_mult_cmplx, cmplx_mag)
// for the common case — rows of at least 256 floats, everything 16-byte aligned: one CTA walks a row, so there is
// no per-element division and every access is a 16-byte vector.  Same rounded operations as the scalar kernels.
__device__ __forceinline__ bool is16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__global__ void __launch_bounds__(256) k_mult_rows(const float* __restrict__ a, size_t sa, const float* __restrict__ b, size_t sb,
                                                   float* __restrict__ dst, size_t sd, uint32_t len4, uint32_t batch) {
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        const float4* pa = reinterpret_cast<const float4*>(a + (size_t) v * sa);
        const float4* pb = reinterpret_cast<const float4*>(b + (size_t) v * sb);
        float4* pd = reinterpret_cast<float4*>(dst + (size_t) v * sd);
        for (uint32_t e = threadIdx.x; e < len4; e += blockDim.x) {
            const float4 x = pa[e], y = pb[e];
            pd[e] = make_float4(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y), __fmul_rn(x.z, y.z), __fmul_rn(x.w, y.w));
        }
    }
}
__global__ void __launch_bounds__(256) k_cmul_rows(const float* __restrict__ a, size_t sa, const float* __restrict__ b, size_t sb,
                                                   float* __restrict__ dst, size_t sd, uint32_t len4, uint32_t batch) {
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        const float4* pa = reinterpret_cast<const float4*>(a + (size_t) v * sa);
        const float4* pb = reinterpret_cast<const float4*>(b + (size_t) v * sb);
        float4* pd = reinterpret_cast<float4*>(dst + (size_t) v * sd);
        for (uint32_t e = threadIdx.x; e < len4; e += blockDim.x) {          // two complex products per float4
            const float4 x = pa[e], y = pb[e];
            float4 r;
            cmul(x.x, x.y, y.x, y.y, r.x, r.y);
            cmul(x.z, x.w, y.z, y.w, r.z, r.w);
            pd[e] = r;
        }
    }
}
__global__ void __launch_bounds__(256) k_cmag_rows(const float* __restrict__ src, size_t ss, float* __restrict__ dst, size_t sd,
                                                   uint32_t out4, uint32_t batch) {
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        const float4* ps = reinterpret_cast<const float4*>(src + (size_t) v * ss);
        float4* pd = reinterpret_cast<float4*>(dst + (size_t) v * sd);
        for (uint32_t e = threadIdx.x; e < out4; e += blockDim.x) {           // four magnitudes from eight floats
            const float4 x = ps[2 * e], y = ps[2 * e + 1];
            pd[e] = make_float4(cmag(x.x, x.y), cmag(x.z, x.w), cmag(y.x, y.y), cmag(y.z, y.w));
        }
    }
}

// arm_scale_f32 — arm_math.h:2508
__global__ void k_scale(const float* src, float scale, float* dst, size_t total) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x)
        dst[i] = __fmul_rn(src[i], scale);
}

// arm_cmplx_mult_cmplx_f32 — arm_math.h:6579-6583
__global__ void k_cmul(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                       uint32_t ncplx, uint32_t batch) {
    const size_t total = (size_t) ncplx * batch;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        size_t v = i / ncplx, e = i - v * ncplx;
        const float* pa = a + v * sa + 2 * e;
        const float* pb = b + v * sb + 2 * e;
        float ar = pa[0], ai = pa[1], br = pb[0], bi = pb[1], re, im;
        cmul(ar, ai, br, bi, re, im);
        float* pd = dst + v * sd + 2 * e;
        pd[0] = re;
        pd[1] = im;
    }
}

// arm_cmplx_mult_real_f32 — arm_math.h:6425-6429
__global__ void k_cmul_real(const float* c, size_t sc, const float* r, size_t sr, float* dst, size_t sd,
                            uint32_t ncplx, uint32_t batch) {
    const size_t total = (size_t) ncplx * batch;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        size_t v = i / ncplx, e = i - v * ncplx;
        float w = r[v * sr + e];
        float re = c[v * sc + 2 * e], im = c[v * sc + 2 * e + 1];
        dst[v * sd + 2 * e] = __fmul_rn(re, w);
        dst[v * sd + 2 * e + 1] = __fmul_rn(im, w);
    }
}

// arm_cmplx_mag_f32 — arm_math.h:6312-6315
__global__ void k_cmag(const float* src, size_t ss, float* dst, size_t sd, uint32_t ncplx, uint32_t batch) {
    const size_t total = (size_t) ncplx * batch;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        size_t v = i / ncplx, e = i - v * ncplx;
        float re = src[v * ss + 2 * e], im = src[v * ss + 2 * e + 1];
        dst[v * sd + e] = cmag(re, im);
    }
}

// In-place form (dst == src, the way the reference always calls it: receiver/Src/main.c:178
// `arm_cmplx_mag_f32(signal, signal, NN)`): magnitude e overwrites a float that is the input of magnitude e/2.
// One CTA owns a row and walks it in ascending chunks of blockDim.x magnitudes; every chunk is read into registers,
// then a barrier, then written — a chunk's outputs land only on inputs of chunks already consumed (or its own).
__global__ void k_cmag_inplace(float* data, size_t stride, uint32_t ncplx, uint32_t batch) {
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        float* p = data + (size_t) v * stride;
        for (uint32_t base = 0; base < ncplx; base += blockDim.x) {
            const uint32_t e = base + threadIdx.x;
            float m = 0.0f;
            if (e < ncplx) m = cmag(p[2 * e], p[2 * e + 1]);
            __syncthreads();
            if (e < ncplx) p[e] = m;
            __syncthreads();
        }
    }
}

// arm_max_f32 — arm_math.h:6537-6541.  One warp per vector; first occurrence of the maximum.
__global__ void k_max(const float* src, size_t ss, uint32_t len, float* result, uint32_t* index, uint32_t batch) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
    for (size_t v = warp; v < batch; v += nwarps) {
        const float* p = src + v * ss;
        float best = -INFINITY;
        uint32_t bi = 0xffffffffu;
        for (uint32_t e = lane; e < len; e += 32) {
            float x = p[e];
            if (bi == 0xffffffffu || best < x) { best = x; bi = e; }
        }
        warp_argmax(best, bi);
        if (lane == 0) {
            result[v] = best;
            if (index) index[v] = bi;
        }
    }
}

// arm_mean_f32 — arm_math.h:6192.  The sum is sequential left to right (canonical order), so one
// thread per vector; vectors here are short (the reference's call is 8 elements, main.c:431).
__global__ void k_mean(const float* src, size_t ss, uint32_t len, float* result, uint32_t batch) {
    for (size_t v = (size_t) blockIdx.x * blockDim.x + threadIdx.x; v < batch; v += (size_t) gridDim.x * blockDim.x) {
        const float* p = src + v * ss;
        float sum = 0.0f;
        for (uint32_t e = 0; e < len; ++e) sum = __fadd_rn(sum, p[e]);
        result[v] = __fdiv_rn(sum, (float) len);
    }
}

// arm_fir_f32 — arm_math.h:1194-1214.  One CTA per stream; y[n] = sum_i hist[n+i]*coeffs[i] with one
// FMA per tap from acc = 0, hist = [state (taps-1) | src block].  State is updated for the next call.
__global__ void k_fir(const float* __restrict__ coeffs, uint32_t taps, float* state, const float* src,
                      float* dst, uint32_t len, uint32_t batch) {
    extern __shared__ float s_fir[];           // taps coeffs + (taps-1+len) history
    float* s_c = s_fir;
    float* s_h = s_fir + taps;
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < taps; i += blockDim.x) s_c[i] = coeffs[i];
        for (uint32_t i = threadIdx.x; i < taps - 1; i += blockDim.x) s_h[i] = state[(size_t) v * (taps - 1) + i];
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) s_h[taps - 1 + i] = src[(size_t) v * len + i];
        __syncthreads();
        for (uint32_t n = threadIdx.x; n < len; n += blockDim.x) {
            float acc = 0.0f;
            for (uint32_t i = 0; i < taps; ++i) acc = __fmaf_rn(s_h[n + i], s_c[i], acc);
            dst[(size_t) v * len + n] = acc;
        }
        for (uint32_t i = threadIdx.x; i < taps - 1; i += blockDim.x) state[(size_t) v * (taps - 1) + i] = s_h[len + i];
        __syncthreads();
    }
}

// Tail of the spectrum analyser fft() (experiments/basic/Src/main.c:117-142): packed spectrum ->
// magnitude * 1/sqrt(N), bins below the AC-coupling frequency forced to 1.0, dB = 10*log10f(mag),
// arg-max of the magnitudes.  One warp per frame.  mag/db: n/2 floats per frame (either may be NULL).
__global__ void k_spectrum_tail(const float* __restrict__ spec, uint32_t n, float inv_sqrt_n, uint32_t ac_bins,
                                float* __restrict__ mag, float* __restrict__ db, float* __restrict__ peak,
                                uint32_t* __restrict__ peak_idx, uint32_t batch) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
    const uint32_t half = n / 2;
    for (size_t v = warp; v < batch; v += nwarps) {
        const float* p = spec + v * n;
        float best = -INFINITY;
        uint32_t bi = 0xffffffffu;
        for (uint32_t e = lane; e < half; e += 32) {
            float m = __fmul_rn(cmag(p[2 * e], p[2 * e + 1]), inv_sqrt_n);     // arm_cmplx_mag_f32 + arm_scale_f32
            if (e < ac_bins) m = 1.0f;                                          // main.c:126-131
            if (mag) mag[v * half + e] = m;
            if (db) db[v * half + e] = 10.0f * log10f(m);                       // main.c:133-135
            if (bi == 0xffffffffu || best < m) { best = m; bi = e; }
        }
        warp_argmax(best, bi);
        if (lane == 0) {
            if (peak) peak[v] = best;
            if (peak_idx) peak_idx[v] = bi;
        }
    }
}
cudaError_t launch_spectrum_tail(const float* spec, uint32_t n, float inv_sqrt_n, uint32_t ac_bins, float* mag, float* db,
                                 float* peak, uint32_t* peak_idx, uint32_t batch, cudaStream_t st) {
    k_spectrum_tail<<<blocks_for((size_t) batch * 32, 256), 256, 0, st>>>(spec, n, inv_sqrt_n, ac_bins, mag, db, peak, peak_idx, batch);
    return cudaGetLastError();
}

// history[] entry of the 4-offset scan (experiments/chirp_compression_freq_domain/Src/main.c:152-157)
__global__ void k_scan_pack(const float* mr, const uint32_t* ir, const float* ml, const uint32_t* il, uint32_t bw8,
                            float* out, uint32_t slot, uint32_t batch) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < batch; v += gridDim.x * blockDim.x) {
        float* e = out + ((size_t) v * 4 + slot) * 4;
        e[0] = mr[v];
        e[1] = ml[v];
        reinterpret_cast<uint32_t*>(e)[2] = ir[v];
        reinterpret_cast<uint32_t*>(e)[3] = bw8 - il[v];
    }
}
cudaError_t launch_scan_pack(const float* mr, const uint32_t* ir, const float* ml, const uint32_t* il, uint32_t bw8,
                             float* out, uint32_t slot, uint32_t batch, cudaStream_t st) {
    k_scan_pack<<<blocks_for(batch, 128), 128, 0, st>>>(mr, ir, ml, il, bw8, out, slot, batch);
    return cudaGetLastError();
}

// Front end of the receiver chain for any frame length: (float) pcm -> x chirp -> x Hann in one pass
// (receiver/Src/main.c:663-665, chirp.c:47-53, main.c:171).  Tables are broadcast over the batch.
template <typename PCM>
__global__ void k_prep(const PCM* __restrict__ pcm, const float* __restrict__ chirp, const float* __restrict__ hann,
                       float* __restrict__ dst, uint32_t n, size_t total) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const uint32_t e = (uint32_t) (i % n);
        dst[i] = __fmul_rn(__fmul_rn(pcm_cast(pcm[i]), chirp[e]), hann[e]);
    }
}
cudaError_t launch_prep(const void* pcm, uint32_t pcm_format, const float* chirp, const float* hann, float* dst, uint32_t n,
                        size_t total, cudaStream_t st) {
    if (pcm_format == 1u) k_prep<int32_t><<<blocks_for(total, 256), 256, 0, st>>>((const int32_t*) pcm, chirp, hann, dst, n, total);
    else k_prep<float><<<blocks_for(total, 256), 256, 0, st>>>((const float*) pcm, chirp, hann, dst, n, total);
    return cudaGetLastError();
}

// Tail of the receiver chain for any frame length: magnitudes of the packed bins [0, window) and their
// first-occurrence arg-max (arm_cmplx_mag_f32 + arm_max_f32, main.c:178, 208).  One warp per frame.
__global__ void k_mag_max(const float* __restrict__ spec, size_t stride, uint32_t window, float* __restrict__ result,
                          uint32_t* __restrict__ index, uint32_t batch) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t) gridDim.x * blockDim.x) >> 5;
    for (size_t v = warp; v < batch; v += nwarps) {
        const float2* p = reinterpret_cast<const float2*>(spec + v * stride);
        float best = -INFINITY;
        uint32_t bi = 0xffffffffu;
        for (uint32_t e = lane; e < window; e += 32) {
            const float2 z = p[e];
            const float m = cmag(z.x, z.y);
            if (bi == 0xffffffffu || best < m) { best = m; bi = e; }
        }
        warp_argmax(best, bi);
        if (lane == 0) {
            result[v] = best;
            if (index) index[v] = bi;
        }
    }
}
cudaError_t launch_mag_max(const float* spec, size_t stride, uint32_t window, float* result, uint32_t* index, uint32_t batch,
                           cudaStream_t st) {
    k_mag_max<<<blocks_for((size_t) batch * 32, 256), 256, 0, st>>>(spec, stride, window, result, index, batch);
    return cudaGetLastError();
}

// symbol decision of the receiver (receiver/Src/main.c:523): down only if strictly greater
__global__ void k_decide(const float* __restrict__ mu, const float* __restrict__ md, uint8_t* __restrict__ bit, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        bit[i] = md[i] > mu[i] ? 0 : 1;
}
cudaError_t launch_decide(const float* mu, const float* md, uint8_t* bit, size_t n, cudaStream_t st) {
    k_decide<<<blocks_for(n, 256), 256, 0, st>>>(mu, md, bit, n);
    return cudaGetLastError();
}

cudaError_t launch_i32_to_f32(const int32_t* src, float* dst, size_t count, cudaStream_t st) {
    k_i32_to_f32<<<blocks_for(count / 4 + 1, 256), 256, 0, st>>>(src, dst, count);
    return cudaGetLastError();
}
cudaError_t launch_mult(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                        uint32_t len, uint32_t batch, cudaStream_t st) {
    const bool rows = len >= 256 && (len & 3u) == 0 && !((sa | sb | sd) & 3u) && ((uintptr_t) a & 15u) == 0 &&
                      ((uintptr_t) b & 15u) == 0 && ((uintptr_t) dst & 15u) == 0 && (a != dst || sa == sd);
    if (rows) k_mult_rows<<<batch < 148u * 16u ? batch : 148u * 16u, 256, 0, st>>>(a, sa, b, sb, dst, sd, len / 4, batch);
    else k_mult<<<blocks_for((size_t) len * batch, 256), 256, 0, st>>>(a, sa, b, sb, dst, sd, len, batch);
    return cudaGetLastError();
}
cudaError_t launch_scale(const float* src, float scale, float* dst, size_t total, cudaStream_t st) {
    k_scale<<<blocks_for(total, 256), 256, 0, st>>>(src, scale, dst, total);
    return cudaGetLastError();
}
cudaError_t launch_cmul(const float* a, size_t sa, const float* b, size_t sb, float* dst, size_t sd,
                        uint32_t ncplx, uint32_t batch, cudaStream_t st) {
    const bool rows = ncplx >= 128 && (ncplx & 1u) == 0 && !((sa | sb | sd) & 3u) && ((uintptr_t) a & 15u) == 0 &&
                      ((uintptr_t) b & 15u) == 0 && ((uintptr_t) dst & 15u) == 0 && (a != dst || sa == sd) && (b != dst || sb == sd);
    if (rows) k_cmul_rows<<<batch < 148u * 16u ? batch : 148u * 16u, 256, 0, st>>>(a, sa, b, sb, dst, sd, ncplx / 2, batch);
    else k_cmul<<<blocks_for((size_t) ncplx * batch, 256), 256, 0, st>>>(a, sa, b, sb, dst, sd, ncplx, batch);
    return cudaGetLastError();
}
cudaError_t launch_cmul_real(const float* c, size_t sc, const float* r, size_t sr, float* dst, size_t sd,
                             uint32_t ncplx, uint32_t batch, cudaStream_t st) {
    k_cmul_real<<<blocks_for((size_t) ncplx * batch, 256), 256, 0, st>>>(c, sc, r, sr, dst, sd, ncplx, batch);
    return cudaGetLastError();
}
cudaError_t launch_cmag(const float* src, size_t ss, float* dst, size_t sd, uint32_t ncplx, uint32_t batch,
                        cudaStream_t st) {
    const bool disjoint = dst + ((size_t) (batch - 1) * sd + ncplx) <= src || src + ((size_t) (batch - 1) * ss + 2 * (size_t) ncplx) <= dst;
    if (!disjoint) {
        // The one aliased form with defined results is CMSIS's in-place call: same base, same row pitch, rows apart.
        if (dst != src || (batch > 1 && (ss != sd || ss < 2 * (size_t) ncplx))) return cudaErrorInvalidValue;
        k_cmag_inplace<<<batch < 148u * 16u ? batch : 148u * 16u, 256, 0, st>>>(dst, ss, ncplx, batch);
        return cudaGetLastError();
    }
    const bool rows = ncplx >= 256 && (ncplx & 3u) == 0 && !((ss | sd) & 3u) && ((uintptr_t) src & 15u) == 0 &&
                      ((uintptr_t) dst & 15u) == 0;
    if (rows) k_cmag_rows<<<batch < 148u * 16u ? batch : 148u * 16u, 256, 0, st>>>(src, ss, dst, sd, ncplx / 4, batch);
    else k_cmag<<<blocks_for((size_t) ncplx * batch, 256), 256, 0, st>>>(src, ss, dst, sd, ncplx, batch);
    return cudaGetLastError();
}
cudaError_t launch_max(const float* src, size_t ss, uint32_t len, float* result, uint32_t* index,
                       uint32_t batch, cudaStream_t st) {
    k_max<<<blocks_for((size_t) batch * 32, 256), 256, 0, st>>>(src, ss, len, result, index, batch);
    return cudaGetLastError();
}
cudaError_t launch_mean(const float* src, size_t ss, uint32_t len, float* result, uint32_t batch,
                        cudaStream_t st) {
    k_mean<<<blocks_for(batch, 128), 128, 0, st>>>(src, ss, len, result, batch);
    return cudaGetLastError();
}
cudaError_t launch_fir(const float* coeffs_dev, uint32_t taps, float* state, const float* src, float* dst,
                       uint32_t len, uint32_t batch, cudaStream_t st) {
    size_t smem = sizeof(float) * ((size_t) taps + taps - 1 + len);
    int grid = batch < 148u * 8u ? (int) batch : 148 * 8;
    k_fir<<<grid, 256, smem, st>>>(coeffs_dev, taps, state, src, dst, len, batch);
    return cudaGetLastError();
}

}  // namespace usc
