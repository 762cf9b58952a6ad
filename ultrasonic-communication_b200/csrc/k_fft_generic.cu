// k_fft_generic.cu — the canonical mixed-radix FFT for any supported length, one CTA per transform
// with the whole transform resident in shared memory.  Serves the CMSIS-shaped
// operators usc_arm_rfft_fast_f32_batch / usc_arm_cfft_f32_batch; the fused 2048-point chains use
// the warp-per-frame kernels instead.  Same arithmetic as the oracle's fft_rec():
//   level l works on contiguous sub-blocks of n_l points; B = rad[l], A = n_l / B
//     for each a < A: gather x[a + A b] (b < B) -> base FFT -> element d *= W_{n_l}^(a d) (d != 0)
//                     -> scatter to the SAME addresses (position a + A d)
//   after the last level element (d0, d1, ...) of output index k = d0 + B0 (d1 + B1 (...)) sits at
//   position d0 A0 + d1 A1 + ...; a digit-reversal copy restores natural order.
#include "usc_kernels.cuh"
#include "usc_launch.h"

namespace usc {

constexpr int kFftThreads = 256;

// one float2 of padding per 32 keeps the last level (each thread owns 32 CONTIGUOUS points, i.e. a
// stride of 32 float2 between lanes) free of shared-memory bank conflicts
__device__ __forceinline__ uint32_t pad(uint32_t i) { return i + (i >> 5); }

template <int B>
__device__ __forceinline__ void level_items(float2* s, uint32_t n_total, uint32_t n_l, const float2* tw,
                                            uint32_t tw_n, bool last) {
    const uint32_t A = n_l / B;
    const uint32_t items = n_total / B;
    const uint32_t tw_step = tw_n / n_l;
    const uint32_t la = 31u - (uint32_t) __clz(A);          // all lengths are powers of two
    for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
        const uint32_t blk = it >> la, a = it & (A - 1u);
        const uint32_t base = blk * n_l + a;
        float re[B], im[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            float2 v = s[pad(base + A * b)];
            re[b] = v.x;
            im[b] = v.y;
        }
        fft_base<B>(re, im);
#pragma unroll
        for (int d = 0; d < B; ++d) {
            float xr = re[d], xi = im[d];
            if (!last && d != 0) {
                float2 w = tw[(size_t) a * d * tw_step];
                cmul(re[d], im[d], w.x, w.y, xr, xi);
            }
            s[pad(base + A * d)] = make_float2(xr, xi);
        }
    }
}

__device__ __forceinline__ void run_levels(float2* s, const fft_plan_dev& plan) {
    uint32_t n_l = plan.n;
    for (uint32_t l = 0; l < plan.nrad; ++l) {
        const bool last = l + 1 == plan.nrad;
        switch (plan.rad[l]) {
            case 2: level_items<2>(s, plan.n, n_l, plan.tw, plan.tw_n, last); break;
            case 4: level_items<4>(s, plan.n, n_l, plan.tw, plan.tw_n, last); break;
            case 8: level_items<8>(s, plan.n, n_l, plan.tw, plan.tw_n, last); break;
            case 16: level_items<16>(s, plan.n, n_l, plan.tw, plan.tw_n, last); break;
            default: level_items<32>(s, plan.n, n_l, plan.tw, plan.tw_n, last); break;
        }
        n_l /= plan.rad[l];
        __syncthreads();
    }
}

// position in the level buffer of output index k
__device__ __forceinline__ uint32_t out_position(const fft_plan_dev& plan, uint32_t k) {
    uint32_t pos = 0, ln = 31u - (uint32_t) __clz(plan.n);
    for (uint32_t l = 0; l < plan.nrad; ++l) {
        const uint32_t B = plan.rad[l], lb = 31u - (uint32_t) __clz(B);
        ln -= lb;                                      // log2(A)
        pos += (k & (B - 1u)) << ln;
        k >>= lb;
    }
    return pos;
}

// Shared memory: two buffers of n float2 (work, natural-order copy).
template <int MODE>
__global__ void __launch_bounds__(kFftThreads) k_fft_generic(fft_plan_dev plan, const float* in, float* out,
                                                             uint32_t batch) {
    extern __shared__ float2 s_fft[];
    float2* work = s_fft;
    float2* nat = s_fft + pad(plan.n) + 1;
    const uint32_t n = plan.n;
    for (uint32_t v = blockIdx.x; v < batch; v += gridDim.x) {
        const float2* src = reinterpret_cast<const float2*>(in) + (size_t) v * n;
        float2* dst = reinterpret_cast<float2*>(out) + (size_t) v * n;
        if (MODE == FFT_C2R) {
            // merge: packed X -> 2Z (natural) -> swap(re,im) for the inverse-by-forward trick
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
                float zr, zi;
                if (k == 0) {
                    float2 x0 = src[0];
                    zr = __fadd_rn(x0.x, x0.y);
                    zi = __fsub_rn(x0.x, x0.y);
                } else {
                    float2 xk = src[k], xc = src[n - k];
                    float2 w = plan.tw[k];                   // master table has 2n entries: W_N^k
                    rfft_merge(xk.x, xk.y, xc.x, xc.y, w.x, -w.y, zr, zi);
                }
                work[pad(k)] = make_float2(zi, zr);
            }
        } else {
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
                float2 x = src[k];
                work[pad(k)] = MODE == FFT_C2C_INV ? make_float2(x.y, x.x) : x;
            }
        }
        __syncthreads();
        run_levels(work, plan);
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) nat[k] = work[pad(out_position(plan, k))];
        __syncthreads();
        if (MODE == FFT_C2C_FWD) {
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) dst[k] = nat[k];
        } else if (MODE == FFT_C2C_INV) {
            const float sc = 1.0f / (float) n;
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x)
                dst[k] = make_float2(__fmul_rn(nat[k].y, sc), __fmul_rn(nat[k].x, sc));
        } else if (MODE == FFT_C2R) {
            const float sc = 1.0f / (float) (2 * n);
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x)
                dst[k] = make_float2(__fmul_rn(nat[k].y, sc), __fmul_rn(nat[k].x, sc));
        } else {   // FFT_R2C: split into the packed spectrum
            for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
                float xr, xi;
                if (k == 0) {
                    xr = __fadd_rn(nat[0].x, nat[0].y);
                    xi = __fsub_rn(nat[0].x, nat[0].y);
                } else {
                    float2 zk = nat[k], zc = nat[n - k];
                    float2 w = plan.tw[k];
                    rfft_split(zk.x, zk.y, zc.x, zc.y, w.x, -w.y, xr, xi);
                }
                dst[k] = make_float2(xr, xi);
            }
        }
        __syncthreads();
    }
}

// ---- transforms too large for one CTA's shared memory (complex length 16384 / 32768) ---------------
// Level 0 of the canonical plan runs in global memory (one thread per (vector, a): gather B elements
// at stride A — coalesced across a —, base FFT, twiddle, scatter), the B sub-transforms of length
// A <= 1024... 8192 then run through k_fft_generic in place, and a last pass undoes the digit
// reversal of level 0 (k = B c + d) and, for real input, applies the split.
template <int B>
__global__ void k_fft_level0(fft_plan_dev plan, const float2* __restrict__ in, float2* __restrict__ work, uint32_t batch) {
    const uint32_t A = plan.n / B;
    const size_t items = (size_t) batch * A;
    const uint32_t tw_step = plan.tw_n / plan.n;
    for (size_t it = (size_t) blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (size_t) gridDim.x * blockDim.x) {
        const size_t v = it / A;
        const uint32_t a = (uint32_t) (it - v * A);
        const float2* src = in + v * plan.n + a;
        float2* dst = work + v * plan.n + a;
        float re[B], im[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            float2 x = src[(size_t) A * b];
            re[b] = x.x;
            im[b] = x.y;
        }
        fft_base<B>(re, im);
#pragma unroll
        for (int d = 0; d < B; ++d) {
            float xr = re[d], xi = im[d];
            if (d != 0) {
                float2 w = plan.tw[(size_t) a * d * tw_step];
                cmul(re[d], im[d], w.x, w.y, xr, xi);
            }
            dst[(size_t) A * d] = make_float2(xr, xi);
        }
    }
}

// work holds, per vector, B sub-spectra of length A in natural order: Z[B c + d] = work[d A + c]
template <int MODE>
__global__ void k_fft_final(fft_plan_dev plan, const float2* __restrict__ work, float2* __restrict__ out, uint32_t batch) {
    const uint32_t n = plan.n, B = plan.rad[0], A = n / B;
    const size_t total = (size_t) batch * n;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const size_t v = i / n;
        const uint32_t k = (uint32_t) (i - v * n);
        const float2* w = work + v * n;
        const float2 zk = w[(size_t) (k % B) * A + k / B];
        if (MODE == FFT_C2C_FWD) {
            out[i] = zk;
        } else {   // FFT_R2C
            float xr, xi;
            if (k == 0) {
                xr = __fadd_rn(zk.x, zk.y);
                xi = __fsub_rn(zk.x, zk.y);
            } else {
                const uint32_t kc = n - k;
                const float2 zc = w[(size_t) (kc % B) * A + kc / B];
                const float2 t = plan.tw[k];
                rfft_split(zk.x, zk.y, zc.x, zc.y, t.x, -t.y, xr, xi);
            }
            out[i] = make_float2(xr, xi);
        }
    }
}

cudaError_t launch_fft_large(int mode, const fft_plan_dev& plan, const float* in, float* out, float* work,
                             uint32_t batch, cudaStream_t st) {
    if (mode != FFT_C2C_FWD && mode != FFT_R2C) return cudaErrorInvalidValue;
    const uint32_t B = plan.rad[0], A = plan.n / B;
    const size_t items = (size_t) batch * A;
    int grid = (int) ((items + 255) / 256 < 148u * 16u ? (items + 255) / 256 : 148u * 16u);
    if (grid < 1) grid = 1;
    const float2* src = reinterpret_cast<const float2*>(in);
    float2* w = reinterpret_cast<float2*>(work);
    switch (B) {
        case 2: k_fft_level0<2><<<grid, 256, 0, st>>>(plan, src, w, batch); break;
        case 4: k_fft_level0<4><<<grid, 256, 0, st>>>(plan, src, w, batch); break;
        case 8: k_fft_level0<8><<<grid, 256, 0, st>>>(plan, src, w, batch); break;
        case 16: k_fft_level0<16><<<grid, 256, 0, st>>>(plan, src, w, batch); break;
        default: k_fft_level0<32><<<grid, 256, 0, st>>>(plan, src, w, batch); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    fft_plan_dev sub = plan;                      // B*batch sub-transforms of length A, same master table
    sub.n = A;
    sub.nrad = plan.nrad - 1;
    for (uint32_t i = 0; i + 1 < plan.nrad; ++i) sub.rad[i] = plan.rad[i + 1];
    e = launch_fft_generic(FFT_C2C_FWD, sub, work, work, batch * B, st);
    if (e != cudaSuccess) return e;
    const size_t total = (size_t) batch * plan.n;
    int g2 = (int) ((total + 255) / 256 < 148u * 32u ? (total + 255) / 256 : 148u * 32u);
    if (mode == FFT_R2C) k_fft_final<FFT_R2C><<<g2, 256, 0, st>>>(plan, w, reinterpret_cast<float2*>(out), batch);
    else k_fft_final<FFT_C2C_FWD><<<g2, 256, 0, st>>>(plan, w, reinterpret_cast<float2*>(out), batch);
    return cudaGetLastError();
}

cudaError_t fft_generic_prepare() {
    const int max_smem = 136 * 1024;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_fft_generic<FFT_C2C_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))) return e;
    if ((e = cudaFuncSetAttribute(k_fft_generic<FFT_C2C_INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))) return e;
    if ((e = cudaFuncSetAttribute(k_fft_generic<FFT_R2C>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))) return e;
    if ((e = cudaFuncSetAttribute(k_fft_generic<FFT_C2R>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem))) return e;
    return cudaSuccess;
}

cudaError_t launch_fft_generic(int mode, const fft_plan_dev& plan, const float* in, float* out, uint32_t batch,
                               cudaStream_t st) {
    const size_t smem = sizeof(float2) * (2 * (size_t) plan.n + plan.n / 32 + 2);
    if (smem > 136 * 1024) return cudaErrorInvalidValue;
    int grid = batch < 148u * 16u ? (int) batch : 148 * 16;
    if (grid < 1) grid = 1;
    switch (mode) {
        case FFT_C2C_FWD: k_fft_generic<FFT_C2C_FWD><<<grid, kFftThreads, smem, st>>>(plan, in, out, batch); break;
        case FFT_C2C_INV: k_fft_generic<FFT_C2C_INV><<<grid, kFftThreads, smem, st>>>(plan, in, out, batch); break;
        case FFT_R2C: k_fft_generic<FFT_R2C><<<grid, kFftThreads, smem, st>>>(plan, in, out, batch); break;
        case FFT_C2R: k_fft_generic<FFT_C2R><<<grid, kFftThreads, smem, st>>>(plan, in, out, batch); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace usc
