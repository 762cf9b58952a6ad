// k_iq.cu — K5 front end of the I/Q baseband path (experiments/iq_modulation/Src/iq_modem.c:52-66):
// carrier mix (arm_mult_f32 x2), 27-tap FIR on I and Q (arm_fir_f32, state carried from the previous
// frame of the same stream), decimation by 2 and interleave to R = I + jQ (BASELINE config 3), fused
// so the PCM crosses HBM once.  One CTA per (stream, frame); the mixed samples (with the numTaps-1
// history samples from the previous frame) are staged in shared memory.  The rest of the chain
// (x conj/plain baseband chirp, Hann, 1024-pt complex FFT, magnitude, windowed arg-max) runs on the
// batched operators; k_iq_pick applies the left/right choice rule of the receiver (main.c:191-197).
#include "usc_kernels.cuh"
#include "usc_launch.h"

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

}  // namespace usc
