// Microbenchmark: fp32 pipe throughput on B200 (scalar FFMA vs packed FFMA2 / FADD2 / FMUL2),
// with and without interleaved LDS traffic. Informs the FFT butterfly design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float a[16];
    float2 p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + i + threadIdx.x; p[i] = make_float2(a[i], a[i] * 0.5f); }
    float w = seed * 0.999f, v = seed * 0.001f;
    float2 w2 = make_float2(w, w), v2 = make_float2(v, v);
    __shared__ float2 sm[256 * 4];
    sm[threadIdx.x] = p[0]; sm[threadIdx.x + 256] = p[1]; sm[threadIdx.x + 512] = p[2]; sm[threadIdx.x + 768] = p[3];
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {           // scalar FFMA, 16 independent chains
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], w, v);
        } else if (MODE == 1) {    // packed FFMA2, 16 independent chains (32 flops-lanes)
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = __ffma2_rn(p[i], w2, v2);
        } else if (MODE == 2) {    // packed FADD2
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = __fadd2_rn(p[i], v2);
        } else if (MODE == 3) {    // scalar FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], v);
        } else if (MODE == 4) {    // scalar FFMA with immediate multiplier
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], 0.99951171875f, v);
        } else if (MODE == 5) {    // 16 FFMA2 + 4 LDS.64 per iteration
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = __ffma2_rn(p[i], w2, v2);
#pragma unroll
            for (int i = 0; i < 4; ++i) { float2 t = sm[(threadIdx.x + it + i * 256) & 1023]; p[i].x += t.x; p[i+4].y += t.y; }
        } else if (MODE == 6) {    // 32 scalar FFMA + 4 LDS.64 per iteration
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i] = __fmaf_rn(a[i], w, v); p[i].x = __fmaf_rn(p[i].x, w, v); }
#pragma unroll
            for (int i = 0; i < 4; ++i) { float2 t = sm[(threadIdx.x + it + i * 256) & 1023]; a[i] += t.x; a[i+4] += t.y; }
        } else if (MODE == 7) {    // mixed: 8 FFMA2 + 8 FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = __ffma2_rn(p[i], w2, v2); p[i+8] = __fadd2_rn(p[i+8], v2); }
        } else if (MODE == 8) {    // scalar mixed: 16 FFMA + 16 FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i] = __fmaf_rn(a[i], w, v); p[i].x = __fadd_rn(p[i].x, v); }
        } else if (MODE == 9) {    // LDS.128 only bandwidth
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 t = reinterpret_cast<float4*>(sm)[(threadIdx.x + it * 3 + i * 64) & 511]; a[i] += t.x; a[i+8] += t.w; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter_per_thread, int nblk, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<nblk, 256>>>(d, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<nblk, 256>>>(d, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = lane_ops_per_iter_per_thread * ITERS * 256.0 * nblk;
    printf("%-34s blocks=%5d  %8.3f ms  %8.2f T lane-ops/s  (%.1f lane-ops/clk/SM @1.9GHz,148SM)\n", name, nblk, ms,
           ops / ms / 1e9, ops / (ms * 1e-3) / 148.0 / 1.9e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float) * 4);
    for (int occ = 2; occ <= 8; occ *= 2) {
        int nb = 148 * occ;
        printf("--- %d CTAs of 256 threads per SM ---\n", occ);
        run<0>("scalar FFMA (3-reg)", 16, nb, d);
        run<4>("scalar FFMA (imm)", 16, nb, d);
        run<3>("scalar FADD", 16, nb, d);
        run<1>("packed FFMA2", 32, nb, d);
        run<2>("packed FADD2", 32, nb, d);
        run<7>("packed 8 FFMA2 + 8 FADD2", 32, nb, d);
        run<8>("scalar 16 FFMA + 16 FADD", 32, nb, d);
        run<5>("16 FFMA2 + 4 LDS.64", 32, nb, d);
        run<6>("32 FFMA + 4 LDS.64", 32, nb, d);
        run<9>("8 LDS.128 (bytes as ops/4)", 8 * 16 / 4.0, nb, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
