// Microbenchmark: shared-memory load / store throughput per SM by access width (32 / 64 / 128 bit per lane), conflict-free
// unit-stride rows, 8 warps per SM — the exchange-tile and table traffic of the fused kernels is priced with these numbers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_widths smem_widths.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int WIDTH, bool STORE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float* out, float seed) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my = s_raw + warp * 8192;                        // 8 KB per warp
    for (int i = lane; i < 2048; i += 32) reinterpret_cast<float*>(my)[i] = seed * i;
    __syncthreads();
    float acc = 0.f;
    constexpr int PER = 8192 / (32 * WIDTH);                        // instructions per 8 KB sweep
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int row = (j + it) & (PER - 1);
            if (WIDTH == 4) {
                float* q = reinterpret_cast<float*>(my) + row * 32 + lane;
                if (STORE) *q = acc + j; else acc += *q;
            } else if (WIDTH == 8) {
                float2* q = reinterpret_cast<float2*>(my) + row * 32 + lane;
                if (STORE) *q = make_float2(acc, j); else { float2 t = *q; acc += t.x + t.y; }
            } else {
                float4* q = reinterpret_cast<float4*>(my) + row * 32 + lane;
                if (STORE) *q = make_float4(acc, j, it, seed); else { float4 t = *q; acc += (t.x + t.y) + (t.z + t.w); }
            }
        }
        if (STORE) acc += 1.f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + reinterpret_cast<float*>(my)[lane];
}
template <int WIDTH, bool STORE, int WARPS>
void run(float* d) {
    cudaFuncSetAttribute(k<WIDTH, STORE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * 8192);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WIDTH, STORE, WARPS><<<148, WARPS * 32, WARPS * 8192>>>(d, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<WIDTH, STORE, WARPS><<<148, WARPS * 32, WARPS * 8192>>>(d, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%s.%-3d %2d warps/SM  %8.3f ms  %6.1f B/clk/SM\n", STORE ? "STS" : "LDS", WIDTH * 8, WARPS, ms, (double) ITERS * WARPS * 8192 / clk);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    run<4, false, 8>(d); run<8, false, 8>(d); run<16, false, 8>(d);
    run<4, true, 8>(d); run<8, true, 8>(d); run<16, true, 8>(d);
    run<4, false, 16>(d); run<8, false, 16>(d); run<16, false, 16>(d);
    run<8, false, 4>(d); run<16, false, 4>(d);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
}
