// Microbenchmark: can tensor memory (TMEM) serve per-lane tables beside the shared-memory data path?
// tcgen05.ld.32x32b (SASS LDTM) gives thread i of a warp N consecutive 32-bit columns of TMEM lane
// 32*(warp%4)+i: a lane-private table store that does not go through the LSU / L1 data pipe.
// Measures, per SM and per iteration of 8 warps: LDTM alone, LDS.64 alone, both, and each with FFMA2 work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_tables tmem_tables.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldtm8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void ldtm16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sttm8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}

// MODE bits: 1 = LDTM (8 KB per warp-iteration), 2 = LDS.64 (8 KB per warp-iteration), 4 = 256 FFMA2, 8 = LDTM as x16
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, float seed, int* ok) {
    __shared__ uint32_t s_taddr;
    extern __shared__ __align__(16) float2 sm[];                // 64 KB: 1024 float2 per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_taddr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 32 * 32 * 8; i += 256) sm[i] = make_float2(i * 0.001f, seed);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_taddr + ((uint32_t) (32 * (warp & 3)) << 16);
    if (warp < 4) {                                             // fill this quadrant: column c of lane l holds l*1000 + c
        for (int c = 0; c < 256; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint((float) ((32 * warp + lane) * 1000 + c + j));
            sttm8(tbase + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    {                                                           // correctness: every warp reads its quadrant back
        uint32_t v[8];
        ldtm8(tbase + 40, v);
        ldtm_wait();
        if (__uint_as_float(v[3]) != (float) ((32 * (warp & 3) + lane) * 1000 + 43)) atomicAdd(ok, 1);
    }

    float2 p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i);
    const float2 w2 = make_float2(seed * 0.999f, seed * 0.999f), v2 = make_float2(seed * 0.001f, seed * 0.001f);
    float acc = 0.f;
    const uint2* my = reinterpret_cast<const uint2*>(sm) + warp * 1024;
    uint32_t xa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < ITERS; ++it) {
        if (MODE & 1) {
            if (MODE & 8) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[16];
                    ldtm16(tbase + ((c * 16 + it) & 255 & ~15), v);
                    ldtm_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) xa[j & 7] ^= v[j];
                }
            } else {
                uint32_t v[8][8];
#pragma unroll
                for (int c = 0; c < 8; ++c) ldtm8(tbase + ((c * 8 + it * 8) & 255), v[c]);
                ldtm_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int j = 0; j < 8; ++j) xa[j] ^= v[c][j];
            }
        }
        if (MODE & 2) {
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const uint2 t = my[((b + it) & 31) * 32 + lane];
                xa[b & 7] ^= t.x ^ t.y;
            }
        }
        if (MODE & 4) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = __ffma2_rn(p[i], w2, v2);
        }
    }
    float s = acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __uint_as_float(xa[j]);
#pragma unroll
    for (int i = 0; i < 16; ++i) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(s_taddr));
}

template <int MODE>
void run(const char* name, float* d, int* ok) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<MODE><<<148, 256, 65536>>>(d, 1.0f, ok);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s ERROR %s\n", name, cudaGetErrorString(e)); return; }
    cudaEventRecord(e0);
    k<MODE><<<148, 256, 65536>>>(d, 1.0f, ok);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * 1.965e9 / ITERS;             // SM clocks per iteration (8 warps x the per-warp work)
    printf("%-40s %8.3f ms  %8.1f clk per iteration of 8 warps", name, ms, clk);
    if (MODE & 1) printf("  LDTM %.1f B/clk/SM", 8 * 8192.0 / clk);
    if (MODE & 2) printf("  LDS %.1f B/clk/SM", 8 * 8192.0 / clk);
    if (MODE & 4) printf("  FFMA2 %.1f lane-ops/clk/SM", 8 * 256 * 64.0 / clk);
    printf("\n");
}

int main() {
    float* d;
    int* ok;
    cudaMalloc(&d, 148 * 256 * sizeof(float));
    cudaMalloc(&ok, 4);
    cudaMemset(ok, 0, 4);
    run<1>("LDTM x8 (8 per wait)", d, ok);
    run<9>("LDTM x16 (1 per wait)", d, ok);
    run<2>("LDS.64", d, ok);
    run<3>("LDTM x8 + LDS.64", d, ok);
    run<4>("FFMA2", d, ok);
    run<5>("LDTM x8 + FFMA2", d, ok);
    run<6>("LDS.64 + FFMA2", d, ok);
    run<7>("LDTM x8 + LDS.64 + FFMA2", d, ok);
    int h = -1;
    cudaMemcpy(&h, ok, 4, cudaMemcpyDeviceToHost);
    printf("readback mismatches: %d\nstatus: %s\n", h, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
