// feasibility probe: host -> device copy rate from default pinned, write-combined pinned and registered pageable memory,
// one and two copy streams (the e2e path is bound by this rate).  nvcc -O2 -o h2d_probe h2d_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
static float run(void* d, const void* h, size_t bytes, int streams) {
    cudaStream_t st[4];
    for (int i = 0; i < streams; ++i) cudaStreamCreate(&st[i]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, st[0]);
        const size_t part = bytes / streams;
        for (int i = 0; i < streams; ++i) {
            if (i) cudaStreamWaitEvent(st[i], e0, 0);
            cudaMemcpyAsync((char*) d + i * part, (const char*) h + i * part, part, cudaMemcpyHostToDevice, st[i]);
        }
        for (int i = 1; i < streams; ++i) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st[i]); cudaStreamWaitEvent(st[0], e, 0); }
        cudaEventRecord(e1, st[0]);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}
int main() {
    const size_t bytes = (size_t) 155648 * 8192;
    void *d, *hp, *hw, *hr;
    cudaMalloc(&d, bytes);
    cudaHostAlloc(&hp, bytes, cudaHostAllocDefault);
    cudaHostAlloc(&hw, bytes, cudaHostAllocWriteCombined);
    hr = aligned_alloc(4096, bytes);
    memset(hp, 1, bytes); memset(hw, 1, bytes); memset(hr, 1, bytes);
    cudaHostRegister(hr, bytes, cudaHostRegisterDefault);
    const char* names[3] = {"pinned (default)", "pinned (write-combined)", "registered pageable"};
    void* hs[3] = {hp, hw, hr};
    for (int k = 0; k < 3; ++k)
        for (int s = 1; s <= 2; ++s) {
            const float ms = run(d, hs[k], bytes, s);
            printf("%-26s %d stream(s): %7.2f ms  %6.1f GB/s\n", names[k], s, ms, bytes / ms / 1e6);
        }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
