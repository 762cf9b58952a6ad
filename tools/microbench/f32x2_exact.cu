// Bit-exactness check of packed f32x2 operations against scalar round-to-nearest operations, in the
// operand forms the kernels use (pair*pair, pair*broadcast scalar, chained multiplies, fma forms).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(const float* x, const float* c, const float* w, uint32_t* bad, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float xa = x[2 * i], xb = x[2 * i + 1], ca = c[2 * i], cb = c[2 * i + 1], ww = w[i];
    float2 X = make_float2(xa, xb);
    // form 1: chained multiply, mixed pair then broadcast
    float2 p1 = __fmul2_rn(__fmul2_rn(X, make_float2(ca, cb)), make_float2(ww, ww));
    float s1a = __fmul_rn(__fmul_rn(xa, ca), ww), s1b = __fmul_rn(__fmul_rn(xb, cb), ww);
    // form 2: same table value in both halves
    float2 p2 = __fmul2_rn(__fmul2_rn(X, make_float2(ca, ca)), make_float2(ww, ww));
    float s2a = __fmul_rn(__fmul_rn(xa, ca), ww), s2b = __fmul_rn(__fmul_rn(xb, ca), ww);
    // form 3: fma with broadcast
    float2 p3 = __ffma2_rn(X, make_float2(ca, ca), make_float2(-ww, -ww));
    float s3a = __fmaf_rn(xa, ca, -ww), s3b = __fmaf_rn(xb, ca, -ww);
    uint32_t m = 0;
    m |= (__float_as_uint(p1.x) != __float_as_uint(s1a)) << 0;
    m |= (__float_as_uint(p1.y) != __float_as_uint(s1b)) << 1;
    m |= (__float_as_uint(p2.x) != __float_as_uint(s2a)) << 2;
    m |= (__float_as_uint(p2.y) != __float_as_uint(s2b)) << 3;
    m |= (__float_as_uint(p3.x) != __float_as_uint(s3a)) << 4;
    m |= (__float_as_uint(p3.y) != __float_as_uint(s3b)) << 5;
    if (m) atomicOr(bad, m), atomicAdd(bad + 1, 1u);
}

int main() {
    const int n = 1 << 22;
    float *x, *c, *w; uint32_t* bad;
    cudaMallocManaged(&x, 2 * n * 4); cudaMallocManaged(&c, 2 * n * 4); cudaMallocManaged(&w, n * 4); cudaMallocManaged(&bad, 8);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s; };
    for (int i = 0; i < 2 * n; ++i) { x[i] = (float) ((int32_t) rnd() >> 6); c[i] = (float) ((int32_t) rnd()) / 2147483648.0f; }
    for (int i = 0; i < n; ++i) w[i] = (float) (rnd() >> 8) / 16777216.0f;
    bad[0] = bad[1] = 0;
    k<<<(n + 255) / 256, 256>>>(x, c, w, bad, n);
    cudaDeviceSynchronize();
    printf("f32x2 exactness: mismatch mask 0x%x, count %u of %d  (%s)\n", bad[0], bad[1], n, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
