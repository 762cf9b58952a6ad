// Reproducer: ptxas 12.9 (sm_100a) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 — also with -fmad=false —
// although scalar mul.rn.f32 + add.rn.f32 are never contracted.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -fmad=false -cubin ptxas_f32x2_contraction.cu && cuobjdump -sass *.cubin | grep -E "FMUL2|FADD2|FFMA2"
// -> the last FMUL2 and the FADD2 appear as one FFMA2.  The kernels avoid the pattern (usc_arith.cuh).
#include <cuda_runtime.h>
__device__ __forceinline__ float2 bc2(float c) { return make_float2(c, c); }
__global__ void k(const float4* __restrict__ c4, const float2* __restrict__ w2, const int2* __restrict__ x2, float2* out) {
    int i = threadIdx.x;
    float4 c = c4[i]; float2 w = w2[i]; int2 r = x2[i];
    float x0 = __int2float_rn(r.x), x1 = __int2float_rn(r.y);
    float2 re = __fmul2_rn(__fmul2_rn(make_float2(c.x, c.y), bc2(x0)), bc2(w.x));
    float2 im = __fmul2_rn(__fmul2_rn(make_float2(c.z, c.w), bc2(x1)), bc2(w.y));
    out[i] = __fadd2_rn(re, im);
}
