// usc_arith.cuh — canonical fp32 arithmetic of the demodulation chain on the device.
//
// Every floating-point operation on the hot path is written with the round-to-nearest intrinsics
// (__fadd_rn/__fsub_rn/__fmul_rn/__fmaf_rn/__fsqrt_rn): the compiler never contracts, splits or
// reassociates them, so the operation order below IS the specification.  The CPU oracle
// (oracle/ref_dsp.c, test infrastructure) states the same order independently in C; parity tests
// require bit-identical floats.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace usc {

// W_32^k = (cos, -sin)(2*pi*k/32), k < 16, rounded once from double (glibc), as hex floats.
// tests/test_abi.py checks these against the host table builder.
__host__ __device__ constexpr float w32_re(int k) {
    constexpr float t[16] = {0x1p+0f,          0x1.f6297cp-1f,  0x1.d906bcp-1f,  0x1.a9b662p-1f,
                             0x1.6a09e6p-1f,   0x1.1c73b4p-1f,  0x1.87de2ap-2f,  0x1.8f8b84p-3f,
                             0x1.1a6264p-54f,  -0x1.8f8b84p-3f, -0x1.87de2ap-2f, -0x1.1c73b4p-1f,
                             -0x1.6a09e6p-1f,  -0x1.a9b662p-1f, -0x1.d906bcp-1f, -0x1.f6297cp-1f};
    return t[k];
}
__host__ __device__ constexpr float w32_im(int k) {
    constexpr float t[16] = {-0x0p+0f,         -0x1.8f8b84p-3f, -0x1.87de2ap-2f, -0x1.1c73b4p-1f,
                             -0x1.6a09e6p-1f,  -0x1.a9b662p-1f, -0x1.d906bcp-1f, -0x1.f6297cp-1f,
                             -0x1p+0f,         -0x1.f6297cp-1f, -0x1.d906bcp-1f, -0x1.a9b662p-1f,
                             -0x1.6a09e6p-1f,  -0x1.1c73b4p-1f, -0x1.87de2ap-2f, -0x1.8f8b84p-3f};
    return t[k];
}

__host__ __device__ constexpr int brev(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// (a + jb)(c + jd): re = fma(a, c, -(b*d)), im = fma(a, d, b*c)        [arm_cmplx_mult_cmplx_f32]
__device__ __forceinline__ void cmul(float ar, float ai, float br, float bi, float& re, float& im) {
    float t0 = __fmul_rn(ai, bi);
    float t1 = __fmul_rn(ai, br);
    re = __fmaf_rn(ar, br, -t0);
    im = __fmaf_rn(ar, bi, t1);
}

// sqrt(fma(re, re, im*im)) with IEEE sqrt                                   [arm_cmplx_mag_f32]
__device__ __forceinline__ float cmag(float re, float im) {
    return __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
}

// Radix-2 DIT butterflies of the base kernel.  (E, O) -> (s, d) in place.
__device__ __forceinline__ void bfly_one(float& er, float& ei, float& or_, float& oi) {   // w = 1
    float sr = __fadd_rn(er, or_), si = __fadd_rn(ei, oi);
    float dr = __fsub_rn(er, or_), di = __fsub_rn(ei, oi);
    er = sr; ei = si; or_ = dr; oi = di;
}
__device__ __forceinline__ void bfly_mj(float& er, float& ei, float& or_, float& oi) {    // w = -j
    float sr = __fadd_rn(er, oi), si = __fsub_rn(ei, or_);
    float dr = __fsub_rn(er, oi), di = __fadd_rn(ei, or_);
    er = sr; ei = si; or_ = dr; oi = di;
}
__device__ __forceinline__ void bfly_gen(float& er, float& ei, float& or_, float& oi, float wr, float wi) {
    float sr = __fmaf_rn(or_, wr, __fmaf_rn(-oi, wi, er));
    float si = __fmaf_rn(or_, wi, __fmaf_rn(oi, wr, ei));
    float dr = __fmaf_rn(2.0f, er, -sr);
    float di = __fmaf_rn(2.0f, ei, -si);
    er = sr; ei = si; or_ = dr; oi = di;
}

// Base kernel: forward FFT of R <= 32 points held in registers, natural order in and out.
// Radix-2 decimation in time; twiddle j of a size-2h block is W_32^(j*16/h); trivial when j == 0
// (w = 1) or 2j == h (w = -j).  Outputs that the caller never reads are removed by the compiler
// (everything is unrolled into straight-line register code), which is how the pruned last pass of
// the fused demodulator is obtained without changing a single retained operation.
template <int R>
__device__ __forceinline__ void fft_base(float (&re)[R], float (&im)[R]) {
    constexpr int BITS = ilog2(R);
    float tr[R], ti[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        tr[i] = re[brev(i, BITS)];
        ti[i] = im[brev(i, BITS)];
    }
#pragma unroll
    for (int h = 1; h < R; h <<= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; ++j) {
                const int a = blk + j, b = blk + j + h;
                if (j == 0) bfly_one(tr[a], ti[a], tr[b], ti[b]);
                else if (2 * j == h) bfly_mj(tr[a], ti[a], tr[b], ti[b]);
                else bfly_gen(tr[a], ti[a], tr[b], ti[b], w32_re(j * (16 / h)), w32_im(j * (16 / h)));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        re[i] = tr[i];
        im[i] = ti[i];
    }
}

// ---- packed (f32x2) twin of the base kernel -------------------------------------------------------
// Blackwell executes add/mul/fma.f32x2 (SASS FADD2/FMUL2/FFMA2) on register pairs with immediate
// and negated operands.  Each half is an independent IEEE round-to-nearest operation, so running
// two transforms side by side in the halves of a float2 gives bit-identical results to two scalar
// runs while halving the issue slots; the fused demodulator carries the up-chirp hypothesis in .x
// and the down-chirp hypothesis in .y.
// CAVEAT (ptxas 12.9): a mul.rn.f32x2 whose only consumer is an add.rn.f32x2 is contracted into FFMA2 even
// with -fmad=false (scalar mul.rn/add.rn never are).  Every packed multiply here therefore feeds an FMA
// addend/multiplicand or a scalar operation, never a packed add; products that must reach a packed add
// (front-end window multiplies) are done with scalar __fmul_rn.  tools/microbench shows the reproducer.
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 bc2(float c) { return make_float2(c, c); }

__device__ __forceinline__ void bfly2_one(float2& er, float2& ei, float2& or_, float2& oi) {
    float2 sr = __fadd2_rn(er, or_), si = __fadd2_rn(ei, oi);
    float2 dr = __fadd2_rn(er, neg2(or_)), di = __fadd2_rn(ei, neg2(oi));
    er = sr; ei = si; or_ = dr; oi = di;
}
__device__ __forceinline__ void bfly2_mj(float2& er, float2& ei, float2& or_, float2& oi) {
    float2 sr = __fadd2_rn(er, oi), si = __fadd2_rn(ei, neg2(or_));
    float2 dr = __fadd2_rn(er, neg2(oi)), di = __fadd2_rn(ei, or_);
    er = sr; ei = si; or_ = dr; oi = di;
}
__device__ __forceinline__ void bfly2_gen(float2& er, float2& ei, float2& or_, float2& oi, float wr, float wi) {
    float2 sr = __ffma2_rn(or_, bc2(wr), __ffma2_rn(neg2(oi), bc2(wi), er));
    float2 si = __ffma2_rn(or_, bc2(wi), __ffma2_rn(oi, bc2(wr), ei));
    float2 dr = __ffma2_rn(bc2(2.0f), er, neg2(sr));
    float2 di = __ffma2_rn(bc2(2.0f), ei, neg2(si));
    er = sr; ei = si; or_ = dr; oi = di;
}

// First-stage butterfly whose inputs are PACKED PRODUCTS (the window multiplies of the fused front ends): written as
// fused multiply-adds by `one` (= 1.0f read from a table, so the compiler cannot fold it): fma(O, 1, E) rounds E + O
// once — the same bits as the addition — and a product can be neither the addend nor a multiplicand of a contracted
// FMA, so ptxas's mul.f32x2 + add.f32x2 -> FFMA2 contraction (see CAVEAT above) has nothing to grab.
__device__ __forceinline__ void bfly2_one_fma(float2& er, float2& ei, float2& or_, float2& oi, float one) {
    const float2 p1 = bc2(one), m1 = bc2(-one);
    float2 sr = __ffma2_rn(or_, p1, er), si = __ffma2_rn(oi, p1, ei);
    float2 dr = __ffma2_rn(or_, m1, er), di = __ffma2_rn(oi, m1, ei);
    er = sr; ei = si; or_ = dr; oi = di;
}

// fft_base2 for inputs that are packed products (see bfly2_one_fma); identical results
template <int R>
__device__ __forceinline__ void fft_base2_prod(float2 (&re)[R], float2 (&im)[R], float one) {
    constexpr int BITS = ilog2(R);
    float2 tr[R], ti[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        tr[i] = re[brev(i, BITS)];
        ti[i] = im[brev(i, BITS)];
    }
#pragma unroll
    for (int h = 1; h < R; h <<= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; ++j) {
                const int a = blk + j, b = blk + j + h;
                if (h == 1) bfly2_one_fma(tr[a], ti[a], tr[b], ti[b], one);
                else if (j == 0) bfly2_one(tr[a], ti[a], tr[b], ti[b]);
                else if (2 * j == h) bfly2_mj(tr[a], ti[a], tr[b], ti[b]);
                else bfly2_gen(tr[a], ti[a], tr[b], ti[b], w32_re(j * (16 / h)), w32_im(j * (16 / h)));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        re[i] = tr[i];
        im[i] = ti[i];
    }
}

template <int R>
__device__ __forceinline__ void fft_base2(float2 (&re)[R], float2 (&im)[R]) {
    constexpr int BITS = ilog2(R);
    float2 tr[R], ti[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        tr[i] = re[brev(i, BITS)];
        ti[i] = im[brev(i, BITS)];
    }
#pragma unroll
    for (int h = 1; h < R; h <<= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; ++j) {
                const int a = blk + j, b = blk + j + h;
                if (j == 0) bfly2_one(tr[a], ti[a], tr[b], ti[b]);
                else if (2 * j == h) bfly2_mj(tr[a], ti[a], tr[b], ti[b]);
                else bfly2_gen(tr[a], ti[a], tr[b], ti[b], w32_re(j * (16 / h)), w32_im(j * (16 / h)));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        re[i] = tr[i];
        im[i] = ti[i];
    }
}

// Forward real-FFT split for one bin k in [1, N/2): Zk = Z[k], Zc = Z[N/2-k], W_N^k = cr - j*si.
//   2X = (Zk + conj Zc) + W_N^k * (Zk - conj Zc)/j ;  X = 0.5 * 2X
__device__ __forceinline__ void rfft_split(float zkr, float zki, float zcr, float zci, float cr, float si,
                                           float& xr, float& xi) {
    float pr = __fadd_rn(zkr, zcr), pi = __fsub_rn(zki, zci);
    float qr = __fadd_rn(zki, zci), qi = __fsub_rn(zcr, zkr);
    float tr = __fmaf_rn(qr, cr, __fmaf_rn(qi, si, pr));
    float ti = __fmaf_rn(qi, cr, __fmaf_rn(-qr, si, pi));
    xr = __fmul_rn(0.5f, tr);
    xi = __fmul_rn(0.5f, ti);
}

// Inverse real-FFT merge for one bin k in [1, N/2): Xk = X[k], Xc = X[N/2-k]  ->  2Z[k]
__device__ __forceinline__ void rfft_merge(float xkr, float xki, float xcr, float xci, float cr, float si,
                                           float& zr, float& zi) {
    float pr = __fadd_rn(xkr, xcr), pi = __fsub_rn(xki, xci);
    float qr = __fsub_rn(xkr, xcr), qi = __fadd_rn(xki, xci);
    zr = __fmaf_rn(-qi, cr, __fmaf_rn(-qr, si, pr));
    zi = __fmaf_rn(qr, cr, __fmaf_rn(-qi, si, pi));
}

// ---- packed forms with a scalar (per-lane) second operand broadcast to both halves ------------------
// (SASS: FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32x2 — no repacking moves).  Half by half identical to the
// scalar functions above.
__device__ __forceinline__ void cmul2(float2 ar, float2 ai, float br, float bi, float2& re, float2& im) {
    const float2 t0 = __fmul2_rn(ai, bc2(bi)), t1 = __fmul2_rn(ai, bc2(br));
    re = __ffma2_rn(ar, bc2(br), neg2(t0));
    im = __ffma2_rn(ar, bc2(bi), t1);
}
__device__ __forceinline__ void rfft_split2(float2 zkr, float2 zki, float2 zcr, float2 zci, float cr, float si,
                                            float2& xr, float2& xi) {
    const float2 pr = __fadd2_rn(zkr, zcr), pi = __fadd2_rn(zki, neg2(zci));
    const float2 qr = __fadd2_rn(zki, zci), qi = __fadd2_rn(zcr, neg2(zkr));
    const float2 tr = __ffma2_rn(qr, bc2(cr), __ffma2_rn(qi, bc2(si), pr));
    const float2 ti = __ffma2_rn(qi, bc2(cr), __ffma2_rn(neg2(qr), bc2(si), pi));
    xr = __fmul2_rn(bc2(0.5f), tr);
    xi = __fmul2_rn(bc2(0.5f), ti);
}
__device__ __forceinline__ void rfft_merge2(float2 xkr, float2 xki, float2 xcr, float2 xci, float cr, float si,
                                            float2& zr, float2& zi) {
    const float2 pr = __fadd2_rn(xkr, xcr), pi = __fadd2_rn(xki, neg2(xci));
    const float2 qr = __fadd2_rn(xkr, neg2(xcr)), qi = __fadd2_rn(xki, xci);
    zr = __ffma2_rn(neg2(qi), bc2(cr), __ffma2_rn(neg2(qr), bc2(si), pr));
    zi = __ffma2_rn(qr, bc2(cr), __ffma2_rn(neg2(qi), bc2(si), pi));
}

// ---- fully packed forms: both halves carry their own second operand (K8's spectral stage: bin k in .x, bin N/2 - k in .y) ----
__device__ __forceinline__ void cmul2v(float2 ar, float2 ai, float2 br, float2 bi, float2& re, float2& im) {
    const float2 t0 = __fmul2_rn(ai, bi), t1 = __fmul2_rn(ai, br);
    re = __ffma2_rn(ar, br, neg2(t0));
    im = __ffma2_rn(ar, bi, t1);
}
__device__ __forceinline__ void rfft_split2v(float2 zkr, float2 zki, float2 zcr, float2 zci, float2 cr, float2 si, float2& xr, float2& xi) {
    const float2 pr = __fadd2_rn(zkr, zcr), pi = __fadd2_rn(zki, neg2(zci));
    const float2 qr = __fadd2_rn(zki, zci), qi = __fadd2_rn(zcr, neg2(zkr));
    const float2 tr = __ffma2_rn(qr, cr, __ffma2_rn(qi, si, pr));
    const float2 ti = __ffma2_rn(qi, cr, __ffma2_rn(neg2(qr), si, pi));
    xr = __fmul2_rn(bc2(0.5f), tr);
    xi = __fmul2_rn(bc2(0.5f), ti);
}
__device__ __forceinline__ void rfft_merge2v(float2 xkr, float2 xki, float2 xcr, float2 xci, float2 cr, float2 si, float2& zr, float2& zi) {
    const float2 pr = __fadd2_rn(xkr, xcr), pi = __fadd2_rn(xki, neg2(xci));
    const float2 qr = __fadd2_rn(xkr, neg2(xcr)), qi = __fadd2_rn(xki, xci);
    zr = __ffma2_rn(neg2(qi), cr, __ffma2_rn(neg2(qr), si, pr));
    zi = __ffma2_rn(qr, cr, __ffma2_rn(neg2(qi), si, pi));
}
__device__ __forceinline__ float2 swap2(float2 a) { return make_float2(a.y, a.x); }

// arm_max_f32 combine: keep the larger value; on equal values keep the lower index.
__device__ __forceinline__ void argmax_combine(float& v, uint32_t& i, float ov, uint32_t oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void warp_argmax(float& v, uint32_t& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        uint32_t oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmax_combine(v, i, ov, oi);
    }
}

}  // namespace usc
