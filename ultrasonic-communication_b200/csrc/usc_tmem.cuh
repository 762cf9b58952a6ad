// usc_tmem.cuh — tensor memory (TMEM) as a lane-private table store for the fused kernels.
//
// The fused demodulators read the same per-lane table entries for every frame: lane a of a warp always needs the
// reference-chirp, Hann and inter-pass twiddle values of m = a + 32 b (b = 0..31).  Served from shared memory these
// reads are 254 of the 633 L1 wavefronts a dual-hypothesis frame costs — the binding resource.
// Blackwell's tensor memory (256 KB per SM, 128 lanes x 512 columns of 32 bits) is reached by its own instructions
// (tcgen05.ld / tcgen05.st, SASS LDTM / STTM): with the .32x32b shape thread i of a warp reads N consecutive columns
// of TMEM lane 32 (warp % 4) + i, i.e. exactly a lane-private row.  Measured (tools/microbench/tmem_tables.cu): LDTM
// delivers 322 B per clock and SM against 118-124 for LDS.64, and costs about half as much again when it runs beside
// shared-memory loads — so the tables go to TMEM (one replica per lane quadrant, written once per CTA) and the
// shared-memory path is left to the PCM stage and the exchange tile.  No tensor-core instruction is involved.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace usc {

// ---- allocation (one warp, whole CTA waits on the barrier that follows) ------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
    static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "TMEM columns: power of two >= 32");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(smem_slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// address of this warp's lane quadrant: bits 31..16 = lane, 15..0 = column
__device__ __forceinline__ uint32_t tmem_quadrant(uint32_t base, int warp) { return base + ((uint32_t) (32 * (warp & 3)) << 16); }

// ---- stores (table set-up) -----------------------------------------------------------------------------------
__device__ __forceinline__ void sttm8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void sttm_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- loads: the wait is part of the same asm statement, so the outputs are defined only once the data has arrived ----
#define USC_R8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
__device__ __forceinline__ void ldtm8(uint32_t taddr, uint32_t (&a)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldtm16(uint32_t taddr, uint32_t (&a)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0), USC_R8(a, 8) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldtm32(uint32_t taddr, uint32_t (&a)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0), USC_R8(a, 8), USC_R8(a, 16), USC_R8(a, 24) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldtm64(uint32_t taddr, uint32_t (&a)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
                 "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0), USC_R8(a, 8), USC_R8(a, 16), USC_R8(a, 24), USC_R8(a, 32), USC_R8(a, 40), USC_R8(a, 48), USC_R8(a, 56)
                 : "r"(taddr) : "memory");
}
// two loads in flight, one wait: 16 + 8 columns (four front-end rows: (up, down) chirp pairs of two samples + Hann pair)
__device__ __forceinline__ void ldtm16_8(uint32_t ta, uint32_t (&a)[16], uint32_t tb, uint32_t (&b)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%24];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%25];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0), USC_R8(a, 8), USC_R8(b, 0) : "r"(ta), "r"(tb) : "memory");
}
// 8 + 4 columns (two front-end rows)
__device__ __forceinline__ void ldtm8_4(uint32_t ta, uint32_t (&a)[8], uint32_t tb, uint32_t (&b)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%12];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8,%9,%10,%11}, [%13];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0), "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(ta), "r"(tb) : "memory");
}
// 32 + 16 columns (eight front-end rows)
__device__ __forceinline__ void ldtm32_16(uint32_t ta, uint32_t (&a)[32], uint32_t tb, uint32_t (&b)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%48];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%49];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : USC_R8(a, 0), USC_R8(a, 8), USC_R8(a, 16), USC_R8(a, 24), USC_R8(b, 0), USC_R8(b, 8) : "r"(ta), "r"(tb) : "memory");
}
#undef USC_R8

}  // namespace usc
