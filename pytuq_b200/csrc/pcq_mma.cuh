// Shared pieces of the packed-operand DMMA kernels (pcq_syrk.cu: S = Z^T Z;  pcq_qform.cu: psi^T C psi):
// pipeline geometry, mbarrier / TMA-bulk / DMMA wrappers.
//
// Packed operand layout (both kernels):  P[k-block (8 values of the contraction index)][row][8]  doubles, so
// that a (128 rows x 8) operand chunk is one contiguous 8 KB run for cp.async.bulk and a lane's two k-step
// values sit in one 16-byte word (conflict-free LDS.128).  The contraction index of mma.m8n8k4 is a summation
// index, so lane tig may feed values (2 tig, 2 tig + 1) of a block in its two k-steps as long as both
// operands agree.
#pragma once
#include "pcq_internal.cuh"

namespace {

// The output tile edge ST is a template parameter of the SYRK kernels: 128 (warp tile 32 x 64), or 96 (warp
// tile 24 x 48) where 128 would execute clearly more flops.  Since the diagonal tiles have their own
// sub-block pass (128 only) that is rare: L just below a multiple of 96 and small.
constexpr int SB = 8;              // contraction values per block (two DMMA k-steps)
constexpr int KS = 4;              // blocks per pipeline stage
constexpr int NSTAGE = 3;
constexpr int CONSUMER_WARPS = 8;
constexpr int SY_THREADS = CONSUMER_WARPS * 32;
constexpr int STAGE_SAMPLES = KS * SB;
constexpr int PACK_TS = 18;        // slot-table row stride of the pack kernels

// dynamic shared memory of one pipelined CTA: ring of NSTAGE x (2 operands x KS blocks x st rows x 8) + barriers
inline size_t syrk_smem_bytes(int st) { return (size_t)NSTAGE * 2 * KS * st * SB * sizeof(double) + 2 * NSTAGE * sizeof(uint64_t); }

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

#endif

}  // namespace
