// tc_common.cuh -- tcgen05 building blocks of the warp-specialised kernel (grad_kernel_ws.cuh): shared-memory matrix and
// instruction descriptors for kind::tf32, the canonical K-major operand layout, MMA issue with the A operand in Tensor
// Memory, commit to an mbarrier.  Every form was verified on B200 in tools/experiments/umma_tf32_probe.cu (3xTF32 product
// within 2e-6 of fp64).  History: a first hybrid that kept both phases in one thread (grad_kernel_tc.cuh,
// round 1) passed every parity test in round 2 but ran 21 % slower than the FFMA2 kernel (12.1 vs 10.0 ms per full-batch
// launch: each 128-row tile waited for three MMA round trips in sequence) and was removed; so was the TMEM-parked
// exact-fit-tile variant of the FFMA2 kernel (grad_kernel_v4.cuh: -2.8 %, superseded).
#pragma once
#include "grad_kernel.cuh"
#include "tmem_ops.cuh"

namespace rcmarl {

constexpr int TC_N = 32;             // MMA N (20 hidden units + zero padding)

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// canonical K-major layout of the B operand: element (n, k) at (k / 4) * (32 rows * 4 floats) + n * 4 + k % 4
__host__ __device__ constexpr int tc_canon(int n, int k) { return (k >> 2) * (TC_N * 4) + n * 4 + (k & 3); }

// D[tmem_d] (+)= A[tmem_a] . B[desc]   (A from TMEM, 128 lanes x 8 tf32 columns)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// issue the 3 x (K / 8) split products of one GEMM (one thread); the caller commits (umma_commit) after the last product
template <int K>
__device__ __forceinline__ void tc_issue(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, const float* b_hi,
                                         const float* b_lo, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t boff = ks * 2 * (TC_N * 16);                       // two 16-byte K chunks per k-step
        const uint64_t dbh = umma_smem_desc(smem_u32(b_hi) + boff, TC_N * 16, 128);
        const uint64_t dbl = umma_smem_desc(smem_u32(b_lo) + boff, TC_N * 16, 128);
        umma_ts(tmem_d, tmem_a_hi + ks * 8, dbh, idesc, ks > 0 ? 1u : 0u);
        umma_ts(tmem_d, tmem_a_lo + ks * 8, dbh, idesc, 1u);
        umma_ts(tmem_d, tmem_a_hi + ks * 8, dbl, idesc, 1u);
    }
}

__host__ __device__ constexpr int round32(int n) { return (n + 31) & ~31; }   // 128-byte alignment of the B operands

}  // namespace rcmarl
