// tc_common.cuh -- tcgen05 / TMEM helpers for the log-prior contraction (sm_100a only).
//
// Encodings follow the PTX ISA as exposed by CUTLASS's cute/arch/mma_sm100_desc.hpp (instruction and
// shared-memory matrix descriptors) -- restated here, not included, so the library has no header
// dependency outside the CUDA toolkit.
#pragma once

#include "mas_common.cuh"

namespace masb200 {

// ---------------------------------------------------------------------------
// tensor memory: 128 lanes x 512 columns of 32 bits per SM; address = lane << 16 | column.
// A warp may touch lanes 32*(warp_id % 4) .. +31 only.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {     // whole warp, ncols = 2^k >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive columns of the thread's lane  <-  16 registers   (32x32b: thread i of the warp <-> lane base+i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// 32 consecutive columns of the thread's lane -> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// ---------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------
// Instruction descriptor, kind::tf32, fp32 accumulate, A K-major (from TMEM), B K-major (from smem):
//   [4,6) c_format = 1 (F32)   [7,10) a_format = 2 (TF32)   [10,13) b_format = 2 (TF32)
//   [15] a_major = 0 (K)       [16] b_major = 0 (K)         [17,23) N >> 3        [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_tf32_ts(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor, K-major operand, no swizzle ("interleave" canonical layout)
//   ((8,n),2) : ((1,SBO),LBO)   in 16-byte units
// i.e. core matrices of 8 rows (MN) x 16 bytes (4 tf32 along K) stored as 128 contiguous bytes; the two
// K-chunks of one MMA (K = 8) are LBO apart, groups of 8 rows SBO apart.  Verified on B200 with
// scripts/micro/umma_probe.cu (the MN-major 128B-swizzle form read zeros there).
__device__ __forceinline__ uint64_t umma_smem_desc_k_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Shared-memory matrix descriptor, MN-major operand in the 128-byte-swizzle canonical layout
//   ((8,n),(8,k)) : ((1,LBO),(8,SBO))   in 16-byte units
// i.e. rows of 128 bytes (32 fp32 along MN) at 128-byte pitch along K, 8 K-rows (1024 B) per swizzle atom,
// atoms repeating at SBO along K and LBO along MN.  [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B).  The atom must be 1024-byte aligned.
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem_d] (+)= A[tmem_a] * B[smem desc]; one thread issues
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// fp32 -> exact tf32 pair: hi = x rounded to tf32 (10 mantissa bits, to nearest, ties away from zero: what
// cvt.rna.tf32.f32 computes for finite x, as two integer-pipe instructions instead of the four ptxas emits for
// the cvt), lo = (x - hi) truncated to tf32 (x - hi is exact in fp32).  hi + lo carries >= 21 mantissa bits of x;
// both are exactly representable in tf32, so the tensor core's own input handling never acts.
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    const float rem = x - __uint_as_float(hi);
    lo = __float_as_uint(rem) & 0xffffe000u;
}

}  // namespace masb200
