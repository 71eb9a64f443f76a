// mas_common.cuh -- shared device helpers (sm_100a only).
//
// PTX wrappers: mbarrier producer/consumer hand-offs with bounded waits, TMA
// bulk copies (cp.async.bulk -> SASS UBLKCP), release/acquire progress flags
// in shared memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mas_b200.h"

namespace masb200 {

constexpr int kTileFrames = 32;        // mel frames per tile == bits per direction word
constexpr int kTilePitch = 36;         // floats per text-position row of a value tile (32 + 4 pad, see below)
constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------
// shared-memory address helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------
// mbarrier (shared::cta).  arrive = release.cta, try_wait = acquire.cta.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must fail the launch (trap -> cudaErrorLaunchFailure),
// never hang the GPU.  2^31 cycles (~1 s) is far beyond any legitimate wait here.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
}

// ---------------------------------------------------------------------------
// TMA bulk copy global -> shared (1-D, no tensor map): one row segment per
// instruction, completion counted in bytes on an mbarrier.  src, dst and
// `bytes` must be multiples of 16.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------
// progress flags in shared memory (one writer warp, one reader warp)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void flag_release(int *flag, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(flag)), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_acquire(const int *flag) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(flag)) : "memory");
    return v;
}
__device__ __forceinline__ int flag_wait_ge(const int *flag, int target) {
    int v = flag_acquire(flag);
    if (v >= target) return v;
    const long long t0 = clock64();
    while ((v = flag_acquire(flag)) < target) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
    return v;
}

// ---------------------------------------------------------------------------
// Value tile in shared memory: kTileFrames frames of up to XP text positions.
//
// Row x (text position, local to the row pass) is stored as 32 consecutive
// frames at SLOT sigma(x) with a pitch of kTilePitch = 36 floats (144 bytes):
//     sigma(x) = (x % R) * (XP / R) + x / R          (R = rows per DP lane)
// so the rows that the 32 lanes of a warp read in the same instruction
// (same r = x % R, consecutive lanes) sit in consecutive slots, and because
// 144 B = 9 * 16 B, a quarter-warp's eight 16-byte loads hit eight distinct
// 16-byte bank groups: conflict-free LDS.128 of 4 consecutive frames per row.
// Each row is an independent TMA bulk copy, so the permutation costs nothing.
// ---------------------------------------------------------------------------
template <int R, int XP>
__device__ __forceinline__ int tile_slot(int x_local) {
    return (x_local % R) * (XP / R) + x_local / R;
}

}  // namespace masb200
