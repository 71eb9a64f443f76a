// mas_common.cuh -- shared device helpers (sm_100a only).
//
// PTX wrappers: mbarrier producer/consumer hand-offs with bounded waits, TMA
// bulk copies (cp.async.bulk -> SASS UBLKCP), release/acquire progress flags
// in shared memory.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mas_b200.h"

namespace masb200 {

constexpr int kTileFrames = 32;        // mel frames per tile == bits per direction word
constexpr int kTilePitch = 32;         // floats per text-position row of a value tile (one 128-byte swizzle row)
constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------
// shared-memory address helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------
// programmatic dependent launch (griddepcontrol): see lp_mas_fused.cu / path_ops.cu
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
// non-blocking probe of a phase (mbarrier.test_wait never suspends)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
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
// The same for a whole (converged) warp with a WARP-UNIFORM exit: every lane polls, the loop
// condition is a vote, so ptxas can prove the warp is still converged afterwards (a per-lane exit
// makes it emit a divergent fallback copy of every following shuffle region and schedule the
// common path conservatively).
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity) {
    if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
    const long long t0 = clock64();
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
}
__device__ __forceinline__ bool mbar_test_warp(uint64_t *bar, uint32_t parity) {
    return __all_sync(0xffffffffu, mbar_test(bar, parity));
}
// one lane of the (converged) warp; the predicate is produced afresh by the instruction itself
__device__ __forceinline__ bool elect_one() {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " elect.sync _|p, 0xffffffff;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok));
    return ok != 0;
}
// lane-predicated (branch-free) arrive / release: `pred` selects the one lane that acts
__device__ __forceinline__ void mbar_arrive_if(uint64_t *bar, uint32_t pred) {
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 " setp.ne.u32 p, %1, 0;\n"
                 " @p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
                 "}\n" ::"r"(smem_u32(bar)), "r"(pred) : "memory");
}

// ---------------------------------------------------------------------------
// TMA tensor-map load global -> shared (3-D tiled box), completion counted in bytes
// on an mbarrier.  SASS: UTMALDG.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// shared -> global tensor-map store (bulk async-group completion).  SASS: UTMASTG.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tmap, const void *src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(src_smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all committed groups are complete (their global writes are performed)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP): src/dst 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void tma_bulk_load_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 1-D bulk copy shared -> global (bulk async-group completion): dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void tma_bulk_store_1d(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
// at most N committed bulk groups may still be reading shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
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
__device__ __forceinline__ void flag_release_if(int *flag, int v, uint32_t pred) {
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 " setp.ne.u32 p, %2, 0;\n"
                 " @p st.release.cta.shared::cta.s32 [%0], %1;\n"
                 "}\n" ::"r"(smem_u32(flag)), "r"(v), "r"(pred) : "memory");
}
// device-scope flags in global memory (cross-CTA hand-off between the log-prior and MAS kernels)
__device__ __forceinline__ void gflag_release(int *flag, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ void gflag_add_release(int *flag, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ int gflag_acquire(const int *flag) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}
// whole-warp wait with a warp-uniform exit (see mbar_wait_warp); returns the smallest value seen
__device__ __forceinline__ int flag_wait_ge_warp(const int *flag, int target) {
    int v = flag_acquire(flag);
    if (__all_sync(0xffffffffu, v >= target)) return v;
    const long long t0 = clock64();
    while (true) {
        v = flag_acquire(flag);
        if (__all_sync(0xffffffffu, v >= target)) return v;
        if (clock64() - t0 > (1ll << 31)) __trap();
    }
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
// Row x (text position, local to the row pass) is stored as 32 consecutive frames
// (128 bytes) at SLOT sigma(x):
//     sigma(x) = (x % R) * (XP / R) + x / R          (R = rows per DP lane)
// so the rows the 32 lanes of a warp read in the same instruction (same r = x % R,
// consecutive lanes) sit in consecutive slots.  Inside a slot the eight 16-byte
// chunks (4 frames each) are XOR-permuted by the slot index, chunk c at position
// c ^ (slot & 7): exactly the TMA/UMMA 128-byte swizzle, so a tensor-map load with
// box {32 frames, NB*R rows} and elementStrides {1, R} drops the rows
// {r, r+R, r+2R, ...} of a tile straight into this layout -- R (or a few more)
// TMA requests per 32-frame tile, each several KB.  A quarter-warp's eight 16-byte
// reads of one chunk then hit eight distinct bank groups: conflict-free LDS.128.
// slot & 7 == lane & 7 for every row of a lane (XP/R is a multiple of 8).
// ---------------------------------------------------------------------------
template <int R, int XP>
__device__ __forceinline__ int tile_slot(int x_local) {
    return (x_local % R) * (XP / R) + x_local / R;
}

// lanes (= rows of one residue r) covered by one TMA box: a multiple of 32 that divides
// XP/R and keeps the traversed extent NB*R within the 256-element box limit.
__host__ __device__ constexpr int tma_box_lanes(int R, int W) {
    for (int wb = 4; wb >= 1; --wb)
        if (W % wb == 0 && 32 * wb * R <= 256) return 32 * wb;
    return 32;
}

}  // namespace masb200
