// mas_forward.cuh -- Monotonic Alignment Search forward DP + backtrack, sm_100a.
//
// Restates (bit-exactly) reference model/monotonic_align/core.pyx:9-35
// (maximum_path_each), one utterance per CTA:
//
//   forward   Q[x,y] = max(v_cur, v_prev) + value[x,y]                    core.pyx:17-30
//               v_cur  = (x == y) ? max_neg_val : Q[x,y-1]
//               v_prev = (x == 0) ? (y == 0 ? 0 : max_neg_val) : Q[x-1,y-1]
//               max(a,b) == (b > a) ? b : a   (Cython's lowering; NaN -> v_cur)
//   backtrack index = t_x-1; for y = t_y-1..0: path[index,y] = 1;
//               if index != 0 and (index == y or Q[index,y-1] < Q[index-1,y-1]) index--   :32-35
//
// Formulation (validated against the compiled reference: oracle.maximum_path_numpy,
// tests/test_oracle.py):
//   * only the previous column of Q is kept (registers), never the matrix;
//   * one DIRECTION BIT per cell, d[x,y] = (v_prev > v_cur) -- the very predicate the
//     backtrack re-derives at core.pyx:34 for 1 <= x < y;
//   * no lower band bound (core.pyx:18's max(0, t_x+y-t_y) is an optimisation: the
//     in-band recursion is closed and the backtrack never leaves the band);
//   * cells with x > y, and frames of the last 32-frame tile beyond t_y, are computed as
//     garbage and never consumed; the cell x == y substitutes v_cur = max_neg_val exactly
//     like the reference.
//
// CTA organisation (template R rows per lane, W DP warps; XP = 32*R*W rows):
//   producer warp W      a handful of TMA tensor-map loads (cp.async.bulk.tensor, box
//                        {32 frames, NB*R rows}, elementStrides {1,R}, 128B swizzle) per
//                        32-frame tile into an NS-deep shared-memory ring, completion
//                        counted on an mbarrier; the strided boxes land the rows in the
//                        lane-major permuted layout of mas_common.cuh so DP reads are
//                        conflict-free LDS.128.
//   DP warps 0..W-1      lane l of warp w owns the R consecutive text positions
//                        x = rows_base + (32w + l)*R + r.  A 32-frame tile is ONE
//                        straight-line block (no branch, no divergence): per frame the
//                        lane updates its rows bottom-up in place; the bottom row goes
//                        first and its SHFL.UP to the next lane is issued immediately, so
//                        the ~26-cycle shuffle is covered by the 2R-2 cell updates until
//                        the next frame's top row consumes it.  Lane 0 takes the halo (row
//                        above the warp) with one FSEL; lane 31 leaves it with one
//                        predicated STS.
//   warp skew            warp w trails warp w-1 by one tile; the boundary row travels
//                        through a shared-memory ring of 32-float slots published with
//                        st.release / ld.acquire progress flags (no __syncthreads, no
//                        barrier instruction in the tile loop).  Only warp 0 waits on the
//                        TMA mbarrier; the others inherit the ordering through the flag chain.
//   direction bits       32 frames x 1 bit per row per tile, kept in shared memory when
//                        they fit (LRS2 shapes), otherwise in an L2-resident global
//                        scratch that is staged back through the idle ring.
//   backtrack            one thread walks TOKENS, not frames: the words of the next 8
//                        tokens of the current tile are fetched with 8 independent LDS,
//                        then each token costs one masked find-leading-one (pure ALU
//                        chain), emitting [start, duration] per token.
//   text longer than XP  processed in passes of XP rows; the last row of pass p is
//                        carried to pass p+1 through a global line (L2).
#pragma once

#include "mas_common.cuh"

namespace masb200 {

struct MasParams {
    const float *value;      // [B,Tx,Ty], y contiguous
    long long stride_b, stride_x;
    const int *t_x, *t_y;    // [B]
    int B, Tx, Ty;
    float neg;               // max_neg_val
    int aligned;             // 1: every row segment is 16-byte aligned -> TMA path
    int ring_stages;         // NS
    int *start;              // [B,Tx] first frame of each token (workspace)
    int *dur;                // [B,Tx] frames per token (user buffer or workspace)
    int *frame_token;        // [B,Ty] or nullptr
    int *status;             // [B] or nullptr
    uint32_t *gbits;         // [B][tiles][rows_pitch] or nullptr when bits live in smem
    int gbits_rows_pitch;
    long long gbits_stride_b;
    float *gline;            // [B][2][line_pitch] carry line between row passes
    int line_pitch;
    void *path;              // optional in-kernel dense path write
    int path_dtype;          // MAS_B200_PATH_*
    long long *dbg;          // diagnostics: [B][8] clock64 phase stamps (nullptr normally)
    // one-sided duration gather (multi-GPU, optional): peer_dur[r] = rank r's [peer_world * B, Tx] int32 buffer as mapped into
    // THIS device's address space (peer-to-peer over NVLink); this rank's durations also go to rows [peer_rank * B, +B) of
    // every one of them, straight from the output stage -- no collective kernel, nothing on the host path of a step
    int *const *peer_dur;
    int peer_world, peer_rank;
};

// One cell of the recurrence, two formulations (template parameter CK):
//
// kCellExact   setp.gt p, v_prev, v_cur ; q = v_cur + v ; @p q = v_prev + v ; @p bits |= 1 << BITPOS
//   = (v_prev > v_cur ? v_prev : v_cur) + v in fp32 RN, NaN -> v_cur: core.pyx:22-30 with Cython's max(v_cur, v_prev)
//   lowering, for ANY input (NaN, infinities, signed zeros).  The select is folded into the two (independent) adds, so
//   the frame-to-frame dependency chain is compare -> predicated add.  Bit BITPOS of `bits` is set iff the diagonal
//   predecessor wins.  This is what maximum_path(value, mask) runs: its values are the caller's.
//
// kCellSign    q = fmax(v_cur, v_prev) + v ; bits = (bits << 1) | signbit(v_cur - v_prev)
//   predicate-free: the seven predicate registers no longer bound how many cells ptxas keeps in flight (static
//   schedule 23.5 instead of 29.4 cycles per frame, measured 30.8 instead of 38.2 in isolation), and the word is
//   accumulated by a funnel shift, newest frame at bit 0 -- after the 32 frames of a tile it IS the bit-reversed word
//   the backtrack wants (no per-block shifts, no BREV, no reset).  Identical to kCellExact whenever no NaN enters the
//   recurrence: for non-NaN operands fmax is the select (Q is never -0: (+0) + v and fmax of non-(-0) values are
//   not -0, by induction from Q = 0 / max_neg_val), a - b of distinct floats never rounds to zero and a - a = +0, so
//   signbit(v_cur - v_prev) == (v_prev > v_cur); inf - inf = NaN (canonical, sign 0) where inf > inf is false.
//   A NaN operand differs (fmax drops a NaN v_cur, the select keeps it).  Used by the fused kernel, whose values are
//   its own finite log-prior sums; tests/test_gpu_logprior.py checks its paths bit-exactly against the reference MAS
//   of exactly those values.
#ifndef MAS_CELL_VARIANT
#define MAS_CELL_VARIANT 0
#endif
constexpr int kCellExact = 0, kCellSign = 1;
template <int CK, int BITPOS>
__device__ __forceinline__ float mas_cell(float v_cur, float v_prev, float v, uint32_t &bits) {
    if constexpr (CK == kCellSign) {
        const float q = fmaxf(v_cur, v_prev) + v;
        bits = __funnelshift_l(__float_as_uint(v_cur - v_prev), bits, 1);
        return q;
    } else {
#if MAS_CELL_VARIANT == 0
    float q;
    asm("{\n"
        " .reg .pred p;\n"
        " setp.gt.f32 p, %3, %2;\n"
        " add.rn.f32 %0, %2, %4;\n"
        " @p add.rn.f32 %0, %3, %4;\n"
        " @p or.b32 %1, %1, %5;\n"
        "}\n"
        : "=&f"(q), "+r"(bits)
        : "f"(v_cur), "f"(v_prev), "f"(v), "n"(1u << BITPOS));
    return q;
#elif MAS_CELL_VARIANT == 1
    // predicate-free: d = all-ones iff v_prev > v_cur; both sums off the chain; bitwise select
    uint32_t d, q;
    const float s0 = v_cur + v, s1 = v_prev + v;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(v_prev), "f"(v_cur));
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(q) : "r"(__float_as_uint(s0)), "r"(__float_as_uint(s1)), "r"(d));
    asm("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+r"(bits) : "r"(d), "n"(1u << BITPOS));
    return __uint_as_float(q);
#elif MAS_CELL_VARIANT == 2
    uint32_t d, m;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(v_prev), "f"(v_cur));
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(m) : "r"(__float_as_uint(v_cur)), "r"(__float_as_uint(v_prev)), "r"(d));
    asm("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+r"(bits) : "r"(d), "n"(1u << BITPOS));
    return __uint_as_float(m) + v;
#else
    const bool d = v_prev > v_cur;
    bits |= d ? (1u << BITPOS) : 0u;
    return (d ? v_prev : v_cur) + v;
#endif
    }
}

// Predicate registers are the scarce resource of the tile body: every cell needs one for its
// compare -> predicated add, and ptxas serialises the cells on a single predicate when long-lived
// booleans occupy the other six (measured: 85-137 instead of 34-49 cycles/frame).  So nothing in
// the body is predicated on a loop-invariant condition:
//   * lane 31's halo hand-off is an UNCONDITIONAL 4-byte store whose address is a per-lane register
//     (the real slot for lane 31 of a warp with a consumer, a dump row otherwise);
//   * lane 0's halo take-over is a bitwise select on a per-lane mask register;
//   * loop-invariant flags consumed after the body are laundered through opaque() so their
//     predicates are materialised after the body, not kept alive across it.
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}
__device__ __forceinline__ void sts_s32(uint32_t addr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float bitselect(uint32_t mask, float a, float b) {      // mask ? a : b, bitwise
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(mask));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t opaque(uint32_t x) {
    asm volatile("" : "+r"(x));
    return x;
}

template <int I>
__device__ __forceinline__ float f4_get(const float4 &v) {
    return I == 0 ? v.x : I == 1 ? v.y : I == 2 ? v.z : v.w;
}

// The lane's R rows of frame group g (frames 4g..4g+3 of the tile, g a RUNTIME value): one LDS.128 per row.
//   lane_tile = &stage[lane_cta * kTilePitch]; chunk c of a row sits at position c ^ (lane & 7).
template <int R, int XP>
__device__ __forceinline__ void load_group(float4 (&v)[R], const float *lane_tile, int lane7, int g) {
    constexpr int kRowStride = (XP / R) * kTilePitch;      // floats between the lane's consecutive rows
    const int o = (g ^ lane7) << 2;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = *reinterpret_cast<const float4 *>(lane_tile + r * kRowStride + o);
}

// One frame of the lane's R rows.  K = position of the frame inside its 8-frame block (compile time): the
// direction bit goes to bit 24 + K of the accumulator (the block loop shifts the word right by 8 per block, so
// after four blocks bit k of the word is frame k of the tile), the halo goes to hout_addr + 4*K.
//   q[r]    running previous-column Q of the lane's rows, updated in place bottom-up
//   up      Q of the row above the lane's first row at the previous frame
//   hq      Q of the row above the WARP at this frame (what lane 0 takes instead of a shuffle)
//   dlb     DIAG only: lane_global - (first frame of the block) / R; the lane owns the diagonal cell of this
//           frame, in row K % R, exactly when dlb == K / R      (R | 8 | first frame of the block)
template <int R, int K, bool DIAG, int CK>
__device__ __forceinline__ void dp_frame(float (&q)[R], uint32_t (&acc)[R], float &up, const float (&v)[R], float hq,
                                         uint32_t lane0_mask, int dlb, float neg, uint32_t hout_addr) {
    float up_next = up;
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        float v_cur = q[r];
        if (DIAG && r == K % R) v_cur = (dlb == K / R) ? neg : v_cur;      // x == y (core.pyx:19-20)
        const float v_prev = (r == 0) ? up : q[r - 1];
        q[r] = mas_cell<CK, 24 + K>(v_cur, v_prev, v[r], acc[r]);
        if (r == R - 1) {
            sts_f32(hout_addr + 4u * K, q[R - 1]);
            const float s = __shfl_up_sync(kFullMask, q[R - 1], 1);
            up_next = bitselect(lane0_mask, hq, s);
        }
    }
    up = up_next;
}

// four frames (group H = 0 / 1 of an 8-frame block)
template <int R, int H, bool DIAG, int CK>
__device__ __forceinline__ void dp_group(float (&q)[R], uint32_t (&acc)[R], float &up, const float4 (&v4)[R],
                                         const float4 &h4, uint32_t lane0_mask, int dlb, float neg,
                                         uint32_t hout_addr) {
    float v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = v4[r].x;
    dp_frame<R, 4 * H + 0, DIAG, CK>(q, acc, up, v, h4.x, lane0_mask, dlb, neg, hout_addr);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = v4[r].y;
    dp_frame<R, 4 * H + 1, DIAG, CK>(q, acc, up, v, h4.y, lane0_mask, dlb, neg, hout_addr);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = v4[r].z;
    dp_frame<R, 4 * H + 2, DIAG, CK>(q, acc, up, v, h4.z, lane0_mask, dlb, neg, hout_addr);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = v4[r].w;
    dp_frame<R, 4 * H + 3, DIAG, CK>(q, acc, up, v, h4.w, lane0_mask, dlb, neg, hout_addr);
}

// A whole 32-frame tile as a loop of four 8-frame blocks (value and halo groups register double-buffered across the
// loop).  The block body is ~2.7 KB of code and stays in the instruction cache; the fully unrolled tile (11 KB per
// variant, beside the other roles' loops) did not.
//   va, ha   group 0 of THIS tile, already loaded (dp_tile_prefetch: for tile j + 1 that happens before the tail work
//            of tile j, so the tile starts without a shared-memory round trip); garbage on exit
//   hin      32 floats in shared memory: Q of the row above the warp at the tile's frames
//   acc      kCellExact: must be zero on entry; on exit bit k of acc[r] is the direction bit of frame k
//            kCellSign:  any value on entry; on exit bit 31 - k is the direction bit of frame k (dp_store_words)
template <int R, int XP>
__device__ __forceinline__ void dp_tile_prefetch(float4 (&va)[R], float4 &ha, const float *lane_tile, const float *hin,
                                                 int lane7) {
    load_group<R, XP>(va, lane_tile, lane7, 0);
    ha = *reinterpret_cast<const float4 *>(hin);
}

template <int R, int XP, bool DIAG, int CK = kCellExact>
__device__ __forceinline__ void dp_tile_pre(float (&q)[R], uint32_t (&acc)[R], float &up, float4 (&va)[R], float4 &ha,
                                            const float *lane_tile, const float *hin, int lane7, uint32_t lane0_mask,
                                            int dl0, float neg, uint32_t hout_addr) {
    static_assert(8 % R == 0, "rows per lane must divide the 8-frame block");
    float4 vb[R];
    float4 hb;
    const float4 *h4 = reinterpret_cast<const float4 *>(hin);
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        load_group<R, XP>(vb, lane_tile, lane7, 2 * i + 1); hb = h4[2 * i + 1];
        if constexpr (CK == kCellExact) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] >>= 8;
        }
        // the block's halo address is formed afresh from the loop index: advancing ONE register in place right behind the
        // halo STS.128 made the add wait for the store to read its address operand (2 cycles per frame in the profile)
        const uint32_t hout_i = hout_addr + 32u * (uint32_t)i;
        dp_group<R, 0, DIAG, CK>(q, acc, up, va, ha, lane0_mask, dl0, neg, hout_i);
        // next block's first group; the last block re-reads group 0 (no branch in the body)
        load_group<R, XP>(va, lane_tile, lane7, (2 * i + 2) & 7); ha = h4[(2 * i + 2) & 7];
        dp_group<R, 1, DIAG, CK>(q, acc, up, vb, hb, lane0_mask, dl0, neg, hout_i);
        dl0 -= 8 / R;
    }
}

template <int R, int XP, bool DIAG, int CK = kCellExact>
__device__ __forceinline__ void dp_tile(float (&q)[R], uint32_t (&acc)[R], float &up, const float *lane_tile,
                                        const float *hin, int lane7, uint32_t lane0_mask, int dl0, float neg,
                                        uint32_t hout_addr) {
    float4 va[R];
    float4 ha;
    dp_tile_prefetch<R, XP>(va, ha, lane_tile, hin, lane7);
    dp_tile_pre<R, XP, DIAG, CK>(q, acc, up, va, ha, lane_tile, hin, lane7, lane0_mask, dl0, neg, hout_addr);
}

// Direction words of a finished tile, "walk ready" for the backtrack: bit-reversed (bit 31-k <-> frame 32j + k), the
// forced move of the diagonal cell (index == y, core.pyx:34) OR-ed in, the word of token 0 (which never moves) zero.
// Leaves acc ready for the next tile.
template <int R, int CK>
__device__ __forceinline__ void dp_finish_words(uint32_t (&acc)[R], uint32_t (&out)[R], int x0, int j, bool diag) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t w = acc[r];
        if constexpr (CK == kCellExact) { w = __brev(w); acc[r] = 0u; }
        if (diag && ((x0 + r) >> 5) == j) w |= 0x80000000u >> ((x0 + r) & 31);
        out[r] = w;
    }
    if (x0 == 0) out[0] = 0u;
}

template <int R>
__device__ __forceinline__ void store_words(uint32_t *dst, const uint32_t (&acc)[R]) {
    if constexpr (R == 1) {
        dst[0] = acc[0];
    } else if constexpr (R == 2) {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(acc[0], acc[1]);
    } else if constexpr (R == 4) {
        *reinterpret_cast<uint4 *>(dst) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    } else {
        static_assert(R == 8, "R must be 1, 2, 4 or 8");
        *reinterpret_cast<uint4 *>(dst) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<uint4 *>(dst + 4) = make_uint4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// Shared-memory carve-up; host (launcher) and device agree through these functions.
//   [ring: NS value tiles][halo: (W+1) rings of NS+1 slots x 32 floats, W dump areas][ctrl][direction bits]
// halo ring i < W is written by DP warp i; ring W holds the constant / carried-line input of warp 0.
template <int R, int W>
struct MasSmem {
    static constexpr int XP = 32 * R * W;
    static constexpr int kTileFloats = XP * kTilePitch;
    __host__ __device__ static constexpr size_t ring_bytes(int ns) { return sizeof(float) * (size_t)ns * kTileFloats; }
    __host__ __device__ static constexpr int halo_slots(int ns) { return ns + 1; }
    // halo rings + per DP warp a 640-byte dump area: lanes without a consumer store their (unused) halo value to the
    // 16-byte slot (lane + 4-frame group) of it -- 32 distinct slots per store instruction.  (All lanes storing to ONE
    // address serialised into 32 shared-memory passes per STS.128: 2-8 cycles per frame of write-after-read stalls.)
    static constexpr int kDumpFloats = 160;
    __host__ __device__ static constexpr size_t halo_bytes(int ns) {
        return sizeof(float) * (((size_t)(W + 1) * halo_slots(ns)) * kTileFrames + (size_t)W * kDumpFloats);
    }
    __host__ __device__ static constexpr size_t ctrl_bytes(int ns) { return 8 * (size_t)(2 * ns) + 4 * (size_t)(W + 2) + 64; }
    __host__ __device__ static constexpr size_t fixed_bytes(int ns) {
        return ((ring_bytes(ns) + halo_bytes(ns) + ctrl_bytes(ns) + 127) / 128) * 128;
    }
};

template <typename T> __device__ __forceinline__ uint32_t one_bits();
template <> __device__ __forceinline__ uint32_t one_bits<float>() { return 0x3f800000u; }
template <> __device__ __forceinline__ uint32_t one_bits<int>() { return 1u; }

// One dense path row [Ty] from its (start, duration) entry, written by one warp: zero the row with plain
// 16-byte streaming stores, then patch the few chunks that overlap [s, e).  A chunk is always written by the
// same lane (c & 31), so the two stores to it are ordered.
template <typename T>
__device__ __forceinline__ void write_path_row(T *row_ptr, int s, int d, int Ty, int lane) {
    const uint32_t one = one_bits<T>();
    const int e = s + d;                                 // exclusive; d == 0 -> empty
    uint32_t *out = reinterpret_cast<uint32_t *>(row_ptr);
    if ((Ty & 3) == 0 && (reinterpret_cast<uintptr_t>(row_ptr) & 15) == 0) {
        uint4 *row = reinterpret_cast<uint4 *>(out);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int c = lane; c < (Ty >> 2); c += 32) __stcs(row + c, z);
        for (int c = (s >> 2) + ((lane - (s >> 2)) & 31); c <= ((e - 1) >> 2) && e > s; c += 32) {
            const int y = c << 2;
            uint4 o;
            o.x = (y + 0 >= s && y + 0 < e) ? one : 0u;
            o.y = (y + 1 >= s && y + 1 < e) ? one : 0u;
            o.z = (y + 2 >= s && y + 2 < e) ? one : 0u;
            o.w = (y + 3 >= s && y + 3 < e) ? one : 0u;
            __stcs(row + c, o);
        }
    } else {
        for (int y = lane; y < Ty; y += 32) out[y] = (y >= s && y < e) ? one : 0u;
    }
}

// Dense path rows from the [start, dur] table (4-byte elements): a pure streaming write, one warp per row.
// Lane k fetches the table entry of the warp's k-th row up front (one round trip to L2 per 32 rows).
template <typename T>
__device__ __forceinline__ void write_path_rows(T *path_b, const int *start_b, const int *dur_b, int Tx, int Ty,
                                                int tid, int nthreads) {
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    for (int k0 = 0; warp + k0 * nwarps < Tx; k0 += 32) {
        const int xl = warp + (k0 + lane) * nwarps;
        int s_l = 0, d_l = 0;
        if (xl < Tx) { s_l = start_b[xl]; d_l = dur_b[xl]; }
        for (int k = 0; k < 32; ++k) {
            const int x = warp + (k0 + k) * nwarps;
            if (x >= Tx) break;
            write_path_row<T>(path_b + (size_t)x * Ty, __shfl_sync(kFullMask, s_l, k), __shfl_sync(kFullMask, d_l, k), Ty, lane);
        }
    }
}

__device__ __forceinline__ void write_path_any(const MasParams &P, int b, const int *start_b, const int *dur_b,
                                               int tid, int nthreads) {
    if (P.path == nullptr || P.path_dtype == MAS_B200_PATH_NONE) return;
    const size_t off = (size_t)b * P.Tx * P.Ty;
    if (P.path_dtype == MAS_B200_PATH_F32)
        write_path_rows<float>(reinterpret_cast<float *>(P.path) + off, start_b, dur_b, P.Tx, P.Ty, tid, nthreads);
    else
        write_path_rows<int>(reinterpret_cast<int *>(P.path) + off, start_b, dur_b, P.Tx, P.Ty, tid, nthreads);
}

// ---------------------------------------------------------------------------------------------------
// Backtrack, latency path (direction words in shared memory).  The token walk below is a chain of
// ~t_x + tiles dependent steps run by one thread; most of it moves off the critical path like this:
//   transfer table   nj[j][x] = number of tokens the path completes inside tile j when it enters the tile
//                    (at its last frame) on token x.  Computed for ALL x of a tile by one warp, 32 rows per
//                    lane group, as soon as the DP warps are done with the tile -- by the TMA producer warp,
//                    which is otherwise idle and learns exactly that from the ring's `empty` barrier.
//   after the DP     x_in(j-1) = x_in(j) - nj[j][x_in(j)]: one dependent shared-memory load per TILE, then one
//                    thread per tile re-walks its own tile from x_in(j) and emits the token start frames.
// Same m / mask' = m ^ -m chain as backtrack_walk (words bit-reversed, forced move on the diagonal,
// token 0 never moves).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bt_tile_mask(int j, int ntiles, int t_y) {
    return (j == ntiles - 1) ? (0u - (1u << (31 - ((t_y - 1) & 31)))) : 0xffffffffu;     // frames <= t_y-1
}

// The DP stores the words of the shared-memory path "walk ready": the forced move of the diagonal cell
// (core.pyx:34, index == y) is OR-ed in and the word of token 0 (which never moves) is zero.
// Entry tokens far from the path complete a token per frame (their words are all ones), and real paths run
// along such stretches too, so every chain is followed to its end (at most 32 tokens per tile).  The row
// groups of a tile are split between the helper warp (groups [0, kGH)) and the TMA producer warp (the rest,
// in the slack the NS-deep ring gives it); what the producer did not get to is shared by all warps after the DP.
// Per token the chain is TWO dependent ALU ops: with m = word & mask,
//     mask' = ~(m ^ (m - 1))          (the bits strictly above the lowest set bit of m; m == 0 -> mask' == 0, sticky)
// and the next token's m' = word' & mask' folds into the same LOP3:  m' = word' & ~(m ^ (m - 1)).
template <int NG>
__device__ __forceinline__ void bt_transfer_groups(const uint32_t *p, unsigned char *nj, int xbase, int xmax, uint32_t mask0,
                                                   int lane) {
    uint32_t m[NG], m1[NG];       // m of the previous token and m - 1 (start: m = mask + ... such that ~(m ^ m1) == mask)
    int n[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        // ~(m ^ m1) = mask0 (or 0 for lanes beyond xmax): m = 0, m1 = ~mask
        const uint32_t mk = (xbase + 32 * g + lane <= xmax) ? mask0 : 0u;
        m[g] = 0u; m1[g] = ~mk; n[g] = 0;
    }
    // token 0's word is zero, so a chain that reaches it stops there (sticky zero): indices below row 0 are
    // only ever read with a zero mask and need no bounds check (they stay inside the CTA's shared memory)
    for (int k0 = 0; k0 < 32; k0 += 8) {
        uint32_t w[8][NG];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
            for (int g = 0; g < NG; ++g) w[kk][g] = p[32 * g - (k0 + kk)];
        uint32_t alive = 0u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                uint32_t mn;
                asm("lop3.b32 %0, %1, %2, %3, 0x90;" : "=r"(mn) : "r"(w[kk][g]), "r"(m[g]), "r"(m1[g]));    // a & ~(b ^ c)
                m[g] = mn;
                m1[g] = mn - 1u;
                if (mn != 0u) n[g] = k0 + kk + 1;
                if (kk == 7) alive |= mn;
            }
        }
        if (!__any_sync(kFullMask, alive != 0u)) break;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) nj[32 * g + lane] = (unsigned char)n[g];
}

// groups [G0, G1) of tile j, four at a time, only as many as hold tokens that can be entered (x <= y, x < t_x)
template <int G0, int G1>
__device__ __forceinline__ void bt_tile_transfer(const uint32_t *bits_j, unsigned char *nj_j, int j, int t_x,
                                                 uint32_t mask0, int lane) {
    const int xmax = min(t_x - 1, 32 * j + 31);
#pragma unroll
    for (int gq = G0; gq < G1; gq += 4) {
        const int ng = min(min(4, G1 - gq), (xmax >> 5) + 1 - gq);  // warp-uniform
        const uint32_t *p = bits_j + 32 * gq + lane;
        unsigned char *nj = nj_j + 32 * gq;
        if (ng >= 4) bt_transfer_groups<4>(p, nj, 32 * gq, xmax, mask0, lane);
        else if (ng == 3) bt_transfer_groups<3>(p, nj, 32 * gq, xmax, mask0, lane);
        else if (ng == 2) bt_transfer_groups<2>(p, nj, 32 * gq, xmax, mask0, lane);
        else if (ng == 1) bt_transfer_groups<1>(p, nj, 32 * gq, xmax, mask0, lane);
    }
}

// The same for a RUNTIME share of the tile's row groups: helper h of H takes groups [n*h/H, n*(h+1)/H) of the
// n = (xmax >> 5) + 1 groups that hold enterable tokens, so the helpers' loads are balanced whatever t_x is.
__device__ __forceinline__ void bt_tile_transfer_share(const uint32_t *bits_j, unsigned char *nj_j, int j, int t_x,
                                                       uint32_t mask0, int lane, int h, int H) {
    const int xmax = min(t_x - 1, 32 * j + 31);
    const int n = (xmax >> 5) + 1;
    const int g1 = (n * (h + 1)) / H;
    for (int gq = (n * h) / H; gq < g1; gq += 4) {
        const int ng = min(4, g1 - gq);                              // warp-uniform
        const uint32_t *p = bits_j + 32 * gq + lane;
        unsigned char *nj = nj_j + 32 * gq;
        if (ng >= 4) bt_transfer_groups<4>(p, nj, 32 * gq, xmax, mask0, lane);
        else if (ng == 3) bt_transfer_groups<3>(p, nj, 32 * gq, xmax, mask0, lane);
        else if (ng == 2) bt_transfer_groups<2>(p, nj, 32 * gq, xmax, mask0, lane);
        else bt_transfer_groups<1>(p, nj, 32 * gq, xmax, mask0, lane);
    }
}

// Token walk over direction words wb[(j - jlo) * wpitch + x] (tile j, text position x), the backtrack
// of core.pyx:32-35 restated per TOKEN: token x ends where token x+1 starts, and starts at the highest
// frame y' of its span with d[x,y'] set, or y' == x (core.pyx:34's index == y).
//
// The DP stores the words BIT-REVERSED (bit 31-k <-> frame 32j + k), so "highest frame" is "lowest set
// bit"; with m = word & mask,
//     mask' = m ^ -m            (the bits strictly above the lowest set bit of m; 0 stays 0)
// is the mask of the next token in the same tile: a token costs two dependent ALU ops (IADD3, LOP3)
// once the words are in registers.  KB tokens are fetched per batch with independent loads; a batch
// ends early (m == 0, sticky) when the walk leaves the tile.  One thread runs this chain and leaves
// only raw material behind -- m per token, the entry token per tile -- which all threads turn into
// start frames afterwards.
//   state  x     current token (its start is not known yet)
//          j     tile the walk is in
//          mask  mask of the (reversed) frames of tile j still available to token x
//   tokm[x]  receives m of the tile token x starts in (1 <= x)
//   xin[j]   receives the token the walk enters tile j with (pre-zeroed by the caller)
// Returns true when token 0 is reached, false when the walk needs a tile below jlo.
__device__ __forceinline__ bool backtrack_walk(const uint32_t *wb, int wpitch, int jlo, int &x, int &j,
                                               uint32_t &mask, uint32_t *tokm, int *xin) {
    constexpr int KB = 8;
    while (x > 0) {
        if (mask == 0u) { --j; mask = 0xffffffffu; if (j >= 0) xin[j] = x; }
        if (j < jlo) return false;
        const uint32_t *row = wb + (size_t)(j - jlo) * wpitch + x;
        uint32_t w[KB], m[KB];
        if (x >= KB && (x >> 5) < j) {
            // common case: 8 real tokens, tile strictly above the diagonal (no forced move possible)
#pragma unroll
            for (int k = 0; k < KB; ++k) w[k] = row[-k];
        } else {
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int xi = x - k;
                uint32_t v = row[-min(k, x)];
                if ((xi >> 5) == j) v |= 0x80000000u >> (xi & 31);     // index == y forces the move
                w[k] = (xi > 0) ? v : 0u;                               // token 0 never moves (index != 0)
            }
        }
        uint32_t mk = mask;
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            m[k] = w[k] & mk;
            mk = m[k] ^ (0u - m[k]);
        }
        // unconditional: an unresolved token (m == 0) is overwritten when it is resolved in a lower tile
#pragma unroll
        for (int k = 0; k < KB; ++k)
            if (k == 0 || x - k > 0) tokm[x - k] = m[k];
        // m != 0 is a prefix property (sticky zero): count it by bisection
        int n;
        if (m[3] != 0u) n = (m[5] != 0u) ? ((m[7] != 0u) ? 8 : ((m[6] != 0u) ? 7 : 6)) : ((m[4] != 0u) ? 5 : 4);
        else n = (m[1] != 0u) ? ((m[2] != 0u) ? 3 : 2) : ((m[0] != 0u) ? 1 : 0);
        x -= n;
        mask = (n == KB) ? mk : 0u;                                      // 0: continue in tile j-1 (top of loop)
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Tail shared by mas_forward_kernel (SMEM_BITS) and the fused log-prior + MAS kernel (lp_mas_fused.cu).
// ---------------------------------------------------------------------------------------------------
// Backtrack over walk-ready direction words in shared memory: the transfer tables of row groups [G0, G1) still owed
// for tiles [jt_owed, ntiles) are shared by all warps, then one dependent shared-memory load per TILE gives the
// token each tile is entered with (xin[]), then one LANE per tile walks the tokens that begin in its tile and leaves
// their start frames as one 32-bit START MASK per tile:
//   sflags[j]  bit 31-k set  <=>  a token begins at frame 32j + k      (token 0, which begins at frame 0, has no bit)
// The tokens that begin in tile j are (xin[j-1], xin[j]], in frame order; everything the outputs need follows from the
// masks by population counts (mas_emit_outputs_flags) -- no per-token find-first-set, no scattered stores and no scan.
template <int XP, int NTHREADS, int G0, int G1>
__device__ __forceinline__ void mas_backtrack_smem(const uint32_t *bits_s, unsigned char *nj_s, int *xin, uint32_t *sflags,
                                                   int jt_owed, int ntiles, int t_x, int t_y, int tid,
                                                   long long *dbg = nullptr) {
    const int warp = tid >> 5, lane = tid & 31;
    if constexpr (G0 < G1) {
        constexpr int nwarps = NTHREADS / 32;
        for (int jt = jt_owed + warp; jt < ntiles; jt += nwarps)
            bt_tile_transfer<G0, G1>(bits_s + (size_t)jt * XP, nj_s + (size_t)jt * XP, jt, t_x, bt_tile_mask(jt, ntiles, t_y), lane);
        __syncthreads();
    }
    if (tid == 0) {                                                       // one dependent load per tile
        int x = t_x - 1;
        for (int jt = ntiles - 1; jt >= 0; --jt) {
            xin[jt] = x;
            const int n = nj_s[(size_t)jt * XP + x];
            x -= n;
        }
        if (dbg) dbg[10] = clock64();
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[14] = clock64();
    // All lanes of a warp run the same loop (trip count = the largest token count among its tiles), eight tokens per
    // round: their words are fetched with independent loads, then each token costs two dependent ALU ops
    // (bt_transfer_groups' chain) and one more to add its lowest set bit -- its start frame -- to the mask.  A lane that
    // has run out of tokens reads the word of token 0 of tile 0 instead, which is zero by construction: ONE address for
    // all such lanes (a broadcast), where "its own tile, token hi" would be one bank for every tile a long token spans
    // (a 25-way bank conflict per load on the LRS2-shaped batch).
    for (int j0 = warp * 32; j0 < ntiles; j0 += NTHREADS) {
        const int jt = j0 + lane;
        const bool live = jt < ntiles;
        const uint32_t *bj = bits_s + (size_t)(live ? jt : 0) * XP;
        const int hi = live ? xin[jt] : 0;
        const int lo = (live && jt > 0) ? xin[jt - 1] : 0;
        const int cnt = live ? hi - lo : 0;
        const int maxcnt = __reduce_max_sync(kFullMask, cnt);
        if (dbg && tid == 0) dbg[31] = maxcnt;
        uint32_t m = 0u, m1 = ~(live ? bt_tile_mask(jt, ntiles, t_y) : 0u);      // ~(m ^ m1) == the tile's frame mask
        uint32_t starts = 0u;
        for (int c = 0; c < maxcnt; c += 8) {
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = (c + k < cnt) ? bj[hi - c - k] : bits_s[0];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t mn;
                asm("lop3.b32 %0, %1, %2, %3, 0x90;" : "=r"(mn) : "r"(w[k]), "r"(m), "r"(m1));   // a & ~(b ^ c)
                m = mn;
                m1 = mn - 1u;
                starts |= mn & ~m1;                                        // lowest set bit of mn: where the token begins
            }
        }
        if (live) sflags[jt] = starts;
    }
    if (dbg && tid == 0) dbg[15] = clock64();
    __syncthreads();
}

// Outputs from the per-tile start masks.  The tokens are consecutive along the frames, so with S = sflags[j], hi = xin[j]
//     frame_token[32j + k] = hi - popc(S & ((1 << (31-k)) - 1))        (the starts at LATER frames of the tile)
// one independent population count per frame: no scan, no head marks.  A frame whose own bit is set is the first frame
// of its token (tok[]); [start, duration] per token follow from tok[].  One thread per 4 frames (16-byte stores).
//   tok        scratch in shared memory [XP]: start frame per token
//   path_ones  optional: the (already zero-filled) dense path [Tx,Ty] of this utterance as 32-bit words; every frame's
//              thread stores `one` at (its token, its frame)
template <int NTHREADS>
__device__ __forceinline__ void mas_emit_outputs_flags(const MasParams &P, int b, int *tok, const int *xin, const uint32_t *sflags,
                                                       int ntiles, int t_x, int t_y, int tid, long long *dbg = nullptr,
                                                       uint32_t *path_ones = nullptr, uint32_t one = 0u) {
    int *start_b = P.start + (size_t)b * P.Tx;
    int *dur_b = P.dur + (size_t)b * P.Tx;
    int *ft = P.frame_token ? P.frame_token + (size_t)b * P.Ty : nullptr;
    const bool vec = ((P.Ty & 3) == 0) && ((reinterpret_cast<uintptr_t>(ft) & 15) == 0);
    if (tid == 0) tok[0] = 0;                                              // token 0 begins at frame 0 (it has no bit)
    for (int t0 = 4 * tid; t0 < P.Ty; t0 += 4 * NTHREADS) {
        const int j = t0 >> 5;
        const bool in_tiles = j < ntiles;
        const uint32_t S = in_tiles ? sflags[j] : 0u;
        const int hi = in_tiles ? xin[j] : 0;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = t0 + k;
            const uint32_t bit = 0x80000000u >> (t & 31);
            const int x = hi - __popc(S & (bit - 1u));
            const bool valid = t < t_y;
            if (valid && (S & bit) != 0u) tok[x] = t;
            if (path_ones != nullptr && valid) path_ones[(size_t)x * P.Ty + t] = one;
            v[k] = valid ? x : -1;
        }
        if (ft != nullptr) {
            if (vec) {
                *reinterpret_cast<int4 *>(ft + t0) = make_int4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (t0 + k < P.Ty) ft[t0 + k] = v[k];
            }
        }
    }
    if (dbg && tid == 0) dbg[11] = clock64();
    __syncthreads();
    for (int x = tid; x < P.Tx; x += NTHREADS) {
        int s = 0, d = 0;
        if (x < t_x) {
            s = tok[x];
            d = ((x + 1 < t_x) ? tok[x + 1] : t_y) - s;
        }
        start_b[x] = s;
        dur_b[x] = d;
        if (P.peer_dur != nullptr) {
            const size_t row = ((size_t)P.peer_rank * P.B + b) * P.Tx + x;
            for (int r = 0; r < P.peer_world; ++r) P.peer_dur[r][row] = d;
        }
    }
    __syncthreads();                         // start_b / dur_b of every thread are in place for the path writer
}

// ints of shared-memory scratch the two functions above need: tok [XP] + xin [tiles] + start masks [tiles]
__host__ __device__ constexpr size_t mas_tail_scratch_ints(int XP, int ntiles, int /*Ty*/) {
    return (size_t)XP + 2 * (size_t)((ntiles + 3) & ~3);
}

// MULTIPASS: text longer than XP rows (carry line between row passes); its runtime role flags cost
// the single-pass instantiations nothing.
// Warps: 0..W-1 DP, W TMA producer, W+1 (SMEM_BITS only) backtrack helper.
template <int R, int W, bool SMEM_BITS, bool MULTIPASS>
__global__ void __launch_bounds__((W + 1 + (SMEM_BITS ? 1 : 0)) * 32, 1)
mas_forward_kernel(const MasParams P, const __grid_constant__ CUtensorMap tmap) {
    using S = MasSmem<R, W>;
    constexpr int XP = S::XP;
    constexpr int NT = kTileFrames;
    constexpr int kTileFloats = S::kTileFloats;
    constexpr int nthreads = (W + 1 + (SMEM_BITS ? 1 : 0)) * 32;
    constexpr int kG = XP / 32;                    // row groups of a tile (backtrack transfer tables)
    constexpr int kGH = (kG + 1) / 2;              // groups [0, kGH) -> helper warp, [kGH, kG) -> producer warp

    const int NS = P.ring_stages;
    const int HS = S::halo_slots(NS);

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    float *hbuf = reinterpret_cast<float *>(smem_raw + S::ring_bytes(NS));                 // [W+1][HS][32]
    uint64_t *ring_full = reinterpret_cast<uint64_t *>(smem_raw + S::ring_bytes(NS) + S::halo_bytes(NS));
    uint64_t *ring_empty = ring_full + NS;
    int *hprog = reinterpret_cast<int *>(ring_empty + NS);                                // [W]
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem_raw + S::fixed_bytes(NS));        // SMEM_BITS only
    __shared__ int bt_state[4];

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);     // provably warp-uniform for ptxas
    const int lane = tid & 31;

    // every loop bound below derives from these: broadcast them so the bounds are warp-uniform values
    const int t_x = __shfl_sync(kFullMask, P.t_x[b], 0);
    const int t_y = __shfl_sync(kFullMask, P.t_y[b], 0);
    int *start_b = P.start + (size_t)b * P.Tx;
    int *dur_b = P.dur + (size_t)b * P.Tx;

    // ---- per-item validation (the reference is undefined here: core.pyx:34) ----
    if (t_x < 1 || t_y < 1 || t_x > P.Tx || t_y > P.Ty || t_x > t_y) {
        for (int x = tid; x < P.Tx; x += nthreads) { start_b[x] = 0; dur_b[x] = 0; }
        if (P.frame_token)
            for (int y = tid; y < P.Ty; y += nthreads) P.frame_token[(size_t)b * P.Ty + y] = -1;
        if (P.status && tid == 0) P.status[b] = MAS_B200_ITEM_BAD_LENGTH;
        __syncthreads();
        write_path_any(P, b, start_b, dur_b, tid, nthreads);
        return;
    }
    if (P.status && tid == 0) P.status[b] = MAS_B200_ITEM_OK;

    const int ntiles = (t_y + NT - 1) / NT;
    unsigned char *nj_s = reinterpret_cast<unsigned char *>(bits_s + (size_t)ntiles * XP);   // [ntiles][XP] transfer table
    const int npass = (t_x + XP - 1) / XP;
    long long *dbg = P.dbg ? P.dbg + (size_t)b * 16 : nullptr;
    if (dbg && tid == 0) { dbg[0] = clock64(); long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[12] = t; }
    uint32_t *gbits_b = SMEM_BITS ? nullptr : P.gbits + (size_t)b * P.gbits_stride_b;
    float *gline_b = P.gline ? P.gline + (size_t)b * 2 * P.line_pitch : nullptr;
    const float *vb = P.value + (size_t)b * P.stride_b;
    float *hconst = hbuf + (size_t)W * HS * NT;              // warp 0's halo input ring
    float *hdump = hconst + (size_t)HS * NT;                 // [W][160] where lanes without a consumer store

    for (int pass = 0; pass < npass; ++pass) {
        const int rows_base = pass * XP;
        const int rows_here = min(XP, t_x - rows_base);
        const int w_act = (rows_here + 32 * R - 1) / (32 * R);      // DP warps owning at least one valid row
        const bool last_pass = (pass + 1 == npass);

        if (tid == 0) {
            for (int s = 0; s < NS; ++s) {
                if (pass > 0) {
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&ring_full[s])) : "memory");
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&ring_empty[s])) : "memory");
                }
                mbar_init(&ring_full[s], 1);
                mbar_init(&ring_empty[s], w_act);
            }
            for (int i = 0; i < W; ++i) hprog[i] = 0;
            hprog[W] = 0x7fffffff;                               // the flag a warp without predecessor polls
            hprog[W + 1] = 0;                                    // where a warp without consumer publishes
            mbar_fence_init();
        }
        // pass 0: the row above text position 0 is max_neg_val at every frame (core.pyx:26-27)
        if (pass == 0 && tid < NT) hconst[tid] = P.neg;
        __syncthreads();

        if (warp == W) {
            // ============================ producer warp ============================
            int stage = 0;
            uint32_t phase = 0;
            int jt_next = 0;                                   // next tile whose transfer table (upper groups) is owed
            const int *flag_last = hprog + (w_act - 1);
            for (int j = 0; j < ntiles; ++j) {
                mbar_wait(&ring_empty[stage], phase ^ 1);
                float *dst = ring + (size_t)stage * kTileFloats;
                const int t0 = j * NT;
                const int nfr = min(NT, P.Ty - t0);                 // frames that exist in memory
                if (P.aligned) {
                    // lane i issues request (r, qb): rows {rows_base + (qb*NB + l)*R + r : l < NB}
                    constexpr int NB = tma_box_lanes(R, W);
                    constexpr int REQS = (XP / R) / NB;
                    const int rr = lane / REQS, qb = lane - rr * REQS;
                    const int row0 = rows_base + qb * NB * R + rr;           // first row of the box
                    const bool issue = (lane < R * REQS) && (row0 < t_x);
                    const unsigned m = __ballot_sync(kFullMask, issue);
                    if (lane == 0) mbar_arrive_expect_tx(&ring_full[stage], (uint32_t)__popc(m) * NB * 128u);
                    __syncwarp();
                    if (issue)
                        tma_load_3d(dst + (rr * (XP / R) + qb * NB) * kTilePitch, &tmap, t0, row0, b,
                                    &ring_full[stage]);
                } else {
                    // unaligned fallback: lane = frame, 8 rows in flight
                    const bool tin = lane < nfr;
                    const float *src = vb + t0 + lane;
                    for (int xl0 = 0; xl0 < rows_here; xl0 += 8) {
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            v[u] = (tin && xl0 + u < rows_here)
                                       ? __ldcs(src + (size_t)(rows_base + xl0 + u) * P.stride_x) : 0.f;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (xl0 + u < rows_here) {
                                const int sl = tile_slot<R, XP>(xl0 + u);
                                dst[sl * kTilePitch + ((((lane >> 2) ^ (sl & 7)) << 2) | (lane & 3))] = v[u];
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ring_full[stage]);
                }
                if (++stage == NS) { stage = 0; phase ^= 1; }
                if (SMEM_BITS && kGH < kG && j + 1 < ntiles) {
                    // while the ring is full (the DP is not waiting for this warp) use the slack for the transfer
                    // tables this warp owes; tiles the DP warps have all released are final
                    while (jt_next < ntiles && !mbar_test_warp(&ring_empty[stage], phase ^ 1)) {
                        const int done_tiles = __shfl_sync(kFullMask, flag_acquire(flag_last), 0);
                        if (done_tiles <= jt_next) break;
                        bt_tile_transfer<kGH, kG>(bits_s + (size_t)jt_next * XP, nj_s + (size_t)jt_next * XP, jt_next, t_x,
                                                  bt_tile_mask(jt_next, ntiles, t_y), lane);
                        ++jt_next;
                    }
                }
            }
            if (lane == 0) bt_state[3] = jt_next;      // the rest is shared by all warps once the DP is done
            if (dbg && lane == 0) { dbg[11] = jt_next; dbg[15] = clock64(); }
        } else if (SMEM_BITS && warp == W + 1) {
            // ========================== backtrack helper warp ==========================
            // trails the LAST active DP warp (its progress flag implies every earlier warp is past the tile too)
            const int *flag_last = hprog + (w_act - 1);
            int known = 0;
            for (int jt = 0; jt < ntiles; ++jt) {
                if (known < jt + 1) known = flag_wait_ge_warp(flag_last, jt + 1);
                bt_tile_transfer<0, kGH>(bits_s + (size_t)jt * XP, nj_s + (size_t)jt * XP, jt, t_x,
                                         bt_tile_mask(jt, ntiles, t_y), lane);
            }
            if (dbg && lane == 0) dbg[14] = clock64();
        } else if (warp < w_act) {
            // ============================== DP warps ==============================
            // Nothing in the tile loop may branch (or predicate) on a loop-invariant condition: ptxas hoists
            // the compare out of the loop and parks it in one of the seven predicate registers the cells of
            // the tile body need (see opaque-free notes at sts_f32).  Role differences between warps are
            // therefore expressed as ADDRESSES: a warp without a predecessor polls a pre-satisfied flag,
            // a warp without a consumer publishes to a dump flag and stores its halo to a dump row.
            const int w = warp;
            const int lane_cta = 32 * w + lane;
            const int x0 = rows_base + lane_cta * R;                 // lane's first text position
            const int xw0 = rows_base + 32 * R * w;                  // warp's first text position
            const int lane_glob = x0 / R;
            const int lane7 = lane & 7;
            const uint32_t lane0_mask = (lane == 0) ? 0xffffffffu : 0u;
            const bool has_consumer = (w + 1 < w_act);               // boundary (w -> w+1) active
            const bool line_out = MULTIPASS && (w == W - 1) && !last_pass;        // feeds the next row pass
            const bool line_in = MULTIPASS && (w == 0) && (pass > 0);             // halo from the previous row pass
            const float *hb_in = (w > 0) ? hbuf + (size_t)(w - 1) * HS * NT : hconst;
            const int hin_step = (w > 0 || line_in) ? NT : 0;        // warp 0 of pass 0 re-reads one constant row
            float *hb_out = hbuf + (size_t)w * HS * NT;
            const uint32_t hout_base = ((has_consumer || line_out) && lane == 31) ? smem_u32(hb_out) : smem_u32(hdump + w * S::kDumpFloats + 4 * lane);
            const uint32_t hout_step = ((has_consumer || line_out) && lane == 31) ? NT * 4u : 0u;
            const int *flag_in = (w > 0) ? hprog + (w - 1) : hprog + W;           // hprog[W] is pre-satisfied
            int *flag_out = hprog + w;                                            // also read by the backtrack helper
            const float *gl_in = (MULTIPASS && pass > 0) ? gline_b + ((pass - 1) & 1) * P.line_pitch : nullptr;
            float *gl_out = line_out ? gline_b + (pass & 1) * P.line_pitch : nullptr;

            float q[R];
            uint32_t acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { q[r] = P.neg; acc[r] = 0u; }
            // neighbour value for frame 0: only text position 0 has a defined one (core.pyx:24-25, y == 0)
            float up = (x0 == 0) ? 0.f : P.neg;
            int known = 0;                 // last observed progress of warp w-1 (tiles completed)
            int stage = 0;
            uint32_t phase = 0;
            int hs = 0;                    // halo ring slot of the current tile
            bool tile_ready = false;       // ring_full of the current tile already observed
            if (MULTIPASS && line_in) {    // carried row of tile 0 into slot 0 of warp 0's input ring
                hconst[lane] = gl_in[lane];
                __syncwarp();
            }

            for (int j = 0; j < ntiles; ++j) {
                const int t0 = j * NT;
                if (!tile_ready) mbar_wait_warp(&ring_full[stage], phase);
                if (known < j + 1) known = flag_wait_ge_warp(flag_in, j + 1);
                const int next_stage = (stage + 1 == NS) ? 0 : stage + 1;
                const uint32_t next_phase = (stage + 1 == NS) ? (phase ^ 1) : phase;
                const int next_hs = (hs + 1 == HS) ? 0 : hs + 1;
                // early, non-blocking probe of the next tile's TMA barrier: its latency hides under the tile
                tile_ready = (j + 1 < ntiles) && mbar_test_warp(&ring_full[next_stage], next_phase);
                float hv_next = 0.f;
                if (MULTIPASS) { if (line_in && j + 1 < ntiles) hv_next = gl_in[t0 + NT + lane]; }

                const float *lane_tile = ring + (size_t)stage * kTileFloats + lane_cta * kTilePitch;
                const float *hin = hb_in + hs * hin_step;
                const uint32_t hout_addr = hout_base + hs * hout_step;
                // Tiles entirely below the diagonal (every row of the warp has x > y) are computed like any
                // other: their results are never consumed (the x == y cell substitutes max_neg_val) and the
                // warp would only be waiting for its predecessor anyway.
                const bool diag = (t0 < xw0 + 32 * R) && (t0 + NT - 1 >= xw0);
                const int dl0 = lane_glob - t0 / R;
                if (diag) dp_tile<R, XP, true>(q, acc, up, lane_tile, hin, lane7, lane0_mask, dl0, P.neg, hout_addr);
                else dp_tile<R, XP, false>(q, acc, up, lane_tile, hin, lane7, lane0_mask, dl0, P.neg, hout_addr);

                // ---- direction words of this tile, walk-ready: forced move of the diagonal cell (index == y,
                // core.pyx:34) OR-ed in, token 0 (never moves) cleared, bit-reversed (bit 31-k <-> frame k) ----
                uint32_t words[R];
                dp_finish_words<R, kCellExact>(acc, words, x0, j, diag);
                if (SMEM_BITS) store_words<R>(bits_s + (size_t)j * XP + lane_cta * R, words);
                else store_words<R>(gbits_b + (size_t)j * P.gbits_rows_pitch + x0, words);

                __syncwarp();                                   // lane 31's halo stores, everyone's ring reads
                if (elect_one()) {                              // a fresh predicate every tile, nothing to hoist
                    flag_release(flag_out, j + 1);
                    mbar_arrive(&ring_empty[stage]);
                }
                if (MULTIPASS) {
                    if (line_out) gl_out[t0 + lane] = hb_out[hs * NT + lane];
                    if (line_in && j + 1 < ntiles) {
                        hconst[next_hs * NT + lane] = hv_next;
                        __syncwarp();
                    }
                }
                stage = next_stage;
                phase = next_phase;
                hs = next_hs;
            }
        }
        if (dbg && tid == 0) { dbg[2] = clock64(); }   // warp 0 done with this pass
        __syncthreads();
    }
    if (dbg && tid == 0) dbg[4] = clock64();                               // all DP warps done

    // ================================ backtrack ================================
    // per-token / per-tile scratch: the idle ring when the bits have their own shared-memory region
    // (4*(Tx + tiles) bytes always fit below two value tiles there), else global (start table, carry line)
    int *tok = SMEM_BITS ? reinterpret_cast<int *>(ring) : start_b;
    int *xin = SMEM_BITS ? tok + XP : reinterpret_cast<int *>(gline_b);
    // per-tile start masks of the tokens, behind tok / xin
    uint32_t *sflags = reinterpret_cast<uint32_t *>(xin + ((ntiles + 3) & ~3));
    if constexpr (SMEM_BITS) {
        // transfer tables (upper row groups) the producer warp did not get to, tile entry tokens, start frames
        mas_backtrack_smem<XP, nthreads, kGH, kG>(bits_s, nj_s, xin, sflags, bt_state[3], ntiles, t_x, t_y, tid, nullptr);
    } else {
        const int rows_pitch = P.gbits_rows_pitch;
        uint32_t *stage_bits = reinterpret_cast<uint32_t *>(ring);           // the ring is idle now
        const int rows_cp = min(rows_pitch, ((t_x + 3) >> 2) << 2);
        const int chunk_tiles = max(1, (int)((S::ring_bytes(NS) / 4) / rows_cp));
        for (int jj = tid; jj < ntiles; jj += nthreads) xin[jj] = 0;
        __syncthreads();
        if (tid == 0) {
            bt_state[0] = t_x - 1;                                           // token
            bt_state[1] = ntiles - 1;                                        // tile
            bt_state[2] = (int)bt_tile_mask(ntiles - 1, ntiles, t_y);        // frames <= t_y-1 of the last tile
            bt_state[3] = 0;
            xin[ntiles - 1] = t_x - 1;
        }
        __syncthreads();
        int jhi = ntiles;
        while (true) {
            const int jlo = max(0, jhi - chunk_tiles);
            const int n4 = rows_cp >> 2;
            for (int i = tid; i < (jhi - jlo) * n4; i += nthreads) {
                const int jj = i / n4, cc = i - jj * n4;
                reinterpret_cast<uint4 *>(stage_bits + (size_t)jj * rows_cp)[cc] =
                    reinterpret_cast<const uint4 *>(gbits_b + (size_t)(jlo + jj) * rows_pitch)[cc];
            }
            __syncthreads();
            if (tid == 0) {
                int x = bt_state[0], j = bt_state[1];
                uint32_t mask = (uint32_t)bt_state[2];
                const bool done = backtrack_walk(stage_bits, rows_cp, jlo, x, j, mask, reinterpret_cast<uint32_t *>(tok), xin);
                bt_state[0] = x; bt_state[1] = j; bt_state[2] = (int)mask; bt_state[3] = done ? 1 : 0;
            }
            __syncthreads();
            if (bt_state[3]) break;
            jhi = jlo;
            __syncthreads();
        }
        // start frames, one thread per tile: tile j holds the starts of tokens (xin[j-1], xin[j]]
        for (int jj = tid; jj < ntiles; jj += nthreads) {
            const int hi = xin[jj], lo = jj > 0 ? xin[jj - 1] : 0;
            for (int x = hi; x > lo; --x) tok[x] = (jj << 5) + 32 - __ffs(tok[x]);
        }
        if (tid == 0) tok[0] = 0;
        __syncthreads();
    }

    if (dbg && tid == 0) dbg[5] = clock64();                               // backtrack done
    // ================================= outputs =================================
    // [start, duration] per token from the start frames; frame -> token index; dense path
    int *ft = P.frame_token ? P.frame_token + (size_t)b * P.Ty : nullptr;
    if (SMEM_BITS) {
        mas_emit_outputs_flags<nthreads>(P, b, tok, xin, sflags, ntiles, t_x, t_y, tid);
    } else {
        for (int x = tid; x < P.Tx; x += nthreads) {
            int s = 0, d = 0;
            if (x < t_x) {
                s = tok[x];
                d = ((x + 1 < t_x) ? tok[x + 1] : t_y) - s;
            }
            if (SMEM_BITS) start_b[x] = s;
            dur_b[x] = d;
            if (ft)
                for (int y = s; y < s + d; ++y) ft[y] = x;
        }
        if (ft)
            for (int y = t_y + tid; y < P.Ty; y += nthreads) ft[y] = -1;
        __syncthreads();
        if (!SMEM_BITS)                                                    // tok aliases start_b: zero the padding rows
            for (int x = t_x + tid; x < P.Tx; x += nthreads) start_b[x] = 0;
        __syncthreads();
    }
    write_path_any(P, b, start_b, dur_b, tid, nthreads);
    if (dbg && tid == 0) {
        dbg[6] = clock64(); dbg[7] = ((long long)t_x << 32) | (unsigned)t_y;
        long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[13] = t;
    }
}

}  // namespace masb200
