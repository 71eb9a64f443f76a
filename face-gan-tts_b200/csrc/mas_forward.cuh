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
//   * cells with x > y are computed as garbage and never consumed; the cell x == y
//     substitutes v_cur = max_neg_val exactly like the reference.
//
// CTA organisation (template R rows per lane, W DP warps; XP = 32*R*W rows):
//   producer warp W      a handful of TMA tensor-map loads (cp.async.bulk.tensor, box
//                        {32 frames, NB*R rows}, elementStrides {1,R}, 128B swizzle) per
//                        32-frame tile into an NS-deep shared-memory ring, completion
//                        counted on an mbarrier; the strided boxes land the rows in the
//                        lane-major permuted layout of mas_common.cuh so DP reads are
//                        conflict-free.  The ring depth keeps ~50-150 KB in flight per SM
//                        -- enough to stream a CTA's value matrix at HBM latency.  (A first
//                        version issued one 128-byte bulk copy per row: 6080 requests per
//                        utterance at ~30 ns each made the kernel request-rate bound.)
//   DP warps 0..W-1      lane l of warp w owns the R consecutive text positions
//                        x = rows_base + (32w + l)*R + r.  Per mel frame the lane
//                        updates its R rows top-down from registers; only row 0 needs a
//                        neighbour (lane-1's last row of the previous frame) via
//                        __shfl_up_sync, issued R-1 cell updates ahead of its use, so
//                        the shuffle latency is off the dependency chain.
//   warp skew            warp w trails warp w-1 by one 8-frame slab; the boundary row
//                        travels through a small shared-memory ring published with
//                        st.release / ld.acquire progress flags (no __syncthreads, no
//                        barrier instruction in the frame loop).  Only warp 0 waits on
//                        the TMA mbarrier; the others inherit the ordering through the
//                        flag chain.
//   direction bits       32 frames x 1 bit per row per tile, kept in shared memory when
//                        they fit (LRS2 shapes), otherwise in an L2-resident global
//                        scratch that is staged back through the idle ring.
//   backtrack            one thread walks TOKENS, not frames: for the current token it
//                        finds the frame where the path leaves it with one masked
//                        find-leading-one per 32-frame word (~t_x + t_y/32 dependent
//                        steps instead of t_y), emitting [start, duration] per token.
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
    int aligned;             // 1: every row segment is 16-byte aligned -> TMA bulk path
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
};

// One cell of the recurrence; the bit is set iff the diagonal predecessor wins.
// CELL 0: portable C.  CELL 1: inline PTX that keeps the select off the ALU pipe:
//   setp.gt p, v_prev, v_cur ; q = v_cur + v ; @p q = v_prev + v ; @p bits |= mask
// (1 ALU-pipe compare + 1 predicated logic op, both adds on the FMA pipe; the
// dependency chain per frame is compare -> predicated add).  Both are the same
// arithmetic: (v_prev > v_cur ? v_prev : v_cur) + v in fp32 RN, NaN -> v_cur.
template <int CELL>
__device__ __forceinline__ float mas_cell(float v_cur, float v_prev, float v, uint32_t &bits, int bitpos) {
    if constexpr (CELL == 0) {
        const bool d = v_prev > v_cur;   // core.pyx max(v_cur, v_prev) -> (v_prev > v_cur) ? v_prev : v_cur
        bits |= d ? (1u << bitpos) : 0u;
        return (d ? v_prev : v_cur) + v; // plain fp32 RN add, same association as core.pyx:30
    } else {
        float q;
        const uint32_t m = 1u << bitpos; // folds to an immediate after unrolling
        asm("{\n"
            " .reg .pred p;\n"
            " setp.gt.f32 p, %3, %2;\n"
            " add.rn.f32 %0, %2, %4;\n"
            " @p add.rn.f32 %0, %3, %4;\n"
            " @p or.b32 %1, %1, %5;\n"
            "}\n"
            : "=&f"(q), "+r"(bits)
            : "f"(v_cur), "f"(v_prev), "f"(v), "r"(m));
        return q;
    }
}

// 8 consecutive frames kb..kb+7 of one tile, fully unrolled.
//   q[r]    running previous-column Q of the lane's rows
//   acc[r]  direction word being assembled (bit k <-> frame t0+k)
//   h[i]    halo: Q of the row above the warp's first row at frame t0+kb+i-1 (same in all lanes)
//   src     lane the rotate-shuffle reads from: (lane + 31) & 31; lane 31 injects the halo
//   dl      lane_global - (t0+kb)/R  (DIAG only): the lane owns the diagonal cell of frame
//           t0+kb+i, in row r = i % R, exactly when dl == i / R
//   lbase   &stage[lane_cta * kTilePitch]; o0 = swizzled float offset of frames kb..kb+3 in the
//           lane's rows, ((kb >> 2) ^ (lane & 7)) << 2; frames kb+4..kb+7 sit at o0 ^ 4
template <int R, int XP, int CELL, bool DIAG, bool HALO_OUT>
__device__ __forceinline__ void dp_frames8(float (&q)[R], uint32_t (&acc)[R], const float *lbase, int kb, int o0,
                                           const float (&h)[8], int lane, int src, int dl, float neg,
                                           float *halo_out) {
    constexpr int kRowStride = (XP / R) * kTilePitch;      // floats between the lane's consecutive rows
    uint32_t a8[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a8[r] = 0u;
    float4 v4[2][R];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int r = 0; r < R; ++r)
            v4[hh][r] = *reinterpret_cast<const float4 *>(lbase + r * kRowStride + (hh == 0 ? o0 : (o0 ^ 4)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // one shuffle per frame: lane l reads lane l-1's last row; lane 0 reads lane 31, which sends the halo
        const float send = (lane == 31) ? h[i] : q[R - 1];
        const float up = __shfl_sync(kFullMask, send, src);
        float n[R];
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            const float4 vv = v4[i >> 2][r];
            const float v = (i & 3) == 0 ? vv.x : (i & 3) == 1 ? vv.y : (i & 3) == 2 ? vv.z : vv.w;
            float v_cur = q[r];
            if (DIAG && r == (i % R)) {                    // (t0+kb+i) % R == i % R since R | 8 | (t0+kb)
                if (dl == i / R) v_cur = neg;              // x == y  (core.pyx:19-20)
            }
            const float v_prev = (r == 0) ? up : q[r - 1];
            n[r] = mas_cell<CELL>(v_cur, v_prev, v, a8[r], i);
        }
        if (HALO_OUT && lane == 31) halo_out[i] = n[R - 1];
#pragma unroll
        for (int r = 0; r < R; ++r) q[r] = n[r];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] |= a8[r] << kb;
}

// Generic single frame (partial last slab): every row checks x >= t.
template <int R, int XP, bool HALO_OUT>
__device__ __forceinline__ void dp_frame_generic(float (&q)[R], uint32_t (&acc)[R], const float *lane_row0, int k,
                                                 float hk, int lane, int src, int x0, int t, float neg,
                                                 float *halo_out_k) {
    constexpr int kRowStride = (XP / R) * kTilePitch;
    const float send = (lane == 31) ? hk : q[R - 1];
    const float up = __shfl_sync(kFullMask, send, src);
    float n[R];
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        const float v = lane_row0[r * kRowStride + ((((k >> 2) ^ (lane & 7)) << 2) | (k & 3))];
        const float v_cur = (x0 + r >= t) ? neg : q[r];
        const float v_prev = (r == 0) ? up : q[r - 1];
        n[r] = mas_cell<0>(v_cur, v_prev, v, acc[r], k);
    }
    if (HALO_OUT && lane == 31) *halo_out_k = n[R - 1];
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] = n[r];
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
template <int R, int W>
struct MasSmem {
    static constexpr int XP = 32 * R * W;
    static constexpr int kTileFloats = XP * kTilePitch;
    __host__ __device__ static constexpr size_t ring_bytes(int ns) { return sizeof(float) * (size_t)ns * kTileFloats; }
    __host__ __device__ static constexpr int halo_slabs(int ns) { return 4 * (ns + 1); }
    __host__ __device__ static constexpr size_t halo_bytes(int ns) { return sizeof(float) * (size_t)W * halo_slabs(ns) * 8; }
    __host__ __device__ static constexpr size_t ctrl_bytes(int ns) { return 8 * (size_t)(2 * ns) + 4 * (size_t)W + 64; }
    __host__ __device__ static constexpr size_t fixed_bytes(int ns) {
        return ((ring_bytes(ns) + halo_bytes(ns) + ctrl_bytes(ns) + 127) / 128) * 128;
    }
    static size_t bits_bytes(int ntiles_max) { return sizeof(uint32_t) * (size_t)ntiles_max * XP; }
};

template <typename T> __device__ __forceinline__ uint32_t one_bits();
template <> __device__ __forceinline__ uint32_t one_bits<float>() { return 0x3f800000u; }
template <> __device__ __forceinline__ uint32_t one_bits<int>() { return 1u; }

// Dense path rows of one utterance from the [start, dur] table (4-byte elements).
template <typename T>
__device__ __forceinline__ void write_path_rows(T *path_b, const int *start_b, const int *dur_b, int Tx, int Ty,
                                                int tid, int nthreads) {
    const uint32_t one = one_bits<T>();
    uint32_t *out = reinterpret_cast<uint32_t *>(path_b);
    if ((Ty & 3) == 0 && (reinterpret_cast<uintptr_t>(path_b) & 15) == 0) {
        const int Ty4 = Ty >> 2;
        const int total = Tx * Ty4;
        for (int i = tid; i < total; i += nthreads) {
            const int x = i / Ty4;
            const int y = (i - x * Ty4) << 2;
            const int s = start_b[x];
            const int e = s + dur_b[x];                  // exclusive; dur == 0 -> empty
            uint4 o;
            o.x = (y + 0 >= s && y + 0 < e) ? one : 0u;
            o.y = (y + 1 >= s && y + 1 < e) ? one : 0u;
            o.z = (y + 2 >= s && y + 2 < e) ? one : 0u;
            o.w = (y + 3 >= s && y + 3 < e) ? one : 0u;
            __stcs(reinterpret_cast<uint4 *>(out) + i, o);
        }
    } else {
        const long long total = (long long)Tx * Ty;
        for (long long i = tid; i < total; i += nthreads) {
            const int x = (int)(i / Ty);
            const int y = (int)(i - (long long)x * Ty);
            const int s = start_b[x];
            const int e = s + dur_b[x];
            out[i] = (y >= s && y < e) ? one : 0u;
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

template <int R, int W, bool SMEM_BITS, int CELL>
__global__ void __launch_bounds__((W + 1) * 32, 1)
mas_forward_kernel(const MasParams P, const __grid_constant__ CUtensorMap tmap) {
    using S = MasSmem<R, W>;
    constexpr int XP = S::XP;
    constexpr int NT = kTileFrames;
    constexpr int kTileFloats = S::kTileFloats;
    constexpr int nthreads = (W + 1) * 32;

    const int NS = P.ring_stages;
    const int HS = S::halo_slabs(NS);

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    float *hbuf = reinterpret_cast<float *>(smem_raw + S::ring_bytes(NS));                 // [W][HS*8]
    uint64_t *ring_full = reinterpret_cast<uint64_t *>(smem_raw + S::ring_bytes(NS) + S::halo_bytes(NS));
    uint64_t *ring_empty = ring_full + NS;
    int *hprog = reinterpret_cast<int *>(ring_empty + NS);                                // [W]
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem_raw + S::fixed_bytes(NS));        // SMEM_BITS only
    __shared__ int bt_state[4];

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const int t_x = P.t_x[b];
    const int t_y = P.t_y[b];
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
    const int npass = (t_x + XP - 1) / XP;
    long long *dbg = P.dbg ? P.dbg + (size_t)b * 8 : nullptr;
    long long dbg_wait = 0;
    if (dbg && tid == 0) dbg[0] = clock64();
    uint32_t *gbits_b = SMEM_BITS ? nullptr : P.gbits + (size_t)b * P.gbits_stride_b;
    float *gline_b = P.gline ? P.gline + (size_t)b * 2 * P.line_pitch : nullptr;
    const float *vb = P.value + (size_t)b * P.stride_b;

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
            mbar_fence_init();
        }
        __syncthreads();

        if (warp == W) {
            // ============================ producer warp ============================
            int stage = 0;
            uint32_t phase = 0;
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
            }
        } else if (warp < w_act) {
            // ============================== DP warps ==============================
            const int w = warp;
            const int lane_cta = 32 * w + lane;
            const int x0 = rows_base + lane_cta * R;                 // lane's first text position
            const int xw0 = rows_base + 32 * R * w;                  // warp's first text position
            const int lane_glob = x0 / R;
            const bool has_consumer = (w + 1 < w_act);               // boundary (w -> w+1) active
            const bool line_out = (w == W - 1) && !last_pass;        // feeds the next row pass
            const bool halo_out = has_consumer || line_out;
            float *hb_in = hbuf + (size_t)(w > 0 ? w - 1 : 0) * HS * 8;
            float *hb_out = hbuf + (size_t)w * HS * 8;
            const int *flag_in = hprog + (w > 0 ? w - 1 : 0);
            int *flag_out = hprog + w;
            const float *gl_in = (pass > 0) ? gline_b + ((pass - 1) & 1) * P.line_pitch : nullptr;
            float *gl_out = line_out ? gline_b + (pass & 1) * P.line_pitch : nullptr;

            float q[R];
            uint32_t acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { q[r] = P.neg; acc[r] = 0u; }
            float carry = P.neg;                                     // halo value of the frame before the slab/tile
            int known = 0;                                           // last observed producer progress (slabs)
            int stage = 0;
            uint32_t phase = 0;
            int hs = 0;                                              // halo ring slot of the current slab (c % HS)
            const int src = (lane + 31) & 31;                        // rotate-shuffle source lane

            for (int j = 0; j < ntiles; ++j) {
                const int t0 = j * NT;
                const int kmax = min(NT, t_y - t0);
                float hv_tile = P.neg;                                // w == 0: halo of frame t0+lane-1
                if (w == 0) {
                    if (pass == 0) {
                        hv_tile = (t0 + lane == 0) ? 0.f : P.neg;    // core.pyx:23-27 (x == 0)
                    } else {
                        hv_tile = (lane == 0) ? carry : gl_in[t0 + lane - 1];
                        carry = gl_in[t0 + 31];
                    }
                    if (dbg) {
                        const long long c0 = clock64();
                        mbar_wait(&ring_full[stage], phase);
                        dbg_wait += clock64() - c0;
                        if (j == 0 && lane == 0) dbg[1] = clock64();
                    } else {
                        mbar_wait(&ring_full[stage], phase);
                    }
                }
                const float *st_lane = ring + (size_t)stage * kTileFloats + lane_cta * kTilePitch;

                for (int kb = 0; kb < kmax; kb += 8) {
                    const int c = 4 * j + (kb >> 3);                 // slab counter within the pass
                    // halo registers: h[i] = Q[row above the warp, frame t0+kb+i-1], identical in all lanes
                    float h[8];
                    if (w > 0) {
                        if (known < c + 1) known = flag_wait_ge(flag_in, c + 1);
                        const float4 s0 = *reinterpret_cast<const float4 *>(hb_in + hs * 8);
                        const float4 s1 = *reinterpret_cast<const float4 *>(hb_in + hs * 8 + 4);
                        h[0] = carry; h[1] = s0.x; h[2] = s0.y; h[3] = s0.z;
                        h[4] = s0.w;  h[5] = s1.x; h[6] = s1.y; h[7] = s1.z;
                        carry = s1.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) h[i] = __shfl_sync(kFullMask, hv_tile, kb + i);
                    }
                    float *hout = hb_out + hs * 8;
                    const int tb = t0 + kb;
                    const int o0 = ((kb >> 2) ^ (lane & 7)) << 2;
                    const bool below_diag = (tb + 7 < xw0);           // every row of the warp has x > y
                    if (!below_diag) {
                        if (kb + 8 <= kmax) {
                            const bool diag = (tb < xw0 + 32 * R);
                            const int dl = lane_glob - tb / R;
                            if (halo_out) {
                                if (diag) dp_frames8<R, XP, CELL, true, true>(q, acc, st_lane, kb, o0, h, lane, src, dl, P.neg, hout);
                                else dp_frames8<R, XP, CELL, false, true>(q, acc, st_lane, kb, o0, h, lane, src, dl, P.neg, hout);
                            } else {
                                if (diag) dp_frames8<R, XP, CELL, true, false>(q, acc, st_lane, kb, o0, h, lane, src, dl, P.neg, hout);
                                else dp_frames8<R, XP, CELL, false, false>(q, acc, st_lane, kb, o0, h, lane, src, dl, P.neg, hout);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int k = kb + i;
                                if (k < kmax) {
                                    if (halo_out) dp_frame_generic<R, XP, true>(q, acc, st_lane, k, h[i], lane, src, x0, t0 + k, P.neg, hout + i);
                                    else dp_frame_generic<R, XP, false>(q, acc, st_lane, k, h[i], lane, src, x0, t0 + k, P.neg, hout);
                                }
                            }
                        }
                    }
                    if (has_consumer) {
                        __syncwarp();
                        if (lane == 0) flag_release(flag_out, c + 1);
                    }
                    if (++hs == HS) hs = 0;
                }
                // ---- direction words of this tile ----
                if (SMEM_BITS) store_words<R>(bits_s + (size_t)j * XP + lane_cta * R, acc);
                else store_words<R>(gbits_b + (size_t)j * P.gbits_rows_pitch + x0, acc);
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0u;

                __syncwarp();
                if (lane == 0) mbar_arrive(&ring_empty[stage]);
                if (line_out) {
                    // the tile's 4 slabs are contiguous in the halo ring (HS is a multiple of 4)
                    const int hs0 = (4 * j) % HS;
                    gl_out[t0 + lane] = hb_out[hs0 * 8 + lane];
                    __syncwarp();
                }
                if (++stage == NS) { stage = 0; phase ^= 1; }
            }
        }
        if (dbg && tid == 0) { dbg[2] = clock64(); dbg[3] = dbg_wait; }   // warp 0 done with this pass
        __syncthreads();
    }
    if (dbg && tid == 0) dbg[4] = clock64();                               // all DP warps done

    // ================================ backtrack ================================
    // Token walk.  State (x, y_end, y): token x owns frames (y, y_end] so far; find the
    // highest frame y' <= y where the path leaves the token: d[x,y'] set, or y' == x.
    const int rows_pitch = SMEM_BITS ? XP : P.gbits_rows_pitch;
    uint32_t *stage_bits = reinterpret_cast<uint32_t *>(ring);           // the ring is idle now
    const int rows_cp = min(rows_pitch, ((t_x + 3) >> 2) << 2);
    const int chunk_tiles = SMEM_BITS ? ntiles : max(1, (int)((S::ring_bytes(NS) / 4) / rows_cp));
    if (tid == 0) { bt_state[0] = t_x - 1; bt_state[1] = t_y - 1; bt_state[2] = t_y - 1; bt_state[3] = 0; }
    __syncthreads();
    int jhi = ntiles;
    while (true) {
        const int jlo = max(0, jhi - chunk_tiles);
        if (!SMEM_BITS) {
            const int n4 = rows_cp >> 2;
            for (int i = tid; i < (jhi - jlo) * n4; i += nthreads) {
                const int jj = i / n4, cc = i - jj * n4;
                reinterpret_cast<uint4 *>(stage_bits + (size_t)jj * rows_cp)[cc] =
                    reinterpret_cast<const uint4 *>(gbits_b + (size_t)(jlo + jj) * rows_pitch)[cc];
            }
            __syncthreads();
        }
        if (tid == 0) {
            const uint32_t *wb = SMEM_BITS ? bits_s : stage_bits;
            const int wpitch = SMEM_BITS ? XP : rows_cp;
            int x = bt_state[0], y_end = bt_state[1], y = bt_state[2];
            bool done = false;
            while (true) {
                if (x == 0) { start_b[0] = 0; dur_b[0] = y_end + 1; done = true; break; }
                const int j = y >> 5;
                if (j < jlo) break;                                        // need the next chunk
                uint32_t wd = wb[(size_t)(j - jlo) * wpitch + x] & (0xffffffffu >> (31 - (y & 31)));
                if ((x >> 5) == j) wd |= 1u << (x & 31);                   // index == y forces the move (core.pyx:34)
                if (wd == 0u) { y = (j << 5) - 1; continue; }
                const int ys = (j << 5) + (31 - __clz(wd));
                start_b[x] = ys;
                dur_b[x] = y_end - ys + 1;
                --x;
                y = y_end = ys - 1;
            }
            bt_state[0] = x; bt_state[1] = y_end; bt_state[2] = y; bt_state[3] = done ? 1 : 0;
        }
        __syncthreads();
        if (bt_state[3]) break;
        jhi = jlo;
        __syncthreads();
    }

    if (dbg && tid == 0) dbg[5] = clock64();                               // backtrack done
    // ================================= outputs =================================
    for (int x = t_x + tid; x < P.Tx; x += nthreads) { start_b[x] = 0; dur_b[x] = 0; }
    if (P.frame_token) {
        int *ft = P.frame_token + (size_t)b * P.Ty;
        for (int x = tid; x < t_x; x += nthreads) {
            const int s = start_b[x], e = s + dur_b[x];
            for (int y = s; y < e; ++y) ft[y] = x;
        }
        for (int y = t_y + tid; y < P.Ty; y += nthreads) ft[y] = -1;
    }
    __syncthreads();
    write_path_any(P, b, start_b, dur_b, tid, nthreads);
    if (dbg && tid == 0) { dbg[6] = clock64(); dbg[7] = ((long long)t_x << 32) | (unsigned)t_y; }
}

}  // namespace masb200
