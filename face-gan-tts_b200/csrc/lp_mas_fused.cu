// lp_mas_fused.cu -- the fused log-prior + Monotonic Alignment Search kernel: ONE launch, one CTA -- or, for texts of
// 129..256 tokens in batches that fit the SMs twice, a 2-CTA cluster (PAIR: one 128-row M-tile and one DP warp per CTA,
// the halo row and the direction words cross with st.async) -- per utterance; the [Tx,Ty] value matrix never leaves the SM.
//
// Replaces reference model/face_tts.py:165-174 (log_prior -> maximum_path) for F = n_feats in {64, 80, 96, 128} and
// Tx <= 256 (the LRS2 training shapes).  The tcgen05 front end (TMA y tiles -> exact tf32 hi/lo split -> 3xTF32
// tcgen05.mma with A = mu_x parked in TENSOR MEMORY, lp_tc_frontend.cuh) produces the value matrix one 32-frame
// tile at a time; its epilogue writes every tile STRAIGHT INTO THE SHARED-MEMORY RING the alignment search reads
// (mas_forward.cuh: register-resident column, SHFL halo, 1 direction bit per cell in shared memory, per-token
// backtrack).  The y^2 / mu^2 / constant terms ride along in the contraction as one extra K step,
//   A_extra[x] = (1, 1, mc_hi, mc_lo, 0, 0, 0, 0),  B_extra[t] = (ysq_hi, ysq_lo, 1, 1, 0, 0, 0, 0),   mc = musq[x] + const,
// (exact tf32 pairs again), so the accumulator leaves TMEM as the finished log-prior value and the epilogue is a pure
// TMEM -> shared-memory copy: it shares its scheduler partition with a DP warp and must stay out of its way.
// The path is the bit-exact MAS of exactly these values (tests dump them through the fused_dump_ptr option); the search
// runs the predicate-free cell (mas_forward.cuh: kCellSign), identical to the reference's for the finite values it sees.
//
// Warp roles (16 warps; warp % 4 = scheduler partition, the arbiter favours the higher warp id):
//    0..3   epilogue: per tile TMEM -> registers -> ring (in the prologue: musq + const)
//    6, 7   operand split: raw y tile [F][32] -> hi/lo K-major core matrices + ysq
//    10     TMA loads (mu_x block, y tiles) + tcgen05.mma issue
//    11, 14, 15 backtrack helpers: per-tile transfer tables in the shadow of the DP (an even share of the row groups each)
//    12, 13 DP warps (text rows 0..127 / 128..255): alone with one epilogue warp on partitions 0 / 1, highest ids there
//           (PAIR: warp 12 is the CTA's DP warp; warp 13 of rank 0 forwards the halo row to the peer CTA)
//    9      dense path (when requested): streams the all-zero [Tx,Ty] block out with bulk copies from an 8 KB zero
//           buffer while the search runs -- the result only adds ~t_y ones to it (written by all warps in the tail)
//    4, 5, 8  parked until the tail (they only keep the latency-critical DP warps' partitions quiet)
//    prologue: warps 4, 5, 10, 11 / 8, 9, 14, 15 (one per TMEM lane quadrant each) park the lower / upper half of the mel
//    bins of mu_x in tensor memory while the split warps already work on the first y tiles
// TMEM lane m of M-tile mt holds text row x = 128*mt + 4*(m & 31) + (m >> 5): the epilogue thread of that lane then owns
// ring slot (m >> 5) * 32W + 32*mt + (m & 31) -- the lane-major permuted layout of mas_common.cuh -- so its eight 16-byte
// stores are bank-conflict free, M-tile mt is exactly DP warp mt's rows, and each (stage, M-tile) has its own
// full / empty mbarrier pair.
//
// Pipeline per 32-frame tile g:  TMA raw[g&1] -> split -> hi/lo[g&1] -> MMA -> D[g&1] (TMEM, two stages)
//                                -> epilogue -> ring[g % NS] -> DP -> direction words -> helpers.
#include <atomic>
#include <cstring>

#include <cudaTypedefs.h>

#include "lp_tc_frontend.cuh"
#include "mas_forward.cuh"
#include "mas_host.h"

namespace masb200 {

float log_prior_const(int F);                                                                        // log_prior_ffma.cu
int make_y_tensor_map(const float *y, int B, int F, int Ty, int box_frames, CUtensorMap *out);       // log_prior_tc.cu

namespace {

constexpr int kFR = 4;                       // text rows per DP lane
constexpr int kFusedCell = kCellSign;      // the search runs on the kernel's own finite log-prior values (mas_forward.cuh: mas_cell)
constexpr int kFusedWarps = 16;
constexpr int kFusedThreads = kFusedWarps * 32;
constexpr int kWarpSplit = 6;                // 6, 7
constexpr int kWarpMma = 10;
constexpr int kWarpHelpA = 11;
constexpr int kWarpDp = 12;                  // 12, 13
constexpr int kWarpHelpB = 14;
constexpr int kWarpHelpC = 15;
constexpr int kWarpZero = 9;                 // dense-path zero fill
constexpr uint32_t kZeroBytes = 8192;        // zero buffer = size of one bulk store

struct FusedParams {
    MasParams mas;       // t_x, t_y, B, Tx, Ty, neg, ring_stages, start, dur, frame_token, status, path, path_dtype, dbg
    const float *mu;     // [B,F,Tx]
    float cst;           // -0.5 * F * log(2 pi)
    float *value_dump;   // tests only: [B,Tx,Ty], receives the value tiles the search consumed (null normally)
    const float *y_pf;   // y [B,F,Ty]: only for the L2 prefetch hints ahead of the dependent-launch wait (the tiles come by TMA)
    int exp;             // diagnostics (MASB200_FUSED_PROF builds): see option fused_exp
};

// Shared-memory carve-up (bytes from a 1024-aligned base):
//   [ring: NS value tiles][halo rings][raw y: 2][hi: 2][lo: 2][ysq partials][mbarriers][flags][zero buffer][direction words + transfer tables]
// The [F][Tx] staging of mu_x for the prologue aliases the front (ring, halo, possibly raw).
// PAIR (2-CTA cluster per utterance, one M-tile per CTA): the home CTA (rank 0) keeps the direction words of BOTH
// M-tiles (XPT rows); behind them every CTA has the one-shot mbarriers of the cross-CTA hand-offs and the peer's halo row.
template <int KS, int W, bool PAIR = false>
struct FusedSmem {
    static constexpr int F = 8 * KS;
    static constexpr int XP = 32 * kFR * W;
    static constexpr int XPT = PAIR ? 2 * XP : XP;               // text rows of the utterance covered by the direction words
    using M = MasSmem<kFR, W>;
    static constexpr uint32_t kRaw = (uint32_t)F * 32u * 4u;      // one raw y tile [F][32 frames]
    static constexpr int FE = F + 8;                              // K extent of an operand tile: F mel bins + the extra K step
    static constexpr uint32_t kOp = (uint32_t)FE * 32u * 4u;      // one hi (or lo) operand tile
    __host__ __device__ static constexpr size_t off_halo(int ns) { return M::ring_bytes(ns); }
    __host__ __device__ static constexpr size_t off_raw(int ns) { return off_halo(ns) + M::halo_bytes(ns); }
    __host__ __device__ static constexpr size_t off_hi(int ns) { return off_raw(ns) + 2 * kRaw; }
    __host__ __device__ static constexpr size_t off_lo(int ns) { return off_hi(ns) + 2 * kOp; }
    __host__ __device__ static constexpr size_t off_part(int ns) { return off_lo(ns) + 2 * kOp; }             // [2][2][32] f32
    __host__ __device__ static constexpr size_t off_bars(int ns) { return off_part(ns) + 2 * 2 * 32 * 4; }    // 16 + 2*ns mbarriers (room for 4*ns)
    __host__ __device__ static constexpr size_t off_flags(int ns) { return off_bars(ns) + 8 * (size_t)(16 + 4 * ns); }
    __host__ __device__ static constexpr size_t off_zero(int ns) { return ((off_flags(ns) + 128 + 127) / 128) * 128; }
    __host__ __device__ static constexpr size_t off_bits(int ns) { return off_zero(ns) + kZeroBytes; }
    // direction words (4 B) + transfer table (1 B) per row and tile
    __host__ __device__ static constexpr size_t off_pair(int ns, int ntiles) { return ((off_bits(ns) + (size_t)5 * ntiles * XPT + 15) / 16) * 16; }
    // PAIR: [ntiles] halo mbarriers + [ntiles] direction-word mbarriers + [ntiles][32] halo floats
    __host__ __device__ static constexpr size_t total(int ns, int ntiles) {
        return off_pair(ns, ntiles) + (PAIR ? (size_t)ntiles * (8 + 8 + 128) + 128 : 0);     // + one slot the prefetch after the last tile reads
    }
};

// ---- 2-CTA cluster helpers (PAIR) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// the shared::cluster address of `p` (an address in THIS CTA's shared memory) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 16 bytes into a peer CTA's shared memory, counted on the peer's mbarrier when they have landed
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar) : "memory");
}

// in-loop wait / body cycle accumulators (scripts/fused_phases.py); off in production builds: the clock reads sit in
// the latency-critical loops
#ifndef MASB200_FUSED_PROF
#define MASB200_FUSED_PROF 0
#endif
#if MASB200_FUSED_PROF
#define PROF_EXP(bit) ((FP.exp >> (bit)) & 1)
#define PROF_DECL(...) long long __VA_ARGS__
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(acc, t0) acc += clock64() - (t0)
#else
#define PROF_EXP(bit) 0
#define PROF_DECL(...)
#define PROF_T(var)
#define PROF_ADD(acc, t0)
#endif

// {x0, x1} = {a0, a1} + {b0, b1}: one packed fp32x2 add (two RN adds, identical results to two add.rn.f32)
__device__ __forceinline__ void add_f32x2(float &x0, float &x1, float a0, float a1, float b0, float b1) {
    uint64_t a, b2, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d));
}

template <int KS, int W, bool PAIR>
__global__ void __launch_bounds__(kFusedThreads, 1)
lp_mas_fused_kernel(const FusedParams FP, const __grid_constant__ CUtensorMap ymap) {
    static_assert(!PAIR || W == 1, "a CTA of a pair holds one M-tile");
    using FS = FusedSmem<KS, W, PAIR>;
    constexpr int R = kFR;
    constexpr int F = 8 * KS;
    constexpr int XP = FS::XP;
    constexpr int XPT = FS::XPT;
    constexpr int NT = kTileFrames;
    constexpr int kTileFloats = XP * kTilePitch;
    constexpr uint32_t kSbo = (uint32_t)FS::FE * 32u;         // 8 frames x (F + 8) K values x 4 B per row group of an operand tile
    const MasParams &P = FP.mas;
    const int NS = P.ring_stages;
    const int HS = NS + 1;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    float *hbuf = reinterpret_cast<float *>(smem_raw + FS::off_halo(NS));                 // [W+1][HS][32] + [W][32]
    unsigned char *raw = smem_raw + FS::off_raw(NS);
    unsigned char *ophi = smem_raw + FS::off_hi(NS);
    unsigned char *oplo = smem_raw + FS::off_lo(NS);
    float *part = reinterpret_cast<float *>(smem_raw + FS::off_part(NS));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + FS::off_bars(NS));
    uint64_t *bar_aready = bars /*[M-tile]*/, *bar_xready = bars + 14 /*[M-tile]*/, *bar_raw = bars + 2, *bar_split = bars + 4, *bar_bfree = bars + 6;
    uint64_t *bar_dfull = bars + 8;              // [D stage][M-tile]
    uint64_t *bar_dempty = bars + 12;            // [D stage]
    uint64_t *ring_empty = bars + 16;            // [ring stage][M-tile]
    int *hprog = reinterpret_cast<int *>(smem_raw + FS::off_flags(NS));                   // [W+2] DP progress flags
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(hprog + 8);
    int *eprog = hprog + 12;                                                              // [W][4] epilogue progress flags
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem_raw + FS::off_bits(NS));
    unsigned char *zbuf = smem_raw + FS::off_zero(NS);
    // prologue staging of mu_x, transposed for the TMEM lane order: element (f, x) with x = 128*mt + 4*l + q at
    //   ((f*W + mt)*4 + q)*32 + (l ^ 8q)      -- coalesced global reads land conflict-free, and so do the readers
    float *mu_s = reinterpret_cast<float *>(smem_raw);

    // PAIR: CTA rank r of the cluster owns text rows [128 r, 128 r + 128) of utterance blockIdx.x / 2; rank 0 is the home
    // CTA (direction words of all rows, backtrack, outputs)
    const int rank = PAIR ? __shfl_sync(kFullMask, (int)cluster_ctarank(), 0) : 0;     // provably warp-uniform for ptxas
    const int row0 = PAIR ? 128 * rank : 0;
    const int b = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);     // provably warp-uniform for ptxas
    const int lane = tid & 31;
    // Programmatic dependent launch: this grid may have been placed while the previous kernel of the stream was still
    // running (its launch latency and the set-up below are hidden); the next kernel of the stream may be placed as soon
    // as every CTA of this grid is past the trigger.
    pdl_launch_dependents();
    // ---- set-up that touches no global memory, BEFORE the wait on the previous kernel of the stream (programmatic
    // dependent launch: this grid is usually resident while the previous one is still running)
    if (tid == 0) {
        mbar_init(&bar_aready[0], 8 * 32); mbar_init(&bar_aready[1], 8 * 32);
        mbar_init(&bar_xready[0], 4 * 32); mbar_init(&bar_xready[1], 4 * 32);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_raw[i], 1); mbar_init(&bar_split[i], 64); mbar_init(&bar_bfree[i], 1); mbar_init(&bar_dempty[i], 128);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&bar_dfull[i], 1);
        for (int i = 0; i < 2 * NS; ++i) mbar_init(&ring_empty[i], 1);
        for (int i = 0; i < W; ++i) hprog[i] = 0;
        hprog[W] = 0x7fffffff;                               // the flag a lane without anything to wait for polls
        hprog[W + 1] = 0;
        for (int i = 0; i < 4 * W; ++i) eprog[i] = 0;
        mbar_fence_init();
    }
    if (P.path != nullptr) {
        for (uint32_t i = tid; i < kZeroBytes / 16; i += kFusedThreads) reinterpret_cast<uint4 *>(zbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();                            // generic-proxy zeros -> visible to the bulk stores reading them
    }
    // ... and L2 prefetches of what the prologue is about to read (hints, not accesses: L2 is the point of coherence, a
    // line the previous kernel still writes is simply updated there): the lengths, this CTA's mu_x block, the first two y
    // tiles, the tensor map.  After the wait the prologue's loads are L2 hits instead of HBM round trips.
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ymap)) : "memory");
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.t_x + b));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.t_y + b));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(FP.mu + (size_t)b * F * P.Tx), "r"((uint32_t)(F * P.Tx * 4)) : "memory");
    }
    if (tid >= 32 && tid < 32 + 2 * F && 64 <= P.Ty)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(FP.y_pf + ((size_t)b * F + ((tid - 32) >> 1)) * P.Ty + 32 * ((tid - 32) & 1)));
    // PAIR: the one-shot hand-off barriers -- one per tile the PADDED length allows, so that nothing here depends on the
    // lengths -- are armed for the bytes the peer will send (rank 1 receives the halo row of every tile: 32 floats; rank 0
    // rank 1's direction words: 128 rows x 4 bytes), and the peer is told (cluster barrier, arrive half; the matching wait
    // sits behind the prologue).  The cluster-scope fence + release cost ~1.2 k cycles when they sat behind the wait.
    // Barriers of tiles beyond t_y, or of an utterance whose second CTA has nothing to do, simply stay armed.
    const int ntmax = (P.Ty + NT - 1) / NT;
    uint64_t *bar_h = reinterpret_cast<uint64_t *>(smem_raw + FS::off_pair(NS, ntmax));      // PAIR [ntmax]: halo row of tile j has landed (rank 1)
    uint64_t *bar_b = bar_h + ntmax;                                                           // PAIR [ntmax]: rank 1's direction words of tile j have landed (rank 0)
    float *halo_full = reinterpret_cast<float *>(bar_b + ntmax);                               // PAIR [ntmax][32]: Q of text row 127 (rank 1)
    if (PAIR) {
        if (tid >= 32 && tid < 32 + ntmax) {
            uint64_t *bar = (rank == 0 ? bar_b : bar_h) + (tid - 32);
            mbar_init(bar, 1);
            mbar_arrive_expect_tx(bar, rank == 0 ? 512u : 128u);
            mbar_fence_init();
        }
        cluster_arrive();
    }
    // nothing the previous kernel may have written is touched before this wait returns
    pdl_wait();
    long long *dbg = P.dbg ? P.dbg + ((size_t)rank * P.B + b) * 32 : nullptr;   // diagnostics ([2B][32] for a pair): phase stamps [0..15], wait cycles [16..31]
    if (dbg && tid == 0) { dbg[0] = clock64(); long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[12] = t; }

    // ---- mu_x block of this utterance: coalesced global loads issued first (they do not wait for the lengths: the
    // padding of a row is valid memory, masked once t_x is here), consumed after the set-up below.
    // One warp per mel-bin row f (f = warp, warp + 15, ...), lane + 32j the text position: no index arithmetic per load.
    constexpr int kMuRows = (F + kFusedWarps - 1) / kFusedWarps;
    constexpr int kMuCols = 4 * W;                            // 32-position column groups of a row
    float mu_reg[kMuRows][kMuCols];
    {
        const float *mu_b = FP.mu + (size_t)b * F * P.Tx + row0 + lane;
#pragma unroll
        for (int k = 0; k < kMuRows; ++k) {
            const int f = warp + kFusedWarps * k;
            const float *rp = mu_b + (size_t)f * P.Tx;
#pragma unroll
            for (int j = 0; j < kMuCols; ++j)
                mu_reg[k][j] = (f < F && row0 + lane + 32 * j < P.Tx) ? __ldg(rp + 32 * j) : 0.f;
        }
    }
    const int t_x = __shfl_sync(kFullMask, P.t_x[b], 0);
    const int t_y = __shfl_sync(kFullMask, P.t_y[b], 0);
    int *start_b = P.start + (size_t)b * P.Tx;
    int *dur_b = P.dur + (size_t)b * P.Tx;

    // ---- per-item validation (the reference is undefined here: core.pyx:34) ----
    if (t_x < 1 || t_y < 1 || t_x > P.Tx || t_y > P.Ty || t_x > t_y) {
        if (rank != 0) return;
        for (int x = tid; x < P.Tx; x += kFusedThreads) {
            start_b[x] = 0; dur_b[x] = 0;
            if (P.peer_dur != nullptr)
                for (int r = 0; r < P.peer_world; ++r) P.peer_dur[r][((size_t)P.peer_rank * P.B + b) * P.Tx + x] = 0;
        }
        if (P.frame_token)
            for (int y = tid; y < P.Ty; y += kFusedThreads) P.frame_token[(size_t)b * P.Ty + y] = -1;
        if (P.status && tid == 0) P.status[b] = MAS_B200_ITEM_BAD_LENGTH;
        __syncthreads();
        write_path_any(P, b, start_b, dur_b, tid, kFusedThreads);
        return;
    }
    if (P.status && tid == 0 && rank == 0) P.status[b] = MAS_B200_ITEM_OK;

    const int w_tot = (t_x + 32 * R - 1) / (32 * R);          // M-tiles of the utterance with valid rows
    // a pair's second CTA has nothing to do for a text of <= 128 tokens: it leaves (its half of the cluster barrier is
    // done: the arrive ahead of the launch wait; nobody waits for it) before it owns tensor memory.  (Both early exits are plain
    // returns on provably warp-uniform conditions; an exit that has to hand tensor memory back first -- a block barrier in
    // the exit path -- makes ptxas give up on the convergence of every warp behind it, and the MMA warp then issues through
    // the divergent elect fallback: ~90 instead of ~20 cycles per instruction.  Hence the allocation comes after them.)
    if (PAIR && rank >= w_tot) return;
    const int ntiles = (t_y + NT - 1) / NT;
    const int w_act = PAIR ? 1 : w_tot;                       // active DP warps == M-tiles of THIS CTA
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, kLpTmemCols); tmem_relinquish(); }
    const bool peer = PAIR && w_tot == 2;                     // the other CTA of the pair is at work too
    unsigned char *nj_s = reinterpret_cast<unsigned char *>(bits_s + (size_t)ntiles * XPT);  // [ntiles][XPT] transfer table
    // the prologue's mu_x staging reaches into the raw y buffers: y tiles (and everything behind them) start late
    const bool late_start = (size_t)F * W * 128 * 4 > FS::off_raw(NS);
#pragma unroll
    for (int k = 0; k < kMuRows; ++k)
#pragma unroll
        for (int j = 0; j < kMuCols; ++j)
            if (row0 + lane + 32 * j >= t_x) mu_reg[k][j] = 0.f;

    {
        // x = 32j + lane = 128*mt + 4*l + q with mt = j >> 2, l = 8*(j & 3) + (lane >> 2), q = lane & 3:
        //   position ((f*W + mt)*4 + q)*32 + (l ^ 8q) = f*W*128 + mt*128 + [q*32 + (lane >> 2) + ((8*(j & 3)) ^ 8q)]
        const int q = lane & 3;
        int e[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = q * 32 + (lane >> 2) + ((8 * c) ^ (8 * q));
#pragma unroll
        for (int k = 0; k < kMuRows; ++k) {
            const int f = warp + kFusedWarps * k;
            if (f < F) {
#pragma unroll
                for (int j = 0; j < kMuCols; ++j) mu_s[f * W * 128 + (j >> 2) * 128 + e[j & 3]] = mu_reg[k][j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(kFullMask, *tmem_slot, 0);
    if (dbg && tid == 0) dbg[8] = clock64();
    // TMEM columns: M-tile mt: A hi at mt*(2F+8), A lo at +F, the extra K step at +2F; D of (stage p, M-tile mt) behind them
    auto col_a = [](int mt, int part) { return (uint32_t)(mt * (2 * F + 8) + part * F); };
    auto col_d = [](int p, int mt) { return (uint32_t)(W * (2 * F + 8) + (p * W + mt) * 32); };
    static_assert(W * (2 * F + 8) + 2 * W * 32 <= kLpTmemCols, "A (hi, lo, extra) + two D stages must fit the 512 TMEM columns");

    if (warp == kWarpMma && !late_start) {
        // the first two y tiles are on their way while the A operand is being parked
        if (elect_one()) {
            for (int q = 0; q < 2 && q < ntiles; ++q) {
                mbar_arrive_expect_tx(&bar_raw[q], FS::kRaw);
                tma_load_3d(raw + (size_t)q * FS::kRaw, &ymap, q * NT, 0, b, &bar_raw[q]);
            }
        }
        __syncwarp();
    }

    // ---- prologue (warp % 4 = TMEM lane quadrant): mu_x rows -> exact tf32 hi/lo -> TMEM, the A operand for the whole
    // CTA.  Warps 4, 5, 10, 11 take the lower half of the mel bins, warps 8, 9, 14, 15 the upper half; warps 0..3 (the
    // epilogue warps) compute musq[x] = -0.5 sum_f mu^2.  The split warps (6, 7) are NOT among them: when the staging
    // does not reach into the y buffers they split the first y tiles meanwhile.
    const int park_grp = warp < 4 ? 0 : ((warp == 4 || warp == 5 || warp == 10 || warp == 11) ? 1
                                         : ((warp == 8 || warp == 9 || warp == 14 || warp == 15) ? 2 : -1));
    if (park_grp >= 0) {
        const int q = warp & 3, grp = park_grp;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
#pragma unroll
        for (int mt = 0; mt < W; ++mt) {
            if (mt < w_act) {
                const float *src = mu_s + (mt * 4 + q) * 32 + (lane ^ (8 * q));       // + f * W * 128
                if (grp == 0) {
                    // musq = -0.5 sum_f mu^2 over four interleaved partial sums (a single chain of F dependent FMAs
                    // was the longest thing in the prologue)
                    float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int f = 0; f < F; f += 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float v = src[(f + k) * W * 128];
                            sq[k] = fmaf(-0.5f * v, v, sq[k]);
                        }
                    }
                    // the extra K step of this row: (1, 1, mc_hi, mc_lo, 0, 0, 0, 0),  mc = musq + const
                    uint32_t ex[8] = {0x3f800000u, 0x3f800000u, 0u, 0u, 0u, 0u, 0u, 0u};
                    tf32_split(((sq[0] + sq[1]) + (sq[2] + sq[3])) + FP.cst, ex[2], ex[3]);
                    tmem_st8(tmem + lane_base + col_a(mt, 2), ex);
                } else {
                    constexpr int KH = KS / 2;
                    const int fbase = (grp - 1) * KH * 8;
#pragma unroll 1
                    for (int c = 0; c < KH; ++c) {
                        const int f0 = fbase + 8 * c;
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) tf32_split(src[(f0 + k) * W * 128], hi[k], lo[k]);
                        tmem_st8(tmem + lane_base + col_a(mt, 0) + f0, hi);
                        tmem_st8(tmem + lane_base + col_a(mt, 1) + f0, lo);
                    }
                }
            }
            // M-tile by M-tile: the first tile's MMAs on M-tile 0 start while M-tile 1 is still being parked
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(grp == 0 ? &bar_xready[mt] : &bar_aready[mt]);
        }
    }
    if (dbg && tid == 0) dbg[9] = clock64();
    if (PAIR && peer) cluster_wait();

    // diagnostics (fused_exp): what slows the DP warps down?  Roles can be parked; the DP warps then run over whatever is in the ring
    const bool is_help = warp == kWarpHelpA || warp == kWarpHelpB || warp == kWarpHelpC;
    const bool is_dp = warp >= kWarpDp && warp < kWarpDp + W;
    const int rw = ((PROF_EXP(1) && is_help) || (PROF_EXP(2) && !is_help && !is_dp)) ? 99 : warp;
    if (rw == kWarpMma) {
        // ======================= TMA loads + MMA issue (warp-uniform; one elected lane acts) =======================
        if (late_start) {
            mbar_wait_warp(&bar_aready[W - 1], 0);
            mbar_wait_warp(&bar_xready[W - 1], 0);
            if (elect_one()) {
                for (int q = 0; q < 2 && q < ntiles; ++q) {
                    mbar_arrive_expect_tx(&bar_raw[q], FS::kRaw);
                    tma_load_3d(raw + (size_t)q * FS::kRaw, &ymap, q * NT, 0, b, &bar_raw[q]);
                }
            }
            __syncwarp();
        }
        const uint32_t idesc = umma_idesc_tf32_ts(128, NT);
        PROF_DECL(w_split = 0, w_dempty = 0);
        for (int g = 0; g < ntiles; ++g) {
            const int p = g & 1;
            const uint32_t par = (uint32_t)(g >> 1) & 1u;
            PROF_T(c0);
            mbar_wait_warp(&bar_split[p], par);
            PROF_ADD(w_split, c0);
            // every split thread is done with raw buffer p: fetch tile g + 2 into it
            if (g + 2 < ntiles && elect_one()) {
                mbar_arrive_expect_tx(&bar_raw[p], FS::kRaw);
                tma_load_3d(raw + (size_t)p * FS::kRaw, &ymap, (g + 2) * NT, 0, b, &bar_raw[p]);
            }
            __syncwarp();
            PROF_T(c1);
            if (g >= 2) mbar_wait_warp(&bar_dempty[p], par ^ 1u);           // epilogue drained D stage p (tile g - 2)
            PROF_ADD(w_dempty, c1);
            tc_fence_after();
            const uint32_t bh = smem_u32(ophi) + (uint32_t)p * FS::kOp;
            const uint32_t bl = smem_u32(oplo) + (uint32_t)p * FS::kOp;
#pragma unroll
            for (int mt = 0; mt < W; ++mt) {
                if (mt < w_act) {
                    if (g == 0) {                                   // A of this M-tile is parked
                        mbar_wait_warp(&bar_aready[mt], 0);
                        tc_fence_after();
                        if (dbg && lane == 0 && mt == 0) dbg[1] = clock64();
                    }
                    const uint32_t dcol = tmem + col_d(p, mt);
                    const uint32_t ah = tmem + col_a(mt, 0), al = tmem + col_a(mt, 1);
                    // a compact loop (two K steps per iteration), not 6*KS unrolled instructions: the code of all
                    // roles has to share the instruction cache with the latency-critical DP warps
                    const uint64_t dh0 = umma_smem_desc_k_nosw(bh, 128u, kSbo);
                    const uint64_t dl0 = umma_smem_desc_k_nosw(bl, 128u, kSbo);
#pragma unroll 2
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint64_t dh = dh0 + (uint64_t)(ks * 16);          // + ks * 256 bytes (address field, >> 4)
                        const uint64_t dl = dl0 + (uint64_t)(ks * 16);
                        umma_tf32_ts_elect(dcol, ah + 8u * ks, dh, idesc, ks > 0 ? 1u : 0u);     // hi * hi
                        umma_tf32_ts_elect(dcol, ah + 8u * ks, dl, idesc, 1u);                    // hi * lo
                        umma_tf32_ts_elect(dcol, al + 8u * ks, dh, idesc, 1u);                    // lo * hi
                    }
                    // + ysq[t] + (musq[x] + const): the extra K step (operand chunks 2KS, 2KS + 1 of the hi tile)
                    if (g == 0) { mbar_wait_warp(&bar_xready[mt], 0); tc_fence_after(); }
                    umma_tf32_ts_elect(dcol, tmem + col_a(mt, 2), dh0 + (uint64_t)(KS * 16), idesc, 1u);
                    umma_commit_elect(&bar_dfull[p * 2 + mt]);     // D of (tile g, M-tile mt) complete -> epilogue warps
                }
            }
            umma_commit_elect(&bar_bfree[p]);                      // operand buffer p may be refilled -> split warps
            __syncwarp();
        }
#if MASB200_FUSED_PROF
        if (dbg && lane == 0) { dbg[16] = w_split; dbg[17] = w_dempty; }
#endif
    } else if (rw == kWarpSplit || rw == kWarpSplit + 1) {
        // ======================= operand split: raw [F][32] -> hi/lo K-major core matrices + the ysq K step =======================
        // thread = (frame n = lane, mel-bin chunks kc = sw, sw + 2, ...): 4 conflict-free LDS.32 down a column of the raw
        // tile, one STS.128 per operand into core matrix (n / 8, kc), row n % 8.
        const int sw = warp - kWarpSplit;
        if (late_start) { mbar_wait(&bar_aready[W - 1], 0); mbar_wait(&bar_xready[W - 1], 0); }
        const int n = lane;
        const uint32_t row_off = (uint32_t)(n >> 3) * kSbo + (uint32_t)(n & 7) * 16u;
        PROF_DECL(w_bfree = 0, w_raw = 0);
        for (int g = 0; g < ntiles; ++g) {
            const int p = g & 1;
            const uint32_t par = (uint32_t)(g >> 1) & 1u;
            PROF_T(c0);
            if (g >= 2) mbar_wait(&bar_bfree[p], par ^ 1u);        // MMA(g - 2) has read operand buffer p
            PROF_ADD(w_bfree, c0);
            PROF_T(c1);
            mbar_wait(&bar_raw[p], par);
            PROF_ADD(w_raw, c1);
            const float *rw = reinterpret_cast<const float *>(raw + (size_t)p * FS::kRaw);
            unsigned char *hb = ophi + (size_t)p * FS::kOp, *lb = oplo + (size_t)p * FS::kOp;
            // two halves of KS/2 chunks; in each all loads go first (the compiler cannot move shared-memory loads
            // above the operand stores itself)
            constexpr int KH = KS / 2;
            float q = 0.f;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int kc0 = sw + 2 * KH * half;
                float v[KH][4];
#pragma unroll
                for (int i = 0; i < KH; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[i][k] = rw[(4 * (kc0 + 2 * i) + k) * NT + n];
#pragma unroll
                for (int i = 0; i < KH; ++i) {
                    const uint32_t kc = (uint32_t)(kc0 + 2 * i);
                    uint4 h, l;
                    tf32_split(v[i][0], h.x, l.x); tf32_split(v[i][1], h.y, l.y);
                    tf32_split(v[i][2], h.z, l.z); tf32_split(v[i][3], h.w, l.w);
                    *reinterpret_cast<uint4 *>(hb + row_off + kc * 128u) = h;
                    *reinterpret_cast<uint4 *>(lb + row_off + kc * 128u) = l;
#pragma unroll
                    for (int k = 0; k < 4; ++k) q = fmaf(-0.5f * v[i][k], v[i][k], q);
                }
            }
            float *pd = part + p * 64;
            pd[sw * 32 + n] = q;
            asm volatile("bar.sync 1, 64;" ::: "memory");
            // the extra K step of frame n: (ysq_hi, ysq_lo, 1, 1 | 0, 0, 0, 0), ysq = -0.5 sum_f y^2
            if (sw == 0) {
                uint4 e = make_uint4(0u, 0u, 0x3f800000u, 0x3f800000u);
                tf32_split(pd[n] + pd[32 + n], e.x, e.y);
                *reinterpret_cast<uint4 *>(hb + row_off + (uint32_t)(2 * KS) * 128u) = e;
            } else {
                *reinterpret_cast<uint4 *>(hb + row_off + (uint32_t)(2 * KS + 1) * 128u) = make_uint4(0u, 0u, 0u, 0u);
            }
            fence_proxy_async_smem();              // operand stores -> visible to the tensor core's smem reads
            mbar_arrive(&bar_split[p]);            // also: raw buffer p may be refilled
        }
#if MASB200_FUSED_PROF
        if (dbg && sw == 0 && lane == 0) { dbg[18] = w_bfree; dbg[19] = w_raw; }
#endif
    } else if (rw < 4) {
        // ======================= epilogue warps: TMEM lane quadrant = warp =======================
        // per tile: D (TMEM) = the finished log-prior values -> ring slot of this lane's text row
        const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
        int stage = 0;
        uint32_t sphase = 0;                                       // parity of the ring stage's CURRENT use
        PROF_DECL(w_dfull = 0, w_rempty = 0);
        for (int g = 0; g < ntiles; ++g) {
            const int p = g & 1;
            const uint32_t par = (uint32_t)(g >> 1) & 1u;
#pragma unroll
            for (int mt = 0; mt < W; ++mt) {
                if (mt < w_act) {
                    uint32_t d[32];
                    PROF_T(c0);
                    mbar_wait(&bar_dfull[p * 2 + mt], par);
                    PROF_ADD(w_dfull, c0);
                    tc_fence_after();
                    tmem_ld32(tmem + lane_base + col_d(p, mt), d);
                    tmem_wait_ld();
                    if (mt == w_act - 1) { tc_fence_before(); mbar_arrive(&bar_dempty[p]); }
                    PROF_T(c1);
                    if (g >= NS) mbar_wait(&ring_empty[stage * 2 + mt], sphase ^ 1u);     // DP warp mt released the stage
                    PROF_ADD(w_rempty, c1);
                    float *rowp = ring + (size_t)stage * kTileFloats + (size_t)(warp * (32 * W) + 32 * mt + lane) * kTilePitch;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4 *>(rowp + ((c ^ (lane & 7)) << 2)) = make_uint4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]);
                    if (FP.value_dump != nullptr) {                // tests only
                        const int x = row0 + 128 * mt + 4 * lane + warp;
                        float *dst = FP.value_dump + ((size_t)b * P.Tx + x) * P.Ty + g * NT;
                        if (x < P.Tx)
                            for (int c = 0; c < 8; ++c)
                                if (g * NT + 4 * c < P.Ty)
                                    *reinterpret_cast<uint4 *>(dst + 4 * c) = make_uint4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]);
                    }
                    __syncwarp();
                    if (lane == 0) flag_release(&eprog[mt * 4 + warp], g + 1);      // this warp's rows of (tile g, M-tile mt) are in the ring
                }
            }
            if (++stage == NS) { stage = 0; sphase ^= 1u; }
        }
#if MASB200_FUSED_PROF
        if (dbg && tid == 0) { dbg[20] = w_dfull; dbg[21] = w_rempty; }
#endif
    } else if (rw == kWarpZero) {
        // ======================= dense path, part 1: the all-zero [Tx,Ty] block, in the shadow of the search =======================
        if (P.path != nullptr && rank == 0) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(P.path) + (size_t)b * P.Tx * P.Ty * 4;
            const size_t total = (size_t)P.Tx * P.Ty * 4;                      // Ty % 4 == 0: a multiple of 16 bytes
            if (elect_one()) {
                for (size_t off = 0; off < total; off += kZeroBytes) {
                    const uint32_t n = (uint32_t)(total - off < kZeroBytes ? total - off : kZeroBytes);
                    tma_bulk_store_1d(dst + off, zbuf, n);
                    tma_store_commit();
                }
                tma_store_wait_all();                                          // the zeros are in memory ...
                asm volatile("fence.proxy.async;" ::: "memory");               // ... and ordered before the generic-proxy ones
            }
            __syncwarp();
        }
    } else if ((rw == kWarpHelpA || rw == kWarpHelpB || rw == kWarpHelpC) && rank == 0) {
        // ======================= backtrack helpers: transfer tables behind the LAST active DP warp =======================
        // three warps share the row groups of every finished tile evenly (they have to keep up with the DP warps: what
        // they have not done when the search ends is on the critical path)
        const int *flag_last = hprog + (w_act - 1);
        const int h = warp == kWarpHelpA ? 0 : (warp == kWarpHelpB ? 1 : 2);
        int known = 0;
#ifdef MASB200_HELP_PROF
        long long hw = 0, hk = 0; int nwait = 0;
#endif
        for (int jt = 0; jt < ntiles; ++jt) {
#ifdef MASB200_HELP_PROF
            const long long c0 = clock64();
            if (known < jt + 1) { known = flag_wait_ge_warp(flag_last, jt + 1); ++nwait; }
            if (PAIR && peer) mbar_wait_warp(&bar_b[jt], 0);
            const long long c1 = clock64();
#else
            if (known < jt + 1) known = flag_wait_ge_warp(flag_last, jt + 1);
            if (PAIR && peer) mbar_wait_warp(&bar_b[jt], 0);             // the peer CTA's direction words of the tile are here
#endif
            bt_tile_transfer_share(bits_s + (size_t)jt * XPT, nj_s + (size_t)jt * XPT, jt, t_x, bt_tile_mask(jt, ntiles, t_y), lane, h, 3);
#ifdef MASB200_HELP_PROF
            hw += c1 - c0; hk += clock64() - c1;
#endif
        }
#ifdef MASB200_HELP_PROF
        if (dbg && lane == 0) { dbg[16 + 3 * h] = hw; dbg[17 + 3 * h] = hk; dbg[18 + 3 * h] = nwait; }
#endif
        if (dbg && lane == 0 && h < 2) dbg[28 + h] = clock64();
    } else if (PAIR && rw == kWarpDp + 1 && rank == 0 && peer) {
        // ======================= PAIR, rank 0: halo forwarding.  The DP warp leaves its bottom row (Q of text row 127) in
        // the CTA's own halo buffer like a DP warp with a local consumer would; this warp sends every finished tile's
        // 32 floats to the peer with st.async -- 16 bytes per lane, delivered into the peer's shared memory and counted on
        // its mbarrier of the tile.  Nothing of the hand-off is in the DP warp's instruction stream. =======================
        const uint32_t raddr = mapa_u32(halo_full, 1) + 16u * (lane & 7), rbar = mapa_u32(bar_h, 1);
        // (polling the DP warp's progress flag, like the helpers: measured faster for the DP warp than sleeping on an mbarrier)
        int known = 0;
        for (int jt = 0; jt < ntiles; ++jt) {
            if (known < jt + 1) known = flag_wait_ge_warp(hprog, jt + 1);
            if (dbg && jt == 0 && lane == 0) dbg[27] = clock64();
            if (lane < 8) {
                const uint4 v = *reinterpret_cast<const uint4 *>(halo_full + jt * NT + 4 * lane);
                st_async_v4(raddr + (uint32_t)jt * (NT * 4u), rbar + (uint32_t)jt * 8u, v.x, v.y, v.z, v.w);
            }
            __syncwarp();
        }
    } else if (rw >= kWarpDp && rw < kWarpDp + w_act) {
        // ======================= DP warps (see mas_forward_kernel: nothing in the tile loop may branch or predicate
        // on a loop-invariant condition; role differences are ADDRESSES) =======================
        const int w = warp - kWarpDp;
        const int lane_cta = 32 * w + lane;                      // lane among the CTA's DP lanes (ring rows)
        const int lane_utt = row0 / R + lane_cta;                // lane among the utterance's (direction words, diagonal)
        const int x0 = lane_utt * R;                             // lane's first text position
        const int xw0 = row0 + 32 * R * w;                       // warp's first text position
        const int lane7 = lane & 7;
        const uint32_t lane0_mask = (lane == 0) ? 0xffffffffu : 0u;
        const bool has_consumer = PAIR ? (rank == 0 && peer) : (w + 1 < w_act);
        float *hconst = hbuf + (size_t)W * HS * NT;              // warp 0's halo input (one constant row)
        float *hdump = hconst + (size_t)HS * NT;                 // [W][160] where lanes without a consumer store
        // halo input: the previous DP warp's ring; the constant row above text position 0; PAIR rank 1: the row the peer
        // CTA's DP warp sends, one slot per tile
        const bool halo_remote = PAIR && rank == 1;
        const float *hb_in = halo_remote ? halo_full : ((w > 0) ? hbuf + (size_t)(w - 1) * HS * NT : hconst);
        const int hin_step = (halo_remote || w > 0) ? NT : 0;
        float *hb_out = PAIR ? halo_full : hbuf + (size_t)w * HS * NT;
        const uint32_t hout_base = (has_consumer && lane == 31) ? smem_u32(hb_out) : smem_u32(hdump + w * FS::M::kDumpFloats + 4 * lane);
        const uint32_t hout_step = (has_consumer && lane == 31) ? NT * 4u : 0u;
        int *flag_out = hprog + w;
        // PAIR: shared::cluster addresses of what this warp sends to the peer CTA
        const uint32_t words_raddr = PAIR ? mapa_u32(bits_s + lane_utt * R, 0) : 0u, words_rbar = PAIR ? mapa_u32(bar_b, 0) : 0u;
        // what tile j needs, one word per lane: lanes 0..3 the four epilogue warps' progress on this warp's M-tile
        // (>= j + 1: the tile is in the ring), lane 4 the progress of DP warp w - 1 (its halo row), the other lanes a
        // word that is always satisfied.  ONE shared-memory load covers everything, and the load for tile j + 1 is
        // issued before the body of tile j, so its latency never stalls the in-order warp.
        const int *sync_word = PROF_EXP(0) ? hprog + W : ((lane < 4) ? eprog + 4 * w + lane : ((lane == 4 && w > 0) ? hprog + (w - 1) : hprog + W));

        float q[R];
        uint32_t acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { q[r] = P.neg; acc[r] = 0u; }
        // neighbour value for frame 0: only text position 0 has a defined one (core.pyx:24-25, y == 0)
        float up = (x0 == 0) ? 0.f : P.neg;
        int stage = 0, hs = 0;
        PROF_DECL(w_full = 0, w_body = 0);
        // the first tile of this warp's M-tile is in the ring: the prologue is over, and with it every read of the mu_x
        // staging that aliases the halo area -- only now may the halo rows be initialised
        flag_wait_ge_warp(sync_word, 1);
        if (halo_remote) mbar_wait_warp(&bar_h[0], 0);
        if (w == 0 && !halo_remote) {
            hconst[lane] = P.neg;                                // the row above text position 0 (core.pyx:26-27)
            __syncwarp();
        }
        if (dbg && lane == 0 && w == 0) dbg[2] = clock64();
        // the first two value / halo groups of a tile are loaded as soon as the tile is known to be in the ring --
        // for tile j + 1 that is before the tail work of tile j
        float4 va[R], ha;
        dp_tile_prefetch<R, XP>(va, ha, ring + lane_cta * kTilePitch, hb_in, lane7);
        // Everything the tile loop needs beside the recurrence advances by additions (the per-tile overhead is serial with
        // the T_mel-long chain: ~310 cycles of index arithmetic, votes and branches per 32 frames before this).
        const float *const ring_lane = ring + lane_cta * kTilePitch;
        const float *lane_tile = ring_lane;                         // this lane's rows of the current ring stage
        const float *hin = hb_in;                                   // halo input row of the current tile
        uint32_t hout_addr = hout_base;                             // where the bottom row of the current tile goes
        const uint32_t hin_wrap = halo_remote ? 0x7fffffffu : (uint32_t)HS;   // halo slots before the ring wraps (PAIR rank 1: never)
        const uint32_t hout_wrap = PAIR ? 0x7fffffffu : (uint32_t)HS;
        const int jd0 = xw0 / NT;                                   // tiles [jd0, jd0 + 4) hold the diagonal cells of the warp's rows
        int dl0 = lane_utt;                                         // lane_utt - (first frame of the tile) / R
        uint32_t bits_addr = smem_u32(bits_s + lane_utt * R);       // the lane's direction words of the current tile
        uint32_t words_ra = words_raddr, words_rb = words_rbar;     // ... in the home CTA (PAIR rank 1), and their mbarrier
        uint64_t *empty_bar = &ring_empty[w];
        for (int j = 0; j < ntiles; ++j) {
            const int sync_next = flag_acquire(sync_word);       // consumed after the body
            // PAIR rank 1: has the peer's halo row of tile j + 1 landed?  (a non-blocking probe, consumed after the body too)
            const bool halo_next = halo_remote ? mbar_test_warp(&bar_h[min(j + 1, ntiles - 1)], 0) : true;
            const bool wrap = stage + 1 == NS;
            const int next_stage = wrap ? 0 : stage + 1;
            const bool hwrap_in = (uint32_t)(hs + 1) == hin_wrap, hwrap_out = (uint32_t)(hs + 1) == hout_wrap;
            const float *lane_tile_next = wrap ? ring_lane : lane_tile + kTileFloats;
            const float *hin_next = hwrap_in ? hb_in : hin + hin_step;
            const bool diag = (unsigned)(j - jd0) < (unsigned)((32 * R) / NT);
            PROF_T(cb0);
            if (diag) dp_tile_pre<R, XP, true, kFusedCell>(q, acc, up, va, ha, lane_tile, hin, lane7, lane0_mask, dl0, P.neg, hout_addr);
            else dp_tile_pre<R, XP, false, kFusedCell>(q, acc, up, va, ha, lane_tile, hin, lane7, lane0_mask, dl0, P.neg, hout_addr);
            PROF_ADD(w_body, cb0);
            // tile j + 1 ready?  (normally yes: the flags were read before the body) -> its first groups are on their
            // way while this tile's direction words are stored and the tile is released.  After the last tile the same
            // code re-reads a resident stage (no branch on "last tile": one vote covers everything).
            PROF_T(cw0);
            const int need = min(j + 2, ntiles);
            if (!__all_sync(kFullMask, sync_next >= need)) flag_wait_ge_warp(sync_word, need);
            // pair rank 1 runs one to two tiles behind rank 0, so the early probe of the peer's halo row fails about every
            // other tile: wait for it alone (the probe's result is warp-uniform), not for the flags again
            if (!halo_next) mbar_wait_warp(&bar_h[min(j + 1, ntiles - 1)], 0);
            dp_tile_prefetch<R, XP>(va, ha, lane_tile_next, hin_next, lane7);
            PROF_ADD(w_full, cw0);

            // direction words of this tile, walk-ready (see mas_forward_kernel)
            uint32_t words[R];
            dp_finish_words<R, kFusedCell>(acc, words, x0, j, diag);
            if (PAIR && rank == 1)      // to the home CTA, counted on its mbarrier of this tile
                st_async_v4(words_ra, words_rb, words[0], words[1], words[2], words[3]);
            else
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bits_addr), "r"(words[0]), "r"(words[1]), "r"(words[2]), "r"(words[3]) : "memory");

            __syncwarp();                                   // lane 31's halo stores, everyone's ring reads
            if (dbg && j == 0 && lane == 0) dbg[26] = clock64();
            if (elect_one()) {
                flag_release(flag_out, j + 1);
                mbar_arrive(empty_bar);
            }
            stage = next_stage;
            hs = (hwrap_in || hwrap_out) ? 0 : hs + 1;
            lane_tile = lane_tile_next;
            hin = hin_next;
            hout_addr = hwrap_out ? hout_base : hout_addr + hout_step;
            dl0 -= NT / R;
            bits_addr += XPT * 4u;
            words_ra += XPT * 4u;
            words_rb += 8u;
            empty_bar = wrap ? &ring_empty[w] : empty_bar + 2;
        }
#if MASB200_FUSED_PROF
        if (dbg && lane == 0) { dbg[22 + 2 * w] = w_full; dbg[26 + w] = w_body; }
#endif
        if (dbg && lane == 0 && w == w_act - 1) dbg[4] = clock64();
        if (dbg && lane == 0 && w == 0) dbg[30] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, kLpTmemCols); }
    if (dbg && tid == 0) dbg[3] = clock64();
    // PAIR: rank 0's helpers have seen every tile's direction words of the peer land (and the block barrier above hands
    // that to all threads); rank 1's DP warp has consumed every halo row rank 0 sent.  Nothing is in flight INTO either
    // CTA any more, and nobody reads rank 1's shared memory: rank 1 is done.
    if (PAIR && rank != 0) return;

    // ================================ backtrack + outputs (the ring is idle now) ================================
    int *tok = reinterpret_cast<int *>(ring);
    int *xin = tok + XPT;
    uint32_t *sflags = reinterpret_cast<uint32_t *>(xin + ((ntiles + 3) & ~3));
    mas_backtrack_smem<XPT, kFusedThreads, 0, 0>(bits_s, nj_s, xin, sflags, ntiles, ntiles, t_x, t_y, tid, dbg);
    if (dbg && tid == 0) dbg[5] = clock64();
    // durations, frame -> token index, and (dense path, part 2) the ones: frame t belongs to token frame_token[t], written
    // by the thread that produces frame_token[t]
    uint32_t *pb = P.path != nullptr ? reinterpret_cast<uint32_t *>(P.path) + (size_t)b * P.Tx * P.Ty : nullptr;
    const uint32_t one = P.path_dtype == MAS_B200_PATH_F32 ? 0x3f800000u : 1u;
    mas_emit_outputs_flags<kFusedThreads>(P, b, tok, xin, sflags, ntiles, t_x, t_y, tid, dbg, pb, one);
    if (dbg && tid == 0) {
        dbg[6] = clock64(); dbg[7] = ((long long)t_x << 32) | (unsigned)t_y;
        long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[13] = t;
    }
}

constexpr size_t kFusedMaxSmem = 232448 - 1024;

template <int KS, int W, bool PAIR>
int fused_plan(int Tx, int Ty, int *ns_out, size_t *smem_out) {
    using FS = FusedSmem<KS, W, PAIR>;
    const int ntiles = (Ty + kTileFrames - 1) / kTileFrames;
    int cap = option("mas_ring_stages");
    if (cap <= 0 || cap > 6) cap = 6;
    int ns = 0;
    while (ns < cap && FS::total(ns + 1, ntiles) <= kFusedMaxSmem) ++ns;
    if (ns < 2) return MAS_B200_ERR_UNSUPPORTED;
    // the tail's scratch (token starts, tile entry tokens, frame heads) lives in the idle ring
    if (mas_tail_scratch_ints(FS::XPT, ntiles, Ty) * sizeof(int) > FS::M::ring_bytes(ns)) return MAS_B200_ERR_UNSUPPORTED;
    if (PAIR && ntiles > kFusedThreads - 32) return MAS_B200_ERR_UNSUPPORTED;      // one thread arms each hand-off barrier
    *ns_out = ns;
    *smem_out = FS::total(ns, ntiles);
    return MAS_B200_OK;
}

template <int KS, int W, bool PAIR>
int fused_launch(FusedParams &FP, const CUtensorMap &ymap, cudaStream_t stream, bool dry_run) {
    int ns = 0;
    size_t smem = 0;
    int rc = fused_plan<KS, W, PAIR>(FP.mas.Tx, FP.mas.Ty, &ns, &smem);
    if (rc != MAS_B200_OK || dry_run) return rc;
    FP.mas.ring_stages = ns;
    static std::atomic<int> configured[16];
    int dev = 0;
    MASB200_CUDA_TRY(cudaGetDevice(&dev));
    auto kern = lp_mas_fused_kernel<KS, W, PAIR>;
    if (dev < 0 || dev >= 16 || !configured[dev].load(std::memory_order_acquire)) {
        MASB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedMaxSmem));
        if (dev >= 0 && dev < 16) configured[dev].store(1, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)FP.mas.B * (PAIR ? 2u : 1u));
    cfg.blockDim = dim3(kFusedThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (PAIR) {          // the two CTAs of an utterance: one cluster, co-scheduled, shared memory reachable from each other
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (option("pdl") != 0) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    MASB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, FP, ymap));
    return MAS_B200_OK;
}

// TMEM budget: W M-tiles of A (hi, lo, extra K step: 2F + 8 columns each) + two D stages of W x 32 columns <= 512.
// pair: a text of more than 128 tokens as a 2-CTA cluster, one M-tile per CTA (see lp_mas_fused_pair_wanted)
int fused_dispatch(int F, FusedParams &FP, const CUtensorMap &ymap, cudaStream_t stream, bool dry_run, bool pair) {
    const bool one = FP.mas.Tx <= 128;
    if (!one && pair) {
        switch (F) {
            case 64: return fused_launch<8, 1, true>(FP, ymap, stream, dry_run);
            case 80: return fused_launch<10, 1, true>(FP, ymap, stream, dry_run);
            case 96: return fused_launch<12, 1, true>(FP, ymap, stream, dry_run);
            case 128: return fused_launch<16, 1, true>(FP, ymap, stream, dry_run);
            default: return MAS_B200_ERR_UNSUPPORTED;
        }
    }
    switch (F) {
        case 64: return one ? fused_launch<8, 1, false>(FP, ymap, stream, dry_run) : fused_launch<8, 2, false>(FP, ymap, stream, dry_run);
        case 80: return one ? fused_launch<10, 1, false>(FP, ymap, stream, dry_run) : fused_launch<10, 2, false>(FP, ymap, stream, dry_run);
        case 96: return one ? fused_launch<12, 1, false>(FP, ymap, stream, dry_run) : MAS_B200_ERR_UNSUPPORTED;
        case 128: return one ? fused_launch<16, 1, false>(FP, ymap, stream, dry_run) : MAS_B200_ERR_UNSUPPORTED;
        default: return MAS_B200_ERR_UNSUPPORTED;
    }
}

// A text of 129..256 tokens runs as a PAIR of CTAs (one 128-row M-tile, one DP warp each; the halo row and the direction
// words cross between them with st.async) when both CTAs of every utterance are resident at once: two SMs per utterance
// halve what the latency-critical DP warp has to share its SM with (62 -> ~48 cycles per frame) and F = 96 / 128 fit the
// tensor memory.  Beyond that (2B > SMs) the batch is throughput-bound and one CTA per utterance does more per SM.
bool fused_pair_wanted(int B, int Tx) {
    if (Tx <= 128 || option("fused_pair") == 0) return false;
    DeviceInfo di;
    if (device_info(&di) != MAS_B200_OK) return false;
    return 2 * B <= di.sm_count || option("fused_pair") == 2;
}

}  // namespace

// Shapes the fused kernel covers (everything else runs the serial form: log-prior kernel -> HBM -> MAS kernel):
// F in {64, 80} with Tx <= 256 (two M-tiles / two DP warps in one CTA, or a pair of CTAs) or F in {64, 80, 96, 128} with
// Tx <= 128 (one M-tile) or Tx <= 256 as a pair of CTAs -- the A operand (mu_x hi/lo + the extra K step) and two D
// stages have to fit the 512 TMEM columns --, Ty % 4 == 0 and 16-byte aligned operands (TMA), direction words + a
// >= 2-stage ring within 227 KB of shared memory (Ty up to ~2900 frames at Tx <= 128, ~1500 at Tx <= 256).
bool lp_mas_fused_supported(const float *mu_x, const float *y, int B, int F, int Tx, int Ty) {
    if (!(F == 64 || F == 80 || F == 96 || F == 128) || Tx > 256 || Ty % 4 != 0 || B <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(mu_x) & 15)) return false;
    FusedParams FP{};
    FP.mas.Tx = Tx; FP.mas.Ty = Ty;
    CUtensorMap dummy;
    std::memset(&dummy, 0, sizeof(dummy));
    if (fused_pair_wanted(B, Tx) && fused_dispatch(F, FP, dummy, nullptr, true, true) == MAS_B200_OK) return true;
    return fused_dispatch(F, FP, dummy, nullptr, true, false) == MAS_B200_OK;
}

int launch_lp_mas_fused(const float *mu_x, const float *y, const int *t_x, const int *t_y, int B, int F, int Tx, int Ty,
                        float neg, void *path, int path_dtype, int *durations, int *frame_token, int *status,
                        void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!mu_x || !y || !t_x || !t_y || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (path_dtype != MAS_B200_PATH_NONE && path_dtype != MAS_B200_PATH_F32 && path_dtype != MAS_B200_PATH_I32)
        return MAS_B200_ERR_ARG;
    if (path_dtype != MAS_B200_PATH_NONE && !path) return MAS_B200_ERR_ARG;
    if (!lp_mas_fused_supported(mu_x, y, B, F, Tx, Ty)) return MAS_B200_ERR_UNSUPPORTED;
    const Workspace ws = workspace_layout(B, Tx, Ty);
    if (!workspace || workspace_bytes < ws.total) return MAS_B200_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return MAS_B200_ERR_ALIGN;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    CUtensorMap ymap;
    std::memset(&ymap, 0, sizeof(ymap));
    rc = make_y_tensor_map(y, B, F, Ty, kTileFrames, &ymap);
    if (rc != MAS_B200_OK) return rc;

    char *wsb = static_cast<char *>(workspace);
    FusedParams FP{};
    MasParams &P = FP.mas;
    P.t_x = t_x; P.t_y = t_y; P.B = B; P.Tx = Tx; P.Ty = Ty; P.neg = neg;
    P.start = reinterpret_cast<int *>(wsb + ws.start_off);
    P.dur = durations ? durations : reinterpret_cast<int *>(wsb + ws.dur_off);
    P.frame_token = frame_token;
    P.status = status;
    {
        const unsigned lo = (unsigned)option("mas_debug_ptr_lo"), hi = (unsigned)option("mas_debug_ptr_hi");
        P.dbg = reinterpret_cast<long long *>(((unsigned long long)hi << 32) | lo);
    }
    // dense path: written by the kernel itself (zeros streamed out in the shadow of the search, ones in the tail) unless
    // the buffer is not 16-byte aligned or option mas_fused_path_write = 0 asks for the separate expansion kernel
    int fuse = option("mas_fused_path_write");
    fuse = (fuse != 0 && (reinterpret_cast<uintptr_t>(path) & 15) == 0) ? 1 : 0;
    const bool want_path = path_dtype != MAS_B200_PATH_NONE;
    P.path = (want_path && fuse) ? path : nullptr;
    P.path_dtype = (want_path && fuse) ? path_dtype : MAS_B200_PATH_NONE;
    FP.mu = mu_x;
    FP.y_pf = y;
    FP.cst = log_prior_const(F);
    {   // tests only: device pointer to a [B,Tx,Ty] buffer that receives the value tiles (two int options)
        const unsigned lo = (unsigned)option("fused_dump_ptr_lo"), hi = (unsigned)option("fused_dump_ptr_hi");
        FP.value_dump = reinterpret_cast<float *>(((unsigned long long)hi << 32) | lo);
    }
    FP.exp = option("fused_exp");
    if (option("peer_world") > 0) {      // one-sided duration gather (see MasParams::peer_dur)
        const unsigned lo = (unsigned)option("peer_dur_ptrs_lo"), hi = (unsigned)option("peer_dur_ptrs_hi");
        P.peer_dur = reinterpret_cast<int *const *>(((unsigned long long)hi << 32) | lo);
        P.peer_world = option("peer_world");
        P.peer_rank = option("peer_rank");
        if (P.peer_dur == nullptr || P.peer_rank < 0 || P.peer_rank >= P.peer_world) return MAS_B200_ERR_ARG;
    }
    {
        FusedParams probe = FP;
        const bool pair = fused_pair_wanted(B, Tx) && fused_dispatch(F, probe, ymap, stream, true, true) == MAS_B200_OK;
        rc = fused_dispatch(F, FP, ymap, stream, false, pair);
        if (pair && rc == MAS_B200_ERR_CUDA) {
            // the cluster launch was refused (e.g. a partitioned device that cannot co-schedule two such CTAs): the same
            // computation as one CTA per utterance where that form exists, else the caller's serial form
            (void)cudaGetLastError();
            FusedParams retry = FP;
            if (fused_dispatch(F, retry, ymap, stream, true, false) != MAS_B200_OK) return MAS_B200_ERR_UNSUPPORTED;
            rc = fused_dispatch(F, FP, ymap, stream, false, false);
        }
    }
    if (rc != MAS_B200_OK) return rc;
    if (want_path && !fuse) return launch_path_expand(P.start, P.dur, B, Tx, Ty, path, path_dtype, stream);
    return MAS_B200_OK;
}

}  // namespace masb200
