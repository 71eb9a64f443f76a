// lp_tc_frontend.cuh -- tcgen05/TMEM front end of the Grad-TTS log-prior, shared by the unfused kernel
// (log_prior_tc.cu: epilogue -> HBM) and the fused kernel (lp_mas_fused.cu: epilogue -> MAS value ring).
//
// Replaces reference model/face_tts.py:165-171 (term-by-term mapping in log_prior_ffma.cu):
//   log_prior[x,t] = (ysq[t] + dot[x,t]) + (musq[x] + const),   dot = sum_f mu_x[f,x] * y[f,t]
// The K = n_feats contraction runs as 3xTF32 (hi*hi + hi*lo + lo*hi over exact tf32 hi/lo pairs, fp32
// accumulation in TMEM): ~22 mantissa bits per operand, fp32-class accuracy (3.8e-7 max relative error
// measured), far inside the 1e-4 bar.  ysq / musq are fp32 FMAs on the CUDA cores.
//
// Per CTA (one utterance, Tx <= 256, F = 8*KS <= 96; or -- split-M, F <= 128 -- one 128-row M-tile of it):
//   A  (M = text positions)  mu_x rows: one cp.async.bulk of the utterance's [F][Tx] block into shared
//                            memory, split hi/lo in registers and parked in TENSOR MEMORY for the whole CTA
//                            (tcgen05.st; lane = text position, column = mel bin) -- the operand costs no
//                            shared memory after the prologue and is never re-read.
//   B  (N = 64 frames)       y groups [F x 64] by TMA; the aux warps split them hi/lo and, in the same pass,
//                            transpose them into the K-major core-matrix layout UMMA reads (8 frames x 4 mel
//                            bins per 128-byte core matrix), double-buffered.
//   D  (fp32, TMEM)          up to 2 M-tiles x 64 columns, single stage, drained by tcgen05.ld.
// MMA issue: tcgen05.mma costs the issuing thread ~100 cycles when its operands have to be moved into
// uniform registers per instruction (scripts/micro/umma_rate.cu), so the MMA warp runs warp-uniform code,
// the K loop is fully unrolled (KS is a template argument) and only the instruction itself is predicated
// on elect.sync; N = 64 keeps the instruction count at 6*KS per 64 frames.
// TMEM map (columns): [0,4F) A hi/lo of M-tile 0 then 1; [4F, 4F+128) D of M-tile 0 then 1.
// Split-M (n_feats = 128, the reference default, needs 4F = 512 columns for A alone): a CTA holds ONE M-tile,
// A in [0,2F), D in [2F, 2F+64); the CTAs of the two M-tiles of an utterance run on different SMs and each
// splits the y groups itself.
#pragma once

#include "tc_common.cuh"

namespace masb200 {

constexpr int kLpGroup = 64;          // frames per MMA group
constexpr int kLpAux = 128;           // aux / epilogue threads (4 warps = the 4 TMEM lane quadrants)
constexpr int kLpTmemCols = 512;

struct LpFrontSmem {
    // byte offsets from a 1024-aligned base; F = 8*KS
    __host__ __device__ static constexpr int nraw(int F) { return F <= 96 ? 2 : 1; }      // raw y-group buffers
    __host__ __device__ static constexpr uint32_t raw_bytes(int F) { return (uint32_t)F * kLpGroup * 4u; }
    __host__ __device__ static constexpr uint32_t op_bytes(int F) { return (uint32_t)F * kLpGroup * 4u; }
    __host__ __device__ static constexpr uint32_t off_raw() { return 0; }                                   // [nraw]
    __host__ __device__ static constexpr uint32_t off_hi(int F) { return (uint32_t)nraw(F) * raw_bytes(F); }   // [2]
    __host__ __device__ static constexpr uint32_t off_lo(int F) { return off_hi(F) + 2 * op_bytes(F); }     // [2]
    __host__ __device__ static constexpr uint32_t off_part(int F) { return off_lo(F) + 2 * op_bytes(F); }   // [2][2][64] f32
    __host__ __device__ static constexpr uint32_t off_ysq(int F) { return off_part(F) + 2 * 2 * 64 * 4; }   // [4][64] f32 ring
    __host__ __device__ static constexpr uint32_t off_bars(int F) { return off_ysq(F) + 4 * 64 * 4; }       // 16 mbarrier slots
    __host__ __device__ static constexpr uint32_t off_tmem(int F) { return off_bars(F) + 16 * 8; }
    __host__ __device__ static constexpr uint32_t total(int F) { return off_tmem(F) + 16; }
};

struct LpFront {
    unsigned char *raw, *hi, *lo;
    float *part, *ysq;
    uint64_t *bar_raw /*[2]*/, *bar_split /*[2]*/, *bar_dfull, *bar_dempty, *bar_aready, *bar_mu, *bar_bfree /*[2]*/,
        *bar_staged /*[2 parities][2 M-tiles]*/, *bar_stfree /*[2 M-tiles]*/;
    uint32_t *tmem_slot;
    long long *prof = nullptr;     // diagnostics: per-CTA wait-cycle accumulators (see scripts/timeline.py), normally null
    __device__ __forceinline__ void carve(unsigned char *base, int F) {
        raw = base + LpFrontSmem::off_raw();
        hi = base + LpFrontSmem::off_hi(F);
        lo = base + LpFrontSmem::off_lo(F);
        part = reinterpret_cast<float *>(base + LpFrontSmem::off_part(F));
        ysq = reinterpret_cast<float *>(base + LpFrontSmem::off_ysq(F));
        uint64_t *b = reinterpret_cast<uint64_t *>(base + LpFrontSmem::off_bars(F));
        bar_raw = b; bar_split = b + 2; bar_dfull = b + 4; bar_dempty = b + 5; bar_aready = b + 6; bar_mu = b + 7;
        bar_bfree = b + 8; bar_staged = b + 10; bar_stfree = b + 14;
        tmem_slot = reinterpret_cast<uint32_t *>(base + LpFrontSmem::off_tmem(F));
    }
    __device__ __forceinline__ void init_barriers() {       // one thread
        mbar_init(&bar_raw[0], 1); mbar_init(&bar_raw[1], 1);
        mbar_init(&bar_split[0], kLpAux); mbar_init(&bar_split[1], kLpAux);
        mbar_init(bar_dfull, 1); mbar_init(bar_dempty, kLpAux);
        mbar_init(bar_aready, kLpAux); mbar_init(bar_mu, 1);
        mbar_init(&bar_bfree[0], 1); mbar_init(&bar_bfree[1], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bar_staged[i], kLpAux);
        mbar_init(&bar_stfree[0], 1); mbar_init(&bar_stfree[1], 1);
    }
};

__device__ __forceinline__ uint32_t lp_col_a(int F, int mt, int lo) { return (uint32_t)((mt * 2 + lo) * F); }
// mtmax = M-tiles one CTA holds: 2 (Tx <= 256 in one CTA, F <= 96) or 1 (split-M: one CTA per M-tile, F <= 128)
__device__ __forceinline__ uint32_t lp_col_d(int F, int mt, int mtmax = 2) { return (uint32_t)(2 * mtmax * F + mt * kLpGroup); }

// named barriers of the two 128-thread groups: 1 = split warps, 2 = epilogue warps
__device__ __forceinline__ void lp_aux_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void lp_epi_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// elect-predicated MMA (the whole warp executes this; one lane issues)
__device__ __forceinline__ void umma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                   uint32_t accumulate) {
    asm volatile(
        "{\n"
        " .reg .pred p, q;\n"
        " elect.sync _|q, 0xffffffff;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n"
        " .reg .pred q;\n"
        " elect.sync _|q, 0xffffffff;\n"
        " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------------------
// MMA / TMA warp: all 32 lanes run this (warp-uniform); ng groups of 64 frames: frames t_begin + k*t_stride.
// ---------------------------------------------------------------------------------------------------
template <int KS, int MTMAX = 2>
__device__ __forceinline__ void lp_mma_warp(const LpFront &S, const CUtensorMap *ymap, const float *mu_b, float *mu_stage,
                                            int Tx, int b, int t_begin, int t_stride, int ng, int MT, uint32_t tmem) {
    constexpr int F = 8 * KS;
    constexpr uint32_t kSbo = (uint32_t)F * 32u;                 // 8 frames x F mel bins x 4 B per row group
    const uint32_t idesc = umma_idesc_tf32_ts(128, kLpGroup);
    constexpr int NRAW = LpFrontSmem::nraw(F);
    if (elect_one()) {
        // the utterance's whole mu_x block [F][Tx] in one bulk copy (16-byte aligned since F % 4 == 0)
        if (mu_stage != nullptr) {
            const uint32_t mu_bytes = (uint32_t)F * (uint32_t)Tx * 4u;
            mbar_arrive_expect_tx(S.bar_mu, mu_bytes);
            tma_bulk_load_1d(mu_stage, mu_b, mu_bytes, S.bar_mu);
        }
        for (int q = 0; q < NRAW && q < ng; ++q) {
            mbar_arrive_expect_tx(&S.bar_raw[q], LpFrontSmem::raw_bytes(F));
            tma_load_3d(S.raw + (size_t)q * LpFrontSmem::raw_bytes(F), ymap, t_begin + q * t_stride, 0, b, &S.bar_raw[q]);
        }
    }
    __syncwarp();
    long long w_split = 0, w_dempty = 0, c_start = clock64();
    mbar_wait_warp(S.bar_aready, 0);
    const long long c_loop = clock64();
    for (int g = 0; g < ng; ++g) {
        const int p = g & 1;
        long long c0 = clock64();
        mbar_wait_warp(&S.bar_split[p], (uint32_t)(g >> 1) & 1u);
        w_split += clock64() - c0;
        // every split thread is done with raw buffer g % NRAW: fetch group g + NRAW into it
        if (g + NRAW < ng && elect_one()) {
            const int q = g % NRAW;
            mbar_arrive_expect_tx(&S.bar_raw[q], LpFrontSmem::raw_bytes(F));
            tma_load_3d(S.raw + (size_t)q * LpFrontSmem::raw_bytes(F), ymap, t_begin + (g + NRAW) * t_stride, 0, b, &S.bar_raw[q]);
        }
        __syncwarp();
        c0 = clock64();
        if (g >= 1) mbar_wait_warp(S.bar_dempty, (uint32_t)(g - 1) & 1u);
        w_dempty += clock64() - c0;
        tc_fence_after();
        const uint32_t bh = smem_u32(S.hi) + (uint32_t)p * LpFrontSmem::op_bytes(F);
        const uint32_t bl = smem_u32(S.lo) + (uint32_t)p * LpFrontSmem::op_bytes(F);
        for (int mt = 0; mt < MT; ++mt) {
            const uint32_t dcol = tmem + lp_col_d(F, mt, MTMAX);
            const uint32_t ah = tmem + lp_col_a(F, mt, 0), al = tmem + lp_col_a(F, mt, 1);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint64_t dh = umma_smem_desc_k_nosw(bh + ks * 256u, 128u, kSbo);
                const uint64_t dl = umma_smem_desc_k_nosw(bl + ks * 256u, 128u, kSbo);
                umma_tf32_ts_elect(dcol, ah + 8u * ks, dh, idesc, ks > 0 ? 1u : 0u);     // hi * hi
                umma_tf32_ts_elect(dcol, ah + 8u * ks, dl, idesc, 1u);                    // hi * lo
                umma_tf32_ts_elect(dcol, al + 8u * ks, dh, idesc, 1u);                    // lo * hi
            }
        }
        umma_commit_elect(S.bar_dfull);           // D of group g complete -> epilogue warps
        umma_commit_elect(&S.bar_bfree[p]);       // ... and operand buffer p may be refilled -> split warps
        __syncwarp();
    }
    if (S.prof && elect_one()) {
        S.prof[10] = w_split; S.prof[11] = w_dempty; S.prof[12] = clock64() - c_loop; S.prof[13] = c_loop - c_start;
    }
}

// ---------------------------------------------------------------------------------------------------
// epilogue warps (threads 0..127; warp w owns TMEM lanes 32w..32w+31) and split warps (a second group of 128
// threads, passed in as tid 0..127 / warp 0..3).  The two groups run concurrently: split(g+1) overlaps
// MMA(g) and epilogue(g-1); hand-offs are mbarriers only (raw -> split -> MMA -> epilogue, bfree: MMA -> split).
// ---------------------------------------------------------------------------------------------------
// A prologue: row_of(mt, m) gives the text position parked in lane m of M-tile mt (any permutation).
// mu_stage: the utterance's [F][Tx] block, in shared memory (GLOBAL_MU = false: staged by the MMA warp's bulk copy,
// bar_mu) or straight in global memory (GLOBAL_MU = true: the split-M form reads only its own 128 columns, which
// no 16-byte-aligned bulk copy can express for odd Tx; consecutive threads read consecutive x: coalesced).
template <int KS, bool GLOBAL_MU = false, class RowOf>
__device__ __forceinline__ void lp_aux_prologue(const LpFront &S, const float *mu_stage, int Tx, int MT, uint32_t tmem,
                                                int tid, int warp, RowOf row_of, float (&musq)[2]) {
    constexpr int F = 8 * KS;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    if (!GLOBAL_MU) mbar_wait(S.bar_mu, 0);
    musq[0] = musq[1] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        if (mt < MT) {
            const int x = row_of(mt, tid);
            const bool xin = x < Tx;
            float sq = 0.f;
#pragma unroll
            for (int f0 = 0; f0 < F; f0 += 8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float v = xin ? (GLOBAL_MU ? __ldg(mu_stage + (size_t)(f0 + k) * Tx + x) : mu_stage[(f0 + k) * Tx + x]) : 0.f;
                    tf32_split(v, hi[k], lo[k]);
                    sq = fmaf(-0.5f * v, v, sq);
                }
                tmem_st8(tmem + lane_base + lp_col_a(F, mt, 0) + f0, hi);
                tmem_st8(tmem + lane_base + lp_col_a(F, mt, 1) + f0, lo);
            }
            musq[mt] = sq;
        }
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(S.bar_aready);
}

// split group g: raw [F][64] -> hi/lo K-major core matrices, ysq per frame.
// thread = (frame n = 32*(warp&1) + lane, mel-bin chunk kc = warp>>1, +2, ...): 4 conflict-free LDS.32 down a
// column of the raw tile, one STS.128 per operand into core matrix (n/8, kc), row n%8.
template <int KS>
__device__ __forceinline__ void lp_aux_split(const LpFront &S, int g, int tid, int warp, int lane, bool skip_math = false) {
    constexpr int F = 8 * KS;
    constexpr uint32_t kSbo = (uint32_t)F * 32u;
    constexpr int NRAW = LpFrontSmem::nraw(F);
    const int p = g & 1;
    long long c0 = clock64();
    if (g >= 2) mbar_wait(&S.bar_bfree[p], (uint32_t)((g >> 1) - 1) & 1u);      // MMA(g-2) has read operand buffer p
    long long c1 = clock64();
    mbar_wait(&S.bar_raw[g % NRAW], (uint32_t)(g / NRAW) & 1u);
    if (S.prof && tid == 0) { S.prof[8] += c1 - c0; S.prof[7] += clock64() - c1; }
    const float *raw = reinterpret_cast<const float *>(S.raw + (size_t)(g % NRAW) * LpFrontSmem::raw_bytes(F));
    unsigned char *hb = S.hi + (size_t)p * LpFrontSmem::op_bytes(F), *lb = S.lo + (size_t)p * LpFrontSmem::op_bytes(F);
    const int n = 32 * (warp & 1) + lane;
    const uint32_t row_off = (uint32_t)(n >> 3) * kSbo + (uint32_t)(n & 7) * 16u;
    float q = 0.f;
#pragma unroll
    for (int kc = (warp >> 1); kc < (skip_math ? 0 : 2 * KS); kc += 2) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = raw[(4 * kc + k) * kLpGroup + n];
        uint4 h, l;
        tf32_split(v[0], h.x, l.x); tf32_split(v[1], h.y, l.y);
        tf32_split(v[2], h.z, l.z); tf32_split(v[3], h.w, l.w);
        *reinterpret_cast<uint4 *>(hb + row_off + (uint32_t)kc * 128u) = h;
        *reinterpret_cast<uint4 *>(lb + row_off + (uint32_t)kc * 128u) = l;
#pragma unroll
        for (int k = 0; k < 4; ++k) q = fmaf(-0.5f * v[k], v[k], q);
    }
    float *pd = S.part + p * 128;
    pd[(warp >> 1) * 64 + n] = q;
    fence_proxy_async_smem();              // hi/lo stores -> visible to the tensor core's smem reads
    lp_aux_bar();
    if (tid < 64) S.ysq[(g & 3) * 64 + tid] = pd[tid] + pd[64 + tid];      // ring of 4: epilogue(g-3) is long done
    mbar_arrive(&S.bar_split[p]);          // also: the raw buffer may be refilled
}

// drain D of group g into registers (both M-tiles, 64 columns each) and hand the accumulator back.
__device__ __forceinline__ void lp_aux_drain(const LpFront &S, int F, int g, int warp, int MT, uint32_t tmem,
                                             uint32_t (&d0)[2][32], uint32_t (&d1)[2][32], int mtmax = 2) {
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    const long long c0 = clock64();
    mbar_wait(S.bar_dfull, (uint32_t)g & 1u);
    if (S.prof && threadIdx.x == 0) S.prof[5] += clock64() - c0;
    tc_fence_after();
    tmem_ld32(tmem + lane_base + lp_col_d(F, 0, mtmax), d0[0]);
    tmem_ld32(tmem + lane_base + lp_col_d(F, 0, mtmax) + 32, d0[1]);
    if (MT > 1) {
        tmem_ld32(tmem + lane_base + lp_col_d(F, 1, mtmax), d1[0]);
        tmem_ld32(tmem + lane_base + lp_col_d(F, 1, mtmax) + 32, d1[1]);
    }
    tmem_wait_ld();
    tc_fence_before();
    mbar_arrive(S.bar_dempty);
}

}  // namespace masb200
