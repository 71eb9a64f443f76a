// log_prior_tc.cu -- Grad-TTS log-prior on the 5th-generation tensor cores (tcgen05 + TMEM), unfused:
// the front end of lp_tc_frontend.cuh with an epilogue that streams [B,Tx,Ty] to HBM.
// One CTA = one utterance x a run of 64-frame groups; warps 0-3 epilogue, warps 4-7 operand split, warp 8 TMA loads + MMA issue,
// warps 9-10 (even / odd groups) TMA stores of the finished tiles.
#include <atomic>
#include <cstring>

#include <cudaTypedefs.h>

#include "lp_tc_frontend.cuh"
#include "mas_forward.cuh"
#include "mas_host.h"

namespace masb200 {

float log_prior_const(int F);   // log_prior_ffma.cu

namespace {

constexpr int kTcThreads = 352;        // warps 0-3 epilogue, 4-7 split, 8 TMA loads + MMA issue, 9-10 TMA stores
constexpr int kMaxF = 96;             // 4F (A hi/lo, two M-tiles) + 128 (D) <= 512 columns; beyond: split-M (2F + 64)

struct LpTcParams {
    const float *mu;     // [B,F,Tx]
    float *out;          // [B,Tx,Ty]
    int B, Tx, Ty;
    float cst;
    int groups_per_cta, ngroups;
    int chunks;          // CTAs per utterance
    long long *dbg;      // diagnostics: [ctas][32] globaltimer stamps / wait-cycle accumulators
    int skip;            // diagnostics (option lp_debug_skip): 1 no global stores, 2 no MMA issue, 4 no split math, 8 no staging
};

// SPLITM: one CTA per (utterance, group run, M-tile) -- blockIdx.z is the M-tile; see lp_tc_frontend.cuh.
template <int KS, bool SPLITM>
__global__ void __launch_bounds__(kTcThreads, 1)
log_prior_tc_kernel(const LpTcParams P, const __grid_constant__ CUtensorMap ymap, const __grid_constant__ CUtensorMap omap) {
    constexpr int F = 8 * KS;
    constexpr int MTMAX = SPLITM ? 1 : 2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    LpFront S;
    S.carve(smem_raw, F);
    // epilogue staging: per (M-tile, 32-frame half) a [128 rows][32 frames] fp32 box in the 128-byte-swizzle layout
    // (16-byte chunk ^= row & 7): written row-per-thread (what tcgen05.ld hands out) without bank conflicts and
    // stored by TMA (tensor map over out[B][Tx][Ty], box {32,128,1}), which also clips rows >= Tx and frames >= Ty.
    unsigned char *stage = smem_raw + ((LpFrontSmem::total(F) + 1023) / 1024) * 1024;
    // [F][Tx] staging of mu_x for the prologue: in the epilogue boxes when it fits there (first written after the
    // prologue, by the same threads) -- then the split warps start on the first y groups while the prologue runs --
    // else in the hi/lo operand buffers (4 * F * 64 * 4 bytes >= F * Tx * 4 for Tx <= 256), which the split warps
    // may only fill once every epilogue thread has left the prologue (bar_aready).
    // The split-M form stages nothing: its prologue reads the CTA's 128 columns of mu_x from global memory.
    const bool mu_in_stage = SPLITM || (size_t)F * P.Tx * 4 <= (size_t)MTMAX * 32768;
    float *mu_s = SPLITM ? nullptr : reinterpret_cast<float *>(mu_in_stage ? stage : S.hi);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
    const int lane = tid & 31;
    const int b = blockIdx.y;
    constexpr int gs = 1;                                                     // group stride of this CTA
    const int g0 = (int)blockIdx.x * P.groups_per_cta;
    const int ng = min(P.groups_per_cta, P.ngroups - g0);                     // groups of this CTA
    if (ng <= 0) return;
    const int MT = SPLITM ? 1 : (P.Tx + 127) >> 7;                           // M-tiles of 128 text positions in this CTA
    const int mt0 = SPLITM ? (int)blockIdx.z : 0;

    long long *dbg = P.dbg ? P.dbg + ((size_t)(blockIdx.z * gridDim.y + b) * gridDim.x + blockIdx.x) * 32 : nullptr;
    S.prof = dbg;
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[0] = t; }
    if (tid == 0) { S.init_barriers(); mbar_fence_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(S.tmem_slot, kLpTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(kFullMask, *S.tmem_slot, 0);

    if (warp == 8) {
        lp_mma_warp<KS, MTMAX>(S, &ymap, P.mu + (size_t)b * F * P.Tx, mu_s, P.Tx, b, g0 * kLpGroup, gs * kLpGroup, ng,
                               (P.skip & 2) ? 0 : MT, tmem);
    } else if (warp >= 4 && warp < 8) {
        // ---- split warps: raw y group -> hi/lo K-major operands + ysq
        if (!mu_in_stage) mbar_wait(S.bar_aready, 0);
        const long long cs0 = clock64();
        for (int g = 0; g < ng; ++g) lp_aux_split<KS>(S, g, tid - 128, warp - 4, lane, (P.skip & 4) != 0);
        if (dbg && tid == 128) dbg[9] = clock64() - cs0;
    } else if (warp >= 9) {
        // ---- store warps: warp 9 takes the even groups, warp 10 the odd ones.  One elected lane per step,
        // warp-uniform control flow.
        for (int gg = warp - 9; gg < ng; gg += 2) {
            const int gidx = g0 + gg * gs;
            const int t0 = gidx * kLpGroup;
            for (int mt = 0; mt < MT; ++mt) {
                // M-tile by M-tile: the boxes of tile 0 are on their way while the epilogue warps still stage tile 1
                mbar_wait_warp(&S.bar_staged[(gg & 1) * 2 + mt], (uint32_t)(gg >> 1) & 1u);
                if (elect_one()) {
                    if (!(P.skip & 1)) {
                        for (int h = 0; h < 2; ++h)
                            tma_store_3d(&omap, stage + (size_t)(mt * 2 + h) * 16384, t0 + 32 * h, (mt0 + mt) * 128, b);
                    }
                    tma_store_commit();
                    tma_store_wait_read();
                    mbar_arrive(&S.bar_stfree[mt]);              // the boxes of this M-tile may be refilled
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {
        float musq[2];
        lp_aux_prologue<KS, SPLITM>(S, SPLITM ? P.mu + (size_t)b * F * P.Tx : mu_s, P.Tx, MT, tmem, tid, warp,
                                    [mt0](int mt, int m) { return (mt0 + mt) * 128 + m; }, musq);
        if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[4] = t; }
        if (dbg && tid == 0) dbg[3] = clock64();
        long long w_stfree = 0;
        for (int gg = 0; gg < ng; ++gg) {
            // ---- epilogue of group gg: thread = text position, 64 consecutive frames per M-tile
            const int p = gg & 3;
            uint32_t d0[2][32], d1[2][32];
            lp_aux_drain(S, F, gg, warp, MT, tmem, d0, d1, MTMAX);
#pragma unroll
            for (int mt = 0; mt < MTMAX; ++mt) {
                if (mt >= MT) break;
                if (gg >= 1) {                                   // the store warp has read this tile's previous boxes
                    const long long c0 = clock64();
                    mbar_wait(&S.bar_stfree[mt], (uint32_t)(gg - 1) & 1u);
                    w_stfree += clock64() - c0;
                }
                const int x = (mt0 + mt) * 128 + tid;
                if (x < P.Tx && !(P.skip & 8)) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        unsigned char *box = stage + (size_t)(mt * 2 + h) * 16384 + (size_t)tid * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 yq = *reinterpret_cast<const float4 *>(S.ysq + p * 64 + 32 * h + 4 * c);
                            const uint32_t *r = (mt == 0) ? &d0[h][4 * c] : &d1[h][4 * c];
                            float4 o;
                            // (ysq + dot) + (musq + const): the association the fused kernel uses (lp_mas_fused.cu)
                            const float ms = musq[mt] + P.cst;
                            o.x = (yq.x + __uint_as_float(r[0])) + ms;
                            o.y = (yq.y + __uint_as_float(r[1])) + ms;
                            o.z = (yq.z + __uint_as_float(r[2])) + ms;
                            o.w = (yq.w + __uint_as_float(r[3])) + ms;
                            *reinterpret_cast<float4 *>(box + ((c ^ (tid & 7)) << 4)) = o;
                        }
                    }
                }
                fence_proxy_async_smem();                        // staging stores -> visible to the TMA store
                mbar_arrive(&S.bar_staged[(gg & 1) * 2 + mt]);
            }
        }
        if (dbg && tid == 0) dbg[14] = w_stfree;
    }
    if (dbg && tid == 0) dbg[6] = clock64() - dbg[3];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kLpTmemCols);
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[1] = t; }
}

PFN_cuTensorMapEncodeTiled_v12000 encoder() {
    static std::atomic<void *> cached{nullptr};
    void *fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        cached.store(fn, std::memory_order_release);
    }
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

}  // namespace

// 3-D map over y[b, f, t]: box {box_frames, F mel bins, 1}, no swizzle, zero fill beyond Ty.
int make_y_tensor_map(const float *y, int B, int F, int Ty, int box_frames, CUtensorMap *out) {
    auto enc = encoder();
    if (!enc) return MAS_B200_ERR_CUDA;
    cuuint64_t gdim[3] = {(cuuint64_t)Ty, (cuuint64_t)F, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)Ty * 4, (cuuint64_t)F * Ty * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_frames, (cuuint32_t)F, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(y), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MAS_B200_OK : MAS_B200_ERR_ARG;
}

// 3-D map over out[b, x, t] for the epilogue's TMA stores: box {32 frames, 128 rows, 1}, 128-byte swizzle.
static int make_out_tensor_map(float *out, int B, int Tx, int Ty, CUtensorMap *map) {
    auto enc = encoder();
    if (!enc) return MAS_B200_ERR_CUDA;
    cuuint64_t gdim[3] = {(cuuint64_t)Ty, (cuuint64_t)Tx, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)Ty * 4, (cuuint64_t)Tx * Ty * 4};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MAS_B200_OK : MAS_B200_ERR_ARG;
}

// shapes the tensor-core kernel takes; everything else goes to the FFMA kernel
bool log_prior_tc_supported(const float *mu_x, const float *y, const float *out, int B, int F, int Tx, int Ty) {
    if (!(F == 64 || F == 80 || F == 96 || F == 128) || Tx > 128 * 65535 || Ty % 4 != 0 || B > 65535) return false;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(mu_x) & 15))
        return false;
    return true;
}

// split-M (one CTA per 128-row M-tile): n_feats = 128, whose A operand alone fills TMEM, and texts longer than the
// two M-tiles one CTA holds
static bool lp_split_m(int F, int Tx) { return F > kMaxF || Tx > 256; }

int launch_log_prior_tc(const float *mu_x, const float *y, int B, int F, int Tx, int Ty, float *out,
                        cudaStream_t stream) {
    if (!mu_x || !y || !out || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!log_prior_tc_supported(mu_x, y, out, B, F, Tx, Ty)) return MAS_B200_ERR_UNSUPPORTED;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    CUtensorMap ymap;
    std::memset(&ymap, 0, sizeof(ymap));
    rc = make_y_tensor_map(y, B, F, Ty, kLpGroup, &ymap);
    if (rc != MAS_B200_OK) return rc;

    CUtensorMap omap;
    std::memset(&omap, 0, sizeof(omap));
    rc = make_out_tensor_map(out, B, Tx, Ty, &omap);
    if (rc != MAS_B200_OK) return rc;

    LpTcParams P{};
    P.mu = mu_x; P.out = out; P.B = B; P.Tx = Tx; P.Ty = Ty; P.cst = log_prior_const(F);
    P.ngroups = (Ty + kLpGroup - 1) / kLpGroup;
    {
        const unsigned lo = (unsigned)option("lp_debug_ptr_lo"), hi = (unsigned)option("lp_debug_ptr_hi");
        P.dbg = reinterpret_cast<long long *>(((unsigned long long)hi << 32) | lo);
    }
    P.skip = option("lp_debug_skip");
    const bool splitm = lp_split_m(F, Tx);
    const int mtiles = splitm ? (Tx + 127) / 128 : 1;          // grid.z
    const int cta_budget = di.sm_count;
    int chunks = cta_budget / (B * mtiles);                                // one wave of CTAs (one CTA per SM: TMEM + smem)
    chunks = chunks < 1 ? 1 : (chunks > P.ngroups ? P.ngroups : chunks);
    P.groups_per_cta = (P.ngroups + chunks - 1) / chunks;
    chunks = (P.ngroups + P.groups_per_cta - 1) / P.groups_per_cta;
    P.chunks = chunks;
    size_t smem = ((LpFrontSmem::total(F) + 1023) / 1024) * 1024 + (size_t)(splitm ? 1 : 2) * 32768 + 1024;
    if (smem < (size_t)120 * 1024) smem = (size_t)120 * 1024;   // > half an SM: one CTA per SM (each allocates all of TMEM)

    void (*kern)(const LpTcParams, const CUtensorMap, const CUtensorMap) = nullptr;
    switch (F) {
        case 64: kern = splitm ? log_prior_tc_kernel<8, true> : log_prior_tc_kernel<8, false>; break;
        case 80: kern = splitm ? log_prior_tc_kernel<10, true> : log_prior_tc_kernel<10, false>; break;
        case 96: kern = splitm ? log_prior_tc_kernel<12, true> : log_prior_tc_kernel<12, false>; break;
        case 128: kern = log_prior_tc_kernel<16, true>; break;
        default: return MAS_B200_ERR_UNSUPPORTED;
    }
    static std::atomic<int> configured[16][8];
    int dev = 0;
    MASB200_CUDA_TRY(cudaGetDevice(&dev));
    const int ki = (F == 64 ? 0 : (F == 80 ? 1 : (F == 96 ? 2 : 3))) + (splitm ? 4 : 0);
    if (dev < 0 || dev >= 16 || !configured[dev][ki].load(std::memory_order_acquire)) {
        MASB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        if (dev >= 0 && dev < 16) configured[dev][ki].store(1, std::memory_order_release);
    }
    kern<<<dim3(chunks, B, mtiles), kTcThreads, smem, stream>>>(P, ymap, omap);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

}  // namespace masb200
