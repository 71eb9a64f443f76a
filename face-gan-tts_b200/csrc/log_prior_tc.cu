// log_prior_tc.cu -- Grad-TTS log-prior on the 5th-generation tensor cores (tcgen05 + TMEM), unfused.
//
// Replaces reference model/face_tts.py:165-171 (see log_prior_ffma.cu for the term-by-term mapping):
//   log_prior[x,t] = ((ysq[t] + dot[x,t]) + musq[x]) + const,   dot = sum_f mu_x[f,x] * y[f,t]
// The K = n_feats contraction runs as 3xTF32 (hi*hi + hi*lo + lo*hi with exact tf32 hi/lo pairs,
// fp32 accumulation in TMEM): ~22 mantissa bits per operand, i.e. fp32-class accuracy, far inside the
// 1e-4 relative bar.  ysq / musq are plain fp32 FMAs on the CUDA cores.
//
// One CTA = one utterance x a run of 32-frame tiles:
//   A  (M = text positions)   mu_x rows, split hi/lo in registers and parked in TENSOR MEMORY for the
//                             whole CTA (tcgen05.st; lane = text position, column = mel bin), so the
//                             operand costs no shared memory and no re-reads.
//   B  (N = 32 frames)        y tiles [F x 32] brought by TMA; the aux warps split them hi/lo and, in the
//                             same pass, transpose them into the K-major core-matrix layout UMMA reads
//                             (8 frames x 4 mel bins per 128-byte core matrix).
//   D  (fp32, TMEM)           2 stages x up to 2 M-tiles x 32 columns.
// warps 0-3  aux: A prologue, y split + ysq, epilogue (tcgen05.ld -> fuse -> global); warp w owns TMEM
//            lanes 32w..32w+31.
// warp 4     one lane: TMA producer for y and MMA issuer (tcgen05.mma kind::tf32, A from TMEM).
// The same front end feeds the fused kernel, whose epilogue writes the MAS ring instead of HBM.
#include <atomic>
#include <cstring>

#include <cudaTypedefs.h>

#include "mas_host.h"
#include "tc_common.cuh"

namespace masb200 {

float log_prior_const(int F);   // log_prior_ffma.cu

namespace {

constexpr int kTcAux = 128;           // aux / epilogue threads (4 warps = the 4 TMEM lane quadrants)
constexpr int kTcThreads = 160;
constexpr int kNSY = 3;               // y stages
constexpr int kTmemCols = 512;
constexpr int kMaxF = 96;             // 4F (A hi/lo, two M-tiles) + 128 (D) <= 512 columns

struct LpTcParams {
    const float *mu;     // [B,F,Tx]
    float *out;          // [B,Tx,Ty]
    int B, F, Tx, Ty;
    float cst;
    int tiles_per_cta, ntiles;
};

// TMEM column map
__device__ __forceinline__ uint32_t col_a(int F, int mt, int lo) { return (uint32_t)((mt * 2 + lo) * F); }
__device__ __forceinline__ uint32_t col_d(int F, int d, int mt) { return (uint32_t)(4 * F + (d * 2 + mt) * 32); }

__device__ __forceinline__ void tma_load_y(void *dst, const CUtensorMap *tmap, int t0, int b, uint64_t *bar) {
    tma_load_3d(dst, tmap, t0, 0, b, bar);
}

__device__ __forceinline__ void aux_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(kTcThreads, 1)
log_prior_tc_kernel(const LpTcParams P, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int F = P.F;
    const uint32_t stage_bytes = (uint32_t)F * 128u;                          // [F][32] fp32
    unsigned char *y_raw = smem_raw;                                          // [kNSY][stage_bytes]  TMA destination
    unsigned char *y_hi = y_raw + (size_t)kNSY * stage_bytes;                 // [2][stage_bytes]     UMMA B operand
    unsigned char *y_lo = y_hi + (size_t)2 * stage_bytes;                     // [2][stage_bytes]
    float *part = reinterpret_cast<float *>(y_lo + (size_t)2 * stage_bytes);  // [2][4 warps][32]
    float *ysq = part + 2 * 4 * 32;                                               // [2][32]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ysq + 2 * 32);
    uint64_t *bar_yfull = bars, *bar_yfree = bars + kNSY, *bar_ysplit = bars + 2 * kNSY;   // ysplit: [2]
    uint64_t *bar_dfull = bar_ysplit + 2, *bar_dempty = bar_dfull + 2, *bar_aready = bar_dempty + 2;
    uint64_t *bar_mu = bar_aready + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mu + 1);
    float *mu_s = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(tmem_slot) + 16);   // [F][Tx] staging

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
    const int lane = tid & 31;
    const int b = blockIdx.y;
    const int j0 = blockIdx.x * P.tiles_per_cta;
    const int n = min(P.tiles_per_cta, P.ntiles - j0);                        // tiles of this CTA
    if (n <= 0) return;
    const int MT = (P.Tx + 127) >> 7;                                         // M-tiles of 128 text positions
    const int KS = F >> 3;                                                    // k-steps of 8 mel bins

    if (tid == 0) {
        for (int s = 0; s < kNSY; ++s) { mbar_init(&bar_yfull[s], 1); mbar_init(&bar_yfree[s], kTcAux); }
        for (int d = 0; d < 2; ++d) { mbar_init(&bar_ysplit[d], kTcAux); mbar_init(&bar_dfull[d], 1); mbar_init(&bar_dempty[d], kTcAux); }
        mbar_init(bar_aready, kTcAux);
        mbar_init(bar_mu, 1);
        mbar_fence_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer + MMA issuer (one lane) =====================
        if (lane == 0) {
            // the utterance's whole mu_x block [F][Tx] in one bulk copy (16-byte aligned since F % 4 == 0)
            const uint32_t mu_bytes = (uint32_t)F * (uint32_t)P.Tx * 4u;
            mbar_arrive_expect_tx(bar_mu, mu_bytes);
            tma_bulk_load_1d(mu_s, P.mu + (size_t)b * F * P.Tx, mu_bytes, bar_mu);
            for (int i = 0; i < min(n, kNSY); ++i) {
                mbar_arrive_expect_tx(&bar_yfull[i], stage_bytes);
                tma_load_y(y_raw + (size_t)i * stage_bytes, &ymap, (j0 + i) * 32, b, &bar_yfull[i]);
            }
            const uint32_t idesc = umma_idesc_tf32_ts(128, 32);
            const uint32_t sbo = (uint32_t)F * 32u;              // 8 frames x F mel bins x 4 B per row group
            mbar_wait(bar_aready, 0);
            for (int i = 0; i < n; ++i) {
                const int s = i % kNSY, d = i & 1;
                mbar_wait(&bar_ysplit[d], (uint32_t)(i >> 1) & 1u);
                // the raw stage is free again: refill it with tile i + kNSY
                if (i + kNSY < n) {
                    mbar_wait(&bar_yfree[s], (uint32_t)(i / kNSY) & 1u);
                    mbar_arrive_expect_tx(&bar_yfull[s], stage_bytes);
                    tma_load_y(y_raw + (size_t)s * stage_bytes, &ymap, (j0 + i + kNSY) * 32, b, &bar_yfull[s]);
                }
                if (i >= 2) mbar_wait(&bar_dempty[d], (uint32_t)((i >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t bh = smem_u32(y_hi + (size_t)d * stage_bytes), bl = smem_u32(y_lo + (size_t)d * stage_bytes);
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t dcol = tmem + col_d(F, d, mt);
                    const uint32_t ah = tmem + col_a(F, mt, 0), al = tmem + col_a(F, mt, 1);
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint64_t dh = umma_smem_desc_k_nosw(bh + ks * 256u, 128u, sbo);
                        const uint64_t dl = umma_smem_desc_k_nosw(bl + ks * 256u, 128u, sbo);
                        umma_tf32_ts(dcol, ah + 8u * ks, dh, idesc, ks > 0 ? 1u : 0u);     // hi * hi
                        umma_tf32_ts(dcol, ah + 8u * ks, dl, idesc, 1u);                    // hi * lo
                        umma_tf32_ts(dcol, al + 8u * ks, dh, idesc, 1u);                    // lo * hi
                    }
                }
                umma_commit(&bar_dfull[d]);
            }
        }
    } else {
        // ============================ aux / epilogue warps ============================
        const int m = tid;                                      // TMEM lane == text position within the M-tile
        const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
        float musq[2] = {0.f, 0.f};
        // ---- A prologue: mu rows -> hi/lo -> tensor memory
        mbar_wait(bar_mu, 0);
        for (int mt = 0; mt < MT; ++mt) {
            const int x = mt * 128 + m;
            const bool xin = x < P.Tx;
            float sq = 0.f;
            for (int f0 = 0; f0 < F; f0 += 8) {
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = xin ? mu_s[(f0 + k) * P.Tx + x] : 0.f;
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    tf32_split(v[k], hi[k], lo[k]);
                    sq = fmaf(-0.5f * v[k], v[k], sq);
                }
                tmem_st8(tmem + lane_base + col_a(F, mt, 0) + f0, hi);
                tmem_st8(tmem + lane_base + col_a(F, mt, 1) + f0, lo);
            }
            musq[mt] = sq;
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_aready);

        float *outb = P.out + (size_t)b * P.Tx * P.Ty;
        const uint32_t sbo = (uint32_t)F * 32u;
        for (int i = 0; i <= n; ++i) {
            if (i < n) {
                // ---- split y tile i hi/lo and transpose it into the K-major core-matrix layout; ysq per frame.
                // thread = (frame n = lane, mel-bin chunk kc = warp, warp+4, ...): 4 conflict-free LDS.32 down a
                // column of the raw tile, one STS.128 per operand into core matrix (n/8, kc), row n%8.
                const int s = i % kNSY, d = i & 1;
                mbar_wait(&bar_yfull[s], (uint32_t)(i / kNSY) & 1u);
                const float *raw = reinterpret_cast<const float *>(y_raw + (size_t)s * stage_bytes);
                unsigned char *hb = y_hi + (size_t)d * stage_bytes, *lb = y_lo + (size_t)d * stage_bytes;
                const uint32_t row_off = (uint32_t)(lane >> 3) * sbo + (uint32_t)(lane & 7) * 16u;
                float q = 0.f;
                for (int kc = warp; kc < (F >> 2); kc += 4) {
                    float v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = raw[(4 * kc + k) * 32 + lane];
                    uint4 h, l;
                    tf32_split(v[0], h.x, l.x); tf32_split(v[1], h.y, l.y);
                    tf32_split(v[2], h.z, l.z); tf32_split(v[3], h.w, l.w);
                    *reinterpret_cast<uint4 *>(hb + row_off + (uint32_t)kc * 128u) = h;
                    *reinterpret_cast<uint4 *>(lb + row_off + (uint32_t)kc * 128u) = l;
#pragma unroll
                    for (int k = 0; k < 4; ++k) q = fmaf(-0.5f * v[k], v[k], q);
                }
                float *pd = part + d * 128;
                pd[warp * 32 + lane] = q;
                fence_proxy_async_smem();              // hi/lo stores -> visible to the tensor core's smem reads
                mbar_arrive(&bar_yfree[s]);            // the raw stage may be refilled
                aux_bar();
                if (tid < 32) ysq[d * 32 + tid] = (pd[tid] + pd[32 + tid]) + (pd[64 + tid] + pd[96 + tid]);
                mbar_arrive(&bar_ysplit[d]);
            }
            if (i >= 1) {
                // ---- epilogue of tile i-1
                const int j = i - 1, d = j & 1;
                const int t0 = (j0 + j) * 32;
                mbar_wait(&bar_dfull[d], (uint32_t)(j >> 1) & 1u);
                tc_fence_after();
                float yq[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4 *>(ysq + d * 32 + 4 * c);
                    yq[4 * c] = v.x; yq[4 * c + 1] = v.y; yq[4 * c + 2] = v.z; yq[4 * c + 3] = v.w;
                }
                for (int mt = 0; mt < MT; ++mt) {
                    uint32_t r[32];
                    tmem_ld32(tmem + lane_base + col_d(F, d, mt), r);
                    tmem_wait_ld();
                    const int x = mt * 128 + m;
                    if (x < P.Tx) {
                        float *dst = outb + (size_t)x * P.Ty + t0;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float4 o;
                            o.x = ((yq[4 * c + 0] + __uint_as_float(r[4 * c + 0])) + musq[mt]) + P.cst;
                            o.y = ((yq[4 * c + 1] + __uint_as_float(r[4 * c + 1])) + musq[mt]) + P.cst;
                            o.z = ((yq[4 * c + 2] + __uint_as_float(r[4 * c + 2])) + musq[mt]) + P.cst;
                            o.w = ((yq[4 * c + 3] + __uint_as_float(r[4 * c + 3])) + musq[mt]) + P.cst;
                            if (t0 + 4 * c + 3 < P.Ty) *reinterpret_cast<float4 *>(dst + 4 * c) = o;     // Ty % 4 == 0
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bar_dempty[d]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
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

// 3-D map over y[b, f, t]: box {32 frames, F mel bins, 1}, no swizzle, zero fill beyond Ty.
int make_y_tensor_map(const float *y, int B, int F, int Ty, CUtensorMap *out) {
    auto enc = encoder();
    if (!enc) return MAS_B200_ERR_CUDA;
    cuuint64_t gdim[3] = {(cuuint64_t)Ty, (cuuint64_t)F, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)Ty * 4, (cuuint64_t)F * Ty * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)F, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(y), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MAS_B200_OK : MAS_B200_ERR_ARG;
}

int launch_log_prior_tc(const float *mu_x, const float *y, int B, int F, int Tx, int Ty, float *out,
                        cudaStream_t stream) {
    if (!mu_x || !y || !out || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    // shapes the tensor-core kernel takes; everything else goes to the FFMA kernel
    if (F % 8 != 0 || F > kMaxF || Tx > 256 || Ty % 4 != 0 || B > 65535) return MAS_B200_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(mu_x) & 15))
        return MAS_B200_ERR_UNSUPPORTED;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    CUtensorMap ymap;
    std::memset(&ymap, 0, sizeof(ymap));
    rc = make_y_tensor_map(y, B, F, Ty, &ymap);
    if (rc != MAS_B200_OK) return rc;

    LpTcParams P{};
    P.mu = mu_x; P.out = out; P.B = B; P.F = F; P.Tx = Tx; P.Ty = Ty; P.cst = log_prior_const(F);
    P.ntiles = (Ty + 31) / 32;
    int chunks = di.sm_count / B;                               // one wave of CTAs (one CTA per SM: TMEM + smem)
    chunks = chunks < 1 ? 1 : (chunks > P.ntiles ? P.ntiles : chunks);
    P.tiles_per_cta = (P.ntiles + chunks - 1) / chunks;
    chunks = (P.ntiles + P.tiles_per_cta - 1) / P.tiles_per_cta;
    const size_t smem = (size_t)(kNSY + 4) * F * 128 + sizeof(float) * (8 * 32 + 2 * 32) + 8 * (2 * kNSY + 8) + 32 +
                        (size_t)F * Tx * 4 + 1024;

    static std::atomic<int> configured[16];
    int dev = 0;
    MASB200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16 || !configured[dev].load(std::memory_order_acquire)) {
        // the opt-in maximum keeps one CTA per SM (each CTA allocates all 512 TMEM columns)
        MASB200_CUDA_TRY(cudaFuncSetAttribute(log_prior_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 16) configured[dev].store(1, std::memory_order_release);
    }
    const size_t smem_launch = smem < (size_t)120 * 1024 ? (size_t)120 * 1024 : smem;     // > half an SM: 1 CTA/SM
    log_prior_tc_kernel<<<dim3(chunks, B), kTcThreads, smem_launch, stream>>>(P, ymap);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

}  // namespace masb200
