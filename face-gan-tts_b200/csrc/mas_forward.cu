// mas_forward.cu -- plan selection + launch of the MAS kernel (mas_forward.cuh).
#include <atomic>
#include <cstdio>
#include <cstring>

#include <cudaTypedefs.h>

#include "mas_forward.cuh"
#include "mas_host.h"

namespace masb200 {

// one translation unit per rows-per-lane value (mas_forward_inst.cu, -DMASB200_INST_R=...)
int launch_mas_r1(const MasParams &P, const CUtensorMap &tmap, int W, int mode, size_t smem, cudaStream_t stream);
int launch_mas_r2(const MasParams &P, const CUtensorMap &tmap, int W, int mode, size_t smem, cudaStream_t stream);
int launch_mas_r4(const MasParams &P, const CUtensorMap &tmap, int W, int mode, size_t smem, cudaStream_t stream);
int launch_mas_r8(const MasParams &P, const CUtensorMap &tmap, int W, int mode, size_t smem, cudaStream_t stream);

namespace {

struct Plan {
    int R, W;
    bool smem_bits;
    int NS;
    size_t smem;
};

constexpr size_t kMaxSmem = 232448 - 1024;   // dynamic-smem attribute set on every kernel (227 KB opt-in minus slack)
constexpr size_t kStaticSmem = 64;           // bt_state + slack

size_t fixed_bytes_rw(int R, int W, int ns) {
    const int XP = 32 * R * W;
    const size_t ring = sizeof(float) * (size_t)ns * XP * kTilePitch;
    const size_t halo = sizeof(float) * (((size_t)(W + 1) * (ns + 1)) * kTileFrames + (size_t)W * 160);   // MasSmem::halo_bytes
    const size_t ctrl = 8 * (size_t)(2 * ns) + 4 * (size_t)(W + 2) + 64;
    return ((ring + halo + ctrl + 127) / 128) * 128;
}

bool valid_rw(int R, int W) { return (R == 1 || R == 2 || R == 4 || R == 8) && W >= 1 && W <= 4 && 32 * R * W <= 768; }

// Pick rows-per-lane / DP warps / ring depth for (B, Tx, Ty).
bool make_plan(int B, int Tx, int Ty, int sm_count, Plan *p) {
    // R = 4 rows per lane hides the lane-to-lane shuffle behind the other rows' updates (29 static
    // cycles/frame against 33 for R = 2 and 41 for R = 8, scripts/tools/sass_stalls.py)
    static const int table[][2] = {{1, 1}, {2, 1}, {4, 1}, {4, 2}, {4, 3}, {4, 4}};
    int R = 4, W = 4;
    for (auto &rw : table) {
        if (32 * rw[0] * rw[1] >= Tx) { R = rw[0]; W = rw[1]; break; }
    }
    const int oR = option("mas_rows_per_lane"), oW = option("mas_dp_warps");
    if (oR > 0 && oW > 0 && valid_rw(oR, oW)) { R = oR; W = oW; }
    const int XP = 32 * R * W;
    const int tiles = (Ty + kTileFrames - 1) / kTileFrames;

    // shared-memory budget: one CTA per SM for small batches (deepest ring, lowest latency),
    // several co-resident CTAs once the batch exceeds the SM count (fill idle issue slots).
    int per_sm = option("mas_ctas_per_sm");
    if (per_sm <= 0) per_sm = B <= sm_count ? 1 : (B <= 2 * sm_count ? 2 : 3);
    size_t budget = kMaxSmem / per_sm - kStaticSmem - (per_sm > 1 ? 1024 : 0);

    const size_t tile_bytes = sizeof(float) * (size_t)XP * kTilePitch;
    // direction words + the per-tile transfer table (one byte per row and tile) of the backtrack
    const size_t bits_bytes = (sizeof(uint32_t) + 1) * (size_t)tiles * XP;
    const bool single_pass = Tx <= XP;
    const int ns_cap = option("mas_ring_stages") > 0 ? option("mas_ring_stages") : 8;

    auto ring_for = [&](size_t avail) {
        int ns = 0;
        while (ns < ns_cap && fixed_bytes_rw(R, W, ns + 1) <= avail) ++ns;
        return ns;
    };
    bool smem_bits = single_pass && option("mas_force_global_bits") <= 0 && bits_bytes + 2 * tile_bytes + 4096 <= budget;
    int ns = smem_bits ? ring_for(budget - bits_bytes) : ring_for(budget);
    if (ns < 2) {
        // fall back to the full opt-in budget
        budget = kMaxSmem - kStaticSmem;
        smem_bits = single_pass && option("mas_force_global_bits") <= 0 && bits_bytes + 2 * tile_bytes + 4096 <= budget;
        ns = smem_bits ? ring_for(budget - bits_bytes) : ring_for(budget);
        if (ns < 2) return false;
    }
    p->R = R; p->W = W; p->smem_bits = smem_bits; p->NS = ns;
    p->smem = fixed_bytes_rw(R, W, ns) + (smem_bits ? bits_bytes : 0);
    return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
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

// 3-D map over value[b, x, t]: box {32 frames, NB*R rows, 1}, traversal stride R along x, 128B swizzle.
// Rows/frames beyond (Tx, Ty) are zero-filled by the TMA unit.
int make_value_tensor_map(const MasLaunch &L, int R, int W, CUtensorMap *out) {
    auto enc = tensor_map_encoder();
    if (!enc) return MAS_B200_ERR_CUDA;
    const int NB = tma_box_lanes(R, W);
    cuuint64_t gdim[3] = {(cuuint64_t)L.Ty, (cuuint64_t)L.Tx, (cuuint64_t)L.B};
    cuuint64_t gstr[2] = {(cuuint64_t)L.stride_x * 4, (cuuint64_t)(L.B > 1 ? L.stride_b : (long long)L.Tx * L.stride_x) * 4};
    cuuint32_t box[3] = {(cuuint32_t)kTileFrames, (cuuint32_t)(NB * R), 1};
    cuuint32_t estr[3] = {1, (cuuint32_t)R, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(L.value), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MAS_B200_OK : MAS_B200_ERR_ARG;
}

}  // namespace

const int *mas_start_table(void *workspace, int B, int Tx, int Ty) {
    return reinterpret_cast<const int *>(static_cast<char *>(workspace) + workspace_layout(B, Tx, Ty).start_off);
}
const int *mas_dur_table(void *workspace, int B, int Tx, int Ty, const int *user_durations) {
    return user_durations ? user_durations
                          : reinterpret_cast<const int *>(static_cast<char *>(workspace) + workspace_layout(B, Tx, Ty).dur_off);
}

Workspace workspace_layout(int B, int Tx, int Ty) {
    Workspace w{};
    auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    w.tiles = (Ty + kTileFrames - 1) / kTileFrames;
    w.rows_pitch = (int)up((size_t)Tx + 511, 512);
    w.line_pitch = (int)up((size_t)Ty, 32) + 32;
    size_t off = 0;
    w.start_off = off; off = up(off + sizeof(int) * (size_t)B * Tx, 256);
    w.dur_off = off;   off = up(off + sizeof(int) * (size_t)B * Tx, 256);
    w.gbits_off = off; off = up(off + sizeof(uint32_t) * (size_t)B * w.tiles * w.rows_pitch, 256);
    w.gline_off = off; off = up(off + sizeof(float) * (size_t)B * 2 * w.line_pitch, 256);
    w.total = off;
    return w;
}

int launch_mas(const MasLaunch &L) {
    if (!L.value || !L.t_x || !L.t_y || L.B <= 0 || L.Tx <= 0 || L.Ty <= 0) return MAS_B200_ERR_ARG;
    if (L.path_dtype != MAS_B200_PATH_NONE && L.path_dtype != MAS_B200_PATH_F32 && L.path_dtype != MAS_B200_PATH_I32)
        return MAS_B200_ERR_ARG;
    if (L.path_dtype != MAS_B200_PATH_NONE && !L.path) return MAS_B200_ERR_ARG;
    const Workspace ws = workspace_layout(L.B, L.Tx, L.Ty);
    if (!L.workspace || L.workspace_bytes < ws.total) return MAS_B200_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(L.workspace) & 255) return MAS_B200_ERR_ALIGN;

    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    Plan plan;
    if (!make_plan(L.B, L.Tx, L.Ty, di.sm_count, &plan)) return MAS_B200_ERR_UNSUPPORTED;
    const int XP = 32 * plan.R * plan.W;
    const long long txp = ((long long)L.Tx + XP - 1) / XP * XP;
    if (txp > ws.rows_pitch) return MAS_B200_ERR_UNSUPPORTED;
    // the backtrack stages one tile of direction words (rows_pitch of them) through the ring
    if ((size_t)ws.rows_pitch * 4 > sizeof(float) * (size_t)plan.NS * XP * kTilePitch) return MAS_B200_ERR_UNSUPPORTED;

    char *wsb = static_cast<char *>(L.workspace);
    MasParams P{};
    P.value = L.value; P.stride_b = L.stride_b; P.stride_x = L.stride_x;
    P.t_x = L.t_x; P.t_y = L.t_y; P.B = L.B; P.Tx = L.Tx; P.Ty = L.Ty; P.neg = L.neg;
    P.aligned = ((reinterpret_cast<uintptr_t>(L.value) & 15) == 0 && (L.stride_b & 3) == 0 && (L.stride_x & 3) == 0 &&
                 option("mas_force_unaligned") <= 0) ? 1 : 0;
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    if (P.aligned && make_value_tensor_map(L, plan.R, plan.W, &tmap) != MAS_B200_OK) P.aligned = 0;
    P.ring_stages = plan.NS;
    P.start = reinterpret_cast<int *>(wsb + ws.start_off);
    P.dur = L.durations ? L.durations : reinterpret_cast<int *>(wsb + ws.dur_off);
    P.frame_token = L.frame_token;
    P.status = L.status;
    P.gbits = reinterpret_cast<uint32_t *>(wsb + ws.gbits_off);
    P.gbits_rows_pitch = ws.rows_pitch;
    P.gbits_stride_b = (long long)ws.tiles * ws.rows_pitch;
    P.gline = reinterpret_cast<float *>(wsb + ws.gline_off);
    P.line_pitch = ws.line_pitch;
    {   // diagnostics: device pointer to a [B][8] int64 buffer smuggled through two int options
        const unsigned lo = (unsigned)option("mas_debug_ptr_lo"), hi = (unsigned)option("mas_debug_ptr_hi");
        P.dbg = reinterpret_cast<long long *>(((unsigned long long)hi << 32) | lo);
    }

    int fuse = option("mas_fused_path_write");
    if (fuse < 0) fuse = (L.B >= 2 * di.sm_count) ? 1 : 0;
    const bool want_path = L.path_dtype != MAS_B200_PATH_NONE;
    P.path = (want_path && fuse) ? L.path : nullptr;
    P.path_dtype = (want_path && fuse) ? L.path_dtype : MAS_B200_PATH_NONE;

    if (L.dry_run) P.B = 0;     // launchers only load the kernel image
    const int mode = plan.smem_bits ? MAS_MODE_SMEM_BITS : (L.Tx > XP ? MAS_MODE_MULTIPASS : MAS_MODE_GLOBAL_BITS);
    switch (plan.R) {
        case 1: rc = launch_mas_r1(P, tmap, plan.W, mode, plan.smem, L.stream); break;
        case 2: rc = launch_mas_r2(P, tmap, plan.W, mode, plan.smem, L.stream); break;
        case 4: rc = launch_mas_r4(P, tmap, plan.W, mode, plan.smem, L.stream); break;
        case 8: rc = launch_mas_r8(P, tmap, plan.W, mode, plan.smem, L.stream); break;
        default: rc = MAS_B200_ERR_UNSUPPORTED;
    }
    if (rc != MAS_B200_OK || L.dry_run) return rc;

    if (want_path && !fuse)
        return launch_path_expand(P.start, P.dur, L.B, L.Tx, L.Ty, L.path, L.path_dtype, L.stream);
    return MAS_B200_OK;
}

}  // namespace masb200
