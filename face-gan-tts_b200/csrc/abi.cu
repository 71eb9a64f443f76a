// abi.cu -- the extern "C" surface of libmas_b200.so (include/mas_b200.h).
#include <atomic>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "mas_host.h"

namespace masb200 {

// ------------------------------------------------------------------ errors
static thread_local int g_last_cuda_error = 0;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = (int)e; }

// ------------------------------------------------------------------ options
namespace {
struct Opt { const char *key; std::atomic<int> value; };
Opt g_opts[] = {
    {"mas_rows_per_lane", {0}},      // R: 1,2,4,8 (0 = auto)
    {"mas_dp_warps", {0}},           // W: 1..4    (0 = auto)
    {"mas_ring_stages", {0}},        // cap on NS  (0 = auto, up to 8)
    {"mas_ctas_per_sm", {0}},        // smem budget divisor (0 = auto from B)
    {"mas_force_global_bits", {0}},  // 1: direction bits always in global scratch
    {"mas_force_unaligned", {0}},    // 1: never use the TMA bulk path
    {"mas_fused_path_write", {-1}},  // -1 auto, 0 separate expand kernel, 1 in-kernel
    {"mas_debug_ptr_lo", {0}},       // diagnostics only: clock64 phase stamps buffer (device pointer halves)
    {"mas_debug_ptr_hi", {0}},
    {"lp_debug_ptr_lo", {0}},        // diagnostics only: [ctas][4] globaltimer stamps of the tcgen05 log-prior kernel
    {"lp_debug_ptr_hi", {0}},
    {"lp_debug_skip", {0}},          // diagnostics only: phases of the tcgen05 log-prior kernel to skip (results invalid)
    {"lp_impl", {0}},                // default log-prior implementation for MAS_B200_LP_AUTO
    {"upload_impl", {0}},            // 0/1 SM zero-copy pull kernel, 2 copy engine (one 2-D copy per utterance and tensor), 3 TMA bulk copies for y
    {"upload_l2_256b", {0}},         // 1: zero-copy loads carry the L2::256B fetch hint
    {"upload_ctas", {0}},            // CTAs of the zero-copy upload kernel (0 = one per SM)
    {"fused_dump_ptr_lo", {0}},      // tests only: [B,Tx,Ty] float buffer receiving the fused kernel's value tiles (device pointer halves)
    {"fused_dump_ptr_hi", {0}},
    {"pdl", {1}},                    // 1: fused kernel / path expansion launch with programmatic stream serialization
    {"fused_exp", {0}},              // diagnostics only (MASB200_PROF builds): bit 0 the DP warps ignore the readiness flags, bit 1 park the helper warps, bit 2 park the producers (results invalid)
    {"peer_dur_ptrs_lo", {0}},       // one-sided duration gather of the fused call (multi-GPU): DEVICE array of peer_world pointers, rank r's
    {"peer_dur_ptrs_hi", {0}},       //   [peer_world * B, Tx] int32 gather buffer as mapped on this device (mas_b200_set_pointer_option "peer_dur_ptrs")
    {"peer_world", {0}},             //   number of ranks (0 = off)
    {"peer_rank", {0}},              //   this rank: its durations go to rows [peer_rank * B, +B) of every buffer
    {"fused_pair", {1}},             // 1: texts of 129..256 tokens run as a 2-CTA cluster per utterance when 2B <= SMs, 0: never, 2: always
    {"fused_impl", {0}},             // 0 auto (fused kernel when the shape is covered), 1 force the serial form
};
}  // namespace

int option(const char *key) {
    for (auto &o : g_opts)
        if (std::strcmp(o.key, key) == 0) return o.value.load(std::memory_order_relaxed);
    return INT_MIN;
}

// ------------------------------------------------------------------ device
int device_info(DeviceInfo *out) {
    static std::mutex mu;
    static DeviceInfo cache[16];
    static bool have[16] = {};
    int dev = 0;
    MASB200_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 16 && have[dev]) { *out = cache[dev]; return MAS_B200_OK; }
    DeviceInfo di{};
    MASB200_CUDA_TRY(cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev));
    MASB200_CUDA_TRY(cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (dev >= 0 && dev < 16) { cache[dev] = di; have[dev] = true; }
    *out = di;
    return MAS_B200_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace masb200

using namespace masb200;

extern "C" {

int mas_b200_abi_version(void) { return MAS_B200_ABI_VERSION; }

const char *mas_b200_error_string(int status) {
    switch (status) {
        case MAS_B200_OK: return "ok";
        case MAS_B200_ERR_ARG: return "invalid argument";
        case MAS_B200_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case MAS_B200_ERR_WORKSPACE: return "workspace missing or too small";
        case MAS_B200_ERR_CUDA: return "CUDA runtime error (see mas_b200_last_cuda_error)";
        case MAS_B200_ERR_ALIGN: return "pointer alignment";
        default: return status > 0 ? "items rejected" : "unknown error";
    }
}

int mas_b200_last_cuda_error(void) { return g_last_cuda_error; }

int mas_b200_set_option(const char *key, int value) {
    if (!key) return INT_MIN;
    for (auto &o : g_opts)
        if (std::strcmp(o.key, key) == 0) return o.value.exchange(value, std::memory_order_relaxed);
    return INT_MIN;
}

int mas_b200_get_option(const char *key) { return key ? option(key) : INT_MIN; }

int mas_b200_set_pointer_option(const char *key, void *ptr) {
    if (!key) return MAS_B200_ERR_ARG;
    const std::string lo = std::string(key) + "_lo", hi = std::string(key) + "_hi";
    if (option(lo.c_str()) == INT_MIN && option(hi.c_str()) == INT_MIN) {
        // INT_MIN is also a legal half of a pointer: look the keys up instead of trusting the value
        bool found = false;
        for (auto &o : g_opts) found = found || lo == o.key;
        if (!found) return MAS_B200_ERR_ARG;
    }
    const unsigned long long v = reinterpret_cast<unsigned long long>(ptr);
    for (auto &o : g_opts) {
        if (lo == o.key) o.value.store((int)(unsigned)(v & 0xffffffffu), std::memory_order_relaxed);
        if (hi == o.key) o.value.store((int)(unsigned)(v >> 32), std::memory_order_relaxed);
    }
    return MAS_B200_OK;
}

size_t mas_b200_workspace_bytes(int B, int Tx, int Ty) {
    if (B <= 0 || Tx <= 0 || Ty <= 0) return 0;
    return workspace_layout(B, Tx, Ty).total;
}

int mas_b200_debug_occupy_sms(int ctas, long long cycles, void *stream) {
    return launch_debug_spin(ctas, cycles, static_cast<cudaStream_t>(stream));
}

int mas_b200_lengths_from_mask(const float *mask_dev, int B, int Tx, int Ty, int *t_x_dev, int *t_y_dev,
                               void *stream) {
    return launch_lengths_from_mask(mask_dev, B, Tx, Ty, t_x_dev, t_y_dev, static_cast<cudaStream_t>(stream));
}

int mas_b200_maximum_path(const float *value_dev, long long stride_b, long long stride_x, const int *t_x_dev,
                          const int *t_y_dev, int B, int Tx, int Ty, float max_neg_val, void *path_dev,
                          int path_dtype, int *durations_dev, int *frame_token_dev, int *status_dev,
                          void *workspace_dev, size_t workspace_bytes, void *stream) {
    MasLaunch L{};
    L.value = value_dev; L.stride_b = stride_b; L.stride_x = stride_x;
    L.t_x = t_x_dev; L.t_y = t_y_dev; L.B = B; L.Tx = Tx; L.Ty = Ty; L.neg = max_neg_val;
    L.path = path_dev; L.path_dtype = path_dtype;
    L.durations = durations_dev; L.frame_token = frame_token_dev; L.status = status_dev;
    L.workspace = workspace_dev; L.workspace_bytes = workspace_bytes;
    L.stream = static_cast<cudaStream_t>(stream);
    if (stride_x < Ty || stride_b < 0) return MAS_B200_ERR_ARG;
    return launch_mas(L);
}

int mas_b200_log_prior(const float *mu_x_dev, const float *y_dev, int B, int F, int Tx, int Ty,
                       float *log_prior_dev, int impl, void *stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (impl == MAS_B200_LP_AUTO) {
        const int o = option("lp_impl");
        impl = (o == MAS_B200_LP_FFMA || o == MAS_B200_LP_TCGEN05) ? o : MAS_B200_LP_AUTO;
    }
    if (impl == MAS_B200_LP_FFMA) return launch_log_prior_ffma(mu_x_dev, y_dev, B, F, Tx, Ty, log_prior_dev, s);
    if (impl == MAS_B200_LP_TCGEN05) return launch_log_prior_tc(mu_x_dev, y_dev, B, F, Tx, Ty, log_prior_dev, s);
    if (impl != MAS_B200_LP_AUTO) return MAS_B200_ERR_ARG;
    const int rc = launch_log_prior_tc(mu_x_dev, y_dev, B, F, Tx, Ty, log_prior_dev, s);
    if (rc != MAS_B200_ERR_UNSUPPORTED) return rc;
    return launch_log_prior_ffma(mu_x_dev, y_dev, B, F, Tx, Ty, log_prior_dev, s);
}

size_t mas_b200_fused_workspace_bytes(int B, int F, int Tx, int Ty) {
    if (B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return 0;
    // MAS workspace + the [B,Tx,Ty] value matrix of the serial form (shapes the fused kernel does not cover)
    return align_up(workspace_layout(B, Tx, Ty).total, 256) + align_up(sizeof(float) * (size_t)B * Tx * Ty, 256);
}

int mas_b200_fused_workspace_prepare(void *workspace_dev, size_t workspace_bytes, int B, int F, int Tx, int Ty, void *stream) {
    // kept for ABI compatibility: the fused call keeps no state in the workspace between calls any more
    (void)stream;
    if (B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!workspace_dev || workspace_bytes < mas_b200_fused_workspace_bytes(B, F, Tx, Ty)) return MAS_B200_ERR_WORKSPACE;
    return MAS_B200_OK;
}

int mas_b200_log_prior_maximum_path(const float *mu_x_dev, const float *y_dev, const int *t_x_dev,
                                    const int *t_y_dev, int B, int F, int Tx, int Ty, float max_neg_val,
                                    void *path_dev, int path_dtype, int *durations_dev, int *frame_token_dev,
                                    int *status_dev, void *workspace_dev, size_t workspace_bytes, int impl,
                                    void *stream) {
    if (!mu_x_dev || !y_dev || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!workspace_dev || workspace_bytes < mas_b200_fused_workspace_bytes(B, F, Tx, Ty)) return MAS_B200_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 255) return MAS_B200_ERR_ALIGN;
    impl &= ~MAS_B200_WS_PREPARED;                                  // accepted for compatibility, no effect
    const size_t mas_ws = align_up(workspace_layout(B, Tx, Ty).total, 256);
    float *value = reinterpret_cast<float *>(static_cast<char *>(workspace_dev) + mas_ws);

    // One kernel -- one CTA, or a 2-CTA cluster, per utterance: the tcgen05 epilogue writes the value tiles into the
    // shared-memory ring of the alignment search (lp_mas_fused.cu).  Shapes it does not cover (texts longer than 256
    // tokens, n_feats = 96 / 128 with more than 128 tokens in a batch too large for the pair form, very long utterances)
    // run the serial form: log-prior kernel -> [B,Tx,Ty] in the workspace -> MAS kernel.
    const bool tc_wanted = (impl == MAS_B200_LP_AUTO || impl == MAS_B200_LP_TCGEN05) && option("lp_impl") != MAS_B200_LP_FFMA;
    if (tc_wanted && option("fused_impl") != 1 && lp_mas_fused_supported(mu_x_dev, y_dev, B, F, Tx, Ty)) {
        const int rc = launch_lp_mas_fused(mu_x_dev, y_dev, t_x_dev, t_y_dev, B, F, Tx, Ty, max_neg_val, path_dev, path_dtype,
                                           durations_dev, frame_token_dev, status_dev, workspace_dev, mas_ws,
                                           static_cast<cudaStream_t>(stream));
        if (rc != MAS_B200_ERR_UNSUPPORTED) return rc;      // (a refused cluster launch without a one-CTA form: serial form)
    }
    int rc = mas_b200_log_prior(mu_x_dev, y_dev, B, F, Tx, Ty, value, impl, stream);
    if (rc != MAS_B200_OK) return rc;
    return mas_b200_maximum_path(value, (long long)Tx * Ty, Ty, t_x_dev, t_y_dev, B, Tx, Ty, max_neg_val, path_dev,
                                 path_dtype, durations_dev, frame_token_dev, status_dev, workspace_dev, mas_ws, stream);
}

int mas_b200_generate_path(const int *durations_dev, const int *t_x_dev, const int *t_y_dev, int B, int Tx, int Ty,
                           void *path_dev, int path_dtype, void *stream) {
    if (path_dtype == MAS_B200_PATH_NONE) return MAS_B200_ERR_ARG;
    return launch_generate_path(durations_dev, 0, t_x_dev, t_y_dev, B, Tx, Ty, path_dev, path_dtype, nullptr,
                                static_cast<cudaStream_t>(stream));
}

int mas_b200_generate_path_f32(const float *durations_dev, const int *t_x_dev, const int *t_y_dev, int B, int Tx, int Ty,
                               void *path_dev, int path_dtype, int *frame_token_dev, void *stream) {
    return launch_generate_path(durations_dev, 1, t_x_dev, t_y_dev, B, Tx, Ty, path_dev, path_dtype, frame_token_dev,
                                static_cast<cudaStream_t>(stream));
}

int mas_b200_put_durations(const int *durations_dev, int B, int Tx, void *const *peer_ptrs_dev, int world, int rank,
                           void *stream) {
    return launch_put_rows(durations_dev, B, Tx, peer_ptrs_dev, world, rank, static_cast<cudaStream_t>(stream));
}

int mas_b200_sequence_mask(const int *lengths_dev, int B, int T, float *mask_dev, void *stream) {
    return launch_sequence_mask(lengths_dev, B, T, mask_dev, static_cast<cudaStream_t>(stream));
}

int mas_b200_crop_frames(const float *y_dev, const int *frame_token_dev, const int *y_lengths_dev,
                         const int *offsets_dev, int B, int F, int Ty, int out_size, float *y_cut_dev,
                         int *frame_token_cut_dev, int *cut_lengths_dev, float *cut_mask_dev, void *stream) {
    return launch_crop_frames(y_dev, frame_token_dev, y_lengths_dev, offsets_dev, B, F, Ty, out_size, y_cut_dev,
                              frame_token_cut_dev, cut_lengths_dev, cut_mask_dev, static_cast<cudaStream_t>(stream));
}

int mas_b200_gather_mu_y(const float *mu_x_dev, const int *frame_token_dev, int B, int F, int Tx, int Ty,
                         float *mu_y_dev, void *stream) {
    return launch_gather_mu_y(mu_x_dev, frame_token_dev, B, F, Tx, Ty, mu_y_dev, static_cast<cudaStream_t>(stream));
}

int mas_b200_gather_mu_y_backward(const float *grad_mu_y_dev, const int *start_dev, const int *durations_dev,
                                  const int *offsets_dev, const int *lengths_dev, int B, int F, int Tx, int Ty,
                                  float *grad_mu_x_dev, void *stream) {
    return launch_gather_mu_y_bwd(grad_mu_y_dev, start_dev, durations_dev, offsets_dev, lengths_dev, B, F, Tx, Ty,
                                  grad_mu_x_dev, static_cast<cudaStream_t>(stream));
}

size_t mas_b200_prior_loss_workspace_bytes(int B, int F, int Ty) { return prior_loss_workspace_bytes(B, F, Ty); }

int mas_b200_prior_loss(const float *y_dev, const float *mu_x_dev, const int *frame_token_dev,
                        const int *y_lengths_dev, int B, int F, int Tx, int Ty, float *mu_y_dev, float *loss_dev,
                        void *workspace_dev, size_t workspace_bytes, void *stream) {
    return launch_prior_loss(y_dev, mu_x_dev, frame_token_dev, y_lengths_dev, B, F, Tx, Ty, mu_y_dev, loss_dev,
                             workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mas_b200_prior_loss_backward(const float *y_dev, const float *mu_x_dev, const int *start_dev,
                                 const int *durations_dev, const int *offsets_dev, const int *y_lengths_dev,
                                 const float *grad_loss_dev, int B, int F, int Tx, int Ty, float *grad_mu_x_dev,
                                 void *stream) {
    return launch_prior_loss_bwd(y_dev, mu_x_dev, start_dev, durations_dev, offsets_dev, y_lengths_dev, grad_loss_dev,
                                 B, F, Tx, Ty, grad_mu_x_dev, static_cast<cudaStream_t>(stream));
}

int mas_b200_duration_loss(const float *logw_dev, const int *durations_dev, const int *x_lengths_dev, int B, int Tx,
                           float *loss_dev, float *logw_target_dev, float *grad_logw_dev, void *stream) {
    return launch_duration_loss(logw_dev, durations_dev, x_lengths_dev, B, Tx, loss_dev, logw_target_dev, grad_logw_dev,
                                static_cast<cudaStream_t>(stream));
}

int mas_b200_upload_batch(const float *mu_x_pinned, const float *y_pinned, const int *t_xs_pinned,
                          const int *t_ys_pinned, int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev,
                          int *t_x_dev, int *t_y_dev, void *stream) {
    return launch_upload_batch(mu_x_pinned, y_pinned, t_xs_pinned, t_ys_pinned, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev,
                               t_y_dev, static_cast<cudaStream_t>(stream));
}

size_t mas_b200_packed_batch_bytes(const int *t_xs, const int *t_ys, int B, int F) {
    if (!t_xs || !t_ys || B <= 0 || F <= 0) return 0;
    size_t n = 0;
    for (int b = 0; b < B; ++b) n += (size_t)(t_xs[b] > 0 ? t_xs[b] : 0) + (size_t)(t_ys[b] > 0 ? t_ys[b] : 0);
    return packed_batch_header_bytes(B) + sizeof(float) * (size_t)F * n;
}

int mas_b200_pack_batch_host(const float *mu_x, const float *y, const int *t_xs, const int *t_ys, int B, int F, int Tx,
                             int Ty, void *packed, size_t packed_bytes) {
    if (!mu_x || !y || !t_xs || !t_ys || !packed || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    for (int b = 0; b < B; ++b)
        if (t_xs[b] < 0 || t_xs[b] > Tx || t_ys[b] < 0 || t_ys[b] > Ty) return MAS_B200_ERR_ARG;
    if (packed_bytes < mas_b200_packed_batch_bytes(t_xs, t_ys, B, F)) return MAS_B200_ERR_WORKSPACE;
    int *hdr = static_cast<int *>(packed);
    std::memcpy(hdr, t_xs, sizeof(int) * (size_t)B);
    std::memcpy(hdr + B, t_ys, sizeof(int) * (size_t)B);
    float *dst = reinterpret_cast<float *>(static_cast<char *>(packed) + packed_batch_header_bytes(B));
    for (int b = 0; b < B; ++b)
        for (int f = 0; f < F; ++f) {
            std::memcpy(dst, mu_x + ((size_t)b * F + f) * Tx, sizeof(float) * (size_t)t_xs[b]);
            dst += t_xs[b];
        }
    for (int b = 0; b < B; ++b)
        for (int f = 0; f < F; ++f) {
            std::memcpy(dst, y + ((size_t)b * F + f) * Ty, sizeof(float) * (size_t)t_ys[b]);
            dst += t_ys[b];
        }
    return MAS_B200_OK;
}

int mas_b200_unpack_batch(const void *packed_dev, int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev,
                          int *t_x_dev, int *t_y_dev, void *stream) {
    return launch_unpack_batch(packed_dev, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev, t_y_dev, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- host-buffer drop-ins
namespace {
// One device arena + one stream per (host thread, device), grown on demand and reused: the host-buffer drop-ins are
// called once per training step, and six cudaMalloc / cudaFree pairs per call cost more than the alignment itself.
struct HostArena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    cudaStream_t stream = nullptr;
    ~HostArena() {
        if (base) cudaFree(base);
        if (stream) cudaStreamDestroy(stream);
    }
    cudaError_t reserve(size_t bytes) {
        used = 0;
        if (!stream) {
            cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
            if (e != cudaSuccess) { stream = nullptr; return e; }
        }
        if (bytes <= cap) return cudaSuccess;
        if (base) { cudaFree(base); base = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&base), bytes);
        if (e != cudaSuccess) { base = nullptr; return e; }
        cap = bytes;
        return cudaSuccess;
    }
    void *take(size_t bytes) {
        char *p = base + used;
        used += align_up(bytes ? bytes : 1, 256);
        return p;
    }
};
HostArena *host_arena() {
    static thread_local HostArena cache[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    return &cache[dev];
}
int count_bad(const int *status, int B) {
    int bad = 0;
    for (int i = 0; i < B; ++i) bad += status[i] != MAS_B200_ITEM_OK;
    return bad;
}
size_t pad256(size_t n) { return align_up(n ? n : 1, 256); }
}  // namespace

int mas_b200_maximum_path_host(int *paths, const float *values, const int *t_xs, const int *t_ys, int B, int Tx,
                               int Ty, float max_neg_val) {
    if (!paths || !values || !t_xs || !t_ys || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    const size_t cells = (size_t)B * Tx * Ty;
    const size_t ws_bytes = mas_b200_workspace_bytes(B, Tx, Ty);
    HostArena *A = host_arena();
    if (!A) return MAS_B200_ERR_CUDA;
    MASB200_CUDA_TRY(A->reserve(2 * pad256(cells * 4) + 3 * pad256((size_t)B * 4) + pad256(ws_bytes)));
    void *d_ws = A->take(ws_bytes);                 // first: 256-byte aligned like the arena itself
    float *d_val = static_cast<float *>(A->take(cells * 4));
    void *d_path = A->take(cells * 4);
    int *d_tx = static_cast<int *>(A->take((size_t)B * 4)), *d_ty = static_cast<int *>(A->take((size_t)B * 4));
    int *d_status = static_cast<int *>(A->take((size_t)B * 4));
    cudaStream_t st = A->stream;
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_val, values, cells * 4, cudaMemcpyHostToDevice, st));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_tx, t_xs, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_ty, t_ys, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    int rc = mas_b200_maximum_path(d_val, (long long)Tx * Ty, Ty, d_tx, d_ty, B, Tx, Ty, max_neg_val, d_path,
                                   MAS_B200_PATH_I32, nullptr, nullptr, d_status, d_ws, ws_bytes, st);
    if (rc != MAS_B200_OK) { cudaStreamSynchronize(st); return rc; }
    int *status = new int[B];
    cudaError_t e = cudaMemcpyAsync(paths, d_path, cells * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete[] status; set_last_cuda_error(e); return MAS_B200_ERR_CUDA; }
    const int bad = count_bad(status, B);
    delete[] status;
    return bad;
}

int mas_b200_log_prior_maximum_path_host(const float *mu_x, const float *y, const int *t_xs, const int *t_ys,
                                         int B, int F, int Tx, int Ty, float max_neg_val, int *paths,
                                         int *durations, int *frame_token) {
    if (!mu_x || !y || !t_xs || !t_ys || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    const size_t cells = (size_t)B * Tx * Ty;
    const size_t ws_bytes = mas_b200_fused_workspace_bytes(B, F, Tx, Ty);
    HostArena *A = host_arena();
    if (!A) return MAS_B200_ERR_CUDA;
    MASB200_CUDA_TRY(A->reserve(pad256(ws_bytes) + pad256((size_t)B * F * Tx * 4) + pad256((size_t)B * F * Ty * 4) +
                                (paths ? pad256(cells * 4) : 0) + 3 * pad256((size_t)B * 4) + pad256((size_t)B * Tx * 4) +
                                pad256((size_t)B * Ty * 4)));
    void *d_ws = A->take(ws_bytes);
    float *d_mu = static_cast<float *>(A->take((size_t)B * F * Tx * 4)), *d_y = static_cast<float *>(A->take((size_t)B * F * Ty * 4));
    void *d_path = paths ? A->take(cells * 4) : nullptr;
    int *d_tx = static_cast<int *>(A->take((size_t)B * 4)), *d_ty = static_cast<int *>(A->take((size_t)B * 4));
    int *d_status = static_cast<int *>(A->take((size_t)B * 4));
    int *d_dur = static_cast<int *>(A->take((size_t)B * Tx * 4)), *d_ft = static_cast<int *>(A->take((size_t)B * Ty * 4));
    cudaStream_t st = A->stream;
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_mu, mu_x, (size_t)B * F * Tx * 4, cudaMemcpyHostToDevice, st));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_y, y, (size_t)B * F * Ty * 4, cudaMemcpyHostToDevice, st));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_tx, t_xs, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_ty, t_ys, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    int rc = mas_b200_log_prior_maximum_path(d_mu, d_y, d_tx, d_ty, B, F, Tx, Ty, max_neg_val, d_path,
                                             paths ? MAS_B200_PATH_I32 : MAS_B200_PATH_NONE, d_dur, d_ft, d_status, d_ws,
                                             ws_bytes, MAS_B200_LP_AUTO, st);
    if (rc != MAS_B200_OK) { cudaStreamSynchronize(st); return rc; }
    int *status = new int[B];
    cudaError_t e = cudaSuccess;
    if (paths) e = cudaMemcpyAsync(paths, d_path, cells * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && durations) e = cudaMemcpyAsync(durations, d_dur, (size_t)B * Tx * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && frame_token) e = cudaMemcpyAsync(frame_token, d_ft, (size_t)B * Ty * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete[] status; set_last_cuda_error(e); return MAS_B200_ERR_CUDA; }
    const int bad = count_bad(status, B);
    delete[] status;
    return bad;
}

}  // extern "C"
