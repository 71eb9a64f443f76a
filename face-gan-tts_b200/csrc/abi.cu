// abi.cu -- the extern "C" surface of libmas_b200.so (include/mas_b200.h).
#include <atomic>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>

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
    {"fused_impl", {0}},             // 0 auto, 1 force unfused pipeline, 2 force fused kernel
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

// The overlapped pipeline runs two kernels that wait for each other's flags; a tool that serialises kernel
// launches (ncu replay, compute-sanitizer) would starve it, so it is switched off when one is attached.
// MAS_B200_PIPELINE=serial|overlap overrides the detection.
static bool kernels_may_overlap() {
    static const bool v = [] {
        const char *e = std::getenv("MAS_B200_PIPELINE");
        if (e && std::strcmp(e, "serial") == 0) return false;
        if (e && std::strcmp(e, "overlap") == 0) return true;
        // what ncu 2025.x sets in the profiled process (scripts/ncu_env_probe.py), plus the CUDA injection hooks
        for (const char *k : {"NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NVIDIA_PROCESS_INJECTION_XML_TARGET_SETTINGS",
                              "NV_CUDA_START_SUSPENDED", "CUDA_INJECTION64_PATH", "NV_SANITIZER_INJECTION_PORT_BASE"})
            if (std::getenv(k) != nullptr) return false;
        return true;
    }();
    return v;
}

// One auxiliary stream + fork/join events per (host thread, device): the overlapped log-prior || MAS pipeline
// forks from and joins back into the caller's stream, so the call stays stream-ordered for the caller.
struct AuxStream { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static AuxStream *aux_stream() {
    static thread_local AuxStream cache[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    AuxStream &a = cache[dev];
    if (a.stream == nullptr) {
        if (cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking) != cudaSuccess) { a.stream = nullptr; return nullptr; }
        if (cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) != cudaSuccess) {
            cudaStreamDestroy(a.stream); a.stream = nullptr; return nullptr;
        }
    }
    return &a;
}

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

size_t mas_b200_workspace_bytes(int B, int Tx, int Ty) {
    if (B <= 0 || Tx <= 0 || Ty <= 0) return 0;
    return workspace_layout(B, Tx, Ty).total;
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

// flag area of the fused workspace: [B][groups][slots] group flags + [B] done flags (slots = log-prior M-tile CTAs)
static size_t fused_flag_ints(int B, int Tx, int Ty) {
    return (size_t)B * ((Ty + 63) / 64) * (size_t)((Tx + 127) / 128) + (size_t)B;
}

size_t mas_b200_fused_workspace_bytes(int B, int F, int Tx, int Ty) {
    if (B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return 0;
    // MAS workspace + the [B,Tx,Ty] value matrix (L2-resident hand-off) + per-group ready flags
    return align_up(workspace_layout(B, Tx, Ty).total, 256) + align_up(sizeof(float) * (size_t)B * Tx * Ty, 256) +
           align_up(sizeof(int) * fused_flag_ints(B, Tx, Ty), 256);
}

int mas_b200_fused_workspace_prepare(void *workspace_dev, size_t workspace_bytes, int B, int F, int Tx, int Ty, void *stream) {
    if (B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!workspace_dev || workspace_bytes < mas_b200_fused_workspace_bytes(B, F, Tx, Ty)) return MAS_B200_ERR_WORKSPACE;
    char *flags = static_cast<char *>(workspace_dev) + align_up(workspace_layout(B, Tx, Ty).total, 256) +
                  align_up(sizeof(float) * (size_t)B * Tx * Ty, 256);
    MASB200_CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int) * fused_flag_ints(B, Tx, Ty), static_cast<cudaStream_t>(stream)));
    return MAS_B200_OK;
}

// One nonce per overlapped call, process-wide and never 0: group / done flags hold the nonce of the call that set
// them, so entries left by earlier calls (all older nonces) or the zeros of a prepared workspace never look "set".
static int next_nonce() {
    static std::atomic<unsigned> counter{0};
    unsigned v;
    do { v = counter.fetch_add(1, std::memory_order_relaxed) + 1; } while ((v & 0x7fffffffu) == 0);
    return (int)(v & 0x7fffffffu);
}

int mas_b200_log_prior_maximum_path(const float *mu_x_dev, const float *y_dev, const int *t_x_dev,
                                    const int *t_y_dev, int B, int F, int Tx, int Ty, float max_neg_val,
                                    void *path_dev, int path_dtype, int *durations_dev, int *frame_token_dev,
                                    int *status_dev, void *workspace_dev, size_t workspace_bytes, int impl,
                                    void *stream) {
    if (!mu_x_dev || !y_dev || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!workspace_dev || workspace_bytes < mas_b200_fused_workspace_bytes(B, F, Tx, Ty)) return MAS_B200_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 255) return MAS_B200_ERR_ALIGN;
    const bool ws_prepared = (impl & MAS_B200_WS_PREPARED) != 0;      // flag area known clean: no memset in this call
    impl &= ~MAS_B200_WS_PREPARED;
    const size_t mas_ws = align_up(workspace_layout(B, Tx, Ty).total, 256);
    float *value = reinterpret_cast<float *>(static_cast<char *>(workspace_dev) + mas_ws);
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    // Overlapped pipeline: the tcgen05 log-prior kernel (aux stream, on the SMs the B MAS CTAs leave free)
    // publishes every 64-frame group of every utterance with a device-scope flag; the MAS kernel's TMA
    // producer acquires the flag before loading the tiles of that group.  The value matrix is handed over
    // through L2, and the alignment search starts while most of the log-prior is still being computed.
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    const int fi = option("fused_impl");
    const bool tc_ok = (impl == MAS_B200_LP_AUTO || impl == MAS_B200_LP_TCGEN05) && option("lp_impl") != MAS_B200_LP_FFMA &&
                       log_prior_tc_supported(mu_x_dev, y_dev, value, B, F, Tx, Ty);
    if (tc_ok && fi != 1 && B + log_prior_tc_min_ctas(B, F, Tx) <= di.sm_count && (fi == 2 || kernels_may_overlap())) {
        AuxStream *aux = aux_stream();
        if (aux != nullptr) {
            const int ngroups = (Ty + 63) / 64;
            int *flags = reinterpret_cast<int *>(reinterpret_cast<char *>(value) + align_up(sizeof(float) * (size_t)B * Tx * Ty, 256));
            const int slots = log_prior_tc_flag_target(F, Tx);
            int *done = flags + (size_t)B * ngroups * slots;    // [B] "table final" flags
            const int nonce = next_nonce();
            const bool want_path = path_dtype != MAS_B200_PATH_NONE && path_dev != nullptr;
            PathJob job{};
            if (want_path) {
                job.start = mas_start_table(workspace_dev, B, Tx, Ty);
                job.dur = mas_dur_table(workspace_dev, B, Tx, Ty, durations_dev);
                job.done = done; job.done_value = nonce; job.path = path_dev; job.path_dtype = path_dtype;
            }
            if (fi == 3) {
                // diagnostics only: log-prior first (serial), then the GATED MAS kernel with every flag already set --
                // isolates the cost of the gating code path from the cost of waiting for the producer
                rc = launch_log_prior_tc(mu_x_dev, y_dev, B, F, Tx, Ty, value, s);
                if (rc != MAS_B200_OK) return rc;
                MASB200_CUDA_TRY(cudaMemsetAsync(flags, 1, sizeof(int) * ((size_t)B * ngroups * slots + B), s));
                MasLaunch G{};
                G.value = value; G.stride_b = (long long)Tx * Ty; G.stride_x = Ty;
                G.t_x = t_x_dev; G.t_y = t_y_dev; G.B = B; G.Tx = Tx; G.Ty = Ty; G.neg = max_neg_val;
                G.path = path_dev; G.path_dtype = path_dtype;
                G.durations = durations_dev; G.frame_token = frame_token_dev; G.status = status_dev;
                G.workspace = workspace_dev; G.workspace_bytes = mas_ws; G.stream = s;
                G.gate = flags; G.gate_pitch = ngroups; G.gate_slots = slots; G.flag_value = 0x01010101;
                return launch_mas(G);
            }
            // A workspace the caller prepared once (mas_b200_fused_workspace_prepare) and has only used for fused
            // calls since needs no clearing: stale flags hold older nonces.  Otherwise clear it now.
            // (A CUDA-graph capture bakes this call's nonce into the kernel nodes, and every replay would then find
            // its own previous flags "set": under capture the clearing always stays in, as a node of the graph.)
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
            if (!ws_prepared || cap != cudaStreamCaptureStatusNone)
                MASB200_CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t)B * ngroups * slots + B), s));
            MASB200_CUDA_TRY(cudaEventRecord(aux->fork, s));
            MASB200_CUDA_TRY(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
            // validate the consumer's plan before anything is launched (each kernel waits for the other's flags)
            MasLaunch L{};
            L.value = value; L.stride_b = (long long)Tx * Ty; L.stride_x = Ty;
            L.t_x = t_x_dev; L.t_y = t_y_dev; L.B = B; L.Tx = Tx; L.Ty = Ty; L.neg = max_neg_val;
            L.path = path_dev; L.path_dtype = path_dtype;
            L.durations = durations_dev; L.frame_token = frame_token_dev; L.status = status_dev;
            L.workspace = workspace_dev; L.workspace_bytes = mas_ws; L.stream = s;
            L.gate = flags; L.gate_pitch = ngroups; L.gate_slots = slots; L.flag_value = nonce;
            L.done = want_path ? done : nullptr;
            L.dry_run = 1;
            rc = launch_mas(L);
            if (rc != MAS_B200_OK) return rc;
            L.dry_run = 0;
            // producer first (the block scheduler must place its CTAs before the consumer's start spinning)
            // The log-prior kernel is on the critical path (the search cannot start before its first group), so it
            // stays on the caller's stream right behind the memset; the MAS kernel takes the cross-stream hop.
            rc = launch_log_prior_tc(mu_x_dev, y_dev, B, F, Tx, Ty, value, s, flags, ngroups, di.sm_count - B,
                                     want_path ? &job : nullptr, nonce);
            L.stream = aux->stream;
            if (rc == MAS_B200_OK) rc = launch_mas(L);
            MASB200_CUDA_TRY(cudaEventRecord(aux->join, aux->stream));
            // join even on error so the aux stream never runs ahead of the caller's stream
            MASB200_CUDA_TRY(cudaStreamWaitEvent(s, aux->join, 0));
            return rc;
        }
    }
    rc = mas_b200_log_prior(mu_x_dev, y_dev, B, F, Tx, Ty, value, impl, stream);
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

// ---------------------------------------------------------------- host-buffer drop-ins
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
};
struct Stream {
    cudaStream_t s = nullptr;
    ~Stream() { if (s) cudaStreamDestroy(s); }
};
int count_bad(const int *status, int B) {
    int bad = 0;
    for (int i = 0; i < B; ++i) bad += status[i] != MAS_B200_ITEM_OK;
    return bad;
}
}  // namespace

int mas_b200_maximum_path_host(int *paths, const float *values, const int *t_xs, const int *t_ys, int B, int Tx,
                               int Ty, float max_neg_val) {
    if (!paths || !values || !t_xs || !t_ys || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    const size_t cells = (size_t)B * Tx * Ty;
    const size_t ws_bytes = mas_b200_workspace_bytes(B, Tx, Ty);
    DevBuf d_val, d_path, d_tx, d_ty, d_status, d_ws;
    Stream st;
    MASB200_CUDA_TRY(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    MASB200_CUDA_TRY(d_val.alloc(cells * 4));
    MASB200_CUDA_TRY(d_path.alloc(cells * 4));
    MASB200_CUDA_TRY(d_tx.alloc((size_t)B * 4));
    MASB200_CUDA_TRY(d_ty.alloc((size_t)B * 4));
    MASB200_CUDA_TRY(d_status.alloc((size_t)B * 4));
    MASB200_CUDA_TRY(d_ws.alloc(ws_bytes));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_val.p, values, cells * 4, cudaMemcpyHostToDevice, st.s));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_tx.p, t_xs, (size_t)B * 4, cudaMemcpyHostToDevice, st.s));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_ty.p, t_ys, (size_t)B * 4, cudaMemcpyHostToDevice, st.s));
    int rc = mas_b200_maximum_path(static_cast<float *>(d_val.p), (long long)Tx * Ty, Ty, static_cast<int *>(d_tx.p),
                                   static_cast<int *>(d_ty.p), B, Tx, Ty, max_neg_val, d_path.p, MAS_B200_PATH_I32,
                                   nullptr, nullptr, static_cast<int *>(d_status.p), d_ws.p, ws_bytes, st.s);
    if (rc != MAS_B200_OK) return rc;
    int *status = new int[B];
    cudaError_t e = cudaMemcpyAsync(paths, d_path.p, cells * 4, cudaMemcpyDeviceToHost, st.s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st.s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st.s);
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
    DevBuf d_mu, d_y, d_path, d_tx, d_ty, d_status, d_dur, d_ft, d_ws;
    Stream st;
    MASB200_CUDA_TRY(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    MASB200_CUDA_TRY(d_mu.alloc((size_t)B * F * Tx * 4));
    MASB200_CUDA_TRY(d_y.alloc((size_t)B * F * Ty * 4));
    if (paths) MASB200_CUDA_TRY(d_path.alloc(cells * 4));
    MASB200_CUDA_TRY(d_tx.alloc((size_t)B * 4));
    MASB200_CUDA_TRY(d_ty.alloc((size_t)B * 4));
    MASB200_CUDA_TRY(d_status.alloc((size_t)B * 4));
    MASB200_CUDA_TRY(d_dur.alloc((size_t)B * Tx * 4));
    MASB200_CUDA_TRY(d_ft.alloc((size_t)B * Ty * 4));
    MASB200_CUDA_TRY(d_ws.alloc(ws_bytes));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_mu.p, mu_x, (size_t)B * F * Tx * 4, cudaMemcpyHostToDevice, st.s));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_y.p, y, (size_t)B * F * Ty * 4, cudaMemcpyHostToDevice, st.s));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_tx.p, t_xs, (size_t)B * 4, cudaMemcpyHostToDevice, st.s));
    MASB200_CUDA_TRY(cudaMemcpyAsync(d_ty.p, t_ys, (size_t)B * 4, cudaMemcpyHostToDevice, st.s));
    int rc = mas_b200_log_prior_maximum_path(
        static_cast<float *>(d_mu.p), static_cast<float *>(d_y.p), static_cast<int *>(d_tx.p),
        static_cast<int *>(d_ty.p), B, F, Tx, Ty, max_neg_val, paths ? d_path.p : nullptr,
        paths ? MAS_B200_PATH_I32 : MAS_B200_PATH_NONE, static_cast<int *>(d_dur.p), static_cast<int *>(d_ft.p),
        static_cast<int *>(d_status.p), d_ws.p, ws_bytes, MAS_B200_LP_AUTO, st.s);
    if (rc != MAS_B200_OK) return rc;
    int *status = new int[B];
    cudaError_t e = cudaSuccess;
    if (paths) e = cudaMemcpyAsync(paths, d_path.p, cells * 4, cudaMemcpyDeviceToHost, st.s);
    if (e == cudaSuccess && durations)
        e = cudaMemcpyAsync(durations, d_dur.p, (size_t)B * Tx * 4, cudaMemcpyDeviceToHost, st.s);
    if (e == cudaSuccess && frame_token)
        e = cudaMemcpyAsync(frame_token, d_ft.p, (size_t)B * Ty * 4, cudaMemcpyDeviceToHost, st.s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st.s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st.s);
    if (e != cudaSuccess) { delete[] status; set_last_cuda_error(e); return MAS_B200_ERR_CUDA; }
    const int bad = count_bad(status, B);
    delete[] status;
    return bad;
}

}  // extern "C"
