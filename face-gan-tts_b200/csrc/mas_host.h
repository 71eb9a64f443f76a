// mas_host.h -- host-side declarations shared by the translation units of libmas_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mas_b200.h"

namespace masb200 {

// ---- error plumbing -------------------------------------------------------
void set_last_cuda_error(cudaError_t e);
#define MASB200_CUDA_TRY(expr)                                   \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) {                                 \
            ::masb200::set_last_cuda_error(_e);                  \
            return MAS_B200_ERR_CUDA;                            \
        }                                                        \
    } while (0)

// ---- options (process-wide tuning knobs, mas_b200_set_option) --------------
int option(const char *key);   // INT32_MIN if unknown

// ---- device properties cached per device ----------------------------------
struct DeviceInfo {
    int sm_count;
    int max_smem_optin;
};
int device_info(DeviceInfo *out);

// ---- workspace layout -------------------------------------------------------
// [start: B*Tx i32][dur: B*Tx i32][gbits: B*tiles*rows_pitch u32][gline: B*2*line_pitch f32]
struct Workspace {
    size_t start_off, dur_off, gbits_off, gline_off, total;
    int tiles, rows_pitch, line_pitch;
};
Workspace workspace_layout(int B, int Tx, int Ty);

// ---- MAS kernel flavours (template arguments SMEM_BITS, MULTIPASS of mas_forward_kernel) ----
enum { MAS_MODE_SMEM_BITS = 0, MAS_MODE_GLOBAL_BITS = 1, MAS_MODE_MULTIPASS = 2 };

// ---- launchers ----------------------------------------------------------------
struct MasLaunch {
    const float *value;
    long long stride_b, stride_x;
    const int *t_x, *t_y;
    int B, Tx, Ty;
    float neg;
    void *path;
    int path_dtype;
    int *durations;     // may be null
    int *frame_token;   // may be null
    int *status;        // may be null
    void *workspace;
    size_t workspace_bytes;
    cudaStream_t stream;
    int dry_run;        // 1: validate arguments / plan only, launch nothing
};

const int *mas_start_table(void *workspace, int B, int Tx, int Ty);
const int *mas_dur_table(void *workspace, int B, int Tx, int Ty, const int *user_durations);
int launch_mas(const MasLaunch &L);

int launch_put_rows(const int *src, int rows, int cols, void *const *peer_ptrs_dev, int world, int rank, cudaStream_t stream);
int launch_path_expand(const int *start, const int *dur, int B, int Tx, int Ty, void *path, int path_dtype,
                       cudaStream_t stream);
int launch_debug_spin(int ctas, long long cycles, cudaStream_t stream);   // tests only (path_ops.cu)
int launch_lengths_from_mask(const float *mask, int B, int Tx, int Ty, int *t_x, int *t_y, cudaStream_t stream);
int launch_generate_path(const void *durations, int dur_is_float, const int *t_x, const int *t_y, int B, int Tx, int Ty,
                         void *path, int path_dtype, int *frame_token, cudaStream_t stream);
// loss_ops.cu: consumers of the alignment in index form
int launch_sequence_mask(const int *lengths, int B, int T, float *mask, cudaStream_t stream);
int launch_crop_frames(const float *y, const int *frame_token, const int *y_lengths, const int *offsets, int B, int F,
                       int Ty, int out_size, float *y_cut, int *ft_cut, int *cut_lengths, float *cut_mask,
                       cudaStream_t stream);
size_t prior_loss_workspace_bytes(int B, int F, int Ty);
int launch_gather_mu_y(const float *mu_x, const int *frame_token, int B, int F, int Tx, int Ty, float *mu_y,
                       cudaStream_t stream);
int launch_gather_mu_y_bwd(const float *grad_mu_y, const int *start, const int *dur, const int *offsets,
                           const int *lengths, int B, int F, int Tx, int Ty, float *grad_mu_x, cudaStream_t stream);
int launch_prior_loss(const float *y, const float *mu_x, const int *frame_token, const int *y_lengths, int B, int F,
                      int Tx, int Ty, float *mu_y, float *loss, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream);
int launch_prior_loss_bwd(const float *y, const float *mu_x, const int *start, const int *dur, const int *offsets,
                          const int *y_lengths, const float *grad_loss, int B, int F, int Tx, int Ty, float *grad_mu_x,
                          cudaStream_t stream);
int launch_duration_loss(const float *logw, const int *durations, const int *x_lengths, int B, int Tx, float *loss,
                         float *logw_target, float *grad_logw, cudaStream_t stream);
// upload.cu: ragged zero-copy pull of a padded host batch (page-locked memory)
int launch_upload_batch(const float *mu_x_pinned, const float *y_pinned, const int *t_xs_pinned, const int *t_ys_pinned,
                        int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev, int *t_x_dev, int *t_y_dev,
                        cudaStream_t stream);
// packed (ragged) batch -> zero-padded device tensors (upload.cu)
size_t packed_batch_header_bytes(int B);
int launch_unpack_batch(const void *packed_dev, int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev, int *t_x_dev,
                        int *t_y_dev, cudaStream_t stream);
int launch_log_prior_ffma(const float *mu_x, const float *y, int B, int F, int Tx, int Ty, float *out,
                          cudaStream_t stream);
// tcgen05 implementation; returns MAS_B200_ERR_UNSUPPORTED when the shape is not covered.
int launch_log_prior_tc(const float *mu_x, const float *y, int B, int F, int Tx, int Ty, float *out,
                        cudaStream_t stream);
// lp_mas_fused.cu: one CTA per utterance, value tiles handed over in shared memory (F in {64,80,96}, Tx <= 256)
bool lp_mas_fused_supported(const float *mu_x, const float *y, int B, int F, int Tx, int Ty);
int launch_lp_mas_fused(const float *mu_x, const float *y, const int *t_x, const int *t_y, int B, int F, int Tx, int Ty,
                        float neg, void *path, int path_dtype, int *durations, int *frame_token, int *status,
                        void *workspace, size_t workspace_bytes, cudaStream_t stream);
bool log_prior_tc_supported(const float *mu_x, const float *y, const float *out, int B, int F, int Tx, int Ty);

}  // namespace masb200
