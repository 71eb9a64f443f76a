// loss_ops.cu -- the consumers of the alignment in FaceTTS.compute_loss, driven by the INDEX form of the path
// (durations / start / frame_token emitted by the backtrack) instead of the dense [B,Tx,Ty] tensor:
//   sequence_mask        lengths -> 0/1 mask                                   (reference model/utils.py:6-11)
//   crop_frames          random out_size-frame window of y and of the path     (model/face_tts.py:181-215)
//   gather_mu_y (+bwd)   mu_y = attn^T @ mu_x^T as a gather / segmented sum     (model/face_tts.py:217-218)
//   prior_loss (+bwd)    0.5*((y-mu_y)^2+log 2pi) masked mean, fused with the gather (model/face_tts.py:233-234)
//   duration_loss        logw_ = log(1e-8+dur)*x_mask, sum((logw-logw_)^2)/sum(len)  (face_tts.py:176-179, utils.py:43-45)
// All of them are small streaming kernels (HBM/L2 bound, a few hundred KB to a few MB per call); reductions are
// two-stage with a fixed order (no atomics), so every result is deterministic run to run.
#include "mas_common.cuh"
#include "mas_host.h"

namespace masb200 {

namespace {

constexpr int kF = 8;            // feature rows per CTA (grid.y tiles F)
constexpr int kThreads = 256;    // frames (or tokens) per CTA
constexpr float kHalfLog2Pi = 0.91893853320467274178f;   // 0.5*log(2*pi)

__device__ __forceinline__ double block_sum(double v, double *red /*[kThreads/32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    return s;   // valid in thread 0
}

__global__ void __launch_bounds__(kThreads) sequence_mask_kernel(const int *__restrict__ len, int T,
                                                                 float *__restrict__ mask) {
    const int b = blockIdx.y, t = blockIdx.x * kThreads + threadIdx.x;
    if (t < T) mask[(size_t)b * T + t] = (t < len[b]) ? 1.f : 0.f;
}

// y_cut[b,f,t'] = y[b,f,off+t'], ft_cut[b,t'] = ft[b,off+t'] for t' < cut_len = min(y_len, out_size); 0 / -1 beyond.
__global__ void __launch_bounds__(kThreads) crop_frames_kernel(const float *__restrict__ y, const int *__restrict__ ft,
                                                               const int *__restrict__ y_len, const int *__restrict__ off,
                                                               int F, int Ty, int out_size, float *__restrict__ y_cut,
                                                               int *__restrict__ ft_cut, int *__restrict__ cut_len,
                                                               float *__restrict__ cut_mask) {
    const int b = blockIdx.z, f0 = blockIdx.y * kF;
    const int t = blockIdx.x * kThreads + threadIdx.x;
    const int len = min(max(y_len[b], 0), min(out_size, Ty));
    const int o = min(max(off[b], 0), Ty - len);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && cut_len) cut_len[b] = len;
    if (t >= out_size) return;
    const bool in = t < len;
    if (blockIdx.y == 0) {
        if (ft_cut) ft_cut[(size_t)b * out_size + t] = in ? ft[(size_t)b * Ty + o + t] : -1;
        if (cut_mask) cut_mask[(size_t)b * out_size + t] = in ? 1.f : 0.f;
    }
    const float *yb = y + ((size_t)b * F + f0) * Ty + o + t;
    float *ob = y_cut + ((size_t)b * F + f0) * out_size + t;
#pragma unroll
    for (int k = 0; k < kF; ++k)
        if (f0 + k < F) ob[(size_t)k * out_size] = in ? __ldg(yb + (size_t)k * Ty) : 0.f;
}

// mu_y[b,f,t] = mu_x[b,f,ft[b,t]] (0 where ft < 0).  With y != nullptr it also accumulates the prior-loss
// numerator sum_{t < len} 0.5*((y-mu_y)^2 + log 2pi) into one double per CTA (second stage below).
template <bool kLoss>
__global__ void __launch_bounds__(kThreads) gather_mu_y_kernel(const float *__restrict__ mu_x, const int *__restrict__ ft,
                                                               const float *__restrict__ y, const int *__restrict__ len,
                                                               int F, int Tx, int Ty, float *__restrict__ mu_y,
                                                               double *__restrict__ partial) {
    const int b = blockIdx.z, f0 = blockIdx.y * kF;
    const int t = blockIdx.x * kThreads + threadIdx.x;
    double acc = 0.0;
    if (t < Ty) {
        const int idx = ft[(size_t)b * Ty + t];
        const bool hit = idx >= 0 && idx < Tx;
        const bool valid = kLoss && t < len[b];
        const float *mb = mu_x + ((size_t)b * F + f0) * Tx;
        const size_t o = ((size_t)b * F + f0) * Ty + t;
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kF; ++k) {
            if (f0 + k >= F) break;
            const float m = hit ? __ldg(mb + (size_t)k * Tx + idx) : 0.f;
            if (mu_y) mu_y[o + (size_t)k * Ty] = m;
            if (kLoss && valid) {
                const float d = __ldg(y + o + (size_t)k * Ty) - m;
                a += 0.5f * (d * d + 2.f * kHalfLog2Pi);
            }
        }
        acc = (double)a;
    }
    if (kLoss) {
        __shared__ double red[kThreads / 32];
        const double s = block_sum(acc, red);
        if (threadIdx.x == 0) partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

// loss = sum(partials) / (sum_b clamp(len[b]) * F)   -- one CTA, fixed order
__global__ void __launch_bounds__(kThreads) prior_loss_finish_kernel(const double *__restrict__ partial, int n,
                                                                     const int *__restrict__ len, int B, int Ty, int F,
                                                                     float *__restrict__ loss) {
    __shared__ double red[kThreads / 32];
    double s = 0.0, l = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) s += partial[i];
    for (int i = threadIdx.x; i < B; i += kThreads) l += (double)min(max(len[i], 0), Ty);
    const double st = block_sum(s, red);
    __syncthreads();
    const double lt = block_sum(l, red);
    if (threadIdx.x == 0) loss[0] = (float)(st / (lt * (double)F));
}

// grad_mu_x[b,f,x] = sum over the frames token x owns of grad_mu_y[b,f,t].  A token's frames are contiguous
// ([start, start+dur), shifted by the crop window when there is one): a segmented sum in a fixed order.
// One WARP per token (lanes stride the token's frames, so a 200-frame silence token costs 7 coalesced passes instead
// of one thread looping 200 times), kF feature rows per warp, xor-shuffle tree at the end: the order of the
// additions is fixed by (lane, tree), hence deterministic.
// kPrior: instead of reading grad_mu_y, the summand is the prior-loss derivative  -(y - mu_x[x]) * g / denom.
constexpr int kSegWarps = kThreads / 32;      // tokens per CTA
template <bool kPrior>
__global__ void __launch_bounds__(kThreads) segment_sum_kernel(const float *__restrict__ g, const float *__restrict__ mu_x,
                                                               const int *__restrict__ start, const int *__restrict__ dur,
                                                               const int *__restrict__ off, const int *__restrict__ len,
                                                               const float *__restrict__ gscale, int B, int F, int Tx,
                                                               int Ty, float *__restrict__ gx) {
    const int b = blockIdx.z, f0 = blockIdx.y * kF;
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * kSegWarps + (threadIdx.x >> 5);
    if (x >= Tx) return;                                   // warp-uniform
    const int o = off ? off[b] : 0;
    const int hi = len ? min(max(len[b], 0), Ty) : Ty;
    const int s0 = start[(size_t)b * Tx + x] - o;
    const int e = min(s0 + dur[(size_t)b * Tx + x], hi);
    const int s = max(s0, 0);
    const float *gb = g + ((size_t)b * F + f0) * Ty;
    float scale = 1.f;
    if (kPrior) {
        double l = 0.0;      // sum of the (clamped) lengths: B/32 additions per lane + a shuffle tree
        for (int i = lane; i < B; i += 32) l += (double)min(max(len ? len[i] : Ty, 0), Ty);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) l += __shfl_xor_sync(kFullMask, l, d);
        scale = -gscale[0] / (float)(l * (double)F);
    }
    float acc[kF], m[kF];
#pragma unroll
    for (int k = 0; k < kF; ++k) {
        acc[k] = 0.f;
        m[k] = (kPrior && f0 + k < F) ? __ldg(mu_x + ((size_t)b * F + f0 + k) * Tx + x) : 0.f;
    }
    for (int t = s + lane; t < e; t += 32) {
#pragma unroll
        for (int k = 0; k < kF; ++k)
            if (f0 + k < F) acc[k] += __ldg(gb + (size_t)k * Ty + t) - m[k];
    }
#pragma unroll
    for (int k = 0; k < kF; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[k] += __shfl_xor_sync(kFullMask, acc[k], d);
    }
    if (lane < kF && f0 + lane < F) {
        float v = acc[0];
#pragma unroll
        for (int k = 1; k < kF; ++k) v = (lane == k) ? acc[k] : v;
        gx[((size_t)b * F + f0 + lane) * Tx + x] = v * scale;
    }
}

// One CTA: loss = sum_{b,x} (logw - logw_)^2 / sum_b x_len,  logw_ = log(1e-8 + dur) * (x < x_len).
__global__ void __launch_bounds__(1024) duration_loss_kernel(const float *__restrict__ logw, const int *__restrict__ dur,
                                                             const int *__restrict__ x_len, int B, int Tx,
                                                             float *__restrict__ loss, float *__restrict__ logw_target,
                                                             float *__restrict__ grad_logw) {
    __shared__ double red[32];
    __shared__ double denom_s;
    double l = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) l += (double)x_len[i];
    const double lt = block_sum(l, red);
    if (threadIdx.x == 0) denom_s = lt;
    __syncthreads();
    const float inv = (float)(1.0 / denom_s);
    double acc = 0.0;
    const int n = B * Tx;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int b = i / Tx, x = i - b * Tx;
        const float tgt = (x < x_len[b]) ? logf(1e-8f + (float)dur[i]) : 0.f;
        const float d = logw[i] - tgt;
        if (logw_target) logw_target[i] = tgt;
        if (grad_logw) grad_logw[i] = 2.f * d * inv;
        acc += (double)(d * d);
    }
    __syncthreads();
    const double st = block_sum(acc, red);
    if (threadIdx.x == 0) loss[0] = (float)(st / denom_s);
}

inline bool grid_ok(int B, int F) { return B <= 65535 && (F + kF - 1) / kF <= 65535; }

}  // namespace

int launch_sequence_mask(const int *lengths, int B, int T, float *mask, cudaStream_t stream) {
    if (!lengths || !mask || B <= 0 || T <= 0) return MAS_B200_ERR_ARG;
    if (B > 65535) return MAS_B200_ERR_UNSUPPORTED;
    sequence_mask_kernel<<<dim3((T + kThreads - 1) / kThreads, B), kThreads, 0, stream>>>(lengths, T, mask);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_crop_frames(const float *y, const int *frame_token, const int *y_lengths, const int *offsets, int B, int F,
                       int Ty, int out_size, float *y_cut, int *ft_cut, int *cut_lengths, float *cut_mask,
                       cudaStream_t stream) {
    if (!y || !y_lengths || !offsets || !y_cut || B <= 0 || F <= 0 || Ty <= 0 || out_size <= 0) return MAS_B200_ERR_ARG;
    if (ft_cut && !frame_token) return MAS_B200_ERR_ARG;
    if (!grid_ok(B, F)) return MAS_B200_ERR_UNSUPPORTED;
    dim3 grid((out_size + kThreads - 1) / kThreads, (F + kF - 1) / kF, B);
    crop_frames_kernel<<<grid, kThreads, 0, stream>>>(y, frame_token, y_lengths, offsets, F, Ty, out_size, y_cut, ft_cut,
                                                      cut_lengths, cut_mask);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

size_t prior_loss_workspace_bytes(int B, int F, int Ty) {
    if (B <= 0 || F <= 0 || Ty <= 0) return 0;
    return sizeof(double) * (size_t)B * ((F + kF - 1) / kF) * ((Ty + kThreads - 1) / kThreads);
}

int launch_gather_mu_y(const float *mu_x, const int *frame_token, int B, int F, int Tx, int Ty, float *mu_y,
                       cudaStream_t stream) {
    if (!mu_x || !frame_token || !mu_y || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!grid_ok(B, F)) return MAS_B200_ERR_UNSUPPORTED;
    dim3 grid((Ty + kThreads - 1) / kThreads, (F + kF - 1) / kF, B);
    gather_mu_y_kernel<false><<<grid, kThreads, 0, stream>>>(mu_x, frame_token, nullptr, nullptr, F, Tx, Ty, mu_y, nullptr);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_prior_loss(const float *y, const float *mu_x, const int *frame_token, const int *y_lengths, int B, int F,
                      int Tx, int Ty, float *mu_y, float *loss, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
    if (!y || !mu_x || !frame_token || !y_lengths || !loss || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0)
        return MAS_B200_ERR_ARG;
    if (!grid_ok(B, F)) return MAS_B200_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < prior_loss_workspace_bytes(B, F, Ty)) return MAS_B200_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 7) return MAS_B200_ERR_ALIGN;
    dim3 grid((Ty + kThreads - 1) / kThreads, (F + kF - 1) / kF, B);
    double *partial = static_cast<double *>(workspace);
    gather_mu_y_kernel<true><<<grid, kThreads, 0, stream>>>(mu_x, frame_token, y, y_lengths, F, Tx, Ty, mu_y, partial);
    MASB200_CUDA_TRY(cudaGetLastError());
    prior_loss_finish_kernel<<<1, kThreads, 0, stream>>>(partial, (int)(grid.x * grid.y * grid.z), y_lengths, B, Ty, F, loss);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_gather_mu_y_bwd(const float *grad_mu_y, const int *start, const int *dur, const int *offsets,
                           const int *lengths, int B, int F, int Tx, int Ty, float *grad_mu_x, cudaStream_t stream) {
    if (!grad_mu_y || !start || !dur || !grad_mu_x || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (!grid_ok(B, F)) return MAS_B200_ERR_UNSUPPORTED;
    dim3 grid((Tx + kSegWarps - 1) / kSegWarps, (F + kF - 1) / kF, B);
    segment_sum_kernel<false><<<grid, kThreads, 0, stream>>>(grad_mu_y, nullptr, start, dur, offsets, lengths, nullptr, B,
                                                             F, Tx, Ty, grad_mu_x);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_prior_loss_bwd(const float *y, const float *mu_x, const int *start, const int *dur, const int *offsets,
                          const int *y_lengths, const float *grad_loss, int B, int F, int Tx, int Ty, float *grad_mu_x,
                          cudaStream_t stream) {
    if (!y || !mu_x || !start || !dur || !y_lengths || !grad_loss || !grad_mu_x || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0)
        return MAS_B200_ERR_ARG;
    if (!grid_ok(B, F)) return MAS_B200_ERR_UNSUPPORTED;
    dim3 grid((Tx + kSegWarps - 1) / kSegWarps, (F + kF - 1) / kF, B);
    segment_sum_kernel<true><<<grid, kThreads, 0, stream>>>(y, mu_x, start, dur, offsets, y_lengths, grad_loss, B, F, Tx,
                                                            Ty, grad_mu_x);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_duration_loss(const float *logw, const int *durations, const int *x_lengths, int B, int Tx, float *loss,
                         float *logw_target, float *grad_logw, cudaStream_t stream) {
    if (!logw || !durations || !x_lengths || !loss || B <= 0 || Tx <= 0) return MAS_B200_ERR_ARG;
    if ((long long)B * Tx > INT32_MAX) return MAS_B200_ERR_UNSUPPORTED;
    duration_loss_kernel<<<1, 1024, 0, stream>>>(logw, durations, x_lengths, B, Tx, loss, logw_target, grad_logw);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

}  // namespace masb200
