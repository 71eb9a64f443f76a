// path_ops.cu -- small HBM-bound kernels either side of the MAS kernel:
//   path_expand       [start,dur] table -> dense [B,Tx,Ty] path (pure streaming write, all SMs)
//   lengths_from_mask dense prefix mask -> t_x, t_y   (reference monotonic_align/__init__.py:20-21)
//   generate_path     integer durations -> dense path (reference model/utils.py:27-40)
#include "mas_forward.cuh"
#include "mas_host.h"

namespace masb200 {

namespace {

constexpr int kExpandRows = 8;
constexpr int kExpandThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kExpandThreads) path_expand_kernel(const int *__restrict__ start,
                                                                     const int *__restrict__ dur, int Tx, int Ty,
                                                                     T *__restrict__ path) {
    const int b = blockIdx.y;
    const int x0 = blockIdx.x * kExpandRows;
    const int rows = min(kExpandRows, Tx - x0);
    const int *sb = start + (size_t)b * Tx + x0;
    const int *db = dur + (size_t)b * Tx + x0;
    T *pb = path + ((size_t)b * Tx + x0) * Ty;
    write_path_rows<T>(pb, sb, db, rows, Ty, threadIdx.x, kExpandThreads);
}

__global__ void __launch_bounds__(256) lengths_from_mask_kernel(const float *__restrict__ mask, int Tx, int Ty,
                                                                int *__restrict__ t_x, int *__restrict__ t_y) {
    // fp32 sums like numpy's mask.sum(1)[:,0] / mask.sum(2)[:,0]; exact for 0/1 masks (< 2^24 terms)
    const int b = blockIdx.x;
    const float *m = mask + (size_t)b * Tx * Ty;
    float sx = 0.f, sy = 0.f;
    for (int x = threadIdx.x; x < Tx; x += blockDim.x) sx += m[(size_t)x * Ty];
    for (int y = threadIdx.x; y < Ty; y += blockDim.x) sy += m[y];
    __shared__ float red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(kFullMask, sx, o);
        sy += __shfl_xor_sync(kFullMask, sy, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sx; red[1][threadIdx.x >> 5] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ax = 0.f, ay = 0.f;
        for (int i = 0; i < 8; ++i) { ax += red[0][i]; ay += red[1][i]; }
        t_x[b] = (int)ax;     // .astype(np.int32) truncates
        t_y[b] = (int)ay;
    }
}

// One CTA per utterance: inclusive scan of the durations in shared memory, then the rows.
template <typename T>
__global__ void __launch_bounds__(256) generate_path_kernel(const int *__restrict__ durations,
                                                            const int *__restrict__ t_x, const int *__restrict__ t_y,
                                                            int Tx, int Ty, T *__restrict__ path) {
    extern __shared__ int sm[];
    int *start_s = sm;          // [Tx]
    int *dur_s = sm + Tx;       // [Tx]
    const int b = blockIdx.x;
    const int tx = min(max(t_x[b], 0), Tx), ty = min(max(t_y[b], 0), Ty);
    if (threadIdx.x == 0) {
        // Tx is a few hundred: a serial scan costs less than the dense write that follows
        long long cum = 0;
        for (int x = 0; x < Tx; ++x) {
            const long long d = max(durations[(size_t)b * Tx + x], 0);
            const long long s = cum < ty ? cum : ty;
            cum += d;
            const long long e = cum < ty ? cum : ty;           // sequence_mask(cum, t_y) then * mask
            start_s[x] = (int)s;
            dur_s[x] = (x < tx) ? (int)(e - s) : 0;
        }
    }
    __syncthreads();
    write_path_rows<T>(path + (size_t)b * Tx * Ty, start_s, dur_s, Tx, Ty, threadIdx.x, blockDim.x);
}

}  // namespace

int launch_path_expand(const int *start, const int *dur, int B, int Tx, int Ty, void *path, int path_dtype,
                       cudaStream_t stream) {
    if (!start || !dur || !path || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    dim3 grid((Tx + kExpandRows - 1) / kExpandRows, B);
    if (path_dtype == MAS_B200_PATH_F32)
        path_expand_kernel<float><<<grid, kExpandThreads, 0, stream>>>(start, dur, Tx, Ty, static_cast<float *>(path));
    else if (path_dtype == MAS_B200_PATH_I32)
        path_expand_kernel<int><<<grid, kExpandThreads, 0, stream>>>(start, dur, Tx, Ty, static_cast<int *>(path));
    else
        return MAS_B200_ERR_ARG;
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_lengths_from_mask(const float *mask, int B, int Tx, int Ty, int *t_x, int *t_y, cudaStream_t stream) {
    if (!mask || !t_x || !t_y || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    lengths_from_mask_kernel<<<B, 256, 0, stream>>>(mask, Tx, Ty, t_x, t_y);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_generate_path(const int *durations, const int *t_x, const int *t_y, int B, int Tx, int Ty, void *path,
                         int path_dtype, cudaStream_t stream) {
    if (!durations || !t_x || !t_y || !path || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    const size_t smem = sizeof(int) * 2 * (size_t)Tx;
    if (smem > 48 * 1024) return MAS_B200_ERR_UNSUPPORTED;
    if (path_dtype == MAS_B200_PATH_F32)
        generate_path_kernel<float><<<B, 256, smem, stream>>>(durations, t_x, t_y, Tx, Ty, static_cast<float *>(path));
    else if (path_dtype == MAS_B200_PATH_I32)
        generate_path_kernel<int><<<B, 256, smem, stream>>>(durations, t_x, t_y, Tx, Ty, static_cast<int *>(path));
    else
        return MAS_B200_ERR_ARG;
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

}  // namespace masb200
