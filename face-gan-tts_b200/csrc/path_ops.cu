// path_ops.cu -- small HBM-bound kernels either side of the MAS kernel:
//   path_expand       [start,dur] table -> dense [B,Tx,Ty] path (pure streaming write, all SMs)
//   lengths_from_mask dense prefix mask -> t_x, t_y   (reference monotonic_align/__init__.py:20-21)
//   generate_path     integer durations -> dense path (reference model/utils.py:27-40)
#include <algorithm>
#include <atomic>

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
    pdl_wait();                   // launched programmatically behind the kernel that produces the table
    pdl_launch_dependents();
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
// D = int: integer durations.  D = float: the reference's float durations (w_ceil * length_scale,
// model/face_tts.py:118-119): cumsum in index order with a double accumulator rounded to fp32 per element (what
// torch.cumsum does on the CPU; its CUDA scan associates differently, within an fp32 rounding), and since the
// reference tests `t < cum` on integer t (sequence_mask, model/utils.py:10), token x owns [ceil(cum[x-1]), ceil(cum[x])).
template <typename T, typename D>
__global__ void __launch_bounds__(256) generate_path_kernel(const D *__restrict__ durations,
                                                            const int *__restrict__ t_x, const int *__restrict__ t_y,
                                                            int Tx, int Ty, T *__restrict__ path,
                                                            int *__restrict__ frame_token) {
    extern __shared__ int sm[];
    int *start_s = sm;          // [Tx]
    int *dur_s = sm + Tx;       // [Tx]
    const int b = blockIdx.x;
    const int tx = min(max(t_x[b], 0), Tx), ty = min(max(t_y[b], 0), Ty);
    if (threadIdx.x == 0) {
        // Tx is a few hundred: a serial scan costs less than the dense write that follows
        double acc = 0.0;
        long long e_prev = 0;
        for (int x = 0; x < Tx; ++x) {
            acc += (double)durations[(size_t)b * Tx + x];
            const D cum = (D)acc;
            long long e;
            if (sizeof(D) == sizeof(float) && !(cum == cum)) e = 0;                 // NaN: `t < nan` is false
            else if ((double)cum >= (double)ty) e = ty;
            else if ((double)cum <= 0.0) e = 0;
            else e = (long long)ceil((double)cum);                                  // sequence_mask(cum, t_y)
            // path = mask(cum[x]) - mask(cum[x-1]): ones on [e_prev, e); a negative duration would give -1s in
            // the reference, which no caller produces (durations are ceil(exp(.)) >= 0) -- clamped to empty here
            const long long s = e_prev < e ? e_prev : e;
            start_s[x] = (int)s;
            dur_s[x] = (x < tx) ? (int)(e - s) : 0;                                 // * mask
            e_prev = e > e_prev ? e : e_prev;
        }
    }
    __syncthreads();
    if (path != nullptr) write_path_rows<T>(path + (size_t)b * Tx * Ty, start_s, dur_s, Tx, Ty, threadIdx.x, blockDim.x);
    if (frame_token != nullptr) {
        int *ft = frame_token + (size_t)b * Ty;
        for (int t = threadIdx.x; t < Ty; t += blockDim.x) ft[t] = -1;
        __syncthreads();
        for (int x = threadIdx.x; x < Tx; x += blockDim.x)
            for (int t = start_s[x], e = start_s[x] + dur_s[x]; t < e; ++t) ft[t] = x;
    }
}

}  // namespace

int launch_path_expand(const int *start, const int *dur, int B, int Tx, int Ty, void *path, int path_dtype,
                       cudaStream_t stream) {
    if (!start || !dur || !path || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((Tx + kExpandRows - 1) / kExpandRows, B);
    cfg.blockDim = dim3(kExpandThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = option("pdl") != 0 ? 1 : 0;
    if (path_dtype == MAS_B200_PATH_F32)
        MASB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, path_expand_kernel<float>, start, dur, Tx, Ty, static_cast<float *>(path)));
    else if (path_dtype == MAS_B200_PATH_I32)
        MASB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, path_expand_kernel<int>, start, dur, Tx, Ty, static_cast<int *>(path)));
    else
        return MAS_B200_ERR_ARG;
    return MAS_B200_OK;
}

// tests only: `ctas` CTAs that each hold a whole SM (max dynamic shared memory) for `cycles` clock cycles
__global__ void __launch_bounds__(32, 1) debug_spin_kernel(long long cycles) {
    extern __shared__ unsigned char hog[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) __nanosleep(200);
    if (cycles < 0) hog[threadIdx.x] = 0;
}

int launch_debug_spin(int ctas, long long cycles, cudaStream_t stream) {
    if (ctas <= 0 || cycles < 0) return MAS_B200_ERR_ARG;
    DeviceInfo di;
    int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    static std::atomic<int> configured[16];
    int dev = 0;
    MASB200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16 || !configured[dev].load(std::memory_order_acquire)) {
        MASB200_CUDA_TRY(cudaFuncSetAttribute(debug_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        if (dev >= 0 && dev < 16) configured[dev].store(1, std::memory_order_release);
    }
    debug_spin_kernel<<<ctas, 32, di.max_smem_optin, stream>>>(cycles);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_lengths_from_mask(const float *mask, int B, int Tx, int Ty, int *t_x, int *t_y, cudaStream_t stream) {
    if (!mask || !t_x || !t_y || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    lengths_from_mask_kernel<<<B, 256, 0, stream>>>(mask, Tx, Ty, t_x, t_y);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

int launch_generate_path(const void *durations, int dur_is_float, const int *t_x, const int *t_y, int B, int Tx, int Ty,
                         void *path, int path_dtype, int *frame_token, cudaStream_t stream) {
    if (!durations || !t_x || !t_y || B <= 0 || Tx <= 0 || Ty <= 0) return MAS_B200_ERR_ARG;
    if (path_dtype != MAS_B200_PATH_NONE && path_dtype != MAS_B200_PATH_F32 && path_dtype != MAS_B200_PATH_I32)
        return MAS_B200_ERR_ARG;
    if (path_dtype != MAS_B200_PATH_NONE && !path) return MAS_B200_ERR_ARG;
    if (path_dtype == MAS_B200_PATH_NONE && !frame_token) return MAS_B200_ERR_ARG;
    const size_t smem = sizeof(int) * 2 * (size_t)Tx;
    if (smem > 48 * 1024) return MAS_B200_ERR_UNSUPPORTED;
    float *pf = path_dtype == MAS_B200_PATH_F32 ? static_cast<float *>(path) : nullptr;
    int *pi = path_dtype == MAS_B200_PATH_I32 ? static_cast<int *>(path) : nullptr;
    if (dur_is_float) {
        auto *d = static_cast<const float *>(durations);
        if (pi) generate_path_kernel<int, float><<<B, 256, smem, stream>>>(d, t_x, t_y, Tx, Ty, pi, frame_token);
        else generate_path_kernel<float, float><<<B, 256, smem, stream>>>(d, t_x, t_y, Tx, Ty, pf, frame_token);
    } else {
        auto *d = static_cast<const int *>(durations);
        if (pi) generate_path_kernel<int, int><<<B, 256, smem, stream>>>(d, t_x, t_y, Tx, Ty, pi, frame_token);
        else generate_path_kernel<float, int><<<B, 256, smem, stream>>>(d, t_x, t_y, Tx, Ty, pf, frame_token);
    }
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}


// ---------------------------------------------------------------------------------------------------
// One-sided gather of a [rows, cols] int32 block (the durations of this rank) into every rank's
// [world * rows, cols] buffer: peer[r] is rank r's buffer as mapped into THIS device's address space
// (symmetric memory, peer-to-peer stores over NVLink).  Multi-GPU loss bookkeeping without a collective:
// utterances are sharded across ranks (reference config.py:144-145 shards the batch the same way) and
// nothing else is ever exchanged.  16-byte stores when the block allows.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) put_rows_kernel(const int *__restrict__ src, int *const *__restrict__ peer, long long n,
                                                       long long dst_off, int world) {
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const bool vec = (n & 3) == 0 && (dst_off & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    for (int r = 0; r < world; ++r) {
        int *dst = peer[r] + dst_off;
        if (vec && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (long long i = i0; i < (n >> 2); i += stride)
                reinterpret_cast<int4 *>(dst)[i] = reinterpret_cast<const int4 *>(src)[i];
        } else {
            for (long long i = i0; i < n; i += stride) dst[i] = src[i];
        }
    }
}

int launch_put_rows(const int *src, int rows, int cols, void *const *peer_ptrs_dev, int world, int rank, cudaStream_t stream) {
    if (!src || !peer_ptrs_dev || rows <= 0 || cols <= 0 || world <= 0 || rank < 0 || rank >= world) return MAS_B200_ERR_ARG;
    const long long n = (long long)rows * cols;
    const int blocks = (int)std::min<long long>(32, (n / 4 + 255) / 256 + 1);
    put_rows_kernel<<<blocks, 256, 0, stream>>>(src, reinterpret_cast<int *const *>(peer_ptrs_dev), n, (long long)rank * n, world);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

}  // namespace masb200
