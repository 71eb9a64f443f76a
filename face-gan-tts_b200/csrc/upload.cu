// upload.cu -- host -> device transfer of one padded batch that moves only the VALID part over PCIe.
//
// The call sites hand over mu_x [B,F,Tx] and y [B,F,Ty] padded to the batch maximum (reference
// text_encoder.py:417 zeroes mu_x beyond t_x; lrs2_dataset.py:256,265 zero-pads y beyond t_y): with LRS2 lengths
// ~35 % of those bytes are padding.  A plain cudaMemcpy of the padded tensors is PCIe-bound and ~4x longer than
// the alignment itself, so the e2e path pulls the rows with SM loads straight from page-locked host memory
// (UVA zero-copy), reads only [0,t_x) / [0,t_y) of every row, and writes zeros for the padding on the device --
// the device tensors are exactly the padded tensors the reference contract describes.  PCIe-bound:
// algorithmic bytes = 4*F*sum_b(t_x+t_y) + 8*B read over PCIe, 4*F*B*(Tx+Ty) written to HBM.
#include <algorithm>

#include "mas_common.cuh"
#include "mas_host.h"

namespace masb200 {

namespace {

constexpr int kRows = 4;          // feature rows per CTA: independent loads in flight per thread
constexpr int kThreads = 256;

// rows of one utterance: src/dst [F, T] with `valid` leading elements per row to copy, the rest zero-filled
// 16-byte load from mapped host memory; hint = 1 asks L2 for 256-byte fetches (fewer, larger PCIe read requests)
__device__ __forceinline__ float4 ld_host16(const float4 *p, int hint) {
    float4 v;
    if (hint)
        asm volatile("ld.global.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else
        asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int VEC>
__device__ __forceinline__ void pull_rows(const float *__restrict__ src, float *__restrict__ dst, int f0, int F, int T,
                                          int valid, int hint, int from = 0) {
    // elements [0, from) (from % 4 == 0) are someone else's job (the bulk-copy kernel below)
    if (VEC == 4) {
        const int nvec = T >> 2;
        for (int i = (from >> 2) + threadIdx.x; i < nvec; i += kThreads) {
            const int c = i << 2;
            float4 v[kRows];
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f0 + k < F && c < valid)     // the only PCIe reads: 16 bytes, all rows issued before any store
                    v[k] = ld_host16(reinterpret_cast<const float4 *>(src + (size_t)(f0 + k) * T + c), hint);
            }
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                if (f0 + k >= F) break;
                if (c + 1 >= valid) v[k].y = 0.f;
                if (c + 2 >= valid) v[k].z = 0.f;
                if (c + 3 >= valid) v[k].w = 0.f;
                *reinterpret_cast<float4 *>(dst + (size_t)(f0 + k) * T + c) = v[k];
            }
        }
    } else {
        for (int c = from + threadIdx.x; c < T; c += kThreads) {
            float v[kRows];
#pragma unroll
            for (int k = 0; k < kRows; ++k) v[k] = (f0 + k < F && c < valid) ? __ldcs(src + (size_t)(f0 + k) * T + c) : 0.f;
#pragma unroll
            for (int k = 0; k < kRows; ++k)
                if (f0 + k < F) dst[(size_t)(f0 + k) * T + c] = v[k];
        }
    }
}

template <int VECX, int VECY>
__global__ void __launch_bounds__(kThreads) upload_batch_kernel(const float *__restrict__ mu_h, const float *__restrict__ y_h,
                                                                const int *__restrict__ tx_h, const int *__restrict__ ty_h,
                                                                int B, int F, int Tx, int Ty, float *__restrict__ mu_d,
                                                                float *__restrict__ y_d, int *__restrict__ tx_d,
                                                                int *__restrict__ ty_d, int hint, int y_bulk) {
    // Persistent, one small CTA per SM: it only has to keep PCIe busy (a few hundred KB in flight), and it must
    // leave the thread slots of every SM to the alignment kernels of the previous step running beside it.
    const int groups = (F + kRows - 1) / kRows;
    for (int item = blockIdx.x; item < B * groups; item += gridDim.x) {
        const int b = item / groups, f0 = (item - b * groups) * kRows;
        const int tx = tx_h[b], ty = ty_h[b];      // 8 bytes over PCIe per item
        if (f0 == 0 && threadIdx.x == 0) { tx_d[b] = tx; ty_d[b] = ty; }
        // y_bulk: the leading 16-byte multiples of every y row travel by bulk copies (upload_bulk_kernel)
        pull_rows<VECY>(y_h + (size_t)b * F * Ty, y_d + (size_t)b * F * Ty, f0, F, Ty, min(max(ty, 0), Ty), hint,
                        y_bulk ? (min(max(ty, 0), Ty) & ~3) : 0);
        if (mu_h != nullptr)
            pull_rows<VECX>(mu_h + (size_t)b * F * Tx, mu_d + (size_t)b * F * Tx, f0, F, Tx, min(max(tx, 0), Tx), hint);
    }
}

// y rows through the TMA engine: global(host) -> shared -> global(device) bulk copies, one warp per CTA, a ring of
// kBulkStages row buffers with kBulkAhead loads in flight.  Only whole 16-byte multiples of the valid part of a row
// ([0, t_y & ~3)); the tail and the zero padding are the pull kernel's job.
constexpr int kBulkStages = 8, kBulkAhead = 6, kBulkChunk = 4096, kBulkMaxB = 1024;
__global__ void __launch_bounds__(32) upload_bulk_kernel(const float *__restrict__ y_h, const int *__restrict__ ty_h, int B, int F,
                                                         int Ty, float *__restrict__ y_d) {
    __shared__ __align__(128) unsigned char ring[kBulkStages][kBulkChunk];
    __shared__ uint64_t full[kBulkStages];
    __shared__ int ty_s[kBulkMaxB];
    for (int i = threadIdx.x; i < B; i += 32) ty_s[i] = ty_h[i];      // the lengths cross PCIe once per CTA
    __syncwarp();
    if (threadIdx.x != 0) return;
    for (int i = 0; i < kBulkStages; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    // work items of this CTA: (row, chunk) pairs in order; rows c, c + G, ...
    const int rows = B * F;
    int lrow = blockIdx.x, loff = 0;          // next item to LOAD
    int srow = blockIdx.x, soff = 0;          // next item to STORE
    int kl = 0, ks = 0;
    auto row_bytes = [&](int row) { return (min(max(ty_s[row / F], 0), Ty) & ~3) * 4; };
    auto issue_load = [&]() -> bool {
        while (lrow < rows) {
            const int nb = row_bytes(lrow);
            if (loff < nb) {
                const int n = min(kBulkChunk, nb - loff);
                const int st = kl % kBulkStages;
                mbar_arrive_expect_tx(&full[st], (uint32_t)n);
                tma_bulk_load_1d(ring[st], reinterpret_cast<const unsigned char *>(y_h + (size_t)lrow * Ty) + loff, (uint32_t)n, &full[st]);
                loff += n; ++kl;
                return true;
            }
            lrow += gridDim.x; loff = 0;
        }
        return false;
    };
    for (int i = 0; i < kBulkAhead; ++i) issue_load();
    while (ks < kl) {
        // item ks: find its (row, off, n) by replaying the same walk
        int nb = row_bytes(srow);
        while (soff >= nb) { srow += gridDim.x; soff = 0; nb = row_bytes(srow); }
        const int n = min(kBulkChunk, nb - soff);
        const int st = ks % kBulkStages;
        mbar_wait(&full[st], (uint32_t)(ks / kBulkStages) & 1u);
        tma_bulk_store_1d(reinterpret_cast<unsigned char *>(y_d + (size_t)srow * Ty) + soff, ring[st], (uint32_t)n);
        tma_store_commit();
        soff += n; ++ks;
        // the stage the next load goes to was stored kBulkStages - kBulkAhead items ago
        tma_store_wait_read_n<kBulkStages - kBulkAhead - 1>();
        issue_load();
    }
    tma_store_wait_all();
}

// zero-fill of the padding only: elements [valid, T) of every row (copy-engine variant; lengths from the mapped host arrays)
__global__ void __launch_bounds__(kThreads) zero_padding_kernel(const int *__restrict__ tx_h, const int *__restrict__ ty_h,
                                                                int B, int F, int Tx, int Ty, float *__restrict__ mu_d,
                                                                float *__restrict__ y_d) {
    for (int row = blockIdx.x; row < B * F; row += gridDim.x) {
        const int b = row / F;
        const int tx = min(max(tx_h[b], 0), Tx), ty = min(max(ty_h[b], 0), Ty);
        float *yr = y_d + (size_t)row * Ty, *mr = mu_d + (size_t)row * Tx;
        for (int t = ty + threadIdx.x; t < Ty; t += kThreads) yr[t] = 0.f;
        if (mu_d != nullptr)
            for (int t = tx + threadIdx.x; t < Tx; t += kThreads) mr[t] = 0.f;
    }
}

bool device_view(const void *host, const void **dev) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;   // pageable memory: not accessible
    *dev = a.devicePointer;
    return true;
}

bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int launch_upload_batch(const float *mu_x_pinned, const float *y_pinned, const int *t_xs_pinned, const int *t_ys_pinned,
                        int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev, int *t_x_dev, int *t_y_dev,
                        cudaStream_t stream) {
    if (!y_pinned || !t_xs_pinned || !t_ys_pinned || !y_dev || !t_x_dev || !t_y_dev || (mu_x_pinned && !mu_x_dev) ||
        B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0)
        return MAS_B200_ERR_ARG;
    const bool with_mu = mu_x_pinned != nullptr;     // NULL: the caller moves mu_x itself (e.g. on the copy engine, in parallel)
    DeviceInfo di;
    const int rc = device_info(&di);
    if (rc != MAS_B200_OK) return rc;
    const void *mu_v = nullptr, *y_v, *tx_v, *ty_v;
    if ((with_mu && !device_view(mu_x_pinned, &mu_v)) || !device_view(y_pinned, &y_v) || !device_view(t_xs_pinned, &tx_v) ||
        !device_view(t_ys_pinned, &ty_v))
        return MAS_B200_ERR_ARG;      // the host buffers must be page-locked (cudaHostAlloc / cudaHostRegister)
    if (option("upload_impl") == 2) {
        // Copy-engine variant: one pitched 2-D copy per utterance and tensor (F rows of t_y / t_x valid floats), the
        // padding zero-filled by a small kernel.  DMA reads move ~50 GB/s over PCIe gen5 where SM-issued zero-copy
        // loads reach ~40 GB/s, at the price of 2B+2 driver calls per batch on the host.
        zero_padding_kernel<<<std::min(B * F, 4 * di.sm_count), kThreads, 0, stream>>>(
            static_cast<const int *>(tx_v), static_cast<const int *>(ty_v), B, F, Tx, Ty, with_mu ? mu_x_dev : nullptr, y_dev);
        MASB200_CUDA_TRY(cudaGetLastError());
        for (int b = 0; b < B; ++b) {
            const size_t tx = with_mu ? (size_t)std::min(std::max(t_xs_pinned[b], 0), Tx) : 0, ty = (size_t)std::min(std::max(t_ys_pinned[b], 0), Ty);
            if (ty)
                MASB200_CUDA_TRY(cudaMemcpy2DAsync(y_dev + (size_t)b * F * Ty, sizeof(float) * Ty, y_pinned + (size_t)b * F * Ty,
                                                   sizeof(float) * Ty, sizeof(float) * ty, (size_t)F, cudaMemcpyHostToDevice, stream));
            if (tx)
                MASB200_CUDA_TRY(cudaMemcpy2DAsync(mu_x_dev + (size_t)b * F * Tx, sizeof(float) * Tx, mu_x_pinned + (size_t)b * F * Tx,
                                                   sizeof(float) * Tx, sizeof(float) * tx, (size_t)F, cudaMemcpyHostToDevice, stream));
        }
        MASB200_CUDA_TRY(cudaMemcpyAsync(t_x_dev, t_xs_pinned, sizeof(int) * B, cudaMemcpyHostToDevice, stream));
        MASB200_CUDA_TRY(cudaMemcpyAsync(t_y_dev, t_ys_pinned, sizeof(int) * B, cudaMemcpyHostToDevice, stream));
        return MAS_B200_OK;
    }
    const bool vx = with_mu && (Tx % 4 == 0) && al16(mu_v) && al16(mu_x_dev);
    const bool vy = (Ty % 4 == 0) && al16(y_v) && al16(y_dev);
    const long long items = (long long)B * ((F + kRows - 1) / kRows);
    const int ctas = option("upload_ctas");
    dim3 grid((unsigned)std::min<long long>(items, ctas > 0 ? ctas : di.sm_count));
    auto *mu = static_cast<const float *>(mu_v);
    auto *yy = static_cast<const float *>(y_v);
    auto *txp = static_cast<const int *>(tx_v);
    auto *typ = static_cast<const int *>(ty_v);
    // The alignment kernels of the previous step run beside this one and need (almost) all of an SM's shared
    // memory; an SM only changes its L1/shared split when idle, so this kernel asks for the same split -- otherwise
    // its persistent CTAs would keep every SM in the small-shared configuration and the two would serialise.
    static const bool carveout_set = [] {
        cudaFuncSetAttribute(upload_batch_kernel<4, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(upload_batch_kernel<1, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(upload_batch_kernel<4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(upload_batch_kernel<1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return true;
    }();
    (void)carveout_set;
    const int hint = option("upload_l2_256b");
    const int y_bulk = (option("upload_impl") == 3 && vy && B <= kBulkMaxB) ? 1 : 0;
    if (y_bulk) {
        upload_bulk_kernel<<<di.sm_count, 32, 0, stream>>>(yy, typ, B, F, Ty, y_dev);
        MASB200_CUDA_TRY(cudaGetLastError());
    }
    if (vx && vy)
        upload_batch_kernel<4, 4><<<grid, kThreads, 0, stream>>>(mu, yy, txp, typ, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev, t_y_dev, hint, y_bulk);
    else if (vy)
        upload_batch_kernel<1, 4><<<grid, kThreads, 0, stream>>>(mu, yy, txp, typ, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev, t_y_dev, hint, y_bulk);
    else if (vx)
        upload_batch_kernel<4, 1><<<grid, kThreads, 0, stream>>>(mu, yy, txp, typ, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev, t_y_dev, hint, y_bulk);
    else
        upload_batch_kernel<1, 1><<<grid, kThreads, 0, stream>>>(mu, yy, txp, typ, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev, t_y_dev, hint, y_bulk);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}


// ---------------------------------------------------------------------------------------------------
// Packed (ragged) batch: what a collate function that does NOT pad would hand over -- one contiguous buffer
//   [t_x: B int32][t_y: B int32][pad to 16 bytes][mu: for b: F rows of t_x[b] floats][y: for b: F rows of t_y[b] floats]
// It crosses PCIe with ONE copy-engine cudaMemcpyAsync (only valid data: ~8.1 of the 12.2 MB of a padded LRS2 batch),
// and this kernel expands it on the device into the zero-padded tensors the reference contract describes
// (mu_x [B,F,Tx] zero beyond t_x: text_encoder.py:417; y [B,F,Ty] zero-padded: lrs2_dataset.py:256,265).
// HBM-bound: reads 4F*sum(t_x + t_y), writes 4FB(Tx + Ty) bytes.  One CTA per (mel bin, utterance) row pair.
// ---------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) unpack_batch_kernel(const int *__restrict__ hdr, const float *__restrict__ body, int B,
                                                           int F, int Tx, int Ty, float *__restrict__ mu_x,
                                                           float *__restrict__ y, int *__restrict__ t_x_out,
                                                           int *__restrict__ t_y_out) {
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    __shared__ long long red[2][8];
    // exclusive prefix sums of the lengths up to utterance b, and the total of t_x (start of the y section)
    long long sx = 0, sy = 0, totx = 0;
    for (int i = tid; i < B; i += 256) {
        const long long a = max(0, min(hdr[i], Tx)), c = max(0, min(hdr[B + i], Ty));
        totx += a;
        if (i < b) { sx += a; sy += c; }
    }
    long long v[3] = {sx, sy, totx};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(kFullMask, v[k], o);
    }
    __shared__ long long part[3][8];
    if ((tid & 31) == 0) { part[0][tid >> 5] = v[0]; part[1][tid >> 5] = v[1]; part[2][tid >> 5] = v[2]; }
    __syncthreads();
    sx = sy = totx = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sx += part[0][w]; sy += part[1][w]; totx += part[2][w]; }
    (void)red;
    const int tx = max(0, min(hdr[b], Tx)), ty = max(0, min(hdr[B + b], Ty));
    if (f == 0 && tid == 0) { t_x_out[b] = hdr[b]; t_y_out[b] = hdr[B + b]; }      // unclamped: bad lengths are reported downstream
    const float *mu_src = body + (size_t)F * sx + (size_t)f * tx;
    const float *y_src = body + (size_t)F * totx + (size_t)F * sy + (size_t)f * ty;
    float *mu_dst = mu_x + ((size_t)b * F + f) * Tx;
    float *y_dst = y + ((size_t)b * F + f) * Ty;
    for (int x = tid; x < Tx; x += 256) mu_dst[x] = x < tx ? __ldcs(mu_src + x) : 0.f;
    for (int t = tid; t < Ty; t += 256) y_dst[t] = t < ty ? __ldcs(y_src + t) : 0.f;
}

}  // namespace

size_t packed_batch_header_bytes(int B) { return (((size_t)8 * B + 15) / 16) * 16; }

int launch_unpack_batch(const void *packed_dev, int B, int F, int Tx, int Ty, float *mu_x_dev, float *y_dev, int *t_x_dev,
                        int *t_y_dev, cudaStream_t stream) {
    if (!packed_dev || !mu_x_dev || !y_dev || !t_x_dev || !t_y_dev || B <= 0 || F <= 0 || Tx <= 0 || Ty <= 0 || B > 65535)
        return MAS_B200_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(packed_dev) & 15) return MAS_B200_ERR_ALIGN;
    const int *hdr = static_cast<const int *>(packed_dev);
    const float *body = reinterpret_cast<const float *>(static_cast<const char *>(packed_dev) + packed_batch_header_bytes(B));
    unpack_batch_kernel<<<dim3(F, B), 256, 0, stream>>>(hdr, body, B, F, Tx, Ty, mu_x_dev, y_dev, t_x_dev, t_y_dev);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

}  // namespace masb200
