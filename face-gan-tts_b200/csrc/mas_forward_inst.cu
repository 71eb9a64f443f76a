// mas_forward_inst.cu -- kernel instantiations for one rows-per-lane value R
// (compiled once per R with -DMASB200_INST_R=R so the translation units build in parallel).
#include <atomic>

#include "mas_forward.cuh"
#include "mas_host.h"

#ifndef MASB200_INST_R
#error "compile with -DMASB200_INST_R=1|2|4|8"
#endif

namespace masb200 {

namespace {

constexpr size_t kMaxSmem = 232448 - 1024;   // must match mas_forward.cu

template <int R, int W, bool SB, bool MP>
int launch_inst(const MasParams &P, const CUtensorMap &tmap, size_t smem, cudaStream_t stream) {
    static std::atomic<int> configured[16];
    int dev = 0;
    MASB200_CUDA_TRY(cudaGetDevice(&dev));
    auto kern = mas_forward_kernel<R, W, SB, MP>;
    if (dev < 0 || dev >= 16 || !configured[dev].load(std::memory_order_acquire)) {
        MASB200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
        if (dev >= 0 && dev < 16) configured[dev].store(1, std::memory_order_release);
    }
    if (P.B <= 0) {
        // prepare only: force the (lazily loaded) kernel image onto the device now.  A later launch that has to
        // load it would wait for running kernels -- a deadlock when one of them is waiting for THIS kernel.
        static std::atomic<int> loaded[16];
        if (dev < 0 || dev >= 16 || !loaded[dev].load(std::memory_order_acquire)) {
            cudaFuncAttributes attr;
            MASB200_CUDA_TRY(cudaFuncGetAttributes(&attr, kern));
            if (dev >= 0 && dev < 16) loaded[dev].store(1, std::memory_order_release);
        }
        return MAS_B200_OK;
    }
    kern<<<P.B, (W + 1 + (SB ? 1 : 0)) * 32, smem, stream>>>(P, tmap);
    MASB200_CUDA_TRY(cudaGetLastError());
    return MAS_B200_OK;
}

template <int R, int W>
int launch_w(const MasParams &P, const CUtensorMap &tmap, int mode, size_t smem, cudaStream_t stream) {
    switch (mode) {
        case MAS_MODE_SMEM_BITS: return launch_inst<R, W, true, false>(P, tmap, smem, stream);
        case MAS_MODE_GLOBAL_BITS: return launch_inst<R, W, false, false>(P, tmap, smem, stream);
        case MAS_MODE_MULTIPASS: return launch_inst<R, W, false, true>(P, tmap, smem, stream);
        default: return MAS_B200_ERR_UNSUPPORTED;
    }
}

template <int R>
int launch_r(const MasParams &P, const CUtensorMap &tmap, int W, int mode, size_t smem, cudaStream_t stream) {
    switch (W) {
        case 1: return launch_w<R, 1>(P, tmap, mode, smem, stream);
        case 2: return launch_w<R, 2>(P, tmap, mode, smem, stream);
        case 3: return launch_w<R, 3>(P, tmap, mode, smem, stream);
        case 4: return launch_w<R, 4>(P, tmap, mode, smem, stream);
        default: return MAS_B200_ERR_UNSUPPORTED;
    }
}

}  // namespace

#define MASB200_CAT2(a, b) a##b
#define MASB200_CAT(a, b) MASB200_CAT2(a, b)

int MASB200_CAT(launch_mas_r, MASB200_INST_R)(const MasParams &P, const CUtensorMap &tmap, int W, int mode,
                                              size_t smem, cudaStream_t stream) {
    return launch_r<MASB200_INST_R>(P, tmap, W, mode, smem, stream);
}

}  // namespace masb200
