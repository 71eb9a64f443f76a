// tcgen05.mma kind::tf32 throughput vs N (A from TMEM, B K-major no-swizzle in smem), M=128, K=8.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate umma_rate.cu
#include <cstdio>
#include "../../face-gan-tts_b200/csrc/tc_common.cuh"
using namespace masb200;

__global__ void __launch_bounds__(160, 1) rate(long long *cyc, int N, int reps, int a_smem, int nacc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(&slot, 512); tmem_relinquish(); }
    for (int e = tid; e < 16384; e += blockDim.x) reinterpret_cast<float *>(smem)[e] = 1.0f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 4) {
        const uint32_t idesc = umma_idesc_tf32_ts(128, N);
        const uint64_t bd = umma_smem_desc_k_nosw(smem_u32(smem), 128, 256);
        const uint64_t ad = umma_smem_desc_k_nosw(smem_u32(smem + 32768), 128, 256);
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            if (a_smem) {
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem + 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile("{\n .reg .pred p, q;\n elect.sync _|q, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n @q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(tmem + 256 + (i & (nacc - 1)) * 32), "r"(tmem + (i & 7) * 8), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        if (elect_one()) umma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (tid == 128) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
    for (int nacc : {1, 2, 4, 8})
        for (int N : {32, 128}) {
            const int reps = 600, a_smem = 0;
            rate<<<1, 160, 130 * 1024>>>(d, N, reps, a_smem, nacc);
            rate<<<1, 160, 130 * 1024>>>(d, N, reps, a_smem, nacc);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("nacc=%d A from %s  N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)  -> %.0f MAC/cycle\n", nacc, a_smem ? "smem" : "TMEM", N,
                   (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(e), 128.0 * N * 8 * reps / h[1]);
        }
    return 0;
}
