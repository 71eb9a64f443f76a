// Probe of tcgen05.mma kind::tf32 operand layouts (one CTA, one MMA M=128 N=32 K=8).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../face-gan-tts_b200/csrc/tc_common.cuh"
using namespace masb200;

// mode bit0: A from smem (K-major SW none?) not used; we test A from TMEM.
__global__ void __launch_bounds__(160, 1) probe(const float *A /*[128][8]*/, const float *Bm /*[8][32] k-major rows = MN contiguous*/,
                                              float *D /*[128][32]*/, uint32_t idesc, uint32_t lbo, uint32_t sbo, int use_mask_form, int ss, uint32_t idesc_ss) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *bs = reinterpret_cast<float *>(smem);                  // 8 rows x 128 B, 128B swizzle
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(&slot, 512); tmem_relinquish(); }
    // B tile: element (k, n) at k*128 + ((n/4) ^ (k%8))*16 + (n%4)*4
    for (int e = tid; e < 8 * 32; e += blockDim.x) {
        const int k = e / 32, n = e % 32;
        bs[k * 32 + (((n >> 2) ^ (k & 7)) << 2) + (n & 3)] = Bm[k * 32 + n];
    }
    // B K-major, no swizzle: core matrix (n/8, k/4) at (n/8)*256 + (k/4)*128, row (n%8)*16 B, elem (k%4)
    float *bk = reinterpret_cast<float *>(smem + 16384);
    for (int e = tid; e < 8 * 32; e += blockDim.x) {
        const int k = e / 32, n = e % 32;
        bk[(n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = Bm[k * 32 + n];
    }
    // A tile for SS mode, MN-major SW128: element (k, m): atom (m/32) at 1024*atom, row k at 128*k, chunk ((m%32)/4 ^ k), elem m%4
    float *as = reinterpret_cast<float *>(smem + 4096);
    for (int e = tid; e < 8 * 128; e += blockDim.x) {
        const int k = e / 128, m = e % 128;
        as[(m >> 5) * 256 + k * 32 + ((((m & 31) >> 2) ^ (k & 7)) << 2) + (m & 3)] = A[m * 8 + k];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp < 4) {
        uint32_t r[8];
        for (int k = 0; k < 8; ++k) r[k] = __float_as_uint(A[tid * 8 + k]);
        tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + 0, r);
        tmem_wait_st();
        // zero D region
        uint32_t z[8]; for (int q = 0; q < 8; ++q) z[q] = __float_as_uint(7.0f);
        for (int c = 0; c < 32; c += 8) tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + 64 + c, z);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 128) {
        const uint64_t bd = umma_smem_desc_mn_sw128(smem_u32(bs), lbo, sbo);
        if (ss == 3) {
            // TS, B K-major no swizzle: LBO = 128 (k chunk), SBO = 256 (8-row group)
            const uint64_t bkd = (uint64_t)((smem_u32(smem + 16384) & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
            umma_tf32_ts(tmem + 64, tmem + 0, bkd, idesc & ~(1u << 16), 0u);
        } else if (ss == 2) {
            // no MMA at all: st/ld round trip of the sentinel
        } else if (ss) {
            const uint64_t ad = umma_smem_desc_mn_sw128(smem_u32(smem + 4096), 1024, 1024);
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                         " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem + 64), "l"(ad), "l"(bd), "r"(idesc_ss), "r"(0u) : "memory");
        } else if (use_mask_form) {
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                         " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                         ::"r"(tmem + 64), "r"(tmem + 0), "l"(bd), "r"(idesc), "r"(0u), "r"(0u) : "memory");
        } else {
            umma_tf32_ts(tmem + 64, tmem + 0, bd, idesc, 0u);
        }
        umma_commit(&bar);
    }
    if (warp < 4) {
        mbar_wait(&bar, 0);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 64, r);
        tmem_wait_ld();
        for (int c = 0; c < 32; ++c) D[tid * 32 + c] = __uint_as_float(r[c]);
        // read back A too
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    std::vector<float> A(128 * 8), B(8 * 32), D(128 * 32), E(128 * 32);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 8; ++k) A[m * 8 + k] = (float)((m % 7) + 1) * (k == 2 ? 1.f : (k == 5 ? 0.5f : 0.f));
    for (int k = 0; k < 8; ++k) for (int n = 0; n < 32; ++n) B[k * 32 + n] = (k == 2 ? (float)(n + 1) : (k == 5 ? 100.f : 0.25f));
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { double s = 0; for (int k = 0; k < 8; ++k) s += (double)A[m * 8 + k] * B[k * 32 + n]; E[m * 32 + n] = (float)s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
    for (int mask_form = 0; mask_form < 5; ++mask_form) {
        cudaMemset(dD, 0xff, D.size() * 4);
        probe<<<1, 160, 130 * 1024>>>(dA, dB, dD, umma_idesc_tf32_ts(128, 32), 1024, 1024, mask_form & 1, mask_form == 2 ? 1 : (mask_form == 3 ? 2 : (mask_form == 4 ? 3 : 0)), umma_idesc_tf32_ts(128, 32) | (1u << 15));
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int nz = 0;
        for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs((double)D[i] - E[i])); nz += D[i] != 0.f; }
        printf("mask_form=%d: %s  max|D-E| = %g  nonzero=%d   D[0][0..3]= %g %g %g %g  E= %g %g %g %g   D[5][0..1]= %g %g  E= %g %g\n", mask_form, cudaGetErrorString(e), maxerr, nz,
               D[0], D[1], D[2], D[3], E[0], E[1], E[2], E[3], D[5 * 32], D[5 * 32 + 1], E[5 * 32], E[5 * 32 + 1]);
    }
    return 0;
}
