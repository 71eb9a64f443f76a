// Microbenchmark: what slows the MAS tile body down inside a busy CTA?  Two DP warps (warps 0, 1 = scheduler
// partitions 0, 1) run dp_tile<4> over a resident tile while noise warps run beside them:
//   bit 0  warps 2, 3 (partitions 2, 3) stream a large straight-line code region      (instruction fetch)
//   bit 1  warps 2, 3 hammer shared memory with LDS.128 / STS.128                       (shared-memory pipe)
//   bit 2  warp 4 (partition 0, the partition of DP warp 0) runs a light ALU loop       (issue slots)
//   bit 3  warp 4 polls an mbarrier with try_wait                                        (waiting neighbours)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o dp_noise dp_noise.cu
#include <cstdio>
#include "../../face-gan-tts_b200/csrc/mas_forward.cuh"

using namespace masb200;

template <int N> struct Rep {
    __device__ __forceinline__ static void run(float &a, float &b, float c) {
#pragma unroll
        for (int i = 0; i < N; ++i) { a = fmaf(a, c, b); b = fmaf(b, c, a); }
    }
};

__global__ void __launch_bounds__(160, 1) bench(const float *in, float *out, long long *cyc, int ntiles, int noise) {
    constexpr int R = 4, W = 2, XP = 32 * R * W;
    extern __shared__ __align__(1024) float smem[];
    float *tile = smem;                    // [XP][32]
    float *halo = smem + XP * 32;          // [W+1][32]
    float *scratch = halo + (W + 1) * 32;  // 16 KB for the shared-memory noise
    for (int i = threadIdx.x; i < XP * 32; i += blockDim.x) tile[i] = in[i & 1023];
    for (int i = threadIdx.x; i < (W + 1) * 32; i += blockDim.x) halo[i] = -1e9f;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) scratch[i] = 1.f;
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w >= W) {
        float a = in[lane], b2 = in[lane + 32], acc = 0.f;
        if (w < 4) {
            if (noise & 1) {
                while (!stop) { Rep<1500>::run(a, b2, 1.0001f); }        // 3000 FFMA = 48 KB of code per pass
            } else if (noise & 2) {
                float4 *sp = reinterpret_cast<float4 *>(scratch) + (w - 2) * 512 + lane;
                while (!stop) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { float4 v = sp[32 * k]; v.x += 1.f; sp[32 * ((k + 1) & 7)] = v; }
                }
            }
        } else {
            if (noise & 4) { while (!stop) { Rep<8>::run(a, b2, 1.0001f); } }
            else if (noise & 8) { while (!stop) { if (mbar_try_wait(&bar, 0)) break; } }
        }
        out[threadIdx.x] = a + b2 + acc;
        return;
    }
    float q[R]; uint32_t acc[R];
    for (int r = 0; r < R; ++r) { q[r] = in[threadIdx.x + r]; acc[r] = 0; }
    float up = -1e9f;
    const float *lane_tile = tile + (32 * w + lane) * 32;
    const uint32_t hout = smem_u32(halo + (w + 1) * 32);
    const long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < ntiles; ++j) {
        dp_tile<R, XP, false>(q, acc, up, lane_tile, halo, lane & 7, lane == 0 ? 0xffffffffu : 0u, (int)(threadIdx.x) - j, -1e9f, hout);
        if (j & 1024) { for (int r = 0; r < R; ++r) acc[r] = 0; }
    }
    const long long t1 = clock64();
    __syncwarp();
    if (threadIdx.x == 0) stop = 1;
    float s = up; uint32_t a = 0;
    for (int r = 0; r < R; ++r) { s += q[r]; a ^= acc[r]; }
    out[threadIdx.x] = s + (float)a;
    if (lane == 0) cyc[w] = t1 - t0;
}

int main() {
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 8192); cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
    float h[2048]; for (int i = 0; i < 2048; ++i) h[i] = (float)((i * 37) % 11) - 5.0f;
    cudaMemcpy(in, h, 8192, cudaMemcpyHostToDevice);
    const int ntiles = 256;
    const size_t smem = sizeof(float) * (32 * 4 * 2 * 32 + 3 * 32 + 4096);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int noise : {0, 1, 2, 4, 8, 1 | 4, 2 | 4, 1 | 8}) {
        for (int k = 0; k < 2; ++k) bench<<<1, 160, smem>>>(in, out, cyc, ntiles, noise);
        cudaDeviceSynchronize();
        long long c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
        printf("noise %2d (icache %d smem %d same-partition alu %d poll %d): warp0 %.2f warp1 %.2f cyc/frame   (%s)\n", noise, noise & 1,
               (noise >> 1) & 1, (noise >> 2) & 1, (noise >> 3) & 1, (double)c[0] / (ntiles * 32), (double)c[1] / (ntiles * 32),
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
