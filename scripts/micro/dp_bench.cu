// Microbenchmark of the MAS tile body in isolation: W warps of one CTA each run dp_tile<R> over a
// resident shared-memory value tile N times (no TMA, no flags); prints cycles per frame.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMAS_CELL_VARIANT=k -o dp_bench dp_bench.cu
#include <cstdio>
#include "../../face-gan-tts_b200/csrc/mas_forward.cuh"

using namespace masb200;

#ifndef SPIN
#define SPIN 0
#endif
template <int R, int W, bool DIAG>
__global__ void __launch_bounds__((W + 1) * 32, 1) bench(const float *in, float *out, long long *cyc, int ntiles) {
    constexpr int XP = 32 * R * W;
    extern __shared__ __align__(1024) float smem[];
    float *tile = smem;                    // [XP][32]
    float *halo = smem + XP * 32;          // [W+1][32]
    for (int i = threadIdx.x; i < XP * 32; i += blockDim.x) tile[i] = in[i & 1023];
    for (int i = threadIdx.x; i < (W + 1) * 32; i += blockDim.x) halo[i] = -1e9f;
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w == W) {
#if SPIN == 1
        while (!stop) { if (mbar_try_wait(&bar, 0)) break; }
#elif SPIN == 2
        while (!stop) { if (mbar_try_wait(&bar, 0)) break; if (clock64() < 0) __trap(); }
#elif SPIN == 3
        while (!stop) { if (mbar_try_wait(&bar, 0)) break; __nanosleep(100); }
#endif
        return;
    }
    float q[R]; uint32_t acc[R];
    for (int r = 0; r < R; ++r) { q[r] = in[threadIdx.x + r]; acc[r] = 0; }
    float up = -1e9f;
    const float *lane_tile = tile + (32 * w + lane) * 32;
    const uint32_t hout = smem_u32(halo + (w + 1) * 32);   // all lanes store (same address)
        const long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < ntiles; ++j) {
        dp_tile<R, XP, DIAG>(q, acc, up, lane_tile, halo, lane & 7, lane == 0 ? 0xffffffffu : 0u, (int)(threadIdx.x) - j, -1e9f, hout);
        if (j & 1024) { for (int r = 0; r < R; ++r) acc[r] = 0; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) stop = 1;
    float s = up; uint32_t a = 0;
    for (int r = 0; r < R; ++r) { s += q[r]; a ^= acc[r]; }
    out[threadIdx.x] = s + (float)a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int R, int W, bool DIAG>
void run(const float *in, float *out, long long *cyc) {
    const int ntiles = 256;
    const size_t smem = sizeof(float) * (32 * R * W * 32 + (W + 1) * 32);
    cudaFuncSetAttribute(bench<R, W, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int k = 0; k < 2; ++k) bench<R, W, DIAG><<<1, (W + 1) * 32, smem>>>(in, out, cyc, ntiles);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("spin %d variant %d R=%d W=%d diag=%d  %.2f cyc/frame  %.2f cyc/cell   (%s)\n", SPIN, MAS_CELL_VARIANT, R, W, (int)DIAG,
           (double)c / (ntiles * 32), (double)c / (ntiles * 32) / R, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 8192); cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
    float h[2048]; for (int i = 0; i < 2048; ++i) h[i] = (float)((i * 37) % 11) - 5.0f;
    cudaMemcpy(in, h, 8192, cudaMemcpyHostToDevice);
    run<1, 1, false>(in, out, cyc);
    run<2, 1, false>(in, out, cyc);
    run<4, 1, false>(in, out, cyc);
    run<8, 1, false>(in, out, cyc);
    run<4, 2, false>(in, out, cyc);
    run<2, 4, false>(in, out, cyc);
    run<4, 1, true>(in, out, cyc);
    return 0;
}
