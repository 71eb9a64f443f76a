// Microbenchmarks (1 warp, 1 CTA): dependent-chain latency and per-frame cost of candidate MAS cell
// formulations on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#define FULL 0xffffffffu
constexpr int N = 4096;

__device__ __forceinline__ float cell_pred(float vc, float vp, float v, unsigned &bits, unsigned m) {
    float q;
    asm volatile("{\n .reg .pred p;\n setp.gt.f32 p, %3, %2;\n add.rn.f32 %0, %2, %4;\n @p add.rn.f32 %0, %3, %4;\n @p or.b32 %1, %1, %5;\n}\n"
        : "=&f"(q), "+r"(bits) : "f"(vc), "f"(vp), "f"(v), "r"(m));
    return q;
}
__device__ __forceinline__ float cell_sel(float vc, float vp, float v, unsigned &bits, unsigned m) {
    const bool d = vp > vc;
    bits |= d ? m : 0u;
    return (d ? vp : vc) + v;
}
__device__ __forceinline__ float cell_max(float vc, float vp, float v, unsigned &bits, unsigned m) {
    // fmaxf on the chain; direction bit computed off-chain
    const float q = fmaxf(vc, vp) + v;
    bits |= (vp > vc) ? m : 0u;
    return q;
}
__device__ __forceinline__ float cell_max_sign(float vc, float vp, float v, unsigned &bits, unsigned m) {
    const float q = fmaxf(vc, vp) + v;
    const float d = vc - vp;                        // sign bit == (vp > vc) for finite, non-(-0) operands
    bits = __funnelshift_l(__float_as_uint(d), bits, 1);
    return q;
}

template <int MODE, int R>
__global__ void frame_kernel(const float *in, float *out, long long *cyc, int src_lane_off) {
    float q[R];
    unsigned acc[R];
    for (int r = 0; r < R; ++r) { q[r] = in[threadIdx.x * R + r]; acc[r] = 0; }
    const int lane = threadIdx.x & 31;
    const int src = (lane + 31) & 31;
    const float vbase = in[64 + lane];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = (lane == 31) ? vbase : q[R - 1];
            const float up = __shfl_sync(FULL, send, src);
            float n[R];
#pragma unroll
            for (int r = R - 1; r >= 0; --r) {
                const float v = vbase + (float)(r + i);
                const float vp = (r == 0) ? up : q[r - 1];
                if (MODE == 0) n[r] = cell_pred(q[r], vp, v, acc[r], 1u << i);
                if (MODE == 1) n[r] = cell_sel(q[r], vp, v, acc[r], 1u << i);
                if (MODE == 2) n[r] = cell_max(q[r], vp, v, acc[r], 1u << i);
                if (MODE == 3) n[r] = cell_max_sign(q[r], vp, v, acc[r], 1u << i);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) q[r] = n[r];
        }
    }
    long long t1 = clock64();
    float s = 0; unsigned a = 0;
    for (int r = 0; r < R; ++r) { s += q[r]; a ^= acc[r]; }
    out[threadIdx.x] = s + (float)a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// pure dependent chains
template <int MODE>
__global__ void chain_kernel(const float *in, float *out, long long *cyc) {
    float q = in[threadIdx.x], c = in[32 + threadIdx.x], v = in[64 + threadIdx.x];
    unsigned bits = 0;
    const int lane = threadIdx.x & 31, src = (lane + 31) & 31;
    __shared__ int sidx[64];
    sidx[threadIdx.x] = (threadIdx.x * 7 + 1) & 31;
    __syncthreads();
    int idx = lane;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < N; ++it) {
        if (MODE == 0) q = cell_pred(q, c, v, bits, 1u);         // setp -> @p add
        if (MODE == 1) q = cell_sel(q, c, v, bits, 1u);          // setp -> sel -> add
        if (MODE == 2) q = fmaxf(q, c) + v;                      // fmnmx -> add
        if (MODE == 3) q = __shfl_sync(FULL, q, src);            // shfl chain
        if (MODE == 4) q = q + v;                                // fadd chain
        if (MODE == 5) idx = sidx[idx];                          // lds chain
        if (MODE == 6) q = __shfl_sync(FULL, fmaxf(q, c) + v, src);   // shfl + fmnmx + add
    }
    long long t1 = clock64();
    out[threadIdx.x] = q + (float)bits + (float)idx;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 37) % 11) - 5.0f;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    long long c;
    const char *cn[] = {"setp->@p add", "setp->sel->add", "fmnmx->add", "shfl", "fadd", "lds", "shfl+fmnmx+add"};
#define CHAIN(M) chain_kernel<M><<<1, 32>>>(in, out, cyc); chain_kernel<M><<<1, 32>>>(in, out, cyc); cudaDeviceSynchronize(); \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("chain %-18s %.2f cyc/iter\n", cn[M], (double)c / N);
    CHAIN(0) CHAIN(1) CHAIN(2) CHAIN(3) CHAIN(4) CHAIN(5) CHAIN(6)
    const char *fn[] = {"pred", "sel", "max+setp", "max+sign"};
#define FRAME(M, R) frame_kernel<M, R><<<1, 32>>>(in, out, cyc, 0); frame_kernel<M, R><<<1, 32>>>(in, out, cyc, 0); cudaDeviceSynchronize(); \
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("frame %-9s R=%d  %.2f cyc/frame  %.2f cyc/cell\n", fn[M], R, (double)c / N, (double)c / N / R);
    FRAME(0, 1) FRAME(0, 2) FRAME(0, 4) FRAME(0, 8)
    FRAME(1, 1) FRAME(1, 2) FRAME(1, 4) FRAME(1, 8)
    FRAME(2, 1) FRAME(2, 2) FRAME(2, 4) FRAME(2, 8)
    FRAME(3, 1) FRAME(3, 2) FRAME(3, 4) FRAME(3, 8)
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
