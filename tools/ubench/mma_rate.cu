// Micro-benchmark: issue rate of mma.sync.m16n8k8 tf32 vs FFMA2 on one SM (all 4 SMSPs, W warps per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int CH>
__global__ void k_mma(float* out, long long* cyc, int iters) {
    float c[CH][4];
    for (int j = 0; j < CH; ++j) for (int i = 0; i < 4; ++i) c[j][i] = threadIdx.x * 1e-3f + j;
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) mma_tf32(c[j], a0, a1, a2, a3, b0, b1);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < CH; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void k_ffma2(float* out, long long* cyc, int iters) {
    float2 c[CH];
    for (int j = 0; j < CH; ++j) c[j] = make_float2(threadIdx.x * 1e-3f + j, 1.0f);
    float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(1e-6f, 2e-6f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) c[j] = __ffma2_rn(c[j], a, b);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < CH; ++j) s += c[j].x + c[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void k_ffma(float* out, long long* cyc, int iters) {
    float c[CH];
    for (int j = 0; j < CH; ++j) c[j] = threadIdx.x * 1e-3f + j;
    float a = 1.0001f + threadIdx.x * 1e-9f, b = 1e-6f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) c[j] = fmaf(c[j], a, b);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < CH; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        long long h;
        k_mma<8><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mma.sync m16n8k8 tf32, %2d warps/SM, 8 independent accumulators: %.2f cycles per mma per SMSP (%.0f MAC/clk/SM)\n", warps,
               (double)h / (iters * 8.0 * warps / 4.0), 1024.0 * iters * 8.0 * warps / h);
        k_mma<3><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mma.sync m16n8k8 tf32, %2d warps/SM, 3 independent accumulators: %.2f cycles per mma per SMSP\n", warps,
               (double)h / (iters * 3.0 * warps / 4.0));
        k_ffma2<16><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FFMA2, %2d warps/SM, 16 chains: %.2f cycles per warp-instruction per SMSP (%.0f MAC/clk/SM)\n", warps,
               (double)h / (iters * 16.0 * warps / 4.0), 64.0 * iters * 16.0 * warps / h);
        k_ffma<16><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FFMA,  %2d warps/SM, 16 chains: %.2f cycles per warp-instruction per SMSP (%.0f MAC/clk/SM)\n", warps,
               (double)h / (iters * 16.0 * warps / 4.0), 32.0 * iters * 16.0 * warps / h);
    }
    k_mma<1><<<1, 32>>>(out, cyc, iters); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mma.sync dependent-chain latency: %.1f cycles\n", (double)h / iters);
    return 0;
}
