// FP64 latency / throughput microbenchmark for B200 (sm_100a): dependent DFMA chain, div, sqrt; and throughput with N warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, long long *cyc, int iters, double seed) {
    double a = seed + threadIdx.x, b = 1.0000001, c = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a = __fma_rn(a, b, c);
    }
    long long t1 = clock64();
    double d = seed + 3.0 + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) d = 1.0 / d + 1.5;
    }
    long long t2 = clock64();
    double s = seed + 2.0 + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s = sqrt(s) + 2.0;
    }
    long long t3 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + s;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cyc[0] = t1 - t0;
        cyc[1] = t2 - t1;
        cyc[2] = t3 - t2;
    }
}
// throughput: each warp runs K independent DFMA chains
template <int K>
__global__ void k_tp(double *out, long long *cyc, int iters, double seed) {
    double a[K];
    for (int k = 0; k < K; ++k) a[k] = seed + k + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = __fma_rn(a[k], 1.0000001, 0.5);
    }
    long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < K; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out;
    long long *cyc, h[3];
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 64);
    int iters = 2000;
    k_lat<<<1, 32>>>(out, cyc, iters, 1.0);
    cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    printf("1 warp: dependent DFMA %.1f cyc, (1/d + c) %.1f cyc, (sqrt + c) %.1f cyc\n", (double)h[0] / (iters * 16), (double)h[1] / (iters * 4), (double)h[2] / (iters * 4));
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        k_tp<4><<<1, 32 * warps>>>(out, cyc, iters, 1.0);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double per = (double)h[0] / (iters * 8 * 4);
        printf("%2d warps/SM x 4 chains: %.2f cyc per DFMA per warp -> %.2f warp-DFMA/cyc/SM\n", warps, per, warps / per);
    }
    for (int warps : {4, 16}) {
        k_lat<<<1, 32 * warps>>>(out, cyc, iters, 1.0);
        cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
        printf("%2d warps: dependent DFMA %.1f cyc, (1/d + c) %.1f cyc, (sqrt + c) %.1f cyc per op per warp\n", warps, (double)h[0] / (iters * 16), (double)h[1] / (iters * 4), (double)h[2] / (iters * 4));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
