// Dependent-chain latencies of the instructions the solver's critical path is made of (one warp).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k(double *out, long long *cyc, double a0, double b0, int *idx) {
    __shared__ double sm[64];
    __shared__ int si[64];
    const int lane = threadIdx.x;
    sm[lane] = a0 + lane; sm[lane + 32] = b0;
    si[lane] = (lane + 1) & 31; si[lane + 32] = lane;
    __syncwarp();
    double x = a0 + lane, y = b0;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, y, y);
    t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + y;
    t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * y;
    t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
    // shfl double chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
    // rsqrt.approx.f64 chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
    // LDS pointer chase (int index -> double)
    int j = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) j = si[j];
    t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
    // LDS.64 dependent through address computed from loaded double
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { x = sm[(j + (int)x) & 63]; }
    t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
    // 3 independent DFMA chains
    double u = x + 1, v = x + 2, w = x + 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { u = fma(u, y, y); v = fma(v, y, y); w = fma(w, y, y); }
    t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
    // FFMA chain
    float f = (float)x, g = (float)y;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, g, g);
    t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
    // full 64-bit rcp (1.0/x) chain and sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; i++) x = 1.0 / x;
    t1 = clock64(); if (lane == 0) cyc[9] = (t1 - t0) * 8;
    // int shfl chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) j = __shfl_xor_sync(0xffffffffu, j, 1);
    t1 = clock64(); if (lane == 0) cyc[10] = t1 - t0;
    // ballot+popc chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) j = __popc(__ballot_sync(0xffffffffu, j & 1)) + lane;
    t1 = clock64(); if (lane == 0) cyc[11] = t1 - t0;
    out[lane] = x + u + v + w + f + j; idx[lane] = j;
}
int main() {
    double *out; long long *cyc; int *idx;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&idx, 64 * 4);
    for (int warps = 1; warps <= 16; warps *= 4) {
        k<<<148 * (warps > 4 ? 4 : 1), 32 * (warps > 4 ? 4 : warps)>>>(out, cyc, 1.0000001, 0.9999999, idx);
        k<<<148 * (warps > 4 ? 4 : 1), 32 * (warps > 4 ? 4 : warps)>>>(out, cyc, 1.0000001, 0.9999999, idx);
        long long h[16];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const char *names[] = {"DFMA", "DADD", "DMUL", "SHFL.f64", "RSQRT64.approx", "LDS chase int", "LDS.64 chase", "3x DFMA (per trip)", "FFMA", "1.0/x f64", "SHFL.i32", "ballot+popc"};
        printf("warps/SM = %d\n", warps);
        for (int i = 0; i < 12; i++) printf("  %-20s %7.1f cycles\n", names[i], (double)h[i] / N);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
