// FP64 / FP32 pipe probe for the B200: dependent-chain latency and per-SM throughput of DFMA, DADD, DMUL, FFMA;
// and the cost of a named barrier among 7 warps.  nvcc -arch=sm_100a -O3 -o fp64 fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = fma(x[i], b, a);
            if (OP == 1) x[i] = x[i] + b;
            if (OP == 2) x[i] = x[i] * b;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void kf(float *out, long long *cyc, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], b, a);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void kbar(long long *cyc, int iters) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <typename F> void run(const char *name, F f, int threads, int blocks, int iters, int ops_per_iter) {
    long long *cyc; double *out;
    cudaMalloc(&cyc, 8 * blocks); cudaMalloc(&out, 8 * threads * blocks);
    f(out, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); f(out, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double total_ops = (double)threads * blocks * iters * ops_per_iter;
    printf("%-28s threads=%4d blocks=%4d: %7.2f cycles/iter (block 0), %8.2f Gop/s, %.3f ms\n", name, threads, blocks,
           (double)h / iters, total_ops / ms / 1e6, ms);
    cudaFree(cyc); cudaFree(out);
}
int main() {
    const int it = 20000;
    run("DFMA latency (1 warp, ILP1)", [&](double *o, long long *c) { k<0, 1><<<1, 32>>>(o, c, it, 1.0, 0.999); }, 32, 1, it, 1);
    run("DADD latency (1 warp, ILP1)", [&](double *o, long long *c) { k<1, 1><<<1, 32>>>(o, c, it, 1.0, 0.999); }, 32, 1, it, 1);
    run("DMUL latency (1 warp, ILP1)", [&](double *o, long long *c) { k<2, 1><<<1, 32>>>(o, c, it, 1.0, 0.999); }, 32, 1, it, 1);
    run("DFMA 1 warp ILP8", [&](double *o, long long *c) { k<0, 8><<<1, 32>>>(o, c, it, 1.0, 0.999); }, 32, 1, it, 8);
    run("DFMA 4 warps ILP8 (1 SM)", [&](double *o, long long *c) { k<0, 8><<<1, 128>>>(o, c, it, 1.0, 0.999); }, 128, 1, it, 8);
    run("DFMA 16 warps ILP8 (1 SM)", [&](double *o, long long *c) { k<0, 8><<<1, 512>>>(o, c, it, 1.0, 0.999); }, 512, 1, it, 8);
    run("DFMA full GPU", [&](double *o, long long *c) { k<0, 8><<<148 * 4, 512>>>(o, c, it, 1.0, 0.999); }, 512, 148 * 4, it, 8);
    run("FFMA latency (1 warp, ILP1)", [&](double *o, long long *c) { kf<1><<<1, 32>>>((float *)o, c, it, 1.0f, 0.999f); }, 32, 1, it, 1);
    run("FFMA full GPU", [&](double *o, long long *c) { kf<8><<<148 * 4, 512>>>((float *)o, c, it, 1.0f, 0.999f); }, 512, 148 * 4, it, 8);
    run("bar.sync 7 warps", [&](double *o, long long *c) { kbar<<<1, 224>>>(c, it); }, 224, 1, it, 1);
    run("bar.sync 2 warps", [&](double *o, long long *c) { kbar<<<1, 64>>>(c, it); }, 64, 1, it, 1);
    return 0;
}
