// micro-benchmarks that shaped the kernel design: fp64 FMA latency / throughput, SHFL and LDS latency, REDUX latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double* out, double a, double b, int iters, long long* cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dfma_ilp(double* out, double a, double b, int iters, long long* cyc)
{
    double x0 = out[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double* out, int iters, long long* cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = __shfl_sync(0xffffffffu, x, (i + 1) & 31); x = __shfl_sync(0xffffffffu, x, (i + 2) & 31); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_dfma_chain(double* out, double a, int iters, long long* cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { double y = __shfl_sync(0xffffffffu, x, i & 31); x = __fma_rn(y, a, x); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(int* out, int iters, long long* cyc)
{
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) s[i] = (i * 33 + 7) & 1023;
    __syncwarp();
    int x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = s[x]; x = s[x]; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void redux_chain(unsigned* out, int iters, long long* cyc)
{
    unsigned x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = __reduce_max_sync(0xffffffffu, x + threadIdx.x); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ddiv_chain(double* out, double a, int iters, long long* cyc)
{
    double x = out[threadIdx.x] + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = a / x; x = a / x; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8); cudaMemset(d, 0, 1 << 24);
    long long h; const int it = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        dfma_chain<<<1, 32>>>(d, 1.0000001, 1e-9, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency      : %.1f cycles\n", (double)h / (4.0 * it));
        dfma_ilp<<<1, 32>>>(d, 1.0000001, 1e-9, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DFMA 1 warp, ILP 8          : %.2f cycles/instr\n", (double)h / (8.0 * it));
        dfma_ilp<<<1, 128>>>(d, 1.0000001, 1e-9, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DFMA 4 warps/SM, ILP 8      : %.2f cycles/warp-instr/SMSP\n", (double)h / (8.0 * it));
        dfma_ilp<<<1, 512>>>(d, 1.0000001, 1e-9, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DFMA 16 warps/SM, ILP 8     : %.2f cycles per 4 warp-instr/SMSP\n", (double)h / (8.0 * it));
        shfl_chain<<<1, 32>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("SHFL(double) dependent       : %.1f cycles\n", (double)h / (2.0 * it));
        shfl_dfma_chain<<<1, 32>>>(d, 1e-9, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("SHFL+DFMA chain              : %.1f cycles\n", (double)h / (1.0 * it));
        lds_chain<<<1, 32>>>((int*)d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("LDS dependent                : %.1f cycles\n", (double)h / (2.0 * it));
        redux_chain<<<1, 32>>>((unsigned*)d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("REDUX.MAX dependent          : %.1f cycles\n", (double)h / (1.0 * it));
        ddiv_chain<<<1, 32>>>(d, 1.7, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DDIV dependent               : %.1f cycles\n", (double)h / (2.0 * it));
    }
    // aggregate fp64 throughput
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_ilp<<<148 * 8, 256>>>(d, 1.0000001, 1e-9, 1 << 16, c);
    cudaEventRecord(e0); dfma_ilp<<<148 * 8, 256>>>(d, 1.0000001, 1e-9, 1 << 16, c); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("fp64 FMA throughput: %.2f TFLOP/s (%s)\n", 2.0 * 148 * 8 * 256 * 8.0 * (1 << 16) / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
