// fp64 tensor-core (DMMA m8n8k4) latency / throughput on B200, next to the LDS.64 + DFMA mat-vec inner loop and the fp64
// reciprocal — the numbers behind the tile-based (fast arithmetic) LDL^T, csrc/pmb_qp_fast.hpp
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma_chain(double* out, int iters, long long* cyc)
{
    double c0 = out[threadIdx.x], c1 = c0 + 1, a = 1.0000001, b = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = c0 + c1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void dmma_ilp(double* out, int iters, long long* cyc)
{
    double c0[ILP], c1[ILP]; const double a = 1.0000001, b = 0.5;
    for (int k = 0; k < ILP; ++k) { c0[k] = out[threadIdx.x] + k; c1[k] = c0[k] + 1; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) dmma(c0[k], c1[k], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// thread-per-row mat-vec inner loop out of shared memory: acc = fma(L[j*ld + row], b[j], acc)
__global__ void lds_matvec(double* out, int iters, int n, long long* cyc)
{
    extern __shared__ double sm[];
    double* L = sm; double* b = sm + n * 128;
    for (int i = threadIdx.x; i < n * 128 + n; i += blockDim.x) sm[i] = 1.0 / (1 + i);
    __syncthreads();
    double acc0 = 0, acc1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int j = 0; j < n; j += 2) { acc0 = __fma_rn(L[j * 128 + threadIdx.x], b[j], acc0); acc1 = __fma_rn(L[(j + 1) * 128 + threadIdx.x], b[j + 1], acc1); }
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc0 + acc1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void rcp_chain(double* out, int iters, long long* cyc)
{
    double x = out[threadIdx.x] + 3.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = __drcp_rn(x) + 1.5; x = __drcp_rn(x) + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void bar_chain(double* out, int iters, long long* cyc)
{
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { __syncthreads(); __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8); cudaMemset(d, 0, 1 << 24);
    long long h; const int it = 4096;
    cudaFuncSetAttribute(lds_matvec, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        dmma_chain<<<1, 32>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DMMA m8n8k4 dependent latency     : %.1f cycles\n", (double)h / (4.0 * it));
        dmma_ilp<4><<<1, 32>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DMMA 1 warp, ILP 4               : %.2f cycles/instr\n", (double)h / (4.0 * it));
        dmma_ilp<8><<<1, 32>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DMMA 1 warp, ILP 8               : %.2f cycles/instr\n", (double)h / (8.0 * it));
        dmma_ilp<8><<<1, 128>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DMMA 4 warps/SM, ILP 8           : %.2f cycles/warp-instr/SMSP\n", (double)h / (8.0 * it));
        dmma_ilp<8><<<1, 512>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("DMMA 16 warps/SM, ILP 8          : %.2f cycles/warp-instr/SMSP\n", (double)h / (8.0 * it * 4));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); dmma_ilp<8><<<148 * 4, 512>>>(d, it * 4, c); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA whole GPU                   : %.2f TFLOP/s\n", 148.0 * 4 * 16 * 8 * it * 4 * 512.0 / (ms * 1e-3) / 1e12);
        lds_matvec<<<1, 128, (104 * 128 + 104) * 8>>>(d, 256, 104, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("LDS.64 x2 + DFMA mat-vec, 4 warps : %.2f cycles per (row, column) step\n", (double)h / (256.0 * 104));
        lds_matvec<<<3, 128, (104 * 128 + 104) * 8>>>(d, 256, 104, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("   same, 3 CTAs (different SMs)   : %.2f\n", (double)h / (256.0 * 104));
        rcp_chain<<<1, 32>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("__drcp_rn + DADD dependent        : %.1f cycles\n", (double)h / (2.0 * it));
        bar_chain<<<1, 128>>>(d, it, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("__syncthreads, 4 warps            : %.1f cycles\n", (double)h / (2.0 * it));
    }
    return 0;
}
