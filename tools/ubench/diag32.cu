// variants of the serial part of the triangular solve: one warp substitutes through a 32x32 unit-lower diagonal block
#include <cstdio>
#include <cuda_runtime.h>
__device__ long long g_t[16];
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
__global__ void k(const double* Lin, double* out)
{
    __shared__ double L[32 * 33];     // L[i + 33*j], column-major with padding
    const int lane = threadIdx.x;
    for (int e = lane; e < 32 * 33; e += 32) L[e] = Lin[e];
    __syncwarp();
    double y0 = 1.0 + lane;
    long long t[8];
    // V1: rolled loop, LDS inside
    double y = y0;
    t[0] = clock64();
#pragma unroll 1
    for (int jj = 0; jj < 32; ++jj) {
        const double yj = __shfl_sync(0xffffffffu, y, jj);
        if (lane > jj) y = fma_(-L[lane + 33 * jj], yj, y);
    }
    t[1] = clock64();
    double r1 = y;
    // V2: column entries preloaded in registers, unrolled
    y = y0;
    double l[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) l[jj] = -L[lane + 33 * jj];
    t[2] = clock64();
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
        const double yj = __shfl_sync(0xffffffffu, y, jj);
        if (lane > jj) y = fma_(l[jj], yj, y);
    }
    t[3] = clock64();
    double r2 = y;
    // V3: unrolled, unpredicated (l = 0 above the diagonal)
    y = y0;
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) l[jj] = lane > jj ? -L[lane + 33 * jj] : 0.0;
    t[4] = clock64();
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
        const double yj = __shfl_sync(0xffffffffu, y, jj);
        y = fma_(l[jj], yj, y);
    }
    t[5] = clock64();
    double r3 = y;
    // V4: rolled, LDS inside, unroll 4
    y = y0;
    t[6] = clock64();
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
        const double yj = __shfl_sync(0xffffffffu, y, jj);
        if (lane > jj) y = fma_(-L[lane + 33 * jj], yj, y);
    }
    t[7] = clock64();
    out[lane] = r1 + r2 + r3 + y;
    if (lane == 0) for (int i = 0; i < 8; ++i) g_t[i] = t[i];
    out[32 + lane] = (r1 == r2 && r2 == r3 && r3 == y) ? 1.0 : 0.0;
}
int main()
{
    double h[32 * 33];
    for (int j = 0; j < 32; ++j) for (int i = 0; i < 33; ++i) h[i + 33 * j] = 0.01 * ((i * 7 + j * 3) % 11 - 5);
    double *d, *o; cudaMalloc(&d, sizeof h); cudaMalloc(&o, 64 * 8); cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 32>>>(d, o); cudaDeviceSynchronize();
        long long t[16]; cudaMemcpyFromSymbol(t, g_t, sizeof t);
        double ho[64]; cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
        printf("32 steps: V1 rolled+LDS %lld | V2 preloaded unrolled %lld (preload %lld) | V3 unpredicated %lld | V4 unroll4+LDS %lld   same=%g  %s\n",
               t[1] - t[0], t[3] - t[2], t[2] - t[1], t[5] - t[4], t[7] - t[6], ho[40], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
