// experiments on the triangular-solve inner loops (one warp, n = 104, packed lower factor in shared memory)
#include <cstdio>
#include <vector>
#include "../../polympc_b200/csrc/pmb_qp.hpp"
using namespace pmb;
constexpr int R = 4;
__device__ long long g_t[8];
__global__ void __launch_bounds__(128, 4) k(int n, const double* Kin, double* out)
{
    extern __shared__ __align__(16) unsigned char sm[];
    Warp w;
    double* Lp = reinterpret_cast<double*>(sm);
    const int fac = n * (n + 1) / 2;
    double* sol = Lp + fac;
    for (int i = threadIdx.x; i < fac; i += 128) Lp[i] = Kin[i] * 1e-3;
    for (int i = threadIdx.x; i < n; i += 128) sol[i] = 1.0 + i;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    double y[R]; int offr[R];
    for (int r = 0; r < R; ++r) { const int i = lane + 32 * r; y[r] = i < n ? sol[i] : 0.0; offr[r] = i < n ? i * n - ((i * (i + 1)) >> 1) : 0; }
    long long t0 = clock64();
    {   // V0: product loop (forward)
        const double* colp = Lp + lane;
#pragma unroll
        for (int jb = 0; jb < R; ++jb) {
            const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
#pragma unroll 1
            for (int jj = 0; jj < jend; ++jj) {
                const int j = jb * 32 + jj;
                const double yj = w.shfl(y[jb], jj);
                if (lane > jj && 32 * jb + lane < n) y[jb] = dm::fma(-colp[32 * jb], yj, y[jb]);
#pragma unroll
                for (int r = jb + 1; r < R; ++r) if (r < R - 1 || lane + 32 * r < n) y[r] = dm::fma(-colp[32 * r], yj, y[r]);
                colp += n - j - 1;
            }
        }
    }
    long long t1 = clock64();
    {   // V1: forward, critical chunk first, only chunk jb (to see the pure chain)
        const double* colp = Lp + lane;
#pragma unroll
        for (int jb = 0; jb < R; ++jb) {
            const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
#pragma unroll 1
            for (int jj = 0; jj < jend; ++jj) {
                const int j = jb * 32 + jj;
                const double yj = w.shfl(y[jb], jj);
                if (lane > jj && 32 * jb + lane < n) y[jb] = dm::fma(-colp[32 * jb], yj, y[jb]);
                colp += n - j - 1;
            }
        }
    }
    long long t2 = clock64();
    {   // V2: forward, no shared loads at all (chain only)
#pragma unroll
        for (int jb = 0; jb < R; ++jb) {
            const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
#pragma unroll 1
            for (int jj = 0; jj < jend; ++jj) {
                const double yj = w.shfl(y[jb], jj);
                if (lane > jj) y[jb] = dm::fma(-1e-3, yj, y[jb]);
            }
        }
    }
    long long t3 = clock64();
    {   // V3: backward product loop
#pragma unroll
        for (int jb = R - 1; jb >= 0; --jb) {
            const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
#pragma unroll 1
            for (int jj = jend - 1; jj >= 0; --jj) {
                const int j = jb * 32 + jj;
                const double yj = w.shfl(y[jb], jj);
#pragma unroll
                for (int r = 0; r < jb; ++r) y[r] = dm::fma(-Lp[offr[r] + j], yj, y[r]);
                if (lane < jj) y[jb] = dm::fma(-Lp[offr[jb] + j], yj, y[jb]);
            }
        }
    }
    long long t4 = clock64();
    {   // V4: forward with prefetch of the next column's critical entry (software pipelining) and crit DFMA first
        const double* colp = Lp + lane;
#pragma unroll
        for (int jb = 0; jb < R; ++jb) {
            const int jend = (n - jb * 32) < 32 ? (n - jb * 32) : 32;
            double lc = (32 * jb + lane < n) ? colp[32 * jb] : 0.0;
#pragma unroll 1
            for (int jj = 0; jj < jend; ++jj) {
                const int j = jb * 32 + jj;
                const double yj = w.shfl(y[jb], jj);
                const double* nxt = colp + (n - j - 1);
                if (lane > jj) y[jb] = dm::fma(-lc, yj, y[jb]);
                lc = (32 * jb + lane < n && jj + 1 < jend) ? nxt[32 * jb] : 0.0;
#pragma unroll
                for (int r = jb + 1; r < R; ++r) if (r < R - 1 || lane + 32 * r < n) y[r] = dm::fma(-colp[32 * r], yj, y[r]);
                colp = nxt;
            }
        }
    }
    long long t5 = clock64();
    if (lane == 0) { g_t[0] = t1 - t0; g_t[1] = t2 - t1; g_t[2] = t3 - t2; g_t[3] = t4 - t3; g_t[4] = t5 - t4; }
    for (int r = 0; r < R; ++r) { const int i = lane + 32 * r; if (i < n) out[i] = y[r]; }
}
int main()
{
    const int n = 104, fac = n * (n + 1) / 2;
    std::vector<double> K(fac);
    for (int j = 0, e = 0; j < n; ++j) for (int i = j; i < n; ++i, ++e) K[e] = (i == j) ? 4.0 : 0.01 * ((i * 7 + j * 3) % 11 - 5);
    double *dKin, *dout;
    cudaMalloc(&dKin, fac * 8); cudaMalloc(&dout, n * 8);
    cudaMemcpy(dKin, K.data(), fac * 8, cudaMemcpyHostToDevice);
    const size_t smem = (fac + n) * 8 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[8];
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, smem>>>(n, dKin, dout); cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_t, sizeof h);
        printf("fwd product %lld | fwd crit-only %lld | chain-only %lld | bwd product %lld | fwd prefetch %lld  cycles (104 steps each)  %s\n", h[0], h[1], h[2], h[3], h[4], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
