// isolates the phases of the QP kernel (pivot order, factorisation, one triangular solve pair) on one CTA: cycles by clock64
#include <cstdio>
#include <vector>
__device__ long long g_tick[8];
#define PMB_TICK(k) if (threadIdx.x == 0) g_tick[k] = clock64();
#include "../../polympc_b200/csrc/pmb_qp.hpp"
using namespace pmb;
#ifndef NREP
#define NREP 10
#endif
template <int R, int NT>
__global__ void __launch_bounds__(NT, 4) phases(int n, const double* Kin, long long* cyc, double* out)
{
    extern __shared__ __align__(16) unsigned char sm[];
    Warp w; 
    double* scratch = reinterpret_cast<double*>(sm);
    Cta c(w, scratch);
    double* Lp = scratch + Cta::SCRATCH_DOUBLES;
    const int fac = n * (n + 1) / 2;
    double* dK = Lp + fac; double* tmp = dK + n; double* sol = tmp + n; int* perm = reinterpret_cast<int*>(sol + n);
    for (int i = threadIdx.x; i < fac; i += NT) Lp[i] = Kin[i];
    for (int i = threadIdx.x; i < n; i += NT) { dK[i] = Kin[packed_off(i, n)]; sol[i] = 1.0 + i; }
    __syncthreads();
    long long t0 = clock64();
    ldlt_pivot_order<R>(c, n, dK, perm, reinterpret_cast<int*>(dK + 4 * n));
    long long t1 = clock64();
    for (int i = threadIdx.x; i < n; i += NT) perm[i] = i;
    __syncthreads();
    long long t2 = clock64();
    ldlt_factor_packed<R>(c, n, Lp);
    long long t3 = clock64();
    for (int rep = 0; rep < NREP; ++rep) ldlt_solve_packed<R>(c, n, Lp, perm, sol, tmp);
    long long t4 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = (t4 - t3) / NREP; }
    for (int i = threadIdx.x; i < n; i += NT) out[i] = sol[i];
}
int main()
{
    const int n = 104, fac = n * (n + 1) / 2;
    std::vector<double> K(fac);
    for (int j = 0, e = 0; j < n; ++j) for (int i = j; i < n; ++i, ++e) K[e] = (i == j) ? (j < 65 ? 4.0 + 0.01 * j : -10.0) : 0.01 * ((i * 7 + j * 3) % 11 - 5);
    double *dKin, *dout; long long* dc;
    cudaMalloc(&dKin, fac * 8); cudaMalloc(&dout, n * 8); cudaMalloc(&dc, 64);
    cudaMemcpy(dKin, K.data(), fac * 8, cudaMemcpyHostToDevice);
    const size_t smem = (Cta::SCRATCH_DOUBLES + fac + 3 * n) * 8 + n * 4 + 64;
    cudaFuncSetAttribute(phases<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(phases<4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[3];
    for (int rep = 0; rep < 2; ++rep) {
        phases<4, 128><<<1, 128, smem>>>(n, dKin, dc, dout); cudaMemcpy(h, dc, 24, cudaMemcpyDeviceToHost);
        printf("128 threads: pivot %lld  factor %lld  solve(pair) %lld cycles   (%s)\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
        { long long tk[8]; cudaMemcpyFromSymbol(tk, g_tick, sizeof tk); printf("   last solve: fwd %lld diag %lld bwd %lld epilogue+sync %lld\n", tk[1]-tk[0], tk[2]-tk[1], tk[3]-tk[2], tk[4]-tk[3]); }
        phases<4, 32><<<1, 32, smem>>>(n, dKin, dc, dout); cudaMemcpy(h, dc, 24, cudaMemcpyDeviceToHost);
        printf(" 32 threads: pivot %lld  factor %lld  solve(pair) %lld cycles   (%s)\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
