// phases of the fast-arithmetic KKT linear algebra (csrc/pmb_qp_fast.hpp) on the mobile-robot size n = 104: cycles by clock64
// for one CTA alone and for one CTA of a full grid (3 CTAs per SM on every SM: shared-memory / tensor-pipe contention)
#include <cstdio>
#include <vector>
#include "../../polympc_b200/csrc/pmb_qp.hpp"
using namespace pmb;
#ifndef NREP
#define NREP 20
#endif
template <int R>
__global__ void __launch_bounds__(128, 3) phases(int n, const double* Hin, long long* cyc, double* out)
{
    extern __shared__ __align__(16) unsigned char sm[];
    Warp w;
    double* scratch = reinterpret_cast<double*>(sm);
    Cta c(w, scratch, blockIdx.x);
    double* Lp = scratch + Cta::SCRATCH_DOUBLES;
    const fast::Ws fw(Lp, n);
    double* dK = Lp + fast::workspace_doubles(n); double* sol = dK + n; int* perm = reinterpret_cast<int*>(sol + n); int* rk = perm + n;
    for (int i = threadIdx.x; i < n; i += 128) { dK[i] = Hin[i + (size_t)i * n]; sol[i] = 1.0 + i; }
    __syncthreads();
    long long t0 = clock64();
    ldlt_pivot_order<R>(c, n, dK, perm, rk);
    long long t1 = clock64();
    fast::gather<R>(c, n, 0, Hin, nullptr, dK, perm, fw);
    long long t2 = clock64();
    fast::FactorProf fp;
    fast::factor(c, fw, &fp);
    long long t3 = clock64();
    fast::invert(c, fw);
    long long t4 = clock64();
    for (int rep = 0; rep < NREP; ++rep) {
        for (int e = threadIdx.x; e < fw.T * 8; e += 128) fw.tb[e] = e < n ? sol[perm[e]] * 0.5 : 0.0;
        __syncthreads();
        fast::solve_rows<4, 4>(c, fw, perm, sol);
    }
    long long t5 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x / 2) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = (t5 - t4) / NREP;
        cyc[5] = fp.diag; cyc[6] = fp.panel; cyc[7] = fp.trail;
    }
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < n; i += 128) out[i] = sol[i];
}
int main()
{
    const int n = 104;
    std::vector<double> H((size_t)n * n);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) H[i + (size_t)j * n] = (i == j) ? (j < 65 ? 4.0 + 0.01 * j : -10.0 - 0.001 * j) : 0.01 * (((i > j ? i * 7 + j * 3 : j * 7 + i * 3)) % 11 - 5);
    double *dH, *dout; long long* dc;
    cudaMalloc(&dH, H.size() * 8); cudaMalloc(&dout, n * 8); cudaMalloc(&dc, 128);
    cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = (Cta::SCRATCH_DOUBLES + fast::workspace_doubles(n) + 2 * n) * 8 + 2 * n * 4 + 64;
    cudaFuncSetAttribute(phases<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[8];
    for (int grid : {1, 148 * 3}) {
        for (int rep = 0; rep < 2; ++rep) {
            phases<4><<<grid, 128, smem>>>(n, dH, dc, dout); cudaMemcpy(h, dc, 64, cudaMemcpyDeviceToHost);
            printf("grid %3d: pivot %lld  gather %lld  factor %lld (diag %lld panel %lld trailing %lld)  invert %lld  solve(pair) %lld cycles   (%s)\n", grid, h[0], h[1], h[2], h[5], h[6],
                   h[7], h[3], h[4], cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
