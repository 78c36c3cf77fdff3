// HBM bandwidth of the materialising KKT kernel (csrc/pmb_kernels.hpp::KktDenseBody) at the mobile-robot size, batch 8192
#include <cstdio>
#include "../../polympc_b200/csrc/pmb_kernels.hpp"

int main()
{
    const int N = 65, M = 39, n = N + M, B = 8192;
    double *H, *A, *rb, *ri, *K;
    cudaMalloc(&H, (size_t)B * N * N * 8); cudaMalloc(&A, (size_t)B * M * N * 8); cudaMalloc(&rb, (size_t)B * N * 8); cudaMalloc(&ri, (size_t)B * M * 8);
    cudaMalloc(&K, (size_t)B * n * n * 8);
    cudaMemset(H, 0, (size_t)B * N * N * 8); cudaMemset(A, 0, (size_t)B * M * N * 8); cudaMemset(rb, 0, (size_t)B * N * 8); cudaMemset(ri, 0, (size_t)B * M * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) pmb::launch<pmb::KktDenseBody>(B, 0, 0, N, M, (const double*)H, (const double*)A, (const double*)rb, (const double*)ri, 1e-6, K);
    cudaDeviceSynchronize();
    const int reps = 20;
    cudaEventRecord(e0);
    for (int rep = 0; rep < reps; ++rep) pmb::launch<pmb::KktDenseBody>(B, 0, 0, N, M, (const double*)H, (const double*)A, (const double*)rb, (const double*)ri, 1e-6, K);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double bytes = (double)B * 8 * (N * N + M * N + n * n);
    printf("kkt_assemble_dense: %.4f ms per launch, %.1f GB/s (%s)\n", ms, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    // plain device-to-device copy of the same number of bytes, for reference
    cudaEventRecord(e0);
    for (int rep = 0; rep < reps; ++rep) cudaMemcpyAsync(K, K + (size_t)B * n * n / 2, (size_t)B * n * n * 4, cudaMemcpyDeviceToDevice, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("cudaMemcpy D2D of %.0f MB (read + write %.0f MB): %.4f ms, %.1f GB/s\n", B * n * n * 4 / 1e6, B * n * n * 8 / 1e6, ms, (double)B * n * n * 8 / (ms * 1e-3) / 1e9);
    return 0;
}
