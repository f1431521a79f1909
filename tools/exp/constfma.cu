// Micro-benchmark: throughput of the small-operator contraction acc[i] += D[i][j]*f[j] on sm_100a for different
// ways of delivering the operator D (15 x 48): constant bank (ptxas emits one LDCU.64 per DFMA), shared memory
// broadcast, and register-blocking over 2 or 4 elements per thread so each constant load feeds several DFMAs.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NI = 15, NF = 48;
__constant__ double cD[NI][NF];
template <int EPT, bool SMEM_OP>
__global__ void __launch_bounds__(128) k(const double* __restrict__ F, double* __restrict__ out, int iters) {
    __shared__ double sD[NI][NF];
    if (SMEM_OP) { for (int t = threadIdx.x; t < NI * NF; t += blockDim.x) sD[t / NF][t % NF] = cD[t / NF][t % NF]; __syncthreads(); }
    double f[EPT][4];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int e = 0; e < EPT; e++) for (int q = 0; q < 4; q++) f[e][q] = F[(tid * EPT + e) * 4 + q];
    double acc[EPT][NI];
    for (int e = 0; e < EPT; e++) for (int i = 0; i < NI; i++) acc[e][i] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NF; j++) {
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const double d = SMEM_OP ? sD[i][j] : cD[i][j];
#pragma unroll
                for (int e = 0; e < EPT; e++) acc[e][i] = fma(d, f[e][j & 3], acc[e][i]);
            }
        }
        for (int e = 0; e < EPT; e++) f[e][it & 3] += 1e-9 * acc[e][it % NI];
    }
    double s = 0;
    for (int e = 0; e < EPT; e++) for (int i = 0; i < NI; i++) s += acc[e][i];
    out[tid] = s;
}
template <int EPT, bool SMEM_OP> void run(const double* F, double* out, const char* name) {
    int blocks = 148 * 8, iters = 40;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<EPT, SMEM_OP><<<blocks, 128>>>(F, out, 2);
    cudaEventRecord(a);
    k<EPT, SMEM_OP><<<blocks, 128>>>(F, out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fma_total = (double)blocks * 128 * EPT * NI * NF * iters;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<EPT, SMEM_OP>, 128, 0);
    printf("%-28s %.3f ms  %.2f TFMA/s (%.1f%% of 18.6 TFMA/s = 64 lanes x 148 SM x 1.965 GHz)  occ %d CTA/SM\n", name, ms,
           fma_total / (ms * 1e-3) / 1e12, 100 * fma_total / (ms * 1e-3) / 18.6e12, occ);
}
int main() {
    double *F, *out; cudaMalloc(&F, 148 * 8 * 128 * 4 * 4 * 8); cudaMalloc(&out, 148 * 8 * 128 * 8);
    cudaMemset(F, 0, 148 * 8 * 128 * 4 * 4 * 8);
    double h[NI][NF]; for (int i = 0; i < NI; i++) for (int j = 0; j < NF; j++) h[i][j] = 1e-3 * (i + j);
    cudaMemcpyToSymbol(cD, h, sizeof(h));
    run<1, false>(F, out, "const bank, 1 elem/thread");
    run<2, false>(F, out, "const bank, 2 elem/thread");
    run<4, false>(F, out, "const bank, 4 elem/thread");
    run<1, true>(F, out, "smem broadcast, 1 elem/thread");
    run<2, true>(F, out, "smem broadcast, 2 elem/thread");
    run<4, true>(F, out, "smem broadcast, 4 elem/thread");
    return 0;
}
