// Micro-benchmark: FP64 tensor-core throughput on sm_100a (mma.sync m8n8k4 and m16n8k8/k16 f64) in the shape the
// DFR contraction needs: A = operator fragments resident in registers, B streamed, several independent accumulators.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int NACC>
__global__ void __launch_bounds__(128) k884(const double *in, double *out, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double a[12], b[4], c[NACC][2];
    for (int i = 0; i < 12; i++) a[i] = in[(tid + i) & 1023];
    for (int i = 0; i < 4; i++) b[i] = in[(tid + 7 * i) & 1023];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int ks = 0; ks < 12; ks++)
#pragma unroll
            for (int t = 0; t < NACC; t++) dmma884(c[t][0], c[t][1], a[ks], b[(t + ks) & 3]);
    }
    double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[tid] = s;
}
template <int NACC>
__global__ void __launch_bounds__(128) k1688(const double *in, double *out, int iters) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double a[6][4], b[4][2], c[NACC][4];
    for (int i = 0; i < 6; i++) for (int q = 0; q < 4; q++) a[i][q] = in[(tid + i * 4 + q) & 1023];
    for (int i = 0; i < 4; i++) for (int q = 0; q < 2; q++) b[i][q] = in[(tid + 7 * i + q) & 1023];
    for (int i = 0; i < NACC; i++) for (int q = 0; q < 4; q++) c[i][q] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int ks = 0; ks < 6; ks++)
#pragma unroll
            for (int t = 0; t < NACC; t++) dmma1688(c[t], a[ks], b[(t + ks) & 3]);
    }
    double s = 0; for (int i = 0; i < NACC; i++) for (int q = 0; q < 4; q++) s += c[i][q];
    out[tid] = s;
}
template <typename K> void run(K kern, const char *name, double fma_per_thread_iter, const double *in, double *out, int cta_per_sm) {
    int blocks = 148 * cta_per_sm, iters = 200;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<blocks, 128>>>(in, out, 2);
    cudaEventRecord(a);
    kern<<<blocks, 128>>>(in, out, iters);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fma_total = (double)blocks * 128 * fma_per_thread_iter * iters;
    printf("%-34s %8.3f ms  %6.2f TFMA/s (%5.1f%% of 18.6)  %s\n", name, ms, fma_total / (ms * 1e-3) / 1e12,
           100 * fma_total / (ms * 1e-3) / 18.6e12, cudaGetErrorString(e));
}
int main() {
    double *in, *out; cudaMalloc(&in, 1024 * 8); cudaMalloc(&out, 148 * 16 * 128 * 8); cudaMemset(in, 0, 1024 * 8);
    // FMA per thread per iteration = (#mma) * (M*N*K) / 32
    run(k884<4>, "m8n8k4 4 acc, 4 CTA/SM", 12 * 4 * 256.0 / 32, in, out, 4);
    run(k884<8>, "m8n8k4 8 acc, 4 CTA/SM", 12 * 8 * 256.0 / 32, in, out, 4);
    run(k884<8>, "m8n8k4 8 acc, 2 CTA/SM", 12 * 8 * 256.0 / 32, in, out, 2);
    run(k884<8>, "m8n8k4 8 acc, 1 CTA/SM", 12 * 8 * 256.0 / 32, in, out, 1);
    run(k1688<4>, "m16n8k8 4 acc, 4 CTA/SM", 6 * 4 * 1024.0 / 32, in, out, 4);
    run(k1688<4>, "m16n8k8 4 acc, 2 CTA/SM", 6 * 4 * 1024.0 / 32, in, out, 2);
    run(k1688<4>, "m16n8k8 4 acc, 1 CTA/SM", 6 * 4 * 1024.0 / 32, in, out, 1);
    return 0;
}
