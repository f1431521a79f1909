// Micro-benchmark: streaming [rows][K] arrays tile by tile, as k_elem does (E elements per CTA,
// each thread handles a row subset).  Measures how tile width affects achieved HBM bandwidth.
#include <cstdio>
#include <cuda_runtime.h>
template <int E, int ROWS_IN, int ROWS_OUT, int TPB>
__global__ void __launch_bounds__(TPB) k(const double* __restrict__ in, double* __restrict__ out, size_t K) {
    constexpr int G = TPB / E;               // row groups
    const int e = threadIdx.x % E, g = threadIdx.x / E;
    const size_t k0 = (size_t)blockIdx.x * E + e;
    if (k0 >= K) return;
    double acc = 0;
#pragma unroll
    for (int r = g; r < ROWS_IN; r += G) acc += in[(size_t)r * K + k0];
#pragma unroll
    for (int r = g; r < ROWS_OUT; r += G) out[(size_t)r * K + k0] = acc + r;
}
template <int E, int TPB> float run(const double* in, double* out, size_t K) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = (int)((K + E - 1) / E);
    k<E, 132, 132, TPB><<<blocks, TPB>>>(in, out, K);
    cudaEventRecord(a);
    for (int it = 0; it < 5; it++) k<E, 132, 132, TPB><<<blocks, TPB>>>(in, out, K);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    double gb = 2.0 * 132 * K * 8 / 1e9;
    printf("E=%3d TPB=%4d: %.3f ms  %.0f GB/s\n", E, TPB, ms, gb / (ms * 1e-3));
    return ms;
}
int main() {
    size_t K = 8000000;
    double *in, *out;
    cudaMalloc(&in, 132 * K * 8); cudaMalloc(&out, 132 * K * 8);
    cudaMemset(in, 0, 132 * K * 8);
    run<32, 128>(in, out, K); run<32, 256>(in, out, K); run<32, 512>(in, out, K);
    run<64, 128>(in, out, K); run<64, 256>(in, out, K); run<64, 512>(in, out, K);
    run<128, 128>(in, out, K); run<128, 256>(in, out, K); run<128, 512>(in, out, K);
    run<256, 256>(in, out, K); run<256, 1024>(in, out, K);
    // plain contiguous copy for reference
    return 0;
}
