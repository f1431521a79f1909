// Checks the operand/accumulator fragment layouts of mma.sync f64 on the running GPU against a host product:
//   m8n8k4  : A[8x4] lane l holds A[l/4][l%4];  B[4x8] lane l holds B[l%4][l/4];  C lane l holds C[l/4][2(l%4)], [..+1]
//             (the layout every DMMA kernel of this repo relies on)
//   m16n8k8 : hypothesis (PTX ISA, same pattern as the tf32 m16n8k8 tile):
//             a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]      (g = l/4, t = l%4)
//             b0 = B[t][g], b1 = B[t+4][g]
//             c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o dmma_layout dmma_layout.cu ; prints PASS/FAIL per shape.
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k884(const double *A, const double *B, double *C) {
    const int l = threadIdx.x, g = l >> 2, t = l & 3;
    double c0 = 0, c1 = 0;
    const double a = A[g * 4 + t], b = B[t * 8 + g];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    C[g * 8 + 2 * t] = c0;
    C[g * 8 + 2 * t + 1] = c1;
}

__global__ void k1688(const double *A, const double *B, double *C) {
    const int l = threadIdx.x, g = l >> 2, t = l & 3;
    double c[4] = {0, 0, 0, 0};
    const double a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    const double b0 = B[t * 8 + g], b1 = B[(t + 4) * 8 + g];
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    C[g * 8 + 2 * t] = c[0];
    C[g * 8 + 2 * t + 1] = c[1];
    C[(g + 8) * 8 + 2 * t] = c[2];
    C[(g + 8) * 8 + 2 * t + 1] = c[3];
}

static bool check(const char *name, int M, int K, void (*launch)(const double *, const double *, double *)) {
    double hA[16 * 8], hB[8 * 8], hC[16 * 8], ref[16 * 8];
    for (int i = 0; i < M * K; i++) hA[i] = std::sin(1.0 + 0.37 * i);
    for (int i = 0; i < K * 8; i++) hB[i] = std::cos(2.0 + 0.53 * i);
    for (int m = 0; m < M; m++)
        for (int n = 0; n < 8; n++) {
            double s = 0;
            for (int k = 0; k < K; k++) s += hA[m * K + k] * hB[k * 8 + n];
            ref[m * 8 + n] = s;
        }
    double *dA, *dB, *dC;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, sizeof(hC));
    cudaMemcpy(dA, hA, sizeof(double) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(double) * K * 8, cudaMemcpyHostToDevice);
    launch(dA, dB, dC);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hC, dC, sizeof(double) * M * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < M * 8; i++) worst = std::fmax(worst, std::fabs(hC[i] - ref[i]));
    const bool ok = e == cudaSuccess && worst < 1e-13;
    printf("%-8s %s  (max |diff| %.3e, %s)\n", name, ok ? "PASS" : "FAIL", worst, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return ok;
}

int main() {
    bool ok = check("m8n8k4", 8, 4, [](const double *a, const double *b, double *c) { k884<<<1, 32>>>(a, b, c); });
    ok &= check("m16n8k8", 16, 8, [](const double *a, const double *b, double *c) { k1688<<<1, 32>>>(a, b, c); });
    return ok ? 0 : 1;
}
