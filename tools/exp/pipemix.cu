// Micro-benchmark (round 2): how the shared FP64 / DMMA pipe of an sm_100a SM behaves under the instruction MIX of the
// element and gradient kernels -- DMMA.8x8x4 interleaved with shared-memory operand loads (LDS.64 / LDS.128) and plain
// DFMA -- for 1, 2 and 4 warps per scheduler.  ncu shows both kernels at 55-60 % "pipe_shared" utilisation with
// mio_throttle / math_pipe_throttle as the top stalls of the tensor warps; this isolates the instruction stream from
// the copy pipeline and HBM.  Output: achieved FMA rate as a fraction of 64 FMA/clk/SM, and cycles per DMMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipemix pipemix.cu && ./pipemix
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE bits: 1 = A operand by LDS.64 per DMMA pair, 2 = B operand by LDS.64 per DMMA pair, 4 = A and B pairs by LDS.128
// NF = independent DFMAs issued per DMMA; NACC = accumulator chains (even)
template <int MODE, int NF, int NACC>
__global__ void __launch_bounds__(512, 1) k_mix(const double *in, double *out, int iters, long long *cyc) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 16384; i += blockDim.x) sm[i] = in[i & 1023];
    __syncthreads();
    double a[12], b[4], c[NACC][2], f[8];
    for (int i = 0; i < 12; i++) a[i] = in[(tid + i) & 1023];
    for (int i = 0; i < 4; i++) b[i] = in[(tid + 7 * i) & 1023];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0;
    for (int i = 0; i < 8; i++) f[i] = 0;
    unsigned off = (unsigned)lane;            // opaque per-iteration shared-memory offset (as in the kernels)
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        asm volatile("" : "+r"(off));
        const double *pA = sm + off, *pB = sm + 8192 + off;
#pragma unroll
        for (int ks = 0; ks < 12; ks++) {
#pragma unroll
            for (int t = 0; t < NACC; t += 2) {
                double av = a[ks], bv0 = b[(t + ks) & 3], bv1 = b[(t + ks + 1) & 3];
                if (MODE & 4) {
                    const double2 aa = *reinterpret_cast<const double2 *>(pA + lane + (ks * NACC + t) * 64);
                    av = aa.x; bv1 = aa.y;
                } else {
                    if (MODE & 1) av = pA[(ks * NACC + t) * 32];
                    if (MODE & 2) bv0 = pB[(ks * NACC + t) * 32];
                }
                dmma884(c[t][0], c[t][1], av, bv0);
                dmma884(c[t + 1][0], c[t + 1][1], av, bv1);
#pragma unroll
                for (int q = 0; q < 2 * NF; q++) {
                    const int ch = (q + 2 * NF * (t / 2 + ks * (NACC / 2))) % 8;       // eight independent DFMA chains, round robin
                    f[ch] = fma(f[ch], a[(ks + q) % 12], b[q & 3]);
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    for (int i = 0; i < 8; i++) s += f[i];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename K> void run(K kern, const char *name, int warps, int nacc, int nf, const double *in, double *out, long long *cyc) {
    const int iters = 400;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    kern<<<148, warps * 32, 16384 * 8>>>(in, out, 2, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<148, warps * 32, 16384 * 8>>>(in, out, iters, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double dmma_per_warp = 12.0 * nacc * iters, dfma_per_warp = dmma_per_warp * nf;
    const double pipe_cycles = warps * (dmma_per_warp * 4.0 + dfma_per_warp * 0.5);     // at 64 FMA/clk/SM
    printf("%-46s warps %2d  %7.3f ms  %9lld clk  pipe-cycle model %5.1f%% of elapsed, %.2f clk per DMMA(+%d DFMA) per SM  %s\n",
           name, warps, ms, c, 100.0 * pipe_cycles / (double)c, (double)c / (warps * dmma_per_warp), nf, cudaGetErrorString(e));
}

int main() {
    double *in, *out; long long *cyc;
    cudaMalloc(&in, 1024 * 8); cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8); cudaMemset(in, 0, 1024 * 8);
    for (int w : {4, 8, 16}) {
        run(k_mix<0, 0, 8>, "DMMA only, 8 chains", w, 8, 0, in, out, cyc);
        run(k_mix<1, 0, 8>, "DMMA + LDS.64 (A) per pair", w, 8, 0, in, out, cyc);
        run(k_mix<3, 0, 8>, "DMMA + LDS.64 (A) + LDS.64 (B) per pair", w, 8, 0, in, out, cyc);
        run(k_mix<4, 0, 8>, "DMMA + one LDS.128 per pair", w, 8, 0, in, out, cyc);
        run(k_mix<0, 1, 8>, "DMMA + 1 DFMA each", w, 8, 1, in, out, cyc);
        run(k_mix<0, 3, 8>, "DMMA + 3 DFMA each", w, 8, 3, in, out, cyc);
        run(k_mix<3, 3, 8>, "DMMA + 2 LDS.64 per pair + 3 DFMA each", w, 8, 3, in, out, cyc);
        run(k_mix<0, 0, 4>, "DMMA only, 4 chains", w, 4, 0, in, out, cyc);
        run(k_mix<0, 0, 2>, "DMMA only, 2 chains", w, 2, 0, in, out, cyc);
        run(k_mix<0, 8, 2>, "2 chains + 8 DFMA each", w, 2, 8, in, out, cyc);
    }
    return 0;
}
