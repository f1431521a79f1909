// Utility kernels (halo pack/unpack, residual reduction) and the artificial-dissipation path.
#pragma once
#include "dfr2d_kernels.cuh"

namespace dfr2d {

struct DissBuffers {
    int *etov = nullptr;                 // [3][Kp] global vertex ids
    double *hk = nullptr;                // [Kp]
    double *nxk = nullptr, *nyk = nullptr;   // [3][Kp] own face normals
    double *eooLen = nullptr;            // [NEp] 1 / edge length
    double *sigma = nullptr, *epsk = nullptr, *se = nullptr;   // [Kp]
    double *sigmaV = nullptr, *epsV = nullptr;                  // [NV]
    double *dissX = nullptr, *dissY = nullptr;                  // [4][NpFlux][Kp]
    double *vflux = nullptr;             // [4][NpEdge][NEp]
    double *aggv = nullptr;              // [NEp]
    double *DTVisc = nullptr;            // [Kp]
};

// message layout per cut edge: [4 vars][NpEdge points], in the sender's own edge-point order
__global__ void k_halo_pack(int total, int npEdge, int Kp, const double *qface, const int *elem, const int *row0, double *buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int per = 4 * npEdge;
    const int c = t / per, r = t % per, n = r / npEdge, i = r % npEdge;
    buf[t] = qface[((size_t)n * 3 * npEdge + row0[c] + i) * Kp + elem[c]];
}

__global__ void k_halo_unpack(int total, int npEdge, int Kp, double *qface, const int *col, const int *row0, const double *buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int per = 4 * npEdge;
    const int c = t / per, r = t % per, n = r / npEdge, i = r % npEdge;
    qface[((size_t)n * 3 * npEdge + row0[c] + i) * Kp + col[c]] = buf[t];
}

// Matrix.Max of each Residual[n] (utils/matrix_extended.go:1085-1096): signed max, one CTA per variable
__global__ void __launch_bounds__(1024) k_signed_max(const double *R, int npInt, int K, int Kp, double *out) {
    const int n = blockIdx.x;
    const double *r = R + (size_t)n * npInt * Kp;
    double m = r[0];
    for (size_t t = threadIdx.x; t < (size_t)npInt * K; t += blockDim.x) {
        const size_t i = t / K, k = t % K;
        const double v = r[i * Kp + k];
        if (v > m) m = v;
    }
    __shared__ double sm[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = sm[threadIdx.x];
        v = warp_max(v);
        if (threadIdx.x == 0) out[n] = v;
    }
}

template <int N> static size_t elem_smem_diss() { return (size_t)12 * Dim<N>::NpInt * kElemsPerBlock * sizeof(double); }

}  // namespace dfr2d
