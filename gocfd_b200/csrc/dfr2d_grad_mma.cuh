// k_grad_mma<N>: phase P5 of the dissipation path (GetSolutionGradientUsingRTElement, euler.go:864-918, and the C0
// branch of CalculateEpsilonGradient, dissipation.go:244-272) with the Div contraction on the FP64 tensor cores.
//
// The DFMA version (k_grad, dfr2d_diss_kernels.cuh) is FP64-issue bound: 2 * NpFlux * (NpInt + 3 NpEdge) constant-operand
// DFMAs per variable (25 kflop per element at N=4, 41 % of the FP64 peak, 3.37 ms of the 8.8 ms stage on 2M triangles).
// Two changes:
//   1  GradX and GradY share the contraction.  DOFX_j = mX_j U_j and DOFY_j = mY_j U_j, and the metric of RT point j
//      takes only five values per element (CalculateRTBasedDerivativeMetrics, DG2D/dfr_startup.go:213-254): Jinv[0|1]
//      on the first interior block, Jinv[2|3] on the second, n_e IInII_e / Jdet on the points of edge e.  So with
//      S_r[row] = sum_{j in block r} Div[row][j] U_j  (one pass of Div over U, no metric),
//      GradX[row] = sum_r mX_r S_r[row],  GradY[row] = sum_r mY_r S_r[row]:  half the contraction flops.
//   2  S_r is a DMMA.8x8x4 product: A = Div rows (only the rows consumed later: [0,NpInt) and the 3 NpEdge edge rows),
//      zero padded to 8 x 4 tiles per block r and kept in shared memory in A-fragment lane order; B = U of one variable
//      for 8 elements (k = RT point, n = element), read from the same stride-36 shared-memory rows as k_elem_mma.
// CTA = 32 elements x 8 warps; warp = (conserved variable, 16-element half of the tile).  The metric combination,
// the Epsilon product (Bary . vertex eps, InterpolateEpsilonSigma dissipation.go:219-242) and the 128-bit stores of
// DissX / DissY happen in the accumulator fragment layout.  Opt-in until measured: DFR2D_GRAD_KERNEL=2.
#pragma once
#include "dfr2d_diss_kernels.cuh"
#include "dfr2d_elem_mma.cuh"

namespace dfr2d {

template <int N> struct GradMmaDim {
    static constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3;
    static constexpr int NOUT = NI + NF3, MT = (NOUT + 7) / 8;          // produced rows, m-tiles
    static constexpr int KI = (NI + 3) / 4, KE = (NEd + 3) / 4;         // k-steps of an interior / an edge block
    static constexpr int KS = 2 * KI + 3 * KE;                          // k-steps of all five blocks
    static constexpr int UROWS = 4 * KI + 12 * KE;                      // B rows per variable (interior block stored once)
    static constexpr int SE = kElemsPerBlock + 4;
    static constexpr int kFragDoubles = MT * KS * 32;
    static constexpr int kTableDoubles = kFragDoubles + 8 * MT * 3;     // + Bary of the produced rows
    static constexpr int MG = 3;                                        // m-tiles accumulated at a time
    static constexpr size_t kSmemBytes =
        (size_t)(4 * UROWS * SE + kTableDoubles + 10 * kElemsPerBlock + 3 * kElemsPerBlock) * sizeof(double);
    __host__ __device__ static constexpr int out_row(int m) { return m < NI ? m : m + NI; }
    __host__ __device__ static constexpr int blk_cols(int r) { return r < 2 ? NI : NEd; }
    __host__ __device__ static constexpr int blk_col0(int r) { return r < 2 ? r * NI : 2 * NI + (r - 2) * NEd; }
    __host__ __device__ static constexpr int blk_ks(int r) { return r < 2 ? KI : KE; }
    __host__ __device__ static constexpr int blk_k0(int r) { return r < 2 ? r * KI : 2 * KI + (r - 2) * KE; }
    __host__ __device__ static constexpr int blk_urow0(int r) { return r < 2 ? 0 : 4 * KI + (r - 2) * 4 * KE; }
};

constexpr int kGradMmaThreads = 8 * kElemsPerBlock;

template <int N>
__global__ void __launch_bounds__(kGradMmaThreads, 2) k_grad_mma(GradArgs a, const double *__restrict__ table) {
    using GD = GradMmaDim<N>;
    constexpr int NI = GD::NI, NEd = GD::NEd, NF = GD::NF, NF3 = GD::NF3, E = kElemsPerBlock, SE = GD::SE;
    constexpr int MT = GD::MT, KS = GD::KS, KI = GD::KI, KE = GD::KE, MG = GD::MG, UROWS = GD::UROWS;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    extern __shared__ double smem[];
    double *sU = smem;                                 // [4][UROWS][SE]
    double *sT = sU + 4 * UROWS * SE;                  // A fragments [MT][KS][32], then Bary [8 MT][3]
    double *sMet = sT + GD::kTableDoubles;             // [5 blocks][x|y][E]
    double *sEv = sMet + 10 * E;                       // [3][E] vertex epsilon
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, n = w & 3, half = w >> 2;
    const int k0 = blockIdx.x * E, k = k0 + lane;
    const int kc = k < a.K ? k : a.K - 1;
    const size_t Kp = a.Kp;

    for (int t = tid; t < GD::kTableDoubles; t += kGradMmaThreads) sT[t] = table[t];
    double *myU = sU + (size_t)n * UROWS * SE;
    // distinct solution values of the RT points: interior rows (both interior blocks read them), then per edge the
    // owner's Q_Face values (EdgeQValues), reversed for the non-owner (euler.go:896-912); rows split between the halves
#pragma unroll
    for (int i = 0; i < 4 * KI; i++)
        if ((i & 1) == half) myU[i * SE + lane] = (i < NI) ? a.q[((size_t)n * NI + i) * Kp + kc] : 0.0;
    const size_t qplane = (size_t)NF3 * Kp;
#pragma unroll
    for (int le = 0; le < 3; le++) {
        const int s = a.etoe[(size_t)le * Kp + kc];
        const bool owner = s >= 0;
        int kS = kc, numS = le;
        if (!owner) {
            const int slot = -1 - s;
            kS = a.ekL[slot];
            numS = a.emeta[slot] & 3;
        }
        const double *src = a.qface + n * qplane + (size_t)(numS * NEd) * Kp + kS;
#pragma unroll
        for (int i = 0; i < 4 * KE; i++)
            if ((i & 1) == half) {
                double v = 0.0;
                if (i < NEd) v = src[(size_t)(owner ? i : NEd - 1 - i) * Kp];
                myU[(4 * KI + le * 4 * KE + i) * SE + lane] = v;
            }
    }
    if (w == 0) {
        // metric of each block (DXMetric / DYMetric, DG2D/dfr_startup.go:213-254)
        const double oojd = 1.0 / a.Jdet[kc];
        sMet[0 * E + lane] = a.Jinv[0 * Kp + kc];
        sMet[1 * E + lane] = a.Jinv[1 * Kp + kc];
        sMet[2 * E + lane] = a.Jinv[2 * Kp + kc];
        sMet[3 * E + lane] = a.Jinv[3 * Kp + kc];
#pragma unroll
        for (int le = 0; le < 3; le++) {
            const double iin = a.IInII[(size_t)le * Kp + kc];
            sMet[(4 + 2 * le) * E + lane] = oojd * a.nxk[(size_t)le * Kp + kc] * iin;
            sMet[(5 + 2 * le) * E + lane] = oojd * a.nyk[(size_t)le * Kp + kc] * iin;
        }
    } else if (w == 4) {
#pragma unroll
        for (int v = 0; v < 3; v++) sEv[v * E + lane] = a.epsV[a.etov[(size_t)v * Kp + kc]];
    }
    __syncthreads();

    const int fr = lane >> 2, fc = lane & 3;
    const int nt0 = 2 * half;
#pragma unroll
    for (int m0 = 0; m0 < MT; m0 += MG) {
        double gx[MG][2][2], gy[MG][2][2];
#pragma unroll
        for (int mt = 0; mt < MG; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++) gx[mt][nt][0] = gx[mt][nt][1] = gy[mt][nt][0] = gy[mt][nt][1] = 0.0;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            double S[MG][2][2];
#pragma unroll
            for (int mt = 0; mt < MG; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++) S[mt][nt][0] = S[mt][nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < (KI > KE ? KI : KE); ks++) {
                if (ks >= GD::blk_ks(r)) continue;
                double b[2];
#pragma unroll
                for (int nt = 0; nt < 2; nt++) b[nt] = myU[(GD::blk_urow0(r) + 4 * ks + fc) * SE + 8 * (nt0 + nt) + fr];
#pragma unroll
                for (int mt = 0; mt < MG; mt++)
                    if (m0 + mt < MT) {
                        const double av = sT[((m0 + mt) * KS + GD::blk_k0(r) + ks) * 32 + lane];
#pragma unroll
                        for (int nt = 0; nt < 2; nt++) dmma884(S[mt][nt][0], S[mt][nt][1], av, b[nt]);
                    }
            }
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const int e0 = 8 * (nt0 + nt) + 2 * fc;
                const double2 mX = *reinterpret_cast<const double2 *>(&sMet[(2 * r) * E + e0]);
                const double2 mY = *reinterpret_cast<const double2 *>(&sMet[(2 * r + 1) * E + e0]);
#pragma unroll
                for (int mt = 0; mt < MG; mt++)
                    if (m0 + mt < MT) {
                        gx[mt][nt][0] = fma(mX.x, S[mt][nt][0], gx[mt][nt][0]);
                        gx[mt][nt][1] = fma(mX.y, S[mt][nt][1], gx[mt][nt][1]);
                        gy[mt][nt][0] = fma(mY.x, S[mt][nt][0], gy[mt][nt][0]);
                        gy[mt][nt][1] = fma(mY.y, S[mt][nt][1], gy[mt][nt][1]);
                    }
            }
        }
        // Diss = Epsilon (.) Grad on the rows of this group; lane holds row 8 mt + fr, elements e0, e0 + 1
#pragma unroll
        for (int mt = 0; mt < MG; mt++) {
            const int m = 8 * (m0 + mt) + fr;
            if (m0 + mt < MT && m < GD::NOUT) {
                const int row = GD::out_row(m);
                const double b0 = sT[GD::kFragDoubles + m * 3 + 0], b1 = sT[GD::kFragDoubles + m * 3 + 1],
                             b2 = sT[GD::kFragDoubles + m * 3 + 2];
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    const int e0 = 8 * (nt0 + nt) + 2 * fc;
                    const double2 v0 = *reinterpret_cast<const double2 *>(&sEv[0 * E + e0]);
                    const double2 v1 = *reinterpret_cast<const double2 *>(&sEv[1 * E + e0]);
                    const double2 v2 = *reinterpret_cast<const double2 *>(&sEv[2 * E + e0]);
                    const double epsA = b0 * v0.x + b1 * v1.x + b2 * v2.x;
                    const double epsB = b0 * v0.y + b1 * v1.y + b2 * v2.y;
                    const size_t o = ((size_t)n * NF + row) * Kp + k0 + e0;
                    if (k0 + e0 + 1 < a.K) {
                        *reinterpret_cast<double2 *>(a.dissX + o) = make_double2(gx[mt][nt][0] * epsA, gx[mt][nt][1] * epsB);
                        *reinterpret_cast<double2 *>(a.dissY + o) = make_double2(gy[mt][nt][0] * epsA, gy[mt][nt][1] * epsB);
                    } else if (k0 + e0 < a.K) {
                        a.dissX[o] = gx[mt][nt][0] * epsA;
                        a.dissY[o] = gy[mt][nt][0] * epsA;
                    }
                }
            }
        }
    }
}

// Host side: Div -> A-operand fragments per (m-tile, k-step) in lane order (lane l holds A[l/4][l%4]), k-steps grouped
// by metric block and zero padded per block; then the Bary rows of the produced RT rows.
template <int N> void build_grad_table(const double *Div, const double *Bary, std::vector<double> &out) {
    using GD = GradMmaDim<N>;
    out.assign(GD::kTableDoubles, 0.0);
    for (int mt = 0; mt < GD::MT; mt++)
        for (int r = 0; r < 5; r++)
            for (int ks = 0; ks < GD::blk_ks(r); ks++)
                for (int l = 0; l < 32; l++) {
                    const int m = 8 * mt + l / 4, c = 4 * ks + l % 4;
                    if (m < GD::NOUT && c < GD::blk_cols(r))
                        out[((size_t)mt * GD::KS + GD::blk_k0(r) + ks) * 32 + l] =
                            Div[(size_t)GD::out_row(m) * GD::NF + GD::blk_col0(r) + c];
                }
    for (int m = 0; m < GD::NOUT; m++)
        for (int c = 0; c < 3; c++) out[GD::kFragDoubles + m * 3 + c] = Bary[(size_t)GD::out_row(m) * 3 + c];
}

}  // namespace dfr2d
