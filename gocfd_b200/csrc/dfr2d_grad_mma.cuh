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
// Warp = (conserved variable, 16-element half of the tile).  The metric combination, the Epsilon product (Bary . vertex
// eps, InterpolateEpsilonSigma dissipation.go:219-242) and the stores happen in the accumulator fragment layout.  Round 1
// measured a one-tile-per-CTA version of this (k_grad_mma: 3.52 ms against 3.86 ms of the DFMA kernel for k_edge +
// gradient, 2M triangles, N=4; profiles/r01l_*); the pipelined persistent k_grad_pipe below (2.74 ms) replaced it.
#pragma once
#include "dfr2d_diss_kernels.cuh"
#include "dfr2d_elem_mma.cuh"
#include "dfr2d_elem_ws.cuh"

namespace dfr2d {

template <int N> struct GradMmaDim {
    static constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3;
    static constexpr int NOUT = NI + NF3;                                // produced rows
    // (r2) the produced rows fill NOUT / 8 m-tiles plus NOUT mod 8 ragged rows (N=4: 33 = 4 x 8 + 1; N=3: 25 = 3 x 8 + 1;
    // N=2: 18 = 2 x 8 + 2).  A single ragged row is evaluated by plain DFMA (grad_tail) instead of a DMMA m-tile that is
    // 7/8 zeros: 20-25 % fewer DMMA and A-fragment loads.
    // measured (2M triangles, profiles/r02m_ab_N*.json): N=4 2.44 -> 2.34 ms, N=3 1.84 -> 1.65 ms; with TWO ragged rows
    // (N=2) the DFMA pass costs more than the m-tile it replaces (1.12 -> 1.15 ms), so only a single ragged row is cut.
    static constexpr int NTAIL = (NOUT % 8 == 1) ? 1 : 0;
    static constexpr int MT = (NOUT - NTAIL + 7) / 8;                   // m-tiles on the tensor cores
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
                    if (m < GD::NOUT - GD::NTAIL && c < GD::blk_cols(r))
                        out[((size_t)mt * GD::KS + GD::blk_k0(r) + ks) * 32 + l] =
                            Div[(size_t)GD::out_row(m) * GD::NF + GD::blk_col0(r) + c];
                }
    for (int m = 0; m < GD::NOUT - GD::NTAIL; m++)
        for (int c = 0; c < 3; c++) out[GD::kFragDoubles + m * 3 + c] = Bary[(size_t)GD::out_row(m) * 3 + c];
}


// ------------------------------------------------------------------------------------------------------------------
// k_grad_pipe<N> (DFR2D_GRAD_KERNEL=3): k_grad_mma as a persistent, software-pipelined kernel.
//
// Measured on B200 (profiles/r01l_grad_ab_k_grad_vs_k_grad_mma.json): k_grad_mma is only 10 % faster than the DFMA kernel although its
// tensor-pipe floor is ~1 ms for 2M elements at N=4 -- a CTA loads its tile (three dependent global loads deep:
// etoe -> edge table -> Q_Face), synchronises, computes, stores; nothing overlaps and the operator table is re-read
// from L2 by every 32-element tile.  Here one CTA per SM holds the table for its lifetime and runs two independent
// groups of 8 warps; each group walks its own tiles through a two-stage shared-memory ring:
//   * every byte of tile t+1 (solution rows, gathered edge values, metric rows, vertex epsilon) is moved by cp.async
//     issued BEFORE the DMMA work of tile t, so it lands underneath it and never occupies a register;
//   * the index chain is pipelined one level per iteration: etoe of tile t+3 and the edge-table entries of tile t+2 are
//     plain loads whose results are first used one iteration later;
//   * per-edge metrics n_e IInII_e / Jdet are precomputed once at create ([6][Kp]) so they are plain rows as well;
//   * the two groups synchronise only among themselves (named barriers), so one group's epilogue stores overlap the
//     other's tensor work.
// ------------------------------------------------------------------------------------------------------------------
template <int N> struct GradPipeDim {
    using GD = GradMmaDim<N>;
    static constexpr int E = kElemsPerBlock;
    static constexpr int kGroups = 2, kGroupThreads = 8 * E, kThreads = kGroups * kGroupThreads;
    static constexpr int kUDoubles = 4 * GD::UROWS * GD::SE;
    static constexpr int kStageDoubles = kUDoubles + 10 * E + 3 * E + 9 * E;  // U, metric rows, vertex epsilon, and per local
                                                                          // edge: owner normal (x, y) and the etoe entry
    static constexpr size_t kSmemBytes = (size_t)(GD::kTableDoubles + 2 * kGroups * kStageDoubles) * sizeof(double);
};

struct GradPipeArgs {
    GradArgs a;
    const double *table;     // build_grad_table
    const double *mxy;       // [6][Kp]: (x, y) metric of the points of edge 0, 1, 2
    int nTiles;
    int skewNs;              // start delay of group 1: keeps the two groups' tensor phases out of step
};

__device__ __forceinline__ void group_bar(int group) {
    if (group == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 2, 256;" ::: "memory");
}

// One metric block of the contraction for the m-tiles [M0, M0+MG) and this warp's two n-tiles: S = Div_block . U_block
// (NKS k-steps of DMMA), then g += metric (.) S.  pUb / pAb / pMb point at the block's first B row, A fragment and
// metric row pair.  Called from ROLLED loops over the blocks: with the whole tile body unrolled, ptxas batches the
// shared-memory loads of many k-steps ahead of the DMMA stream and spills them.
template <int N, int MG, int M0, int NKS>
__device__ __forceinline__ void grad_block(const double *pUb, const double *pAb, const double *pMb, double (&gx)[MG][2][2],
                                           double (&gy)[MG][2][2]) {
    using GD = GradMmaDim<N>;
    constexpr int E = kElemsPerBlock, SE = GD::SE, KS = GD::KS, MT = GD::MT;
    double S[MG][2][2];
#pragma unroll
    for (int mt = 0; mt < MG; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) S[mt][nt][0] = S[mt][nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < NKS; ks++) {
        double b[2];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) b[nt] = pUb[4 * ks * SE + 8 * nt];
#pragma unroll
        for (int mt = 0; mt < MG; mt++)
            if (M0 + mt < MT) {
                const double av = pAb[((M0 + mt) * KS + ks) * 32];
#pragma unroll
                for (int nt = 0; nt < 2; nt++) dmma884(S[mt][nt][0], S[mt][nt][1], av, b[nt]);
            }
    }
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
        const double2 mX = *reinterpret_cast<const double2 *>(&pMb[8 * nt]);
        const double2 mY = *reinterpret_cast<const double2 *>(&pMb[E + 8 * nt]);
#pragma unroll
        for (int mt = 0; mt < MG; mt++)
            if (M0 + mt < MT) {
                gx[mt][nt][0] = fma(mX.x, S[mt][nt][0], gx[mt][nt][0]);
                gx[mt][nt][1] = fma(mX.y, S[mt][nt][1], gx[mt][nt][1]);
                gy[mt][nt][0] = fma(mY.x, S[mt][nt][0], gy[mt][nt][0]);
                gy[mt][nt][1] = fma(mY.y, S[mt][nt][1], gy[mt][nt][1]);
            }
    }
}

// unroll factors of the two block loops of grad_mgroup (1 = rolled).  Measured with k_grad_ws<N,2> on 2M triangles
// (profiles/r02ab_*): the two interior blocks unrolled (their DMMA chains and metric FMAs overlap) 2.66 -> 2.59 ms at N=4,
// 2.04 -> 1.98 ms at N=3 for gradient + fused interior edges; unrolling the three edge blocks instead is slower (2.73 /
// 2.06 ms), both together spill (288 B of local memory at 104 registers).
#ifndef DFR2D_GRAD_UNROLL_INT
#define DFR2D_GRAD_UNROLL_INT 2
#endif
#ifndef DFR2D_GRAD_UNROLL_EDGE
#define DFR2D_GRAD_UNROLL_EDGE 1
#endif

// Rows [M0, M0+MG) of one tile: all five blocks, Epsilon product, stores.
template <int N, int MG, int M0>
__device__ __forceinline__ void grad_mgroup(const GradArgs &a, const double *pU, const double *pA, const double *pM,
                                            const double *pB, int n, int fr, int fc, int nt0, int k0, size_t KpL) {
    using GD = GradMmaDim<N>;
    constexpr int E = kElemsPerBlock, SE = GD::SE, KI = GD::KI, KE = GD::KE, MT = GD::MT;
    double gx[MG][2][2], gy[MG][2][2];
#pragma unroll
    for (int mt = 0; mt < MG; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) gx[mt][nt][0] = gx[mt][nt][1] = gy[mt][nt][0] = gy[mt][nt][1] = 0.0;
    constexpr int kUnrollInt = DFR2D_GRAD_UNROLL_INT, kUnrollEdge = DFR2D_GRAD_UNROLL_EDGE;
#pragma unroll kUnrollInt
    for (int r = 0; r < 2; r++)                 // the two interior blocks share the B rows
        grad_block<N, MG, M0, KI>(pU, pA + r * KI * 32, pM + 2 * r * E, gx, gy);
#pragma unroll kUnrollEdge
    for (int le = 0; le < 3; le++)              // points of edge le
        grad_block<N, MG, M0, KE>(pU + (4 * KI + le * 4 * KE) * SE, pA + (2 * KI + le * KE) * 32, pM + (4 + 2 * le) * E, gx, gy);
    // vertex epsilon of this lane's element pairs (rows 10..12 of the metric block)
    double2 ev[2][3];
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int v = 0; v < 3; v++) ev[nt][v] = *reinterpret_cast<const double2 *>(&pM[(10 + v) * E + 8 * nt]);
#pragma unroll
    for (int mt = 0; mt < MG; mt++) {
        const int m = 8 * (M0 + mt) + fr;
        if (M0 + mt < MT && m < GD::NOUT - GD::NTAIL) {
            const int row = GD::out_row(m);
            const double b0 = pB[8 * (M0 + mt) * 3], b1 = pB[8 * (M0 + mt) * 3 + 1], b2 = pB[8 * (M0 + mt) * 3 + 2];
            constexpr int NI = GD::NI, NEd = GD::NEd;
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const int e0 = 8 * (nt0 + nt) + 2 * fc;
                const double epsA = b0 * ev[nt][0].x + b1 * ev[nt][1].x + b2 * ev[nt][2].x;
                const double epsB = b0 * ev[nt][0].y + b1 * ev[nt][1].y + b2 * ev[nt][2].y;
                if (m < NI) {
                    // interior rows: DissX / DissY [4][NpInt][Kp], read back by AddDissipation in the element kernel
                    const size_t o = ((size_t)n * NI + row) * KpL + k0 + e0;
                    if (k0 + e0 + 1 < a.K) {
                        *reinterpret_cast<double2 *>(a.dissX + o) = make_double2(gx[mt][nt][0] * epsA, gx[mt][nt][1] * epsB);
                        *reinterpret_cast<double2 *>(a.dissY + o) = make_double2(gy[mt][nt][0] * epsA, gy[mt][nt][1] * epsB);
                    } else if (k0 + e0 < a.K) {
                        a.dissX[o] = gx[mt][nt][0] * epsA;
                        a.dissY[o] = gy[mt][nt][0] * epsA;
                    }
                } else {
                    // edge rows: owner-normal component into the edge slot (GradArgs::vn), owner's point order
                    const int le = (m - NI) / NEd, ii = (m - NI) - le * NEd;
                    const double2 nxo = *reinterpret_cast<const double2 *>(&pM[(13 + le) * E + 8 * nt]);
                    const double2 nyo = *reinterpret_cast<const double2 *>(&pM[(16 + le) * E + 8 * nt]);
                    const double2 sl = *reinterpret_cast<const double2 *>(&pM[(19 + le) * E + 8 * nt]);
                    const long long sA = __double_as_longlong(sl.x), sB = __double_as_longlong(sl.y);
                    if (k0 + e0 < a.K) {
                        const bool own = sA >= 0;
                        const size_t slot = (size_t)(own ? sA : -1 - sA);
                        a.vn[((size_t)((own ? 0 : 4) + n) * NEd + (own ? ii : NEd - 1 - ii)) * a.NEp + slot] =
                            nxo.x * (gx[mt][nt][0] * epsA) + nyo.x * (gy[mt][nt][0] * epsA);
                    }
                    if (k0 + e0 + 1 < a.K) {
                        const bool own = sB >= 0;
                        const size_t slot = (size_t)(own ? sB : -1 - sB);
                        a.vn[((size_t)((own ? 0 : 4) + n) * NEd + (own ? ii : NEd - 1 - ii)) * a.NEp + slot] =
                            nxo.y * (gx[mt][nt][1] * epsB) + nyo.y * (gy[mt][nt][1] * epsB);
                    }
                }
            }
        }
    }
}

// The NTAIL ragged rows (all of them edge rows) of one tile by DFMA.  Warp = (variable n, 16-element half); lane l works
// on element 16 half + (l & 15); the two half warps split the five metric blocks (part 0: interior block 0 and edges 0, 1;
// part 1: interior block 1 and edge 2) and combine with one shuffle.  Div and Bary entries are constant-bank operands
// (compile-time indices), the U values / metrics / vertex epsilons / slot data are read from the tile's stage.
template <int N>
__device__ __forceinline__ void grad_tail(const GradArgs &a, const double *pUn /* stage U rows of variable n */,
                                          const double *pMs /* stage metric block */, int n, int half, int lane, int k0) {
    using GD = GradMmaDim<N>;
    constexpr int NI = GD::NI, NEd = GD::NEd, E = kElemsPerBlock, SE = GD::SE, KI = GD::KI, KE = GD::KE;
    const Ops<N> &op = ops<N>();
    const int e = 16 * half + (lane & 15), part = lane >> 4;
#pragma unroll
    for (int t = 0; t < GD::NTAIL; t++) {
        const int m = 8 * GD::MT + t;
        const int row = GD::out_row(m);
        const int le = (m - NI) / NEd, ii = (m - NI) - le * NEd;
        double gx = 0.0, gy = 0.0;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            if ((r == 0 || r == 2 || r == 3) != (part == 0)) continue;
            double S = 0.0;
            const int urow0 = GD::blk_urow0(r), col0 = GD::blk_col0(r);
#pragma unroll
            for (int j = 0; j < GD::blk_cols(r); j++) S = fma(op.Div[row][col0 + j], pUn[(urow0 + j) * SE + e], S);
            gx = fma(pMs[(2 * r) * E + e], S, gx);
            gy = fma(pMs[(2 * r + 1) * E + e], S, gy);
        }
        gx += __shfl_xor_sync(0xffffffffu, gx, 16);
        gy += __shfl_xor_sync(0xffffffffu, gy, 16);
        if (part == 0 && k0 + e < a.K) {
            const double eps = op.Bary[row][0] * pMs[10 * E + e] + op.Bary[row][1] * pMs[11 * E + e] + op.Bary[row][2] * pMs[12 * E + e];
            const long long s = __double_as_longlong(pMs[(19 + le) * E + e]);
            const bool own = s >= 0;
            const size_t slot = (size_t)(own ? s : -1 - s);
            a.vn[((size_t)((own ? 0 : 4) + n) * NEd + (own ? ii : NEd - 1 - ii)) * a.NEp + slot] =
                pMs[(13 + le) * E + e] * (gx * eps) + pMs[(16 + le) * E + e] * (gy * eps);
        }
    }
    (void)KI; (void)KE;
}

template <int N, int MG>      // MG: m-tiles accumulated at a time (2 or 3)
__global__ void __launch_bounds__(GradPipeDim<N>::kThreads, 1) k_grad_pipe(GradPipeArgs args) {
    using GD = GradMmaDim<N>;
    using PD = GradPipeDim<N>;
    constexpr int NI = GD::NI, NEd = GD::NEd, NF3 = GD::NF3, E = kElemsPerBlock, SE = GD::SE;
    constexpr int MT = GD::MT, KI = GD::KI, KE = GD::KE, UROWS = GD::UROWS;
    const GradArgs &a = args.a;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    extern __shared__ double smem[];
    double *sT = smem;                                              // operator table, whole CTA
    const int group = threadIdx.x / PD::kGroupThreads, tg = threadIdx.x % PD::kGroupThreads;
    double *sStage = sT + GD::kTableDoubles + (size_t)group * 2 * PD::kStageDoubles;   // this group's two stages
    const int lane = tg & 31, wg = tg >> 5, n = wg & 3, half = wg >> 2;
    const size_t Kp = a.Kp;
    const int nTiles = args.nTiles;
    const int stride = gridDim.x * PD::kGroups;
    const int tile0 = blockIdx.x * PD::kGroups + group;

    for (int t = threadIdx.x; t < GD::kTableDoubles; t += PD::kThreads) sT[t] = args.table[t];
    // padding rows of both stages are zeroed once; the copies only ever write the real rows
    for (int st = 0; st < 2; st++) {
        double *u = sStage + (size_t)st * PD::kStageDoubles + (size_t)n * UROWS * SE;
        for (int r = half; r < UROWS; r += 2) {
            const bool pad = (r < 4 * KI) ? (r >= NI) : (((r - 4 * KI) % (4 * KE)) >= NEd);
            if (pad) u[r * SE + lane] = 0.0;
        }
    }
    __syncthreads();

    // ---- pipeline registers ---------------------------------------------------------------------------------------
    int sIdx[3];                 // etoe of the tile three ahead
    int eKL[3], eMeta[3];        // raw edge-table entries of the tile two ahead (owner column, meta); decoded at issue
                                 // time, one iteration after the loads were posted, so that nothing waits on them
    int eKc = 0;                 // this lane's (clamped) element of that tile
    unsigned evIdx = 0;          // vertex of this lane's element (warps 5..7 of the group: vertex wg - 5)
    int eSlot = 0;               // etoe entry of local edge wg - 5 of this lane's element (same warps): where the edge rows go

    auto elem_of = [&](int tile) {
        const int tc = tile < nTiles ? tile : nTiles - 1;
        const int k = tc * E + lane;
        return k < a.K ? k : a.K - 1;
    };
    auto load_s = [&](int tile) {
        const int kc = elem_of(tile);
#pragma unroll
        for (int le = 0; le < 3; le++) sIdx[le] = a.etoe[(size_t)le * Kp + kc];
    };
    auto load_l2 = [&](int tile) {          // consumes sIdx (loaded one iteration earlier)
        eKc = elem_of(tile);
#pragma unroll
        for (int le = 0; le < 3; le++) {
            const int s = sIdx[le];
            eKL[le] = -1;
            eMeta[le] = 0;
            if (s < 0) {
                eKL[le] = a.ekL[-1 - s];
                eMeta[le] = a.emeta[-1 - s];
            }
        }
        if (wg >= 5) {
            evIdx = (unsigned)a.etov[(size_t)(wg - 5) * Kp + eKc];
            eSlot = (wg == 5) ? sIdx[0] : ((wg == 6) ? sIdx[1] : sIdx[2]);
        }
    };
    auto issue = [&](int tile, int st) {    // every byte of `tile` -> stage st, asynchronously
        if (tile < nTiles) {
            const int k0 = tile * E;
            double *base = sStage + (size_t)st * PD::kStageDoubles;
            unsigned uB = smem_u32(base);
            size_t KpL = Kp;
            // keep ptxas from hoisting (and spilling) every loop-invariant address of the unrolled copy lists
            asm volatile("" : "+r"(uB), "+l"(KpL));
            // solution rows: 4 NI rows of 256 B as 16-byte chunks
            for (int c = tg; c < 4 * NI * 16; c += PD::kGroupThreads) {
                const int row = c >> 4, ch = c & 15, var = row / NI, i = row - var * NI;
                cp_async16_u32(uB + (unsigned)(((var * UROWS + i) * SE + 2 * ch) * sizeof(double)),
                               a.q + ((size_t)var * NI + i) * KpL + k0 + 2 * ch);
            }
            // edge values of this lane's element, variable n, rows of parity `half`
            const double *qf = a.qface + (size_t)(n * NF3) * KpL;
#pragma unroll
            for (int le = 0; le < 3; le++) {
                // owner: own edge rows; neighbour: the owner's rows, points reversed (euler.go:896-912)
                const bool rev = eKL[le] >= 0;
                const double *src = qf + (rev ? (size_t)((eMeta[le] & 3) * NEd) * KpL + eKL[le] : (size_t)(le * NEd) * KpL + eKc);
#pragma unroll
                for (int i = 0; i < NEd; i++)
                    if ((i & 1) == half)
                        cp_async8_u32(uB + (unsigned)(((n * UROWS + 4 * KI + le * 4 * KE + i) * SE + lane) * sizeof(double)),
                                      src + (size_t)(rev ? NEd - 1 - i : i) * KpL);
            }
            // metric rows: Jinv[4][Kp], mxy[6][Kp]
            if (tg < 160) {
                const int row = tg >> 4, ch = tg & 15;
                const double *src = (row < 4 ? a.Jinv + (size_t)row * KpL : args.mxy + (size_t)(row - 4) * KpL) + k0 + 2 * ch;
                cp_async16_u32(uB + (unsigned)((PD::kUDoubles + row * E + 2 * ch) * sizeof(double)), src);
            } else {
                cp_async8_u32(uB + (unsigned)((PD::kUDoubles + 10 * E + (wg - 5) * E + lane) * sizeof(double)), a.epsV + evIdx);
                // local edge wg - 5 of this lane's element: owner normal of its slot and the etoe entry itself (as the
                // bit pattern of a 64-bit integer, so that the epilogue reads it with the same double2 pattern)
                const int slot = eSlot >= 0 ? eSlot : -1 - eSlot;
                cp_async8_u32(uB + (unsigned)((PD::kUDoubles + (13 + (wg - 5)) * E + lane) * sizeof(double)), a.enx + slot);
                cp_async8_u32(uB + (unsigned)((PD::kUDoubles + (16 + (wg - 5)) * E + lane) * sizeof(double)), a.eny + slot);
                base[PD::kUDoubles + (19 + (wg - 5)) * E + lane] = __longlong_as_double((long long)eSlot);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- prologue -----------------------------------------------------------------------------------------------------
    // Both groups run the same instruction stream on the same amount of work, so started together they stay in lock
    // step: both in the DMMA phase (fighting for the one FP64/tensor pipe of the SM), then both in the epilogue with the
    // pipe idle.  Delaying group 1 by about half a tile time interleaves the phases.
    if (group == 1 && args.skewNs > 0) __nanosleep((unsigned)args.skewNs);
    load_s(tile0);
    load_l2(tile0);
    issue(tile0, 0);
    load_s(tile0 + stride);
    load_l2(tile0 + stride);
    load_s(tile0 + 2 * stride);
    asm volatile("cp.async.wait_all;" ::: "memory");
    group_bar(group);

    const int fr = lane >> 2, fc = lane & 3;
    const int nt0 = 2 * half;
    // Shared-memory addressing of the steady state: every read is smem[opaque per-iteration offset + compile-time
    // constant], i.e. LDS [R + imm].  Left alone, the compiler hoists the ~200 per-fragment addresses of the unrolled
    // body out of the tile loop and spills them (1 KB of local memory per thread); laundering a POINTER instead turns
    // the loads into generic LD (long-scoreboard stalls, profiles/r01n).
    const unsigned groupOff = (unsigned)(GD::kTableDoubles + group * 2 * PD::kStageDoubles);
    const unsigned laneU = (unsigned)(n * UROWS * SE + fc * SE + fr + 8 * nt0);          // B fragments
    const unsigned laneM = (unsigned)(PD::kUDoubles + 8 * nt0 + 2 * fc);                 // metric / vertex-eps pairs
    const unsigned laneB = (unsigned)(GD::kFragDoubles + fr * 3);                        // Bary rows
    int st = 0;
    for (int tile = tile0; tile < nTiles; tile += stride, st ^= 1) {
        issue(tile + stride, st ^ 1);              // uses eKL / eMeta / eKc / evIdx of tile + stride
        load_l2(tile + 2 * stride);                // uses sIdx of tile + 2 stride
        load_s(tile + 3 * stride);

        const int k0 = tile * E;
        unsigned oU = groupOff + (unsigned)st * (unsigned)PD::kStageDoubles + laneU;
        unsigned oM = groupOff + (unsigned)st * (unsigned)PD::kStageDoubles + laneM;
        unsigned oA = (unsigned)lane, oB = laneB;
        size_t KpL = Kp;
        asm volatile("" : "+r"(oU), "+r"(oM), "+r"(oA), "+r"(oB), "+l"(KpL));
        const double *pU = smem + oU, *pM = smem + oM, *pA = smem + oA, *pB = smem + oB;
        grad_mgroup<N, MG, 0>(a, pU, pA, pM, pB, n, fr, fc, nt0, k0, KpL);
        if (MG < MT) grad_mgroup<N, MG, MG>(a, pU, pA, pM, pB, n, fr, fc, nt0, k0, KpL);
        if (2 * MG < MT) grad_mgroup<N, MG, 2 * MG>(a, pU, pA, pM, pB, n, fr, fc, nt0, k0, KpL);
        static_assert(3 * MG >= MT, "m-groups");
        if (GD::NTAIL > 0) {
            unsigned oT = groupOff + (unsigned)st * (unsigned)PD::kStageDoubles;
            asm volatile("" : "+r"(oT));
            grad_tail<N>(a, smem + oT + n * UROWS * SE, smem + oT + PD::kUDoubles, n, half, lane, k0);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        group_bar(group);          // stage st^1 is complete and every warp of the group has finished reading stage st
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_grad_ws<N, MG> (DFR2D_GRAD_KERNEL=4): k_grad_pipe with the data movement taken out of the tensor-core warps, the way
// k_elem_ws did it for the element kernel.  ncu on k_grad_pipe (profiles/r02i_*, r02z_*): the shared FP64/DMMA pipe is only
// 45 % busy and the 16 warps issue 0.31 instructions per cycle and scheduler; every warp spends its tile on 1,850
// instructions of which only 112 are DMMA -- the rest is the copy list of the next tile with its index chain, address
// arithmetic, and two group-wide synchronisations (cp.async.wait_all + bar.sync) that keep the eight warps of a group in
// lock step, so that they all want the pipe at once (math_pipe_throttle) and then all leave it idle.  Here
//   * TWO producer warps per group issue every copy of the group's next tile -- one streams the dense rows, the other owns
//     the index chain (etoe -> edge table -> Q_Face rows, vertex ids, slots) and the gathers; the copies complete on the
//     stage's FULL mbarrier;
//   * the eight consumer warps of a group (variable, 16-element half) only wait for FULL, run the unchanged DMMA / epilogue
//     / ragged-row code, and arrive on EMPTY.  They never write shared memory and never synchronise with each other, so
//     their DMMA bursts drift apart.
// Same stage layout, table, accumulation groups and results (bitwise) as k_grad_pipe.
template <int N> struct GradWsDim {
    using PD = GradPipeDim<N>;
    // two producer warps per group (register files are handed out per 4 warps: 18 warps would be allotted as 20 anyway):
    // role 0 streams the dense rows (solution, metrics), role 1 owns the index chain and the 8-byte gathers
    static constexpr int kConsWarps = 16, kProdWarps = 2 * PD::kGroups, kThreads = (kConsWarps + kProdWarps) * 32;
    static constexpr int kFullCount = 128;         // per producer lane (2 warps): its cp.async completions + one plain arrive
    // 640 threads are allotted 96 registers each; the producers keep 40 and hand the rest to the consumers
#ifndef DFR2D_GRAD_CONS_REGS
#define DFR2D_GRAD_CONS_REGS 104
#endif
    static constexpr int kProdRegs = 40, kConsRegs = DFR2D_GRAD_CONS_REGS;
    // setmaxnreg only redistributes what the CTA was given at launch (65,536 / threads, rounded down to 8): a consumer request
    // beyond that waits for registers nobody will ever release -- the kernel hangs (measured the hard way at 112)
    static constexpr int kLaunchRegs = (65536 / kThreads) & ~7;
    static_assert(kConsWarps * 32 * kConsRegs + kProdWarps * 32 * kProdRegs <= kLaunchRegs * kThreads, "setmaxnreg budget");
};

template <int N, int MG>
__global__ void __launch_bounds__(GradWsDim<N>::kThreads, 1) k_grad_ws(GradPipeArgs args) {
    using GD = GradMmaDim<N>;
    using PD = GradPipeDim<N>;
    using WD = GradWsDim<N>;
    constexpr int NI = GD::NI, NEd = GD::NEd, NF3 = GD::NF3, E = kElemsPerBlock, SE = GD::SE;
    constexpr int MT = GD::MT, KI = GD::KI, KE = GD::KE, UROWS = GD::UROWS;
    const GradArgs &a = args.a;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    extern __shared__ double smem[];
    __shared__ __align__(8) unsigned long long fullBar[PD::kGroups][2], emptyBar[PD::kGroups][2];
    double *sT = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp >= WD::kConsWarps;
    const int group = producer ? (warp - WD::kConsWarps) >> 1 : warp >> 3;
    double *sStage = sT + GD::kTableDoubles + (size_t)group * 2 * PD::kStageDoubles;
    const size_t Kp = a.Kp;
    const int nTiles = args.nTiles;
    const int stride = gridDim.x * PD::kGroups;
    const int tile0 = blockIdx.x * PD::kGroups + group;

    for (int t = threadIdx.x; t < GD::kTableDoubles; t += WD::kThreads) sT[t] = args.table[t];
    // padding rows of every stage are zeroed once; the copies only ever write the real rows
    for (int t = threadIdx.x; t < PD::kGroups * 2 * 4 * UROWS * E; t += WD::kThreads) {
        const int col = t % E, r = (t / E) % UROWS, rest = t / (E * UROWS);          // rest = (group, stage, variable)
        const bool pad = (r < 4 * KI) ? (r >= NI) : (((r - 4 * KI) % (4 * KE)) >= NEd);
        if (pad) sT[GD::kTableDoubles + (size_t)(rest >> 2) * PD::kStageDoubles + ((rest & 3) * UROWS + r) * SE + col] = 0.0;
    }
    if (threadIdx.x == 0) {
        for (int g = 0; g < PD::kGroups; g++)
            for (int st = 0; st < 2; st++) {
                mbar_init(&fullBar[g][st], WD::kFullCount);
                mbar_init(&emptyBar[g][st], 8);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (producer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WD::kProdRegs));
        const int role = (warp - WD::kConsWarps) & 1;
        if (role == 0) {
            // ================================ dense rows: solution slabs and metric rows ==================================
            int it = 0;
            for (int tile = tile0; tile < nTiles; tile += stride, it++) {
                const int st = it & 1;
                if (it >= 2) mbar_wait(&emptyBar[group][st], (unsigned)(((it >> 1) + 1) & 1));
                const int k0 = tile * E;
                const unsigned uB = smem_u32(sStage + (size_t)st * PD::kStageDoubles);
#pragma unroll
                for (int v = 0; v < 4; v++)
                    slab_g2s<NI, SE>(uB + (unsigned)((v * UROWS) * SE * sizeof(double)), a.q + (size_t)v * NI * Kp + k0, Kp, lane);
                slab_g2s<4, E>(uB + (unsigned)(PD::kUDoubles * sizeof(double)), a.Jinv + k0, Kp, lane);
                slab_g2s<6, E>(uB + (unsigned)((PD::kUDoubles + 4 * E) * sizeof(double)), args.mxy + k0, Kp, lane);
                cp_async_arrive_noinc(&fullBar[group][st]);
                mbar_arrive(&fullBar[group][st]);
            }
        } else {
            // ================================ index chain and gathers: lane = element =====================================
            int sIdx[3], vIdx[3];            // etoe / etov of this lane's element in the NEXT tile (prefetched one tile ahead)
            auto elem_of = [&](int tile) {
                const int tc = tile < nTiles ? tile : nTiles - 1;
                const int k = tc * E + lane;
                return k < a.K ? k : a.K - 1;
            };
            auto load_idx = [&](int tile) {
                const int kc = elem_of(tile);
#pragma unroll
                for (int le = 0; le < 3; le++) {
                    sIdx[le] = a.etoe[(size_t)le * Kp + kc];
                    vIdx[le] = a.etov[(size_t)le * Kp + kc];
                }
            };
            load_idx(tile0);
            int it = 0;
            for (int tile = tile0; tile < nTiles; tile += stride, it++) {
                const int st = it & 1;
                if (it >= 2) mbar_wait(&emptyBar[group][st], (unsigned)(((it >> 1) + 1) & 1));
                const int kc = elem_of(tile);
                double *base = sStage + (size_t)st * PD::kStageDoubles;
                const unsigned uB = smem_u32(base);
                // edge-table entries of the edges this element does not own (second level of the index chain)
                int eKL[3], eNum[3];
#pragma unroll
                for (int le = 0; le < 3; le++) {
                    eKL[le] = -1; eNum[le] = 0;
                    if (sIdx[le] < 0) { eKL[le] = a.ekL[-1 - sIdx[le]]; eNum[le] = a.emeta[-1 - sIdx[le]] & 3; }
                }
                // per local edge: vertex epsilon, owner normal of the slot, the etoe entry (as the bits of a 64-bit integer)
#pragma unroll
                for (int le = 0; le < 3; le++) {
                    const int slot = sIdx[le] >= 0 ? sIdx[le] : -1 - sIdx[le];
                    cp_async8_u32(uB + (unsigned)((PD::kUDoubles + (10 + le) * E + lane) * sizeof(double)), a.epsV + vIdx[le]);
                    cp_async8_u32(uB + (unsigned)((PD::kUDoubles + (13 + le) * E + lane) * sizeof(double)), a.enx + slot);
                    cp_async8_u32(uB + (unsigned)((PD::kUDoubles + (16 + le) * E + lane) * sizeof(double)), a.eny + slot);
                    base[PD::kUDoubles + (19 + le) * E + lane] = __longlong_as_double((long long)sIdx[le]);
                }
                // edge values: own edge rows, or the owner's rows with the points reversed (euler.go:896-912)
#pragma unroll
                for (int le = 0; le < 3; le++) {
                    const bool rev = eKL[le] >= 0;
                    const double *src0 = a.qface + (rev ? (size_t)(eNum[le] * NEd) * Kp + eKL[le] : (size_t)(le * NEd) * Kp + kc);
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        const double *src = src0 + (size_t)(v * NF3) * Kp;
#pragma unroll
                        for (int i = 0; i < NEd; i++)
                            cp_async8_u32(uB + (unsigned)(((v * UROWS + 4 * KI + le * 4 * KE + i) * SE + lane) * sizeof(double)),
                                          src + (size_t)(rev ? NEd - 1 - i : i) * Kp);
                    }
                }
                cp_async_arrive_noinc(&fullBar[group][st]);
                mbar_arrive(&fullBar[group][st]);
                load_idx(tile + stride);
            }
        }
    } else {
        // ================================ consumer warps: (variable n, 16-element half) ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WD::kConsRegs));
        const int wg = warp & 7, n = wg & 3, half = wg >> 2;
        const int fr = lane >> 2, fc = lane & 3;
        const int nt0 = 2 * half;
        const unsigned groupOff = (unsigned)(GD::kTableDoubles + group * 2 * PD::kStageDoubles);
        const unsigned laneU = (unsigned)(n * UROWS * SE + fc * SE + fr + 8 * nt0);          // B fragments
        const unsigned laneM = (unsigned)(PD::kUDoubles + 8 * nt0 + 2 * fc);                 // metric / vertex-eps pairs
        const unsigned laneB = (unsigned)(GD::kFragDoubles + fr * 3);                        // Bary rows
        int it = 0;
        for (int tile = tile0; tile < nTiles; tile += stride, it++) {
            const int st = it & 1;
            mbar_wait(&fullBar[group][st], (unsigned)((it >> 1) & 1));
            const int k0 = tile * E;
            unsigned oU = groupOff + (unsigned)st * (unsigned)PD::kStageDoubles + laneU;
            unsigned oM = groupOff + (unsigned)st * (unsigned)PD::kStageDoubles + laneM;
            unsigned oA = (unsigned)lane, oB = laneB;
            size_t KpL = Kp;
            asm volatile("" : "+r"(oU), "+r"(oM), "+r"(oA), "+r"(oB), "+l"(KpL));
            const double *pU = smem + oU, *pM = smem + oM, *pA = smem + oA, *pB = smem + oB;
            grad_mgroup<N, MG, 0>(a, pU, pA, pM, pB, n, fr, fc, nt0, k0, KpL);
            if (MG < MT) grad_mgroup<N, MG, MG>(a, pU, pA, pM, pB, n, fr, fc, nt0, k0, KpL);
            if (2 * MG < MT) grad_mgroup<N, MG, 2 * MG>(a, pU, pA, pM, pB, n, fr, fc, nt0, k0, KpL);
            if (GD::NTAIL > 0) {
                unsigned oT = groupOff + (unsigned)st * (unsigned)PD::kStageDoubles;
                asm volatile("" : "+r"(oT));
                grad_tail<N>(a, smem + oT + n * UROWS * SE, smem + oT + PD::kUDoubles, n, half, lane, k0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyBar[group][st]);
        }
    }
}

}  // namespace dfr2d
