// k_elem_pipe<N>: the inviscid element kernel as ONE persistent, software-pipelined CTA per SM.
//
// Measurements that led here (profiles/r01b_*, r01c_*): every earlier variant of the element kernel spent
// ~12-15 us per 32-element tile whatever the arithmetic cost, because a tile needed 4-5 *dependent* trips to
// HBM (edge slot -> edge flux gather; stage input; geometry; RK registers) with only 2-4 tiles resident per SM,
// so bytes in flight per SM stayed at about half of what Little's law needs for 6.4 TB/s.  Here a CTA of 16
// warps owns an SM and, while it computes tile t out of shared memory, the loads of tile t+1 (and the edge
// slots of tile t+2) are already in flight into registers.  The two thin contractions run on the FP64 tensor
// cores (DMMA.8x8x4, see dfr2d_elem_mma.cuh for why).
//
// Per tile:  registers(t) -> smem (stage input rows, sign/IInII-scaled edge DOFs); cp.async of the RK registers(t)
//            issue loads(t+1) -> registers ; issue edge-slot loads(t+2)
//            flux per point -> smem ; DMMA DivInt.F ; epilogue (dt, -1/J, SSP-RK update, stores) ;
//            DMMA FluxEdgeInterp.q_new ; Q_Face stores
//
// Tiles are always full: every element array is padded to a multiple of 32 columns and the padding columns hold a
// benign constant state with zero metrics (dfr2d.cu: fill_padding), so no tail predicates are needed.
// Warp w owns conserved variable (w & 3) and the rows / n-tile selected by part = w >> 2; all row <-> warp maps are
// shifts and adds.
#pragma once
#include "dfr2d_elem_mma.cuh"

namespace dfr2d {

#ifndef DFR2D_PIPE_MINBLOCKS
#define DFR2D_PIPE_MINBLOCKS 2
#endif
#ifndef DFR2D_PIPE_LATE_WAIT
#define DFR2D_PIPE_LATE_WAIT 1
#endif
#ifndef DFR2D_PIPE_PACED
#define DFR2D_PIPE_PACED 1
#endif
#ifndef DFR2D_PIPE_WARPS
#define DFR2D_PIPE_WARPS 8
#endif
constexpr int kPipeWarps = DFR2D_PIPE_WARPS;          // 8 or 16
constexpr int kPipeThreads = kPipeWarps * 32;
constexpr int kPipeParts = kPipeWarps / 4;            // warps per conserved variable
constexpr int kPipeNT = 4 / kPipeParts;               // 8-element n-tiles per warp

template <int N> struct PipeDim {
    using MD = MmaDim<N>;
    static constexpr int NI = MD::NI, NF = MD::NF, NF3 = MD::NF3;
    static constexpr int QR = (NI + kPipeParts - 1) / kPipeParts;       // stage-input rows per thread
    static constexpr int ER = (NF3 + kPipeParts - 1) / kPipeParts;      // edge rows per thread
    static constexpr int XROWS = MD::QROWS;
    static size_t smem_bytes(int nExtra) {
        return (size_t)(4 * (MD::QROWS + MD::FROWS) * MD::SE + nExtra * 4 * XROWS * MD::SE + 2 * kElemsPerBlock +
                        MD::kFragDoubles) * sizeof(double);
    }
};

template <typename T> __device__ __forceinline__ T pick3(int i, T a0, T a1, T a2) { return i == 0 ? a0 : (i == 1 ? a1 : a2); }

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc) {
    const unsigned sd = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sd), "l"(gsrc) : "memory");
}

#ifdef DFR2D_PIPE_TIMING
__device__ unsigned long long g_pipe_clk[8];
#define PIPE_TICK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); clkAcc[i] += t_ - clkLast; clkLast = t_; } } while (0)
#else
#define PIPE_TICK(i)
#endif

template <int N>
__global__ void __launch_bounds__(kPipeThreads, DFR2D_PIPE_MINBLOCKS) k_elem_pipe(ElemMmaArgs args) {
#ifdef DFR2D_PIPE_TIMING
    long long clkAcc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, clkLast = clock64();
#endif
    using MD = MmaDim<N>;
    using PD = PipeDim<N>;
    constexpr int NI = MD::NI, NEd = Dim<N>::NpEdge, NF = MD::NF, NF3 = MD::NF3, E = kElemsPerBlock, SE = MD::SE;
    constexpr int M1 = MD::M1, K1 = MD::K1, M2 = MD::M2, K2 = MD::K2, QR = PD::QR, ER = PD::ER;
    constexpr int P = kPipeParts, NT = kPipeNT;
    const ElemArgs &a = args.a;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) {
        if (a.rk == 4 && a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            a.sc->time[a.par ^ 1] = a.sc->time[a.par];
            a.sc->finished = 1;
        }
        return;
    }
    const int nExtra = (a.rk == 0 || a.rhsOut != nullptr) ? 0 : (a.rk == 4 ? 4 : 1);
    extern __shared__ double smem[];
    double *sQ = smem;                                  // [4][QROWS][SE]
    double *sF = sQ + 4 * MD::QROWS * SE;               // [4][FROWS][SE]
    double *sDT = sF + 4 * MD::FROWS * SE;              // [E]
    double *sMOOJ = sDT + E;                            // [E]
    double *sA = sMOOJ + E;                             // operator fragments [frag][32]
    double *sX = sA + MD::kFragDoubles;                 // [nExtra][4][XROWS][SE]  q0 (, q2, q3, R)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int var = w & 3, part = w >> 2;
    const size_t Kp = a.Kp;

    for (int t = threadIdx.x; t < MD::kFragDoubles; t += kPipeThreads) sA[t] = args.frags[t];
    const double *sA1 = sA + lane, *sA2 = sA + (size_t)M1 * K1 * 32 + lane;
    double *myQ = sQ + (size_t)var * MD::QROWS * SE;
    double *myF = sF + (size_t)var * MD::FROWS * SE;
    if (part == 0) {
        for (int r = NI; r < MD::QROWS; r++) myQ[r * SE + lane] = 0.0;
        for (int r = NF; r < MD::FROWS; r++) myF[r * SE + lane] = 0.0;
    }
    double dtGlobal = 0.0;
    if (!a.ph.localDT) {
        const double gw = __longlong_as_double((long long)a.sc->wave[a.slot][0]);
        dtGlobal = a.ph.CFL / gw;
        const double t = a.sc->time[a.par];
        if (t + dtGlobal > a.ph.FinalTime) dtGlobal = a.ph.FinalTime - t;
    }
    bool bad = false;
    const int fr = lane >> 2, fc = lane & 3;
    double *dst = (a.rk == 0) ? a.q1 : (a.rk == 1) ? a.q2 : (a.rk == 2) ? a.q3 : (a.rk == 3) ? a.q4 : a.q0;

    // loop-invariant addresses (the tile's element offset is added per tile)
    const double *qsRow = a.qs + ((size_t)var * NI + part) * Kp + lane;          // + c*P*Kp
    const double *efRow = a.eflux + (size_t)var * NEd * a.NEp;
    const int cpHalf = lane >> 4, cpCh = lane & 15;     // cp.async role: row parity, 16-byte chunk of the row

    // ---- prefetch registers ---------------------------------------------------------------------------
    double pq[QR], pe[ER], pIin0 = 0, pIin1 = 0, pIin2 = 0, pJd = 1, pJ0 = 0, pJ1 = 0, pJ2 = 0, pJ3 = 0;
    double pAgg0 = 0, pAgg1 = 0, pAgg2 = 0, pDT = 0;
    int sN0 = 0, sN1 = 0, sN2 = 0, sC0 = 0, sC1 = 0, sC2 = 0;      // edge slots of tile t+2 / of the tile in flight
    int oC0 = 0, oC1 = 0, oC2 = 0;                                   // edge slots of the tile being consumed
#pragma unroll
    for (int c = 0; c < QR; c++) pq[c] = 0.0;
#pragma unroll
    for (int c = 0; c < ER; c++) pe[c] = 0.0;

#define PIPE_LOAD_SLOTS(TILE, S0, S1, S2)                                   \
    if ((TILE) < args.nTiles) {                                             \
        const size_t kk_ = (size_t)(TILE) * E + lane;                       \
        S0 = a.etoe[kk_]; S1 = a.etoe[Kp + kk_]; S2 = a.etoe[2 * Kp + kk_]; \
    }
#define PIPE_ISSUE_LOADS(TILE)                                                                          \
    if ((TILE) < args.nTiles) {                                                                         \
        const size_t kk_ = (size_t)(TILE) * E;                                                          \
        _Pragma("unroll") for (int c = 0; c < QR; c++)                                                  \
            if (part + P * c < NI) pq[c] = qsRow[(size_t)c * P * Kp + kk_];                             \
        _Pragma("unroll") for (int c = 0; c < ER; c++) {                                                \
            const int row = part + P * c;                                                               \
            if (row < NF3) {                                                                            \
                const int le = (row >= NEd) + (row >= 2 * NEd), i = row - le * NEd;                     \
                const int sl = pick3(le, sC0, sC1, sC2);                                                \
                const bool own = sl >= 0;                                                               \
                pe[c] = efRow[(size_t)(own ? i : NEd - 1 - i) * a.NEp + (own ? sl : -1 - sl)];          \
            }                                                                                           \
        }                                                                                               \
        pIin0 = a.IInII[kk_ + lane]; pIin1 = a.IInII[Kp + kk_ + lane]; pIin2 = a.IInII[2 * Kp + kk_ + lane]; \
        pJd = a.Jdet[kk_ + lane];                                                                       \
        pJ0 = a.Jinv[kk_ + lane]; pJ1 = a.Jinv[Kp + kk_ + lane];                                        \
        pJ2 = a.Jinv[2 * Kp + kk_ + lane]; pJ3 = a.Jinv[3 * Kp + kk_ + lane];                           \
        if (a.ph.localDT && w == 0) {                                                                   \
            pAgg0 = a.agg[sC0 >= 0 ? sC0 : -1 - sC0]; pAgg1 = a.agg[sC1 >= 0 ? sC1 : -1 - sC1];         \
            pAgg2 = a.agg[sC2 >= 0 ? sC2 : -1 - sC2]; pDT = a.DT[kk_ + lane];                           \
        }                                                                                               \
    }

    // the same loads one at a time, so that they can be paced through the compute phases of the current tile (a warp
    // that issues ~25 loads back to back sits in the LSU queue for ~2000 cycles while HBM back-pressures)
#define PIPE_LQ(c)  do { if (haveN && (c) < QR && part + P * (c) < NI) pq[(c)] = qsRow[(size_t)(c) * P * Kp + kkN]; } while (0)
#define PIPE_LE(c)  do { if (haveN && (c) < ER) { const int row = part + P * (c);                                      \
        if (row < NF3) { const int le = (row >= NEd) + (row >= 2 * NEd), i = row - le * NEd;                          \
            const int sl = pick3(le, sC0, sC1, sC2); const bool own = sl >= 0;                                         \
            pe[(c)] = efRow[(size_t)(own ? i : NEd - 1 - i) * a.NEp + (own ? sl : -1 - sl)]; } } } while (0)
#define PIPE_LG()   do { if (haveN) {                                                                                  \
        pIin0 = a.IInII[kkN + lane]; pIin1 = a.IInII[Kp + kkN + lane]; pIin2 = a.IInII[2 * Kp + kkN + lane];           \
        pJd = a.Jdet[kkN + lane]; pJ0 = a.Jinv[kkN + lane]; pJ1 = a.Jinv[Kp + kkN + lane];                             \
        pJ2 = a.Jinv[2 * Kp + kkN + lane]; pJ3 = a.Jinv[3 * Kp + kkN + lane];                                          \
        if (a.ph.localDT && w == 0) {                                                                                  \
            pAgg0 = a.agg[sC0 >= 0 ? sC0 : -1 - sC0]; pAgg1 = a.agg[sC1 >= 0 ? sC1 : -1 - sC1];                        \
            pAgg2 = a.agg[sC2 >= 0 ? sC2 : -1 - sC2]; pDT = a.DT[kkN + lane]; } } } while (0)
#define PIPE_FENCE() asm volatile("" ::: "memory")

    PIPE_LOAD_SLOTS(blockIdx.x, sC0, sC1, sC2);
    PIPE_ISSUE_LOADS(blockIdx.x);
    PIPE_LOAD_SLOTS(blockIdx.x + gridDim.x, sN0, sN1, sN2);

    for (int tile = blockIdx.x; tile < args.nTiles; tile += gridDim.x) {
        const size_t k0 = (size_t)tile * E;
        __syncthreads();                       // previous tile fully consumed
        PIPE_TICK(0);

        // ---- RK registers of THIS tile -> smem by cp.async (consumed in the epilogue) ------------------------
        if (nExtra > 0) {
#pragma unroll
            for (int cc = 0; cc < (QR + 1) / 2; cc++) {
                const int i = part + P * (2 * cc + cpHalf);
                if (i < NI) {
                    const size_t g = ((size_t)var * NI + i) * Kp + k0 + 2 * cpCh;
                    double *d = &sX[((size_t)var * PD::XROWS + i) * SE + 2 * cpCh];
                    cp_async16(d, a.q0 + g);
                    if (nExtra == 4) {
                        cp_async16(d + 1 * 4 * PD::XROWS * SE, a.q2 + g);
                        cp_async16(d + 2 * 4 * PD::XROWS * SE, a.q3 + g);
                        cp_async16(d + 3 * 4 * PD::XROWS * SE, a.R + g);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        PIPE_TICK(1);
        // ---- registers(t) -> shared memory ------------------------------------------------------------------
        oC0 = sC0; oC1 = sC1; oC2 = sC2;
#pragma unroll
        for (int c = 0; c < QR; c++)
            if (part + P * c < NI) myQ[(part + P * c) * SE + lane] = pq[c];
#pragma unroll
        for (int c = 0; c < ER; c++) {
            const int row = part + P * c;
            if (row < NF3) {
                const int le = (row >= NEd) + (row >= 2 * NEd);
                const bool own = pick3(le, oC0, oC1, oC2) >= 0;
                const double iin = pick3(le, pIin0, pIin1, pIin2);
                // SetRTFluxOnEdges (edges.go:469-479): +flux*IInII for the owner, -flux(reversed)*IInII otherwise
                myF[(2 * NI + row) * SE + lane] = own ? pe[c] * iin : -pe[c] * iin;
            }
        }
        const double jdet = pJd, j0 = pJ0, j1 = pJ1, j2 = pJ2, j3 = pJ3;
        if (w == 0) {
            double dtk = dtGlobal;
            if (a.ph.localDT) {
                const double wmaxk = fmax(fmax(pAgg0, pAgg1), pAgg2);
                const double d = (a.rk == 0) ? -100.0 : pDT;
                dtk = a.ph.CFL / fmax(d, wmaxk);
                if (a.rhsOut == nullptr) a.DT[k0 + lane] = dtk;
            }
            sDT[lane] = dtk;
            sMOOJ[lane] = -(1.0 / jdet);
        }
        PIPE_TICK(2);
        // ---- put tile t+1 in flight, and the edge slots of tile t+2 -------------------------------------------
        sC0 = sN0; sC1 = sN1; sC2 = sN2;
        const bool haveN = tile + gridDim.x < args.nTiles;
        const size_t kkN = (size_t)(tile + gridDim.x) * E;
#if DFR2D_PIPE_PACED
        PIPE_LG();
        PIPE_LOAD_SLOTS(tile + 2 * gridDim.x, sN0, sN1, sN2);
#else
        PIPE_ISSUE_LOADS(tile + gridDim.x);
        PIPE_LOAD_SLOTS(tile + 2 * gridDim.x, sN0, sN1, sN2);
#endif
        PIPE_TICK(3);
        __syncthreads();
        PIPE_TICK(4);

        // ---- SetRTFluxInternal: point j handled by warp j mod kPipeWarps ------------------------------------------
#pragma unroll
        for (int jj = 0; jj < (NI + kPipeWarps - 1) / kPipeWarps; jj++) {
            const int j = w + kPipeWarps * jj;
            if (j < NI) {
                double Q[4], Fx[4], Fy[4];
#pragma unroll
                for (int m = 0; m < 4; m++) Q[m] = sQ[(m * MD::QROWS + j) * SE + lane];
                flux_calc(a.ph.gamma, Q, Fx, Fy);
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    sF[(m * MD::FROWS + j) * SE + lane] = jdet * (j0 * Fx[m] + j1 * Fy[m]);
                    sF[(m * MD::FROWS + j + NI) * SE + lane] = jdet * (j2 * Fx[m] + j3 * Fy[m]);
                }
            }
#if DFR2D_PIPE_PACED
            PIPE_LQ(2 * jj); PIPE_LQ(2 * jj + 1); PIPE_FENCE();
#endif
        }
#if DFR2D_PIPE_PACED
#pragma unroll
        for (int c = 2 * ((NI + kPipeWarps - 1) / kPipeWarps); c < QR; c++) PIPE_LQ(c);
        PIPE_FENCE();
#endif
#if !DFR2D_PIPE_LATE_WAIT
        asm volatile("cp.async.wait_all;" ::: "memory");
#endif
        __syncthreads();
        PIPE_TICK(5);

        // ---- C1 = DivInt . F (tensor cores): variable `var`, n-tiles NT*part .. -----------------------------------
        double c1[M1][NT][2];
#pragma unroll
        for (int mt = 0; mt < M1; mt++)
#pragma unroll
            for (int t = 0; t < NT; t++) c1[mt][t][0] = c1[mt][t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < K1; ks++) {
            double b[NT], a1[M1];
#pragma unroll
            for (int t = 0; t < NT; t++) b[t] = myF[(4 * ks + fc) * SE + 8 * (NT * part + t) + fr];
#pragma unroll
            for (int mt = 0; mt < M1; mt++) a1[mt] = sA1[(mt * K1 + ks) * 32];
#pragma unroll
            for (int mt = 0; mt < M1; mt++)
#pragma unroll
                for (int t = 0; t < NT; t++) dmma884(c1[mt][t][0], c1[mt][t][1], a1[mt], b[t]);
#if DFR2D_PIPE_PACED
            PIPE_LE(ks); PIPE_FENCE();
#endif
        }
#if DFR2D_PIPE_PACED
#pragma unroll
        for (int c = K1; c < ER; c++) PIPE_LE(c);
#endif
#if DFR2D_PIPE_LATE_WAIT
        // the RK registers are only needed from here on: give their cp.async the whole DMMA loop to land
        if (nExtra > 0) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        }
#endif
        // epilogue in fragment layout: lane holds rows i = 8 mt + fr, elements e0 = 8 (NT part + t) + 2 fc, e0 + 1
#pragma unroll
        for (int mt = 0; mt < M1; mt++) {
            const int i = 8 * mt + fr;
            if (i < NI) {
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int e0 = 8 * (NT * part + t) + 2 * fc;
                    const double2 mo = *reinterpret_cast<const double2 *>(&sMOOJ[e0]);
                    const double rhs0 = c1[mt][t][0] * mo.x, rhs1 = c1[mt][t][1] * mo.y;
                    const size_t o = ((size_t)var * NI + i) * Kp + k0 + e0;
                    if (a.rhsOut != nullptr) {
                        *reinterpret_cast<double2 *>(a.rhsOut + o) = make_double2(rhs0, rhs1);
                        continue;
                    }
                    const double2 dt = *reinterpret_cast<const double2 *>(&sDT[e0]);
                    const double2 qs = *reinterpret_cast<const double2 *>(&myQ[i * SE + e0]);
                    const double *xq = &sX[((size_t)var * PD::XROWS + i) * SE + e0];
                    double2 qn;
                    if (a.rk == 0) {
                        qn.x = qs.x + RK0_A * (dt.x * rhs0);
                        qn.y = qs.y + RK0_A * (dt.y * rhs1);
                    } else if (a.rk < 4) {
                        const double2 q0v = *reinterpret_cast<const double2 *>(xq);
                        const double ca = (a.rk == 1) ? RK1_A : (a.rk == 2) ? RK2_A : RK3_A;
                        const double cb = (a.rk == 1) ? RK1_B : (a.rk == 2) ? RK2_B : RK3_B;
                        const double cc = (a.rk == 1) ? RK1_C : (a.rk == 2) ? RK2_C : RK3_C;
                        qn.x = ca * q0v.x + cb * qs.x + cc * (dt.x * rhs0);
                        qn.y = ca * q0v.y + cb * qs.y + cc * (dt.y * rhs1);
                        if (a.rk == 3) *reinterpret_cast<double2 *>(a.R + o) = make_double2(rhs0, rhs1);
                    } else {
                        const double2 q0v = *reinterpret_cast<const double2 *>(xq);
                        const double2 q2v = *reinterpret_cast<const double2 *>(xq + 1 * 4 * PD::XROWS * SE);
                        const double2 q3v = *reinterpret_cast<const double2 *>(xq + 2 * 4 * PD::XROWS * SE);
                        const double2 rv = *reinterpret_cast<const double2 *>(xq + 3 * 4 * PD::XROWS * SE);
                        double2 r;
                        r.x = -q0v.x + RK4_A * q2v.x + RK4_B * q3v.x + RK4_C * qs.x + RK4_D * (dt.x * rv.x) + RK4_E * (dt.x * rhs0);
                        r.y = -q0v.y + RK4_A * q2v.y + RK4_B * q3v.y + RK4_C * qs.y + RK4_D * (dt.y * rv.y) + RK4_E * (dt.y * rhs1);
                        qn.x = q0v.x + r.x;
                        qn.y = q0v.y + r.y;
                        *reinterpret_cast<double2 *>(a.R + o) = r;
                    }
                    bad |= (qn.x != qn.x) || (qn.y != qn.y);
                    *reinterpret_cast<double2 *>(dst + o) = qn;
                    *reinterpret_cast<double2 *>(&myQ[i * SE + e0]) = qn;
                }
            }
        }
        PIPE_TICK(6);
        if (a.rhsOut != nullptr || a.qface == nullptr) continue;
        __syncwarp();          // phase 4 only reads this warp's own columns of the fresh register

        // ---- C2 = FluxEdgeInterp . q_new (tensor cores): next stage's Q_Face ------------------------------------
        double c2[M2][NT][2];
#pragma unroll
        for (int mt = 0; mt < M2; mt++)
#pragma unroll
            for (int t = 0; t < NT; t++) c2[mt][t][0] = c2[mt][t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < K2; ks++) {
            double b[NT], a2[M2];
#pragma unroll
            for (int t = 0; t < NT; t++) b[t] = myQ[(4 * ks + fc) * SE + 8 * (NT * part + t) + fr];
#pragma unroll
            for (int mt = 0; mt < M2; mt++) a2[mt] = sA2[(mt * K2 + ks) * 32];
#pragma unroll
            for (int mt = 0; mt < M2; mt++)
#pragma unroll
                for (int t = 0; t < NT; t++) dmma884(c2[mt][t][0], c2[mt][t][1], a2[mt], b[t]);
        }
#pragma unroll
        for (int mt = 0; mt < M2; mt++) {
            const int m = 8 * mt + fr;
            if (m < NF3) {
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    const int e0 = 8 * (NT * part + t) + 2 * fc;
                    *reinterpret_cast<double2 *>(a.qface + ((size_t)var * NF3 + m) * Kp + k0 + e0) =
                        make_double2(c2[mt][t][0], c2[mt][t][1]);
                }
            }
        }
        PIPE_TICK(7);
    }
#ifdef DFR2D_PIPE_TIMING
    if (threadIdx.x == 0) for (int i = 0; i < 8; i++) atomicAdd(&g_pipe_clk[i], (unsigned long long)clkAcc[i]);
#endif
#undef PIPE_LQ
#undef PIPE_LE
#undef PIPE_LG
#undef PIPE_FENCE
#undef PIPE_LOAD_SLOTS
#undef PIPE_ISSUE_LOADS
    if (bad) a.sc->nanFlag = 1;

    if (a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        a.sc->wave[a.slot ^ 1][0] = 0ull;
        a.sc->wave[a.slot ^ 1][1] = 0ull;
        if (!a.ph.localDT) a.sc->globalDT = dtGlobal;
        if (a.rk == 4) {
            const double tnew = a.sc->time[a.par] + (a.ph.localDT ? a.sc->globalDT : dtGlobal);
            a.sc->time[a.par ^ 1] = tnew;
            a.sc->timeOut = tnew;
            const long long st = a.sc->steps + 1;
            a.sc->steps = st;
            if (tnew >= a.ph.FinalTime || st >= (long long)a.ph.maxIter) a.sc->finished = 1;
        }
    }
}

}  // namespace dfr2d
