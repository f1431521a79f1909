// k_elem_mma_diss<N>: element kernel of the PerssonC0 path with its contractions on the FP64 tensor cores
// (DFR2D_DISS_ELEM_KERNEL=3; opt-in until measured -- the default stays k_elem<N,true>).
//
// k_elem<N,true> is the top kernel of the dissipation stage (39 % at 2M triangles, N=4, profiles/r01q_*): per element and
// variable it runs DivInt[NpInt x NpFlux] over (F - F_visc) and the modal limiter Vinv / V (2 NpInt^2) as constant-operand
// DFMA, the form that tops out at 33-43 % of the FP64 pipe on sm_100a.  This kernel is k_elem_mma
// (persistent CTAs, warp = conserved variable, operator fragments resident in registers, stride-36 smem rows) with
//   * SetRTFluxOnEdges minus the viscous edge DOFs and SetRTFluxInternal minus the interior dissipation DOFs
//     (AddDissipation, dissipation.go:274-346, folded in with the opposite sign: RHS = -(1/J) DivInt (F - F_visc)),
//   * LimitFilterSolution(RHSQ) (euler.go:499; limitAndFilterSolution dissipation.go:606-622) as two more DMMA products
//     through the warp's own (dead) F rows:  RHS -> smem -> Vinv . RHS -> scale modes i >= 1 by mf_i (1 - sin(pi sigma/2))
//     in the accumulator layout -> smem -> V . (...),
//   * the viscous time-step limit (euler.go:956-966, :989-999).
// The next stage's edge interpolation is not fused (k_diss_prepare needs the vertex-merged sigma first).
#pragma once
#include "dfr2d_elem_mma.cuh"
#include "dfr2d_diss_kernels.cuh"

namespace dfr2d {

template <int N> struct MmaDissDim {
    static constexpr int NI = Dim<N>::NpInt, NF = Dim<N>::NpFlux;
    static constexpr int M1 = (NI + 7) / 8, K1 = (NF + 3) / 4;     // DivInt tiles
    static constexpr int M3 = (NI + 7) / 8, K3 = (NI + 3) / 4;     // Vinv / V tiles
    static constexpr int SE = kElemsPerBlock + 4;
    static constexpr int QROWS = 4 * K3, FROWS = 4 * K1;
    static constexpr int kFragDoubles = (M1 * K1 + 2 * M3 * K3) * 32;
    static constexpr size_t kSmemBytes = (size_t)(4 * (QROWS + FROWS) * SE + 3 * kElemsPerBlock) * sizeof(double);
    static_assert(8 * M3 >= QROWS && FROWS >= QROWS, "the limiter reuses the first QROWS rows of the F block");
};

template <int N>
__global__ void __launch_bounds__(kElemThreads, DFR2D_MMA_MINBLOCKS) k_elem_mma_diss(ElemMmaArgs args) {
    using MD = MmaDissDim<N>;
    constexpr int NI = MD::NI, NEd = Dim<N>::NpEdge, NF = MD::NF, E = kElemsPerBlock, SE = MD::SE;
    constexpr int M1 = MD::M1, K1 = MD::K1, M3 = MD::M3, K3 = MD::K3;
    const ElemArgs &a = args.a;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) {
        if (a.rk == 4 && a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            a.sc->time[a.par ^ 1] = a.sc->time[a.par];
            a.sc->finished = 1;
        }
        return;
    }
    extern __shared__ double smem[];
    double *sQ = smem;                                 // [4][QROWS][SE]
    double *sF = smem + 4 * MD::QROWS * SE;            // [4][FROWS][SE]
    double *sDT = sF + 4 * MD::FROWS * SE;             // [E]
    double *sMOOJ = sDT + E;                           // [E]  -1/Jdet
    double *sOMA = sMOOJ + E;                          // [E]  1 - sin(pi sigma / 2)
    const int lane = threadIdx.x & 31, n = threadIdx.x >> 5;
    const size_t Kp = a.Kp;

    // operator fragments -> registers, once per CTA
    double A1[M1][K1], AVi[M3][K3], AV[M3][K3];
#pragma unroll
    for (int mt = 0; mt < M1; mt++)
#pragma unroll
        for (int ks = 0; ks < K1; ks++) A1[mt][ks] = args.frags[(size_t)(mt * K1 + ks) * 32 + lane];
#pragma unroll
    for (int mt = 0; mt < M3; mt++)
#pragma unroll
        for (int ks = 0; ks < K3; ks++) {
            AVi[mt][ks] = args.frags[(size_t)(M1 * K1 + mt * K3 + ks) * 32 + lane];
            AV[mt][ks] = args.frags[(size_t)(M1 * K1 + M3 * K3 + mt * K3 + ks) * 32 + lane];
        }
    // zero the padding rows once (never written afterwards; the limiter only rewrites rows < QROWS of the F block with
    // values whose padding rows are exact zeros of the zero-padded operators)
    for (int r = NI; r < MD::QROWS; r++) sQ[(n * MD::QROWS + r) * SE + lane] = 0.0;
    for (int r = NF; r < MD::FROWS; r++) sF[(n * MD::FROWS + r) * SE + lane] = 0.0;

    double *myQ = sQ + (size_t)n * MD::QROWS * SE;
    double *myF = sF + (size_t)n * MD::FROWS * SE;
    // calculateGlobalDT (euler.go:945-971) incl. the viscous limit; the same for every tile
    double dtGlobal = 0.0;
    if (!a.ph.localDT) {
        const double gw = __longlong_as_double((long long)a.sc->wave[a.slot][0]);
        dtGlobal = a.ph.CFL / gw;
        const double gv = __longlong_as_double((long long)a.sc->wave[a.slot][1]);
        dtGlobal = fmin(dtGlobal, a.ph.Cdiff / gv);
        const double t = a.sc->time[a.par];
        if (t + dtGlobal > a.ph.FinalTime) dtGlobal = a.ph.FinalTime - t;
    }
    bool bad = false;
    const int fr = lane >> 2, fc = lane & 3;     // fragment row / column of this lane
    // mode filter of this lane's fragment rows (ModeFilter, dfr_shock_capturing.go:43-69); mode 0 is never scaled
    double mfRow[M3];
#pragma unroll
    for (int mt = 0; mt < M3; mt++) {
        const int i = 8 * mt + fr;
        mfRow[mt] = (i >= 1 && i < NI) ? ops<N>().mf[i] : 0.0;
    }

    for (int tile = blockIdx.x; tile < args.nTiles; tile += gridDim.x) {
        const int k0 = tile * E;
        const int k = k0 + lane;
        const bool valid = k < a.K;
        const int kc = valid ? k : a.K - 1;
        const bool fullTile = k0 + E <= a.K;
        __syncthreads();       // previous tile fully consumed before sQ / sF are overwritten

        if (a.pfTiles > 0) {
            // bulk L2 prefetch of this CTA's NEXT tile (the kernel is load -> barrier -> compute -> store per tile, ncu:
            // 26 % DRAM throughput, long-scoreboard bound): stage input, interior DissX / DissY, the extra RK registers,
            // geometry, and the neighbourhood of the edge-flux / viscous-flux slots
            const long long kt = (long long)(tile + gridDim.x) * E;
            if (kt + E <= a.K) {
                constexpr unsigned RB = E * sizeof(double);
                for (int r = threadIdx.x; r < 4 * NI; r += kElemThreads) {
                    prefetch_l2(a.qs + (size_t)r * Kp + kt, RB);
                    prefetch_l2(a.dissX + (size_t)r * Kp + kt, RB);
                    prefetch_l2(a.dissY + (size_t)r * Kp + kt, RB);
                    if (a.rk >= 1) prefetch_l2(a.q0 + (size_t)r * Kp + kt, RB);
                    if (a.rk == 4) {
                        prefetch_l2(a.q2 + (size_t)r * Kp + kt, RB);
                        prefetch_l2(a.q3 + (size_t)r * Kp + kt, RB);
                        prefetch_l2(a.R + (size_t)r * Kp + kt, RB);
                    }
                }
                if (threadIdx.x < 4) prefetch_l2(a.Jinv + (size_t)threadIdx.x * Kp + kt, RB);
                else if (threadIdx.x < 7) prefetch_l2(a.IInII + (size_t)(threadIdx.x - 4) * Kp + kt, RB);
                else if (threadIdx.x < 10) prefetch_l2(a.etoe + (size_t)(threadIdx.x - 7) * Kp + kt, E * sizeof(int));
                else if (threadIdx.x == 10) prefetch_l2(a.Jdet + kt, RB);
                else if (threadIdx.x == 11) prefetch_l2(a.sigma + kt, RB);
                else if (threadIdx.x >= 32 && threadIdx.x < 32 + 4 * NEd) {
                    const long long s0 = ((kt * 3 / 2) / 16) * 16;       // owner slots follow the element numbering (1.5 per element)
                    if (s0 + 64 <= a.NEp) {
                        prefetch_l2(a.eflux + (size_t)(threadIdx.x - 32) * a.NEp + s0, 64 * sizeof(double));
                    }
                }
            }
        }

        // ---- phase 1: stage input row, edge DOFs of (F - F_visc), dt, -1/J, limiter factor ------------------------
#pragma unroll
        for (int i = 0; i < NI; i++) myQ[i * SE + lane] = a.qs[((size_t)n * NI + i) * Kp + kc];
        {
            double wmaxk = -1.7976931348623157e308, vmaxk = -1.7976931348623157e308;
#pragma unroll
            for (int le = 0; le < 3; le++) {
                const int s = a.etoe[(size_t)le * Kp + kc];
                const bool owner = s >= 0;
                const int slot = owner ? s : -1 - s;
                const double iin = a.IInII[(size_t)le * Kp + kc];
                const double *f = a.eflux + ((size_t)n * NEd) * a.NEp + slot;
#pragma unroll
                for (int i = 0; i < NEd; i++) {
                    const size_t o = (size_t)(owner ? i : NEd - 1 - i) * a.NEp;
                    // SetRTFluxOnEdges (edges.go:454-483) on F - F_visc: k_visc_edge subtracted the viscous normal flux from
                    // eflux in place (AddDissipation's edge DOFs, dissipation.go:316-333)
                    const double v = f[o];
                    myF[(2 * NI + le * NEd + i) * SE + lane] = owner ? v * iin : -v * iin;
                }
                if (a.ph.localDT && n == 0) {
                    wmaxk = fmax(wmaxk, a.agg[slot]);
                    vmaxk = fmax(vmaxk, a.aggv[slot]);
                }
            }
            if (n == 0) {
                double dtk = dtGlobal;
                if (a.ph.localDT) {
                    // InitializeDT at stage 0, DT = max(DT, aggregates), CalculateLocalDT (euler.go:637-643, :973-1002)
                    const double d = (a.rk == 0) ? -100.0 : a.DT[kc];
                    dtk = a.ph.CFL / fmax(d, wmaxk);
                    double dtv = fmax(a.DTVisc[kc], vmaxk);
                    if (dtv > 1.e-9) { dtv = a.ph.Cdiff / dtv; dtk = fmin(dtk, dtv); }
                    if (valid && a.rhsOut == nullptr) { a.DT[k] = dtk; a.DTVisc[k] = dtv; }
                }
                sDT[lane] = dtk;
                sMOOJ[lane] = -(1.0 / a.Jdet[kc]);
                sOMA[lane] = 1.0 - sin(0.5 * 3.14159265358979323846 * a.sigma[kc]);
            }
        }
        __syncthreads();

        // ---- phase 2: SetRTFluxInternal minus the interior dissipation DOFs, point j handled by warp j mod 4 ------
        {
            const double jdet = a.Jdet[kc];
            const double j0 = a.Jinv[0 * Kp + kc], j1 = a.Jinv[1 * Kp + kc], j2 = a.Jinv[2 * Kp + kc], j3 = a.Jinv[3 * Kp + kc];
#pragma unroll
            for (int jj = 0; jj < (NI + 3) / 4; jj++) {
                const int j = n + 4 * jj;
                if (j < NI) {
                    double Q[4], Fx[4], Fy[4];
#pragma unroll
                    for (int m = 0; m < 4; m++) Q[m] = sQ[(m * MD::QROWS + j) * SE + lane];
                    flux_calc(a.ph.gamma, Q, Fx, Fy);
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        double frr = jdet * (j0 * Fx[m] + j1 * Fy[m]);
                        double fss = jdet * (j2 * Fx[m] + j3 * Fy[m]);
                        const double dix = a.dissX[((size_t)m * NI + j) * Kp + kc];
                        const double diy = a.dissY[((size_t)m * NI + j) * Kp + kc];
                        frr -= jdet * (j0 * dix + j1 * diy);                 // dissipation.go:310-315
                        fss -= jdet * (j2 * dix + j3 * diy);
                        sF[(m * MD::FROWS + j) * SE + lane] = frr;
                        sF[(m * MD::FROWS + j + NI) * SE + lane] = fss;
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3: C1 = DivInt . (F - F_visc) on the tensor cores -------------------------------------------------
        double c1[M1][4][2];
#pragma unroll
        for (int mt = 0; mt < M1; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) c1[mt][nt][0] = c1[mt][nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < K1; ks++) {
            double b[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = myF[(4 * ks + fc) * SE + 8 * nt + fr];
#pragma unroll
            for (int mt = 0; mt < M1; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(c1[mt][nt][0], c1[mt][nt][1], A1[mt][ks], b[nt]);
        }

        // ---- LimitFilterSolution(RHSQ): RHS = -(1/J) C1 -> Vinv -> mode scaling -> V, through rows [0, QROWS) of this
        //      warp's F block (dead after the loop above; only this warp ever touches them) --------------------------
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < M1; mt++) {
            const int i = 8 * mt + fr;
            if (i < MD::QROWS) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const int e0 = 8 * nt + 2 * fc;
                    const double2 mo = *reinterpret_cast<const double2 *>(&sMOOJ[e0]);
                    *reinterpret_cast<double2 *>(&myF[i * SE + e0]) = make_double2(c1[mt][nt][0] * mo.x, c1[mt][nt][1] * mo.y);
                }
            }
        }
        __syncwarp();
        double c3[M3][4][2];
#pragma unroll
        for (int mt = 0; mt < M3; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) c3[mt][nt][0] = c3[mt][nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < K3; ks++) {
            double b[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = myF[(4 * ks + fc) * SE + 8 * nt + fr];
#pragma unroll
            for (int mt = 0; mt < M3; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(c3[mt][nt][0], c3[mt][nt][1], AVi[mt][ks], b[nt]);
        }
        __syncwarp();          // every lane has read the RHS rows
#pragma unroll
        for (int mt = 0; mt < M3; mt++) {
            const int i = 8 * mt + fr;
            if (i < MD::QROWS) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const int e0 = 8 * nt + 2 * fc;
                    double2 uh = make_double2(c3[mt][nt][0], c3[mt][nt][1]);
                    if (i >= 1) {      // modes i >= 1: mf_i (1 - alpha_k)
                        const double2 oma = *reinterpret_cast<const double2 *>(&sOMA[e0]);
                        uh.x *= mfRow[mt] * oma.x;
                        uh.y *= mfRow[mt] * oma.y;
                    }
                    *reinterpret_cast<double2 *>(&myF[i * SE + e0]) = uh;
                }
            }
        }
        __syncwarp();
        double c4[M3][4][2];
#pragma unroll
        for (int mt = 0; mt < M3; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) c4[mt][nt][0] = c4[mt][nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < K3; ks++) {
            double b[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = myF[(4 * ks + fc) * SE + 8 * nt + fr];
#pragma unroll
            for (int mt = 0; mt < M3; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(c4[mt][nt][0], c4[mt][nt][1], AV[mt][ks], b[nt]);
        }

        // ---- epilogue: lane holds rows i = 8 mt + fr, elements e0 = 8 nt + 2 fc (+1); SSP54 (euler.go:502-565) --------
        double *dst = (a.rk == 0) ? a.q1 : (a.rk == 1) ? a.q2 : (a.rk == 2) ? a.q3 : (a.rk == 3) ? a.q4 : a.q0;
#pragma unroll
        for (int mt = 0; mt < M3; mt++) {
            const int i = 8 * mt + fr;
            if (i < NI) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const int e0 = 8 * nt + 2 * fc;
                    const double rhs0 = c4[mt][nt][0], rhs1 = c4[mt][nt][1];
                    const size_t o = ((size_t)n * NI + i) * Kp + k0 + e0;
                    const bool v0 = k0 + e0 < a.K, v1 = k0 + e0 + 1 < a.K;
                    if (a.rhsOut != nullptr) {
                        if (v0) a.rhsOut[o] = rhs0;
                        if (v1) a.rhsOut[o + 1] = rhs1;
                        continue;
                    }
                    const double2 dt = *reinterpret_cast<const double2 *>(&sDT[e0]);
                    const double2 qs = *reinterpret_cast<const double2 *>(&myQ[i * SE + e0]);
                    double2 q0v = make_double2(0, 0), q2v = q0v, q3v = q0v, rv = q0v;
                    if (fullTile) {
                        if (a.rk >= 1) q0v = *reinterpret_cast<const double2 *>(a.q0 + o);
                        if (a.rk == 4) {
                            q2v = *reinterpret_cast<const double2 *>(a.q2 + o);
                            q3v = *reinterpret_cast<const double2 *>(a.q3 + o);
                            rv = *reinterpret_cast<const double2 *>(a.R + o);
                        }
                    } else {
                        if (a.rk >= 1) { if (v0) q0v.x = a.q0[o]; if (v1) q0v.y = a.q0[o + 1]; }
                        if (a.rk == 4) {
                            if (v0) { q2v.x = a.q2[o]; q3v.x = a.q3[o]; rv.x = a.R[o]; }
                            if (v1) { q2v.y = a.q2[o + 1]; q3v.y = a.q3[o + 1]; rv.y = a.R[o + 1]; }
                        }
                    }
                    double2 qn, rout = make_double2(rhs0, rhs1);
                    switch (a.rk) {
                        case 0:
                            qn.x = qs.x + RK0_A * (dt.x * rhs0);
                            qn.y = qs.y + RK0_A * (dt.y * rhs1);
                            break;
                        case 1:
                            qn.x = RK1_A * q0v.x + RK1_B * qs.x + RK1_C * (dt.x * rhs0);
                            qn.y = RK1_A * q0v.y + RK1_B * qs.y + RK1_C * (dt.y * rhs1);
                            break;
                        case 2:
                            qn.x = RK2_A * q0v.x + RK2_B * qs.x + RK2_C * (dt.x * rhs0);
                            qn.y = RK2_A * q0v.y + RK2_B * qs.y + RK2_C * (dt.y * rhs1);
                            break;
                        case 3:
                            qn.x = RK3_A * q0v.x + RK3_B * qs.x + RK3_C * (dt.x * rhs0);
                            qn.y = RK3_A * q0v.y + RK3_B * qs.y + RK3_C * (dt.y * rhs1);
                            break;
                        default: {
                            rout.x = -q0v.x + RK4_A * q2v.x + RK4_B * q3v.x + RK4_C * qs.x + RK4_D * (dt.x * rv.x) + RK4_E * (dt.x * rhs0);
                            rout.y = -q0v.y + RK4_A * q2v.y + RK4_B * q3v.y + RK4_C * qs.y + RK4_D * (dt.y * rv.y) + RK4_E * (dt.y * rhs1);
                            qn.x = q0v.x + rout.x;
                            qn.y = q0v.y + rout.y;
                        } break;
                    }
                    bad |= (v0 && qn.x != qn.x) || (v1 && qn.y != qn.y);
                    if (fullTile) {
                        *reinterpret_cast<double2 *>(dst + o) = qn;
                        if (a.rk >= 3) *reinterpret_cast<double2 *>(a.R + o) = rout;
                    } else {
                        if (v0) { dst[o] = qn.x; if (a.rk >= 3) a.R[o] = rout.x; }
                        if (v1) { dst[o + 1] = qn.y; if (a.rk >= 3) a.R[o + 1] = rout.y; }
                    }
                }
            }
        }
    }
    if (bad) a.sc->nanFlag = 1;

    if (a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        // controller bookkeeping (euler.go:177-182, :796-801)
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

// Host side: A-operand fragment table [DivInt | Vinv | V]; fragment (mt, ks), lane l holds Op[8 mt + l/4][4 ks + l%4].
template <int N> void build_mma_diss_frags(const double *DivInt, const double *Vinv, const double *V, std::vector<double> &out) {
    using MD = MmaDissDim<N>;
    out.assign(MD::kFragDoubles, 0.0);
    size_t f = 0;
    for (int mt = 0; mt < MD::M1; mt++)
        for (int ks = 0; ks < MD::K1; ks++, f++)
            for (int l = 0; l < 32; l++) {
                const int r = 8 * mt + l / 4, c = 4 * ks + l % 4;
                if (r < MD::NI && c < MD::NF) out[f * 32 + l] = DivInt[(size_t)r * MD::NF + c];
            }
    for (int which = 0; which < 2; which++) {
        const double *op = which == 0 ? Vinv : V;
        for (int mt = 0; mt < MD::M3; mt++)
            for (int ks = 0; ks < MD::K3; ks++, f++)
                for (int l = 0; l < 32; l++) {
                    const int r = 8 * mt + l / 4, c = 4 * ks + l % 4;
                    if (r < MD::NI && c < MD::NI) out[f * 32 + l] = op[(size_t)r * MD::NI + c];
                }
    }
}

}  // namespace dfr2d
