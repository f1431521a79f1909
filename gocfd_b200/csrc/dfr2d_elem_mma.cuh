// k_elem_mma<N>: inviscid element kernel with both thin contractions on the FP64 tensor cores.
//
//   RHS[NpInt x E]    = DivInt[NpInt x NpFlux]          . F_RT_DOF[NpFlux x E]   (RHSInternalPoints, euler.go:665-699)
//   Q_Face[3NpEdge x E] = FluxEdgeInterp[3NpEdge x NpInt] . q_new[NpInt x E]     (InterpolateSolutionToEdges, edges.go:485-491)
//
// Why tensor cores (north star: "FP64 DMMA ... only if ncu shows the high-P divergence is compute-bound"):
// on sm_100a ptxas feeds every DFMA of the unrolled contraction through its own LDCU.64 (uniform constant
// load); ncu showed k_elem<4> pinned at 31 % FP64-pipe utilisation and a stand-alone micro-benchmark of that
// loop tops out at 33 % of the FP64 peak, while DMMA.8x8x4 with register-resident operator fragments reaches
// 97 % (profiles/r01b_microbench_operator_delivery.txt).
//
// Structure: persistent CTAs (grid = SMs x resident CTAs), each looping over 32-element tiles.  CTA = 4 warps, warp w
// owns conserved variable w.  Operator fragments (A operands, zero padded to 8 x 4 tiles) are loaded once per CTA into
// registers.  Per tile:
//   1  row of the stage input -> smem sQ; gathered edge DOFs -> smem sF (SetRTFluxOnEdges); dt, -1/J per element
//   2  physical flux per point -> (Fr,Fs) into sF                                         (SetRTFluxInternal)
//   3  DMMA: C1 = DivInt . sF[var]; epilogue in fragment layout: -1/J, SSP-RK update with 128-bit global
//      accesses, fresh register -> global and -> sQ
//   4  DMMA: C2 = FEI . sQ[var]; 128-bit stores of Q_Face
// smem rows are padded to 36 doubles so that both the row-wise phase-1/2 accesses and the (4 k x 8 n) B-fragment
// loads are bank-conflict free.
#pragma once
#include "dfr2d_kernels.cuh"

#ifndef DFR2D_MMA_MINBLOCKS
#define DFR2D_MMA_MINBLOCKS 3
#endif

namespace dfr2d {

template <int N> struct MmaDim {
    static constexpr int NI = Dim<N>::NpInt, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3;
    static constexpr int M1 = (NI + 7) / 8, K1 = (NF + 3) / 4;     // DivInt tiles
    static constexpr int M2 = (NF3 + 7) / 8, K2 = (NI + 3) / 4;    // FluxEdgeInterp tiles
    static constexpr int SE = kElemsPerBlock + 4;                  // padded row stride (doubles)
    static constexpr int QROWS = 4 * K2, FROWS = 4 * K1;           // rows per variable incl. zero padding
    static constexpr int kFragDoubles = (M1 * K1 + M2 * K2) * 32;  // per-lane fragment table [frag][lane]
    static constexpr size_t kSmemBytes = (size_t)(4 * (QROWS + FROWS) * SE + 2 * kElemsPerBlock) * sizeof(double);
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct ElemMmaArgs {
    ElemArgs a;
    const double *frags;     // [M1*K1 + M2*K2][32] operator fragments in A-operand lane order
    int nTiles;
};

template <int N>
__global__ void __launch_bounds__(kElemThreads, DFR2D_MMA_MINBLOCKS) k_elem_mma(ElemMmaArgs args) {
    using MD = MmaDim<N>;
    constexpr int NI = MD::NI, NEd = Dim<N>::NpEdge, NF = MD::NF, NF3 = MD::NF3, E = kElemsPerBlock, SE = MD::SE;
    constexpr int M1 = MD::M1, K1 = MD::K1, M2 = MD::M2, K2 = MD::K2;
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
    const int lane = threadIdx.x & 31, n = threadIdx.x >> 5;
    const size_t Kp = a.Kp;

    // operator fragments -> registers, once per CTA
    double A1[M1][K1], A2[M2][K2];
#pragma unroll
    for (int mt = 0; mt < M1; mt++)
#pragma unroll
        for (int ks = 0; ks < K1; ks++) A1[mt][ks] = args.frags[(size_t)(mt * K1 + ks) * 32 + lane];
#pragma unroll
    for (int mt = 0; mt < M2; mt++)
#pragma unroll
        for (int ks = 0; ks < K2; ks++) A2[mt][ks] = args.frags[(size_t)(M1 * K1 + mt * K2 + ks) * 32 + lane];
    // zero the padding rows once (they are never written afterwards)
    for (int r = NI; r < MD::QROWS; r++) sQ[(n * MD::QROWS + r) * SE + lane] = 0.0;
    for (int r = NF; r < MD::FROWS; r++) sF[(n * MD::FROWS + r) * SE + lane] = 0.0;

    double *myQ = sQ + (size_t)n * MD::QROWS * SE;
    double *myF = sF + (size_t)n * MD::FROWS * SE;
    // global dt is the same for every tile (calculateGlobalDT, euler.go:945-971)
    double dtGlobal = 0.0;
    if (!a.ph.localDT) {
        const double gw = __longlong_as_double((long long)a.sc->wave[a.slot][0]);
        dtGlobal = a.ph.CFL / gw;
        const double t = a.sc->time[a.par];
        if (t + dtGlobal > a.ph.FinalTime) dtGlobal = a.ph.FinalTime - t;
    }
    bool bad = false;
    const int fr = lane >> 2, fc = lane & 3;     // fragment row / column of this lane

    for (int tile = blockIdx.x; tile < args.nTiles; tile += gridDim.x) {
        const int k0 = tile * E;
        const int k = k0 + lane;
        const bool valid = k < a.K;
        const int kc = valid ? k : a.K - 1;
        const bool fullTile = k0 + E <= a.K;
        __syncthreads();       // previous tile fully consumed before sQ / sF are overwritten

        if (a.pfTiles > 0) {
            // bulk L2 prefetch of this CTA's NEXT tile: stage input, the extra RK registers, geometry, edge-flux slots
            const long long kt = (long long)(tile + gridDim.x) * E;
            if (kt + E <= a.K) {
                constexpr unsigned RB = E * sizeof(double);
                for (int r = threadIdx.x; r < 4 * NI; r += kElemThreads) {
                    prefetch_l2(a.qs + (size_t)r * Kp + kt, RB);
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
                else if (threadIdx.x >= 32 && threadIdx.x < 32 + 4 * NEd) {
                    const long long s0 = ((kt * 3 / 2) / 16) * 16;
                    if (s0 + 64 <= a.NEp) prefetch_l2(a.eflux + (size_t)(threadIdx.x - 32) * a.NEp + s0, 64 * sizeof(double));
                }
            }
        }

        // ---- phase 1 -----------------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < NI; i++) myQ[i * SE + lane] = a.qs[((size_t)n * NI + i) * Kp + kc];
        {
            double wmaxk = -1.7976931348623157e308;
#pragma unroll
            for (int le = 0; le < 3; le++) {
                const int s = a.etoe[(size_t)le * Kp + kc];
                const bool owner = s >= 0;
                const int slot = owner ? s : -1 - s;
                const double iin = a.IInII[(size_t)le * Kp + kc];
                const double *f = a.eflux + ((size_t)n * NEd) * a.NEp + slot;
#pragma unroll
                for (int i = 0; i < NEd; i++) {
                    const double v = f[(size_t)(owner ? i : NEd - 1 - i) * a.NEp];
                    myF[(2 * NI + le * NEd + i) * SE + lane] = owner ? v * iin : -v * iin;
                }
                if (a.ph.localDT && n == 0) wmaxk = fmax(wmaxk, a.agg[slot]);
            }
            if (n == 0) {
                double dtk = dtGlobal;
                if (a.ph.localDT) {
                    const double d = (a.rk == 0) ? -100.0 : a.DT[kc];
                    dtk = a.ph.CFL / fmax(d, wmaxk);
                    if (valid && a.rhsOut == nullptr) a.DT[k] = dtk;
                }
                sDT[lane] = dtk;
                sMOOJ[lane] = -(1.0 / a.Jdet[kc]);
            }
        }
        __syncthreads();

        // ---- phase 2: SetRTFluxInternal, point j handled by warp j mod 4 ---------------------------------
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
                        sF[(m * MD::FROWS + j) * SE + lane] = jdet * (j0 * Fx[m] + j1 * Fy[m]);
                        sF[(m * MD::FROWS + j + NI) * SE + lane] = jdet * (j2 * Fx[m] + j3 * Fy[m]);
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3: C1 = DivInt . F on the tensor cores ------------------------------------------------
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
        // epilogue: lane holds rows i = 8 mt + fr, elements e0 = 8 nt + 2 fc (+1)
        double *dst = (a.rk == 0) ? a.q1 : (a.rk == 1) ? a.q2 : (a.rk == 2) ? a.q3 : (a.rk == 3) ? a.q4 : a.q0;
#pragma unroll
        for (int mt = 0; mt < M1; mt++) {
            const int i = 8 * mt + fr;
            if (i < NI) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const int e0 = 8 * nt + 2 * fc;
                    const double2 mo = *reinterpret_cast<const double2 *>(&sMOOJ[e0]);
                    const double rhs0 = c1[mt][nt][0] * mo.x, rhs1 = c1[mt][nt][1] * mo.y;
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
                    *reinterpret_cast<double2 *>(&myQ[i * SE + e0]) = qn;     // fresh register for the fused interpolation
                }
            }
        }
        if (a.rhsOut != nullptr || a.qface == nullptr) continue;
        __syncwarp();          // this warp wrote all of its variable's fresh rows

        // ---- phase 4: C2 = FluxEdgeInterp . q_new on the tensor cores -------------------------------------
        double c2[M2][4][2];
#pragma unroll
        for (int mt = 0; mt < M2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) c2[mt][nt][0] = c2[mt][nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < K2; ks++) {
            double b[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = myQ[(4 * ks + fc) * SE + 8 * nt + fr];
#pragma unroll
            for (int mt = 0; mt < M2; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(c2[mt][nt][0], c2[mt][nt][1], A2[mt][ks], b[nt]);
        }
#pragma unroll
        for (int mt = 0; mt < M2; mt++) {
            const int m = 8 * mt + fr;
            if (m < NF3) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const int e0 = 8 * nt + 2 * fc;
                    const size_t o = ((size_t)n * NF3 + m) * Kp + k0 + e0;
                    if (fullTile) {
                        *reinterpret_cast<double2 *>(a.qface + o) = make_double2(c2[mt][nt][0], c2[mt][nt][1]);
                    } else {
                        if (k0 + e0 < a.K) a.qface[o] = c2[mt][nt][0];
                        if (k0 + e0 + 1 < a.K) a.qface[o + 1] = c2[mt][nt][1];
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

// Host side: operator -> A-operand fragment table.  Fragment (mt, ks), lane l holds Op[8 mt + l/4][4 ks + l%4].
template <int N> void build_mma_frags(const double *DivInt, const double *FEI, std::vector<double> &out) {
    using MD = MmaDim<N>;
    out.assign(MD::kFragDoubles, 0.0);
    size_t f = 0;
    for (int mt = 0; mt < MD::M1; mt++)
        for (int ks = 0; ks < MD::K1; ks++, f++)
            for (int l = 0; l < 32; l++) {
                const int r = 8 * mt + l / 4, c = 4 * ks + l % 4;
                if (r < MD::NI && c < MD::NF) out[f * 32 + l] = DivInt[(size_t)r * MD::NF + c];
            }
    for (int mt = 0; mt < MD::M2; mt++)
        for (int ks = 0; ks < MD::K2; ks++, f++)
            for (int l = 0; l < 32; l++) {
                const int r = 8 * mt + l / 4, c = 4 * ks + l % 4;
                if (r < MD::NF3 && c < MD::NI) out[f * 32 + l] = FEI[(size_t)r * MD::NI + c];
            }
}

}  // namespace dfr2d
