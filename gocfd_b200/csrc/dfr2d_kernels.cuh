// Kernels of the inviscid stage.  Two launches per RK stage:
//   k_edge : numerical normal flux on every edge (+BCs) and wave-speed maxima      (edges.go:246-452)
//   k_elem : interior flux -> RT DOFs -> DivInt contraction -> dt -> SSP-RK update,
//            fused with the NEXT stage's solution-to-edge interpolation            (euler.go:460-566,
//            :665-726; edges.go:454-491)
// Layout: every element-indexed array is [row][Kp] with the element index fastest (the
// reference's own utils.Matrix layout), Kp = padded (own + ghost) element count.  Edge arrays are
// [var][point][NEp], edge index fastest.
#pragma once
#include "dfr2d_device.cuh"

namespace dfr2d {

constexpr int kElemsPerBlock = 32;              // E: elements per CTA of k_elem / k_interp
constexpr int kElemThreads = 4 * kElemsPerBlock; // one thread per (element, conserved variable)

struct EdgeArgs {
    int ne, NEp, Kp;
    int Kown;                           // own columns; an edge with a column >= Kown (a ghost) is a cut edge, left to the list pass
    const int *kL, *kR, *meta;          // kR < 0: boundary edge, boundary-point slot = -1 - kR
    const int *list;                    // optional: process only these edge slots (boundary and cut edges)
    int nlist;
    const double *nx, *ny, *oohk;
    const double *bpx, *bpy;            // [NBPloc][NpEdge]
    const double *qface;                // [4][3NpEdge][Kp]
    double *eflux;                      // [4][NpEdge][NEp]
    double *agg;                        // [NEp] edge max wave speed
    // VISC instantiations (PerssonC0 path, run after the gradient kernel): StoreEdgeViscousFlux (edges.go:151-244) is
    // evaluated on the spot and F - F_visc is what gets stored -- no separate viscous edge kernel, no read-modify-write
    const double *vn;                   // [2][4][NpEdge][NEp] owner-normal component of Epsilon (.) Grad (GradArgs::vn)
    const int *etov;                    // [3][Kp]
    const double *epsV;                 // vertex epsilon (ghost columns: private slots)
    const double *ooLen;                // [NEp] 1 / edge length
    double *aggv;                       // [NEp] viscous aggregate
    DevScalars *sc;
    int slot, par;
    long long stepIndex;
    Phys ph;
};

// The reference's loop is `for !finished { Step; Time += GlobalDT; steps++; finished = Time >= FinalTime ||
// steps >= MaxIterations }` (euler.go:175-182): step 0 always runs, step t >= 1 runs iff the time at its start is
// below FinalTime and t < MaxIterations.  time[par] is stable for the whole step, so every thread of every launch of
// a step evaluates the same predicate without any host round trip.
__device__ __forceinline__ bool step_is_noop(const DevScalars *sc, const Phys &ph, int par, long long stepIndex) {
    return stepIndex >= 1 && (sc->time[par] >= ph.FinalTime || stepIndex >= (long long)ph.maxIter);
}

// SSP54 coefficients (euler.go:511-563)
#define RK0_A 0.391752226571890
#define RK1_A 0.444370493651235
#define RK1_B 0.555629506348765
#define RK1_C 0.368410593050371
#define RK2_A 0.620101851488403
#define RK2_B 0.379898148511597
#define RK2_C 0.251891774271694
#define RK3_A 0.178079954393132
#define RK3_B 0.821920045606868
#define RK3_C 0.544974750228521
#define RK4_A 0.517231671970585
#define RK4_B 0.096059710526146
#define RK4_C 0.386708617503269
#define RK4_D 0.063692468666290
#define RK4_E 0.226007483236906

// ------------------------------------------------------------------------------------------------
// Edge kernel: one thread per edge, grid-stride.  CalculateEdgeEulerFlux + StoreEdgeAggregates.
// The BC routines overwrite Q_Face in place in the reference; the only later reader of the
// overwritten rows is StoreEdgeAggregates, so the post-BC state stays in registers here.
// ------------------------------------------------------------------------------------------------
#ifndef DFR2D_EDGEINT_MINBLOCKS
#define DFR2D_EDGEINT_MINBLOCKS 3
#endif
#ifndef DFR2D_EDGE_MINBLOCKS
#define DFR2D_EDGE_MINBLOCKS 2
#endif
// PPT = edge points per thread.  PPT == NpEdge: one thread per edge (per-edge aggregate written directly).  Smaller PPT:
// the points of an edge are split over NpEdge/PPT threads in different warps -- fewer registers, more warps in flight
// for this latency-bound gather kernel; the per-edge aggregate (only consumed with local time stepping) is then
// combined with atomicMax on the bit pattern (agg is zeroed by the host beforehand).
// Viscous state of one edge for the VISC edge kernels: vertex epsilons of both elements (InterpolateEpsilonSigma,
// dissipation.go:219-242, re-derived with Bary at the edge rows), penalty coefficient, viscous aggregate.
template <int N>
struct EdgeVisc {
    double eL0, eL1, eL2, eR0, eR1, eR2, ooLen, oohk2, vmax;
    int numL, numR;
    bool shared;
    __device__ __forceinline__ void load(const EdgeArgs &a, int e, int kL, int kR, int nL, int nR, double oohk) {
        const size_t Kp = a.Kp;
        shared = kR >= 0;
        numL = nL; numR = nR;
        eL0 = a.epsV[a.etov[kL]]; eL1 = a.epsV[a.etov[Kp + kL]]; eL2 = a.epsV[a.etov[2 * Kp + kL]];
        eR0 = eR1 = eR2 = 0.0;
        if (shared) { eR0 = a.epsV[a.etov[kR]]; eR1 = a.epsV[a.etov[Kp + kR]]; eR2 = a.epsV[a.etov[2 * Kp + kR]]; }
        ooLen = a.ooLen[e];
        oohk2 = oohk * oohk;
        vmax = -1.7976931348623157e308;
    }
    // F[n] -= viscous normal flux at owner point i; QLpre = the owner's stored edge values (EdgeQValues, pre-BC)
    __device__ __forceinline__ void apply(const EdgeArgs &a, int e, int i, int kL, const double (&QLpre)[4], double (&F)[4]) {
        constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge;
        const Ops<N> &op = ops<N>();
        const int rowL = 2 * NI + numL * NEd + i;
        const double epsL = op.Bary[rowL][0] * eL0 + op.Bary[rowL][1] * eL1 + op.Bary[rowL][2] * eL2;
        vmax = fmax(oohk2 * epsL, vmax);
        const size_t fplane = (size_t)NEd * a.NEp, qplane = (size_t)Dim<N>::NF3 * a.Kp;
        double lam = 0.0;
        if (shared) {
            const int rowR = 2 * NI + numR * NEd + (NEd - 1 - i);
            const double epsR = op.Bary[rowR][0] * eR0 + op.Bary[rowR][1] * eR1 + op.Bary[rowR][2] * eR2;
            lam = 0.5 * (epsL + epsR);
        }
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const size_t idx = ((size_t)n * NEd + i) * a.NEp + e;
            const double vFL = a.vn[idx];
            double vf = vFL;
            if (shared) {
                const double vFR = a.vn[idx + 4 * fplane];
                vf = 0.5 * (vFL + vFR);
                // both "sides" of the jump resolve to the owner's stored edge values (edges.go:225-236)
                const double qb = a.qface[n * qplane + (size_t)(numL * NEd + (NEd - 1 - i)) * a.Kp + kL];
                vf -= (a.ph.Omega * lam * ooLen) * (QLpre[n] - qb);
            }
            F[n] -= vf;
        }
    }
};

template <int N, int PPT, bool VISC>
__global__ void __launch_bounds__(256, DFR2D_EDGE_MINBLOCKS) k_edge(EdgeArgs a) {
    constexpr int NE_ = Dim<N>::NpEdge;
    constexpr int G = NE_ / PPT;
    static_assert(NE_ % PPT == 0, "points per thread must divide NpEdge");
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    const double gamma = a.ph.gamma;
    const size_t qplane = (size_t)Dim<N>::NF3 * a.Kp;
    const size_t fplane = (size_t)NE_ * a.NEp;
    double blockmax = 0.0, blockmaxV = 0.0;
    const int span = a.list ? a.nlist : a.NEp;
    const long long total = (long long)G * span;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t / span);
        const int e = a.list ? a.list[t % span] : (int)(t % span);
        if (e >= a.ne) continue;
        const int kL = a.kL[e], kR = a.kR[e], meta = a.meta[e];
        const int numL = meta & 3, numR = (meta >> 2) & 3, bc = (meta >> 4) & 15;
        const double nx = a.nx[e], ny = a.ny[e], oohk = a.oohk[e];
        double wmax = -1.7976931348623157e308;
        EdgeVisc<N> ev;
        if (VISC) ev.load(a, e, kL, kR >= 0 ? kR : -1, numL, numR, oohk);
#pragma unroll
        for (int ii = 0; ii < PPT; ii++) {
            const int i = g * PPT + ii;
            double QL[4], F[4], QLpre[4];
            double wL = 0.0;
            bool haveW = false;
            const size_t offL = (size_t)(numL * NE_ + i) * a.Kp + kL;
#pragma unroll
            for (int n = 0; n < 4; n++) { QL[n] = a.qface[n * qplane + offL]; QLpre[n] = QL[n]; }
            if (kR >= 0) {
                double QR[4];
                const size_t offR = (size_t)(numR * NE_ + (NE_ - 1 - i)) * a.Kp + kR;
#pragma unroll
                for (int n = 0; n < 4; n++) QR[n] = a.qface[n * qplane + offR];
                switch (a.ph.fluxType) {
                    case DFR2D_FLUX_Average: avg_flux(gamma, QL, QR, nx, ny, F); break;
                    case DFR2D_FLUX_LaxFriedrichs: lax_flux(gamma, QL, QR, nx, ny, F); break;
                    case DFR2D_FLUX_Roe: roe_flux(gamma, QL, QR, nx, ny, F, wL); haveW = true; break;
                    default: roe_er_flux(gamma, QL, QR, nx, ny, F); break;
                }
            } else {
                // calculateNonSharedEdgeFlux (edges.go:413-452)
                bool normalFlux = true;
                if (bc == DFR2D_BC_Far || bc == DFR2D_BC_In || bc == DFR2D_BC_Out) {
                    const dfr2d_freestream &FS = a.ph.fs[bc == DFR2D_BC_Far ? 0 : (bc == DFR2D_BC_In ? 1 : 2)];
                    double QB[4];
                    riemann_bc(FS, QL, FS.Qinf, nx, ny, QB);
#pragma unroll
                    for (int n = 0; n < 4; n++) QL[n] = QB[n];
                } else if (bc == DFR2D_BC_IVortex) {
                    const int b = -1 - kR;
                    double QX[4], QB[4];
                    ivortex_state(a.ph.vortex, a.sc->time[a.par], a.bpx[(size_t)b * NE_ + i], a.bpy[(size_t)b * NE_ + i], QX);
                    riemann_bc(a.ph.fs[0], QL, QX, nx, ny, QB);
#pragma unroll
                    for (int n = 0; n < 4; n++) QL[n] = QB[n];
                } else if (bc == DFR2D_BC_Wall || bc == DFR2D_BC_Cyl) {
                    normalFlux = false;
                    double p = static_pressure(gamma, QL[0], QL[1], QL[2], QL[3]);
                    F[0] = 0; F[1] = nx * p; F[2] = ny * p; F[3] = 0;
                } else if (bc == DFR2D_BC_Periodic || bc == DFR2D_BC_PeriodicReversed) {
                    normalFlux = false;
                    F[0] = F[1] = F[2] = F[3] = 0;
                }
                if (normalFlux) {
                    double Fx[4], Fy[4];
                    flux_calc(gamma, QL, Fx, Fy);
#pragma unroll
                    for (int n = 0; n < 4; n++) F[n] = nx * Fx[n] + ny * Fy[n];
                }
            }
            if (VISC) ev.apply(a, e, i, kL, QLpre, F);
#pragma unroll
            for (int n = 0; n < 4; n++) a.eflux[n * fplane + (size_t)i * a.NEp + e] = F[n];
            // StoreEdgeAggregates: owner side, post-BC state (edges.go:260-273)
            if (!haveW) wL = speed_plus_sound(gamma, QL[0], QL[1], QL[2], QL[3]);
            const double w = oohk * wL;
            if (w > wmax) wmax = w;
        }
        if (G == 1) a.agg[e] = wmax;
        else if (a.ph.localDT) atomic_max_nonneg(reinterpret_cast<unsigned long long *>(&a.agg[e]), wmax);
        blockmax = fmax(blockmax, wmax);
        if (VISC) {
            if (G == 1) a.aggv[e] = ev.vmax;
            else if (a.ph.localDT) atomic_max_nonneg(reinterpret_cast<unsigned long long *>(&a.aggv[e]), fmax(ev.vmax, 0.0));
            blockmaxV = fmax(blockmaxV, ev.vmax);
        }
    }
    __shared__ double smax[8], smaxV[8];
    blockmax = warp_max(blockmax);
    if (VISC) blockmaxV = warp_max(blockmaxV);
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = blockmax; if (VISC) smaxV[threadIdx.x >> 5] = blockmaxV; }
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? smax[threadIdx.x] : 0.0;
        v = warp_max(v);
        if (threadIdx.x == 0) atomic_max_nonneg(&a.sc->wave[a.slot][0], v);
        if (VISC) {
            double vv = threadIdx.x < (blockDim.x >> 5) ? smaxV[threadIdx.x] : 0.0;
            vv = warp_max(vv);
            if (threadIdx.x == 0) atomic_max_nonneg(&a.sc->wave[a.slot][1], vv);
        }
    }
}

// Interior (shared) edges only, flux type fixed at compile time: none of the boundary-condition code (pow/exp, three
// free-stream records) is in this kernel, which halves its register count and doubles the warps in flight.
// Boundary edges are skipped here and handled by k_edge<N,PPT> over the compact boundary list.
template <int N, int FLUX, int PPT, bool VISC>
__global__ void __launch_bounds__(256, VISC ? 2 : DFR2D_EDGEINT_MINBLOCKS) k_edge_int(EdgeArgs a) {
    constexpr int NE_ = Dim<N>::NpEdge;
    constexpr int G = NE_ / PPT;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    const double gamma = a.ph.gamma;
    const size_t qplane = (size_t)Dim<N>::NF3 * a.Kp;
    const size_t fplane = (size_t)NE_ * a.NEp;
    double blockmax = 0.0, blockmaxV = 0.0;
    const long long total = (long long)G * a.NEp;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t % a.NEp), g = (int)(t / a.NEp);
        if (e >= a.ne) continue;
        const int kR = a.kR[e];
        if (kR < 0 || kR >= a.Kown) continue;
        const int kL = a.kL[e], meta = a.meta[e];
        if (kL >= a.Kown) continue;         // cut edge: needs the halo, evaluated after the exchange
        const int numL = meta & 3, numR = (meta >> 2) & 3;
        const double nx = a.nx[e], ny = a.ny[e], oohk = a.oohk[e];
        double wmax = -1.7976931348623157e308;
        EdgeVisc<N> ev;
        if (VISC) ev.load(a, e, kL, kR, numL, numR, oohk);
#pragma unroll
        for (int ii = 0; ii < PPT; ii++) {
            const int i = g * PPT + ii;
            double QL[4], QR[4], F[4], wL;
            const size_t offL = (size_t)(numL * NE_ + i) * a.Kp + kL;
            const size_t offR = (size_t)(numR * NE_ + (NE_ - 1 - i)) * a.Kp + kR;
#pragma unroll
            for (int n = 0; n < 4; n++) { QL[n] = a.qface[n * qplane + offL]; QR[n] = a.qface[n * qplane + offR]; }
            if (FLUX == DFR2D_FLUX_Roe) roe_flux(gamma, QL, QR, nx, ny, F, wL);
            else {
                if (FLUX == DFR2D_FLUX_Average) avg_flux(gamma, QL, QR, nx, ny, F);
                else if (FLUX == DFR2D_FLUX_LaxFriedrichs) lax_flux(gamma, QL, QR, nx, ny, F);
                else roe_er_flux(gamma, QL, QR, nx, ny, F);
                wL = speed_plus_sound(gamma, QL[0], QL[1], QL[2], QL[3]);
            }
            if (VISC) ev.apply(a, e, i, kL, QL, F);
#pragma unroll
            for (int n = 0; n < 4; n++) a.eflux[n * fplane + (size_t)i * a.NEp + e] = F[n];
            const double w = oohk * wL;
            if (w > wmax) wmax = w;
        }
        if (G == 1) a.agg[e] = wmax;
        else if (a.ph.localDT) atomic_max_nonneg(reinterpret_cast<unsigned long long *>(&a.agg[e]), wmax);
        blockmax = fmax(blockmax, wmax);
        if (VISC) {
            if (G == 1) a.aggv[e] = ev.vmax;
            else if (a.ph.localDT) atomic_max_nonneg(reinterpret_cast<unsigned long long *>(&a.aggv[e]), fmax(ev.vmax, 0.0));
            blockmaxV = fmax(blockmaxV, ev.vmax);
        }
    }
    __shared__ double smax[8], smaxV[8];
    blockmax = warp_max(blockmax);
    if (VISC) blockmaxV = warp_max(blockmaxV);
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = blockmax; if (VISC) smaxV[threadIdx.x >> 5] = blockmaxV; }
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? smax[threadIdx.x] : 0.0;
        v = warp_max(v);
        if (threadIdx.x == 0) atomic_max_nonneg(&a.sc->wave[a.slot][0], v);
        if (VISC) {
            double vv = threadIdx.x < (blockDim.x >> 5) ? smaxV[threadIdx.x] : 0.0;
            vv = warp_max(vv);
            if (threadIdx.x == 0) atomic_max_nonneg(&a.sc->wave[a.slot][1], vv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Q_Face[n][m][k] = sum_i FluxEdgeInterp[m][i] * q[i]   (edges.go:485-491), thread = (element, n)
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void interp_store(const double (&q)[Dim<N>::NpInt], double *qface_n, int Kp, int k) {
    const Ops<N> &op = ops<N>();
#pragma unroll
    for (int m = 0; m < Dim<N>::NF3; m++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < Dim<N>::NpInt; i++) s = fma(op.FEI[m][i], q[i], s);
        qface_n[(size_t)m * Kp + k] = s;
    }
}

template <int N>
__global__ void __launch_bounds__(kElemThreads) k_interp(int K, int Kp, const double *q, double *qface) {
    const int e = threadIdx.x % kElemsPerBlock, n = threadIdx.x / kElemsPerBlock;
    const int k = blockIdx.x * kElemsPerBlock + e;
    if (k >= K) return;
    double qs[Dim<N>::NpInt];
#pragma unroll
    for (int i = 0; i < Dim<N>::NpInt; i++) qs[i] = q[((size_t)n * Dim<N>::NpInt + i) * Kp + k];
    interp_store<N>(qs, qface + (size_t)n * Dim<N>::NF3 * Kp, Kp, k);
}

template <int N> __device__ __forceinline__ void limit_filter_row(double (&u)[Dim<N>::NpInt], double sigmaK);

// Bulk L2 prefetch of a contiguous byte range (sm_90+): decouples the HBM->L2 stream of a FUTURE tile from
// the load phase of the CTA that will consume it (the kernel is bulk-synchronous with few resident warps).
__device__ __forceinline__ void prefetch_l2(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct ElemArgs {
    int K, Kp, NEp;
    int pfTiles;                            // prefetch distance in tiles (0 = off)
    const double *qs;                       // stage input register [4][NpInt][Kp]
    double *q0, *q1, *q2, *q3, *q4, *R;     // c.Q, Q1..Q4, Residual
    double *qface;                          // [4][3NpEdge][Kp], written for the next stage when non-null
    const double *eflux;                    // [4][NpEdge][NEp] numerical normal flux
    const double *agg, *aggv;               // [NEp]
    double *DT, *DTVisc;                    // [Kp] (local time stepping)
    const double *Jdet, *Jinv, *IInII;      // [Kp], [4][Kp], [3][Kp]
    const int *etoe;                        // [3][Kp]: edge slot, or -1 - slot when the neighbour owns it
    const double *dissX, *dissY;            // [4][NpInt][Kp]: interior rows of Epsilon (.) Grad (dissipation)
    const double *sigma;                    // [Kp] vertex-averaged sigma (dissipation)
    double *rhsOut;                         // test hook: RHSQ only, no update
    DevScalars *sc;
    int rk, slot, par;
    long long stepIndex;
    Phys ph;
};

// Element kernel.  CTA = 32 elements x 4 conserved variables.
//   phase 1  thread (e,n) loads its solution row into registers and shared memory, and gathers
//            the numerical flux of its three edges (sign / reversal / IInII applied)
//   phase 2  per-point physical flux -> (Fr, Fs) RT DOFs into shared memory (euler.go:701-726)
//   phase 3  DivInt contraction from shared memory with the operator as constant-bank operands,
//            -1/Jdet, dt, SSP-RK update, NaN flag, next stage's edge interpolation
template <int N, bool DISS>
__global__ void __launch_bounds__(kElemThreads) k_elem(ElemArgs a) {
    constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF3 = Dim<N>::NF3;
    constexpr int E = kElemsPerBlock;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) {
        if (a.rk == 4 && a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            a.sc->time[a.par ^ 1] = a.sc->time[a.par];     // keep both parities at the final time
            a.sc->finished = 1;
        }
        return;
    }
    extern __shared__ double smem[];
    double *sQ = smem;                    // [4][NI][E]
    double *sF = smem + 4 * NI * E;       // [4][2NI][E]
    const Ops<N> &op = ops<N>();
    const int e = threadIdx.x % E, n = threadIdx.x / E;
    const int k = blockIdx.x * E + e;
    const bool valid = k < a.K;
    const int kc = valid ? k : a.K - 1;   // clamp: out-of-range lanes compute on a valid element, store nothing
    const size_t Kp = a.Kp;

    if (a.pfTiles > 0) {
        // rows of the tile `pfTiles` ahead: stage input, the extra RK registers of this stage, geometry, and the
        // edge-flux slots around the tile (edges are sorted by owner column: slot ~ 1.5 * column)
        const long long kt = (long long)(blockIdx.x + a.pfTiles) * E;
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

    double qs[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) {
        qs[i] = a.qs[((size_t)n * NI + i) * Kp + kc];
        sQ[(n * NI + i) * E + e] = qs[i];
    }
    const double jdet = a.Jdet[kc];
    // edge rows of F_RT_DOF for variable n (SetRTFluxOnEdges, edges.go:454-483)
    double fe[NF3];
    double dtk = 0.0, dtv = 0.0;
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
                const double v = f[(size_t)(owner ? i : NEd - 1 - i) * a.NEp];
                fe[le * NEd + i] = owner ? v * iin : -v * iin;
            }
            // (DISS: k_visc_edge has already subtracted the viscous normal flux from eflux in place -- AddDissipation's
            // edge DOFs, dissipation.go:316-333, enter the same contraction with the opposite sign)
            if (a.ph.localDT) {
                wmaxk = fmax(wmaxk, a.agg[slot]);
                if (DISS) vmaxk = fmax(vmaxk, a.aggv[slot]);
            }
        }
        if (a.ph.localDT) {
            // InitializeDT at stage 0, then DT = max(DT, edge aggregates) (euler.go:637-643, edges.go:291-323)
            double d = (a.rk == 0) ? -100.0 : a.DT[kc];
            dtk = a.ph.CFL / fmax(d, wmaxk);                       // CalculateLocalDT (euler.go:985-988)
            if (DISS) {
                dtv = fmax(a.DTVisc[kc], vmaxk);
                if (dtv > 1.e-9) { dtv = a.ph.Cdiff / dtv; dtk = fmin(dtk, dtv); }
            }
        } else {
            // calculateGlobalDT (euler.go:945-971); every thread derives the same scalar
            const double gw = __longlong_as_double((long long)a.sc->wave[a.slot][0]);
            dtk = a.ph.CFL / gw;
            if (DISS) {
                const double gv = __longlong_as_double((long long)a.sc->wave[a.slot][1]);
                dtk = fmin(dtk, a.ph.Cdiff / gv);
            }
            const double t = a.sc->time[a.par];
            if (t + dtk > a.ph.FinalTime) dtk = a.ph.FinalTime - t;
        }
    }
    __syncthreads();

    // phase 2: SetRTFluxInternal for points j = n, n+4, ...
    {
        const double j0 = a.Jinv[0 * Kp + kc], j1 = a.Jinv[1 * Kp + kc], j2 = a.Jinv[2 * Kp + kc], j3 = a.Jinv[3 * Kp + kc];
#pragma unroll
        for (int j = n; j < NI; j += 4) {
            double Q[4], Fx[4], Fy[4];
#pragma unroll
            for (int m = 0; m < 4; m++) Q[m] = sQ[(m * NI + j) * E + e];
            flux_calc(a.ph.gamma, Q, Fx, Fy);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                double fr = jdet * (j0 * Fx[m] + j1 * Fy[m]);
                double fs = jdet * (j2 * Fx[m] + j3 * Fy[m]);
                if (DISS) {   // interior dissipation DOFs (dissipation.go:310-315)
                    const double dix = a.dissX[((size_t)m * NI + j) * Kp + kc];
                    const double diy = a.dissY[((size_t)m * NI + j) * Kp + kc];
                    fr -= jdet * (j0 * dix + j1 * diy);
                    fs -= jdet * (j2 * dix + j3 * diy);
                }
                sF[(m * 2 * NI + j) * E + e] = fr;
                sF[(m * 2 * NI + j + NI) * E + e] = fs;
            }
        }
    }
    __syncthreads();

    // phase 3: RHSInternalPoints (euler.go:665-699)
    double acc[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) acc[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 2 * NI; j++) {
        const double f = sF[(n * 2 * NI + j) * E + e];
#pragma unroll
        for (int i = 0; i < NI; i++) acc[i] = fma(op.DivInt[i][j], f, acc[i]);
    }
#pragma unroll
    for (int j = 0; j < NF3; j++) {
#pragma unroll
        for (int i = 0; i < NI; i++) acc[i] = fma(op.DivInt[i][2 * NI + j], fe[j], acc[i]);
    }
    const double moojd = -(1.0 / jdet);
#pragma unroll
    for (int i = 0; i < NI; i++) acc[i] *= moojd;
    if (DISS) limit_filter_row<N>(acc, a.sigma[kc]);     // LimitFilterSolution(RHSQ) (euler.go:499)

    if (a.rhsOut != nullptr) {
        if (valid) {
#pragma unroll
            for (int i = 0; i < NI; i++) a.rhsOut[((size_t)n * NI + i) * Kp + k] = acc[i];
        }
        return;
    }

    // SSP-RK(5,4) combination (euler.go:502-565); qs is the stage input register
    const size_t base = (size_t)n * NI * Kp + kc;
    double qn[NI];
    bool bad = false;
    double *dst;
    switch (a.rk) {
        case 0:
            dst = a.q1;
#pragma unroll
            for (int i = 0; i < NI; i++) qn[i] = qs[i] + RK0_A * (dtk * acc[i]);
            break;
        case 1:
            dst = a.q2;
#pragma unroll
            for (int i = 0; i < NI; i++)
                qn[i] = RK1_A * a.q0[base + i * Kp] + RK1_B * qs[i] + RK1_C * (dtk * acc[i]);
            break;
        case 2:
            dst = a.q3;
#pragma unroll
            for (int i = 0; i < NI; i++)
                qn[i] = RK2_A * a.q0[base + i * Kp] + RK2_B * qs[i] + RK2_C * (dtk * acc[i]);
            break;
        case 3:
            dst = a.q4;
#pragma unroll
            for (int i = 0; i < NI; i++) {
                qn[i] = RK3_A * a.q0[base + i * Kp] + RK3_B * qs[i] + RK3_C * (dtk * acc[i]);
                if (valid) a.R[base + i * Kp] = acc[i];      // keep RHS(q3) for the last stage
            }
            break;
        default:
            dst = a.q0;
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const double q0 = a.q0[base + i * Kp];
                const double dtR3 = dtk * a.R[base + i * Kp];
                const double r = -q0 + RK4_A * a.q2[base + i * Kp] + RK4_B * a.q3[base + i * Kp] + RK4_C * qs[i] +
                                 RK4_D * dtR3 + RK4_E * (dtk * acc[i]);
                if (valid) a.R[base + i * Kp] = r;
                qn[i] = q0 + r;
            }
            break;
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
        bad |= (qn[i] != qn[i]);
        if (valid) dst[base + i * Kp] = qn[i];
    }
    if (bad && valid) a.sc->nanFlag = 1;     // utils.IsNanPanic (euler.go:471-486)
    if (a.ph.localDT && valid && n == 0) {
        a.DT[k] = dtk;
        if (DISS) a.DTVisc[k] = dtv;
    }
    if (a.qface != nullptr && valid) interp_store<N>(qn, a.qface + (size_t)n * NF3 * Kp, a.Kp, k);

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // controller bookkeeping (euler.go:177-182, :796-801), one thread per launch
        a.sc->wave[a.slot ^ 1][0] = 0ull;
        a.sc->wave[a.slot ^ 1][1] = 0ull;
        if (!a.ph.localDT) a.sc->globalDT = dtk;
        if (a.rk == 4) {
            const double tnew = a.sc->time[a.par] + (a.ph.localDT ? a.sc->globalDT : dtk);
            a.sc->time[a.par ^ 1] = tnew;
            a.sc->timeOut = tnew;
            const long long st = a.sc->steps + 1;
            a.sc->steps = st;
            if (tnew >= a.ph.FinalTime || st >= (long long)a.ph.maxIter) a.sc->finished = 1;   // CheckIfFinished
        }
    }
}


// ------------------------------------------------------------------------------------------------
// k_elem2: the inviscid element kernel with each (element, variable) row split over two threads
// (even / odd output nodes).  Same arithmetic as k_elem<N,false>; twice the resident warps at half
// the registers, the gathered edge fluxes staged in shared memory, and the fresh register exchanged
// through shared memory for the fused edge interpolation.  CTA = 32 elements x 4 variables x 2 halves
// = 8 warps; warp w handles variable (w & 3), half (w >> 2) so the operator stays a warp-uniform
// constant-bank operand.  ncu on the first version showed k_elem latency-bound at 12 warps/SM
// (profiles/r01a_*): issue slots 41 %, FP64 pipe 31 %, DRAM 30 %.
// ------------------------------------------------------------------------------------------------
constexpr int kElem2Threads = 8 * kElemsPerBlock;

template <int N, int HF>
__device__ __forceinline__ void elem2_tail(const ElemArgs &a, double *sQ, const double *sF, int e, int n, int k, int kc,
                                           bool valid, double jdet, double dtk) {
    constexpr int NI = Dim<N>::NpInt, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3, E = kElemsPerBlock;
    constexpr int NO = (NI - HF + 1) / 2;          // my output nodes i = HF + 2c
    constexpr int NM = (NF3 - HF + 1) / 2;         // my Q_Face rows m = HF + 2c
    const Ops<N> &op = ops<N>();
    const size_t Kp = a.Kp;
    double acc[NO > 0 ? NO : 1];
#pragma unroll
    for (int c = 0; c < NO; c++) acc[c] = 0.0;
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const double f = sF[(n * NF + j) * E + e];
#pragma unroll
        for (int c = 0; c < NO; c++) acc[c] = fma(op.DivInt[HF + 2 * c][j], f, acc[c]);
    }
    const double moojd = -(1.0 / jdet);
#pragma unroll
    for (int c = 0; c < NO; c++) acc[c] *= moojd;
    if (a.rhsOut != nullptr) {
        if (valid) {
#pragma unroll
            for (int c = 0; c < NO; c++) a.rhsOut[((size_t)n * NI + HF + 2 * c) * Kp + k] = acc[c];
        }
        return;
    }
    const size_t base = (size_t)n * NI * Kp + kc;
    double *dst = (a.rk == 0) ? a.q1 : (a.rk == 1) ? a.q2 : (a.rk == 2) ? a.q3 : (a.rk == 3) ? a.q4 : a.q0;
    bool bad = false;
#pragma unroll
    for (int c = 0; c < NO; c++) {
        const int i = HF + 2 * c;
        const size_t o = base + (size_t)i * Kp;
        const double qsi = sQ[(n * NI + i) * E + e];
        const double rhs = acc[c];
        double qn;
        switch (a.rk) {
            case 0: qn = qsi + RK0_A * (dtk * rhs); break;
            case 1: qn = RK1_A * a.q0[o] + RK1_B * qsi + RK1_C * (dtk * rhs); break;
            case 2: qn = RK2_A * a.q0[o] + RK2_B * qsi + RK2_C * (dtk * rhs); break;
            case 3:
                qn = RK3_A * a.q0[o] + RK3_B * qsi + RK3_C * (dtk * rhs);
                if (valid) a.R[o] = rhs;
                break;
            default: {
                const double q0 = a.q0[o];
                const double dtR3 = dtk * a.R[o];
                const double r = -q0 + RK4_A * a.q2[o] + RK4_B * a.q3[o] + RK4_C * qsi + RK4_D * dtR3 + RK4_E * (dtk * rhs);
                if (valid) a.R[o] = r;
                qn = q0 + r;
            } break;
        }
        bad |= (qn != qn);
        if (valid) dst[o] = qn;
        acc[c] = qn;
    }
    if (bad && valid) a.sc->nanFlag = 1;
    if (a.qface == nullptr) return;            // (uniform) no fused interpolation requested
    __syncthreads();                           // every thread has read its stage-input values from sQ
#pragma unroll
    for (int c = 0; c < NO; c++) sQ[(n * NI + HF + 2 * c) * E + e] = acc[c];
    __syncthreads();
    if (!valid) return;
    double *qf = a.qface + (size_t)n * NF3 * Kp + k;
#pragma unroll
    for (int c = 0; c < NM; c++) {
        const int m = HF + 2 * c;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NI; i++) s = fma(op.FEI[m][i], sQ[(n * NI + i) * E + e], s);
        qf[(size_t)m * Kp] = s;
    }
}

template <int N>
__global__ void __launch_bounds__(kElem2Threads, 3) k_elem2(ElemArgs a) {
    constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF = Dim<N>::NpFlux, E = kElemsPerBlock;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) {
        if (a.rk == 4 && a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            a.sc->time[a.par ^ 1] = a.sc->time[a.par];
            a.sc->finished = 1;
        }
        return;
    }
    extern __shared__ double smem[];
    double *sQ = smem;                    // [4][NI][E]   stage input, later the fresh register
    double *sF = smem + 4 * NI * E;       // [4][NF][E]   RT DOFs: interior (Fr, Fs) rows then the edge rows
    const int e = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = w & 3, hf = w >> 2;
    const int k = blockIdx.x * E + e;
    const bool valid = k < a.K;
    const int kc = valid ? k : a.K - 1;
    const size_t Kp = a.Kp;

    // phase 1: my half of the row, my half of the edge DOFs (SetRTFluxOnEdges, edges.go:454-483), dt
    for (int i = hf; i < NI; i += 2) sQ[(n * NI + i) * E + e] = a.qs[((size_t)n * NI + i) * Kp + kc];
    const double jdet = a.Jdet[kc];
    double dtk;
    {
        double wmaxk = -1.7976931348623157e308;
#pragma unroll
        for (int le = 0; le < 3; le++) {
            const int s = a.etoe[(size_t)le * Kp + kc];
            const bool owner = s >= 0;
            const int slot = owner ? s : -1 - s;
            const double iin = a.IInII[(size_t)le * Kp + kc];
            const double *f = a.eflux + ((size_t)n * NEd) * a.NEp + slot;
            for (int r = le * NEd + ((le * NEd + hf) & 1); r < (le + 1) * NEd; r += 2) {
                const int i = r - le * NEd;
                const double v = f[(size_t)(owner ? i : NEd - 1 - i) * a.NEp];
                sF[(n * NF + 2 * NI + r) * E + e] = owner ? v * iin : -v * iin;
            }
            if (a.ph.localDT) wmaxk = fmax(wmaxk, a.agg[slot]);
        }
        if (a.ph.localDT) {
            const double d = (a.rk == 0) ? -100.0 : a.DT[kc];
            dtk = a.ph.CFL / fmax(d, wmaxk);
        } else {
            const double gw = __longlong_as_double((long long)a.sc->wave[a.slot][0]);
            dtk = a.ph.CFL / gw;
            const double t = a.sc->time[a.par];
            if (t + dtk > a.ph.FinalTime) dtk = a.ph.FinalTime - t;
        }
    }
    __syncthreads();

    // phase 2: SetRTFluxInternal, point j handled by warp j mod 8
    {
        const double j0 = a.Jinv[0 * Kp + kc], j1 = a.Jinv[1 * Kp + kc], j2 = a.Jinv[2 * Kp + kc], j3 = a.Jinv[3 * Kp + kc];
        for (int j = w; j < NI; j += 8) {
            double Q[4], Fx[4], Fy[4];
#pragma unroll
            for (int m = 0; m < 4; m++) Q[m] = sQ[(m * NI + j) * E + e];
            flux_calc(a.ph.gamma, Q, Fx, Fy);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                sF[(m * NF + j) * E + e] = jdet * (j0 * Fx[m] + j1 * Fy[m]);
                sF[(m * NF + j + NI) * E + e] = jdet * (j2 * Fx[m] + j3 * Fy[m]);
            }
        }
    }
    __syncthreads();

    // phase 3/4: contraction, update, fused interpolation -- per half so that operator rows are compile-time
    if (hf == 0) elem2_tail<N, 0>(a, sQ, sF, e, n, k, kc, valid, jdet, dtk);
    else elem2_tail<N, 1>(a, sQ, sF, e, n, k, kc, valid, jdet, dtk);

    if (a.rhsOut == nullptr) {
        if (a.ph.localDT && valid && w == 0) a.DT[k] = dtk;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.sc->wave[a.slot ^ 1][0] = 0ull;
            a.sc->wave[a.slot ^ 1][1] = 0ull;
            if (!a.ph.localDT) a.sc->globalDT = dtk;
            if (a.rk == 4) {
                const double tnew = a.sc->time[a.par] + (a.ph.localDT ? a.sc->globalDT : dtk);
                a.sc->time[a.par ^ 1] = tnew;
                a.sc->timeOut = tnew;
                const long long st = a.sc->steps + 1;
                a.sc->steps = st;
                if (tnew >= a.ph.FinalTime || st >= (long long)a.ph.maxIter) a.sc->finished = 1;
            }
        }
    }
}

}  // namespace dfr2d
