// Utility kernels (halo pack/unpack, residual reduction) and the artificial-dissipation path.
#pragma once
#include "dfr2d_kernels.cuh"
#include "dfr2d_peer.cuh"

namespace dfr2d {

struct DissBuffers {
    int *etov = nullptr;                 // [3][Kp] global vertex ids
    double *hk = nullptr;                // [Kp]
    double *nxk = nullptr, *nyk = nullptr;   // [3][Kp] own face normals
    double *eooLen = nullptr;            // [NEp] 1 / edge length
    double *sigma = nullptr, *epsk = nullptr, *se = nullptr;   // [Kp]
    double *sigmaV = nullptr, *epsV = nullptr;                  // [NV]
    double *dissX = nullptr, *dissY = nullptr;                  // [4][NpInt][Kp]: interior rows of Epsilon (.) Grad
    double *vn = nullptr;                // [2 sides][4][NpEdge][NEp]: owner-normal component of Epsilon (.) Grad on the edge
                                         // points, in the OWNER's point order (side 0 = owner, 1 = neighbour)
    double *aggv = nullptr;              // [NEp]
    double *DTVisc = nullptr;            // [Kp]
};

// Columns [K, Kp) of a [4][npInt][Kp] register get a benign constant state (rho=1, E=2.5, no momentum) so that the
// persistent element kernel can treat every 32-element tile as full; their metrics are zero, so they stay constant.
__global__ void k_fill_pad(double *q, int npInt, int K, int Kp) {
    const int pad = Kp - K;
    const long long total = (long long)4 * npInt * pad;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(t / pad), c = (int)(t % pad), v = r / npInt;
        q[(size_t)r * Kp + K + c] = (v == 0) ? 1.0 : (v == 3 ? 2.5 : 0.0);
    }
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


// ------------------------------------------------------------------------------------------------
// Phase P1: Persson modal sensor on the density of the stage input, sigma, element viscosity, and
// the element -> vertex max merge.  One thread per element.
//   UpdateSeMoment (DG2D/dfr_shock_capturing.go:142-166), UpdateShockFinderSigma
//   (dissipation.go:491-518), CalculateElementViscosity (:397-412),
//   MergeElementScalarToVertices with max (euler.go:1048-1065): every incident element contributes,
//   values are >= 0, so zero + atomicMax on the bit pattern gives the same vertex value.
// ------------------------------------------------------------------------------------------------
struct SensorArgs {
    int K, Kp;
    const double *rho;                 // [NpInt][Kp] density rows of the stage input
    const int *etov;                   // [3][Kp]
    const double *hk;
    double *se, *sigma, *epsk;
    unsigned long long *sigmaV, *epsV; // [NV] double bits
    DevScalars *sc;
    int par;
    long long stepIndex;
    Phys ph;
};

template <int N>
__global__ void __launch_bounds__(128) k_sensor(SensorArgs a) {
    constexpr int NI = Dim<N>::NpInt;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const Ops<N> &op = ops<N>();
    double rho[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) rho[i] = a.rho[(size_t)i * a.Kp + k];
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int i = 0; i < NI; i++) {
        double x = 0.0, y = 0.0, dq = 0.0;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            x = fma(op.P[i][j], rho[j], x);
            y = fma(op.M[i][j], rho[j], y);
            dq = fma(op.D[i][j], rho[j], dq);
        }
        num += dq * x;
        den += rho[i] * y;
    }
    const double se = log10(num / den);
    const double kappa = a.ph.sdKappa, S0 = a.ph.S0;
    const double left = S0 - kappa, right = S0 + kappa, ookappa = 0.5 / kappa;
    double sigma = 0.0;                 // (a NaN Se matches no case in the reference; unreachable for rho > 0)
    if (se < left) sigma = 0.0;
    else if (se >= left && se <= right) sigma = 0.5 * (1.0 + sin(3.14159265358979323846 * ookappa * (se - S0)));
    else if (se > right) sigma = 1.0;
    const double eps = a.ph.Eps0 * a.hk[k] * sigma;
    a.se[k] = se;
    a.sigma[k] = sigma;
    a.epsk[k] = eps;
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const int vid = a.etov[(size_t)v * a.Kp + k];
        if (sigma > 0.0) atomic_max_nonneg(&a.sigmaV[vid], sigma);
        if (eps > 0.0) atomic_max_nonneg(&a.epsV[vid], eps);
    }
}

// limitAndFilterSolution (dissipation.go:606-622) on one (element, variable) row held in registers
template <int N>
__device__ __forceinline__ void limit_filter_row(double (&u)[Dim<N>::NpInt], double sigmaK) {
    constexpr int NI = Dim<N>::NpInt;
    const Ops<N> &op = ops<N>();
    const double alpha = sin(0.5 * 3.14159265358979323846 * sigmaK);
    double uh[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NI; j++) s = fma(op.Vinv[i][j], u[j], s);
        uh[i] = (i >= 1) ? s * (op.mf[i] * (1.0 - alpha)) : s;
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NI; j++) s = fma(op.V[i][j], uh[j], s);
        u[i] = s;
    }
}

// Phase P3: sigma_k = mean of its three vertex values (euler.go:1067-1086); at rk == 2 the stage
// input is limited/filtered in place (euler.go:605-609); then the edge interpolation.
struct PrepArgs {
    int K, Kp;
    double *q;                         // stage input register (modified in place at rk == 2)
    double *qface;
    const int *etov;
    const double *sigmaV;
    double *sigma;
    DevScalars *sc;
    int rk, par;
    long long stepIndex;
    Phys ph;
};

template <int N>
__global__ void __launch_bounds__(kElemThreads) k_diss_prepare(PrepArgs a) {
    constexpr int NI = Dim<N>::NpInt;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    const int e = threadIdx.x % kElemsPerBlock, n = threadIdx.x / kElemsPerBlock;
    const int k = blockIdx.x * kElemsPerBlock + e;
    if (k >= a.K) return;
    double acc = 0.0;
#pragma unroll
    for (int v = 0; v < 3; v++) acc = acc + a.sigmaV[a.etov[(size_t)v * a.Kp + k]];
    const double sig = acc / 3.;
    if (n == 0) a.sigma[k] = sig;
    double qs[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) qs[i] = a.q[((size_t)n * NI + i) * a.Kp + k];
    if (a.rk == 2) {
        limit_filter_row<N>(qs, sig);
#pragma unroll
        for (int i = 0; i < NI; i++) a.q[((size_t)n * NI + i) * a.Kp + k] = qs[i];
    }
    interp_store<N>(qs, a.qface + (size_t)n * Dim<N>::NF3 * a.Kp, a.Kp, k);
}

// Epsilon at RT row r of an element from its three vertex values (InterpolateEpsilonSigma,
// dissipation.go:219-242): Bary[NpFlux x 3] . vertexVals[3]
template <int N>
__device__ __forceinline__ double eps_row(int r, double e0, double e1, double e2) {
    const Ops<N> &op = ops<N>();
    return op.Bary[r][0] * e0 + op.Bary[r][1] * e1 + op.Bary[r][2] * e2;
}

// Phase P5: RT-element gradient of every conserved variable, times Epsilon
//   GetSolutionGradientUsingRTElement (euler.go:864-918) + CalculateEpsilonGradient C0 branch
//   (dissipation.go:244-272).  Only the rows later consumed are produced: [0, NpInt) (AddDissipation)
//   and the 3*NpEdge edge rows (StoreEdgeViscousFlux); the duplicate interior rows are never read.
struct GradArgs {
    int K, Kp;
    const double *q;                   // stage input (after the rk == 2 limiter)
    const double *qface;               // Q_Face incl. neighbours (pre-BC = EdgeQValues of the owner)
    const int *etoe;                   // [3][Kp] edge slot (or -1-slot when not the owner)
    const int *ekL, *ekR, *emeta;      // edge table
    const double *Jdet, *Jinv, *IInII, *nxk, *nyk;
    const int *etov;
    const double *epsV;
    double *dissX, *dissY;             // [4][NpInt][Kp]: rows [0, NpInt) of DissX / DissY (AddDissipation reads them)
    // The 3 NpEdge edge rows of DissX / DissY have exactly one consumer, StoreEdgeViscousFlux, which only ever forms
    // normalL . (DissX, DissY) of both sides with the OWNER's normal (normalR := normalL, edges.go:175).  So the gradient
    // kernels store that one number per edge point straight into edge-indexed arrays, in the owner's point order:
    // half the bytes of the two row sets, and the viscous edge kernel reads them fully coalesced instead of gathering
    // 8-byte words from two element columns (its 25 % sector efficiency was item 3 of round 1's open list).
    const double *enx, *eny;           // [NEp] owner normal of every local edge
    double *vn;                        // [2][4][NpEdge][NEp]
    int NEp;
    DevScalars *sc;
    int par;
    long long stepIndex;
    Phys ph;
};

// row (>= 2 NpInt) of element k, local edge le, point ii, conserved variable n: store nx dX + ny dY into the edge slot
template <int N>
__device__ __forceinline__ void store_vn(const GradArgs &a, int n, int le, int ii, int s, double dX, double dY) {
    constexpr int NEd = Dim<N>::NpEdge;
    const bool owner = s >= 0;
    const int slot = owner ? s : -1 - s;
    const double v = a.enx[slot] * dX + a.eny[slot] * dY;
    a.vn[((size_t)((owner ? 0 : 4) + n) * NEd + (owner ? ii : NEd - 1 - ii)) * a.NEp + slot] = v;
}

template <int N>
__global__ void __launch_bounds__(kElemThreads) k_grad(GradArgs a) {
    constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3;
    constexpr int E = kElemsPerBlock;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    extern __shared__ double smem[];
    double *sU = smem;                 // [4][NI + NF3][E]: distinct solution values of the RT points
    const Ops<N> &op = ops<N>();
    const int e = threadIdx.x % E, n = threadIdx.x / E;
    const int k = blockIdx.x * E + e;
    const bool valid = k < a.K;
    const int kc = valid ? k : a.K - 1;
    const size_t Kp = a.Kp;
    constexpr int NU = NI + NF3;
    double *u = sU + (size_t)n * NU * E + e;          // u[j * E]
#pragma unroll
    for (int i = 0; i < NI; i++) u[i * E] = a.q[((size_t)n * NI + i) * Kp + kc];
    const size_t qplane = (size_t)NF3 * Kp;
    int sl[3];
#pragma unroll
    for (int le = 0; le < 3; le++) {
        const int s = a.etoe[(size_t)le * Kp + kc];
        sl[le] = s;
        if (s >= 0) {                   // owner: own edge values
#pragma unroll
            for (int i = 0; i < NEd; i++) u[(NI + le * NEd + i) * E] = a.qface[n * qplane + (size_t)(le * NEd + i) * Kp + kc];
        } else {                        // neighbour: the owner's values in reversed order (euler.go:896-912)
            const int slot = -1 - s;
            const int kO = a.ekL[slot];
            const int numO = a.emeta[slot] & 3;
#pragma unroll
            for (int i = 0; i < NEd; i++)
                u[(NI + le * NEd + i) * E] = a.qface[n * qplane + (size_t)(numO * NEd + (NEd - 1 - i)) * Kp + kO];
        }
    }
    // per-element metric scalars (CalculateRTBasedDerivativeMetrics, DG2D/dfr_startup.go:213-254)
    const double j0 = a.Jinv[0 * Kp + kc], j1 = a.Jinv[1 * Kp + kc], j2 = a.Jinv[2 * Kp + kc], j3 = a.Jinv[3 * Kp + kc];
    const double oojd = 1.0 / a.Jdet[kc];
    double mx[3], my[3];
#pragma unroll
    for (int le = 0; le < 3; le++) {
        const double iin = a.IInII[(size_t)le * Kp + kc];
        mx[le] = oojd * a.nxk[(size_t)le * Kp + kc] * iin;
        my[le] = oojd * a.nyk[(size_t)le * Kp + kc] * iin;
    }
    const double ev0 = a.epsV[a.etov[0 * Kp + kc]], ev1 = a.epsV[a.etov[1 * Kp + kc]], ev2 = a.epsV[a.etov[2 * Kp + kc]];
    __syncthreads();   // (each thread only reads what it wrote; the barrier keeps the phases aligned across warps)

    // output rows in chunks; DOF_j = metric_j * U_j is formed on the fly
    constexpr int NOUT = NI + NF3;                 // rows [0,NI) and [2NI, NF)
    constexpr int CH = (NOUT % 11 == 0) ? 11 : ((NOUT % 9 == 0) ? 9 : ((NOUT % 7 == 0) ? 7 : ((NOUT % 5 == 0) ? 5 : 3)));
    static_assert(NOUT % CH == 0, "chunking");
#pragma unroll
    for (int c0 = 0; c0 < NOUT; c0 += CH) {
        double gx[CH], gy[CH];
#pragma unroll
        for (int r = 0; r < CH; r++) { gx[r] = 0.0; gy[r] = 0.0; }
#pragma unroll
        for (int j = 0; j < NF; j++) {
            // value and metric of RT point j
            const int ju = (j < NI) ? j : ((j < 2 * NI) ? j - NI : j - NI);
            const double uj = u[ju * E];
            double dx, dy;
            if (j < NI) { dx = j0 * uj; dy = j1 * uj; }
            else if (j < 2 * NI) { dx = j2 * uj; dy = j3 * uj; }
            else { const int le = (j - 2 * NI) / NEd; dx = mx[le] * uj; dy = my[le] * uj; }
#pragma unroll
            for (int r = 0; r < CH; r++) {
                const int row = (c0 + r < NI) ? (c0 + r) : (c0 + r + NI);
                const double d = op.Div[row][j];
                gx[r] = fma(d, dx, gx[r]);
                gy[r] = fma(d, dy, gy[r]);
            }
        }
        if (valid) {
#pragma unroll
            for (int r = 0; r < CH; r++) {
                const int row = (c0 + r < NI) ? (c0 + r) : (c0 + r + NI);
                const double eps = op.Bary[row][0] * ev0 + op.Bary[row][1] * ev1 + op.Bary[row][2] * ev2;
                if (row < NI) {
                    a.dissX[((size_t)n * NI + row) * Kp + k] = gx[r] * eps;
                    a.dissY[((size_t)n * NI + row) * Kp + k] = gy[r] * eps;
                } else {
                    const int le = (row - 2 * NI) / NEd, ii = (row - 2 * NI) % NEd;
                    store_vn<N>(a, n, le, ii, sl[le], gx[r] * eps, gy[r] * eps);
                }
            }
        }
    }
}

// Phase P6: StoreEdgeViscousFlux (edges.go:151-244) + the viscous half of StoreEdgeAggregates
// (edges.go:274-286).  One thread per edge, grid-stride.
struct ViscEdgeArgs {
    int ne, NEp, Kp;
    const int *kL, *kR, *meta;
    const double *nx, *ny, *oohk, *ooLen;
    const double *qface, *vn;          // vn: [2][4][NpEdge][NEp], see GradArgs
    const int *etov;
    const double *epsV;
    double *eflux, *aggv;              // eflux: numerical normal flux, the viscous one is subtracted in place
    DevScalars *sc;
    int slot, par;
    long long stepIndex;
    Phys ph;
};

// Thread = (edge, conserved variable): block (64 edges, 4 variables), every warp walks 32 consecutive edges of one variable,
// so all vn / eflux accesses are full 256-byte rows (one thread per edge with 50+ dependent loads was latency bound:
// 45 % DRAM throughput, 38 long-scoreboard stall cycles per instruction, profiles/r02i_*).  The viscous normal flux is
// SUBTRACTED FROM THE STORED NUMERICAL FLUX in place: its only consumer, the element kernel, needs F - F_visc on the edge
// DOFs (AddDissipation, dissipation.go:316-333, folded into the DivInt contraction), so it gathers one array, not two.
constexpr int kViscEdges = 64;

template <int N>
__global__ void __launch_bounds__(4 * kViscEdges) k_visc_edge(ViscEdgeArgs a) {
    constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF3 = Dim<N>::NF3;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    const size_t Kp = a.Kp;
    const size_t qplane = (size_t)NF3 * Kp, fplane = (size_t)NEd * a.NEp;
    const int n = threadIdx.y;
    double blockmax = 0.0;
    for (int e = blockIdx.x * kViscEdges + threadIdx.x; e < a.ne; e += gridDim.x * kViscEdges) {
        const int kL = a.kL[e], kRraw = a.kR[e], meta = a.meta[e];
        const bool shared = kRraw >= 0;
        const int kR = shared ? kRraw : 0;
        const int numL = meta & 3, numR = (meta >> 2) & 3;
        const double oohk = a.oohk[e];
        const double eL0 = a.epsV[a.etov[0 * Kp + kL]], eL1 = a.epsV[a.etov[1 * Kp + kL]], eL2 = a.epsV[a.etov[2 * Kp + kL]];
        double eR0 = 0, eR1 = 0, eR2 = 0;
        if (shared) { eR0 = a.epsV[a.etov[0 * Kp + kR]]; eR1 = a.epsV[a.etov[1 * Kp + kR]]; eR2 = a.epsV[a.etov[2 * Kp + kR]]; }
        const double ooLen = a.ooLen[e];
        double vmax = -1.7976931348623157e308;
        const double *qrow = a.qface + n * qplane + (size_t)(numL * NEd) * Kp + kL;
        double qv[NEd];
#pragma unroll
        for (int i = 0; i < NEd; i++) qv[i] = shared ? qrow[(size_t)i * Kp] : 0.0;
#pragma unroll
        for (int i = 0; i < NEd; i++) {
            const int rowL = 2 * NI + numL * NEd + i;
            const int rowR = 2 * NI + numR * NEd + (NEd - 1 - i);
            const Ops<N> &op = ops<N>();
            const double epsL = op.Bary[rowL][0] * eL0 + op.Bary[rowL][1] * eL1 + op.Bary[rowL][2] * eL2;
            vmax = fmax(oohk * oohk * epsL, vmax);
            // n . (DissX, DissY) of both sides with the owner's normal (normalR := normalL in the reference,
            // edges.go:175), formed by the gradient kernel
            const size_t idx = ((size_t)n * NEd + i) * a.NEp + e;
            const double vFL = a.vn[idx];
            double vf = vFL;
            if (shared) {
                const double epsR = op.Bary[rowR][0] * eR0 + op.Bary[rowR][1] * eR1 + op.Bary[rowR][2] * eR2;
                const double lam = 0.5 * (epsL + epsR);
                const double vFR = a.vn[idx + 4 * fplane];
                vf = 0.5 * (vFL + vFR);
                // both "sides" of the jump resolve to the owner's stored edge values (edges.go:225-236)
                vf -= (a.ph.Omega * lam * ooLen) * (qv[i] - qv[NEd - 1 - i]);
            }
            a.eflux[n * fplane + (size_t)i * a.NEp + e] -= vf;
        }
        if (n == 0) {
            a.aggv[e] = vmax;
            blockmax = fmax(blockmax, vmax);
        }
    }
    __shared__ double smax[8];
    const int tid = threadIdx.y * kViscEdges + threadIdx.x;
    blockmax = warp_max(blockmax);
    if ((tid & 31) == 0) smax[tid >> 5] = blockmax;
    __syncthreads();
    if (tid < 32) {
        double v = tid < 8 ? smax[tid] : 0.0;
        v = warp_max(v);
        if (tid == 0) atomic_max_nonneg(&a.sc->wave[a.slot][1], v);
    }
}

// ------------------------------------------------------------------------------------------------
// Field read-back: Euler.GetPlotField for the GetFlowFunction family (plot.go:14-86) --
// flow function per solution node (fluids.go:289-336), GraphInterp product (DG2D/dfr_startup.go:62-63),
// AverageGraphFieldVertices (DG2D/graphics_support2.go:184-199), transpose to [element][graph node] and the
// float32 narrowing of the AVS writer (DG2D/graphics_support.go:80-93), so that a plot costs one pass over the
// state and NpGraph floats per element over PCIe instead of four float64 registers.
// ------------------------------------------------------------------------------------------------
struct PlotArgs {
    int K, Kp, ff;
    const double *q;       // [4][NpInt][Kp]
    const double *gi;      // [NpGraph][NpInt] row-major (device copy of DFR.GraphInterp)
    float *out;            // [K][NpGraph]
    double gamma, Pinf, QQinf;
    double kappa;          // ShockFunction: sf.Kappa of the finder GetPlotField uses (plot.go:31-37)
};

// ModeAliasShockFinder.ShockIndicator (DG2D/dfr_shock_capturing.go:189-235) on one element's density: the L2 moment of
// U - Clipper U with the diagonal of the mass matrix, Persson's ramp with S0 = 4 / N^4 (N, not N+1: this is the plot
// path's indicator, not UpdateShockFinderSigma).  Clipper = I - D (the resident operator set carries D).
template <int N>
__device__ __forceinline__ double shock_indicator(const double (&u)[Dim<N>::NpInt], double kappa) {
    constexpr int NI = Dim<N>::NpInt;
    const Ops<N> &op = ops<N>();
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int i = 0; i < NI; i++) {
        double clipped = 0.0;
#pragma unroll
        for (int j = 0; j < NI; j++) clipped = fma(((i == j) ? 1.0 : 0.0) - op.D[i][j], u[j], clipped);
        const double t1 = u[i] - clipped, mass = op.M[i][i];
        num += mass * (t1 * t1);
        den += mass * (u[i] * u[i]);
    }
    const double Se = log10(num / den);
    const double S0 = 4.0 / pow((double)N, 4.0);
    const double left = S0 - kappa, right = S0 + kappa;
    if (Se < left) return 0.0;
    if (Se <= right) return 0.5 * (1.0 + sin(3.14159265358979323846 * (0.5 / kappa) * (Se - S0)));
    if (Se > right) return 1.0;
    return 0.0;                                    // NaN: none of the reference's switch cases fires
}

__device__ __forceinline__ double flow_function(int pf, double gamma, double Pinf, double QQinf, double rho, double rhoU,
                                                double rhoV, double E) {
    const double GM1 = gamma - 1.0, oorho = 1.0 / rho;
    switch (pf) {
        case 0: return rho;
        case 1: return rhoU;
        case 2: return rhoV;
        case 3: return E;
        case 10: return rhoU * oorho;
        case 11: return rhoV * oorho;
        default: break;
    }
    const double u = rhoU * oorho, v = rhoV * oorho;
    const double U2 = u * u + v * v;
    const double q = 0.5 * rho * U2;
    const double p = GM1 * (E - q);
    switch (pf) {
        case 9: return sqrt(U2);
        case 6: return q;
        case 5: return p;
        case 7: return (p - Pinf) / QQinf;
        case 8: return sqrt(fabs(gamma * p * oorho));
        case 12: return (E + p) / rho;
        case 13: return log(p) - gamma * log(rho);
        case 4: return sqrt(U2) / sqrt(fabs(gamma * p * oorho));
        default: return 0.0;
    }
}

constexpr int kPlotThreads = 128;

template <int N>
__global__ void __launch_bounds__(kPlotThreads) k_plot_field(PlotArgs a) {
    constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NG = 3 * (1 + NEd) + NI, ST = NG | 1;   // odd smem stride
    __shared__ float sOut[kPlotThreads * ST];
    const int kb = blockIdx.x * kPlotThreads, k = kb + threadIdx.x;
    if (k < a.K) {
        double f[NI];
        if (a.ff == 100) {                          // ShockFunction: the element's indicator at every node (plot.go:30-47)
#pragma unroll
            for (int i = 0; i < NI; i++) f[i] = a.q[(size_t)i * a.Kp + k];
            const double m = shock_indicator<N>(f, a.kappa);
#pragma unroll
            for (int i = 0; i < NI; i++) f[i] = m;
        } else {
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const size_t o = (size_t)i * a.Kp + k;
                const size_t plane = (size_t)NI * a.Kp;
                f[i] = flow_function(a.ff, a.gamma, a.Pinf, a.QQinf, a.q[o], a.q[plane + o], a.q[2 * plane + o], a.q[3 * plane + o]);
            }
        }
        double g[NG];
#pragma unroll
        for (int r = 0; r < NG; r++) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NI; j++) s = fma(a.gi[r * NI + j], f[j], s);
            g[r] = s;
        }
        // the three vertex nodes take the mean of their two neighbours along the element boundary
        constexpr int NPE = NEd + 2;
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const int iV = e * (NPE - 1), iVp = iV + 1, iVm = (e == 0) ? 3 * (NPE - 1) - 1 : e * (NPE - 1) - 1;
            g[iV] = 0.5 * (g[iVp] + g[iVm]);
        }
#pragma unroll
        for (int r = 0; r < NG; r++) sOut[threadIdx.x * ST + r] = (float)g[r];
    }
    __syncthreads();
    const int nk = min(kPlotThreads, a.K - kb);
    float *dst = a.out + (size_t)kb * NG;
    for (int t = threadIdx.x; t < nk * NG; t += kPlotThreads) dst[t] = sOut[(t / NG) * ST + (t % NG)];
}

// ------------------------------------------------------------------------------------------------
// Initial condition on the device: InitializeSolution (euler.go:728-794) -- InitializeFS and
// InitializeIVortex (initialization.go:50-83), shock-tube split at x < 0.5 (euler.go:742-768) -- evaluated at the
// solution points X = 0.5 (-(r+s) v0 + (1+r) v1 + (1+s) v2) (CalculateElementLocalGeometry), so that the host neither
// builds nor uploads the [4][NpInt x K] state.  One thread per element.
// ------------------------------------------------------------------------------------------------
struct InitArgs {
    int K, Kp, npInt, initCase;
    const double *vx, *vy;     // [NV]
    const int *etov;           // [K][3] own elements, global vertex ids
    const double *r, *s;       // [NpInt] solution points of the reference element
    double *q;                 // [4][NpInt][Kp]
    Phys ph;
};

__global__ void k_init_state(InitArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const int v0 = a.etov[3 * (size_t)k], v1 = a.etov[3 * (size_t)k + 1], v2 = a.etov[3 * (size_t)k + 2];
    const double ax = a.vx[v0], bx = a.vx[v1], cx = a.vx[v2], ay = a.vy[v0], by = a.vy[v1], cy = a.vy[v2];
    const size_t plane = (size_t)a.npInt * a.Kp;
    for (int i = 0; i < a.npInt; i++) {
        const double r = a.r[i], s = a.s[i];
        const double x = (((r + s) * -1.0) * ax + (r + 1.0) * bx + (s + 1.0) * cx) * 0.5;
        const double y = (((r + s) * -1.0) * ay + (r + 1.0) * by + (s + 1.0) * cy) * 0.5;
        double Q[4];
        if (a.initCase == DFR2D_CASE_IVortex) {
            ivortex_state(a.ph.vortex, 0.0, x, y, Q);
        } else {
            const dfr2d_freestream &f = (a.initCase == DFR2D_CASE_ShockTube) ? (x < 0.5 ? a.ph.fs[1] : a.ph.fs[2]) : a.ph.fs[0];
            Q[0] = f.Qinf[0]; Q[1] = f.Qinf[1]; Q[2] = f.Qinf[2]; Q[3] = f.Qinf[3];
        }
#pragma unroll
        for (int n = 0; n < 4; n++) a.q[n * plane + (size_t)i * a.Kp + k] = Q[n];
    }
}

// EpsilonDissipation / EpsilonDissipationC0 plot fields (plot.go:48-53; dissipation.go:377-395): [NpFlux][K] doubles, not
// interpolated -- the element's scalar epsilon at every RT point, or Bary . (vertex epsilon) as InterpolateEpsilonSigma
// left it (dissipation.go:219-242).
template <int N>
__global__ void k_epsilon_field(int K, int Kp, int c0, const double *epsk, const double *epsV, const int *etov, double *out) {
    constexpr int NF = Dim<N>::NpFlux;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    if (!c0) {
        const double e = epsk[k];
#pragma unroll
        for (int i = 0; i < NF; i++) out[(size_t)i * K + k] = e;
    } else {
        const double e0 = epsV[etov[k]], e1 = epsV[etov[(size_t)Kp + k]], e2 = epsV[etov[2 * (size_t)Kp + k]];
#pragma unroll
        for (int i = 0; i < NF; i++) out[(size_t)i * K + k] = eps_row<N>(i, e0, e1, e2);
    }
}

// ---- gradient plot fields (plot.go:54-77): XGradient* / YGradient* = GetSolutionGradientUsingRTElement(-1, n, c.Q, ...) ----
// The reference evaluates them from the CURRENT c.Q at the interior RT points and from the EdgeQValues store at the edge
// points -- the owner side's Q_Face of the last CalculateEdgeEulerFlux, i.e. of the input of stage 5 of the last step.
// Kernel 5 of that stage overwrites Q_Face with the next stage's interpolation, so a host that wants these fields turns
// the capture on (dfr2d_capture_edge_values) and this copy runs between the edge phase and the update of every stage 5.
// It honours the step predicate: a step issued after the run has finished must leave the store alone.
__global__ void __launch_bounds__(256) k_capture_qface(size_t n2, const double2 *src, double2 *dst, const DevScalars *sc, Phys ph, int par,
                                                       long long stepIndex) {
    if (step_is_noop(sc, ph, par, stepIndex)) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

struct GradPlotArgs {
    int K, Kp, var, dirY;              // dirY = 0: GradX (DXMetric), 1: GradY (DYMetric)
    const double *q;                   // c.Q
    const double *qfaceSaved;          // [4][3NpEdge][Kp] captured Q_Face (own + ghost columns)
    const int *etoe, *ekL, *emeta;
    const double *Jdet, *Jinv, *IInII, *nk;   // nk = the element normals' x (dirY = 0) or y (dirY = 1) component, [3][Kp]
    double *grad;                      // [NpFlux][K]
};

// One thread per element and all NpFlux rows of the requested direction: DOF_j = metric_j U_j
// (CalculateRTBasedDerivativeMetrics, DG2D/dfr_startup.go:213-254), Grad = Div . DOF (raviart_thomas_element.go:249-297).
// A read-back utility that runs once per plotted field, not a stage kernel: the operator comes straight from constant
// memory (warp-uniform index).
template <int N>
__global__ void __launch_bounds__(128) k_grad_plot(GradPlotArgs a) {
    constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const Ops<N> &op = ops<N>();
    const size_t Kp = a.Kp;
    const int n = a.var;
    double dof[NF];
    // DXMetric rows: {rx, sx} = Jinv[0], Jinv[2]; DYMetric rows: {ry, sy} = Jinv[1], Jinv[3]
    const double ja = a.Jinv[(size_t)(a.dirY ? 1 : 0) * Kp + k], jb = a.Jinv[(size_t)(a.dirY ? 3 : 2) * Kp + k];
#pragma unroll
    for (int i = 0; i < NI; i++) {
        const double u = a.q[((size_t)n * NI + i) * Kp + k];
        dof[i] = ja * u;
        dof[NI + i] = jb * u;
    }
    const double oojd = 1.0 / a.Jdet[k];
    const size_t qplane = (size_t)NF3 * Kp;
#pragma unroll
    for (int le = 0; le < 3; le++) {
        const double m = oojd * a.nk[(size_t)le * Kp + k] * a.IInII[(size_t)le * Kp + k];
        const int s = a.etoe[(size_t)le * Kp + k];
        int col = k, row0 = le * NEd, dir = 1;
        if (s < 0) {                    // neighbour: the owner's values in reversed order (euler.go:896-912)
            const int slot = -1 - s;
            col = a.ekL[slot];
            row0 = (a.emeta[slot] & 3) * NEd + NEd - 1;
            dir = -1;
        }
#pragma unroll
        for (int i = 0; i < NEd; i++)
            dof[2 * NI + le * NEd + i] = m * a.qfaceSaved[n * qplane + (size_t)(row0 + dir * i) * Kp + col];
    }
#pragma unroll 1
    for (int r = 0; r < NF; r++) {
        double g = 0.0;
#pragma unroll
        for (int j = 0; j < NF; j++) g = fma(op.Div[r][j], dof[j], g);
        a.grad[(size_t)r * a.K + k] = g;
    }
}

template <int N> static size_t elem_smem_diss() { return (size_t)12 * Dim<N>::NpInt * kElemsPerBlock * sizeof(double); }

}  // namespace dfr2d
