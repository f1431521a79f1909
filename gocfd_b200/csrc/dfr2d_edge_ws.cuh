// k_edge_ws<N, FLUX>: the interior-edge kernel (k_edge_int: calculateSharedEdgeFlux edges.go:379-411 + StoreEdgeAggregates
// :246-289) as a warp-specialised producer / consumer pipeline, like kernel 5 and k_grad_ws.
//
// k_edge_int is a latency-bound gather (ncu, profiles/r02zz_ncu_full_c5.txt: long_scoreboard 10.6 of 16 stall cycles per
// issued instruction, 24 warps per SM at 80 registers, HBM at 62 % of its peak, FP64 pipe 44 % busy): every thread loads the
// eight Q_Face values of an edge point, waits the full memory latency, then runs ~170 FP64 instructions of Roe flux; the
// registers that flux needs are the ones more loads in flight would need.  Both floors of the kernel are ~1.1 ms at C5
// (7.4 GB of HBM traffic; 170 x 0.83 pipe cycles per warp-point), it runs at 1.49 ms.  Here
//   * four producer warps (lane = edge of the tile) own the index chain (edge -> owner / neighbour column and edge
//     number, loaded one tile ahead) and gather the 8 NpEdge values of every edge of the tile with 8-byte cp.async into a
//     ring of shared-memory stages ([value row][edge], completing on the stage's FULL mbarrier) -- three to five tiles in
//     flight per SM, none of it in registers;
//   * the consumer warps (thread = edge x pair of points) read their operands from shared memory, hand the stage back
//     (EMPTY) before the arithmetic starts, evaluate the flux with the same device functions as k_edge_int (bitwise the
//     same results) and store coalesced rows of the edge-indexed flux array.
// Only for even NpEdge (N = 0, 2, 4: two points per thread); the other orders, the VISC form and the boundary / cut-edge
// list keep k_edge_int / k_edge.
#pragma once
#include "dfr2d_kernels.cuh"
#include "dfr2d_elem_ws.cuh"

namespace dfr2d {

template <int N> struct EdgeWsDim {
    static constexpr int NEd = Dim<N>::NpEdge;
    static constexpr int TE = 128;                       // edges per tile
    static constexpr int PT = 2;                         // points per consumer thread
    static constexpr int G = NEd / PT;                   // consumer threads per edge
    static constexpr int kProdWarps = 4, kConsWarps = G * (TE / 32);
    static constexpr int kThreads = (kProdWarps + kConsWarps) * 32;
    static constexpr int kRows = 8 * NEd + 4;            // value rows (point, side, variable) + nx, ny, 1/hK, valid flag
    static constexpr int kStageDoubles = kRows * TE;
    static constexpr int kFullCount = kProdWarps * 64;   // per producer lane: its cp.async completions + one plain arrive
    static constexpr int kMaxStages = 6;
    static_assert(NEd % PT == 0, "k_edge_ws: even NpEdge only");
    static int stages() {                                // as many as fit beside ~64 KB of L1 for the gather, 3..6
        const int fit = (int)((232448 - 512 - 65536) / (kStageDoubles * sizeof(double)));
        return fit < 3 ? 3 : (fit > kMaxStages ? kMaxStages : fit);
    }
};

template <int N, int FLUX>
__global__ void __launch_bounds__(EdgeWsDim<N>::kThreads, 1) k_edge_ws(EdgeArgs a, int nTiles, int S) {
    using ED = EdgeWsDim<N>;
    constexpr int NEd = ED::NEd, TE = ED::TE, PT = ED::PT, G = ED::G;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) return;
    extern __shared__ __align__(16) double smem_e[];
    __shared__ __align__(8) unsigned long long fullBar[ED::kMaxStages], emptyBar[ED::kMaxStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&fullBar[s], ED::kFullCount);
            mbar_init(&emptyBar[s], ED::kConsWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nLocal = (nTiles > (int)blockIdx.x) ? (nTiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const size_t qplane = (size_t)Dim<N>::NF3 * a.Kp;

    if (warp < ED::kProdWarps) {
        // =================================== producers: lane = edge 32 warp + lane of the tile ==========================
        const int el = 32 * warp + lane;
        int kL = 0, kR = -1, meta = 0;
        auto load_idx = [&](int n) {
            kR = -1;
            if (n >= nLocal) return;
            const long long e = (long long)(blockIdx.x + (long long)n * gridDim.x) * TE + el;
            if (e < a.ne) { kL = a.kL[e]; kR = a.kR[e]; meta = a.meta[e]; }
        };
        load_idx(0);
        int s = 0;
        unsigned ph = 1;                        // parity of the previous phase of emptyBar[s]
        for (int n = 0; n < nLocal; n++) {
            if (n >= S) mbar_wait(&emptyBar[s], ph);
            const long long e = (long long)(blockIdx.x + (long long)n * gridDim.x) * TE + el;
            double *st = smem_e + (size_t)s * ED::kStageDoubles;
            // interior edge of this partition: both columns own (a ghost column makes it a cut edge: list pass)
            const bool valid = kR >= 0 && kR < a.Kown && kL < a.Kown;
            st[(8 * NEd + 3) * TE + el] = valid ? 1.0 : 0.0;
            if (valid) {
                const unsigned stU = smem_u32(st) + (unsigned)(el * sizeof(double));
                const int numL = meta & 3, numR = (meta >> 2) & 3;
                const double *pL = a.qface + (size_t)(numL * NEd) * a.Kp + kL;
                const double *pR = a.qface + (size_t)(numR * NEd + NEd - 1) * a.Kp + kR;        // points reversed (edges.go:396-400)
#pragma unroll
                for (int i = 0; i < NEd; i++) {
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        cp_async8_u32(stU + (unsigned)(((i * 8 + v) * TE) * sizeof(double)), pL + v * qplane + (size_t)i * a.Kp);
                        cp_async8_u32(stU + (unsigned)(((i * 8 + 4 + v) * TE) * sizeof(double)), pR + v * qplane - (size_t)i * a.Kp);
                    }
                }
                cp_async8_u32(stU + (unsigned)(((8 * NEd + 0) * TE) * sizeof(double)), a.nx + e);
                cp_async8_u32(stU + (unsigned)(((8 * NEd + 1) * TE) * sizeof(double)), a.ny + e);
                cp_async8_u32(stU + (unsigned)(((8 * NEd + 2) * TE) * sizeof(double)), a.oohk + e);
            }
            cp_async_arrive_noinc(&fullBar[s]);
            mbar_arrive(&fullBar[s]);
            load_idx(n + 1);
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else {
        // =================================== consumers: thread = (edge, pair of points) =================================
        const int cw = warp - ED::kProdWarps;
        const int g = cw / (TE / 32), el = 32 * (cw % (TE / 32)) + lane;
        const double gamma = a.ph.gamma;
        const size_t fplane = (size_t)NEd * a.NEp;
        double blockmax = 0.0;
        int s = 0;
        unsigned ph = 0;
        for (int n = 0; n < nLocal; n++) {
            const long long e = (long long)(blockIdx.x + (long long)n * gridDim.x) * TE + el;
            const double *st = smem_e + (size_t)s * ED::kStageDoubles + el;
            mbar_wait(&fullBar[s], ph);
            const bool valid = st[(8 * NEd + 3) * TE] != 0.0;
            double QL[PT][4], QR[PT][4], nx = 0.0, ny = 0.0, oohk = 0.0;
            if (valid) {
#pragma unroll
                for (int ii = 0; ii < PT; ii++) {
                    const int i = g * PT + ii;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        QL[ii][v] = st[(i * 8 + v) * TE];
                        QR[ii][v] = st[(i * 8 + 4 + v) * TE];
                    }
                }
                nx = st[(8 * NEd + 0) * TE]; ny = st[(8 * NEd + 1) * TE]; oohk = st[(8 * NEd + 2) * TE];
            }
            // operands are in registers: the stage goes back to the producers before the arithmetic
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyBar[s]);
            if (valid) {
                double wmax = -1.7976931348623157e308;
#pragma unroll
                for (int ii = 0; ii < PT; ii++) {
                    const int i = g * PT + ii;
                    double F[4], wL;
                    if (FLUX == DFR2D_FLUX_Roe) roe_flux(gamma, QL[ii], QR[ii], nx, ny, F, wL);
                    else {
                        if (FLUX == DFR2D_FLUX_Average) avg_flux(gamma, QL[ii], QR[ii], nx, ny, F);
                        else if (FLUX == DFR2D_FLUX_LaxFriedrichs) lax_flux(gamma, QL[ii], QR[ii], nx, ny, F);
                        else roe_er_flux(gamma, QL[ii], QR[ii], nx, ny, F);
                        wL = speed_plus_sound(gamma, QL[ii][0], QL[ii][1], QL[ii][2], QL[ii][3]);
                    }
#pragma unroll
                    for (int v = 0; v < 4; v++) a.eflux[v * fplane + (size_t)i * a.NEp + e] = F[v];
                    const double w = oohk * wL;
                    if (w > wmax) wmax = w;
                }
                if (G == 1) a.agg[e] = wmax;
                else if (a.ph.localDT) atomic_max_nonneg(reinterpret_cast<unsigned long long *>(&a.agg[e]), wmax);
                blockmax = fmax(blockmax, wmax);
            }
            if (++s == S) { s = 0; ph ^= 1u; }
        }
        blockmax = warp_max(blockmax);
        if (lane == 0 && blockmax > 0.0) atomic_max_nonneg(&a.sc->wave[a.slot][0], blockmax);
    }
}

}  // namespace dfr2d
