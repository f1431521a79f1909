// k_elem_ws<N>: the inviscid element kernel as a Warp-Specialised producer/consumer pipeline (kernel #5).  (Round 1 called
// it k_elem_tma; it never used the TMA engine -- the data moves by per-thread cp.async that complete on mbarriers -- so
// round 2 renamed it.  Profiles r01* / r02a-j still carry the old name.)
//
// What the measurements of kernel #4 (k_elem_pipe) said (profiles/r01e_*): per 32-element tile an SM spent
// ~7,700 cycles = (HBM time of the tile's 97 KB at the SM's bandwidth share, ~4,400) + (DMMA floor, ~2,150) + (flux and
// epilogue ALU work) -- the parts ADD because every warp does all of them in lock-step phases separated by CTA
// barriers, so when HBM back-pressures the LSU queue the tensor pipe idles and vice versa.  Here they overlap:
//
//   * FOUR producer warps per CTA move every byte asynchronously: the [row][32 elements] slabs of the stage input, of
//     the RK registers and of the metrics arrive by coalesced 16-byte cp.async (two 256-byte rows per warp instruction,
//     shared-memory rows at a stride of 36 doubles), the edge-flux gather by 8-byte cp.async; all complete on the
//     stage's mbarrier (cp.async.mbarrier.arrive.noinc).  No load ever occupies a register or an issue slot of a warp
//     that does arithmetic.  (One cp.async.bulk -- the TMA engine -- per 256-byte row was measured first: 13 ms, a bulk
//     copy is a uniform-datapath instruction and lane-dependent addresses serialise.)
//   * 8 consumer warps work in two groups of 4 on two different tiles; inside a group every warp owns 8 columns
//     (= one DMMA n-tile) of the tile for ALL four conserved variables and never synchronises with another warp:
//     no __syncthreads in the steady state, phases of different warps drift apart and fill each other's bubbles.
//   * The physical flux is evaluated directly in B-fragment layout: lane (fr, fc) of the m8n8k4 DMMA holds B[k = fc][n = fr],
//     so it evaluates point p = 4m + fc of element fr once and feeds Fr to k-step 2m and Fs to k-step 2m+1 for all four
//     variables (the k-order of the contraction is permuted accordingly in the operator fragments).  F_RT_DOF never
//     exists in shared memory; the operator fragments (A operands) stay in registers for the whole kernel.
//   * The fresh register is written in place into the warp's own columns of the stage-input slab, which then serves as
//     the B operand of the fused FluxEdgeInterp contraction (next stage's Q_Face).
//
// Reference semantics are those of k_elem_pipe / k_elem (SetRTFluxInternal euler.go:701-726, SetRTFluxOnEdges
// edges.go:454-483, RHSInternalPoints euler.go:665-699, rkAdvance :502-565, InterpolateSolutionToEdges edges.go:485-491).
#pragma once
#include "dfr2d_elem_mma.cuh"

namespace dfr2d {

constexpr int kWsProdWarps = 4;                         // warp 0: row slabs + dt; warps 1-3: gather of local edge 0-2
constexpr int kWsFullCount = kWsProdWarps * 64;        // per producer lane: its cp.async completions + one plain arrive
constexpr int kWsMaxStages = 6;

template <int N> struct WsDim {
    static constexpr int NI = Dim<N>::NpInt, NEd = Dim<N>::NpEdge, NF3 = Dim<N>::NF3;
    static constexpr int SE = 36;                                  // row stride of the slabs (doubles): 4 k-rows x 8 n conflict free
    static constexpr int M1 = (NI + 7) / 8, KI = (NI + 3) / 4;
    // Tensor-pipe work is co-limiting with HBM here (ncu: shared FP64/DMMA pipe 65 % busy, profiles/r02c_*), so the zero
    // padding of the 8 x 8 x 4 tiles is cut where it is cheap:
    //  * the interior k-steps have 4 KI - NI idle point slots each for Fr and for Fs; the LAST nMoved edge rows ride in
    //    them (lane fc of k-step 2m / 2m+1 carries edge row NE3k + 2 (4m + fc - NI) [+ 1] instead of a zero), which
    //    shortens the edge block from ceil(NF3 / 4) to KE k-steps (N=4: 13 -> 12 k-steps of C1);
    //  * the interpolation operator has NF3 rows: when NF3 mod 8 is 1..4 the ragged last m-tile (N=4: rows 16, 17 of 24) is
    //    evaluated with plain DFMA on constant-bank operands instead of a mostly empty DMMA tile (48 -> 32 DMMA of C2).
    static constexpr int padI = 4 * KI - NI;
    static constexpr int nMoved = (2 * padI < NF3) ? 2 * padI : NF3;
    static constexpr int NE3k = NF3 - nMoved;                      // edge rows that stay in the edge k-steps
    static constexpr int KE = (NE3k + 3) / 4, K1 = 2 * KI + KE;
    static constexpr int remF = NF3 % 8;
    static constexpr bool tailDfma = remF >= 1 && remF <= 4;
    static constexpr int M2 = tailDfma ? NF3 / 8 : (NF3 + 7) / 8, NTail = tailDfma ? remF : 0, K2 = (NI + 3) / 4;
    static constexpr int qOff = 0;                                 // [4 NI][SE] stage input
    static constexpr int eOff = 4 * NI * SE;                       // [4 NF3][SE] gathered numerical edge flux (raw)
    static constexpr int gOff = eOff + 4 * NF3 * SE;               // [9][32]: Jdet, Jinv0..3, (+-)IInII0..2, dt
    static constexpr int xOff = gOff + 9 * 32;                     // [nExtra][4 NI][SE]: q0 (, q2, q3, R)
    static constexpr int xSize = 4 * NI * SE;
    static int stage_doubles(int nExtra) { return xOff + nExtra * xSize; }
    static size_t smem_bytes(int nExtra, int stages) { return (size_t)stage_doubles(nExtra) * stages * sizeof(double); }
    // PerssonC0 variant (DISS): one more metric row (1 - sin(pi sigma / 2) of the limiter) and two more slabs, the interior
    // rows of DissX and DissY.  At rk 4 only q0 and q2 are staged; q3 and R are read from global memory in the epilogue
    // (four extra slabs would not leave room for a second ring stage).
    static constexpr int xOffD = gOff + 10 * 32;
    __host__ __device__ static int extras_in_smem(int nExtra, bool diss) { return (diss && nExtra == 4) ? 2 : nExtra; }
    __host__ __device__ static int stage_doubles_diss(int nExtraS) { return xOffD + (nExtraS + 2) * xSize; }
    static size_t smem_bytes_diss(int nExtraS, int stages) { return (size_t)stage_doubles_diss(nExtraS) * stages * sizeof(double); }
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void cp_async16_u32(unsigned dst, const double *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// one [rows][32] slab of a [.][Kp] array -> shared rows of stride SE: every warp instruction moves two full 256-byte rows
template <int ROWS, int SE_>
__device__ __forceinline__ void slab_g2s(unsigned dstBase, const double *srcBase, size_t Kp, int lane) {
    const int half = lane >> 4, ch = lane & 15;
    const double *src = srcBase + (size_t)half * Kp + 2 * ch;
    unsigned d = dstBase + (unsigned)((half * SE_ + 2 * ch) * sizeof(double));
#pragma unroll 5
    for (int r = 0; r < ROWS; r += 2) {
        if (ROWS % 2 == 0 || r + half < ROWS) cp_async16_u32(d, src);
        src += 2 * Kp;
        d += (unsigned)(2 * SE_ * sizeof(double));
    }
}
__device__ __forceinline__ void cp_async8_u32(unsigned dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct ElemWsArgs {
    ElemArgs a;
    int nTiles, nStages, nExtra;
    // (r2) measured variant DFR2D_WS_TMA=1: the extra RK-register slabs (q0, q2, q3, R) are fetched by the TMA engine --
    // one cp.async.bulk.tensor.2d per [4 NpInt rows x 32 columns] slab, issued by one lane (UTMALDG in SASS), completing
    // by byte count on the stage's mbarrier -- instead of 4 NpInt / 2 coalesced LDGSTS warp instructions.  tmaps = four
    // CUtensorMap objects (q0, q2, q3, R) in global memory, nullptr = off.  TMA lands dense 256-byte rows, so these slabs
    // then have a row stride of 32 doubles; only the epilogue reads them (LDS.128 in accumulator layout), never the DMMA
    // B-operand path whose 4 k-rows x 8 columns pattern needs the stride of 36.
    const void *tmaps;
};

// 2D tile {c0 = column, c1 = row} of a tensor map -> shared memory, completion by bytes on the mbarrier
__device__ __forceinline__ void tma_load_2d(unsigned dst, const void *tmap, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// DFR2D_WS_KNOCKOUT (timing experiments only, tools/elem_knockout.py -- results are WRONG when non-zero): bit 0 drops the fused
// edge interpolation (C2 and the Q_Face stores), bit 1 the physical flux evaluation, bit 3 the DMMA of the first
// contraction, bit 4 all consumer work (the producers then run against the ring alone: the copy / HBM floor of the kernel),
// bit 2 the stores of the new stage register, bit 5 the Q_Face stores but not C2, bit 6 C2 but not its stores
#ifndef DFR2D_WS_KNOCKOUT
#define DFR2D_WS_KNOCKOUT 0
#endif
__device__ __forceinline__ void ko_keep(double &acc, double v) {      // keeps v alive without touching the FP64 pipe
    acc = __longlong_as_double(__double_as_longlong(acc) ^ __double_as_longlong(v));
}

// CW = consumer warps (8: two groups of four, 224 registers each; 12: three groups, 152 registers each)
// DISS = the element kernel of the PerssonC0 path (AddDissipation, dissipation.go:274-346, and LimitFilterSolution on
// RHSQ, euler.go:496-501, dissipation.go:520-542) on the same ring: the interior dissipation DOFs join Fr / Fs before
// the DivInt contraction (the edge DOFs already carry F - F_visc: k_visc_edge subtracts in place), the modal limiter is
// two more DMMA products through the warp's own columns of the dead DissX / DissY slabs, the viscous dt limit joins the
// dt logic (euler.go:945-1002), and there is no fused interpolation (k_diss_prepare needs the vertex-merged sigma first).
// SPLIT (r2, inviscid kernel only): the fused edge interpolation (C2 + the Q_Face stores) moves to FOUR MORE WARPS.  The
// knock-out timings of tools/elem_knockout.py (profiles/r02v_*, r02w_*: 2M triangles, N=4) showed the kernel bound by the
// latency of its consumer warps, not by the FP64 pipe or by HBM: dropping the flux evaluation changes nothing (1.26 ->
// 1.25 ms), dropping 3/4 of the DMMA buys 10 %, but dropping C2 and its stores buys 29 % (0.90 ms, 83-87 % of the HBM
// peak) -- 150 instructions of a dependent chain at ~10 cycles each that nothing else in the warp can hide.  With SPLIT
// a consumer warp hands its 8 columns of the fresh register over (mbarrier qnewBar[stage][wq]) and moves on to its next
// tile; interpolation warp wq contracts them with FluxEdgeInterp, releases the stage and streams Q_Face out.
template <int N, int CW, bool DISS, bool SPLIT = false>
__global__ void __launch_bounds__((CW + (SPLIT ? 4 : 0) + kWsProdWarps) * 32, 1) k_elem_ws(ElemWsArgs args) {
    static_assert(!SPLIT || (CW == 8 && !DISS), "SPLIT: inviscid kernel with two consumer groups");
    constexpr int kWsConsWarps = CW, kWsInterpWarps = SPLIT ? 4 : 0, kGroups = CW / 4;
    constexpr int kWsThreads = (CW + kWsInterpWarps + kWsProdWarps) * 32;
    // register budget (65,536 per SM): 8 x 32 x 224 + 4 x 32 x 56 | SPLIT: 8 x 32 x 176 + 4 x 32 x 96 + 4 x 32 x 56
    constexpr int kProdRegs = (CW == 8) ? 56 : 40, kConsRegs = SPLIT ? 176 : ((CW == 8) ? 224 : 152), kInterpRegs = 96;
    // setmaxnreg only redistributes the registers the CTA got at launch; asking for more never returns
    static_assert(CW * 32 * kConsRegs + kWsInterpWarps * 32 * kInterpRegs + kWsProdWarps * 32 * kProdRegs <= ((65536 / kWsThreads) & ~7) * kWsThreads,
                  "setmaxnreg budget");
    using TD = WsDim<N>;
    constexpr int NI = TD::NI, NEd = TD::NEd, NF3 = TD::NF3, SE = TD::SE, E = kElemsPerBlock;
    constexpr int M1 = TD::M1, KI = TD::KI, KE = TD::KE, K1 = TD::K1, M2 = TD::M2, K2 = TD::K2;
    constexpr int KO = DFR2D_WS_KNOCKOUT;
    const ElemArgs &a = args.a;
    if (step_is_noop(a.sc, a.ph, a.par, a.stepIndex)) {
        if (a.rk == 4 && a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            a.sc->time[a.par ^ 1] = a.sc->time[a.par];
            a.sc->finished = 1;
        }
        return;
    }
    extern __shared__ __align__(128) double smem_ws[];
    double *smem = smem_ws;
    __shared__ __align__(8) unsigned long long fullBar[kWsMaxStages], emptyBar[kWsMaxStages];
    __shared__ __align__(8) unsigned long long qnewBar[SPLIT ? kWsMaxStages : 1][4];   // consumer warp wq -> interpolation warp wq
    const int S = args.nStages, nExtra = args.nExtra;
    const int nExtraS = TD::extras_in_smem(nExtra, DISS);           // extra RK registers staged in the ring
    const bool useTma = !DISS && args.tmaps != nullptr;
    const int SEX = useTma ? E : SE;                                // row stride of the extra slabs
    constexpr int XO = DISS ? TD::xOffD : TD::xOff;                  // first extra slab
    const int dOff = XO + nExtraS * TD::xSize;                      // DISS: DissX slab, DissY behind it
    const int stageDoubles = DISS ? TD::stage_doubles_diss(nExtraS) : TD::xOff + nExtra * TD::xSize;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t Kp = a.Kp;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&fullBar[s], kWsFullCount);
            mbar_init(&emptyBar[s], 4);              // the four consumer (SPLIT: interpolation) warps of the tile
            if (SPLIT)
                for (int w = 0; w < 4; w++) mbar_init(&qnewBar[s][w], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the two operators pass through shared memory once (the ring is not in use yet) so that every consumer lane can
    // pick its A-fragment entries with lane-dependent indices without serialised constant-bank reads
    {
        const Ops<N> &op = ops<N>();
        constexpr int NFL = Dim<N>::NpFlux;
        for (int t = threadIdx.x; t < NI * NFL; t += kWsThreads) smem[t] = op.DivInt[t / NFL][t % NFL];
        if (!DISS) {
            for (int t = threadIdx.x; t < NF3 * NI; t += kWsThreads) smem[NI * NFL + t] = op.FEI[t / NI][t % NI];
        } else {
            for (int t = threadIdx.x; t < NI * NI; t += kWsThreads) {
                smem[NI * NFL + t] = op.Vinv[t / NI][t % NI];
                smem[NI * NFL + NI * NI + t] = op.V[t / NI][t % NI];
            }
        }
    }
    __syncthreads();

    double dtGlobal = 0.0;
    if (!a.ph.localDT) {
        const double gw = __longlong_as_double((long long)a.sc->wave[a.slot][0]);
        dtGlobal = a.ph.CFL / gw;
        if (DISS) {              // calculateGlobalDT with the viscous limit (euler.go:956-966)
            const double gv = __longlong_as_double((long long)a.sc->wave[a.slot][1]);
            dtGlobal = fmin(dtGlobal, a.ph.Cdiff / gv);
        }
        const double t = a.sc->time[a.par];
        if (t + dtGlobal > a.ph.FinalTime) dtGlobal = a.ph.FinalTime - t;
    }
    const int nLocal = (args.nTiles > (int)blockIdx.x) ? (args.nTiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp >= kWsConsWarps + kWsInterpWarps) {
        // =================================== producer warps ==================================================
        // 168 registers per thread are allotted at launch (12 warps x 168 x 32 = 64,512); the producers keep 56 and
        // hand the rest to the consumers (setmaxnreg, 4 x 56 + 8 x 224 = the same total)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProdRegs));
        const int pw = warp - kWsConsWarps - kWsInterpWarps;
        if (pw == 0) {
            // ---- row slabs by the TMA engine; dt per element (lane = element) -------------------------------------
            int c0 = 0, c1 = 0, c2 = 0;
            double g0 = 0, g1 = 0, g2 = 0, dtOld = 0, gvMax = 0, dtvOld = 0, sig = 0;
            auto load_dt = [&](int n) {
                if (n >= nLocal) return;
                const size_t kk = (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E + lane;
                if (DISS) sig = a.sigma[kk];
                if (a.ph.localDT) {
                    c0 = a.etoe[kk]; c1 = a.etoe[Kp + kk]; c2 = a.etoe[2 * Kp + kk];
                    c0 = c0 >= 0 ? c0 : -1 - c0; c1 = c1 >= 0 ? c1 : -1 - c1; c2 = c2 >= 0 ? c2 : -1 - c2;
                    g0 = a.agg[c0]; g1 = a.agg[c1]; g2 = a.agg[c2];
                    dtOld = a.DT[kk];
                    if (DISS) {
                        gvMax = fmax(fmax(a.aggv[c0], a.aggv[c1]), a.aggv[c2]);
                        dtvOld = a.DTVisc[kk];
                    }
                }
            };
            load_dt(0);
            asm volatile("bar.sync 1, %0;" ::"n"(kWsThreads) : "memory");      // consumers hold their operator fragments
            int s = 0;
            unsigned ph = 1;                        // parity of the previous phase of emptyBar[s]
            for (int n = 0; n < nLocal; n++) {
                if (n >= S) mbar_wait(&emptyBar[s], ph);
                const size_t k0 = (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E;
                double *st = smem + (size_t)s * stageDoubles;
                const unsigned stU = smem_u32(st);
                // stage input, the stage-3 residual (rk 4) and the metrics: coalesced 16-byte async copies
                slab_g2s<4 * NI, SE>(stU + (unsigned)(TD::qOff * sizeof(double)), a.qs + k0, Kp, lane);
                if (!DISS && nExtra == 4) {
                    if (useTma) {
                        if (lane == 0) {
                            mbar_expect_tx(&fullBar[s], (unsigned)(4 * NI * E * sizeof(double)));
                            tma_load_2d(stU + (unsigned)((TD::xOff + 3 * TD::xSize) * sizeof(double)),
                                        (const char *)args.tmaps + 3 * 128, (int)k0, 0, smem_u32(&fullBar[s]));
                        }
                    } else {
                        slab_g2s<4 * NI, SE>(stU + (unsigned)((TD::xOff + 3 * TD::xSize) * sizeof(double)), a.R + k0, Kp, lane);
                    }
                }
                if (lane < 16) cp_async16_u32(stU + (unsigned)((TD::gOff + 2 * lane) * sizeof(double)), a.Jdet + k0 + 2 * lane);
                slab_g2s<4, 32>(stU + (unsigned)((TD::gOff + 32) * sizeof(double)), a.Jinv + k0, Kp, lane);
                double dtk = dtGlobal;
                if (a.ph.localDT) {
                    // InitializeDT at stage 0, DT = max(DT, aggregates), CalculateLocalDT (euler.go:637-643, :973-1002)
                    const double wmaxk = fmax(fmax(g0, g1), g2);
                    const double d = (a.rk == 0) ? -100.0 : dtOld;
                    dtk = a.ph.CFL / fmax(d, wmaxk);
                    if (DISS) {
                        double dtv = fmax(dtvOld, gvMax);
                        if (dtv > 1.e-9) { dtv = a.ph.Cdiff / dtv; dtk = fmin(dtk, dtv); }
                        if (a.rhsOut == nullptr && k0 + lane < (size_t)a.K) a.DTVisc[k0 + lane] = dtv;
                    }
                    if (a.rhsOut == nullptr && (!DISS || k0 + lane < (size_t)a.K)) a.DT[k0 + lane] = dtk;
                }
                st[TD::gOff + 8 * 32 + lane] = dtk;
                if (DISS) st[TD::gOff + 9 * 32 + lane] = 1.0 - sin(0.5 * 3.14159265358979323846 * sig);
                cp_async_arrive_noinc(&fullBar[s]);
                mbar_arrive(&fullBar[s]);
                load_dt(n + 1);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        } else {
            // ---- gather of the numerical flux of local edge `le` of every element of the tile (lane = element): the
            // source walks the edge's rows upwards (one pointer increment per copy), the destination row runs up for
            // the owner and down (reversed point order) for the neighbour; sign x IInII of SetRTFluxOnEdges
            // (edges.go:469-479) goes along as a per-element scale
            const int le = pw - 1;
            int cs = 0, ns = 0, ns2 = 0;            // edge slot of tile n, n+1, n+2 (loads are issued three tiles ahead)
            double iin = 0.0, iinN = 0.0;           // IInII of tile n, n+1 (two tiles ahead)
            auto load_slot = [&](int n, int &sl) {
                if (n < nLocal) sl = a.etoe[(size_t)le * Kp + (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E + lane];
            };
            auto load_iin = [&](int n, double &v) {
                if (n < nLocal) v = a.IInII[(size_t)le * Kp + (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E + lane];
            };
            load_slot(0, cs);
            load_iin(0, iin);
            load_slot(1, ns);
            load_iin(1, iinN);
            load_slot(2, ns2);
            asm volatile("bar.sync 1, %0;" ::"n"(kWsThreads) : "memory");
            const size_t srcStep = (size_t)a.NEp;
            int s = 0;
            unsigned ph = 1;                        // parity of the previous phase of emptyBar[s]
            for (int n = 0; n < nLocal; n++) {
                if (n >= S) mbar_wait(&emptyBar[s], ph);
                double *st = smem + (size_t)s * stageDoubles;
                if (le < (DISS ? nExtraS : (nExtra == 4 ? 3 : nExtra))) {
                    // RK register slab number `le` of this stage (q0, q2, q3); R travels with the stage input in warp 0
                    const size_t k0 = (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E;
                    if (useTma) {
                        if (lane == 0) {
                            mbar_expect_tx(&fullBar[s], (unsigned)(4 * NI * E * sizeof(double)));
                            tma_load_2d(smem_u32(st) + (unsigned)((XO + le * TD::xSize) * sizeof(double)),
                                        (const char *)args.tmaps + le * 128, (int)k0, 0, smem_u32(&fullBar[s]));
                        }
                    } else {
                        slab_g2s<4 * NI, SE>(smem_u32(st) + (unsigned)((XO + le * TD::xSize) * sizeof(double)),
                                             (le == 0 ? a.q0 : (le == 1 ? a.q2 : a.q3)) + k0, Kp, lane);
                    }
                }
                if (DISS && le >= 1) {
                    // interior rows of DissX (gather warp 1) and DissY (gather warp 2)
                    const size_t k0 = (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E;
                    slab_g2s<4 * NI, SE>(smem_u32(st) + (unsigned)((dOff + (le - 1) * TD::xSize) * sizeof(double)),
                                         (le == 1 ? a.dissX : a.dissY) + k0, Kp, lane);
                }
                const bool own = cs >= 0;
                st[TD::gOff + (5 + le) * 32 + lane] = own ? iin : -iin;
                const double *src = a.eflux + (own ? cs : -1 - cs);
                const int dStep = own ? (int)(SE * sizeof(double)) : -(int)(SE * sizeof(double));
                const unsigned seU = smem_u32(st) + (unsigned)((TD::eOff + lane + (le * NEd + (own ? 0 : NEd - 1)) * SE) * sizeof(double));
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    unsigned d = seU + (unsigned)(v * NF3 * SE * sizeof(double));
#pragma unroll
                    for (int i = 0; i < NEd; i++) {
                        cp_async8_u32(d, src);
                        src += srcStep;
                        d += dStep;
                    }
                }
                cp_async_arrive_noinc(&fullBar[s]);
                mbar_arrive(&fullBar[s]);
                cs = ns; ns = ns2; iin = iinN;
                load_iin(n + 2, iinN);
                load_slot(n + 3, ns2);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
    } else if (SPLIT && warp >= kWsConsWarps) {
        // =================================== interpolation warps (SPLIT) ======================================
        // warp wq owns columns 8 wq .. 8 wq + 7 of EVERY tile of this CTA (both consumer groups), in tile order
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kInterpRegs));
        const int wq = warp - kWsConsWarps;
        const int fr = lane >> 2, fc = lane & 3;
        const int eB = 8 * wq + fr, eC = 8 * wq + 2 * fc;
        constexpr int NFL = Dim<N>::NpFlux;
        const double *opF = smem + NI * NFL;
        double a2[M2][K2];
#pragma unroll
        for (int mt = 0; mt < M2; mt++) {
            const int row = 8 * mt + fr;
#pragma unroll
            for (int ks = 0; ks < K2; ks++) {
                const int j = 4 * ks + fc;
                a2[mt][ks] = (row < NF3 && j < NI) ? opF[row * NI + j] : 0.0;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWsThreads) : "memory");
        // (measured and dropped, profiles/r02ac_*: storing the fresh register from here as well -- the B fragments of a warp
        // are exactly its 8 columns of it -- makes these warps the bottleneck: kernel 4.70 -> 5.13 ms at C5)
        const bool doInterp = a.rhsOut == nullptr && a.qface != nullptr;
        int s = 0;
        unsigned ph = 0;
        for (int n = 0; n < nLocal; n++) {
            const size_t k0 = (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E;
            const double *sQ = smem + (size_t)s * stageDoubles + TD::qOff;
            mbar_wait(&qnewBar[s][wq], ph);
            double tail[TD::NTail > 0 ? TD::NTail : 1];
            double b[K2][4];
            if (doInterp) {
                // ragged last rows of FluxEdgeInterp by DFMA: lane = (variable, column) of the warp's 8 columns
                if (TD::NTail > 0) {
                    const Ops<N> &op = ops<N>();
                    const double *qv = sQ + ((lane >> 3) * NI) * SE + 8 * wq + (lane & 7);
#pragma unroll
                    for (int r = 0; r < TD::NTail; r++) tail[r] = 0.0;
#pragma unroll
                    for (int j = 0; j < NI; j++) {
                        const double q = qv[j * SE];
#pragma unroll
                        for (int r = 0; r < TD::NTail; r++) tail[r] = fma(op.FEI[8 * M2 + r][j], q, tail[r]);
                    }
                }
                // every B fragment of C2 up front: the stage goes back to the producers before the tensor work starts
#pragma unroll
                for (int ks = 0; ks < K2; ks++) {
                    const int j = 4 * ks + fc;
#pragma unroll
                    for (int v = 0; v < 4; v++) b[ks][v] = (4 * ks + 3 < NI || j < NI) ? sQ[(v * NI + j) * SE + eB] : 0.0;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyBar[s]);
            if (doInterp) {
                double c2[4][M2][2];
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M2; mt++) c2[v][mt][0] = c2[v][mt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < K2; ks++)
#pragma unroll
                    for (int v = 0; v < 4; v++)
#pragma unroll
                        for (int mt = 0; mt < M2; mt++) dmma884(c2[v][mt][0], c2[v][mt][1], a2[mt][ks], b[ks][v]);
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M2; mt++) {
                        const int m = 8 * mt + fr;
                        if (m < NF3)
                            *reinterpret_cast<double2 *>(a.qface + ((size_t)v * NF3 + m) * Kp + k0 + eC) =
                                make_double2(c2[v][mt][0], c2[v][mt][1]);
                    }
#pragma unroll
                for (int r = 0; r < TD::NTail; r++)
                    a.qface[((size_t)(lane >> 3) * NF3 + 8 * M2 + r) * Kp + k0 + 8 * wq + (lane & 7)] = tail[r];
            }
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else {
        // =================================== consumer warps ==================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kConsRegs));
        const int group = warp >> 2, wq = warp & 3;
        const int fr = lane >> 2, fc = lane & 3;
        // Row of the first contraction held by accumulator row fr: bits 0 and 1 of fr swapped (0,2,1,3,4,6,5,7).  The 128-bit
        // shared-memory accesses of the epilogue are served per quarter warp (fr = 2j, 2j+1; fc = 0..3); with the natural
        // order the two rows lie 72 words = 8 banks apart and collide 2-way (ncu: 37 % of the shared wavefronts of this
        // kernel were conflicts, profiles/r01j_*), with the swap they lie 144 words = 16 banks apart: conflict free.  The
        // operator fragments below are built with the same permutation, nothing else notices.
        const int frP = (fr & 4) | ((fr & 1) << 1) | ((fr >> 1) & 1);
        const int eB = 8 * wq + fr;                  // B-operand column of this lane
        const int eC = 8 * wq + 2 * fc;              // first of the two accumulator columns of this lane
        constexpr int NFL = Dim<N>::NpFlux;
        const double *opD = smem, *opF = smem + NI * NFL;
        // operator fragments, A[row = fr][k = fc] of every (m-tile, k-step); k-steps 2m / 2m+1 carry Fr / Fs of points
        // 4m..4m+3, k-steps 2KI.. the edge rows
        double a1[M1][K1], a2[(DISS || SPLIT) ? 1 : M2][(DISS || SPLIT) ? 1 : K2];
        // DISS: Vinv and V (LimitFilterSolution, dissipation.go:606-622) as A fragments, both with the permuted row order;
        // mfRow = ModeFilter of this lane's modes (mode 0 is never scaled)
        double aVi[DISS ? M1 : 1][DISS ? KI : 1], aV[DISS ? M1 : 1][DISS ? KI : 1], mfRow[DISS ? M1 : 1];
        if (DISS) {
            const double *opVi = smem + NI * NFL, *opV = opVi + NI * NI;
#pragma unroll
            for (int mt = 0; mt < M1; mt++) {
                const int row = 8 * mt + frP;
                mfRow[mt] = (row >= 1 && row < NI) ? ops<N>().mf[row] : 0.0;
#pragma unroll
                for (int ks = 0; ks < KI; ks++) {
                    const int j = 4 * ks + fc;
                    aVi[mt][ks] = (row < NI && j < NI) ? opVi[row * NI + j] : 0.0;
                    aV[mt][ks] = (row < NI && j < NI) ? opV[row * NI + j] : 0.0;
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < M1; mt++) {
            const int row = 8 * mt + frP;
#pragma unroll
            for (int m = 0; m < KI; m++) {
                const int p = 4 * m + fc;
                if (p < NI) {
                    a1[mt][2 * m] = row < NI ? opD[row * NFL + p] : 0.0;
                    a1[mt][2 * m + 1] = row < NI ? opD[row * NFL + NI + p] : 0.0;
                } else {                // idle point slot: edge rows NE3k + t, NE3k + t + 1 ride here
                    const int t = 2 * (p - NI);
                    a1[mt][2 * m] = (row < NI && t < TD::nMoved) ? opD[row * NFL + 2 * NI + TD::NE3k + t] : 0.0;
                    a1[mt][2 * m + 1] = (row < NI && t + 1 < TD::nMoved) ? opD[row * NFL + 2 * NI + TD::NE3k + t + 1] : 0.0;
                }
            }
#pragma unroll
            for (int ke = 0; ke < KE; ke++) {
                const int r = 4 * ke + fc;
                a1[mt][2 * KI + ke] = (row < NI && r < TD::NE3k) ? opD[row * NFL + 2 * NI + r] : 0.0;
            }
        }
        if (!DISS && !SPLIT) {
#pragma unroll
            for (int mt = 0; mt < M2; mt++) {
                const int row = 8 * mt + fr;
#pragma unroll
                for (int ks = 0; ks < K2; ks++) {
                    const int j = 4 * ks + fc;
                    a2[(DISS || SPLIT) ? 0 : mt][(DISS || SPLIT) ? 0 : ks] = (row < NF3 && j < NI) ? opF[row * NI + j] : 0.0;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWsThreads) : "memory");      // operator table consumed: the ring may be filled
        bool bad = false;
        double resMax[4] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
        double *dst = (a.rk == 0) ? a.q1 : (a.rk == 1) ? a.q2 : (a.rk == 2) ? a.q3 : (a.rk == 3) ? a.q4 : a.q0;

        int s = group % S;
        unsigned ph = (unsigned)((group / S) & 1);
        for (int n = group; n < nLocal; n += kGroups) {
            const size_t k0 = (size_t)(blockIdx.x + (size_t)n * gridDim.x) * E;
            double *st = smem + (size_t)s * stageDoubles;
            double *sQ = st + TD::qOff;
            const double *sE = st + TD::eOff, *g = st + TD::gOff;
            mbar_wait(&fullBar[s], ph);
            if (KO & 16) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&emptyBar[s]);
#pragma unroll
                for (int t = 0; t < kGroups; t++)
                    if (++s == S) { s = 0; ph ^= 1u; }
                continue;
            }

            // ---- C1 = DivInt . F_RT_DOF with the flux evaluated in B-fragment layout ----------------------------
            const double jd = g[eB], j0 = g[32 + eB], j1 = g[64 + eB], j2 = g[96 + eB], j3 = g[128 + eB];
            const double sc0 = g[5 * 32 + eB], sc1 = g[6 * 32 + eB], sc2 = g[7 * 32 + eB];
            double c1[4][M1][2];
#pragma unroll
            for (int v = 0; v < 4; v++)
#pragma unroll
                for (int mt = 0; mt < M1; mt++) c1[v][mt][0] = c1[v][mt][1] = 0.0;
#pragma unroll
            for (int m = 0; m < KI; m++) {
                const int p = 4 * m + fc;
                double Fr[4] = {0.0, 0.0, 0.0, 0.0}, Fs[4] = {0.0, 0.0, 0.0, 0.0};
                if (4 * m + 3 < NI || p < NI) {
                    double Q[4], Fx[4], Fy[4];
#pragma unroll
                    for (int v = 0; v < 4; v++) Q[v] = sQ[(v * NI + p) * SE + eB];
                    if (KO & 2) {
#pragma unroll
                        for (int v = 0; v < 4; v++) { Fx[v] = Q[v]; Fy[v] = Q[(v + 1) & 3]; }
                    } else {
                        flux_calc(a.ph.gamma, Q, Fx, Fy);
                    }
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        Fr[v] = jd * (j0 * Fx[v] + j1 * Fy[v]);
                        Fs[v] = jd * (j2 * Fx[v] + j3 * Fy[v]);
                    }
                    if (DISS) {          // AddDissipation interior DOFs (dissipation.go:310-315), opposite sign
                        const double *sDX = st + dOff, *sDY = sDX + TD::xSize;
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            const double dix = sDX[(v * NI + p) * SE + eB], diy = sDY[(v * NI + p) * SE + eB];
                            Fr[v] -= jd * (j0 * dix + j1 * diy);
                            Fs[v] -= jd * (j2 * dix + j3 * diy);
                        }
                    }
                } else if (TD::nMoved > 0) {
                    // idle point slot of the last interior k-steps: the numerical flux of edge rows NE3k + t (Fr slot) and
                    // NE3k + t + 1 (Fs slot), scaled like the edge block below
                    const int t = 2 * (p - NI);
                    if (t < TD::nMoved) {
                        const int r = TD::NE3k + t;
                        const double scl = pick3((r >= NEd) + (r >= 2 * NEd), sc0, sc1, sc2);
#pragma unroll
                        for (int v = 0; v < 4; v++) Fr[v] = sE[(v * NF3 + r) * SE + eB] * scl;
                    }
                    if (t + 1 < TD::nMoved) {
                        const int r = TD::NE3k + t + 1;
                        const double scl = pick3((r >= NEd) + (r >= 2 * NEd), sc0, sc1, sc2);
#pragma unroll
                        for (int v = 0; v < 4; v++) Fs[v] = sE[(v * NF3 + r) * SE + eB] * scl;
                    }
                }
                if (KO & 8) {
#pragma unroll
                    for (int v = 0; v < 4; v++) { ko_keep(c1[v][0][0], Fr[v]); ko_keep(c1[v][M1 - 1][1], Fs[v]); }
                    continue;
                }
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) dmma884(c1[v][mt][0], c1[v][mt][1], a1[mt][2 * m], Fr[v]);
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) dmma884(c1[v][mt][0], c1[v][mt][1], a1[mt][2 * m + 1], Fs[v]);
            }
#pragma unroll
            for (int ke = 0; ke < KE; ke++) {
                const int r = 4 * ke + fc;
                double b[4] = {0.0, 0.0, 0.0, 0.0};
                if (4 * ke + 3 < TD::NE3k || r < TD::NE3k) {
                    const int le = (r >= NEd) + (r >= 2 * NEd);
                    const double scl = pick3(le, sc0, sc1, sc2);
#pragma unroll
                    for (int v = 0; v < 4; v++) b[v] = sE[(v * NF3 + r) * SE + eB] * scl;
                }
                if (KO & 8) {
#pragma unroll
                    for (int v = 0; v < 4; v++) ko_keep(c1[v][0][0], b[v]);
                    continue;
                }
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) dmma884(c1[v][mt][0], c1[v][mt][1], a1[mt][2 * KI + ke], b[v]);
            }

            // ---- epilogue in accumulator layout: rows 8 mt + fr, columns eC, eC + 1 ------------------------------
            const bool fullTile = k0 + E <= (size_t)a.K;
            const double2 jdc = *reinterpret_cast<const double2 *>(&g[eC]);
            const double2 dt = *reinterpret_cast<const double2 *>(&g[8 * 32 + eC]);
            double mo0 = -(1.0 / jdc.x), mo1 = -(1.0 / jdc.y);
            const double *sX = st + XO;
            if (DISS) {
                // LimitFilterSolution(RHSQ): RHS = -(1/J) C1 -> Vinv -> modes i >= 1 times mf_i (1 - sin(pi sigma / 2)) -> V,
                // through this warp's own 8 columns of the DissX slab (RHS) and the DissY slab (modes): both are dead
                // once the flux loop above has read them, and no other warp ever touches these columns
                double *sR = st + dOff, *sU = sR + TD::xSize;
                const double2 oma = *reinterpret_cast<const double2 *>(&g[9 * 32 + eC]);
                __syncwarp();
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) {
                        const int i = 8 * mt + frP;
                        if (i < NI)
                            *reinterpret_cast<double2 *>(&sR[(v * NI + i) * SE + eC]) = make_double2(c1[v][mt][0] * mo0, c1[v][mt][1] * mo1);
                    }
                __syncwarp();
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) c1[v][mt][0] = c1[v][mt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KI; ks++) {
                    const int j = 4 * ks + fc;
                    double b[4] = {0.0, 0.0, 0.0, 0.0};
                    if (4 * ks + 3 < NI || j < NI) {
#pragma unroll
                        for (int v = 0; v < 4; v++) b[v] = sR[(v * NI + j) * SE + eB];
                    }
#pragma unroll
                    for (int v = 0; v < 4; v++)
#pragma unroll
                        for (int mt = 0; mt < M1; mt++) dmma884(c1[v][mt][0], c1[v][mt][1], aVi[mt][ks], b[v]);
                }
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) {
                        const int i = 8 * mt + frP;
                        if (i < NI) {
                            double2 uh = make_double2(c1[v][mt][0], c1[v][mt][1]);
                            if (i >= 1) { uh.x *= mfRow[mt] * oma.x; uh.y *= mfRow[mt] * oma.y; }
                            *reinterpret_cast<double2 *>(&sU[(v * NI + i) * SE + eC]) = uh;
                        }
                    }
                __syncwarp();
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M1; mt++) c1[v][mt][0] = c1[v][mt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KI; ks++) {
                    const int j = 4 * ks + fc;
                    double b[4] = {0.0, 0.0, 0.0, 0.0};
                    if (4 * ks + 3 < NI || j < NI) {
#pragma unroll
                        for (int v = 0; v < 4; v++) b[v] = sU[(v * NI + j) * SE + eB];
                    }
#pragma unroll
                    for (int v = 0; v < 4; v++)
#pragma unroll
                        for (int mt = 0; mt < M1; mt++) dmma884(c1[v][mt][0], c1[v][mt][1], aV[mt][ks], b[v]);
                }
                mo0 = mo1 = 1.0;         // c1 now holds the limited RHS itself
            }
#pragma unroll
            for (int v = 0; v < 4; v++) {
#pragma unroll
                for (int mt = 0; mt < M1; mt++) {
                    const int i = 8 * mt + frP;
                    if (i < NI) {
                        const double rhs0 = c1[v][mt][0] * mo0, rhs1 = c1[v][mt][1] * mo1;
                        const size_t o = ((size_t)v * NI + i) * Kp + k0 + eC;
                        if (a.rhsOut != nullptr) {
                            *reinterpret_cast<double2 *>(a.rhsOut + o) = make_double2(rhs0, rhs1);
                            continue;
                        }
                        const int so = (v * NI + i) * SE + eC;
                        const int sx = (v * NI + i) * SEX + eC;          // extra slabs: stride 32 when the TMA engine wrote them
                        const double2 qs = *reinterpret_cast<const double2 *>(&sQ[so]);
                        double2 qn;
                        if (a.rk == 0) {
                            qn.x = qs.x + RK0_A * (dt.x * rhs0);
                            qn.y = qs.y + RK0_A * (dt.y * rhs1);
                        } else if (a.rk < 4) {
                            const double2 q0v = *reinterpret_cast<const double2 *>(&sX[sx]);
                            const double ca = (a.rk == 1) ? RK1_A : (a.rk == 2) ? RK2_A : RK3_A;
                            const double cb = (a.rk == 1) ? RK1_B : (a.rk == 2) ? RK2_B : RK3_B;
                            const double cc = (a.rk == 1) ? RK1_C : (a.rk == 2) ? RK2_C : RK3_C;
                            qn.x = ca * q0v.x + cb * qs.x + cc * (dt.x * rhs0);
                            qn.y = ca * q0v.y + cb * qs.y + cc * (dt.y * rhs1);
                            if (a.rk == 3) *reinterpret_cast<double2 *>(a.R + o) = make_double2(rhs0, rhs1);
                        } else {
                            const double2 q0v = *reinterpret_cast<const double2 *>(&sX[sx]);
                            const double2 q2v = *reinterpret_cast<const double2 *>(&sX[sx + 1 * TD::xSize]);
                            // (DISS: q3 and the stage-3 residual are not staged at rk 4 -- no room for two ring stages)
                            const double2 q3v = DISS ? *reinterpret_cast<const double2 *>(a.q3 + o)
                                                     : *reinterpret_cast<const double2 *>(&sX[sx + 2 * TD::xSize]);
                            const double2 rv = DISS ? *reinterpret_cast<const double2 *>(a.R + o)
                                                    : *reinterpret_cast<const double2 *>(&sX[sx + 3 * TD::xSize]);
                            double2 r;
                            r.x = -q0v.x + RK4_A * q2v.x + RK4_B * q3v.x + RK4_C * qs.x + RK4_D * (dt.x * rv.x) + RK4_E * (dt.x * rhs0);
                            r.y = -q0v.y + RK4_A * q2v.y + RK4_B * q3v.y + RK4_C * qs.y + RK4_D * (dt.y * rv.y) + RK4_E * (dt.y * rhs1);
                            qn.x = q0v.x + r.x;
                            qn.y = q0v.y + r.y;
                            // the residual register is only ever reduced to its per-variable maximum (PrintUpdate,
                            // euler.go:821-835): reduce here instead of writing it out (pad columns excluded)
                            if (fullTile || k0 + eC < (size_t)a.K) resMax[v] = fmax(resMax[v], r.x);
                            if (fullTile || k0 + eC + 1 < (size_t)a.K) resMax[v] = fmax(resMax[v], r.y);
                        }
                        bad |= (qn.x != qn.x) || (qn.y != qn.y);
                        if (!(KO & 4)) *reinterpret_cast<double2 *>(dst + o) = qn;
                        *reinterpret_cast<double2 *>(&sQ[so]) = qn;      // in place: B operand of the interpolation below
                    }
                }
            }
            if (SPLIT) {
                // the fresh register is in this warp's columns of the stage-input slab: over to interpolation warp wq
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&qnewBar[s][wq]);
#pragma unroll
                for (int t = 0; t < kGroups; t++)
                    if (++s == S) { s = 0; ph ^= 1u; }
                continue;
            }
            const bool doInterp = !(KO & 1) && !DISS && a.rhsOut == nullptr && a.qface != nullptr;
            double c2[4][M2][2];
            double tail[TD::NTail > 0 ? TD::NTail : 1];
            if (KO & 64) {            // stores without the contraction
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M2; mt++) c2[v][mt][0] = c2[v][mt][1] = 0.0;
#pragma unroll
                for (int r = 0; r < TD::NTail; r++) tail[r] = 0.0;
            }
            if (doInterp && !(KO & 64)) {
                __syncwarp();
                // ---- C2 = FluxEdgeInterp . q_new: next stage's Q_Face ---------------------------------------------
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M2; mt++) c2[v][mt][0] = c2[v][mt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < K2; ks++) {
                    const int j = 4 * ks + fc;
                    double b[4] = {0.0, 0.0, 0.0, 0.0};
                    if (4 * ks + 3 < NI || j < NI) {
#pragma unroll
                        for (int v = 0; v < 4; v++) b[v] = sQ[(v * NI + j) * SE + eB];
                    }
#pragma unroll
                    for (int v = 0; v < 4; v++)
#pragma unroll
                        for (int mt = 0; mt < M2; mt++) dmma884(c2[v][mt][0], c2[v][mt][1], a2[(DISS || SPLIT) ? 0 : mt][(DISS || SPLIT) ? 0 : ks], b[v]);
                }
                // ragged last rows of FluxEdgeInterp by DFMA: lane = (variable, column) of the warp's 8 columns
                if (TD::NTail > 0) {
                    const Ops<N> &op = ops<N>();
                    const double *qv = sQ + ((lane >> 3) * NI) * SE + 8 * wq + (lane & 7);
#pragma unroll
                    for (int r = 0; r < TD::NTail; r++) tail[r] = 0.0;
#pragma unroll
                    for (int j = 0; j < NI; j++) {
                        const double q = qv[j * SE];
#pragma unroll
                        for (int r = 0; r < TD::NTail; r++) tail[r] = fma(op.FEI[8 * M2 + r][j], q, tail[r]);
                    }
                }
            }
            // ---- the stage is consumed: hand it back to the producer (generic-proxy accesses ordered before the
            // async-proxy refill), then stream out Q_Face from registers
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyBar[s]);
            if (doInterp && (KO & 32)) {          // the contraction without its stores
                double sink = 0.0;
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M2; mt++) { ko_keep(sink, c2[v][mt][0]); ko_keep(sink, c2[v][mt][1]); }
#pragma unroll
                for (int r = 0; r < TD::NTail; r++) ko_keep(sink, tail[r]);
                bad |= (__double_as_longlong(sink) == 0x123456789abcll);
            } else if (doInterp) {
#pragma unroll
                for (int v = 0; v < 4; v++)
#pragma unroll
                    for (int mt = 0; mt < M2; mt++) {
                        const int m = 8 * mt + fr;
                        if (m < NF3)
                            *reinterpret_cast<double2 *>(a.qface + ((size_t)v * NF3 + m) * Kp + k0 + eC) =
                                make_double2(c2[v][mt][0], c2[v][mt][1]);
                    }
#pragma unroll
                for (int r = 0; r < TD::NTail; r++)
                    a.qface[((size_t)(lane >> 3) * NF3 + 8 * M2 + r) * Kp + k0 + 8 * wq + (lane & 7)] = tail[r];
            }
#pragma unroll
            for (int t = 0; t < kGroups; t++)       // this group's next tile is kGroups fills further on
                if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (bad) a.sc->nanFlag = 1;
        if (a.rk == 4 && a.rhsOut == nullptr) {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double m = warp_max(resMax[v]);
                if (lane == 0 && nLocal > group) atomicMax(&a.sc->resMax[v], res_encode(m));
            }
        }
    }

    if (a.rhsOut == nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        a.sc->wave[a.slot ^ 1][0] = 0ull;
        a.sc->wave[a.slot ^ 1][1] = 0ull;
        // the residual maxima of this step are accumulated by the rk 4 launch: cleared here, inside a launch that only runs
        // when the step does (a host-side memset would also clear them on the no-op steps after FinalTime)
        if (a.rk == 3) a.sc->resMax[0] = a.sc->resMax[1] = a.sc->resMax[2] = a.sc->resMax[3] = 0ull;
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
