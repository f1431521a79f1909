/*
 * dfr2d.h -- C ABI of the B200 device library for gocfd's 2D Euler DFR right-hand side and
 * SSP-RK(5,4) time step (float64, sm_100a CUDA, no CPU fallback).
 *
 * The Go host (cmd/2D.go -> Euler2D.NewEuler -> Euler.Solve) keeps mesh input, YAML input and
 * all DG2D operator construction.  It hands the flattened problem to dfr2d_create() once and
 * then replaces `c.RK.Step(c)` (model_problems/Euler2D/euler.go:177) with dfr2d_step().
 * Citations below are file:line in the gocfd reference tree.
 *
 * Conventions
 *   - every matrix is row-major; field arrays are [node rows][element columns] exactly like
 *     utils.Matrix (utils/matrix_extended.go:33-55): index = k + i*K.
 *   - all pointers in dfr2d_problem are host memory owned by the caller and only read during
 *     dfr2d_create(); the library copies what it needs (cgo pointer rules are respected).
 *   - every entry point returns 0 on success, non-zero on failure; dfr2d_last_error() gives the
 *     message.  The Go shim panics on non-zero, matching the reference's panic-on-error style
 *     (euler.go:149, utils/system.go:22 "NAN found").
 *   - a handle is not re-entrant: one caller at a time (the reference's single controller
 *     goroutine, euler.go:408-418).
 */
#ifndef DFR2D_H
#define DFR2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* types/cfd.go:13-26 -- BC flag numbering is part of the ABI */
enum {
    DFR2D_BC_None = 0, DFR2D_BC_In = 1, DFR2D_BC_Dirichlet = 2, DFR2D_BC_Slip = 3, DFR2D_BC_Far = 4,
    DFR2D_BC_Wall = 5, DFR2D_BC_Cyl = 6, DFR2D_BC_Neuman = 7, DFR2D_BC_Out = 8, DFR2D_BC_IVortex = 9,
    DFR2D_BC_Periodic = 10, DFR2D_BC_PeriodicReversed = 11
};
/* fluxes.go:18-23 */
enum { DFR2D_FLUX_Average = 0, DFR2D_FLUX_LaxFriedrichs = 1, DFR2D_FLUX_Roe = 2, DFR2D_FLUX_RoeER = 3 };
/* initialization.go:17-23 */
enum { DFR2D_CASE_Freestream = 0, DFR2D_CASE_IVortex = 1, DFR2D_CASE_ShockTube = 2 };
/* dfr2d_get_field selectors */
enum { DFR2D_FIELD_DT = 0, DFR2D_FIELD_SigmaScalar = 1, DFR2D_FIELD_EpsilonScalar = 2, DFR2D_FIELD_Se = 3 };

#define DFR2D_MAX_ORDER 4
#define DFR2D_ERR_NAN 7
#define DFR2D_ERR_PEER 8   /* a partner partition never delivered its message (peer exchange timed out) */

/* fluids.go:237-243 FreeStream */
typedef struct dfr2d_freestream {
    double Gamma;
    double Qinf[4];
    double Pinf, QQinf, Cinf;
    double Alpha, Minf;
} dfr2d_freestream;

/* isentropic_vortex/analytic_vortex.go:9-12 */
typedef struct dfr2d_vortex {
    double Beta, X0, Y0, Gamma, Ufs;
} dfr2d_vortex;

/*
 * Everything NewEuler + NewRungeKuttaSSP leave in place for the time loop (euler.go:98-117,
 * :343-406), flattened.  Np* follow from N: NpInt=(N+1)(N+2)/2, NpEdge=N+2,
 * NpFlux=(N+2)(N+4) (raviart_thomas_element.go:199-203 with RT order N+1).
 */
typedef struct dfr2d_problem {
    int32_t N;                    /* PolynomialOrder, 0..DFR2D_MAX_ORDER */
    int32_t flux_type;            /* c.FluxCalcAlgo */
    int32_t init_case;            /* c.Case (informational) */
    int32_t local_time_stepping;  /* c.LocalTimeStepping */
    int32_t max_iterations;       /* c.MaxIterations */
    int32_t dissipation;          /* c.Dissipation != nil: PerssonC0 limiter requested and N != 0 (euler.go:110) */
    int64_t K, NV, NE;            /* elements, vertices, edges */
    int64_t NBP;                  /* boundary edges for which edge-point coordinates are given */
    double CFL, FinalTime;
    double Kappa;                 /* ip.Kappa as given (0 -> the dissipation default 5, dissipation.go:140-147) */
    dfr2d_freestream FSFar, FSIn, FSOut;
    dfr2d_vortex vortex;          /* analytic state of BC_IVortex edges */

    /* reference-element operators */
    const double *FluxEdgeInterp; /* [3NpEdge x NpInt]  DFR.FluxEdgeInterp          dfr_startup.go:64   */
    const double *DivInt;         /* [NpInt x NpFlux]   DFR.FluxElement.DivInt      raviart_thomas_element.go:243-247 */
    const double *Div;            /* [NpFlux x NpFlux]  DFR.FluxElement.Div         (dissipation only) */
    const double *V, *Vinv;       /* [NpInt x NpInt]    SolutionElement.JB2D        basis_polynomials.go:41-43 */
    const double *MassMatrix, *D, *P; /* [NpInt x NpInt] ModeAliasShockFinder       dfr_shock_capturing.go:84-105 */
    const double *ModeFilter;     /* [NpInt]                                        dfr_shock_capturing.go:43-69 */
    const double *Bary;           /* [NpFlux x 3]       Dissipation.BaryCentricCoords dissipation.go:414-449 */

    /* geometry, global (un-sharded) */
    const double *Jdet;           /* [K]        dfr_startup.go:256-287 */
    const double *Jinv;           /* [K x 4]    element-major {rx,ry,sx,sy} */
    const double *FaceNormX, *FaceNormY; /* [3 x K]   DFR.FaceNorm[0|1], index k + K*edge, dfr_startup.go:289-310 */
    const double *IInII;          /* [3 x K]    dfr_startup.go:301 */
    const double *EdgeLenMax;     /* [K]        dfr_startup.go:125-147 (hK = EdgeLenMax/(N+1)^2) */

    /* topology */
    const int32_t *EToV;          /* [K x 3]    Tris.EToV */
    const int32_t *edge_kL, *edge_kR;       /* [NE] Edge.ConnectedTris[0|1]; kR = -1 when NumConnectedTris == 1 */
    const int32_t *edge_numL, *edge_numR;   /* [NE] Edge.ConnectedTriEdgeNumber[0|1] */
    const int32_t *edge_nconn;    /* [NE]       Edge.NumConnectedTris */
    const int32_t *edge_bc;       /* [NE]       Edge.BCType (after the IVORTEX wall->ivortex rewrite, euler.go:781-787) */
    const double *edge_len;       /* [NE]       Edge.GetEdgeLength() triangulation.go:189-198 */
    const int32_t *EtoEdge;       /* [K x 3]    position in this table of DFR.EdgeNumber[k + K*e] */

    /* DFR.FluxX/FluxY rows 2NpInt.. of boundary elements only (bcs.go:33-37) */
    const int32_t *bp_edge;       /* [NBP]          edge index */
    const double *bp_x, *bp_y;    /* [NBP x NpEdge] in the owner's edge-point order */
} dfr2d_problem;

typedef struct dfr2d_step_info {
    double time;        /* rk.Time */
    double dt;          /* rk.GlobalDT of the last stage (0 with local time stepping) */
    int64_t steps;      /* rk.StepCount */
    int32_t finished;   /* CheckIfFinished euler.go:796-801 */
    int32_t nan_found;  /* utils.IsNanPanic would have fired */
} dfr2d_step_info;

typedef struct dfr2d_handle dfr2d_handle;

/*
 * Build one partition of the solver on CUDA device `device`.  n_parts/part select the
 * contiguous element range of utils.PartitionMap.Split1D (utils/parallel_utils.go:172-192);
 * n_parts = 1 is the whole mesh on one GPU.  Mirrors the tail of NewEuler + NewRungeKuttaSSP.
 */
int dfr2d_create(const dfr2d_problem *p, int n_parts, int part, int device, dfr2d_handle **out);
/*
 * Partition-local set-up: the same, but `p` describes only a WINDOW of the mesh -- the elements
 * [k_offset, k_offset + p->K) of a mesh of K_global elements -- so that no process has to materialise the global
 * problem of an 8M-element run.  The window must contain the partition's own PartitionMap range and every element that
 * shares an edge (with the PerssonC0 limiter: a vertex) with it.  Inside `p`, element indices (edge_kL, edge_kR, the
 * rows of Jdet / Jinv / EToV / EtoEdge, the column stride of FaceNorm / IInII) are window-local; the edge table lists at
 * least every edge touching an own element; vertex ids (EToV) and p->NV stay global.  Host arrays of set_state /
 * get_state / get_field / plot_field then have the WINDOW's columns ([..][p->K]).  Results are bitwise those of
 * dfr2d_create on the global problem (tests/test_window.py).
 */
int dfr2d_create_window(const dfr2d_problem *p, int64_t K_global, int64_t k_offset, int n_parts, int part, int device,
                        dfr2d_handle **out);
void dfr2d_destroy(dfr2d_handle *h);
const char *dfr2d_last_error(const dfr2d_handle *h); /* h may be NULL: error of the last failed create on this thread */

/* c.Q shards <-> global [4][NpInt x K] (parallelism.go:61-98).  get_state writes only this partition's columns. */
int dfr2d_set_state(dfr2d_handle *h, const double *Q);
int dfr2d_get_state(dfr2d_handle *h, double *Q);

/* c.Time / c.StepCount (euler.go:175-182) as the next dfr2d_step sees them: restart from a saved state, or rewind. */
int dfr2d_set_clock(dfr2d_handle *h, double time, int64_t steps);

/* nsteps x { RK.Step; Time += GlobalDT; StepCount++ } (euler.go:177-182); stops early when finished.
 * info may be NULL (no host synchronisation at all). */
int dfr2d_step(dfr2d_handle *h, int nsteps, dfr2d_step_info *info);

/* Signed max of the Residual arrays per variable, as PrintUpdate reports it (euler.go:821-835), for the last step that
 * ran.  With the default element kernel the four maxima are reduced inside the last stage's launch and this call only
 * reads four scalars; the Residual register itself (euler.go:553-560) is then never materialised. */
int dfr2d_residual(dfr2d_handle *h, double maxR[4]);

/* Test hooks: RHSQ of stage rk evaluated on register `rk` ({c.Q,Q1..Q4}[rk], euler.go:422) without
 * advancing; register access; per-element fields. Global layouts, own columns only. */
int dfr2d_rhs(dfr2d_handle *h, int rk, double *RHS_out);
int dfr2d_set_register(dfr2d_handle *h, int reg, const double *Q);
int dfr2d_get_register(dfr2d_handle *h, int reg, double *Q);
int dfr2d_get_field(dfr2d_handle *h, int which, double *out /* [K] */);

/* Initial condition evaluated on the device instead of dfr2d_set_state: Euler.InitializeSolution (euler.go:728-794) --
 * InitializeFS / InitializeIVortex (initialization.go:50-83; beta=5, x0=5, y0=0 as carried in dfr2d_problem.vortex) /
 * the shock-tube split at x < 0.5 (euler.go:742-768, FSIn / FSOut) -- at the solution points of every own element,
 * X = 0.5(-(r+s) v0 + (1+r) v1 + (1+s) v2).  VX, VY = [nv] vertex coordinates, EToV = [K x 3] (global), R, S = [NpInt]
 * SolutionElement.R/S. */
int dfr2d_init_state(dfr2d_handle *h, int init_case, int64_t nv, const double *VX, const double *VY, const int32_t *EToV,
                     const double *R, const double *S);

/* Field read-back for plots: Euler.GetPlotField (model_problems/Euler2D/plot.go:14-86) for the flow functions that go
 * through FreeStream.GetFlowFunction (fluids.go:209-223: Density=0 .. Entropy=13) evaluated on c.Q on the device:
 * node values -> DFR.GraphInterp product (DG2D/dfr_startup.go:62-63) -> AverageGraphFieldVertices
 * (DG2D/graphics_support2.go:184-199) -> transpose -> float32 (what AVSFieldWriter.saveField stores,
 * DG2D/graphics_support.go:80-93).  graph_interp = [np_graph x NpInt] row-major with np_graph = 3(1+NpEdge)+NpInt;
 * out = [K x np_graph] (global element index; own rows only are written). */
int dfr2d_plot_field(dfr2d_handle *h, int flow_function, const double *graph_interp, int np_graph, float *out);
/* flow_function 100 = ShockFunction (plot.go:30-47): ModeAliasShockFinder.ShockIndicator of the element's density
 * (DG2D/dfr_shock_capturing.go:189-235; Kappa of c.ShockFinder, or 2 without the limiter) at every node, then the same
 * GraphInterp / vertex-average / float32 path.
 * The two epsilon fields bypass the interpolation in the reference (plot.go:48-53): out = [NpFlux x K] doubles, global
 * column index, own columns written.  c0 = 0: EpsilonDissipation = the element's scalar epsilon at every RT point
 * (dissipation.go:377-390); c0 = 1: EpsilonDissipationC0 = Bary . vertex epsilon as the last stage left it (:392-395). */
int dfr2d_epsilon_field(dfr2d_handle *h, int c0, double *out);

/* The remaining GetPlotField cases (plot.go:54-77; fluids.go:227-234): flow_function 200..203 = XGradientDensity ..
 * XGradientEnergy, 300..303 = YGradient*: GetSolutionGradientUsingRTElement(-1, n, c.Q, ...) (euler.go:864-918) -- the RT
 * gradient of conserved variable n = flow_function % 100 of the CURRENT c.Q at the interior RT points and of the
 * EdgeQValues store at the edge points, Div . (DXMetric|DYMetric (.) U), not interpolated: out = [NpFlux x K] doubles,
 * global column index, own columns written.  The store holds the owner side's Q_Face of the last stage that ran
 * (edges.go:344-350: the input of stage 5 of the last step; zeros before the first step).  The device reuses Q_Face for
 * the next stage, so these values are kept only on request: dfr2d_capture_edge_values(h, 1, ...) makes every following
 * step copy Q_Face (own and ghost columns) aside before its last update, at the cost of one device copy per step; turn it
 * on before the step whose fields are plotted and off again afterwards (on = 0 keeps what was captured).  FaceNormX/Y =
 * DFR.FaceNorm[0|1] exactly as the problem struct given at create carried them; they are read once, and may be
 * NULL on a handle with the limiter (which keeps them already).  Needs dfr2d_problem.Div. */
int dfr2d_capture_edge_values(dfr2d_handle *h, int on, const double *FaceNormX, const double *FaceNormY);
int dfr2d_gradient_field(dfr2d_handle *h, int flow_function, double *out);

/* ---- plumbing for one-process-per-GPU hosts (torch.distributed / NCCL) ---------------------
 * The library never calls a collective itself: per stage the host moves the halo bytes and
 * max-reduces two doubles, between the three phases below.  Single-partition users only need
 * dfr2d_step().  All work is enqueued on the stream given to dfr2d_set_stream (default 0). */
int dfr2d_set_stream(dfr2d_handle *h, void *cuda_stream);
int dfr2d_partition_range(const dfr2d_handle *h, int64_t *k_begin, int64_t *k_end);
/* doubles sent to / received from every partition each stage (arrays of n_parts) */
int dfr2d_halo_counts(const dfr2d_handle *h, int64_t *send_counts, int64_t *recv_counts);
/* device buffers, contiguous, ordered by peer partition */
int dfr2d_halo_buffers(dfr2d_handle *h, void **send_dev, void **recv_dev);
/* device address of {max wave speed, max viscous wave speed} of the stage in flight: MAX-allreduce in place */
int dfr2d_wavespeed_buffer(dfr2d_handle *h, void **dev_two_doubles);
int dfr2d_stage_prepare(dfr2d_handle *h, int rk); /* edge interpolation if stale + pack the halo send buffer */
int dfr2d_stage_edges(dfr2d_handle *h, int rk);   /* unpack halo + numerical edge fluxes + wave-speed maxima */
int dfr2d_stage_update(dfr2d_handle *h, int rk);  /* divergence, dt, SSP-RK update (+ next stage's interpolation) */
/* With the PerssonC0 limiter a stage has three exchange points instead of one:
 *   dfr2d_stage_sensor  [sensor, element->vertex max merge, pack]           -> exchange DFR2D_XCHG_VERTEX
 *   dfr2d_stage_prepare [unpack(max), vertex->element, limiter, interp, pack] -> exchange DFR2D_XCHG_EDGE
 *   dfr2d_stage_edges   [unpack, edge flux, RT gradient x epsilon, pack]    -> exchange DFR2D_XCHG_DISS
 *   dfr2d_stage_visc    [unpack, viscous edge flux]  -> MAX-allreduce of the wave-speed pair -> dfr2d_stage_update
 * (phases P1-P8 of StepWorker, euler.go:574-651; the goroutines' shared memory becomes these messages).
 * stage_sensor / stage_visc are no-ops without the limiter, so a host may always call all five. */
enum { DFR2D_XCHG_EDGE = 0, DFR2D_XCHG_VERTEX = 1, DFR2D_XCHG_DISS = 2 };
int dfr2d_stage_sensor(dfr2d_handle *h, int rk);
int dfr2d_stage_visc(dfr2d_handle *h, int rk);
/* Optional overlap hook: the numerical flux of the edges that touch no ghost column does not need the halo.  Call it
 * after posting the EDGE exchange and before waiting for it; dfr2d_stage_edges then only unpacks and evaluates the
 * boundary and cut edges.  Results are identical with or without this call. */
int dfr2d_stage_edges_interior(dfr2d_handle *h, int rk);
/* per-peer doubles (arrays of n_parts; every exchange is symmetric) and device buffers of exchange `which`;
 * which = DFR2D_XCHG_EDGE gives the same answers as dfr2d_halo_counts / dfr2d_halo_buffers */
int dfr2d_exchange_counts(const dfr2d_handle *h, int which, int64_t *send_counts, int64_t *recv_counts);
int dfr2d_exchange_buffers(dfr2d_handle *h, int which, void **send_dev, void **recv_dev);
int dfr2d_step_finish(dfr2d_handle *h, dfr2d_step_info *info); /* after stage 4: read back time/steps (info may be NULL) */

/* ---- partition-to-partition exchange over peer memory (NVLink P2P), no host and no collective library in the loop ----
 * Replaces the shared memory through which the reference's partition goroutines read each other's Q_Face / edge store
 * (RungeKutta5SSP.Step, euler.go:408-418; calculateSharedEdgeFlux, edges.go:379-411) and the serial max over partitions
 * of calculateGlobalDT (euler.go:951-955).  Every partition owns a MAILBOX in device memory (receive buffers of the
 * three exchanges, arrival flags, a wave-speed inbox).  Once partitions are connected, the pack kernels store their
 * messages straight into the partner's mailbox and publish the stage's sequence number in its arrival flag; the unpack
 * kernels spin on the flag (csrc/dfr2d_peer.cuh).  After connecting, dfr2d_step() works on every partition of a
 * multi-partition run -- each owner just calls it -- and the stage calls exchange by themselves (the host must then NOT
 * move the halo buffers; it calls dfr2d_stage_wave instead of max-reducing dfr2d_wavespeed_buffer).
 *
 *   one process per partition (torchrun / MPI style): dfr2d_peer_export on every partition, all-gather the blobs by any
 *       means, dfr2d_peer_connect(h, blobs, n_parts) everywhere (cudaIpcOpenMemHandle under the hood), barrier, step.
 *   one process owning all partitions (the Go controller goroutine): dfr2d_multi_step connects them itself
 *       (cudaDeviceEnablePeerAccess) and issues the stages of all partitions from the calling thread. */
#define DFR2D_PEER_BLOB_BYTES 1280
int dfr2d_peer_export(dfr2d_handle *h, void *blob /* [DFR2D_PEER_BLOB_BYTES] */);
int dfr2d_peer_connect(dfr2d_handle *h, const void *blobs /* [n_parts][DFR2D_PEER_BLOB_BYTES], index = partition */, int n_parts);
/* back to the host-moved exchange of the plain stage API (on = 0) or to the mailboxes again (on = 1); mappings stay */
int dfr2d_peer_enable(dfr2d_handle *h, int on);
int dfr2d_stage_wave(dfr2d_handle *h, int rk);   /* connected hosts: put + gather of the wave-speed pair, between visc and update */

/* hs[g] = dfr2d_create(p, n, g, device_g, ...) for g = 0..n-1, all in this process (several partitions may share a device).
 * dfr2d_multi_step runs nsteps x { 5 stages } over all partitions with the mailbox protocol above; the numerical flux of
 * the edges that touch no ghost column runs while the halo is in flight.  No host synchronisation inside the call;
 * `info` (may be NULL) is read back from partition 0 at the end.  Results are bitwise those of a single partition. */
int dfr2d_multi_step(dfr2d_handle **hs, int n, int nsteps, dfr2d_step_info *info);
/* The same through one call each (the shape SURVEY.md 8(b) gives the boundary: one create for the run, the fan-out over the
 * GPUs inside the library): hs_out[g] = partition g of n_parts on devices[g]; devices = NULL spreads them round-robin over
 * the visible devices.  On failure nothing is left allocated and dfr2d_last_error(NULL) has the message.
 * dfr2d_multi_residual = the signed maximum over the partitions of dfr2d_residual (PrintUpdate's loop, euler.go:823-829). */
int dfr2d_multi_create(const dfr2d_problem *p, int n_parts, const int *devices, dfr2d_handle **hs_out);
void dfr2d_multi_destroy(dfr2d_handle **hs, int n);
int dfr2d_multi_residual(dfr2d_handle **hs, int n, double maxR[4]);

/* c.Q <-> all partitions of this process; the copies of the n devices cross PCIe concurrently */
int dfr2d_multi_set_state(dfr2d_handle **hs, int n, const double *Q);
int dfr2d_multi_get_state(dfr2d_handle **hs, int n, double *Q);
/* One profiled step: ms_out[n][5 stages][6 phases] = CUDA-event duration of {sensor+prepare+pack/put, interior edges,
 * halo wait + boundary/cut edges (+ RT gradient), viscous edges, wave put+gather, element update} on each partition. */
int dfr2d_multi_step_profile(dfr2d_handle **hs, int n, float *ms_out);

/* ---- host-only partition plan (no CUDA): the bookkeeping dfr2d_create performs for (n_parts, part), exposed so
 * the decomposition can be verified bit-exactly on a CPU-only machine.  Mirrors utils.PartitionMap +
 * PartitionEdgesByK (parallelism.go:179-190) plus the ghost/halo lists that replace the goroutines' shared memory. */
typedef struct dfr2d_plan dfr2d_plan;
int dfr2d_plan_create(const dfr2d_problem *p, int n_parts, int part, dfr2d_plan **out);
int dfr2d_plan_create_window(const dfr2d_problem *p, int64_t K_global, int64_t k_offset, int n_parts, int part, dfr2d_plan **out);
void dfr2d_plan_destroy(dfr2d_plan *pl);
/* out = {k_begin, k_end, n_ghost_columns, padded_columns, n_local_edges, padded_edges, n_cut_edges, n_boundary_edges} */
int dfr2d_plan_sizes(const dfr2d_plan *pl, int64_t out[8]);
/* per local edge: owner / neighbour column (ghost columns >= k_end-k_begin; neighbour < 0 on boundaries), packed
 * meta (numL | numR<<2 | bc<<4), global edge index; etoe = [3][padded_columns] edge slot (or -1-slot if not owner) */
int dfr2d_plan_edges(const dfr2d_plan *pl, int32_t *kL, int32_t *kR, int32_t *meta, int64_t *global_edge, int32_t *etoe);
/* per peer counts (doubles), global element id of every ghost column, and per cut edge the send (column,row0) and
 * receive (ghost column,row0) addresses inside Q_Face; message order = (peer, global edge index) */
int dfr2d_plan_halo(const dfr2d_plan *pl, int64_t *send_counts, int64_t *recv_counts, int64_t *ghost_global,
                    int32_t *send_elem, int32_t *send_row0, int32_t *recv_col, int32_t *recv_row0);
/* dissipation only: shared vertices grouped by peer (ascending global vertex id within a peer); counts are doubles
 * per peer (2 per vertex: sigma, epsilon).  vertex_ids may be NULL to query the counts first. */
int dfr2d_plan_vertices(const dfr2d_plan *pl, int64_t *counts, int32_t *vertex_ids);

/* Host-only helper for partition quality (no CUDA): reverse Cuthill-McKee order of the elements over the edge table
 * (elements sharing an edge are adjacent).  order[new index] = old element.  Renumber the mesh with it before
 * NewDFR2D so that the contiguous PartitionMap.Split1D ranges (utils/parallel_utils.go:172-192) are spatially compact;
 * the reference wraps METIS for the same purpose in 3D only (DG3D/mesh/partitioner/mesh_partitioner.go). */
int dfr2d_rcm_order(int64_t K, int64_t NE, const int32_t *edge_kL, const int32_t *edge_kR, const int32_t *edge_nconn,
                    int32_t *order);

/* Host-only (no CUDA), the stronger of the two orderings on the reference's airfoil meshes: order[new index] = old element
 * along a Hilbert curve through the element centroids, with each coordinate replaced by its rank so that the curve
 * resolves surface cells and far field alike.  Takes what the host holds right after the mesh is read (EToV = [K x 3],
 * VX, VY = [NV]).  Cut edges of the contiguous Split1D ranges at 8 partitions: nacaAirfoil-base 10,983 as numbered by the
 * mesh generator, 1,020 after dfr2d_rcm_order, 606 after this; mesh_NACA0012_inv 2,973 / 1,104 / 582. */
int dfr2d_hilbert_order(int64_t K, int64_t NV, const int32_t *EToV, const double *VX, const double *VY, int32_t *order);

/* Host-only (no CUDA): the operator table the tensor-core gradient kernel (k_grad_mma, DFR2D_GRAD_KERNEL=2) stages in
 * shared memory -- DFR.FluxElement.Div (raviart_thomas_element.go:249-297) cut to the rows and metric blocks that
 * GetSolutionGradientUsingRTElement (euler.go:864-918) consumes, in DMMA.8x8x4 A-fragment lane order, followed by the
 * Bary rows of those RT points.  Writes min(cap, needed) doubles to `out` and returns the needed count (<0: bad order).
 * Exposed so that the layout can be checked on the CPU (tests/test_grad_mma_layout.py). */
int64_t dfr2d_grad_mma_table(int N, const double *Div, const double *Bary, double *out, int64_t cap);

/* Host-only: the A-fragment table [DivInt | Vinv | V] of the tensor-core dissipation element kernel (k_elem_mma_diss,
 * DFR2D_DISS_ELEM_KERNEL=3), same conventions as dfr2d_grad_mma_table; for the CPU layout check. */
int64_t dfr2d_mma_diss_table(int N, const double *DivInt, const double *Vinv, const double *V, double *out, int64_t cap);

/* number of kernels this handle has launched (for the benchmark's gpu_launches claim) */
int64_t dfr2d_launch_count(const dfr2d_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* DFR2D_H */
