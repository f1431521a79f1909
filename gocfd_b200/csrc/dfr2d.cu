// C ABI (include/dfr2d.h) of the B200 DFR2D library: problem upload, partition extraction,
// stage sequencing.  All arithmetic lives in the kernels; this file never touches solution data
// on the host except to copy it in or out.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include <unistd.h>

#include <cuda.h>          // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link)

#include "dfr2d_kernels.cuh"
#include "dfr2d_diss_kernels.cuh"
#include "dfr2d_elem_mma.cuh"
#include "dfr2d_elem_pipe.cuh"
#include "dfr2d_elem_ws.cuh"
#include "dfr2d_grad_mma.cuh"
#include "dfr2d_elem_mma_diss.cuh"
#include "dfr2d_edge_ws.cuh"

using namespace dfr2d;

static bool grad_table_for(int N, const double *Div, const double *Bary, std::vector<double> &tb);

static thread_local std::string g_create_error;

struct dfr2d_handle {
    int N = 0, device = 0, nParts = 1, part = 0;
    int64_t Kglobal = 0, k0 = 0, k1 = 0;
    int64_t winOff = 0;               // dfr2d_create_window: global id of the first element the problem arrays describe
    int64_t hostOff = 0, hostPitch = 0;   // host-side [rows][hostPitch] arrays: this partition's columns start at hostOff
    int K = 0, G = 0, Kp = 0, NE = 0, NEp = 0, NV = 0, NBP = 0;
    int NpInt = 0, NpEdge = 0, NpFlux = 0;
    Phys ph{};
    cudaStream_t stream = 0;
    std::string err;
    std::vector<void *> allocs;
    std::vector<double> opsHost;      // Ops<N> image for constant memory
    uint64_t opsFp = 0;               // fingerprint of opsHost (which table is resident on the device)
    // state
    double *q[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // q[4] aliases q[1] (Q1 is dead after stage 1)
    double *R = nullptr, *qface = nullptr, *eflux = nullptr, *agg = nullptr, *DT = nullptr, *rhsScratch = nullptr;
    double *Jdet = nullptr, *Jinv = nullptr, *IInII = nullptr;
    int *etoe = nullptr;
    // edges
    int *ekL = nullptr, *ekR = nullptr, *emeta = nullptr, *bndList = nullptr;
    int nBnd = 0;
    int edgeSplit = 1;                // 1: lean interior-edge kernel + boundary-list kernel; 0: one generic kernel
    double *enx = nullptr, *eny = nullptr, *eoohk = nullptr, *bpx = nullptr, *bpy = nullptr;
    // dissipation
    DissBuffers ds{};
    // halo
    std::vector<int64_t> sendCounts, recvCounts;
    int64_t sendTotal = 0, recvTotal = 0;
    double *sendBuf = nullptr, *recvBuf = nullptr;
    int *sendElem = nullptr, *sendRow0 = nullptr, *recvCol = nullptr, *recvRow0 = nullptr;
    int *cutSlot = nullptr, *cutSide = nullptr;
    int nSendEdges = 0, nRecvEdges = 0;
    int ePer = 0;                     // doubles per cut edge in the Q_Face message (4 NpEdge, +3 vertex eps with dissipation)
    // dissipation exchanges (multi-partition): shared-vertex max merge and DissX/DissY edge rows
    std::vector<int64_t> vtxCounts, dissCounts;
    int nVtx = 0;
    int *vtxList = nullptr;
    double *vSendBuf = nullptr, *vRecvBuf = nullptr, *dSendBuf = nullptr, *dRecvBuf = nullptr;
    // time loop
    DevScalars *sc = nullptr;
    DevScalars *scHost = nullptr;     // pinned
    long long stageCounter = 0, stepIndex = 0, launches = 0;
    bool qfaceValid = false;          // Q_Face holds the interpolation of the next stage's input register
    // gradient plot fields (plot.go:54-77): the EdgeQValues store = Q_Face of the last stage 5, captured on request
    bool captureEdgeQ = false;
    double *qfaceSaved = nullptr;     // [4][3NpEdge][Kp], zero until the first captured step (the reference's store starts at 0)
    double *nxkPlot = nullptr, *nykPlot = nullptr;   // [3][Kp] element face normals (the limiter's copies when it is on)
    bool haveDiv = false;             // dfr2d_problem.Div was given
    bool interiorDone = false;        // the interior-edge kernel of the stage in flight has been launched (overlap with the halo)
    int edgeBlocks = 0;
    bool smemAttrSet = false;
    int pfTiles = 0;
    int elemKernel = 4;               // 1: row-per-thread DFMA, 2: split-row DFMA, 3: DMMA, 4: pipelined DMMA (default)
    double *mmaFrags = nullptr;
    int gradKernel = 1;               // 1: constant-operand DFMA k_grad, 3: pipelined persistent DMMA k_grad_pipe, 4: k_grad_ws
                                      // (DFR2D_GRAD_KERNEL; 2 was round 1's one-tile-per-CTA DMMA kernel, retired)
    double *gradTable = nullptr, *gradMxy = nullptr;
    int gradMG = 3;                   // m-tiles per accumulation group of k_grad_pipe (DFR2D_GRAD_MG = 2 | 3)
    int dissElemKernel = 1;           // element kernel of the PerssonC0 path: 1 = k_elem<N,true> (DFMA), 3 = k_elem_mma_diss, 5 = k_elem_ws<N,8,true>
                                      // (DMMA, opt-in until measured; DFR2D_DISS_ELEM_KERNEL)
    double *mmaDissFrags = nullptr;
    int mmaDissGrid = 0;
    int gradSkewNs = 0;               // start delay of k_grad_pipe's second warp group (DFR2D_GRAD_SKEW_NS)
    bool gradAttrSet = false, gradWsAttrSet = false;
    int sms = 148, mmaGrid = 148;
    int pipeOcc[3] = {0, 0, 0};
    int wsStages = 0;                // DFR2D_WS_STAGES override of the ring depth of kernel 5
    void *tmaps = nullptr;            // DFR2D_WS_TMA=1: four CUtensorMap (q0, q2, q3, R) for the TMA variant of kernel 5
    int edgeViscFused = 1;            // PerssonC0: viscous edge flux inside the edge kernels (DFR2D_EDGE_VISC_FUSED=0: k_visc_edge)
    bool dissWsAttrSet = false;
    int dissPrefetch = 0;             // k_elem_mma_diss: L2 prefetch of the next tile (DFR2D_DISS_PREFETCH, measured slower)
    int wsCW = 8;                    // consumer warps of kernel 5: 8 (two groups) or 12 (three groups, DFR2D_WS_CW)
    // single partition: the short boundary-edge list kernel (BC transcendentals, a few dozen CTAs, 14 us at C2) runs on a
    // second stream beside the interior-edge kernel instead of after it (DFR2D_EDGE_OVERLAP=0 turns it off)
    cudaStream_t auxStream = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    int edgeOverlap = 1;
    int edgeWs = 0, edgeWsStages = 0; // interior edges by the warp-specialised k_edge_ws (DFR2D_EDGE_WS, DFR2D_EDGE_WS_STAGES)
    bool edgeWsAttrSet = false;
    bool resInScalars = false;       // the residual maxima of the last step are in DevScalars::resMax (kernel 5), not in R
    int wsSplit = 1;                 // kernel 5 with four extra interpolation warps (k_elem_ws<N,8,false,true>, DFR2D_WS_SPLIT)
    int edgePPT = 0;
    // peer exchange (dfr2d_peer.cuh): one allocation = the three receive buffers + arrival flags + wave inbox, so that a
    // single IPC handle / peer pointer gives a partner everything it writes
    unsigned long long *mailbox = nullptr;
    size_t mailboxWords = 0;
    int64_t mbOff[3] = {0, 0, 0}, mbFlags = 0, mbWaveFlag = 0, mbWaveIn = 0;     // word offsets inside the mailbox
    bool connected = false;
    PutTab *putTab[3] = {nullptr, nullptr, nullptr};      // device
    WaitTab *waitTab[3] = {nullptr, nullptr, nullptr};    // device
    unsigned int *putDone = nullptr;                      // device [3]
    WaveTab waveTab{};
    std::vector<void *> ipcMapped;
    double kappaGiven = 0.0;          // ip.Kappa as given: c.ShockFinder of the plot path (euler.go:82)
    void *scratch = nullptr;          // call-spanning scratch of residual / plot_field / init_state / rhs (scratch_reserve)
    size_t scratchBytes = 0;
};

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return 2;                                                                                   \
        }                                                                                               \
    } while (0)

template <typename T> static int dev_alloc(dfr2d_handle *h, T **p, size_t n) {
    void *d = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc(&d, n * sizeof(T));
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return 2;
    }
    h->allocs.push_back(d);
    *p = (T *)d;
    return 0;
}

template <typename T> static int dev_upload(dfr2d_handle *h, T **p, const std::vector<T> &v, size_t padTo = 0) {
    size_t n = std::max(v.size(), padTo);
    if (int rc = dev_alloc(h, p, n)) return rc;
    CK(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static void split1d(int64_t maxIndex, int nparts, int n, int64_t *lo, int64_t *hi) {
    // utils/parallel_utils.go:172-192
    int64_t npart = maxIndex / nparts, rem = maxIndex % nparts, startAdd = 0, endAdd = 0;
    if (rem != 0) {
        if (n + 1 > rem) { startAdd = rem; endAdd = 0; } else { startAdd = n; endAdd = 1; }
    }
    *lo = n * npart + startAdd;
    *hi = *lo + npart + endAdd;
}

static int bucket_of(int64_t k, int64_t maxIndex, int nparts) {
    int b = (int)((double)(nparts * k) / (double)maxIndex);
    b = std::min(std::max(b, 0), nparts - 1);
    for (;;) {
        int64_t lo, hi;
        split1d(maxIndex, nparts, b, &lo, &hi);
        if (k < lo) b--; else if (k >= hi) b++; else return b;
    }
}

template <int N> static void pack_ops(const dfr2d_problem *p, std::vector<double> &out) {
    out.assign(kOpsDoubles, 0.0);
    Ops<N> *o = reinterpret_cast<Ops<N> *>(out.data());
    constexpr int NI = Dim<N>::NpInt, NF = Dim<N>::NpFlux, NF3 = Dim<N>::NF3;
    auto cp = [](double *dst, const double *src, size_t n) { if (src) memcpy(dst, src, n * sizeof(double)); };
    cp(&o->FEI[0][0], p->FluxEdgeInterp, (size_t)NF3 * NI);
    cp(&o->DivInt[0][0], p->DivInt, (size_t)NI * NF);
    cp(&o->V[0][0], p->V, (size_t)NI * NI);
    cp(&o->Vinv[0][0], p->Vinv, (size_t)NI * NI);
    cp(&o->M[0][0], p->MassMatrix, (size_t)NI * NI);
    cp(&o->D[0][0], p->D, (size_t)NI * NI);
    cp(&o->P[0][0], p->P, (size_t)NI * NI);
    cp(&o->mf[0], p->ModeFilter, (size_t)NI);
    cp(&o->Bary[0][0], p->Bary, (size_t)NF * 3);
    cp(&o->Div[0][0], p->Div, (size_t)NF * NF);
}

// The operator tables live in __constant__ memory, which exists once per DEVICE.  Per device we remember a
// fingerprint of the resident table, the stream that uploaded it and an event behind the upload:
//   - same table, same stream: nothing to do (the common case: every partition of a run has the same operators);
//   - same table, other stream: that stream waits for the upload event (free once it has completed);
//   - another table (handles of different order interleaved on one device): the device is drained first, because kernels
//     of other streams may still be reading the old table, then the new one is uploaded.
// Every stage_* entry calls this, so the stage API is safe for any interleaving of handles.
struct OpsResident {
    uint64_t fp = 0;
    bool valid = false;
    cudaStream_t stream = 0;
    cudaEvent_t ev = nullptr;
};
static std::mutex g_ops_mutex;
static std::unordered_map<int, OpsResident> g_ops_resident;

static uint64_t ops_fingerprint(const std::vector<double> &v) {
    uint64_t hsh = 1469598103934665603ull;               // FNV-1a over the bit patterns
    for (double d : v) {
        uint64_t b;
        memcpy(&b, &d, sizeof(b));
        hsh = (hsh ^ b) * 1099511628211ull;
    }
    return hsh ? hsh : 1;
}

static int ensure_ops(dfr2d_handle *h) {
    std::lock_guard<std::mutex> lock(g_ops_mutex);
    OpsResident &r = g_ops_resident[h->device];
    if (r.valid && r.fp == h->opsFp) {
        if (r.stream != h->stream) CK(cudaStreamWaitEvent(h->stream, r.ev, 0));
        return 0;
    }
    if (r.valid) CK(cudaDeviceSynchronize());
    if (!r.ev) CK(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
    CK(cudaMemcpyToSymbolAsync(c_ops_raw, h->opsHost.data(), kOpsDoubles * sizeof(double), 0, cudaMemcpyHostToDevice,
                               h->stream));
    CK(cudaEventRecord(r.ev, h->stream));
    r.fp = h->opsFp;
    r.stream = h->stream;
    r.valid = true;
    return 0;
}

extern "C" const char *dfr2d_last_error(const dfr2d_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" void dfr2d_destroy(dfr2d_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (void *p : h->ipcMapped) cudaIpcCloseMemHandle(p);
    for (void *p : h->allocs) cudaFree(p);
    if (h->scratch) cudaFree(h->scratch);
    if (h->scHost) cudaFreeHost(h->scHost);
    if (h->auxStream) cudaStreamDestroy(h->auxStream);
    if (h->evFork) cudaEventDestroy(h->evFork);
    if (h->evJoin) cudaEventDestroy(h->evJoin);
    delete h;
}

// Host-only partition plan: which elements/edges/ghost columns a partition holds and what crosses the
// cut each stage.  Pure integer/geometry bookkeeping, no CUDA -- also exported (dfr2d_plan_*) so the
// multi-partition logic can be tested on a CPU-only machine.
struct dfr2d_plan {
    int N = 0, nParts = 1, part = 0;
    int64_t Kglobal = 0, k0 = 0, k1 = 0;
    int64_t kOff = 0;                      // window mode: the problem arrays describe elements [kOff, kOff + p->K) only
    int K = 0, G = 0, Kp = 0, NE = 0, NEp = 0, NV = 0, NBP = 0;
    std::vector<int> ekL, ekR, emeta, etoe, sendElem, sendRow0, recvCol, recvRow0, bndList;
    std::vector<int> cutSlot, cutSide;     // per cut edge (message order): local edge slot, my side (0 = I own it)
    std::vector<double> enx, eny, eoohk, eooLen, bpx, bpy, Jdet, Jinv, IInII;
    std::vector<int64_t> ghostGlobal, edgeGlobal, sendCounts, recvCounts;
    int nSendEdges = 0, nRecvEdges = 0;
    int64_t sendTotal = 0, recvTotal = 0;
    int ePer = 0;
    std::vector<int> vtxList;              // shared vertices grouped by peer, ascending vertex id within a peer
    std::vector<int64_t> vtxCounts;        // doubles per peer in the vertex message (2 per listed vertex)
    std::string err;
};

static int build_plan(const dfr2d_problem *p, dfr2d_plan &pl) {
    const int N = p->N;
    pl.N = N;
    const int NEd = N + 2;
    // ---- partition (utils.PartitionMap, parallelism.go:179-190) ------------------------------
    // The problem arrays may describe the whole mesh (kOff = 0, Kglobal = p->K) or a WINDOW of it: elements
    // [kOff, kOff + p->K) of a mesh of Kglobal elements, which must contain the partition's own range and every element
    // that shares an edge (with the limiter: a vertex) with it.  Element indices inside the arrays (edge_kL / edge_kR,
    // rows of Jdet, EToV, ...) are window-local; vertex ids stay global.  Everything below works on GLOBAL element ids
    // kg and indexes the arrays with kg - kOff.
    if (pl.Kglobal <= 0) pl.Kglobal = p->K;
    const int64_t kOff = pl.kOff, Kg = pl.Kglobal, Kw = p->K;
    split1d(Kg, pl.nParts, pl.part, &pl.k0, &pl.k1);
    const int64_t k0 = pl.k0, k1 = pl.k1;
    pl.K = (int)(k1 - k0);
    if (k0 < kOff || k1 > kOff + Kw) { pl.err = "the window does not contain the partition's own element range"; return 1; }
    auto mine = [&](int64_t k) { return k >= k0 && k < k1; };
    auto eL = [&](int64_t e) -> int64_t { return (int64_t)p->edge_kL[e] + kOff; };
    auto eR = [&](int64_t e) -> int64_t { return (p->edge_nconn[e] == 2) ? (int64_t)p->edge_kR[e] + kOff : -1; };

    // local edges = every edge touching an owned element; remote sides become ghost columns
    struct LE { int64_t ge; int l, r; };
    std::vector<LE> led;
    led.reserve((size_t)(pl.nParts == 1 ? p->NE : 2 * (p->NE / pl.nParts) + 1024));
    std::unordered_map<int64_t, int> ghostOf;
    auto &ghostGlobal = pl.ghostGlobal;
    auto local_col = [&](int64_t kg) -> int {
        if (mine(kg)) return (int)(kg - k0);
        auto it = ghostOf.find(kg);
        if (it != ghostOf.end()) return pl.K + it->second;
        int g = (int)ghostGlobal.size();
        ghostOf.emplace(kg, g);
        ghostGlobal.push_back(kg);
        return pl.K + g;
    };
    for (int64_t e = 0; e < p->NE; e++) {
        const int64_t kl = eL(e), kr = eR(e);
        if (!(mine(kl) || (kr >= 0 && mine(kr)))) continue;
        led.push_back({e, 0, 0});
    }
    // ghosts are numbered in global-edge order so that both sides of a cut agree on the message order
    for (auto &le : led) {
        const int64_t e = le.ge;
        le.l = local_col(eL(e));
        le.r = (p->edge_nconn[e] == 2) ? local_col(eR(e)) : -1;
    }
    pl.G = (int)ghostGlobal.size();
    pl.Kp = ((pl.K + pl.G + 31) / 32) * 32;
    // default: sort by the owner-side column so that owner-side loads of neighbouring threads coalesce.
    // DFR2D_EDGE_SORT=1 sorts by (owner's local edge number, owner column) instead: one Q_Face row per warp on the
    // owner side.
    {
        const char *ev = getenv("DFR2D_EDGE_SORT");
        const bool byEdgeNumber = (ev && atoi(ev) == 1);     // measured: no net gain on B200 (k_edge slower, k_elem faster)
        const int32_t *numL = p->edge_numL;
        if (byEdgeNumber)
            std::stable_sort(led.begin(), led.end(), [numL](const LE &a, const LE &b) {
                const int na = numL[a.ge], nb = numL[b.ge];
                return na != nb ? na < nb : a.l < b.l;
            });
        else
            std::stable_sort(led.begin(), led.end(), [](const LE &a, const LE &b) { return a.l < b.l; });
    }
    pl.NE = (int)led.size();
    pl.NEp = ((pl.NE + 31) / 32) * 32;
    pl.NV = (int)p->NV;
    const int Kp = pl.Kp, K = pl.K;

    std::unordered_map<int64_t, int> slotOfGlobalEdge;
    if (pl.nParts > 1) slotOfGlobalEdge.reserve(led.size() * 2);
    std::vector<int> slotOfEdgeDense;
    if (pl.nParts == 1) slotOfEdgeDense.assign((size_t)p->NE, -1);
    auto &ekL = pl.ekL; auto &ekR = pl.ekR; auto &emeta = pl.emeta;
    auto &enx = pl.enx; auto &eny = pl.eny; auto &eoohk = pl.eoohk; auto &eooLen = pl.eooLen;
    ekL.assign(pl.NE, 0); ekR.assign(pl.NE, 0); emeta.assign(pl.NE, 0);
    enx.assign(pl.NE, 0.0); eny.assign(pl.NE, 0.0); eoohk.assign(pl.NE, 0.0); eooLen.assign(pl.NE, 0.0);
    pl.edgeGlobal.assign(pl.NE, 0);
    std::unordered_map<int64_t, int64_t> bpOfEdge;
    for (int64_t b = 0; b < p->NBP; b++) bpOfEdge.emplace(p->bp_edge[b], b);
    auto &bpx = pl.bpx; auto &bpy = pl.bpy;
    int nbp = 0;
    const double np12 = (double)((N + 1) * (N + 1));
    for (int s = 0; s < pl.NE; s++) {
        const int64_t e = led[s].ge;
        pl.edgeGlobal[s] = e;
        if (pl.nParts == 1) slotOfEdgeDense[e] = s; else slotOfGlobalEdge.emplace(e, s);
        const int64_t kl = eL(e) - kOff;          // window row of the owner
        const int numL = p->edge_numL[e], numR = p->edge_numR[e];
        ekL[s] = led[s].l;
        emeta[s] = (numL & 3) | ((numR & 3) << 2) | ((p->edge_bc[e] & 15) << 4);
        enx[s] = p->FaceNormX[kl + p->K * numL];
        eny[s] = p->FaceNormY[kl + p->K * numL];
        const double hK = p->EdgeLenMax[kl] / np12;          // DFR.GetHk (dfr_startup.go:89-98)
        eoohk[s] = 1.0 / hK;
        eooLen[s] = 1.0 / p->edge_len[e];
        if (led[s].r >= 0) {
            ekR[s] = led[s].r;
        } else {
            auto it = bpOfEdge.find(e);
            if (it == bpOfEdge.end()) {
                pl.err = "boundary edge without edge-point coordinates (bp_edge)";
                return 3;
            }
            ekR[s] = -1 - nbp;
            pl.bndList.push_back(s);
            for (int i = 0; i < NEd; i++) {
                bpx.push_back(p->bp_x[it->second * NEd + i]);
                bpy.push_back(p->bp_y[it->second * NEd + i]);
            }
            nbp++;
        }
    }
    pl.NBP = nbp;
    auto slot_of = [&](int64_t e) -> int {
        return pl.nParts == 1 ? slotOfEdgeDense[e] : slotOfGlobalEdge.at(e);
    };

    // element arrays (SoA, stride Kp)
    auto &Jdet = pl.Jdet; auto &Jinv = pl.Jinv; auto &IInII = pl.IInII; auto &etoe = pl.etoe;
    Jdet.assign((size_t)Kp, 1.0); Jinv.assign((size_t)4 * Kp, 0.0); IInII.assign((size_t)3 * Kp, 0.0);
    etoe.assign((size_t)3 * Kp, 0);
    for (int k = 0; k < K; k++) {
        const int64_t kg = k0 + k;
        const int64_t kw = kg - kOff;
        Jdet[k] = p->Jdet[kw];
        for (int c = 0; c < 4; c++) Jinv[(size_t)c * Kp + k] = p->Jinv[kw * 4 + c];
        for (int le = 0; le < 3; le++) {
            IInII[(size_t)le * Kp + k] = p->IInII[kw + p->K * le];
            const int64_t e = p->EtoEdge[kw * 3 + le];
            const int s = slot_of(e);
            etoe[(size_t)le * Kp + k] = (eL(e) == kg) ? s : -1 - s;
        }
    }

    // ---- halo lists: for every cut edge the 4 x NpEdge Q_Face values of each side cross once per stage
    pl.sendCounts.assign(pl.nParts, 0);
    pl.recvCounts.assign(pl.nParts, 0);
    auto &sendElem = pl.sendElem; auto &sendRow0 = pl.sendRow0; auto &recvCol = pl.recvCol; auto &recvRow0 = pl.recvRow0;
    if (pl.nParts > 1) {
        struct Cut { int peer; int64_t ge; int myCol, myNum, ghCol, ghNum, slot, side; };
        std::vector<Cut> cuts;
        for (int s = 0; s < pl.NE; s++) {
            const int64_t e = led[s].ge;
            if (p->edge_nconn[e] != 2) continue;
            const int64_t kl = eL(e), kr = eR(e);
            if (mine(kl) && mine(kr)) continue;
            const bool lMine = mine(kl);
            const int64_t remote = lMine ? kr : kl;
            Cut c;
            c.peer = bucket_of(remote, Kg, pl.nParts);
            c.ge = kl * 4 + p->edge_numL[e];       // global key of the edge (owner element, owner's edge number): both
                                                   // sides of a cut order their messages by it, whatever window they see
            c.myCol = lMine ? led[s].l : led[s].r;
            c.myNum = lMine ? p->edge_numL[e] : p->edge_numR[e];
            c.ghCol = lMine ? led[s].r : led[s].l;
            c.ghNum = lMine ? p->edge_numR[e] : p->edge_numL[e];
            c.slot = s;
            c.side = lMine ? 0 : 1;
            cuts.push_back(c);
            pl.bndList.push_back(s);      // evaluated with the boundary edges after the halo exchange, so that the
                                          // interior-edge kernel can run while the halo is in flight
        }
        std::stable_sort(cuts.begin(), cuts.end(), [](const Cut &a, const Cut &b) {
            return a.peer != b.peer ? a.peer < b.peer : a.ge < b.ge;
        });
        const int64_t per = 4 * NEd + (p->dissipation ? 3 : 0);
        pl.ePer = (int)per;
        for (auto &c : cuts) {
            pl.sendCounts[c.peer] += per;
            pl.recvCounts[c.peer] += per;
            sendElem.push_back(c.myCol);
            sendRow0.push_back(c.myNum * NEd);
            recvCol.push_back(c.ghCol);
            recvRow0.push_back(c.ghNum * NEd);
            pl.cutSlot.push_back(c.slot);
            pl.cutSide.push_back(c.side);
        }
        pl.nSendEdges = pl.nRecvEdges = (int)cuts.size();
        pl.sendTotal = pl.recvTotal = per * (int64_t)cuts.size();
    }

    // ---- shared vertices (dissipation): the element -> vertex max merge (euler.go:1048-1065) must see the
    // incident elements of every partition.  Each pair of partitions that both touch a vertex exchanges its
    // local maxima for it; both sides list the vertices in ascending global id, so the messages line up.
    pl.vtxCounts.assign(pl.nParts, 0);
    if (pl.nParts > 1 && p->dissipation) {
        std::vector<int> firstPart((size_t)p->NV, -1);
        std::vector<std::pair<int64_t, int>> multi;      // (vertex, partition) for vertices touched by >= 2 partitions
        // (ascending global element id = partition by partition: the same visiting order as the reference's per-partition
        // loops; in window mode only the window's elements are seen, which by contract include every element touching a
        // vertex of this partition)
        for (int64_t kw = 0; kw < Kw; kw++) {
            const int pt = bucket_of(kw + kOff, Kg, pl.nParts);
            for (int v = 0; v < 3; v++) {
                const int64_t vid = p->EToV[kw * 3 + v];
                if (firstPart[vid] < 0) firstPart[vid] = pt;
                else if (firstPart[vid] != pt) multi.emplace_back(vid, pt);
            }
        }
        std::sort(multi.begin(), multi.end());
        multi.erase(std::unique(multi.begin(), multi.end()), multi.end());
        std::vector<std::vector<int>> byPeer(pl.nParts);
        for (size_t i = 0; i < multi.size();) {
            size_t j = i;
            const int64_t vid = multi[i].first;
            bool mineTouches = firstPart[vid] == pl.part;
            while (j < multi.size() && multi[j].first == vid) { mineTouches |= multi[j].second == pl.part; j++; }
            if (mineTouches) {
                if (firstPart[vid] != pl.part) byPeer[firstPart[vid]].push_back((int)vid);
                for (size_t t = i; t < j; t++)
                    if (multi[t].second != pl.part) byPeer[multi[t].second].push_back((int)vid);
            }
            i = j;
        }
        for (int pt = 0; pt < pl.nParts; pt++) {
            pl.vtxCounts[pt] = 2 * (int64_t)byPeer[pt].size();
            pl.vtxList.insert(pl.vtxList.end(), byPeer[pt].begin(), byPeer[pt].end());
        }
    }

    return 0;
}

// Tensor maps of the RK registers for the TMA variant of kernel 5 (ElemWsArgs::tmaps): rank 2, inner dimension = element
// columns, outer = the 4 NpInt rows of a register, box = one [4 NpInt x 32] slab, no swizzle (dense 256-byte rows).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tmaps(dfr2d_handle *h) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { h->err = "cuTensorMapEncodeTiled is not available in this driver"; return 2; }
    alignas(64) CUtensorMap maps[4];
    double *base[4] = {h->q[0], h->q[2], h->q[3], h->R};
    const cuuint64_t dims[2] = {(cuuint64_t)h->Kp, (cuuint64_t)(4 * h->NpInt)};
    const cuuint64_t strides[1] = {(cuuint64_t)h->Kp * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)kElemsPerBlock, (cuuint32_t)(4 * h->NpInt)};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 4; i++) {
        CUresult r = ((EncodeTiledFn)fn)(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base[i], dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return 2; }
    }
    unsigned char *d = nullptr;
    if (int rc = dev_alloc(h, &d, sizeof(maps))) return rc;
    CK(cudaMemcpy(d, maps, sizeof(maps), cudaMemcpyHostToDevice));
    h->tmaps = d;
    return 0;
}

static int create_impl(dfr2d_handle *h, const dfr2d_problem *p) {
    const int N = p->N;
    h->NpInt = (N + 1) * (N + 2) / 2;
    h->NpEdge = N + 2;
    h->NpFlux = (N + 2) * (N + 4);
    const int NI = h->NpInt, NEd = h->NpEdge;
    CK(cudaSetDevice(h->device));
    switch (N) {
        case 0: pack_ops<0>(p, h->opsHost); break;
        case 1: pack_ops<1>(p, h->opsHost); break;
        case 2: pack_ops<2>(p, h->opsHost); break;
        case 3: pack_ops<3>(p, h->opsHost); break;
        default: pack_ops<4>(p, h->opsHost); break;
    }
    h->opsFp = ops_fingerprint(h->opsHost);
    h->haveDiv = p->Div != nullptr;
    // physics block
    Phys &ph = h->ph;
    ph.gamma = p->FSFar.Gamma;
    ph.CFL = p->CFL;
    ph.FinalTime = p->FinalTime;
    ph.fs[0] = p->FSFar; ph.fs[1] = p->FSIn; ph.fs[2] = p->FSOut;
    ph.vortex = p->vortex;
    ph.sdKappa = (p->Kappa != 0.0) ? p->Kappa : 5.0;         // dissipation.go:140-147
    h->kappaGiven = p->Kappa;
    ph.Eps0 = 5.0 / 1.5;                                     // fixed before the kappa override
    ph.S0 = 4.0 / pow((double)(N + 1), 4.0);                 // dissipation.go:500
    ph.Cdiff = 1.0 / (double)((N + 1) * (N + 1));            // euler.go:958-960
    ph.Omega = 1.0 * (double)(N * N);                        // edges.go:218-219
    ph.fluxType = p->flux_type;
    ph.localDT = p->local_time_stepping ? 1 : 0;
    ph.dissipation = p->dissipation ? 1 : 0;
    ph.N = N;
    ph.maxIter = p->max_iterations;

    // ---- partition plan (host only) -------------------------------------------------------------
    dfr2d_plan pl;
    pl.nParts = h->nParts; pl.part = h->part;
    pl.kOff = h->winOff; pl.Kglobal = h->Kglobal;           // (0, 0) unless dfr2d_create_window
    if (int rc = build_plan(p, pl)) { h->err = pl.err; return rc; }
    h->hostPitch = p->K;                                    // host arrays of set/get_state etc. have the problem's columns
    h->hostOff = pl.k0 - pl.kOff;
    h->Kglobal = pl.Kglobal; h->k0 = pl.k0; h->k1 = pl.k1;
    h->K = pl.K; h->G = pl.G; h->Kp = pl.Kp; h->NE = pl.NE; h->NEp = pl.NEp; h->NV = pl.NV; h->NBP = pl.NBP;
    h->sendCounts = pl.sendCounts; h->recvCounts = pl.recvCounts;
    h->nSendEdges = pl.nSendEdges; h->nRecvEdges = pl.nRecvEdges; h->sendTotal = pl.sendTotal; h->recvTotal = pl.recvTotal;
    h->ePer = pl.ePer;
    h->vtxCounts = pl.vtxCounts;
    h->nVtx = (int)pl.vtxList.size();
    h->dissCounts.assign(h->nParts, 0);
    const int Kp = h->Kp, K = h->K;
    const int64_t k0 = h->k0;
    const double np12 = (double)((N + 1) * (N + 1));
    auto &ekL = pl.ekL; auto &ekR = pl.ekR; auto &emeta = pl.emeta; auto &etoe = pl.etoe;
    auto &enx = pl.enx; auto &eny = pl.eny; auto &eoohk = pl.eoohk; auto &eooLen = pl.eooLen;
    auto &bpx = pl.bpx; auto &bpy = pl.bpy; auto &Jdet = pl.Jdet; auto &Jinv = pl.Jinv; auto &IInII = pl.IInII;
    auto &sendElem = pl.sendElem; auto &sendRow0 = pl.sendRow0; auto &recvCol = pl.recvCol; auto &recvRow0 = pl.recvRow0;

    // ---- device memory -------------------------------------------------------------------------
    const size_t reg = (size_t)4 * NI * Kp;
    for (int r = 0; r < 4; r++) {
        if (int rc = dev_alloc(h, &h->q[r], reg)) return rc;
        CK(cudaMemset(h->q[r], 0, reg * sizeof(double)));
    }
    h->q[4] = h->q[1];
    if (h->Kp > h->K)
        for (int r = 0; r < 4; r++) k_fill_pad<<<64, 256>>>(h->q[r], NI, h->K, h->Kp);
    if (int rc = dev_alloc(h, &h->R, reg)) return rc;
    CK(cudaMemset(h->R, 0, reg * sizeof(double)));
    if (int rc = dev_alloc(h, &h->qface, (size_t)4 * 3 * NEd * Kp)) return rc;
    CK(cudaMemset(h->qface, 0, (size_t)4 * 3 * NEd * Kp * sizeof(double)));
    if (int rc = dev_alloc(h, &h->eflux, (size_t)4 * NEd * h->NEp)) return rc;
    CK(cudaMemset(h->eflux, 0, (size_t)4 * NEd * h->NEp * sizeof(double)));
    if (int rc = dev_alloc(h, &h->agg, (size_t)h->NEp)) return rc;
    CK(cudaMemset(h->agg, 0, (size_t)h->NEp * sizeof(double)));
    if (int rc = dev_alloc(h, &h->DT, (size_t)Kp)) return rc;
    CK(cudaMemset(h->DT, 0, (size_t)Kp * sizeof(double)));
    if (int rc = dev_upload(h, &h->Jdet, Jdet)) return rc;
    if (int rc = dev_upload(h, &h->Jinv, Jinv)) return rc;
    if (int rc = dev_upload(h, &h->IInII, IInII)) return rc;
    if (int rc = dev_upload(h, &h->etoe, etoe)) return rc;
    if (int rc = dev_upload(h, &h->ekL, ekL, h->NEp)) return rc;
    if (int rc = dev_upload(h, &h->ekR, ekR, h->NEp)) return rc;
    if (int rc = dev_upload(h, &h->emeta, emeta, h->NEp)) return rc;
    if (int rc = dev_upload(h, &h->enx, enx, h->NEp)) return rc;
    if (int rc = dev_upload(h, &h->eny, eny, h->NEp)) return rc;
    if (int rc = dev_upload(h, &h->eoohk, eoohk, h->NEp)) return rc;
    if (int rc = dev_upload(h, &h->bndList, pl.bndList)) return rc;
    h->nBnd = (int)pl.bndList.size();
    if (int rc = dev_upload(h, &h->bpx, bpx)) return rc;
    if (int rc = dev_upload(h, &h->bpy, bpy)) return rc;
    if (h->nParts > 1) {
        // mailbox layout (8-byte words): [recv EDGE | recv VERTEX | recv DISS | flags[3][32] | wave flags[32] |
        // wave inbox[32][2 slots][2]]; every region starts on a 16-word boundary
        auto up16 = [](int64_t w) { return (w + 15) & ~(int64_t)15; };
        const int64_t nE = h->recvTotal, nV = ph.dissipation ? 2 * (int64_t)h->nVtx : 0,
                      nD = ph.dissipation ? (int64_t)4 * NEd * h->nRecvEdges : 0;
        h->mbOff[DFR2D_XCHG_EDGE] = 0;
        h->mbOff[DFR2D_XCHG_VERTEX] = up16(nE);
        h->mbOff[DFR2D_XCHG_DISS] = h->mbOff[DFR2D_XCHG_VERTEX] + up16(nV);
        h->mbFlags = h->mbOff[DFR2D_XCHG_DISS] + up16(nD);
        h->mbWaveFlag = h->mbFlags + 3 * kMaxParts;
        h->mbWaveIn = h->mbWaveFlag + kMaxParts;
        h->mailboxWords = (size_t)(h->mbWaveIn + 4 * kMaxParts);
        if (int rc = dev_alloc(h, &h->mailbox, h->mailboxWords)) return rc;
        CK(cudaMemset(h->mailbox, 0, h->mailboxWords * sizeof(unsigned long long)));
        h->recvBuf = (double *)(h->mailbox + h->mbOff[DFR2D_XCHG_EDGE]);
        if (int rc = dev_alloc(h, &h->sendBuf, (size_t)h->sendTotal)) return rc;
        if (int rc = dev_upload(h, &h->sendElem, sendElem)) return rc;
        if (int rc = dev_upload(h, &h->sendRow0, sendRow0)) return rc;
        if (int rc = dev_upload(h, &h->recvCol, recvCol)) return rc;
        if (int rc = dev_upload(h, &h->recvRow0, recvRow0)) return rc;
        if (int rc = dev_upload(h, &h->cutSlot, pl.cutSlot)) return rc;
        if (int rc = dev_upload(h, &h->cutSide, pl.cutSide)) return rc;
    }
    if (ph.dissipation) {
        // vertices: keep global numbering (vertex arrays are small next to the element arrays)
        std::vector<int> etov((size_t)3 * Kp, 0);
        std::vector<double> hk((size_t)Kp, 1.0);
        for (int k = 0; k < K; k++) {
            const int64_t kg = k0 + k - pl.kOff;            // row of the problem arrays
            for (int v = 0; v < 3; v++) etov[(size_t)v * Kp + k] = p->EToV[kg * 3 + v];
            hk[k] = p->EdgeLenMax[kg] / np12;
        }
        // a ghost column's three vertex values arrive with the Q_Face message into private slots behind the real vertices
        for (int g = 0; g < h->G; g++)
            for (int v = 0; v < 3; v++) etov[(size_t)v * Kp + K + g] = h->NV + 3 * g + v;
        std::vector<double> nxk((size_t)3 * Kp, 0.0), nyk((size_t)3 * Kp, 0.0);
        for (int k = 0; k < K; k++)
            for (int le = 0; le < 3; le++) {
                nxk[(size_t)le * Kp + k] = p->FaceNormX[(k0 + k - pl.kOff) + p->K * le];
                nyk[(size_t)le * Kp + k] = p->FaceNormY[(k0 + k - pl.kOff) + p->K * le];
            }
        DissBuffers &d = h->ds;
        if (int rc = dev_upload(h, &d.etov, etov)) return rc;
        if (int rc = dev_upload(h, &d.hk, hk)) return rc;
        if (int rc = dev_upload(h, &d.nxk, nxk)) return rc;
        if (int rc = dev_upload(h, &d.nyk, nyk)) return rc;
        if (int rc = dev_upload(h, &d.eooLen, eooLen, h->NEp)) return rc;
        if (int rc = dev_alloc(h, &d.sigma, (size_t)Kp)) return rc;
        if (int rc = dev_alloc(h, &d.epsk, (size_t)Kp)) return rc;
        if (int rc = dev_alloc(h, &d.se, (size_t)Kp)) return rc;
        if (int rc = dev_alloc(h, &d.sigmaV, (size_t)h->NV)) return rc;
        if (int rc = dev_alloc(h, &d.epsV, (size_t)h->NV + 3 * (size_t)h->G)) return rc;
        CK(cudaMemset(d.epsV, 0, ((size_t)h->NV + 3 * (size_t)h->G) * sizeof(double)));
        if (h->nParts > 1) {
            for (int pt = 0; pt < h->nParts; pt++) h->dissCounts[pt] = (h->sendCounts[pt] / h->ePer) * 4 * NEd;
            if (int rc = dev_upload(h, &h->vtxList, pl.vtxList)) return rc;
            if (int rc = dev_alloc(h, &h->vSendBuf, (size_t)2 * h->nVtx)) return rc;
            h->vRecvBuf = (double *)(h->mailbox + h->mbOff[DFR2D_XCHG_VERTEX]);
            if (int rc = dev_alloc(h, &h->dSendBuf, (size_t)4 * NEd * h->nSendEdges)) return rc;
            h->dRecvBuf = (double *)(h->mailbox + h->mbOff[DFR2D_XCHG_DISS]);
        }
        if (int rc = dev_alloc(h, &d.dissX, (size_t)4 * NI * Kp)) return rc;
        if (int rc = dev_alloc(h, &d.dissY, (size_t)4 * NI * Kp)) return rc;
        if (int rc = dev_alloc(h, &d.vn, (size_t)8 * NEd * h->NEp)) return rc;
        CK(cudaMemset(d.vn, 0, (size_t)8 * NEd * h->NEp * sizeof(double)));
        if (int rc = dev_alloc(h, &d.aggv, (size_t)h->NEp)) return rc;
        if (int rc = dev_alloc(h, &d.DTVisc, (size_t)Kp)) return rc;
        CK(cudaMemset(d.sigma, 0, (size_t)Kp * sizeof(double)));
        CK(cudaMemset(d.epsk, 0, (size_t)Kp * sizeof(double)));
        CK(cudaMemset(d.se, 0, (size_t)Kp * sizeof(double)));
        CK(cudaMemset(d.sigmaV, 0, (size_t)std::max(h->NV, 1) * sizeof(double)));
        CK(cudaMemset(d.epsV, 0, (size_t)std::max(h->NV, 1) * sizeof(double)));
        CK(cudaMemset(d.dissX, 0, (size_t)4 * NI * Kp * sizeof(double)));
        CK(cudaMemset(d.dissY, 0, (size_t)4 * NI * Kp * sizeof(double)));
        CK(cudaMemset(d.aggv, 0, (size_t)h->NEp * sizeof(double)));
        CK(cudaMemset(d.DTVisc, 0, (size_t)Kp * sizeof(double)));
    }
    if (int rc = dev_alloc(h, &h->sc, 1)) return rc;
    CK(cudaMemset(h->sc, 0, sizeof(DevScalars)));
    CK(cudaMallocHost(&h->scHost, sizeof(DevScalars)));
    memset(h->scHost, 0, sizeof(DevScalars));

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    h->edgeBlocks = sms * 8;
    if (const char *ev = getenv("DFR2D_EDGE_PPT")) h->edgePPT = atoi(ev);
    if (const char *ev = getenv("DFR2D_EDGE_SPLIT")) h->edgeSplit = atoi(ev);
    if (const char *ev = getenv("DFR2D_EDGE_BLOCKS_PER_SM")) h->edgeBlocks = sms * std::max(1, atoi(ev));
    h->pfTiles = 1;
    if (const char *ev = getenv("DFR2D_PREFETCH_TILES")) h->pfTiles = atoi(ev);
    h->sms = sms;
    // measured on B200 (profiles/r01h_*, r01i_*): the warp-specialised DMMA kernel wins at N >= 2 (N=4 8M: 5.2 vs 8.1 ms;
    // N=3 2M: 0.92 vs 1.04 ms; N=2 2M: 0.63 vs 0.67 ms; N=2 200K: 71 vs 73 us); N <= 1 keeps the row-per-thread DFMA kernel
    // (operators of 3 x 15 / 1 x 8 entries: the 8 x 8 x 4 tiles would be mostly padding)
    h->elemKernel = (N >= 2) ? 5 : 1;
    if (const char *ev = getenv("DFR2D_ELEM_KERNEL")) h->elemKernel = atoi(ev);
    if (const char *ev = getenv("DFR2D_WS_STAGES")) h->wsStages = atoi(ev);
    if (const char *ev = getenv("DFR2D_WS_CW")) h->wsCW = atoi(ev) == 12 ? 12 : 8;
    // measured (profiles/r02x_bench_c5_split{0,1}.json, two alternating runs each at C5): 32.94 / 33.15 ms per step without,
    // 32.01 / 31.74 ms with the interpolation warps; kernel 5 itself 4.99 -> 4.65-4.70 ms (74.5 % -> 79-80 % of the HBM roofline);
    // N=2 (2M triangles) 0.629 -> 0.626 ms: neutral.  Default on.
    h->wsSplit = 1;
    if (const char *ev = getenv("DFR2D_WS_SPLIT")) h->wsSplit = atoi(ev) != 0;
    if (const char *ev = getenv("DFR2D_EDGE_OVERLAP")) h->edgeOverlap = atoi(ev) != 0;
    if (const char *ev = getenv("DFR2D_EDGE_WS")) h->edgeWs = atoi(ev) != 0;
    if (const char *ev = getenv("DFR2D_EDGE_WS_STAGES")) h->edgeWsStages = atoi(ev);
    if (h->edgeOverlap && h->nParts == 1) {
        CK(cudaStreamCreateWithFlags(&h->auxStream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
    }
    if (const char *ev = getenv("DFR2D_DISS_PREFETCH")) h->dissPrefetch = atoi(ev) > 0 ? 1 : 0;
    if (const char *ev = getenv("DFR2D_EDGE_VISC_FUSED")) h->edgeViscFused = atoi(ev) != 0 ? 1 : 0;
    {
        std::vector<double> fr;
        switch (N) {
            case 0: build_mma_frags<0>(p->DivInt, p->FluxEdgeInterp, fr); break;
            case 1: build_mma_frags<1>(p->DivInt, p->FluxEdgeInterp, fr); break;
            case 2: build_mma_frags<2>(p->DivInt, p->FluxEdgeInterp, fr); break;
            case 3: build_mma_frags<3>(p->DivInt, p->FluxEdgeInterp, fr); break;
            default: build_mma_frags<4>(p->DivInt, p->FluxEdgeInterp, fr); break;
        }
        if (int rc = dev_upload(h, &h->mmaFrags, fr)) return rc;
    }
    if (const char *ev = getenv("DFR2D_PREFETCH_TILES")) h->pfTiles = atoi(ev);
    // measured on B200, 2M triangles (profiles/r01p_grad_ab_N*.json): the pipelined tensor-core gradient kernel wins at
    // N = 4 (k_edge + gradient 3.86 -> 2.74 ms), N = 3 (2.32 -> 2.15) and N = 2 (1.56 -> 1.43).  N = 1 keeps the DFMA
    // kernel: its operators are 15 x 15, and its summation order is the one the 1e-11 parity bar at N = 1 relies on
    // (the RT2 divergence operator is ill conditioned, tests/test_noise_floor.py)
    // (r2, profiles/r02t_ab_N*.json) the warp-specialised form of the same kernel (k_grad_ws: producer warps own the copies
    // and the index chain) is bitwise identical and faster at every N >= 2: stage 5.15 -> 4.88 ms at N=4, 3.86 -> 3.75 at
    // N=3, 2.72 -> 2.53 at N=2 (2M triangles); k_grad_pipe stays selectable as 3 (2 maps to it as well)
    h->gradKernel = (N >= 2) ? 4 : 1;
    if (const char *ev = getenv("DFR2D_GRAD_KERNEL")) h->gradKernel = atoi(ev) >= 4 ? 4 : (atoi(ev) >= 2 ? 3 : 1);
    // accumulation groups of the m-tiles: k_grad_pipe is best with {3,1} (profiles/r02s), k_grad_ws with {2,2} -- N=4: stage
    // 4.87 -> 4.79 ms, N=3 ({2,1}): 3.73 -> 3.62 ms (profiles/r02u_ab_N*.json)
    h->gradMG = (h->gradKernel == 4 && N >= 3) ? 2 : 3;
    if (const char *ev = getenv("DFR2D_GRAD_MG")) h->gradMG = atoi(ev) == 2 ? 2 : 3;
    if (const char *ev = getenv("DFR2D_GRAD_SKEW_NS")) h->gradSkewNs = std::max(0, std::min(atoi(ev), 100000));
    // measured (2M triangles, profiles/r02l_ab_N*.json): the warp-specialised ring with the PerssonC0 terms
    // (k_elem_ws<N,8,true>) wins at every N >= 2 -- N=4: 1.54 ms against 2.37 (k_elem_mma_diss) / 2.60 (k_elem<4,true>);
    // N=3: 1.23 / 1.85 / 1.54; N=2: 0.77 / 1.19 / 0.82.  N = 1 keeps the DFMA kernel (3 x 15 operators).
    h->dissElemKernel = (N >= 2) ? 5 : 1;
    if (const char *ev = getenv("DFR2D_DISS_ELEM_KERNEL")) h->dissElemKernel = atoi(ev);
    if (ph.dissipation && h->dissElemKernel == 3) {
        std::vector<double> fr;
        switch (N) {
            case 1: build_mma_diss_frags<1>(p->DivInt, p->Vinv, p->V, fr); break;
            case 2: build_mma_diss_frags<2>(p->DivInt, p->Vinv, p->V, fr); break;
            case 3: build_mma_diss_frags<3>(p->DivInt, p->Vinv, p->V, fr); break;
            default: build_mma_diss_frags<4>(p->DivInt, p->Vinv, p->V, fr); break;
        }
        if (int rc = dev_upload(h, &h->mmaDissFrags, fr)) return rc;
    }
    if (ph.dissipation && h->gradKernel >= 2) {
        std::vector<double> tb;
        grad_table_for(N, p->Div, p->Bary, tb);
        if (int rc = dev_upload(h, &h->gradTable, tb)) return rc;
        // (x, y) metric of the RT points of edge 0, 1, 2 (DXMetric / DYMetric edge rows, DG2D/dfr_startup.go:213-254)
        // as plain rows [mx0, my0, mx1, my1, mx2, my2][Kp], same operation order as k_grad
        std::vector<double> mxy((size_t)6 * Kp, 0.0);
        for (int k = 0; k < K; k++) {
            const double oojd = 1.0 / pl.Jdet[k];
            for (int le = 0; le < 3; le++) {
                const double iin = pl.IInII[(size_t)le * Kp + k];
                mxy[(size_t)(2 * le) * Kp + k] = oojd * p->FaceNormX[(k0 + k - pl.kOff) + p->K * le] * iin;
                mxy[(size_t)(2 * le + 1) * Kp + k] = oojd * p->FaceNormY[(k0 + k - pl.kOff) + p->K * le] * iin;
            }
        }
        if (int rc = dev_upload(h, &h->gradMxy, mxy)) return rc;
    }
    if (const char *ev = getenv("DFR2D_WS_TMA"))
        if (atoi(ev) > 0 && !ph.dissipation && h->elemKernel == 5)
            if (int rc = make_tmaps(h)) return rc;
    CK(cudaDeviceSynchronize());
    return 0;
}

static int create_common(const dfr2d_problem *p, int64_t K_global, int64_t k_offset, int n_parts, int part, int device,
                         dfr2d_handle **out) {
    if (!p || !out) { g_create_error = "null argument"; return 1; }
    if (p->N < 0 || p->N > DFR2D_MAX_ORDER) { g_create_error = "polynomial order out of range 0..4"; return 1; }
    if (n_parts < 1 || part < 0 || part >= n_parts || K_global < n_parts || k_offset < 0 || k_offset + p->K > K_global) {
        g_create_error = "bad partition request";
        return 1;
    }
    dfr2d_handle *h = new dfr2d_handle();
    h->N = p->N; h->device = device; h->nParts = n_parts; h->part = part;
    h->Kglobal = K_global; h->winOff = k_offset;
    int rc = create_impl(h, p);
    if (rc) {
        g_create_error = h->err;
        dfr2d_destroy(h);
        return rc;
    }
    *out = h;
    return 0;
}

extern "C" int dfr2d_create(const dfr2d_problem *p, int n_parts, int part, int device, dfr2d_handle **out) {
    return create_common(p, p ? p->K : 0, 0, n_parts, part, device, out);
}

// Partition-local set-up: `p` describes only the window [k_offset, k_offset + p->K) of a mesh of K_global elements.
extern "C" int dfr2d_create_window(const dfr2d_problem *p, int64_t K_global, int64_t k_offset, int n_parts, int part,
                                   int device, dfr2d_handle **out) {
    return create_common(p, K_global, k_offset, n_parts, part, device, out);
}

// ---- state I/O -------------------------------------------------------------------------------------
static int copy_in(dfr2d_handle *h, double *dst, const double *Q) {
    CK(cudaSetDevice(h->device));
    const size_t rows = (size_t)4 * h->NpInt;
    CK(cudaMemcpy2DAsync(dst, (size_t)h->Kp * sizeof(double), Q + h->hostOff, (size_t)h->hostPitch * sizeof(double),
                         (size_t)h->K * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
static int copy_out(dfr2d_handle *h, const double *src, double *Q) {
    CK(cudaSetDevice(h->device));
    const size_t rows = (size_t)4 * h->NpInt;
    CK(cudaMemcpy2DAsync(Q + h->hostOff, (size_t)h->hostPitch * sizeof(double), src, (size_t)h->Kp * sizeof(double),
                         (size_t)h->K * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dfr2d_set_state(dfr2d_handle *h, const double *Q) {
    if (!h || !Q) return 1;
    h->qfaceValid = false;
    return copy_in(h, h->q[0], Q);
}
extern "C" int dfr2d_get_state(dfr2d_handle *h, double *Q) {
    if (!h || !Q) return 1;
    return copy_out(h, h->q[0], Q);
}
// All partitions of one process at once: the copies of different devices are enqueued before anything is waited for, so
// they cross PCIe concurrently (one controller thread, n GPUs).
static int multi_copy(dfr2d_handle **hs, int n, double *Q, bool in) {
    if (!hs || !Q || n < 1) return 1;
    for (int g = 0; g < n; g++) {
        dfr2d_handle *h = hs[g];
        if (!h) return 1;
        CK(cudaSetDevice(h->device));
        const size_t rows = (size_t)4 * h->NpInt;
        if (in) {
            h->qfaceValid = false;
            CK(cudaMemcpy2DAsync(h->q[0], (size_t)h->Kp * sizeof(double), Q + h->hostOff, (size_t)h->hostPitch * sizeof(double),
                                 (size_t)h->K * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream));
        } else {
            CK(cudaMemcpy2DAsync(Q + h->hostOff, (size_t)h->hostPitch * sizeof(double), h->q[0], (size_t)h->Kp * sizeof(double),
                                 (size_t)h->K * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    for (int g = 0; g < n; g++) {
        dfr2d_handle *h = hs[g];
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}
extern "C" int dfr2d_multi_set_state(dfr2d_handle **hs, int n, const double *Q) { return multi_copy(hs, n, (double *)Q, true); }
extern "C" int dfr2d_multi_get_state(dfr2d_handle **hs, int n, double *Q) { return multi_copy(hs, n, Q, false); }

// rk.Time / rk.StepCount as the host sees them (restart from a saved state; euler.go:175-182 keeps them in the Euler struct)
extern "C" int dfr2d_set_clock(dfr2d_handle *h, double time, int64_t steps) {
    if (!h || steps < 0) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(h->scHost, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost));
    h->scHost->time[0] = h->scHost->time[1] = time;
    h->scHost->timeOut = time;
    h->scHost->steps = steps;
    h->scHost->finished = 0;
    CK(cudaMemcpy(h->sc, h->scHost, sizeof(DevScalars), cudaMemcpyHostToDevice));
    h->stepIndex = steps;
    return 0;
}

extern "C" int dfr2d_set_register(dfr2d_handle *h, int reg, const double *Q) {
    if (!h || !Q || reg < 0 || reg > 4) return 1;
    h->qfaceValid = false;
    return copy_in(h, h->q[reg], Q);
}
extern "C" int dfr2d_get_register(dfr2d_handle *h, int reg, double *Q) {
    if (!h || !Q || reg < 0 || reg > 5) return 1;
    return copy_out(h, reg == 5 ? h->R : h->q[reg], Q);
}

// ---- launches ---------------------------------------------------------------------------------------
#define DISPATCH_N(N_, ...)                              \
    switch (N_) {                                            \
        case 0: { constexpr int NN = 0; __VA_ARGS__; } break; \
        case 1: { constexpr int NN = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int NN = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int NN = 3; __VA_ARGS__; } break; \
        default: { constexpr int NN = 4; __VA_ARGS__; } break; \
    }

static int launch_check(dfr2d_handle *h, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        h->err = std::string(what) + ": " + cudaGetErrorString(e);
        return 2;
    }
    h->launches++;
    return 0;
}

static int run_interp(dfr2d_handle *h, const double *reg) {
    const int blocks = (h->K + kElemsPerBlock - 1) / kElemsPerBlock;
    DISPATCH_N(h->N, (k_interp<NN><<<blocks, kElemThreads, 0, h->stream>>>(h->K, h->Kp, reg, h->qface)));
    return launch_check(h, "k_interp");
}

static unsigned long long xchg_seq(const dfr2d_handle *h) { return (unsigned long long)h->stageCounter + 1ull; }

static int run_pack(dfr2d_handle *h) {
    if (h->nSendEdges == 0) return 0;
    HaloPackArgs a{};
    a.total = h->nSendEdges * h->ePer; a.per = h->ePer; a.npEdge = h->NpEdge; a.planeRows = 3 * h->NpEdge; a.rowBase = 0;
    a.Kp = h->Kp; a.nTail = h->ePer - 4 * h->NpEdge;
    a.src = h->qface; a.src2 = nullptr; a.elem = h->sendElem; a.row0 = h->sendRow0; a.etov = h->ds.etov; a.epsV = h->ds.epsV;
    a.buf = h->sendBuf; a.put = h->connected ? h->putTab[DFR2D_XCHG_EDGE] : nullptr; a.seq = xchg_seq(h);
    k_halo_pack<<<(a.total + 255) / 256, 256, 0, h->stream>>>(a);
    return launch_check(h, "k_halo_pack");
}

static int run_unpack(dfr2d_handle *h) {
    if (h->nRecvEdges == 0) return 0;
    HaloUnpackArgs a{};
    a.total = h->nRecvEdges * h->ePer; a.per = h->ePer; a.npEdge = h->NpEdge; a.planeRows = 3 * h->NpEdge; a.rowBase = 0;
    a.Kp = h->Kp; a.nTail = h->ePer - 4 * h->NpEdge; a.K = h->K; a.NV = h->NV;
    a.dst = h->qface; a.dst2 = nullptr; a.col = h->recvCol; a.row0 = h->recvRow0; a.epsV = h->ds.epsV;
    a.buf = h->recvBuf; a.wait = h->connected ? h->waitTab[DFR2D_XCHG_EDGE] : nullptr; a.seq = xchg_seq(h);
    k_halo_unpack<<<(a.total + 255) / 256, 256, 0, h->stream>>>(a);
    return launch_check(h, "k_halo_unpack");
}

// viscous exchange: the owner-normal components of my side of every cut edge (k_vn_pack / k_vn_unpack, dfr2d_peer.cuh)
static int run_pack_diss(dfr2d_handle *h) {
    if (h->nSendEdges == 0) return 0;
    const int total = h->nSendEdges * 4 * h->NpEdge;
    k_vn_pack<<<(total + 255) / 256, 256, 0, h->stream>>>(h->nSendEdges, h->NpEdge, h->NEp, h->ds.vn, h->cutSlot, h->cutSide, h->dSendBuf,
                                                          h->connected ? h->putTab[DFR2D_XCHG_DISS] : nullptr, xchg_seq(h));
    return launch_check(h, "k_vn_pack");
}

static int run_unpack_diss(dfr2d_handle *h) {
    if (h->nRecvEdges == 0) return 0;
    const int total = h->nRecvEdges * 4 * h->NpEdge;
    k_vn_unpack<<<(total + 255) / 256, 256, 0, h->stream>>>(h->nRecvEdges, h->NpEdge, h->NEp, h->ds.vn, h->cutSlot, h->cutSide, h->dRecvBuf,
                                                            h->connected ? h->waitTab[DFR2D_XCHG_DISS] : nullptr, xchg_seq(h));
    return launch_check(h, "k_vn_unpack");
}

static int run_pack_vertex(dfr2d_handle *h) {
    if (h->nVtx == 0) return 0;
    k_vertex_pack<<<(h->nVtx + 255) / 256, 256, 0, h->stream>>>(h->nVtx, h->vtxList, h->ds.sigmaV, h->ds.epsV, h->vSendBuf,
                                                                h->connected ? h->putTab[DFR2D_XCHG_VERTEX] : nullptr, xchg_seq(h));
    return launch_check(h, "k_vertex_pack");
}

static int run_unpack_vertex(dfr2d_handle *h) {
    if (h->nVtx == 0) return 0;
    k_vertex_unpack_max<<<(h->nVtx + 255) / 256, 256, 0, h->stream>>>(h->nVtx, h->vtxList, (unsigned long long *)h->ds.sigmaV,
                                                                      (unsigned long long *)h->ds.epsV, h->vRecvBuf,
                                                                      h->connected ? h->waitTab[DFR2D_XCHG_VERTEX] : nullptr, xchg_seq(h));
    return launch_check(h, "k_vertex_unpack_max");
}

// part: 1 = interior edges (no ghost column involved: may run while the halo is in flight), 2 = boundary + cut edges
// (after the halo has been unpacked), 3 = both
static int run_edges(dfr2d_handle *h, int rk, int part) {
    // PerssonC0 path with the fused viscous edge flux: the VISC instantiations, run AFTER the gradient kernel
    const bool visc = h->ph.dissipation && h->edgeViscFused;
    EdgeArgs a{};
    a.ne = h->NE; a.NEp = h->NEp; a.Kp = h->Kp; a.Kown = h->K;
    a.kL = h->ekL; a.kR = h->ekR; a.meta = h->emeta;
    a.nx = h->enx; a.ny = h->eny; a.oohk = h->eoohk;
    a.bpx = h->bpx; a.bpy = h->bpy;
    a.qface = h->qface; a.eflux = h->eflux; a.agg = h->agg;
    a.vn = h->ds.vn; a.etov = h->ds.etov; a.epsV = h->ds.epsV; a.ooLen = h->ds.eooLen; a.aggv = h->ds.aggv;
    a.sc = h->sc;
    a.slot = (int)(h->stageCounter & 1);
    a.par = (int)(h->stepIndex & 1);
    a.stepIndex = h->stepIndex;
    a.ph = h->ph;
    const int blocks = std::max(1, std::min(h->edgeBlocks, (h->NE + 255) / 256));
    // points per thread: DFR2D_EDGE_PPT overrides (must divide N+2); agg is combined by atomicMax when an edge is split
    // measured (tools/sweep_edge_ppt.sh): whole edges per thread on big meshes, one point per thread when there are
    // too few edges to fill the machine (C2: 300K edges)
    // (r2) the switch-over was 64 sms 256 = 2.4M edges in round 1, bracketed only by C2 (300K edges: one point per
    // thread wins) and C5 (12M: whole edges win).  The 8-GPU partitions of C5 (1.5M edges) fell on the wrong side:
    // 0.285 ms where whole edges take 0.19 ms (profiles/r02h_*).  16 sms 256 = 606K edges.
    int ppt = h->edgePPT;
    if (ppt <= 0) ppt = (h->NE < 16 * h->sms * 256) ? 1 : h->N + 2;
    if ((h->N + 2) % ppt != 0) ppt = h->N + 2;
    // the boundary / cut-edge list is short (O(sqrt K) edges): one point per thread there, or a few dozen CTAs walk whole
    // edges through the BC transcendental functions one point after the other (measured at C5: 57 us for 8,000 edges)
    int pptList = (h->nBnd < 64 * h->sms * 256 && h->edgePPT <= 0) ? 1 : ppt;
    if ((h->N + 2) % pptList != 0) pptList = h->N + 2;
    // interior edges by the producer / consumer pipeline (two points per thread): even NpEdge, inviscid form only
    const bool useWs = h->edgeWs && h->edgeSplit && !visc && (h->N % 2 == 0);
    if (useWs) ppt = 2;
    if ((ppt != h->N + 2 || (h->edgeSplit && pptList != h->N + 2)) && h->ph.localDT && (part & 1))
    {
        CK(cudaMemsetAsync(h->agg, 0, (size_t)h->NEp * sizeof(double), h->stream));
        if (visc) CK(cudaMemsetAsync(h->ds.aggv, 0, (size_t)h->NEp * sizeof(double), h->stream));
    }
    a.list = nullptr; a.nlist = 0;
#define EDGE_LAUNCH(KERN)                                                                      \
    DISPATCH_N(h->N, {                                                                          \
        constexpr int NEd_ = NN + 2;                                                            \
        if (ppt == 1) KERN(NN, 1);                                                              \
        else if (NEd_ % 2 == 0 && ppt == 2) KERN(NN, (NEd_ % 2 == 0 ? 2 : 1));                  \
        else if (NEd_ % 3 == 0 && ppt == 3) KERN(NN, (NEd_ % 3 == 0 ? 3 : 1));                  \
        else KERN(NN, NEd_);                                                                    \
    })
    if (h->edgeSplit) {
        const int ib = std::max(1, std::min(h->edgeBlocks * 2, (int)(((long long)h->NEp * ((h->N + 2) / ppt) + 255) / 256)));
        // both kernels in one call on a single partition: the list kernel goes to the second stream, forked here (after
        // everything both depend on) and joined below; they write disjoint edge slots and share only the atomicMax words
        const bool overlap = part == 3 && h->auxStream != nullptr && h->nBnd > 0;
        if (overlap) {
            CK(cudaEventRecord(h->evFork, h->stream));
            CK(cudaStreamWaitEvent(h->auxStream, h->evFork, 0));
        }
        cudaStream_t listStream = overlap ? h->auxStream : h->stream;
        if ((part & 1) && useWs) {
            const int nTiles = (h->NEp + 127) / 128;
#define KWS(NN_, FL_) do {                                                                                              \
                using ED = EdgeWsDim<NN_>;                                                                              \
                int st = h->edgeWsStages >= 2 ? std::min(h->edgeWsStages, (int)ED::kMaxStages) : ED::stages();          \
                while (st > 2 && (size_t)st * ED::kStageDoubles * sizeof(double) > 232448 - 512) st--;                  \
                cudaFuncSetAttribute(k_edge_ws<NN_, FL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 512);   \
                k_edge_ws<NN_, FL_><<<std::min(nTiles, h->sms), ED::kThreads, (size_t)st * ED::kStageDoubles * sizeof(double), h->stream>>>(a, nTiles, st); \
            } while (0)
#define KWS_N(FL_) do { if (h->N == 4) KWS(4, FL_); else if (h->N == 2) KWS(2, FL_); else KWS(0, FL_); } while (0)
            switch (h->ph.fluxType) {
                case DFR2D_FLUX_Average: KWS_N(DFR2D_FLUX_Average); break;
                case DFR2D_FLUX_LaxFriedrichs: KWS_N(DFR2D_FLUX_LaxFriedrichs); break;
                case DFR2D_FLUX_Roe: KWS_N(DFR2D_FLUX_Roe); break;
                default: KWS_N(DFR2D_FLUX_RoeER); break;
            }
            if (int rc = launch_check(h, "k_edge_ws")) return rc;
        } else if (part & 1) {
        switch (h->ph.fluxType) {
#define KI_AVG(NN_, P_) do { if (visc) k_edge_int<NN_, DFR2D_FLUX_Average, P_, true><<<ib, 256, 0, h->stream>>>(a); else k_edge_int<NN_, DFR2D_FLUX_Average, P_, false><<<ib, 256, 0, h->stream>>>(a); } while (0)
#define KI_LAX(NN_, P_) do { if (visc) k_edge_int<NN_, DFR2D_FLUX_LaxFriedrichs, P_, true><<<ib, 256, 0, h->stream>>>(a); else k_edge_int<NN_, DFR2D_FLUX_LaxFriedrichs, P_, false><<<ib, 256, 0, h->stream>>>(a); } while (0)
#define KI_ROE(NN_, P_) do { if (visc) k_edge_int<NN_, DFR2D_FLUX_Roe, P_, true><<<ib, 256, 0, h->stream>>>(a); else k_edge_int<NN_, DFR2D_FLUX_Roe, P_, false><<<ib, 256, 0, h->stream>>>(a); } while (0)
#define KI_RER(NN_, P_) do { if (visc) k_edge_int<NN_, DFR2D_FLUX_RoeER, P_, true><<<ib, 256, 0, h->stream>>>(a); else k_edge_int<NN_, DFR2D_FLUX_RoeER, P_, false><<<ib, 256, 0, h->stream>>>(a); } while (0)
            case DFR2D_FLUX_Average: EDGE_LAUNCH(KI_AVG); break;
            case DFR2D_FLUX_LaxFriedrichs: EDGE_LAUNCH(KI_LAX); break;
            case DFR2D_FLUX_Roe: EDGE_LAUNCH(KI_ROE); break;
            default: EDGE_LAUNCH(KI_RER); break;
        }
        if (int rc = launch_check(h, "k_edge_int")) return rc;
        }
        if (!(part & 2) || h->nBnd == 0) return 0;
        a.list = h->bndList; a.nlist = h->nBnd;
        {
            const int ppt = pptList;       // (EDGE_LAUNCH dispatches on `ppt`)
            const int bb = std::max(1, std::min(h->edgeBlocks, (h->nBnd * ((h->N + 2) / ppt) + 255) / 256));
#define KB(NN_, P_) do { if (visc) k_edge<NN_, P_, true><<<bb, 256, 0, listStream>>>(a); else k_edge<NN_, P_, false><<<bb, 256, 0, listStream>>>(a); } while (0)
            EDGE_LAUNCH(KB);
        }
        if (overlap) {
            CK(cudaEventRecord(h->evJoin, h->auxStream));
            CK(cudaStreamWaitEvent(h->stream, h->evJoin, 0));
        }
        return launch_check(h, "k_edge(boundary)");
    }
    if (!(part & 2)) return 0;          // single generic kernel: everything happens in the second part
#define KG(NN_, P_) do { if (visc) k_edge<NN_, P_, true><<<blocks, 256, 0, h->stream>>>(a); else k_edge<NN_, P_, false><<<blocks, 256, 0, h->stream>>>(a); } while (0)
    a.Kown = 0x7fffffff;
    EDGE_LAUNCH(KG);
    (void)rk;
    return launch_check(h, "k_edge");
}

template <int N> static size_t elem_smem() { return (size_t)12 * Dim<N>::NpInt * kElemsPerBlock * sizeof(double); }

static int run_elem(dfr2d_handle *h, int rk, double *rhsOut, bool fuseInterp) {
    ElemArgs a{};
    a.K = h->K; a.Kp = h->Kp; a.NEp = h->NEp;
    a.pfTiles = 0;
    a.qs = h->q[rk];
    a.q0 = h->q[0]; a.q1 = h->q[1]; a.q2 = h->q[2]; a.q3 = h->q[3]; a.q4 = h->q[4]; a.R = h->R;
    a.qface = fuseInterp ? h->qface : nullptr;
    a.eflux = h->eflux;
    a.agg = h->agg; a.aggv = h->ds.aggv;
    a.DT = h->DT; a.DTVisc = h->ds.DTVisc;
    a.Jdet = h->Jdet; a.Jinv = h->Jinv; a.IInII = h->IInII;
    a.etoe = h->etoe;
    a.dissX = h->ds.dissX; a.dissY = h->ds.dissY; a.sigma = h->ds.sigma;
    a.rhsOut = rhsOut;
    a.sc = h->sc;
    a.rk = rk;
    a.slot = (int)(h->stageCounter & 1);
    a.par = (int)(h->stepIndex & 1);
    a.stepIndex = h->stepIndex;
    a.ph = h->ph;
    const int blocks = (h->K + kElemsPerBlock - 1) / kElemsPerBlock;
    if (rk == 4 && rhsOut == nullptr) {
        // kernel 5 reduces the residual register to its four maxima inside the launch (DevScalars::resMax) instead of
        // writing it out; every other element kernel stores it and dfr2d_residual reduces it afterwards
        h->resInScalars = h->ph.dissipation ? (h->dissElemKernel == 5) : (h->elemKernel == 5);
    }
    if (h->ph.dissipation && h->dissElemKernel == 5) {
        // the warp-specialised ring of kernel 5 with the PerssonC0 terms (k_elem_ws<N, 8, true>, dfr2d_elem_ws.cuh)
        ElemWsArgs ta{};
        ta.a = a;
        ta.nTiles = blocks;
        ta.nExtra = (rk == 0 || rhsOut != nullptr) ? 0 : (rk == 4 ? 4 : 1);
        DISPATCH_N(h->N, {
            using TD = WsDim<NN>;
            const size_t maxSmem = 232448 - 256;
            const int nExtraS = TD::extras_in_smem(ta.nExtra, true);
            int stages = (int)(maxSmem / TD::smem_bytes_diss(nExtraS, 1));
            stages = std::max(2, std::min(stages, h->wsStages > 1 ? h->wsStages : 2));
            if (TD::smem_bytes_diss(nExtraS, stages) > maxSmem) { h->err = "k_elem_ws<diss>: ring does not fit"; return 2; }
            ta.nStages = stages;
            if (!h->dissWsAttrSet) {
                cudaFuncSetAttribute(k_elem_ws<NN, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxSmem);
                h->dissWsAttrSet = true;
            }
            k_elem_ws<NN, 8, true><<<std::min(blocks, h->sms), (8 + kWsProdWarps) * 32, TD::smem_bytes_diss(nExtraS, stages), h->stream>>>(ta);
        });
        return launch_check(h, "k_elem_ws<diss>");
    }
    if (h->ph.dissipation && h->dissElemKernel == 3) {
        ElemMmaArgs ma{};
        ma.a = a;
        ma.a.pfTiles = h->dissPrefetch;       // bulk L2 prefetch of the next tile: measured SLOWER (3.00 vs 2.84 ms, 2M triangles,
                                              // profiles/r02i_*), so off unless DFR2D_DISS_PREFETCH=1
        ma.frags = h->mmaDissFrags;
        ma.nTiles = blocks;
        DISPATCH_N(h->N, {
            const size_t sm = MmaDissDim<NN>::kSmemBytes;
            if (h->mmaDissGrid == 0) {
                cudaFuncSetAttribute(k_elem_mma_diss<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                cudaFuncSetAttribute(k_elem_mma_diss<NN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                int occ = 1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elem_mma_diss<NN>, kElemThreads, sm);
                h->mmaDissGrid = h->sms * std::max(occ, 1);
            }
            k_elem_mma_diss<NN><<<std::min(blocks, h->mmaDissGrid), kElemThreads, sm, h->stream>>>(ma);
        });
        return launch_check(h, "k_elem_mma_diss");
    }
    if (h->ph.dissipation) {
        DISPATCH_N(h->N, {
            const size_t sm = elem_smem_diss<NN>();
            if (!h->smemAttrSet) {
                cudaFuncSetAttribute(k_elem<NN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                cudaFuncSetAttribute(k_elem<NN, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            }
            k_elem<NN, true><<<blocks, kElemThreads, sm, h->stream>>>(a);
        });
    } else {
        if (h->elemKernel == 5) {
            ElemWsArgs ta{};
            ta.a = a;
            ta.nTiles = blocks;
            ta.nExtra = (rk == 0 || rhsOut != nullptr) ? 0 : (rk == 4 ? 4 : 1);
            ta.tmaps = h->tmaps;
            DISPATCH_N(h->N, {
                using TD = WsDim<NN>;
                const size_t maxSmem = 232448 - 512;      // 227 KB per CTA minus the static mbarrier words
                int stages = (int)(maxSmem / TD::smem_bytes(ta.nExtra, 1));
                stages = std::max(2, std::min(stages, kWsMaxStages));     // two consumer groups: a waiter may be at most one phase ahead
                // measured (profiles/r01i_ring_depth.txt): a deeper ring is SLOWER -- the shared memory it takes comes out
                // of L1, and the 8-byte edge-flux gather lives on L1 hits (neighbouring elements share sectors and
                // edges).  3 stages for rk 0-2 (121-173 KB), 2 for rk 3 (115 KB) and rk 4 (219 KB, all that fits)
                stages = std::min(stages, (rk >= 3 && rhsOut == nullptr) ? 2 : 3);
                if (h->wsStages > 1) stages = std::min(std::max(2, (int)(maxSmem / TD::smem_bytes(ta.nExtra, 1))), h->wsStages);
                ta.nStages = stages;
                // three consumer groups (12 warps, 152 registers each) need a ring of >= 3 stages: every stage but the last
                // (rk 4: 109 KB per stage, two fit).  DFR2D_WS_CW=12 selects them; measured A/B in profiles/r02f_*
                const bool cw12 = h->wsCW == 12 && (int)(maxSmem / TD::smem_bytes(ta.nExtra, 1)) >= 3;
                if (cw12) stages = std::max(3, std::min(h->wsStages > 1 ? h->wsStages : 3, (int)(maxSmem / TD::smem_bytes(ta.nExtra, 1))));
                ta.nStages = stages;
                const size_t sm = TD::smem_bytes(ta.nExtra, stages);
                if (!h->smemAttrSet) {
                    cudaFuncSetAttribute(k_elem_ws<NN, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxSmem);
                    cudaFuncSetAttribute(k_elem_ws<NN, 12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxSmem);
                    cudaFuncSetAttribute(k_elem_ws<NN, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxSmem);
                }
                if (h->wsSplit && !cw12)
                    k_elem_ws<NN, 8, false, true><<<std::min(blocks, h->sms), (8 + 4 + kWsProdWarps) * 32, sm, h->stream>>>(ta);
                else if (cw12) k_elem_ws<NN, 12, false><<<std::min(blocks, h->sms), (12 + kWsProdWarps) * 32, sm, h->stream>>>(ta);
                else k_elem_ws<NN, 8, false><<<std::min(blocks, h->sms), (8 + kWsProdWarps) * 32, sm, h->stream>>>(ta);
            });
            h->smemAttrSet = true;
            return launch_check(h, "k_elem_ws");
        }
        if (h->elemKernel == 4) {
            ElemMmaArgs ma{};
            ma.a = a;
            ma.frags = h->mmaFrags;
            ma.nTiles = blocks;
            const int nExtra = (rk == 0) ? 0 : (rk == 4 ? 4 : 1);
            DISPATCH_N(h->N, {
                if (!h->smemAttrSet) {
                    cudaFuncSetAttribute(k_elem_pipe<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PipeDim<NN>::smem_bytes(4));
                    cudaFuncSetAttribute(k_elem_pipe<NN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                }
                const int oi = (nExtra == 0) ? 0 : (nExtra == 1 ? 1 : 2);
                if (h->pipeOcc[oi] == 0) {
                    int occ = 1;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elem_pipe<NN>, kPipeThreads, PipeDim<NN>::smem_bytes(nExtra));
                    h->pipeOcc[oi] = std::max(occ, 1);
                }
                k_elem_pipe<NN><<<std::min(blocks, h->sms * h->pipeOcc[oi]), kPipeThreads, PipeDim<NN>::smem_bytes(nExtra), h->stream>>>(ma);
            });
            h->smemAttrSet = true;
            return launch_check(h, "k_elem_pipe");
        }
        if (h->elemKernel == 3) {
            ElemMmaArgs ma{};
            ma.a = a;
            ma.frags = h->mmaFrags;
            ma.nTiles = blocks;
            DISPATCH_N(h->N, {
                const size_t sm = MmaDim<NN>::kSmemBytes;
                if (!h->smemAttrSet) {
                    cudaFuncSetAttribute(k_elem_mma<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                    cudaFuncSetAttribute(k_elem_mma<NN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    int occ = 1;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elem_mma<NN>, kElemThreads, sm);
                    h->mmaGrid = h->sms * std::max(occ, 1);
                }
                k_elem_mma<NN><<<std::min(blocks, h->mmaGrid), kElemThreads, sm, h->stream>>>(ma);
            });
            h->smemAttrSet = true;
            return launch_check(h, "k_elem_mma");
        }
        if (h->elemKernel == 2) {
            DISPATCH_N(h->N, {
                const size_t sm = (size_t)4 * (Dim<NN>::NpInt + Dim<NN>::NpFlux) * kElemsPerBlock * sizeof(double);
                if (!h->smemAttrSet) {
                    cudaFuncSetAttribute(k_elem2<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                    cudaFuncSetAttribute(k_elem2<NN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                }
                k_elem2<NN><<<blocks, kElem2Threads, sm, h->stream>>>(a);
            });
            h->smemAttrSet = true;
            return launch_check(h, "k_elem2");
        }
        DISPATCH_N(h->N, {
            const size_t sm = elem_smem<NN>();
            if (!h->smemAttrSet) {
                cudaFuncSetAttribute(k_elem<NN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                cudaFuncSetAttribute(k_elem<NN, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            }
            k_elem<NN, false><<<blocks, kElemThreads, sm, h->stream>>>(a);
        });
    }
    h->smemAttrSet = true;
    return launch_check(h, "k_elem");
}

// sensor + element -> vertex max merge: phases P1-P2 of StepWorker (euler.go:574-600)
static int run_diss_sensor(dfr2d_handle *h, int rk) {
    DissBuffers &d = h->ds;
    CK(cudaMemsetAsync(d.sigmaV, 0, (size_t)h->NV * sizeof(double), h->stream));
    CK(cudaMemsetAsync(d.epsV, 0, (size_t)h->NV * sizeof(double), h->stream));
    SensorArgs sa{};
    sa.K = h->K; sa.Kp = h->Kp;
    sa.rho = h->q[rk];
    sa.etov = d.etov; sa.hk = d.hk;
    sa.se = d.se; sa.sigma = d.sigma; sa.epsk = d.epsk;
    sa.sigmaV = (unsigned long long *)d.sigmaV; sa.epsV = (unsigned long long *)d.epsV;
    sa.sc = h->sc; sa.par = (int)(h->stepIndex & 1); sa.stepIndex = h->stepIndex; sa.ph = h->ph;
    DISPATCH_N(h->N, (k_sensor<NN><<<(h->K + 127) / 128, 128, 0, h->stream>>>(sa)));
    return launch_check(h, "k_sensor");
}

// vertex -> element mean, (rk == 2) limiter, edge interpolation: phase P3 (euler.go:601-616)
static int run_diss_prepare(dfr2d_handle *h, int rk) {
    DissBuffers &d = h->ds;
    PrepArgs pa{};
    pa.K = h->K; pa.Kp = h->Kp;
    pa.q = h->q[rk]; pa.qface = h->qface;
    pa.etov = d.etov; pa.sigmaV = d.sigmaV; pa.sigma = d.sigma;
    pa.sc = h->sc; pa.rk = rk; pa.par = (int)(h->stepIndex & 1); pa.stepIndex = h->stepIndex; pa.ph = h->ph;
    const int blocks = (h->K + kElemsPerBlock - 1) / kElemsPerBlock;
    DISPATCH_N(h->N, (k_diss_prepare<NN><<<blocks, kElemThreads, 0, h->stream>>>(pa)));
    return launch_check(h, "k_diss_prepare");
}

// RT gradient x epsilon: phase P5 (euler.go:624-628)
static int run_diss_grad(dfr2d_handle *h, int rk) {
    DissBuffers &d = h->ds;
    GradArgs ga{};
    ga.K = h->K; ga.Kp = h->Kp;
    ga.q = h->q[rk]; ga.qface = h->qface;
    ga.etoe = h->etoe; ga.ekL = h->ekL; ga.ekR = h->ekR; ga.emeta = h->emeta;
    ga.Jdet = h->Jdet; ga.Jinv = h->Jinv; ga.IInII = h->IInII; ga.nxk = d.nxk; ga.nyk = d.nyk;
    ga.etov = d.etov; ga.epsV = d.epsV;
    ga.dissX = d.dissX; ga.dissY = d.dissY;
    ga.enx = h->enx; ga.eny = h->eny; ga.vn = d.vn; ga.NEp = h->NEp;
    ga.sc = h->sc; ga.par = (int)(h->stepIndex & 1); ga.stepIndex = h->stepIndex; ga.ph = h->ph;
    const int blocks = (h->K + kElemsPerBlock - 1) / kElemsPerBlock;
    if (h->gradKernel == 4) {
        DISPATCH_N(h->N, {
            using PD = GradPipeDim<NN>;
            using WD = GradWsDim<NN>;
            if (!h->gradWsAttrSet) {
                cudaFuncSetAttribute(k_grad_ws<NN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD::kSmemBytes);
                cudaFuncSetAttribute(k_grad_ws<NN, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_grad_ws<NN, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD::kSmemBytes);
                cudaFuncSetAttribute(k_grad_ws<NN, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                h->gradWsAttrSet = true;
            }
            GradPipeArgs pa{};
            pa.a = ga; pa.table = h->gradTable; pa.mxy = h->gradMxy; pa.nTiles = blocks; pa.skewNs = 0;
            const int grid = std::max(1, std::min(h->sms, (blocks + PD::kGroups - 1) / PD::kGroups));
            if (h->gradMG == 2) k_grad_ws<NN, 2><<<grid, WD::kThreads, PD::kSmemBytes, h->stream>>>(pa);
            else k_grad_ws<NN, 3><<<grid, WD::kThreads, PD::kSmemBytes, h->stream>>>(pa);
        });
        return launch_check(h, "k_grad_ws");
    }
    if (h->gradKernel == 3) {
        DISPATCH_N(h->N, {
            using PD = GradPipeDim<NN>;
            if (!h->gradAttrSet) {
                cudaFuncSetAttribute(k_grad_pipe<NN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD::kSmemBytes);
                cudaFuncSetAttribute(k_grad_pipe<NN, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaFuncSetAttribute(k_grad_pipe<NN, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD::kSmemBytes);
                cudaFuncSetAttribute(k_grad_pipe<NN, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                h->gradAttrSet = true;
            }
            GradPipeArgs pa{};
            pa.a = ga; pa.table = h->gradTable; pa.mxy = h->gradMxy; pa.nTiles = blocks; pa.skewNs = h->gradSkewNs;
            const int grid = std::max(1, std::min(h->sms, (blocks + PD::kGroups - 1) / PD::kGroups));
            if (h->gradMG == 2) k_grad_pipe<NN, 2><<<grid, PD::kThreads, PD::kSmemBytes, h->stream>>>(pa);
            else k_grad_pipe<NN, 3><<<grid, PD::kThreads, PD::kSmemBytes, h->stream>>>(pa);
        });
        return launch_check(h, "k_grad_pipe");
    }
    DISPATCH_N(h->N, {
        const size_t sm = (size_t)4 * (Dim<NN>::NpInt + Dim<NN>::NF3) * kElemsPerBlock * sizeof(double);
        k_grad<NN><<<blocks, kElemThreads, sm, h->stream>>>(ga);
    });
    return launch_check(h, "k_grad");
}

// viscous edge flux: phase P6 (euler.go:629-635)
static int run_diss_visc(dfr2d_handle *h) {
    DissBuffers &d = h->ds;
    ViscEdgeArgs va{};
    va.ne = h->NE; va.NEp = h->NEp; va.Kp = h->Kp;
    va.kL = h->ekL; va.kR = h->ekR; va.meta = h->emeta;
    va.nx = h->enx; va.ny = h->eny; va.oohk = h->eoohk; va.ooLen = d.eooLen;
    va.qface = h->qface; va.vn = d.vn;
    va.etov = d.etov; va.epsV = d.epsV;
    va.eflux = h->eflux; va.aggv = d.aggv;
    va.sc = h->sc; va.slot = (int)(h->stageCounter & 1); va.par = (int)(h->stepIndex & 1); va.stepIndex = h->stepIndex; va.ph = h->ph;
    const int eb = std::max(1, std::min(4 * h->edgeBlocks, (h->NE + kViscEdges - 1) / kViscEdges));
    DISPATCH_N(h->N, (k_visc_edge<NN><<<eb, dim3(kViscEdges, 4), 0, h->stream>>>(va)));
    return launch_check(h, "k_visc_edge");
}

// ---- stage phases (also the multi-process API) ---------------------------------------------------------
// inviscid:     prepare [interp, pack E]  ->E->  edges [unpack E, k_edge]  ->allreduce->  update
// dissipation:  sensor [k_sensor, pack V] ->V->  prepare [unpack V, k_diss_prepare, pack E] ->E->
//               edges [unpack E, k_edge, k_grad, pack D] ->D-> visc [unpack D, k_visc_edge] ->allreduce-> update
static int stage_sensor(dfr2d_handle *h, int rk) {
    if (!h->ph.dissipation) return 0;
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    if (int rc = run_diss_sensor(h, rk)) return rc;
    return run_pack_vertex(h);
}

static int stage_prepare(dfr2d_handle *h, int rk) {
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    if (h->ph.dissipation) {
        if (int rc = run_unpack_vertex(h)) return rc;
        if (int rc = run_diss_prepare(h, rk)) return rc;     // vertex mean, limiter, interpolation
    } else if (!h->qfaceValid) {
        if (int rc = run_interp(h, h->q[rk])) return rc;
    }
    h->qfaceValid = false;
    return run_pack(h);
}

// interior edges only: independent of the halo, so a multi-GPU host calls it between posting the EDGE exchange and
// waiting for it (north star: halo transfer overlapped with interior work); optional -- stage_edges catches up
static int stage_edges_interior(dfr2d_handle *h, int rk) {
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    if (h->interiorDone || !h->edgeSplit) return 0;
    if (h->ph.dissipation && h->edgeViscFused) return 0;      // fused viscous edges run after the gradient (stage_edges)
    if (int rc = run_edges(h, rk, 1)) return rc;
    h->interiorDone = true;
    return 0;
}

static int stage_edges(dfr2d_handle *h, int rk) {
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    if (int rc = run_unpack(h)) return rc;
    if (h->ph.dissipation && h->edgeViscFused) {
        // gradient first (it needs Q_Face only), its cut-edge values go out, and the interior edges -- numerical flux
        // minus viscous flux in one kernel -- overlap that exchange; the boundary / cut-edge list follows in stage_visc
        if (int rc = run_diss_grad(h, rk)) return rc;
        if (int rc = run_pack_diss(h)) return rc;
        return run_edges(h, rk, h->edgeSplit ? 1 : 0);
    }
    if (int rc = run_edges(h, rk, h->interiorDone ? 2 : 3)) return rc;
    h->interiorDone = false;
    if (h->ph.dissipation) {
        if (int rc = run_diss_grad(h, rk)) return rc;
        return run_pack_diss(h);
    }
    return 0;
}

static int stage_visc(dfr2d_handle *h, int rk) {
    if (!h->ph.dissipation) return 0;
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    if (int rc = run_unpack_diss(h)) return rc;
    if (h->edgeViscFused) return run_edges(h, rk, h->edgeSplit ? 2 : 3);
    return run_diss_visc(h);
}

static int stage_update(dfr2d_handle *h, int rk, double *rhsOut) {
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    const bool fuse = (rhsOut == nullptr) && !h->ph.dissipation;
    if (rk == 4 && rhsOut == nullptr && h->captureEdgeQ) {
        // EdgeQValues of the stage that is about to complete (edges.go:344-350), before kernel 5 reuses Q_Face
        const size_t n2 = (size_t)4 * 3 * h->NpEdge * h->Kp / 2;       // Kp is even (padded to the element-block size)
        const int blocks = (int)std::min<size_t>((n2 + 255) / 256, (size_t)h->sms * 8);
        k_capture_qface<<<std::max(blocks, 1), 256, 0, h->stream>>>(n2, (const double2 *)h->qface, (double2 *)h->qfaceSaved, h->sc, h->ph,
                                                                  (int)(h->stepIndex & 1), h->stepIndex);
        if (int rc = launch_check(h, "k_capture_qface")) return rc;
    }
    if (int rc = run_elem(h, rk, rhsOut, fuse)) return rc;
    if (rhsOut == nullptr) {
        h->qfaceValid = fuse;
        h->stageCounter++;
        if (rk == 4) h->stepIndex++;
    }
    return 0;
}

extern "C" int dfr2d_stage_sensor(dfr2d_handle *h, int rk) { return h ? stage_sensor(h, rk) : 1; }
extern "C" int dfr2d_stage_visc(dfr2d_handle *h, int rk) { return h ? stage_visc(h, rk) : 1; }
extern "C" int dfr2d_stage_edges_interior(dfr2d_handle *h, int rk) { return h ? stage_edges_interior(h, rk) : 1; }
extern "C" int dfr2d_stage_prepare(dfr2d_handle *h, int rk) { return h ? stage_prepare(h, rk) : 1; }
extern "C" int dfr2d_stage_edges(dfr2d_handle *h, int rk) { return h ? stage_edges(h, rk) : 1; }
extern "C" int dfr2d_stage_update(dfr2d_handle *h, int rk) { return h ? stage_update(h, rk, nullptr) : 1; }

extern "C" int dfr2d_step_finish(dfr2d_handle *h, dfr2d_step_info *info) {
    if (!h) return 1;
    if (!info) return 0;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->scHost, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    info->time = h->scHost->timeOut;
    info->dt = h->scHost->globalDT;
    info->steps = h->scHost->steps;
    info->finished = h->scHost->finished;
    info->nan_found = h->scHost->nanFlag;
    if (h->scHost->nanFlag == 2) {
        h->err = "peer exchange timed out: a partner partition never delivered its halo / wave-speed message";
        return DFR2D_ERR_PEER;
    }
    if (h->scHost->nanFlag) {
        h->err = "NAN found";
        return DFR2D_ERR_NAN;
    }
    return 0;
}

// {max wave speed, max viscous wave speed} over all partitions (calculateGlobalDT, euler.go:945-971): put + gather over
// the peer mailboxes; the gather is also the stage barrier of the mailbox protocol (dfr2d_peer.cuh).  Two launches so
// that partitions sharing one stream (tests: all partitions on one device) can post every put before the first wait.
static int stage_wave_put(dfr2d_handle *h) {
    if (!h->connected) return 0;
    CK(cudaSetDevice(h->device));
    const int slot = (int)(h->stageCounter & 1);
    k_wave_put<<<1, 32, 0, h->stream>>>(&h->sc->wave[slot][0], h->waveTab, slot, xchg_seq(h));
    return launch_check(h, "k_wave_put");
}
static int stage_wave_gather(dfr2d_handle *h) {
    if (!h->connected) return 0;
    CK(cudaSetDevice(h->device));
    const int slot = (int)(h->stageCounter & 1);
    k_wave_gather<<<1, 32, 0, h->stream>>>(&h->sc->wave[slot][0], h->waveTab, slot, xchg_seq(h));
    return launch_check(h, "k_wave_gather");
}
extern "C" int dfr2d_stage_wave(dfr2d_handle *h, int rk) {
    (void)rk;
    if (!h) return 1;
    if (!h->connected) { h->err = "dfr2d_stage_wave needs dfr2d_peer_connect (unconnected hosts max-reduce dfr2d_wavespeed_buffer)"; return 1; }
    if (int rc = stage_wave_put(h)) return rc;
    return stage_wave_gather(h);
}

static bool host_finished(const dfr2d_handle *h) { return h->stepIndex >= 1 && h->stepIndex >= (long long)h->ph.maxIter; }

extern "C" int dfr2d_step(dfr2d_handle *h, int nsteps, dfr2d_step_info *info) {
    if (!h) return 1;
    if (h->nParts > 1 && !h->connected) {
        h->err = "dfr2d_step on one partition of several needs dfr2d_peer_connect first (or drive the stage calls with "
                 "your own exchange, or use dfr2d_multi_step)";
        return 1;
    }
    for (int s = 0; s < nsteps; s++) {
        if (host_finished(h)) break;      // host-visible half of CheckIfFinished
        for (int rk = 0; rk < 5; rk++) {
            if (int rc = stage_sensor(h, rk)) return rc;              // (+ put VERTEX)
            if (int rc = stage_prepare(h, rk)) return rc;             // (wait VERTEX) ... (+ put EDGE)
            if (h->connected)
                if (int rc = stage_edges_interior(h, rk)) return rc;  // overlaps the halo flight
            if (int rc = stage_edges(h, rk)) return rc;               // (wait EDGE) ... (+ put DISS)
            if (int rc = stage_visc(h, rk)) return rc;                // (wait DISS)
            if (int rc = stage_wave_put(h)) return rc;
            if (int rc = stage_wave_gather(h)) return rc;
            if (int rc = stage_update(h, rk, nullptr)) return rc;
        }
    }
    int rc = dfr2d_step_finish(h, info);
    if (info && host_finished(h)) info->finished = 1;
    return rc;
}

// ---- peer connection -----------------------------------------------------------------------------------------------
// What a partner needs to know about a partition's mailbox.  Fixed-size, plain data: it travels through whatever the host
// has (an all-gather between processes; nothing at all inside one process).
struct PeerBlob {
    uint64_t magic;
    int32_t device, part, nParts, pad;
    int64_t hostPid;
    cudaIpcMemHandle_t mem;
    int64_t words;
    int64_t off[3], offFlags, offWaveFlag, offWaveIn;
    int64_t cnt[3][kMaxParts];          // doubles I receive from partition p per exchange (symmetric: = what I send to p)
};
static_assert(sizeof(PeerBlob) <= DFR2D_PEER_BLOB_BYTES, "PeerBlob must fit the ABI blob");
constexpr uint64_t kPeerMagic = 0x3244524644504231ull;

static void fill_blob(const dfr2d_handle *h, PeerBlob &b) {
    memset(&b, 0, sizeof(b));
    b.magic = kPeerMagic;
    b.device = h->device; b.part = h->part; b.nParts = h->nParts;
    b.hostPid = (int64_t)getpid();
    b.words = (int64_t)h->mailboxWords;
    for (int w = 0; w < 3; w++) b.off[w] = h->mbOff[w];
    b.offFlags = h->mbFlags; b.offWaveFlag = h->mbWaveFlag; b.offWaveIn = h->mbWaveIn;
    for (int p = 0; p < h->nParts; p++) {
        b.cnt[DFR2D_XCHG_EDGE][p] = h->recvCounts[p];
        b.cnt[DFR2D_XCHG_VERTEX][p] = h->vtxCounts[p];
        b.cnt[DFR2D_XCHG_DISS][p] = h->dissCounts[p];
    }
}

extern "C" int dfr2d_peer_export(dfr2d_handle *h, void *blob) {
    if (!h || !blob) return 1;
    if (h->nParts < 2 || !h->mailbox) { h->err = "dfr2d_peer_export: a single partition has no peers"; return 1; }
    CK(cudaSetDevice(h->device));
    PeerBlob b;
    fill_blob(h, b);
    CK(cudaIpcGetMemHandle(&b.mem, h->mailbox));
    memset(blob, 0, DFR2D_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof(b));
    return 0;
}

// base[p] = address of partition p's mailbox as seen from h's device
static int connect_with(dfr2d_handle *h, const PeerBlob *blobs, unsigned long long *const *base) {
    const int n = h->nParts, me = h->part;
    CK(cudaSetDevice(h->device));
    PutTab put[3];
    WaitTab wait[3];
    memset(put, 0, sizeof(put));
    memset(wait, 0, sizeof(wait));
    if (!h->putDone) {
        if (int rc = dev_alloc(h, &h->putDone, 4)) return rc;
        CK(cudaMemset(h->putDone, 0, 4 * sizeof(unsigned int)));
        for (int w = 0; w < 3; w++) {
            if (int rc = dev_alloc(h, &h->putTab[w], 1)) return rc;
            if (int rc = dev_alloc(h, &h->waitTab[w], 1)) return rc;
        }
    }
    for (int w = 0; w < 3; w++) {
        const std::vector<int64_t> &mine = (w == DFR2D_XCHG_EDGE) ? h->sendCounts : (w == DFR2D_XCHG_VERTEX ? h->vtxCounts : h->dissCounts);
        put[w].nParts = wait[w].nParts = n;
        put[w].done = h->putDone + w;
        wait[w].flag = h->mailbox + h->mbFlags + (size_t)w * kMaxParts;
        wait[w].err = &h->sc->nanFlag;
        int64_t acc = 0;
        for (int p = 0; p < n; p++) {
            put[w].first[p] = wait[w].first[p] = acc;
            const int64_t cnt = (p == me) ? 0 : mine[p];
            if (cnt != blobs[p].cnt[w][me]) { h->err = "dfr2d_peer_connect: partner disagrees on the message size (different mesh or partitioning?)"; return 1; }
            if (cnt > 0) {
                int64_t segOff = 0;                   // my segment inside p's receive buffer: behind those of partitions < me
                for (int k = 0; k < me; k++) segOff += blobs[p].cnt[w][k];
                put[w].dst[p] = (double *)(base[p] + blobs[p].off[w]) + segOff;
                put[w].flag[p] = base[p] + blobs[p].offFlags + (size_t)w * kMaxParts + me;
            }
            acc += cnt;
        }
        for (int p = n; p <= kMaxParts; p++) put[w].first[p] = wait[w].first[p] = acc;
        CK(cudaMemcpy(h->putTab[w], &put[w], sizeof(PutTab), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->waitTab[w], &wait[w], sizeof(WaitTab), cudaMemcpyHostToDevice));
    }
    WaveTab &wt = h->waveTab;
    memset(&wt, 0, sizeof(wt));
    wt.nParts = n; wt.me = me;
    wt.myInbox = h->mailbox + h->mbWaveIn;
    wt.myFlag = h->mailbox + h->mbWaveFlag;
    wt.err = &h->sc->nanFlag;
    for (int p = 0; p < n; p++)
        if (p != me) {
            wt.inbox[p] = base[p] + blobs[p].offWaveIn + (size_t)me * 4;
            wt.flag[p] = base[p] + blobs[p].offWaveFlag + me;
        }
    h->connected = true;
    return 0;
}

// Between processes: blobs = [n_parts][DFR2D_PEER_BLOB_BYTES] gathered from every partition (index = partition).  The
// mailboxes are mapped with cudaIpcOpenMemHandle (which also enables peer access); partitions of this very process are
// used directly.  Collective in spirit: every partition must connect before anyone steps.
extern "C" int dfr2d_peer_connect(dfr2d_handle *h, const void *blobs_raw, int n) {
    if (!h || !blobs_raw) return 1;
    if (n != h->nParts || n < 2 || n > kMaxParts) { h->err = "dfr2d_peer_connect: need one blob per partition (2..32)"; return 1; }
    if (h->putDone) { h->connected = true; return 0; }
    CK(cudaSetDevice(h->device));
    std::vector<PeerBlob> blobs((size_t)n);
    std::vector<unsigned long long *> base((size_t)n, nullptr);
    for (int p = 0; p < n; p++) {
        memcpy(&blobs[p], (const char *)blobs_raw + (size_t)p * DFR2D_PEER_BLOB_BYTES, sizeof(PeerBlob));
        if (blobs[p].magic != kPeerMagic || blobs[p].part != p || blobs[p].nParts != n) { h->err = "dfr2d_peer_connect: bad blob"; return 1; }
    }
    for (int p = 0; p < n; p++) {
        if (p == h->part) { base[p] = h->mailbox; continue; }
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, blobs[p].mem, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            h->err = std::string("cudaIpcOpenMemHandle (partition ") + std::to_string(p) + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            return 2;
        }
        h->ipcMapped.push_back(ptr);
        base[p] = (unsigned long long *)ptr;
    }
    return connect_with(h, blobs.data(), base.data());
}

// Switch between the peer-memory exchange (on) and the host-moved exchange of the plain stage API (off) on a handle that
// has been connected; the mappings stay.  Every partition must switch at the same stage boundary.
extern "C" int dfr2d_peer_enable(dfr2d_handle *h, int on) {
    if (!h) return 1;
    if (on && !h->putDone) { h->err = "dfr2d_peer_enable: not connected (dfr2d_peer_connect / dfr2d_multi_step first)"; return 1; }
    h->connected = on != 0;
    return 0;
}

// Inside one process (the Go controller goroutine owning all partitions): plain peer access, no IPC.
static int connect_local(dfr2d_handle **hs, int n) {
    std::vector<PeerBlob> blobs((size_t)n);
    std::vector<unsigned long long *> base((size_t)n, nullptr);
    for (int i = 0; i < n; i++) {
        dfr2d_handle *h = hs[i];
        if (!h || h->nParts != n || h->part != i) return 1;
        if (!h->mailbox) { h->err = "dfr2d_multi_step: handle has no mailbox"; return 1; }
        fill_blob(h, blobs[i]);
        base[i] = h->mailbox;
    }
    for (int i = 0; i < n; i++) {
        dfr2d_handle *h = hs[i];
        if (h->putDone) { h->connected = true; continue; }
        CK(cudaSetDevice(h->device));
        for (int j = 0; j < n; j++)
            if (hs[j]->device != h->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, h->device, hs[j]->device);
                if (!can) { h->err = "dfr2d_multi_step needs peer access between the devices"; return 2; }
                cudaError_t e = cudaDeviceEnablePeerAccess(hs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->err = cudaGetErrorString(e); return 2; }
                cudaGetLastError();
            }
        if (int rc = connect_with(h, blobs.data(), base.data())) return rc;
    }
    return 0;
}

// ---- single-process multi-GPU driver ------------------------------------------------------------------------------
// The same stage protocol as dfr2d_step, issued phase by phase over all partitions by the one calling thread: every
// partition's put of an exchange is enqueued before any partition's wait for it, so partitions that share a device (and
// hence a stream) cannot deadlock, and partitions on different devices run concurrently.  prof (optional): CUDA events
// at the phase boundaries of one step on every partition's stream.
struct MultiProf {
    std::vector<cudaEvent_t> ev;       // [n][5 stages][kProfPhases + 1]
};
constexpr int kProfPhases = 6;         // sensor+prepare(pack,put) | interior edges | halo wait+cut/boundary edges(+grad) | visc | wave | update

static int multi_step_impl(dfr2d_handle **hs, int n, int nsteps, dfr2d_step_info *info, MultiProf *prof) {
    if (!hs || n < 1 || n > kMaxParts) return 1;
    if (n == 1) return dfr2d_step(hs[0], nsteps, info);
    if (int rc = connect_local(hs, n)) return rc;
    auto mark = [&](int g, int rk, int ph) -> int {
        if (!prof) return 0;
        dfr2d_handle *h = hs[g];
        CK(cudaSetDevice(h->device));
        CK(cudaEventRecord(prof->ev[((size_t)g * 5 + rk) * (kProfPhases + 1) + ph], h->stream));
        return 0;
    };
#define MULTI_ALL(call, ph)                              \
    for (int g = 0; g < n; g++) {                        \
        if (int rc = call) return rc;                    \
        if (ph >= 0) if (int rc = mark(g, rk, ph)) return rc; \
    }
    for (int s = 0; s < nsteps; s++) {
        if (host_finished(hs[0])) break;
        for (int rk = 0; rk < 5; rk++) {
            MULTI_ALL(mark(g, rk, 0), -1);
            MULTI_ALL(stage_sensor(hs[g], rk), -1);
            MULTI_ALL(stage_prepare(hs[g], rk), 1);
            MULTI_ALL(stage_edges_interior(hs[g], rk), 2);
            MULTI_ALL(stage_edges(hs[g], rk), 3);
            MULTI_ALL(stage_visc(hs[g], rk), 4);
            MULTI_ALL(stage_wave_put(hs[g]), -1);
            MULTI_ALL(stage_wave_gather(hs[g]), 5);
            MULTI_ALL(stage_update(hs[g], rk, nullptr), 6);
        }
    }
#undef MULTI_ALL
    if (!info) return 0;                              // like dfr2d_step: no host synchronisation without info
    int rc = 0;
    for (int g = n - 1; g >= 0; g--) {
        dfr2d_step_info tmp{};
        int r = dfr2d_step_finish(hs[g], &tmp);       // synchronises partition g; surfaces "NAN found"
        if (r) rc = r;
        if (g == 0) {
            *info = tmp;
            if (host_finished(hs[0])) info->finished = 1;
        }
    }
    return rc;
}

// SURVEY.md 8(b)'s shape of the boundary -- one create for the whole run, the fan-out over GPUs inside the library:
// hs_out[g] = partition g of n_parts on devices[g] (NULL: g modulo the number of visible devices).
extern "C" int dfr2d_multi_create(const dfr2d_problem *p, int n_parts, const int *devices, dfr2d_handle **hs_out) {
    if (!p || !hs_out || n_parts < 1 || n_parts > kMaxParts) { g_create_error = "dfr2d_multi_create: bad arguments"; return 1; }
    int ndev = 0;
    if (!devices) {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev < 1) {
            g_create_error = std::string("dfr2d_multi_create: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device");
            return 2;
        }
    }
    for (int g = 0; g < n_parts; g++) hs_out[g] = nullptr;
    for (int g = 0; g < n_parts; g++) {
        int rc = dfr2d_create(p, n_parts, g, devices ? devices[g] : g % ndev, &hs_out[g]);
        if (rc) {
            for (int j = 0; j < g; j++) { dfr2d_destroy(hs_out[j]); hs_out[j] = nullptr; }
            return rc;
        }
    }
    return 0;
}

extern "C" void dfr2d_multi_destroy(dfr2d_handle **hs, int n) {
    if (!hs) return;
    for (int g = 0; g < n; g++) { dfr2d_destroy(hs[g]); hs[g] = nullptr; }
}

// PrintUpdate's loop over the partitions (euler.go:823-829) without its clamp at zero: the signed maximum per variable
extern "C" int dfr2d_multi_residual(dfr2d_handle **hs, int n, double maxR[4]) {
    if (!hs || n < 1 || !maxR) return 1;
    for (int g = 0; g < n; g++) {
        double r[4];
        if (int rc = dfr2d_residual(hs[g], r)) return rc;
        for (int v = 0; v < 4; v++)
            if (g == 0 || r[v] > maxR[v]) maxR[v] = r[v];
    }
    return 0;
}

extern "C" int dfr2d_multi_step(dfr2d_handle **hs, int n, int nsteps, dfr2d_step_info *info) {
    return multi_step_impl(hs, n, nsteps, info, nullptr);
}

// One profiled step: ms_out[n][5][6] = duration of each phase of each stage on each partition (CUDA events on the
// partition's stream; waiting for a partner's halo or wave value is inside the phase that waits).
extern "C" int dfr2d_multi_step_profile(dfr2d_handle **hs, int n, float *ms_out) {
    if (!hs || n < 1 || n > kMaxParts || !ms_out) return 1;
    MultiProf prof;
    prof.ev.assign((size_t)n * 5 * (kProfPhases + 1), nullptr);
    int rc = 0;
    for (int g = 0; g < n && !rc; g++) {
        dfr2d_handle *h = hs[g];
        if (!h) { rc = 1; break; }
        if (cudaSetDevice(h->device) != cudaSuccess) { rc = 2; break; }
        for (int i = 0; i < 5 * (kProfPhases + 1); i++)
            if (cudaEventCreate(&prof.ev[(size_t)g * 5 * (kProfPhases + 1) + i]) != cudaSuccess) { rc = 2; break; }
    }
    if (!rc) rc = multi_step_impl(hs, n, 1, nullptr, n > 1 ? &prof : nullptr);
    if (!rc && n > 1)
        for (int g = 0; g < n; g++) {
            cudaSetDevice(hs[g]->device);
            cudaStreamSynchronize(hs[g]->stream);
            for (int rk = 0; rk < 5; rk++)
                for (int ph = 0; ph < kProfPhases; ph++) {
                    const size_t i = ((size_t)g * 5 + rk) * (kProfPhases + 1) + ph;
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, prof.ev[i], prof.ev[i + 1]);
                    ms_out[((size_t)g * 5 + rk) * kProfPhases + ph] = ms;
                }
        }
    for (size_t i = 0; i < prof.ev.size(); i++)
        if (prof.ev[i]) cudaEventDestroy(prof.ev[i]);
    cudaGetLastError();
    return rc;
}

// Scratch that outlives a call: grown on demand, freed with the handle (no cudaMalloc/cudaFree -- and the implicit device
// synchronisation they bring -- on calls the host makes every few steps).
static int scratch_reserve(dfr2d_handle *h, size_t bytes) {
    if (bytes <= h->scratchBytes) return 0;
    if (h->scratch) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaFree(h->scratch));
        h->scratch = nullptr;
        h->scratchBytes = 0;
    }
    bytes = (bytes + 255) & ~(size_t)255;
    CK(cudaMalloc(&h->scratch, bytes));
    h->scratchBytes = bytes;
    return 0;
}

extern "C" int dfr2d_rhs(dfr2d_handle *h, int rk, double *RHS_out) {
    if (!h || !RHS_out || rk < 0 || rk > 4) return 1;
    if (h->nParts > 1) { h->err = "dfr2d_rhs is a single-partition test hook"; return 1; }
    CK(cudaSetDevice(h->device));
    const size_t regBytes = (size_t)4 * h->NpInt * h->Kp * sizeof(double);
    if (!h->rhsScratch) {
        if (int rc = dev_alloc(h, &h->rhsScratch, (size_t)4 * h->NpInt * h->Kp)) return rc;
    }
    // once CheckIfFinished holds every launch is a no-op: report that instead of copying out stale scratch
    if (h->stepIndex >= 1) {
        dfr2d_step_info fin{};
        if (int rc = dfr2d_step_finish(h, &fin)) return rc;
        if (fin.finished || h->stepIndex >= (long long)h->ph.maxIter) {
            h->err = "dfr2d_rhs: the run has finished (FinalTime or MaxIterations reached); no RHS is evaluated";
            return 1;
        }
    }
    CK(cudaMemsetAsync(h->rhsScratch, 0, regBytes, h->stream));
    // the hook must not leak its wave-speed maxima into the next real stage: the edge kernels atomicMax into
    // wave[stageCounter & 1] and the rhsOut path of the element kernel returns before that slot is reset
    unsigned long long *slot = &h->sc->wave[h->stageCounter & 1][0];
    CK(cudaMemsetAsync(slot, 0, 2 * sizeof(unsigned long long), h->stream));
    // with the limiter, stage 2 filters its input register in place (euler.go:605-609); the hook evaluates it on the
    // caller's register like the reference would, then puts the unfiltered register back ("without advancing")
    double *saved = nullptr;
    if (h->ph.dissipation && rk == 2) {
        if (int rc = scratch_reserve(h, regBytes)) return rc;
        saved = (double *)h->scratch;
        CK(cudaMemcpyAsync(saved, h->q[rk], regBytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    h->qfaceValid = false;
    if (int rc = stage_sensor(h, rk)) return rc;
    if (int rc = stage_prepare(h, rk)) return rc;
    if (int rc = stage_edges(h, rk)) return rc;
    if (int rc = stage_visc(h, rk)) return rc;
    if (int rc = stage_update(h, rk, h->rhsScratch)) return rc;
    h->qfaceValid = false;
    CK(cudaMemsetAsync(slot, 0, 2 * sizeof(unsigned long long), h->stream));
    if (saved) CK(cudaMemcpyAsync(h->q[rk], saved, regBytes, cudaMemcpyDeviceToDevice, h->stream));
    return copy_out(h, h->rhsScratch, RHS_out);
}

extern "C" int dfr2d_residual(dfr2d_handle *h, double maxR[4]) {
    if (!h || !maxR) return 1;
    CK(cudaSetDevice(h->device));
    if (h->resInScalars) {
        unsigned long long enc[4];
        CK(cudaMemcpyAsync(enc, &h->sc->resMax[0], sizeof(enc), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int n = 0; n < 4; n++) maxR[n] = res_decode(enc[n]);
        return 0;
    }
    if (int rc = scratch_reserve(h, 4 * sizeof(double))) return rc;
    double *tmp = (double *)h->scratch;
    k_signed_max<<<4, 1024, 0, h->stream>>>(h->R, h->NpInt, h->K, h->Kp, tmp);
    if (int rc = launch_check(h, "k_signed_max")) return rc;
    CK(cudaMemcpyAsync(maxR, tmp, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dfr2d_get_field(dfr2d_handle *h, int which, double *out) {
    if (!h || !out) return 1;
    CK(cudaSetDevice(h->device));
    const double *src = nullptr;
    switch (which) {
        case DFR2D_FIELD_DT: src = h->DT; break;
        case DFR2D_FIELD_SigmaScalar: src = h->ds.sigma; break;
        case DFR2D_FIELD_EpsilonScalar: src = h->ds.epsk; break;
        case DFR2D_FIELD_Se: src = h->ds.se; break;
        default: break;
    }
    if (!src) { h->err = "field not available for this configuration"; return 1; }
    CK(cudaMemcpyAsync(out + h->hostOff, src, (size_t)h->K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dfr2d_init_state(dfr2d_handle *h, int init_case, int64_t nv, const double *VX, const double *VY,
                                const int32_t *EToV, const double *R, const double *S) {
    if (!h || !VX || !VY || !EToV || !R || !S || nv <= 0) return 1;
    if (init_case < DFR2D_CASE_Freestream || init_case > DFR2D_CASE_ShockTube) { h->err = "unknown case type"; return 1; }
    CK(cudaSetDevice(h->device));
    // one scratch block: vx | vy | r,s | etov
    const size_t oVy = (size_t)nv, oRs = 2 * (size_t)nv, oEt = oRs + 2 * (size_t)h->NpInt;
    const size_t bytes = oEt * sizeof(double) + (size_t)3 * std::max(h->K, 1) * sizeof(int);
    if (int rc = scratch_reserve(h, bytes)) return rc;
    double *vx = (double *)h->scratch, *vy = vx + oVy, *rs = vx + oRs;
    int *etov = (int *)(vx + oEt);
    CK(cudaMemcpyAsync(vx, VX, (size_t)nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(vy, VY, (size_t)nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(rs, R, (size_t)h->NpInt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(rs + h->NpInt, S, (size_t)h->NpInt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(etov, EToV + 3 * h->hostOff, (size_t)3 * h->K * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    InitArgs ia{};
    ia.K = h->K; ia.Kp = h->Kp; ia.npInt = h->NpInt; ia.initCase = init_case;
    ia.vx = vx; ia.vy = vy; ia.etov = etov; ia.r = rs; ia.s = rs + h->NpInt; ia.q = h->q[0]; ia.ph = h->ph;
    k_init_state<<<(h->K + 127) / 128, 128, 0, h->stream>>>(ia);
    if (int rc = launch_check(h, "k_init_state")) return rc;
    h->qfaceValid = false;
    CK(cudaStreamSynchronize(h->stream));       // the host arrays are only valid during the call (cgo pointer rules)
    return 0;
}

extern "C" int dfr2d_plot_field(dfr2d_handle *h, int flow_function, const double *graph_interp, int np_graph, float *out) {
    if (!h || !graph_interp || !out) return 1;
    const int NG = 3 * (1 + h->NpEdge) + h->NpInt;
    if (np_graph != NG) { h->err = "GraphInterp must have 3(1+NpEdge)+NpInt rows (GetRSForGraphMesh)"; return 1; }
    if ((flow_function < 0 || flow_function > 13) && flow_function != 100) {
        h->err = "dfr2d_plot_field evaluates the GetFlowFunction family (Density..Entropy = 0..13, fluids.go:209-223) and "
                 "ShockFunction (100); the epsilon fields are dfr2d_epsilon_field";
        return 1;
    }
    CK(cudaSetDevice(h->device));
    if (flow_function == 100)
        if (int rc = ensure_ops(h)) return rc;
    const size_t nOut = (size_t)h->K * NG;
    const size_t giBytes = ((size_t)NG * h->NpInt * sizeof(double) + 255) & ~(size_t)255;
    if (int rc = scratch_reserve(h, giBytes + std::max<size_t>(nOut, 1) * sizeof(float))) return rc;
    double *gi = (double *)h->scratch;
    float *dout = (float *)((char *)h->scratch + giBytes);
    CK(cudaMemcpyAsync(gi, graph_interp, (size_t)NG * h->NpInt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PlotArgs pa{};
    pa.K = h->K; pa.Kp = h->Kp; pa.ff = flow_function;
    pa.q = h->q[0]; pa.gi = gi; pa.out = dout;
    pa.gamma = h->ph.fs[0].Gamma; pa.Pinf = h->ph.fs[0].Pinf; pa.QQinf = h->ph.fs[0].QQinf;
    pa.kappa = h->ph.dissipation ? h->kappaGiven : 2.0;      // c.ShockFinder (euler.go:82) or NewAliasShockFinder(2) (plot.go:36)
    const int blocks = (h->K + kPlotThreads - 1) / kPlotThreads;
    DISPATCH_N(h->N, (k_plot_field<NN><<<blocks, kPlotThreads, 0, h->stream>>>(pa)));
    if (int rc = launch_check(h, "k_plot_field")) return rc;
    CK(cudaMemcpyAsync(out + (size_t)h->hostOff * NG, dout, nOut * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dfr2d_epsilon_field(dfr2d_handle *h, int c0, double *out) {
    if (!h || !out) return 1;
    if (!h->ph.dissipation) { h->err = "the epsilon fields exist only with the PerssonC0 limiter (c.Dissipation != nil)"; return 1; }
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    const size_t n = (size_t)h->NpFlux * std::max(h->K, 1);
    if (int rc = scratch_reserve(h, n * sizeof(double))) return rc;
    double *tmp = (double *)h->scratch;
    DISPATCH_N(h->N, (k_epsilon_field<NN><<<(h->K + 127) / 128, 128, 0, h->stream>>>(h->K, h->Kp, c0, h->ds.epsk, h->ds.epsV, h->ds.etov, tmp)));
    if (int rc = launch_check(h, "k_epsilon_field")) return rc;
    CK(cudaMemcpy2DAsync(out + h->hostOff, (size_t)h->hostPitch * sizeof(double), tmp, (size_t)h->K * sizeof(double),
                         (size_t)h->K * sizeof(double), (size_t)h->NpFlux, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- gradient plot fields --------------------------------------------------------------------------------------------
extern "C" int dfr2d_capture_edge_values(dfr2d_handle *h, int on, const double *FaceNormX, const double *FaceNormY) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    if (!on) { h->captureEdgeQ = false; return 0; }
    if (!h->qfaceSaved) {
        const size_t n = (size_t)4 * 3 * h->NpEdge * h->Kp;
        if (h->Kp % 2) { h->err = "internal: odd column padding"; return 1; }
        if (int rc = dev_alloc(h, &h->qfaceSaved, n)) return rc;
        CK(cudaMemsetAsync(h->qfaceSaved, 0, n * sizeof(double), h->stream));
    }
    if (!h->nxkPlot) {
        if (h->ph.dissipation) {
            h->nxkPlot = h->ds.nxk; h->nykPlot = h->ds.nyk;
        } else {
            if (!FaceNormX || !FaceNormY) {
                h->err = "dfr2d_capture_edge_values: DFR.FaceNorm[0|1] are needed once (the handle keeps element normals only "
                         "with the limiter)";
                return 1;
            }
            std::vector<double> nxk((size_t)3 * h->Kp, 0.0), nyk((size_t)3 * h->Kp, 0.0);
            for (int le = 0; le < 3; le++)
                for (int k = 0; k < h->K; k++) {
                    nxk[(size_t)le * h->Kp + k] = FaceNormX[(size_t)le * h->hostPitch + h->hostOff + k];
                    nyk[(size_t)le * h->Kp + k] = FaceNormY[(size_t)le * h->hostPitch + h->hostOff + k];
                }
            if (int rc = dev_upload(h, &h->nxkPlot, nxk)) return rc;
            if (int rc = dev_upload(h, &h->nykPlot, nyk)) return rc;
        }
    }
    h->captureEdgeQ = true;
    return 0;
}

extern "C" int dfr2d_gradient_field(dfr2d_handle *h, int flow_function, double *out) {
    if (!h || !out) return 1;
    const bool isX = flow_function >= 200 && flow_function <= 203, isY = flow_function >= 300 && flow_function <= 303;
    if (!isX && !isY) { h->err = "dfr2d_gradient_field: flow_function must be 200..203 (XGradient*) or 300..303 (YGradient*)"; return 1; }
    if (!h->qfaceSaved || !h->nxkPlot) {
        h->err = "dfr2d_gradient_field: the edge values of the last stage are not kept by default; call "
                 "dfr2d_capture_edge_values(h, 1, ...) before the step whose fields are plotted";
        return 1;
    }
    if (!h->haveDiv) { h->err = "dfr2d_gradient_field: dfr2d_problem.Div (DFR.FluxElement.Div) was not given at create"; return 1; }
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_ops(h)) return rc;
    const size_t n = (size_t)h->NpFlux * std::max(h->K, 1);
    if (int rc = scratch_reserve(h, n * sizeof(double))) return rc;
    GradPlotArgs a{};
    a.K = h->K; a.Kp = h->Kp; a.var = flow_function % 100; a.dirY = isY ? 1 : 0;
    a.q = h->q[0]; a.qfaceSaved = h->qfaceSaved;
    a.etoe = h->etoe; a.ekL = h->ekL; a.emeta = h->emeta;
    a.Jdet = h->Jdet; a.Jinv = h->Jinv; a.IInII = h->IInII; a.nk = isY ? h->nykPlot : h->nxkPlot;
    a.grad = (double *)h->scratch;
    if (h->K > 0) {
        DISPATCH_N(h->N, (k_grad_plot<NN><<<(h->K + 127) / 128, 128, 0, h->stream>>>(a)));
        if (int rc = launch_check(h, "k_grad_plot")) return rc;
    }
    CK(cudaMemcpy2DAsync(out + h->hostOff, (size_t)h->hostPitch * sizeof(double), a.grad, (size_t)h->K * sizeof(double),
                         (size_t)h->K * sizeof(double), (size_t)h->NpFlux, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dfr2d_set_stream(dfr2d_handle *h, void *s) {
    if (!h) return 1;
    h->stream = (cudaStream_t)s;
    return 0;
}
extern "C" int dfr2d_partition_range(const dfr2d_handle *h, int64_t *b, int64_t *e) {
    if (!h) return 1;
    if (b) *b = h->k0;
    if (e) *e = h->k1;
    return 0;
}
extern "C" int dfr2d_halo_counts(const dfr2d_handle *h, int64_t *s, int64_t *r) {
    if (!h) return 1;
    for (int i = 0; i < h->nParts; i++) {
        if (s) s[i] = h->sendCounts[i];
        if (r) r[i] = h->recvCounts[i];
    }
    return 0;
}
extern "C" int dfr2d_halo_buffers(dfr2d_handle *h, void **s, void **r) {
    if (!h) return 1;
    if (s) *s = h->sendBuf;
    if (r) *r = h->recvBuf;
    return 0;
}
extern "C" int dfr2d_exchange_counts(const dfr2d_handle *h, int which, int64_t *s, int64_t *r) {
    if (!h || which < 0 || which > 2) return 1;
    const std::vector<int64_t> &c = which == DFR2D_XCHG_EDGE ? h->sendCounts : (which == DFR2D_XCHG_VERTEX ? h->vtxCounts : h->dissCounts);
    for (int i = 0; i < h->nParts; i++) {      // every exchange is symmetric: what goes to a peer comes back from it
        if (s) s[i] = c[i];
        if (r) r[i] = c[i];
    }
    return 0;
}
extern "C" int dfr2d_exchange_buffers(dfr2d_handle *h, int which, void **s, void **r) {
    if (!h || which < 0 || which > 2) return 1;
    if (s) *s = which == DFR2D_XCHG_EDGE ? h->sendBuf : (which == DFR2D_XCHG_VERTEX ? h->vSendBuf : h->dSendBuf);
    if (r) *r = which == DFR2D_XCHG_EDGE ? h->recvBuf : (which == DFR2D_XCHG_VERTEX ? h->vRecvBuf : h->dRecvBuf);
    return 0;
}
extern "C" int dfr2d_wavespeed_buffer(dfr2d_handle *h, void **p) {
    if (!h || !p) return 1;
    *p = (void *)&h->sc->wave[h->stageCounter & 1][0];
    return 0;
}
extern "C" int64_t dfr2d_launch_count(const dfr2d_handle *h) { return h ? h->launches : 0; }

// ---- host-only plan API (CPU tests of the multi-partition bookkeeping) ---------------------------------------
static int plan_create_common(const dfr2d_problem *p, int64_t K_global, int64_t k_offset, int n_parts, int part, dfr2d_plan **out) {
    if (!p || !out || n_parts < 1 || part < 0 || part >= n_parts || K_global < n_parts || k_offset < 0 || k_offset + p->K > K_global) {
        g_create_error = "bad plan request";
        return 1;
    }
    dfr2d_plan *pl = new dfr2d_plan();
    pl->nParts = n_parts; pl->part = part; pl->Kglobal = K_global; pl->kOff = k_offset;
    int rc = build_plan(p, *pl);
    if (rc) { g_create_error = pl->err; delete pl; return rc; }
    *out = pl;
    return 0;
}
extern "C" int dfr2d_plan_create(const dfr2d_problem *p, int n_parts, int part, dfr2d_plan **out) {
    return plan_create_common(p, p ? p->K : 0, 0, n_parts, part, out);
}
extern "C" int dfr2d_plan_create_window(const dfr2d_problem *p, int64_t K_global, int64_t k_offset, int n_parts, int part,
                                        dfr2d_plan **out) {
    return plan_create_common(p, K_global, k_offset, n_parts, part, out);
}
extern "C" void dfr2d_plan_destroy(dfr2d_plan *pl) { delete pl; }
extern "C" int dfr2d_plan_sizes(const dfr2d_plan *pl, int64_t out[8]) {
    if (!pl || !out) return 1;
    out[0] = pl->k0; out[1] = pl->k1; out[2] = pl->G; out[3] = pl->Kp; out[4] = pl->NE; out[5] = pl->NEp;
    out[6] = pl->nSendEdges; out[7] = pl->NBP;
    return 0;
}
extern "C" int dfr2d_plan_edges(const dfr2d_plan *pl, int32_t *kL, int32_t *kR, int32_t *meta, int64_t *global_edge, int32_t *etoe) {
    if (!pl) return 1;
    for (int s = 0; s < pl->NE; s++) {
        if (kL) kL[s] = pl->ekL[s];
        if (kR) kR[s] = pl->ekR[s];
        if (meta) meta[s] = pl->emeta[s];
        if (global_edge) global_edge[s] = pl->edgeGlobal[s];
    }
    if (etoe) memcpy(etoe, pl->etoe.data(), pl->etoe.size() * sizeof(int));
    return 0;
}
extern "C" int dfr2d_plan_halo(const dfr2d_plan *pl, int64_t *send_counts, int64_t *recv_counts, int64_t *ghost_global,
                               int32_t *send_elem, int32_t *send_row0, int32_t *recv_col, int32_t *recv_row0) {
    if (!pl) return 1;
    for (int i = 0; i < pl->nParts; i++) {
        if (send_counts) send_counts[i] = pl->sendCounts[i];
        if (recv_counts) recv_counts[i] = pl->recvCounts[i];
    }
    for (int g = 0; g < pl->G; g++) if (ghost_global) ghost_global[g] = pl->ghostGlobal[g];
    for (int c = 0; c < pl->nSendEdges; c++) {
        if (send_elem) send_elem[c] = pl->sendElem[c];
        if (send_row0) send_row0[c] = pl->sendRow0[c];
        if (recv_col) recv_col[c] = pl->recvCol[c];
        if (recv_row0) recv_row0[c] = pl->recvRow0[c];
    }
    return 0;
}

extern "C" int dfr2d_plan_vertices(const dfr2d_plan *pl, int64_t *counts, int32_t *vertex_ids) {
    if (!pl) return 1;
    for (int i = 0; i < pl->nParts; i++) if (counts) counts[i] = pl->vtxCounts[i];
    if (vertex_ids) for (size_t i = 0; i < pl->vtxList.size(); i++) vertex_ids[i] = pl->vtxList[i];
    return 0;
}

// ---- element renumbering for partition quality -------------------------------------------------
// Reverse Cuthill-McKee over the element adjacency graph (elements sharing an edge).  PartitionMap.Split1D cuts the
// element index range into contiguous pieces, so the cut size of a mesh is decided by its numbering: mesh-generator
// numbering of the shipped NACA meshes is essentially random in space.  The host renumbers the elements with the
// returned order BEFORE building the DG2D tables; nothing in the time loop changes.
static bool grad_table_for(int N, const double *Div, const double *Bary, std::vector<double> &tb) {
    switch (N) {
        case 0: build_grad_table<0>(Div, Bary, tb); return true;
        case 1: build_grad_table<1>(Div, Bary, tb); return true;
        case 2: build_grad_table<2>(Div, Bary, tb); return true;
        case 3: build_grad_table<3>(Div, Bary, tb); return true;
        case 4: build_grad_table<4>(Div, Bary, tb); return true;
        default: return false;
    }
}

extern "C" int64_t dfr2d_grad_mma_table(int N, const double *Div, const double *Bary, double *out, int64_t cap) {
    std::vector<double> tb;
    if (!Div || !Bary || !grad_table_for(N, Div, Bary, tb)) return -1;
    if (out)
        for (int64_t i = 0; i < cap && i < (int64_t)tb.size(); i++) out[i] = tb[i];
    return (int64_t)tb.size();
}

extern "C" int64_t dfr2d_mma_diss_table(int N, const double *DivInt, const double *Vinv, const double *V, double *out, int64_t cap) {
    std::vector<double> fr;
    if (!DivInt || !Vinv || !V) return -1;
    switch (N) {
        case 0: build_mma_diss_frags<0>(DivInt, Vinv, V, fr); break;
        case 1: build_mma_diss_frags<1>(DivInt, Vinv, V, fr); break;
        case 2: build_mma_diss_frags<2>(DivInt, Vinv, V, fr); break;
        case 3: build_mma_diss_frags<3>(DivInt, Vinv, V, fr); break;
        case 4: build_mma_diss_frags<4>(DivInt, Vinv, V, fr); break;
        default: return -1;
    }
    if (out)
        for (int64_t i = 0; i < cap && i < (int64_t)fr.size(); i++) out[i] = fr[i];
    return (int64_t)fr.size();
}

extern "C" int dfr2d_rcm_order(int64_t K, int64_t NE, const int32_t *edge_kL, const int32_t *edge_kR, const int32_t *edge_nconn,
                               int32_t *order) {
    if (K <= 0 || NE < 0 || !edge_kL || !edge_kR || !edge_nconn || !order) { g_create_error = "bad rcm request"; return 1; }
    std::vector<int> deg((size_t)K, 0), start((size_t)K + 1, 0), adj;
    for (int64_t e = 0; e < NE; e++)
        if (edge_nconn[e] == 2) { deg[edge_kL[e]]++; deg[edge_kR[e]]++; }
    for (int64_t k = 0; k < K; k++) start[k + 1] = start[k] + deg[k];
    adj.assign((size_t)start[K], 0);
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int64_t e = 0; e < NE; e++)
            if (edge_nconn[e] == 2) {
                adj[fill[edge_kL[e]]++] = edge_kR[e];
                adj[fill[edge_kR[e]]++] = edge_kL[e];
            }
    }
    std::vector<char> seen((size_t)K, 0);
    std::vector<int> level((size_t)K, -1), out, queue;
    out.reserve((size_t)K);
    auto bfs_far = [&](int root) {            // last node of a BFS from root with minimal degree in the last level
        queue.assign(1, root);
        std::vector<int> touched(1, root);
        level[root] = 0;
        int far = root;
        for (size_t qi = 0; qi < queue.size(); qi++) {
            const int u = queue[qi];
            if (level[u] > level[far] || (level[u] == level[far] && deg[u] < deg[far])) far = u;
            for (int t = start[u]; t < start[u + 1]; t++) {
                const int w = adj[t];
                if (!seen[w] && level[w] < 0) { level[w] = level[u] + 1; queue.push_back(w); touched.push_back(w); }
            }
        }
        for (int t : touched) level[t] = -1;
        return far;
    };
    for (int64_t s0 = 0; s0 < K; s0++) {
        if (seen[s0]) continue;
        int root = (int)s0;
        for (int it = 0; it < 3; it++) root = bfs_far(root);       // pseudo-peripheral start node
        const size_t first = out.size();
        out.push_back(root);
        seen[root] = 1;
        std::vector<int> nb;
        for (size_t qi = first; qi < out.size(); qi++) {
            const int u = out[qi];
            nb.clear();
            for (int t = start[u]; t < start[u + 1]; t++)
                if (!seen[adj[t]]) { seen[adj[t]] = 1; nb.push_back(adj[t]); }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
            out.insert(out.end(), nb.begin(), nb.end());
        }
        std::reverse(out.begin() + (std::ptrdiff_t)first, out.end());
    }
    for (int64_t k = 0; k < K; k++) order[k] = out[(size_t)k];
    return 0;
}

// Host-only: elements ordered along a Hilbert curve through their centroids.  The coordinates are replaced by their RANKS
// first (position in the sorted list, scaled to 16 bits), so the curve resolves the thin cells at an airfoil surface as
// well as the far field; ties are broken by the element index, which makes the order a pure function of the mesh.
extern "C" int dfr2d_hilbert_order(int64_t K, int64_t NV, const int32_t *EToV, const double *VX, const double *VY, int32_t *order) {
    if (K <= 0 || NV <= 0 || !EToV || !VX || !VY || !order) { g_create_error = "bad hilbert request"; return 1; }
    std::vector<double> cx((size_t)K), cy((size_t)K);
    for (int64_t k = 0; k < K; k++) {
        double sx = 0.0, sy = 0.0;
        for (int v = 0; v < 3; v++) {
            const int32_t id = EToV[3 * k + v];
            if (id < 0 || id >= NV) { g_create_error = "hilbert request: vertex id out of range"; return 1; }
            sx += VX[id]; sy += VY[id];
        }
        cx[(size_t)k] = sx; cy[(size_t)k] = sy;      // 3 x centroid: the ranks are the same
    }
    constexpr int64_t kSide = 65536;
    auto ranks = [&](const std::vector<double> &c, std::vector<int64_t> &r) {
        std::vector<int32_t> idx((size_t)K);
        for (int64_t k = 0; k < K; k++) idx[(size_t)k] = (int32_t)k;
        std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return c[a] != c[b] ? c[a] < c[b] : a < b; });
        r.resize((size_t)K);
        for (int64_t pos = 0; pos < K; pos++) r[(size_t)idx[(size_t)pos]] = pos * (kSide - 1) / std::max<int64_t>(K - 1, 1);
    };
    std::vector<int64_t> rx, ry, d((size_t)K);
    ranks(cx, rx);
    ranks(cy, ry);
    for (int64_t k = 0; k < K; k++) {
        int64_t x = rx[(size_t)k], y = ry[(size_t)k], dd = 0;
        for (int64_t sft = kSide / 2; sft > 0; sft /= 2) {
            const int64_t bx = (x & sft) ? 1 : 0, by = (y & sft) ? 1 : 0;
            dd += sft * sft * ((3 * bx) ^ by);
            if (by == 0) {                       // rotate the quadrant
                if (bx == 1) { x = kSide - 1 - x; y = kSide - 1 - y; }
                std::swap(x, y);
            }
        }
        d[(size_t)k] = dd;
    }
    std::vector<int32_t> idx((size_t)K);
    for (int64_t k = 0; k < K; k++) idx[(size_t)k] = (int32_t)k;
    std::sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return d[a] != d[b] ? d[a] < d[b] : a < b; });
    for (int64_t k = 0; k < K; k++) order[k] = idx[(size_t)k];
    return 0;
}

#ifdef DFR2D_PIPE_TIMING
extern "C" int dfr2d_debug_pipe_clocks(unsigned long long out[8], int reset) {
    cudaMemcpyFromSymbol(out, g_pipe_clk, 8 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_pipe_clk, z, sizeof(z)); }
    return 0;
}
#endif
