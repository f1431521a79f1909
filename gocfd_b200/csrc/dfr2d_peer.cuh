// Partition-to-partition exchange over peer memory (NVLink / NVSwitch P2P stores), no host in the loop and no collective
// library: the goroutines' shared memory of the reference (RungeKutta5SSP.Step fans out over partitions that read each
// other's Q_Face / edge store directly, model_problems/Euler2D/euler.go:408-418, edges.go:379-411) becomes
//
//   sender    pack kernel: gathers the message of every cut edge / shared vertex and STORES IT STRAIGHT INTO THE RECEIVER'S
//             MAILBOX (peer-mapped address: cudaDeviceEnablePeerAccess inside one process, cudaIpcOpenMemHandle between
//             processes); the last CTA to finish publishes the stage's sequence number into the receiver's arrival flag
//             (fence.sys + st.release.sys)
//   receiver  unpack kernel: every thread spins on the arrival flag of the partition its element of the mailbox comes from
//             (ld.acquire.sys), then scatters into the ghost columns / vertex maxima
//   max       the pair {max wave speed, max viscous wave speed} of calculateGlobalDT (euler.go:945-971) is put into every
//             peer's inbox the same way; the gather kernel waits for all inboxes, so it is also the barrier that makes the
//             mailboxes safe to overwrite in the next stage
//
// Sequence numbers are the host's stage counter + 1: every partition runs the same stages in the same order, so both
// sides agree without talking.  A mailbox segment written for stage s is consumed before the consumer's wave put of stage
// s, and a producer only writes stage s+1 after its wave gather of stage s: no double buffering is needed.
#pragma once
#include "dfr2d_device.cuh"

namespace dfr2d {

constexpr int kMaxParts = 32;

// where the doubles [first[p], first[p+1]) of my send order go: straight into partition p's mailbox
struct PutTab {
    double *dst[kMaxParts];                  // peer p: start of my segment inside its receive buffer
    unsigned long long *flag[kMaxParts];     // peer p: its arrival flag for messages of this exchange from me
    long long first[kMaxParts + 1];
    unsigned int *done;                      // local: CTAs of the running pack kernel that have finished
    int nParts;
};

// my receive buffer: doubles [first[p], first[p+1]) come from partition p, which raises flag[p]
struct WaitTab {
    const unsigned long long *flag;          // [kMaxParts] local
    long long first[kMaxParts + 1];
    int *err;                                // DevScalars.nanFlag
    int nParts;
};

struct WaveTab {
    unsigned long long *inbox[kMaxParts];    // peer p: its inbox row for me, [2 slots][2 values]
    unsigned long long *flag[kMaxParts];     // peer p: its wave arrival flag for me
    const unsigned long long *myInbox;       // local [kMaxParts][2 slots][2 values]
    const unsigned long long *myFlag;        // local [kMaxParts]
    int *err;                                // DevScalars.nanFlag
    int nParts, me;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// A partner that never arrives (crashed process, mismatched stage sequence) must not hang the GPU: after kPeerTimeoutNs
// the waiter gives up and raises err (DevScalars.nanFlag = 2, reported by dfr2d_step_finish as a failed exchange); every
// later wait of this partition then returns at once, so a dead partner costs one time-out, not one per kernel.
constexpr unsigned long long kPeerTimeoutNs = 5ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void wait_flag(const unsigned long long *p, unsigned long long seq, int *err) {
    if (ld_acquire_sys(p) >= seq) return;
    if (err && *(volatile int *)err == 2) return;        // an earlier wait already gave up: do not stack time-outs
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(p) < seq) {
        __nanosleep(100);
        if (global_ns() - t0 > kPeerTimeoutNs) {
            if (err) *(volatile int *)err = 2;
            return;
        }
    }
}

__device__ __forceinline__ int part_of(const long long *first, int nParts, long long b) {
    int p = 0;
    while (p + 1 < nParts && b >= first[p + 1]) p++;
    return p;
}

// destination of element b of the send order: the local send buffer (host-moved exchange) or the peer's mailbox
__device__ __forceinline__ void put_store(const PutTab *put, double *localBuf, long long b, double v) {
    if (put == nullptr) {
        localBuf[b] = v;
    } else {
        const int p = part_of(put->first, put->nParts, b);
        put->dst[p][b - put->first[p]] = v;
    }
}

// last CTA out publishes the sequence number to every peer that received something
__device__ __forceinline__ void put_signal(const PutTab *put, unsigned long long seq) {
    if (put == nullptr) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(put->done, 1u);
        if (prev == gridDim.x - 1) {
            *put->done = 0u;
            __threadfence_system();
            for (int p = 0; p < put->nParts; p++)
                if (put->first[p + 1] > put->first[p]) st_release_sys(put->flag[p], seq);
        }
    }
}

__device__ __forceinline__ void wait_segment(const WaitTab *w, long long b, unsigned long long seq) {
    if (w == nullptr) return;
    wait_flag(w->flag + part_of(w->first, w->nParts, b), seq, w->err);
}

// Halo messages.  One message slot per cut edge, `per` doubles each:
//   [4 vars][NpEdge points] rows `rowBase + row0[c] + i` of a [4][planeRows][Kp] array (Q_Face; or DissX then DissY when
//   src2 is given), in the sender's own edge-point order, followed (Q_Face message of the dissipation path only,
//   nTail = 3) by the three vertex epsilon values of the sender's element.
struct HaloPackArgs {
    int total, per, npEdge, planeRows, rowBase, Kp, nTail;
    const double *src, *src2;
    const int *elem, *row0, *etov;
    const double *epsV;
    double *buf;
    const PutTab *put;
    unsigned long long seq;
};

__global__ void k_halo_pack(HaloPackArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.total) {
        const int body = 4 * a.npEdge;
        const int c = t / a.per;
        int r = t % a.per;
        const double *src = a.src;
        double v;
        if (a.src2 != nullptr && r >= body) { src = a.src2; r -= body; }
        if (r < body) {
            const int n = r / a.npEdge, i = r % a.npEdge;
            v = src[((size_t)n * a.planeRows + a.rowBase + a.row0[c] + i) * a.Kp + a.elem[c]];
        } else {
            v = a.epsV[a.etov[(size_t)(r - body) * a.Kp + a.elem[c]]];
        }
        put_store(a.put, a.buf, (long long)t, v);
    }
    put_signal(a.put, a.seq);
}

// ghost element g = col - K gets the private vertex slots NV + 3g + {0,1,2} (its etov entries point there)
struct HaloUnpackArgs {
    int total, per, npEdge, planeRows, rowBase, Kp, nTail, K, NV;
    double *dst, *dst2;
    const int *col, *row0;
    double *epsV;
    const double *buf;
    const WaitTab *wait;
    unsigned long long seq;
};

__global__ void k_halo_unpack(HaloUnpackArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.total) return;
    wait_segment(a.wait, (long long)t, a.seq);
    const int body = 4 * a.npEdge;
    const int c = t / a.per;
    int r = t % a.per;
    const double v = __ldcg(a.buf + t);            // written by the partner over NVLink: read at L2
    double *dst = a.dst;
    if (a.dst2 != nullptr && r >= body) { dst = a.dst2; r -= body; }
    if (r < body) {
        const int n = r / a.npEdge, i = r % a.npEdge;
        dst[((size_t)n * a.planeRows + a.rowBase + a.row0[c] + i) * a.Kp + a.col[c]] = v;
    } else {
        a.epsV[(size_t)a.NV + 3 * (size_t)(a.col[c] - a.K) + (r - body)] = v;
    }
}

// Viscous exchange: for every cut edge the 4 x NpEdge owner-normal components of Epsilon (.) Grad of MY side
// (GradArgs::vn, already in the owner's point order) -> the partner's copy of the other side.
__global__ void k_vn_pack(int nCut, int npEdge, int NEp, const double *vn, const int *slot, const int *side, double *buf,
                          const PutTab *put, unsigned long long seq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 4 * npEdge;
    if (t < nCut * per) {
        const int c = t / per, r = t % per;              // r = n * NpEdge + i
        put_store(put, buf, (long long)t, vn[((size_t)side[c] * per + r) * NEp + slot[c]]);
    }
    put_signal(put, seq);
}
__global__ void k_vn_unpack(int nCut, int npEdge, int NEp, double *vn, const int *slot, const int *side, const double *buf,
                            const WaitTab *wait, unsigned long long seq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 4 * npEdge;
    if (t >= nCut * per) return;
    wait_segment(wait, (long long)t, seq);
    const int c = t / per, r = t % per;
    vn[((size_t)(1 - side[c]) * per + r) * NEp + slot[c]] = __ldcg(buf + t);
}

// shared-vertex exchange of the element -> vertex max merge: message = (sigma, eps) per listed vertex
__global__ void k_vertex_pack(int n, const int *vid, const double *sigmaV, const double *epsV, double *buf, const PutTab *put,
                              unsigned long long seq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        put_store(put, buf, 2 * (long long)t, sigmaV[vid[t]]);
        put_store(put, buf, 2 * (long long)t + 1, epsV[vid[t]]);
    }
    put_signal(put, seq);
}
__global__ void k_vertex_unpack_max(int n, const int *vid, unsigned long long *sigmaV, unsigned long long *epsV,
                                    const double *buf, const WaitTab *wait, unsigned long long seq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    wait_segment(wait, 2 * (long long)t, seq);
    const double s = __ldcg(buf + 2 * (size_t)t), e = __ldcg(buf + 2 * (size_t)t + 1);
    if (s > 0.0) atomic_max_nonneg(&sigmaV[vid[t]], s);
    if (e > 0.0) atomic_max_nonneg(&epsV[vid[t]], e);
}

// {max wave speed, max viscous wave speed} of this partition's edge phase -> every peer's inbox
__global__ void k_wave_put(const unsigned long long *mine /* sc->wave[slot] */, WaveTab w, int slot, unsigned long long seq) {
    const int p = threadIdx.x;
    if (p < w.nParts && p != w.me) {
        w.inbox[p][slot * 2 + 0] = mine[0];
        w.inbox[p][slot * 2 + 1] = mine[1];
        __threadfence_system();
        st_release_sys(w.flag[p], seq);
    }
}

// wait for every peer's pair, reduce: bit patterns of non-negative doubles order like the doubles
__global__ void k_wave_gather(unsigned long long *mine, WaveTab w, int slot, unsigned long long seq) {
    const int p = threadIdx.x;
    unsigned long long v0 = 0ull, v1 = 0ull;
    if (p < w.nParts) {
        if (p == w.me) {
            v0 = mine[0];
            v1 = mine[1];
        } else {
            wait_flag(w.myFlag + p, seq, w.err);
            v0 = __ldcg(w.myInbox + (p * 2 + slot) * 2 + 0);
            v1 = __ldcg(w.myInbox + (p * 2 + slot) * 2 + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u0 = __shfl_xor_sync(0xffffffffu, v0, o), u1 = __shfl_xor_sync(0xffffffffu, v1, o);
        v0 = u0 > v0 ? u0 : v0;
        v1 = u1 > v1 ? u1 : v1;
    }
    if (p == 0) {
        mine[0] = v0;
        mine[1] = v1;
    }
}

}  // namespace dfr2d
