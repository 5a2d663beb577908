// peer.cu -- the sharded path's one exchange as a kernel over NVLink peer memory.
//
// A training / evaluation step of the image-sharded path all-reduces three doubles, [sum smooth-L1, sum CE, sum
// positives] (SSD_trainer.py:105,108,600).  A library collective for 24 bytes costs a host-side call per step that is
// longer than the step's kernels (~93 us at B = 256) and cannot sit inside the step's CUDA graph on this stack.
// peer_allreduce_kernel does the exchange itself: every rank owns a 1 KB mailbox in its own HBM, mapped into the
// other ranks' address spaces through CUDA IPC (NVLink 5 / NVSwitch peer access).  One warp per rank:
//   1. lane r stores this rank's three sums into slot [generation][rank] of rank r's mailbox (peer stores), fences,
//      then stores the step's sequence number as the slot's tag;
//   2. lane r polls slot [generation][r] of the OWN mailbox until its tag is the sequence number, then reads the sums;
//   3. the world's sums are added in rank order (identical bits on every rank) and written back in place.
// The sequence number lives in device memory and is advanced by the kernel, so the launch is a plain graph node that
// can be replayed.  A rank that waits longer than ~10 s (a peer that never launched) raises bit 2 of dev_flags and
// writes NaN instead of hanging the GPU; later calls then do not wait at all.
// lag = 1 delivers the reduced sums of the PREVIOUS call instead (zeros on the first call): step 2 then collects slots
// that were posted a whole step earlier, so it practically never waits and the ranks are not re-synchronised every step
// (they may drift one step apart) -- the overlap a side-stream collective gives, without the host work.
// Four slot generations (seq mod 4): a rank can be at most one call ahead of the slowest reader of its slots (it needs
// every peer's post of call s to finish call s + lag), so the generation it overwrites was read at least two calls ago.
#include "common.cuh"

namespace ssdhot {

constexpr int kPeerMax = SSDHOT_PEER_MAX_RANKS;     // 8: one NVSwitch domain
constexpr int kGenerations = 4;
constexpr unsigned long long kPeerTimeoutNs = 10000000000ull;   // 10 s: ranks of one job are expected to arrive far closer together
constexpr int kMailboxDoubles = kGenerations * kPeerMax * 4;    // [generation][rank][sum loc, sum CE, sum pos, tag]

struct PeerSlots {
    double* box[kPeerMax];      // box[r] = rank r's mailbox as mapped in THIS process (box[rank] = the local one)
};

// img_part != null: the kernel also does finalize_sums_kernel's job first -- it folds the per-image partial sums [B][2] and
// the per-image positive counts of the loss kernel it was launched behind (PDL) into this rank's three sums, lane-strided
// and then over a fixed shuffle tree (deterministic), so that the step's loss branch has one dependent launch less.
__global__ void __launch_bounds__(32) peer_allreduce_kernel(double* __restrict__ sums, const PeerSlots ps, int rank, int world, int lag,
                                                            unsigned long long* __restrict__ seq_counter, int32_t* __restrict__ flags,
                                                            const double* __restrict__ img_part, const int32_t* __restrict__ n_pos, int B) {
    pdl_wait();                                     // (launched with PDL behind finalize_sums_kernel or the loss kernel itself)
    const int lane = threadIdx.x;
    const unsigned long long seq = *seq_counter + 1ull;
    const int par = (int)(seq % kGenerations);
    double a, c, n;
    if (img_part) {
        a = 0.0; c = 0.0; n = 0.0;
        for (int i = lane; i < B; i += 32) { a += img_part[2ll * i]; c += img_part[2ll * i + 1]; n += (double)n_pos[i]; }
        a = warp_sum(a); c = warp_sum(c); n = warp_sum(n);
    } else { a = sums[0]; c = sums[1]; n = sums[2]; }
    __syncwarp();
    if (lane < world) {
        double* q = ps.box[0];
#pragma unroll
        for (int i = 1; i < kPeerMax; ++i) if (lane == i) q = ps.box[i];
        volatile double* dst = q + (par * kPeerMax + rank) * 4;
        dst[0] = a; dst[1] = c; dst[2] = n;
        __threadfence_system();
        reinterpret_cast<volatile unsigned long long*>(dst)[3] = seq;
    }
    double ra = 0.0, rc = 0.0, rn = 0.0;
    const unsigned long long want = seq - (unsigned long long)lag;       // the call whose sums are delivered now
    const int wpar = (int)(want % kGenerations);
    if (lane < world && want > 0ull) {
        double* mine = ps.box[0];
#pragma unroll
        for (int i = 1; i < kPeerMax; ++i) if (rank == i) mine = ps.box[i];
        volatile double* src = mine + (wpar * kPeerMax + lane) * 4;
        volatile unsigned long long* tag = reinterpret_cast<volatile unsigned long long*>(src) + 3;
        const unsigned long long t0 = globaltimer_ns();
        bool ok = true;
        const bool gave_up_before = flags && (*reinterpret_cast<volatile int32_t*>(flags) & 2);    // then do not wait again
        while (*tag != want) {
            if (gave_up_before || globaltimer_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }
            __nanosleep(64);
        }
        __threadfence_system();
        if (ok) { ra = src[0]; rc = src[1]; rn = src[2]; }
        else {
            if (flags) atomicOr(flags, 2);
            ra = rc = rn = __longlong_as_double(0x7ff8000000000000ll);
        }
    }
    double A = 0.0, C = 0.0, N = 0.0;               // rank order: the same additions on every rank
    for (int r = 0; r < world; ++r) {
        A += __shfl_sync(FULL, ra, r);
        C += __shfl_sync(FULL, rc, r);
        N += __shfl_sync(FULL, rn, r);
    }
    if (lane == 0) { sums[0] = A; sums[1] = C; sums[2] = N; *seq_counter = seq; }
}

}  // namespace ssdhot

using namespace ssdhot;

extern "C" unsigned long long ssdhot_peer_mailbox_bytes(void) { return (unsigned long long)kMailboxDoubles * sizeof(double) + 64; }

// A mailbox (plus the sequence counter in its last 64 bytes) in this rank's HBM, zero-filled, from cudaMalloc so that it
// can be exported with CUDA IPC.
extern "C" int ssdhot_peer_alloc(void** mailbox_out) {
    if (!mailbox_out) return SSDHOT_ERR_NULL;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)ssdhot_peer_mailbox_bytes());
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, (size_t)ssdhot_peer_mailbox_bytes());
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    *mailbox_out = p;
    return SSDHOT_OK;
}
extern "C" int ssdhot_peer_free(void* mailbox) {
    if (!mailbox) return SSDHOT_OK;
    cudaError_t e = cudaFree(mailbox);
    return e == cudaSuccess ? SSDHOT_OK : (int)e;
}
// 64-byte CUDA IPC handle of a mailbox (HOST buffer), to be sent to the other ranks of the node.
extern "C" int ssdhot_peer_export(const void* mailbox, void* handle64_host) {
    if (!mailbox || !handle64_host) return SSDHOT_ERR_NULL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(mailbox));
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64_host, &h, 64);
    return SSDHOT_OK;
}
// Map another rank's mailbox (its exported handle) into this process; enables peer access over NVLink.
extern "C" int ssdhot_peer_open(const void* handle64_host, void** mailbox_out) {
    if (!handle64_host || !mailbox_out) return SSDHOT_ERR_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *mailbox_out = p;
    return SSDHOT_OK;
}
extern "C" int ssdhot_peer_close(void* mapped_mailbox) {
    if (!mapped_mailbox) return SSDHOT_OK;
    cudaError_t e = cudaIpcCloseMemHandle(mapped_mailbox);
    return e == cudaSuccess ? SSDHOT_OK : (int)e;
}

// All-reduce (sum) sums[3] in place across the `world` ranks whose mailboxes are mailboxes_host[0..world) (HOST array
// of DEVICE pointers as mapped in this process; entry `rank` is the local mailbox).  lag = 0: the sums of this call;
// lag = 1: the reduced sums of the previous call (zeros on the first).  Asynchronous on `stream`, capturable in a CUDA
// graph.  Every rank must call it the same number of times with the same lag.
extern "C" int ssdhot_allreduce_sums_peer(double* sums, void* const* mailboxes_host, int rank, int world, int lag, int32_t* dev_flags,
                                          ssdhot_stream_t stream) {
    if (!sums || !mailboxes_host) return SSDHOT_ERR_NULL;
    if (world < 1 || world > kPeerMax || rank < 0 || rank >= world) return SSDHOT_ERR_SHAPE;
    if (lag != 0 && lag != 1) return SSDHOT_ERR_VALUE;
    PeerSlots ps = {};
    for (int r = 0; r < world; ++r) {
        if (!mailboxes_host[r]) return SSDHOT_ERR_NULL;
        ps.box[r] = reinterpret_cast<double*>(mailboxes_host[r]);
    }
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(ps.box[rank] + kMailboxDoubles);
    cudaError_t e = launch_pdl(peer_allreduce_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, sums, ps, rank, world, lag, seq, dev_flags,
                               (const double*)nullptr, (const int32_t*)nullptr, 0);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    return SSDHOT_OK;
}

// The same exchange fed by the per-image partial sums a loss forward left in its workspace (call the forward with
// sums == NULL: it then skips its own final reduction): loss_work = that call's `work`, n_pos = that call's n_pos (NULL: the
// workspace copy).  sums receives the all-reduced sums.  One dependent launch less per step than forward + ssdhot_allreduce_sums_peer.
extern "C" int ssdhot_allreduce_partials_peer(const void* loss_work, int B, const int32_t* n_pos, double* sums, void* const* mailboxes_host,
                                              int rank, int world, int lag, int32_t* dev_flags, ssdhot_stream_t stream) {
    if (!loss_work || !sums || !mailboxes_host) return SSDHOT_ERR_NULL;
    if (B <= 0 || world < 1 || world > kPeerMax || rank < 0 || rank >= world) return SSDHOT_ERR_SHAPE;
    if (lag != 0 && lag != 1) return SSDHOT_ERR_VALUE;
    PeerSlots ps = {};
    for (int r = 0; r < world; ++r) {
        if (!mailboxes_host[r]) return SSDHOT_ERR_NULL;
        ps.box[r] = reinterpret_cast<double*>(mailboxes_host[r]);
    }
    // workspace layout of the loss forward (train_path.cu): img_part [B][2] double | n_pos [B] int32 | ...
    const double* img_part = reinterpret_cast<const double*>(loss_work);
    const int32_t* np = n_pos ? n_pos : reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(loss_work) + (size_t)B * 2 * sizeof(double));
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(ps.box[rank] + kMailboxDoubles);
    cudaError_t e = launch_pdl(peer_allreduce_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, sums, ps, rank, world, lag, seq, dev_flags,
                               img_part, np, B);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    return SSDHOT_OK;
}
