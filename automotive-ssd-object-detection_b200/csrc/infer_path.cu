// infer_path.cu -- decode, score threshold, ranking and greedy DIoU/CIoU NMS
// (SURVEY.md section 8a rows a6-a8).
//
// SSD300 with C == 6 (the reference's case) is ONE kernel, predict_image_kernel (second half of this file): the stream
// keeps one 16-bit key per row in shared memory -- or takes the keys the loss kernel's stream of the same logits left in
// the share buffer (ssdhot.h: ssdhot_share_bytes; train_path.cu) --, a histogram cut picks the few hundred "hot" rows, one
// thread per hot row evaluates it exactly, the pulled keys are ranked by counting (rank_keys_by_counting) and the shared
// back end (nms_round_backend) decodes, tests the pairs, resolves and emits.  Images that do not fit that mould are redone
// by the same CTA on the generic path below.
//
// The generic predict (other class counts, candidate ids >= 65536, ssdhot_predict_stages) = two kernels:
//  * score_kernel -- the HBM-bound stage.  The class logits are streamed once (two CTAs x 8 warps per image,
//    16-byte loads, software-pipelined); every (prior, foreground class) pair is tested against the score
//    threshold and the survivors are appended to the warp's own segment of the image's candidate list as
//    64-bit keys (order-preserving score bits << 32 | ~candidate id), so sorting the keys descending orders
//    by score and breaks ties by candidate id prior*(C-1)+class.  For C == 6 the scores are approximate
//    (ex2.approx / rcp.approx); decisions within 1e-4 of the threshold are re-made with the exact eager-CUDA
//    arithmetic, so the candidate set is exact.
//  * nms_image_kernel -- one CTA of 512 threads per image walks the image's candidates of ALL classes in
//    one global score order and applies class-aware greedy NMS, so the walk stops as soon as max_per_img
//    boxes survive.  That is exactly the reference's result -- per-class greedy NMS, then a global score
//    sort and `keep[:max_per_img]` (SSD_from_scratch.py:439-465): a candidate's fate depends only on
//    higher-scored candidates of its own class, all of which precede it in the global order, and survivors
//    appear in global score order -- but the thousands of low-score candidates that cannot reach the output
//    are never ranked or tested.  Rounds of <= 512 candidates: score histogram -> cut -> gather -> exact
//    scores of the pulled candidates -> bitonic sort -> decode -> class-local order -> pair tests (IoU gate,
//    then the exact metric for the few pairs that pass) into class-local bit rows -> per-class resolve.
// Both kernels are templated on the source of the head outputs (heads.cuh): the packed [B,8732,D] tensors, or the six
// per-level tensors of each branch as channels_last rows or NCHW planes (ssdhot_predict_heads) -- same candidate ids,
// same lists, bit-identical results.
// The stand-alone NMS entry point (mySSD.iou_nms, nms_sets_kernel / nms_unit) handles arbitrary set sizes
// and unbounded survivor counts with tiles of 64 and a CTA-local radix select.
#include <cstdlib>

#include "boxmath.cuh"
#include "heads.cuh"

namespace ssdhot {

constexpr int CHUNK = 256;     // candidates ranked per round
constexpr int TILE = 64;
constexpr float kFilterSlack = 0.999f;   // see suppresses()
constexpr int HBINS = 4096;    // score histogram: 16 octaves below 1.0 x 256 mantissa steps
constexpr int HBASE = (127 - 16) << 8;

struct UnitShared {
    unsigned hist[256];
    unsigned long long rowmask[TILE];   // bit j of rowmask[i]: tile member i suppresses tile member j (j > i)
    unsigned long long best[32];
    unsigned long long keepbits;
    unsigned char suppf[TILE];          // tile member suppressed by an earlier survivor
    unsigned sel_prefix, sel_need, sel_eq;
    int counter;                        // gather cursor
    int cut_bin, cut_count;
    int iscratch[32];
};

// ---- the pair predicate --------------------------------------------------------------------------
// "S suppresses c" iff NOT(metric(S, c) <= thr)  (SSD_from_scratch.py:690; NaN suppresses).
// Every metric here is <= IoU (DIoU and CIoU subtract non-negative penalties), so a pair whose
// intersection is below thr_lo * union, thr_lo = 0.999 * thr, has IoU < thr by a margin of 1e-3
// -- four orders of magnitude above fp32 rounding -- and survives without the two IEEE divisions.
// Pairs that pass (or whose union is not positive, where the exact test may yield NaN) take the
// exact path, so the decision is bit-identical to evaluating the metric everywhere.
// The cheap half of the predicate: false = the pair certainly does not suppress (IoU < thr by a margin).
__device__ __forceinline__ bool iou_gate(const BoxC& S, const BoxC& c, float thr_lo, float& inter, float& uni) {
    const float w = fmaxf(fsub(fminf(S.x2, c.x2), fmaxf(S.x1, c.x1)), 0.0f);
    const float h = fmaxf(fsub(fminf(S.y2, c.y2), fmaxf(S.y1, c.y1)), 0.0f);
    inter = fmul(w, h);
    uni = fsub(fadd(S.area, c.area), inter);
    return !(inter < fmul(thr_lo, uni));
}
// The exact half (two IEEE divisions), for pairs that passed the gate.
template <int METRIC>
__device__ __forceinline__ bool suppresses_exact(const BoxC& S, const BoxC& c, float inter, float uni, float thr) {
    const float iou = fdiv(inter, uni);
    float m = iou;
    if (METRIC != SSDHOT_METRIC_IOU) {
        m = pair_diou_from_iou(S, c, iou);
        if (METRIC == SSDHOT_METRIC_CIOU) {
            const float da = fsub(S.at, c.at);
            const float v = fmul(kFourOverPiSq, fmul(da, da));
            const float alpha = fdiv(v, fadd(fadd(fsub(1.0f, iou), v), kEps));
            m = fsub(m, fmul(alpha, v));
        }
    }
    return !(m <= thr);
}
// the same with the metric as a (CTA-uniform) runtime value: the per-image kernels carry it in their parameters, which keeps
// their instantiations to the head sources
__device__ __forceinline__ bool suppresses_exact_rt(int metric, const BoxC& S, const BoxC& c, float inter, float uni, float thr) {
    const float iou = fdiv(inter, uni);
    float m = iou;
    if (metric != SSDHOT_METRIC_IOU) {
        m = pair_diou_from_iou(S, c, iou);
        if (metric == SSDHOT_METRIC_CIOU) {
            const float da = fsub(S.at, c.at);
            const float v = fmul(kFourOverPiSq, fmul(da, da));
            const float alpha = fdiv(v, fadd(fadd(fsub(1.0f, iou), v), kEps));
            m = fsub(m, fmul(alpha, v));
        }
    }
    return !(m <= thr);
}
__device__ __forceinline__ bool suppresses_rt(int metric, const BoxC& S, const BoxC& c, float thr, float thr_lo) {
    float inter, uni;
    if (!iou_gate(S, c, thr_lo, inter, uni)) return false;
    return suppresses_exact_rt(metric, S, c, inter, uni, thr);
}
template <int METRIC>
__device__ __forceinline__ bool suppresses(const BoxC& S, const BoxC& c, float thr, float thr_lo) {
    float inter, uni;
    if (!iou_gate(S, c, thr_lo, inter, uni)) return false;
    return suppresses_exact<METRIC>(S, c, inter, uni, thr);
}

// ---- candidate sources ---------------------------------------------------------------------------
// A source is the set of candidates of one unit: entry i carries a 32-bit order-preserving key
// (0 = absent / consumed) and a candidate id.  for_each visits every entry once, spread over the CTA.
struct DenseSource {           // key array indexed by candidate id (stand-alone NMS; shared memory)
    unsigned* keys; int n;
    template <int NT, typename F>
    __device__ __forceinline__ void for_each(F f) const { for (int i = threadIdx.x; i < n; i += NT) f(i, (unsigned long long)keys[i]); }
    __device__ __forceinline__ void consume(int i) const { keys[i] = 0u; }
    __device__ __forceinline__ void replace(int i, unsigned long long r) const { keys[i] = (unsigned)r; }
    static __device__ __forceinline__ unsigned key(unsigned long long r) { return (unsigned)r; }
    static __device__ __forceinline__ unsigned long long sortkey(unsigned long long r, int i) {
        return (r << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    }
};
constexpr int SEGS = 16;       // list segments per image = warps of score_kernel per image = warps of nms_image_kernel
struct SegSource {             // 32 per-warp segments of (key << 32 | ~id) written by score_kernel (global memory)
    unsigned long long* base; const int* counts; int seg_cap;
    template <int NT, typename F>
    __device__ __forceinline__ void for_each(F f) const {
        static_assert((SEGS * 32) % NT == 0, "whole segments per warp");
        const int lane = threadIdx.x & 31;
        for (int seg = threadIdx.x >> 5; seg < SEGS; seg += NT / 32) {
            const int c = counts[seg];
            for (int j = lane; j < c; j += 128) {             // four independent loads in flight per lane
                unsigned long long r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) r[k] = j + 32 * k < c ? base[seg * seg_cap + j + 32 * k] : 0ull;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (j + 32 * k < c) f(seg * seg_cap + j + 32 * k, r[k]);
            }
        }
    }
    __device__ __forceinline__ void consume(int i) const { base[i] = 0ull; }
    __device__ __forceinline__ void replace(int i, unsigned long long r) const { base[i] = r; }
    static __device__ __forceinline__ unsigned key(unsigned long long r) { return (unsigned)(r >> 32); }
    static __device__ __forceinline__ unsigned long long sortkey(unsigned long long r, int) { return r; }
};

// ---- ranking helpers ---------------------------------------------------------------------------

// histogram bin of an ord_encode()d score in (0, 1]: exponent and the top 8 mantissa bits,
// clamped to 16 octaves below 1.0 (everything smaller shares bin 0)
__device__ __forceinline__ int score_bin(unsigned key) {
    const int v = (int)((key & 0x7fffffffu) >> 15) - HBASE;
    return v < 0 ? 0 : (v >= HBINS ? HBINS - 1 : v);
}
__device__ __forceinline__ unsigned bin_floor_key(int bin) {      // smallest key that falls in `bin`
    return bin <= 0 ? 1u : ((unsigned)(bin + HBASE) << 15) | 0x80000000u;
}
__device__ __forceinline__ void hist_add(unsigned* hist16, unsigned key) {
    const int bin = score_bin(key);
    atomicAdd(&hist16[bin >> 1], 1u << ((bin & 1) * 16));         // two 16-bit counters per word
}

// Find the lowest bin whose suffix count (candidates in this bin and above) is still <= cap:
// us.cut_bin / us.cut_count (cut_count == 0 if even the top non-empty bin exceeds cap).  The bins
// from the cut upwards are zeroed: they are consumed by the gather that follows.  Thread 0 owns the
// top HBINS / NT bins, thread 1 the next ones, ...
template <int NT>
__device__ __forceinline__ void hist_cut(unsigned* hist16, int cap, UnitShared& us) {
    constexpr int WPT = (HBINS / 2) / NT;                   // words (two bins each) per thread
    static_assert(WPT >= 1 && WPT * NT * 2 == HBINS, "thread count must divide the histogram");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w0 = (NT - 1 - tid) * WPT;
    unsigned wv[WPT];
    int c[2 * WPT];                                         // top bin first
    int mine = 0;
#pragma unroll
    for (int j = 0; j < WPT; ++j) {
        wv[j] = hist16[w0 + WPT - 1 - j];
        c[2 * j] = (int)(wv[j] >> 16);
        c[2 * j + 1] = (int)(wv[j] & 0xffffu);
        mine += c[2 * j] + c[2 * j + 1];
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    if (tid == 0) { us.cut_bin = -1; us.cut_count = 0; }
    if (lane == 31) us.iscratch[warp] = incl;
    __syncthreads();
    int run = incl - mine;                                  // candidates in bins above mine
    for (int w = 0; w < warp; ++w) run += us.iscratch[w];
    // cumulative counts grow monotonically going down, so "my lowest bin that still fits" is a valid
    // proposal and the cut is the minimum proposal over all threads
    const int top_bin = (w0 + WPT) * 2 - 1;
    int proposal = -1, prop_count = 0;
#pragma unroll
    for (int j = 0; j < 2 * WPT; ++j) {
        run += c[j];
        if (run <= cap) { proposal = top_bin - j; prop_count = run; }
    }
    unsigned long long packed = proposal >= 0 ? (((unsigned long long)(unsigned)proposal << 32) | (unsigned)prop_count) : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(FULL, packed, o);
        packed = y < packed ? y : packed;
    }
    if (lane == 0) us.best[warp] = packed;
    __syncthreads();
    if (tid < 32) {
        unsigned long long v = tid < NT / 32 ? us.best[tid] : ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(FULL, v, o);
            v = y < v ? y : v;
        }
        if (tid == 0 && v != ~0ull) { us.cut_bin = (int)(v >> 32); us.cut_count = (int)(v & 0xffffffffu); }
    }
    __syncthreads();
    const int cut = us.cut_bin;
    if (cut >= 0) {                                         // consume the bins from the cut upwards
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
            const int wi = w0 + WPT - 1 - j;
            unsigned v = wv[j];
            if (2 * wi + 1 >= cut) v &= 0x0000ffffu;
            if (2 * wi >= cut) v &= 0xffff0000u;
            hist16[wi] = v;
        }
    }
}

// Radix select of the K-th largest of the non-zero 32-bit keys keyfn(i, raw) over the entries of src.
// Returns the key, how many entries equal to it are needed (`need`) and how many exist (`eq`).
template <int NT, typename Src, typename KeyFn>
__device__ __forceinline__ void select_kth(const Src& src, KeyFn keyfn, unsigned K, UnitShared& us,
                                           unsigned& thr, unsigned& need, unsigned& eq) {
    const int tid = threadIdx.x;
    unsigned prefix = 0u, remaining = K, count_eq = 0u;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < 256; i += NT) us.hist[i] = 0u;
        __syncthreads();
        src.template for_each<NT>([&](int i, unsigned long long r) {
            const unsigned k = keyfn(i, r);
            if (k != 0u && (k & himask) == prefix) atomicAdd(&us.hist[(k >> shift) & 255u], 1u);
        });
        __syncthreads();
        if (tid < 32) {
            unsigned mine = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) mine += us.hist[255 - (tid * 8 + j)];
            unsigned incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(FULL, incl, o);
                if (tid >= o) incl += y;
            }
            const unsigned excl = incl - mine;
            if (excl < remaining && remaining <= incl) {
                unsigned run = excl;
                for (int j = 0; j < 8; ++j) {
                    const int bin = 255 - (tid * 8 + j);
                    const unsigned c = us.hist[bin];
                    if (run + c >= remaining) { us.sel_prefix = (unsigned)bin; us.sel_need = remaining - run; us.sel_eq = c; break; }
                    run += c;
                }
            }
        }
        __syncthreads();
        prefix |= us.sel_prefix << shift;
        remaining = us.sel_need;
        count_eq = us.sel_eq;
        __syncthreads();
    }
    thr = prefix; need = remaining; eq = count_eq;
}

// Bitonic sort (descending) of N 64-bit keys held in shared memory, by the first N threads (one key
// each): compare-exchange distances < 32 run on registers through warp shuffles, the larger ones
// through shared memory.  Every thread of the CTA must call it.
template <int N>
__device__ __forceinline__ void bitonic_desc(unsigned long long* keys) {
    const int tid = threadIdx.x;
    unsigned long long v = tid < N ? keys[tid] : 0ull;
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool desc = (tid & k) == 0;
            const bool lower = (tid & j) == 0;             // this thread holds the lower-indexed element
            const bool take_max = (lower == desc);         // descending run: lower index keeps the larger key
            if (j >= 32) {
                __syncthreads();
                if (tid < N) keys[tid] = v;
                __syncthreads();
                if (tid < N) {
                    const unsigned long long o = keys[tid ^ j];
                    v = take_max ? (o > v ? o : v) : (o < v ? o : v);
                }
            } else if (tid < N) {
                const unsigned long long o = __shfl_xor_sync(FULL, v, j);
                v = take_max ? (o > v ? o : v) : (o < v ? o : v);
            }
        }
    }
    __syncthreads();
    if (tid < N) keys[tid] = v;
    __syncthreads();
}

// ---- the unit ------------------------------------------------------------------------------------
// src:     the candidates (consumed as they are ranked)
// hist16:  packed 16-bit score histogram of the candidates (HIST only)
// Fetch:   BoxC operator()(unsigned id)       -- pixel box + constants of candidate id
// Group:   int operator()(unsigned id)        -- suppression group (class) of candidate id; only
//                                                members of the same group suppress each other
// Emit:    void operator()(int pos, unsigned long long key, unsigned id, const BoxC&)
// kept:    the surviving boxes in output order, capacity max_keep (shared or global memory)
// kidx:    [n_groups][max_keep] per-group lists of positions in `kept`; ngroup / cmask [n_groups]
// returns the number of survivors (valid in every thread)
struct UnitBuffers {
    unsigned* hist16; unsigned long long* ckey; BoxC* cbox; unsigned char* cgroup;
    BoxC* kept; unsigned short* kidx; int* ngroup; unsigned long long* cmask; int* cidx;
};

template <int METRIC, int NT, bool GROUPS, bool HIST, bool APPROX, typename Src, typename Fetch, typename Group, typename Emit, typename ExactKey>
__device__ int nms_unit(const Src src, const UnitBuffers buf, int n_cand, int n_groups, int max_keep, float thr,
                        UnitShared& us, Fetch fetch, Group group_of, Emit emit, ExactKey exact_key) {
    static_assert(NT >= CHUNK, "one thread per chunk entry in the sort");
    constexpr int SUB = NT / TILE;                 // threads cooperating on one tile member (16 or 8)
    constexpr int KEY_MARGIN = 512;                // ulps: >= 3e-5 relative, 3x the worst error of an approximate score
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* ckey = buf.ckey;
    BoxC* cbox = buf.cbox;
    BoxC* kept = buf.kept;
    const float thr_lo = fmul(thr, kFilterSlack);
    int kept_n = 0;
    int remaining = n_cand;
    bool use_hist = HIST && NT == 1024 && n_cand < 65536;  // 16-bit bin counters, one thread per four bins
    // APPROX: the keys of the source are approximate scores (score_kernel) until make_exact() has run;
    // exact scores are computed only for the candidates a chunk actually pulls.
    bool keys_exact = !APPROX;
    auto make_exact = [&]() {                      // every thread of the CTA
        src.template for_each<NT>([&](int i, unsigned long long r) {
            if (Src::key(r) == 0u) return;
            const unsigned long long sk = Src::sortkey(r, i);
            const unsigned id = 0xffffffffu - (unsigned)(sk & 0xffffffffull);
            src.replace(i, ((unsigned long long)exact_key(id) << 32) | (sk & 0xffffffffull));
        });
        if (HIST) for (int i = tid; i < HBINS / 2; i += NT) buf.hist16[i] = 0u;
        __syncthreads();
        if (use_hist) src.template for_each<NT>([&](int, unsigned long long r) { if (Src::key(r) != 0u) hist_add(buf.hist16, Src::key(r)); });
        __syncthreads();
        keys_exact = true;
    };
    if (GROUPS) for (int g = tid; g < n_groups; g += NT) buf.ngroup[g] = 0;
    while (remaining > 0 && kept_n < max_keep) {
        // ---- pull the next best candidates ---------------------------------------------------
        int K = remaining < CHUNK ? remaining : CHUNK;
        unsigned tkey = 1u, need = 0u, eq = 0u;
        bool all = remaining <= CHUNK;
        if (!all && use_hist) {
            hist_cut<NT == 1024 ? 1024 : 512>(buf.hist16, keys_exact ? CHUNK : CHUNK - 16, us);     // (room for the few margin strays)
            if (us.cut_count > 0) { K = us.cut_count; tkey = bin_floor_key(us.cut_bin); all = true; }
            else use_hist = false;                 // a single bin holds more than a chunk: exact select from here on
            __syncthreads();
        }
        unsigned tie_floor = 0u;        // among entries with key == tkey only those with ~id >= tie_floor are taken
        if (!all) {
            if (APPROX && !keys_exact) make_exact();       // the exact selection needs exact keys (use_hist stays off)
            select_kth<NT>(src, [&](int, unsigned long long r) { return Src::key(r); }, (unsigned)K, us, tkey, need, eq);
            if (need != eq) {
                // more entries share the threshold score than are needed: take the lowest candidate ids
                // (= the largest ~id, which is the low half of the sort key) -- a second select over ids
                unsigned n2, e2;
                select_kth<NT>(src, [&](int i, unsigned long long r) {
                    return Src::key(r) == tkey ? (unsigned)(Src::sortkey(r, i) & 0xffffffffull) : 0u;
                }, need, us, tie_floor, n2, e2);
            }
        }
        const unsigned gkey = (keys_exact || tkey <= (unsigned)KEY_MARGIN) ? tkey : tkey - (unsigned)KEY_MARGIN;
        if (tid == 0) us.counter = 0;
        __syncthreads();
        src.template for_each<NT>([&](int i, unsigned long long r) {
            const unsigned k = Src::key(r);
            if (k == 0u || k < gkey) return;
            const unsigned long long sk = Src::sortkey(r, i);
            if (keys_exact && k == tkey && (unsigned)(sk & 0xffffffffull) < tie_floor) return;
            const int pos = atomicAdd(&us.counter, 1);
            if (pos < CHUNK) {
                ckey[pos] = sk;
                if (APPROX) buf.cidx[pos] = i;
            }
            if (keys_exact) src.consume(i);
        });
        __syncthreads();
        if (APPROX && !keys_exact) {
            const int gathered = us.counter;
            if (gathered > CHUNK) {                // more margin strays than the slack: settle it with exact keys
                __syncthreads();
                make_exact();
                continue;
            }
            // exact scores of the pulled candidates; one that falls below the cut stays in the pool
            bool valid = false;
            if (tid < CHUNK) {
                unsigned long long out = 0ull;
                if (tid < gathered) {
                    const unsigned long long sk = ckey[tid];
                    const unsigned ek = exact_key(0xffffffffu - (unsigned)(sk & 0xffffffffull));
                    valid = ek >= tkey;
                    if (valid) { out = ((unsigned long long)ek << 32) | (sk & 0xffffffffull); src.consume(buf.cidx[tid]); }
                }
                ckey[tid] = out;
            }
            K = __syncthreads_count(valid);
        } else {
            for (int i = K + tid; i < CHUNK; i += NT) ckey[i] = 0ull;
            __syncthreads();
        }
        bitonic_desc<CHUNK>(ckey);
        for (int i = tid; i < K; i += NT) {
            const unsigned id = 0xffffffffu - (unsigned)(ckey[i] & 0xffffffffull);
            cbox[i] = fetch(id);
            buf.cgroup[i] = GROUPS ? (unsigned char)group_of(id) : (unsigned char)0;
        }
        if (tid < TILE) { us.rowmask[tid] = 0ull; us.suppf[tid] = 0; }
        __syncthreads();

        // ---- greedy NMS over the sorted chunk, tile by tile ----------------------------------
        // SUB threads share one tile member j: they split (a) the survivors of j's group and (b) the
        // earlier tile members; (a) sets suppf[j], (b) sets bit j of rowmask[i].
        for (int t0 = 0; t0 < K && kept_n < max_keep; t0 += TILE) {
            const int m = (K - t0) < TILE ? (K - t0) : TILE;
            const int j = tid / SUB, sub = tid % SUB;
            if (j < m) {
                const BoxC c = cbox[t0 + j];
                const int g = (int)buf.cgroup[t0 + j];
                bool hit = false;
                if (GROUPS) {
                    const unsigned short* list = buf.kidx + (size_t)g * max_keep;
                    const int ng = buf.ngroup[g];
                    for (int i = sub; i < ng; i += SUB) hit |= suppresses<METRIC>(kept[list[i]], c, thr, thr_lo);
                } else {
                    for (int i = sub; i < kept_n; i += SUB) hit |= suppresses<METRIC>(kept[i], c, thr, thr_lo);
                }
                for (int i = sub; i < j; i += SUB) {
                    if ((!GROUPS || (int)buf.cgroup[t0 + i] == g) && suppresses<METRIC>(cbox[t0 + i], c, thr, thr_lo))
                        atomicOr(&us.rowmask[i], 1ull << j);
                }
                if (hit) us.suppf[j] = 1;            // benign race: every writer stores 1
            }
            __syncthreads();
            if (warp == 0) {
                const bool in_lo = lane < m, in_hi = lane + 32 < m;
                const unsigned s_lo = __ballot_sync(FULL, in_lo && us.suppf[lane] != 0);
                const unsigned s_hi = __ballot_sync(FULL, in_hi && us.suppf[lane + 32] != 0);
                const unsigned z_lo = __ballot_sync(FULL, in_lo && us.rowmask[lane] != 0ull);
                const unsigned z_hi = __ballot_sync(FULL, in_hi && us.rowmask[lane + 32] != 0ull);
                if (lane == 0) {
                    const unsigned long long valid = m == 64 ? ~0ull : ((1ull << m) - 1ull);
                    const unsigned long long supp = ((unsigned long long)s_hi << 32) | s_lo;
                    const unsigned long long nz = ((unsigned long long)z_hi << 32) | z_lo;
                    unsigned long long alive = ~supp & valid;
                    // only members that suppress somebody need the serial walk, in index order
                    unsigned long long pending = alive & nz;
                    while (pending) {
                        const int i = __ffsll((long long)pending) - 1;
                        alive &= ~us.rowmask[i];
                        pending &= alive & ~((2ull << i) - 1ull);
                    }
                    // truncate to the room that is left (lowest indices = highest scores first)
                    int room = max_keep - kept_n;
                    if (__popcll(alive) > room) {
                        unsigned long long t = alive, keepb = 0ull;
                        while (room-- > 0) { const int i = __ffsll((long long)t) - 1; keepb |= 1ull << i; t &= t - 1ull; }
                        alive = keepb;
                    }
                    us.keepbits = alive;
                }
            }
            __syncthreads();
            const unsigned long long keepb = us.keepbits;
            if (GROUPS) {
                // tile members per group (one thread per group; at most 64 byte loads each)
                for (int g = tid; g < n_groups; g += NT) {
                    unsigned long long mk = 0ull;
                    for (int i = 0; i < m; ++i) mk |= (unsigned long long)((int)buf.cgroup[t0 + i] == g) << i;
                    buf.cmask[g] = mk;
                }
                __syncthreads();
            }
            if (tid < m && ((keepb >> tid) & 1ull)) {
                const unsigned long long lower = keepb & ((1ull << tid) - 1ull);
                const int pos = kept_n + __popcll(lower);
                const BoxC bx = cbox[t0 + tid];
                kept[pos] = bx;
                if (GROUPS) {
                    const int g = (int)buf.cgroup[t0 + tid];
                    buf.kidx[(size_t)g * max_keep + buf.ngroup[g] + __popcll(lower & buf.cmask[g])] = (unsigned short)pos;
                }
                const unsigned long long key = ckey[t0 + tid];
                emit(pos, key, 0xffffffffu - (unsigned)(key & 0xffffffffull), bx);
            }
            if (tid >= NT - TILE) { us.rowmask[tid - (NT - TILE)] = 0ull; us.suppf[tid - (NT - TILE)] = 0; }
            kept_n += __popcll(keepb);
            __syncthreads();
            if (GROUPS) {
                for (int g = tid; g < n_groups; g += NT) buf.ngroup[g] += __popcll(keepb & buf.cmask[g]);
                __syncthreads();
            }
        }
        remaining -= K;
    }
    return kept_n;
}

// ---- predict: parameters ---------------------------------------------------------------------------
struct PredictParams {
    const float* pri; int P; const float* loc_all; const float* conf_all; int B, C;
    float score_thresh, nms_thresh; int max_keep; float vc, vs, img_w, img_h;
    int metric, agnostic;           // SSDHOT_METRIC_*, class-agnostic NMS (the per-image kernels read them at run time)
    unsigned long long* cand;       // [B][P*(C-1)] candidate keys
    int* cand_count;                // [B]
    int64_t* out_labels; float* out_scores; float* out_boxes; int32_t* out_cand; int32_t* out_count;
    unsigned long long* timeline;   // debug (ssdhot_debug_timeline) or null
    HeadView loc_h, conf_h;         // head sources (SRC_LEVEL_ROWS / SRC_LEVEL_PLANES) instead of loc_all / conf_all
    // key hand-off from train_image_kernel of the same eval step (ssdhot_share_bytes; null = none)
    int* share_flag; const unsigned* share_keys;
};

// exp(x_i - max) of one row and their sum in eager torch-CUDA order (persistent warp softmax:
// lanes = min(next_pow2(C), 32); lane l accumulates elements l, l+32, ... in order; then butterfly
// adds over xor offsets lanes/2 .. 1).  C == 6 is the reference's class count.
__device__ __forceinline__ float row_exps6(const float* x, float* e) {       // x: the six logits (registers)
    const float mx = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(x[4], x[5]));
#pragma unroll
    for (int i = 0; i < 6; ++i) e[i] = expf(fsub(x[i], mx));
    return fadd(fadd(fadd(e[0], e[4]), e[2]), fadd(fadd(e[1], e[5]), e[3]));
}

__device__ __forceinline__ float row_sum_generic(const float* __restrict__ row, int C, float& mx) {
    int lanes = 1;
    while (lanes < C && lanes < 32) lanes <<= 1;
    mx = __ldg(row);
    for (int i = 1; i < C; ++i) mx = fmaxf(mx, __ldg(row + i));
    float part[32];
    for (int l = 0; l < 32; ++l) part[l] = 0.0f;
    for (int i = 0; i < C; ++i) part[i & (lanes - 1)] = fadd(part[i & (lanes - 1)], expf(fsub(__ldg(row + i), mx)));
    for (int off = lanes >> 1; off > 0; off >>= 1)
        for (int l = 0; l < off; ++l) part[l] = fadd(part[l], part[l + off]);
    return part[0];
}

constexpr int ST = 256;        // threads of score_kernel
constexpr int SCS = SEGS / (ST / 32);   // CTAs per image: every warp owns one list segment
constexpr int PT = 1024;       // threads of nms_image_kernel
constexpr int UT = 512;        // threads of stand-alone NMS

// rows of one list segment (even, so that a segment starts on a 48-byte row pair when C == 6)
__host__ __device__ inline int seg_rows(int P) { return (((P + SEGS - 1) / SEGS) + 1) & ~1; }
// rows a list segment has room for: the plane-wise order of the NCHW head source (heads.cuh) hands a warp up to
// 10 chunks of 32 row pairs
__host__ __device__ inline int seg_cap_rows(int P) { return P == 8732 ? 640 : seg_rows(P); }

__device__ __forceinline__ float ex2_approx_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exact softmax score of class k+1 of a 6-logit row, eager torch-CUDA order (SFS:388), as an order-preserving key
template <int SRC>
__device__ __forceinline__ unsigned exact_score_key6(const HeadReader<SRC, 6>& rd, int p, int k) {
    float x[6], e[6];
    rd.row(p, x);
    const float sum = row_exps6(x, e);
    const float ek = k == 0 ? e[1] : k == 1 ? e[2] : k == 2 ? e[3] : k == 3 ? e[4] : e[5];
    return __float_as_uint(fdiv(ek, sum)) | 0x80000000u;
}

// score_kernel: the HBM-bound half of predict.  Warp w of the image's SCS CTAs streams rows
// [w * seg_rows, (w+1) * seg_rows) and appends the (prior, class) pairs that pass the score threshold to
// its own segment of the candidate list -- no atomics, no shared memory, counts written at the end.
// C == 6: softmax is evaluated with ex2.approx / rcp.approx (relative error <= 1.1e-5, see nms_unit's
// KEY_MARGIN); the strict test `score > thresh` (SFS:402) is decided by the approximate score when it is
// more than 1e-4 (relative) away from the threshold and by the exact eager-CUDA arithmetic otherwise, so
// the candidate SET is exact; the keys are approximate and nms_image_kernel refines the ones it pulls.
// One warp's share of the stream: list segment `seg` of image b (htab / regions: the CTA's head tables, already filled).
template <int CT, int SRC>
__device__ __forceinline__ void score_segment(const PredictParams& prm, int b, int seg, int lane, const HeadTable& htab,
                                              const PlaneRegions& regions) {
    const int P = prm.P, n_fg = prm.C - 1;
    const int rows = seg_rows(P);
    const int r0 = min(P, seg * rows), r1 = min(P, r0 + rows);
    const float* conf_b = SRC == SRC_PACKED ? prm.conf_all + (long long)b * P * prm.C : nullptr;
    unsigned long long* list = prm.cand + ((long long)b * SEGS + seg) * seg_cap_rows(P) * n_fg;
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;                                            // warp-uniform
    if (CT == 6) {
        const float thr = prm.score_thresh, thr_hi = thr * 1.0001f, thr_lo = thr * 0.9999f;
        const HeadReader<SRC, 6> rd = {conf_b, &htab};
        const int q0 = r0 >> 1, q1 = r1 >> 1;               // P is even on this path
        // iteration `it` of this warp: 32 row pairs -- consecutive pairs of the segment's row range, or (NCHW planes) chunk
        // seg + 16 * it of the plane-wise order (heads.cuh).  -> this lane has a pair; p0 = its first row
        const int n_it = SRC == SRC_LEVEL_PLANES ? (kPlaneChunks - seg + SEGS - 1) / SEGS : (q1 - q0 + 31) >> 5;
        PlaneWalk walk(&regions);
        auto load_it = [&](int it, float* xx, int& p0) -> bool {
            if (SRC == SRC_LEVEL_PLANES) {
                return walk.load(seg + SEGS * it, lane, xx, p0);
            } else {
                const int q = q0 + 32 * it + lane;
                p0 = 2 * q;
                if (q < q1) rd.pair(q, xx);
                return q < q1;
            }
        };
        // software pipeline: the loads of the next row pair (three 16-byte loads when rows are contiguous) are in
        // flight while this one is scored
        float nx[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) nx[j] = 0.f;
        int n_p0 = 0;
        bool n_live = n_it > 0 && load_it(0, nx, n_p0);
        for (int it = 0; it < n_it; ++it) {
            unsigned pass = 0u;                             // bit (h * 5 + k): row p0+h, class k+1 is a candidate
            float sc[10];
            float x[2][6];
#pragma unroll
            for (int j = 0; j < 12; ++j) x[j / 6][j % 6] = nx[j];
            const bool live = n_live;
            const int p0 = n_p0;
            n_live = it + 1 < n_it && load_it(it + 1, nx, n_p0);
            if (live) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float mx = fmaxf(fmaxf(fmaxf(x[h][0], x[h][1]), fmaxf(x[h][2], x[h][3])), fmaxf(x[h][4], x[h][5]));
                    float e[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) e[i] = ex2_approx_ftz((x[h][i] - mx) * 1.4426950408889634f);
                    const float rs = rcp_approx_ftz(((e[0] + e[1]) + (e[2] + e[3])) + (e[4] + e[5]));
                    unsigned maybe = 0u;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const float s = e[k + 1] * rs;
                        sc[h * 5 + k] = s;
                        if (s > thr_hi) pass |= 1u << (h * 5 + k);
                        else if (s >= thr_lo) maybe |= 1u << k;
                    }
                    if (maybe) {                            // within 1e-4 of the threshold: the exact arithmetic decides
                        float ee[6];
                        const float mxe = fmaxf(fmaxf(fmaxf(x[h][0], x[h][1]), fmaxf(x[h][2], x[h][3])), fmaxf(x[h][4], x[h][5]));
#pragma unroll
                        for (int i = 0; i < 6; ++i) ee[i] = expf(fsub(x[h][i], mxe));
                        const float sum = fadd(fadd(fadd(ee[0], ee[4]), ee[2]), fadd(fadd(ee[1], ee[5]), ee[3]));
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (((maybe >> k) & 1u) && fdiv(ee[k + 1], sum) > thr) pass |= 1u << (h * 5 + k);
                    }
                }
            }
            // append: lane l's entries follow those of lanes < l
            const int mine = __popc(pass);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += y;
            }
            int at = cnt + incl - mine;
            cnt += __shfl_sync(FULL, incl, 31);
            if (pass) {
                const unsigned id0 = (unsigned)p0 * 5u;
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    if ((pass >> j) & 1u) {
                        list[at++] = ((unsigned long long)(__float_as_uint(sc[j]) | 0x80000000u) << 32) |
                                     (unsigned long long)(0xffffffffu - (id0 + (unsigned)j));
                    }
                }
            }
        }
    } else {
        for (int base = r0; base < r1; base += 32) {
            const int p = base + lane;
            const bool live = p < r1;
            float mx = 0.0f, sum = 1.0f;
            if (live) sum = row_sum_generic(conf_b + (long long)p * prm.C, prm.C, mx);
            for (int k = 0; k < n_fg; ++k) {
                bool ok = false;
                unsigned long long key = 0ull;
                if (live) {
                    const float s = fdiv(expf(fsub(__ldg(conf_b + (long long)p * prm.C + k + 1), mx)), sum);
                    ok = s > prm.score_thresh;
                    key = ((unsigned long long)(__float_as_uint(s) | 0x80000000u) << 32) |
                          (unsigned long long)(0xffffffffu - (unsigned)(p * n_fg + k));
                }
                const unsigned bal = __ballot_sync(FULL, ok);
                if (ok) list[cnt + __popc(bal & lt)] = key;
                cnt += __popc(bal);
            }
        }
    }
    if (lane == 0) prm.cand_count[b * SEGS + seg] = cnt;
}

template <int CT, int SRC>
__global__ void __launch_bounds__(ST, 4) score_kernel(const PredictParams prm) {
    pdl_trigger();                                 // nms_image_kernel's CTAs may be scheduled as SMs free up (they wait for this grid)
    const int b = blockIdx.x / SCS, part = blockIdx.x % SCS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ HeadTable htab;                     // per-level bases of this image (head sources only)
    __shared__ PlaneRegions regions;               // (NCHW heads)
    if (SRC != SRC_PACKED) {
        if (SRC == SRC_LEVEL_PLANES) plane_regions_fill(regions, prm.conf_h, b, threadIdx.x);
        else head_table_fill<SRC, 6>(htab, prm.conf_h, b, threadIdx.x);
        __syncthreads();
    }
    score_segment<CT, SRC>(prm, b, part * (ST / 32) + warp, lane, htab, regions);
}

// shared-memory carve-up of the stand-alone unit (nms_sets_kernel carves by hand)
// ---- predict: per-image ranking + class-aware greedy NMS -------------------------------------------
// One CTA of 512 threads per image (two or three CTAs per SM, so a batch of 256 is one wave).  Per round:
//   pull    the best <= 496 unconsumed candidates: score histogram -> cut -> gather -> exact scores of
//           the pulled ones (the list holds approximate scores when APPROX) -> bitonic sort;
//   decode  their boxes (SFS:419-425);
//   order   class-local positions: position m of entry j among the round's entries of its class;
//   test    entry j against the survivors of its class from earlier rounds and against the earlier
//           entries of its class in this round -- bit m of row i = "i suppresses the entry at position m";
//   resolve one warp per class walks its rows in score order (only rows that are alive and suppress
//           something), the classes in parallel;
//   emit    survivors in global score order until max_per_img boxes are out.
// The first round pulls ~1.5 max_per_img candidates, which is enough for typical inputs; further rounds
// run only while fewer than max_per_img boxes have survived and candidates remain.  This is exactly the
// reference's per-class NMS + global sort + truncation (SFS:439-465): a candidate's fate depends only on
// higher-scored candidates of its own class, all of which precede it in the global order.
constexpr int CH = 512;        // candidates per round
constexpr int MW = CH / 64;    // 64-bit words of a class-local bit row
constexpr int PAIRS_CAP = 2048; // listed gate survivors of a round (rest: settled inline)

struct ImgBuffers {
    unsigned long long* ckey;      // [CH] sort keys of the round
    unsigned long long* mat;       // [CH][MW] suppression rows (bits = class-local positions)
    unsigned long long* aliveW;    // [n_groups][MW] class-local alive bits
    unsigned long long* nzW;       // [n_groups][MW] class-local "row suppresses something" bits
    BoxC* cbox;                    // [CH]
    BoxC* kept;                    // [max_keep] survivors in output order
    unsigned* hist16;              // [HBINS / 2]
    int* cidx;                     // [CH] list index of a pulled candidate
    int* coff;                     // [n_groups + 1] offsets of the class lists inside clist
    int* ngroup;                   // [n_groups] survivors per class so far
    int* wcnt;                     // [16][n_groups] per-warp class counts of the round
    unsigned short* clist;         // [CH] round entries grouped by class, score order inside a class
    unsigned short* cpos;          // [CH] class-local position of entry j
    unsigned short* kidx;          // [n_groups][max_keep] positions in `kept` of each class's survivors
    unsigned char* cgroup;         // [CH]
    unsigned short* cmask;         // [CH] cells of a 4 x 4 grid over the image that the box touches (pair pre-filter)
    unsigned* plist;               // [PAIRS_CAP] pairs that passed the IoU gate: aq | m << 9 | class << 18
};
__host__ __device__ inline size_t img_smem_bytes(int max_keep, int n_groups) {
    const int ng = n_groups > 0 ? n_groups : 1;
    return (size_t)CH * 8 + (size_t)CH * MW * 8 + (size_t)ng * MW * 16 + (size_t)CH * sizeof(BoxC) + (size_t)max_keep * sizeof(BoxC) +
           (size_t)HBINS * 2 + (size_t)CH * 4 + (size_t)PAIRS_CAP * 4 + (size_t)(ng + 1) * 4 + (size_t)ng * 4 + (size_t)16 * ng * 4 +
           (size_t)CH * 2 * 3 + (size_t)ng * max_keep * 2 + (size_t)CH + 64;
}
__device__ __forceinline__ ImgBuffers carve_img(unsigned char* dyn, int max_keep, int n_groups) {
    const int ng = n_groups > 0 ? n_groups : 1;
    ImgBuffers b;
    b.ckey = reinterpret_cast<unsigned long long*>(dyn);
    b.mat = b.ckey + CH;
    b.aliveW = b.mat + (size_t)CH * MW;
    b.nzW = b.aliveW + (size_t)ng * MW;
    b.cbox = reinterpret_cast<BoxC*>(b.nzW + (size_t)ng * MW);
    b.kept = b.cbox + CH;
    b.hist16 = reinterpret_cast<unsigned*>(b.kept + max_keep);
    b.cidx = reinterpret_cast<int*>(b.hist16 + HBINS / 2);
    b.plist = reinterpret_cast<unsigned*>(b.cidx + CH);
    b.coff = reinterpret_cast<int*>(b.plist + PAIRS_CAP);
    b.ngroup = b.coff + ng + 1;
    b.wcnt = b.ngroup + ng;
    b.clist = reinterpret_cast<unsigned short*>(b.wcnt + 16 * ng);
    b.cpos = b.clist + CH;
    b.cmask = b.cpos + CH;
    b.kidx = b.cmask + CH;
    b.cgroup = reinterpret_cast<unsigned char*>(b.kidx + (size_t)ng * max_keep);
    return b;
}

constexpr int IT = 512;        // threads of nms_image_kernel

// ---- cold paths of nms_image_kernel, kept out of line so that the code every image runs stays compact -------
// Replace every approximate key of the list by the exact one and rebuild the score histogram (every thread of the CTA).
template <int SRC>
__device__ __noinline__ void nms_make_exact(const SegSource src, unsigned* hist16, const HeadReader<SRC, 6> conf_rd, bool use_hist) {
    src.for_each<IT>([&](int i, unsigned long long r) {
        if (SegSource::key(r) == 0u) return;
        const unsigned id = 0xffffffffu - (unsigned)(r & 0xffffffffull);
        src.replace(i, ((unsigned long long)exact_score_key6(conf_rd, (int)(id / 5u), (int)(id % 5u)) << 32) | (r & 0xffffffffull));
    });
    for (int i = threadIdx.x; i < HBINS / 2; i += IT) hist16[i] = 0u;
    __syncthreads();
    if (use_hist) src.for_each<IT>([&](int, unsigned long long r) { if (SegSource::key(r) != 0u) hist_add(hist16, SegSource::key(r)); });
    __syncthreads();
}
// Exact cut for the K best entries when the histogram cannot give one (a single bin holds more than a round).
__device__ __noinline__ void nms_select_cut(const SegSource src, UnitShared& us, int K, unsigned* tkey_out, unsigned* tie_floor_out) {
    unsigned tkey, need, eq, tie_floor = 0u;
    select_kth<IT>(src, [&](int, unsigned long long r) { return SegSource::key(r); }, (unsigned)K, us, tkey, need, eq);
    if (need != eq) {                               // ties at the cut: the lowest candidate ids (largest ~id) first
        unsigned n2, e2;
        select_kth<IT>(src, [&](int, unsigned long long r) { return SegSource::key(r) == tkey ? (unsigned)(r & 0xffffffffull) : 0u; },
                       need, us, tie_floor, n2, e2);
    }
    *tkey_out = tkey;
    *tie_floor_out = tie_floor;
}
#define SSDHOT_FAST_EXIT(code, info) do { if (prm.timeline && threadIdx.x == 0) prm.timeline[(long long)blockIdx.x * 16 + 15] = (unsigned long long)(code) | ((unsigned long long)(info) << 8); } while (0)
#define SSDHOT_NSTAMP(k) do { if (prm.timeline && threadIdx.x == 0 && first) prm.timeline[(long long)blockIdx.x * 16 + (k)] = globaltimer_ns(); } while (0)

// ---- one round's back end (every thread of the CTA) ------------------------------------------------------------------
// buf.ckey[0 .. K) holds the round's keys (exact score bits << 32 | ~candidate id), zero beyond K up to CH.  Sorts them,
// decodes the boxes, orders the entries class by class, runs the pair tests against earlier survivors and inside the round,
// resolves every class and emits the survivors in global score order behind the kept_n earlier ones.
// -> the number of entries of this round that survive (the caller clamps kept_n + that to max_keep).
template <int SRC>
__device__ __forceinline__ int nms_round_backend(const PredictParams& prm, const int b, const ImgBuffers& buf, UnitShared& us,
                                                 const HeadReader<SRC, 4>& loc_rd, const int K, const int kept_n, const bool first,
                                                 const bool sorted = false) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_fg = prm.C - 1, max_keep = prm.max_keep;
    const int METRIC = prm.metric;
    const bool AGN = prm.agnostic != 0;
    const int n_groups = AGN ? 1 : n_fg;
    const float thr = prm.nms_thresh, thr_lo = fmul(thr, kFilterSlack);
    const bool want_atan = METRIC == SSDHOT_METRIC_CIOU;
    const long long o = (long long)b * max_keep;
    const unsigned lt = (1u << lane) - 1u;
    if (sorted) {}                                       // (the caller ranked the keys: rank_keys_by_counting)
    else if (K <= 128) bitonic_desc<128>(buf.ckey);      // (entries beyond K are zero and stay behind)
    else if (K <= 256) bitonic_desc<256>(buf.ckey);
    else bitonic_desc<CH>(buf.ckey);

    SSDHOT_NSTAMP(5);
    if (prm.timeline && tid == 0 && first) prm.timeline[(long long)blockIdx.x * 16 + 13] = (unsigned long long)K;
    // ---- decode + class-local order --------------------------------------------------------------
    int my_g = -1;
    BoxC my_box = {};
    unsigned my_cells = 0u;
    float lv[4] = {0.f, 0.f, 0.f, 0.f};
    float4 my_prior = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < K) {
        // the box offsets and the prior are requested here and used behind the class ordering (three barriers on): their
        // latency passes under it
        const unsigned id = 0xffffffffu - (unsigned)(buf.ckey[tid] & 0xffffffffull);
        const unsigned p = id / (unsigned)n_fg;
        loc_rd.row((int)p, lv);
        my_prior = ldg4(prm.pri + 4ll * p);
        my_g = AGN ? 0 : (int)(id % (unsigned)n_fg);
        buf.cgroup[tid] = (unsigned char)my_g;
    }
    auto decode_mine = [&]() {
        const float4 box = decode_box(make_float4(lv[0], lv[1], lv[2], lv[3]), my_prior, prm.vc, prm.vs);
        const float4 px = to_pixel_xyxy(box, prm.img_w, prm.img_h);
        my_box = box_consts(px.x, px.y, px.z, px.w, want_atan);
        {   // which quarter-columns / quarter-rows of the image the (clamped) box reaches: boxes that share no cell
            // have an empty intersection, so the pair test can skip them on a 16-bit AND
            const float qx = 4.0f / prm.img_w, qy = 4.0f / prm.img_h;
            const int cx0 = min(3, max(0, (int)(px.x * qx))), cx1 = min(3, max(0, (int)(px.z * qx)));
            const int cy0 = min(3, max(0, (int)(px.y * qy))), cy1 = min(3, max(0, (int)(px.w * qy)));
            const unsigned xm = ((2u << cx1) - 1u) & ~((1u << cx0) - 1u);
            unsigned m16 = 0u;
            for (int r = cy0; r <= cy1; ++r) m16 |= xm << (4 * r);
            // an empty (or NaN) box has union 0 with another empty box: IoU = NaN, which suppresses (SFS:690) -- always test it
            if (!(fmul(fsub(px.z, px.x), fsub(px.w, px.y)) > 0.0f)) m16 = 0xffffu;
            my_cells = m16;
        }
    };
    for (int i = tid; i < 16 * n_groups; i += IT) buf.wcnt[i] = 0;
    for (int i = tid; i < n_groups * MW; i += IT) buf.nzW[i] = 0ull;
    {
        ulonglong2* m2 = reinterpret_cast<ulonglong2*>(buf.mat);
        for (int i = tid; i < K * MW / 2; i += IT) m2[i] = make_ulonglong2(0ull, 0ull);
    }
    __syncthreads();
    const unsigned peers = __match_any_sync(FULL, my_g);            // lanes of this warp with the same class
    const int local_rank = __popc(peers & lt);
    if (my_g >= 0 && local_rank == 0) buf.wcnt[warp * n_groups + my_g] = __popc(peers);
    __syncthreads();
    if (tid < n_groups) {                                           // class totals -> offsets (n_groups <= 255 < IT)
        int tot = 0;
        for (int w = 0; w < 16; ++w) tot += buf.wcnt[w * n_groups + tid];
        us.hist[tid] = (unsigned)tot;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int g = 0; g < n_groups; ++g) { buf.coff[g] = run; run += (int)us.hist[g]; }
        buf.coff[n_groups] = run;
    }
    int my_m = 0;
    if (my_g >= 0) {
        for (int w = 0; w < warp; ++w) my_m += buf.wcnt[w * n_groups + my_g];
        my_m += local_rank;
        buf.cpos[tid] = (unsigned short)my_m;
        decode_mine();
    }
    __syncthreads();
    if (my_g >= 0) {                                                // boxes and cell masks in class order
        const int at = buf.coff[my_g] + my_m;
        buf.cbox[at] = my_box;
        buf.cmask[at] = (unsigned short)my_cells;
    }
    for (int i = tid; i < n_groups * MW; i += IT) {                 // alive = every position the class has this round
        const int g = i / MW, w = i % MW, n_c = buf.coff[g + 1] - buf.coff[g];
        const int bits = min(max(n_c - 64 * w, 0), 64);
        buf.aliveW[i] = bits == 64 ? ~0ull : ((1ull << bits) - 1ull);
    }
    __syncthreads();

    SSDHOT_NSTAMP(6);
    // ---- pair tests --------------------------------------------------------------------------------
    // (a) later rounds only: entry j against the survivors of its class from earlier rounds
    if (kept_n > 0 && tid < K) {
        const unsigned short* kl = buf.kidx + (size_t)my_g * max_keep;
        const int ng = buf.ngroup[my_g];
        bool hit = false;
        for (int i = 0; i < ng && !hit; ++i) hit = suppresses_rt(METRIC, buf.kept[kl[i]], my_box, thr, thr_lo);
        if (hit) atomicAnd(reinterpret_cast<unsigned*>(buf.aliveW + my_g * MW) + (my_m >> 5), ~(1u << (my_m & 31)));
    }
    // (b) every entry against the earlier entries of its class (boxes and cell masks sit in class order).  The loop
    //     only runs the cheap tests (common cell, IoU gate); the ~3 % of pairs that pass are listed and get the
    //     exact metric in a second, dense pass -- in one divergent loop every warp would execute the exact path
    //     in most iterations.
    if (tid == 0) us.counter = 0;
    __syncthreads();
    auto settle = [&](int g, int cbase, int aq, int m) {          // bit m of row aq (rows in class order) = "aq suppresses m"
        const BoxC S = buf.cbox[cbase + aq], c = buf.cbox[cbase + m];
        float inter, uni;
        iou_gate(S, c, thr_lo, inter, uni);
        if (suppresses_exact_rt(METRIC, S, c, inter, uni, thr)) {
            atomicOr(reinterpret_cast<unsigned*>(buf.mat + (size_t)(cbase + aq) * MW) + (m >> 5), 1u << (m & 31));
            atomicOr(reinterpret_cast<unsigned*>(buf.nzW + g * MW) + (aq >> 5), 1u << (aq & 31));
        }
    };
    {
        // All (earlier, later) pairs of class positions, class after class, form one index space of
        // T = sum n_c (n_c - 1) / 2 tests; every thread takes a contiguous run of ceil(T / IT) of them and walks it
        // (later position m, earlier position aq: aq = 0 .. m-1, then m + 1, then the next class), so the work is
        // balanced no matter how large individual boxes or classes are.
        int T = 0;
        for (int g = 0; g < n_groups; ++g) { const int n_c = buf.coff[g + 1] - buf.coff[g]; T += n_c * (n_c - 1) / 2; }
        const int W = (T + IT - 1) / IT;
        int t = tid * W;
        const int t_end = min(T, t + W);
        if (t < t_end) {
            int g = 0, n_c = 0, r = t;
            for (;; ++g) {                                          // the class that owns test t
                n_c = buf.coff[g + 1] - buf.coff[g];
                const int pairs = n_c * (n_c - 1) / 2;
                if (r < pairs) break;
                r -= pairs;
            }
            int m = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)r)) * 0.5f);      // r = m (m - 1) / 2 + aq, 0 <= aq < m
            while (m * (m - 1) / 2 > r) --m;
            while ((m + 1) * m / 2 <= r) ++m;
            int aq = r - m * (m - 1) / 2;
            int cbase = buf.coff[g];
            BoxC c = buf.cbox[cbase + m];
            unsigned cells = buf.cmask[cbase + m];
            for (; t < t_end; ++t) {
                if ((buf.cmask[cbase + aq] & cells) != 0u) {        // a common cell (else: empty intersection)
                    float inter, uni;
                    if (iou_gate(buf.cbox[cbase + aq], c, thr_lo, inter, uni)) {
                        const int slot = atomicAdd(&us.counter, 1);
                        if (slot < PAIRS_CAP) buf.plist[slot] = (unsigned)aq | ((unsigned)m << 9) | ((unsigned)g << 18);
                        else settle(g, cbase, aq, m);
                    }
                }
                if (++aq == m) {                                    // next later position / next class
                    aq = 0;
                    if (++m == n_c && t + 1 < t_end) {
                        do { ++g; n_c = buf.coff[g + 1] - buf.coff[g]; } while (n_c < 2);
                        cbase = buf.coff[g];
                        m = 1;
                    }
                    if (t + 1 < t_end) { c = buf.cbox[cbase + m]; cells = buf.cmask[cbase + m]; }
                }
            }
        }
    }
    __syncthreads();
    {
        const int n_list = min(us.counter, PAIRS_CAP);
        for (int e = tid; e < n_list; e += IT) {
            const unsigned pr = buf.plist[e];
            const int g = (int)(pr >> 18);
            settle(g, buf.coff[g], (int)(pr & 511u), (int)((pr >> 9) & 511u));
        }
    }
    __syncthreads();

    SSDHOT_NSTAMP(7);
    // ---- resolve: one warp per class, rows in score order ----------------------------------------------
    for (int g = warp; g < n_groups; g += IT / 32) {
        const int cbase = buf.coff[g], n_c = buf.coff[g + 1] - cbase;
        if (n_c == 0) continue;
        const int words = (n_c + 63) >> 6;
        if (words == 1) {                                   // the usual case: everything in the registers of one lane
            if (lane == 0) {
                unsigned long long a = buf.aliveW[g * MW];
                const unsigned long long nz = buf.nzW[g * MW];
                unsigned long long pend = a & nz;
                while (pend) {
                    const int bit = __ffsll((long long)pend) - 1;
                    a &= ~buf.mat[(size_t)(cbase + bit) * MW];
                    pend = a & nz & ~((2ull << bit) - 1ull);
                }
                buf.aliveW[g * MW] = a;
            }
            continue;
        }
        unsigned long long a = lane < MW ? buf.aliveW[g * MW + lane] : 0ull;
        for (int w0 = 0; w0 < words; ++w0) {
            const unsigned long long nz = buf.nzW[g * MW + w0];
            unsigned long long pend = __shfl_sync(FULL, a, w0) & nz;
            while (pend) {
                const int bit = __ffsll((long long)pend) - 1;
                if (lane < MW) a &= ~buf.mat[(size_t)(cbase + w0 * 64 + bit) * MW + lane];
                pend = __shfl_sync(FULL, a, w0) & nz & ~((2ull << bit) - 1ull);
            }
        }
        if (lane < MW) buf.aliveW[g * MW + lane] = a;
    }
    __syncthreads();

    SSDHOT_NSTAMP(8);
    // ---- emit survivors in global score order ------------------------------------------------------------
    bool alive = false;
    if (my_g >= 0) alive = (buf.aliveW[my_g * MW + (my_m >> 6)] >> (my_m & 63)) & 1ull;
    const unsigned ab = __ballot_sync(FULL, alive);
    if (lane == 0) us.iscratch[warp] = __popc(ab);
    __syncthreads();
    int before = __popc(ab & lt), total_alive = 0;
    for (int w = 0; w < IT / 32; ++w) { const int c = us.iscratch[w]; if (w < warp) before += c; total_alive += c; }
    const int pos = kept_n + before;
    if (alive && pos < max_keep) {
        const BoxC bx = my_box;
        buf.kept[pos] = bx;
        // survivors of the class ahead of this one in the round: alive positions below m
        int crank = 0;
        for (int w = 0; w < (my_m >> 6); ++w) crank += __popcll(buf.aliveW[my_g * MW + w]);
        crank += __popcll(buf.aliveW[my_g * MW + (my_m >> 6)] & ((1ull << (my_m & 63)) - 1ull));
        buf.kidx[(size_t)my_g * max_keep + buf.ngroup[my_g] + crank] = (unsigned short)pos;
        const unsigned long long key = buf.ckey[tid];
        const unsigned id = 0xffffffffu - (unsigned)(key & 0xffffffffull);
        prm.out_labels[o + pos] = (int64_t)(id % (unsigned)n_fg);
        prm.out_scores[o + pos] = __uint_as_float((unsigned)(key >> 32) & 0x7fffffffu);
        reinterpret_cast<float4*>(prm.out_boxes)[o + pos] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
        if (prm.out_cand) prm.out_cand[o + pos] = (int32_t)id;
    }
    __syncthreads();
    if (alive && pos < max_keep) atomicAdd(&buf.ngroup[my_g], 1);
    return total_alive;
}

// The rounds of one image on the GLOBAL candidate lists (every thread of the CTA).  Preconditions: buf.hist16 and buf.ngroup
// are zero, the head tables are filled, n_cand = candidates in `src`, and a barrier has published all of it.
template <bool APPROX, int SRC, typename Src>
__device__ __forceinline__ void nms_image_body(const PredictParams& prm, const int b, const ImgBuffers& buf, UnitShared& us,
                                               const HeadTable& loc_tab, const HeadTable& conf_tab, const Src& src, const int n_cand) {
    constexpr int KEY_MARGIN = 512;                // ulps: >= 3e-5 relative, 3x the worst error of an approximate score
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_fg = prm.C - 1, P = prm.P, max_keep = prm.max_keep;
    const int METRIC = prm.metric;
    const bool AGN = prm.agnostic != 0;
    const int n_groups = AGN ? 1 : n_fg;
    bool first = true;
    const float thr = prm.nms_thresh, thr_lo = fmul(thr, kFilterSlack);
    const HeadReader<SRC, 4> loc_rd = {SRC == SRC_PACKED ? prm.loc_all + 4ll * b * P : nullptr, &loc_tab};
    const HeadReader<SRC, 6> conf_rd = {SRC == SRC_PACKED ? prm.conf_all + (long long)b * P * prm.C : nullptr, &conf_tab};
    const bool want_atan = METRIC == SSDHOT_METRIC_CIOU;
    const long long o = (long long)b * max_keep;
    const unsigned lt = (1u << lane) - 1u;
    auto exact_key = [&](unsigned id) -> unsigned { return exact_score_key6(conf_rd, (int)(id / 5u), (int)(id % 5u)); };

    bool use_hist = n_cand > CH && n_cand < 65536;          // 16-bit bin counters
    bool keys_exact = !APPROX;
    if (use_hist) src.template for_each<IT>([&](int, unsigned long long r) { hist_add(buf.hist16, Src::key(r)); });
    __syncthreads();
    auto make_exact = [&]() { nms_make_exact<SRC>(src, buf.hist16, conf_rd, use_hist); keys_exact = true; };

    SSDHOT_NSTAMP(1);
    int kept_n = 0, remaining = n_cand;
    while (remaining > 0 && kept_n < max_keep) {
        // ---- pull ----------------------------------------------------------------------------------
        int K = remaining < CH ? remaining : CH;
        unsigned tkey = 1u, tie_floor = 0u;
        bool all = remaining <= CH;
        if (!all && use_hist) {
            int cap = keys_exact ? CH : CH - 16;            // (room for the few margin strays)
            if (first) cap = min(cap, max(64, max_keep + (max_keep >> 1) + 16));
            hist_cut<IT>(buf.hist16, cap, us);
            if (us.cut_count > 0) { K = us.cut_count; tkey = bin_floor_key(us.cut_bin); all = true; }
            else use_hist = false;                          // a single bin holds more than the cap: exact select from here on
            __syncthreads();
        }
        if (!all) {
            if (APPROX && !keys_exact) make_exact();
            nms_select_cut(src, us, K, &tkey, &tie_floor);
        }
        SSDHOT_NSTAMP(2);
        const unsigned gkey = (keys_exact || tkey <= (unsigned)KEY_MARGIN) ? tkey : tkey - (unsigned)KEY_MARGIN;
        if (tid == 0) us.counter = 0;
        __syncthreads();
        src.for_each<IT>([&](int i, unsigned long long r) {
            const unsigned k = Src::key(r);
            if (k == 0u || k < gkey) return;
            if (keys_exact && k == tkey && (unsigned)(r & 0xffffffffull) < tie_floor) return;
            const int pos = atomicAdd(&us.counter, 1);
            if (pos < CH) { buf.ckey[pos] = Src::sortkey(r, i); buf.cidx[pos] = i; }
            if (keys_exact) src.consume(i);
        });
        __syncthreads();
        SSDHOT_NSTAMP(3);
        if (APPROX && !keys_exact) {
            const int gathered = us.counter;
            if (gathered > CH) {                            // more margin strays than the slack: settle it with exact keys
                __syncthreads();
                make_exact();
                continue;
            }
            bool valid = false;                             // exact scores; a candidate that falls below the cut stays in the pool
            unsigned long long out = 0ull;
            if (tid < gathered) {
                const unsigned long long sk = buf.ckey[tid];
                const unsigned ek = exact_key(0xffffffffu - (unsigned)(sk & 0xffffffffull));
                valid = ek >= tkey;
                if (valid) { out = ((unsigned long long)ek << 32) | (sk & 0xffffffffull); src.consume(buf.cidx[tid]); }
            }
            buf.ckey[tid] = out;
            K = __syncthreads_count(valid);
        } else {
            if (tid >= K) buf.ckey[tid] = 0ull;
            __syncthreads();
        }
        SSDHOT_NSTAMP(4);
        if (K == 0) {                                                           // (only margin strays: the next cut is lower)
            first = false;
            continue;
        }
        const int total_alive = nms_round_backend<SRC>(prm, b, buf, us, loc_rd, K, kept_n, first);
        kept_n = min(max_keep, kept_n + total_alive);
        remaining -= K;
        __syncthreads();
        SSDHOT_NSTAMP(9);
        first = false;
    }
    if (tid == 0) { prm.out_count[b] = kept_n; if (prm.timeline) { prm.timeline[(long long)b * 16 + 10] = globaltimer_ns(); unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); prm.timeline[(long long)b * 16 + 11] = sm; prm.timeline[(long long)b * 16 + 12] = (unsigned long long)n_cand; } }
}

template <bool APPROX, int SRC>
__global__ void __launch_bounds__(IT, 2) nms_image_kernel(const PredictParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ UnitShared us;
    __shared__ HeadTable loc_tab, conf_tab;        // per-level bases of this image (head sources only)
    const int tid = threadIdx.x, b = blockIdx.x;
    const int n_fg = prm.C - 1, n_groups = prm.agnostic ? 1 : n_fg;
    const ImgBuffers buf = carve_img(dyn, prm.max_keep, n_groups);
    SegSource src;
    src.seg_cap = seg_cap_rows(prm.P) * n_fg;
    src.base = prm.cand + (long long)b * SEGS * src.seg_cap;
    src.counts = prm.cand_count + b * SEGS;
    for (int i = tid; i < HBINS / 2; i += IT) buf.hist16[i] = 0u;
    for (int g = tid; g < n_groups; g += IT) buf.ngroup[g] = 0;
    pdl_wait();                                    // score_kernel's lists and counts are complete from here on
    if (prm.timeline && tid == 0) prm.timeline[(long long)b * 16 + 0] = globaltimer_ns();
    const int n_cand = block_sum<int>(tid < SEGS ? src.counts[tid] : 0, us.iscratch);
    if (SRC != SRC_PACKED) {                       // (published by the __syncthreads() below)
        head_table_fill<SRC, 4>(loc_tab, prm.loc_h, b, tid);
        head_table_fill<SRC, 6>(conf_tab, prm.conf_h, b, tid - 32);
    }
    __syncthreads();
    nms_image_body<APPROX, SRC>(prm, b, buf, us, loc_tab, conf_tab, src, n_cand);
}


// ---- predict in ONE kernel: stream -> row keys in shared memory -> hot rows -> one NMS round -----------------------------
// predict_image_kernel (one CTA of 512 threads per image, two per SM) serves the case every typical image is: the
// detections come out of ONE round over the few hundred best candidates.  Those candidates sit in few rows, so the stream
// does not list candidates at all.  Per row it keeps ONE 16-bit key -- the histogram bin of the row's best approximate
// foreground score (0: no class can pass the score threshold) -- in shared memory: no per-candidate work, no compaction, no
// list in HBM.  Then: histogram of the row keys -> the cut that leaves <= ~1.5 max_per_img rows -> the "hot" rows (key at most
// one bin below the cut: an approximate score is within 1.1e-5 of the exact one, a bin is 3.9e-3 wide) -> one thread per
// hot row evaluates the row with the exact eager-CUDA arithmetic (SFS:388, :402) and appends every class whose exact score
// passes the threshold and reaches the cut.  Every candidate at or above the cut lives in a hot row, so the pulled set is
// exactly the set the two-kernel path pulls for the same cut, and the shared back end (nms_round_backend) does the rest.
// Whatever does not fit that mould -- no usable cut, more than CH pulled candidates, a second round -- makes the CTA redo
// its image on the generic path: score_segment into the global lists, then nms_image_body on them.
constexpr int HOT_CAP = CH;        // hot rows of a round (one thread each)

// One warp's share of the stream: key16[p] for the rows of list segment `seg` (same row ranges as score_segment).
template <int SRC>
__device__ __forceinline__ void stream_row_keys(const PredictParams& prm, int b, int seg, int lane, const HeadTable& htab,
                                                const PlaneRegions& regions, unsigned* __restrict__ rowkey32,
                                                unsigned* __restrict__ hist16, int* __restrict__ n_rows_cand,
                                                const float* x_first = nullptr) {
    const int P = prm.P;
    const int rows = seg_rows(P);
    const int r0 = min(P, seg * rows), r1 = min(P, r0 + rows);
    const float* conf_b = SRC == SRC_PACKED ? prm.conf_all + (long long)b * P * 6 : nullptr;
    const float thr_lo = prm.score_thresh * 0.9999f;    // a row whose best approximate score is below this has no candidate
    const HeadReader<SRC, 6> rd = {conf_b, &htab};
    const int q0 = r0 >> 1, q1 = r1 >> 1;               // P is even on this path
    const int n_it = SRC == SRC_LEVEL_PLANES ? (kPlaneChunks - seg + SEGS - 1) / SEGS : (q1 - q0 + 31) >> 5;
    PlaneWalk walk(&regions);
    auto load_it = [&](int it, float* xx, int& p0) -> bool {
        if (it >= n_it) return false;
        if (SRC == SRC_LEVEL_PLANES) {
            return walk.load(seg + SEGS * it, lane, xx, p0);
        } else {
            const int q = q0 + 32 * it + lane;
            p0 = 2 * q;
            if (q < q1) rd.pair(q, xx);
            return q < q1;
        }
    };
    constexpr float kL2E = 1.4426950408889634f;
    int n_cand_rows = 0;                                // this lane's rows that can hold a candidate
    auto process = [&](const float* x, int p0, bool live) {
        if (!live) return;
        unsigned packed = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float* xr = x + 6 * h;
            const float mx = fmaxf(fmaxf(fmaxf(xr[0], xr[1]), fmaxf(xr[2], xr[3])), fmaxf(xr[4], xr[5]));
            const float nm = -mx * kL2E;                // the common shift cancels in e_k / sum (a NaN / Inf row stays NaN)
            float e[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) e[i] = ex2_approx_ftz(fmaf(xr[i], kL2E, nm));
            const float sum = ((e[0] + e[1]) + (e[2] + e[3])) + (e[4] + e[5]);
            const float emax = fmaxf(fmaxf(fmaxf(e[1], e[2]), fmaxf(e[3], e[4])), e[5]);
            const float sc = emax * rcp_approx_ftz(sum);                    // the row's best approximate foreground score
            unsigned k16 = max(__float_as_uint(sc) >> 15, 1u);              // its histogram bin (+ HBASE); never 0 for a candidate row
            if (emax >= thr_lo * sum) {                                     // (NaN rows: the comparison is false -> no candidate)
                packed |= k16 << (16 * h);
                hist_add(hist16, 0x80000000u | (k16 << 15));                // the histogram of the row keys rides along
                ++n_cand_rows;
            }
        }
        rowkey32[p0 >> 1] = packed;
    };
    float xa[12], xb[12];
    int pa = 0, pb = 0;
    bool la, lb = false;
    if (SRC == SRC_PACKED && x_first) {                // (the first row pair was requested at the top of the kernel)
#pragma unroll
        for (int j = 0; j < 12; ++j) xa[j] = x_first[j];
        pa = 2 * (q0 + lane);
        la = n_it > 0 && q0 + lane < q1;
    } else la = load_it(0, xa, pa);
    for (int it = 0; it < n_it; it += 2) {             // two iterations per trip: the next pair's loads are in flight
        lb = load_it(it + 1, xb, pb);
        process(xa, pa, la);
        la = load_it(it + 2, xa, pa);
        process(xb, pb, lb);
    }
    n_cand_rows = __reduce_add_sync(FULL, n_cand_rows);
    if (lane == 0 && n_cand_rows) atomicAdd(n_rows_cand, n_cand_rows);
}

// The row keys of image b as left by train_image_kernel's stream of the same logits (flag 2), or false when they are not coming:
// flag still 0 a few microseconds after this CTA started (the loss kernel took another path, or its CTA is not resident yet) or
// not delivered within kShareWaitNs.  Called by one thread.
constexpr unsigned long long kShareGraceNs = 6000ull, kShareWaitNs = 150000ull;
__device__ __forceinline__ bool share_keys_arrive(const int* flag) {
    int f = ld_acquire_gpu(flag);
    if (f == 2) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        __nanosleep(200);
        f = ld_acquire_gpu(flag);
        if (f == 2) return true;
        const unsigned long long dt = globaltimer_ns() - t0;
        if ((f == 0 && dt > kShareGraceNs) || dt > kShareWaitNs) return false;
    }
}
// Row keys from the hand-off buffer into shared memory: rows under the threshold's bin are blanked, the others counted and
// entered in the histogram -- what stream_row_keys leaves, with the candidate test at the resolution of a bin (a superset:
// the exact pass applies SFS:402 to every row it looks at).
template <int NT>
__device__ __forceinline__ void load_row_keys(const unsigned* __restrict__ keys, int P, float score_thresh, int tid,
                                              unsigned* __restrict__ rowkey32, unsigned* __restrict__ hist16, int* __restrict__ n_rows_cand) {
    const float thr_lo = score_thresh * 0.9999f;
    const unsigned kthr = thr_lo > 0.0f ? max(__float_as_uint(thr_lo) >> 15, 1u) : 1u;
    const int n_words = P / 2;
    int n = 0;
    for (int i = 2 * tid; i < n_words; i += 2 * NT) {
        const uint2 w2 = __ldcg(reinterpret_cast<const uint2*>(keys + i));
        unsigned out[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned w = j ? w2.y : w2.x;
            unsigned lo = w & 0xffffu, hi = w >> 16;
            if (i + j >= n_words) { lo = 0u; hi = 0u; }
            if (lo >= kthr) { hist_add(hist16, 0x80000000u | (lo << 15)); ++n; } else lo = 0u;
            if (hi >= kthr) { hist_add(hist16, 0x80000000u | (hi << 15)); ++n; } else hi = 0u;
            out[j] = lo | (hi << 16);
        }
        rowkey32[i] = out[0];
        if (i + 1 < n_words) rowkey32[i + 1] = out[1];
    }
    n = __reduce_add_sync(FULL, n);
    if ((tid & 31) == 0 && n) atomicAdd(n_rows_cand, n);
}

// buf.ckey[0 .. K) (K <= CH, distinct 64-bit keys: exact score bits << 32 | ~candidate id) into descending order by COUNTING
// instead of a CH-wide bitonic sort (45 compare-exchange stages, ~11.7 k warp-instructions an image): a key's position is the
// number of keys in higher score bins -- one pass over the 4096-bin histogram, 8 bins per thread -- plus its rank among the
// one to three keys of its own bin.  hist16 (left cleared), buf.plist and buf.cbox serve as scratch: all free between the exact
// keys and the decode.  Every thread of the CTA; ends with a barrier.
__device__ __forceinline__ void rank_keys_by_counting(const ImgBuffers& buf, UnitShared& us, const int K) {
    static_assert(IT * 8 == HBINS && PAIRS_CAP * 4 >= HBINS * 2 && CH * sizeof(BoxC) >= CH * 8, "scratch sizes");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned short* base16 = reinterpret_cast<unsigned short*>(buf.plist);         // [HBINS] keys in higher bins
    unsigned long long* tmp = reinterpret_cast<unsigned long long*>(buf.cbox);     // [CH] keys grouped by bin
    for (int i = tid; i < HBINS / 2; i += IT) buf.hist16[i] = 0u;
    const unsigned long long key = tid < K ? buf.ckey[tid] : 0ull;
    __syncthreads();
    int bin = 0, arrival = 0;
    if (tid < K) {
        bin = score_bin((unsigned)(key >> 32));
        const unsigned old = atomicAdd(&buf.hist16[bin >> 1], 1u << ((bin & 1) * 16));
        arrival = (int)((old >> ((bin & 1) * 16)) & 0xffffu);
    }
    __syncthreads();
    {
        // thread t owns bins [HBINS - 8 (t + 1), HBINS - 8 t), walked from the top
        const uint4 w = *reinterpret_cast<const uint4*>(buf.hist16 + (HBINS / 2 - 4 * (tid + 1)));
        const unsigned c[8] = {w.w >> 16, w.w & 0xffffu, w.z >> 16, w.z & 0xffffu, w.y >> 16, w.y & 0xffffu, w.x >> 16, w.x & 0xffffu};
        int tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) tot += (int)c[j];
        int incl = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) us.iscratch[warp] = incl;
        __syncthreads();
        int run = incl - tot;
        for (int w2 = 0; w2 < warp; ++w2) run += us.iscratch[w2];
        if (tot) {                                   // (bases of empty bins are never read)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                base16[HBINS - 1 - 8 * tid - j] = (unsigned short)run;
                run += (int)c[j];
            }
        }
    }
    __syncthreads();
    int base = 0;
    if (tid < K) {
        base = (int)base16[bin];
        tmp[base + arrival] = key;
    }
    __syncthreads();
    int pos = 0;
    if (tid < K) {
        const int cnt = (int)((buf.hist16[bin >> 1] >> ((bin & 1) * 16)) & 0xffffu);
        pos = base;
        for (int j = 0; j < cnt; ++j) pos += tmp[base + j] > key ? 1 : 0;
    }
    __syncthreads();
    if (tid < K) {
        buf.ckey[pos] = key;
        buf.hist16[bin >> 1] = 0u;                   // (every thread of a bin pair stores the same zero)
    }
    __syncthreads();
}

template <int SRC>
__global__ void __launch_bounds__(IT, 2) predict_image_kernel(const PredictParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ UnitShared us;
    __shared__ HeadTable loc_tab, conf_tab;        // per-level bases of this image (head sources only)
    __shared__ PlaneRegions regions;               // (NCHW heads)
    __shared__ int n_rows_cand, n_hot, more_low, keys_given;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const int n_fg = prm.C - 1, n_groups = prm.agnostic ? 1 : n_fg, P = prm.P, max_keep = prm.max_keep;
    const ImgBuffers buf = carve_img(dyn, max_keep, n_groups);
    // the row keys and the hot-row list live where the suppression rows and class-ordered boxes will be (dead before those are written)
    unsigned* rowkey32 = reinterpret_cast<unsigned*>(buf.mat);                 // [P / 2] two 16-bit keys per word
    int* hot = reinterpret_cast<int*>(rowkey32 + ((P / 2 + 3) & ~3));            // [HOT_CAP]
    // a CTA that streams itself requests its first row pairs before anything else: their DRAM latency passes under the set-up
    float x_first[12];
    const bool pre = SRC == SRC_PACKED && !prm.share_keys;
    if (pre) {
        const int rows = seg_rows(P), r0 = min(P, warp * rows), r1 = min(P, r0 + rows), q = (r0 >> 1) + lane;
        if (q < (r1 >> 1)) {
            const HeadReader<SRC, 6> rd0 = {prm.conf_all + (long long)b * P * 6, nullptr};
            rd0.pair(q, x_first);
        }
    }
    for (int i = tid; i < HBINS / 2; i += IT) buf.hist16[i] = 0u;
    for (int g = tid; g < n_groups; g += IT) buf.ngroup[g] = 0;
    if (tid == 0) { n_rows_cand = 0; n_hot = 0; more_low = 0; us.counter = 0; }
    if (SRC != SRC_PACKED) {
        head_table_fill<SRC, 4>(loc_tab, prm.loc_h, b, tid);
        head_table_fill<SRC, 6>(conf_tab, prm.conf_h, b, tid - 32);
        if (SRC == SRC_LEVEL_PLANES) plane_regions_fill(regions, prm.conf_h, b, tid - 64);
    }
    if (prm.timeline && tid == 0) prm.timeline[(long long)b * 16 + 14] = globaltimer_ns();
    if (tid == IT - 1) {
        const bool got = prm.share_keys && share_keys_arrive(prm.share_flag + b);
        keys_given = got ? 1 : 0;
        if (got) prm.share_flag[b] = 3;             // (picked up: what the tests count; nobody waits for it)
    }
    __syncthreads();
    if (keys_given) load_row_keys<IT>(prm.share_keys + (size_t)b * share_stride_words(P), P, prm.score_thresh, tid, rowkey32, buf.hist16, &n_rows_cand);
    else stream_row_keys<SRC>(prm, b, warp, lane, conf_tab, regions, rowkey32, buf.hist16, &n_rows_cand, pre ? x_first : nullptr);
    __syncthreads();
    bool first = true;
    SSDHOT_NSTAMP(0);
    const HeadReader<SRC, 4> loc_rd = {SRC == SRC_PACKED ? prm.loc_all + 4ll * b * P : nullptr, &loc_tab};
    const HeadReader<SRC, 6> conf_rd = {SRC == SRC_PACKED ? prm.conf_all + (long long)b * P * prm.C : nullptr, &conf_tab};
    const int rows_cand = n_rows_cand;
    int fail = 0;                                   // (CTA-uniform) why the image leaves the fused mould, 0 = it does not
    unsigned tkey = 1u, floor16 = 1u;               // pulled: exact key >= tkey; hot: row key >= floor16
    const int cap_rows = min(HOT_CAP - 16, max(64, max_keep + (max_keep >> 2) + 16));      // rows of the round (~1.1 candidates each)
    if (rows_cand > cap_rows) {
        if (rows_cand >= 65536) fail = 2;           // (16-bit bin counters)
        else {
            hist_cut<IT>(buf.hist16, cap_rows, us);
            if (us.cut_count > 0) {
                tkey = bin_floor_key(us.cut_bin);
                floor16 = us.cut_bin - 1 >= 1 ? (unsigned)(us.cut_bin - 1 + HBASE) : 1u;      // one bin below the cut
            } else fail = 2;                        // a single bin holds more rows than a round takes
        }
    }
    SSDHOT_NSTAMP(2);
    int kept_n = 0;
    if (!fail && rows_cand > 0) {
        // ---- hot rows ----------------------------------------------------------------------------------------------------
        for (int q = tid; q < P / 2; q += IT) {
            const unsigned w = rowkey32[q];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned k = h ? (w >> 16) : (w & 0xffffu);
                if (k >= floor16 && k != 0u) {
                    const int at = atomicAdd(&n_hot, 1);
                    if (at < HOT_CAP) hot[at] = 2 * q + h;
                }
            }
        }
        __syncthreads();
        SSDHOT_NSTAMP(3);
        const int hot_n = n_hot;
        if (hot_n > HOT_CAP) fail = 3;
        else {
            // ---- one thread per hot row: exact scores (SFS:388), exact threshold test (SFS:402), exact keys ------------------
            if (tid < hot_n) {
                const int p = hot[tid];
                float x[6], e[6];
                conf_rd.row(p, x);
                // the decode of this row's candidates comes three barriers later and finds the box offsets cold in HBM: ask for them now
                // (step 60.4 -> 59.2 us; the same for the hot rows' logits, one barrier ahead of the exact pass: no gain)
                if (SRC != SRC_LEVEL_PLANES) asm volatile("prefetch.global.L2 [%0];" ::"l"(loc_rd.row_ptr(p)));
                const float sum = row_exps6(x, e);
                bool low = false;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float sc = fdiv(e[k + 1], sum);
                    if (sc > prm.score_thresh) {
                        const unsigned ek = __float_as_uint(sc) | 0x80000000u;
                        if (ek >= tkey) {
                            const int at = atomicAdd(&us.counter, 1);
                            if (at < CH) buf.ckey[at] = ((unsigned long long)ek << 32) | (unsigned long long)(0xffffffffu - (unsigned)(p * 5 + k));
                        } else low = true;          // a candidate below the cut: it would belong to a later round
                    }
                }
                if (low) more_low = 1;              // (benign race: every writer stores 1)
            }
            __syncthreads();
            const int K = us.counter;
            if (K > CH) fail = 3;
            else if (K == 0) fail = 4;              // (only strays below the cut)
            else {
                for (int i = K + tid; i < CH; i += IT) buf.ckey[i] = 0ull;
                SSDHOT_NSTAMP(4);
                rank_keys_by_counting(buf, us, K);
                const int total_alive = nms_round_backend<SRC>(prm, b, buf, us, loc_rd, K, 0, first, true);
                kept_n = min(max_keep, total_alive);
                __syncthreads();
                SSDHOT_NSTAMP(9);
                // candidates outside the round: rows that were not hot, or classes of hot rows below the cut
                if (kept_n < max_keep && (rows_cand > hot_n || more_low)) fail = 5;
            }
        }
    }
    if (!fail) {
        if (tid == 0) {
            prm.out_count[b] = kept_n;
            if (prm.timeline) { prm.timeline[(long long)b * 16 + 10] = globaltimer_ns(); unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); prm.timeline[(long long)b * 16 + 11] = sm; prm.timeline[(long long)b * 16 + 12] = (unsigned long long)rows_cand; }
        }
        return;
    }
    SSDHOT_FAST_EXIT(fail, kept_n);
    // ---- the generic path, by this CTA alone: global lists, as many rounds as it takes -------------------------------
    __syncthreads();
    score_segment<6, SRC>(prm, b, warp, lane, conf_tab, regions);
    for (int i = tid; i < HBINS / 2; i += IT) buf.hist16[i] = 0u;
    for (int g = tid; g < n_groups; g += IT) buf.ngroup[g] = 0;
    __syncthreads();                               // (the lists and counts this CTA wrote are visible to all its threads)
    SegSource gsrc;
    gsrc.seg_cap = seg_cap_rows(P) * n_fg;
    gsrc.base = prm.cand + (long long)b * SEGS * gsrc.seg_cap;
    gsrc.counts = prm.cand_count + b * SEGS;
    const int n_all = block_sum<int>(tid < SEGS ? gsrc.counts[tid] : 0, us.iscratch);
    __syncthreads();
    nms_image_body<true, SRC>(prm, b, buf, us, loc_tab, conf_tab, gsrc, n_all);
}

// ---- stand-alone NMS -------------------------------------------------------------------------------
template <int METRIC>
__global__ void __launch_bounds__(UT) nms_sets_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                      const int32_t* __restrict__ set_offsets, int n_max, float thr, int max_keep,
                                                      int64_t* __restrict__ keep, int32_t* __restrict__ keep_count,
                                                      BoxC* __restrict__ kept_all) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ UnitShared us;
    const int tid = threadIdx.x;
    const int begin = set_offsets[blockIdx.x];
    const int n = set_offsets[blockIdx.x + 1] - begin;
    UnitBuffers buf = {};
    buf.cbox = reinterpret_cast<BoxC*>(dyn);
    buf.ckey = reinterpret_cast<unsigned long long*>(buf.cbox + CHUNK);
    DenseSource src;
    src.keys = reinterpret_cast<unsigned*>(buf.ckey + CHUNK);
    src.n = n;
    buf.cgroup = reinterpret_cast<unsigned char*>(src.keys + n_max);
    buf.kept = kept_all + begin;            // survivors of a stand-alone call are unbounded: global memory
    for (int i = tid; i < n; i += UT) {
        const unsigned k = ord_encode(__ldg(scores + begin + i));
        src.keys[i] = k == 0u ? 1u : k;
    }
    __syncthreads();
    const bool want_atan = METRIC == SSDHOT_METRIC_CIOU;
    auto fetch = [&](unsigned id) -> BoxC {
        const float4 bx = ldg4(boxes + 4ll * (begin + (long long)id));
        return box_consts(bx.x, bx.y, bx.z, bx.w, want_atan);
    };
    auto group_of = [&](unsigned) -> int { return 0; };
    int64_t* out = keep + begin;
    auto emit = [&](int pos, unsigned long long, unsigned id, const BoxC&) { out[pos] = (int64_t)id; };
    const int cap = (max_keep > 0 && max_keep < n) ? max_keep : n;
    const int kept_n = nms_unit<METRIC, UT, false, false, false>(src, buf, n, 0, cap, thr, us, fetch, group_of, emit,
                                                                 [](unsigned) -> unsigned { return 0u; });
    if (tid == 0) keep_count[blockIdx.x] = kept_n;
}

__global__ void decode_kernel(const float* __restrict__ loc, const float* __restrict__ pri, int M, float vc, float vs,
                              float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    reinterpret_cast<float4*>(out)[i] = decode_box(ldg4(loc + 4ll * i), ldg4(pri + 4ll * i), vc, vs);
}


constexpr size_t kMaxDynSmem = 227 * 1024 - sizeof(UnitShared) - 1024;

template <typename K>
static int set_smem(K kern, size_t bytes) {
    if (bytes > kMaxDynSmem) return SSDHOT_ERR_SHAPE;
    if (bytes > 40 * 1024)        // static + dynamic shared memory share the 48 KB default limit
        return ensure_dyn_smem(reinterpret_cast<const void*>(kern), kMaxDynSmem);
    return SSDHOT_OK;
}

template <int SRC>
static int launch_predict_image(const PredictParams& prm, size_t dyn, cudaStream_t stream) {
    int rc;
    if ((rc = set_smem(predict_image_kernel<SRC>, dyn))) return rc;
    predict_image_kernel<SRC><<<prm.B, IT, dyn, stream>>>(prm);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

template <bool APPROX, int SRC>
static int launch_nms_image(const PredictParams& prm, size_t dyn, cudaStream_t stream) {
    int rc;
    if ((rc = set_smem(nms_image_kernel<APPROX, SRC>, dyn))) return rc;
    // PDL: the CTAs may start (shared-memory carve-up, histogram clear) while score_kernel drains
    cudaError_t e = launch_pdl(nms_image_kernel<APPROX, SRC>, dim3(prm.B), dim3(IT), dyn, stream, prm);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    return SSDHOT_OK;
}

}  // namespace ssdhot

using namespace ssdhot;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int ssdhot_decode(const float* loc, const float* priors_cxcywh, int M, float var_center, float var_size,
                             float* out, ssdhot_stream_t stream) {
    if (!loc || !priors_cxcywh || !out) return SSDHOT_ERR_NULL;
    if (M < 0) return SSDHOT_ERR_SHAPE;
    if (!al16(loc) || !al16(priors_cxcywh) || !al16(out)) return SSDHOT_ERR_ALIGN;
    if (M == 0) return SSDHOT_OK;
    decode_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(loc, priors_cxcywh, M, var_center, var_size, out);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

extern "C" unsigned long long ssdhot_nms_workspace_bytes(long long total_boxes) {
    return total_boxes <= 0 ? 64ull : (unsigned long long)total_boxes * sizeof(BoxC) + 64ull;
}

extern "C" int ssdhot_nms(const float* boxes, const float* scores, const int32_t* set_offsets, int n_sets,
                          long long total_boxes, int max_set_size, float thresh, int metric, int max_keep,
                          int64_t* keep, int32_t* keep_count, void* work, ssdhot_stream_t stream) {
    if (!set_offsets || !keep_count) return SSDHOT_ERR_NULL;
    if (n_sets <= 0 || total_boxes < 0 || max_set_size < 0) return SSDHOT_ERR_SHAPE;
    if (total_boxes > 0 && (!boxes || !scores || !keep || !work)) return SSDHOT_ERR_NULL;
    if (!al16(boxes) || !al16(work)) return SSDHOT_ERR_ALIGN;
    const int n_max = max_set_size > 0 ? max_set_size : 1;
    const size_t dyn = (size_t)CHUNK * (sizeof(BoxC) + 8) + (size_t)n_max * 4 + (size_t)CHUNK + 16;
    BoxC* kept_all = reinterpret_cast<BoxC*>(work);
    int rc;
    cudaStream_t s = (cudaStream_t)stream;
    switch (metric) {
        case SSDHOT_METRIC_DIOU:
            if ((rc = set_smem(nms_sets_kernel<SSDHOT_METRIC_DIOU>, dyn))) return rc;
            nms_sets_kernel<SSDHOT_METRIC_DIOU><<<n_sets, UT, dyn, s>>>(boxes, scores, set_offsets, n_max, thresh, max_keep, keep, keep_count, kept_all);
            break;
        case SSDHOT_METRIC_CIOU:
            if ((rc = set_smem(nms_sets_kernel<SSDHOT_METRIC_CIOU>, dyn))) return rc;
            nms_sets_kernel<SSDHOT_METRIC_CIOU><<<n_sets, UT, dyn, s>>>(boxes, scores, set_offsets, n_max, thresh, max_keep, keep, keep_count, kept_all);
            break;
        case SSDHOT_METRIC_IOU:
            if ((rc = set_smem(nms_sets_kernel<SSDHOT_METRIC_IOU>, dyn))) return rc;
            nms_sets_kernel<SSDHOT_METRIC_IOU><<<n_sets, UT, dyn, s>>>(boxes, scores, set_offsets, n_max, thresh, max_keep, keep, keep_count, kept_all);
            break;
        default: return SSDHOT_ERR_VALUE;
    }
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

// workspace layout: cand_count [B][SEGS] int32 | cand [B][SEGS][seg_cap_rows(P)*(C-1)] uint64
static size_t pw_cand_off(int B) { return ((size_t)B * SEGS * 4 + 15) & ~(size_t)15; }

extern "C" unsigned long long ssdhot_predict_workspace_bytes(int B, int P, int C) {
    if (B <= 0 || P <= 0 || C < 2) return 64ull;
    return (unsigned long long)(pw_cand_off(B) + (size_t)B * SEGS * seg_cap_rows(P) * (C - 1) * 8 + 64);
}

extern "C" int ssdhot_predict(const float* priors_cxcywh, int P, const float* loc_all, const float* conf_all,
                              int B, int C, float score_thresh, float nms_thresh, int max_per_img,
                              int class_agnostic, int metric, float var_center, float var_size,
                              float img_w, float img_h,
                              int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                              int32_t* out_count, void* work, ssdhot_stream_t stream) {
    return ssdhot_predict_stages(priors_cxcywh, P, loc_all, conf_all, B, C, score_thresh, nms_thresh, max_per_img, class_agnostic,
                                 metric, var_center, var_size, img_w, img_h, out_labels, out_scores, out_boxes, out_cand, out_count,
                                 work, SSDHOT_STAGE_SCORES | SSDHOT_STAGE_NMS, nullptr, stream);
}

// Shared back end of ssdhot_predict_stages (src = SRC_PACKED) and ssdhot_predict_heads (per-level sources).
static int predict_launch(PredictParams& prm, int src, int class_agnostic, int metric, void* work, int stages, void* share, cudaStream_t s) {
    const int B = prm.B, C = prm.C, P = prm.P;
    if ((stages & ~(SSDHOT_STAGE_SCORES | SSDHOT_STAGE_NMS)) != 0 || stages == 0) return SSDHOT_ERR_VALUE;
    if (!prm.pri || !prm.out_labels || !prm.out_scores || !prm.out_boxes || !prm.out_count || !work) return SSDHOT_ERR_NULL;
    if (P <= 0 || B <= 0 || C < 2 || C > SSDHOT_MAX_CLASSES || prm.max_keep <= 0 || prm.max_keep > 65535) return SSDHOT_ERR_SHAPE;
    if ((long long)P * (C - 1) > 0x7fffffffll) return SSDHOT_ERR_SHAPE;
    // same validation as SSD_from_scratch.py:369-373
    if (!(prm.score_thresh >= 0.0f && prm.score_thresh < 1.0f) || !(prm.nms_thresh > 0.0f && prm.nms_thresh < 1.0f)) return SSDHOT_ERR_VALUE;
    if (metric != SSDHOT_METRIC_DIOU && metric != SSDHOT_METRIC_CIOU && metric != SSDHOT_METRIC_IOU) return SSDHOT_ERR_VALUE;
    if (!al16(prm.pri) || !al16(prm.out_boxes) || !al16(work)) return SSDHOT_ERR_ALIGN;
    const size_t dyn = img_smem_bytes(prm.max_keep, class_agnostic ? 1 : C - 1);
    if (dyn > kMaxDynSmem) return SSDHOT_ERR_SHAPE;
    prm.timeline = g_timeline;
    unsigned char* w = reinterpret_cast<unsigned char*>(work);
    prm.cand_count = reinterpret_cast<int*>(w);
    prm.cand = reinterpret_cast<unsigned long long*>(w + pw_cand_off(B));
    // C == 6 with 48-byte-aligned row pairs: approximate scores, refined by the NMS kernel; otherwise exact scores
    const bool approx = src != SRC_PACKED || (C == 6 && (P % 2) == 0 && al16(prm.conf_all));
    // both stages of an approximable input (C == 6, candidate ids < 65536): ONE kernel, the candidate lists stay in shared
    // memory (SSDHOT_PREDICT_TWO_KERNELS=1 keeps the two-kernel path, for A/B measurements)
    static const bool two_kernels = getenv("SSDHOT_PREDICT_TWO_KERNELS") != nullptr;
    prm.metric = metric;
    prm.agnostic = class_agnostic ? 1 : 0;
    if (stages == (SSDHOT_STAGE_SCORES | SSDHOT_STAGE_NMS) && approx && (long long)P * (C - 1) <= 65535 && !two_kernels) {
        if (share && al16(share) && C == 6) {           // the loss kernel of this step leaves the row keys of the same logits
            prm.share_flag = reinterpret_cast<int*>(share);
            prm.share_keys = reinterpret_cast<const unsigned*>(reinterpret_cast<unsigned char*>(share) + share_flags_bytes(B));
        }
        if (src == SRC_LEVEL_ROWS) return launch_predict_image<SRC_LEVEL_ROWS>(prm, dyn, s);
        if (src == SRC_LEVEL_PLANES) return launch_predict_image<SRC_LEVEL_PLANES>(prm, dyn, s);
        return launch_predict_image<SRC_PACKED>(prm, dyn, s);
    }
    if (stages & SSDHOT_STAGE_SCORES) {
        if (src == SRC_LEVEL_ROWS) score_kernel<6, SRC_LEVEL_ROWS><<<B * SCS, ST, 0, s>>>(prm);
        else if (src == SRC_LEVEL_PLANES) score_kernel<6, SRC_LEVEL_PLANES><<<B * SCS, ST, 0, s>>>(prm);
        else if (approx) score_kernel<6, SRC_PACKED><<<B * SCS, ST, 0, s>>>(prm);
        else score_kernel<0, SRC_PACKED><<<B * SCS, ST, 0, s>>>(prm);
        SSDHOT_CHECK_LAUNCH();
    }
    if (!(stages & SSDHOT_STAGE_NMS)) return SSDHOT_OK;
    if (src == SRC_LEVEL_ROWS) return launch_nms_image<true, SRC_LEVEL_ROWS>(prm, dyn, s);
    if (src == SRC_LEVEL_PLANES) return launch_nms_image<true, SRC_LEVEL_PLANES>(prm, dyn, s);
    return approx ? launch_nms_image<true, SRC_PACKED>(prm, dyn, s) : launch_nms_image<false, SRC_PACKED>(prm, dyn, s);
}

extern "C" int ssdhot_predict_stages(const float* priors_cxcywh, int P, const float* loc_all, const float* conf_all,
                                     int B, int C, float score_thresh, float nms_thresh, int max_per_img,
                                     int class_agnostic, int metric, float var_center, float var_size,
                                     float img_w, float img_h,
                                     int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                                     int32_t* out_count, void* work, int stages, void* share, ssdhot_stream_t stream) {
    if (!loc_all || !conf_all) return SSDHOT_ERR_NULL;
    if (!al16(loc_all) || (reinterpret_cast<uintptr_t>(conf_all) & 7u)) return SSDHOT_ERR_ALIGN;
    PredictParams prm = {};
    prm.pri = priors_cxcywh; prm.P = P; prm.loc_all = loc_all; prm.conf_all = conf_all; prm.B = B; prm.C = C;
    prm.score_thresh = score_thresh; prm.nms_thresh = nms_thresh; prm.max_keep = max_per_img;
    prm.vc = var_center; prm.vs = var_size; prm.img_w = img_w; prm.img_h = img_h;
    prm.out_labels = out_labels; prm.out_scores = out_scores; prm.out_boxes = out_boxes; prm.out_cand = out_cand;
    prm.out_count = out_count;
    return predict_launch(prm, SRC_PACKED, class_agnostic, metric, work, stages, share, (cudaStream_t)stream);
}

// predict straight from the six head outputs of each branch (SSD300 layout, C == 6): see heads.cuh.
extern "C" int ssdhot_predict_heads(const float* priors_cxcywh, const float* const* loc_heads_host,
                                    const float* const* conf_heads_host, int head_layout, int B, int C,
                                    float score_thresh, float nms_thresh, int max_per_img,
                                    int class_agnostic, int metric, float var_center, float var_size,
                                    float img_w, float img_h,
                                    int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                                    int32_t* out_count, void* work, int stages, void* share, ssdhot_stream_t stream) {
    if (!loc_heads_host || !conf_heads_host) return SSDHOT_ERR_NULL;
    if (head_layout != SSDHOT_HEADS_NHWC && head_layout != SSDHOT_HEADS_NCHW) return SSDHOT_ERR_VALUE;
    if (C != 6) return SSDHOT_ERR_SHAPE;                       // other class counts: ssdhot_pack_heads + ssdhot_predict
    PredictParams prm = {};
    for (int l = 0; l < kHeadLevels; ++l) {
        if (!loc_heads_host[l] || !conf_heads_host[l]) return SSDHOT_ERR_NULL;
        if (!al16(loc_heads_host[l]) || !al16(conf_heads_host[l])) return SSDHOT_ERR_ALIGN;
        prm.loc_h.base[l] = loc_heads_host[l];
        prm.conf_h.base[l] = conf_heads_host[l];
    }
    prm.pri = priors_cxcywh; prm.P = 8732; prm.B = B; prm.C = C;
    prm.score_thresh = score_thresh; prm.nms_thresh = nms_thresh; prm.max_keep = max_per_img;
    prm.vc = var_center; prm.vs = var_size; prm.img_w = img_w; prm.img_h = img_h;
    prm.out_labels = out_labels; prm.out_scores = out_scores; prm.out_boxes = out_boxes; prm.out_cand = out_cand;
    prm.out_count = out_count;
    return predict_launch(prm, head_layout == SSDHOT_HEADS_NHWC ? SRC_LEVEL_ROWS : SRC_LEVEL_PLANES, class_agnostic, metric, work,
                          stages, share, (cudaStream_t)stream);
}
