// infer_path.cu -- decode, score threshold, ranking and greedy DIoU/CIoU NMS
// (SURVEY.md section 8a rows a6-a8).
//
// Unit of work = one IMAGE: one CTA of 1024 threads walks the image's candidates of ALL classes in
// one global score order and applies class-aware greedy NMS, so the walk can stop as soon as
// max_per_img boxes survive.  That is exactly the reference's result -- per-class greedy NMS, then
// a global score sort and `keep[:max_per_img]` (SSD_from_scratch.py:439-465): a candidate's fate
// depends only on higher-scored candidates of its own class, all of which precede it in the
// global order, and survivors are produced in global score order -- but it never ranks or tests
// the thousands of low-score candidates that cannot reach the output.
//
// The image's scores live in shared memory as a dense array of order-preserving keys indexed by
// candidate id prior*(C-1)+class (0 = not a candidate), so thresholding needs no compaction and
// equal scores are ordered by candidate id for free.  Candidates are consumed best-first in
// chunks: a 4-pass radix select pulls the next CHUNK highest keys, a bitonic network sorts them,
// their boxes are decoded, and greedy NMS walks the chunk in tiles of 64 -- every (kept box |
// earlier tile member) x (tile member) predicate is evaluated in parallel into 64-bit suppression
// masks (ballot), then one thread resolves the tile serially with bit operations.
// When the dense array does not fit in shared memory (many classes) the same unit runs per
// (image, class) and a merge kernel combines the per-class survivor lists.
#include <map>
#include <mutex>

#include "boxmath.cuh"

namespace ssdhot {

constexpr int CHUNK = 256;     // candidates ranked per round
constexpr int TILE = 64;
constexpr float kFilterSlack = 0.999f;   // see suppresses()
constexpr int HBINS = 4096;    // score histogram: 16 octaves below 1.0 x 256 mantissa steps
constexpr int HBASE = (127 - 16) << 8;

struct UnitShared {
    unsigned hist[256];
    unsigned long long rowmask[TILE];   // bit j of rowmask[i]: tile member i suppresses tile member j (j > i)
    unsigned long long cmask[32];       // tile members per suppression group
    unsigned long long keepbits;
    int ngroup[32];                     // survivors per suppression group
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
template <int METRIC>
__device__ __forceinline__ bool suppresses(const BoxC& S, const BoxC& c, float thr, float thr_lo) {
    const float w = fmaxf(fsub(fminf(S.x2, c.x2), fmaxf(S.x1, c.x1)), 0.0f);
    const float h = fmaxf(fsub(fminf(S.y2, c.y2), fmaxf(S.y1, c.y1)), 0.0f);
    const float inter = fmul(w, h);
    const float uni = fsub(fadd(S.area, c.area), inter);
    if (inter < fmul(thr_lo, uni)) return false;
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
// from the cut upwards are zeroed: they are consumed by the gather that follows.  1024 threads,
// four bins each, thread 0 owns the top four.
__device__ __forceinline__ void hist_cut(unsigned* hist16, int cap, UnitShared& us) {
    __shared__ unsigned long long best[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w0 = (1023 - tid) * 2;
    const unsigned lo = hist16[w0], hi = hist16[w0 + 1];
    const int c[4] = {(int)(hi >> 16), (int)(hi & 0xffffu), (int)(lo >> 16), (int)(lo & 0xffffu)};   // top first
    const int mine = c[0] + c[1] + c[2] + c[3];
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
    const int top_bin = (1023 - tid) * 4 + 3;
    int proposal = -1, prop_count = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        run += c[j];
        if (run <= cap) { proposal = top_bin - j; prop_count = run; }
    }
    unsigned long long packed = proposal >= 0 ? (((unsigned long long)(unsigned)proposal << 32) | (unsigned)prop_count) : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(FULL, packed, o);
        packed = y < packed ? y : packed;
    }
    if (lane == 0) best[warp] = packed;
    __syncthreads();
    if (tid < 32) {
        unsigned long long v = best[tid];
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
        unsigned nlo = lo, nhi = hi;
        const int b0 = (1023 - tid) * 4;
        if (b0 + 0 >= cut) nlo &= 0xffff0000u;
        if (b0 + 1 >= cut) nlo &= 0x0000ffffu;
        if (b0 + 2 >= cut) nhi &= 0xffff0000u;
        if (b0 + 3 >= cut) nhi &= 0x0000ffffu;
        hist16[w0] = nlo; hist16[w0 + 1] = nhi;
    }
}

// Radix select over the non-zero entries of dense[0..n): the K-th largest key.  Returns the key,
// how many entries equal to it are needed (`need`) and how many exist (`eq`).
template <int NT>
__device__ __forceinline__ void select_kth(const unsigned* dense, int n, unsigned K, UnitShared& us,
                                           unsigned& thr, unsigned& need, unsigned& eq) {
    const int tid = threadIdx.x;
    unsigned prefix = 0u, remaining = K, count_eq = 0u;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < 256; i += NT) us.hist[i] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const unsigned k = dense[i];
            if (k != 0u && (k & himask) == prefix) atomicAdd(&us.hist[(k >> shift) & 255u], 1u);
        }
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

// In-place bitonic sort (descending) of n_pad (power of two <= CHUNK) 64-bit keys in shared memory.
__device__ __forceinline__ void bitonic_desc(unsigned long long* keys, int n_pad) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            if (tid < n_pad) {
                const int partner = tid ^ j;
                if (partner > tid) {
                    const unsigned long long a = keys[tid], b = keys[partner];
                    const bool desc = (tid & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[tid] = b; keys[partner] = a; }
                }
            }
        }
    }
    __syncthreads();
}

// ---- the unit ------------------------------------------------------------------------------------
// dense:   [n] keys (0 = absent), consumed (zeroed) as candidates are ranked
// hist16:  packed 16-bit score histogram of the candidates (HIST only)
// Fetch:   BoxC operator()(unsigned idx)      -- pixel box + constants of candidate idx
// Group:   int operator()(unsigned idx)       -- suppression group (class) of candidate idx; only
//                                                members of the same group suppress each other
// Emit:    void operator()(int pos, unsigned long long key, unsigned idx, const BoxC&)
// kept:    the surviving boxes in output order, capacity max_keep (shared or global memory)
// kidx:    [n_groups][max_keep] per-group lists of positions in `kept` (GROUPS only)
// returns the number of survivors (valid in every thread)
struct UnitBuffers {
    unsigned* dense; unsigned* hist16; unsigned long long* ckey; BoxC* cbox; unsigned char* cgroup;
    BoxC* kept; unsigned short* kidx;
};

template <int METRIC, int NT, bool GROUPS, bool HIST, typename Fetch, typename Group, typename Emit>
__device__ int nms_unit(const UnitBuffers buf, int n, int n_cand, int n_groups, int max_keep, float thr, UnitShared& us,
                        Fetch fetch, Group group_of, Emit emit) {
    constexpr int NW = NT / 32;
    constexpr int SUB = NT / TILE;                 // threads cooperating on one tile member (16 or 8)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned* dense = buf.dense;
    unsigned long long* ckey = buf.ckey;
    BoxC* cbox = buf.cbox;
    BoxC* kept = buf.kept;
    const float thr_lo = fmul(thr, kFilterSlack);
    int kept_n = 0;
    int remaining = n_cand;
    bool use_hist = HIST && NT == 1024 && n < 65536;       // 16-bit bin counters, one thread per four bins
    if (tid < 32) us.ngroup[tid] = 0;
    while (remaining > 0 && kept_n < max_keep) {
        // ---- pull the next best candidates ---------------------------------------------------
        int K = remaining < CHUNK ? remaining : CHUNK;
        unsigned tkey = 1u, need = 0u, eq = 0u;
        bool all = remaining <= CHUNK;
        if (!all && use_hist) {
            hist_cut(buf.hist16, CHUNK, us);
            if (us.cut_count > 0) { K = us.cut_count; tkey = bin_floor_key(us.cut_bin); need = eq = 0u; all = true; }
            else use_hist = false;                 // a single bin holds more than a chunk: exact select from here on
            __syncthreads();
        }
        if (!all) select_kth<NT>(dense, n, (unsigned)K, us, tkey, need, eq);
        if (tid == 0) us.counter = 0;
        __syncthreads();
        if (all || need == eq) {
            for (int i = tid; i < n; i += NT) {
                const unsigned k = dense[i];
                if (k != 0u && k >= tkey) {
                    const int pos = atomicAdd(&us.counter, 1);
                    ckey[pos] = ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
                    dense[i] = 0u;
                }
            }
        } else {
            // more entries equal to the threshold key than needed: take the lowest ids first
            int taken_eq = 0;
            for (int base = 0; base < n; base += NT) {
                const int i = base + tid;
                const unsigned k = i < n ? dense[i] : 0u;
                const bool is_eq = k != 0u && k == tkey;
                const unsigned bal = __ballot_sync(FULL, is_eq);
                __syncthreads();
                if (lane == 0) us.iscratch[warp] = __popc(bal);
                __syncthreads();
                int before = taken_eq, tot = 0;
                for (int w = 0; w < NW; ++w) { if (w < warp) before += us.iscratch[w]; tot += us.iscratch[w]; }
                const int my_rank = before + __popc(bal & ((1u << lane) - 1u));
                if (k != 0u && (k > tkey || (is_eq && (unsigned)my_rank < need))) {
                    const int pos = atomicAdd(&us.counter, 1);
                    ckey[pos] = ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
                    dense[i] = 0u;
                }
                taken_eq += tot;
            }
        }
        int n_pad = 64;
        while (n_pad < K) n_pad <<= 1;
        __syncthreads();
        for (int i = K + tid; i < n_pad; i += NT) ckey[i] = 0ull;
        bitonic_desc(ckey, n_pad);
        for (int i = tid; i < K; i += NT) {
            const unsigned idx = 0xffffffffu - (unsigned)(ckey[i] & 0xffffffffull);
            cbox[i] = fetch(idx);
            buf.cgroup[i] = GROUPS ? (unsigned char)group_of(idx) : (unsigned char)0;
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
                    const int ng = us.ngroup[g];
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
                if (GROUPS) {
                    const int g_lo = in_lo ? (int)buf.cgroup[t0 + lane] : -1, g_hi = in_hi ? (int)buf.cgroup[t0 + lane + 32] : -1;
                    for (int g = 0; g < n_groups; ++g) {
                        const unsigned a = __ballot_sync(FULL, g_lo == g), b2 = __ballot_sync(FULL, g_hi == g);
                        if (lane == 0) us.cmask[g] = ((unsigned long long)b2 << 32) | a;
                    }
                }
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
            if (tid < m && ((keepb >> tid) & 1ull)) {
                const unsigned long long lower = keepb & ((1ull << tid) - 1ull);
                const int pos = kept_n + __popcll(lower);
                const BoxC bx = cbox[t0 + tid];
                kept[pos] = bx;
                if (GROUPS) {
                    const int g = (int)buf.cgroup[t0 + tid];
                    buf.kidx[(size_t)g * max_keep + us.ngroup[g] + __popcll(lower & us.cmask[g])] = (unsigned short)pos;
                }
                const unsigned long long key = ckey[t0 + tid];
                emit(pos, key, 0xffffffffu - (unsigned)(key & 0xffffffffull), bx);
            }
            if (tid >= NT - TILE) { us.rowmask[tid - (NT - TILE)] = 0ull; us.suppf[tid - (NT - TILE)] = 0; }
            kept_n += __popcll(keepb);
            __syncthreads();
            if (GROUPS && tid < n_groups) us.ngroup[tid] += __popcll(keepb & us.cmask[tid]);
            __syncthreads();
        }
        remaining -= K;
    }
    return kept_n;
}

// shared-memory carve-up of one unit: kept | cbox | ckey | dense | hist16 | kidx | cgroup
__host__ __device__ inline size_t unit_smem_bytes(long long n, int max_keep, int n_groups, bool hist) {
    return (size_t)max_keep * sizeof(BoxC) + (size_t)CHUNK * (sizeof(BoxC) + 8) + (size_t)n * 4 +
           (hist ? (size_t)HBINS * 2 : 0) + (size_t)n_groups * max_keep * 2 + (size_t)CHUNK + 32;
}
__device__ __forceinline__ UnitBuffers carve_unit(unsigned char* dyn, long long n, int max_keep, int n_groups, bool hist) {
    UnitBuffers b;
    b.kept = reinterpret_cast<BoxC*>(dyn);
    b.cbox = b.kept + max_keep;
    b.ckey = reinterpret_cast<unsigned long long*>(b.cbox + CHUNK);
    b.dense = reinterpret_cast<unsigned*>(b.ckey + CHUNK);
    b.hist16 = b.dense + n;
    b.kidx = reinterpret_cast<unsigned short*>(b.hist16 + (hist ? HBINS / 2 : 0));
    b.cgroup = reinterpret_cast<unsigned char*>(b.kidx + (size_t)n_groups * max_keep);
    return b;
}

// ---- predict -------------------------------------------------------------------------------------
struct PredictParams {
    const float* pri; int P; const float* loc_all; const float* conf_all; int B, C;
    float score_thresh, nms_thresh; int max_keep; float vc, vs, img_w, img_h;
    // final outputs (per-image kernel)
    int64_t* out_labels; float* out_scores; float* out_boxes; int32_t* out_cand; int32_t* out_count;
    // per-class lists (fallback path)
    unsigned long long* list_key;   // [B*units][max_keep]
    float4* list_box;               // [B*units][max_keep]
    int* list_count;                // [B*units]
};

// exp(x_i - max) of one row and their sum in eager torch-CUDA order (persistent warp softmax: one
// element per lane, lanes = next_pow2(C), butterfly adds over xor offsets lanes/2 .. 1).  CT == 6 is
// the reference's class count; CT == 0 handles any C <= 32 through `e` in local memory.
template <int CT>
__device__ __forceinline__ float row_exps(const float* __restrict__ row, int C, float* e) {
    if (CT == 6) {
        const float2 a = ldg2(row), b = ldg2(row + 2), c = ldg2(row + 4);
        const float mx = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)), fmaxf(c.x, c.y));
        e[0] = expf(fsub(a.x, mx)); e[1] = expf(fsub(a.y, mx)); e[2] = expf(fsub(b.x, mx));
        e[3] = expf(fsub(b.y, mx)); e[4] = expf(fsub(c.x, mx)); e[5] = expf(fsub(c.y, mx));
        return fadd(fadd(fadd(e[0], e[4]), e[2]), fadd(fadd(e[1], e[5]), e[3]));
    } else {
        int lanes = 1;
        while (lanes < C) lanes <<= 1;
        float mx = __ldg(row);
        for (int i = 1; i < C; ++i) mx = fmaxf(mx, __ldg(row + i));
        float part[32];
        for (int l = 0; l < 32; ++l) part[l] = 0.0f;
        for (int i = 0; i < C; ++i) { e[i] = expf(fsub(__ldg(row + i), mx)); part[i] = e[i]; }
        for (int off = lanes >> 1; off > 0; off >>= 1)
            for (int l = 0; l < off; ++l) part[l] = fadd(part[l], part[l + off]);
        return part[0];
    }
}

constexpr int PT = 1024;   // threads of the per-image kernel
constexpr int UT = 512;    // threads of the per-(image, class) fallback and of stand-alone NMS

// One CTA per image, all classes in one score-ordered stream (class-aware unless AGN).
template <int METRIC, bool AGN, int CT>
__global__ void __launch_bounds__(PT) predict_image_kernel(const PredictParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ UnitShared us;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int n_fg = prm.C - 1, P = prm.P;
    const int n = P * n_fg;
    const int n_groups = AGN ? 1 : n_fg;
    const UnitBuffers buf = carve_unit(dyn, n, prm.max_keep, AGN ? 0 : n_groups, true);
    for (int i = tid; i < HBINS / 2; i += PT) buf.hist16[i] = 0u;
    __syncthreads();

    // A score can only pass `s > thresh` if e_k > thr_pre * sum, thr_pre = thresh * (1 - 1e-5): the
    // margin is two orders of magnitude above the rounding of the product and of the division, so
    // rows / classes below it skip the IEEE division without changing any decision.
    const float thr_pre = fmul(prm.score_thresh, 0.99999f);
    int mine = 0;
    const float* conf_b = prm.conf_all + (long long)b * P * prm.C;
    for (int p = tid; p < P; p += PT) {
        float e[CT > 0 ? CT : 32];
        const float sum = row_exps<CT>(conf_b + (long long)p * prm.C, prm.C, e);
        const float gate = fmul(thr_pre, sum);
        const int nf = CT > 0 ? CT - 1 : n_fg;
#pragma unroll
        for (int k = 0; k < nf; ++k) {
            unsigned key = 0u;
            if (e[k + 1] > gate) {
                const float s = fdiv(e[k + 1], sum);       // softmax(conf)[..., 1:]  (SFS:388)
                if (s > prm.score_thresh) {                // strict (SFS:402)
                    key = ord_encode(s);
                    hist_add(buf.hist16, key);
                    mine += 1;
                }
            }
            buf.dense[p * n_fg + k] = key;
        }
    }
    const int n_cand = block_sum<int>(mine, us.iscratch);
    __syncthreads();

    const float* loc_b = prm.loc_all + 4ll * b * P;
    const bool want_atan = METRIC == SSDHOT_METRIC_CIOU;
    auto fetch = [&](unsigned idx) -> BoxC {
        const unsigned p = idx / (unsigned)n_fg;
        const float4 box = decode_box(ldg4(loc_b + 4ll * p), ldg4(prm.pri + 4ll * p), prm.vc, prm.vs);
        const float4 px = to_pixel_xyxy(box, prm.img_w, prm.img_h);
        return box_consts(px.x, px.y, px.z, px.w, want_atan);
    };
    auto group_of = [&](unsigned idx) -> int { return (int)(idx % (unsigned)n_fg); };
    const long long o = (long long)b * prm.max_keep;
    auto emit = [&](int pos, unsigned long long key, unsigned idx, const BoxC& bx) {
        prm.out_labels[o + pos] = (int64_t)(idx % (unsigned)n_fg);
        prm.out_scores[o + pos] = ord_decode((unsigned)(key >> 32));
        reinterpret_cast<float4*>(prm.out_boxes)[o + pos] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
        if (prm.out_cand) prm.out_cand[o + pos] = (int32_t)idx;
    };
    const int kept_n = nms_unit<METRIC, PT, !AGN, true>(buf, n, n_cand, n_groups, prm.max_keep, prm.nms_thresh, us, fetch, group_of, emit);
    if (tid == 0) prm.out_count[b] = kept_n;
}

// Fallback for class counts whose dense score array does not fit in shared memory: one CTA per
// (image, class) (or per image when class-agnostic and it fits), lists merged afterwards.
template <int METRIC>
__global__ void __launch_bounds__(UT) predict_class_kernel(const PredictParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ UnitShared us;
    const int tid = threadIdx.x;
    const int n_fg = prm.C - 1;
    const int b = blockIdx.x / n_fg, c = blockIdx.x % n_fg;
    const int P = prm.P;
    const UnitBuffers buf = carve_unit(dyn, P, prm.max_keep, 0, false);

    int mine = 0;
    const float* conf_b = prm.conf_all + (long long)b * P * prm.C;
    for (int p = tid; p < P; p += UT) {
        const float* row = conf_b + (long long)p * prm.C;
        float sum;
        float ec;
        if (prm.C <= 32) {
            float e[32];
            sum = row_exps<0>(row, prm.C, e);
            ec = e[c + 1];
        } else {
            // C > 32: eager torch-CUDA keeps ceil(C/32) elements per lane, summed in order, then the butterfly
            float mx = __ldg(row);
            for (int i = 1; i < prm.C; ++i) mx = fmaxf(mx, __ldg(row + i));
            float part[32];
            for (int l = 0; l < 32; ++l) part[l] = 0.0f;
            for (int i = 0; i < prm.C; ++i) part[i & 31] = fadd(part[i & 31], expf(fsub(__ldg(row + i), mx)));
            for (int off = 16; off > 0; off >>= 1)
                for (int l = 0; l < off; ++l) part[l] = fadd(part[l], part[l + off]);
            sum = part[0];
            ec = expf(fsub(__ldg(row + c + 1), mx));
        }
        const float s = fdiv(ec, sum);
        const bool on = s > prm.score_thresh;
        buf.dense[p] = on ? ord_encode(s) : 0u;
        mine += on ? 1 : 0;
    }
    const int n_cand = block_sum<int>(mine, us.iscratch);
    __syncthreads();

    const float* loc_b = prm.loc_all + 4ll * b * P;
    const bool want_atan = METRIC == SSDHOT_METRIC_CIOU;
    auto fetch = [&](unsigned idx) -> BoxC {
        const float4 box = decode_box(ldg4(loc_b + 4ll * idx), ldg4(prm.pri + 4ll * idx), prm.vc, prm.vs);
        const float4 px = to_pixel_xyxy(box, prm.img_w, prm.img_h);
        return box_consts(px.x, px.y, px.z, px.w, want_atan);
    };
    auto group_of = [&](unsigned) -> int { return 0; };
    unsigned long long* out_key = prm.list_key + (long long)blockIdx.x * prm.max_keep;
    float4* out_box = prm.list_box + (long long)blockIdx.x * prm.max_keep;
    auto emit = [&](int pos, unsigned long long key, unsigned idx, const BoxC& bx) {
        // re-key by the flat candidate id prior*(C-1)+class so that lists of different classes merge
        const unsigned flat = idx * (unsigned)n_fg + (unsigned)c;
        out_key[pos] = (key & 0xffffffff00000000ull) | (unsigned long long)(0xffffffffu - flat);
        out_box[pos] = make_float4(bx.x1, bx.y1, bx.x2, bx.y2);
    };
    const int kept_n = nms_unit<METRIC, UT, false, false>(buf, P, n_cand, 1, prm.max_keep, prm.nms_thresh, us, fetch, group_of, emit);
    if (tid == 0) prm.list_count[blockIdx.x] = kept_n;
}

// Merge the per-class survivor lists of one image (each already score-descending) and keep the
// max_keep best (SSD_from_scratch.py:461-474).  Rank of an element = its position in its own list
// + the number of strictly larger keys in every other list (keys are unique).
__global__ void __launch_bounds__(256) merge_lists_kernel(const unsigned long long* __restrict__ list_key,
                                                          const float4* __restrict__ list_box,
                                                          const int* __restrict__ list_count, int units, int max_keep, int n_fg,
                                                          int64_t* __restrict__ out_labels, float* __restrict__ out_scores,
                                                          float* __restrict__ out_boxes, int32_t* __restrict__ out_cand,
                                                          int32_t* __restrict__ out_count) {
    const int b = blockIdx.x;
    const unsigned long long* keys = list_key + (long long)b * units * max_keep;
    const float4* boxes = list_box + (long long)b * units * max_keep;
    const int* counts = list_count + (long long)b * units;
    int total = 0;
    for (int u = 0; u < units; ++u) total += counts[u];
    for (int e = threadIdx.x; e < units * max_keep; e += blockDim.x) {
        const int u = e / max_keep, i = e % max_keep;
        if (i >= counts[u]) continue;
        const unsigned long long key = keys[(long long)u * max_keep + i];
        int rank = i;
        for (int v = 0; v < units; ++v) {
            if (v == u) continue;
            int lo = 0, hi = counts[v];          // first position whose key is < key
            const unsigned long long* kv = keys + (long long)v * max_keep;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (kv[mid] > key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < max_keep) {
            const long long o = (long long)b * max_keep + rank;
            const unsigned flat = 0xffffffffu - (unsigned)(key & 0xffffffffull);
            out_labels[o] = (int64_t)(flat % (unsigned)n_fg);
            out_scores[o] = ord_decode((unsigned)(key >> 32));
            reinterpret_cast<float4*>(out_boxes)[o] = boxes[(long long)u * max_keep + i];
            if (out_cand) out_cand[o] = (int32_t)flat;
        }
    }
    if (threadIdx.x == 0) out_count[b] = total < max_keep ? total : max_keep;
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
    UnitBuffers buf;
    buf.cbox = reinterpret_cast<BoxC*>(dyn);
    buf.ckey = reinterpret_cast<unsigned long long*>(buf.cbox + CHUNK);
    buf.dense = reinterpret_cast<unsigned*>(buf.ckey + CHUNK);
    buf.cgroup = reinterpret_cast<unsigned char*>(buf.dense + n_max);
    buf.kept = kept_all + begin;            // survivors of a stand-alone call are unbounded: global memory
    buf.kidx = nullptr; buf.hist16 = nullptr;
    for (int i = tid; i < n; i += UT) {
        const unsigned k = ord_encode(__ldg(scores + begin + i));
        buf.dense[i] = k == 0u ? 1u : k;
    }
    __syncthreads();
    const bool want_atan = METRIC == SSDHOT_METRIC_CIOU;
    auto fetch = [&](unsigned idx) -> BoxC {
        const float4 bx = ldg4(boxes + 4ll * (begin + (long long)idx));
        return box_consts(bx.x, bx.y, bx.z, bx.w, want_atan);
    };
    auto group_of = [&](unsigned) -> int { return 0; };
    int64_t* out = keep + begin;
    auto emit = [&](int pos, unsigned long long, unsigned idx, const BoxC&) { out[pos] = (int64_t)idx; };
    const int cap = (max_keep > 0 && max_keep < n) ? max_keep : n;
    const int kept_n = nms_unit<METRIC, UT, false, false>(buf, n, n, 1, cap, thr, us, fetch, group_of, emit);
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
    // the opt-in is sticky per kernel, so it is raised once and never inside a graph capture
    if (bytes > kMaxDynSmem) return SSDHOT_ERR_SHAPE;
    static std::mutex mu;
    static std::map<const void*, size_t> configured;
    if (bytes > 40 * 1024) {      // static + dynamic shared memory share the 48 KB default limit
        std::lock_guard<std::mutex> lock(mu);
        size_t& have = configured[reinterpret_cast<const void*>(kern)];
        if (bytes > have) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
            if (e != cudaSuccess) return (int)e;
            have = kMaxDynSmem;
        }
    }
    return SSDHOT_OK;
}

template <int METRIC, bool AGN>
static int launch_predict_image(const PredictParams& prm, size_t dyn, cudaStream_t stream) {
    int rc;
    if (prm.C == 6) {
        if ((rc = set_smem(predict_image_kernel<METRIC, AGN, 6>, dyn))) return rc;
        predict_image_kernel<METRIC, AGN, 6><<<prm.B, PT, dyn, stream>>>(prm);
    } else {
        if ((rc = set_smem(predict_image_kernel<METRIC, AGN, 0>, dyn))) return rc;
        predict_image_kernel<METRIC, AGN, 0><<<prm.B, PT, dyn, stream>>>(prm);
    }
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

template <int METRIC>
static int launch_predict_class(const PredictParams& prm, size_t dyn, cudaStream_t stream) {
    int rc;
    if ((rc = set_smem(predict_class_kernel<METRIC>, dyn))) return rc;
    predict_class_kernel<METRIC><<<prm.B * (prm.C - 1), UT, dyn, stream>>>(prm);
    SSDHOT_CHECK_LAUNCH();
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

extern "C" unsigned long long ssdhot_predict_workspace_bytes(int B, int C, int max_per_img) {
    if (B <= 0 || C < 2 || max_per_img <= 0) return 64ull;
    const unsigned long long lists = (unsigned long long)B * (C - 1);
    return lists * max_per_img * (8ull + 16ull) + lists * 4ull + 256ull;
}

extern "C" int ssdhot_predict(const float* priors_cxcywh, int P, const float* loc_all, const float* conf_all,
                              int B, int C, float score_thresh, float nms_thresh, int max_per_img,
                              int class_agnostic, int metric, float var_center, float var_size,
                              float img_w, float img_h,
                              int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                              int32_t* out_count, void* work, ssdhot_stream_t stream) {
    if (!priors_cxcywh || !loc_all || !conf_all || !out_labels || !out_scores || !out_boxes || !out_count || !work)
        return SSDHOT_ERR_NULL;
    if (P <= 0 || B <= 0 || C < 2 || C > SSDHOT_MAX_CLASSES || max_per_img <= 0) return SSDHOT_ERR_SHAPE;
    // same validation as SSD_from_scratch.py:369-373
    if (!(score_thresh >= 0.0f && score_thresh < 1.0f) || !(nms_thresh > 0.0f && nms_thresh < 1.0f)) return SSDHOT_ERR_VALUE;
    if (metric != SSDHOT_METRIC_DIOU && metric != SSDHOT_METRIC_CIOU && metric != SSDHOT_METRIC_IOU) return SSDHOT_ERR_VALUE;
    if (!al16(priors_cxcywh) || !al16(loc_all) || !al16(out_boxes) || !al16(work) ||
        (reinterpret_cast<uintptr_t>(conf_all) & 7u)) return SSDHOT_ERR_ALIGN;
    PredictParams prm = {};
    prm.pri = priors_cxcywh; prm.P = P; prm.loc_all = loc_all; prm.conf_all = conf_all; prm.B = B; prm.C = C;
    prm.score_thresh = score_thresh; prm.nms_thresh = nms_thresh; prm.max_keep = max_per_img;
    prm.vc = var_center; prm.vs = var_size; prm.img_w = img_w; prm.img_h = img_h;
    prm.out_labels = out_labels; prm.out_scores = out_scores; prm.out_boxes = out_boxes; prm.out_cand = out_cand;
    prm.out_count = out_count;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    // preferred: one CTA per image, all classes in one score-ordered stream
    const size_t dyn_image = unit_smem_bytes((long long)P * (C - 1), max_per_img, class_agnostic ? 0 : C - 1, true);
    if (dyn_image <= kMaxDynSmem && C <= 32 && max_per_img < 65536) {
#define SSDHOT_DISPATCH(M) rc = class_agnostic ? launch_predict_image<M, true>(prm, dyn_image, s) : launch_predict_image<M, false>(prm, dyn_image, s)
        if (metric == SSDHOT_METRIC_DIOU) { SSDHOT_DISPATCH(SSDHOT_METRIC_DIOU); }
        else if (metric == SSDHOT_METRIC_CIOU) { SSDHOT_DISPATCH(SSDHOT_METRIC_CIOU); }
        else { SSDHOT_DISPATCH(SSDHOT_METRIC_IOU); }
#undef SSDHOT_DISPATCH
        return rc;
    }
    // fallback (many classes): one CTA per (image, class), then a merge; class-agnostic NMS over a
    // candidate set that does not fit in shared memory is not supported
    if (class_agnostic) return SSDHOT_ERR_SHAPE;
    const int units = C - 1;
    const long long lists = (long long)B * units;
    unsigned char* w = reinterpret_cast<unsigned char*>(work);
    prm.list_box = reinterpret_cast<float4*>(w); w += (size_t)lists * max_per_img * 16;
    prm.list_key = reinterpret_cast<unsigned long long*>(w); w += (size_t)lists * max_per_img * 8;
    prm.list_count = reinterpret_cast<int*>(w);
    const size_t dyn = unit_smem_bytes(P, max_per_img, 0, false);
    if (metric == SSDHOT_METRIC_DIOU) rc = launch_predict_class<SSDHOT_METRIC_DIOU>(prm, dyn, s);
    else if (metric == SSDHOT_METRIC_CIOU) rc = launch_predict_class<SSDHOT_METRIC_CIOU>(prm, dyn, s);
    else rc = launch_predict_class<SSDHOT_METRIC_IOU>(prm, dyn, s);
    if (rc) return rc;
    merge_lists_kernel<<<B, 256, 0, s>>>(prm.list_key, prm.list_box, prm.list_count, units, max_per_img, C - 1,
                                         out_labels, out_scores, out_boxes, out_cand, out_count);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}
