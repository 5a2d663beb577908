// train_path.cu -- match + encode + mined multibox loss.
//
// The SSD300 fast path is ONE kernel, train_image_kernel (second half of this file): box-centric matching over
// candidate rectangles beside the logit stream, approximate-then-exact hard-negative mining, one CTA per image.
// Given a share buffer (ssdhot.h: ssdhot_share_bytes) its logit stream also leaves the 16-bit row keys that
// predict_image_kernel of the same eval step needs, and raises one flag per image (template flag SHARE: 14 + 10 warps
// instead of 18 + 6).
// It is templated on the source of the head outputs (heads.cuh: packed tensors, channels_last rows or NCHW planes of the six
// per-level tensors -- ssdhot_multibox_loss_heads_fwd; loss_bwd_heads_kernel writes the gradients in the same layouts).
// The layout-agnostic path (any priors / class count / > 64 boxes per image, and encode_ssd's "every prior"
// outputs) is two kernels:
//  * match_kernel -- one thread-block CLUSTER per image (8 CTAs of 256 threads); each CTA owns a
//    contiguous slice of the priors, each thread keeps KP of them in registers, the image's ground
//    truth lives in shared memory with its per-box constants, and the column arg-max (best prior
//    per ground truth) is combined across the cluster through distributed shared memory.  On the
//    fused-loss path its only output is a 2-byte code per prior (0 = negative, 1 + matched box).
//  * loss_image_kernel -- one CTA of 768 threads per image streams the class logits once
//    (the HBM-bound part), forms softmax cross-entropy and smooth-L1 in registers and picks the
//    hard negatives with a CTA-local 4-pass radix select; no [B,P] float tensor touches HBM.
//
// The CIoU sweep runs in one of two forms:
//  * exact-everywhere (PRUNE = false): every (prior, box) pair is evaluated; needed only when the
//    caller wants the matched box / offsets of NEGATIVE priors too (the encode_ssd signature);
//  * pruned (PRUNE = true, the hot path): a pair is evaluated only if its IoU can matter.  Per box g
//    a lower bound cb0[g] of the column maximum is taken from ~30 "seed" priors (the ones whose
//    cell contains the box centre on each pyramid level); CIoU <= IoU always, so a pair with
//    intersection < lim[g] * union, lim[g] = 0.999 * min(cb0[g], iou_thresh), can neither be the
//    column champion (its CIoU < cb0[g] <= max) nor make its prior positive (CIoU < iou_thresh).
//    Positives, their matched boxes and the forced champions are therefore bit-identical to the
//    exact sweep while ~98 % of the pairs cost 13 instructions instead of ~75.
#include <type_traits>

#include "boxmath.cuh"
#include "heads.cuh"

namespace ssdhot {

constexpr int TT = 256;          // threads per CTA
constexpr int KP = 5;            // priors per thread
constexpr int SLOTS = TT * KP;   // prior slots per CTA
constexpr int MAX_CS = 8;        // portable cluster size
constexpr float kPruneSlack = 0.999f;


struct TrainParams {
    // priors
    const float* pri; const float* pri_xyxy; const float* pri_aux; int P;
    // ground truth
    const float* gt_boxes; const int64_t* gt_labels; const int32_t* gt_offsets;
    int B, max_gt; float norm_w, norm_h;
    float thresh, inv_vc, inv_vs;
    // head outputs
    const float* loc_all; const float* conf_all; int C;
    double ratio;
    // given targets (MODE_LOSS)
    const int64_t* in_cls; const uint8_t* in_pos;
    // match outputs
    float* loc_t; int loc_pos_only; int64_t* cls_t; uint8_t* pos_mask; int32_t* matched32;
    float* matched_box; int32_t* n_pos;
    // loss outputs
    int8_t* sel_cls; int16_t* matched16;
    float4* gt_rec;        // [B*max_gt][3]: (x1 y1 x2 y2), (area xc yc atan), (lim, label bits, -, -)
    uint16_t* code;        // [B,P] 0 = negative, 1 + matched box (fused path: match -> loss hand-off)
    double* img_part;      // [B][2] per-image (smooth-L1, CE) sums
    int32_t* flags;
    unsigned long long* timeline;   // debug: [B][16] %globaltimer stamps of the fused kernel's phases (or null)
    HeadView loc_h, conf_h;         // head sources (SRC_LEVEL_ROWS / SRC_LEVEL_PLANES) instead of loc_all / conf_all
    // key hand-off to predict_image_kernel of the same eval step (ssdhot_share_bytes; null = none): the stream also leaves the
    // 16-bit row keys of the image's logits and raises the image's flag
    int* share_flag; unsigned* share_keys;
};

// SSD300 pyramid (SSD_from_scratch.py:289-290): used only to pick seed priors, never for results
__constant__ int kLevelSide[6] = {38, 19, 10, 5, 3, 1};
__constant__ int kLevelShapes[6] = {4, 6, 6, 6, 4, 4};
__constant__ int kLevelOffset[6] = {0, 5776, 7942, 8542, 8692, 8728};
__constant__ int kSeedLevel[32] = {0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 1};
__constant__ int kSeedShape[32] = {0, 1, 2, 3, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 0, 1, 2, 3, 0, 0};

// ---------------------------------------------------------------------------------------------
// shared-memory carve-up (dynamic): ground-truth records, column keys, champions, forced table
// ---------------------------------------------------------------------------------------------
struct Smem {
    float4* gt_a;              // [G] x1 y1 x2 y2 (normalised)
    float4* gt_b;              // [G] area xc yc atan
    unsigned long long* col;   // [G] packed (ord(ciou) << 32 | ~prior) of this CTA's best prior per GT
    int* gt_label;             // [G]
    unsigned* champ;           // [G] cluster-wide best prior per GT
    float* lim;                // [G] prune bound (PRUNE) -- see header
    int* forced;               // [SLOTS] lowest GT index that forces this prior, or INT_MAX
};

__host__ __device__ inline size_t smem_bytes(int g_cap) {
    return (size_t)g_cap * (16 + 16 + 8 + 4 + 4 + 4) + (size_t)SLOTS * 4 + 64;
}

__device__ __forceinline__ Smem carve(unsigned char* base, int g_cap) {
    Smem s;
    s.gt_a = reinterpret_cast<float4*>(base); base += (size_t)g_cap * 16;
    s.gt_b = reinterpret_cast<float4*>(base); base += (size_t)g_cap * 16;
    s.col = reinterpret_cast<unsigned long long*>(base); base += (size_t)g_cap * 8;
    s.gt_label = reinterpret_cast<int*>(base); base += (size_t)g_cap * 4;
    s.champ = reinterpret_cast<unsigned*>(base); base += (size_t)g_cap * 4;
    s.lim = reinterpret_cast<float*>(base); base += (size_t)g_cap * 4;
    s.forced = reinterpret_cast<int*>(base);
    return s;
}

struct Static {                 // static shared state used by the cluster exchanges
    unsigned hist[2][256];
    unsigned total[256];
    double dscratch[32];
    int iscratch[32];
    int npos_cta;               // positives in this CTA (read remotely)
    int ties_cta;               // elements equal to the threshold in this CTA (read remotely)
    int first_nan;              // lowest GT index whose CIoU column is NaN, or INT_MAX
    int refill;                 // some column needs the exact sweep after all (PRUNE fallback)
    unsigned sel_digit, sel_need;
};

// ---------------------------------------------------------------------------------------------
// per-row softmax statistics in eager torch-CUDA order: returns (max, log(sum exp(x - max)))
// ---------------------------------------------------------------------------------------------
template <int CT>
__device__ __forceinline__ void row_lse(const float* __restrict__ row, int C, float& mx, float& lg) {
    if (CT == 6) {
        const float2 a = ldg2(row), b = ldg2(row + 2), c = ldg2(row + 4);
        mx = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)), fmaxf(c.x, c.y));
        const float e0 = expf(fsub(a.x, mx)), e1 = expf(fsub(a.y, mx)), e2 = expf(fsub(b.x, mx));
        const float e3 = expf(fsub(b.y, mx)), e4 = expf(fsub(c.x, mx)), e5 = expf(fsub(c.y, mx));
        const float s = fadd(fadd(fadd(e0, e4), e2), fadd(fadd(e1, e5), e3));   // 8-lane butterfly
        lg = logf(s);
    } else {
        // generic C: lane l of a 32-wide (or next_pow2(C)-wide) warp sums elements l, l+W, ...
        int lanes = 1;
        while (lanes < C && lanes < 32) lanes <<= 1;
        mx = __ldg(row);
        for (int i = 1; i < C; ++i) mx = fmaxf(mx, __ldg(row + i));
        float part[32];
        for (int l = 0; l < 32; ++l) part[l] = 0.0f;
        for (int i = 0; i < C; ++i) {
            const int l = i & (lanes - 1);
            part[l] = fadd(part[l], expf(fsub(__ldg(row + i), mx)));
        }
        for (int off = lanes >> 1; off > 0; off >>= 1)
            for (int l = 0; l < off; ++l) part[l] = fadd(part[l], part[l + off]);
        lg = logf(part[0]);
    }
}

__device__ __forceinline__ BoxC load_prior(const float* __restrict__ xyxy, const float* __restrict__ aux, int p) {
    const float4 a = ldg4(xyxy + 4ll * p), x = ldg4(aux + 4ll * p);
    BoxC r;
    r.x1 = a.x; r.y1 = a.y; r.x2 = a.z; r.y2 = a.w; r.area = x.x; r.xc = x.y; r.yc = x.z; r.at = x.w;
    return r;
}

__device__ __forceinline__ BoxC gt_box(const float4 ga, const float4 gb) {
    BoxC g;
    g.x1 = ga.x; g.y1 = ga.y; g.x2 = ga.z; g.y2 = ga.w; g.area = gb.x; g.xc = gb.y; g.yc = gb.z; g.at = gb.w;
    return g;
}

// one column of the exact sweep for this thread's priors (used by PRUNE = false and by the fallback)
struct ColBest { float v; unsigned p; };

__device__ __forceinline__ void publish_column(const ColBest cb, unsigned long long* slot) {
    const unsigned enc = (cb.p == 0xffffffffu) ? 0u : ord_encode(cb.v);
    const unsigned wmax = __reduce_max_sync(FULL, enc);
    const unsigned wmin = __reduce_min_sync(FULL, enc == wmax ? cb.p : 0xffffffffu);
    if ((threadIdx.x & 31) == 0 && wmin != 0xffffffffu) {
        const unsigned long long key = ((unsigned long long)wmax << 32) | (unsigned long long)(0xffffffffu - wmin);
        if (key > *slot) atomicMax(slot, key);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool PRUNE>
__global__ void __launch_bounds__(TT, PRUNE ? 4 : 2) match_kernel(const TrainParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ Static st;

    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / cs;
    const int tid = threadIdx.x;
    const int P = prm.P;
    const int chunk = (P + cs - 1) / cs;
    const int p0 = rank * chunk;
    const int p1 = min(P, p0 + chunk);

    const int g_begin = prm.gt_offsets[b];
    int G = prm.gt_offsets[b + 1] - g_begin;
    if (G > prm.max_gt) {
        if (tid == 0 && rank == 0 && prm.flags) atomicOr(prm.flags, 1);
        G = prm.max_gt;
    }
    Smem sm = carve(dyn, prm.max_gt > 0 ? prm.max_gt : 1);

    // ---- per-thread prior state -------------------------------------------------------------
    int best_g[KP];
    float best_v[KP];
    bool pos[KP];
    int cls[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) { best_g[k] = 0; best_v[k] = -INFINITY; pos[k] = false; cls[k] = 0; }

    {
        // ---- stage ground truth ------------------------------------------------------------
        if (tid == 0) { st.first_nan = INT_MAX; st.refill = 0; }
        for (int i = tid; i < SLOTS; i += TT) sm.forced[i] = INT_MAX;
        __syncthreads();
        for (int g = tid; g < G; g += TT) {
            const float4* rec = prm.gt_rec + 3ll * ((long long)b * prm.max_gt + g);     // gt_prepare_kernel
            const float4 ra = __ldg(rec), rb = __ldg(rec + 1), rc = __ldg(rec + 2);
            sm.gt_a[g] = ra;
            sm.gt_b[g] = rb;
            sm.lim[g] = rc.x;
            sm.gt_label[g] = __float_as_int(rc.y);
            sm.col[g] = 0ull;
            if (rb.w != rb.w) atomicMin(&st.first_nan, g);
        }
        __syncthreads();

        // ---- CIoU sweep: row arg-max in registers, column arg-max per warp -> shared ---------
        float4 pb[KP];      // x1 y1 x2 y2 of this thread's priors
        float pa[KP];       // their areas
        bool valid[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int p = p0 + k * TT + tid;
            valid[k] = p < p1;
            const int q = valid[k] ? p : p0;
            pb[k] = ldg4(prm.pri_xyxy + 4ll * q);
            pa[k] = __ldg(prm.pri_aux + 4ll * q);
        }
        for (int g = 0; g < G; ++g) {
            const float4 ga = sm.gt_a[g], gb = sm.gt_b[g];
            if (gb.w != gb.w) continue;            // NaN column (degenerate box): handled below
            const BoxC gc = gt_box(ga, gb);
            const float lim = PRUNE ? sm.lim[g] : 0.0f;
            ColBest cb = {-INFINITY, 0xffffffffu};
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                bool go = true;
                if (PRUNE) {
                    // no vertical overlap anywhere in the warp: intersection 0 < lim * union for all lanes
                    const float hraw = fsub(fminf(pb[k].w, gc.y2), fmaxf(pb[k].y, gc.y1));
                    if (!__any_sync(FULL, hraw > 0.0f)) continue;
                    const float w = fmaxf(fsub(fminf(pb[k].z, gc.x2), fmaxf(pb[k].x, gc.x1)), 0.0f);
                    const float h = fmaxf(hraw, 0.0f);
                    const float inter = fmul(w, h);
                    const float uni = fsub(fadd(pa[k], gc.area), inter);
                    go = !(inter < fmul(lim, uni));          // NaN-safe: anything odd takes the exact path
                }
                if (!PRUNE || __any_sync(FULL, go)) {
                    if (go) {
                        const int q = valid[k] ? p0 + k * TT + tid : p0;
                        const float4 x = ldg4(prm.pri_aux + 4ll * q);
                        BoxC pr;
                        pr.x1 = pb[k].x; pr.y1 = pb[k].y; pr.x2 = pb[k].z; pr.y2 = pb[k].w;
                        pr.area = x.x; pr.xc = x.y; pr.yc = x.z; pr.at = x.w;
                        const float v = pair_ciou(pr, gc);
                        if (v > best_v[k]) { best_v[k] = v; best_g[k] = g; }
                        if (valid[k] && (v > cb.v || cb.p == 0xffffffffu)) { cb.v = v; cb.p = (unsigned)(p0 + k * TT + tid); }   // (a column of -inf still has a champion)
                    }
                }
            }
            if (!PRUNE || __any_sync(FULL, cb.p != 0xffffffffu)) publish_column(cb, &sm.col[g]);
        }
        // ---- cluster-wide champions --------------------------------------------------------
        cluster.sync();
        for (int g = tid; g < G; g += TT) {
            unsigned long long key = 0ull;
            for (int r = 0; r < cs; ++r) {
                const unsigned long long kr = cluster.map_shared_rank(sm.col, r)[g];
                key = kr > key ? kr : key;
            }
            const float4 gb = sm.gt_b[g];
            const bool nan_col = gb.w != gb.w;
            // PRUNE fallback: a column whose pruned maximum is not positive (box outside every prior,
            // or seeds that bound nothing) is redone exactly
            if (PRUNE && !nan_col && !((unsigned)(key >> 32) > ord_encode(0.0f))) st.refill = 1;
            const unsigned champ = nan_col ? 0u : (0xffffffffu - (unsigned)(key & 0xffffffffull));
            sm.champ[g] = champ;
        }
        __syncthreads();
        if (PRUNE) {
            // every CTA of the cluster reads the same keys, so `refill` is cluster-uniform
            if (st.refill) {
                cluster.sync();                       // all peers are done reading sm.col
                for (int g = 0; g < G; ++g) {
                    const float4 ga = sm.gt_a[g], gb = sm.gt_b[g];
                    if (gb.w != gb.w) continue;
                    const BoxC gc = gt_box(ga, gb);
                    ColBest cb = {-INFINITY, 0xffffffffu};
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        const int q = valid[k] ? p0 + k * TT + tid : p0;
                        const float v = pair_ciou(load_prior(prm.pri_xyxy, prm.pri_aux, q), gc);
                        if (v > best_v[k] || (v == best_v[k] && g < best_g[k])) { best_v[k] = v; best_g[k] = g; }
                        if (valid[k] && (v > cb.v || cb.p == 0xffffffffu)) { cb.v = v; cb.p = (unsigned)q; }
                    }
                    publish_column(cb, &sm.col[g]);
                }
                cluster.sync();
                for (int g = tid; g < G; g += TT) {
                    unsigned long long key = 0ull;
                    for (int r = 0; r < cs; ++r) {
                        const unsigned long long kr = cluster.map_shared_rank(sm.col, r)[g];
                        key = kr > key ? kr : key;
                    }
                    const float4 gb = sm.gt_b[g];
                    sm.champ[g] = (gb.w != gb.w) ? 0u : (0xffffffffu - (unsigned)(key & 0xffffffffull));
                }
                __syncthreads();
            }
        }
        for (int g = tid; g < G; g += TT) {
            const int champ = (int)sm.champ[g];
            if (champ >= p0 && champ < p1) atomicMin(&sm.forced[champ - p0], g);
        }
        __syncthreads();
        const int first_nan = st.first_nan;

        // ---- resolve rows, write targets ----------------------------------------------------
        int my_pos = 0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int p = p0 + k * TT + tid;
            if (!valid[k]) continue;
            if (G > 0) {
                const int f = sm.forced[p - p0];
                if (first_nan != INT_MAX && p != 0) { best_g[k] = first_nan; best_v[k] = __int_as_float(0x7fc00000); }
                else if (f != INT_MAX) { best_g[k] = f; best_v[k] = 2.0f; }
                pos[k] = best_v[k] >= prm.thresh;
                cls[k] = pos[k] ? sm.gt_label[best_g[k]] + 1 : 0;
            }
            my_pos += pos[k] ? 1 : 0;
            const long long row = (long long)b * P + p;
            if (prm.code) prm.code[row] = pos[k] ? (uint16_t)(best_g[k] + 1) : (uint16_t)0;
            {
                if (prm.pos_mask) prm.pos_mask[row] = pos[k] ? 1 : 0;
                if (prm.cls_t) prm.cls_t[row] = (int64_t)cls[k];
                if (prm.matched32) prm.matched32[row] = best_g[k];
                const bool want_loc = prm.loc_t && (!prm.loc_pos_only || pos[k]);
                if (want_loc || prm.matched_box) {
                    float4 gbox = make_float4(0.f, 0.f, 0.f, 0.f), t = gbox;
                    if (G > 0) {
                        const float4 ga = sm.gt_a[best_g[k]], gb = sm.gt_b[best_g[k]];
                        gbox = make_float4(gb.y, gb.z, fsub(ga.z, ga.x), fsub(ga.w, ga.y));
                        t = encode_offsets(gbox, ldg4(prm.pri + 4ll * p), prm.inv_vc, prm.inv_vs);
                    }
                    if (want_loc) reinterpret_cast<float4*>(prm.loc_t)[row] = t;
                    if (prm.matched_box) reinterpret_cast<float4*>(prm.matched_box)[row] = gbox;
                }
            }
        }
        const int cta_pos = block_sum<int>(my_pos, st.iscratch);
        if (tid == 0) st.npos_cta = cta_pos;
    }

    // ---- positives of the whole image (build_targets form only) ---------------------------
    if (prm.n_pos) {
        cluster.sync();
        int n_pos_img = 0;
        for (int r = 0; r < cs; ++r) n_pos_img += cluster.map_shared_rank(&st, r)->npos_cta;
        if (rank == 0 && tid == 0) prm.n_pos[b] = n_pos_img;
    }
    cluster.sync();          // keep this CTA's shared memory alive until every peer has read it
}

// ---------------------------------------------------------------------------------------------
// gt_prepare_kernel: one warp per ground-truth box -- normalised box, its CIoU constants, label and
// the prune bound lim (header comment) from the ~30 seed priors.  Done once per box instead of
// once per CTA of the image's cluster.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gt_prepare_kernel(const TrainParams prm, int want_seeds) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= prm.B * prm.max_gt) return;
    const int b = w / prm.max_gt, g = w % prm.max_gt;
    const int g_begin = prm.gt_offsets[b];
    if (g >= prm.gt_offsets[b + 1] - g_begin) return;
    const float4 px = ldg4(prm.gt_boxes + 4ll * (g_begin + g));
    const BoxC c = box_consts(fdiv(px.x, prm.norm_w), fdiv(px.y, prm.norm_h), fdiv(px.z, prm.norm_w), fdiv(px.w, prm.norm_h), true);
    float lim = 1e-30f;          // overlap-only pruning unless a seed raises it
    if (want_seeds && prm.P == 8732) {
        const int lv = kSeedLevel[lane], side = kLevelSide[lv];
        int ix = (int)floorf(c.xc * (float)side), iy = (int)floorf(c.yc * (float)side);
        ix = min(max(ix, 0), side - 1);
        iy = min(max(iy, 0), side - 1);
        const int ps = kLevelOffset[lv] + (iy * side + ix) * kLevelShapes[lv] + kSeedShape[lane];
        const float v = pair_ciou(load_prior(prm.pri_xyxy, prm.pri_aux, ps), c);
        const float cb0 = ord_decode(__reduce_max_sync(FULL, ord_encode(v)));   // NaN sorts to the top
        if (cb0 > 0.0f) lim = fmaxf(fmul(kPruneSlack, fminf(cb0, prm.thresh)), 1e-30f);
    }
    if (lane == 0) {
        float4* rec = prm.gt_rec + 3ll * ((long long)b * prm.max_gt + g);
        rec[0] = make_float4(c.x1, c.y1, c.x2, c.y2);
        rec[1] = make_float4(c.area, c.xc, c.yc, c.at);
        rec[2] = make_float4(lim, __int_as_float((int)prm.gt_labels[g_begin + g]), 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// loss_image_kernel: CE + smooth-L1 + hard-negative mining of one image per CTA
// ---------------------------------------------------------------------------------------------
constexpr int LT = 768;                        // threads; 2 CTAs per SM keep every image of a 256-batch resident

struct LossShared {
    unsigned hist[256];
    double dscratch[32];
    int iscratch[32];
    unsigned sel_digit, sel_need, sel_eq;
};

// per-prior target of the loss kernel: -1 = negative, >= 0 target class of a positive (mg = matched box)
template <bool FROM_TARGETS>
__device__ __forceinline__ int load_target(const TrainParams& prm, long long row, int g_begin, int& mg) {
    mg = -1;
    if (FROM_TARGETS) return prm.in_pos[row] != 0 ? (int)prm.in_cls[row] : -1;
    const int code = (int)prm.code[row];
    if (code == 0) return -1;
    mg = code - 1;
    return (int)prm.gt_labels[g_begin + mg] + 1;
}

constexpr unsigned kNotNegative = 0xffffffffu;   // key of a positive prior (never a float >= 0 bit pattern)

// ---------------------------------------------------------------------------------------------
// Exact hard-negative selection of one image, shared by both loss kernels.  `key_at(p)` is the CE
// bit pattern of prior p if it is a negative (CE >= 0: bit order = value order) and kNotNegative
// otherwise; ls.hist must already hold the histogram of key >> 24 over the negatives (the first
// radix pass rides along with the pass that produced the keys).  `tgt_of(p, mg)` returns the
// target class of a positive (and its matched box in mg) or -1.  Returns this thread's share of
// the sum of the selected negatives' CE; writes sel_cls / matched16 when they are requested.
// Every thread of the CTA must call it.
// ---------------------------------------------------------------------------------------------
template <int NT, typename KeyAt, typename TgtOf>
__device__ __forceinline__ double mined_exact_tail(const TrainParams& prm, int b, int P, int n_pos_img, LossShared& ls,
                                                   KeyAt key_at, TgtOf tgt_of) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc = 0.0;
    // ---- hard-negative budget (SSD_trainer.py:585-596) --------------------------------------
    const long long n_neg = (long long)P - n_pos_img;
    long long want = (n_pos_img == 0) ? (long long)prm.ratio : (long long)(prm.ratio * (double)n_pos_img);
    if (want < 0) want = 0;
    const long long kk = want < n_neg ? want : n_neg;
    unsigned thr_key = 0u;        // selected negatives: key > thr_key, plus `need` of those == thr_key
    unsigned need = 0u, n_eq = 0u;
    const bool take_all = (kk >= n_neg);
    if (kk > 0 && !take_all) {
        // 4-pass MSD radix select of the kk-th largest key
        unsigned prefix = 0u, remaining = (unsigned)kk;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (pass > 0) {
                for (int i = tid; i < 256; i += NT) ls.hist[i] = 0u;
                __syncthreads();
                const unsigned himask = 0xffffffffu << (shift + 8);
                for (int p = tid; p < P; p += NT) {
                    const unsigned key = key_at(p);
                    if (key != kNotNegative && ((key & himask) == prefix)) atomicAdd(&ls.hist[(key >> shift) & 255u], 1u);
                }
                __syncthreads();
            }
            if (tid < 32) {
                // bins 255..0: lane l owns bins [255-8l-7 .. 255-8l]; find where the suffix count reaches `remaining`
                unsigned mine = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) mine += ls.hist[255 - (tid * 8 + j)];
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
                        const unsigned c = ls.hist[bin];
                        if (run + c >= remaining) { ls.sel_digit = (unsigned)bin; ls.sel_need = remaining - run; ls.sel_eq = c; break; }
                        run += c;
                    }
                }
            }
            __syncthreads();
            prefix |= ls.sel_digit << shift;
            remaining = ls.sel_need;
            n_eq = ls.sel_eq;
            __syncthreads();
        }
        thr_key = prefix;
        need = remaining;
    }

    // ---- sums (and the backward selection) --------------------------------------------------
    if (kk > 0) {
        for (int p = tid; p < P; p += NT) {
            const unsigned key = key_at(p);
            if (key != kNotNegative && (take_all || key > thr_key)) acc += (double)__uint_as_float(key);
        }
    }
    if (prm.sel_cls) {
        // of the negatives equal to the threshold value the first `need` in prior order are taken
        const bool all_ties = take_all || kk == 0 || need == n_eq;
        int before = 0;
        for (int base = 0; base < P; base += NT) {
            const int p = base + tid;
            const unsigned key = p < P ? key_at(p) : kNotNegative;
            const bool tie = kk > 0 && !take_all && key != kNotNegative && key == thr_key;
            int rank_tie = 0;
            if (!all_ties) {
                const unsigned bal = __ballot_sync(FULL, tie);
                __syncthreads();
                if (lane == 0) ls.iscratch[warp] = __popc(bal);
                __syncthreads();
                int slot_total = 0;
                for (int w = 0; w < NT / 32; ++w) { if (w < warp) rank_tie += ls.iscratch[w]; slot_total += ls.iscratch[w]; }
                rank_tie += before + __popc(bal & ((1u << lane) - 1u));
                before += slot_total;
            }
            if (p >= P) continue;
            const long long row = (long long)b * P + p;
            int mg;
            const int tgt = tgt_of(p, mg);
            int8_t sel = -1;
            if (tgt >= 0) sel = (int8_t)tgt;
            else if (kk > 0 && (take_all || key > thr_key || (tie && (all_ties || (unsigned)rank_tie < need)))) sel = 0;
            prm.sel_cls[row] = sel;
            if (prm.matched16) prm.matched16[row] = (int16_t)mg;
        }
    }
    // the `need` threshold-valued negatives are counted once
    if (tid == 0 && kk > 0 && !take_all) acc += (double)need * (double)__uint_as_float(thr_key);

    return acc;
}

template <int CT, bool FROM_TARGETS>
__global__ void __launch_bounds__(LT, 2) loss_image_kernel(const TrainParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ LossShared ls;
    unsigned* keys = reinterpret_cast<unsigned*>(dyn);     // [P] CE bits of negatives (CE >= 0: bit order = value order)
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = prm.P;
    const int g_begin = FROM_TARGETS ? 0 : prm.gt_offsets[b];

    for (int i = tid; i < 256; i += LT) ls.hist[i] = 0u;
    __syncthreads();

    double acc_loc = 0.0, acc_ce = 0.0;
    int my_pos = 0;
#pragma unroll 2
    for (int p = tid; p < P; p += LT) {
        const long long row = (long long)b * P + p;
        int mg;
        const int tgt = load_target<FROM_TARGETS>(prm, row, g_begin, mg);
        const int cls = tgt < 0 ? 0 : tgt;
        float mx, lg;
        row_lse<CT>(prm.conf_all + row * prm.C, prm.C, mx, lg);
        const float xc = __ldg(prm.conf_all + row * prm.C + cls);
        // -log_softmax[c] = -((x_c - max) - log(sum))  (ATen PersistentSoftmax.cuh, nll_loss)
        const float ce = -fsub(fsub(xc, mx), lg);
        if (tgt >= 0) {
            my_pos += 1;
            acc_ce += (double)ce;
            keys[p] = kNotNegative;
            if (!FROM_TARGETS && prm.loc_all) {
                const float4 px = ldg4(prm.gt_boxes + 4ll * (g_begin + mg));
                const float x1 = fdiv(px.x, prm.norm_w), y1 = fdiv(px.y, prm.norm_h);
                const float x2 = fdiv(px.z, prm.norm_w), y2 = fdiv(px.w, prm.norm_h);
                const float4 gbox = make_float4(fmul(fadd(x1, x2), 0.5f), fmul(fadd(y1, y2), 0.5f), fsub(x2, x1), fsub(y2, y1));
                const float4 t = encode_offsets(gbox, ldg4(prm.pri + 4ll * p), prm.inv_vc, prm.inv_vs);
                const float4 l = ldg4(prm.loc_all + 4ll * row);
                const float d[4] = {fsub(l.x, t.x), fsub(l.y, t.y), fsub(l.z, t.z), fsub(l.w, t.w)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float z = fabsf(d[j]);
                    acc_loc += (double)(z < 1.0f ? fmul(fmul(0.5f, z), z) : fsub(z, 0.5f));
                }
            }
        } else {
            const unsigned key = __float_as_uint(ce) & 0x7fffffffu;      // -0.0 (sum == 1, target is the max) counts as 0
            keys[p] = key;
            atomicAdd(&ls.hist[key >> 24], 1u);             // first radix pass rides along
        }
    }
    const int n_pos_img = block_sum<int>(my_pos, ls.iscratch);      // (two barriers: keys and hist are complete)
    if (tid == 0 && prm.n_pos) prm.n_pos[b] = n_pos_img;

    acc_ce += mined_exact_tail<LT>(prm, b, P, n_pos_img, ls, [&](int p) { return keys[p]; },
                                   [&](int p, int& mg) { return load_target<FROM_TARGETS>(prm, (long long)b * P + p, g_begin, mg); });

    const double s_loc = block_sum<double>(acc_loc, ls.dscratch);
    const double s_ce = block_sum<double>(acc_ce, ls.dscratch);
    if (tid == 0) {
        prm.img_part[2ll * b + 0] = s_loc;
        prm.img_part[2ll * b + 1] = s_ce;
    }
}

// =============================================================================================
// train_image_kernel -- the SSD300 fast path: sparse matching + mined loss, ONE CTA PER IMAGE
// =============================================================================================
// Matching is box-centric.  The priors of SSD300 sit on six regular grids (SSD_from_scratch.py:
// 289-323), so for a ground-truth box and one (level, shape) combination the priors whose IoU can
// reach lim (the same bound the pruned sweep of match_kernel uses) form a small rectangle of
// cells: the 2-D IoU is at most each 1-D IoU, and IoU >= lim needs inter (1 + lim) >= lim (a_p + a_g)
// with inter <= w_ov * min(h_p, h_g); both are evaluated with the exact clamped prior extents
// (axis_hull: closed form for interior cells, the few border-clamped cells one by one).
// Those ~60 cells per box (instead of 8732 priors) pass the same cheap IoU gate as before; the
// ~35 survivors get the exact CIoU and update a packed (CIoU, ~box) maximum per prior in shared
// memory plus the packed (CIoU, ~prior) maximum per box.  Boxes the rectangles cannot handle (not
// finite, empty, or whose best CIoU is not positive) are swept densely, exactly like the refill of
// match_kernel.  Results are bit-identical to the exact-everywhere sweep for every positive prior.
//
// The loss then streams the image's logits once.  Hard-negative mining only needs the k largest
// cross-entropies of ~8.7k negatives, so the stream evaluates CE with ex2.approx / lg2.approx
// (|error| <= 1e-5 + 1e-6 CE, see approx_ce6), a 4096-bin histogram brackets the k-th largest
// approximate value in a bin 0.4 % wide, and only the negatives above that bracket
// and the handful inside a +-3-error band around it are re-evaluated with the exact eager-CUDA
// arithmetic: every negative above the band is certainly among the k hardest, every one below it
// certainly not, and the band members are ranked exactly (ties: lower prior index first).  Images
// where that does not apply (budget >= #negatives, an over-full band, non-finite bracket) take the
// exact path of loss_image_kernel (mined_exact_tail).  The sums are therefore those of the exact
// arithmetic in every case.
constexpr int FT = 768;                      // threads; 2 CTAs per SM
constexpr int FAST_MAX_GT = 64;
constexpr int GT_ROUND = 32;                 // boxes enumerated per round
constexpr int NSEG = GT_ROUND * 32;          // (box, level-shape) segments per round
constexpr int PAIR_CAP = 1536;               // listed gate survivors of an image (rest: settled inline, box recomputed)
#ifndef SSDHOT_MT_LOSS
#define SSDHOT_MT_LOSS 576
#endif
constexpr int MT_LOSS = SSDHOT_MT_LOSS;      // threads (warps 0..17) that match while the other six warps stream the logits (A/B on B200: 512 -> 43.5 us, 576 -> 41.6 us, 448 -> 45.6 us at B = 256)
#ifndef SSDHOT_MT_SHARE
#define SSDHOT_MT_SHARE 448
#endif
// ... and when predict waits for this stream's row keys (share buffer): ten stream warps deliver them ~5 us earlier; the kernel
// alone is slower that way (44.1 vs 41.1 us), the forked step faster (B = 256: 69.1 -> 64.0 us)
constexpr int MT_SHARE = SSDHOT_MT_SHARE;
constexpr int POS_CAP = 2048;                // listed positive priors of an image (more: the exact tail walks the slots)
constexpr int SEL_CAP = 4096;                // listed certainly-mined negatives
constexpr int BAND_CAP = 1024;
constexpr unsigned kOrdTwo = 0xC0000000u;    // ord_encode(2.0f): the forced-match value (SFS:747)
constexpr int FUSED_SCRATCH = 4112 + 8192 + 8 * PAIR_CAP + 2 * NSEG;     // 26640: matching view; the mining lists reuse it

struct FusedStatic {
    float4 gt_a[FAST_MAX_GT];                // x1 y1 x2 y2 (normalised)
    float4 gt_b[FAST_MAX_GT];                // area xc yc atan
    unsigned long long col[FAST_MAX_GT];     // (ord(ciou) << 32 | ~prior) best prior per box
    float lim[FAST_MAX_GT];
    int label[FAST_MAX_GT];
    int champ[FAST_MAX_GT];
    unsigned char kind[FAST_MAX_GT];         // 0 = rectangles, 1 = dense sweep, 2 = all-NaN column
    // the 30 (level, shape) combinations: grid side, shapes per cell, first prior (level offset + shape), (w, h)
    int cb_side[32], cb_shapes[32], cb_base[32];
    float cb_w[32], cb_h[32], cb_inv[32];      // (w, h) of the shape, 1 / side
    float cb_hw[32], cb_hh[32];                // fl(0.5 w), fl(0.5 h): the half extents the clamped corners are built from
    int cb_coff[32];                           // first entry of the level's centres in `centre`
    float centre[80];                          // (i + 0.5) / side of every level, back to back (38 + 19 + 10 + 5 + 3 + 1 = 76)
    float4 gt_px[FAST_MAX_GT];                // the image's boxes as given (prefetched while the table is cleared)
    LossShared ls;
    int first_nan, n_dense, n_pair, n_work, n_seg, n_chunk, n_band, n_sure, pair_overflow, n_pos_img;
    int n_pos_listed, n_sel;                 // listed positives (claim), listed certainly-mined negatives (scan)
    double fin_odd[2][FT / 32];              // the final exchange's double parts
    unsigned r_bin, r_above;
};

__host__ __device__ inline size_t fused_smem_bytes(int P) {
    return (size_t)P * 8 + FUSED_SCRATCH + 8192;         // table | scratch | CE histogram
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Approximate -log_softmax(row)[0] of a 6-logit row.  d_i = x_i - max is the exact fp32 difference
// the reference forms; each ex2.approx term is within 2.6e-7 absolute, the sum s in [1, 6] within
// 4e-6 relative, ln(s) within 5e-6, so |result - exact CE| <= 6e-6 + 2e-7 CE  (bound used: 1e-5 + 1e-6 CE).
template <bool RK>
__device__ __forceinline__ float approx_ce6_rk(float x0, float x1, float x2, float x3, float x4, float x5, unsigned& rk) {
    const float mx = fmaxf(fmaxf(fmaxf(x0, x1), fmaxf(x2, x3)), fmaxf(x4, x5));
    const float k = 1.4426950408889634f;
    const float d0 = x0 - mx;
    const float e1 = ex2_approx((x1 - mx) * k), e2 = ex2_approx((x2 - mx) * k), e3 = ex2_approx((x3 - mx) * k);
    const float e4 = ex2_approx((x4 - mx) * k), e5 = ex2_approx((x5 - mx) * k);
    const float s = ((ex2_approx(d0 * k) + e1) + (e2 + e3)) + (e4 + e5);
    const float ce = lg2_approx(s) * 0.6931471805599453f - d0;
    if (RK) {
        // the row key of predict_image_kernel (infer_path.cu: stream_row_keys): the bin of the row's best approximate foreground
        // score, from the exps already at hand; never 0 for a row of numbers, 0 for a row that holds a NaN / +Inf (no candidate)
        float rs;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s));
        const float sc = fmaxf(fmaxf(fmaxf(e1, e2), fmaxf(e3, e4)), e5) * rs;
        rk = (s == s) ? max(__float_as_uint(sc) >> 15, 1u) : 0u;
    }
    // a NaN / +-Inf logit makes the exact CE NaN or +Inf, which torch.topk ranks above every number (TR:597): keep it non-finite
    // here (fmaxf would turn NaN into 0) so that its key lands in the top, unresolved bin and the exact arithmetic decides
    return (ce != ce) ? __int_as_float(0x7fffffff) : fmaxf(ce, 0.0f);
}
__device__ __forceinline__ float approx_ce6(float x0, float x1, float x2, float x3, float x4, float x5) {
    unsigned rk;
    return approx_ce6_rk<false>(x0, x1, x2, x3, x4, x5, rk);
}
__device__ __forceinline__ float ce_error_bound(float ce) { return 1e-5f + 1e-6f * ce; }

// Histogram bin of a CE bit pattern: 16 octaves [2^-9, 2^7) x 256 mantissa steps = 4096 bins (a bin is 0.4 % wide,
// ~30 of 8.7k negatives); smaller / larger values share the end bins, which the selection treats as "unresolved".
constexpr int kCeBinBase = (127 - 9) << 8;
__device__ __forceinline__ int ce_bin(unsigned key) {
    const int v = (int)(key >> 15) - kCeBinBase;
    return min(max(v, 0), 4095);
}
__device__ __forceinline__ void ce_hist_add(unsigned* hist16, unsigned key) {
    const int bin = ce_bin(key);
    atomicAdd(&hist16[bin >> 1], 1u << ((bin & 1) * 16));
}

// exact CE of one row for target class cls, in eager torch-CUDA order (as loss_image_kernel)
__device__ __forceinline__ float exact_ce6(const float* __restrict__ row, int cls) {
    float mx, lg;
    row_lse<6>(row, 6, mx, lg);
    return -fsub(fsub(__ldg(row + cls), mx), lg);
}

// the same from a head source (heads.cuh): row p of the image
template <int SRC>
__device__ __forceinline__ float exact_ce6(const HeadReader<SRC, 6>& rd, int p, int cls) {
    if (SRC == SRC_PACKED) return exact_ce6(rd.packed + 6ll * p, cls);
    float x[6];
    rd.row(p, x);
    const float mx = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(x[4], x[5]));
    const float e0 = expf(fsub(x[0], mx)), e1 = expf(fsub(x[1], mx)), e2 = expf(fsub(x[2], mx));
    const float e3 = expf(fsub(x[3], mx)), e4 = expf(fsub(x[4], mx)), e5 = expf(fsub(x[5], mx));
    const float lg = logf(fadd(fadd(fadd(e0, e4), e2), fadd(fadd(e1, e5), e3)));    // 8-lane butterfly order (row_lse)
    const float xc = cls == 0 ? x[0] : cls == 1 ? x[1] : cls == 2 ? x[2] : cls == 3 ? x[3] : cls == 4 ? x[4] : x[5];
    return -fsub(fsub(xc, mx), lg);
}

// From the top bin downwards, find the bin in which the running count reaches k (1 <= k <= total).
// count(bin) reads a bin; every thread of the CTA must call it; two barriers inside.
template <int NT, int NBINS, typename CountFn>
__device__ __forceinline__ void find_kth_from_top(CountFn count, unsigned k, FusedStatic& fs, unsigned& bin_out, unsigned& above_out) {
    constexpr int BPT = (NBINS + NT - 1) / NT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned c[BPT];
    unsigned mine = 0u;
#pragma unroll
    for (int j = 0; j < BPT; ++j) {
        const int bin = NBINS - 1 - (tid * BPT + j);
        c[j] = bin >= 0 ? count(bin) : 0u;
        mine += c[j];
    }
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    __syncthreads();
    if (lane == 31) fs.ls.iscratch[warp] = (int)incl;
    __syncthreads();
    unsigned excl = incl - mine;
    for (int w = 0; w < warp; ++w) excl += (unsigned)fs.ls.iscratch[w];
    if (excl < k && k <= excl + mine) {
        unsigned run = excl;
#pragma unroll
        for (int j = 0; j < BPT; ++j) {
            if (run < k && run + c[j] >= k) { fs.r_bin = (unsigned)(NBINS - 1 - (tid * BPT + j)); fs.r_above = run; }
            run += c[j];
        }
    }
    __syncthreads();
    bin_out = fs.r_bin;
    above_out = fs.r_above;
}

// Extent of the clamped prior interval of cell i on a grid of `side` cells: [max(0, c - w/2), min(1, c + w/2)], c = (i + .5)/side
// (bound arithmetic only: approximate reciprocals are fine, every comparison built on it is slackened)
__device__ __forceinline__ float clamped_extent(int i, float inv_side, float w) {
    const float c = ((float)i + 0.5f) * inv_side;
    return fminf(1.0f, c + 0.5f * w) - fmaxf(0.0f, c - 0.5f * w);
}

// Hull [lo, hi] of the cells whose (clamped) prior interval can satisfy  ov * A >= Bc * extent + Cc,  ov = its
// overlap with [g1, g2]; empty if hi < lo.  Un-clamped cells have extent w and ov <= min(w, gw, (w + gw)/2 - |c - gc|),
// which gives a closed-form centre range.  The nb cells the image border clamps on each side have a smaller extent -- at
// least w/2, their centre lies inside the image -- and an overlap no larger than the un-clamped one, so the same closed
// form with w/2 in the requirement bounds them; only border cells take that looser range.  Every comparison is
// slackened, so the hull errs on the inclusive side (a superset costs a few more gate tests, never a match).
__device__ __forceinline__ void axis_hull(float g1, float g2, int side, float inv_side, float w, float A, float Bc, float Cc, int& lo, int& hi) {
    const float fS = (float)side, gw = g2 - g1, gc = 0.5f * (g1 + g2);
    lo = side;
    hi = -1;
    const float rA = __fdividef(1.0f, A), wmin = fminf(w, gw), half = 0.5f * (w + gw);
    const float areq = (Bc * w + Cc) * rA * 0.9999f;
    if (wmin >= areq) {
        const float r = half - areq;
        const int ia = max(0, (int)ceilf(fmaxf((gc - r) * fS - 0.51f, -1.0f)));
        const int ib = min(side - 1, (int)floorf(fminf((gc + r) * fS - 0.49f, fS)));
        if (ia <= ib) { lo = ia; hi = ib; }
    }
    const int nb = min(side, max(0, (int)ceilf(0.5f * w * fS - 0.49f)));      // cells clamped at each border
    const bool left = g1 < w, right = g2 > 1.0f - w;
    if (nb > 0 && (left || right)) {
        const float areq2 = (Bc * 0.5f * w + Cc) * rA * 0.999f;
        if (wmin >= areq2) {
            const float r = half - areq2;
            const int ja = max(0, (int)ceilf(fmaxf((gc - r) * fS - 0.51f, -1.0f)));
            const int jb = min(side - 1, (int)floorf(fminf((gc + r) * fS - 0.49f, fS)));
            if (left && ja <= min(jb, nb - 1)) { lo = min(lo, ja); hi = max(hi, min(jb, nb - 1)); }
            if (right && max(ja, side - nb) <= jb) { lo = min(lo, max(ja, side - nb)); hi = max(hi, jb); }
        }
    }
    (void)inv_side;
}

#define SSDHOT_STAMP(k) do { if (prm.timeline && (threadIdx.x & 31) == 0) { if (threadIdx.x == 0 || (k) == 2) prm.timeline[(long long)blockIdx.x * 16 + (k)] = globaltimer_ns(); } } while (0)

// Barrier of the threads that run the matching: the whole CTA (MT == FT) or its first MT threads (named barrier 1),
// while the remaining warps stream the logits.
template <int MT>
__device__ __forceinline__ void role_sync() {
    if (MT == FT) __syncthreads();
    else asm volatile("bar.sync 1, %0;" ::"n"(MT) : "memory");
}

// Prior (level-shape combination, cell i, j) exactly as the device tables hold it -- clamped corners and the per-prior CIoU
// constants -- from arithmetic alone, so that matching never waits on the prior tables in L2 under the stream's load.
// Bit-identical to prior_tables_kernel: the centres of SSD300 are fl32((i + 0.5) / side) computed in float64 (SFS:318-325),
// which equals the correctly rounded fp32 quotient (2 i + 1) / (2 side) (the double's rounding error, 2^-53, is far below
// the distance 1 / (2 side 2^25) of a non-dyadic small-integer ratio from any fp32 rounding boundary); the host layout check
// admits only priors that satisfy this exactly.
__device__ __forceinline__ BoxC grid_prior(const FusedStatic& fs, int combo, int i, int j, bool want_atan) {
    const int c0 = fs.cb_coff[combo];
    const float cx = fs.centre[c0 + i], cy = fs.centre[c0 + j];
    const float hw = fs.cb_hw[combo], hh = fs.cb_hh[combo];
    const float x1 = fminf(fmaxf(fsub(cx, hw), 0.f), 1.f), y1 = fminf(fmaxf(fsub(cy, hh), 0.f), 1.f);
    const float x2 = fminf(fmaxf(fadd(cx, hw), 0.f), 1.f), y2 = fminf(fmaxf(fadd(cy, hh), 0.f), 1.f);
    return box_consts(x1, y1, x2, y2, want_atan);
}

// Shared-memory views of the fused kernel.  The table has one 64-bit slot per prior: the HIGH word is the best
// ord(CIoU) any box reached at that prior (matching, 32-bit atomicMax), the LOW word is written by the logit
// stream (CE bits of the prior as a negative) and, once a prior is known to be positive, replaced by
// 0x80000000 | (63 - matched box) -- so the two roles never touch the same word.
struct FusedViews {
    unsigned* lo;              // lo[2p]
    unsigned* hi;              // hi[2p]  (= lo + 1)
    uint16_t* seg_n;           // [NSEG] cells of a listed (box, level-shape) rectangle
    unsigned* seg_info;        // [NSEG] i0 | j0<<6 | ni<<12 | combination<<18 | box<<23
    uint16_t* chunk_list;      // [2 NSEG] rectangle | (chunk of 32 cells)<<10: the work items of the gate, one warp each
    unsigned* pair_pg;         // [PAIR_CAP] prior | box << 16 : every pair that passed the IoU gate (all rounds)
    unsigned* pair_ov;         // [PAIR_CAP] ord(CIoU) of that pair (until 1d: its grid position, combination | i<<5 | j<<11)
    uint16_t* work_list;       // [NSEG] (box in round) << 5 | level-shape
};

// ---- matching (sections 1 and 2 of the kernel), run by the first MT threads of the CTA --------------------
template <int MT>
__device__ __forceinline__ void match_phase(const TrainParams& prm, FusedStatic& fs, const FusedViews& v, int G, int g_begin) {
    const int mtid = threadIdx.x, lane = mtid & 31, mwarp = mtid >> 5;
    const int P = prm.P;
    const float thresh = prm.thresh;
    // record a pair whose CIoU may matter: best value per prior (high word), best prior per box, list entry
    auto settle_box = [&](const BoxC& pr, int p, int g, int slot) {
        const float val = pair_ciou(pr, gt_box(fs.gt_a[g], fs.gt_b[g]));
        const unsigned o = ord_encode(val);
        atomicMax(&v.hi[2 * p], o);
        const unsigned long long ck = ((unsigned long long)o << 32) | (unsigned long long)(0xffffffffu - (unsigned)p);
        if (ck > *reinterpret_cast<volatile unsigned long long*>(&fs.col[g])) atomicMax(&fs.col[g], ck);
        if (slot < PAIR_CAP) v.pair_ov[slot] = o;
    };
    // cell t of listed rectangle sg: the IoU gate; a survivor joins the pair list with its grid position.  Called by a whole
    // warp (lane = cell of a chunk; `solo` = false) or by a single thread working off an overfull chunk list (`solo`).
    auto gate_cell = [&](int sg, int t, bool solo) {
        const unsigned info = v.seg_info[sg];
        const int n = (int)v.seg_n[sg];
        const int i0 = (int)(info & 63u), j0 = (int)((info >> 6) & 63u), ni = (int)((info >> 12) & 63u);
        const int combo = (int)((info >> 18) & 31u), g = (int)(info >> 23);
        bool pass = false;
        int p = 0, i = 0, j = 0;
        if (t < n) {
            const int lj = (int)(((float)t + 0.5f) * __fdividef(1.0f, (float)ni));      // t / ni  (t < 4096, ni < 64: exact)
            i = i0 + t - lj * ni;
            j = j0 + lj;
            p = fs.cb_base[combo] + (j * fs.cb_side[combo] + i) * fs.cb_shapes[combo];
            const BoxC pr = grid_prior(fs, combo, i, j, false);
            const float4 ga = fs.gt_a[g];
            const float w = fmaxf(fsub(fminf(pr.x2, ga.z), fmaxf(pr.x1, ga.x)), 0.0f);
            const float h = fmaxf(fsub(fminf(pr.y2, ga.w), fmaxf(pr.y1, ga.y)), 0.0f);
            const float inter = fmul(w, h);
            const float uni = fsub(fadd(pr.area, fs.gt_b[g].x), inter);
            pass = !(inter < fmul(fs.lim[g], uni));      // NaN-safe: anything odd takes the exact path
        }
        int dst = 0;
        if (solo) {
            if (!pass) return;
            dst = atomicAdd(&fs.n_pair, 1);
        } else {
            const unsigned bal = __ballot_sync(FULL, pass);
            if (!bal) return;
            if (lane == 0) dst = atomicAdd(&fs.n_pair, __popc(bal));
            dst = __shfl_sync(FULL, dst, 0) + __popc(bal & ((1u << lane) - 1u));
            if (!pass) return;
        }
        if (dst < PAIR_CAP) {
            v.pair_pg[dst] = (unsigned)p | ((unsigned)g << 16);
            v.pair_ov[dst] = (unsigned)combo | ((unsigned)i << 5) | ((unsigned)j << 11);
        } else { settle_box(grid_prior(fs, combo, i, j, true), p, g, dst); fs.pair_overflow = 1; }   // unlisted: its box is recovered by recomputation
    };
    for (int g0 = 0; g0 < G; g0 += GT_ROUND) {
        const int gn = min(GT_ROUND, G - g0);
        if (mtid == 0) { fs.n_work = 0; fs.n_seg = 0; fs.n_chunk = 0; }
        const int round_begin = min(fs.n_pair, PAIR_CAP);       // (stable: the previous round ended with a barrier)
        role_sync<MT>();
        // 1a. a warp per box (one of the 30 seed priors per lane) while the round's boxes fit the matching warps, else half a
        //     warp per box (two seed priors per lane: the longest dependent chain of the phase, twice): constants, seed
        //     bound, and the (level, shape) combinations whose sizes can reach IoU >= lim at all (1-D and area ratios; a
        //     clamped extent is at least half the nominal one)
        const bool wide = gn <= MT / 32;
        const int sub = wide ? lane : (lane & 15), nh = wide ? 1 : 2;
        for (int gl = wide ? mwarp : 2 * mwarp + (lane >> 4); gl < (wide ? gn : ((gn + 1) & ~1));
             gl += wide ? MT / 32 : 2 * (MT / 32)) {                                              // warp-uniform trip count
            const bool live = gl < gn;
            const int g = g0 + (live ? gl : gn - 1);
            const float4 px = fs.gt_px[g];
            const BoxC c = box_consts(fdiv(px.x, prm.norm_w), fdiv(px.y, prm.norm_h), fdiv(px.z, prm.norm_w), fdiv(px.w, prm.norm_h), true);
            const float gw = fsub(c.x2, c.x1), gh = fsub(c.y2, c.y1);
            int kind = 0;
            if (c.at != c.at) kind = 2;
            else if (!(gw > 0.0f && gh > 0.0f) || !(fabsf(c.x1) < 1e30f && fabsf(c.y1) < 1e30f && fabsf(c.x2) < 1e30f && fabsf(c.y2) < 1e30f)) kind = 1;
            float lim = 1e-30f;
            unsigned keep = 0u;                              // bit h: combination sub + 16 h stays
            unsigned best = 0u;
            if (kind == 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= nh) break;
                    const int combo = sub + 16 * h;          // 30, 31 repeat 0, 4 as seeds
                    const int side = fs.cb_side[combo];
                    int ix = (int)floorf(c.xc * (float)side), iy = (int)floorf(c.yc * (float)side);
                    ix = min(max(ix, 0), side - 1);
                    iy = min(max(iy, 0), side - 1);
                    best = max(best, ord_encode(pair_ciou(grid_prior(fs, combo, ix, iy, true), c)));
                }
            }
            // (the two halves of a warp may hold different boxes: the shuffles below must run in every lane)
            if (wide) best = max(best, __shfl_xor_sync(FULL, best, 16));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
            if (kind == 0) {
                const float cb0 = ord_decode(best);
                if (cb0 > 0.0f) lim = fmaxf(fmul(kPruneSlack, fminf(cb0, thresh)), 1e-30f);
                const float l2 = 0.99f * lim;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= nh) break;
                    const int combo = sub + 16 * h;
                    const float w = fs.cb_w[combo], hh = fs.cb_h[combo];
                    if (combo < 30 && !(gw < l2 * 0.5f * w || w < l2 * gw || gh < l2 * 0.5f * hh || hh < l2 * gh ||
                                        c.area < l2 * 0.25f * w * hh || w * hh < l2 * c.area)) keep |= 1u << h;
                }
            }
            if (!live) keep = 0u;
            const unsigned b0 = __ballot_sync(FULL, keep & 1u), b1 = __ballot_sync(FULL, keep & 2u);
            int dst = 0;
            if (lane == 0 && (b0 | b1)) dst = atomicAdd(&fs.n_work, __popc(b0) + __popc(b1));
            dst = __shfl_sync(FULL, dst, 0);
            const unsigned ltm = (1u << lane) - 1u;
            if (keep & 1u) v.work_list[dst + __popc(b0 & ltm)] = (uint16_t)((gl << 5) | sub);
            if (keep & 2u) v.work_list[dst + __popc(b0) + __popc(b1 & ltm)] = (uint16_t)((gl << 5) | (sub + 16));
            if (live && sub == 0) {
                fs.gt_a[g] = make_float4(c.x1, c.y1, c.x2, c.y2);
                fs.gt_b[g] = make_float4(c.area, c.xc, c.yc, c.at);
                fs.lim[g] = lim;
                fs.col[g] = 0ull;
                fs.champ[g] = 0;
                fs.kind[g] = (unsigned char)kind;
                if (kind == 2) atomicMin(&fs.first_nan, g);
                if (kind == 1) atomicAdd(&fs.n_dense, 1);
            }
        }
        role_sync<MT>();
        SSDHOT_STAMP(11);
        // 1a'. one thread per surviving (box, level-shape): its rectangle of candidate cells; the non-empty ones are listed
        {
            const int n_work = fs.n_work;
            for (int e = mtid; e < n_work; e += MT) {
                const int wk = (int)v.work_list[e], g = g0 + (wk >> 5), combo = wk & 31;
                const int side = fs.cb_side[combo];
                const float w = fs.cb_w[combo], h = fs.cb_h[combo], l2 = 0.99f * fs.lim[g];
                const float inv_side = fs.cb_inv[combo];
                const float4 ga = fs.gt_a[g];
                const float gw = ga.z - ga.x, gh = ga.w - ga.y, ag = fs.gt_b[g].x;
                // rows whose 1-D IoU with the box can reach lim (2-D IoU <= each 1-D IoU) ...
                int j0, j1, i0, i1, a0, a1;
                axis_hull(ga.y, ga.w, side, inv_side, h, 1.0f + l2, l2, l2 * gh, j0, j1);
                if (j1 < j0) continue;
                // ... then columns / rows that can satisfy inter (1 + lim) >= lim (a_p + a_g) given the best
                // overlap and the smallest clamped extent the other axis offers
                const float hcmin = fminf(clamped_extent(j0, inv_side, h), clamped_extent(j1, inv_side, h));
                axis_hull(ga.x, ga.z, side, inv_side, w, fminf(h, gh) * (1.0f + l2), l2 * hcmin, l2 * ag, i0, i1);
                if (i1 < i0) continue;
                const float wcmin = fminf(clamped_extent(i0, inv_side, w), clamped_extent(i1, inv_side, w));
                axis_hull(ga.y, ga.w, side, inv_side, h, fminf(w, gw) * (1.0f + l2), l2 * wcmin, l2 * ag, a0, a1);
                j0 = max(j0, a0);
                j1 = min(j1, a1);
                if (j1 < j0) continue;
                const int ni = i1 - i0 + 1, n = ni * (j1 - j0 + 1);
                const int at = atomicAdd(&fs.n_seg, 1);           // (n_work <= NSEG)
                v.seg_n[at] = (uint16_t)n;
                v.seg_info[at] = (unsigned)i0 | ((unsigned)j0 << 6) | ((unsigned)ni << 12) | ((unsigned)combo << 18) | ((unsigned)g << 23);
                // the rectangle's cells in chunks of 32: the gate's work items (a full list leaves the rest to this thread)
                const int nc = (n + 31) >> 5;
                const int c0 = atomicAdd(&fs.n_chunk, nc);
                for (int c = 0; c < nc; ++c) {
                    if (c0 + c < 2 * NSEG) v.chunk_list[c0 + c] = (uint16_t)(at | (c << 10));
                    else for (int t = 32 * c; t < min(n, 32 * c + 32); ++t) gate_cell(at, t, true);
                }
            }
        }
        role_sync<MT>();
        SSDHOT_STAMP(12);
        // 1c. one warp per chunk of 32 cells: the cheap IoU gate, the prior built from its grid position; survivors join
        //     the pair list with that position
        {
            const int n_chunk = min(fs.n_chunk, 2 * NSEG);
            for (int c = mwarp; c < n_chunk; c += MT / 32) {
                const int e = (int)v.chunk_list[c];
                gate_cell(e & 1023, 32 * (e >> 10) + lane, false);
            }
        }
        role_sync<MT>();
        SSDHOT_STAMP(13);
        // 1d. exact CIoU of this round's listed pairs
        {
            const int round_end = min(fs.n_pair, PAIR_CAP);
            for (int e = round_begin + mtid; e < round_end; e += MT) {
                const unsigned pr = v.pair_pg[e], at = v.pair_ov[e];
                settle_box(grid_prior(fs, (int)(at & 31u), (int)((at >> 5) & 63u), (int)((at >> 11) & 63u), true), (int)(pr & 0xffffu), (int)(pr >> 16), e);
            }
        }
        role_sync<MT>();
    }

    SSDHOT_STAMP(14);
    // ---- 2. champions; dense sweep of the columns the rectangles could not settle ------------------
    if (mtid < G) {
        const int g = mtid;
        if (fs.kind[g] == 0) {
            const unsigned long long key = fs.col[g];
            if ((unsigned)(key >> 32) > ord_encode(0.0f)) fs.champ[g] = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
            else { fs.kind[g] = 1; atomicAdd(&fs.n_dense, 1); }
        }
    }
    role_sync<MT>();
    auto list_pair = [&](int p, int g, unsigned o) {         // a pair found outside the rectangles
        const int dst = atomicAdd(&fs.n_pair, 1);
        if (dst < PAIR_CAP) { v.pair_pg[dst] = (unsigned)p | ((unsigned)g << 16); v.pair_ov[dst] = o; }
        else fs.pair_overflow = 1;
    };
    if (fs.n_dense > 0) {
        for (int g = 0; g < G; ++g) {
            if (fs.kind[g] != 1) continue;              // uniform over the role
            const BoxC gc = gt_box(fs.gt_a[g], fs.gt_b[g]);
            if (mtid == 0) fs.col[g] = 0ull;
            role_sync<MT>();
            unsigned long long best = 0ull;
            for (int p = mtid; p < P; p += MT) {
                const float val = pair_ciou(load_prior(prm.pri_xyxy, prm.pri_aux, p), gc);
                const unsigned long long ck = ((unsigned long long)ord_encode(val) << 32) | (unsigned long long)(0xffffffffu - (unsigned)p);
                best = ck > best ? ck : best;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long y = __shfl_xor_sync(FULL, best, o);
                best = y > best ? y : best;
            }
            if (lane == 0) atomicMax(&fs.col[g], best);
            role_sync<MT>();
            const int champ = (int)(0xffffffffu - (unsigned)(fs.col[g] & 0xffffffffull));
            if (mtid == 0) fs.champ[g] = champ;
            // rows: the champion's entry becomes 2.0 below (SFS:747 overwrites it); the others keep their CIoU
            for (int p = mtid; p < P; p += MT) {
                if (p == champ) continue;
                const float val = pair_ciou(load_prior(prm.pri_xyxy, prm.pri_aux, p), gc);
                if (val != val) atomicMax(&v.hi[2 * p], 0xffffffffu);
                else if (val >= thresh) { atomicMax(&v.hi[2 * p], ord_encode(val)); list_pair(p, g, ord_encode(val)); }
            }
        }
        role_sync<MT>();
    }
    // forced matches: row champ[g], column g = 2.0 (the lowest box index wins a shared champion: resolved with the list)
    if (mtid < G) {
        atomicMax(&v.hi[2 * fs.champ[mtid]], kOrdTwo);
        list_pair(fs.champ[mtid], mtid, kOrdTwo);
    }
    role_sync<MT>();
}

template <bool LOSS, int SRC, bool SHARE>
__global__ void __launch_bounds__(FT, 2) train_image_kernel(const TrainParams prm) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ FusedStatic fs;
    __shared__ HeadTable loc_tab, conf_tab;        // per-level bases of this image (head sources only)
    __shared__ PlaneRegions regions;               // (NCHW heads: the stream's plane-wise order)
    pdl_trigger();                                 // finalize_sums_kernel may be scheduled early (it waits for this grid)
    SSDHOT_STAMP(0);
    constexpr int MT = LOSS ? (SHARE ? MT_SHARE : MT_LOSS) : FT;       // threads that run the matching
    static_assert(MT % 32 == 0 && MT >= 256 && MT <= FT, "the matching prologue is spread over the first 204 threads");
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = prm.P;
    unsigned long long* table = reinterpret_cast<unsigned long long*>(dyn);
    unsigned char* scratch = dyn + (size_t)P * 8;
    unsigned* hist16 = reinterpret_cast<unsigned*>(scratch + FUSED_SCRATCH);       // [2048] 4096 bins x 16 bit (own region)
    FusedViews v;
    v.lo = reinterpret_cast<unsigned*>(table);
    v.hi = v.lo + 1;
    v.seg_n = reinterpret_cast<uint16_t*>(scratch);
    v.seg_info = reinterpret_cast<unsigned*>(scratch + 4112);
    v.chunk_list = reinterpret_cast<uint16_t*>(scratch + 4112 + 4096);
    v.pair_pg = reinterpret_cast<unsigned*>(scratch + 4112 + 8192);
    v.pair_ov = v.pair_pg + PAIR_CAP;
    v.work_list = reinterpret_cast<uint16_t*>(v.pair_ov + PAIR_CAP);
    // after the matching the scratch holds the lists of the mining
    // the positives are listed while the pair list is still being read: their list sits in the rectangle / chunk region
    uint16_t* pos_list = reinterpret_cast<uint16_t*>(scratch);                     // [POS_CAP]   (4096 B of the first 12304)
    uint16_t* sel_list = pos_list + POS_CAP;                                       // [SEL_CAP]   (filled after the pair list is dead)
    unsigned* band_v = reinterpret_cast<unsigned*>(scratch + 4112 + 8192);         // [BAND_CAP] exact CE bits
    unsigned* band_sorted = band_v + BAND_CAP;                                     // [BAND_CAP] the band's winners by rank
    uint16_t* band_p = reinterpret_cast<uint16_t*>(band_sorted + BAND_CAP);        // [BAND_CAP]
    static_assert(2 * (POS_CAP + SEL_CAP) <= 4112 + 8192 && 4112 + 8192 + 10 * BAND_CAP <= FUSED_SCRATCH, "mining lists fit the scratch");

    // the stream warps' first row pair is requested before anything else: its (cold) DRAM latency passes under the clear
    float x_first[12];
    if (LOSS && SRC == SRC_PACKED && tid >= MT && tid - MT < P / 2) {
        const HeadReader<SRC, 6> rd0 = {prm.conf_all + (long long)b * P * 6, nullptr};
        rd0.pair(tid - MT, x_first);
    }
    // (requested here, first used behind the clear -- and by the matching warps only: the stream does not wait for the boxes)
    const int g_begin = __ldg(prm.gt_offsets + b);
    const int G_given = __ldg(prm.gt_offsets + b + 1) - g_begin;

    // ---- 0. clear ------------------------------------------------------------------------------
    if (SRC != SRC_PACKED && LOSS) {
        head_table_fill<SRC, 4>(loc_tab, prm.loc_h, b, tid - 64);
        head_table_fill<SRC, 6>(conf_tab, prm.conf_h, b, tid - 96);
        if (SRC == SRC_LEVEL_PLANES) plane_regions_fill(regions, prm.conf_h, b, tid - 128);
    }
    if (tid == 0) { fs.first_nan = INT_MAX; fs.n_dense = 0; fs.n_band = 0; fs.n_pair = 0; fs.pair_overflow = 0; fs.n_pos_img = 0; fs.n_sure = 0; fs.n_pos_listed = 0; fs.n_sel = 0; }
    float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 32) sh = ldg4(prm.pri + 4ll * (kLevelOffset[kSeedLevel[tid]] + kSeedShape[tid]));   // (w, h) of this level-shape: the same in every cell (checked on the host)
    {
        ulonglong2* t2 = reinterpret_cast<ulonglong2*>(table);
        for (int i = tid; i < P / 2; i += FT) t2[i] = make_ulonglong2(0ull, 0ull);
        if ((P & 1) && tid == 0) table[P - 1] = 0ull;
        if (LOSS) for (int i = tid; i < 2048; i += FT) hist16[i] = 0u;
    }
    __syncthreads();                                        // table, histogram and head tables are ready: the stream starts
    const int G = min(G_given, prm.max_gt);
    if (tid < MT) {
        // what only the matching needs (the boxes arrive from HBM ~1 us after the offsets): behind its own barrier
        if (tid == 0 && G_given > prm.max_gt && prm.flags) atomicOr(prm.flags, 1);
        if (tid >= 32 && tid < 32 + G) {
            fs.gt_px[tid - 32] = ldg4(prm.gt_boxes + 4ll * (g_begin + tid - 32));
            fs.label[tid - 32] = (int)prm.gt_labels[g_begin + tid - 32];
        }
        if (tid < 32) {
            const int lv = kSeedLevel[tid], base = kLevelOffset[lv] + kSeedShape[tid];
            fs.cb_side[tid] = kLevelSide[lv]; fs.cb_shapes[tid] = kLevelShapes[lv]; fs.cb_base[tid] = base;
            fs.cb_w[tid] = sh.z; fs.cb_h[tid] = sh.w; fs.cb_inv[tid] = 1.0f / (float)kLevelSide[lv];
            fs.cb_hw[tid] = fmul(0.5f, sh.z); fs.cb_hh[tid] = fmul(0.5f, sh.w);
            fs.cb_coff[tid] = lv == 0 ? 0 : lv == 1 ? 38 : lv == 2 ? 57 : lv == 3 ? 67 : lv == 4 ? 72 : 75;
        }
        if (tid >= 128 && tid < 128 + 76) {                 // the centres of the six grids: fl32((2 i + 1) / (2 side))
            const int e = tid - 128;
            const int lv = (e >= 38) + (e >= 57) + (e >= 67) + (e >= 72) + (e >= 75);
            const int first = lv == 0 ? 0 : lv == 1 ? 38 : lv == 2 ? 57 : lv == 3 ? 67 : lv == 4 ? 72 : 75;
            fs.centre[e] = fdiv((float)(2 * (e - first) + 1), (float)(2 * kLevelSide[lv]));
        }
        role_sync<MT>();
    }
    const float thresh = prm.thresh;
    const HeadReader<SRC, 4> loc_rd = {SRC == SRC_PACKED && LOSS ? prm.loc_all + 4ll * b * P : nullptr, &loc_tab};
    const HeadReader<SRC, 6> conf_rd = {SRC == SRC_PACKED && LOSS ? prm.conf_all + (long long)b * P * 6 : nullptr, &conf_tab};

    // ---- 1, 2 (warps 0..7): matching  ||  3 (the other warps): approximate CE of every prior ----------------
    SSDHOT_STAMP(1);
    if (tid < MT) {
        match_phase<MT>(prm, fs, v, G, g_begin);
        SSDHOT_STAMP(3);
    } else if (LOSS) {
        const int n_pairs = P / 2;
        // key hand-off (eval step with one conf_all for both halves): flag 1 = this image's keys are on their way, 2 = delivered
        unsigned* share_keys = SHARE ? prm.share_keys + (size_t)b * share_stride_words(P) : nullptr;
        if (SHARE && tid == MT) st_release_gpu(prm.share_flag + b, 1);
        // (two instances of the loop, so that neither carries the other's branch)
        auto stream = [&](auto with_keys) {
            constexpr bool RK = decltype(with_keys)::value;
            auto row_pair = [&](const float* x, int q) {        // rows 2q, 2q + 1
                unsigned r0 = 0u, r1 = 0u;
                const unsigned k0 = __float_as_uint(approx_ce6_rk<RK>(x[0], x[1], x[2], x[3], x[4], x[5], r0));
                const unsigned k1 = __float_as_uint(approx_ce6_rk<RK>(x[6], x[7], x[8], x[9], x[10], x[11], r1));
                if (RK) share_keys[q] = r0 | (r1 << 16);
                v.lo[4 * q] = k0;
                v.lo[4 * q + 2] = k1;
                ce_hist_add(hist16, k0);
                ce_hist_add(hist16, k1);
            };
            if (SRC == SRC_LEVEL_PLANES) {
                // NCHW heads: plane-wise order (heads.cuh), a chunk of 32 cells per warp and iteration
                PlaneWalk walk(&regions);
#pragma unroll 2
                for (int k = (tid - MT) >> 5; k < kPlaneChunks; k += (FT - MT) >> 5) {
                    float x[12];
                    int p0;
                    if (walk.load(k, lane, x, p0)) row_pair(x, p0 >> 1);
                }
            } else {
                int q = tid - MT;
                if (SRC == SRC_PACKED) {                            // (requested at the top of the kernel)
                    if (q < n_pairs) row_pair(x_first, q);
                    q += FT - MT;
                }
#pragma unroll 2
                for (; q < n_pairs; q += FT - MT) {
                    float x[12];
                    conf_rd.pair(q, x);
                    row_pair(x, q);
                }
            }
        };
        stream(std::integral_constant<bool, SHARE>{});
        if ((P & 1) && tid == MT) {                          // odd P: the last row (never SSD300)
            float r[6];
            conf_rd.row(P - 1, r);
            const unsigned k = __float_as_uint(approx_ce6(r[0], r[1], r[2], r[3], r[4], r[5]));
            v.lo[2 * (P - 1)] = k;
            ce_hist_add(hist16, k);
        }
        if (SHARE) {                                         // every stream warp's keys are out: publish
            asm volatile("bar.sync 2, %0;" ::"n"(FT - MT) : "memory");
            if (tid == MT) { __threadfence(); st_release_gpu(prm.share_flag + b, 2); }
        }
        if (tid == FT - 32) SSDHOT_STAMP(2);                 // (the last stream warp)
    }
    __syncthreads();
    SSDHOT_STAMP(4);

    // ---- 2'. which box each positive prior matched ---------------------------------------------------
    const int first_nan = fs.first_nan;
    const unsigned ord_thr = ord_encode(thresh);
    // prior p is positive iff its best entry is a number >= thresh (an all-NaN column makes every row but 0 NaN)
    auto positive = [&](unsigned hi, int p) { return hi >= ord_thr && hi != 0xffffffffu && (first_nan == INT_MAX || p == 0); };
    // A positive's low word becomes 0x80000000 | (63 - box); atomicMax keeps the lowest box among equal maxima
    // (torch's arg-max).  The first claimer sees the old low word: the CE key the stream wrote, which leaves the histogram.
    auto claim = [&](int p, int g) {
        const unsigned old = atomicMax(&v.lo[2 * p], 0x80000000u | (unsigned)(63 - g));
        if (!(old >> 31)) {
            const int at = atomicAdd(&fs.n_pos_img, 1);
            if (LOSS) {
                if (at < POS_CAP) pos_list[at] = (uint16_t)p;      // the positives are listed here, not found again by a scan
                const int bin = ce_bin(old);
                atomicSub(&hist16[bin >> 1], 1u << ((bin & 1) * 16));
            }
        }
    };
    {
        const int n_list = min(fs.n_pair, PAIR_CAP);
        for (int e = tid; e < n_list; e += FT) {
            const unsigned pr = v.pair_pg[e];
            const int p = (int)(pr & 0xffffu);
            const unsigned hi = v.hi[2 * p];
            if (v.pair_ov[e] == hi && positive(hi, p)) claim(p, (int)(pr >> 16));
        }
    }
    __syncthreads();
    if (fs.pair_overflow) {
        // (rare: more gate survivors than the list holds) a positive whose winning pair was not listed recomputes it
        for (int p = tid; p < P; p += FT) {
            const unsigned hi = v.hi[2 * p];
            if (!positive(hi, p) || (v.lo[2 * p] >> 31)) continue;
            const BoxC pr = load_prior(prm.pri_xyxy, prm.pri_aux, p);
            for (int g = 0; g < G; ++g) {
                const unsigned o = fs.champ[g] == p ? kOrdTwo : ord_encode(pair_ciou(pr, gt_box(fs.gt_a[g], fs.gt_b[g])));
                if (o == hi) { claim(p, g); break; }
            }
        }
        __syncthreads();
    }
    auto matched_box = [&](unsigned lo) { return 63 - (int)(lo & 63u); };

    if (!LOSS) {
        // ---- build_targets-shaped outputs (SSD_trainer.py:547): masks, classes, positives' offsets ------
        int my_pos = 0;
        for (int p = tid; p < P; p += FT) {
            const unsigned lo = v.lo[2 * p];
            const bool pos = (lo >> 31) != 0u;
            const int g = matched_box(lo);
            const long long row = (long long)b * P + p;
            my_pos += pos ? 1 : 0;
            if (prm.pos_mask) prm.pos_mask[row] = pos ? 1 : 0;
            if (prm.cls_t) prm.cls_t[row] = pos ? (int64_t)(fs.label[g] + 1) : (int64_t)0;
            if (prm.code) prm.code[row] = pos ? (uint16_t)(g + 1) : (uint16_t)0;
            if (prm.loc_t && pos) {
                const float4 ga = fs.gt_a[g], gb = fs.gt_b[g];
                const float4 gbox = make_float4(gb.y, gb.z, fsub(ga.z, ga.x), fsub(ga.w, ga.y));
                reinterpret_cast<float4*>(prm.loc_t)[row] = encode_offsets(gbox, ldg4(prm.pri + 4ll * p), prm.inv_vc, prm.inv_vs);
            }
        }
        const int n_pos_img = block_sum<int>(my_pos, fs.ls.iscratch);
        if (tid == 0 && prm.n_pos) prm.n_pos[b] = n_pos_img;
        return;
    }

    SSDHOT_STAMP(5);
    // ---- 3'. hard-negative budget and bracket (the number of positives is the number of first claims) ------
    const int n_pos_img = fs.n_pos_img;
    if (tid == 0 && prm.n_pos) prm.n_pos[b] = n_pos_img;
    const long long n_neg = (long long)P - n_pos_img;
    long long want = (n_pos_img == 0) ? (long long)prm.ratio : (long long)(prm.ratio * (double)n_pos_img);
    if (want < 0) want = 0;
    const long long kk = want < n_neg ? want : n_neg;
    bool exact_all = kk >= n_neg && kk > 0;
    bool fast = kk > 0 && !exact_all;
    unsigned band_lo_key = 0xffffffffu, band_hi_key = 0xffffffffu;
    if (fast) {
        unsigned b1, above1;
        find_kth_from_top<FT, 4096>([&](int bin) { return (hist16[bin >> 1] >> ((bin & 1) * 16)) & 0xffffu; }, (unsigned)kk, fs, b1, above1);
        if (b1 == 0u || b1 == 4095u) { exact_all = true; fast = false; }   // the k-th value lies outside the resolved range: CTA-uniform
        else {
            const unsigned lo_key = (b1 + (unsigned)kCeBinBase) << 15, hi_key = lo_key | 0x7fffu;
            const float t_lo = __uint_as_float(lo_key), t_hi = __uint_as_float(hi_key);
            band_lo_key = __float_as_uint(fmaxf(t_lo - 3.0f * ce_error_bound(t_lo), 0.0f));
            band_hi_key = __float_as_uint(t_hi + 3.0f * ce_error_bound(t_hi));
        }
    }

    SSDHOT_STAMP(6);
    // ---- 4. one scan of the slots: the certainly-mined negatives and the few negatives inside the error band are listed (in
    //         whatever order the threads arrive; the positives were listed by the claims).  The loss terms are summed as
    //         64-bit fixed point, so the sums do not depend on that order: the same bits on every run and for every split of
    //         the batch.  Terms the fixed point cannot hold (>= 2^14, Inf, NaN) go to a double beside it.
    // Two fixed-point accumulators per sum: terms below 1 at 2^-48 (exact down to 2^-25; at most 8732 of them: < 2^62),
    // terms in [1, 2^14) at 2^-34 (exact: their ulp is >= 2^-23).  What neither holds (>= 2^14, Inf, NaN) goes to a double.
    struct Fx { long long small, big; double odd; };
    Fx fx_loc = {0, 0, 0.0}, fx_ce = {0, 0, 0.0};
    auto fx_add = [](Fx& a, float term) {
        if (term < 1.0f) a.small += __float2ll_rn(term * 281474976710656.0f);          // 2^48 (a power of two: the product is exact)
        else if (term < 16384.0f) a.big += __float2ll_rn(term * 17179869184.0f);       // 2^34
        else a.odd += (double)term;                                                    // (NaN fails both comparisons)
    };
    // a positive prior: exact CE of its class, smooth-L1 of its offsets (TR:108, :577-580)
    auto positive_terms = [&](int p, int g, bool with_ce) {
        if (with_ce) fx_add(fx_ce, exact_ce6(conf_rd, p, fs.label[g] + 1));
        const float4 ga = fs.gt_a[g], gb = fs.gt_b[g];
        const float4 gbox = make_float4(gb.y, gb.z, fsub(ga.z, ga.x), fsub(ga.w, ga.y));
        const float4 t = encode_offsets(gbox, ldg4(prm.pri + 4ll * p), prm.inv_vc, prm.inv_vs);
        float l[4];
        loc_rd.row(p, l);
        const float d[4] = {fsub(l[0], t.x), fsub(l[1], t.y), fsub(l[2], t.z), fsub(l[3], t.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float z = fabsf(d[j]);
            fx_add(fx_loc, z < 1.0f ? fmul(fmul(0.5f, z), z) : fsub(z, 0.5f));
        }
    };
    const bool pos_listed = n_pos_img <= POS_CAP;
    if (!pos_listed) { exact_all = true; fast = false; }               // (more positives than the list holds: the exact tail walks the slots)
    {
        const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(table);
        for (int q = tid; q < P / 2; q += FT) {                      // two adjacent slots per thread and trip (P is even here)
            const ulonglong2 sl = t2[q];
            const unsigned lo[2] = {(unsigned)sl.x, (unsigned)sl.y};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = 2 * q + h;
                const bool pos = (lo[h] >> 31) != 0u;
                if (prm.sel_cls) {
                    const long long row = (long long)b * P + p;
                    prm.sel_cls[row] = pos ? (int8_t)(fs.label[matched_box(lo[h])] + 1) : (int8_t)-1;
                    if (prm.matched16) prm.matched16[row] = pos ? (int16_t)matched_box(lo[h]) : (int16_t)-1;
                }
                if (fast && !pos && lo[h] >= band_lo_key) {          // (fast path off: nothing qualifies)
                    if (lo[h] > band_hi_key) {
                        const int at = atomicAdd(&fs.n_sel, 1);
                        if (at < SEL_CAP) sel_list[at] = (uint16_t)p;
                    } else {
                        const int at = atomicAdd(&fs.n_band, 1);
                        if (at < BAND_CAP) band_p[at] = (uint16_t)p;
                    }
                }
            }
        }
    }
    __syncthreads();
    SSDHOT_STAMP(7);
    {
        const int n_sure = fs.n_sel, n_band = fs.n_band;
        if (fast && (n_band > BAND_CAP || n_sure > SEL_CAP)) { exact_all = true; fast = false; }   // (massive ties): CTA-uniform
        // one pass over the lists: positives (cross-entropy) | positives (smooth-L1) | certain negatives | band members --
        // the two sides of a positive are the longest chains of the pass, so they go to different threads
        const int tot_pos = pos_listed ? n_pos_img : 0;
        const int tot_sel = fast ? n_sure : 0;
        const int tot_band = fast ? n_band : 0;
        for (int e = tid; e < 2 * tot_pos + tot_sel + tot_band; e += FT) {
            if (e < 2 * tot_pos) {
                const int p = (int)pos_list[e < tot_pos ? e : e - tot_pos];
                const int g = matched_box(v.lo[2 * p]);
                if (e < tot_pos) fx_add(fx_ce, exact_ce6(conf_rd, p, fs.label[g] + 1));
                else positive_terms(p, g, false);
            } else if (e < 2 * tot_pos + tot_sel) {
                const int p = (int)sel_list[e - 2 * tot_pos];
                fx_add(fx_ce, exact_ce6(conf_rd, p, 0));
                if (prm.sel_cls) prm.sel_cls[(long long)b * P + p] = 0;
            } else {
                const int eb = e - 2 * tot_pos - tot_sel;
                band_v[eb] = __float_as_uint(exact_ce6(conf_rd, (int)band_p[eb], 0)) & 0x7fffffffu;
            }
        }
        if (fast) {
            __syncthreads();
            const int r = (int)kk - n_sure;              // members still to take from the band (1 <= r <= n_band)
            for (int e = tid; e < n_band; e += FT) {
                const unsigned val = band_v[e];
                const int p = (int)band_p[e];
                int rank = 0;
                for (int j = 0; j < n_band; ++j) {
                    const unsigned vj = band_v[j];
                    rank += (vj > val || (vj == val && (int)band_p[j] < p)) ? 1 : 0;
                }
                if (rank < r) {
                    band_sorted[rank] = val;             // ranks are distinct: (value, prior) is a total order
                    if (prm.sel_cls) prm.sel_cls[(long long)b * P + p] = 0;
                }
            }
            __syncthreads();
            for (int e = tid; e < r; e += FT) fx_add(fx_ce, __uint_as_float(band_sorted[e]));
        }
    }
    SSDHOT_STAMP(8);
    double acc_ce = 0.0;
    if (exact_all) {
        // exact keys for every negative, then the exact radix selection (as loss_image_kernel); the cross-entropy side is
        // redone in prior order (double accumulation), the smooth-L1 side only for positives the list could not hold
        fx_ce = {0, 0, 0.0};
        __syncthreads();
        for (int i = tid; i < 256; i += FT) fs.ls.hist[i] = 0u;
        __syncthreads();
        for (int p = tid; p < P; p += FT) {
            const unsigned lo = v.lo[2 * p];
            if (!(lo >> 31)) {
                const unsigned key = __float_as_uint(exact_ce6(conf_rd, p, 0)) & 0x7fffffffu;
                v.lo[2 * p] = key;
                atomicAdd(&fs.ls.hist[key >> 24], 1u);
            } else {
                acc_ce += (double)exact_ce6(conf_rd, p, fs.label[matched_box(lo)] + 1);
                if (!pos_listed) positive_terms(p, matched_box(lo), false);
            }
        }
        __syncthreads();
        acc_ce += mined_exact_tail<FT>(prm, b, P, n_pos_img, fs.ls,
                                       [&](int p) { const unsigned lo = v.lo[2 * p]; return (lo >> 31) ? kNotNegative : lo; },
                                       [&](int p, int& mg) {
                                           const unsigned lo = v.lo[2 * p];
                                           mg = -1;
                                           if (!(lo >> 31)) return -1;
                                           mg = matched_box(lo);
                                           return fs.label[mg] + 1;
                                       });
    }

    // fixed point: integer addition is associative, so the order in which threads met the list entries does not matter.
    // One exchange for the six partial sums (warp sums -> shared memory -> the first warp), instead of six block sums.
    {
        long long* isum = reinterpret_cast<long long*>(fs.ls.hist);       // [4][FT / 32]  (the radix histogram is dead by now)
        const long long a0 = warp_sum(fx_loc.small), a1 = warp_sum(fx_loc.big), a2 = warp_sum(fx_ce.small), a3 = warp_sum(fx_ce.big);
        const double d0 = warp_sum(fx_loc.odd), d1 = warp_sum(fx_ce.odd + acc_ce);
        __syncthreads();
        if (lane == 0) {
            isum[warp] = a0; isum[FT / 32 + warp] = a1; isum[2 * (FT / 32) + warp] = a2; isum[3 * (FT / 32) + warp] = a3;
            fs.fin_odd[0][warp] = d0; fs.fin_odd[1][warp] = d1;
        }
        __syncthreads();
    }
    double s_loc = 0.0, s_ce = 0.0;
    if (warp == 0) {
        const long long* isum = reinterpret_cast<const long long*>(fs.ls.hist);
        const bool in = lane < FT / 32;
        const long long t0 = warp_sum(in ? isum[lane] : 0ll), t1 = warp_sum(in ? isum[FT / 32 + lane] : 0ll);
        const long long t2 = warp_sum(in ? isum[2 * (FT / 32) + lane] : 0ll), t3 = warp_sum(in ? isum[3 * (FT / 32) + lane] : 0ll);
        const double o0 = warp_sum(in ? fs.fin_odd[0][lane] : 0.0), o1 = warp_sum(in ? fs.fin_odd[1][lane] : 0.0);
        s_loc = (double)t0 * (1.0 / 281474976710656.0) + (double)t1 * (1.0 / 17179869184.0) + o0;
        s_ce = (double)t2 * (1.0 / 281474976710656.0) + (double)t3 * (1.0 / 17179869184.0) + o1;
    }
    if (tid == 0) {
        prm.img_part[2ll * b + 0] = s_loc;
        prm.img_part[2ll * b + 1] = s_ce;
        if (prm.timeline) { prm.timeline[(long long)b * 16 + 9] = globaltimer_ns(); unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); prm.timeline[(long long)b * 16 + 10] = sm; }
    }
}

// Fixed-order final reduction: sums[0..1] = sum of the per-CTA partials, sums[2] = sum n_pos.
__global__ void __launch_bounds__(256) finalize_sums_kernel(const double* __restrict__ cta_part, int n_part,
                                                            const int32_t* __restrict__ n_pos, int B,
                                                            double* __restrict__ sums) {
    __shared__ double scratch[32];
    pdl_wait();                                    // (launched with PDL: the producers' partial sums are complete from here on)
    double a = 0.0, c = 0.0, n = 0.0;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) { a += cta_part[2ll * i]; c += cta_part[2ll * i + 1]; }
    for (int i = threadIdx.x; i < B; i += blockDim.x) n += (double)n_pos[i];
    a = block_sum<double>(a, scratch);
    c = block_sum<double>(c, scratch);
    n = block_sum<double>(n, scratch);
    if (threadIdx.x == 0) { sums[0] = a; sums[1] = c; sums[2] = n; }
}

// per-prior constants of the clamped xyxy priors
__global__ void prior_tables_kernel(const float* __restrict__ pri, int P, float* __restrict__ xyxy, float* __restrict__ aux) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float4 c = ldg4(pri + 4ll * p);
    const float hw = fmul(0.5f, c.z), hh = fmul(0.5f, c.w);
    const float x1 = fminf(fmaxf(fsub(c.x, hw), 0.f), 1.f), y1 = fminf(fmaxf(fsub(c.y, hh), 0.f), 1.f);
    const float x2 = fminf(fmaxf(fadd(c.x, hw), 0.f), 1.f), y2 = fminf(fmaxf(fadd(c.y, hh), 0.f), 1.f);
    const BoxC k = box_consts(x1, y1, x2, y2, true);
    if (xyxy) reinterpret_cast<float4*>(xyxy)[p] = make_float4(x1, y1, x2, y2);
    if (aux) reinterpret_cast<float4*>(aux)[p] = make_float4(k.area, k.xc, k.yc, k.at);
}

// only aux, from caller-provided xyxy
__global__ void prior_aux_kernel(const float* __restrict__ xyxy, int P, float* __restrict__ aux) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float4 a = ldg4(xyxy + 4ll * p);
    const BoxC k = box_consts(a.x, a.y, a.z, a.w, true);
    reinterpret_cast<float4*>(aux)[p] = make_float4(k.area, k.xc, k.yc, k.at);
}

// loc_t[pos_mask]: one CTA per image, ordered compaction of the rows whose mask byte is set
__global__ void __launch_bounds__(256) compact_rows_kernel(const float* __restrict__ loc_t, const uint8_t* __restrict__ mask,
                                                           const int32_t* __restrict__ n_pos, int P, float* __restrict__ out) {
    __shared__ int warp_cnt[8];
    __shared__ long long base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        long long base = 0;
        for (int i = 0; i < b; ++i) base += n_pos[i];
        base_s = base;
    }
    __syncthreads();
    long long run = base_s;
    for (int p_base = 0; p_base < P; p_base += 256) {
        const int p = p_base + tid;
        const bool on = p < P && mask[(long long)b * P + p] != 0;
        const unsigned bal = __ballot_sync(FULL, on);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) before += warp_cnt[w]; total += warp_cnt[w]; }
        if (on) {
            const long long dst = run + before + __popc(bal & ((1u << lane) - 1u));
            reinterpret_cast<float4*>(out)[dst] = reinterpret_cast<const float4*>(loc_t)[(long long)b * P + p];
        }
        run += total;
        __syncthreads();
    }
}

// gradients of both losses w.r.t. the head outputs
template <int CT>
__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ pri, int P,
                                                       const float* __restrict__ gt_boxes, const int32_t* __restrict__ gt_offsets,
                                                       long long rows, float norm_w, float norm_h,
                                                       const float* __restrict__ loc_all, const float* __restrict__ conf_all, int C,
                                                       float inv_vc, float inv_vs,
                                                       const int8_t* __restrict__ sel, const int16_t* __restrict__ matched,
                                                       const double* __restrict__ scales,
                                                       float* __restrict__ g_loc, float* __restrict__ g_conf) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int s = sel[row];
    const float k_loc = (float)scales[0], k_conf = (float)scales[1];
    if (g_conf) {
        float* out = g_conf + row * C;
        if (s < 0) {
            for (int i = 0; i < C; ++i) out[i] = 0.0f;
        } else {
            const float* in = conf_all + row * C;
            float mx = __ldg(in);
            for (int i = 1; i < C; ++i) mx = fmaxf(mx, __ldg(in + i));
            float sum = 0.0f;
            for (int i = 0; i < C; ++i) sum += expf(__ldg(in + i) - mx);
            const float inv = 1.0f / sum;
            for (int i = 0; i < C; ++i) {
                const float pr = expf(__ldg(in + i) - mx) * inv;
                out[i] = k_conf * (pr - (i == s ? 1.0f : 0.0f));
            }
        }
    }
    if (g_loc) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = matched[row];
        if (m >= 0) {
            const int b = (int)(row / P), p = (int)(row % P);
            const float4 px = ldg4(gt_boxes + 4ll * (gt_offsets[b] + m));
            const float x1 = fdiv(px.x, norm_w), y1 = fdiv(px.y, norm_h), x2 = fdiv(px.z, norm_w), y2 = fdiv(px.w, norm_h);
            const float4 gbox = make_float4(fmul(fadd(x1, x2), 0.5f), fmul(fadd(y1, y2), 0.5f), fsub(x2, x1), fsub(y2, y1));
            const float4 t = encode_offsets(gbox, ldg4(pri + 4ll * p), inv_vc, inv_vs);
            const float4 l = ldg4(loc_all + 4ll * row);
            g.x = k_loc * fminf(fmaxf(l.x - t.x, -1.f), 1.f);
            g.y = k_loc * fminf(fmaxf(l.y - t.y, -1.f), 1.f);
            g.z = k_loc * fminf(fmaxf(l.z - t.z, -1.f), 1.f);
            g.w = k_loc * fminf(fmaxf(l.w - t.w, -1.f), 1.f);
        }
        reinterpret_cast<float4*>(g_loc)[row] = g;
    }
}


// loss_bwd_kernel on the head layouts (heads.cuh): reads the head outputs and writes the gradients in the same six
// tensors' layout, so autograd never needs a packed [B,P,D] gradient.  One thread per prior; in the NCHW layout the
// thread index enumerates (level, shape, cell) with the cell fastest, so a warp reads and writes runs of one plane.
__device__ __forceinline__ const float* head_pick(const HeadView& hv, int l) {
    const float* q = hv.base[0];
#pragma unroll
    for (int i = 1; i < kHeadLevels; ++i) if (l == i) q = hv.base[i];
    return q;
}
template <int SRC>
__global__ void __launch_bounds__(256) loss_bwd_heads_kernel(const float* __restrict__ pri,
                                                             const float* __restrict__ gt_boxes, const int32_t* __restrict__ gt_offsets,
                                                             int B, float norm_w, float norm_h,
                                                             const HeadView loc_h, const HeadView conf_h,
                                                             float inv_vc, float inv_vs,
                                                             const int8_t* __restrict__ sel, const int16_t* __restrict__ matched,
                                                             const double* __restrict__ scales,
                                                             const HeadView g_loc_h, const HeadView g_conf_h, int want_loc, int want_conf) {
    constexpr int P = 8732;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * P) return;
    const int b = (int)(t / P), u = (int)(t % P);
    const int l = head_level(u), r = u - head_level_off(l), hw = head_level_hw(l), A = head_level_six(l) ? 6 : 4;
    int cell, a;
    if (SRC == SRC_LEVEL_PLANES) { a = r / hw; cell = r - a * hw; }     // cell fastest across threads
    else { cell = r / A; a = r - cell * A; }
    const int p = head_level_off(l) + cell * A + a;
    const long long row = (long long)b * P + p;
    // element j of this prior's row in a branch with D values per prior: first + j * stride
    auto first = [&](int D) -> long long {
        return SRC == SRC_LEVEL_PLANES ? ((long long)b * A * D + a * D) * hw + cell : ((long long)b * hw * A + (p - head_level_off(l))) * D;
    };
    const int stride = SRC == SRC_LEVEL_PLANES ? hw : 1;
    const int s = sel[row];
    const float k_loc = (float)scales[0], k_conf = (float)scales[1];
    if (want_conf) {
        const long long f = first(6);
        float* out = const_cast<float*>(head_pick(g_conf_h, l)) + f;
        if (s < 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) out[(long long)i * stride] = 0.0f;
        } else {
            const float* in = head_pick(conf_h, l) + f;
            float x[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) x[i] = __ldg(in + (long long)i * stride);
            float mx = x[0];
#pragma unroll
            for (int i = 1; i < 6; ++i) mx = fmaxf(mx, x[i]);
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < 6; ++i) sum += expf(x[i] - mx);
            const float inv = 1.0f / sum;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float pr = expf(x[i] - mx) * inv;
                out[(long long)i * stride] = k_conf * (pr - (i == s ? 1.0f : 0.0f));
            }
        }
    }
    if (want_loc) {
        const long long f = first(4);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        const int m = matched[row];
        if (m >= 0) {
            const float4 px = ldg4(gt_boxes + 4ll * (gt_offsets[b] + m));
            const float x1 = fdiv(px.x, norm_w), y1 = fdiv(px.y, norm_h), x2 = fdiv(px.z, norm_w), y2 = fdiv(px.w, norm_h);
            const float4 gbox = make_float4(fmul(fadd(x1, x2), 0.5f), fmul(fadd(y1, y2), 0.5f), fsub(x2, x1), fsub(y2, y1));
            const float4 tg = encode_offsets(gbox, ldg4(pri + 4ll * p), inv_vc, inv_vs);
            const float* in = head_pick(loc_h, l) + f;
            const float tt[4] = {tg.x, tg.y, tg.z, tg.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = k_loc * fminf(fmaxf(__ldg(in + (long long)j * stride) - tt[j], -1.f), 1.f);
        }
        float* out = const_cast<float*>(head_pick(g_loc_h, l)) + f;
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(long long)j * stride] = g[j];
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

static int train_cluster_size(int P) {
    int cs = MAX_CS;
    while (cs > 1 && (cs / 2) * SLOTS >= P && P <= 2048) cs >>= 1;
    return cs;
}

template <bool PRUNE>
static int launch_match(const TrainParams& prm, cudaStream_t stream) {
    // a fixed cluster of 8 balances best at SSD300 sizes (1092 priors per CTA); small P shrink it
    const int cs = train_cluster_size(prm.P);
    if (cs * SLOTS < prm.P) return SSDHOT_ERR_SHAPE;
    if (prm.max_gt > 0) {
        const long long warps = (long long)prm.B * prm.max_gt;
        gt_prepare_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(prm, PRUNE ? 1 : 0);
        SSDHOT_CHECK_LAUNCH();
    }
    const size_t dyn = smem_bytes(prm.max_gt > 0 ? prm.max_gt : 1);
    auto kern = match_kernel<PRUNE>;
    if (dyn > 48 * 1024) {
        const int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), dyn);
        if (rc) return rc;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(prm.B * cs));
    cfg.blockDim = dim3(TT);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, prm);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    return SSDHOT_OK;
}

template <bool FROM_TARGETS>
static int launch_loss(const TrainParams& prm, cudaStream_t stream) {
    const size_t dyn = (size_t)prm.P * 4;       // <= 40 KB: no opt-in needed
    if (prm.C == 6) loss_image_kernel<6, FROM_TARGETS><<<prm.B, LT, dyn, stream>>>(prm);
    else loss_image_kernel<0, FROM_TARGETS><<<prm.B, LT, dyn, stream>>>(prm);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

template <bool LOSS, int SRC = SRC_PACKED>
static int launch_train_image(const TrainParams& prm_in, cudaStream_t stream) {
    TrainParams prm = prm_in;
    prm.timeline = g_timeline;
    const size_t dyn = fused_smem_bytes(prm.P);
    auto kern = train_image_kernel<LOSS, SRC, false>;
    if (LOSS && prm.share_keys) kern = train_image_kernel<LOSS, SRC, LOSS>;     // (<false, ., true> is never instantiated)
    {
        const int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), dyn);   // sticky per device; first raised by the warm-up call
        if (rc) return rc;
    }
    kern<<<prm.B, FT, dyn, stream>>>(prm);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

// the box-centric kernel needs the SSD300 grid structure, at most FAST_MAX_GT boxes per image and a positive threshold
static bool fast_path_ok(int prior_layout, int P, int max_gt, float iou_thresh) {
    return prior_layout == SSDHOT_LAYOUT_SSD300 && P == 8732 && max_gt <= FAST_MAX_GT && iou_thresh > 0.0f;
}

}  // namespace ssdhot

using namespace ssdhot;

// Host-side check that `priors` (HOST memory) have the SSD300 structure train_image_kernel relies on:
// levels of 38/19/10/5/3/1 cells with 4/6/6/6/4/4 shapes, cells row-major, shapes innermost
// (SSD_from_scratch.py:289-323); centres exactly fl32((i + 0.5)/side); one (w, h) in (0, 1] per (level, shape).
extern "C" int ssdhot_ssd300_layout_host(const float* priors_cxcywh_host, int P) {
    if (!priors_cxcywh_host || P != 8732) return 0;
    static const int side[6] = {38, 19, 10, 5, 3, 1}, shapes[6] = {4, 6, 6, 6, 4, 4};
    int off = 0;
    for (int l = 0; l < 6; ++l) {
        for (int j = 0; j < side[l]; ++j)
            for (int i = 0; i < side[l]; ++i)
                for (int k = 0; k < shapes[l]; ++k) {
                    const float* q = priors_cxcywh_host + 4ll * (off + (j * side[l] + i) * shapes[l] + k);
                    const float* q0 = priors_cxcywh_host + 4ll * (off + k);
                    // exactly the fp32 quotient the kernel rebuilds the priors from (grid_prior); SFS:318-325 satisfies it
                    const float cx = (float)(2 * i + 1) / (float)(2 * side[l]), cy = (float)(2 * j + 1) / (float)(2 * side[l]);
                    if (!(q[0] == cx && q[1] == cy)) return 0;
                    if (!(q[2] == q0[2] && q[3] == q0[3] && q[2] > 0.0f && q[2] <= 1.0f && q[3] > 0.0f && q[3] <= 1.0f)) return 0;
                }
        off += side[l] * side[l] * shapes[l];
    }
    return off == P ? 1 : 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int ssdhot_prior_tables(const float* priors_cxcywh, int P, float* priors_xyxy, float* prior_aux,
                                   ssdhot_stream_t stream) {
    if (!priors_cxcywh || (!priors_xyxy && !prior_aux)) return SSDHOT_ERR_NULL;
    if (P <= 0 || P > SSDHOT_MAX_PRIORS) return SSDHOT_ERR_SHAPE;
    if (!aligned16(priors_cxcywh) || !aligned16(priors_xyxy) || !aligned16(prior_aux)) return SSDHOT_ERR_ALIGN;
    prior_tables_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(priors_cxcywh, P, priors_xyxy, prior_aux);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

extern "C" int ssdhot_prior_aux(const float* priors_xyxy, int P, float* prior_aux, ssdhot_stream_t stream) {
    if (!priors_xyxy || !prior_aux) return SSDHOT_ERR_NULL;
    if (P <= 0 || P > SSDHOT_MAX_PRIORS) return SSDHOT_ERR_SHAPE;
    if (!aligned16(priors_xyxy) || !aligned16(prior_aux)) return SSDHOT_ERR_ALIGN;
    prior_aux_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(priors_xyxy, P, prior_aux);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

static int check_gt_args(const float* pri, const float* pri_xyxy, const float* aux, int P, const float* gt_boxes,
                         const int64_t* gt_labels, const int32_t* gt_offsets, int B, int max_gt,
                         float norm_w, float norm_h, float var_center, float var_size) {
    if (!pri || !pri_xyxy || !aux || !gt_offsets) return SSDHOT_ERR_NULL;
    if (max_gt > 0 && (!gt_boxes || !gt_labels)) return SSDHOT_ERR_NULL;
    if (P <= 0 || P > SSDHOT_MAX_PRIORS || B <= 0 || max_gt < 0 || max_gt > SSDHOT_MAX_GT) return SSDHOT_ERR_SHAPE;
    if (!(norm_w > 0.f) || !(norm_h > 0.f) || !(var_center > 0.f) || !(var_size > 0.f)) return SSDHOT_ERR_VALUE;
    if (!aligned16(pri) || !aligned16(pri_xyxy) || !aligned16(aux) || !aligned16(gt_boxes)) return SSDHOT_ERR_ALIGN;
    return SSDHOT_OK;
}

extern "C" int ssdhot_match_encode(const float* priors_cxcywh, const float* priors_xyxy, const float* prior_aux, int P,
                                   int prior_layout, const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                                   int B, int max_gt, float norm_w, float norm_h,
                                   float iou_thresh, float var_center, float var_size,
                                   float* loc_t, int loc_positives_only, int64_t* cls_t, uint8_t* pos_mask,
                                   int32_t* matched_gt, float* matched_cxcywh, int32_t* n_pos,
                                   int32_t* dev_flags, void* work, ssdhot_stream_t stream) {
    int rc = check_gt_args(priors_cxcywh, priors_xyxy, prior_aux, P, gt_boxes, gt_labels, gt_offsets, B, max_gt,
                           norm_w, norm_h, var_center, var_size);
    if (rc) return rc;
    if (!aligned16(loc_t) || !aligned16(matched_cxcywh) || !aligned16(work)) return SSDHOT_ERR_ALIGN;
    if (max_gt > 0 && !work) return SSDHOT_ERR_NULL;
    TrainParams prm = {};
    prm.gt_rec = reinterpret_cast<float4*>(work);
    prm.pri = priors_cxcywh; prm.pri_xyxy = priors_xyxy; prm.pri_aux = prior_aux; prm.P = P;
    prm.gt_boxes = gt_boxes; prm.gt_labels = gt_labels; prm.gt_offsets = gt_offsets;
    prm.B = B; prm.max_gt = max_gt; prm.norm_w = norm_w; prm.norm_h = norm_h;
    prm.thresh = iou_thresh; prm.inv_vc = 1.0f / var_center; prm.inv_vs = 1.0f / var_size;
    prm.loc_t = loc_t; prm.loc_pos_only = loc_positives_only; prm.cls_t = cls_t; prm.pos_mask = pos_mask;
    prm.matched32 = matched_gt; prm.matched_box = matched_cxcywh; prm.n_pos = n_pos; prm.flags = dev_flags;
    // the pruned sweep is exact for positives; negatives' matches need the exact-everywhere sweep
    const bool prune = (!loc_t || loc_positives_only) && !matched_gt && !matched_cxcywh;
    if (prune && fast_path_ok(prior_layout, P, max_gt, iou_thresh)) return launch_train_image<false>(prm, (cudaStream_t)stream);
    return prune ? launch_match<true>(prm, (cudaStream_t)stream) : launch_match<false>(prm, (cudaStream_t)stream);
}

extern "C" int ssdhot_compact_rows(const float* loc_t, const uint8_t* pos_mask, const int32_t* n_pos, int B, int P,
                                   float* out, ssdhot_stream_t stream) {
    if (!loc_t || !pos_mask || !n_pos || !out) return SSDHOT_ERR_NULL;
    if (B <= 0 || P <= 0) return SSDHOT_ERR_SHAPE;
    if (!aligned16(loc_t) || !aligned16(out)) return SSDHOT_ERR_ALIGN;
    compact_rows_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(loc_t, pos_mask, n_pos, P, out);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

// workspace layout: img_part [B][2] double | n_pos [B] int32 | gt_rec [B*max_gt][3] float4 | code [B,P] uint16
static size_t ws_npos_off(int B) { return (size_t)B * 2 * sizeof(double); }
static size_t ws_rec_off(int B) { return ws_npos_off(B) + (((size_t)B * sizeof(int32_t) + 15) & ~(size_t)15); }
static size_t ws_code_off(int B, int max_gt) { return ws_rec_off(B) + (size_t)B * (max_gt > 0 ? max_gt : 0) * 48; }

extern "C" unsigned long long ssdhot_match_workspace_bytes(int B, int max_gt) {
    if (B <= 0 || max_gt < 0) return 0;
    return (unsigned long long)((size_t)B * max_gt * 48 + 64);
}

extern "C" unsigned long long ssdhot_loss_workspace_bytes(int B, int P, int max_gt) {
    if (B <= 0 || P <= 0 || max_gt < 0) return 0;
    return (unsigned long long)(ws_code_off(B, max_gt) + (size_t)B * P * sizeof(uint16_t) + 64);
}

// ---- key hand-off between the halves of an eval step -----------------------------------------------------------------------
// SSD_test_step feeds ONE conf_all to the loss and to predict (SSD_trainer.py:208-256).  With a share buffer the loss kernel's
// logit stream also leaves predict's 16-bit row keys (B x P/2 words) and raises one flag per image; predict_image_kernel of the
// same step (on a second stream) picks them up instead of streaming conf_all again.  The flags are cleared by
// ssdhot_share_reset, issued before the two launches fork; a predict CTA whose image's keys are not on their way streams itself.
extern "C" unsigned long long ssdhot_share_bytes(int B, int P) {
    if (B <= 0 || P <= 0) return 0;
    return (unsigned long long)(share_flags_bytes(B) + (size_t)B * share_stride_words(P) * sizeof(unsigned));
}
extern "C" int ssdhot_share_reset(void* share, int B, ssdhot_stream_t stream) {
    if (!share) return SSDHOT_ERR_NULL;
    if (B <= 0) return SSDHOT_ERR_SHAPE;
    const cudaError_t e = cudaMemsetAsync(share, 0, share_flags_bytes(B), (cudaStream_t)stream);
    return e == cudaSuccess ? SSDHOT_OK : (int)e;
}
static void set_share(TrainParams& prm, void* share) {
    if (!share || (prm.P & 1) || !aligned16(share)) return;
    prm.share_flag = reinterpret_cast<int*>(share);
    prm.share_keys = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(share) + share_flags_bytes(prm.B));
}

static int finalize(const TrainParams& prm, const int32_t* n_pos, double* sums, cudaStream_t stream) {
    if (!sums) return SSDHOT_OK;        // the caller folds the per-image partials itself (ssdhot_allreduce_partials_peer)
    cudaError_t e = launch_pdl(finalize_sums_kernel, dim3(1), dim3(256), 0, stream, (const double*)prm.img_part, prm.B, n_pos, prm.B, sums);
    ++g_launches;
    if (e != cudaSuccess) return (int)e;
    return SSDHOT_OK;
}

extern "C" int ssdhot_multibox_loss_fwd(const float* priors_cxcywh, const float* priors_xyxy, const float* prior_aux, int P,
                                        int prior_layout, const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                                        int B, int max_gt, float norm_w, float norm_h,
                                        const float* loc_all, const float* conf_all, int C,
                                        float iou_thresh, float var_center, float var_size, double neg_pos_ratio,
                                        double* sums, void* work, int8_t* sel_cls, int16_t* matched_gt, int32_t* n_pos,
                                        int32_t* dev_flags, void* share, ssdhot_stream_t stream) {
    int rc = check_gt_args(priors_cxcywh, priors_xyxy, prior_aux, P, gt_boxes, gt_labels, gt_offsets, B, max_gt,
                           norm_w, norm_h, var_center, var_size);
    if (rc) return rc;
    if (!loc_all || !conf_all || !work) return SSDHOT_ERR_NULL;
    if (C < 2 || C > SSDHOT_MAX_CLASSES) return SSDHOT_ERR_SHAPE;
    if (sel_cls && C > SSDHOT_MAX_CLASSES_BWD) return SSDHOT_ERR_SHAPE;       // sel_cls holds the target class in an int8
    if (!(neg_pos_ratio >= 0.0)) return SSDHOT_ERR_VALUE;
    if (!aligned16(loc_all) || (reinterpret_cast<uintptr_t>(conf_all) & 7u) || !aligned16(work)) return SSDHOT_ERR_ALIGN;
    unsigned char* w = reinterpret_cast<unsigned char*>(work);
    TrainParams prm = {};
    prm.pri = priors_cxcywh; prm.pri_xyxy = priors_xyxy; prm.pri_aux = prior_aux; prm.P = P;
    prm.gt_boxes = gt_boxes; prm.gt_labels = gt_labels; prm.gt_offsets = gt_offsets;
    prm.B = B; prm.max_gt = max_gt; prm.norm_w = norm_w; prm.norm_h = norm_h;
    prm.thresh = iou_thresh; prm.inv_vc = 1.0f / var_center; prm.inv_vs = 1.0f / var_size;
    prm.loc_all = loc_all; prm.conf_all = conf_all; prm.C = C; prm.ratio = neg_pos_ratio;
    prm.img_part = reinterpret_cast<double*>(w);
    prm.gt_rec = reinterpret_cast<float4*>(w + ws_rec_off(B));
    prm.code = reinterpret_cast<uint16_t*>(w + ws_code_off(B, max_gt));
    prm.flags = dev_flags;
    int32_t* np = n_pos ? n_pos : reinterpret_cast<int32_t*>(w + ws_npos_off(B));
    if (C == 6 && aligned16(conf_all) && fast_path_ok(prior_layout, P, max_gt, iou_thresh)) {
        // SSD300 fast path: box-centric matching + mined loss, one kernel, one CTA per image
        prm.n_pos = np; prm.sel_cls = sel_cls; prm.matched16 = matched_gt; prm.code = nullptr;
        set_share(prm, share);
        rc = launch_train_image<true>(prm, (cudaStream_t)stream);
        if (rc) return rc;
        return finalize(prm, np, sums, (cudaStream_t)stream);
    }
    // generic priors / class counts:
    // 1) match: cluster per image, pruned sweep, 2-byte code per prior
    rc = launch_match<true>(prm, (cudaStream_t)stream);
    if (rc) return rc;
    // 2) loss: one CTA per image
    prm.n_pos = np; prm.sel_cls = sel_cls; prm.matched16 = matched_gt;
    rc = launch_loss<false>(prm, (cudaStream_t)stream);
    if (rc) return rc;
    return finalize(prm, np, sums, (cudaStream_t)stream);
}

// match + mined loss straight from the six head outputs of each branch (heads.cuh): the SSD300 fast path only.
extern "C" int ssdhot_multibox_loss_heads_fwd(const float* priors_cxcywh, const float* priors_xyxy, const float* prior_aux,
                                              int prior_layout, const float* gt_boxes, const int64_t* gt_labels,
                                              const int32_t* gt_offsets, int B, int max_gt, float norm_w, float norm_h,
                                              const float* const* loc_heads_host, const float* const* conf_heads_host,
                                              int head_layout, int C,
                                              float iou_thresh, float var_center, float var_size, double neg_pos_ratio,
                                              double* sums, void* work, int8_t* sel_cls, int16_t* matched_gt, int32_t* n_pos,
                                              int32_t* dev_flags, void* share, ssdhot_stream_t stream) {
    const int P = 8732;
    int rc = check_gt_args(priors_cxcywh, priors_xyxy, prior_aux, P, gt_boxes, gt_labels, gt_offsets, B, max_gt,
                           norm_w, norm_h, var_center, var_size);
    if (rc) return rc;
    if (!loc_heads_host || !conf_heads_host || !work) return SSDHOT_ERR_NULL;
    if (head_layout != SSDHOT_HEADS_NHWC && head_layout != SSDHOT_HEADS_NCHW) return SSDHOT_ERR_VALUE;
    if (!(neg_pos_ratio >= 0.0)) return SSDHOT_ERR_VALUE;
    // other class counts, priors or box counts: ssdhot_pack_heads + ssdhot_multibox_loss_fwd
    if (C != 6 || !fast_path_ok(prior_layout, P, max_gt, iou_thresh)) return SSDHOT_ERR_SHAPE;
    if (!aligned16(work)) return SSDHOT_ERR_ALIGN;
    unsigned char* w = reinterpret_cast<unsigned char*>(work);
    TrainParams prm = {};
    for (int l = 0; l < kHeadLevels; ++l) {
        if (!loc_heads_host[l] || !conf_heads_host[l]) return SSDHOT_ERR_NULL;
        if (!aligned16(loc_heads_host[l]) || !aligned16(conf_heads_host[l])) return SSDHOT_ERR_ALIGN;
        prm.loc_h.base[l] = loc_heads_host[l];
        prm.conf_h.base[l] = conf_heads_host[l];
    }
    prm.pri = priors_cxcywh; prm.pri_xyxy = priors_xyxy; prm.pri_aux = prior_aux; prm.P = P;
    prm.gt_boxes = gt_boxes; prm.gt_labels = gt_labels; prm.gt_offsets = gt_offsets;
    prm.B = B; prm.max_gt = max_gt; prm.norm_w = norm_w; prm.norm_h = norm_h;
    prm.thresh = iou_thresh; prm.inv_vc = 1.0f / var_center; prm.inv_vs = 1.0f / var_size;
    prm.C = C; prm.ratio = neg_pos_ratio;
    prm.img_part = reinterpret_cast<double*>(w);
    prm.gt_rec = reinterpret_cast<float4*>(w + ws_rec_off(B));
    prm.flags = dev_flags;
    int32_t* np = n_pos ? n_pos : reinterpret_cast<int32_t*>(w + ws_npos_off(B));
    prm.n_pos = np; prm.sel_cls = sel_cls; prm.matched16 = matched_gt; prm.code = nullptr;
    set_share(prm, share);
    rc = head_layout == SSDHOT_HEADS_NHWC ? launch_train_image<true, SRC_LEVEL_ROWS>(prm, (cudaStream_t)stream)
                                          : launch_train_image<true, SRC_LEVEL_PLANES>(prm, (cudaStream_t)stream);
    if (rc) return rc;
    return finalize(prm, np, sums, (cudaStream_t)stream);
}

extern "C" int ssdhot_mined_ce_fwd(const float* conf_all, const int64_t* cls_t, const uint8_t* pos_mask,
                                   int B, int P, int C, double neg_pos_ratio,
                                   double* sums, void* work, int8_t* sel_cls, ssdhot_stream_t stream) {
    if (!conf_all || !cls_t || !pos_mask || !sums || !work) return SSDHOT_ERR_NULL;
    if (P <= 0 || P > SSDHOT_MAX_PRIORS || B <= 0 || C < 2 || C > SSDHOT_MAX_CLASSES) return SSDHOT_ERR_SHAPE;
    if (sel_cls && C > SSDHOT_MAX_CLASSES_BWD) return SSDHOT_ERR_SHAPE;       // sel_cls holds the target class in an int8
    if (!(neg_pos_ratio >= 0.0)) return SSDHOT_ERR_VALUE;
    if ((reinterpret_cast<uintptr_t>(conf_all) & 7u) || !aligned16(work)) return SSDHOT_ERR_ALIGN;
    unsigned char* w = reinterpret_cast<unsigned char*>(work);
    TrainParams prm = {};
    prm.P = P; prm.B = B; prm.max_gt = 0; prm.conf_all = conf_all; prm.C = C; prm.ratio = neg_pos_ratio;
    prm.in_cls = cls_t; prm.in_pos = pos_mask;
    prm.img_part = reinterpret_cast<double*>(w);
    prm.n_pos = reinterpret_cast<int32_t*>(w + ws_npos_off(B));
    prm.sel_cls = sel_cls;
    int rc = launch_loss<true>(prm, (cudaStream_t)stream);
    if (rc) return rc;
    return finalize(prm, prm.n_pos, sums, (cudaStream_t)stream);
}

extern "C" int ssdhot_multibox_loss_bwd(const float* priors_cxcywh, int P,
                                        const float* gt_boxes, const int32_t* gt_offsets, int B,
                                        float norm_w, float norm_h,
                                        const float* loc_all, const float* conf_all, int C,
                                        float var_center, float var_size,
                                        const int8_t* sel_cls, const int16_t* matched_gt, const double* scales,
                                        float* grad_loc, float* grad_conf, ssdhot_stream_t stream) {
    if (!sel_cls || !scales || (!grad_loc && !grad_conf)) return SSDHOT_ERR_NULL;
    if (grad_conf && !conf_all) return SSDHOT_ERR_NULL;
    if (grad_loc && (!loc_all || !matched_gt || !priors_cxcywh || !gt_offsets)) return SSDHOT_ERR_NULL;
    if (P <= 0 || B <= 0 || C < 2 || C > SSDHOT_MAX_CLASSES_BWD) return SSDHOT_ERR_SHAPE;
    if (!aligned16(loc_all) || !aligned16(grad_loc) || !aligned16(priors_cxcywh) || !aligned16(gt_boxes)) return SSDHOT_ERR_ALIGN;
    const long long rows = (long long)B * P;
    loss_bwd_kernel<0><<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        priors_cxcywh, P, gt_boxes, gt_offsets, rows, norm_w, norm_h, loc_all, conf_all, C,
        1.0f / var_center, 1.0f / var_size, sel_cls, matched_gt, scales, grad_loc, grad_conf);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}

// ssdhot_multibox_loss_bwd on the head layouts: the head outputs in, the gradients out, both as six per-level tensors.
extern "C" int ssdhot_multibox_loss_heads_bwd(const float* priors_cxcywh, const float* gt_boxes, const int32_t* gt_offsets, int B,
                                              float norm_w, float norm_h,
                                              const float* const* loc_heads_host, const float* const* conf_heads_host,
                                              int head_layout, int C, float var_center, float var_size,
                                              const int8_t* sel_cls, const int16_t* matched_gt, const double* scales,
                                              float* const* grad_loc_heads_host, float* const* grad_conf_heads_host,
                                              ssdhot_stream_t stream) {
    if (!sel_cls || !scales || (!grad_loc_heads_host && !grad_conf_heads_host)) return SSDHOT_ERR_NULL;
    if (grad_conf_heads_host && !conf_heads_host) return SSDHOT_ERR_NULL;
    if (grad_loc_heads_host && (!loc_heads_host || !matched_gt || !priors_cxcywh || !gt_offsets || !gt_boxes)) return SSDHOT_ERR_NULL;
    if (head_layout != SSDHOT_HEADS_NHWC && head_layout != SSDHOT_HEADS_NCHW) return SSDHOT_ERR_VALUE;
    if (B <= 0 || C != 6) return SSDHOT_ERR_SHAPE;
    if (!aligned16(priors_cxcywh) || !aligned16(gt_boxes)) return SSDHOT_ERR_ALIGN;
    HeadView lh = {}, ch = {}, glh = {}, gch = {};
    for (int l = 0; l < kHeadLevels; ++l) {
        if (grad_loc_heads_host) {
            if (!loc_heads_host[l] || !grad_loc_heads_host[l]) return SSDHOT_ERR_NULL;
            lh.base[l] = loc_heads_host[l]; glh.base[l] = grad_loc_heads_host[l];
        }
        if (grad_conf_heads_host) {
            if (!conf_heads_host[l] || !grad_conf_heads_host[l]) return SSDHOT_ERR_NULL;
            ch.base[l] = conf_heads_host[l]; gch.base[l] = grad_conf_heads_host[l];
        }
    }
    const long long rows = (long long)B * 8732;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    const int wl = grad_loc_heads_host ? 1 : 0, wc = grad_conf_heads_host ? 1 : 0;
    if (head_layout == SSDHOT_HEADS_NHWC)
        loss_bwd_heads_kernel<SRC_LEVEL_ROWS><<<grid, 256, 0, (cudaStream_t)stream>>>(
            priors_cxcywh, gt_boxes, gt_offsets, B, norm_w, norm_h, lh, ch, 1.0f / var_center, 1.0f / var_size, sel_cls, matched_gt,
            scales, glh, gch, wl, wc);
    else
        loss_bwd_heads_kernel<SRC_LEVEL_PLANES><<<grid, 256, 0, (cudaStream_t)stream>>>(
            priors_cxcywh, gt_boxes, gt_offsets, B, norm_w, norm_h, lh, ch, 1.0f / var_center, 1.0f / var_size, sel_cls, matched_gt,
            scales, glh, gch, wl, wc);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}
