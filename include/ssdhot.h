/*
 * ssdhot.h -- C ABI of libssdhot.so, the B200 (sm_100a) implementation of the SSD300 multibox
 * post-backbone hot path of ElliotBlackstone/automotive-ssd-object-detection.
 *
 * The reference has no FFI of its own: the boundary it offers is a set of Python callables
 * (SURVEY.md section 8b).  Each entry point below replaces the device work behind one of them;
 * the citation says which (SFS = SSD_from_scratch.py, TR = SSD_trainer.py, tv = torchvision/ops).
 * INTEGRATION.md shows the ctypes stubs a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all
 *     buffers, nothing is allocated, freed or cached inside the library;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *     synchronisation and is safe to capture in a CUDA graph;
 *   - return value: 0 = SSDHOT_OK, negative = ssdhot_status (bad argument), positive =
 *     cudaError_t of the failed launch;
 *   - thread-compatible: concurrent calls must use different streams and different output buffers;
 *   - floating point is fp32 with IEEE round-to-nearest +,-,*,/ in the operation order of the
 *     reference (no FMA contraction, no fast-math); labels are int64, masks are one byte per
 *     element (torch.bool layout);
 *   - P (number of priors) must be <= SSDHOT_MAX_PRIORS; the reference uses 8732.
 */
#ifndef SSDHOT_H_
#define SSDHOT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SSDHOT_API __attribute__((visibility("default")))
#else
#define SSDHOT_API
#endif

#define SSDHOT_ABI_VERSION 2
#define SSDHOT_MAX_PRIORS 10240   /* 8 CTAs x 256 threads x 5 priors per image cluster */
#define SSDHOT_MAX_GT 2048        /* ground-truth boxes per image held in shared memory */
#define SSDHOT_MAX_CLASSES 256
#define SSDHOT_MAX_CLASSES_BWD 127   /* with sel_cls (the backward's int8 class record): SSDHOT_ERR_SHAPE beyond it */

typedef enum {
    SSDHOT_OK = 0,
    SSDHOT_ERR_NULL = -1,        /* a required pointer is NULL */
    SSDHOT_ERR_SHAPE = -2,       /* a size is out of the supported range */
    SSDHOT_ERR_VALUE = -3,       /* a scalar violates the reference's own validation */
    SSDHOT_ERR_DEVICE = -4,      /* not an sm_100 device */
    SSDHOT_ERR_ALIGN = -5        /* a pointer is not aligned as documented (16 B for [.,4] fp32) */
} ssdhot_status;

typedef void* ssdhot_stream_t;   /* cudaStream_t */

/* prior_layout argument of the matching entry points: 0 = arbitrary priors (generic kernels);
 * 1 = the SSD300 grid structure of SFS:289-323 (levels 38/19/10/5/3/1, shapes innermost), as
 * verified once by ssdhot_ssd300_layout_host -- enables the box-centric fast path.  Passing 1 for
 * priors that do not have that structure gives wrong matches. */
#define SSDHOT_LAYOUT_GENERIC 0
#define SSDHOT_LAYOUT_SSD300 1

/* metric used by the NMS predicates: the reference CODE uses DIoU (SFS:688); CIoU is offered
 * because README.md:17 / BASELINE.json name it. */
#define SSDHOT_METRIC_DIOU 0
#define SSDHOT_METRIC_CIOU 1
#define SSDHOT_METRIC_IOU 2

SSDHOT_API int ssdhot_abi_version(void);
/* debug: device buffer [B][16] uint64 receiving %globaltimer stamps of ssdhot_multibox_loss_fwd's fused kernel
 * phases (NULL switches it off, the default); tools/timeline.py prints them */
SSDHOT_API int ssdhot_debug_timeline(void* dev_buffer);

/* Measurement aid (csrc/probe.cu): one CTA per image pulls `image_bytes` contiguous bytes (a multiple of 48, `conf` 16-byte
 * aligned) and writes one checksum per image to out [B]; mode 0 = the 16-byte read-only loads the kernels use, mode 1 =
 * 1-D bulk-async copies (cp.async.bulk + mbarrier) through a shared-memory ring.  tools/stream_probe.py times both. */
SSDHOT_API int ssdhot_debug_stream_probe(const float* conf, int B, long long image_bytes, int mode, float* out,
                                         ssdhot_stream_t stream);
SSDHOT_API const char* ssdhot_status_string(int status);
/* number of kernels this library has launched so far in this process (bench.py gpu_launches) */
SSDHOT_API unsigned long long ssdhot_launch_count(void);

/* ---- a1: per-prior constants -------------------------------------------------------------
 * priors_xyxy [P,4] = clamp(cxcywh -> xyxy, 0, 1) as registered at SFS:34-35
 * (tv _box_convert.py:5-24); prior_aux [P,4] = (area, x centre, y centre, atan(w/h)) of the
 * clamped box, the per-row constants of complete_box_iou (tv boxes.py:296, :470-471, :430). */
SSDHOT_API int ssdhot_prior_tables(const float* priors_cxcywh, int P, float* priors_xyxy, float* prior_aux,
                        ssdhot_stream_t stream);
/* prior_aux only, from caller-owned (already clamped) xyxy priors such as mySSD.priors_xyxy. */
SSDHOT_API int ssdhot_prior_aux(const float* priors_xyxy, int P, float* prior_aux, ssdhot_stream_t stream);
/* HOST function on a HOST copy of the priors: 1 if they have the SSD300 structure
 * (SSDHOT_LAYOUT_SSD300), else 0.  No GPU work. */
SSDHOT_API int ssdhot_ssd300_layout_host(const float* priors_cxcywh_host, int P);

/* ---- a2/a3: match + encode ----------------------------------------------------------------
 * Replaces the per-image loop of build_targets (TR:525-545) around mySSD.encode_ssd
 * (SFS:697-773): CIoU of every prior against the image's ground truth, forced best-prior
 * match, best ground truth per prior, positives, centre-size offsets, class targets.
 *   gt_boxes   [sumG,4] xyxy, divided in-kernel by (norm_w,norm_h,norm_w,norm_h) (TR:519,532);
 *              pass 1,1 for already-normalised boxes (the encode_ssd signature)
 *   gt_labels  [sumG] int64 foreground ids 0..C-2
 *   gt_offsets [B+1] int32, image b owns rows gt_offsets[b] .. gt_offsets[b+1]
 *   max_gt     host-known upper bound of boxes per image (<= SSDHOT_MAX_GT)
 * Outputs (each may be NULL to skip it):
 *   loc_t [B,P,4] (every prior, like encode_ssd; or only positives' rows when
 *   loc_positives_only != 0, the other rows are left untouched), cls_t [B,P] int64,
 *   pos_mask [B,P] bytes, matched_gt [B,P] int32 (index inside the image),
 *   matched_cxcywh [B,P,4], n_pos [B] int32.
 * An image whose box count exceeds max_gt sets bit 0 of *dev_flags (optional).
 *   work: scratch of ssdhot_match_workspace_bytes(B, max_gt) bytes. */
SSDHOT_API int ssdhot_match_encode(const float* priors_cxcywh, const float* priors_xyxy, const float* prior_aux, int P,
                        int prior_layout, const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                        int B, int max_gt, float norm_w, float norm_h,
                        float iou_thresh, float var_center, float var_size,
                        float* loc_t, int loc_positives_only, int64_t* cls_t, uint8_t* pos_mask,
                        int32_t* matched_gt, float* matched_cxcywh, int32_t* n_pos,
                        int32_t* dev_flags, void* work, ssdhot_stream_t stream);
/* scratch for ssdhot_match_encode (per-box records: 48 bytes each), 16-byte aligned */
SSDHOT_API unsigned long long ssdhot_match_workspace_bytes(int B, int max_gt);

/* Ordered compaction loc_t[pos_mask] (TR:547): rows of loc_t [B,P,4] whose mask byte is set, in
 * (image, prior) order, written to out[sum(n_pos),4].  n_pos [B] as written by
 * ssdhot_match_encode. */
SSDHOT_API int ssdhot_compact_rows(const float* loc_t, const uint8_t* pos_mask, const int32_t* n_pos, int B, int P,
                        float* out, ssdhot_stream_t stream);

/* ---- a4/a5: losses ------------------------------------------------------------------------
 * Fused forward of the post-backbone training step (TR:92-117): matching as above, smooth-L1
 * over positives (TR:108), softmax cross-entropy over positives plus the hardest
 * k = min(int(ratio*n_pos) [int(ratio) if n_pos == 0], #negatives) negatives of each image
 * (TR:577-598).  Nothing of shape [B,P] is materialised unless asked for.
 *   sums [3] double, OVERWRITTEN: { sum smooth-L1, sum CE (positives + mined), sum n_pos } --
 *   un-normalised, ready for one all-reduce; loss = sums[0..1] / max(sums[2], 1) (TR:105,600).
 *   work: scratch of ssdhot_loss_workspace_bytes(B, P, max_gt) bytes, 16-byte aligned.
 *   sums may be NULL: the final reduction is then left to ssdhot_allreduce_partials_peer (the partials stay in `work`).
 *   Optional, for the backward pass: sel_cls [B,P] int8 (-1 = prior not in the loss, else its
 *   target class), matched_gt [B,P] int16 (positives only, -1 elsewhere), n_pos [B]. */
SSDHOT_API unsigned long long ssdhot_loss_workspace_bytes(int B, int P, int max_gt);
SSDHOT_API int ssdhot_multibox_loss_fwd(const float* priors_cxcywh, const float* priors_xyxy, const float* prior_aux, int P,
                             int prior_layout, const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_offsets,
                             int B, int max_gt, float norm_w, float norm_h,
                             const float* loc_all, const float* conf_all, int C,
                             float iou_thresh, float var_center, float var_size, double neg_pos_ratio,
                             double* sums, void* work, int8_t* sel_cls, int16_t* matched_gt, int32_t* n_pos,
                             int32_t* dev_flags, void* share, ssdhot_stream_t stream);

/* ---- key hand-off inside an eval step (SSD_test_step, TR:208-256: ONE conf_all feeds the loss and predict) ------------
 * `share` (device, ssdhot_share_bytes(B, P) bytes, 16-byte aligned; NULL = none) ties a ssdhot_multibox_loss_fwd /
 * _heads_fwd launch to the ssdhot_predict_stages / _heads launch of the SAME logits issued on a second stream: the loss
 * kernel's logit stream also leaves predict's 16-bit row keys there and raises one flag per image, and the predict kernel
 * picks them up instead of reading conf_all a second time (SSD300 fast path and C == 6 on both sides; ignored otherwise --
 * a predict CTA whose keys are not on their way streams the logits itself, so results never depend on it).  Results are
 * identical with and without.  Per step: ssdhot_share_reset on the stream BEFORE the two launches fork (it clears the
 * flags), then both launches with the same `share`; the next step's reset must be ordered after both. */
SSDHOT_API unsigned long long ssdhot_share_bytes(int B, int P);
SSDHOT_API int ssdhot_share_reset(void* share, int B, ssdhot_stream_t stream);

/* CELoss_w_neg_mining (TR:551-600) with the targets given: sums[1] and sums[2] are written
 * (sums[0] = 0).  cls_t [B,P] int64, pos_mask [B,P] bytes. */
SSDHOT_API int ssdhot_mined_ce_fwd(const float* conf_all, const int64_t* cls_t, const uint8_t* pos_mask,
                        int B, int P, int C, double neg_pos_ratio,
                        double* sums, void* work, int8_t* sel_cls, ssdhot_stream_t stream);

/* Backward of the two losses w.r.t. the head outputs (autograd of TR:108 and TR:577-600):
 *   grad_conf [B,P,C] = scale_conf * (softmax(conf) - onehot(sel_cls)) where sel_cls >= 0, else 0
 *   grad_loc  [B,P,4] = scale_loc * clamp(loc - loc_t, -1, 1) on positives, else 0
 * scale_* = upstream gradient / max(total positives, 1), read from scales [2] (device, double).
 * Positives' loc targets are re-encoded from matched_gt (int16) and the ground truth.
 * Either gradient may be NULL. */
SSDHOT_API int ssdhot_multibox_loss_bwd(const float* priors_cxcywh, int P,
                             const float* gt_boxes, const int32_t* gt_offsets, int B,
                             float norm_w, float norm_h,
                             const float* loc_all, const float* conf_all, int C,
                             float var_center, float var_size,
                             const int8_t* sel_cls, const int16_t* matched_gt, const double* scales,
                             float* grad_loc, float* grad_conf, ssdhot_stream_t stream);

/* ---- a6: decode ---------------------------------------------------------------------------
 * mySSD.decode_ssd (SFS:776-800): out [M,4] cxcywh from loc [M,4] and priors [M,4]. */
SSDHOT_API int ssdhot_decode(const float* loc, const float* priors_cxcywh, int M, float var_center, float var_size,
                  float* out, ssdhot_stream_t stream);

/* ---- a8: greedy NMS -----------------------------------------------------------------------
 * mySSD.iou_nms (SFS:664-692) for `n_sets` independent box sets stored back to back:
 * set s owns boxes[set_offsets[s] .. set_offsets[s+1]) (xyxy) and the matching scores.
 * A box is dropped iff NOT(metric(kept, box) <= thresh) for an earlier kept box, boxes visited by
 * descending score, equal scores by ascending index.  keep [total] int64 receives, per set and
 * starting at set_offsets[s], the surviving indices (relative to the set) in that order;
 * keep_count [n_sets] int32 their number.  max_keep > 0 stops each set after that many survivors.
 * work: ssdhot_nms_workspace_bytes(total) bytes. */
SSDHOT_API unsigned long long ssdhot_nms_workspace_bytes(long long total_boxes);
SSDHOT_API int ssdhot_nms(const float* boxes, const float* scores, const int32_t* set_offsets, int n_sets,
               long long total_boxes, int max_set_size, float thresh, int metric, int max_keep,
               int64_t* keep, int32_t* keep_count, void* work, ssdhot_stream_t stream);

/* ---- a7: post-process ---------------------------------------------------------------------
 * mySSD.predict after the forward pass (SFS:388-476): softmax, strict score threshold on every
 * (prior, foreground class) pair, decode + clamp + scale to pixels of the survivors, greedy
 * NMS per class (or over all classes when class_agnostic != 0), global score sort and
 * truncation to max_per_img.  Outputs are padded to max_per_img per image:
 *   out_labels [B,max_per_img] int64 (0-based foreground id), out_scores [B,max_per_img],
 *   out_boxes [B,max_per_img,4] pixel xyxy, out_cand [B,max_per_img] int32 (optional: flat
 *   candidate id prior*(C-1)+class), out_count [B] int32 valid entries per image.
 * work: ssdhot_predict_workspace_bytes(B, P, C) bytes (16 candidate-list segments per image), 16-byte aligned.
 * SSD300 with C == 6 runs as ONE kernel (predict_image_kernel: the stream keeps one 16-bit key per row in shared memory, the
 * candidates of the round come from the few "hot" rows; `work` is touched only by images that need more than one round);
 * other shapes run the two-kernel path below.  Same results either way. */
SSDHOT_API unsigned long long ssdhot_predict_workspace_bytes(int B, int P, int C);
SSDHOT_API int ssdhot_predict(const float* priors_cxcywh, int P, const float* loc_all, const float* conf_all,
                   int B, int C, float score_thresh, float nms_thresh, int max_per_img,
                   int class_agnostic, int metric, float var_center, float var_size,
                   float img_w, float img_h,
                   int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                   int32_t* out_count, void* work, ssdhot_stream_t stream);

/* ---- f.3 (first step): head-output packing ---------------------------------------------------------
 * The tail of mySSD.forward (SFS:249-269): the six NCHW head outputs [B, shapes_l * D, side_l, side_l] of one branch
 * (D = 4 box offsets, or D = C class logits; levels 38/19/10/5/3/1 with 4/6/6/6/4/4 shapes) -> out [B, 8732, D], the
 * layout every other entry point reads -- one launch instead of 6 permute().contiguous() + view + cat, every byte read
 * and written once.  heads_host is a HOST array of the six DEVICE pointers. */
SSDHOT_API int ssdhot_pack_heads(const float* const* heads_host, int B, int D, float* out, ssdhot_stream_t stream);

/* The two stages of ssdhot_predict separately (same arguments): SSDHOT_STAGE_SCORES streams the logits and fills the
 * candidate lists in `work` (score_kernel, the HBM-bound stage); SSDHOT_STAGE_NMS ranks them and runs the greedy NMS
 * (nms_image_kernel) -- it consumes the lists, so it needs a fresh SCORES stage before every call.  ssdhot_predict is
 * stages = SCORES | NMS (for SSD300 / C == 6 that request takes the one-kernel path instead).  Used by bench.py to time the
 * generic path's streaming kernel against the HBM roofline on its own.  `share`: see ssdhot_share_bytes (NULL = none). */
#define SSDHOT_STAGE_SCORES 1
#define SSDHOT_STAGE_NMS 2
SSDHOT_API int ssdhot_predict_stages(const float* priors_cxcywh, int P, const float* loc_all, const float* conf_all,
                          int B, int C, float score_thresh, float nms_thresh, int max_per_img,
                          int class_agnostic, int metric, float var_center, float var_size,
                          float img_w, float img_h,
                          int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                          int32_t* out_count, void* work, int stages, void* share, ssdhot_stream_t stream);

/* The hot path straight from the head outputs (SSD_from_scratch.py:249-269): instead of the packed
 * loc_all / conf_all, the six per-level tensors of each branch (levels 38/19/10/5/3/1 with 4/6/6/6/4/4 shapes; HOST
 * arrays of six DEVICE pointers, each 16-byte aligned).  head_layout:
 *   SSDHOT_HEADS_NCHW  [B, A_l*D, H_l, W_l] contiguous, exactly what the conv heads return (SFS:249-262);
 *   SSDHOT_HEADS_NHWC  [B, H_l, W_l, A_l*D] contiguous = permute(0,2,3,1) of a channels_last head output.
 * Neither the 12 permute().contiguous() nor the 2 cat of mySSD.forward run: the streaming kernels address row p of
 * level l directly.  SSD300 layout (P = 8732) and C == 6 only (SSDHOT_ERR_SHAPE otherwise: use ssdhot_pack_heads and the
 * packed entry points).  Results are bit-identical to the packed entry points on the packed tensors.  `work` as for
 * ssdhot_predict (ssdhot_predict_workspace_bytes(B, 8732, C)); all other arguments as ssdhot_predict_stages. */
#define SSDHOT_HEADS_NCHW 0
#define SSDHOT_HEADS_NHWC 1
SSDHOT_API int ssdhot_predict_heads(const float* priors_cxcywh, const float* const* loc_heads_host,
                         const float* const* conf_heads_host, int head_layout, int B, int C,
                         float score_thresh, float nms_thresh, int max_per_img,
                         int class_agnostic, int metric, float var_center, float var_size,
                         float img_w, float img_h,
                         int64_t* out_labels, float* out_scores, float* out_boxes, int32_t* out_cand,
                         int32_t* out_count, void* work, int stages, void* share, ssdhot_stream_t stream);

/* ssdhot_multibox_loss_fwd from the head outputs (see ssdhot_predict_heads for the two layouts).  SSD300 fast path only:
 * prior_layout == SSDHOT_LAYOUT_SSD300, C == 6, max_gt <= 64, iou_thresh > 0 -- SSDHOT_ERR_SHAPE otherwise (pack the heads
 * and call ssdhot_multibox_loss_fwd).  sums / sel_cls / matched_gt / n_pos are bit-identical to ssdhot_multibox_loss_fwd on
 * the packed tensors; `work` as there (ssdhot_loss_workspace_bytes(B, 8732, max_gt)). */
SSDHOT_API int ssdhot_multibox_loss_heads_fwd(const float* priors_cxcywh, const float* priors_xyxy, const float* prior_aux,
                         int prior_layout, const float* gt_boxes, const int64_t* gt_labels,
                         const int32_t* gt_offsets, int B, int max_gt, float norm_w, float norm_h,
                         const float* const* loc_heads_host, const float* const* conf_heads_host,
                         int head_layout, int C,
                         float iou_thresh, float var_center, float var_size, double neg_pos_ratio,
                         double* sums, void* work, int8_t* sel_cls, int16_t* matched_gt, int32_t* n_pos,
                         int32_t* dev_flags, void* share, ssdhot_stream_t stream);

/* ssdhot_multibox_loss_bwd on the head layouts: reads the head outputs, writes the gradients as six per-level tensors
 * of the same layout as the inputs (HOST arrays of DEVICE pointers; either gradient array may be NULL).  sel_cls /
 * matched_gt [B,8732] as written by ssdhot_multibox_loss_heads_fwd.  Values equal ssdhot_multibox_loss_bwd's. */
SSDHOT_API int ssdhot_multibox_loss_heads_bwd(const float* priors_cxcywh, const float* gt_boxes, const int32_t* gt_offsets, int B,
                         float norm_w, float norm_h,
                         const float* const* loc_heads_host, const float* const* conf_heads_host,
                         int head_layout, int C, float var_center, float var_size,
                         const int8_t* sel_cls, const int16_t* matched_gt, const double* scales,
                         float* const* grad_loc_heads_host, float* const* grad_conf_heads_host,
                         ssdhot_stream_t stream);

/* ---- e: the sharded path's exchange over NVLink peer memory -----------------------------------------------------
 * The only collective of the image-sharded path is the all-reduce (sum) of sums[3] = [sum smooth-L1, sum CE, sum
 * positives] (SSD_trainer.py:105,108,600).  ssdhot_allreduce_sums_peer does it in ONE 32-thread kernel over peer
 * memory: every rank owns a mailbox (ssdhot_peer_alloc) that the other ranks of the node map through CUDA IPC
 * (ssdhot_peer_export on the owner, ssdhot_peer_open on the others; the 64-byte handles travel over any host channel,
 * e.g. torch.distributed.all_gather).  The kernel stores this rank's sums into every mailbox, waits for the others'
 * and adds them in rank order, so all ranks end with identical bits; it keeps its step counter in device memory and can
 * therefore be captured in a CUDA graph and replayed.  mailboxes_host: HOST array of `world` DEVICE pointers as mapped
 * in the calling process (entry `rank` = the local mailbox).  lag = 0: sums becomes the all-reduced sums of this call.
 * lag = 1: sums becomes the all-reduced sums of the PREVIOUS call (zeros on the first call) -- the slots it collects were
 * posted a step earlier, so the kernel does not wait and the ranks are not re-synchronised every step.  Every rank must
 * issue the same number of calls with the same lag; a rank that waits more than ~10 s sets bit 2 of *dev_flags (if given)
 * and produces NaN.  world <= SSDHOT_PEER_MAX_RANKS. */
#define SSDHOT_PEER_MAX_RANKS 8
SSDHOT_API unsigned long long ssdhot_peer_mailbox_bytes(void);
SSDHOT_API int ssdhot_peer_alloc(void** mailbox_out);
SSDHOT_API int ssdhot_peer_free(void* mailbox);
SSDHOT_API int ssdhot_peer_export(const void* mailbox, void* handle64_host);
SSDHOT_API int ssdhot_peer_open(const void* handle64_host, void** mailbox_out);
SSDHOT_API int ssdhot_peer_close(void* mapped_mailbox);
SSDHOT_API int ssdhot_allreduce_sums_peer(double* sums, void* const* mailboxes_host, int rank, int world, int lag, int32_t* dev_flags,
                               ssdhot_stream_t stream);
/* The same exchange fed by the per-image partial sums of a loss forward that was called with sums == NULL (it then skips
 * its own final reduction): loss_work = that call's `work`, B = its batch, n_pos = its n_pos (NULL: the workspace copy).
 * The kernel folds the partials (fixed order) and exchanges them: one dependent launch less per step. */
SSDHOT_API int ssdhot_allreduce_partials_peer(const void* loss_work, int B, const int32_t* n_pos, double* sums,
                                              void* const* mailboxes_host, int rank, int world, int lag, int32_t* dev_flags,
                                              ssdhot_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSDHOT_H_ */
