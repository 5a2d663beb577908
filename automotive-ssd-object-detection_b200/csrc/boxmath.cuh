// boxmath.cuh -- pairwise IoU / DIoU / CIoU in exactly the operation order of torchvision
// (ops/boxes.py:273-341 box_area/_box_inter_union, :344-371 box_iou, :462-480 _box_diou_iou,
// :437-459 distance_box_iou, :404-434 complete_box_iou), one rounding per reference op.
//
// "row" is boxes1 (the prior in matching, the kept higher-score box in NMS), "col" is boxes2.
// Per-box constants (area, centre, atan(w/h)) are computed once per box by box_consts(); they
// are the [N,1] / [1,M] tensors torchvision forms before broadcasting.
#pragma once
#include "common.cuh"

namespace ssdhot {

struct __align__(16) BoxC {   // one box + its broadcast constants
    float x1, y1, x2, y2;
    float area, xc, yc, at;
};

constexpr float kEps = 1e-7f;                      // eps of distance/complete_box_iou, cast to fp32
constexpr float kFourOverPiSq = 0.4052847345693511f;  // float(4 / pi**2)   (boxes.py:430)

__device__ __forceinline__ BoxC box_consts(float x1, float y1, float x2, float y2, bool want_atan) {
    BoxC b;
    b.x1 = x1; b.y1 = y1; b.x2 = x2; b.y2 = y2;
    const float w = fsub(x2, x1), h = fsub(y2, y1);
    b.area = fmul(w, h);                              // boxes.py:296
    b.xc = fmul(fadd(x1, x2), 0.5f);                  // :470  (x1 + x2) / 2  (exact either way)
    b.yc = fmul(fadd(y1, y2), 0.5f);                  // :471
    b.at = want_atan ? atanf(fdiv(w, h)) : 0.0f;      // :430  atan(w / h)
    return b;
}

// IoU; also returns the overlap extents' product.
__device__ __forceinline__ float pair_iou(const BoxC& r, const BoxC& c) {
    const float w = fmaxf(fsub(fminf(r.x2, c.x2), fmaxf(r.x1, c.x1)), 0.0f);   // :321-322, :336
    const float h = fmaxf(fsub(fminf(r.y2, c.y2), fmaxf(r.y1, c.y1)), 0.0f);
    const float inter = fmul(w, h);                                            // :337
    const float uni = fsub(fadd(r.area, c.area), inter);                       // :339
    return fdiv(inter, uni);                                                   // :370
}

__device__ __forceinline__ float pair_diou_from_iou(const BoxC& r, const BoxC& c, float iou) {
    const float ex = fmaxf(fsub(fmaxf(r.x2, c.x2), fminf(r.x1, c.x1)), 0.0f);  // :465-467
    const float ey = fmaxf(fsub(fmaxf(r.y2, c.y2), fminf(r.y1, c.y1)), 0.0f);
    const float diag2 = fadd(fadd(fmul(ex, ex), fmul(ey, ey)), kEps);          // :468
    const float dx = fsub(r.xc, c.xc), dy = fsub(r.yc, c.yc);
    const float dist2 = fadd(fmul(dx, dx), fmul(dy, dy));                      // :475-477
    return fsub(iou, fdiv(dist2, diag2));                                      // :480
}

__device__ __forceinline__ float pair_diou(const BoxC& r, const BoxC& c) {
    return pair_diou_from_iou(r, c, pair_iou(r, c));
}

__device__ __forceinline__ float pair_ciou(const BoxC& r, const BoxC& c) {
    const float iou = pair_iou(r, c);
    const float diou = pair_diou_from_iou(r, c, iou);
    const float da = fsub(r.at, c.at);
    const float v = fmul(kFourOverPiSq, fmul(da, da));                         // :430
    const float alpha = fdiv(v, fadd(fadd(fsub(1.0f, iou), v), kEps));         // :432
    return fsub(diou, fmul(alpha, v));                                         // :433
}

template <int METRIC>
__device__ __forceinline__ float pair_metric(const BoxC& r, const BoxC& c) {
    if (METRIC == SSDHOT_METRIC_DIOU) return pair_diou(r, c);
    if (METRIC == SSDHOT_METRIC_CIOU) return pair_ciou(r, c);
    return pair_iou(r, c);
}

// Centre-size offsets of ground truth g (cx, cy, w, h) w.r.t. prior p (cx, cy, w, h)
// (SSD_from_scratch.py:759-762).  `/ v` with a Python scalar is executed by eager torch-CUDA as a
// multiplication by fl(1/v) (ATen BinaryDivTrueKernel), which is what inv_vc / inv_vs hold.
__device__ __forceinline__ float4 encode_offsets(float4 g, float4 p, float inv_vc, float inv_vs) {
    float4 t;
    t.x = fmul(fdiv(fsub(g.x, p.x), p.z), inv_vc);
    t.y = fmul(fdiv(fsub(g.y, p.y), p.w), inv_vc);
    t.z = fmul(logf(fmaxf(fdiv(g.z, p.z), 1e-12f)), inv_vs);
    t.w = fmul(logf(fmaxf(fdiv(g.w, p.w), 1e-12f)), inv_vs);
    return t;
}

// mySSD.decode_ssd (SSD_from_scratch.py:793-797): cx = ((l0*v_c)*pw)+pcx, w = pw*exp(l2*v_s).
__device__ __forceinline__ float4 decode_box(float4 l, float4 p, float vc, float vs) {
    float4 o;
    o.x = fadd(fmul(fmul(l.x, vc), p.z), p.x);
    o.y = fadd(fmul(fmul(l.y, vc), p.w), p.y);
    o.z = fmul(p.z, expf(fmul(l.z, vs)));
    o.w = fmul(p.w, expf(fmul(l.w, vs)));
    return o;
}

// torch.clamp(v, 0, 1): NaN propagates (ATen clamp = min(max(v, lo), hi) with NaN-propagating min / max), unlike fminf / fmaxf.
__device__ __forceinline__ float clamp01_nan(float v) {
    const float c = fminf(fmaxf(v, 0.0f), 1.0f);
    return (v != v) ? v : c;
}

// cxcywh (normalised) -> clamped pixel xyxy (SSD_from_scratch.py:422-425).  A NaN coordinate stays NaN (its box then has a
// NaN area, fails no IoU gate and suppresses / is suppressed exactly as in the reference, SFS:690).
__device__ __forceinline__ float4 to_pixel_xyxy(float4 c, float img_w, float img_h) {
    const float hw = fmul(0.5f, c.z), hh = fmul(0.5f, c.w);
    float4 o;
    o.x = fmul(clamp01_nan(fsub(c.x, hw)), img_w);
    o.y = fmul(clamp01_nan(fsub(c.y, hh)), img_h);
    o.z = fmul(clamp01_nan(fadd(c.x, hw)), img_w);
    o.w = fmul(clamp01_nan(fadd(c.y, hh)), img_h);
    return o;
}

// Softmax pieces in the order of eager torch-CUDA's persistent warp softmax (dim <= 1024:
// one element per lane for C <= 32, lanes = next_pow2(C), butterfly add over xor offsets
// lanes/2 .. 1, masked lanes contribute 0).  `e` holds exp(x_i - max) for i < C (C <= 32).
__device__ __forceinline__ float softmax_denominator(const float* e, int C) {
    int lanes = 1;
    while (lanes < C) lanes <<= 1;
    float part[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) part[i] = (i < C) ? e[i] : 0.0f;
    for (int off = lanes >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < off) part[i] = fadd(part[i], part[i + off]);   // lane i's value after the xor step
    }
    return part[0];
}

}  // namespace ssdhot
