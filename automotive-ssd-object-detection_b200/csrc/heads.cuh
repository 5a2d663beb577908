// heads.cuh -- where the streaming kernels find row p of a head branch.
//
// mySSD.forward (SSD_from_scratch.py:249-269) turns the six head outputs of a branch, NCHW tensors [B, A_l*D, H_l, W_l]
// (D = 4 box offsets or D = C logits), into loc_all / conf_all [B, 8732, D] with 12 permute().contiguous() and two cat.
// Row p of level l (p = off_l + (y*W_l + x)*A_l + a) is the D channels a*D .. a*D+D-1 of cell (y, x).  The kernels read
// that row from one of three sources, chosen at compile time:
//   SRC_PACKED       the reference's packed tensor [B, 8732, D]                          (row = 4*D contiguous bytes)
//   SRC_LEVEL_ROWS   six per-level tensors [B, H_l*W_l*A_l, D]: what permute(0,2,3,1) of a channels_last head output
//                    already is in memory -- rows are contiguous, only the two cats are skipped
//   SRC_LEVEL_PLANES six per-level NCHW tensors exactly as the conv heads return them: element j of row p lives at
//                    (b*A_l*D + a*D + j)*HW_l + cell; consecutive lanes read consecutive cells of a plane
// so that neither the transposes nor the cats ever touch HBM.  All level offsets and per-image level sizes of SSD300
// are even numbers of rows and multiples of 16 bytes, so a row pair never straddles two levels and the 16-byte loads
// of the packed path stay legal in SRC_LEVEL_ROWS.
#pragma once
#include "common.cuh"

namespace ssdhot {

constexpr int SRC_PACKED = 0, SRC_LEVEL_ROWS = 1, SRC_LEVEL_PLANES = 2;
constexpr int kHeadLevels = 6;

struct HeadView {                      // kernel parameter: the six tensors of one branch (device pointers)
    const float* base[kHeadLevels];
};

__device__ __forceinline__ int head_level(int p) {          // SSD300 level of prior p (a1 table: 0/5776/7942/8542/8692/8728)
    return (p >= 5776) + (p >= 7942) + (p >= 8542) + (p >= 8692) + (p >= 8728);
}
__device__ __forceinline__ int head_level_off(int l) { return l == 0 ? 0 : l == 1 ? 5776 : l == 2 ? 7942 : l == 3 ? 8542 : l == 4 ? 8692 : 8728; }
__device__ __forceinline__ int head_level_hw(int l) { return l == 0 ? 1444 : l == 1 ? 361 : l == 2 ? 100 : l == 3 ? 25 : l == 4 ? 9 : 1; }
__device__ __forceinline__ bool head_level_six(int l) { return l >= 1 && l <= 3; }            // 6 shapes per cell (else 4)

// Per-CTA table in shared memory (one image per CTA): filled by head_table_fill, read by HeadReader.
struct HeadTable {
    const float* vbase[kHeadLevels];
};

// SRC_LEVEL_ROWS: vbase[l] + p*D is row p.  SRC_LEVEL_PLANES: vbase[l] is the image's first plane of level l.
template <int SRC, int D>
__device__ __forceinline__ void head_table_fill(HeadTable& t, const HeadView& hv, int b, int tid) {
    if (SRC == SRC_PACKED) return;
    if ((unsigned)tid < (unsigned)kHeadLevels) {
        const int l = tid;
        const long long rows = (long long)head_level_hw(l) * (head_level_six(l) ? 6 : 4);
        const float* img = hv.base[0];
#pragma unroll
        for (int i = 1; i < kHeadLevels; ++i) if (l == i) img = hv.base[i];         // (no dynamic index into the parameter)
        img += (long long)b * rows * D;
        t.vbase[l] = SRC == SRC_LEVEL_ROWS ? img - (long long)head_level_off(l) * D : img;
    }
}

template <int SRC, int D>
struct HeadReader {
    const float* packed;               // SRC_PACKED: row 0 of the image
    const HeadTable* tab;

    // address pieces of row p in SRC_LEVEL_PLANES
    __device__ __forceinline__ const float* plane0(int p, int& hw) const {
        const int l = head_level(p);
        const unsigned r = (unsigned)(p - head_level_off(l));
        hw = head_level_hw(l);
        const unsigned cell = head_level_six(l) ? (__umulhi(r, 0xAAAAAAABu) >> 2) : (r >> 2);
        const unsigned a = r - cell * (head_level_six(l) ? 6u : 4u);
        return tab->vbase[l] + (long long)(a * D) * hw + cell;
    }
    __device__ __forceinline__ const float* row_ptr(int p) const {      // SRC_PACKED / SRC_LEVEL_ROWS
        return SRC == SRC_PACKED ? packed + (long long)p * D : tab->vbase[head_level(p)] + (long long)p * D;
    }
    // the D values of row p
    __device__ __forceinline__ void row(int p, float* x) const {
        if (SRC == SRC_LEVEL_PLANES) {
            int hw;
            const float* q = plane0(p, hw);
#pragma unroll
            for (int j = 0; j < D; ++j) x[j] = __ldg(q + (long long)j * hw);
        } else {
            const float* q = row_ptr(p);
            if (D == 4) {
                const float4 v = ldg4(q);
                x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            } else if (D == 6) {
                const float2 a = ldg2(q), b = ldg2(q + 2), c = ldg2(q + 4);
                x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y;
            } else {
#pragma unroll
                for (int j = 0; j < D; ++j) x[j] = __ldg(q + j);
            }
        }
    }
    // the 2*D values of rows 2q and 2q+1 (same cell, shapes a and a+1; 16-byte loads where rows are contiguous)
    __device__ __forceinline__ void pair(int q, float* x) const {
        if (SRC == SRC_LEVEL_PLANES) {
            int hw;
            const float* s = plane0(2 * q, hw);
#pragma unroll
            for (int j = 0; j < 2 * D; ++j) x[j] = __ldg(s + (long long)j * hw);
        } else {
            const float4* s = reinterpret_cast<const float4*>(row_ptr(2 * q));
#pragma unroll
            for (int j = 0; j < D / 2; ++j) {
                const float4 v = __ldg(s + j);
                x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
            }
        }
    }
};


// ---- SRC_LEVEL_PLANES streaming order -------------------------------------------------------------------------------
// The streaming kernels visit the row pairs (rows p, p+1 = shapes 2s, 2s+1 of one cell) of an image; in the NCHW layout
// they do it plane-wise: a "chunk" is 32 consecutive cells of one (level, shape pair) region, lane = cell, so each of the
// 12 loads of a chunk reads 128 contiguous bytes of one channel plane.  15 regions (level, s), 147 chunks per image:
//   level 0: 2 x 46 chunks | level 1: 3 x 12 | level 2: 3 x 4 | level 3: 3 x 1 | level 4: 2 x 1 | level 5: 2 x 1
constexpr int kPlaneChunks = 147;
__constant__ int kRegChunk0[16] = {0, 46, 92, 104, 116, 128, 132, 136, 140, 141, 142, 143, 144, 145, 146, 147};
__constant__ int kRegLevel[15] = {0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 5};
__constant__ int kRegPair[15] = {0, 1, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 0, 1};

// One (level, shape pair) region of an image's D = 6 branch, prepared once per CTA in shared memory.
struct PlaneRegion {
    const float* ptr0;     // cell 0 of the first plane of the pair's 12 channels
    int hw;                // cells = plane stride
    int pbase;             // first row of the pair in cell 0: level offset + 2 s
    int shapes;            // rows per cell (4 or 6)
    int chunk0, chunk1;    // the region's chunks are [chunk0, chunk1)
    int pad;
};
struct PlaneRegions { PlaneRegion r[15]; };

__device__ __forceinline__ void plane_regions_fill(PlaneRegions& t, const HeadView& hv, int b, int tid) {
    if ((unsigned)tid < 15u) {
        const int l = kRegLevel[tid], s = kRegPair[tid], hw = head_level_hw(l), shapes = head_level_six(l) ? 6 : 4;
        const float* img = hv.base[0];
#pragma unroll
        for (int i = 1; i < kHeadLevels; ++i) if (l == i) img = hv.base[i];
        PlaneRegion r;
        r.ptr0 = img + ((long long)b * shapes * 6 + 12 * s) * hw;
        r.hw = hw; r.pbase = head_level_off(l) + 2 * s; r.shapes = shapes;
        r.chunk0 = kRegChunk0[tid]; r.chunk1 = kRegChunk0[tid + 1]; r.pad = 0;
        t.r[tid] = r;
    }
}

// the 12 values of a row pair: plane stride known at compile time, so the loads take immediate offsets
template <int HW>
__device__ __forceinline__ void plane_load12_hw(const float* q, float* x) {
#pragma unroll
    for (int j = 0; j < 12; ++j) x[j] = __ldg(q + j * HW);
}
__device__ __forceinline__ void plane_load12(const float* q, int hw, float* x) {       // hw is warp-uniform
    if (hw == 1444) plane_load12_hw<1444>(q, x);
    else if (hw == 361) plane_load12_hw<361>(q, x);
    else if (hw == 100) plane_load12_hw<100>(q, x);
    else if (hw == 25) plane_load12_hw<25>(q, x);
    else if (hw == 9) plane_load12_hw<9>(q, x);
    else plane_load12_hw<1>(q, x);
}

// Walks the chunks k0, k0 + step, ... of an image (warp-uniform); load() fetches this lane's pair of the current chunk.
struct PlaneWalk {
    const PlaneRegions* t;
    int reg;
    __device__ __forceinline__ PlaneWalk(const PlaneRegions* tab) : t(tab), reg(0) {}
    // -> this lane's cell exists; p0 = first row of its pair.  (A lane past the end of the region reads the region's last
    // cell, so the loads need no predicate.)
    __device__ __forceinline__ bool load(int k, int lane, float* x, int& p0) {
        while (k >= t->r[reg].chunk1) ++reg;
        const PlaneRegion r = t->r[reg];
        const int cell = (k - r.chunk0) * 32 + lane;
        p0 = r.pbase + cell * r.shapes;
        plane_load12(r.ptr0 + min(cell, r.hw - 1), r.hw, x);
        return cell < r.hw;
    }
};

}  // namespace ssdhot
