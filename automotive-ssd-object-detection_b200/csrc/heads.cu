// heads.cu -- head-output packing (first step).
//
// mySSD.forward ends with 12 permute(0,2,3,1).contiguous() calls and two torch.cat (SSD_from_scratch.py:249-269)
// to turn the six NCHW head outputs [B, A*D, H, W] of a branch (D = 4 box offsets or D = C class logits) into
// loc_all / conf_all [B, 8732, D].  pack_heads_kernel does one branch in ONE launch: every (image, level, tile) CTA
// transposes a 32 x 32 tile [channel k][cell hw] -> [cell hw][channel k] through padded shared memory, so both the
// NCHW reads and the [B, P, D] writes are coalesced and every byte is read and written once.
#include "common.cuh"

namespace ssdhot {

constexpr int kLevels = 6;
struct PackParams {
    const float* head[kLevels];     // [B, K_l, HW_l]
    int hw[kLevels];                // cells of the level
    int k[kLevels];                 // channels of the level = shapes per cell * D
    int out_off[kLevels];           // offset (in floats) of the level inside one image's output row
    int tile_begin[kLevels + 1];    // prefix of tiles per image over the levels
    int tiles_x[kLevels];           // tiles along hw
    int per_image;                  // floats of one image's output = 8732 * D
    float* out;
};

__global__ void __launch_bounds__(256) pack_heads_kernel(const PackParams prm) {
    __shared__ float tile[32][33];
    const int b = blockIdx.y;
    int t = blockIdx.x, l = 0;
    while (t >= prm.tile_begin[l + 1]) ++l;
    t -= prm.tile_begin[l];
    const int hw0 = (t % prm.tiles_x[l]) * 32, k0 = (t / prm.tiles_x[l]) * 32;
    const int HW = prm.hw[l], K = prm.k[l];
    const float* in = prm.head[l] + (long long)b * K * HW;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {                       // rows = channels, columns = cells (contiguous in NCHW)
        const int k = k0 + r, hw = hw0 + tx;
        if (k < K && hw < HW) tile[r][tx] = __ldg(in + (long long)k * HW + hw);
    }
    __syncthreads();
    float* out = prm.out + (long long)b * prm.per_image + prm.out_off[l];
#pragma unroll
    for (int r = ty; r < 32; r += 8) {                       // rows = cells, columns = channels (contiguous in the output)
        const int hw = hw0 + r, k = k0 + tx;
        if (k < K && hw < HW) out[(long long)hw * K + k] = tile[tx][r];
    }
}

}  // namespace ssdhot

using namespace ssdhot;

// heads_host: HOST array of the six DEVICE pointers of one branch, level order 38/19/10/5/3/1 (SFS:249-262).
extern "C" int ssdhot_pack_heads(const float* const* heads_host, int B, int D, float* out, ssdhot_stream_t stream) {
    if (!heads_host || !out) return SSDHOT_ERR_NULL;
    if (B <= 0 || B > 65535 || D <= 0 || D > SSDHOT_MAX_CLASSES) return SSDHOT_ERR_SHAPE;
    static const int side[kLevels] = {38, 19, 10, 5, 3, 1}, shapes[kLevels] = {4, 6, 6, 6, 4, 4};
    PackParams prm = {};
    int off = 0, tiles = 0;
    for (int l = 0; l < kLevels; ++l) {
        if (!heads_host[l]) return SSDHOT_ERR_NULL;
        prm.head[l] = heads_host[l];
        prm.hw[l] = side[l] * side[l];
        prm.k[l] = shapes[l] * D;
        prm.out_off[l] = off;
        prm.tiles_x[l] = (prm.hw[l] + 31) / 32;
        prm.tile_begin[l] = tiles;
        tiles += prm.tiles_x[l] * ((prm.k[l] + 31) / 32);
        off += prm.hw[l] * prm.k[l];
    }
    prm.tile_begin[kLevels] = tiles;
    prm.per_image = off;                                      // = 8732 * D
    prm.out = out;
    pack_heads_kernel<<<dim3((unsigned)tiles, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(prm);
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}
