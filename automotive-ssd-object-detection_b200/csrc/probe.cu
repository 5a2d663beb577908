// probe.cu -- measurement aid, not part of the hot path: how fast can ONE CTA per image pull its 209,568 contiguous bytes of
// logits from HBM, (0) with the 16-byte read-only loads the kernels use (three LDG.128 per lane and row pair, two row pairs in
// flight per lane) or (1) with 1-D bulk-async copies (cp.async.bulk + mbarrier, the TMA path without a tensor map) into a
// shared-memory ring that all warps then read?  Both variants reduce the image to one checksum so that every byte is
// consumed.  tools/stream_probe.py times them; DESIGN.md section 4 records the outcome.
#include "common.cuh"

namespace ssdhot {

constexpr int PT_THREADS = 512;                 // as predict_image_kernel: 16 warps per image, two CTAs per SM
constexpr int PROBE_CHUNK = 16 * 32 * 48;       // one row pair (48 B) per lane, all 16 warps: 24,576 B
constexpr int PROBE_STAGES = 3;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(PT_THREADS, 2) stream_probe_kernel(const float* __restrict__ conf, long long image_bytes, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ float wsum[PT_THREADS / 32];
    __shared__ unsigned long long full[PROBE_STAGES], empty[PROBE_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char* img = reinterpret_cast<const unsigned char*>(conf) + (long long)blockIdx.x * image_bytes;
    float acc = 0.0f;
    if (MODE == 0) {
        // warp w owns a contiguous sixteenth of the image (as the kernels' list segments), 1536 B per warp and trip
        const long long seg = ((image_bytes / 16 + 47) / 48) * 48;            // bytes per warp, a multiple of 48
        const long long b0 = min(image_bytes, warp * seg), b1 = min(image_bytes, b0 + seg);
        const int n_it = (int)((b1 - b0 + 1535) / 1536);
        float4 xa[3], xb[3];
        auto load = [&](int it, float4* x) -> bool {
            const long long off = b0 + 1536ll * it + 48 * lane;
            if (it >= n_it || off + 48 > b1) return false;
            const float4* s = reinterpret_cast<const float4*>(img + off);
            x[0] = __ldg(s); x[1] = __ldg(s + 1); x[2] = __ldg(s + 2);
            return true;
        };
        auto use = [&](const float4* x, bool live) {
            if (live) acc += ((x[0].x + x[0].y) + (x[0].z + x[0].w)) + ((x[1].x + x[1].y) + (x[1].z + x[1].w)) + ((x[2].x + x[2].y) + (x[2].z + x[2].w));
        };
        bool la = load(0, xa), lb = false;
        for (int it = 0; it < n_it; it += 2) {
            lb = load(it + 1, xb);
            use(xa, la);
            la = load(it + 2, xa);
            use(xb, lb);
        }
    } else {
        if (tid == 0) {
            for (int s = 0; s < PROBE_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], PT_THREADS / 32); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const int n_chunk = (int)((image_bytes + PROBE_CHUNK - 1) / PROBE_CHUNK);
        auto issue = [&](int c) {                                             // (thread 0) chunk c -> stage c % STAGES
            const int s = c % PROBE_STAGES;
            const unsigned bytes = (unsigned)min((long long)PROBE_CHUNK, image_bytes - (long long)c * PROBE_CHUNK);
            mbar_expect_tx(&full[s], bytes);
            bulk_g2s(dyn + (size_t)s * PROBE_CHUNK, img + (long long)c * PROBE_CHUNK, bytes, &full[s]);
        };
        if (tid == 0) for (int c = 0; c < min(PROBE_STAGES, n_chunk); ++c) issue(c);
        for (int c = 0; c < n_chunk; ++c) {
            const int s = c % PROBE_STAGES;
            const unsigned phase = (unsigned)(c / PROBE_STAGES) & 1u;
            mbar_wait(&full[s], phase);
            const long long left = image_bytes - (long long)c * PROBE_CHUNK;
            const int off = 48 * tid;
            if (off + 48 <= left) {
                const float4* sp = reinterpret_cast<const float4*>(dyn + (size_t)s * PROBE_CHUNK + off);
                const float4 a = sp[0], b4 = sp[1], d = sp[2];
                acc += ((a.x + a.y) + (a.z + a.w)) + ((b4.x + b4.y) + (b4.z + b4.w)) + ((d.x + d.y) + (d.z + d.w));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (tid == 0 && c + PROBE_STAGES < n_chunk) {                      // refill the stage once every warp has left it
                mbar_wait(&empty[s], phase);
                issue(c + PROBE_STAGES);
            }
        }
    }
    acc = (float)warp_sum((double)acc);
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int w = 0; w < PT_THREADS / 32; ++w) t += wsum[w];
        out[blockIdx.x] = t;
    }
}

}  // namespace ssdhot

using namespace ssdhot;

extern "C" int ssdhot_debug_stream_probe(const float* conf, int B, long long image_bytes, int mode, float* out, ssdhot_stream_t stream) {
    if (!conf || !out) return SSDHOT_ERR_NULL;
    if (B <= 0 || image_bytes <= 0 || (image_bytes % 48) != 0) return SSDHOT_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(conf) & 15u) != 0) return SSDHOT_ERR_ALIGN;
    if (mode == 0) {
        stream_probe_kernel<0><<<B, PT_THREADS, 0, (cudaStream_t)stream>>>(conf, image_bytes, out);
    } else if (mode == 1) {
        const size_t dyn = (size_t)PROBE_STAGES * PROBE_CHUNK;
        int rc = ensure_dyn_smem(reinterpret_cast<const void*>(stream_probe_kernel<1>), dyn);
        if (rc) return rc;
        stream_probe_kernel<1><<<B, PT_THREADS, dyn, (cudaStream_t)stream>>>(conf, image_bytes, out);
    } else return SSDHOT_ERR_VALUE;
    SSDHOT_CHECK_LAUNCH();
    return SSDHOT_OK;
}
