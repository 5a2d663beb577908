// common.cuh -- shared device helpers for libssdhot (sm_100a only).
//
// Arithmetic rule of the whole library: every +,-,*,/ of the reference is a separately rounded
// fp32 ATen op, so the kernels spell each one with the __f*_rn intrinsics, which nvcc never
// contracts into FMA.  Transcendentals (atanf, expf, logf) are the CUDA libdevice ones, i.e.
// the same functions eager torch-CUDA kernels call.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

#include "../../include/ssdhot.h"

namespace cg = cooperative_groups;

namespace ssdhot {

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libssdhot is written for sm_100a (B200) only"
#endif

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

constexpr unsigned FULL = 0xffffffffu;

// Order-preserving map fp32 -> u32 (larger float <=> larger integer); every NaN maps to the top.
__device__ __forceinline__ unsigned ord_encode(float v) {
    unsigned u = __float_as_uint(v);
    u ^= (u & 0x80000000u) ? 0xffffffffu : 0x80000000u;
    return (v != v) ? 0xffffffffu : u;
}
__device__ __forceinline__ float ord_decode(unsigned u) {
    u ^= (u & 0x80000000u) ? 0x80000000u : 0xffffffffu;
    return __uint_as_float(u);
}

// ---- key hand-off between the halves of an eval step (ssdhot_share_bytes): [B flags, 256-byte padded | B x stride words] ----
__host__ __device__ inline int share_stride_words(int P) { return ((P / 2) + 3) & ~3; }
__host__ __device__ inline size_t share_flags_bytes(int B) { return ((size_t)B * 4 + 255) & ~(size_t)255; }
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(FULL, v); }
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Block-wide sums in a fixed (deterministic) order.  `scratch` holds >= 32 elements; the result
// is returned to every thread.  Contains two __syncthreads().
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    // every warp folds the <= 32 partials in the same butterfly order: identical in all threads
    return warp_sum(lane < nwarp ? scratch[lane] : T(0));
}

// 128-bit read-only loads.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor on the
// stream is still running; it must call pdl_wait() before touching anything the predecessor writes.  The
// predecessor calls pdl_trigger() once the dependent grid may be scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

int ensure_dyn_smem(const void* kern, size_t bytes);   // per-(device, kernel) opt-in to > 48 KB of dynamic shared memory (abi.cu)
extern unsigned long long g_launches;   // host-side counter (abi.cu)
extern unsigned long long* g_timeline;  // debug hook (ssdhot_debug_timeline, abi.cu): device buffer [units][16] or null

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

}  // namespace ssdhot

#define SSDHOT_CHECK_LAUNCH()                              \
    do {                                                   \
        ++ssdhot::g_launches;                              \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)
