// Shared helpers for libvqcpc_b200: error reporting, warp primitives, the low-latency ("LL")
// cross-SM exchange used by the persistent recurrent kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqcpc.h"

namespace vqcpc {

void set_error(const char* fmt, ...);

#define VQ_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            vqcpc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                        \
            return VQCPC_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

#define VQ_ARG(cond, ...)                \
    do {                                 \
        if (!(cond)) {                   \
            vqcpc::set_error(__VA_ARGS__); \
            return VQCPC_ERR_ARG;        \
        }                                \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Programmatic dependent launch (latency paths: chains of small kernels on one stream).  A kernel launched with launch_pdl() may
// start while its predecessor is still running; pdl_sync() at its top -- before the first access to anything the predecessor
// touches -- waits for the predecessor's completion and memory flush, then lets ITS successor be scheduled, so every kernel's
// launch latency and prologue run under the kernel before it.  In a kernel launched normally both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// sigmoid / tanh from ex2.approx: abs error ~1e-7, saturate correctly for large |x|.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, __fadd_rn(1.0f, __expf(-x))); }
__device__ __forceinline__ float tanh_fast(float x) {
    return __fsub_rn(1.0f, __fdividef(2.0f, __fadd_rn(1.0f, __expf(__fmul_rn(2.0f, x)))));
}

// Named barriers (ids 1..15; 0 is __syncthreads).  arrive = producer side, sync = consumer side.
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// LL exchange: every value travels as one 8-byte word {tag:32 | float:32}.  An aligned 8-byte
// store is single-copy atomic, so data and flag arrive together: no fence, one L2 round trip.
// Tags are step numbers (>= 1); the buffer is zeroed before launch; slots are double-buffered by
// step parity ("LL exchange", for the WAR argument).
// ---------------------------------------------------------------------------------------------
typedef unsigned long long ll_word;

__device__ __forceinline__ void ll_store(ll_word* p, float v, uint32_t tag) {
    ll_word w = (static_cast<ll_word>(tag) << 32) | static_cast<ll_word>(__float_as_uint(v));
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ ll_word ll_load(const ll_word* p) {
    ll_word w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void ll_load2(const ll_word* p, ll_word& a, ll_word& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ll_tag(ll_word w) { return static_cast<uint32_t>(w >> 32); }
__device__ __forceinline__ float ll_val(ll_word w) { return __uint_as_float(static_cast<uint32_t>(w)); }

// Spin budget for one exchange before a persistent kernel gives up (cycles of clock64): ~1 s.
constexpr long long LL_TIMEOUT_CYCLES = 2000000000LL;

}  // namespace vqcpc
