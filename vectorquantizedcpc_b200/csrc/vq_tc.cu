// VQ nearest-code search on the tensor cores (VQEmbeddingEMA.encode, /root/reference/model.py:103-115).
//
//   score[n, m] = |e_m|^2 - 2 x_n . e_m      idx[n] = first argmin_m score[n, m]      q[n] = e[idx[n]]
//
// HBM-bound by construction: x is read once (256 B/frame), q and idx written once (264 B/frame); the 512 x 64
// distance contraction never leaves the SM:
//   * the codebook lives in shared memory for the whole kernel as bf16 hi/lo planes of (-2 e) (128 KB, loaded once
//     by TMA, SWIZZLE_128B) -- the factor -2 is exact in bf16, so the MMA yields -2 x.e directly;
//   * converter warps turn each 128-frame tile of fp32 x into bf16 hi/lo planes written straight into the UMMA
//     K-major SWIZZLE_128B layout (generic-proxy stores + fence.proxy.async), double buffered;
//   * one thread issues tcgen05.mma (M=128, N=256, K=16): three K-segments (x_hi e_hi + x_hi e_lo + x_lo e_hi) per
//     256-code half; the two halves of a tile are the two TMEM accumulators, so the MMAs of one half overlap the
//     argmin epilogue of the other;
//   * epilogue warps (thread = frame) read the accumulators with tcgen05.ld (|e|^2 already folded in by one extra MMA
//     step) and, per 32-column chunk, take the minimum (FMNMX3, half an ALU-pipe instruction per score) and then COUNT
//     the scores within the margin of it on the FMA pipe: c = sat(BIG (min + margin - s)) is exactly 1 or 0 and
//     acc += c (1024 + column) carries count and index in one fp32 (two full-rate FFMA per score; the earlier top-2
//     tracking with the index packed into the mantissa cost 3.5 half-rate instructions per score and bounded the kernel).
//     The coarse scores carry an error eps <= 2^-15 |x| max|e| + 2^-21 max|e|^2; the margin is 2 eps.  Exactly one score
//     within the margin of the minimum: that code is provably the exact fp32 winner.  Anything else (near-ties, exact
//     ties, duplicated codebook rows, NaN) puts the frame on a list and vq_rescan_kernel gives it exact fp32 scores of
//     all 512 codes (the fp32 kernel's arithmetic), first minimum wins -- so the result equals the fp32 path's.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

// PTX wrappers shared with gemm_tc.cu (kept local: both files are self-contained translation units)
namespace vqtc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    // suspend-time hint (ns): the warp may sleep in hardware until the phase completes instead of re-issuing the probe -- the
    // kernel is issue-slot bound and a sixth of its executed instructions were spin probes
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { atomicExch(err, VQCPC_ERR_TIMEOUT); return false; }
    }
    return true;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// K-major SWIZZLE_32B tile (rows of 32 bytes, 8-row groups 256 bytes apart): one K = 16 step of bf16
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(256 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(6) << 61;
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));     // FMNMX3
    return d;
}
__device__ __forceinline__ float ffma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));       // FFMA.SAT: NaN -> +0
    return d;
}
}  // namespace vqtc
using namespace vqtc;

constexpr int VT_D = 64, VT_M = 512, VT_TF = 128;
// 12 warps = 384 threads: the register file then allows 168 registers per thread (13 warps round to 512 threads = 128), which the
// epilogue needs to keep two 32-column TMEM reads in flight
constexpr int VT_MMA_WARP = 0, VT_CONV_WARP0 = 1, VT_CONV_WARPS = 3, VT_EPI_WARP0 = 4, VT_EPI_WARPS = 8, VT_THREADS = (4 + VT_EPI_WARPS) * 32;
constexpr int VT_CONV_ROWS = VT_CONV_WARPS * 4;              // rows covered per converter pass (8 threads per row)
constexpr int VT_CONV_ITEMS = (VT_TF + VT_CONV_ROWS - 1) / VT_CONV_ROWS;
constexpr uint32_t VT_CB_HALF = 256 * 128;                 // bytes of one (plane, half) block of the codebook: 256 rows x 128 B
constexpr uint32_t VT_CB_BYTES = 4 * VT_CB_HALF;           // hi/lo x two halves = 128 KB
constexpr uint32_t VT_PLANE = VT_TF * 128;                 // 16 KB: one bf16 plane of an x tile
constexpr uint32_t VT_XBUF = 2 * VT_PLANE;                 // hi + lo
constexpr uint32_t VT_AUG_A = VT_TF * 32;                  // 4 KB: constant ones block, 128 rows x 32 B (one K = 16 step)
constexpr uint32_t VT_AUG_B = VT_M * 32;                   // 16 KB: |e|^2 block, 512 rows x 32 B
constexpr size_t VT_SMEM = VT_CB_BYTES + 2 * VT_XBUF + VT_AUG_A + VT_AUG_B + 1024;
constexpr uint32_t VT_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) |
                              (static_cast<uint32_t>(128 >> 4) << 24);

struct VqTcParams {
    const float* x;          // (n, 64)
    const float* codebook;   // (512, 64) fp32
    const float* e2;         // (512,) |e_m|^2, fp32 kernel's arithmetic (vq_prepare_kernel)
    const uint4* aug_b;      // (512, 2) the |e|^2 MMA block, 32 bytes per code
    float* out_q;            // (n, 64)
    int64_t* out_idx;        // (n,)
    unsigned* flags;         // [0] = number of flagged frames, [4 ..] their frame numbers (first VQ_TC_FLAG_CAP of them)
    int* err;
    long long n;
    int debug;               // bit 0: skip the argmin math, bit 1: skip the gather/output, bit 2: converters skip the x loads
    long long* trace;        // optional (VQCPC_VQ_TRACE): clock64 stamps of CTA 0, [tile iteration][16]
    int trace_iters;
};

__global__ void __launch_bounds__(VT_THREADS, 1) vq_tc_kernel(const __grid_constant__ CUtensorMap map_cb, VqTcParams p) {
    extern __shared__ __align__(1024) unsigned char vt_smem[];
    __shared__ __align__(8) uint64_t cb_bar, xfull_bar[2], xempty_bar[2], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float e2s[VT_M];
    __shared__ float xnorm[2][VT_TF];
    __shared__ float emax_s;
    __shared__ float2 part[2][2][VT_TF];              // [tile parity][column half][row]: (minimum, count/index sum) of that warp's 256 codes

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(vt_smem) + 1023) & ~uintptr_t(1023));
    unsigned char* cb_s = smem;                       // [plane][half][256 rows][128 B]
    unsigned char* x_s = smem + VT_CB_BYTES;          // [buf][plane][128 rows][128 B]
    unsigned char* aug_a = x_s + 2 * VT_XBUF;         // [128 rows][16 bf16]   (SWIZZLE_32B tile)
    unsigned char* aug_b = aug_a + VT_AUG_A;          // [512 rows][16 bf16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[14] = clock64();      // kernel entry
    if (p.trace != nullptr && threadIdx.x == 0) {            // every CTA: entry stamps (clock, global timer), SM id
        long long* c = p.trace + 16 * p.trace_iters + 8 * blockIdx.x;
        unsigned smid; unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        c[0] = clock64(); c[2] = static_cast<long long>(gt); c[4] = smid;
    }

    if (threadIdx.x == 0) {
        mbar_init(&cb_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&xfull_bar[i], VT_CONV_WARPS);  // one arrival per converter warp
            mbar_init(&xempty_bar[i], 1);             // tcgen05.commit
            mbar_init(&tfull_bar[i], 1);              // tcgen05.commit
            mbar_init(&tempty_bar[i], VT_EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == VT_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // |e_m|^2 and its MMA block were prepared once by vq_prepare_kernel: |e_m|^2 rides into the accumulator through one
    // extra K = 16 MMA step (A row = ones, B row m = 0.5 |e_m|^2 split into three bf16 pieces, exact to fp32 precision;
    // both 16-byte halves of a row carry the same three entries, so the product is |e_m|^2 whatever the swizzle does
    // with the two halves of either operand's rows).
    for (int m = threadIdx.x; m < VT_M; m += VT_THREADS) e2s[m] = __ldg(p.e2 + m);
    for (int i = threadIdx.x; i < VT_M * 2; i += VT_THREADS) reinterpret_cast<uint4*>(aug_b)[i] = __ldg(p.aug_b + i);
    for (int r = threadIdx.x; r < VT_TF; r += VT_THREADS) {
        __nv_bfloat16* arow = reinterpret_cast<__nv_bfloat16*>(aug_a + r * 32);
#pragma unroll
        for (int k = 0; k < 16; ++k) arow[k] = __float2bfloat16_rn((k & 7) < 3 ? 1.0f : 0.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) {
        float mx = 0.f;
        for (int m = lane; m < VT_M; m += 32) mx = fmaxf(mx, e2s[m]);
        mx = warp_max(mx);
        if (lane == 0) emax_s = sqrtf(mx);
    }
    if (threadIdx.x == 0) {
        mbar_expect_tx(&cb_bar, VT_CB_BYTES);
        for (int plane = 0; plane < 2; ++plane)
            for (int half = 0; half < 2; ++half)
                tma_load_2d(cb_s + (plane * 2 + half) * VT_CB_HALF, &map_cb, plane * VT_D, half * 256, &cb_bar);
    }
    __syncthreads();
    const uint32_t tmem_base = tmem_base_slot;
    const long long n_tiles = (p.n + VT_TF - 1) / VT_TF;
    // phase stamps of CTA 0: slots 0-3 MMA issuer, 4-6 converter warp 1, 7-13 epilogue warp (quarter 0, ch 0)
#define VT_TRACE(slot) if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && it < p.trace_iters) p.trace[it * 16 + (slot)] = clock64();

    if (warp == VT_MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            bool ok = mbar_wait(&cb_bar, 0, p.err);
            uint32_t xphase[2] = {0, 0}, tphase[2] = {0, 0};
            int it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                VT_TRACE(0)
                ok = mbar_wait(&xfull_bar[buf], xphase[buf], p.err);       // planes of this tile are in shared memory
                if (!ok) break;
                VT_TRACE(1)
                xphase[buf] ^= 1;
                tc_fence_after();
                const uint32_t xa = smem_u32(x_s + buf * VT_XBUF);
                const uint64_t a_hi = umma_desc_sw128(xa), a_lo = umma_desc_sw128(xa + VT_PLANE);
                for (int half = 0; half < 2 && ok; ++half) {
                    ok = mbar_wait(&tempty_bar[half], tphase[half] ^ 1, p.err);   // epilogue drained this accumulator
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t cb = smem_u32(cb_s);
                    const uint64_t b_hi = umma_desc_sw128(cb + (0 * 2 + half) * VT_CB_HALF);
                    const uint64_t b_lo = umma_desc_sw128(cb + (1 * 2 + half) * VT_CB_HALF);
                    const uint32_t d = tmem_base + half * 256;
                    if (!(p.debug & 8)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(d, a_hi + 2 * k, b_hi + 2 * k, VT_IDESC, k > 0 ? 1u : 0u);
                    if (!(p.debug & 32)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(d, a_hi + 2 * k, b_lo + 2 * k, VT_IDESC, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(d, a_lo + 2 * k, b_hi + 2 * k, VT_IDESC, 1u);
                    }
                    tc_mma_f16(d, umma_desc_sw32(smem_u32(aug_a)), umma_desc_sw32(smem_u32(aug_b) + half * 256 * 32), VT_IDESC, 1u);   // + |e|^2
                    }
                    tc_commit(&tfull_bar[half]);
                    tphase[half] ^= 1;
                    if (half == 0) { VT_TRACE(2) } else { VT_TRACE(3) }
                }
                tc_commit(&xempty_bar[buf]);                               // planes buffer reusable
            }
        }
    } else if (warp < VT_EPI_WARP0) {
        // ------------------------------------------------------------------ converters: fp32 x -> bf16 hi/lo planes
        // item (row r, 16-byte chunk c): thread t of the 96 handles c = t & 7, rows r = (t >> 3) + 12 j, j < 11 (r < 128).
        const int t = threadIdx.x - 32 * VT_CONV_WARP0;
        const int c = t & 7, r0 = t >> 3;
        uint32_t ephase[2] = {0, 0};
        bool ok = true;
        int it = 0;
        float4 cur[VT_CONV_ITEMS][2];
        auto load_tile = [&](long long tile, float4 (&dst)[VT_CONV_ITEMS][2]) {
#pragma unroll
            for (int j = 0; j < VT_CONV_ITEMS; ++j) {
                const int r = r0 + VT_CONV_ROWS * j;
                const long long fr = tile * VT_TF + r;
                if (r < VT_TF && fr < p.n && !(p.debug & 4)) {
                    const float4* src = reinterpret_cast<const float4*>(p.x + fr * VT_D + 8 * c);
                    dst[j][0] = __ldg(src); dst[j][1] = __ldg(src + 1);
                } else {
                    dst[j][0] = make_float4(0.f, 0.f, 0.f, 0.f); dst[j][1] = dst[j][0];
                }
            }
        };
        if (blockIdx.x < n_tiles) load_tile(blockIdx.x, cur);
        for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (warp == VT_CONV_WARP0) { VT_TRACE(4) }
            ok = mbar_wait(&xempty_bar[buf], ephase[buf] ^ 1, p.err);      // MMAs of the tile two iterations ago are done
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            ephase[buf] ^= 1;
            if (warp == VT_CONV_WARP0) { VT_TRACE(5) }
            unsigned char* hi = x_s + buf * VT_XBUF;
            unsigned char* lo = hi + VT_PLANE;
            float ssq[VT_CONV_ITEMS];
#pragma unroll
            for (int j = 0; j < VT_CONV_ITEMS; ++j) {
                ssq[j] = 0.f;
                if (p.debug & 16) break;
                const int r = r0 + VT_CONV_ROWS * j;
                if (r >= VT_TF) break;                     // last pass: only the first rows exist (uniform per 8-lane row group)
                const float v[8] = {cur[j][0].x, cur[j][0].y, cur[j][0].z, cur[j][0].w, cur[j][1].x, cur[j][1].y, cur[j][1].z, cur[j][1].w};
                __nv_bfloat162 h2[4], l2[4];
                float ss = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float a = v[2 * q], b = v[2 * q + 1];
                    h2[q] = __floats2bfloat162_rn(a, b);                                  // one cvt.rn.bf16x2.f32
                    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2[q]);
                    const float ha = __uint_as_float(hb << 16), hbv = __uint_as_float(hb & 0xffff0000u);
                    l2[q] = __floats2bfloat162_rn(a - ha, b - hbv);
                    ss = fmaf(a, a, ss);
                    ss = fmaf(b, b, ss);
                }
                ssq[j] = ss;
                // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk (c ^ (r & 7)) of that row's 128 bytes
                const uint32_t off = r * 128 + ((c ^ (r & 7)) << 4);
                // st.shared, not a generic store: x_s comes out of an integer alignment round trip, which hides its address space
                {
                    const uint4 hv = *reinterpret_cast<const uint4*>(h2), lv = *reinterpret_cast<const uint4*>(l2);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(hi) + off), "r"(hv.x), "r"(hv.y), "r"(hv.z), "r"(hv.w) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(lo) + off), "r"(lv.x), "r"(lv.y), "r"(lv.z), "r"(lv.w) : "memory");
                }
            }
            // row norms: the 8 lanes of a row reduce their partial sums (all rows' butterflies interleaved)
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
                for (int j = 0; j < VT_CONV_ITEMS; ++j) ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
            }
            if (c == 0) {
#pragma unroll
                for (int j = 0; j < VT_CONV_ITEMS; ++j)
                    if (r0 + VT_CONV_ROWS * j < VT_TF) xnorm[buf][r0 + VT_CONV_ROWS * j] = sqrtf(ssq[j]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(&xfull_bar[buf]);
            if (warp == VT_CONV_WARP0) { VT_TRACE(6) }
            if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x, cur);   // prefetch the next tile's rows
        }
    } else {
        // ------------------------------------------------------------------ epilogue: min + count-within-margin, gather
        // 8 warps: warp e handles TMEM lanes 32*(warp & 3) (its frame rows) and columns [128*ch, 128*ch + 128) of each
        // 256-code accumulator, ch = e >> 2.  State per subset of scores: (m, acc) = (minimum, sum over the scores within the
        // margin of m of 1024 + code) -- acc in [1024, 1536) means "exactly one candidate, code acc - 1024".  Two subsets
        // merge exactly: the one whose minimum is lower by more than the margin wins outright (the other holds no candidate);
        // minima within the margin of each other mean at least two candidates (acc := 4096, invalid).
        const int e = warp - VT_EPI_WARP0, quarter = warp & 3, ch = e >> 2;
        const int row = quarter * 32 + lane;
        uint32_t tphase[2] = {0, 0};
        bool ok = true;
        int it = 0;
        const float emax = emax_s;
        constexpr float BIG = 1.2676506e30f;               // 2^100: BIG * (thr - s) saturates to exactly 1 for any s < thr (|thr| >= 2^-76)
        int pend_code = 0;                                  // decided code of this lane's row of the PREVIOUS tile
        int pend_tile = -1;                                 // tile whose output rows are still to be written (-1: none); 32-bit on purpose
#define PEND_FR0() (static_cast<long long>(pend_tile) * VT_TF + quarter * 32 + ch * 16)
#define VT_MERGE(m_, acc_, mo_, acco_)                                                              \
        {                                                                                           \
            const float d_ = (mo_) - (m_);                                                          \
            (acc_) = fabsf(d_) <= margin ? 4096.f : (d_ < -margin ? (acco_) : (acc_));              \
            (m_) = fminf((m_), (mo_));                                                              \
        }
        for (int tile = blockIdx.x; tile < static_cast<int>(n_tiles) && ok; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            float mt = INFINITY, acct = 0.f, margin = 0.f, mgB = 0.f;
            if (e == 0) { VT_TRACE(7) }
            for (int half = 0; half < 2 && ok; ++half) {
                ok = mbar_wait(&tfull_bar[half], tphase[half], p.err);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tphase[half] ^= 1;
                tc_fence_after();
                if (e == 0) { if (half == 0) { VT_TRACE(8) } else { VT_TRACE(10) } }
                if (half == 0) {
                    // read now: the converters may refill this slot two tiles later.  margin = 2 eps:
                    // 2 x (bf16x3 product error 2^-15 |x| max|e|  +  accumulation of |e|^2 into the fp32 accumulator 2^-21 max|e|^2)
                    const float xn = xnorm[buf][row];
                    margin = 6.2e-5f * xn * emax + 1.0e-6f * emax * emax;
                    mgB = margin * BIG;
                }
                // software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is reduced
                const int col0 = half * 256 + ch * 128;
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + col0;
                float mh = INFINITY, acch = 0.f;
                uint32_t va[32], vb[32];
                tc_ld32(tbase, va);
                // deferred output of the previous tile, half of it per accumulator half: 4 gather loads (two rows each) are issued
                // here and stored after this half's pass, so their latency hides behind the min/count work instead of ending
                // every tile (all 8 at once held 32 registers through the pass and cost the TMEM double buffering)
                float4 gq[4];
                const bool flush_now = (pend_tile >= 0) && !(p.debug & 2);
                if (flush_now) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int r = 2 * (4 * half + k) + (lane >> 4);      // row within this warp's 16
                        const int code = __shfl_sync(0xffffffffu, pend_code, ch * 16 + r);
                        // volatile asm: the load must be ISSUED here (the compiler would otherwise sink it to its first use,
                        // the store after the pass, and expose the whole latency again)
                        const float4* src = reinterpret_cast<const float4*>(p.codebook + (code < 0 ? 0 : code) * VT_D) + (lane & 15);
                        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(gq[k].x), "=f"(gq[k].y), "=f"(gq[k].z), "=f"(gq[k].w) : "l"(src));
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (p.debug & 1) { tc_wait_ld(); break; }
                    uint32_t (&v)[32] = (cc & 1) ? vb : va;
                    uint32_t (&vn)[32] = (cc & 1) ? va : vb;
                    tc_wait_ld();
                    if (cc + 1 < 4) tc_ld32(tbase + 32 * (cc + 1), vn);
                    // pass 1 (ALU pipe): minimum of the 32 scores, two chains of FMNMX3
                    float ma = fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
                    float mb = fmin3(__uint_as_float(v[16]), __uint_as_float(v[17]), __uint_as_float(v[18]));
#pragma unroll
                    for (int j = 3; j < 15; j += 2) {
                        ma = fmin3(ma, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        mb = fmin3(mb, __uint_as_float(v[16 + j]), __uint_as_float(v[16 + j + 1]));
                    }
                    const float mc = fmin3(ma, mb, fminf(__uint_as_float(v[15]), __uint_as_float(v[31])));
                    // pass 2 (FMA pipe): c = sat(BIG (mc + margin - s)) in {0, 1};  acc += c (1024 + local column)
                    const float tb = fmaf(mc, BIG, mgB);
                    float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        // in place: the score's register becomes its indicator (keeps both TMEM buffers within the register budget)
                        asm("fma.rn.sat.f32 %0, %0, %1, %2;" : "+r"(v[j]) : "f"(-BIG), "f"(tb));
                        a4[j & 3] = fmaf(__uint_as_float(v[j]), static_cast<float>(1024 + 32 * cc + j), a4[j & 3]);
                    }
                    const float accc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
                    VT_MERGE(mh, acch, mc, accc)
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[half]);
                if (flush_now) {
                    // one fully coalesced 256-byte row per 16 lanes
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int r = 2 * (4 * half + k) + (lane >> 4);
                        const int code = __shfl_sync(0xffffffffu, pend_code, ch * 16 + r);      // < 0: flagged row, left to the rescan
                        if (code >= 0 && PEND_FR0() + r < p.n) reinterpret_cast<float4*>(p.out_q + (PEND_FR0() + r) * VT_D)[lane & 15] = gq[k];
                    }
                    if (half == 1) pend_tile = -1;
                }
                // fold this half's 128 codes into the tile state; the column index becomes the code index (an invalid acc stays invalid:
                // 0 + col0 < 1024, 4096 + col0 >= 2048)
                const float accg = acch + static_cast<float>(col0);
                VT_MERGE(mt, acct, mh, accg)
                if (e == 0) { if (half == 0) { VT_TRACE(9) } else { VT_TRACE(11) } }
            }
            if (!ok) break;
            // merge the two column halves of the row: both warps publish their state and read the other's (the merge is
            // symmetric, so both arrive at the same result) and each takes half of the output work.
            part[buf][ch][row] = make_float2(mt, acct);
            bar_sync(1 + quarter, 64);            // only the two warps of this row quarter meet
            if (e == 0) { VT_TRACE(12) }
            {
                const float2 o = part[buf][ch ^ 1][row];
                VT_MERGE(mt, acct, o.x, o.y)
            }
            const long long fr = static_cast<long long>(tile) * VT_TF + row;
            // warp (quarter, ch) finishes rows quarter*32 + ch*16 + [0, 16): lanes ch*16 .. ch*16+15 own them
            const bool owner = (lane >> 4) == ch && fr < p.n && !(p.debug & 2);
            // Exactness.  Every coarse score is within eps of its exact fp32 score and margin = 2 eps, so a code can only beat
            // (or tie) the coarse winner exactly if its coarse score is <= min + margin.  acc in [1024, 1536): the minimum is the
            // ONLY such score, hence the strict exact winner.  Otherwise the frame is listed for vq_rescan_kernel (next launch in
            // the stream): exact fp32 argmin over all 512 codes.  |BIG thr| < 1e10 (scores of magnitude < 1e-20, where BIG (thr - s)
            // might not saturate) goes the same way.  ~0.1-0.3 % of the frames of an init-like codebook, none of a trained one.
            const bool valid = acct >= 1024.f && acct < 1536.f && fabsf(fmaf(mt, BIG, mgB)) >= 1e10f;
            const int i1 = static_cast<int>(acct) - 1024;
            const bool flagged = owner && !valid;
            if (owner) p.out_idx[fr] = flagged ? -1LL : static_cast<long long>(i1);
            if (flagged) {
                const unsigned pos = atomicAdd(p.flags, 1u);
                if (pos < static_cast<unsigned>(VQ_TC_FLAG_CAP)) p.flags[4 + pos] = static_cast<unsigned>(fr);
            }
            __syncwarp();
            pend_code = flagged ? -1 : i1;                      // a flagged row is written by the rescan kernel, not by the deferred flush
            pend_tile = tile;
            if (e == 0) { VT_TRACE(13) }
        }
        if (pend_tile >= 0 && !(p.debug & 2)) {               // output of the last tile
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = 2 * k + (lane >> 4);
                const int code = __shfl_sync(0xffffffffu, pend_code, ch * 16 + r);
                if (code >= 0 && PEND_FR0() + r < p.n) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p.codebook + code * VT_D) + (lane & 15));
                    reinterpret_cast<float4*>(p.out_q + (PEND_FR0() + r) * VT_D)[lane & 15] = v;
                }
            }
        }
    }
#undef VT_MERGE
#undef VT_TRACE
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[15] = clock64();      // all roles done
    if (p.trace != nullptr && threadIdx.x == 0) {
        long long* c = p.trace + 16 * p.trace_iters + 8 * blockIdx.x;
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        c[1] = clock64(); c[3] = static_cast<long long>(gt);
    }
    if (warp == VT_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// Exact pass over the frames vq_tc_kernel flagged: one warp per flagged frame, exact fp32 scores of ALL 512 codes with the fp32
// kernel's arithmetic (sequential FMA over k, score = |e|^2 - 2 x.e), first minimum wins (torch.argmin semantics); writes the
// index and the quantised row.  The frames come from the list the main kernel appended to (under one per thousand); if the list
// overflowed (or frame numbers do not fit 32 bits) every frame's index is inspected instead (out_idx < 0 = flagged).  A CTA
// that has work stages the fp32 codebook in shared memory once (row pitch 65 words: lane = code reads are conflict-free), so a
// frame costs ~2 000 warp instructions instead of 16 dependent L2 round trips.
constexpr int VT_RS_PITCH = VT_D + 1;
constexpr size_t VT_RS_SMEM = sizeof(float) * VT_M * VT_RS_PITCH;
constexpr int VT_RS_WARPS = 8;

__device__ __forceinline__ void vq_rescan_frame(const float* __restrict__ x, const float* __restrict__ codebook, const float* Es,
                                                const float* __restrict__ e2, float* __restrict__ out_q,
                                                long long* __restrict__ out_idx, long long fr, int lane) {
    float xv[VT_D];
    {
        const float4* xr = reinterpret_cast<const float4*>(x + fr * VT_D);
#pragma unroll
        for (int k4 = 0; k4 < VT_D / 4; ++k4) {
            const float4 v = __ldg(xr + k4);
            xv[4 * k4] = v.x; xv[4 * k4 + 1] = v.y; xv[4 * k4 + 2] = v.z; xv[4 * k4 + 3] = v.w;
        }
    }
    float best = INFINITY;
    int besti = 0;                                             // all-NaN scores: index 0, as the fp32 kernel
#pragma unroll 1
    for (int i = 0; i < VT_M / 32; i += 4) {
        // four codes per lane at a time (independent chains); codes ascend per lane: strict < keeps the first minimum
        const float* er = Es + (32 * i + lane) * VT_RS_PITCH;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < VT_D; ++k) {
#pragma unroll
            for (int c = 0; c < 4; ++c) d[c] = fmaf(xv[k], er[c * 32 * VT_RS_PITCH + k], d[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int code = 32 * (i + c) + lane;
            const float sc = fmaf(-2.0f, d[c], __ldg(e2 + code));
            if (sc < best) { best = sc; besti = code; }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
        if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) out_idx[fr] = besti;
    if (lane < 16) reinterpret_cast<float4*>(out_q + fr * VT_D)[lane] = __ldg(reinterpret_cast<const float4*>(codebook + besti * VT_D) + lane);
}

__global__ void __launch_bounds__(VT_RS_WARPS * 32) vq_rescan_kernel(const float* __restrict__ x, const float* __restrict__ codebook,
                                                                     const float* __restrict__ e2, float* __restrict__ out_q,
                                                                     long long* __restrict__ out_idx, long long n,
                                                                     const unsigned* __restrict__ flags) {
    extern __shared__ float rs_Es[];                           // [512][65]
    const int lane = threadIdx.x & 31;
    const long long warp = static_cast<long long>(blockIdx.x) * VT_RS_WARPS + (threadIdx.x >> 5);
    const long long n_warps = static_cast<long long>(gridDim.x) * VT_RS_WARPS;
    const unsigned count = flags[0];
    if (count == 0) return;
    const bool listed = count <= static_cast<unsigned>(VQ_TC_FLAG_CAP) && n <= 0xffffffffLL;
    if (listed && static_cast<long long>(blockIdx.x) * VT_RS_WARPS >= count) return;      // no entry for any warp of this CTA
    for (int f = threadIdx.x; f < VT_M * (VT_D / 4); f += VT_RS_WARPS * 32) {
        const int m = f >> 4, q = f & 15;
        const float4 v = __ldg(reinterpret_cast<const float4*>(codebook) + f);
        float* dst = rs_Es + m * VT_RS_PITCH + 4 * q;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    if (listed) {
        for (long long i = warp; i < count; i += n_warps) vq_rescan_frame(x, codebook, rs_Es, e2, out_q, out_idx, flags[4 + i], lane);
        return;
    }
    for (long long base = warp * 32; base < n; base += n_warps * 32) {
        const long long mine = base + lane < n ? out_idx[base + lane] : 0;
        unsigned fm = __ballot_sync(0xffffffffu, mine < 0);
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            vq_rescan_frame(x, codebook, rs_Es, e2, out_q, out_idx, base + src, lane);
        }
    }
}

// Preparation of the codebook for vq_tc_kernel (one warp per code, 14 -> ~3 us per call): bf16 hi/lo planes of -2 e_m (lanes: two
// dimensions each, coalesced), |e_m|^2 with the fp32 kernel's arithmetic (ONE lane, sequential FMA over k -- the order is part
// of the exactness contract), and the 32-byte row of the |e|^2 MMA block.
__global__ void __launch_bounds__(256) vq_prepare_kernel(const float* __restrict__ codebook, __nv_bfloat16* __restrict__ planes,
                                                         float* __restrict__ e2, __nv_bfloat16* __restrict__ aug_b,
                                                         unsigned* __restrict__ flags) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) flags[0] = 0;       // the main kernel's list of flagged frames starts empty
    if (m >= VT_M) return;
    const float* e = codebook + m * VT_D;
    const float2 ev = __ldg(reinterpret_cast<const float2*>(e) + lane);
    {
        const float v0 = -2.0f * ev.x, v1 = -2.0f * ev.y;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        reinterpret_cast<__nv_bfloat162*>(planes + m * 2 * VT_D)[lane] = __halves2bfloat162(h0, h1);
        reinterpret_cast<__nv_bfloat162*>(planes + m * 2 * VT_D + VT_D)[lane] =
            __halves2bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
    }
    // sequential sum over k on lane 0, the values handed over by shuffles
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VT_D / 2; ++k) {
        const float a = __shfl_sync(0xffffffffu, ev.x, k), b = __shfl_sync(0xffffffffu, ev.y, k);
        s = fmaf(a, a, s);
        s = fmaf(b, b, s);
    }
    if (lane == 0) {
        e2[m] = s;
        const float hs = 0.5f * s;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(hs);
        const float r1 = hs - __bfloat162float(h0);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
        __nv_bfloat16* brow = aug_b + m * 16;
        for (int k = 0; k < 16; ++k) brow[k] = __float2bfloat16_rn(0.f);
        brow[0] = h0; brow[1] = h1; brow[2] = h2; brow[8] = h0; brow[9] = h1; brow[10] = h2;
    }
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// planes_ws: 128 KB device scratch for the codebook planes; err: device int
int vq_lookup_tc(const float* x, const float* codebook, int64_t n, float* q, int64_t* idx, void* planes_ws, int* err,
                 cudaStream_t stream) {
    if (n == 0) return VQCPC_OK;
    VQ_ARG(x && codebook && q && idx && planes_ws && err, "vq_lookup_tc: null pointer");
    static PFN_encodeTiled2 fn = nullptr;
    if (fn == nullptr) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess) {
            set_error("vq_lookup_tc: cuTensorMapEncodeTiled is unavailable");
            return VQCPC_ERR_CUDA;
        }
        fn = reinterpret_cast<PFN_encodeTiled2>(fp);
    }
    float* e2_ws = reinterpret_cast<float*>(static_cast<unsigned char*>(planes_ws) + VT_CB_BYTES);
    __nv_bfloat16* aug_ws = reinterpret_cast<__nv_bfloat16*>(static_cast<unsigned char*>(planes_ws) + VT_CB_BYTES + VT_M * 4);
    unsigned* flags_ws = reinterpret_cast<unsigned*>(static_cast<unsigned char*>(planes_ws) + VT_CB_BYTES + VT_M * 4 + VT_M * 32);
    vq_prepare_kernel<<<VT_M * 32 / 256, 256, 0, stream>>>(codebook, static_cast<__nv_bfloat16*>(planes_ws), e2_ws, aug_ws, flags_ws);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    CUtensorMap map;
    cuuint64_t gdim[2] = {2 * VT_D, VT_M};
    cuuint64_t gstr[1] = {2 * VT_D * 2};
    cuuint32_t box[2] = {VT_D, 256};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, planes_ws, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("vq_lookup_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return VQCPC_ERR_CUDA; }
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(vq_tc_kernel), static_cast<int>(VT_SMEM))) return rc_attr;
    const long long n_tiles = (n + VT_TF - 1) / VT_TF;
    VQ_ARG(n_tiles < (1LL << 31) - 1024, "vq_lookup_tc: too many frames for one call (%lld)", static_cast<long long>(n));
    const int sms = device_sm_count();
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VQCPC_VQ_DEBUG"); dbg = e ? atoi(e) : 0; }
    // VQCPC_VQ_TRACE=<iterations>: phase stamps of CTA 0 (MMA issuer, one converter warp, one epilogue warp), printed to
    // stderr as median cycles per tile -- the diagnostic behind the pipeline numbers in DESIGN.md 4.3
    static int trace_iters = -1;
    if (trace_iters < 0) { const char* e = getenv("VQCPC_VQ_TRACE"); trace_iters = e ? atoi(e) : 0; }
    long long* d_trace = nullptr;
    if (trace_iters > 0) {
        VQ_CUDA(cudaMalloc(&d_trace, sizeof(long long) * (16 * trace_iters + 8 * sms)));
        VQ_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(long long) * (16 * trace_iters + 8 * sms), stream));
    }
    VqTcParams p{x, codebook, e2_ws, reinterpret_cast<const uint4*>(aug_ws), q, idx, flags_ws, err, static_cast<long long>(n), dbg, d_trace, trace_iters};
    vq_tc_kernel<<<static_cast<unsigned>(n_tiles < sms ? n_tiles : sms), VT_THREADS, VT_SMEM, stream>>>(map, p);
    VQ_CUDA(cudaGetLastError());
    {
        if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(vq_rescan_kernel), static_cast<int>(VT_RS_SMEM))) return rc_attr;
        vq_rescan_kernel<<<sms, VT_RS_WARPS * 32, VT_RS_SMEM, stream>>>(x, codebook, e2_ws, q, reinterpret_cast<long long*>(idx), n, flags_ws);
        VQ_CUDA(cudaGetLastError());
    }
    count_launch(2);
    if (d_trace != nullptr) {
        std::vector<long long> h(16 * static_cast<size_t>(trace_iters) + 8 * static_cast<size_t>(sms));
        VQ_CUDA(cudaMemcpyAsync(h.data(), d_trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, stream));
        VQ_CUDA(cudaStreamSynchronize(stream));
        VQ_CUDA(cudaFree(d_trace));
        const long long per_cta = (n_tiles + (n_tiles < sms ? n_tiles : sms) - 1) / (n_tiles < sms ? n_tiles : sms);
        const int iters = static_cast<int>(per_cta - 1 < trace_iters ? per_cta - 1 : trace_iters);
        {
            // per-CTA spans: cycles entry -> done, and the global-timer window of the whole grid
            const int ncta = static_cast<int>(n_tiles < sms ? n_tiles : sms);
            std::vector<long long> cyc;
            long long g0 = -1, g1 = -1, gfirst_end = -1;
            for (int b = 0; b < ncta; ++b) {
                const long long* c = h.data() + 16 * trace_iters + 8 * b;
                cyc.push_back(c[1] - c[0]);
                if (g0 < 0 || c[2] < g0) g0 = c[2];
                if (c[3] > g1) g1 = c[3];
                if (gfirst_end < 0 || c[3] < gfirst_end) gfirst_end = c[3];
            }
            std::sort(cyc.begin(), cyc.end());
            fprintf(stderr, "[vq_tc trace] %d CTAs: cycles entry->done min %lld median %lld max %lld; global timer: grid span %lld ns, first CTA done after %lld ns\n",
                    ncta, cyc.front(), cyc[cyc.size() / 2], cyc.back(), g1 - g0, gfirst_end - g0);
        }
        if (iters > 4) {
            // differences between stamps, median over iterations 2 .. iters-1 (steady state)
            auto med = [&](int a_slot, int a_it_off, int b_slot) {
                std::vector<long long> d;
                for (int it = 2; it + a_it_off < iters; ++it) d.push_back(h[(it + a_it_off) * 16 + a_slot] - h[it * 16 + b_slot]);
                std::sort(d.begin(), d.end());
                return d.empty() ? 0LL : d[d.size() / 2];
            };
            fprintf(stderr, "[vq_tc trace] CTA 0: first tile's epilogue starts %lld cycles after its converter's first stamp; tile starts (epilogue) relative to it:",
                    h[0 * 16 + 7] - h[0 * 16 + 4]);
            for (int it = 0; it < iters; it += 8) fprintf(stderr, " t%d=%lld", it, h[it * 16 + 7] - h[0 * 16 + 4]);
            fprintf(stderr, "\n[vq_tc trace] kernel entry -> first converter stamp %lld cycles; entry -> all roles done %lld cycles (%lld tiles on this CTA)\n",
                    h[4] - h[14], h[15] - h[14], per_cta);
            fprintf(stderr,
                    "[vq_tc trace, CTA 0, cycles, median of %d tiles]\n"
                    "  tile period (epilogue warp)      %lld\n"
                    "  MMA issuer : wait x planes %lld | issue+commit half0 %lld | half1 (incl. wait for a free accumulator) %lld\n"
                    "  converter  : wait free plane buffer %lld | convert + publish %lld | rest of its loop (prefetch issue) %lld\n"
                    "  epilogue   : wait acc0 %lld | top-2 half0 %lld | wait acc1 %lld | top-2 half1 %lld | merge+barrier %lld | re-decide + output %lld\n",
                    iters - 2, med(7, 1, 7), med(1, 0, 0), med(2, 0, 1), med(3, 0, 2), med(5, 0, 4), med(6, 0, 5), med(4, 1, 6),
                    med(8, 0, 7), med(9, 0, 8), med(10, 0, 9), med(11, 0, 10), med(12, 0, 11), med(13, 0, 12));
        }
    }
    return VQCPC_OK;
}

}  // namespace vqcpc
