// Vocoder.generate / Vocoder.forward sample loop of /root/reference/network_vocoder.py:41-78 (rnnms AR part restated
// per SURVEY.md App. A.3) for ONE utterance: the latency kernel of BASELINE configs[2].
//
// What B200 measures (tools/hop_microbench.cu, profiles/r02_hop_microbench.txt): any L2 exchange with more than two
// parties costs ~2100 cycles whatever its fan-in (pairs: 750 same die / 1300 across), a DSMEM hop inside a cluster
// ~330-400 one way (16-way all-gather of 16 words: ~620 per exchange), and only 7 clusters of 16 CTAs are co-resident
// at one CTA per SM.  Hence: ONE grid-scope exchange per step (the all-gather of h_t that W_hh needs anyway) and
// everything after it -- fc1, fc2, softmax, sampling -- redundantly inside every 16-CTA cluster over DSMEM:
//
//   grid   : 7 clusters x 16 CTAs = 112 CTAs x 256 threads; CTA c owns hidden units 8c..8c+7
//   step t : C warp   gates_t(x_{t-1})  ->  publish h_t[8c..8c+7] as LL words (L2)                       [grid hop]
//            M warps  poll their 128-word segment of h_t (4 words per lane) and, straight from those registers, form
//                     the partial sums of this CTA's 16 fc1 rows over the lane's 4 columns (packed FFMA2: two rows per
//                     instruction), reduce them over the warp with a transposing butterfly -> fpart[warp][16]
//            all      r = relu(sum of the 7 warps' partials + b) -> fc2 column-partials over the 16 own r values for
//                     all 256 classes -> reduce-scatter into the owning rank's inbox (st.async)          [DSMEM hop]
//            C warp   sums 16 partials -> its 16 logits -> all-gather to the 16 ranks (DSMEM LL words)   [DSMEM hop]
//                     softmax + inverse-CDF sample x_t (every CTA redundantly, bit-identical) -> gates_{t+1}
//            M warps  (off the critical path) W_hh rows of the own units x h_t -> shared memory for gates_{t+1}
// All weights live in registers: per M-warp lane 24 W_hh rows x 4 columns and 16 fc1 rows x 4 columns (as row pairs for
// fma.rn.f32x2), 16 fc2 weights per thread.  h_t never goes through shared memory.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {
namespace cg = cooperative_groups;

constexpr int CL_S = 16;                 // CTAs per cluster
constexpr int CL_K = 7;                  // clusters
constexpr int CL_CTAS = CL_S * CL_K;     // 112
constexpr int CL_H = 896, CL_G = 3 * CL_H, CL_FC = 256, CL_Q = 256;
constexpr int CL_U = CL_H / CL_CTAS;     // 8 hidden units per CTA
constexpr int CL_ROWS = 3 * CL_U;        // 24 W_hh rows per CTA, gate-major: row = g * 8 + u
constexpr int CL_FR = CL_FC / CL_S;      // 16 fc1 rows / logits per CTA
constexpr int CL_MW = 7;                 // matvec warps, warp q <-> h words [128 q, 128 q + 128) = the units of cluster q
constexpr int CL_MT = 32 * CL_MW;        // 224
constexpr int CL_THREADS = CL_MT + 32;   // + the chain warp
constexpr int CL_CHUNKS = CL_H / 64;     // 14 fc1 column chunks of 64
constexpr int CL_X_INIT = 128;
constexpr int CL_TRACE_STRIDE = 32;     // slots 0..7 chain warp, 8..11 M warp 0, 16+q / 24+q: poll-done clock / poll rounds of M warp q
static_assert(CL_U == 8 && CL_MW * 128 == CL_H && CL_CHUNKS * 16 == CL_MT, "layout");

struct ClParams {
    const float* w_hh; const float* b_hh; const float* fc1_w; const float* fc1_b; const float* fc2_w; const float* fc2_b;
    const float* eprime; const float* lut;
    const float* G;            // (T2, 2688)
    const float* uniforms;     // (L,)   generate mode
    const int64_t* x_in;       // (L,)   teacher-forced mode
    float* out_wav; int32_t* out_codes; float* out_logits;
    ll_word* hbuf;             // [2][896] LL words, zeroed before launch
    int* status;
    long long* trace; int trace_cta, trace_t0, trace_n;
    int L, upsample;
    int poll_delay, poll_mode;
};

__device__ __forceinline__ unsigned cl_mapa(unsigned local, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// DSMEM exchanges: st.async carries plain fp32 payload to the peer's shared memory and completes `bytes` of transaction count
// on the peer's mbarrier -- data and signal in one instruction, no tags, no polling of 256 words, half the DSMEM bytes of
// 8-byte LL words (DSMEM moves only ~20 B/clk per SM).  The receiver arms expect_tx once per phase and waits in hardware.
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned mbar, unsigned parity) {
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void st_async_f32(unsigned dst, float v, unsigned mbar_remote) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(dst), "f"(v), "r"(mbar_remote) : "memory");
}
__device__ __forceinline__ void st_async_v4(unsigned dst, float a, float b, float c, float d, unsigned mbar_remote) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst), "f"(a), "f"(b),
                 "f"(c), "f"(d), "r"(mbar_remote) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// sigmoid / tanh from ex2.approx / rcp.approx (each ~2^-22 relative): no denormal rescue, no division sequence on the chain
__device__ __forceinline__ float sigmoid_chain(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_chain(float x) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f); }
__device__ __forceinline__ float redux_max(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
__device__ __forceinline__ unsigned redux_min(unsigned v) {
    unsigned m;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(v));
    return m;
}

// packed fp32 FMA (FFMA2 on sm_100a): two independent round-to-nearest FMAs per instruction -- bit-identical to two fmaf
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

constexpr unsigned CL_XBYTES = CL_Q * 4;         // bytes per exchange phase: 256 fp32 into every CTA

// TRACE = true is the instrumented build (tools/cl_trace.py): every trace point reads the clock only after the value named as
// its dependency is available, so the stamps mark COMPLETION of a phase (a bare clock read issues in order right after the
// previous instruction issues, and bar.sync does not block at issue).  The production instantiation contains none of it.
template <bool TRACE>
__global__ void __launch_bounds__(CL_THREADS, 1) ar_cluster_kernel(ClParams p) {
    __shared__ __align__(16) float Es[CL_Q * CL_ROWS];        // E'[x][g*8+u] of the own units (24 KB)
    __shared__ __align__(16) float fpart[CL_MW * CL_FR];      // fc1: partial sums of the 16 own rows over each M warp's 128 columns
    __shared__ __align__(16) float hhpart[CL_MW * CL_ROWS];
    __shared__ __align__(16) float inboxf[2 * CL_S * CL_FR];  // [par][source rank][own logit]: fc2 column-partials
    __shared__ __align__(16) float lgf[2 * CL_Q];             // [par][class]: all 256 logits
    __shared__ __align__(8) unsigned long long mbars[4];      // reduce-scatter par 0/1, all-gather par 0/1
    __shared__ volatile int abort_flag;

    cg::cluster_group cluster = cg::this_cluster();
    // the warp index goes through a shuffle so that the compiler knows it is warp-uniform: code under `if (warp < ...)` is then
    // convergent and every later __shfl_sync / redux compiles to the bare instruction (no WARPSYNC.COLLECTIVE slow-path check)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int cta = blockIdx.x, rank = static_cast<int>(cluster.block_rank());
    const int L = p.L;

    for (int i = tid; i < CL_Q * CL_ROWS; i += CL_THREADS) {
        const int x = i / CL_ROWS, j = i % CL_ROWS, g = j / CL_U, u = j % CL_U;
        Es[i] = __ldg(p.eprime + static_cast<int64_t>(x) * CL_G + g * CL_H + cta * CL_U + u);
    }
    for (int i = tid; i < CL_MW * CL_ROWS; i += CL_THREADS) hhpart[i] = 0.f;
    const unsigned mbar_a = static_cast<unsigned>(__cvta_generic_to_shared(&mbars[0]));
    if (tid == 0) {
        abort_flag = 0;
        for (int i = 0; i < 4; ++i) mbar_init(mbar_a + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();      // every CTA of the cluster is running and its mbarriers exist

    const unsigned inbox_a = static_cast<unsigned>(__cvta_generic_to_shared(&inboxf[0]));
    const unsigned lg_a = static_cast<unsigned>(__cvta_generic_to_shared(&lgf[0]));
    const bool teacher = (p.x_in != nullptr);
    const bool pipelined_poll = (p.poll_mode & 1) != 0;
    const bool tracing = TRACE && (p.trace != nullptr) && (cta == p.trace_cta) && (lane == 0);
#define CL_TRACE(k, tt, dep)                                                                   \
    if constexpr (TRACE) {                                                                     \
        long long clk_;                                                                        \
        asm volatile("mov.u64 %0, %%clock64; // after %1" : "=l"(clk_) : "r"(dep) : "memory"); \
        if (tracing && (tt) >= p.trace_t0 && (tt) < p.trace_t0 + p.trace_n)                    \
            p.trace[((tt) - p.trace_t0) * CL_TRACE_STRIDE + (k)] = clk_;                       \
    }
#define FU(v) __float_as_uint(v)
    // A timed-out wait sets the CTA's abort flag and the status word; from then on every wait of this CTA falls through at its
    // first probe, the arithmetic runs on with garbage (the host raises VQCPC_ERR_TIMEOUT and discards the outputs) and all
    // barriers keep being served -- no thread-divergent `dead` state, so the compute stays provably convergent.
#define CL_FAIL()                                                          \
    do {                                                                   \
        abort_flag = 1;                                                    \
        if (lane == 0) atomicExch(p.status, VQCPC_ERR_TIMEOUT);            \
    } while (0)

    // ---- per-thread weights shared by all 8 warps
    // fc2 column-partials: thread = (output group og = tid / 4 -> classes 4 og .. 4 og + 3, column group cg = tid % 4 -> own r
    // values 4 cg .. 4 cg + 3); the four column groups meet by shuffles and lane cg == 0 sends ONE 16-byte st.async (a DSMEM
    // store costs its issue slot per thread, not per byte: 64 operations per CTA instead of 256).
    const int og = tid >> 2, cg4 = tid & 3;
    float w2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.fc2_w + static_cast<int64_t>(4 * og + i) * CL_FC + rank * CL_FR + 4 * cg4));
        w2[i][0] = v.x; w2[i][1] = v.y; w2[i][2] = v.z; w2[i][3] = v.w;
    }
    const unsigned rs_dst = cl_mapa(inbox_a + (rank * CL_FR + ((4 * og) & 15)) * 4, static_cast<unsigned>(og >> 2));
    const unsigned rs_mbar = cl_mapa(mbar_a, static_cast<unsigned>(og >> 2));
    const float4 b1v = __ldg(reinterpret_cast<const float4*>(p.fc1_b + rank * CL_FR) + cg4);   // biases of the r values fc2 reads

    // r = relu(fc1 h_t + b) of the own rows 4 cg4 .. +3: the seven M warps' partial sums meet here
    auto fc1_r = [&]() -> float4 {
        float4 pq[CL_MW];
#pragma unroll
        for (int q = 0; q < CL_MW; ++q) pq[q] = *reinterpret_cast<const float4*>(&fpart[q * CL_FR + 4 * cg4]);
        float4 rv;
        rv.x = fmaxf((((pq[0].x + pq[1].x) + (pq[2].x + pq[3].x)) + ((pq[4].x + pq[5].x) + pq[6].x)) + b1v.x, 0.f);
        rv.y = fmaxf((((pq[0].y + pq[1].y) + (pq[2].y + pq[3].y)) + ((pq[4].y + pq[5].y) + pq[6].y)) + b1v.y, 0.f);
        rv.z = fmaxf((((pq[0].z + pq[1].z) + (pq[2].z + pq[3].z)) + ((pq[4].z + pq[5].z) + pq[6].z)) + b1v.z, 0.f);
        rv.w = fmaxf((((pq[0].w + pq[1].w) + (pq[2].w + pq[3].w)) + ((pq[4].w + pq[5].w) + pq[6].w)) + b1v.w, 0.f);
        return rv;
    };
    // fc2 column-partials of classes 4 og .. +3 over those four r values -> the owning rank's inbox; returns s[0] (trace dependency)
    auto fc2_partial_send = [&](const float4 rv, int par) -> float {
        float s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = fmaf(w2[i][3], rv.w, fmaf(w2[i][2], rv.z, fmaf(w2[i][1], rv.y, w2[i][0] * rv.x)));
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], 2);
        if (cg4 == 0) st_async_v4(rs_dst + par * CL_XBYTES, s[0], s[1], s[2], s[3], rs_mbar + 8 * par);
        return s[0];
    };

    if (warp < CL_MW) {
        // =============================================================== M warps
        const int q = warp;
        // W_hh rows (r, r + 12) and fc1 rows (i, i + 8) of this CTA as register pairs: one FFMA2 advances both rows
        float2 whp[CL_ROWS / 2][4], f1p[CL_FR / 2][4];
#pragma unroll
        for (int r = 0; r < CL_ROWS / 2; ++r) {
            const int ra = r, rb = r + CL_ROWS / 2;
            const float4 va = __ldg(reinterpret_cast<const float4*>(
                p.w_hh + static_cast<int64_t>((ra / CL_U) * CL_H + cta * CL_U + ra % CL_U) * CL_H + 128 * q + 4 * lane));
            const float4 vb = __ldg(reinterpret_cast<const float4*>(
                p.w_hh + static_cast<int64_t>((rb / CL_U) * CL_H + cta * CL_U + rb % CL_U) * CL_H + 128 * q + 4 * lane));
            whp[r][0] = make_float2(va.x, vb.x); whp[r][1] = make_float2(va.y, vb.y);
            whp[r][2] = make_float2(va.z, vb.z); whp[r][3] = make_float2(va.w, vb.w);
        }
#pragma unroll
        for (int i = 0; i < CL_FR / 2; ++i) {
            const float4 va = __ldg(reinterpret_cast<const float4*>(p.fc1_w + static_cast<int64_t>(rank * CL_FR + i) * CL_H + 128 * q + 4 * lane));
            const float4 vb = __ldg(reinterpret_cast<const float4*>(p.fc1_w + static_cast<int64_t>(rank * CL_FR + i + CL_FR / 2) * CL_H + 128 * q + 4 * lane));
            f1p[i][0] = make_float2(va.x, vb.x); f1p[i][1] = make_float2(va.y, vb.y);
            f1p[i][2] = make_float2(va.z, vb.z); f1p[i][3] = make_float2(va.w, vb.w);
        }
        const int w0 = 128 * q + 4 * lane;                       // first of this lane's four h words
        const ll_word* hsrc = p.hbuf + w0;
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
        for (int t = 0; t < L; ++t) {
            const uint32_t tag = static_cast<uint32_t>(t) + 1u;
            const int par = t & 1;
            float hv[4] = {0.f, 0.f, 0.f, 0.f};
            // this CTA publishes h_t at (about) the same time as everybody else: do not poll L2 before that.  A hardware
            // barrier, not a shared-memory spin: seven spinning warps would steal issue and LDS slots from the chain warp.
            bar_sync(5, CL_THREADS);
            if (TRACE && warp == 0) { CL_TRACE(8, t, static_cast<unsigned>(abort_flag)) }
            if (p.poll_delay) { const long long t1 = clock64(); while (clock64() - t1 < p.poll_delay) {} }
            {
                const ll_word* src = hsrc + par * CL_H;
                const long long t0 = clock64();
                unsigned n = 0;
                ll_word a0, a1, b0, b1w;
                unsigned rounds = 0;
                if (!pipelined_poll) {
                    for (;;) {
                        ++rounds;
                        ll_load2(src, a0, a1);
                        ll_load2(src + 2, b0, b1w);
                        const bool ok = ll_tag(a0) == tag && ll_tag(a1) == tag && ll_tag(b0) == tag && ll_tag(b1w) == tag;
                        if (__all_sync(0xffffffffu, ok)) break;
                        if (abort_flag || ((++n & 255u) == 0 && clock64() - t0 > LL_TIMEOUT_CYCLES)) { CL_FAIL(); break; }
                    }
                } else {
                    // two poll rounds in flight (measured: slower -- the extra polls delay the stores they wait for)
                    ll_word c0, c1, d0, d1;
                    ll_load2(src, a0, a1);
                    ll_load2(src + 2, b0, b1w);
                    for (;;) {
                        ll_load2(src, c0, c1);
                        ll_load2(src + 2, d0, d1);
                        bool ok = ll_tag(a0) == tag && ll_tag(a1) == tag && ll_tag(b0) == tag && ll_tag(b1w) == tag;
                        if (__all_sync(0xffffffffu, ok)) break;
                        ll_load2(src, a0, a1);
                        ll_load2(src + 2, b0, b1w);
                        ok = ll_tag(c0) == tag && ll_tag(c1) == tag && ll_tag(d0) == tag && ll_tag(d1) == tag;
                        if (__all_sync(0xffffffffu, ok)) { a0 = c0; a1 = c1; b0 = d0; b1w = d1; break; }
                        if (abort_flag || ((++n & 255u) == 0 && clock64() - t0 > LL_TIMEOUT_CYCLES)) { CL_FAIL(); break; }
                    }
                }
                hv[0] = ll_val(a0); hv[1] = ll_val(a1); hv[2] = ll_val(b0); hv[3] = ll_val(b1w);
                CL_TRACE(16 + q, t, FU(hv[0]))
                if (tracing && t >= p.trace_t0 && t < p.trace_t0 + p.trace_n) p.trace[(t - p.trace_t0) * CL_TRACE_STRIDE + 24 + q] = rounds;
            }
            if (TRACE && warp == 0) { CL_TRACE(9, t, FU(hv[3])) }
            const float2 hd[4] = {make_float2(hv[0], hv[0]), make_float2(hv[1], hv[1]), make_float2(hv[2], hv[2]), make_float2(hv[3], hv[3])};
            {
                // ---- critical path: fc1 partials of the 16 own rows over this lane's 4 columns, then the warp-wide sums
                float2 fa[CL_FR / 2];
#pragma unroll
                for (int i = 0; i < CL_FR / 2; ++i)
                    fa[i] = ffma2(f1p[i][3], hd[3], ffma2(f1p[i][2], hd[2], ffma2(f1p[i][1], hd[1], ffma2(f1p[i][0], hd[0], make_float2(0.f, 0.f)))));
                // transposing butterfly: 16 -> 8 -> 4 -> 2 -> 1 values per lane, then one plain round: lane pair l/2 holds row l/2
                float v8[8], v4[4], v2[2];
#pragma unroll
                for (int i = 0; i < 8; ++i) v8[i] = (b4 ? fa[i].y : fa[i].x) + __shfl_xor_sync(0xffffffffu, b4 ? fa[i].x : fa[i].y, 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) v4[i] = (b3 ? v8[i + 4] : v8[i]) + __shfl_xor_sync(0xffffffffu, b3 ? v8[i] : v8[i + 4], 8);
#pragma unroll
                for (int i = 0; i < 2; ++i) v2[i] = (b2 ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, b2 ? v4[i] : v4[i + 2], 4);
                float v1 = (b1 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? v2[0] : v2[1], 2);
                v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                if (!(lane & 1)) fpart[q * CL_FR + (lane >> 1)] = v1;
                if (TRACE && warp == 0) { CL_TRACE(10, t, FU(v1)) }
            }
            bar_sync(3, CL_THREADS);                 // fpart holds all seven warps' fc1 partial sums
            {
                const float4 rv = fc1_r();
                if (TRACE && warp == 0) { CL_TRACE(13, t, FU(rv.x)) }
                const float s0 = fc2_partial_send(rv, par);
                if (TRACE && warp == 0) { CL_TRACE(14, t, FU(s0)) }
                // W_hh below must not be scheduled in front of the send: its inputs pass through an ordered no-op
                asm volatile("" : "+f"(hv[0]), "+f"(hv[1]), "+f"(hv[2]), "+f"(hv[3]));
            }
            const float2 hd2[4] = {make_float2(hv[0], hv[0]), make_float2(hv[1], hv[1]), make_float2(hv[2], hv[2]), make_float2(hv[3], hv[3])};
            {
                // ---- off the critical path: W_hh rows of the own units x h_t (this warp's 128 columns)
                float2 acc[CL_ROWS / 2];
#pragma unroll
                for (int r = 0; r < CL_ROWS / 2; ++r)
                    acc[r] = ffma2(whp[r][3], hd2[3], ffma2(whp[r][2], hd2[2], ffma2(whp[r][1], hd2[1], ffma2(whp[r][0], hd2[0], make_float2(0.f, 0.f)))));
                // transposing butterfly: 24 -> 12 -> 6 -> 3 values per lane, then two plain rounds
                float v12[12], v6[6], v3[3];
#pragma unroll
                for (int i = 0; i < 12; ++i) v12[i] = (b4 ? acc[i].y : acc[i].x) + __shfl_xor_sync(0xffffffffu, b4 ? acc[i].x : acc[i].y, 16);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const float send = b3 ? v12[i] : v12[i + 6];
                    v6[i] = (b3 ? v12[i + 6] : v12[i]) + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float send = b2 ? v6[i] : v6[i + 3];
                    v3[i] = (b2 ? v6[i + 3] : v6[i]) + __shfl_xor_sync(0xffffffffu, send, 4);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    v3[i] += __shfl_xor_sync(0xffffffffu, v3[i], 2);
                    v3[i] += __shfl_xor_sync(0xffffffffu, v3[i], 1);
                }
                const int base = (b4 ? 12 : 0) + (b3 ? 6 : 0) + (b2 ? 3 : 0);
                const int sel = lane & 3;
                if (sel < 3) hhpart[q * CL_ROWS + base + sel] = sel == 0 ? v3[0] : (sel == 1 ? v3[1] : v3[2]);
            }
            bar_arrive(4, CL_THREADS);               // -> C warp: this warp's share of W_hh h_t is in hhpart
            if (TRACE && warp == 0) { CL_TRACE(11, t, 0u) }
        }
    } else {
        // =============================================================== C warp: the sequential chain
        const int gu = cta * CL_U + (lane & 7);
        const float bh_r = __ldg(p.b_hh + gu), bh_z = __ldg(p.b_hh + CL_H + gu), bh_n = __ldg(p.b_hh + 2 * CL_H + gu);
        const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.fc2_b + rank * CL_FR) + (lane & 3));
        // reduce-scatter read: source ranks lane/4 + 8j, own logits 4 (lane%4) .. +3 ; all-gather send: piece lane%4 -> ranks lane/4 + 8j
        const unsigned rs_src = inbox_a + (((lane >> 2) * CL_FR) + 4 * (lane & 3)) * 4;
        unsigned ag_dst[2], ag_mbar[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            ag_dst[j] = cl_mapa(lg_a + (rank * CL_FR + 4 * (lane & 3)) * 4, static_cast<unsigned>((lane >> 2) + 8 * j));
            ag_mbar[j] = cl_mapa(mbar_a + 16, static_cast<unsigned>((lane >> 2) + 8 * j));
        }
        float hown = 0.f, hb_r = bh_r, hb_z = bh_z, hb_n = bh_n;      // W_hh h_{-1} = 0
        float gh_r = 0.f, gh_z = 0.f;                                   // G_r + hb_r, G_z + hb_z: everything but E'[x]
        int x = CL_X_INIT, x_out = -1;
        int x_next = (teacher && L > 0) ? static_cast<int>(__ldg(p.x_in)) & (CL_Q - 1) : 0;
        int frame = 0, frame_left = 0;
        float g_r = 0.f, g_z = 0.f, g_n = 0.f, gn_r = 0.f, gn_z = 0.f, gn_n = 0.f;
        const int n_frames = (L + p.upsample - 1) / p.upsample;
        if (n_frames > 0) {
            const float* g = p.G + gu;
            gn_r = __ldg(g); gn_z = __ldg(g + CL_H); gn_n = __ldg(g + 2 * CL_H);
        }
        for (int t = 0; t < L; ++t) {
            const uint32_t tag = static_cast<uint32_t>(t) + 1u;
            const int par = t & 1;
            const unsigned phase = (static_cast<unsigned>(t) >> 1) & 1u;
            if (frame_left == 0) {
                g_r = gn_r; g_z = gn_z; g_n = gn_n;
                ++frame; frame_left = p.upsample;
                if (frame < n_frames) {           // prefetch the next frame's conditioning (used 160 steps from now)
                    const float* g = p.G + static_cast<int64_t>(frame) * CL_G + gu;
                    gn_r = __ldg(g); gn_z = __ldg(g + CL_H); gn_n = __ldg(g + 2 * CL_H);
                }
            }
            --frame_left;
            if (t == 0) { gh_r = __fadd_rn(g_r, hb_r); gh_z = __fadd_rn(g_z, hb_z); }
            CL_TRACE(0, t, static_cast<unsigned>(x))
            float u_t = 0.f;
            {
                // ---- GRU gates of step t for the own units (lanes 0..7; the other lanes mirror them harmlessly)
                if (teacher) x = x_next;
                const float* e = &Es[x * CL_ROWS + (lane & 7)];
                const float r = sigmoid_chain(__fadd_rn(e[0], gh_r));
                const float z = sigmoid_chain(__fadd_rn(e[CL_U], gh_z));
                const float n = tanh_chain(__fmaf_rn(r, hb_n, __fadd_rn(e[2 * CL_U], g_n)));
                hown = __fmaf_rn(z, __fsub_rn(hown, n), n);            // (1-z) n + z h
                if (lane < CL_U) ll_store(p.hbuf + par * CL_H + cta * CL_U + lane, hown, tag);
            }
            bar_arrive(5, CL_THREADS);               // M warps: h_t is on its way, start polling
            {
                CL_TRACE(1, t, FU(hown))
                if (lane == 0) {                       // arm this step's exchange phases (each receives 256 fp32)
                    mbar_expect_tx(mbar_a + 8 * par, CL_XBYTES);
                    if (!teacher) mbar_expect_tx(mbar_a + 16 + 8 * par, CL_XBYTES);
                }
                if (!teacher) u_t = __ldg(p.uniforms + t);
                else if (t + 1 < L) x_next = static_cast<int>(__ldg(p.x_in + t + 1)) & (CL_Q - 1);
                if (cta == 0 && lane == 0 && x_out >= 0) {             // the previous step's sample leaves here, off the chain
                    if (p.out_wav) p.out_wav[t - 1] = __ldg(p.lut + x_out);
                    if (p.out_codes) p.out_codes[t - 1] = x_out;
                }
            }
            bar_sync(3, CL_THREADS);                 // fpart complete
            float lo[4] = {0.f, 0.f, 0.f, 0.f};
            {
                const float4 rv = fc1_r();
                CL_TRACE(2, t, FU(rv.x))
                const float s0 = fc2_partial_send(rv, par);
                CL_TRACE(3, t, FU(s0))
                // ---- reduce-scatter: 16 ranks x 16 own logits of column-partials
                const long long t0 = clock64();
                unsigned n = 0;
                while (!mbar_try_wait(mbar_a + 8 * par, phase)) {
                    if (abort_flag || ((++n & 63u) == 0 && clock64() - t0 > LL_TIMEOUT_CYCLES)) { CL_FAIL(); break; }
                }
            }
            {
                const float4 pa = *reinterpret_cast<const float4*>(&inboxf[0] + (rs_src - inbox_a) / 4 + par * CL_Q);
                const float4 pb = *reinterpret_cast<const float4*>(&inboxf[0] + (rs_src - inbox_a) / 4 + par * CL_Q + 8 * CL_FR);
                CL_TRACE(4, t, FU(pa.x))
                lo[0] = pa.x + pb.x; lo[1] = pa.y + pb.y; lo[2] = pa.z + pb.z; lo[3] = pa.w + pb.w;
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) lo[i] += __shfl_xor_sync(0xffffffffu, lo[i], o);
                }
                lo[0] += b2.x; lo[1] += b2.y; lo[2] += b2.z; lo[3] += b2.w;   // logits of classes 16 rank + 4 (lane % 4) + {0..3}
                CL_TRACE(5, t, FU(lo[0]))
                if (p.out_logits != nullptr && cta < CL_S && lane < 4)
                    *reinterpret_cast<float4*>(p.out_logits + static_cast<int64_t>(t) * CL_Q + rank * CL_FR + 4 * lane) =
                        make_float4(lo[0], lo[1], lo[2], lo[3]);
                if (!teacher) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) st_async_v4(ag_dst[j] + par * CL_XBYTES, lo[0], lo[1], lo[2], lo[3], ag_mbar[j] + 8 * par);
                }
            }
            bar_sync(4, CL_THREADS);                 // W_hh h_t partial sums of the 7 M warps are in hhpart
            {
                float s = 0.f;
                if (lane < CL_ROWS) {
                    s = hhpart[lane];
#pragma unroll
                    for (int q = 1; q < CL_MW; ++q) s += hhpart[q * CL_ROWS + lane];
                }
                hb_r = __fadd_rn(__shfl_sync(0xffffffffu, s, lane & 7), bh_r);
                hb_z = __fadd_rn(__shfl_sync(0xffffffffu, s, 8 + (lane & 7)), bh_z);
                hb_n = __fadd_rn(__shfl_sync(0xffffffffu, s, 16 + (lane & 7)), bh_n);
                // the next step's conditioning frame is already known here
                const bool nf = (frame_left == 0);
                gh_r = __fadd_rn(nf ? gn_r : g_r, hb_r);
                gh_z = __fadd_rn(nf ? gn_z : g_z, hb_z);
                CL_TRACE(6, t, FU(gh_z))
            }
            if (!teacher) {
                // ---- all-gather of the logits, then softmax + inverse-CDF sample: lane l holds classes 8l..8l+7
                const long long t0 = clock64();
                unsigned n = 0;
                while (!mbar_try_wait(mbar_a + 16 + 8 * par, phase)) {
                    if (abort_flag || ((++n & 63u) == 0 && clock64() - t0 > LL_TIMEOUT_CYCLES)) { CL_FAIL(); break; }
                }
                {
                    const float4 oa = *reinterpret_cast<const float4*>(&lgf[par * CL_Q + 8 * lane]);
                    const float4 ob = *reinterpret_cast<const float4*>(&lgf[par * CL_Q + 8 * lane + 4]);
                    const float o[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
                    CL_TRACE(7, t, FU(oa.x))
                    constexpr float LOG2E = 1.4426950408889634f;
                    const float m = redux_max(fmaxf(fmaxf(fmaxf(o[0], o[1]), fmaxf(o[2], o[3])), fmaxf(fmaxf(o[4], o[5]), fmaxf(o[6], o[7]))));
                    const float ml = -m * LOG2E;
                    float c[8];
                    c[0] = ex2_approx(fmaf(o[0], LOG2E, ml));
#pragma unroll
                    for (int i = 1; i < 8; ++i) c[i] = c[i - 1] + ex2_approx(fmaf(o[i], LOG2E, ml));
                    // radix-4 inclusive scan of the lane sums over the warp (3 shuffle rounds)
                    float v = c[7];
                    {
                        const float s1 = __shfl_up_sync(0xffffffffu, v, 1), s2 = __shfl_up_sync(0xffffffffu, v, 2),
                                    s3 = __shfl_up_sync(0xffffffffu, v, 3);
                        v += ((lane >= 1 ? s1 : 0.f) + (lane >= 2 ? s2 : 0.f)) + (lane >= 3 ? s3 : 0.f);
                        const float s4 = __shfl_up_sync(0xffffffffu, v, 4), s8 = __shfl_up_sync(0xffffffffu, v, 8),
                                    s12 = __shfl_up_sync(0xffffffffu, v, 12);
                        v += ((lane >= 4 ? s4 : 0.f) + (lane >= 8 ? s8 : 0.f)) + (lane >= 12 ? s12 : 0.f);
                        const float s16 = __shfl_up_sync(0xffffffffu, v, 16);
                        v += (lane >= 16 ? s16 : 0.f);
                    }
                    const float excl = v - c[7];
                    const float thr = u_t * __shfl_sync(0xffffffffu, v, 31);
                    unsigned hit = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) hit |= (excl + c[i] > thr) ? (1u << i) : 0u;
                    unsigned cand = hit ? static_cast<unsigned>(8 * lane + __ffs(hit) - 1) : static_cast<unsigned>(CL_Q);
                    cand = redux_min(cand);
                    x = cand < CL_Q ? static_cast<int>(cand) : CL_Q - 1;
                    x_out = x;
                    CL_TRACE(12, t, static_cast<unsigned>(x))
                }
            }
        }
        if (!abort_flag && cta == 0 && lane == 0 && x_out >= 0 && L > 0) {
            if (p.out_wav) p.out_wav[L - 1] = __ldg(p.lut + x_out);
            if (p.out_codes) p.out_codes[L - 1] = x_out;
        }
    }
#undef CL_TRACE
#undef FU
#undef CL_FAIL
    __syncthreads();
    cluster.sync();      // nobody's shared memory disappears while a peer may still store into it
}

// ------------------------------------------------------------------------------------------------ host
static int g_cl_poll_delay = 300, g_cl_poll_mode = 0;
int g_cl_enable = 1;

// 1 = the device can hold the 7 x 16 cluster grid at one CTA per SM, 0 = it cannot (fall back to ar_kernel), cached per device
int ar_cluster_supported() {
    static int cache[64] = {0};     // 0 unknown, 1 yes, 2 no
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cache[dev]) return cache[dev] == 1;
    int ok = 0;
    do {
        if (cudaFuncSetAttribute(ar_cluster_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) break;
        if (cudaFuncSetAttribute(ar_cluster_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) break;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL_CTAS); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = 0;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL_S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, ar_cluster_kernel<false>, &cfg) != cudaSuccess) break;
        ok = ncl >= CL_K;
    } while (0);
    cudaGetLastError();
    cache[dev] = ok ? 1 : 2;
    return ok;
}

size_t ar_cluster_ll_bytes() { return sizeof(ll_word) * 2 * CL_H; }

int ar_cluster_launch(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int L,
                      ll_word* hbuf, int* status, float* out_wav, int32_t* out_codes, float* out_logits, long long* trace,
                      int trace_cta, int trace_t0, int trace_n, cudaStream_t stream) {
    VQ_CUDA(cudaMemsetAsync(hbuf, 0, ar_cluster_ll_bytes(), stream));
    ClParams p{};
    p.w_hh = w->ar_w_hh; p.b_hh = w->ar_b_hh; p.fc1_w = w->fc1_w; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b;
    p.eprime = w->eprime; p.lut = w->mulaw_lut;
    p.G = G; p.uniforms = uniforms; p.x_in = x_in;
    p.out_wav = out_wav; p.out_codes = out_codes; p.out_logits = out_logits;
    p.hbuf = hbuf; p.status = status;
    p.trace = trace; p.trace_cta = trace_cta; p.trace_t0 = trace_t0; p.trace_n = trace_n;
    p.L = L; p.upsample = w->upsample_t;
    p.poll_delay = g_cl_poll_delay; p.poll_mode = g_cl_poll_mode;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL_CTAS); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL_S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;      // all 112 CTAs co-resident, or the launch fails (never a silent hang)
    attr[1].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    if (trace != nullptr) {
        VQ_CUDA(cudaLaunchKernelEx(&cfg, ar_cluster_kernel<true>, p));
    } else {
        VQ_CUDA(cudaLaunchKernelEx(&cfg, ar_cluster_kernel<false>, p));
    }
    count_launch(1);
    return VQCPC_OK;
}

void ar_cluster_set_poll(int delay, int mode) { g_cl_poll_delay = delay; g_cl_poll_mode = mode; }

}  // namespace vqcpc
