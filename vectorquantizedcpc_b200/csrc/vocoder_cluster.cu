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

#include <cstdlib>

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
__device__ __forceinline__ void st_async_v2(unsigned dst, float a, float b, unsigned mbar_remote) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(dst), "f"(a), "f"(b),
                 "r"(mbar_remote) : "memory");
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
    __shared__ __align__(16) float h_s[CL_H];                 // h_t gathered from the grid
    __shared__ __align__(16) float r_s[CL_FR];                // relu(fc1 h_t + b) of the 16 own rows
    __shared__ __align__(16) float hb_s[CL_ROWS];             // W_hh h_t + b_hh of the own 24 rows, gate-major
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
    const unsigned mbar_a = static_cast<unsigned>(__cvta_generic_to_shared(&mbars[0]));
    if (tid == 0) {
        abort_flag = 0;
        for (int i = 0; i < 4; ++i) mbar_init(mbar_a + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();      // every CTA of the cluster is running and its mbarriers exist

    const bool teacher = (p.x_in != nullptr);
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

    // ---- roles.  Everything on an SM is bound by how many instructions its four schedulers issue (tools/sm_microbench.cu:
    // SHFL 4.7 cycles of a scheduler each, FFMA2 2.65, FFMA 1.5, BAR ~40, LDS 30 latency), so the critical fc1 / fc2 work runs on
    // ONE warp per scheduler (warps 4..7, the highest warp ids: the arbiter prefers them), whole rows per warp so that a row
    // needs 5 shuffle rounds once instead of a 7-warp partial-sum tree, and the W_hh rows for the NEXT step's gates run on
    // warps 0..3 underneath, in the issue slots the critical warps leave free.
    //   warps 0..6 : poll the 128-word segment `warp` of h_t (4 words per lane) -> h_s
    //   warps 4..7 : fc1 rows 4 (warp-4) .. +3 of this CTA's 16 (lane: 28 columns of all four rows), then fc2 column-partials of
    //                classes 64 (warp-4) + 2 lane + {0,1} over the 16 own r values -> 8-byte st.async into the owning rank's inbox
    //   warps 0..3 : W_hh rows 6 warp .. +5 of the CTA's 24 (lane: 28 columns of all six rows) + b_hh -> hb_s
    //   warp 7     : additionally the sequential chain (gates, reduce-scatter sum, all-gather, softmax, sample)
    const bool is_f = warp >= 4;
    const int fw = warp - 4;
    bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    // fc1 / W_hh weights: row-major pairs of adjacent columns, columns 128 j + 4 lane + {0,1 | 2,3}, j = 0..6 (conflict-free LDS.128)
    float2 wrow[6][14];
    {
        const int nrow = is_f ? 4 : 6;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const float* src;
            if (is_f) {
                src = p.fc1_w + static_cast<int64_t>(rank * CL_FR + 4 * fw + (r < 4 ? r : 0)) * CL_H;
            } else {
                const int row = 6 * warp + r, g = row / CL_U, u = row % CL_U;
                src = p.w_hh + static_cast<int64_t>(g * CL_H + cta * CL_U + u) * CL_H;
            }
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrow) v = __ldg(reinterpret_cast<const float4*>(src + 128 * j + 4 * lane));
                wrow[r][2 * j] = make_float2(v.x, v.y); wrow[r][2 * j + 1] = make_float2(v.z, v.w);
            }
        }
    }
    // fc2 (warps 4..7): classes c0 = 64 fw + 2 lane and c0 + 1 over the own r values k = 0..15, as (c0, c0 + 1) pairs
    float2 w2p[CL_FR];
    float2 b1p[2];                      // fc1 bias of the two rows this lane's 8-lane group finishes (see below) -- only [0] used
    unsigned rs_dst = 0, rs_mbar = 0;
    const unsigned inbox_a = static_cast<unsigned>(__cvta_generic_to_shared(&inboxf[0]));
    const unsigned lg_a = static_cast<unsigned>(__cvta_generic_to_shared(&lgf[0]));
    float fc1_bias = 0.f;
    if (is_f) {
        const int c0 = 64 * fw + 2 * lane;
#pragma unroll
        for (int k = 0; k < CL_FR; ++k)
            w2p[k] = make_float2(__ldg(p.fc2_w + static_cast<int64_t>(c0) * CL_FC + rank * CL_FR + k),
                                 __ldg(p.fc2_w + static_cast<int64_t>(c0 + 1) * CL_FC + rank * CL_FR + k));
        const unsigned dst_rank = static_cast<unsigned>(c0 >> 4);
        rs_dst = cl_mapa(inbox_a + (rank * CL_FR + (c0 & 15)) * 4, dst_rank);
        rs_mbar = cl_mapa(mbar_a, dst_rank);
        fc1_bias = __ldg(p.fc1_b + rank * CL_FR + 4 * fw + (b4 ? 2 : 0) + (b3 ? 1 : 0));
    } else {
#pragma unroll
        for (int k = 0; k < CL_FR; ++k) w2p[k] = make_float2(0.f, 0.f);
    }
    (void)b1p;
    float whh_bias = 0.f;               // W warps: b_hh of the row this lane's 4-lane group finishes
    {
        const int row = 6 * (warp & 3) + (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
        if (!is_f && row < 6 * (warp & 3) + 6) whh_bias = __ldg(p.b_hh + (row / CL_U) * CL_H + cta * CL_U + row % CL_U);
    }

    // ---- h_t : poll this warp's segment and put it into shared memory (warps 0..6)
    const ll_word* hsrc = p.hbuf + 128 * warp + 4 * lane;
    auto poll_h = [&](int t) {
        const uint32_t tag = static_cast<uint32_t>(t) + 1u;
        const ll_word* src = hsrc + (t & 1) * CL_H;
        if (p.poll_delay) { const long long t1 = clock64(); while (clock64() - t1 < p.poll_delay) {} }
        const long long t0 = clock64();
        unsigned n = 0, rounds = 0;
        ll_word a0, a1, c0, c1;
        // one probe in flight.  Measured alternatives, both slower: back-to-back probes (r02 v2) and a second probe staggered
        // 100-200 cycles behind the first (2.21 vs 1.88 us/step) -- loads pending on a line delay the stores they wait for.
        for (;;) {
            ++rounds;
            ll_load2(src, a0, a1);
            ll_load2(src + 2, c0, c1);
            const bool ok = ll_tag(a0) == tag && ll_tag(a1) == tag && ll_tag(c0) == tag && ll_tag(c1) == tag;
            if (__all_sync(0xffffffffu, ok)) break;
            if (abort_flag || ((++n & 255u) == 0 && clock64() - t0 > LL_TIMEOUT_CYCLES)) { CL_FAIL(); break; }
        }
        const float4 hv = make_float4(ll_val(a0), ll_val(a1), ll_val(c0), ll_val(c1));
        *reinterpret_cast<float4*>(&h_s[128 * warp + 4 * lane]) = hv;
        CL_TRACE(16 + warp, t, FU(hv.x))
        if constexpr (TRACE) {
            if (tracing && t >= p.trace_t0 && t < p.trace_t0 + p.trace_n) p.trace[(t - p.trace_t0) * CL_TRACE_STRIDE + 24 + warp] = rounds;
        }
        return hv.w;
    };
    // ---- NR row sums of (wrow . h_t) over this lane's 28 columns
    auto row_dots = [&](float (&s)[8], const int nr) {
        float2 acc[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const float4 hv = *reinterpret_cast<const float4*>(&h_s[128 * j + 4 * lane]);
            const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                if (r < nr) acc[r] = ffma2(wrow[r][2 * j + 1], h23, ffma2(wrow[r][2 * j], h01, acc[r]));
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) s[r] = (r < nr) ? acc[r < 6 ? r : 0].x + acc[r < 6 ? r : 0].y : 0.f;
    };
    // ---- fc1 rows of this warp (4 values per lane -> transposing butterfly: 16, 8, then plain 4, 2, 1), ReLU -> r_s
    auto fc1_rows = [&]() -> float {
        float s[8];
        row_dots(s, 4);
        float v2[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) v2[i] = (b4 ? s[i + 2] : s[i]) + __shfl_xor_sync(0xffffffffu, b4 ? s[i] : s[i + 2], 16);
        float v1 = (b3 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b3 ? v2[0] : v2[1], 8);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 4);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        const float r = fmaxf(v1 + fc1_bias, 0.f);
        if ((lane & 7) == 0) r_s[4 * fw + (b4 ? 2 : 0) + (b3 ? 1 : 0)] = r;
        return r;
    };
    // ---- fc2 column-partials of this lane's two classes over the 16 own r values -> the owning rank's inbox
    auto fc2_partial_send = [&](int par) -> float {
        const float4 r0 = *reinterpret_cast<const float4*>(&r_s[0]), r1 = *reinterpret_cast<const float4*>(&r_s[4]);
        const float4 r2 = *reinterpret_cast<const float4*>(&r_s[8]), r3 = *reinterpret_cast<const float4*>(&r_s[12]);
        const float rr[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
        float2 a[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) a[c] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CL_FR; ++k) a[k & 3] = ffma2(w2p[k], make_float2(rr[k], rr[k]), a[k & 3]);
        const float sx = (a[0].x + a[1].x) + (a[2].x + a[3].x), sy = (a[0].y + a[1].y) + (a[2].y + a[3].y);
        st_async_v2(rs_dst + par * CL_XBYTES, sx, sy, rs_mbar + 8 * par);
        return sx;
    };

    // ---- reduce-scatter sum + all-gather send (warps 4..7, each redundantly -- they sit on four different schedulers): wait for
    // the 16 ranks' column-partials of the 16 own logits, sum them (lane: sources lane/4 and lane/4 + 8, logit quad lane % 4),
    // and send the 16 logits to ranks 4 fw .. 4 fw + 3 only: a DSMEM store occupies the LSU per active lane, so one warp sending
    // to all 16 ranks (64 lane-operations) took ~250 cycles; four warps x 16 lanes take ~64.
    const float4 b2q = is_f ? __ldg(reinterpret_cast<const float4*>(p.fc2_b + rank * CL_FR) + (lane & 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned ag_dst = cl_mapa(lg_a + (rank * CL_FR + 4 * (lane & 3)) * 4, static_cast<unsigned>(4 * (fw & 3) + ((lane >> 2) & 3)));
    const unsigned ag_mbar = cl_mapa(mbar_a + 16, static_cast<unsigned>(4 * (fw & 3) + ((lane >> 2) & 3)));
    auto rs_sum_ag_send = [&](int t, float (&lo)[4]) {
        const int par = t & 1;
        const unsigned phase = (static_cast<unsigned>(t) >> 1) & 1u;
        const long long t0 = clock64();
        unsigned n = 0;
        while (!mbar_try_wait(mbar_a + 8 * par, phase)) {
            if (abort_flag || ((++n & 63u) == 0 && clock64() - t0 > LL_TIMEOUT_CYCLES)) { CL_FAIL(); break; }
        }
        const float* ib = &inboxf[par * CL_Q + (lane >> 2) * CL_FR + 4 * (lane & 3)];
        const float4 pa = *reinterpret_cast<const float4*>(ib);
        const float4 pb = *reinterpret_cast<const float4*>(ib + 8 * CL_FR);
        if (TRACE && warp == 7) { CL_TRACE(4, t, FU(pa.x)) }
        lo[0] = pa.x + pb.x; lo[1] = pa.y + pb.y; lo[2] = pa.z + pb.z; lo[3] = pa.w + pb.w;
        // (reading four sources per lane to save a shuffle round was measured slower: 225 vs 200 cycles)
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) lo[i] += __shfl_xor_sync(0xffffffffu, lo[i], o);
        }
        lo[0] += b2q.x; lo[1] += b2q.y; lo[2] += b2q.z; lo[3] += b2q.w;   // logits of classes 16 rank + 4 (lane % 4) + {0..3}
        if (TRACE && warp == 7) { CL_TRACE(5, t, FU(lo[0])) }
        if (!teacher && lane < 16) st_async_v4(ag_dst + par * CL_XBYTES, lo[0], lo[1], lo[2], lo[3], ag_mbar + 8 * par);
    };

    if (warp < 4) {
        // =============================================================== W warps: poll, then W_hh rows for the next step's gates
        for (int t = 0; t < L; ++t) {
            // this CTA publishes h_t at (about) the same time as everybody else: do not poll L2 before that.  A hardware
            // barrier, not a shared-memory spin: spinning warps would steal issue and LDS slots from the chain warp.
            bar_sync(5, CL_THREADS);
            if (TRACE && warp == 0) { CL_TRACE(8, t, static_cast<unsigned>(abort_flag)) }
            const float hw = poll_h(t);
            if (TRACE && warp == 0) { CL_TRACE(9, t, FU(hw)) }
            bar_sync(1, CL_THREADS);                 // all 896 values of h_t are in h_s
            bar_sync(6, CL_THREADS);                 // the critical warps have sent their logits: the schedulers are free now
            float s[8];
            row_dots(s, 6);
            if (TRACE && warp == 0) { CL_TRACE(10, t, FU(s[0])) }
            // 8 slots (rows 6, 7 are zero padding): transposing butterfly 16, 8, 4, then plain 2, 1
            float v4[4], v2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) v4[i] = (b4 ? s[i + 4] : s[i]) + __shfl_xor_sync(0xffffffffu, b4 ? s[i] : s[i + 4], 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) v2[i] = (b3 ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? v4[i] : v4[i + 2], 8);
            float v1 = (b2 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? v2[0] : v2[1], 4);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            const int slot = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
            if ((lane & 3) == 0 && slot < 6) hb_s[6 * warp + slot] = v1 + whh_bias;
            bar_arrive(4, 160);                      // -> chain warp: W_hh h_t + b_hh of this warp's six rows is in hb_s
            if (TRACE && warp == 0) { CL_TRACE(11, t, FU(v1)) }
        }
    } else if (warp < CL_MW) {
        // =============================================================== F warps 4..6: poll, fc1 rows, fc2 partials
        for (int t = 0; t < L; ++t) {
            bar_sync(5, CL_THREADS);
            poll_h(t);
            bar_sync(1, CL_THREADS);
            const float r = fc1_rows();
            if (TRACE && warp == 4) { CL_TRACE(13, t, FU(r)) }
            bar_sync(3, 128);                        // r_s holds all 16 relu(fc1) values of this CTA
            const float s0 = fc2_partial_send(t & 1);
            if (TRACE && warp == 4) { CL_TRACE(14, t, FU(s0)) }
            if (!teacher) {
                float lo[4];
                rs_sum_ag_send(t, lo);
            }
            bar_arrive(6, CL_THREADS);
        }
    } else {
        // =============================================================== C warp: the sequential chain (+ its share of fc1 / fc2)
        const int gu = cta * CL_U + (lane & 7);
        float hown = 0.f;
        float hb_r = __ldg(p.b_hh + gu), hb_z = __ldg(p.b_hh + CL_H + gu), hb_n = __ldg(p.b_hh + 2 * CL_H + gu);   // W_hh h_{-1} = 0
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
        // W_hh h_t + b_hh of the own units for the next step's gates (hb_s is written by warps 0..3 while the logits travel)
        auto hb_load = [&]() {
            bar_sync(4, 160);
            hb_r = hb_s[lane & 7]; hb_z = hb_s[CL_U + (lane & 7)]; hb_n = hb_s[2 * CL_U + (lane & 7)];
        };
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
            bar_arrive(5, CL_THREADS);               // polling warps: h_t is on its way
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
            bar_sync(1, CL_THREADS);                 // all 896 values of h_t are in h_s
            float lo[4] = {0.f, 0.f, 0.f, 0.f};
            {
                const float r = fc1_rows();
                CL_TRACE(2, t, FU(r))
                bar_sync(3, 128);                    // r_s complete
                const float s0 = fc2_partial_send(par);
                CL_TRACE(3, t, FU(s0))
            }
            {
                rs_sum_ag_send(t, lo);
                if (p.out_logits != nullptr && cta < CL_S && lane < 4)
                    *reinterpret_cast<float4*>(p.out_logits + static_cast<int64_t>(t) * CL_Q + rank * CL_FR + 4 * lane) =
                        make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            bar_arrive(6, CL_THREADS);               // W warps: go
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
                    hb_load();      // issued under the last shuffle's latency; the W warps finished during the all-gather hop
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
            if (teacher) hb_load();
            {
                // the next step's conditioning frame is already known here
                const bool nf = (frame_left == 0);
                gh_r = __fadd_rn(nf ? gn_r : g_r, hb_r);
                gh_z = __fadd_rn(nf ? gn_z : g_z, hb_z);
                CL_TRACE(6, t, FU(gh_z))
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

// ------------------------------------------------------------------------------------------------
// Exchange floor of THIS kernel (diagnostic, bench.py): the bare grid-scope all-gather of a step -- every CTA publishes its 8
// LL words, warps 0..6 poll the 896 words after the same first-probe delay, one CTA barrier -- with no compute in between.
// per-step floor = t_smem + ONE such exchange.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS, 1) exchange_floor_cluster_kernel(ll_word* buf, int iters, int poll_delay, long long* cycles) {
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), cta = blockIdx.x;
    __shared__ int fail;
    if (tid == 0) fail = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        if (warp == CL_MW) {
            if (lane < CL_U) ll_store(buf + par * CL_H + cta * CL_U + lane, 0.f, static_cast<uint32_t>(it));
            bar_arrive(5, CL_THREADS);
        } else {
            bar_sync(5, CL_THREADS);
            if (poll_delay) { const long long t1 = clock64(); while (clock64() - t1 < poll_delay) {} }
            const ll_word* src = buf + par * CL_H + 128 * warp + 4 * lane;
            const long long ts = clock64();
            for (;;) {
                ll_word a0, a1, c0, c1;
                ll_load2(src, a0, a1);
                ll_load2(src + 2, c0, c1);
                const uint32_t tag = static_cast<uint32_t>(it);
                const bool ok = ll_tag(a0) == tag && ll_tag(a1) == tag && ll_tag(c0) == tag && ll_tag(c1) == tag;
                if (__all_sync(0xffffffffu, ok)) break;
                if (fail || clock64() - ts > LL_TIMEOUT_CYCLES) { fail = 1; break; }
            }
        }
        bar_sync(1, CL_THREADS);
    }
    if (tid == 0) cycles[cta] = fail ? -1 : clock64() - t0;
}

int ar_cluster_exchange_floor(void* workspace, size_t workspace_bytes, int iters, double* mean_cycles, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ host
// first probe of the grid hop this many cycles after the CTA's own publish (tools/gen_delay_sweep.py on three boards: 300 -> 1.874 /
// 1.886 / 1.886 us per step, 350-360 -> 1.856 / 1.861 / 1.871, 450 -> - / 1.869 / 1.849; an earlier probe comes back empty and costs a round)
static int g_cl_poll_delay = 360, g_cl_poll_mode = 0;
int g_cl_enable = 1;

// 1 = the device can hold the 7 x 16 cluster grid at one CTA per SM, 0 = it cannot (fall back to ar_kernel), cached per device
int ar_cluster_supported() {
    // VQCPC_AR_CLUSTER=0 / 1: explicit choice.  Otherwise the cluster kernel is used unless Nsight Compute is attached to this
    // process: ncu cannot launch this grid in any replay mode (a cooperative launch that takes every 16-CTA cluster slot of the
    // device; profiles/r02_launches_summary.md) and takes the process down with it, so a profiled run uses the 128-CTA ar_kernel,
    // which it can replay.  The two kernels agree to summation-order noise (logits within 1e-5, tests/test_gpu_parity.py).
    static const int policy = [] {
        const char* e = getenv("VQCPC_AR_CLUSTER");
        if (e != nullptr && e[0] != '\0') return e[0] == '0' ? 0 : 1;
        if (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr || getenv("NV_NSIGHT_INJECTION_PORT_BASE") != nullptr) return 0;
        return 1;
    }();
    if (!policy) return 0;
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

int ar_cluster_exchange_floor(void* workspace, size_t workspace_bytes, int iters, double* mean_cycles, cudaStream_t s) {
    const size_t need = sizeof(ll_word) * 2 * CL_H + sizeof(long long) * CL_CTAS;
    VQ_ARG(workspace_bytes >= need, "exchange_floor: workspace too small");
    if (device_sm_count() < CL_CTAS) { set_error("exchange_floor: needs %d SMs", CL_CTAS); return VQCPC_ERR_DEVICE; }
    ll_word* buf = static_cast<ll_word*>(workspace);
    long long* cyc = reinterpret_cast<long long*>(buf + 2 * CL_H);
    int delay = g_cl_poll_delay;
    VQ_CUDA(cudaMemsetAsync(workspace, 0, need, s));
    void* args[] = {&buf, &iters, &delay, &cyc};
    VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(exchange_floor_cluster_kernel), dim3(CL_CTAS), dim3(CL_THREADS), args, 0, s));
    count_launch(1);
    long long host[CL_CTAS];
    VQ_CUDA(cudaMemcpyAsync(host, cyc, sizeof(host), cudaMemcpyDeviceToHost, s));
    VQ_CUDA(cudaStreamSynchronize(s));
    double sum = 0;
    for (int i = 0; i < CL_CTAS; ++i) {
        if (host[i] < 0) { set_error("exchange_floor: timed out"); return VQCPC_ERR_TIMEOUT; }
        sum += static_cast<double>(host[i]);
    }
    *mean_cycles = sum / CL_CTAS / iters;
    return VQCPC_OK;
}

}  // namespace vqcpc
