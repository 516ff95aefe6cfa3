// Batched sample loop on tcgen05 for 65..128 utterances per launch.
//
// Same step as vocoder_batch.cu, but the contraction [W_hh ; fc1] (23 rows per CTA, padded to 32) x h_t (896 x 128
// utterances) is ONE tcgen05 GEMM per step and per CTA:
//   D[utterance 0..127][row 0..31]  =  sum_k  H[utterance][k] * W[row][k]        (M = 128, N = 32, K = 896)
//   * A = h_t: every CTA publishes its 7 hidden units of all utterances as bf16 hi/lo into a global (128 x [hi 896 | lo
//     896]) matrix (double buffered by step parity); after grid barrier 1 each CTA streams the whole matrix through a
//     4-stage shared-memory ring with TMA (28 tiles of 128 x 64, SWIZZLE_128B) -- ONE pass over h_t per step where the
//     mma.sync kernels make two per group;
//   * B = the CTA's 32 weight rows as bf16 hi/lo planes, resident in 112 KB of shared memory for the whole kernel
//     (28 tiles of 32 x 64, loaded once by TMA from a per-launch prepared copy);
//   * three MMAs per (tile, k-step): hi*hi, hi*lo, lo*hi -- fp32 accumulation in TMEM (32 columns);
//   * D lands with lane = utterance: epilogue thread u reads its 23 sums with one tcgen05.ld -- the 21 W_hh sums stay in
//     REGISTERS for the next step's gates (no shared-memory reduction at all), the two fc1 sums go out as relu(fc1 h).
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = gates + epilogue + fc2 (thread = utterance), 6 = barrier poller,
// 7 = sampler (decoupled: talks to the rest only through LL words in global memory).
// Exchanges per step as in vocoder_batch.cu: two grid barriers (h_t, relu(fc1 h_t)), logits and codes as LL words.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {
namespace abtc {

constexpr int H = 896, G3 = 2688, FC = 256, Q = 256;
constexpr int CTAS = 128, U = 7, R = 2, NROW = 21, ROWS = 23, NPAD = 32;
constexpr int BM = 128;                       // utterance slots = M of the MMA
constexpr int THREADS = 256;
constexpr int STAGES = 4;
constexpr int KCH = H / 64;                   // 14 k-chunks of 64 columns per plane
constexpr uint32_t WTILE = NPAD * 128;        // 4 KB: 32 rows x 64 bf16
constexpr uint32_t ATILE = BM * 128;          // 16 KB: 128 rows x 64 bf16
constexpr uint32_t W_BYTES = 2 * KCH * WTILE; // 112 KB
constexpr int ES = Q * NROW;
constexpr size_t SMEM = W_BYTES + STAGES * ATILE + sizeof(float) * (ES + (FC / 4) * R * 4) + 1024;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NPAD >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
constexpr int X_INIT = 128;
constexpr int MAIN = 224;                     // warps 0..6 meet at named barrier 1

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (*abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) { *abort_flag = 1; atomicExch(status, VQCPC_ERR_TIMEOUT); return false; }
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
__device__ __forceinline__ float ld_strong(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

struct Params {
    const float* b_hh; const float* fc1_b; const float* fc2_w; const float* fc2_b; const float* eprime; const float* lut;
    const float* G; const float* uniforms; const int64_t* x_in;
    float* out_wav; int32_t* out_codes; float* out_logits;
    __nv_bfloat16* Hbuf;     // [2 parity][128 utterances][hi 896 | lo 896]
    float* rT;               // [256][128]
    ll_word* oLL;            // [128][256]
    ll_word* xLL;            // [128]
    ll_word* flags;          // [128][16]
    int* status;
    long long g_stride;
    int L, upsample, nb;
    long long* trace;        // optional (debug): [n][8] clock64 phase stamps of a gate thread of CTA trace_cta
    int trace_cta, trace_t0, trace_n;
};

// per-launch copy of every CTA's 32 weight rows as bf16 hi/lo planes: Wp[(cta * 32 + r)][hi 896 | lo 896]
__global__ void prep_weights_kernel(const float* __restrict__ w_hh, const float* __restrict__ fc1_w, __nv_bfloat16* __restrict__ Wp) {
    const int row = blockIdx.x;                    // 0 .. 128*32
    const int cta = row / NPAD, r = row % NPAD;
    const float* src = nullptr;
    if (r < NROW) src = w_hh + static_cast<int64_t>((r % 3) * H + cta * U + r / 3) * H;
    else if (r < ROWS) src = fc1_w + static_cast<int64_t>(cta * R + (r - NROW)) * H;
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        const float v = src ? __ldg(src + k) : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        Wp[static_cast<int64_t>(row) * 2 * H + k] = h;
        Wp[static_cast<int64_t>(row) * 2 * H + H + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

__device__ __forceinline__ void grid_signal(ll_word* flags, uint32_t tag) {
    bar_sync(1, MAIN);
    if (threadIdx.x == 0) {
        __threadfence();
        ll_store(flags + blockIdx.x * 16, 0.f, tag);
    }
}
__device__ __forceinline__ bool grid_wait(ll_word* flags, uint32_t tag, volatile int* abort_flag, int* status) {
    if ((threadIdx.x >> 5) == 6) {
        const int ln = threadIdx.x & 31;
        const long long t0 = clock64();
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t seen = ll_tag(ll_load(flags + (32 * k + ln) * 16));
                ok = ok && (static_cast<int32_t>(seen - tag) >= 0);
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (*abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) {
                *abort_flag = 1;
                if (ln == 0) atomicExch(status, VQCPC_ERR_TIMEOUT);
                break;
            }
        }
        __threadfence();
    }
    bar_sync(1, MAIN);
    return *abort_flag == 0;
}

__global__ void __launch_bounds__(THREADS, 1)
ar_batch_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h0,
                   const __grid_constant__ CUtensorMap map_h1, Params p) {
    extern __shared__ __align__(1024) unsigned char t3_smem[];
    __shared__ __align__(8) uint64_t w_bar, full_bar[STAGES], empty_bar[STAGES], dfull_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ volatile int abort_flag;
    __shared__ float bhh_s[NROW], b1_s[R], b2_s[R];

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(t3_smem) + 1023) & ~uintptr_t(1023));
    unsigned char* w_s = smem;                                   // [plane 2][chunk 14][32 rows][128 B]
    unsigned char* a_s = smem + W_BYTES;                         // [stage][128 rows][128 B]
    float* Es = reinterpret_cast<float*>(a_s + STAGES * ATILE);  // [x][21]
    float* W2s = Es + ES;                                        // [c4][r][4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const bool teacher = p.x_in != nullptr;
    const int L = p.L, nb = p.nb;

    if (tid == 0) {
        abort_flag = 0;
        mbar_init(&w_bar, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&dfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < (FC / 4) * R; i += THREADS) {
        const int c4 = i / R, r = i % R;
        reinterpret_cast<float4*>(W2s)[i] = __ldg(reinterpret_cast<const float4*>(p.fc2_w + static_cast<int64_t>(cta * R + r) * FC + 4 * c4));
    }
    for (int i = tid; i < ES; i += THREADS) {
        const int x = i / NROW, j = i % NROW;
        Es[i] = __ldg(p.eprime + static_cast<int64_t>(x) * G3 + (j % 3) * H + cta * U + j / 3);
    }
    if (tid < NROW) bhh_s[tid] = __ldg(p.b_hh + (tid % 3) * H + cta * U + tid / 3);
    if (tid < R) { b1_s[tid] = __ldg(p.fc1_b + cta * R + tid); b2_s[tid] = __ldg(p.fc2_b + cta * R + tid); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 7) {
        // ------------------------------------------------------------------ sampler warp (decoupled): CTA b <-> utterance b
        if (teacher || cta >= nb) return;
        const int b = cta;
        for (int t = 0; t < L; ++t) {
            const ll_word* src = p.oLL + b * Q + 8 * lane;
            float ov[8];
            const long long t0 = clock64();
            for (;;) {
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    ll_word w0, w1;
                    ll_load2(src + k, w0, w1);
                    ok = ok && ll_tag(w0) == static_cast<uint32_t>(t + 1) && ll_tag(w1) == static_cast<uint32_t>(t + 1);
                    ov[k] = ll_val(w0); ov[k + 1] = ll_val(w1);
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if (abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) {
                    abort_flag = 1;
                    if (lane == 0) atomicExch(p.status, VQCPC_ERR_TIMEOUT);
                    return;
                }
            }
            // inverse-CDF sample: lane l holds classes 8l .. 8l+7
            float m = ov[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) m = fmaxf(m, ov[k]);
            m = warp_max(m);
            float c[8];
            float run = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { run += __expf(ov[k] - m); c[k] = run; }
            float incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const float excl = incl - run;
            const float thr = __ldg(p.uniforms + static_cast<int64_t>(b) * L + t) * __shfl_sync(0xffffffffu, incl, 31);
            int loc = 8;
#pragma unroll
            for (int k = 7; k >= 0; --k) if (excl + c[k] > thr) loc = k;
            const unsigned hit = __ballot_sync(0xffffffffu, loc < 8);
            int x = Q - 1;
            if (hit != 0u) {
                const int srcl = __ffs(hit) - 1;
                x = 8 * srcl + __shfl_sync(0xffffffffu, loc, srcl);
            }
            if (lane == 0) {
                ll_store(p.xLL + b, __int_as_float(x), static_cast<uint32_t>(t + 1));
                if (p.out_wav) p.out_wav[static_cast<int64_t>(b) * L + t] = __ldg(p.lut + x);
                if (p.out_codes) p.out_codes[static_cast<int64_t>(b) * L + t] = x;
            }
        }
        return;
    }

    // ---- resident weights: 28 tiles of 32 x 64 bf16 by TMA (row cta * 32 of the prepared copy)
    if (tid == 0) {
        mbar_expect_tx(&w_bar, W_BYTES);
        for (int pl = 0; pl < 2; ++pl)
            for (int c = 0; c < KCH; ++c) tma_load_2d(w_s + (pl * KCH + c) * WTILE, &map_w, pl * H + c * 64, cta * NPAD, &w_bar);
    }

    uint32_t tag = 0;
    // state of the gate / epilogue threads (thread = utterance u): W_hh h + b_hh of the previous step, own hidden units,
    // conditioning of the current frame, previous code
    const int quarter = warp & 3;
    const int u = quarter * 32 + lane;
    float hh[NROW], hown[U], Gc[NROW];
    int xcur = X_INIT;
#pragma unroll
    for (int j = 0; j < NROW; ++j) { hh[j] = 0.f; Gc[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < U; ++j) hown[j] = 0.f;
    int stage = 0;
    uint32_t ring_phase = 0, dphase = 0;
    bool wloaded = false;

    const bool tracing = p.trace != nullptr && cta == p.trace_cta && tid == 64;
#define T3_TRACE(k) if (tracing && t >= p.trace_t0 && t < p.trace_t0 + p.trace_n) p.trace[(t - p.trace_t0) * 8 + (k)] = clock64();
    for (int t = 0; t < L; ++t) {
        const int par = t & 1;
        T3_TRACE(0)
        // ------------------------------------------------------------------ G: gates of step t, h_t published (bf16 hi/lo)
        if (warp >= 2 && warp <= 5) {
            if (teacher) {
                xcur = (u < nb) ? (static_cast<int>(__ldg(p.x_in + static_cast<int64_t>(u) * L + t)) & (Q - 1)) : 0;
            } else if (t > 0) {
                int xv = 0;
                if (u < nb) {
                    const long long t0 = clock64();
                    for (;;) {
                        const ll_word w = ll_load(p.xLL + u);
                        if (ll_tag(w) == static_cast<uint32_t>(t)) { xv = __float_as_int(ll_val(w)); break; }
                        if (abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) { abort_flag = 1; atomicExch(p.status, VQCPC_ERR_TIMEOUT); break; }
                    }
                }
                xcur = xv;
            }
            if (t % p.upsample == 0) {
                const int frame = t / p.upsample;
#pragma unroll
                for (int j = 0; j < NROW; ++j)
                    Gc[j] = (u < nb) ? __ldg(p.G + u * p.g_stride + static_cast<int64_t>(frame) * G3 + (j % 3) * H + cta * U + j / 3) : 0.f;
            }
            __nv_bfloat16* hrow = p.Hbuf + (static_cast<int64_t>(par) * BM + u) * 2 * H + cta * U;
            const float* e = &Es[xcur * NROW];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const float hr = (t == 0 ? 0.f : hh[3 * j]) + bhh_s[3 * j], hz = (t == 0 ? 0.f : hh[3 * j + 1]) + bhh_s[3 * j + 1];
                const float hn_ = (t == 0 ? 0.f : hh[3 * j + 2]) + bhh_s[3 * j + 2];
                const float r = sigmoid_fast(__fadd_rn(__fadd_rn(e[3 * j], Gc[3 * j]), hr));
                const float z = sigmoid_fast(__fadd_rn(__fadd_rn(e[3 * j + 1], Gc[3 * j + 1]), hz));
                const float n = tanh_fast(__fmaf_rn(r, hn_, __fadd_rn(e[3 * j + 2], Gc[3 * j + 2])));
                const float hv = __fmaf_rn(z, __fsub_rn(hown[j], n), n);
                hown[j] = hv;
                const __nv_bfloat16 hb = __float2bfloat16_rn(hv);
                hrow[j] = hb;
                hrow[H + j] = __float2bfloat16_rn(hv - __bfloat162float(hb));
            }
        }
        T3_TRACE(1)
        grid_signal(p.flags, ++tag);
        if (!grid_wait(p.flags, tag, &abort_flag, p.status)) return;                // barrier 1: h_t complete
        T3_TRACE(2)

        // ------------------------------------------------------------------ the step's GEMM: TMA ring -> tcgen05 -> TMEM
        if (warp == 0) {
            if (lane == 0) {
                asm volatile("fence.proxy.async;" ::: "memory");                   // generic-proxy writes (other CTAs) -> TMA reads
                const CUtensorMap* mh = par ? &map_h1 : &map_h0;
                int st = stage; uint32_t ph = ring_phase;
                bool ok = true;
                for (int i = 0; i < 2 * KCH && ok; ++i) {
                    const int c = i >> 1, pl = i & 1;                              // hi tile of chunk c, then its lo tile
                    ok = mbar_wait(&empty_bar[st], ph ^ 1, &abort_flag, p.status);
                    if (!ok) break;
                    mbar_expect_tx(&full_bar[st], ATILE);
                    tma_load_2d(a_s + st * ATILE, mh, pl * H + c * 64, 0, &full_bar[st]);
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                bool ok = true;
                if (!wloaded) ok = mbar_wait(&w_bar, 0, &abort_flag, p.status);
                int st = stage; uint32_t ph = ring_phase;
                for (int i = 0; i < 2 * KCH && ok; ++i) {
                    const int c = i >> 1, pl = i & 1;
                    ok = mbar_wait(&full_bar[st], ph, &abort_flag, p.status);
                    if (!ok) break;
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(a_s + st * ATILE));
                    const uint64_t bhi = umma_desc_sw128(smem_u32(w_s + (0 * KCH + c) * WTILE));
                    const uint64_t blo = umma_desc_sw128(smem_u32(w_s + (1 * KCH + c) * WTILE));
                    if (pl == 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, adesc + 2 * k, bhi + 2 * k, IDESC, (c > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, adesc + 2 * k, blo + 2 * k, IDESC, 1u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, adesc + 2 * k, bhi + 2 * k, IDESC, 1u);
                    }
                    tc_commit(&empty_bar[st]);
                    if (++st == STAGES) { st = 0; ph ^= 1; }
                }
                tc_commit(&dfull_bar);
            }
        }
        wloaded = true;
        // every thread advances its view of the ring by the 28 tiles of this step
        {
            const int adv = 2 * KCH;
            const int ns = stage + adv;
            if (((ns / STAGES) & 1) != 0) ring_phase ^= 1;
            stage = ns % STAGES;
        }
        // ------------------------------------------------------------------ epilogue: lane = utterance
        if (warp >= 2 && warp <= 5) {
            bool ok = mbar_wait(&dfull_bar, dphase, &abort_flag, p.status);
            ok = __all_sync(0xffffffffu, ok);
            T3_TRACE(3)
            if (ok) {
                tc_fence_after();
                uint32_t v[32];
                tc_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16), v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < NROW; ++j) hh[j] = __uint_as_float(v[j]);
#pragma unroll
                for (int r = 0; r < R; ++r) p.rT[(cta * R + r) * BM + u] = fmaxf(__uint_as_float(v[NROW + r]) + b1_s[r], 0.f);
                tc_fence_before();
            }
        }
        dphase ^= 1;
        T3_TRACE(4)
        grid_signal(p.flags, ++tag);
        if (!grid_wait(p.flags, tag, &abort_flag, p.status)) return;                // barrier 2: relu(fc1 h_t) complete
        T3_TRACE(5)

        // ------------------------------------------------------------------ P3: the CTA's two fc2 rows for utterance u
        if (warp >= 2 && warp <= 5) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 1
            for (int k0 = 0; k0 < FC; k0 += 64) {
                float rv[64];
#pragma unroll
                for (int i = 0; i < 64; ++i) rv[i] = ld_strong(p.rT + (k0 + i) * BM + u);
                const float4* w2g = reinterpret_cast<const float4*>(W2s) + (k0 / 4) * R;
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) {
                    const float4 w0 = w2g[c4 * R], w1 = w2g[c4 * R + 1];
                    a0 = fmaf(w0.x, rv[4 * c4], a0); a0 = fmaf(w0.y, rv[4 * c4 + 1], a0);
                    a0 = fmaf(w0.z, rv[4 * c4 + 2], a0); a0 = fmaf(w0.w, rv[4 * c4 + 3], a0);
                    a1 = fmaf(w1.x, rv[4 * c4], a1); a1 = fmaf(w1.y, rv[4 * c4 + 1], a1);
                    a1 = fmaf(w1.z, rv[4 * c4 + 2], a1); a1 = fmaf(w1.w, rv[4 * c4 + 3], a1);
                }
            }
            if (u < nb) {
                const float o0 = a0 + b2_s[0], o1 = a1 + b2_s[1];
                if (!teacher) {
                    ll_store(p.oLL + u * Q + cta * R, o0, static_cast<uint32_t>(t + 1));
                    ll_store(p.oLL + u * Q + cta * R + 1, o1, static_cast<uint32_t>(t + 1));
                }
                if (p.out_logits != nullptr) {
                    float* dst = p.out_logits + (static_cast<int64_t>(u) * L + t) * Q + cta * R;
                    dst[0] = o0; dst[1] = o1;
                }
            }
        }
        T3_TRACE(6)
    }
#undef T3_TRACE
    tc_fence_before();
    bar_sync(1, MAIN);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
    }
}

typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows) {
    static PFN_encodeTiled3 fn = nullptr;
    if (fn == nullptr) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            set_error("ar_batch_tc: cuTensorMapEncodeTiled is unavailable");
            return VQCPC_ERR_CUDA;
        }
        fn = reinterpret_cast<PFN_encodeTiled3>(fp);
    }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("ar_batch_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return VQCPC_ERR_CUDA; }
    return VQCPC_OK;
}

// workspace: [Wp 4096 x 1792 bf16][H 2 x 128 x 1792 bf16][rT 256 x 128][oLL 128 x 256][xLL 128][flags 128 x 16]
constexpr size_t WP_BYTES = static_cast<size_t>(CTAS) * NPAD * 2 * H * 2;
constexpr size_t HB_BYTES = 2ull * BM * 2 * H * 2;
static size_t ws_bytes() {
    return WP_BYTES + HB_BYTES + sizeof(float) * FC * BM + sizeof(ll_word) * (static_cast<size_t>(BM) * Q + BM + CTAS * 16) + 1024;
}

}  // namespace abtc

extern long long* g_ab_trace;
extern int g_ab_trace_cta, g_ab_trace_t0, g_ab_trace_n;

size_t ar_batch_tc_workspace_bytes() { return align_up(abtc::ws_bytes(), 256); }

// one launch of up to 128 utterances; G / uniforms / x_in / outputs already offset to the launch's first utterance
int ar_batch_tc_launch(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int nb, int T2, int L,
                       void* ws, int* status, float* out_wav, int32_t* out_codes, float* out_logits, cudaStream_t stream) {
    using namespace abtc;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(ar_batch_tc_kernel), static_cast<int>(SMEM))) return rc;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    __nv_bfloat16* Wp = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* Hb = reinterpret_cast<__nv_bfloat16*>(base + WP_BYTES);
    float* rT = reinterpret_cast<float*>(base + WP_BYTES + HB_BYTES);
    ll_word* oLL = reinterpret_cast<ll_word*>(rT + FC * BM);
    ll_word* xLL = oLL + static_cast<size_t>(BM) * Q;
    ll_word* flags = xLL + BM;
    VQ_CUDA(cudaMemsetAsync(Hb, 0, ws_bytes() - 1024 - WP_BYTES, stream));
    prep_weights_kernel<<<CTAS * NPAD, 128, 0, stream>>>(w->ar_w_hh, w->fc1_w, Wp);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    CUtensorMap mw, mh0, mh1;
    int rc = make_map(&mw, Wp, CTAS * NPAD, 2 * H, NPAD);
    if (rc) return rc;
    if ((rc = make_map(&mh0, Hb, BM, 2 * H, BM))) return rc;
    if ((rc = make_map(&mh1, Hb + static_cast<size_t>(BM) * 2 * H, BM, 2 * H, BM))) return rc;
    Params p{};
    p.b_hh = w->ar_b_hh; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b; p.eprime = w->eprime; p.lut = w->mulaw_lut;
    p.G = G; p.uniforms = uniforms; p.x_in = x_in; p.out_wav = out_wav; p.out_codes = out_codes; p.out_logits = out_logits;
    p.Hbuf = Hb; p.rT = rT; p.oLL = oLL; p.xLL = xLL; p.flags = flags; p.status = status;
    p.g_stride = static_cast<long long>(T2) * G3;
    p.L = L; p.upsample = w->upsample_t; p.nb = nb;
    p.trace = g_ab_trace; p.trace_cta = g_ab_trace_cta; p.trace_t0 = g_ab_trace_t0; p.trace_n = g_ab_trace_n;
    void* args[] = {&mw, &mh0, &mh1, &p};
    VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ar_batch_tc_kernel), dim3(CTAS), dim3(THREADS), args, SMEM, stream));
    count_launch(1);
    return VQCPC_OK;
}

}  // namespace vqcpc
