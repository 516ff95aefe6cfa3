// Batched sample loop: up to 64 utterances advance together through one persistent cooperative kernel.
//
// The single-utterance kernel (vocoder.cu) is bound by three ~1 us grid-wide exchanges per step.  With a batch the
// same exchanges carry 64 utterances and the per-step contraction  [W_hh ; fc1] (23 x 896 per CTA) x h_t (896 x 64)
// becomes a small GEMM that runs on the TENSOR CORES:
//   * 128 CTAs x 256 threads.  CTA j owns hidden units 7j..7j+6 (21 W_hh rows), fc1 rows 2j,2j+1, fc2 rows 2j,2j+1.
//   * warp cc owns 112 of the 896 columns.  The utterances are the M side of mma.sync m16n8k16 (4 tiles of 16), the
//     weight rows the N side (23 rows -> 3 tiles of 8, one padding row).  The warp's slice of the weight rows lives in
//     REGISTERS for the whole kernel as B-fragments, split into bf16 hi/lo planes (84 registers per lane).
//   * the producers of h_t (the gate phase) publish it already split:  hi = bf16(h), lo = bf16(h - hi), stored in mma
//     A-fragment order [k-tile][utterance tile][lane][hi a0..a3, lo a0..a3], so a consumer lane fetches the eight A
//     registers of a (16 x 16) tile with two 16-byte strong loads and a warp reads 1 KB contiguous.  Three MMAs per (16 x 8 x 16) tile -- hi*hi, hi*lo, lo*hi -- accumulate in fp32
//     (the dropped lo*lo term is 2^-16 relative; logits stay within 1e-5 of the fp32 kernels).
//   * the eight column-chunk partials of a (row, utterance) meet in shared memory.
//   * h_t and relu(fc1 h_t) travel through plain global buffers separated by grid barriers (__threadfence + one LL
//     flag per CTA in its own 128-byte slot; split into signal / wait): 2 barriers per step.  The logits (to the 64
//     sampling CTAs) and the sampled codes (to everyone) travel as LL words -- value and step tag in one 8-byte
//     store, polled by the reader -- which costs one ~1 us hop instead of a fence + flag + poll round.
//   * a ninth warp per CTA polls the barriers and does the sampling, so the eight tensor-core warps never wait on it.
//   * step:  G   gates -> h_t                                               | barrier 1 |
//            P2a rows 16..23 (W_hh rows 16-20 + the two fc1 rows) -> r       | signal 2  | P2b | wait 2 |
//            P3  fc2 rows -> logits (LL)
//            sampler warp of CTA b: wait logits, sample utterance b, publish code (LL);  everyone polls the 64 codes
//     P2b = rows 0..15 (W_hh rows needed only by the NEXT step's gates): runs while barrier 2 is in flight and, after
//     P3, while the sampler warp works.

#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

constexpr int AB_H = 896, AB_G = 2688, AB_FC = 256, AB_Q = 256;
constexpr int AB_CTAS = 128, AB_U = 7, AB_R = 2, AB_NROW = 21, AB_ROWS = AB_NROW + AB_R;   // 23 rows ride the h stream
constexpr int AB_B = 64;                          // utterance slots per launch
constexpr int AB_THREADS = 288;                   // 8 tensor-core warps + 1 sampler / barrier-poller warp
constexpr int AB_SW = 8;                          // index of the sampler warp
constexpr int AB_CC = 8, AB_CHUNK = AB_H / AB_CC; // 112 columns per warp in the h stream
constexpr int AB_X_INIT = 128;

struct AbParams {
    const float* w_hh; const float* b_hh; const float* fc1_w; const float* fc1_b; const float* fc2_w; const float* fc2_b;
    const float* eprime; const float* lut;
    const float* G;          // (B, T2, 2688)
    const float* uniforms;   // (B, L)   generate mode
    const int64_t* x_in;     // (B, L)   teacher-forced mode
    float* out_wav; int32_t* out_codes; float* out_logits;
    uint32_t* hP;            // [2 parity][k-tile 56][m-tile 4][lane 32][hi a0..a3, lo a0..a3] bf16x2 (mma A-fragment order)
    float* rT;               // [256][64]
    ll_word* oLL;            // [64][256]  logits as LL words {tag = step + 1, value}, utterance-major
    ll_word* xLL;            // [64]       sampled codes as LL words
    ll_word* flags;          // [128][16]
    int* status;
    long long g_stride;
    int L, upsample, nb;     // nb = active utterances (<= 64)
    long long* trace;        // optional (debug): [n][8] clock64 phase stamps of CTA trace_cta
    int trace_cta, trace_t0, trace_n;
};

// dynamic shared memory layout (floats)
constexpr int AB_MROWS = 24;                        // 23 weight rows padded to three n8 tiles
constexpr int AB_PSTR = 68;                         // utterance stride of a `part` row (conflict-free C-fragment stores)
constexpr int AB_W2S = (AB_FC / 4) * AB_R * 4;      // fc2 groups:    [64][2][4]
constexpr int AB_ES = AB_Q * AB_NROW;
constexpr int AB_PART = AB_CC * AB_MROWS * AB_PSTR;  // [cc][row 0..23][b (stride 68)]
constexpr int AB_PART2 = 4 * AB_R * AB_B;           // fc2 partials [cc3][r][b]
constexpr int AB_HH = AB_NROW * AB_B;
constexpr int AB_HOWN = AB_U * AB_B;
constexpr int AB_GC = AB_NROW * AB_B;
constexpr size_t AB_SMEM = sizeof(float) * (AB_W2S + AB_ES + AB_PART + AB_PART2 + AB_HH + AB_HOWN + AB_GC) + sizeof(int) * AB_B;
constexpr int AB_PLANE = (AB_H / 2) * AB_B;         // 2 * AB_PLANE uint32 words = one parity buffer of h (hi and lo halves of every element)

__device__ __forceinline__ float ld_strong(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// grid barrier, split: signal = publish this CTA's arrival in its own 128-byte slot (after a fence), wait = the sampler
// warp polls all 128 arrivals.  Work placed between the two hides the ~4000-cycle barrier latency.
__device__ __forceinline__ void ab_signal(ll_word* flags, uint32_t tag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ll_store(flags + blockIdx.x * 16, 0.f, tag);
    }
}
__device__ __forceinline__ bool ab_wait(ll_word* flags, uint32_t tag, volatile int* abort_flag, int* status) {
    if (threadIdx.x >= AB_SW * 32) {                  // the sampler warp polls; the tensor-core warps go straight to bar.sync
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
            if (clock64() - t0 > LL_TIMEOUT_CYCLES) {
                *abort_flag = 1;
                if (ln == 0) atomicExch(status, VQCPC_ERR_TIMEOUT);
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
    return *abort_flag == 0;
}

__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// One m16 tile of weight rows x utterance tiles [n_begin, n_end) x this warp's 7 k-tiles (112 columns).
// hp_hi / hp_lo -> plane word of (column pair of the chunk's first column + lane % 4, utterance lane / 4).
// B fragment of (k-tile, n-tile): b0 = pair k*8 + lane%4, b1 = pair k*8 + 4 + lane%4, utterance n*8 + lane/4.
// A fragments (h) of k-tile k for the four utterance tiles: a[m][0..3] = hi a0..a3, a[m][4..7] = lo a0..a3
template <int MT>
__device__ __forceinline__ void ab_load_a(const uint4* hq, int k, uint32_t (&a)[4][8]) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint4* src = hq + ((k * 4 + m) * 32) * 2;
        if (m < MT) {                                   // utterance tiles beyond the batch: no loads, no MMAs
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a[m][0]), "=r"(a[m][1]), "=r"(a[m][2]), "=r"(a[m][3]) : "l"(src) : "memory");
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a[m][4]), "=r"(a[m][5]), "=r"(a[m][6]), "=r"(a[m][7]) : "l"(src + 1) : "memory");
        }
    }
}
__device__ __forceinline__ void mma_bf16_a(float (&d)[4], const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// one k-tile x weight-row tiles [N0, N1): consecutive MMAs go to different accumulators
template <int N0, int N1, int MT>
__device__ __forceinline__ void ab_mma_ktile(const uint32_t (&w_hi)[3][2], const uint32_t (&w_lo)[3][2], const uint32_t (&a)[4][8],
                                             float (&acc)[4][3][4]) {
#pragma unroll
    for (int n = N0; n < N1; ++n)
#pragma unroll
        for (int m = 0; m < 4; ++m) if (m < MT) mma_bf16_a(acc[m][n], &a[m][0], w_hi[n][0], w_hi[n][1]);
#pragma unroll
    for (int n = N0; n < N1; ++n)
#pragma unroll
        for (int m = 0; m < 4; ++m) if (m < MT) mma_bf16_a(acc[m][n], &a[m][4], w_hi[n][0], w_hi[n][1]);
#pragma unroll
    for (int n = N0; n < N1; ++n)
#pragma unroll
        for (int m = 0; m < 4; ++m) if (m < MT) mma_bf16_a(acc[m][n], &a[m][0], w_lo[n][0], w_lo[n][1]);
}
// k-tiles [K0, K1) of this warp's column chunk x all 64 utterances x weight-row tiles [N0, N1): the A fragments of
// k-tile k+1 are in flight while k-tile k runs on the tensor cores.
template <int K0, int K1, int N0, int N1, int MT = 4>
__device__ __forceinline__ void ab_mma_pass(const uint32_t (&w_hi)[7][3][2], const uint32_t (&w_lo)[7][3][2], const uint4* hq,
                                            float (&acc)[4][3][4]) {
    // three rotating fragment buffers: the loads of k-tiles k+1 and k+2 are in flight while k-tile k runs on the
    // tensor cores (an L2 round trip under load is longer than the MMAs of one k-tile)
    uint32_t a[3][4][8];
    ab_load_a<MT>(hq, K0, a[0]);
    if (K0 + 1 < K1) ab_load_a<MT>(hq, K0 + 1, a[1]);
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        if (k + 2 < K1) ab_load_a<MT>(hq, k + 2, a[(k - K0 + 2) % 3]);
        ab_mma_ktile<N0, N1, MT>(w_hi[k], w_lo[k], a[(k - K0) % 3], acc);
    }
}
// C fragment of (utterance tile m, weight-row tile n): c0,c1 = (utt m*16 + lane/4, rows n*8 + 2*(lane%4) + {0,1}),
// c2,c3 = utt + 8
template <int N0, int N1, int MT = 4>
__device__ __forceinline__ void ab_store_c(float* part_cc, int lane, const float (&acc)[4][3][4]) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = N0; n < N1; ++n) {
            float* d = part_cc + (n * 8 + 2 * (lane & 3)) * AB_PSTR + m * 16 + (lane >> 2);
            if (m < MT) {
                d[0] = acc[m][n][0];
                d[AB_PSTR] = acc[m][n][1];
                d[8] = acc[m][n][2];
                d[AB_PSTR + 8] = acc[m][n][3];
            }
        }
}

// inverse-CDF sample of one utterance by one warp: lane l holds the logits of classes 8l .. 8l+7, u = the step's uniform
__device__ __forceinline__ int ab_sample(const float (&ov)[8], float u, int lane) {
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
    const float thr = u * __shfl_sync(0xffffffffu, incl, 31);
    int loc = 8;
#pragma unroll
    for (int k = 7; k >= 0; --k) if (excl + c[k] > thr) loc = k;
    const unsigned hit = __ballot_sync(0xffffffffu, loc < 8);
    int x = AB_Q - 1;
    if (hit != 0u) {
        const int src = __ffs(hit) - 1;
        x = 8 * src + __shfl_sync(0xffffffffu, loc, src);
    }
    return x;
}

template <int MT>
__global__ void __launch_bounds__(AB_THREADS, 1) ar_batch_kernel(AbParams p) {
    extern __shared__ __align__(16) float ab_smem[];
    float* W2s = ab_smem;                // [c4][r][4]
    float* Es = W2s + AB_W2S;            // [x][21]
    float* part = Es + AB_ES;            // [cc][row 0..31][b]  column-chunk partials of the weight rows
    float* part2 = part + AB_PART;       // [cc3][r][b]         fc2 partials
    float* hh = part2 + AB_PART2;        // [row][b]   W_hh h + b_hh
    float* hown = hh + AB_HH;            // [u][b]
    float* Gc = hown + AB_HOWN;          // [row][b]   conditioning of the current frame
    int* xcur = reinterpret_cast<int*>(Gc + AB_GC);   // [b]
    __shared__ volatile int abort_flag;
    __shared__ float bhh_s[AB_NROW], b1_s[AB_R], b2_s[AB_R];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const int cc = warp;                              // column chunk of this warp in the h stream
    const int ug = warp & 1, cc3 = warp >> 1;         // P3: utterance group / 64-column chunk
    const int slot = ug * 32 + lane;                  // P3: this lane's utterance
    const bool teacher = p.x_in != nullptr;
    const int L = p.L, nb = p.nb;
    constexpr int BU = 16 * MT;                       // utterance slots computed: MT = tiles of 16 in use (1..4), the others cost nothing

    // ---- one-time: this warp's slice of the 23 weight rows as bf16 hi/lo mma B-fragments (registers)
    // B fragment of (k-tile k, row tile n): b[0] = (cols k*16 + 2*(lane%4) + {0,1}, row n*8 + lane/4), b[1] = cols + 8.
    // Rows 0..20 = W_hh (3u + g), 21..22 = fc1, 23 = zero.
    uint32_t w_hi[7][3][2], w_lo[7][3][2];
    {
        auto wrow = [&](int r) -> const float* {
            if (r < AB_NROW) return p.w_hh + static_cast<int64_t>((r % 3) * AB_H + cta * AB_U + r / 3) * AB_H;
            if (r < AB_ROWS) return p.fc1_w + static_cast<int64_t>(cta * AB_R + (r - AB_NROW)) * AB_H;
            return nullptr;
        };
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int n = 0; n < 3; ++n)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float* src = wrow(n * 8 + (lane >> 2));
                    const int c = cc * AB_CHUNK + k * 16 + 2 * (lane & 3) + q * 8;
                    float w0 = 0.f, w1 = 0.f;
                    if (src != nullptr) { w0 = __ldg(src + c); w1 = __ldg(src + c + 1); }
                    split_pair(w0, w1, w_hi[k][n][q], w_lo[k][n][q]);
                }
    }
    for (int i = tid; i < (AB_FC / 4) * AB_R; i += AB_THREADS) {
        const int c4 = i / AB_R, r = i % AB_R;
        reinterpret_cast<float4*>(W2s)[i] = __ldg(reinterpret_cast<const float4*>(p.fc2_w + static_cast<int64_t>(cta * AB_R + r) * AB_FC + 4 * c4));
    }
    for (int i = tid; i < AB_ES; i += AB_THREADS) {
        const int x = i / AB_NROW, j = i % AB_NROW;
        Es[i] = __ldg(p.eprime + static_cast<int64_t>(x) * AB_G + (j % 3) * AB_H + cta * AB_U + j / 3);
    }
    if (tid < AB_NROW) bhh_s[tid] = __ldg(p.b_hh + (tid % 3) * AB_H + cta * AB_U + tid / 3);
    if (tid < AB_R) { b1_s[tid] = __ldg(p.fc1_b + cta * AB_R + tid); b2_s[tid] = __ldg(p.fc2_b + cta * AB_R + tid); }
    if (tid == 0) abort_flag = 0;
    __syncthreads();
    for (int i = tid; i < AB_HH; i += AB_THREADS) hh[i] = bhh_s[i / AB_B];     // W_hh h_{-1} + b_hh, h_{-1} = 0
    for (int i = tid; i < AB_HOWN; i += AB_THREADS) hown[i] = 0.f;
    if (tid < AB_B) xcur[tid] = AB_X_INIT;
    __syncthreads();

    uint32_t tag = 0;
    int frame_left = 0, frame = 0;
    const bool tracing = p.trace != nullptr && cta == p.trace_cta && tid == 0;
#define AB_TRACE(k) if (tracing && t >= p.trace_t0 && t < p.trace_t0 + p.trace_n) p.trace[(t - p.trace_t0) * 8 + (k)] = clock64();
    const bool tracing_s = p.trace != nullptr && cta == p.trace_cta && tid == AB_SW * 32;
#define AB_TRACE_S(k) if (tracing_s && t >= p.trace_t0 && t < p.trace_t0 + p.trace_n) p.trace[(t - p.trace_t0) * 8 + (k)] = clock64();
    for (int t = 0; t < L; ++t) {
        AB_TRACE(0)
        uint32_t* hp = p.hP + static_cast<int64_t>(t & 1) * 2 * AB_PLANE;          // h_t fragments, double buffered by parity
        const uint4* hq = reinterpret_cast<const uint4*>(hp) + ((cc * 7 * 4) * 32 + lane) * 2;   // this warp's first (k-tile, utterance tile)
        // ------------------------------------------------------------------ G: conditioning reload, gates, h_t planes
        if (frame_left == 0) {
            for (int i = tid; i < AB_NROW * BU; i += AB_THREADS) {
                const int row = i / BU, b = i % BU;
                Gc[row * AB_B + b] = (b < nb) ? __ldg(p.G + b * p.g_stride + static_cast<int64_t>(frame) * AB_G + (row % 3) * AB_H + cta * AB_U + row / 3) : 0.f;
            }
            frame_left = p.upsample; ++frame;
        }
        --frame_left;
        if (teacher && tid < AB_B) xcur[tid] = (tid < nb) ? (static_cast<int>(__ldg(p.x_in + static_cast<int64_t>(tid) * L + t)) & (AB_Q - 1)) : 0;
        __syncthreads();
        for (int j = tid; j < AB_U * BU; j += AB_THREADS) {
            const int u = j / BU, b = j % BU, i = u * AB_B + b;
            const float* e = &Es[xcur[b] * AB_NROW + 3 * u];
            const float r = sigmoid_fast(__fadd_rn(__fadd_rn(e[0], Gc[(3 * u) * AB_B + b]), hh[(3 * u) * AB_B + b]));
            const float z = sigmoid_fast(__fadd_rn(__fadd_rn(e[1], Gc[(3 * u + 1) * AB_B + b]), hh[(3 * u + 1) * AB_B + b]));
            const float n = tanh_fast(__fmaf_rn(r, hh[(3 * u + 2) * AB_B + b], __fadd_rn(e[2], Gc[(3 * u + 2) * AB_B + b])));
            const float hn = __fmaf_rn(z, __fsub_rn(hown[i], n), n);
            hown[i] = hn;
            // publish h already split for the tensor cores: element (column, utterance) of the hi / lo planes
            const int col = cta * AB_U + u;
            const __nv_bfloat16 hb = __float2bfloat16_rn(hn);
            const __nv_bfloat16 lb = __float2bfloat16_rn(hn - __bfloat162float(hb));
            // fragment-major layout [k-tile 56][utterance tile 4][lane 32][hi a0..a3, lo a0..a3] of bf16x2 words: the
            // consumer lane (utterance % 8, column pair % 4) of tile (col / 16, b / 16) finds its eight A registers in
            // 32 contiguous bytes; register a(rsel + 2 csel) holds (utterance + 8 rsel, columns + 8 csel)
            const int c16 = col & 15, u16 = b & 15;
            const size_t word = ((static_cast<size_t>(col >> 4) * 4 + (b >> 4)) * 32 + ((u16 & 7) * 4 + ((c16 & 7) >> 1))) * 8 +
                                (u16 >> 3) + 2 * (c16 >> 3);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(hp) + word * 2 + (col & 1);
            dst[0] = hb;
            dst[8] = lb;
        }
        AB_TRACE(1)
        ab_signal(p.flags, ++tag);
        if (!ab_wait(p.flags, tag, &abort_flag, p.status)) return;                  // barrier 1: h_t complete
        AB_TRACE(2)

        // ------------------------------------------------------------------ P2a: row tile 2 = W_hh rows 16..20 + the fc1 rows
        // (the only part of the contraction the critical path needs: legacy HMMA issues at ~24 cycles per SMSP on this
        // part, so every MMA moved out of here is ~50 cycles off the step)
        float* part_cc = part + cc * AB_MROWS * AB_PSTR;
        float acc[4][3][4];
        if (warp < AB_SW) {
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 3; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
            ab_mma_pass<0, 7, 2, 3, MT>(w_hi, w_lo, hq, acc);
            ab_store_c<2, 3, MT>(part_cc, lane, acc);
        }
        __syncthreads();
        if (tid < AB_R * AB_B) {
            const int r = tid / AB_B, b = tid % AB_B;
            float sum = 0.f;
            if (b < BU) {
#pragma unroll
                for (int c = 0; c < AB_CC; ++c) sum += part[(c * AB_MROWS + AB_NROW + r) * AB_PSTR + b];
            }
            p.rT[(cta * AB_R + r) * AB_B + b] = b < BU ? fmaxf(sum + b1_s[r], 0.f) : 0.f;
        }
        AB_TRACE(3)
        ab_signal(p.flags, ++tag);                                                  // barrier 2 (r complete) ...
        // ------------------------------------------------------------------ P2b: row tiles 0, 1 (W_hh rows 0..15, needed only
        // by the NEXT step's gates): k-tiles 0..2 while barrier 2 is in flight, the rest after P3 while the sampler works
        if (warp < AB_SW) ab_mma_pass<0, 3, 0, 2, MT>(w_hi, w_lo, hq, acc);
        if (!ab_wait(p.flags, tag, &abort_flag, p.status)) return;                  // ... barrier 2 wait
        AB_TRACE(4)

        // ------------------------------------------------------------------ P3: fc2 rows over relu(fc1 h_t)
        if (warp < AB_SW) {
            float a0 = 0.f, a1 = 0.f;
            const float* rcol = p.rT + static_cast<int64_t>(cc3 * 64) * AB_B + slot;
            const float4* w2g = reinterpret_cast<const float4*>(W2s) + static_cast<int64_t>(cc3 * 16) * AB_R;
            float rv[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) rv[i] = ld_strong(rcol + i * AB_B);       // all 64 loads in flight at once
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
                const float4 w0 = w2g[c4 * AB_R], w1 = w2g[c4 * AB_R + 1];
                a0 = fmaf(w0.x, rv[4 * c4], a0); a0 = fmaf(w0.y, rv[4 * c4 + 1], a0);
                a0 = fmaf(w0.z, rv[4 * c4 + 2], a0); a0 = fmaf(w0.w, rv[4 * c4 + 3], a0);
                a1 = fmaf(w1.x, rv[4 * c4], a1); a1 = fmaf(w1.y, rv[4 * c4 + 1], a1);
                a1 = fmaf(w1.z, rv[4 * c4 + 2], a1); a1 = fmaf(w1.w, rv[4 * c4 + 3], a1);
            }
            part2[(cc3 * AB_R + 0) * AB_B + slot] = a0;
            part2[(cc3 * AB_R + 1) * AB_B + slot] = a1;
        }
        __syncthreads();
        if (tid < AB_R * AB_B) {
            const int r = tid / AB_B, b = tid % AB_B;
            const float o = (part2[(0 * AB_R + r) * AB_B + b] + part2[(1 * AB_R + r) * AB_B + b]) +
                            (part2[(2 * AB_R + r) * AB_B + b] + part2[(3 * AB_R + r) * AB_B + b]) + b2_s[r];
            // logits travel as LL words (value + step tag in one 8-byte store): no fence, no barrier
            if (!teacher && b < nb) ll_store(p.oLL + b * AB_Q + cta * AB_R + r, o, static_cast<uint32_t>(t + 1));
            if (p.out_logits != nullptr && b < nb)
                p.out_logits[(static_cast<int64_t>(b) * L + t) * AB_Q + cta * AB_R + r] = o;
        }
        AB_TRACE(5)
        if (warp < AB_SW) {
            ab_mma_pass<3, 7, 0, 2, MT>(w_hi, w_lo, hq, acc);
            ab_store_c<0, 2, MT>(part_cc, lane, acc);                                   // the eight column-chunk partials meet in SMEM
        } else if (!teacher && cta < nb) {
            // -------------------------------------------------------------- P4 (sampler warp): CTA b samples utterance b
            const int b = cta;
            const ll_word* src = p.oLL + b * AB_Q + 8 * lane;
            float ov[8];
            {
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
                    if (clock64() - t0 > LL_TIMEOUT_CYCLES) {
                        abort_flag = 1;
                        if (lane == 0) atomicExch(p.status, VQCPC_ERR_TIMEOUT);
                        break;
                    }
                }
            }
            AB_TRACE_S(6)
            const int x = ab_sample(ov, __ldg(p.uniforms + static_cast<int64_t>(b) * L + t), lane);
            if (lane == 0) {
                ll_store(p.xLL + b, __int_as_float(x), static_cast<uint32_t>(t + 1));   // the code travels as an LL word too
                if (p.out_wav) p.out_wav[static_cast<int64_t>(b) * L + t] = __ldg(p.lut + x);
                if (p.out_codes) p.out_codes[static_cast<int64_t>(b) * L + t] = x;
            }
            AB_TRACE_S(7)
        }
        if (!teacher && tid < AB_B) {
            // every CTA picks up the 64 sampled codes (threads of warps 0 and 1, after their W_hh slice)
            int xv = 0;
            if (tid < nb) {
                const long long t0 = clock64();
                for (;;) {
                    const ll_word w = ll_load(p.xLL + tid);
                    if (ll_tag(w) == static_cast<uint32_t>(t + 1)) { xv = __float_as_int(ll_val(w)); break; }
                    if (clock64() - t0 > LL_TIMEOUT_CYCLES) { abort_flag = 1; atomicExch(p.status, VQCPC_ERR_TIMEOUT); break; }
                }
            }
            xcur[tid] = xv;
        }
        __syncthreads();
        if (abort_flag != 0) return;
        for (int j = tid; j < AB_NROW * BU; j += AB_THREADS) {
            const int row = j / BU, b = j % BU;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < AB_CC; ++c) sum += part[(c * AB_MROWS + row) * AB_PSTR + b];
            hh[row * AB_B + b] = sum + bhh_s[row];
        }
        // (the __syncthreads at the top of the next step orders hh / xcur before the gates)
    }
#undef AB_TRACE
#undef AB_TRACE_S
}

// ------------------------------------------------------------------------------------------------------------------
// Two-group variant for 65..128 utterances per launch.  The single-group kernel above spends more than half of a step
// waiting on its four communication hops; here two groups of up to 64 utterances share the CTA's register-resident
// weight fragments and run the same phases interleaved, so one group's hop hides behind the other group's compute:
//
//   [codes A] G(A) s1A | P2b(B, t-1) [codes B] G(B) s1B | w1A P2a(A) s2A | w1B P2a(B) s2B | w2A P3(A) | P2b(A) | w2B P3(B)
//
// (s = barrier signal, w = barrier wait; P2b of group B is software-pipelined into the next step).  320 threads: eight
// tensor-core warps, a barrier-poller warp and a sampler warp.  The sampler warp talks to the rest only through the LL
// words in global memory (logits in, codes out), so it runs fully decoupled; the other nine warps meet at named
// barrier 1.  Same arithmetic per utterance as the single-group kernel (bit-identical results).
// ------------------------------------------------------------------------------------------------------------------
constexpr int AB2_THREADS = 320, AB2_MAIN = 288, AB2_SAMPLER = 9;
struct Ab2Group {
    uint32_t* hP; float* rT; ll_word* oLL; ll_word* xLL; ll_word* flags;
    int nb, b_off, cta_off;      // active utterances, first utterance of the launch, first sampling CTA
};
struct Ab2Params {
    const float* w_hh; const float* b_hh; const float* fc1_w; const float* fc1_b; const float* fc2_w; const float* fc2_b;
    const float* eprime; const float* lut;
    const float* G; const float* uniforms; const int64_t* x_in;
    float* out_wav; int32_t* out_codes; float* out_logits;
    Ab2Group grp[2];
    int* status;
    long long g_stride;
    int L, upsample;
    long long* trace; int trace_cta, trace_t0, trace_n;
};
constexpr int AB2_GRP = AB_HH + AB_HOWN + AB_GC + AB_B;          // per-group shared state: hh, hown, Gc, xcur
constexpr size_t AB2_SMEM = sizeof(float) * (AB_W2S + AB_ES + AB_PART + AB_PART2 + 2 * AB2_GRP);

__device__ __forceinline__ void ab2_signal(ll_word* flags, uint32_t tag) {
    bar_sync(1, AB2_MAIN);
    if (threadIdx.x == 0) {
        __threadfence();
        ll_store(flags + blockIdx.x * 16, 0.f, tag);
    }
}
__device__ __forceinline__ bool ab2_wait(ll_word* flags, uint32_t tag, volatile int* abort_flag, int* status) {
    if ((threadIdx.x >> 5) == AB_SW) {                 // the poller warp
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
    bar_sync(1, AB2_MAIN);
    return *abort_flag == 0;
}

__global__ void __launch_bounds__(AB2_THREADS, 1) ar_batch2_kernel(Ab2Params p) {
    extern __shared__ __align__(16) float ab_smem[];
    float* W2s = ab_smem;                // [c4][r][4]
    float* Es = W2s + AB_W2S;            // [x][21]
    float* part = Es + AB_ES;            // [cc][row 0..23][b]
    float* part2 = part + AB_PART;       // [cc3][r][b]
    float* gstate = part2 + AB_PART2;    // per group: hh [row][b], hown [u][b], Gc [row][b], xcur [b]
    __shared__ volatile int abort_flag;
    __shared__ float bhh_s[AB_NROW], b1_s[AB_R], b2_s[AB_R];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    const int cc = warp;
    const int ug = warp & 1, cc3 = warp >> 1;
    const int slot = ug * 32 + lane;
    const bool teacher = p.x_in != nullptr;
    const int L = p.L;
    const int ngroups = p.grp[1].nb > 0 ? 2 : 1;

    if (tid == 0) abort_flag = 0;
    __syncthreads();

    if (warp == AB2_SAMPLER) {
        // ------------------------------------------------------------------ sampler warp: decoupled from the other nine
        if (teacher) return;
        bool mine[2]; int bb[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) { bb[g] = cta - p.grp[g].cta_off; mine[g] = bb[g] >= 0 && bb[g] < p.grp[g].nb; }
        if (!mine[0] && !mine[1]) return;
        for (int t = 0; t < L; ++t) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (!mine[g]) continue;
                const int b = bb[g];
                const int64_t gb = p.grp[g].b_off + b;
                const ll_word* src = p.grp[g].oLL + b * AB_Q + 8 * lane;
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
                const int x = ab_sample(ov, __ldg(p.uniforms + gb * L + t), lane);
                if (lane == 0) {
                    ll_store(p.grp[g].xLL + b, __int_as_float(x), static_cast<uint32_t>(t + 1));
                    if (p.out_wav) p.out_wav[gb * L + t] = __ldg(p.lut + x);
                    if (p.out_codes) p.out_codes[gb * L + t] = x;
                }
            }
        }
        return;
    }

    // ---- one-time: weight fragments (same layout as the single-group kernel), fc2 rows, E' columns, biases
    uint32_t w_hi[7][3][2], w_lo[7][3][2];
    if (warp < AB_SW) {
        auto wrow = [&](int r) -> const float* {
            if (r < AB_NROW) return p.w_hh + static_cast<int64_t>((r % 3) * AB_H + cta * AB_U + r / 3) * AB_H;
            if (r < AB_ROWS) return p.fc1_w + static_cast<int64_t>(cta * AB_R + (r - AB_NROW)) * AB_H;
            return nullptr;
        };
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int n = 0; n < 3; ++n)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float* src = wrow(n * 8 + (lane >> 2));
                    const int c = cc * AB_CHUNK + k * 16 + 2 * (lane & 3) + q * 8;
                    float w0 = 0.f, w1 = 0.f;
                    if (src != nullptr) { w0 = __ldg(src + c); w1 = __ldg(src + c + 1); }
                    split_pair(w0, w1, w_hi[k][n][q], w_lo[k][n][q]);
                }
    }
    for (int i = tid; i < (AB_FC / 4) * AB_R; i += AB2_MAIN) {
        const int c4 = i / AB_R, r = i % AB_R;
        reinterpret_cast<float4*>(W2s)[i] = __ldg(reinterpret_cast<const float4*>(p.fc2_w + static_cast<int64_t>(cta * AB_R + r) * AB_FC + 4 * c4));
    }
    for (int i = tid; i < AB_ES; i += AB2_MAIN) {
        const int x = i / AB_NROW, j = i % AB_NROW;
        Es[i] = __ldg(p.eprime + static_cast<int64_t>(x) * AB_G + (j % 3) * AB_H + cta * AB_U + j / 3);
    }
    if (tid < AB_NROW) bhh_s[tid] = __ldg(p.b_hh + (tid % 3) * AB_H + cta * AB_U + tid / 3);
    if (tid < AB_R) { b1_s[tid] = __ldg(p.fc1_b + cta * AB_R + tid); b2_s[tid] = __ldg(p.fc2_b + cta * AB_R + tid); }
    bar_sync(1, AB2_MAIN);
    for (int g = 0; g < 2; ++g) {
        float* hh = gstate + g * AB2_GRP;
        float* hown = hh + AB_HH;
        int* xcur = reinterpret_cast<int*>(hown + AB_HOWN + AB_GC);
        for (int i = tid; i < AB_HH; i += AB2_MAIN) hh[i] = bhh_s[i / AB_B];     // W_hh h_{-1} + b_hh, h_{-1} = 0
        for (int i = tid; i < AB_HOWN; i += AB2_MAIN) hown[i] = 0.f;
        if (tid < AB_B) xcur[tid] = AB_X_INIT;
    }
    bar_sync(1, AB2_MAIN);

    uint32_t tag[2] = {0, 0};
    float* part_cc = part + cc * AB_MROWS * AB_PSTR;
    bool dead = false;

    // ---- phases (g is a literal at every call site)
    // codes of step t-1 -> xcur, conditioning reload, gates of step t, h_t published, barrier 1 signalled
    auto begin_step = [&](const int g, const int t) {
        const Ab2Group& q = p.grp[g];
        float* hh = gstate + g * AB2_GRP;
        float* hown = hh + AB_HH;
        float* Gc = hown + AB_HOWN;
        int* xcur = reinterpret_cast<int*>(Gc + AB_GC);
        if (tid < AB_B) {
            if (teacher) {
                xcur[tid] = (tid < q.nb) ? (static_cast<int>(__ldg(p.x_in + static_cast<int64_t>(q.b_off + tid) * L + t)) & (AB_Q - 1)) : 0;
            } else if (t > 0) {
                int xv = 0;
                if (tid < q.nb) {
                    const long long t0 = clock64();
                    for (;;) {
                        const ll_word w = ll_load(q.xLL + tid);
                        if (ll_tag(w) == static_cast<uint32_t>(t)) { xv = __float_as_int(ll_val(w)); break; }
                        if (abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) { abort_flag = 1; atomicExch(p.status, VQCPC_ERR_TIMEOUT); break; }
                    }
                }
                xcur[tid] = xv;
            }
        }
        if (t % p.upsample == 0) {
            const int frame = t / p.upsample;
            for (int i = tid; i < AB_GC; i += AB2_MAIN) {
                const int row = i / AB_B, b = i % AB_B;
                Gc[i] = (b < q.nb) ? __ldg(p.G + (q.b_off + b) * p.g_stride + static_cast<int64_t>(frame) * AB_G + (row % 3) * AB_H + cta * AB_U + row / 3) : 0.f;
            }
        }
        bar_sync(1, AB2_MAIN);
        uint32_t* hp = q.hP + static_cast<int64_t>(t & 1) * 2 * AB_PLANE;
        for (int i = tid; i < AB_U * AB_B; i += AB2_MAIN) {
            const int u = i / AB_B, b = i % AB_B;
            const float* e = &Es[xcur[b] * AB_NROW + 3 * u];
            const float r = sigmoid_fast(__fadd_rn(__fadd_rn(e[0], Gc[(3 * u) * AB_B + b]), hh[(3 * u) * AB_B + b]));
            const float z = sigmoid_fast(__fadd_rn(__fadd_rn(e[1], Gc[(3 * u + 1) * AB_B + b]), hh[(3 * u + 1) * AB_B + b]));
            const float n = tanh_fast(__fmaf_rn(r, hh[(3 * u + 2) * AB_B + b], __fadd_rn(e[2], Gc[(3 * u + 2) * AB_B + b])));
            const float hn = __fmaf_rn(z, __fsub_rn(hown[i], n), n);
            hown[i] = hn;
            const int col = cta * AB_U + u;
            const __nv_bfloat16 hb = __float2bfloat16_rn(hn);
            const __nv_bfloat16 lb = __float2bfloat16_rn(hn - __bfloat162float(hb));
            const int c16 = col & 15, u16 = b & 15;
            const size_t word = ((static_cast<size_t>(col >> 4) * 4 + (b >> 4)) * 32 + ((u16 & 7) * 4 + ((c16 & 7) >> 1))) * 8 +
                                (u16 >> 3) + 2 * (c16 >> 3);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(hp) + word * 2 + (col & 1);
            dst[0] = hb;
            dst[8] = lb;
        }
        ab2_signal(q.flags, ++tag[g]);                                              // barrier 1: h_t complete
    };
    // row tile 2 over all of h_t -> r_t (global) and W_hh rows 16..20 for the next gates
    auto phase_p2a = [&](const int g, const int t) {
        const Ab2Group& q = p.grp[g];
        float* hh = gstate + g * AB2_GRP;
        if (warp < AB_SW) {
            const uint4* hq = reinterpret_cast<const uint4*>(q.hP + static_cast<int64_t>(t & 1) * 2 * AB_PLANE) + ((cc * 7 * 4) * 32 + lane) * 2;
            float acc[4][3][4];
#pragma unroll
            for (int m = 0; m < 4; ++m) { acc[m][2][0] = acc[m][2][1] = acc[m][2][2] = acc[m][2][3] = 0.f; }
            ab_mma_pass<0, 7, 2, 3>(w_hi, w_lo, hq, acc);
            ab_store_c<2, 3>(part_cc, lane, acc);
        }
        bar_sync(1, AB2_MAIN);
        for (int i = tid; i < (AB_ROWS - 16) * AB_B; i += AB2_MAIN) {
            const int row = 16 + i / AB_B, b = i % AB_B;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < AB_CC; ++c) sum += part[(c * AB_MROWS + row) * AB_PSTR + b];
            if (row >= AB_NROW) q.rT[(cta * AB_R + (row - AB_NROW)) * AB_B + b] = fmaxf(sum + b1_s[row - AB_NROW], 0.f);
            else hh[row * AB_B + b] = sum + bhh_s[row];
        }
        ab2_signal(q.flags, ++tag[g]);                                              // barrier 2: r_t complete
    };
    // row tiles 0, 1 over all of h_t -> W_hh rows 0..15 for the next gates
    auto phase_p2b = [&](const int g, const int t) {
        const Ab2Group& q = p.grp[g];
        float* hh = gstate + g * AB2_GRP;
        if (warp < AB_SW) {
            const uint4* hq = reinterpret_cast<const uint4*>(q.hP + static_cast<int64_t>(t & 1) * 2 * AB_PLANE) + ((cc * 7 * 4) * 32 + lane) * 2;
            float acc[4][3][4];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
            ab_mma_pass<0, 7, 0, 2>(w_hi, w_lo, hq, acc);
            ab_store_c<0, 2>(part_cc, lane, acc);
        }
        bar_sync(1, AB2_MAIN);
        for (int i = tid; i < 16 * AB_B; i += AB2_MAIN) {
            const int row = i / AB_B, b = i % AB_B;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < AB_CC; ++c) sum += part[(c * AB_MROWS + row) * AB_PSTR + b];
            hh[row * AB_B + b] = sum + bhh_s[row];
        }
        bar_sync(1, AB2_MAIN);
    };
    // fc2 rows over relu(fc1 h_t) -> logits as LL words
    auto phase_p3 = [&](const int g, const int t) {
        const Ab2Group& q = p.grp[g];
        if (warp < AB_SW) {
            float a0 = 0.f, a1 = 0.f;
            const float* rcol = q.rT + static_cast<int64_t>(cc3 * 64) * AB_B + slot;
            const float4* w2g = reinterpret_cast<const float4*>(W2s) + static_cast<int64_t>(cc3 * 16) * AB_R;
            float rv[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) rv[i] = ld_strong(rcol + i * AB_B);
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
                const float4 w0 = w2g[c4 * AB_R], w1 = w2g[c4 * AB_R + 1];
                a0 = fmaf(w0.x, rv[4 * c4], a0); a0 = fmaf(w0.y, rv[4 * c4 + 1], a0);
                a0 = fmaf(w0.z, rv[4 * c4 + 2], a0); a0 = fmaf(w0.w, rv[4 * c4 + 3], a0);
                a1 = fmaf(w1.x, rv[4 * c4], a1); a1 = fmaf(w1.y, rv[4 * c4 + 1], a1);
                a1 = fmaf(w1.z, rv[4 * c4 + 2], a1); a1 = fmaf(w1.w, rv[4 * c4 + 3], a1);
            }
            part2[(cc3 * AB_R + 0) * AB_B + slot] = a0;
            part2[(cc3 * AB_R + 1) * AB_B + slot] = a1;
        }
        bar_sync(1, AB2_MAIN);
        if (tid < AB_R * AB_B) {
            const int r = tid / AB_B, b = tid % AB_B;
            const float o = (part2[(0 * AB_R + r) * AB_B + b] + part2[(1 * AB_R + r) * AB_B + b]) +
                            (part2[(2 * AB_R + r) * AB_B + b] + part2[(3 * AB_R + r) * AB_B + b]) + b2_s[r];
            if (!teacher && b < q.nb) ll_store(q.oLL + b * AB_Q + cta * AB_R + r, o, static_cast<uint32_t>(t + 1));
            if (p.out_logits != nullptr && b < q.nb)
                p.out_logits[(static_cast<int64_t>(q.b_off + b) * L + t) * AB_Q + cta * AB_R + r] = o;
        }
        bar_sync(1, AB2_MAIN);
    };

    const bool tracing = p.trace != nullptr && cta == p.trace_cta && tid == 0;
#define AB_TRACE(k) if (tracing && t >= p.trace_t0 && t < p.trace_t0 + p.trace_n) p.trace[(t - p.trace_t0) * 8 + (k)] = clock64();
    for (int t = 0; t < L && !dead; ++t) {
        AB_TRACE(0)
        begin_step(0, t);
        AB_TRACE(1)
        if (ngroups == 2) {
            if (t > 0) phase_p2b(1, t - 1);
            begin_step(1, t);
        }
        AB_TRACE(2)
        if (!ab2_wait(p.grp[0].flags, tag[0], &abort_flag, p.status)) { dead = true; break; }
        AB_TRACE(3)
        phase_p2a(0, t);
        AB_TRACE(4)
        if (ngroups == 2) {
            if (!ab2_wait(p.grp[1].flags, tag[1], &abort_flag, p.status)) { dead = true; break; }
            phase_p2a(1, t);
        }
        AB_TRACE(5)
        if (!ab2_wait(p.grp[0].flags, tag[0], &abort_flag, p.status)) { dead = true; break; }
        phase_p3(0, t);
        AB_TRACE(6)
        phase_p2b(0, t);
        AB_TRACE(7)
        if (ngroups == 2) {
            if (!ab2_wait(p.grp[1].flags, tag[1], &abort_flag, p.status)) { dead = true; break; }
            phase_p3(1, t);
        }
    }
#undef AB_TRACE
}

// workspace of one launch / group: [h 2 parity x 896x64 words][rT 256x64][oLL 64x256 LL words][xLL 64 LL words][flags 128x16 words]
static size_t ab_ws_bytes() {
    return sizeof(float) * (2 * AB_H + AB_FC) * AB_B + sizeof(ll_word) * (AB_B * AB_Q + AB_B + AB_CTAS * 16);
}
size_t ar_batch_workspace_bytes() {
    return 2 * align_up(ab_ws_bytes(), 256);
}

long long* g_ab_trace = nullptr;
int g_ab_trace_cta = 0, g_ab_trace_t0 = 0, g_ab_trace_n = 0;
int g_ab_two_group = 1;      // debug switch (vqcpc_debug_set_ar_poll_gap bit 29 clears it)

int ar_batch_run(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int B, int T2,
                 int L, void* ws, int* status, float* out_wav, int32_t* out_codes, float* out_logits, cudaStream_t stream) {
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(ar_batch_kernel<1>), static_cast<int>(AB_SMEM))) return rc_attr;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(ar_batch_kernel<2>), static_cast<int>(AB_SMEM))) return rc_attr;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(ar_batch_kernel<4>), static_cast<int>(AB_SMEM))) return rc_attr;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(ar_batch2_kernel), static_cast<int>(AB2_SMEM))) return rc_attr;
    unsigned char* base = static_cast<unsigned char*>(ws);
    const size_t gbytes = align_up(ab_ws_bytes(), 256);
    auto carve = [&](unsigned char* gb, uint32_t*& hP, float*& rT, ll_word*& oLL, ll_word*& xLL, ll_word*& flags) {
        hP = reinterpret_cast<uint32_t*>(gb);
        rT = reinterpret_cast<float*>(gb) + 2 * AB_H * AB_B;
        oLL = reinterpret_cast<ll_word*>(rT + AB_FC * AB_B);
        xLL = oLL + AB_B * AB_Q;
        flags = xLL + AB_B;
    };
    for (int b0 = 0; b0 < B;) {
        const int left = B - b0;
        const bool two = g_ab_two_group && left > AB_B;          // 65..128 utterances: two interleaved groups
        const int nb = two ? (left < 2 * AB_B ? left : 2 * AB_B) : (left < AB_B ? left : AB_B);
        const float* Gp = G + static_cast<int64_t>(b0) * T2 * AB_G;
        const float* up = uniforms ? uniforms + static_cast<int64_t>(b0) * L : nullptr;
        const int64_t* xp = x_in ? x_in + static_cast<int64_t>(b0) * L : nullptr;
        float* ow = out_wav ? out_wav + static_cast<int64_t>(b0) * L : nullptr;
        int32_t* oc = out_codes ? out_codes + static_cast<int64_t>(b0) * L : nullptr;
        float* ol = out_logits ? out_logits + static_cast<int64_t>(b0) * L * AB_Q : nullptr;
        VQ_CUDA(cudaMemsetAsync(ws, 0, 2 * align_up(ab_ws_bytes(), 256), stream));
        if (!two) {
            AbParams p{};
            p.w_hh = w->ar_w_hh; p.b_hh = w->ar_b_hh; p.fc1_w = w->fc1_w; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b;
            p.eprime = w->eprime; p.lut = w->mulaw_lut;
            p.G = Gp; p.g_stride = static_cast<long long>(T2) * AB_G;
            p.uniforms = up; p.x_in = xp; p.out_wav = ow; p.out_codes = oc; p.out_logits = ol;
            carve(base, p.hP, p.rT, p.oLL, p.xLL, p.flags);
            p.status = status;
            p.L = L; p.upsample = w->upsample_t; p.nb = nb;
            p.trace = g_ab_trace; p.trace_cta = g_ab_trace_cta; p.trace_t0 = g_ab_trace_t0; p.trace_n = g_ab_trace_n;
            void* args[] = {&p};
            // only the utterance tiles of 16 that are in use are loaded and multiplied
            // (three tiles measured no faster than four: 15.5 vs 14.4 us/step, so 33..64 utterances run the 4-tile kernel)
            void* fn = nb <= 16 ? reinterpret_cast<void*>(ar_batch_kernel<1>) : nb <= 32 ? reinterpret_cast<void*>(ar_batch_kernel<2>)
                                                                              : reinterpret_cast<void*>(ar_batch_kernel<4>);
            VQ_CUDA(cudaLaunchCooperativeKernel(fn, dim3(AB_CTAS), dim3(AB_THREADS), args, AB_SMEM, stream));
        } else {
            Ab2Params p{};
            p.w_hh = w->ar_w_hh; p.b_hh = w->ar_b_hh; p.fc1_w = w->fc1_w; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b;
            p.eprime = w->eprime; p.lut = w->mulaw_lut;
            p.G = Gp; p.g_stride = static_cast<long long>(T2) * AB_G;
            p.uniforms = up; p.x_in = xp; p.out_wav = ow; p.out_codes = oc; p.out_logits = ol;
            const int nA = (nb + 1) / 2;
            for (int g = 0; g < 2; ++g) {
                Ab2Group& q = p.grp[g];
                carve(base + g * gbytes, q.hP, q.rT, q.oLL, q.xLL, q.flags);
                q.nb = g == 0 ? nA : nb - nA;
                q.b_off = g == 0 ? 0 : nA;
                q.cta_off = g == 0 ? 0 : AB_B;
            }
            p.status = status;
            p.L = L; p.upsample = w->upsample_t;
            p.trace = g_ab_trace; p.trace_cta = g_ab_trace_cta; p.trace_t0 = g_ab_trace_t0; p.trace_n = g_ab_trace_n;
            void* args[] = {&p};
            VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ar_batch2_kernel), dim3(AB_CTAS), dim3(AB2_THREADS), args,
                                                AB2_SMEM, stream));
        }
        count_launch(1);
        b0 += nb;
    }
    return VQCPC_OK;
}

}  // namespace vqcpc
