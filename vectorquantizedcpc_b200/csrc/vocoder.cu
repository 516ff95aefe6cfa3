// Vocoder.generate / Vocoder.forward of /root/reference/network_vocoder.py:41-78 on sm_100a.
//   conditioning : code/speaker embedding + x2 nearest + concat (network_vocoder.py:73-77) -> 2-layer biGRU
//                  prenet (hidden 128/dir) -> hoisted input projection G (the x160 upsample is t -> t/160)
//   sample loop  : ONE persistent cooperative kernel per utterance over 128 SMs; GRUCell(896) / fc1 / fc2
//                  weights live in registers, three LL exchanges per step (h, relu(fc1), logits).
// The rnnms arithmetic is restated per SURVEY.md App. A.3 (dims from /root/reference/config.py:62-77,199).
#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

constexpr int PRE_H = 128;           // prenet hidden per direction (dim_voc_latent / 2)
constexpr int PRE_G = 3 * PRE_H;     // 384 gate rows per direction
constexpr int AR_H = 896;            // size_h_rnn
constexpr int AR_G = 3 * AR_H;       // 2688
constexpr int AR_FC = 256;           // size_h_fc
constexpr int AR_Q = 256;            // 2**bits_mu_law classes
constexpr int AR_EMB = 256;          // size_i_embed_ar
constexpr int AR_COND = 256;         // dim_voc_latent
constexpr int AR_CTAS = 128;         // persistent grid: one CTA per SM on 128 SMs
constexpr int AR_U = AR_H / AR_CTAS; // 7 hidden units per CTA
constexpr int AR_R = AR_FC / AR_CTAS;  // 2 fc1 rows and 2 fc2 rows per CTA
constexpr int X_INIT = 128;

// ------------------------------------------------------------------------------------------------
// u[b,t,:] = [code_emb[codes[b, t>>1]] ; spk_emb[speaker[b]]]      (B, 2Tc, 128)
// ------------------------------------------------------------------------------------------------
__global__ void embed_concat_kernel(const float* __restrict__ code_emb, const float* __restrict__ spk_emb,
                                    const int64_t* __restrict__ codes, const int64_t* __restrict__ speaker,
                                    float* __restrict__ u, int B, int Tc, int dc, int ds, int n_codes, int n_speakers,
                                    int* __restrict__ index_error) {
    const int width = dc + ds;
    const int64_t total = static_cast<int64_t>(B) * 2 * Tc * (width / 4);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(i % (width / 4));
        const int64_t bt = i / (width / 4);
        const int t = static_cast<int>(bt % (2 * Tc));
        const int b = static_cast<int>(bt / (2 * Tc));
        float4 v;
        // out-of-range ids (nn.Embedding raises): clamped for memory safety and reported through the workspace header, so the
        // host needs no synchronising range check before the launch
        if (q * 4 < dc) {
            int64_t c = codes[static_cast<int64_t>(b) * Tc + (t >> 1)];
            if (c < 0 || c >= n_codes) { *index_error = INDEX_ERROR_MAGIC; c = 0; }
            v = __ldg(reinterpret_cast<const float4*>(code_emb + c * dc) + q);
        } else {
            int64_t sp = speaker[b];
            if (sp < 0 || sp >= n_speakers) { *index_error = INDEX_ERROR_MAGIC; sp = 0; }
            v = __ldg(reinterpret_cast<const float4*>(spk_emb + sp * ds) + (q - dc / 4));
        }
        reinterpret_cast<float4*>(u)[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// One direction of one biGRU layer for one utterance per CTA.  384 threads: thread r owns gate row r of
// W_hh (128 weights in registers); h lives in shared memory and is read as broadcast LDS.128.
// xproj (B, T, 768) = W_ih x + b_ih for [fwd | bwd]; out (B, T, 256) = [h_fwd | h_bwd].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PRE_G, 1)
bigru_layer_kernel(const float* __restrict__ xproj, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                   float* __restrict__ out, int T, const int32_t* __restrict__ code_lengths) {
    __shared__ __align__(16) float h[PRE_H];
    __shared__ float pre[PRE_G];
    __shared__ float an[PRE_H];
    const int tid = threadIdx.x, dir = blockIdx.x, b = blockIdx.y;
    float w[PRE_H];
    {
        const float4* row = reinterpret_cast<const float4*>(w_hh + (static_cast<int64_t>(dir) * PRE_G + tid) * PRE_H);
#pragma unroll
        for (int k = 0; k < PRE_H / 4; ++k) {
            const float4 v = __ldg(row + k);
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
    }
    const float bias = __ldg(b_hh + dir * PRE_G + tid);
    if (tid < PRE_H) h[tid] = 0.f;
    __syncthreads();
    const float* xp = xproj + static_cast<int64_t>(b) * T * (2 * PRE_G) + dir * PRE_G + tid;
    float* o = out + static_cast<int64_t>(b) * T * (2 * PRE_H) + dir * PRE_H + tid;
    // ragged batches: utterance b is Tb = 2 * code_lengths[b] frames long; its backward direction starts at ITS last
    // frame (a padded tail must not leak into the valid frames); the tail of the output is zero-filled
    const int Tb = code_lengths ? min(T, 2 * __ldg(code_lengths + b)) : T;
    if (tid < PRE_H)
        for (int t = Tb; t < T; ++t) o[static_cast<int64_t>(t) * (2 * PRE_H)] = 0.f;
    for (int s = 0; s < Tb; ++s) {
        const int t = dir ? (Tb - 1 - s) : s;
        const float a = __ldg(xp + static_cast<int64_t>(t) * (2 * PRE_G));
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
        for (int k = 0; k < PRE_H / 4; ++k) {
            const float4 hv = *reinterpret_cast<const float4*>(&h[4 * k]);
            acc0 = fmaf(w[4 * k], hv.x, acc0);
            acc1 = fmaf(w[4 * k + 1], hv.y, acc1);
            acc2 = fmaf(w[4 * k + 2], hv.z, acc2);
            acc3 = fmaf(w[4 * k + 3], hv.w, acc3);
        }
        const float hh = (acc0 + acc1) + (acc2 + acc3) + bias;
        if (tid < 2 * PRE_H) pre[tid] = a + hh;          // r and z rows: a + (W_h. h + b_h.)
        else { pre[tid] = hh; an[tid - 2 * PRE_H] = a; }  // n row: keep W_hn h + b_hn apart (gated by r)
        __syncthreads();
        if (tid < PRE_H) {
            const float r = sigmoid_fast(pre[tid]);
            const float z = sigmoid_fast(pre[PRE_H + tid]);
            const float n = tanh_fast(an[tid] + r * pre[2 * PRE_H + tid]);
            const float hn = (1.0f - z) * n + z * h[tid];
            h[tid] = hn;
            o[static_cast<int64_t>(t) * (2 * PRE_H)] = hn;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// The autoregressive sample loop (rnnms AR part restated; hoisted form):
//   a_t   = E'[x_{t-1}] + G[t/160]                      E' = emb . W_ih[:, :256]^T,  G includes b_ih
//   b_t   = W_hh h_{t-1} + b_hh
//   r = s(a_r + b_r) ; z = s(a_z + b_z) ; n = tanh(a_n + r b_n) ; h_t = (1-z) n + z h_{t-1}
//   o_t   = fc2(relu(fc1 h_t))  ;  x_t = min{k : cumsum(exp(o - max))_k > u_t * S}
// Grid: 128 CTAs x 224 threads, cooperative (all co-resident), ONE launch per group of <= NB utterances.
// CTA j owns hidden units 7j..7j+6 (their 21 W_hh rows in the registers of warps 3..6, a column quarter each,
// 7 columns per lane), fc1 rows 2j,2j+1, fc2 rows 2j,2j+1 and the 21 matching columns of E' (21.5 KB smem).
// A step is three all-to-all LL exchanges (h_t, relu(fc1 h_t), logits).  What B200 measures for them
// (profiles/r01_ll_microbench_*.txt): reading a line another SM has just written costs ~750 cycles one way;
// a 128-way exchange costs ~2050 cycles when every producer owns its own 64/128-byte slot and 3600-6400 when
// producers share lines; replicating the buffers does not help; > 32 outstanding strong stores per publish
// serialise at ~21 cycles each; weak (.cg) loads never observe remote stores.  Hence:
//   * every producer owns a padded slot (h: 64 B, r / o: one 128 B line), one copy, <= 8 stores per publish;
//   * three ROLE warps, each polling one buffer and publishing the next one, chained by shared-memory flags:
//       warp 0 "H": poll h_t -> hand h_t to the W_hh warps (smem, named barrier) -> fc1 rows -> publish r_t
//       warp 1 "R": poll r_t -> fc2 rows -> publish o_t          (teacher-forced mode: write logits instead)
//       warp 2 "O": poll o_{t-1} -> softmax + inverse-CDF sample x_{t-1} (every CTA redundantly, bit-identical)
//                   -> GRU gates of step t (needs W_hh h_{t-1} from warps 3..6) -> publish h_t
//   * the per-utterance latency is exchange-bound, so a launch interleaves NB utterances: every role walks
//     the utterances round-robin and works on one while the others' exchanges are in flight.
// ------------------------------------------------------------------------------------------------
constexpr int AR_ROLE_WARPS = 3;
constexpr int AR_HH_WARPS = 4;                            // each takes 224 of the 896 columns of the CTA's 21 W_hh rows
constexpr int AR_THREADS = 32 * (AR_ROLE_WARPS + AR_HH_WARPS);   // 224 (register file is allotted per 128 threads:
                                                                  // 256 threads -> up to 255 registers each)
constexpr int AR_HH_THREADS = 32 * AR_HH_WARPS;           // 128
constexpr int AR_BAR_COUNT = AR_HH_THREADS + 32;          // W_hh warps + one role warp
constexpr int AR_NROW = AR_U * 3;                         // 21 W_hh rows per CTA
constexpr int AR_HSLOT = 8;                               // words per producer in the h exchange (7 + 1 pad = 64 B)
constexpr int AR_RSLOT = 16;                              // words per producer in the r / o exchanges (own 128 B line)
constexpr int AR_HPAD = AR_CTAS * AR_HSLOT;               // 1024: padded length of h (column c*8 + w <-> unit 7c + w)
constexpr int AR_NB_MAX = 4;                              // utterances interleaved per launch
constexpr size_t AR_LL_H = 2 * AR_HPAD;                   // LL words per utterance: h [2][128][8]
constexpr size_t AR_LL_R = 2 * AR_CTAS * AR_RSLOT;        //                          r, o [2][128][16]
constexpr size_t AR_LL_PER_UTT = AR_LL_H + 2 * AR_LL_R;
constexpr size_t AR_LL_WORDS = AR_NB_MAX * AR_LL_PER_UTT;

struct ArParams {
    const float* w_hh;      // (2688, 896)
    const float* b_hh;      // (2688,)
    const float* fc1_w;     // (256, 896)
    const float* fc1_b;
    const float* fc2_w;     // (256, 256)
    const float* fc2_b;
    const float* eprime;    // (256, 2688)
    const float* lut;       // (256,)
    // per-utterance arrays of this launch's group: utterance nb lives at base + nb * stride
    const float* G;         // (T2, 2688)           stride g_stride
    const float* uniforms;  // (L,)   generate mode  stride L
    const int64_t* x_in;    // (L,)   teacher-forced mode (nullptr in generate mode), stride L
    float* out_wav;         // (L,) or null
    int32_t* out_codes;     // (L,) or null
    float* out_logits;      // (L, 256) or null
    ll_word* ll;            // [nb][ h: 2*1024 | r: 2*128*16 | o: 2*128*16 ]
    int* status;
    long long* trace;       // optional phase timestamps (debug): [trace_n][8] clock64 values of CTA trace_cta, utterance 0
    int trace_cta, trace_t0, trace_n;
    long long g_stride;
    int L, upsample, nb_active;
    int poll_delay, poll_backoff;   // cycles before the first poll round / between failed rounds
};

// Poll N2 16-byte pairs (lane reads pairs at base + 2*k*pair_stride) until every word carries `tag`.
// The first round is issued `first_delay` cycles after the call, failed rounds are retried after `backoff`.
template <int N2>
__device__ __forceinline__ bool ll_poll(const ll_word* base, int pair_stride, uint32_t tag, float* out,
                                        int first_delay, int backoff, volatile int* abort_flag) {
    ll_word a0[N2], b0[N2];
    const long long t0 = clock64();
    if (first_delay) while (clock64() - t0 < first_delay) {}
    for (;;) {
#pragma unroll
        for (int k = 0; k < N2; ++k) ll_load2(base + 2 * k * pair_stride, a0[k], b0[k]);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < N2; ++k) ok = ok && (ll_tag(a0[k]) == tag) && (ll_tag(b0[k]) == tag);
        if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
            for (int k = 0; k < N2; ++k) { out[2 * k] = ll_val(a0[k]); out[2 * k + 1] = ll_val(b0[k]); }
            return true;
        }
        if (*abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) return false;
        if (backoff) { const long long t1 = clock64(); while (clock64() - t1 < backoff) {} }
    }
}

// wait until a shared-memory sequence flag reaches `want` (set by another role warp of this CTA)
__device__ __forceinline__ bool wait_seq(volatile int* flag, int want, volatile int* abort_flag) {
    const long long t0 = clock64();
    while (*flag < want) {
        if (*abort_flag || clock64() - t0 > LL_TIMEOUT_CYCLES) return false;
    }
    return true;
}

template <int NB>
__global__ void __launch_bounds__(AR_THREADS, 1) ar_kernel(ArParams p) {
    extern __shared__ __align__(16) float ar_dyn_smem[];
    float (*hs)[2][AR_HPAD] = reinterpret_cast<float (*)[2][AR_HPAD]>(ar_dyn_smem);   // [NB][2][1024]
    __shared__ float hhpart[NB][2][AR_HH_WARPS][AR_NROW + 3];   // per-warp partial sums of W_hh h_t (column quarters)
    __shared__ float Es[AR_Q * AR_NROW];
    __shared__ volatile int abort_flag;
    __shared__ volatile int seq_h[NB], seq_r[NB], seq_o[NB];    // steps whose h / r / o this CTA has published

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const int L = p.L, nba = p.nb_active;

    for (int i = tid; i < AR_Q * AR_NROW; i += AR_THREADS) {
        const int x = i / AR_NROW, j = i % AR_NROW, u = j / 3, g = j % 3;
        Es[i] = __ldg(p.eprime + static_cast<int64_t>(x) * AR_G + g * AR_H + cta * AR_U + u);
    }
    for (int i = tid; i < NB * AR_HH_WARPS * (AR_NROW + 3); i += AR_THREADS) {
        const int nb = i / (AR_HH_WARPS * (AR_NROW + 3)), rest = i % (AR_HH_WARPS * (AR_NROW + 3));
        hhpart[nb][1][rest / (AR_NROW + 3)][rest % (AR_NROW + 3)] = 0.f;   // W_hh h_{-1} with h_{-1} = 0
    }
    if (tid < NB) { seq_h[tid] = 0; seq_r[tid] = 0; seq_o[tid] = 0; }
    if (tid == 0) abort_flag = 0;
    __syncthreads();

    const bool teacher = (p.x_in != nullptr);
    const bool tracing = (p.trace != nullptr) && (cta == p.trace_cta) && (lane == 0);
#define AR_TRACE(k, tt)                                                                              \
    if (tracing && nb == 0 && (tt) >= p.trace_t0 && (tt) < p.trace_t0 + p.trace_n) p.trace[((tt) - p.trace_t0) * 8 + (k)] = clock64();
#define AR_FAIL()                                                          \
    do {                                                                   \
        abort_flag = 1;                                                    \
        if (lane == 0) atomicExch(p.status, VQCPC_ERR_TIMEOUT);            \
        __threadfence_block();                                             \
    } while (0)

    if (warp >= AR_ROLE_WARPS) {
        // ----------------------------------------------------------------- W_hh warps (off the critical path)
        // warp q holds columns {224q + 32k + lane : k < 7} of all 21 rows (unit u, gate g -> row index 3u + g)
        const int q = warp - AR_ROLE_WARPS;
        float w[AR_NROW][7];
        int poff[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int col = 224 * q + 32 * k + lane;
            poff[k] = (col / AR_U) * AR_HSLOT + col % AR_U;
        }
#pragma unroll
        for (int r = 0; r < AR_NROW; ++r) {
            const int u = r / 3, g = r % 3;
            const float* row = p.w_hh + static_cast<int64_t>(g * AR_H + cta * AR_U + u) * AR_H + 224 * q + lane;
#pragma unroll
            for (int k = 0; k < 7; ++k) w[r][k] = __ldg(row + 32 * k);
        }
        bool dead = false;
        for (int t = 0; t < L && !dead; ++t) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                if (nb >= nba || dead) continue;
                bar_sync(1 + 2 * nb, AR_BAR_COUNT);       // h_t[nb] is in hs[nb][t&1]
                if (abort_flag) { dead = true; continue; }
                const float* h = hs[nb][t & 1];
                float hv[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) hv[k] = h[poff[k]];
                float acc[AR_NROW];
#pragma unroll
                for (int r = 0; r < AR_NROW; ++r) {
                    float a = __fmul_rn(w[r][0], hv[0]);
#pragma unroll
                    for (int k = 1; k < 7; ++k) a = __fmaf_rn(w[r][k], hv[k], a);
                    acc[r] = a;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int r = 0; r < AR_NROW; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                }
                float mine = acc[0];
#pragma unroll
                for (int r = 1; r < AR_NROW; ++r) mine = (lane == r) ? acc[r] : mine;
                if (lane < AR_NROW) hhpart[nb][t & 1][q][lane] = mine;
                bar_arrive(2 + 2 * nb, AR_BAR_COUNT);
            }
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) bar_arrive(2 + 2 * nb, AR_BAR_COUNT);   // on abort: never leave role O waiting
        return;
    }

    if (warp == 0) {
        // ----------------------------------------------------------------- role H: h_t -> fc1 rows -> r_t
        // pair k of lane l = words {64k + 2l, 64k + 2l + 1} of the padded h vector: producer c = 8k + l/4,
        // units 7c + 2(l%4) + {0,1} (the second one of l%4 == 3 is the pad word).
        float w1[AR_R][32], b1[AR_R];
#pragma unroll
        for (int r = 0; r < AR_R; ++r) {
            const float* row = p.fc1_w + static_cast<int64_t>(cta * AR_R + r) * AR_H;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int c = 8 * k + (lane >> 2), wd = 2 * (lane & 3);
                w1[r][2 * k] = __ldg(row + AR_U * c + wd);
                w1[r][2 * k + 1] = (wd + 1 < AR_U) ? __ldg(row + AR_U * c + wd + 1) : 0.f;
            }
            b1[r] = __ldg(p.fc1_b + cta * AR_R + r);
        }
        bool dead = false;
        for (int t = 0; t < L && !dead; ++t) {
            const uint32_t tag = static_cast<uint32_t>(t) + 1u;
            const int par = t & 1;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                if (nb >= nba || dead) continue;
                ll_word* llb = p.ll + nb * AR_LL_PER_UTT;
                if (!wait_seq(&seq_h[nb], t + 1, &abort_flag)) { AR_FAIL(); dead = true; continue; }
                AR_TRACE(1, t)
                float hv[32];
                if (!ll_poll<16>(llb + par * AR_HPAD + 2 * lane, 32, tag, hv, p.poll_delay, p.poll_backoff, &abort_flag)) {
                    AR_FAIL(); dead = true; continue;
                }
                AR_TRACE(2, t)
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    *reinterpret_cast<float2*>(&hs[nb][par][64 * k + 2 * lane]) = make_float2(hv[2 * k], hv[2 * k + 1]);
                bar_arrive(1 + 2 * nb, AR_BAR_COUNT);     // W_hh warps may start on h_t[nb]
                float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    s0a = fmaf(w1[0][2 * k], hv[2 * k], s0a); s0b = fmaf(w1[0][2 * k + 1], hv[2 * k + 1], s0b);
                    s1a = fmaf(w1[1][2 * k], hv[2 * k], s1a); s1b = fmaf(w1[1][2 * k + 1], hv[2 * k + 1], s1b);
                }
                float s0 = s0a + s0b, s1 = s1a + s1b;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {       // two interleaved butterflies
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (lane < AR_R) {
                    const float v = fmaxf((lane ? s1 : s0) + (lane ? b1[1] : b1[0]), 0.f);
                    ll_store(llb + AR_LL_H + (par * AR_CTAS + cta) * AR_RSLOT + lane, v, tag);
                }
                __syncwarp();
                if (lane == 0) seq_r[nb] = t + 1;
                AR_TRACE(3, t)
            }
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) bar_arrive(1 + 2 * nb, AR_BAR_COUNT);   // on abort: release the W_hh warps
        return;
    }

    if (warp == 1) {
        // ----------------------------------------------------------------- role R: r_t -> fc2 rows -> o_t
        // pair k of lane l = the two r values of producer 32k + l: columns 2(32k + l) + {0,1}
        float w2[AR_R][8], b2[AR_R];
#pragma unroll
        for (int r = 0; r < AR_R; ++r) {
            const float* row = p.fc2_w + static_cast<int64_t>(cta * AR_R + r) * AR_FC;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(row + 64 * k) + lane);
                w2[r][2 * k] = v.x; w2[r][2 * k + 1] = v.y;
            }
            b2[r] = __ldg(p.fc2_b + cta * AR_R + r);
        }
        bool dead = false;
        for (int t = 0; t < L && !dead; ++t) {
            const uint32_t tag = static_cast<uint32_t>(t) + 1u;
            const int par = t & 1;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                if (nb >= nba || dead) continue;
                ll_word* llb = p.ll + nb * AR_LL_PER_UTT;
                if (!wait_seq(&seq_r[nb], t + 1, &abort_flag)) { AR_FAIL(); dead = true; continue; }
                float rv[8];
                if (!ll_poll<4>(llb + AR_LL_H + (par * AR_CTAS + lane) * AR_RSLOT, 32 * AR_RSLOT / 2, tag, rv, p.poll_delay,
                                p.poll_backoff, &abort_flag)) {
                    AR_FAIL(); dead = true; continue;
                }
                AR_TRACE(4, t)
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) { s0 = fmaf(w2[0][k], rv[k], s0); s1 = fmaf(w2[1][k], rv[k], s1); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                const float v = (lane & 1 ? s1 : s0) + (lane & 1 ? b2[1] : b2[0]);
                if (p.out_logits != nullptr && lane < AR_R)
                    p.out_logits[(static_cast<int64_t>(nb) * L + t) * AR_Q + cta * AR_R + lane] = v;
                if (!teacher && lane < AR_R) ll_store(llb + AR_LL_H + AR_LL_R + (par * AR_CTAS + cta) * AR_RSLOT + lane, v, tag);
                __syncwarp();
                if (lane == 0) seq_o[nb] = t + 1;    // teacher-forced mode too: role O throttles on it (see below)
                AR_TRACE(5, t)
            }
        }
        return;
    }

    // --------------------------------------------------------------------- role O
    // for t = 0..L:  [t > 0: poll o_{t-1} -> sample x_{t-1} -> output]  [t < L: gates_t -> publish h_t]
    // lanes 0..6 <-> hidden units 7*cta + lane; lane 7 publishes the pad word.
    const bool gate_lane = lane < AR_U;
    const int gu = cta * AR_U + (gate_lane ? lane : 0);
    float hown[NB], g_r[NB], g_z[NB], g_n[NB], u_next[NB];
    int x[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) { hown[nb] = 0.f; g_r[nb] = g_z[nb] = g_n[nb] = 0.f; u_next[nb] = 0.f; x[nb] = X_INIT; }
    const float bh_r = __ldg(p.b_hh + gu), bh_z = __ldg(p.b_hh + AR_H + gu), bh_n = __ldg(p.b_hh + 2 * AR_H + gu);
    int frame_left = 0, frame = 0;
    bool dead = false;
    for (int t = 0; t <= L && !dead; ++t) {
        const int par = t & 1;
        if (t < L) {
            if (frame_left == 0) {
                if (gate_lane) {
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        if (nb >= nba) continue;
                        const float* g = p.G + nb * p.g_stride + static_cast<int64_t>(frame) * AR_G + gu;
                        g_r[nb] = __ldg(g); g_z[nb] = __ldg(g + AR_H); g_n[nb] = __ldg(g + 2 * AR_H);
                    }
                }
                frame_left = p.upsample; ++frame;
            }
            --frame_left;
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            if (nb >= nba || dead) continue;
            ll_word* llb = p.ll + nb * AR_LL_PER_UTT;
            AR_TRACE(0, t)
            if (t > 0 && !teacher) {
                // ---- sample x_{t-1}[nb] from o_{t-1}: lane l holds classes {64k + 2l + j : k < 4, j < 2}
                const int tp = t - 1, ppar = tp & 1;
                const float u_t = u_next[nb];
                if (!wait_seq(&seq_o[nb], t, &abort_flag)) { AR_FAIL(); dead = true; continue; }
                float ov[8];
                if (!ll_poll<4>(llb + AR_LL_H + AR_LL_R + (ppar * AR_CTAS + lane) * AR_RSLOT, 32 * AR_RSLOT / 2,
                                static_cast<uint32_t>(t), ov, p.poll_delay, p.poll_backoff, &abort_flag)) {
                    AR_FAIL(); dead = true; continue;
                }
                AR_TRACE(6, tp)
                float m = fmaxf(fmaxf(fmaxf(ov[0], ov[1]), fmaxf(ov[2], ov[3])), fmaxf(fmaxf(ov[4], ov[5]), fmaxf(ov[6], ov[7])));
                m = warp_max(m);
                float e0[4], e1[4], inc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    e0[k] = __expf(ov[2 * k] - m); e1[k] = __expf(ov[2 * k + 1] - m);
                    inc[k] = e0[k] + e1[k];
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {       // four interleaved inclusive scans over lanes
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float v = __shfl_up_sync(0xffffffffu, inc[k], o);
                        if (lane >= o) inc[k] += v;
                    }
                }
                float pre[4];                             // sum of all classes below block k
                float run = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) { pre[k] = run; run += __shfl_sync(0xffffffffu, inc[k], 31); }
                const float thr = u_t * run;
                int xs = AR_Q - 1;
                bool found = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float c1 = pre[k] + inc[k];                 // cumulative sum through class 64k + 2l + 1
                    const float c0 = pre[k] + (inc[k] - e1[k]);       //                 through class 64k + 2l
                    const unsigned h1 = __ballot_sync(0xffffffffu, c1 > thr);
                    const unsigned h0 = __ballot_sync(0xffffffffu, c0 > thr);
                    if (!found && h1 != 0u) {
                        const int src = __ffs(h1) - 1;
                        xs = 64 * k + 2 * src + (((h0 >> src) & 1u) ? 0 : 1);
                        found = true;
                    }
                }
                x[nb] = xs;
                if (cta == 0 && lane == 0) {
                    if (p.out_wav) p.out_wav[static_cast<int64_t>(nb) * L + tp] = __ldg(p.lut + xs);
                    if (p.out_codes) p.out_codes[static_cast<int64_t>(nb) * L + tp] = xs;
                }
                AR_TRACE(7, tp)
            }
            if (t < L) {
                // ---- gates of step t for utterance nb
                if (teacher) {
                    x[nb] = static_cast<int>(__ldg(p.x_in + static_cast<int64_t>(nb) * L + t)) & (AR_Q - 1);
                    // Teacher-forced mode has no sampling, so nothing in the h -> W_hh -> gates chain waits for role R.  The
                    // double-buffered r slots are only WAR-safe if no CTA publishes h_t before its own role R has consumed
                    // r_{t-2}: a CTA that sees every h_t (and may then overwrite its r slot of step t-2 with r_t) knows that
                    // every role R is past step t-2.
                    if (t >= 2 && !wait_seq(&seq_o[nb], t - 1, &abort_flag)) { AR_FAIL(); dead = true; continue; }
                }
                if (t > 0) bar_sync(2 + 2 * nb, AR_BAR_COUNT);   // quarter sums of W_hh h_{t-1}[nb] are in hhpart[nb][par ^ 1]
                if (abort_flag) { dead = true; continue; }
                const uint32_t tag = static_cast<uint32_t>(t) + 1u;
                float hnew = 0.f;
                if (gate_lane) {
                    const float* e = &Es[x[nb] * AR_NROW + lane * 3];
                    const float* hp = &hhpart[nb][par ^ 1][0][lane * 3];
                    constexpr int HS = AR_NROW + 3;
                    const float hb_r = __fadd_rn(__fadd_rn(__fadd_rn(hp[0], hp[HS]), __fadd_rn(hp[2 * HS], hp[3 * HS])), bh_r);
                    const float hb_z = __fadd_rn(__fadd_rn(__fadd_rn(hp[1], hp[HS + 1]), __fadd_rn(hp[2 * HS + 1], hp[3 * HS + 1])), bh_z);
                    const float hb_n = __fadd_rn(__fadd_rn(__fadd_rn(hp[2], hp[HS + 2]), __fadd_rn(hp[2 * HS + 2], hp[3 * HS + 2])), bh_n);
                    // explicit FMA forms: identical bits in every template instantiation / batch composition
                    const float r = sigmoid_fast(__fadd_rn(__fadd_rn(e[0], g_r[nb]), hb_r));
                    const float z = sigmoid_fast(__fadd_rn(__fadd_rn(e[1], g_z[nb]), hb_z));
                    const float n = tanh_fast(__fmaf_rn(r, hb_n, __fadd_rn(e[2], g_n[nb])));
                    hnew = __fmaf_rn(z, __fsub_rn(hown[nb], n), n);          // (1-z) n + z h
                    hown[nb] = hnew;
                }
                if (lane < AR_HSLOT) ll_store(llb + par * AR_HPAD + cta * AR_HSLOT + lane, hnew, tag);
                __syncwarp();
                if (lane == 0) seq_h[nb] = t + 1;
                if (!teacher) u_next[nb] = __ldg(p.uniforms + static_cast<int64_t>(nb) * L + t);   // consumed one step later
            }
        }
    }
#undef AR_TRACE
#undef AR_FAIL
}


// ------------------------------------------------------------------------------------------------
// Exchange floor (diagnostic, used by bench.py): the bare 128-way all-to-all of the sample loop -- every CTA
// publishes 2 LL words into its own 128-byte slot and polls all 128 slots -- with no compute in between.
// Three of these per step are the latency floor of the role-warp design on this GPU.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1) exchange_floor_kernel(ll_word* buf, int iters, long long* cycles) {
    const int lane = threadIdx.x, cta = blockIdx.x;
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        const int par = it & 1;
        if (lane < 2) ll_store(buf + (par * AR_CTAS + cta) * AR_RSLOT + lane, 0.f, static_cast<uint32_t>(it));
        const long long ts = clock64();
        for (;;) {
            ll_word a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) ll_load2(buf + (par * AR_CTAS + 32 * k + lane) * AR_RSLOT, a[k], b[k]);
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) ok = ok && ll_tag(a[k]) == static_cast<uint32_t>(it) && ll_tag(b[k]) == static_cast<uint32_t>(it);
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - ts > LL_TIMEOUT_CYCLES) { if (lane == 0) cycles[cta] = -1; return; }
        }
    }
    if (lane == 0) cycles[cta] = clock64() - t0;
}

// ------------------------------------------------------------------------------------------------ host
int ar_batch_run(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int B, int T2,
                 int L, void* ws, int* status, float* out_wav, int32_t* out_codes, float* out_logits, cudaStream_t stream);
constexpr int AR_BATCH_MIN_B = 8;     // from this many utterances on, the grid-barrier batched kernel wins
// vocoder_cluster.cu: the cluster / DSMEM latency kernel (one utterance per launch)
extern int g_cl_enable;
int ar_cluster_supported();
size_t ar_cluster_ll_bytes();
int ar_cluster_launch(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int L,
                      ll_word* hbuf, int* status, float* out_wav, int32_t* out_codes, float* out_logits, long long* trace,
                      int trace_cta, int trace_t0, int trace_n, cudaStream_t stream);
void ar_cluster_set_poll(int delay, int mode);
int ar_cluster_exchange_floor(void* workspace, size_t workspace_bytes, int iters, double* mean_cycles, cudaStream_t s);
static size_t ar_ws_bytes() {
    const size_t ll = align_up(sizeof(ll_word) * AR_LL_WORDS, 256), ab = ar_batch_workspace_bytes();
    return sizeof(WorkspaceHeader) + (ll > ab ? ll : ab);
}

// workspace layout: [header][u B*2Tc*128][xproj B*2Tc*768][p0 B*2Tc*256][p1 B*2Tc*256][ll words]
static size_t vocoder_ws_bytes(int B, int Tc) {
    const size_t rows = static_cast<size_t>(B) * 2 * Tc;
    return sizeof(WorkspaceHeader) + align_up(rows * 128 * 4, 256) + align_up(rows * 768 * 4, 256) +
           2 * align_up(rows * 256 * 4, 256) + ar_ws_bytes();
}
static int check_vocoder_dims(const vqcpc_vocoder_weights* w) {
    VQ_ARG(w != nullptr, "vocoder: null weights");
    VQ_ARG(w->dim_code + w->dim_speaker == 128 && w->dim_code % 4 == 0 && w->dim_speaker % 4 == 0,
           "vocoder: dim_i_embedding + dim_speaker_embedding must be 128 (config.py:199)");
    VQ_ARG(w->upsample_t > 0, "vocoder: upsampling_t must be positive");
    return VQCPC_OK;
}

int vocoder_pack(const vqcpc_vocoder_weights* w, float* eprime_out, cudaStream_t stream) {
    VQ_ARG(w && eprime_out && w->ar_emb && w->ar_w_ih, "vocoder_pack: null pointer");
    // E'[x, row] = sum_k emb[x,k] * W_ih[row, k], k < 256  (W_ih row stride 512)
    return gemm_dense(w->ar_emb, AR_EMB, w->ar_w_ih, AR_EMB + AR_COND, nullptr, eprime_out, AR_G, AR_Q, AR_G, AR_EMB,
                      stream);
}

int vocoder_condition(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker, const int32_t* code_lengths,
                      int B, int Tc, void* ws, size_t ws_bytes, float* out_G, float* out_p, cudaStream_t stream) {
    int rc = check_vocoder_dims(w);
    if (rc) return rc;
    VQ_ARG(codes && speaker && ws && out_G, "vocoder_condition: null pointer");
    VQ_ARG(B >= 0 && Tc >= 1, "vocoder_condition: bad shape B=%d Tc=%d", B, Tc);
    VQ_ARG(ws_bytes >= vocoder_ws_bytes(B, Tc), "vocoder_condition: workspace too small");
    if (B == 0) return VQCPC_OK;
    const int T2 = 2 * Tc;
    const int64_t rows = static_cast<int64_t>(B) * T2;
    unsigned char* base = static_cast<unsigned char*>(ws);
    size_t off = sizeof(WorkspaceHeader);
    float* u = reinterpret_cast<float*>(base + off); off += align_up(rows * 128 * 4, 256);
    float* xproj = reinterpret_cast<float*>(base + off); off += align_up(rows * 768 * 4, 256);
    float* p0 = reinterpret_cast<float*>(base + off); off += align_up(rows * 256 * 4, 256);
    float* p1 = out_p ? out_p : reinterpret_cast<float*>(base + off);

    {
        const int64_t total = rows * 32;
        const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        WorkspaceHeader* hdr = reinterpret_cast<WorkspaceHeader*>(base);
        VQ_CUDA(cudaMemsetAsync(hdr, 0, sizeof(WorkspaceHeader), stream));
        embed_concat_kernel<<<grid, 256, 0, stream>>>(w->code_emb, w->spk_emb, codes, speaker, u, B, Tc, w->dim_code,
                                                      w->dim_speaker, w->n_codes, w->n_speakers, &hdr->index_error);
        VQ_CUDA(cudaGetLastError());
        count_launch(1);
    }
    // layer 0
    if ((rc = gemm_dense(u, 128, w->pre_w_ih[0], 128, w->pre_b_ih[0], xproj, 768, rows, 768, 128, stream))) return rc;
    bigru_layer_kernel<<<dim3(2, B), PRE_G, 0, stream>>>(xproj, w->pre_w_hh[0], w->pre_b_hh[0], p0, T2, code_lengths);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    // layer 1 (input = [fwd;bwd] of layer 0)
    if ((rc = gemm_dense(p0, 256, w->pre_w_ih[1], 256, w->pre_b_ih[1], xproj, 768, rows, 768, 256, stream))) return rc;
    bigru_layer_kernel<<<dim3(2, B), PRE_G, 0, stream>>>(xproj, w->pre_w_hh[1], w->pre_b_hh[1], p1, T2, code_lengths);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    // hoisted conditioning half of the AR input projection: G = p . W_ih[:, 256:]^T + b_ih
    return gemm_dense(p1, 256, w->ar_w_ih + AR_EMB, AR_EMB + AR_COND, w->ar_b_ih, out_G, AR_G, rows, AR_G, AR_COND,
                      stream);
}

static int g_poll_gap = 400;     // bits 0..11: first-poll delay, bits 12..23: backoff (cycles); see vqcpc_debug_set_ar_poll_gap
static int g_nb_cap = AR_NB_MAX;
static int g_batch_min_b = AR_BATCH_MIN_B;
struct ArTrace { long long* buf; int cta, t0, n; };
static ArTrace g_trace = {nullptr, 0, 0, 0};

template <int NB>
static int ar_launch(ArParams& p, cudaStream_t stream) {
    void* args[] = {&p};
    constexpr size_t dyn = sizeof(float) * NB * 2 * AR_HPAD;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(ar_kernel<NB>), static_cast<int>(dyn))) return rc_attr;
    VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ar_kernel<NB>), dim3(AR_CTAS), dim3(AR_THREADS), args, dyn,
                                        stream));
    count_launch(1);
    return VQCPC_OK;
}

static int ar_run(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int B,
                  int T2, int L, void* ws, size_t ws_bytes, float* out_wav, int32_t* out_codes, float* out_logits,
                  cudaStream_t stream) {
    int rc = check_vocoder_dims(w);
    if (rc) return rc;
    VQ_ARG(G && ws, "vocoder: null pointer");
    VQ_ARG(w->eprime && w->mulaw_lut, "vocoder: weights not packed (eprime / mulaw_lut missing)");
    VQ_ARG(B >= 0 && T2 >= 1 && L >= 0, "vocoder: bad shape");
    VQ_ARG(static_cast<int64_t>(L) <= static_cast<int64_t>(T2) * w->upsample_t,
           "vocoder: L=%d exceeds conditioning length %d*%d", L, T2, w->upsample_t);
    VQ_ARG(ws_bytes >= ar_ws_bytes(), "vocoder: workspace too small");
    if (B == 0 || L == 0) return VQCPC_OK;
    if (device_sm_count() < AR_CTAS) {
        set_error("vocoder: the persistent sample loop needs %d co-resident CTAs (device has %d SMs)", AR_CTAS,
                  device_sm_count());
        return VQCPC_ERR_DEVICE;
    }
    unsigned char* base = static_cast<unsigned char*>(ws);
    WorkspaceHeader* hdr = reinterpret_cast<WorkspaceHeader*>(base);
    ll_word* ll = reinterpret_cast<ll_word*>(base + sizeof(WorkspaceHeader));
    VQ_CUDA(cudaMemsetAsync(&hdr->status, 0, sizeof(int), stream));   // index_error of a preceding vocoder_condition survives
    if (B >= g_batch_min_b)
        return ar_batch_run(w, G, uniforms, x_in, B, T2, L, base + sizeof(WorkspaceHeader), &hdr->status, out_wav, out_codes,
                            out_logits, stream);
    if (g_cl_enable && ar_cluster_supported()) {
        // one cluster-kernel launch per utterance (7 clusters x 16 CTAs, one grid-scope exchange per step)
        for (int b = 0; b < B; ++b) {
            rc = ar_cluster_launch(w, G + static_cast<int64_t>(b) * T2 * AR_G, uniforms ? uniforms + static_cast<int64_t>(b) * L : nullptr,
                                   x_in ? x_in + static_cast<int64_t>(b) * L : nullptr, L, ll, &hdr->status,
                                   out_wav ? out_wav + static_cast<int64_t>(b) * L : nullptr,
                                   out_codes ? out_codes + static_cast<int64_t>(b) * L : nullptr,
                                   out_logits ? out_logits + static_cast<int64_t>(b) * L * AR_Q : nullptr,
                                   b == 0 ? g_trace.buf : nullptr, g_trace.cta, g_trace.t0, g_trace.n, stream);
            if (rc) return rc;
        }
        return VQCPC_OK;
    }
    // fallback (device cannot co-schedule 7 clusters of 16): groups of up to NB utterances share one persistent launch
    for (int b = 0; b < B;) {
        int nb = B - b < g_nb_cap ? B - b : g_nb_cap;
        VQ_CUDA(cudaMemsetAsync(ll, 0, sizeof(ll_word) * AR_LL_PER_UTT * nb, stream));
        ArParams p{};
        p.w_hh = w->ar_w_hh; p.b_hh = w->ar_b_hh;
        p.fc1_w = w->fc1_w; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b;
        p.eprime = w->eprime;
        p.lut = w->mulaw_lut;
        p.G = G + static_cast<int64_t>(b) * T2 * AR_G;
        p.g_stride = static_cast<long long>(T2) * AR_G;
        p.uniforms = uniforms ? uniforms + static_cast<int64_t>(b) * L : nullptr;
        p.x_in = x_in ? x_in + static_cast<int64_t>(b) * L : nullptr;
        p.out_wav = out_wav ? out_wav + static_cast<int64_t>(b) * L : nullptr;
        p.out_codes = out_codes ? out_codes + static_cast<int64_t>(b) * L : nullptr;
        p.out_logits = out_logits ? out_logits + static_cast<int64_t>(b) * L * AR_Q : nullptr;
        p.ll = ll;
        p.trace = g_trace.buf; p.trace_cta = g_trace.cta; p.trace_t0 = g_trace.t0; p.trace_n = g_trace.n;
        p.status = &hdr->status;
        p.L = L; p.upsample = w->upsample_t; p.nb_active = nb;
        p.poll_delay = g_poll_gap & 0xfff; p.poll_backoff = (g_poll_gap >> 12) & 0xfff;
        if (nb == 1) rc = ar_launch<1>(p, stream);
        else if (nb == 2) rc = ar_launch<2>(p, stream);
        else rc = ar_launch<4>(p, stream);
        if (rc) return rc;
        b += nb;
    }
    return VQCPC_OK;
}

}  // namespace vqcpc

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int vqcpc_debug_exchange_floor(void* workspace, size_t workspace_bytes, int32_t iters, double* mean_cycles,
                                          void* stream) {
    using namespace vqcpc;
    VQ_ARG(workspace && mean_cycles && iters > 0, "exchange_floor: bad arguments");
    // the exchange of the kernel generate actually runs: the cluster kernel's 112-CTA all-gather of h_t when it is in use
    if (g_cl_enable && ar_cluster_supported())
        return ar_cluster_exchange_floor(workspace, workspace_bytes, iters, mean_cycles, static_cast<cudaStream_t>(stream));
    const size_t need = sizeof(ll_word) * AR_LL_R + sizeof(long long) * AR_CTAS;
    VQ_ARG(workspace_bytes >= need, "exchange_floor: workspace too small");
    if (device_sm_count() < AR_CTAS) { set_error("exchange_floor: needs %d SMs", AR_CTAS); return VQCPC_ERR_DEVICE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ll_word* buf = static_cast<ll_word*>(workspace);
    long long* cyc = reinterpret_cast<long long*>(buf + AR_LL_R);
    VQ_CUDA(cudaMemsetAsync(workspace, 0, need, s));
    void* args[] = {&buf, &iters, &cyc};
    VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(exchange_floor_kernel), dim3(AR_CTAS), dim3(32), args, 0, s));
    count_launch(1);
    long long host[AR_CTAS];
    VQ_CUDA(cudaMemcpyAsync(host, cyc, sizeof(host), cudaMemcpyDeviceToHost, s));
    VQ_CUDA(cudaStreamSynchronize(s));
    double sum = 0;
    for (int i = 0; i < AR_CTAS; ++i) {
        if (host[i] < 0) { set_error("exchange_floor: timed out"); return VQCPC_ERR_TIMEOUT; }
        sum += static_cast<double>(host[i]);
    }
    *mean_cycles = sum / AR_CTAS / iters;
    return VQCPC_OK;
}
namespace vqcpc { extern int g_ab_two_group; }
extern "C" int vqcpc_debug_set_ar_poll_gap(int32_t packed) {
    vqcpc::g_poll_gap = packed < 0 ? 0 : (packed & 0xffffff);
    const int cap = (packed >> 24) & 0xf;          // bits 24..27: cap on utterances per launch (0 = default)
    vqcpc::g_nb_cap = (cap >= 1 && cap <= vqcpc::AR_NB_MAX) ? cap : vqcpc::AR_NB_MAX;
    vqcpc::g_batch_min_b = ((packed >> 28) & 1) ? (1 << 30) : vqcpc::AR_BATCH_MIN_B;   // bit 28: disable the batched kernel
    vqcpc::g_ab_two_group = ((packed >> 29) & 1) ? 0 : 1;                                // bit 29: disable its two-group variant
    return VQCPC_OK;
}
extern "C" int vqcpc_debug_set_ar_cluster(int32_t enable, int32_t first_poll_delay, int32_t poll_mode) {
    vqcpc::g_cl_enable = enable ? 1 : 0;
    vqcpc::ar_cluster_set_poll(first_poll_delay < 0 ? 0 : first_poll_delay, poll_mode);
    return VQCPC_OK;
}
// 1 = single-utterance generate runs on the cluster kernel on the current device, 0 = on the round-1 128-CTA kernel (forced by
// vqcpc_debug_set_ar_cluster(0, ..) / VQCPC_AR_CLUSTER=0, a device that cannot co-schedule 7 x 16 CTAs, or Nsight Compute attached)
extern "C" int vqcpc_ar_cluster_active(void) { return (vqcpc::g_cl_enable && vqcpc::ar_cluster_supported()) ? 1 : 0; }
namespace vqcpc { extern long long* g_ab_trace; extern int g_ab_trace_cta, g_ab_trace_t0, g_ab_trace_n; }
extern "C" int vqcpc_debug_set_ar_trace(long long* device_buf, int32_t cta, int32_t first_step, int32_t n_steps) {
    vqcpc::g_trace = vqcpc::ArTrace{device_buf, cta, first_step, n_steps};
    vqcpc::g_ab_trace = device_buf; vqcpc::g_ab_trace_cta = cta; vqcpc::g_ab_trace_t0 = first_step; vqcpc::g_ab_trace_n = n_steps;
    return VQCPC_OK;
}
extern "C" int vqcpc_vocoder_pack(const vqcpc_vocoder_weights* w, float* eprime_out, void* stream) {
    return vqcpc::vocoder_pack(w, eprime_out, static_cast<cudaStream_t>(stream));
}
extern "C" size_t vqcpc_vocoder_workspace_bytes(int32_t B, int32_t Tc) { return vqcpc::vocoder_ws_bytes(B, Tc); }
extern "C" int vqcpc_vocoder_condition(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker,
                                       int32_t B, int32_t Tc, void* workspace, size_t workspace_bytes, float* out_G,
                                       float* out_p, void* stream) {
    return vqcpc::vocoder_condition(w, codes, speaker, nullptr, B, Tc, workspace, workspace_bytes, out_G, out_p,
                                    static_cast<cudaStream_t>(stream));
}
extern "C" int vqcpc_vocoder_condition_ragged(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker,
                                              const int32_t* code_lengths, int32_t B, int32_t Tc, void* workspace,
                                              size_t workspace_bytes, float* out_G, float* out_p, void* stream) {
    if (code_lengths == nullptr) { vqcpc::set_error("vocoder_condition_ragged: code_lengths is required"); return VQCPC_ERR_ARG; }
    return vqcpc::vocoder_condition(w, codes, speaker, code_lengths, B, Tc, workspace, workspace_bytes, out_G, out_p,
                                    static_cast<cudaStream_t>(stream));
}
extern "C" int vqcpc_vocoder_generate(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, int32_t B,
                                      int32_t T2, int32_t L, void* workspace, size_t workspace_bytes, float* out_wav,
                                      int32_t* out_codes, float* out_logits, void* stream) {
    if (uniforms == nullptr || out_wav == nullptr) {
        vqcpc::set_error("vocoder_generate: uniforms and out_wav are required");
        return VQCPC_ERR_ARG;
    }
    return vqcpc::ar_run(w, G, uniforms, nullptr, B, T2, L, workspace, workspace_bytes, out_wav, out_codes, out_logits,
                         static_cast<cudaStream_t>(stream));
}
extern "C" int vqcpc_vocoder_logits_tf(const vqcpc_vocoder_weights* w, const float* G, const int64_t* x_in, int32_t B,
                                       int32_t T2, int32_t L, void* workspace, size_t workspace_bytes,
                                       float* out_logits, void* stream) {
    if (x_in == nullptr || out_logits == nullptr) {
        vqcpc::set_error("vocoder_logits_tf: x_in and out_logits are required");
        return VQCPC_ERR_ARG;
    }
    return vqcpc::ar_run(w, G, nullptr, x_in, B, T2, L, workspace, workspace_bytes, nullptr, nullptr, out_logits,
                         static_cast<cudaStream_t>(stream));
}
