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
constexpr int AR_THREADS = 32 * (1 + AR_U);   // warp 0 = chain warp, warps 1..7 = W_hh warps
constexpr int X_INIT = 128;

// ------------------------------------------------------------------------------------------------
// u[b,t,:] = [code_emb[codes[b, t>>1]] ; spk_emb[speaker[b]]]      (B, 2Tc, 128)
// ------------------------------------------------------------------------------------------------
__global__ void embed_concat_kernel(const float* __restrict__ code_emb, const float* __restrict__ spk_emb,
                                    const int64_t* __restrict__ codes, const int64_t* __restrict__ speaker,
                                    float* __restrict__ u, int B, int Tc, int dc, int ds) {
    const int width = dc + ds;
    const int64_t total = static_cast<int64_t>(B) * 2 * Tc * (width / 4);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(i % (width / 4));
        const int64_t bt = i / (width / 4);
        const int t = static_cast<int>(bt % (2 * Tc));
        const int b = static_cast<int>(bt / (2 * Tc));
        float4 v;
        if (q * 4 < dc) v = __ldg(reinterpret_cast<const float4*>(code_emb + codes[static_cast<int64_t>(b) * Tc + (t >> 1)] * dc) + q);
        else v = __ldg(reinterpret_cast<const float4*>(spk_emb + speaker[b] * ds) + (q - dc / 4));
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
                   float* __restrict__ out, int T) {
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
    for (int s = 0; s < T; ++s) {
        const int t = dir ? (T - 1 - s) : s;
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
// Grid: 128 CTAs x 256 threads, cooperative (all co-resident).  CTA j owns hidden units 7j..7j+6 (their 21
// W_hh rows in the registers of warps 1..7, one unit per warp, 28 columns per lane), fc1 rows 2j,2j+1 and
// fc2 rows 2j,2j+1 (registers of warp 0), and the 21 matching columns of E' (21.5 KB shared memory).
// Warp 0 ("chain warp") runs the sequential dependency chain of a step alone:
//   gates -> publish h_t -> poll all 896 h -> fc1 rows -> publish -> poll 256 -> fc2 rows -> publish ->
//   poll 256 logits -> softmax + inverse-CDF sample (every CTA redundantly: bit-identical) -> next gates.
// It hands h_t to warps 1..7 through shared memory (named barrier 1); they compute W_hh h_t + b_hh for the
// NEXT step off the critical path and hand the 21 sums back (named barrier 2).
// Teacher-forced mode (Vocoder.forward): x_{t-1} comes from x_in, logits are written out, no third exchange.
// ------------------------------------------------------------------------------------------------
struct ArParams {
    const float* w_hh;      // (2688, 896)
    const float* b_hh;      // (2688,)
    const float* fc1_w;     // (256, 896)
    const float* fc1_b;
    const float* fc2_w;     // (256, 256)
    const float* fc2_b;
    const float* eprime;    // (256, 2688)
    const float* G;         // (T2, 2688) for this utterance
    const float* uniforms;  // (L,)   generate mode
    const int64_t* x_in;    // (L,)   teacher-forced mode (nullptr in generate mode)
    const float* lut;       // (256,)
    float* out_wav;         // (L,) or null
    int32_t* out_codes;     // (L,) or null
    float* out_logits;      // (L, 256) or null
    ll_word* ll_h;          // [2][896]
    ll_word* ll_r;          // [2][256]
    ll_word* ll_o;          // [2][256]
    int* status;
    int L, upsample;
};

constexpr int AR_LL_WORDS = 2 * AR_H + 2 * AR_FC + 2 * AR_Q;

// poll `N2` 16-byte pairs (2 slots each) starting at `base + 2*first_pair`, stride `pair_stride` pairs
template <int N2>
__device__ __forceinline__ bool ll_poll_pairs(const ll_word* base, int pair_stride, uint32_t tag, float* out) {
    const long long t0 = clock64();
    for (;;) {
        ll_word a[N2], b[N2];
#pragma unroll
        for (int k = 0; k < N2; ++k) ll_load2(base + 2 * k * pair_stride, a[k], b[k]);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < N2; ++k) {
            ok = ok && (ll_tag(a[k]) == tag) && (ll_tag(b[k]) == tag);
            out[2 * k] = ll_val(a[k]);
            out[2 * k + 1] = ll_val(b[k]);
        }
        if (__all_sync(0xffffffffu, ok)) return true;
        if (clock64() - t0 > LL_TIMEOUT_CYCLES) return false;
    }
}

__global__ void __launch_bounds__(AR_THREADS, 1) ar_kernel(ArParams p) {
    __shared__ __align__(16) float hs[2][AR_H];
    __shared__ float hhres[2][AR_U * 3 + 3];
    __shared__ float Es[AR_Q * AR_U * 3];
    __shared__ volatile int abort_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const int L = p.L;

    for (int i = tid; i < AR_Q * AR_U * 3; i += AR_THREADS) {
        const int x = i / (AR_U * 3), j = i % (AR_U * 3), u = j / 3, g = j % 3;
        Es[i] = __ldg(p.eprime + static_cast<int64_t>(x) * AR_G + g * AR_H + cta * AR_U + u);
    }
    if (tid < AR_U * 3) {
        const int u = tid / 3, g = tid % 3;
        hhres[1][tid] = __ldg(p.b_hh + g * AR_H + cta * AR_U + u);   // W_hh h_{-1} + b_hh with h_{-1} = 0
    }
    if (tid == 0) abort_flag = 0;
    __syncthreads();

    if (warp > 0) {
        // ----------------------------------------------------------------- W_hh warps (off the critical path)
        const int u = warp - 1, gu = cta * AR_U + u;
        float w[3][28];
        float bias[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float* row = p.w_hh + static_cast<int64_t>(g * AR_H + gu) * AR_H;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + 128 * k) + lane);
                w[g][4 * k] = v.x; w[g][4 * k + 1] = v.y; w[g][4 * k + 2] = v.z; w[g][4 * k + 3] = v.w;
            }
            bias[g] = __ldg(p.b_hh + g * AR_H + gu);
        }
        for (int t = 0; t < L; ++t) {
            bar_sync(1, AR_THREADS);                  // h_t is in hs[t&1]
            if (abort_flag) break;
            const float* h = hs[t & 1];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float4 hv = *reinterpret_cast<const float4*>(&h[128 * k + 4 * lane]);
                a0 = fmaf(w[0][4 * k], hv.x, a0); a0 = fmaf(w[0][4 * k + 1], hv.y, a0);
                a0 = fmaf(w[0][4 * k + 2], hv.z, a0); a0 = fmaf(w[0][4 * k + 3], hv.w, a0);
                a1 = fmaf(w[1][4 * k], hv.x, a1); a1 = fmaf(w[1][4 * k + 1], hv.y, a1);
                a1 = fmaf(w[1][4 * k + 2], hv.z, a1); a1 = fmaf(w[1][4 * k + 3], hv.w, a1);
                a2 = fmaf(w[2][4 * k], hv.x, a2); a2 = fmaf(w[2][4 * k + 1], hv.y, a2);
                a2 = fmaf(w[2][4 * k + 2], hv.z, a2); a2 = fmaf(w[2][4 * k + 3], hv.w, a2);
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane < 3) hhres[t & 1][u * 3 + lane] = (lane == 0 ? a0 : (lane == 1 ? a1 : a2)) + bias[lane];
            bar_arrive(2, AR_THREADS);
        }
        return;
    }

    // --------------------------------------------------------------------- chain warp
    float w1[AR_R][28], w2[AR_R][8], b1[AR_R], b2[AR_R];
#pragma unroll
    for (int r = 0; r < AR_R; ++r) {
        const int row = cta * AR_R + r;
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(p.fc1_w + static_cast<int64_t>(row) * AR_H + 64 * k) + lane);
            w1[r][2 * k] = v.x; w1[r][2 * k + 1] = v.y;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.fc2_w + static_cast<int64_t>(row) * AR_FC + 8 * lane) + k);
            w2[r][4 * k] = v.x; w2[r][4 * k + 1] = v.y; w2[r][4 * k + 2] = v.z; w2[r][4 * k + 3] = v.w;
        }
        b1[r] = __ldg(p.fc1_b + row);
        b2[r] = __ldg(p.fc2_b + row);
    }
    const bool teacher = (p.x_in != nullptr);
    const int gu = cta * AR_U + (lane < AR_U ? lane : 0);
    float hown = 0.f, g_r = 0.f, g_z = 0.f, g_n = 0.f;
    int x = X_INIT;
    bool failed = false;

    for (int t = 0; t < L; ++t) {
        const uint32_t tag = static_cast<uint32_t>(t) + 1u;
        const int par = t & 1;
        if (t % p.upsample == 0 && lane < AR_U) {
            const float* g = p.G + static_cast<int64_t>(t / p.upsample) * AR_G + gu;
            g_r = __ldg(g); g_z = __ldg(g + AR_H); g_n = __ldg(g + 2 * AR_H);
        }
        const float u_t = teacher ? 0.f : __ldg(p.uniforms + t);
        if (teacher) x = static_cast<int>(__ldg(p.x_in + t)) & (AR_Q - 1);

        // gates: lane u <-> hidden unit 7*cta + u
        if (lane < AR_U) {
            const float* e = &Es[x * (AR_U * 3) + lane * 3];
            const float* hb = &hhres[par ^ 1][lane * 3];
            const float r = sigmoid_fast(e[0] + g_r + hb[0]);
            const float z = sigmoid_fast(e[1] + g_z + hb[1]);
            const float n = tanh_fast(e[2] + g_n + r * hb[2]);
            hown = (1.0f - z) * n + z * hown;
            ll_store(p.ll_h + par * AR_H + gu, hown, tag);
        }
        // gather h_t: lane l takes columns {64k + 2l, 64k + 2l + 1}, k = 0..13 (512 B contiguous per warp load)
        float hv[28];
        if (!ll_poll_pairs<14>(p.ll_h + par * AR_H + 2 * lane, 32, tag, hv)) { failed = true; break; }
#pragma unroll
        for (int k = 0; k < 14; ++k)
            *reinterpret_cast<float2*>(&hs[par][64 * k + 2 * lane]) = make_float2(hv[2 * k], hv[2 * k + 1]);
        bar_arrive(1, AR_THREADS);                   // W_hh warps may start on h_t

        // fc1 rows
        float r1[AR_R];
#pragma unroll
        for (int r = 0; r < AR_R; ++r) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int k = 0; k < 14; ++k) { a = fmaf(w1[r][2 * k], hv[2 * k], a); b = fmaf(w1[r][2 * k + 1], hv[2 * k + 1], b); }
            r1[r] = fmaxf(warp_sum(a + b) + b1[r], 0.f);
        }
        if (lane < AR_R) ll_store(p.ll_r + par * AR_FC + cta * AR_R + lane, lane == 0 ? r1[0] : r1[1], tag);
        // gather relu(fc1): lane l takes columns 8l..8l+7
        float rv[8];
        if (!ll_poll_pairs<4>(p.ll_r + par * AR_FC + 8 * lane, 1, tag, rv)) { failed = true; break; }
        float o2[AR_R];
#pragma unroll
        for (int r = 0; r < AR_R; ++r) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) a = fmaf(w2[r][k], rv[k], a);
            o2[r] = warp_sum(a) + b2[r];
        }
        if (p.out_logits != nullptr && lane < AR_R)
            p.out_logits[static_cast<int64_t>(t) * AR_Q + cta * AR_R + lane] = lane == 0 ? o2[0] : o2[1];

        if (!teacher) {
            if (lane < AR_R) ll_store(p.ll_o + par * AR_Q + cta * AR_R + lane, lane == 0 ? o2[0] : o2[1], tag);
            float ov[8];
            if (!ll_poll_pairs<4>(p.ll_o + par * AR_Q + 8 * lane, 1, tag, ov)) { failed = true; break; }
            // softmax + inverse-CDF sample over classes in index order (lane l holds classes 8l..8l+7)
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
            const float S = __shfl_sync(0xffffffffu, incl, 31);
            const float thr = u_t * S;
            int loc = 8;
#pragma unroll
            for (int k = 7; k >= 0; --k) if (excl + c[k] > thr) loc = k;
            const unsigned hit = __ballot_sync(0xffffffffu, loc < 8);
            if (hit == 0u) x = AR_Q - 1;
            else {
                const int src = __ffs(hit) - 1;
                x = 8 * src + __shfl_sync(0xffffffffu, loc, src);
            }
            if (cta == 0 && lane == 0) {
                if (p.out_wav) p.out_wav[t] = __ldg(p.lut + x);
                if (p.out_codes) p.out_codes[t] = x;
            }
        }
        bar_sync(2, AR_THREADS);                     // W_hh h_t + b_hh is in hhres[par]
    }
    if (failed) {
        abort_flag = 1;
        if (lane == 0) atomicExch(p.status, VQCPC_ERR_TIMEOUT);
        __threadfence_block();
        bar_arrive(1, AR_THREADS);                   // release the W_hh warps so the CTA can exit
    }
}

// ------------------------------------------------------------------------------------------------ host
// workspace layout: [header][u B*2Tc*128][xproj B*2Tc*768][p0 B*2Tc*256][p1 B*2Tc*256][ll words]
static size_t vocoder_ws_bytes(int B, int Tc) {
    const size_t rows = static_cast<size_t>(B) * 2 * Tc;
    return sizeof(WorkspaceHeader) + align_up(rows * 128 * 4, 256) + align_up(rows * 768 * 4, 256) +
           2 * align_up(rows * 256 * 4, 256) + align_up(sizeof(ll_word) * AR_LL_WORDS, 256);
}
static size_t ar_ws_bytes() { return sizeof(WorkspaceHeader) + align_up(sizeof(ll_word) * AR_LL_WORDS, 256); }

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

int vocoder_condition(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker, int B, int Tc,
                      void* ws, size_t ws_bytes, float* out_G, float* out_p, cudaStream_t stream) {
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
        embed_concat_kernel<<<grid, 256, 0, stream>>>(w->code_emb, w->spk_emb, codes, speaker, u, B, Tc, w->dim_code,
                                                      w->dim_speaker);
        VQ_CUDA(cudaGetLastError());
    }
    // layer 0
    if ((rc = gemm_dense(u, 128, w->pre_w_ih[0], 128, w->pre_b_ih[0], xproj, 768, rows, 768, 128, stream))) return rc;
    bigru_layer_kernel<<<dim3(2, B), PRE_G, 0, stream>>>(xproj, w->pre_w_hh[0], w->pre_b_hh[0], p0, T2);
    VQ_CUDA(cudaGetLastError());
    // layer 1 (input = [fwd;bwd] of layer 0)
    if ((rc = gemm_dense(p0, 256, w->pre_w_ih[1], 256, w->pre_b_ih[1], xproj, 768, rows, 768, 256, stream))) return rc;
    bigru_layer_kernel<<<dim3(2, B), PRE_G, 0, stream>>>(xproj, w->pre_w_hh[1], w->pre_b_hh[1], p1, T2);
    VQ_CUDA(cudaGetLastError());
    // hoisted conditioning half of the AR input projection: G = p . W_ih[:, 256:]^T + b_ih
    return gemm_dense(p1, 256, w->ar_w_ih + AR_EMB, AR_EMB + AR_COND, w->ar_b_ih, out_G, AR_G, rows, AR_G, AR_COND,
                      stream);
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
    VQ_CUDA(cudaMemsetAsync(hdr, 0, sizeof(WorkspaceHeader), stream));
    for (int b = 0; b < B; ++b) {
        VQ_CUDA(cudaMemsetAsync(ll, 0, sizeof(ll_word) * AR_LL_WORDS, stream));
        ArParams p{};
        p.w_hh = w->ar_w_hh; p.b_hh = w->ar_b_hh;
        p.fc1_w = w->fc1_w; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b;
        p.eprime = w->eprime;
        p.G = G + static_cast<int64_t>(b) * T2 * AR_G;
        p.uniforms = uniforms ? uniforms + static_cast<int64_t>(b) * L : nullptr;
        p.x_in = x_in ? x_in + static_cast<int64_t>(b) * L : nullptr;
        p.lut = w->mulaw_lut;
        p.out_wav = out_wav ? out_wav + static_cast<int64_t>(b) * L : nullptr;
        p.out_codes = out_codes ? out_codes + static_cast<int64_t>(b) * L : nullptr;
        p.out_logits = out_logits ? out_logits + static_cast<int64_t>(b) * L * AR_Q : nullptr;
        p.ll_h = ll; p.ll_r = ll + 2 * AR_H; p.ll_o = ll + 2 * AR_H + 2 * AR_FC;
        p.status = &hdr->status;
        p.L = L; p.upsample = w->upsample_t;
        void* args[] = {&p};
        VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ar_kernel), dim3(AR_CTAS), dim3(AR_THREADS), args,
                                            0, stream));
    }
    return VQCPC_OK;
}

}  // namespace vqcpc

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int vqcpc_vocoder_pack(const vqcpc_vocoder_weights* w, float* eprime_out, void* stream) {
    return vqcpc::vocoder_pack(w, eprime_out, static_cast<cudaStream_t>(stream));
}
extern "C" size_t vqcpc_vocoder_workspace_bytes(int32_t B, int32_t Tc) { return vqcpc::vocoder_ws_bytes(B, Tc); }
extern "C" int vqcpc_vocoder_condition(const vqcpc_vocoder_weights* w, const int64_t* codes, const int64_t* speaker,
                                       int32_t B, int32_t Tc, void* workspace, size_t workspace_bytes, float* out_G,
                                       float* out_p, void* stream) {
    return vqcpc::vocoder_condition(w, codes, speaker, B, Tc, workspace, workspace_bytes, out_G, out_p,
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
