// Batched sample loop: up to 64 utterances advance together through one persistent cooperative kernel.
//
// The single-utterance kernel (vocoder.cu) is bound by three ~1 us grid-wide exchanges per step.  With a batch the
// same exchanges carry 64 utterances, so the cost per sample falls by the batch size and the fp32 FMA work
// (2.7 M MAC x 64 per step) becomes the other half of the step.  Layout:
//   * 128 CTAs x 256 threads.  CTA j owns hidden units 7j..7j+6 (21 W_hh rows), fc1 rows 2j,2j+1, fc2 rows 2j,2j+1.
//     Its weight slice lives in SHARED memory as float4 groups of 4 consecutive columns; a warp reads a weight
//     group with one broadcast LDS.128 and every LANE multiplies it with the h values of ITS OWN utterance
//     (utterance slots lane and lane + 32), so there are no cross-lane reductions at all.
//   * warp cc streams 112 of the 896 columns for all 64 utterances (two per lane: one weight LDS.128 feeds 8
//     FMAs, keeping the shared-memory return path below the FMA pipe); the eight chunk partials of a
//     (row, utterance) meet in shared memory.
//   * h_t, relu(fc1 h_t), logits and the sampled codes are exchanged through plain global buffers stored
//     utterance-minor ([row][64]: coalesced for producers and consumers) and separated by grid barriers
//     (__threadfence + one LL flag per CTA in its own 128-byte slot, polled by one warp): 4 barriers per step
//     in generate mode, 2 in teacher-forced mode.  h_t is streamed from L2 (229 KB per CTA per step), never
//     staged: every lane consumes its own utterance's column values straight from coalesced 128-byte loads.
//   * step phases:  G  gates -> h_t                                  | barrier |
//                   P2 W_hh h_t (for the next step) + fc1 rows -> r  | barrier |
//                   P3 fc2 rows -> logits                            | barrier |
//                   P4 CTA b samples utterance b (softmax + inverse CDF), writes code + waveform | barrier |
// Arithmetic is fp32 FMA throughout (same definitions as vocoder.cu; summation order differs).
#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

constexpr int AB_H = 896, AB_G = 2688, AB_FC = 256, AB_Q = 256;
constexpr int AB_CTAS = 128, AB_U = 7, AB_R = 2, AB_NROW = 21, AB_ROWS = AB_NROW + AB_R;   // 23 rows ride the h stream
constexpr int AB_B = 64;                          // utterance slots per launch
constexpr int AB_THREADS = 256;
constexpr int AB_CC = 8, AB_CHUNK = AB_H / AB_CC; // 112 columns per warp in the h stream
constexpr int AB_X_INIT = 128;

struct AbParams {
    const float* w_hh; const float* b_hh; const float* fc1_w; const float* fc1_b; const float* fc2_w; const float* fc2_b;
    const float* eprime; const float* lut;
    const float* G;          // (B, T2, 2688)
    const float* uniforms;   // (B, L)   generate mode
    const int64_t* x_in;     // (B, L)   teacher-forced mode
    float* out_wav; int32_t* out_codes; float* out_logits;
    float* hT;               // [2][896][64]  double buffered by step parity
    float* rT;               // [256][64]
    float* oT;               // [256][64]
    int* xs;                 // [64]
    ll_word* flags;          // [128][16]
    int* status;
    long long g_stride;
    int L, upsample, nb;     // nb = active utterances (<= 64)
    long long* trace;        // optional (debug): [n][8] clock64 phase stamps of CTA trace_cta
    int trace_cta, trace_t0, trace_n;
};

// dynamic shared memory layout (floats)
constexpr int AB_WS = (AB_H / 4) * AB_ROWS * 4;     // weight groups: [224][23][4]
constexpr int AB_W2S = (AB_FC / 4) * AB_R * 4;      // fc2 groups:    [64][2][4]
constexpr int AB_ES = AB_Q * AB_NROW;
constexpr int AB_PART = AB_CC * AB_ROWS * AB_B;
constexpr int AB_HH = AB_NROW * AB_B;
constexpr int AB_HOWN = AB_U * AB_B;
constexpr int AB_GC = AB_NROW * AB_B;
constexpr size_t AB_SMEM = sizeof(float) * (AB_WS + AB_W2S + AB_ES + AB_PART + AB_HH + AB_HOWN + AB_GC) + sizeof(int) * AB_B;

__device__ __forceinline__ float ld_strong(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// grid barrier: every CTA publishes `tag` in its own 128-byte slot after a fence; warp 0 polls all 128 slots.
// Returns false (CTA-uniform) on timeout.
__device__ __forceinline__ bool ab_grid_sync(ll_word* flags, uint32_t tag, volatile int* abort_flag, int* status) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ll_store(flags + blockIdx.x * 16, 0.f, tag);
    }
    if (threadIdx.x < 32) {
        const long long t0 = clock64();
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t seen = ll_tag(ll_load(flags + (32 * k + threadIdx.x) * 16));
                ok = ok && (static_cast<int32_t>(seen - tag) >= 0);
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - t0 > LL_TIMEOUT_CYCLES) {
                *abort_flag = 1;
                if (threadIdx.x == 0) atomicExch(status, VQCPC_ERR_TIMEOUT);
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
    return *abort_flag == 0;
}

// split barrier: signal = publish this CTA's arrival (after a fence), wait = poll all 128 arrivals.  Work placed
// between the two hides the ~4000-cycle barrier latency.
__device__ __forceinline__ void ab_signal(ll_word* flags, uint32_t tag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        ll_store(flags + blockIdx.x * 16, 0.f, tag);
    }
}
__device__ __forceinline__ bool ab_wait(ll_word* flags, uint32_t tag, volatile int* abort_flag, int* status) {
    if (threadIdx.x < 32) {
        const long long t0 = clock64();
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t seen = ll_tag(ll_load(flags + (32 * k + threadIdx.x) * 16));
                ok = ok && (static_cast<int32_t>(seen - tag) >= 0);
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (clock64() - t0 > LL_TIMEOUT_CYCLES) {
                *abort_flag = 1;
                if (threadIdx.x == 0) atomicExch(status, VQCPC_ERR_TIMEOUT);
                break;
            }
        }
        __threadfence();
    }
    __syncthreads();
    return *abort_flag == 0;
}

// Stream column groups [c4_begin, c4_end) of this warp's 112-column chunk of h_t from L2 and accumulate rows
// R0 .. R0+NR-1 of the CTA's weight slice for this lane's two utterances (slots lane, lane + 32).
// hcol -> hT[chunk column 0][lane]; wg -> weight groups of the chunk ([c4][23 rows] float4, shared memory).
template <int R0, int NR, int PF>
__device__ __forceinline__ void ab_stream(const float* hcol, const float4* wg, int c4_begin, int c4_end, float (&acc0)[NR],
                                          float (&acc1)[NR]) {
    float ha[PF][4], hb[PF][4];
#pragma unroll
    for (int s = 0; s < PF - 1; ++s) {
        if (c4_begin + s < c4_end) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ha[s][i] = ld_strong(hcol + (4 * (c4_begin + s) + i) * AB_B);
                hb[s][i] = ld_strong(hcol + (4 * (c4_begin + s) + i) * AB_B + 32);
            }
        }
    }
    for (int base = c4_begin; base < c4_end; base += PF) {
#pragma unroll
        for (int s = 0; s < PF; ++s) {
            const int c4 = base + s;
            if (c4 < c4_end) {
                constexpr int dummy = 0; (void)dummy;
                const int nxt = (s + PF - 1) % PF;
                if (c4 + PF - 1 < c4_end) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ha[nxt][i] = ld_strong(hcol + (4 * (c4 + PF - 1) + i) * AB_B);
                        hb[nxt][i] = ld_strong(hcol + (4 * (c4 + PF - 1) + i) * AB_B + 32);
                    }
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const float4 w4 = wg[c4 * AB_ROWS + R0 + r];                    // broadcast LDS.128
                    acc0[r] = fmaf(w4.x, ha[s][0], acc0[r]); acc1[r] = fmaf(w4.x, hb[s][0], acc1[r]);
                    acc0[r] = fmaf(w4.y, ha[s][1], acc0[r]); acc1[r] = fmaf(w4.y, hb[s][1], acc1[r]);
                    acc0[r] = fmaf(w4.z, ha[s][2], acc0[r]); acc1[r] = fmaf(w4.z, hb[s][2], acc1[r]);
                    acc0[r] = fmaf(w4.w, ha[s][3], acc0[r]); acc1[r] = fmaf(w4.w, hb[s][3], acc1[r]);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(AB_THREADS, 1) ar_batch_kernel(AbParams p) {
    extern __shared__ __align__(16) float ab_smem[];
    float* Ws = ab_smem;                 // [c4][row][4]
    float* W2s = Ws + AB_WS;             // [c4][r][4]
    float* Es = W2s + AB_W2S;            // [x][21]
    float* part = Es + AB_ES;            // [cc][row][b]
    float* hh = part + AB_PART;          // [row][b]   W_hh h + b_hh
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

    // ---- one-time staging of the CTA's weight slice
    for (int i = tid; i < (AB_H / 4) * AB_ROWS; i += AB_THREADS) {
        const int c4 = i / AB_ROWS, row = i % AB_ROWS;
        const float* src = row < AB_NROW
            ? p.w_hh + static_cast<int64_t>((row % 3) * AB_H + cta * AB_U + row / 3) * AB_H + 4 * c4
            : p.fc1_w + static_cast<int64_t>(cta * AB_R + (row - AB_NROW)) * AB_H + 4 * c4;
        reinterpret_cast<float4*>(Ws)[i] = __ldg(reinterpret_cast<const float4*>(src));
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
    constexpr int NIT = AB_CHUNK / 4, S1 = 10, S2 = 19;          // column groups per warp; slices of the W_hh rows
    const float4* wg = reinterpret_cast<const float4*>(Ws) + static_cast<int64_t>(cc * NIT) * AB_ROWS;
    for (int t = 0; t < L; ++t) {
        AB_TRACE(0)
        float* hT = p.hT + static_cast<int64_t>(t & 1) * AB_H * AB_B;             // double buffered by step parity
        const float* hcol = hT + static_cast<int64_t>(cc * AB_CHUNK) * AB_B + lane;
        // ------------------------------------------------------------------ G: conditioning reload, gates, h_t
        if (frame_left == 0) {
            for (int i = tid; i < AB_GC; i += AB_THREADS) {
                const int row = i / AB_B, b = i % AB_B;
                Gc[i] = (b < nb) ? __ldg(p.G + b * p.g_stride + static_cast<int64_t>(frame) * AB_G + (row % 3) * AB_H + cta * AB_U + row / 3) : 0.f;
            }
            frame_left = p.upsample; ++frame;
        }
        --frame_left;
        if (teacher && tid < AB_B) xcur[tid] = (tid < nb) ? (static_cast<int>(__ldg(p.x_in + static_cast<int64_t>(tid) * L + t)) & (AB_Q - 1)) : 0;
        __syncthreads();
        for (int i = tid; i < AB_U * AB_B; i += AB_THREADS) {
            const int u = i / AB_B, b = i % AB_B;
            const float* e = &Es[xcur[b] * AB_NROW + 3 * u];
            const float r = sigmoid_fast(__fadd_rn(__fadd_rn(e[0], Gc[(3 * u) * AB_B + b]), hh[(3 * u) * AB_B + b]));
            const float z = sigmoid_fast(__fadd_rn(__fadd_rn(e[1], Gc[(3 * u + 1) * AB_B + b]), hh[(3 * u + 1) * AB_B + b]));
            const float n = tanh_fast(__fmaf_rn(r, hh[(3 * u + 2) * AB_B + b], __fadd_rn(e[2], Gc[(3 * u + 2) * AB_B + b])));
            const float hn = __fmaf_rn(z, __fsub_rn(hown[i], n), n);
            hown[i] = hn;
            hT[(cta * AB_U + u) * AB_B + b] = hn;
        }
        AB_TRACE(1)
        ab_signal(p.flags, ++tag);
        if (!ab_wait(p.flags, tag, &abort_flag, p.status)) return;                  // barrier 1: h_t complete
        AB_TRACE(2)

        // ------------------------------------------------------------------ P2a: the two fc1 rows first (critical path)
        {
            float f0[AB_R] = {0.f, 0.f}, f1[AB_R] = {0.f, 0.f};
            ab_stream<AB_NROW, AB_R, 7>(hcol, wg, 0, NIT, f0, f1);
#pragma unroll
            for (int r = 0; r < AB_R; ++r) {
                part[(cc * AB_R + r) * AB_B + lane] = f0[r];
                part[(cc * AB_R + r) * AB_B + lane + 32] = f1[r];
            }
        }
        __syncthreads();
        if (tid < AB_R * AB_B) {
            const int r = tid / AB_B, b = tid % AB_B;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < AB_CC; ++c) sum += part[(c * AB_R + r) * AB_B + b];
            p.rT[(cta * AB_R + r) * AB_B + b] = fmaxf(sum + b1_s[r], 0.f);
        }
        AB_TRACE(3)
        ab_signal(p.flags, ++tag);                                                  // barrier 2 (r complete) ...
        // ------------------------------------------------------------------ P2b: W_hh rows for the NEXT step, in three
        // slices that run while barriers 2, 3 and 4 are in flight
        float acc0[AB_NROW], acc1[AB_NROW];
#pragma unroll
        for (int r = 0; r < AB_NROW; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
        ab_stream<0, AB_NROW, 3>(hcol, wg, 0, S1, acc0, acc1);
        if (!ab_wait(p.flags, tag, &abort_flag, p.status)) return;                  // ... barrier 2 wait
        AB_TRACE(4)

        // ------------------------------------------------------------------ P3: fc2 rows over relu(fc1 h_t)
        {
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
            __syncthreads();                                                        // fc1 partials in `part` fully consumed
            part[(cc3 * AB_R + 0) * AB_B + slot] = a0;
            part[(cc3 * AB_R + 1) * AB_B + slot] = a1;
        }
        __syncthreads();
        if (tid < AB_R * AB_B) {
            const int r = tid / AB_B, b = tid % AB_B;
            const float o = (part[(0 * AB_R + r) * AB_B + b] + part[(1 * AB_R + r) * AB_B + b]) +
                            (part[(2 * AB_R + r) * AB_B + b] + part[(3 * AB_R + r) * AB_B + b]) + b2_s[r];
            p.oT[(cta * AB_R + r) * AB_B + b] = o;
            if (p.out_logits != nullptr && b < nb)
                p.out_logits[(static_cast<int64_t>(b) * L + t) * AB_Q + cta * AB_R + r] = o;
        }
        AB_TRACE(5)
        if (!teacher) {
            ab_signal(p.flags, ++tag);                                              // barrier 3 (logits complete) ...
            ab_stream<0, AB_NROW, 3>(hcol, wg, S1, S2, acc0, acc1);
            if (!ab_wait(p.flags, tag, &abort_flag, p.status)) return;
            AB_TRACE(6)
            // -------------------------------------------------------------- P4: CTA b samples utterance b
            if (warp == 0 && cta < nb) {
                const int b = cta;
                float ov[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) ov[k] = ld_strong(p.oT + (8 * lane + k) * AB_B + b);
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
                const float thr = __ldg(p.uniforms + static_cast<int64_t>(b) * L + t) * S;
                int loc = 8;
#pragma unroll
                for (int k = 7; k >= 0; --k) if (excl + c[k] > thr) loc = k;
                const unsigned hit = __ballot_sync(0xffffffffu, loc < 8);
                int x = AB_Q - 1;
                if (hit != 0u) {
                    const int src = __ffs(hit) - 1;
                    x = 8 * src + __shfl_sync(0xffffffffu, loc, src);
                }
                if (lane == 0) {
                    p.xs[b] = x;
                    if (p.out_wav) p.out_wav[static_cast<int64_t>(b) * L + t] = __ldg(p.lut + x);
                    if (p.out_codes) p.out_codes[static_cast<int64_t>(b) * L + t] = x;
                }
            }
            AB_TRACE(7)
            ab_signal(p.flags, ++tag);                                              // barrier 4 (codes complete) ...
            ab_stream<0, AB_NROW, 3>(hcol, wg, S2, NIT, acc0, acc1);
        } else {
            ab_stream<0, AB_NROW, 3>(hcol, wg, S1, NIT, acc0, acc1);
            __syncthreads();                                                        // fc2 partials in `part` consumed
        }
        // W_hh partials of the eight column chunks meet in shared memory
#pragma unroll
        for (int r = 0; r < AB_NROW; ++r) {
            part[(cc * AB_ROWS + r) * AB_B + lane] = acc0[r];
            part[(cc * AB_ROWS + r) * AB_B + lane + 32] = acc1[r];
        }
        if (!teacher) {
            if (!ab_wait(p.flags, tag, &abort_flag, p.status)) return;              // ... barrier 4 wait (ends in __syncthreads)
            if (tid < AB_B) {
                int xv = 0;
                if (tid < nb) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(xv) : "l"(p.xs + tid) : "memory");
                xcur[tid] = xv;
            }
        } else {
            __syncthreads();
        }
        for (int i = tid; i < AB_NROW * AB_B; i += AB_THREADS) {
            const int row = i / AB_B, b = i % AB_B;
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < AB_CC; ++c) sum += part[(c * AB_ROWS + row) * AB_B + b];
            hh[i] = sum + bhh_s[row];
        }
        // (the __syncthreads at the top of the next step orders hh / xcur before the gates)
    }
#undef AB_TRACE
}

// workspace of one launch: [hT 2x896x64][rT 256x64][oT 256x64][xs 64][pad][flags 128x16 words]
static size_t ab_ws_bytes() {
    return sizeof(float) * (2 * AB_H + 2 * AB_FC) * AB_B + 256 + sizeof(ll_word) * AB_CTAS * 16;
}
size_t ar_batch_workspace_bytes() { return align_up(ab_ws_bytes(), 256); }

long long* g_ab_trace = nullptr;
int g_ab_trace_cta = 0, g_ab_trace_t0 = 0, g_ab_trace_n = 0;

int ar_batch_run(const vqcpc_vocoder_weights* w, const float* G, const float* uniforms, const int64_t* x_in, int B, int T2,
                 int L, void* ws, int* status, float* out_wav, int32_t* out_codes, float* out_logits, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        VQ_CUDA(cudaFuncSetAttribute(ar_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(AB_SMEM)));
        attr_set = true;
    }
    unsigned char* base = static_cast<unsigned char*>(ws);
    for (int b0 = 0; b0 < B; b0 += AB_B) {
        const int nb = B - b0 < AB_B ? B - b0 : AB_B;
        VQ_CUDA(cudaMemsetAsync(ws, 0, ab_ws_bytes(), stream));
        AbParams p{};
        p.w_hh = w->ar_w_hh; p.b_hh = w->ar_b_hh; p.fc1_w = w->fc1_w; p.fc1_b = w->fc1_b; p.fc2_w = w->fc2_w; p.fc2_b = w->fc2_b;
        p.eprime = w->eprime; p.lut = w->mulaw_lut;
        p.G = G + static_cast<int64_t>(b0) * T2 * AB_G;
        p.g_stride = static_cast<long long>(T2) * AB_G;
        p.uniforms = uniforms ? uniforms + static_cast<int64_t>(b0) * L : nullptr;
        p.x_in = x_in ? x_in + static_cast<int64_t>(b0) * L : nullptr;
        p.out_wav = out_wav ? out_wav + static_cast<int64_t>(b0) * L : nullptr;
        p.out_codes = out_codes ? out_codes + static_cast<int64_t>(b0) * L : nullptr;
        p.out_logits = out_logits ? out_logits + static_cast<int64_t>(b0) * L * AB_Q : nullptr;
        p.hT = reinterpret_cast<float*>(base);
        p.rT = p.hT + 2 * AB_H * AB_B;
        p.oT = p.rT + AB_FC * AB_B;
        p.xs = reinterpret_cast<int*>(p.oT + AB_FC * AB_B);
        p.flags = reinterpret_cast<ll_word*>(base + sizeof(float) * (2 * AB_H + 2 * AB_FC) * AB_B + 256);
        p.status = status;
        p.L = L; p.upsample = w->upsample_t; p.nb = nb;
        p.trace = g_ab_trace; p.trace_cta = g_ab_trace_cta; p.trace_t0 = g_ab_trace_t0; p.trace_n = g_ab_trace_n;
        void* args[] = {&p};
        VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ar_batch_kernel), dim3(AB_CTAS), dim3(AB_THREADS), args,
                                            AB_SMEM, stream));
        count_launch(1);
    }
    return VQCPC_OK;
}

}  // namespace vqcpc
