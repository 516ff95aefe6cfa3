// Encoder.encode of /root/reference/model.py:59-70 on sm_100a, fp32 parity path:
//   conv (implicit-im2col GEMM) -> LN+ReLU -> 4x[Linear -> LN+ReLU] -> Linear(C,64)+b -> VQ lookup -> LSTM.
// This file: LayerNorm+ReLU rows, the exact fp32 VQ nearest-code search + gather, the persistent LSTM
// (LL exchange between 32-CTA groups, W_hh resident in registers) and the host-side orchestration.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

// Code indices come from the caller in the public LSTM entry: gathers clamp them (memory safety) and
// check_codes_kernel reports an out-of-range index through the workspace status word (vqcpc_check_status).
__device__ __forceinline__ int64_t clamp_code(int64_t id) { return id < 0 ? 0 : (id > 511 ? 511 : id); }
__global__ void check_codes_kernel(const int64_t* __restrict__ idx, int64_t n, int* status) {
    bool bad = false;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        bad |= (idx[i] < 0 || idx[i] > 511);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(status, VQCPC_ERR_ARG);
}


// ------------------------------------------------------------------------------------------------
// relu(LayerNorm(x)) in place, one warp per row  (nn.LayerNorm(C) + nn.ReLU, model.py:47-48,51-52).
// mean, then biased variance of (x - mean), eps = 1e-5: the oracle's two-pass formula.
// ------------------------------------------------------------------------------------------------
template <int N4>
__global__ void __launch_bounds__(256) ln_relu_kernel(float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ b, int64_t rows) {
    constexpr int C = N4 * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_sync();
    if (row >= rows) return;
    float4* p = reinterpret_cast<float4*>(x + row * C);
    float4 v[N4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N4; ++j) {
        v[j] = p[j * 32 + lane];
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < N4; ++j) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int j = 0; j < N4; ++j) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + j * 32 + lane);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + j * 32 + lane);
        float4 o;
        o.x = fmaxf(fmaf(v[j].x * rstd, ww.x, bb.x), 0.f);
        o.y = fmaxf(fmaf(v[j].y * rstd, ww.y, bb.y), 0.f);
        o.z = fmaxf(fmaf(v[j].z * rstd, ww.z, bb.z), 0.f);
        o.w = fmaxf(fmaf(v[j].w * rstd, ww.w, bb.w), 0.f);
        p[j * 32 + lane] = o;
    }
}

int layernorm_relu(float* x, const float* w, const float* b, int64_t rows, int C, cudaStream_t stream, bool pdl) {
    if (rows == 0) return VQCPC_OK;
    VQ_ARG(x && w && b, "layernorm: null pointer");
    VQ_ARG(C % 128 == 0 && C >= 128 && C <= 1024, "layernorm: C=%d must be a multiple of 128 in [128,1024]", C);
    if (rows == 0) return VQCPC_OK;
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    switch (C / 128) {
#define LN_CASE(n) case n: VQ_CUDA(launch_pdl(pdl, ln_relu_kernel<n>, dim3(grid), dim3(256), 0, stream, x, w, b, rows)); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
    }
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

// ------------------------------------------------------------------------------------------------
// Tensor-core mode helpers: the A operand of the next tcgen05 GEMM is a pair of bf16 planes [hi | lo].
//   ln_relu_split : relu(LayerNorm(y)) of an fp32 row -> planes (and, optionally, the fp32 row itself)
//   im2col_split  : the conv's implicit im2col rows (K = 320) -> planes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store4(float4 o, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(o.x), h1 = __float2bfloat16_rn(o.y), h2 = __float2bfloat16_rn(o.z),
                        h3 = __float2bfloat16_rn(o.w);
    reinterpret_cast<__nv_bfloat162*>(hi)[0] = __halves2bfloat162(h0, h1);
    reinterpret_cast<__nv_bfloat162*>(hi)[1] = __halves2bfloat162(h2, h3);
    reinterpret_cast<__nv_bfloat162*>(lo)[0] = __halves2bfloat162(__float2bfloat16_rn(o.x - __bfloat162float(h0)),
                                                                   __float2bfloat16_rn(o.y - __bfloat162float(h1)));
    reinterpret_cast<__nv_bfloat162*>(lo)[1] = __halves2bfloat162(__float2bfloat16_rn(o.z - __bfloat162float(h2)),
                                                                   __float2bfloat16_rn(o.w - __bfloat162float(h3)));
}

template <int N4>
__global__ void __launch_bounds__(256) ln_relu_split_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, __nv_bfloat16* __restrict__ planes,
                                                            float* __restrict__ out_f32, int64_t rows) {
    constexpr int C = N4 * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* p = reinterpret_cast<const float4*>(x + row * C);
    float4 v[N4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N4; ++j) {
        v[j] = p[j * 32 + lane];
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < N4; ++j) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
    __nv_bfloat16* prow = planes + row * 2 * C;
#pragma unroll
    for (int j = 0; j < N4; ++j) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + j * 32 + lane);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + j * 32 + lane);
        float4 o;
        o.x = fmaxf(fmaf(v[j].x * rstd, ww.x, bb.x), 0.f);
        o.y = fmaxf(fmaf(v[j].y * rstd, ww.y, bb.y), 0.f);
        o.z = fmaxf(fmaf(v[j].z * rstd, ww.z, bb.z), 0.f);
        o.w = fmaxf(fmaf(v[j].w * rstd, ww.w, bb.w), 0.f);
        const int col = (j * 32 + lane) * 4;
        split_store4(o, prow + col, prow + C + col);
        if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32 + row * C)[j * 32 + lane] = o;
    }
}

static int layernorm_relu_split(const float* x, const float* w, const float* b, void* planes, float* out_f32,
                                int64_t rows, int C, cudaStream_t stream) {
    if (rows == 0) return VQCPC_OK;
    VQ_ARG(C % 128 == 0 && C >= 128 && C <= 1024, "layernorm: C=%d must be a multiple of 128 in [128,1024]", C);
    const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
    __nv_bfloat16* pl = static_cast<__nv_bfloat16*>(planes);
    switch (C / 128) {
#define LN_CASE(n) case n: ln_relu_split_kernel<n><<<grid, 256, 0, stream>>>(x, w, b, pl, out_f32, rows); break;
        LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
    }
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

// One CTA per (utterance, 64 output frames): the 80 x 130 window of mel it needs is read coalesced along time into shared memory
// (the channel-major input has its channels T floats apart: one thread per (row, channel) reading its 4 taps straight from global
// memory touched a different 1.2 KB row per lane), then every thread writes the 4 taps mel[b, i, 2t-1 .. 2t+2] (zero padded) of a
// (row, channel) as 8 bytes of the hi and 8 bytes of the lo plane -- consecutive channels, consecutive addresses.
constexpr int IM_TT = 64;                 // output frames per CTA
constexpr int IM_W = 2 * IM_TT + 2;       // input frames they touch
constexpr int IM_PITCH = IM_W + 3;        // 133: odd pitch, channel-strided reads of the write phase are conflict-free
constexpr int IM_CIN = 80;
__global__ void __launch_bounds__(256) im2col_split_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ planes, int T,
                                                           int Tp, int tiles_per_utt) {
    __shared__ float win[IM_CIN * IM_PITCH];
    const int b = blockIdx.x / tiles_per_utt, t0 = (blockIdx.x % tiles_per_utt) * IM_TT;
    const float* src = mel + static_cast<int64_t>(b) * IM_CIN * T;
    for (int idx = threadIdx.x; idx < IM_CIN * IM_W; idx += 256) {
        const int i = idx / IM_W, j = idx - i * IM_W;
        const int tin = 2 * t0 - 1 + j;
        win[i * IM_PITCH + j] = (tin >= 0 && tin < T) ? __ldg(src + static_cast<int64_t>(i) * T + tin) : 0.0f;
    }
    __syncthreads();
    const int K = IM_CIN * 4;
    const int nt = min(IM_TT, Tp - t0);
    for (int idx = threadIdx.x; idx < nt * IM_CIN; idx += 256) {
        const int tt = idx / IM_CIN, i = idx - tt * IM_CIN;
        const float* w = win + i * IM_PITCH + 2 * tt;
        const float4 v = make_float4(w[0], w[1], w[2], w[3]);
        __nv_bfloat16* row = planes + (static_cast<int64_t>(b) * Tp + t0 + tt) * 2 * K;
        split_store4(v, row + 4 * i, row + K + 4 * i);
    }
}

// ------------------------------------------------------------------------------------------------
// VQ lookup (VQEmbeddingEMA.encode, model.py:103-115), exact fp32 path.
// score[n,m] = |e_m|^2 - 2 x_n.e_m, dot accumulated over k = 0..63 in order with FMA; first minimum wins.
// Persistent CTAs; the whole codebook is transposed into shared memory once per CTA (k-major, 128 KB) and
// each 128-frame tile of x is transposed into a 33 KB k-major tile; 8x8 register micro-tiles, four passes
// of 128 codes.  The gather re-reads codebook rows from L2 (256 B contiguous per frame).
// ------------------------------------------------------------------------------------------------
constexpr int VQ_D = 64, VQ_M = 512, VQ_TF = 128, VQ_LDX = VQ_TF + 4;
constexpr size_t VQ_SMEM = sizeof(float) * (VQ_D * VQ_M + VQ_D * VQ_LDX + VQ_M) + sizeof(int) * VQ_TF;

__global__ void __launch_bounds__(256, 1)
vq_lookup_kernel(const float* __restrict__ x, const float* __restrict__ codebook, int64_t n_frames,
                 float* __restrict__ out_q, int64_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Es = reinterpret_cast<float*>(smem_raw);          // [64][512]
    float* Xs = Es + VQ_D * VQ_M;                             // [64][132]
    float* e2 = Xs + VQ_D * VQ_LDX;                           // [512]
    int* best = reinterpret_cast<int*>(e2 + VQ_M);            // [128]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    // codebook -> shared, transposed; |e|^2 in the same k order as the oracle's sum (sequential).
    for (int f = tid; f < VQ_M * (VQ_D / 4); f += 256) {
        const int m = f >> 4, q = f & 15;
        const float4 v = __ldg(reinterpret_cast<const float4*>(codebook + m * VQ_D) + q);
        Es[(4 * q + 0) * VQ_M + m] = v.x;
        Es[(4 * q + 1) * VQ_M + m] = v.y;
        Es[(4 * q + 2) * VQ_M + m] = v.z;
        Es[(4 * q + 3) * VQ_M + m] = v.w;
    }
    __syncthreads();
    for (int m = tid; m < VQ_M; m += 256) {
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < VQ_D; ++k) s = fmaf(Es[k * VQ_M + m], Es[k * VQ_M + m], s);
        e2[m] = s;
    }

    const int64_t n_tiles = (n_frames + VQ_TF - 1) / VQ_TF;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t f0 = tile * VQ_TF;
        __syncthreads();   // previous tile's Xs / best fully consumed (also orders e2 on the first pass)
        {
            const int row = tid & 127, half = tid >> 7;
            const int64_t fr = f0 + row;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = half * 8 + j;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fr < n_frames) v = __ldg(reinterpret_cast<const float4*>(x + fr * VQ_D) + q);
                Xs[(4 * q + 0) * VQ_LDX + row] = v.x;
                Xs[(4 * q + 1) * VQ_LDX + row] = v.y;
                Xs[(4 * q + 2) * VQ_LDX + row] = v.z;
                Xs[(4 * q + 3) * VQ_LDX + row] = v.w;
            }
        }
        __syncthreads();

        float bestv[8];
        int besti[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { bestv[i] = INFINITY; besti[i] = 0; }

        for (int pass = 0; pass < VQ_M / 128; ++pass) {
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            const float* eb = Es + pass * 128;
#pragma unroll 4
            for (int k = 0; k < VQ_D; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&Xs[k * VQ_LDX + ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&Xs[k * VQ_LDX + 64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&eb[k * VQ_M + tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&eb[k * VQ_M + 64 + tx * 4]);
                const float af[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bf[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
            }
            // candidates visited in increasing code index inside a thread; strict < keeps the first minimum
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int m = pass * 128 + (j >> 2) * 64 + tx * 4 + (j & 3);
                const float em = e2[m];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float sc = fmaf(-2.0f, acc[i][j], em);
                    if (sc < bestv[i]) { bestv[i] = sc; besti[i] = m; }
                }
            }
        }
        // combine the 16 tx lanes that share a frame: lexicographic (score, index) minimum
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bestv[i], o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
                if (ov < bestv[i] || (ov == bestv[i] && oi < besti[i])) { bestv[i] = ov; besti[i] = oi; }
            }
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) best[(i >> 2) * 64 + ty * 4 + (i & 3)] = besti[i];
        }
        __syncthreads();
        if (tid < VQ_TF && f0 + tid < n_frames) out_idx[f0 + tid] = best[tid];
        {
            const int q = tid & 15;
#pragma unroll
            for (int it = 0; it < VQ_TF / 16; ++it) {
                const int row = it * 16 + (tid >> 4);
                const int64_t fr = f0 + row;
                if (fr < n_frames) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(codebook + best[row] * VQ_D) + q);
                    reinterpret_cast<float4*>(out_q + fr * VQ_D)[q] = v;
                }
            }
        }
    }
}

// Latency variant for a few frames (configs[0]: one 2 s utterance = 100 frames; the tiled kernel above would run ONE CTA that
// first transposes the whole 128 KB codebook: 39 us).  A CTA takes 8 frames (one per warp, the frame's row in registers) and
// walks the codebook in chunks of 32 codes staged through shared memory with coalesced loads (rows padded to 65 words: lane =
// code reads its row conflict-free); |e|^2 of a chunk is formed once, by warp 0.  Per code exactly the tiled kernel's
// arithmetic -- sequential FMA over k for x.e and for |e|^2, score = fma(-2, x.e, |e|^2), lexicographic (score, index)
// minimum -- so the indices are bit-identical to it.
__global__ void __launch_bounds__(256) vq_lookup_small_kernel(const float* __restrict__ x, const float* __restrict__ codebook,
                                                              int64_t n_frames, float* __restrict__ out_q, int64_t* __restrict__ out_idx) {
    __shared__ float Ec[2][32][VQ_D + 1];
    __shared__ float e2c[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t fr = static_cast<int64_t>(blockIdx.x) * 8 + warp;
    const bool valid = fr < n_frames;
    pdl_sync();
    float xr[VQ_D];
#pragma unroll
    for (int k4 = 0; k4 < VQ_D / 4; ++k4) {
        const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(x + fr * VQ_D) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        xr[4 * k4] = v.x; xr[4 * k4 + 1] = v.y; xr[4 * k4 + 2] = v.z; xr[4 * k4 + 3] = v.w;
    }
    auto stage = [&](int chunk, int buf) {            // 32 codes x 64 floats = 512 float4: two per thread, coalesced
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int f = tid + 256 * r, code = f >> 4, q = f & 15;
            const float4 v = __ldg(reinterpret_cast<const float4*>(codebook + (chunk * 32 + code) * VQ_D) + q);
            Ec[buf][code][4 * q] = v.x; Ec[buf][code][4 * q + 1] = v.y; Ec[buf][code][4 * q + 2] = v.z; Ec[buf][code][4 * q + 3] = v.w;
        }
    };
    float best = INFINITY;
    int besti = 0;                                    // all-NaN scores: index 0, as the large fp32 kernel (never out of range)
    stage(0, 0);
    __syncthreads();
    for (int chunk = 0; chunk < VQ_M / 32; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < VQ_M / 32) stage(chunk + 1, buf ^ 1);
        if (warp == 0) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < VQ_D; ++k) s = fmaf(Ec[buf][lane][k], Ec[buf][lane][k], s);
            e2c[buf][lane] = s;
        }
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < VQ_D; ++k) d = fmaf(xr[k], Ec[buf][lane][k], d);
        __syncthreads();                              // e2c[buf] written, next chunk staged, this chunk's rows read
        const float sc = fmaf(-2.0f, d, e2c[buf][lane]);
        if (sc < best) { best = sc; besti = chunk * 32 + lane; }          // codes ascend per lane: strict < keeps the first minimum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov < best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (!valid) return;
    if (lane == 0) out_idx[fr] = besti;
    if (lane < 16) reinterpret_cast<float4*>(out_q + fr * VQ_D)[lane] = __ldg(reinterpret_cast<const float4*>(codebook + besti * VQ_D) + lane);
}
constexpr int64_t VQ_SMALL_MAX_FRAMES = 1024;

int vq_lookup(const float* x, const float* codebook, int64_t n, int n_codes, int dim, float* q, int64_t* idx,
              cudaStream_t stream, bool pdl) {
    if (n == 0) return VQCPC_OK;
    VQ_ARG(x && codebook && q && idx, "vq_lookup: null pointer");
    VQ_ARG(n_codes == VQ_M && dim == VQ_D, "vq_lookup: only a 512x64 codebook is supported (got %dx%d)", n_codes, dim);
    VQ_ARG(n >= 0, "vq_lookup: negative frame count");
    if (n == 0) return VQCPC_OK;
    if (n <= VQ_SMALL_MAX_FRAMES) {
        VQ_CUDA(launch_pdl(pdl, vq_lookup_small_kernel, dim3(static_cast<unsigned>((n + 7) / 8)), dim3(256), 0, stream, x, codebook, n, q, idx));
        count_launch(1);
        return VQCPC_OK;
    }
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(vq_lookup_kernel), static_cast<int>(VQ_SMEM))) return rc_attr;
    const int64_t n_tiles = (n + VQ_TF - 1) / VQ_TF;
    const int sms = device_sm_count();
    const unsigned grid = static_cast<unsigned>(n_tiles < sms ? n_tiles : sms);
    vq_lookup_kernel<<<grid, 256, VQ_SMEM, stream>>>(x, codebook, n, q, idx);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

int vq_lookup_auto(const float* x, const float* codebook, int64_t n, float* q, int64_t* idx, void* planes_ws, int* err,
                   cudaStream_t stream, bool pdl) {
    if (planes_ws != nullptr && err != nullptr && n >= VQ_TC_MIN_FRAMES)
        return vq_lookup_tc(x, codebook, n, q, idx, planes_ws, err, stream);
    return vq_lookup(x, codebook, n, VQ_M, VQ_D, q, idx, stream, pdl);
}

// ------------------------------------------------------------------------------------------------
// LSTM(64 -> 256) over the quantised codes (model.py:57,69).
//   x-projection: z_q takes only 512 values, so  W_ih z_q[t] + b_ih + b_hh = table[idx[t]]  with
//   table = codebook . W_ih^T + b  (512 x 1024), one small GEMM per call.
//   Recurrence: groups of 32 CTAs; CTA c of a group owns hidden units 8c..8c+7 (warp w <-> unit 8c+w, its 4
//   gate rows of W_hh live in registers, 8 columns per lane).  Each step every CTA publishes its 8 new h
//   values through the LL exchange and polls all 256.  A group advances NB utterances in lockstep to
//   amortise the exchange latency; groups are independent (utterances never interact).
// ------------------------------------------------------------------------------------------------
constexpr int LSTM_H = 256, LSTM_G = 1024, LSTM_GROUP = 32;

struct LstmParams {
    const float* table;     // (512, 1024)
    const int64_t* idx;     // (B, Tp)
    const float* w_hh;      // (1024, 256)
    float* out;             // (B, Tp, 256)
    ll_word* ll;            // [n_groups][2][NB][256]
    int* status;
    int B, Tp, n_groups;
};

template <int NB>
__global__ void __launch_bounds__(256) lstm_kernel(LstmParams p) {
    __shared__ __align__(16) float hs[2][NB][LSTM_H];
    __shared__ int abort_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = blockIdx.x / LSTM_GROUP, cta = blockIdx.x % LSTM_GROUP;
    const int unit = cta * 8 + warp;

    // W_hh rows (gate*256 + unit), columns {4*lane..+3, 128+4*lane..+3}
    float w[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float* row = p.w_hh + static_cast<int64_t>(g * LSTM_H + unit) * LSTM_H;
        const float4 a = __ldg(reinterpret_cast<const float4*>(row) + lane);
        const float4 b = __ldg(reinterpret_cast<const float4*>(row + 128) + lane);
        w[g][0] = a.x; w[g][1] = a.y; w[g][2] = a.z; w[g][3] = a.w;
        w[g][4] = b.x; w[g][5] = b.y; w[g][6] = b.z; w[g][7] = b.w;
    }
    if (tid == 0) abort_flag = 0;

    ll_word* ll = p.ll + static_cast<size_t>(group) * 2 * NB * LSTM_H;
    const int n_chunks = (p.B + NB - 1) / NB;
    uint32_t tag = 0;

    for (int chunk = group; chunk < n_chunks; chunk += p.n_groups) {
        const int b0 = chunk * NB;
        // lane nb of every warp runs the gate math of utterance slot nb
        const int my_b = b0 + lane;
        const bool my_valid = (lane < NB) && (my_b < p.B);
        float cst = 0.f;
        __syncthreads();
        for (int i = tid; i < 2 * NB * LSTM_H; i += 256) (&hs[0][0][0])[i] = 0.f;   // h_{-1} = 0
        __syncthreads();

        float xp[4] = {0.f, 0.f, 0.f, 0.f};
        if (my_valid) {
            const int64_t id = clamp_code(p.idx[static_cast<int64_t>(my_b) * p.Tp]);
            const float* trow = p.table + id * LSTM_G + unit;
#pragma unroll
            for (int g = 0; g < 4; ++g) xp[g] = __ldg(trow + g * LSTM_H);
        }

        for (int t = 0; t < p.Tp; ++t) {
            ++tag;
            // slot parity follows the group-global step counter (NOT t): consecutive steps always alternate,
            // also across utterance chunks with odd Tp
            const int cur = static_cast<int>(tag & 1u), prev = cur ^ 1;
            // prefetch next step's x-projection
            float xn[4] = {0.f, 0.f, 0.f, 0.f};
            if (my_valid && t + 1 < p.Tp) {
                const int64_t id = clamp_code(p.idx[static_cast<int64_t>(my_b) * p.Tp + t + 1]);
                const float* trow = p.table + id * LSTM_G + unit;
#pragma unroll
                for (int g = 0; g < 4; ++g) xn[g] = __ldg(trow + g * LSTM_H);
            }
            float mine[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const float4 h0 = *reinterpret_cast<const float4*>(&hs[prev][nb][4 * lane]);
                const float4 h1 = *reinterpret_cast<const float4*>(&hs[prev][nb][128 + 4 * lane]);
                const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                float s[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) a = fmaf(w[g][k], hv[k], a);
                    s[g] = a;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) s[g] = warp_sum(s[g]);
                if (lane == nb) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) mine[g] = s[g];
                }
            }
            if (lane < NB) {
                // PyTorch gate order i, f, g, o
                const float ig = sigmoid_fast(mine[0] + xp[0]);
                const float fg = sigmoid_fast(mine[1] + xp[1]);
                const float gg = tanh_fast(mine[2] + xp[2]);
                const float og = sigmoid_fast(mine[3] + xp[3]);
                cst = fg * cst + ig * gg;
                const float hn = og * tanh_fast(cst);
                // every producer CTA owns one contiguous slot [NB][8] (64 B at NB = 1): a line is written by one CTA only
                ll_store(ll + ((static_cast<size_t>(cur) * LSTM_GROUP + cta) * NB + lane) * 8 + warp, hn, tag);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) xp[g] = xn[g];

            // gather all 256 hidden values of every slot: thread tid polls column tid
            float got[NB];
            {
                const ll_word* src = ll + ((static_cast<size_t>(cur) * LSTM_GROUP + (tid >> 3)) * NB) * 8 + (tid & 7);
                const long long t0 = clock64();
                while (clock64() - t0 < 300) {}          // the earliest a remote value can be visible (~770-cycle one-way)
                bool done = false;
                while (!done) {
                    ll_word wv[NB];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) wv[nb] = ll_load(src + nb * 8);
                    done = true;
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) {
                        done = done && (ll_tag(wv[nb]) == tag);
                        got[nb] = ll_val(wv[nb]);
                    }
                    if (!done && clock64() - t0 > LL_TIMEOUT_CYCLES) {
                        abort_flag = 1;
                        atomicExch(p.status, VQCPC_ERR_TIMEOUT);
                        break;
                    }
                }
            }
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) hs[cur][nb][tid] = got[nb];
            // CTA nb writes slot nb's output row (coalesced 1 KB)
            if (cta < NB && b0 + cta < p.B) {
                float v = got[0];
#pragma unroll
                for (int nb = 1; nb < NB; ++nb) v = (cta == nb) ? got[nb] : v;
                p.out[(static_cast<int64_t>(b0 + cta) * p.Tp + t) * LSTM_H + tid] = v;
            }
            __syncthreads();
            if (abort_flag) return;
        }
    }
}

static size_t lstm_ll_bytes(int n_groups, int nb) { return sizeof(ll_word) * n_groups * 2 * nb * LSTM_H; }
constexpr int LSTM_MAX_GROUPS = 32;
constexpr int LSTM_MAX_NB = 8;

template <int NB>
static int lstm_launch(LstmParams prm, void* ll_mem, cudaStream_t stream) {
    int per_sm = 0;
    VQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_kernel<NB>, 256, 0));
    const int capacity = per_sm * device_sm_count() / LSTM_GROUP;
    if (capacity < 1) {
        set_error("lstm: device cannot co-schedule one 32-CTA group");
        return VQCPC_ERR_DEVICE;
    }
    const int n_chunks = (prm.B + NB - 1) / NB;
    int n_groups = n_chunks < capacity ? n_chunks : capacity;
    if (n_groups > LSTM_MAX_GROUPS) n_groups = LSTM_MAX_GROUPS;
    prm.n_groups = n_groups;
    prm.ll = static_cast<ll_word*>(ll_mem);
    VQ_CUDA(cudaMemsetAsync(ll_mem, 0, lstm_ll_bytes(n_groups, NB), stream));
    void* args[] = {&prm};
    VQ_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_kernel<NB>), dim3(n_groups * LSTM_GROUP),
                                        dim3(256), args, 0, stream));
    count_launch(1);
    return VQCPC_OK;
}

// ------------------------------------------------------------------------------------------------
// Large batches run the LSTM time-major instead: per step ONE batched product  gates = H_{t-1} . W_hh^T
// (B x 1024, K = 256; tcgen05 over bf16 hi/lo planes of h, or the fp32 GEMM) followed by this fused gate kernel,
// which adds the gathered input projection table[idx[b, t]], applies the gate non-linearities, updates the cell
// state, writes h_t into the output sequence and re-splits it into the planes the next step's GEMM reads.
// ------------------------------------------------------------------------------------------------
constexpr int LSTM_BATCHED_MIN_B = 64;
// the persistent tcgen05 kernel (tensor-core modes) already pays from 33 utterances: 0.70 ms per 150 steps at 17 .. 63 against
// 0.74 .. 0.86 for the L2-exchange groups below at 33 .. 63 (which win below that: 0.47 .. 0.53 at 17 .. 32, 0.28 .. 0.31 at 10 .. 16,
// and stay the path of the fp32 mode)
constexpr int LSTM_PERSIST_MIN_B = 33;
constexpr int LSTM_CLUSTER_MAX_B = 9;              // one wave of 16-CTA clusters on 148 SMs
constexpr int LSTM_FUSED_MAX_B = 2048;            // below: fused tcgen05 step kernel; from here on: GEMM + coalesced gate kernel

__global__ void lstm_gate_kernel(const float* __restrict__ gates, const float* __restrict__ table,
                                 const int64_t* __restrict__ idx, int t, int Tp, float* __restrict__ cstate,
                                 float* __restrict__ out, __nv_bfloat16* __restrict__ hplanes, int B) {
    const int64_t total = static_cast<int64_t>(B) * (LSTM_H / 4);
    // programmatic dependent launch (no-ops otherwise): wait for the step's GEMM, then let the next GEMM start its prologue
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / (LSTM_H / 4)), q = static_cast<int>(i % (LSTM_H / 4));
        const float4* trow = reinterpret_cast<const float4*>(table + clamp_code(idx[static_cast<int64_t>(b) * Tp + t]) * LSTM_G);
        float4 g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            g[k] = __ldg(trow + k * (LSTM_H / 4) + q);
            if (gates != nullptr) {
                const float4 r = reinterpret_cast<const float4*>(gates + static_cast<int64_t>(b) * LSTM_G)[k * (LSTM_H / 4) + q];
                g[k].x += r.x; g[k].y += r.y; g[k].z += r.z; g[k].w += r.w;
            }
        }
        float4 c = (t == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(cstate)[i];
        float4 h;
#define LSTM_CELL(e)                                                                        \
        c.e = sigmoid_fast(g[1].e) * c.e + sigmoid_fast(g[0].e) * tanh_fast(g[2].e);         \
        h.e = sigmoid_fast(g[3].e) * tanh_fast(c.e);
        LSTM_CELL(x) LSTM_CELL(y) LSTM_CELL(z) LSTM_CELL(w)
#undef LSTM_CELL
        reinterpret_cast<float4*>(cstate)[i] = c;
        reinterpret_cast<float4*>(out + (static_cast<int64_t>(b) * Tp + t) * LSTM_H)[q] = h;
        if (hplanes != nullptr) {
            __nv_bfloat16* row = hplanes + static_cast<int64_t>(b) * 2 * LSTM_H;
            split_store4(h, row + 4 * q, row + LSTM_H + 4 * q);
        }
    }
}

static bool lstm_persist_enabled() {          // VQCPC_LSTM_PERSIST=0 selects the per-step-launch paths (A/B measurements)
    static int v = -1;
    if (v < 0) { const char* e = getenv("VQCPC_LSTM_PERSIST"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// workspace: [header][table 512x1024][ll][gates B x 1024][cstate B x 256][h planes B x 512 bf16] x 2
static size_t lstm_batched_bytes(int B) {
    if (B < LSTM_PERSIST_MIN_B) return 0;
    return align_up(sizeof(float) * B * LSTM_G, 256) + align_up(sizeof(float) * B * LSTM_H, 256) +
           2 * align_up(2 * static_cast<size_t>(B) * 2 * LSTM_H, 256) +      // two h-plane buffers (ping-pong across steps)
           align_up(lstm_persist_table_bytes(), 256);                         // permuted input-projection table of lstm_persist
}
static size_t lstm_ws_bytes(int B) {
    return sizeof(WorkspaceHeader) + align_up(sizeof(float) * VQ_M * LSTM_G, 256) +
           lstm_ll_bytes(LSTM_MAX_GROUPS, LSTM_MAX_NB) + lstm_batched_bytes(B);
}

int lstm_forward(const vqcpc_encoder_weights* w, const int64_t* idx, int B, int Tp, void* ws, size_t ws_bytes,
                 float* out_c, cudaStream_t stream, bool reset_status = true, int mode = VQCPC_GEMM_FP32) {
    VQ_ARG(w && idx && ws && out_c, "lstm: null pointer");
    VQ_ARG(w->n_embeddings == VQ_M && w->z_dim == VQ_D && w->c_dim == LSTM_H,
           "lstm: only 512 codes x 64 -> 256 is supported");
    VQ_ARG(ws_bytes >= lstm_ws_bytes(B), "lstm: workspace too small (%zu < %zu)", ws_bytes, lstm_ws_bytes(B));
    if (B == 0 || Tp == 0) return VQCPC_OK;
    unsigned char* base = static_cast<unsigned char*>(ws);
    WorkspaceHeader* hdr = reinterpret_cast<WorkspaceHeader*>(base);
    float* table_ws = reinterpret_cast<float*>(base + sizeof(WorkspaceHeader));
    void* ll = base + sizeof(WorkspaceHeader) + align_up(sizeof(float) * VQ_M * LSTM_G, 256);
    if (reset_status) VQ_CUDA(cudaMemsetAsync(hdr, 0, sizeof(WorkspaceHeader), stream));
    // the 512-row input-projection table: precomputed with the weights (vqcpc_encoder_weights::lstm_table) or built here
    int rc = VQCPC_OK;
    if (w->lstm_table == nullptr)
        rc = gemm_dense(w->codebook, VQ_D, w->lstm_w_ih, VQ_D, w->lstm_b, table_ws, LSTM_G, VQ_M, LSTM_G, VQ_D, stream);
    const float* table = w->lstm_table != nullptr ? w->lstm_table : table_ws;
    if (rc) return rc;
    const bool tc_mode = (mode != VQCPC_GEMM_FP32) && (w->lstm_whh_p != nullptr);
    if (B >= LSTM_BATCHED_MIN_B || (tc_mode && lstm_persist_enabled() && B >= LSTM_PERSIST_MIN_B)) {
        unsigned char* bb = static_cast<unsigned char*>(ll) + lstm_ll_bytes(LSTM_MAX_GROUPS, LSTM_MAX_NB);
        float* gates = reinterpret_cast<float*>(bb); bb += align_up(sizeof(float) * B * LSTM_G, 256);
        float* cstate = reinterpret_cast<float*>(bb); bb += align_up(sizeof(float) * B * LSTM_H, 256);
        __nv_bfloat16* hplanes = reinterpret_cast<__nv_bfloat16*>(bb); bb += align_up(2 * static_cast<size_t>(B) * 2 * LSTM_H, 256);
        __nv_bfloat16* hplanes2 = reinterpret_cast<__nv_bfloat16*>(bb); bb += align_up(2 * static_cast<size_t>(B) * 2 * LSTM_H, 256);
        float* table_perm = reinterpret_cast<float*>(bb);
        const bool tc = tc_mode;                                                    // bf16 mode too: the recurrence stays bf16x3
        const int64_t total = static_cast<int64_t>(B) * (LSTM_H / 4);
        const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        if (tc && lstm_persist_enabled()) {
            // one persistent launch for all T' steps (lstm_persist.cu); the gates buffer doubles as its counter scratch
            return lstm_persist(table, idx, w->lstm_whh_p, B, Tp, hplanes, hplanes2, reinterpret_cast<unsigned*>(gates), table_perm, out_c,
                                &hdr->status, stream);
        }
        if (tc && B >= LSTM_FUSED_MAX_B) {
            // Large batches: the step's product through the plain tcgen05 GEMM, then the fully coalesced gate kernel (the
            // fused epilogue below touches 32-byte pieces per thread, which loses to coalesced traffic once the step is
            // bandwidth- rather than launch-bound: 4096 utterances 14.6 ms unfused vs 16.1 ms fused).
            TcPlan plan;
            if ((rc = gemm_tc_plan(&plan, hplanes, w->lstm_whh_p, nullptr, gates, LSTM_G, B, LSTM_G, LSTM_H, 3, &hdr->status))) return rc;
            for (int t = 0; t < Tp; ++t) {
                if (t > 0 && (rc = gemm_tc_run(&plan, stream, true))) return rc;
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(256);
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr;
                cfg.numAttrs = t > 0 ? 1 : 0;
                const float* gates_in = t > 0 ? gates : nullptr;
                VQ_CUDA(cudaLaunchKernelEx(&cfg, lstm_gate_kernel, gates_in, static_cast<const float*>(table), idx, t, Tp, cstate, out_c,
                                           hplanes, B));
                count_launch(1);
            }
            return VQCPC_OK;
        }
        if (tc) {
            // Small / medium batches (launch- and fill-bound steps): step 0 has h = 0 (gate kernel alone, writes the first h
            // planes); every later step is ONE kernel -- W_hh h_{t-1} on tcgen05 with the LSTM cell fused into its epilogue
            // -- reading one plane buffer and writing the other, each launched as a programmatic dependent of its
            // predecessor (512 utterances x 3 s: 3.54 -> 3.40 ms).
            __nv_bfloat16* pb[2] = {hplanes, hplanes2};
            TcPlan plan[2];
            for (int i = 0; i < 2; ++i)
                if ((rc = gemm_tc_plan_lstm(&plan[i], pb[i], w->lstm_whh_p, B, LSTM_H, &hdr->status))) return rc;
            lstm_gate_kernel<<<grid, 256, 0, stream>>>(nullptr, table, idx, 0, Tp, cstate, out_c, pb[0], B);
            VQ_CUDA(cudaGetLastError());
            count_launch(1);
            for (int t = 1; t < Tp; ++t)
                if ((rc = gemm_tc_run_lstm(&plan[(t - 1) & 1], table, idx, cstate, out_c, pb[t & 1], t, Tp, stream, true))) return rc;
            return VQCPC_OK;
        }
        for (int t = 0; t < Tp; ++t) {
            if (t > 0) {
                rc = gemm_dense(out_c + static_cast<int64_t>(t - 1) * LSTM_H, static_cast<int64_t>(Tp) * LSTM_H, w->lstm_w_hh,
                                LSTM_H, nullptr, gates, LSTM_G, B, LSTM_G, LSTM_H, stream);
                if (rc) return rc;
            }
            lstm_gate_kernel<<<grid, 256, 0, stream>>>(t > 0 ? gates : nullptr, table, idx, t, Tp, cstate, out_c, nullptr, B);
            VQ_CUDA(cudaGetLastError());
            count_launch(1);
        }
        return VQCPC_OK;
    }
    // a few utterances: one 16-CTA cluster each, h over DSMEM (lstm_cluster.cu); VQCPC_LSTM_CLUSTER16=0 keeps the L2 kernel
    static const int use_cluster = [] { const char* e = getenv("VQCPC_LSTM_CLUSTER16"); return (e && e[0] == '0') ? 0 : 1; }();
    if (use_cluster && B <= LSTM_CLUSTER_MAX_B && lstm_cluster_supported())
        return lstm_cluster_launch(table, idx, w->lstm_w_hh, B, Tp, out_c, &hdr->status, stream);
    LstmParams prm{};
    prm.table = table;
    prm.idx = idx;
    prm.w_hh = w->lstm_w_hh;
    prm.out = out_c;
    prm.status = &hdr->status;
    prm.B = B;
    prm.Tp = Tp;
    // lockstep width: latency case keeps one utterance per group, large batches amortise the exchange
    // lockstep width NB (utterances per 32-CTA group), measured on B200 (ms per 150 steps, tools/lstm_time.py with LP_MODE=0):
    //   NB = 1: 10-12 utterances 0.28, 16: 0.63, 24: 1.18, 40: 1.81 (more groups than fit at once run in rounds)
    //   NB = 4: 10-16: 0.31, 17-32: 0.52-0.68, 40: 0.74, 47: 0.81, 63: 0.86        NB = 8: 17-32: 0.47-0.53, 40: 0.81, 63: 0.88
    if (B >= 8 * 12) return lstm_launch<8>(prm, ll, stream);
    if (B >= 33) return lstm_launch<4>(prm, ll, stream);
    if (B >= 17) return lstm_launch<8>(prm, ll, stream);
    if (B >= 13) return lstm_launch<4>(prm, ll, stream);
    return lstm_launch<1>(prm, ll, stream);
}

// ------------------------------------------------------------------------------------------------
// Orchestration of Encoder.encode.
// workspace: [lstm workspace][act0 M*C][act1 M*C][z_pre M*64]
// ------------------------------------------------------------------------------------------------
static size_t encoder_ws_bytes(int B, int T, int C, int mode) {
    const int Tp = T >= 2 ? (T - 2) / 2 + 1 : 0;
    const size_t M = static_cast<size_t>(B) * Tp;
    size_t n = align_up(lstm_ws_bytes(B), 256) + 2 * align_up(M * C * sizeof(float), 256) +
               align_up(M * VQ_D * sizeof(float), 256) + align_up(VQ_TC_PLANES_BYTES, 1024);
    if (mode != VQCPC_GEMM_FP32)          // two bf16 plane buffers (ping-pong A operands) + the fused-LN kernel's scratch rows
        n += 2 * (align_up(M * 2 * (C > 320 ? C : 320) * 2, 256) + 1024) + align_up(gemm_ln_pair_scratch_bytes(C), 256);
    else                                  // split-K planes of the latency case (0 unless M <= 256)
        n += gemm_splitk_ws_bytes(static_cast<int64_t>(M), C);
    return n;
}

int encoder_forward(const vqcpc_encoder_weights* w, const float* mel, int B, int T, void* ws, size_t ws_bytes,
                    float* out_z, float* out_c, int64_t* out_idx, float* out_prevq, float* out_hidden, int mode,
                    cudaStream_t stream) {
    if (B == 0) return VQCPC_OK;
    VQ_ARG(w && mel && ws && out_z && out_idx, "encoder: null pointer");
    VQ_ARG(B >= 0 && T >= 2, "encoder: bad shape B=%d T=%d (T must be >= 2)", B, T);
    VQ_ARG(mode == VQCPC_GEMM_FP32 || mode == VQCPC_GEMM_BF16X3 || mode == VQCPC_GEMM_BF16, "encoder: unknown gemm_mode %d", mode);
    const int C = w->channels;
    VQ_ARG(w->in_channels == 80, "encoder: in_channels must be 80");
    VQ_ARG(C % 128 == 0 && C >= 128 && C <= 1024, "encoder: channels=%d must be a multiple of 128 in [128,1024]", C);
    VQ_ARG(ws_bytes >= encoder_ws_bytes(B, T, C, mode), "encoder: workspace too small");
    const int Tp = (T - 2) / 2 + 1;
    const int64_t M = static_cast<int64_t>(B) * Tp;
    unsigned char* base = static_cast<unsigned char*>(ws);
    void* lstm_ws = base;
    WorkspaceHeader* hdr = reinterpret_cast<WorkspaceHeader*>(base);
    size_t off = align_up(lstm_ws_bytes(B), 256);
    float* act[2];
    act[0] = reinterpret_cast<float*>(base + off); off += align_up(M * C * sizeof(float), 256);
    act[1] = reinterpret_cast<float*>(base + off); off += align_up(M * C * sizeof(float), 256);
    float* zpre_ws = reinterpret_cast<float*>(base + off); off += align_up(M * VQ_D * sizeof(float), 256);
    off = align_up(off, 1024);
    void* vq_planes = base + off; off += align_up(VQ_TC_PLANES_BYTES, 1024);
    float* zpre = out_prevq ? out_prevq : zpre_ws;

    int rc;
    bool vq_pdl = false;
    if (mode == VQCPC_GEMM_FP32) {
        VQ_CUDA(cudaMemsetAsync(hdr, 0, sizeof(WorkspaceHeader), stream));
        // latency case (M <= 256 frames): split-K planes behind the other buffers, tile counters in the header's reserved words
        void* sk = base + off;
        const size_t sk_bytes = gemm_splitk_ws_bytes(M, C);
        unsigned* sk_ctr = reinterpret_cast<unsigned*>(hdr->reserved);
        static_assert(sizeof(hdr->reserved) >= SPLITK_COUNTERS * sizeof(unsigned), "header too small for the split-K counters");
        // one utterance (M <= 256): eleven small kernels in a row -- each is launched as a programmatic dependent of the one before
        // it (launch latency and prologue under the predecessor).  The first one follows the memset and is launched normally.
        const bool pdl = M <= 256;
        if ((rc = gemm_conv(mel, B, T, 80, w->conv_w, act[0], C, stream, sk, sk_bytes, sk_ctr))) return rc;
        if ((rc = layernorm_relu(act[0], w->ln_w[0], w->ln_b[0], M, C, stream, pdl))) return rc;
        int cur = 0;
        for (int j = 0; j < 4; ++j) {
            if ((rc = gemm_dense(act[cur], C, w->fc_w[j], C, nullptr, act[cur ^ 1], C, M, C, C, stream, sk, sk_bytes, sk_ctr, pdl))) return rc;
            cur ^= 1;
            if ((rc = layernorm_relu(act[cur], w->ln_w[j + 1], w->ln_b[j + 1], M, C, stream, pdl))) return rc;
        }
        if (out_hidden)
            VQ_CUDA(cudaMemcpyAsync(out_hidden, act[cur], M * C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        if ((rc = gemm_dense(act[cur], C, w->proj_w, C, w->proj_b, zpre, VQ_D, M, VQ_D, C, stream, sk, sk_bytes, sk_ctr, pdl && !out_hidden))) return rc;
        vq_pdl = pdl;
    } else {
        // tensor-core mode: every GEMM is tcgen05 over bf16 hi/lo planes; LN+ReLU re-splits its output for the next one
        VQ_ARG(w->conv_wp && w->fc_wp[0] && w->fc_wp[1] && w->fc_wp[2] && w->fc_wp[3] && w->proj_wp,
               "encoder: weight planes missing (vqcpc_split_planes) for VQCPC_GEMM_BF16X3");
        VQ_ARG(M < (1LL << 31), "encoder: too many frames for the tensor-core mode");
        const size_t plane_bytes = align_up(M * 2 * (C > 320 ? C : 320) * 2, 256) + 1024;
        void* planes = base + off; off += plane_bytes;
        void* planes2 = base + off; off += plane_bytes;
        void* ln_scratch = base + off;
        const bool fused = gemm_ln_pair_supported(C);
        const int nseg = mode == VQCPC_GEMM_BF16 ? 1 : 3;                 // single-pass bf16: hi planes only
        const int out_pitch = nseg == 3 ? 2 * C : C;
        VQ_CUDA(cudaMemsetAsync(hdr, 0, sizeof(WorkspaceHeader), stream));
        {
            const int tiles_per_utt = (Tp + IM_TT - 1) / IM_TT;
            im2col_split_kernel<<<static_cast<unsigned>(B) * tiles_per_utt, 256, 0, stream>>>(mel, static_cast<__nv_bfloat16*>(planes), T, Tp, tiles_per_utt);
            VQ_CUDA(cudaGetLastError());
            count_launch(1);
        }
        if (fused) {
            // every layer is ONE kernel on CTA pairs (gemm_pair.cu): GEMM + LayerNorm + ReLU, written straight as the next
            // layer's bf16 planes (ping-pong between the two plane buffers)
            void* cur = planes;
            void* nxt = planes2;
            if ((rc = gemm_ln_pair(cur, 640, 320, w->conv_wp, w->ln_w[0], w->ln_b[0], nxt, out_pitch, nullptr, ln_scratch,
                                   static_cast<int>(M), C, 320, nseg, &hdr->status, stream))) return rc;
            for (int j = 0; j < 4; ++j) {
                void* t = cur; cur = nxt; nxt = t;
                if ((rc = gemm_ln_pair(cur, out_pitch, C, w->fc_wp[j], w->ln_w[j + 1], w->ln_b[j + 1], nxt, out_pitch,
                                       j == 3 ? out_hidden : nullptr, ln_scratch, static_cast<int>(M), C, C, nseg, &hdr->status,
                                       stream))) return rc;
            }
            planes = nxt;
            if ((rc = gemm_tc(planes, w->proj_wp, w->proj_b, zpre, VQ_D, static_cast<int>(M), VQ_D, C, nseg, &hdr->status, stream))) return rc;
        } else {
            if ((rc = gemm_tc(planes, w->conv_wp, nullptr, act[0], C, static_cast<int>(M), C, 320, 3, &hdr->status, stream))) return rc;
            if ((rc = layernorm_relu_split(act[0], w->ln_w[0], w->ln_b[0], planes, nullptr, M, C, stream))) return rc;
            for (int j = 0; j < 4; ++j) {
                if ((rc = gemm_tc(planes, w->fc_wp[j], nullptr, act[0], C, static_cast<int>(M), C, C, 3, &hdr->status, stream))) return rc;
                if ((rc = layernorm_relu_split(act[0], w->ln_w[j + 1], w->ln_b[j + 1], planes, j == 3 ? out_hidden : nullptr, M, C,
                                               stream))) return rc;
            }
            if ((rc = gemm_tc(planes, w->proj_wp, w->proj_b, zpre, VQ_D, static_cast<int>(M), VQ_D, C, 3, &hdr->status, stream))) return rc;
        }
    }
    // the nearest-code search is exact in both modes (tensor-core coarse pass + exact recheck, or the fp32 kernel)
    if ((rc = vq_lookup_auto(zpre, w->codebook, M, out_z, out_idx, vq_planes, &hdr->status, stream, vq_pdl))) return rc;
    if (out_c == nullptr) return VQCPC_OK;      // front part only (the caller runs the recurrence later: vqcpc_lstm_forward_ex)
    return lstm_forward(w, out_idx, B, Tp, lstm_ws, lstm_ws_bytes(B), out_c, stream, false, mode);
}

}  // namespace vqcpc

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int vqcpc_layernorm_relu_f32(float* x, const float* w, const float* b, int64_t rows, int32_t C,
                                        void* stream) {
    return vqcpc::layernorm_relu(x, w, b, rows, C, static_cast<cudaStream_t>(stream));
}
extern "C" size_t vqcpc_vq_workspace_bytes(void) { return 1024 + vqcpc::VQ_TC_PLANES_BYTES; }
// workspace layout: [status int | pad to 1 KB | codebook planes].  No hidden state: the tensor-core search (n_frames >= 8192)
// runs only when the caller passes a workspace; its status word is read with vqcpc_check_status(workspace, stream).
extern "C" int vqcpc_vq_lookup(const float* x, const float* codebook, int64_t n_frames, int32_t n_codes, int32_t dim,
                               float* out_q, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace vqcpc;
    if (n_frames == 0) return VQCPC_OK;
    VQ_ARG(n_codes == VQ_M && dim == VQ_D, "vq_lookup: only a 512x64 codebook is supported (got %dx%d)", n_codes, dim);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (workspace == nullptr) return vq_lookup(x, codebook, n_frames, n_codes, dim, out_q, out_idx, st);
    VQ_ARG(workspace_bytes >= vqcpc_vq_workspace_bytes(), "vq_lookup: workspace too small (%zu < %zu)", workspace_bytes,
           vqcpc_vq_workspace_bytes());
    unsigned char* s = static_cast<unsigned char*>(workspace);
    VQ_CUDA(cudaMemsetAsync(s, 0, 2 * sizeof(int), st));        // WorkspaceHeader: status, index_error
    return vq_lookup_auto(x, codebook, n_frames, out_q, out_idx, s + 1024, reinterpret_cast<int*>(s), st);
}
extern "C" size_t vqcpc_encoder_workspace_bytes(int32_t B, int32_t T, int32_t channels) {
    return vqcpc::encoder_ws_bytes(B, T, channels, VQCPC_GEMM_FP32);
}
extern "C" size_t vqcpc_encoder_workspace_bytes_ex(int32_t B, int32_t T, int32_t channels, int32_t gemm_mode) {
    return vqcpc::encoder_ws_bytes(B, T, channels, gemm_mode);
}
extern "C" int vqcpc_encoder_forward_ex(const vqcpc_encoder_weights* w, const float* mel, int32_t B, int32_t T,
                                        void* workspace, size_t workspace_bytes, float* out_z, float* out_c,
                                        int64_t* out_idx, float* out_prevq, float* out_hidden, int32_t gemm_mode,
                                        void* stream) {
    return vqcpc::encoder_forward(w, mel, B, T, workspace, workspace_bytes, out_z, out_c, out_idx, out_prevq,
                                  out_hidden, gemm_mode, static_cast<cudaStream_t>(stream));
}
extern "C" int vqcpc_encoder_forward(const vqcpc_encoder_weights* w, const float* mel, int32_t B, int32_t T,
                                     void* workspace, size_t workspace_bytes, float* out_z, float* out_c,
                                     int64_t* out_idx, float* out_prevq, float* out_hidden, void* stream) {
    return vqcpc::encoder_forward(w, mel, B, T, workspace, workspace_bytes, out_z, out_c, out_idx, out_prevq,
                                  out_hidden, VQCPC_GEMM_FP32, static_cast<cudaStream_t>(stream));
}
extern "C" size_t vqcpc_lstm_workspace_bytes(int32_t B, int32_t Tp) {
    (void)Tp;
    return vqcpc::lstm_ws_bytes(B);
}
extern "C" int vqcpc_lstm_forward_ex(const vqcpc_encoder_weights* w, const int64_t* idx, int32_t B, int32_t Tp,
                                     void* workspace, size_t workspace_bytes, float* out_c, int32_t gemm_mode, void* stream) {
    VQ_ARG(gemm_mode == VQCPC_GEMM_FP32 || gemm_mode == VQCPC_GEMM_BF16X3 || gemm_mode == VQCPC_GEMM_BF16,
           "lstm: unknown gemm_mode %d", gemm_mode);
    const int rc = vqcpc::lstm_forward(w, idx, B, Tp, workspace, workspace_bytes, out_c, static_cast<cudaStream_t>(stream), true,
                                       gemm_mode);
    if (rc == VQCPC_OK && idx != nullptr && workspace != nullptr && B > 0 && Tp > 0) {
        // caller-supplied indices: an index outside [0, 512) is clamped by the gathers and reported here (after the run, so that
        // the workspace header the run resets carries it): vqcpc_check_status(workspace) returns VQCPC_ERR_ARG
        const int64_t n = static_cast<int64_t>(B) * Tp;
        vqcpc::check_codes_kernel<<<static_cast<unsigned>((n + 1023) / 1024 < 256 ? (n + 1023) / 1024 : 256), 256, 0,
                                    static_cast<cudaStream_t>(stream)>>>(idx, n, static_cast<int*>(workspace));
        if (cudaGetLastError() != cudaSuccess) return VQCPC_ERR_CUDA;
        vqcpc::count_launch(1);
    }
    return rc;
}
extern "C" int vqcpc_lstm_forward(const vqcpc_encoder_weights* w, const int64_t* idx, int32_t B, int32_t Tp,
                                  void* workspace, size_t workspace_bytes, float* out_c, void* stream) {
    return vqcpc_lstm_forward_ex(w, idx, B, Tp, workspace, workspace_bytes, out_c, VQCPC_GEMM_FP32, stream);
}
