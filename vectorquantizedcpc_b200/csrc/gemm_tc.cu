// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[m, n] = sum_k A'[m, k] * W'[n, k]  (+ bias[n]),  fp32 out.
//
// A' and W' are bf16, K-major.  fp32-grade accuracy comes from a hi/lo split: every fp32 operand x is stored as
// two bf16 planes  hi = bf16(x), lo = bf16(x - hi)  side by side ([hi | lo], 2K columns), and the kernel walks
// THREE K-segments  (A_hi, W_hi), (A_hi, W_lo), (A_lo, W_hi)  accumulating all of them into the same fp32 TMEM
// accumulator -- the dropped lo*lo term is 2^-16 relative.  nseg = 1 gives the plain bf16 product.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor.2d of a 128 x 64 A tile and a BN x 64 W tile per stage
//                               (SWIZZLE_128B, 4 stages, mbarrier complete_tx)
//   warp 1   : MMA issuer     -- one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x 4 per
//                               stage, accumulators in TMEM (2 x BN columns: double buffered against the epilogue),
//                               tcgen05.commit releases smem stages / publishes finished accumulators
//   warps 2-5: epilogue       -- tcgen05.ld 32 lanes x 32 columns, (+bias), fp32 stores; each thread owns one row
// Every mbarrier wait is bounded; on timeout the kernel raises a device-side error flag and drains.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vqcpc {

struct TcParams {
    float* C;
    const float* bias;      // nullable
    int* err;
    long long ldc;
    int M, N, K;            // K = columns of ONE plane (multiple of 64)
    int nseg;               // 1: plain bf16 product; 3: hi/lo split, planes stored [hi | lo] (2K columns)
    // Fused LSTM step (lstm_table != nullptr): the product is W_hh h_{t-1} for N = 4 * hidden gate rows.  A column tile of
    // BN columns holds the four gates of BN/4 hidden units (the producer loads four BN/4-row boxes of W, one per gate), the
    // epilogue adds the gathered input projection, applies the cell and writes h_t (fp32 into the output sequence and
    // bf16 hi/lo planes for the next step, into the OTHER plane buffer: this step's A operand is still being read).
    const float* lstm_table;        // (n_codes, 4 * hidden) = W_ih e + b_ih + b_hh per code
    const int64_t* lstm_idx;        // (M, Tp) code indices
    float* lstm_c;                  // (M, hidden) cell state, in place
    float* lstm_out;                // (M, Tp, hidden)
    __nv_bfloat16* lstm_planes_out; // (M, 2 * hidden) [hi | lo] of h_t
    int lstm_t, lstm_Tp;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, TcParams p) {
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;        // 16 KB
    constexpr uint32_t W_BYTES = BN * TC_BK * 2;           // 32 KB at BN = 256
    constexpr uint32_t STAGE_BYTES = A_BYTES + W_BYTES;
    constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;     // power of two >= 32 (BN in {64, 128, 256})
    // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                               (static_cast<uint32_t>(TC_BM >> 4) << 24);

    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) may overlap the tail of the
    // previous kernel in the stream; from here on its results are read and its inputs overwritten.  (No-ops when the
    // kernel was launched without the attribute.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int tiles_m = (p.M + TC_BM - 1) / TC_BM, tiles_n = p.N / BN;
    const int n_tiles = tiles_m * tiles_n;
    const int kb_per_seg = p.K / TC_BK;
    const int n_kb = kb_per_seg * p.nseg;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
                const int tm = tile / tiles_n, tn = tile % tiles_n;       // consecutive CTAs share an A row block
                for (int kb = 0; kb < n_kb && ok; ++kb) {
                    ok = mbar_wait(&empty_bar[stage], phase ^ 1, p.err);
                    if (!ok) break;
                    const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * TC_BK;
                    const int a_col = (seg == 2 ? p.K : 0) + kk;           // (hi, hi, lo)
                    const int w_col = (seg == 1 ? p.K : 0) + kk;           // (hi, lo, hi)
                    unsigned char* sa = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(sa, &map_a, a_col, tm * TC_BM, &full_bar[stage]);
                    if (p.lstm_table == nullptr) {
                        tma_load_2d(sa + A_BYTES, &map_w, w_col, tn * BN, &full_bar[stage]);
                    } else {
                        // gate g of units tn*BN/4 .. : rows g * hidden + tn * BN/4 of W_hh, BN/4 rows each (the W map's box)
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            tma_load_2d(sa + A_BYTES + g * (BN / 4) * TC_BK * 2, &map_w, w_col, g * (p.N / 4) + tn * (BN / 4), &full_bar[stage]);
                    }
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0, 0};
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
                ok = mbar_wait(&tempty_bar[acc], acc_phase[acc] ^ 1, p.err);     // epilogue has drained this accumulator
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < n_kb && ok; ++kb) {
                    ok = mbar_wait(&full_bar[stage], phase, p.err);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)        // +32 bytes (>> 4 = 2) per K = 16 step inside the 128 B atom
                        tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                    tc_commit(&empty_bar[stage]);               // smem stage reusable once these MMAs have read it
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull_bar[acc]);                     // accumulator complete
                acc_phase[acc] ^= 1;
                acc ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp & 3;                           // TMEM lanes 32*quarter .. +31 are this warp's
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        bool ok = true;
        for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
            const int tm = tile / tiles_n, tn = tile % tiles_n;
            ok = mbar_wait(&tfull_bar[acc], acc_phase[acc], p.err);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            tc_fence_after();
            const int row = tm * TC_BM + quarter * 32 + lane;
            if (p.lstm_table != nullptr) {
                // ---- fused LSTM cell: 8 hidden units at a time, their four gates from four column blocks of the accumulator
                constexpr int UT = BN / 4;                      // hidden units of this column tile
                const int hidden = p.N / 4;
                const uint32_t tb = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
                const float* trow = nullptr;
                if (row < p.M) trow = p.lstm_table + __ldg(p.lstm_idx + static_cast<long long>(row) * p.lstm_Tp + p.lstm_t) * p.N;
#pragma unroll 1
                for (int ug = 0; ug < UT; ug += 8) {
                    uint32_t v[4][8];
#pragma unroll
                    for (int g = 0; g < 4; ++g) tc_ld8(tb + g * UT + ug, v[g]);
                    tc_wait_ld();
                    if (row < p.M) {
                        const int u0 = tn * UT + ug;
                        float x[4][8];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(trow + g * hidden + u0));
                            const float4 b = __ldg(reinterpret_cast<const float4*>(trow + g * hidden + u0) + 1);
                            x[g][0] = a.x; x[g][1] = a.y; x[g][2] = a.z; x[g][3] = a.w; x[g][4] = b.x; x[g][5] = b.y; x[g][6] = b.z; x[g][7] = b.w;
                        }
                        float* cp = p.lstm_c + static_cast<long long>(row) * hidden + u0;
                        const float4 c0v = *reinterpret_cast<const float4*>(cp), c1v = *(reinterpret_cast<const float4*>(cp) + 1);
                        float c[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
                        float h[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            // same expression order as lstm_gate_kernel: table + gates, then the cell
                            const float gi = x[0][k] + __uint_as_float(v[0][k]), gf = x[1][k] + __uint_as_float(v[1][k]);
                            const float gg = x[2][k] + __uint_as_float(v[2][k]), go = x[3][k] + __uint_as_float(v[3][k]);
                            c[k] = sigmoid_fast(gf) * c[k] + sigmoid_fast(gi) * tanh_fast(gg);
                            h[k] = sigmoid_fast(go) * tanh_fast(c[k]);
                        }
                        *reinterpret_cast<float4*>(cp) = make_float4(c[0], c[1], c[2], c[3]);
                        *(reinterpret_cast<float4*>(cp) + 1) = make_float4(c[4], c[5], c[6], c[7]);
                        float* op = p.lstm_out + (static_cast<long long>(row) * p.lstm_Tp + p.lstm_t) * hidden + u0;
                        *reinterpret_cast<float4*>(op) = make_float4(h[0], h[1], h[2], h[3]);
                        *(reinterpret_cast<float4*>(op) + 1) = make_float4(h[4], h[5], h[6], h[7]);
                        __nv_bfloat16* pr = p.lstm_planes_out + static_cast<long long>(row) * 2 * hidden + u0;
                        tc_split_store4(make_float4(h[0], h[1], h[2], h[3]), pr, pr + hidden);
                        tc_split_store4(make_float4(h[4], h[5], h[6], h[7]), pr + 4, pr + hidden + 4);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                acc_phase[acc] ^= 1;
                acc ^= 1;
                continue;
            }
            float* crow = p.C + static_cast<long long>(row) * p.ldc + tn * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c0, v);
                tc_wait_ld();
                if (row < p.M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]);
                        o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
                        if (p.bias != nullptr) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + tn * BN + c0 + j));
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        *reinterpret_cast<float4*>(crow + c0 + j) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc_phase[acc] ^= 1;
            acc ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------
// GEMM with a fused LayerNorm + ReLU + hi/lo-split epilogue:   planes = split(relu(LN(A' . W'^T)))
// (nn.Linear(C, C, bias=False) -> nn.LayerNorm(C) -> nn.ReLU of /root/reference/model.py:50-52, and the conv ->
// LN -> ReLU head, model.py:43-48).  N = NT * 256 <= 768: a CTA owns a whole 128-row block and walks its NT
// column tiles through the two TMEM accumulators, so an epilogue thread (= one row) sees the complete row:
//   pass A (per tile, overlapped with the next tile's MMAs): tcgen05.ld -> stash raw y in an L2-resident
//          per-CTA scratch row, accumulate shifted sums  s1 = sum(y - K), s2 = sum((y - K)^2),  K = y[row][0];
//   pass B (after the last tile, overlapped with the next row block's MMAs): mean = K + s1/N,
//          var = s2/N - (s1/N)^2 (biased, eps 1e-5), re-read the row, affine, ReLU, and write it straight as the
//          bf16 hi/lo planes the next tcgen05 GEMM consumes (optionally also as fp32).
// The fp32 activation never makes a round trip through a separate LayerNorm kernel.
// ----------------------------------------------------------------------------------------------
struct TcLnParams {
    const float* ln_w; const float* ln_b;
    __nv_bfloat16* out_planes;   // (M, 2N): [hi | lo]
    float* out_f32;              // (M, N) or null
    float* scratch;              // gridDim.x * 128 * N floats
    int* err;
    int M, N, K, nseg;
    int debug;                   // bit 0: skip the scratch stores of pass A, bit 1: skip pass B (timing experiments)
};

template <int NT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, TcLnParams p) {
    constexpr int BN = 256;
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2, W_BYTES = BN * TC_BK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                               (static_cast<uint32_t>(TC_BM >> 4) << 24);
    constexpr int N = NT * BN;

    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float ep_tile[4][32 * 33];                 // per epilogue warp: 32 x 32 transpose tile (padded)

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int n_blocks = (p.M + TC_BM - 1) / TC_BM;
    const int kb_per_seg = p.K / TC_BK;
    const int n_kb = kb_per_seg * p.nseg;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int rb = blockIdx.x; rb < n_blocks && ok; rb += gridDim.x) {
                for (int tn = 0; tn < NT && ok; ++tn) {
                    for (int kb = 0; kb < n_kb && ok; ++kb) {
                        ok = mbar_wait(&empty_bar[stage], phase ^ 1, p.err);
                        if (!ok) break;
                        const int seg = kb / kb_per_seg, kk = (kb - seg * kb_per_seg) * TC_BK;
                        const int a_col = (seg == 2 ? p.K : 0) + kk, w_col = (seg == 1 ? p.K : 0) + kk;
                        unsigned char* sa = smem + stage * STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(sa, &map_a, a_col, rb * TC_BM, &full_bar[stage]);
                        tma_load_2d(sa + A_BYTES, &map_w, w_col, tn * BN, &full_bar[stage]);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0, 0};
            bool ok = true;
            for (int rb = blockIdx.x; rb < n_blocks && ok; rb += gridDim.x) {
                for (int tn = 0; tn < NT && ok; ++tn) {
                    ok = mbar_wait(&tempty_bar[acc], acc_phase[acc] ^ 1, p.err);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    for (int kb = 0; kb < n_kb && ok; ++kb) {
                        ok = mbar_wait(&full_bar[stage], phase, p.err);
                        if (!ok) break;
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                        tc_commit(&empty_bar[stage]);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&tfull_bar[acc]);
                    acc_phase[acc] ^= 1;
                    acc ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        // pass A: thread = row (TMEM lane).  Each 32 x 32 chunk goes through a padded shared tile so that the raw y
        //         lands in the per-CTA scratch with fully coalesced 128-byte row segments.
        // pass B: warp walks its 32 rows; the whole warp reads one row (coalesced), lane l owns columns
        //         {128 j + 4 l .. +3}; the row's mean / rstd come from the lane that owns the row (shuffle).
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        bool ok = true;
        float* tile = &ep_tile[quarter][0];
        float* sblk = p.scratch + static_cast<size_t>(blockIdx.x) * TC_BM * N;       // this CTA's 128 scratch rows
        constexpr int NJ = N / 128;
        float4 lw[NJ], lb[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            lw[j] = __ldg(reinterpret_cast<const float4*>(p.ln_w + 128 * j) + lane);
            lb[j] = __ldg(reinterpret_cast<const float4*>(p.ln_b + 128 * j) + lane);
        }
        for (int rb = blockIdx.x; rb < n_blocks && ok; rb += gridDim.x) {
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            for (int tn = 0; tn < NT && ok; ++tn) {
                ok = mbar_wait(&tfull_bar[acc], acc_phase[acc], p.err);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c0, v);
                    tc_wait_ld();
                    if (tn == 0 && c0 == 0) shift = __uint_as_float(v[0]);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float y = __uint_as_float(v[j]);
                        const float d = y - shift;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                        tile[lane * 33 + j] = y;                    // [row = lane][col j], padded: conflict free
                    }
                    __syncwarp();
                    if (!(p.debug & 1))
#pragma unroll 8
                    for (int r = 0; r < 32; ++r)                   // row r of the chunk: 32 consecutive floats, one line
                        sblk[static_cast<size_t>(quarter * 32 + r) * N + tn * BN + c0 + lane] = tile[r * 33 + lane];
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                acc_phase[acc] ^= 1;
                acc ^= 1;
            }
            if (!ok) break;
            // per-row statistics live in the lane that owns the row
            const float m1 = s1 * (1.0f / N);
            const float mean_own = shift + m1;
            const float rstd_own = 1.0f / sqrtf(fmaxf(s2 * (1.0f / N) - m1 * m1, 0.f) + 1e-5f);
            __syncwarp();                                          // this warp's scratch rows are complete (same-warp writes)
            // software pipelined over rows: the loads of row r+1 are issued before row r is normalised and stored
            // (the scratch loads may not be reordered across the plane stores by the compiler, so do it by hand)
            const float* src0 = sblk + static_cast<size_t>(quarter * 32) * N;
            float4 ycur[NJ], ynext[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) ycur[j] = *(reinterpret_cast<const float4*>(src0 + 128 * j) + lane);
#pragma unroll 1
            for (int r = 0; r < ((p.debug & 2) ? 0 : 32); ++r) {
                const float mean = __shfl_sync(0xffffffffu, mean_own, r), rstd = __shfl_sync(0xffffffffu, rstd_own, r);
                const int row = rb * TC_BM + quarter * 32 + r;
                if (row >= p.M) break;                             // warp-uniform
                if (r + 1 < 32) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) ynext[j] = *(reinterpret_cast<const float4*>(src0 + static_cast<size_t>(r + 1) * N + 128 * j) + lane);
                }
                __nv_bfloat16* prow = p.out_planes + static_cast<size_t>(row) * 2 * N;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float4 y = ycur[j];
                    float4 o;
                    o.x = fmaxf(fmaf((y.x - mean) * rstd, lw[j].x, lb[j].x), 0.f);
                    o.y = fmaxf(fmaf((y.y - mean) * rstd, lw[j].y, lb[j].y), 0.f);
                    o.z = fmaxf(fmaf((y.z - mean) * rstd, lw[j].z, lb[j].z), 0.f);
                    o.w = fmaxf(fmaf((y.w - mean) * rstd, lw[j].w, lb[j].w), 0.f);
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
                    const uint32_t hb0 = *reinterpret_cast<const uint32_t*>(&h0), hb1 = *reinterpret_cast<const uint32_t*>(&h1);
                    const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - __uint_as_float(hb0 << 16), o.y - __uint_as_float(hb0 & 0xffff0000u));
                    const __nv_bfloat162 l1 = __floats2bfloat162_rn(o.z - __uint_as_float(hb1 << 16), o.w - __uint_as_float(hb1 & 0xffff0000u));
                    uint2 hv, lv;
                    hv.x = hb0; hv.y = hb1;
                    lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
                    *(reinterpret_cast<uint2*>(prow + 128 * j) + lane) = hv;              // 256 B coalesced per warp
                    *(reinterpret_cast<uint2*>(prow + N + 128 * j) + lane) = lv;
                    if (p.out_f32 != nullptr) *(reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(row) * N + 128 * j) + lane) = o;
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) ycur[j] = ynext[j];
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// bf16 row-major (rows x cols), box = 64 columns (128 B) x box_rows, SWIZZLE_128B, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld_elems, int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) { set_error("gemm_tc: cuTensorMapEncodeTiled is unavailable"); return VQCPC_ERR_CUDA; }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
    cuuint32_t box[2] = {TC_BK, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return VQCPC_ERR_CUDA; }
    return VQCPC_OK;
}

int make_map_bf16(void* map, const void* base, long long rows, long long cols, long long ld_elems, int box_rows) {
    return make_map(static_cast<CUtensorMap*>(map), base, rows, cols, ld_elems, box_rows);
}
// general box / swizzle (TMA-store maps of the pair kernel)
int make_map_bf16_box(void* map, const void* base, long long rows, long long cols, long long ld_elems, int box_cols, int box_rows,
                      int swizzle_bytes) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) { set_error("gemm_tc: cuTensorMapEncodeTiled is unavailable"); return VQCPC_ERR_CUDA; }
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(static_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return VQCPC_ERR_CUDA; }
    return VQCPC_OK;
}

template <int BN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mw, const TcParams& p, cudaStream_t stream, bool pdl) {
    constexpr size_t smem = TC_STAGES * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 1024;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_tc_kernel<BN>), static_cast<int>(smem))) return rc_attr;
    const int tiles = ((p.M + TC_BM - 1) / TC_BM) * (p.N / BN);
    const int sms = device_sm_count();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles < sms ? tiles : sms);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    VQ_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN>, ma, mw, p));
    count_launch(1);
    return VQCPC_OK;
}

// A planes (M x nplanes*K) bf16 with nplanes = (nseg == 3 ? 2 : 1); W planes are ALWAYS stored (N x 2K) = [hi | lo] (weights are
// split once at pack time) -- nseg = 1 reads only their hi half.
// A plan holds the two tensor maps so that a GEMM repeated on the same buffers (the LSTM's per-step product)
// encodes them once.
int gemm_tc_plan(TcPlan* plan, const void* a_planes, const void* w_planes, const float* bias, float* C, long long ldc, int M,
                 int N, int K, int nseg, int* err_flag) {
    VQ_ARG(plan && a_planes && w_planes && C && err_flag, "gemm_tc: null pointer");
    VQ_ARG(M > 0, "gemm_tc: empty problem");
    VQ_ARG(nseg == 1 || nseg == 3, "gemm_tc: nseg must be 1 or 3");
    VQ_ARG(K % TC_BK == 0 && K > 0, "gemm_tc: K=%d must be a multiple of %d", K, TC_BK);
    VQ_ARG(N % 64 == 0 && ldc % 4 == 0, "gemm_tc: N=%d must be a multiple of 64", N);
    static_assert(sizeof(CUtensorMap) == sizeof(plan->map_a), "TcPlan map storage");
    const int planes = nseg == 3 ? 2 : 1;
    plan->bn = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
    {
        // small problems (the LSTM's per-step product at a few hundred utterances): narrower column tiles until at
        // least half of the 148 SMs have a tile
        long long tiles = static_cast<long long>((M + TC_BM - 1) / TC_BM) * (N / plan->bn);
        while (plan->bn > 64 && tiles * 2 <= 148) { plan->bn /= 2; tiles *= 2; }
    }
    int rc = make_map(reinterpret_cast<CUtensorMap*>(plan->map_a), a_planes, M, static_cast<long long>(planes) * K,
                      static_cast<long long>(planes) * K, TC_BM);
    if (rc) return rc;
    rc = make_map(reinterpret_cast<CUtensorMap*>(plan->map_w), w_planes, N, 2LL * K, 2LL * K, plan->bn);
    if (rc) return rc;
    plan->C = C; plan->bias = bias; plan->err = err_flag; plan->ldc = ldc; plan->M = M; plan->N = N; plan->K = K; plan->nseg = nseg;
    return VQCPC_OK;
}

// Fused LSTM step: plan over (h planes of step t-1, W_hh planes); the W map's box is BN/4 rows (one gate of a tile)
int gemm_tc_plan_lstm(TcPlan* plan, const void* h_planes, const void* whh_planes, int M, int hidden, int* err_flag) {
    VQ_ARG(plan && h_planes && whh_planes && err_flag, "gemm_tc(lstm): null pointer");
    VQ_ARG(M > 0 && hidden % 64 == 0, "gemm_tc(lstm): bad shape M=%d hidden=%d", M, hidden);
    const int N = 4 * hidden, K = hidden;
    plan->bn = 256;
    {
        long long tiles = static_cast<long long>((M + TC_BM - 1) / TC_BM) * (N / plan->bn);
        while (plan->bn > 64 && tiles * 2 <= 148) { plan->bn /= 2; tiles *= 2; }
    }
    int rc = make_map(reinterpret_cast<CUtensorMap*>(plan->map_a), h_planes, M, 2LL * K, 2LL * K, TC_BM);
    if (rc) return rc;
    rc = make_map(reinterpret_cast<CUtensorMap*>(plan->map_w), whh_planes, N, 2LL * K, 2LL * K, plan->bn / 4);
    if (rc) return rc;
    plan->C = nullptr; plan->bias = nullptr; plan->err = err_flag; plan->ldc = 0; plan->M = M; plan->N = N; plan->K = K; plan->nseg = 3;
    return VQCPC_OK;
}
int gemm_tc_run_lstm(const TcPlan* plan, const float* table, const int64_t* idx, float* cstate, float* out, void* planes_out, int t,
                     int Tp, cudaStream_t stream, bool pdl) {
    const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(plan->map_a);
    const CUtensorMap& mw = *reinterpret_cast<const CUtensorMap*>(plan->map_w);
    TcParams p{nullptr, nullptr, plan->err, 0, plan->M, plan->N, plan->K, plan->nseg,
               table, idx, cstate, out, static_cast<__nv_bfloat16*>(planes_out), t, Tp};
    if (plan->bn == 256) return launch_tc<256>(ma, mw, p, stream, pdl);
    if (plan->bn == 128) return launch_tc<128>(ma, mw, p, stream, pdl);
    return launch_tc<64>(ma, mw, p, stream, pdl);
}

int gemm_tc_run(const TcPlan* plan, cudaStream_t stream, bool pdl) {
    const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(plan->map_a);
    const CUtensorMap& mw = *reinterpret_cast<const CUtensorMap*>(plan->map_w);
    TcParams p{plan->C, plan->bias, plan->err, plan->ldc, plan->M, plan->N, plan->K, plan->nseg, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    if (plan->bn == 256) return launch_tc<256>(ma, mw, p, stream, pdl);
    if (plan->bn == 128) return launch_tc<128>(ma, mw, p, stream, pdl);
    return launch_tc<64>(ma, mw, p, stream, pdl);
}

int gemm_tc(const void* a_planes, const void* w_planes, const float* bias, float* C, long long ldc, int M, int N, int K,
            int nseg, int* err_flag, cudaStream_t stream) {
    if (M == 0) return VQCPC_OK;
    TcPlan plan;
    int rc = gemm_tc_plan(&plan, a_planes, w_planes, bias, C, ldc, M, N, K, nseg, err_flag);
    if (rc) return rc;
    return gemm_tc_run(&plan, stream, false);
}


// relu(LayerNorm(A . W^T)) written as bf16 hi/lo planes (and optionally fp32): N must be 512 or 768.
// scratch: gemm_tc_ln_scratch_bytes(N) bytes of device memory.
size_t gemm_tc_ln_scratch_bytes(int N) { return static_cast<size_t>(148) * TC_BM * N * sizeof(float); }
bool gemm_tc_ln_supported(int N) { return N == 512 || N == 768; }

template <int NT>
static int launch_tc_ln(const CUtensorMap& ma, const CUtensorMap& mw, const TcLnParams& p, cudaStream_t stream) {
    constexpr size_t smem = TC_STAGES * (TC_BM * TC_BK * 2 + 256 * TC_BK * 2) + 1024;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_tc_ln_kernel<NT>), static_cast<int>(smem))) return rc_attr;
    const int blocks = (p.M + TC_BM - 1) / TC_BM;
    int sms = device_sm_count();
    if (sms > 148) sms = 148;
    gemm_tc_ln_kernel<NT><<<blocks < sms ? blocks : sms, TC_THREADS, smem, stream>>>(ma, mw, p);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

int gemm_tc_ln(const void* a_planes, const void* w_planes, const float* ln_w, const float* ln_b, void* out_planes,
               float* out_f32, void* scratch, int M, int N, int K, int nseg, int* err_flag, cudaStream_t stream) {
    if (M == 0) return VQCPC_OK;
    VQ_ARG(a_planes && w_planes && ln_w && ln_b && out_planes && scratch && err_flag, "gemm_tc_ln: null pointer");
    VQ_ARG(gemm_tc_ln_supported(N), "gemm_tc_ln: N=%d must be 512 or 768", N);
    VQ_ARG(K % TC_BK == 0 && K > 0 && (nseg == 1 || nseg == 3), "gemm_tc_ln: bad K / nseg");
    const int planes = nseg == 3 ? 2 : 1;
    CUtensorMap ma, mw;
    int rc = make_map(&ma, a_planes, M, static_cast<long long>(planes) * K, static_cast<long long>(planes) * K, TC_BM);
    if (rc) return rc;
    rc = make_map(&mw, w_planes, N, 2LL * K, 2LL * K, 256);
    if (rc) return rc;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VQCPC_LN_DEBUG"); dbg = e ? atoi(e) : 0; }
    TcLnParams p{ln_w, ln_b, static_cast<__nv_bfloat16*>(out_planes), out_f32, static_cast<float*>(scratch), err_flag, M, N, K, nseg, dbg};
    return N == 768 ? launch_tc_ln<3>(ma, mw, p, stream) : launch_tc_ln<2>(ma, mw, p, stream);
}

// ---------------------------------------------------------------------------------------------- bf16 hi/lo planes
__device__ __forceinline__ void split2(float x, float y, __nv_bfloat162& hi, __nv_bfloat162& lo) {
    const __nv_bfloat16 hx = __float2bfloat16_rn(x), hy = __float2bfloat16_rn(y);
    hi = __halves2bfloat162(hx, hy);
    lo = __halves2bfloat162(__float2bfloat16_rn(x - __bfloat162float(hx)), __float2bfloat16_rn(y - __bfloat162float(hy)));
}

// fp32 (rows x K, leading dimension ld) -> bf16 planes (rows x 2K): [hi | lo]
__global__ void split_planes_kernel(const float* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ out, long long rows, int K) {
    const long long total = rows * (K / 4);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / (K / 4);
        const int q = static_cast<int>(i - r * (K / 4));
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld) + q);
        __nv_bfloat162 h0, l0, h1, l1;
        split2(v.x, v.y, h0, l0);
        split2(v.z, v.w, h1, l1);
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(out + r * 2 * K + 4 * q);
        __nv_bfloat162* ol = reinterpret_cast<__nv_bfloat162*>(out + r * 2 * K + K + 4 * q);
        oh[0] = h0; oh[1] = h1; ol[0] = l0; ol[1] = l1;
    }
}

int split_planes(const float* x, long long ld, void* out, long long rows, int K, cudaStream_t stream) {
    if (rows == 0) return VQCPC_OK;
    VQ_ARG(x && out && K % 4 == 0 && ld % 4 == 0, "split_planes: bad arguments");
    const long long total = rows * (K / 4);
    const unsigned grid = static_cast<unsigned>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    split_planes_kernel<<<grid, 256, 0, stream>>>(x, ld, static_cast<__nv_bfloat16*>(out), rows, K);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

}  // namespace vqcpc

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" int vqcpc_split_planes(const float* x, int64_t ld, void* out_planes, int64_t rows, int32_t K, void* stream) {
    return vqcpc::split_planes(x, ld, out_planes, rows, K, static_cast<cudaStream_t>(stream));
}
// C = A . W^T (+bias) through the tensor-core path; A (M x K) and W (N x K) fp32 in, split on the fly into the
// caller-provided plane buffers (a_planes: M x 2K bf16, w_planes: N x 2K bf16).  mode 3 = hi/lo split (fp32-grade),
// mode 1 = plain bf16.
extern "C" int vqcpc_linear_tc(const float* A, const float* W, const float* bias, float* C, int64_t M, int32_t N, int32_t K,
                               int32_t mode, void* a_planes, void* w_planes, int32_t* err_flag, void* stream) {
    using namespace vqcpc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc;
    if ((rc = split_planes(A, K, a_planes, M, K, s))) return rc;
    if ((rc = split_planes(W, K, w_planes, N, K, s))) return rc;
    // mode 1 reads only the hi plane (plane row stride stays 2K)
    if (mode == 3) return gemm_tc(a_planes, w_planes, bias, C, N, static_cast<int>(M), N, K, 3, err_flag, s);
    vqcpc::set_error("vqcpc_linear_tc: only mode 3 is exposed");
    return VQCPC_ERR_ARG;
}
