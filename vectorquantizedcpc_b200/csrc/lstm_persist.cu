// Persistent batched LSTM recurrence of Encoder.encode (/root/reference/model.py:57,69: nn.LSTM(64, 256, batch_first)) for
// B >= 64 utterances: ONE launch for all T' steps.
//
// Why: per step the recurrence is  gates = H_{t-1} . W_hh^T  (B x 1024, K = 256)  + the gathered input projection of the
// frame's code, i.e. 0.27 GFLOP at 512 utterances -- 0.2 us of tensor time.  The per-step-launch path (gemm_tc_kernel in its
// fused-LSTM mode) costs 14.5 us per step at 512 utterances (launch, TMEM allocation, pipeline fill, W_hh re-fetched from L2
// 150 times), which is 62 % of the whole 512 x 3 s encode.  Here
//   * the grid is  (row tiles of 128 utterances) x (slices of UT hidden units), at most one CTA per SM, co-resident
//     (cooperative launch); CTA (mt, ns) keeps the 4 UT gate rows of its units resident in shared memory for the whole
//     sequence as bf16 hi/lo planes in UMMA K-major SWIZZLE_128B layout (loaded once by TMA: UT = 8 / 16 / 32 -> 32 / 64 / 128 KB);
//   * per step: TMA streams the row tile's h_{t-1} planes (128 x 256 hi + lo = 128 KB) through a ring of 32 KB stages, one
//     thread issues tcgen05.mma (M = 128, N = 4 UT, K = 16; the three hi/lo products per k-block from one stage), four or
//     eight epilogue warps read the accumulator from TMEM in mma-fragment layout (16x256b: four threads per utterance row), add
//     the gathered row of the input-projection table (prefetched a chunk / a step ahead), apply the cell with the cell state in
//     REGISTERS for the whole sequence, and write h_t as fp32 into the output sequence and as bf16 hi/lo planes into the other
//     plane buffer;
//   * the CTAs of a row tile (and only those: utterances are independent) meet at a release/acquire counter before the next
//     step's TMA loads -- no grid-wide barrier, no host involvement;
//   * every CTA of a row tile needs the SAME 128 KB of h_{t-1}: optionally (VQCPC_LSTM_CLUSTER=4, measured no faster) clusters of
//     4 CTAs (neighbouring unit slices) load a quarter each and TMA-multicast it to all four.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vqcpc {

// threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2 .. 2 + EW - 1 epilogue of row tile 0 (then EW more for row tile 1)
constexpr int LP_H = 256, LP_G = 4 * LP_H;
constexpr int LP_KB = LP_H / TC_BK;   // 4 k-blocks of 64

struct LpParams {
    const float* table;          // (512, 1024) = W_ih e + b_ih + b_hh per code, PERMUTED: [code][unit / 8][gate][unit % 8]
    const int64_t* idx;          // (B, Tp)
    float* out;                  // (B, Tp, 256)
    __nv_bfloat16* planes[2];    // (B, 512) [hi | lo] of h_t, ping-pong across steps
    unsigned* counters;          // one per row tile, zeroed before launch
    int* err;
    int B, Tp, n_ns;
    int dbg;                     // VQCPC_LP_DEBUG ablation bits (timing only; results are wrong with any bit set)
};

__device__ long long lp_trace[8 * 8];     // VQCPC_LP_DEBUG & 32: clock64 stamps of CTA 0, row tile 0, steps 60..67 (printed by the host)
#define LP_STAMP(slot) do { if ((p.dbg & 32) && blockIdx.x == 0 && t >= 60 && t < 68) lp_trace[(t - 60) * 8 + (slot)] = clock64(); } while (0)
// LSTM cell with 7 MUFU instead of 10: the three sigmoids / two tanhs of  c' = s(f) c + s(i) tanh(g),  h = s(o) tanh(c')  put on
// common denominators -- c' = [c (1+Ei)(1+Eg) + (1-Eg)(1+Ef)] / [(1+Ef)(1+Ei)(1+Eg)],  h = (1-Ec) / [(1+Eo)(1+Ec)]  with
// Ex = e^-x (sigmoid) or e^-2x (tanh) from ex2.approx, one rcp.approx each.  Arguments are clamped from below (-28 / -14) so that
// a product of three (1 + E) terms stays finite (<= e^84); there sigmoid / 1 + tanh are < 7e-13, below fp32 resolution of the sums.
__device__ __forceinline__ float lp_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lp_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lp_cell(float gi, float gf, float gg, float go, float& c) {
    constexpr float L2E = 1.4426950408889634f;
    const float ef = lp_ex2(-L2E * fmaxf(gf, -28.f)), ei = lp_ex2(-L2E * fmaxf(gi, -28.f)), eg = lp_ex2(-2.f * L2E * fmaxf(gg, -14.f));
    const float a = 1.f + ef, big = (1.f + ei) * (1.f + eg);
    c = fmaf(c, big, (1.f - eg) * a) * lp_rcp(a * big);
    const float eo = lp_ex2(-L2E * fmaxf(go, -28.f)), ec = lp_ex2(-2.f * L2E * fmaxf(c, -14.f));
    return (1.f - ec) * lp_rcp((1.f + eo) * (1.f + ec));
}
__device__ __forceinline__ int lp_clamp_code(int64_t id) { return id < 0 ? 0 : (id > 511 ? 511 : static_cast<int>(id)); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// NT = row tiles per CTA (1 or 2).  With two, the CTA alternates between two INDEPENDENT groups of 128 utterances: the loads and
// MMAs of one tile run under the epilogue (and the counter hop) of the other -- used when the batch has more row tiles than
// the grid has room for (4096 utterances: 16 x 8 CTAs x 2 tiles instead of two launches one after the other).
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// all prior MMAs of this CTA complete -> one arrival on the barrier at this offset in every CTA of the mask
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// CS = cluster size (1: every CTA loads its own copy of h_{t-1}; 4: each loads a quarter and multicasts it)
// EW = epilogue warps per row tile (4, 8 or 16)
template <int UT, int NT, int CS, int EW>
__global__ void __launch_bounds__(64 + 32 * EW * NT, 1)
lstm_persist_kernel(const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_h1,
                    const __grid_constant__ CUtensorMap map_w, LpParams p) {
    constexpr int N = 4 * UT;                                   // gate columns of this CTA: [i | f | g | o] x UT units
    constexpr int HSPLIT = (UT / (EW / 4) < 8 || (NT == 2 && EW == 8) || EW == 16) ? 2 : 1;   // 2: the warps of a lane quarter split its two 16-lane halves, not the columns
    constexpr int CGN = EW / 4 / HSPLIT;                        // column groups
    constexpr int UW = UT / CGN;                                // hidden units per epilogue warp
    constexpr int NK = 4 / HSPLIT;                              // utterance rows per thread (each with 2 units per 8-unit group)
    static_assert(EW % 4 == 0 && UW % 8 == 0, "epilogue split: whole 8-unit groups per warp");
    constexpr uint32_t A_TILE = TC_BM * TC_BK * 2;              // 16 KB
    constexpr uint32_t STAGE_BYTES = 2 * A_TILE;                // hi + lo tile of one k-block
    constexpr int STAGES = UT == 32 ? 3 : 4;
    constexpr uint32_t W_TILE = N * TC_BK * 2;                  // one (plane, k-block) tile of the resident W_hh slice
    constexpr uint32_t W_BYTES = 2 * LP_KB * W_TILE;
    constexpr uint32_t TMEM_COLS = NT * N < 32 ? 32 : NT * N;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                               (static_cast<uint32_t>(TC_BM >> 4) << 24);

    extern __shared__ __align__(1024) unsigned char lp_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[NT], tempty_bar[NT], w_bar;
    __shared__ uint32_t tmem_base_slot;

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(lp_smem) + 1023) & ~uintptr_t(1023));
    unsigned char* w_s = smem;                                  // [plane][k-block][N rows x 128 B]
    unsigned char* a_s = smem + W_BYTES;                        // ring of (hi, lo) tiles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt0 = (blockIdx.x / p.n_ns) * NT, ns = blockIdx.x % p.n_ns;      // row tiles mt0 .. mt0 + NT - 1
    const int u0 = ns * UT;
    const int Tp = p.Tp;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CS); }
        for (int j = 0; j < NT; ++j) { mbar_init(&tfull_bar[j], 1); mbar_init(&tempty_bar[j], EW); }
        mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();             // the peers' barriers exist before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t CMASK = static_cast<uint16_t>((1u << CS) - 1);
    constexpr int SH_ROWS = TC_BM / CS;         // rows of an A tile this CTA fetches (and multicasts)

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            // the resident W_hh slice, once: per (plane, k-block) four boxes of UT rows, one per gate
            mbar_expect_tx(&w_bar, W_BYTES);
            for (int pl = 0; pl < 2; ++pl)
                for (int kb = 0; kb < LP_KB; ++kb)
                    for (int g = 0; g < 4; ++g)
                        tma_load_2d(w_s + (pl * LP_KB + kb) * W_TILE + g * UT * TC_BK * 2, &map_w, pl * LP_H + kb * TC_BK,
                                    g * LP_H + u0, &w_bar);
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int t = 1; t < Tp && ok; ++t)
            for (int j = 0; j < NT && ok; ++j) {
                const int mt = mt0 + j;
                if (mt * TC_BM >= p.B) continue;                      // an odd number of row tiles: the last CTA row has one
                const unsigned* ctr = p.counters + mt;
                // h_{t-1} of this row tile is complete once all n_ns CTAs of the tile have finished step t-1
                const unsigned want = static_cast<unsigned>(EW) * static_cast<unsigned>(p.n_ns) * static_cast<unsigned>(t);   // EW epilogue warps per CTA
                const long long t0 = clock64();
                while (!(p.dbg & 8) && ld_acquire_u32(ctr) < want) {
                    if (clock64() - t0 > TC_TIMEOUT) { atomicExch(p.err, VQCPC_ERR_TIMEOUT); ok = false; break; }
                }
                if (!ok) break;
                if (j == 0) LP_STAMP(0);
                asm volatile("fence.proxy.async;" ::: "memory");          // other CTAs' generic-proxy stores -> TMA reads
                const CUtensorMap* mh = ((t - 1) & 1) ? &map_h1 : &map_h0;
                for (int kb = 0; kb < LP_KB && ok; ++kb) {
                    ok = mbar_wait(&empty_bar[stage], phase ^ 1, p.err);
                    if (!ok) break;
                    unsigned char* st = a_s + stage * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    if (CS == 1) {
                        tma_load_2d(st, mh, kb * TC_BK, mt * TC_BM, &full_bar[stage]);
                        tma_load_2d(st + A_TILE, mh, LP_H + kb * TC_BK, mt * TC_BM, &full_bar[stage]);
                    } else {
                        const uint32_t off = crank * SH_ROWS * TC_BK * 2;
                        tma_load_2d_mc(st + off, mh, kb * TC_BK, mt * TC_BM + crank * SH_ROWS, &full_bar[stage], CMASK);
                        tma_load_2d_mc(st + A_TILE + off, mh, LP_H + kb * TC_BK, mt * TC_BM + crank * SH_ROWS, &full_bar[stage], CMASK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (j == 0) LP_STAMP(1);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, tphase[NT] = {};
            bool ok = mbar_wait(&w_bar, 0, p.err);
            for (int t = 1; t < Tp && ok; ++t)
            for (int j = 0; j < NT && ok; ++j) {
                if ((mt0 + j) * TC_BM >= p.B) continue;
                ok = mbar_wait(&tempty_bar[j], tphase[j], p.err);         // the epilogue of step t-1 has drained this accumulator
                if (!ok) break;
                tphase[j] ^= 1;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + j * N;
                for (int kb = 0; kb < LP_KB && ok; ++kb) {
                    ok = mbar_wait(&full_bar[stage], phase, p.err);
                    if (!ok) break;
                    if (j == 0 && kb == 0) LP_STAMP(2);
                    if (j == 0 && kb == LP_KB - 1) LP_STAMP(3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(a_s + stage * STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + A_TILE);
                    const uint64_t w_hi = umma_desc_sw128(smem_u32(w_s + kb * W_TILE));
                    const uint64_t w_lo = umma_desc_sw128(smem_u32(w_s + (LP_KB + kb) * W_TILE));
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) tc_mma_f16(d_tmem, a_hi + 2 * k, w_hi + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) tc_mma_f16(d_tmem, a_hi + 2 * k, w_lo + 2 * k, IDESC, 1u);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) tc_mma_f16(d_tmem, a_lo + 2 * k, w_hi + 2 * k, IDESC, 1u);
                    if (CS == 1) tc_commit(&empty_bar[stage]); else tc_commit_mc(&empty_bar[stage], CMASK);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull_bar[j]);
            }
        }
    } else {
        {
        // ------------------------------------------------------------------ epilogue, fragment layout (tcgen05.ld 16x256b)
        // With thread = utterance every global access of a warp touches 32 different rows = 32 L1 wavefronts per instruction, and
        // the epilogue of the large-batch configurations is bound by exactly that.  The 16x256b shape hands the
        // accumulator out like an mma C fragment instead: per 16 lanes x 8 columns, thread l holds (row l / 4, columns
        // 2 (l % 4) + {0, 1}) and (row l / 4 + 8, same columns).  Four threads then share a row and its 32-byte sectors: a table
        // read or an output store of a warp touches 8 rows instead of 32.  Per 8-unit group a thread owns 4 rows x 2 units.
        const int quarter = warp & 3;
        const int e = warp - 2;
        const int j = e / EW;
        const int wq = (e % EW) >> 2;                           // which of the EW / 4 warps of this TMEM lane quarter
        const int ub = (wq % CGN) * UW;                         // column group: units ub .. ub + UW - 1 of the CTA's UT
        const int hsel = wq / CGN;                              // HSPLIT == 2: this warp owns one 16-lane half of the quarter
        const int mt = mt0 + j;
        const int cp = lane & 3, r8 = lane >> 2;
        const bool tile_live = mt * TC_BM < p.B;
        int rowk[NK];
        bool validk[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            rowk[k] = mt * TC_BM + quarter * 32 + 8 * (k + 2 * hsel) + r8;      // k = 2 * (half of the quarter) + (upper 8 rows of the half)
            validk[k] = rowk[k] < p.B;
            if (!validk[k]) rowk[k] = 0;
        }
        const uint32_t tb = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + j * N;
        float c[UW / 8][NK][2];
#pragma unroll
        for (int q = 0; q < UW / 8; ++q)
#pragma unroll
            for (int k = 0; k < NK; ++k) c[q][k][0] = c[q][k][1] = 0.f;
        uint32_t tphase = 0;
        bool ok = true;
        // code indices are read TWO steps ahead: the table address of step t + 1 (prefetched during step t) must not wait for them
        int code[NK], code_next[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            code[k] = validk[k] ? lp_clamp_code(__ldg(p.idx + static_cast<int64_t>(rowk[k]) * Tp)) : 0;
            code_next[k] = (validk[k] && 1 < Tp) ? lp_clamp_code(__ldg(p.idx + static_cast<int64_t>(rowk[k]) * Tp + 1)) : 0;
        }
        // permuted table, fragment flavour: [code][unit / 8][gate / 2][unit pair][gate % 2][2] -- per instruction the four threads of a
        // row read 64 contiguous bytes (two whole sectors) of the 8-unit group's 128-byte line
        float4 xq[NK][2];
        auto fetch = [&](const int (&cd)[NK], int ug) {
            const int u = u0 + ub + ug;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float4* tr = reinterpret_cast<const float4*>(p.table + static_cast<int64_t>((p.dbg & 1) ? 0 : cd[k]) * LP_G + (u >> 3) * 32) + cp;
                xq[k][0] = __ldg(tr);          // gates i, f of the thread's two units: the four threads of a row read 64 contiguous bytes
                xq[k][1] = __ldg(tr + 4);      // gates g, o
            }
        };
        fetch(code, 0);
        for (int t = 0; t < Tp && ok && tile_live; ++t) {
            int code_nn[NK];
#pragma unroll
            for (int k = 0; k < NK; ++k)
                code_nn[k] = (validk[k] && t + 2 < Tp) ? lp_clamp_code(__ldg(p.idx + static_cast<int64_t>(rowk[k]) * Tp + t + 2)) : 0;
            if (t > 0) {
                ok = mbar_wait(&tfull_bar[j], tphase, p.err);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tphase ^= 1;
                tc_fence_after();
                if (warp == 2 && lane == 0) LP_STAMP(4);
            }
            float2 hout[NT == 1 ? UW / 8 : 1][NK];
#pragma unroll
            for (int q = 0; q < UW / 8; ++q) {
                const int ug = 8 * q;
                uint32_t v[NK / 2][4][4];                                    // [half][gate][row r8: cols 2cp, 2cp+1 | row r8 + 8: same]
                if (t > 0) {
#pragma unroll
                    for (int hf = 0; hf < NK / 2; ++hf)
#pragma unroll
                        for (int g = 0; g < 4; ++g) tc_ld_16x256b(tb + (static_cast<uint32_t>(16 * (hf + hsel)) << 16) + g * UT + ub + ug, v[hf][g]);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int hf = 0; hf < NK / 2; ++hf)
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[hf][g][i] = 0u;
                }
                float x[NK][4][2];                                       // [row k][gate][unit]
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    x[k][0][0] = xq[k][0].x; x[k][0][1] = xq[k][0].y; x[k][1][0] = xq[k][0].z; x[k][1][1] = xq[k][0].w;
                    x[k][2][0] = xq[k][1].x; x[k][2][1] = xq[k][1].y; x[k][3][0] = xq[k][1].z; x[k][3][1] = xq[k][1].w;
                }
                if (q + 1 < UW / 8) fetch(code, ug + 8);
                else if (t + 1 < Tp) fetch(code_next, 0);
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    float h2[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int hf = k >> 1, vi = 2 * (k & 1) + i;
                        const float gi = x[k][0][i] + __uint_as_float(v[hf][0][vi]), gf = x[k][1][i] + __uint_as_float(v[hf][1][vi]);
                        const float gg = x[k][2][i] + __uint_as_float(v[hf][2][vi]), go = x[k][3][i] + __uint_as_float(v[hf][3][vi]);
                        float& cc = c[q][k][i];
                        if (p.dbg & 4) { cc = 0.5f * cc + gf * gi + gg; h2[i] = go * 0.001f + cc * 0.001f; continue; }
                        h2[i] = lp_cell(gi, gf, gg, go, cc);
                    }
                    if constexpr (NT == 1) hout[q][k] = make_float2(h2[0], h2[1]);
                    else if (validk[k] && !(p.dbg & 2))      // two tiles per CTA: registers are short, the output goes out at once
                        *reinterpret_cast<float2*>(p.out + (static_cast<int64_t>(rowk[k]) * Tp + t) * LP_H + u0 + ub + ug + 2 * cp) = make_float2(h2[0], h2[1]);
                    if (validk[k] && t + 1 < Tp && !(p.dbg & 16)) {
                        __nv_bfloat16* pr = ((t & 1) ? p.planes[1] : p.planes[0]) + static_cast<int64_t>(rowk[k]) * 2 * LP_H + u0 + ub + ug + 2 * cp;
                        const __nv_bfloat16 a = __float2bfloat16_rn(h2[0]), b = __float2bfloat16_rn(h2[1]);
                        const __nv_bfloat162 hh = __halves2bfloat162(a, b);
                        const __nv_bfloat162 ll = __halves2bfloat162(__float2bfloat16_rn(h2[0] - __bfloat162float(a)),
                                                                     __float2bfloat16_rn(h2[1] - __bfloat162float(b)));
                        asm volatile("st.global.b32 [%0], %1;" ::"l"(pr), "r"(*reinterpret_cast<const uint32_t*>(&hh)) : "memory");
                        asm volatile("st.global.b32 [%0], %1;" ::"l"(pr + LP_H), "r"(*reinterpret_cast<const uint32_t*>(&ll)) : "memory");
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NK; ++k) { code[k] = code_next[k]; code_next[k] = code_nn[k]; }
            if (t + 1 < Tp) {
                if (warp == 2 && lane == 0) LP_STAMP(5);
                tc_fence_before();
                if (p.dbg & 64) __threadfence();
                __syncwarp();
                // ONE release per CTA and tile: the tile's epilogue warps meet at a named barrier, then a single thread fences and
                // adds EW to the counter (eight MEMBAR.GPU per CTA and step cost 0.5 us per step: 512 utterances 5.50 -> 4.95 us)
                if (lane == 0) mbar_arrive(&tempty_bar[j]);
                bar_sync(1 + j, EW * 32);
                if ((e % EW) == 0 && lane == 0) red_release_add_u32(p.counters + mt, static_cast<unsigned>(EW));
                if (warp == 2 && lane == 0) LP_STAMP(6);
            }
            // (one tile per CTA) the fp32 output sequence is nobody's input: stored AFTER the release, which then waits for the planes only
            if (NT == 1 && !(p.dbg & 2)) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    if (!validk[k]) continue;
                    float* op = p.out + (static_cast<int64_t>(rowk[k]) * Tp + t) * LP_H + u0 + ub + 2 * cp;
#pragma unroll
                    for (int q = 0; q < UW / 8; ++q) *reinterpret_cast<float2*>(op + 8 * q) = hout[q][k];
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();             // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// table[code][gate * 256 + unit]  ->  perm[code][unit / 8][gate / 2][unit pair][gate % 2][unit % 2]: the 4 gates x 8 units of an
// aligned 8-unit group are ONE 128-byte line (instead of four 32-byte pieces 1 KB apart), ordered so that the four threads that
// share a row in the fragment-layout epilogue read 64 contiguous bytes per instruction
__global__ void lp_permute_table_kernel(const float* __restrict__ table, float* __restrict__ perm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int code = i / LP_G, r = i % LP_G, g = r / LP_H, u = r % LP_H;
    const int o = (g >> 1) * 16 + ((u & 7) >> 1) * 4 + (g & 1) * 2 + (u & 1);
    perm[static_cast<int64_t>(code) * LP_G + (u >> 3) * 32 + o] = table[i];
}

// ------------------------------------------------------------------------------------------------ host
int make_map_bf16(void* map, const void* base, long long rows, long long cols, long long ld_elems, int box_rows);

template <int UT, int NT, int CS, int EW>
static size_t lp_smem() {
    constexpr int N = 4 * UT;
    constexpr int STAGES = UT == 32 ? 3 : 4;
    return 2 * LP_KB * N * TC_BK * 2 + STAGES * 2 * TC_BM * TC_BK * 2 + 1024;
}
template <int UT, int NT, int CS, int EW>
static void lp_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int n_ctas, cudaStream_t stream) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3(n_ctas); cfg->blockDim = dim3(64 + 32 * EW * NT); cfg->dynamicSmemBytes = lp_smem<UT, NT, CS, EW>(); cfg->stream = stream;
    attr[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident, or the launch fails (never a silent hang)
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CS; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg->attrs = attr; cfg->numAttrs = CS > 1 ? 2 : 1;
}
// how many CTAs of this instantiation the device holds at once (0 on error), cached per device
template <int UT, int NT, int CS, int EW>
static int lp_capacity() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cache[dev]) return cache[dev] > 0 ? cache[dev] : 0;
    int cap = 0;
    if (ensure_dyn_smem(reinterpret_cast<const void*>(lstm_persist_kernel<UT, NT, CS, EW>), static_cast<int>(lp_smem<UT, NT, CS, EW>())) == VQCPC_OK) {
        if (CS == 1) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_persist_kernel<UT, NT, CS, EW>, 64 + 32 * EW * NT, lp_smem<UT, NT, CS, EW>()) == cudaSuccess)
                cap = per_sm * device_sm_count();
        } else {
            cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
            lp_config<UT, NT, CS, EW>(&cfg, attr, CS, nullptr);
            cfg.numAttrs = 2; attr[0] = attr[1]; cfg.numAttrs = 1;      // occupancy query: cluster dimension only
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, lstm_persist_kernel<UT, NT, CS, EW>, &cfg) == cudaSuccess) cap = ncl * CS;
        }
    }
    cudaGetLastError();
    cache[dev] = cap > 0 ? cap : -1;
    return cap;
}
template <int UT, int NT, int CS, int EW>
static int lp_launch(const CUtensorMap& h0, const CUtensorMap& h1, const CUtensorMap& mw, LpParams p, int n_mt, cudaStream_t stream) {
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(lstm_persist_kernel<UT, NT, CS, EW>), static_cast<int>(lp_smem<UT, NT, CS, EW>()))) return rc;
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
    lp_config<UT, NT, CS, EW>(&cfg, attr, ((n_mt + NT - 1) / NT) * p.n_ns, stream);
    VQ_CUDA(cudaLaunchKernelEx(&cfg, lstm_persist_kernel<UT, NT, CS, EW>, h0, h1, mw, p));
    count_launch(1);
    return VQCPC_OK;
}

// rows per launch: (row tiles / 2) x 8 unit slices must fit the device
int lstm_persist_max_rows() {
    const int sms = device_sm_count();
    return 2 * (sms / 8) * TC_BM;
}
size_t lstm_persist_counter_bytes() { return 64 * sizeof(unsigned); }
size_t lstm_persist_table_bytes() { return sizeof(float) * 512 * LP_G; }

// planes: two (B, 512) bf16 buffers; counters: >= lstm_persist_counter_bytes(), any content (zeroed here); table_perm:
// lstm_persist_table_bytes() of scratch for the permuted copy of the input-projection table
int lstm_persist(const float* table_in, const int64_t* idx, const void* whh_planes, int B, int Tp, void* planes0, void* planes1,
                 unsigned* counters, float* table_perm, float* out, int* err_flag, cudaStream_t stream) {
    VQ_ARG(table_in && idx && whh_planes && planes0 && planes1 && counters && table_perm && out && err_flag, "lstm_persist: null pointer");
    lp_permute_table_kernel<<<(512 * LP_G + 255) / 256, 256, 0, stream>>>(table_in, table_perm, 512 * LP_G);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    const float* table = table_perm;
    const int max_rows = lstm_persist_max_rows();
    VQ_ARG(max_rows >= TC_BM, "lstm_persist: device too small");
    const int n_chunks = (B + max_rows - 1) / max_rows;
    const int per = ((B + n_chunks - 1) / n_chunks + TC_BM - 1) / TC_BM * TC_BM;
    alignas(64) CUtensorMap mw;
    for (int b0 = 0; b0 < B; b0 += per) {
        const int nb = B - b0 < per ? B - b0 : per;
        const int n_mt = (nb + TC_BM - 1) / TC_BM;
        VQ_ARG(n_mt <= 64, "lstm_persist: too many row tiles");
        // configuration: the most unit slices (least work per CTA) whose grid is co-resident.  The multicast variant (clusters of
        // 4) is opt-in (VQCPC_LSTM_CLUSTER=4): measured on B200 it is no faster (512 utterances: 2.49 vs 2.45 ms per encode, 4096:
        // 12.1 vs 11.9) -- the step is a latency chain (counter hop, TMA first byte, 128 KB of shared-memory ingest per CTA either
        // way, MMA operand reads, epilogue, store visibility), not L2-bandwidth bound.
        static const int lp_cs = [] { const char* e = getenv("VQCPC_LSTM_CLUSTER"); return (e && e[0] == '4') ? 4 : 1; }();
        static const int lp_dbg = [] { const char* e = getenv("VQCPC_LP_DEBUG"); return e ? atoi(e) : 0; }();
        int rc = VQCPC_ERR_ARG;
        bool done = false;
        // epilogue warps per row tile: sixteen (two column groups x the two 16-lane halves of each TMEM lane quarter) with one row
        // tile of 16 or 32 units per CTA, eight (the two halves) with 8-unit slices and with two tiles per CTA (576 threads either
        // way).  VQCPC_LP_EW=1 / 2: four / eight everywhere (A/B; 2048 utterances 7.2 us per step with eight, 6.6 with sixteen;
        // 4096 utterances 10.3 with four, 9.3 with eight).
        static const int lp_ew = [] { const char* e = getenv("VQCPC_LP_EW"); return (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : 3; }();
        auto try_cfg = [&](auto utc, auto ntc, auto csc, auto ewc) {
            constexpr int UT_ = decltype(utc)::value, NT_ = decltype(ntc)::value, CS_ = decltype(csc)::value, EW_ = decltype(ewc)::value;
            if (done || (CS_ > 1) != (lp_cs > 1)) return;
            if (EW_ != (lp_ew == 1 ? 4 : (lp_ew == 3 && NT_ == 1 && UT_ >= 16) ? 16 : 8)) return;   // lp_ew: 1 four, 2 eight, 3 (default) as many as fit
            const int n_ns = LP_H / UT_;
            const int ctas = ((n_mt + NT_ - 1) / NT_) * n_ns;
            if (ctas > lp_capacity<UT_, NT_, CS_, EW_>()) return;
            alignas(64) CUtensorMap h0, h1;
            __nv_bfloat16* p0 = static_cast<__nv_bfloat16*>(planes0) + static_cast<int64_t>(b0) * 2 * LP_H;
            __nv_bfloat16* p1 = static_cast<__nv_bfloat16*>(planes1) + static_cast<int64_t>(b0) * 2 * LP_H;
            if ((rc = make_map_bf16(&h0, p0, nb, 2LL * LP_H, 2LL * LP_H, TC_BM / CS_))) { done = true; return; }
            if ((rc = make_map_bf16(&h1, p1, nb, 2LL * LP_H, 2LL * LP_H, TC_BM / CS_))) { done = true; return; }
            if ((rc = make_map_bf16(&mw, whh_planes, LP_G, 2LL * LP_H, 2LL * LP_H, UT_))) { done = true; return; }
            if (cudaMemsetAsync(counters, 0, lstm_persist_counter_bytes(), stream) != cudaSuccess) { rc = VQCPC_ERR_CUDA; done = true; return; }
            LpParams p{};
            p.table = table; p.idx = idx + static_cast<int64_t>(b0) * Tp; p.out = out + static_cast<int64_t>(b0) * Tp * LP_H;
            p.planes[0] = p0; p.planes[1] = p1; p.counters = counters; p.err = err_flag; p.B = nb; p.Tp = Tp; p.n_ns = n_ns; p.dbg = lp_dbg;
            rc = lp_launch<UT_, NT_, CS_, EW_>(h0, h1, mw, p, n_mt, stream);
            done = true;
        };
        using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>; using I4 = std::integral_constant<int, 4>;
        using U8 = std::integral_constant<int, 8>; using U16 = std::integral_constant<int, 16>; using U32 = std::integral_constant<int, 32>;
        using E4 = std::integral_constant<int, 4>; using E8 = std::integral_constant<int, 8>;
        auto try_ut = [&](auto utc, auto ntc, auto csc) {
            try_cfg(utc, ntc, csc, E4{});
            try_cfg(utc, ntc, csc, E8{});
            if constexpr (decltype(ntc)::value == 1 && decltype(utc)::value >= 16 && decltype(csc)::value == 1) try_cfg(utc, ntc, csc, std::integral_constant<int, 16>{});
        };
        try_ut(U8{}, I1{}, I4{});  try_ut(U16{}, I1{}, I4{}); try_ut(U32{}, I1{}, I4{}); try_ut(U32{}, I2{}, I4{});
        try_ut(U8{}, I1{}, I1{});  try_ut(U16{}, I1{}, I1{}); try_ut(U32{}, I1{}, I1{}); try_ut(U32{}, I2{}, I1{});
        if (!done) { set_error("lstm_persist: no configuration fits the device (%d row tiles)", n_mt); return VQCPC_ERR_CUDA; }
        if (rc) return rc;
        if (lp_dbg & 32) {
            long long tr[64];
            cudaStreamSynchronize(stream);
            if (cudaMemcpyFromSymbol(tr, lp_trace, sizeof(tr)) == cudaSuccess)
                for (int s = 0; s + 1 < 8; ++s)
                    fprintf(stderr, "lp step %d: counter ok 0 | TMA issued %lld | first stage full %lld | last stage full %lld | accumulator full %lld | "
                            "epilogue done %lld | published %lld | next counter ok %lld\n", 60 + s, tr[s * 8 + 1] - tr[s * 8], tr[s * 8 + 2] - tr[s * 8],
                            tr[s * 8 + 3] - tr[s * 8], tr[s * 8 + 4] - tr[s * 8], tr[s * 8 + 5] - tr[s * 8], tr[s * 8 + 6] - tr[s * 8],
                            tr[(s + 1) * 8] - tr[s * 8]);
        }
    }
    return VQCPC_OK;
}

}  // namespace vqcpc
