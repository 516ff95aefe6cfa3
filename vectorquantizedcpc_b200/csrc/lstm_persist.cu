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
//     thread issues tcgen05.mma (M = 128, N = 4 UT, K = 16; the three hi/lo products per k-block from one stage), the four
//     epilogue warps (thread = utterance) read the accumulator from TMEM, add the gathered table row (prefetched during the
//     wait), apply the cell with the cell state in REGISTERS for the whole sequence, and write h_t as fp32 into the output
//     sequence and as bf16 hi/lo planes into the other plane buffer;
//   * the CTAs of a row tile (and only those: utterances are independent) meet at a release/acquire counter before the next
//     step's TMA loads -- no grid-wide barrier, no host involvement;
//   * every CTA of a row tile needs the SAME 128 KB of h_{t-1}: clusters of 4 CTAs (neighbouring unit slices) load a quarter
//     each and TMA-multicast it to all four (the L2 -> SM traffic, 16 MB per step at 512 utterances, was what bound the step).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vqcpc {

// threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue of row tile 0 (6-9: row tile 1)
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
    constexpr int UW = UT / (EW / 4);                           // hidden units per epilogue warp
    constexpr int CH = (UW >= 8 && 64 + 32 * EW * NT <= 384) ? 8 : 4;   // units per chunk (register budget)
    static_assert(EW % 4 == 0 && UW >= 4 && UW % CH == 0, "epilogue split");
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
        // ------------------------------------------------------------------ epilogue: thread = utterance (TMEM lane)
        // EW warps per row tile: warp = (TMEM lane quarter, column group); a column group is UW = UT / (EW / 4) hidden units,
        // worked through in chunks of CH units.  With ONE warp per scheduler the epilogue was a serial latency chain (TMEM read ->
        // table lines -> ten MUFU-deep gate math -> stores: 3 500 cycles per 8 units, 14 300 per step at UT = 32); several
        // warps per scheduler overlap those chains.
        const int quarter = warp & 3;                           // TMEM lanes 32 quarter .. +31 (a warp may only touch its own quarter)
        const int e = warp - 2;
        const int j = e / EW;                                   // row tile of this warp
        const int ub = ((e % EW) >> 2) * UW;                    // first unit (within the CTA's UT) of this warp's column group
        const int mt = mt0 + j;
        const int row = mt * TC_BM + quarter * 32 + lane;
        const bool valid = row < p.B;
        const bool tile_live = mt * TC_BM < p.B;
        const uint32_t tb = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + j * N;
        const int64_t* irow = p.idx + static_cast<int64_t>(valid ? row : 0) * Tp;
        float c[UW], hout[UW];
#pragma unroll
        for (int k = 0; k < UW; ++k) c[k] = 0.f;
        uint32_t tphase = 0;
        bool ok = true;
        int code = valid ? lp_clamp_code(__ldg(irow)) : 0;
        // table values of the first chunk of step 0 (every later chunk is prefetched one chunk / one step ahead)
        float4 xq[4][CH / 4];
        auto fetch = [&](int cd, int ug) {
            // permuted table (lp_permute_table_kernel): the 4 gates x 8 units of an aligned 8-unit group are ONE 128-byte line
            const int u = u0 + ub + ug;
            const float4* tr = reinterpret_cast<const float4*>(p.table + static_cast<int64_t>((p.dbg & 1) ? 0 : cd) * LP_G + (u >> 3) * 32 + (u & 7));
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int q = 0; q < CH / 4; ++q) xq[g][q] = __ldg(tr + 2 * g + q);
        };
        fetch(code, 0);
        for (int t = 0; t < Tp && ok && tile_live; ++t) {
            const int code_next = (valid && t + 1 < Tp) ? lp_clamp_code(__ldg(irow + t + 1)) : 0;
            if (t > 0) {
                ok = mbar_wait(&tfull_bar[j], tphase, p.err);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tphase ^= 1;
                tc_fence_after();
                if (warp == 2 && lane == 0) LP_STAMP(4);
            }
            float* op = p.out + (static_cast<int64_t>(valid ? row : 0) * Tp + t) * LP_H + u0 + ub;
            __nv_bfloat16* pr = p.planes[t & 1] + static_cast<int64_t>(valid ? row : 0) * 2 * LP_H + u0 + ub;
#pragma unroll
            for (int ug = 0; ug < UW; ug += CH) {
                uint32_t v[4][CH];
                if (t > 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if constexpr (CH == 8) tc_ld8(tb + g * UT + ub + ug, v[g]); else tc_ld4(tb + g * UT + ub + ug, v[g]);
                    }
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
#pragma unroll
                        for (int k = 0; k < CH; ++k) v[g][k] = 0u;
                }
                float x[4][CH];
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        x[g][4 * q + 0] = xq[g][q].x; x[g][4 * q + 1] = xq[g][q].y; x[g][4 * q + 2] = xq[g][q].z; x[g][4 * q + 3] = xq[g][q].w;
                    }
                // prefetch the next chunk (of this step, or chunk 0 of the next step) while this one is computed
                if (ug + CH < UW) fetch(code, ug + CH);
                else if (t + 1 < Tp) fetch(code_next, 0);
                float h[CH];
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    // same expression order as lstm_gate_kernel: table + gates, then the cell
                    const float gi = x[0][k] + __uint_as_float(v[0][k]), gf = x[1][k] + __uint_as_float(v[1][k]);
                    const float gg = x[2][k] + __uint_as_float(v[2][k]), go = x[3][k] + __uint_as_float(v[3][k]);
                    if (p.dbg & 4) { c[ug + k] = 0.5f * c[ug + k] + gf * gi + gg; h[k] = go * 0.001f + c[ug + k] * 0.001f; continue; }
                    c[ug + k] = sigmoid_fast(gf) * c[ug + k] + sigmoid_fast(gi) * tanh_fast(gg);
                    h[k] = sigmoid_fast(go) * tanh_fast(c[ug + k]);
                }
#pragma unroll
                for (int k = 0; k < CH; ++k) hout[ug + k] = h[k];
                if (valid) {
                    if (t + 1 < Tp && !(p.dbg & 16)) {
                        // hi / lo planes of the chunk: one 16-byte (8 units) or 8-byte (4 units) store each; same rounding as tc_split_store4
                        uint32_t ph[CH / 2], pl[CH / 2];
#pragma unroll
                        for (int k = 0; k < CH / 2; ++k) {
                            const __nv_bfloat16 a = __float2bfloat16_rn(h[2 * k]), b = __float2bfloat16_rn(h[2 * k + 1]);
                            const __nv_bfloat162 hh = __halves2bfloat162(a, b);
                            const __nv_bfloat162 ll = __halves2bfloat162(__float2bfloat16_rn(h[2 * k] - __bfloat162float(a)),
                                                                         __float2bfloat16_rn(h[2 * k + 1] - __bfloat162float(b)));
                            ph[k] = *reinterpret_cast<const uint32_t*>(&hh);
                            pl[k] = *reinterpret_cast<const uint32_t*>(&ll);
                        }
                        if constexpr (CH == 8) {
                            *reinterpret_cast<uint4*>(pr + ug) = make_uint4(ph[0], ph[1], ph[CH / 2 - 2], ph[CH / 2 - 1]);
                            *reinterpret_cast<uint4*>(pr + LP_H + ug) = make_uint4(pl[0], pl[1], pl[CH / 2 - 2], pl[CH / 2 - 1]);
                        } else {
                            *reinterpret_cast<uint2*>(pr + ug) = make_uint2(ph[0], ph[1]);
                            *reinterpret_cast<uint2*>(pr + LP_H + ug) = make_uint2(pl[0], pl[1]);
                        }
                    }
                }
            }
            code = code_next;
            if (t + 1 < Tp) {
                // this warp's h_t is written: publish and hand the accumulator back.  The lanes' stores are ordered before lane
                // 0's gpu-scope release by the warp barrier (the pattern of a grid barrier: block barrier, then ONE thread fences);
                // a fence in every lane (VQCPC_LP_DEBUG & 64) cost ~1 000 cycles more per step
                if (warp == 2 && lane == 0) LP_STAMP(5);
                tc_fence_before();
                if (p.dbg & 64) __threadfence();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&tempty_bar[j]);
                    red_release_add_u32(p.counters + mt, 1u);              // EW arrivals per CTA and step
                }
                if (warp == 2 && lane == 0) LP_STAMP(6);
            }
            // the fp32 output sequence is nobody's input: stored AFTER the release, so that the release waits only for the planes
            if (valid && !(p.dbg & 2)) {
#pragma unroll
                for (int q = 0; q < UW / 4; ++q)
                    *reinterpret_cast<float4*>(op + 4 * q) = make_float4(hout[4 * q], hout[4 * q + 1], hout[4 * q + 2], hout[4 * q + 3]);
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

// table[code][gate * 256 + unit]  ->  perm[code][unit / 8][gate][unit % 8]: what a thread of the epilogue needs for one
// 8-unit chunk (4 gates x 8 units) is one aligned 128-byte line instead of four 32-byte pieces 1 KB apart
__global__ void lp_permute_table_kernel(const float* __restrict__ table, float* __restrict__ perm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int code = i / LP_G, r = i % LP_G, g = r / LP_H, u = r % LP_H;
    perm[static_cast<int64_t>(code) * LP_G + (u >> 3) * 32 + g * 8 + (u & 7)] = table[i];
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
        // epilogue warps per row tile: eight (two per scheduler) with one row tile per CTA, four with two (the two tiles' epilogues
        // already overlap).  Measured on random codes, us/step at 512 / 1024 / 2048 / 4096 utterances: four warps 6.25 / 8.27 / 12.9 /
        // 16.6, eight 6.11 / 7.32 / 11.0 / 17.5, sixteen - / 7.91 / 11.4 / -.  VQCPC_LP_EW=1 forces four everywhere (A/B).
        static const int lp_ew = [] { const char* e = getenv("VQCPC_LP_EW"); return (e && e[0] == '1') ? 1 : 2; }();
        auto try_cfg = [&](auto utc, auto ntc, auto csc, auto ewc) {
            constexpr int UT_ = decltype(utc)::value, NT_ = decltype(ntc)::value, CS_ = decltype(csc)::value, EW_ = decltype(ewc)::value;
            if (done || (CS_ > 1) != (lp_cs > 1)) return;
            if (EW_ != ((NT_ == 2 || lp_ew == 1) ? 4 : 8)) return;
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
            constexpr int NT_ = decltype(ntc)::value;
            try_cfg(utc, ntc, csc, E4{});
            if constexpr (NT_ == 1) try_cfg(utc, ntc, csc, E8{});
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
