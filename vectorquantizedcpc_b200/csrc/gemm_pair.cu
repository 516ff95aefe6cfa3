// Fused  relu(LayerNorm(A . W^T))  layer of the encoder MLP (/root/reference/model.py:43-55,65-68) on CTA PAIRS:
// tcgen05.mma.cta_group::2 (M = 256 rows per pair, 128 per CTA; N = 256-column tiles; fp32 accumulators in TMEM).
//
// Why pairs: the 128-row single-CTA kernel (gemm_tc_ln_kernel) re-reads the whole weight matrix per 128 rows and every A tile
// once per segment -- 17 GB of L2 -> SM operand traffic per 614 400 x 768 x 768 layer, which bound it.  Here
//   * a pair shares every W tile (each CTA loads HALF of it; the MMA reads both halves)  -> W traffic / 2,
//   * a stage holds A_hi, A_lo, W_hi, W_lo of one 64-wide k-block ONCE and the three products of the hi/lo split
//     (A_hi W_hi + A_hi W_lo + A_lo W_hi) are issued from it                             -> tile loads 4 instead of 6,
// i.e. 7.6 GB per layer.  NSEG = 1 is the single-pass bf16 mode (hi planes only).
//
// Roles per CTA (320 threads): warp 0 TMA producer (own A rows, own half of the W tile; all transaction bytes complete on the
// LEADER's full barrier), warp 1 MMA issuer (leader CTA only; commits multicast to both CTAs' barriers), warps 2-9 epilogue,
// thread = row (TMEM lane), two warps per lane quarter (128 of a tile's 256 columns each):
//   pass A, per column tile as it completes: shifted row sums straight from TMEM.  The LAST TWO tiles of a row block stay in the
//           two TMEM accumulators; earlier tiles (tile 0 of N = 768) are parked in a per-warp scratch (thread-private layout,
//           128 KB per CTA, L2-resident) and their accumulator is handed back to the MMA warp;
//   pass B, once the statistics are complete (the two column halves of a row meet in shared memory): normalise, affine, ReLU
//           and the bf16 hi/lo split from TMEM (tile NT-2 first: its accumulator is the one the next row block needs), then
//           from the scratch.  Output leaves through shared memory as 32 x 32 bf16 boxes (SWIZZLE_64B, conflict-free 16-byte
//           stores) and TMA stores (cp.async.bulk.tensor .. bulk_group) -- no uncoalesced global stores, no fp32 round trip.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace vqcpc {

constexpr int PR_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

struct PairParams {
    const float* ln_w; const float* ln_b;
    __nv_bfloat16* out_planes;   // (M, out_pitch): [hi | lo] (out_pitch = 2N) or hi only (out_pitch = N)
    float* out_f32;              // (M, N) or null
    float* scratch;              // gridDim.x * 128 * N floats
    int* err;
    int M, K;
    int out_pitch;               // elements per output row
    int write_lo;                // 1: write the lo plane at column offset N
    int a_lo_col, w_lo_col;      // column offset of the lo plane inside the A / W plane rows (NSEG = 3)
    int debug;                   // VQCPC_PAIR_DEBUG bit 0: no L2 cache hints, bit 1: no evict_first on the output stores
};

template <int NT, int NSEG>
// (10 warps: one scheduler hosts three of them, so 16384 / 3 / 32 = 170 registers per thread is the ceiling)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PR_THREADS, 1)
gemm_ln_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_out, PairParams p) {
    constexpr int BN = 256;                                   // columns per tile (pair-wide); each CTA stages 128 W rows of it
    constexpr int N = NT * BN;
    constexpr uint32_t TILE_BYTES = TC_BM * TC_BK * 2;        // 16 KB: 128 rows x 64 bf16
    constexpr int NPL = NSEG == 3 ? 2 : 1;                    // planes per operand in a stage
    constexpr uint32_t STAGE_BYTES = 2 * NPL * TILE_BYTES;    // 64 KB (hi/lo) or 32 KB (bf16)
    constexpr int STAGES = NSEG == 3 ? 3 : 6;
    constexpr uint32_t TMEM_COLS = 512;
    // instruction descriptor: D = F32, A = B = BF16, K-major, N = 256, M = 256 (pair)
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                               (static_cast<uint32_t>(256 >> 4) << 24);

    extern __shared__ __align__(1024) unsigned char pr_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pr_smem) + 1023) & ~uintptr_t(1023));
    unsigned char* stage_out = smem + STAGES * STAGE_BYTES;     // 8 x 4 KB: per epilogue warp the bf16 hi / lo boxes of a TMA store
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 16); }   // 8 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers are initialised and both have their TMEM before anyone signals
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int n_pblocks = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int n_kb = p.K / TC_BK;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            const uint32_t full0 = mapa_u32(smem_u32(&full_bar[0]), 0);          // the leader's full barriers
            // L2 residency: a row block's A tiles are fetched once per column tile -- keep them until the last one; W is hot
            const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
            const bool hints = !(p.debug & 1);
            for (int pb = pair; pb < n_pblocks && ok; pb += n_pairs) {
                const int row0 = pb * 2 * TC_BM + static_cast<int>(rank) * TC_BM;
                for (int tn = 0; tn < NT && ok; ++tn) {
                    const int wrow0 = tn * BN + static_cast<int>(rank) * (BN / 2);
                    for (int kb = 0; kb < n_kb && ok; ++kb) {
                        ok = mbar_wait(&empty_bar[stage], phase ^ 1, p.err);
                        if (!ok) break;
                        unsigned char* st = smem + stage * STAGE_BYTES;
                        const uint32_t fb = full0 + static_cast<uint32_t>(stage) * 8u;
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);     // both CTAs' bytes
                        const int kk = kb * TC_BK;
                        if (hints) {
                            const uint64_t pa = tn == NT - 1 ? pol_stream : pol_keep;
                            tma_load_2d_pair_hint(st, &map_a, kk, row0, fb, pa);
                            if (NSEG == 3) tma_load_2d_pair_hint(st + TILE_BYTES, &map_a, p.a_lo_col + kk, row0, fb, pa);
                            tma_load_2d_pair_hint(st + NPL * TILE_BYTES, &map_w, kk, wrow0, fb, pol_keep);
                            if (NSEG == 3) tma_load_2d_pair_hint(st + 3 * TILE_BYTES, &map_w, p.w_lo_col + kk, wrow0, fb, pol_keep);
                        } else {
                            tma_load_2d_pair(st, &map_a, kk, row0, fb);
                            if (NSEG == 3) tma_load_2d_pair(st + TILE_BYTES, &map_a, p.a_lo_col + kk, row0, fb);
                            tma_load_2d_pair(st + NPL * TILE_BYTES, &map_w, kk, wrow0, fb);
                            if (NSEG == 3) tma_load_2d_pair(st + 3 * TILE_BYTES, &map_w, p.w_lo_col + kk, wrow0, fb);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0, 0};
            bool ok = true;
            for (int pb = pair; pb < n_pblocks && ok; pb += n_pairs) {
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
                        const uint64_t a_hi = umma_desc_sw128(sa), w_hi = umma_desc_sw128(sa + NPL * TILE_BYTES);
                        const uint64_t a_lo = umma_desc_sw128(sa + TILE_BYTES), w_lo = umma_desc_sw128(sa + 3 * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k) {
                            tc_mma_f16_pair(d_tmem, a_hi + 2 * k, w_hi + 2 * k, IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                            if (NSEG == 3) {
                                tc_mma_f16_pair(d_tmem, a_hi + 2 * k, w_lo + 2 * k, IDESC, 1u);
                                tc_mma_f16_pair(d_tmem, a_lo + 2 * k, w_hi + 2 * k, IDESC, 1u);
                            }
                        }
                        tc_commit_pair(&empty_bar[stage]);           // frees the stage in BOTH CTAs
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit_pair(&tfull_bar[acc]);                 // accumulator ready, both CTAs' epilogues
                    acc_phase[acc] ^= 1;
                    acc ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (each CTA: its own 128 rows)
        // thread = row throughout (TMEM lane).  Warp (quarter, half) owns rows 32 quarter .. +31 and, of every 256-column tile,
        // columns 128 half .. +127.
        const int ew = warp - 2;
        const int quarter = warp & 3, half = ew >> 2;
        int acc = 0;
        uint32_t acc_phase[2] = {0, 0};
        bool ok = true;
        const uint32_t tempty0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);               // the leader's accumulator-free barriers
        unsigned char* stg = stage_out + ew * 4096;                                    // [hi 2 KB | lo 2 KB], SWIZZLE_64B boxes
        const uint32_t stg_u32 = smem_u32(stg);
        const uint32_t swz = static_cast<uint32_t>((lane >> 1) & 3);
        float4* sc = reinterpret_cast<float4*>(p.scratch) +
                     (static_cast<size_t>(blockIdx.x) * 8 + ew) * (NT > 2 ? NT - 2 : 1) * 32 * 32 + lane;
        const uint64_t pol_out = l2_policy_evict_first();
        constexpr float NH = static_cast<float>(NT * 128);                              // columns per half-row
        for (int pb = pair; pb < n_pblocks && ok; pb += n_pairs) {
            const int row_base = pb * 2 * TC_BM + static_cast<int>(rank) * TC_BM;
            const int row_w = row_base + quarter * 32;                                 // first row of this warp
            const int row = row_w + lane;
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            int tile_acc[NT];
            // ---- pass A: statistics (tiles 0 .. NT-3 also leave TMEM for the scratch, freeing their accumulator)
#pragma unroll
            for (int tn = 0; tn < NT; ++tn) {
                if (ok) ok = mbar_wait(&tfull_bar[acc], acc_phase[acc], p.err);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * 128;
                // software pipelined: the tcgen05.ld of chunk c+1 is in flight while chunk c is summed
                uint32_t vbuf[2][32];
                tc_ld32(taddr, vbuf[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t (&v)[32] = vbuf[c & 1];
                    tc_wait_ld();
                    if (c < 3) tc_ld32(taddr + 32 * (c + 1), vbuf[(c + 1) & 1]);
                    if (tn == 0 && c == 0) shift = __uint_as_float(v[0]);
                    if (!(p.debug & 16))
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = __uint_as_float(v[j]) - shift;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                    if (tn < NT - 2) {
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4)
                            __stcg(&sc[((tn * 4 + c) * 8 + q4) * 32],                   // L2 only: L1 keeps the LayerNorm weights
                                   make_float4(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]), __uint_as_float(v[4 * q4 + 2]),
                                               __uint_as_float(v[4 * q4 + 3])));
                    }
                }
                tile_acc[tn] = acc;
                if (tn < NT - 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + static_cast<uint32_t>(acc) * 8u);
                }
                acc_phase[acc] ^= 1;
                acc ^= 1;
            }
            if (!ok) break;
            // ---- the row's statistics: this thread has 128 NT of its N columns, the partner warp's thread the rest.  The two
            //      exchange (mean, M2) through the first 256 bytes of their staging buffers (idle between row blocks).
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            {
                const float mh = s1 * (1.0f / NH);
                reinterpret_cast<float2*>(stg)[lane] = make_float2(shift + mh, fmaxf(s2 - s1 * mh, 0.f));
            }
            bar_sync(1 + quarter, 64);
            float mean, rstd;
            {
                const float2 a = reinterpret_cast<const float2*>(stg)[lane];
                const float2 b = reinterpret_cast<const float2*>(stage_out + (ew ^ 4) * 4096)[lane];
                const float dm = a.x - b.x;
                mean = 0.5f * (a.x + b.x);
                const float var = (a.y + b.y + dm * dm * (0.5f * NH)) * (1.0f / N);
                rstd = 1.0f / sqrtf(var + 1e-5f);
            }
            bar_sync(1 + quarter, 64);                                               // both have read before either stages output
            const float nmr = -mean * rstd;
            // normalise + affine + ReLU + split of 32 columns starting at global column gc; TMA store of the two bf16 boxes
            // (the LayerNorm weights of the 32 columns are warp-uniform L1 hits: all 16 loads are issued by the caller BEFORE it
            // waits for the TMEM / scratch data, so their latency is paid once per chunk, not once per use)
            auto emit = [&](const uint32_t (&v)[32], const float4 (&w4)[8], const float4 (&b4)[8], int gc) {
                if (p.debug & 32) return;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float o0 = fmaxf(fmaf(fmaf(__uint_as_float(v[4 * j4]), rstd, nmr), w4[j4].x, b4[j4].x), 0.f);
                    float o1 = fmaxf(fmaf(fmaf(__uint_as_float(v[4 * j4 + 1]), rstd, nmr), w4[j4].y, b4[j4].y), 0.f);
                    float o2 = fmaxf(fmaf(fmaf(__uint_as_float(v[4 * j4 + 2]), rstd, nmr), w4[j4].z, b4[j4].z), 0.f);
                    float o3 = fmaxf(fmaf(fmaf(__uint_as_float(v[4 * j4 + 3]), rstd, nmr), w4[j4].w, b4[j4].w), 0.f);
                    if (p.out_f32 != nullptr && row < p.M)
                        __stcs(reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(row) * N + gc) + j4, make_float4(o0, o1, o2, o3));
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
                    const uint32_t hb0 = *reinterpret_cast<const uint32_t*>(&h0), hb1 = *reinterpret_cast<const uint32_t*>(&h1);
                    hi[2 * j4] = hb0; hi[2 * j4 + 1] = hb1;
                    if (p.write_lo) {
                        const __nv_bfloat162 l0 = __floats2bfloat162_rn(o0 - __uint_as_float(hb0 << 16), o1 - __uint_as_float(hb0 & 0xffff0000u));
                        const __nv_bfloat162 l1 = __floats2bfloat162_rn(o2 - __uint_as_float(hb1 << 16), o3 - __uint_as_float(hb1 & 0xffff0000u));
                        lo[2 * j4] = *reinterpret_cast<const uint32_t*>(&l0); lo[2 * j4 + 1] = *reinterpret_cast<const uint32_t*>(&l1);
                    }
                }
                if (lane == 0) bulk_wait_read0();                 // the previous boxes have left the staging buffer
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {                     // row = lane: 64 B per plane, 16-byte chunk j at (j ^ swz)
                    const uint32_t off = static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(j) ^ swz) << 4);
                    st_shared_v4(stg_u32 + off, hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    if (p.write_lo) st_shared_v4(stg_u32 + 2048u + off, lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(p.debug & 8)) {
                    if (p.debug & 2) {
                        tma_store_2d(&map_out, stg, gc, row_w);
                        if (p.write_lo) tma_store_2d(&map_out, stg + 2048, N + gc, row_w);
                    } else {
                        tma_store_2d_hint(&map_out, stg, gc, row_w, pol_out);
                        if (p.write_lo) tma_store_2d_hint(&map_out, stg + 2048, N + gc, row_w, pol_out);
                    }
                    bulk_commit();
                }
            };
            // ---- pass B: the two tiles still in TMEM -- tile NT-2 first (the next row block's first tile waits for its
            //      accumulator), then tile NT-1 --, then the tiles parked in the scratch
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int tn = (NT >= 2) ? (k == 0 ? NT - 2 : NT - 1) : 0;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + tile_acc[tn] * BN + half * 128;
                uint32_t vbuf[2][32];
                tc_ld32(taddr, vbuf[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 w4[8], b4[8];
                    const int gc = tn * BN + half * 128 + 32 * c;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        if (p.debug & 4) { w4[j4] = make_float4(1.f, 1.f, 1.f, 1.f); b4[j4] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
                        w4[j4] = __ldg(reinterpret_cast<const float4*>(p.ln_w + gc) + j4);
                        b4[j4] = __ldg(reinterpret_cast<const float4*>(p.ln_b + gc) + j4);
                    }
                    tc_wait_ld();
                    if (c < 3) tc_ld32(taddr + 32 * (c + 1), vbuf[(c + 1) & 1]);      // in flight while chunk c is processed
                    emit(vbuf[c & 1], w4, b4, gc);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(tempty0 + static_cast<uint32_t>(tile_acc[tn]) * 8u);
            }
#pragma unroll 1
            for (int g = 0; g < (NT - 2) * 4; ++g) {
                uint32_t v[32];
                float4 w4[8], b4[8], y4[8];
                const int gc = (g >> 2) * BN + half * 128 + (g & 3) * 32;
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) y4[q4] = __ldcg(&sc[(g * 8 + q4) * 32]);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    w4[j4] = __ldg(reinterpret_cast<const float4*>(p.ln_w + gc) + j4);
                    b4[j4] = __ldg(reinterpret_cast<const float4*>(p.ln_b + gc) + j4);
                }
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    v[4 * q4] = __float_as_uint(y4[q4].x); v[4 * q4 + 1] = __float_as_uint(y4[q4].y);
                    v[4 * q4 + 2] = __float_as_uint(y4[q4].z); v[4 * q4 + 3] = __float_as_uint(y4[q4].w);
                }
                emit(v, w4, b4, gc);
            }
        }
        if (lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still signal this CTA's barriers / read its shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
int make_map_bf16(void* map, const void* base, long long rows, long long cols, long long ld_elems, int box_rows);   // gemm_tc.cu
int make_map_bf16_box(void* map, const void* base, long long rows, long long cols, long long ld_elems, int box_cols, int box_rows,
                      int swizzle_bytes);

// per CTA: the tiles that cannot stay in TMEM until the row statistics are complete (all but the last two), 128 x 256 fp32 each
size_t gemm_ln_pair_scratch_bytes(int N) { return static_cast<size_t>(148) * TC_BM * (N > 512 ? N - 512 : 256) * sizeof(float); }
bool gemm_ln_pair_supported(int N) { return N == 512 || N == 768; }

template <int NT, int NSEG>
static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mo, const PairParams& p, cudaStream_t stream) {
    constexpr size_t smem = (NSEG == 3 ? 3 * 64 : 6 * 32) * 1024 + 8 * 4096 + 1024;
    if (int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(gemm_ln_pair_kernel<NT, NSEG>), static_cast<int>(smem))) return rc_attr;
    const int pblocks = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
    int pairs = device_sm_count() / 2;
    if (pairs > 74) pairs = 74;
    if (pblocks < pairs) pairs = pblocks;
    gemm_ln_pair_kernel<NT, NSEG><<<2 * pairs, PR_THREADS, smem, stream>>>(ma, mw, mo, p);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}

// a_planes: (M, a_pitch) bf16 with the hi plane in columns [0, K) and (nseg = 3) the lo plane in [a_lo_col, a_lo_col + K);
// w_planes: (N, 2K) [hi | lo].  out_planes: (M, out_pitch); out_pitch = 2N writes [hi | lo], out_pitch = N hi only.
int gemm_ln_pair(const void* a_planes, long long a_pitch, int a_lo_col, const void* w_planes, const float* ln_w, const float* ln_b,
                 void* out_planes, int out_pitch, float* out_f32, void* scratch, int M, int N, int K, int nseg, int* err_flag,
                 cudaStream_t stream) {
    if (M == 0) return VQCPC_OK;
    VQ_ARG(a_planes && w_planes && ln_w && ln_b && out_planes && scratch && err_flag, "gemm_ln_pair: null pointer");
    VQ_ARG(gemm_ln_pair_supported(N), "gemm_ln_pair: N=%d must be 512 or 768", N);
    VQ_ARG(K % TC_BK == 0 && K > 0 && (nseg == 1 || nseg == 3), "gemm_ln_pair: bad K / nseg");
    VQ_ARG(out_pitch == N || out_pitch == 2 * N, "gemm_ln_pair: out_pitch must be N or 2N");
    VQ_ARG(a_pitch % 8 == 0 && (nseg == 1 || a_lo_col + K <= a_pitch), "gemm_ln_pair: bad A plane layout");
    CUtensorMap ma, mw, mo;
    int rc = make_map_bf16(&ma, a_planes, M, a_pitch, a_pitch, TC_BM);
    if (rc) return rc;
    rc = make_map_bf16(&mw, w_planes, N, 2LL * K, 2LL * K, TC_BM);
    if (rc) return rc;
    rc = make_map_bf16_box(&mo, out_planes, M, out_pitch, out_pitch, 32, 32, 64);     // 32 x 32 bf16 boxes, SWIZZLE_64B
    if (rc) return rc;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VQCPC_PAIR_DEBUG"); dbg = e ? atoi(e) : 0; }
    PairParams p{ln_w, ln_b, static_cast<__nv_bfloat16*>(out_planes), out_f32, static_cast<float*>(scratch), err_flag, M, K,
                 out_pitch, out_pitch == 2 * N ? 1 : 0, a_lo_col, K, dbg};
    if (N == 768) return nseg == 3 ? launch_pair<3, 3>(ma, mw, mo, p, stream) : launch_pair<3, 1>(ma, mw, mo, p, stream);
    return nseg == 3 ? launch_pair<2, 3>(ma, mw, mo, p, stream) : launch_pair<2, 1>(ma, mw, mo, p, stream);
}

}  // namespace vqcpc
