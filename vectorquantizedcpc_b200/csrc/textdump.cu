// Text dump of encode.py (/root/reference/encode.py:48-52, 57-67: np.savetxt(file, z, fmt="%.16f") of z, c and the pre-VQ
// auxiliary embedding) on the GPU -- SURVEY.md 8f row 3: once encode takes a fraction of a millisecond, numpy's per-value
// "%.16f" % float(v) formatting (~1 us per value) is what the wall clock of encode.py is made of.
//
// Byte work, bit-exact: every fp32 value v = m 2^e has a finite decimal expansion, and "%.16f" is that expansion rounded to 16
// fractional digits, ties to even (glibc / CPython repr arithmetic).  With integer arithmetic only:
//   fraction F = k / 2^q  (k < 2^24):   F 10^16 = k 5^16 2^(16-q),   k 5^16 < 2^62 fits 64 bits; the right shift by q - 16 carries
//   the remainder that decides the rounding (> half: up; == half: to even; a carry out of 10^16 goes into the integer part);
//   integer part m 2^e (e >= 0) fits 128 bits (fp32 max < 2^128), converted by repeated division.
// Layout = np.savetxt's: values of a row separated by ' ', every row ended by '\n'.  Token lengths differ (sign, integer digits,
// nan / inf), so the dump is length pass -> exclusive scan over 256-value blocks -> write pass.
#include "common.cuh"
#include "kernels.cuh"

namespace vqcpc {

constexpr int TD_BLOCK = 256;
constexpr unsigned long long TD_5P16 = 152587890625ULL;              // 5^16
constexpr unsigned long long TD_10P16 = 10000000000000000ULL;        // 10^16

struct TdValue {
    unsigned __int128 ipart;      // integer part of |v|
    unsigned long long frac;      // 16 fractional digits as an integer in [0, 10^16)
    int special;                  // 0 finite, 1 nan, 2 inf
    bool neg;
};

__device__ __forceinline__ TdValue td_decompose(float v) {
    TdValue d;
    const unsigned bits = __float_as_uint(v);
    d.neg = (bits >> 31) != 0;
    const unsigned ex = (bits >> 23) & 0xffu, man = bits & 0x7fffffu;
    d.ipart = 0; d.frac = 0; d.special = 0;
    if (ex == 0xffu) { d.special = man ? 1 : 2; return d; }
    const unsigned m = ex ? (man | 0x800000u) : man;
    const int e = ex ? static_cast<int>(ex) - 150 : -149;
    if (e >= 0) { d.ipart = static_cast<unsigned __int128>(m) << e; return d; }
    const int q = -e;                                                  // 1 .. 149
    unsigned k = m;
    if (q < 32) { d.ipart = m >> q; k = m & ((1u << q) - 1u); }
    unsigned long long D;
    if (q <= 16) {
        D = (static_cast<unsigned long long>(k) * TD_5P16) << (16 - q);      // exact: k < 2^q
    } else {
        const unsigned long long N = static_cast<unsigned long long>(k) * TD_5P16;      // < 2^62
        const int sh = q - 16;
        if (sh > 62) {
            D = 0;                                                     // N < 2^62 <= half: rounds down
        } else {
            D = N >> sh;
            const unsigned long long rem = N & ((1ULL << sh) - 1ULL), half = 1ULL << (sh - 1);
            if (rem > half || (rem == half && (D & 1ULL))) ++D;
        }
    }
    if (D >= TD_10P16) { D -= TD_10P16; d.ipart += 1; }
    d.frac = D;
    return d;
}

__device__ __forceinline__ int td_int_digits(unsigned __int128 x) {
    if (x < 10) return 1;
    int n = 0;
    if ((x >> 64) == 0) {
        unsigned long long y = static_cast<unsigned long long>(x);
        while (y) { y /= 10; ++n; }
        return n;
    }
    while (x) { x /= 10; ++n; }
    return n;
}

// length of "%.16f" % v, Python semantics: nan -> "nan" (no sign), inf -> "inf" / "-inf"
__device__ __forceinline__ int td_token_len(float v) {
    const TdValue d = td_decompose(v);
    if (d.special == 1) return 3;
    if (d.special == 2) return 3 + (d.neg ? 1 : 0);
    return (d.neg ? 1 : 0) + td_int_digits(d.ipart) + 1 + 16;
}

__device__ __forceinline__ void td_token_write(float v, unsigned char* dst) {
    const TdValue d = td_decompose(v);
    if (d.special == 1) { dst[0] = 'n'; dst[1] = 'a'; dst[2] = 'n'; return; }
    if (d.neg) *dst++ = '-';
    if (d.special == 2) { dst[0] = 'i'; dst[1] = 'n'; dst[2] = 'f'; return; }
    const int nd = td_int_digits(d.ipart);
    if ((d.ipart >> 64) == 0) {
        unsigned long long y = static_cast<unsigned long long>(d.ipart);
        for (int i = nd - 1; i >= 0; --i) { dst[i] = static_cast<unsigned char>('0' + y % 10); y /= 10; }
    } else {
        unsigned __int128 y = d.ipart;
        for (int i = nd - 1; i >= 0; --i) { dst[i] = static_cast<unsigned char>('0' + static_cast<unsigned>(y % 10)); y /= 10; }
    }
    dst += nd;
    *dst++ = '.';
    unsigned long long f = d.frac;
#pragma unroll
    for (int i = 15; i >= 0; --i) { dst[i] = static_cast<unsigned char>('0' + f % 10); f /= 10; }
}

// pass 1: bytes of every block of 256 values (token + its separator)
__global__ void __launch_bounds__(TD_BLOCK) textdump_len_kernel(const float* __restrict__ x, long long n,
                                                                 unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned warp_sums[TD_BLOCK / 32];
    const long long i = static_cast<long long>(blockIdx.x) * TD_BLOCK + threadIdx.x;
    unsigned len = i < n ? static_cast<unsigned>(td_token_len(__ldg(x + i))) + 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = len;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
#pragma unroll
        for (int w = 0; w < TD_BLOCK / 32; ++w) s += warp_sums[w];
        block_sums[blockIdx.x] = s;
    }
}

// pass 2: exclusive scan of the block sums in place (one CTA; 1024 threads x contiguous chunks), total behind the last entry
__global__ void __launch_bounds__(1024) textdump_scan_kernel(unsigned long long* __restrict__ block_sums, long long n_blocks) {
    __shared__ unsigned long long part[1024];
    const long long per = (n_blocks + 1023) / 1024;
    const long long lo = min(n_blocks, static_cast<long long>(threadIdx.x) * per), hi = min(n_blocks, lo + per);
    unsigned long long s = 0;
    for (long long i = lo; i < hi; ++i) s += block_sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; ++t) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        block_sums[n_blocks] = run;                                     // total bytes
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (long long i = lo; i < hi; ++i) { const unsigned long long v = block_sums[i]; block_sums[i] = run; run += v; }
}

// pass 3: offsets inside the block, then every thread writes its token and the separator (' ' inside a row, '\n' at its end)
__global__ void __launch_bounds__(TD_BLOCK) textdump_write_kernel(const float* __restrict__ x, long long n, int cols,
                                                                   const unsigned long long* __restrict__ block_offsets,
                                                                   unsigned char* __restrict__ out) {
    __shared__ unsigned warp_sums[TD_BLOCK / 32];
    const long long i = static_cast<long long>(blockIdx.x) * TD_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float v = i < n ? __ldg(x + i) : 0.f;
    const unsigned len = i < n ? static_cast<unsigned>(td_token_len(v)) + 1u : 0u;
    unsigned incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < warp; ++w) base += warp_sums[w];
    if (i >= n) return;
    unsigned char* dst = out + block_offsets[blockIdx.x] + base + (incl - len);
    td_token_write(v, dst);
    dst[len - 1] = ((i + 1) % cols == 0) ? '\n' : ' ';
}

}  // namespace vqcpc

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" size_t vqcpc_textdump_workspace_bytes(int64_t rows, int32_t cols) {
    const long long n = static_cast<long long>(rows) * cols;
    return sizeof(unsigned long long) * static_cast<size_t>((n + vqcpc::TD_BLOCK - 1) / vqcpc::TD_BLOCK + 2);
}

// out_text == NULL: only the length is computed.  Returns VQCPC_ERR_ARG (with *out_len set) when out_capacity is too small.
extern "C" int vqcpc_textdump_f16(const float* x, int64_t rows, int32_t cols, unsigned char* out_text, size_t out_capacity,
                                  int64_t* out_len, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace vqcpc;
    VQ_ARG(out_len != nullptr, "textdump: out_len is null");
    *out_len = 0;
    VQ_ARG(rows >= 0 && cols >= 1, "textdump: bad shape %lld x %d", static_cast<long long>(rows), cols);
    const long long n = static_cast<long long>(rows) * cols;
    if (n == 0) return VQCPC_OK;
    VQ_ARG(x != nullptr && workspace != nullptr, "textdump: null pointer");
    VQ_ARG(workspace_bytes >= vqcpc_textdump_workspace_bytes(rows, cols), "textdump: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n_blocks = (n + TD_BLOCK - 1) / TD_BLOCK;
    VQ_ARG(n_blocks < (1LL << 31), "textdump: too many values for one call (%lld)", n);
    unsigned long long* sums = static_cast<unsigned long long*>(workspace);
    textdump_len_kernel<<<static_cast<unsigned>(n_blocks), TD_BLOCK, 0, st>>>(x, n, sums);
    VQ_CUDA(cudaGetLastError());
    textdump_scan_kernel<<<1, 1024, 0, st>>>(sums, n_blocks);
    VQ_CUDA(cudaGetLastError());
    count_launch(2);
    unsigned long long total = 0;
    VQ_CUDA(cudaMemcpyAsync(&total, sums + n_blocks, sizeof(total), cudaMemcpyDeviceToHost, st));
    VQ_CUDA(cudaStreamSynchronize(st));
    *out_len = static_cast<int64_t>(total);
    if (out_text == nullptr) return VQCPC_OK;
    VQ_ARG(out_capacity >= total, "textdump: output buffer too small (%zu < %llu)", out_capacity, total);
    textdump_write_kernel<<<static_cast<unsigned>(n_blocks), TD_BLOCK, 0, st>>>(x, n, cols, sums, out_text);
    VQ_CUDA(cudaGetLastError());
    count_launch(1);
    return VQCPC_OK;
}
