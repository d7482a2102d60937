// mxb_sort.cuh — key sorting kernels behind mxb_sort / mxb_unique (compiled into api.cu; plain tensors only, so nothing
// here goes through the expression code generator).
//
// Reference: sort_impl -> matxCubPlan_t::ExecSort / OptimizedExecSort (transforms/cub.h:428-560,2145-2190:
// cub::DeviceRadixSort::SortKeys for one row, DeviceSegmentedSort / per-row launches for batches) and the HostExecutor
// overload (std::sort per row, transforms/cub.h:2192-2240).  Keys only, ascending or descending, every row of the innermost
// dim sorted separately.
//
// Keys are mapped to unsigned integers whose order is the key order (floats: flip the sign bit of positives, all bits of
// negatives; signed integers: flip the sign bit; descending: complement), sorted, and mapped back on the last store.
//   rows of <= 4096 keys   one CTA per row: bitonic network in shared memory (one read, one write);
//   longer rows            least-significant-digit radix sort, 8 bits per pass.  Per pass: `count` (digit histogram of every
//                          chunk of a row), `scan` (one CTA per row: exclusive offsets of (digit, chunk) in digit-major
//                          order), `scatter` (a CTA walks its chunk in order, ranks 256 keys at a time — warp match + per-warp
//                          digit counters — and writes every key to its digit's running offset: stable).  2 reads + 1 write
//                          per pass and key; CUB's one-sweep does 1 + 1 — see DESIGN.md for the measured gap.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mxbsort {

typedef long long i64;
typedef unsigned int u32;
typedef unsigned long long u64;

// kind: 0 unsigned, 1 signed integer, 2 floating point
template <class K> __device__ __forceinline__ K to_key(K bits, int kind, bool desc) {
  constexpr K SIGN = (K)1 << (sizeof(K) * 8 - 1);
  K k = bits;
  if (kind == 1) k = bits ^ SIGN;
  else if (kind == 2) k = (bits & SIGN) ? (K)~bits : (K)(bits ^ SIGN);
  return desc ? (K)~k : k;
}
template <class K> __device__ __forceinline__ K from_key(K k, int kind, bool desc) {
  constexpr K SIGN = (K)1 << (sizeof(K) * 8 - 1);
  if (desc) k = (K)~k;
  if (kind == 1) return k ^ SIGN;
  if (kind == 2) return (k & SIGN) ? (K)(k ^ SIGN) : (K)~k;
  return k;
}

struct SortParams {
  const void *in;      // [B][L] keys (pass 0: raw element bits, later passes: mapped keys)
  void *out;
  i64 B, L;
  int kind, desc;
  int shift;           // bit offset of this pass's digit
  int first, last;     // first pass maps raw bits to keys, the last pass maps them back
  int cpr;             // chunks per row
  i64 chunk;           // keys per chunk (multiple of 256)
  u32 *counts;         // [B][cpr][256] keys per (chunk, digit)
  u32 *offsets;        // [B][cpr][256] where the chunk's keys of a digit start in the row
};

// ---- short rows: bitonic network in shared memory -------------------------------------------------------------------------
template <class K>
__global__ void __launch_bounds__(256) bitonic_rows_kernel(const SortParams p, int n_pow2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K *s = (K *)smem_raw;
  const bool desc = p.desc != 0;
  for (i64 b = blockIdx.x; b < p.B; b += gridDim.x) {
    const K *in = (const K *)p.in + b * p.L;
    K *out = (K *)p.out + b * p.L;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) s[i] = i < p.L ? to_key<K>(in[i], p.kind, desc) : (K)~(K)0;   // padding sorts last
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < n_pow2 / 2; i += blockDim.x) {
          const int lo = 2 * i - (i & (j - 1));      // index with bit j cleared
          const int hi = lo + j;
          const bool up = (lo & k) == 0;
          const K a = s[lo], c = s[hi];
          if ((a > c) == up) { s[lo] = c; s[hi] = a; }
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < p.L; i += blockDim.x) out[i] = from_key<K>(s[i], p.kind, desc);
    __syncthreads();
  }
}

// ---- long rows: one radix pass = count, scan, scatter ---------------------------------------------------------------------
template <class K> __device__ __forceinline__ K load_key(const SortParams &p, const K *row, i64 i) {
  const K v = row[i];
  return p.first ? to_key<K>(v, p.kind, p.desc != 0) : v;
}

template <class K>
__global__ void __launch_bounds__(256) radix_count_kernel(const SortParams p) {
  __shared__ u32 s_cnt[256];
  const i64 work = p.B * p.cpr;
  for (i64 w = blockIdx.x; w < work; w += gridDim.x) {
    const i64 b = w / p.cpr, c = w - b * p.cpr;
    const K *row = (const K *)p.in + b * p.L;
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const i64 i0 = c * p.chunk, i1 = (i0 + p.chunk < p.L) ? i0 + p.chunk : p.L;
    for (i64 i = i0 + threadIdx.x; i < i1; i += 256) {
      const u32 d = (u32)(load_key<K>(p, row, i) >> p.shift) & 255u;
      atomicAdd(&s_cnt[d], 1u);
    }
    __syncthreads();
    p.counts[(b * p.cpr + c) * 256 + threadIdx.x] = s_cnt[threadIdx.x];
    __syncthreads();
  }
}

// one CTA per row, thread d = digit d: offsets[c][d] = (keys of smaller digits in the row) + (keys of digit d in earlier chunks)
__global__ void __launch_bounds__(256) radix_scan_kernel(const SortParams p) {
  __shared__ u32 s_tot[256];
  const int d = threadIdx.x;
  for (i64 b = blockIdx.x; b < p.B; b += gridDim.x) {
    const u32 *__restrict__ cnt = p.counts + b * p.cpr * 256;
    u32 *__restrict__ off = p.offsets + b * p.cpr * 256;
    u32 tot = 0;
#pragma unroll 8
    for (int c = 0; c < p.cpr; ++c) tot += cnt[c * 256 + d];
    s_tot[d] = tot;
    __syncthreads();
    u32 run = 0;
    for (int k = 0; k < d; ++k) run += s_tot[k];
#pragma unroll 8
    for (int c = 0; c < p.cpr; ++c) { off[c * 256 + d] = run; run += cnt[c * 256 + d]; }
    __syncthreads();
  }
}

template <class K>
__global__ void __launch_bounds__(256) radix_scatter_kernel(const SortParams p) {
  __shared__ u32 s_off[256];        // running output offset of every digit for this chunk
  __shared__ u32 s_cnt[8][256];     // keys per (warp, digit) of the current round
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 work = p.B * p.cpr;
  for (i64 w = blockIdx.x; w < work; w += gridDim.x) {
    const i64 b = w / p.cpr, c = w - b * p.cpr;
    const K *row = (const K *)p.in + b * p.L;
    K *orow = (K *)p.out + b * p.L;
    s_off[tid] = p.offsets[(b * p.cpr + c) * 256 + tid];
    const i64 i0 = c * p.chunk, i1 = (i0 + p.chunk < p.L) ? i0 + p.chunk : p.L;
    for (i64 r0 = i0; r0 < i1; r0 += 256) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s_cnt[k][tid] = 0;
      __syncthreads();
      const i64 i = r0 + tid;
      const bool live = i < i1;
      K key = 0;
      u32 d = 256;   // dead lanes match nobody
      if (live) { key = load_key<K>(p, row, i); d = (u32)(key >> p.shift) & 255u; }
      // lanes of the warp holding the same digit: rank among them = lower lanes, the lowest of them records the count
      const u32 peers = __match_any_sync(0xffffffffu, d);
      const u32 rank = (u32)__popc(peers & ((1u << lane) - 1u));
      if (live && rank == 0) s_cnt[warp][d] = (u32)__popc(peers);
      __syncthreads();
      if (live) {
        u32 pos = s_off[d] + rank;
        for (int k = 0; k < warp; ++k) pos += s_cnt[k][d];
        orow[pos] = p.last ? from_key<K>(key, p.kind, p.desc != 0) : key;
      }
      __syncthreads();
      u32 add = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) add += s_cnt[k][tid];
      s_off[tid] += add;
    }
    __syncthreads();
  }
}

}  // namespace mxbsort
