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
//                          chunk of a row), `scan` (a warp per (row, digit): exclusive offsets over the chunks), `scatter`
//                          (a CTA walks its chunk in rounds of 2048 keys, ranks them with warp matches + per-warp digit
//                          counters and writes every key to its digit's running offset: stable).  2 reads + 1 write per
//                          pass and key; CUB's one-sweep does 1 + 1 — see DESIGN.md for the measured gap.
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
  u32 *counts;         // [B][256][cpr] keys per (digit, chunk)
  u32 *offsets;        // [B][256][cpr] keys of the digit in earlier chunks of the row
  u32 *totals;         // [B][256] keys of the digit in the row
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
// Counters are digit-major: counts[b][d][c] = keys of digit d in chunk c of row b, so the scan of one digit over the chunks
// of a row is a walk over contiguous words.
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
    // four loads in flight per thread; plain shared-memory atomics (a warp match costs one step per DISTINCT digit in the
    // warp — 140 us instead of 48 per pass over 2^24 uniformly random keys)
    for (i64 ii = i0 + threadIdx.x; ii < i1; ii += 1024) {
      K k4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) if (ii + j * 256 < i1) k4[j] = load_key<K>(p, row, ii + j * 256);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (ii + j * 256 < i1) atomicAdd(&s_cnt[(u32)(k4[j] >> p.shift) & 255u], 1u);
    }
    __syncthreads();
    p.counts[(b * 256 + threadIdx.x) * p.cpr + c] = s_cnt[threadIdx.x];
    __syncthreads();
  }
}

// one WARP per (row, digit): exclusive scan of the digit's counts over the row's chunks (contiguous words), in place into
// `offsets`, and the digit's total into totals[b][d] (the scatter kernel turns the totals into the digits' bases)
__global__ void __launch_bounds__(256) radix_scan_kernel(const SortParams p) {
  const int lane = threadIdx.x & 31;
  const i64 nw = (i64)gridDim.x * 8, w0 = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (i64 w = w0; w < p.B * 256; w += nw) {
    const u32 *__restrict__ cnt = p.counts + w * p.cpr;
    u32 *__restrict__ off = p.offsets + w * p.cpr;
    u32 run = 0;
    for (int c0 = 0; c0 < p.cpr; c0 += 128) {   // four loads in flight: the walk is one warp against L2 latency
      u32 x[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int c = c0 + k * 32 + lane; x[k] = c < p.cpr ? cnt[c] : 0u; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + k * 32 + lane;
        u32 incl = x[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        if (c < p.cpr) off[c] = run + incl - x[k];
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if (lane == 0) p.totals[w] = run;
  }
}

// A CTA walks its chunk in rounds of 2048 keys: warp w owns the round's keys [w * 256, w * 256 + 256), lane l its keys
// j * 32 + l (coalesced), so the order inside the round is (warp, j, lane) = the index order.  Ranks: per j the lanes of a
// warp with the same digit are ranked by votes and added to the warp's running counter of that digit (no CTA barrier
// inside the round).  The round is then sorted by digit IN SHARED MEMORY (position = digits before + warps before + rank)
// and written out from there, consecutive threads to consecutive addresses of a digit's run: scattering the keys straight
// from the registers cost one 32-byte sector transaction per 4-byte key (146 us of a 175 us pass on random keys).  Stable.
template <class K>
__global__ void __launch_bounds__(256) radix_scatter_kernel(const SortParams p) {
  constexpr int KPT = 8, ROUND = 256 * KPT;
  __shared__ u32 s_off[256];        // where the next key of every digit goes (row-relative), for this chunk
  __shared__ u32 s_cnt[8][256];     // keys per (warp, digit) of the current round
  __shared__ u32 s_base[8][256];    // position inside the round's sorted order of the first key of (warp, digit)
  __shared__ u32 s_delta[256];      // output position of a key of digit d = s_delta[d] + its position inside the round
  __shared__ u32 s_scan[2][8];
  __shared__ K s_keys[ROUND];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = (1u << lane) - 1u;
  const i64 work = p.B * p.cpr;
  for (i64 w = blockIdx.x; w < work; w += gridDim.x) {
    const i64 b = w / p.cpr, c = w - b * p.cpr;
    const K *row = (const K *)p.in + b * p.L;
    K *orow = (K *)p.out + b * p.L;
    // base of digit `tid` in the row = total of the smaller digits (exclusive scan of the 256 totals over the CTA)
    {
      const u32 t = p.totals[b * 256 + tid];
      u32 incl = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) s_scan[0][warp] = incl;
      __syncthreads();
      u32 before = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < warp) before += s_scan[0][k];
      s_off[tid] = before + incl - t + p.offsets[(b * 256 + tid) * p.cpr + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) s_cnt[k][tid] = 0;
    }
    __syncthreads();
    const i64 i0 = c * p.chunk, i1 = (i0 + p.chunk < p.L) ? i0 + p.chunk : p.L;
    int par = 1;
    for (i64 r0 = i0; r0 < i1; r0 += ROUND, par ^= 1) {
      const int nround = (int)((i1 - r0 < ROUND) ? (i1 - r0) : ROUND);
      K key[KPT];
      u32 dig[KPT], rank[KPT];
      const i64 wbase = r0 + (i64)warp * 32 * KPT + lane;
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const i64 i = wbase + (i64)j * 32;
        dig[j] = 256u;   // dead lanes: excluded from every vote below
        if (i < i1) { key[j] = load_key<K>(p, row, i); dig[j] = (u32)(key[j] >> p.shift) & 255u; }
      }
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const bool live = dig[j] < 256u;
        // lanes holding the same digit, from one vote per digit bit (a fixed nine instructions; match.any takes one step
        // per distinct value in the warp, 32 of them on random keys)
        u32 peers = __ballot_sync(0xffffffffu, live);
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
          const u32 vote = __ballot_sync(0xffffffffu, (dig[j] >> bit) & 1u);
          peers &= ((dig[j] >> bit) & 1u) ? vote : ~vote;
        }
        u32 before = 0;
        if (live) before = s_cnt[warp][dig[j]];
        __syncwarp();
        rank[j] = before + (u32)__popc(peers & lt);
        if (live && (peers & lt) == 0) s_cnt[warp][dig[j]] = before + (u32)__popc(peers);
        __syncwarp();
      }
      __syncthreads();
      {   // thread d: the round's keys of digit d per warp -> where (warp, d) starts inside the round's sorted order
        u32 n[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { n[k] = s_cnt[k][tid]; tot += n[k]; s_cnt[k][tid] = 0; }
        u32 incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        if (lane == 31) s_scan[par][warp] = incl;
        __syncthreads();
        u32 lstart = incl - tot;
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < warp) lstart += s_scan[par][k];
        u32 run = lstart;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_base[k][tid] = run; run += n[k]; }
        s_delta[tid] = s_off[tid] - lstart;
        s_off[tid] += tot;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < KPT; ++j)
        if (dig[j] < 256u) s_keys[s_base[warp][dig[j]] + rank[j]] = key[j];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KPT; ++k) {
        const int i = k * 256 + tid;
        if (i < nround) {
          const K kk = s_keys[i];
          const u32 pos = s_delta[(u32)(kk >> p.shift) & 255u] + (u32)i;
          orow[pos] = p.last ? from_key<K>(kk, p.kind, p.desc != 0) : kk;
        }
      }
      // the next round touches s_keys / s_base / s_delta only behind its own barriers; s_scan alternates by round parity
    }
    __syncthreads();
  }
}

}  // namespace mxbsort
