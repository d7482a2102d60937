// stream_probe.cu — measurement scaffold, not product code: streaming-read (full-tensor sum) and read+write (vector add)
// kernel SHAPES timed side by side on one B200 with cub::DeviceReduce::Sum / a one-vector-per-thread add as the bar.
// Decides the launch shape of reduce_inner / ew in mxb_device.cuh (round 2: VERDICT "beat the reference where it is
// already at the roofline").  Build: tools/probe/build.sh  ->  tools/_bin/stream_probe ; run under gpurun.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef unsigned int u32;
typedef unsigned long long u64;

// ---- load / store flavours -------------------------------------------------------------------------
// F: 0 = nc + L1::no_allocate (ours), 1 = plain ld.global, 2 = nc, 3 = nc + no_allocate + L2::256B, 4 = L1::evict_first
template <int F> __device__ __forceinline__ float4 ld16(const float *p) {
  float4 r;
  if (F == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (F == 1) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (F == 2) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (F == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else asm volatile("ld.global.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
struct f8 { float v[8]; };
template <int F> __device__ __forceinline__ f8 ld32(const float *p) {
  f8 r;
  if (F == 0 || F == 3)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  else if (F == 1)
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  else if (F == 2)
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  else
    asm volatile("ld.global.L1::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
// S: 0 = st.global.L1::no_allocate (ours), 1 = plain, 2 = .cs (streaming), 3 = .wt
template <int S> __device__ __forceinline__ void st16(float *p, float4 r) {
  if (S == 0) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
  else if (S == 1) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
  else if (S == 2) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
  else asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
}
template <int S> __device__ __forceinline__ void st32(float *p, const f8 &r) {
  if (S == 0)
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
  else if (S == 1)
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
  else
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}

__device__ __forceinline__ float warp_sum(float a) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
  return a;
}
template <int BLOCK> __device__ __forceinline__ void cta_finish(float a, float *partial) {
  __shared__ float s[32];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < BLOCK / 32 ? s[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

// ---- read-only sum ---------------------------------------------------------------------------------
// MODE 0: tiles of BLOCK*U vectors dealt round-robin to the CTAs (what reduce_inner does today)
// MODE 1: every CTA owns one contiguous range of tiles (CUB's even-share)
// MODE 2: CTAs draw chunks of CH tiles from an atomic counter
template <int V, int U, int BLOCK, int MODE, int F, int CH>
__global__ void __launch_bounds__(BLOCK) k_sum(const float *__restrict__ x, size_t n, float *partial, u32 *counter) {
  const size_t tile = (size_t)BLOCK * U * V;
  const size_t ntiles = n / tile;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  auto do_tile = [&](size_t t) {
    const float *p = x + t * tile + (size_t)threadIdx.x * V;
    if (V == 4) {
      float4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = ld16<F>(p + (size_t)u * BLOCK * V);
#pragma unroll
      for (int u = 0; u < U; ++u) { acc[0] += r[u].x; acc[1] += r[u].y; acc[2] += r[u].z; acc[3] += r[u].w; }
    } else {
      f8 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = ld32<F>(p + (size_t)u * BLOCK * V);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += r[u].v[v];
    }
  };
  if (MODE == 0) {
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) do_tile(t);
  } else if (MODE == 1) {
    const size_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const size_t t0 = (size_t)blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
    for (size_t t = t0; t < t1; ++t) do_tile(t);
  } else {
    __shared__ u32 s_chunk;
    const size_t nchunks = (ntiles + CH - 1) / CH;
    while (true) {
      if (threadIdx.x == 0) s_chunk = atomicAdd(counter, 1u);
      __syncthreads();
      const size_t c = s_chunk;
      __syncthreads();
      if (c >= nchunks) break;
      const size_t t1 = (c + 1) * CH < ntiles ? (c + 1) * CH : ntiles;
      for (size_t t = c * CH; t < t1; ++t) do_tile(t);
    }
  }
  float a = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) a += acc[v];
  cta_finish<BLOCK>(a, partial);
}

// ---- read-only sum through a TMA bulk-copy ring ----------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}
// NC consumer threads + one producer warp; ring of STAGES buffers of SB bytes; chunks dealt round-robin to the CTAs
template <int NC, int STAGES, int SB>
__global__ void __launch_bounds__(NC + 32) k_sum_tma(const float *__restrict__ x, size_t n, float *partial) {
  extern __shared__ __align__(128) unsigned char smem[];
  u64 *full = (u64 *)smem, *empty = full + 8;
  unsigned char *ring = smem + 128;
  const int tid = threadIdx.x;
  const size_t nchunks = n * 4 / SB;
  const size_t mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= NC) {
    if (tid == NC) {
      int s = 0; u32 round = 0;
      for (size_t k = 0; k < mine; ++k) {
        if (round > 0) mbar_wait(&empty[s], (round - 1u) & 1u);
        const size_t c = blockIdx.x + k * gridDim.x;
        mbar_expect_tx(&full[s], SB);
        bulk_g2s(ring + (size_t)s * SB, (const char *)x + c * SB, SB, &full[s]);
        if (++s == STAGES) { s = 0; ++round; }
      }
    }
    return;
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int s = 0; u32 phase = 0;
  for (size_t k = 0; k < mine; ++k) {
    mbar_wait(&full[s], phase);
    const float4 *b = (const float4 *)(ring + (size_t)s * SB);
#pragma unroll 4
    for (int i = tid; i < SB / 16; i += NC) { const float4 v = b[i]; a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    if (++s == STAGES) { s = 0; phase ^= 1u; }
  }
  __shared__ float sp[32];
  float a = warp_sum((a0 + a1) + (a2 + a3));
  if ((tid & 31) == 0) sp[tid >> 5] = a;
  asm volatile("bar.sync 1, %0;" ::"r"(NC) : "memory");
  if (tid < 32) {
    float v = tid < NC / 32 ? sp[tid] : 0.f;
    v = warp_sum(v);
    if (tid == 0) partial[blockIdx.x] = v;
  }
}

// ---- vector add --------------------------------------------------------------------------------------
// PERSIST 0: one batch of U vectors per thread, grid covers N;  1: grid-stride
template <int V, int U, int BLOCK, int PERSIST, int F, int S>
__global__ void __launch_bounds__(BLOCK) k_add(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ o, size_t n) {
  const size_t nv = n / V;
  const size_t nthr = (size_t)gridDim.x * BLOCK;
  size_t q = (size_t)blockIdx.x * BLOCK * (PERSIST ? 1 : U) + threadIdx.x;
  const size_t ustride = PERSIST ? nthr : BLOCK;
  do {
    if (V == 4) {
      float4 x[U], y[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (q + u * ustride < nv) { x[u] = ld16<F>(a + (q + u * ustride) * V); y[u] = ld16<F>(b + (q + u * ustride) * V); }
#pragma unroll
      for (int u = 0; u < U; ++u) if (q + u * ustride < nv) {
        float4 r; r.x = x[u].x + y[u].x; r.y = x[u].y + y[u].y; r.z = x[u].z + y[u].z; r.w = x[u].w + y[u].w;
        st16<S>(o + (q + u * ustride) * V, r);
      }
    } else {
      f8 x[U], y[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (q + u * ustride < nv) { x[u] = ld32<F>(a + (q + u * ustride) * V); y[u] = ld32<F>(b + (q + u * ustride) * V); }
#pragma unroll
      for (int u = 0; u < U; ++u) if (q + u * ustride < nv) {
        f8 r;
#pragma unroll
        for (int v = 0; v < 8; ++v) r.v[v] = x[u].v[v] + y[u].v[v];
        st32<S>(o + (q + u * ustride) * V, r);
      }
    }
    q += (size_t)U * nthr;
  } while (PERSIST && q < nv);
}

// ---- harness -------------------------------------------------------------------------------------------
template <class L> float time_ms(L launch, int iters = 20, int reps = 3) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms / iters < best) best = ms / iters;
  }
  CK(cudaGetLastError());
  return best;
}
void report(const char *grp, const std::string &name, float ms, double bytes) {
  printf("{\"group\": \"%s\", \"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", grp, name.c_str(), ms, bytes / ms * 1e-6);
  fflush(stdout);
}

int main(int argc, char **argv) {
  const size_t n = (size_t)1 << 30;   // C2: 2^30 fp32
  const size_t na = (size_t)1 << 28;  // vector add: 2^28 fp32 per operand
  int sm = 148;
  CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
  float *x, *partial, *o; u32 *counter;
  CK(cudaMalloc(&x, n * 4));
  CK(cudaMalloc(&o, na * 4));
  CK(cudaMalloc(&partial, 1 << 22));
  CK(cudaMalloc(&counter, 4));
  CK(cudaMemset(x, 0, n * 4));
  const double rb = (double)n * 4;

  // --- bar: cub::DeviceReduce::Sum ---
  {
    void *tmp = nullptr; size_t tb = 0;
    cub::DeviceReduce::Sum(tmp, tb, x, partial, (long long)n);
    CK(cudaMalloc(&tmp, tb));
    report("sum", "cub::DeviceReduce::Sum", time_ms([&] { cub::DeviceReduce::Sum(tmp, tb, x, partial, (long long)n); }), rb);
    cudaFree(tmp);
  }
#define SUM(V, U, BLOCK, MODE, F, CH, CPS) \
  report("sum", "V" #V " U" #U " B" #BLOCK " mode" #MODE " F" #F " CH" #CH " cps" #CPS, \
         time_ms([&] { if (MODE == 2) cudaMemsetAsync(counter, 0, 4); k_sum<V, U, BLOCK, MODE, F, CH><<<sm * CPS, BLOCK>>>(x, n, partial, counter); }), rb)
  // today's shape and its neighbours
  SUM(4, 4, 256, 0, 0, 1, 8);
  SUM(4, 4, 256, 0, 0, 1, 4);
  SUM(4, 4, 256, 0, 0, 1, 6);
  SUM(4, 4, 256, 0, 0, 1, 16);
  SUM(4, 4, 512, 0, 0, 1, 4);
  SUM(4, 8, 256, 0, 0, 1, 4);
  SUM(4, 4, 256, 0, 1, 1, 8);
  SUM(4, 4, 256, 0, 2, 1, 8);
  SUM(4, 4, 256, 0, 3, 1, 8);
  SUM(4, 4, 256, 0, 4, 1, 8);
  // contiguous ranges (even-share)
  SUM(4, 4, 256, 1, 0, 1, 8);
  SUM(4, 4, 256, 1, 0, 1, 40);
  SUM(4, 4, 512, 1, 1, 1, 20);
  SUM(4, 4, 256, 1, 1, 1, 40);
  // dynamic chunks
  SUM(4, 4, 256, 2, 0, 4, 8);
  SUM(4, 4, 256, 2, 0, 16, 8);
  SUM(4, 4, 256, 2, 0, 64, 8);
  SUM(4, 4, 256, 2, 0, 16, 4);
  SUM(4, 4, 512, 2, 0, 8, 4);
  // 32-byte loads
  SUM(8, 2, 256, 0, 0, 1, 8);
  SUM(8, 4, 256, 0, 0, 1, 4);
  SUM(8, 2, 256, 0, 1, 1, 8);
  SUM(8, 4, 256, 0, 1, 1, 4);
  SUM(8, 2, 256, 2, 0, 16, 8);
  SUM(8, 4, 256, 2, 0, 8, 4);
  SUM(8, 2, 512, 0, 0, 1, 4);
  SUM(8, 2, 128, 0, 0, 1, 16);
#define TMA(NC, ST, SB, CPS) do { \
    const int smem = 128 + ST * SB; \
    CK(cudaFuncSetAttribute(k_sum_tma<NC, ST, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    report("sum", "tma NC" #NC " ST" #ST " SB" #SB " cps" #CPS, time_ms([&] { k_sum_tma<NC, ST, SB><<<sm * CPS, NC + 32, smem>>>(x, n, partial); }), rb); } while (0)
  TMA(128, 4, 16384, 2);
  TMA(128, 6, 16384, 2);
  TMA(256, 4, 16384, 2);
  TMA(256, 3, 32768, 2);
  TMA(256, 6, 32768, 1);
  TMA(128, 6, 8192, 4);
  TMA(128, 3, 16384, 4);
  TMA(256, 8, 8192, 2);
  TMA(512, 6, 32768, 1);

  // --- vector add ---
  float *a = x, *b = x + na;
  const double ab = (double)na * 12;
#define ADD(V, U, BLOCK, PERSIST, F, S, CPS) \
  report("add", "V" #V " U" #U " B" #BLOCK " persist" #PERSIST " F" #F " S" #S " cps" #CPS, \
         time_ms([&] { const size_t nv = na / V; \
                       const unsigned g = PERSIST ? (unsigned)(sm * CPS) : (unsigned)((nv + (size_t)BLOCK * U - 1) / ((size_t)BLOCK * U)); \
                       k_add<V, U, BLOCK, PERSIST, F, S><<<g, BLOCK>>>(a, b, o, na); }), ab)
  ADD(8, 1, 256, 0, 1, 1, 0);   // the reference's shape: EPT 8, one vector per thread, default cache ops
  ADD(4, 4, 256, 0, 0, 0, 0);   // ours today
  ADD(8, 1, 256, 0, 0, 0, 0);
  ADD(8, 2, 256, 0, 0, 0, 0);
  ADD(8, 1, 256, 0, 0, 1, 0);
  ADD(8, 1, 256, 0, 1, 0, 0);
  ADD(8, 1, 256, 0, 2, 1, 0);
  ADD(8, 1, 256, 0, 4, 2, 0);
  ADD(8, 1, 256, 0, 1, 2, 0);
  ADD(8, 1, 128, 0, 1, 1, 0);
  ADD(8, 1, 512, 0, 1, 1, 0);
  ADD(8, 2, 256, 0, 1, 1, 0);
  ADD(4, 1, 256, 0, 1, 1, 0);
  ADD(4, 2, 256, 0, 1, 1, 0);
  ADD(4, 4, 256, 0, 1, 1, 0);
  ADD(8, 1, 256, 1, 1, 1, 8);
  ADD(8, 2, 256, 1, 1, 1, 8);
  ADD(8, 1, 256, 1, 1, 1, 16);
  ADD(4, 4, 256, 1, 0, 0, 8);
  ADD(8, 2, 256, 1, 0, 0, 8);
  return 0;
}
