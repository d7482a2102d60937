// mxb_device.cuh — device-side skeletons of the B200 reduce / fused-elementwise engine.
//
// This header is compiled two ways from the same text:
//   * ahead of time by nvcc (-gencode arch=compute_100a,code=sm_100a) for the named expressions, and
//   * at run time by NVRTC for any other expression program (see jit.cpp),
// so it must stay free of host headers.  A generated `struct Expr` (codegen.cpp) supplies the leaf
// loads and the per-element arithmetic; the kernels below supply everything the reference gets from
// the generic executor kernels (executors/kernel.h:41-223) and from CUB's device / segmented reduce
// (transforms/cub.h:647-894,1281-1328): index decomposition, vectorised coalesced loads, the
// per-thread / warp / CTA / grid reduction stages and the output conversion.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mxb {

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int KMAXD = 4;      // collapsed dims per group handled on the device
constexpr int KMAXLEAF = 12;  // == MXB_MAX_LEAVES
constexpr int KMAXCONST = 24; // == MXB_MAX_CONSTS

// ------------------------------------------------------------------------------------------------
// value types
// ------------------------------------------------------------------------------------------------
struct __align__(8) cfloat {
  float re, im;
  cfloat() = default;
  __device__ __forceinline__ cfloat(float r, float i = 0.f) : re(r), im(i) {}
};
__device__ __forceinline__ cfloat operator+(cfloat a, cfloat b) { return cfloat(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cfloat operator-(cfloat a, cfloat b) { return cfloat(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cfloat operator-(cfloat a) { return cfloat(-a.re, -a.im); }
__device__ __forceinline__ cfloat operator*(cfloat a, cfloat b) {
  return cfloat(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
__device__ __forceinline__ cfloat operator*(cfloat a, float b) { return cfloat(a.re * b, a.im * b); }
__device__ __forceinline__ cfloat operator*(float a, cfloat b) { return cfloat(a * b.re, a * b.im); }
__device__ __forceinline__ cfloat operator/(cfloat a, float b) { return cfloat(a.re / b, a.im / b); }
__device__ __forceinline__ cfloat operator+(cfloat a, float b) { return cfloat(a.re + b, a.im); }
__device__ __forceinline__ cfloat operator+(float a, cfloat b) { return cfloat(a + b.re, b.im); }
__device__ __forceinline__ cfloat operator-(cfloat a, float b) { return cfloat(a.re - b, a.im); }
__device__ __forceinline__ cfloat operator-(float a, cfloat b) { return cfloat(a - b.re, -b.im); }
__device__ __forceinline__ cfloat operator/(cfloat a, cfloat b) {
  // Smith's algorithm, as libcu++'s complex division does for finite operands
  if (fabsf(b.re) >= fabsf(b.im)) {
    float r = b.im / b.re, d = b.re + b.im * r;
    return cfloat((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  } else {
    float r = b.re / b.im, d = b.re * r + b.im;
    return cfloat((a.re * r + a.im) / d, (a.im * r - a.re) / d);
  }
}
__device__ __forceinline__ cfloat operator/(float a, cfloat b) { return cfloat(a, 0.f) / b; }
__device__ __forceinline__ bool operator==(cfloat a, cfloat b) { return a.re == b.re && a.im == b.im; }
__device__ __forceinline__ bool operator!=(cfloat a, cfloat b) { return !(a == b); }

template <class T> struct is_complex { enum { value = 0 }; };
template <> struct is_complex<cfloat> { enum { value = 1 }; };
template <class A, class B> struct same_t { enum { value = 0 }; };
template <class A> struct same_t<A, A> { enum { value = 1 }; };

// ------------------------------------------------------------------------------------------------
// kernel parameter blocks (plain data, passed by value as __grid_constant__)
// ------------------------------------------------------------------------------------------------
struct LeafDev {
  const void *ptr;
  i64 bs[KMAXD];  // strides over the (collapsed) batch / outer dims, elements
  i64 rs[KMAXD];  // strides over the (collapsed) reduce dims, elements
};
struct OutDev {
  void *ptr;
  i64 bs[KMAXD];
};
struct ConstDev {
  double dre[KMAXCONST], dim[KMAXCONST];
  float fre[KMAXCONST], fim[KMAXCONST];
  i64 ire[KMAXCONST];
};

constexpr int KMAXITEMS = 8;  // statements per exchange
// where a finished 32-byte record is pushed in the fused multi-GPU exchange
struct PeerPush {
  void *rec[8];        // rank r's PartialRec[2 slots][world][KMAXITEMS], as mapped into this device
  u32 *flag[8];        // rank r's arrival counters u32[world]
  const u32 *epoch;    // this rank's count of completed exchanges
  int world, rank, item, pad_;
};

// reduction: batch dims (nb of them, row-major) x reduce dims (nr of them, row-major, innermost last)
struct RedParams {
  int nb, nr;
  i64 bsz[KMAXD], rsz[KMAXD];
  i64 B, R;        // products of bsz / rsz
  i64 bflat[KMAXD]; // weight of each batch dim in the ORIGINAL row-major batch index (dims may be rotated)
  int nleaf;
  int splits;      // CTAs cooperating on one row (inner_cta) or on one output tile (outer)
  LeafDev leaf[KMAXLEAF];
  OutDev out, idx;
  void *out2, *idx2;  // argminmax: the max value / max index outputs (same batch strides as out / idx)
  void *ws;        // splits > 1: B*splits partial records
  u32 *tickets;    // splits > 1: one self-resetting counter per row / tile
  i64 idx_base;    // added to every reported flat index (slab offset in multi-GPU partials)
  float post_scale_f; double post_scale_d; // MEAN: divide by this; VAR: N - ddof
  int post_div;    // 1 = divide the sum by post_scale
  int post_sqrt;   // STDD
  int raw_partial; // 1 = write the 32-byte partial record instead of the finalised value (multi-GPU)
  int all_unit;    // 1 = every leaf is unit-stride along the vector dim
  int tx;          // outer family: threads along the vector (column) dim; blockDim.x / tx reduce lanes
  ConstDev c;
  // raw_partial == 2: the finished 32-byte record is pushed straight into every rank's exchange buffer over
  // NVLink peer mappings (no collective call): rec[r] = rank r's PartialRec[2 slots][world][KMAXITEMS],
  // flag[r] = rank r's arrival counters u32[world], epoch = this rank's count of completed exchanges
  PeerPush peer;
  i64 out_rs[KMAXD];  // softmax / scan: strides of the output over the reduce dims
  // scan family only: the handle's scan workspace (tile / group totals, epoch-coded flags)
  // TILES mode: published totals {tag : value} of tiles / groups of 32 tiles / supergroup starts (8-byte slots for 4-byte
  // values, 16-byte slots for 8-byte values), the device-resident launch epoch and the exit ticket
  void *scan_agg, *scan_gagg, *scan_sagg, *scan_own;
  int scan_plain;  // the operand is ONE contiguous tensor of the value type: the warp-tile scan may read it with L2 policies
  u32 *scan_ctl;
  // hist family: HistogramEven bounds and bin count (reference: cub::DeviceHistogram::HistogramEven behind hist_impl)
  double hist_lo_d, hist_hi_d;
  i64 hist_lo_i, hist_hi_i;
  int hist_bins;
  int hist_smem;    // 1 = privatised bins in shared memory, 0 = straight to global memory (more bins than fit)
  int scan_depth;   // tiles of one CTA between their phase 1 (local scan, parked in shared memory) and phase 2 (carry + store)
  int scan_group;   // warp team, rows of <= 32 vectors: lanes per row (power of two), 0 = a whole warp per row
  u32 scan_flags;   // TILES mode: bit 0 issues the next tile's loads before the publish instead of after (sweep knob)
  int tma_rt;       // reduce_outer_tma: reduce rows per ring stage (splits = ring depth, tx = 16-byte chunks per strip)
  int tma_mode;     // reduce_outer_tma: 1 = tile copies through `tmap`, 0 = cp.async.bulk copies (whole stage or per row)
  // CUtensorMap over the leaf as {vector dim in 8-byte elements, reduce dim, outer batch dim} (opaque 128 bytes, must
  // stay in the kernel's parameter space: the copy instruction takes its address)
  __align__(64) unsigned long long tmap[16];
  // reduce_inner, CTA-per-item flavour: dynamic work distribution.  work_ctr != nullptr: a CTA takes item blockIdx.x first
  // and then draws gridDim.x + atomicAdd(work_ctr, 1) until the items run out (the SMs never run at the same speed — HBM
  // channel conflicts, the other die — and a static deal ends at the speed of the slowest); work_ctr[1] is the
  // self-resetting exit ticket whose last holder zeroes work_ctr[0] for the next launch.  Which CTA runs an item does
  // not change what the item computes, so results stay run-to-run deterministic.
  u32 *work_ctr;
  int chunk_tiles;  // > 0 (one contiguous reduce run): split s of a row owns tiles [s * chunk_tiles, (s + 1) * chunk_tiles)
  // 1 (dynamic deal, one row, an op whose combine is exact in any order — max / min / arg / any / all, integer sums): a
  // CTA keeps ONE accumulator across all the items it draws and writes one partial at exit (no per-item CTA stage)
  int carry_items;
};

// elementwise: up to KMAXD collapsed dims, innermost last
struct EwParams {
  int nd;
  i64 sz[KMAXD];
  i64 N;           // product
  int nleaf;
  int all_unit;    // 1 = every leaf is unit-stride along the vector dim
  LeafDev leaf[KMAXLEAF]; // bs[] used
  OutDev out;
  ConstDev c;
  // transposing family (ew_tr) only
  int tr_ydim;        // the dim the staged leaves are unit-stride along
  unsigned tr_ymask;  // bit k: leaf k is staged through shared memory
  int tr_yvec;        // staged leaves may be read in aligned 16-byte chunks
  int tr_xvec;        // X-walking leaves may be read in aligned V-element vectors
  int tr_ovec;        // the output may be written in aligned V-element vectors
  // select family (find / find_idx) only
  int sel_op;                      // mxb_select_op_t: x < c, x > c, x == c, x != c, x <= c, x >= c
  double sel_thr_d;                // the threshold c for floating value types ...
  i64 sel_thr_i;                   // ... and for integer ones
  unsigned long long *sel_counts;  // selected elements per CTA (count pass; both passes give a CTA the same run of tiles)
  unsigned long long *sel_offsets; // exclusive prefix of sel_counts (written by the count pass's last CTA)
  u32 *sel_ticket;                 // self-resetting arrival counter of the count pass
  int *sel_total;                  // num_found (clamped to INT_MAX like the reference's int count)
  i64 sel_cap;                     // capacity of the output: elements beyond it are counted but not written
  // single-pass select (select1p): one 64-bit status word per tile {(epoch << 2 | state) : running count}, the launch
  // epoch (device-resident: bumped by the last CTA out, so a CUDA-graph replay sees a fresh epoch too) and the exit ticket
  unsigned long long *sel_status;
  u32 *sel_epoch;
  int sel_depth;                   // tiles of one CTA between their phase 1 (rank + stage) and phase 2 (offset + copy-out): 4..8
};

// ------------------------------------------------------------------------------------------------
// vector register bundles and global loads
// ------------------------------------------------------------------------------------------------
template <class T, int V> struct Vec { T v[V]; };

// streaming read-only loads: data is touched once, keep it out of L1

template <int BYTES> struct LdBytes;
template <> struct LdBytes<1> {
  static __device__ __forceinline__ void ld(void *d, const void *s) { *(unsigned char *)d = __ldg((const unsigned char *)s); }
};
template <> struct LdBytes<2> {
  static __device__ __forceinline__ void ld(void *d, const void *s) { *(unsigned short *)d = __ldg((const unsigned short *)s); }
};
template <> struct LdBytes<4> {
  static __device__ __forceinline__ void ld(void *d, const void *s) {
    u32 r;
    asm("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(s));
    *(u32 *)d = r;
  }
};
template <> struct LdBytes<8> {
  static __device__ __forceinline__ void ld(void *d, const void *s) {
    u32 a, b;
    asm("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(s));
    ((u32 *)d)[0] = a; ((u32 *)d)[1] = b;
  }
};
// MXB_LD_FLAVOR (development knob, JIT only): cache policy of the 128-bit streaming load
#ifndef MXB_LD_FLAVOR
#define MXB_LD_FLAVOR 0
#endif
#if MXB_LD_FLAVOR == 1
#define MXB_LD16 "ld.global.nc.v4.b32"
#elif MXB_LD_FLAVOR == 2
#define MXB_LD16 "ld.global.L1::no_allocate.v4.b32"
#elif MXB_LD_FLAVOR == 3
#define MXB_LD16 "ld.global.nc.L1::no_allocate.L2::256B.v4.b32"
#elif MXB_LD_FLAVOR == 4
#define MXB_LD16 "ld.global.nc.L1::evict_first.v4.b32"
#elif MXB_LD_FLAVOR == 5
#define MXB_LD16 "ld.global.nc.L1::no_allocate.L2::128B.v4.b32"
#else
#define MXB_LD16 "ld.global.nc.L1::no_allocate.v4.b32"
#endif
template <> struct LdBytes<16> {
  static __device__ __forceinline__ void ld(void *d, const void *s) {
    u32 a, b, c, e;
    asm(MXB_LD16 " {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(e) : "l"(s));
    ((u32 *)d)[0] = a; ((u32 *)d)[1] = b; ((u32 *)d)[2] = c; ((u32 *)d)[3] = e;
  }
};
#if defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ < 1209)
// 256-bit global accesses need PTX ISA 8.8 (CUDA 12.9); an older NVRTC gets two 128-bit accesses instead
#define MXB_NO_256BIT 1
#endif
template <> struct LdBytes<32> {  // LDG.E.256 on sm_100
  static __device__ __forceinline__ void ld(void *d, const void *s) {
#ifdef MXB_NO_256BIT
    LdBytes<16>::ld(d, s);
    LdBytes<16>::ld((char *)d + 16, (const char *)s + 16);
    return;
#else
    u32 r0, r1, r2, r3, r4, r5, r6, r7;
    asm("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(s));
    u32 *o = (u32 *)d;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = r5; o[6] = r6; o[7] = r7;
#endif
  }
};
template <> struct LdBytes<64> {
  static __device__ __forceinline__ void ld(void *d, const void *s) {
    LdBytes<32>::ld(d, s);
    LdBytes<32>::ld((char *)d + 32, (const char *)s + 32);
  }
};

// load V consecutive elements (address must be aligned to V*sizeof(T), or V == 1)
template <class T, int V> __device__ __forceinline__ void ldv(Vec<T, V> &r, const T *p) {
  LdBytes<(int)sizeof(T) * V>::ld(&r, p);
}
// leaf whose innermost stride is 0: one scalar load, replicated
template <class T, int V> __device__ __forceinline__ void ldsplat(Vec<T, V> &r, const T *p) {
  Vec<T, 1> s;
  LdBytes<(int)sizeof(T)>::ld(&s, p);
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = s.v[0];
}
// UNIT: every leaf of the expression is unit-stride along the vector dim (decided once per launch on the
// host), so the hot loops carry no per-leaf stride test and the loads of an unrolled batch can all be in flight.
template <class T, int V, bool UNIT> __device__ __forceinline__ void ldleaf(Vec<T, V> &r, const void *base, i64 j, i64 inner) {
  const T *p = (const T *)base;
  if (UNIT) {
    ldv<T, V>(r, p + j);
  } else if (V == 1) {
    LdBytes<(int)sizeof(T)>::ld(&r, p + j * inner);
  } else if (inner != 0) {
    ldv<T, V>(r, p + j);
  } else {
    ldsplat<T, V>(r, p);
  }
}

// streaming stores
template <int BYTES> struct StBytes;
template <> struct StBytes<1> { static __device__ __forceinline__ void st(void *d, const void *s) { *(unsigned char *)d = *(const unsigned char *)s; } };
template <> struct StBytes<2> { static __device__ __forceinline__ void st(void *d, const void *s) { *(unsigned short *)d = *(const unsigned short *)s; } };
template <> struct StBytes<4> {
  static __device__ __forceinline__ void st(void *d, const void *s) {
    asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(d), "r"(*(const u32 *)s) : "memory");
  }
};
template <> struct StBytes<8> {
  static __device__ __forceinline__ void st(void *d, const void *s) {
    const u32 *x = (const u32 *)s;
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(d), "r"(x[0]), "r"(x[1]) : "memory");
  }
};
template <> struct StBytes<16> {
  static __device__ __forceinline__ void st(void *d, const void *s) {
    const u32 *x = (const u32 *)s;
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(d), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
  }
};
template <> struct StBytes<32> {  // STG.E.256 on sm_100
  static __device__ __forceinline__ void st(void *d, const void *s) {
#ifdef MXB_NO_256BIT
    StBytes<16>::st(d, s);
    StBytes<16>::st((char *)d + 16, (const char *)s + 16);
    return;
#endif
    const u32 *x = (const u32 *)s;
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "r"(x[0]), "r"(x[1]),
                 "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]) : "memory");
  }
};
template <> struct StBytes<64> {
  static __device__ __forceinline__ void st(void *d, const void *s) {
    StBytes<32>::st(d, s);
    StBytes<32>::st((char *)d + 32, (const char *)s + 32);
  }
};

// ------------------------------------------------------------------------------------------------
// conversions between the arithmetic type of an expression and a stored element type
// ------------------------------------------------------------------------------------------------
template <class To, class From> struct Cvt {
  static __device__ __forceinline__ To go(From x) { return (To)x; }
};
template <class From> struct Cvt<__nv_bfloat16, From> {
  static __device__ __forceinline__ __nv_bfloat16 go(From x) { return __float2bfloat16_rn((float)x); }
};
template <class From> struct Cvt<__half, From> {
  static __device__ __forceinline__ __half go(From x) { return __float2half_rn((float)x); }
};
template <class To> struct Cvt<To, __nv_bfloat16> {
  static __device__ __forceinline__ To go(__nv_bfloat16 x) { return (To)__bfloat162float(x); }
};
template <class To> struct Cvt<To, __half> {
  static __device__ __forceinline__ To go(__half x) { return (To)__half2float(x); }
};
template <> struct Cvt<__nv_bfloat16, __nv_bfloat16> { static __device__ __forceinline__ __nv_bfloat16 go(__nv_bfloat16 x) { return x; } };
template <> struct Cvt<__half, __half> { static __device__ __forceinline__ __half go(__half x) { return x; } };
template <> struct Cvt<__half, __nv_bfloat16> { static __device__ __forceinline__ __half go(__nv_bfloat16 x) { return __float2half_rn(__bfloat162float(x)); } };
template <> struct Cvt<__nv_bfloat16, __half> { static __device__ __forceinline__ __nv_bfloat16 go(__half x) { return __float2bfloat16_rn(__half2float(x)); } };
template <> struct Cvt<cfloat, cfloat> { static __device__ __forceinline__ cfloat go(cfloat x) { return x; } };
template <class From> struct Cvt<cfloat, From> {
  static __device__ __forceinline__ cfloat go(From x) { return cfloat(Cvt<float, From>::go(x), 0.f); }
};
template <class To> struct Cvt<To, cfloat> {  // complex -> real keeps the real part (only reached for any/all style 0/1 outputs)
  static __device__ __forceinline__ To go(cfloat x) { return Cvt<To, float>::go(x.re); }
};
template <class To, class From> __device__ __forceinline__ To cvt(From x) { return Cvt<To, From>::go(x); }

template <class T> __device__ __forceinline__ bool nonzero(T x) { return x != (T)0; }
template <> __device__ __forceinline__ bool nonzero<cfloat>(cfloat x) { return x.re != 0.f || x.im != 0.f; }

// ------------------------------------------------------------------------------------------------
// scalar math used by generated expressions (reference functors: operators/scalar_ops.h:434-503,
// scalar_internal.h:44-297 which forward to cuda::std / CUDA math)
// ------------------------------------------------------------------------------------------------
// normcdf(x) for fp32, hand-written: the CUDA library's normcdff costs ~63 SASS instructions per call and the fused
// Black-Scholes chain (examples/black_scholes.cu:122-138, two calls per option) is instruction-issue bound on B200, not
// HBM bound, with it.  This one is ~29: with z = min(|x|, 14.5),
//   Phi(-z) = exp(-z^2/2) * Q(z),   Q(z) = erfcx(z/sqrt2)/2 = u * P(t),  u = 1/(z+2),  t = (z-2)/(z+2) = 1 - 4u
// P = degree-10 minimax fit (relative error 2.4e-8 on [0, 14.6], fitted here from scipy's erfcx, not copied from
// anywhere).  The pole offset 2 (not a larger one that would need fewer terms) keeps the map z -> u well conditioned
// near z = 0, where half an ulp of u is worth (z+2) * 2^-25 in z.  exp(-z^2/2) = ex2(a_hi) * (1 + ln2 * a_lo) with
// the rounding residue of z*z and of the product with -0.5*log2(e) carried in a_lo, so the tail keeps its RELATIVE
// accuracy (|x| = 10: a = -72, an uncompensated product would be off by 3e-6).  Measured on B200 against fp64 truth
// (tools/math_accuracy.py, profiles/): see DESIGN.md section 4; the library function documents 5 ulp.  NaN
// propagates (min.NaN), +-inf give 1 / 0, results below 2^-126 flush to zero (|x| > 13.2, where the library returns
// denormals).
__device__ __forceinline__ float f_normcdf(float x) {
  float z;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(z) : "f"(fabsf(x)), "f"(14.5f));
  const float s = z * z, sl = fmaf(z, z, -s);
  const float d = z + 2.0f;
  float u;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(d));
  u = fmaf(u, fmaf(-d, u, 1.0f), u);  // one Newton step: u = 1/d to half an ulp
  const float t = fmaf(-4.0f, u, 1.0f);
  float p = -7.689554332e-05f;
  p = fmaf(p, t, 2.991535075e-05f);
  p = fmaf(p, t, 7.246770547e-04f);
  p = fmaf(p, t, 8.852431783e-04f);
  p = fmaf(p, t, -1.890715910e-03f);
  p = fmaf(p, t, -6.751993671e-03f);
  p = fmaf(p, t, -4.832238483e-04f);
  p = fmaf(p, t, 3.672166169e-02f);
  p = fmaf(p, t, 2.879839949e-02f);
  p = fmaf(p, t, -3.314045668e-01f);
  p = fmaf(p, t, 6.724079847e-01f);
  const float q = p * u;
  const float C = -0.72134752044448170368f;                       // -0.5 * log2(e)
  const float CL = (float)(-0.72134752044448170368 - (double)C);  // and what fp32 dropped of it
  const float ahi = s * C;
  const float alo = fmaf(sl, C, fmaf(s, CL, fmaf(s, C, -ahi)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ahi));
  float r = e * q;
  r = fmaf(r * 0.69314718055994530942f, alo, r);
  return x < 0.0f ? r : 1.0f - r;
}
// log(x) for fp32: the same range reduction and polynomial degree as every fp32 logf (m in [2/3, 4/3), log1p(f) =
// f + f*f*p(f), p a degree-8 minimax fit made here), with zero / subnormal / negative / inf / NaN sent to the library
// function by ONE unsigned compare instead of five select instructions on the hot path.
static __device__ __noinline__ float f_log_special(float x) { return logf(x); }  // out of line: keeps the cold block out of unrolled loops
__device__ __forceinline__ float f_log(float x) {
  const unsigned ix = __float_as_uint(x);
  if (ix - 0x00800000u >= 0x7f000000u) return f_log_special(x);
  const unsigned e = (ix - 0x3f2aaaabu) & 0xff800000u;
  const float f = __uint_as_float(ix - e) - 1.0f;
  const float fe = (float)(int)e;
  float p = -0.1294892579317093f;
  p = fmaf(p, f, 0.1400475949048996f);
  p = fmaf(p, f, -0.1216716319322586f);
  p = fmaf(p, f, 0.14001160860061646f);
  p = fmaf(p, f, -0.16682304441928864f);
  p = fmaf(p, f, 0.20010747015476227f);
  p = fmaf(p, f, -0.24999716877937317f);
  p = fmaf(p, f, 0.3333320915699005f);
  p = fmaf(p, f, -0.5f);
  p = fmaf(f, p * f, f);
  return fmaf(fe, 8.26295829e-08f, p);  // + exponent * ln2 (e still carries its 2^23 scale)
}
__device__ __forceinline__ double f_log(double x) { return log(x); }

// ------------------------------------------------------------------------------------------------
// Two fp32 lanes per register pair: Blackwell's packed fp32 arithmetic (PTX fma / mul / add .f32x2 -> SASS FFMA2 /
// FMUL2 / FADD2) carries two elements per issue slot.  The fused elementwise kernels are instruction-issue bound for
// transcendental chains (Black-Scholes: 152 instructions per option at 89 % issue utilisation, HBM at 62 %), so the
// generated `eval2` bodies evaluate elements v and v+1 together; division, sqrt and exp stay the scalar IEEE /
// library code per half, log and normcdf are the packed twins of the functions above (same constants, same order
// of operations, so a lane gives the same bits whichever body evaluates it).
// ------------------------------------------------------------------------------------------------
struct f2 {
  float2 v;
  __device__ __forceinline__ f2() {}
  __device__ __forceinline__ f2(float a, float b) { v.x = a; v.y = b; }
  __device__ __forceinline__ explicit f2(float a) { v.x = a; v.y = a; }
  __device__ __forceinline__ explicit f2(float2 a) : v(a) {}
};
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return f2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return f2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ f2 neg2(f2 a) { return f2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return f2(__fadd2_rn(a.v, neg2(b).v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return f2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, float c) { return f2(__ffma2_rn(a.v, b.v, make_float2(c, c))); }
__device__ __forceinline__ f2 fma2(f2 a, float b, float c) { return f2(__ffma2_rn(a.v, make_float2(b, b), make_float2(c, c))); }
__device__ __forceinline__ f2 operator/(f2 a, f2 b) { return f2(a.v.x / b.v.x, a.v.y / b.v.y); }
__device__ __forceinline__ f2 f_sqrt(f2 a) { return f2(sqrtf(a.v.x), sqrtf(a.v.y)); }
__device__ __forceinline__ f2 f_exp(f2 a) { return f2(expf(a.v.x), expf(a.v.y)); }
__device__ __forceinline__ f2 f_abs(f2 a) { return f2(fabsf(a.v.x), fabsf(a.v.y)); }
__device__ __forceinline__ f2 f_normcdf(f2 x) {   // f_normcdf(float) above, two lanes at a time
  f2 z;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(z.v.x) : "f"(fabsf(x.v.x)), "f"(14.5f));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(z.v.y) : "f"(fabsf(x.v.y)), "f"(14.5f));
  const f2 s = z * z, sl = fma2(z, z, neg2(s));
  const f2 d = z + f2(2.0f);
  f2 u;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u.v.x) : "f"(d.v.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u.v.y) : "f"(d.v.y));
  u = fma2(u, fma2(neg2(d), u, 1.0f), u);
  const f2 t = fma2(u, -4.0f, 1.0f);
  f2 p = fma2(t, -7.689554332e-05f, 2.991535075e-05f);
  p = fma2(p, t, 7.246770547e-04f);
  p = fma2(p, t, 8.852431783e-04f);
  p = fma2(p, t, -1.890715910e-03f);
  p = fma2(p, t, -6.751993671e-03f);
  p = fma2(p, t, -4.832238483e-04f);
  p = fma2(p, t, 3.672166169e-02f);
  p = fma2(p, t, 2.879839949e-02f);
  p = fma2(p, t, -3.314045668e-01f);
  p = fma2(p, t, 6.724079847e-01f);
  const f2 q = p * u;
  const float C = -0.72134752044448170368f;
  const float CL = (float)(-0.72134752044448170368 - (double)C);
  const f2 ahi = s * f2(C);
  const f2 alo = fma2(sl, f2(C), fma2(s, f2(CL), fma2(s, f2(C), neg2(ahi))));
  f2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.v.x) : "f"(ahi.v.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.v.y) : "f"(ahi.v.y));
  f2 r = e * q;
  r = fma2(r * f2(0.69314718055994530942f), alo, r);
  const f2 om = f2(1.0f) - r;
  return f2(x.v.x < 0.0f ? r.v.x : om.v.x, x.v.y < 0.0f ? r.v.y : om.v.y);
}
__device__ __forceinline__ f2 f_log(f2 x) {       // f_log(float) above, two lanes at a time
  const unsigned ix = __float_as_uint(x.v.x), iy = __float_as_uint(x.v.y);
  if ((ix - 0x00800000u >= 0x7f000000u) | (iy - 0x00800000u >= 0x7f000000u)) return f2(f_log(x.v.x), f_log(x.v.y));
  const unsigned ex = (ix - 0x3f2aaaabu) & 0xff800000u, ey = (iy - 0x3f2aaaabu) & 0xff800000u;
  const f2 f = f2(__uint_as_float(ix - ex), __uint_as_float(iy - ey)) + f2(-1.0f);
  const f2 fe((float)(int)ex, (float)(int)ey);
  f2 p = fma2(f, -0.1294892579317093f, 0.1400475949048996f);
  p = fma2(p, f, -0.1216716319322586f);
  p = fma2(p, f, 0.14001160860061646f);
  p = fma2(p, f, -0.16682304441928864f);
  p = fma2(p, f, 0.20010747015476227f);
  p = fma2(p, f, -0.24999716877937317f);
  p = fma2(p, f, 0.3333320915699005f);
  p = fma2(p, f, -0.5f);
  p = fma2(f, p * f, f);
  return fma2(fe, f2(8.26295829e-08f), p);
}
__device__ __forceinline__ double f_normcdf(double x) { return normcdf(x); }
__device__ __forceinline__ float f_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double f_abs(double x) { return fabs(x); }
__device__ __forceinline__ int f_abs(int x) { return x < 0 ? -x : x; }
__device__ __forceinline__ i64 f_abs(i64 x) { return x < 0 ? -x : x; }
__device__ __forceinline__ float f_abs(cfloat x) { return hypotf(x.re, x.im); }
__device__ __forceinline__ float f_abs2(float x) { return x * x; }
__device__ __forceinline__ double f_abs2(double x) { return x * x; }
__device__ __forceinline__ int f_abs2(int x) { return x * x; }
__device__ __forceinline__ i64 f_abs2(i64 x) { return x * x; }
__device__ __forceinline__ float f_abs2(cfloat x) { return x.re * x.re + x.im * x.im; }
__device__ __forceinline__ cfloat f_conj(cfloat x) { return cfloat(x.re, -x.im); }
template <class T> __device__ __forceinline__ T f_conj(T x) { return x; }
__device__ __forceinline__ float f_real(cfloat x) { return x.re; }
__device__ __forceinline__ float f_imag(cfloat x) { return x.im; }
template <class T> __device__ __forceinline__ T f_real(T x) { return x; }
template <class T> __device__ __forceinline__ T f_imag(T) { return (T)0; }
__device__ __forceinline__ float f_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double f_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float f_mod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double f_mod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ int f_mod(int a, int b) { return a % b; }
__device__ __forceinline__ i64 f_mod(i64 a, i64 b) { return a % b; }
template <class T> __device__ __forceinline__ T f_max(T a, T b) { return a > b ? a : b; }  // scalar_internal.h: cuda::std::max
template <class T> __device__ __forceinline__ T f_min(T a, T b) { return a < b ? a : b; }
__device__ __forceinline__ float f_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double f_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float f_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double f_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ cfloat f_expj(float x) { float s, c; sincosf(x, &s, &c); return cfloat(c, s); }
__device__ __forceinline__ cfloat f_exp(cfloat x) { float e = expf(x.re), s, c; sincosf(x.im, &s, &c); return cfloat(e * c, e * s); }
__device__ __forceinline__ float f_exp(float x) { return expf(x); }
__device__ __forceinline__ double f_exp(double x) { return exp(x); }
__device__ __forceinline__ float f_round(float x) { return roundf(x); }
__device__ __forceinline__ double f_round(double x) { return round(x); }
template <class T> __device__ __forceinline__ bool f_isnan(T) { return false; }
__device__ __forceinline__ bool f_isnan(float x) { return x != x; }
__device__ __forceinline__ bool f_isnan(double x) { return x != x; }
__device__ __forceinline__ bool f_isnan(cfloat x) { return x.re != x.re || x.im != x.im; }
template <class T> __device__ __forceinline__ bool f_isinf(T) { return false; }
__device__ __forceinline__ bool f_isinf(float x) { return isinf(x); }
__device__ __forceinline__ bool f_isinf(double x) { return isinf(x); }
__device__ __forceinline__ bool f_isinf(cfloat x) { return isinf(x.re) || isinf(x.im); }

// ------------------------------------------------------------------------------------------------
// reduction operators: accumulate (per element), combine (associative, order-preserving), warp stage
// ------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T shfl_xor_t(T x, int m) {
  // bounce any trivially copyable record through 32-bit shuffles
  enum { W = (sizeof(T) + 3) / 4 };
  union { T t; u32 w[W]; } u;
  u.t = x;
#pragma unroll
  for (int i = 0; i < W; ++i) u.w[i] = __shfl_xor_sync(0xffffffffu, u.w[i], m);
  return u.t;
}
template <class T> __device__ __forceinline__ T shfl_down_t(T x, int d) {
  enum { W = (sizeof(T) + 3) / 4 };
  union { T t; u32 w[W]; } u;
  u.t = x;
#pragma unroll
  for (int i = 0; i < W; ++i) u.w[i] = __shfl_down_sync(0xffffffffu, u.w[i], d);
  return u.t;
}

template <class T> struct Limits;
template <> struct Limits<float> {
  static __device__ __forceinline__ float lowest() { return -__int_as_float(0x7f800000); }
  static __device__ __forceinline__ float highest() { return __int_as_float(0x7f800000); }
};
template <> struct Limits<double> {
  static __device__ __forceinline__ double lowest() { return -__longlong_as_double(0x7ff0000000000000LL); }
  static __device__ __forceinline__ double highest() { return __longlong_as_double(0x7ff0000000000000LL); }
};
template <> struct Limits<int> {
  static __device__ __forceinline__ int lowest() { return (int)0x80000000; }
  static __device__ __forceinline__ int highest() { return 0x7fffffff; }
};
template <> struct Limits<i64> {
  static __device__ __forceinline__ i64 lowest() { return (i64)0x8000000000000000ULL; }
  static __device__ __forceinline__ i64 highest() { return 0x7fffffffffffffffLL; }
};
template <> struct Limits<unsigned char> {
  static __device__ __forceinline__ unsigned char lowest() { return 0; }
  static __device__ __forceinline__ unsigned char highest() { return 255; }
};

// Every op exposes:
//   acc_t                       running state of one thread / partial record (<= 16 bytes)
//   init()
//   step(acc, x, flat_index)    fold one element; elements arrive in increasing flat_index per thread
//   merge(a, b)                 a = a (+) b where every index in a precedes / is unrelated to b's (commutative here:
//                               arg ops compare indices explicitly, so order does not matter)
//   warp(acc)                   all-lanes result of the warp-wide merge
//   result_t / finish(acc)      value written to `out`
//   HAS_INDEX, index(acc)
template <class T> struct OpSum {
  typedef T acc_t; typedef T result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ acc_t init() { return (T)0; }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64) { a = a + x; }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) { a = a + b; }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) a = a + shfl_xor_t(a, m);
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};
template <> struct OpSum<cfloat> {
  typedef cfloat acc_t; typedef cfloat result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ acc_t init() { return cfloat(0.f, 0.f); }
  static __device__ __forceinline__ void step(acc_t &a, cfloat x, i64) { a = a + x; }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) { a = a + b; }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) a = a + shfl_xor_t(a, m);
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};
// integer sums ride redux.sync
template <> __device__ __forceinline__ int OpSum<int>::warp(int a) { return __reduce_add_sync(0xffffffffu, a); }

template <class T> struct OpProd {
  typedef T acc_t; typedef T result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ acc_t init() { return (T)1; }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64) { a = a * x; }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) { a = a * b; }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) a = a * shfl_xor_t(a, m);
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};
template <> struct OpProd<cfloat> {
  typedef cfloat acc_t; typedef cfloat result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ acc_t init() { return cfloat(1.f, 0.f); }
  static __device__ __forceinline__ void step(acc_t &a, cfloat x, i64) { a = a * x; }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) { a = a * b; }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) a = a * shfl_xor_t(a, m);
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};

// max / min by value (reference: reduceOpMax / reduceOpMin, transforms/reduce.h:127-177)
template <class T, bool IS_MAX> struct OpExt {
  typedef T acc_t; typedef T result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ bool better(T a, T b) { return IS_MAX ? (a > b) : (a < b); }
  static __device__ __forceinline__ acc_t init() { return IS_MAX ? Limits<T>::lowest() : Limits<T>::highest(); }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64) { a = better(x, a) ? x : a; }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) { a = better(b, a) ? b : a; }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { T o = shfl_xor_t(a, m); a = better(o, a) ? o : a; }
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};
// fp32 max / min: one redux.sync.{max,min}.f32 (sm_100a) instead of five shuffle rounds.  NaN-free
// inputs give the same value as the compare-select form; with a NaN the shuffle form is used.
template <> __device__ __forceinline__ float OpExt<float, true>::warp(float a) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(a));
  return r;
}
template <> __device__ __forceinline__ float OpExt<float, false>::warp(float a) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(a));
  return r;
}
template <> __device__ __forceinline__ int OpExt<int, true>::warp(int a) { return __reduce_max_sync(0xffffffffu, a); }
template <> __device__ __forceinline__ int OpExt<int, false>::warp(int a) { return __reduce_min_sync(0xffffffffu, a); }

// argmax / argmin: (value, absolute flat index), lowest index wins ties — the HostExecutor result
// (std::max_element / std::min_element, transforms/host_algorithms.h:262-296).
template <class T> struct __align__(16) ArgAcc { T val; i64 idx; };
template <class T, bool IS_MAX> struct OpArg {
  typedef ArgAcc<T> acc_t; typedef T result_t; enum { HAS_INDEX = 1 };
  static __device__ __forceinline__ bool better(T a, T b) { return IS_MAX ? (a > b) : (a < b); }
  static __device__ __forceinline__ acc_t init() {
    acc_t a; a.val = IS_MAX ? Limits<T>::lowest() : Limits<T>::highest(); a.idx = 0x7fffffffffffffffLL; return a;
  }
  // per-thread elements arrive in increasing index order: strict compare keeps the first occurrence
  static __device__ __forceinline__ void step(acc_t &a, T x, i64 i) {
    if (better(x, a.val) || (a.idx == 0x7fffffffffffffffLL && x == a.val)) { a.val = x; a.idx = i; }
  }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) {
    if (better(b.val, a.val) || (b.val == a.val && b.idx < a.idx)) a = b;
  }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { acc_t o = shfl_xor_t(a, m); merge(a, o); }
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a.val; }
  static __device__ __forceinline__ i64 index(acc_t a) { return a.idx; }
};
// fp32 arg ops: the warp stage is two redux.sync — value first, then the lowest index among the
// lanes that hold that value (u32 halves of the index, high word first).
template <bool IS_MAX> __device__ __forceinline__ ArgAcc<float> warp_arg_f32(ArgAcc<float> a) {
  float best;
  if (IS_MAX) asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(best) : "f"(a.val));
  else asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(best) : "f"(a.val));
  const bool mine = (a.val == best);
  u32 hi = mine ? (u32)((u64)a.idx >> 32) : 0xffffffffu;
  u32 hmin = __reduce_min_sync(0xffffffffu, hi);
  u32 lo = (mine && hi == hmin) ? (u32)(u64)a.idx : 0xffffffffu;
  u32 lmin = __reduce_min_sync(0xffffffffu, lo);
  ArgAcc<float> r; r.val = best; r.idx = (i64)(((u64)hmin << 32) | (u64)lmin);
  return r;
}
template <> __device__ __forceinline__ ArgAcc<float> OpArg<float, true>::warp(ArgAcc<float> a) { return warp_arg_f32<true>(a); }
template <> __device__ __forceinline__ ArgAcc<float> OpArg<float, false>::warp(ArgAcc<float> a) { return warp_arg_f32<false>(a); }

// argminmax: min and max with their indices in ONE read (reference: argminmax_impl -> cub_dualargreduce,
// transforms/reduce.h:1090-1109, cub.h:1439-1491; it carries 32-byte tuples through CUB and first materialises operator
// inputs).  State = the two arg states side by side; lowest index wins ties on both sides.
template <class T> struct __align__(16) ArgMMAcc { T vmin, vmax; i64 imin, imax; };
template <class T> struct OpArgMinMax {
  typedef ArgMMAcc<T> acc_t; typedef T result_t; enum { HAS_INDEX = 1, DUAL = 1 };
  typedef OpArg<T, false> Mn; typedef OpArg<T, true> Mx;
  static __device__ __forceinline__ acc_t init() {
    acc_t a; a.vmin = Limits<T>::highest(); a.vmax = Limits<T>::lowest(); a.imin = a.imax = 0x7fffffffffffffffLL; return a;
  }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64 i) {
    if (x < a.vmin || (a.imin == 0x7fffffffffffffffLL && x == a.vmin)) { a.vmin = x; a.imin = i; }
    if (x > a.vmax || (a.imax == 0x7fffffffffffffffLL && x == a.vmax)) { a.vmax = x; a.imax = i; }
  }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) {
    if (b.vmin < a.vmin || (b.vmin == a.vmin && b.imin < a.imin)) { a.vmin = b.vmin; a.imin = b.imin; }
    if (b.vmax > a.vmax || (b.vmax == a.vmax && b.imax < a.imax)) { a.vmax = b.vmax; a.imax = b.imax; }
  }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
    ArgAcc<T> lo, hi;
    lo.val = a.vmin; lo.idx = a.imin; hi.val = a.vmax; hi.idx = a.imax;
    lo = Mn::warp(lo);
    hi = Mx::warp(hi);
    acc_t r; r.vmin = lo.val; r.imin = lo.idx; r.vmax = hi.val; r.imax = hi.idx;
    return r;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a.vmin; }
  static __device__ __forceinline__ i64 index(acc_t a) { return a.imin; }
};

// Block steps of the arg ops.  A thread's V x U elements of a tile arrive in increasing index order, and a running
// extremum is replaced less and less often as the walk goes on (the k-th element of random data improves it with
// probability 1 / k), so the common case is decided by ONE min / max per element: the block's extremum (NaNs ignored, as the
// per-element compare ignores them) is tested against the state, and only a block that can change the state takes the exact
// per-element steps.  Same result as stepping every element; ~1 instruction per element instead of ~10.
template <class T> __device__ __forceinline__ T nn_max(T a, T b) { return a > b ? a : b; }
template <class T> __device__ __forceinline__ T nn_min(T a, T b) { return a < b ? a : b; }
template <> __device__ __forceinline__ float nn_max<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ float nn_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double nn_max<double>(double a, double b) { return fmax(a, b); }
template <> __device__ __forceinline__ double nn_min<double>(double a, double b) { return fmin(a, b); }
template <class Op> struct OpBlock { enum { ON = 0 }; };
template <class T, bool IS_MAX> struct OpBlock<OpArg<T, IS_MAX> > {
  enum { ON = 1 };
  typedef OpArg<T, IS_MAX> Op;
  template <int N> static __device__ __forceinline__ void step(typename Op::acc_t &a, const T *x, i64 i0) {
    T m = x[0];
#pragma unroll
    for (int k = 1; k < N; ++k) m = IS_MAX ? nn_max<T>(m, x[k]) : nn_min<T>(m, x[k]);
    if (Op::better(m, a.val) || a.idx == 0x7fffffffffffffffLL) {
#pragma unroll
      for (int k = 0; k < N; ++k) Op::step(a, x[k], i0 + k);
    }
  }
};
template <class T> struct OpBlock<OpArgMinMax<T> > {
  enum { ON = 1 };
  typedef OpArgMinMax<T> Op;
  template <int N> static __device__ __forceinline__ void step(typename Op::acc_t &a, const T *x, i64 i0) {
    T lo = x[0], hi = x[0];
#pragma unroll
    for (int k = 1; k < N; ++k) { lo = nn_min<T>(lo, x[k]); hi = nn_max<T>(hi, x[k]); }
    if (lo < a.vmin || hi > a.vmax || a.imin == 0x7fffffffffffffffLL || a.imax == 0x7fffffffffffffffLL) {
#pragma unroll
      for (int k = 0; k < N; ++k) Op::step(a, x[k], i0 + k);
    }
  }
};

template <class Op, int N, class T, bool ON = (OpBlock<Op>::ON != 0)> struct BlockStepCall {
  static __device__ __forceinline__ void go(typename Op::acc_t &, const T *, i64) {}
};
template <class Op, int N, class T> struct BlockStepCall<Op, N, T, true> {
  static __device__ __forceinline__ void go(typename Op::acc_t &a, const T *x, i64 i0) { OpBlock<Op>::template step<N>(a, x, i0); }
};
template <class Op, int N, class T> __device__ __forceinline__ void block_step(typename Op::acc_t &a, const T *x, i64 i0) { BlockStepCall<Op, N, T>::go(a, x, i0); }

// any / all (reference: reduceOpAny / reduceOpAll, transforms/reduce.h:153-177; result is 0 / 1 in the
// output tensor's type).  The warp stage is a vote.
template <class T, bool IS_ANY> struct OpLogic {
  typedef int acc_t; typedef int result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ acc_t init() { return IS_ANY ? 0 : 1; }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64) { a = IS_ANY ? (a | (int)nonzero(x)) : (a & (int)nonzero(x)); }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) { a = IS_ANY ? (a | b) : (a & b); }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
    return IS_ANY ? (int)__any_sync(0xffffffffu, a) : (int)__all_sync(0xffffffffu, a);
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};

// running (max, sum of exp(x - max)) of a row: the one-pass form of the reference's softmax statistics
// (max_impl, then sum_impl(exp(in - max)), transforms/reduce.h:371-373,440-441).  `out` receives the max, the
// auxiliary output (the index slot of the parameter block, here in the value type) the reciprocal of the sum.
template <class T> struct __align__(2 * sizeof(T)) LseAcc { T m, s; };
template <class T> struct OpLse {
  typedef LseAcc<T> acc_t; typedef T result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ acc_t init() { acc_t a; a.m = Limits<T>::lowest(); a.s = (T)0; return a; }
  // exp(-|hi - lo|), with equal operands (two -inf included) giving exactly 1
  static __device__ __forceinline__ T decay(T d, bool same) { return same ? (T)1 : f_exp(d > (T)0 ? -d : d); }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64) {
    const T d = x - a.m;
    const T e = decay(d, x == a.m);
    if (d > (T)0) { a.s = a.s * e + (T)1; a.m = x; }
    else a.s = a.s + e;
  }
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) {
    if (b.s == (T)0) return;             // identity
    if (a.s == (T)0) { a = b; return; }
    const T d = b.m - a.m;
    const T e = decay(d, b.m == a.m);
    if (d > (T)0) { a.s = a.s * e + b.s; a.m = b.m; }
    else a.s = a.s + b.s * e;
  }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { acc_t o = shfl_xor_t(a, m); merge(a, o); }
    return a;
  }
  static __device__ __forceinline__ result_t finish(acc_t a) { return a.m; }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};
// one-pass variance: every accumulator lane keeps SHIFTED sums about a pivot K — s1 = sum (x - K), s2 = sum |x - K|^2,
// cnt — i.e. the state (mean = K + s1/cnt, M2 = s2 - |s1|^2/cnt, cnt) in a form whose per-element update is three
// arithmetic instructions and no division.  The pivot is the lane's first element and the state is re-centred
// (K <- mean, s1 <- ~0, s2 <- M2) after 8 elements and then every 64, so the cancellation in s2 - |s1|^2/cnt is bounded by the 64
// elements since the last re-centring whatever the data (|mean| >> stddev, outliers).  Partial states are combined by
// re-expressing one about the other's (re-centred) pivot — Chan's parallel formula without its divisions; moving a
// pivot is exact algebra for any offset, so an approximate reciprocal (MUFU.RCP) picks the new pivot.  Numerically the stable way to get a variance from ONE read when the row cannot be
// kept on chip for the reference's two passes (transforms/reduce.h:1406-1444): rows longer than shared memory, full
// tensors, strided / permuted rows through the coalesced walkers.  `out` receives M2 / (N - ddof) (and its sqrt for
// stdd) through the usual post-processing.  fp32 and complex<float>.  (A Welford update with a correctly rounded
// reciprocal per element was measured first: same accuracy, ~30 instructions per element, 0.64-0.72 of the HBM peak.)
template <class T> struct VarAcc { T k; T s1; float s2; int cnt; };   // 16 bytes (fp32) / 24 bytes (complex<float>)
__device__ __forceinline__ float var_abs2dot(float d, float e) { return d * e; }
__device__ __forceinline__ float var_abs2dot(cfloat d, cfloat e) { return d.re * e.re + d.im * e.im; }   // Re(d * conj(e))
template <class T> struct OpVar {
  typedef VarAcc<T> acc_t; typedef float result_t; enum { HAS_INDEX = 0 };
  static __device__ __forceinline__ T zero() { return cvt<T>(0.0f); }
  static __device__ __forceinline__ acc_t init() { acc_t a; a.k = zero(); a.s1 = zero(); a.s2 = 0.f; a.cnt = 0; return a; }
  // move the pivot by m: exact algebra for ANY m (sum (x-K-m) = s1 - n m, sum |x-K-m|^2 = s2 - 2 Re(conj(m) s1) + n |m|^2)
  static __device__ __forceinline__ void shift(acc_t &a, T m) {
    const float n = (float)a.cnt;
    a.k = a.k + m;
    a.s2 = a.s2 + (n * var_abs2dot(m, m) - 2.f * var_abs2dot(m, a.s1));
    a.s1 = a.s1 - m * n;
  }
  // pivot <- (an approximation of) the mean: MUFU.RCP is enough, the shift is exact for whatever m comes out
  static __device__ __forceinline__ void recentre(acc_t &a) {
    if (a.cnt == 0) return;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)a.cnt));
    shift(a, a.s1 * r);
  }
  static __device__ __forceinline__ void step(acc_t &a, T x, i64) {
    if (a.cnt == 0) a.k = x;
    const T d = x - a.k;
    a.s1 = a.s1 + d;
    a.s2 += var_abs2dot(d, d);
    if (((++a.cnt) & 63) == 8) {
      // after 8, 72, 136, ... elements (a bad first pivot — an outlier — is dropped early): K <- K + m, m = s1 / cnt by
      // MUFU.RCP.  The compact form drops the residual s1 - cnt*m = s1 * O(2^-23) (the hot loop stays small: the full
      // shift() inlined V x U times cost 10 % on long rows); merges and the final value use the exact shift
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)a.cnt));
      const T m = a.s1 * r;
      a.k = a.k + m;
      a.s2 -= var_abs2dot(a.s1, m);
      a.s1 = zero();
    }
  }
  // a <- a (+) b: a's pivot goes to a's mean, b is re-expressed about it (no division), the sums add
  static __device__ __forceinline__ void merge(acc_t &a, acc_t b) {
    if (b.cnt == 0) return;               // identity
    if (a.cnt == 0) { a = b; return; }
    recentre(a);
    shift(b, a.k - b.k);
    a.s1 = a.s1 + b.s1;
    a.s2 += b.s2;
    a.cnt += b.cnt;
  }
  static __device__ __forceinline__ acc_t warp(acc_t a) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { acc_t o = shfl_xor_t(a, m); merge(a, o); }
    return a;
  }
  // M2 = s2 - |s1|^2 / n about a pivot that is the mean to rounding: the correction term is tiny, the division exact
  static __device__ __forceinline__ result_t finish(acc_t a) {
    if (a.cnt == 0) return 0.f;
    recentre(a);
    return a.s2 - var_abs2dot(a.s1, a.s1) / (float)a.cnt;
  }
  static __device__ __forceinline__ i64 index(acc_t) { return 0; }
};

// the second (value, index) pair of a dual op (none by default)
template <class Op, class OutT> struct DualStore {
  static __device__ __forceinline__ void go(const RedParams &, i64, i64, const i64 *, typename Op::acc_t) {}
};
template <class T, class OutT> struct DualStore<OpArgMinMax<T>, OutT> {
  static __device__ __forceinline__ void go(const RedParams &p, i64 oo, i64 io, const i64 *bidx, ArgMMAcc<T> a) {
    ((OutT *)p.out2)[oo] = cvt<OutT>(a.vmax);
    i64 ix = a.imax;
    if (ix == 0x7fffffffffffffffLL) {   // nothing compared greater than the identity: first element, as std::max_element
      i64 fb = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) fb += bidx[d] * p.bflat[d];
      ix = p.idx_base + fb * p.R;
    }
    ((i64 *)p.idx2)[io] = ix;
  }
};
// second output of an op (none by default)
template <class Op> struct AuxStore {
  static __device__ __forceinline__ void go(void *, i64, typename Op::acc_t) {}
};
template <class T> struct AuxStore<OpLse<T> > {
  static __device__ __forceinline__ void go(void *ptr, i64 off, LseAcc<T> a) { ((T *)ptr)[off] = (T)1 / a.s; }  // the apply pass multiplies
};


// ------------------------------------------------------------------------------------------------
// post-processing of a finished sum (mean / var divisor, stdd sqrt) and the store
// ------------------------------------------------------------------------------------------------
template <class R> struct Post {
  static __device__ __forceinline__ R go(R x, const RedParams &) { return x; }
};
template <> struct Post<float> {
  static __device__ __forceinline__ float go(float x, const RedParams &p) {
    if (p.post_div) x = x / p.post_scale_f;
    if (p.post_sqrt) x = sqrtf(x);
    return x;
  }
};
template <> struct Post<double> {
  static __device__ __forceinline__ double go(double x, const RedParams &p) {
    if (p.post_div) x = x / p.post_scale_d;
    if (p.post_sqrt) x = sqrt(x);
    return x;
  }
};
template <> struct Post<cfloat> {
  static __device__ __forceinline__ cfloat go(cfloat x, const RedParams &p) {
    if (p.post_div) x = x / p.post_scale_f;
    return x;
  }
};

// 32-byte partial record exchanged between GPUs (mxb_reduce_partial / mxb_reduce_finalize)
struct __align__(16) PartialRec { u32 w[8]; };

// fused exchange: one writer thread stores the record into every rank's buffer (its own included), fences at system
// scope, then bumps that rank's arrival counter for this source — a release over NVLink peer mappings
__device__ __forceinline__ void push_record(const PeerPush &pp, const PartialRec &rec) {
  const u32 e = *(volatile const u32 *)pp.epoch + 1u;
  const int slot = (int)(e & 1u);
  for (int r = 0; r < pp.world; ++r)
    ((PartialRec *)pp.rec[r])[((size_t)slot * pp.world + pp.rank) * KMAXITEMS + pp.item] = rec;
  __threadfence_system();
  for (int r = 0; r < pp.world; ++r) atomicAdd_system(pp.flag[r] + pp.rank, 1u);
}

// flat index -> per-dim indices (row-major over n dims of sizes sz[])
__device__ __forceinline__ void decomp(i64 flat, int n, const i64 *sz, i64 *idx) {
  if (n <= 1) {  // the common case (one collapsed dim): no division at all
    idx[0] = flat;
#pragma unroll
    for (int d = 1; d < KMAXD; ++d) idx[d] = 0;
    return;
  }
#pragma unroll
  for (int d = KMAXD - 1; d >= 0; --d) {
    if (d < n) {
      if (d == 0) { idx[0] = flat; }
      else { const i64 q = flat / sz[d]; idx[d] = flat - q * sz[d]; flat = q; }
    } else idx[d] = 0;
  }
}

// what a slab's 32-byte record holds: the accumulator's own bits, except for the one-pass variance, whose record is
// the (mean re, mean im, M2, n) quadruple of doubles that the multi-GPU fold combines with Chan's formula
template <class Op> struct PartialPack {
  static __device__ __forceinline__ void go(PartialRec &rec, typename Op::acc_t acc) {
    union { PartialRec rec; typename Op::acc_t a; } u;
#pragma unroll
    for (int i = 0; i < 8; ++i) u.rec.w[i] = 0;
    u.a = acc;
    rec = u.rec;
  }
};
__device__ __forceinline__ double var_re(float x) { return (double)x; }
__device__ __forceinline__ double var_im(float) { return 0.0; }
__device__ __forceinline__ double var_re(cfloat x) { return (double)x.re; }
__device__ __forceinline__ double var_im(cfloat x) { return (double)x.im; }
template <class T> struct PartialPack<OpVar<T> > {
  static __device__ __forceinline__ void go(PartialRec &rec, VarAcc<T> a) {
    union { PartialRec rec; double d[4]; } u;
    const double n = (double)a.cnt, inv = a.cnt ? 1.0 / n : 0.0;
    const double s1r = var_re(a.s1), s1i = var_im(a.s1);
    u.d[0] = var_re(a.k) + s1r * inv;                           // mean
    u.d[1] = var_im(a.k) + s1i * inv;
    u.d[2] = (double)a.s2 - (s1r * s1r + s1i * s1i) * inv;      // M2 about that mean
    u.d[3] = n;
    rec = u.rec;
  }
};

template <class Op, class OutT>
__device__ __forceinline__ void store_result(const RedParams &p, i64 b, typename Op::acc_t acc) {
  if (p.raw_partial) {
    union { PartialRec rec; int pad_; } u;
    PartialPack<Op>::go(u.rec, acc);
    if (p.raw_partial == 2) {
      push_record(p.peer, u.rec);
      return;
    }
    ((PartialRec *)p.out.ptr)[b] = u.rec;
    return;
  }
  i64 bidx[KMAXD];
  i64 oo = 0, io = 0;
  if (p.nb == 1) {
    bidx[0] = b;
    bidx[1] = bidx[2] = bidx[3] = 0;
    oo = b * p.out.bs[0];
    io = b * p.idx.bs[0];
  } else {
    decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) {
      if (d < p.nb) { oo += bidx[d] * p.out.bs[d]; io += bidx[d] * p.idx.bs[d]; }
    }
  }
  typename Op::result_t r = Post<typename Op::result_t>::go(Op::finish(acc), p);
  ((OutT *)p.out.ptr)[oo] = cvt<OutT>(r);
  if (Op::HAS_INDEX) {
    i64 ix = Op::index(acc);
    // nothing compared better than the identity (row of -inf / +inf): first element, as std::max_element
    if (ix == 0x7fffffffffffffffLL) {
      i64 fb = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) fb += bidx[d] * p.bflat[d];
      ix = p.idx_base + fb * p.R;
    }
    ((i64 *)p.idx.ptr)[io] = ix;
  }
  DualStore<Op, OutT>::go(p, oo, io, bidx, acc);
  AuxStore<Op>::go(p.idx.ptr, io, acc);
}

// read a record another CTA wrote during this launch (L2, never a stale L1 line)
template <class T> __device__ __forceinline__ T ld_cg_t(const T *p) {
  enum { W = sizeof(T) / 4 };
  union { T t; u32 w[W]; } u;
#pragma unroll
  for (int i = 0; i < W; ++i) u.w[i] = __ldcg((const u32 *)p + i);
  return u.t;
}
template <class T> __device__ __forceinline__ void st_cg_t(T *p, T v) {
  enum { W = sizeof(T) / 4 };
  union { T t; u32 w[W]; } u;
  u.t = v;
#pragma unroll
  for (int i = 0; i < W; ++i) __stcg((u32 *)p + i, u.w[i]);
}

// CTA-wide merge; result valid in thread 0.  `smem` holds one acc_t per warp (32 max).
template <class Op> __device__ __forceinline__ typename Op::acc_t cta_merge(typename Op::acc_t a, typename Op::acc_t *smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  a = Op::warp(a);
  if (nwarp == 1) return a;
  __syncthreads();  // previous use of smem is over
  if (lane == 0) smem[warp] = a;
  __syncthreads();
  if (warp == 0) {
    typename Op::acc_t v = (lane < nwarp) ? smem[lane] : Op::init();
    a = Op::warp(v);
  }
  return a;
}

// merge across the G lanes (power of two) that share a row; G == 32 takes the op's redux / vote path
template <class Op> __device__ __forceinline__ typename Op::acc_t group_merge(typename Op::acc_t a, int G) {
  if (G == 32) return Op::warp(a);
  for (int m = G >> 1; m > 0; m >>= 1) {
    typename Op::acc_t o = shfl_xor_t(a, m);
    Op::merge(a, o);
  }
  return a;
}

// per-leaf byte offset of an outer (non-innermost) reduce index
template <class E>
__device__ __forceinline__ void outer_bases(const RedParams &p, i64 o, const char *const *base, const char **rb) {
  i64 oidx[KMAXD];
  decomp(o, p.nr - 1, p.rsz, oidx);
#pragma unroll
  for (int k = 0; k < E::NL; ++k) {
    i64 off = 0;
#pragma unroll
    for (int d = 0; d < KMAXD - 1; ++d) if (d < p.nr - 1) off += oidx[d] * p.leaf[k].rs[d];
    rb[k] = base[k] + off * E::leaf_bytes(k);
  }
}

// The last CTA out of a dynamically dealt reduce_inner launch folds the partials in item order (deterministic).
template <class Op, class OutT>
__device__ __forceinline__ void dyn_fold(const RedParams &p, i64 nfold, typename Op::acc_t *s_acc) {
  typedef typename Op::acc_t acc_t;
  __threadfence();
  const acc_t *ws = (const acc_t *)p.ws;
  if (p.B == 1) {
    // independent loads per trip: the fold is one CTA against L2 latency
    acc_t a = Op::init();
    const i64 nt = blockDim.x;
    i64 i = threadIdx.x;
    // eight loads in flight per trip (the states here are one or two words): a 512 MB slab leaves 4096 partials, two trips
    for (; i + 7 * nt < nfold; i += 8 * nt) {
      acc_t x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = ld_cg_t(&ws[i + k * nt]);
#pragma unroll
      for (int k = 0; k < 8; ++k) Op::merge(a, x[k]);
    }
    for (; i + 3 * nt < nfold; i += 4 * nt) {
      const acc_t x0 = ld_cg_t(&ws[i]), x1 = ld_cg_t(&ws[i + nt]), x2 = ld_cg_t(&ws[i + 2 * nt]), x3 = ld_cg_t(&ws[i + 3 * nt]);
      Op::merge(a, x0); Op::merge(a, x1); Op::merge(a, x2); Op::merge(a, x3);
    }
    for (; i < nfold; i += nt) Op::merge(a, ld_cg_t(&ws[i]));
    a = cta_merge<Op>(a, s_acc);
    if (threadIdx.x == 0) store_result<Op, OutT>(p, 0, a);
  } else {
    // a warp per row: lane-strided partials, then the op's warp stage
    const i64 S = p.splits;
    const int lane = (int)threadIdx.x & 31, nwarp = (int)blockDim.x >> 5;
    for (i64 b = threadIdx.x >> 5; b < p.B; b += nwarp) {
      acc_t a = Op::init();
      for (i64 i = lane; i < S; i += 32) Op::merge(a, ld_cg_t(&ws[b * S + i]));
      a = Op::warp(a);
      if (lane == 0) store_result<Op, OutT>(p, b, a);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1: reduce_inner — rows whose innermost reduce dim is the vector dim (unit stride or broadcast in
// every leaf when V > 1; any stride when V == 1, which makes V == 1 the universal fallback).
//
// TEAM == 0: one CTA — or `splits` CTAs — per row.  Thread stage: V lane accumulators, U vector
//   loads per leaf in flight.  Warp stage: shuffle / redux.sync / vote.  CTA stage: shared memory.
//   Grid stage (splits > 1), same launch: each CTA drops its partial in the workspace and takes a
//   ticket; the last to arrive folds the partials in a fixed order (deterministic) and stores.  The
//   ticket is an atomicInc that wraps to zero, so no memset is needed between launches.
// TEAM == 1: one warp per row, no shared memory, no barrier (short rows).
// ------------------------------------------------------------------------------------------------
template <class E, class Op, class OutT, int V, int U, int TEAM, bool UNIT>
__device__ __forceinline__ void reduce_inner_body_impl(const RedParams &p) {
  typedef typename Op::acc_t acc_t;
  __shared__ acc_t s_acc[32];
  __shared__ acc_t s_acc2[TEAM == 0 ? 2 : 1][TEAM == 0 ? 32 : 1];
  __shared__ int s_last;

  // TEAM == 1: G lanes per row (G = p.tx, a power of two <= 32), 32 / G rows per warp; the row loop is warp-uniform
  // (groups past the last row redo the last row and skip the store) so the full-mask shuffles stay legal
  const int G = TEAM == 0 ? (int)blockDim.x : p.tx;
  const int nthr = G;
  const int tid = TEAM == 0 ? (int)threadIdx.x : (int)(threadIdx.x & (G - 1));
  const int gid = TEAM == 0 ? 0 : (int)((threadIdx.x & 31) / G);
  const int gpw = TEAM == 0 ? 1 : 32 / G;
  const i64 wb0 = TEAM == 0 ? (i64)blockIdx.x : ((i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * gpw;
  const i64 wstep = TEAM == 0 ? (i64)gridDim.x : (i64)gridDim.x * (blockDim.x >> 5) * gpw;
  const int nr = p.nr;
  const i64 L = p.rsz[nr - 1];  // innermost run
  const i64 Lv = L / V;         // full vectors per run
  const i64 tail = L - Lv * V;  // leftover scalars per run
  const i64 O = p.R / L;        // runs per row
  const i64 S = TEAM == 0 ? (i64)p.splits : 1;
  const i64 work = p.B * S;

  __shared__ i64 s_next[2];
  // the dynamic deal is compiled in for one- and two-word accumulators only (sum / prod / max / min / any / all): with the
  // wide states (arg ops, variance, log-sum-exp) its bookkeeping cost a CTA per SM in registers (64 -> 80), which the deal
  // does not win back; the host never passes a work counter for those
  constexpr bool DYN_OK = TEAM == 0 && sizeof(acc_t) <= 8;
  const bool dyn = DYN_OK && p.work_ctr != nullptr;
  const bool carry = dyn && p.carry_items != 0;
  bool first_item = true;
  acc_t acc[V];
  int par = 0;
  for (i64 wb = wb0; wb < work;) {
    i64 wnext = wb + wstep;
    // dynamic deal: the draw for the NEXT item goes out first and stays in a register until the CTA stage, so its L2
    // round trip hides behind this item's loads (parking it in shared memory right away would stall thread 0 on the
    // atomic's return before it has issued a single load: measured 6 % on 64 KB items)
    u32 drawn = 0;
    if (dyn && threadIdx.x == 0) drawn = atomicAdd(p.work_ctr, 1u);
    i64 w = wb + gid;
    const bool valid = w < work;
    if (!valid) w = work - 1;
    const i64 b = w / S;
    const i64 s = w - b * S;
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 row0 = p.idx_base;  // flat index of the row's first element
    if (p.nb == 1) {
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        base[k] = (const char *)p.leaf[k].ptr + b * p.leaf[k].bs[0] * E::leaf_bytes(k);
        inner[k] = p.leaf[k].rs[nr - 1];
      }
      if (Op::HAS_INDEX) row0 += b * p.bflat[0] * p.R;
    } else {
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
        inner[k] = p.leaf[k].rs[nr - 1];
      }
      if (Op::HAS_INDEX) {
#pragma unroll
        for (int d = 0; d < KMAXD; ++d) if (d < p.nb) row0 += bidx[d] * p.bflat[d] * p.R;
      }
    }
    if (!carry || first_item) {
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = Op::init();
      first_item = false;
    }

    // vector steps of the row are numbered q = o * Lv + jv, Q of them
    const i64 Q = O * Lv;

    if (nr == 1) {
      // Tiles of nthr * U vectors are dealt round-robin to the row's splits (neighbouring CTAs stream neighbouring
      // tiles, which keeps the DRAM pages they open shared); inside a tile every thread has U loads in flight.
      const i64 tile = (i64)nthr * U;
      const i64 nfull = Q / tile;
      const i64 cht = p.chunk_tiles;   // > 0: this split owns a contiguous run of tiles, else tiles are dealt round-robin
      const i64 tfirst = cht > 0 ? s * cht : s;
      const i64 tend = cht > 0 ? ((tfirst + cht < nfull) ? tfirst + cht : nfull) : nfull;
      const i64 tstep = cht > 0 ? 1 : S;
      for (i64 t = tfirst; t < tend; t += tstep) {
        const i64 q = t * tile + tid;
        typename E::template Regs<V> r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) E::template loadv<V, UNIT>(r[u], base, inner, (q + (i64)u * nthr) * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j0 = (q + (i64)u * nthr) * V;
          if (OpBlock<Op>::ON) {   // arg ops: one accumulator, the vector's V elements as a block
            typename E::value_type xs[V];
#pragma unroll
            for (int v = 0; v < V; ++v) xs[v] = E::template eval<V>(r[u], v, p.c);
            block_step<Op, V>(acc[0], xs, row0 + j0);
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) Op::step(acc[v], E::template eval<V>(r[u], v, p.c), row0 + j0 + v);
          }
        }
      }
      if (s == (cht > 0 ? S - 1 : nfull % S)) {  // the ragged last tile goes to the next split in turn (contiguous runs: to the last split)
        for (i64 q = nfull * tile + tid; q < Q; q += nthr) {
          typename E::template Regs<V> r;
          E::template loadv<V, UNIT>(r, base, inner, q * V);
#pragma unroll
          for (int v = 0; v < V; ++v) Op::step(acc[v], E::template eval<V>(r, v, p.c), row0 + q * V + v);
        }
      }
    } else {
      // several runs per row: this split owns the contiguous range [q0, q1) of vector steps
      const i64 per = (Q + S - 1) / S;
      const i64 q0 = s * per;
      const i64 q1 = (q0 + per < Q) ? (q0 + per) : Q;
      for (i64 q = q0 + tid; q < q1; q += nthr) {
        const i64 o = q / Lv, jv = q - o * Lv;
        const char *rb[E::NL];
        outer_bases<E>(p, o, base, rb);
        typename E::template Regs<V> r;
        E::template loadv<V, UNIT>(r, rb, inner, jv * V);
#pragma unroll
        for (int v = 0; v < V; ++v) Op::step(acc[v], E::template eval<V>(r, v, p.c), row0 + o * L + jv * V + v);
      }
    }
    // scalar tails of the runs (L % V elements each), taken by split 0
    if (V > 1 && tail > 0 && s == 0) {
      const i64 nt = O * tail;
      for (i64 t = tid; t < nt; t += nthr) {
        const i64 o = t / tail, j = Lv * V + (t - o * tail);
        const char *rb[E::NL];
        outer_bases<E>(p, o, base, rb);
        typename E::template Regs<1> r;
        E::template loadv<1, UNIT>(r, rb, inner, j);
        // this index may precede ones the thread already folded: merge, do not step
        acc_t one = Op::init();
        Op::step(one, E::template eval<1>(r, 0, p.c), row0 + o * L + j);
        Op::merge(acc[0], one);
      }
    }
    if (carry) {
      // one accumulator across the items of this CTA: only the next draw has to be published
      if (threadIdx.x == 0) s_next[par] = (i64)gridDim.x + (i64)drawn;
      __syncthreads();
      wnext = s_next[par];
      par ^= 1;
      wb = wnext;
      continue;
    }
#pragma unroll
    for (int v = 1; v < V; ++v) Op::merge(acc[0], acc[v]);
    // streaming part of this CTA's last work item is over: let the next kernel on the stream start launching while
    // the warp / CTA / grid stages finish (it still waits for this grid to complete before touching memory)
    if (!dyn && wnext >= work) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (dyn && threadIdx.x == 0) s_next[par] = (i64)gridDim.x + (i64)drawn;   // published by the CTA stage's barriers

    if (TEAM == 1) {
      acc_t tot = group_merge<Op>(acc[0], G);
      if (tid == 0 && valid) store_result<Op, OutT>(p, b, tot);
    } else {
      // CTA stage, ONE barrier per item: the per-warp partials are double-buffered by item parity (the barrier also
      // publishes thread 0's draw of the next item)
      const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5, nwarp = ((int)blockDim.x + 31) >> 5;
      const acc_t wtot = Op::warp(acc[0]);
      if (lane == 0) s_acc2[par][warp] = wtot;
      __syncthreads();
      if (warp == 0) {
        const acc_t v = (lane < nwarp) ? s_acc2[par][lane] : Op::init();
        const acc_t tot = nwarp > 1 ? Op::warp(v) : wtot;
        if (lane == 0) {
          if (S == 1) {
            store_result<Op, OutT>(p, b, tot);
          } else {
            st_cg_t(&((acc_t *)p.ws)[b * S + s], tot);
            if (!dyn) {   // static deal (every CTA has one item): per-row ticket, the last arrival folds the row
              __threadfence();
              const u32 t = atomicInc(&p.tickets[b], (u32)(S - 1));
              s_last = (t == (u32)(S - 1));
            }
          }
        }
      }
      if (S > 1 && !dyn) {
        __syncthreads();
        if (s_last) {
          __threadfence();
          acc_t a = Op::init();
          for (i64 i = tid; i < S; i += nthr) Op::merge(a, ld_cg_t(&((const acc_t *)p.ws)[b * S + i]));
          a = cta_merge<Op>(a, s_acc);
          if (tid == 0) store_result<Op, OutT>(p, b, a);
        }
      }
    }
    if (dyn) {
      wnext = s_next[par];
      if (wnext >= work) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    par ^= 1;
    wb = wnext;
  }
  if (carry) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (first_item) {
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = Op::init();
    }
#pragma unroll
    for (int v = 1; v < V; ++v) Op::merge(acc[0], acc[v]);
    const acc_t tot = cta_merge<Op>(acc[0], s_acc);
    if (threadIdx.x == 0) st_cg_t(&((acc_t *)p.ws)[blockIdx.x], tot);
  }
  if (dyn) {
    // exit ticket: every draw of this CTA has returned and its partials are written; the last CTA out rewinds the work
    // counter for the next launch and folds the partials of every row (S per row; one per CTA in carry mode) in a fixed
    // order (deterministic)
    const i64 nfold = carry ? (i64)gridDim.x : S;
    if (threadIdx.x == 0) {
      __threadfence();
      const bool last = atomicInc(p.work_ctr + 1, gridDim.x - 1) == gridDim.x - 1;
      if (last) atomicExch(p.work_ctr, 0u);
      s_last = last ? 1 : 0;
    }
    __syncthreads();
    if (s_last && (S > 1 || carry)) dyn_fold<Op, OutT>(p, nfold, s_acc);
  }
}

// ------------------------------------------------------------------------------------------------
// K2: reduce_outer — the reduce dims are strided and a BATCH dim is the vector dim (unit stride in
// every leaf): column sums, `sum(permute(t,{2,0,1}),{2})`.  The host rotates that batch dim to the
// last batch position of the parameter block (strides travel with it, so nothing is copied) and the
// output strides follow.  A CTA is blockDim.x/TY column groups x TY reduce lanes; each thread owns V
// adjacent columns, walks the reduce index with stride TY (U loads in flight), keeps V accumulators
// in registers, and the TY partials meet in shared memory.  Loads stay coalesced along the unit-stride
// dim however the reduce dims are strided.
// ------------------------------------------------------------------------------------------------
template <class E, class Op, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void reduce_outer_body_impl(const RedParams &p) {
  typedef typename Op::acc_t acc_t;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  acc_t *s_part = (acc_t *)s_dyn;  // [TY][TX*V]
  __shared__ int s_last;

  const int TX = p.tx, TY = (int)blockDim.x / TX;
  const int tx = (int)threadIdx.x % TX, ty = (int)threadIdx.x / TX;
  const int nb = p.nb, nr = p.nr;
  const i64 C = p.bsz[nb - 1];                 // vector (column) dim extent
  const i64 tile = (i64)TX * V;                // columns per CTA
  const i64 ctiles = (C + tile - 1) / tile;
  const i64 Bo = p.B / C;                      // product of the other batch dims
  // few output tiles and a long reduce dim (column sums of a tall matrix): `splits` CTAs share a tile, each walks the
  // reduce index with stride splits * TY, and the last CTA to finish folds the partial vectors in split order
  const i64 S = p.splits;
  const i64 RS = S * TY;                       // stride of one lane through the reduce index
  const i64 work = Bo * ctiles * S;

  for (i64 w = blockIdx.x; w < work; w += gridDim.x) {
    const i64 tile_id = w / S, s = w - tile_id * S;
    const i64 bo = tile_id / ctiles, ct = tile_id - bo * ctiles;
    const i64 c0 = ct * tile + (i64)tx * V;    // first column of this thread
    const bool active = c0 < C;
    const bool fullvec = c0 + V <= C;
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 bidx[KMAXD];
    decomp(bo, nb - 1, p.bsz, bidx);
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD - 1; ++d) if (d < nb - 1) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      inner[k] = p.leaf[k].bs[nb - 1];
    }
    acc_t acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = Op::init();
    // flat index of (batch row, reduce r) = (sum_d bidx[d]*bflat[d]) * R + r; column c adds c*bflat[nb-1]
    i64 rowflat = p.idx_base + c0 * p.bflat[nb - 1] * p.R;
    if (Op::HAS_INDEX) {
#pragma unroll
      for (int d = 0; d < KMAXD - 1; ++d) if (d < nb - 1) rowflat += bidx[d] * p.bflat[d] * p.R;
    }
    const i64 colflat = p.bflat[nb - 1] * p.R;  // flat-index step between adjacent columns

    if (active) {
      if (fullvec) {
        i64 r = s * TY + ty;
        for (; r + (i64)(U - 1) * RS < p.R; r += (i64)U * RS) {
          typename E::template Regs<V> reg[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const char *rb[E::NL];
            if (nr == 1) {
#pragma unroll
              for (int k = 0; k < E::NL; ++k) rb[k] = base[k] + (r + (i64)u * RS) * p.leaf[k].rs[0] * E::leaf_bytes(k);
            } else {
              i64 ridx[KMAXD];
              decomp(r + (i64)u * RS, nr, p.rsz, ridx);
#pragma unroll
              for (int k = 0; k < E::NL; ++k) {
                i64 off = 0;
#pragma unroll
                for (int d = 0; d < KMAXD; ++d) if (d < nr) off += ridx[d] * p.leaf[k].rs[d];
                rb[k] = base[k] + off * E::leaf_bytes(k);
              }
            }
            E::template loadv<V, UNIT>(reg[u], rb, inner, c0);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int v = 0; v < V; ++v)
              Op::step(acc[v], E::template eval<V>(reg[u], v, p.c), rowflat + (i64)v * colflat + r + (i64)u * RS);
          }
        }
        for (; r < p.R; r += RS) {
          const char *rb[E::NL];
          i64 ridx[KMAXD];
          decomp(r, nr, p.rsz, ridx);
#pragma unroll
          for (int k = 0; k < E::NL; ++k) {
            i64 off = 0;
#pragma unroll
            for (int d = 0; d < KMAXD; ++d) if (d < nr) off += ridx[d] * p.leaf[k].rs[d];
            rb[k] = base[k] + off * E::leaf_bytes(k);
          }
          typename E::template Regs<V> reg;
          E::template loadv<V, UNIT>(reg, rb, inner, c0);
#pragma unroll
          for (int v = 0; v < V; ++v) Op::step(acc[v], E::template eval<V>(reg, v, p.c), rowflat + (i64)v * colflat + r);
        }
      } else {
        // ragged last vector of the column dim: scalar lanes
        for (i64 r = s * TY + ty; r < p.R; r += RS) {
          const char *rb[E::NL];
          i64 ridx[KMAXD];
          decomp(r, nr, p.rsz, ridx);
#pragma unroll
          for (int k = 0; k < E::NL; ++k) {
            i64 off = 0;
#pragma unroll
            for (int d = 0; d < KMAXD; ++d) if (d < nr) off += ridx[d] * p.leaf[k].rs[d];
            rb[k] = base[k] + off * E::leaf_bytes(k);
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            if (c0 + v < C) {
              typename E::template Regs<1> reg;
              E::template loadv<1, UNIT>(reg, rb, inner, c0 + v);
              Op::step(acc[v], E::template eval<1>(reg, 0, p.c), rowflat + (i64)v * colflat + r);
            }
          }
        }
      }
    }
    // TY partials per column meet in shared memory (fixed order: ty = 0, 1, ...)
    if (TY > 1) {
      __syncthreads();
#pragma unroll
      for (int v = 0; v < V; ++v) s_part[(ty * TX + tx) * V + v] = acc[v];
      __syncthreads();
      if (ty == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          acc_t a = acc[v];
          for (int y = 1; y < TY; ++y) Op::merge(a, s_part[(y * TX + tx) * V + v]);
          acc[v] = a;
        }
      }
    }
    if (S == 1) {
      if (ty == 0 && active) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c0 + v < C) store_result<Op, OutT>(p, bo * C + c0 + v, acc[v]);
      }
    } else {
      acc_t *ws = (acc_t *)p.ws + (size_t)tile_id * S * tile;
      if (ty == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) st_cg_t(&ws[(size_t)s * tile + (size_t)tx * V + v], acc[v]);
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const u32 t = atomicInc(&p.tickets[tile_id], (u32)(S - 1));
        s_last = (t == (u32)(S - 1));
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (ty == 0 && active) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            if (c0 + v < C) {
              acc_t a = Op::init();
              for (i64 s2 = 0; s2 < S; ++s2) Op::merge(a, ld_cg_t(&ws[(size_t)s2 * tile + (size_t)tx * V + v]));
              store_result<Op, OutT>(p, bo * C + c0 + v, a);
            }
          }
        }
      }
      __syncthreads();  // s_last is reused by the next work item
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3: var_inner_smem — variance / standard deviation of rows that fit in shared memory, with the
// reference's exact two-pass arithmetic (transforms/reduce.h:1406-1444: mean, then the sum of
// |x - mean|^2, then / (N - ddof)) but ONE read of the row from HBM: pass 1 evaluates the expression,
// parks the values in shared memory and sums them; pass 2 re-reads shared memory only.
// ------------------------------------------------------------------------------------------------
template <class T> struct AbsDev2 {
  typedef T real_t;
  static __device__ __forceinline__ T go(T x, T m) { const T d = x - m; return d * d; }
};
template <> struct AbsDev2<cfloat> {
  typedef float real_t;
  static __device__ __forceinline__ float go(cfloat x, cfloat m) { const float a = x.re - m.re, b = x.im - m.im; return a * a + b * b; }
};
template <class T> struct MeanDiv {
  static __device__ __forceinline__ T go(T s, i64 n) { return s / (T)n; }
};
template <> struct MeanDiv<cfloat> {
  static __device__ __forceinline__ cfloat go(cfloat s, i64 n) { return s / (float)n; }
};

template <class E, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void var_inner_smem_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  typedef typename AbsDev2<T>::real_t RT;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  T *s_row = (T *)s_dyn;  // [R]
  __shared__ T s_sum[32];
  __shared__ RT s_sq[32];
  __shared__ T s_mean;

  const int nthr = blockDim.x, tid = threadIdx.x;
  const int nr = p.nr;
  const i64 L = p.rsz[nr - 1], Lv = L / V, tail = L - Lv * V, O = p.R / L;

  for (i64 b = blockIdx.x; b < p.B; b += gridDim.x) {
    const char *base[E::NL];
    i64 inner[E::NL];
    {
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
        inner[k] = p.leaf[k].rs[nr - 1];
      }
    }
    T acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = OpSum<T>::init();
    const i64 Q = O * Lv;
    if (nr == 1) {
      i64 q = tid;
      for (; q + (i64)(U - 1) * nthr < Q; q += (i64)U * nthr) {
        typename E::template Regs<V> r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) E::template loadv<V, UNIT>(r[u], base, inner, (q + (i64)u * nthr) * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          Vec<T, V> x;
#pragma unroll
          for (int v = 0; v < V; ++v) { x.v[v] = E::template eval<V>(r[u], v, p.c); acc[v] = acc[v] + x.v[v]; }
          *(Vec<T, V> *)(s_row + (q + (i64)u * nthr) * V) = x;
        }
      }
      for (; q < Q; q += nthr) {
        typename E::template Regs<V> r;
        E::template loadv<V, UNIT>(r, base, inner, q * V);
        Vec<T, V> x;
#pragma unroll
        for (int v = 0; v < V; ++v) { x.v[v] = E::template eval<V>(r, v, p.c); acc[v] = acc[v] + x.v[v]; }
        *(Vec<T, V> *)(s_row + q * V) = x;
      }
    } else {
      for (i64 q = tid; q < Q; q += nthr) {
        const i64 o = q / Lv, jv = q - o * Lv;
        const char *rb[E::NL];
        outer_bases<E>(p, o, base, rb);
        typename E::template Regs<V> r;
        E::template loadv<V, UNIT>(r, rb, inner, jv * V);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T x = E::template eval<V>(r, v, p.c);
          acc[v] = acc[v] + x;
          s_row[o * L + jv * V + v] = x;
        }
      }
    }
    if (V > 1 && tail > 0) {
      const i64 nt = O * tail;
      for (i64 t = tid; t < nt; t += nthr) {
        const i64 o = t / tail, j = Lv * V + (t - o * tail);
        const char *rb[E::NL];
        outer_bases<E>(p, o, base, rb);
        typename E::template Regs<1> r;
        E::template loadv<1, UNIT>(r, rb, inner, j);
        const T x = E::template eval<1>(r, 0, p.c);
        acc[0] = acc[0] + x;
        s_row[o * L + j] = x;
      }
    }
#pragma unroll
    for (int v = 1; v < V; ++v) acc[0] = acc[0] + acc[v];
    T tot = cta_merge<OpSum<T> >(acc[0], s_sum);
    if (tid == 0) s_mean = MeanDiv<T>::go(tot, p.R);
    __syncthreads();  // also orders the s_row writes before pass 2
    const T mean = s_mean;
    RT sq = (RT)0;
    for (i64 i = tid; i < p.R; i += nthr) sq += AbsDev2<T>::go(s_row[i], mean);
    sq = cta_merge<OpSum<RT> >(sq, s_sq);
    if (tid == 0) {
      RT r = sq / (RT)p.post_scale_d;
      if (p.post_sqrt) r = f_sqrt(r);
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
      i64 oo = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
      ((OutT *)p.out.ptr)[oo] = cvt<OutT>(r);
    }
    __syncthreads();  // s_row / s_mean are reused by the next row
  }
}

// ------------------------------------------------------------------------------------------------
// K3r: var_inner_reg — the same exact two-pass variance for rows of up to blockDim.x * IPT vectors, with
// the evaluated row parked in REGISTERS instead of shared memory: every thread issues its IPT vector loads
// back to back (all in flight at once), keeps the values, and both passes run out of registers.  One read of
// HBM, no shared-memory traffic beyond the two CTA reductions.  Needs a single contiguous reduce dim whose
// length is a multiple of V (the host checks); everything else goes to var_inner_smem or the two-launch path.
// ------------------------------------------------------------------------------------------------
template <class E, class OutT, int V, int IPT, bool UNIT>
__device__ __forceinline__ void var_inner_reg_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  typedef typename AbsDev2<T>::real_t RT;
  __shared__ T s_sum[32];
  __shared__ RT s_sq[32];
  __shared__ T s_mean;
  const int nthr = blockDim.x, tid = threadIdx.x;
  const i64 Lv = p.rsz[0] / V;

  for (i64 b = blockIdx.x; b < p.B; b += gridDim.x) {
    const char *base[E::NL];
    i64 inner[E::NL];
    {
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
        inner[k] = p.leaf[k].rs[0];
      }
    }
    typename E::template Regs<V> r[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * nthr;
      if (q < Lv) E::template loadv<V, UNIT>(r[i], base, inner, q * V);
    }
    Vec<T, V> x[IPT];
    T acc = OpSum<T>::init();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * nthr;
      if (q < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) { x[i].v[v] = E::template eval<V>(r[i], v, p.c); acc = acc + x[i].v[v]; }
      }
    }
    const T tot = cta_merge<OpSum<T> >(acc, s_sum);
    if (tid == 0) s_mean = MeanDiv<T>::go(tot, p.R);
    __syncthreads();
    const T mean = s_mean;
    RT sq = (RT)0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * nthr;
      if (q < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) sq += AbsDev2<T>::go(x[i].v[v], mean);
      }
    }
    sq = cta_merge<OpSum<RT> >(sq, s_sq);
    if (tid == 0) {
      RT res = sq / (RT)p.post_scale_d;
      if (p.post_sqrt) res = f_sqrt(res);
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
      i64 oo = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
      ((OutT *)p.out.ptr)[oo] = cvt<OutT>(res);
    }
    __syncthreads();  // s_mean is reused by the next row
  }
}

// ------------------------------------------------------------------------------------------------
// K3g: var_group — the exact two-pass variance for SHORT rows (<= 32 * IPT vectors): G lanes of a warp share a row
// (G = p.tx), 32 / G rows per warp, the row lives in registers, both reductions are shuffles inside the group; no
// shared memory, no barrier.  One HBM read.
// ------------------------------------------------------------------------------------------------
template <class E, class OutT, int V, int IPT, bool UNIT>
__device__ __forceinline__ void var_group_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  typedef typename AbsDev2<T>::real_t RT;
  const int G = p.tx, gpw = 32 / G;
  const int tid = (int)(threadIdx.x & (G - 1)), gid = (int)((threadIdx.x & 31) / G);
  const i64 wb0 = ((i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * gpw;
  const i64 wstep = (i64)gridDim.x * (blockDim.x >> 5) * gpw;
  const i64 Lv = p.rsz[0] / V;
  for (i64 wb = wb0; wb < p.B; wb += wstep) {
    i64 b = wb + gid;
    const bool valid = b < p.B;
    if (!valid) b = p.B - 1;
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      inner[k] = p.leaf[k].rs[0];
    }
    typename E::template Regs<V> r[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * G;
      if (q < Lv) E::template loadv<V, UNIT>(r[i], base, inner, q * V);
    }
    Vec<T, V> x[IPT];
    T acc = OpSum<T>::init();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * G < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) { x[i].v[v] = E::template eval<V>(r[i], v, p.c); acc = acc + x[i].v[v]; }
      }
    }
    for (int m = G >> 1; m > 0; m >>= 1) acc = acc + shfl_xor_t(acc, m);
    const T mean = MeanDiv<T>::go(acc, p.R);
    RT sq = (RT)0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * G < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) sq += AbsDev2<T>::go(x[i].v[v], mean);
      }
    }
    for (int m = G >> 1; m > 0; m >>= 1) sq += shfl_xor_t(sq, m);
    if (tid == 0 && valid) {
      RT res = sq / (RT)p.post_scale_d;
      if (p.post_sqrt) res = f_sqrt(res);
      i64 oo = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
      ((OutT *)p.out.ptr)[oo] = cvt<OutT>(res);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// S1: softmax_group / softmax_reg — out(b, :) = exp(x(b, :) - max_b) / sum(exp(x(b, :) - max_b)), the arithmetic of
// the reference's softmax_impl (transforms/reduce.h:362-445: max_impl, sum_impl(exp(in - max)), then the divide) in
// ONE launch with one HBM read and one write: the evaluated row is parked in registers (IPT vectors per lane),
// max -> exp -> sum -> divide run out of them.  softmax_group: G = p.tx lanes of a warp own a row (32 / G rows per
// warp, shuffles only); softmax_reg: a CTA owns a row (two barriers per row, per-warp partials double-buffered by
// row parity and folded by every thread in the same order).  Needs a single contiguous reduce dim; anything else
// takes the two-launch path (OpLse statistics + the elementwise kernel).
// ------------------------------------------------------------------------------------------------
template <class OutT, int V>
__device__ __forceinline__ void softmax_store(OutT *orow, i64 q, i64 ostride, const Vec<OutT, V> &o) {
  if (V == 1) orow[q * ostride] = o.v[0];
  else StBytes<(int)sizeof(OutT) * V>::st(orow + q * V, &o);
}

template <class E, class OutT, int V, int IPT, bool UNIT>
__device__ __forceinline__ void softmax_group_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  const int G = p.tx, gpw = 32 / G;
  const int tid = (int)(threadIdx.x & (G - 1)), gid = (int)((threadIdx.x & 31) / G);
  const i64 wb0 = ((i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * gpw;
  const i64 wstep = (i64)gridDim.x * (blockDim.x >> 5) * gpw;
  const i64 Lv = p.rsz[0] / V;
  const i64 ostride = p.out_rs[0];
  for (i64 wb = wb0; wb < p.B; wb += wstep) {
    i64 b = wb + gid;
    const bool valid = b < p.B;
    if (!valid) b = p.B - 1;
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      inner[k] = p.leaf[k].rs[0];
    }
    i64 oo = 0;
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    OutT *orow = (OutT *)p.out.ptr + oo;
    typename E::template Regs<V> r[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * G;
      if (q < Lv) E::template loadv<V, UNIT>(r[i], base, inner, q * V);
    }
    Vec<T, V> x[IPT];
    T mx = Limits<T>::lowest();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * G < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) { x[i].v[v] = E::template eval<V>(r[i], v, p.c); mx = x[i].v[v] > mx ? x[i].v[v] : mx; }
      }
    }
    for (int m = G >> 1; m > 0; m >>= 1) { const T o = shfl_xor_t(mx, m); mx = o > mx ? o : mx; }
    T sum = (T)0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * G < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) { x[i].v[v] = f_exp(x[i].v[v] - mx); sum = sum + x[i].v[v]; }
      }
    }
    for (int m = G >> 1; m > 0; m >>= 1) sum = sum + shfl_xor_t(sum, m);
    const T inv = (T)1 / sum;   // one IEEE reciprocal per row; exp * inv is within 1 ulp of the reference's exp / sum
    if (valid) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const i64 q = tid + (i64)i * G;
        if (q < Lv) {
          Vec<OutT, V> o;
#pragma unroll
          for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(x[i].v[v] * inv);
          softmax_store<OutT, V>(orow, q, ostride, o);
        }
      }
    }
  }
}

template <class E, class OutT, int V, int IPT, bool UNIT>
__device__ __forceinline__ void softmax_reg_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  __shared__ T s_max[2][32];
  __shared__ T s_sum[2][32];
  const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const i64 Lv = p.rsz[0] / V;
  const i64 ostride = p.out_rs[0];
  int par = 0;
  for (i64 b = blockIdx.x; b < p.B; b += gridDim.x, par ^= 1) {
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      inner[k] = p.leaf[k].rs[0];
    }
    i64 oo = 0;
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    OutT *orow = (OutT *)p.out.ptr + oo;
    typename E::template Regs<V> r[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * nthr;
      if (q < Lv) E::template loadv<V, UNIT>(r[i], base, inner, q * V);
    }
    Vec<T, V> x[IPT];
    T mx = Limits<T>::lowest();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * nthr < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) { x[i].v[v] = E::template eval<V>(r[i], v, p.c); mx = x[i].v[v] > mx ? x[i].v[v] : mx; }
      }
    }
    mx = OpExt<T, true>::warp(mx);
    if (lane == 0) s_max[par][warp] = mx;
    __syncthreads();  // (A)
    for (int w = 0; w < nwarp; ++w) { const T o = s_max[par][w]; mx = o > mx ? o : mx; }
    T sum = (T)0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * nthr < Lv) {
#pragma unroll
        for (int v = 0; v < V; ++v) { x[i].v[v] = f_exp(x[i].v[v] - mx); sum = sum + x[i].v[v]; }
      }
    }
    sum = OpSum<T>::warp(sum);
    if (lane == 0) s_sum[par][warp] = sum;
    __syncthreads();  // (B)
    T tot = (T)0;
    for (int w = 0; w < nwarp; ++w) tot = tot + s_sum[par][w];  // same order in every thread
    const T inv = (T)1 / tot;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 q = tid + (i64)i * nthr;
      if (q < Lv) {
        Vec<OutT, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(x[i].v[v] * inv);
        softmax_store<OutT, V>(orow, q, ostride, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3t: var_inner_tma — variance / stdd of contiguous rows of a plain tensor, rows staged ONCE in shared memory by
// the TMA engine.  Persistent grid, one CTA per SM, a ring of `stages` row buffers: thread 0 arms an mbarrier with
// the row's byte count and issues one `cp.async.bulk.shared::cluster.global` (SASS UBLKCP) per row; the whole CTA
// waits on the barrier's phase, runs the reference's exact two passes (mean, then sum |x - mean|^2) out of shared
// registers (one pass over shared memory), and the buffer is re-armed with the row `stages` ahead as soon as
// every thread holds its share (after the first of the two barriers per row).  The copy engine keeps
// (stages - 1) rows in flight per SM while the SM computes, so HBM never idles between a row's two passes —
// which is what limits var_inner_smem / var_inner_reg (load phase and compute phase alternate there).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class Tin> struct Widen { typedef Tin type; };
template <> struct Widen<__nv_bfloat16> { typedef float type; };
template <> struct Widen<__half> { typedef float type; };

template <class Tin, class OutT, int IPT>
__device__ __forceinline__ void var_inner_tma_body(const RedParams &p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  typedef typename Widen<Tin>::type T;
  typedef typename AbsDev2<T>::real_t RT;
  enum { V = 16 / (int)sizeof(Tin) };
  extern __shared__ __align__(128) unsigned char s_dyn[];
  __shared__ T s_sum[2][32];    // per-warp partials, double-buffered by row parity: two barriers per row suffice
  __shared__ RT s_sq[2][32];
  u64 *full = (u64 *)s_dyn;     // one mbarrier per stage (first 128 bytes)
  unsigned char *buf = s_dyn + 128;
  const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int stages = p.splits;  // ring depth chosen by the host from the row size
  const i64 R = p.R;
  const u32 rowbytes = (u32)(R * (i64)sizeof(Tin));
  const u32 rowstride = (rowbytes + 127u) & ~127u;
  const i64 Rv = R / V;

  auto row_src = [&](i64 b) -> const char * {
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
    i64 off = 0;
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[0].bs[d];
    return (const char *)p.leaf[0].ptr + off * (i64)sizeof(Tin);
  };

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < stages; ++k) {
      const i64 b = (i64)blockIdx.x + (i64)k * gridDim.x;
      if (b < p.B) {
        mbar_expect_tx(&full[k], rowbytes);
        bulk_g2s(buf + (size_t)k * rowstride, row_src(b), rowbytes, &full[k]);
      }
    }
  }
  i64 k = 0;
  for (i64 b = blockIdx.x; b < p.B; b += gridDim.x, ++k) {
    const int s = (int)(k % stages);
    const int par = (int)(k & 1);
    mbar_wait(&full[s], (u32)((k / stages) & 1));
    // the thread's share of the row: shared memory -> registers, once
    const Vec<Tin, V> *x = (const Vec<Tin, V> *)(buf + (size_t)s * rowstride);
    Vec<Tin, V> q[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const i64 j = tid + (i64)i * nthr;
      if (j < Rv) q[i] = x[j];
    }
    // pass 1: mean
    T acc = OpSum<T>::init();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * nthr < Rv) {
#pragma unroll
        for (int v = 0; v < V; ++v) acc = acc + cvt<T>(q[i].v[v]);
      }
    }
    acc = OpSum<T>::warp(acc);
    if (lane == 0) s_sum[par][warp] = acc;
    __syncthreads();  // (A) partials visible; every thread holds its share in registers, so buffer s is free
    if (tid == 0) {
      const i64 nb2 = b + (i64)stages * gridDim.x;
      if (nb2 < p.B) {
        mbar_expect_tx(&full[s], rowbytes);
        bulk_g2s(buf + (size_t)s * rowstride, row_src(nb2), rowbytes, &full[s]);
      }
    }
    T tot = OpSum<T>::init();
    for (int w = 0; w < nwarp; ++w) tot = tot + s_sum[par][w];  // same order in every thread
    const T mean = MeanDiv<T>::go(tot, R);
    // pass 2: sum of |x - mean|^2 out of registers
    RT sq = (RT)0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (tid + (i64)i * nthr < Rv) {
#pragma unroll
        for (int v = 0; v < V; ++v) sq += AbsDev2<T>::go(cvt<T>(q[i].v[v]), mean);
      }
    }
    sq = OpSum<RT>::warp(sq);
    if (lane == 0) s_sq[par][warp] = sq;
    __syncthreads();  // (B)
    if (tid == 0) {
      RT tsq = (RT)0;
      for (int w = 0; w < nwarp; ++w) tsq += s_sq[par][w];
      RT res = tsq / (RT)p.post_scale_d;
      if (p.post_sqrt) res = f_sqrt(res);
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
      i64 oo = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
      ((OutT *)p.out.ptr)[oo] = cvt<OutT>(res);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2t: reduce_outer_tma — reduce_outer (strided / permuted reduce dim, a batch dim is the unit-stride vector dim) for
// a plain tensor with ONE collapsed reduce dim, tiles staged in shared memory by the TMA engine.  A CTA owns a strip
// of TX 16-byte chunks of the vector dim (an "item" = one strip of one outer batch index) and walks the reduce index
// in chunks of RT rows: one stage of the ring = RT rows x TX chunks.  Warp-specialised: the LAST warp of the CTA is
// the producer — it waits for a stage's `empty` mbarrier, arms its `full` mbarrier with the byte count and issues the
// copy: ONE `cp.async.bulk.tensor.3d` (tile mode, tensor map over {vector dim, reduce dim, outer batch dim}: any
// pitch, out-of-range rows / columns zero-filled and never read), or, without a tensor map, `cp.async.bulk` copies
// (one when the strip covers whole contiguous rows, else one per row).  The other warps are consumers: wait on
// `full`, fold their rows out of shared memory (one LDS.128 per row: a warp reads consecutive chunks, conflict
// free), and one lane per warp arrives on `empty`.  No CTA-wide barrier in the loop (the first version, with a
// __syncthreads per stage and warp 0 issuing, stalled 65 % of its issue slots on that barrier).  The ring runs ACROSS
// the items of a CTA, so the copy engine keeps `stages` x RT rows in flight per CTA through the item epilogues —
// 3-4x the bytes the LDG walker can hold in registers, which is what bounds that one (48 KB per SM).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor_g2s_3d(void *dst_smem, const void *tmap, int c0, int c1, int c2, u64 *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

template <class Tin, class Op, class OutT>
__device__ __forceinline__ void reduce_outer_tma_body(const RedParams &p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  typedef typename Widen<Tin>::type T;
  typedef typename Op::acc_t acc_t;
  enum { V = 16 / (int)sizeof(Tin) };
  extern __shared__ __align__(128) unsigned char s_dyn[];
  u64 *full = (u64 *)s_dyn;                  // first 128 bytes: full[0..7], empty[0..7]
  u64 *empty = full + 8;
  const int stages = p.splits;               // ring depth (<= 8)
  const int RT = p.tma_rt;                   // rows per stage
  const int NC = (int)blockDim.x - 32;       // consumer threads; the last warp is the producer
  const int TX = p.tx, TY = NC / TX;
  const u32 rowb = (u32)TX * 16u;            // bytes of one full strip row
  const u32 stage_bytes = (u32)RT * rowb;
  unsigned char *ring = s_dyn + 128;
  acc_t *s_part = (acc_t *)(ring + (size_t)stages * stage_bytes);   // [TY - 1][TX * V]
  const int tid = threadIdx.x, lane = tid & 31;
  const int nb = p.nb;
  const i64 C = p.bsz[nb - 1];
  const i64 tile = (i64)TX * V;
  const i64 ctiles = (C + tile - 1) / tile;
  const i64 items = (p.B / C) * ctiles;
  const i64 R = p.R;
  const i64 nch = (R + RT - 1) / RT;
  const i64 n_items = items > (i64)blockIdx.x ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const bool tensor = p.tma_mode == 1;
  // row pitch inside a stage: the tensor copy always lands a full TX-chunk box, the bulk copies pack the strip's own width

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (u32)(NC / 32)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= NC) {
    // ---------------- producer warp ----------------
    const i64 pitch = p.leaf[0].rs[0] * (i64)sizeof(Tin);   // bytes between consecutive reduce rows
    int s = 0;
    u32 round = 0;
    for (i64 n = 0; n < n_items; ++n) {
      const i64 w = (i64)blockIdx.x + n * gridDim.x;
      const i64 bo = w / ctiles, ct = w - bo * ctiles;
      const i64 cols = (C - ct * tile) < tile ? (C - ct * tile) : tile;
      const u32 wb = (u32)(cols * (i64)sizeof(Tin));
      const char *src = nullptr;
      if (!tensor) {
        i64 bidx[KMAXD];
        decomp(bo, nb - 1, p.bsz, bidx);
        i64 off = ct * tile;
#pragma unroll
        for (int d = 0; d < KMAXD - 1; ++d) if (d < nb - 1) off += bidx[d] * p.leaf[0].bs[d];
        src = (const char *)p.leaf[0].ptr + off * (i64)sizeof(Tin);
      }
      for (i64 ch = 0; ch < nch; ++ch) {
        if (round > 0) {
          if (lane == 0) mbar_wait(&empty[s], (round - 1u) & 1u);
          __syncwarp();
        }
        unsigned char *dst = ring + (size_t)s * stage_bytes;
        const i64 r0 = ch * RT;
        if (tensor) {
          if (lane == 0) {
            mbar_expect_tx(&full[s], stage_bytes);
            // coordinates in 8-byte elements along the vector dim, rows, outer batch index
            tensor_g2s_3d(dst, p.tmap, (int)(ct * TX * 2), (int)r0, (int)bo, &full[s]);
          }
        } else {
          const int rows = (int)((R - r0) < RT ? (R - r0) : RT);
          if (lane == 0) mbar_expect_tx(&full[s], (u32)rows * wb);
          __syncwarp();
          const char *sp = src + r0 * pitch;
          if (pitch == (i64)wb) {
            if (lane == 0) bulk_g2s(dst, sp, (u32)rows * wb, &full[s]);
          } else {
            for (int r = lane; r < rows; r += 32) bulk_g2s(dst + (size_t)r * wb, sp + (i64)r * pitch, wb, &full[s]);
          }
        }
        if (++s == stages) { s = 0; ++round; }
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  const int tx = tid % TX, ty = tid / TX;
  int s = 0;
  u32 phase = 0;
  for (i64 n = 0; n < n_items; ++n) {
    const i64 w = (i64)blockIdx.x + n * gridDim.x;
    const i64 bo = w / ctiles, ct = w - bo * ctiles;
    const i64 c0 = ct * tile + (i64)tx * V;      // first column of this thread (C is a multiple of V: host rule)
    const bool active = c0 < C && ty < TY;
    const i64 cols = (C - ct * tile) < tile ? (C - ct * tile) : tile;
    const u32 spitch = tensor ? rowb : (u32)(cols * (i64)sizeof(Tin));
    acc_t acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = Op::init();
    // flat index of (batch row, reduce r) = (sum_d bidx[d]*bflat[d]) * R + r; column c adds c*bflat[nb-1]*R
    i64 rowflat = p.idx_base + c0 * p.bflat[nb - 1] * R;
    if (Op::HAS_INDEX) {
      i64 bidx[KMAXD];
      decomp(bo, nb - 1, p.bsz, bidx);
#pragma unroll
      for (int d = 0; d < KMAXD - 1; ++d) if (d < nb - 1) rowflat += bidx[d] * p.bflat[d] * R;
    }
    const i64 colflat = p.bflat[nb - 1] * R;
    for (i64 ch = 0; ch < nch; ++ch) {
      mbar_wait(&full[s], phase);
      const i64 r0 = ch * RT;
      const int rows = (int)((R - r0) < RT ? (R - r0) : RT);
      if (active) {
        const unsigned char *sbase = ring + (size_t)s * stage_bytes + (size_t)tx * 16;
        int r = ty;
        // four rows per trip: the LDS.128s go out together (a warp reads consecutive chunks of one row)
        for (; r + 3 * TY < rows; r += 4 * TY) {
          union { uint4 q; Vec<Tin, V> x; } u[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) u[i].q = *(const uint4 *)(sbase + (size_t)(r + i * TY) * spitch);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int v = 0; v < V; ++v) Op::step(acc[v], cvt<T>(u[i].x.v[v]), rowflat + (i64)v * colflat + r0 + r + i * TY);
          }
        }
        for (; r < rows; r += TY) {
          union { uint4 q; Vec<Tin, V> x; } u;
          u.q = *(const uint4 *)(sbase + (size_t)r * spitch);
#pragma unroll
          for (int v = 0; v < V; ++v) Op::step(acc[v], cvt<T>(u.x.v[v]), rowflat + (i64)v * colflat + r0 + r);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);   // this warp is done with stage s
      if (++s == stages) { s = 0; phase ^= 1u; }
    }
    if (n + 1 == n_items) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // TY partials per column meet in shared memory (fixed order: ty = 0, 1, ...); consumer-only named barrier
    if (TY > 1) {
      asm volatile("bar.sync 1, %0;" ::"r"(NC) : "memory");   // the previous item's readers of s_part are done
      if (active && ty > 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) s_part[((size_t)(ty - 1) * TX + tx) * V + v] = acc[v];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(NC) : "memory");
      if (active && ty == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          acc_t a = acc[v];
          for (int y = 1; y < TY; ++y) Op::merge(a, s_part[((size_t)(y - 1) * TX + tx) * V + v]);
          acc[v] = a;
        }
      }
    }
    if (active && ty == 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) store_result<Op, OutT>(p, bo * C + c0 + v, acc[v]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// E1: elementwise — out(idx) = expr(idx).  Persistent grid-stride over V-wide vectors of the innermost
// dim (LDG.128/256 per leaf, each distinct leaf loaded once, STG.128/256), U vectors per thread in
// flight; V == 1 takes any strides.  Replaces the one-vector-per-thread generic kernels of
// executors/kernel.h:41-223.
// ------------------------------------------------------------------------------------------------
template <class E, class OutT, int V>
__device__ __forceinline__ void ew_store(const EwParams &p, char *obase, i64 oinner, i64 j, const typename E::template Regs<V> &r) {
  Vec<OutT, V> o;
  if constexpr (E::PAIR && V % 2 == 0) {
#pragma unroll
    for (int v = 0; v < V; v += 2) {
      const f2 t = E::template eval2<V>(r, v, p.c);
      o.v[v] = cvt<OutT>(t.v.x);
      o.v[v + 1] = cvt<OutT>(t.v.y);
    }
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(E::template eval<V>(r, v, p.c));
  }
  if (V == 1) ((OutT *)obase)[j * oinner] = o.v[0];
  else StBytes<(int)sizeof(OutT) * V>::st((OutT *)obase + j, &o);
}

template <class E, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void ew_body_impl(const EwParams &p) {
  const int nd = p.nd;
  const i64 L = p.sz[nd - 1], Lv = L / V, tail = L - Lv * V, O = p.N / L;
  const i64 Q = O * Lv;
  const i64 nthr = (i64)gridDim.x * blockDim.x;
  const i64 t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 inner[E::NL];
#pragma unroll
  for (int k = 0; k < E::NL; ++k) inner[k] = p.leaf[k].bs[nd - 1];
  const i64 oinner = p.out.bs[nd - 1];

  if (nd == 1) {
    const char *base[E::NL];
#pragma unroll
    for (int k = 0; k < E::NL; ++k) base[k] = (const char *)p.leaf[k].ptr;
    // a CTA takes batches of blockDim.x * U consecutive vectors (its U loads per leaf are neighbours in memory: the
    // streams a CTA has open stay within one DRAM page per operand), batches are dealt round-robin to the grid
    const i64 bsz = (i64)blockDim.x * U;
    const i64 nbatch = Q / bsz;
    for (i64 bi = blockIdx.x; bi < nbatch; bi += gridDim.x) {
      const i64 q = bi * bsz + threadIdx.x;
      typename E::template Regs<V> r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) E::template loadv<V, UNIT>(r[u], base, inner, (q + (i64)u * blockDim.x) * V);
#pragma unroll
      for (int u = 0; u < U; ++u) ew_store<E, OutT, V>(p, (char *)p.out.ptr, oinner, (q + (i64)u * blockDim.x) * V, r[u]);
    }
    for (i64 q = nbatch * bsz + t0; q < Q; q += nthr) {
      typename E::template Regs<V> r;
      E::template loadv<V, UNIT>(r, base, inner, q * V);
      ew_store<E, OutT, V>(p, (char *)p.out.ptr, oinner, q * V, r);
    }
  } else {
    for (i64 q = t0; q < Q; q += nthr) {
      const i64 o = q / Lv, jv = q - o * Lv;
      i64 oidx[KMAXD];
      decomp(o, nd - 1, p.sz, oidx);
      const char *base[E::NL];
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD - 1; ++d) if (d < nd - 1) off += oidx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      }
      i64 ooff = 0;
#pragma unroll
      for (int d = 0; d < KMAXD - 1; ++d) if (d < nd - 1) ooff += oidx[d] * p.out.bs[d];
      typename E::template Regs<V> r;
      E::template loadv<V, UNIT>(r, base, inner, jv * V);
      ew_store<E, OutT, V>(p, (char *)p.out.ptr + ooff * (i64)sizeof(OutT), oinner, jv * V, r);
    }
  }
  if (V > 1 && tail > 0) {
    const i64 nt = O * tail;
    for (i64 t = t0; t < nt; t += nthr) {
      const i64 o = t / tail, j = Lv * V + (t - o * tail);
      i64 oidx[KMAXD];
      decomp(o, nd - 1, p.sz, oidx);
      const char *base[E::NL];
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD - 1; ++d) if (d < nd - 1) off += oidx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      }
      i64 ooff = 0;
#pragma unroll
      for (int d = 0; d < KMAXD - 1; ++d) if (d < nd - 1) ooff += oidx[d] * p.out.bs[d];
      typename E::template Regs<1> r;
      E::template loadv<1, UNIT>(r, base, inner, j);
      ew_store<E, OutT, 1>(p, (char *)p.out.ptr + ooff * (i64)sizeof(OutT), oinner, j, r);
    }
  }
}


// Programmatic dependent launch: let the next kernel on the stream start its launch while this one runs, and do not
// touch memory before the previous kernel has completed and flushed (no-ops without the launch attribute).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// launch-time dispatch on the unit-stride flag
template <class E, class Op, class OutT, int V, int U, int TEAM>
__device__ __forceinline__ void reduce_inner_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) reduce_inner_body_impl<E, Op, OutT, V, U, TEAM, true>(p);
  else reduce_inner_body_impl<E, Op, OutT, V, U, TEAM, false>(p);
}

template <class E, class Op, class OutT, int V, int U>
__device__ __forceinline__ void reduce_outer_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) reduce_outer_body_impl<E, Op, OutT, V, U, true>(p);
  else reduce_outer_body_impl<E, Op, OutT, V, U, false>(p);
}

template <class E, class OutT, int V, int U>
__device__ __forceinline__ void var_inner_smem_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) var_inner_smem_body_impl<E, OutT, V, U, true>(p);
  else var_inner_smem_body_impl<E, OutT, V, U, false>(p);
}

template <class E, class OutT, int V, int U>
__device__ __forceinline__ void ew_body(const EwParams &p) {
  pdl_prologue();
  if (p.all_unit) ew_body_impl<E, OutT, V, U, true>(p);
  else ew_body_impl<E, OutT, V, U, false>(p);
}

template <class E, class OutT, int V, int IPT>
__device__ __forceinline__ void var_inner_reg_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) var_inner_reg_body_impl<E, OutT, V, IPT, true>(p);
  else var_inner_reg_body_impl<E, OutT, V, IPT, false>(p);
}

template <class E, class OutT, int V, int IPT>
__device__ __forceinline__ void var_group_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) var_group_body_impl<E, OutT, V, IPT, true>(p);
  else var_group_body_impl<E, OutT, V, IPT, false>(p);
}

template <class E, class OutT, int V, int IPT>
__device__ __forceinline__ void softmax_group_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) softmax_group_body_impl<E, OutT, V, IPT, true>(p);
  else softmax_group_body_impl<E, OutT, V, IPT, false>(p);
}

template <class E, class OutT, int V, int IPT>
__device__ __forceinline__ void softmax_reg_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) softmax_reg_body_impl<E, OutT, V, IPT, true>(p);
  else softmax_reg_body_impl<E, OutT, V, IPT, false>(p);
}

// ------------------------------------------------------------------------------------------------
// Transposing elementwise kernel: out(idx) = expr(idx) when some leaves are unit-stride along a dim Y that is NOT
// the output's unit-stride dim X — permuted copies `(y = x.Permute(...)).run()` and expressions that mix row- and
// column-walking operands (`a + permute(b)`).  The reference's generic kernel (executors/kernel.h:41-223) walks the
// output index and reads such a leaf one element per 32-byte sector; its own answer is a separate tiled transpose
// for the 2-D case only (kernels/transpose.cuh:28).  Here every CTA owns a TX x TY tile of the (X, Y) plane of one
// outer index:
//   phase 1  the Y-walking leaves are read along Y (16-byte chunks, a warp covers 4 rows x 128 B) and written
//            element-wise into shared memory in [y][x] order;
//   phase 2  threads walk X: a 16-byte chunk of every staged leaf comes back with one LDS.128, the X-walking leaves
//            are loaded straight from global memory, the expression is evaluated and V results leave in one store
//            (a warp covers 2 rows x 256 B).
// Shared-memory layout per staged leaf: TY rows of 256 bytes (TX = 16 chunks of V = 16/EB elements), chunk index
// XOR-swizzled with the row's own chunk number (y / V) & 15, which makes the phase-1 element stores and the phase-2
// 16-byte loads bank-conflict free for EB = 2, 4, 8.  HBM traffic = algorithmic bytes (every sector fully used).
// ------------------------------------------------------------------------------------------------
template <int EB> struct TrTile {
  enum { V = 16 / EB, TX = 16 * V, TY = 16 * V, BYTES = TY * 256 };
};

__device__ __forceinline__ void lds16(void *d, const void *smem_generic) {
  const u32 a = (u32)__cvta_generic_to_shared(smem_generic);
  u32 *o = (u32 *)d;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(a));
}
// leaf read from shared memory (one 16-byte chunk = V elements)
template <class T, int V, bool CHUNK = (sizeof(T) * V == 16)> struct LdTile {
  static __device__ __forceinline__ void go(Vec<T, V> &r, const char *s) { lds16(&r, s); }
};
template <class T, int V> struct LdTile<T, V, false> {  // a leaf of another element size is never staged (host rule)
  static __device__ __forceinline__ void go(Vec<T, V> &, const char *) {}
};
template <class T, int V> __device__ __forceinline__ void ldtile(Vec<T, V> &r, const char *s) { LdTile<T, V>::go(r, s); }
// leaf that walks X itself: V elements at p, p+inner, ... (inner is 0 or 1 in practice); `vec` = one aligned vector
// load is legal, `n` = how many of the V elements exist
template <class T, int V> __device__ __forceinline__ void ldrow(Vec<T, V> &r, const void *base, i64 inner, bool vec, int n) {
  const T *p = (const T *)base;
  if (inner == 0) {
    ldsplat<T, V>(r, p);
  } else if (vec && n == V && inner == 1) {
    ldv<T, V>(r, p);
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (v < n) { Vec<T, 1> s; LdBytes<(int)sizeof(T)>::ld(&s, p + (i64)v * inner); r.v[v] = s.v[0]; }
      else r.v[v] = r.v[0];
    }
  }
}

template <class E, class OutT, int EB>
__device__ __forceinline__ void ew_tr_body(const EwParams &p) {
  typedef TrTile<EB> TT;
  constexpr int V = TT::V, TX = TT::TX, TY = TT::TY;
  extern __shared__ __align__(16) unsigned char tr_smem[];
  pdl_prologue();
  const int nd = p.nd, ydim = p.tr_ydim, xdim = nd - 1;
  const i64 SX = p.sz[xdim], SY = p.sz[ydim];
  // CTA -> (outer index, x tile, y tile), y tiles fastest: neighbouring CTAs read neighbouring 256-byte runs.  The grid
  // fits 31 bits (host rule), so every quotient here is a 32-bit division — the 64-bit ones cost more instructions
  // than moving the tile does.
  unsigned b = blockIdx.x;
  const unsigned nty = (unsigned)((SY + TY - 1) / TY), ntx = (unsigned)((SX + TX - 1) / TX);
  const unsigned ty = b % nty; b /= nty;
  const unsigned tx = b % ntx; b /= ntx;
  i64 lofs[E::NL], oofs = 0;
#pragma unroll
  for (int k = 0; k < E::NL; ++k) lofs[k] = 0;
#pragma unroll
  for (int d = KMAXD - 1; d >= 0; --d) {
    if (d < nd - 1 && d != ydim) {
      const unsigned szd = (unsigned)p.sz[d];
      const unsigned q = b / szd, i = b - q * szd;
      b = q;
#pragma unroll
      for (int k = 0; k < E::NL; ++k) lofs[k] += (i64)i * p.leaf[k].bs[d];
      oofs += (i64)i * p.out.bs[d];
    }
  }
  const i64 x0 = (i64)tx * TX, y0 = (i64)ty * TY;
  const int remx = (int)((SX - x0) < (i64)TX ? (SX - x0) : (i64)TX);   // valid extent of this tile
  const int remy = (int)((SY - y0) < (i64)TY ? (SY - y0) : (i64)TY);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned ymask = p.tr_ymask;

  // ---- phase 1: stage the Y-walking leaves ----
  {
    int slot = 0;
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      if (!((ymask >> k) & 1u)) continue;
      unsigned char *tile = tr_smem + (size_t)slot * TT::BYTES;
      ++slot;
      const i64 sx = p.leaf[k].bs[xdim];
      const char *g = (const char *)p.leaf[k].ptr + (lofs[k] + y0 + x0 * sx) * EB;   // stride along Y is 1
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int unit = warp + 8 * i;
        const int x = (unit >> 1) * 4 + (lane >> 3);
        const int cy = (unit & 1) * 8 + (lane & 7);
        if (x >= remx || cy * V >= remy) continue;
        const char *src = g + ((i64)x * sx + cy * V) * EB;
        unsigned char e[16];
        const int n = (remy - cy * V) < V ? (remy - cy * V) : V;
        if (p.tr_yvec && n == V) {
          LdBytes<16>::ld(e, src);
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) if (j < n) LdBytes<EB>::ld(e + j * EB, src + j * EB);
        }
        // element (x, cy*V + j) -> row cy*V + j, chunk (x / V) ^ cy, slot x % V
        const int col = (((x / V) ^ cy) * 16) + (x % V) * EB;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          unsigned char *dst = tile + (size_t)(cy * V + j) * 256 + col;
          if (EB == 2) *(unsigned short *)dst = *(const unsigned short *)(e + j * EB);
          else if (EB == 4) *(u32 *)dst = *(const u32 *)(e + j * EB);
          else *(unsigned long long *)dst = *(const unsigned long long *)(e + j * EB);
        }
      }
    }
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- phase 2: walk X, evaluate, store ----
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int unit = warp + 8 * i;
    const int yy = unit * 2 + (lane >> 4);
    const int cx = lane & 15;
    if (yy >= remy || cx * V >= remx) continue;
    const i64 gy = y0 + yy, gx = x0 + (i64)cx * V;
    const int n = (remx - cx * V) < V ? (remx - cx * V) : V;
    const char *sp[E::NL];
    const char *gp[E::NL];
    i64 ginner[E::NL];
    int slot = 0;
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      if ((ymask >> k) & 1u) {
        sp[k] = (const char *)tr_smem + (size_t)slot * TT::BYTES + (size_t)yy * 256 + ((cx ^ ((yy / V) & 15)) * 16);
        gp[k] = nullptr;
        ginner[k] = 0;
        ++slot;
      } else {
        sp[k] = nullptr;
        ginner[k] = p.leaf[k].bs[xdim];
        gp[k] = (const char *)p.leaf[k].ptr + (lofs[k] + gy * p.leaf[k].bs[ydim] + gx * ginner[k]) * E::leaf_bytes(k);
      }
    }
    typename E::template Regs<V> r;
    E::template loadmix<V>(r, gp, ginner, sp, ymask, p.tr_xvec != 0, n);
    Vec<OutT, V> o;
#pragma unroll
    for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(E::template eval<V>(r, v, p.c));
    OutT *op = (OutT *)p.out.ptr + oofs + gy * p.out.bs[ydim] + gx;
    if (p.tr_ovec && n == V) {
      StBytes<(int)sizeof(OutT) * V>::st(op, &o);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) if (v < n) op[v] = o.v[v];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K9: scan_inner — inclusive prefix sum along the innermost dim (reference: cumsum_impl -> cub::DeviceScan::InclusiveSum,
// transforms/cub.h:375-408,2367-2395; one CUB launch PER ROW there, plus a second pass inside CUB's own decoupled
// look-back).  One launch here, every element read once and written once.
//
// A tile is NT threads x U chunks x V elements; thread t owns V adjacent elements of every chunk (vector loads stay
// coalesced), scans them in registers, the warp scans the thread totals with shuffles, the eight warp totals of the U
// chunks meet in shared memory.  All additions happen in a FIXED order, so results are run-to-run deterministic.
//   mode ROWS  (p.splits == 1): a CTA owns whole rows and walks their tiles with a carry in a register.
//   mode TILES (p.splits  > 1): few long rows -> the tiles of all rows are dealt round-robin to a grid whose CTAs are all
//     resident (cooperative launch), so a tile only ever waits for tiles that are running or done.  Every tile publishes
//     its total; the last tile of a GROUP of 32 tiles publishes the group's total; the last tile of a SUPERGROUP of 32
//     groups publishes the running total at the start of the next supergroup (the only chained quantity: one link per
//     1024 tiles, far slower than the stream).  A tile's carry = that running total + the totals of the groups before
//     its own in the supergroup + the totals of the tiles before it in its group: at most 1 + 31 + 31 values, fetched by
//     warp 0 with all loads in flight at once and summed by fixed-shape shuffle trees — one L2 round trip when the
//     neighbours are done, no atomics, and no dependence on which neighbour happened to finish first (CUB's look-back adds
//     whatever it finds, so its float sums vary from run to run; these do not).
//   Both modes keep the NEXT tile's loads in flight while the current tile goes through its barriers and its carry
//   exchange (the loads are issued into the registers the evaluation has just freed).
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_NT = 256;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// A published total: tag = (launch epoch << 2) | 1 in the high half of a 64-bit word.  4-byte values travel IN that word
// (one relaxed load both tests and fetches, what CUB's tile status words do); 8-byte values sit beside it in a 16-byte
// slot moved with one 128-bit access (never torn inside an aligned 16-byte sector; CUB's ScanTileState makes the same
// assumption for 8-byte types).  The epoch lives in device memory and is bumped by the last CTA out, so nothing is
// cleared between launches and a CUDA-graph replay gets a fresh epoch as well.
template <class T, int BYTES = (int)sizeof(T)> struct ScanSlot;
template <class T> struct ScanSlot<T, 4> {
  enum { STRIDE = 8 };
  struct Word { unsigned long long w; };
  static __device__ __forceinline__ void publish(void *slots, i64 i, T v, u32 tag) {
    union { T t; u32 w; } u;
    u.t = v;
    st_relaxed_u64((unsigned long long *)slots + i, ((unsigned long long)tag << 32) | u.w);
  }
  static __device__ __forceinline__ Word peek(const void *slots, i64 i) { Word r; r.w = ld_relaxed_u64((const unsigned long long *)slots + i); return r; }
  static __device__ __forceinline__ bool ready(const Word &r, u32 tag) { return (u32)(r.w >> 32) == tag; }
  static __device__ __forceinline__ T value(const Word &r) { union { T t; u32 w; } u; u.w = (u32)r.w; return u.t; }
  static __device__ __forceinline__ T wait(const void *slots, i64 i, u32 tag) {
    Word r = peek(slots, i);
    while (!ready(r, tag)) { __nanosleep(20); r = peek(slots, i); }
    return value(r);
  }
};
template <class T> struct ScanSlot<T, 8> {
  enum { STRIDE = 16 };
  struct Word { unsigned long long f, w; };
  static __device__ __forceinline__ void publish(void *slots, i64 i, T v, u32 tag) {
    union { T t; unsigned long long w; } u;
    u.t = v;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"((unsigned long long *)slots + 2 * i), "l"((unsigned long long)tag << 32), "l"(u.w) : "memory");
  }
  static __device__ __forceinline__ Word peek(const void *slots, i64 i) {
    Word r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.f), "=l"(r.w) : "l"((const unsigned long long *)slots + 2 * i) : "memory");
    return r;
  }
  static __device__ __forceinline__ bool ready(const Word &r, u32 tag) { return (u32)(r.f >> 32) == tag; }
  static __device__ __forceinline__ T value(const Word &r) { union { T t; unsigned long long w; } u; u.w = r.w; return u.t; }
  static __device__ __forceinline__ T wait(const void *slots, i64 i, u32 tag) {
    Word r = peek(slots, i);
    while (!ready(r, tag)) { __nanosleep(20); r = peek(slots, i); }
    return value(r);
  }
};

template <class T> __device__ __forceinline__ T shfl_up_t(T v, int d) {
  enum { W = sizeof(T) / 4 };
  union { T t; u32 w[W]; } a, b;
  a.t = v;
#pragma unroll
  for (int i = 0; i < W; ++i) b.w[i] = __shfl_up_sync(0xffffffffu, a.w[i], d);
  return b.t;
}
template <class T> __device__ __forceinline__ T shfl_idx_t(T v, int src) {
  enum { W = sizeof(T) / 4 };
  union { T t; u32 w[W]; } a, b;
  a.t = v;
#pragma unroll
  for (int i = 0; i < W; ++i) b.w[i] = __shfl_sync(0xffffffffu, a.w[i], src);
  return b.t;
}
template <class T> __device__ __forceinline__ T scan_zero() { return cvt<T>(0.0f); }
template <> __device__ __forceinline__ u32 scan_zero<u32>() { return 0u; }
// fixed-shape butterfly sum over a warp: every lane ends with the same bits (a + b == b + a exactly)
template <class T> __device__ __forceinline__ T warp_tree_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = v + shfl_xor_t(v, m);
  return v;
}

// One in-flight read of a published slot: issue() sends the load, get() spins only if the value was not there yet.  The
// pipelined kernels issue the reads of an iteration's exchange jobs before the iteration's own work and collect them after
// it, so the L2 round trip hides behind the work.
template <class T> struct SlotRead {
  typename ScanSlot<T>::Word w;
  const void *slots;
  i64 i;
  bool on;
  __device__ __forceinline__ void issue(const void *s, i64 idx, bool enable) {
    slots = s; i = idx; on = enable;
    if (on) w = ScanSlot<T>::peek(slots, i);
  }
  __device__ __forceinline__ T get(u32 tag) {
    if (!on) return scan_zero<T>();
    while (!ScanSlot<T>::ready(w, tag)) { __nanosleep(20); w = ScanSlot<T>::peek(slots, i); }
    return ScanSlot<T>::value(w);
  }
};

// The three-level exchange of hier_carry, split into jobs that run in DIFFERENT iterations of a pipelined tile loop
// (depth D >= 4: phase 1 of a tile — load, local work, publish its total — runs D - 1 iterations before its phase 2 — carry,
// store), by warp 0 of the CTA that owns the tile:
//   close   iteration + 1  a tile that ends its group of 32 sums the group's tile totals and publishes the group total
//   chain   iteration + 2  a tile that ends its supergroup of 32 groups publishes the running total at the next one
//   carry   iteration + D - 1  running total at the supergroup start + group totals before + tile totals before
// Everything a job reads was published at least one iteration earlier by CTAs running the same loop, so the reads find
// their data instead of spinning on it (they still spin if a CTA lags), and the sums keep their fixed shape.
template <class T> struct TileExchange {
  char *agg, *gagg, *sagg;   // slots of row 0; row r sits tpr / gpr / spr slots further
  i64 tpr, gpr, spr;
  u32 tag;
  int lane;
  SlotRead<T> ca, cb, cc, ga, sb, sc;   // carry: tiles / groups / supergroup start; close: tiles; chain: groups / supergroup start
  i64 close_ct, chain_ct, close_row, chain_row;
  __device__ __forceinline__ void init(void *a, void *g, void *sp, i64 tiles_per_row, u32 tg, int ln) {
    agg = (char *)a; gagg = (char *)g; sagg = (char *)sp;
    tpr = tiles_per_row; gpr = (tpr + 31) >> 5; spr = (tpr + 1023) >> 10;
    tag = tg; lane = ln;
    close_ct = chain_ct = -1; close_row = chain_row = 0;
  }
  __device__ __forceinline__ char *A(i64 row) const { return agg + (size_t)(row * tpr) * ScanSlot<T>::STRIDE; }
  __device__ __forceinline__ char *G(i64 row) const { return gagg + (size_t)(row * gpr) * ScanSlot<T>::STRIDE; }
  __device__ __forceinline__ char *S(i64 row) const { return sagg + (size_t)(row * spr) * ScanSlot<T>::STRIDE; }
  __device__ __forceinline__ bool closes_group(i64 ct) const { return (ct & 31) == 31 || ct == tpr - 1; }
  __device__ __forceinline__ bool closes_super(i64 ct) const { return (ct & 1023) == 1023 && ct != tpr - 1; }
  // a tile publishes its own total (phase 1)
  __device__ __forceinline__ void publish_tile(i64 row, i64 ct, T total) const { ScanSlot<T>::publish(A(row), ct, total, tag); }
  // top of an iteration: (row, tile within the row) of the close / chain / carry job, tile < 0 = none
  __device__ __forceinline__ void issue(i64 row_close, i64 ct_close, i64 row_chain, i64 ct_chain, i64 row_carry, i64 ct_carry) {
    close_ct = (ct_close >= 0 && closes_group(ct_close)) ? ct_close : -1;
    chain_ct = (ct_chain >= 0 && closes_super(ct_chain)) ? ct_chain : -1;
    close_row = row_close;
    chain_row = row_chain;
    {
      const i64 first = (close_ct >> 5) << 5;
      ga.issue(A(row_close), first + lane, close_ct >= 0 && first + lane < close_ct);
    }
    {
      const i64 sg = chain_ct >> 10;
      sb.issue(G(row_chain), (sg << 5) + lane, chain_ct >= 0);
      sc.issue(S(row_chain), sg, chain_ct >= 0 && sg > 0 && lane == 0);
    }
    {
      const i64 g = ct_carry >> 5, first = g << 5, sg = ct_carry >> 10, gfirst = sg << 5;
      ca.issue(A(row_carry), first + lane, ct_carry >= 0 && first + lane < ct_carry);
      cb.issue(G(row_carry), gfirst + lane, ct_carry >= 0 && gfirst + lane < g);
      cc.issue(S(row_carry), sg, ct_carry >= 0 && sg > 0 && lane == 0);
    }
  }
  // bottom of the iteration: finish the jobs; `close_total` = the closing tile's own total; returns the carry (all lanes)
  __device__ __forceinline__ T finish(T close_total) {
    if (close_ct >= 0) {
      const T gt = warp_tree_sum(ga.get(tag)) + close_total;
      if (lane == 0) ScanSlot<T>::publish(G(close_row), close_ct >> 5, gt, tag);
    }
    if (chain_ct >= 0) {
      const T st = warp_tree_sum(sb.get(tag));
      const T c = shfl_idx_t(sc.get(tag), 0);
      if (lane == 0) ScanSlot<T>::publish(S(chain_row), (chain_ct >> 10) + 1, c + st, tag);
    }
    const T sa = warp_tree_sum(ca.get(tag)), sbb = warp_tree_sum(cb.get(tag));
    const T c = shfl_idx_t(cc.get(tag), 0);
    return (c + sbb) + sa;
  }
};

// Carry of tile `ct` (of a row of `tpr` tiles) from the published totals, by ONE warp (all 32 lanes call it; every lane
// returns the carry).  The caller has already published this tile's own total in agg[ct].  Three levels, all loads in
// flight at once: running total at the start of the tile's supergroup (1024 tiles; chained, published by the last tile
// of the previous supergroup) + totals of the groups (32 tiles) before the tile's own in that supergroup + totals of the
// tiles before it in its group.  The last tile of a group publishes the group total, the last tile of a supergroup the
// next running total.  Sums are fixed-shape shuffle trees: the carry does not depend on timing.
// (A flat CUB-style look-back would walk back through every tile of the same WAVE here — the persistent grid starts a
// wave of tiles together, none of them has a running total yet — one L2 round trip per 32 tiles: measured 5 us per tile.)
template <class T>
__device__ __forceinline__ T hier_carry(void *agg, void *gagg, void *sagg, i64 ct, i64 tpr, T total, u32 tag, int lane) {
  const i64 g = ct >> 5, first = g << 5, sg = ct >> 10, gfirst = sg << 5;
  const int n2 = (int)(ct - first), n1 = (int)(g - gfirst);
  // the group's total goes out as soon as the group's own tiles are in — BEFORE waiting for anything of an earlier
  // group: the first tiles of the next group wait for it, and waiting for earlier groups first would chain the groups
  T a = scan_zero<T>();
  if (lane < n2) a = ScanSlot<T>::wait(agg, first + lane, tag);
  const T sa = warp_tree_sum(a);
  const bool last_tile = ct == tpr - 1;
  const bool closes_group = (ct & 31) == 31 || last_tile;
  const T gt = sa + total;                         // this group's total (meaningful when the tile closes its group)
  if (closes_group && lane == 0) ScanSlot<T>::publish(gagg, g, gt, tag);
  T b = scan_zero<T>(), c = scan_zero<T>();
  if (lane < n1) b = ScanSlot<T>::wait(gagg, gfirst + lane, tag);
  if (lane == 0 && sg > 0) c = ScanSlot<T>::wait(sagg, sg, tag);
  const T sb = warp_tree_sum(b);
  c = shfl_idx_t(c, 0);
  if (closes_group && ((g & 31) == 31) && !last_tile && lane == 0) ScanSlot<T>::publish(sagg, sg + 1, c + (sb + gt), tag);
  return (c + sb) + sa;
}

// warp-per-row flavour for short rows: no shared memory, no barrier; a warp walks its rows in steps of 32 x U x V
template <class E, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void scan_warp_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  const int lane = threadIdx.x & 31;
  const i64 L = p.rsz[0];
  const i64 STEP = (i64)32 * V * U;
  const i64 w0 = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), wstep = (i64)gridDim.x * (blockDim.x >> 5);
  const i64 oinner = p.out_rs[0];
  if (U == 1 && p.scan_group > 0) {
    // rows of at most G vectors (G = p.scan_group, a power of two <= 32): G lanes per row, 32 / G rows per warp, one
    // segmented shuffle scan, no loop — a 64-element fp32 row would otherwise leave half of a warp idle and an
    // 8-element one 30 lanes of 32
    const int G = p.scan_group, lr = lane & (G - 1);
    const i64 rpw = 32 / G;
    for (i64 b0 = w0 * rpw; b0 < p.B; b0 += wstep * rpw) {   // warp-uniform trip count: the shuffles stay full-mask
      const i64 b = b0 + lane / G;
      const bool live = b < p.B;
      const char *base[E::NL];
      i64 inner[E::NL];
      i64 oo = 0;
      {
        i64 bidx[KMAXD];
        decomp(live ? b : 0, p.nb, p.bsz, bidx);
#pragma unroll
        for (int k = 0; k < E::NL; ++k) {
          i64 off = 0;
#pragma unroll
          for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
          base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
          inner[k] = p.leaf[k].rs[0];
        }
#pragma unroll
        for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
      }
      OutT *orow = (OutT *)p.out.ptr + oo;
      const i64 j = (i64)lr * V;
      T x[V];
      const bool full = V == 1 ? (j < L) : (j + V <= L);
      if (live && full) {
        typename E::template Regs<V> r;
        E::template loadv<V, UNIT>(r, base, inner, j);
#pragma unroll
        for (int v = 0; v < V; ++v) x[v] = E::template eval<V>(r, v, p.c);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          x[v] = scan_zero<T>();
          if (live && V > 1 && j + v < L) {
            typename E::template Regs<1> r1;
            E::template loadv<1, false>(r1, base, inner, j + v);
            x[v] = E::template eval<1>(r1, 0, p.c);
          }
        }
      }
#pragma unroll
      for (int v = 1; v < V; ++v) x[v] = x[v - 1] + x[v];
      T incl = x[V - 1];
      for (int d = 1; d < G; d <<= 1) {
        const T o = shfl_up_t(incl, d);
        if (lr >= d) incl = o + incl;
      }
      const T ex = shfl_up_t(incl, 1);
      if (live) {
        Vec<OutT, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(lr == 0 ? x[v] : ex + x[v]);
        if (V > 1 && j + V <= L && oinner == 1 && p.tx) StBytes<(int)sizeof(OutT) * V>::st(orow + j, &o);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) if (j + v < L) orow[(j + v) * oinner] = o.v[v];
        }
      }
    }
    return;
  }
  for (i64 b = w0; b < p.B; b += wstep) {
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 oo = 0;
    {
      i64 bidx[KMAXD];
      decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
        inner[k] = p.leaf[k].rs[0];
      }
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    }
    OutT *orow = (OutT *)p.out.ptr + oo;
    T carry = scan_zero<T>();
    for (i64 j0 = 0; j0 < L; j0 += STEP) {
      T x[U][V];
      {
        typename E::template Regs<V> r[U];
        bool full[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = j0 + ((i64)u * 32 + lane) * V;
          full[u] = V == 1 ? (j < L) : (j + V <= L);
          if (full[u]) E::template loadv<V, UNIT>(r[u], base, inner, j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = j0 + ((i64)u * 32 + lane) * V;
          if (full[u]) {
#pragma unroll
            for (int v = 0; v < V; ++v) x[u][v] = E::template eval<V>(r[u], v, p.c);
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              x[u][v] = scan_zero<T>();
              if (V > 1 && j + v < L) {
                typename E::template Regs<1> r1;
                E::template loadv<1, false>(r1, base, inner, j + v);
                x[u][v] = E::template eval<1>(r1, 0, p.c);
              }
            }
          }
#pragma unroll
          for (int v = 1; v < V; ++v) x[u][v] = x[u][v - 1] + x[u][v];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T incl = x[u][V - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const T o = shfl_up_t(incl, d);
          if (lane >= d) incl = o + incl;
        }
        const T ex = shfl_up_t(incl, 1);
        const T pre = lane == 0 ? carry : carry + ex;
        const i64 j = j0 + ((i64)u * 32 + lane) * V;
        Vec<OutT, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(pre + x[u][v]);
        if (V > 1 && j + V <= L && oinner == 1 && p.tx) StBytes<(int)sizeof(OutT) * V>::st(orow + j, &o);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) if (j + v < L) orow[(j + v) * oinner] = o.v[v];
        }
        // chunk total = the last lane's inclusive value, broadcast
        enum { W = sizeof(T) / 4 };
        union { T t; u32 w[W]; } a, c;
        a.t = incl;
#pragma unroll
        for (int i = 0; i < W; ++i) c.w[i] = __shfl_sync(0xffffffffu, a.w[i], 31);
        carry = carry + c.t;
      }
    }
  }
}

template <class E, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void scan_inner_body_impl(const RedParams &p) {   // mode ROWS: a CTA walks whole rows
  typedef typename E::value_type T;
  constexpr int NT = SCAN_NT, NW = NT / 32;
  __shared__ T s_warp[2][U][NW];   // warp totals of every chunk, double-buffered by tile parity: one barrier per tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 L = p.rsz[0];
  const i64 TILE = (i64)NT * V * U;
  const i64 tpr = (L + TILE - 1) / TILE;          // tiles per row
  const i64 oinner = p.out_rs[0];

  const char *base[E::NL];
  i64 inner[E::NL];
#pragma unroll
  for (int k = 0; k < E::NL; ++k) inner[k] = p.leaf[k].rs[0];
  auto setup_row = [&](i64 b) -> OutT * {
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
    i64 oo = 0;
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
    }
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    return (OutT *)p.out.ptr + oo;
  };

  i64 cb = blockIdx.x, ct = 0;
  if (cb >= p.B) return;
  OutT *orow = setup_row(cb);
  typename E::template Regs<V> r[U];
  bool full[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const i64 j = ct * TILE + ((i64)u * NT + tid) * V;
    full[u] = V == 1 ? (j < L) : (j + V <= L);
    if (full[u]) E::template loadv<V, UNIT>(r[u], base, inner, j);
  }
  T carry = scan_zero<T>();
  int par = 0;
  while (true) {
    const i64 j0 = ct * TILE;
    T x[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 j = j0 + ((i64)u * NT + tid) * V;
      if (full[u]) {
#pragma unroll
        for (int v = 0; v < V; ++v) x[u][v] = E::template eval<V>(r[u], v, p.c);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          x[u][v] = scan_zero<T>();
          if (V > 1 && j + v < L) {
            typename E::template Regs<1> r1;
            E::template loadv<1, false>(r1, base, inner, j + v);
            x[u][v] = E::template eval<1>(r1, 0, p.c);
          }
        }
      }
#pragma unroll
      for (int v = 1; v < V; ++v) x[u][v] = x[u][v - 1] + x[u][v];
    }
    T wexcl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T incl = x[u][V - 1];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const T o = shfl_up_t(incl, d);
        if (lane >= d) incl = o + incl;
      }
      const T ex = shfl_up_t(incl, 1);
      wexcl[u] = lane == 0 ? scan_zero<T>() : ex;
      if (lane == 31) s_warp[par][u][warp] = incl;
    }
    __syncthreads();
    T wpre[U], ctot[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T run = scan_zero<T>();
      wpre[u] = run;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        if (w == warp) wpre[u] = run;
        run = run + s_warp[par][u][w];
      }
      ctot[u] = run;
    }
    // the NEXT tile's loads fly while this tile's stores go out
    const i64 nb = ct + 1 < tpr ? cb : cb + gridDim.x;
    const i64 nt = ct + 1 < tpr ? ct + 1 : 0;
    const bool more = nb < p.B;
    OutT *orow_next = orow;
    bool full_next[U];
    if (more) {
      if (nb != cb) orow_next = setup_row(nb);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = nt * TILE + ((i64)u * NT + tid) * V;
        full_next[u] = V == 1 ? (j < L) : (j + V <= L);
        if (full_next[u]) E::template loadv<V, UNIT>(r[u], base, inner, j);
      }
    }
    T cpre = carry;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 j = j0 + ((i64)u * NT + tid) * V;
      const T pre = (cpre + wpre[u]) + wexcl[u];
      Vec<OutT, V> o;
#pragma unroll
      for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(pre + x[u][v]);
      if (V > 1 && j + V <= L && oinner == 1 && p.tx) StBytes<(int)sizeof(OutT) * V>::st(orow + j, &o);
      else {
#pragma unroll
        for (int v = 0; v < V; ++v) if (j + v < L) orow[(j + v) * oinner] = o.v[v];
      }
      cpre = cpre + ctot[u];
    }
    if (!more) break;
    carry = nb == cb ? cpre : scan_zero<T>();
    cb = nb;
    ct = nt;
    orow = orow_next;
#pragma unroll
    for (int u = 0; u < U; ++u) full[u] = full_next[u];
    par ^= 1;
  }
}

// fixed-order sum over the CTA (shuffle tree per warp, warp totals added in warp order); result in every thread
template <class T, int NW> __device__ __forceinline__ T scan_block_sum(T v, T *s_red) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = v + shfl_xor_t(v, m);
  __syncthreads();                       // previous use of s_red is over
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  T tot = s_red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) tot = tot + s_red[w];
  return tot;
}

// mode TILES, flat exchange (p.splits == 2; the measured default): a tile publishes its total; the LAST tile of each group
// of SCAN_GROUP tiles also publishes the group's total; a tile's carry = (totals of the groups before its own) + (totals of
// the tiles before it in its group): two flat, fixed-order sums of at most a few hundred L2-resident values gathered by all
// 256 threads at once — no serial chain between tiles, no atomics, and no dependence on which neighbour happened to
// finish first.  The next tile's loads are issued right after the publish and fly through the gather.
constexpr int SCAN_GROUP = 128;
template <class E, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void scan_tiles_flat_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  constexpr int NT = SCAN_NT, NW = NT / 32;
  __shared__ T s_warp[2][U][NW];
  __shared__ T s_red[NW + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 L = p.rsz[0];
  const i64 TILE = (i64)NT * V * U;
  const i64 tpr = (L + TILE - 1) / TILE;
  const i64 gpr = (tpr + SCAN_GROUP - 1) / SCAN_GROUP;
  const i64 total_tiles = p.B * tpr;
  const i64 oinner = p.out_rs[0];
  const u32 tag = ((__ldcg(p.scan_ctl) & 0x3fffffffu) << 2) | 1u;

  const char *base[E::NL];
  i64 inner[E::NL];
#pragma unroll
  for (int k = 0; k < E::NL; ++k) inner[k] = p.leaf[k].rs[0];
  auto setup_row = [&](i64 b) -> OutT * {
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
    i64 oo = 0;
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
    }
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    return (OutT *)p.out.ptr + oo;
  };
  i64 gid = blockIdx.x;
  const i64 gq = (i64)gridDim.x / tpr, gr = (i64)gridDim.x - gq * tpr;
  if (gid < total_tiles) {
    i64 cb = gid / tpr, ct = gid - cb * tpr;
    OutT *orow = setup_row(cb);
    typename E::template Regs<V> r[U];
    bool full[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 j = ct * TILE + ((i64)u * NT + tid) * V;
      full[u] = V == 1 ? (j < L) : (j + V <= L);
      if (full[u]) E::template loadv<V, UNIT>(r[u], base, inner, j);
    }
    int par = 0;
    while (true) {
      const i64 j0 = ct * TILE;
      T x[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = j0 + ((i64)u * NT + tid) * V;
        if (full[u]) {
#pragma unroll
          for (int v = 0; v < V; ++v) x[u][v] = E::template eval<V>(r[u], v, p.c);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            x[u][v] = scan_zero<T>();
            if (V > 1 && j + v < L) {
              typename E::template Regs<1> r1;
              E::template loadv<1, false>(r1, base, inner, j + v);
              x[u][v] = E::template eval<1>(r1, 0, p.c);
            }
          }
        }
#pragma unroll
        for (int v = 1; v < V; ++v) x[u][v] = x[u][v - 1] + x[u][v];
      }
      T wexcl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T incl = x[u][V - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const T o = shfl_up_t(incl, d);
          if (lane >= d) incl = o + incl;
        }
        const T ex = shfl_up_t(incl, 1);
        wexcl[u] = lane == 0 ? scan_zero<T>() : ex;
        if (lane == 31) s_warp[par][u][warp] = incl;
      }
      __syncthreads();   // (A)
      T wpre[U], ctot[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T run = scan_zero<T>();
        wpre[u] = run;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          if (w == warp) wpre[u] = run;
          run = run + s_warp[par][u][w];
        }
        ctot[u] = run;
      }
      T total = ctot[0];
#pragma unroll
      for (int u = 1; u < U; ++u) total = total + ctot[u];
      char *agg = (char *)p.scan_agg + (size_t)(cb * tpr) * ScanSlot<T>::STRIDE;
      char *gagg = (char *)p.scan_gagg + (size_t)(cb * gpr) * ScanSlot<T>::STRIDE;
      if (tid == 0) ScanSlot<T>::publish(agg, ct, total, tag);
      // the NEXT tile's loads go out right behind the publish
      const i64 ng = gid + gridDim.x;
      i64 nb = cb + gq, nt = ct + gr;          // (row, tile) of gid + gridDim.x without a division in the loop
      if (nt >= tpr) { nt -= tpr; ++nb; }
      const bool more = ng < total_tiles;
      OutT *orow_next = orow;
      bool full_next[U];
      if (more) {
        if (nb != cb) orow_next = setup_row(nb);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j = nt * TILE + ((i64)u * NT + tid) * V;
          full_next[u] = V == 1 ? (j < L) : (j + V <= L);
          if (full_next[u]) E::template loadv<V, UNIT>(r[u], base, inner, j);
        }
      }
      const i64 g = ct / SCAN_GROUP, first = g * SCAN_GROUP;
      const i64 gcount = (tpr - first) < SCAN_GROUP ? (tpr - first) : SCAN_GROUP;
      const i64 n1 = g, n2 = ct - first;
      T carry;
      if (ct == first + gcount - 1) {
        T a = scan_zero<T>();
        for (i64 i = tid; i < n2; i += NT) a = a + ScanSlot<T>::wait(agg, first + i, tag);
        const T stiles = scan_block_sum<T, NW>(a, s_red);
        if (tid == 0) ScanSlot<T>::publish(gagg, g, stiles + total, tag);
        a = scan_zero<T>();
        for (i64 i = tid; i < n1; i += NT) a = a + ScanSlot<T>::wait(gagg, i, tag);
        carry = scan_block_sum<T, NW>(a, s_red) + stiles;
      } else {
        T acc = scan_zero<T>();
        for (i64 i = tid; i < n1 + n2; i += NT)
          acc = acc + (i < n1 ? ScanSlot<T>::wait(gagg, i, tag) : ScanSlot<T>::wait(agg, first + (i - n1), tag));
        carry = scan_block_sum<T, NW>(acc, s_red);
      }
      T cpre = carry;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = j0 + ((i64)u * NT + tid) * V;
        const T pre = (cpre + wpre[u]) + wexcl[u];
        Vec<OutT, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(pre + x[u][v]);
        if (V > 1 && j + V <= L && oinner == 1 && p.tx) StBytes<(int)sizeof(OutT) * V>::st(orow + j, &o);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) if (j + v < L) orow[(j + v) * oinner] = o.v[v];
        }
        cpre = cpre + ctot[u];
      }
      if (!more) break;
      cb = nb;
      ct = nt;
      gid = ng;
      orow = orow_next;
#pragma unroll
      for (int u = 0; u < U; ++u) full[u] = full_next[u];
      par ^= 1;
    }
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    __threadfence();
    if (atomicInc(p.scan_ctl + 1, gridDim.x - 1) == gridDim.x - 1) {
      u32 e = (__ldcg(p.scan_ctl) + 1u) & 0x3fffffffu;
      *(volatile u32 *)p.scan_ctl = e ? e : 1u;
    }
  }
}

// mode TILES: few long rows.  The tiles of all rows are dealt round-robin to a grid whose CTAs are all resident
// (cooperative launch).  Pipelined like select1p: phase 1 of a tile (load, tile-local inclusive scan, publish the tile
// total, park the scanned tile in a shared-memory slot) runs D - 1 iterations before its phase 2 (carry from the
// TileExchange jobs, add, store), so the exchange's L2 round trips happen while the CTA works on later tiles.
template <class E, class OutT, int V, int U, bool UNIT>
__device__ __forceinline__ void scan_tiles_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  constexpr int NT = SCAN_NT, NW = NT / 32, DMAX = 8;
  extern __shared__ __align__(16) unsigned char scan_smem[];
  T *stage = (T *)scan_smem;        // [D][NT * U * V] tile-local inclusive scans
  __shared__ T s_warp[2][U][NW];
  __shared__ T s_tot[DMAX];
  __shared__ T s_carry[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.scan_depth;
  const i64 L = p.rsz[0];
  const i64 TILE = (i64)NT * V * U;
  const i64 tpr = (L + TILE - 1) / TILE;
  const i64 total_tiles = p.B * tpr;
  const i64 G = gridDim.x;
  const i64 mine = total_tiles > (i64)blockIdx.x ? (total_tiles - blockIdx.x + G - 1) / G : 0;
  const i64 oinner = p.out_rs[0];
  const u32 tag = ((__ldcg(p.scan_ctl) & 0x3fffffffu) << 2) | 1u;
  TileExchange<T> xc;
  xc.init(p.scan_agg, p.scan_gagg, p.scan_sagg, tpr, tag, lane);

  const char *base[E::NL];
  i64 inner[E::NL];
#pragma unroll
  for (int k = 0; k < E::NL; ++k) inner[k] = p.leaf[k].rs[0];
  auto setup_in = [&](i64 b) {
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
    }
  };
  auto out_row = [&](i64 b) -> OutT * {
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
    i64 oo = 0;
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    return (OutT *)p.out.ptr + oo;
  };
  typename E::template Regs<V> r[U];
  bool full[U];
  i64 loaded_row = -1;
  auto issue_loads = [&](i64 b, i64 t) {
    if (b != loaded_row) { setup_in(b); loaded_row = b; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 j = t * TILE + ((i64)u * NT + tid) * V;
      full[u] = V == 1 ? (j < L) : (j + V <= L);
      if (full[u]) E::template loadv<V, UNIT>(r[u], base, inner, j);
    }
  };
  // (row, tile in the row) of this CTA's tiles, stepped by G tiles per iteration without a division in the loop (six 64-bit
  // divisions and three modulos per iteration were most of the instructions of this kernel)
  const i64 gq = G / tpr, gr = G - gq * tpr;
  auto step = [&](i64 &b, i64 &t) { t += gr; b += gq; if (t >= tpr) { t -= tpr; ++b; } };
  i64 cb = (i64)blockIdx.x / tpr, ct = (i64)blockIdx.x - cb * tpr;   // tile of phase 1 (iteration `it`)
  i64 b1 = 0, t1 = -1, b2 = 0, t2 = -1;                               // tiles of iterations it - 1 and it - 2 (none yet)
  i64 cb2 = cb, ct2 = ct;                                             // tile of phase 2 (iteration it - (D - 1))
  int slot = 0;
  if (mine > 0) issue_loads(cb, ct);
  for (i64 it = 0; it < mine + D - 1; ++it) {
    const int par = (int)(it & 1);
    const bool p1 = it < mine;
    const bool p2 = it >= D - 1;
    const int slot2 = slot + 1 == D ? 0 : slot + 1;      // (it - (D - 1)) mod D
    const int slot1 = slot == 0 ? D - 1 : slot - 1;      // (it - 1) mod D
    if (warp == 0) xc.issue(b1, (it >= 1 && it - 1 < mine) ? t1 : -1, b2, (it >= 2 && it - 2 < mine) ? t2 : -1, cb2, p2 ? ct2 : -1);
    i64 nb = cb, nt = ct;
    step(nb, nt);
    if (p1) {
      const i64 j0 = ct * TILE;
      T x[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = j0 + ((i64)u * NT + tid) * V;
        if (full[u]) {
#pragma unroll
          for (int v = 0; v < V; ++v) x[u][v] = E::template eval<V>(r[u], v, p.c);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            x[u][v] = scan_zero<T>();
            if (V > 1 && j + v < L) {
              typename E::template Regs<1> r1;
              E::template loadv<1, false>(r1, base, inner, j + v);
              x[u][v] = E::template eval<1>(r1, 0, p.c);
            }
          }
        }
#pragma unroll
        for (int v = 1; v < V; ++v) x[u][v] = x[u][v - 1] + x[u][v];
      }
      if (it + 1 < mine) issue_loads(nb, nt);   // the next tile's loads fly while this one is scanned and parked
      T wexcl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T incl = x[u][V - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const T o = shfl_up_t(incl, d);
          if (lane >= d) incl = o + incl;
        }
        const T ex = shfl_up_t(incl, 1);
        wexcl[u] = lane == 0 ? scan_zero<T>() : ex;
        if (lane == 31) s_warp[par][u][warp] = incl;
      }
      __syncthreads();   // (A)
      T cpre = scan_zero<T>();
      T *stg = stage + (size_t)slot * TILE;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T run = scan_zero<T>(), wpre = run;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          if (w == warp) wpre = run;
          run = run + s_warp[par][u][w];
        }
        const T pre = (cpre + wpre) + wexcl[u];
        Vec<T, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = pre + x[u][v];
        *(Vec<T, V> *)(stg + ((size_t)u * NT + tid) * V) = o;
        cpre = cpre + run;
      }
      if (tid == 0) {   // the tile's total: every later tile of the row waits for this store
        xc.publish_tile(cb, ct, cpre);
        s_tot[slot] = cpre;
      }
    } else {
      __syncthreads();   // (A) keeps the barrier count uniform while the pipeline drains
    }
    if (warp == 0) {
      const T ctot = xc.close_ct >= 0 ? s_tot[slot1] : scan_zero<T>();
      const T cr = xc.finish(ctot);
      if (lane == 0) s_carry[par] = cr;
    }
    __syncthreads();   // (B)
    if (p2) {
      const T carry = s_carry[par];
      OutT *orow = out_row(cb2);
      const T *stg = stage + (size_t)slot2 * TILE;
      const i64 j0 = ct2 * TILE;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j = j0 + ((i64)u * NT + tid) * V;
        const Vec<T, V> x = *(const Vec<T, V> *)(stg + ((size_t)u * NT + tid) * V);
        Vec<OutT, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(carry + x.v[v]);
        if (V > 1 && j + V <= L && oinner == 1 && p.tx) StBytes<(int)sizeof(OutT) * V>::st(orow + j, &o);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) if (j + v < L) orow[(j + v) * oinner] = o.v[v];
        }
      }
      step(cb2, ct2);
    }
    b2 = b1; t2 = t1;
    b1 = cb; t1 = ct;
    cb = nb; ct = nt;
    slot = slot + 1 == D ? 0 : slot + 1;
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // exit ticket: the last CTA out opens the next epoch (every slot of this launch is stale from then on)
  if (tid == 0) {
    __threadfence();
    if (atomicInc(p.scan_ctl + 1, gridDim.x - 1) == gridDim.x - 1) {
      u32 e = (__ldcg(p.scan_ctl) + 1u) & 0x3fffffffu;
      *(volatile u32 *)p.scan_ctl = e ? e : 1u;
    }
  }
}

// mode TILES on WARP tiles (p.splits == 4, the default): the single-pass design of select1p applied to the prefix sum of
// few long rows.  A warp owns tiles of 32 lanes x EPL elements (4 KB).  Phase 1 of a tile reads it and publishes its total;
// the exchange runs as jobs 1 and 2 iterations later (group totals; supergroup totals + running totals — fixed-shape sums,
// so the result does not depend on timing); phase 2, three iterations later, reads the tile AGAIN — L2 is asked to keep
// it in between — scans it in registers, adds the carry and stores.  Nothing is kept on chip between the phases, so there
// is no ring, no shared memory and no barrier.
template <class T> struct ScanXchg {
  typedef ScanSlot<T> SL;
  char *agg, *gagg, *run, *own;      // slots of row 0; row r sits tpr / gpr / spr + 1 / spr slots further
  i64 tpr, gpr, spr;
  u32 tag;
  int lane;
  __device__ __forceinline__ char *A(i64 row) const { return agg + (size_t)(row * tpr) * SL::STRIDE; }
  __device__ __forceinline__ char *G(i64 row) const { return gagg + (size_t)(row * gpr) * SL::STRIDE; }
  __device__ __forceinline__ char *R(i64 row) const { return run + (size_t)(row * (spr + 1)) * SL::STRIDE; }
  __device__ __forceinline__ char *O(i64 row) const { return own + (size_t)(row * spr) * SL::STRIDE; }
  __device__ __forceinline__ T get(const char *slots, i64 i, bool on) const { return on ? SL::wait(slots, i, tag) : scan_zero<T>(); }
  __device__ __forceinline__ bool closes_group(i64 ct) const { return ct >= 0 && ((ct & 31) == 31 || ct == tpr - 1); }
  __device__ __forceinline__ bool closes_super(i64 ct) const { return ct >= 0 && ct < tpr - 1 && (ct & 1023) == 1023; }
  __device__ __forceinline__ void close(i64 row, i64 ct, T total) const {
    const i64 first = (ct >> 5) << 5;
    const T s = warp_tree_sum(get(A(row), first + lane, first + lane < ct));
    if (lane == 0) SL::publish(G(row), ct >> 5, s + total, tag);
  }
  __device__ __forceinline__ void super(i64 row, i64 ct) const {
    const i64 sg = ct >> 10;
    const T s = warp_tree_sum(get(G(row), (sg << 5) + lane, true));
    if (lane == 0) SL::publish(O(row), sg, s, tag);
  }
  __device__ __forceinline__ void running(i64 row, i64 ct) const {
    const i64 sg = ct >> 10;
    T acc = scan_zero<T>();
    for (i64 i0 = 0; i0 <= sg; i0 += 32) acc = acc + warp_tree_sum(get(O(row), i0 + lane, i0 + lane <= sg));   // fixed order
    if (lane == 0) SL::publish(R(row), sg + 1, acc, tag);
  }
  __device__ __forceinline__ T carry(i64 row, i64 ct) const {
    const i64 g = ct >> 5, first = g << 5, sg = ct >> 10, gfirst = sg << 5;
    const T a = get(A(row), first + lane, first + lane < ct);
    const T b = get(G(row), gfirst + lane, gfirst + lane < g);
    T c = get(R(row), sg, lane == 0 && sg > 0);
    const T sa = warp_tree_sum(a), sb = warp_tree_sum(b);
    c = shfl_idx_t(c, 0);
    return (c + sb) + sa;
  }
};

// 16-byte store that leaves L2 first: the scan's output must not push the tiles waiting for their second read out of L2
__device__ __forceinline__ void st16_evict_first(void *d, const void *s) {
  const u32 *x = (const u32 *)s;
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(d), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "l"(pol) : "memory");
}

__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void ld16_hint(void *dst, const void *src, unsigned long long pol) {
  u32 *x = (u32 *)dst;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "l"(src), "l"(pol));
}

template <class E, class OutT, int V, bool UNIT, int EPL4>
__device__ __forceinline__ void scan_wtiles_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  typedef typename E::template Regs<V> R;
  constexpr int EPL = (sizeof(T) > 4 ? EPL4 / 2 : EPL4) < V ? V : (sizeof(T) > 4 ? EPL4 / 2 : EPL4);      // elements per lane and tile
  constexpr int U = EPL / V;                        // vectors per lane and tile
  constexpr i64 TILE = (i64)32 * EPL;
  static_assert(EPL % V == 0 && U >= 1, "vector width must divide the elements per lane");
  const int lane = threadIdx.x & 31;
  const i64 wpc = blockDim.x >> 5;
  const i64 nwarp = (i64)gridDim.x * wpc, gw = (i64)blockIdx.x * wpc + (threadIdx.x >> 5);
  const i64 L = p.rsz[0];
  const i64 tpr = (L + TILE - 1) / TILE;
  const i64 total_tiles = p.B * tpr;
  const i64 oinner = p.out_rs[0];
  ScanXchg<T> xc;
  xc.agg = (char *)p.scan_agg; xc.gagg = (char *)p.scan_gagg; xc.run = (char *)p.scan_sagg; xc.own = (char *)p.scan_own;
  xc.tpr = tpr; xc.gpr = (tpr + 31) >> 5; xc.spr = (tpr + 1023) >> 10;
  xc.tag = ((__ldcg(p.scan_ctl) & 0x3fffffffu) << 2) | 1u;
  xc.lane = lane;
  i64 inner[E::NL];
#pragma unroll
  for (int k = 0; k < E::NL; ++k) inner[k] = p.leaf[k].rs[0];
  auto row_bases = [&](i64 b, const char **base) -> OutT * {
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
    i64 oo = 0;
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
    }
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    return (OutT *)p.out.ptr + oo;
  };
  // element values of this lane's part of tile (row base, ct): zero beyond the row's end
  // a plain contiguous operand is read with an L2 policy: the first read asks L2 to keep the tile (evict_last), the second
  // one — its last use — releases it (evict_first); so do the output stores
  const bool plain = p.scan_plain != 0 && (int)sizeof(T) * V == 16 && UNIT;
  const unsigned long long pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  auto load_tile = [&](const char *const *base, i64 ct, T (&x)[U][V], unsigned long long pol) {
    const i64 j0 = ct * TILE + (i64)lane * V;
    if (plain && (ct + 1) * TILE <= L) {
#pragma unroll
      for (int u = 0; u < U; ++u) ld16_hint(&x[u][0], base[0] + (j0 + (i64)u * 32 * V) * (i64)sizeof(T), pol);
    } else if ((ct + 1) * TILE <= L) {
      R r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) E::template loadv<V, UNIT>(r[u], base, inner, j0 + (i64)u * 32 * V);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < V; ++v) x[u][v] = E::template eval<V>(r[u], v, p.c);
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const i64 j = j0 + (i64)u * 32 * V + v;
          x[u][v] = scan_zero<T>();
          if (j < L) {
            typename E::template Regs<1> r1;
            E::template loadv<1, false>(r1, base, inner, j);
            x[u][v] = E::template eval<1>(r1, 0, p.c);
          }
        }
      }
    }
  };
  // (row, tile in the row) of the tiles of the last five iterations, stepped by nwarp tiles without a division in the loop
  const i64 gq = nwarp / tpr, gr = nwarp - gq * tpr;
  // DIST iterations lie between a tile's two reads.  4 (one iteration per level of the exchange) never waits but leaves
  // 58 MB between the reads: 0.53 ms for 2^28 fp32; 3 (supergroup total and running total in one iteration) 0.48 ms with
  // 65 % of the second reads served by L2; 2 (all jobs in one iteration) makes the warps wait on each other: 0.63 ms
  constexpr int DIST = 3;
  i64 rb[DIST + 1], rt[DIST + 1];   // [0] this iteration ... [DIST] DIST iterations ago
#pragma unroll
  for (int k = 0; k <= DIST; ++k) { rb[k] = 0; rt[k] = -1; }
  i64 nb = gw / tpr, nt = gw - nb * tpr;   // coordinates of `tile` before the loop's first step
  T prev_total = scan_zero<T>();
  for (i64 tile = gw; tile - DIST * nwarp < total_tiles; tile += nwarp) {
#pragma unroll
    for (int k = DIST; k > 0; --k) { rb[k] = rb[k - 1]; rt[k] = rt[k - 1]; }
    const bool p1 = tile < total_tiles;
    rb[0] = nb; rt[0] = p1 ? nt : -1;
    nt += gr; nb += gq;
    if (nt >= tpr) { nt -= tpr; ++nb; }
    // ---- phase 1: read the tile, its total ----
    T total = scan_zero<T>();
    if (p1) {
      const char *base[E::NL];
      row_bases(rb[0], base);
      T x[U][V];
      load_tile(base, rt[0], x, pol_keep);
      T s = scan_zero<T>();
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < V; ++v) s = s + x[u][v];
      }
      total = warp_tree_sum(s);
      if (lane == 0) ScanSlot<T>::publish(xc.A(rb[0]), rt[0], total, xc.tag);
    }
    // ---- exchange jobs of the tiles read 1, 2 and 3 iterations ago ----
    if (xc.closes_group(rt[1])) xc.close(rb[1], rt[1], prev_total);
    if (xc.closes_super(rt[2])) { xc.super(rb[2], rt[2]); xc.running(rb[2], rt[2]); }
    prev_total = total;
    // ---- phase 2 of the tile read DIST iterations ago: carry, second read (L2), scan in registers, store ----
    if (rt[DIST] >= 0) {
      const char *base[E::NL];
      OutT *orow = row_bases(rb[DIST], base);
      const i64 ct = rt[DIST];
      T x[U][V];
      load_tile(base, ct, x, pol_drop);
      const T carry = xc.carry(rb[DIST], ct);
      T rowbase = carry;
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 1; v < V; ++v) x[u][v] = x[u][v - 1] + x[u][v];
        T incl = x[u][V - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const T o = shfl_up_t(incl, d);
          if (lane >= d) incl = o + incl;
        }
        const T exl = shfl_up_t(incl, 1);
        const T pre = lane == 0 ? rowbase : rowbase + exl;
        const i64 j = ct * TILE + ((i64)u * 32 + lane) * V;
        Vec<OutT, V> o;
#pragma unroll
        for (int v = 0; v < V; ++v) o.v[v] = cvt<OutT>(pre + x[u][v]);
        if (V > 1 && j + V <= L && oinner == 1 && p.tx) {
          if (sizeof(OutT) * V == 16) st16_evict_first(orow + j, &o);
          else StBytes<(int)sizeof(OutT) * V>::st(orow + j, &o);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) if (j + v < L) orow[(j + v) * oinner] = o.v[v];
        }
        rowbase = rowbase + shfl_idx_t(incl, 31);
      }
    }
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // exit ticket: the last CTA out opens the next epoch (every slot of this launch is stale from then on)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(p.scan_ctl + 1, gridDim.x - 1) == gridDim.x - 1) {
      u32 e = (__ldcg(p.scan_ctl) + 1u) & 0x3fffffffu;
      *(volatile u32 *)p.scan_ctl = e ? e : 1u;
    }
  }
}

template <class E, class OutT, int V, int U, int TEAM>
__device__ __forceinline__ void scan_inner_body(const RedParams &p) {
  pdl_prologue();
  if (TEAM == 1) {
    if (p.all_unit) scan_warp_body_impl<E, OutT, V, U, true>(p);
    else scan_warp_body_impl<E, OutT, V, U, false>(p);
  } else if (p.splits == 4) {
    if (p.scan_depth == 16) {
      if (p.all_unit) scan_wtiles_body_impl<E, OutT, V, true, 16>(p);
      else scan_wtiles_body_impl<E, OutT, V, false, 16>(p);
    } else {
      if (p.all_unit) scan_wtiles_body_impl<E, OutT, V, true, 32>(p);
      else scan_wtiles_body_impl<E, OutT, V, false, 32>(p);
    }
  } else if (p.splits == 2) {
    if (p.all_unit) scan_tiles_flat_body_impl<E, OutT, V, U, true>(p);
    else scan_tiles_flat_body_impl<E, OutT, V, U, false>(p);
  } else if (p.splits > 2) {
    if (p.all_unit) scan_tiles_body_impl<E, OutT, V, U, true>(p);
    else scan_tiles_body_impl<E, OutT, V, U, false>(p);
  } else {
    if (p.all_unit) scan_inner_body_impl<E, OutT, V, U, true>(p);
    else scan_inner_body_impl<E, OutT, V, U, false>(p);
  }
}

// ------------------------------------------------------------------------------------------------
// S1: select — stream compaction behind find / find_idx (reference: find_impl / find_idx_impl, transforms/cub.h:2609-2790,
// cub::DeviceSelect::If over the row-major flattened operator with the functors LT / GT / EQ / NEQ / LTE / GTE,
// :2521-2588).  Output order = flat index order (stable), count = number selected.  Two launches, no spinning:
//   count pass    a tile = 256 threads x U chunks x V elements of the flat index space; CTA c owns a contiguous run of
//                 tiles, its threads count their selected elements (popc of the flag masks) with no barrier in the loop,
//                 one CTA total goes to sel_counts[c]; the LAST CTA to finish (self-resetting ticket) turns the totals
//                 into exclusive offsets and writes num_found;
//   scatter pass  CTA c walks the same tiles with a running output position that starts at offsets[c]; an element's
//                 position is that + (chunks before in the tile) + (warps before in its chunk) + (lanes before in its
//                 warp: shuffle scan) + (flags before in its vector); values (MODE 1) or flat indices (MODE 2) go there.
// The input is read twice (the second read comes from L2 when it fits); CUB's single-pass decoupled look-back reads it
// once — the tile-exchange machinery of `scan` is the route to that here.
// ------------------------------------------------------------------------------------------------
constexpr int SEL_NT = 256, SEL_U = 4;
template <class T> __device__ __forceinline__ bool sel_test(T x, int op, T c) {
  switch (op) {
    case 0: return x < c;
    case 1: return x > c;
    case 2: return x == c;
    case 3: return x != c;
    case 4: return x <= c;
    default: return x >= c;
  }
}
// branch-free form (single-pass kernel): category of x against c (less / greater / equal / unordered) indexes
// a 4-bit acceptance mask (LT 0001, GT 0010, EQ 0100, NEQ 1011, LTE 0101, GTE 0110)
__device__ __forceinline__ u32 sel_mask(int op) { return (0x65B421u >> (4 * op)) & 0xFu; }
template <class T> __device__ __forceinline__ u32 sel_flag(T x, T c, u32 mask) {
  const u32 cat = x < c ? 0u : (x > c ? 1u : (x == c ? 2u : 3u));
  return (mask >> cat) & 1u;
}
template <class T> struct SelThr { static __device__ __forceinline__ T get(const EwParams &p) { return (T)p.sel_thr_d; } };
template <> struct SelThr<int> { static __device__ __forceinline__ int get(const EwParams &p) { return (int)p.sel_thr_i; } };
template <> struct SelThr<i64> { static __device__ __forceinline__ i64 get(const EwParams &p) { return p.sel_thr_i; } };
template <> struct SelThr<unsigned char> { static __device__ __forceinline__ unsigned char get(const EwParams &p) { return (unsigned char)p.sel_thr_i; } };

// flags (bit v = element v selected) and values of the V elements starting at flat index j0
template <class E, int V>
__device__ __forceinline__ u32 sel_eval(const EwParams &p, i64 j0, typename E::value_type thr, typename E::value_type *vals) {
  typedef typename E::value_type T;
  u32 flags = 0;
  if (j0 >= p.N) return 0;
  const int nd = p.nd;
  if (nd == 1) {
    const char *base[E::NL];
    i64 inner[E::NL];
#pragma unroll
    for (int k = 0; k < E::NL; ++k) { base[k] = (const char *)p.leaf[k].ptr; inner[k] = p.leaf[k].bs[0]; }
    if (V > 1 && j0 + V <= p.N && p.all_unit) {
      typename E::template Regs<V> r;
      E::template loadv<V, true>(r, base, inner, j0);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        vals[v] = E::template eval<V>(r, v, p.c);
        flags |= (u32)sel_test<T>(vals[v], p.sel_op, thr) << v;
      }
      return flags;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (j0 + v < p.N) {
        typename E::template Regs<1> r;
        E::template loadv<1, false>(r, base, inner, j0 + v);
        vals[v] = E::template eval<1>(r, 0, p.c);
        flags |= (u32)sel_test<T>(vals[v], p.sel_op, thr) << v;
      }
    }
    return flags;
  }
  // N-D view that does not collapse: one element at a time through the row-major decomposition (V == 1 by host rule)
#pragma unroll
  for (int v = 0; v < V; ++v) {
    if (j0 + v < p.N) {
      i64 idx[KMAXD];
      decomp(j0 + v, nd, p.sz, idx);
      const char *base[E::NL];
      i64 inner[E::NL];
#pragma unroll
      for (int k = 0; k < E::NL; ++k) {
        i64 off = 0;
#pragma unroll
        for (int d = 0; d < KMAXD - 1; ++d) if (d < nd - 1) off += idx[d] * p.leaf[k].bs[d];
        base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
        inner[k] = p.leaf[k].bs[nd - 1];
      }
      typename E::template Regs<1> r;
      E::template loadv<1, false>(r, base, inner, idx[nd - 1]);
      vals[v] = E::template eval<1>(r, 0, p.c);
      flags |= (u32)sel_test<T>(vals[v], p.sel_op, thr) << v;
    }
  }
  return flags;
}

template <class E, class OutT, int V, int MODE_IN>
__device__ __forceinline__ void select_body(const EwParams &p) {
  constexpr int MODE = MODE_IN;             // 0 count (+ in-launch scan of the CTA totals), 1 scatter values, 2 scatter flat indices
  pdl_prologue();
  typedef typename E::value_type T;
  constexpr int NT = SEL_NT, U = SEL_U, NW = NT / 32;
  __shared__ u32 s_w[U][NW];
  __shared__ unsigned long long s_seg[NT];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 TILE = (i64)NT * V * U;
  const i64 ntiles = (p.N + TILE - 1) / TILE;
  const T thr = SelThr<T>::get(p);
  // both passes give CTA c the SAME contiguous run of tiles, so only one total per CTA crosses the grid
  const i64 tpc = (ntiles + gridDim.x - 1) / gridDim.x;
  const i64 t0 = (i64)blockIdx.x * tpc, t1 = (t0 + tpc < ntiles) ? (t0 + tpc) : ntiles;

  if (MODE == 0) {
    u32 c = 0;   // per thread: at most tpc * U * V elements
    for (i64 tile = t0; tile < t1; ++tile) {
      T vals[V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const i64 j0 = tile * TILE + ((i64)u * NT + tid) * V;
        c += (u32)__popc(sel_eval<E, V>(p, j0, thr, vals));
      }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_w[0][warp] = c;
    __syncthreads();
    if (tid == 0) {
      unsigned long long t = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += s_w[0][w];
      __stcg(p.sel_counts + blockIdx.x, t);
    }
    // grid stage: the last CTA to arrive scans the per-CTA totals (fixed order) and publishes num_found
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const u32 t = atomicInc(p.sel_ticket, gridDim.x - 1);
      s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const i64 G = gridDim.x;
      const i64 seg = (G + NT - 1) / NT, b0 = (i64)tid * seg, b1 = (b0 + seg < G) ? (b0 + seg) : G;
      unsigned long long sum = 0;
      for (i64 i = b0; i < b1; ++i) sum += __ldcg(p.sel_counts + i);
      s_seg[tid] = sum;
      __syncthreads();
      unsigned long long run = 0;
      for (int t = 0; t < tid; ++t) run += s_seg[t];
      if (tid == NT - 1) {
        const unsigned long long total = run + sum;
        *p.sel_total = total > 0x7fffffffull ? 0x7fffffff : (int)total;
      }
      for (i64 i = b0; i < b1; ++i) {
        const unsigned long long c2 = __ldcg(p.sel_counts + i);
        __stcg(p.sel_offsets + i, run);
        run += c2;
      }
    }
    return;
  }

  unsigned long long pos = __ldcg(p.sel_offsets + blockIdx.x);   // running output position of this CTA
  for (i64 tile = t0; tile < t1; ++tile) {
    u32 flags[U], excl[U];
    T vals[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 j0 = tile * TILE + ((i64)u * NT + tid) * V;
      flags[u] = sel_eval<E, V>(p, j0, thr, vals[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const u32 c = (u32)__popc(flags[u]);
      u32 incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      excl[u] = incl - c;
      if (lane == 31) s_w[u][warp] = incl;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      u32 before = 0, chunk = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const u32 t = s_w[u][w]; chunk += t; if (w < warp) before += t; }
      const i64 j0 = tile * TILE + ((i64)u * NT + tid) * V;
      unsigned long long q = pos + before + excl[u];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if ((flags[u] >> v) & 1u) {
          if ((i64)q < p.sel_cap) ((OutT *)p.out.ptr)[q] = MODE == 1 ? cvt<OutT>(vals[u][v]) : (OutT)(j0 + v);
          ++q;
        }
      }
      pos += chunk;
    }
    __syncthreads();   // s_w is reused by the next tile
  }
}

// ------------------------------------------------------------------------------------------------
// S1p: select1p — find / find_idx in ONE pass: every element is read once, the selected ones are written once.
// (reference: cub::DeviceSelect::If behind find_impl / find_idx_impl, transforms/cub.h:912-1010,2609-2790.)
// The unit of work is a WARP tile: 32 lanes x EPL elements (EPL = 32, 16 for 8-byte values) of the flat index space,
// as 16- or 32-byte vector loads (lane-adjacent vectors: coalesced).  Warp tiles are dealt round-robin to the warps of a
// grid whose CTAs are all resident (cooperative launch), so a tile only ever waits for tiles that are running or done.
// There is no shared memory and no barrier: a warp never waits for another warp of its CTA, only for the published
// counts of earlier tiles, and the other warps of the SM (32 resident) cover that wait.
//   1. branch-free predicate -> one flag bit per element in ONE register per lane;
//   2. ranks inside the tile: the per-vector counts of a lane travel packed (a byte each; 16 bits for vectors of more
//      than 7 elements) through shuffle scans, the per-vector warp totals come from lane 31;
//   3. the tile publishes its count and gathers its output offset from the published counts before it (sel_carry: tile
//      counts of its group of 32, group counts of its supergroup of 32 groups, running count at the supergroup start —
//      no chain: a supergroup's closer sums the supergroups' OWN totals, so the depth of the dependency is constant);
//   4. the selected values (or flat indices) go straight to out[offset + rank]: ranks grow with the lane, so a store
//      instruction covers a run of at most 32 x V consecutive outputs.
// The output is deterministic and stable.  Status words carry the launch epoch: nothing is cleared between launches.
// ------------------------------------------------------------------------------------------------
// predicate of one selection op as a type: the tile loop is instantiated per op (one compare + one predicated OR per
// element) instead of classifying every element against a runtime mask
template <int OP> struct SelPred {
  template <class T> static __device__ __forceinline__ bool test(T x, T c) {
    return OP == 0 ? (x < c) : OP == 1 ? (x > c) : OP == 2 ? (x == c) : OP == 3 ? (x != c) : OP == 4 ? (x <= c) : (x >= c);
  }
};

// The exchange of the published counts, as jobs that a warp runs for tiles it ranked 1, 2, 3 and 4 iterations earlier (so
// every job finds its inputs published by the other warps' previous iteration instead of waiting a round trip for them):
//   close  (+1)  a tile that ends its group of 32 sums the group's tile counts          -> group count
//   super  (+2)  a tile that ends its supergroup of 32 groups sums the group counts     -> the supergroup's OWN count
//   run    (+3)  the same tile sums the own counts of every supergroup up to its own    -> running count at the next start
//   carry  (+4)  running count at the supergroup start + group counts before + tile counts before = the tile's offset
// No job depends on a job of the same kind, so the dependency depth is constant however many tiles there are.
struct SelExchange {
  typedef ScanSlot<u32> SL;
  unsigned long long *agg, *gagg, *run, *own;
  i64 ntiles;
  u32 tag;
  u32 nap;     // back-off of a waiting lane between two polls (ns)
  int lane;
  __device__ __forceinline__ u32 get(const unsigned long long *slots, i64 i, bool on) const {
    u32 v = 0;
    if (on) {
      SL::Word w = SL::peek(slots, i);
      while (!SL::ready(w, tag)) { __nanosleep(nap); w = SL::peek(slots, i); }
      v = SL::value(w);
    }
    return v;
  }
  __device__ __forceinline__ bool closes_group(i64 ct) const { return ct >= 0 && ct < ntiles && ((ct & 31) == 31 || ct == ntiles - 1); }
  __device__ __forceinline__ bool closes_super(i64 ct) const { return ct >= 0 && ct < ntiles - 1 && (ct & 1023) == 1023; }
  __device__ __forceinline__ void close(i64 ct, u32 total) const {
    const i64 first = (ct >> 5) << 5;
    const u32 s = __reduce_add_sync(0xffffffffu, get(agg, first + lane, first + lane < ct));
    if (lane == 0) SL::publish(gagg, ct >> 5, s + total, tag);
  }
  __device__ __forceinline__ void super(i64 ct) const {
    if (!closes_super(ct)) return;
    const i64 sg = ct >> 10;
    const u32 s = __reduce_add_sync(0xffffffffu, get(gagg, (sg << 5) + lane, true));
    if (lane == 0) SL::publish(own, sg, s, tag);
  }
  __device__ __forceinline__ void running(i64 ct) const {
    if (!closes_super(ct)) return;
    const i64 sg = ct >> 10;
    u32 acc = 0;
    for (i64 i = lane; i <= sg; i += 32) acc += get(own, i, true);
    acc = __reduce_add_sync(0xffffffffu, acc);
    if (lane == 0) SL::publish(run, sg + 1, acc, tag);
  }
  __device__ __forceinline__ u32 carry(i64 ct) const {
    const i64 g = ct >> 5, first = g << 5, sg = ct >> 10, gfirst = sg << 5;
    // the three reads go out together
    const bool on_a = first + lane < ct, on_b = gfirst + lane < g, on_c = lane == 0 && sg > 0;
    SL::Word wa, wb, wc;
    if (on_a) wa = SL::peek(agg, first + lane);
    if (on_b) wb = SL::peek(gagg, gfirst + lane);
    if (on_c) wc = SL::peek(run, sg);
    u32 v = 0;
    if (on_a) { while (!SL::ready(wa, tag)) { __nanosleep(nap); wa = SL::peek(agg, first + lane); } v += SL::value(wa); }
    if (on_b) { while (!SL::ready(wb, tag)) { __nanosleep(nap); wb = SL::peek(gagg, gfirst + lane); } v += SL::value(wb); }
    if (on_c) { while (!SL::ready(wc, tag)) { __nanosleep(nap); wc = SL::peek(run, sg); } v += SL::value(wc); }
    return __reduce_add_sync(0xffffffffu, v);
  }
};

constexpr int SEL_RING = 5;     // tile states a warp keeps: ranked in iteration i, written in iteration i + 4
constexpr int SEL_WARPS = 8;    // warps per CTA (256 threads)
template <class T, int V> struct SelGeom {
  // elements per lane and tile: 64 (2048-element warp tiles, read as two halves of eight 16-byte loads), 32 for 8-byte
  // values, 8 on the scalar (strided / broadcast) walk
  enum { EPL = V == 1 ? 8 : (sizeof(T) > 4 ? 32 : 64), U = EPL / V, UH = U / 2, TILE = 32 * EPL,
         CB = (32 * V <= 255) ? 8 : 16, CPW = 32 / CB, NWORD = (U + CPW - 1) / CPW,   // packed lane counts of the vector rows
         FW = (U * V + 31) / 32,                                                       // flag words per lane
         RSW = (U + 1) / 2,                                                            // packed row starts, 16 bits each (per warp)
         LANE_WORDS = FW + NWORD, SLOT_WORDS = LANE_WORDS * 32 + RSW + 1 };
};

// staging buffer of one vector row of a dense tile (32 lanes x V outputs per warp), when it fits 1 KB per warp
template <class OutT, int V> struct SelStage { enum { ON = (32 * V * (int)sizeof(OutT) <= 1024) ? 1 : 0, BYTES = ON ? 32 * V * (int)sizeof(OutT) : 4 }; };

// predicated store without a branch (the compiler turns `if (p) out[i] = v` into a divergence region per element)
template <class OutT> __device__ __forceinline__ void st_if(bool p, OutT *addr, OutT v) { if (p) *addr = v; }
template <> __device__ __forceinline__ void st_if<int>(bool p, int *addr, int v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.u32 [%1], %2; }" ::"r"((u32)p), "l"(addr), "r"(v) : "memory");
}
template <> __device__ __forceinline__ void st_if<float>(bool p, float *addr, float v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.f32 [%1], %2; }" ::"r"((u32)p), "l"(addr), "f"(v) : "memory");
}
template <> __device__ __forceinline__ void st_if<i64>(bool p, i64 *addr, i64 v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.u64 [%1], %2; }" ::"r"((u32)p), "l"(addr), "l"(v) : "memory");
}
template <> __device__ __forceinline__ void st_if<double>(bool p, double *addr, double v) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.f64 [%1], %2; }" ::"r"((u32)p), "l"(addr), "d"(v) : "memory");
}

template <class E, class OutT, int V, int MODE, bool UNIT, int OP>   // MODE 1: values, 2: flat indices; OP < 0: runtime op / unique
__device__ __forceinline__ void select1p_tiles(const EwParams &p, u32 *ring_all, void *stage_all) {
  typedef typename E::value_type T;
  typedef typename E::template Regs<V> R;
  typedef SelGeom<T, V> GEO;
  constexpr int U = GEO::U, UH = GEO::UH, CB = GEO::CB, CPW = GEO::CPW, NWORD = GEO::NWORD, RSW = GEO::RSW, FW = GEO::FW;
  constexpr i64 TILE = GEO::TILE;
  constexpr u32 CMASK = CB == 8 ? 0xffu : 0xffffu;
  constexpr u32 VMASK = V >= 32 ? 0xffffffffu : ((1u << V) - 1u);
  constexpr int VSH = V == 1 ? 0 : V == 2 ? 1 : V == 4 ? 2 : V == 8 ? 3 : V == 16 ? 4 : 5;
  constexpr int RPW = 32 / V;                         // vector rows per flag word
  constexpr bool STAGED = SelStage<OutT, V>::ON != 0; // dense tiles: rows compacted in shared memory before they are stored
  static_assert(U % 2 == 0 && UH >= 1 && (1 << VSH) == V && (U * V) % 32 == 0 || U * V < 32, "tile geometry");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const i64 wpc = blockDim.x >> 5;
  const i64 nwarp = (i64)gridDim.x * wpc, gw = (i64)blockIdx.x * wpc + warp;
  const i64 ntiles = (p.N + TILE - 1) / TILE;
  const T thr = SelThr<T>::get(p);
  const bool unique_mode = p.sel_op == 6;             // MXB_SEL_UNIQUE (mxb_unique): adjacent-difference flags over a sorted operand
  const u32 fmask = sel_mask(unique_mode ? 0 : p.sel_op);
  const u32 epoch = __ldcg(p.sel_epoch) & 0x3fffffffu;
  // slots: tile counts | group counts (32 tiles) | running counts at supergroup starts (1024 tiles) | supergroup counts
  const i64 ngroup = (ntiles + 31) >> 5, nsuper = (ntiles + 1023) >> 10;
  SelExchange xc;
  xc.agg = p.sel_status; xc.gagg = xc.agg + ntiles; xc.run = xc.gagg + ngroup; xc.own = xc.run + nsuper + 1;
  xc.ntiles = ntiles; xc.tag = (epoch << 2) | 1u; xc.lane = lane;
  xc.nap = (u32)(p.sel_depth & 0xffff);
  const u32 sparse_max = (u32)p.sel_depth >> 16;   // tiles with at most this many selected elements walk their set bits
  const char *base[E::NL];
  i64 inner[E::NL];
#pragma unroll
  for (int k = 0; k < E::NL; ++k) { base[k] = (const char *)p.leaf[k].ptr; inner[k] = p.leaf[k].bs[0]; }
  OutT *out = (OutT *)p.out.ptr;
  const u32 cap = p.sel_cap > 0xffffffffll ? 0xffffffffu : (u32)p.sel_cap;
  // ring slot of a tile, per warp: per-lane words [w][lane] — flag words, then the exclusive lane counts of the vector
  // rows (packed) — followed by words every lane shares: the rows' starts inside the tile (16 bits each) and the tile's count
  u32 *wring = ring_all + (size_t)warp * SEL_RING * GEO::SLOT_WORDS;
  auto slot_ptr = [&](int sl) { return wring + (size_t)sl * GEO::SLOT_WORDS; };

  int s0 = 0;   // ring slot of this iteration's tile; the tile of k iterations ago sits in slot (s0 - k) mod SEL_RING
  for (i64 tile = gw; tile - 4 * nwarp < ntiles; tile += nwarp) {
    const bool p1 = tile < ntiles;
    const i64 t0 = tile * TILE;
    const i64 jl = t0 + (i64)lane * V;              // first element of this lane's vector 0; vector u sits 32 * V further
    const bool fast = OP >= 0 && p1 && t0 + TILE <= p.N;
    // ---- the loads of the tile's first half go out first: the exchange jobs below run under their latency ----
    R r[UH];
    if (fast) {
#pragma unroll
      for (int u = 0; u < UH; ++u) E::template loadv<V, UNIT>(r[u], base, inner, jl + (i64)u * 32 * V);
    }
    // ---- exchange jobs of the tiles ranked 1, 2 and 3 iterations ago ----
    {
      const int s1 = s0 >= 1 ? s0 - 1 : s0 - 1 + SEL_RING;
      const i64 tb = tile - nwarp;
      if (xc.closes_group(tb)) xc.close(tb, slot_ptr(s1)[GEO::LANE_WORDS * 32 + RSW]);
      xc.super(tile - 2 * nwarp);
      xc.running(tile - 3 * nwarp);
    }
    // ---- the tile ranked 4 iterations ago: its offset ----
    const i64 td = tile - 4 * nwarp;
    const bool p4 = td >= 0 && td < ntiles;
    const int s4 = s0 >= 4 ? s0 - 4 : s0 - 4 + SEL_RING;
    u32 off = 0;
    if (p4) off = xc.carry(td);
    // ---- flags of this iteration's tile: first half, then the second half's loads, which fly through the write below ----
    u32 f[FW];
#pragma unroll
    for (int w = 0; w < FW; ++w) f[w] = 0;
    auto flags_fast = [&](int h) {
#pragma unroll
      for (int uu = 0; uu < UH; ++uu) {
        const int u = h * UH + uu;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (SelPred<OP < 0 ? 0 : OP>::test(E::template eval<V>(r[uu], v, p.c), thr)) f[(u * V + v) / 32] |= 1u << ((u * V + v) % 32);
        }
      }
    };
    if (fast) {
      flags_fast(0);
#pragma unroll
      for (int u = 0; u < UH; ++u) E::template loadv<V, UNIT>(r[u], base, inner, jl + (i64)(UH + u) * 32 * V);
    }
    // ---- the selected elements of the tile ranked 4 iterations ago go to their final places ----
    if (p4) {
      const u32 *sp = slot_ptr(s4);
      const u32 *shared_words = sp + GEO::LANE_WORDS * 32;
      const u32 cnt = shared_words[RSW];
      if (td == ntiles - 1 && lane == 0) {
        const unsigned long long all = (unsigned long long)off + cnt;
        *p.sel_total = all > 0x7fffffffull ? 0x7fffffff : (int)all;
      }
      const i64 jd = td * TILE + (i64)lane * V;
      const bool walk = cnt <= sparse_max || (MODE == 1 && ((td + 1) * TILE > p.N || V == 1));
      if (walk) {
        // few selected elements (or a tile that cannot be read again as vectors): walk the set bits of the lane; the
        // row's start and the lane's exclusive count come straight from the ring
#pragma unroll
        for (int w = 0; w < FW; ++w) {
          const u32 fw = sp[w * 32 + lane];
          u32 ff = fw;
          while (ff) {
            const int bit = __ffs((int)ff) - 1;
            ff &= ff - 1;
            const int u = w * RPW + (bit >> VSH), v = bit & (V - 1);
            const u32 exw = sp[(FW + u / CPW) * 32 + lane];
            const u32 rsw = shared_words[u >> 1];
            const u32 below = (fw >> ((bit >> VSH) << VSH)) & ((1u << v) - 1u);
            const u32 pos = off + ((rsw >> (16 * (u & 1))) & 0xffffu) + ((exw >> (CB * (u % CPW))) & CMASK) + (u32)__popc(below);
            const i64 j = jd + (i64)u * 32 * V + v;
            if (pos < cap) {                            // beyond the capacity: counted, not written
              if (MODE == 1) {
                typename E::template Regs<1> r1;
                E::template loadv<1, false>(r1, base, inner, j);
                out[pos] = cvt<OutT>(E::template eval<1>(r1, 0, p.c));
              } else {
                out[pos] = (OutT)j;
              }
            }
          }
        }
      } else {
        // many selected elements: every element slot of the tile, predicated stores, no divergence.  Values: the tile is
        // read again, as vectors (it went through L2 four iterations ago) — carrying its values across four iterations
        // would cost the occupancy that hides the loads
        constexpr int UQ = UH >= 2 ? UH / 2 : 1;      // the re-read goes in quarters of the tile: a quarter of the registers
#pragma unroll
        for (int q = 0; q < U / UQ; ++q) {
          R r2[UQ];
          if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < UQ; ++u) E::template loadv<V, UNIT>(r2[u], base, inner, jd + (i64)(q * UQ + u) * 32 * V);
          }
#pragma unroll
          for (int uu = 0; uu < UQ; ++uu) {
            const int u = q * UQ + uu;
            const u32 fu = (sp[((u * V) / 32) * 32 + lane] >> ((u * V) % 32)) & VMASK;
            const u32 exw = sp[(FW + u / CPW) * 32 + lane];
            const u32 rstart = (shared_words[u / 2] >> (16 * (u % 2))) & 0xffffu;
            const u32 lrank = (exw >> (CB * (u % CPW))) & CMASK;     // selected elements of the row in lower lanes
            const i64 j0 = jd + (i64)u * 32 * V;
            if (STAGED) {
              // the row's selected elements meet in the warp's staging buffer, by rank, and leave as full coalesced stores:
              // straight from the registers a store instruction touches every sector of the row's run for a quarter of it
              OutT *stg = (OutT *)stage_all + (size_t)warp * 32 * V;
              u32 lp = lrank;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const bool on = (fu >> v) & 1u;
                const OutT val = MODE == 1 ? cvt<OutT>(E::template eval<V>(r2[uu], v, p.c)) : (OutT)(j0 + v);
                if (on) stg[lp] = val;
                lp += on ? 1u : 0u;
              }
              __syncwarp();
              const u32 rnext = u + 1 < U ? ((shared_words[(u + 1) / 2] >> (16 * ((u + 1) % 2))) & 0xffffu) : cnt;
              const u32 rtot = rnext - rstart;
#pragma unroll
              for (int k = 0; k < V; ++k) {
                const u32 i = (u32)(k * 32 + lane);
                const u32 pos = off + rstart + i;
                if (i < rtot && pos < cap) out[pos] = stg[i];
              }
              __syncwarp();
            } else {
              u32 pos = off + rstart + lrank;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const bool on = (fu >> v) & 1u;
                const OutT val = MODE == 1 ? cvt<OutT>(E::template eval<V>(r2[uu], v, p.c)) : (OutT)(j0 + v);
                st_if<OutT>(on && pos < cap, out + pos, val);
                pos += on ? 1u : 0u;
              }
            }
          }
        }
      }
    }
    // ---- rank this iteration's tile ----
    if (p1) {
      if (fast) {
        flags_fast(1);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const i64 j0 = jl + (i64)u * 32 * V;
          T vals[V];
          u32 fu = 0;
          if (j0 + V <= p.N) {
            R rr;
            E::template loadv<V, UNIT>(rr, base, inner, j0);
#pragma unroll
            for (int v = 0; v < V; ++v) {
              vals[v] = E::template eval<V>(rr, v, p.c);
              fu |= sel_flag<T>(vals[v], thr, fmask) << v;
            }
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              if (j0 + v < p.N) {
                typename E::template Regs<1> r1;
                E::template loadv<1, false>(r1, base, inner, j0 + v);
                vals[v] = E::template eval<1>(r1, 0, p.c);
                fu |= sel_flag<T>(vals[v], thr, fmask) << v;
              }
            }
          }
          if (unique_mode) {
            // unique over a SORTED operand (std::unique / cub::DeviceSelect::Unique): keep x[j] iff j == 0 or x[j] != x[j - 1]
            fu = 0;
            if (j0 < p.N) {
              T prev = vals[0];
              if (j0 > 0) {
                typename E::template Regs<1> r1;
                E::template loadv<1, false>(r1, base, inner, j0 - 1);
                prev = E::template eval<1>(r1, 0, p.c);
              }
#pragma unroll
              for (int v = 0; v < V; ++v) {
                if (j0 + v < p.N) {
                  const bool keep = (j0 + v == 0) || !(vals[v] == prev);
                  fu |= (keep ? 1u : 0u) << v;
                  prev = vals[v];
                }
              }
            }
          }
          f[(u * V) / 32] |= fu << ((u * V) % 32);
        }
      }
      // ranks: packed per-vector counts through shuffle scans; the per-vector totals of the warp come from lane 31
      u32 own_w[NWORD], incl[NWORD];
#pragma unroll
      for (int w = 0; w < NWORD; ++w) own_w[w] = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) own_w[u / CPW] |= (u32)__popc((f[(u * V) / 32] >> ((u * V) % 32)) & VMASK) << (CB * (u % CPW));
#pragma unroll
      for (int w = 0; w < NWORD; ++w) incl[w] = own_w[w];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int w = 0; w < NWORD; ++w) {
          const u32 o = __shfl_up_sync(0xffffffffu, incl[w], d);
          if (lane >= d) incl[w] += o;
        }
      }
      u32 *sp = slot_ptr(s0);
      u32 rs[RSW];
#pragma unroll
      for (int w = 0; w < RSW; ++w) rs[w] = 0;
      u32 total = 0;
#pragma unroll
      for (int w = 0; w < NWORD; ++w) {
        const u32 t = __shfl_sync(0xffffffffu, incl[w], 31);
        sp[(FW + w) * 32 + lane] = incl[w] - own_w[w];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          const int u = w * CPW + c;
          if (u < U) {
            rs[u / 2] |= total << (16 * (u % 2));
            total += (t >> (CB * c)) & CMASK;
          }
        }
      }
#pragma unroll
      for (int w = 0; w < FW; ++w) sp[w * 32 + lane] = f[w];
      if (lane == 0) {
#pragma unroll
        for (int w = 0; w < RSW; ++w) sp[GEO::LANE_WORDS * 32 + w] = rs[w];
        sp[GEO::LANE_WORDS * 32 + RSW] = total;
        ScanSlot<u32>::publish(xc.agg, tile, total, xc.tag);
      }
      __syncwarp();
    }
    s0 = s0 + 1 == SEL_RING ? 0 : s0 + 1;
  }
}

template <class E, class OutT, int V, int MODE>
__device__ __forceinline__ void select1p_body(const EwParams &p) {
  pdl_prologue();
  // per-warp ring of tile states (SelGeom::SLOT_WORDS per tile)
  __shared__ u32 ring[SEL_WARPS * SEL_RING * SelGeom<typename E::value_type, V>::SLOT_WORDS];
  __shared__ __align__(16) unsigned char stage[SEL_WARPS * SelStage<OutT, V>::BYTES];
  const bool unit = p.all_unit != 0;
  if (V > 1 || unit) {
    // V > 1 is only ever launched over unit-stride leaves
    switch (p.sel_op) {
      case 0: select1p_tiles<E, OutT, V, MODE, true, 0>(p, ring, stage); break;
      case 1: select1p_tiles<E, OutT, V, MODE, true, 1>(p, ring, stage); break;
      case 2: select1p_tiles<E, OutT, V, MODE, true, 2>(p, ring, stage); break;
      case 3: select1p_tiles<E, OutT, V, MODE, true, 3>(p, ring, stage); break;
      case 4: select1p_tiles<E, OutT, V, MODE, true, 4>(p, ring, stage); break;
      case 5: select1p_tiles<E, OutT, V, MODE, true, 5>(p, ring, stage); break;
      default: select1p_tiles<E, OutT, V, MODE, true, -1>(p, ring, stage); break;
    }
  } else {
    select1p_tiles<E, OutT, V, MODE, false, -1>(p, ring, stage);
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // exit ticket: the last CTA out opens the next epoch (every status word of this launch is stale from then on)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(p.sel_ticket, gridDim.x - 1) == gridDim.x - 1) {
      const u32 epoch = __ldcg(p.sel_epoch) & 0x3fffffffu;
      const u32 e = (epoch + 1u) & 0x3fffffffu;
      *(volatile u32 *)p.sel_epoch = e ? e : 1u;   // epoch 0 is what freshly zeroed status words carry: never used
    }
  }
}

// ------------------------------------------------------------------------------------------------
// H1: hist — even-width histogram of every row (reference: hist_impl -> cub::DeviceHistogram::HistogramEven,
// transforms/cub.h:320-359,2464-2503; bin arithmetic restated from CCCL's ScaleTransform, cub/device/dispatch/kernels/
// kernel_histogram.cuh:75-245: a sample counts iff lower <= x < upper, bin = int((x - lower) * scale) with scale =
// T(bins) / T(upper - lower) for floating types and ((x - lower) * bins) / (upper - lower) in 64-bit unsigned arithmetic
// for integers).  A work item is a chunk of one row; a CTA keeps the item's bins in shared memory (atomicAdd.shared),
// then adds the non-empty ones to the row's output bins (integer atomics: the result does not depend on the order).  The
// host zeroes the output on the stream first.
// ------------------------------------------------------------------------------------------------
template <class T> struct HistBin {   // floating types
  T lo, hi, scale;
  int bins;
  __device__ __forceinline__ void init(const RedParams &p) { lo = (T)p.hist_lo_d; hi = (T)p.hist_hi_d; bins = p.hist_bins; scale = (T)((T)p.hist_bins / (T)(hi - lo)); }
  // a sample a hair below `upper` can round to bin == bins in (x - lower) * scale (999.99994f * 0.256f = 256.0f): CUB's
  // formula would index past its last bin there; such a sample is dropped (the oracle does the same)
  __device__ __forceinline__ int bin(T x) const {
    if (!(x >= lo && x < hi)) return -1;
    const int b = (int)((x - lo) * scale);
    return b < bins ? b : -1;
  }
};
template <class T> struct HistBinInt {
  i64 lo, hi;
  u64 bins, range;
  __device__ __forceinline__ void init(const RedParams &p) { lo = p.hist_lo_i; hi = p.hist_hi_i; bins = (u64)p.hist_bins; range = (u64)(hi - lo); }
  __device__ __forceinline__ int bin(T x) const { const i64 v = (i64)x; return (v >= lo && v < hi) ? (int)(((u64)(v - lo) * bins) / range) : -1; }
};
template <> struct HistBin<int> : HistBinInt<int> {};
template <> struct HistBin<i64> : HistBinInt<i64> {};
template <> struct HistBin<unsigned char> : HistBinInt<unsigned char> {};

template <class E, int V, int U, bool UNIT>
__device__ __forceinline__ void hist_body_impl(const RedParams &p) {
  typedef typename E::value_type T;
  extern __shared__ __align__(16) unsigned char hist_smem[];
  int *s_bins = (int *)hist_smem;
  const int nthr = blockDim.x, tid = threadIdx.x;
  const int bins = p.hist_bins;
  const bool priv = p.hist_smem != 0;
  HistBin<T> hb;
  hb.init(p);
  const i64 L = p.rsz[0], Lv = L / V, tail = L - Lv * V;
  const i64 S = p.splits;                      // chunks per row
  const i64 per = (Lv + S - 1) / S;            // vector steps per chunk
  const i64 work = p.B * S;
  for (i64 w = blockIdx.x; w < work; w += gridDim.x) {
    const i64 b = w / S, s = w - b * S;
    const char *base[E::NL];
    i64 inner[E::NL];
    i64 bidx[KMAXD];
    decomp(b, p.nb, p.bsz, bidx);
    i64 oo = 0;
#pragma unroll
    for (int k = 0; k < E::NL; ++k) {
      i64 off = 0;
#pragma unroll
      for (int d = 0; d < KMAXD; ++d) if (d < p.nb) off += bidx[d] * p.leaf[k].bs[d];
      base[k] = (const char *)p.leaf[k].ptr + off * E::leaf_bytes(k);
      inner[k] = p.leaf[k].rs[0];
    }
#pragma unroll
    for (int d = 0; d < KMAXD; ++d) if (d < p.nb) oo += bidx[d] * p.out.bs[d];
    int *obins = (int *)p.out.ptr + oo;          // the row's bins are contiguous (host rule)
    int *dst = priv ? s_bins : obins;
    if (priv) {
      for (int i = tid; i < bins; i += nthr) s_bins[i] = 0;
      __syncthreads();
    }
    const i64 q0 = s * per, q1 = (q0 + per < Lv) ? q0 + per : Lv;
    i64 q = q0 + tid;
    for (; q + (i64)(U - 1) * nthr < q1; q += (i64)U * nthr) {
      typename E::template Regs<V> r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) E::template loadv<V, UNIT>(r[u], base, inner, (q + (i64)u * nthr) * V);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int bi = hb.bin(E::template eval<V>(r[u], v, p.c));
          if (bi >= 0) atomicAdd(dst + bi, 1);
        }
      }
    }
    for (; q < q1; q += nthr) {
      typename E::template Regs<V> r;
      E::template loadv<V, UNIT>(r, base, inner, q * V);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int bi = hb.bin(E::template eval<V>(r, v, p.c));
        if (bi >= 0) atomicAdd(dst + bi, 1);
      }
    }
    if (V > 1 && tail > 0 && s == S - 1) {
      for (i64 j = Lv * V + tid; j < L; j += nthr) {
        typename E::template Regs<1> r;
        E::template loadv<1, false>(r, base, inner, j);
        const int bi = hb.bin(E::template eval<1>(r, 0, p.c));
        if (bi >= 0) atomicAdd(dst + bi, 1);
      }
    }
    if (priv) {
      __syncthreads();
      for (int i = tid; i < bins; i += nthr) {
        const int c = s_bins[i];
        if (c) atomicAdd(obins + i, c);
      }
      __syncthreads();
    }
  }
}
template <class E, int V, int U>
__device__ __forceinline__ void hist_body(const RedParams &p) {
  pdl_prologue();
  if (p.all_unit) hist_body_impl<E, V, U, true>(p);
  else hist_body_impl<E, V, U, false>(p);
}

}  // namespace mxb
