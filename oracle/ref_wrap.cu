// ref_wrap.cu — thin extern "C" wrappers around STATEMENTS OF THE UNMODIFIED REFERENCE (MatX), so that the
// Python harness can run the reference's own implementation of the path on raw buffers.
//
// TEST / BASELINE INFRASTRUCTURE ONLY: built by oracle/build_ref.py from the reference headers where they lie
// (/root/reference/include — never copied into this repo) into oracle/_ref/, which is git-ignored.  Used to
// (1) pin oracle/matx_oracle.c, (2) generate tests/golden/, (3) time the reference's HostExecutor as the CPU
// baseline, (4) with -DMREF_CUDA, run the reference's cudaExecutor + CUB path on the GPU box for A/B numbers.
// Every wrapper is one MatX statement of the form the north star names, e.g. `(out = sum(a*b+c, {1})).run(exec)`.
//
// Build-time selection: -DMREF_DTYPE=<0..4> picks the element type of the generic reduce entry point of this
// translation unit (the instantiation sets are kept per-TU so the build parallelises); -DMREF_FUSED adds the
// fused-expression entry points; -DMREF_CUDA switches the executor to matx::cudaExecutor.
#include <matx.h>

#include <cstdint>

using namespace matx;

#ifdef MREF_CUDA
#define MREF_SUFFIX _cuda
static constexpr bool kCuda = true;   // the reference's CUB path does not compile any()/all() of a complex tensor
#else
#define MREF_SUFFIX _host
static constexpr bool kCuda = false;
#endif

#define MREF_CAT2(a, b) a##b
#define MREF_CAT(a, b) MREF_CAT2(a, b)
#define MREF_NAME(base) MREF_CAT(base, MREF_SUFFIX)

// run `body(ex)` with the executor `mode` selects: 0 = HostExecutor<SINGLE>, 1 = HostExecutor<ALL>
template <class F> static int with_exec(int mode, F &&body) {
  try {
#ifdef MREF_CUDA
    (void)mode;
    cudaExecutor ex{(cudaStream_t)0};
    body(ex);
    ex.sync();
#else
    if (mode == 0) { HostExecutor<ThreadsMode::SINGLE> ex; body(ex); }
    else { HostExecutor<ThreadsMode::ALL> ex; body(ex); }
#endif
  } catch (...) {
    return 1;
  }
  return 0;
}

template <typename T, int RANK> static auto make_view(void *p, const int64_t *shape, const int64_t *strides) {
  index_t sh[RANK], st[RANK];
  for (int i = 0; i < RANK; ++i) { sh[i] = shape[i]; st[i] = strides[i]; }
  return make_tensor<T>(reinterpret_cast<T *>(p), sh, st);
}
template <typename T, int RANK> static auto make_out(void *p, const int64_t *shape, const int *dims, int nd, int in_rank) {
  if constexpr (RANK == 0) {
    (void)shape; (void)dims; (void)nd; (void)in_rank;
    return make_tensor<T>(reinterpret_cast<T *>(p), {});
  } else {
    cuda::std::array<index_t, RANK> sh;
    int m = 0;
    for (int d = 0; d < in_rank; ++d) {
      bool red = false;
      for (int j = 0; j < nd; ++j) red = red || dims[j] == d;
      if (!red) sh[m++] = shape[d];
    }
    return make_tensor<T>(reinterpret_cast<T *>(p), sh);
  }
}

enum { R_SUM = 0, R_MEAN, R_VAR, R_STDD, R_MAX, R_MIN, R_ARGMAX, R_ARGMIN, R_ANY, R_ALL, R_PROD };

template <typename T> struct real_of { using type = T; };
template <> struct real_of<cuda::std::complex<float>> { using type = float; };

// one reduction statement; D == RANK uses the full-reduction spelling `sum(x)`
template <typename T, int RANK, int D, class Exec>
static void reduce_stmt(Exec &ex, int op, void *in, const int64_t *shape, const int64_t *strides, const int *dims_rt,
                        void *out, int64_t *idx_out, int ddof) {
  auto v = make_view<T, RANK>(in, shape, strides);
  constexpr int OR = RANK - D;
  using RT = typename real_of<T>::type;
  constexpr bool cplx = !std::is_same_v<T, RT>;
  // the reference's host mean / prod / var do not compile for 16-bit floats (reduce.h:326,747): not built
  constexpr bool half = std::is_same_v<T, matxBf16> || std::is_same_v<T, matxFp16>;
  auto o = make_out<T, OR>(out, shape, dims_rt, D, RANK);
  [[maybe_unused]] auto orl = make_out<RT, OR>(out, shape, dims_rt, D, RANK);
  [[maybe_unused]] auto oi = make_out<index_t, OR>(idx_out ? (void *)idx_out : out, shape, dims_rt, D, RANK);
  if constexpr (D == RANK) {
    switch (op) {
      case R_SUM: (o = sum(v)).run(ex); break;
      case R_MEAN: if constexpr (!half) { (o = mean(v)).run(ex); } break;
      case R_PROD: if constexpr (!half) { (o = prod(v)).run(ex); } break;
      case R_ANY: if constexpr (!(cplx && kCuda)) { (o = any(v)).run(ex); } break;
      case R_ALL: if constexpr (!(cplx && kCuda)) { (o = all(v)).run(ex); } break;
      case R_VAR: if constexpr (!std::is_integral_v<T> && !half) { (orl = var(v, ddof)).run(ex); } break;
      case R_STDD: if constexpr (!std::is_integral_v<T> && !half) { (orl = stdd(v, ddof)).run(ex); } break;
      default:
        if constexpr (!cplx) {
          switch (op) {
            case R_MAX: (o = max(v)).run(ex); break;
            case R_MIN: (o = min(v)).run(ex); break;
            case R_ARGMAX: (mtie(o, oi) = argmax(v)).run(ex); break;
            case R_ARGMIN: (mtie(o, oi) = argmin(v)).run(ex); break;
          }
        }
    }
  } else {
    int dims[D];
    for (int i = 0; i < D; ++i) dims[i] = dims_rt[i];
    switch (op) {
      case R_SUM: (o = sum(v, dims)).run(ex); break;
      case R_MEAN: if constexpr (!half) { (o = mean(v, dims)).run(ex); } break;
      case R_PROD: if constexpr (!half) { (o = prod(v, dims)).run(ex); } break;
      case R_ANY: if constexpr (!(cplx && kCuda)) { (o = any(v, dims)).run(ex); } break;
      case R_ALL: if constexpr (!(cplx && kCuda)) { (o = all(v, dims)).run(ex); } break;
      case R_VAR: if constexpr (!std::is_integral_v<T> && !half) { (orl = var(v, dims, ddof)).run(ex); } break;
      case R_STDD: if constexpr (!std::is_integral_v<T> && !half) { (orl = stdd(v, dims, ddof)).run(ex); } break;
      default:
        if constexpr (!cplx) {
          switch (op) {
            case R_MAX: (o = max(v, dims)).run(ex); break;
            case R_MIN: (o = min(v, dims)).run(ex); break;
            case R_ARGMAX: (mtie(o, oi) = argmax(v, dims)).run(ex); break;
            case R_ARGMIN: (mtie(o, oi) = argmin(v, dims)).run(ex); break;
          }
        }
    }
  }
}

#ifdef MREF_DTYPE
#if MREF_DTYPE == 0
using elem_t = float;
#define MREF_MAXRANK 4
#elif MREF_DTYPE == 1
using elem_t = double;
#define MREF_MAXRANK 2
#elif MREF_DTYPE == 2
using elem_t = matxBf16;
#define MREF_MAXRANK 3
#elif MREF_DTYPE == 4
using elem_t = cuda::std::complex<float>;
#define MREF_MAXRANK 2
#elif MREF_DTYPE == 5
using elem_t = int32_t;
#define MREF_MAXRANK 2
#endif

template <int RANK, int D>
static int reduce_rd(int mode, int op, void *in, const int64_t *shape, const int64_t *strides, const int *dims, void *out,
                     int64_t *idx_out, int ddof) {
  return with_exec(mode, [&](auto &ex) { reduce_stmt<elem_t, RANK, D>(ex, op, in, shape, strides, dims, out, idx_out, ddof); });
}

#define MREF_FN3(a, b) a##b
#define MREF_FN2(a, b) MREF_FN3(a, b)
#define MREF_REDUCE_NAME MREF_NAME(MREF_FN2(MREF_FN2(mref_reduce_dt, MREF_DTYPE), _))

// reduce dims `dims[0..nd)` of the strided view (shape, strides in elements); `out` (and `idx_out`) are
// contiguous over the remaining dims in their original order.  Returns 0, 1 (reference threw) or 2 (not built).
extern "C" int MREF_REDUCE_NAME(int mode, int op, int rank, const int64_t *shape, const int64_t *strides, void *in, int nd,
                                const int *dims, void *out, int64_t *idx_out, int ddof) {
#define MREF_CASE(R, D) if (rank == R && nd == D) return reduce_rd<R, D>(mode, op, in, shape, strides, dims, out, idx_out, ddof);
  MREF_CASE(1, 1)
#if MREF_MAXRANK >= 2
  MREF_CASE(2, 1) MREF_CASE(2, 2)
#endif
#if MREF_MAXRANK >= 3
  MREF_CASE(3, 1) MREF_CASE(3, 2)
#if MREF_DTYPE == 0
  MREF_CASE(3, 3)
#endif
#endif
#if MREF_MAXRANK >= 4
  MREF_CASE(4, 1) MREF_CASE(4, 2) MREF_CASE(4, 3) MREF_CASE(4, 4)
#endif
  return 2;
}
#endif  // MREF_DTYPE

#ifdef MREF_FUSED
// config 1: (out = sum(a*b+c, {1})).run(exec)
extern "C" int MREF_NAME(mref_fma_sum)(int mode, const float *a, const float *b, const float *c, float *out, int64_t rows, int64_t cols) {
  return with_exec(mode, [&](auto &ex) {
    auto ta = make_tensor<float>(const_cast<float *>(a), {rows, cols});
    auto tb = make_tensor<float>(const_cast<float *>(b), {rows, cols});
    auto tc = make_tensor<float>(const_cast<float *>(c), {rows, cols});
    auto to = make_tensor<float>(out, {rows});
    (to = sum(ta * tb + tc, {1})).run(ex);
  });
}
// config 3: (mtie(v, i) = argmax(abs2(x), {1})).run(exec) on complex<float>
extern "C" int MREF_NAME(mref_abs2_argmax)(int mode, const void *x, float *val, int64_t *idx, int64_t rows, int64_t cols) {
  using cf = cuda::std::complex<float>;
  return with_exec(mode, [&](auto &ex) {
    auto tx = make_tensor<cf>(reinterpret_cast<cf *>(const_cast<void *>(x)), {rows, cols});
    auto tv = make_tensor<float>(val, {rows});
    auto ti = make_tensor<index_t>(reinterpret_cast<index_t *>(idx), {rows});
    (mtie(tv, ti) = argmax(abs2(tx), {1})).run(ex);
  });
}
// config 4: the arithmetic-expression form of examples/black_scholes.cu:122-138, stated with the same tokens
extern "C" int MREF_NAME(mref_black_scholes)(int mode, const float *pK, const float *pS, const float *pV, const float *pr,
                                            const float *pT, float *pout, int64_t n) {
  return with_exec(mode, [&](auto &ex) {
    auto K = make_tensor<float>(const_cast<float *>(pK), {n});
    auto S = make_tensor<float>(const_cast<float *>(pS), {n});
    auto V = make_tensor<float>(const_cast<float *>(pV), {n});
    auto r = make_tensor<float>(const_cast<float *>(pr), {n});
    auto T = make_tensor<float>(const_cast<float *>(pT), {n});
    auto output = make_tensor<float>(pout, {n});
    auto VsqrtT = V * sqrt(T);
    auto d1 = (log(S / K) + (r + 0.5f * V * V) * T) / VsqrtT;
    auto d2 = d1 - VsqrtT;
    auto cdf_d1 = normcdf(d1);
    auto cdf_d2 = normcdf(d2);
    auto expRT = exp(-1.f * r * T);
    (output = S * cdf_d1 - K * expRT * cdf_d2).run(ex);
  });
}
// bench/00_operators/operators.cu:10-36 — vector add
extern "C" int MREF_NAME(mref_vector_add)(int mode, const float *a, const float *b, float *out, int64_t n) {
  return with_exec(mode, [&](auto &ex) {
    auto ta = make_tensor<float>(const_cast<float *>(a), {n});
    auto tb = make_tensor<float>(const_cast<float *>(b), {n});
    auto to = make_tensor<float>(out, {n});
    (to = ta + tb).run(ex);
  });
}
// element functors on fp32 (operator_func_*_test.cu style): out = f(a) / f(a, b)
extern "C" int MREF_NAME(mref_unary_f32)(int mode, int opcode, const float *a, float *out, int64_t n) {
  return with_exec(mode, [&](auto &ex) {
    auto ta = make_tensor<float>(const_cast<float *>(a), {n});
    auto to = make_tensor<float>(out, {n});
    switch (opcode) {
      case 40: (to = -ta).run(ex); break;
      case 41: (to = sqrt(ta)).run(ex); break;
      case 42: (to = rsqrt(ta)).run(ex); break;
      case 43: (to = exp(ta)).run(ex); break;
      case 44: (to = log(ta)).run(ex); break;
      case 45: (to = log2(ta)).run(ex); break;
      case 46: (to = log10(ta)).run(ex); break;
      case 47: (to = abs(ta)).run(ex); break;
      case 48: (to = abs2(ta)).run(ex); break;
      case 52: (to = sin(ta)).run(ex); break;
      case 53: (to = cos(ta)).run(ex); break;
      case 54: (to = tan(ta)).run(ex); break;
      case 55: (to = tanh(ta)).run(ex); break;
      case 56: (to = normcdf(ta)).run(ex); break;
      case 60: (to = floor(ta)).run(ex); break;
      case 61: (to = ceil(ta)).run(ex); break;
      case 62: (to = round(ta)).run(ex); break;
      case 63: (to = sinh(ta)).run(ex); break;
      case 64: (to = cosh(ta)).run(ex); break;
      case 65: (to = asin(ta)).run(ex); break;
      case 66: (to = acos(ta)).run(ex); break;
      case 67: (to = atan(ta)).run(ex); break;
      default: throw 1;
    }
  });
}
extern "C" int MREF_NAME(mref_binary_f32)(int mode, int opcode, const float *a, const float *b, float *out, int64_t n) {
  return with_exec(mode, [&](auto &ex) {
    auto ta = make_tensor<float>(const_cast<float *>(a), {n});
    auto tb = make_tensor<float>(const_cast<float *>(b), {n});
    auto to = make_tensor<float>(out, {n});
    switch (opcode) {
      case 10: (to = ta + tb).run(ex); break;
      case 11: (to = ta - tb).run(ex); break;
      case 12: (to = ta * tb).run(ex); break;
      case 13: (to = ta / tb).run(ex); break;
      case 14: (to = fmod(ta, tb)).run(ex); break;
      case 15: (to = pow(ta, tb)).run(ex); break;
      case 16: (to = matx::max(ta, tb)).run(ex); break;
      case 17: (to = matx::min(ta, tb)).run(ex); break;
      default: throw 1;
    }
  });
}
// test/00_operators/ReductionTests.cu:1615-1693 — (mtie(out, num_found) = find / find_idx(t, SEL{thresh})).run(exec)
// on a (possibly strided) rank-1 or rank-2 fp32 view; sel: 0 LT, 1 GT, 2 EQ, 3 NEQ, 4 LTE, 5 GTE
template <int RANK, class Ex> static void ref_find_rank(Ex &ex, int sel, float thr, float *in, const int64_t *shape, const int64_t *strides,
                                                         void *out, int64_t cap, int *num_found, int want_idx) {
  auto t = make_view<float, RANK>(in, shape, strides);
  auto nf = make_tensor<int>(num_found, {});
  auto run = [&](auto functor) {
    if (want_idx) {
      auto o = make_tensor<int>(reinterpret_cast<int *>(out), {cap});
      (mtie(o, nf) = find_idx(t, functor)).run(ex);
    } else {
      auto o = make_tensor<float>(reinterpret_cast<float *>(out), {cap});
      (mtie(o, nf) = find(t, functor)).run(ex);
    }
  };
  switch (sel) {
    case 0: run(LT<float>{thr}); break;
    case 1: run(GT<float>{thr}); break;
    case 2: run(EQ<float>{thr}); break;
    case 3: run(NEQ<float>{thr}); break;
    case 4: run(LTE<float>{thr}); break;
    case 5: run(GTE<float>{thr}); break;
    default: throw 1;
  }
}
extern "C" int MREF_NAME(mref_find_f32)(int mode, int sel, float thr, int rank, const int64_t *shape, const int64_t *strides, float *in,
                                        void *out, int64_t cap, int *num_found, int want_idx) {
  return with_exec(mode, [&](auto &ex) {
    if (rank == 1) ref_find_rank<1>(ex, sel, thr, in, shape, strides, out, cap, num_found, want_idx);
    else if (rank == 2) ref_find_rank<2>(ex, sel, thr, in, shape, strides, out, cap, num_found, want_idx);
    else throw 1;
  });
}
extern "C" int MREF_NAME(mref_threads)(void) {
#if defined(MATX_EN_OMP) && !defined(MREF_CUDA)
  return omp_get_num_procs();
#else
  return 1;
#endif
}
#endif  // MREF_FUSED

#ifdef MREF_SORT
// ---- sort / unique statements of the reference (HostExecutor overloads: transforms/cub.h:2192-2240,2844-2880) -----------
template <class T>
static int ref_sort_rows(int mode, int desc, T *in, T *out, int64_t rows, int64_t cols) {
  return with_exec(mode, [&](auto &ex) {
    if (rows == 1) {
      auto t = make_tensor<T>(in, {cols});
      auto o = make_tensor<T>(out, {cols});
      (o = matx::sort(t, desc ? SORT_DIR_DESC : SORT_DIR_ASC)).run(ex);
    } else {
      auto t = make_tensor<T>(in, {rows, cols});
      auto o = make_tensor<T>(out, {rows, cols});
      (o = matx::sort(t, desc ? SORT_DIR_DESC : SORT_DIR_ASC)).run(ex);
    }
  });
}
extern "C" int MREF_NAME(mref_sort_f32)(int mode, int desc, float *in, float *out, int64_t rows, int64_t cols) { return ref_sort_rows<float>(mode, desc, in, out, rows, cols); }
extern "C" int MREF_NAME(mref_sort_f64)(int mode, int desc, double *in, double *out, int64_t rows, int64_t cols) { return ref_sort_rows<double>(mode, desc, in, out, rows, cols); }
extern "C" int MREF_NAME(mref_sort_i32)(int mode, int desc, int *in, int *out, int64_t rows, int64_t cols) { return ref_sort_rows<int>(mode, desc, in, out, rows, cols); }
template <class T>
static int ref_unique(int mode, T *in, int64_t n, T *out, int *num_found) {
  return with_exec(mode, [&](auto &ex) {
    auto t = make_tensor<T>(in, {n});
    auto o = make_tensor<T>(out, {n});
    auto nf = make_tensor<int>(num_found, {});
    (mtie(o, nf) = unique(t)).run(ex);
  });
}
extern "C" int MREF_NAME(mref_unique_f32)(int mode, float *in, int64_t n, float *out, int *num_found) { return ref_unique<float>(mode, in, n, out, num_found); }
extern "C" int MREF_NAME(mref_unique_i32)(int mode, int *in, int64_t n, int *out, int *num_found) { return ref_unique<int>(mode, in, n, out, num_found); }
#endif  // MREF_SORT
