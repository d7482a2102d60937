// dropin_test.cu — the drop-in proof: UNMODIFIED MatX statements run twice, once with the reference's
// matx::cudaExecutor (CUB path) and once with matx::b200Executor (libmatx_b200.so), on the same inputs, and the
// results are compared.  Also prints which native kernel served each statement.  Built here against the reference
// headers by oracle/build_ref.py --dropin (into oracle/_ref/dropin_test), executed on the GPU box by
// tests/test_gpu_dropin.py.  With `--bench` it times both executors on the BASELINE.json configs (A/B on one box).
#include <matx.h>
#include <matx_b200/executor.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace matx;
using cf = cuda::std::complex<float>;

static int g_fail = 0, g_pass = 0;
static void report(const char *name, bool ok, const char *kernel, double err = 0.0) {
  printf("%s %-44s err=%.3g kernel=%s\n", ok ? "PASS" : "FAIL", name, err, kernel && *kernel ? kernel : "(reference fallback)");
  (ok ? g_pass : g_fail)++;
}
template <class T> static double max_rel(const T &a, const T &b, index_t n) {
  double m = 0;
  for (index_t i = 0; i < n; ++i) {
    const double x = static_cast<double>(a(i)), y = static_cast<double>(b(i));
    m = std::max(m, std::fabs(x - y) / std::max(std::fabs(y), 1e-30));
  }
  return m;
}

static void fill_uniform(tensor_t<float, 2> &t, unsigned seed, float lo, float hi) {
  std::mt19937 g(seed);
  std::uniform_real_distribution<float> d(lo, hi);
  for (index_t i = 0; i < t.Size(0); ++i) for (index_t j = 0; j < t.Size(1); ++j) t(i, j) = d(g);
}

static int run_checks() {
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  cudaExecutor ref{stream};
  b200Executor b200{stream};

  // ---- config 1: (out = sum(a*b+c, {1})).run(exec) ----
  {
    const index_t rows = 64, cols = 4096;
    auto a = make_tensor<float>({rows, cols}), b = make_tensor<float>({rows, cols}), c = make_tensor<float>({rows, cols});
    fill_uniform(a, 1, 0.f, 1.f); fill_uniform(b, 2, 0.f, 1.f); fill_uniform(c, 3, -0.5f, 0.5f);
    auto o1 = make_tensor<float>({rows}), o2 = make_tensor<float>({rows});
    (o1 = sum(a * b + c, {1})).run(ref);
    (o2 = sum(a * b + c, {1})).run(b200);
    ref.sync();
    const double e = max_rel(o2, o1, rows);
    report("sum(a*b+c,{1}) fp32 64x4096", e <= 1e-5 && *b200.last_kernel(), b200.last_kernel(), e);

    // mean / var / stdd / max / min / any / all over the same rows
    auto v1 = make_tensor<float>({rows}), v2 = make_tensor<float>({rows});
    (v1 = mean(a, {1})).run(ref); (v2 = mean(a, {1})).run(b200); ref.sync();
    report("mean(a,{1})", max_rel(v2, v1, rows) <= 1e-5 && *b200.last_kernel(), b200.last_kernel(), max_rel(v2, v1, rows));
    (v1 = var(a, {1}, 1)).run(ref); (v2 = var(a, {1}, 1)).run(b200); ref.sync();
    report("var(a,{1})", max_rel(v2, v1, rows) <= 2e-5 && *b200.last_kernel(), b200.last_kernel(), max_rel(v2, v1, rows));
    (v1 = stdd(a, {1}, 1)).run(ref); (v2 = stdd(a, {1}, 1)).run(b200); ref.sync();
    report("stdd(a,{1})", max_rel(v2, v1, rows) <= 2e-5 && *b200.last_kernel(), b200.last_kernel(), max_rel(v2, v1, rows));
    (v1 = max(a, {1})).run(ref); (v2 = max(a, {1})).run(b200); ref.sync();
    report("max(a,{1}) bit-exact", max_rel(v2, v1, rows) == 0 && *b200.last_kernel(), b200.last_kernel());
    (v1 = min(a, {1})).run(ref); (v2 = min(a, {1})).run(b200); ref.sync();
    report("min(a,{1}) bit-exact", max_rel(v2, v1, rows) == 0 && *b200.last_kernel(), b200.last_kernel());
    // permuted view: reduce the OUTER dim (strided), PermutedReduce of ReductionTests.cu:353-573
    auto w1 = make_tensor<float>({cols}), w2 = make_tensor<float>({cols});
    (w1 = sum(a, {0})).run(ref); (w2 = sum(a, {0})).run(b200); ref.sync();
    report("sum(a,{0}) strided reduce dim", max_rel(w2, w1, cols) <= 1e-5 && *b200.last_kernel(), b200.last_kernel(), max_rel(w2, w1, cols));
    (w1 = sum(permute(a, {1, 0}), {1})).run(ref); (w2 = sum(permute(a, {1, 0}), {1})).run(b200); ref.sync();
    report("sum(permute(a,{1,0}),{1})", max_rel(w2, w1, cols) <= 1e-5 && *b200.last_kernel(), b200.last_kernel(), max_rel(w2, w1, cols));
    // full reductions to a rank-0 tensor
    auto s1 = make_tensor<float>({}), s2 = make_tensor<float>({});
    (s1 = sum(a)).run(ref); (s2 = sum(a)).run(b200); ref.sync();
    report("sum(a) full", std::fabs(s2() - s1()) <= 1e-5 * std::fabs(s1()) && *b200.last_kernel(), b200.last_kernel(), std::fabs(s2() - s1()) / s1());
    auto i1 = make_tensor<index_t>({}), i2 = make_tensor<index_t>({});
    (mtie(s1, i1) = argmax(a)).run(ref); (mtie(s2, i2) = argmax(a)).run(b200); ref.sync();
    report("mtie(v,i)=argmax(a) full", s1() == s2() && i1() == i2() && *b200.last_kernel(), b200.last_kernel());
    auto ir1 = make_tensor<index_t>({rows}), ir2 = make_tensor<index_t>({rows});
    (mtie(v1, ir1) = argmin(a, {1})).run(ref); (mtie(v2, ir2) = argmin(a, {1})).run(b200); ref.sync();
    bool same = true;
    for (index_t r = 0; r < rows; ++r) same = same && v1(r) == v2(r) && ir1(r) == ir2(r);
    report("mtie(v,i)=argmin(a,{1}) absolute index", same && *b200.last_kernel(), b200.last_kernel());
    {
      auto mn1 = make_tensor<float>({rows}), mn2 = make_tensor<float>({rows}), mx1 = make_tensor<float>({rows}), mx2 = make_tensor<float>({rows});
      auto in1 = make_tensor<index_t>({rows}), in2 = make_tensor<index_t>({rows}), ix1 = make_tensor<index_t>({rows}), ix2 = make_tensor<index_t>({rows});
      (mtie(mn1, in1, mx1, ix1) = argminmax(a, {1})).run(ref);
      (mtie(mn2, in2, mx2, ix2) = argminmax(a, {1})).run(b200);
      ref.sync();
      bool ok = true;
      for (index_t r = 0; r < rows; ++r) ok = ok && mn1(r) == mn2(r) && mx1(r) == mx2(r) && in1(r) == in2(r) && ix1(r) == ix2(r);
      report("mtie(mn,imn,mx,imx)=argminmax(a,{1})", ok && *b200.last_kernel(), b200.last_kernel());
    }
    (v1 = any(a > 0.9999f, {1})).run(ref); (v2 = any(a > 0.9999f, {1})).run(b200); ref.sync();
    report("any(a>0.9999f,{1})", max_rel(v2, v1, rows) == 0 && *b200.last_kernel(), b200.last_kernel());
    // elementwise with broadcast scalar, unary chain
    auto e1 = make_tensor<float>({rows, cols}), e2 = make_tensor<float>({rows, cols});
    (e1 = sqrt(a) * 2.f + exp(-b) / (c * c + 1.f)).run(ref);
    (e2 = sqrt(a) * 2.f + exp(-b) / (c * c + 1.f)).run(b200);
    ref.sync();
    double em = 0;
    for (index_t i = 0; i < rows; ++i) for (index_t j = 0; j < cols; j += 17) em = std::max(em, std::fabs((double)e1(i, j) - e2(i, j)) / std::fabs(e1(i, j)));
    report("elementwise sqrt/exp/div chain", em <= 1e-6 && *b200.last_kernel(), b200.last_kernel(), em);
  }
  // ---- config 3: complex rows ----
  {
    const index_t rows = 32, cols = 8192;
    auto x = make_tensor<cf>({rows, cols});
    std::mt19937 g(5);
    std::normal_distribution<float> d(0.f, 1.f);
    for (index_t i = 0; i < rows; ++i) for (index_t j = 0; j < cols; ++j) x(i, j) = cf(d(g), d(g));
    auto v1 = make_tensor<float>({rows}), v2 = make_tensor<float>({rows});
    auto i1 = make_tensor<index_t>({rows}), i2 = make_tensor<index_t>({rows});
    (mtie(v1, i1) = argmax(abs2(x), {1})).run(ref); (mtie(v2, i2) = argmax(abs2(x), {1})).run(b200); ref.sync();
    bool same = true;
    for (index_t r = 0; r < rows; ++r) same = same && i1(r) == i2(r);
    report("argmax(abs2(x),{1}) complex<float>", same && max_rel(v2, v1, rows) <= 1e-6 && *b200.last_kernel(), b200.last_kernel(), max_rel(v2, v1, rows));
    (v1 = var(x, {1}, 1)).run(ref); (v2 = var(x, {1}, 1)).run(b200); ref.sync();
    report("var(x,{1}) complex<float> -> float", max_rel(v2, v1, rows) <= 2e-5 && *b200.last_kernel(), b200.last_kernel(), max_rel(v2, v1, rows));
    auto m1 = make_tensor<cf>({rows}), m2 = make_tensor<cf>({rows});
    (m1 = mean(x, {1})).run(ref); (m2 = mean(x, {1})).run(b200); ref.sync();
    double em = 0;
    for (index_t r = 0; r < rows; ++r) em = std::max(em, (double)cuda::std::abs(m1(r) - m2(r)));
    report("mean(x,{1}) complex<float>", em <= 1e-5 && *b200.last_kernel(), b200.last_kernel(), em);
  }
  // ---- config 4: Black-Scholes expression of examples/black_scholes.cu:122-138 ----
  {
    const index_t n = 1 << 16;
    auto K = make_tensor<float>({n}), S = make_tensor<float>({n}), V = make_tensor<float>({n}), r = make_tensor<float>({n}), T = make_tensor<float>({n});
    std::mt19937 g(6);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    for (index_t i = 0; i < n; ++i) { S(i) = 10 + 90 * u(g); K(i) = 10 + 90 * u(g); V(i) = 0.05f + 0.45f * u(g); r(i) = 0.01f + 0.09f * u(g); T(i) = 0.1f + 1.9f * u(g); }
    auto o1 = make_tensor<float>({n}), o2 = make_tensor<float>({n});
    auto bs = [&](auto &out, auto &exec) {
      auto VsqrtT = V * sqrt(T);
      auto d1 = (log(S / K) + (r + 0.5f * V * V) * T) / VsqrtT;
      auto d2 = d1 - VsqrtT;
      auto cdf_d1 = normcdf(d1);
      auto cdf_d2 = normcdf(d2);
      auto expRT = exp(-1.f * r * T);
      (out = S * cdf_d1 - K * expRT * cdf_d2).run(exec);
    };
    bs(o1, ref); bs(o2, b200); ref.sync();
    double em = 0;
    for (index_t i = 0; i < n; ++i) em = std::max(em, std::fabs((double)o1(i) - o2(i)));
    report("black_scholes expression (abs err)", em <= 1e-4 && !strncmp(b200.last_kernel(), "ew|", 3), b200.last_kernel(), em);
  }
  // ---- config 5: bf16, reduce over a permuted non-contiguous dim ----
  {
    const index_t d0 = 16, d1 = 96, d2 = 256;
    auto t = make_tensor<matxBf16>({d0, d1, d2});
    std::mt19937 g(11);
    std::uniform_real_distribution<float> u(0.f, 0.25f);
    for (index_t i = 0; i < d0; ++i) for (index_t j = 0; j < d1; ++j) for (index_t k = 0; k < d2; ++k) t(i, j, k) = matxBf16(u(g));
    auto o1 = make_tensor<matxBf16>({d2, d0}), o2 = make_tensor<matxBf16>({d2, d0});
    (o1 = sum(permute(t, {2, 0, 1}), {2})).run(ref); (o2 = sum(permute(t, {2, 0, 1}), {2})).run(b200); ref.sync();
    // The reference accumulates in bf16 (core/type_utils_both.h:738-746), so ITS distance to the fp64 truth of the
    // same bf16 inputs is ~1e-2; this path accumulates in fp32 and rounds once.  Report all three distances.
    double em = 0, et = 0, er = 0;
    for (index_t i = 0; i < d2; ++i) for (index_t j = 0; j < d0; ++j) {
      double truth = 0;
      for (index_t k = 0; k < d1; ++k) truth += static_cast<float>(t(j, k, i));
      em = std::max(em, std::fabs((double)static_cast<float>(o1(i, j)) - static_cast<float>(o2(i, j))) / truth);
      et = std::max(et, std::fabs((double)static_cast<float>(o2(i, j)) - truth) / truth);
      er = std::max(er, std::fabs((double)static_cast<float>(o1(i, j)) - truth) / truth);
    }
    printf("     bf16 distances: |b200-ref|=%.3g  |b200-fp64|=%.3g  |ref-fp64|=%.3g\n", em, et, er);
    report("sum(permute(t,{2,0,1}),{2}) bf16 vs reference", em <= er + 1.0 / 256 && *b200.last_kernel(), b200.last_kernel(), em);
    report("sum(permute(t,{2,0,1}),{2}) bf16 vs fp64 truth (2^-8)", et <= 1.0 / 256, b200.last_kernel(), et);
  }
  // ---- trace, allclose, softmax_impl, permuted copy ----
  {
    const index_t n = 512;
    auto m = make_tensor<float>({n, n});
    fill_uniform(m, 21, -1.f, 1.f);
    auto s1 = make_tensor<float>({}), s2 = make_tensor<float>({});
    (s1 = trace(m)).run(ref); (s2 = trace(m)).run(b200); ref.sync();
    report("trace(m) = sum(diag(m))", std::fabs(s2() - s1()) <= 1e-5 * std::max(1.f, std::fabs(s1())) && *b200.last_kernel(), b200.last_kernel(), std::fabs(s2() - s1()));
    (s1 = sum(diag(m * m))).run(ref); (s2 = sum(diag(m * m))).run(b200); ref.sync();
    report("sum(diag(m*m)) fused", std::fabs(s2() - s1()) <= 1e-5 * std::fabs(s1()) && *b200.last_kernel(), b200.last_kernel(), std::fabs(s2() - s1()));
    auto d1 = make_tensor<float>({n - 3}), d2 = make_tensor<float>({n - 3});
    (d1 = diag(m, 3) * 2.f).run(ref); (d2 = diag(m, 3) * 2.f).run(b200); ref.sync();
    report("diag(m,3)*2 (off-diagonal view)", max_rel(d2, d1, n - 3) == 0 && *b200.last_kernel(), b200.last_kernel());

    auto m2 = make_tensor<float>({n, n});
    (m2 = m).run(ref); ref.sync();
    m2(100, 7) += 1e-3f;
    auto f1 = make_tensor<int>({}), f2 = make_tensor<int>({});
    allclose(f1, m, m2, 1e-5, 1e-8, ref); allclose(f2, m, m2, 1e-5, 1e-8, b200); ref.sync();
    report("allclose(m, m2) = 0", f1() == 0 && f2() == 0 && *b200.last_kernel(), b200.last_kernel());
    allclose(f1, m, m2, 1e-1, 1e-2, ref); allclose(f2, m, m2, 1e-1, 1e-2, b200); ref.sync();
    report("allclose(m, m2, loose) = 1", f1() == 1 && f2() == 1 && *b200.last_kernel(), b200.last_kernel());

    auto p1 = make_tensor<float>({n, n}), p2 = make_tensor<float>({n, n});
    softmax_impl(p1, m, cuda::std::array<int, 1>{1}, stream);
    softmax_impl(p2, m, cuda::std::array<int, 1>{1}, b200);
    ref.sync();
    double em = 0;
    for (index_t i = 0; i < n; ++i) for (index_t j = 0; j < n; ++j) em = std::max(em, std::fabs((double)p1(i, j) - p2(i, j)) / p1(i, j));
    report("softmax_impl(p, m, {1}, exec) one launch", em <= 1e-5 && !strncmp(b200.last_kernel(), "softmax", 7), b200.last_kernel(), em);
    // Column softmax.  The reference's own dims overload is not usable as the yardstick here: its statistics line
    // subtracts clone(tmp_max, clone_dims) — laid out for the UNPERMUTED input — from permute(in, perm)
    // (transforms/reduce.h:438-440), so for a square matrix every element meets another column's max (2 % off; a
    // non-square one throws on the size check).  Yardstick: the softmax of each column in fp64 on the host.
    softmax_impl(p2, m, cuda::std::array<int, 1>{0}, b200);
    ref.sync();
    em = 0;
    for (index_t j = 0; j < n; ++j) {
      double mxv = -1e300, sum = 0;
      for (index_t i = 0; i < n; ++i) mxv = std::max(mxv, (double)m(i, j));
      for (index_t i = 0; i < n; ++i) sum += std::exp((double)m(i, j) - mxv);
      for (index_t i = 0; i < n; ++i) { const double t = std::exp((double)m(i, j) - mxv) / sum; em = std::max(em, std::fabs(t - p2(i, j)) / t); }
    }
    report("softmax_impl(p, m, {0}, exec) column softmax vs fp64", em <= 1e-5 && *b200.last_kernel(), b200.last_kernel(), em);

    // cumsum: CUBTests.cu:203-226 (permuted int input) and batched float rows
    {
      auto inv = make_tensor<int>({3, 4});
      inv.SetVals({{1, 2, 3, 4}, {10, 20, 30, 40}, {100, 200, 300, 400}});
      auto c1 = make_tensor<int>({4, 3}), c2 = make_tensor<int>({4, 3});
      (c1 = cumsum(inv.Permute({1, 0}))).run(ref); (c2 = cumsum(inv.Permute({1, 0}))).run(b200); ref.sync();
      bool ok = true;
      for (index_t i = 0; i < 4; ++i) { int run = 0; for (index_t j = 0; j < 3; ++j) { run += inv(j, i); ok = ok && c1(i, j) == run && c2(i, j) == run; } }
      report("cumsum(inv.Permute({1,0})) int", ok && !strncmp(b200.last_kernel(), "scan|", 5), b200.last_kernel());
      auto r1 = make_tensor<float>({n, n}), r2 = make_tensor<float>({n, n});
      (r1 = cumsum(m)).run(ref); (r2 = cumsum(m)).run(b200); ref.sync();
      double ec = 0;
      for (index_t i = 0; i < n; ++i) for (index_t j = 0; j < n; ++j) ec = std::max(ec, std::fabs((double)r1(i, j) - r2(i, j)));
      report("cumsum(m) 512 rows, one launch", ec <= 1e-4 && !strncmp(b200.last_kernel(), "scan|", 5), b200.last_kernel(), ec);
      auto lx = make_tensor<float>({1 << 20}), l1 = make_tensor<float>({1 << 20}), l2 = make_tensor<float>({1 << 20});
      for (index_t i = 0; i < (1 << 20); ++i) lx(i) = float((i * 7) % 5) - 2.f;
      (l1 = cumsum(lx)).run(ref); (l2 = cumsum(lx)).run(b200); ref.sync();
      report("cumsum(x) 2^20 in one row (tile exchange)", max_rel(l2, l1, 1 << 20) == 0 && !strncmp(b200.last_kernel(), "scan|", 5), b200.last_kernel());
    }
    // find / find_idx: ReductionTests.cu:1615-1693 (values in [0, 2), GT 0.5) plus ties and a multi-tile input
    {
      const index_t nf_n = 100000;
      auto fx = make_tensor<float>({nf_n});
      for (index_t i = 0; i < nf_n; ++i) fx(i) = float((i * 7919) % 9) * 0.25f;
      auto v1 = make_tensor<float>({nf_n}), v2 = make_tensor<float>({nf_n});
      auto i1 = make_tensor<int>({nf_n}), i2 = make_tensor<int>({nf_n});
      auto n1 = make_tensor<int>({}), n2 = make_tensor<int>({}), n3 = make_tensor<int>({}), n4 = make_tensor<int>({});
      (mtie(v1, n1) = find(fx, GT<float>{0.5f})).run(ref); (mtie(v2, n2) = find(fx, GT<float>{0.5f})).run(b200); ref.sync();
      const std::string kf = b200.last_kernel();
      (mtie(i1, n3) = find_idx(fx, LTE<float>{1.0f})).run(ref); (mtie(i2, n4) = find_idx(fx, LTE<float>{1.0f})).run(b200); ref.sync();
      bool ok = n1() == n2() && n3() == n4() && n1() > 0 && n3() > 0;
      for (index_t i = 0; ok && i < n1(); ++i) ok = v1(i) == v2(i);
      for (index_t i = 0; ok && i < n3(); ++i) ok = i1(i) == i2(i) && fx(i2(i)) <= 1.0f;
      report("(mtie(out, num_found) = find / find_idx(x, SEL{c})) stream compaction", ok && !strncmp(kf.c_str(), "select|", 7) && !strncmp(b200.last_kernel(), "select|", 7), b200.last_kernel());
    }
    // permuted copy: bench/00_operators/operators.cu:40-59, scaled down
    auto x = make_tensor<float>({50, 40, 6, 30});
    std::mt19937 g(23);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    for (index_t a = 0; a < 50; ++a) for (index_t b = 0; b < 40; ++b) for (index_t c = 0; c < 6; ++c) for (index_t d = 0; d < 30; ++d) x(a, b, c, d) = u(g);
    auto y1 = make_tensor<float>({30, 50, 6, 40}), y2 = make_tensor<float>({30, 50, 6, 40});
    (y1 = x.Permute({3, 0, 2, 1})).run(ref); (y2 = x.Permute({3, 0, 2, 1})).run(b200); ref.sync();
    bool same = true;
    for (index_t a = 0; a < 30; ++a) for (index_t b = 0; b < 50; ++b) for (index_t c = 0; c < 6; ++c) for (index_t d = 0; d < 40; ++d)
      same = same && y1(a, b, c, d) == y2(a, b, c, d) && y2(a, b, c, d) == x(b, d, c, a);
    report("(y = x.Permute({3,0,2,1})) tiled transpose", same && !strncmp(b200.last_kernel(), "ew_tr|", 6), b200.last_kernel());
  }
  // ---- round 2: cast / constant / slice / collapse operator nodes are lowered (no fallback) ----
  {
    const index_t rows = 48, cols = 1056;
    auto xi = make_tensor<int32_t>({rows, cols});
    auto a = make_tensor<float>({rows, cols}), bb = make_tensor<float>({rows, cols});
    fill_uniform(a, 31, 0.f, 1.f); fill_uniform(bb, 32, -1.f, 1.f);
    for (index_t i = 0; i < rows; ++i) for (index_t j = 0; j < cols; ++j) xi(i, j) = int32_t((i * 131 + j * 17) % 23) - 11;
    auto v1 = make_tensor<float>({rows}), v2 = make_tensor<float>({rows});
    const long long fb0 = b200.fallbacks();
    (v1 = sum(as_type<float>(xi) * a, {1})).run(ref); (v2 = sum(as_type<float>(xi) * a, {1})).run(b200); ref.sync();
    // (signed terms: the row sums are ~30x smaller than the sum of |terms|, so the bar for two fp32 summation orders is 1e-4)
    report("sum(as_type<float>(xi) * a, {1})  CastOp", max_rel(v2, v1, rows) <= 1e-4 && !strncmp(b200.last_kernel(), "red_", 4), b200.last_kernel(), max_rel(v2, v1, rows));
    (v1 = sum(as_float(xi), {1})).run(ref); (v2 = sum(as_float(xi), {1})).run(b200); ref.sync();
    report("sum(as_float(xi), {1})", max_rel(v2, v1, rows) == 0 && !strncmp(b200.last_kernel(), "red_", 4), b200.last_kernel());
    (v1 = sum(a * ones<float>({rows, cols}) + zeros<float>({rows, cols}), {1})).run(ref);
    (v2 = sum(a * ones<float>({rows, cols}) + zeros<float>({rows, cols}), {1})).run(b200); ref.sync();
    report("sum(a * ones() + zeros(), {1})  ConstVal", max_rel(v2, v1, rows) <= 1e-5 && !strncmp(b200.last_kernel(), "red_", 4), b200.last_kernel(), max_rel(v2, v1, rows));
    auto c1 = make_tensor<int>({}), c2 = make_tensor<int>({});
    (c1 = sum(ones<int>({rows, cols}))).run(ref); (c2 = sum(ones<int>({rows, cols}))).run(b200); ref.sync();
    report("sum(ones<int>({rows, cols}))  (ReductionTests.cu:221-311)", c1() == c2() && c2() == int(rows * cols) && *b200.last_kernel(), b200.last_kernel());
    // slice of an EXPRESSION (SliceOp), unit and non-unit steps, a dropped dim
    auto s1 = make_tensor<float>({rows}), s2 = make_tensor<float>({rows});
    (s1 = sum(slice(a * bb, {0, 16}, {matxEnd, 1040}), {1})).run(ref); (s2 = sum(slice(a * bb, {0, 16}, {matxEnd, 1040}), {1})).run(b200); ref.sync();
    report("sum(slice(a*b, {0,16}, {end,1040}), {1})  SliceOp", max_rel(s2, s1, rows) <= 1e-4 && !strncmp(b200.last_kernel(), "red_", 4), b200.last_kernel(), max_rel(s2, s1, rows));
    // (a strided slice of an OPERATOR does not compile in the reference itself: slice.h:166 assigns through a const
    // reference; strided slices of tensors are views, i.e. plain strides here)
    auto t1 = make_tensor<float>({rows / 2}), t2 = make_tensor<float>({rows / 2});
    (t1 = sum(slice(a, {0, 0}, {matxEnd, matxEnd}, {2, 3}) * 2.f, {1})).run(ref); (t2 = sum(slice(a, {0, 0}, {matxEnd, matxEnd}, {2, 3}) * 2.f, {1})).run(b200); ref.sync();
    report("sum(slice(a, ..., strides {2,3}) * 2, {1})  strided tensor view", max_rel(t2, t1, rows / 2) <= 1e-4 && *b200.last_kernel(), b200.last_kernel(), max_rel(t2, t1, rows / 2));
    auto r1 = make_tensor<float>({cols}), r2 = make_tensor<float>({cols});
    (r1 = slice<1>(a * 2.f, {5, 0}, {matxDropDim, matxEnd}) + 1.f).run(ref); (r2 = slice<1>(a * 2.f, {5, 0}, {matxDropDim, matxEnd}) + 1.f).run(b200); ref.sync();
    report("(r = slice<1>(a*2, {5,0}, {drop,end}) + 1)  dropped dim", max_rel(r2, r1, cols) == 0 && !strncmp(b200.last_kernel(), "ew", 2), b200.last_kernel());
    // collapse nodes over contiguous operands
    auto t3 = make_tensor<float>({6, 8, 352});
    for (index_t i = 0; i < 6; ++i) for (index_t j = 0; j < 8; ++j) for (index_t k = 0; k < 352; ++k) t3(i, j, k) = float((i * 8 + j) % 5) + 0.001f * float(k);
    auto l1 = make_tensor<float>({48}), l2 = make_tensor<float>({48});
    (l1 = sum(lcollapse<2>(t3 * 2.f), {1})).run(ref); (l2 = sum(lcollapse<2>(t3 * 2.f), {1})).run(b200); ref.sync();
    report("sum(lcollapse<2>(t3*2), {1})  LCollapseOp", max_rel(l2, l1, 48) <= 1e-5 && !strncmp(b200.last_kernel(), "red_", 4), b200.last_kernel(), max_rel(l2, l1, 48));
    auto q1 = make_tensor<float>({6}), q2 = make_tensor<float>({6});
    (q1 = max(rcollapse<2>(t3 + 1.f), {1})).run(ref); (q2 = max(rcollapse<2>(t3 + 1.f), {1})).run(b200); ref.sync();
    report("max(rcollapse<2>(t3+1), {1})  RCollapseOp", max_rel(q2, q1, 6) == 0 && !strncmp(b200.last_kernel(), "red_", 4), b200.last_kernel());
    report("fallbacks() stays 0 over the lowered node types", b200.fallbacks() == fb0, "");
    // a collapse over a view that is NOT contiguous across the collapsed dims cannot be a stride: reference path, counted
    auto tp = t3.Permute({1, 0, 2});
    (l1 = sum(lcollapse<2>(tp), {1})).run(ref); (l2 = sum(lcollapse<2>(tp), {1})).run(b200); ref.sync();
    report("sum(lcollapse<2>(permuted view), {1}) falls back, is counted", max_rel(l2, l1, 48) <= 1e-5 && b200.fallbacks() == fb0 + 1, "");
  }
  // ---- round 2: sort / unique (executor-typed seams) and, with the overlay headers, hist and the softmax OPERATOR ----
  {
    const index_t n = 100000;
    auto x = make_tensor<float>({n});
    std::mt19937 g(91);
    std::uniform_int_distribution<int> u(-500, 500);
    for (index_t i = 0; i < n; ++i) x(i) = 0.25f * float(u(g));
    auto s1 = make_tensor<float>({n}), s2 = make_tensor<float>({n});
    const long long fb0 = b200.fallbacks();
    (s1 = matx::sort(x, SORT_DIR_ASC)).run(ref); (s2 = matx::sort(x, SORT_DIR_ASC)).run(b200); ref.sync();
    report("(out = sort(x, SORT_DIR_ASC)) 100000 keys (radix)", max_rel(s2, s1, n) == 0 && !strncmp(b200.last_kernel(), "sort_radix", 10), b200.last_kernel());
    (s1 = matx::sort(x, SORT_DIR_DESC)).run(ref); (s2 = matx::sort(x, SORT_DIR_DESC)).run(b200); ref.sync();
    report("(out = sort(x, SORT_DIR_DESC))", max_rel(s2, s1, n) == 0 && !strncmp(b200.last_kernel(), "sort_radix", 10), b200.last_kernel());
    auto m = make_tensor<float>({64, 1000}), m1 = make_tensor<float>({64, 1000}), m2 = make_tensor<float>({64, 1000});
    for (index_t i = 0; i < 64; ++i) for (index_t j = 0; j < 1000; ++j) m(i, j) = 0.5f * float(u(g));
    (m1 = matx::sort(m, SORT_DIR_ASC)).run(ref); (m2 = matx::sort(m, SORT_DIR_ASC)).run(b200); ref.sync();
    bool same = true;
    for (index_t i = 0; i < 64; ++i) for (index_t j = 0; j < 1000; ++j) same = same && m1(i, j) == m2(i, j);
    report("(out = sort(m, ASC)) 64 rows of 1000 (one CTA per row)", same && !strncmp(b200.last_kernel(), "sort_bitonic", 12), b200.last_kernel());
    auto u1 = make_tensor<float>({n}), u2 = make_tensor<float>({n});
    auto c1 = make_tensor<int>({}), c2 = make_tensor<int>({});
    (mtie(u1, c1) = unique(x)).run(ref); (mtie(u2, c2) = unique(x)).run(b200); ref.sync();
    report("(mtie(out, num_found) = unique(x))", c1() == c2() && max_rel(u2, u1, c1()) == 0 && !strncmp(b200.last_kernel(), "select|", 7), b200.last_kernel());
#ifdef MXB_OVERLAY
    auto h1 = make_tensor<int>({16}), h2 = make_tensor<int>({16});
    (h1 = hist(x, -125.0f, 125.0f, 17)).run(ref); (h2 = hist(x, -125.0f, 125.0f, 17)).run(b200); ref.sync();
    same = true;
    for (index_t i = 0; i < 16; ++i) same = same && h1(i) == h2(i);
    report("(out = hist(x, lo, hi, levels)) through the overlay header", same && !strncmp(b200.last_kernel(), "hist|", 5), b200.last_kernel());
    auto p1 = make_tensor<float>({64, 1000}), p2 = make_tensor<float>({64, 1000});
    (p1 = softmax(m * 0.01f, {1})).run(ref); (p2 = softmax(m * 0.01f, {1})).run(b200); ref.sync();
    double em = 0;
    for (index_t i = 0; i < 64; ++i) for (index_t j = 0; j < 1000; ++j) em = std::max(em, (double)std::fabs(p2(i, j) - p1(i, j)) / std::max(1e-12, (double)std::fabs(p1(i, j))));
    report("(out = softmax(x, {1})).run(exec) OPERATOR form through the overlay header", em <= 1e-5 && !strncmp(b200.last_kernel(), "softmax", 7), b200.last_kernel(), em);
#endif
    report("fallbacks() stays 0 over sort / unique / hist / softmax", b200.fallbacks() == fb0, "");
  }
  // ---- a node the shim does not lower falls back to the reference, same answer ----
  {
    auto a = make_tensor<float>({1000});
    for (index_t i = 0; i < 1000; ++i) a(i) = float(i % 37);
    auto o1 = make_tensor<float>({1000}), o2 = make_tensor<float>({1000});
    (o1 = shift<0>(a, 3) + 1.f).run(ref); (o2 = shift<0>(a, 3) + 1.f).run(b200); ref.sync();
    const long long fb1 = b200.fallbacks();
    (o2 = shift<0>(a, 3) + 1.f).run(b200); ref.sync();
    report("unknown node (shift) falls back", max_rel(o2, o1, 1000) == 0 && b200.fallbacks() == fb1 + 1, "");
  }
  printf("SUMMARY pass=%d fail=%d native_launches=%lld\n", g_pass, g_fail, b200.native_launches());
  return g_fail;
}

template <class F> static float time_ms(cudaStream_t s, int iters, F &&f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b, s);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

static void run_bench() {
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  cudaExecutor ref{stream};
  b200Executor b200{stream};
  auto line = [](const char *cfg, const char *op, double bytes, float ms_ref, float ms_new) {
    printf("BENCH {\"config\": \"%s\", \"op\": \"%s\", \"algorithmic_bytes\": %.0f, \"reference_cudaExecutor_ms\": %.4f, \"b200Executor_ms\": %.4f, "
           "\"reference_GBps\": %.1f, \"b200_GBps\": %.1f, \"speedup\": %.2f}\n",
           cfg, op, bytes, ms_ref, ms_new, bytes / ms_ref / 1e6, bytes / ms_new / 1e6, ms_ref / ms_new);
    fflush(stdout);
  };
  {  // C2
    const index_t n = index_t(1) << 30;
    auto x = make_tensor<float>({n}, MATX_DEVICE_MEMORY);
    (x = random<float>({n}, UNIFORM)).run(ref);
    auto s = make_tensor<float>({}, MATX_DEVICE_MEMORY);
    auto i = make_tensor<index_t>({}, MATX_DEVICE_MEMORY);
    line("C2 fp32 2^30", "sum", n * 4.0, time_ms(stream, 10, [&] { (s = sum(x)).run(ref); }), time_ms(stream, 10, [&] { (s = sum(x)).run(b200); }));
    line("C2 fp32 2^30", "max", n * 4.0, time_ms(stream, 10, [&] { (s = max(x)).run(ref); }), time_ms(stream, 10, [&] { (s = max(x)).run(b200); }));
    line("C2 fp32 2^30", "argmax", n * 4.0, time_ms(stream, 10, [&] { (mtie(s, i) = argmax(x)).run(ref); }), time_ms(stream, 10, [&] { (mtie(s, i) = argmax(x)).run(b200); }));
    auto s2 = make_tensor<float>({}, MATX_DEVICE_MEMORY);
    auto i2 = make_tensor<index_t>({}, MATX_DEVICE_MEMORY);
    line("C2 fp32 2^30", "argminmax", n * 4.0, time_ms(stream, 5, [&] { (mtie(s, i, s2, i2) = argminmax(x)).run(ref); }),
         time_ms(stream, 10, [&] { (mtie(s, i, s2, i2) = argminmax(x)).run(b200); }));
  }
  {  // C1
    const index_t rows = 16384, cols = 4096;
    auto a = make_tensor<float>({rows, cols}, MATX_DEVICE_MEMORY), b = make_tensor<float>({rows, cols}, MATX_DEVICE_MEMORY), c = make_tensor<float>({rows, cols}, MATX_DEVICE_MEMORY);
    (a = random<float>({rows, cols}, UNIFORM)).run(ref); (b = random<float>({rows, cols}, UNIFORM)).run(ref); (c = random<float>({rows, cols}, UNIFORM)).run(ref);
    auto o = make_tensor<float>({rows}, MATX_DEVICE_MEMORY);
    line("C1 fp32 16384x4096", "sum(a*b+c,{1})", 3.0 * rows * cols * 4 + rows * 4, time_ms(stream, 10, [&] { (o = sum(a * b + c, {1})).run(ref); }),
         time_ms(stream, 10, [&] { (o = sum(a * b + c, {1})).run(b200); }));
  }
  {  // C3
    const index_t rows = 65536, cols = 8192;
    auto x = make_tensor<cf>({rows, cols}, MATX_DEVICE_MEMORY);
    (x = random<cf>({rows, cols}, NORMAL)).run(ref);
    auto m = make_tensor<cf>({rows}, MATX_DEVICE_MEMORY);
    auto v = make_tensor<float>({rows}, MATX_DEVICE_MEMORY);
    auto i = make_tensor<index_t>({rows}, MATX_DEVICE_MEMORY);
    const double bytes = 8.0 * rows * cols;
    line("C3 c64 65536x8192", "mean(x,{1})", bytes, time_ms(stream, 5, [&] { (m = mean(x, {1})).run(ref); }), time_ms(stream, 5, [&] { (m = mean(x, {1})).run(b200); }));
    line("C3 c64 65536x8192", "var(x,{1})", bytes, time_ms(stream, 5, [&] { (v = var(x, {1}, 1)).run(ref); }), time_ms(stream, 5, [&] { (v = var(x, {1}, 1)).run(b200); }));
    line("C3 c64 65536x8192", "argmax(abs2(x),{1})", bytes, time_ms(stream, 5, [&] { (mtie(v, i) = argmax(abs2(x), {1})).run(ref); }),
         time_ms(stream, 5, [&] { (mtie(v, i) = argmax(abs2(x), {1})).run(b200); }));
  }
  {  // C4
    const index_t n = index_t(1) << 28;
    auto K = make_tensor<float>({n}, MATX_DEVICE_MEMORY), S = make_tensor<float>({n}, MATX_DEVICE_MEMORY), V = make_tensor<float>({n}, MATX_DEVICE_MEMORY),
         r = make_tensor<float>({n}, MATX_DEVICE_MEMORY), T = make_tensor<float>({n}, MATX_DEVICE_MEMORY), out = make_tensor<float>({n}, MATX_DEVICE_MEMORY);
    (S = random<float>({n}, UNIFORM) * 90.f + 10.f).run(ref); (K = random<float>({n}, UNIFORM) * 90.f + 10.f).run(ref);
    (V = random<float>({n}, UNIFORM) * 0.45f + 0.05f).run(ref); (r = random<float>({n}, UNIFORM) * 0.09f + 0.01f).run(ref);
    (T = random<float>({n}, UNIFORM) * 1.9f + 0.1f).run(ref);
    auto bs = [&](auto &exec) {
      auto VsqrtT = V * sqrt(T);
      auto d1 = (log(S / K) + (r + 0.5f * V * V) * T) / VsqrtT;
      auto d2 = d1 - VsqrtT;
      (out = S * normcdf(d1) - K * exp(-1.f * r * T) * normcdf(d2)).run(exec);
    };
    line("C4 fp32 2^28", "black_scholes", 6.0 * n * 4, time_ms(stream, 5, [&] { bs(ref); }), time_ms(stream, 5, [&] { bs(b200); }));
    line("vector_add fp32 2^28", "out = S + K", 3.0 * n * 4, time_ms(stream, 5, [&] { (out = S + K).run(ref); }), time_ms(stream, 5, [&] { (out = S + K).run(b200); }));
  }
  {  // permuted copy, the reference's own benchmark (bench/00_operators/operators.cu:40-59)
    auto x = make_tensor<float>({1000, 200, 6, 300}, MATX_DEVICE_MEMORY);
    auto y = make_tensor<float>({300, 1000, 6, 200}, MATX_DEVICE_MEMORY);
    (x = random<float>({1000, 200, 6, 300}, UNIFORM)).run(ref);
    line("permute fp32 {1000,200,6,300}", "y = x.Permute({3,0,2,1})", 2.0 * 4 * 1000 * 200 * 6 * 300, time_ms(stream, 3, [&] { (y = x.Permute({3, 0, 2, 1})).run(ref); }),
         time_ms(stream, 5, [&] { (y = x.Permute({3, 0, 2, 1})).run(b200); }));
    auto a = make_tensor<float>({8192, 8192}, MATX_DEVICE_MEMORY), t = make_tensor<float>({8192, 8192}, MATX_DEVICE_MEMORY);
    (a = random<float>({8192, 8192}, UNIFORM)).run(ref);
    line("transpose fp32 8192x8192", "t = a.Permute({1,0})", 2.0 * 4 * 8192 * 8192, time_ms(stream, 5, [&] { (t = a.Permute({1, 0})).run(ref); }),
         time_ms(stream, 5, [&] { (t = a.Permute({1, 0})).run(b200); }));
    line("transpose fp32 8192x8192", "t = transpose_matrix(a) (reference's tiled kernel)", 2.0 * 4 * 8192 * 8192, time_ms(stream, 5, [&] { (t = transpose_matrix(a)).run(ref); }),
         time_ms(stream, 5, [&] { (t = a.Permute({1, 0})).run(b200); }));
  }
  {  // cumsum: many rows (the reference launches CUB once per row) and one long row
    const index_t rows = 4096, cols = 8192;
    auto x = make_tensor<float>({rows, cols}, MATX_DEVICE_MEMORY), y = make_tensor<float>({rows, cols}, MATX_DEVICE_MEMORY);
    (x = random<float>({rows, cols}, UNIFORM)).run(ref);
    line("cumsum fp32 4096x8192", "y = cumsum(x)", 2.0 * 4 * rows * cols, time_ms(stream, 2, [&] { (y = cumsum(x)).run(ref); }),
         time_ms(stream, 5, [&] { (y = cumsum(x)).run(b200); }));
    const index_t n = index_t(1) << 28;
    auto lx = make_tensor<float>({n}, MATX_DEVICE_MEMORY), ly = make_tensor<float>({n}, MATX_DEVICE_MEMORY);
    (lx = random<float>({n}, UNIFORM)).run(ref);
    line("cumsum fp32 2^28", "y = cumsum(x)", 2.0 * 4 * n, time_ms(stream, 5, [&] { (ly = cumsum(lx)).run(ref); }),
         time_ms(stream, 5, [&] { (ly = cumsum(lx)).run(b200); }));
  }
  {  // find: stream compaction of 2^28 fp32 at 1 % and 50 % selectivity
    const index_t n = index_t(1) << 28;
    auto fx = make_tensor<float>({n}, MATX_DEVICE_MEMORY), fo = make_tensor<float>({n}, MATX_DEVICE_MEMORY);
    auto nf = make_tensor<int>({}, MATX_DEVICE_MEMORY);
    (fx = random<float>({n}, UNIFORM)).run(ref);
    line("find fp32 2^28 (1 % selected)", "mtie(out, n) = find(x, GT{0.99})", 4.0 * n * 1.01, time_ms(stream, 3, [&] { (mtie(fo, nf) = find(fx, GT<float>{0.99f})).run(ref); }),
         time_ms(stream, 5, [&] { (mtie(fo, nf) = find(fx, GT<float>{0.99f})).run(b200); }));
    line("find fp32 2^28 (50 % selected)", "mtie(out, n) = find(x, GT{0.5})", 4.0 * n * 1.5, time_ms(stream, 3, [&] { (mtie(fo, nf) = find(fx, GT<float>{0.5f})).run(ref); }),
         time_ms(stream, 5, [&] { (mtie(fo, nf) = find(fx, GT<float>{0.5f})).run(b200); }));
    auto fi = make_tensor<index_t>({n}, MATX_DEVICE_MEMORY);
    line("find_idx fp32 2^28 (1 % selected)", "mtie(idx, n) = find_idx(x, GT{0.99})", 4.0 * n + 8.0 * n * 0.01, time_ms(stream, 3, [&] { (mtie(fi, nf) = find_idx(fx, GT<float>{0.99f})).run(ref); }),
         time_ms(stream, 5, [&] { (mtie(fi, nf) = find_idx(fx, GT<float>{0.99f})).run(b200); }));
    line("find_idx fp32 2^28 (50 % selected)", "mtie(idx, n) = find_idx(x, GT{0.5})", 4.0 * n + 8.0 * n * 0.5, time_ms(stream, 3, [&] { (mtie(fi, nf) = find_idx(fx, GT<float>{0.5f})).run(ref); }),
         time_ms(stream, 5, [&] { (mtie(fi, nf) = find_idx(fx, GT<float>{0.5f})).run(b200); }));
  }
  {  // sort / unique / hist: the rank-3 neighbours beside cub::DeviceRadixSort / DeviceSelect::Unique / DeviceHistogram
    const index_t n = index_t(1) << 24;
    auto x = make_tensor<float>({n}, MATX_DEVICE_MEMORY), y = make_tensor<float>({n}, MATX_DEVICE_MEMORY);
    (x = random<float>({n}, UNIFORM)).run(ref);
    line("sort fp32 2^24", "y = sort(x, SORT_DIR_ASC)", 2.0 * 4 * n, time_ms(stream, 3, [&] { (y = matx::sort(x, SORT_DIR_ASC)).run(ref); }),
         time_ms(stream, 3, [&] { (y = matx::sort(x, SORT_DIR_ASC)).run(b200); }));
    auto m = make_tensor<float>({16384, 1024}, MATX_DEVICE_MEMORY), ms = make_tensor<float>({16384, 1024}, MATX_DEVICE_MEMORY);
    (m = random<float>({16384, 1024}, UNIFORM)).run(ref);
    line("sort fp32 16384x1024", "ms = sort(m, SORT_DIR_ASC) (rows)", 2.0 * 4 * 16384 * 1024, time_ms(stream, 2, [&] { (ms = matx::sort(m, SORT_DIR_ASC)).run(ref); }),
         time_ms(stream, 3, [&] { (ms = matx::sort(m, SORT_DIR_ASC)).run(b200); }));
    auto q = make_tensor<float>({n}, MATX_DEVICE_MEMORY);
    (q = floor(x * 1000.f)).run(ref);
    auto nu = make_tensor<int>({}, MATX_DEVICE_MEMORY);
    line("unique fp32 2^24 (1000 distinct)", "mtie(u, n) = unique(q)", 3.0 * 4 * n, time_ms(stream, 3, [&] { (mtie(y, nu) = unique(q)).run(ref); }),
         time_ms(stream, 3, [&] { (mtie(y, nu) = unique(q)).run(b200); }));
#ifdef MXB_OVERLAY
    const index_t nh = index_t(1) << 28;
    auto hx = make_tensor<float>({nh}, MATX_DEVICE_MEMORY);
    (hx = random<float>({nh}, UNIFORM)).run(ref);
    auto hb = make_tensor<int>({256}, MATX_DEVICE_MEMORY);
    line("hist fp32 2^28, 256 bins", "h = hist(x, 0, 1, 257)", 4.0 * nh, time_ms(stream, 3, [&] { (hb = hist(hx, 0.0f, 1.0f, 257)).run(ref); }),
         time_ms(stream, 5, [&] { (hb = hist(hx, 0.0f, 1.0f, 257)).run(b200); }));
#endif
  }
  {  // C5
    const index_t d = 1024;
    auto t = make_tensor<matxBf16>({d, d, d}, MATX_DEVICE_MEMORY);
    (t = as_type<matxBf16>(random<float>({d, d, d}, UNIFORM) * 0.25f)).run(ref);
    auto o = make_tensor<matxBf16>({d, d}, MATX_DEVICE_MEMORY);
    line("C5 bf16 1024^3", "sum(permute(t,{2,0,1}),{2})", 2.0 * d * d * d + 2.0 * d * d, time_ms(stream, 3, [&] { (o = sum(permute(t, {2, 0, 1}), {2})).run(ref); }),
         time_ms(stream, 5, [&] { (o = sum(permute(t, {2, 0, 1}), {2})).run(b200); }));
  }
}

int main(int argc, char **argv) {
  MATX_ENTER_HANDLER();
  if (argc > 1 && !strcmp(argv[1], "--bench")) { run_bench(); return 0; }
  return run_checks() ? 1 : 0;
  MATX_EXIT_HANDLER();
}
