// matx_b200/executor.h — the drop-in: a MatX executor that sends the fused-elementwise / reduction hot path to
// libmatx_b200.so (hand-written sm_100a kernels) and leaves everything else to the reference.
//
//     #include <matx.h>
//     #include <matx_b200/executor.h>          // after matx.h; link with -lmatx_b200
//     matx::b200Executor exec{stream};
//     (out = sum(a*b+c, {1})).run(exec);        // unchanged MatX statements
//     (mtie(v, i) = argmax(abs2(x), {1})).run(exec);
//     (o = S*normcdf(d1) - K*exp(-1.f*r*T)*normcdf(d2)).run(exec);
//
// How it plugs in (all reference citations relative to include/matx/):
//   * b200Executor derives from matx::cudaExecutor (executors/cuda.h:60-82), so it IS a CUDA executor for every
//     trait (`matx_executor`, `cuda_executor`, core/type_utils_both.h:393-400) and every transform this path does
//     not cover keeps running through the reference's own cudaExecutor overloads by derived-to-base conversion.
//   * BaseOp::run (operators/base_operator.h:181-285) calls `ex.Exec(set_node)` for plain assignments: Exec below
//     lowers `set<tensor, expr>` to an mxb_expr_t program + output view and calls mxb_elementwise.
//   * the transform nodes call unqualified `sum_impl(dest, in, ex)`, `argmax_impl(dest, idest, in, ex)`, ...
//     (operators/sum.h:304-307, argmax.h:72-76, var.h:90, ...): the overloads at the bottom of this file take a
//     b200Executor (exact match beats the reference's `const cudaExecutor&` overloads in transforms/reduce.h) and
//     call mxb_reduce.
//   * anything the lowering does not know (a node type outside the list below, complex<double>, rank > 8, views that
//     do not collapse) falls back — at compile time where the type says so, at run time on MXB_ERR_NOT_SUPPORTED —
//     to the reference implementation on the same stream.  Other failures become matxException like the
//     reference's (core/error.h:55-75): MXB_ERR_CUDA -> matxCudaError, MXB_ERR_INVALID -> matxInvalidParameter,
//     MXB_ERR_SIZE -> matxInvalidSize.
//
// Lowered node types: tensor views (any strides: permuted / sliced / cloned views are just strides), arithmetic
// scalars, matxBinaryOp, matxUnaryOp with the functors of operators/scalar_ops.h:434-503, PermuteOp, CloneOp, DiagOp
// (so trace(x) = sum(diag(x)) is one strided reduction) and IsCloseOp (so allclose is one fused all-reduction).
// The expression-template nodes keep their operands private, so the walker reads them through layout mirrors
// (same member types in the same order; sizes are checked with static_assert).
#pragma once

#include <matx.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>

#include "../matx_b200.h"

namespace matx {

namespace b200_detail {

// ---- element types -------------------------------------------------------------------------------------------
template <class T> struct dtype_of { static constexpr int value = -1; };
template <> struct dtype_of<float> { static constexpr int value = MXB_F32; };
template <> struct dtype_of<double> { static constexpr int value = MXB_F64; };
template <> struct dtype_of<matxBf16> { static constexpr int value = MXB_BF16; };
template <> struct dtype_of<matxFp16> { static constexpr int value = MXB_F16; };
template <> struct dtype_of<cuda::std::complex<float>> { static constexpr int value = MXB_C64; };
template <> struct dtype_of<int32_t> { static constexpr int value = MXB_I32; };
template <> struct dtype_of<long long> { static constexpr int value = MXB_I64; };
template <> struct dtype_of<long> { static constexpr int value = sizeof(long) == 8 ? MXB_I64 : MXB_I32; };
template <> struct dtype_of<uint8_t> { static constexpr int value = MXB_U8; };
template <> struct dtype_of<bool> { static constexpr int value = MXB_U8; };

// ---- functor -> opcode ------------------------------------------------------------------------------------------
template <class F> struct fn_code { static constexpr int value = -1; };
#define MXB_SHIM_BIN(NAME, CODE) \
  template <class A, class B> struct fn_code<matx::detail::NAME##Op<A, B>> { static constexpr int value = CODE; };
#define MXB_SHIM_UN(NAME, CODE) \
  template <class A> struct fn_code<matx::detail::NAME##Op<A>> { static constexpr int value = CODE; };
MXB_SHIM_BIN(Add, MXB_OP_ADD) MXB_SHIM_BIN(Sub, MXB_OP_SUB) MXB_SHIM_BIN(Mul, MXB_OP_MUL) MXB_SHIM_BIN(Div, MXB_OP_DIV)
MXB_SHIM_BIN(Mod, MXB_OP_MOD) MXB_SHIM_BIN(FMod, MXB_OP_MOD) MXB_SHIM_BIN(Pow, MXB_OP_POW) MXB_SHIM_BIN(Maximum, MXB_OP_MAX)
MXB_SHIM_BIN(Minimum, MXB_OP_MIN) MXB_SHIM_BIN(Atan2, MXB_OP_ATAN2) MXB_SHIM_BIN(LT, MXB_OP_LT) MXB_SHIM_BIN(GT, MXB_OP_GT)
MXB_SHIM_BIN(LTE, MXB_OP_LE) MXB_SHIM_BIN(GTE, MXB_OP_GE) MXB_SHIM_BIN(EQ, MXB_OP_EQ) MXB_SHIM_BIN(NE, MXB_OP_NE)
MXB_SHIM_BIN(AndAnd, MXB_OP_AND) MXB_SHIM_BIN(OrOr, MXB_OP_OR)
MXB_SHIM_UN(Sqrt, MXB_OP_SQRT) MXB_SHIM_UN(RSqrt, MXB_OP_RSQRT) MXB_SHIM_UN(Exp, MXB_OP_EXP) MXB_SHIM_UN(Expj, MXB_OP_EXPJ)
MXB_SHIM_UN(Log10, MXB_OP_LOG10) MXB_SHIM_UN(Log2, MXB_OP_LOG2) MXB_SHIM_UN(Log, MXB_OP_LOG) MXB_SHIM_UN(Conj, MXB_OP_CONJ)
MXB_SHIM_UN(Abs, MXB_OP_ABS) MXB_SHIM_UN(Abs2, MXB_OP_ABS2) MXB_SHIM_UN(Sin, MXB_OP_SIN) MXB_SHIM_UN(Cos, MXB_OP_COS)
MXB_SHIM_UN(Tan, MXB_OP_TAN) MXB_SHIM_UN(Asin, MXB_OP_ASIN) MXB_SHIM_UN(Acos, MXB_OP_ACOS) MXB_SHIM_UN(Atan, MXB_OP_ATAN)
MXB_SHIM_UN(Sinh, MXB_OP_SINH) MXB_SHIM_UN(Cosh, MXB_OP_COSH) MXB_SHIM_UN(Tanh, MXB_OP_TANH) MXB_SHIM_UN(Floor, MXB_OP_FLOOR)
MXB_SHIM_UN(Ceil, MXB_OP_CEIL) MXB_SHIM_UN(Round, MXB_OP_ROUND) MXB_SHIM_UN(NormCdf, MXB_OP_NORMCDF) MXB_SHIM_UN(Real, MXB_OP_REAL)
MXB_SHIM_UN(Imag, MXB_OP_IMAG) MXB_SHIM_UN(SubNeg, MXB_OP_NEG) MXB_SHIM_UN(IsNan, MXB_OP_ISNAN) MXB_SHIM_UN(IsInf, MXB_OP_ISINF)
MXB_SHIM_UN(Not, MXB_OP_NOT)
#undef MXB_SHIM_BIN
#undef MXB_SHIM_UN

// ---- node kinds and layout mirrors -----------------------------------------------------------------------------
template <class T> struct node_kind { static constexpr int value = 0; };  // 0 unknown, 1 binary, 2 unary, 3 permute, 4 clone
template <class I1, class I2, class F> struct node_kind<matx::detail::matxBinaryOp<I1, I2, F>> {
  static constexpr int value = 1;
  using A = I1; using B = I2; using Fn = F;
  struct Mirror { matx::detail::base_type_t<I1> in1_; matx::detail::base_type_t<I2> in2_; matx::detail::base_type_t<F> op_; };
};
template <class I1, class F> struct node_kind<matx::detail::matxUnaryOp<I1, F>> {
  static constexpr int value = 2;
  using A = I1; using Fn = F;
  struct Mirror { matx::detail::base_type_t<I1> in1_; matx::detail::base_type_t<F> op_; cuda::std::array<index_t, matx::detail::get_rank<I1>()> size_; };
};
template <class T> struct node_kind<matx::detail::PermuteOp<T>> {
  static constexpr int value = 3;
  using A = T;
  struct Mirror { matx::detail::base_type_t<T> op_; cuda::std::array<int32_t, T::Rank()> dims_; };
};
template <int CRank, class T> struct node_kind<matx::detail::CloneOp<CRank, T>> {
  static constexpr int value = 4;
  using A = T;
  static constexpr int crank = CRank;
  struct Mirror { matx::detail::base_type_t<T> op_; cuda::std::array<index_t, CRank> sizes_; cuda::std::array<index_t, T::Rank()> dims_; };
};

template <class T, int RANK> struct node_kind<matx::detail::DiagOp<T, RANK>> {   // operators/diag.h:52-56
  static constexpr int value = (RANK >= 2) ? 5 : 0;   // diag(vector) builds a matrix: not on this path
  using A = T;
  struct Mirror { matx::detail::base_type_t<T> op_; index_t k_; };
};
template <class O1, class O2> struct node_kind<matx::detail::IsCloseOp<O1, O2>> {   // operators/isclose.h:234-238
  static constexpr int value = 6;
  using A = O1; using B = O2;
  using inner = typename matx::detail::IsCloseOp<O1, O2>::inner_type;
  struct Mirror { matx::detail::base_type_t<O1> op1_; matx::detail::base_type_t<O2> op2_; inner rtol_; inner atol_; };
};

template <class T> constexpr int rank_of_operand() {
  if constexpr (std::is_arithmetic_v<T> || is_complex_v<T>) return 0;
  else return T::Rank();
}

// compile-time: can this operand type be lowered at all?
template <class T0> constexpr bool lowerable() {
  using T = remove_cvref_t<T0>;
  if constexpr (std::is_arithmetic_v<T>) return dtype_of<T>::value >= 0;
  else if constexpr (std::is_same_v<T, cuda::std::complex<float>>) return true;
  else if constexpr (is_tensor_view_v<T> || matx::is_tensor_impl_v<T>) return dtype_of<typename T::value_type>::value >= 0 && T::Rank() <= MXB_MAX_RANK;
  else if constexpr (node_kind<T>::value == 1)
    return fn_code<typename node_kind<T>::Fn>::value >= 0 && lowerable<typename node_kind<T>::A>() && lowerable<typename node_kind<T>::B>();
  else if constexpr (node_kind<T>::value == 2) return fn_code<typename node_kind<T>::Fn>::value >= 0 && lowerable<typename node_kind<T>::A>();
  else if constexpr (node_kind<T>::value == 3 || node_kind<T>::value == 4 || node_kind<T>::value == 5) return lowerable<typename node_kind<T>::A>();
  else if constexpr (node_kind<T>::value == 6)
    return (std::is_same_v<typename node_kind<T>::inner, float> || std::is_same_v<typename node_kind<T>::inner, double>) &&
           lowerable<typename node_kind<T>::A>() && lowerable<typename node_kind<T>::B>();
  else return false;
}

struct Builder {
  mxb_expr_t e;
  bool ok = true;
  Builder() { std::memset(&e, 0, sizeof e); }
  // Expression templates hold their operands by value, so `auto d1 = ...;` used twice arrives as two identical
  // sub-trees: nodes, leaves and constants are hash-consed here and the program becomes a DAG again (every distinct
  // tensor view is one leaf, i.e. one load per element — the reference re-loads it per occurrence,
  // examples/black_scholes.cu:42-49).
  int node(int opcode, int s0, int s1 = -1, int aux = 0) {
    for (int i = 0; i < e.n_nodes; ++i) {
      const mxb_node_t &n = e.nodes[i];
      if (n.opcode == opcode && n.src[0] == s0 && n.src[1] == s1 && n.aux == aux) return i;
    }
    if (e.n_nodes >= MXB_MAX_NODES) { ok = false; return 0; }
    mxb_node_t &n = e.nodes[e.n_nodes];
    n.opcode = opcode; n.src[0] = s0; n.src[1] = s1; n.aux = aux;
    return e.n_nodes++;
  }
  int leaf(const mxb_leaf_t &lf) {
    for (int k = 0; k < e.n_leaves; ++k) {
      if (e.leaves[k].data == lf.data && e.leaves[k].dtype == lf.dtype &&
          std::memcmp(e.leaves[k].stride, lf.stride, sizeof lf.stride) == 0) return node(MXB_OP_LEAF, k);
    }
    if (e.n_leaves >= MXB_MAX_LEAVES) { ok = false; return 0; }
    e.leaves[e.n_leaves] = lf;
    return node(MXB_OP_LEAF, e.n_leaves++);
  }
  int constant(double re, double im, int dtype) {
    for (int k = 0; k < e.n_consts; ++k) {
      if (e.consts[k].dtype == dtype && std::memcmp(&e.consts[k].re, &re, sizeof re) == 0 && std::memcmp(&e.consts[k].im, &im, sizeof im) == 0)
        return node(MXB_OP_CONST, k);
    }
    if (e.n_consts >= MXB_MAX_CONSTS) { ok = false; return 0; }
    mxb_const_t &c = e.consts[e.n_consts];
    c.re = re; c.im = im; c.dtype = dtype;
    return node(MXB_OP_CONST, e.n_consts++);
  }
};

// axes[i] = dim of the root index space that dim i of `op` walks
template <class Op0> int lower(Builder &b, const Op0 &op, const int *axes) {
  using Op = remove_cvref_t<Op0>;
  if constexpr (std::is_arithmetic_v<Op> || std::is_same_v<Op, cuda::std::complex<float>>) {
    if constexpr (std::is_arithmetic_v<Op>) return b.constant(static_cast<double>(op), 0.0, dtype_of<Op>::value);
    else return b.constant(op.real(), op.imag(), dtype_of<Op>::value);
  } else if constexpr (is_tensor_view_v<Op> || matx::is_tensor_impl_v<Op>) {
    mxb_leaf_t lf;
    std::memset(&lf, 0, sizeof lf);
    lf.data = op.Data();
    lf.dtype = dtype_of<typename Op::value_type>::value;
    for (int d = 0; d < Op::Rank(); ++d) lf.stride[axes[d]] += op.Stride(d);
    return b.leaf(lf);
  } else if constexpr (node_kind<Op>::value == 1) {
    using K = node_kind<Op>;
    static_assert(sizeof(typename K::Mirror) == sizeof(Op), "matxBinaryOp layout changed: update the mirror");
    const auto &m = reinterpret_cast<const typename K::Mirror &>(op);
    constexpr int R = Op::Rank(), RA = rank_of_operand<remove_cvref_t<decltype(m.in1_)>>(), RB = rank_of_operand<remove_cvref_t<decltype(m.in2_)>>();
    const int a = lower(b, m.in1_, axes + (R - RA));  // lower-rank operands line up with the trailing dims
    const int c = lower(b, m.in2_, axes + (R - RB));
    return b.node(fn_code<typename K::Fn>::value, a, c);
  } else if constexpr (node_kind<Op>::value == 2) {
    using K = node_kind<Op>;
    static_assert(sizeof(typename K::Mirror) == sizeof(Op), "matxUnaryOp layout changed: update the mirror");
    const auto &m = reinterpret_cast<const typename K::Mirror &>(op);
    return b.node(fn_code<typename K::Fn>::value, lower(b, m.in1_, axes));
  } else if constexpr (node_kind<Op>::value == 3) {
    using K = node_kind<Op>;
    static_assert(sizeof(typename K::Mirror) == sizeof(Op), "PermuteOp layout changed: update the mirror");
    const auto &m = reinterpret_cast<const typename K::Mirror &>(op);
    int child[MXB_MAX_RANK];
    bool seen[MXB_MAX_RANK] = {false};
    for (int i = 0; i < Op::Rank(); ++i) {
      const int d = m.dims_[i];  // output dim i is input dim dims_[i]  (operators/permute.h:276)
      if (d < 0 || d >= Op::Rank() || seen[d] || op.Size(i) != m.op_.Size(d)) { b.ok = false; return 0; }
      seen[d] = true;
      child[d] = axes[i];
    }
    return lower(b, m.op_, child);
  } else if constexpr (node_kind<Op>::value == 4) {
    using K = node_kind<Op>;
    static_assert(sizeof(typename K::Mirror) == sizeof(Op), "CloneOp layout changed: update the mirror");
    const auto &m = reinterpret_cast<const typename K::Mirror &>(op);
    constexpr int RA = remove_cvref_t<decltype(m.op_)>::Rank();
    int child[MXB_MAX_RANK > 0 ? MXB_MAX_RANK : 1];
    for (int d = 0; d < RA; ++d) {
      const index_t od = m.dims_[d];  // input dim d is output dim dims_[d]  (operators/clone.h:137)
      if (od < 0 || od >= K::crank) { b.ok = false; return 0; }
      child[d] = axes[od];
    }
    return lower(b, m.op_, child);
  } else if constexpr (node_kind<Op>::value == 5) {
    // diag(op, k): the result's last dim walks rows AND columns of the operand (operators/diag.h:145-190), i.e. both
    // of the operand's last two dims map to the same root dim and their strides add up.  k != 0 offsets the data
    // pointer, which only a tensor operand has.
    using K = node_kind<Op>;
    static_assert(sizeof(typename K::Mirror) == sizeof(Op), "DiagOp layout changed: update the mirror");
    const auto &m = reinterpret_cast<const typename K::Mirror &>(op);
    using Child = remove_cvref_t<decltype(m.op_)>;
    constexpr int RA = Child::Rank();
    int child[MXB_MAX_RANK];
    for (int d = 0; d < RA - 2; ++d) child[d] = axes[d];
    child[RA - 2] = child[RA - 1] = axes[RA - 2];
    if (m.k_ == 0) return lower(b, m.op_, child);
    {  // the reference's size rule for k != 0 (diag.h:262-267) runs off a non-square matrix: leave that to the reference
      const index_t rows = m.op_.Size(RA - 2), cols = m.op_.Size(RA - 1);
      const index_t valid = m.k_ > 0 ? cuda::std::min(rows, cols - m.k_) : cuda::std::min(rows + m.k_, cols);
      if (op.Size(RA - 2) > valid) { b.ok = false; return 0; }
    }
    if constexpr (is_tensor_view_v<Child> || matx::is_tensor_impl_v<Child>) {
      mxb_leaf_t lf;
      std::memset(&lf, 0, sizeof lf);
      const index_t off = m.k_ > 0 ? m.k_ * m.op_.Stride(RA - 1) : -m.k_ * m.op_.Stride(RA - 2);
      lf.data = m.op_.Data() + off;
      lf.dtype = dtype_of<typename Child::value_type>::value;
      for (int d = 0; d < RA; ++d) lf.stride[child[d]] += m.op_.Stride(d);
      return b.leaf(lf);
    } else {
      b.ok = false;
      return 0;
    }
  } else if constexpr (node_kind<Op>::value == 6) {
    // isclose: int(|a - b| <= atol + rtol * |b|), tolerances in the operands' inner type (operators/isclose.h)
    using K = node_kind<Op>;
    static_assert(sizeof(typename K::Mirror) == sizeof(Op), "IsCloseOp layout changed: update the mirror");
    const auto &m = reinterpret_cast<const typename K::Mirror &>(op);
    constexpr int R = Op::Rank(), RA = rank_of_operand<remove_cvref_t<decltype(m.op1_)>>(), RB = rank_of_operand<remove_cvref_t<decltype(m.op2_)>>();
    const int a = lower(b, m.op1_, axes + (R - RA));
    const int c = lower(b, m.op2_, axes + (R - RB));
    constexpr int tol_dt = dtype_of<typename K::inner>::value;
    const int diff = b.node(MXB_OP_ABS, b.node(MXB_OP_SUB, a, c));
    const int bound = b.node(MXB_OP_ADD, b.constant(static_cast<double>(m.atol_), 0.0, tol_dt),
                             b.node(MXB_OP_MUL, b.constant(static_cast<double>(m.rtol_), 0.0, tol_dt), b.node(MXB_OP_ABS, c)));
    return b.node(MXB_OP_CAST, b.node(MXB_OP_LE, diff, bound), -1, MXB_I32);
  } else {
    b.ok = false;
    return 0;
  }
}

template <class Op> bool lower_root(Builder &b, const Op &op) {
  constexpr int R = rank_of_operand<remove_cvref_t<Op>>();
  if constexpr (R > MXB_MAX_RANK) return false;
  b.e.rank = R;
  int axes[MXB_MAX_RANK > 0 ? MXB_MAX_RANK : 1];
  for (int d = 0; d < R; ++d) { axes[d] = d; if constexpr (R > 0) b.e.size[d] = op.Size(d); }
  b.e.root = lower(b, op, axes);
  return b.ok;
}

template <class T> bool out_desc(const T &t, mxb_out_t &o) {
  std::memset(&o, 0, sizeof o);
  if constexpr (!(is_tensor_view_v<T> || matx::is_tensor_impl_v<T>)) return false;
  else {
    if (dtype_of<typename T::value_type>::value < 0 || T::Rank() > MXB_MAX_RANK) return false;
    o.data = const_cast<void *>(static_cast<const void *>(t.Data()));
    o.dtype = dtype_of<typename T::value_type>::value;
    o.rank = T::Rank();
    if constexpr (T::Rank() > 0) {   // tensor_t<T, 0>::Stride is a static_assert (core/tensor.h:1026)
      for (int d = 0; d < T::Rank(); ++d) { o.size[d] = t.Size(d); o.stride[d] = t.Stride(d); }
    }
    return true;
  }
}

// status -> reference error convention; returns true when the caller should fall back to the reference path
inline bool check_or_fallback(int st) {
  if (st == MXB_OK) return false;
  if (st == MXB_ERR_NOT_SUPPORTED) return true;
  if (st == MXB_ERR_JIT) {
    // no ahead-of-time kernel for this expression and NVRTC is unavailable / failed: the statement is still valid
    // MatX, so it runs on the reference path; say so once, the fallback is a performance cliff
    static bool warned = false;
    if (!warned) {
      warned = true;
      fprintf(stderr, "matx_b200: falling back to the reference kernels for expressions without an ahead-of-time kernel (%s)\n", mxb_last_error());
    }
    return true;
  }
  const std::string msg = std::string("libmatx_b200: ") + mxb_last_error();
  if (st == MXB_ERR_SIZE) { MATX_THROW(matxInvalidSize, msg); }
  if (st == MXB_ERR_INVALID) { MATX_THROW(matxInvalidParameter, msg); }
  MATX_THROW(matxCudaError, msg);
  return false;
}

}  // namespace b200_detail

class b200Executor : public cudaExecutor {
 public:
  // explicit: the reference itself calls `sum_impl(dest, in, stream)` (transforms/reduce.h:280) and relies on
  // cudaExecutor being the only type a bare cudaStream_t converts to
  explicit b200Executor(cudaStream_t stream, bool profiling = false) : cudaExecutor(stream, profiling) { init(); }
  b200Executor() : cudaExecutor() { init(); }

  mxb_handle_t handle() const { return h_.get(); }
  const char *last_kernel() const { return mxb_last_kernel(h_.get()); }   // "" when the last statement fell back
  long long native_launches() const { return mxb_launch_count(h_.get()); }

  // `(tensor = expression).run(exec)` lands here (operators/base_operator.h:255-262)
  template <typename Op> void Exec(const Op &op) const {
    if constexpr (is_matx_set_op<Op>()) {
      using Rhs = typename Op::op_type;
      if constexpr (b200_detail::lowerable<Rhs>()) {
        Op copy = op;  // get_lhs() / get_rhs() are non-const accessors (operators/set.h:169-175)
        b200_detail::Builder b;
        mxb_out_t out;
        if (b200_detail::out_desc(copy.get_lhs(), out) && lower_for_lhs(b, copy.get_rhs(), out)) {
          if (!b200_detail::check_or_fallback(mxb_elementwise(h_.get(), &b.e, &out))) return;
        }
      }
    }
    cudaExecutor::Exec(op);  // reference generic kernels, same stream
  }

  // one reduction statement; returns false when the reference path has to take it
  template <class Out, class In> bool reduce(int op, Out &dest, const In &in, int ddof = 1) const {
    return reduce_idx<Out, Out, In>(op, dest, nullptr, in, ddof);
  }
  // softmax over the trailing n_axes dims of `in` (already permuted so that the softmax axes are innermost); `dest` is
  // walked in the same permuted order
  template <class Out, class In> bool softmax_trailing(Out &dest, const In &in, int n_axes) const {
    if constexpr (!b200_detail::lowerable<In>()) return false;
    else {
      b200_detail::Builder b;
      mxb_out_t out;
      if (!b200_detail::out_desc(dest, out)) return false;
      if (!b200_detail::lower_root(b, in)) return false;
      return !b200_detail::check_or_fallback(mxb_softmax(h_.get(), &b.e, n_axes, &out));
    }
  }
  template <class Out, class In> bool cumsum(Out &dest, const In &in) const {
    if constexpr (!b200_detail::lowerable<In>()) return false;
    else {
      b200_detail::Builder b;
      mxb_out_t out;
      if (!b200_detail::out_desc(dest, out)) return false;
      if (!b200_detail::lower_root(b, in)) return false;
      return !b200_detail::check_or_fallback(mxb_cumsum(h_.get(), &b.e, &out));
    }
  }
  // find / find_idx with one of the reference's selection functors (LT / GT / EQ / NEQ / LTE / GTE, cub.h:2521-2588)
  template <class Out, class Cnt, class In, class T> bool find(Out &dest, Cnt &num_found, const In &in, int sel_op, T thr, bool want_idx) const {
    if constexpr (!b200_detail::lowerable<In>() || Out::Rank() != 1 || Cnt::Rank() != 0 ||
                  !std::is_same_v<typename Cnt::value_type, int> || !std::is_arithmetic_v<T>) return false;
    else {
      b200_detail::Builder b;
      mxb_out_t out, cnt;
      if (!b200_detail::out_desc(dest, out)) return false;
      if (!b200_detail::out_desc(num_found, cnt)) return false;
      if (!b200_detail::lower_root(b, in)) return false;
      return !b200_detail::check_or_fallback(mxb_find(h_.get(), &b.e, sel_op, static_cast<double>(thr), &out, &cnt, want_idx ? 1 : 0));
    }
  }
  template <class Out, class Idx, class In> bool reduce_idx(int op, Out &dest, Idx *idest, const In &in, int ddof) const {
    if constexpr (!b200_detail::lowerable<In>()) return false;
    else {
      b200_detail::Builder b;
      mxb_out_t out, iout;
      if (!b200_detail::out_desc(dest, out)) return false;
      if (idest && !b200_detail::out_desc(*idest, iout)) return false;
      if (!b200_detail::lower_root(b, in)) return false;
      constexpr int n_reduce = In::Rank() - Out::Rank();
      return !b200_detail::check_or_fallback(mxb_reduce(h_.get(), op, &b.e, n_reduce, &out, idest ? &iout : nullptr, ddof));
    }
  }

 private:
  void init() {
    mxb_handle_t h = nullptr;
    const int st = mxb_create(&h, reinterpret_cast<void *>(getStream()));
    if (st != MXB_OK) { MATX_THROW(matxCudaError, std::string("libmatx_b200: ") + mxb_last_error()); }
    h_ = std::shared_ptr<mxb_context>(h, [](mxb_context *p) { mxb_destroy(p); });
  }
  // rhs of lower rank than the lhs broadcasts over the leading dims of the lhs
  template <class Rhs> static bool lower_for_lhs(b200_detail::Builder &b, const Rhs &rhs, const mxb_out_t &out) {
    constexpr int R = b200_detail::rank_of_operand<remove_cvref_t<Rhs>>();
    if (R > out.rank) return false;
    b.e.rank = out.rank;
    int axes[MXB_MAX_RANK > 0 ? MXB_MAX_RANK : 1];
    for (int d = 0; d < out.rank; ++d) b.e.size[d] = out.size[d];
    for (int d = 0; d < R; ++d) axes[d] = out.rank - R + d;
    b.e.root = b200_detail::lower(b, rhs, axes);
    return b.ok;
  }
  std::shared_ptr<mxb_context> h_;
};

// ---- the transform seam: overloads found by ADL from SumOp::Exec etc. -----------------------------------------------
#define MXB_SHIM_REDUCE(NAME, CODE)                                                                           \
  template <typename OutType, typename InType>                                                                \
  void NAME##_impl(OutType dest, const InType &in, const b200Executor &exec) {                                \
    if (!exec.reduce(CODE, dest, in)) NAME##_impl(dest, in, static_cast<const cudaExecutor &>(exec));         \
  }                                                                                                           \
  template <typename OutType, typename InType>                                                                \
  void NAME##_impl(OutType dest, const InType &in, b200Executor &exec) {                                      \
    NAME##_impl(dest, in, static_cast<const b200Executor &>(exec));                                           \
  }
MXB_SHIM_REDUCE(sum, MXB_RED_SUM)
MXB_SHIM_REDUCE(mean, MXB_RED_MEAN)
MXB_SHIM_REDUCE(prod, MXB_RED_PROD)
MXB_SHIM_REDUCE(max, MXB_RED_MAX)
MXB_SHIM_REDUCE(min, MXB_RED_MIN)
MXB_SHIM_REDUCE(any, MXB_RED_ANY)
MXB_SHIM_REDUCE(all, MXB_RED_ALL)
#undef MXB_SHIM_REDUCE

#define MXB_SHIM_ARGREDUCE(NAME, CODE)                                                                        \
  template <typename OutType, typename TensorIndexType, typename InType>                                      \
  void NAME##_impl(OutType dest, TensorIndexType &idest, const InType &in, const b200Executor &exec) {        \
    if (!exec.reduce_idx(CODE, dest, &idest, in, 1)) NAME##_impl(dest, idest, in, static_cast<const cudaExecutor &>(exec)); \
  }                                                                                                           \
  template <typename OutType, typename TensorIndexType, typename InType>                                      \
  void NAME##_impl(OutType dest, TensorIndexType &idest, const InType &in, b200Executor &exec) {              \
    NAME##_impl(dest, idest, in, static_cast<const b200Executor &>(exec));                                    \
  }
MXB_SHIM_ARGREDUCE(argmax, MXB_RED_ARGMAX)
MXB_SHIM_ARGREDUCE(argmin, MXB_RED_ARGMIN)
#undef MXB_SHIM_ARGREDUCE

// argminmax (transforms/reduce.h:1090-1109): min and max with their indices.  Served as argmin + argmax — two
// single-pass launches at the HBM roofline each (the reference's dual-arg CUB path first materialises operator inputs
// and carries 32-byte tuples); a fused dual-arg operator is listed under "next" in DESIGN.md.
template <typename OutType, typename TensorIndexType, typename InType>
void argminmax_impl(OutType destmin, TensorIndexType &idestmin, OutType destmax, TensorIndexType &idestmax, const InType &in,
                    const b200Executor &exec) {
  if (exec.reduce_idx(MXB_RED_ARGMIN, destmin, &idestmin, in, 1) && exec.reduce_idx(MXB_RED_ARGMAX, destmax, &idestmax, in, 1)) return;
  argminmax_impl(destmin, idestmin, destmax, idestmax, in, static_cast<const cudaExecutor &>(exec));
}
template <typename OutType, typename TensorIndexType, typename InType>
void argminmax_impl(OutType destmin, TensorIndexType &idestmin, OutType destmax, TensorIndexType &idestmax, const InType &in,
                    b200Executor &exec) {
  argminmax_impl(destmin, idestmin, destmax, idestmax, in, static_cast<const b200Executor &>(exec));
}

// var / stdd are generic over the executor in the reference (transforms/reduce.h:1406-1479); these are more specialised
#define MXB_SHIM_VAR(NAME, CODE)                                                                              \
  template <typename OutType, typename InType>                                                                \
  void NAME##_impl(OutType dest, const InType &in, const b200Executor &exec, int ddof = 1) {                  \
    if (!exec.reduce(CODE, dest, in, ddof)) {                                                                 \
      cudaExecutor ref{exec.getStream()};                                                                     \
      NAME##_impl(dest, in, ref, ddof);                                                                       \
    }                                                                                                         \
  }                                                                                                           \
  template <typename OutType, typename InType>                                                                \
  void NAME##_impl(OutType dest, const InType &in, b200Executor &exec, int ddof = 1) {                        \
    NAME##_impl(dest, in, static_cast<const b200Executor &>(exec), ddof);                                     \
  }
MXB_SHIM_VAR(var, MXB_RED_VAR)
MXB_SHIM_VAR(stdd, MXB_RED_STDD)
#undef MXB_SHIM_VAR

// softmax (transforms/reduce.h:362-445).  NOTE the reference's SoftmaxOp::Exec hands softmax_impl the bare STREAM
// (operators/softmax.h:105-108), so `(out = softmax(x)).run(exec)` cannot be told apart by executor type and keeps
// running the reference's three passes; these overloads are for callers of softmax_impl and for an overlay of
// operators/softmax.h that passes `ex` instead of `ex.getStream()`.
template <typename OutType, typename InType>
void softmax_impl(OutType dest, const InType &in, const b200Executor &exec) {
  if (!exec.softmax_trailing(dest, in, InType::Rank())) softmax_impl(dest, in, exec.getStream());
}
template <typename OutType, typename InType, typename PermDims>
void softmax_impl(OutType dest, const InType &in, PermDims dims, const b200Executor &exec) {
  static_assert(OutType::Rank() == InType::Rank(), "softmax output rank must equal input rank");
  if constexpr (is_tensor_view_v<OutType>) {
    const auto perm = detail::getPermuteDims<InType::Rank()>(dims);   // batch dims first, softmax axes last (core/utils.h:96-127)
    auto pdest = dest.Permute(perm);
    if (exec.softmax_trailing(pdest, permute(in, perm), static_cast<int>(dims.size()))) return;
  }
  softmax_impl(dest, in, dims, exec.getStream());
}

// cumsum (transforms/cub.h:2367-2395): CumsumOp::Exec calls `cumsum_impl(out, a_, ex)` (operators/cumsum.h:293-296), so
// `(out = cumsum(x)).run(exec)` lands here.  One launch for all rows (the reference launches CUB once per row).
template <typename OutputTensor, typename InputOperator>
void cumsum_impl(OutputTensor &a_out, const InputOperator &a, const b200Executor &exec) {
  if (!exec.cumsum(a_out, a)) cumsum_impl(a_out, a, static_cast<const cudaExecutor &>(exec));
}
template <typename OutputTensor, typename InputOperator>
void cumsum_impl(OutputTensor &a_out, const InputOperator &a, b200Executor &exec) {
  cumsum_impl(a_out, a, static_cast<const b200Executor &>(exec));
}

// find / find_idx (transforms/cub.h:2609-2625,2705-2721): FindOp::Exec / FindIdxOp::Exec call
// `find_impl(out, num_found, a_, sel_, ex)` (operators/find.h:89, find_idx.h:89), so
// `(mtie(out, num_found) = find(x, GT{0.5f})).run(exec)` lands here when the functor is one of the reference's six
// comparison structs (their threshold is the public member c_); any other callable keeps the reference path.
namespace b200_detail {
template <class S> struct sel_code { static constexpr int value = -1; };
template <class T> struct sel_code<LT<T>> { static constexpr int value = MXB_SEL_LT; };
template <class T> struct sel_code<GT<T>> { static constexpr int value = MXB_SEL_GT; };
template <class T> struct sel_code<EQ<T>> { static constexpr int value = MXB_SEL_EQ; };
template <class T> struct sel_code<NEQ<T>> { static constexpr int value = MXB_SEL_NEQ; };
template <class T> struct sel_code<LTE<T>> { static constexpr int value = MXB_SEL_LTE; };
template <class T> struct sel_code<GTE<T>> { static constexpr int value = MXB_SEL_GTE; };
}  // namespace b200_detail
template <typename SelectType, typename CountTensor, typename OutputTensor, typename InputOperator>
void find_impl(OutputTensor &a_out, CountTensor &num_found, const InputOperator &a, SelectType sel, const b200Executor &exec) {
  bool done = false;
  if constexpr (b200_detail::sel_code<SelectType>::value >= 0)
    done = exec.find(a_out, num_found, a, b200_detail::sel_code<SelectType>::value, sel.c_, false);
  if (!done) find_impl(a_out, num_found, a, sel, static_cast<const cudaExecutor &>(exec));
}
template <typename SelectType, typename CountTensor, typename OutputTensor, typename InputOperator>
void find_impl(OutputTensor &a_out, CountTensor &num_found, const InputOperator &a, SelectType sel, b200Executor &exec) {
  find_impl(a_out, num_found, a, sel, static_cast<const b200Executor &>(exec));
}
template <typename SelectType, typename CountTensor, typename OutputTensor, typename InputOperator>
void find_idx_impl(OutputTensor &a_out, CountTensor &num_found, const InputOperator &a, SelectType sel, const b200Executor &exec) {
  bool done = false;
  if constexpr (b200_detail::sel_code<SelectType>::value >= 0)
    done = exec.find(a_out, num_found, a, b200_detail::sel_code<SelectType>::value, sel.c_, true);
  if (!done) find_idx_impl(a_out, num_found, a, sel, static_cast<const cudaExecutor &>(exec));
}
template <typename SelectType, typename CountTensor, typename OutputTensor, typename InputOperator>
void find_idx_impl(OutputTensor &a_out, CountTensor &num_found, const InputOperator &a, SelectType sel, b200Executor &exec) {
  find_idx_impl(a_out, num_found, a, sel, static_cast<const b200Executor &>(exec));
}

// allclose (transforms/reduce.h:1321-1331): all(isclose(in1, in2, rtol, atol)) into a rank-0 int tensor, one launch
template <typename OutType, typename InType1, typename InType2>
void allclose(OutType dest, const InType1 &in1, const InType2 &in2, double rtol, double atol, const b200Executor &exec) {
  static_assert(OutType::Rank() == 0, "allclose output must be rank 0");
  if (!exec.reduce(MXB_RED_ALL, dest, isclose(in1, in2, rtol, atol)))
    allclose(dest, in1, in2, rtol, atol, static_cast<const cudaExecutor &>(exec));
}
template <typename OutType, typename InType1, typename InType2>
void allclose(OutType dest, const InType1 &in1, const InType2 &in2, double rtol, double atol, b200Executor &exec) {
  allclose(dest, in1, in2, rtol, atol, static_cast<const b200Executor &>(exec));
}

}  // namespace matx
