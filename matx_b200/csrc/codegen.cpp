// codegen.cpp — lowers an expression program (mxb_expr_t, SSA over leaves / constants) to the device
// functor `struct E_<hash>` that the kernel skeletons in mxb_device.cuh are instantiated with.
//
// The same generator feeds both builds: the AOT tool (gen_main.cpp) prints the functors of the named
// programs into .cu files that nvcc compiles into the library, and jit.cpp hands the text to NVRTC for
// any other program.  The typing rules restate what the reference gets from C++ on its functors
// (operators/scalar_ops.h:434-503, operators/binary_operators.h:91-381, unary_operators.h:58-):
// usual arithmetic conversions between operand types, comparisons / logic yield bool, abs / abs2 /
// real / imag of a complex yield its real type.  16-bit float leaves are widened to fp32 on load and
// all arithmetic on them is fp32 (the north star's "fp32 accumulation for fp16/bf16 inputs").
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <sstream>
#include <unordered_map>

#include "mxb_internal.h"

namespace mxbh {

int dtype_bytes(int d) {
  switch (d) {
    case MXB_F32: return 4;
    case MXB_F64: return 8;
    case MXB_BF16: return 2;
    case MXB_F16: return 2;
    case MXB_C64: return 8;
    case MXB_I32: return 4;
    case MXB_I64: return 8;
    case MXB_U8: return 1;
  }
  return 0;
}
const char *dtype_name(int d) {
  static const char *n[] = {"f32", "f64", "bf16", "f16", "c64", "i32", "i64", "u8"};
  return (d >= 0 && d < MXB_DTYPE_COUNT) ? n[d] : "?";
}
const char *dtype_ctype(int d) {
  static const char *n[] = {"float", "double", "__nv_bfloat16", "__half", "mxb::cfloat", "int", "mxb::i64", "unsigned char"};
  return (d >= 0 && d < MXB_DTYPE_COUNT) ? n[d] : "?";
}
const char *reduce_op_name(int op) {
  static const char *n[] = {"sum", "mean", "var", "stdd", "max", "min", "argmax", "argmin", "any", "all", "prod"};
  if (op == KOP_LSE) return "lse";
  if (op == KOP_ARGMINMAX) return "argminmax";
  return (op >= 0 && op < MXB_RED_COUNT) ? n[op] : "?";
}
uint64_t fnv64(const std::string &s) {
  uint64_t h = 1469598103934665603ULL;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

namespace {

int compute_type(int storage) { return (storage == MXB_BF16 || storage == MXB_F16) ? MXB_F32 : storage; }
bool is_int(int t) { return t == MXB_I32 || t == MXB_I64 || t == MXB_U8; }
bool is_real_float(int t) { return t == MXB_F32 || t == MXB_F64; }
int rank_of(int t) {
  switch (t) {
    case MXB_U8: return 0;
    case MXB_I32: return 1;
    case MXB_I64: return 2;
    case MXB_F32: return 3;
    case MXB_F64: return 4;
  }
  return -1;
}
// usual arithmetic conversions; -1 = unsupported combination
int promote(int a, int b) {
  if (a == MXB_C64 || b == MXB_C64) {
    const int o = (a == MXB_C64) ? b : a;
    if (o == MXB_F64) return -1;  // complex<double> is not lowered
    return MXB_C64;
  }
  int r = rank_of(a) > rank_of(b) ? a : b;
  if (r == MXB_U8) r = MXB_I32;  // integral promotion
  return r;
}
// type a math function sees for an operand: integers go to double, as std:: overloads do
int math_type(int t) { return is_int(t) ? MXB_F64 : t; }

std::string val(int id) { return "t" + std::to_string(id); }
std::string as(int want, int have, const std::string &x) {
  if (want == have) return x;
  return std::string("mxb::cvt<") + dtype_ctype(want) + ">(" + x + ")";
}

struct UnaryFn { int opcode; const char *name; const char *ffn; const char *dfn; };
const UnaryFn kUnary[] = {
    {MXB_OP_SQRT, "sqrt", "sqrtf", "sqrt"},     {MXB_OP_LOG, "log", "mxb::f_log", "mxb::f_log"},       {MXB_OP_LOG2, "log2", "log2f", "log2"},
    {MXB_OP_LOG10, "log10", "log10f", "log10"}, {MXB_OP_SIN, "sin", "sinf", "sin"},       {MXB_OP_COS, "cos", "cosf", "cos"},
    {MXB_OP_TAN, "tan", "tanf", "tan"},         {MXB_OP_TANH, "tanh", "tanhf", "tanh"},   {MXB_OP_SINH, "sinh", "sinhf", "sinh"},
    {MXB_OP_COSH, "cosh", "coshf", "cosh"},     {MXB_OP_ASIN, "asin", "asinf", "asin"},   {MXB_OP_ACOS, "acos", "acosf", "acos"},
    {MXB_OP_ATAN, "atan", "atanf", "atan"},     {MXB_OP_FLOOR, "floor", "floorf", "floor"}, {MXB_OP_CEIL, "ceil", "ceilf", "ceil"},
    {MXB_OP_ROUND, "round", "roundf", "round"}, {MXB_OP_RSQRT, "rsqrt", "rsqrtf", "rsqrt"}, {MXB_OP_NORMCDF, "normcdf", "mxb::f_normcdf", "mxb::f_normcdf"},
};
const char *opcode_tag(int op) {
  switch (op) {
    case MXB_OP_LEAF: return "L";
    case MXB_OP_CONST: return "C";
    case MXB_OP_ADD: return "add";
    case MXB_OP_SUB: return "sub";
    case MXB_OP_MUL: return "mul";
    case MXB_OP_DIV: return "div";
    case MXB_OP_MOD: return "mod";
    case MXB_OP_POW: return "pow";
    case MXB_OP_MAX: return "max";
    case MXB_OP_MIN: return "min";
    case MXB_OP_LT: return "lt";
    case MXB_OP_GT: return "gt";
    case MXB_OP_LE: return "le";
    case MXB_OP_GE: return "ge";
    case MXB_OP_EQ: return "eq";
    case MXB_OP_NE: return "ne";
    case MXB_OP_AND: return "and";
    case MXB_OP_OR: return "or";
    case MXB_OP_ATAN2: return "atan2";
    case MXB_OP_NEG: return "neg";
    case MXB_OP_EXP: return "exp";
    case MXB_OP_ABS: return "abs";
    case MXB_OP_ABS2: return "abs2";
    case MXB_OP_CONJ: return "conj";
    case MXB_OP_REAL: return "real";
    case MXB_OP_IMAG: return "imag";
    case MXB_OP_NOT: return "not";
    case MXB_OP_ISNAN: return "isnan";
    case MXB_OP_ISINF: return "isinf";
    case MXB_OP_EXPJ: return "expj";
    case MXB_OP_CAST: return "cast";
  }
  for (const UnaryFn &u : kUnary)
    if (u.opcode == op) return u.name;
  return nullptr;
}
bool is_binary(int op) { return op >= MXB_OP_ADD && op <= MXB_OP_ATAN2; }

}  // namespace

// analyze_expr reads only the STRUCTURE of a program — node opcodes / operands / cast targets, leaf and constant
// dtypes, the root — never its sizes, strides, pointers or constant values, so its result is memoised on exactly those
// bytes: a repeated statement (every iteration of a user's loop) pays a hash lookup instead of regenerating the functor
// source (measured in plan-only mode: Black-Scholes 34 -> 9 us of host time per statement).  Failures are not cached.
static int analyze_expr_uncached(const mxb_expr_t *e, ExprInfo *info, std::string *err);
int analyze_expr(const mxb_expr_t *e, ExprInfo *info, std::string *err) {
  if (!e || e->n_nodes <= 0 || e->n_nodes > MXB_MAX_NODES || e->n_leaves < 0 || e->n_leaves > MXB_MAX_LEAVES || e->n_consts < 0 ||
      e->n_consts > MXB_MAX_CONSTS || e->rank < 0 || e->rank > MXB_MAX_RANK)
    return analyze_expr_uncached(e, info, err);   // malformed: the uncached path names the problem
  std::string key;
  key.reserve(16 + (size_t)e->n_nodes * sizeof(mxb_node_t) + (size_t)(e->n_leaves + e->n_consts) * 4);
  const int32_t head[4] = {e->n_nodes, e->n_leaves, e->n_consts, e->root};
  key.append((const char *)head, sizeof head);
  key.append((const char *)e->nodes, (size_t)e->n_nodes * sizeof(mxb_node_t));
  for (int k = 0; k < e->n_leaves; ++k) key.append((const char *)&e->leaves[k].dtype, 4);
  for (int k = 0; k < e->n_consts; ++k) key.append((const char *)&e->consts[k].dtype, 4);
  static std::mutex mu;
  static std::unordered_map<std::string, std::shared_ptr<const ExprInfo>> cache;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *info = *it->second; return MXB_OK; }
  }
  const int st = analyze_expr_uncached(e, info, err);
  if (st == MXB_OK) {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() < 4096) cache.emplace(std::move(key), std::make_shared<const ExprInfo>(*info));
  }
  return st;
}
static int analyze_expr_uncached(const mxb_expr_t *e, ExprInfo *info, std::string *err) {
  auto fail = [&](int st, const std::string &m) { if (err) *err = m; return st; };
  if (!e) return fail(MXB_ERR_INVALID, "null expression");
  if (e->rank < 0 || e->rank > MXB_MAX_RANK) return fail(MXB_ERR_INVALID, "expression rank out of range");
  if (e->n_nodes <= 0 || e->n_nodes > MXB_MAX_NODES) return fail(MXB_ERR_INVALID, "node count out of range");
  if (e->n_leaves < 0 || e->n_leaves > MXB_MAX_LEAVES) return fail(MXB_ERR_INVALID, "leaf count out of range");
  if (e->n_consts < 0 || e->n_consts > MXB_MAX_CONSTS) return fail(MXB_ERR_INVALID, "constant count out of range");
  if (e->root < 0 || e->root >= e->n_nodes) return fail(MXB_ERR_INVALID, "root id out of range");

  std::vector<int> type(e->n_nodes, -1);
  std::ostringstream sig, body;
  info->nleaf = e->n_leaves;
  info->max_leaf_bytes = 1;
  info->min_leaf_bytes = 16;
  for (int k = 0; k < e->n_leaves; ++k) {
    const int d = e->leaves[k].dtype;
    if (d < 0 || d >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "leaf dtype out of range");
    info->leaf_dtype[k] = d;
    if (dtype_bytes(d) > info->max_leaf_bytes) info->max_leaf_bytes = dtype_bytes(d);
    if (dtype_bytes(d) < info->min_leaf_bytes) info->min_leaf_bytes = dtype_bytes(d);
    sig << "l" << k << ":" << dtype_name(d) << ";";
  }
  if (e->n_leaves == 0) info->min_leaf_bytes = info->max_leaf_bytes = 4;

  for (int i = 0; i < e->n_nodes; ++i) {
    const mxb_node_t &n = e->nodes[i];
    const char *tag = opcode_tag(n.opcode);
    if (!tag) return fail(MXB_ERR_NOT_SUPPORTED, "opcode " + std::to_string(n.opcode) + " is not lowered");
    int t = -1;
    std::string rhs;
    if (n.opcode == MXB_OP_LEAF) {
      const int k = n.src[0];
      if (k < 0 || k >= e->n_leaves) return fail(MXB_ERR_INVALID, "leaf index out of range");
      t = compute_type(e->leaves[k].dtype);
      rhs = as(t, e->leaves[k].dtype, "r.x" + std::to_string(k) + ".v[v]");
      sig << i << "=L" << k << ";";
    } else if (n.opcode == MXB_OP_CONST) {
      const int k = n.src[0];
      if (k < 0 || k >= e->n_consts) return fail(MXB_ERR_INVALID, "constant index out of range");
      const int d = e->consts[k].dtype;
      if (d < 0 || d >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "constant dtype out of range");
      t = compute_type(d);
      const std::string ks = std::to_string(k);
      switch (t) {
        case MXB_F32: rhs = "c.fre[" + ks + "]"; break;
        case MXB_F64: rhs = "c.dre[" + ks + "]"; break;
        case MXB_C64: rhs = "mxb::cfloat(c.fre[" + ks + "], c.fim[" + ks + "])"; break;
        case MXB_I32: rhs = "(int)c.ire[" + ks + "]"; break;
        case MXB_I64: rhs = "c.ire[" + ks + "]"; break;
        case MXB_U8: rhs = "(unsigned char)c.ire[" + ks + "]"; break;
      }
      sig << i << "=C" << k << ":" << dtype_name(t) << ";";
    } else {
      const int a = n.src[0];
      if (a < 0 || a >= i) return fail(MXB_ERR_INVALID, "operand id must precede its use");
      const int ta = type[a];
      if (is_binary(n.opcode)) {
        const int b = n.src[1];
        if (b < 0 || b >= i) return fail(MXB_ERR_INVALID, "operand id must precede its use");
        const int tb = type[b];
        int pt = promote(ta, tb);
        if (pt < 0) return fail(MXB_ERR_NOT_SUPPORTED, "complex<double> arithmetic is not lowered");
        // complex (x) real keeps the real operand real (scalar multiply / divide, like cuda::std::complex)
        auto opnd = [&](int id, int tid) {
          if (pt == MXB_C64 && tid != MXB_C64) return as(MXB_F32, tid, val(id));
          return as(pt, tid, val(id));
        };
        const std::string A = opnd(a, ta), B = opnd(b, tb);
        switch (n.opcode) {
          case MXB_OP_ADD: t = pt; rhs = A + " + " + B; break;
          case MXB_OP_SUB: t = pt; rhs = A + " - " + B; break;
          case MXB_OP_MUL: t = pt; rhs = A + " * " + B; break;
          case MXB_OP_DIV: t = pt; rhs = A + " / " + B; break;
          case MXB_OP_MOD:
            if (pt == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, "mod of complex");
            t = pt; rhs = "mxb::f_mod(" + A + ", " + B + ")"; break;
          case MXB_OP_POW: {
            if (pt == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, "pow of complex");
            t = math_type(pt);
            rhs = "mxb::f_pow(" + as(t, ta, val(a)) + ", " + as(t, tb, val(b)) + ")";
            break;
          }
          case MXB_OP_ATAN2: {
            if (pt == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, "atan2 of complex");
            t = math_type(pt);
            rhs = std::string(t == MXB_F32 ? "atan2f(" : "atan2(") + as(t, ta, val(a)) + ", " + as(t, tb, val(b)) + ")";
            break;
          }
          case MXB_OP_MAX: case MXB_OP_MIN:
            if (pt == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, "max/min of complex");
            t = pt;
            rhs = std::string(n.opcode == MXB_OP_MAX ? "mxb::f_max<" : "mxb::f_min<") + dtype_ctype(pt) + ">(" + A + ", " + B + ")";
            break;
          case MXB_OP_LT: case MXB_OP_GT: case MXB_OP_LE: case MXB_OP_GE: {
            if (pt == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, "ordering of complex");
            const char *o = n.opcode == MXB_OP_LT ? " < " : n.opcode == MXB_OP_GT ? " > " : n.opcode == MXB_OP_LE ? " <= " : " >= ";
            t = MXB_U8; rhs = "(unsigned char)(" + A + o + B + ")"; break;
          }
          case MXB_OP_EQ: case MXB_OP_NE: {
            const std::string A2 = as(pt, ta, val(a)), B2 = as(pt, tb, val(b));
            t = MXB_U8; rhs = "(unsigned char)(" + A2 + (n.opcode == MXB_OP_EQ ? " == " : " != ") + B2 + ")"; break;
          }
          case MXB_OP_AND: case MXB_OP_OR:
            t = MXB_U8;
            rhs = "(unsigned char)(mxb::nonzero(" + val(a) + (n.opcode == MXB_OP_AND ? ") && mxb::nonzero(" : ") || mxb::nonzero(") + val(b) + "))";
            break;
          default: return fail(MXB_ERR_NOT_SUPPORTED, "binary opcode not lowered");
        }
        sig << i << "=" << tag << "(" << a << "," << b << ");";
      } else {
        const std::string X = val(a);
        bool done = false;
        for (const UnaryFn &u : kUnary) {
          if (u.opcode != n.opcode) continue;
          if (ta == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, std::string(u.name) + " of complex is not lowered");
          t = math_type(ta);
          rhs = std::string(t == MXB_F32 ? u.ffn : u.dfn) + "(" + as(t, ta, X) + ")";
          done = true;
        }
        if (!done) {
          switch (n.opcode) {
            case MXB_OP_NEG: t = (ta == MXB_U8) ? MXB_I32 : ta; rhs = "-" + as(t, ta, X); break;
            case MXB_OP_EXP:
              t = math_type(ta); rhs = "mxb::f_exp(" + as(t, ta, X) + ")"; break;
            case MXB_OP_ABS: t = (ta == MXB_C64) ? MXB_F32 : (ta == MXB_U8 ? MXB_I32 : ta); rhs = "mxb::f_abs(" + as(ta == MXB_U8 ? MXB_I32 : ta, ta, X) + ")"; break;
            case MXB_OP_ABS2: t = (ta == MXB_C64) ? MXB_F32 : (ta == MXB_U8 ? MXB_I32 : ta); rhs = "mxb::f_abs2(" + as(ta == MXB_U8 ? MXB_I32 : ta, ta, X) + ")"; break;
            case MXB_OP_CONJ: t = ta; rhs = "mxb::f_conj(" + X + ")"; break;
            case MXB_OP_REAL: t = (ta == MXB_C64) ? MXB_F32 : ta; rhs = "mxb::f_real(" + X + ")"; break;
            case MXB_OP_IMAG: t = (ta == MXB_C64) ? MXB_F32 : ta; rhs = "mxb::f_imag(" + X + ")"; break;
            case MXB_OP_NOT: t = MXB_U8; rhs = "(unsigned char)(!mxb::nonzero(" + X + "))"; break;
            case MXB_OP_ISNAN: t = MXB_U8; rhs = "(unsigned char)mxb::f_isnan(" + X + ")"; break;
            case MXB_OP_ISINF: t = MXB_U8; rhs = "(unsigned char)mxb::f_isinf(" + X + ")"; break;
            case MXB_OP_EXPJ:
              if (ta == MXB_C64) return fail(MXB_ERR_NOT_SUPPORTED, "expj of complex");
              t = MXB_C64; rhs = "mxb::f_expj(" + as(MXB_F32, ta, X) + ")"; break;
            case MXB_OP_CAST: {
              const int d = n.aux;
              if (d < 0 || d >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "cast dtype out of range");
              t = compute_type(d);
              // a cast to a 16-bit float rounds through that format, then the value lives on as fp32
              rhs = (d == t) ? as(t, ta, X) : as(t, d, as(d, ta, X));
              break;
            }
            default: return fail(MXB_ERR_NOT_SUPPORTED, "unary opcode not lowered");
          }
        }
        sig << i << "=" << tag << "(" << a;
        if (n.opcode == MXB_OP_CAST) sig << ":" << dtype_name(n.aux);
        sig << ");";
      }
    }
    type[i] = t;
    body << "    const " << dtype_ctype(t) << " " << val(i) << " = " << rhs << ";\n";
  }
  sig << "root=" << e->root;
  info->value_dtype = type[e->root];
  info->sig = sig.str();
  char nm[32];
  snprintf(nm, sizeof nm, "E_%016llx", (unsigned long long)fnv64(info->sig));
  info->name = nm;

  // ---- emit the functor ----
  std::ostringstream s;
  const int nl = e->n_leaves > 0 ? e->n_leaves : 1;
  s << "// " << info->sig << "\n";
  s << "struct " << info->name << " {\n";
  s << "  typedef " << dtype_ctype(info->value_dtype) << " value_type;\n";
  s << "  enum { NLEAF = " << e->n_leaves << ", NL = " << nl << " };\n";
  s << "  static __device__ __forceinline__ constexpr int leaf_bytes(int k) { return ";
  for (int k = 0; k < e->n_leaves; ++k) s << "k == " << k << " ? " << dtype_bytes(e->leaves[k].dtype) << " : ";
  s << "1; }\n";
  s << "  template <int V> struct Regs {";
  for (int k = 0; k < e->n_leaves; ++k) s << " mxb::Vec<" << dtype_ctype(e->leaves[k].dtype) << ", V> x" << k << ";";
  if (e->n_leaves == 0) s << " int unused_;";
  s << " };\n";
  s << "  template <int V, bool UNIT> static __device__ __forceinline__ void loadv(Regs<V> &r, const char *const *base, const mxb::i64 *inner, mxb::i64 j) {\n";
  for (int k = 0; k < e->n_leaves; ++k)
    s << "    mxb::ldleaf<" << dtype_ctype(e->leaves[k].dtype) << ", V, UNIT>(r.x" << k << ", base[" << k << "], j, inner[" << k << "]);\n";
  if (e->n_leaves == 0) s << "    (void)r; (void)base; (void)inner; (void)j;\n";
  s << "  }\n";
  // transposing family: leaf k comes from a shared-memory tile (bit k of ymask) or straight from global memory
  s << "  template <int V> static __device__ __forceinline__ void loadmix(Regs<V> &r, const char *const *gp, const mxb::i64 *ginner, const char *const *sp, unsigned ymask, bool vec, int n) {\n";
  for (int k = 0; k < e->n_leaves; ++k)
    s << "    if ((ymask >> " << k << ") & 1u) mxb::ldtile<" << dtype_ctype(e->leaves[k].dtype) << ", V>(r.x" << k << ", sp[" << k << "]); else mxb::ldrow<"
      << dtype_ctype(e->leaves[k].dtype) << ", V>(r.x" << k << ", gp[" << k << "], ginner[" << k << "], vec, n);\n";
  s << "    (void)r; (void)gp; (void)ginner; (void)sp; (void)ymask; (void)vec; (void)n;\n";
  s << "  }\n";
  s << "  template <int V> static __device__ __forceinline__ value_type eval(const Regs<V> &r, int v, const mxb::ConstDev &c) {\n";
  s << "    (void)r; (void)v; (void)c;\n";
  s << body.str();
  s << "    return " << val(e->root) << ";\n";
  s << "  }\n";
  // ---- paired evaluation (elements v and v+1 in one pass) on Blackwell's packed fp32 instructions ----
  // Pure-fp32 programs made of + - * / neg sqrt exp log normcdf abs get a second body over mxb::f2 (FFMA2 / FMUL2 /
  // FADD2 carry two lanes per issue slot; the hand-written log / normcdf are packed too).  a*b+c is contracted
  // explicitly (left product first, single-use products only), the way the scalar body is contracted by the compiler.
  {
    bool pair_ok = type[e->root] == MXB_F32 && e->n_leaves > 0;
    std::vector<int> uses(e->n_nodes, 0);
    for (int i = 0; i < e->n_nodes && pair_ok; ++i) {
      const mxb_node_t &n = e->nodes[i];
      if (type[i] != MXB_F32) { pair_ok = false; break; }
      switch (n.opcode) {
        case MXB_OP_LEAF: {
          const int d = e->leaves[n.src[0]].dtype;
          if (d != MXB_F32 && d != MXB_BF16 && d != MXB_F16) pair_ok = false;
          break;
        }
        case MXB_OP_CONST: break;
        case MXB_OP_ADD: case MXB_OP_SUB: case MXB_OP_MUL: case MXB_OP_DIV:
          if (type[n.src[0]] != MXB_F32 || type[n.src[1]] != MXB_F32) pair_ok = false;
          uses[n.src[0]]++; uses[n.src[1]]++;
          break;
        case MXB_OP_NEG: case MXB_OP_SQRT: case MXB_OP_EXP: case MXB_OP_LOG: case MXB_OP_NORMCDF: case MXB_OP_ABS:
          if (type[n.src[0]] != MXB_F32) pair_ok = false;
          uses[n.src[0]]++;
          break;
        default: pair_ok = false;
      }
    }
    s << "  enum { PAIR = " << (pair_ok ? 1 : 0) << " };\n";
    if (pair_ok) {
      std::vector<char> fused(e->n_nodes, 0);  // product folded into the add / sub that consumes it
      auto is_mul1 = [&](int id) { return e->nodes[id].opcode == MXB_OP_MUL && uses[id] == 1; };
      for (int i = 0; i < e->n_nodes; ++i) {
        const mxb_node_t &n = e->nodes[i];
        if (n.opcode != MXB_OP_ADD && n.opcode != MXB_OP_SUB) continue;
        if (is_mul1(n.src[0])) fused[n.src[0]] = 1;
        else if (is_mul1(n.src[1])) fused[n.src[1]] = 1;
      }
      s << "  template <int V> static __device__ __forceinline__ mxb::f2 eval2(const Regs<V> &r, int v, const mxb::ConstDev &c) {\n";
      s << "    (void)r; (void)v; (void)c;\n";
      for (int i = 0; i < e->n_nodes; ++i) {
        const mxb_node_t &n = e->nodes[i];
        if (fused[i]) continue;
        std::string rhs;
        const std::string A = val(n.src[0]), B = val(n.src[1]);
        switch (n.opcode) {
          case MXB_OP_LEAF: {
            const std::string x = "r.x" + std::to_string(n.src[0]);
            const int d = e->leaves[n.src[0]].dtype;
            rhs = "mxb::f2(" + as(MXB_F32, d, x + ".v[v]") + ", " + as(MXB_F32, d, x + ".v[v + 1]") + ")";
            break;
          }
          case MXB_OP_CONST: rhs = "mxb::f2(c.fre[" + std::to_string(n.src[0]) + "])"; break;
          case MXB_OP_ADD: case MXB_OP_SUB: {
            const bool sub = n.opcode == MXB_OP_SUB;
            if (fused[n.src[0]] && is_mul1(n.src[0])) {
              const mxb_node_t &m = e->nodes[n.src[0]];
              rhs = "mxb::fma2(" + val(m.src[0]) + ", " + val(m.src[1]) + ", " + (sub ? "mxb::neg2(" + B + ")" : B) + ")";
            } else if (fused[n.src[1]] && is_mul1(n.src[1])) {
              const mxb_node_t &m = e->nodes[n.src[1]];
              rhs = "mxb::fma2(" + (sub ? "mxb::neg2(" + val(m.src[0]) + ")" : val(m.src[0])) + ", " + val(m.src[1]) + ", " + A + ")";
            } else {
              rhs = A + (sub ? " - " : " + ") + B;
            }
            break;
          }
          case MXB_OP_MUL: rhs = A + " * " + B; break;
          case MXB_OP_DIV: rhs = A + " / " + B; break;
          case MXB_OP_NEG: rhs = "mxb::neg2(" + A + ")"; break;
          case MXB_OP_SQRT: rhs = "mxb::f_sqrt(" + A + ")"; break;
          case MXB_OP_EXP: rhs = "mxb::f_exp(" + A + ")"; break;
          case MXB_OP_LOG: rhs = "mxb::f_log(" + A + ")"; break;
          case MXB_OP_NORMCDF: rhs = "mxb::f_normcdf(" + A + ")"; break;
          case MXB_OP_ABS: rhs = "mxb::f_abs(" + A + ")"; break;
        }
        s << "    const mxb::f2 " << val(i) << " = " << rhs << ";\n";
      }
      s << "    return " << val(e->root) << ";\n";
      s << "  }\n";
    }
  }
  s << "};\n";
  info->src = s.str();
  return MXB_OK;
}

// ---------------------------------------------------------------------------------------------------
// kernel instances
// ---------------------------------------------------------------------------------------------------
int policy_vmax(const ExprInfo &info) {
  // Measured on B200 (profiles/r1_sweeps.md): for 2- and 4-byte elements LDG.128 per leaf is already at the HBM
  // ceiling and LDG.256 changes nothing; for 8-byte elements (complex<float>, double) 32-byte loads cut the index /
  // loop overhead per element and win 1-15 %.
  int v = (info.max_leaf_bytes >= 8 ? 32 : 16) / info.max_leaf_bytes;
  if (v < 1) v = 1;
  if (v > 8) v = 8;
  return v;
}
int policy_unroll(const ExprInfo &info, int V, int family) {
  const int bytes = V * info.max_leaf_bytes;  // widest load of one step
  if (family == FAM_RED_OUTER) return 4;
  if (family == FAM_RED_OUTER_TMA) return 1;
  if (family == FAM_VAR_REG || family == FAM_VAR_TMA || family == FAM_VAR_GROUP || family == FAM_SM_GROUP || family == FAM_SM_REG) return 1;
  if (family == FAM_EW_TR) return 1;
  if (family == FAM_SELECT) return 4;
  if (family == FAM_HIST) return 4;
  if (family == FAM_VAR_SMEM) return bytes >= 32 ? 4 : 8;
  if (family == FAM_EW) return bytes >= 32 ? (info.nleaf <= 2 ? 2 : 1) : (info.nleaf <= 2 ? 4 : 2);
  return bytes >= 32 ? 2 : (info.nleaf <= 2 ? 4 : 2);
}

std::string kernel_key(const ExprInfo &info, const KernelSpec &s) {
  std::ostringstream k;
  static const char *fam[] = {"red_inner", "red_outer", "var_smem", "ew", "var_reg", "var_tma", "var_group", "softmax_group", "softmax_reg", "ew_tr", "scan", "red_outer_tma", "select", "var_tma2", "hist"};
  k << fam[s.family] << "|" << info.name << "|" << (s.op >= 0 ? reduce_op_name(s.op) : "-") << "|" << dtype_name(s.out_dtype)
    << "|V" << s.V << "|U" << s.U << "|T" << s.team;
  if (s.minb > 0) k << "|M" << s.minb;
  return k.str();
}
std::string kernel_symbol(const std::string &key) {
  char nm[40];
  snprintf(nm, sizeof nm, "mxbk_%016llx", (unsigned long long)fnv64(key));
  return nm;
}

int kernel_wrapper_src(const ExprInfo &info, const KernelSpec &s, const std::string &symbol, std::string *out, std::string *err) {
  auto fail = [&](const std::string &m) { if (err) *err = m; return (int)MXB_ERR_NOT_SUPPORTED; };
  const std::string T = dtype_ctype(info.value_dtype);
  const std::string O = dtype_ctype(s.out_dtype);
  const bool cplx = info.value_dtype == MXB_C64;
  std::ostringstream k;
  std::string op;
  if (s.family == FAM_RED_INNER || s.family == FAM_RED_OUTER || s.family == FAM_RED_OUTER_TMA) {
    switch (s.op) {
      case MXB_RED_SUM: op = "mxb::OpSum<" + T + ">"; break;
      case MXB_RED_PROD: op = "mxb::OpProd<" + T + ">"; break;
      case MXB_RED_MAX: case MXB_RED_MIN:
        if (cplx) return fail("max/min of a complex expression (the reference rejects it too)");
        op = "mxb::OpExt<" + T + (s.op == MXB_RED_MAX ? ", true>" : ", false>"); break;
      case MXB_RED_ARGMAX: case MXB_RED_ARGMIN:
        if (cplx) return fail("argmax/argmin of a complex expression (the reference rejects it too)");
        op = "mxb::OpArg<" + T + (s.op == MXB_RED_ARGMAX ? ", true>" : ", false>"); break;
      case KOP_ARGMINMAX:
        if (cplx) return fail("argminmax of a complex expression (the reference rejects it too)");
        op = "mxb::OpArgMinMax<" + T + ">"; break;
      case MXB_RED_ANY: op = "mxb::OpLogic<" + T + ", true>"; break;
      case MXB_RED_ALL: op = "mxb::OpLogic<" + T + ", false>"; break;
      case MXB_RED_VAR:   // one-pass (mean, M2, n) states through the generic walkers
        if (!(info.value_dtype == MXB_F32 || cplx)) return fail("one-pass variance serves fp32 and complex<float>");
        op = "mxb::OpVar<" + T + ">"; break;
      case KOP_LSE:
        if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64)) return fail("softmax of a non-real-floating expression");
        op = "mxb::OpLse<" + T + ">"; break;
      default: return fail("reduce op has no kernel of its own");
    }
  }
  const std::string E = info.name;
  const std::string VU = std::to_string(s.V) + ", " + std::to_string(s.U);
  switch (s.family) {
    case FAM_RED_INNER:
      k << "extern \"C\" __global__ void __launch_bounds__(" << (s.minb > 0 ? "256, " + std::to_string(s.minb) : std::string(s.team == 0 ? "512" : "256")) << ") " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::reduce_inner_body<" << E << ", " << op << ", " << O << ", " << VU
        << ", " << s.team << ">(p); }\n";
      break;
    case FAM_RED_OUTER:
      // 3 CTAs per SM (<= 85 registers): at 2 the column walker has too few loads in flight (ncu: 112 registers,
      // 24 % warps active, 0.435 ms on config 5; 76 registers, 34 %, 0.350 ms before the split-R code joined this body)
      k << "extern \"C\" __global__ void __launch_bounds__(256, 3) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::reduce_outer_body<" << E << ", " << op << ", " << O << ", " << VU << ">(p); }\n";
      break;
    case FAM_RED_OUTER_TMA:
      if (info.nleaf != 1) return fail("red_outer_tma serves plain tensors only");
      // up to 512 consumer threads + the producer warp; the ring, not the register file, holds the bytes in flight
      k << "extern \"C\" __global__ void __launch_bounds__(544) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::reduce_outer_tma_body<" << dtype_ctype(info.leaf_dtype[0]) << ", " << op << ", " << O << ">(p); }\n";
      break;
    case FAM_VAR_SMEM:
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || cplx)) return fail("var of a non-floating expression");
      k << "extern \"C\" __global__ void __launch_bounds__(512) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::var_inner_smem_body<" << E << ", " << O << ", " << VU << ">(p); }\n";
      break;
    case FAM_VAR_REG:
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || cplx)) return fail("var of a non-floating expression");
      k << "extern \"C\" __global__ void __launch_bounds__(512) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::var_inner_reg_body<" << E << ", " << O << ", " << s.V << ", " << s.team << ">(p); }\n";
      break;
    case FAM_VAR_GROUP:
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || cplx)) return fail("var of a non-floating expression");
      k << "extern \"C\" __global__ void __launch_bounds__(256) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::var_group_body<" << E << ", " << O << ", " << s.V << ", " << s.team << ">(p); }\n";
      break;
    case FAM_SM_GROUP: case FAM_SM_REG:
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64)) return fail("softmax of a non-real-floating expression");
      // register budgets from ptxas (no spills): group 78 regs at 4 vectors per lane, CTA family <= 64 regs up to 1024 threads
      k << "extern \"C\" __global__ void __launch_bounds__(" << (s.family == FAM_SM_GROUP ? (s.team >= 4 ? "256, 3" : "256, 4") : "1024") << ") " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::" << (s.family == FAM_SM_GROUP ? "softmax_group_body<" : "softmax_reg_body<") << E << ", " << O
        << ", " << s.V << ", " << s.team << ">(p); }\n";
      break;
    case FAM_VAR_TMA:
      if (info.nleaf != 1) return fail("var_tma serves plain tensors only");
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || cplx)) return fail("var of a non-floating expression");
      k << "extern \"C\" __global__ void __launch_bounds__(1024) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::var_inner_tma_body<" << dtype_ctype(info.leaf_dtype[0]) << ", " << O << ", " << s.team << ">(p); }\n";
      break;
    case FAM_EW:
      // team = minimum resident CTAs per SM asked of the compiler (0 = no cap): arithmetic-heavy programs trade a few
      // registers for occupancy once the packed-fp32 body has taken them off the issue limit
      k << "extern \"C\" __global__ void __launch_bounds__(256" << (s.team > 0 ? ", " + std::to_string(s.team) : std::string()) << ") " << symbol
        << "(const __grid_constant__ mxb::EwParams p) { mxb::ew_body<" << E << ", " << O << ", " << VU << ">(p); }\n";
      break;
    case FAM_SCAN:
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || cplx || info.value_dtype == MXB_I32 || info.value_dtype == MXB_I64))
        return fail("cumsum of this value type is not lowered");
      k << "extern \"C\" __global__ void __launch_bounds__(256, " << (s.minb > 0 ? s.minb : 3) << ") " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::scan_inner_body<" << E << ", " << O << ", " << VU << ", " << s.team << ">(p); }\n";
      break;
    case FAM_EW_TR:
      if (s.V != 2 && s.V != 4 && s.V != 8) return fail("ew_tr moves 16-byte chunks of 2-, 4- or 8-byte elements");
      k << "extern \"C\" __global__ void __launch_bounds__(256) " << symbol
        << "(const __grid_constant__ mxb::EwParams p) { mxb::ew_tr_body<" << E << ", " << O << ", " << (16 / s.V) << ">(p); }\n";
      break;
    case FAM_SELECT:
      // team = kernel: 0 count (+ in-launch scan of the CTA totals), 1 scatter values, 2 scatter flat indices (the two-pass
      // pair for N-D views); 3 / 4 = the single-pass warp-tile kernel writing values / flat indices
      if (cplx || info.value_dtype == MXB_BF16 || info.value_dtype == MXB_F16) return fail("find / find_idx serve real value types");
      if (s.team < 0 || s.team > 4) return fail("select kernel out of range");
      if (s.team >= 3)
        // three CTAs per SM: 80 registers keep the ranking and the vector re-read of the tile being written out of local
        // memory, and the dense tiles (every element slot stored under a predicate) measured faster than at four
        k << "extern \"C\" __global__ void __launch_bounds__(256, 3) " << symbol
          << "(const __grid_constant__ mxb::EwParams p) { mxb::select1p_body<" << E << ", " << O << ", " << s.V << ", " << (s.team - 2) << ">(p); }\n";
      else
        k << "extern \"C\" __global__ void __launch_bounds__(256) " << symbol
          << "(const __grid_constant__ mxb::EwParams p) { mxb::select_body<" << E << ", " << O << ", " << s.V << ", " << s.team << ">(p); }\n";
      break;
    case FAM_HIST:
      if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || info.value_dtype == MXB_I32 || info.value_dtype == MXB_I64 || info.value_dtype == MXB_U8))
        return fail("hist serves real value types");
      k << "extern \"C\" __global__ void __launch_bounds__(256) " << symbol
        << "(const __grid_constant__ mxb::RedParams p) { mxb::hist_body<" << E << ", " << VU << ">(p); }\n";
      break;
    default: return fail("unknown kernel family");
  }
  *out = k.str();
  return MXB_OK;
}

}  // namespace mxbh
