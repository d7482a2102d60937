/*
 * matx_oracle.c — CPU restatement of the reference's algorithm for the reduce / fused-elementwise path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under matx_b200/ may call, link or load this file; it is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.  Parity pinning: this
 * restatement is checked against (a) the known-answer vectors of the reference's own tests
 * (test/00_operators/ReductionTests.cu, test/00_tensor/CUBTests.cu — see tests/test_oracle.py) and
 * (b) the reference's HostExecutor itself, compiled from /root/reference into oracle/_ref/ (bit-exact on
 * fp32 / fp64 / int sums, min/max, arg ops, any/all; tests/test_oracle_vs_ref.py and tests/golden/).
 *
 * What is restated (paths relative to the reference tree):
 *   - HostExecutor reductions: include/matx/transforms/reduce.h:315-336 (mean), 674-691 (sum),
 *     814-832 (max), 892-913 (argmax), 962-979 (min), 1042-1063 (argmin), 1199-1220 (any),
 *     1272-1293 (all), 1406-1444 (var: mean, then sum of pow(abs(x-mean),2), then /(N-ddof)),
 *     1474-1479 (stdd); the single-thread helpers they call, include/matx/transforms/host_algorithms.h:
 *     std::accumulate in the VALUE type, std::max_element / std::min_element (first occurrence wins),
 *     std::any_of / std::all_of.
 *   - the flat-offset walk of the collapsed, permuted input: RandomOperatorIterator + GetIdxFromAbs
 *     (include/matx/core/iterator.h:48-202, include/matx/core/operator_utils.h:207-232): element r of
 *     batch row b is the row-major (b, r) element of the view whose reduced dims are innermost.
 *   - arg-reduce index convention: absolute flat offset b*R + r (test/00_operators/ReductionTests.cu:1296-1311).
 *   - element functors: include/matx/operators/scalar_ops.h:434-503, scalar_internal.h:44-297
 *     (C++ usual arithmetic conversions; abs2 of complex = re*re + im*im; normcdf).
 *   - the elementwise executor HostExecutor::Exec: include/matx/executors/host.h:147-174 (flat loop, EPT 1).
 * 16-bit floats: the reference host path rounds every partial sum through bf16/fp16 (core/half.h:665-672),
 * which stagnates; `half_acc` selects that faithful behaviour, otherwise 16-bit
 * inputs are widened to fp32 once (the arithmetic the B200 path documents).  Parity for 16-bit floats is
 * UNPINNED by any reference test.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/matx_b200.h"

typedef struct {
  int t;        /* MXB_F32 / F64 / C64 / I32 / I64 / U8 (16-bit floats are widened to F32 at the leaf) */
  float f, fi;  /* F32, C64 */
  double d;     /* F64 */
  long long i;  /* I32 / I64 / U8 */
} val_t;

static float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f32_to_bf16(float f) { /* round to nearest even, as __float2bfloat16_rn */
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float f16_to_f32(uint16_t h) {
  uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 0x1f, m = h & 0x3ff, u;
  if (e == 0) {
    if (m == 0) u = s;
    else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } m &= 0x3ff; u = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13); }
  } else if (e == 31) u = s | 0x7f800000u | (m << 13);
  else u = s | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}
static uint16_t f32_to_f16(float f) {
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t s = (x >> 16) & 0x8000u; int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15; uint32_t m = x & 0x7fffffu;
  if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0));
  if (e >= 31) return (uint16_t)(s | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)s;
    m |= 0x800000u; uint32_t sh = (uint32_t)(14 - e); uint32_t r = m >> sh, rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (r & 1))) ++r;
    return (uint16_t)(s | r);
  }
  uint32_t r = ((uint32_t)e << 10) | (m >> 13), rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
  return (uint16_t)(s | r);
}

static int compute_type(int storage) { return (storage == MXB_BF16 || storage == MXB_F16) ? MXB_F32 : storage; }
static int rank_of(int t) { return t == MXB_U8 ? 0 : t == MXB_I32 ? 1 : t == MXB_I64 ? 2 : t == MXB_F32 ? 3 : t == MXB_F64 ? 4 : -1; }
static int is_int(int t) { return t == MXB_I32 || t == MXB_I64 || t == MXB_U8; }
static int promote(int a, int b) {
  if (a == MXB_C64 || b == MXB_C64) return MXB_C64;
  int r = rank_of(a) > rank_of(b) ? a : b;
  return r == MXB_U8 ? MXB_I32 : r;
}
static val_t conv(val_t x, int t) {
  if (x.t == t) return x;
  val_t r; memset(&r, 0, sizeof r); r.t = t;
  double src = x.t == MXB_F32 || x.t == MXB_C64 ? (double)x.f : x.t == MXB_F64 ? x.d : (double)x.i;
  switch (t) {
    case MXB_F32: r.f = x.t == MXB_F64 ? (float)x.d : is_int(x.t) ? (float)x.i : x.f; break;
    case MXB_F64: r.d = src; break;
    case MXB_C64: r.f = x.t == MXB_F64 ? (float)x.d : is_int(x.t) ? (float)x.i : x.f; r.fi = 0.f; break;
    case MXB_I32: r.i = is_int(x.t) ? (long long)(int32_t)x.i : (long long)(int32_t)src; break;
    case MXB_I64: r.i = is_int(x.t) ? x.i : (long long)src; break;
    case MXB_U8: r.i = is_int(x.t) ? (long long)(uint8_t)x.i : (long long)(uint8_t)src; break;
  }
  return r;
}
static int nonzero(val_t x) {
  switch (x.t) {
    case MXB_F32: return x.f != 0.f;
    case MXB_F64: return x.d != 0.0;
    case MXB_C64: return x.f != 0.f || x.fi != 0.f;
    default: return x.i != 0;
  }
}
static val_t mk_f(float f) { val_t r; memset(&r, 0, sizeof r); r.t = MXB_F32; r.f = f; return r; }
static val_t mk_d(double d) { val_t r; memset(&r, 0, sizeof r); r.t = MXB_F64; r.d = d; return r; }
static val_t mk_c(float a, float b) { val_t r; memset(&r, 0, sizeof r); r.t = MXB_C64; r.f = a; r.fi = b; return r; }
static val_t mk_i(int t, long long i) { val_t r; memset(&r, 0, sizeof r); r.t = t; r.i = t == MXB_I32 ? (long long)(int32_t)i : i; return r; }
static val_t mk_b(int b) { val_t r; memset(&r, 0, sizeof r); r.t = MXB_U8; r.i = b ? 1 : 0; return r; }

static val_t load_leaf(const mxb_leaf_t *lf, int64_t off) {
  const char *p = (const char *)lf->data;
  switch (lf->dtype) {
    case MXB_F32: return mk_f(((const float *)p)[off]);
    case MXB_F64: return mk_d(((const double *)p)[off]);
    case MXB_BF16: return mk_f(bf16_to_f32(((const uint16_t *)p)[off]));
    case MXB_F16: return mk_f(f16_to_f32(((const uint16_t *)p)[off]));
    case MXB_C64: return mk_c(((const float *)p)[2 * off], ((const float *)p)[2 * off + 1]);
    case MXB_I32: return mk_i(MXB_I32, ((const int32_t *)p)[off]);
    case MXB_I64: return mk_i(MXB_I64, ((const int64_t *)p)[off]);
    default: return mk_i(MXB_U8, ((const uint8_t *)p)[off]);
  }
}

static double normcdf_d(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

static val_t unary_math(int op, val_t a) {
  int t = is_int(a.t) ? MXB_F64 : a.t;  /* std:: math on an integer computes in double */
  a = conv(a, t);
  if (t == MXB_F32) {
    float x = a.f, r = 0.f;
    switch (op) {
      case MXB_OP_SQRT: r = sqrtf(x); break;
      case MXB_OP_RSQRT: r = (float)(1.0 / sqrt((double)x)); break;  /* scalar_internal.h:111-119 host branch */
      case MXB_OP_LOG: r = logf(x); break;
      case MXB_OP_LOG2: r = log2f(x); break;
      case MXB_OP_LOG10: r = log10f(x); break;
      case MXB_OP_SIN: r = sinf(x); break;
      case MXB_OP_COS: r = cosf(x); break;
      case MXB_OP_TAN: r = tanf(x); break;
      case MXB_OP_TANH: r = tanhf(x); break;
      case MXB_OP_SINH: r = sinhf(x); break;
      case MXB_OP_COSH: r = coshf(x); break;
      case MXB_OP_ASIN: r = asinf(x); break;
      case MXB_OP_ACOS: r = acosf(x); break;
      case MXB_OP_ATAN: r = atanf(x); break;
      case MXB_OP_FLOOR: r = floorf(x); break;
      case MXB_OP_CEIL: r = ceilf(x); break;
      case MXB_OP_ROUND: r = roundf(x); break;
      case MXB_OP_NORMCDF: r = (float)normcdf_d((double)x); break;
      case MXB_OP_EXP: r = expf(x); break;
    }
    return mk_f(r);
  }
  double x = a.d, r = 0.0;
  switch (op) {
    case MXB_OP_SQRT: r = sqrt(x); break;
    case MXB_OP_RSQRT: r = 1.0 / sqrt(x); break;
    case MXB_OP_LOG: r = log(x); break;
    case MXB_OP_LOG2: r = log2(x); break;
    case MXB_OP_LOG10: r = log10(x); break;
    case MXB_OP_SIN: r = sin(x); break;
    case MXB_OP_COS: r = cos(x); break;
    case MXB_OP_TAN: r = tan(x); break;
    case MXB_OP_TANH: r = tanh(x); break;
    case MXB_OP_SINH: r = sinh(x); break;
    case MXB_OP_COSH: r = cosh(x); break;
    case MXB_OP_ASIN: r = asin(x); break;
    case MXB_OP_ACOS: r = acos(x); break;
    case MXB_OP_ATAN: r = atan(x); break;
    case MXB_OP_FLOOR: r = floor(x); break;
    case MXB_OP_CEIL: r = ceil(x); break;
    case MXB_OP_ROUND: r = round(x); break;
    case MXB_OP_NORMCDF: r = normcdf_d(x); break;
    case MXB_OP_EXP: r = exp(x); break;
  }
  return mk_d(r);
}

static val_t arith(int op, val_t a, val_t b) {
  int pt = promote(a.t, b.t);
  if (pt == MXB_C64) {
    /* complex (x) real keeps the real operand real, as cuda::std::complex's mixed operators do */
    int ar = a.t != MXB_C64, br = b.t != MXB_C64;
    a = conv(a, ar ? MXB_F32 : MXB_C64); b = conv(b, br ? MXB_F32 : MXB_C64);
    float are = a.f, aim = ar ? 0.f : a.fi, bre = b.f, bim = br ? 0.f : b.fi;
    switch (op) {
      case MXB_OP_ADD: return mk_c(are + bre, aim + bim);
      case MXB_OP_SUB: return mk_c(are - bre, ar ? -bim : (br ? aim : aim - bim));
      case MXB_OP_MUL:
        if (br) return mk_c(are * bre, aim * bre);
        if (ar) return mk_c(are * bre, are * bim);
        return mk_c(are * bre - aim * bim, are * bim + aim * bre);
      case MXB_OP_DIV:
        if (br) return mk_c(are / bre, aim / bre);
        if (fabsf(bre) >= fabsf(bim)) { float r = bim / bre, d = bre + bim * r; return mk_c((are + aim * r) / d, (aim - are * r) / d); }
        else { float r = bre / bim, d = bre * r + bim; return mk_c((are * r + aim) / d, (aim * r - are) / d); }
    }
    return mk_c(0.f, 0.f);
  }
  a = conv(a, pt); b = conv(b, pt);
  switch (pt) {
    case MXB_F32:
      switch (op) {
        case MXB_OP_ADD: return mk_f(a.f + b.f);
        case MXB_OP_SUB: return mk_f(a.f - b.f);
        case MXB_OP_MUL: return mk_f(a.f * b.f);
        case MXB_OP_DIV: return mk_f(a.f / b.f);
        case MXB_OP_MOD: return mk_f(fmodf(a.f, b.f));
        case MXB_OP_MAX: return mk_f(a.f > b.f ? a.f : b.f);
        case MXB_OP_MIN: return mk_f(a.f < b.f ? a.f : b.f);
      }
      break;
    case MXB_F64:
      switch (op) {
        case MXB_OP_ADD: return mk_d(a.d + b.d);
        case MXB_OP_SUB: return mk_d(a.d - b.d);
        case MXB_OP_MUL: return mk_d(a.d * b.d);
        case MXB_OP_DIV: return mk_d(a.d / b.d);
        case MXB_OP_MOD: return mk_d(fmod(a.d, b.d));
        case MXB_OP_MAX: return mk_d(a.d > b.d ? a.d : b.d);
        case MXB_OP_MIN: return mk_d(a.d < b.d ? a.d : b.d);
      }
      break;
    default:
      switch (op) {
        case MXB_OP_ADD: return mk_i(pt, (long long)((unsigned long long)a.i + (unsigned long long)b.i));
        case MXB_OP_SUB: return mk_i(pt, (long long)((unsigned long long)a.i - (unsigned long long)b.i));
        case MXB_OP_MUL: return mk_i(pt, (long long)((unsigned long long)a.i * (unsigned long long)b.i));
        case MXB_OP_DIV: return mk_i(pt, b.i ? a.i / b.i : 0);
        case MXB_OP_MOD: return mk_i(pt, b.i ? a.i % b.i : 0);
        case MXB_OP_MAX: return mk_i(pt, a.i > b.i ? a.i : b.i);
        case MXB_OP_MIN: return mk_i(pt, a.i < b.i ? a.i : b.i);
      }
  }
  return mk_i(MXB_I32, 0);
}

static int compare(int op, val_t a, val_t b) {
  int pt = promote(a.t, b.t);
  a = conv(a, pt); b = conv(b, pt);
  if (pt == MXB_C64) { int eq = a.f == b.f && a.fi == b.fi; return op == MXB_OP_EQ ? eq : !eq; }
  double x, y;
  if (pt == MXB_F32) { x = a.f; y = b.f; } else if (pt == MXB_F64) { x = a.d; y = b.d; }
  else {
    switch (op) {
      case MXB_OP_LT: return a.i < b.i; case MXB_OP_GT: return a.i > b.i; case MXB_OP_LE: return a.i <= b.i;
      case MXB_OP_GE: return a.i >= b.i; case MXB_OP_EQ: return a.i == b.i; default: return a.i != b.i;
    }
  }
  switch (op) {
    case MXB_OP_LT: return x < y; case MXB_OP_GT: return x > y; case MXB_OP_LE: return x <= y;
    case MXB_OP_GE: return x >= y; case MXB_OP_EQ: return x == y; default: return x != y;
  }
}

/* value of the expression at the N-D index idx[] (reference: op.operator()(idx...), one element at a time) */
static val_t eval_expr(const mxb_expr_t *e, const int64_t *idx) {
  val_t v[MXB_MAX_NODES];
  for (int i = 0; i < e->n_nodes; ++i) {
    const mxb_node_t *n = &e->nodes[i];
    if (n->opcode == MXB_OP_LEAF) {
      const mxb_leaf_t *lf = &e->leaves[n->src[0]];
      int64_t off = 0;
      for (int d = 0; d < e->rank; ++d) off += idx[d] * lf->stride[d];
      v[i] = load_leaf(lf, off);
    } else if (n->opcode == MXB_OP_CONST) {
      const mxb_const_t *c = &e->consts[n->src[0]];
      int t = compute_type(c->dtype);
      switch (t) {
        case MXB_F32: v[i] = mk_f((float)c->re); break;
        case MXB_F64: v[i] = mk_d(c->re); break;
        case MXB_C64: v[i] = mk_c((float)c->re, (float)c->im); break;
        default: v[i] = mk_i(t, (long long)c->re); break;
      }
    } else if (n->opcode >= MXB_OP_ADD && n->opcode <= MXB_OP_ATAN2) {
      val_t a = v[n->src[0]], b = v[n->src[1]];
      switch (n->opcode) {
        case MXB_OP_ADD: case MXB_OP_SUB: case MXB_OP_MUL: case MXB_OP_DIV: case MXB_OP_MOD: case MXB_OP_MAX: case MXB_OP_MIN:
          v[i] = arith(n->opcode, a, b); break;
        case MXB_OP_POW: {
          int t = promote(a.t, b.t); if (is_int(t)) t = MXB_F64;
          a = conv(a, t); b = conv(b, t);
          v[i] = t == MXB_F32 ? mk_f(powf(a.f, b.f)) : mk_d(pow(a.d, b.d));
          break;
        }
        case MXB_OP_ATAN2: {
          int t = promote(a.t, b.t); if (is_int(t)) t = MXB_F64;
          a = conv(a, t); b = conv(b, t);
          v[i] = t == MXB_F32 ? mk_f(atan2f(a.f, b.f)) : mk_d(atan2(a.d, b.d));
          break;
        }
        case MXB_OP_AND: v[i] = mk_b(nonzero(a) && nonzero(b)); break;
        case MXB_OP_OR: v[i] = mk_b(nonzero(a) || nonzero(b)); break;
        default: v[i] = mk_b(compare(n->opcode, a, b)); break;
      }
    } else {
      val_t a = v[n->src[0]];
      switch (n->opcode) {
        case MXB_OP_NEG:
          if (a.t == MXB_F32) v[i] = mk_f(-a.f); else if (a.t == MXB_F64) v[i] = mk_d(-a.d);
          else if (a.t == MXB_C64) v[i] = mk_c(-a.f, -a.fi); else v[i] = mk_i(a.t == MXB_U8 ? MXB_I32 : a.t, -a.i);
          break;
        case MXB_OP_ABS:
          if (a.t == MXB_F32) v[i] = mk_f(fabsf(a.f)); else if (a.t == MXB_F64) v[i] = mk_d(fabs(a.d));
          else if (a.t == MXB_C64) v[i] = mk_f(hypotf(a.f, a.fi)); else v[i] = mk_i(a.t == MXB_U8 ? MXB_I32 : a.t, a.i < 0 ? -a.i : a.i);
          break;
        case MXB_OP_ABS2: /* scalar_internal.h:184-191 */
          if (a.t == MXB_F32) v[i] = mk_f(a.f * a.f); else if (a.t == MXB_F64) v[i] = mk_d(a.d * a.d);
          else if (a.t == MXB_C64) v[i] = mk_f(a.f * a.f + a.fi * a.fi); else v[i] = mk_i(a.t == MXB_U8 ? MXB_I32 : a.t, a.i * a.i);
          break;
        case MXB_OP_CONJ: v[i] = a; if (a.t == MXB_C64) v[i].fi = -a.fi; break;
        case MXB_OP_REAL: v[i] = a.t == MXB_C64 ? mk_f(a.f) : a; break;
        case MXB_OP_IMAG:
          if (a.t == MXB_C64) v[i] = mk_f(a.fi); else if (a.t == MXB_F32) v[i] = mk_f(0.f); else if (a.t == MXB_F64) v[i] = mk_d(0.0); else v[i] = mk_i(a.t, 0);
          break;
        case MXB_OP_NOT: v[i] = mk_b(!nonzero(a)); break;
        case MXB_OP_ISNAN: v[i] = mk_b(a.t == MXB_F32 ? isnan(a.f) : a.t == MXB_F64 ? isnan(a.d) : a.t == MXB_C64 ? (isnan(a.f) || isnan(a.fi)) : 0); break;
        case MXB_OP_ISINF: v[i] = mk_b(a.t == MXB_F32 ? isinf(a.f) : a.t == MXB_F64 ? isinf(a.d) : a.t == MXB_C64 ? (isinf(a.f) || isinf(a.fi)) : 0); break;
        case MXB_OP_EXPJ: { float x = conv(a, MXB_F32).f; v[i] = mk_c(cosf(x), sinf(x)); break; }
        case MXB_OP_EXP:
          if (a.t == MXB_C64) { float m = expf(a.f); v[i] = mk_c(m * cosf(a.fi), m * sinf(a.fi)); }
          else v[i] = unary_math(MXB_OP_EXP, a);
          break;
        case MXB_OP_CAST: {
          int t = compute_type(n->aux);
          val_t r = conv(a, t);
          if (n->aux == MXB_BF16) r.f = bf16_to_f32(f32_to_bf16(r.f));
          if (n->aux == MXB_F16) r.f = f16_to_f32(f32_to_f16(r.f));
          v[i] = r;
          break;
        }
        default: v[i] = unary_math(n->opcode, a); break;
      }
    }
  }
  return v[e->root];
}

static void store_val(void *base, int dtype, int64_t off, val_t x) {
  switch (dtype) {
    case MXB_F32: ((float *)base)[off] = conv(x, MXB_F32).f; break;
    case MXB_F64: ((double *)base)[off] = conv(x, MXB_F64).d; break;
    case MXB_BF16: ((uint16_t *)base)[off] = f32_to_bf16(conv(x, MXB_F32).f); break;
    case MXB_F16: ((uint16_t *)base)[off] = f32_to_f16(conv(x, MXB_F32).f); break;
    case MXB_C64: { val_t c = conv(x, MXB_C64); ((float *)base)[2 * off] = c.f; ((float *)base)[2 * off + 1] = c.fi; break; }
    case MXB_I32: ((int32_t *)base)[off] = (int32_t)conv(x, MXB_I32).i; break;
    case MXB_I64: ((int64_t *)base)[off] = conv(x, MXB_I64).i; break;
    default: ((uint8_t *)base)[off] = (uint8_t)conv(x, MXB_U8).i; break;
  }
}

static void unflatten(int64_t flat, int n, const int64_t *size, int64_t *idx) { /* GetIdxFromAbs, operator_utils.h:207-232 */
  for (int d = n - 1; d >= 0; --d) { idx[d] = size[d] ? flat % size[d] : 0; flat = size[d] ? flat / size[d] : 0; }
}

static int val_less(val_t a, val_t b) { /* operator< of the value type */
  if (a.t == MXB_F32) return a.f < b.f;
  if (a.t == MXB_F64) return a.d < b.d;
  return a.i < b.i;
}

/* out(idx) = expr(idx): HostExecutor::Exec, executors/host.h:147-174 */
int orc_elementwise(const mxb_expr_t *e, const mxb_out_t *out) {
  int64_t N = 1;
  for (int d = 0; d < e->rank; ++d) N *= e->size[d];
  int64_t idx[MXB_MAX_RANK] = {0};
  for (int64_t i = 0; i < N; ++i) {
    unflatten(i, e->rank, e->size, idx);
    int64_t off = 0;
    for (int d = 0; d < e->rank; ++d) off += idx[d] * out->stride[d];
    store_val(out->data, out->dtype, off, eval_expr(e, idx));
  }
  return 0;
}

/* Reduce the trailing n_reduce dims.  half_acc: round the running sum through this 16-bit dtype after every
 * add (MXB_BF16 / MXB_F16), the faithful host behaviour for 16-bit value types; -1 = accumulate in the
 * arithmetic type. */
int orc_reduce(int op, const mxb_expr_t *e, int n_reduce, const mxb_out_t *out, const mxb_out_t *idx_out, int ddof, int half_acc) {
  const int nb = e->rank - n_reduce;
  int64_t B = 1, R = 1;
  for (int d = 0; d < nb; ++d) B *= e->size[d];
  for (int d = nb; d < e->rank; ++d) R *= e->size[d];
  if (R <= 0) return 1;
  int64_t idx[MXB_MAX_RANK] = {0};
  for (int64_t b = 0; b < B; ++b) {
    unflatten(b, nb, e->size, idx);
    int64_t ooff = 0, ioff = 0;
    for (int d = 0; d < nb; ++d) { ooff += idx[d] * out->stride[d]; if (idx_out) ioff += idx[d] * idx_out->stride[d]; }
    val_t acc; memset(&acc, 0, sizeof acc);
    val_t mean; memset(&mean, 0, sizeof mean);
    int64_t best = 0;
    int flag = (op == MXB_RED_ALL) ? 1 : 0;
    const int passes = (op == MXB_RED_VAR || op == MXB_RED_STDD) ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
      for (int64_t r = 0; r < R; ++r) {
        unflatten(r, n_reduce, e->size + nb, idx + nb);
        val_t x = eval_expr(e, idx);
        if (r == 0 && pass == 0) {
          acc = x;  /* typed zero / one below */
          switch (op) {
            case MXB_RED_SUM: case MXB_RED_MEAN: case MXB_RED_VAR: case MXB_RED_STDD: {
              val_t z; memset(&z, 0, sizeof z); z.t = x.t; acc = arith(MXB_OP_ADD, z, x); break;  /* accumulate(first,last,T(0)) */
            }
            case MXB_RED_PROD: { val_t o = conv(mk_i(MXB_I32, 1), x.t); acc = arith(MXB_OP_MUL, o, x); break; }
            default: break;
          }
          best = 0;
          if (op == MXB_RED_ANY) flag = nonzero(x);
          if (op == MXB_RED_ALL) flag = nonzero(x);
          if (half_acc == MXB_BF16) acc.f = bf16_to_f32(f32_to_bf16(acc.f));
          if (half_acc == MXB_F16) acc.f = f16_to_f32(f32_to_f16(acc.f));
          continue;
        }
        if (pass == 1) {
          /* pow(abs(x - mean), 2): reduce.h:1431 */
          val_t dlt = arith(MXB_OP_SUB, x, mean);
          val_t sq;
          if (dlt.t == MXB_C64) { float a = hypotf(dlt.f, dlt.fi); sq = mk_f(powf(a, 2.f)); }
          else if (dlt.t == MXB_F32) sq = mk_f(powf(fabsf(dlt.f), 2.f));
          else sq = mk_d(pow(fabs(conv(dlt, MXB_F64).d), 2.0));
          if (r == 0) { val_t z; memset(&z, 0, sizeof z); z.t = sq.t; acc = arith(MXB_OP_ADD, z, sq); }
          else acc = arith(MXB_OP_ADD, acc, sq);
          continue;
        }
        switch (op) {
          case MXB_RED_SUM: case MXB_RED_MEAN: case MXB_RED_VAR: case MXB_RED_STDD:
            acc = arith(MXB_OP_ADD, acc, x);
            if (half_acc == MXB_BF16) acc.f = bf16_to_f32(f32_to_bf16(acc.f));
            if (half_acc == MXB_F16) acc.f = f16_to_f32(f32_to_f16(acc.f));
            break;
          case MXB_RED_PROD: acc = arith(MXB_OP_MUL, acc, x); break;
          case MXB_RED_MAX: case MXB_RED_ARGMAX: if (val_less(acc, x)) { acc = x; best = r; } break;  /* std::max_element */
          case MXB_RED_MIN: case MXB_RED_ARGMIN: if (val_less(x, acc)) { acc = x; best = r; } break;  /* std::min_element */
          case MXB_RED_ANY: flag = flag || nonzero(x); break;
          case MXB_RED_ALL: flag = flag && nonzero(x); break;
        }
      }
      if (passes == 2 && pass == 0) {
        /* mean_impl: sum / N in the value type (complex / real N) */
        mean = acc.t == MXB_C64 ? mk_c(acc.f / (float)R, acc.fi / (float)R) : acc.t == MXB_F32 ? mk_f(acc.f / (float)R) : mk_d(conv(acc, MXB_F64).d / (double)R);
      }
    }
    val_t res = acc;
    switch (op) {
      case MXB_RED_MEAN:
        res = acc.t == MXB_C64 ? mk_c(acc.f / (float)R, acc.fi / (float)R) : acc.t == MXB_F32 ? mk_f(acc.f / (float)R) : mk_d(conv(acc, MXB_F64).d / (double)R);
        break;
      case MXB_RED_VAR: case MXB_RED_STDD:
        if (acc.t == MXB_F32) { res = mk_f(acc.f / (float)(R - ddof)); if (op == MXB_RED_STDD) res.f = sqrtf(res.f); }
        else { res = mk_d(acc.d / (double)(R - ddof)); if (op == MXB_RED_STDD) res.d = sqrt(res.d); }
        break;
      case MXB_RED_ANY: case MXB_RED_ALL: res = mk_b(flag); break;
      default: break;
    }
    store_val(out->data, out->dtype, ooff, res);
    if (idx_out && (op == MXB_RED_ARGMAX || op == MXB_RED_ARGMIN)) ((int64_t *)idx_out->data)[ioff] = b * R + best;
  }
  return 0;
}

/* softmax over the trailing n_reduce dims — softmax_impl, transforms/reduce.h:362-445: tmp_max = max(in);
 * tmp_sum = sum(exp(in - tmp_max)); dest = exp(in - tmp_max) / tmp_sum, all in the value type (the reference has a
 * CUDA implementation only; the reductions are restated sequentially here, so sums agree with it to rounding).
 * `out` has the rank and sizes of `e`. */
int orc_softmax(const mxb_expr_t *e, int n_reduce, const mxb_out_t *out) {
  const int nb = e->rank - n_reduce;
  int64_t B = 1, R = 1;
  for (int d = 0; d < nb; ++d) B *= e->size[d];
  for (int d = nb; d < e->rank; ++d) R *= e->size[d];
  if (n_reduce < 1) return 1;
  int64_t idx[MXB_MAX_RANK] = {0};
  for (int64_t b = 0; b < B; ++b) {
    unflatten(b, nb, e->size, idx);
    val_t mx; memset(&mx, 0, sizeof mx);
    for (int64_t r = 0; r < R; ++r) {
      unflatten(r, n_reduce, e->size + nb, idx + nb);
      val_t x = eval_expr(e, idx);
      if (x.t != MXB_F32 && x.t != MXB_F64) return 2;
      if (r == 0 || val_less(mx, x)) mx = x;
    }
    val_t sum; memset(&sum, 0, sizeof sum); sum.t = mx.t;
    for (int64_t r = 0; r < R; ++r) {
      unflatten(r, n_reduce, e->size + nb, idx + nb);
      val_t ex = unary_math(MXB_OP_EXP, arith(MXB_OP_SUB, eval_expr(e, idx), mx));
      sum = arith(MXB_OP_ADD, sum, ex);
    }
    for (int64_t r = 0; r < R; ++r) {
      unflatten(r, n_reduce, e->size + nb, idx + nb);
      val_t ex = unary_math(MXB_OP_EXP, arith(MXB_OP_SUB, eval_expr(e, idx), mx));
      int64_t off = 0;
      for (int d = 0; d < e->rank; ++d) off += idx[d] * out->stride[d];
      store_val(out->data, out->dtype, off, arith(MXB_OP_DIV, ex, sum));
    }
  }
  return 0;
}

/* cumsum along the last dim — cumsum_impl(HostExecutor), transforms/cub.h:2397-2440 -> host_inclusive_scan,
 * transforms/host_algorithms.h:231-249: std::partial_sum per row, i.e. a sequential running sum in the value type
 * (16-bit float leaves arrive widened to fp32 here, like everywhere on this path).  Pinned by the reference's own
 * known answers: CUBTests.cu:203-226 (permuted int matrix), :536-575 (running float sums), stack_test.cu:72-80. */
int orc_cumsum(const mxb_expr_t *e, const mxb_out_t *out) {
  if (e->rank < 1) return 1;
  const int nb = e->rank - 1;
  int64_t B = 1;
  for (int d = 0; d < nb; ++d) B *= e->size[d];
  const int64_t L = e->size[nb];
  int64_t idx[MXB_MAX_RANK] = {0};
  for (int64_t b = 0; b < B; ++b) {
    unflatten(b, nb, e->size, idx);
    val_t run; memset(&run, 0, sizeof run);
    for (int64_t j = 0; j < L; ++j) {
      idx[nb] = j;
      val_t x = eval_expr(e, idx);
      run = j == 0 ? x : arith(MXB_OP_ADD, run, x);
      int64_t off = 0;
      for (int d = 0; d < e->rank; ++d) off += idx[d] * out->stride[d];
      store_val(out->data, out->dtype, off, run);
    }
  }
  return 0;
}

/* find / find_idx — find_impl / find_idx_impl(HostExecutor), transforms/cub.h:2656-2675,2752-2770: walk the operator in
 * flat row-major order (cbegin .. cend), keep the elements the selection functor accepts (LT / GT / EQ / NEQ / LTE / GTE
 * with the threshold in the VALUE type, :2521-2588) — or their flat index as static_cast<int> — and count them in an int.
 * select_op follows mxb_select_op_t.  Elements beyond the output's capacity are counted but not stored. */
int orc_find(const mxb_expr_t *e, int select_op, double threshold, const mxb_out_t *out, int32_t *count_out, int want_indices) {
  static const int kOp[6] = {MXB_OP_LT, MXB_OP_GT, MXB_OP_EQ, MXB_OP_NE, MXB_OP_LE, MXB_OP_GE};
  if (select_op < 0 || select_op > 5 || out->rank != 1) return 1;
  int64_t N = 1;
  for (int d = 0; d < e->rank; ++d) N *= e->size[d];
  int64_t idx[MXB_MAX_RANK] = {0};
  int32_t cnt = 0;
  for (int64_t f = 0; f < N; ++f) {
    unflatten(f, e->rank, e->size, idx);
    val_t x = eval_expr(e, idx);
    val_t c;
    if (x.t == MXB_F32) c = mk_f((float)threshold);
    else if (x.t == MXB_F64) c = mk_d(threshold);
    else c = mk_i(x.t, (long long)threshold);
    if (compare(kOp[select_op], x, c)) {
      if (cnt < out->size[0]) store_val(out->data, out->dtype, cnt, want_indices ? mk_i(MXB_I64, f) : x);
      ++cnt;
    }
  }
  *count_out = cnt;
  return 0;
}

/* ---- sort / unique / hist -----------------------------------------------------------
 * sort: the HostExecutor overload of sort_impl (transforms/cub.h:2192-2240) — std::sort of every row of the last dim,
 *       ascending or descending (std::greater).  Restated with qsort on the value type (sorting is a function of the
 *       multiset of values; only the relative order of -0.0 / +0.0 and of NaNs is unspecified, and the tests avoid both).
 * unique: the HostExecutor overload of unique_impl (:2844-2880) — std::sort, then std::unique, count in an int.
 * hist: the reference has NO host implementation; the arithmetic is third-party: cub::DeviceHistogram::HistogramEven of
 *       NVIDIA/cccl 3.3.0 (pinned in cmake/versions.json:3-8, not vendored under /root/reference; the nearest copy in this
 *       image is the 3.3.2 tree inside the flashinfer wheel).  Restated from its published algorithm
 *       (cub/device/dispatch/kernels/kernel_histogram.cuh, ScaleTransform): a sample counts iff lower <= x < upper; bin =
 *       (int)((x - lower) * scale), scale = T(bins) / T(upper - lower), for floating T; ((x - lower) * bins) / (upper - lower)
 *       in unsigned 64-bit arithmetic for integers.  Pinned by the reference's own known answers (test/00_tensor/CUBTests.cu:
 *       153-197), not by a run of the reference: parity for hist is "pinned to third party" (DESIGN.md section 5). */
static int cmp_f32(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }
static int cmp_f64(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }
static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }
static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }

static int sort_buffer(void *buf, int dtype, int64_t n, int descending) {
  size_t esz;
  int (*cmp)(const void *, const void *);
  switch (dtype) {
    case MXB_F32: esz = 4; cmp = cmp_f32; break;
    case MXB_F64: esz = 8; cmp = cmp_f64; break;
    case MXB_I32: esz = 4; cmp = cmp_i32; break;
    case MXB_I64: esz = 8; cmp = cmp_i64; break;
    default: return 1;
  }
  qsort(buf, (size_t)n, esz, cmp);
  if (descending) {
    char *p = (char *)buf, tmp[8];
    for (int64_t i = 0, j = n - 1; i < j; ++i, --j) {
      memcpy(tmp, p + i * esz, esz); memcpy(p + i * esz, p + j * esz, esz); memcpy(p + j * esz, tmp, esz);
    }
  }
  return 0;
}

/* out (contiguous, the expression's shape and value type) = every row of the last dim of `e`, sorted */
int orc_sort(const mxb_expr_t *e, const mxb_out_t *out, int descending) {
  if (e->rank < 1 || out->rank != e->rank) return 1;
  int64_t N = 1;
  for (int d = 0; d < e->rank; ++d) N *= e->size[d];
  if (N == 0) return 0;
  const int64_t L = e->size[e->rank - 1];
  int64_t idx[MXB_MAX_RANK] = {0};
  for (int64_t f = 0; f < N; ++f) {
    unflatten(f, e->rank, e->size, idx);
    store_val(out->data, out->dtype, f, conv(eval_expr(e, idx), out->dtype));
  }
  const size_t esz = (out->dtype == MXB_F64 || out->dtype == MXB_I64) ? 8 : 4;
  for (int64_t b = 0; b < N / L; ++b)
    if (sort_buffer((char *)out->data + (size_t)(b * L) * esz, out->dtype, L, descending)) return 1;
  return 0;
}

int orc_unique(const mxb_expr_t *e, const mxb_out_t *out, int32_t *count_out) {
  if (e->rank != 1 || out->rank != 1) return 1;
  const int64_t N = e->size[0];
  const size_t esz = (out->dtype == MXB_F64 || out->dtype == MXB_I64) ? 8 : 4;
  char *tmp = (char *)malloc((size_t)(N > 0 ? N : 1) * esz);
  if (!tmp) return 1;
  int64_t idx[MXB_MAX_RANK] = {0};
  for (int64_t f = 0; f < N; ++f) { idx[0] = f; store_val(tmp, out->dtype, f, conv(eval_expr(e, idx), out->dtype)); }
  if (sort_buffer(tmp, out->dtype, N, 0)) { free(tmp); return 1; }
  int32_t cnt = 0;
  for (int64_t f = 0; f < N; ++f) {
    if (f == 0 || memcmp(tmp + f * esz, tmp + (f - 1) * esz, esz) != 0) {
      /* std::unique compares with ==: -0.0 == +0.0 and NaN != NaN; the tests use neither */
      if (cnt < out->size[0]) memcpy((char *)out->data + (size_t)cnt * esz, tmp + f * esz, esz);
      ++cnt;
    }
  }
  free(tmp);
  *count_out = cnt;
  return 0;
}

/* out(b..., k) int32 counts; bins = out->size[last]; out walks its own strides */
int orc_hist(const mxb_expr_t *e, double lower, double upper, const mxb_out_t *out) {
  if (e->rank < 1 || out->rank != e->rank || out->dtype != MXB_I32) return 1;
  const int64_t bins = out->size[out->rank - 1], L = e->size[e->rank - 1];
  int64_t B = 1;
  for (int d = 0; d + 1 < e->rank; ++d) B *= e->size[d];
  int64_t idx[MXB_MAX_RANK] = {0}, bidx[MXB_MAX_RANK] = {0};
  for (int64_t b = 0; b < B; ++b) {
    unflatten(b, e->rank - 1, e->size, bidx);
    int64_t obase = 0;
    for (int d = 0; d + 1 < e->rank; ++d) { idx[d] = bidx[d]; obase += bidx[d] * out->stride[d]; }
    int32_t *ob = (int32_t *)out->data + obase;
    for (int64_t k = 0; k < bins; ++k) ob[k * out->stride[out->rank - 1]] = 0;
    for (int64_t j = 0; j < L; ++j) {
      idx[e->rank - 1] = j;
      val_t x = eval_expr(e, idx);
      int bin = -1;
      if (x.t == MXB_F32) {
        const float lo = (float)lower, hi = (float)upper, scale = (float)bins / (hi - lo);
        if (x.f >= lo && x.f < hi) bin = (int)((x.f - lo) * scale);
      } else if (x.t == MXB_F64) {
        const double scale = (double)bins / (upper - lower);
        if (x.d >= lower && x.d < upper) bin = (int)((x.d - lower) * scale);
      } else if (is_int(x.t)) {
        const long long lo = (long long)lower, hi = (long long)upper;
        if (x.i >= lo && x.i < hi) bin = (int)(((unsigned long long)(x.i - lo) * (unsigned long long)bins) / (unsigned long long)(hi - lo));
      } else return 1;
      if (bin >= 0 && bin < bins) ob[bin * out->stride[out->rank - 1]] += 1;
    }
  }
  return 0;
}
