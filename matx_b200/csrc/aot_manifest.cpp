// aot_manifest.cpp — the named expression programs and the list of kernel instances that are compiled
// ahead of time into libmatx_b200.so.  Anything not listed here is built on first use by NVRTC from the
// same generator and the same skeleton header, so the manifest is a warm-start list, not a feature list.
#include <cstring>
#include <map>
#include <tuple>

#include "mxb_internal.h"

namespace mxbh {

ExprBuilder::ExprBuilder() { memset(&e, 0, sizeof e); }
int ExprBuilder::leaf(int dtype) {
  const int k = e.n_leaves++;
  e.leaves[k].dtype = dtype;
  mxb_node_t &n = e.nodes[e.n_nodes];
  n.opcode = MXB_OP_LEAF; n.src[0] = k; n.src[1] = -1; n.aux = 0;
  return e.n_nodes++;
}
int ExprBuilder::cst(double v, int dtype) {
  const int k = e.n_consts++;
  e.consts[k].re = v; e.consts[k].im = 0; e.consts[k].dtype = dtype;
  mxb_node_t &n = e.nodes[e.n_nodes];
  n.opcode = MXB_OP_CONST; n.src[0] = k; n.src[1] = -1; n.aux = 0;
  return e.n_nodes++;
}
int ExprBuilder::un(int opcode, int a, int aux) {
  mxb_node_t &n = e.nodes[e.n_nodes];
  n.opcode = opcode; n.src[0] = a; n.src[1] = -1; n.aux = aux;
  return e.n_nodes++;
}
int ExprBuilder::bin(int opcode, int a, int b) {
  mxb_node_t &n = e.nodes[e.n_nodes];
  n.opcode = opcode; n.src[0] = a; n.src[1] = b; n.aux = 0;
  return e.n_nodes++;
}
mxb_expr_t ExprBuilder::finish(int root) {
  e.root = root;
  e.rank = 1;
  e.size[0] = 1;
  mxb_expr_t c;
  std::string err;
  // leaves of a structural program carry distinct fake addresses so canonicalize keeps them apart
  for (int k = 0; k < e.n_leaves; ++k) e.leaves[k].data = (const void *)(uintptr_t)(0x1000 * (k + 1));
  canonicalize(&e, &c, &err);
  return c;
}

int canonicalize(const mxb_expr_t *in, mxb_expr_t *out, std::string *err) {
  auto fail = [&](const std::string &m) { if (err) *err = m; return (int)MXB_ERR_INVALID; };
  if (in->n_nodes <= 0 || in->n_nodes > MXB_MAX_NODES || in->root < 0 || in->root >= in->n_nodes) return fail("malformed program");
  if (in->n_leaves < 0 || in->n_leaves > MXB_MAX_LEAVES || in->n_consts < 0 || in->n_consts > MXB_MAX_CONSTS) return fail("malformed program");
  if (in->rank < 0 || in->rank > MXB_MAX_RANK) return fail("expression rank out of range");
  memset(out, 0, sizeof *out);
  out->rank = in->rank;
  memcpy(out->size, in->size, sizeof in->size);

  // leaves: merge identical views
  int leaf_map[MXB_MAX_LEAVES];
  int leaf_new[MXB_MAX_LEAVES];  // canonical index assigned on first use
  for (int k = 0; k < in->n_leaves; ++k) {
    leaf_map[k] = k;
    leaf_new[k] = -1;
    for (int j = 0; j < k; ++j) {
      if (in->leaves[j].data == in->leaves[k].data && in->leaves[j].dtype == in->leaves[k].dtype &&
          memcmp(in->leaves[j].stride, in->leaves[k].stride, sizeof(int64_t) * (size_t)(in->rank > 0 ? in->rank : 0)) == 0) {
        leaf_map[k] = leaf_map[j];
        break;
      }
    }
  }
  int const_map[MXB_MAX_CONSTS];
  int const_new[MXB_MAX_CONSTS];
  for (int k = 0; k < in->n_consts; ++k) {
    const_map[k] = k;
    const_new[k] = -1;
    for (int j = 0; j < k; ++j) {
      if (in->consts[j].dtype == in->consts[k].dtype && memcmp(&in->consts[j].re, &in->consts[k].re, sizeof(double)) == 0 &&
          memcmp(&in->consts[j].im, &in->consts[k].im, sizeof(double)) == 0) {
        const_map[k] = const_map[j];
        break;
      }
    }
  }
  std::map<std::tuple<int, int, int, int>, int> cons;
  int memo[MXB_MAX_NODES];
  for (int i = 0; i < in->n_nodes; ++i) memo[i] = -1;

  // iterative post-order (programs are small, but recursion depth is still user-controlled)
  struct Frame { int node; int stage; };
  std::vector<Frame> st;
  st.push_back({in->root, 0});
  while (!st.empty()) {
    Frame &f = st.back();
    const mxb_node_t &n = in->nodes[f.node];
    if (memo[f.node] >= 0) { st.pop_back(); continue; }
    const bool leafish = n.opcode == MXB_OP_LEAF || n.opcode == MXB_OP_CONST;
    const bool binary = n.opcode >= MXB_OP_ADD && n.opcode <= MXB_OP_ATAN2;
    if (!leafish) {
      if (n.src[0] < 0 || n.src[0] >= in->n_nodes || n.src[0] == f.node) return fail("operand id out of range");
      if (binary && (n.src[1] < 0 || n.src[1] >= in->n_nodes || n.src[1] == f.node)) return fail("operand id out of range");
      // the walk of a DAG never holds more frames than there are nodes: a deeper stack means a cycle (a -> b -> a)
      if (st.size() > (size_t)in->n_nodes) return fail("cyclic program");
      if (f.stage == 0) { f.stage = 1; if (memo[n.src[0]] < 0) { const int c = n.src[0]; st.push_back({c, 0}); continue; } }
      if (f.stage == 1) { f.stage = 2; if (binary && memo[n.src[1]] < 0) { const int c = n.src[1]; st.push_back({c, 0}); continue; } }
    }
    int s0, s1 = -1;
    if (n.opcode == MXB_OP_LEAF) {
      if (n.src[0] < 0 || n.src[0] >= in->n_leaves) return fail("leaf index out of range");
      const int k = leaf_map[n.src[0]];
      if (leaf_new[k] < 0) { leaf_new[k] = out->n_leaves; out->leaves[out->n_leaves++] = in->leaves[k]; }
      s0 = leaf_new[k];
    } else if (n.opcode == MXB_OP_CONST) {
      if (n.src[0] < 0 || n.src[0] >= in->n_consts) return fail("constant index out of range");
      const int k = const_map[n.src[0]];
      if (const_new[k] < 0) { const_new[k] = out->n_consts; out->consts[out->n_consts++] = in->consts[k]; }
      s0 = const_new[k];
    } else {
      s0 = memo[n.src[0]];
      if (binary) s1 = memo[n.src[1]];
    }
    const auto key = std::make_tuple(n.opcode, s0, s1, n.opcode == MXB_OP_CAST ? n.aux : 0);
    auto it = cons.find(key);
    int id;
    if (it != cons.end()) id = it->second;
    else {
      if (out->n_nodes >= MXB_MAX_NODES) return fail("program too large");
      id = out->n_nodes++;
      out->nodes[id].opcode = n.opcode;
      out->nodes[id].src[0] = s0;
      out->nodes[id].src[1] = s1;
      out->nodes[id].aux = n.opcode == MXB_OP_CAST ? n.aux : 0;
      cons[key] = id;
    }
    memo[f.node] = id;
    st.pop_back();
  }
  out->root = memo[in->root];
  return MXB_OK;
}

mxb_expr_t prog_identity(int dtype) { ExprBuilder b; return b.finish(b.leaf(dtype)); }
mxb_expr_t prog_fma3(int dtype) {
  ExprBuilder b;
  const int a = b.leaf(dtype), x = b.leaf(dtype), c = b.leaf(dtype);
  return b.finish(b.bin(MXB_OP_ADD, b.bin(MXB_OP_MUL, a, x), c));
}
mxb_expr_t prog_abs2(int dtype) { ExprBuilder b; return b.finish(b.un(MXB_OP_ABS2, b.leaf(dtype))); }
mxb_expr_t prog_vector_add(int dtype) {
  ExprBuilder b;
  const int a = b.leaf(dtype), x = b.leaf(dtype);
  return b.finish(b.bin(MXB_OP_ADD, a, x));
}
mxb_expr_t prog_black_scholes() {
  // output = S*normcdf(d1) - K*exp(-1.f*r*T)*normcdf(d2)
  //   VsqrtT = V*sqrt(T);  d1 = (log(S/K) + (r + 0.5f*V*V)*T) / VsqrtT;  d2 = d1 - VsqrtT
  ExprBuilder b;
  const int K = b.leaf(MXB_F32), S = b.leaf(MXB_F32), V = b.leaf(MXB_F32), r = b.leaf(MXB_F32), T = b.leaf(MXB_F32);
  const int half = b.cst(0.5, MXB_F32), neg1 = b.cst(-1.0, MXB_F32);
  const int VsqrtT = b.bin(MXB_OP_MUL, V, b.un(MXB_OP_SQRT, T));
  const int lsk = b.un(MXB_OP_LOG, b.bin(MXB_OP_DIV, S, K));
  const int hvv = b.bin(MXB_OP_MUL, b.bin(MXB_OP_MUL, half, V), V);
  const int num = b.bin(MXB_OP_ADD, lsk, b.bin(MXB_OP_MUL, b.bin(MXB_OP_ADD, r, hvv), T));
  const int d1 = b.bin(MXB_OP_DIV, num, VsqrtT);
  const int d2 = b.bin(MXB_OP_SUB, d1, VsqrtT);
  const int c1 = b.un(MXB_OP_NORMCDF, d1), c2 = b.un(MXB_OP_NORMCDF, d2);
  const int expRT = b.un(MXB_OP_EXP, b.bin(MXB_OP_MUL, b.bin(MXB_OP_MUL, neg1, r), T));
  const int out = b.bin(MXB_OP_SUB, b.bin(MXB_OP_MUL, S, c1), b.bin(MXB_OP_MUL, b.bin(MXB_OP_MUL, K, expRT), c2));
  return b.finish(out);
}

void aot_manifest(std::vector<ManifestItem> *items) {
  auto add = [&](const mxb_expr_t &e, int family, int op, int out_dtype, int team, bool scalar_too) {
    ExprInfo info;
    std::string err;
    if (analyze_expr(&e, &info, &err) != MXB_OK) return;
    for (int pass = 0; pass < (scalar_too ? 2 : 1); ++pass) {
      ManifestItem it;
      it.expr = e;
      it.spec.family = family;
      it.spec.op = op;
      it.spec.out_dtype = out_dtype;
      it.spec.V = pass == 0 ? policy_vmax(info) : 1;
      if (family == FAM_EW_TR || family == FAM_RED_OUTER_TMA) it.spec.V = 16 / info.max_leaf_bytes;  // one 16-byte chunk
      it.spec.U = policy_unroll(info, it.spec.V, family);
      it.spec.team = team;
      items->push_back(it);
    }
  };
  // explicit vector width / loads in flight (the dispatcher's choices that differ from the shared policy)
  auto add_vu = [&](const mxb_expr_t &e, int family, int op, int out_dtype, int team, int V, int U) {
    ExprInfo info;
    std::string err;
    if (analyze_expr(&e, &info, &err) != MXB_OK) return;
    ManifestItem it;
    it.expr = e;
    it.spec.family = family;
    it.spec.op = op;
    it.spec.out_dtype = out_dtype;
    it.spec.V = V;
    it.spec.U = U;
    it.spec.team = team;
    items->push_back(it);
  };
  const int kAllOps[] = {MXB_RED_SUM, MXB_RED_PROD, MXB_RED_MAX, MXB_RED_MIN, MXB_RED_ARGMAX, MXB_RED_ARGMIN, MXB_RED_ANY, MXB_RED_ALL};
  // CTA-per-item streaming of a single contiguous reduce run: 32-byte loads for the sums, four loads in flight for all
  add_vu(prog_identity(MXB_F32), FAM_RED_INNER, MXB_RED_SUM, MXB_F32, 0, 8, 4);
  add_vu(prog_identity(MXB_F32), FAM_RED_INNER, MXB_RED_PROD, MXB_F32, 0, 8, 4);
  add_vu(prog_identity(MXB_F32), FAM_RED_INNER, MXB_RED_MAX, MXB_F32, 0, 8, 4);
  add_vu(prog_identity(MXB_F32), FAM_RED_INNER, MXB_RED_MIN, MXB_F32, 0, 8, 4);
  add_vu(prog_identity(MXB_I32), FAM_RED_INNER, MXB_RED_MAX, MXB_I32, 0, 8, 4);
  add_vu(prog_identity(MXB_I32), FAM_RED_INNER, MXB_RED_MIN, MXB_I32, 0, 8, 4);
  add_vu(prog_identity(MXB_I32), FAM_RED_INNER, MXB_RED_SUM, MXB_I32, 0, 8, 4);
  for (int op : kAllOps) {
    if (op != MXB_RED_PROD) add_vu(prog_identity(MXB_F64), FAM_RED_INNER, op, MXB_F64, 0, 4, 4);
    if (op == MXB_RED_SUM || op == MXB_RED_ANY || op == MXB_RED_ALL) add_vu(prog_identity(MXB_C64), FAM_RED_INNER, op, MXB_C64, 0, 4, 4);
  }
  add_vu(prog_identity(MXB_C64), FAM_RED_INNER, MXB_RED_VAR, MXB_F32, 0, 4, 4);
  for (int op : {MXB_RED_SUM, MXB_RED_MAX, MXB_RED_ARGMAX}) add_vu(prog_abs2(MXB_C64), FAM_RED_INNER, op, MXB_F32, 0, 4, 4);
  // identity programs: full / per-row reductions of plain tensors
  for (int d : {MXB_F32, MXB_F64, MXB_BF16, MXB_C64, MXB_I32}) {
    const mxb_expr_t e = prog_identity(d);
    for (int op : kAllOps) {
      if (d == MXB_C64 && (op == MXB_RED_MAX || op == MXB_RED_MIN || op == MXB_RED_ARGMAX || op == MXB_RED_ARGMIN)) continue;
      if (d != MXB_F32 && op == MXB_RED_PROD) continue;
      for (int team : {0, 1}) add(e, FAM_RED_INNER, op, d, team, d == MXB_F32);
      if (d == MXB_F32 || d == MXB_BF16) add(e, FAM_RED_OUTER, op, d, 0, d == MXB_F32);
      // TMA-staged tiles for a strided / permuted reduce dim (config 5 and column reductions of plain tensors)
      if ((d == MXB_F32 || d == MXB_BF16) && op != MXB_RED_PROD) add(e, FAM_RED_OUTER_TMA, op, d, 0, false);
    }
  }
  // fused argminmax of plain fp32 tensors
  for (int team : {0, 1}) {
    add(prog_identity(MXB_F32), FAM_RED_INNER, KOP_ARGMINMAX, MXB_F32, team, false);
    if (team == 0) items->back().spec.minb = 4;   // the dispatcher's register cap for the CTA-per-item shape
  }
  // one-pass variance (Welford + Chan) for rows that cannot stay on chip and for strided rows
  for (int d : {MXB_F32, MXB_C64})
    for (int team : {0, 1}) add(prog_identity(d), FAM_RED_INNER, MXB_RED_VAR, MXB_F32, team, d == MXB_F32);
  add(prog_identity(MXB_F32), FAM_RED_OUTER, MXB_RED_VAR, MXB_F32, 0, false);
  add(prog_identity(MXB_F32), FAM_RED_OUTER_TMA, MXB_RED_VAR, MXB_F32, 0, false);
  for (int d : {MXB_F32, MXB_C64}) add(prog_identity(d), FAM_VAR_SMEM, MXB_RED_VAR, MXB_F32, 0, false);
  add(prog_identity(MXB_F64), FAM_VAR_SMEM, MXB_RED_VAR, MXB_F64, 0, false);
  for (int ipt : {1, 2, 4, 8}) {
    for (int d : {MXB_F32, MXB_C64, MXB_BF16}) add(prog_identity(d), FAM_VAR_TMA, MXB_RED_VAR, MXB_F32, ipt, false);
    add(prog_identity(MXB_F64), FAM_VAR_TMA, MXB_RED_VAR, MXB_F64, ipt, false);
  }
  for (int ipt : {1, 2, 4, 8}) {
    for (int d : {MXB_F32, MXB_C64}) add(prog_identity(d), FAM_VAR_GROUP, MXB_RED_VAR, MXB_F32, ipt, false);
    for (int d : {MXB_F32, MXB_C64}) add(prog_identity(d), FAM_VAR_REG, MXB_RED_VAR, MXB_F32, ipt, false);
    add(prog_identity(MXB_F64), FAM_VAR_REG, MXB_RED_VAR, MXB_F64, ipt, false);
  }
  // softmax: rows in registers, plus the statistics + apply pair for everything else
  for (int ipt : {1, 2, 4, 8}) {
    for (int fam : {FAM_SM_GROUP, FAM_SM_REG}) {
      add(prog_identity(MXB_F32), fam, -1, MXB_F32, ipt, true);
      add(prog_identity(MXB_BF16), fam, -1, MXB_BF16, ipt, false);
      add(prog_identity(MXB_F64), fam, -1, MXB_F64, ipt, false);
    }
  }
  for (int team : {0, 1}) add(prog_identity(MXB_F32), FAM_RED_INNER, KOP_LSE, MXB_F32, team, true);
  add(prog_identity(MXB_F32), FAM_RED_OUTER, KOP_LSE, MXB_F32, 0, false);
  {
    ExprBuilder b;
    const int x = b.leaf(MXB_F32), m = b.leaf(MXB_F32), sm = b.leaf(MXB_F32);
    add(b.finish(b.bin(MXB_OP_MUL, b.un(MXB_OP_EXP, b.bin(MXB_OP_SUB, x, m)), sm)), FAM_EW, -1, MXB_F32, 0, false);
  }
  // config 1: sum(a*b+c, {1})
  for (int op : {MXB_RED_SUM, MXB_RED_MAX, MXB_RED_ARGMAX})
    for (int team : {0, 1}) add(prog_fma3(MXB_F32), FAM_RED_INNER, op, MXB_F32, team, false);
  // config 3: argmax(abs2(x), {1}) on complex<float>
  for (int op : {MXB_RED_SUM, MXB_RED_MAX, MXB_RED_ARGMAX})
    for (int team : {0, 1}) add(prog_abs2(MXB_C64), FAM_RED_INNER, op, MXB_F32, team, false);
  // config 4 and the reference's own elementwise benches
  add(prog_black_scholes(), FAM_EW, -1, MXB_F32, 0, false);
  add(prog_black_scholes(), FAM_EW, -1, MXB_F32, 4, false);   // 4 CTAs per SM (64 registers): the dispatcher's choice for heavy programs
  add(prog_fma3(MXB_F32), FAM_EW, -1, MXB_F32, 0, false);
  for (int d : {MXB_F32, MXB_F64, MXB_C64}) add(prog_vector_add(d), FAM_EW, -1, d, 0, false);
  for (int d : {MXB_F32, MXB_BF16}) add(prog_identity(d), FAM_EW, -1, d, 0, d == MXB_F32);
  // cumsum of plain tensors: long-row (U = 4) and short-row (U = 1) tiles
  for (int d : {MXB_F32, MXB_F64, MXB_C64, MXB_I32, MXB_BF16})
    for (int u : {(d == MXB_F64 || d == MXB_C64) ? 2 : 4, 1}) {
      const mxb_expr_t e = prog_identity(d);
      ExprInfo info;
      std::string err;
      if (analyze_expr(&e, &info, &err) != MXB_OK) continue;
      ManifestItem it;
      it.expr = e;
      it.spec.family = FAM_SCAN;
      it.spec.op = -1;
      it.spec.out_dtype = d;
      it.spec.V = policy_vmax(info);
      it.spec.U = u;
      it.spec.team = 0;
      items->push_back(it);
      // warp-per-row twins for short rows: U = 1, and U = 2 for the 2- and 4-byte types
      it.spec.team = 1;
      const bool one_only = d == MXB_F64 || d == MXB_C64 || d == MXB_BF16;
      it.spec.U = (u == 1 || one_only) ? 1 : 2;
      if (!(u != 1 && one_only)) items->push_back(it);
    }
  // find / find_idx of plain tensors: count pass, value scatter, index scatter (int and index_t outputs)
  for (int d : {MXB_F32, MXB_F64, MXB_I32}) {
    add(prog_identity(d), FAM_SELECT, -1, d, 0, d == MXB_F32);
    add(prog_identity(d), FAM_SELECT, -1, d, 1, d == MXB_F32);
    add(prog_identity(d), FAM_SELECT, -1, MXB_I32, 2, d == MXB_F32);
    add(prog_identity(d), FAM_SELECT, -1, MXB_I64, 2, false);
  }
  // single-pass look-back kernels: values, int / index_t flat indices
  for (int d : {MXB_F32, MXB_F64, MXB_I32}) {
    add(prog_identity(d), FAM_SELECT, -1, d, 3, d == MXB_F32);
    add(prog_identity(d), FAM_SELECT, -1, MXB_I32, 4, d == MXB_F32);
    add(prog_identity(d), FAM_SELECT, -1, MXB_I64, 4, false);
  }
  // histograms of plain tensors
  for (int d : {MXB_F32, MXB_F64, MXB_I32}) add(prog_identity(d), FAM_HIST, -1, MXB_I32, 0, d == MXB_F32);
  // permuted copies (bench/00_operators/operators.cu:40-59) and the row + column mix
  for (int d : {MXB_F32, MXB_F64, MXB_C64, MXB_BF16, MXB_I32}) add(prog_identity(d), FAM_EW_TR, -1, d, 0, false);
  add(prog_vector_add(MXB_F32), FAM_EW_TR, -1, MXB_F32, 0, false);
}

}  // namespace mxbh
