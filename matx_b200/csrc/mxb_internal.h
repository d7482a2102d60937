// mxb_internal.h — host-side internals shared by the C-ABI (api.cu), the expression code generator
// (codegen.cpp), the NVRTC loader (jit.cpp) and the ahead-of-time manifest (aot_manifest.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/matx_b200.h"

namespace mxbh {

// ---- expression analysis -------------------------------------------------------------------------
struct ExprInfo {
  std::string sig;        // canonical structural signature (opcodes, operand ids, leaf / const dtypes)
  std::string name;       // E_<fnv64(sig)>
  std::string src;        // `struct <name> { ... };`
  int value_dtype = -1;   // arithmetic type the root evaluates to (bf16 / f16 leaves compute in f32)
  int nleaf = 0;
  int leaf_dtype[MXB_MAX_LEAVES] = {0};
  int max_leaf_bytes = 1; // widest leaf element
  int min_leaf_bytes = 16;
};

int dtype_bytes(int dtype);
const char *dtype_name(int dtype);    // "f32", ...
const char *dtype_ctype(int dtype);   // storage C type in device code
const char *reduce_op_name(int op);
uint64_t fnv64(const std::string &s);

// Analyse + generate.  Returns MXB_OK or an error status with `err` filled.
int analyze_expr(const mxb_expr_t *e, ExprInfo *info, std::string *err);

// ---- kernel instances ----------------------------------------------------------------------------
enum Family { FAM_RED_INNER = 0, FAM_RED_OUTER = 1, FAM_VAR_SMEM = 2, FAM_EW = 3, FAM_VAR_REG = 4, FAM_VAR_TMA = 5, FAM_VAR_GROUP = 6,
              FAM_SM_GROUP = 7, FAM_SM_REG = 8, FAM_EW_TR = 9, FAM_SCAN = 10, FAM_RED_OUTER_TMA = 11, FAM_SELECT = 12, FAM_UNUSED_13 = 13, FAM_HIST = 14 };
// internal kernel op: running (max, sum exp) of a row — the statistics pass of a long-row softmax
constexpr int KOP_LSE = MXB_RED_COUNT;
// internal kernel op: min and max with their indices in one read (mxb_argminmax)
constexpr int KOP_ARGMINMAX = MXB_RED_COUNT + 1;

struct KernelSpec {
  int family = 0;
  int op = -1;        // mxb_reduce_op_t for reductions (SUM also serves MEAN; VAR serves STDD), -1 for elementwise
  int out_dtype = 0;
  int V = 1, U = 1;
  int team = 0;       // FAM_RED_INNER: 0 = CTA per row, 1 = warp per row; FAM_VAR_REG: vectors per thread (IPT)
  int minb = 0;       // > 0: minimum resident CTAs per SM asked of the compiler (__launch_bounds__ second argument)
};

// unique key of (expression, spec); also yields the extern "C" symbol name
std::string kernel_key(const ExprInfo &info, const KernelSpec &spec);
std::string kernel_symbol(const std::string &key);
// source of the extern "C" __global__ wrapper for this instance (without the Expr struct)
int kernel_wrapper_src(const ExprInfo &info, const KernelSpec &spec, const std::string &symbol, std::string *out,
                       std::string *err);

// vector width / unroll policy shared by the dispatcher and the AOT manifest
int policy_vmax(const ExprInfo &info);
int policy_unroll(const ExprInfo &info, int V, int family);

// ---- registry: AOT table + JIT cache -------------------------------------------------------------
struct AotEntry { const char *key; const void *fn; };
void register_aot(const AotEntry *entries, int n);
const void *lookup_aot(const std::string &key);

// JIT: returns a cudaKernel_t-compatible handle usable with cudaLaunchKernel, or nullptr (+err)
const void *jit_get_kernel(const std::string &key, const std::string &symbol, const std::string &source,
                           std::string *err);
extern const char *const kDeviceHeaderText;  // mxb_device.cuh embedded at build time (for NVRTC)

// ---- built-in programs (AOT manifest and tests) ---------------------------------------------------
struct ManifestItem { mxb_expr_t expr; KernelSpec spec; };
void aot_manifest(std::vector<ManifestItem> *items);

}  // namespace mxbh

// ---- program utilities ---------------------------------------------------------------------------
namespace mxbh {
// Renumber a program into canonical form: post-order from the root (left operand first), identical
// sub-expressions / leaves (same pointer, dtype and strides) / constants merged.  Front ends may number
// nodes any way they like; the kernel registry and the AOT manifest only ever see canonical programs.
int canonicalize(const mxb_expr_t *in, mxb_expr_t *out, std::string *err);

// tiny builder used by the AOT manifest (the C++ shim and the Python mirror have their own)
struct ExprBuilder {
  mxb_expr_t e;
  ExprBuilder();
  int leaf(int dtype);
  int cst(double v, int dtype);
  int un(int opcode, int a, int aux = 0);
  int bin(int opcode, int a, int b);
  mxb_expr_t finish(int root);
};
mxb_expr_t prog_identity(int dtype);
mxb_expr_t prog_fma3(int dtype);           // a*b+c            (BASELINE config 1)
mxb_expr_t prog_abs2(int dtype);           // abs2(x)          (config 3: argmax(abs2(x)))
mxb_expr_t prog_black_scholes();           // examples/black_scholes.cu:122-138 (config 4)
mxb_expr_t prog_vector_add(int dtype);     // a + b            (bench/00_operators/operators.cu:10-36)
}  // namespace mxbh

namespace mxbh {
int jit_launch(const void *fn, unsigned grid, unsigned block, unsigned smem, void *stream, void *params, std::string *err, bool pdl = true, bool coop = false);
int jit_occupancy(const void *fn, unsigned block, unsigned smem);   // resident CTAs per SM of a JIT-built kernel (0 = unknown)
int jit_compile_only(const std::string &source, std::string *log);
}  // namespace mxbh
