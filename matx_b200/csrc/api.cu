// api.cu — the extern "C" boundary of libmatx_b200.so (declared in include/matx_b200.h): argument
// checking, view collapsing, kernel-family selection, workspace ownership and the launches.
//
// Host-side counterpart of, in the reference:
//   cudaExecutor::Exec + find_best_launch_params + get_grid_dims
//       (executors/cuda.h:84-230, executors/cuda_executor_common.h:325-493, core/get_grid_dims.h:45-174)
//   the *_impl(cudaExecutor) reductions and matxCubPlan_t
//       (transforms/reduce.h, transforms/cub.h:119-298,647-894,1281-1328)
//   ReduceInput / lcollapse / rcollapse (core/reduce_utils.h:45-102, operators/collapse.h)
// Differences by design: one launch per statement (no size-query call, no cudaMallocAsync per call, no
// second "single tile" launch, no separate scale kernel for mean), int64 sizes, deterministic results.
#include <cuda_runtime.h>
#include <cuda.h>  // CUtensorMap types only; the encoder's entry point is queried at run time

#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mxb_device.cuh"
#include "mxb_internal.h"
#include "mxb_sort.cuh"

using namespace mxbh;
using mxb::EwParams;
using mxb::KMAXD;
using mxb::RedParams;

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct mxb_context {
  // One handle = one stream + one set of scratch buffers.  Entry points take this lock, so a handle (a b200Executor and
  // its copies) may be used from several host threads: their statements are serialised onto the handle's stream, which
  // is what keeps the shared scratch safe (the reference keys its plan caches per thread, core/cache.h:224-245; threads
  // that want to overlap use one executor each, as there).
  std::recursive_mutex mu;
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int max_smem_optin = 227 * 1024;
  void *ws = nullptr;
  size_t ws_bytes = 0;
  unsigned *tickets = nullptr;
  size_t n_tickets = 0;
  void *tmp = nullptr;  // mean scratch of the two-launch variance; the radix sort's second key buffer
  size_t tmp_bytes = 0;
  void *tmp2 = nullptr; // unique: the sorted copy of the operand
  size_t tmp2_bytes = 0;
  unsigned *sort_ctr = nullptr;   // radix sort: per-(row, chunk, digit) counts and offsets
  size_t sort_ctr_words = 0;
  std::string last_kernel;
  int64_t launches = 0;
  // single-pass select / look-back scan: tile status words + {epoch, exit ticket} (device-resident epoch: graph-replay safe)
  void *lb_status = nullptr;
  size_t lb_cap_tiles = 0;
  unsigned *lb_ctl = nullptr;   // [0] epoch (starts at 1), [1] exit ticket
  // tensor maps of the TMA-tiled column reductions, keyed by everything the encoder sees: a repeated statement costs a
  // 96-byte compare instead of a driver call (round-robin over a few entries; maps are plain values, nothing to release)
  struct TmapEntry {
    unsigned long long key[12];
    CUtensorMap map;
    bool used = false;
  };
  TmapEntry tmaps[8];
  int tmap_next = 0;
};

namespace {

thread_local std::string g_err;

int fail(int st, const std::string &m) { g_err = m; return st; }
int cuda_fail(cudaError_t e, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return MXB_ERR_CUDA;
}
// MXB_PLAN_ONLY=1 (set before mxb_create): the host side runs as usual — view collapsing, kernel-family choice, launch
// geometry — but no CUDA call is made and nothing is launched; mxb_last_kernel() reports the kernel key plus
// grid / block / dynamic shared memory.  It exists so that the dispatch policy can be tested on a box without a GPU
// (tests/test_dispatch_plan.py); it computes nothing and is not a fallback.
bool g_plan = false;
#define MXB_CUDA(call) do { if (!g_plan) { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } } while (0)

int ensure_ws(mxb_context *h, size_t bytes, size_t tickets) {
  if (bytes > h->ws_bytes) {
    if (h->ws) MXB_CUDA(cudaFreeAsync(h->ws, h->stream));
    const size_t nb = std::max<size_t>(bytes, 64 * 1024);
    MXB_CUDA(cudaMallocAsync(&h->ws, nb, h->stream));
    h->ws_bytes = nb;
  }
  if (tickets > h->n_tickets) {
    if (h->tickets) MXB_CUDA(cudaFreeAsync(h->tickets, h->stream));
    const size_t nt = std::max<size_t>(tickets, 1024);
    MXB_CUDA(cudaMallocAsync((void **)&h->tickets, nt * sizeof(unsigned), h->stream));
    // zeroed once: the grid stage leaves every ticket at zero again (atomicInc wrap-around)
    MXB_CUDA(cudaMemsetAsync(h->tickets, 0, nt * sizeof(unsigned), h->stream));
    h->n_tickets = nt;
  }
  return MXB_OK;
}
int ensure_tmp(mxb_context *h, size_t bytes) {
  if (bytes > h->tmp_bytes) {
    if (h->tmp) MXB_CUDA(cudaFreeAsync(h->tmp, h->stream));
    const size_t nb = std::max<size_t>(bytes, 64 * 1024);
    MXB_CUDA(cudaMallocAsync(&h->tmp, nb, h->stream));
    h->tmp_bytes = nb;
  }
  return MXB_OK;
}

// Published-total slots of the look-back kernels (single-pass select, TILES-mode scan) and their control words
// {launch epoch, exit ticket}.  Two regions that never overlap: 8-byte slots {tag : 32-bit value} and 16-byte slots
// {tag, 64-bit value}, so a stale VALUE can never sit where a later launch reads a tag.  Zeroed once: a zero slot belongs
// to epoch 0, which no launch uses (the device-side epoch starts at 1 and skips 0 when it wraps).
int ensure_lb(mxb_context *h, size_t slots) {
  if (!h->lb_ctl) {
    MXB_CUDA(cudaMallocAsync((void **)&h->lb_ctl, 64, h->stream));
    MXB_CUDA(cudaMemsetAsync(h->lb_ctl, 0, 64, h->stream));
    static const unsigned one = 1;
    MXB_CUDA(cudaMemcpyAsync(h->lb_ctl, &one, sizeof one, cudaMemcpyHostToDevice, h->stream));
  }
  if (slots > h->lb_cap_tiles) {
    if (h->lb_status) MXB_CUDA(cudaFreeAsync(h->lb_status, h->stream));
    const size_t nt = (std::max<size_t>(slots + slots / 4, 16384) + 15) & ~(size_t)15;   // even: the 16-byte slots behind the 8-byte ones stay aligned
    MXB_CUDA(cudaMallocAsync(&h->lb_status, nt * 24, h->stream));
    MXB_CUDA(cudaMemsetAsync(h->lb_status, 0, nt * 24, h->stream));
    h->lb_cap_tiles = nt;
  }
  return MXB_OK;
}
void *lb_region(mxb_context *h, int slot_bytes) {   // 8-byte slots first, 16-byte slots behind them
  return slot_bytes == 8 ? h->lb_status : (void *)((char *)h->lb_status + h->lb_cap_tiles * 8);
}

// ---------------------------------------------------------------------------------------------------
// view collapsing: drop size-1 dims, merge neighbours that every leaf (and the outputs) walk contiguously
// ---------------------------------------------------------------------------------------------------
struct Group {
  int n = 0;
  int64_t size[MXB_MAX_RANK];
  int64_t ls[MXB_MAX_LEAVES][MXB_MAX_RANK];  // leaf strides
  int64_t os[MXB_MAX_RANK];                  // out strides (0 when unused)
  int64_t is[MXB_MAX_RANK];                  // index-out strides
};

void collapse(Group &g, int nleaf) {
  // 1. drop size-1 dims
  int m = 0;
  for (int d = 0; d < g.n; ++d) {
    if (g.size[d] == 1) continue;
    g.size[m] = g.size[d];
    for (int k = 0; k < nleaf; ++k) g.ls[k][m] = g.ls[k][d];
    g.os[m] = g.os[d];
    g.is[m] = g.is[d];
    ++m;
  }
  g.n = m;
  // 2. merge (d, d+1) when stride[d] == stride[d+1] * size[d+1] everywhere
  int d = 0;
  while (d + 1 < g.n) {
    bool ok = (g.os[d] == g.os[d + 1] * g.size[d + 1]) && (g.is[d] == g.is[d + 1] * g.size[d + 1]);
    for (int k = 0; ok && k < nleaf; ++k) ok = (g.ls[k][d] == g.ls[k][d + 1] * g.size[d + 1]);
    if (!ok) { ++d; continue; }
    g.size[d] *= g.size[d + 1];
    for (int k = 0; k < nleaf; ++k) g.ls[k][d] = g.ls[k][d + 1];
    g.os[d] = g.os[d + 1];
    g.is[d] = g.is[d + 1];
    for (int j = d + 1; j + 1 < g.n; ++j) {
      g.size[j] = g.size[j + 1];
      for (int k = 0; k < nleaf; ++k) g.ls[k][j] = g.ls[k][j + 1];
      g.os[j] = g.os[j + 1];
      g.is[j] = g.is[j + 1];
    }
    --g.n;
  }
  if (g.n == 0) {
    g.n = 1;
    g.size[0] = 1;
    for (int k = 0; k < nleaf; ++k) g.ls[k][0] = 0;
    g.os[0] = 0;
    g.is[0] = 0;
  }
}

void fill_consts(const mxb_expr_t &e, mxb::ConstDev &c) {
  memset(&c, 0, sizeof c);
  for (int k = 0; k < e.n_consts; ++k) {
    c.dre[k] = e.consts[k].re;
    c.dim[k] = e.consts[k].im;
    c.fre[k] = (float)e.consts[k].re;
    c.fim[k] = (float)e.consts[k].im;
    c.ire[k] = (long long)e.consts[k].re;
  }
}

void fill_peer_push(mxb::PeerPush &pp, const mxb_peers_t &peers, int item) {
  memset(&pp, 0, sizeof pp);
  for (int r = 0; r < peers.world; ++r) {
    pp.rec[r] = peers.rec[r];
    pp.flag[r] = (unsigned *)peers.flag[r];
  }
  pp.epoch = (const unsigned *)peers.epoch;
  pp.world = peers.world;
  pp.rank = peers.rank;
  pp.item = item;
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no link against libcuda; the library still
// loads on a CPU-only box).  nullptr when the driver does not have it.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return (EncodeTiledFn) nullptr;
    }
    return (EncodeTiledFn)f;
  }();
  return fn;
}

bool aligned_to(const void *p, int64_t bytes) { return ((uintptr_t)p % (uintptr_t)bytes) == 0; }
// Development / tuning knobs (MXB_TUNE_*, MXB_VAR_*, ...) are environment variables.  They are read ONCE per thread and
// knob — a dispatch used to cost ~15 getenv() scans of the environment — and again only after mxb_reload_env() (the
// Python test harness calls it when it has changed a knob).  Names are string literals: the cache is keyed by pointer.
std::atomic<unsigned> g_env_gen{1};
struct EnvCache {
  unsigned gen = 0;
  std::unordered_map<const void *, std::pair<bool, int>> v;   // literal -> (set, value)
};
const std::pair<bool, int> &env_lookup(const char *name) {
  thread_local EnvCache c;
  const unsigned g = g_env_gen.load(std::memory_order_relaxed);
  if (c.gen != g) { c.v.clear(); c.gen = g; }
  auto it = c.v.find((const void *)name);
  if (it != c.v.end()) return it->second;
  const char *v = getenv(name);
  return c.v.emplace((const void *)name, std::make_pair(v && *v, (v && *v) ? atoi(v) : 0)).first->second;
}
int env_int(const char *name, int dflt) {
  const std::pair<bool, int> &e = env_lookup(name);
  return e.first ? e.second : dflt;
}
bool env_set(const char *name) { return env_lookup(name).first; }

int acc_bytes(int op, int value_dtype) {
  switch (op) {
    case MXB_RED_ARGMAX: case MXB_RED_ARGMIN: return 16;
    case KOP_ARGMINMAX: return 32;
    case MXB_RED_ANY: case MXB_RED_ALL: return 4;
    case KOP_LSE: return 2 * dtype_bytes(value_dtype);
    case MXB_RED_VAR: return value_dtype == MXB_C64 ? 32 : 16;   // one-pass (pivot, s1, s2, count) state: 16 / 24 bytes
    default: return dtype_bytes(value_dtype);
  }
}

// ---------------------------------------------------------------------------------------------------
// kernel lookup (AOT table, then NVRTC) and launch
// ---------------------------------------------------------------------------------------------------
struct Kernel { const void *fn = nullptr; bool jit = false; std::string key; };

int get_kernel(const ExprInfo &info, const KernelSpec &spec, Kernel *k) {
  k->key = kernel_key(info, spec);
  const char *flavor = env_set("MXB_LD_FLAVOR") ? getenv("MXB_LD_FLAVOR") : nullptr;  // development knob: forces a JIT build with another load cache policy
  if (flavor && *flavor) k->key += std::string("|F") + flavor;
  k->fn = (flavor && *flavor) ? nullptr : lookup_aot(k->key);
  k->jit = false;
  if (k->fn) return MXB_OK;
  if (g_plan) {   // plan only: prove that the instance can be generated, do not build or load it
    std::string wrap, err;
    const int st = kernel_wrapper_src(info, spec, kernel_symbol(k->key), &wrap, &err);
    if (st != MXB_OK) return fail(st, err);
    k->jit = true;
    return MXB_OK;
  }
  if (env_set("MXB_DISABLE_JIT")) return fail(MXB_ERR_JIT, "no ahead-of-time kernel for " + k->key + " and MXB_DISABLE_JIT is set");
  const std::string sym = kernel_symbol(k->key);
  std::string wrap, err;
  int st = kernel_wrapper_src(info, spec, sym, &wrap, &err);
  if (st != MXB_OK) return fail(st, err);
  const std::string src = "#include \"mxb_device.cuh\"\n" + info.src + "\n" + wrap;
  k->fn = jit_get_kernel(k->key, sym, src, &err);
  if (!k->fn) return fail(MXB_ERR_JIT, "JIT of " + k->key + " failed: " + err);
  k->jit = true;
  return MXB_OK;
}

// coop: cooperative launch — the runtime guarantees that every CTA of the grid is resident at once (or refuses the
// launch), which the look-back kernels' tile waits rely on; such launches do not use programmatic dependent launch
template <class P>
int launch(mxb_context *h, const Kernel &k, unsigned grid, unsigned block, unsigned smem, P &params, bool coop = false) {
  if (grid == 0) return MXB_OK;
  if (g_plan) {
    h->launches++;
    h->last_kernel = k.key + (k.jit ? "|jit" : "|aot") + "|grid=" + std::to_string(grid) + "|block=" + std::to_string(block) + "|smem=" + std::to_string(smem);
    return MXB_OK;
  }
  if (k.jit) {
    std::string err;
    int st = jit_launch(k.fn, grid, block, smem, (void *)h->stream, (void *)&params, &err, !coop && env_int("MXB_PDL", 1) != 0, coop);
    if (st != MXB_OK) return fail(st, err);
  } else {
    // static + dynamic shared memory above 48 KB needs the opt-in; the kernels carry up to ~3 KB of static smem
    if (smem > 40 * 1024) MXB_CUDA(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {(void *)&params};
    if (coop) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(block);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = h->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      MXB_CUDA(cudaLaunchKernelExC(&cfg, k.fn, args));
    } else if (env_int("MXB_PDL", 1)) {
      // programmatic dependent launch: the kernel may be scheduled while its predecessor on the stream drains; every
      // kernel body starts with griddepcontrol.wait, so nothing is read or written before the predecessor has completed
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(block);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = h->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      MXB_CUDA(cudaLaunchKernelExC(&cfg, k.fn, args));
    } else {
      MXB_CUDA(cudaLaunchKernel(k.fn, dim3(grid), dim3(block), args, smem, h->stream));
    }
  }
  h->launches++;
  h->last_kernel = k.key + (k.jit ? "|jit" : "|aot");
  return MXB_OK;
}

// CTAs of `k` that fit on one SM at once (launch-time occupancy; `dflt` in plan-only mode or when the query fails)
int resident_ctas(const Kernel &k, unsigned block, unsigned smem, int dflt) {
  if (g_plan || !k.fn) return dflt;
  int occ = 0;
  if (k.jit) {
    occ = jit_occupancy(k.fn, block, smem);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k.fn, (int)block, smem) != cudaSuccess) { occ = 0; cudaGetLastError(); }
  }
  return occ > 0 ? occ : dflt;
}

// ---------------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------------
struct RedOptions {
  bool post_div = false;
  double post_scale = 1.0;
  bool post_sqrt = false;
  bool raw_partial = false;
  int64_t idx_base = 0;
  const mxb_peers_t *peers = nullptr;  // raw_partial + peers: push the record into every rank's exchange buffer
  int peer_item = 0;
  void *out2 = nullptr, *idx2 = nullptr;   // KOP_ARGMINMAX: the max value / max index outputs (strides of out / idx_out)
};

int check_expr_shape(const mxb_expr_t *e) {
  if (!e) return fail(MXB_ERR_INVALID, "null expression");
  if (e->rank < 0 || e->rank > MXB_MAX_RANK) return fail(MXB_ERR_INVALID, "expression rank out of range");
  for (int d = 0; d < e->rank; ++d)
    if (e->size[d] < 0) return fail(MXB_ERR_INVALID, "negative size");
  for (int k = 0; k < e->n_leaves && k < MXB_MAX_LEAVES; ++k)
    if (!e->leaves[k].data) return fail(MXB_ERR_INVALID, "leaf " + std::to_string(k) + " has a null data pointer");
  return MXB_OK;
}

// kernel op actually launched for a public op
int kernel_op(int op) {
  switch (op) {
    case MXB_RED_MEAN: return MXB_RED_SUM;
    case MXB_RED_STDD: return MXB_RED_VAR;
    default: return op;
  }
}

// One reduction launch (everything except the variance family).  `e` is canonical.
int reduce_launch(mxb_context *h, int kop, const mxb_expr_t &e, const ExprInfo &info, int n_reduce, const mxb_out_t *out,
                  const mxb_out_t *idx_out, const RedOptions &opt, bool var_smem = false) {
  const int nbd = e.rank - n_reduce;
  Group gb, gr;
  gb.n = nbd;
  gr.n = n_reduce;
  for (int d = 0; d < nbd; ++d) {
    gb.size[d] = e.size[d];
    for (int k = 0; k < e.n_leaves; ++k) gb.ls[k][d] = e.leaves[k].stride[d];
    gb.os[d] = (out && !opt.raw_partial) ? out->stride[d] : 0;
    gb.is[d] = idx_out ? idx_out->stride[d] : 0;
  }
  for (int d = 0; d < n_reduce; ++d) {
    gr.size[d] = e.size[nbd + d];
    for (int k = 0; k < e.n_leaves; ++k) gr.ls[k][d] = e.leaves[k].stride[nbd + d];
    gr.os[d] = 0;
    gr.is[d] = 0;
  }
  // size-1 batch dims carry no information for the output either
  collapse(gb, e.n_leaves);
  collapse(gr, e.n_leaves);
  if (gb.n > KMAXD || gr.n > KMAXD)
    return fail(MXB_ERR_NOT_SUPPORTED, "view does not collapse to <= 4 batch and <= 4 reduce dims");

  int64_t B = 1, R = 1;
  for (int d = 0; d < gb.n; ++d) B *= gb.size[d];
  for (int d = 0; d < gr.n; ++d) R *= gr.size[d];
  if (B == 0) return MXB_OK;
  if (R == 0) return fail(MXB_ERR_INVALID, "reduction over zero elements");

  const int out_dtype = opt.raw_partial ? (kop == MXB_RED_VAR ? (int)MXB_F32 : info.value_dtype) : out->dtype;
  const int vmax = env_int("MXB_TUNE_V", 0) > 0 ? env_int("MXB_TUNE_V", 0) : policy_vmax(info);
  const int nl = e.n_leaves;

  // ---- can the innermost reduce dim be the vector dim? ----
  auto inner_ok = [&](int V) {
    bool any_unit = false;
    for (int k = 0; k < nl; ++k) {
      const int64_t in = gr.ls[k][gr.n - 1];
      if (in != 0 && in != 1) return false;
      if (in == 0) continue;
      any_unit = true;
      const int64_t bytes = (int64_t)V * dtype_bytes(e.leaves[k].dtype);
      if (!aligned_to(e.leaves[k].data, bytes)) return false;
      for (int d = 0; d < gb.n; ++d) if (gb.ls[k][d] % V) return false;
      for (int d = 0; d + 1 < gr.n; ++d) if (gr.ls[k][d] % V) return false;
    }
    return any_unit;
  };
  // ---- or a batch dim (rotated last)? ----
  auto outer_dim = [&](int V) {
    for (int c = gb.n - 1; c >= 0; --c) {
      if (gb.size[c] < 2) continue;
      bool any_unit = false, ok = true;
      for (int k = 0; ok && k < nl; ++k) {
        const int64_t in = gb.ls[k][c];
        if (in != 0 && in != 1) { ok = false; break; }
        if (in == 0) continue;
        any_unit = true;
        const int64_t bytes = (int64_t)V * dtype_bytes(e.leaves[k].dtype);
        if (!aligned_to(e.leaves[k].data, bytes)) ok = false;
        for (int d = 0; ok && d < gb.n; ++d) if (d != c && gb.ls[k][d] % V) ok = false;
        for (int d = 0; ok && d < gr.n; ++d) if (gr.ls[k][d] % V) ok = false;
      }
      if (ok && any_unit) return c;
    }
    return -1;
  };

  KernelSpec spec;
  spec.op = kop;
  spec.out_dtype = out_dtype;
  int rot = -1;
  int var_ipt = 0, var_threads = 0, tma_ctas = 1, var_group_g = 0;
  // variance of rows that are NOT vectorisable along their innermost dim (strided / permuted rows, column variance):
  // the two-pass families would walk them one element per sector, so they go to the coalesced generic walkers with
  // the one-pass (mean, M2, n) op instead (fp32 / complex<float>)
  const bool chan_ok = kop == MXB_RED_VAR && (info.value_dtype == MXB_F32 || info.value_dtype == MXB_C64) && env_int("MXB_VAR_CHAN", 1);
  if (var_smem && chan_ok && vmax > 1 && !inner_ok(vmax) && outer_dim(vmax) >= 0 && R < (1ll << 31)) var_smem = false;
  if (var_smem) {
    spec.family = FAM_VAR_SMEM;
    spec.V = (vmax > 1 && inner_ok(vmax)) ? vmax : 1;
    // rows that fit in the registers of one CTA: blockDim.x * IPT vectors, single contiguous run, no ragged tail
    if (gr.n == 1 && inner_ok(spec.V) && gr.size[0] % spec.V == 0 && !env_set("MXB_VAR_SMEM_ONLY")) {
      const int64_t Lv = gr.size[0] / spec.V;
      if (Lv <= 256 && !env_set("MXB_VAR_NO_GROUP")) {
        // short rows: G lanes of a warp per row, the row in registers, shuffles only
        int G = 1;
        while (G < 32 && G < Lv) G <<= 1;
        int ipt = 1;
        while ((int64_t)G * ipt < Lv) ipt <<= 1;
        var_group_g = G;
        var_ipt = ipt;
      } else {
        const int vpt = env_int("MXB_TUNE_VAR_VPT", 4);   // vectors per thread the CTA size is chosen for
        int thr = 64;
        while (thr < 512 && (int64_t)thr * vpt < Lv) thr <<= 1;
        int ipt = 1;
        while ((int64_t)thr * ipt < Lv && ipt < 8) ipt <<= 1;
        if ((int64_t)thr * ipt >= Lv) { var_threads = thr; var_ipt = ipt; }
      }
      if (var_ipt) spec.family = var_group_g ? FAM_VAR_GROUP : FAM_VAR_REG;
    }
    // plain tensor, long contiguous rows: TMA-staged ring of rows in shared memory (one HBM read, copy engine keeps
    // rows in flight while the SM runs the two passes)
    if (e.n_nodes == 1 && nl == 1 && gr.n == 1 && gr.ls[0][0] == 1 && !env_set("MXB_VAR_SMEM_ONLY") && !env_set("MXB_VAR_NO_TMA")) {
      const int64_t esz = dtype_bytes(e.leaves[0].dtype);
      const int64_t rowbytes = gr.size[0] * esz;
      const int64_t rowstride = (rowbytes + 127) & ~int64_t(127);
      bool ok = rowbytes % 16 == 0 && rowbytes >= env_int("MXB_VAR_TMA_MIN_ROW", 16 * 1024) && aligned_to(e.leaves[0].data, 16) && rowbytes <= 128 * 1024;
      for (int d = 0; ok && d < gb.n; ++d) ok = (gb.ls[0][d] * esz) % 16 == 0;
      // Rows resident per SM = CTAs per SM x ring depth.  One CTA works on one row at a time, and a row costs a
      // fixed latency chain (barrier wait, two CTA reductions), so short rows want several CTAs per SM; long rows
      // only fit one CTA with a 2-3 deep ring (profiles/r1_sweeps.md).
      const int64_t per_sm = (int64_t)228 * 1024;
      int64_t ctas = 1, stages = 0;
      for (int64_t c = 4; c >= 1; --c) {
        const int64_t budget = std::min<int64_t>(per_sm / c - 1024 - 2048, (int64_t)h->max_smem_optin - 2048) - 128;
        int64_t st = budget / rowstride;
        if (st > 4) st = 4;
        if (st >= 2) { ctas = c; stages = st; break; }
      }
      if (env_int("MXB_TUNE_STAGES", 0) > 0 && env_int("MXB_TUNE_STAGES", 0) < stages) stages = env_int("MXB_TUNE_STAGES", 0);
      if (ok && stages >= 2) {
        spec.family = FAM_VAR_TMA;
        var_group_g = 0;
        spec.V = policy_vmax(info);
        var_ipt = (int)stages;
        tma_ctas = (int)ctas;
      }
    }
  } else if (vmax > 1 && inner_ok(vmax)) {
    spec.family = FAM_RED_INNER;
    spec.V = vmax;
  } else if (vmax > 1 && (rot = outer_dim(vmax)) >= 0) {
    spec.family = FAM_RED_OUTER;
    spec.V = vmax;
  } else if (vmax > 2 && inner_ok(vmax / 2)) {   // half-width vectors before giving up on them (8-byte aligned views)
    spec.family = FAM_RED_INNER;
    spec.V = vmax / 2;
  } else if (inner_ok(1)) {
    spec.family = FAM_RED_INNER;
    spec.V = 1;
  } else if ((rot = outer_dim(1)) >= 0) {
    spec.family = FAM_RED_OUTER;
    spec.V = 1;
  } else {
    spec.family = FAM_RED_INNER;  // V == 1 walks any strides
    spec.V = 1;
  }
  if (kop == KOP_ARGMINMAX && spec.family != FAM_RED_INNER) return MXB_ERR_NOT_SUPPORTED;   // mxb_argminmax runs argmin + argmax
  // ---- strided / permuted reduce dim of a plain tensor: TMA-staged tiles instead of the LDG column walker ----
  // (one collapsed reduce dim, at most one other batch dim, 16-byte granularity everywhere, enough strips to fill
  // the machine without splitting R).  Tile copies go through a tensor map (any pitch); without one, plain bulk
  // copies serve the case where a strip is whole contiguous rows (per-row bulk copies of 0.5-1 KB measured slower
  // than the LDG walker: profiles/r1_outer_tma_sweep.jsonl).
  int ot_tx = 0, ot_rt = 0, ot_stages = 0, ot_block = 256, ot_ctas = 2, ot_mode = 0;
  CUtensorMap ot_map;
  if (spec.family == FAM_RED_OUTER && spec.V > 1 && e.n_nodes == 1 && nl == 1 && gr.n == 1 && gb.n <= 2 && kop != KOP_LSE &&
      env_int("MXB_OUTER_TMA", 1)) {
    const int64_t esz = dtype_bytes(e.leaves[0].dtype);
    const int64_t C = gb.size[rot];
    const int64_t cv = C * esz / 16;                      // 16-byte chunks per row of the vector dim
    const int other = gb.n == 2 ? 1 - rot : -1;           // the other batch dim, if any
    const int64_t pitch = gr.ls[0][0] * esz;
    bool ok = gb.ls[0][rot] == 1 && (C * esz) % 16 == 0 && aligned_to(e.leaves[0].data, 16) && gr.ls[0][0] > 0 &&
              pitch % 16 == 0 && R >= env_int("MXB_OUTER_TMA_MIN_R", 64) && esz <= 16;
    if (ok && other >= 0) ok = gb.ls[0][other] > 0 && (gb.ls[0][other] * esz) % 16 == 0;
    const int want_mode = env_int("MXB_OUTER_TMA_MODE", -1);   // -1 auto, 0 bulk copies, 1 tensor map
    if (ok) {
      ot_block = env_int("MXB_TUNE_BLOCK", 0) > 0 ? env_int("MXB_TUNE_BLOCK", 0) : 256;
      ot_ctas = env_int("MXB_TUNE_OT_CTAS", 2);
      const int64_t grid_max = (int64_t)h->sm_count * ot_ctas;
      // strip width: the widest of 128 / 64 / 32 chunks whose item count spreads evenly over the persistent grid
      // (the ring makes every item cost the same, so the last wave's fill is the whole imbalance)
      double best = -1.0;
      for (int tx = 128; tx >= 32; tx >>= 1) {
        if (tx > ot_block) continue;
        int t = tx;
        while (t > 1 && t / 2 >= cv) t >>= 1;             // never wider than the row itself
        if (t < 32) break;
        const int64_t items = (B / C) * ((cv + t - 1) / t);
        if (items < (int64_t)h->sm_count) continue;
        const int64_t waves = (items + grid_max - 1) / grid_max;
        // ragged last strip of a row (cv % t) wastes its threads, not bandwidth: only the wave fill is scored
        const double fill = (double)items / (double)(waves * std::min<int64_t>(items, grid_max));
        if (fill > best + env_int("MXB_TUNE_OT_FILL_PCT", 4) * 0.01) { best = fill; ot_tx = t; }
        if (t != tx) break;
      }
      if (env_int("MXB_TUNE_TX", 0) >= 32 && env_int("MXB_TUNE_TX", 0) <= 128) ot_tx = env_int("MXB_TUNE_TX", 0);
    }
    if (ot_tx > 0) {
      const int ty = ot_block / ot_tx;
      const int64_t strip = (int64_t)ot_tx * 16;
      const int64_t spart = (int64_t)(ty - 1) * ot_tx * spec.V * acc_bytes(kop, info.value_dtype);
      const int64_t budget = std::min<int64_t>((int64_t)228 * 1024 / ot_ctas - 1024, (int64_t)h->max_smem_optin) - 1024 - 128 - spart;
      int64_t stage = (int64_t)env_int("MXB_TUNE_OT_STAGE_KB", 32) * 1024;
      int64_t want = env_int("MXB_TUNE_STAGES", 0) > 0 ? env_int("MXB_TUNE_STAGES", 0) : 3;
      while (stage > strip * ty && stage * want > budget) stage >>= 1;
      int64_t rt = std::max<int64_t>(ty, stage / strip / ty * ty);
      if (rt > ((R + ty - 1) / ty) * ty) rt = ((R + ty - 1) / ty) * ty;
      if (rt > 256) rt = 256 / ty * ty;                   // box dims of a tensor map are at most 256
      int64_t st = std::min<int64_t>(budget / (rt * strip), env_int("MXB_TUNE_STAGES", 0) > 0 ? env_int("MXB_TUNE_STAGES", 0) : 4);
      if (st > 8) st = 8;
      // tile copies: tensor map over {vector dim in 8-byte elements, reduce rows, other batch dim}
      bool have_map = false;
      EncodeTiledFn enc = (want_mode == 0 || g_plan) ? nullptr : tensor_map_encoder();
      if (enc && st >= 2) {
        const int64_t Bo = other >= 0 ? gb.size[other] : 1;
        const cuuint64_t gdim[3] = {(cuuint64_t)(C * esz / 8), (cuuint64_t)R, (cuuint64_t)Bo};
        const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)(other >= 0 ? gb.ls[0][other] * esz : pitch * R)};
        const cuuint32_t box[3] = {(cuuint32_t)(ot_tx * 2), (cuuint32_t)rt, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const int l2p = env_int("MXB_TUNE_OT_L2PROMO", 2);
        const bool fits = gdim[0] < (1ull << 31) && gdim[1] < (1ull << 31) && gdim[2] < (1ull << 31) && gstr[0] < (1ull << 40) && gstr[1] < (1ull << 40);
        if (fits) {
          const unsigned long long key[12] = {(unsigned long long)(uintptr_t)e.leaves[0].data, gdim[0], gdim[1], gdim[2], gstr[0], gstr[1],
                                              box[0], box[1], box[2], (unsigned long long)l2p, 3ull, (unsigned long long)CU_TENSOR_MAP_DATA_TYPE_UINT64};
          for (auto &t : h->tmaps)
            if (t.used && !memcmp(t.key, key, sizeof key)) { ot_map = t.map; have_map = true; break; }
          if (!have_map && enc(&ot_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(e.leaves[0].data), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2p,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
            have_map = true;
            auto &t = h->tmaps[h->tmap_next];
            h->tmap_next = (h->tmap_next + 1) % 8;
            memcpy(t.key, key, sizeof key);
            t.map = ot_map;
            t.used = true;
          }
        }
      }
      if (g_plan && want_mode != 0 && st >= 2) have_map = true;   // plan only: assume the driver encodes the map, as on the B200 box
      const bool contiguous_strips = pitch == C * esz && cv <= ot_tx;   // one bulk copy per stage
      if (st >= 2 && (have_map ? want_mode != 0 : (contiguous_strips || want_mode == 0))) {
        spec.family = FAM_RED_OUTER_TMA;
        spec.V = (int)(16 / esz);
        ot_rt = (int)rt;
        ot_stages = (int)st;
        ot_mode = have_map ? 1 : 0;
      }
    }
  }
  spec.U = policy_unroll(info, spec.V, spec.family);
  if (spec.family == FAM_VAR_REG || spec.family == FAM_VAR_GROUP) spec.team = var_ipt;
  // development knobs (tools/sweep.py): override the unroll / launch shape; any combination is JIT-compiled on demand
  const int tune_u = env_int("MXB_TUNE_U", 0), tune_block = env_int("MXB_TUNE_BLOCK", 0), tune_cps = env_int("MXB_TUNE_CTAS_PER_SM", 0);
  const int tune_tx = env_int("MXB_TUNE_TX", 0);
  int tma_block = 0;
  if (spec.family == FAM_VAR_TMA) {
    // 16-byte vectors of the row per thread, held in registers: IPT in {1,2,4,8} x up to 1024 threads
    const int64_t Rv16 = gr.size[0] * dtype_bytes(e.leaves[0].dtype) / 16;
    tma_block = 128;
    while (tma_block < 1024 && (int64_t)tma_block * 8 < Rv16) tma_block <<= 1;   // 8 vectors per thread when the row allows
    if (tune_block > 0) tma_block = tune_block;
    int ipt = 1;
    while ((int64_t)ipt * tma_block < Rv16) ipt <<= 1;
    if (ipt > 8) { tma_block = 1024; ipt = 8; }
    spec.team = ipt;
  }
  if (tune_u > 0 && spec.family != FAM_VAR_REG) spec.U = tune_u;

  RedParams p;
  memset(&p, 0, sizeof p);
  p.nb = gb.n;
  p.nr = gr.n;
  p.B = B;
  p.R = R;
  p.nleaf = nl;
  p.splits = 1;
  // original row-major weights of the batch dims
  int64_t bflat[MXB_MAX_RANK];
  {
    int64_t w = 1;
    for (int d = gb.n - 1; d >= 0; --d) { bflat[d] = w; w *= gb.size[d]; }
  }
  // batch dim order on the device (rot moved last for the outer family)
  int order[MXB_MAX_RANK];
  {
    int m = 0;
    for (int d = 0; d < gb.n; ++d) if (d != rot) order[m++] = d;
    if (rot >= 0) order[m++] = rot;
  }
  for (int i = 0; i < gb.n; ++i) {
    const int d = order[i];
    p.bsz[i] = gb.size[d];
    p.bflat[i] = bflat[d];
    p.out.bs[i] = gb.os[d];
    p.idx.bs[i] = gb.is[d];
    for (int k = 0; k < nl; ++k) p.leaf[k].bs[i] = gb.ls[k][d];
  }
  for (int d = 0; d < gr.n; ++d) {
    p.rsz[d] = gr.size[d];
    for (int k = 0; k < nl; ++k) p.leaf[k].rs[d] = gr.ls[k][d];
  }
  for (int k = 0; k < nl; ++k) p.leaf[k].ptr = e.leaves[k].data;
  p.out.ptr = out ? out->data : nullptr;
  p.idx.ptr = idx_out ? idx_out->data : nullptr;
  p.idx_base = opt.idx_base;
  p.out2 = opt.out2;
  p.idx2 = opt.idx2;
  p.post_div = opt.post_div ? 1 : 0;
  // one-pass variance through the generic walkers: the stored value is M2 / (N - ddof)
  if (kop == MXB_RED_VAR && (spec.family == FAM_RED_INNER || spec.family == FAM_RED_OUTER || spec.family == FAM_RED_OUTER_TMA)) p.post_div = 1;
  p.post_scale_f = (float)opt.post_scale;
  p.post_scale_d = opt.post_scale;
  p.post_sqrt = opt.post_sqrt ? 1 : 0;
  p.raw_partial = opt.raw_partial ? 1 : 0;
  if (opt.raw_partial && opt.peers) {
    p.raw_partial = 2;
    fill_peer_push(p.peer, *opt.peers, opt.peer_item);
  }
  fill_consts(e, p.c);
  const int vec_is_batch_dim = (rot >= 0);

  {
    // every leaf unit-stride along the vector dim -> the hot loops drop the per-leaf stride test
    bool unit = nl > 0;
    for (int k = 0; k < nl; ++k) {
      const int64_t in = vec_is_batch_dim ? p.leaf[k].bs[p.nb - 1] : p.leaf[k].rs[p.nr - 1];
      unit = unit && (in == 1);
    }
    p.all_unit = unit ? 1 : 0;
  }
  const int sm = h->sm_count;
  unsigned grid = 1, block = tune_block > 0 ? (unsigned)tune_block : 256u, smem = 0;
  // persistent grids: CTAs per SM x SM count (every CTA loops over its share of the rows / tiles)
  if (spec.family == FAM_VAR_TMA) {
    const int64_t rowstride = (R * dtype_bytes(e.leaves[0].dtype) + 127) & ~int64_t(127);
    p.splits = var_ipt;  // ring depth
    block = (unsigned)tma_block;
    smem = (unsigned)(128 + (int64_t)var_ipt * rowstride);
    grid = (unsigned)std::min<int64_t>(B, (int64_t)sm * (tune_cps > 0 ? tune_cps : tma_ctas));
  } else if (spec.family == FAM_VAR_GROUP) {
    p.tx = var_group_g;
    block = 256;
    const int64_t rows_per_cta = (int64_t)(block / 32) * (32 / var_group_g);
    grid = (unsigned)std::min<int64_t>((B + rows_per_cta - 1) / rows_per_cta, (int64_t)sm * (tune_cps > 0 ? tune_cps : 8));
  } else if (spec.family == FAM_VAR_REG) {
    block = (unsigned)var_threads;
    grid = (unsigned)std::min<int64_t>(B, (int64_t)sm * (tune_cps > 0 ? tune_cps : 32));
  } else if (spec.family == FAM_VAR_SMEM) {
    if (tune_block <= 0) block = R >= 4096 ? 512 : 256;
    smem = (unsigned)(R * dtype_bytes(info.value_dtype));
    grid = (unsigned)std::min<int64_t>(B, (int64_t)sm * (tune_cps > 0 ? tune_cps : 16));
  } else if (spec.family == FAM_RED_INNER) {
    const int64_t row_bytes = R * info.max_leaf_bytes;
    // rows under 32 KB (arg ops, whose CTA stage is the expensive one: under 128 KB): a warp — or a slice of one —
    // per row, no shared memory, no barrier; longer rows: a CTA, or several, per row (tools/shape_sweep*.py)
    const bool arg_op = (kop == MXB_RED_ARGMAX || kop == MXB_RED_ARGMIN || kop == KOP_ARGMINMAX);
    const int64_t t1_limit = env_int("MXB_TUNE_T1_BYTES", arg_op ? 131072 : 32768);
    spec.team = (row_bytes >= t1_limit || B < 8 * (int64_t)sm) ? 0 : 1;
    if (row_bytes < 4096) spec.team = 1;   // short rows never want a whole CTA
    if (env_int("MXB_TUNE_TEAM", -1) >= 0) spec.team = env_int("MXB_TUNE_TEAM", -1);
    // ops whose state is a multi-word record with a costly merge (one-pass variance, running log-sum-exp) keep the
    // round-1 shape: the per-item CTA stage is what they pay for (full-tensor var: 0.65 ms static vs 1.35 ms with 64 KB items)
    const bool heavy_merge = kop == MXB_RED_VAR || kop == KOP_LSE;
    // accumulators wider than two words (arg ops, variance) are compiled without the dynamic deal (mxb_device.cuh, DYN_OK):
    // they and log-sum-exp take the static round-robin deal of round 1
    const bool static_deal = heavy_merge || acc_bytes(kop, info.value_dtype) > 8;
    if (spec.team == 0 && tune_u <= 0 && env_int("MXB_TUNE_V", 0) <= 0 && nl <= 2 && gr.n == 1 && !heavy_merge) {
      // CTA-per-item streaming of one or two operands (profiles/r2_sweeps.md): four loads per leaf in flight, and — for
      // the ops whose per-element work is one add or compare (sum, prod, max, min) — 32-byte loads (fp32 2^30 sum: 0.5877
      // vs 0.5991 ms); the arg ops keep 16 bytes: their (value, index) compare chains want the registers
      spec.U = 4;
      const bool sum_like = kop == MXB_RED_SUM || kop == MXB_RED_PROD || kop == MXB_RED_MAX || kop == MXB_RED_MIN;   // one add / compare per element
      const int wide = 32 / info.max_leaf_bytes;
      if (sum_like && wide > spec.V && wide <= 8 && spec.V > 1 && inner_ok(wide)) spec.V = wide;
    }
    // many long rows, one CTA per row at a time: 128-thread CTAs (8 per SM) hide a row's CTA stage behind the other
    // rows' loads better than 4 CTAs of 256 (complex<float> 65536 x 8192 mean: 0.5829 vs 0.6080 ms)
    if (spec.team == 0 && tune_block <= 0 && B >= 2 * (int64_t)sm && !heavy_merge) block = 128;
    if (spec.team == 0) {
      const int64_t L = gr.size[gr.n - 1];
      const int64_t Q = (R / L) * (L / spec.V);  // vector steps per row
      int64_t S = 1;
      // CTAs that are co-resident per SM: the grid is exactly one resident wave and the work items are dealt
      // dynamically (RedParams::work_ctr), so no CTA ever waits behind another for an SM
      // the dual arg state: without a register cap the compiler takes 118 registers (two CTAs per SM); at 64 the few spills
      // sit on the rare path that replaces an extremum (2^30 fp32: 0.82 -> 0.65 ms)
      spec.minb = env_int("MXB_TUNE_MINB", kop == KOP_ARGMINMAX ? 4 : 0);
      Kernel kq;
      int st = get_kernel(info, spec, &kq);
      if (st != MXB_OK) return st;
      const int resident = resident_ctas(kq, block, 0, 4);
      int cps = tune_cps > 0 ? tune_cps : std::min(resident, 8);
      const int64_t grid_max = (int64_t)sm * cps;
      bool rr_static = false;
      if (B < 2 * (int64_t)sm) {
        // few long rows: a row is cut into S items.  One contiguous reduce run: an item is a contiguous chunk of tiles
        // (64 KB of the widest leaf; more for inputs so large that the final fold of the S partials would show), drawn
        // dynamically.  Several runs per row: S ranges of vector steps, one per resident CTA.
        if (gr.n == 1) {
          const int64_t nfull = Q / ((int64_t)block * spec.U);
          const int64_t want = env_int("MXB_TUNE_CHUNK_TILES", 0) > 0 ? env_int("MXB_TUNE_CHUNK_TILES", 0)
                                                                       : (B * nfull >= 8 * grid_max ? 4 : (B * nfull >= 2 * grid_max ? 2 : 1));
          int64_t cht = std::max<int64_t>(want, (nfull + 16383) / 16384);
          // the round-1 deal — tiles round-robin over sm x 8 static splits, no chunks (MXB_TUNE_RR=1 forces it for the A/B)
          if (static_deal || env_int("MXB_TUNE_RR", 0)) cht = 0;
          if (cht > 0) {
            S = std::max<int64_t>(1, (nfull + cht - 1) / cht);
          } else {
            rr_static = true;   // 8 CTAs per SM in two waves, one split each, tiles dealt round-robin to the splits
            S = std::max<int64_t>(1, std::min<int64_t>(((int64_t)sm * 8 + B - 1) / B, std::max<int64_t>(1, nfull / 2)));
          }
          p.chunk_tiles = (int)cht;
        } else {
          S = (grid_max + B - 1) / B;
          const int64_t maxS = std::max<int64_t>(1, Q / ((int64_t)block * spec.U * 2));
          S = std::max<int64_t>(1, std::min(S, maxS));
        }
      }
      p.splits = (int)S;
      grid = (unsigned)std::min<int64_t>(B * S, rr_static ? (int64_t)sm * 8 : grid_max);
      // dynamic deal for the chunks of split rows; whole rows (S == 1) stay on the static round-robin: their CTA stage is
      // per row either way and the draw only adds latency there (complex<float> 65536 x 8192 mean: 0.592 static, 0.614 dynamic)
      const bool dynamic = !rr_static && !static_deal && B * S > (int64_t)grid && (S > 1 || env_int("MXB_TUNE_DYNAMIC", 0) == 2) && env_int("MXB_TUNE_DYNAMIC", 1) && B * S < (1ll << 31);
      // combine exact in any order (max / min / arg / any / all, integer sums): one accumulator per CTA across its items
      const bool exact_op = kop == MXB_RED_MAX || kop == MXB_RED_MIN || kop == MXB_RED_ARGMAX || kop == MXB_RED_ARGMIN || kop == KOP_ARGMINMAX || kop == MXB_RED_ANY ||
                            kop == MXB_RED_ALL || ((kop == MXB_RED_SUM || kop == MXB_RED_PROD) && (info.value_dtype == MXB_I32 || info.value_dtype == MXB_I64));
      if (dynamic && B == 1 && S > 1 && exact_op && env_int("MXB_TUNE_CARRY", 1)) p.carry_items = 1;
      {
        int st2 = ensure_ws(h, S > 1 ? (size_t)std::max<int64_t>(B * S, grid) * (size_t)acc_bytes(kop, info.value_dtype) : 0, (size_t)(S > 1 ? B : 0) + 2);
        if (st2 != MXB_OK) return st2;
        if (S > 1) {
          p.ws = h->ws;
          p.tickets = h->tickets;
        }
        // the two words after the per-row tickets: work counter + exit ticket (both return to zero by themselves)
        if (dynamic) p.work_ctr = h->tickets + (S > 1 ? B : 0);
      }
      return launch(h, kq, grid, block, smem, p);
    } else {
      // G lanes per row: enough lanes for one vector step each, at most a warp; 32 / G rows share a warp
      const int64_t L = gr.size[gr.n - 1];
      const int64_t steps = (R / L) * std::max<int64_t>(1, (L + spec.V - 1) / spec.V);
      // up to `spl` vector steps per lane: adjacent lanes still read adjacent rows' bytes, and fewer lanes per row
      // means fewer shuffle rounds and more rows per warp
      const int spl = env_int("MXB_TUNE_STEPS_PER_LANE", 4);
      int G = 1;
      while (G < 32 && (int64_t)G * spl < steps) G <<= 1;
      if (env_int("MXB_TUNE_G", 0) > 0) G = env_int("MXB_TUNE_G", 0);
      p.tx = G;
      const int64_t rows_per_cta = (int64_t)(block / 32) * (32 / G);
      grid = (unsigned)std::min<int64_t>((B + rows_per_cta - 1) / rows_per_cta, (int64_t)sm * (tune_cps > 0 ? tune_cps : 8));
    }
  } else if (spec.family == FAM_RED_OUTER_TMA) {
    const int64_t C = p.bsz[p.nb - 1];
    const int64_t tile = (int64_t)ot_tx * spec.V;
    const int64_t items = (B / C) * ((C + tile - 1) / tile);
    const int ty = ot_block / ot_tx;
    p.tx = ot_tx;
    p.tma_rt = ot_rt;
    p.splits = ot_stages;
    p.tma_mode = ot_mode;
    if (ot_mode == 1) memcpy(p.tmap, &ot_map, sizeof ot_map);
    block = (unsigned)ot_block + 32u;   // consumers + the producer warp
    smem = (unsigned)(128 + (int64_t)ot_stages * ot_rt * ot_tx * 16 + (int64_t)(ty - 1) * ot_tx * spec.V * acc_bytes(kop, info.value_dtype));
    grid = (unsigned)std::min<int64_t>(items, (int64_t)sm * (tune_cps > 0 ? tune_cps : ot_ctas));
  } else {  // FAM_RED_OUTER
    const int64_t C = p.bsz[p.nb - 1];
    const int64_t cv = (C + spec.V - 1) / spec.V;
    int tx = 1;
    const int txmax = tune_tx > 0 ? tune_tx : 64;   // 64 x V columns per CTA, 4 reduce lanes (sweep: profiles/r1_sweeps.md)
    while (tx < txmax && tx < cv && tx < (int)block) tx <<= 1;
    p.tx = tx;
    const int ty = (int)block / tx;
    smem = ty > 1 ? (unsigned)(ty * tx * spec.V * acc_bytes(kop, info.value_dtype)) : 0;
    const int64_t tile = (int64_t)tx * spec.V;
    const int64_t tiles = (B / C) * ((C + tile - 1) / tile);
    // few output tiles, long reduce dim (column sums of a tall matrix): several CTAs per tile + in-launch combine
    int64_t S = 1;
    const int cps = tune_cps > 0 ? tune_cps : 8;
    if (tiles < 2 * (int64_t)sm) {
      S = ((int64_t)sm * cps + tiles - 1) / tiles;
      const int64_t maxS = std::max<int64_t>(1, R / ((int64_t)ty * spec.U * 4));
      S = std::max<int64_t>(1, std::min(S, maxS));
      if (env_int("MXB_TUNE_SPLITS", 0) > 0) S = env_int("MXB_TUNE_SPLITS", 0);
    }
    p.splits = (int)S;
    if (S > 1) {
      int st = ensure_ws(h, (size_t)(tiles * S * tile) * (size_t)acc_bytes(kop, info.value_dtype), (size_t)tiles);
      if (st != MXB_OK) return st;
      p.ws = h->ws;
      p.tickets = h->tickets;
    }
    grid = (unsigned)std::min<int64_t>(tiles * S, (int64_t)sm * cps);
  }

  Kernel k;
  int st = get_kernel(info, spec, &k);
  if (st != MXB_OK) return st;
  return launch(h, k, grid, block, smem, p);
}

bool is_floating(int t) { return t == MXB_F32 || t == MXB_F64 || t == MXB_C64; }

// ---- slab-sharded variance: the slab's (mean, M2 = sum |x - mean|^2, n) as one 32-byte record of doubles ----
struct VarRec { double mre, mim, m2, n; };
static_assert(sizeof(VarRec) == MXB_PARTIAL_BYTES, "variance record is one partial record");

template <class T, class RT>
__global__ void pack_var_kernel(const T *mean, const RT *m2, double n, VarRec *dst, int push, const __grid_constant__ mxb::PeerPush pp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  VarRec r;
  if constexpr (mxb::is_complex<T>::value) { r.mre = mean->re; r.mim = mean->im; }
  else { r.mre = (double)*mean; r.mim = 0.0; }
  r.m2 = (double)*m2;
  r.n = n;
  if (push) {
    union { VarRec v; mxb::PartialRec p; } u;
    u.v = r;
    mxb::push_record(pp, u.p);
  } else {
    *dst = r;
  }
}

// the slab's record: mean (one launch), sum |x - mean|^2 with the mean as a broadcast leaf (one launch), pack / push
int var_partial(mxb_context *h, const mxb_expr_t &e, const ExprInfo &info, const RedOptions &popt, void *record) {
  int64_t n = 1;
  for (int d = 0; d < e.rank; ++d) n *= e.size[d];
  if (n <= 0) return fail(MXB_ERR_INVALID, "variance of an empty slab");
  if ((info.value_dtype == MXB_F32 || info.value_dtype == MXB_C64) && env_int("MXB_VAR_CHAN", 1) && n < (1ll << 31)) {
    // ONE launch, one read: the one-pass state of the slab, written (or pushed to the peers) by the last CTA as the
    // (mean, M2, n) record of doubles the fold expects
    mxb_out_t rec_out;
    memset(&rec_out, 0, sizeof rec_out);
    rec_out.data = record;
    rec_out.dtype = MXB_F32;
    return reduce_launch(h, MXB_RED_VAR, e, info, e.rank, &rec_out, nullptr, popt, false);
  }
  int st = ensure_tmp(h, 64);
  if (st != MXB_OK) return st;
  const int rdt = info.value_dtype == MXB_F64 ? MXB_F64 : MXB_F32;
  mxb_out_t mean_out;
  memset(&mean_out, 0, sizeof mean_out);
  mean_out.data = h->tmp;
  mean_out.dtype = info.value_dtype;
  RedOptions mo;
  mo.post_div = true;
  mo.post_scale = (double)n;
  st = reduce_launch(h, MXB_RED_SUM, e, info, e.rank, &mean_out, nullptr, mo, false);
  if (st != MXB_OK) return st;
  if (e.n_leaves >= MXB_MAX_LEAVES || e.n_nodes + 3 > MXB_MAX_NODES) return fail(MXB_ERR_NOT_SUPPORTED, "expression too large for the sharded variance");
  mxb_expr_t e2 = e;
  const int lk = e2.n_leaves++;
  memset(&e2.leaves[lk], 0, sizeof e2.leaves[lk]);
  e2.leaves[lk].data = h->tmp;
  e2.leaves[lk].dtype = info.value_dtype;
  const int nleafnode = e2.n_nodes++;
  e2.nodes[nleafnode] = mxb_node_t{MXB_OP_LEAF, {lk, -1}, 0};
  const int nsub = e2.n_nodes++;
  e2.nodes[nsub] = mxb_node_t{MXB_OP_SUB, {e.root, nleafnode}, 0};
  const int nabs2 = e2.n_nodes++;
  e2.nodes[nabs2] = mxb_node_t{MXB_OP_ABS2, {nsub, -1}, 0};
  e2.root = nabs2;
  mxb_expr_t e2c;
  std::string err;
  st = canonicalize(&e2, &e2c, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info2;
  st = analyze_expr(&e2c, &info2, &err);
  if (st != MXB_OK) return fail(st, err);
  mxb_out_t m2_out;
  memset(&m2_out, 0, sizeof m2_out);
  m2_out.data = (char *)h->tmp + 16;
  m2_out.dtype = rdt;
  st = reduce_launch(h, MXB_RED_SUM, e2c, info2, e.rank, &m2_out, nullptr, RedOptions(), false);
  if (st != MXB_OK) return st;
  mxb::PeerPush pp;
  memset(&pp, 0, sizeof pp);
  const int push = popt.peers ? 1 : 0;
  if (push) fill_peer_push(pp, *popt.peers, popt.peer_item);
  VarRec *dst = (VarRec *)record;
  switch (info.value_dtype) {
    case MXB_F32: pack_var_kernel<float, float><<<1, 32, 0, h->stream>>>((const float *)h->tmp, (const float *)((char *)h->tmp + 16), (double)n, dst, push, pp); break;
    case MXB_F64: pack_var_kernel<double, double><<<1, 32, 0, h->stream>>>((const double *)h->tmp, (const double *)((char *)h->tmp + 16), (double)n, dst, push, pp); break;
    default: pack_var_kernel<mxb::cfloat, float><<<1, 32, 0, h->stream>>>((const mxb::cfloat *)h->tmp, (const float *)((char *)h->tmp + 16), (double)n, dst, push, pp); break;
  }
  MXB_CUDA(cudaGetLastError());
  h->launches++;
  return MXB_OK;
}

int reduce_impl(mxb_context *h, int op, const mxb_expr_t *expr_in, int n_reduce, const mxb_out_t *out, const mxb_out_t *idx_out,
                int ddof, const RedOptions *partial_opt) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  if (op < 0 || op >= MXB_RED_COUNT) return fail(MXB_ERR_INVALID, "unknown reduce op");
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (n_reduce < 0 || n_reduce > expr_in->rank) return fail(MXB_ERR_INVALID, "n_reduce_dims out of range");
  const int nbd = expr_in->rank - n_reduce;
  const bool arg = (op == MXB_RED_ARGMAX || op == MXB_RED_ARGMIN);
  if (!partial_opt) {
    if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
    if (out->rank != nbd) return fail(MXB_ERR_SIZE, "output rank must be expr rank - n_reduce_dims");
    for (int d = 0; d < nbd; ++d)
      if (out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
    if (out->dtype < 0 || out->dtype >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "output dtype out of range");
    if (arg) {
      if (!idx_out || !idx_out->data) return fail(MXB_ERR_INVALID, "argmax/argmin need an index output");
      if (idx_out->dtype != MXB_I64) return fail(MXB_ERR_INVALID, "index output must be MXB_I64 (matx::index_t)");
      if (idx_out->rank != nbd) return fail(MXB_ERR_SIZE, "index output rank mismatch");
      for (int d = 0; d < nbd; ++d)
        if (idx_out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "index output size mismatch");
    }
  }
  if (!arg) idx_out = nullptr;
  MXB_CUDA(cudaSetDevice(h->device));

  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  if (out && !partial_opt && info.value_dtype == MXB_C64 && out->dtype != MXB_C64 &&
      (op == MXB_RED_SUM || op == MXB_RED_MEAN || op == MXB_RED_PROD))
    return fail(MXB_ERR_INVALID, "complex reduction needs a complex output");

  RedOptions opt;
  if (partial_opt) opt = *partial_opt;
  int64_t R = 1;
  for (int d = nbd; d < e.rank; ++d) R *= e.size[d];

  if (op == MXB_RED_VAR || op == MXB_RED_STDD) {
    if (!is_floating(info.value_dtype)) return fail(MXB_ERR_NOT_SUPPORTED, "var/stdd of a non-floating expression");
    if (partial_opt) return var_partial(h, e, info, *partial_opt, out->data);
    if (out->dtype == MXB_C64) return fail(MXB_ERR_INVALID, "var/stdd output is real (the reference uses the inner type)");
    opt.post_scale = (double)(R - ddof);
    opt.post_sqrt = (op == MXB_RED_STDD);
    const int64_t row_bytes = R * dtype_bytes(info.value_dtype);
    // MXB_VAR_ONEPASS=1: every fp32 / complex<float> variance through the one-pass op (A/B knob, tools/var_onepass_ab.py)
    const bool onepass_all = (info.value_dtype == MXB_F32 || info.value_dtype == MXB_C64) && env_int("MXB_VAR_CHAN", 1) &&
                             env_int("MXB_VAR_ONEPASS", 0) && !env_set("MXB_VAR_TWO_LAUNCH") && !env_set("MXB_VAR_SMEM_ONLY") && R < (1ll << 31);
    // fp32 rows of 16..128 elements: the warp-team walker with the one-pass op beats the register-resident two-pass
    // group kernel (8388608x32: 0.54 vs 0.38 of peak, 4194304x64: 0.48 vs 0.36; profiles/r1_var_onepass_ab.jsonl)
    const bool onepass_short = info.value_dtype == MXB_F32 && env_int("MXB_VAR_CHAN", 1) && R >= env_int("MXB_VAR_ONEPASS_MIN_R", 16) &&
                               R <= env_int("MXB_VAR_ONEPASS_MAX_R", 128) && !env_set("MXB_VAR_TWO_LAUNCH") && !env_set("MXB_VAR_SMEM_ONLY") &&
                               !env_set("MXB_VAR_NO_GROUP");
    if (onepass_all || onepass_short) {
      opt.post_div = true;
      return reduce_launch(h, MXB_RED_VAR, e, info, n_reduce, out, nullptr, opt);
    }
    if (row_bytes <= (int64_t)h->max_smem_optin - 4096 && !env_set("MXB_VAR_TWO_LAUNCH")) {
      return reduce_launch(h, MXB_RED_VAR, e, info, n_reduce, out, nullptr, opt, /*var_smem=*/true);
    }
    // Row does not fit in shared memory.  fp32 / complex<float>: ONE read through the generic walkers with the
    // one-pass (mean, M2, n) op (Welford per thread, Chan's combine across threads / CTAs).
    if ((info.value_dtype == MXB_F32 || info.value_dtype == MXB_C64) && env_int("MXB_VAR_CHAN", 1) && !env_set("MXB_VAR_TWO_LAUNCH") &&
        R < (1ll << 31)) {   // the state counts elements in an int
      opt.post_div = true;
      return reduce_launch(h, MXB_RED_VAR, e, info, n_reduce, out, nullptr, opt);
    }
    // fp64: the reference's own scheme — mean, then the sum of |x - mean|^2 — as two launches over the same
    // skeletons (two reads of the input, like the reference).
    int64_t B = 1;
    for (int d = 0; d < nbd; ++d) B *= e.size[d];
    st = ensure_tmp(h, (size_t)std::max<int64_t>(B, 1) * (size_t)dtype_bytes(info.value_dtype));
    if (st != MXB_OK) return st;
    mxb_out_t mean_out;
    memset(&mean_out, 0, sizeof mean_out);
    mean_out.data = h->tmp;
    mean_out.dtype = info.value_dtype;
    mean_out.rank = nbd;
    {
      int64_t w = 1;
      for (int d = nbd - 1; d >= 0; --d) { mean_out.size[d] = e.size[d]; mean_out.stride[d] = w; w *= e.size[d]; }
    }
    RedOptions mo;
    mo.post_div = true;
    mo.post_scale = (double)R;
    st = reduce_launch(h, MXB_RED_SUM, e, info, n_reduce, &mean_out, nullptr, mo);
    if (st != MXB_OK) return st;
    if (e.n_leaves >= MXB_MAX_LEAVES || e.n_nodes + 3 > MXB_MAX_NODES) return fail(MXB_ERR_NOT_SUPPORTED, "expression too large for the two-launch variance");
    mxb_expr_t e2 = e;
    const int lk = e2.n_leaves++;
    memset(&e2.leaves[lk], 0, sizeof e2.leaves[lk]);
    e2.leaves[lk].data = h->tmp;
    e2.leaves[lk].dtype = info.value_dtype;
    for (int d = 0; d < nbd; ++d) e2.leaves[lk].stride[d] = mean_out.stride[d];
    const int nleafnode = e2.n_nodes++;
    e2.nodes[nleafnode] = mxb_node_t{MXB_OP_LEAF, {lk, -1}, 0};
    const int nsub = e2.n_nodes++;
    e2.nodes[nsub] = mxb_node_t{MXB_OP_SUB, {e.root, nleafnode}, 0};
    const int nabs2 = e2.n_nodes++;
    e2.nodes[nabs2] = mxb_node_t{MXB_OP_ABS2, {nsub, -1}, 0};
    e2.root = nabs2;
    mxb_expr_t e2c;
    st = canonicalize(&e2, &e2c, &err);
    if (st != MXB_OK) return fail(st, err);
    ExprInfo info2;
    st = analyze_expr(&e2c, &info2, &err);
    if (st != MXB_OK) return fail(st, err);
    opt.post_div = true;
    return reduce_launch(h, MXB_RED_SUM, e2c, info2, n_reduce, out, nullptr, opt);
  }

  if (op == MXB_RED_MEAN && !partial_opt) {
    if (!is_floating(info.value_dtype)) return fail(MXB_ERR_NOT_SUPPORTED, "mean of a non-floating expression");
    opt.post_div = true;
    opt.post_scale = (double)R;
  }
  return reduce_launch(h, kernel_op(op), e, info, n_reduce, out, idx_out, opt);
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU finalize: fold `world` 32-byte records in rank order, one thread
// ---------------------------------------------------------------------------------------------------
template <class Op, class OutT>
__global__ void finalize_kernel(const uint4 *recs, int world, int stride16, const __grid_constant__ RedParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  typename Op::acc_t a = Op::init();
  for (int r = 0; r < world; ++r) {
    union { uint4 q; typename Op::acc_t acc; } u;  // every acc_t fits the first 16 bytes of the 32-byte record
    u.q = recs[(size_t)r * stride16];
    Op::merge(a, u.acc);
  }
  mxb::store_result<Op, OutT>(p, 0, a);
}

// Chan's parallel combination of (n, mean, M2) triples, in fp64, rank order
__device__ inline double combine_var(const uint4 *recs, int world, int stride16, int ddof) {
  double n_tot = 0, mre = 0, mim = 0;
  for (int r = 0; r < world; ++r) {
    const volatile double *v = (const volatile double *)(recs + (size_t)r * stride16);
    n_tot += v[3];
    mre += v[3] * v[0];
    mim += v[3] * v[1];
  }
  mre /= n_tot;
  mim /= n_tot;
  double m2 = 0;
  for (int r = 0; r < world; ++r) {
    const volatile double *v = (const volatile double *)(recs + (size_t)r * stride16);
    const double dr = v[0] - mre, di = v[1] - mim;
    m2 += v[2] + v[3] * (dr * dr + di * di);
  }
  return m2 / (n_tot - (double)ddof);
}
template <class OutT>
__global__ void finalize_var_kernel(const uint4 *recs, int world, int stride16, int ddof, int take_sqrt, OutT *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v = combine_var(recs, world, stride16, ddof);
  if (take_sqrt) v = sqrt(v);
  *out = (OutT)v;
}

template <class T, class OutT>
int finalize_dispatch_op(mxb_context *h, int kop, const void *recs, int world, int stride16, RedParams &p) {
#define MXB_FIN(...) finalize_kernel<__VA_ARGS__, OutT><<<1, 32, 0, h->stream>>>((const uint4 *)recs, world, stride16, p)
  switch (kop) {
    case MXB_RED_SUM: MXB_FIN(mxb::OpSum<T>); break;
    case MXB_RED_PROD: MXB_FIN(mxb::OpProd<T>); break;
    case MXB_RED_ANY: MXB_FIN(mxb::OpLogic<T, true>); break;
    case MXB_RED_ALL: MXB_FIN(mxb::OpLogic<T, false>); break;
    default:
      if constexpr (!mxb::is_complex<T>::value) {
        switch (kop) {
          case MXB_RED_MAX: MXB_FIN(mxb::OpExt<T, true>); break;
          case MXB_RED_MIN: MXB_FIN(mxb::OpExt<T, false>); break;
          case MXB_RED_ARGMAX: MXB_FIN(mxb::OpArg<T, true>); break;
          case MXB_RED_ARGMIN: MXB_FIN(mxb::OpArg<T, false>); break;
          default: return fail(MXB_ERR_NOT_SUPPORTED, "finalize: op not supported");
        }
      } else {
        return fail(MXB_ERR_NOT_SUPPORTED, "finalize: op not supported for complex");
      }
  }
#undef MXB_FIN
  MXB_CUDA(cudaGetLastError());
  h->launches++;
  h->last_kernel = "finalize";
  return MXB_OK;
}

// ---------------------------------------------------------------------------------------------------
// fused exchange: wait for every rank's arrival counter, then fold all statements (one warp, one launch per step)
// ---------------------------------------------------------------------------------------------------
struct FoldItemDev { int op, dtype; void *out; long long *idx; int ddof, pad_; };
struct ExchangeParams {
  const uint4 *rec;        // this rank's buffer: PartialRec[2][world][KMAXITEMS]
  const unsigned *flag;    // this rank's arrival counters [world]
  unsigned *epoch;
  int world, n_items;
  double scale;            // MEAN divisor (global element count)
  FoldItemDev item[mxb::KMAXITEMS];
};

template <class Op> __device__ void fold_one(const uint4 *recs, int world, int stride16, void *out, long long *idx, bool div, double scale) {
  typename Op::acc_t a = Op::init();
  for (int r = 0; r < world; ++r) {
    union { uint4 q; typename Op::acc_t acc; } u;
    const volatile uint4 *src = recs + (size_t)r * stride16;   // written by peers during this launch: never from L1
    u.q.x = src->x; u.q.y = src->y; u.q.z = src->z; u.q.w = src->w;
    Op::merge(a, u.acc);
  }
  typename Op::result_t v = Op::finish(a);
  if (div) {
    RedParams fake;
    fake.post_div = 1; fake.post_sqrt = 0; fake.post_scale_d = scale; fake.post_scale_f = (float)scale;
    v = mxb::Post<typename Op::result_t>::go(v, fake);
  }
  *(typename Op::result_t *)out = v;
  if (Op::HAS_INDEX) *idx = Op::index(a);
}
template <class T> __device__ void fold_dtype(int op, const uint4 *recs, int world, int stride16, void *out, long long *idx, double scale) {
  switch (op) {
    case MXB_RED_SUM: fold_one<mxb::OpSum<T> >(recs, world, stride16, out, idx, false, scale); break;
    case MXB_RED_MEAN: fold_one<mxb::OpSum<T> >(recs, world, stride16, out, idx, true, scale); break;
    case MXB_RED_PROD: fold_one<mxb::OpProd<T> >(recs, world, stride16, out, idx, false, scale); break;
    default:
      if constexpr (!mxb::is_complex<T>::value) {
        switch (op) {
          case MXB_RED_MAX: fold_one<mxb::OpExt<T, true> >(recs, world, stride16, out, idx, false, scale); break;
          case MXB_RED_MIN: fold_one<mxb::OpExt<T, false> >(recs, world, stride16, out, idx, false, scale); break;
          case MXB_RED_ARGMAX: fold_one<mxb::OpArg<T, true> >(recs, world, stride16, out, idx, false, scale); break;
          case MXB_RED_ARGMIN: fold_one<mxb::OpArg<T, false> >(recs, world, stride16, out, idx, false, scale); break;
          default: break;
        }
      }
  }
}
// any / all records hold an int whatever the value type; the result is written in the value type
template <class T> __device__ void fold_logic(int op, const uint4 *recs, int world, int stride16, void *out) {
  int a = op == MXB_RED_ANY ? 0 : 1;
  for (int r = 0; r < world; ++r) {
    const int v = (int)((const volatile uint4 *)(recs + (size_t)r * stride16))->x;
    a = op == MXB_RED_ANY ? (a | v) : (a & v);
  }
  *(T *)out = mxb::cvt<T>(a);
}

// this rank's exchange control block (the zeroed words behind `epoch`): [0] completed exchanges, [1] sticky error word,
// [16 + r] records of source rank r consumed so far.  The arrival counters count records cumulatively, so a step waits for
// arrived[r] - consumed[r] >= n_items: steps of different sizes can share one exchange.
constexpr int EXC_ERR = 1, EXC_CONSUMED = 16;

__global__ void exchange_finalize_kernel(const __grid_constant__ ExchangeParams p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x;
  const unsigned e = *p.epoch + 1u;
  // lanes 0..world-1 each wait for one source rank (bounded: a dead peer must not hang this GPU)
  bool ok = true;
  if (lane < p.world) {
    const unsigned seen = p.epoch[EXC_CONSUMED + lane];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flag + lane) : "memory");
      if ((int)(v - seen - (unsigned)p.n_items) >= 0) break;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t1 - t0 > 5000000000ull) { ok = false; break; }  // 5 s
      __nanosleep(100);
    }
    p.epoch[EXC_CONSUMED + lane] = seen + (unsigned)p.n_items;
  }
  ok = __all_sync(0xffffffffu, ok);
  const int slot = (int)(e & 1u);
  if (!ok) {
    // a peer never delivered: nothing is folded (the slot may hold a stale record); the outputs get a sentinel and the
    // sticky error word makes mxb_exchange_check return MXB_ERR_CUDA
    if (lane == 0) p.epoch[EXC_ERR] = e;
    if (lane < p.n_items) {
      const FoldItemDev it = p.item[lane];
      const bool real64 = it.dtype == MXB_F64;
      if (it.dtype == MXB_F32 || it.dtype == MXB_C64 || real64) {
        if (real64) *(double *)it.out = __longlong_as_double(0x7ff8000000000000LL);
        else *(float *)it.out = __int_as_float(0x7fc00000);
      }
      if (it.idx) *it.idx = -1;
    }
  } else if (lane < p.n_items) {
    const FoldItemDev it = p.item[lane];
    const uint4 *recs = p.rec + ((size_t)slot * p.world * mxb::KMAXITEMS + lane) * 2;  // 2 x uint4 per 32-byte record
    const int stride16 = mxb::KMAXITEMS * 2;
    if (it.op == MXB_RED_VAR || it.op == MXB_RED_STDD) {
      double v = combine_var(recs, p.world, stride16, it.ddof);
      if (it.op == MXB_RED_STDD) v = sqrt(v);
      if (it.dtype == MXB_F64) *(double *)it.out = v;
      else *(float *)it.out = (float)v;   // fp32 and complex<float> inputs: fp32 result (the reference's inner type)
    } else if (it.op == MXB_RED_ANY || it.op == MXB_RED_ALL) {
      switch (it.dtype) {
        case MXB_F32: fold_logic<float>(it.op, recs, p.world, stride16, it.out); break;
        case MXB_F64: fold_logic<double>(it.op, recs, p.world, stride16, it.out); break;
        case MXB_I32: fold_logic<int>(it.op, recs, p.world, stride16, it.out); break;
        case MXB_I64: fold_logic<long long>(it.op, recs, p.world, stride16, it.out); break;
        default: break;
      }
    } else {
      switch (it.dtype) {
        case MXB_F32: fold_dtype<float>(it.op, recs, p.world, stride16, it.out, it.idx, p.scale); break;
        case MXB_F64: fold_dtype<double>(it.op, recs, p.world, stride16, it.out, it.idx, p.scale); break;
        case MXB_C64: fold_dtype<mxb::cfloat>(it.op, recs, p.world, stride16, it.out, it.idx, p.scale); break;
        case MXB_I32: fold_dtype<int>(it.op, recs, p.world, stride16, it.out, it.idx, p.scale); break;
        case MXB_I64: fold_dtype<long long>(it.op, recs, p.world, stride16, it.out, it.idx, p.scale); break;
        default: break;
      }
    }
  }
  __syncwarp();
  if (lane == 0) *p.epoch = e;
}

}  // namespace

// ===================================================================================================
// extern "C"
// ===================================================================================================
extern "C" {

int mxb_version(void) { return MXB_VERSION_MAJOR * 1000 + MXB_VERSION_MINOR; }
void mxb_reload_env(void) { g_env_gen.fetch_add(1); }
const char *mxb_last_error(void) { return g_err.c_str(); }

int mxb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mxb_create(mxb_handle_t *out_handle, void *stream) {
  if (!out_handle) return fail(MXB_ERR_INVALID, "null handle pointer");
  *out_handle = nullptr;
  if (env_int("MXB_PLAN_ONLY", 0)) {
    g_plan = true;
    mxb_context *hp = new mxb_context();   // B200 figures: 148 SMs, 227 KB of opt-in shared memory per CTA
    hp->device = -1;
    *out_handle = hp;
    return MXB_OK;
  }
  if (mxb_device_count() <= 0) return fail(MXB_ERR_NO_DEVICE, "no CUDA device visible: this library has no CPU fallback");
  mxb_context *h = new mxb_context();
  cudaError_t e = cudaGetDevice(&h->device);
  if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaGetDevice"); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, h->device);
  if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaGetDeviceProperties"); }
  if (prop.major != 10) {
    delete h;
    return fail(MXB_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                       ", this library is built for sm_100a (B200) only");
  }
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->stream = (cudaStream_t)stream;
  *out_handle = h;
  return MXB_OK;
}

int mxb_destroy(mxb_handle_t h) {
  if (!h) return MXB_OK;
  if (h->device >= 0) cudaSetDevice(h->device);
  if (h->ws) cudaFreeAsync(h->ws, h->stream);
  if (h->tickets) cudaFreeAsync(h->tickets, h->stream);
  if (h->tmp) cudaFreeAsync(h->tmp, h->stream);
  if (h->tmp2) cudaFreeAsync(h->tmp2, h->stream);
  if (h->sort_ctr) cudaFreeAsync(h->sort_ctr, h->stream);
  if (h->lb_status) cudaFreeAsync(h->lb_status, h->stream);
  if (h->lb_ctl) cudaFreeAsync(h->lb_ctl, h->stream);
  delete h;
  return MXB_OK;
}

int mxb_set_stream(mxb_handle_t h, void *stream) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  if ((cudaStream_t)stream != h->stream) {
    // scratch is stream-ordered: hand it over only once the old stream is done with it
    MXB_CUDA(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)stream;
  }
  return MXB_OK;
}

int mxb_sync(mxb_handle_t h) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  MXB_CUDA(cudaStreamSynchronize(h->stream));
  return MXB_OK;
}

const char *mxb_last_kernel(mxb_handle_t h) { return h ? h->last_kernel.c_str() : ""; }
int64_t mxb_launch_count(mxb_handle_t h) { return h ? h->launches : 0; }

int mxb_reduce(mxb_handle_t h, int reduce_op, const mxb_expr_t *expr, int n_reduce_dims, const mxb_out_t *out,
               const mxb_out_t *idx_out, int ddof) {
  return reduce_impl(h, reduce_op, expr, n_reduce_dims, out, idx_out, ddof, nullptr);
}

int mxb_argminmax(mxb_handle_t h, const mxb_expr_t *expr_in, int n_reduce, const mxb_out_t *out_min, const mxb_out_t *idx_min,
                  const mxb_out_t *out_max, const mxb_out_t *idx_max) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (n_reduce < 0 || n_reduce > expr_in->rank) return fail(MXB_ERR_INVALID, "n_reduce_dims out of range");
  const int nbd = expr_in->rank - n_reduce;
  const mxb_out_t *outs[4] = {out_min, idx_min, out_max, idx_max};
  for (int k = 0; k < 4; ++k) {
    if (!outs[k] || !outs[k]->data) return fail(MXB_ERR_INVALID, "argminmax needs four outputs");
    if (outs[k]->rank != nbd) return fail(MXB_ERR_SIZE, "output rank must be expr rank - n_reduce_dims");
    for (int d = 0; d < nbd; ++d)
      if (outs[k]->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
  }
  if (idx_min->dtype != MXB_I64 || idx_max->dtype != MXB_I64) return fail(MXB_ERR_INVALID, "index outputs must be MXB_I64 (matx::index_t)");
  if (out_min->dtype != out_max->dtype) return fail(MXB_ERR_INVALID, "the two value outputs must have one dtype");
  if (out_min->dtype < 0 || out_min->dtype >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "output dtype out of range");
  // one launch writes both pairs through the strides of the first: the min and max outputs must be laid out alike
  bool same_layout = true;
  for (int d = 0; d < nbd; ++d)
    if (out_min->size[d] > 1 && (out_min->stride[d] != out_max->stride[d] || idx_min->stride[d] != idx_max->stride[d])) same_layout = false;
  MXB_CUDA(cudaSetDevice(h->device));
  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  // the dual state rides the row walkers only (reduce_inner): a strided reduce dim (reduce_launch answers
  // MXB_ERR_NOT_SUPPORTED before launching anything), differently laid out output pairs and complex values take two launches
  if (same_layout && info.value_dtype != MXB_C64) {
    RedOptions opt;
    opt.out2 = out_max->data;
    opt.idx2 = idx_max->data;
    st = reduce_launch(h, KOP_ARGMINMAX, e, info, n_reduce, out_min, idx_min, opt);
    if (st != MXB_ERR_NOT_SUPPORTED) return st;
  }
  st = reduce_impl(h, MXB_RED_ARGMIN, expr_in, n_reduce, out_min, idx_min, 1, nullptr);
  if (st != MXB_OK) return st;
  return reduce_impl(h, MXB_RED_ARGMAX, expr_in, n_reduce, out_max, idx_max, 1, nullptr);
}

int mxb_reduce_partial(mxb_handle_t h, int reduce_op, const mxb_expr_t *expr, int64_t slab_offset, void *partial_record) {
  if (!partial_record) return fail(MXB_ERR_INVALID, "null partial record");
  if (!expr) return fail(MXB_ERR_INVALID, "null expression");
  RedOptions opt;
  opt.raw_partial = true;
  opt.idx_base = slab_offset;
  mxb_out_t o;
  memset(&o, 0, sizeof o);
  o.data = partial_record;
  mxb_out_t io = o;
  io.dtype = MXB_I64;
  int op = reduce_op == MXB_RED_MEAN ? MXB_RED_SUM : reduce_op;
  return reduce_impl(h, op, expr, expr->rank, &o, &io, 0, &opt);
}

int mxb_reduce_finalize(mxb_handle_t h, int reduce_op, int32_t value_dtype, const void *gathered_records, int world,
                        int64_t record_stride_bytes, int64_t global_count, int ddof, const mxb_out_t *out, const mxb_out_t *idx_out) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  if (!gathered_records || world <= 0) return fail(MXB_ERR_INVALID, "no records to fold");
  if (record_stride_bytes == 0) record_stride_bytes = MXB_PARTIAL_BYTES;
  if (record_stride_bytes < MXB_PARTIAL_BYTES || record_stride_bytes % 16) return fail(MXB_ERR_INVALID, "record stride must be a multiple of 16 and >= 32");
  const int rstride = (int)(record_stride_bytes / 16);
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  const bool arg = reduce_op == MXB_RED_ARGMAX || reduce_op == MXB_RED_ARGMIN;
  if (arg && (!idx_out || !idx_out->data || idx_out->dtype != MXB_I64)) return fail(MXB_ERR_INVALID, "argmax/argmin need an MXB_I64 index output");
  if (reduce_op == MXB_RED_VAR || reduce_op == MXB_RED_STDD) {
    MXB_CUDA(cudaSetDevice(h->device));
    const int sq = reduce_op == MXB_RED_STDD;
    if (out->dtype == MXB_F32) finalize_var_kernel<float><<<1, 32, 0, h->stream>>>((const uint4 *)gathered_records, world, rstride, ddof, sq, (float *)out->data);
    else if (out->dtype == MXB_F64) finalize_var_kernel<double><<<1, 32, 0, h->stream>>>((const uint4 *)gathered_records, world, rstride, ddof, sq, (double *)out->data);
    else return fail(MXB_ERR_INVALID, "var/stdd output is real (fp32 or fp64)");
    MXB_CUDA(cudaGetLastError());
    h->launches++;
    h->last_kernel = "finalize_var";
    return MXB_OK;
  }
  if (out->dtype != value_dtype) return fail(MXB_ERR_NOT_SUPPORTED, "finalize writes the value dtype only");
  MXB_CUDA(cudaSetDevice(h->device));
  RedParams p;
  memset(&p, 0, sizeof p);
  p.nb = 1; p.nr = 1; p.bsz[0] = 1; p.rsz[0] = 1; p.B = 1; p.R = global_count; p.bflat[0] = 1;
  p.out.ptr = out->data;
  p.idx.ptr = arg ? idx_out->data : nullptr;
  if (reduce_op == MXB_RED_MEAN) { p.post_div = 1; p.post_scale_f = (float)global_count; p.post_scale_d = (double)global_count; }
  const int kop = kernel_op(reduce_op);
  switch (value_dtype) {
    case MXB_F32: return finalize_dispatch_op<float, float>(h, kop, gathered_records, world, rstride, p);
    case MXB_F64: return finalize_dispatch_op<double, double>(h, kop, gathered_records, world, rstride, p);
    case MXB_C64: return finalize_dispatch_op<mxb::cfloat, mxb::cfloat>(h, kop, gathered_records, world, rstride, p);
    case MXB_I32: return finalize_dispatch_op<int, int>(h, kop, gathered_records, world, rstride, p);
    case MXB_I64: return finalize_dispatch_op<mxb::i64, mxb::i64>(h, kop, gathered_records, world, rstride, p);
  }
  return fail(MXB_ERR_NOT_SUPPORTED, "finalize: value dtype not supported");
}

int mxb_reduce_partial_push(mxb_handle_t h, int reduce_op, const mxb_expr_t *expr, int64_t slab_offset, const mxb_peers_t *peers,
                            int item, int n_items) {
  if (!expr) return fail(MXB_ERR_INVALID, "null expression");
  if (!peers || peers->world < 1 || peers->world > MXB_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world || !peers->epoch)
    return fail(MXB_ERR_INVALID, "bad peer table");
  if (item < 0 || item >= n_items || n_items > MXB_MAX_ITEMS) return fail(MXB_ERR_INVALID, "item index out of range");
  for (int r = 0; r < peers->world; ++r)
    if (!peers->rec[r] || !peers->flag[r]) return fail(MXB_ERR_INVALID, "peer buffer missing");
  RedOptions opt;
  opt.raw_partial = true;
  opt.idx_base = slab_offset;
  opt.peers = peers;
  opt.peer_item = item;
  mxb_out_t o;
  memset(&o, 0, sizeof o);
  o.data = peers->rec[peers->rank];
  mxb_out_t io = o;
  io.dtype = MXB_I64;
  const int op = reduce_op == MXB_RED_MEAN ? MXB_RED_SUM : reduce_op;
  return reduce_impl(h, op, expr, expr->rank, &o, &io, 0, &opt);
}

int mxb_exchange_alloc(mxb_handle_t h, size_t bytes, void **out_ptr, unsigned char ipc_handle_out[MXB_IPC_HANDLE_BYTES]) {
  if (!h || !out_ptr || !ipc_handle_out || bytes == 0) return fail(MXB_ERR_INVALID, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == MXB_IPC_HANDLE_BYTES, "IPC handle size");
  MXB_CUDA(cudaSetDevice(h->device));
  void *p = nullptr;
  MXB_CUDA(cudaMalloc(&p, bytes));
  MXB_CUDA(cudaMemset(p, 0, bytes));
  MXB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(ipc_handle_out, &hd, sizeof hd);
  *out_ptr = p;
  return MXB_OK;
}

int mxb_exchange_open(mxb_handle_t h, const unsigned char ipc_handle[MXB_IPC_HANDLE_BYTES], void **out_peer_ptr) {
  if (!h || !ipc_handle || !out_peer_ptr) return fail(MXB_ERR_INVALID, "bad arguments");
  MXB_CUDA(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, sizeof hd);
  void *p = nullptr;
  MXB_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *out_peer_ptr = p;
  return MXB_OK;
}

int mxb_exchange_close(mxb_handle_t h, void *peer_ptr) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  if (!peer_ptr) return MXB_OK;
  MXB_CUDA(cudaSetDevice(h->device));
  MXB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return MXB_OK;
}

int mxb_exchange_free(mxb_handle_t h, void *ptr) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  if (!ptr) return MXB_OK;
  MXB_CUDA(cudaSetDevice(h->device));
  MXB_CUDA(cudaFree(ptr));
  return MXB_OK;
}

int mxb_exchange_finalize(mxb_handle_t h, const mxb_peers_t *peers, const mxb_fold_item_t *items, int n_items, int64_t global_count) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  if (!peers || !items || n_items < 1 || n_items > MXB_MAX_ITEMS) return fail(MXB_ERR_INVALID, "bad exchange arguments");
  if (peers->world < 1 || peers->world > MXB_MAX_PEERS) return fail(MXB_ERR_INVALID, "bad peer table");
  MXB_CUDA(cudaSetDevice(h->device));
  ExchangeParams p;
  memset(&p, 0, sizeof p);
  p.rec = (const uint4 *)peers->rec[peers->rank];
  p.flag = (const unsigned *)peers->flag[peers->rank];
  p.epoch = (unsigned *)peers->epoch;
  p.world = peers->world;
  p.n_items = n_items;
  p.scale = (double)global_count;
  for (int k = 0; k < n_items; ++k) {
    const mxb_fold_item_t &it = items[k];
    const bool arg = it.reduce_op == MXB_RED_ARGMAX || it.reduce_op == MXB_RED_ARGMIN;
    if (!it.out || (arg && !it.idx_out)) return fail(MXB_ERR_INVALID, "null output in fold item");
    switch (it.reduce_op) {
      case MXB_RED_SUM: case MXB_RED_MEAN: case MXB_RED_PROD: case MXB_RED_MAX: case MXB_RED_MIN: case MXB_RED_ARGMAX:
      case MXB_RED_ARGMIN: case MXB_RED_ANY: case MXB_RED_ALL: case MXB_RED_VAR: case MXB_RED_STDD: break;
      default: return fail(MXB_ERR_NOT_SUPPORTED, "fold: reduce op not supported");
    }
    if (it.value_dtype != MXB_F32 && it.value_dtype != MXB_F64 && it.value_dtype != MXB_C64 && it.value_dtype != MXB_I32 && it.value_dtype != MXB_I64)
      return fail(MXB_ERR_NOT_SUPPORTED, "fold: value dtype not supported");
    if (it.value_dtype == MXB_C64 && it.reduce_op != MXB_RED_SUM && it.reduce_op != MXB_RED_MEAN && it.reduce_op != MXB_RED_PROD &&
        it.reduce_op != MXB_RED_VAR && it.reduce_op != MXB_RED_STDD)
      return fail(MXB_ERR_NOT_SUPPORTED, "fold: op not defined for complex");
    p.item[k].op = it.reduce_op;
    p.item[k].dtype = it.value_dtype;
    p.item[k].out = it.out;
    p.item[k].idx = (long long *)it.idx_out;
    p.item[k].ddof = it.ddof;
  }
  {
    // programmatic dependent launch: the fold kernel's launch latency hides behind the tail of the last slab kernel (it
    // starts with griddepcontrol.wait, so it reads nothing before that kernel has completed)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env_int("MXB_PDL", 1) ? 1 : 0;
    if (!g_plan) {
      void *args[] = {(void *)&p};
      MXB_CUDA(cudaLaunchKernelExC(&cfg, (const void *)exchange_finalize_kernel, args));
    }
  }
  h->launches++;
  h->last_kernel = "exchange_finalize";
  return MXB_OK;
}

int mxb_exchange_check(mxb_handle_t h, const mxb_peers_t *peers) {
  if (!h || !peers || !peers->epoch) return fail(MXB_ERR_INVALID, "bad exchange arguments");
  MXB_CUDA(cudaSetDevice(h->device));
  MXB_CUDA(cudaStreamSynchronize(h->stream));
  unsigned ctl[2] = {0, 0};
  MXB_CUDA(cudaMemcpy(ctl, peers->epoch, sizeof ctl, cudaMemcpyDeviceToHost));
  if (ctl[EXC_ERR] != 0)
    return fail(MXB_ERR_CUDA, "fused exchange: a peer's records did not arrive within 5 s at exchange " + std::to_string(ctl[EXC_ERR]) +
                                  " of " + std::to_string(ctl[0]) + "; the outputs of that step hold NaN / -1");
  return MXB_OK;
}

int mxb_elementwise(mxb_handle_t h, const mxb_expr_t *expr_in, const mxb_out_t *out) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  if (out->rank != expr_in->rank) return fail(MXB_ERR_SIZE, "output rank must equal the expression rank");
  for (int d = 0; d < out->rank; ++d)
    if (out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
  if (out->dtype < 0 || out->dtype >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "output dtype out of range");
  MXB_CUDA(cudaSetDevice(h->device));

  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  if (info.value_dtype == MXB_C64 && out->dtype != MXB_C64) return fail(MXB_ERR_INVALID, "complex expression needs a complex output");

  Group g;
  g.n = e.rank;
  for (int d = 0; d < e.rank; ++d) {
    g.size[d] = e.size[d];
    for (int k = 0; k < e.n_leaves; ++k) g.ls[k][d] = e.leaves[k].stride[d];
    g.os[d] = out->stride[d];
    g.is[d] = 0;
  }
  // An elementwise statement may walk its index space in any order: put the dims in the order the OUTPUT is laid out
  // (largest stride first), so that a permuted left-hand side `(y.Permute(...) = x)` is written along its own
  // contiguous dim like any other and the transposing family below sees the read side as the permuted one.
  {
    bool sorted = true;
    for (int d = 0; d + 1 < g.n; ++d) if (g.size[d] > 1 && g.size[d + 1] > 1 && g.os[d] < g.os[d + 1]) sorted = false;
    if (!sorted) {
      int order[MXB_MAX_RANK];
      for (int d = 0; d < g.n; ++d) order[d] = d;
      std::stable_sort(order, order + g.n, [&](int a, int b) { return g.os[a] > g.os[b]; });
      Group s2 = g;
      for (int d = 0; d < g.n; ++d) {
        g.size[d] = s2.size[order[d]];
        g.os[d] = s2.os[order[d]];
        g.is[d] = s2.is[order[d]];
        for (int k = 0; k < e.n_leaves; ++k) g.ls[k][d] = s2.ls[k][order[d]];
      }
    }
  }
  collapse(g, e.n_leaves);
  if (g.n > KMAXD) return fail(MXB_ERR_NOT_SUPPORTED, "view does not collapse to <= 4 dims");
  int64_t N = 1;
  for (int d = 0; d < g.n; ++d) N *= g.size[d];
  if (N == 0) return MXB_OK;

  const int nl = e.n_leaves;
  const int obytes = dtype_bytes(out->dtype);

  // ---- transposing family: some leaf is unit-stride along a dim that is not the output's -------------------
  if (g.n >= 2 && g.os[g.n - 1] == 1 && nl > 0 && env_int("MXB_NO_TRANSPOSE", 0) == 0) {
    const int X = g.n - 1;
    unsigned nonx = 0;
    for (int k = 0; k < nl; ++k)
      if (g.ls[k][X] != 0 && g.ls[k][X] != 1) nonx |= 1u << k;
    int Y = -1;
    int64_t best = 0;
    for (int d = 0; d < X; ++d) {
      int64_t score = 0;
      for (int k = 0; k < nl; ++k)
        if (((nonx >> k) & 1u) && g.ls[k][d] == 1) score += dtype_bytes(e.leaves[k].dtype);
      if (score > best) { best = score; Y = d; }
    }
    bool ok = nonx != 0 && Y >= 0;
    int eb = 0, staged = 0;
    for (int k = 0; ok && k < nl; ++k) {
      if (!((nonx >> k) & 1u)) continue;
      const int b = dtype_bytes(e.leaves[k].dtype);
      if (g.ls[k][Y] != 1 || (b != 2 && b != 4 && b != 8) || (eb != 0 && b != eb)) ok = false;
      eb = b;
      ++staged;
    }
    if (ok && staged > 3) ok = false;                       // shared-memory budget: <= 3 tiles
    if (ok) {
      const int V = 16 / eb, T = 16 * V;
      const int64_t SX = g.size[X], SY = g.size[Y];
      if (SX * 4 < T || SY * 4 < T) ok = false;             // tiles would be mostly empty: the row walker is as good
      if (ok && V * obytes > 64) ok = false;
      int64_t ntx = (SX + T - 1) / T, nty = (SY + T - 1) / T, outer = 1;
      for (int d = 0; d < X; ++d) if (d != Y) outer *= g.size[d];
      const int64_t ctas = ntx * nty * outer;
      if (ok && ctas > 0x7fffffff) ok = false;
      if (ok) {
        EwParams p;
        memset(&p, 0, sizeof p);
        p.nd = g.n;
        p.N = N;
        p.nleaf = nl;
        for (int d = 0; d < g.n; ++d) {
          p.sz[d] = g.size[d];
          p.out.bs[d] = g.os[d];
          for (int k = 0; k < nl; ++k) p.leaf[k].bs[d] = g.ls[k][d];
        }
        for (int k = 0; k < nl; ++k) p.leaf[k].ptr = e.leaves[k].data;
        p.out.ptr = out->data;
        p.tr_ydim = Y;
        p.tr_ymask = nonx;
        // alignment of the three kinds of vector access (chunk starts are multiples of V along X and Y by construction)
        bool yvec = true, xvec = true, ovec = aligned_to(out->data, (int64_t)V * obytes) && V * obytes <= 32;
        for (int d = 0; d < g.n; ++d) if (d != X && g.os[d] % V) ovec = false;
        for (int k = 0; k < nl; ++k) {
          const int b = dtype_bytes(e.leaves[k].dtype);
          if ((nonx >> k) & 1u) {
            if (!aligned_to(e.leaves[k].data, 16)) yvec = false;
            for (int d = 0; d < g.n; ++d) if (d != Y && g.ls[k][d] % V) yvec = false;
          } else if (g.ls[k][X] == 1) {
            if (!aligned_to(e.leaves[k].data, (int64_t)V * b)) xvec = false;
            for (int d = 0; d < g.n; ++d) if (d != X && g.ls[k][d] % V) xvec = false;
          }
        }
        p.tr_yvec = yvec ? 1 : 0;
        p.tr_xvec = xvec ? 1 : 0;
        p.tr_ovec = ovec ? 1 : 0;
        fill_consts(e, p.c);
        KernelSpec spec;
        spec.family = FAM_EW_TR;
        spec.op = -1;
        spec.out_dtype = out->dtype;
        spec.V = V;
        spec.U = 1;
        Kernel k;
        st = get_kernel(info, spec, &k);
        if (st != MXB_OK) return st;
        return launch(h, k, (unsigned)ctas, 256u, (unsigned)(staged * T * 256), p);
      }
    }
  }

  int vmax = env_int("MXB_TUNE_V", 0) > 0 ? env_int("MXB_TUNE_V", 0) : policy_vmax(info);
  // the store must fit one instruction too (STG.256 at most)
  while (vmax > 1 && vmax * obytes > 32) vmax >>= 1;
  auto vec_ok = [&](int V) {
    if (g.os[g.n - 1] != 1) return false;
    if (!aligned_to(out->data, (int64_t)V * obytes)) return false;
    for (int d = 0; d + 1 < g.n; ++d) if (g.os[d] % V) return false;
    for (int k = 0; k < nl; ++k) {
      const int64_t in = g.ls[k][g.n - 1];
      if (in != 0 && in != 1) return false;
      if (in == 0) continue;
      if (!aligned_to(e.leaves[k].data, (int64_t)V * dtype_bytes(e.leaves[k].dtype))) return false;
      for (int d = 0; d + 1 < g.n; ++d) if (g.ls[k][d] % V) return false;
    }
    return true;
  };
  KernelSpec spec;
  spec.family = FAM_EW;
  spec.op = -1;
  spec.out_dtype = out->dtype;
  spec.V = (vmax > 1 && vec_ok(vmax)) ? vmax : 1;
  spec.U = policy_unroll(info, spec.V, FAM_EW);
  if (env_int("MXB_TUNE_U", 0) > 0) spec.U = env_int("MXB_TUNE_U", 0);
  spec.team = env_int("MXB_TUNE_MINBLOCKS", -1) >= 0 ? env_int("MXB_TUNE_MINBLOCKS", -1) : (e.n_nodes >= 16 ? 4 : 0);

  EwParams p;
  memset(&p, 0, sizeof p);
  p.nd = g.n;
  p.N = N;
  p.nleaf = nl;
  for (int d = 0; d < g.n; ++d) {
    p.sz[d] = g.size[d];
    p.out.bs[d] = g.os[d];
    for (int k = 0; k < nl; ++k) p.leaf[k].bs[d] = g.ls[k][d];
  }
  for (int k = 0; k < nl; ++k) p.leaf[k].ptr = e.leaves[k].data;
  p.out.ptr = out->data;
  {
    bool unit = nl > 0;
    for (int k = 0; k < nl; ++k) unit = unit && (p.leaf[k].bs[g.n - 1] == 1);
    p.all_unit = unit ? 1 : 0;
  }
  fill_consts(e, p.c);

  const unsigned block = env_int("MXB_TUNE_BLOCK", 0) > 0 ? (unsigned)env_int("MXB_TUNE_BLOCK", 0) : 256u;
  const int64_t items = (N + spec.V - 1) / spec.V;
  const int64_t want = (items + (int64_t)block * spec.U - 1) / ((int64_t)block * spec.U);
  // one U-batch per thread (no grid-stride wrap) unless told otherwise: the hardware block scheduler balances the
  // SMs better than a static persistent loop for pure streaming (sweep: profiles/r1_sweeps.md)
  // ... except for arithmetic-heavy programs (Black-Scholes: 27 nodes), which are bound by instruction issue, not by
  // HBM: there a persistent grid of 32 CTAs per SM amortises the per-thread prologue (parameter loads, 64-bit index
  // setup, ~100 instructions) over many batches — measured 1.196 -> 1.138 ms on config 4 (profiles/r1_sweeps.md)
  const int ew_default_cps = e.n_nodes >= 16 ? 32 : 1000000;
  const int ew_cps = env_int("MXB_TUNE_CTAS_PER_SM", 0) > 0 ? env_int("MXB_TUNE_CTAS_PER_SM", 0) : ew_default_cps;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, (int64_t)h->sm_count * ew_cps), 0x7fffffff));
  Kernel k;
  st = get_kernel(info, spec, &k);
  if (st != MXB_OK) return st;
  return launch(h, k, grid, block, 0, p);
}

// ---------------------------------------------------------------------------------------------------
// softmax over the trailing n_reduce dims (reference: softmax_impl, transforms/reduce.h:362-445)
// ---------------------------------------------------------------------------------------------------
int mxb_softmax(mxb_handle_t h, const mxb_expr_t *expr_in, int n_reduce, const mxb_out_t *out) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (n_reduce < 1 || n_reduce > expr_in->rank) return fail(MXB_ERR_INVALID, "softmax needs 1 <= n_reduce_dims <= rank");
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  if (out->rank != expr_in->rank) return fail(MXB_ERR_SIZE, "softmax output rank must equal the input rank");
  for (int d = 0; d < out->rank; ++d)
    if (out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
  if (out->dtype != MXB_F32 && out->dtype != MXB_F64 && out->dtype != MXB_BF16 && out->dtype != MXB_F16)
    return fail(MXB_ERR_NOT_SUPPORTED, "softmax output must be a real floating type");
  MXB_CUDA(cudaSetDevice(h->device));

  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  if (info.value_dtype != MXB_F32 && info.value_dtype != MXB_F64)
    return fail(MXB_ERR_NOT_SUPPORTED, "softmax of a non-real-floating expression (the reference's max_impl rejects complex too)");

  const int nbd = e.rank - n_reduce, nl = e.n_leaves;
  Group gb, gr;
  gb.n = nbd;
  gr.n = n_reduce;
  for (int d = 0; d < nbd; ++d) {
    gb.size[d] = e.size[d];
    for (int k = 0; k < nl; ++k) gb.ls[k][d] = e.leaves[k].stride[d];
    gb.os[d] = out->stride[d];
    gb.is[d] = 0;
  }
  for (int d = 0; d < n_reduce; ++d) {
    gr.size[d] = e.size[nbd + d];
    for (int k = 0; k < nl; ++k) gr.ls[k][d] = e.leaves[k].stride[nbd + d];
    gr.os[d] = out->stride[nbd + d];
    gr.is[d] = 0;
  }
  collapse(gb, nl);
  collapse(gr, nl);
  int64_t B = 1, R = 1;
  for (int d = 0; d < gb.n; ++d) B *= gb.size[d];
  for (int d = 0; d < gr.n; ++d) R *= gr.size[d];
  if (B == 0 || R == 0) return MXB_OK;

  // ---- one launch, row in registers: a single reduce run that every leaf and the output walk with stride 0 / 1 ----
  const int obytes = dtype_bytes(out->dtype);
  bool rows_ok = gr.n == 1 && gb.n <= KMAXD && gr.os[0] == 1 && !env_set("MXB_SOFTMAX_TWO_LAUNCH");
  for (int k = 0; rows_ok && k < nl; ++k) rows_ok = (gr.ls[k][0] == 0 || gr.ls[k][0] == 1);
  if (rows_ok) {
    int vmax = env_int("MXB_TUNE_V", 0) > 0 ? env_int("MXB_TUNE_V", 0) : policy_vmax(info);
    while (vmax > 1 && vmax * obytes > 32) vmax >>= 1;
    auto vec_ok = [&](int V) {
      if (R % V) return false;
      if (!aligned_to(out->data, (int64_t)V * obytes)) return false;
      for (int d = 0; d < gb.n; ++d) if (gb.os[d] % V) return false;
      for (int k = 0; k < nl; ++k) {
        if (gr.ls[k][0] == 0) continue;
        if (!aligned_to(e.leaves[k].data, (int64_t)V * dtype_bytes(e.leaves[k].dtype))) return false;
        for (int d = 0; d < gb.n; ++d) if (gb.ls[k][d] % V) return false;
      }
      return true;
    };
    int V = vmax;
    while (V > 1 && !vec_ok(V)) V >>= 1;
    const int64_t Lv = R / V;
    KernelSpec spec;
    spec.op = -1;
    spec.out_dtype = out->dtype;
    spec.V = V;
    spec.U = 1;
    int G = 0, thr = 0, ipt = 0;
    const int max_ipt = env_int("MXB_SOFTMAX_MAX_IPT", 4);   // vectors per lane held in registers (8 spills under the budgets)
    if (Lv <= 32 * max_ipt) {
      G = 1;
      while (G < 32 && G < Lv) G <<= 1;
      ipt = 1;
      while ((int64_t)G * ipt < Lv) ipt <<= 1;
      spec.family = FAM_SM_GROUP;
    } else {
      const int vpt = env_int("MXB_TUNE_VAR_VPT", 4);
      thr = 64;
      while (thr < 1024 && (int64_t)thr * vpt < Lv) thr <<= 1;
      ipt = 1;
      while ((int64_t)thr * ipt < Lv && ipt < max_ipt) ipt <<= 1;
      if ((int64_t)thr * ipt >= Lv) spec.family = FAM_SM_REG;
      else ipt = 0;
    }
    if (ipt) {
      spec.team = ipt;
      RedParams p;
      memset(&p, 0, sizeof p);
      p.nb = gb.n;
      p.nr = 1;
      p.B = B;
      p.R = R;
      p.nleaf = nl;
      p.splits = 1;
      for (int d = 0; d < gb.n; ++d) {
        p.bsz[d] = gb.size[d];
        p.out.bs[d] = gb.os[d];
        for (int k = 0; k < nl; ++k) p.leaf[k].bs[d] = gb.ls[k][d];
      }
      p.rsz[0] = R;
      p.out_rs[0] = gr.os[0];
      bool unit = nl > 0;
      for (int k = 0; k < nl; ++k) {
        p.leaf[k].rs[0] = gr.ls[k][0];
        p.leaf[k].ptr = e.leaves[k].data;
        unit = unit && gr.ls[k][0] == 1;
      }
      p.all_unit = unit ? 1 : 0;
      p.out.ptr = out->data;
      fill_consts(e, p.c);
      const int sm = h->sm_count;
      const int tune_cps = env_int("MXB_TUNE_CTAS_PER_SM", 0);
      unsigned grid, block;
      if (spec.family == FAM_SM_GROUP) {
        p.tx = G;
        block = 256;
        const int64_t rows_per_cta = (int64_t)(block / 32) * (32 / G);
        grid = (unsigned)std::min<int64_t>((B + rows_per_cta - 1) / rows_per_cta, (int64_t)sm * (tune_cps > 0 ? tune_cps : 8));
      } else {
        block = (unsigned)thr;
        grid = (unsigned)std::min<int64_t>(B, (int64_t)sm * (tune_cps > 0 ? tune_cps : 32));
      }
      Kernel k;
      st = get_kernel(info, spec, &k);
      if (st != MXB_OK) return st;
      return launch(h, k, grid, block, 0, p);
    }
  }

  // ---- any other shape: statistics launch (running max / sum of exp through the ordinary reduction families, so long
  // rows split over CTAs and strided reduce dims ride reduce_outer), then out = exp(x - max) * (1 / sum) as an elementwise
  // statement with the two statistics as broadcast leaves.  Two reads (the second mostly from L2), one write; the
  // reference does three reads, two temporaries and five launches.
  const int64_t esz = dtype_bytes(info.value_dtype);
  const size_t s_off = ((size_t)(B * esz) + 255) & ~(size_t)255;
  st = ensure_tmp(h, 2 * s_off);
  if (st != MXB_OK) return st;
  mxb_out_t m_out;
  memset(&m_out, 0, sizeof m_out);
  m_out.data = h->tmp;
  m_out.dtype = info.value_dtype;
  m_out.rank = nbd;
  {
    int64_t w = 1;
    for (int d = nbd - 1; d >= 0; --d) { m_out.size[d] = e.size[d]; m_out.stride[d] = w; w *= e.size[d]; }
  }
  mxb_out_t s_out = m_out;
  s_out.data = (char *)h->tmp + s_off;
  st = reduce_launch(h, KOP_LSE, e, info, n_reduce, &m_out, &s_out, RedOptions());
  if (st != MXB_OK) return st;
  if (e.n_leaves + 2 > MXB_MAX_LEAVES || e.n_nodes + 5 > MXB_MAX_NODES) return fail(MXB_ERR_NOT_SUPPORTED, "expression too large for the two-launch softmax");
  mxb_expr_t e2 = e;
  int stat_node[2];
  for (int j = 0; j < 2; ++j) {
    const int lk = e2.n_leaves++;
    memset(&e2.leaves[lk], 0, sizeof e2.leaves[lk]);
    e2.leaves[lk].data = j == 0 ? m_out.data : s_out.data;
    e2.leaves[lk].dtype = info.value_dtype;
    for (int d = 0; d < nbd; ++d) e2.leaves[lk].stride[d] = m_out.stride[d];
    stat_node[j] = e2.n_nodes++;
    e2.nodes[stat_node[j]] = mxb_node_t{MXB_OP_LEAF, {lk, -1}, 0};
  }
  const int nsub = e2.n_nodes++;
  e2.nodes[nsub] = mxb_node_t{MXB_OP_SUB, {e.root, stat_node[0]}, 0};
  const int nexp = e2.n_nodes++;
  e2.nodes[nexp] = mxb_node_t{MXB_OP_EXP, {nsub, -1}, 0};
  const int ndiv = e2.n_nodes++;
  e2.nodes[ndiv] = mxb_node_t{MXB_OP_MUL, {nexp, stat_node[1]}, 0};   // stat 1 holds 1 / sum
  e2.root = ndiv;
  return mxb_elementwise(h, &e2, out);
}

// ---------------------------------------------------------------------------------------------------
// cumsum: inclusive prefix sum along the innermost dim (reference: cumsum_impl, transforms/cub.h:2367-2395 ->
// ExecPrefixScanEx :375-408: cub::DeviceScan::InclusiveSum, one launch per row for rank > 1)
// ---------------------------------------------------------------------------------------------------
int mxb_cumsum(mxb_handle_t h, const mxb_expr_t *expr_in, const mxb_out_t *out) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (expr_in->rank < 1) return fail(MXB_ERR_INVALID, "cumsum needs rank >= 1");
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  if (out->rank != expr_in->rank) return fail(MXB_ERR_SIZE, "cumsum output rank must equal the input rank");
  for (int d = 0; d < out->rank; ++d)
    if (out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
  if (out->dtype < 0 || out->dtype >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "output dtype out of range");
  MXB_CUDA(cudaSetDevice(h->device));

  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  const int vt = info.value_dtype;
  if (!(vt == MXB_F32 || vt == MXB_F64 || vt == MXB_C64 || vt == MXB_I32 || vt == MXB_I64))
    return fail(MXB_ERR_NOT_SUPPORTED, "cumsum of this value type is not lowered");
  if (vt == MXB_C64 && out->dtype != MXB_C64) return fail(MXB_ERR_INVALID, "complex expression needs a complex output");

  const int nbd = e.rank - 1, nl = e.n_leaves;
  Group gb;
  gb.n = nbd;
  for (int d = 0; d < nbd; ++d) {
    gb.size[d] = e.size[d];
    for (int k = 0; k < nl; ++k) gb.ls[k][d] = e.leaves[k].stride[d];
    gb.os[d] = out->stride[d];
    gb.is[d] = 0;
  }
  collapse(gb, nl);
  if (gb.n > KMAXD) return fail(MXB_ERR_NOT_SUPPORTED, "batch dims do not collapse to <= 4");
  const int64_t L = e.size[e.rank - 1];
  int64_t B = 1;
  for (int d = 0; d < gb.n; ++d) B *= gb.size[d];
  if (B == 0 || L == 0) return MXB_OK;
  const int64_t ostride = out->stride[e.rank - 1];
  const int obytes = dtype_bytes(out->dtype);

  // vector width: every leaf unit-stride (or broadcast) along the scan dim, rows 16-byte aligned
  int V = policy_vmax(info);
  while (V > 1 && V * obytes > 32) V >>= 1;
  auto in_ok = [&](int v) {
    for (int k = 0; k < nl; ++k) {
      const int64_t in = e.leaves[k].stride[e.rank - 1];
      if (in != 0 && in != 1) return false;
      if (in == 0) continue;
      if (!aligned_to(e.leaves[k].data, (int64_t)v * dtype_bytes(e.leaves[k].dtype))) return false;
      for (int d = 0; d < gb.n; ++d) if (gb.ls[k][d] % v) return false;
    }
    return true;
  };
  while (V > 1 && !in_ok(V)) V = 1;
  bool ovec = V > 1 && ostride == 1 && aligned_to(out->data, (int64_t)V * obytes);
  for (int d = 0; ovec && d < gb.n; ++d) if (gb.os[d] % V) ovec = false;

  KernelSpec spec;
  spec.family = FAM_SCAN;
  spec.op = -1;
  spec.out_dtype = out->dtype;
  spec.V = V;
  // tile = 256 threads x U chunks x V elements; short rows take the one-chunk instance
  // (8-byte values: two chunks, or the 16 running values per thread spill)
  spec.U = (L > (int64_t)256 * V) ? (dtype_bytes(vt) >= 8 ? 2 : 4) : 1;
  if (env_int("MXB_TUNE_U", 0) > 0) spec.U = env_int("MXB_TUNE_U", 0);
  spec.minb = env_int("MXB_TUNE_MINB", 0);
  // short rows, many of them: a warp per row (no shared memory, no barrier)
  const bool warp_team = env_int("MXB_TUNE_TEAM", -1) >= 0 ? env_int("MXB_TUNE_TEAM", -1) == 1
                                                          : (L <= (int64_t)2048 && B >= 4 * (int64_t)h->sm_count);
  if (warp_team) {
    spec.team = 1;
    // two vector steps per lane and pass (one for 8-byte values): four spill at the 3-CTAs-per-SM register budget
    spec.U = (dtype_bytes(vt) >= 8 || V > 4 || L <= (int64_t)32 * V) ? 1 : 2;
  }
  const int64_t tile = (int64_t)256 * V * spec.U;
  const int64_t tpr = (L + tile - 1) / tile;
  const int sm = h->sm_count;
  // few long rows: one CTA per tile with the in-launch carry exchange; otherwise a CTA walks whole rows
  bool tiles_mode = !warp_team && tpr > 1 && B < 2 * (int64_t)sm;
  if (env_int("MXB_SCAN_MODE", 0) == 1) tiles_mode = false;
  if (env_int("MXB_SCAN_MODE", 0) == 2) tiles_mode = tpr > 1;
  if (tiles_mode && B * tpr > 0x7fffffff) tiles_mode = false;

  RedParams p;
  memset(&p, 0, sizeof p);
  p.nb = gb.n;
  p.nr = 1;
  p.B = B;
  p.R = L;
  p.nleaf = nl;
  // TILES mode exchange: 2 = flat two-level gather by the whole CTA (the measured default), 3 = pipelined three-level jobs by one warp
  const bool scan_pipe = env_int("MXB_SCAN_PIPELINE", 0) != 0;
  // 4 = warp tiles (4 KB per warp), phase 2 re-reads its tile through L2: no ring, no barrier
  const bool scan_wt = tiles_mode && !scan_pipe && env_int("MXB_SCAN_WTILES", 1) != 0;   // 0: the flat CTA-tile exchange of round 1
  p.splits = tiles_mode ? (scan_wt ? 4 : (scan_pipe ? 3 : 2)) : 1;
  for (int d = 0; d < gb.n; ++d) {
    p.bsz[d] = gb.size[d];
    p.out.bs[d] = gb.os[d];
    for (int k = 0; k < nl; ++k) p.leaf[k].bs[d] = gb.ls[k][d];
  }
  p.rsz[0] = L;
  p.out_rs[0] = ostride;
  bool unit = nl > 0;
  for (int k = 0; k < nl; ++k) {
    p.leaf[k].rs[0] = e.leaves[k].stride[e.rank - 1];
    p.leaf[k].ptr = e.leaves[k].data;
    unit = unit && p.leaf[k].rs[0] == 1;
  }
  p.all_unit = unit ? 1 : 0;
  p.tx = ovec ? 1 : 0;
  p.out.ptr = out->data;
  fill_consts(e, p.c);

  int64_t rows_per_warp = 1;
  if (warp_team && spec.U == 1) {
    // short rows share a warp: G lanes per row, G = smallest power of two that covers the row's vectors
    const int64_t nv = (L + V - 1) / V;
    if (nv <= 32 && env_int("MXB_SCAN_NO_GROUP", 0) == 0) {
      int G = 1;
      while (G < nv) G <<= 1;
      p.scan_group = G;
      rows_per_warp = 32 / G;
    }
  }
  Kernel k;
  st = get_kernel(info, spec, &k);
  if (st != MXB_OK) return st;
  unsigned grid, scan_smem = 0;
  if (scan_wt) {
    // warp tiles of 32 lanes x 32 elements (16 for 8-byte values); slots per row: tile totals | group totals | running
    // totals at supergroup starts | supergroup totals
    const int epl4 = env_int("MXB_SCAN_WT_EPL", 32) == 16 ? 16 : 32;
    p.scan_plain = (e.n_nodes == 1 && nl == 1 && e.nodes[0].opcode == MXB_OP_LEAF && e.leaves[0].dtype == vt && unit && env_int("MXB_SCAN_WT_L2", 1)) ? 1 : 0;
    p.scan_depth = epl4;
    const int64_t wtile = 32 * std::max<int64_t>(V, dtype_bytes(vt) > 4 ? epl4 / 2 : epl4);
    const int64_t wtpr = (L + wtile - 1) / wtile, gpr = (wtpr + 31) / 32, spr = (wtpr + 1023) / 1024;
    const size_t need_t = (size_t)(B * wtpr), need_g = (size_t)(B * gpr), need_r = (size_t)(B * (spr + 1)), need_o = (size_t)(B * spr);
    st = ensure_lb(h, need_t + need_g + need_r + need_o);
    if (st != MXB_OK) return st;
    const int sb = dtype_bytes(vt) >= 8 ? 16 : 8;
    char *w = (char *)lb_region(h, sb);
    p.scan_agg = w;
    p.scan_gagg = w + need_t * sb;
    p.scan_sagg = w + (need_t + need_g) * sb;
    p.scan_own = w + (need_t + need_g + need_r) * sb;
    p.scan_ctl = h->lb_ctl;
    // the grid is one resident wave (cooperative launch): 3 CTAs per SM x 8 warps x 4 KB x 3 iterations = 43 MB of input lie
    // between a tile's two reads, which L2 holds for the most part when the first read asks for it (evict_last) and the
    // second read and the stores give way (evict_first): 1.45 GB instead of 2.1 GB read from DRAM for a 1 GiB row
    const int res = resident_ctas(k, 256, 0, 3);
    const int per_sm = std::min(env_int("MXB_SCAN_WT_CTAS", 3), res);
    grid = (unsigned)std::min<int64_t>((B * wtpr + 7) / 8, (int64_t)sm * per_sm);
  } else if (tiles_mode) {
    // slots: tile totals, group totals (32 tiles), running totals at supergroup starts (1024 tiles), per row
    const int64_t gpr = (tpr + 31) / 32, spr = (tpr + 1023) / 1024;
    const size_t need_t = (size_t)(B * tpr), need_g = (size_t)(B * gpr), need_s = (size_t)(B * spr);
    st = ensure_lb(h, need_t + need_g + need_s);
    if (st != MXB_OK) return st;
    const int sb = dtype_bytes(vt) >= 8 ? 16 : 8;
    char *w = (char *)lb_region(h, sb);
    p.scan_agg = w;
    p.scan_gagg = w + need_t * sb;
    p.scan_sagg = w + (need_t + need_g) * sb;
    p.scan_ctl = h->lb_ctl;
    // tiles are dealt round-robin to the grid and a tile waits for its predecessors, so every CTA must be resident:
    // the grid is sized from the launch-time occupancy and launched cooperatively (the runtime refuses a grid that
    // cannot be co-resident instead of letting it spin)
    int depth = env_int("MXB_TUNE_SCAN_DEPTH", 0) > 0 ? env_int("MXB_TUNE_SCAN_DEPTH", 0) : 4;
    while (depth > 2 && (int64_t)depth * tile * dtype_bytes(vt) > 72 * 1024) --depth;
    p.scan_depth = std::min(depth, 8);
    scan_smem = scan_pipe ? (unsigned)((int64_t)p.scan_depth * tile * dtype_bytes(vt)) : 0u;
    const int res = resident_ctas(k, 256, scan_smem, spec.minb > 0 ? spec.minb : 3);
    const int per_sm = env_int("MXB_SCAN_GRID_PER_SM", 0) > 0 ? std::min(env_int("MXB_SCAN_GRID_PER_SM", 0), res) : res;
    grid = (unsigned)std::min<int64_t>(B * tpr, (int64_t)sm * per_sm);
  } else {
    const int tune_cps = env_int("MXB_TUNE_CTAS_PER_SM", 0);
    const int64_t rows_per_cta = warp_team ? 8 * rows_per_warp : 1;
    // CTA-per-row walk: a whole number of resident waves (3 CTAs per SM at 80 registers; 8 per SM was 2.67 waves and left
    // the last one a third empty: 65536 x 4096 fp32 0.866 -> 0.915 of the copy peak, profiles/r2_scan_rows_grid.jsonl)
    const int cps = tune_cps > 0 ? tune_cps : (warp_team ? 8 : 4 * resident_ctas(k, 256, 0, 3));
    grid = (unsigned)std::min<int64_t>((B + rows_per_cta - 1) / rows_per_cta, (int64_t)sm * cps);
  }
  return launch(h, k, grid, 256u, scan_smem, p, /*coop=*/tiles_mode);
}

// ---------------------------------------------------------------------------------------------------
// find / find_idx: stream compaction in the flat (row-major) order of the expression
// ---------------------------------------------------------------------------------------------------
static int find_impl(mxb_handle_t h, const mxb_expr_t *expr_in, int select_op, double threshold, const mxb_out_t *out,
                     const mxb_out_t *count_out, int want_indices, bool unique) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  if (!expr_in) return fail(MXB_ERR_INVALID, "null expression");
  if (expr_in->rank >= 0 && expr_in->rank <= MXB_MAX_RANK && count_out && count_out->data && count_out->rank == 0 && count_out->dtype == MXB_I32) {
    int64_t n0 = 1;
    for (int d = 0; d < expr_in->rank; ++d) n0 *= expr_in->size[d];
    if (n0 == 0) {   // reference: num_found() = 0 and nothing else (cub.h:2660-2663); an empty tensor may have a null pointer
      MXB_CUDA(cudaSetDevice(h->device));
      MXB_CUDA(cudaMemsetAsync(count_out->data, 0, sizeof(int), h->stream));
      return MXB_OK;
    }
  }
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (unique) select_op = MXB_SEL_COUNT;   // internal: adjacent-difference flags over a sorted operand (mxb_unique)
  else if (select_op < 0 || select_op >= MXB_SEL_COUNT) return fail(MXB_ERR_INVALID, "unknown selection op");
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  if (out->rank != 1 || (out->size[0] > 1 && out->stride[0] != 1)) return fail(MXB_ERR_INVALID, "find output must be rank 1 and contiguous");
  if (!count_out || !count_out->data) return fail(MXB_ERR_INVALID, "null count output");
  if (count_out->rank != 0) return fail(MXB_ERR_INVALID, "num_found must be rank 0 (the reference static_asserts it)");
  if (count_out->dtype != MXB_I32) return fail(MXB_ERR_INVALID, "num_found must be MXB_I32 (tensor_t<int, 0>)");
  if (out->dtype < 0 || out->dtype >= MXB_DTYPE_COUNT) return fail(MXB_ERR_INVALID, "output dtype out of range");
  if (want_indices && out->dtype != MXB_I32 && out->dtype != MXB_I64) return fail(MXB_ERR_INVALID, "find_idx output must be MXB_I32 or MXB_I64");
  if (out->dtype == MXB_C64 || out->dtype == MXB_BF16 || out->dtype == MXB_F16) return fail(MXB_ERR_NOT_SUPPORTED, "find output dtype is not lowered");
  MXB_CUDA(cudaSetDevice(h->device));

  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  if (!(info.value_dtype == MXB_F32 || info.value_dtype == MXB_F64 || info.value_dtype == MXB_I32 || info.value_dtype == MXB_I64 || info.value_dtype == MXB_U8))
    return fail(MXB_ERR_NOT_SUPPORTED, "find / find_idx of this value type is not lowered (the reference orders real types only)");

  // collapse neighbours only: the flat order of the selection is the row-major order of the expression's dims
  Group g;
  g.n = e.rank;
  for (int d = 0; d < e.rank; ++d) {
    g.size[d] = e.size[d];
    for (int k = 0; k < e.n_leaves; ++k) g.ls[k][d] = e.leaves[k].stride[d];
    g.os[d] = 0;
    g.is[d] = 0;
  }
  int64_t N = 1;
  for (int d = 0; d < e.rank; ++d) N *= e.size[d];
  // size-1 dims carry no order information; broadcast (stride-0) dims do, so the generic collapse rule applies as is
  collapse(g, e.n_leaves);
  if (g.n > KMAXD) return fail(MXB_ERR_NOT_SUPPORTED, "view does not collapse to <= 4 dims");

  const int nl = e.n_leaves;
  int V = 1;
  bool unit = nl > 0;
  if (g.n == 1) {
    const int vmax = policy_vmax(info);
    bool ok = vmax > 1;
    for (int k = 0; k < nl; ++k) {
      const int64_t in = g.ls[k][0];
      if (in != 1) unit = false;
      if (in != 0 && in != 1) ok = false;
      if (in == 1 && !aligned_to(e.leaves[k].data, (int64_t)vmax * dtype_bytes(e.leaves[k].dtype))) ok = false;
    }
    if (ok && unit) V = vmax;
  } else {
    unit = false;
  }
  if (env_int("MXB_TUNE_V", 0) == 1) V = 1;

  const int64_t TILE = (int64_t)256 * V * 4;
  const int64_t ntiles = (N + TILE - 1) / TILE;

  EwParams p;
  memset(&p, 0, sizeof p);
  p.nd = g.n;
  p.N = N;
  p.nleaf = nl;
  p.all_unit = (unit && V > 1) ? 1 : 0;
  for (int d = 0; d < g.n; ++d) {
    p.sz[d] = g.size[d];
    for (int k = 0; k < nl; ++k) p.leaf[k].bs[d] = g.ls[k][d];
  }
  for (int k = 0; k < nl; ++k) p.leaf[k].ptr = e.leaves[k].data;
  p.out.ptr = out->data;
  fill_consts(e, p.c);
  p.sel_op = select_op;
  p.sel_thr_d = threshold;
  p.sel_thr_i = (int64_t)threshold;
  p.sel_total = (int *)count_out->data;
  p.sel_cap = out->size[0];

  KernelSpec spec;
  spec.family = FAM_SELECT;
  spec.op = -1;
  spec.V = V;
  spec.U = 4;
  Kernel k;

  // ---- one pass (select1p): a view that collapses to one dim, counts that fit 32 bits --------------------------------
  if (unique && !(g.n == 1 && N < (1ll << 32) - TILE)) return fail(MXB_ERR_NOT_SUPPORTED, "unique of more than 2^32 elements");
  if (g.n == 1 && N < (1ll << 32) - TILE && (unique || env_int("MXB_SEL_TWO_PASS", 0) == 0)) {
    if (V == 1) p.all_unit = 0;
    spec.team = want_indices ? 4 : 3;
    spec.out_dtype = out->dtype;
    st = get_kernel(info, spec, &k);
    if (st != MXB_OK) return st;
    // warp tiles of 32 lanes x 64 elements (32 for 8-byte values, 8 on the scalar walk), dealt round-robin to the warps of a grid that is
    // resident as a whole (cooperative launch: a tile only waits for tiles whose warps are running)
    const int64_t wtile = 32 * (V == 1 ? 8 : (dtype_bytes(info.value_dtype) > 4 ? 32 : 64));   // SelGeom::TILE (mxb_device.cuh)
    const int64_t nwt = (N + wtile - 1) / wtile;
    const unsigned smem = 0;
    const int res = resident_ctas(k, 256, smem, 3);
    const int cps = env_int("MXB_TUNE_SEL_CTAS", 0) > 0 ? std::min(env_int("MXB_TUNE_SEL_CTAS", 0), res) : res;
    const unsigned grid = (unsigned)std::min<int64_t>((nwt + 7) / 8, (int64_t)h->sm_count * cps);
    // status words: tile counts | group counts | running counts at supergroup starts | supergroup counts
    st = ensure_lb(h, (size_t)(nwt + (nwt + 31) / 32 + 2 * ((nwt + 1023) / 1024) + 4));
    if (st != MXB_OK) return st;
    p.sel_status = (unsigned long long *)lb_region(h, 8);
    p.sel_epoch = h->lb_ctl;
    p.sel_ticket = h->lb_ctl + 1;
    // low half: back-off between two polls of a slot that is not published yet (ns); high half: tiles with at most this
    // many selected elements walk their set bits, fuller ones store every element slot under a predicate
    // (2048-element tiles, profiles/r2_find_warp_tiles_sweep.jsonl: indices walk up to a quarter of the tile, values — whose
    // walk re-reads one element per set bit — up to 5 %)
    p.sel_depth = (env_int("MXB_TUNE_SEL_NAP", 200) & 0xffff) | (env_int("MXB_TUNE_SEL_SPARSE", want_indices ? 512 : 96) << 16);
    return launch(h, k, grid, 256, smem, p, /*coop=*/true);
  }

  // ---- two passes: N-D views that do not collapse (row-major decomposition per element), counts beyond 32 bits -------
  const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)h->sm_count * (env_int("MXB_TUNE_SEL_CTAS", 0) > 0 ? env_int("MXB_TUNE_SEL_CTAS", 0) : 8));
  st = ensure_ws(h, (size_t)grid * 16 + 64, 1);
  if (st != MXB_OK) return st;
  p.sel_offsets = (unsigned long long *)h->ws;
  p.sel_counts = (unsigned long long *)h->ws + grid;
  p.sel_ticket = h->tickets;
  spec.team = 0;
  spec.out_dtype = info.value_dtype;
  st = get_kernel(info, spec, &k);
  if (st != MXB_OK) return st;
  st = launch(h, k, grid, 256, 0, p);
  if (st != MXB_OK) return st;
  spec.team = want_indices ? 2 : 1;
  spec.out_dtype = out->dtype;
  st = get_kernel(info, spec, &k);
  if (st != MXB_OK) return st;
  return launch(h, k, grid, 256, 0, p);
}

int mxb_find(mxb_handle_t h, const mxb_expr_t *expr_in, int select_op, double threshold, const mxb_out_t *out,
             const mxb_out_t *count_out, int want_indices) {
  return find_impl(h, expr_in, select_op, threshold, out, count_out, want_indices, false);
}

// ---------------------------------------------------------------------------------------------------
// hist: even-width histogram of every row (reference: hist_impl, transforms/cub.h:2464-2503)
// ---------------------------------------------------------------------------------------------------
int mxb_hist(mxb_handle_t h, const mxb_expr_t *expr_in, double lower, double upper, const mxb_out_t *out) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (expr_in->rank < 1) return fail(MXB_ERR_INVALID, "hist needs rank >= 1");
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  if (out->dtype != MXB_I32) return fail(MXB_ERR_INVALID, "histogram output must be MXB_I32 (the reference static_asserts int)");
  if (out->rank != expr_in->rank) return fail(MXB_ERR_SIZE, "histogram output rank must equal the input rank");
  for (int d = 0; d + 1 < out->rank; ++d)
    if (out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
  const int64_t bins = out->size[out->rank - 1];
  if (bins < 1 || bins > (1 << 24)) return fail(MXB_ERR_INVALID, "bin count out of range");
  if (!(upper > lower)) return fail(MXB_ERR_INVALID, "hist needs lower < upper");
  if (bins > 1 && out->stride[out->rank - 1] != 1) return fail(MXB_ERR_NOT_SUPPORTED, "the bins of a row must be contiguous");
  MXB_CUDA(cudaSetDevice(h->device));
  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  const int vt = info.value_dtype;
  if (!(vt == MXB_F32 || vt == MXB_F64 || vt == MXB_I32 || vt == MXB_I64 || vt == MXB_U8))
    return fail(MXB_ERR_NOT_SUPPORTED, "hist of this value type is not lowered");

  const int nbd = e.rank - 1, nl = e.n_leaves;
  Group gb;
  gb.n = nbd;
  for (int d = 0; d < nbd; ++d) {
    gb.size[d] = e.size[d];
    for (int k = 0; k < nl; ++k) gb.ls[k][d] = e.leaves[k].stride[d];
    gb.os[d] = out->stride[d];
    gb.is[d] = 0;
  }
  collapse(gb, nl);
  if (gb.n > KMAXD) return fail(MXB_ERR_NOT_SUPPORTED, "batch dims do not collapse to <= 4");
  const int64_t L = e.size[e.rank - 1];
  int64_t B = 1;
  for (int d = 0; d < gb.n; ++d) B *= gb.size[d];
  if (B == 0) return MXB_OK;
  // the output is an accumulator: zero every row's bins on the stream first (rows may be strided)
  {
    bool dense = true;
    int64_t w = bins;
    for (int d = nbd - 1; d >= 0; --d) { if (out->size[d] > 1 && out->stride[d] != w) dense = false; w *= out->size[d]; }
    if (dense) MXB_CUDA(cudaMemsetAsync(out->data, 0, (size_t)(B * bins) * 4, h->stream));
    else if (gb.n <= 1) MXB_CUDA(cudaMemset2DAsync(out->data, (size_t)(gb.n ? gb.os[0] : bins) * 4, 0, (size_t)bins * 4, (size_t)B, h->stream));
    else return fail(MXB_ERR_NOT_SUPPORTED, "histogram rows must be dense or singly strided");
  }
  if (L == 0) return MXB_OK;

  int V = policy_vmax(info);
  auto in_ok = [&](int v) {
    for (int k = 0; k < nl; ++k) {
      const int64_t in = e.leaves[k].stride[e.rank - 1];
      if (in != 0 && in != 1) return false;
      if (in == 0) continue;
      if (!aligned_to(e.leaves[k].data, (int64_t)v * dtype_bytes(e.leaves[k].dtype))) return false;
      for (int d = 0; d < gb.n; ++d) if (gb.ls[k][d] % v) return false;
    }
    return true;
  };
  if (V > 1 && !in_ok(V)) V = 1;
  KernelSpec spec;
  spec.family = FAM_HIST;
  spec.op = -1;
  spec.out_dtype = MXB_I32;
  spec.V = V;
  spec.U = 4;
  Kernel k;
  st = get_kernel(info, spec, &k);
  if (st != MXB_OK) return st;

  RedParams p;
  memset(&p, 0, sizeof p);
  p.nb = gb.n;
  p.nr = 1;
  p.B = B;
  p.R = L;
  p.nleaf = nl;
  for (int d = 0; d < gb.n; ++d) {
    p.bsz[d] = gb.size[d];
    p.out.bs[d] = gb.os[d];
    for (int kk = 0; kk < nl; ++kk) p.leaf[kk].bs[d] = gb.ls[kk][d];
  }
  p.rsz[0] = L;
  bool unit = nl > 0;
  for (int kk = 0; kk < nl; ++kk) {
    p.leaf[kk].rs[0] = e.leaves[kk].stride[e.rank - 1];
    p.leaf[kk].ptr = e.leaves[kk].data;
    unit = unit && p.leaf[kk].rs[0] == 1;
  }
  p.all_unit = unit ? 1 : 0;
  p.out.ptr = out->data;
  fill_consts(e, p.c);
  p.hist_lo_d = lower;
  p.hist_hi_d = upper;
  p.hist_lo_i = (int64_t)lower;
  p.hist_hi_i = (int64_t)upper;
  p.hist_bins = (int)bins;
  const int64_t smem_need = bins * 4;
  p.hist_smem = smem_need <= 128 * 1024 ? 1 : 0;
  const unsigned smem = p.hist_smem ? (unsigned)smem_need : 0u;
  // chunks per row: enough items to fill the machine, at least 64 K elements each (the flush costs `bins` atomics)
  const int64_t grid_max = (int64_t)h->sm_count * 8;
  int64_t S = 1;
  if (B < grid_max) S = std::max<int64_t>(1, std::min<int64_t>((grid_max + B - 1) / B, (L + 65535) / 65536));
  p.splits = (int)S;
  const unsigned grid = (unsigned)std::min<int64_t>(B * S, grid_max);
  return launch(h, k, grid, 256u, smem, p);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// sort: every row of the innermost dim, keys only (reference: sort_impl, transforms/cub.h:2145-2190)
// ---------------------------------------------------------------------------------------------------
namespace {
int ensure_buf(mxb_context *h, void **buf, size_t *have, size_t bytes) {
  if (bytes > *have) {
    if (*buf) MXB_CUDA(cudaFreeAsync(*buf, h->stream));
    const size_t nb = std::max<size_t>(bytes, 64 * 1024);
    MXB_CUDA(cudaMallocAsync(buf, nb, h->stream));
    *have = nb;
  }
  return MXB_OK;
}

template <class K>
int sort_rows_typed(mxb_context *h, const void *in, void *out, int kind, int64_t B, int64_t L, bool desc) {
  mxbsort::SortParams p;
  memset(&p, 0, sizeof p);
  p.B = B;
  p.L = L;
  p.kind = kind;
  p.desc = desc ? 1 : 0;
  const int sm = h->sm_count;
  if (L <= 4096) {
    int n2 = 1;
    while (n2 < L) n2 <<= 1;
    if (n2 < 2) n2 = 2;
    p.in = in;
    p.out = out;
    const unsigned grid = (unsigned)std::min<int64_t>(B, (int64_t)sm * 8);
    if (!g_plan) mxbsort::bitonic_rows_kernel<K><<<grid, 256, (size_t)n2 * sizeof(K), h->stream>>>(p, n2);
    MXB_CUDA(cudaGetLastError());
    h->launches++;
    h->last_kernel = std::string("sort_bitonic|") + (sizeof(K) == 4 ? "k32" : "k64");
    return MXB_OK;
  }
  if (L >= (1ll << 32)) return fail(MXB_ERR_NOT_SUPPORTED, "rows of 2^32 or more keys");
  const int passes = (int)sizeof(K);
  // chunks of whole 2048-key scatter rounds: about one chunk per resident CTA for long rows, never more than 4096 per row
  int64_t want = std::max<int64_t>(2048, (B * L + (int64_t)sm * 8 - 1) / ((int64_t)sm * 8));
  want = std::max<int64_t>(want, (L + 4095) / 4096);
  const int64_t chunk = (want + 2047) / 2048 * 2048;
  int64_t cpr = (L + chunk - 1) / chunk;
  p.cpr = (int)cpr;
  p.chunk = chunk;
  int st = ensure_buf(h, &h->tmp, &h->tmp_bytes, (size_t)(B * L) * sizeof(K));
  if (st != MXB_OK) return st;
  const size_t words = (size_t)(B * cpr) * 256, tot_words = (size_t)B * 256;
  if (2 * words + tot_words > h->sort_ctr_words) {
    if (h->sort_ctr) MXB_CUDA(cudaFreeAsync(h->sort_ctr, h->stream));
    MXB_CUDA(cudaMallocAsync((void **)&h->sort_ctr, (2 * words + tot_words) * 4, h->stream));
    h->sort_ctr_words = 2 * words + tot_words;
  }
  p.counts = h->sort_ctr;
  p.offsets = h->sort_ctr + words;
  p.totals = h->sort_ctr + 2 * words;
  const unsigned grid = (unsigned)std::min<int64_t>(B * cpr, (int64_t)sm * 8);
  const void *src = in;
  for (int pass = 0; pass < passes; ++pass) {
    // ping-pong: the even number of passes ends in `out`
    void *dst = (pass & 1) ? out : h->tmp;
    p.in = src;
    p.out = dst;
    p.shift = 8 * pass;
    p.first = pass == 0;
    p.last = pass == passes - 1;
    if (!g_plan) {
      mxbsort::radix_count_kernel<K><<<grid, 256, 0, h->stream>>>(p);
      mxbsort::radix_scan_kernel<<<(unsigned)std::min<int64_t>(B * 32, (int64_t)sm * 8), 256, 0, h->stream>>>(p);   // a warp per (row, digit)
      mxbsort::radix_scatter_kernel<K><<<grid, 256, 0, h->stream>>>(p);
    }
    MXB_CUDA(cudaGetLastError());
    h->launches += 3;
    src = dst;
  }
  h->last_kernel = std::string("sort_radix|") + (sizeof(K) == 4 ? "k32" : "k64") + "|passes=" + std::to_string(passes);
  return MXB_OK;
}

int sort_rows(mxb_context *h, const void *in, void *out, int dtype, int64_t B, int64_t L, bool desc) {
  if (B == 0 || L == 0) return MXB_OK;
  switch (dtype) {
    case MXB_F32: return sort_rows_typed<uint32_t>(h, in, out, 2, B, L, desc);
    case MXB_I32: return sort_rows_typed<uint32_t>(h, in, out, 1, B, L, desc);
    case MXB_F64: return sort_rows_typed<unsigned long long>(h, in, out, 2, B, L, desc);
    case MXB_I64: return sort_rows_typed<unsigned long long>(h, in, out, 1, B, L, desc);
    default: return fail(MXB_ERR_NOT_SUPPORTED, "sort of this key type is not lowered (fp32, fp64, int32, int64)");
  }
}

// a program that is ONE leaf walking `dims` row-major-contiguously with the output's dtype: sortable in place of a copy
const void *plain_contiguous_leaf(const mxb_expr_t &e, int out_dtype) {
  if (e.n_nodes != 1 || e.nodes[0].opcode != MXB_OP_LEAF || e.n_leaves != 1) return nullptr;
  const mxb_leaf_t &lf = e.leaves[0];
  if (lf.dtype != out_dtype) return nullptr;
  int64_t w = 1;
  for (int d = e.rank - 1; d >= 0; --d) {
    if (e.size[d] > 1 && lf.stride[d] != w) return nullptr;
    w *= e.size[d];
  }
  return lf.data;
}
}  // namespace

extern "C" {

int mxb_sort(mxb_handle_t h, const mxb_expr_t *expr_in, const mxb_out_t *out, int descending) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (expr_in->rank < 1) return fail(MXB_ERR_INVALID, "sort needs rank >= 1");
  if (!out || !out->data) return fail(MXB_ERR_INVALID, "null output");
  if (out->rank != expr_in->rank) return fail(MXB_ERR_SIZE, "sort output rank must equal the input rank");
  int64_t w = 1, N = 1;
  for (int d = out->rank - 1; d >= 0; --d) {
    if (out->size[d] != expr_in->size[d]) return fail(MXB_ERR_SIZE, "output size mismatch in dim " + std::to_string(d));
    if (out->size[d] > 1 && out->stride[d] != w) return fail(MXB_ERR_NOT_SUPPORTED, "sort output must be contiguous (the reference requires it too)");
    w *= out->size[d];
    N *= out->size[d];
  }
  MXB_CUDA(cudaSetDevice(h->device));
  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  if (info.value_dtype != out->dtype) return fail(MXB_ERR_INVALID, "sort output dtype must equal the operand's value type");
  if (N == 0) return MXB_OK;
  const int64_t L = e.size[e.rank - 1], B = N / L;
  const void *src = plain_contiguous_leaf(e, out->dtype);
  if (!src) {
    // an expression / a strided view: evaluate it into the output (contiguous) and sort there, like the reference's copy
    st = mxb_elementwise(h, expr_in, out);
    if (st != MXB_OK) return st;
    src = out->data;
  }
  return sort_rows(h, src, out->data, out->dtype, B, L, descending != 0);
}

// ---------------------------------------------------------------------------------------------------
// unique: sorted distinct values of a rank-1 operand + their count (reference: unique_impl, transforms/cub.h:2796-2842)
// ---------------------------------------------------------------------------------------------------
int mxb_unique(mxb_handle_t h, const mxb_expr_t *expr_in, const mxb_out_t *out, const mxb_out_t *count_out) {
  if (!h) return fail(MXB_ERR_INVALID, "null handle");
  std::lock_guard<std::recursive_mutex> lock_(h->mu);
  int st = check_expr_shape(expr_in);
  if (st != MXB_OK) return st;
  if (expr_in->rank != 1) return fail(MXB_ERR_NOT_SUPPORTED, "unique serves rank-1 operands");
  if (!out || !out->data || out->rank != 1 || (out->size[0] > 1 && out->stride[0] != 1)) return fail(MXB_ERR_INVALID, "unique output must be rank 1 and contiguous");
  if (!count_out || !count_out->data || count_out->rank != 0 || count_out->dtype != MXB_I32) return fail(MXB_ERR_INVALID, "num_found must be a rank-0 MXB_I32");
  MXB_CUDA(cudaSetDevice(h->device));
  const int64_t N = expr_in->size[0];
  if (N == 0) {
    MXB_CUDA(cudaMemsetAsync(count_out->data, 0, sizeof(int), h->stream));
    return MXB_OK;
  }
  mxb_expr_t e;
  std::string err;
  st = canonicalize(expr_in, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  const int vt = info.value_dtype;
  if (vt != out->dtype) return fail(MXB_ERR_INVALID, "unique output dtype must equal the operand's value type");
  if (!(vt == MXB_F32 || vt == MXB_F64 || vt == MXB_I32 || vt == MXB_I64)) return fail(MXB_ERR_NOT_SUPPORTED, "unique of this value type is not lowered");
  st = ensure_buf(h, &h->tmp2, &h->tmp2_bytes, (size_t)N * dtype_bytes(vt));
  if (st != MXB_OK) return st;
  mxb_out_t sorted;
  memset(&sorted, 0, sizeof sorted);
  sorted.data = h->tmp2;
  sorted.dtype = vt;
  sorted.rank = 1;
  sorted.size[0] = N;
  sorted.stride[0] = 1;
  st = mxb_sort(h, expr_in, &sorted, 0);
  if (st != MXB_OK) return st;
  mxb_expr_t se;
  memset(&se, 0, sizeof se);
  se.rank = 1;
  se.size[0] = N;
  se.n_nodes = 1;
  se.n_leaves = 1;
  se.nodes[0] = mxb_node_t{MXB_OP_LEAF, {0, -1}, 0};
  se.leaves[0].data = h->tmp2;
  se.leaves[0].dtype = vt;
  se.leaves[0].stride[0] = 1;
  return find_impl(h, &se, 0, 0.0, out, count_out, 0, /*unique=*/true);
}

int mxb_is_aot(const mxb_expr_t *expr, int reduce_op_or_minus1) {
  if (!expr) return 0;
  mxb_expr_t e;
  std::string err;
  if (canonicalize(expr, &e, &err) != MXB_OK) return 0;
  ExprInfo info;
  if (analyze_expr(&e, &info, &err) != MXB_OK) return 0;
  KernelSpec spec;
  if (reduce_op_or_minus1 < 0) {
    spec.family = FAM_EW;
    spec.op = -1;
    spec.out_dtype = info.value_dtype;
  } else {
    spec.family = (reduce_op_or_minus1 == MXB_RED_VAR || reduce_op_or_minus1 == MXB_RED_STDD) ? FAM_VAR_SMEM : FAM_RED_INNER;
    spec.op = kernel_op(reduce_op_or_minus1);
    spec.out_dtype = info.value_dtype;
  }
  spec.V = policy_vmax(info);
  spec.U = policy_unroll(info, spec.V, spec.family);
  spec.team = (reduce_op_or_minus1 < 0 && e.n_nodes >= 16) ? 4 : 0;   // the dispatcher's occupancy hint for heavy elementwise programs
  return lookup_aot(kernel_key(info, spec)) ? 1 : 0;
}

// ---- test hooks (not part of the reference-facing surface) -----------------------------------------
// Canonical signature + generated functor of a program, for the CPU-side tests of the lowering.
int mxb_debug_codegen(const mxb_expr_t *expr, char *buf, size_t buflen) {
  if (!expr || !buf || buflen == 0) return fail(MXB_ERR_INVALID, "bad arguments");
  mxb_expr_t e;
  std::string err;
  int st = canonicalize(expr, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  const std::string s = std::string("value_dtype=") + dtype_name(info.value_dtype) + "\n" + info.src;
  snprintf(buf, buflen, "%s", s.c_str());
  return MXB_OK;
}

// Build (NVRTC only, no device needed) the kernel a call would use; proves on a CPU box that the
// generated source compiles for sm_100a.  family: 0 red_inner, 1 red_outer, 2 var_smem, 3 elementwise, 9 transposing elementwise.
int mxb_debug_compile(const mxb_expr_t *expr, int family, int reduce_op, int out_dtype, int V, int team, char *log, size_t loglen) {
  if (!expr) return fail(MXB_ERR_INVALID, "null expression");
  mxb_expr_t e;
  std::string err;
  int st = canonicalize(expr, &e, &err);
  if (st != MXB_OK) return fail(st, err);
  ExprInfo info;
  st = analyze_expr(&e, &info, &err);
  if (st != MXB_OK) return fail(st, err);
  KernelSpec spec;
  spec.family = family;
  spec.op = (family == FAM_EW || family == FAM_EW_TR || family == FAM_SCAN || family == FAM_SM_GROUP || family == FAM_SM_REG || family == FAM_SELECT) ? -1 : kernel_op(reduce_op);
  spec.out_dtype = out_dtype;
  spec.V = V > 0 ? V : policy_vmax(info);
  spec.U = policy_unroll(info, spec.V, family);
  spec.team = team;
  const std::string key = kernel_key(info, spec);
  std::string wrap;
  st = kernel_wrapper_src(info, spec, kernel_symbol(key), &wrap, &err);
  if (st != MXB_OK) return fail(st, err);
  const std::string src = "#include \"mxb_device.cuh\"\n" + info.src + "\n" + wrap;
  std::string lg;
  st = jit_compile_only(src, &lg);
  if (log && loglen) snprintf(log, loglen, "%s", lg.c_str());
  if (st != MXB_OK) return fail(st, "NVRTC: " + lg);
  return MXB_OK;
}

}  // extern "C"
