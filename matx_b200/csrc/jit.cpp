// jit.cpp — kernel registry: the ahead-of-time table plus an NVRTC-backed cache for every expression
// program that is not in it.  The reference resolves its kernels through C++ templates in the user's
// translation unit (executors/cuda_executor_common.h:190-312 builds a table of kernel pointers per
// operator type); behind a C ABI the operator type is gone, so the program is re-materialised as a
// functor (codegen.cpp), dropped into the same skeleton header the AOT kernels use, and compiled for
// sm_100a once per (program, kernel family, reduce op, dtypes).  NVRTC and the driver are dlopen'ed on
// first use so that the library itself loads on a machine with neither.
#include <cuda.h>   // types only (CUlaunchConfig, CUfunction ...): every driver entry point is resolved with dlsym
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

#include "mxb_internal.h"

namespace mxbh {

// ---- AOT table ------------------------------------------------------------------------------------
namespace {
std::map<std::string, const void *> &aot_table() {
  static std::map<std::string, const void *> t;
  return t;
}
std::mutex g_mu;
}  // namespace

void register_aot(const AotEntry *entries, int n) {
  for (int i = 0; i < n; ++i) aot_table()[entries[i].key] = entries[i].fn;
}
const void *lookup_aot(const std::string &key) {
  auto it = aot_table().find(key);
  return it == aot_table().end() ? nullptr : it->second;
}

// ---- NVRTC + driver entry points, resolved lazily ---------------------------------------------------
namespace {
typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;

struct Dyn {
  bool tried = false, ok = false;
  std::string why;
  nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
  const char *(*GetErrorString)(nvrtcResult) = nullptr;
  CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
  CUresult (*GetErrorStringDrv)(CUresult, const char **) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig *, CUfunction, void **, void **) = nullptr;   // optional (PDL launches)
  CUresult (*CtxGetCurrent)(CUcontext *) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*OccupancyMaxActiveBlocks)(int *, CUfunction, int, size_t) = nullptr;
  std::string include_dir;
};
Dyn g_dyn;

void *open_first(const std::vector<std::string> &names) {
  for (const std::string &n : names) {
    void *h = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (h) return h;
  }
  return nullptr;
}

bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

bool dyn_init(std::string *err) {
  if (g_dyn.tried) { if (!g_dyn.ok && err) *err = g_dyn.why; return g_dyn.ok; }
  g_dyn.tried = true;
  std::vector<std::string> roots;
  for (const char *ev : {"MXB_CUDA_HOME", "CUDA_HOME", "CUDA_PATH"}) if (const char *v = getenv(ev)) roots.push_back(v);
  roots.push_back("/usr/local/cuda");
  // The toolkit's own NVRTC first, by absolute path: a bare "libnvrtc.so.12" resolves to whatever copy the process
  // already holds (PyTorch bundles NVRTC 12.8, whose ptxas rejects the 256-bit loads of PTX ISA 8.8).
  std::vector<std::string> nv;
  for (const std::string &r : roots) { nv.push_back(r + "/lib64/libnvrtc.so.12"); nv.push_back(r + "/lib64/libnvrtc.so"); }
  nv.push_back("libnvrtc.so.12");
  nv.push_back("libnvrtc.so");
  void *hn = open_first(nv);
  if (!hn) { g_dyn.why = "libnvrtc.so.12 not found (set MXB_CUDA_HOME)"; if (err) *err = g_dyn.why; return false; }
  void *hc = open_first({"libcuda.so.1", "libcuda.so"});
  if (!hc) { g_dyn.why = "libcuda.so.1 not found"; if (err) *err = g_dyn.why; return false; }
#define MXB_SYM(h, field, name) *(void **)(&g_dyn.field) = dlsym(h, name); if (!g_dyn.field) { g_dyn.why = std::string("missing symbol ") + name; if (err) *err = g_dyn.why; return false; }
  MXB_SYM(hn, CreateProgram, "nvrtcCreateProgram")
  MXB_SYM(hn, CompileProgram, "nvrtcCompileProgram")
  MXB_SYM(hn, GetCUBINSize, "nvrtcGetCUBINSize")
  MXB_SYM(hn, GetCUBIN, "nvrtcGetCUBIN")
  MXB_SYM(hn, GetProgramLogSize, "nvrtcGetProgramLogSize")
  MXB_SYM(hn, GetProgramLog, "nvrtcGetProgramLog")
  MXB_SYM(hn, DestroyProgram, "nvrtcDestroyProgram")
  MXB_SYM(hn, GetErrorString, "nvrtcGetErrorString")
  MXB_SYM(hc, ModuleLoadData, "cuModuleLoadData")
  MXB_SYM(hc, ModuleGetFunction, "cuModuleGetFunction")
  MXB_SYM(hc, GetErrorStringDrv, "cuGetErrorString")
  MXB_SYM(hc, LaunchKernel, "cuLaunchKernel")
  MXB_SYM(hc, FuncSetAttribute, "cuFuncSetAttribute")
  MXB_SYM(hc, CtxGetCurrent, "cuCtxGetCurrent")
  MXB_SYM(hc, ModuleUnload, "cuModuleUnload")
  MXB_SYM(hc, OccupancyMaxActiveBlocks, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
#undef MXB_SYM
  *(void **)(&g_dyn.LaunchKernelEx) = dlsym(hc, "cuLaunchKernelEx");
  // cuda_fp16.h / cuda_bf16.h live next to the toolkit
  for (const std::string &r : roots)
    if (file_exists(r + "/include/cuda_bf16.h")) { g_dyn.include_dir = r + "/include"; break; }
  if (g_dyn.include_dir.empty()) {
    Dl_info di;
    if (dladdr((void *)g_dyn.CreateProgram, &di) && di.dli_fname) {
      std::string p = di.dli_fname;  // .../lib64/libnvrtc.so.12
      for (int up = 0; up < 2; ++up) { const size_t s = p.find_last_of('/'); if (s == std::string::npos) break; p = p.substr(0, s); }
      for (const char *sub : {"/include", "/targets/x86_64-linux/include"})
        if (file_exists(p + sub + "/cuda_bf16.h")) { g_dyn.include_dir = p + sub; break; }
    }
  }
  if (g_dyn.include_dir.empty()) { g_dyn.why = "CUDA include directory (cuda_bf16.h) not found; set MXB_CUDA_HOME"; if (err) *err = g_dyn.why; return false; }
  g_dyn.ok = true;
  return true;
}

std::string cache_dir() {
  if (const char *v = getenv("MXB_CACHE_DIR")) return v;
  const char *home = getenv("HOME");
  return std::string(home ? home : "/tmp") + "/.cache/matx_b200";
}
void mkdirs(const std::string &p) {
  std::string cur;
  for (size_t i = 0; i < p.size(); ++i) {
    cur.push_back(p[i]);
    if (p[i] == '/' || i + 1 == p.size()) mkdir(cur.c_str(), 0755);
  }
}

std::map<std::string, const void *> g_jit;  // key -> CUfunction
int64_t g_jit_compiles = 0;
}  // namespace

const void *jit_get_kernel(const std::string &key, const std::string &symbol, const std::string &source, std::string *err) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!dyn_init(err)) return nullptr;
  // a CUfunction belongs to the context its module was loaded in: one entry per (context, kernel), so a second handle on
  // another GPU of the same process gets a function of ITS device
  CUcontext ctx = nullptr;
  g_dyn.CtxGetCurrent(&ctx);
  char ctxs[32];
  snprintf(ctxs, sizeof ctxs, "@%p|", (void *)ctx);
  const std::string ckey = ctxs + key;
  auto it = g_jit.find(ckey);
  if (it != g_jit.end()) return it->second;

  // disk cache keyed by the full source text (skeleton + functor + wrapper)
  char hname[64];
  snprintf(hname, sizeof hname, "%016llx_%016llx.cubin", (unsigned long long)fnv64(source + (getenv("MXB_LD_FLAVOR") ? getenv("MXB_LD_FLAVOR") : "")),
           (unsigned long long)fnv64(std::string(kDeviceHeaderText)));
  const std::string cpath = cache_dir() + "/" + hname;
  std::string cubin;
  bool from_disk = false;
  if (!getenv("MXB_NO_DISK_CACHE")) {
    std::ifstream f(cpath, std::ios::binary);
    if (f) { std::stringstream ss; ss << f.rdbuf(); cubin = ss.str(); from_disk = !cubin.empty(); }
  }
  CUmodule mod = nullptr;
  if (from_disk && g_dyn.ModuleLoadData(&mod, cubin.data()) != 0) {
    // truncated / corrupt cache entry: drop it and build again, once
    mod = nullptr;
    cubin.clear();
    unlink(cpath.c_str());
  }
  if (cubin.empty()) {
    nvrtcProgram prog = nullptr;
    const char *hdr_src[] = {kDeviceHeaderText};
    const char *hdr_name[] = {"mxb_device.cuh"};
    nvrtcResult r = g_dyn.CreateProgram(&prog, source.c_str(), "mxb_jit.cu", 1, hdr_src, hdr_name);
    if (r != 0) { if (err) *err = std::string("nvrtcCreateProgram: ") + g_dyn.GetErrorString(r); return nullptr; }
    const std::string inc = "--include-path=" + g_dyn.include_dir;
    std::string flavor = "-DMXB_LD_FLAVOR=0";
    if (const char *fv = getenv("MXB_LD_FLAVOR")) flavor = std::string("-DMXB_LD_FLAVOR=") + fv;
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", inc.c_str(), "-lineinfo", flavor.c_str()};
    r = g_dyn.CompileProgram(prog, 5, opts);
    if (r != 0) {
      size_t n = 0;
      g_dyn.GetProgramLogSize(prog, &n);
      std::string log(n, '\0');
      if (n) g_dyn.GetProgramLog(prog, &log[0]);
      if (err) *err = std::string("nvrtcCompileProgram: ") + g_dyn.GetErrorString(r) + "\n" + log;
      g_dyn.DestroyProgram(&prog);
      return nullptr;
    }
    size_t n = 0;
    g_dyn.GetCUBINSize(prog, &n);
    cubin.resize(n);
    g_dyn.GetCUBIN(prog, &cubin[0]);
    g_dyn.DestroyProgram(&prog);
    ++g_jit_compiles;
    if (!getenv("MXB_NO_DISK_CACHE")) {
      mkdirs(cache_dir());
      const std::string tmp = cpath + "." + std::to_string((long)getpid());
      std::ofstream f(tmp, std::ios::binary);
      if (f) { f.write(cubin.data(), (std::streamsize)cubin.size()); f.close(); rename(tmp.c_str(), cpath.c_str()); }
    }
  }
  CUresult cr = mod ? CUDA_SUCCESS : g_dyn.ModuleLoadData(&mod, cubin.data());
  if (cr != 0) {
    const char *s = nullptr;
    g_dyn.GetErrorStringDrv(cr, &s);
    if (err) *err = std::string("cuModuleLoadData: ") + (s ? s : "?");
    return nullptr;
  }
  CUfunction fn = nullptr;
  cr = g_dyn.ModuleGetFunction(&fn, mod, symbol.c_str());
  if (cr != 0) { if (err) *err = "cuModuleGetFunction failed for " + symbol; return nullptr; }
  g_jit[ckey] = (const void *)fn;
  return (const void *)fn;
}

int jit_occupancy(const void *fn, unsigned block, unsigned smem) {
  if (!g_dyn.ok || !g_dyn.OccupancyMaxActiveBlocks) return 0;
  int n = 0;
  if (g_dyn.OccupancyMaxActiveBlocks(&n, (CUfunction)fn, (int)block, (size_t)smem) != 0) return 0;
  return n;
}


int jit_launch(const void *fn, unsigned grid, unsigned block, unsigned smem, void *stream, void *params, std::string *err, bool pdl, bool coop) {
  if (!g_dyn.ok) { if (err) *err = "JIT runtime not initialised"; return MXB_ERR_JIT; }
  CUfunction f = (CUfunction)fn;
  if (smem > 40 * 1024) {
    const CUresult a = g_dyn.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
    if (a != 0) { if (err) *err = "cuFuncSetAttribute(max dynamic smem) failed"; return MXB_ERR_CUDA; }
  }
  void *args[] = {params};
  CUresult r;
  if (coop && !g_dyn.LaunchKernelEx) { if (err) *err = "cuLaunchKernelEx missing: cooperative launch unavailable"; return MXB_ERR_CUDA; }
  if (g_dyn.LaunchKernelEx && (pdl || coop)) {
    // programmatic dependent launch, like the ahead-of-time kernels (every body starts with griddepcontrol.wait)
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDimX = grid; cfg.gridDimY = 1; cfg.gridDimZ = 1;
    cfg.blockDimX = block; cfg.blockDimY = 1; cfg.blockDimZ = 1;
    cfg.sharedMemBytes = smem;
    cfg.hStream = (CUstream)stream;
    CUlaunchAttribute attr[1];
    memset(attr, 0, sizeof attr);
    if (coop) {
      attr[0].id = CU_LAUNCH_ATTRIBUTE_COOPERATIVE;
      attr[0].value.cooperative = 1;
    } else {
      attr[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
      attr[0].value.programmaticStreamSerializationAllowed = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    r = g_dyn.LaunchKernelEx(&cfg, f, args, nullptr);
  } else {
    r = g_dyn.LaunchKernel(f, grid, 1, 1, block, 1, 1, smem, (CUstream)stream, args, nullptr);
  }
  if (r != 0) {
    const char *s = nullptr;
    g_dyn.GetErrorStringDrv(r, &s);
    if (err) *err = std::string("cuLaunchKernel: ") + (s ? s : "?");
    return MXB_ERR_CUDA;
  }
  return MXB_OK;
}

// NVRTC only (no driver): used by the CPU-side tests to prove that a generated kernel builds for sm_100a
int jit_compile_only(const std::string &source, std::string *log) {
  std::lock_guard<std::mutex> lock(g_mu);
  std::string err;
  // the driver library may be absent on a CPU box: load NVRTC by hand
  std::vector<std::string> roots;
  for (const char *ev : {"MXB_CUDA_HOME", "CUDA_HOME", "CUDA_PATH"}) if (const char *v = getenv(ev)) roots.push_back(v);
  roots.push_back("/usr/local/cuda");
  std::vector<std::string> nv;
  for (const std::string &r : roots) { nv.push_back(r + "/lib64/libnvrtc.so.12"); nv.push_back(r + "/lib64/libnvrtc.so"); }
  nv.push_back("libnvrtc.so.12");
  nv.push_back("libnvrtc.so");
  void *hn = open_first(nv);
  if (!hn) { if (log) *log = "libnvrtc not found"; return MXB_ERR_JIT; }
  Dyn d;
#define MXB_SYM2(field, name) *(void **)(&d.field) = dlsym(hn, name); if (!d.field) { if (log) *log = std::string("missing symbol ") + name; return MXB_ERR_JIT; }
  MXB_SYM2(CreateProgram, "nvrtcCreateProgram")
  MXB_SYM2(CompileProgram, "nvrtcCompileProgram")
  MXB_SYM2(GetProgramLogSize, "nvrtcGetProgramLogSize")
  MXB_SYM2(GetProgramLog, "nvrtcGetProgramLog")
  MXB_SYM2(DestroyProgram, "nvrtcDestroyProgram")
  MXB_SYM2(GetErrorString, "nvrtcGetErrorString")
#undef MXB_SYM2
  std::string incdir;
  for (const std::string &r : roots) if (file_exists(r + "/include/cuda_bf16.h")) { incdir = r + "/include"; break; }
  nvrtcProgram prog = nullptr;
  const char *hdr_src[] = {kDeviceHeaderText};
  const char *hdr_name[] = {"mxb_device.cuh"};
  nvrtcResult r = d.CreateProgram(&prog, source.c_str(), "mxb_jit.cu", 1, hdr_src, hdr_name);
  if (r != 0) { if (log) *log = "nvrtcCreateProgram failed"; return MXB_ERR_JIT; }
  const std::string inc = "--include-path=" + incdir;
  const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", inc.c_str(), "-lineinfo"};
  r = d.CompileProgram(prog, 4, opts);
  size_t n = 0;
  d.GetProgramLogSize(prog, &n);
  std::string lg(n, '\0');
  if (n) d.GetProgramLog(prog, &lg[0]);
  if (log) *log = lg;
  d.DestroyProgram(&prog);
  return r == 0 ? MXB_OK : MXB_ERR_JIT;
}

}  // namespace mxbh
