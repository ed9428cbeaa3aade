// jit.cpp — run-time specialisation of the lean kernel (lean_kernel.cuh) on one plan shape.
//
// The ahead-of-time build interprets the lowered program: one dispatch per instruction and tile.  A plan that runs
// repeatedly (a prepared aggregate over resident columns) is recompiled here with NVRTC: the very same kernel source,
// with the LeanShape (program + layout + geometry) as a `constexpr` so that every dispatch, operand decode and
// shared-memory offset folds at compile time.  Literals, column pointers, row ranges and the MVCC snapshot stay run-time
// kernel parameters, so re-running a query with other constants or more rows reuses the specialised kernel.
//
// NVRTC is loaded with dlopen (libnvrtc.so.12 ships with the CUDA toolkit of this image); when it is missing the lean
// path simply keeps interpreting on the GPU.  The cubin is loaded through the runtime's library API
// (cudaLibraryLoadData / cudaLibraryGetKernel) and launched with cudaLaunchKernel.
#include "jit.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace llkv {

// the kernel sources, embedded at build time (build.sh writes build/lean_sources.inc)
#include "build/lean_sources.inc"

namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
  void* lib = nullptr;
  bool tried = false;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NvrtcApi g_nvrtc;
std::mutex g_mu;

bool load_nvrtc(std::string& err) {
  if (g_nvrtc.tried) {
    if (!g_nvrtc.lib) err = "NVRTC is not available (libnvrtc.so.12 not found)";
    return g_nvrtc.lib != nullptr;
  }
  g_nvrtc.tried = true;
  const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) {
    err = "NVRTC is not available (libnvrtc.so.12 not found)";
    return false;
  }
#define LLKV_SYM(field, name)                                             \
  do {                                                                    \
    *(void**)(&g_nvrtc.field) = dlsym(h, name);                           \
    if (!g_nvrtc.field) {                                                 \
      err = std::string("NVRTC symbol missing: ") + name;                 \
      dlclose(h);                                                         \
      return false;                                                       \
    }                                                                     \
  } while (0)
  LLKV_SYM(CreateProgram, "nvrtcCreateProgram");
  LLKV_SYM(CompileProgram, "nvrtcCompileProgram");
  LLKV_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  LLKV_SYM(GetCUBIN, "nvrtcGetCUBIN");
  LLKV_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  LLKV_SYM(GetProgramLog, "nvrtcGetProgramLog");
  LLKV_SYM(DestroyProgram, "nvrtcDestroyProgram");
  LLKV_SYM(GetErrorString, "nvrtcGetErrorString");
#undef LLKV_SYM
  g_nvrtc.lib = h;
  return true;
}

// LeanShape is made of uint32_t fields only: brace elision lets it be written as one flat initializer list
static_assert(sizeof(LeanShape) % 4 == 0, "LeanShape must be an array of 32-bit words");

std::string shape_source(const LeanShape& s, int ctas_per_sm) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&s);
  size_t n = sizeof(LeanShape) / 4;
  while (n > 1 && w[n - 1] == 0) --n;  // trailing zeros are value-initialised
  int n_stash = 0;
  for (uint32_t i = 0; i < s.n_code && i < (uint32_t)kMaxFastInstr; ++i)
    if (s.code[i].op >= FO_COUNT_STAR && s.code[i].op <= FO_FIRSTNAN) ++n_stash;
  if (n_stash == 0) n_stash = 1;
  int n_tmps = 1;  // temporaries live in registers in a specialised build (LeanTile::treg)
  for (uint32_t i = 0; i < (uint32_t)kMaxFastInstr && s.code[i].op != FO_END; ++i)
    if (s.code[i].op == FO_ST_TMP && (int)s.code[i].a + 1 > n_tmps) n_tmps = (int)s.code[i].a + 1;
  std::ostringstream o;
  o << "#include \"lean_kernel.cuh\"\n"
       "namespace llkv {\n"
       "__device__ constexpr LeanShape kJitShape = {";
  for (size_t i = 0; i < n; ++i) {
    if (i) o << ',';
    if ((i & 31) == 0) o << '\n';
    o << w[i] << 'u';
  }
  o << "};\n"
       "struct LeanJitCfg {\n"
       "  static constexpr bool kStatic = true;\n"
       "  static __device__ __forceinline__ const LeanShape& shape(const LeanPlan&) { return kJitShape; }\n"
       "  static __host__ __device__ constexpr FInstr code(int pc) { return kJitShape.code[pc]; }\n"
       "  static constexpr bool kPartition = kJitShape.partition != 0;\n"
       "  static constexpr bool kPacked = kJitShape.partition == 2;\n"
       "  static constexpr bool kDefer = kJitShape.n_keys != 0 && (kJitShape.direct_global == 0 || kPartition);\n"
       "  static constexpr bool kSplitSlow = kJitShape.n_keys != 0 && kJitShape.direct_global == 0 && !kPartition;\n"
       "  static constexpr int kStash = " << n_stash << ";\n"
       "  static constexpr int kTmps = " << n_tmps << ";\n"
       "};\n"
       "}  // namespace llkv\n"
       "extern \"C\" __global__ void __launch_bounds__("
    << (s.nc + 32) << ", " << ctas_per_sm
    << ") llkv_lean_jit(const __grid_constant__ llkv::LeanPlan p) {\n"
       "  llkv::lean_body<"
    << s.rows_per_thread
    << ", llkv::LeanJitCfg>(p);\n"
       "}\n";
  return o.str();
}

struct Entry {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  bool failed = false;
  std::string log;
};
std::map<std::string, Entry> g_cache;  // key: device + ctas + shape bytes

}  // namespace

int jit_compile_cubin(const LeanShape& shape, int ctas_per_sm, std::vector<char>& cubin, std::string& log) {
  std::string err;
  if (!load_nvrtc(err)) {
    log = err;
    return -1;
  }
  const std::string src = shape_source(shape, ctas_per_sm);
  const char* headers[] = {kSrc_lean_kernel_cuh, kSrc_device_util_cuh, kSrc_plan_h};
  const char* names[] = {"lean_kernel.cuh", "device_util.cuh", "plan.h"};
  nvrtcProgram prog = nullptr;
  int rc = g_nvrtc.CreateProgram(&prog, src.c_str(), "llkv_lean_jit.cu", 3, headers, names);
  if (rc != 0) {
    log = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(rc);
    return -1;
  }
  // -lineinfo only on request (ncu source view): cicc 12.9 has crashed generating line info for some instantiations
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--device-int128", "-lineinfo"};
  rc = g_nvrtc.CompileProgram(prog, getenv("LLKV_GPU_JIT_LINEINFO") ? 4 : 3, opts);
  size_t log_size = 0;
  g_nvrtc.GetProgramLogSize(prog, &log_size);
  if (log_size > 1) {
    log.resize(log_size);
    g_nvrtc.GetProgramLog(prog, &log[0]);
  }
  if (rc != 0) {
    log = std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(rc) + "\n" + log;
    g_nvrtc.DestroyProgram(&prog);
    return -1;
  }
  size_t n = 0;
  g_nvrtc.GetCUBINSize(prog, &n);
  cubin.resize(n);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  g_nvrtc.DestroyProgram(&prog);
  if (const char* dir = getenv("LLKV_GPU_JIT_DUMP")) {  // keep the specialised cubin + source for cuobjdump -sass
    static int serial = 0;
    char path[512];
    snprintf(path, sizeof(path), "%s/llkv_lean_jit_%d.cubin", dir, serial);
    if (FILE* f = fopen(path, "wb")) {
      fwrite(cubin.data(), 1, cubin.size(), f);
      fclose(f);
    }
    snprintf(path, sizeof(path), "%s/llkv_lean_jit_%d.cu", dir, serial++);
    if (FILE* f = fopen(path, "w")) {
      fwrite(src.data(), 1, src.size(), f);
      fclose(f);
    }
  }
  return 0;
}

static Entry* jit_entry(int device, const LeanPlan& plan, int ctas_per_sm) {
  std::string key;
  key.reserve(sizeof(LeanShape) + 16);
  key.append(reinterpret_cast<const char*>(&device), sizeof(device));
  key.append(reinterpret_cast<const char*>(&ctas_per_sm), sizeof(ctas_per_sm));
  key.append(reinterpret_cast<const char*>(&plan.s), sizeof(LeanShape));
  Entry* e;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      Entry ne;
      std::vector<char> cubin;
      if (jit_compile_cubin(plan.s, ctas_per_sm, cubin, ne.log) != 0) {
        ne.failed = true;
      } else {
        cudaError_t ce = cudaLibraryLoadData(&ne.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (ce == cudaSuccess) ce = cudaLibraryGetKernel(&ne.kernel, ne.lib, "llkv_lean_jit");
        if (ce == cudaSuccess)
          ce = cudaFuncSetAttribute((const void*)ne.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.s.smem_total);
        if (ce != cudaSuccess) {
          ne.failed = true;
          ne.log = std::string("loading the specialised kernel: ") + cudaGetErrorString(ce);
          cudaGetLastError();
        }
      }
      if (ne.failed && getenv("LLKV_GPU_JIT_VERBOSE")) fprintf(stderr, "[llkv jit] %s\n", ne.log.c_str());
      it = g_cache.emplace(key, ne).first;
    }
    e = &it->second;
  }
  return e;
}

bool jit_ready(int device, const LeanPlan& plan, int ctas_per_sm) { return !jit_entry(device, plan, ctas_per_sm)->failed; }

cudaError_t jit_launch(int device, const LeanPlan& plan, int ctas_per_sm, uint32_t grid, cudaStream_t stream, bool* used, std::string* why) {
  *used = false;
  Entry* e = jit_entry(device, plan, ctas_per_sm);
  if (e->failed) {
    if (why) *why = e->log;
    return cudaSuccess;  // the caller interprets instead
  }
  LeanPlan copy = plan;
  void* args[] = {&copy};
  cudaError_t ce = cudaLaunchKernel((const void*)e->kernel, dim3(grid), dim3(plan.s.nc + 32), args, plan.s.smem_total, stream);
  if (ce == cudaSuccess) *used = true;
  return ce;
}

}  // namespace llkv
