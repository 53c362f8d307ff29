// zpq_spec.cpp -- finds (or builds) the specialised kernels for a block header:
//   1. ahead-of-time instances generated at build time for the built-in models,
//   2. otherwise NVRTC: zpq_codegen.cpp writes the model source, libnvrtc compiles it for
//      sm_100a, the cubin is loaded through the runtime's library API; cached per header.
// ZPQ_SPECIALIZE=0 disables both and leaves the run-time model walker (GenericModel) in charge.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "zpq_aot.h"
#include "zpq_host.h"

namespace zpq {

std::string generate_model_source(const Header& hdr, const std::string& name, const std::string& enc_kernel,
                                  const std::string& dec_kernel, bool* compiled_hcomp, int* duo_g, bool* fdec);

std::string generate_post_source(const Bytes& prog, int ph, int pm);

namespace {

#include "gen/zpq_embed.inc"   // kEmbedPlan[], kEmbedDevcore[]: the headers NVRTC needs, as text

struct Registry {
  std::mutex mu;
  std::map<Bytes, SpecKernels> table;
  std::map<Bytes, std::string> failed;   // headers NVRTC could not build (message kept)
};
Registry& registry() { static Registry r; return r; }

// ---- NVRTC through dlopen: no link-time dependency, absence is reported, not fatal ----
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* so = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  bool ok = false;
  Nvrtc() {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names) if ((so = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!so) return;
#define ZPQ_SYM(f) *(void**)(&f) = dlsym(so, "nvrtc" #f)
    ZPQ_SYM(CreateProgram); ZPQ_SYM(CompileProgram); ZPQ_SYM(GetCUBINSize); ZPQ_SYM(GetCUBIN);
    ZPQ_SYM(GetProgramLogSize); ZPQ_SYM(GetProgramLog); ZPQ_SYM(DestroyProgram);
#undef ZPQ_SYM
    ok = CreateProgram && CompileProgram && GetCUBINSize && GetCUBIN && GetProgramLogSize && GetProgramLog && DestroyProgram;
  }
};
Nvrtc& nvrtc() { static Nvrtc n; return n; }

bool specialize_enabled() {
  const char* e = getenv("ZPQ_SPECIALIZE");
  return !(e && *e == '0');
}

}  // namespace

AotRegistrar::AotRegistrar(const unsigned char* header, size_t len, const void* enc, const void* enc_lanes, const void* enc_duo, int duo_g,
                           const void* dec, const void* dec_fast, const char* origin) {
  Registry& r = registry();
  std::lock_guard<std::mutex> g(r.mu);
  SpecKernels k; k.enc = enc; k.enc_lanes = enc_lanes; k.enc_duo = enc_duo; k.duo_g = duo_g; k.dec = dec; k.dec_fast = dec_fast; k.origin = origin;
  r.table[Bytes(header, header + len)] = k;
}

// Compile a model source to a cubin with NVRTC.  Throws Failure with the compiler log.
Bytes nvrtc_compile(const std::string& src) {
  Nvrtc& n = nvrtc();
  if (!n.ok) throw Failure(ZPQ_E_UNSUPPORTED, "libnvrtc is not available");
  nvrtcProgram prog = nullptr;
  const char* hdr_src[5] = {kEmbedPlan, kEmbedDevcore, kEmbedPipe, kEmbedDuo, kEmbedFdec};
  const char* hdr_name[5] = {"zpq_plan.h", "zpq_devcore.cuh", "zpq_pipe.cuh", "zpq_duo.cuh", "zpq_fdec.cuh"};
  if (n.CreateProgram(&prog, src.c_str(), "zpq_model.cu", 5, hdr_src, hdr_name) != 0)
    throw Failure(ZPQ_E_CUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const int rc = n.CompileProgram(prog, (int)(sizeof opts / sizeof opts[0]), opts);
  if (rc != 0) {
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.GetProgramLog(prog, &log[0]);
    n.DestroyProgram(&prog);
    throw Failure(ZPQ_E_CONFIG, "NVRTC: " + log.substr(0, 2000));
  }
  size_t sz = 0;
  n.GetCUBINSize(prog, &sz);
  Bytes cubin(sz);
  n.GetCUBIN(prog, reinterpret_cast<char*>(cubin.data()));
  n.DestroyProgram(&prog);
  return cubin;
}

// Kernels specialised for `hdr`, or false when the run-time model walker has to be used.
bool find_spec_kernels(const Header& hdr, uint32_t smem_limit, SpecKernels& out, std::string* why_not) {
  if (!specialize_enabled()) { if (why_not) *why_not = "ZPQ_SPECIALIZE=0"; return false; }
  if (hdr.n < 1 || hdr.n > 32) { if (why_not) *why_not = "component count outside 1..32"; return false; }
  Registry& r = registry();
  std::lock_guard<std::mutex> g(r.mu);
  auto it = r.table.find(hdr.wire);
  if (it != r.table.end()) { out = it->second; return true; }
  auto bad = r.failed.find(hdr.wire);
  if (bad != r.failed.end()) { if (why_not) *why_not = bad->second; return false; }
  const char* e = getenv("ZPQ_NVRTC");
  if (e && *e == '0') { if (why_not) *why_not = "ZPQ_NVRTC=0"; return false; }
  try {
    int duo_g = 0;
    bool fd = false;
    std::string src = generate_model_source(hdr, "Model_rt", "zpq_enc_rt", "zpq_dec_rt", nullptr, &duo_g, &fd);
    Bytes cubin = nvrtc_compile(src);
    cudaLibrary_t lib = nullptr;
    cudaError_t ce = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce != cudaSuccess) throw Failure(ZPQ_E_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce));
    cudaKernel_t ke = nullptr, kl = nullptr, kd = nullptr, kq = nullptr, kf = nullptr;
    if ((ce = cudaLibraryGetKernel(&ke, lib, "zpq_enc_rt")) != cudaSuccess || (ce = cudaLibraryGetKernel(&kl, lib, "zpq_enc_rt_l")) != cudaSuccess ||
        (ce = cudaLibraryGetKernel(&kd, lib, "zpq_dec_rt")) != cudaSuccess)
      throw Failure(ZPQ_E_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce));
    if (duo_g && (ce = cudaLibraryGetKernel(&kq, lib, "zpq_enc_rt_d")) != cudaSuccess)
      throw Failure(ZPQ_E_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce));
    if (fd && (ce = cudaLibraryGetKernel(&kf, lib, "zpq_dec_rt_f")) != cudaSuccess)
      throw Failure(ZPQ_E_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce));
    // (the dynamic shared memory limit is a per-device attribute: launch_codec raises it on the device it launches on)
    SpecKernels k; k.enc = (const void*)ke; k.enc_lanes = (const void*)kl; k.enc_duo = (const void*)kq; k.duo_g = duo_g; k.dec = (const void*)kd;
    k.dec_fast = (const void*)kf; k.origin = "nvrtc";
    r.table[hdr.wire] = k;
    out = k;
    return true;
  } catch (const std::exception& ex) {
    cudaGetLastError();
    r.failed[hdr.wire] = ex.what();
    if (why_not) *why_not = ex.what();
    return false;
  }
}

// The post-processing kernel compiled for one PCOMP program (zpq_codegen.cpp: generate_post_source), or null when the program
// cannot be translated / NVRTC is not there (the interpreter pass then runs the program).  Cached per (program, ph, pm).
const void* find_post_kernel(const Bytes& prog, int ph, int pm, std::string* why_not) {
  static std::mutex mu;
  static std::map<Bytes, const void*> table;          // null = known not to build
  static std::map<Bytes, std::string> reason;
  const char* e = getenv("ZPQ_POST_NVRTC");
  if (e && *e == '0') { if (why_not) *why_not = "ZPQ_POST_NVRTC=0"; return nullptr; }
  Bytes key = prog;
  key.push_back((uint8_t)ph); key.push_back((uint8_t)pm);
  std::lock_guard<std::mutex> g(mu);
  auto it = table.find(key);
  if (it != table.end()) { if (!it->second && why_not) *why_not = reason[key]; return it->second; }
  const void* k = nullptr;
  try {
    const std::string src = generate_post_source(prog, ph, pm);
    if (src.empty()) throw Failure(ZPQ_E_UNSUPPORTED, "the program jumps into the middle of an instruction");
    Bytes cubin = nvrtc_compile(src);
    cudaLibrary_t lib = nullptr;
    cudaError_t ce = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce != cudaSuccess) throw Failure(ZPQ_E_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce));
    cudaKernel_t kk = nullptr;
    if ((ce = cudaLibraryGetKernel(&kk, lib, "zpq_post_rt")) != cudaSuccess)
      throw Failure(ZPQ_E_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce));
    k = (const void*)kk;
  } catch (const std::exception& ex) {
    cudaGetLastError();
    reason[key] = ex.what();
    if (why_not) *why_not = ex.what();
  }
  table[key] = k;
  return k;
}

// NVRTC compile of a PCOMP program without loading it (needs no GPU): the cubin size, or a negative code with the log in *log.
int64_t post_program_cubin(const Bytes& prog, int ph, int pm, std::string* source, std::string* log) {
  try {
    const std::string src = generate_post_source(prog, ph, pm);
    if (source) *source = src;
    if (src.empty()) { if (log) *log = "the program jumps into the middle of an instruction"; return ZPQ_E_UNSUPPORTED; }
    return (int64_t)nvrtc_compile(src).size();
  } catch (const Failure& f) { if (log) *log = f.what(); return f.code; }
}

// Ahead-of-time kernels need their dynamic shared memory limit raised once per device.
void spec_set_smem_limit(uint32_t bytes) {
  Registry& r = registry();
  std::lock_guard<std::mutex> g(r.mu);
  for (auto& kv : r.table) {
    cudaFuncSetAttribute(kv.second.enc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (kv.second.enc_lanes) cudaFuncSetAttribute(kv.second.enc_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (kv.second.enc_duo) cudaFuncSetAttribute(kv.second.enc_duo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    cudaFuncSetAttribute(kv.second.dec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (kv.second.dec_fast) cudaFuncSetAttribute(kv.second.dec_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
  cudaGetLastError();
}

}  // namespace zpq
