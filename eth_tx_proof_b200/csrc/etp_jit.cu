// NVRTC back end of the constraint programs (cprog.h): CUDA source -> sm_100a cubin -> cudaLibrary -> kernel.
// The device headers the generated kernel includes (gl.cuh, powtable.cuh, quotient_rt.cuh) are embedded in the
// library as source text by the Makefile (jit_headers.inc), so a registered table compiles with exactly the field
// arithmetic of the built-in kernels.  Product code: if NVRTC or the device is missing the call fails loudly.
#include <nvrtc.h>

#include <mutex>
#include <unordered_map>

#include "cprog.h"
#include "ctx.cuh"

namespace {
const char* const kHeaderNames[] = {"gl.cuh", "powtable.cuh", "quotient_rt.cuh"};
const char* const kHeaderSources[] = {
#include "jit_headers.inc"
};
}  // namespace

// Compiles `source` (one extern "C" kernel `entry`) for sm_100a.  cubin_out receives the image.
// Process-wide cache of compiled programs: the contexts of one process (one per worker thread / parallel.ProverPool)
// register the same tables, and a wide table takes seconds to compile (keccak shape: 2400 columns, 600 constraints).
namespace {
std::mutex g_cubin_mutex;
std::unordered_map<std::string, std::vector<char>> g_cubin_cache;
}  // namespace

int jit_compile(etp_ctx* ctx, const std::string& source, std::vector<char>* cubin_out, std::string* log_out) {
  {
    std::lock_guard<std::mutex> lock(g_cubin_mutex);
    auto it = g_cubin_cache.find(source);
    if (it != g_cubin_cache.end()) {
      *cubin_out = it->second;
      if (log_out) log_out->clear();
      return ETP_OK;
    }
  }
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source.c_str(), "etp_cprog.cu", 3, kHeaderSources, kHeaderNames) != NVRTC_SUCCESS)
    return etp_fail(ctx, ETP_ERR_CUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "--device-int128"};
  const nvrtcResult rc = nvrtcCompileProgram(prog, 5, opts);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  std::string log(log_size, '\0');
  if (log_size > 1) nvrtcGetProgramLog(prog, &log[0]);
  if (log_out) *log_out = log;
  if (rc != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    return etp_fail(ctx, ETP_ERR_CUDA, "NVRTC compilation of the constraint program failed: %s: %.300s", nvrtcGetErrorString(rc), log.c_str());
  }
  size_t n = 0;
  if (nvrtcGetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
    nvrtcDestroyProgram(&prog);
    return etp_fail(ctx, ETP_ERR_CUDA, "NVRTC produced no cubin");
  }
  cubin_out->resize(n);
  nvrtcGetCUBIN(prog, cubin_out->data());
  nvrtcDestroyProgram(&prog);
  {
    std::lock_guard<std::mutex> lock(g_cubin_mutex);
    g_cubin_cache.emplace(source, *cubin_out);
  }
  return ETP_OK;
}

int jit_load(etp_ctx* ctx, const std::vector<char>& cubin, const char* entry, JitKernel* out) {
  ETP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaLibrary_t lib;
  ETP_CUDA(ctx, cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  cudaKernel_t k;
  cudaError_t e = cudaLibraryGetKernel(&k, lib, entry);
  if (e != cudaSuccess) {
    cudaLibraryUnload(lib);
    return etp_fail(ctx, ETP_ERR_CUDA, "cudaLibraryGetKernel(%s) failed: %s", entry, cudaGetErrorString(e));
  }
  out->library = (void*)lib;
  out->kernel = (void*)k;
  return ETP_OK;
}

void jit_unload(JitKernel* k) {
  if (k && k->library) cudaLibraryUnload((cudaLibrary_t)k->library);
  if (k) { k->library = nullptr; k->kernel = nullptr; }
}

// C ABI helper used by the CPU tests (no device needed): program words -> cubin size; proves that a table can be
// compiled for sm_100a in this process.
extern "C" int etp_cprog_compile_check(const uint64_t* program, size_t n_words, size_t* cubin_bytes_out, char* err, size_t err_len) {
  cprog::Program p;
  const std::string why = cprog::parse(program, n_words, 16, 8, &p);
  auto fail = [&](const std::string& m) { if (err && err_len) snprintf(err, err_len, "%s", m.c_str()); return ETP_ERR_INVALID; };
  if (!why.empty()) return fail(why);
  etp_ctx tmp;  // only its error string is used
  std::vector<char> cubin;
  std::string log;
  const int rc = jit_compile(&tmp, cprog::generate_cuda(p), &cubin, &log);
  if (rc != ETP_OK) { fail(tmp.err); return rc; }
  if (cubin_bytes_out) *cubin_bytes_out = cubin.size();
  return ETP_OK;
}
