// NVRTC back end of the constraint programs (cprog.h): CUDA source -> sm_100a cubin -> cudaLibrary -> kernel.
// The device headers the generated kernel includes (gl.cuh, powtable.cuh, quotient_rt.cuh) are embedded in the
// library as source text by the Makefile (jit_headers.inc), so a registered table compiles with exactly the field
// arithmetic of the built-in kernels.  Product code: if NVRTC or the device is missing the call fails loudly.
#include <nvrtc.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>

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

// On-disk cache of compiled programs (the analogue of the reference's persisted prover state,
// /root/reference/common/src/prover_state/persistence.rs:32-38: circuits are built once and reloaded at worker start): with
// ETP_CUBIN_CACHE=<dir> a program compiled by any worker process is loaded from <dir>/etp_<key>.cubin by the next one instead
// of going through NVRTC again (seconds for a 2400-column table).  Key = FNV-1a of the source, the NVRTC version and the
// target; file = "ETPCUBN1", payload size, FNV-1a of the payload, payload — a truncated or foreign file is ignored.
namespace {
uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ULL) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
  return h;
}
std::string disk_cache_path(const std::string& source) {
  const char* dir = getenv("ETP_CUBIN_CACHE");
  if (!dir || !*dir) return "";
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  uint64_t h = fnv1a(source.data(), source.size());
  const char target[] = "sm_100a";
  h = fnv1a(target, sizeof target, h);
  h = fnv1a(&major, sizeof major, h);
  h = fnv1a(&minor, sizeof minor, h);
  char name[64];
  snprintf(name, sizeof name, "/etp_%016llx.cubin", (unsigned long long)h);
  return std::string(dir) + name;
}
bool disk_cache_load(const std::string& path, std::vector<char>* cubin) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  uint64_t hdr[3] = {0, 0, 0};
  bool ok = fread(hdr, 8, 3, f) == 3 && hdr[0] == 0x314E425543505445ULL && hdr[1] > 0 && hdr[1] < ((uint64_t)1 << 31);
  if (ok) {
    cubin->resize(hdr[1]);
    ok = fread(cubin->data(), 1, hdr[1], f) == hdr[1] && fgetc(f) == EOF && fnv1a(cubin->data(), cubin->size()) == hdr[2];
  }
  fclose(f);
  if (!ok) cubin->clear();
  return ok;
}
void disk_cache_store(const std::string& path, const std::vector<char>& cubin) {
  char tmp[32];
  snprintf(tmp, sizeof tmp, ".tmp%ld", (long)getpid());
  const std::string part = path + tmp;
  FILE* f = fopen(part.c_str(), "wb");
  if (!f) return;  // the cache is best effort: an unwritable directory only costs the next start-up its compile time
  const uint64_t hdr[3] = {0x314E425543505445ULL, (uint64_t)cubin.size(), fnv1a(cubin.data(), cubin.size())};
  const bool ok = fwrite(hdr, 8, 3, f) == 3 && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  if (fclose(f) != 0 || !ok || rename(part.c_str(), path.c_str()) != 0) remove(part.c_str());
}
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
  const std::string disk = disk_cache_path(source);
  if (!disk.empty() && disk_cache_load(disk, cubin_out)) {
    if (log_out) log_out->clear();
    std::lock_guard<std::mutex> lock(g_cubin_mutex);
    g_cubin_cache.emplace(source, *cubin_out);
    return ETP_OK;
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
  if (!disk.empty()) disk_cache_store(disk, *cubin_out);
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

// The CUDA source the library generates for a program (what NVRTC is given): for inspection and for tests that re-interpret the
// generated code against the program.  Returns the source length (without the terminator) or a negative error; copies at most
// cap - 1 bytes.
extern "C" int64_t etp_cprog_generate_cuda(const uint64_t* program, size_t n_words, char* out, size_t cap) {
  cprog::Program p;
  if (!cprog::parse(program, n_words, 16, 8, &p).empty()) return ETP_ERR_INVALID;
  const std::string src = cprog::generate_cuda(p);
  if (out && cap) {
    const size_t n = src.size() < cap - 1 ? src.size() : cap - 1;
    memcpy(out, src.data(), n);
    out[n] = 0;
  }
  return (int64_t)src.size();
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
