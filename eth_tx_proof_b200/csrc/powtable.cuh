// Two-level power tables shared by the NTT passes, the quotient kernels (static and NVRTC-compiled) and FRI.
#pragma once
#include "gl.cuh"

namespace ntt {

// base^e for e < 2^bits via two tables: hi[e >> lo_bits] * lo[e & mask].  `hi` may carry a scale.
struct PowTable {
  const uint64_t* lo;
  const uint64_t* hi;
  int lo_bits;
  uint32_t mask;
#if defined(__CUDACC__)
  __device__ __forceinline__ uint64_t get(uint32_t e) const {
    return gl::mul(__ldg(hi + (e >> lo_bits)), __ldg(lo + (e & mask)));
  }
#endif
};

}  // namespace ntt
