// Pipe-throughput microbenchmarks for sm_100a (B200): which instruction mix should the Poseidon MDS
// and the Goldilocks reduction use?  Reports cycles per warp-instruction per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_ITER 4096
#define CHAINS 8

__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b) {
  uint64_t acc[CHAINS];
  uint32_t x = a + threadIdx.x;
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      // dependent through the multiplicand so ptxas cannot split the accumulate away
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 1) % CHAINS]), "r"(b));
    }
  }
  uint64_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
}
__global__ void k_imul_wide(uint64_t* out, uint32_t a, uint32_t b) {  // IMAD.WIDE with RZ addend
  uint64_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x + a;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"((uint32_t)acc[i] + 1), "r"(b));
  }
  uint64_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad32(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x + a;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
  }
  uint32_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imadhi(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x + a;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
  }
  uint32_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x + a;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("add.u32 %0, %0, %1; xor.b32 %0, %0, %2;" : "+r"(acc[i]) : "r"(b), "r"(a));
  }
  uint32_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_add64cc(uint64_t* out, uint32_t a, uint32_t b) {  // 64-bit add with carry chain (IADD3 + IADD3.X)
  uint64_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x + a;
  uint64_t bb = ((uint64_t)b << 32) | a;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"(bb));
  }
  uint64_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) acc[i] = i + threadIdx.x;
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
  }
  double s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s;
}
// mixes: do the pipes overlap?
__global__ void k_mix_dfma_iadd(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS]; uint32_t ia[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; ia[i] = i + a + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ia[i]) : "r"(b));
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += ia[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
__global__ void k_mix_dfma_imad(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS]; uint32_t ia[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; ia[i] = i + a + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(ia[i]) : "r"(b), "r"(a));
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += ia[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
__global__ void k_mix3(uint64_t* out, uint32_t a, uint32_t b) {  // DFMA + IMAD + IADD
  double acc[CHAINS]; uint32_t ia[CHAINS], ib[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; ia[i] = i + a + threadIdx.x; ib[i] = i * 3 + a + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(ia[i]) : "r"(b), "r"(a));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"(b));
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += ia[i] + ib[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
__global__ void k_mix_imad_iadd(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t ia[CHAINS], ib[CHAINS];
  for (int i = 0; i < CHAINS; i++) { ia[i] = i + a + threadIdx.x; ib[i] = i * 3 + a + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(ia[i]) : "r"(b), "r"(a));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"(b));
    }
  }
  uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) t += ia[i] + ib[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}


// DFMA with three distinct register operands per instruction (register-file bandwidth check)
__global__ void k_dfma3(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS], x[CHAINS], y[CHAINS];
  for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; x[i] = 1.0 + i * 1e-9 + a * 1e-12; y[i] = 1.0 - i * 1e-9 + b * 1e-12; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i]) : "d"(x[i]), "d"(y[(i + 3) % CHAINS]));
  }
  double s = 0;
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s;
}
// the instruction mix of the hybrid Poseidon round: 2 DFMA : 3 IMAD (imm) : 4 IADD3 per group
__global__ void k_mix_234(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS]; uint32_t ia[CHAINS], ib[CHAINS], ic[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; ia[i] = i + a + threadIdx.x; ib[i] = i * 3 + a + threadIdx.x; ic[i] = i * 5 + b + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("mad.lo.u32 %0, %1, 17, %0;" : "+r"(ia[i]) : "r"(ib[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"(b));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(ic[i]) : "r"(ib[i]));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(c), "d"(m));
      asm volatile("mad.lo.u32 %0, %1, 41, %0;" : "+r"(ia[i]) : "r"(ic[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ic[i]) : "r"(a));
      asm volatile("mad.lo.u32 %0, %1, 13, %0;" : "+r"(ia[i]) : "r"(ib[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"(ic[i]));
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += ia[i] + ib[i] + ic[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
// the current mix: 6 DFMA/DADD : 2 IMAD.WIDE : 5 IADD3
__global__ void k_mix_cur(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS]; uint32_t ib[CHAINS], ic[CHAINS]; uint64_t w[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; ib[i] = i * 3 + a + threadIdx.x; ic[i] = i * 5 + b + threadIdx.x; w[i] = i + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(ib[i]), "r"(ic[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"((uint32_t)w[i]));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(c), "d"(m));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ic[i]) : "r"((uint32_t)(w[i] >> 32)));
      asm volatile("add.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(m));
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(ic[i]), "r"(ib[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"((uint32_t)w[i]));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(ic[i]) : "r"((uint32_t)(w[i] >> 32)));
      asm volatile("add.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(c));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"(ic[i]));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(c), "d"(c));
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += ib[i] + ic[i] + (uint32_t)w[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
// integer only: 1 IMAD.WIDE : 1 IMAD : 3 IADD3
__global__ void k_mix_int(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t ia[CHAINS], ib[CHAINS], ic[CHAINS]; uint64_t w[CHAINS];
  for (int i = 0; i < CHAINS; i++) { ia[i] = i + a + threadIdx.x; ib[i] = i * 3 + a + threadIdx.x; ic[i] = i * 5 + b + threadIdx.x; w[i] = i + threadIdx.x; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(ib[i]), "r"(ic[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ib[i]) : "r"((uint32_t)w[i]));
      asm volatile("mad.lo.u32 %0, %1, 41, %0;" : "+r"(ia[i]) : "r"(ic[i]));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ic[i]) : "r"((uint32_t)(w[i] >> 32)));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(ib[i]) : "r"(ia[i]));
    }
  }
  uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) t += ia[i] + ib[i] + ic[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <class K>
void run(const char* name, K k, int instr_per_iter_chain, uint64_t* d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int warps_per_smsp = 8, threads = 256, blocks = sms * (warps_per_smsp * 4 * 32 / threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(d, 3, 5);
  k<<<blocks, threads>>>(d, 3, 5);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0); k<<<blocks, threads>>>(d, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double winstr_per_smsp = (double)warps_per_smsp * N_ITER * CHAINS * instr_per_iter_chain;
  // SM clock under load is unknown here: report ns per warp-instr per SMSP and the cycles at max clock
  double ns = best * 1e6 / winstr_per_smsp;
  printf("%-18s %8.3f ms  %.3f ns/warp-instr/SMSP  = %.2f clk @%d MHz (max)\n", name, best, ns, ns * clk / 1e6, clk / 1000);
}

int main() {
  uint64_t* d; cudaMalloc(&d, 1 << 26);
  run("imad.wide acc", k_imad_wide, 1, d);
  run("imul.wide", k_imul_wide, 2, d);   // + 1 IADD for the dependency
  run("imad.lo 32", k_imad32, 1, d);
  run("imad.hi 32", k_imadhi, 1, d);
  run("iadd+xor", k_iadd3, 2, d);
  run("add.u64", k_add64cc, 2, d);
  run("dfma", k_dfma, 1, d);
  run("dfma+iadd", k_mix_dfma_iadd, 2, d);
  run("dfma+imad", k_mix_dfma_imad, 2, d);
  run("dfma+imad+iadd", k_mix3, 3, d);
  run("imad+iadd", k_mix_imad_iadd, 2, d);
  run("dfma 3 regs", k_dfma3, 1, d);
  run("2dfma:3imad:4iadd", k_mix_234, 9, d);
  run("6fp64:2imw:5iadd", k_mix_cur, 13, d);
  run("1imw:1imad:3iadd", k_mix_int, 5, d);
  return 0;
}
