// Conversion-pipe microbenchmarks for sm_100a (B200): can the u32 -> f64 plane conversions of the Poseidon
// MDS layer move from the FMA pipe (two IMAD.MOV per value: the 2^52 bit trick) to I2F.F64.U32?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_ITER 4096
#define CHAINS 8

__global__ void k_i2f(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t x[CHAINS]; double acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) { x[i] = i + threadIdx.x + a; acc[i] = 0; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      double d;
      asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(x[i]));
      x[i] = (uint32_t)__double2loint(d) + b;   // 1 ALU op keeps the chain alive
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < CHAINS; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_i2f_dfma(uint64_t* out, uint32_t a, uint32_t b) {  // 1 I2F : 4 DFMA
  uint32_t x[CHAINS]; double acc[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { x[i] = i + threadIdx.x + a; acc[i] = i; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      double d;
      asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(x[i]));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(d));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(c), "d"(m));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      x[i] += b;
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
__global__ void k_dfma4_iadd(uint64_t* out, uint32_t a, uint32_t b) {  // same without the I2F
  uint32_t x[CHAINS]; double acc[CHAINS];
  double m = (double)b, c = (double)a;
  for (int i = 0; i < CHAINS; i++) { x[i] = i + threadIdx.x + a; acc[i] = i; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(c), "d"(m));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
      x[i] += b;
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}
__global__ void k_i2f_imad(uint64_t* out, uint32_t a, uint32_t b) {  // 1 I2F : 4 IMAD
  uint32_t x[CHAINS], y[CHAINS]; double acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) { x[i] = i + threadIdx.x + a; y[i] = i; acc[i] = 0; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      double d;
      asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(x[i]));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(a));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(a));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(a));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(a));
      x[i] = (uint32_t)__double2hiint(d) + y[i];
    }
  }
  uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) t += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// f64 -> integer the other way: F2I.U64.F64 (one instruction instead of the 7-instruction plane combine?)
__global__ void k_f2i64(uint64_t* out, uint32_t a, uint32_t b) {
  double d[CHAINS]; uint64_t acc[CHAINS];
  for (int i = 0; i < CHAINS; i++) { d[i] = i + threadIdx.x + a; acc[i] = 0; }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      uint64_t v;
      asm volatile("cvt.rzi.u64.f64 %0, %1;" : "=l"(v) : "d"(d[i]));
      d[i] = __hiloint2double(0x43300000, (int)(uint32_t)v + b);
    }
  }
  uint64_t t = 0;
  for (int i = 0; i < CHAINS; i++) t += (uint64_t)d[i];
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
  double groups_per_smsp = (double)warps_per_smsp * N_ITER * CHAINS;
  double ns = best * 1e6 / groups_per_smsp;
  printf("%-22s %8.3f ms  %.2f clk per group of %d instr (warp, SMSP) @%d MHz (max)\n", name, best, ns * clk / 1e6, instr_per_iter_chain, clk / 1000);
}

int main() {
  uint64_t* d; cudaMalloc(&d, 1 << 26);
  run("i2f.f64.u32 + iadd", k_i2f, 2, d);
  run("4 dfma + iadd", k_dfma4_iadd, 5, d);
  run("i2f + 4 dfma + iadd", k_i2f_dfma, 6, d);
  run("i2f + 4 imad + iadd", k_i2f_imad, 6, d);
  run("f2i.u64.f64 + 2", k_f2i64, 3, d);
  return 0;
}
