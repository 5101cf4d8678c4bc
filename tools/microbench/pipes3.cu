// Does a DFMA cost more than one issue/operand slot on B200 when its three operands are distinct registers?
// Each kernel runs N DFMA-type ops + N IADD3 per chain step; operands differ per variant.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_ITER 2048
#define CHAINS 8

// variant 0: acc = acc*m + c (m, c shared)   1: acc = x_i*y_i + acc (distinct)   2: acc = x_i * 15.0 + acc (immediate)
// variant 3: acc = acc + x_i (DADD)          4: acc = xs*15.0 + acc with xs shared by all chains
template <int V, int WITH_INT, int WITH_FP>
__global__ void k(uint64_t* out, uint32_t a, uint32_t b) {
  double acc[CHAINS], x[CHAINS], y[CHAINS];
  uint32_t ia[CHAINS], ib[CHAINS];
  double m = (double)b * 1e-3, c = (double)a;
  for (int i = 0; i < CHAINS; i++) {
    acc[i] = i + threadIdx.x; x[i] = 1.0 + 1e-9 * (i + a); y[i] = 1.0 - 1e-9 * (i + b);
    ia[i] = i + a + threadIdx.x; ib[i] = 3 * i + b;
  }
  for (int it = 0; it < N_ITER; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (WITH_FP) {
        if (V == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));
        if (V == 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i]) : "d"(x[i]), "d"(y[(i + 3) % CHAINS]));
        if (V == 2) asm volatile("fma.rn.f64 %0, %1, 0d402E000000000000, %0;" : "+d"(acc[i]) : "d"(x[i]));
        if (V == 3) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(x[i]));
        if (V == 4) asm volatile("fma.rn.f64 %0, %1, 0d402E000000000000, %0;" : "+d"(acc[i]) : "d"(m));
      }
      if (WITH_INT) asm volatile("add.u32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ib[(i + 1) % CHAINS]));
      if (WITH_INT == 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(ib[i]) : "r"(ia[(i + 5) % CHAINS]));
    }
  }
  double s = 0; uint32_t t = 0;
  for (int i = 0; i < CHAINS; i++) { s += acc[i] + x[i] + y[i]; t += ia[i] + ib[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;
}

template <class K>
void run(const char* name, K kern, uint64_t* d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int warps_per_smsp = 8, threads = 256, blocks = sms * (warps_per_smsp * 4 * 32 / threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(d, 3, 5); kern<<<blocks, threads>>>(d, 3, 5);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0); kern<<<blocks, threads>>>(d, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ns = best * 1e6 / ((double)warps_per_smsp * N_ITER * CHAINS);
  printf("%-44s %.2f clk per chain step (warp, SMSP) @%d MHz\n", name, ns * clk / 1e6, clk / 1000);
}

int main() {
  uint64_t* d; cudaMalloc(&d, 1 << 26);
  run("iadd3 alone", k<0, 1, 0>, d);
  run("iadd3 + lop3 alone", k<0, 2, 0>, d);
  run("dfma shared operands", k<0, 0, 1>, d);
  run("dfma 3 distinct regs", k<1, 0, 1>, d);
  run("dfma reg*imm+reg", k<2, 0, 1>, d);
  run("dadd", k<3, 0, 1>, d);
  run("dfma shared*imm+reg", k<4, 0, 1>, d);
  run("dfma shared operands | iadd3", k<0, 1, 1>, d);
  run("dfma 3 distinct regs | iadd3", k<1, 1, 1>, d);
  run("dfma reg*imm+reg | iadd3", k<2, 1, 1>, d);
  run("dadd | iadd3", k<3, 1, 1>, d);
  run("dfma shared*imm+reg | iadd3", k<4, 1, 1>, d);
  run("dfma shared operands | iadd3 + lop3", k<0, 2, 1>, d);
  run("dfma 3 distinct regs | iadd3 + lop3", k<1, 2, 1>, d);
  run("dfma reg*imm+reg | iadd3 + lop3", k<2, 2, 1>, d);
  run("dadd | iadd3 + lop3", k<3, 2, 1>, d);
  run("dfma shared*imm+reg | iadd3 + lop3", k<4, 2, 1>, d);
  return 0;
}
