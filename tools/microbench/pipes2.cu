// Which integer-multiply forms collide with DFMA on B200?  Pairs of independent chains: one DFMA per
// chain step next to one integer multiply of each form.  "clk per pair" close to max(a, b) = overlap,
// close to a + b = the two share a datapath / dispatch port.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_ITER 4096
#define CHAINS 8

#define KERNEL(name, INT_ASM, WITH_DFMA)                                                              \
  __global__ void name(uint64_t* out, uint32_t a, uint32_t b) {                                       \
    double acc[CHAINS]; uint64_t w[CHAINS]; uint32_t x[CHAINS];                                       \
    double m = (double)b, c = (double)a;                                                              \
    for (int i = 0; i < CHAINS; i++) { acc[i] = i + threadIdx.x; w[i] = i * 7 + threadIdx.x; x[i] = i + a + threadIdx.x; } \
    for (int it = 0; it < N_ITER; it++) {                                                             \
      _Pragma("unroll") for (int i = 0; i < CHAINS; i++) {                                            \
        if (WITH_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(m), "d"(c));    \
        INT_ASM;                                                                                      \
      }                                                                                               \
    }                                                                                                 \
    double s = 0; uint64_t t = 0;                                                                     \
    for (int i = 0; i < CHAINS; i++) { s += acc[i]; t += w[i] + x[i]; }                               \
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)s + t;                                     \
  }

#define WIDE_RZ asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(b))
#define WIDE_ACC asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(b))
#define HI32 asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b))
#define LO32 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a))
#define ADD3 asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b))
#define FFMA_ asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&x[i]) : "f"(1.0001f), "f"(0.5f))

KERNEL(k_wide_rz, WIDE_RZ, 0)
KERNEL(k_wide_acc, WIDE_ACC, 0)
KERNEL(k_hi, HI32, 0)
KERNEL(k_lo, LO32, 0)
KERNEL(k_add, ADD3, 0)
KERNEL(k_ffma, FFMA_, 0)
KERNEL(k_d_wide_rz, WIDE_RZ, 1)
KERNEL(k_d_wide_acc, WIDE_ACC, 1)
KERNEL(k_d_hi, HI32, 1)
KERNEL(k_d_lo, LO32, 1)
KERNEL(k_d_add, ADD3, 1)
KERNEL(k_d_ffma, FFMA_, 1)
KERNEL(k_lo_wide, LO32; WIDE_RZ, 0)
KERNEL(k_add_wide, ADD3; WIDE_RZ, 0)
KERNEL(k_add_wide_acc, ADD3; WIDE_ACC, 0)
KERNEL(k_ffma_wide, FFMA_; WIDE_RZ, 0)
KERNEL(k_d_only, , 1)

template <class K>
void run(const char* name, K k, uint64_t* d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int warps_per_smsp = 8, threads = 256, blocks = sms * (warps_per_smsp * 4 * 32 / threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(d, 3, 5); k<<<blocks, threads>>>(d, 3, 5);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0); k<<<blocks, threads>>>(d, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ns = best * 1e6 / ((double)warps_per_smsp * N_ITER * CHAINS);
  printf("%-24s %.2f clk per chain step (warp, SMSP) @%d MHz\n", name, ns * clk / 1e6, clk / 1000);
}

int main() {
  uint64_t* d; cudaMalloc(&d, 1 << 26);
  run("dfma", k_d_only, d);
  run("imad.wide rz", k_wide_rz, d);
  run("imad.wide acc", k_wide_acc, d);
  run("imad.hi", k_hi, d);
  run("imad.lo", k_lo, d);
  run("iadd3", k_add, d);
  run("ffma", k_ffma, d);
  run("dfma | imad.wide rz", k_d_wide_rz, d);
  run("dfma | imad.wide acc", k_d_wide_acc, d);
  run("dfma | imad.hi", k_d_hi, d);
  run("dfma | imad.lo", k_d_lo, d);
  run("dfma | iadd3", k_d_add, d);
  run("dfma | ffma", k_d_ffma, d);
  run("imad.lo | imad.wide rz", k_lo_wide, d);
  run("iadd3 | imad.wide rz", k_add_wide, d);
  run("iadd3 | imad.wide acc", k_add_wide_acc, d);
  run("ffma | imad.wide rz", k_ffma_wide, d);
  return 0;
}
