// Do the S-box (integer pipes) and the MDS layer (FP64 pipe) of the Poseidon kernel overlap on B200, or do
// their times add?  Times the full permutation, the S-boxes alone and the MDS layers alone (same loop
// structure, 128-thread CTAs, grid sized for 7 CTAs/SM x 148 SMs x 8 waves).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../eth_tx_proof_b200/csrc/poseidon.cuh"

template <int MODE>  // 0 full, 1 S-boxes only, 2 MDS only
__global__ void __launch_bounds__(128, 7) k_parts(uint64_t* st, int reps) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  uint64_t s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = st[i] + t * (i + 1);
  for (int rep = 0; rep < reps; rep++) {
    if (MODE == 0) {
      poseidon::permute(s);
    } else if (MODE == 1) {
      // 118 S-boxes in the same shape: 8 x 12 through the out-of-line pair + 22 single ones
#pragma unroll 1
      for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int k = 0; k < 12; k += 2) {
          const ulonglong2 q = poseidon::sbox7_pair(s[k], s[k + 1]);
          s[k] = q.x; s[k + 1] = q.y;
        }
      }
#pragma unroll 1
      for (int r = 0; r < 22; r++) s[r % 12 == 0 ? 0 : 0] = poseidon::sbox7(s[0] + r);
    } else {
      // the FP64 part in the same shape as poseidon::permute: 8 full layers + 11 partial-round pairs, no S-boxes
#pragma unroll 1
      for (int f = 0; f < 8; f++) {
        double dl[12], dh[12], ol[12], oh[12];
#pragma unroll
        for (int k = 0; k < 12; k++) { dl[k] = poseidon::plane_lo(s[k]); dh[k] = poseidon::plane_hi(s[k]); }
        poseidon::mds_plane(dl, poseidon::FULL_F64 + 24 * f, ol);
        poseidon::mds_plane(dh, poseidon::FULL_F64 + 24 * f + 12, oh);
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = poseidon::combine_planes(ol[k], oh[k]);
      }
#pragma unroll 1
      for (int i = 0; i < 11; i++) {
        double dl[12], dh[12], ol[12], oh[12];
#pragma unroll
        for (int k = 0; k < 12; k++) { dl[k] = poseidon::plane_lo(s[k]); dh[k] = poseidon::plane_hi(s[k]); }
        poseidon::Chains cl, ch;
        const double ul = poseidon::pair_phase1(dl, poseidon::PAIR_F64 + 26 * i, cl);
        const double uh = poseidon::pair_phase1(dh, poseidon::PAIR_F64 + 26 * i + 13, ch);
        const uint64_t m0 = poseidon::combine_planes(ul, uh);
        poseidon::pair_phase2(dl[0], ul, poseidon::plane_lo(m0), cl, ol);
        poseidon::pair_phase2(dh[0], uh, poseidon::plane_hi(m0), ch, oh);
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = poseidon::combine_planes(ol[k], oh[k]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 12; i++) st[12 * t + i] = s[i];
}

template <int MODE>
void run(const char* name, uint64_t* d, int blocks_per_sm) {
  const int reps = 16, blocks = 148 * blocks_per_sm * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_parts<MODE><<<blocks, 128>>>(d, reps);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int i = 0; i < 3; i++) {
    cudaEventRecord(e0); k_parts<MODE><<<blocks, 128>>>(d, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double perms = (double)blocks * 128 * reps;
  printf("%-12s %d CTA/SM-equivalent grid: %8.3f ms  %.3f G perm-equivalents/s  %.0f clk per warp-perm per SMSP @1965 MHz\n", name, blocks_per_sm,
         best, perms / best / 1e6, best * 1e-3 * 1.965e9 * 148 * 4 * 32 / perms);
}

int main() {
  uint64_t* d; cudaMalloc(&d, (size_t)148 * 7 * 4 * 128 * 96 + 4096);
  cudaMemset(d, 1, 4096);
  run<0>("full", d, 7);
  run<1>("sbox only", d, 7);
  run<2>("mds only", d, 7);
  return 0;
}
