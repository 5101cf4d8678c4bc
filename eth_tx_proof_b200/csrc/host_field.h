// Host-side helpers of the product library: Poseidon permutation + Fiat-Shamir Challenger for the
// (tiny, strictly sequential) transcript, and small host transforms for the FRI final polynomial.
//
// Replaces plonky2 0.2.2 `Challenger<GoldilocksField, PoseidonHash>` (plonky2/src/iop/challenger.rs):
// duplex sponge in overwrite mode, rate 8, challenges popped from the END of the output buffer; and
// `PoseidonPermutation::permute` (plonky2/src/hash/poseidon.rs) — crate pinned at
// /root/reference/Cargo.lock:3441, reached from /root/reference/ops/src/lib.rs:52.
// The transcript is strictly sequential, so it stays on a host core (host_poseidon.cpp: ~3 us per permutation;
// a 2400-column table makes the challenger absorb ~9600 opening words = 1200 permutations);
// each FRI layer costs one cap download (512 B) and nothing is uploaded but the 16-byte beta.
// This is NOT the oracle: nothing here includes oracle/.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "gl.cuh"
#include "poseidon_constants.h"

namespace hostf {

}  // namespace hostf
// host_poseidon.cpp (plain C++, AVX2 when the CPU has it): in place, any u64 in, canonical out
extern "C" void etp_host_poseidon_permute(uint64_t s[12]);
namespace hostf {

inline void poseidon(uint64_t s[12]) { etp_host_poseidon_permute(s); }

struct Challenger {
  uint64_t state[12] = {0};
  uint64_t in[8];
  int n_in = 0;
  uint64_t out[8];
  int n_out = 0;
  void duplexing() {
    for (int i = 0; i < n_in; i++) state[i] = in[i];
    n_in = 0;
    poseidon(state);
    memcpy(out, state, sizeof out);
    n_out = 8;
  }
  void observe(uint64_t e) {
    n_out = 0;
    in[n_in++] = gl::canon(e);
    if (n_in == 8) duplexing();
  }
  void observe(const uint64_t* e, size_t n) { for (size_t i = 0; i < n; i++) observe(e[i]); }
  uint64_t get() {
    if (n_in != 0 || n_out == 0) duplexing();
    return out[--n_out];
  }
  gl::Ext get_ext() { uint64_t a = get(); uint64_t b = get(); return gl::ext(a, b); }
};

// in-place radix-2 DFT on ext values (natural order in/out), root = w or w^-1
inline void ext_fft(std::vector<gl::Ext>& a, int log_n, bool inverse) {
  const size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = gl::bitrev32((uint32_t)i, log_n);
    if (i < j) std::swap(a[i], a[j]);
  }
  uint64_t w_n = gl::root_of_unity(log_n);
  if (inverse) w_n = gl::canon(gl::inv(w_n));
  for (int s = 1; s <= log_n; s++) {
    const size_t m = (size_t)1 << s, half = m >> 1;
    const uint64_t wm = gl::canon(gl::pow(w_n, n >> s));
    for (size_t k = 0; k < n; k += m) {
      uint64_t w = 1;
      for (size_t j = 0; j < half; j++) {
        gl::Ext u = a[k + j], v = gl::emul_base(a[k + j + half], w);
        a[k + j] = gl::eadd(u, v);
        a[k + j + half] = gl::esub(u, v);
        w = gl::mul(w, wm);
      }
    }
  }
  if (inverse) {
    const uint64_t n_inv = gl::inv((uint64_t)n);
    for (auto& x : a) x = gl::emul_base(x, n_inv);
  }
  for (auto& x : a) x = gl::ecanon(x);
}

}  // namespace hostf
