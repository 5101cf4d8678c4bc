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

#include "../../include/etp_b200.h"
#include "gl.cuh"
#include "poseidon_constants.h"

namespace hostf {

}  // namespace hostf
// host_poseidon.cpp (plain C++, AVX2 when the CPU has it): in place, any u64 in, canonical out
extern "C" void etp_host_poseidon_permute(uint64_t s[12]);
namespace hostf {

inline void poseidon(uint64_t s[12]) { etp_host_poseidon_permute(s); }

// The transcript state IS the public plain-data struct etp_challenger (include/etp_b200.h): callers hand it in and get
// it back (challenger state in / out), exactly the fields of plonky2's Challenger.
struct Challenger : etp_challenger {
  Challenger() { memset(static_cast<etp_challenger*>(this), 0, sizeof(etp_challenger)); }
  explicit Challenger(const etp_challenger& c) : etp_challenger(c) {}
  void duplexing() {
    for (uint32_t i = 0; i < input_len; i++) sponge_state[i] = input_buffer[i];
    input_len = 0;
    poseidon(sponge_state);
    memcpy(output_buffer, sponge_state, sizeof output_buffer);
    output_len = 8;
  }
  void observe(uint64_t e) {
    output_len = 0;
    input_buffer[input_len++] = gl::canon(e);
    if (input_len == 8) duplexing();
  }
  void observe(const uint64_t* e, size_t n) { for (size_t i = 0; i < n; i++) observe(e[i]); }
  uint64_t get() {
    if (input_len != 0 || output_len == 0) duplexing();
    return output_buffer[--output_len];
  }
  gl::Ext get_ext() { uint64_t a = get(); uint64_t b = get(); return gl::ext(a, b); }
  // Challenger::compact: flush pending inputs, drop buffered outputs; the sponge state is what a recursion circuit resumes from
  void compact() {
    if (input_len != 0) duplexing();
    output_len = 0;
  }
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
