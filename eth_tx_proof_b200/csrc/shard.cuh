// The column-split table of etp_shard.cu, shared with etp_stark.cu (quotient / openings / FRI over a split table).
#pragma once
#include "ctx.cuh"

struct etp_shard {
  etp_ctx* ctx;
  size_t n_cols_total, cps, c0, local_cols;
  int log_n, rate_bits, cap_height, rank, world;
  uint64_t* coeffs = nullptr;   // local_cols x n      (cudaMalloc: exportable)
  uint64_t* lde = nullptr;      // local_cols x L      (cudaMalloc: exportable)
  uint64_t* levels = nullptr;   // digests of the own rows, level by level, down to the own cap entries
  const uint64_t* peer[merkle::MAX_SRC] = {};
  bool committed = false;
  size_t n() const { return (size_t)1 << log_n; }
  size_t lde_n() const { return (size_t)1 << (log_n + rate_bits); }
  size_t rows() const { return lde_n() / world; }
  size_t row0() const { return rows() * rank; }
  int local_cap_height() const { int lw = 0; while ((1 << lw) < world) lw++; return cap_height - lw; }
  // LDE column c of the whole table as addressable from this rank (own HBM or a mapped peer), nullptr if not mapped
  const uint64_t* column(size_t c) const {
    const size_t g = c / cps;
    return c < n_cols_total && peer[g] ? peer[g] + (c - g * cps) * lde_n() : nullptr;
  }
};
