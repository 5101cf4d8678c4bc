/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * MerkleTree::new / prove / verify_merkle_proof_to_cap.  Restates plonky2 0.2.2
 * (/root/reference/Cargo.lock:3441; not on disk):
 *   plonky2/src/hash/merkle_tree.rs    MerkleTree::new, fill_digests_buf, fill_subtree, prove
 *   plonky2/src/hash/merkle_proofs.rs  verify_merkle_proof_to_cap
 * Digest layout per cap subtree (recursive): left recursive output || left child digest ||
 * right child digest || right recursive output.
 */
#include "oracle.h"
#include <string.h>

static int log2_strict(size_t n) {
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

/* fill_subtree: digests_buf holds 2*(n_leaves-1) digests; returns the subtree root in out */
static void fill_subtree(uint64_t *digests_buf, size_t n_digests, const uint64_t *leaves, size_t n_leaves,
                         size_t leaf_len, uint64_t out[4]) {
  if (n_digests == 0) {
    orc_hash_or_noop(leaves, leaf_len, out);
    return;
  }
  size_t half = n_digests / 2;
  uint64_t *left_buf = digests_buf;                       /* half - 1 digests             */
  uint64_t *left_digest_mem = digests_buf + (half - 1) * 4;
  uint64_t *right_digest_mem = digests_buf + half * 4;
  uint64_t *right_buf = digests_buf + (half + 1) * 4;     /* half - 1 digests             */
  uint64_t l[4], r[4];
#pragma omp task shared(l) if (n_leaves >= 512)
  fill_subtree(left_buf, half - 1, leaves, n_leaves / 2, leaf_len, l);
#pragma omp task shared(r) if (n_leaves >= 512)
  fill_subtree(right_buf, half - 1, leaves + (n_leaves / 2) * leaf_len, n_leaves / 2, leaf_len, r);
#pragma omp taskwait
  memcpy(left_digest_mem, l, 32);
  memcpy(right_digest_mem, r, 32);
  orc_two_to_one(l, r, out);
}

void orc_merkle_new(const uint64_t *leaves, size_t n_leaves, size_t leaf_len, int cap_height,
                    uint64_t *digests, uint64_t *cap) {
  size_t n_cap = (size_t)1 << cap_height;
  size_t num_digests = 2 * (n_leaves - n_cap);
  uint64_t warm[12] = {0};
  orc_poseidon_permute(warm); /* force round-constant init before going parallel */
  if (num_digests == 0) {
    for (size_t i = 0; i < n_leaves; i++) orc_hash_or_noop(leaves + i * leaf_len, leaf_len, cap + 4 * i);
    return;
  }
  size_t sub_digests = num_digests >> cap_height, sub_leaves = n_leaves >> cap_height;
#pragma omp parallel
#pragma omp single
  for (size_t s = 0; s < n_cap; s++) {
#pragma omp task
    fill_subtree(digests + s * sub_digests * 4, sub_digests, leaves + s * sub_leaves * leaf_len, sub_leaves,
                 leaf_len, cap + 4 * s);
  }
}

void orc_merkle_prove(const uint64_t *digests, size_t n_leaves, int cap_height, size_t leaf_index,
                      uint64_t *siblings) {
  int num_layers = log2_strict(n_leaves) - cap_height;
  size_t subtree_digest_size = ((size_t)1 << (num_layers + 1)) - 2;
  size_t subtree_idx = leaf_index >> num_layers;
  const uint64_t *sub = digests + subtree_idx * subtree_digest_size * 4;
  size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
  for (int i = 0; i < num_layers; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    size_t sibling_index = 2 * siblings_index + (1 - parity);
    memcpy(siblings + 4 * i, sub + 4 * sibling_index, 32);
  }
}

int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index, const uint64_t *siblings,
                      int n_siblings, const uint64_t *cap, int cap_height) {
  (void)cap_height;
  uint64_t cur[4], nxt[4];
  size_t index = leaf_index;
  orc_hash_or_noop(leaf, leaf_len, cur);
  for (int i = 0; i < n_siblings; i++) {
    if (index & 1) orc_two_to_one(siblings + 4 * i, cur, nxt);
    else orc_two_to_one(cur, siblings + 4 * i, nxt);
    memcpy(cur, nxt, 32);
    index >>= 1;
  }
  for (int i = 0; i < 4; i++)
    if (cur[i] != gl_canon(cap[4 * index + i])) return 0;
  return 1;
}
