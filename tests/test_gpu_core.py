"""GPU parity tests (through the C ABI) for the commit path: Poseidon, NTT / LDE, MerkleTree::new,
PolynomialBatch::from_values / from_coeffs — bit-exact against the oracle on the same seeded inputs,
plus size-independent properties at BASELINE sizes.  Mirrors the upstream tests that would pin this
path if the crates were on disk: poseidon_goldilocks.rs test_vectors, fft.rs
fft_and_ifft / test_lde, merkle_tree.rs test_merkle_trees (+ cap-height variants), oracle.rs."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def ctx():
    import eth_tx_proof_b200 as etp

    c = etp.Context(0)
    yield c
    c.close()


def rand_cols(rng, cols, n, full_range=False):
    hi = (1 << 64) if full_range else P
    return np.array([[rng.randrange(hi) for _ in range(n)] for _ in range(cols)], dtype=np.uint64)


def np_rand(seed, shape):
    from eth_tx_proof_b200 import synthetic as syn

    x = syn._rand(seed, 0, int(np.prod(shape))).reshape(shape)
    return np.where(x >= np.uint64(P), x - np.uint64(P), x)


def test_poseidon_kats(ctx):
    with open(os.path.join(os.path.dirname(__file__), "golden", "poseidon_kat.json")) as f:
        kat = json.load(f)
    inputs = {"zeros": [0] * 12, "range12": list(range(12)), "neg_one": [P - 1] * 12}
    states = np.array([inputs[v["input"]] for v in kat["permutation"]], dtype=np.uint64)
    out = ctx.poseidon_permute(states)
    for row, v in zip(out, kat["permutation"]):
        assert [f"{int(x):016x}" for x in row] == v["output"]


def test_poseidon_random_vs_oracle(ctx):
    import oracle

    rng = random.Random(1)
    states = np.array([[rng.randrange(1 << 64) for _ in range(12)] for _ in range(1000)], dtype=np.uint64)
    states[0] = 0xFFFFFFFFFFFFFFFF  # non-canonical extremes
    states[1] = P
    states[2] = P - 1
    out = ctx.poseidon_permute(states)
    for i in range(0, 1000, 7):
        assert (out[i] == oracle.poseidon_permute(states[i])).all(), i
    assert (out < np.uint64(P)).all()


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 7, 10, 12, 13, 14, 17])
def test_fft_ifft_vs_oracle(ctx, log_n):
    import oracle

    rng = random.Random(log_n)
    cols = 3 if log_n < 14 else 2
    a = np_rand(log_n, (cols, 1 << log_n)) if log_n > 10 else rand_cols(rng, cols, 1 << log_n, full_range=True)
    f = ctx.fft(a)
    i = ctx.ifft(a)
    for c in range(cols):
        assert (f[c] == oracle.fft(a[c])).all()
        assert (i[c] == oracle.ifft(a[c])).all()
    back = ctx.ifft(f)
    assert (back == a % np.uint64(P)).all()


@pytest.mark.parametrize("log_n,rate_bits", [(0, 1), (3, 1), (5, 1), (4, 2), (4, 0), (11, 1), (12, 1), (13, 1), (16, 1), (12, 3)])
def test_coset_lde_vs_oracle(ctx, log_n, rate_bits):
    import oracle

    a = np_rand(100 + log_n, (2, 1 << log_n))
    out = ctx.coset_lde(a, rate_bits)
    for c in range(2):
        assert (out[c] == oracle.lde(a[c], rate_bits)).all()
    back = ctx.coset_ifft(ctx.coset_lde(a, 0))
    assert (back == a).all()


def test_edge_columns(ctx):
    import oracle

    n = 1 << 9
    cols = np.zeros((4, n), dtype=np.uint64)
    cols[1] = P - 1
    cols[2] = 0xFFFFFFFFFFFFFFFF  # non-canonical
    cols[3, 5] = 1                # single one
    got = ctx.ifft(cols)
    for c in range(4):
        assert (got[c] == oracle.ifft(cols[c])).all()


@pytest.mark.parametrize("log_n,cap,width", [(5, 2, 11), (4, 0, 3), (6, 4, 20), (3, 3, 9), (4, 4, 5), (2, 0, 135),
                                             (7, 1, 4), (10, 4, 8), (9, 0, 32), (3, 0, 0)])
def test_merkle_tree_new(ctx, log_n, cap, width):
    import eth_tx_proof_b200 as etp
    import oracle

    leaves = np_rand(log_n * 31 + width, (1 << log_n, width))
    t = etp.MerkleTree.new(ctx, leaves, cap)
    digests, capv = oracle.merkle_new(leaves, cap)
    assert (t.cap == capv).all()
    assert (t.digests == digests).all()
    for i in {0, 1, (1 << log_n) - 1, (1 << log_n) // 3}:
        sib = t.prove(i)
        assert (sib == oracle.merkle_prove(digests, 1 << log_n, cap, i)).all()
        assert oracle.merkle_verify(leaves[i], i, sib, capv)


def test_merkle_cap_height_too_large_is_rejected(ctx):
    import eth_tx_proof_b200 as etp

    with pytest.raises(etp.EtpError):
        etp.MerkleTree.new(ctx, np.zeros((8, 3), dtype=np.uint64), 4)
    with pytest.raises(etp.EtpError):
        etp.MerkleTree.new(ctx, np.zeros((6, 3), dtype=np.uint64), 1)  # not a power of two


@pytest.mark.parametrize("log_n,cols,rate_bits,cap", [(4, 3, 1, 2), (5, 9, 1, 4), (3, 4, 1, 0), (4, 2, 2, 1), (4, 1, 1, 4),
                                                     (10, 21, 1, 4), (12, 16, 1, 4), (13, 7, 1, 4), (14, 130, 1, 4), (3, 12, 1, 4)])
def test_batch_from_values_vs_oracle(ctx, log_n, cols, rate_bits, cap):
    import eth_tx_proof_b200 as etp
    import oracle

    vals = np_rand(log_n * 1000 + cols, (cols, 1 << log_n))
    if log_n == 10:
        vals[0] = 0xFFFFFFFFFFFFFFFF
        vals[1] = 0
        vals[2] = P - 1
    b = etp.PolynomialBatch.from_values(ctx, vals, rate_bits, False, cap)
    o = oracle.Batch.from_values(vals, rate_bits, cap)
    assert (b.cap == o.cap).all()
    assert (b.polynomials == o.coeffs).all()
    assert (b.leaves == o.leaves).all()
    assert (b.digests == o.digests).all()
    big = 1 << (log_n + rate_bits)
    idx = [0, 1, big - 1, big // 2 + 1]
    rows = b.leaves_at(idx)
    for q, i in enumerate(idx):
        assert (rows[q] == o.leaves[i]).all()
        assert oracle.merkle_verify(rows[q], i, b.prove(i), b.cap)
    # get_lde_values(index, step) == leaves[bitrev(index*step)]
    from oracle import pyref as R

    assert (b.get_lde_values(3, 1) == o.leaves[R.bitrev(3, log_n + rate_bits)]).all()
    b2 = etp.PolynomialBatch.from_coeffs(ctx, o.coeffs, rate_bits, False, cap)
    assert (b2.cap == o.cap).all() and (b2.digests == o.digests).all()


def test_blinding_is_rejected(ctx):
    import eth_tx_proof_b200 as etp

    with pytest.raises(etp.EtpError):
        etp.PolynomialBatch.from_values(ctx, np.zeros((2, 8), dtype=np.uint64), 1, True, 0)


def test_quotient_shaped_batch_leaves_are_not_hashed(ctx):
    """4 columns => hash_or_noop copies the row; cap of a 2^cap-leaf tree is the rows themselves."""
    import eth_tx_proof_b200 as etp

    c = np_rand(77, (4, 8))
    b = etp.PolynomialBatch.from_coeffs(ctx, c, 1, False, 4)  # 16 leaves, cap 16
    assert (b.cap == b.leaves).all()
    assert b.digests.shape[0] == 0


def test_from_values_dev_and_properties_at_baseline_size(ctx):
    """BASELINE config 2 shape (2^20 x 128 would take the oracle minutes): check size-independent
    properties instead — linearity of the committed LDE, Merkle paths verify to the cap, and the
    leaf digests equal a row-major MerkleTree::new over downloaded rows on a sampled subtree."""
    import torch

    import eth_tx_proof_b200 as etp
    import oracle

    log_n, cols = 18, 128
    n = 1 << log_n
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda", generator=g)
    b_ = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda", generator=g)
    s = a + b_  # < 2^63 < p: plain integer sum == field sum
    torch.cuda.synchronize()  # the library works on its own (non-blocking) stream
    ba = etp.PolynomialBatch.from_values_dev(ctx, a.data_ptr(), n, cols, log_n, 1, False, 4)
    bb = etp.PolynomialBatch.from_values_dev(ctx, b_.data_ptr(), n, cols, log_n, 1, False, 4)
    bs = etp.PolynomialBatch.from_values_dev(ctx, s.data_ptr(), n, cols, log_n, 1, False, 4)
    idx = [0, 1, 12345, (2 << log_n) - 1, 1 << log_n]
    ra, rb, rs = ba.leaves_at(idx).astype(object), bb.leaves_at(idx).astype(object), bs.leaves_at(idx).astype(object)
    assert ((ra + rb) % P == rs).all()
    for q, i in enumerate(idx):
        assert oracle.merkle_verify(bs.leaves_at([i])[0], i, bs.prove(i), bs.cap)
    # oracle on a few columns of the same data
    host = a[:3].cpu().numpy().astype(np.uint64)
    o = oracle.Batch.from_values(host, 1, 4)
    assert (ba.polynomials[:3] == o.coeffs).all()
    assert (ba.leaves_at(idx)[:, :3] == o.leaves[idx]).all()


def test_block_cache_reuses_and_trims(ctx):
    """dev_alloc / dev_free go through the context's block cache: a second commit of the same shape allocates nothing new,
    etp_ctx_trim gives the cached blocks back, and results do not depend on whether a block is fresh or reused."""
    import eth_tx_proof_b200 as etp
    from eth_tx_proof_b200 import synthetic as syn

    vals = syn.random_columns(9, 11, seed=5)
    b1 = etp.PolynomialBatch.from_values(ctx, vals, 1, False, 4)
    cap1 = b1.cap.copy()
    del b1
    held = ctx.cached_bytes
    assert held >= 9 * (1 << 11) * 8 * 3  # coefficients + LDE of the freed batch are cached
    b2 = etp.PolynomialBatch.from_values(ctx, vals, 1, False, 4)
    assert ctx.cached_bytes < held  # served from the cache
    assert (b2.cap == cap1).all()
    del b2
    ctx.trim()
    assert ctx.cached_bytes == 0
    b3 = etp.PolynomialBatch.from_values(ctx, vals, 1, False, 4)
    assert (b3.cap == cap1).all()


def test_properties_at_the_full_baseline_size(ctx):
    """BASELINE.json configs[1] at its full size, 2^22 x 128, rate_bits 1, cap_height 4 (the oracle would need minutes for the
    whole batch): size-independent properties — the commitment is linear on sampled leaf rows, sampled Merkle paths verify
    to the cap, the cap is reproducible, and three columns agree with the oracle's ifft / coset LDE word for word."""
    import torch

    import eth_tx_proof_b200 as etp
    import oracle

    log_n, cols = 22, 128
    n = 1 << log_n
    g = torch.Generator(device="cuda").manual_seed(22)
    a = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda", generator=g)
    b_ = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda", generator=g)
    idx = [0, 1, 2, 987654, n - 1, n, n + 1, (2 << log_n) - 1]
    torch.cuda.synchronize()  # the library works on its own (non-blocking) stream
    ba = etp.PolynomialBatch.from_values_dev(ctx, a.data_ptr(), n, cols, log_n, 1, False, 4)
    ra, cap_a = ba.leaves_at(idx).astype(object), ba.cap.copy()
    # three columns against the oracle: coefficients and the whole LDE column in leaf (bit-reversed) order
    host = a[[0, 63, 127]].cpu().numpy().astype(np.uint64)
    want_coeffs = np.stack([oracle.ifft(c) for c in host])
    for k, c in enumerate([0, 63, 127]):
        lde = oracle.lde(want_coeffs[k], 1)  # natural order: value k = P(7 w^k)
        for q, i in enumerate(idx):
            assert int(ra[q][c]) == int(lde[int(format(i, f"0{log_n + 1}b")[::-1], 2)])
    for i in idx[:4]:
        assert oracle.merkle_verify(ba.leaves_at([i])[0], i, ba.prove(i), ba.cap)
    ba.recommit_values_dev(a.data_ptr(), n)
    assert (ba.cap == cap_a).all()
    del ba
    bb = etp.PolynomialBatch.from_values_dev(ctx, b_.data_ptr(), n, cols, log_n, 1, False, 4)
    rb = bb.leaves_at(idx).astype(object)
    del bb
    s = a + b_  # < 2^63 < p: plain integer sum == field sum
    torch.cuda.synchronize()
    del a, b_
    bs = etp.PolynomialBatch.from_values_dev(ctx, s.data_ptr(), n, cols, log_n, 1, False, 4)
    rs = bs.leaves_at(idx).astype(object)
    assert ((ra + rb) % P == rs).all()
    assert oracle.merkle_verify(bs.leaves_at([idx[-1]])[0], idx[-1], bs.prove(idx[-1]), bs.cap)
    del bs
    ctx.trim()
