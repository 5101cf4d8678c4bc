"""CPU tests: pin the oracle (C restatement) against the upstream known-answer vectors and against
the independent pure-Python restatement (oracle/pyref.py), and exercise the edge cases upstream tests
(plonky2 merkle_tree.rs test_merkle_trees / cap-height variants, fft.rs fft_and_ifft / test_lde,
poseidon_goldilocks.rs test_vectors; none of which is on disk — SURVEY.md section 4)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import oracle
from oracle import pyref as R

P = R.P
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _kat():
    with open(os.path.join(GOLD, "poseidon_kat.json")) as f:
        return json.load(f)


def test_field_constants():
    assert R.root_of_unity(32) == int(_kat()["power_of_two_generator"], 16)
    assert pow(7, (P - 1) >> 32, P) == R.POWER_OF_TWO_GENERATOR
    assert pow(R.POWER_OF_TWO_GENERATOR, 1 << 31, P) == P - 1  # order exactly 2^32
    assert (1 << 64) % P == (1 << 32) - 1 and (1 << 96) % P == P - 1


def test_round_constants_pinned():
    k = _kat()
    rc = oracle.poseidon_constants()
    assert hashlib.sha256(rc.astype("<u8").tobytes()).hexdigest() == k["round_constants_sha256_le_u64"]
    assert [f"{int(x):016x}" for x in rc[:4]] == k["round_constants_head"]
    assert [f"{int(x):016x}" for x in rc[-4:]] == k["round_constants_tail"]
    assert [int(x) for x in rc] == R.round_constants()  # C and Python derivations agree


def test_permutation_kats():
    inputs = {"zeros": [0] * 12, "range12": list(range(12)), "neg_one": [P - 1] * 12}
    for v in _kat()["permutation"]:
        want = [int(x, 16) for x in v["output"]]
        assert [int(x) for x in oracle.poseidon_permute(inputs[v["input"]])] == want
        assert R.poseidon(inputs[v["input"]]) == want


def test_permutation_c_vs_python_random_and_noncanonical():
    rng = random.Random(1)
    for _ in range(20):
        s = [rng.randrange(1 << 64) for _ in range(12)]  # includes values >= p
        assert [int(x) for x in oracle.poseidon_permute(s)] == R.poseidon(s)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 8, 9, 16, 17, 21, 128])
def test_sponge(n):
    rng = random.Random(n)
    x = [rng.randrange(P) for _ in range(n)]
    assert [int(v) for v in oracle.hash_no_pad(x)] == R.hash_no_pad(x)
    assert [int(v) for v in oracle.hash_or_noop(x)] == R.hash_or_noop(x)
    if n <= 4:  # hash_or_noop copies short inputs
        assert R.hash_or_noop(x) == x + [0] * (4 - n)


def test_two_to_one():
    rng = random.Random(5)
    l, r = [rng.randrange(P) for _ in range(4)], [rng.randrange(P) for _ in range(4)]
    assert [int(v) for v in oracle.two_to_one(l, r)] == R.two_to_one(l, r) == R.poseidon(l + r + [0] * 4)[:4]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_fft_matches_definition(log_n):
    rng = random.Random(log_n)
    a = [rng.randrange(1 << 64) for _ in range(1 << log_n)]
    assert [int(v) for v in oracle.fft(a)] == R.dft(a)
    assert [int(v) for v in oracle.ifft(a)] == R.idft(a)
    assert [int(v) for v in oracle.ifft(oracle.fft(a))] == [v % P for v in a]


@pytest.mark.parametrize("log_n,rate_bits", [(0, 1), (3, 1), (5, 1), (4, 2), (4, 0)])
def test_lde_matches_definition(log_n, rate_bits):
    rng = random.Random(log_n * 7 + rate_bits)
    c = [rng.randrange(P) for _ in range(1 << log_n)]
    assert [int(v) for v in oracle.lde(c, rate_bits)] == R.lde_values(c, rate_bits)
    back = oracle.coset_ifft(oracle.coset_fft(c))
    assert [int(v) for v in back] == c


@pytest.mark.parametrize("log_n,cap,width", [(5, 2, 11), (4, 0, 3), (6, 4, 20), (3, 3, 9), (4, 4, 5), (2, 0, 135), (7, 1, 4)])
def test_merkle_layout_and_proofs(log_n, cap, width):
    rng = random.Random(log_n * 100 + cap)
    n = 1 << log_n
    leaves = [[rng.randrange(P) for _ in range(width)] for _ in range(n)]
    digests, capv = oracle.merkle_new(np.array(leaves, dtype=np.uint64), cap)
    pd, pc = R.merkle_tree(leaves, cap)
    assert [[int(x) for x in d] for d in digests] == pd
    assert [[int(x) for x in d] for d in capv] == pc
    assert digests.shape[0] == 2 * (n - (1 << cap))
    for i in range(n):
        sib = oracle.merkle_prove(digests, n, cap, i)
        assert sib.shape[0] == log_n - cap
        assert oracle.merkle_verify(leaves[i], i, sib, capv)
        assert R.merkle_verify(leaves[i], i, [[int(x) for x in s] for s in sib], pc)
    if log_n > cap:  # a wrong leaf must not verify
        bad = list(leaves[0])
        bad[0] = (bad[0] + 1) % P
        assert not oracle.merkle_verify(bad, 0, oracle.merkle_prove(digests, n, cap, 0), capv)


def test_merkle_closed_form_scatter_map():
    """node j of layer i sits at 2*(((j>>1)<<(i+1)) + 2^i - 1) + (j&1) in its subtree."""
    rng = random.Random(9)
    log_n, cap, width = 6, 1, 7
    n = 1 << log_n
    leaves = [[rng.randrange(P) for _ in range(width)] for _ in range(n)]
    digests, _ = oracle.merkle_new(np.array(leaves, dtype=np.uint64), cap)
    level = [R.hash_or_noop(l) for l in leaves]
    num_layers = log_n - cap
    sub_size = (1 << (num_layers + 1)) - 2
    for i in range(num_layers):
        per_sub = len(level) >> cap
        for j, d in enumerate(level):
            s, jj = divmod(j, per_sub)
            slot = s * sub_size + 2 * (((jj >> 1) << (i + 1)) + (1 << i) - 1) + (jj & 1)
            assert [int(x) for x in digests[slot]] == d
        level = [R.two_to_one(level[2 * k], level[2 * k + 1]) for k in range(len(level) // 2)]


@pytest.mark.parametrize("log_n,cols,rate_bits,cap", [(4, 3, 1, 2), (5, 9, 1, 4), (3, 4, 1, 0), (4, 2, 2, 1), (4, 1, 1, 4)])
def test_batch_from_values(log_n, cols, rate_bits, cap):
    rng = random.Random(log_n + cols)
    n = 1 << log_n
    vals = [[rng.randrange(1 << 64) for _ in range(n)] for _ in range(cols)]
    b = oracle.Batch.from_values(np.array(vals, dtype=np.uint64), rate_bits, cap)
    coeffs = [R.idft(v) for v in vals]
    assert [[int(x) for x in c] for c in b.coeffs] == coeffs
    ldes = [R.lde_values(c, rate_bits) for c in coeffs]
    big_log = log_n + rate_bits
    rows = [[ldes[c][R.bitrev(i, big_log)] for c in range(cols)] for i in range(n << rate_bits)]
    assert [[int(x) for x in r] for r in b.leaves] == rows
    pd, pc = R.merkle_tree(rows, cap)
    assert [[int(x) for x in d] for d in b.digests] == pd
    assert [[int(x) for x in d] for d in b.cap] == pc
    b2 = oracle.Batch.from_coeffs(np.array(coeffs, dtype=np.uint64), rate_bits, cap)
    assert (b2.cap == b.cap).all() and (b2.leaves == b.leaves).all()


def test_challenger_c_vs_python():
    import ctypes as C

    L = oracle.lib()
    ch = oracle.Challenger()
    L.orc_challenger_init(C.byref(ch))
    py = R.Challenger()
    rng = random.Random(3)
    for step in range(40):
        if rng.random() < 0.6:
            xs = [rng.randrange(1 << 64) for _ in range(rng.randrange(1, 13))]
            arr = np.array(xs, dtype=np.uint64)
            L.orc_challenger_observe(C.byref(ch), arr.ctypes.data_as(C.POINTER(C.c_uint64)), len(xs))
            py.observe(xs)
        else:
            for _ in range(rng.randrange(1, 11)):
                assert int(L.orc_challenger_get(C.byref(ch))) == py.get()


def test_pow_grind_is_smallest():
    rng = random.Random(11)
    st = [rng.randrange(P) for _ in range(12)]
    w = oracle.pow_grind(st, 3, 8)
    for cand in range(w + 1):
        s = list(st)
        s[3] = cand
        ok = (R.poseidon(s)[7] >> 56) == 0
        assert ok == (cand == w)


def test_fri_fold_matches_definition():
    rng = random.Random(13)
    n, ab = 64, 4
    coeffs = [(rng.randrange(P), rng.randrange(P)) for _ in range(n)]
    beta = (rng.randrange(P), rng.randrange(P))
    got = oracle.fri_fold_coeffs(np.array(coeffs, dtype=np.uint64), ab, beta)
    for k in range(n >> ab):
        acc = (0, 0)
        for j in range(1 << ab):
            acc = R.e_add(acc, R.e_mul(coeffs[(k << ab) + j], R.e_pow(beta, j)))
        assert (int(got[k][0]), int(got[k][1])) == acc


def test_fast_permutation_equals_the_naive_definition():
    """orc_poseidon_permute (lazy reduction, split MDS: the form the timed CPU baseline runs) against the
    round-by-round definition, on edge states and random ones."""
    import random

    import oracle

    rng = random.Random(11)
    p = 0xFFFFFFFF00000001
    edge = [[0] * 12, [p - 1] * 12, [2**64 - 1] * 12, [2**32 - 1] * 12, [2**32] * 12, [p] * 12, [0xFFFFFFFF00000000] * 12]
    for st in edge + [[rng.randrange(2**64) for _ in range(12)] for _ in range(500)]:
        a = oracle.poseidon_permute(np.array(st, dtype=np.uint64))
        b = oracle.poseidon_permute_naive(np.array(st, dtype=np.uint64))
        assert (a == b).all() and (a < np.uint64(p)).all()
