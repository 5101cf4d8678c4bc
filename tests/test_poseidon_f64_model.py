"""The FP64-pipe schedule of csrc/poseidon.cuh (three-level circulant split, partial rounds two at a time
through C*C), replayed on Python integers by tools/gen_poseidon_constants.model_permute: it must equal the
oracle's plain permutation and the upstream known-answer vectors, and no modelled double may leave the
exactly-representable range."""
import json
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_poseidon_constants as gen  # noqa: E402

from oracle import pyref  # noqa: E402


def test_model_matches_known_answer_vectors():
    with open(os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")) as f:
        kat = json.load(f)
    inputs = {"zeros": [0] * 12, "range12": list(range(12)), "neg_one": [gen.P - 1] * 12}
    assert len(kat["permutation"]) == 3
    for v in kat["permutation"]:
        st = inputs[v["input"]] if v["input"] in inputs else [gen.P - 1] * 12
        assert gen.model_permute(st) == [int(x, 16) for x in v["output"]], v["input"]


def test_model_matches_oracle_and_stays_exact():
    rng = random.Random(7)
    tr = gen.Track()
    edge = [[0] * 12, [gen.P - 1] * 12, [2**64 - 1] * 12, [2**32 - 1] * 12, [2**32] * 12, [0xFFFFFFFF00000000] * 12]
    for st in edge + [[rng.randrange(2**64) for _ in range(12)] for _ in range(40)]:
        assert gen.model_permute(st, tr) == pyref.poseidon([x % gen.P for x in st])
    assert tr.max < 2**53
    assert gen.worst_case_bound() < 2**53, math.log2(gen.worst_case_bound())


def test_split_coefficients_are_integral():
    c1 = gen.split_coeffs(gen.CIRC)
    c2 = gen.split_coeffs(gen.circ_square(gen.CIRC))
    assert c1 == ([2, 1, 1, -1, -16, 4], [-1, -2, 8], [16, 16, 32])
    assert c2[2].count(c2[2][0]) == 2  # cyclic(3) part keeps the "sum + one term" shape
    # round-constant planes: congruent to the constant, multiples of 4, non-negative
    for k in gen.read_rc():
        lo, hi = gen.mod4_planes(k)
        assert lo % 4 == 0 and hi % 4 == 0 and lo >= 0 and hi >= 0 and (lo + (hi << 32)) % gen.P == k
