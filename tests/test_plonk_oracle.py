"""The permutation argument of plonky2's circuit prover as restated by the oracle (oracle/plonk.c): on a witness that satisfies
its copy constraints the running product Z closes, on a broken one it does not; layout = [Z per challenge] ++ partial products."""
import numpy as np

import oracle
from eth_tx_proof_b200 import recursion as rec

P = rec.P


def closing_value(wires, sigmas, k_is, beta, gamma, z_last):
    n = wires.shape[1]
    x = pow(rec.root_of_unity(int(n).bit_length() - 1), n - 1, P)
    r = z_last
    for j in range(wires.shape[0]):
        num = (int(wires[j, n - 1]) + beta * int(k_is[j]) * x + gamma) % P
        den = (int(wires[j, n - 1]) + beta * int(sigmas[j, n - 1]) + gamma) % P
        r = r * num * pow(den, P - 2, P) % P
    return r


def test_partial_products_and_zs_close_on_a_consistent_witness():
    for degree_bits, routed, factor in ((4, 10, 4), (6, 80, 8), (5, 7, 8)):
        w, s, k = rec.permutation_witness(degree_bits, routed, seed=degree_bits)
        betas, gammas = [11, 22], [33, 44]
        out = oracle.plonk_partial_products_and_zs(w, s, k, factor, betas, gammas)
        n_pp = -(-routed // factor) - 1
        assert out.shape == (2 * (1 + n_pp), 1 << degree_bits)
        for c in range(2):
            assert int(out[c, 0]) == 1
            assert closing_value(w, s, k, betas[c], gammas[c], int(out[c, -1])) == 1
        # partial product c of row i = Z(x_i) * (first c+1 chunk products): the last one times the last chunk = Z(x_{i+1})
        if n_pp:
            i = 3
            acc = int(out[0, i])
            for cidx in range(n_pp + 1):
                q = 1
                x = pow(rec.root_of_unity(degree_bits), i, P)
                for j in range(cidx * factor, min((cidx + 1) * factor, routed)):
                    num = (int(w[j, i]) + betas[0] * int(k[j]) * x + gammas[0]) % P
                    den = (int(w[j, i]) + betas[0] * int(s[j, i]) + gammas[0]) % P
                    q = q * num * pow(den, P - 2, P) % P
                acc = acc * q % P
                if cidx < n_pp:
                    assert int(out[2 + cidx, i]) == acc
            assert int(out[0, i + 1]) == acc
        # break one copy constraint: Z no longer closes
        w2 = w.copy()
        w2[0, 0] = np.uint64((int(w2[0, 0]) + 1) % P)
        out2 = oracle.plonk_partial_products_and_zs(w2, s, k, factor, betas, gammas)
        assert closing_value(w2, s, k, betas[0], gammas[0], int(out2[0, -1])) != 1
