"""ORACLE — TEST INFRASTRUCTURE ONLY.

Pure-Python big-integer restatement of the same path as ``oracle/*.c`` — slow, "obviously right",
written independently of the C code (``%`` on Python ints, O(n^2) DFT, recursive Merkle tree) and
used by ``tests/`` to cross-check the C oracle on small cases, and by the STARK verifier.

Upstream files restated (plonky2 0.2.2 / plonky2_field 0.2.2 / starky 0.4.0; pins at
/root/reference/Cargo.lock:3441,3466,4529; reached from /root/reference/ops/src/lib.rs:52):
field/src/goldilocks_field.rs, field/src/fft.rs, plonky2/src/hash/{poseidon,hashing,merkle_tree,
merkle_proofs}.rs, plonky2/src/iop/challenger.rs.
"""
from __future__ import annotations

P = 0xFFFFFFFF00000001
GENERATOR = 7
POWER_OF_TWO_GENERATOR = 1753635133440165772
W = 7  # extension: X^2 = 7

MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_DIAG = [8] + [0] * 11


def root_of_unity(n_log: int) -> int:
    return pow(POWER_OF_TWO_GENERATOR, 1 << (32 - n_log), P)


def bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


# ---------------------------------------------------------------- round constants (ChaCha8Rng(0))
def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def _chacha8_block(key, counter):
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [counter & 0xFFFFFFFF, counter >> 32, 0, 0]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(4):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, s)]


def derive_round_constants():
    """ChaCha8Rng::seed_from_u64(0) then 360 x gen_range(0..p) (rand 0.8.5)."""
    state = 0
    key = []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & (2**64 - 1)
        xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        key.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
    words = []
    ctr = 0
    out = []
    while len(out) < 360:
        if len(words) < 2:
            words += _chacha8_block(key, ctr)
            ctr += 1
        v = words[0] | (words[1] << 32)
        words = words[2:]
        m = v * P
        if (m & (2**64 - 1)) <= P - 1:
            out.append(m >> 64)
    return out


_RC = None


def round_constants():
    global _RC
    if _RC is None:
        _RC = derive_round_constants()
    return _RC


# ---------------------------------------------------------------- Poseidon
def _mds(s):
    return [(sum(s[(i + r) % 12] * MDS_CIRC[i] for i in range(12)) + s[r] * MDS_DIAG[r]) % P for r in range(12)]


def poseidon(state):
    rc = round_constants()
    s = [x % P for x in state]
    k = 0
    for rnd in range(30):
        s = [(s[i] + rc[k + i]) % P for i in range(12)]
        k += 12
        if rnd < 4 or rnd >= 26:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = _mds(s)
    return s


def hash_no_pad(inputs):
    st = [0] * 12
    for off in range(0, len(inputs), 8):
        chunk = inputs[off:off + 8]
        for i, v in enumerate(chunk):
            st[i] = v % P
        st = poseidon(st)
    return st[:4]


def hash_or_noop(inputs):
    if len(inputs) <= 4:
        return [x % P for x in inputs] + [0] * (4 - len(inputs))
    return hash_no_pad(inputs)


def two_to_one(l, r):
    return poseidon(list(l) + list(r) + [0] * 4)[:4]


# ---------------------------------------------------------------- Merkle tree (recursive definition)
def merkle_tree(leaves, cap_height):
    """Returns (digests in plonky2 layout, cap)."""
    n = len(leaves)
    n_cap = 1 << cap_height
    assert n >= n_cap

    def fill(sub):  # -> (digest list for this subtree, root)
        if len(sub) == 1:
            return [], hash_or_noop(sub[0])
        half = len(sub) // 2
        ld, lroot = fill(sub[:half])
        rd, rroot = fill(sub[half:])
        return ld + [lroot, rroot] + rd, two_to_one(lroot, rroot)

    digests, cap = [], []
    per = n // n_cap
    for s in range(n_cap):
        d, root = fill(leaves[s * per:(s + 1) * per])
        digests += d
        cap.append(root)
    return digests, cap


def merkle_verify(leaf, index, siblings, cap):
    cur = hash_or_noop(leaf)
    for sib in siblings:
        cur = two_to_one(sib, cur) if index & 1 else two_to_one(cur, sib)
        index >>= 1
    return cur == [x % P for x in cap[index]]


# ---------------------------------------------------------------- DFT by definition
def dft(a):
    n = len(a)
    w = root_of_unity(n.bit_length() - 1)
    return [sum(a[j] * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]


def idft(v):
    n = len(v)
    w_inv = pow(root_of_unity(n.bit_length() - 1), P - 2, P)
    n_inv = pow(n, P - 2, P)
    return [sum(v[j] * pow(w_inv, j * k, P) for j in range(n)) * n_inv % P for k in range(n)]


def eval_poly(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def lde_values(coeffs, rate_bits):
    """values[k] = P(7 * w_{n << rate_bits}^k), natural order"""
    big = len(coeffs) << rate_bits
    w = root_of_unity(big.bit_length() - 1)
    return [eval_poly(coeffs, GENERATOR * pow(w, k, P) % P) for k in range(big)]


# ---------------------------------------------------------------- extension field F_p[X]/(X^2 - 7)
def e_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def e_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def e_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def e_scalar(a, s):
    return (a[0] * s % P, a[1] * s % P)


def e_inv(a):
    norm = (a[0] * a[0] - W * a[1] * a[1]) % P
    ni = pow(norm, P - 2, P)
    return (a[0] * ni % P, (-a[1]) * ni % P)


def e_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = e_mul(r, a)
        a = e_mul(a, a)
        e >>= 1
    return r


def e_from(x):
    return (x % P, 0)


# ---------------------------------------------------------------- Challenger
class Challenger:
    def __init__(self):
        self.state = [0] * 12
        self.inb = []
        self.out = []

    def _duplex(self):
        for i, v in enumerate(self.inb):
            self.state[i] = v
        self.inb = []
        self.state = poseidon(self.state)
        self.out = list(self.state[:8])

    def observe(self, elems):
        for e in elems:
            self.out = []
            self.inb.append(int(e) % P)
            if len(self.inb) == 8:
                self._duplex()

    def get(self):
        if self.inb or not self.out:
            self._duplex()
        return self.out.pop()

    def get_n(self, n):
        return [self.get() for _ in range(n)]

    def compact(self):
        """Challenger::compact: flush pending inputs, drop buffered outputs, return the sponge state."""
        if self.inb:
            self._duplex()
        self.out = []
        return list(self.state)

    def get_ext(self):
        a = self.get_n(2)
        return (a[0], a[1])
