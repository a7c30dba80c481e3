"""Independent big-integer restatement of the Marlin protocol layer.  TEST INFRASTRUCTURE ONLY.

What simpleworks calls through src/marlin/mod.rs:45-94 (universal_setup / index / prove and the
canonical bytes of serialization.rs:5-31) lives in third-party crates that are not under
/root/reference: ark-marlin (git fork Entropy1729/marlin, branch use-constraint-system-directly, of
arkworks-rs/marlin master: joint 6-polynomial arithmetisation, SimpleHashFiatShamirRng) and
ark-poly-commit 0.3.0 (marlin_pc::MarlinKZG10 over kzg10).  This file restates those algorithms with
python integers only -- O(n log n) FFT, schoolbook polynomial products, double-and-add group
arithmetic, hashlib BLAKE2s, the ChaCha word streams of oracle/golden.py -- and was written from the
published protocol (SURVEY.md Appendix A.6-A.11 plus the upstream sources as the author knows them),
WITHOUT reading simpleworks_b200/csrc/marlin/marlin.hpp, so that it is a second, independent
implementation of the protocol layer: tests/golden/marlin_proofs.json is generated from it and both
engines of the C++ protocol code (CPU arm, CUDA) must reproduce its bytes.

PARITY UNPINNED against real arkworks: no Rust toolchain exists in this image, so these bytes have
never been diffed against the reference's own output.  Everything observable is a unique mathematical
object given the RNG streams (commitments are affine points, evaluations field elements), so what can
still differ from arkworks are conventions, each marked "convention:" below.

Upstream functions restated (file names of the crates, as cited in SURVEY.md 8c):
  marlin/src/lib.rs                    Marlin::{universal_setup, index, prove}
  marlin/src/ahp/mod.rs                max_degree, get_degree_bounds, construct_linear_combinations
  marlin/src/ahp/indexer.rs            AHPForR1CS::index
  marlin/src/ahp/constraint_systems.rs pad_input, make_matrices_square, sum_matrices, arithmetize_matrix
  marlin/src/ahp/prover.rs             prover_init, prover_{first,second,third}_round
  marlin/src/ahp/verifier.rs           verifier_{first,second,third}_round, verifier_query_set
  marlin/src/rng.rs                    SimpleHashFiatShamirRng<Blake2s, ChaChaRng>
  poly-commit/src/kzg10/mod.rs         setup, commit, compute_witness_polynomial, open
  poly-commit/src/marlin/marlin_pc/mod.rs  trim, commit, open (degree bounds, hiding)
  poly-commit/src/marlin/mod.rs        Marlin::open_combinations;  poly-commit/src/lib.rs batch_open
  ark-ec short_weierstrass_jacobian.rs UniformRand for GroupProjective, (de)serialisation flags
"""
from __future__ import annotations

import struct

from oracle import golden as G

P = G.R_MOD          # scalar field Fr
Q = G.Q_MOD          # base field Fq
PROTOCOL_NAME = b"MARLIN-2019"


# ------------------------------------------------------------------------------------------------
# Fr, polynomials (coefficient lists, low degree first, no trailing zeros)
# ------------------------------------------------------------------------------------------------
def finv(a: int) -> int:
    return pow(a, -1, P)


def batch_inverse(v):
    """ark_ff::batch_inversion: zeros stay zero"""
    return [0 if x == 0 else finv(x) for x in v]


def pnorm(p):
    p = [c % P for c in p]
    while p and p[-1] == 0:
        p.pop()
    return p


def padd(a, b):
    n = max(len(a), len(b))
    return pnorm([(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def psub(a, b):
    n = max(len(a), len(b))
    return pnorm([(a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0) for i in range(n)])


def pscale(a, k):
    return pnorm([c * k for c in a])


def pmul(a, b):
    if not a or not b:
        return []
    if min(len(a), len(b)) > 64:                 # same product through the FFT (checked against the schoolbook loop in self_check)
        log = (len(a) + len(b) - 2).bit_length()
        ea, eb = G.fft_fast(a, log), G.fft_fast(b, log)
        return pnorm(G.fft_fast([x * y % P for x, y in zip(ea, eb)], log, inverse=True))
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] += x * y
    return pnorm(out)


def peval(p, x):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % P
    return acc


def pdivmod_vanishing(p, n):
    """DensePolynomial::divide_by_vanishing_poly: (quotient, remainder) of p by X^n - 1"""
    p = list(p)
    if len(p) < n:
        return [], pnorm(p)
    q = [0] * (len(p) - n)
    for i in range(len(p) - 1, n - 1, -1):
        q[i - n] = p[i] % P
        p[i - n] = (p[i - n] + p[i]) % P
    return pnorm(q), pnorm(p[:n])


def pdiv_linear(p, z):
    """quotient of p by (X - z) (the remainder p(z) is dropped): kzg10::compute_witness_polynomial"""
    if len(p) < 2:
        return []
    q = [0] * (len(p) - 1)
    acc = 0
    for i in range(len(p) - 1, 0, -1):
        acc = (acc * z + p[i]) % P
        q[i - 1] = acc
    return pnorm(q)


class Domain:
    """ark_poly::Radix2EvaluationDomain::new(min_size) + the marlin crate's EvaluationDomainExt"""

    def __init__(self, min_size: int):
        self.log = 0
        while (1 << self.log) < min_size:
            self.log += 1
        assert self.log <= G.FR_TWO_ADICITY
        self.size = 1 << self.log
        self.gen = G.domain_gen(self.log)

    def elements(self):
        out, x = [], 1
        for _ in range(self.size):
            out.append(x)
            x = x * self.gen % P
        return out

    def fft(self, coeffs):
        assert len(coeffs) <= self.size
        return G.fft_fast(coeffs, self.log)

    def ifft(self, evals):
        return pnorm(G.fft_fast(evals, self.log, inverse=True))

    def vanishing(self, x):
        return (pow(x, self.size, P) - 1) % P

    def reindex_by_subdomain(self, other: "Domain", index: int) -> int:
        period = self.size // other.size
        if index < other.size:
            return index * period
        i = index - other.size
        x = period - 1
        return i + i // x + 1

    def bivariate_lagrange(self, x, y):
        """eval_unnormalized_bivariate_lagrange_poly: (v_H(x) - v_H(y)) / (x - y)"""
        if x != y:
            return (self.vanishing(x) - self.vanishing(y)) * finv((x - y) % P) % P
        return self.size * pow(x, self.size - 1, P) % P

    def lagrange_coefficients(self, tau):
        """evaluate_all_lagrange_coefficients"""
        z = self.vanishing(tau)
        els = self.elements()
        if z == 0:
            return [1 if e == tau else 0 for e in els]
        # L_i(tau) = v(tau) * w^i / (n * (tau - w^i))
        ninv = finv(self.size)
        return [z * e % P * ninv % P * finv((tau - e) % P) % P for e in els]


# ------------------------------------------------------------------------------------------------
# Fq, Fq2 = Fq[u] / (u^2 + 5), curves G1: y^2 = x^3 + 1 and G2: y^2 = x^3 + B2 (affine; None = identity)
# ------------------------------------------------------------------------------------------------
def qinv(a: int) -> int:
    return pow(a, -1, Q)


def fq_sqrt(a: int):
    """some square root of a in Fq, or None (Tonelli-Shanks; q = 1 mod 2^46)"""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 0, Q - 1
    while t % 2 == 0:
        t //= 2
        s += 1
    z = 2
    while pow(z, (Q - 1) // 2, Q) == 1:
        z += 1
    c = pow(z, t, Q)
    x = pow(a, (t + 1) // 2, Q)
    b = pow(a, t, Q)
    m = s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % Q
            i += 1
        e = pow(c, 1 << (m - i - 1), Q)
        x = x * e % Q
        c = e * e % Q
        b = b * c % Q
        m = i
    return x


FQ2_NONRESIDUE = Q - 5          # u^2 = -5


def f2(a, b=0):
    return (a % Q, b % Q)


def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] + FQ2_NONRESIDUE * a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a):
    n = qinv((a[0] * a[0] - FQ2_NONRESIDUE * a[1] * a[1]) % Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


def f2_sqrt(a):
    """some square root in Fq2 or None"""
    if a == (0, 0):
        return (0, 0)
    if a[1] == 0:
        r = fq_sqrt(a[0])
        if r is not None:
            return (r, 0)
        # a0 is a non-residue: sqrt = sqrt(a0 / nonresidue) * u
        r = fq_sqrt(a[0] * qinv(FQ2_NONRESIDUE) % Q)
        return None if r is None else (0, r)
    norm = (a[0] * a[0] - FQ2_NONRESIDUE * a[1] * a[1]) % Q
    alpha = fq_sqrt(norm)
    if alpha is None:
        return None
    two_inv = qinv(2)
    delta = (a[0] + alpha) * two_inv % Q
    c0 = fq_sqrt(delta)
    if c0 is None:
        delta = (a[0] - alpha) * two_inv % Q
        c0 = fq_sqrt(delta)
        if c0 is None:
            return None
    c1 = a[1] * qinv(2 * c0 % Q) % Q
    r = (c0, c1)
    assert f2_mul(r, r) == a
    return r


def f2_lt(a, b):
    """Ord of ark_ff::QuadExtField: c1 first, then c0 (canonical integers)"""
    return (a[1], a[0]) < (b[1], b[0])


class Curve:
    """short Weierstrass curve y^2 = x^3 + b with a = 0 over Fq (elements: ints) or Fq2 (pairs)"""

    def __init__(self, b, ext: bool):
        self.b, self.ext = b, ext
        if ext:
            self.add_, self.sub_, self.mul_, self.inv_, self.neg_ = f2_add, f2_sub, f2_mul, f2_inv, f2_neg
            self.zero, self.three, self.two = (0, 0), (3, 0), (2, 0)
        else:
            self.add_ = lambda a, c: (a + c) % Q
            self.sub_ = lambda a, c: (a - c) % Q
            self.mul_ = lambda a, c: a * c % Q
            self.inv_ = qinv
            self.neg_ = lambda a: (-a) % Q
            self.zero, self.three, self.two = 0, 3, 2

    def on_curve(self, p):
        if p is None:
            return True
        x, y = p
        return self.mul_(y, y) == self.add_(self.mul_(self.mul_(x, x), x), self.b)

    def neg(self, p):
        return None if p is None else (p[0], self.neg_(p[1]))

    def add(self, p, q):
        if p is None:
            return q
        if q is None:
            return p
        x1, y1 = p
        x2, y2 = q
        if x1 == x2:
            if self.add_(y1, y2) == self.zero:
                return None
            lam = self.mul_(self.mul_(self.three, self.mul_(x1, x1)), self.inv_(self.mul_(self.two, y1)))
        else:
            lam = self.mul_(self.sub_(y2, y1), self.inv_(self.sub_(x2, x1)))
        x3 = self.sub_(self.sub_(self.mul_(lam, lam), x1), x2)
        y3 = self.sub_(self.mul_(lam, self.sub_(x1, x3)), y1)
        return (x3, y3)

    def mul(self, p, k: int):
        acc, addend = None, p
        while k:
            if k & 1:
                acc = self.add(acc, addend)
            addend = self.add(addend, addend)
            k >>= 1
        return acc


E1 = Curve(1, False)
# ark-bls12-377 curves/g2.rs COEFF_B = (0, 1558...906); equals 1 / u = -u / 5 (checked in self_check)
G2_B = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
E2 = Curve(G2_B, True)
G1_COFACTOR = G.G1_COFACTOR


def _g2_cofactor() -> int:
    """#E'(Fq2) / r from the trace of Frobenius (no constant copied): t = q + 1 - h1 r; t2 = t^2 - 2q;
    t2^2 - 4 q^2 = -3 f^2; the sextic twists have q^2 + 1 - (+-t2 +- 3f) / 2 points; ours is the one whose
    order r divides and on which random points of y^2 = x^3 + B2 vanish (self_check)."""
    import math
    t = Q + 1 - G1_COFACTOR * P
    t2 = t * t - 2 * Q
    f2_ = (4 * Q * Q - t2 * t2) // 3
    f = math.isqrt(f2_)
    assert f * f == f2_
    cands = []
    for st in (1, -1):
        for sf in (1, -1):
            num = st * t2 + sf * 3 * f
            if num % 2 == 0:
                n = Q * Q + 1 - num // 2
                if n % P == 0:
                    cands.append(n // P)
    return cands


G2_COFACTOR_CANDIDATES = _g2_cofactor()
# ark-bls12-377 curves/g2.rs COFACTOR (little-endian u64 limbs, as the author recalls them); must be one of the
# candidates derived above
G2_COFACTOR = G.from_limbs([0x0000000000000001, 0x452217cc90000000, 0xa0f3622fba094800, 0xd693e8c36676bd09,
                            0x8c505634fae2e189, 0xfbb36b00e1dcc40c, 0xddd88d99a6f6a829, 0x26ba558ae9562a])


def fq_rand(rng) -> int:
    return G.fq_from_mont(G.fq_rand_mont(rng))


def fr_rand(rng) -> int:
    return G.fr_from_mont(G.fr_rand_mont(rng))


def rand_bool(rng) -> bool:
    """rand 0.8 Standard for bool: (next_u32 as i32) < 0"""
    return rng.next_u32() >= 0x80000000


def g1_rand(rng):
    """UniformRand for GroupProjective (ark-ec 0.3 short_weierstrass_jacobian.rs): x <- Fq::rand, greatest <- bool,
    get_point_from_x: y = sqrt(x^3 + b), the root selected by ((y < -y) ^ greatest ? y : -y); then the cofactor"""
    while True:
        x = fq_rand(rng)
        greatest = rand_bool(rng)
        y = fq_sqrt((x * x * x + 1) % Q)
        if y is None:
            continue
        negy = (-y) % Q
        y = y if ((y < negy) ^ greatest) else negy
        return E1.mul((x, y), G1_COFACTOR)


def g2_rand(rng):
    while True:
        x = (fq_rand(rng), fq_rand(rng))          # Fp2::rand: c0 then c1
        greatest = rand_bool(rng)
        y = f2_sqrt(f2_add(f2_mul(f2_mul(x, x), x), G2_B))
        if y is None:
            continue
        negy = f2_neg(y)
        y = y if (f2_lt(y, negy) ^ greatest) else negy
        return E2.mul((x, y), G2_COFACTOR)


# ------------------------------------------------------------------------------------------------
# bytes: ark_ff::ToBytes (transcript) and ark_serialize::CanonicalSerialize (proof / key bytes)
# ------------------------------------------------------------------------------------------------
def u64le(x: int) -> bytes:
    return struct.pack("<Q", x)


def fr_bytes(x: int) -> bytes:
    return (x % P).to_bytes(32, "little")


def g1_tobytes(p) -> bytes:
    """ToBytes for GroupAffine: x | y | infinity flag; the identity is (0, 1, true)"""
    if p is None:
        return (0).to_bytes(48, "little") + (1).to_bytes(48, "little") + b"\x01"
    return p[0].to_bytes(48, "little") + p[1].to_bytes(48, "little") + b"\x00"


def g1_compressed(p) -> bytes:
    """CanonicalSerialize for GroupAffine: x with SWFlags in the top two bits of the last byte
    (bit 7: y > -y, bit 6: infinity)"""
    if p is None:
        b = bytearray(48)
        b[47] |= 1 << 6
        return bytes(b)
    b = bytearray(p[0].to_bytes(48, "little"))
    if p[1] > (-p[1]) % Q:
        b[47] |= 1 << 7
    return bytes(b)


def g2_compressed(p) -> bytes:
    if p is None:
        b = bytearray(96)
        b[95] |= 1 << 6
        return bytes(b)
    x, y = p
    b = bytearray(x[0].to_bytes(48, "little") + x[1].to_bytes(48, "little"))
    if f2_lt(f2_neg(y), y):
        b[95] |= 1 << 7
    return bytes(b)


class FiatShamirRng:
    """SimpleHashFiatShamirRng<Blake2s, ChaChaRng> (marlin/src/rng.rs)"""

    def __init__(self, initial: bytes):
        self.seed = G.blake2s(initial)
        self.r = G.ChaChaRng(self.seed, rounds=20)

    def absorb(self, data: bytes):
        self.seed = G.blake2s(data + self.seed)
        self.r = G.ChaChaRng(self.seed, rounds=20)

    def next_u32(self): return self.r.next_u32()
    def next_u64(self): return self.r.next_u64()


# ------------------------------------------------------------------------------------------------
# R1CS as ConstraintSystem::to_matrices() exposes it
# ------------------------------------------------------------------------------------------------
class R1CS:
    def __init__(self, num_instance: int, num_witness: int):
        self.num_instance = num_instance            # includes the constant one
        self.num_witness = num_witness
        self.a, self.b, self.c = [], [], []           # rows: lists of (coefficient, column)
        self.instance, self.witness = None, None

    def enforce(self, a, b, c):
        self.a.append([(k % P, j) for k, j in a])
        self.b.append([(k % P, j) for k, j in b])
        self.c.append([(k % P, j) for k, j in c])

    def assign(self, instance, witness):
        assert len(instance) == self.num_instance and len(witness) == self.num_witness and instance[0] == 1
        self.instance, self.witness = [x % P for x in instance], [x % P for x in witness]

    def copy(self) -> "R1CS":
        o = R1CS(self.num_instance, self.num_witness)
        o.a, o.b, o.c = [list(r) for r in self.a], [list(r) for r in self.b], [list(r) for r in self.c]
        o.instance = None if self.instance is None else list(self.instance)
        o.witness = None if self.witness is None else list(self.witness)
        return o

    def is_satisfied(self) -> bool:
        z = self.instance + self.witness
        dot = lambda row: sum(k * z[j] for k, j in row) % P
        return all(dot(ra) * dot(rb) % P == dot(rc) for ra, rb, rc in zip(self.a, self.b, self.c))

    def pad_and_square(self):
        """pad_input_for_indexer_and_prover + make_matrices_square: instance count to a power of two with
        zero inputs (witness columns move up), then dummy 0 * 0 = 0 constraints or dummy witnesses (value 1)"""
        target = Domain(self.num_instance).size
        add = target - self.num_instance
        if add:
            sh = lambda m: [[(k, j + add if j >= self.num_instance else j) for k, j in row] for row in m]
            self.a, self.b, self.c = sh(self.a), sh(self.b), sh(self.c)
            if self.instance is not None:
                self.instance += [0] * add
            self.num_instance = target
        nv, nc = self.num_instance + self.num_witness, len(self.a)
        if nv > nc:
            for _ in range(nv - nc):
                self.a.append([]); self.b.append([]); self.c.append([])
        elif nc > nv:
            self.num_witness += nc - nv
            if self.witness is not None:
                self.witness += [1] * (nc - nv)


def circuit_manual_constraints(a_val: int, b_val: int) -> R1CS:
    """examples/manual-constraints.rs:16-31: instance [1, a], witness [b], (a - b) * 1 = 0.
    (the example test passes Fr::new(BigInteger256::new([1, 0, 0, 0])), i.e. the value R^-1 mod r, :90-99)"""
    cs = R1CS(2, 1)
    cs.enforce([(1, 1), (P - 1, 2)], [(1, 0)], [])
    cs.assign([1, a_val], [b_val])
    return cs


def circuit_uint8_equality(a_val: int, b_val: int) -> R1CS:
    """examples/test-circuit.rs:13-26: UInt8::new_witness twice (8 Boolean witnesses each, LSB first, each with
    (1 - x) * x = 0), then enforce_equal = 8 constraints (b_i - a_i) * 1 = 0"""
    cs = R1CS(1, 16)
    for i in range(16):
        cs.enforce([(1, 0), (P - 1, 1 + i)], [(1, 1 + i)], [])
    for i in range(8):
        cs.enforce([(P - 1, 1 + i), (1, 1 + 8 + i)], [(1, 0)], [])
    cs.assign([1], [(a_val >> i) & 1 for i in range(8)] + [(b_val >> i) & 1 for i in range(8)])
    return cs


def circuit_mul_chain(n: int, seed0: int, seed1: int) -> R1CS:
    """synthetic chain x_i * x_{i+1} = x_{i+2} with one public input x_0 (BASELINE configs[3])"""
    x = [seed0 % P, seed1 % P]
    for i in range(n):
        x.append(x[i] * x[i + 1] % P)
    cs = R1CS(2, n + 1)
    col = lambda i: 1 if i == 0 else 1 + i
    for i in range(n):
        cs.enforce([(1, col(i))], [(1, col(i + 1))], [(1, col(i + 2))])
    cs.assign([1, x[0]], x[1:])
    return cs


# ------------------------------------------------------------------------------------------------
# KZG10 / MarlinKZG10
# ------------------------------------------------------------------------------------------------
def max_degree(num_constraints: int, num_variables: int, num_non_zero: int) -> int:
    """AHPForR1CS::max_degree with zk_bound = 1"""
    h = Domain(max(num_variables, num_constraints)).size
    k = Domain(num_non_zero).size
    return max(2 * h + 1 - 2, 3 * h + 2 - 3, h, h, 3 * k - 3)


class SRS:
    """kzg10::UniversalParams; the python model keeps the trapdoor so that a commitment is one scalar
    multiplication: MSM(powers_of_g, coeffs) = p(beta) * g.  `powers(i)` materialises individual SRS points."""

    def __init__(self, beta, g, gamma_g, h, max_deg):
        self.beta, self.g, self.gamma_g, self.h, self.max_degree = beta, g, gamma_g, h, max_deg
        self.beta_h = E2.mul(h, beta)

    def power_of_g(self, i: int):
        return E1.mul(self.g, pow(self.beta, i, P))

    def power_of_gamma_g(self, i: int):
        return E1.mul(self.gamma_g, pow(self.beta, i, P))


def universal_setup(num_constraints: int, num_variables: int, num_non_zero: int, rng) -> SRS:
    """Marlin::universal_setup -> MarlinKZG10::setup -> KZG10::setup(max_degree, false, rng): draws beta, g,
    gamma_g, h in this order"""
    d = max_degree(num_constraints, num_variables, num_non_zero)
    beta = fr_rand(rng)
    g = g1_rand(rng)
    gamma_g = g1_rand(rng)
    h = g2_rand(rng)
    return SRS(beta, g, gamma_g, h, d)


class LabeledPoly:
    def __init__(self, label, coeffs, degree_bound=None, hiding_bound=None):
        self.label, self.coeffs, self.degree_bound, self.hiding_bound = label, pnorm(coeffs), degree_bound, hiding_bound


class Rand:
    """marlin_pc::Randomness: blinding polynomial of the commitment and of its shifted twin"""

    def __init__(self, rand=None, shifted=None):
        self.rand = rand or []
        self.shifted = shifted          # None = no degree bound


class CommitterKey:
    def __init__(self, srs: SRS, supported_degree: int, bounds):
        self.srs, self.supported_degree = srs, supported_degree
        self.bounds = sorted(set(bounds))
        self.max_degree = srs.max_degree

    def commit_plain(self, coeffs, shift: int = 0):
        """MSM(powers_of_g[shift..], coeffs) via the trapdoor"""
        return E1.mul(self.srs.g, peval(coeffs, self.srs.beta) * pow(self.srs.beta, shift, P) % P)

    def commit_gamma(self, coeffs):
        assert len(coeffs) <= 3           # trim keeps powers_of_gamma_g[0 ..= hiding_bound + 1]
        return E1.mul(self.srs.gamma_g, peval(coeffs, self.srs.beta))

    def shift_of(self, bound: int) -> int:
        """shifted_powers(bound) = powers_of_g[max_degree - bound ..] (max_degree of the SRS, not of the index)"""
        assert bound in self.bounds
        return self.max_degree - bound


def kzg_commit(ck: CommitterKey, coeffs, hiding_bound, rng, shift=0):
    """kzg10::commit: plain MSM, then (if hiding) blinding polynomial = DensePolynomial::rand(hiding_bound + 1)
    against powers_of_gamma_g"""
    assert len(coeffs) - 1 + shift <= ck.max_degree
    c = ck.commit_plain(coeffs, shift)
    blind = []
    if hiding_bound is not None:
        blind = pnorm([fr_rand(rng) for _ in range(hiding_bound + 2)])
        c = E1.add(c, ck.commit_gamma(blind))
    return c, blind


def pc_commit(ck: CommitterKey, polys, rng):
    """marlin_pc::commit: per polynomial [unshifted commit][shifted commit with a fresh blinder if degree-bounded]"""
    comms, rands = [], []
    for p in polys:
        assert len(p.coeffs) - 1 <= ck.supported_degree
        c, r = kzg_commit(ck, p.coeffs, p.hiding_bound, rng)
        sc, sr = None, None
        if p.degree_bound is not None:
            assert len(p.coeffs) - 1 <= p.degree_bound
            sc, sr = kzg_commit(ck, p.coeffs, p.hiding_bound, rng, shift=ck.shift_of(p.degree_bound))
        comms.append((c, sc))
        rands.append(Rand(r, sr))
    return comms, rands


def comm_tobytes(c) -> bytes:
    """ToBytes for marlin_pc::Commitment: comm | shifted_exists | shifted_comm or kzg10::Commitment::empty()"""
    return g1_tobytes(c[0]) + (b"\x01" if c[1] is not None else b"\x00") + g1_tobytes(c[1])


def comm_serialize(c) -> bytes:
    """CanonicalSerialize (derived): comm, Option<shifted_comm>"""
    return g1_compressed(c[0]) + (b"\x01" + g1_compressed(c[1]) if c[1] is not None else b"\x00")


# ------------------------------------------------------------------------------------------------
# AHP indexer
# ------------------------------------------------------------------------------------------------
INDEXER_POLYNOMIALS = ["row", "col", "a_val", "b_val", "c_val", "row_col"]


class Index:
    pass


def ahp_index(cs_in: R1CS) -> Index:
    cs = cs_in.copy()
    cs.pad_and_square()
    nv, nc = cs.num_instance + cs.num_witness, len(cs.a)
    assert nv == nc, "NonSquareMatrix"
    joint = [sorted(set(j for _, j in ra) | set(j for _, j in rb) | set(j for _, j in rc)) for ra, rb, rc in zip(cs.a, cs.b, cs.c)]
    nnz = sum(len(r) for r in joint)
    ix = Index()
    ix.num_variables, ix.num_constraints, ix.num_non_zero, ix.num_instance = nv, nc, nnz, cs.num_instance
    ix.a, ix.b, ix.c = cs.a, cs.b, cs.c
    H, K, X = Domain(nc), Domain(nnz), Domain(cs.num_instance)
    ix.H, ix.K, ix.X = H, K, X
    elems = H.elements()
    # u_H(x, x) = |H| x^(|H| - 1)
    eq = {e: H.size * pow(e, H.size - 1, P) % P for e in elems}
    row, col, va, vb, vc, inv = [], [], [], [], [], []
    lookup = lambda m, r, j: sum(k for k, jj in m[r] if jj == j) % P
    for r, cols in enumerate(joint):
        for j in cols:
            row_val = elems[r]
            col_val = elems[H.reindex_by_subdomain(X, j)]
            # arithmetisation of M*(i, j) = M(j, i) u_H(j, j): the transpose
            row.append(col_val)
            col.append(row_val)
            va.append(lookup(cs.a, r, j)); vb.append(lookup(cs.b, r, j)); vc.append(lookup(cs.c, r, j))
            inv.append(eq[col_val])
    inv = batch_inverse(inv)
    va = [v * i % P for v, i in zip(va, inv)]
    vb = [v * i % P for v, i in zip(vb, inv)]
    vc = [v * i % P for v, i in zip(vc, inv)]
    pad = K.size - len(row)
    row += [elems[0]] * pad; col += [elems[0]] * pad
    va += [0] * pad; vb += [0] * pad; vc += [0] * pad
    row_col = [r * c % P for r, c in zip(row, col)]
    ix.evals = {"row": row, "col": col, "a_val": va, "b_val": vb, "c_val": vc, "row_col": row_col}
    ix.polys = [LabeledPoly(l, K.ifft(ix.evals[l])) for l in INDEXER_POLYNOMIALS]
    return ix


def index(srs: SRS, cs: R1CS):
    """Marlin::index: AHP index, trim, commit to the six index polynomials (no hiding, no bounds, no rng)"""
    ix = ahp_index(cs)
    deg = max_degree(ix.num_constraints, ix.num_variables, ix.num_non_zero)
    assert srs.max_degree >= deg, "IndexTooLarge"
    bounds = [ix.H.size - 2, ix.K.size - 2]            # get_degree_bounds
    ck = CommitterKey(srs, deg, bounds)
    comms, rands = pc_commit(ck, ix.polys, None)
    vk = {"info": (ix.num_variables, ix.num_constraints, ix.num_non_zero, ix.num_instance), "comms": comms,
          "g": srs.g, "gamma_g": srs.gamma_g, "h": srs.h, "beta_h": srs.beta_h,
          "shift_powers": [(b, srs.power_of_g(srs.max_degree - b)) for b in ck.bounds],
          "max_degree": srs.max_degree, "supported_degree": deg}
    pk = {"index": ix, "ck": ck, "vk": vk, "rands": rands}
    return pk, vk


def vk_tobytes(vk) -> bytes:
    """ToBytes for IndexVerifierKey (what the transcript absorbs): three u64 sizes + the index commitments"""
    out = u64le(vk["info"][0]) + u64le(vk["info"][1]) + u64le(vk["info"][2])
    for c in vk["comms"]:
        out += comm_tobytes(c)
    return out


def vk_serialize(vk) -> bytes:
    """serialize_verifying_key (reference src/marlin/serialization.rs:19-25): derived CanonicalSerialize of
    IndexVerifierKey { index_info, index_comms, verifier_key: marlin_pc::VerifierKey { vk: kzg10::VerifierKey
    { g, gamma_g, h, beta_h }, degree_bounds_and_shift_powers, max_degree, supported_degree } }.
    convention: IndexInfo is written as four u64 (num_variables, num_constraints, num_non_zero,
    num_instance_variables)."""
    out = b"".join(u64le(x) for x in vk["info"])
    out += u64le(len(vk["comms"])) + b"".join(comm_serialize(c) for c in vk["comms"])
    out += g1_compressed(vk["g"]) + g1_compressed(vk["gamma_g"]) + g2_compressed(vk["h"]) + g2_compressed(vk["beta_h"])
    out += b"\x01" + u64le(len(vk["shift_powers"]))
    for b, p in vk["shift_powers"]:
        out += u64le(b) + g1_compressed(p)
    out += u64le(vk["max_degree"]) + u64le(vk["supported_degree"])
    return out


# ------------------------------------------------------------------------------------------------
# prover
# ------------------------------------------------------------------------------------------------
def prove(pk, cs_in: R1CS, zk_rng, trace: dict | None = None) -> bytes:
    """Marlin::prove; returns serialize_proof(proof) (reference src/marlin/mod.rs:70-77, serialization.rs:5-12)"""
    ix, ck, vk = pk["index"], pk["ck"], pk["vk"]
    H, K, X = ix.H, ix.K, ix.X
    # ---- prover_init
    cs = cs_in.copy()
    cs.pad_and_square()
    assert len(cs.a) == ix.num_constraints and cs.num_instance + cs.num_witness == ix.num_variables, "InstanceDoesNotMatchIndex"
    formatted_input, witness = cs.instance, cs.witness
    z = formatted_input + witness
    if not all(sum(k * z[j] for k, j in ra) * sum(k * z[j] for k, j in rb) % P == sum(k * z[j] for k, j in rc) % P
               for ra, rb, rc in zip(ix.a, ix.b, ix.c)):
        raise ValueError("unsatisfied instance")
    z_a = [sum(k * z[j] for k, j in r) % P for r in ix.a]
    z_b = [sum(k * z[j] for k, j in r) % P for r in ix.b]
    public_input = formatted_input[1:]                # unformat_public_input: without the leading one, WITH the padding
    fs = FiatShamirRng(PROTOCOL_NAME + vk_tobytes(vk) + b"".join(fr_bytes(x) for x in public_input))

    # ---- first round
    v_H = [P - 1] + [0] * (H.size - 1) + [1]
    x_poly = X.ifft(formatted_input)
    x_evals = H.fft(x_poly)
    ratio = H.size // X.size
    w_ext = witness + [0] * (H.size - X.size - len(witness))
    w_evals = [0 if k % ratio == 0 else (w_ext[k - k // ratio - 1] - x_evals[k]) % P for k in range(H.size)]
    w_poly = padd(H.ifft(w_evals), pscale(v_H, fr_rand(zk_rng)))
    w_poly, rem = pdivmod_vanishing(w_poly, X.size)
    assert rem == []
    z_a_poly = padd(H.ifft(z_a), pscale(v_H, fr_rand(zk_rng)))
    z_b_poly = padd(H.ifft(z_b), pscale(v_H, fr_rand(zk_rng)))
    mask_degree = 3 * H.size + 2 - 3
    mask = [fr_rand(zk_rng) for _ in range(mask_degree + 1)]
    r0 = sum(mask[H.size * i] for i in range(mask_degree // H.size + 1)) % P
    mask[0] = (mask[0] - r0) % P
    first = [LabeledPoly("w", w_poly, None, 1), LabeledPoly("z_a", z_a_poly, None, 1), LabeledPoly("z_b", z_b_poly, None, 1),
             LabeledPoly("mask_poly", mask, None, None)]
    first_comms, first_rands = pc_commit(ck, first, zk_rng)
    fs.absorb(b"".join(comm_tobytes(c) for c in first_comms))          # + EmptyMessage = no bytes

    def sample_outside(dom):
        t = fr_rand(fs)
        while dom.vanishing(t) == 0:
            t = fr_rand(fs)
        return t
    alpha = sample_outside(H)
    eta_a, eta_b, eta_c = fr_rand(fs), fr_rand(fs), fr_rand(fs)

    # ---- second round
    z_c_poly = pmul(z_a_poly, z_b_poly)
    summed_z_m = padd(pscale(z_c_poly, eta_c), padd(pscale(z_a_poly, eta_a), pscale(z_b_poly, eta_b)))
    elems = H.elements()
    v_H_alpha = H.vanishing(alpha)
    r_alpha_x = [v_H_alpha * finv((alpha - e) % P) % P for e in elems]
    r_alpha_poly = H.ifft(r_alpha_x)
    t_evals = [0] * H.size
    for m, eta in ((ix.a, eta_a), (ix.b, eta_b), (ix.c, eta_c)):
        for r, rowm in enumerate(m):
            for k, j in rowm:
                t_evals[H.reindex_by_subdomain(X, j)] += eta * k % P * r_alpha_x[r]
    t_poly = H.ifft(t_evals)
    v_X = [P - 1] + [0] * (X.size - 1) + [1]
    z_poly = padd(pmul(w_poly, v_X), x_poly)
    q_1 = padd(mask, psub(pmul(r_alpha_poly, summed_z_m), pmul(t_poly, z_poly)))
    h_1, x_g_1 = pdivmod_vanishing(q_1, H.size)
    g_1 = pnorm(x_g_1[1:])
    assert len(g_1) - 1 <= H.size - 2
    second = [LabeledPoly("t", t_poly, None, None), LabeledPoly("g_1", g_1, H.size - 2, 1), LabeledPoly("h_1", h_1, None, 1)]
    second_comms, second_rands = pc_commit(ck, second, zk_rng)
    fs.absorb(b"".join(comm_tobytes(c) for c in second_comms))
    beta = sample_outside(H)

    # ---- third round
    v_H_beta = H.vanishing(beta)
    vv = v_H_alpha * v_H_beta % P
    ev = ix.evals
    polys_ix = {p.label: p for p in ix.polys}
    a_poly = padd(pscale(polys_ix["a_val"].coeffs, eta_a * vv), padd(pscale(polys_ix["b_val"].coeffs, eta_b * vv),
                                                                    pscale(polys_ix["c_val"].coeffs, eta_c * vv)))
    b_evals = [(alpha * beta - alpha * r - beta * c + rc) % P for r, c, rc in zip(ev["row"], ev["col"], ev["row_col"])]
    b_poly = K.ifft(b_evals)
    invs = batch_inverse([(beta - r) * (alpha - c) % P for r, c in zip(ev["row"], ev["col"])])
    f_evals = [i * (eta_a * vv % P * a + eta_b * vv % P * b + eta_c * vv % P * c) % P
               for i, a, b, c in zip(invs, ev["a_val"], ev["b_val"], ev["c_val"])]
    f = K.ifft(f_evals)
    g_2 = pnorm(f[1:])
    h_2, rem2 = pdivmod_vanishing(psub(a_poly, pmul(b_poly, f)), K.size)
    assert rem2 == []
    third = [LabeledPoly("g_2", g_2, K.size - 2, None), LabeledPoly("h_2", h_2, None, None)]
    third_comms, third_rands = pc_commit(ck, third, zk_rng)
    fs.absorb(b"".join(comm_tobytes(c) for c in third_comms))
    gamma = fr_rand(fs)

    # ---- query set, linear combinations, evaluations
    polys = {p.label: p for p in ix.polys + first + second + third}
    rands = dict(zip([p.label for p in ix.polys + first + second + third], pk["rands"] + first_rands + second_rands + third_rands))
    ev_at = lambda label, x: peval(polys[label].coeffs, x)
    z_b_beta, t_beta, g_1_beta, g_2_gamma = ev_at("z_b", beta), ev_at("t", beta), ev_at("g_1", beta), ev_at("g_2", gamma)
    r_alpha_beta = H.bivariate_lagrange(alpha, beta)
    v_X_beta = X.vanishing(beta)
    x_beta = sum(l * x for l, x in zip(X.lagrange_coefficients(beta), formatted_input)) % P
    ONE = None      # LCTerm::One
    outer = [(1, "mask_poly"), (r_alpha_beta * (eta_a + eta_c * z_b_beta) % P, "z_a"), (r_alpha_beta * eta_b % P * z_b_beta % P, ONE),
             ((-t_beta * v_X_beta) % P, "w"), ((-t_beta * x_beta) % P, ONE), ((-v_H_beta) % P, "h_1"), ((-beta * g_1_beta) % P, ONE)]
    scal = (gamma * g_2_gamma + t_beta * finv(K.size)) % P
    inner = [(eta_a * vv % P, "a_val"), (eta_b * vv % P, "b_val"), (eta_c * vv % P, "c_val"),
             ((-scal * alpha % P * beta) % P, ONE), (scal * alpha % P, "row"), (scal * beta % P, "col"), ((-scal) % P, "row_col"),
             ((-K.vanishing(gamma)) % P, "h_2")]
    lcs = {"g_1": [(1, "g_1")], "g_2": [(1, "g_2")], "t": [(1, "t")], "z_b": [(1, "z_b")], "outer_sumcheck": outer, "inner_sumcheck": inner}

    def lc_eval(lc, x):
        return sum(c * (1 if l is ONE else ev_at(l, x)) for c, l in lc) % P
    assert lc_eval(outer, beta) == 0, "outer sumcheck does not vanish"
    assert lc_eval(inner, gamma) == 0, "inner sumcheck does not vanish"
    # QuerySet is a BTreeSet<(label, (point label, point))>; evaluations of the LCs that are not known to be zero, by label
    evaluations = [g_1_beta, g_2_gamma, t_beta, z_b_beta]
    fs.absorb(b"".join(fr_bytes(x) for x in evaluations))
    lo = fs.next_u64()
    hi = fs.next_u64()
    xi = ((hi << 64) | lo) % P                  # opening_challenge: F = u128::rand(&mut fs_rng).into()

    # ---- open_combinations: LC polynomials without their constant terms, then batch_open per point label
    def lc_poly(lc):
        poly, rnd, shifted = [], [], None
        terms = [(c, l) for c, l in lc if l is not ONE]
        bound, hiding = None, None
        for c, l in terms:
            p = polys[l]
            if len(lc) == 1 and p.degree_bound is not None:
                assert c == 1
                bound = p.degree_bound
            else:
                assert p.degree_bound is None, "EquationHasDegreeBounds"
            if p.hiding_bound is not None:
                hiding = max(hiding or 0, p.hiding_bound)
            poly = padd(poly, pscale(p.coeffs, c))
            rnd = padd(rnd, pscale(rands[l].rand, c))
            if rands[l].shifted is not None:
                shifted = padd(shifted or [], pscale(rands[l].shifted, c))
        return LabeledPoly("", poly, bound, hiding), Rand(rnd, shifted)

    def open_at(labels, point):
        """marlin_pc::open (ark-poly-commit 0.3.0 open_individual_opening_challenges): challenge exponents from a
        counter that advances once per polynomial and once more for a degree-bounded one's shifted part"""
        p, r = [], []
        shifted_w, shifted_r, shifted_r_witness = [], [], []
        enforce, counter = False, 0
        for label in labels:
            lp, rnd = lc_poly(lcs[label])
            ch = pow(xi, counter, P)
            counter += 1
            p = padd(p, pscale(lp.coeffs, ch))
            r = padd(r, pscale(rnd.rand, ch))
            if lp.degree_bound is not None:
                enforce = True
                ch1 = pow(xi, counter, P)
                counter += 1
                witness = pdiv_linear(lp.coeffs, point)
                shift = ck.bounds[-1] - lp.degree_bound
                shifted_witness = ([0] * shift + witness) if witness else []
                shifted_w = padd(shifted_w, pscale(shifted_witness, ch1))
                shifted_r = padd(shifted_r, pscale(rnd.shifted, ch1))
                if rnd.shifted:
                    shifted_r_witness = padd(shifted_r_witness, pscale(pdiv_linear(rnd.shifted, point), ch1))
        # kzg10::open on the unshifted combination
        w = ck.commit_plain(pdiv_linear(p, point))
        random_v = None
        if r:
            w = E1.add(w, ck.commit_gamma(pdiv_linear(r, point)))
            random_v = peval(r, point)
        if enforce:
            # open_with_witness_polynomial over shifted_powers(None) = powers_of_g[max_degree - largest bound ..]
            sw = ck.commit_plain(shifted_w, shift=ck.max_degree - ck.bounds[-1])
            sw = E1.add(sw, ck.commit_gamma(shifted_r_witness))
            w = E1.add(w, sw)
            if random_v is not None:
                random_v = (random_v + peval(shifted_r, point)) % P
        return w, random_v

    proof_beta = open_at(["g_1", "outer_sumcheck", "t", "z_b"], beta)            # BTreeMap by point label, BTreeSet of labels
    proof_gamma = open_at(["g_2", "inner_sumcheck"], gamma)

    if trace is not None:
        trace.update(alpha=alpha, eta=(eta_a, eta_b, eta_c), beta=beta, gamma=gamma, xi=xi, evaluations=evaluations,
                     comms=[first_comms, second_comms, third_comms], polys={k: v.coeffs for k, v in polys.items()})
    # ---- serialize_proof: Proof { commitments: Vec<Vec<Commitment>>, evaluations: Vec<F>, prover_messages: Vec<ProverMsg>,
    # pc_proof: BatchLCProof { proof: Vec<kzg10::Proof { w, random_v: Option<F> }>, evals: Option<Vec<F>> } }
    out = u64le(3)
    for group in (first_comms, second_comms, third_comms):
        out += u64le(len(group)) + b"".join(comm_serialize(c) for c in group)
    out += u64le(len(evaluations)) + b"".join(fr_bytes(x) for x in evaluations)
    out += u64le(3) + b"\x00" * 3                                        # three EmptyMessage = Option::None
    out += u64le(2)
    for w, v in (proof_beta, proof_gamma):
        out += g1_compressed(w) + (b"\x00" if v is None else b"\x01" + fr_bytes(v))
    out += b"\x00"                                                       # evals: None
    return out


# ------------------------------------------------------------------------------------------------
def self_check():
    """constants and conventions this file relies on"""
    # G2: B2 = 1/u, the cofactor recalled from ark-bls12-377 is the order of the right twist divided by r
    assert f2_mul(G2_B, (0, 1)) == (1, 0)
    assert G2_COFACTOR in G2_COFACTOR_CANDIDATES, "recalled G2 cofactor is not #E'(Fq2)/r"
    rng = G.ChaChaRng(bytes(range(32)), 20)
    for _ in range(2):
        pt = g2_rand(rng)
        assert E2.on_curve(pt) and E2.mul(pt, P) is None and pt is not None
    pt = g1_rand(rng)
    assert E1.on_curve(pt) and E1.mul(pt, P) is None
    # polynomial helpers
    p = [3, 1, 4, 1, 5, 9, 2, 6, 5]
    q, r = pdivmod_vanishing(p, 4)
    assert padd(pmul(q, [P - 1, 0, 0, 0, 1]), r) == pnorm(p)
    z = 123456789
    assert padd(pmul(pdiv_linear(p, z), [P - z, 1]), [peval(p, z)]) == pnorm(p)
    d = Domain(8)
    assert d.ifft(d.fft(p[:8])) == pnorm(p[:8])
    big_a, big_b = [pow(3, i, P) for i in range(70)], [pow(5, i, P) for i in range(90)]
    school = [0] * 159
    for i, x in enumerate(big_a):
        for j, y in enumerate(big_b):
            school[i + j] += x * y
    assert pmul(big_a, big_b) == pnorm(school)
    tau = 987654321
    assert sum(l * peval(p[:8], e) for l, e in zip(d.lagrange_coefficients(tau), d.elements())) % P == peval(p[:8], tau)
    return True


if __name__ == "__main__":
    self_check()
    print("golden_marlin.py self-check OK")
