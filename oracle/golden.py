"""Big-integer golden model for the BLS12-377 Marlin hot path.  TEST INFRASTRUCTURE ONLY.

This file is the slowest, most obviously-correct layer of the parity chain

    golden.py (python ints)  ->  oracle/*.c (C restatement of arkworks)  ->  CUDA kernels

It is imported only by tests/, by tests/golden/make_golden.py (which wrote the committed
fixtures) and by oracle self-checks.  The product (simpleworks_b200/) never imports it.

PARITY UNPINNED against real arkworks: the reference (/root/reference) contains no field, curve,
FFT or MSM code -- those live in the un-vendored crates ark-ff / ark-ec / ark-poly /
ark-bls12-377 ^0.3.0 (reference Cargo.toml:15-27) and the Entropy1729/marlin git fork
(Cargo.toml:29-30), and no Rust toolchain exists in this image.  What IS pinned here:
  * constants re-derived from first principles (primality, orders, Montgomery constants);
  * DFT values and affine MSM results are unique mathematical objects, independent of
    algorithm, so any correct implementation agrees with arkworks on them;
  * the RNG word streams follow the published ChaCha / rand_core::BlockRng definitions.
The call sites that consume these values in the reference are src/marlin/mod.rs:52,75,85,92.
"""
from __future__ import annotations

import hashlib
import struct

# ----------------------------------------------------------------------------------------------
# A.1 constants (ark-bls12-377 0.3 fields/fr.rs, fields/fq.rs, curves/g1.rs)
# ----------------------------------------------------------------------------------------------
R_MOD = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001   # Fr modulus r
Q_MOD = int(
    "01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800"
    "170b5d44300000008508c00000000001", 16)                                    # Fq modulus q
FR_TWO_ADICITY = 47
FR_GENERATOR = 22
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)
FR_MONT_R = (1 << 256) % R_MOD
FQ_MONT_R = (1 << 384) % Q_MOD
G1_X = 81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695
G1_Y = 241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030
G1_COFACTOR = 0x170B5D44300000000000000000000000
G1_B = 1

assert FR_ROOT_OF_UNITY == 8065159656716812877374967518403273466521432693661810619979959746626482506078
assert (G1_Y * G1_Y - G1_X ** 3 - G1_B) % Q_MOD == 0


def fr_inv(a: int) -> int:
    return pow(a, R_MOD - 2, R_MOD)


def fq_inv(a: int) -> int:
    return pow(a, Q_MOD - 2, Q_MOD)


# Montgomery encode/decode: ark_ff::Fp256 stores a*R mod p in 4 LE u64 limbs (A.2)
def fr_to_mont(a: int) -> int:
    return (a * FR_MONT_R) % R_MOD


def fr_from_mont(m: int) -> int:
    return (m * fr_inv(FR_MONT_R)) % R_MOD


def fq_to_mont(a: int) -> int:
    return (a * FQ_MONT_R) % Q_MOD


def fq_from_mont(m: int) -> int:
    return (m * fq_inv(FQ_MONT_R)) % Q_MOD


def limbs_le(x: int, n: int) -> list[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs(ls) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(ls))


# ----------------------------------------------------------------------------------------------
# A.3 G1: y^2 = x^3 + 1 over Fq.  Affine law; None is the identity.
# ----------------------------------------------------------------------------------------------
def g1_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - G1_B) % Q_MOD == 0


def g1_neg(p):
    if p is None:
        return None
    return (p[0], (-p[1]) % Q_MOD)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = (3 * x1 * x1) * fq_inv(2 * y1) % Q_MOD
    else:
        lam = (y2 - y1) * fq_inv((x2 - x1) % Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def g1_mul(p, k: int):
    k %= R_MOD * G1_COFACTOR  # group exponent of E(Fq) divides r*h for prime-order part usage
    acc = None
    add = p
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


G1_GEN = (G1_X, G1_Y)


def msm_naive(bases, scalars):
    """sum_i scalars[i] * bases[i]; truncates to min(len) like
    ark_ec::msm::VariableBaseMSM::multi_scalar_mul (A.4)."""
    acc = None
    for p, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


# ----------------------------------------------------------------------------------------------
# A.5 Radix2EvaluationDomain
# ----------------------------------------------------------------------------------------------
def domain_gen(log_n: int) -> int:
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_TWO_ADICITY - log_n), R_MOD)


def dft_naive(v, log_n: int, inverse: bool = False, coset: bool = False):
    """fft / ifft / coset_fft / coset_ifft of ark_poly::Radix2EvaluationDomain, O(n^2)."""
    n = 1 << log_n
    v = list(v) + [0] * (n - len(v))
    w = domain_gen(log_n)
    g = FR_GENERATOR
    if not inverse:
        if coset:
            v = [(x * pow(g, j, R_MOD)) % R_MOD for j, x in enumerate(v)]
        return [sum(v[j] * pow(w, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]
    wi = fr_inv(w)
    ninv = fr_inv(n)
    out = [sum(v[j] * pow(wi, i * j, R_MOD) for j in range(n)) * ninv % R_MOD for i in range(n)]
    if coset:
        gi = fr_inv(g)
        out = [(x * pow(gi, j, R_MOD)) % R_MOD for j, x in enumerate(out)]
    return out


def fft_fast(v, log_n: int, inverse: bool = False, coset: bool = False):
    """Same function as dft_naive, O(n log n) recursive; cross-checked against it in tests."""
    n = 1 << log_n
    v = list(v) + [0] * (n - len(v))
    g = FR_GENERATOR
    if coset and not inverse:
        s = 1
        for j in range(n):
            v[j] = v[j] * s % R_MOD
            s = s * g % R_MOD
    w = domain_gen(log_n)
    if inverse:
        w = fr_inv(w)

    def rec(a, w):
        m = len(a)
        if m == 1:
            return a
        e = rec(a[0::2], w * w % R_MOD)
        o = rec(a[1::2], w * w % R_MOD)
        out = [0] * m
        t = 1
        for i in range(m // 2):
            x = o[i] * t % R_MOD
            out[i] = (e[i] + x) % R_MOD
            out[i + m // 2] = (e[i] - x) % R_MOD
            t = t * w % R_MOD
        return out

    out = rec(v, w)
    if inverse:
        ninv = fr_inv(n)
        out = [x * ninv % R_MOD for x in out]
        if coset:
            gi = fr_inv(g)
            s = 1
            for j in range(n):
                out[j] = out[j] * s % R_MOD
                s = s * gi % R_MOD
    return out


# ----------------------------------------------------------------------------------------------
# A.2 RNGs: rand_chacha 0.3 ChaCha{12,20}Rng over rand_core::BlockRng (64 u32 words buffered)
# ----------------------------------------------------------------------------------------------
def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def _chacha_block(key_words, counter: int, stream: int, rounds: int):
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF]
    x = st[:]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)]


class ChaChaRng:
    """rand_chacha::ChaCha{8,12,20}Rng: 32-byte seed = key, 64-bit block counter from 0,
    stream 0, four blocks (64 words) generated per refill; next_u64 straddles refills the
    way rand_core::block::BlockRng does."""

    def __init__(self, seed: bytes, rounds: int = 20):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.rounds = rounds
        self.counter = 0
        self.buf: list[int] = []
        self.index = 64

    def _refill(self):
        self.buf = []
        for i in range(4):
            self.buf += _chacha_block(self.key, self.counter + i, 0, self.rounds)
        self.counter += 4
        self.index = 0

    def next_u32(self) -> int:
        if self.index >= 64:
            self._refill()
        v = self.buf[self.index]
        self.index += 1
        return v

    def next_u64(self) -> int:
        if self.index < 63:
            lo = self.buf[self.index]
            hi = self.buf[self.index + 1]
            self.index += 2
            return (hi << 32) | lo
        if self.index >= 64:
            self._refill()
            lo, hi = self.buf[0], self.buf[1]
            self.index = 2
            return (hi << 32) | lo
        lo = self.buf[63]
        self._refill()
        hi = self.buf[0]
        self.index = 1
        return (hi << 32) | lo

    def fill_bytes(self, n: int) -> bytes:
        out = b""
        while len(out) < n:
            out += struct.pack("<I", self.next_u32())
        return out[:n]


TEST_RNG_SEED = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)


def test_rng() -> ChaChaRng:
    """ark_std::test_rng(): StdRng (= ChaCha12Rng in rand 0.8) from the fixed seed; this is what
    simpleworks' generate_rand returns (reference src/marlin/mod.rs:33-35)."""
    return ChaChaRng(TEST_RNG_SEED, rounds=12)


def fr_rand_mont(rng: ChaChaRng) -> int:
    """ark_ff Fp256::rand: rejection-sample raw limbs (top 3 bits shaved) and keep them AS the
    Montgomery representation (A.2).  Returns the raw Montgomery integer."""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= 0xFFFFFFFFFFFFFFFF >> 3
        raw = from_limbs(limbs)
        if raw < R_MOD:
            return raw


def fq_rand_mont(rng: ChaChaRng) -> int:
    while True:
        limbs = [rng.next_u64() for _ in range(6)]
        limbs[5] &= 0xFFFFFFFFFFFFFFFF >> 7
        raw = from_limbs(limbs)
        if raw < Q_MOD:
            return raw


def blake2s(data: bytes) -> bytes:
    return hashlib.blake2s(data, digest_size=32).digest()


# ----------------------------------------------------------------------------------------------
# self-check of the constants (run: python oracle/golden.py)
# ----------------------------------------------------------------------------------------------
def _is_probable_prime(n: int) -> bool:
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def self_check():
    assert _is_probable_prime(R_MOD) and _is_probable_prime(Q_MOD)
    assert R_MOD.bit_length() == 253 and Q_MOD.bit_length() == 377
    assert (R_MOD - 1) % (1 << 47) == 0 and (R_MOD - 1) % (1 << 48) != 0
    assert pow(FR_ROOT_OF_UNITY, 1 << 47, R_MOD) == 1 and pow(FR_ROOT_OF_UNITY, 1 << 46, R_MOD) != 1
    assert (-pow(R_MOD, -1, 1 << 64)) % (1 << 64) == 725501752471715839
    assert (-pow(Q_MOD, -1, 1 << 64)) % (1 << 64) == 9586122913090633727
    assert (-pow(R_MOD, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF
    assert (-pow(Q_MOD, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF
    assert FR_MONT_R == 6014086494747379908336260804527802945383293308637734276299549080986809532403
    assert g1_is_on_curve(G1_GEN) and g1_mul(G1_GEN, R_MOD) is None
    # ChaCha20 known-answer: RFC 7539 2.3.2 uses a 32-bit counter + 96-bit nonce layout; with
    # counter=1 and nonce words (0x09000000, 0x4a000000, 0) the 64-bit-counter/64-bit-stream
    # layout used by rand_chacha coincides when counter = 1 | 0x09000000<<32, stream = 0x4a000000.
    key = bytes(range(32))
    blk = _chacha_block(struct.unpack("<8I", key), 1 | (0x09000000 << 32), 0x4A000000, 20)
    assert blk[0] == 0xE4E7F110 and blk[15] == 0x4E3C50A2
    return True


if __name__ == "__main__":
    self_check()
    print("golden.py self-check OK")
