"""GPU parity tests proper: libswb200 (through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and -- at BASELINE sizes -- through
size-independent properties.  Integer work: the bar is bit-exact."""
import json
import os
import random

import numpy as np
import pytest

from oracle import golden as G
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from simpleworks_b200 import build
    from simpleworks_b200.binding import Backend
    build.build()
    b = Backend(0)
    yield b
    b.close()


def _rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 2 ** 63, size=(n, 4), dtype=np.int64).astype(np.uint64) * np.uint64(2) + \
        rs.randint(0, 2, size=(n, 4)).astype(np.uint64)
    a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)          # < 2^252 < r: valid canonical AND Montgomery values
    return a


def _rand_fq(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 2 ** 63, size=(n, 6), dtype=np.int64).astype(np.uint64) * np.uint64(2) + \
        rs.randint(0, 2, size=(n, 6)).astype(np.uint64)
    a[:, 5] &= np.uint64(0x00FFFFFFFFFFFFFF)
    return a


def hx(s):
    return int(s, 16)


# ---------------------------------------------------------------------------------------------
# field arithmetic
# ---------------------------------------------------------------------------------------------
def test_field_golden(be, golden_dir):
    kats = json.load(open(os.path.join(golden_dir, "field.json")))
    for field, mont, unmont, mul, add, sub in (("fr", O.fr_mont, O.fr_unmont, be.fr_mul, be.fr_add, be.fr_sub),
                                               ("fq", O.fq_mont, O.fq_unmont, be.fq_mul, be.fq_add, be.fq_sub)):
        a = be.to_device(mont([hx(k["a"]) for k in kats[field]]))
        b = be.to_device(mont([hx(k["b"]) for k in kats[field]]))
        assert unmont(be.to_host(mul(a, b))) == [hx(k["mul"]) for k in kats[field]]
        assert unmont(be.to_host(add(a, b))) == [hx(k["add"]) for k in kats[field]]
        assert unmont(be.to_host(sub(a, b))) == [hx(k["sub"]) for k in kats[field]]


def test_field_mul_bulk_vs_oracle(be):
    n = 1 << 18
    a, b = _rand_fr(n, 1), _rand_fr(n, 2)
    assert np.array_equal(be.to_host(be.fr_mul(be.to_device(a), be.to_device(b))), O.fr_mul_vec(a, b))
    a, b = _rand_fq(n, 3), _rand_fq(n, 4)
    assert np.array_equal(be.to_host(be.fq_mul(be.to_device(a), be.to_device(b))), O.fq_mul_vec(a, b))


def test_batch_inverse(be):
    a = _rand_fr(1000, 5)
    a[0] = 0
    a[17] = 0
    a[999] = 0
    got = be.to_host(be.fr_batch_inverse_(be.to_device(a)))
    assert np.array_equal(got, O.fr_batch_inverse(a))


# ---------------------------------------------------------------------------------------------
# NTT
# ---------------------------------------------------------------------------------------------
def test_ntt_golden(be, golden_dir):
    data = json.load(open(os.path.join(golden_dir, "ntt.json")))
    for case in data["cases"]:
        log_n = case["log_n"]
        x = O.fr_mont([hx(s) for s in case["input"]])
        for inverse in (0, 1):
            for coset in (0, 1):
                want = [hx(s) for s in case["inv%d_coset%d" % (inverse, coset)]]
                got_host = be.ntt_(x.copy(), log_n, bool(inverse), bool(coset))          # host-buffer ABI
                got_dev = be.to_host(be.ntt_(be.to_device(x), log_n, bool(inverse), bool(coset)))
                assert O.fr_unmont(got_host) == want
                assert np.array_equal(got_dev, got_host)


@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 19])
def test_ntt_vs_oracle_all_modes(be, log_n):
    x = _rand_fr(1 << log_n, 100 + log_n)
    for inverse in (False, True):
        for coset in (False, True):
            got = be.to_host(be.ntt_(be.to_device(x), log_n, inverse, coset))
            assert np.array_equal(got, O.ntt(x, log_n, inverse, coset)), (log_n, inverse, coset)


def test_ntt_batch(be):
    log_n, batch = 10, 5
    x = _rand_fr(batch << log_n, 7)
    got = be.to_host(be.ntt_(be.to_device(x), log_n, False, False, batch=batch))
    for b in range(batch):
        sl = slice(b << log_n, (b + 1) << log_n)
        assert np.array_equal(got[sl], O.ntt(np.ascontiguousarray(x[sl]), log_n))


def test_ntt_large_properties(be):
    """2^22 (a BASELINE sweep size): inverse(forward(x)) == x, coset round trip, linearity, and a
    spot check of 16 outputs against direct evaluation sum_j x_j w^(ij)."""
    import torch
    log_n = 22
    n = 1 << log_n
    x = _rand_fr(n, 42)
    dx = be.to_device(x)
    y = be.ntt_(dx.clone(), log_n)
    assert torch.equal(be.ntt_(y.clone(), log_n, inverse=True), dx)
    yc = be.ntt_(dx.clone(), log_n, coset=True)
    assert torch.equal(be.ntt_(yc.clone(), log_n, inverse=True, coset=True), dx)
    x2 = _rand_fr(n, 43)
    dx2 = be.to_device(x2)
    lhs = be.ntt_(be.fr_add(dx, dx2), log_n)
    rhs = be.fr_add(y, be.ntt_(dx2.clone(), log_n))
    assert torch.equal(lhs, rhs)
    # spot check with the oracle's full transform on a sparse input: x' = x on 64 positions, else 0
    sparse = np.zeros_like(x)
    pos = np.random.RandomState(1).choice(n, 64, replace=False)
    sparse[pos] = x[pos]
    ys = be.to_host(be.ntt_(be.to_device(sparse), log_n))
    w = G.domain_gen(log_n)
    vals = O.fr_unmont(x[pos])
    for i in (0, 1, 2, n // 2, n - 1, 123457):
        want = sum(v * pow(w, (i * int(p)) % n, G.R_MOD) for v, p in zip(vals, pos)) % G.R_MOD
        assert O.fr_unmont(ys[i:i + 1])[0] == want


# ---------------------------------------------------------------------------------------------
# MSM
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def srs_points():
    """2^16 distinct affine points beta^i * G from the oracle's fixed-base routine."""
    g = O.g1_mul(O.g1_generator(), 1)
    beta = O.fr_mont([0x1234567890ABCDEF1234567890ABCDEF1234567])
    return O.fixed_base_powers(g, beta, 1 << 16)


def test_msm_golden(be, golden_dir):
    for case in json.load(open(os.path.join(golden_dir, "msm.json")))["cases"]:
        pts = [None if b is None else (hx(b[0]), hx(b[1])) for b in case["bases"]]
        bases = be.load_bases(O.affine_from_points(pts))
        scalars = O.ints_to_limbs([hx(s) for s in case["scalars"]], 4)
        want = None if case["result"] is None else (hx(case["result"][0]), hx(case["result"][1]))
        try:
            for c in (0, 3, 7):
                be.set_msm_window_bits(c)
                assert O.points_from_jacobian(be.msm(bases, scalars))[0] == want, (case["tag"], c)
                assert O.points_from_jacobian(be.msm(bases, be.to_device(scalars)))[0] == want
        finally:
            be.set_msm_window_bits(0)
            bases.free()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 1000, 1 << 14, 1 << 16])
def test_msm_vs_oracle(be, srs_points, n):
    bases = be.load_bases(srs_points[:n])
    scalars = _rand_fr(n, 200 + n)
    if n >= 31:
        scalars[0] = 0
        scalars[1] = 0
        scalars[1, 0] = 1
        scalars[2] = O.ints_to_limbs([G.R_MOD - 1], 4)[0]
        scalars[5:20, 1:] = 0                          # small scalars
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[:n]), scalars))
    got = O.g1_to_affine(be.msm(bases, scalars))
    assert np.array_equal(got, want)
    got_dev = O.g1_to_affine(be.msm(bases, be.to_device(scalars)))
    assert np.array_equal(got_dev, want)
    # Montgomery-form scalars converted on the device
    mont = O.fr_mont(O.limbs_to_ints(scalars))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, be.to_device(mont), montgomery=True)), want)
    bases.free()


def test_msm_offset_and_truncation(be, srs_points):
    bases = be.load_bases(srs_points[:4096])
    scalars = _rand_fr(1000, 9)
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[100:1100]), scalars))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, scalars, offset=100)), want)
    bases.free()


def test_msm_skewed_and_empty(be, srs_points):
    n = 1 << 14
    bases = be.load_bases(srs_points[:n])
    # Marlin-like: 50% zero, 25% one, 25% uniform (SURVEY 8d distribution M)
    rs = np.random.RandomState(3)
    scalars = _rand_fr(n, 11)
    kind = rs.randint(0, 4, size=n)
    scalars[kind < 2] = 0
    scalars[kind == 2] = 0
    scalars[kind == 2, 0] = 1
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[:n]), scalars))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, scalars)), want)
    # all-zero scalars and n = 0 give the identity
    assert O.points_from_jacobian(be.msm(bases, np.zeros((n, 4), dtype=np.uint64)))[0] is None
    assert O.points_from_jacobian(be.msm(bases, np.zeros((0, 4), dtype=np.uint64)))[0] is None
    # all scalars equal: one bucket per window gets everything
    same = np.repeat(_rand_fr(1, 12), n, axis=0)
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[:n]), same))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, same)), want)
    bases.free()


@pytest.mark.parametrize("n,c", [(1 << 14, 0), (1 << 14, 8), (1 << 16, 0), (1 << 16, 12), (5000, 10)])
def test_msm_window_tables_vs_oracle(be, srs_points, n, c):
    """swb_bases_precompute: same results as the plain path / the oracle, including scalars above
    r/2 (negated), 0, 1, r - 1, the (r +- 1)/2 boundary, offsets and sub-ranges."""
    bases = be.load_bases(srs_points[:n]).precompute(c)
    cb, levels = bases.table_info()
    assert levels == -(-253 // cb) and (c == 0 or cb == c)
    scalars = _rand_fr(n, 900 + n + c)
    scalars[0] = 0
    scalars[1] = O.ints_to_limbs([1], 4)[0]
    scalars[2] = O.ints_to_limbs([G.R_MOD - 1], 4)[0]
    scalars[3] = O.ints_to_limbs([(G.R_MOD - 1) // 2], 4)[0]
    scalars[4] = O.ints_to_limbs([(G.R_MOD + 1) // 2], 4)[0]
    scalars[5] = O.ints_to_limbs([(1 << 252) - 1], 4)[0]
    scalars[6:40, 1:] = 0
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[:n]), scalars))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, scalars)), want)
    mont = O.fr_mont(O.limbs_to_ints(scalars))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, be.to_device(mont), montgomery=True)), want)
    # sub-range with an offset (still large enough to take the table path) and a tiny one (plain path)
    m = n // 2
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[100:100 + m]), scalars[:m]))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, np.ascontiguousarray(scalars[:m]), offset=100)), want)
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[7:7 + 9]), scalars[:9]))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, np.ascontiguousarray(scalars[:9]), offset=7)), want)
    # exported bases are still the originals
    assert np.array_equal(be.export_bases(bases, 0, 64), srs_points[:64])
    bases.free()


def test_msm_window_tables_skew_and_identity(be, srs_points):
    n = 1 << 14
    pts = srs_points[:n].copy()
    pts[5] = O.affine_from_points([None])[0]           # a base at infinity stays at infinity in every level
    bases = be.load_bases(pts).precompute(9)
    rs = np.random.RandomState(5)
    scalars = _rand_fr(n, 13)
    kind = rs.randint(0, 4, size=n)
    scalars[kind < 2] = 0
    scalars[kind == 2] = 0
    scalars[kind == 2, 0] = 1
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(pts), scalars))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, scalars)), want)
    same = np.repeat(_rand_fr(1, 14), n, axis=0)
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(pts), same))
    assert np.array_equal(O.g1_to_affine(be.msm(bases, same)), want)
    assert O.points_from_jacobian(be.msm(bases, np.zeros((n, 4), dtype=np.uint64)))[0] is None
    bases.free()


@pytest.mark.parametrize("tables", [False, True])
def test_msm_repeated_negated_and_infinite_bases(be, srs_points, tables):
    """Bases repeated (P + P inside a bucket needs a doubling), negated (P - P cancels), at infinity, odd
    lengths and position ranges cut mid-run, on the plain path and with window tables."""
    n = 6001
    pts = srs_points[:n].copy()
    ref = O.points_from_affine(pts[:4])
    neg0 = O.affine_from_points([(ref[0][0], G.Q_MOD - ref[0][1])])[0]
    for i in range(40, 80):
        pts[i] = pts[0]                                  # many copies of P ...
    for i in range(80, 100):
        pts[i] = neg0                                    # ... and of -P
    pts[100] = O.affine_from_points([None])[0]
    pts[101] = O.affine_from_points([None])[0]
    scalars = _rand_fr(n, 4242)
    scalars[40:100] = scalars[0]                         # same digits => same buckets, adjacent after the sort
    scalars[100:103] = scalars[0]
    scalars[200:1200] = scalars[200]                     # one long run per window
    want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(pts), scalars))
    bases = be.load_bases(pts)
    if tables:
        bases.precompute(9)
    try:
        assert np.array_equal(O.g1_to_affine(be.msm(bases, scalars)), want)
        for m in (1, 2, 3, 17, 4097):
            w = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(pts[:m]), scalars[:m]))
            assert np.array_equal(O.g1_to_affine(be.msm(bases, np.ascontiguousarray(scalars[:m]))), w), m
    finally:
        bases.free()


@pytest.mark.parametrize("tables", [False, True])
def test_msm_batch_equals_single_calls(be, srs_points, tables):
    """swb_msm_g1_batch_dev: MSMs overlapped on the two slots give the results of single calls, for mixed
    sizes (including empty), offsets and both scalar forms."""
    bases = be.load_bases(srs_points[:20000])
    if tables:
        bases.precompute(10)
    try:
        sizes = [5000, 0, 1, 16384, 777, 12000, 3]
        offs = [0, 5, 100, 1000, 19000, 8000, 19997]
        host = [_rand_fr(n, 3100 + i) for i, n in enumerate(sizes)]
        dev = [be.to_device(h) for h in host]
        got = be.msm_batch(bases, dev, offs)
        for i, (h, o) in enumerate(zip(host, offs)):
            single = be.msm(bases, h, offset=o)
            assert np.array_equal(got[i:i + 1], single), i
        mont = [be.to_device(O.fr_mont(O.limbs_to_ints(h))) if len(h) else be.to_device(h) for h in host]
        assert np.array_equal(be.msm_batch(bases, mont, offs, montgomery=True), got)
        want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[1000:1000 + 16384]), host[3]))
        assert np.array_equal(O.g1_to_affine(got[3:4]), want)
    finally:
        bases.free()


def test_msm_randomised_shapes_vs_oracle(be, srs_points):
    """40 seeded random cases: size, offset, forced window width, window tables or not, and a mix of scalar
    kinds (0, 1, r - 1, small, top-heavy, uniform) -- each against the oracle's Pippenger."""
    rs = np.random.RandomState(20261017)
    special = [0, 1, 2, G.R_MOD - 1, G.R_MOD - 2, (G.R_MOD - 1) // 2, (G.R_MOD + 1) // 2, (1 << 252), (1 << 252) - 1, 1 << 128]
    for case in range(40):
        n = int(rs.choice([1, 2, 3, 5, 16, 17, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 9000]))
        off = int(rs.randint(0, 200))
        scalars = _rand_fr(n, 5000 + case)
        kind = rs.randint(0, 6, size=n)
        scalars[kind == 0] = 0
        for i in np.nonzero(kind == 1)[0]:
            scalars[i] = O.ints_to_limbs([special[int(rs.randint(len(special)))]], 4)[0]
        scalars[kind == 2, 1:] = 0                                   # 64-bit scalars
        scalars[kind == 3, :3] = 0                                   # only the top limb
        pts = np.ascontiguousarray(srs_points[off:off + n])
        want = O.g1_to_affine(O.msm_variable_base(pts, scalars))
        tables = bool(rs.randint(2))
        bases = be.load_bases(srs_points[:off + n + 7])
        try:
            if tables:
                bases.precompute(int(rs.choice([0, 8, 9, 11, 13])))
            else:
                be.set_msm_window_bits(int(rs.choice([0, 2, 3, 5, 8, 11, 14])))
            got = O.g1_to_affine(be.msm(bases, scalars, offset=off))
            assert np.array_equal(got, want), (case, n, off, tables)
        finally:
            be.set_msm_window_bits(0)
            bases.free()


def test_ntt_randomised_round_trips_and_linearity(be):
    """Seeded random sizes and modes: inverse(forward(x)) == x in all four (inverse, coset) pairings,
    forward is linear, and a random subset of outputs equals the direct evaluation x(w^i) / x(g w^i)."""
    import torch
    rs = np.random.RandomState(7)
    for case in range(12):
        log_n = int(rs.randint(1, 21))
        n = 1 << log_n
        x = _rand_fr(n, 9000 + case)
        dx = be.to_device(x)
        coset = bool(rs.randint(2))
        y = be.ntt_(dx.clone(), log_n, coset=coset)
        assert torch.equal(be.ntt_(y.clone(), log_n, inverse=True, coset=coset), dx), (log_n, coset)
        x2 = _rand_fr(n, 9100 + case)
        lhs = be.ntt_(be.fr_add(dx, be.to_device(x2)), log_n, coset=coset)
        assert torch.equal(lhs, be.fr_add(y, be.ntt_(be.to_device(x2), log_n, coset=coset)))
        if log_n <= 12:                                             # direct evaluation with python ints
            w = G.domain_gen(log_n)
            xs = O.fr_unmont(x)
            ys = O.fr_unmont(be.to_host(y))
            shift = G.FR_GENERATOR if coset else 1
            for i in rs.choice(n, min(n, 4), replace=False):
                pt = shift * pow(w, int(i), G.R_MOD) % G.R_MOD
                acc = 0
                for c in reversed(xs):
                    acc = (acc * pt + c) % G.R_MOD
                assert ys[int(i)] == acc, (log_n, coset, i)


def test_fixed_base_powers_vs_oracle(be):
    g = O.g1_mul(O.g1_generator(), 5)
    beta = O.fr_mont([0xDEADBEEFCAFEBABE1234])
    for n in (1, 33, 3000):
        assert np.array_equal(be.fixed_base_powers(g, beta, n), O.fixed_base_powers(g, beta, n))


def test_msm_large_properties(be):
    """2^20 points (BASELINE sweep size): bases built on the GPU, spot-checked against the oracle;
    MSM(all) == MSM(first half) + MSM(second half); scaling all scalars by 2 doubles the result;
    unit scalars give the plain sum; result equals the oracle's Pippenger."""
    n = 1 << 20
    g = O.g1_mul(O.g1_generator(), 1)
    beta = O.fr_mont([0x5357423230300001])
    pts = be.fixed_base_powers(g, beta, n)
    ref = O.fixed_base_powers(g, beta, 4096)
    assert np.array_equal(pts[:4096], ref)
    bases = be.load_bases(pts)
    scalars = _rand_fr(n, 77)
    scalars[:, 3] &= np.uint64(0x07FFFFFFFFFFFFFF)      # < 2^251 so that 2*s stays below r
    full = be.msm(bases, scalars)
    h = n // 2
    lo = be.msm(bases, np.ascontiguousarray(scalars[:h]))
    hi = be.msm(bases, np.ascontiguousarray(scalars[h:]), offset=h)
    assert np.array_equal(be.g1_sum(np.concatenate([lo, hi])), full)
    ints2 = np.zeros_like(scalars)
    carry = np.zeros(n, dtype=np.uint64)
    for j in range(4):
        ints2[:, j] = (scalars[:, j] << np.uint64(1)) | carry
        carry = scalars[:, j] >> np.uint64(63)
    assert np.array_equal(be.msm(bases, ints2), be.g1_sum(np.concatenate([full, full])))
    assert np.array_equal(O.g1_to_affine(full), O.g1_to_affine(O.msm_variable_base(pts, scalars)))
    bases.free()


# ---------------------------------------------------------------------------------------------
# the configurations bench.py times: wide windows (three radix-sort passes from c = 21 on), one shared
# bucket set with W = 11 / 13 table levels, and NTTs beyond 2^24
# ---------------------------------------------------------------------------------------------
def _uniform_mod_r(n, seed):
    """canonical scalars uniform in [0, r): 253-bit draws, rejected when >= r (SURVEY 8d distribution U)"""
    rs = np.random.Generator(np.random.PCG64(seed))
    a = rs.integers(0, 2 ** 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)
    r_limbs = O.ints_to_limbs([G.R_MOD], 4)[0]
    while True:
        ge = np.ones(n, dtype=bool)                      # lexicographic compare from the top limb down
        decided = np.zeros(n, dtype=bool)
        for j in (3, 2, 1, 0):
            lt = ~decided & (a[:, j] < r_limbs[j])
            gt = ~decided & (a[:, j] > r_limbs[j])
            ge[lt] = False
            decided |= lt | gt
        bad = np.nonzero(ge)[0]
        if len(bad) == 0:
            return a
        a[bad] = rs.integers(0, 2 ** 64, size=(len(bad), 4), dtype=np.uint64)
        a[bad, 3] &= np.uint64((1 << 61) - 1)


@pytest.mark.parametrize("c", [21, 23])
def test_msm_wide_window_tables_vs_oracle(be, c):
    """The headline path of bench.py: window tables with c = 23 (W = 11) -- and c = 21 (W = 13), what four
    GPUs use -- one shared bucket set, three-pass radix sort; 2^22 points against the oracle's Pippenger,
    uniform scalars mod r plus the edge values, on the table path and (same handle) forced onto the plain one."""
    n = 1 << 22
    g = O.g1_mul(O.g1_generator(), 1)
    beta = O.fr_mont([0x5357423230300001])
    bases = be.bases_from_powers(g, beta, n)
    try:
        host = be.export_bases(bases, 0, n)
        scalars = _uniform_mod_r(n, 2300 + c)
        for i, v in enumerate([0, 1, G.R_MOD - 1, (G.R_MOD - 1) // 2, (G.R_MOD + 1) // 2, (1 << 252) - 1, 1 << 252]):
            scalars[i] = O.ints_to_limbs([v], 4)[0]
        scalars[100:164, 1:] = 0
        want = O.g1_to_affine(O.msm_variable_base(host, scalars))
        bases.precompute(c)
        assert bases.table_info() == (c, -(-253 // c))
        dev = be.to_device(scalars)
        be.profile(True)
        got = be.msm(bases, dev)
        assert "accumulate" in be.last_stages()
        assert np.array_equal(O.g1_to_affine(got), want)
        assert np.array_equal(be.msm(bases, scalars), got)                       # host-buffer entry point
        be.set_msm_table_policy(-1)
        assert np.array_equal(be.msm(bases, dev), got)                          # plain path, automatic width
        be.set_msm_window_bits(c)
        assert np.array_equal(be.msm(bases, dev), got)                          # plain path at the same width (3-pass sort per window)
        be.set_msm_window_bits(0)
        # a sub-range that the automatic rule would send down the plain path, forced through the tables
        m = 1 << 16
        be.set_msm_table_policy(1)
        sub = be.msm(bases, np.ascontiguousarray(scalars[:m]), offset=12345)
        be.set_msm_table_policy(0)
        want_sub = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(host[12345:12345 + m]), np.ascontiguousarray(scalars[:m])))
        assert np.array_equal(O.g1_to_affine(sub), want_sub)
    finally:
        be.set_msm_window_bits(0)
        be.set_msm_table_policy(0)
        be.profile(False)
        bases.free()


def test_trim_keeps_results(be, srs_points):
    """swb_trim: scratch arenas, vector cache and the NTT twiddle table go back to the driver; the same calls then
    give the same bytes (everything is re-allocated and the table rebuilt), also when nothing was allocated yet."""
    import torch
    n = 5000
    bases = be.load_bases(srs_points[:n])
    scalars = _uniform_mod_r(n, 99)
    x = be.to_device(_rand_fr(1 << 13, 7))
    try:
        be.trim()
        want_msm = be.msm(bases, scalars)
        want_ntt = be.ntt_(x.clone(), 13)
        free0 = torch.cuda.mem_get_info()[0]
        be.trim()
        be.trim()
        assert torch.cuda.mem_get_info()[0] >= free0
        assert np.array_equal(be.msm(bases, scalars), want_msm)
        assert torch.equal(be.ntt_(x.clone(), 13), want_ntt)
        assert torch.equal(be.ntt_(be.ntt_(x.clone(), 13, coset=True), 13, inverse=True, coset=True), x)
    finally:
        bases.free()


@pytest.mark.parametrize("log_n", [11, 14, 17, 21, 23])
def test_ntt_twiddle_paths_agree(be, log_n, monkeypatch):
    """Inter-digit twiddles by table lookup (default) and by running powers (SWB_NTT_TABLE_MAX_LOG=0: the path of
    transforms above 2^26 points and of a refused table allocation) give the same bytes, forward, inverse and on the
    coset; odd and even digit plans (a leftover radix-2 stage beside the stage pairs)."""
    import torch
    n = 1 << log_n
    x = torch.randint(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=torch.int64, device=f"cuda:{be.device}")
    x[:, 3] &= 0x0FFFFFFFFFFFFFFF
    outs = {}
    for cap in ("26", "0"):
        monkeypatch.setenv("SWB_NTT_TABLE_MAX_LOG", cap)
        be.trim()                                            # drops the table, so the cap decides again
        f = be.ntt_(x.clone(), log_n)
        outs[cap] = (f, be.ntt_(x.clone(), log_n, inverse=True), be.ntt_(x.clone(), log_n, coset=True),
                     be.ntt_(x.clone(), log_n, inverse=True, coset=True))
        assert torch.equal(be.ntt_(f.clone(), log_n, inverse=True), x)
    for a, b in zip(outs["26"], outs["0"]):
        assert torch.equal(a, b)
    monkeypatch.delenv("SWB_NTT_TABLE_MAX_LOG")
    be.trim()


@pytest.mark.parametrize("log_n", [25, 26])
def test_ntt_four_pass_sizes(be, log_n):
    """log n >= 25 (more passes than anything the oracle comparison reaches): inverse(forward(x)) == x with and
    without the coset shift, and outputs of a 64-sparse input against direct evaluation with python integers."""
    import torch
    n = 1 << log_n
    rs = np.random.RandomState(log_n)
    x = torch.randint(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=torch.int64, device=f"cuda:{be.device}")
    x[:, 3] &= 0x0FFFFFFFFFFFFFFF
    y = be.ntt_(x.clone(), log_n)
    assert not torch.equal(y, x)
    assert torch.equal(be.ntt_(y, log_n, inverse=True), x)
    del y
    yc = be.ntt_(x.clone(), log_n, coset=True)
    assert torch.equal(be.ntt_(yc, log_n, inverse=True, coset=True), x)
    del yc
    pos = rs.choice(n, 64, replace=False)
    vals_m = _rand_fr(64, 31 + log_n)
    sparse = torch.zeros_like(x)
    sparse[torch.from_numpy(pos).to(sparse.device)] = be.to_device(vals_m)
    vals = O.fr_unmont(vals_m)
    w = G.domain_gen(log_n)
    idx = [0, 1, 2, 3, n // 2, n // 2 + 1, n - 1, 123457, (1 << 24) + 5, n - (1 << 17) - 3] + [int(i) for i in rs.choice(n, 6)]
    for coset in (False, True):
        ys = be.ntt_(sparse.clone(), log_n, coset=coset)
        got = O.fr_unmont(be.to_host(ys[torch.tensor(idx, device=ys.device)]))
        for i, gv in zip(idx, got):
            if coset:
                want = sum(v * pow(G.FR_GENERATOR, int(p), G.R_MOD) * pow(w, (i * int(p)) % n, G.R_MOD) for v, p in zip(vals, pos)) % G.R_MOD
            else:
                want = sum(v * pow(w, (i * int(p)) % n, G.R_MOD) for v, p in zip(vals, pos)) % G.R_MOD
            assert gv == want, (log_n, coset, i)
        # and back: the inverse restores the sparse input
        assert torch.equal(be.ntt_(ys, log_n, inverse=True, coset=coset), sparse)


@pytest.mark.parametrize("tables", [False, True])
def test_msm_bucket_shards_add_up(be, srs_points, tables):
    """swb_msm_set_bucket_shard: the shares of the bucket ranges of all ranks add up to the full result -- on the
    plain path (marked pairs) and with window tables (pairs compacted in the digits kernel), for uniform scalars and
    for a distribution where almost everything lands in one rank's range; a rank that receives nothing returns the
    identity.  (Run rank by rank on one GPU: the shares do not depend on where they are computed.)"""
    n = 1 << 15
    bases = be.load_bases(srs_points[:n])
    if tables:
        bases.precompute(11)
    try:
        uni = _uniform_mod_r(n, 77)
        small = np.zeros((n, 4), dtype=np.uint64)
        small[:, 0] = np.random.RandomState(5).randint(1, 4, size=n) * 8 + 1  # every digit in rank 0's buckets (b = 0 mod 8)
        small[::97] = uni[::97]
        for scalars in (uni, small):
            dev = be.to_device(scalars)
            full = be.msm(bases, dev)
            want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(srs_points[:n]), scalars))
            assert np.array_equal(O.g1_to_affine(full), want)
            for world in (2, 8):
                parts = []
                for rank in range(world):
                    be.set_msm_bucket_shard(rank, world)
                    parts.append(be.msm(bases, dev))
                    parts.append(be.msm(bases, scalars, offset=0))                 # host-buffer entry point, same share
                be.set_msm_bucket_shard(0, 1)
                parts = np.concatenate(parts)
                assert np.array_equal(parts[0::2], parts[1::2])
                assert np.array_equal(be.g1_sum(parts[0::2]), full), (tables, world)
        # a rank whose buckets stay empty: the value 3 only has the digit 3 = bucket 2, which is rank 2's
        be.set_msm_bucket_shard(7, 8)
        only_small = np.zeros((n, 4), dtype=np.uint64)
        only_small[:, 0] = 3
        assert O.points_from_jacobian(be.msm(bases, be.to_device(only_small)))[0] is None
    finally:
        be.set_msm_bucket_shard(0, 1)
        bases.free()


def test_msm_bucket_shards_tiny_msms(be, srs_points):
    """MSMs too small to split by bucket (narrow windows): rank 0 returns everything, the other ranks the identity,
    so the shares still add up (the prover's three-term blinder commitments go this way)."""
    for n in (1, 3, 20, 40):
        bases = be.load_bases(srs_points[:n])
        scalars = _uniform_mod_r(n, 500 + n)
        full = be.msm(bases, scalars)
        try:
            for world in (2, 8):
                parts = []
                for rank in range(world):
                    be.set_msm_bucket_shard(rank, world)
                    parts.append(be.msm(bases, scalars))
                assert np.array_equal(be.g1_sum(np.concatenate(parts)), full), (n, world)
        finally:
            be.set_msm_bucket_shard(0, 1)
            bases.free()


@pytest.mark.parametrize("tables", [False, True])
def test_msm_pair_sums_forced(be, srs_points, tables):
    """swb_msm_set_pair_sums(2 / 3 / 5): one, two and four levels of batch-affine pair sums in front of the accumulation, forced on inputs of every
    shape -- uniform, all-equal scalars (one bucket per window holds everything), many zeros and ones, bases that are
    repeated (P + P: x2 == x1, must fall back to the doubling), negated (P - P), at infinity, the point (0, 1) whose
    x coordinate is zero like the identity marker's, odd lengths -- against the oracle and against the pass switched off."""
    n = 6001
    pts = srs_points[:n].copy()
    ref = O.points_from_affine(pts[:4])
    neg0 = O.affine_from_points([(ref[0][0], G.Q_MOD - ref[0][1])])[0]
    for i in range(40, 80):
        pts[i] = pts[0]
    for i in range(80, 100):
        pts[i] = neg0
    pts[100] = O.affine_from_points([None])[0]
    pts[101] = O.affine_from_points([None])[0]
    if not tables:
        pts[102] = O.affine_from_points([(0, 1)])[0]          # on the curve (order 3), x = 0; the table path needs subgroup points
        pts[103] = O.affine_from_points([(0, G.Q_MOD - 1)])[0]
    bases = be.load_bases(pts)
    if tables:
        bases.precompute(9)
    uni = _uniform_mod_r(n, 31337)
    same = np.repeat(_uniform_mod_r(1, 5), n, axis=0)
    runs = uni.copy()
    runs[40:104] = runs[0]                                      # the special bases share every digit: neighbours after the sort
    runs[200:1200] = runs[200]
    marl = uni.copy()
    kind = np.random.RandomState(3).randint(0, 4, size=n)
    marl[kind < 2] = 0
    marl[kind == 2] = 0
    marl[kind == 2, 0] = 1
    try:
        for scalars in (uni, same, runs, marl):
            want = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(pts), scalars))
            for m in (n, 4097, 2, 1):
                sub = np.ascontiguousarray(scalars[:m])
                if m != n:
                    want_m = O.g1_to_affine(O.msm_variable_base(np.ascontiguousarray(pts[:m]), sub))
                else:
                    want_m = want
                be.set_msm_pair_sums(0)
                off = be.msm(bases, be.to_device(sub))
                assert np.array_equal(O.g1_to_affine(off), want_m), m
                for pol in (2, 3, 5):
                    be.set_msm_pair_sums(pol)
                    got = be.msm(bases, be.to_device(sub))
                    assert np.array_equal(got, off), (m, pol)
        # with bucket shards on top
        be.set_msm_pair_sums(4)
        full = be.msm(bases, be.to_device(uni))
        parts = []
        for rank in range(4):
            be.set_msm_bucket_shard(rank, 4)
            parts.append(be.msm(bases, be.to_device(uni)))
        be.set_msm_bucket_shard(0, 1)
        assert np.array_equal(be.g1_sum(np.concatenate(parts)), full)
    finally:
        be.set_msm_pair_sums(1)
        be.set_msm_bucket_shard(0, 1)
        bases.free()
