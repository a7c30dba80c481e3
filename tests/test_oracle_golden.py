"""CPU: the C oracle (oracle/liboracle.so) against the committed big-int fixtures in tests/golden/
and against oracle/golden.py on fresh seeded inputs.  This is what pins the oracle."""
import json
import os
import random

import numpy as np
import pytest

from oracle import golden as G
from oracle import pyoracle as O


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def hx(s):
    return int(s, 16)


def test_golden_self_check():
    assert G.self_check()


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_kats(golden_dir, field):
    kats = _load(golden_dir, "field.json")[field]
    mont = O.fr_mont if field == "fr" else O.fq_mont
    unmont = O.fr_unmont if field == "fr" else O.fq_unmont
    mulv = O.fr_mul_vec if field == "fr" else O.fq_mul_vec
    a = mont([hx(k["a"]) for k in kats])
    b = mont([hx(k["b"]) for k in kats])
    assert unmont(mulv(a, b)) == [hx(k["mul"]) for k in kats]
    lib = O.lib()
    r = np.zeros_like(a)
    for i, k in enumerate(kats):
        getattr(lib, f"orc_{field}_add")(O._p(r[i:i + 1]), O._p(a[i:i + 1]), O._p(b[i:i + 1]))
        assert unmont(r[i:i + 1])[0] == hx(k["add"])
        getattr(lib, f"orc_{field}_sub")(O._p(r[i:i + 1]), O._p(a[i:i + 1]), O._p(b[i:i + 1]))
        assert unmont(r[i:i + 1])[0] == hx(k["sub"])
        if hx(k["a"]):
            getattr(lib, f"orc_{field}_inv")(O._p(r[i:i + 1]), O._p(a[i:i + 1]))
            assert unmont(r[i:i + 1])[0] == hx(k["inv_a"])


def test_montgomery_representation_matches_arkworks_layout():
    # Fr::one() in memory is R mod r
    one = O.fr_mont([1])
    assert O.limbs_to_ints(one)[0] == G.FR_MONT_R
    big = np.zeros((1, 4), dtype=np.uint64)
    O.lib().orc_fr_to_canon(O._p(big), O._p(one))
    assert O.limbs_to_ints(big)[0] == 1


def test_batch_inverse_skips_zeros():
    rnd = random.Random(5)
    vals = [rnd.randrange(G.R_MOD) for _ in range(50)]
    vals[0] = vals[17] = vals[49] = 0
    out = O.fr_unmont(O.fr_batch_inverse(O.fr_mont(vals)))
    assert out == [G.fr_inv(v) if v else 0 for v in vals]


def test_g1_generator_multiples(golden_dir):
    g = O.g1_generator()
    assert O.points_from_affine(g)[0] == G.G1_GEN
    for k in _load(golden_dir, "g1.json")["mul_gen"]:
        jac = O.g1_mul(g, hx(k["k"]))
        got = O.points_from_jacobian(jac)[0]
        want = None if k["inf"] else (hx(k["x"]), hx(k["y"]))
        assert got == want
        aff = O.g1_to_affine(jac)
        assert O.points_from_affine(aff)[0] == want
        assert O.lib().orc_g1_affine_on_curve(O._p(aff)) == 1


def _msm_case_arrays(case):
    pts = [None if b is None else (hx(b[0]), hx(b[1])) for b in case["bases"]]
    bases = O.affine_from_points(pts)
    scalars = O.ints_to_limbs([hx(s) for s in case["scalars"]], 4)
    want = None if case["result"] is None else (hx(case["result"][0]), hx(case["result"][1]))
    return bases, scalars, want


def test_msm_golden(golden_dir):
    for case in _load(golden_dir, "msm.json")["cases"]:
        bases, scalars, want = _msm_case_arrays(case)
        for threads in (1, 0):
            got = O.points_from_jacobian(O.msm_variable_base(bases, scalars, threads))[0]
            assert got == want, case["tag"]


def test_msm_random_vs_bigint():
    rnd = random.Random(11)
    n = 200
    ks = [rnd.randrange(1, G.R_MOD) for _ in range(n)]
    g = O.g1_generator()
    jac = np.concatenate([O.g1_mul(g, k) for k in ks])
    bases = O.g1_batch_normalize(jac)
    scalars = [rnd.randrange(G.R_MOD) for _ in range(n)]
    got = O.points_from_jacobian(O.msm_variable_base(bases, O.ints_to_limbs(scalars, 4)))[0]
    want = G.g1_mul(G.G1_GEN, sum(k * s for k, s in zip(ks, scalars)) % G.R_MOD)
    assert got == want


def test_msm_empty():
    out = O.msm_variable_base(np.zeros((0, 13), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert O.points_from_jacobian(out)[0] is None


def test_fixed_base_powers():
    beta = 0x1234567890ABCDEF1234567890ABCDEF
    g = O.g1_generator()
    gj = O.g1_mul(g, 1)
    for n in (1, 5, 40):
        out = O.points_from_affine(O.fixed_base_powers(gj, O.fr_mont([beta]), n))
        assert out == [G.g1_mul(G.G1_GEN, pow(beta, i, G.R_MOD)) for i in range(n)]


def test_ntt_golden(golden_dir):
    data = _load(golden_dir, "ntt.json")
    for k, v in data["domain_gen"].items():
        w = np.zeros((1, 4), dtype=np.uint64)
        O.lib().orc_domain_generator(O._p(w), int(k))
        assert O.fr_unmont(w)[0] == hx(v)
    for case in data["cases"]:
        log_n = case["log_n"]
        x = O.fr_mont([hx(s) for s in case["input"]])
        for inverse in (0, 1):
            for coset in (0, 1):
                got = O.fr_unmont(O.ntt(x, log_n, bool(inverse), bool(coset)))
                assert got == [hx(s) for s in case["inv%d_coset%d" % (inverse, coset)]]


def test_ntt_vs_python_fft_and_roundtrip():
    rnd = random.Random(3)
    for log_n in (8, 11):
        v = [rnd.randrange(G.R_MOD) for _ in range(1 << log_n)]
        x = O.fr_mont(v)
        for coset in (False, True):
            y = O.ntt(x, log_n, False, coset)
            assert O.fr_unmont(y) == G.fft_fast(v, log_n, False, coset)
            assert np.array_equal(O.ntt(y, log_n, True, coset), x)
