#!/usr/bin/env python3
"""Writes tests/golden/*.json from oracle/golden.py (python big ints only -- independent of the
C oracle and of the CUDA library, which are both checked AGAINST these files).

The reference holds no golden vectors for this path, so these are
manufactured pins: unique mathematical values (field products, affine group elements, DFT
values, affine MSM results) and RNG word streams from the published ChaCha definition.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import golden as G  # noqa: E402


def hx(v):
    return "%x" % v


def main():
    rnd = random.Random(0x5357423230300001)
    # --- field KATs (values given as canonical integers) ---
    edge_r = [0, 1, 2, G.R_MOD - 1, G.R_MOD - 2, G.FR_MONT_R, (1 << 252), G.FR_ROOT_OF_UNITY]
    edge_q = [0, 1, 2, G.Q_MOD - 1, G.Q_MOD - 2, G.FQ_MONT_R, (1 << 376), G.G1_X]
    fr_pairs = [(a, b) for a in edge_r for b in edge_r[:4]] + [
        (rnd.randrange(G.R_MOD), rnd.randrange(G.R_MOD)) for _ in range(64)]
    fq_pairs = [(a, b) for a in edge_q for b in edge_q[:4]] + [
        (rnd.randrange(G.Q_MOD), rnd.randrange(G.Q_MOD)) for _ in range(64)]
    field = {
        "fr": [{"a": hx(a), "b": hx(b), "mul": hx(a * b % G.R_MOD), "add": hx((a + b) % G.R_MOD),
                "sub": hx((a - b) % G.R_MOD), "inv_a": hx(G.fr_inv(a) if a else 0)} for a, b in fr_pairs],
        "fq": [{"a": hx(a), "b": hx(b), "mul": hx(a * b % G.Q_MOD), "add": hx((a + b) % G.Q_MOD),
                "sub": hx((a - b) % G.Q_MOD), "inv_a": hx(G.fq_inv(a) if a else 0)} for a, b in fq_pairs],
        "fr_mont_R": hx(G.FR_MONT_R), "fq_mont_R": hx(G.FQ_MONT_R),
    }
    json.dump(field, open(os.path.join(HERE, "field.json"), "w"), indent=0)

    # --- G1 scalar multiples of the generator ---
    ks = [1, 2, 3, 5, 0xFFFF, (1 << 64) + 1, G.R_MOD - 1, G.R_MOD, rnd.randrange(G.R_MOD), rnd.randrange(G.R_MOD)]
    g1 = []
    for k in ks:
        p = G.g1_mul(G.G1_GEN, k)
        g1.append({"k": hx(k), "inf": p is None, "x": hx(p[0]) if p else "0", "y": hx(p[1]) if p else "0"})
    json.dump({"mul_gen": g1}, open(os.path.join(HERE, "g1.json"), "w"), indent=0)

    # --- MSM cases: bases k_i*G, scalars with the arkworks special cases (0, 1), r-1, dupes ---
    msm = []
    for n, tag in [(1, "single"), (7, "tiny"), (33, "window-rule-switch"), (96, "mixed")]:
        bks = [rnd.randrange(1, G.R_MOD) for _ in range(n)]
        bases = [G.g1_mul(G.G1_GEN, k) for k in bks]
        scalars = [rnd.randrange(G.R_MOD) for _ in range(n)]
        if n >= 7:
            scalars[0] = 0
            scalars[1] = 1
            scalars[2] = G.R_MOD - 1
            scalars[3] = 1
            bases[5] = bases[4]                     # duplicate base -> exercises doubling in buckets
            bks[5] = bks[4]
            scalars[5] = scalars[4]
            bases[6] = None                         # identity base
            bks[6] = 0
        if n >= 33:
            for i in range(8, 24):                  # small scalars (boolean-heavy witness polys)
                scalars[i] = rnd.randrange(4)
        expect_k = sum(k * s for k, s in zip(bks, scalars)) % G.R_MOD
        res = G.msm_naive(bases, scalars)
        assert res == G.g1_mul(G.G1_GEN, expect_k)
        msm.append({"tag": tag, "n": n,
                    "bases": [None if b is None else [hx(b[0]), hx(b[1])] for b in bases],
                    "scalars": [hx(s) for s in scalars],
                    "result": None if res is None else [hx(res[0]), hx(res[1])]})
    # cancellation to identity: s*P + (r-s)*P
    p = G.g1_mul(G.G1_GEN, 12345)
    msm.append({"tag": "cancel", "n": 2, "bases": [[hx(p[0]), hx(p[1])]] * 2,
                "scalars": [hx(77), hx(G.R_MOD - 77)], "result": None})
    json.dump({"cases": msm}, open(os.path.join(HERE, "msm.json"), "w"), indent=0)

    # --- NTT cases ---
    ntt = []
    for log_n in (0, 1, 3, 6):
        n = 1 << log_n
        v = [rnd.randrange(G.R_MOD) for _ in range(n)]
        case = {"log_n": log_n, "input": [hx(x) for x in v]}
        for inverse in (False, True):
            for coset in (False, True):
                out = G.dft_naive(v, log_n, inverse, coset)
                assert out == G.fft_fast(v, log_n, inverse, coset)
                case["inv%d_coset%d" % (inverse, coset)] = [hx(x) for x in out]
        ntt.append(case)
    json.dump({"cases": ntt, "domain_gen": {str(k): hx(G.domain_gen(k)) for k in (1, 2, 10, 20, 26, 47)}},
              open(os.path.join(HERE, "ntt.json"), "w"), indent=0)

    # --- RNG streams ---
    r12 = G.test_rng()
    words = [r12.next_u64() for _ in range(40)]          # crosses the 64-word refill at word 32
    r12b = G.test_rng()
    mixed = [r12b.next_u32() for _ in range(63)] + [r12b.next_u64()]   # straddling next_u64
    r12c = G.test_rng()
    frs = [G.fr_rand_mont(r12c) for _ in range(8)]
    r20 = G.ChaChaRng(bytes(range(32)), rounds=20)
    w20 = [r20.next_u32() for _ in range(20)]
    json.dump({"test_rng_u64": [hx(w) for w in words], "test_rng_straddle": [hx(w) for w in mixed],
               "test_rng_fr_rand_mont": [hx(x) for x in frs], "chacha20_seed_0_31_u32": [hx(w) for w in w20],
               "blake2s_abc": G.blake2s(b"abc").hex()},
              open(os.path.join(HERE, "rng.json"), "w"), indent=0)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
