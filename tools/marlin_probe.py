#!/usr/bin/env python3
"""Time setup / index / prove / verify of the synthetic mul-chain circuit on the GPU engine and
(optionally) the CPU arm.  python tools/marlin_probe.py 14,16 [--cpu]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleworks_b200.binding import Backend, ConstraintSystem, Marlin, Rng  # noqa: E402


def fr_mont(v):
    R = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
    m = v * (1 << 256) % R
    return np.array([[(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)


def main():
    logs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "12,14").split(",")]
    do_cpu = "--cpu" in sys.argv
    be = Backend(0)
    m = Marlin(be)
    for lg in logs:
        n = (1 << lg) - 2                 # constraints; variables = n + 3 -> |H| = 2^lg
        bounds = (1 << lg, 1 << lg, 3 << lg)
        rng = Rng()
        t0 = time.perf_counter(); srs = m.generate_universal_srs(*bounds, rng); t1 = time.perf_counter()
        cs = ConstraintSystem.builtin("mul-chain", n, 3, 5); t2 = time.perf_counter()
        pk, vk = m.generate_proving_and_verifying_keys(srs, cs); t3 = time.perf_counter()
        proof = m.generate_proof(cs, pk, rng); t4 = time.perf_counter()
        proof2 = m.generate_proof(cs, pk, rng); t5 = time.perf_counter()
        more = []
        for _ in range(int(os.environ.get("PROBE_EXTRA_PROOFS", "0"))):
            ta = time.perf_counter(); m.generate_proof(cs, pk, rng); more.append(time.perf_counter() - ta)
        t5b = time.perf_counter()
        ok = m.verify_proof(vk, fr_mont(3), proof); t6 = time.perf_counter()
        t5 = t5 if not more else t5
        if more:
            print("  further proofs: " + " ".join(f"{x:.3f}s" for x in more), flush=True)
        t6 = t5 + (t6 - t5b)
        print(f"GPU 2^{lg}: setup {t1-t0:.3f}s synth {t2-t1:.3f}s index {t3-t2:.3f}s prove {t4-t3:.3f}s prove#2 {t5-t4:.3f}s "
              f"verify {t6-t5:.3f}s ok={ok} proof={len(proof)}B", flush=True)
        if do_cpu:
            from oracle import pymarlin as C
            crng = C.Rng()
            t0 = time.perf_counter(); csrs = C.universal_setup(*bounds, crng); t1 = time.perf_counter()
            ccs = C.R1cs("chain", size=n, v0=3, v1=5)
            t2 = time.perf_counter(); cpk, cvk = C.index(csrs, ccs); t3 = time.perf_counter()
            cproof = C.prove(cpk, ccs, crng); t4 = time.perf_counter()
            print(f"CPU 2^{lg}: setup {t1-t0:.3f}s index {t3-t2:.3f}s prove {t4-t3:.3f}s same_bytes={cproof == proof}", flush=True)


if __name__ == "__main__":
    main()
