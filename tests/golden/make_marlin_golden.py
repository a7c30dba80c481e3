#!/usr/bin/env python3
"""Regression fixtures for the protocol layer: proof and verifying-key bytes of the toy circuits under
the fixed test_rng() seed, produced by the CPU arm (oracle/pymarlin.py).

    python tests/golden/make_marlin_golden.py          -> tests/golden/marlin_proofs.json

These pin OUR restatement against itself across refactors (both engines must keep producing exactly
these bytes); they are NOT arkworks outputs -- no Rust toolchain exists here (DESIGN.md section 2:
parity unpinned).  If real arkworks dumps ever become available, they replace this file unchanged in
format.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pymarlin as M  # noqa: E402

CASES = [
    # tag, universal_setup bounds, circuit kind and arguments (reference examples: manual-constraints.rs:89-99,
    # test-circuit.rs:76; the chain is BASELINE config 4 at toy size)
    ("manual_constraints_a1_b1", (100, 25, 300), ("manual", dict(v0=1, v1=1))),
    ("uint8_equality_1_1", (100, 25, 300), ("uint8_eq", dict(v0=1, v1=1))),
    ("mul_chain_20", (100, 25, 300), ("chain", dict(size=20, v0=3, v1=5))),
    ("mul_chain_1000", (1 << 10, 1 << 10, 3 << 10), ("chain", dict(size=1000, v0=7, v1=11))),
]


def run_case(bounds, kind, kw):
    rng = M.Rng()                                  # generate_rand(): one stream for setup and proving
    srs = M.universal_setup(*bounds, rng)
    cs = M.R1cs(kind, **kw)
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    return proof, M.vk_serialize(vk)


def main():
    out = {"_comment": "sha256 of proof / verifying-key bytes from the CPU arm; see make_marlin_golden.py", "cases": {}}
    for tag, bounds, (kind, kw) in CASES:
        proof, vkb = run_case(bounds, kind, kw)
        out["cases"][tag] = {"bounds": list(bounds), "kind": kind, "args": kw, "proof_len": len(proof),
                             "proof_sha256": hashlib.sha256(proof).hexdigest(), "vk_len": len(vkb),
                             "vk_sha256": hashlib.sha256(vkb).hexdigest(), "proof_head_hex": proof[:48].hex()}
    json.dump(out, open(os.path.join(HERE, "marlin_proofs.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
