#!/usr/bin/env python3
"""Golden vectors of the protocol layer: proof and verifying-key bytes of the reference's toy circuits
under the fixed test_rng() seed, produced by the INDEPENDENT python restatement oracle/golden_marlin.py
(python integers only; it shares no code with simpleworks_b200/csrc/marlin/marlin.hpp).

    python tests/golden/make_marlin_golden.py          -> tests/golden/marlin_proofs.json

Each case follows the reference's example tests (examples/manual-constraints.rs:86-100,
examples/test-circuit.rs:72-81): ONE test_rng() stream for universal_setup(100, 25, 300) and for proving.
Both engines of the C++ protocol code (the CPU arm oracle/pymarlin.py and the CUDA library) must reproduce
exactly these bytes: tests/test_marlin_cpu.py, tests/test_gpu_marlin.py.  The generator itself checks the CPU
arm against the python bytes before writing.

These are not arkworks outputs -- no Rust toolchain exists here (parity unpinned); if
real arkworks dumps ever become available they replace this file unchanged in format.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import golden as G  # noqa: E402
from oracle import golden_marlin as M  # noqa: E402

R_INV = pow(G.FR_MONT_R, -1, G.R_MOD)      # Fr::new(BigInteger256::new([1, 0, 0, 0])): the raw Montgomery word 1

# tag, universal_setup bounds, circuit kind, arguments, whether the full bytes are stored (else sha256 only)
CASES = [
    ("manual_constraints_a1_b1", (100, 25, 300), "manual", dict(v0=1, v1=1), True),
    # examples/manual-constraints.rs:90-99 as written: a = b = Fr::new(BigInteger256::new([1, 0, 0, 0])) = R^-1 mod r
    ("manual_constraints_reference_values", (100, 25, 300), "manual", dict(v0=R_INV, v1=R_INV), True),
    ("uint8_equality_1_1", (100, 25, 300), "uint8_eq", dict(v0=1, v1=1), True),
    ("uint8_equality_200_200", (100, 25, 300), "uint8_eq", dict(v0=200, v1=200), False),
    ("mul_chain_20", (100, 25, 300), "chain", dict(size=20, v0=3, v1=5), True),
    ("mul_chain_1000", (1 << 10, 1 << 10, 3 << 10), "chain", dict(size=1000, v0=7, v1=11), False),
]


def python_circuit(kind, kw):
    if kind == "manual":
        return M.circuit_manual_constraints(kw["v0"], kw["v1"])
    if kind == "uint8_eq":
        return M.circuit_uint8_equality(kw["v0"], kw["v1"])
    return M.circuit_mul_chain(kw["size"], kw["v0"], kw["v1"])


def run_python(bounds, kind, kw):
    rng = G.test_rng()                              # generate_rand(): one stream for setup and proving
    srs = M.universal_setup(*bounds, rng)
    cs = python_circuit(kind, kw)
    pk, vk = M.index(srs, cs)
    return M.prove(pk, cs, rng), M.vk_serialize(vk)


def cpu_arm_circuit(C, kind, kw):
    """the same instance on the C++ CPU arm: built-ins where they can express it, the generic ABI otherwise"""
    if kind == "manual" and max(kw["v0"], kw["v1"]) >= 1 << 64:
        cs = python_circuit(kind, kw)
        return C.R1cs.custom(cs.num_instance, cs.num_witness, list(zip(cs.a, cs.b, cs.c)), cs.instance, cs.witness)
    return C.R1cs(kind, **kw)


def run_cpu_arm(bounds, kind, kw):
    from oracle import pymarlin as C
    rng = C.Rng()
    srs = C.universal_setup(*bounds, rng)
    cs = cpu_arm_circuit(C, kind, kw)
    pk, vk = C.index(srs, cs)
    return C.prove(pk, cs, rng), C.vk_serialize(vk)


def main():
    out = {"_comment": "proof / verifying-key bytes from oracle/golden_marlin.py (independent python restatement); "
                       "see make_marlin_golden.py", "cases": {}}
    for tag, bounds, kind, kw, full in CASES:
        proof, vkb = run_python(bounds, kind, kw)
        cproof, cvkb = run_cpu_arm(bounds, kind, kw)
        assert proof == cproof, f"{tag}: CPU arm proof differs from the python oracle"
        assert vkb == cvkb, f"{tag}: CPU arm verifying key differs from the python oracle"
        case = {"bounds": list(bounds), "kind": kind, "args": kw, "proof_len": len(proof),
                "proof_sha256": hashlib.sha256(proof).hexdigest(), "vk_len": len(vkb),
                "vk_sha256": hashlib.sha256(vkb).hexdigest(), "proof_head_hex": proof[:48].hex()}
        if full:
            case["proof_hex"] = proof.hex()
            case["vk_hex"] = vkb.hex()
        out["cases"][tag] = case
        print(tag, "ok", case["proof_sha256"][:16], case["vk_sha256"][:16], flush=True)
    json.dump(out, open(os.path.join(HERE, "marlin_proofs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
