"""GPU: the protocol level through the C ABI -- simpleworks::marlin's setup / index / prove /
verify with every NTT, MSM and fixed-base table on the B200 -- against the CPU arm (same RNG
streams => byte-identical proofs) and the reference's own round-trip tests."""
import numpy as np
import pytest

from oracle import pymarlin as CPU
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from simpleworks_b200 import build
    from simpleworks_b200.binding import Backend, Marlin
    build.build()
    be = Backend(0)
    yield Marlin(be)
    be.close()


CASES = [("manual-constraints", "manual", 0, 1, 1, [1]),
         ("test-circuit", "uint8_eq", 0, 9, 9, []),
         ("mul-chain", "chain", 50, 3, 5, [3]),
         ("mul-chain", "chain", 3000, 7, 11, [7])]


@pytest.mark.parametrize("gname,cname,size,v0,v1,pub", CASES)
def test_proof_bytes_gpu_equals_cpu_and_verifies(gpu, gname, cname, size, v0, v1, pub):
    from simpleworks_b200.binding import ConstraintSystem, Rng
    nc = max(100, size + 2)
    bounds = (nc, nc, 3 * nc)
    # GPU arm
    rng = Rng()
    srs = gpu.generate_universal_srs(*bounds, rng)
    cs = ConstraintSystem.builtin(gname, size, v0, v1)
    assert cs.is_satisfied()
    pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
    proof = gpu.generate_proof(cs, pk, rng)
    pi = O.fr_mont(pub) if pub else np.zeros((0, 4), dtype=np.uint64)
    assert gpu.verify_proof(vk, pi, proof)
    # CPU arm, same seeds
    crng = CPU.Rng()
    csrs = CPU.universal_setup(*bounds, crng)
    assert CPU.lib().orc_srs_max_degree(csrs) == gpu.srs_max_degree(srs)
    ccs = CPU.R1cs(cname, size=size, v0=v0, v1=v1)
    cpk, cvk = CPU.index(csrs, ccs)
    cproof = CPU.prove(cpk, ccs, crng)
    assert proof == cproof
    # cross-verification and tamper rejection
    assert CPU.verify(cvk, pi, proof)
    assert gpu.verify_proof(vk, pi, cproof)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 0x10
    assert not gpu.verify_proof(vk, pi, bytes(bad))
    if pub:
        assert not gpu.verify_proof(vk, O.fr_mont([pub[0] + 1]), proof)


def test_unsatisfied_instance_is_refused(gpu):
    from simpleworks_b200.binding import ConstraintSystem, Rng, SwbError
    rng = Rng()
    srs = gpu.generate_universal_srs(100, 25, 300, rng)
    good = ConstraintSystem.builtin("test-circuit", 0, 5, 5)
    bad = ConstraintSystem.builtin("test-circuit", 0, 5, 4)
    pk, _ = gpu.generate_proving_and_verifying_keys(srs, good)
    with pytest.raises(SwbError):
        gpu.generate_proof(bad, pk, rng)


def test_custom_constraint_system_through_the_abi(gpu):
    """x * y = z with public z, built row by row like ConstraintSystem::enforce_constraint."""
    from simpleworks_b200.binding import ConstraintSystem, Rng
    one = O.fr_mont([1])[0]
    cs = ConstraintSystem.new(2, 2)                       # instance [1, z], witness [x, y]
    cs.enforce_constraint([(one, 2)], [(one, 3)], [(one, 1)])
    cs.assign(O.fr_mont([1, 35]), O.fr_mont([5, 7]))
    assert cs.is_satisfied()
    rng = Rng()
    srs = gpu.generate_universal_srs(100, 25, 300, rng)
    pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
    proof = gpu.generate_proof(cs, pk, rng)
    assert gpu.verify_proof(vk, O.fr_mont([35]), proof)
    assert not gpu.verify_proof(vk, O.fr_mont([36]), proof)


def test_vk_bytes_and_r1cs_interchange(gpu):
    """Proofs and verifying keys travel as bytes (simple_merkle_tree.rs:122-148); a constraint system
    imported from the SWBR1CS1 interchange format yields the same proof."""
    from simpleworks_b200.binding import ConstraintSystem, Rng
    rng = Rng()
    srs = gpu.generate_universal_srs(100, 25, 300, rng)
    cs = ConstraintSystem.builtin("mul-chain", 20, 2, 9)
    pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
    proof = gpu.generate_proof(cs, pk, Rng())
    vk2 = gpu.deserialize_verifying_key(gpu.serialize_verifying_key(vk))
    assert gpu.verify_proof(vk2, O.fr_mont([2]), proof, Rng())
    assert not gpu.verify_proof(vk2, O.fr_mont([5]), proof)
    cs2 = ConstraintSystem.from_bytes(cs.to_bytes())
    pk2, _ = gpu.generate_proving_and_verifying_keys(srs, cs2)
    assert gpu.generate_proof(cs2, pk2, Rng()) == proof
    # the CPU arm reads the same interchange bytes and agrees byte for byte
    crng = CPU.Rng()
    csrs = CPU.universal_setup(100, 25, 300, crng)
    ccs = CPU.R1cs.from_bytes(cs.to_bytes())
    cpk, cvk = CPU.index(csrs, ccs)
    assert CPU.prove(cpk, ccs, CPU.Rng()) == proof
    assert CPU.vk_serialize(cvk) == gpu.serialize_verifying_key(vk)


def _gpu_circuit(case):
    """the fixture's instance through the C ABI: built-ins where they can express the values, swb_r1cs_new /
    add_constraint / set_assignment otherwise (examples/manual-constraints.rs passes the raw Montgomery word 1)"""
    from oracle import golden_marlin as PM
    from simpleworks_b200.binding import ConstraintSystem
    names = {"manual": "manual-constraints", "uint8_eq": "test-circuit", "chain": "mul-chain"}
    a = case["args"]
    if max(a["v0"], a["v1"]) < 1 << 64:
        return ConstraintSystem.builtin(names[case["kind"]], a.get("size", 0), a["v0"], a["v1"])
    assert case["kind"] == "manual"
    pc = PM.circuit_manual_constraints(a["v0"], a["v1"])
    cs = ConstraintSystem.new(pc.num_instance, pc.num_witness)
    lc = lambda row: [(O.fr_mont([k])[0], j) for k, j in row]
    for ra, rb, rc in zip(pc.a, pc.b, pc.c):
        cs.enforce_constraint(lc(ra), lc(rb), lc(rc))
    cs.assign(O.fr_mont(pc.instance), O.fr_mont(pc.witness))
    return cs


def test_gpu_proofs_match_committed_fixtures(gpu):
    """The GPU engine reproduces tests/golden/marlin_proofs.json -- proof and verifying-key bytes of the reference's
    toy circuits under the fixed test_rng() seed, computed by the independent python restatement
    oracle/golden_marlin.py -- without the CPU arm in the loop."""
    import hashlib
    import json
    import os
    from simpleworks_b200.binding import Rng
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "marlin_proofs.json")))["cases"]
    assert len(g) >= 6
    for tag, case in g.items():
        rng = Rng()
        srs = gpu.generate_universal_srs(*case["bounds"], rng)
        cs = _gpu_circuit(case)
        pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
        proof = gpu.generate_proof(cs, pk, rng)
        assert hashlib.sha256(proof).hexdigest() == case["proof_sha256"], tag
        assert hashlib.sha256(gpu.serialize_verifying_key(vk)).hexdigest() == case["vk_sha256"], tag
        if "proof_hex" in case:
            assert proof.hex() == case["proof_hex"] and gpu.serialize_verifying_key(vk).hex() == case["vk_hex"], tag


def test_tuned_srs_gives_the_same_proofs(gpu):
    """swb_srs_set_tune_after: with window tables built over the SRS powers before the first commitment,
    index and prove return the same bytes as on the plain MSM path (and as the committed fixture)."""
    import hashlib
    import json
    import os
    from simpleworks_b200.binding import ConstraintSystem, Rng
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "marlin_proofs.json")))["cases"]["mul_chain_1000"]
    out = []
    for tune in (0, 1):
        rng = Rng()
        srs = gpu.generate_universal_srs(*case["bounds"], rng)
        gpu.srs_set_tune_after(srs, tune)
        cs = ConstraintSystem.builtin("mul-chain", 1000, 7, 11)
        pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
        proof = gpu.generate_proof(cs, pk, rng)
        assert gpu.verify_proof(vk, O.fr_mont([7]), proof)
        tab_c, tab_w = gpu.srs_table_info(srs)                   # swb_srs_table_info: (0, 0) on the plain path
        assert (tab_w > 0 and tab_c > 0) if tune else (tab_c, tab_w) == (0, 0)
        if tune:
            before = gpu.generate_proof(cs, pk, Rng())
            gpu.be.trim()                                        # swb_trim between proofs: scratch is re-grown, bytes stay
            assert gpu.generate_proof(cs, pk, Rng()) == before
        out.append((proof, gpu.serialize_verifying_key(vk)))
    assert out[0] == out[1]
    assert hashlib.sha256(out[1][0]).hexdigest() == case["proof_sha256"]


@pytest.mark.parametrize("size,terms", [(300, 5), (5000, 3), (70000, 4)])
def test_general_shape_circuit_gpu_equals_cpu(gpu, size, terms):
    """Multi-term rows with repeated columns and three public inputs: the device-resident CSR products
    (k_csr_spmv) and the index must give the CPU arm's bytes."""
    from simpleworks_b200.binding import ConstraintSystem, Rng
    bounds = (2 * size, 2 * size, 2 * terms * size + 2 * size)
    rng = Rng()
    srs = gpu.generate_universal_srs(*bounds, rng)
    cs = ConstraintSystem.builtin("random-sparse", size, terms, 77)
    assert cs.is_satisfied()
    pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
    proof = gpu.generate_proof(cs, pk, rng)
    pub = O.fr_mont(CPU.random_sparse_public_inputs(77))
    assert gpu.verify_proof(vk, pub, proof)
    crng = CPU.Rng()
    csrs = CPU.universal_setup(*bounds, crng)
    ccs = CPU.R1cs("random_sparse", size=size, v0=terms, v1=77)
    cpk, cvk = CPU.index(csrs, ccs)
    assert CPU.prove(cpk, ccs, crng) == proof
    assert gpu.serialize_verifying_key(vk) == CPU.vk_serialize(cvk)


def test_sharded_proving_two_gpus():
    """Multi-GPU proving (swb_set_msm_shard): two ranks, MSMs sharded, NCCL all-gather of the partial results --
    same proof bytes as one GPU.  Needs two visible GPUs (skipped on a single-GPU box)."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(os.path.dirname(__file__), "dist_marlin_check.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", script, "14"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert json.loads(r.stdout.strip().splitlines()[-1])["ok"]


def test_proving_key_bytes_round_trip(gpu):
    """serialize_proving_key / deserialize_proving_key (reference src/marlin/serialization.rs:33-45): a key loaded from
    its bytes -- on a context that never saw the SRS -- proves the same bytes; tampered or truncated bytes are refused."""
    from simpleworks_b200.binding import ConstraintSystem, Rng
    rng = Rng()
    srs = gpu.generate_universal_srs(400, 400, 1200, rng)
    cs = ConstraintSystem.builtin("random-sparse", 300, 3, 41)
    pk, vk = gpu.generate_proving_and_verifying_keys(srs, cs)
    want = gpu.generate_proof(cs, pk, Rng())
    blob = gpu.serialize_proving_key(pk, vk)
    assert blob[:8] == b"SWBPK001" and len(blob) > (gpu.srs_max_degree(srs) + 1) * 96
    vkb = gpu.serialize_verifying_key(vk)
    pk.close(); srs.close()                                        # the loaded key must stand on its own
    pk2, vk2 = gpu.deserialize_proving_key(blob)
    assert gpu.serialize_verifying_key(vk2) == vkb
    assert gpu.generate_proof(cs, pk2, Rng()) == want
    assert gpu.serialize_proving_key(pk2, vk2) == blob
    from simpleworks_b200._lib import SwbError
    for bad in (blob[:-1], blob[:1000], b"SWBPK002" + blob[8:], blob[:200] + bytes([blob[200] ^ 1]) + blob[201:],
                blob[:-40] + bytes([blob[-40] ^ 1]) + blob[-39:]):
        with pytest.raises(SwbError):
            gpu.deserialize_proving_key(bad)


def test_proof_objects(gpu):
    """deserialize_proof / serialize_proof / verify_proof on the object (serialization.rs:5-17, mod.rs:79-86)"""
    import json
    import os
    from simpleworks_b200._lib import SwbError
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "marlin_proofs.json")))["cases"]["manual_constraints_a1_b1"]
    proof_bytes, vk_bytes = bytes.fromhex(case["proof_hex"]), bytes.fromhex(case["vk_hex"])
    vk = gpu.deserialize_verifying_key(vk_bytes)
    proof = gpu.deserialize_proof(proof_bytes)
    assert gpu.serialize_proof(proof) == proof_bytes
    assert gpu.verify_proof_object(vk, O.fr_mont([1]), proof)
    assert not gpu.verify_proof_object(vk, O.fr_mont([2]), proof)
    for bad in (proof_bytes[:-1], proof_bytes + b"\x00", b"", proof_bytes[:16] + b"\xff" * 48 + proof_bytes[64:]):
        with pytest.raises(SwbError):
            gpu.deserialize_proof(bad)
