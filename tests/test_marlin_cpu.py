"""CPU: the Marlin protocol layer (shared templates) on the CPU engine -- RNG / hash pins against
the big-int golden model, then the reference's own Marlin tests restated:
verify(prove(x)) == true for manual-constraints and test-circuit under the (100, 25, 300) SRS,
unsatisfied circuits abort the prover, tampered proofs / inputs are rejected."""
import json
import os

import numpy as np
import pytest

from oracle import golden as G
from oracle import pymarlin as M
from oracle import pyoracle as O


def hx(s):
    return int(s, 16)


def test_rng_streams_match_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "rng.json")))
    r = M.Rng()
    assert [r.next_u64() for _ in range(40)] == [hx(w) for w in g["test_rng_u64"]]
    r = M.Rng()
    got = [r.next_u32() for _ in range(63)] + [r.next_u64()]
    assert got == [hx(w) for w in g["test_rng_straddle"]]
    r = M.Rng()
    assert [r.fr_rand_mont() for _ in range(8)] == [hx(w) for w in g["test_rng_fr_rand_mont"]]
    r = M.Rng(bytes(range(32)), 20)
    assert [r.next_u32() for _ in range(20)] == [hx(w) for w in g["chacha20_seed_0_31_u32"]]


def test_blake2s_matches_hashlib(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "rng.json")))
    assert M.blake2s(b"abc").hex() == g["blake2s_abc"]
    for n in (0, 1, 63, 64, 65, 127, 128, 129, 1000):
        data = bytes((i * 7 + 3) & 0xFF for i in range(n))
        assert M.blake2s(data) == G.blake2s(data)


def test_pairing_tower_and_bilinearity():
    """BLS12-377 pairing used by verify: tower relations, G2 cofactor (derived, not transcribed),
    e(aP, bQ) = e(P, Q)^(ab), non-degeneracy, e^r = 1, pairing-product accept and reject."""
    assert M.lib().orc_pairing_selftest() == 0


@pytest.fixture(scope="module")
def toy_srs():
    rng = M.Rng()                                   # generate_rand()
    srs = M.universal_setup(100, 25, 300, rng)      # examples/manual-constraints.rs:89
    assert M.lib().orc_srs_max_degree(srs) == 1533
    return srs, rng


def test_manual_constraints_prove_verify(toy_srs):
    srs, rng = toy_srs
    cs = M.R1cs("manual", v0=1, v1=1)               # a = b = 1, manual-constraints.rs:90-99
    assert cs.is_satisfied()
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    assert M.verify(vk, O.fr_mont([1]), proof)
    # wrong public input, tampered evaluation, tampered commitment, truncated proof
    assert not M.verify(vk, O.fr_mont([2]), proof)
    bad = bytearray(proof)
    bad[-40] ^= 1
    assert not M.verify(vk, O.fr_mont([1]), bytes(bad))
    bad = bytearray(proof)
    bad[20] ^= 1
    assert not M.verify(vk, O.fr_mont([1]), bytes(bad))
    assert not M.verify(vk, O.fr_mont([1]), proof[:-1])


def test_uint8_equality_prove_verify(toy_srs):
    srs, rng = toy_srs
    cs = M.R1cs("uint8_eq", v0=1, v1=1)             # examples/test-circuit.rs:76
    assert cs.is_satisfied()
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    assert M.verify(vk, np.zeros((0, 4), dtype=np.uint64), proof)
    assert not M.verify(vk, O.fr_mont([1]), proof)   # wrong number of public inputs


def test_unsatisfied_circuit_aborts_prover(toy_srs):
    """examples/schnorr-signature/main.rs:214-217 (#[should_panic]): proving an unsatisfied
    instance aborts instead of producing a proof."""
    srs, rng = toy_srs
    good = M.R1cs("uint8_eq", v0=5, v1=5)
    bad = M.R1cs("uint8_eq", v0=5, v1=4)
    assert not bad.is_satisfied()
    pk, _ = M.index(srs, good)
    with pytest.raises(M.MarlinError):
        M.prove(pk, bad, rng)


def test_proofs_are_deterministic_functions_of_the_rng():
    """SimpleMerkleTree::{prove,verify} create a fresh test_rng() per call
    (src/merkle_tree/simple_merkle_tree.rs:117,146): same seed, same bytes."""
    out = []
    for _ in range(2):
        rng = M.Rng()
        srs = M.universal_setup(100, 25, 300, rng)
        cs = M.R1cs("chain", size=20, v0=3, v1=5)
        pk, vk = M.index(srs, cs)
        out.append(M.prove(pk, cs, M.Rng()))
        assert M.verify(vk, O.fr_mont([3]), out[-1])
    assert out[0] == out[1]


def test_chain_circuit_medium(toy_srs):
    rng = M.Rng()
    srs = M.universal_setup(1 << 10, 1 << 10, 3 << 10, rng)
    cs = M.R1cs("chain", size=1000, v0=7, v1=11)
    assert cs.is_satisfied()
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    assert M.verify(vk, O.fr_mont([7]), proof)
    assert not M.verify(vk, O.fr_mont([8]), proof)


def test_verifying_key_and_r1cs_round_trip(toy_srs):
    """serialize_verifying_key -> deserialize_verifying_key still verifies (the shape of
    src/merkle_tree/simple_merkle_tree.rs:122-148, which ships proofs as bytes); the R1CS interchange
    format reproduces the same proof."""
    srs, _ = toy_srs
    cs = M.R1cs("chain", size=30, v0=2, v1=9)
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, M.Rng())
    vk2 = M.vk_deserialize(M.vk_serialize(vk))
    assert M.vk_serialize(vk2) == M.vk_serialize(vk)
    assert M.verify(vk2, O.fr_mont([2]), proof)
    assert not M.verify(vk2, O.fr_mont([3]), proof)
    with pytest.raises(M.MarlinError):
        M.vk_deserialize(M.vk_serialize(vk)[:-3])
    blob = cs.to_bytes()
    cs2 = M.R1cs.from_bytes(blob)
    assert cs2.is_satisfied() and cs2.to_bytes() == blob
    pk2, _ = M.index(srs, cs2)
    assert M.prove(pk2, cs2, M.Rng()) == proof
    with pytest.raises(M.MarlinError):
        M.R1cs.from_bytes(blob[:-1])


def _python_circuit(case):
    from oracle import golden_marlin as PM
    a = case["args"]
    if case["kind"] == "manual":
        return PM.circuit_manual_constraints(a["v0"], a["v1"])
    if case["kind"] == "uint8_eq":
        return PM.circuit_uint8_equality(a["v0"], a["v1"])
    return PM.circuit_mul_chain(a["size"], a["v0"], a["v1"])


def _cpu_arm_circuit(case):
    a = case["args"]
    if max(a["v0"], a["v1"]) >= 1 << 64:             # values the built-ins cannot express: through the generic entry points
        cs = _python_circuit(case)
        return M.R1cs.custom(cs.num_instance, cs.num_witness, list(zip(cs.a, cs.b, cs.c)), cs.instance, cs.witness)
    return M.R1cs(case["kind"], **a)


def test_proof_and_vk_bytes_match_committed_fixtures(golden_dir):
    """tests/golden/marlin_proofs.json is made by the independent python restatement (oracle/golden_marlin.py,
    make_marlin_golden.py): the C++ protocol code on the CPU engine must produce exactly those bytes."""
    import hashlib
    g = json.load(open(os.path.join(golden_dir, "marlin_proofs.json")))["cases"]
    assert len(g) >= 6
    for tag, case in g.items():
        rng = M.Rng()
        srs = M.universal_setup(*case["bounds"], rng)
        cs = _cpu_arm_circuit(case)
        pk, vk = M.index(srs, cs)
        proof = M.prove(pk, cs, rng)
        assert len(proof) == case["proof_len"], tag
        assert hashlib.sha256(proof).hexdigest() == case["proof_sha256"], tag
        assert hashlib.sha256(M.vk_serialize(vk)).hexdigest() == case["vk_sha256"], tag
        if "proof_hex" in case:
            assert proof.hex() == case["proof_hex"] and M.vk_serialize(vk).hex() == case["vk_hex"], tag


def test_python_oracle_reproduces_the_fixtures(golden_dir):
    """The committed vectors are what oracle/golden_marlin.py computes today (the small cases; mul_chain_1000 is
    only run by the generator), its self-check passes, and its sumcheck identities hold -- prove() asserts that the
    outer and inner linear combinations vanish at beta / gamma, the debug_assert quoted at
    examples/schnorr-signature/main.rs:214-217."""
    from oracle import golden as G
    from oracle import golden_marlin as PM
    assert PM.self_check()
    g = json.load(open(os.path.join(golden_dir, "marlin_proofs.json")))["cases"]
    for tag, case in g.items():
        if "proof_hex" not in case:
            continue
        rng = G.test_rng()
        srs = PM.universal_setup(*case["bounds"], rng)
        cs = _python_circuit(case)
        pk, vk = PM.index(srs, cs)
        assert PM.prove(pk, cs, rng).hex() == case["proof_hex"], tag
        assert PM.vk_serialize(vk).hex() == case["vk_hex"], tag
    # an unsatisfied instance is refused
    bad = PM.circuit_manual_constraints(1, 2)
    pk, _ = PM.index(srs, bad)
    with pytest.raises(ValueError):
        PM.prove(pk, bad, G.test_rng())


def test_python_oracle_explicit_srs_points():
    """The python model commits through the trapdoor (p(beta) * g); here the same commitments are recomputed as
    real multi-scalar multiplications over explicit SRS points beta^i * g, as kzg10::commit does."""
    from oracle import golden as G
    from oracle import golden_marlin as PM
    rng = G.test_rng()
    srs = PM.universal_setup(100, 25, 300, rng)
    cs = PM.circuit_manual_constraints(1, 1)
    pk, vk = PM.index(srs, cs)
    ix = pk["index"]
    powers = [srs.power_of_g(i) for i in range(8)]
    for lp, comm in zip(ix.polys, vk["comms"]):
        acc = None
        for c, pt in zip(lp.coeffs, powers):
            acc = PM.E1.add(acc, PM.E1.mul(pt, c))
        assert acc == comm[0], lp.label
    # shifted powers: a degree-bounded commitment of X^0 against powers_of_g[D - bound ..] is the shift power itself
    ck = pk["ck"]
    bound = ck.bounds[-1]
    assert ck.commit_plain([1], shift=ck.shift_of(bound)) == dict(vk["shift_powers"])[bound]


def test_general_shape_circuit_prove_verify():
    """Rows with several terms, repeated columns and three public inputs (the shape gadget synthesis
    produces; the toy circuits above have one term per row)."""
    rng = M.Rng()
    srs = M.universal_setup(600, 600, 6000, rng)
    cs = M.R1cs("random_sparse", size=300, v0=5, v1=77)
    assert cs.is_satisfied()
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    pub = M.random_sparse_public_inputs(77)
    assert M.verify(vk, O.fr_mont(pub), proof)
    assert not M.verify(vk, O.fr_mont([pub[0], pub[1], pub[2] + 1]), proof)
    # the SWBR1CS1 round trip keeps it provable
    cs2 = M.R1cs.from_bytes(cs.to_bytes())
    assert M.prove(pk, cs2, M.Rng()) == M.prove(pk, cs, M.Rng())


def test_corrupted_proofs_and_keys_are_rejected_not_crashed(toy_srs):
    """Every sampled single-bit flip and every truncation of a valid proof is rejected (never accepted,
    never a crash); corrupted verifying-key bytes either fail to parse or stop verifying."""
    srs, rng = toy_srs
    cs = M.R1cs("manual", v0=1, v1=1)
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    pub = O.fr_mont([1])
    assert M.verify(vk, pub, proof)
    rs = np.random.RandomState(11)
    for pos in rs.choice(len(proof), 40, replace=False):
        bad = bytearray(proof)
        bad[int(pos)] ^= 1 << int(rs.randint(8))
        assert not M.verify(vk, pub, bytes(bad)), int(pos)
    for cut in (0, 1, 7, 8, 100, len(proof) - 48, len(proof) - 1):
        assert not M.verify(vk, pub, proof[:cut])
    assert not M.verify(vk, pub, proof + b"\x00")
    vkb = M.vk_serialize(vk)
    # every byte the verifier consumes; the last 16 (max_degree, supported_degree of the committer key) and
    # num_instance_variables (bytes 24..32) are carried along but not used by verification, as upstream
    for pos in list(rs.choice(len(vkb) - 16, 24, replace=False)) + [236, 250, 283]:      # 236..283: c_val = identity
        if 24 <= pos < 32:
            continue
        bad = bytearray(vkb)
        bad[int(pos)] ^= 0x04
        try:
            vk2 = M.vk_deserialize(bytes(bad))
        except M.MarlinError:
            continue
        assert not M.verify(vk2, pub, proof), int(pos)
    # non-canonical encodings are refused at parse time: x >= q, junk under the infinity flag
    g_at = 32 + 8 + 6 * 49
    bad = bytearray(vkb)
    bad[g_at:g_at + 48] = b"\xff" * 47 + bytes([0x3f | (vkb[g_at + 47] & 0x80)])
    with pytest.raises(M.MarlinError):
        M.vk_deserialize(bytes(bad))
    bad = bytearray(vkb)
    bad[236] ^= 1
    with pytest.raises(M.MarlinError):
        M.vk_deserialize(bytes(bad))


def test_r1cs_reader_rejects_malformed_input():
    """SWBR1CS1 (swb_r1cs_read / the CPU arm's reader): truncations, a wrong magic, out-of-range columns and
    non-canonical field elements are refused instead of read."""
    cs = M.R1cs("random_sparse", size=12, v0=3, v1=9)
    good = cs.to_bytes()
    assert M.R1cs.from_bytes(good).is_satisfied()
    for cut in list(range(0, 40)) + [len(good) // 2, len(good) - 1]:
        with pytest.raises(M.MarlinError):
            M.R1cs.from_bytes(good[:cut])
    with pytest.raises(M.MarlinError):
        M.R1cs.from_bytes(b"SWBR1CS2" + good[8:])
    with pytest.raises(M.MarlinError):
        M.R1cs.from_bytes(good + b"\x00")
    huge = bytearray(good)
    huge[24:32] = (1 << 31).to_bytes(8, "little")  # a constraint count the data cannot possibly hold
    with pytest.raises(M.MarlinError):
        M.R1cs.from_bytes(bytes(huge))
    bad = bytearray(good)
    first_entry = 8 + 24 + 8                       # magic, three counts, nnz of the first row
    bad[first_entry + 32:first_entry + 40] = (10 ** 6).to_bytes(8, "little")   # column out of range
    with pytest.raises(M.MarlinError):
        M.R1cs.from_bytes(bytes(bad))
    bad = bytearray(good)
    bad[first_entry:first_entry + 32] = b"\xff" * 32                           # coefficient >= r
    with pytest.raises(M.MarlinError):
        M.R1cs.from_bytes(bytes(bad))
