"""ctypes binding of oracle/liboracle_marlin.so (CPU arm of the Marlin protocol layer).
TEST INFRASTRUCTURE ONLY -- see marlin_oracle.cpp."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        L = ctypes.CDLL(os.path.join(_HERE, "liboracle_marlin.so"))
        vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.orc_marlin_last_error.restype = ctypes.c_char_p
        for name in ("orc_rng_test_rng", "orc_r1cs_new", "orc_r1cs_builtin", "orc_rng_from_seed"):
            getattr(L, name).restype = vp
        L.orc_rng_from_seed.argtypes = [ctypes.c_char_p, i32]
        L.orc_rng_next_u64.restype = ctypes.c_uint64
        L.orc_rng_next_u64.argtypes = [vp]
        L.orc_rng_next_u32.restype = ctypes.c_uint32
        L.orc_rng_next_u32.argtypes = [vp]
        L.orc_rng_fr_rand.argtypes = [vp, vp]
        L.orc_rng_free.argtypes = [vp]
        L.orc_blake2s.argtypes = [ctypes.c_char_p, sz, ctypes.c_char_p]
        L.orc_r1cs_new.argtypes = [sz, sz]
        L.orc_r1cs_builtin.argtypes = [i32, sz, ctypes.c_uint64, ctypes.c_uint64]
        L.orc_r1cs_add_constraint.argtypes = [vp, vp, vp, sz, vp, vp, sz, vp, vp, sz]
        L.orc_r1cs_set_assignment.argtypes = [vp, vp, sz, vp, sz]
        L.orc_r1cs_is_satisfied.argtypes = [vp]
        L.orc_r1cs_free.argtypes = [vp]
        L.orc_marlin_universal_setup.argtypes = [sz, sz, sz, vp, ctypes.POINTER(vp)]
        L.orc_srs_max_degree.restype = sz
        L.orc_srs_max_degree.argtypes = [vp]
        L.orc_srs_free.argtypes = [vp]
        L.orc_marlin_index.argtypes = [vp, vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
        L.orc_pk_free.argtypes = [vp]
        L.orc_vk_free.argtypes = [vp]
        L.orc_marlin_prove.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)), ctypes.POINTER(sz)]
        L.orc_marlin_verify.argtypes = [vp, vp, sz, ctypes.c_char_p, sz, vp, ctypes.POINTER(i32)]
        L.orc_bytes_free.argtypes = [ctypes.POINTER(ctypes.c_uint8)]
        L.orc_r1cs_read.restype = vp
        L.orc_r1cs_read.argtypes = [ctypes.c_char_p, sz]
        L.orc_r1cs_write.restype = ctypes.POINTER(ctypes.c_uint8)
        L.orc_r1cs_write.argtypes = [vp, ctypes.POINTER(sz)]
        L.orc_vk_serialize.restype = ctypes.POINTER(ctypes.c_uint8)
        L.orc_vk_serialize.argtypes = [vp, ctypes.POINTER(sz)]
        L.orc_vk_deserialize.restype = vp
        L.orc_vk_deserialize.argtypes = [ctypes.c_char_p, sz]
        _LIB = L
    return _LIB


class MarlinError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise MarlinError(lib().orc_marlin_last_error().decode())


class Rng:
    def __init__(self, seed: bytes | None = None, rounds: int = 20):
        self.h = ctypes.c_void_p(lib().orc_rng_test_rng() if seed is None else lib().orc_rng_from_seed(seed, rounds))

    def next_u64(self):
        return lib().orc_rng_next_u64(self.h)

    def next_u32(self):
        return lib().orc_rng_next_u32(self.h)

    def fr_rand_mont(self) -> int:
        out = np.zeros(4, dtype=np.uint64)
        lib().orc_rng_fr_rand(self.h, out.ctypes.data_as(ctypes.c_void_p))
        return sum(int(v) << (64 * i) for i, v in enumerate(out))


def blake2s(data: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().orc_blake2s(data, len(data), out)
    return out.raw


def random_sparse_public_inputs(seed: int) -> list:
    """the three public inputs of R1cs('random_sparse', v1=seed): first draws of its xorshift generator"""
    m = (1 << 64) - 1
    st = (seed * 0x9E3779B97F4A7C15 + 0x1234567) & m
    out = []
    for _ in range(3):
        st ^= (st << 13) & m
        st ^= st >> 7
        st ^= (st << 17) & m
        out.append(st >> 8)
    return out


class R1cs:
    """kind: 'manual' (examples/manual-constraints.rs), 'uint8_eq' (examples/test-circuit.rs),
    'chain' (synthetic x_i * x_{i+1} = x_{i+2}), 'random_sparse' (size constraints, v0 terms per row, seed v1;
    3 public inputs, see random_sparse_public_inputs)."""

    def __init__(self, kind: str, size: int = 0, v0: int = 1, v1: int = 1):
        k = {"manual": 0, "uint8_eq": 1, "chain": 2, "random_sparse": 3}[kind]
        self.h = ctypes.c_void_p(lib().orc_r1cs_builtin(k, size, v0, v1))

    @classmethod
    def custom(cls, num_instance: int, num_witness: int, constraints, instance, witness) -> "R1cs":
        """a constraint system given as python integers: constraints = [(a_row, b_row, c_row)] with rows
        [(coefficient, column)], assignment values mod r (instance[0] must be 1)"""
        from oracle import pyoracle as O
        L = lib()
        obj = cls.__new__(cls)
        obj.h = ctypes.c_void_p(L.orc_r1cs_new(num_instance, num_witness))

        def pack(row):
            coef = O.fr_mont([k for k, _ in row]) if row else np.zeros((0, 4), np.uint64)
            col = np.ascontiguousarray([j for _, j in row], dtype=np.uint32)
            return coef, col
        for ra, rb, rc in constraints:
            (ac, ai), (bc, bi), (cc, ci) = pack(ra), pack(rb), pack(rc)
            if L.orc_r1cs_add_constraint(obj.h, ac.ctypes.data, ai.ctypes.data, len(ra), bc.ctypes.data, bi.ctypes.data, len(rb),
                                         cc.ctypes.data, ci.ctypes.data, len(rc)):
                raise MarlinError("add_constraint: column out of range")
        inst, wit = O.fr_mont(list(instance)), O.fr_mont(list(witness))
        if L.orc_r1cs_set_assignment(obj.h, inst.ctypes.data, len(instance), wit.ctypes.data, len(witness)):
            raise MarlinError("set_assignment: wrong sizes")
        return obj

    def is_satisfied(self) -> bool:
        return bool(lib().orc_r1cs_is_satisfied(self.h))

    def to_bytes(self) -> bytes:
        n = ctypes.c_size_t()
        p = lib().orc_r1cs_write(self.h, ctypes.byref(n))
        out = bytes(p[:n.value])
        lib().orc_bytes_free(p)
        return out

    @classmethod
    def from_bytes(cls, data: bytes) -> "R1cs":
        h = lib().orc_r1cs_read(data, len(data))
        if not h:
            raise MarlinError("malformed SWBR1CS1 data")
        obj = cls.__new__(cls)
        obj.h = ctypes.c_void_p(h)
        return obj


def vk_serialize(vk) -> bytes:
    n = ctypes.c_size_t()
    p = lib().orc_vk_serialize(vk, ctypes.byref(n))
    out = bytes(p[:n.value])
    lib().orc_bytes_free(p)
    return out


def vk_deserialize(data: bytes):
    h = lib().orc_vk_deserialize(data, len(data))
    if not h:
        raise MarlinError("malformed verifying key")
    return ctypes.c_void_p(h)


def universal_setup(nc, nv, nnz, rng: Rng):
    srs = ctypes.c_void_p()
    _chk(lib().orc_marlin_universal_setup(nc, nv, nnz, rng.h, ctypes.byref(srs)))
    return srs


def index(srs, cs: R1cs):
    pk, vk = ctypes.c_void_p(), ctypes.c_void_p()
    _chk(lib().orc_marlin_index(srs, cs.h, ctypes.byref(pk), ctypes.byref(vk)))
    return pk, vk


_shard_cb = None


def set_msm_shard(rank: int, world: int, combine=None):
    """CPU-arm mirror of swb_set_msm_shard: every prover MSM only covers this process's share of the index
    range; combine(partial (1,18) uint64 Jacobian) must return the sum over all processes.  world <= 1 = off."""
    global _shard_cb
    L = lib()
    L.orc_set_msm_shard.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    if world <= 1 or combine is None:
        L.orc_set_msm_shard(0, 1, None, None)
        _shard_cb = None
        return
    cb_t = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)

    def tramp(_user, mine, out):
        try:
            part = np.ctypeslib.as_array(ctypes.cast(mine, ctypes.POINTER(ctypes.c_uint64)), shape=(1, 18)).copy()
            total = np.ascontiguousarray(combine(part), dtype=np.uint64)
            ctypes.memmove(out, total.ctypes.data, 144)
            return 0
        except Exception:
            return 1
    _shard_cb = cb_t(tramp)
    L.orc_set_msm_shard(rank, world, ctypes.cast(_shard_cb, ctypes.c_void_p), None)


def prove(pk, cs: R1cs, rng: Rng) -> bytes:
    p = ctypes.POINTER(ctypes.c_uint8)()
    n = ctypes.c_size_t()
    _chk(lib().orc_marlin_prove(pk, cs.h, rng.h, ctypes.byref(p), ctypes.byref(n)))
    out = bytes(p[:n.value])
    lib().orc_bytes_free(p)
    return out


def verify(vk, public_inputs_mont: np.ndarray, proof: bytes, rng: Rng | None = None) -> bool:
    ok = ctypes.c_int()
    pi = np.ascontiguousarray(public_inputs_mont, dtype=np.uint64).reshape(-1, 4)
    _chk(lib().orc_marlin_verify(vk, pi.ctypes.data_as(ctypes.c_void_p), pi.shape[0], proof, len(proof),
                                 rng.h if rng else None, ctypes.byref(ok)))
    return bool(ok.value)
