"""ctypes binding of oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY (see swb_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  Data layout helpers convert between python ints and the numpy arrays that both the oracle
and the product C-ABI use:
    Fr            (n, 4)  uint64   Montgomery limbs, little endian
    BigInteger256 (n, 4)  uint64   canonical integer
    Fq            (n, 6)  uint64
    G1 affine     (n, 13) uint64   x[6] y[6] infinity(byte0 of word 12)       = 104 B
    G1 jacobian   (n, 18) uint64   x[6] y[6] z[6]                              = 144 B
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

R_MOD = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
Q_MOD = int("01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800"
            "170b5d44300000008508c00000000001", 16)
FR_R = (1 << 256) % R_MOD
FQ_R = (1 << 384) % Q_MOD
FR_RINV = pow(FR_R, -1, R_MOD)
FQ_RINV = pow(FQ_R, -1, Q_MOD)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.uint64
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- int <-> limb arrays -------------------------------------------------------------------
def ints_to_limbs(vals, nlimbs: int) -> np.ndarray:
    out = np.zeros((len(vals), nlimbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(nlimbs):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr: np.ndarray) -> list[int]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, arr.shape[-1])
    return [sum(int(x) << (64 * j) for j, x in enumerate(row)) for row in arr]


def fr_mont(vals) -> np.ndarray:
    return ints_to_limbs([(v % R_MOD) * FR_R % R_MOD for v in vals], 4)


def fr_unmont(arr) -> list[int]:
    return [v * FR_RINV % R_MOD for v in limbs_to_ints(arr)]


def fq_mont(vals) -> np.ndarray:
    return ints_to_limbs([(v % Q_MOD) * FQ_R % Q_MOD for v in vals], 6)


def fq_unmont(arr) -> list[int]:
    return [v * FQ_RINV % Q_MOD for v in limbs_to_ints(arr)]


def affine_from_points(points) -> np.ndarray:
    """points: list of (x, y) python-int tuples or None (identity)."""
    out = np.zeros((len(points), 13), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None:
            out[i, 12] = 1
        else:
            out[i, 0:6] = fq_mont([p[0]])[0]
            out[i, 6:12] = fq_mont([p[1]])[0]
    return out


def points_from_affine(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 13)
    out = []
    for row in arr:
        if int(row[12]) & 0xFF:
            out.append(None)
        else:
            out.append((fq_unmont(row[None, 0:6])[0], fq_unmont(row[None, 6:12])[0]))
    return out


def points_from_jacobian(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 18)
    out = []
    for row in arr:
        x, y, z = (fq_unmont(row[None, 6 * k:6 * k + 6])[0] for k in range(3))
        if z == 0:
            out.append(None)
        else:
            zi = pow(z, -1, Q_MOD)
            out.append((x * zi * zi % Q_MOD, y * zi * zi * zi % Q_MOD))
    return out


# ---- oracle calls ------------------------------------------------------------------------
def fr_mul_vec(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    r = np.empty_like(a)
    lib().orc_fr_mul_vec(_p(r), _p(a), _p(b), ctypes.c_size_t(a.shape[0]))
    return r


def fq_mul_vec(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    r = np.empty_like(a)
    lib().orc_fq_mul_vec(_p(r), _p(a), _p(b), ctypes.c_size_t(a.shape[0]))
    return r


def fr_batch_inverse(a: np.ndarray) -> np.ndarray:
    r = np.ascontiguousarray(a.copy())
    lib().orc_fr_batch_inverse(_p(r), ctypes.c_size_t(r.shape[0]))
    return r


def g1_generator() -> np.ndarray:
    g = np.zeros((1, 13), dtype=np.uint64)
    lib().orc_g1_generator(_p(g))
    return g


def g1_mul(base_affine: np.ndarray, k: int) -> np.ndarray:
    out = np.zeros((1, 18), dtype=np.uint64)
    kk = ints_to_limbs([k], 4)
    lib().orc_g1_mul(_p(out), _p(np.ascontiguousarray(base_affine.reshape(1, 13))), _p(kk))
    return out


def g1_to_affine(jac: np.ndarray) -> np.ndarray:
    jac = np.ascontiguousarray(jac.reshape(-1, 18))
    out = np.zeros((jac.shape[0], 13), dtype=np.uint64)
    for i in range(jac.shape[0]):
        lib().orc_g1_to_affine(_p(out[i:i + 1]), _p(jac[i:i + 1]))
    return out


def g1_batch_normalize(jac: np.ndarray) -> np.ndarray:
    jac = np.ascontiguousarray(jac.reshape(-1, 18))
    out = np.zeros((jac.shape[0], 13), dtype=np.uint64)
    lib().orc_g1_batch_normalize(_p(out), _p(jac), ctypes.c_size_t(jac.shape[0]))
    return out


def msm_variable_base(bases: np.ndarray, scalars: np.ndarray, threads: int = 0) -> np.ndarray:
    """bases (n,13) affine Montgomery, scalars (n,4) canonical -> (1,18) Jacobian."""
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros((1, 18), dtype=np.uint64)
    lib().orc_msm_variable_base(_p(out), _p(bases), _p(scalars), ctypes.c_size_t(n), ctypes.c_int(threads))
    return out


def fixed_base_powers(g_jac: np.ndarray, beta_mont: np.ndarray, n: int, threads: int = 0) -> np.ndarray:
    out = np.zeros((n, 13), dtype=np.uint64)
    lib().orc_fixed_base_powers(_p(out), _p(np.ascontiguousarray(g_jac.reshape(1, 18))),
                                _p(np.ascontiguousarray(beta_mont.reshape(1, 4))), ctypes.c_size_t(n),
                                ctypes.c_int(threads))
    return out


def ntt(v: np.ndarray, log_n: int, inverse: bool = False, coset: bool = False, threads: int = 0) -> np.ndarray:
    r = np.ascontiguousarray(v.copy())
    assert r.shape == (1 << log_n, 4)
    lib().orc_ntt(_p(r), ctypes.c_uint32(log_n), ctypes.c_int(int(inverse)), ctypes.c_int(int(coset)),
                  ctypes.c_int(threads))
    return r


def num_threads() -> int:
    return lib().orc_num_threads()
