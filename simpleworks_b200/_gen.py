"""Constant helpers for host-side callers (generator point, Montgomery encoding of small ints)."""
from __future__ import annotations

import numpy as np

R_MOD = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
Q_MOD = int("01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800"
            "170b5d44300000008508c00000000001", 16)
G1_X = 81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695
G1_Y = 241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030


def _limbs(v: int, n: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def fr_mont(v: int) -> np.ndarray:
    """(1,4) Montgomery representation of an integer mod r."""
    return _limbs((v % R_MOD) * (1 << 256) % R_MOD, 4).reshape(1, 4)


def fq_mont(v: int) -> np.ndarray:
    return _limbs((v % Q_MOD) * (1 << 384) % Q_MOD, 6).reshape(1, 6)


def g1_generator_jacobian() -> np.ndarray:
    """(1,18): the BLS12-377 G1 generator (ark-bls12-377 curves/g1.rs) as (x, y, 1)."""
    return np.concatenate([fq_mont(G1_X), fq_mont(G1_Y), fq_mont(1)], axis=1)
