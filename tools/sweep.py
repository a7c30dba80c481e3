#!/usr/bin/env python3
"""BASELINE.json config 5 on one GPU: BLS12-377 G1 MSM and Fr NTT sweep 2^18..2^26 with roofline fractions.

    python tools/sweep.py [18,20,22,24,26]     -> gpurun_out/sweep.json

MSM: scalar distributions U (uniform) and M ("Marlin-like": 50 % zero, 25 % one, 25 % uniform -- SURVEY 8d), plain
path and window tables (swb_bases_precompute).  The fraction is k_msm_accumulate's limb-product rate over the IMAD.WIDE
issue rate measured in the same process; for M only the non-zero digits are counted as work.
NTT: forward transform in place; limb-product rate over the same peak and algorithmic bytes over the HBM peak.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleworks_b200 import _gen  # noqa: E402
from simpleworks_b200.binding import Backend  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand_fr(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randint(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    a[:, 3] &= 0x0FFFFFFFFFFFFFFF
    return a


def marlin_like(n, seed):
    s = rand_fr(n, seed)
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    kind = torch.randint(0, 4, (n,), device="cuda", generator=g)
    s[kind < 3] = 0
    s[kind == 2, 0] = 1
    return s


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    logs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "18,20,22,24,26").split(",")]
    be = Backend(0)
    peak = max(be.measure_imad_peak("wide", 20000), be.measure_imad_peak("wide_carry", 20000))
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    out = {"device": be.device_info(), "imad_wide_peak_per_s": peak, "hbm_peak_gbs": hbm, "msm": [], "ntt": []}
    for lg in logs:
        n = 1 << lg
        x = rand_fr(n, lg)
        ms = timed(lambda: be.ntt_(x, lg))
        row = {"log_n": lg, "ms": ms, "elems_per_s": n / ms * 1e3, "int_frac": (n / 2) * lg * 128 / ms * 1e3 / peak,
               "hbm_frac": 64.0 * n / ms / 1e6 / hbm}
        out["ntt"].append(row)
        print("ntt", row, flush=True)
    be.profile(True)
    for lg in logs:
        n = 1 << lg
        for tables in (False, True):
            bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(0x5357423230300001), n)
            if tables:
                bases.precompute(0)
                c, w = bases.table_info()
            else:
                c, w = be.msm_plan(n)
            for dist in ("U", "M"):
                s = rand_fr(n, 100 + lg) if dist == "U" else marlin_like(n, 200 + lg)
                ms = timed(lambda: be.msm(bases, s))
                acc_ms = be.last_stages().get("accumulate", float("nan"))
                # work actually done: for M only a quarter of the scalars have non-zero digits beyond the lowest
                adds = n * w if dist == "U" else n * (0.25 * w + 0.25)
                row = {"log_n": lg, "dist": dist, "tables": tables, "window_bits": c, "windows": w, "ms": ms,
                       "points_per_s": n / ms * 1e3, "accumulate_ms": acc_ms,
                       "accumulate_frac_of_imad_wide_peak": adds * 2880 / (acc_ms / 1e3) / peak}
                out["msm"].append(row)
                print("msm", row, flush=True)
            bases.free()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
