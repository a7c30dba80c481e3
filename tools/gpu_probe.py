#!/usr/bin/env python3
"""Quick GPU probe: limb-product peak, NTT and MSM timings at a few sizes (CUDA events)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleworks_b200.binding import Backend  # noqa: E402


def rand_fr(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randint(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    a[:, 3] &= 0x0FFFFFFFFFFFFFFF
    return a


def timeit(fn, reps=3):
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
    out = {}
    be = Backend(0)
    be.use_torch_stream()
    out["device"] = be.device_info()
    for f in ("fr", "fq"):
        out["peak_" + f] = be.measure_mul_peak(f, 4000)
    for k in ("lo", "wide", "wide_carry", "addc"):
        out["imad_" + k] = be.measure_imad_peak(k, 20000)
    print(json.dumps(out), flush=True)
    sizes = [int(s) for s in os.environ.get("NTT_LOGS", "16,20,22,24").split(",")]
    for log_n in sizes:
        x = rand_fr(1 << log_n, log_n)
        ms = timeit(lambda: be.ntt_(x, log_n))
        out[f"ntt_2^{log_n}_ms"] = ms
        print(f"ntt 2^{log_n}: {ms:.3f} ms  {(1 << log_n) / ms / 1e6:.2f} Gelem/s", flush=True)
    msizes = [int(s) for s in os.environ.get("MSM_LOGS", "16,18,20,22").split(",")]
    nmax = 1 << max(msizes)
    g = np.zeros((1, 18), dtype=np.uint64)
    # generator in Montgomery form comes from the fixed-base routine: use beta^i * G with G = (x,y,1)
    from simpleworks_b200 import _gen
    t0 = time.time()
    pts = be.fixed_base_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(0x5357423230300001), nmax)
    out["fixed_base_s"] = time.time() - t0
    print(f"fixed-base 2^{max(msizes)} points: {out['fixed_base_s']:.2f} s", flush=True)
    bases = be.load_bases(pts)
    for log_n in msizes:
        s = rand_fr(1 << log_n, 100 + log_n)
        ms = timeit(lambda: be.msm(bases, s))
        out[f"msm_2^{log_n}_ms"] = ms
        print(f"msm 2^{log_n}: {ms:.3f} ms  {(1 << log_n) / ms / 1e3:.2f} Mpts/s", flush=True)
    bases.free()
    if os.environ.get("MSM_TABLES"):
        # window tables sized for each MSM (swb_bases_precompute); MSM_TABLES=auto or a list of digit widths
        spec = os.environ["MSM_TABLES"]
        widths = [0] if spec == "auto" else [int(x) for x in spec.split(",")]
        for log_n in msizes:
            s = rand_fr(1 << log_n, 100 + log_n)
            for cw in widths:
                tb = be.load_bases(pts[: 1 << log_n])
                t0 = time.time()
                try:
                    tb.precompute(cw)
                except Exception as e:  # e.g. too many levels for a narrow width
                    print(f"tables c={cw} 2^{log_n}: {e}", flush=True)
                    tb.free()
                    continue
                dt = time.time() - t0
                ms = timeit(lambda: be.msm(tb, s))
                out[f"msm_tables_c{cw}_2^{log_n}_ms"] = ms
                print(f"msm+tables{tb.table_info()} 2^{log_n}: {ms:.3f} ms  {(1 << log_n) / ms / 1e3:.2f} Mpts/s  (precompute {dt:.2f} s)",
                      flush=True)
                tb.free()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
