#!/usr/bin/env python3
"""Kernel-time breakdown of ONE steady-state proof (SRS with window tables) for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/prove_launches.csv python tools/prove_kernels.py 20
    python tools/prove_kernels.py --summarise gpurun_out/prove_launches.csv

Only the proof between cudaProfilerStart / Stop is captured."""
import csv
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    acc, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r is hdr or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "").replace("swb::", "")
        acc[name] += v
        cnt[name] += 1
    tot = sum(acc.values())
    print(f"total kernel time {tot:.2f} ms over {sum(cnt.values())} launches")
    for k, v in acc.most_common(40):
        print(f"{v:9.3f} ms {100 * v / tot:5.1f} %  x{cnt[k]:<5d} {k}")


def main():
    if sys.argv[1] == "--summarise":
        return summarise(sys.argv[2])
    import torch
    from simpleworks_b200.binding import Backend, ConstraintSystem, Marlin, Rng
    lg = int(sys.argv[1])
    be = Backend(0)
    m = Marlin(be)
    srs = m.generate_universal_srs(1 << lg, 1 << lg, 3 << lg, Rng())
    cs = ConstraintSystem.builtin("mul-chain", (1 << lg) - 2, 3, 5)
    pk, vk = m.generate_proving_and_verifying_keys(srs, cs)
    m.srs_set_tune_after(srs, 1)
    for _ in range(3):
        m.generate_proof(cs, pk, Rng())
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m.generate_proof(cs, pk, Rng())
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
