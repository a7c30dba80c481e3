#!/usr/bin/env python3
"""One GPU playing rank r of `world` under bucket-range sharding: per-stage times of the 2^log_n MSM share
(what each GPU of an N-GPU run would spend), checked by adding up all shares once at a smaller size."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from simpleworks_b200 import _gen
    from simpleworks_b200.binding import Backend
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=26)
    ap.add_argument("--pairs", type=int, default=1, help="swb_msm_set_pair_sums policy")
    args = ap.parse_args()
    be = Backend(0)
    n = 1 << args.log_n
    bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(bench.BETA_SEED), n)
    bases.precompute(0)
    print("tables", bases.table_info(), flush=True)
    dev = torch.from_numpy(bench.synth_scalars_host(n, 1234).view(np.int64)).to("cuda:0")
    be.profile(True)
    be.set_msm_pair_sums(args.pairs)
    out = {}
    for world in (1, 2, 4, 8):
        rows = []
        for rank in sorted({0, world - 1}):
            be.set_msm_bucket_shard(rank, world)
            for _ in range(2):
                be.msm(bases, dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            be.msm(bases, dev)
            e1.record()
            torch.cuda.synchronize()
            rows.append({"rank": rank, "ms": e0.elapsed_time(e1), "stages": be.last_stages()})
        out[world] = rows
        print(world, json.dumps(rows), flush=True)
    be.set_msm_bucket_shard(0, 1)
    base = out[1][0]["ms"]
    for world in (2, 4, 8):
        worst = max(r["ms"] for r in out[world])
        print(f"world {world}: {worst:.1f} ms per rank -> efficiency {base / (world * worst):.3f}")


if __name__ == "__main__":
    main()
