#!/usr/bin/env python3
"""2^log_n MSM with and without the batch-affine pair-sum pass: per-stage times, results compared."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from simpleworks_b200 import _gen
    from simpleworks_b200.binding import Backend
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", default="26")
    ap.add_argument("--no-tables", action="store_true")
    args = ap.parse_args()
    be = Backend(0)
    be.profile(True)
    for log_n in [int(x) for x in args.log_n.split(",")]:
        n = 1 << log_n
        bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(bench.BETA_SEED), n)
        if not args.no_tables:
            bases.precompute(0)
        dev = torch.from_numpy(bench.synth_scalars_host(n, 1234).view(np.int64)).to("cuda:0")
        res = {}
        for pol in (0, 2, 3, 4, 5):
            be.set_msm_pair_sums(pol)
            for _ in range(2):
                out = be.msm(bases, dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = be.msm(bases, dev); e1.record(); torch.cuda.synchronize()
            res[pol] = (e0.elapsed_time(e1), be.last_stages(), out)
            print(log_n, "tables" if not args.no_tables else "plain", bases.table_info(), f"pair levels {pol - 1}" if pol else "xyzz only",
                  round(res[pol][0], 2), json.dumps({k: round(v, 2) for k, v in res[pol][1].items()}), flush=True)
        print("  equal:", all(bool(np.array_equal(res[0][2], res[k][2])) for k in res), " speedup", {k - 1: round(res[0][0] / res[k][0], 3) for k in res if k}, flush=True)
        be.set_msm_pair_sums(1)
        bases.free()
        del dev


if __name__ == "__main__":
    main()
