#!/usr/bin/env python3
"""A prover round's commit list (three 2^20 and one 3*2^20 vector) as one batched MSM vs single calls, per table width"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from simpleworks_b200 import _gen
from simpleworks_b200.binding import Backend
be = Backend(0)
be.profile(True)
n = 1 << 22
sizes = [1 << 20, 1 << 20, 1 << 20, 3 << 20]
vecs = [torch.from_numpy(bench.synth_scalars_host(s, 7 + i).view(np.int64)).to("cuda:0") for i, s in enumerate(sizes)]
for c in (0, 17, 18, 19, 20, 21):
    bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(bench.BETA_SEED), n)
    if c:
        bases.precompute(c)
    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3, out
    tb, ob = timed(lambda: be.msm_batch(bases, vecs, montgomery=True))
    st = be.last_stages()
    ts, os_ = timed(lambda: np.concatenate([be.msm(bases, v, montgomery=True) for v in vecs]))
    print("c", c, bases.table_info(), "batch ms", round(tb, 2), "singles ms", round(ts, 2), "equal", bool(np.array_equal(ob, os_)),
          json.dumps({k: round(v, 2) for k, v in st.items()}), flush=True)
    bases.free()
