#!/usr/bin/env python3
"""Small fixed workload for ncu captures: python tools/prof_target.py [msm|ntt] [log_n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleworks_b200 import _gen  # noqa: E402
from simpleworks_b200.binding import Backend  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "msm"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
be = Backend(0)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randint(-2 ** 63, 2 ** 63 - 1, (1 << log_n, 4), dtype=torch.int64, device="cuda", generator=g)
x[:, 3] &= 0x0FFFFFFFFFFFFFFF
if what == "msm":
    bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(0x5357423230300001), 1 << log_n)
    if os.environ.get("MSM_TABLES"):
        bases.precompute(0)
    for _ in range(3):
        be.msm(bases, x)
else:
    for _ in range(3):
        be.ntt_(x, log_n)
torch.cuda.synchronize()
print("done", what, log_n)
