import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from simpleworks_b200 import _gen
from simpleworks_b200.binding import Backend
be = Backend(0)
n = 1 << 24
bases = be.bases_from_powers(_gen.g1_generator_jacobian(), _gen.fr_mont(bench.BETA_SEED), n)
bases.precompute(0)
dev = torch.from_numpy(bench.synth_scalars_host(n, 1234).view(np.int64)).to("cuda:0")
be.set_msm_pair_sums(3)
for _ in range(2):
    be.msm(bases, dev)
torch.cuda.synchronize()
