#!/usr/bin/env python3
"""Multi-GPU proving check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_marlin_check.py [log_n]

Every rank runs the same prover with its commit / open MSMs sharded over the N GPUs (swb_set_msm_shard), the
partial commitments combined by the library's own NCCL communicator (swb_comm_init / swb_comm_sum_g1).
Checks: the proof of the `mul_chain_1000` fixture has the committed sha256 on every rank (i.e. the bytes of
a single GPU), a 2^log_n proof verifies, and prints the sharded proving time next to the single-GPU one.
Launched by tests/test_gpu_marlin.py::test_sharded_proving_two_gpus when two GPUs are visible."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpleworks_b200 import _gen  # noqa: E402
from simpleworks_b200.binding import Backend, ConstraintSystem, Marlin, Rng  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    be = Backend(local)
    m = Marlin(be)

    def run(bounds, size, v0, v1, proofs):
        rng = Rng()
        srs = m.generate_universal_srs(*bounds, rng)
        cs = ConstraintSystem.builtin("mul-chain", size, v0, v1)
        pk, vk = m.generate_proving_and_verifying_keys(srs, cs)
        ts, proof = [], None
        for _ in range(proofs):
            t0 = time.perf_counter(); proof = m.generate_proof(cs, pk, Rng()); ts.append(time.perf_counter() - t0)
        return proof, vk, min(ts)

    case = json.load(open(os.path.join(ROOT, "tests", "golden", "marlin_proofs.json")))["cases"]["mul_chain_1000"]
    # single-GPU timing first (no sharding), then sharded
    _, _, t_single = run((1 << log_n, 1 << log_n, 3 << log_n), (1 << log_n) - 2, 3, 5, 3)
    # the library's own communicator carries the partial commitments (one all-gather per prover round);
    # SWB_DIST_CALLBACK=1 uses the python callback over torch.distributed instead
    if os.environ.get("SWB_DIST_CALLBACK"):
        be.set_msm_shard(rank, world, dev)
    else:
        be.comm_init(rank, world)
        be.set_msm_shard(rank, world, use_comm=True)
        # the exchange primitive itself: sum of k * G over the ranks, two points per call
        from oracle import pyoracle as O
        mine = np.concatenate([O.g1_mul(O.g1_generator(), rank + 1), O.g1_mul(O.g1_generator(), 10 * (rank + 1))])
        got = be.comm_sum_g1(mine)
        tot = world * (world + 1) // 2
        want = np.concatenate([O.g1_mul(O.g1_generator(), tot), O.g1_mul(O.g1_generator(), 10 * tot)])
        assert np.array_equal(O.g1_to_affine(got), O.g1_to_affine(want)), "swb_comm_sum_g1 gave a wrong sum"
    rng = Rng()
    srs = m.generate_universal_srs(*case["bounds"], rng)
    cs = ConstraintSystem.builtin("mul-chain", 1000, 7, 11)
    pk, vk = m.generate_proving_and_verifying_keys(srs, cs)
    proof = m.generate_proof(cs, pk, rng)
    assert hashlib.sha256(proof).hexdigest() == case["proof_sha256"], f"rank {rank}: sharded proof differs from the fixture"
    assert hashlib.sha256(m.serialize_verifying_key(vk)).hexdigest() == case["vk_sha256"]
    big, vk_big, t_sharded = run((1 << log_n, 1 << log_n, 3 << log_n), (1 << log_n) - 2, 3, 5, 3)
    assert m.verify_proof(vk_big, _gen.fr_mont(3), big)
    h = torch.tensor(list(hashlib.sha256(big).digest()), dtype=torch.int64, device=dev)
    allh = torch.empty((world, 32), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allh, h.reshape(1, 32))
    assert bool((allh == allh[0]).all()), "ranks produced different proofs"
    be.set_msm_shard(0, 1)
    if rank == 0:
        print(json.dumps({"ok": True, "world": world, "log_constraints": log_n, "prove_s_single_gpu": t_single,
                          "prove_s_msm_sharded": t_sharded}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
