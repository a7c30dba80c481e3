"""CPU, world_size 2 over gloo: the multi-GPU MSM path's host logic -- index sharding, the
all-gather of 144-byte partial results and their summation inside libswb200 -- with each rank's
partial MSM computed by the CPU oracle (no GPU here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as O
    from simpleworks_b200 import binding
    g = O.g1_mul(O.g1_generator(), 1)
    bases = O.fixed_base_powers(g, O.fr_mont([0x5357423230300001]), n)
    rs = np.random.RandomState(5)
    scalars = rs.randint(0, 2 ** 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
    scalars[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    lo, hi = binding.shard_range(n, rank, world)
    partial = O.msm_variable_base(np.ascontiguousarray(bases[lo:hi]), np.ascontiguousarray(scalars[lo:hi]), threads=1)
    total = binding.combine_partials(partial, world)
    full = O.g1_to_affine(O.msm_variable_base(bases, scalars, threads=1))
    q.put((rank, bool(np.array_equal(O.g1_to_affine(total), full)), int(lo), int(hi)))
    dist.destroy_process_group()


def test_sharded_msm_combine_world2():
    from simpleworks_b200 import build
    build.build()
    world, n = 2, 1001          # odd n: uneven shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == n


def _prove_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hashlib
    from oracle import pymarlin as M
    from simpleworks_b200 import binding
    M.set_msm_shard(rank, world, lambda part: binding.combine_partials(part, world))
    rng = M.Rng()
    srs = M.universal_setup(100, 25, 300, rng)
    cs = M.R1cs("chain", size=20, v0=3, v1=5)
    pk, vk = M.index(srs, cs)
    proof = M.prove(pk, cs, rng)
    M.set_msm_shard(0, 1)
    q.put((rank, hashlib.sha256(proof).hexdigest(), hashlib.sha256(M.vk_serialize(vk)).hexdigest()))
    dist.destroy_process_group()


def test_sharded_proving_world2_gives_the_single_process_proof():
    """Multi-process proving (the CPU-arm mirror of swb_set_msm_shard): two gloo ranks run the same prover,
    every commit / open MSM is split by index range and the partial results are all-gathered and summed
    by libswb200 -- both ranks must produce the committed single-process proof and verifying key."""
    import json
    from simpleworks_b200 import build
    build.build()
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "marlin_proofs.json")))["cases"]["mul_chain_20"]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_prove_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, ph, vh in res:
        assert ph == case["proof_sha256"] and vh == case["vk_sha256"]


def test_shard_range_covers_everything():
    from simpleworks_b200.binding import shard_range
    for n in (0, 1, 7, 8, 1 << 20):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_g1_sum_host_only():
    """swb_g1_sum_jacobian is host arithmetic: identity handling, doubling, cancellation."""
    from oracle import pyoracle as O
    from simpleworks_b200 import binding
    g = O.g1_generator()
    p3, p5 = O.g1_mul(g, 3), O.g1_mul(g, 5)
    zero = np.zeros((1, 18), dtype=np.uint64)
    zero[0, 6:12] = O.fq_mont([1])[0]
    got = binding.g1_sum(np.concatenate([p3, zero, p5]))
    assert np.array_equal(O.g1_to_affine(got), O.g1_to_affine(O.g1_mul(g, 8)))
    assert np.array_equal(O.g1_to_affine(binding.g1_sum(np.concatenate([p3, p3]))), O.g1_to_affine(O.g1_mul(g, 6)))
    neg = O.g1_mul(g, O.R_MOD - 3)
    assert O.points_from_jacobian(binding.g1_sum(np.concatenate([p3, neg])))[0] is None
    assert O.points_from_jacobian(binding.g1_sum(np.zeros((0, 18), dtype=np.uint64)))[0] is None
