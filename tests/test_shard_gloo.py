"""The N>1 path on CPU: two processes (gloo), each evaluating its site shard with the C
restatement standing in for the GPU (same sharding contract: contiguous site ranges, global
terms on rank 0 only), one all-reduce of [∇ϕ ; components] -- must equal the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, c_oracle_value_and_grad, hvi_b200


def _worker(rank, world, port, nb, M, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hvi_b200.shard import allreduce_result, shard_batch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    d, z, lo = shard_batch(data, zMs, prob.n_M, rank, world)
    nL, comps, g = c_oracle_value_and_grad(prob, ϕ, d, zP, z, count_globals=(rank == 0))
    out = torch.from_numpy(np.concatenate([g, comps, [nL]]))
    allreduce_result(out, dist)
    if rank == 0:
        ret.put(out.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_site_range_partition():
    from hvi_b200.shard import site_range
    for n in (1, 7, 20, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            r = [site_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_allreduce_equals_unsharded():
    nb, M, world = 37, 4, 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, world, port, nb, M, ret)) for r in range(world)]
    for p in procs:
        p.start()
    out = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    nL, comps, g = c_oracle_value_and_grad(prob, prob.init_ϕ(perturbed=True), data, zP, zMs)
    n = prob.n_phi
    np.testing.assert_allclose(out[:n], g, rtol=1e-10, atol=1e-10 * np.abs(g).max())
    np.testing.assert_allclose(out[n:n + 6], comps, rtol=1e-12)
    assert out[n + 6] == pytest.approx(nL, rel=1e-12)
