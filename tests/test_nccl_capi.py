"""The collective of site-sharded jobs behind the C ABI (hvi_nccl_unique_id, hvi_plan_comm_init, hvi_neg_elbo_grad_sharded,
hvi_allreduce_result): the reference is single-device (src/HybridSolver.jl:184-267); sharding sums
[∇ϕ; components; nL; non-finite] over the ranks with ONE ncclAllReduce.

One GPU: a communicator of one rank (the all-reduce is the identity) against the unsharded evaluation.
Two or more GPUs: tests/nccl_two_ranks.py (one process per GPU) is launched and must report agreement with the unsharded
result on every rank; skipped on a single-GPU box.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import hvi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f32", 1e-6), ("f64", 1e-13)])
def test_communicator_of_one_rank_is_the_unsharded_evaluation(dtype, tol):
    from hvi_b200 import shard
    prob = hvi_b200.DoubleMMCase(dtype=dtype)
    nb, M = 700, 5
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    ref = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    with pytest.raises(RuntimeError):          # no communicator yet
        plan.neg_elbo_withgradient_sharded(ϕ, zP, zMs)
    plan.comm_init(shard.nccl_unique_id(), 0, 1)
    got = plan.neg_elbo_withgradient_sharded(ϕ, zP, zMs)
    assert abs(got[0] - ref[0]) <= tol * abs(ref[0])
    np.testing.assert_allclose(np.array(got[1]), np.array(ref[1]), rtol=tol, atol=0)
    np.testing.assert_allclose(got[2], ref[2], rtol=tol, atol=tol * np.max(np.abs(ref[2])))
    # device draws through the same entry point
    a = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=3)
    b = plan.neg_elbo_withgradient_sharded(ϕ, n_MC=M, seed=3)
    assert abs(a[0] - b[0]) <= tol * abs(a[0])
    plan.comm_destroy()
    plan.comm_destroy()                        # idempotent


@pytest.mark.gpu
def test_two_ranks_sum_to_the_unsharded_result():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "nccl_two_ranks.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("nccl_two_ranks ok") == 2, out.stdout[-2000:]
