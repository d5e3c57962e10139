"""One process per GPU (torchrun): the C-ABI collective of a site-sharded job.  Every rank evaluates its contiguous site range
and hvi_neg_elbo_grad_sharded all-reduces the result with NCCL inside the library; every rank must hold the unsharded result
(computed on rank-local GPU by a second, unsharded plan).  torch.distributed (gloo) only carries the 128-byte NCCL id."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hvi_b200  # noqa: E402
from hvi_b200 import shard  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for dtype, tol in (("f32", 2e-5), ("f64", 1e-12)):
        for scen in ((), ("covarK2",)):
            prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
            nb, M = 5000, 6
            data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
            zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
            ϕ = prob.init_ϕ(perturbed=True)
            whole = hvi_b200.NegElboPlan(prob, device=local)
            whole.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
            ref = whole.neg_elbo_withgradient(ϕ, zP, zMs)
            ref_rng = whole.neg_elbo_withgradient(ϕ, n_MC=M, seed=11)
            d, z, lo = shard.shard_batch(data, zMs, prob.n_M, rank, world)
            plan = hvi_b200.NegElboPlan(prob, device=local, count_globals=(rank == 0))
            plan.set_data(d["xM"], d["xP"], d["y_o"], d["y_unc"])
            uid = [shard.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            plan.comm_init(uid[0], rank, world)
            got = plan.neg_elbo_withgradient_sharded(ϕ, zP, z, site_offset=lo)
            got_rng = plan.neg_elbo_withgradient_sharded(ϕ, n_MC=M, seed=11, site_offset=lo)
            for g, r in ((got, ref), (got_rng, ref_rng)):
                assert abs(g[0] - r[0]) <= tol * abs(r[0]), (dtype, scen, g[0], r[0])
                assert np.max(np.abs(np.array(g[1]) - np.array(r[1]))) <= tol * np.max(np.abs(np.array(r[1])))
                for blk, rr in prob.positions().items():
                    if len(rr):
                        err = np.max(np.abs(g[2][rr].astype(np.float64) - r[2][rr].astype(np.float64)))
                        assert err <= 5 * tol * np.max(np.abs(r[2][rr])), (dtype, scen, blk, err)
            plan.comm_destroy()
    dist.barrier()
    print(f"nccl_two_ranks ok rank {rank}/{world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
