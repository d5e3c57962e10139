"""Site sharding for multi-GPU jobs.  The reference is single-device; given ϕ
the sites are independent, so rank r of R owns the contiguous site range
[r·n/R, (r+1)·n/R) of (xM, xP, y_o, y_unc, zMs), ϕ and zP are replicated, the terms that exist
once per evaluation (prior / log-Jacobian / entropy of the global block) are counted on rank 0
only (`count_globals`), and ONE all-reduce(sum) of [∇ϕ ; 6 components ; nL ; non-finite count]
(|ϕ| + 8 doubles) finishes the evaluation.  No other data-path collective.
"""
from __future__ import annotations

from typing import Tuple


def site_range(n_site: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced site range of a rank (first `n_site % world` ranks get one more)."""
    base, rem = divmod(n_site, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(data: dict, zMs, n_M: int, rank: int, world: int):
    """Slice a batch and its draws zMs (n_MC × n_M·n_site, site-major columns) for one rank."""
    n_site = data["xM"].shape[1]
    lo, hi = site_range(n_site, rank, world)
    d = {k: data[k][:, lo:hi] for k in ("xM", "xP", "y_o", "y_unc")}
    return d, zMs[:, lo * n_M:hi * n_M], lo


def nccl_unique_id() -> bytes:
    """128 bytes that identify a new NCCL communicator (hvi_nccl_unique_id): rank 0 creates them and hands them to every
    rank by any means (a torch.distributed / MPI broadcast, a file); each rank then calls `plan.comm_init(id, rank, world)`."""
    import ctypes as C
    from . import _abi
    lib = _abi.load_library()
    buf = C.create_string_buffer(128)
    rc = lib.hvi_nccl_unique_id(buf)
    if rc != 0:
        raise RuntimeError("hvi_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


def allreduce_result(out, dist=None):
    """Sum [∇ϕ ; components ; nL ; non-finite] over the ranks (torch tensor, in place)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out)
    return out


def adam_step_sharded(plan, out_dev, n_MC: int, seed: int, site_offset: int, dist=None, stream_ptr: int = 0):
    """One optimiser iteration of a site-sharded job without a host round trip: every rank evaluates its shard with
    the replicated device ϕ of its plan (`adam_init` first), the result vectors are summed over the ranks and every rank
    applies the same Adam update (`hvi_adam_apply_dev`), so ϕ stays replicated.  `out_dev`: float64 CUDA tensor of
    n_phi + 8 elements; afterwards out_dev[n_phi + 6] is the job's loss of this iteration."""
    plan.neg_elbo_async(plan.adam_phi_dev(), None, None, out_dev, n_MC, stream_ptr, seed=seed, site_offset=site_offset)
    allreduce_result(out_dev, dist)
    plan.adam_apply_dev(out_dev, stream_ptr)
    return out_dev

