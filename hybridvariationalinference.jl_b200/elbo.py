"""Host-side mirror of the reference's cost-function interface for the neg-ELBO path, on top of
the C-ABI library (include/hvi_b200.h).

Reference interface mirrored (paths relative to /root/reference):
  * ``get_loss_elbo`` closure ``loss_elbo(ϕ, rng, xM, xP, y_o, y_unc, i_sites)`` (src/HybridSolver.jl:293-324)
  * ``neg_elbo_gtf`` / ``neg_elbo_gtf_components`` (src/elbo.jl:30-88)
  * ``generate_ζ`` (src/elbo.jl:472-521), ``transformU_block_cholesky1`` positions
    (src/cholesky.jl:30-50,233-247), ``ComponentArrayInterpreter`` positions
    (src/ComponentArrayInterpreter.jl:281-288)

The draws ξ = (zP, zMs) of ``sample_ζresid_norm`` (src/elbo.jl:605-606) are either passed in (bit
reproducible, used for parity) or produced on the device from a seed (like the reference's
CUDA.randn override, ext/HybridVariationalInferenceCUDAExt.jl:82-86).

No CPU fallback: every call goes through the CUDA library and raises if it is not available.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import numpy as np

from . import _abi as A
from .problem import HybridProblem

COMPONENT_NAMES = ("nLjoint", "entropy_ζ", "loss_penalty", "nLy", "neg_log_prior", "neg_log_jac")


class HVIError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hvi_b200 error {code}: {msg}")
        self.code = code


class NegElboComponents(NamedTuple):
    """NamedTuple returned by neg_elbo_gtf_components (src/elbo.jl:200)."""
    nLjoint: float
    entropy_ζ: float
    loss_penalty: float
    nLy: float
    neg_log_prior: float
    neg_log_jac: float


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor (device memory handle only)
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def _is_device(a):
    return hasattr(a, "data_ptr") and a.is_cuda


def _check_colmajor(tensors):
    """Device tensors are read as the reference's column-major (rows x n_site) arrays: a 2-D tensor must have
    stride (1, rows) -- e.g. `torch.empty(n_site, rows).T` -- a C-contiguous (rows, n_site) tensor is refused."""
    for t in tensors:
        if t is None:
            continue
        if t.dim() == 2 and t.shape[1] > 1 and t.shape[0] > 1 and not (t.stride(0) == 1 and t.stride(1) == t.shape[0]):
            raise ValueError(f"device array of shape {tuple(t.shape)} has strides {t.stride()}: expected column-major "
                             f"(1, {t.shape[0]})")
        if t.dim() == 1 and not t.is_contiguous():
            raise ValueError("1-D device array must be contiguous")


class NegElboPlan:
    """One problem description bound to one GPU: owns the `hvi_plan`.  Stands in for the keyword
    set that `get_loss_elbo` captures (src/HybridSolver.jl:240-245)."""

    def __init__(self, prob: HybridProblem, device: int = 0, count_globals: bool = True, lib_path: Optional[str] = None):
        self.lib = A.load_library(lib_path)
        self.prob = prob
        self.desc = prob.to_desc(device, count_globals)
        self.device = device
        h = C.c_void_p()
        rc = self.lib.hvi_plan_create(C.byref(self.desc), C.byref(h))
        if rc != A.HVI_OK:
            raise HVIError(rc, (self.lib.hvi_last_error(None) or b"").decode())
        self._h = h
        self.n_phi = int(self.lib.hvi_phi_len(self._h))
        self.n_site = 0
        self.dt = prob.np_dtype

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.hvi_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow_nonfinite=False):
        if rc == A.HVI_OK or (allow_nonfinite and rc == A.HVI_ENONFINITE):
            return rc
        raise HVIError(rc, (self.lib.hvi_last_error(self._h) or b"").decode())

    # -- index maps ----------------------------------------------------------------------
    def positions(self):
        """0-based (offset, length) of (ϕg, logσ2_ζP, coef_logσ2_ζMs, ρsP, ρsM, μP) as the library
        lays ϕ out -- must equal `prob.positions()` bit for bit."""
        off = (C.c_int64 * 6)()
        ln = (C.c_int64 * 6)()
        self._check(self.lib.hvi_phi_positions(self._h, off, ln))
        names = ("ϕg", "logσ2_ζP", "logσ2_ζMs" if self.prob.is_sepvar else "coef_logσ2_ζMs", "ρsP", "ρsM", "μP")
        return {n: range(off[i], off[i] + ln[i]) for i, n in enumerate(names)}

    @property
    def launch_count(self):
        return int(self.lib.hvi_launch_count(self._h))

    def set_profiling(self, on: bool):
        self._check(self.lib.hvi_plan_set_profiling(self._h, int(on)))

    def kernel_times_ms(self):
        """(prologue, g forward, fused stream, g backward, finalize) of the last profiled evaluation."""
        ms = (C.c_float * 5)()
        self._check(self.lib.hvi_plan_kernel_times(self._h, ms))
        return dict(zip(("prologue", "mlp_fwd", "stream", "mlp_bwd", "finalize"), map(float, ms)))

    # -- data ----------------------------------------------------------------------------
    def set_data(self, xM, xP, y_o=None, y_unc=None, i_sites=None):
        """Register one batch (xM, xP, y_o, y_unc) -- a DataLoader element of the reference
        (src/AbstractHybridProblem.jl:179-184).  numpy arrays (host) or CUDA tensors (device),
        column = site."""
        dev = _is_device(xM)
        if dev:
            _check_colmajor((xM, xP, y_o, y_unc))
            n_site = xM.shape[-1] if xM.dim() == 2 else xM.numel() // self.prob.n_covar
            args = (xM, xP, y_o, y_unc)
        else:
            f = lambda a: None if a is None else np.asfortranarray(a, dtype=self.dt)
            args = tuple(map(f, (xM, xP, y_o, y_unc)))
            n_site = args[0].shape[1]
            assert args[1].shape == (getattr(self, "n_x", 2 * self.prob.n_obs), n_site) and args[0].shape[0] == self.prob.n_covar
            if args[2] is not None:
                assert args[2].shape == (self.prob.n_obs, n_site) and args[3].shape == args[2].shape
        self._keep = args
        self._check(self.lib.hvi_plan_set_data(self._h, n_site, *map(_ptr, args),
                                               A.HVI_MEM_DEVICE if dev else A.HVI_MEM_HOST))
        self.n_site = n_site
        if i_sites is not None:
            # the batch's site indices, 0-based (the `i_sites` of loss_elbo; used by MeanVarSep, src/elbo_sepvec.jl:23)
            idx = np.ascontiguousarray(i_sites, dtype=np.int64)
            self._check(self.lib.hvi_plan_set_sites(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size))

    def set_data_bcast(self, xM, xP, y_o, y_unc, i_sites=None):
        """set_data where xP and / or y_unc may be ONE column shared by all sites (shape (rows, 1) or (rows,)): shared
        drivers (PBMPopulationGlobalApplicator, src/PBMApplicator.jl:305-375) and a constant observation uncertainty
        are uploaded once and expanded on the device.  Host arrays."""
        f = lambda a: np.asfortranarray(np.asarray(a, dtype=self.dt).reshape(a.shape[0], -1))
        xM_, xP_, yo_, yu_ = map(f, (xM, xP, y_o, y_unc))
        n_site = xM_.shape[1]
        nx = getattr(self, "n_x", 2 * self.prob.n_obs)
        assert xP_.shape in ((nx, n_site), (nx, 1)) and yu_.shape in ((self.prob.n_obs, n_site), (self.prob.n_obs, 1))
        assert yo_.shape == (self.prob.n_obs, n_site) and xM_.shape[0] == self.prob.n_covar
        xs, us = int(xP_.shape[1] == 1 and n_site > 1), int(yu_.shape[1] == 1 and n_site > 1)
        self._keep = (xM_, xP_, yo_, yu_)
        self._check(self.lib.hvi_plan_set_data_bcast(self._h, n_site, _ptr(xM_), _ptr(xP_), xs, _ptr(yo_), _ptr(yu_), us,
                                                     A.HVI_MEM_HOST))
        self.n_site = n_site
        if i_sites is not None:
            idx = np.ascontiguousarray(i_sites, dtype=np.int64)
            self._check(self.lib.hvi_plan_set_sites(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size))

    def prefetch_data(self, xM, xP, y_o=None, y_unc=None):
        """Start the upload of the NEXT batch (host arrays, ideally page-locked) while the current one is
        evaluated -- the DataLoader iteration of `solve` (src/HybridSolver.jl:240-260) double-buffered.
        `commit_data()` makes it the current batch.  The arrays must stay unchanged until then."""
        f = lambda a: None if a is None else np.asfortranarray(a, dtype=self.dt)
        args = tuple(map(f, (xM, xP, y_o, y_unc)))
        n_site = args[0].shape[1]
        assert args[1].shape == (2 * self.prob.n_obs, n_site) and args[0].shape[0] == self.prob.n_covar
        if args[2] is not None:
            assert args[2].shape == (self.prob.n_obs, n_site) and args[3].shape == args[2].shape
        self._keep_next = (args, n_site)
        self._check(self.lib.hvi_plan_prefetch_data(self._h, n_site, *map(_ptr, args)))

    def commit_data(self, i_sites=None):
        self._check(self.lib.hvi_plan_commit_data(self._h))
        self._keep, self.n_site = self._keep_next
        if i_sites is not None:
            idx = np.ascontiguousarray(i_sites, dtype=np.int64)
            self._check(self.lib.hvi_plan_set_sites(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size))

    # -- generic process-based model --------------------------------------------------------------
    def set_pbm(self, pbm: str, n_x: Optional[int] = None, fix=()):
        """Replace the DoubleMM process model of this plan (AbstractPBMApplicator, src/PBMApplicator.jl:26-32) by
        "builtin:doubleMM", "builtin:mm1" or CUDA C++ source defining `struct UserPbm` (include/hvi_b200.h).  `n_x`:
        driver rows per site (xP is then n_x × n_site), `fix`: the fixed parameters θFix.  Call set_data afterwards."""
        fx = np.ascontiguousarray(fix, dtype=np.float64)
        nx = 2 * self.prob.n_obs if n_x is None else int(n_x)
        self._check(self.lib.hvi_plan_set_pbm(self._h, pbm.encode(), nx, fx.size,
                                              fx.ctypes.data_as(C.POINTER(C.c_double)) if fx.size else None))
        self.n_x = nx
        self.n_site = 0

    # -- device-resident training set ------------------------------------------------------------
    def set_dataset(self, xM, xP, y_o=None, y_unc=None):
        """Upload the whole training set once (gdev_hybridproblem_dataloader, src/AbstractHybridProblem.jl:220-250);
        minibatches are then chosen on the device by `select_range` / `select_batch` / `adam_run_minibatches`."""
        dev = _is_device(xM)
        if dev:
            _check_colmajor((xM, xP, y_o, y_unc))
            args = (xM, xP, y_o, y_unc)
            n = xM.shape[-1]
        else:
            f = lambda a: None if a is None else np.asfortranarray(a, dtype=self.dt)
            args = tuple(map(f, (xM, xP, y_o, y_unc)))
            n = args[0].shape[1]
            assert args[1].shape == (2 * self.prob.n_obs, n) and args[0].shape[0] == self.prob.n_covar
        self._check(self.lib.hvi_plan_set_dataset(self._h, n, *map(_ptr, args),
                                                  A.HVI_MEM_DEVICE if dev else A.HVI_MEM_HOST))
        self.n_site_dataset = n

    def select_range(self, first, n):
        """Sites [first, first+n) of the dataset become the registered batch (one DataLoader element)."""
        self._check(self.lib.hvi_plan_select_range(self._h, int(first), int(n)))
        self.n_site = int(n)

    def select_batch(self, i_sites):
        """Sites i_sites (0-based; numpy or CUDA int64 tensor) of the dataset become the registered batch."""
        if _is_device(i_sites):
            assert i_sites.dtype.is_floating_point is False and i_sites.element_size() == 8 and i_sites.is_contiguous()
            n = i_sites.numel()
            self._check(self.lib.hvi_plan_select_batch(self._h, _ptr(i_sites), n, A.HVI_MEM_DEVICE))
        else:
            idx = np.ascontiguousarray(i_sites, dtype=np.int64)
            n = idx.size
            self._check(self.lib.hvi_plan_select_batch(self._h, _ptr(idx), n, A.HVI_MEM_HOST))
        self.n_site = int(n)

    def set_stream(self, stream_ptr=None):
        """Make the plan work on the caller's CUDA stream (None: back to the plan's own)."""
        self._check(self.lib.hvi_plan_set_stream(self._h, C.c_void_p(stream_ptr or 0), int(stream_ptr is None)))

    def use_graphs(self, on: bool):
        self._check(self.lib.hvi_plan_use_graphs(self._h, int(on)))

    def fused_phase_clocks(self, start_only=False):
        """SM-clock stamps of the last single-launch evaluation at its phase boundaries (first call arms the stamping)."""
        if start_only:
            self._check(self.lib.hvi_plan_fused_phase_clocks(self._h, None))
            return None
        c = (C.c_int64 * 10)()
        self._check(self.lib.hvi_plan_fused_phase_clocks(self._h, c))
        names = ("gather", "zP_draws", "prologue", "g_fwd", "stream", "g_bwd", "reduce", "finalize", "adam")
        return {n: int(c[i + 1] - c[i]) for i, n in enumerate(names)}

    def use_fused(self, on):
        """Small batches with device draws run as one launch (k_small_iter): False / 0 forces the multi-kernel path, True / 1
        (default) uses the single launch where it is the faster one (n_site <= 48), 2 wherever it is supported (n_site <= 1024)."""
        self._check(self.lib.hvi_plan_use_fused(self._h, int(on)))

    def adam_run_minibatches(self, n_iter, n_MC, batch_size, seed0=0, first_batch=0, trace=True):
        """`solve`'s minibatch loop (src/HybridSolver.jl:259-260) on the device-resident dataset: iteration i takes
        minibatch (first_batch + i) mod (n_site // batch_size), evaluates and applies Adam; CUDA-graph replayed."""
        tr = np.empty(n_iter, dtype=np.float64) if trace else None
        self._check(self.lib.hvi_adam_run_minibatches(self._h, n_iter, n_MC, seed0, int(batch_size), int(first_batch),
                                                      tr.ctypes.data_as(C.POINTER(C.c_double)) if trace else None))
        self.n_site = int(batch_size)
        return tr

    # -- point solver ---------------------------------------------------------------------------
    def loss_gf_withgradient(self, ϕ_gf, want_grad=True):
        """(nLjoint_pen, (nLjoint_pen, nLy, neg_log_prior), ∇ϕ) of `loss_gf` (src/gf.jl:227-295) for ϕ = [ϕg; ϕP] --
        the cost function of `solve(prob, HybridPointSolver)` (src/HybridSolver.jl:9-140) on the registered batch."""
        ϕ_gf = np.ascontiguousarray(ϕ_gf, dtype=self.dt)
        assert ϕ_gf.size == self.prob.n_phig + self.prob.n_P
        comps = (C.c_double * 3)()
        grad = np.empty_like(ϕ_gf) if want_grad else None
        self._check(self.lib.hvi_loss_gf_grad(self._h, _ptr(ϕ_gf), A.HVI_MEM_HOST, comps, _ptr(grad) if want_grad else None))
        return comps[0], np.array(list(comps)), grad

    # -- optimiser steps on the device -------------------------------------------------------
    def adam_init(self, ϕ0, eta=0.02, betas=(0.9, 0.999), eps=1e-8):
        """Device-resident `Adam(eta)` state for `adam_run` (Optimization.solve of src/HybridSolver.jl:259-260)."""
        ϕ0 = np.ascontiguousarray(ϕ0, dtype=self.dt)
        assert ϕ0.size == self.prob.n_phi
        self._check(self.lib.hvi_adam_init(self._h, _ptr(ϕ0), eta, betas[0], betas[1], eps))

    def adam_run(self, n_iter, n_MC, seed0=0, site_offset=0, trace=True):
        """n_iter iterations of (neg-ELBO gradient with device draws seeded seed0+i, Adam update) on the registered
        batch without leaving the GPU; returns the loss of every iteration (or None: nothing is synchronised)."""
        tr = np.empty(n_iter, dtype=np.float64) if trace else None
        self._check(self.lib.hvi_adam_run(self._h, n_iter, n_MC, seed0, site_offset,
                                          tr.ctypes.data_as(C.POINTER(C.c_double)) if trace else None))
        return tr

    def adam_phi_dev(self) -> int:
        """Device address of the optimiser's ϕ (the `phi_dev` of `neg_elbo_async` in sharded jobs)."""
        ptr = C.c_void_p()
        self._check(self.lib.hvi_adam_phi_dev(self._h, C.byref(ptr)))
        return ptr.value

    def adam_apply_dev(self, out_dev, stream_ptr=0):
        """One Adam update from an (all-reduced) device result vector [∇ϕ; components; nL; non-finite] of doubles."""
        self._check(self.lib.hvi_adam_apply_dev(self._h, _ptr(out_dev), C.c_void_p(stream_ptr)))

    def adam_phi(self):
        ϕ = np.empty(self.prob.n_phi, dtype=self.dt)
        self._check(self.lib.hvi_adam_get_phi(self._h, _ptr(ϕ)))
        return ϕ

    # -- the hot path ----------------------------------------------------------------------
    def neg_elbo_withgradient(self, ϕ, zP=None, zMs=None, *, n_MC=None, seed=None, site_offset=0,
                              want_grad=True, allow_nonfinite=False):
        """(nL, components, ∇ϕ) -- `Zygote.withgradient(loss_elbo, ϕ)` of the reference
        (src/HybridSolver.jl:256-258).  Host arrays in, host results out."""
        comps = (C.c_double * 6)()
        nL = C.c_double()
        dev = _is_device(ϕ)
        if dev:
            import torch
            grad = torch.empty_like(ϕ) if want_grad else None
            ϕ_ = ϕ
        else:
            ϕ_ = np.ascontiguousarray(ϕ, dtype=self.dt)
            assert ϕ_.shape == (self.n_phi,), (ϕ_.shape, self.n_phi)
            grad = np.empty(self.n_phi, dtype=self.dt) if want_grad else None
        mem = A.HVI_MEM_DEVICE if dev else A.HVI_MEM_HOST
        if zMs is None:
            assert seed is not None and n_MC is not None, "pass ξ=(zP, zMs) or (n_MC, seed)"
            rc = self.lib.hvi_neg_elbo_grad_rng(self._h, n_MC, _ptr(ϕ_), seed, site_offset, mem, comps, C.byref(nL), _ptr(grad))
        else:
            if dev:
                zP_, zMs_ = zP, zMs
                M = zMs.shape[0] if zMs.dim() == 2 and zMs.stride(0) == 1 else int(n_MC)
            else:
                zP_ = np.asfortranarray(zP, dtype=self.dt)
                zMs_ = np.asfortranarray(zMs, dtype=self.dt)
                M = zMs_.shape[0]
                assert zMs_.shape == (M, self.prob.n_M * self.n_site), zMs_.shape
                assert zP_.shape == (M, self.prob.n_P), zP_.shape
            rc = self.lib.hvi_neg_elbo_grad(self._h, M, _ptr(ϕ_), _ptr(zP_), _ptr(zMs_), mem, comps, C.byref(nL), _ptr(grad))
        self._check(rc, allow_nonfinite)
        return float(nL.value), NegElboComponents(*comps), grad

    def neg_elbo_async(self, ϕ_dev, zP_dev, zMs_dev, out_dev, n_MC, stream_ptr, seed=0, site_offset=0):
        """Stream-ordered device variant (multi-GPU jobs): out_dev (float64, n_phi+8)."""
        self._check(self.lib.hvi_neg_elbo_grad_async(self._h, n_MC, _ptr(ϕ_dev), _ptr(zP_dev), _ptr(zMs_dev), seed,
                                                     site_offset, _ptr(out_dev), C.c_void_p(stream_ptr)))

    # -- site-sharded jobs: the collective behind the C ABI ------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the NCCL communicator of a site-sharded job (`shard.nccl_unique_id()` on rank 0, handed to every rank)."""
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.hvi_plan_comm_init(self._h, buf, rank, world))

    def comm_destroy(self):
        self._check(self.lib.hvi_plan_comm_destroy(self._h))

    def allreduce_result(self, out_dev, stream_ptr=0):
        """Σ over the ranks of out_dev (float64 CUDA tensor, n_phi + 8) on the given stream (0: the plan's)."""
        self._check(self.lib.hvi_allreduce_result(self._h, _ptr(out_dev), C.c_void_p(stream_ptr)))

    def neg_elbo_withgradient_sharded(self, ϕ, zP=None, zMs=None, *, n_MC=None, seed=0, site_offset=0, allow_nonfinite=False):
        """(nL, components, ∇ϕ) of the whole job on every rank: this rank's shard is evaluated, the result vector is summed
        over the ranks by the library (ncclAllReduce on the evaluation's stream).  zMs: this rank's columns of the draws
        (host arrays), or None for device draws keyed by the global site index (`site_offset` = first site of the shard)."""
        comps = (C.c_double * 6)()
        nL = C.c_double()
        ϕ_ = np.ascontiguousarray(ϕ, dtype=self.dt)
        grad = np.empty_like(ϕ_)
        if zMs is None:
            assert n_MC is not None
            rc = self.lib.hvi_neg_elbo_grad_sharded(self._h, int(n_MC), _ptr(ϕ_), None, None, seed, site_offset, A.HVI_MEM_HOST,
                                                    comps, C.byref(nL), _ptr(grad))
        else:
            zP_ = np.asfortranarray(zP, dtype=self.dt)
            zMs_ = np.asfortranarray(zMs, dtype=self.dt)
            M = zMs_.shape[0]
            assert zMs_.shape == (M, self.prob.n_M * self.n_site), zMs_.shape
            rc = self.lib.hvi_neg_elbo_grad_sharded(self._h, M, _ptr(ϕ_), _ptr(zP_), _ptr(zMs_), 0, site_offset, A.HVI_MEM_HOST,
                                                    comps, C.byref(nL), _ptr(grad))
        self._check(rc, allow_nonfinite)
        return float(nL.value), NegElboComponents(*comps), grad

    def generate_ζ(self, ϕ, zP, zMs):
        """generate_ζ (src/elbo.jl:472-521): ζsP (n_θP×n_MC), ζsMs_tr (n_site×n_θM×n_MC), σ."""
        ϕ_ = np.ascontiguousarray(ϕ, dtype=self.dt)
        zP_ = np.asfortranarray(zP, dtype=self.dt)
        zMs_ = np.asfortranarray(zMs, dtype=self.dt)
        M = zMs_.shape[0]
        nP, nM, nb = self.prob.n_P, self.prob.n_M, self.n_site
        ζP = np.empty((nP, M), dtype=self.dt, order="F")
        ζMs = np.empty((nb, nM, M), dtype=self.dt, order="F")
        σ = np.empty(nP + nM * nb, dtype=self.dt)
        self._check(self.lib.hvi_generate_zeta(self._h, M, _ptr(ϕ_), _ptr(zP_), _ptr(zMs_), A.HVI_MEM_HOST,
                                               _ptr(ζP), _ptr(ζMs), _ptr(σ)))
        return ζP, ζMs, σ

    def predict_hvi(self, ϕ, zP=None, zMs=None, *, n_sample_pred=200, seed=None, site_offset=0):
        """predict_hvi (src/elbo.jl:320-352): returns (y, θsP, θsMs_tr, entropy_ζ) with
        y (n_obs × n_site × n_sample), θsP (n_θP × n_sample), θsMs_tr (n_site × n_θM × n_sample)."""
        ϕ_ = np.ascontiguousarray(ϕ, dtype=self.dt)
        nP, nM, nb, nO = self.prob.n_P, self.prob.n_M, self.n_site, self.prob.n_obs
        if zMs is not None:
            zP_ = np.asfortranarray(zP, dtype=self.dt)
            zMs_ = np.asfortranarray(zMs, dtype=self.dt)
            M = zMs_.shape[0]
        else:
            assert seed is not None
            zP_ = zMs_ = None
            M = int(n_sample_pred)
        θP = np.empty((nP, M), dtype=self.dt, order="F")
        θMs = np.empty((nb, nM, M), dtype=self.dt, order="F")
        y = np.empty((nO, nb, M), dtype=self.dt, order="F")
        ent = C.c_double()
        self._check(self.lib.hvi_predict(self._h, M, _ptr(ϕ_), _ptr(zP_), _ptr(zMs_), int(seed or 0), site_offset,
                                         A.HVI_MEM_HOST, _ptr(θP), _ptr(θMs), _ptr(y), C.byref(ent)))
        return y, θP, θMs, float(ent.value)

    def randn(self, n_MC: int, n_cols: int, seed: int, col_offset: int = 0):
        """The device generator's draws (n_MC × n_cols), column = global site·n_θM + k -- what
        `neg_elbo_withgradient(seed=…)` uses (stand-in for CUDA.randn)."""
        out = np.empty((n_MC, n_cols), dtype=self.dt, order="F")
        self._check(self.lib.hvi_randn(self._h, n_MC, n_cols, seed, col_offset, A.HVI_MEM_HOST, _ptr(out)))
        return out

    def apply_g(self, ϕ):
        """g(xM, ϕg) → μ_ζMs (n_θM × n_site) (src/ModelApplicator.jl:19-23,163-176)."""
        ϕ_ = np.ascontiguousarray(ϕ, dtype=self.dt)
        mu = np.empty((self.prob.n_M, self.n_site), dtype=self.dt, order="F")
        self._check(self.lib.hvi_apply_g(self._h, _ptr(ϕ_), A.HVI_MEM_HOST, _ptr(mu)))
        return mu


# ---- functional spellings that read like the reference's -----------------------------------
def pbm_compile_check(pbm: str, prob: HybridProblem, n_x: Optional[int] = None):
    """NVRTC-compile a process model for `prob`'s sizes without a device; returns (ok, compiler message)."""
    lib = A.load_library()
    log = C.create_string_buffer(4096)
    rc = lib.hvi_pbm_compile_check(pbm.encode(), A.HVI_F32 if prob.dtype == "f32" else A.HVI_F64, prob.n_P, prob.n_M,
                                   prob.n_obs, 2 * prob.n_obs if n_x is None else int(n_x), log, 4096)
    return rc == A.HVI_OK, log.value.decode(errors="replace")


def neg_elbo_gtf_components(plan: NegElboPlan, ϕ, zP, zMs) -> NegElboComponents:
    """src/elbo.jl:46-88"""
    return plan.neg_elbo_withgradient(ϕ, zP, zMs, want_grad=False)[1]


def neg_elbo_gtf(plan: NegElboPlan, ϕ, zP, zMs) -> float:
    """src/elbo.jl:30-44: nL = nLjoint − entropy_ζ + loss_penalty (a Float64 even for Float32 ϕ)."""
    return plan.neg_elbo_withgradient(ϕ, zP, zMs, want_grad=False)[0]


def get_loss_elbo(prob: HybridProblem, device: int = 0, n_MC: int = 12):
    """`get_loss_elbo` (src/HybridSolver.jl:293-324): returns
    ``loss_elbo(ϕ, rng_seed, xM, xP, y_o, y_unc, i_sites) -> (nL, components, ∇ϕ)``.

    Every call registers its batch (set_data is a copy + one re-tile kernel): arrays refilled in place by a loader
    are therefore always seen.  Pass ``batch_token`` (any hashable that changes when the batch does) to skip the
    re-registration of an unchanged batch.  `i_sites` (0-based) is required with MeanVarSepHVIApproximation, whose
    per-site log-variances are addressed through it (src/elbo_sepvec.jl:23)."""
    plan = NegElboPlan(prob, device)
    state = {"token": None}

    def loss_elbo(ϕ, rng_seed, xM, xP, y_o, y_unc, i_sites=None, *, batch_token=None):
        if prob.is_sepvar and i_sites is None:
            raise ValueError("MeanVarSepHVIApproximation needs i_sites (src/elbo_sepvec.jl:23)")
        if batch_token is None or batch_token != state["token"]:
            plan.set_data(xM, xP, y_o, y_unc, i_sites=i_sites if prob.is_sepvar else None)
            state["token"] = batch_token
        return plan.neg_elbo_withgradient(ϕ, n_MC=n_MC, seed=int(rng_seed))

    loss_elbo.plan = plan
    return loss_elbo


# ---- integer index maps exported by the library (bit-exact contract) ------------------------
def vec2utri_pos(s: int):
    """src/cholesky.jl:30-37 (one-based)"""
    lib = A.load_library()
    r, c = C.c_int64(), C.c_int64()
    if lib.hvi_vec2utri_pos(s, C.byref(r), C.byref(c)) != 0:
        raise ValueError(s)
    return int(r.value), int(c.value)


def utri2vec_pos(row: int, col: int) -> int:
    """src/cholesky.jl:42-50"""
    v = A.load_library().hvi_utri2vec_pos(row, col)
    if v < 0:
        raise AssertionError("row <= col")
    return int(v)


def get_cor_count(cor_ends) -> int:
    """src/cholesky.jl:233-247"""
    arr = (C.c_int32 * max(1, len(cor_ends)))(*cor_ends)
    return int(A.load_library().hvi_cor_count(arr, len(cor_ends)))
