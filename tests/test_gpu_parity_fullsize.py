"""Oracle parity at the sizes BASELINE.json quotes (configs 2, 3, 3', the 8-GPU shard of config 4 and the shape of
config 5): the CUDA path through the C-ABI, Float32 AND Float64, against the float64 C restatement of the reference
(oracle/hvi_oracle.c, all host threads) on identical inputs and identical standard-normal draws ξ -- value, the six
components and every ϕ block (reference path: src/elbo.jl:46-88, its Zygote gradient src/HybridSolver.jl:256-258).

Tolerances (BASELINE.json north_star): rel 1e-10 (Float64), 1e-4 (Float32); "rel" per ϕ block is
max|g − g_ref| / max|g_ref|.  The C restatement itself is re-pinned at size against the independent float64
torch-autograd oracle (oracle/hvi_oracle.py) in `test_c_port_against_autograd_at_size` (no GPU needed).
"""
import os

import numpy as np
import pytest

from helpers import O, c_oracle_value_and_grad, hvi_b200, oracle_spec

TOL = {"f64": 1e-10, "f32": 1e-4}
NTHREADS = os.cpu_count() or 1

# name: (scenario, hidden, n_site, n_MC, NaN-driver fraction)
CONFIGS = {
    # config 2: DoubleMM 10^5 sites, n_MC = 32
    "C2_1e5x32": ((), None, 100_000, 32, 0.0),
    # config 3: 10^6 sites, full P and M correlation blocks with ρ ≠ 0 (perturbed ϕ), 10 % of the sites with missing drivers
    "C3_1e6x32_nan": ((), None, 1_000_000, 32, 0.1),
    # config 3': θM–θP dependence through pbm_covars = (:K2,) (docs/src/tutorials/corr_site_global.md): g per (site, draw)
    "C3p_covarK2_2e5x32_nan": (("covarK2",), None, 200_000, 32, 0.1),
    # config 4: 10^7 sites x 64 draws, the shard one of 8 GPUs owns
    "C4_shard_1.25e6x64": ((), None, 1_250_000, 64, 0.0),
    # config 5 shape: 4 x 512 hidden, pbm covariate so that g runs on sites x draws columns
    "C5_shape_4x512_1.5e3x8": (("covarK2",), (512, 512, 512, 512), 1_500, 8, 0.0),
}

_oracle_cache = {}


def _inputs_and_oracle(name):
    """inputs representable in both element types + the float64 C-port result (computed once per config)"""
    if name in _oracle_cache:
        return _oracle_cache[name]
    scen, hidden, nb, M, nan_frac = CONFIGS[name]
    p64 = hvi_b200.DoubleMMCase(scen, dtype="f64", hidden=hidden)
    p32 = hvi_b200.DoubleMMCase(scen, dtype="f32", hidden=hidden)
    d32 = hvi_b200.gen_hybridproblem_synthetic(p32, nb, seed=111, driver_nan_frac=nan_frac)   # float32 values ...
    data = {k: np.asfortranarray(d32[k].astype(np.float64)) for k in ("xM", "xP", "y_o", "y_unc")}   # ... exact in float64
    zP, zMs = hvi_b200.draw_ξ(p32, nb, M)
    ϕ = p64.init_ϕ(perturbed=True).astype(np.float32).astype(np.float64)
    ref = c_oracle_value_and_grad(p64, ϕ, data, zP.astype(np.float64), zMs.astype(np.float64), nthreads=NTHREADS)
    _oracle_cache.clear()          # one config's arrays at a time (config 4's draws alone are 1.3 GB in float64)
    _oracle_cache[name] = (data, zP, zMs, ϕ, ref)
    return _oracle_cache[name]


def _compare(prob, got, ref, tol):
    nL, comps, g = got
    nL_o, comps_o, g_o = ref
    assert abs(nL - nL_o) <= tol * abs(nL_o), (nL, nL_o)
    assert np.max(np.abs(np.array(comps) - comps_o)) <= tol * np.max(np.abs(comps_o)), (comps, comps_o)
    for blk, r in prob.positions().items():
        if len(r) == 0:
            continue
        err = np.max(np.abs(np.asarray(g[r], dtype=np.float64) - g_o[r]))
        scale = np.max(np.abs(g_o[r]))
        assert err <= tol * scale, (blk, err, scale)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_baseline_size_parity(name, dtype):
    scen, hidden, nb, M, _ = CONFIGS[name]
    if hidden is not None and dtype == "f64":
        pytest.skip("the wide (tensor-core) applicator is the Float32 path of config 5")
    data, zP, zMs, ϕ, ref = _inputs_and_oracle(name)
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype, hidden=hidden)
    ft = prob.np_dtype
    plan = hvi_b200.NegElboPlan(prob, device=0)
    try:
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        got = plan.neg_elbo_withgradient(ϕ.astype(ft), zP.astype(ft), zMs.astype(ft))
    finally:
        plan.close()
    assert np.isfinite(got[0]) and np.all(np.isfinite(got[2]))
    _compare(prob, got, ref, TOL[dtype])


@pytest.mark.parametrize("scen", [(), ("covarK2",)])
def test_c_port_against_autograd_at_size(scen):
    """The C restatement (hand adjoint) against the independent float64 torch-autograd oracle at 2·10^4 sites:
    the checker of the BASELINE-size tests is itself checked at size.  CPU only."""
    nb, M = (20_000, 16) if not scen else (5_000, 8)
    prob = hvi_b200.DoubleMMCase(scen, dtype="f64")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=7, driver_nan_frac=0.1)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ϕ = prob.init_ϕ(perturbed=True)
    ref = O.value_and_grad(oracle_spec(prob), ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs, "f64")
    got = c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=NTHREADS)
    _compare(prob, got, ref, 1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("block_sites", [512, 1024])
def test_site_blocked_wide_evaluation(block_sites, monkeypatch):
    """Config 5 shape with the evaluation walking blocks of sites (forward → stream → backward per block, the schedule that
    keeps the activations of pass 2 in HBM at 10^6 sites x 64 draws): against the oracle, and equal to the one-block
    schedule up to the order of the float sums; passed-in draws and device draws."""
    name = "C5_shape_4x512_1.5e3x8"
    scen, hidden, nb, M, _ = CONFIGS[name]
    data, zP, zMs, ϕ, ref = _inputs_and_oracle(name)
    prob = hvi_b200.DoubleMMCase(scen, dtype="f32", hidden=hidden)
    ft = prob.np_dtype

    def run(seed=None):
        plan = hvi_b200.NegElboPlan(prob, device=0)
        try:
            plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
            if seed is None:
                return plan.neg_elbo_withgradient(ϕ.astype(ft), zP.astype(ft), zMs.astype(ft)), plan.launch_count
            return plan.neg_elbo_withgradient(ϕ.astype(ft), n_MC=M, seed=seed), plan.launch_count
        finally:
            plan.close()

    whole, n_whole = run()
    whole_rng, _ = run(seed=5)
    monkeypatch.setenv("HVI_WIDE_BLOCK_SITES", str(block_sites))
    blocked, n_blocked = run()
    blocked_rng, _ = run(seed=5)
    assert n_blocked > n_whole          # the blocks really were walked
    _compare(prob, blocked, ref, TOL["f32"])
    _compare(prob, blocked, whole, 2e-5)
    _compare(prob, blocked_rng, whole_rng, 2e-5)


@pytest.mark.gpu
def test_wide_backward_fused_kernels_against_their_unfused_forms(monkeypatch):
    """The one-pass last-layer backward kernel (k_last_bwd_fused_bf3) and the bias gradients taken as column sums in the
    epilogue of the data-gradient GEMM, against the kernels they replace (k_last_bf3<1> + k_thin_bwd_bf3 + skinny
    tensor-core products): every entry of ∇ϕg, relative to the largest entry of its layer's weights or biases."""
    name = "C5_shape_4x512_1.5e3x8"
    scen, hidden, nb, M, _ = CONFIGS[name]
    data, zP, zMs, ϕ, ref = _inputs_and_oracle(name)
    prob = hvi_b200.DoubleMMCase(scen, dtype="f32", hidden=hidden)
    ft = prob.np_dtype

    def run():
        plan = hvi_b200.NegElboPlan(prob, device=0)
        try:
            plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
            return plan.neg_elbo_withgradient(ϕ.astype(ft), zP.astype(ft), zMs.astype(ft))
        finally:
            plan.close()

    fused = run()
    monkeypatch.setenv("HVI_WIDE_LAST_UNFUSED", "1")
    monkeypatch.setenv("HVI_WIDE_BIAS_SKINNY", "1")
    unfused = run()
    _compare(prob, fused, ref, TOL["f32"])
    g, g_u, g_o = (np.asarray(x[2], dtype=np.float64) for x in (fused, unfused, ref))
    off = 0
    for n_in, n_out, _, has_bias in prob.layers:           # per layer: weights [in][out], then the biases
        for n in (n_in * n_out, n_out if has_bias else 0):
            if n:
                r = slice(off, off + n)
                scale = np.max(np.abs(g_o[r]))
                assert np.max(np.abs(g[r] - g_u[r])) <= 2e-5 * scale, (off, n)
                assert np.max(np.abs(g[r] - g_o[r])) <= TOL["f32"] * scale, (off, n)
                off += n


@pytest.mark.gpu
def test_site_blocks_chosen_from_the_cache_size(monkeypatch):
    """The block size the library derives itself from the memory it may take for the activation cache (here capped with
    HVI_WIDE_CACHE_MB so that 10 000 sites x 4 draws need two blocks): same result as the one-block schedule."""
    nb, M = 10_000, 4
    prob = hvi_b200.DoubleMMCase(("covarK2",), dtype="f32", hidden=(512, 512, 512, 512))
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=5)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    ft = prob.np_dtype
    ϕ = prob.init_ϕ(perturbed=True).astype(ft)

    def run():
        plan = hvi_b200.NegElboPlan(prob, device=0)
        try:
            plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
            return plan.neg_elbo_withgradient(ϕ, zP.astype(ft), zMs.astype(ft)), plan.launch_count
        finally:
            plan.close()

    whole, n_whole = run()
    monkeypatch.setenv("HVI_WIDE_CACHE_MB", "160")     # 160 MB / (8 KB per column x 4 draws) = 5 120 sites per block
    blocked, n_blocked = run()
    assert n_blocked > n_whole
    assert np.isfinite(blocked[0]) and np.all(np.isfinite(blocked[2]))
    _compare(prob, blocked, whole, 2e-5)
