"""Device-resident training set, minibatch selection on the device, the graph-replayed minibatch loop and the stream
contract of include/hvi_b200.h.

Reference behaviour mirrored: gdev_hybridproblem_dataloader moves the whole loader to the GPU once
(src/AbstractHybridProblem.jl:220-250); solve walks MLUtils.DataLoader(...; batchsize = n_batch, partial = false)
(:206-207) -- consecutive column ranges, no shuffling -- and applies Adam per minibatch (src/HybridSolver.jl:256-260)."""
import numpy as np
import pytest

import hvi_b200

pytestmark = pytest.mark.gpu

KEYS = ("xM", "xP", "y_o", "y_unc")


def _setup(dtype="f64", n=230, scen=(), nan_frac=0.2, seed=5):
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, n, seed=seed, driver_nan_frac=nan_frac)
    return prob, data


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("scen", [(), ("covarK2",)])
def test_select_equals_set_data(scen, dtype):
    """a minibatch selected on the device gives bit-identical results to registering the same columns from the host"""
    prob, data = _setup(dtype, 230, scen)
    M = 8
    ϕ = prob.init_ϕ(perturbed=True)
    ds = hvi_b200.NegElboPlan(prob)
    ds.set_dataset(*(data[k] for k in KEYS))
    ref = hvi_b200.NegElboPlan(prob)
    rng = np.random.default_rng(1)
    for sel in (("range", 0, 230), ("range", 37, 64), ("range", 229, 1), ("idx", rng.permutation(230)[:50]),
                ("idx", np.array([7, 7, 3, 229, 0]))):
        if sel[0] == "range":
            idx = np.arange(sel[1], sel[1] + sel[2])
            ds.select_range(sel[1], sel[2])
        else:
            idx = sel[1]
            ds.select_batch(idx)
        ref.set_data(*(data[k][:, idx] for k in KEYS))
        zP, zMs = hvi_b200.draw_ξ(prob, len(idx), M)
        a = ds.neg_elbo_withgradient(ϕ, zP, zMs)
        b = ref.neg_elbo_withgradient(ϕ, zP, zMs)
        assert a[0] == b[0] and tuple(a[1]) == tuple(b[1]) and np.array_equal(a[2], b[2])
        # the prediction path reads the gathered raw drivers
        ya = ds.predict_hvi(ϕ, zP, zMs)[0]
        yb = ref.predict_hvi(ϕ, zP, zMs)[0]
        assert np.array_equal(ya, yb, equal_nan=True)


def test_select_batch_device_indices_and_errors():
    import torch
    prob, data = _setup("f32", 100)
    plan = hvi_b200.NegElboPlan(prob)
    with pytest.raises(hvi_b200.HVIError):
        plan.select_range(0, 10)                       # no dataset yet
    plan.set_dataset(*(data[k] for k in KEYS))
    with pytest.raises(hvi_b200.HVIError):
        plan.select_range(95, 10)                      # runs past the end
    with pytest.raises(hvi_b200.HVIError):
        plan.select_batch(np.array([1, 100]))          # index out of range
    idx = np.array([5, 99, 0, 42], dtype=np.int64)
    zP, zMs = hvi_b200.draw_ξ(prob, 4, 6)
    ϕ = prob.init_ϕ()
    plan.select_batch(idx)
    a = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    plan.select_batch(torch.as_tensor(idx, device="cuda"))
    b = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    assert a[0] == b[0] and np.array_equal(a[2], b[2])


def test_dataset_from_device_tensors_and_layout_check():
    import torch
    prob, data = _setup("f32", 64, nan_frac=0.0)
    plan = hvi_b200.NegElboPlan(prob)
    col = lambda a: torch.as_tensor(np.ascontiguousarray(a.T), device="cuda").T     # column-major view (rows, n)
    plan.set_dataset(*(col(data[k]) for k in KEYS))
    torch.cuda.synchronize()
    plan.select_range(0, 64)
    ref = hvi_b200.NegElboPlan(prob)
    ref.set_data(*(data[k] for k in KEYS))
    zP, zMs = hvi_b200.draw_ξ(prob, 64, 4)
    ϕ = prob.init_ϕ()
    assert plan.neg_elbo_withgradient(ϕ, zP, zMs)[0] == ref.neg_elbo_withgradient(ϕ, zP, zMs)[0]
    # a C-contiguous (rows, n_site) device tensor would silently be read transposed: refused
    with pytest.raises(ValueError):
        plan.set_data(*(torch.as_tensor(np.ascontiguousarray(data[k]), device="cuda") for k in KEYS))
    with pytest.raises(ValueError):
        plan.set_dataset(*(torch.as_tensor(np.ascontiguousarray(data[k]), device="cuda") for k in KEYS))


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_select_with_meanvarsep(dtype):
    """MeanVarSep: the selected dataset indices are the batch's i_sites (src/elbo_sepvec.jl:23)"""
    prob = hvi_b200.DoubleMMCase(("sepvar", "few_sites"), dtype=dtype)        # n_site = 100
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 100, seed=2)
    ϕ = prob.init_ϕ(perturbed=True)
    ϕ[prob.positions()["logσ2_ζMs"]] += np.random.default_rng(0).normal(0, 0.3, 200).astype(prob.np_dtype)
    ds = hvi_b200.NegElboPlan(prob)
    ds.set_dataset(*(data[k] for k in KEYS))
    ref = hvi_b200.NegElboPlan(prob)
    for idx in (np.arange(40, 60), np.array([99, 3, 57, 20, 21])):
        if idx[1] == idx[0] + 1:
            ds.select_range(int(idx[0]), len(idx))
        else:
            ds.select_batch(idx)
        ref.set_data(*(data[k][:, idx] for k in KEYS), i_sites=idx)
        zP, zMs = hvi_b200.draw_ξ(prob, len(idx), 5)
        a, b = ds.neg_elbo_withgradient(ϕ, zP, zMs), ref.neg_elbo_withgradient(ϕ, zP, zMs)
        assert a[0] == b[0] and np.array_equal(a[2], b[2])
        assert np.count_nonzero(a[2][prob.positions()["logσ2_ζMs"]]) == 2 * len(idx)


def test_get_loss_elbo_meanvarsep_minibatches():
    """loss_elbo(ϕ, rng, xM, xP, y_o, y_unc, i_sites) passes i_sites on (src/HybridSolver.jl:312-322): every minibatch
    updates its own sites' log-variances; a missing i_sites is an error; a batch refilled in place is seen"""
    prob = hvi_b200.DoubleMMCase(("sepvar", "few_sites"), dtype="f64")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 100, seed=2)
    ϕ = prob.init_ϕ(perturbed=True)
    loss = hvi_b200.get_loss_elbo(prob, n_MC=4)
    r = prob.positions()["logσ2_ζMs"]
    bufs = [np.asfortranarray(data[k][:, :20].copy()) for k in KEYS]
    for b in range(3):
        idx = np.arange(20 * b, 20 * b + 20)
        for buf, k in zip(bufs, KEYS):
            buf[...] = data[k][:, idx]                          # the loader reuses its buffers
        nL, comps, g = loss(ϕ, 7, *bufs, idx)
        nz = np.flatnonzero(g[r])
        assert nz.min() == 2 * idx[0] and nz.max() == 2 * idx[-1] + 1
        ref = hvi_b200.NegElboPlan(prob)
        ref.set_data(*(data[k][:, idx] for k in KEYS), i_sites=idx)
        assert ref.neg_elbo_withgradient(ϕ, n_MC=4, seed=7)[0] == nL
    with pytest.raises(ValueError):
        loss(ϕ, 7, *bufs)


def _host_minibatch_adam(plan, ϕ, n_iter, M, bs, seed0, eta, first_batch=0, b1=0.9, b2=0.999, eps=1e-8):
    n_batches = plan.n_site_dataset // bs
    ϕ = ϕ.astype(np.float64).copy()
    m, v = np.zeros_like(ϕ), np.zeros_like(ϕ)
    trace, t = [], 0
    for it in range(n_iter):
        b = (first_batch + it) % n_batches
        plan.select_range(b * bs, bs)
        try:
            nL, _, g = plan.neg_elbo_withgradient(ϕ.astype(plan.dt), n_MC=M, seed=seed0 + it)
        except hvi_b200.HVIError as e:       # the reference raises / returns NaN there: the iteration is skipped
            assert e.code == hvi_b200._abi.HVI_ENONFINITE
            trace.append(np.nan)
            continue
        t += 1
        g = g.astype(np.float64)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        ϕ = (ϕ.astype(plan.dt) - eta * (m / (1 - b1 ** t)) / (np.sqrt(v / (1 - b2 ** t)) + eps)).astype(plan.dt).astype(np.float64)
        trace.append(nL)
    return ϕ, np.array(trace)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-10), ("f32", 3e-4)])
@pytest.mark.parametrize("scen", [(), ("covarK2",), ("sepvar", "few_sites")])
def test_adam_run_minibatches_matches_host_loop(scen, dtype, tol):
    """the reference's operating point -- 20-site minibatches, few draws -- walked on the device (one captured CUDA
    graph per iteration) equals the host loop batch by batch; with and without graph replay bit-identical"""
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
    n = 100
    data = hvi_b200.gen_hybridproblem_synthetic(prob, n, seed=3)
    ϕ0 = prob.init_ϕ()
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_dataset(*(data[k] for k in KEYS))
    bs, M, n_iter = 20, 3, 13       # 5 minibatches: the walk wraps around twice
    want_ϕ, want_trace = _host_minibatch_adam(plan, ϕ0, n_iter, M, bs, 11, 0.02, first_batch=2)
    results = []
    for graphs in (True, False):
        plan.use_graphs(graphs)
        plan.adam_init(ϕ0, eta=0.02)
        n0 = plan.launch_count
        tr = plan.adam_run_minibatches(n_iter, M, bs, seed0=11, first_batch=2)
        results.append((tr, plan.adam_phi(), plan.launch_count - n0))
    (tr, got, _), (tr2, got2, _) = results
    assert np.array_equal(tr, tr2) and np.array_equal(got, got2)
    assert np.all(np.isfinite(tr))
    assert np.max(np.abs(tr - want_trace)) <= tol * np.max(np.abs(want_trace))
    assert np.max(np.abs(got.astype(np.float64) - want_ϕ)) <= tol * max(1.0, np.max(np.abs(want_ϕ)))
    # a second run continues the trajectory (step count and moments persist) and reuses the captured graph
    tr3 = plan.adam_run_minibatches(4, M, bs, seed0=100)
    assert np.all(np.isfinite(tr3))


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_optimiser_loop_inside_the_launch_equals_one_launch_per_iteration(dtype, monkeypatch):
    """k_small_iter runs the iterations of a call as a loop inside ONE launch (counters, seed and minibatch index in device
    memory); HVI_SMALL_NO_LOOP=1 launches it once per iteration: bit-identical loss trace and parameters, fewer launches"""
    prob = hvi_b200.DoubleMMCase(dtype=dtype)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 100, seed=3)
    ϕ0 = prob.init_ϕ()
    bs, M, n_iter = 20, 3, 23

    def run():
        plan = hvi_b200.NegElboPlan(prob)
        try:
            plan.set_dataset(*(data[k] for k in KEYS))
            plan.adam_init(ϕ0, eta=0.02)
            n0 = plan.launch_count
            tr = plan.adam_run_minibatches(n_iter, M, bs, seed0=5, first_batch=1)
            tr_b = plan.adam_run_minibatches(7, M, bs, seed0=50)          # a second call continues the trajectory
            return np.concatenate([tr, tr_b]), plan.adam_phi(), plan.launch_count - n0
        finally:
            plan.close()

    tr, ϕ, n_loop = run()
    monkeypatch.setenv("HVI_SMALL_NO_LOOP", "1")
    tr1, ϕ1, n_single = run()
    assert np.all(np.isfinite(tr))
    assert np.array_equal(tr, tr1) and np.array_equal(ϕ, ϕ1)
    assert n_loop < n_single and n_loop <= 8 and n_single >= n_iter + 7


def test_adam_skips_anomalous_minibatch_without_counting_it():
    """a finite observation under a missing driver makes the reference's loss NaN: on the device
    loop that minibatch is skipped (NaN in the trace) and does not advance the Adam step count"""
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    data = hvi_b200.gen_hybridproblem_synthetic(prob, 60, seed=3)
    data["xP"][0, 25] = np.nan                               # site 25 (minibatch 1): driver missing, observation finite
    ϕ0 = prob.init_ϕ()
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_dataset(*(data[k] for k in KEYS))
    plan.adam_init(ϕ0, eta=0.02)
    tr = plan.adam_run_minibatches(7, 3, 20, seed0=5)
    assert list(np.isnan(tr)) == [False, True, False, False, True, False, False]
    got = plan.adam_phi()
    want_ϕ, want_tr = _host_minibatch_adam(plan, ϕ0, 7, 3, 20, 5, 0.02)
    assert np.array_equal(np.isnan(want_tr), np.isnan(tr))
    assert np.max(np.abs(got - want_ϕ)) <= 1e-10 * max(1.0, np.max(np.abs(want_ϕ)))
    # the single-batch loop refuses such a batch up front once the count is known
    plan.select_range(20, 20)
    with pytest.raises(hvi_b200.HVIError):
        plan.neg_elbo_withgradient(ϕ0, n_MC=3, seed=1)
    with pytest.raises(hvi_b200.HVIError):
        plan.adam_run(2, 3)


def test_stream_contract_without_host_synchronisation():
    """async evaluation and Adam update on a caller's stream, interleaved with calls on the plan's own stream, with no
    torch.cuda.synchronize() in between: the library orders the streams itself (include/hvi_b200.h, stream contract)"""
    import torch
    prob = hvi_b200.DoubleMMCase(dtype="f64")
    n, M, steps = 3000, 8, 5
    data = hvi_b200.gen_hybridproblem_synthetic(prob, n, seed=8)
    ϕ0 = prob.init_ϕ(perturbed=True)
    whole = hvi_b200.NegElboPlan(prob)
    whole.set_data(*(data[k] for k in KEYS))
    whole.adam_init(ϕ0, eta=0.01)
    want_trace = whole.adam_run(steps, M, seed0=21)
    want = whole.adam_phi()

    plan = hvi_b200.NegElboPlan(prob)
    plan.set_dataset(*(data[k] for k in KEYS))
    plan.adam_init(ϕ0, eta=0.01)
    side = torch.cuda.Stream()
    out = torch.zeros(prob.n_phi + 8, dtype=torch.float64, device="cuda")
    outs = []
    for it in range(steps):
        plan.select_range(0, n)                                   # plan stream: rewrites the batch buffers
        plan.neg_elbo_async(plan.adam_phi_dev(), None, None, out, M, side.cuda_stream, seed=21 + it)   # caller's stream
        plan.adam_apply_dev(out, side.cuda_stream)
        with torch.cuda.stream(side):
            outs.append(out[prob.n_phi + 6].clone())
    got = plan.adam_phi()                                         # plan stream again: must see the last update
    torch.cuda.synchronize()
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    assert np.allclose([float(o) for o in outs], want_trace, rtol=1e-12)
    # the plan can also be moved onto the caller's stream altogether
    plan.set_stream(side.cuda_stream)
    with torch.cuda.stream(side):
        ϕd = torch.as_tensor(ϕ0, device="cuda") * 1.0             # produced on `side`, consumed by the plan on `side`
        nL, _, g = plan.neg_elbo_withgradient(ϕd, n_MC=M, seed=3)
    plan.set_stream(None)
    nL2, _, g2 = plan.neg_elbo_withgradient(ϕ0, n_MC=M, seed=3)
    assert nL == nL2 and np.array_equal(g.cpu().numpy(), g2)
