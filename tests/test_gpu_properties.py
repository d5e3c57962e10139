"""Size-independent properties of the CUDA path at sizes far beyond what the float64 oracle can
check in seconds (BASELINE configs 2-4): shard additivity, invariance under duplicated draws and
under a permutation of the sites, directional derivative against central differences, bit
reproducibility.  Through the C ABI, on the GPU."""
import numpy as np
import pytest

from helpers import hvi_b200

pytestmark = pytest.mark.gpu


def _setup(dtype, nb, M, scen=(), nan=0.1, seed=111):
    prob = hvi_b200.DoubleMMCase(scen, dtype=dtype)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=seed, driver_nan_frac=nan)
    zP, zMs = hvi_b200.draw_ξ(prob, nb, M)
    return prob, data, zP, zMs, prob.init_ϕ(perturbed=True)


def _eval(prob, data, ϕ, zP, zMs, count_globals=True):
    plan = hvi_b200.NegElboPlan(prob, count_globals=count_globals)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    out = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    plan.close()
    return out


def test_config3_size_shards_add_up_and_reproduce():
    """10^6 sites x 32 draws, Float32, 10 % missing drivers (BASELINE config 3): two site shards sum to
    the whole (what the all-reduce relies on) and a second evaluation is bit-identical."""
    from hvi_b200.shard import shard_batch
    prob, data, zP, zMs, ϕ = _setup("f32", 1_000_000, 32)
    nL, comps, g = _eval(prob, data, ϕ, zP, zMs)
    nL2, comps2, g2 = _eval(prob, data, ϕ, zP, zMs)
    assert nL == nL2 and np.array_equal(g, g2)
    tot_c, tot_g = np.zeros(6), np.zeros(prob.n_phi)
    for r in range(2):
        d, z, lo = shard_batch(data, zMs, prob.n_M, r, 2)
        _, c, gg = _eval(prob, d, ϕ, zP, z, count_globals=(r == 0))
        tot_c += np.array(c)
        tot_g += gg
    np.testing.assert_allclose(tot_c, np.array(comps), rtol=1e-6)
    for name, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(tot_g[r] - g[r])) <= 2e-5 * np.max(np.abs(g[r])), name


def test_duplicated_draws_leave_the_estimate_unchanged():
    """the ELBO is a mean over draws: evaluating on [ξ; ξ] (2·n_MC draws) gives the same value and
    gradient as on ξ."""
    prob, data, zP, zMs, ϕ = _setup("f64", 20_000, 8)
    a = _eval(prob, data, ϕ, zP, zMs)
    b = _eval(prob, data, ϕ, np.vstack([zP, zP]), np.vstack([zMs, zMs]))
    assert abs(a[0] - b[0]) <= 1e-11 * abs(a[0])
    np.testing.assert_allclose(b[2], a[2], rtol=1e-9, atol=1e-10 * np.abs(a[2]).max())


def test_site_permutation_invariance():
    prob, data, zP, zMs, ϕ = _setup("f64", 10_000, 8)
    perm = np.random.default_rng(3).permutation(10_000)
    d2 = {k: np.asfortranarray(data[k][:, perm]) for k in ("xM", "xP", "y_o", "y_unc")}
    cols = (perm[:, None] * prob.n_M + np.arange(prob.n_M)[None, :]).reshape(-1)
    a = _eval(prob, data, ϕ, zP, zMs)
    b = _eval(prob, d2, ϕ, zP, np.asfortranarray(zMs[:, cols]))
    assert abs(a[0] - b[0]) <= 1e-11 * abs(a[0])
    np.testing.assert_allclose(b[2], a[2], rtol=1e-8, atol=1e-9 * np.abs(a[2]).max())


@pytest.mark.parametrize("scen", [(), ("covarK2",)])
def test_directional_derivative_at_scale(scen):
    """Float64, 2·10^5 sites: (f(ϕ + h v) − f(ϕ − h v)) / 2h = ∇f·v."""
    prob, data, zP, zMs, ϕ = _setup("f64", 200_000 if not scen else 20_000, 8, scen)
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    nL, comps, g = plan.neg_elbo_withgradient(ϕ, zP, zMs)
    rng = np.random.default_rng(5)
    for _ in range(3):
        v = rng.standard_normal(prob.n_phi)
        v /= np.linalg.norm(v)
        h = 1e-6
        fp = plan.neg_elbo_withgradient(ϕ + h * v, zP, zMs, want_grad=False)[0]
        fm = plan.neg_elbo_withgradient(ϕ - h * v, zP, zMs, want_grad=False)[0]
        fd = (fp - fm) / (2 * h)
        assert abs(fd - g @ v) <= 1e-5 * abs(g @ v) + 1e-7 * np.linalg.norm(g), (fd, g @ v)
