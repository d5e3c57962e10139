"""The real-reference pin of the whole path, one command away: `julia --project=<reference> julia/dump_golden.jl
tests/golden/julia` writes value, six components and Zygote gradient of the reference's OWN neg_elbo_gtf
(src/elbo.jl:30-88) for the DoubleMM scenarios below; these tests then compare (CPU) the oracle and the C restatement
and (GPU) the CUDA path with those files.  No Julia in this image: the tests skip when no `.bin` is present -- until
then parity stays "unpinned" for the whole path.  The reader/writer round trip is tested regardless.

Tolerance: the reference computes in eltype(ϕ): for Float32 files its own rounding (~1e-5 of the gradient scale) bounds
what can be asserted, so 1e-4 is used for everything Float32 and 1e-9 / 1e-10 for Float64 files."""
import glob
import os

import numpy as np
import pytest

from helpers import O, c_oracle_value_and_grad, hvi_b200, oracle_spec
from golden.from_julia_bin import load_bin, write_bin

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "julia", "*.bin")))
# file name (julia/dump_golden.jl) -> scenario of DoubleMMCase
SCENARIOS = {
    "doublemm_default": (), "doublemm_default_init": (), "doublemm_nmc12": (), "doublemm_covark2": ("covarK2",),
    "doublemm_omit_r0": ("omit_r0",), "doublemm_k1global": ("omit_r0", "K1global"), "doublemm_stackedms": ("stackedMS",),
    "doublemm_neglect_cor": ("neglect_cor",), "doublemm_drivernan": (), "doublemm_sepvar": ("sepvar", "few_sites"),
}


def _problem(path, G, dtype=None):
    name = os.path.splitext(os.path.basename(path))[0]
    prob = hvi_b200.DoubleMMCase(SCENARIOS[name], dtype=dtype or G["dtype"])
    assert prob.n_phi == G["phi"].size, "ϕ layout / applicator size differs from the reference's"
    return prob


def _check(prob, got, G, tol):
    nL, comps, g = got
    assert abs(nL - G["nL"]) <= tol * abs(G["nL"]), (nL, G["nL"])
    assert np.max(np.abs(np.asarray(comps) - G["comps"])) <= tol * np.max(np.abs(G["comps"]))
    for blk, r in prob.positions().items():
        if len(r):
            assert np.max(np.abs(np.asarray(g[r], dtype=np.float64) - G["grad"][r])) <= tol * np.max(np.abs(G["grad"][r])), blk


def test_bin_round_trip(tmp_path):
    """the reader against a file written in the documented layout from a committed golden .npz"""
    G = dict(np.load(os.path.join(HERE, "golden", "doublemm_default.npz")))
    p = str(tmp_path / "doublemm_default.bin")
    write_bin(p, G)
    B = load_bin(p)
    for k in ("phi", "xM", "xP", "y_o", "y_unc", "zP", "zMs", "comps", "grad"):
        assert np.array_equal(B[k], G[k]) and B[k].shape == G[k].shape, k
    assert B["nL"] == float(G["nL"]) and B["dtype"] == "f64" and np.array_equal(B["i_sites0"], np.arange(20))
    with open(p, "ab") as f:
        f.write(b"\0" * 8)
    with pytest.raises(ValueError):
        load_bin(p)


@pytest.mark.skipif(not FILES, reason="no tests/golden/julia/*.bin (run julia/dump_golden.jl on a Julia host)")
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracles_match_the_reference(path):
    G = load_bin(path)
    prob = _problem(path, G, "f64")
    tol = 1e-4 if G["dtype"] == "f32" else 1e-9
    isites = G["i_sites0"] if prob.is_sepvar else None
    got = O.value_and_grad(oracle_spec(prob), G["phi"], G["xM"], G["xP"], G["y_o"], G["y_unc"], G["zP"], G["zMs"], "f64",
                           **({"i_sites": isites} if isites is not None else {}))
    _check(prob, got, G, tol)
    got = c_oracle_value_and_grad(prob, G["phi"], G, G["zP"], G["zMs"], i_sites=isites)
    _check(prob, got, G, tol)


@pytest.mark.gpu
@pytest.mark.skipif(not FILES, reason="no tests/golden/julia/*.bin (run julia/dump_golden.jl on a Julia host)")
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_matches_the_reference(path):
    G = load_bin(path)
    prob = _problem(path, G)
    ft = prob.np_dtype
    tol = 1e-4 if G["dtype"] == "f32" else 1e-10
    plan = hvi_b200.NegElboPlan(prob)
    plan.set_data(*(G[k].astype(ft) for k in ("xM", "xP", "y_o", "y_unc")), i_sites=G["i_sites0"] if prob.is_sepvar else None)
    got = plan.neg_elbo_withgradient(G["phi"].astype(ft), G["zP"].astype(ft), G["zMs"].astype(ft))
    _check(prob, got, G, tol)
