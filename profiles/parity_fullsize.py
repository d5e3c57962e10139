"""Parity at (a shard of) the full BASELINE size against the C restatement of the reference (oracle/hvi_oracle.c,
float64, all host threads): the CUDA path in Float32 and Float64 with the SAME standard-normal draws ξ passed in.
Usage (GPU box): python profiles/parity_fullsize.py [n_site] [n_MC]     (default 2 000 000 x 64)
Prints the relative deviations of the value, of the six components and of every ϕ block."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import hvi_b200  # noqa: E402
from helpers import c_oracle_value_and_grad  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for scen in ((), ("covarK2",)):
    if scen and nb > 200_000:
        nb_s = 200_000      # g runs on every (draw, site) column on the CPU too
    else:
        nb_s = nb
    p64 = hvi_b200.DoubleMMCase(scen, dtype="f64")
    data = hvi_b200.gen_hybridproblem_synthetic(p64, nb_s, seed=111, driver_nan_frac=0.1)
    zP, zMs = hvi_b200.draw_ξ(hvi_b200.DoubleMMCase(scen, dtype="f32"), nb_s, M)   # float32 draws, exact in float64
    ϕ = p64.init_ϕ(perturbed=True).astype(np.float32).astype(np.float64)            # representable in both types
    t0 = time.perf_counter()
    nL_o, comps_o, g_o = c_oracle_value_and_grad(p64, ϕ, data, zP.astype(np.float64), zMs.astype(np.float64),
                                                 nthreads=os.cpu_count() or 1)
    t_o = time.perf_counter() - t0
    for dt in ("f64", "f32"):
        prob = hvi_b200.DoubleMMCase(scen, dtype=dt)
        plan = hvi_b200.NegElboPlan(prob, device=0)
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        nL, comps, g = plan.neg_elbo_withgradient(ϕ.astype(prob.np_dtype), zP.astype(prob.np_dtype), zMs.astype(prob.np_dtype))
        dev = {name: float(np.max(np.abs(g[r] - g_o[r])) / np.max(np.abs(g_o[r]))) for name, r in prob.positions().items() if len(r)}
        print(f"scenario={scen} {nb_s} sites x {M} draws {dt}: |nL-nL_o|/|nL_o| = {abs(nL - nL_o) / abs(nL_o):.2e}, "
              f"components {np.max(np.abs(comps - comps_o)) / np.max(np.abs(comps_o)):.2e}, gradient blocks "
              + ", ".join(f"{k} {v:.1e}" for k, v in dev.items()) + f"   (oracle {t_o:.1f} s)", flush=True)
        del plan
