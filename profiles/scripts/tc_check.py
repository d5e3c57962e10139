"""tcgen05 small applicator (hvi_mlp_tc.cu) against the mma.sync kernels and the Float64 CUDA-core path, same inputs.
Usage: python profiles/scripts/tc_check.py [sites] [n_MC]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hvi_b200


def run(prob, data, ϕ, zP, zMs, mode):
    os.environ["HVI_MLP_TC"] = str(mode)
    ft = prob.np_dtype
    plan = hvi_b200.NegElboPlan(prob, device=0)
    plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
    out = plan.neg_elbo_withgradient(ϕ.astype(ft), zP.astype(ft), zMs.astype(ft))
    t0 = time.perf_counter()
    for _ in range(3):
        out = plan.neg_elbo_withgradient(ϕ.astype(ft), zP.astype(ft), zMs.astype(ft))
    dt = (time.perf_counter() - t0) / 3
    plan.close() if hasattr(plan, "close") else None
    return out, dt


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for scen in ((), ("covarK2",)):
        p32 = hvi_b200.DoubleMMCase(scen, dtype="f32")
        p64 = hvi_b200.DoubleMMCase(scen, dtype="f64")
        d32 = hvi_b200.gen_hybridproblem_synthetic(p32, nb, seed=111, driver_nan_frac=0.1)
        data = {k: np.asfortranarray(d32[k].astype(np.float64)) for k in ("xM", "xP", "y_o", "y_unc")}
        zP, zMs = hvi_b200.draw_ξ(p32, nb, M)
        ϕ = p64.init_ϕ(perturbed=True).astype(np.float32).astype(np.float64)
        ref, _ = run(p64, data, ϕ, zP, zMs, 0)
        for mode in (0, 1, 3):
            got, dt = run(p32, d32, ϕ, zP, zMs, mode)
            line = f"scen={scen} nb={nb} M={M} HVI_MLP_TC={mode}: nL rel {abs(got[0] - ref[0]) / abs(ref[0]):.2e}"
            for blk, r in p32.positions().items():
                if len(r):
                    err = np.max(np.abs(np.asarray(got[2][r], dtype=np.float64) - ref[2][r])) / np.max(np.abs(ref[2][r]))
                    line += f" {blk} {err:.1e}"
            print(line + f"  [{dt * 1e3:.2f} ms/call]", flush=True)


if __name__ == "__main__":
    main()
