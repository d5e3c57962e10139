"""Times g forward of the BASELINE config-5 network (5-512-512-512-512-2) through hvi_apply_g:
hidden layers on tcgen05 (3xTF32).  Usage: python profiles/bench_wide_g.py [n_site]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hvi_b200

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
prob = hvi_b200.DoubleMMCase(dtype="f32", hidden=(512, 512, 512, 512))
data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
ϕ = prob.init_ϕ()
plan = hvi_b200.NegElboPlan(prob)
plan.set_data(data["xM"], data["xP"])
plan.apply_g(ϕ)
t0 = time.perf_counter()
n = 3
for _ in range(n):
    mu = plan.apply_g(ϕ)
dt = (time.perf_counter() - t0) / n
flop = 2.0 * nb * (5 * 512 + 3 * 512 * 512 + 512 * 2)
print(f"n_site={nb} apply_g {dt * 1e3:.2f} ms  {flop / dt / 1e12:.1f} TFLOP/s fp32-equivalent "
      f"({3 * 2.0 * nb * 3 * 512 * 512 / dt / 1e12:.1f} TFLOP/s of tf32 tensor work incl. the 3x split)")
