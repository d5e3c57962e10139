"""Latency of one neg-ELBO + gradient evaluation at the batch sizes the reference itself runs
(n_batch = 20 of 800 sites, n_MC = 3 … 12; test/test_HybridProblem.jl:252, src/HybridSolver.jl:147):
host ϕ in, (nL, ∇ϕ) out, draws on the device.  Usage: python profiles/bench_small_batch.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hvi_b200

for dtype in ("f32", "f64"):
    for nb, M in ((20, 3), (20, 12), (800, 12), (10_000, 32), (100_000, 32)):
        prob = hvi_b200.DoubleMMCase(dtype=dtype)
        data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
        ϕ = prob.init_ϕ(perturbed=True)
        plan = hvi_b200.NegElboPlan(prob)
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        for _ in range(20):
            plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=1)
        n = 200
        t0 = time.perf_counter()
        for i in range(n):
            plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=i)
        dt = (time.perf_counter() - t0) / n
        print(f"{dtype} n_batch={nb:6d} n_MC={M:3d}: {dt * 1e6:8.1f} us per evaluation  ({nb * M / dt:.3e} units/s)", flush=True)

# the same iteration with the optimiser on the device (hvi_adam_run): no ϕ upload, no gradient download, no host step
print("device-resident Adam iterations (hvi_adam_run, 500 iterations, one synchronisation at the end):")
for dtype in ("f32", "f64"):
    for nb, M in ((20, 3), (800, 12), (10_000, 32), (100_000, 32)):
        prob = hvi_b200.DoubleMMCase(dtype=dtype)
        data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
        plan = hvi_b200.NegElboPlan(prob)
        plan.set_data(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        plan.adam_init(prob.init_ϕ(), eta=0.02)
        plan.adam_run(50, M, seed0=0)
        n = 500
        t0 = time.perf_counter()
        tr = plan.adam_run(n, M, seed0=100)
        dt = (time.perf_counter() - t0) / n
        print(f"{dtype} n_batch={nb:6d} n_MC={M:3d}: {dt * 1e6:8.1f} us per iteration  ({nb * M / dt:.3e} units/s)  "
              f"loss {tr[0]:.4g} -> {tr[-1]:.4g}", flush=True)

# the reference's training loop itself: 800 sites on the device, 20-site minibatches walked by hvi_adam_run_minibatches
# (src/HybridSolver.jl:259-260); graph replay on / off
print("minibatch loop on the device-resident dataset (hvi_adam_run_minibatches, 800 sites, 2000 iterations):")
for dtype in ("f32", "f64"):
    # fused: 2 = the single launch forced, 0 = the multi-kernel path, 1 = the library's choice (single launch up to 32 sites)
    for bs, M, graphs, fused in ((20, 3, True, 1), (20, 3, False, 1), (20, 3, True, 0), (20, 12, True, 1), (20, 12, True, 0),
                                 (32, 12, True, 2), (32, 12, True, 0), (50, 12, True, 2), (50, 12, True, 0), (200, 12, True, 2),
                                 (200, 12, True, 0), (200, 12, True, 1)):
        if True:
            prob = hvi_b200.DoubleMMCase(dtype=dtype)
            data = hvi_b200.gen_hybridproblem_synthetic(prob, 800)
            plan = hvi_b200.NegElboPlan(prob)
            plan.set_dataset(data["xM"], data["xP"], data["y_o"], data["y_unc"])
            plan.use_fused(fused)
            plan.use_graphs(graphs)
            plan.adam_init(prob.init_ϕ(), eta=0.02)
            plan.adam_run_minibatches(50, M, bs, seed0=0)
            n = 2000
            t0 = time.perf_counter()
            tr = plan.adam_run_minibatches(n, M, bs, seed0=100)
            dt = (time.perf_counter() - t0) / n
            print(f"{dtype} n_batch={bs:4d} n_MC={M:3d} graphs={graphs!s:5} fused={fused}: {dt * 1e6:8.1f} us per iteration  "
                  f"loss {np.nanmean(tr[:40]):.4g} -> {np.nanmean(tr[-40:]):.4g}", flush=True)

# where the single launch spends its time: SM clocks between the phase boundaries of k_small_iter (thread 0)
print("phase clocks of the single-launch iteration (SM cycles; 1965 MHz):")
for dtype in ("f32", "f64"):
    for bs, M in ((20, 3), (200, 12)):
        prob = hvi_b200.DoubleMMCase(dtype=dtype)
        data = hvi_b200.gen_hybridproblem_synthetic(prob, 800)
        plan = hvi_b200.NegElboPlan(prob)
        plan.set_dataset(data["xM"], data["xP"], data["y_o"], data["y_unc"])
        plan.use_fused(2)
        plan.adam_init(prob.init_ϕ(), eta=0.02)
        plan.fused_phase_clocks(start_only=True)
        plan.adam_run_minibatches(50, M, bs, seed0=0)
        print(f"{dtype} n_batch={bs} n_MC={M}: {plan.fused_phase_clocks()}", flush=True)
