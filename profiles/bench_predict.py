"""predict_hvi / sample_posterior forward path (hvi_predict, src/elbo.jl:320-458) with device-resident outputs:
an output-bandwidth-bound kernel (40 B written per site·sample in Float32: 8 predictions + 2 site parameters).
Usage: python profiles/bench_predict.py [n_site] [n_sample]      (default 1 000 000 x 200 = 8 GB of output)"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import hvi_b200  # noqa: E402

A = hvi_b200._abi
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 200
prob = hvi_b200.DoubleMMCase(dtype="f32")
data = hvi_b200.gen_hybridproblem_synthetic(prob, nb)
plan = hvi_b200.NegElboPlan(prob, device=0)
plan.set_data(data["xM"], data["xP"])            # prediction-only batch: no observations
dev = torch.device("cuda", 0)
ϕ = torch.from_numpy(prob.init_ϕ(perturbed=True)).to(dev)
θP = torch.empty(prob.n_P * M, device=dev)
θMs = torch.empty(nb * prob.n_M * M, device=dev)
y = torch.empty(prob.n_obs * nb * M, device=dev)
ent = C.c_double()
p = lambda t: C.c_void_p(t.data_ptr())


def run(seed):
    rc = plan.lib.hvi_predict(plan._h, M, p(ϕ), None, None, seed, 0, A.HVI_MEM_DEVICE, p(θP), p(θMs), p(y), C.byref(ent))
    assert rc == 0, plan.lib.hvi_last_error(plan._h)


for i in range(3):
    run(i)
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for i in range(n):
    run(10 + i)            # synchronous at return
dt = (time.perf_counter() - t0) / n
out_bytes = 4 * (prob.n_obs + prob.n_M) * nb * M + 4 * prob.n_P * M
in_bytes = 4 * (2 * prob.n_obs + prob.n_covar + prob.n_M) * nb
peak = 6552.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
gbs = (out_bytes + in_bytes) / dt / 1e9
print(f"hvi_predict {nb} sites x {M} samples: {dt * 1e3:.3f} ms per call (prologue + g + k_predict, device draws), "
      f"{nb * M / dt:.3e} site·samples/s, {gbs:.0f} GB/s algorithmic = {gbs / peak:.3f} of the measured {peak:.0f} GB/s; "
      f"y[0:3] = {y[:3].tolist()}")
