#!/usr/bin/env python
"""bench.py -- site·MC-sample neg-ELBO+gradient evaluations per second (BASELINE.json metric).

One "step" = one full evaluation (value and ∇ϕ) of the DoubleMM negative ELBO over the batch
registered on each GPU.  Workload (config.workload): BASELINE config 4's per-evaluation shape,
DoubleMM default scenario, Float32, n_MC = 64, `--sites` sites PER GPU (weak scaling; default
10^7/8 · 8 = 1.25e6·8 → see --sites), draws ξ resident in HBM for `value`.

  value   : device-timed (CUDA events on the launching stream, max over ranks), inputs resident
            in HBM (site records, xM, ξ).  Inputs are ≫ the 126 MB L2, so no flush is needed.
  e2e     : the same evaluation through the public host API (`NegElboPlan.set_data` +
            `neg_elbo_withgradient`) with the batch (xM, xP, y_o, y_unc) and ϕ in pinned HOST
            memory every step, H2D copies and the D2H read of (nL, ∇ϕ) inside the timed region;
            ξ is drawn on the device from a seed (as the reference's CUDA.randn override does).
  roofline: the fused streaming kernel timed alone with CUDA events inside the library
            (hvi_plan_set_profiling), algorithmic bytes = nb·(4·(2·nO + nO + nO) + 2·4·nM + 4·nM·M)
            (site records + μ_ζ/μ̄_ζ hand-off + ξ) ÷ its duration ÷ MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: oracle/_ref/liboracle.so (C restatement, "port") on a bounded sample.

`--impl reference` times the CPU restatement with all host threads (the Julia reference cannot
run here: no julia binary) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "site·MC-sample neg-ELBO+grad evals/sec (DoubleMM)"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--sites", type=int, default=10_000_000, help="sites per GPU")
    ap.add_argument("--n-mc", type=int, default=64)
    ap.add_argument("--scenario", default="", help="comma-separated DoubleMMCase scenario flags")
    ap.add_argument("--hidden", default="", help="comma-separated hidden widths of g (BASELINE config 5: 512,512,512,512)")
    ap.add_argument("--nan-frac", type=float, default=0.0, help="fraction of sites with missing drivers (:driverNAN generalised)")
    ap.add_argument("--cpu-sample-sites", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --sites per GPU (default); strong: --sites in total, split over the GPUs (BASELINE config 4)")
    ap.add_argument("--p2-sample-sites", type=int, default=100_000, help="sites of the torch-autograd (P2) CPU timing")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def sites_of_rank(args, rank, world):
    """weak scaling: --sites on every GPU; strong scaling: --sites in total, contiguous balanced site ranges"""
    if args.scaling == "weak":
        return args.sites
    base, rem = divmod(args.sites, world)
    return base + (1 if rank < rem else 0)


def workload_config(args, prob):
    per = "sites/GPU" if args.scaling == "weak" else f"sites in total over {args.gpus} GPU(s)"
    return {"workload": f"DoubleMM (BASELINE config 4 shape) {args.sites:.0e} {per} x n_MC={args.n_mc}, "
                        f"scenario=({args.scenario}), MLP {'-'.join(str(l[0]) for l in prob.layers)}-{prob.layers[-1][1]}",
            "sites_per_gpu": args.sites if args.scaling == "weak" else args.sites / max(args.gpus, 1),
            "n_MC": args.n_mc, "n_phi": prob.n_phi, "nan_frac": args.nan_frac,
            "parallelism": f"site-sharded x{args.gpus}, allreduce of [grad; components]",
            "l2": "inputs (>= 600 B/site x sites) exceed the 126 MB L2; no flush needed"}


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_baseline(prob, n_MC, sample_sites, nthreads):
    """Time the C restatement (oracle) on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hvi_b200
    from helpers import c_oracle_value_and_grad
    data = hvi_b200.gen_hybridproblem_synthetic(prob, sample_sites, seed=111)
    zP, zMs = hvi_b200.draw_ξ(prob, sample_sites, n_MC)
    ϕ = prob.init_ϕ(perturbed=True)
    small = {k: np.ascontiguousarray(v[..., :1000]) for k, v in data.items() if hasattr(v, "shape") and v.ndim == 2}
    c_oracle_value_and_grad(prob, ϕ, small, zP, zMs[:, :1000 * prob.n_M], nthreads=nthreads)  # warm (library load)
    t0 = time.perf_counter()
    c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return sample_sites * n_MC / dt, dt


def p2_baseline(prob, n_MC, sample_sites):
    """P2 of BASELINE.md: the float32 torch-autograd oracle (taped reverse mode, structurally the closest thing to the
    reference's Zygote path that can run here) on a bounded sample, all torch CPU threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import hvi_b200
    from helpers import O, oracle_spec
    data = hvi_b200.gen_hybridproblem_synthetic(prob, sample_sites, seed=111)
    zP, zMs = hvi_b200.draw_ξ(prob, sample_sites, n_MC)
    ϕ = prob.init_ϕ(perturbed=True)
    spec = oracle_spec(prob)
    n0 = min(1000, sample_sites)
    O.value_and_grad(spec, ϕ, data["xM"][:, :n0], data["xP"][:, :n0], data["y_o"][:, :n0], data["y_unc"][:, :n0], zP,
                     zMs[:, :n0 * prob.n_M], "f32")
    t0 = time.perf_counter()
    O.value_and_grad(spec, ϕ, data["xM"], data["xP"], data["y_o"], data["y_unc"], zP, zMs, "f32")
    dt = time.perf_counter() - t0
    return {"value": sample_sites * n_MC / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_sites} sites x n_MC={n_MC}, one evaluation, {dt:.1f} s",
            "note": "P2: oracle/hvi_oracle.py, float32 torch autograd on the CPU (Zygote-like proxy)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot run here (Julia absent), so
    the CPU restatement of it (oracle port) is timed with all host threads on the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import hvi_b200
    prob = hvi_b200.DoubleMMCase(tuple(s for s in args.scenario.split(",") if s), dtype="f32")
    cores = os.cpu_count() or 1
    sample = args.cpu_sample_sites
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import c_oracle_value_and_grad
    data = hvi_b200.gen_hybridproblem_synthetic(prob, sample, seed=111)
    zP, zMs = hvi_b200.draw_ξ(prob, sample, args.n_mc)
    ϕ = prob.init_ϕ(perturbed=True)
    for _ in range(args.warmup):
        c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle_value_and_grad(prob, ϕ, data, zP, zMs, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample * args.n_mc / dt
    sample_txt = f"{sample} sites x n_MC={args.n_mc} per step (bounded sample of the {args.sites:.0e}-site workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, prob),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt,
                         "note": "C restatement of the reference (oracle/hvi_oracle.c, OpenMP); the Julia reference "
                                 "itself cannot run in this image"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def dominant_roofline(args, prob, nb, M, kt, peaks, sampler, local):
    """Roofline of the kernel (group) that dominates THIS run, from the library's own per-kernel CUDA events:
    the fused streaming kernel against HBM, or the applicator g (forward + backward) against the tensor pipe."""
    nO, nM = prob.n_obs, prob.n_M
    g_ms = kt["mlp_fwd"] + kt["mlp_bwd"]
    mac = sum(i * o for i, o, _, _ in prob.layers)
    cols = nb * (M + 1) if prob.pbm_covars else nb
    wide = max(max(i, o) for i, o, _, _ in prob.layers) > 64
    traffic = None
    # wide applicator with pbm covariates: the evaluation walks blocks of sites (pass-2 forward -> stream -> pass-2 backward
    # per block) and the library reports the whole loop in the "stream" slot; the streaming kernel itself is
    # below 1 % of it, so the loop counts as applicator time
    blocked = bool(wide and prob.pbm_covars)
    if blocked:
        g_ms += kt["stream"]
    if kt["stream"] >= g_ms and not blocked:
        # site records + μ_ζ/μ̄_ζ hand-off + ξ (+ per-draw means and their cotangents with pbm covariates)
        alg_bytes = nb * (4 * 4 * nO + 2 * 4 * nM + 4 * nM * M + (2 * 4 * nM * M if prob.pbm_covars else 0))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / (kt["stream"] * 1e-3) / 1e9
        kernel = "k_stream"
        # DRAM bytes per launch from the committed `ncu --set full` capture of this workload and this kernel variant
        # (profiles/traffic.json, written from the capture by profiles/summarize_ncu.py); null without a capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(f"k_stream:{args.scenario}:{nb}x{M}", {})
            traffic = ent.get("dram_bytes")
            kernel = ent.get("kernel", kernel)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kt}
        # the kernel's real limiter is the FP32 pipe: pipe slots per site·draw counted in the SASS of the
        # hot loop (profiles/sass_summary.txt) against 128 lanes x SMs x clock
        try:
            import torch
            sm = torch.cuda.get_device_properties(local).multi_processor_count
            clk_hz = float(sampler.summary().get("sm_mhz") or 1965.0) * 1e6
            if nO == 8 and nM == 2 and not prob.pbm_covars:
                slots = FP32_SLOTS_PER_UNIT
                roofline["fp32_pipe"] = {"slots_per_unit": slots, "peak_lane_slots_per_s": sm * 128 * clk_hz,
                                         "frac": slots * nb * M / (kt["stream"] * 1e-3) / (sm * 128 * clk_hz)}
        except Exception:
            pass
        return roofline
    # the applicator dominates: 2·Σ in·out flop per column forward, x3 with the backward pass; the tensor
    # work is 3 passes of that (3xBF16 / 3xTF32 error-compensated split) -- the algorithmic figure is the fp32-equivalent one
    alg_flop = 3.0 * 2.0 * mac * cols
    peak = float(peaks.get("bf16_tflops_sustained", 1384.0))
    achieved = alg_flop / (g_ms * 1e-3) / 1e12
    tc_small = os.environ.get("HVI_MLP_TC", "2") != "0"
    kernel = ("k_gemm_bf3 chain (tcgen05 kind::f16, TMA)" if wide else
              "k_mlp3_tc_fwd + k_mlp3_tc_bwd (tcgen05, TMEM accumulators)" if tc_small else "k_mlp3_mma_fwd + k_mlp3_mma_bwd (mma.sync)")
    return {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)" if peaks
                           else "fallback 1384 TFLOP/s",
            "algorithmic_flop_per_step": alg_flop, "tensor_passes": 3, "columns": cols, "kernel_ms": kt,
            **({"kernel_ms_note": "site-blocked schedule: 'stream' holds the per-block loop (g forward, k_stream, g backward)"}
               if blocked else {})}


# FP32-pipe slots per site·draw of k_stream's hot loop (DoubleMM, 8 observations): packed instructions count twice
FP32_SLOTS_PER_UNIT = 149


def bind_to_gpu_numa_node(gpu):
    """Run this rank on the CPUs next to its GPU, so that the pinned batch buffers are allocated on that NUMA node and the
    host -> device uploads of the ranks do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(gpu), (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import hvi_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    hidden = tuple(int(w) for w in args.hidden.split(",") if w) or None
    prob = hvi_b200.DoubleMMCase(tuple(s for s in args.scenario.split(",") if s), dtype="f32", hidden=hidden)
    nb, M = sites_of_rank(args, rank, world), args.n_mc
    n_total = world * nb if args.scaling == "weak" else args.sites
    # synthetic batch of this rank's site shard (generator restates gen_hybridproblem_synthetic)
    data = hvi_b200.gen_hybridproblem_synthetic(prob, nb, seed=111 + rank, driver_nan_frac=args.nan_frac)
    ϕ = prob.init_ϕ(perturbed=True)
    plan = hvi_b200.NegElboPlan(prob, device=local, count_globals=(rank == 0))
    pin = {k: torch.from_numpy(np.ascontiguousarray(data[k].T)).pin_memory() for k in ("xM", "xP", "y_o", "y_unc")}
    host = {k: pin[k].numpy().T for k in pin}  # column-major views of the pinned buffers
    plan.set_data(host["xM"], host["xP"], host["y_o"], host["y_unc"])
    n_phi = plan.n_phi
    ϕ_dev = torch.from_numpy(ϕ).to(dev)
    out_dev = torch.zeros(n_phi + 8, dtype=torch.float64, device=dev)
    # ξ resident in HBM (generated once, outside the timed region, by the library's generator via
    # a seeded evaluation that leaves the draws in the plan -- then re-used through explicit buffers)
    g = torch.Generator(device=dev)
    g.manual_seed(211 + rank)
    zP_dev = torch.randn(M * max(prob.n_P, 1), device=dev, generator=g)
    zMs_dev = torch.randn(M * prob.n_M * nb, device=dev, generator=g)
    stream = torch.cuda.current_stream()

    def step():
        plan.neg_elbo_async(ϕ_dev, zP_dev, zMs_dev, out_dev, M, stream.cuda_stream)
        if dist is not None:
            dist.all_reduce(out_dev)  # one all-reduce of [∇ϕ ; components]: SURVEY 8(e)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = plan.launch_count - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_total * M / (ms_step * 1e-3)
    res = out_dev.cpu().numpy()

    # ---- roofline of the dominant kernel, timed alone by the library's own events ----------
    plan.set_profiling(True)
    acc = {}
    for _ in range(max(3, args.steps // 2)):
        plan.neg_elbo_async(ϕ_dev, zP_dev, zMs_dev, out_dev, M, stream.cuda_stream)
        torch.cuda.synchronize()
        for k, v in plan.kernel_times_ms().items():
            acc.setdefault(k, []).append(v)
    plan.set_profiling(False)
    sampler.stop_flag = True
    kt = {k: float(np.mean(v)) for k, v in acc.items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = dominant_roofline(args, prob, nb, M, kt, peaks, sampler, local)
    # ---- end to end through the public host API ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(pin[k].numel() * 4 for k in pin) + n_phi * 4
        d2h = n_phi * 4 + 8 * 8
        n_e2e = max(2, min(args.steps, 5))
        plan.set_data(host["xM"], host["xP"], host["y_o"], host["y_unc"])
        plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=1)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            plan.set_data(host["xM"], host["xP"], host["y_o"], host["y_unc"])     # the step's batch: host → device
            nL, comps, grad = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=2 + i)   # ϕ host → device, (nL, ∇ϕ) → host
            if dist is not None:
                gt = torch.from_numpy(grad).to(dev)
                dist.all_reduce(gt)
                grad = gt.cpu().numpy()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        serial = n_total * M / float(tt.item())
        serial_ms = float(tt.item()) * 1e3
        # the minibatch loop as the library is meant to be driven: the upload of batch i+1 (prefetch_data, copy
        # stream) runs under the evaluation of batch i; every step still moves one whole batch host → device and
        # reads (nL, ∇ϕ) back, all inside the timed region (the final commit waits for the last upload)
        plan.prefetch_data(host["xM"], host["xP"], host["y_o"], host["y_unc"])
        plan.commit_data()
        plan.prefetch_data(host["xM"], host["xP"], host["y_o"], host["y_unc"])
        plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=1)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            plan.commit_data()                                                       # batch i: uploaded during step i-1
            plan.prefetch_data(host["xM"], host["xP"], host["y_o"], host["y_unc"])  # batch i+1: host → device
            nL, comps, grad = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=2 + i)     # ϕ host → device, (nL, ∇ϕ) → host
            if dist is not None:
                gt = torch.from_numpy(grad).to(dev)
                dist.all_reduce(gt)
                grad = gt.cpu().numpy()
        plan.commit_data()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * M / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": n_e2e, "ms_per_step": float(tt.item()) * 1e3,
               "what": "per step: commit_data + prefetch_data(next host batch, pinned) + neg_elbo_withgradient(host ϕ, "
                       "device-drawn ξ); the upload overlaps the evaluation; wall clock",
               "serial_value": serial, "serial_ms_per_step": serial_ms,
               "serial_what": "set_data(host batch) then neg_elbo_withgradient, no overlap"}
        # host batch with the arrays that are the same for every site (DoubleMM drivers, src/DoubleMM/f_doubleMM.jl:12-13,375,
        # and the constant y_unc, :392) passed as ONE column (hvi_plan_set_data_bcast): 52 instead of 148 bytes per site
        try:
            xs, us = np.asfortranarray(host["xP"][:, :1]), np.asfortranarray(host["y_unc"][:, :1])
            if args.nan_frac == 0.0 and np.all(host["xP"][:, :64] == xs) and np.all(host["y_unc"][:, :64] == us):
                plan.set_data_bcast(host["xM"], xs, host["y_o"], us)
                plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=1)
                barrier()
                t0 = time.perf_counter()
                for i in range(n_e2e):
                    plan.set_data_bcast(host["xM"], xs, host["y_o"], us)
                    nL, comps, grad = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=2 + i)
                    if dist is not None:
                        gt = torch.from_numpy(grad).to(dev)
                        dist.all_reduce(gt)
                        grad = gt.cpu().numpy()
                barrier()
                tt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
                if dist is not None:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e["broadcast_value"] = n_total * M / float(tt.item())
                e2e["broadcast_ms_per_step"] = float(tt.item()) * 1e3
                e2e["broadcast_h2d_bytes_per_step"] = int(host["xM"].size * 4 + host["y_o"].size * 4 + xs.size * 4 + us.size * 4 + n_phi * 4)
                e2e["broadcast_what"] = ("set_data_bcast(host xM, ONE shared column of xP, host y_o, ONE shared column of y_unc) "
                                         "then neg_elbo_withgradient, no overlap")
        except Exception as ex:
            e2e["broadcast_what"] = f"skipped: {ex}"
        # data-resident variant (batch registered once, as gdev_hybridproblem_dataloader does)
        t0 = time.perf_counter()
        for i in range(n_e2e):
            plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=2 + i)
        barrier()
        e2e["resident_value"] = n_total * M / ((time.perf_counter() - t0) / n_e2e)
        # the kernels of one such call (device draws: Philox + Box-Muller inside k_stream instead of ξ read from HBM)
        plan.set_profiling(True)
        plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=99)
        e2e["resident_kernel_ms"] = plan.kernel_times_ms()
        plan.set_profiling(False)
        # training set resident on the device, the minibatch chosen there (hvi_plan_set_dataset + select_range --
        # gdev_hybridproblem_dataloader, src/AbstractHybridProblem.jl:220-250): per step one gather kernel instead
        # of an upload; ϕ host -> device and (nL, ∇ϕ) -> host stay inside the timed region
        try:
            plan.set_dataset(host["xM"], host["xP"], host["y_o"], host["y_unc"])
            plan.select_range(0, nb)
            plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=1)
            barrier()
            t0 = time.perf_counter()
            for i in range(n_e2e):
                plan.select_range(0, nb)
                nL, comps, grad = plan.neg_elbo_withgradient(ϕ, n_MC=M, seed=2 + i)
                if dist is not None:
                    gt = torch.from_numpy(grad).to(dev)
                    dist.all_reduce(gt)
                    grad = gt.cpu().numpy()
            barrier()
            tt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e["device_dataset_value"] = n_total * M / float(tt.item())
            e2e["device_dataset_ms_per_step"] = float(tt.item()) * 1e3
            e2e["device_dataset_what"] = ("set_dataset once (outside the timed region); per step select_range (device "
                                          "gather + re-tile) + neg_elbo_withgradient(host ϕ): h2d = 4·n_phi bytes")
        except Exception as ex:   # e.g. not enough HBM for a second copy of a very large batch
            e2e["device_dataset_value"] = None
            e2e["device_dataset_what"] = f"skipped: {ex}"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, prob),
                "clocks": sampler.summary(), "gpu_launches": int(launches), "roofline": roofline,
                "nL": float(res[n_phi + 6]), "nonfinite": float(res[n_phi + 7])}
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            v, dt = cpu_baseline(prob, M, args.cpu_sample_sites, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{args.cpu_sample_sites} sites x n_MC={M}, one evaluation, {dt:.1f} s",
                                    "note": "oracle/hvi_oracle.c (C restatement of the Julia reference), f32, 1 thread"}
            if args.p2_sample_sites > 0 and not prob.pbm_covars and hidden is None:
                try:
                    line["cpu_baseline"]["p2_torch_autograd"] = p2_baseline(prob, M, args.p2_sample_sites)
                except Exception as ex:
                    line["cpu_baseline"]["p2_torch_autograd"] = {"unavailable": str(ex)}
        print(json.dumps(line, ensure_ascii=False))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
