"""A few single-launch Adam iterations at the reference's batch size (20 sites x 3 draws, Float32) for an ncu capture of
k_small_iter.  Usage: ncu ... -k regex:k_small_iter python profiles/scripts/small_iter_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hvi_b200

prob = hvi_b200.DoubleMMCase(dtype="f32")
data = hvi_b200.gen_hybridproblem_synthetic(prob, 800)
plan = hvi_b200.NegElboPlan(prob)
plan.set_dataset(data["xM"], data["xP"], data["y_o"], data["y_unc"])
plan.adam_init(prob.init_ϕ(), eta=0.02)
plan.use_fused(1)
tr = plan.adam_run_minibatches(400, 3, 20, seed0=1)
print(tr[-3:])
