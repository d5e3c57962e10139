# ref_cpu.jl -- times the REAL reference CPU path (Zygote) for BASELINE.md row R0.
# Not runnable in this repository's image (no Julia).  On a Julia host:
#   julia -t auto --project=<reference> bench_ref/ref_cpu.jl
# prints site·MC-sample neg-ELBO+grad evaluations per second for the batch sizes the reference can
# hold (its block-diagonal construction, src/elbo.jl:721-743, limits n_batch), with the thread
# and core counts the north star asks for.
using HybridVariationalInference, BenchmarkTools, Zygote, StableRNGs, LinearAlgebra
import HybridVariationalInference as HVI

prob = HVI.DoubleMM.DoubleMMCase()
println("julia ", VERSION, " threads=", Threads.nthreads(), " cpu_threads=", Sys.CPU_THREADS,
        " blas_threads=", BLAS.get_num_threads())
for (scenario, n_MC) in ((Val(()), 3), (Val(()), 32), (Val((:few_sites,)), 32))
    solver = HybridPosteriorSolver(; alg=nothing, n_MC)
    rng = StableRNG(111)
    # build loss_elbo exactly as solve does (src/HybridSolver.jl:240-258)
    (; loss_elbo, ϕ0, batch) = HVI.setup_loss_elbo(prob, solver; scenario, rng)   # helper: see INTEGRATION.md
    t = @belapsed Zygote.withgradient(ϕ -> first($loss_elbo(ϕ, $rng, $batch...; is_testmode=false)), $ϕ0)
    n_batch = size(batch[1], 2)
    println("scenario=", scenario, " n_batch=", n_batch, " n_MC=", n_MC, "  t=", t, " s  evals/s=", n_batch * n_MC / t)
end
