# ref_cpu.jl -- times the REAL reference CPU path (Zygote) for BASELINE.md row R0.
# Not runnable in this repository's image (no Julia; never executed).  On a Julia host:
#     julia -t auto --project=<reference checkout> bench_ref/ref_cpu.jl
# prints site·MC-sample neg-ELBO+gradient evaluations per second at the batch sizes the reference can hold (its
# block-diagonal construction, src/elbo.jl:721-743, limits n_batch), with the thread and core counts the north star
# asks for.  The cost function is built with the reference's own `get_loss_elbo` exactly as `solve` does
# (src/HybridSolver.jl:193-258), and differentiated with Zygote like its AutoZygote objective.
using HybridVariationalInference, BenchmarkTools, Zygote, StableRNGs, LinearAlgebra, MLUtils
import HybridVariationalInference as HVI
import ComponentArrays as CA

"the `loss_elbo` closure, ϕ0 and one minibatch of `n_batch` sites, assembled like src/HybridSolver.jl:193-245"
function setup_loss_elbo(prob0, n_MC; scenario=Val(()), n_batch=nothing)
    prob = HybridProblem(prob0; scenario)
    pt = get_hybridproblem_par_templates(prob; scenario)
    cor_ends = get_hybridproblem_cor_ends(prob; scenario)
    g, ϕg0 = get_hybridproblem_MLapplicator(prob; scenario)
    (; transP, transM) = get_hybridproblem_transforms(prob; scenario)
    pbm_covars = get_hybridproblem_pbmpar_covars(prob; scenario)
    n_site, n_batch0 = get_hybridproblem_n_site_and_batch(prob; scenario)
    nb = isnothing(n_batch) ? n_batch0 : n_batch
    (; ϕ, interpreters) = HVI.init_hybrid_params(ϕg0, get_hybridproblem_ϕq(prob; scenario))
    priors = get_hybridproblem_priors(prob; scenario)
    priorsP = Tuple(priors[k] for k in keys(pt.θP)); priorsM = Tuple(priors[k] for k in keys(pt.θM))
    f = HVI.create_nsite_applicator(get_hybridproblem_PBmodel(prob; scenario), nb)
    py = get_hybridproblem_neg_logden_obs(prob; scenario)
    loader = MLUtils.DataLoader(get_hybridproblem_train_dataloader(prob; scenario).data; batchsize=nb, partial=false)
    loss_elbo = HVI.get_loss_elbo(g, transP, HVI.StackedArray(transM, nb), f, py;
        n_MC, n_MC_cap=n_MC, cor_ends, priors_θP_mean=[], priors_θMs_mean=[], cdev=identity, pbm_covars, pt.θP,
        int_ϕq=interpreters.ϕq, int_ϕg_ϕq=interpreters.ϕg_ϕq, priorsP, priorsM, is_omit_priors=Val(false),
        zero_prior_logdensity=HVI.get_zero_prior_logdensity(priorsP, priorsM, pt.θP, pt.θM), approx=prob.approx)
    (; loss_elbo, ϕ0=Vector(CA.getdata(ϕ)), batch=first(loader))
end

println("julia ", VERSION, " threads=", Threads.nthreads(), " cpu_threads=", Sys.CPU_THREADS,
        " blas_threads=", BLAS.get_num_threads())
case = HVI.DoubleMM.DoubleMMCase()
for (n_batch, n_MC) in ((20, 3), (20, 12), (20, 32), (200, 32), (800, 32))
    rng = StableRNG(111)
    (; loss_elbo, ϕ0, batch) = setup_loss_elbo(case, n_MC; n_batch)
    fobj = ϕ -> first(loss_elbo(ϕ, rng, batch...; is_testmode=false))
    Zygote.withgradient(fobj, ϕ0)     # compile
    t = @belapsed Zygote.withgradient($fobj, $ϕ0)
    println("n_batch=", n_batch, " n_MC=", n_MC, "  t=", t, " s  evals/s=", n_batch * n_MC / t)
end
