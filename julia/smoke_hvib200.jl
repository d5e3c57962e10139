# smoke_hvib200.jl -- first thing to run on a host with Julia, the reference package and a B200 (never run in this
# repository's image: no Julia).   HVI_B200_LIB=<repo>/hybridvariationalinference.jl_b200/csrc/libhvi_b200.so \
#     julia --project=<reference checkout> julia/smoke_hvib200.jl
# 1. loads the binding (struct-layout assertions), 2. evaluates the reference's own neg_elbo_gtf + Zygote gradient and
# the library on the SAME draws (ReplayRNG, see dump_golden.jl) and prints the relative deviations, 3. runs two epochs
# of the unchanged `solve` on the wrapped problem.
using HybridVariationalInference, ComponentArrays, StableRNGs, Zygote, Random, Optimisers
import HybridVariationalInference as HVI
import ComponentArrays as CA
include(joinpath(@__DIR__, "HVIB200.jl"))
include(joinpath(@__DIR__, "replay_rng.jl"))

scenario = Val(())
prob = HybridProblem(HVI.DoubleMM.DoubleMMCase(); scenario)
pt = get_hybridproblem_par_templates(prob; scenario)
g, ϕg0 = get_hybridproblem_MLapplicator(prob; scenario)
(; ϕ, interpreters) = HVI.init_hybrid_params(ϕg0, get_hybridproblem_ϕq(prob; scenario))
(; transP, transM) = get_hybridproblem_transforms(prob; scenario)
cor_ends = get_hybridproblem_cor_ends(prob; scenario)
f = get_hybridproblem_PBmodel(prob; scenario)
priors = get_hybridproblem_priors(prob; scenario)
priorsP = Tuple(priors[k] for k in keys(pt.θP)); priorsM = Tuple(priors[k] for k in keys(pt.θM))
xM, xP, y_o, y_unc, i_sites = first(get_hybridproblem_train_dataloader(prob; scenario))
xM, xP, y_o, y_unc = Matrix(CA.getdata(xM)), Matrix(CA.getdata(xP)), Matrix(y_o), Matrix(y_unc)
n_batch, n_MC, FT = size(xM, 2), 3, eltype(ϕ)
ϕv = Vector(CA.getdata(ϕ))
srng = StableRNG(211)
zP = randn(srng, FT, n_MC, length(pt.θP)); zMs = randn(srng, FT, n_MC, length(pt.θM) * n_batch)
kw = (; int_ϕq=interpreters.ϕq, int_ϕg_ϕq=interpreters.ϕg_ϕq, n_MC, cor_ends,
    pbm_covar_indices=HVI.get_pbm_covar_indices(pt.θP, ()), transP, transMs=HVI.StackedArray(transM, n_batch),
    priorsP, priorsM, is_testmode=false, is_omit_priors=Val(false), zero_prior_logdensity=zero(FT), approx=prob.approx)
rng = ReplayRNG(zP, zMs)
nL_ref = HVI.neg_elbo_gtf(rng, ϕv, g, f, neg_logden_indep_normal, xM, xP, y_o, y_unc, collect(i_sites); kw...)
g_ref = Zygote.gradient(v -> HVI.neg_elbo_gtf(rng, v, g, f, neg_logden_indep_normal, xM, xP, y_o, y_unc, collect(i_sites); kw...), ϕv)[1]

plan = HVIB200.Plan(HVIB200.make_desc(prob; scenario))
HVIB200.check_layout(plan, interpreters.ϕg_ϕq, interpreters.ϕq)
HVIB200.set_data!(plan, xM, xP, y_o, y_unc)
nL, comps, gr = HVIB200.neg_elbo_withgradient(plan, ϕv, zP, zMs)
println("nL reference ", nL_ref, "  library ", nL, "  rel ", abs(nL - nL_ref) / abs(nL_ref))
println("gradient rel deviation ", maximum(abs, gr .- g_ref) / maximum(abs, g_ref), "   (Float32: both sides carry ~1e-5)")

probb = HVIB200.b200_problem(prob; scenario)
res = solve(probb, HybridPosteriorSolver(; alg=Optimisers.Adam(0.02), n_MC=3); scenario, rng=StableRNG(111), epochs=2)
println("solve on the wrapped problem: final objective ", res.resopt.objective)
