# dump_golden.jl -- golden input/output vectors of the REAL reference for the neg-ELBO path.
#
# Run on a host that has Julia and HybridVariationalInference.jl (this repository's build image has neither):
#
#     julia --project=<reference checkout> julia/dump_golden.jl tests/golden/julia
#     python -m pytest tests/test_julia_golden.py            # CPU: oracle vs reference;  -m gpu: CUDA vs reference
#
# What is evaluated is the reference's OWN `neg_elbo_gtf_components` (src/elbo.jl:46-88) and its Zygote gradient
# (as in src/HybridSolver.jl:256-258) -- not a re-derivation.  The only intervention is the random number
# generator: `ReplayRNG` hands `_create_randn` (src/elbo.jl:771-773, called at :605-606 first for zP, then for zMs)
# arrays drawn beforehand, so that the very same standard-normal draws ξ can be given to the CUDA path.
#
# File format, <name>.bin = flat little-endian Float64 stream:
#   header (10 values): magic 20260.0, n_phi, n_covar (rows of xM), n_obs, n_batch, n_MC, n_P, n_M,
#                       sizeof(eltype ϕ) (4 | 8), n_site_total
#   then ϕ | xM | xP (2 n_obs x n_batch) | y_o | y_unc | i_sites (one-based) | zP (n_MC x n_P) | zMs (n_MC x n_M n_batch) |
#        nL | nLjoint, entropy_ζ, loss_penalty, nLy, neg_log_prior, neg_log_jac | ∇ϕ            (matrices column-major)
# tests/golden/from_julia_bin.py reads it.
using HybridVariationalInference, ComponentArrays, StableRNGs, Zygote, Random
import HybridVariationalInference as HVI
import ComponentArrays as CA

include(joinpath(@__DIR__, "replay_rng.jl"))

function dump_case(name, prob0; scenario=Val(()), n_MC=3, out_dir=".", perturb=true,
        approx=(:sepvar ∈ HVI._val_value(scenario)) ? HVI.MeanVarSepHVIApproximation() : HVI.MeanHVIApproximationMat())
    prob = HybridProblem(prob0; scenario, approx)
    # the pieces `solve` assembles (src/HybridSolver.jl:193-245)
    pt = get_hybridproblem_par_templates(prob; scenario)
    g, ϕg0 = get_hybridproblem_MLapplicator(prob; scenario)
    ϕq = get_hybridproblem_ϕq(prob; scenario)
    (; ϕ, interpreters) = HVI.init_hybrid_params(ϕg0, ϕq)
    int_ϕq, int_ϕg_ϕq = interpreters.ϕq, interpreters.ϕg_ϕq
    (; transP, transM) = get_hybridproblem_transforms(prob; scenario)
    cor_ends = get_hybridproblem_cor_ends(prob; scenario)
    f = get_hybridproblem_PBmodel(prob; scenario)
    py = get_hybridproblem_neg_logden_obs(prob; scenario)
    priors = get_hybridproblem_priors(prob; scenario)
    n_site, _ = get_hybridproblem_n_site_and_batch(prob; scenario)
    xM, xP, y_o, y_unc, i_sites = first(get_hybridproblem_train_dataloader(prob; scenario))
    xM, xP, y_o, y_unc = Matrix(CA.getdata(xM)), Matrix(CA.getdata(xP)), Matrix(y_o), Matrix(y_unc)
    i_sites = collect(i_sites)
    n_batch = size(xM, 2)
    transMs = HVI.StackedArray(transM, n_batch)
    pbm_covars = get_hybridproblem_pbmpar_covars(prob; scenario)
    pbm_covar_indices = HVI.get_pbm_covar_indices(pt.θP, pbm_covars)
    priorsP = Tuple(priors[k] for k in keys(pt.θP))
    priorsM = Tuple(priors[k] for k in keys(pt.θM))
    FT = eltype(ϕ)
    ϕv = Vector(CA.getdata(ϕ))
    if perturb   # non-zero correlations and σ slopes, as test/test_elbo.jl:133-149 does, so that every ϕ block matters
        pos = HVI.get_positions(int_ϕq)
        off = length(ϕg0)
        isempty(pos.ρsP) || (ϕv[off .+ pos.ρsP] .= FT(1.33))
        isempty(pos.ρsM) || (ϕv[off .+ pos.ρsM] .= FT(-0.75))
        if haskey(pos, :coef_logσ2_ζMs)
            cf = vec(pos.coef_logσ2_ζMs)
            ϕv[off .+ cf] .= FT.(repeat([log(0.1^2), -0.3], length(cf) ÷ 2) .* (1 .+ 0.1 .* (0:length(cf)-1)))
        end
        ϕv[off .+ pos.logσ2_ζP] .= FT(log(0.2^2))
    end
    srng = StableRNG(211)
    zP = randn(srng, FT, n_MC, length(pt.θP))
    zMs = randn(srng, FT, n_MC, length(pt.θM) * n_batch)
    rng = ReplayRNG(zP, zMs)
    kw = (; int_ϕq, int_ϕg_ϕq, n_MC, cor_ends, pbm_covar_indices, transP, transMs, priorsP, priorsM,
        is_testmode=false, is_omit_priors=Val(false), zero_prior_logdensity=zero(FT), approx)
    c = HVI.neg_elbo_gtf_components(rng, ϕv, g, f, py, xM, xP, y_o, y_unc, i_sites; kw...)
    nL = c.nLjoint - c.entropy_ζ + c.loss_penalty
    gr = Zygote.gradient(v -> HVI.neg_elbo_gtf(rng, v, g, f, py, xM, xP, y_o, y_unc, i_sites; kw...), ϕv)[1]
    hdr = Float64[20260, length(ϕv), size(xM, 1), size(y_o, 1), n_batch, n_MC, length(pt.θP), length(pt.θM),
        sizeof(FT), n_site]
    mkpath(out_dir)
    open(joinpath(out_dir, name * ".bin"), "w") do io
        for a in (hdr, ϕv, xM, xP, y_o, y_unc, i_sites, zP, zMs, [nL],
                  [c.nLjoint, c.entropy_ζ, c.loss_penalty, c.nLy, c.neg_log_prior, c.neg_log_jac], gr)
            write(io, htol.(Float64.(vec(collect(a)))))
        end
    end
    println(name, ": nL = ", nL, "  |∇ϕ|∞ = ", maximum(abs, gr))
end

out_dir = length(ARGS) > 0 ? ARGS[1] : "."
case = HVI.DoubleMM.DoubleMMCase()
# names must match PROBLEMS in tests/test_julia_golden.py
dump_case("doublemm_default", case; out_dir)
dump_case("doublemm_default_init", case; out_dir, perturb=false)
dump_case("doublemm_covark2", case; scenario=Val((:covarK2,)), out_dir)
dump_case("doublemm_omit_r0", case; scenario=Val((:omit_r0,)), out_dir)
dump_case("doublemm_k1global", case; scenario=Val((:omit_r0, :K1global)), out_dir)
dump_case("doublemm_stackedms", case; scenario=Val((:stackedMS,)), out_dir)
dump_case("doublemm_neglect_cor", case; scenario=Val((:neglect_cor,)), out_dir)
dump_case("doublemm_drivernan", case; scenario=Val((:driverNAN,)), out_dir)
dump_case("doublemm_sepvar", case; scenario=Val((:sepvar, :few_sites)), out_dir)
dump_case("doublemm_nmc12", case; n_MC=12, out_dir)
