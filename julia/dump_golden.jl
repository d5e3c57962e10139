# dump_golden.jl -- writes golden input/output vectors of the REAL reference for the neg-ELBO path.
#
# Run on a host that has Julia and HybridVariationalInference.jl (this repository's build image has
# neither):   julia --project=<reference> julia/dump_golden.jl out_dir
# For each case it writes <name>.bin = a flat little-endian Float64 stream
#   [n_phi, n_covar, n_obs, n_site, n_MC, n_P, n_M | ϕ | xM | xP | y_o | y_unc | zP | zMs | nL | 6 comps | ∇ϕ]
# (matrices column-major), the same quantities tests/golden/*.npz hold.  tests/golden/make_golden.py
# documents the matching .npz layout; converting is `np.fromfile(..., '<f8')` plus the header.
using HybridVariationalInference, ComponentArrays, StableRNGs, Zygote, Random
import HybridVariationalInference as HVI
import ComponentArrays as CA

function dump_case(name, prob; scenario=Val(()), n_MC=3, out_dir=".")
    solver = HybridPosteriorSolver(; alg=nothing, n_MC)
    # the pieces `solve` assembles (src/HybridSolver.jl:193-245)
    pt = get_hybridproblem_par_templates(prob; scenario)
    g, ϕg0 = get_hybridproblem_MLapplicator(prob; scenario)
    ϕq = get_hybridproblem_ϕq(prob; scenario)
    (; ϕ, interpreters) = HVI.init_hybrid_params(ϕg0, ϕq)
    (; transP, transM) = get_hybridproblem_transforms(prob; scenario)
    cor_ends = get_hybridproblem_cor_ends(prob; scenario)
    f = get_hybridproblem_PBmodel(prob; scenario)
    py = get_hybridproblem_neg_logden_obs(prob; scenario)
    priors = get_hybridproblem_priors(prob; scenario)
    xM, xP, y_o, y_unc, i_sites = first(get_hybridproblem_train_dataloader(prob; scenario))
    n_batch = size(xM, 2)
    transMs = HVI.StackedArray(transM, n_batch)
    pbm_covars = get_hybridproblem_pbmpar_covars(prob; scenario)
    pbm_covar_indices = HVI.get_pbm_covar_indices(pt.θP, pbm_covars)
    priorsP = Tuple(priors[k] for k in keys(pt.θP)); priorsM = Tuple(priors[k] for k in keys(pt.θM))
    FT = eltype(ϕ)
    rng = StableRNG(211)
    zP = randn(rng, FT, n_MC, length(pt.θP)); zMs = randn(rng, FT, n_MC, length(pt.θM) * n_batch)
    approx = HVI.MeanHVIApproximationMat()
    # same body as neg_elbo_gtf_components (src/elbo.jl:46-88) with the draws passed in
    function cost(ϕv)
        ϕc = interpreters.ϕg_ϕq(ϕv)
        ϕqv = CA.getdata(ϕc.ϕq); μP = interpreters.ϕq(ϕqv)[Val(:μP)]
        xMP0 = HVI._append_each_covars(xM, CA.getdata(μP), pbm_covar_indices)
        μ0 = g(xMP0, CA.getdata(ϕc.ϕg); is_testmode=false)
        rP, rM, σ = HVI.sample_ζresid_norm(approx, i_sites, zP, zMs, μ0, ϕqv; cor_ends, int_ϕq=interpreters.ϕq)
        ζsP = μP .+ rP
        ζsMs = isempty(pbm_covars) ? permutedims(μ0 .+ rM, (2, 1, 3)) :
               stack(map((ζP, r) -> (g(HVI._append_each_covars(xM, CA.getdata(ζP), pbm_covar_indices), CA.getdata(ϕc.ϕg); is_testmode=false) .+ r)',
                         eachcol(ζsP), eachslice(rM; dims=3)))
        c = HVI.neg_elbo_ζtf(ζsP, ζsMs, σ, f, py, xP, y_o, y_unc; transP, transMs, priorsP, priorsM,
            floss_penalty=HVI.zero_penalty_loss, ϕg=CA.getdata(ϕc.ϕg), ϕq=ϕqv, is_omit_priors=Val(false),
            zero_prior_logdensity=zero(FT))
        c.nLjoint - c.entropy_ζ + c.loss_penalty, c
    end
    ϕv = Vector(CA.getdata(ϕ))
    (nL, c), gr = Zygote.withgradient(v -> cost(v)[1], ϕv)[1], Zygote.gradient(v -> cost(v)[1], ϕv)[1]
    hdr = Float64[length(ϕv), size(xM, 1), size(y_o, 1), n_batch, n_MC, length(pt.θP), length(pt.θM)]
    open(joinpath(out_dir, name * ".bin"), "w") do io
        for a in (hdr, ϕv, xM, CA.getdata(xP), y_o, y_unc, zP, zMs, [nL],
                  [c.nLjoint, c.entropy_ζ, c.loss_penalty, c.nLy, c.neg_log_prior, c.neg_log_jac], gr)
            write(io, Float64.(vec(collect(a))))
        end
    end
end

out_dir = length(ARGS) > 0 ? ARGS[1] : "."
dump_case("doublemm_default", HVI.DoubleMM.DoubleMMCase(); out_dir)
dump_case("doublemm_covark2", HVI.DoubleMM.DoubleMMCase(); scenario=Val((:covarK2,)), out_dir)
