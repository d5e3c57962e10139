# HVIB200.jl -- the reference-side binding of libhvi_b200.so (include/hvi_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is the stub a
# maintainer of HybridVariationalInference.jl would add (as a package extension, like
# ext/HybridVariationalInferenceCUDAExt.jl) so that `solve(prob, HybridPosteriorSolver(...))`
# evaluates the negative ELBO and its gradient on B200 instead of through Zygote.
#
# What it replaces, with file:line in the reference:
#   * the closure built by get_loss_elbo (src/HybridSolver.jl:293-324) and differentiated by
#     AutoZygote (src/HybridSolver.jl:256-258)   -> loss_elbo_b200 + ChainRulesCore.rrule
#   * neg_elbo_gtf / neg_elbo_gtf_components (src/elbo.jl:30-88)   -> hvi_neg_elbo_grad
#   * generate_ζ (src/elbo.jl:472-521)                             -> hvi_generate_zeta
# What stays in Julia untouched: HybridProblem, the accessor generics, the DataLoader, the
# optimiser (Optimization.jl + Optimisers.Adam), the ComponentVector layout of ϕ.
module HVIB200

using ChainRulesCore
import HybridVariationalInference as HVI
using HybridVariationalInference: AbstractHybridProblem, HybridPosteriorSolver
using Bijectors, Distributions, ComponentArrays
import ComponentArrays as CA

const libhvi = get(ENV, "HVI_B200_LIB", "libhvi_b200.so")

const HVI_MAX_PAR = 8
const HVI_MAX_LAYERS = 8

struct hvi_layer
    n_in::Int32; n_out::Int32; act::Int32; has_bias::Int32
end
struct hvi_prior
    kind::Int32; _pad::Int32; a::Float64; b::Float64
end
# field-for-field mirror of `struct hvi_desc` (include/hvi_b200.h); NTuples give the C layout
struct hvi_desc
    struct_size::Int32; dtype::Int32; device::Int32; flags::Int32
    n_P::Int32; n_M::Int32; n_obs::Int32; n_covar::Int32
    n_layers::Int32
    layers::NTuple{HVI_MAX_LAYERS,hvi_layer}
    scale_kind::Int32; _pad0::Int32
    scale_a::NTuple{HVI_MAX_PAR,Float64}; scale_b::NTuple{HVI_MAX_PAR,Float64}
    trans_P::NTuple{HVI_MAX_PAR,Int32}; trans_M::NTuple{HVI_MAX_PAR,Int32}
    prior_P::NTuple{HVI_MAX_PAR,hvi_prior}; prior_M::NTuple{HVI_MAX_PAR,hvi_prior}
    n_cor_ends_P::Int32; cor_ends_P::NTuple{HVI_MAX_PAR,Int32}
    n_cor_ends_M::Int32; cor_ends_M::NTuple{HVI_MAX_PAR,Int32}
    slot_src::NTuple{4,Int32}; slot_idx::NTuple{4,Int32}; slot_fix::NTuple{4,Float64}
    n_pbm_covar::Int32; pbm_covar_idx::NTuple{HVI_MAX_PAR,Int32}
    count_globals::Int32
    approx::Int32
    n_site_total::Int64
end

pad(v, n, z) = ntuple(i -> i <= length(v) ? v[i] : z, n)

trans_code(::typeof(identity)) = Int32(0)
trans_code(::HVI.Exp) = Int32(1)
trans_code(::HVI.Logistic) = Int32(2)
"one code per parameter column of a `Stacked` (src/bijectors_utils.jl:65-114)"
function trans_codes(t::Bijectors.Stacked, n)
    codes = zeros(Int32, n)
    for (b, r) in zip(t.bs, t.ranges_in)
        codes[r] .= trans_code(b)
    end
    codes
end
prior_entry(d::Normal) = hvi_prior(1, 0, d.μ, d.σ)
prior_entry(d::LogNormal) = hvi_prior(2, 0, d.μ, d.σ)

"""
    make_desc(prob; scenario, device=0, count_globals=true)

Flatten what `solve` reads from the problem accessors (src/HybridSolver.jl:193-234) into the C
description.  The MLP shape is the canonical 3-layer network of
`construct_3layer_MLApplicator` (ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52).
"""
function make_desc(prob::AbstractHybridProblem; scenario=Val(()), device=0, count_globals=true)
    pt = HVI.get_hybridproblem_par_templates(prob; scenario)
    (; transP, transM) = HVI.get_hybridproblem_transforms(prob; scenario)
    cor_ends = HVI.get_hybridproblem_cor_ends(prob; scenario)
    priors = HVI.get_hybridproblem_priors(prob; scenario)
    pbm_covars = HVI.get_hybridproblem_pbmpar_covars(prob; scenario)
    g, ϕg0 = HVI.get_hybridproblem_MLapplicator(prob; scenario)      # NormalScalingModelApplicator
    FT = eltype(ϕg0)
    nP, nM = length(pt.θP), length(pt.θM)
    n_covar = HVI.get_hybridproblem_n_covar(prob; scenario)
    n_in = n_covar + length(pbm_covars)
    layers = (hvi_layer(n_in, 4n_in, 1, 1), hvi_layer(4n_in, 4n_in, 1, 1), hvi_layer(4n_in, nM, 2, 0))
    f = HVI.get_hybridproblem_PBmodel(prob; scenario)
    θFix = f.θFixm[1, :]
    slot(name) = name ∈ keys(pt.θP) ? (Int32(0), Int32(findfirst(==(name), keys(pt.θP)) - 1), 0.0) :
                 name ∈ keys(pt.θM) ? (Int32(1), Int32(findfirst(==(name), keys(pt.θM)) - 1), 0.0) :
                 (Int32(2), Int32(0), Float64(θFix[name]))
    slots = map(slot, (:r0, :r1, :K1, :K2))                          # src/PBMApplicator.jl:234-239
    idx = collect(Int32, HVI.get_pbm_covar_indices(pt.θP, pbm_covars)) .- Int32(1)
    zp = hvi_prior(0, 0, 0.0, 1.0)
    hvi_desc(sizeof(hvi_desc), FT == Float32 ? 0 : 1, device, 0,
        nP, nM, 8, n_covar, 3, pad(collect(layers), HVI_MAX_LAYERS, hvi_layer(0, 0, 0, 0)),
        g isa HVI.RangeScalingModelApplicator ? 1 : 0, 0,
        pad(Float64.(g isa HVI.RangeScalingModelApplicator ? g.offset : g.μ), HVI_MAX_PAR, 0.0),
        pad(Float64.(g isa HVI.RangeScalingModelApplicator ? g.width : g.σ), HVI_MAX_PAR, 1.0),
        pad(trans_codes(transP, nP), HVI_MAX_PAR, Int32(0)), pad(trans_codes(transM, nM), HVI_MAX_PAR, Int32(0)),
        pad([prior_entry(priors[k]) for k in keys(pt.θP)], HVI_MAX_PAR, zp),
        pad([prior_entry(priors[k]) for k in keys(pt.θM)], HVI_MAX_PAR, zp),
        length(cor_ends.P), pad(Int32.(cor_ends.P), HVI_MAX_PAR, Int32(0)),
        length(cor_ends.M), pad(Int32.(cor_ends.M), HVI_MAX_PAR, Int32(0)),
        map(first, slots), map(s -> s[2], slots), map(last, slots),
        length(idx), pad(idx, HVI_MAX_PAR, Int32(0)), count_globals ? 1 : 0,
        prob.approx isa HVI.AbstractMeanVarSepHVIApproximation ? 1 : 0,
        first(HVI.get_hybridproblem_n_site_and_batch(prob; scenario)))
end

mutable struct Plan
    h::Ptr{Cvoid}
    n_phi::Int
    FT::DataType
    next_batch::Any   # host arrays of a prefetched batch, kept alive until commit_data!
    function Plan(desc::hvi_desc)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hvi_plan_create, libhvi), Cint, (Ref{hvi_desc}, Ref{Ptr{Cvoid}}), desc, ref)
        rc == 0 || error("hvi_plan_create: " * unsafe_string(ccall((:hvi_last_error, libhvi), Cstring, (Ptr{Cvoid},), C_NULL)))
        p = new(ref[], ccall((:hvi_phi_len, libhvi), Int64, (Ptr{Cvoid},), ref[]), desc.dtype == 0 ? Float32 : Float64, nothing)
        finalizer(q -> ccall((:hvi_plan_destroy, libhvi), Cint, (Ptr{Cvoid},), q.h), p)
    end
end

check(p::Plan, rc) = rc == 0 ? nothing :
    error("hvi_b200 ($rc): " * unsafe_string(ccall((:hvi_last_error, libhvi), Cstring, (Ptr{Cvoid},), p.h)))

"register one DataLoader batch (src/AbstractHybridProblem.jl:179-184)"
function set_data!(p::Plan, xM::Matrix{T}, xP::Matrix{T}, y_o::Matrix{T}, y_unc::Matrix{T}) where {T}
    check(p, ccall((:hvi_plan_set_data, libhvi), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Cint), p.h, size(xM, 2), xM, xP, y_o, y_unc, 0))
end

"value, six components and gradient; ξ drawn on the device from `seed` (cf. ext/…CUDAExt.jl:82-86)"
# double-buffered registration: upload the next DataLoader element while the current one is evaluated
function prefetch_data!(p::Plan, xM::Matrix{T}, xP::Matrix{T}, y_o::Matrix{T}, y_unc::Matrix{T}) where {T}
    p.next_batch = (xM, xP, y_o, y_unc)   # keep the host arrays alive and untouched until commit_data!
    check(p, ccall((:hvi_plan_prefetch_data, libhvi), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}),
        p.h, size(xM, 2), xM, xP, y_o, y_unc))
end
commit_data!(p::Plan) = check(p, ccall((:hvi_plan_commit_data, libhvi), Cint, (Ptr{Cvoid},), p.h))

function neg_elbo_withgradient(p::Plan, ϕ::Vector{T}, n_MC::Integer, seed::Integer) where {T}
    comps = zeros(Float64, 6); nL = Ref(0.0); grad = similar(ϕ)
    check(p, ccall((:hvi_neg_elbo_grad_rng, libhvi), Cint,
        (Ptr{Cvoid}, Int32, Ptr{T}, UInt64, Int64, Cint, Ptr{Float64}, Ref{Float64}, Ptr{T}),
        p.h, n_MC, ϕ, seed, 0, 0, comps, nL, grad))
    nL[], (; nLjoint=comps[1], entropy_ζ=comps[2], loss_penalty=comps[3], nLy=comps[4],
             neg_log_prior=comps[5], neg_log_jac=comps[6]), grad
end

"same with the draws passed in (bit-reproducible): zP (n_MC × n_θP), zMs (n_MC × n_θM·n_batch), src/elbo.jl:605-606"
function neg_elbo_withgradient(p::Plan, ϕ::Vector{T}, zP::Matrix{T}, zMs::Matrix{T}) where {T}
    comps = zeros(Float64, 6); nL = Ref(0.0); grad = similar(ϕ)
    check(p, ccall((:hvi_neg_elbo_grad, libhvi), Cint,
        (Ptr{Cvoid}, Int32, Ptr{T}, Ptr{T}, Ptr{T}, Cint, Ptr{Float64}, Ref{Float64}, Ptr{T}),
        p.h, size(zMs, 1), ϕ, zP, zMs, 0, comps, nL, grad))
    nL[], comps, grad
end

"""
    get_loss_elbo_b200(prob; scenario, n_MC, device)

Drop-in for `get_loss_elbo` (src/HybridSolver.jl:293-324): returns `loss_elbo(ϕ, rng, xM, xP, y_o,
y_unc, i_sites; is_testmode)` with the same signature; its `rrule` hands Zygote the gradient
computed by the library, so `OptimizationFunction(..., AutoZygote())` (src/HybridSolver.jl:256-258)
works unchanged.
"""
"loss_gf and its gradient for ϕ = [ϕg; ϕP] (src/gf.jl:227-295), the point solver's cost function"
function loss_gf_withgradient(p::Plan, ϕ::Vector{T}) where {T}
    comps = zeros(Float64, 3); grad = similar(ϕ)
    check(p, ccall((:hvi_loss_gf_grad, libhvi), Cint, (Ptr{Cvoid}, Ptr{T}, Cint, Ptr{Float64}, Ptr{T}), p.h, ϕ, 0, comps, grad))
    (; nLjoint_pen = comps[1], nLy = comps[2], neg_log_prior = comps[3]), grad
end

"n_iter Adam iterations on the registered batch without leaving the GPU (Optimization.solve with Adam, src/HybridSolver.jl:259-260)"
function adam_solve!(p::Plan, ϕ0::Vector{T}, n_iter::Integer, n_MC::Integer; eta=0.02, betas=(0.9, 0.999), eps=1e-8, seed0=0) where {T}
    check(p, ccall((:hvi_adam_init, libhvi), Cint, (Ptr{Cvoid}, Ptr{T}, Cdouble, Cdouble, Cdouble, Cdouble),
        p.h, ϕ0, eta, betas[1], betas[2], eps))
    trace = Vector{Float64}(undef, n_iter)
    check(p, ccall((:hvi_adam_run, libhvi), Cint, (Ptr{Cvoid}, Int32, Int32, UInt64, Int64, Ptr{Float64}),
        p.h, n_iter, n_MC, seed0, 0, trace))
    ϕ = similar(ϕ0)
    check(p, ccall((:hvi_adam_get_phi, libhvi), Cint, (Ptr{Cvoid}, Ptr{T}), p.h, ϕ))
    ϕ, trace
end

function get_loss_elbo_b200(prob; scenario=Val(()), n_MC=12, device=0)
    plan = Plan(make_desc(prob; scenario, device))
    last_batch = Ref{Any}(nothing)
    function loss_elbo(ϕ, rng, xM, xP, y_o, y_unc, i_sites; is_testmode=false)
        first(_eval(plan, last_batch, ϕ, rng, xM, xP, y_o, y_unc, n_MC))
    end
    loss_elbo, plan
end

function _eval(plan, last_batch, ϕ, rng, xM, xP, y_o, y_unc, n_MC)
    if last_batch[] !== xM            # a new minibatch: re-register it
        set_data!(plan, Matrix(xM), Matrix(CA.getdata(xP)), Matrix(y_o), Matrix(y_unc))
        last_batch[] = xM
    end
    neg_elbo_withgradient(plan, Vector(CA.getdata(ϕ)), n_MC, rand(rng, UInt64))
end

function ChainRulesCore.rrule(::typeof(_eval), plan, last_batch, ϕ, rng, xM, xP, y_o, y_unc, n_MC)
    nL, comps, grad = _eval(plan, last_batch, ϕ, rng, xM, xP, y_o, y_unc, n_MC)
    pullback(Δ) = (NoTangent(), NoTangent(), NoTangent(), first(Δ) .* grad, ntuple(_ -> NoTangent(), 6)...)
    (nL, comps, grad), pullback
end

end # module
