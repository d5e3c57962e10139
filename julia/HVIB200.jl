# HVIB200.jl -- the reference-side binding of libhvi_b200.so (include/hvi_b200.h).
#
# STATUS: written against the reference's sources (file:line below are relative to the HybridVariationalInference.jl
# checkout) and read field by field against include/hvi_b200.h, but NEVER EXECUTED -- the build image of this
# repository has no Julia.  julia/smoke_hvib200.jl is the two-minute check to run first on a Julia host.
#
# How it drops in.  `solve(prob, HybridPosteriorSolver(...))` (src/HybridSolver.jl:184-267) builds its cost function
# with `get_loss_elbo(g_dev, transP, transMs, f_dev, py; n_MC, cor_ends, pbm_covars, θP, int_ϕq, int_ϕg_ϕq, priorsP,
# priorsM, is_omit_priors, approx, ...)` (:240-245, definition :293-324) and differentiates it with AutoZygote
# (:256-258).  This module adds ONE method of `get_loss_elbo` that dispatches on the type of `g`:
#
#     prob_b200 = HVIB200.b200_problem(prob; scenario)      # same problem, g wrapped in a B200ModelApplicator
#     solve(prob_b200, HybridPosteriorSolver(; alg = Adam(0.02), n_MC = 3); scenario, rng, epochs = 2)   # unchanged call
#
# Everything the method needs is what `solve` hands to `get_loss_elbo`; the remaining sizes (n_covar, n_obs) are read
# from the first minibatch, so the plan is created lazily at the first `loss_elbo` call.  The closure has the
# reference's signature `loss_elbo(ϕ, rng, xM, xP, y_o, y_unc, i_sites; is_testmode)`; its `ChainRulesCore.rrule`
# hands Zygote the gradient computed by the library.
#
# What stays in Julia untouched: HybridProblem, the accessor generics, the DataLoader, the optimiser
# (Optimization.jl + Optimisers.Adam), the ComponentVector layout of ϕ.
module HVIB200

using ChainRulesCore
import HybridVariationalInference as HVI
using HybridVariationalInference: AbstractHybridProblem, AbstractModelApplicator, HybridProblem
using Bijectors, Distributions, Random
import ComponentArrays as CA

const libhvi = get(ENV, "HVI_B200_LIB", "libhvi_b200.so")

const HVI_MAX_PAR = 8
const HVI_MAX_LAYERS = 8
const HVI_ENONFINITE = 3

# ---- C structs (include/hvi_b200.h) --------------------------------------------------------------------------
struct hvi_layer
    n_in::Int32; n_out::Int32; act::Int32; has_bias::Int32
end
struct hvi_prior
    kind::Int32; _pad::Int32; a::Float64; b::Float64
end
# field-for-field mirror of `struct hvi_desc`; NTuples of isbits give the C array layout
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
# Julia lays an isbits struct out by the C rules, so this must reproduce the C side (plan_create checks struct_size):
# 944 bytes, scale_a at 176 (4 bytes of padding after _pad0), prior_P at 368, slot_fix at 856, n_site_total at 936 --
# the numbers tests/test_abi.py::test_desc_layout_known_offsets pins for the C struct.
@assert sizeof(hvi_desc) == 944 "hvi_desc layout differs from include/hvi_b200.h"
@assert fieldoffset(hvi_desc, findfirst(==(:scale_a), fieldnames(hvi_desc))) == 176
@assert fieldoffset(hvi_desc, findfirst(==(:prior_P), fieldnames(hvi_desc))) == 368
@assert fieldoffset(hvi_desc, findfirst(==(:slot_fix), fieldnames(hvi_desc))) == 856
@assert fieldoffset(hvi_desc, findfirst(==(:n_site_total), fieldnames(hvi_desc))) == 936

pad(v, n, z) = ntuple(i -> i <= length(v) ? v[i] : z, n)

# ---- problem pieces → description -----------------------------------------------------------------------------
trans_code(::typeof(identity)) = Int32(0)
trans_code(::HVI.Exp) = Int32(1)
trans_code(::HVI.Logistic) = Int32(2)
trans_code(b) = error("HVIB200: bijector $(typeof(b)) is not one of identity / Exp / Logistic")

"One code per parameter column of a `Stacked` (src/bijectors_utils.jl:65-114)."
function trans_codes(t::Bijectors.Stacked, n)
    codes = zeros(Int32, n)
    for (b, r) in zip(t.bs, t.ranges_in)
        codes[r] .= trans_code(b)
    end
    codes
end

prior_entry(d::Normal) = hvi_prior(1, 0, Float64(d.μ), Float64(d.σ))
prior_entry(d::LogNormal) = hvi_prior(2, 0, Float64(d.μ), Float64(d.σ))
prior_entry(d) = error("HVIB200: prior $(typeof(d)) is not Normal / LogNormal")

"Scaling wrapper of g → (kind, a, b, inner applicator) (src/ModelApplicator.jl:132-176, 184-208)."
scaling(g::HVI.NormalScalingModelApplicator) = (Int32(0), Float64.(g.μ), Float64.(g.σ), g.app)
scaling(g::HVI.RangeScalingModelApplicator) = (Int32(1), Float64.(g.offset), Float64.(g.width), g.app)
scaling(g) = error("HVIB200: g must be wrapped in NormalScalingModelApplicator or RangeScalingModelApplicator, got $(typeof(g))")

"""
    mlp_layers(n_in, n_out, n_ϕg; layers = nothing)

Layer table of g.  Default: the canonical network of `construct_3layer_MLApplicator`
(ext/HybridVariationalInferenceSimpleChainsExt.jl:43-52: n_in → 4n_in tanh → 4n_in tanh → n_out logistic, no bias on the
last layer; parameters per layer `vec(W[out × in])` then `b`), accepted only if its parameter count equals `n_ϕg`.
Any other dense network: pass `layers = [(n_in, n_out, act, has_bias), ...]` with act ∈ (:identity, :tanh, :logistic).
"""
function mlp_layers(n_in, n_out, n_ϕg; layers=nothing)
    act_code = Dict(:identity => 0, :tanh => 1, :logistic => 2)
    tab = isnothing(layers) ?
          [hvi_layer(n_in, 4n_in, 1, 1), hvi_layer(4n_in, 4n_in, 1, 1), hvi_layer(4n_in, n_out, 2, 0)] :
          [hvi_layer(l[1], l[2], act_code[l[3]], l[4] ? 1 : 0) for l in layers]
    n_par = sum(l -> l.n_in * l.n_out + (l.has_bias == 1 ? l.n_out : 0), tab)
    n_par == n_ϕg || error("HVIB200: the layer table has $n_par parameters but ϕg has $n_ϕg; pass `layers = ...`")
    length(tab) <= HVI_MAX_LAYERS || error("HVIB200: more than $HVI_MAX_LAYERS layers")
    tab
end

"""
    make_desc(; FT, pt, g, n_ϕg, f, transP, transM, cor_ends, priorsP, priorsM, pbm_covars, approx, n_covar, n_obs,
              n_site_total, is_omit_priors = false, device = 0, count_globals = true, layers = nothing)

Flatten the keyword set of `get_loss_elbo` / `neg_elbo_gtf` (src/HybridSolver.jl:240-245, src/elbo.jl:46-66) into the C
description.  `pt = (; θP, θM)` are the parameter templates, `f` the PBM applicator whose fixed parameters complete the
(r0, r1, K1, K2) slot map (src/PBMApplicator.jl:230-245; src/DoubleMM/f_doubleMM.jl:47-65).
"""
function make_desc(; FT, pt, g, n_ϕg, f, transP, transM, cor_ends, priorsP, priorsM, pbm_covars, approx, n_covar, n_obs,
        n_site_total, is_omit_priors=false, device=0, count_globals=true, layers=nothing)
    nP, nM = length(pt.θP), length(pt.θM)
    (nP <= HVI_MAX_PAR && nM <= HVI_MAX_PAR) || error("HVIB200: more than $HVI_MAX_PAR parameters per block")
    scale_kind, sa, sb, _ = scaling(g)
    n_in = n_covar + length(pbm_covars)
    tab = mlp_layers(n_in, nM, n_ϕg; layers)
    θFix = hasproperty(f, :θFixm) ? f.θFixm[1, :] : (hasproperty(f, :θFix) ? f.θFix : CA.ComponentVector())
    function slot(name)
        name ∈ keys(pt.θP) && return (Int32(0), Int32(findfirst(==(name), keys(pt.θP)) - 1), 0.0)
        name ∈ keys(pt.θM) && return (Int32(1), Int32(findfirst(==(name), keys(pt.θM)) - 1), 0.0)
        name ∈ keys(θFix) && return (Int32(2), Int32(0), Float64(θFix[name]))
        error("HVIB200: process-model parameter $name is neither in θP, θM nor θFix")
    end
    slots = map(slot, (:r0, :r1, :K1, :K2))
    idx = collect(Int32, HVI.get_pbm_covar_indices(pt.θP, pbm_covars)) .- Int32(1)
    zp = hvi_prior(0, 0, 0.0, 1.0)
    sepvar = approx isa HVI.AbstractMeanVarSepHVIApproximation
    hvi_desc(sizeof(hvi_desc), FT == Float32 ? 0 : 1, device, is_omit_priors ? 1 : 0,
        nP, nM, n_obs, n_covar, length(tab), pad(tab, HVI_MAX_LAYERS, hvi_layer(0, 0, 0, 0)),
        scale_kind, 0, pad(sa, HVI_MAX_PAR, 0.0), pad(sb, HVI_MAX_PAR, 1.0),
        pad(trans_codes(transP, nP), HVI_MAX_PAR, Int32(0)), pad(trans_codes(transM, nM), HVI_MAX_PAR, Int32(0)),
        pad([prior_entry(d) for d in priorsP], HVI_MAX_PAR, zp), pad([prior_entry(d) for d in priorsM], HVI_MAX_PAR, zp),
        length(cor_ends.P), pad(collect(Int32, cor_ends.P), HVI_MAX_PAR, Int32(0)),
        length(cor_ends.M), pad(collect(Int32, cor_ends.M), HVI_MAX_PAR, Int32(0)),
        map(s -> s[1], slots), map(s -> s[2], slots), map(s -> s[3], slots),
        length(idx), pad(idx, HVI_MAX_PAR, Int32(0)), count_globals ? 1 : 0,
        sepvar ? 1 : 0, n_site_total)
end

"""
    make_desc(prob; scenario, approx, n_MC-independent keywords...)

The same description assembled from the problem accessors, the way `solve` reads them (src/HybridSolver.jl:193-234).
`approx` defaults to `prob.approx` for a `HybridProblem` (src/HybridProblem.jl:46) and to the scenario's choice for
other problem types (`:sepvar`, src/DoubleMM/f_doubleMM.jl:408).
"""
function make_desc(prob::AbstractHybridProblem; scenario=Val(()),
        approx=prob isa HybridProblem ? prob.approx :
               (:sepvar ∈ HVI._val_value(scenario) ? HVI.MeanVarSepHVIApproximation() : HVI.MeanHVIApproximationMat()),
        kwargs...)
    pt = HVI.get_hybridproblem_par_templates(prob; scenario)
    (; transP, transM) = HVI.get_hybridproblem_transforms(prob; scenario)
    priors = HVI.get_hybridproblem_priors(prob; scenario)
    g, ϕg0 = HVI.get_hybridproblem_MLapplicator(prob; scenario)
    g isa B200ModelApplicator && (g = g.app)
    xM, xP, y_o, y_unc, i_sites = first(HVI.get_hybridproblem_train_dataloader(prob; scenario))
    make_desc(; FT=eltype(ϕg0), pt, g, n_ϕg=length(ϕg0), f=HVI.get_hybridproblem_PBmodel(prob; scenario), transP, transM,
        cor_ends=HVI.get_hybridproblem_cor_ends(prob; scenario),
        priorsP=[priors[k] for k in keys(pt.θP)], priorsM=[priors[k] for k in keys(pt.θM)],
        pbm_covars=HVI.get_hybridproblem_pbmpar_covars(prob; scenario), approx,
        n_covar=size(xM, 1), n_obs=size(y_o, 1),
        n_site_total=first(HVI.get_hybridproblem_n_site_and_batch(prob; scenario)), kwargs...)
end

# ---- plan -----------------------------------------------------------------------------------------------------
mutable struct Plan
    h::Ptr{Cvoid}
    n_phi::Int
    FT::DataType
    sepvar::Bool
    keep::Any   # host arrays the library may still read (a prefetched batch), kept alive
    function Plan(desc::hvi_desc)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hvi_plan_create, libhvi), Cint, (Ref{hvi_desc}, Ref{Ptr{Cvoid}}), desc, ref)
        rc == 0 || error("hvi_plan_create ($rc): " *
                         unsafe_string(ccall((:hvi_last_error, libhvi), Cstring, (Ptr{Cvoid},), C_NULL)))
        n_phi = ccall((:hvi_phi_len, libhvi), Int64, (Ptr{Cvoid},), ref[])
        p = new(ref[], n_phi, desc.dtype == 0 ? Float32 : Float64, desc.approx == 1, nothing)
        finalizer(q -> ccall((:hvi_plan_destroy, libhvi), Cint, (Ptr{Cvoid},), q.h), p)
    end
end

last_error(p::Plan) = unsafe_string(ccall((:hvi_last_error, libhvi), Cstring, (Ptr{Cvoid},), p.h))
check(p::Plan, rc) = rc == 0 ? nothing : error("hvi_b200 ($rc): " * last_error(p))

"0-based (offset, length) of (ϕg, logσ2_ζP, coef_logσ2_ζMs | logσ2_ζMs, ρsP, ρsM, μP) inside ϕ as the library lays it out."
function phi_positions(p::Plan)
    off = zeros(Int64, 6); len = zeros(Int64, 6)
    check(p, ccall((:hvi_phi_positions, libhvi), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), p.h, off, len))
    off, len
end

"""
    check_layout(p, int_ϕg_ϕq, int_ϕq)

The bit-exact part of the contract: the library's positions must equal those of the reference's interpreters
(`get_positions`, src/ComponentArrayInterpreter.jl:281-288; layout src/init_hybrid_params.jl:25-33).
"""
function check_layout(p::Plan, int_ϕg_ϕq, int_ϕq)
    off, len = phi_positions(p)
    pos = HVI.get_positions(int_ϕg_ϕq)
    posq = HVI.get_positions(int_ϕq)
    nϕg = length(pos.ϕg)
    names = (:logσ2_ζP, p.sepvar ? :logσ2_ζMs : :coef_logσ2_ζMs, :ρsP, :ρsM, :μP)
    (off[1] == 0 && len[1] == nϕg) || error("HVIB200: ϕg position mismatch")
    for (i, nm) in enumerate(names)
        r = vec(collect(posq[nm]))
        want = isempty(r) ? (off[i+1], 0) : (nϕg + first(r) - 1, length(r))
        (len[i+1] == want[2] && (want[2] == 0 || off[i+1] == want[1])) ||
            error("HVIB200: position of $nm is $(off[i+1])+$(len[i+1]) in the library, $(want) in the interpreter")
    end
    length(int_ϕg_ϕq) == p.n_phi || error("HVIB200: length(ϕ) mismatch")
    nothing
end

"Register one DataLoader batch (src/AbstractHybridProblem.jl:179-184); the arrays are copied before the call returns."
function set_data!(p::Plan, xM::Matrix{T}, xP::Matrix{T}, y_o::Matrix{T}, y_unc::Matrix{T}) where {T}
    T == p.FT || error("HVIB200: batch eltype $T differs from eltype(ϕ) $(p.FT)")
    check(p, ccall((:hvi_plan_set_data, libhvi), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Cint), p.h, size(xM, 2), xM, xP, y_o, y_unc, 0))
end

"MeanVarSep: the batch's site indices, ONE-based as in Julia (`i_sites` of loss_elbo, src/elbo_sepvec.jl:23)."
function set_sites!(p::Plan, i_sites)
    idx = collect(Int64, i_sites) .- 1
    check(p, ccall((:hvi_plan_set_sites, libhvi), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), p.h, idx, length(idx)))
end

"Upload the whole training set once (gdev_hybridproblem_dataloader, src/AbstractHybridProblem.jl:220-250)."
function set_dataset!(p::Plan, xM::Matrix{T}, xP::Matrix{T}, y_o::Matrix{T}, y_unc::Matrix{T}) where {T}
    T == p.FT || error("HVIB200: dataset eltype $T differs from eltype(ϕ) $(p.FT)")
    check(p, ccall((:hvi_plan_set_dataset, libhvi), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Cint), p.h, size(xM, 2), xM, xP, y_o, y_unc, 0))
end

"Make the sites `i_sites` (ONE-based) of the device-resident dataset the registered batch: one gather kernel, no upload."
function select_batch!(p::Plan, i_sites)
    if i_sites isa AbstractUnitRange
        check(p, ccall((:hvi_plan_select_range, libhvi), Cint, (Ptr{Cvoid}, Int64, Int64), p.h, first(i_sites) - 1, length(i_sites)))
    else
        idx = collect(Int64, i_sites) .- 1
        check(p, ccall((:hvi_plan_select_batch, libhvi), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Cint), p.h, idx, length(idx), 0))
    end
end

"Double-buffered registration: start the upload of the next DataLoader element while the current one is evaluated."
function prefetch_data!(p::Plan, xM::Matrix{T}, xP::Matrix{T}, y_o::Matrix{T}, y_unc::Matrix{T}) where {T}
    p.keep = (xM, xP, y_o, y_unc)   # must stay alive and unchanged until commit_data!
    check(p, ccall((:hvi_plan_prefetch_data, libhvi), Cint,
        (Ptr{Cvoid}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}), p.h, size(xM, 2), xM, xP, y_o, y_unc))
end

"Make the prefetched batch the current one."
function commit_data!(p::Plan)
    check(p, ccall((:hvi_plan_commit_data, libhvi), Cint, (Ptr{Cvoid},), p.h))
    p.keep = nothing
end

components(c) = (; nLjoint=c[1], entropy_ζ=c[2], loss_penalty=c[3], nLy=c[4], neg_log_prior=c[5], neg_log_jac=c[6])

"""
    neg_elbo_withgradient(p, ϕ, n_MC, seed) -> (nL, components, ∇ϕ)

Value, the six components of `neg_elbo_gtf_components` (src/elbo.jl:200) and the gradient on the registered batch; the
draws ξ come from the device generator keyed by `seed` (stand-in for CUDA.randn, ext/HybridVariationalInferenceCUDAExt.jl:82-86).
"""
function neg_elbo_withgradient(p::Plan, ϕ::Vector{T}, n_MC::Integer, seed::Integer) where {T}
    length(ϕ) == p.n_phi || error("HVIB200: length(ϕ) = $(length(ϕ)), expected $(p.n_phi)")
    comps = zeros(Float64, 6); nL = Ref(0.0); grad = similar(ϕ)
    check(p, ccall((:hvi_neg_elbo_grad_rng, libhvi), Cint,
        (Ptr{Cvoid}, Int32, Ptr{T}, UInt64, Int64, Cint, Ptr{Float64}, Ref{Float64}, Ptr{T}),
        p.h, n_MC, ϕ, seed % UInt64, 0, 0, comps, nL, grad))
    nL[], components(comps), grad
end

"""
    neg_elbo_withgradient(p, ϕ, zP, zMs) -> (nL, components, ∇ϕ)

The same with the draws passed in (bit-reproducible): zP (n_MC × n_θP), zMs (n_MC × n_θM·n_batch), the arrays
`sample_ζresid_norm` draws at src/elbo.jl:605-606.
"""
function neg_elbo_withgradient(p::Plan, ϕ::Vector{T}, zP::Matrix{T}, zMs::Matrix{T}) where {T}
    length(ϕ) == p.n_phi || error("HVIB200: length(ϕ) = $(length(ϕ)), expected $(p.n_phi)")
    comps = zeros(Float64, 6); nL = Ref(0.0); grad = similar(ϕ)
    check(p, ccall((:hvi_neg_elbo_grad, libhvi), Cint,
        (Ptr{Cvoid}, Int32, Ptr{T}, Ptr{T}, Ptr{T}, Cint, Ptr{Float64}, Ref{Float64}, Ptr{T}),
        p.h, size(zMs, 1), ϕ, zP, zMs, 0, comps, nL, grad))
    nL[], components(comps), grad
end

"`loss_gf` and its gradient for ϕ = [ϕg; ϕP] (src/gf.jl:227-295), the point solver's cost function."
# ---- site-sharded jobs: the collective lives in the library (ncclAllReduce on the evaluation's stream) ------------
# rank 0: id = nccl_unique_id(); hand the 128 bytes to every rank (MPI.bcast, a file, ...); every rank: comm_init!(p, id, rank, world)
# with a plan created with count_globals = (rank == 0) and its own site range registered by set_data!.
function nccl_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:hvi_nccl_unique_id, libhvi), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("hvi_nccl_unique_id failed ($rc): libnccl.so.2 not loadable?")
    id
end
comm_init!(p::Plan, id::Vector{UInt8}, rank::Integer, world::Integer) =
    check(p, ccall((:hvi_plan_comm_init, libhvi), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), p.h, id, rank, world))
comm_destroy!(p::Plan) = check(p, ccall((:hvi_plan_comm_destroy, libhvi), Cint, (Ptr{Cvoid},), p.h))

"(nL, comps, ∇ϕ) of the whole job on every rank; zMs_shard: this rank's columns of the draws; first_site: 0-based"
function neg_elbo_withgradient_sharded(p::Plan, ϕ::Vector{T}, zP::Matrix{T}, zMs_shard::Matrix{T}, first_site::Integer) where {T}
    comps = zeros(Float64, 6); nL = Ref{Float64}(0.0); grad = similar(ϕ)
    check(p, ccall((:hvi_neg_elbo_grad_sharded, libhvi), Cint,
                   (Ptr{Cvoid}, Int32, Ptr{T}, Ptr{T}, Ptr{T}, UInt64, Int64, Cint, Ptr{Float64}, Ref{Float64}, Ptr{T}),
                   p.h, size(zMs_shard, 1), ϕ, zP, zMs_shard, 0, first_site, 0, comps, nL, grad))
    nL[], comps, grad
end

function loss_gf_withgradient(p::Plan, ϕ::Vector{T}) where {T}
    comps = zeros(Float64, 3); grad = similar(ϕ)
    check(p, ccall((:hvi_loss_gf_grad, libhvi), Cint, (Ptr{Cvoid}, Ptr{T}, Cint, Ptr{Float64}, Ptr{T}), p.h, ϕ, 0, comps, grad))
    (; nLjoint_pen=comps[1], nLy=comps[2], neg_log_prior=comps[3]), grad
end

"""
    adam_solve_minibatches!(p, ϕ0, n_iter, n_MC, n_batch; eta, betas, eps, seed0) -> (ϕ, nL_trace)

The whole minibatch loop of `solve` (src/HybridSolver.jl:259-260 over DataLoader(...; batchsize = n_batch,
partial = false), src/AbstractHybridProblem.jl:206-207) on the device-resident dataset (`set_dataset!` first).
"""
function adam_solve_minibatches!(p::Plan, ϕ0::Vector{T}, n_iter::Integer, n_MC::Integer, n_batch::Integer;
        eta=0.02, betas=(0.9, 0.999), eps=1e-8, seed0=0) where {T}
    check(p, ccall((:hvi_adam_init, libhvi), Cint, (Ptr{Cvoid}, Ptr{T}, Cdouble, Cdouble, Cdouble, Cdouble),
        p.h, ϕ0, eta, betas[1], betas[2], eps))
    trace = Vector{Float64}(undef, n_iter)
    check(p, ccall((:hvi_adam_run_minibatches, libhvi), Cint,
        (Ptr{Cvoid}, Int32, Int32, UInt64, Int64, Int64, Ptr{Float64}), p.h, n_iter, n_MC, seed0, n_batch, 0, trace))
    ϕ = similar(ϕ0)
    check(p, ccall((:hvi_adam_get_phi, libhvi), Cint, (Ptr{Cvoid}, Ptr{T}), p.h, ϕ))
    ϕ, trace
end

# ---- the drop-in: a method of get_loss_elbo selected by the type of g -------------------------------------------
"""
    B200ModelApplicator(app, pt; device = 0, dataset = nothing, layers = nothing)

Wraps the problem's ML applicator.  As a model applicator it behaves exactly like `app` (so `predict_hvi`, the point
solver, ... keep working); as the first argument of `get_loss_elbo` it selects the B200 cost function.  `pt` are the
parameter templates `(; θP, θM)`.  `dataset = (xM, xP, y_o, y_unc)` (the DataLoader's `.data`) uploads the training
set once; every `loss_elbo` call then only sends `i_sites`.
"""
struct B200ModelApplicator{A,PT} <: AbstractModelApplicator
    app::A
    pt::PT
    device::Int
    dataset::Any
    layers::Any
end
B200ModelApplicator(app, pt; device=0, dataset=nothing, layers=nothing) = B200ModelApplicator(app, pt, device, dataset, layers)
HVI.apply_model(a::B200ModelApplicator, x, ϕ; kwargs...) = HVI.apply_model(a.app, x, ϕ; kwargs...)

"""
    b200_problem(prob; scenario, device = 0, resident_dataset = true, layers = nothing)

`prob` with its ML applicator wrapped: `solve(b200_problem(prob; scenario), solver; scenario, ...)` runs the
neg-ELBO and its gradient on the GPU library; nothing else about the problem changes.
"""
function b200_problem(prob::AbstractHybridProblem; scenario=Val(()), device=0, resident_dataset=true, layers=nothing)
    g, ϕg = HVI.get_hybridproblem_MLapplicator(prob; scenario)
    pt = HVI.get_hybridproblem_par_templates(prob; scenario)
    dataset = nothing
    if resident_dataset
        xM, xP, y_o, y_unc, _ = HVI.get_hybridproblem_train_dataloader(prob; scenario).data
        dataset = (Matrix(CA.getdata(xM)), Matrix(CA.getdata(xP)), Matrix(y_o), Matrix(y_unc))
    end
    approx = prob isa HybridProblem ? prob.approx :
             (:sepvar ∈ HVI._val_value(scenario) ? HVI.MeanVarSepHVIApproximation() : HVI.MeanHVIApproximationMat())
    HybridProblem(prob; scenario, approx, g=B200ModelApplicator(g, pt; device, dataset, layers), ϕg)
end

mutable struct LossState
    plan::Union{Nothing,Plan}
    make::Any     # (n_covar, n_obs, FT) -> Plan
end

"""
    get_loss_elbo(g::B200ModelApplicator, transP, transMs, f, py; n_MC, cor_ends, pbm_covars, θP, int_ϕq, int_ϕg_ϕq,
                  priorsP, priorsM, is_omit_priors, approx, ...)

Method of the reference's `get_loss_elbo` (src/HybridSolver.jl:293-324) with the same keyword set (the ones this path
does not use -- n_MC_mean, n_MC_cap, priors_θP_mean, priors_θMs_mean, cdev, zero_prior_logdensity -- are accepted and
ignored; a non-default `floss_penalty` is refused).  Returns `loss_elbo(ϕ, rng, xM, xP, y_o, y_unc, i_sites; is_testmode)`.
"""
function HVI.get_loss_elbo(g::B200ModelApplicator, transP, transMs, f, py;
        n_MC, cor_ends, pbm_covars, θP, int_ϕq, int_ϕg_ϕq, priorsP, priorsM,
        is_omit_priors, approx, floss_penalty=HVI.zero_penalty_loss, kwargs...)
    floss_penalty === HVI.zero_penalty_loss || error("HVIB200: floss_penalty is not supported on this path")
    py === HVI.neg_logden_indep_normal || error("HVIB200: the likelihood must be neg_logden_indep_normal")
    omit = is_omit_priors isa Val ? HVI._val_value(is_omit_priors) : Bool(is_omit_priors)
    sepvar = approx isa HVI.AbstractMeanVarSepHVIApproximation
    n_ϕg = length(HVI.get_positions(int_ϕg_ϕq).ϕg)
    n_site_total = sepvar ? size(HVI.get_positions(int_ϕq).logσ2_ζMs, 2) : 0
    make = (n_covar, n_obs, FT) -> begin
        desc = make_desc(; FT, pt=g.pt, g=g.app, n_ϕg, f, transP, transM=transMs.stacked, cor_ends, priorsP, priorsM,
            pbm_covars, approx, n_covar, n_obs, n_site_total, is_omit_priors=omit, device=g.device, layers=g.layers)
        plan = Plan(desc)
        check_layout(plan, int_ϕg_ϕq, int_ϕq)
        isnothing(g.dataset) || set_dataset!(plan, g.dataset...)
        plan
    end
    state = LossState(nothing, make)
    resident = !isnothing(g.dataset)
    function loss_elbo(ϕ, rng, xM, xP, y_o, y_unc, i_sites; is_testmode=false)
        _neg_elbo(state, resident, n_MC, ϕ, rng, xM, xP, y_o, y_unc, i_sites)
    end
end

# primal and gradient in one library call; returns (nL::Float64, ∇ϕ)
function _evaluate(state::LossState, resident, n_MC, ϕ, rng, xM, xP, y_o, y_unc, i_sites)
    ϕv = Vector(CA.getdata(ϕ))
    if isnothing(state.plan)
        state.plan = state.make(size(xM, 1), size(y_o, 1), eltype(ϕv))
    end
    plan = state.plan
    if resident
        select_batch!(plan, i_sites)          # the batch is chosen on the device; with MeanVarSep this also sets i_sites
    else
        # every call registers its batch: a loader that refills its buffers in place is always seen
        set_data!(plan, Matrix(CA.getdata(xM)), Matrix(CA.getdata(xP)), Matrix(y_o), Matrix(y_unc))
        plan.sepvar && set_sites!(plan, i_sites)
    end
    nL, comps, grad = neg_elbo_withgradient(plan, ϕv, n_MC, rand(rng, UInt64))
    nL, grad
end

_neg_elbo(state, resident, n_MC, ϕ, rng, xM, xP, y_o, y_unc, i_sites) =
    first(_evaluate(state, resident, n_MC, ϕ, rng, xM, xP, y_o, y_unc, i_sites))

function ChainRulesCore.rrule(::typeof(_neg_elbo), state, resident, n_MC, ϕ, rng, xM, xP, y_o, y_unc, i_sites)
    nL, grad = _evaluate(state, resident, n_MC, ϕ, rng, xM, xP, y_o, y_unc, i_sites)
    # one tangent per argument of _neg_elbo (after the function itself): only ϕ is differentiated
    pullback(Δ) = (NoTangent(), NoTangent(), NoTangent(), NoTangent(), eltype(grad)(Δ) .* grad,
                   NoTangent(), NoTangent(), NoTangent(), NoTangent(), NoTangent(), NoTangent())
    nL, pullback
end

end # module
