# replay_rng.jl -- an RNG that hands the reference pre-drawn standard normals (shared by dump_golden.jl and
# smoke_hvib200.jl).  `_create_randn(rng, v, dims...) = randn(rng, eltype(v), dims...)` (src/elbo.jl:771-773) is the only
# place the neg-ELBO path draws; it is called first for zP (n_MC × n_θP), then for zMs (n_MC × n_θM·n_batch) (:605-606).
using Random, Zygote

"An RNG whose `randn(rng, T, dims...)` calls return the queued arrays in order (they must match in type and size)."
mutable struct ReplayRNG <: Random.AbstractRNG
    queue::Vector{Any}
    pos::Int
end
ReplayRNG(arrays...) = ReplayRNG(Any[arrays...], 0)
function Random.randn(r::ReplayRNG, ::Type{T}, d1::Integer, dims::Integer...) where {T}
    r.pos = r.pos % length(r.queue) + 1          # wraps: every evaluation (primal, gradient) replays the same draws
    a = r.queue[r.pos]
    (eltype(a) == T && size(a) == (d1, dims...)) ||
        error("ReplayRNG: asked for $(T) $((d1, dims...)), queued $(eltype(a)) $(size(a))")
    copy(a)
end
Zygote.@nograd Random.randn
