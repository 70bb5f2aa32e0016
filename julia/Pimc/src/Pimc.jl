# Pimc.jl -- drop-in shim: the exported API of oameye/PIMC.jl (src/Pimc.jl:15-26) over libpimc_b200.so (include/pimc_b200.h).
#
# UNEXECUTED in this repository's CI: the build image has no Julia.  The same C entry points, in the same order, are exercised through the
# Python mirror pimc_jl_b200/pimc.py and by the C drivers tests/c/example_*.c (the call sequence of the shipped scripts through the C ABI).
# The shipped example scripts (examples/energy_2d_*.jl, density_2d_*.jl, density_SRL_lattice.jl) are written against the surface this
# module provides: `System`, `build_prop_int`, the update constructors, `Energy`/`Density`, `run!`, `acceptance`, and the fields the
# example tools read (s.N, s.M, s.L, s.β, s.τ, s.μ, s.a, s.Ncycle, s.N_MC, s.world[n].{r,V,bins,next}, s.nn, s.nbs, s.lnV/lnK/lnU,
# u.var.size/m, u.counter_var.queue, d.dens ...).
#
# Environment: PIMC_B200_LIB (path of the .so), PIMC_CHAINS (independent replicas, default 1), PIMC_CHAIN_OFFSET (first global chain of this worker), PIMC_SEED,
#              PIMC_SCHED = "faithful" (default: exactly run!) | "sweep" (batched schedule, DESIGN.md section 3).
module Pimc

using Libdl, Random

export System, run!, build_prop_int, acceptance, Coord, determine_nnrange
export Worldline, Particle, levy!, distance, update_nnbins!, disallowmissing, apply!, bin, lnV
export subcycle, pcycle, Update, Updates
export Counter, Step, NumbOfSlices
export SingleCenterOfMass, PolymerCenterOfMass, ReshapeLinear, ReshapeSwapLinear
export Density, Measurement, ZMeasurement, Energy
export PairCorrelation, Winding, superfluid_fraction, StructureFactor, compressibility   # the TODOs of src/measurement.jl:125-127

const LIB = get(ENV, "PIMC_B200_LIB", joinpath(@__DIR__, "..", "..", "..", "pimc_jl_b200", "libpimc_b200.so"))

# ---- src/types.jl, src/utils.jl -------------------------------------------------------------------------------------------
abstract type Worldline end
abstract type Measurement end
abstract type ZMeasurement <: Measurement end
abstract type Update end
abstract type UpdateC <: Update end
Updates = Vector{Tuple{Int64,A}} where {A<:Update}
ZMeasurements = Vector{ZMeasurement}
VectorMissing{T} = Vector{Union{Missing,T}}
Coord = Union{Vector{Float64},SubArray{Float64,1}}
disallowmissing(x::AbstractArray{T}) where {T} = convert(AbstractArray{nonmissingtype(T)}, x)

# ---- C structs (include/pimc_b200.h) --------------------------------------------------------------------------------------
struct CPotential
    kind::Int32; dv_kind::Int32; k::Float64; depth::Float64; scale::Float64; sgn::Float64
    nang::Int32; helical::Int32; ang::NTuple{32,Float64}
end
struct CConfig
    dim::Int32; M::Int32; N::Int32; chains::Int32; chain_offset::UInt32
    mu::Float64; lambda::Float64; L::Float64; T::Float64
    interactions::Int32; g::Float64; r_a::Float64; Ncycle::Int32; compat::Int32; init::Int32
    seed::UInt64; pot::CPotential; tab::Ptr{Float64}; tab_n::Int32; tab_lo::Float64; tab_hi::Float64; device::Int32
end
mutable struct RunStats
    iterations::Int64; proposals::Int64; accepted::Int64; bead_moves::Int64; measurements::Int64; launches::Int64; kernel_ms::Float64
    RunStats() = new(0, 0, 0, 0, 0, 0, 0.0)
end

function check(h::Ptr{Cvoid}, rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:pimc_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("libpimc_b200 ($rc): $msg")
end

# ---- potential closures -> descriptors (a CUDA kernel cannot call a Julia closure) ----------------------------------------
const POT_ZERO, POT_HARMONIC, POT_SIN2_1D, POT_LATTICE = Int32(0), Int32(1), Int32(2), Int32(3)
pad32(v) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 32)
function descriptor_value(p::CPotential, r::Vector{Float64})
    if p.kind == POT_ZERO
        return 0.0
    elseif p.kind == POT_HARMONIC
        return (0.5 * p.k) * sum(abs2, r)
    elseif p.kind == POT_SIN2_1D
        return p.depth * sin(2π * r[1] * p.scale)^2
    else
        s = c = 0.0
        for i in 1:p.nang
            a = p.ang[i]; rr = r[1] * sin(a) + (length(r) > 1 ? r[2] : 0.0) * cos(a)
            ph = 2π * rr * p.scale + (p.helical != 0 ? a : 0.0)
            s += sin(ph); c += cos(ph)
        end
        s /= p.nang; c /= p.nang
        return (p.sgn * p.depth) * (s * s + c * c)
    end
end
capt(f, name) = name in fieldnames(typeof(f)) ? (x = getfield(f, name); x isa Core.Box ? x.contents : x) : nothing
"""Map the closures the shipped scripts use to a descriptor and verify the match at random points (<= 1e-12)."""
function lower_potential(V::Function, dV::Function, dim::Int, L::Float64)
    probe = [L .* (2 .* rand(dim) .- 1) for _ in 1:16]
    vals = [Float64(V(r)) for r in probe]
    ang, scale, depth, sgn = capt(V, :ang), capt(V, :scale), capt(V, :depth), capt(V, :sgn)
    cands = CPotential[]
    push!(cands, CPotential(POT_ZERO, 0, 1.0, 0.0, 1.0, 1.0, 0, 0, pad32(Float64[])))
    push!(cands, CPotential(POT_HARMONIC, 0, 1.0, 0.0, 1.0, 1.0, 0, 0, pad32(Float64[])))
    if ang !== nothing && scale !== nothing && depth !== nothing
        for hel in (0, 1)
            push!(cands, CPotential(POT_LATTICE, 0, 1.0, depth, scale, sgn === nothing ? 1.0 : Float64(sgn), length(ang), hel, pad32(ang)))
        end
    end
    if scale !== nothing && depth !== nothing
        push!(cands, CPotential(POT_SIN2_1D, 0, 1.0, depth, scale, 1.0, 0, 0, pad32(Float64[])))
    end
    for p in cands
        if all(isapprox(descriptor_value(p, r), v; rtol = 1e-12, atol = 1e-13) for (r, v) in zip(probe, vals))
            # dV: `zero` (default), `identity` (the energy examples), anything else -> analytic gradient of the descriptor
            g = dV(probe[1])
            dvk = all(iszero, g) ? Int32(0) : (g == probe[1] ? Int32(1) : Int32(2))
            return CPotential(p.kind, dvk, p.k, p.depth, p.scale, p.sgn, p.nang, p.helical, p.ang)
        end
    end
    error("Pimc (B200): the potential closure is none of the supported families (zero, harmonic, sin^2 1-D, plane-wave lattice); " *
          "arbitrary Julia closures cannot run inside a CUDA kernel")
end

# ---- state (src/system.jl) ------------------------------------------------------------------------------------------------
mutable struct Particle <: Worldline
    r::Matrix{Float64}; V::Vector{Float64}; bins::Vector{Int64}; next::Int64
end
mutable struct System
    h::Ptr{Cvoid}
    dim::Int64; M::Int64; N::Int64; Ninit::Int64; μ::Float64; λ::Float64; L::Float64; vol::Float64; β::Float64; τ::Float64
    a::Float64; nbins::Int64; Ncycle::Int64; measure_scheme::Symbol; chains::Int64; sched::Int32
    V::Function; dV::Function
    tab::Matrix{Float64}; tab_lo::Float64; tab_hi::Float64     # keeps the propagator table alive
    function System(potential::Function; dV::Function = zero, dim::Int64 = 2, M::Int64 = 100, N::Int64 = 2, μ::Float64 = 0.0,
                    L::Float64 = 4.0, T::Float64 = 1.0, λ::Float64 = 1.0, interactions::Bool = false, propint = _ -> 0.0,
                    g::Float64 = 0.0, rₐ::Float64 = 0.0, length_measurement_cycle::Int64 = 10, measure_scheme::Symbol = :c,
                    chains::Int64 = parse(Int64, get(ENV, "PIMC_CHAINS", "1")),
                    seed::UInt64 = parse(UInt64, get(ENV, "PIMC_SEED", string(0x5EEDB200))))
        pot = lower_potential(potential, dV, dim, L)
        tab = zeros(0, 0); lo = hi = 0.0
        if interactions
            if propint isa PropInt               # this module's build_prop_int
                tab = propint.tab; lo = propint.lo; hi = propint.hi
            else                                 # the reference package's closure over the interpolated term table `terms` (src/propagator.jl:80)
                terms = capt(propint, :terms)
                terms === nothing && error("interactions=true needs propint = build_prop_int(L, g, τ)")
                tab = Matrix{Float64}(terms.itp.coefs); lo = first(terms.ranges[1]); hi = last(terms.ranges[1])
            end
            # rₐ == 0.0: init_int (src/system.jl:29-31) takes the cut-off from the propagator; pimc_create does the same (pimc_determine_nnrange).
            # Note the shipped script never forwards g (examples/density_SRL_lattice.jl:18-19): g = 0.0 => a = exp(-2π/0.0) = 0.0 (system.jl:151),
            # no hard core, the interaction enters through lnU only -- reproduced as is.
        end
        cfg = Ref(CConfig(dim, M, N, chains, parse(UInt32, get(ENV, "PIMC_CHAIN_OFFSET", "0")), μ, λ, L, T, interactions, g, rₐ, length_measurement_cycle, 15, 1, seed, pot,
                          isempty(tab) ? C_NULL : pointer(tab), size(tab, 1), lo, hi, -1))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve tab check(C_NULL, ccall((:pimc_create, LIB), Cint, (Ref{CConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        sc = zeros(5); isc = zeros(Int64, 5)
        check(h[], ccall((:pimc_get_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), h[], sc, isc))
        sched = get(ENV, "PIMC_SCHED", "faithful") == "sweep" ? Int32(1) : Int32(0)
        s = new(h[], dim, M, N, N, μ, λ, L, sc[3], sc[1], sc[2], sc[4], isc[1], length_measurement_cycle, measure_scheme, chains, sched,
                potential, dV, tab, lo, hi)
        finalizer(x -> (ccall((:pimc_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), s)
        return s
    end
end
function scalars(s::System)
    sc = zeros(5); isc = zeros(Int64, 5)
    check(s.h, ccall((:pimc_get_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), s.h, sc, isc))
    return sc, isc
end
"""s.world of one replica: Vector{Particle} with r (M x dim), V, bins, next exactly as src/system.jl:1-6 lays them out."""
function world(s::System; chain::Integer = 0)
    r = Array{Float64}(undef, s.M, s.dim, s.N); V = Array{Float64}(undef, s.M, s.N)
    bins = Array{Int64}(undef, s.M, s.N); nxt = Vector{Int64}(undef, s.N)
    check(s.h, ccall((:pimc_get_paths, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
                     s.h, chain, 1, r, V, bins, nxt))
    return Worldline[Particle(r[:, :, n], V[:, n], bins[:, n], nxt[n]) for n in 1:s.N]
end
function Base.getproperty(s::System, f::Symbol)
    f === :world && return world(s)
    f === :N_MC && return Dict{Int64,Int64}(getfield(s, :N) => scalars(s)[2][2])
    f === :Nctr && return Dict{Int64,Int64}(getfield(s, :N) => scalars(s)[2][3])
    f === :ctr && return scalars(s)[2][4]
    f === :lnV && return (x1, x2) -> lnV(x1, x2, getfield(s, :τ), getfield(s, :V))
    f === :lnK && return (x1, x2, τ) -> lnK(x1, x2, getfield(s, :λ), τ, getfield(s, :L))          # slot order of system.jl:163
    f === :lnU && return (r1, r2) -> lnU(s, r1, r2)
    f === :rₐ && return scalars(s)[1][5]
    f === :nbs && return Vector{Int64}[bin_neighbors(i, getfield(s, :nbins), getfield(s, :dim)) for i in 1:getfield(s, :nbins)^getfield(s, :dim)]
    f === :nn && return nn_lists(s)
    return getfield(s, f)
end
"""lnK(x1, x2, λ, τ, L) (src/propagator.jl:16-19), evaluated on the device"""
function lnK(x1::AbstractVector{Float64}, x2::AbstractVector{Float64}, λ::Float64, τ::Float64, L::Float64)::Float64
    out = Ref(0.0); a = Vector{Float64}(x1); b = Vector{Float64}(x2)
    check(C_NULL, ccall((:pimc_lnK, LIB), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Float64, Float64, Ref{Float64}), 1, a, b, length(a), τ, λ, L, out))
    out[]
end
"""s.lnU(r1_rel, r2_rel) (src/system.jl:25-27): 0 without interactions; propint < 0 ? -μ : log(propint)"""
function lnU(s::System, r1::AbstractVector{Float64}, r2::AbstractVector{Float64})::Float64
    isempty(getfield(s, :tab)) && return 0.0
    p = PropInt(getfield(s, :tab), getfield(s, :tab_lo), getfield(s, :tab_hi))(r1, r2, getfield(s, :τ))
    p < 0.0 ? -getfield(s, :μ) : log(p)
end
bin_index(x::Int64, y::Int64, w::Int64)::Int64 = x + w * y + 1                     # src/nearest_neighbours.jl:6-16
bin_index(x::Int64, w::Int64)::Int64 = x + 1
function bin_neighbors(b::Int64, nbins::Int64, dim::Int64)::Vector{Int64}          # src/nearest_neighbours.jl:52-63
    x = rem(b - 1, nbins)
    dim == 2 || return [bin_index(mod(x + dx, nbins), nbins) for dx in (0, -1, 1)]
    y = div(b - 1, nbins)
    [bin_index(mod(x + dx, nbins), mod(y + dy, nbins), nbins) for (dx, dy) in ((0, 0), (-1, 1), (0, 1), (1, 1), (-1, 0), (1, 0), (-1, -1), (0, -1), (1, -1))]
end
"""s.nn (src/system.jl:109): per time slice, per cell, the particles in it -- rebuilt from the device's bins (list order: ascending index)"""
function nn_lists(s::System; chain::Integer = 0)
    w = world(s; chain = chain); nb = getfield(s, :nbins)^getfield(s, :dim)
    nn = [[Int64[] for _ in 1:nb] for _ in 1:getfield(s, :M)]
    for (i, p) in enumerate(w), m in 1:getfield(s, :M)
        p.bins[m] >= 1 && push!(nn[m][p.bins[m]], i)
    end
    nn
end
# ---- multi-GPU: one worker per GPU (the scripts' `addprocs` + `pmap`, examples/density_SRL_lattice.jl:1-2,46) -------------------------------
# Build every worker's System with `chains` = its share and ENV["PIMC_CHAIN_OFFSET"] = first global chain, create the id on one worker, ship it
# (e.g. `remotecall_fetch`) and attach: afterwards `mea.energy[N]`, `d.dens`, `d.ndata` are reduced over all workers inside the library
# (ncclAllReduce on a side stream, per measurement block).  Every worker must issue the same read-outs in the same order.
function comm_unique_id()::Vector{UInt8}
    id = zeros(UInt8, 128)
    check(C_NULL, ccall((:pimc_comm_get_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    id
end
comm_init!(s::System, nranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(s.h, ccall((:pimc_comm_init, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), s.h, nranks, rank, id))
# complete chain state (resumes bit for bit): checkpoint(s) -> Vector{UInt8}; restore!(s2, blob) on a System built with the same arguments and objects
function checkpoint(s::System)::Vector{UInt8}
    n = Ref{Int64}(0)
    check(s.h, ccall((:pimc_state_size, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), s.h, n))
    buf = Vector{UInt8}(undef, n[])
    check(s.h, ccall((:pimc_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), s.h, buf, n[]))
    buf
end
restore!(s::System, buf::Vector{UInt8}) = check(s.h, ccall((:pimc_set_state, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), s.h, buf, length(buf)))
update_nnbins!(s::System) = check(s.h, ccall((:pimc_update_nnbins, LIB), Cint, (Ptr{Cvoid},), s.h))

# ---- updates (src/updates/helper.jl:6-52, com.jl, reshape.jl) -------------------------------------------------------------
struct UpdateStats; var::Float64; tries::Int64; tries_var::Int64; acc_window::Float64; accepted::Int64; bead_moves::Int64; end
function update_get(s::System, id::Int32; chain::Integer = 0)
    var = Ref(0.0); acc = Ref(0.0); tr = Ref{Int64}(0); tv = Ref{Int64}(0); a = Ref{Int64}(0); bm = Ref{Int64}(0)
    check(s.h, ccall((:pimc_update_get, LIB), Cint,
                     (Ptr{Cvoid}, Int32, Int32, Ref{Float64}, Ref{Int64}, Ref{Int64}, Ref{Float64}, Ref{Int64}, Ref{Int64}),
                     s.h, id, chain, var, tr, tv, acc, a, bm))
    return UpdateStats(var[], tr[], tv[], acc[], a[], bm[])
end
struct QueueView; s::System; id::Int32; end                          # stands in for Counter.queue
struct Counter; s::System; id::Int32; which::Symbol; end
struct Step; s::System; id::Int32; end
struct NumbOfSlices; s::System; id::Int32; end
Base.getproperty(c::Counter, f::Symbol) = f === :tries ? (st = update_get(getfield(c, :s), getfield(c, :id)); getfield(c, :which) === :counter ? st.tries : st.tries_var) :
                                          f === :queue ? QueueView(getfield(c, :s), getfield(c, :id)) : getfield(c, f)
Base.getproperty(v::Step, f::Symbol) = f === :size ? update_get(getfield(v, :s), getfield(v, :id)).var : getfield(v, f)
Base.getproperty(v::NumbOfSlices, f::Symbol) = f === :m ? Int64(update_get(getfield(v, :s), getfield(v, :id)).var) : getfield(v, f)
acceptance(q::QueueView)::Float64 = update_get(q.s, q.id).acc_window  # src/simulation.jl:1

function new_update(s::System, kind::Integer, var0, vmin, vmax, minacc, maxacc, adj, range)
    id = Ref{Int32}(0)
    check(s.h, ccall((:pimc_update_create, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Ref{Int32}), s.h, kind, var0, id))
    check(s.h, ccall((:pimc_update_configure, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Float64, Int64, Int64),
                     s.h, id[], vmin, vmax, minacc, maxacc, adj, range))
    return id[]
end
for (T, kind, slices) in ((:ReshapeLinear, 0, true), (:ReshapeSwapLinear, 1, true), (:SingleCenterOfMass, 2, false), (:PolymerCenterOfMass, 3, false))
    VarT = slices ? :NumbOfSlices : :Step
    @eval struct $T <: UpdateC
        s::System; id::Int32; counter::Counter; var::$VarT; counter_var::Counter
    end
    if slices
        @eval function $T(s::System, slices::Int64; minslices = 2, maxslices = s.M - 2, minacc = 0.6, maxacc = 0.8, adj = 10, range = 10_000)
            id = new_update(s, $kind, min(slices, maxslices), minslices, maxslices, minacc, maxacc, adj, range)
            $T(s, id, Counter(s, id, :counter), NumbOfSlices(s, id), Counter(s, id, :counter_var))
        end
    else
        @eval function $T(s::System, step::Float64; minstep = 1e-1, maxstep = s.L / 2, minacc = 0.4, maxacc = 0.6, adj = 10, range = 10_000)
            id = new_update(s, $kind, step, minstep, maxstep, minacc, maxacc, adj, range)
            $T(s, id, Counter(s, id, :counter), Step(s, id), Counter(s, id, :counter_var))
        end
    end
    # functor call f(s)::Bool (src/simulation.jl:20): one faithful iteration of this update on every replica
    @eval function (u::$T)(s::System)::Bool
        before = update_get(s, u.id).accepted
        run!(s, 1, [(1, u)]; sched = Int32(0))
        return update_get(s, u.id).accepted > before
    end
end
apply!(s::System, f::UpdateC) = (f(s); nothing)                        # src/simulation.jl:19-27 (bookkeeping happens on the device)

# ---- measurements (src/measurement.jl) -------------------------------------------------------------------------------------
struct Energy <: ZMeasurement
    s::System; id::Int32; n::Int64
    function Energy(s::System, n = 20_000)
        id = Ref{Int32}(0)
        check(s.h, ccall((:pimc_energy_create, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Int32}), s.h, n, id))
        new(s, id[], n)
    end
end
function energy_series(e::Energy)
    s = getfield(e, :s); n = getfield(e, :n)
    E = Vector{Float64}(undef, n); Ev = Vector{Float64}(undef, n); cnt = Ref{Int64}(0)
    check(s.h, ccall((:pimc_energy_read, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Int64}),
                     s.h, getfield(e, :id), -1, E, Ev, n, cnt))
    m = min(cnt[], n)
    pad(x) = VectorMissing{Float64}(vcat(x[1:m], fill(missing, n - m)))   # same shape as the reference's pre-sized vectors
    return pad(E), pad(Ev)
end
function Base.getproperty(e::Energy, f::Symbol)
    f === :energy && return Dict{Int64,VectorMissing{Float64}}(getfield(getfield(e, :s), :N) => energy_series(e)[1])
    f === :energy_virial && return Dict{Int64,VectorMissing{Float64}}(getfield(getfield(e, :s), :N) => energy_series(e)[2])
    return getfield(e, f)
end
struct Density <: ZMeasurement
    s::System; id::Int32; nbins::Int64; bin::Float64
    function Density(s::System; nbins = 500)
        id = Ref{Int32}(0)
        check(s.h, ccall((:pimc_density_create, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Int32}), s.h, nbins, id))
        new(s, id[], nbins, (2 * s.L) / nbins)
    end
end
function Base.getproperty(d::Density, f::Symbol)
    if f === :dens || f === :ndata
        s = getfield(d, :s); nb = getfield(d, :nbins)
        dens = zeros(Float64, ntuple(i -> nb, s.dim)); nd = Ref{Int64}(0); b = Ref(0.0)
        check(s.h, ccall((:pimc_density_read, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Int64}, Ref{Float64}), s.h, getfield(d, :id), dens, nd, b))
        return f === :dens ? dens : nd[]
    end
    return getfield(d, f)
end
(d::Density)(s::System) = check(s.h, ccall((:pimc_density_measure, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, d.id))

# ---- the estimators src/measurement.jl:125-127 lists as TODO, in the style of the functors above ---------------------------
struct PairCorrelation <: ZMeasurement                                  # `#TODO radial distribution`
    s::System; id::Int32; nbins::Int64; rmax::Float64; bin::Float64
    function PairCorrelation(s::System; nbins = 200, rmax = s.L)
        id = Ref{Int32}(0)
        check(s.h, ccall((:pimc_paircorr_create, LIB), Cint, (Ptr{Cvoid}, Int64, Float64, Ref{Int32}), s.h, nbins, rmax, id))
        new(s, id[], nbins, rmax, rmax / nbins)
    end
end
function Base.getproperty(g::PairCorrelation, f::Symbol)
    if f === :hist || f === :ndata || f === :g
        s = getfield(g, :s); nb = getfield(g, :nbins); bin = getfield(g, :bin)
        hist = zeros(Float64, nb); nd = Ref{Int64}(0); b = Ref(0.0)
        check(s.h, ccall((:pimc_paircorr_read, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Int64}, Ref{Float64}), s.h, getfield(g, :id), hist, nd, b))
        f === :hist && return hist
        f === :ndata && return nd[]
        edges = (0:nb) .* bin
        shell = s.dim == 2 ? π .* (edges[2:end] .^ 2 .- edges[1:end-1] .^ 2) : fill(2 * bin, nb)
        return hist ./ max.(nd[] * s.N * (s.N - 1) / 2 .* shell ./ s.vol, 1e-300)
    end
    return getfield(g, f)
end
(g::PairCorrelation)(s::System) = check(s.h, ccall((:pimc_paircorr_measure, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, g.id))

struct Winding <: ZMeasurement                                          # `#TODO Superfluid Fraction`
    s::System; id::Int32; n::Int64
    function Winding(s::System, n = 20_000)
        id = Ref{Int32}(0)
        check(s.h, ccall((:pimc_winding_create, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Int32}), s.h, n, id))
        new(s, id[], n)
    end
end
function winding_W2(w::Winding)::Vector{Float64}                        # chain-mean of W^2 per measurement
    s = w.s; cnt = Ref{Int64}(0); out = Vector{Float64}(undef, w.n)
    check(s.h, ccall((:pimc_winding_read, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64, Ref{Int64}), s.h, w.id, -1, out, w.n, cnt))
    return out[1:min(cnt[], w.n)]
end
# rho_s / rho = <W^2> (2L)^2 / (2 dim lambda beta N)  (Pollock & Ceperley, PRB 36, 8343)
superfluid_fraction(w::Winding) = (W2 = winding_W2(w); isempty(W2) ? NaN : sum(W2) / length(W2) * (2 * w.s.L)^2 / (2 * w.s.dim * w.s.λ * w.s.β * w.s.N))

struct StructureFactor <: ZMeasurement                                  # `#TODO Compressibilty`
    s::System; id::Int32; kmax::Int32
    function StructureFactor(s::System; kmax = 4)
        id = Ref{Int32}(0)
        check(s.h, ccall((:pimc_structure_create, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}), s.h, kmax, id))
        new(s, id[], kmax)
    end
end
function Base.getproperty(k::StructureFactor, f::Symbol)
    if f === :S || f === :ndata
        s = getfield(k, :s); km = Int(getfield(k, :kmax))
        sums = zeros(Float64, 2km + 1, km + 1); nd = Ref{Int64}(0); kk = Ref{Int32}(0)      # column-major: sums[b + km + 1, a + 1]
        check(s.h, ccall((:pimc_structure_read, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ref{Int64}, Ref{Int32}), s.h, getfield(k, :id), sums, nd, kk))
        return f === :ndata ? nd[] : sums ./ max(1, nd[] * s.N)      # S(k) at k = (π / L)(a, b)
    end
    return getfield(k, f)
end
(k::StructureFactor)(s::System) = check(s.h, ccall((:pimc_structure_measure, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, k.id))
function compressibility(k::StructureFactor)::Float64                   # κ_T = β S(k_min) / ρ on the smallest shell of the box
    κ = Ref(0.0); s0 = Ref(0.0)
    check(k.s.h, ccall((:pimc_compressibility, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}, Ref{Float64}), k.s.h, k.id, κ, s0))
    return κ[]
end

struct CMeasurements                                                    # pimc_measurements (include/pimc_b200.h)
    energy_ids::Ptr{Int32}; nenergy::Int32; density_ids::Ptr{Int32}; ndensity::Int32
    paircorr_ids::Ptr{Int32}; npaircorr::Int32; winding_ids::Ptr{Int32}; nwinding::Int32
    structure_ids::Ptr{Int32}; nstructure::Int32
end

# ---- run! (src/simulation.jl:29-42) ----------------------------------------------------------------------------------------
function run!(s::System, n::Int64, updates; Zmeasurements = ZMeasurement[], sched::Int32 = s.sched)::Nothing
    ids = Int32[u.id for (_, u) in updates]; every = Int64[e for (e, _) in updates]
    en = Int32[m.id for m in Zmeasurements if m isa Energy]; de = Int32[m.id for m in Zmeasurements if m isa Density]
    pc = Int32[m.id for m in Zmeasurements if m isa PairCorrelation]; wi = Int32[m.id for m in Zmeasurements if m isa Winding]
    sk = Int32[m.id for m in Zmeasurements if m isa StructureFactor]
    st = RunStats()
    if isempty(pc) && isempty(wi) && isempty(sk)
        GC.@preserve ids every en de check(s.h, ccall((:pimc_run, LIB), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int64}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Int32, Int32, Ref{RunStats}),
            s.h, n, ids, every, length(ids), en, length(en), de, length(de), sched, st))
    else
        GC.@preserve ids every en de pc wi sk begin
            z = CMeasurements(pointer(en), length(en), pointer(de), length(de), pointer(pc), length(pc), pointer(wi), length(wi), pointer(sk), length(sk))
            check(s.h, ccall((:pimc_run_ex, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int64}, Int32, Ref{CMeasurements}, Int32, Ref{RunStats}),
                             s.h, n, ids, every, length(ids), z, sched, st))
        end
    end
    nothing
end

# ---- debug exports (src/Pimc.jl:18) ----------------------------------------------------------------------------------------
function distance(x1::Float64, x2::Float64, L::Float64)::Float64      # src/propagator.jl:6-9, evaluated on the device
    out = Ref(0.0)
    check(C_NULL, ccall((:pimc_distance, LIB), Cint, (Int64, Ref{Float64}, Ref{Float64}, Float64, Ref{Float64}), 1, x1, x2, L, out))
    out[]
end
lnV(r1::Coord, r2::Coord, τ::Float64, V::Function)::Float64 = -0.5 * τ * (V(r1) + V(r2))   # src/propagator.jl:26-28 (host closure)
bin(r::AbstractVector{Float64}, nbins::Int64, L::Float64)::Int64 = (ib = floor.(Int64, (r .+ L) ./ (2 * L / nbins)); length(ib) == 2 ? ib[1] + nbins * ib[2] + 1 : ib[1] + 1)
"""levy!(r′, τ, L, λ) (src/updates/helper.jl:118-139): Gaussians drawn by Julia's randn, bridge arithmetic on the device."""
function levy!(r′::Matrix{Float64}, τ::Float64, L::Float64, λ::Float64)
    rows, dim = size(r′)
    xi = permutedims(randn(rows - 2, dim))                            # (rows-2) x dim row-major for the C ABI
    check(C_NULL, ccall((:pimc_levy_bridge, LIB), Cint, (Ptr{Float64}, Int32, Int32, Float64, Float64, Float64, Ptr{Float64}, Int64),
                        r′, rows, dim, τ, L, λ, xi, 1))
    r′
end
function subcycle(p::Vector{Worldline}, n::Int64)::Tuple{Int64,Vector{Int64}}   # src/updates/helper.jl:64-85
    cycle = Int64[n]; i = n
    while p[i].next != 0 && p[i].next != n && length(cycle) <= length(p)
        i = p[i].next; push!(cycle, i)
    end
    return length(cycle), cycle
end
pcycle(j::Int64, pol::Vector{Int64}, Npol::Int64, M::Int64)::Int64 = pol[mod1(1 + floor(Int64, (j - 1) / M), Npol)]  # helper.jl:113-115
# ---- pair propagator (src/propagator.jl:34-89): table built by the library's host code, csrc/pimc_propint.cu --------------
"""What `build_prop_int` returns: callable like the reference's closure `prop_int(r1_rel, r2_rel, τ)`, and carrying the sampled
term table (`terms` of src/propagator.jl:80) that `System(...; interactions = true, propint)` hands to the device."""
struct PropInt <: Function
    tab::Matrix{Float64}; lo::Float64; hi::Float64
end
function build_prop_int(L::Float64, g0::Float64, τ::Float64; Δ::Integer = 600)::PropInt
    tab = Matrix{Float64}(undef, Δ, Δ); lo = Ref(0.0); hi = Ref(0.0)
    check(C_NULL, ccall((:pimc_build_prop_table, LIB), Cint, (Float64, Float64, Float64, Int32, Ptr{Float64}, Ref{Float64}, Ref{Float64}),
                        L, g0, τ, Δ, tab, lo, hi))
    PropInt(tab, lo[], hi[])
end
function (p::PropInt)(r1_rel::AbstractVector{Float64}, r2_rel::AbstractVector{Float64}, τ::Float64)::Float64
    out = Ref(0.0); a = Vector{Float64}(r1_rel); b = Vector{Float64}(r2_rel)
    check(C_NULL, ccall((:pimc_prop_int, LIB), Cint, (Ptr{Float64}, Int32, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ref{Float64}),
                        p.tab, size(p.tab, 1), p.lo, p.hi, a, b, length(a), τ, out))
    out[]
end
function determine_nnrange(p::PropInt, τ::Float64, a::Float64, b::Float64)::Float64   # src/system.jl:10-15
    out = Ref(0.0)
    rc = ccall((:pimc_determine_nnrange, LIB), Cint, (Ptr{Float64}, Int32, Float64, Float64, Float64, Float64, Float64, Ref{Float64}),
               p.tab, size(p.tab, 1), p.lo, p.hi, τ, a, b, out)
    rc == 0 || error("determine_nnrange: no sign change of propint - 0.999 on (r_min, b)")
    out[]
end
end # module
