# SemiClassicalMCB200.jl -- the reference-side binding a maintainer would add: keeps the exported names of
# SemiClassicalMC.jl (src/SemiClassicalMC.jl:19-87) for the Metropolis / annealing path and forwards them to
# libscmc_b200.so (include/scmc.h) with `ccall`.  NOT executed in this repo's CI: the build image has no Julia.
# Geometry still comes from LatticePhysics.jl exactly as in src/Configuration.jl:37-48; only tables cross the ABI.
module SemiClassicalMCB200

using LinearAlgebra
using LatticePhysics

const LIB = get(ENV, "SCMC_LIB", joinpath(@__DIR__, "..", "semiclassicalmc.jl_b200", "libscmc_b200.so"))
const NOBS = 21

check(st::Int32) = st == 0 || error("libscmc_b200: " * unsafe_string(ccall((:scmc_last_error, LIB), Cstring, ())))

# generators exactly as the reference builds them (src/Generators.jl); any package providing getGenerators works
import SemiClassicalMC: getGenerators

mutable struct Configuration
    handle            :: Ptr{Cvoid}
    latticename       :: String
    L                 :: Int64
    nSites            :: Int64
    nReplicas         :: Int64
    d                 :: Int64
    positions         :: Vector{Vector{Float64}}
    basis             :: Vector{Vector{Float64}}
    interactionSites  :: Vector{Vector{Int64}}
    interactionLabels :: Vector{Vector{Int64}}
    seed              :: UInt64
    sweep             :: UInt64
end

Base.length(cfg::Configuration) = cfg.nSites

# C order [a][j][k] from a Vector of d x d matrices (Julia is column-major: permute before handing over)
parts(gen) = (Float64[real(gen[a][j, k]) for k in 1:size(gen[1], 2), j in 1:size(gen[1], 1), a in 1:16][:],
              Float64[imag(gen[a][j, k]) for k in 1:size(gen[1], 2), j in 1:size(gen[1], 1), a in 1:16][:])

# Configuration(latticename, L, uc, interactions, onsiteInteraction, n)  (src/Configuration.jl:27-34) + replica keywords
function Configuration(latticename::String, L::Int64, uc, interactions::Vector{Matrix{Float64}}, onsiteInteraction::Vector{Float64},
                       n::Int64; replicas::Int = 1, precision::Int = 32, device::Int = 0, replica_offset::Int = 0,
                       seed::Integer = abs(rand(Int)))
    l = getLatticePeriodic(uc, L)
    N = length(l.sites)
    orgBonds = organizedBondsFrom(l)
    iSites = [to.(b) for b in orgBonds]
    iLabels = [label.(b) for b in orgBonds]
    z = maximum(length.(iSites))
    nbr = fill(Int32(-1), z, N); lab = zeros(Int32, z, N)          # column-major [z, N] == C order [N][z]
    for i in 1:N, k in eachindex(iSites[i])
        nbr[k, i] = iSites[i][k] - 1
        lab[k, i] = iLabels[i][k] - 1
    end
    J = Float64[interactions[l][a, b] for b in 1:16, a in 1:16, l in eachindex(interactions)][:]   # C order [label][a][b]
    gen = getGenerators(n); gensq = gen .^ 2
    gre, gim = parts(gen); qre, qim = parts(gensq)
    d = size(gen[1], 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:scmc_create, LIB), Int32,
                (Ref{Ptr{Cvoid}}, Int32, Int32, Int64, Int32, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                h, device, precision, N, d, replicas, replica_offset, z, nbr, lab, length(interactions), J, onsiteInteraction,
                gre, gim, qre, qim, C_NULL))
    cfg = Configuration(h[], latticename, L, N, replicas, d, point.(l.sites), point.(uc.sites), iSites, iLabels, UInt64(seed), UInt64(0))
    check(ccall((:scmc_randomize, LIB), Int32, (Ptr{Cvoid}, UInt64), cfg.handle, cfg.seed))     # src/Configuration.jl:62
    finalizer(c -> ccall((:scmc_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), cfg)
    return cfg
end

# getEnergy(cfg) (src/Configuration.jl:236-253): replica 1
function getEnergy(cfg::Configuration)
    E = zeros(cfg.nReplicas)
    check(ccall((:scmc_energy, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), cfg.handle, E))
    return E[1]
end

# getSpinExpectation(cfg) (src/Configuration.jl:128-134)
function getSpinExpectation(cfg::Configuration; replica::Int = 1)
    T = zeros(16, cfg.nSites)
    check(ccall((:scmc_get_T, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), cfg.handle, replica - 1, T, C_NULL))
    return [T[:, i] for i in 1:cfg.nSites]
end

# localUpdate!(cfg, E, accepted, β, i, σ) (src/MonteCarlo.jl:263-286): randoms drawn here, update on the device
function localUpdate!(cfg::Configuration, E::Float64, accepted_updates::Int64, β::Float64, i::Int64, σ::Float64; replica::Int = 1)
    xi = randn(2 * cfg.d); u = rand()
    dE = Ref{Float64}(0.0); acc = Ref{Int32}(0)
    check(ccall((:scmc_update_deterministic, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Float64, Float64, Float64, Ref{Float64}, Ref{Int32}),
                cfg.handle, replica - 1, i - 1, xi, u, β, σ, dE, acc))
    return acc[] == 1 ? (E + dE[], accepted_updates + 1) : (E, accepted_updates)
end

struct RunParams
    n_therm::Int64; n_measure::Int64; measurement_rate::Int64; hits::Int32; reserved::Int32
    seed::UInt64; sweep0::UInt64; T::Ptr{Float64}; T_i::Ptr{Float64}; sigma::Ptr{Float64}
end
mutable struct RunResults
    series::Ptr{Float64}; n_rows::Int64; mean::Ptr{Float64}; acc_rate::Ptr{Float64}; therm_energy::Ptr{Float64}
end

# run!(cfg, obs, T, T_i, N_therm, N_measure, filename; ...) (src/MonteCarlo.jl:3-18).  T / T_i / σ may be per-replica vectors.
function run!(cfg::Configuration, obs, T, T_i, N_therm::Int, N_measure::Int, filename::String; measurement_rate::Int = 10,
              checkpoint_rate::Int = 100000, report_rate::Int = 100, current_sweep::Int = 0, energy::Vector{Float64} = Float64[],
              βs::Vector{Float64} = Float64[], σ = 60.0)
    R = cfg.nReplicas
    Ts = T isa Number ? fill(Float64(T), R) : Float64.(T)
    Tis = T_i isa Number ? fill(Float64(T_i), R) : Float64.(T_i)
    σs = σ isa Number ? fill(Float64(σ), R) : Float64.(σ)
    rows = (N_therm + N_measure) ÷ measurement_rate - max(N_therm, current_sweep) ÷ measurement_rate
    series = zeros(NOBS, R, max(rows, 1)); means = zeros(NOBS, R); acc = zeros(R)
    GC.@preserve Ts Tis σs series means acc begin
        p = RunParams(N_therm, N_measure, measurement_rate, 2, 0, cfg.seed, current_sweep, pointer(Ts), pointer(Tis), pointer(σs))
        out = RunResults(pointer(series), 0, pointer(means), pointer(acc), C_NULL)
        check(ccall((:scmc_run_metropolis, LIB), Int32, (Ptr{Cvoid}, Ref{RunParams}, Ref{RunResults}), cfg.handle, p, out))
    end
    # series[:, r, m] is the measure! row (E/N, (E/N)^2, |m_σ|, |m_τ|, |m_στ|, m[1:16]) of replica r at measurement m:
    # push it into the BinningAnalysis accumulators of `obs` exactly as measure! does (src/Observables/ObservablesGeneric.jl:32-46)
    return (series = series, means = means, acceptance = acc, σ = σs)
end

struct AnnealParams
    T_i::Float64; T_fac::Float64; N_max::Int64; N_per_T::Int64; min_acc_per_site::Int64
    min_accrate::Float64; sigma_min::Float64; sigma0::Float64; hits::Int32; reserved::Int32; seed::UInt64
end
mutable struct AnnealResults
    E::Ptr{Float64}; beta::Ptr{Float64}; sigma::Ptr{Float64}; sweeps::Ptr{Int64}
end

# runAnnealing!(cfg, T_i, filename; ...) (src/MonteCarlo.jl:112-129)
function runAnnealing!(cfg::Configuration, T_i, filename::String; T_fac = 0.98, N_max::Int = 10000000, N_o::Int = 0, N_per_T::Int = 10000,
                       min_acc_per_site::Int = 100, min_accrate = 1e-5, checkpoint_rate::Int = 100000, σ_min::Float64 = 0.05,
                       verbose::Bool = false, σ::Float64 = 60.0)
    R = cfg.nReplicas
    E = zeros(R); β = zeros(R); σs = zeros(R); sweeps = zeros(Int64, R)
    GC.@preserve E β σs sweeps begin
        p = AnnealParams(T_i, T_fac, N_max, N_per_T, min_acc_per_site, min_accrate, σ_min, σ, 2, 0, cfg.seed)
        out = AnnealResults(pointer(E), pointer(β), pointer(σs), pointer(sweeps))
        check(ccall((:scmc_anneal, LIB), Int32, (Ptr{Cvoid}, Ref{AnnealParams}, Ref{AnnealResults}), cfg.handle, p, out))
        N_o > 0 && check(ccall((:scmc_optimize, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}), cfg.handle, N_o, E))
    end
    return (E = E, β = β, σ = σs, sweeps = sweeps)
end

# cfg.state as the d x N complex matrix of src/IO.jl:146-150 (f["state"], src/MonteCarlo.jl:232-234)
function getStateMatrix(cfg::Configuration; replica::Int = 1)
    re = zeros(cfg.d, cfg.nSites); im_ = zeros(cfg.d, cfg.nSites)
    check(ccall((:scmc_get_state, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}), cfg.handle, replica - 1, 1, re, im_))
    return complex.(re, im_)
end

export Configuration, getEnergy, getSpinExpectation, localUpdate!, run!, runAnnealing!, getStateMatrix
end
