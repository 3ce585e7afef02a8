# SemiClassicalMCB200.jl -- the reference-side binding a maintainer adds to run SemiClassicalMC.jl's Metropolis / annealing path on B200s:
# every on-path exported name of src/SemiClassicalMC.jl:19-87 keeps its signature and its mutation contract and forwards to
# libscmc_b200.so (include/scmc.h) with `ccall`.  NOT executed in this repo's CI (the build image has no Julia) -- INTEGRATION.md lists,
# name by name, what binds to what; the same C entry points are exercised from C (tests/abi_client.c) and Python (tests/).
#
#   using SemiClassicalMC            # geometry (LatticePhysics), generators, BinningAnalysis accumulators, HDF5 post-processing stay
#   using SemiClassicalMCB200        # Configuration / run! / runAnnealing! / launch! / measure! ... of THIS module run on the GPU
#
# Conventions: replica 1 of a configuration plays the role of the reference's single chain; `replicas = R` packs R independent chains
# (temperatures, restarts) into one launch -- then T, T_i, σ may be vectors and `obs` a vector of observables, one per replica.
# σ crosses this API in the REFERENCE's units (its randn! has variance 1/2: σ_reference = sqrt(2) σ_library, see DESIGN.md section 2).
module SemiClassicalMCB200

using LinearAlgebra
using Random
using Serialization
using HDF5
using BinningAnalysis
using LatticePhysics
import SemiClassicalMC
import SemiClassicalMC: Observables, ObservablesGeneric, ObservablesTgHbn, getGenerators, computeStructureFactor, computeStructureFactorVec,
                        getKsInBox, getmeans, getenergies, vectorToMatrix

const LIB = get(ENV, "SCMC_LIB", joinpath(@__DIR__, "..", "semiclassicalmc.jl_b200", "libscmc_b200.so"))
const NOBS = 21
const SQRT2 = sqrt(2.0)

check(st::Int32) = st == 0 || error("libscmc_b200: " * unsafe_string(ccall((:scmc_last_error, LIB), Cstring, ())))

## ------------------------------------------------------------------------------------------------ Configuration (src/Configuration.jl:2-77)
mutable struct Configuration
    handle            :: Ptr{Cvoid}
    latticename       :: String
    L                 :: Int64
    n                 :: Int64
    nSites            :: Int64
    nReplicas         :: Int64
    d                 :: Int64
    unitVectors       :: Vector{Vector{Float64}}
    positions         :: Vector{Vector{Float64}}
    basis             :: Vector{Vector{Float64}}
    basisLabels       :: Vector{Int64}
    interactionSites  :: Vector{Vector{Int64}}
    interactionLabels :: Vector{Vector{Int64}}
    interactions      :: Vector{Matrix{Float64}}
    onsiteInteraction :: Vector{Float64}
    generators        :: Vector{Matrix{ComplexF64}}
    generatorsSq      :: Vector{Matrix{ComplexF64}}
    uc                :: Any                    # the LatticePhysics unit cell the configuration was built from (kept for checkpoints)
    precision         :: Int64
    device            :: Int64
    replicaOffset     :: Int64
    seed              :: UInt64
    streamEpoch       :: Int64                  # fresh runs after the first use another stream key (no counter is ever replayed)
    runs              :: Int64
end

Base.length(cfg::Configuration) = cfg.nSites

# C order [a][j][k] from a Vector of d x d matrices (Julia is column-major: build [k, j, a])
parts(gen) = (Float64[real(gen[a][j, k]) for k in 1:size(gen[1], 2), j in 1:size(gen[1], 1), a in 1:16][:],
              Float64[imag(gen[a][j, k]) for k in 1:size(gen[1], 2), j in 1:size(gen[1], 1), a in 1:16][:])

# tables of a model as the C ABI takes them (0-based, C order); shared by Configuration and MultiConfiguration
function abiTables(uc, L::Int64, interactions::Vector{Matrix{Float64}}, n::Int64)
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
    gen = getGenerators(n); gensq = gen .^ 2                     # matrix squares, src/Configuration.jl:52
    return l, N, iSites, iLabels, z, nbr, lab, J, gen, gensq
end

# Configuration(latticename, L, uc, interactions, onsiteInteraction, n)  (src/Configuration.jl:27-34) + replica keywords
function Configuration(latticename::String, L::Int64, uc, interactions::Vector{Matrix{Float64}}, onsiteInteraction::Vector{Float64},
                       n::Int64; replicas::Int = 1, precision::Int = 32, device::Int = 0, replica_offset::Int = 0,
                       seed::Integer = abs(rand(Random.RandomDevice(), Int)))
    l, N, iSites, iLabels, z, nbr, lab, J, gen, gensq = abiTables(uc, L, interactions, n)
    gre, gim = parts(gen); qre, qim = parts(gensq)
    d = size(gen[1], 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:scmc_create, LIB), Int32,
                (Ref{Ptr{Cvoid}}, Int32, Int32, Int64, Int32, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                h, device, precision, N, d, replicas, replica_offset, z, nbr, lab, length(interactions), J, onsiteInteraction,
                gre, gim, qre, qim, C_NULL))
    cfg = Configuration(h[], latticename, L, n, N, replicas, d, [Vector{Float64}(v) for v in uc.lattice_vectors], point.(l.sites),
                        point.(uc.sites), label.(l.sites), iSites, iLabels, interactions, onsiteInteraction, gen, gensq, uc,
                        precision, device, replica_offset, UInt64(seed), 0, 0)
    check(ccall((:scmc_randomize, LIB), Int32, (Ptr{Cvoid}, UInt64), cfg.handle, cfg.seed))     # random initial state, src/Configuration.jl:62
    finalizer(c -> ccall((:scmc_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), cfg)
    return cfg
end

## ------------------------------------------------------------------------------------------------ models (src/Models.jl:2-113)
function initializeCfg(model::String, latticename::String, J::AbstractVector, L::Int, n::Int; kwargs...)
    if model == "nearest-neighbor"
        initializeCfgNN(J, latticename, L, n; kwargs...)
    elseif model == "tg-hbn"
        initializeCfgTgHbn(J, L; n = n, kwargs...)
    else
        @error "model $model not implemented"       # like the reference: logs and returns nothing (src/Models.jl:15)
    end
end

function initializeCfgNN(J::AbstractVector, latticename::String, L::Int, n::Int; kwargs...)
    if length(J) == 1
        onsite = zeros(16); interactions = [Matrix{Float64}(J[1])]
    elseif length(J) == 2
        onsite = Vector{Float64}(diag(J[1])); interactions = [Matrix{Float64}(J[2])]      # src/Models.jl:30-36
    else
        error("Need length(J) = 1 (nn interactions) or length(J) = 2 (on-site and nearest-neighbor interactions).")
    end
    uc = getUnitcell(Symbol(latticename))
    return Configuration(latticename, L, uc, interactions, onsite, n; kwargs...)
end

function initializeCfgTgHbn(J::AbstractVector, L::Int; n::Int = 2, kwargs...)
    @assert length(J) == 5 "Need J = [J1, J2, Jp1, Jp2, JH]"
    J1, J2, Jp1, Jp2, JH = J ./ 8                                        # src/Models.jl:55
    onsite = vcat([0.0, 0.0, 0.0, JH / 2], repeat([-JH / 2, 0.0, 0.0, JH / 2], 3))
    Jv1 = [J1 0 0 0; 0 J1+Jp1 Jp2 0; 0 -Jp2 J1+Jp1 0; 0 0 0 J1]
    J_nn1 = kron(Matrix(1.0I, 4, 4), Jv1); J_nn1[1, 1] = 0.0
    J_nn2 = kron(Matrix(1.0I, 4, 4), Matrix(Jv1')); J_nn2[1, 1] = 0.0
    J_nnn = diagm(vcat(0.0, fill(J2, 15)))
    uc = getUnitcell(:triangular)
    empty!(uc.bonds)
    for off in ((1, 0), (0, -1), (-1, 1)); addBond!(uc, 1, 1, 1, off, false); end      # src/Models.jl:98-110
    for off in ((-1, 0), (0, 1), (1, -1)); addBond!(uc, 1, 1, 2, off, false); end
    for off in ((1, 1), (-1, 2), (-2, 1)); addBond!(uc, 1, 1, 3, off, true); end
    return Configuration("triangular", L, uc, [J_nn1, J_nn2, J_nnn], onsite, n; kwargs...)
end

## ------------------------------------------------------------------------------------------------ accessors (src/Configuration.jl:80-134)
getPosition(cfg::Configuration, i::Int) = cfg.positions[i]
getBasis(cfg::Configuration) = cfg.basis
getInteractionSites(cfg::Configuration, i::Int) = cfg.interactionSites[i]
getInteractionLabels(cfg::Configuration, i::Int) = cfg.interactionLabels[i]
getInteraction(cfg::Configuration, label::Int) = cfg.interactions[label]
getOnsiteInteraction(cfg::Configuration) = cfg.onsiteInteraction
SemiClassicalMC.getGenerators(cfg::Configuration) = cfg.generators
getGeneratorsSq(cfg::Configuration) = cfg.generatorsSq

# cfg.state as the d x N complex matrix of src/IO.jl:146-150 (f["state"], src/MonteCarlo.jl:232-234); getState(cfg, i) is a column of it
function getStateMatrix(cfg::Configuration; replica::Int = 1)
    re = zeros(cfg.d, cfg.nSites); im_ = zeros(cfg.d, cfg.nSites)
    check(ccall((:scmc_get_state, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}), cfg.handle, replica - 1, 1, re, im_))
    return complex.(re, im_)
end
function setStateMatrix!(cfg::Configuration, psi::AbstractMatrix{<:Complex}; replica::Int = 1)
    re = Matrix{Float64}(real.(psi)); im_ = Matrix{Float64}(imag.(psi))
    check(ccall((:scmc_set_state, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}), cfg.handle, replica - 1, 1, re, im_))
    return nothing
end
getState(cfg::Configuration, i::Int64; replica::Int = 1) = getStateMatrix(cfg; replica = replica)[:, i]

function spinTables(cfg::Configuration; replica::Int = 1)
    T = zeros(16, cfg.nSites); Tsq = zeros(16, cfg.nSites)
    check(ccall((:scmc_get_T, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), cfg.handle, replica - 1, T, Tsq))
    return T, Tsq
end
getSpinExpectation(cfg::Configuration; replica::Int = 1) = (T = spinTables(cfg; replica = replica)[1]; [T[:, i] for i in 1:cfg.nSites])
getSpinSqExpectation(cfg::Configuration; replica::Int = 1) = (T = spinTables(cfg; replica = replica)[2]; [T[:, i] for i in 1:cfg.nSites])
getSpinExpectation(cfg::Configuration, i::Int64; replica::Int = 1) = getSpinExpectation(cfg; replica = replica)[i]
getSpinSqExpectation(cfg::Configuration, i::Int64; replica::Int = 1) = getSpinSqExpectation(cfg; replica = replica)[i]

# getEnergy(cfg) (src/Configuration.jl:236-253): replica 1; getEnergies: all replicas
function getEnergies(cfg::Configuration)
    E = zeros(cfg.nReplicas)
    check(ccall((:scmc_energy, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), cfg.handle, E))
    return E
end
getEnergy(cfg::Configuration) = getEnergies(cfg)[1]

# localUpdate!(cfg, E, accepted, β, i, σ) (src/MonteCarlo.jl:263-286): randoms drawn here, update on the device.  σ in reference units.
function localUpdate!(cfg::Configuration, E::Float64, accepted_updates::Int64, β::Float64, i::Int64, σ::Float64; replica::Int = 1)
    xi = randn(2 * cfg.d); u = rand()
    dE = Ref{Float64}(0.0); acc = Ref{Int32}(0)
    check(ccall((:scmc_update_deterministic, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Float64, Float64, Float64, Ref{Float64}, Ref{Int32}),
                cfg.handle, replica - 1, i - 1, xi, u, β, σ / SQRT2, dE, acc))
    return acc[] == 1 ? (E + dE[], accepted_updates + 1) : (E, accepted_updates)
end

# kernel / stream selection (include/scmc.h)
setSweepMode!(cfg::Configuration, mode::Int, batch::Int = 0) = check(ccall((:scmc_set_sweep_mode, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), cfg.handle, mode, batch))
setStreamVersion!(cfg::Configuration, v::Int) = check(ccall((:scmc_set_stream_version, LIB), Int32, (Ptr{Cvoid}, Int32), cfg.handle, v))
setLanes!(cfg::Configuration, lanes::Int) = check(ccall((:scmc_set_lanes, LIB), Int32, (Ptr{Cvoid}, Int32), cfg.handle, lanes))
"Number of violations found by the on-device self-check of colouring, neighbour tables and tile plan (0 = consistent)."
function verifyTables(cfg::Configuration)
    n = Ref{Int64}(0)
    check(ccall((:scmc_verify_tables, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), cfg.handle, n))
    return n[]
end

## ------------------------------------------------------------------------------------------------ observables (src/Observables/*.jl)
# measure! rows of all replicas, NOBS x R: E/N, (E/N)^2, |m_σ|, |m_τ|, |m_στ|, m[1:16]
function measureRows(cfg::Configuration)
    rows = zeros(NOBS, cfg.nReplicas)
    check(ccall((:scmc_measure, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), cfg.handle, rows))
    return rows
end
getMagnetization(cfg::Configuration; replica::Int = 1) = measureRows(cfg)[6:21, replica]          # src/Observables/Observables.jl:16-18

getRs(cfg::Configuration) = vcat([cfg.positions .- Ref(cfg.positions[1] .+ b) for b in cfg.basis]...)        # :44-46

function getPrefactors(v::Vector{<:Number}, basis::Vector{<:Vector{<:Number}})                             # :83-88
    A = [dot(e1, e2) for e1 in basis, e2 in basis]
    b = [dot(e, v) for e in basis]
    return round.((A \ b), digits = 5)
end
function periodicDistance(i::Int64, j::Int64, cfg::Configuration)                                          # :65-80
    db = cfg.basis[cfg.basisLabels[i]] .- cfg.basis[cfg.basisLabels[j]]
    r = cfg.positions[i] - cfg.positions[j] .- db
    ns = getPrefactors(r, cfg.unitVectors)
    for k in eachindex(ns)
        ns[k] = ((ns[k] % cfg.L) + cfg.L) % cfg.L
    end
    return sum(ns .* cfg.unitVectors) .+ db
end
function getProject(cfg::Configuration)                                                                    # :49-62
    rs = getRs(cfg)
    project = zeros(Int64, length(cfg), length(cfg))
    for i in 1:length(cfg), j in 1:length(cfg)
        project[i, j] = findfirst(isapprox.(Ref(periodicDistance(i, j, cfg)), rs; atol = 1e-10))
    end
    return project
end
# getCorrelations(cfg, project) (:22-41): pair sums on the device, 16 x (nBasis N)
function getCorrelations(cfg::Configuration, project::Matrix{Int64}; replica::Int = 1)
    n_rs = length(cfg.basis) * cfg.nSites
    p0 = Matrix{Int32}(permutedims(project) .- 1)          # C order [i][j], 0-based
    C = zeros(n_rs, 16)                                    # C order [16][n_rs] = Julia (n_rs, 16)
    check(ccall((:scmc_correlations, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Int64, Ptr{Float64}), cfg.handle, replica - 1, p0, n_rs, C))
    return permutedims(C)
end
# S^a(k) of every replica for arbitrary momenta, straight from the spin vectors (equals computeStructureFactorVec of the correlations)
function structureFactor(cfg::Configuration, ks::Vector{<:Vector{<:Number}})
    dim = length(ks[1]); nk = length(ks)
    kmat = Float64[ks[m][c] for c in 1:dim, m in 1:nk]; pmat = Float64[cfg.positions[i][c] for c in 1:dim, i in 1:cfg.nSites]
    out = zeros(nk, 16, cfg.nReplicas)
    check(ccall((:scmc_structure_factor, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64, Ptr{Float64}),
                cfg.handle, kmat, pmat, nk, dim, Float64(length(cfg.basis) * cfg.nSites), out))
    return out
end
# allowed momenta of the periodic cluster k = (m1 b1 + m2 b2) / L: separable DFT over the cell indices (all momenta of a cell at once);
# ms: 2 x nk integer momentum indices, comp_mask != 0 sums the selected components on the device (bit a-1 set: component a)
function structureFactorGrid(cfg::Configuration, ms::Matrix{Int32}, ks::Vector{<:Vector{<:Number}}; comp_mask::UInt32 = UInt32(0))
    nb = length(cfg.basis); nk = size(ms, 2)
    cells = Matrix{Int32}(undef, 2, cfg.nSites)
    for i in 1:cfg.nSites
        cells[:, i] = Int32.(round.(getPrefactors(cfg.positions[i] .- cfg.basis[cfg.basisLabels[i]], cfg.unitVectors)))
    end
    basis0 = Int32.(cfg.basisLabels .- 1)
    phase = Float64[f(dot(ks[m], cfg.basis[b])) for f in (cos, sin), b in 1:nb, m in 1:nk]      # C order [nk][nb][2]
    out = comp_mask == 0 ? zeros(nk, 16, cfg.nReplicas) : zeros(nk, cfg.nReplicas)
    check(ccall((:scmc_structure_factor_grid, LIB), Int32,
                (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64}, Float64, UInt32, Ptr{Float64}),
                cfg.handle, cfg.L, cfg.L, nb, cells, basis0, nk, mod.(ms, Int32(cfg.L)), phase, Float64(nb * cfg.nSites), comp_mask, out))
    return out
end
# TG/h-BN reductions (src/Observables/ObservablesTgHbn.jl:211-249)
function getSublatticeMagnetization(cfg::Configuration, siteLabels::Vector{Int64}; replica::Int = 1)
    labels = sort(unique(siteLabels)); lab0 = Int32[findfirst(==(s), labels) - 1 for s in siteLabels]
    out = zeros(16, length(labels), cfg.nReplicas)
    check(ccall((:scmc_sublattice_magnetization, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float64}), cfg.handle, lab0, length(labels), out))
    return [out[:, s, replica] for s in eachindex(labels)]
end
function getChirality(cfg::Configuration; replica::Int = 1)
    tri = Matrix{Int32}(undef, 4, cfg.nSites)
    for i in 1:cfg.nSites
        tri[:, i] = Int32.(getInteractionSites(cfg, i)[[1, 5, 4, 2]] .- 1)
    end
    out = zeros(3, cfg.nReplicas)
    check(ccall((:scmc_chirality, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}), cfg.handle, tri, out))
    return out[:, replica]
end

# initializeObservables(obstype, cfg) (src/Observables/Observables.jl:9-11): the reference's accumulator structs, built from this cfg
function ObservablesGeneric(cfg::Configuration)
    return ObservablesGeneric(ErrorPropagator(Float64), LogBinner(Float64), LogBinner(Float64), LogBinner(Float64), LogBinner(zeros(Float64, 16)),
                              LogBinner(zeros(Float64, 16, length(cfg.basis) * length(cfg))), getRs(cfg), getProject(cfg))
end
initializeObservables(obstype::Type{T}, cfg::Configuration) where {T <: Observables} = obstype(cfg)

# push one measure! row (src/Observables/ObservablesGeneric.jl:32-46); E/N and its square come re-summed from the device
function pushRow!(obs::ObservablesGeneric, row::AbstractVector{Float64})
    push!(obs.energy, row[1], row[2])
    push!(obs.σ, row[3]); push!(obs.τ, row[4]); push!(obs.στ, row[5])
    push!(obs.magnetizationVec, Vector{Float64}(row[6:21]))
    return nothing
end
function measure!(obs::ObservablesGeneric, cfg::Configuration, E::Number = NaN; replica::Int = 1)
    pushRow!(obs, measureRows(cfg)[:, replica])
    push!(obs.correlations, getCorrelations(cfg, obs.project; replica = replica))
    return nothing
end

function saveMeans!(filename::String, obs::ObservablesGeneric, cfg::Configuration, β::Float64)      # ObservablesGeneric.jl:49-72, same datasets
    h5open(filename, "cw") do file
        file["means/energy/mean"] = means(obs.energy)[1]
        file["means/energy/error"] = std_errors(obs.energy)[1]
        file["means/σ/mean"] = mean(obs.σ); file["means/σ/error"] = std_error(obs.σ)
        file["means/τ/mean"] = mean(obs.τ); file["means/τ/error"] = std_error(obs.τ)
        file["means/στ/mean"] = mean(obs.στ); file["means/στ/error"] = std_error(obs.στ)
        file["means/magnetizationVec/mean"] = mean(obs.magnetizationVec)
        file["means/magnetizationVec/error"] = std_error(obs.magnetizationVec)
        file["means/correlations/mean"] = mean(obs.correlations)
        file["means/correlations/error"] = std_error(obs.correlations)
        c(e) = β * β * (e[2] - e[1] * e[1]) * length(cfg)
        ∇c(e) = [-2.0 * β * β * e[1] * length(cfg), β * β * length(cfg)]
        lvl = BinningAnalysis._reliable_level(obs.energy)
        file["means/heat/mean"] = mean(obs.energy, c)
        file["means/heat/error"] = sqrt(abs(var(obs.energy, ∇c, lvl)) / obs.energy.count[lvl])
        file["rs"] = vectorToMatrix(obs.rs)
    end
    return nothing
end

## ------------------------------------------------------------------------------------------------ IO (src/IO.jl:5-61)
# what is serialised in place of the reference's cfg: everything needed to rebuild the handle and continue the SAME chain
struct ConfigurationSnapshot
    latticename::String; L::Int64; n::Int64; uc::Any; interactions::Vector{Matrix{Float64}}; onsiteInteraction::Vector{Float64}
    replicas::Int64; precision::Int64; device::Int64; replicaOffset::Int64; seed::UInt64; streamEpoch::Int64
    states::Vector{Matrix{ComplexF64}}
end
snapshot(cfg::Configuration) = ConfigurationSnapshot(cfg.latticename, cfg.L, cfg.n, cfg.uc, cfg.interactions, cfg.onsiteInteraction, cfg.nReplicas,
                                                     cfg.precision, cfg.device, cfg.replicaOffset, cfg.seed, cfg.streamEpoch,
                                                     [getStateMatrix(cfg; replica = r) for r in 1:cfg.nReplicas])
function restore(s::ConfigurationSnapshot)
    cfg = Configuration(s.latticename, s.L, s.uc, s.interactions, s.onsiteInteraction, s.n; replicas = s.replicas, precision = s.precision,
                        device = s.device, replica_offset = s.replicaOffset, seed = s.seed)
    for r in 1:s.replicas
        setStateMatrix!(cfg, s.states[r]; replica = r)
    end
    cfg.streamEpoch = s.streamEpoch; cfg.runs = 1
    return cfg
end

function checkpoint!(filename, cfg::Configuration, obs, sweep::Int64, βs::Vector{Float64}, energy::Vector{Float64}, σ::Float64)
    h5open(filename, "w") do file
        io = IOBuffer()
        serialize(io, snapshot(cfg)); file["cfg"] = take!(io)
        serialize(io, obs); file["obs"] = take!(io)
        file["current_sweep"] = sweep; file["energy"] = energy; file["βs"] = βs; file["σ"] = σ
        file["state"] = getStateMatrix(cfg)                  # readable without Julia deserialisation (and by the Python mirror)
    end
    return nothing
end
function checkpoint!(filename, cfg::Configuration, sweep::Int64, βs::Vector{Float64}, energy::Vector{Float64}, σ::Float64)      # annealing, IO.jl:37-49
    h5open(filename, "w") do file
        io = IOBuffer()
        serialize(io, snapshot(cfg)); file["cfg"] = take!(io)
        file["current_sweep"] = sweep; file["energy"] = energy; file["βs"] = βs; file["σ"] = σ
    end
    return nothing
end
function readCheckpoint(filename, ::Type{Obs}) where {Obs}
    h5open(filename) do file
        cfg = restore(deserialize(IOBuffer(read(file["cfg"]))))
        obs = deserialize(IOBuffer(read(file["obs"])))
        return cfg, obs, read(file["current_sweep"]), read(file["energy"]), read(file["βs"]), read(file["σ"])
    end
end
function readCheckpointSA(filename)
    h5open(filename) do file
        cfg = restore(deserialize(IOBuffer(read(file["cfg"]))))
        return cfg, read(file["current_sweep"]), read(file["energy"]), read(file["βs"]), read(file["σ"])
    end
end

## ------------------------------------------------------------------------------------------------ Monte Carlo (src/MonteCarlo.jl)
struct RunParams
    n_therm::Int64; n_measure::Int64; measurement_rate::Int64; hits::Int32; stop_sweep::Int32
    seed::UInt64; sweep0::UInt64; T::Ptr{Float64}; T_i::Ptr{Float64}; sigma::Ptr{Float64}
end
mutable struct RunResults
    series::Ptr{Float64}; n_rows::Int64; mean::Ptr{Float64}; acc_rate::Ptr{Float64}; therm_energy::Ptr{Float64}
end

perReplica(x, R) = x isa Number ? fill(Float64(x), R) : Vector{Float64}(x)
function newEpoch!(cfg::Configuration, resumed::Bool)
    if !resumed && cfg.runs > 0
        cfg.streamEpoch += 1
    end
    cfg.runs += 1
    return cfg.seed + 0x9E3779B97F4A7C15 * UInt64(cfg.streamEpoch)        # wraps mod 2^64 like the Python mirror's stream_seed
end
obsOf(obs, r) = obs isa AbstractVector ? obs[r] : (r == 1 ? obs : nothing)

# run!(cfg, obs, T, T_i, N_therm, N_measure, filename; ...) (src/MonteCarlo.jl:3-108).  Mutates cfg (device state), obs (one push per
# measurement, :82-83), energy and βs (one push per sigma-adaptation and per measurement point, :48-49, 84-85; replica 1), writes
# checkpoint! + saveMeans! every checkpoint_rate sweeps and at the end (:56-59, 88-91, 103-104), prints the reference's progress lines.
function run!(cfg::Configuration, obs, T, T_i, N_therm::Int, N_measure::Int, filename::String; measurement_rate::Int = 10,
              checkpoint_rate::Int = 100000, report_rate::Int = 100, current_sweep::Int = 0, energy::Vector{Float64} = Float64[],
              βs::Vector{Float64} = Float64[], σ = 60.0, hits::Int = 2, replica_filenames = nothing)::Nothing
    R = cfg.nReplicas
    Ts = perReplica(T, R); Tis = perReplica(T_i, R); σs = perReplica(σ, R) ./ SQRT2
    rate = measurement_rate
    total = N_therm + N_measure
    N_anneal = ceil(Int, 0.75 * N_therm)
    seed = newEpoch!(cfg, current_sweep > 0)
    cp = cld(max(checkpoint_rate, 1), rate) * rate               # a chunk never cuts a sigma-adaptation / measurement interval
    println("$filename: Starting thermalization sweeps"); flush(stdout)
    s = current_sweep
    while s < total
        stop = min(total, (s ÷ cp + 1) * cp)
        therm_end = min(N_therm, stop)
        n_log = max(0, therm_end ÷ rate - s ÷ rate)
        rows_max = stop > N_therm ? max(0, stop ÷ rate - max(N_therm, s) ÷ rate) : 0
        series = zeros(NOBS, R, max(rows_max, 1)); means_ = zeros(NOBS, R); acc = zeros(R); therm = zeros(R, max(n_log, 1))
        out = RunResults(C_NULL, 0, C_NULL, C_NULL, C_NULL)
        GC.@preserve Ts Tis σs series means_ acc therm begin
            p = RunParams(N_therm, N_measure, rate, hits, stop >= total ? 0 : stop, seed, s, pointer(Ts), pointer(Tis), pointer(σs))
            out = RunResults(pointer(series), 0, pointer(means_), pointer(acc), pointer(therm))
            check(ccall((:scmc_run_metropolis, LIB), Int32, (Ptr{Cvoid}, Ref{RunParams}, Ref{RunResults}), cfg.handle, p, out))
        end
        # thermalisation log of this chunk (push!(βs, β); push!(energy, E / N), :48-49)
        k = 0
        for q in (s + 1):therm_end
            q % rate == 0 || continue
            k += 1
            Tq = (q <= N_anneal && N_anneal > 1) ? Tis[1] * (Ts[1] / Tis[1])^((q - 1) / (N_anneal - 1)) : Ts[1]
            push!(βs, 1 / Tq); push!(energy, therm[1, k])
        end
        # measurement rows of this chunk (measure!(obs, cfg, E); push!(energy, E / N); push!(βs, β), :82-85)
        for m in 1:out.n_rows
            for r in 1:R
                o = obsOf(obs, r)
                o === nothing || pushRow!(o, series[:, r, m])
            end
            push!(energy, series[1, 1, m]); push!(βs, 1 / Ts[1])
        end
        if out.n_rows > 0 && any(o -> o isa ObservablesGeneric, obs isa AbstractVector ? obs : [obs])
            # the reference measures getCorrelations at every measure!; the packed driver keeps only the 21-value rows, so the pair sums are
            # taken once per chunk from the current states (set checkpoint_rate = measurement_rate for the reference's cadence on small lattices)
            for r in 1:R
                o = obsOf(obs, r)
                o isa ObservablesGeneric && push!(o.correlations, getCorrelations(cfg, o.project; replica = r))
            end
        end
        s = stop
        if s % cp == 0 || s == total
            writeFiles(filename, replica_filenames, cfg, obs, max(current_sweep, s), βs, energy, σs .* SQRT2, Ts, s > N_therm)
        end
        if s <= N_therm
            println("$filename: $s/$N_therm thermalization sweeps. T = $(round(1 / βs[end], sigdigits = 4)), σ = $(round(σs[1] * SQRT2, sigdigits = 4)).")
        else
            println("$filename: $(s - N_therm)/$N_measure measurement sweeps at $(round(acc[1], sigdigits = 4)) acceptance rate.")
        end
        flush(stdout)
    end
    println("$filename: Calculation finished"); flush(stdout)
    return nothing
end

# checkpoint! + saveMeans! of replica 1 into `filename` (the reference's file); replica r > 1 into replica_filenames[r] when given
function writeFiles(filename, replica_filenames, cfg, obs, sweep, βs, energy, σs, Ts, have_rows)
    o1 = obsOf(obs, 1)
    checkpoint!(filename, cfg, o1, sweep, βs, energy, σs[1])
    (have_rows && o1 !== nothing) && saveMeans!(filename, o1, cfg, 1 / Ts[1])
    replica_filenames === nothing && return nothing
    for r in 2:cfg.nReplicas
        o = obsOf(obs, r)
        (have_rows && o !== nothing) && saveMeans!(replica_filenames[r], o, cfg, 1 / Ts[r])
    end
    return nothing
end

struct AnnealParams
    T_i::Float64; T_fac::Float64; N_max::Int64; N_per_T::Int64; min_acc_per_site::Int64
    min_accrate::Float64; sigma_min::Float64; sigma0::Float64; hits::Int32; log_rate::Int32; seed::UInt64
end
mutable struct AnnealResults
    E::Ptr{Float64}; beta::Ptr{Float64}; sigma::Ptr{Float64}; sweeps::Ptr{Int64}
    energy_log::Ptr{Float64}; beta_log::Ptr{Float64}; n_log_cap::Int64; n_log::Int64
end

# runAnnealing!(cfg, T_i, filename; ...) (src/MonteCarlo.jl:112-238): cooling loop, N_o optimisation sweeps, checkpoint! + f["state"] at the
# end (:230-234).  Every replica (restart) follows its own schedule on the device; energy / βs receive the final E/N and β of replica 1.
# The best restart is replica argmin(results.E).
function runAnnealing!(cfg::Configuration, T_i, filename::String; T_fac = 0.98, N_max::Int = 10000000, N_o::Int = 0, N_per_T::Int = 10000,
                       min_acc_per_site::Int = 100, min_accrate = 1e-5, checkpoint_rate::Int = 100000, σ_min::Float64 = 0.05,
                       verbose::Bool = false, βs::Vector{Float64} = Float64[], energy::Vector{Float64} = Float64[], current_sweep::Int = 1,
                       σ::Float64 = 60.0, hits::Int = 2, results = nothing)::Nothing
    R = cfg.nReplicas
    E = zeros(R); β = zeros(R); σs = zeros(R); sweeps = zeros(Int64, R)
    seed = newEpoch!(cfg, false)
    GC.@preserve E β σs sweeps begin
        p = AnnealParams(T_i, T_fac, N_max, N_per_T, min_acc_per_site, min_accrate, σ_min / SQRT2, σ / SQRT2, hits, 0, seed)
        out = AnnealResults(pointer(E), pointer(β), pointer(σs), pointer(sweeps), C_NULL, C_NULL, 0, 0)
        check(ccall((:scmc_anneal, LIB), Int32, (Ptr{Cvoid}, Ref{AnnealParams}, Ref{AnnealResults}), cfg.handle, p, out))
        N_o > 0 && check(ccall((:scmc_optimize, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}), cfg.handle, N_o, E))
    end
    push!(energy, E[1] / length(cfg)); push!(βs, β[1])
    checkpoint!(filename, cfg, Int64(sweeps[1]), βs, energy, σs[1] * SQRT2)
    h5open(filename, "cw") do f
        f["state"] = getStateMatrix(cfg)
    end
    if results !== nothing                 # optional out-parameter (a Dict): per-restart results of a packed launch
        results[:E] = E; results[:β] = β; results[:σ] = σs .* SQRT2; results[:sweeps] = sweeps
    end
    println("$filename: Finished calculation. E/N = $(E[1] / length(cfg))"); flush(stdout)
    return nothing
end

# optimisation sweeps (localOptimization!, src/MonteCarlo.jl:289-329): the device moves every site to the minimiser of its local energy
function localOptimization!(cfg::Configuration, E::Float64 = 0.0, i::Int64 = 0; n_sweeps::Int = 1)
    En = zeros(cfg.nReplicas)
    check(ccall((:scmc_optimize, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}), cfg.handle, n_sweeps, En))
    return En[1]
end

## ------------------------------------------------------------------------------------------------ launchers (src/Launcher.jl:49-90, 142-182)
function launch!(filename::String, model::String, n::Int, obstype::Type{<:Observables}, latticename::String, L::Int, J::AbstractVector,
                 T::Number, T_i::Number, N_therm::Int, N_measure::Int; measurement_rate::Int = 10, checkpoint_rate::Int = 100000,
                 report_rate::Int = 100000, seed::Int = abs(rand(Random.RandomDevice(), Int)), overwrite::Bool = true, kwargs...)::Nothing
    Random.seed!(seed)
    if overwrite
        println("$filename: Overwrite = true, starting from scratch.")
        cfg = initializeCfg(model, latticename, J, L, n; seed = seed, kwargs...)
        obs = initializeObservables(obstype, cfg)
        run!(cfg, obs, T, T_i, N_therm, N_measure, filename; measurement_rate = measurement_rate, checkpoint_rate = checkpoint_rate, report_rate = report_rate)
    else
        println("$filename: Overwrite = false, trying to load data and continue calculations")
        cfg, obs, current_sweep, energy, βs, σ = readCheckpoint(filename, obstype)
        if current_sweep >= N_therm + N_measure
            println("$filename: Already reached $(N_therm + N_measure) sweeps, terminating.")
        else
            # unlike the reference (which reseeds, SURVEY Q13) this continues the SAME chain: the stream is counter based
            run!(cfg, obs, T, T_i, N_therm, N_measure, filename; measurement_rate = measurement_rate, checkpoint_rate = checkpoint_rate,
                 report_rate = report_rate, energy = energy, βs = βs, current_sweep = current_sweep, σ = σ)
        end
    end
    return nothing
end

# A temperature grid as ONE launch: what the reference does with one Slurm job step per temperature (src/Launcher.jl:239-345).
# filename(T) names the result file of temperature T, as in notebook cell 22.
function launch!(filename::Function, model::String, n::Int, obstype::Type{<:Observables}, latticename::String, L::Int, J::AbstractVector,
                 Ts::AbstractVector{<:Number}, T_i::Number, N_therm::Int, N_measure::Int; measurement_rate::Int = 10,
                 checkpoint_rate::Int = 100000, report_rate::Int = 100000, seed::Int = abs(rand(Random.RandomDevice(), Int)), kwargs...)::Nothing
    cfg = initializeCfg(model, latticename, J, L, n; seed = seed, replicas = length(Ts), kwargs...)
    obs = [initializeObservables(obstype, cfg) for _ in Ts]
    files = [filename(T) for T in Ts]
    run!(cfg, obs, Ts, T_i, N_therm, N_measure, files[1]; measurement_rate = measurement_rate, checkpoint_rate = checkpoint_rate,
         report_rate = report_rate, replica_filenames = files)
    return nothing
end

function launchAnnealing!(filename::String, model::String, n::Int, latticename::String, L::Int, J::AbstractVector, T_i::Number;
                          T_fac::Number = 0.98, N_max::Int = 10000, N_o::Int = 0, N_per_T::Int = 100, min_acc_per_site::Int = 100,
                          min_accrate::Number = 0.0, checkpoint_rate::Int = 10000, σ_min::Float64 = 0.05,
                          seed::Int = abs(rand(Random.RandomDevice(), Int)), verbose::Bool = false, kwargs...)::Nothing
    Random.seed!(seed)
    cfg = initializeCfg(model, latticename, J, L, n; seed = seed, kwargs...)          # replicas = K: K annealing restarts in one launch
    runAnnealing!(cfg, T_i, filename; T_fac = T_fac, N_max = N_max, N_o = N_o, N_per_T = N_per_T, min_acc_per_site = min_acc_per_site,
                  min_accrate = min_accrate, checkpoint_rate = checkpoint_rate, σ_min = σ_min, verbose = verbose)
    return nothing
end

## ------------------------------------------------------------------------------------------------ several GPUs of one box (scmc_multi_*)
# One Julia process, replicas split over the GPUs of the box; the only exchange between GPUs is one ncclAllGather of the measure! rows.
mutable struct MultiConfiguration
    handle    :: Ptr{Cvoid}
    nDevices  :: Int64
    nSites    :: Int64
    nReplicas :: Int64          # over all devices
    d         :: Int64
    seed      :: UInt64
end
Base.length(m::MultiConfiguration) = m.nSites

function MultiConfiguration(n_devices::Int, L::Int64, uc, interactions::Vector{Matrix{Float64}}, onsiteInteraction::Vector{Float64}, n::Int64;
                            replicas::Int, precision::Int = 32, replica_offset::Int = 0, seed::Integer = abs(rand(Random.RandomDevice(), Int)))
    l, N, iSites, iLabels, z, nbr, lab, J, gen, gensq = abiTables(uc, L, interactions, n)
    gre, gim = parts(gen); qre, qim = parts(gensq)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:scmc_multi_create, LIB), Int32,
                (Ref{Ptr{Cvoid}}, Int32, Ptr{Int32}, Int32, Int64, Int32, Int64, Int64, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                h, n_devices, C_NULL, precision, N, size(gen[1], 1), replicas, replica_offset, z, nbr, lab, length(interactions), J,
                onsiteInteraction, gre, gim, qre, qim, C_NULL))
    m = MultiConfiguration(h[], n_devices, N, replicas, size(gen[1], 1), UInt64(seed))
    check(ccall((:scmc_multi_randomize, LIB), Int32, (Ptr{Cvoid}, UInt64), m.handle, m.seed))
    finalizer(c -> ccall((:scmc_multi_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), m)
    return m
end

# measure! rows of all replicas of all devices (NOBS x R): per-device reductions + one ncclAllGather over NVLink
function measureRows(m::MultiConfiguration)
    rows = zeros(NOBS, m.nReplicas)
    check(ccall((:scmc_multi_measure, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), m.handle, rows))
    return rows
end

# run! for a temperature grid over all GPUs: obs is a vector of observables, one per replica; returns nothing, mutates obs / σ
function run!(m::MultiConfiguration, obs::AbstractVector, Ts::AbstractVector{<:Number}, T_i, N_therm::Int, N_measure::Int, filename::String;
              measurement_rate::Int = 10, σ = 60.0, hits::Int = 2)::Nothing
    R = m.nReplicas
    T = Vector{Float64}(Ts); Tis = perReplica(T_i, R); σs = perReplica(σ, R) ./ SQRT2
    rows_max = max((N_therm + N_measure) ÷ measurement_rate - N_therm ÷ measurement_rate, 1)
    series = zeros(NOBS, R, rows_max); means_ = zeros(NOBS, R); acc = zeros(R)
    out = RunResults(C_NULL, 0, C_NULL, C_NULL, C_NULL)
    GC.@preserve T Tis σs series means_ acc begin
        p = RunParams(N_therm, N_measure, measurement_rate, hits, 0, m.seed, 0, pointer(T), pointer(Tis), pointer(σs))
        out = RunResults(pointer(series), 0, pointer(means_), pointer(acc), C_NULL)
        check(ccall((:scmc_multi_run_metropolis, LIB), Int32, (Ptr{Cvoid}, Ref{RunParams}, Ref{RunResults}), m.handle, p, out))
    end
    for k in 1:out.n_rows, r in 1:R
        pushRow!(obs[r], series[:, r, k])
    end
    println("$filename: Calculation finished on $(m.nDevices) GPUs"); flush(stdout)
    return nothing
end

# annealing restarts over all GPUs; returns (E, β, σ, sweeps) per restart, energies after N_o optimisation sweeps
function runAnnealing!(m::MultiConfiguration, T_i; T_fac = 0.98, N_max::Int = 10000000, N_o::Int = 0, N_per_T::Int = 10000,
                       min_acc_per_site::Int = 100, min_accrate = 1e-5, σ_min::Float64 = 0.05, σ::Float64 = 60.0, hits::Int = 2)
    R = m.nReplicas
    E = zeros(R); β = zeros(R); σs = zeros(R); sweeps = zeros(Int64, R)
    GC.@preserve E β σs sweeps begin
        p = AnnealParams(T_i, T_fac, N_max, N_per_T, min_acc_per_site, min_accrate, σ_min / SQRT2, σ / SQRT2, hits, 0, m.seed)
        out = AnnealResults(pointer(E), pointer(β), pointer(σs), pointer(sweeps), C_NULL, C_NULL, 0, 0)
        check(ccall((:scmc_multi_anneal, LIB), Int32, (Ptr{Cvoid}, Ref{AnnealParams}, Ref{AnnealResults}), m.handle, p, out))
        N_o > 0 && check(ccall((:scmc_multi_optimize, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}), m.handle, N_o, E))
    end
    return (E = E, β = β, σ = σs .* SQRT2, sweeps = sweeps)
end

export Configuration, MultiConfiguration, initializeCfg, getPosition, getBasis, getInteractionSites, getInteractionLabels, getInteraction,
       getOnsiteInteraction, getGeneratorsSq, getState, getStateMatrix, setStateMatrix!, getSpinExpectation, getSpinSqExpectation, getEnergy,
       getEnergies, initializeObservables, measure!, measureRows, getMagnetization, getCorrelations, getRs, getProject, structureFactor,
       structureFactorGrid, getSublatticeMagnetization, getChirality, saveMeans!, checkpoint!, readCheckpoint, readCheckpointSA, run!,
       runAnnealing!, localUpdate!, localOptimization!, launch!, launchAnnealing!, setSweepMode!, setStreamVersion!, setLanes!, verifyTables
end
