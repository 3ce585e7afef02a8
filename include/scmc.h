/* scmc.h -- C ABI of libscmc_b200.so: the B200-native Metropolis / simulated-annealing sweep for su(4)
 * product-state spin-valley models, a drop-in for that ONE path of lgresista/SemiClassicalMC.jl.
 *
 * The reference (Julia, single-threaded CPU) has no FFI seam; the seam is its exported Julia API
 * (src/SemiClassicalMC.jl:19-87).  Each entry point below names the reference function(s) it replaces
 * (paths relative to /root/reference/).  A Julia `ccall` shim (julia/SemiClassicalMCB200.jl) and the Python
 * `ctypes` mirror (semiclassicalmc.jl_b200/) bind exactly these symbols; see INTEGRATION.md.
 *
 * Conventions
 *  - every function returns int32 status: 0 = ok, < 0 = error (text via scmc_last_error(), thread-local);
 *  - the caller owns all host pointers; the library owns device memory behind the opaque handle;
 *  - a handle is bound to one CUDA device and is NOT thread-safe; calls are synchronous on return unless noted;
 *  - all host-side reals are FP64; `precision` selects device storage/arithmetic (32: FP32 state + FP64 accumulators);
 *  - sites are numbered 0-based in the caller's order (the library permutes internally, colour-major);
 *  - matrices are C order (row-major): J[label][a][b] = J^{ab}  (T_i^a J^{ab} T_j^b, i = bond source);
 *    generators g[a][j][k] = G^a_{jk}.  Julia callers pass `permutedims` of their column-major arrays;
 *  - "replica" = one independent Configuration (temperature point / restart / copy); `replica_offset` is the
 *    global id of this handle's replica 0 and only enters the random stream, so results do not depend on how
 *    replicas are sharded over GPUs.
 *  - there is no CPU fallback: without a CUDA device every compute entry fails with SCMC_ERR_CUDA.
 */
#ifndef SCMC_H
#define SCMC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct scmc_handle scmc_t;

enum {
    SCMC_OK = 0,
    SCMC_ERR_ARG = -1,      /* invalid argument / inconsistent tables (e.g. J_label(j->i) != J_label(i->j)^T) */
    SCMC_ERR_CUDA = -2,     /* CUDA runtime error (including "no device") */
    SCMC_ERR_UNSUPPORTED = -3,
    SCMC_ERR_ALLOC = -4
};

#define SCMC_NOBS 21 /* per-replica measurement row: E/N, (E/N)^2, |m_sigma|, |m_tau|, |m_sigmatau|, m[16]  (measure!, ObservablesGeneric.jl:32-46) */

/* ---- library-level ---------------------------------------------------------------------------------------------- */
const char* scmc_last_error(void);
int32_t scmc_version(void);
int32_t scmc_device_count(int32_t* n);

/* ---- Configuration (struct Configuration + constructor, src/Configuration.jl:2-77; tables from src/Models.jl:23-113) */
int32_t scmc_create(scmc_t** h, int32_t device, int32_t precision /* 32 | 64 */, int64_t n_sites, int32_t d /* 4 | 6 */,
                    int64_t n_replicas, int64_t replica_offset, int32_t z_max,
                    const int32_t* nbr_site /* [N][z_max], 0-based, -1 = padding  (interactionSites,  Configuration.jl:46-47) */,
                    const int32_t* nbr_label /* [N][z_max], 0-based               (interactionLabels, Configuration.jl:48)    */,
                    int32_t n_labels, const double* J /* [n_labels][16][16]       (interactions,      Configuration.jl:12)    */,
                    const double* K /* [16] on-site couplings                     (onsiteInteraction, Configuration.jl:13)    */,
                    const double* gen_re, const double* gen_im /* [16][d][d]      (getGenerators(n),  Generators.jl:2-23)     */,
                    const double* gensq_re, const double* gensq_im /* [16][d][d]  (gen .^ 2,          Configuration.jl:52)    */,
                    const int32_t* colour /* [N] graph colouring or NULL -> greedy */);
int32_t scmc_destroy(scmc_t* h);
/* Use an existing CUDA stream (cudaStream_t as void*) for all launches/copies of this handle; NULL = legacy default stream. */
int32_t scmc_set_stream(scmc_t* h, void* cuda_stream);
/* info[0]=n_colours, [1]=generator table id (1,2,3), [2]=device bytes allocated, [3]=resident-kernel eligibility bits (1: T-gather family, 2: psi-gather family),
 * [4]=cluster size of the resident kernel, [5]=kernel launches issued so far, [6]=n_sites, [7]=n_replicas */
int32_t scmc_info(scmc_t* h, int64_t* info8);
/* Self-check of the DEVICE copies of the tables the sweep kernels read (no counterpart in the reference; stands in for a race checker): no bond inside
 * a colour class, every bond has its reverse, per-class tables and the bulk-copy tile plan resolve to exactly the neighbour table.  *n_errors = number of
 * violations (0 for every handle scmc_create returns). */
int32_t scmc_verify_tables(scmc_t* h, int64_t* n_errors);

/* cfg.state (Configuration.jl:17); layout [count][N][d], the d x N complex matrix of IO.jl:146-150 per replica */
int32_t scmc_set_state(scmc_t* h, int64_t replica0, int64_t count, const double* re, const double* im);
int32_t scmc_get_state(scmc_t* h, int64_t replica0, int64_t count, double* re, double* im);
/* Haar-random product state on every site of every replica (getRandomState, MonteCarlo.jl:244-246; Configuration.jl:62) */
int32_t scmc_randomize(scmc_t* h, uint64_t seed);
/* cfg.spinExpectation / spinSqExpectation of one replica, [N][16] each; Tsq may be NULL (Configuration.jl:18-19, 65-66) */
int32_t scmc_get_T(scmc_t* h, int64_t replica, double* T, double* Tsq);

/* ---- energies ----------------------------------------------------------------------------------------------------- */
/* getEnergy (Configuration.jl:236-253) for every replica: E[R] (total energy, not per site) */
int32_t scmc_energy(scmc_t* h, double* E);
/* parity probe: h_i^a = sum_j J_{label(i,j)}^{ab} T_j^b for one replica, [N][16] (inner sum of getEnergyDifference!, Configuration.jl:292-296) */
int32_t scmc_local_field(scmc_t* h, int64_t replica, double* hfield);

/* ---- local update (localUpdate!, MonteCarlo.jl:263-286) ------------------------------------------------------------- */
/* One update of one site with injected randoms: xi[2d] = d real-part then d imaginary-part normals (getRandomState!,
 * MonteCarlo.jl:249-260), u = the uniform of the acceptance test `rand() < exp(-beta dE)` (MonteCarlo.jl:277-278). */
int32_t scmc_update_deterministic(scmc_t* h, int64_t replica, int64_t site, const double* xi, double u, double beta,
                                  double sigma, double* dE, int32_t* accepted);
/* One colour-ordered sweep (colours ascending; `hits` consecutive updates per site visit) of all replicas with injected
 * randoms: xi [hits][R][N][2d], u [hits][R][N]; outputs accept [hits][R][N] (0/1) and dE [hits][R][N] (may be NULL). */
int32_t scmc_sweep_deterministic(scmc_t* h, int32_t hits, const double* beta /* [R] */, const double* sigma /* [R] */,
                                 const double* xi, const double* u, uint8_t* accept, double* dE);
/* n_sweeps colour-ordered sweeps driven by the counter-based stream ("stream v1"): the 2N-update loops of
 * run!/runAnnealing! (MonteCarlo.jl:40-44, 76-80, 152-156) with hits = 2.  sweep_counter = index of the first sweep
 * (visit number = sweep * hits + hit).  accepted/attempted: [R] totals over the call (may be NULL). */
int32_t scmc_sweep(scmc_t* h, int64_t n_sweeps, int32_t hits, const double* beta /* [R] */, const double* sigma /* [R] */,
                   uint64_t seed, uint64_t sweep_counter, int64_t* accepted, int64_t* attempted);
/* Same, but betas/sigmas stay as last set and nothing is copied back (for graphs / async pipelines). */
int32_t scmc_sweep_async(scmc_t* h, int64_t n_sweeps, int32_t hits, uint64_t seed, uint64_t sweep_counter);
/* kernel selection for the stream-driven sweeps: 0 = auto; T-gather family (neighbours enter through the cached T vectors):
 * 1 = streaming colour-pass kernel, 2 = replica-resident cluster kernel; psi-gather family (neighbours enter through their
 * states via density-matrix sums; one coupling label, or several at d = 4 for mode 3): 3 = streaming with a per-thread gather, 4 = replica-resident,
 * 5 = streaming, TMA-pipelined (cp.async.bulk + mbarrier ring; classes that do not decompose into bulk copies fall back to the
 * gather of mode 3), 6 = mode 5 with the 16x16 mat-vec of a site visit on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split,
 * accumulator in tensor memory; FP32 only -- this is what auto uses for streaming FP32 passes of the psi family unless SCMC_TC=0).
 * Kernels of one family produce bit-identical chains, except that mode 6 follows the psi-family chain to rounding (different
 * summation order in the mat-vec); the two families differ by rounding.  Auto picks the family from the model alone (never from
 * the replica count).  batch = replicas per batch for the streaming path (0 = all). */
int32_t scmc_set_sweep_mode(scmc_t* h, int32_t mode, int64_t batch);
/* Random stream of the stream-driven sweeps (the chain definition; stands in for randn!(local_rng(), ...) and rand(),
 * MonteCarlo.jl:255-256, 278): 1 (default) = Philox4x32-10, two calls and 24-bit Box-Muller pairs per update; 2 = Philox4x32-7, one call
 * and 16-bit Box-Muller pairs per update, acceptance uniforms of four visits from one further call.  Both are pure functions of
 * (seed, site, global replica, visit); a checkpoint must be resumed with the version it was written with.  scmc_randomize always uses 1. */
int32_t scmc_set_stream_version(scmc_t* h, int32_t version);
/* Replica lanes of the streaming kernels: 0 = auto (two lanes whenever a call is a chain of several launches), 1 = everything on
 * the handle's stream, 2 = the replicas of a call are split over two internal side streams that fork from / join into the
 * handle's stream, so the tail of one lane's colour pass is filled by the other lane's next pass.  Replicas never interact:
 * chains are bit-identical for every setting (no reference counterpart -- the reference runs one chain per process). */
int32_t scmc_set_lanes(scmc_t* h, int32_t lanes);

/* ---- observables (measure!, ObservablesGeneric.jl:32-46; getMagnetization, Observables.jl:16-18) -------------------- */
/* obs [R][SCMC_NOBS]; E is re-summed from the configuration */
int32_t scmc_measure(scmc_t* h, double* obs);
/* S^a(k) = |sum_i T_i^a exp(i k.r_i)|^2 / n_rs for every replica: equals computeStructureFactorVec(getCorrelations(...))
 * (Observables.jl:22-41, 116-128) on the allowed momenta of the periodic cluster.  ks [nk][dim], pos [N][dim], out [R][16][nk]. */
int32_t scmc_structure_factor(scmc_t* h, const double* ks, const double* pos, int32_t nk, int32_t dim, double n_rs, double* Sak);
/* The same S^a(k) for allowed momenta of a two-dimensional periodic L1 x L2 cluster with n_basis sites per cell, as a separable DFT over the
 * cell indices (O(N (L1 + L2)) for ALL momenta of a cell instead of O(N) per momentum): site i sits in cell (cell[i][0], cell[i][1]) with basis
 * index basis[i]; momentum k = (m1' b1 + m2' b2) / L is requested as m = (m1' mod L1, m2' mod L2) together with its basis phases
 * phase[k][b] = (cos, sin)(k . delta_b).  comp_mask = 0: output [R][16][nk] as scmc_structure_factor (computeStructureFactorVec,
 * Observables.jl:116-128); comp_mask != 0: output [R][nk], summed over the components a whose bit a is set (computeStructureFactor with
 * `components`, Observables.jl:101-113; the reference's default 1:15 is mask 0x7FFF) -- 16 times less data to copy back.  No direct
 * reference counterpart for the algorithm: the reference evaluates these formulas on averaged O(N^2) correlations. */
int32_t scmc_structure_factor_grid(scmc_t* h, int32_t L1, int32_t L2, int32_t n_basis, const int32_t* cell /* [N][2] */, const int32_t* basis /* [N] */,
                                   int32_t nk, const int32_t* m /* [nk][2] */, const double* phase /* [nk][n_basis][2] */, double n_rs, uint32_t comp_mask,
                                   double* Sak);
/* Running sum over measurements on the device: the component-summed S(k) of this call (same arguments, comp_mask != 0) is ADDED to a buffer
 * [R][nk] owned by the handle instead of being copied out (a measurement then costs the DFT only, not the copy).  scmc_structure_factor_fetch
 * returns the SUM (divide by *count for the mean; Sk may be NULL to read the count only) and, with reset != 0, drops the buffer.  The momentum
 * list and mask must stay the same between resets.  Mirrors how the reference averages correlations before computeStructureFactor
 * (ObservablesGeneric.jl:32-46, 49-72). */
int32_t scmc_structure_factor_accumulate(scmc_t* h, int32_t L1, int32_t L2, int32_t n_basis, const int32_t* cell, const int32_t* basis, int32_t nk,
                                         const int32_t* m, const double* phase, double n_rs, uint32_t comp_mask);
int32_t scmc_structure_factor_fetch(scmc_t* h, double* Sk /* [R][nk] or NULL */, int64_t* count, int32_t reset);
/* getCorrelations (Observables.jl:22-41) of one replica for a caller-supplied project map [N][N] (0-based), C [16][n_rs] */
int32_t scmc_correlations(scmc_t* h, int64_t replica, const int32_t* project, int64_t n_rs, double* C);

/* ---- model-specific reductions of the same T vectors (ObservablesTgHbn.jl) ----------------------------------------------- */
/* getSublatticeMagnetization (ObservablesTgHbn.jl:211-221): mean T over the sites of every label; site_label [N] in 0..n_sub-1,
 * out [R][n_sub][16] */
int32_t scmc_sublattice_magnetization(scmc_t* h, const int32_t* site_label, int32_t n_sub, double* out);
/* getChirality (ObservablesTgHbn.jl:224-249): staggered z vector chirality averaged over the two triangles of every site;
 * tri [N][4] = sites (i2, i3, i4, i5) of getInteractionSites(cfg, i)[[1, 5, 4, 2]]; out [R][3] = kappa_sigma (components
 * 5, 9, 13), kappa_tau (2:4), kappa_sigmatau (6:8 + 10:12 + 14:16) in the reference's 1-based component numbering */
int32_t scmc_chirality(scmc_t* h, const int32_t* tri, double* out);

/* ---- drivers -------------------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t n_therm;            /* N_therm  (run!, MonteCarlo.jl:3-18) */
    int64_t n_measure;          /* N_measure */
    int64_t measurement_rate;   /* default 10 */
    int32_t hits;               /* updates per site visit; 2 reproduces the reference's 2N updates per sweep */
    int32_t stop_sweep;         /* 0 = run to the end; else return after this sweep (a multiple of measurement_rate): the caller writes its
                                 * checkpoint_rate files (MonteCarlo.jl:56-59, 88-91) and continues with sweep0 = stop_sweep */
    uint64_t seed;
    uint64_t sweep0;            /* current_sweep (resume): sweeps sweep0 + 1 ... are run */
    const double* T;            /* [R] target temperatures */
    const double* T_i;          /* [R] initial temperatures of the thermalisation ramp */
    double* sigma;              /* [R] in/out cone widths (default 60.0) */
} scmc_run_params;
typedef struct {
    double* series;             /* [n_rows][R][SCMC_NOBS] measurement rows (caller-allocated, n_rows >= n_measure/measurement_rate) or NULL */
    int64_t n_rows;             /* out: rows written */
    double* mean;               /* [R][SCMC_NOBS] out: means over the measurement rows, or NULL */
    double* acc_rate;           /* [R] out: acceptance rate over the measurement phase, or NULL */
    double* therm_energy;       /* [n_therm/measurement_rate][R] E/N logged at sigma-adaptation points (push!, MonteCarlo.jl:48-49) or NULL */
} scmc_run_results;
/* run! (MonteCarlo.jl:3-108) for all replicas at once: 3/4 geometric ramp T_i -> T, sigma adaptation every measurement_rate
 * sweeps during thermalisation (min(sigma * 0.5 / (1 - R), 60)), then measurement sweeps with measure! every measurement_rate.
 * Where the replica-resident kernel is the preferred one (a replica in one CTA or one cluster of CTAs, few enough clusters: see
 * scmc_set_sweep_mode and DESIGN.md section 4) the whole schedule runs inside that persistent kernel (<= 512 sweeps per launch,
 * acceptance counts and measurement sums of a cluster reduced through distributed shared memory); otherwise controller kernels run
 * between sweep launches.  Both give the same chains, adapted sigma and acceptance rates; rows agree to rounding. */
int32_t scmc_run_metropolis(scmc_t* h, const scmc_run_params* p, scmc_run_results* out);

typedef struct {
    double T_i, T_fac;          /* runAnnealing! kwargs (MonteCarlo.jl:112-129) */
    int64_t N_max, N_per_T, min_acc_per_site;
    double min_accrate, sigma_min, sigma0;
    int32_t hits;
    int32_t log_rate;           /* > 0: E/N and beta of every replica are logged every log_rate sweeps (energy_log / beta_log below) */
    uint64_t seed;
} scmc_anneal_params;
typedef struct {
    double* E;                  /* [R] final energies (total) */
    double* beta;               /* [R] final inverse temperatures */
    double* sigma;              /* [R] final cone widths */
    int64_t* sweeps;            /* [R] annealing sweeps performed per replica */
    double* energy_log;         /* [n_log_cap][R] E/N series (push!(energy, E / N), MonteCarlo.jl:158-159, on a grid of log_rate sweeps) or NULL */
    double* beta_log;           /* [n_log_cap][R] beta series or NULL */
    int64_t n_log_cap;          /* rows the two logs can hold */
    int64_t n_log;              /* out: rows written (replicas that stopped earlier repeat their final values) */
} scmc_anneal_results;
/* cooling loop of runAnnealing! (MonteCarlo.jl:131-204), every replica with its own beta / sigma / stopping flag (device-side
 * inside the resident kernel where that kernel is the preferred one, one CTA or one cluster per replica, else one controller kernel
 * per sweep; identical chains) */
int32_t scmc_anneal(scmc_t* h, const scmc_anneal_params* p, scmc_anneal_results* out);
/* optimisation sweeps (localOptimization!, MonteCarlo.jl:214-225, 289-329): colour-ordered; every site is moved to the
 * minimiser of its local energy <psi|H_loc|psi> (the fixed point of the reference's per-site gradient descent). E [R] or NULL. */
int32_t scmc_optimize(scmc_t* h, int64_t n_sweeps, double* E);

/* ---- several GPUs of one box behind one object ------------------------------------------------------------------------
 * Replaces the reference's only parallelism, one Slurm job step per temperature merged from files (Launcher.jl:239-345, IO.jl:64-143), for
 * callers that are not one-process-per-GPU Python programs (the Julia shim, C).  Replicas / temperatures / restarts never interact: the
 * object is one scmc_t per device, each owning a contiguous range of the n_replicas_total GLOBAL replica ids (device i: first_replica ..
 * first_replica + n_replicas - 1; the ids enter the random stream, so a chain does not depend on the split).  Every array argument below is
 * indexed by the global replica and lives on the host.  Calls fan out to the devices from one host thread per device; the only exchange
 * between GPUs is in scmc_multi_measure: ONE ncclAllGather of the SCMC_NOBS-value measure! rows over NVLink (single process,
 * ncclCommInitAll; NCCL is dlopen'ed at first use).  devices = NULL means 0 .. n_dev-1.  A per-device handle obtained from
 * scmc_multi_handle can be used with every single-device entry point (structure factors, get/set_state, ...). */
typedef struct scmc_multi scmc_multi_t;
int32_t scmc_multi_create(scmc_multi_t** m, int32_t n_dev, const int32_t* devices, int32_t precision, int64_t n_sites, int32_t d,
                          int64_t n_replicas_total, int64_t replica_offset, int32_t z_max, const int32_t* nbr_site, const int32_t* nbr_label,
                          int32_t n_labels, const double* J, const double* K, const double* gen_re, const double* gen_im,
                          const double* gensq_re, const double* gensq_im, const int32_t* colour);
int32_t scmc_multi_destroy(scmc_multi_t* m);
int32_t scmc_multi_count(scmc_multi_t* m, int32_t* n_dev, int64_t* n_replicas_total);
int32_t scmc_multi_handle(scmc_multi_t* m, int32_t i, scmc_t** h, int64_t* first_replica, int64_t* n_replicas);
int32_t scmc_multi_randomize(scmc_multi_t* m, uint64_t seed);
int32_t scmc_multi_set_stream_version(scmc_multi_t* m, int32_t version);
int32_t scmc_multi_set_sweep_mode(scmc_multi_t* m, int32_t mode, int64_t batch);
int32_t scmc_multi_sweep(scmc_multi_t* m, int64_t n_sweeps, int32_t hits, const double* beta, const double* sigma, uint64_t seed,
                         uint64_t sweep_counter, int64_t* accepted, int64_t* attempted);
int32_t scmc_multi_energy(scmc_multi_t* m, double* E /* [R_total] */);
int32_t scmc_multi_optimize(scmc_multi_t* m, int64_t n_sweeps, double* E /* [R_total] */);
int32_t scmc_multi_get_state(scmc_multi_t* m, int64_t r0, int64_t count, double* re, double* im);
/* measure! rows of all replicas [R_total][SCMC_NOBS]: per-device reductions + one ncclAllGather (the final observable reduction) */
int32_t scmc_multi_measure(scmc_multi_t* m, double* obs);
/* run! / runAnnealing! cooling loop on all devices at once; parameter and result arrays have R_total entries per row */
int32_t scmc_multi_run_metropolis(scmc_multi_t* m, const scmc_run_params* p, scmc_run_results* out);
int32_t scmc_multi_anneal(scmc_multi_t* m, const scmc_anneal_params* p, scmc_anneal_results* out);

#ifdef __cplusplus
}
#endif
#endif /* SCMC_H */
