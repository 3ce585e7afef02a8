// handle.h -- host-side state behind the opaque scmc_t and small shared helpers (internal to libscmc_b200).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/scmc.h"
#include "device_core.cuh"

namespace scmc {

void set_error(const std::string& msg);
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SCMC_CK(call)                                                             \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return ::scmc::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)
#define SCMC_TRY(expr)                 \
    do {                               \
        int32_t _s = (expr);           \
        if (_s != SCMC_OK) return _s;  \
    } while (0)
#define SCMC_ARG(cond, msg)                                   \
    do {                                                      \
        if (!(cond)) { ::scmc::set_error(msg); return SCMC_ERR_ARG; } \
    } while (0)

// Device view of one model + its replicas, passed BY VALUE to kernels (couplings land in the constant bank).
template <class R, int NL>
struct DevModel {
    vec4<R>* psi;         // [D/2][R][Ns]  site states, colour-major site order
    vec4<R>* T;           // [4][R][Ns]    cached <G^a>
    vec4<R>* Tsq;         // [4][R][Ns]    cached <(G^a)^2> (only with on-site couplings at n = 2) else nullptr
    const int32_t* nbr;   // [z][N]        neighbour (permuted index) or -1
    const int8_t* lab;    // [z][N]        coupling label of that bond
    const int32_t* orig;  // [N]           permuted index -> caller's site number
    const int32_t* col_list;  // [N]       permuted indices grouped by colour: class c = col_list[col_off[c] .. col_off[c+1])
    const int32_t* cls_site;  // [N]       caller's site number by class position (= orig[col_list[.]])
    const int32_t* cls_nbr;   // [z][N]    neighbour (permuted index) by class position: one round trip instead of two
    int32_t N, Ns, z, n_col;
    int64_t nrep;
    int32_t col_off[MAXC + 1];
    Couplings<R, NL> cp;
};

// Arguments of one colour pass (streaming kernels of scmc_core.cu and scmc_tma.cu)
struct SweepArgs {
    int32_t colour, colour_last, hits;   // classes [colour, colour_last] are visited in order by ONE block per replica when colour_last > colour
    int64_t r0, rcount;        // replica batch [r0, r0 + rcount)
    uint64_t visit0, seed;     // visit number of hit 0 of this pass; stream key
    int32_t stream;            // random stream version (1 | 2, device_core.cuh)
    int64_t roff;              // global replica id of replica 0
    const void *beta, *sigma;  // [R] reals
    const void* Jqt;           // [16][16] psi family with 16 density entries: transposed Jq = C^T J C, or nullptr
    const void* Jt;            // [NL][16][16] transposed couplings Jt[l][b][a] = J_l^{ab} in the handle's precision (global)
    const uint8_t* active;     // [R] or nullptr
    unsigned long long* acc;   // [R]
    // deterministic test mode (nullptr in production): injected randoms / recorded decisions, caller's site order
    const double *inj_xi, *inj_u;
    uint8_t* out_acc;
    double* out_dE;
    int32_t lab_major;            // several labels: label l owns neighbour slots [lab_off[l], lab_off[l+1]) at every site
    int32_t lab_off[MAXL + 1];
};

}  // namespace scmc

struct scmc_handle {
    int device = 0, prec = 32, d = 4, z = 0, n_labels = 1, gen_id = 1, n_col = 0;
    bool onsite = false;      // Tsq planes live and on-site term evaluated per update (n = 2 with K != 0)
    int64_t N = 0, Ns = 0, R = 0, roff = 0;
    cudaStream_t stream = nullptr;
    std::vector<int32_t> colour, perm, orig;   // perm[orig site] = permuted index
    int32_t col_off[scmc::MAXC + 1] = {0};
    double J[scmc::MAXL][256] = {{0}}, K[16] = {0};
    double e_const = 0.0;     // N * sum_a K^a c_a when (G^a)^2 = c_a 1 (n = 1, 3): constant on-site energy
    scmc::CustomTables* custom = nullptr;     // caller-supplied generator tables (gen_id 4: d = 4, 5: d = 6), else null
    // device memory
    void *psi = nullptr, *T = nullptr, *Tsq = nullptr;
    int32_t *nbr = nullptr, *d_orig = nullptr, *d_col_list = nullptr, *d_cls_site = nullptr, *d_cls_nbr = nullptr;
    uint16_t* d_nbr16 = nullptr;                 // [z][N] resident-kernel neighbour codes: owner CTA << 12 | local index
    int8_t* lab = nullptr;
    void *beta = nullptr, *sigma = nullptr;      // [R] reals of the handle's precision
    void* d_Jt = nullptr;                        // [MAXL][16][16] transposed couplings (kernel stages them into shared memory)
    void* d_Jqt = nullptr;                       // [MAXL][16][16] transposed couplings in the density basis, Jq_l = C^T J_l C (psi family, 16 density entries)
    // several labels: neighbour slots are stably sorted by label at create; when every site has the same number of bonds per label,
    // label l owns slots [lab_off[l], lab_off[l+1]) at every site and the kernels accumulate one label at a time in registers
    int32_t lab_off[scmc::MAXL + 1] = {0};
    bool label_major = false;
    bool fused_density = false;                  // psi kernels go straight from density sums to energy coefficients (d_Jqt valid)
    int32_t n_fwd = 0;                           // one label: number of "forward" neighbour slots (z / 2) such that every bond i-j appears exactly once as (i, slot in fwd_slot)
    uint8_t fwd_slot[8] = {0};                   //            -- the energy measurement sums each bond once instead of twice (0: no such set, all slots with a factor 1/2)
    bool psi_labels = false;                     // several labels, d = 4, label-major slots: the psi-gather family is available through k_sweep_colour_psi_nl (d_Jqt holds one table per label)
    unsigned long long* acc = nullptr;           // [R] accepted-update counters
    uint8_t* active = nullptr;                   // [R] replica still running (annealing) -- nullptr semantics handled by callers
    double* obs = nullptr;                       // [R][SCMC_NOBS] scratch rows
    // grow-only device scratch of the API calls (every call that uses it synchronises the stream before it returns, so a slot is free again
    // at the next call).  cudaMalloc / cudaFree per call cost 0.3-1 ms each on a B200 box and now and then 10-300 ms (profiles/README.md).
    static constexpr int N_SCRATCH = 8;
    void* scratch[N_SCRATCH] = {nullptr};
    size_t scratch_bytes[N_SCRATCH] = {0};
    // running sum of the component-summed structure factor over measurements (scmc_structure_factor_accumulate / _fetch)
    double* sf_acc = nullptr;                    // [R][sf_acc_nk]
    int64_t sf_acc_nk = 0, sf_acc_count = 0;
    uint32_t sf_acc_mask = 0;
    double* d_meas_partial = nullptr;            // [R][parts][17] per-CTA partial sums of the psi-family measurement
    size_t meas_partial_bytes = 0;
    void* pinned = nullptr;                      // pinned host staging for the beta / sigma tables
    int64_t bytes = 0, launches = 0;
    int32_t stream_version = 1;                  // counter-based random stream of the sweeps: 1 = Philox4x32-10 / 24-bit Box-Muller, 2 = Philox4x32-7 / 16-bit (scmc_set_stream_version)
    int32_t sweep_mode = 0;                      // 0 auto, 1 streaming T-gather, 2 resident T-gather, 3 streaming psi-gather, 4 resident psi-gather
    bool res_family_T = false;                   // reserved: force the T-gather arithmetic family in auto mode
    bool T_stale = false;                        // psi-gather sweeps leave the T / Tsq planes behind; ensure_T() refreshes them on demand
    int32_t async_pipeline = 0;                  // streaming path: 1 = cp.async-pipelined kernel (measured slower on B200: issue-bound at 16 warps/SM, profiles/README.md)
    // replica-resident kernel plan (scmc_resident.cu): a cluster of `cs` CTAs keeps one replica in distributed shared memory
    struct {
        int eligible = 0, eligible_psi = 0, cs = 0, n_own_max = 0, cap = 0, nthreads = 0, max_clusters = 0, max_clusters_psi = 0;
        size_t smem = 0, smem_psi = 0;            // T-gather family (psi + T resident) / psi-gather family (psi only)
        int cta_off[17] = {0};                    // first permuted site of every CTA's chunk
        int cta_col_off[16][scmc::MAXC + 1] = {{0}};   // local offsets of the colour classes inside a chunk
    } res;
    int64_t batch = 0;
    // TMA-pipelined streaming psi-gather kernel (scmc_tma.cu): per colour class, the class positions are cut into tiles of <= 128 sites; the
    // states a tile needs (its own sites and all their neighbours, each once) are a few runs of consecutive permuted indices, which one
    // cp.async.bulk per run and plane brings into a shared-memory stage.  ok = 0: the class keeps the direct-gather kernel (seam classes).
    struct TmaClass { int ok = 0, n_tiles = 0, tile = 0, first = 0; };
    struct {
        int ready = 0, cap = 0, max_runs = 0;        // cap: float4 slots per plane and stage (largest tile + the all-zero slot, multiple of 8)
        TmaClass cls[scmc::MAXC];
        void* d_meta = nullptr;                      // [tiles][128][2] uint4: stage slots of a thread's own site and neighbours, caller's site number, permuted index
        void* d_runs = nullptr;                      // [tiles][TMA_MAX_RUNS] uint2: first permuted index, stage slot | length << 16
        int32_t* d_tile = nullptr;                   // [tiles][4]: sites in the tile, runs, slots copied per plane, first class position
    } tma;
    // replica lanes of the streaming path: the replicas of a call are split over side streams so that the tail of one lane's
    // colour pass overlaps with the other lane's next pass (colour passes of different replicas are independent)
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    int32_t lanes = 0;                           // 0 = auto, 1 = single stream, 2 = two lanes
    size_t esz() const { return prec == 64 ? 8 : 4; }
};

namespace scmc {
// implemented in scmc_core.cu
int32_t upload_scalars(scmc_handle* h, const double* beta, const double* sigma);
int32_t launch_sweeps(scmc_handle* h, int64_t n_sweeps, int32_t hits, uint64_t seed, uint64_t sweep_counter, const uint8_t* active);
int32_t launch_measure(scmc_handle* h, double* d_rows /* device [R][SCMC_NOBS] or null */, double* d_Etot = nullptr /* device [R] or null */);
int32_t dev_alloc(scmc_handle* h, void** p, size_t bytes);
int32_t ensure_T(scmc_handle* h);      // recompute the cached T (Tsq) planes from psi if a psi-gather sweep left them stale
bool uses_psi_gather(const scmc_handle* h);
// device scratch slot of at least `bytes` bytes (kept in the handle, grown on demand); a DevBuf above 64 MB releases its own slot when it goes out of scope
cudaError_t scratch_get(scmc_handle* h, int slot, size_t bytes, void** p);
void scratch_release(scmc_handle* h, int slot);      // frees ONE slot (the caller is done with it and it is oversized)
bool resident_preferred(const scmc_handle* h);      // the auto policy's cluster predicate: clusters of <= 2 CTAs, or all clusters on the chip at once
int32_t launch_resident_anneal(scmc_handle* h, int64_t chunk, int32_t hits, uint64_t seed, double att_per_sweep, double min_accepted,
                               double beta_fac, double min_accrate, double sigma_min, int64_t n_per_T, int64_t n_total,
                               int64_t* d_sweeps_at_T, int64_t* d_sweeps_done, int* d_n_active);
// implemented in scmc_tma.cu
#ifndef SCMC_TMA_STAGES
#define SCMC_TMA_STAGES 2
#endif
constexpr int TMA_TILE = 128, TMA_MAX_RUNS = 16, TMA_STAGES = SCMC_TMA_STAGES, TMA_RC = 128;   // TMA_RC: most replicas one CTA walks over
void plan_tma(scmc_handle* h, const std::vector<int32_t>& col_list, const std::vector<int32_t>& nb /* [z][N] permuted neighbour by permuted site */);
bool tma_pass_available(const scmc_handle* h, const SweepArgs& A);
bool tma_use_tc(const scmc_handle* h);      // the handle's TMA passes run the tensor-core mat-vec variant (mode 6, auto unless SCMC_TC=0)
int32_t launch_colour_pass_tma(scmc_handle* h, const SweepArgs& A, cudaStream_t stream);
// implemented in scmc_resident.cu
void plan_resident(scmc_handle* h, const std::vector<int>& owner_of_perm);
int32_t launch_resident(scmc_handle* h, int64_t n_sweeps, int32_t hits, uint64_t seed, uint64_t sweep_counter, const uint8_t* active, bool psi);
int32_t launch_resident_run(scmc_handle* h, int mode, int64_t chunk, int32_t hits, uint64_t seed, const double* d_T, const double* d_Ti,
                            int64_t n_anneal, int64_t rate, int64_t sweep_first, int64_t row_base, double* d_rows);
}  // namespace scmc
