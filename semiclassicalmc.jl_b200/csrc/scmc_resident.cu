// scmc_resident.cu -- replica-resident Metropolis sweeps: one thread-block cluster keeps ONE replica's whole lattice
// (psi, T, (Tsq), neighbour codes, couplings) in distributed shared memory for all sweeps of a launch.
//
// Replaces the same reference loops as the streaming kernel (2N-update loops + localUpdate!, /root/reference/src/MonteCarlo.jl:40-44,
// 76-80, 152-156, 263-286) and produces bit-identical chains (same stream v1, same colour order, same arithmetic).
//
// Why: the streaming colour-pass kernel re-reads every neighbour T vector from HBM in every colour pass (272-332 B of DRAM
// traffic per site visit, profiles/).  Here HBM is touched once per launch (state in, state out = the algorithmic 4 d w bytes
// per site, plus the T cache), every neighbour read is a shared-memory read (own CTA) or a DSMEM read (ld.shared::cluster into
// the neighbouring strip's CTA), and colour passes are separated by cluster barriers instead of kernel launches.
//   C5 (128 x 128 triangular, FP32): 16384 sites x 96 B = 1.5 MB  ->  cluster of 8 CTAs x 216 KB.
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>
#include <cooperative_groups.h>
#include "handle.h"

namespace cg = cooperative_groups;

namespace scmc {

struct ResidentArgs {
    int64_t nrep, roff;
    int32_t n_sweeps, hits, cs, n_own_max, cap, n_clusters;
    uint64_t visit0, seed;
    int32_t stream;                 // random stream version (1 | 2)
    const void *beta, *sigma, *Jt, *Jqt;
    const uint8_t* active;
    unsigned long long* acc;
    const uint16_t* nbr16;          // [z][N]
    int32_t cta_off[17];
    int32_t cta_col_off[16][MAXC + 1];
    // device-side annealing controller (runAnnealing!, src/MonteCarlo.jl:145-204); single-CTA replicas only
    int32_t anneal;                 // 0 = plain sweeps
    double att_per_sweep, min_accepted, beta_fac, min_accrate, sigma_min;
    int64_t n_per_T, n_total;       // sweeps per temperature cap; total sweep budget (N_max - 1)
    int64_t *sweeps_at_T, *sweeps_done;
    uint8_t* active_rw;
    int* n_active;                  // out: replicas still running after this launch
    // device-side run! schedule (src/MonteCarlo.jl:38-99); single-CTA replicas only.  run_mode 1 = thermalisation chunk (beta ramp
    // per sweep, sigma adaptation every `rate` sweeps), 2 = measurement chunk (beta, sigma frozen; measure! row every `rate` sweeps)
    int32_t run_mode;
    const double *run_T, *run_Ti;   // [R]
    int64_t n_anneal, rate, sweep_first, row_base;   // sweep_first: 1-based index of this launch's first sweep; row = g / rate - row_base
    double att_window, e_const;     // attempts between two adaptations; constant on-site energy
    double* rows;                   // [rows of this launch][R][SCMC_NOBS]
};

// ---- shared-memory / DSMEM access helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
template <bool CLUSTER> __device__ __forceinline__ float4 ld_v4(uint32_t addr, float) {
    float4 v;
    if (CLUSTER) asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    else asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <bool CLUSTER> __device__ __forceinline__ double4 ld_v4(uint32_t addr, double) {
    double4 v;
    if (CLUSTER) {
        asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
        asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(addr + 16));
    } else {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(addr + 16));
    }
    return v;
}

// FP64 value of another CTA of the cluster (or of this CTA when CLUSTER is false) at shared address `addr`
template <bool CLUSTER> __device__ __forceinline__ double ld_f64(uint32_t addr) {
    double v;
    if (CLUSTER) asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Sum of one FP64 value per CTA over the cluster, added in rank order in every CTA (identical result everywhere, deterministic).
// `slot` is a __shared__ double of this CTA; two cluster barriers.  Without a cluster: the value itself.
template <bool CLUSTER> __device__ __forceinline__ double cluster_sum(double x, double* slot, int cs, int tid) {
    if (!CLUSTER) return x;
    if (tid == 0) *slot = x;
    cg::this_cluster().sync();
    double v = 0.0;
    for (int b = 0; b < cs; b++) v += ld_f64<CLUSTER>(map_to_rank(smem_u32(slot), (uint32_t)b));
    cg::this_cluster().sync();
    return v;
}

// One replica at a time per cluster; grid = n_clusters * cs CTAs, persistent over replicas.  T-gather family: psi, T (Tsq) of the replica
// live in (distributed) shared memory.  Besides plain sweeps the kernel runs the device-side schedules of k_resident_psi: runAnnealing!'s
// cooling loop (A.anneal) and run! (A.run_mode 1 = thermalisation chunk, 2 = measurement chunk), with the same control flow and formulas.
template <class R, int GEN, bool ONSITE, int NL, bool CLUSTER>
__global__ void __launch_bounds__(512, 1) k_resident(const __grid_constant__ DevModel<R, NL> M, const __grid_constant__ ResidentArgs A) {
    constexpr int D = GenOps<GEN>::D;
    constexpr int NPV = D / 2;
    using V = vec4<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cap = A.cap, nmax = A.n_own_max;
    V* sT = reinterpret_cast<V*>(smem_raw);                       // [4][cap]
    V* sPsi = sT + 4 * cap;                                       // [NPV][cap]
    V* sTsq = sPsi + NPV * cap;                                   // [4][cap] (ONSITE only)
    R* sJt = reinterpret_cast<R*>(sTsq + (ONSITE ? 4 * cap : 0)); // [NL][16][16]
    uint16_t* sNbr = reinterpret_cast<uint16_t*>(sJt + NL * NG * NG);        // [z][nmax]
    int8_t* sLab = reinterpret_cast<int8_t*>(sNbr + (size_t)M.z * nmax);     // [z][nmax] (NL > 1 only)
    __shared__ unsigned s_acc;
    __shared__ double s_red[16][NG + 1];                          // measurement reduction (one row per warp)
    __shared__ double s_xchg[NG + 1];                             // this CTA's partial sums, read by the other CTAs of the cluster

    unsigned rank = 0, cluster_id = blockIdx.x;
    if (CLUSTER) {
        cg::cluster_group cl = cg::this_cluster();
        rank = cl.block_rank();
        cluster_id = blockIdx.x / A.cs;
    }
    const int my_off = A.cta_off[rank];
    const int n_own = A.cta_off[rank + 1] - my_off;
    const int tid = threadIdx.x, nth = blockDim.x;

    // ---- one-time staging: couplings, neighbour codes (and labels) of my sites, the all-zero padding slot
    for (int t = tid; t < NL * NG * NG; t += nth) sJt[t] = ((const R*)A.Jt)[t];
    for (int k = 0; k < M.z; k++)
        for (int q = tid; q < n_own; q += nth) {
            sNbr[k * nmax + q] = A.nbr16[(size_t)k * M.N + my_off + q];
            if (NL > 1) sLab[k * nmax + q] = M.lab[(size_t)k * M.N + my_off + q];
        }
    if (tid < 4) sT[tid * cap + nmax] = make_v4<R>(R(0), R(0), R(0), R(0));
    if (tid == 0) s_acc = 0;
    __syncthreads();

    const size_t ps = (size_t)M.nrep * M.Ns;
    const uint32_t sT_u32 = smem_u32(sT);
    const uint32_t plane_bytes = (uint32_t)cap * (uint32_t)sizeof(V);

    // local field of my site q from the neighbours' T vectors in (distributed) shared memory: h = sum_l J_l sum_{j in l} T_j
    auto field_at = [&](int q, R* h) {
        R S[NL][NG];
#pragma unroll
        for (int l = 0; l < NL; l++)
#pragma unroll
            for (int a = 0; a < NG; a++) S[l][a] = R(0);
#pragma unroll 3
        for (int k = 0; k < M.z; k++) {
            const unsigned code = sNbr[k * nmax + q];
            const int l = NL == 1 ? 0 : sLab[k * nmax + q];
            uint32_t addr = sT_u32 + (code & 0x1FFFu) * (uint32_t)sizeof(V);
            if (CLUSTER) addr = map_to_rank(addr, code >> 13);
            V v[4];
#pragma unroll
            for (int p4 = 0; p4 < 4; p4++) v[p4] = ld_v4<CLUSTER>(addr + p4 * plane_bytes, R(0));
#pragma unroll
            for (int ll = 0; ll < NL; ll++)
                if (NL == 1 || l == ll) {
#pragma unroll
                    for (int p4 = 0; p4 < 4; p4++) {
                        S[ll][4 * p4] += v[p4].x; S[ll][4 * p4 + 1] += v[p4].y; S[ll][4 * p4 + 2] += v[p4].z; S[ll][4 * p4 + 3] += v[p4].w;
                    }
                }
        }
        field_from_sums_smem<NL>(sJt, S, h);
    };

    for (int64_t r = cluster_id; r < A.nrep; r += A.n_clusters) {
        if (A.active != nullptr && !A.active[r]) continue;          // uniform over the cluster
        const int64_t gbase = r * M.Ns + my_off;
        // ---- load my strip of the replica (coalesced 128-bit loads, state in)
        for (int q = tid; q < n_own; q += nth) {
#pragma unroll
            for (int v = 0; v < NPV; v++) sPsi[v * cap + q] = M.psi[gbase + q + v * ps];
#pragma unroll
            for (int v = 0; v < 4; v++) sT[v * cap + q] = M.T[gbase + q + v * ps];
            if (ONSITE) {
#pragma unroll
                for (int v = 0; v < 4; v++) sTsq[v * cap + q] = M.Tsq[gbase + q + v * ps];
            }
        }
        R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
        R nbeta = neg_beta_scaled(beta);
        // annealing state of this replica (every thread of every CTA keeps identical copies; the controller is deterministic)
        const bool anneal = A.anneal != 0;
        int64_t an_done = anneal ? A.sweeps_done[r] : 0, an_sat = anneal ? A.sweeps_at_T[r] : 0;
        double an_acc = anneal ? (double)A.acc[r] : 0.0;
        bool an_still = true;
        // run! schedule of this replica; run_acc is this CTA's share of the window's accepted updates (rank 0 carries the count over)
        const int run = A.run_mode;
        double run_acc = (run && rank == 0) ? (double)A.acc[r] : 0.0;
        const double run_T = run == 1 ? A.run_T[r] : 1.0, run_Ti = run == 1 ? A.run_Ti[r] : 1.0;
        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
        unsigned nacc = 0;
        for (int sweep = 0; sweep < A.n_sweeps && an_still; sweep++) {
            if (anneal && an_done >= A.n_total) break;
            const int64_t g = A.sweep_first + sweep;               // 1-based global sweep index (run modes)
            if (run == 1) {                                        // geometric ramp T_i -> T over the first n_anneal sweeps (MonteCarlo.jl:24-27)
                double t = run_T;
                if (g <= A.n_anneal && A.n_anneal > 1) t = run_Ti * pow(run_T / run_Ti, (double)(g - 1) / (double)(A.n_anneal - 1));
                beta = (R)(1.0 / t);
                nbeta = neg_beta_scaled(beta);
            }
            const uint64_t visit_base = anneal ? (uint64_t)an_done * (uint64_t)A.hits
                                        : run  ? (uint64_t)(g - 1) * (uint64_t)A.hits
                                               : A.visit0 + (uint64_t)sweep * (uint64_t)A.hits;
            unsigned sweep_acc = 0;
            for (int c = 0; c < M.n_col; c++) {
                const int c0 = A.cta_col_off[rank][c], cnt = A.cta_col_off[rank][c + 1] - c0;
                for (int s = tid; s < cnt; s += nth) {
                    const int q = c0 + s;
                    R re[D], im[D], T[NG], Tsq[NG], h[NG];
                    load_psi<R, D>(sPsi + q, (size_t)cap, re, im);
                    load_T<R>(sT + q, (size_t)cap, T);
                    if (ONSITE) load_T<R>(sTsq + q, (size_t)cap, Tsq);
                    const int site = M.orig[my_off + q];
                    field_at(q, h);
                    R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
                    unsigned mine = 0;
                    for (int hit = 0; hit < A.hits; hit++) {
                        R xr[D], xi[D], u, dE;
                        stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), visit_base + hit, xr, xi, u);
                        mine += metropolis_hit<R, GEN, ONSITE>(re, im, T, Tsq, h, M.cp.K, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
                    }
                    if (mine) {
                        store_psi<R, D>(sPsi + q, (size_t)cap, re, im);
                        store_T<R>(sT + q, (size_t)cap, T);
                        if (ONSITE) store_T<R>(sTsq + q, (size_t)cap, Tsq);
                    }
                    sweep_acc += mine;
                }
                // all CTAs finish the colour class before anyone reads its T vectors (release/acquire at cluster scope)
                if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
            }
            if (!run && !anneal) { nacc += sweep_acc; continue; }
            // ---- schedules: accepted updates of this sweep in my CTA
            {
                const unsigned wt = __reduce_add_sync(0xffffffffu, sweep_acc);
                if ((tid & 31) == 0 && wt) atomicAdd(&s_acc, wt);
                __syncthreads();
            }
            const double mine_sweep = (double)s_acc;
            __syncthreads();
            if (tid == 0) s_acc = 0;
            if (run) {
                run_acc += mine_sweep;
                if (g % A.rate == 0) {
                    if (run == 1) {                                // sigma <- min(sigma * 0.5 / (1 - R), 60)  (MonteCarlo.jl:50-53)
                        const double acc_all = cluster_sum<CLUSTER>(run_acc, &s_xchg[0], A.cs, tid);
                        const double ratio = acc_all / A.att_window;
                        sigma = (R)fmin((double)sigma * 0.5 / (1.0 - ratio), 60.0);
                        run_acc = 0.0;
                    }
                    // E and sum_i T_i from the resident T (Tsq) vectors (measure!, ObservablesGeneric.jl:32-46; thermalisation energy log,
                    // MonteCarlo.jl:48-49): run == 2 writes a measure! row, run == 1 logs E/N
                    if (run == 2 || A.rows != nullptr) {
                        double eacc = 0.0, macc[NG];
#pragma unroll
                        for (int t = 0; t < NG; t++) macc[t] = 0.0;
                        for (int q = tid; q < n_own; q += nth) {
                            R T[NG], h[NG];
                            load_T<R>(sT + q, (size_t)cap, T);
                            field_at(q, h);
                            R e = R(0);
#pragma unroll
                            for (int t = 0; t < NG; t++) e = fma(T[t], h[t], e);
                            e *= R(0.5);
                            if (ONSITE) {
                                R Tsq[NG];
                                load_T<R>(sTsq + q, (size_t)cap, Tsq);
#pragma unroll
                                for (int t = 0; t < NG; t++) e = fma(Tsq[t], M.cp.K[t], e);
                            }
                            eacc += (double)e;
#pragma unroll
                            for (int t = 0; t < NG; t++) macc[t] += (double)T[t];
                        }
                        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
                        for (int t = 0; t <= NG; t++) {
                            double v = t < NG ? macc[t < NG ? t : 0] : eacc;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0) s_red[warp][t] = v;
                        }
                        __syncthreads();
                        if (tid <= NG) {
                            double v = 0.0;
                            for (int w = 0; w < (nth >> 5); w++) v += s_red[w][tid];
                            s_xchg[tid] = v;
                        }
                        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
                        if (tid == 0 && rank == 0) {
                            double tot[NG + 1];
                            for (int t = 0; t <= NG; t++) {
                                double v = 0.0;
                                for (int b = 0; b < (CLUSTER ? A.cs : 1); b++) v += ld_f64<CLUSTER>(CLUSTER ? map_to_rank(smem_u32(&s_xchg[t]), (uint32_t)b) : smem_u32(&s_xchg[t]));
                                tot[t] = v;
                            }
                            const double E = tot[NG] + A.e_const, invN = 1.0 / (double)M.N;
                            const size_t slot = (size_t)(g / A.rate - A.row_base) * A.nrep + r;
                            if (run == 1) {
                                A.rows[slot] = E * invN;
                            } else {
                                double* row = A.rows + slot * SCMC_NOBS;
                                double m[NG];
                                for (int t = 0; t < NG; t++) m[t] = tot[t] * invN;
                                row[0] = E * invN;
                                row[1] = (E * invN) * (E * invN);
                                row[2] = sqrt(m[4] * m[4] + m[8] * m[8] + m[12] * m[12]);
                                row[3] = sqrt(m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
                                row[4] = sqrt(m[5] * m[5] + m[6] * m[6] + m[7] * m[7] + m[9] * m[9] + m[10] * m[10] + m[11] * m[11] + m[13] * m[13] +
                                              m[14] * m[14] + m[15] * m[15]);
                                for (int t = 0; t < NG; t++) row[5 + t] = m[t];
                            }
                        }
                        // nobody overwrites T vectors or the exchange slots before every CTA is done reading them
                        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
                    }
                }
            } else {
                // accepted updates of this sweep over the whole replica, then the cooling rule (MonteCarlo.jl:151-203)
                an_acc += cluster_sum<CLUSTER>(mine_sweep, &s_xchg[0], A.cs, tid);
                an_sat++; an_done++;
                const bool last = an_done >= A.n_total;
                if (!(an_acc < A.min_accepted && an_sat < A.n_per_T && !last)) {
                    const double ratio = an_acc / (A.att_per_sweep * (double)an_sat);
                    if (ratio > A.min_accrate) {
                        beta = (R)((double)beta * A.beta_fac);
                        sigma = (R)fmax(fmin((double)sigma * 0.5 / (1.0 - ratio), 60.0), A.sigma_min);
                        nbeta = neg_beta_scaled(beta);
                        an_acc = 0.0; an_sat = 0;
                    } else {
                        an_still = false;
                    }
                }
                __syncthreads();
            }
        }
        // ---- store my strip (state out)
        for (int q = tid; q < n_own; q += nth) {
#pragma unroll
            for (int v = 0; v < NPV; v++) M.psi[gbase + q + v * ps] = sPsi[v * cap + q];
#pragma unroll
            for (int v = 0; v < 4; v++) M.T[gbase + q + v * ps] = sT[v * cap + q];
            if (ONSITE) {
#pragma unroll
                for (int v = 0; v < 4; v++) M.Tsq[gbase + q + v * ps] = sTsq[v * cap + q];
            }
        }
        if (run) {
            const double acc_all = cluster_sum<CLUSTER>(run_acc, &s_xchg[0], A.cs, tid);
            if (tid == 0 && rank == 0) {
                ((R*)A.sigma)[r] = sigma;
                A.acc[r] = (unsigned long long)acc_all;
            }
            __syncthreads();
            continue;
        }
        if (anneal) {
            if (tid == 0 && rank == 0) {
                ((R*)A.beta)[r] = beta; ((R*)A.sigma)[r] = sigma;
                A.acc[r] = (unsigned long long)an_acc;
                A.sweeps_at_T[r] = an_sat; A.sweeps_done[r] = an_done;
                const bool running = an_still && an_done < A.n_total;
                A.active_rw[r] = an_still ? 1 : 0;
                if (running) atomicAdd(A.n_active, 1);
            }
            __syncthreads();
            continue;
        }
        const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
        if ((tid & 31) == 0 && tot) atomicAdd(&s_acc, tot);
        __syncthreads();
        if (tid == 0) {
            if (s_acc) atomicAdd(A.acc + r, (unsigned long long)s_acc);
            s_acc = 0;
        }
        // the next replica's loads overwrite shared memory that remote CTAs may still be reading in their last colour pass:
        // the cluster barrier that ended that pass already ordered those reads before this point.
        __syncthreads();
    }
}

// psi-gather family (same arithmetic as k_sweep_colour_psi, bit-identical chains): only psi lives in shared memory
// (d/2 x 16 B per site instead of psi + T), neighbours are read as states through DSMEM and enter through density-matrix sums.
// C5 (128 x 128, FP32): 16384 x (32 + 12) B = 704 KB -> cluster of 4 CTAs x 177 KB, all 148 SMs busy.
template <class R, int GEN, bool ONSITE, bool CLUSTER>
__global__ void __launch_bounds__(512, 1) k_resident_psi(const __grid_constant__ DevModel<R, 1> M, const __grid_constant__ ResidentArgs A) {
    constexpr int D = GenOps<GEN>::D, NPV = D / 2, NP = GenOps<GEN>::NDENS;
    using V = vec4<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cap = A.cap, nmax = A.n_own_max;
    V* sPsi = reinterpret_cast<V*>(smem_raw);                     // [NPV][cap]
    R* sJt = reinterpret_cast<R*>(sPsi + NPV * cap);              // [16][16]
    uint16_t* sNbr = reinterpret_cast<uint16_t*>(sJt + NG * NG);  // [z][nmax]
    __shared__ unsigned s_acc;
    __shared__ double s_red[16][NG + 1];                          // measurement reduction (one row per warp)
    __shared__ double s_xchg[NG + 1];                             // this CTA's partial sums, read by the other CTAs of the cluster (run modes)
    unsigned rank = 0, cluster_id = blockIdx.x;
    if (CLUSTER) {
        rank = cg::this_cluster().block_rank();
        cluster_id = blockIdx.x / A.cs;
    }
    const int my_off = A.cta_off[rank];
    const int n_own = A.cta_off[rank + 1] - my_off;
    const int tid = threadIdx.x, nth = blockDim.x;
    const bool fused = !ONSITE && NP == NG && A.Jqt != nullptr;
    for (int t = tid; t < NG * NG; t += nth) sJt[t] = ((const R*)(fused ? A.Jqt : A.Jt))[t];
    for (int k = 0; k < M.z; k++)
        for (int q = tid; q < n_own; q += nth) sNbr[k * nmax + q] = A.nbr16[(size_t)k * M.N + my_off + q];
    if (tid < NPV) sPsi[tid * cap + nmax] = make_v4<R>(R(0), R(0), R(0), R(0));     // all-zero padding slot
    if (tid == 0) s_acc = 0;
    __syncthreads();
    const size_t ps = (size_t)M.nrep * M.Ns;
    const uint32_t sPsi_u32 = smem_u32(sPsi);
    const uint32_t plane_bytes = (uint32_t)cap * (uint32_t)sizeof(V);
    for (int64_t r = cluster_id; r < A.nrep; r += A.n_clusters) {
        if (A.active != nullptr && !A.active[r]) continue;
        const int64_t gbase = r * M.Ns + my_off;
        for (int q = tid; q < n_own; q += nth) {
#pragma unroll
            for (int v = 0; v < NPV; v++) sPsi[v * cap + q] = M.psi[gbase + q + v * ps];
        }
        R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
        R nbeta = neg_beta_scaled(beta);
        // annealing state of this replica (all threads keep identical copies; the controller is deterministic)
        const bool anneal = A.anneal != 0;                     // clusters: the sweep's accepted count is reduced through DSMEM, every CTA takes the same decision
        int64_t an_done = anneal ? A.sweeps_done[r] : 0, an_sat = anneal ? A.sweeps_at_T[r] : 0;
        double an_acc = anneal ? (double)A.acc[r] : 0.0;
        bool an_still = true;
        // run! schedule of this replica
        const int run = A.run_mode;                            // clusters: acceptance counts and measurement sums are reduced through DSMEM
        double run_acc = (run && rank == 0) ? (double)A.acc[r] : 0.0;      // this CTA's share of the window's accepted updates (rank 0 carries the count over from the previous launch)
        const double run_T = run == 1 ? A.run_T[r] : 1.0, run_Ti = run == 1 ? A.run_Ti[r] : 1.0;
        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
        unsigned nacc = 0;
        for (int sweep = 0; sweep < A.n_sweeps && an_still; sweep++) {
            if (anneal && an_done >= A.n_total) break;
            const int64_t g = A.sweep_first + sweep;               // 1-based global sweep index (run modes)
            if (run == 1) {                                        // geometric ramp T_i -> T over the first n_anneal sweeps (MonteCarlo.jl:24-27)
                double t = run_T;
                if (g <= A.n_anneal && A.n_anneal > 1) t = run_Ti * pow(run_T / run_Ti, (double)(g - 1) / (double)(A.n_anneal - 1));
                beta = (R)(1.0 / t);
                nbeta = neg_beta_scaled(beta);
            }
            const uint64_t visit_base = anneal ? (uint64_t)an_done * (uint64_t)A.hits
                                        : run  ? (uint64_t)(g - 1) * (uint64_t)A.hits
                                               : A.visit0 + (uint64_t)sweep * (uint64_t)A.hits;
            unsigned sweep_acc = 0;
            for (int c = 0; c < M.n_col; c++) {
                const int c0 = A.cta_col_off[rank][c], cnt = A.cta_col_off[rank][c + 1] - c0;
                for (int s = tid; s < cnt; s += nth) {
                    const int q = c0 + s;
                    R re[D], im[D], T[NG], Tsq[NG], h[NG];
                    load_psi<R, D>(sPsi + q, (size_t)cap, re, im);
                    const int site = M.orig[my_off + q];
                    {
                        R P[NP];
#pragma unroll
                        for (int t = 0; t < NP; t++) P[t] = R(0);
#pragma unroll 3
                        for (int k = 0; k < M.z; k++) {
                            const unsigned code = sNbr[k * nmax + q];
                            uint32_t addr = sPsi_u32 + (code & 0x1FFFu) * (uint32_t)sizeof(V);
                            if (CLUSTER) addr = map_to_rank(addr, code >> 13);
                            R nr[D], ni[D];
#pragma unroll
                            for (int v = 0; v < NPV; v++) {
                                const V w4 = ld_v4<CLUSTER>(addr + v * plane_bytes, R(0));
                                nr[2 * v] = w4.x; ni[2 * v] = w4.y; nr[2 * v + 1] = w4.z; ni[2 * v + 1] = w4.w;
                            }
                            GenOps<GEN>::density_acc(nr, ni, P);
                        }
                        if (!fused) {
                            R S[1][NG];
                            GenOps<GEN>::expect_from_density(P, S[0]);
                            field_from_sums_smem<1>(sJt, S, h);
                        } else if constexpr (NP == NG) {
                            field_from_sums_smem<1>(sJt, (const R(*)[NG])P, h);
                        }
                    }
                    unsigned mine = 0;
                    if (ONSITE) {
                        GenOps<GEN>::expect(re, im, T);
                        GenOps<GEN>::expect_sq(re, im, Tsq);
                        R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
                        for (int hit = 0; hit < A.hits; hit++) {
                            R xr[D], xi[D], u, dE;
                            stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), visit_base + hit, xr, xi, u);
                            mine += metropolis_hit<R, GEN, ONSITE>(re, im, T, Tsq, h, M.cp.K, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
                        }
                    } else {
                        R Hq[GenOps<GEN>::NDENS];
                        if (fused) {
#pragma unroll
                            for (int t = 0; t < (NP == NG ? NG : 0); t++) Hq[t] = h[t];
                        } else {
                            GenOps<GEN>::hq_from_field(h, Hq);
                        }
                        R eloc = GenOps<GEN>::energy_from_state(re, im, Hq);
                        for (int hit = 0; hit < A.hits; hit++) {
                            R xr[D], xi[D], u, dE;
                            stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), visit_base + hit, xr, xi, u);
                            mine += metropolis_hit_psi<R, GEN>(re, im, Hq, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
                        }
                    }
                    if (mine) store_psi<R, D>(sPsi + q, (size_t)cap, re, im);
                    sweep_acc += mine;
                }
                if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
            }
            if (run) {
                const unsigned wt = __reduce_add_sync(0xffffffffu, sweep_acc);
                if ((tid & 31) == 0 && wt) atomicAdd(&s_acc, wt);
                __syncthreads();
                run_acc += (double)s_acc;
                __syncthreads();
                if (tid == 0) s_acc = 0;
                if (g % A.rate == 0) {
                    if (run == 1) {                                // sigma <- min(sigma * 0.5 / (1 - R), 60)  (MonteCarlo.jl:50-53)
                        double acc_all = run_acc;
                        if (CLUSTER) {                              // accepted updates of the whole replica: every CTA adds the shares in rank order
                            if (tid == 0) s_xchg[0] = run_acc;
                            cg::this_cluster().sync();
                            acc_all = 0.0;
                            for (int b = 0; b < A.cs; b++) acc_all += ld_f64<CLUSTER>(map_to_rank(smem_u32(&s_xchg[0]), (uint32_t)b));
                            cg::this_cluster().sync();
                        }
                        const double ratio = acc_all / A.att_window;
                        sigma = (R)fmin((double)sigma * 0.5 / (1.0 - ratio), 60.0);
                        run_acc = 0.0;
                    }
                    // E and sum_i T_i straight from the resident states (measure!, ObservablesGeneric.jl:32-46; the thermalisation
                    // energy log, MonteCarlo.jl:48-49): run == 2 writes a measure! row, run == 1 logs E/N before sigma is adapted
                    if (run == 2 || A.rows != nullptr) {
                        double eacc = 0.0, macc[NG];      // FP64 partial sums (a thread owns only a few sites)
#pragma unroll
                        for (int t = 0; t < NG; t++) macc[t] = 0.0;
                        for (int q = tid; q < n_own; q += nth) {
                            R re[D], im[D], h[NG], P[NP];
                            load_psi<R, D>(sPsi + q, (size_t)cap, re, im);
#pragma unroll
                            for (int t = 0; t < NP; t++) P[t] = R(0);
                            for (int k = 0; k < M.z; k++) {
                                const unsigned code = sNbr[k * nmax + q];
                                uint32_t addr = sPsi_u32 + (code & 0x1FFFu) * (uint32_t)sizeof(V);
                                if (CLUSTER) addr = map_to_rank(addr, code >> 13);
                                R nr[D], ni[D];
#pragma unroll
                                for (int v = 0; v < NPV; v++) {
                                    const V w4 = ld_v4<CLUSTER>(addr + v * plane_bytes, R(0));
                                    nr[2 * v] = w4.x; ni[2 * v] = w4.y; nr[2 * v + 1] = w4.z; ni[2 * v + 1] = w4.w;
                                }
                                GenOps<GEN>::density_acc(nr, ni, P);
                            }
                            if (!fused) {
                                R S[1][NG];
                                GenOps<GEN>::expect_from_density(P, S[0]);
                                field_from_sums_smem<1>(sJt, S, h);
                                R T[NG];
                                GenOps<GEN>::expect(re, im, T);
                                R e = R(0);
#pragma unroll
                                for (int t = 0; t < NG; t++) e = fma(T[t], h[t], e);
                                e *= R(0.5);
                                if (ONSITE) {
                                    R Tsq[NG];
                                    GenOps<GEN>::expect_sq(re, im, Tsq);
#pragma unroll
                                    for (int t = 0; t < NG; t++) e = fma(Tsq[t], M.cp.K[t], e);
                                }
                                eacc += (double)e;
#pragma unroll
                                for (int t = 0; t < NG; t++) macc[t] += (double)T[t];
                            } else if constexpr (NP == NG) {
                                field_from_sums_smem<1>(sJt, (const R(*)[NG])P, h);           // h = Hq
                                eacc += 0.5 * (double)GenOps<GEN>::energy_from_state(re, im, h);
                                R Pown[NP];
#pragma unroll
                                for (int t = 0; t < NP; t++) Pown[t] = R(0);
                                GenOps<GEN>::density_acc(re, im, Pown);
#pragma unroll
                                for (int t = 0; t < NG; t++) macc[t] += (double)Pown[t];
                            }
                        }
                        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
                        for (int t = 0; t <= NG; t++) {
                            double v = t < NG ? macc[t < NG ? t : 0] : eacc;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0) s_red[warp][t] = v;
                        }
                        __syncthreads();
                        if (CLUSTER) {
                            if (tid <= NG) {
                                double v = 0.0;
                                for (int w = 0; w < (nth >> 5); w++) v += s_red[w][tid];
                                s_xchg[tid] = v;
                            }
                            cg::this_cluster().sync();
                        }
                        if (tid == 0 && rank == 0) {
                            double tot[NG + 1], Tsum[NG];
                            for (int t = 0; t <= NG; t++) {
                                double v = 0.0;
                                if (CLUSTER) {
                                    for (int b = 0; b < A.cs; b++) v += ld_f64<CLUSTER>(map_to_rank(smem_u32(&s_xchg[t]), (uint32_t)b));
                                } else {
                                    for (int w = 0; w < (nth >> 5); w++) v += s_red[w][t];
                                }
                                tot[t] = v;
                            }
                            if (fused) GenOps<GEN>::expect_from_density(tot, Tsum);           // magnetisation is linear in the density sums
                            else for (int t = 0; t < NG; t++) Tsum[t] = tot[t];
                            const double E = tot[NG] + A.e_const, invN = 1.0 / (double)M.N;
                            const size_t slot = (size_t)(g / A.rate - A.row_base) * A.nrep + r;
                            if (run == 1) {
                                A.rows[slot] = E * invN;
                            } else {
                                double* row = A.rows + slot * SCMC_NOBS;
                                double m[NG];
                                for (int t = 0; t < NG; t++) m[t] = Tsum[t] * invN;
                                row[0] = E * invN;
                                row[1] = (E * invN) * (E * invN);
                                row[2] = sqrt(m[4] * m[4] + m[8] * m[8] + m[12] * m[12]);
                                row[3] = sqrt(m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
                                row[4] = sqrt(m[5] * m[5] + m[6] * m[6] + m[7] * m[7] + m[9] * m[9] + m[10] * m[10] + m[11] * m[11] + m[13] * m[13] +
                                              m[14] * m[14] + m[15] * m[15]);
                                for (int t = 0; t < NG; t++) row[5 + t] = m[t];
                            }
                        }
                        if (CLUSTER) cg::this_cluster().sync(); else __syncthreads();
                    }
                }
            } else if (!anneal) {
                nacc += sweep_acc;
            } else {
                // accepted updates of this sweep over the whole replica, then the cooling rule (MonteCarlo.jl:151-203)
                const unsigned wt = __reduce_add_sync(0xffffffffu, sweep_acc);
                if ((tid & 31) == 0 && wt) atomicAdd(&s_acc, wt);
                __syncthreads();
                double sweep_all = (double)s_acc;
                if (CLUSTER) {                                      // accepted updates of the whole replica in this sweep, added in rank order
                    if (tid == 0) s_xchg[0] = (double)s_acc;
                    cg::this_cluster().sync();
                    sweep_all = 0.0;
                    for (int b = 0; b < A.cs; b++) sweep_all += ld_f64<CLUSTER>(map_to_rank(smem_u32(&s_xchg[0]), (uint32_t)b));
                    cg::this_cluster().sync();
                }
                an_acc += sweep_all;
                __syncthreads();
                if (tid == 0) s_acc = 0;
                an_sat++; an_done++;
                const bool last = an_done >= A.n_total;
                if (!(an_acc < A.min_accepted && an_sat < A.n_per_T && !last)) {
                    const double ratio = an_acc / (A.att_per_sweep * (double)an_sat);
                    if (ratio > A.min_accrate) {
                        beta = (R)((double)beta * A.beta_fac);
                        sigma = (R)fmax(fmin((double)sigma * 0.5 / (1.0 - ratio), 60.0), A.sigma_min);
                        nbeta = neg_beta_scaled(beta);
                        an_acc = 0.0; an_sat = 0;
                    } else {
                        an_still = false;
                    }
                }
                __syncthreads();
            }
        }
        for (int q = tid; q < n_own; q += nth) {
#pragma unroll
            for (int v = 0; v < NPV; v++) M.psi[gbase + q + v * ps] = sPsi[v * cap + q];
        }
        if (run) {
            double acc_all = run_acc;
            if (CLUSTER) {
                if (tid == 0) s_xchg[0] = run_acc;
                cg::this_cluster().sync();
                acc_all = 0.0;
                for (int b = 0; b < A.cs; b++) acc_all += ld_f64<CLUSTER>(map_to_rank(smem_u32(&s_xchg[0]), (uint32_t)b));
                cg::this_cluster().sync();
            }
            if (tid == 0 && rank == 0) {
                ((R*)A.sigma)[r] = sigma;
                A.acc[r] = (unsigned long long)acc_all;
            }
            __syncthreads();
            continue;
        }
        if (anneal) {
            if (tid == 0 && rank == 0) {
                ((R*)A.beta)[r] = beta; ((R*)A.sigma)[r] = sigma;
                A.acc[r] = (unsigned long long)an_acc;
                A.sweeps_at_T[r] = an_sat; A.sweeps_done[r] = an_done;
                const bool running = an_still && an_done < A.n_total;
                A.active_rw[r] = an_still ? 1 : 0;
                if (running) atomicAdd(A.n_active, 1);
            }
            __syncthreads();
            continue;
        }
        const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
        if ((tid & 31) == 0 && tot) atomicAdd(&s_acc, tot);
        __syncthreads();
        if (tid == 0) {
            if (s_acc) atomicAdd(A.acc + r, (unsigned long long)s_acc);
            s_acc = 0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <int V> using ic = std::integral_constant<int, V>;
template <class R, int NL>
static DevModel<R, NL> make_model_r(const scmc_handle* h) {
    DevModel<R, NL> M;
    memset(&M, 0, sizeof M);
    M.psi = (vec4<R>*)h->psi; M.T = (vec4<R>*)h->T; M.Tsq = (vec4<R>*)h->Tsq;
    M.nbr = h->nbr; M.lab = h->lab; M.orig = h->d_orig; M.col_list = h->d_col_list; M.cls_site = h->d_cls_site; M.cls_nbr = h->d_cls_nbr;
    M.N = (int32_t)h->N; M.Ns = (int32_t)h->Ns; M.z = h->z; M.n_col = h->n_col; M.nrep = h->R;
    for (int c = 0; c <= MAXC; c++) M.col_off[c] = h->col_off[c];
    for (int l = 0; l < NL; l++)
        for (int t = 0; t < 256; t++) M.cp.J[l][t] = l < h->n_labels ? (R)h->J[l][t] : R(0);
    for (int a = 0; a < NG; a++) M.cp.K[a] = (R)h->K[a];
    return M;
}
template <class F>
static int32_t dispatch_r(const scmc_handle* h, F&& f) {
    auto with_nl = [&](auto rt, auto gen, auto ons) -> int32_t {
        if (h->n_labels == 1) return f(rt, gen, ons, ic<1>{});
        return f(rt, gen, ons, ic<MAXL>{});
    };
    auto with_gen = [&](auto rt) -> int32_t {
        switch (h->gen_id) {
            case 1: return with_nl(rt, ic<1>{}, ic<0>{});
            case 2: return h->onsite ? with_nl(rt, ic<2>{}, ic<1>{}) : with_nl(rt, ic<2>{}, ic<0>{});
            case 3: return with_nl(rt, ic<3>{}, ic<0>{});
        }
        return SCMC_ERR_ARG;
    };
    return h->prec == 64 ? with_gen(double{}) : with_gen(float{});
}

template <class K>
static int32_t configure_and_count(K kernel, scmc_handle* h, size_t smem, int* max_clusters) {
    auto& P = h->res;
    if (smem > (size_t)227 * 1024) return SCMC_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return SCMC_ERR_UNSUPPORTED; }
    int n = 0;
    if (P.cs > 1) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(P.cs * 148); cfg.blockDim = dim3(P.nthreads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = P.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return SCMC_ERR_UNSUPPORTED; }
    } else {
        int per_sm = 0, dev = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, P.nthreads, smem) != cudaSuccess) { cudaGetLastError(); return SCMC_ERR_UNSUPPORTED; }
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        n = per_sm * sms;
    }
    *max_clusters = n;
    return n > 0 ? SCMC_OK : SCMC_ERR_UNSUPPORTED;
}

// Called at the end of scmc_create: neighbour codes, chunk tables, launch geometry.  Leaves res.eligible = 0 when the lattice
// does not fit (the streaming kernel is then the only path).
void plan_resident(scmc_handle* h, const std::vector<int>& owner_of_perm) {
    auto& P = h->res;
    P.eligible = 0;
    if (P.cs == 0) return;
    const int64_t N = h->N;
    const int cs = P.cs;
    std::vector<int> n_own(cs, 0);
    for (int64_t q = 0; q < N; q++) n_own[owner_of_perm[q]]++;
    P.cta_off[0] = 0;
    for (int b = 0; b < cs; b++) P.cta_off[b + 1] = P.cta_off[b] + n_own[b];
    P.n_own_max = *std::max_element(n_own.begin(), n_own.end());
    P.cap = (P.n_own_max + 1 + 7) / 8 * 8;
    int max_cnt = 1;
    for (int b = 0; b < cs; b++) {
        int at = 0;
        for (int c = 0; c < h->n_col; c++) {
            P.cta_col_off[b][c] = at;
            int cnt = 0;
            for (int q = P.cta_off[b]; q < P.cta_off[b + 1]; q++) cnt += (h->colour[h->orig[q]] == c);
            at += cnt;
            max_cnt = std::max(max_cnt, cnt);
        }
        for (int c = h->n_col; c <= MAXC; c++) P.cta_col_off[b][c] = at;
    }
    const int rounds = (max_cnt + 511) / 512;
    P.nthreads = std::min(512, std::max(32, ((max_cnt + rounds - 1) / rounds + 31) / 32 * 32));
    const size_t w = h->esz();
    P.smem = (size_t)P.cap * 4 * w * (4 + h->d / 2 + (h->onsite ? 4 : 0)) + (size_t)MAXL * 256 * w + (size_t)h->z * P.n_own_max * 2 +
             (h->n_labels > 1 ? (size_t)h->z * P.n_own_max : 0) + 64;
    P.smem = (P.smem + 15) / 16 * 16;
    P.smem_psi = (size_t)P.cap * 4 * w * (h->d / 2) + (size_t)256 * w + (size_t)h->z * P.n_own_max * 2 + 64;
    P.smem_psi = (P.smem_psi + 15) / 16 * 16;
    // neighbour codes in permuted order: owner CTA << 13 | local index; padding slots -> my own zero slot (local index n_own_max)
    std::vector<uint16_t> code((size_t)h->z * N);
    std::vector<int32_t> nb((size_t)h->z * N);
    if (cudaMemcpy(nb.data(), h->nbr, nb.size() * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return; }
    for (int k = 0; k < h->z; k++)
        for (int64_t q = 0; q < N; q++) {
            const int32_t j = nb[(size_t)k * N + q];
            if (j >= N) code[(size_t)k * N + q] = (uint16_t)((owner_of_perm[q] << 13) | P.n_own_max);
            else code[(size_t)k * N + q] = (uint16_t)((owner_of_perm[j] << 13) | (j - P.cta_off[owner_of_perm[j]]));
        }
    if (dev_alloc(h, (void**)&h->d_nbr16, code.size() * sizeof(uint16_t)) != SCMC_OK) return;
    if (cudaMemcpy(h->d_nbr16, code.data(), code.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return; }
    int maxc = 0, maxc_psi = 0;
    const int32_t st = dispatch_r(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        if (cs > 1) return configure_and_count(k_resident<R, GEN, ONS, NL, true>, h, P.smem, &maxc);
        return configure_and_count(k_resident<R, GEN, ONS, NL, false>, h, P.smem, &maxc);
    });
    if (st == SCMC_OK) { P.max_clusters = maxc; P.eligible = 1; }
    if (h->n_labels == 1) {
        const int32_t st2 = dispatch_r(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
            using R = decltype(rt);
            constexpr int GEN = decltype(gen)::value;
            constexpr bool ONS = decltype(ons)::value != 0;
            if (cs > 1) return configure_and_count(k_resident_psi<R, GEN, ONS, true>, h, P.smem_psi, &maxc_psi);
            return configure_and_count(k_resident_psi<R, GEN, ONS, false>, h, P.smem_psi, &maxc_psi);
        });
        if (st2 == SCMC_OK) { P.max_clusters_psi = maxc_psi; P.eligible_psi = 1; }
    }
}

int32_t launch_resident(scmc_handle* h, int64_t n_sweeps, int32_t hits, uint64_t seed, uint64_t sweep_counter, const uint8_t* active, bool psi) {
    auto& P = h->res;
    if (n_sweeps <= 0) return SCMC_OK;
    if (psi) h->T_stale = true;
    return dispatch_r(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        auto M = make_model_r<R, NL>(h);
        ResidentArgs A;
        memset(&A, 0, sizeof A);
        A.nrep = h->R; A.roff = h->roff; A.hits = hits; A.cs = P.cs; A.n_own_max = P.n_own_max; A.cap = P.cap;
        A.seed = seed; A.stream = h->stream_version; A.beta = h->beta; A.sigma = h->sigma; A.Jt = h->d_Jt; A.Jqt = h->fused_density ? h->d_Jqt : nullptr; A.active = active; A.acc = h->acc; A.nbr16 = h->d_nbr16;
        memcpy(A.cta_off, P.cta_off, sizeof A.cta_off);
        memcpy(A.cta_col_off, P.cta_col_off, sizeof A.cta_col_off);
        A.n_clusters = (int32_t)std::min<int64_t>(h->R, psi ? P.max_clusters_psi : P.max_clusters);
        const size_t smem = psi ? P.smem_psi : P.smem;
        // sweeps per launch bounded so that visit numbers stay in int32 loop counters and launches stay preemptible
        for (int64_t done = 0; done < n_sweeps;) {
            const int64_t chunk = std::min<int64_t>(n_sweeps - done, 4096);
            A.n_sweeps = (int32_t)chunk;
            A.visit0 = (sweep_counter + (uint64_t)done) * (uint64_t)hits;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(A.n_clusters * P.cs)); cfg.blockDim = dim3(P.nthreads); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
            cudaLaunchAttribute at[1];
            cfg.attrs = at; cfg.numAttrs = 0;
            if (P.cs > 1) {
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = P.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.numAttrs = 1;
                if (psi) { if constexpr (NL == 1) SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident_psi<R, GEN, ONS, true>, M, A)); }
                else SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident<R, GEN, ONS, NL, true>, M, A));
            } else {
                if (psi) { if constexpr (NL == 1) SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident_psi<R, GEN, ONS, false>, M, A)); }
                else SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident<R, GEN, ONS, NL, false>, M, A));
            }
            h->launches++;
            done += chunk;
        }
        return SCMC_OK;
    });
}


// Launch of the persistent kernel of one arithmetic family with a filled ResidentArgs (cluster attribute when a replica spans several CTAs)
template <class R, int GEN, bool ONS, int NL>
static int32_t launch_family_kernel(scmc_handle* h, const DevModel<R, NL>& M, ResidentArgs& A, bool psi) {
    auto& P = h->res;
    A.n_clusters = (int32_t)std::min<int64_t>(h->R, psi ? P.max_clusters_psi : P.max_clusters);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(A.n_clusters * P.cs)); cfg.blockDim = dim3(P.nthreads); cfg.dynamicSmemBytes = psi ? P.smem_psi : P.smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    cfg.attrs = at; cfg.numAttrs = 0;
    if (P.cs > 1) {
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = P.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.numAttrs = 1;
    }
    if (psi) {
        if constexpr (NL == 1) {
            if (P.cs > 1) SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident_psi<R, GEN, ONS, true>, M, A));
            else SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident_psi<R, GEN, ONS, false>, M, A));
        } else {
            return SCMC_ERR_UNSUPPORTED;
        }
    } else {
        if (P.cs > 1) SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident<R, GEN, ONS, NL, true>, M, A));
        else SCMC_CK(cudaLaunchKernelEx(&cfg, k_resident<R, GEN, ONS, NL, false>, M, A));
    }
    h->launches++;
    return SCMC_OK;
}

// Device-side cooling loop for single-CTA psi-family replicas: up to `chunk` sweeps per launch, controller inside the kernel.
// Returns through *n_active how many replicas are still cooling and have budget left.
int32_t launch_resident_anneal(scmc_handle* h, int64_t chunk, int32_t hits, uint64_t seed, double att_per_sweep, double min_accepted,
                               double beta_fac, double min_accrate, double sigma_min, int64_t n_per_T, int64_t n_total,
                               int64_t* d_sweeps_at_T, int64_t* d_sweeps_done, int* d_n_active) {
    auto& P = h->res;
    const bool psi = uses_psi_gather(h);
    if (psi) h->T_stale = true; else SCMC_TRY(ensure_T(h));      // the T-gather kernel keeps the T (Tsq) planes current
    return dispatch_r(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        {
            auto M = make_model_r<R, NL>(h);
            ResidentArgs A;
            memset(&A, 0, sizeof A);
            A.nrep = h->R; A.roff = h->roff; A.hits = hits; A.cs = P.cs; A.n_own_max = P.n_own_max; A.cap = P.cap;
            A.seed = seed; A.stream = h->stream_version; A.beta = h->beta; A.sigma = h->sigma; A.Jt = h->d_Jt; A.Jqt = h->fused_density ? h->d_Jqt : nullptr;
            A.active = h->active; A.acc = h->acc; A.nbr16 = h->d_nbr16;
            memcpy(A.cta_off, P.cta_off, sizeof A.cta_off);
            memcpy(A.cta_col_off, P.cta_col_off, sizeof A.cta_col_off);
            A.n_sweeps = (int32_t)chunk; A.visit0 = 0;
            A.anneal = 1; A.att_per_sweep = att_per_sweep; A.min_accepted = min_accepted; A.beta_fac = beta_fac; A.min_accrate = min_accrate;
            A.sigma_min = sigma_min; A.n_per_T = n_per_T; A.n_total = n_total; A.sweeps_at_T = d_sweeps_at_T; A.sweeps_done = d_sweeps_done;
            A.active_rw = h->active; A.n_active = d_n_active;
            return launch_family_kernel<R, GEN, ONS, NL>(h, M, A, psi);
        }
    });
}


// One chunk of the device-side run! schedule (either family, any cluster size): mode 1 = thermalisation, 2 = measurement.
int32_t launch_resident_run(scmc_handle* h, int mode, int64_t chunk, int32_t hits, uint64_t seed, const double* d_T, const double* d_Ti,
                            int64_t n_anneal, int64_t rate, int64_t sweep_first, int64_t row_base, double* d_rows) {
    auto& P = h->res;
    const bool psi = uses_psi_gather(h);
    if (psi) h->T_stale = true; else SCMC_TRY(ensure_T(h));      // the T-gather kernel keeps the T (Tsq) planes current
    return dispatch_r(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        {
            auto M = make_model_r<R, NL>(h);
            ResidentArgs A;
            memset(&A, 0, sizeof A);
            A.nrep = h->R; A.roff = h->roff; A.hits = hits; A.cs = P.cs; A.n_own_max = P.n_own_max; A.cap = P.cap;
            A.seed = seed; A.stream = h->stream_version; A.beta = h->beta; A.sigma = h->sigma; A.Jt = h->d_Jt; A.Jqt = h->fused_density ? h->d_Jqt : nullptr;
            A.active = nullptr; A.acc = h->acc; A.nbr16 = h->d_nbr16;
            memcpy(A.cta_off, P.cta_off, sizeof A.cta_off);
            memcpy(A.cta_col_off, P.cta_col_off, sizeof A.cta_col_off);
            A.n_sweeps = (int32_t)chunk;
            A.run_mode = mode; A.run_T = d_T; A.run_Ti = d_Ti; A.n_anneal = n_anneal; A.rate = rate; A.sweep_first = sweep_first; A.row_base = row_base;
            A.att_window = (double)hits * (double)h->N * (double)rate; A.e_const = h->e_const; A.rows = d_rows;
            return launch_family_kernel<R, GEN, ONS, NL>(h, M, A, psi);
        }
    });
}

}  // namespace scmc
