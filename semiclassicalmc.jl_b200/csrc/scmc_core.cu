// scmc_core.cu -- Configuration storage, streaming colour-pass Metropolis kernel, energy / measurement reductions
// and the C-ABI entry points around them (include/scmc.h).  sm_100a only; no CPU fallback.
//
// Reference functions replaced here (paths relative to /root/reference/):
//   k_sweep_colour   <- the 2N-update loops + localUpdate!           src/MonteCarlo.jl:40-44, 76-80, 152-156, 263-286
//   k_measure        <- getEnergy, getMagnetization, measure!         src/Configuration.jl:236-253, src/Observables/Observables.jl:16-18,
//                                                                     src/Observables/ObservablesGeneric.jl:32-46
//   k_local_field    <- inner sum of getEnergyDifference!             src/Configuration.jl:292-296
//   k_pack / k_unpack / k_randomize <- Configuration constructor      src/Configuration.jl:62-66, src/MonteCarlo.jl:244-246
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "handle.h"

namespace scmc {

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    return SCMC_ERR_CUDA;
}

int32_t dev_alloc(scmc_handle* h, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        g_err = buf;
        return SCMC_ERR_ALLOC;
    }
    h->bytes += (int64_t)bytes;
    return SCMC_OK;
}

// ------------------------------------------------------------------------------------------------ kernels
#ifndef SCMC_SWEEP_THREADS
#define SCMC_SWEEP_THREADS 128
#define SCMC_SWEEP_BLOCKS 5
#endif
// resident CTAs per SM the FP64 instantiations are compiled for: at 5 (96 registers) they spill 0.4-2.4 KB per thread.  Measured on
// B200 (C5 lattice, 2048 replicas, FP64): T-gather 0.92 / 1.10 / 1.21e10 updates/s at 5 / 4 / 3 CTAs, psi-gather 1.39 / 1.60 / 1.45e10
#ifndef SCMC_SWEEP_BLOCKS_F64
#define SCMC_SWEEP_BLOCKS_F64 3
#endif
#ifndef SCMC_SWEEP_BLOCKS_MMA
#define SCMC_SWEEP_BLOCKS_MMA 8                     // the tensor-core mat-vec experiment (SCMC_MMA=1): 64 registers, no spills; measured best of 4 ... 9
#endif
#ifndef SCMC_SWEEP_BLOCKS_NL
#define SCMC_SWEEP_BLOCKS_NL SCMC_SWEEP_BLOCKS      // FP32 instantiations with several coupling labels
#endif
#ifndef SCMC_SWEEP_PSI_BLOCKS_F64
#define SCMC_SWEEP_PSI_BLOCKS_F64 4
#endif
#ifndef SCMC_GATHER6
#define SCMC_GATHER6 0        // experiment: the FFMA psi-gather kernel with z = 6 issues all six neighbours' loads before the first density product
#endif
#ifndef SCMC_HIT_UNROLL2
#define SCMC_HIT_UNROLL2 0
#endif
#ifndef SCMC_ABLATE
#define SCMC_ABLATE 0            // performance ablations of the direct streaming kernel (profiles/README.md); 0 = production
#endif
#ifndef SCMC_ASYNC_THREADS
#define SCMC_ASYNC_THREADS 128
#define SCMC_ASYNC_BLOCKS 4
#endif
#ifndef SCMC_ASYNC_CP
#define SCMC_ASYNC_CP "cp.async.cg.shared.global"
#endif
#ifndef SCMC_ASYNC_TILES
#define SCMC_ASYNC_TILES 11      // visits per thread in the cp.async-pipelined kernel
#endif
#ifndef SCMC_SWEEP_PREFETCH
#define SCMC_SWEEP_PREFETCH 0  // measured on B200: L2 prefetch of the next tile costs more issue slots than it hides (2.57e10 vs 2.90e10 upd/s)
#endif
#ifndef SCMC_SWEEP_TILES
#define SCMC_SWEEP_TILES 16     // site tiles (of SCMC_SWEEP_THREADS sites) a block walks over
#endif
#define SCMC_MEASURE_THREADS 128
// struct SweepArgs: see handle.h (shared with scmc_tma.cu)

// S_l = sum of T_j over the neighbours with label l.  The neighbour table is padded to a multiple of 3 slots; padding slots
// (and missing bonds) point at site N, whose T planes are identically zero, so the loop is branch-free and every chunk
// issues 3 index loads followed by 12 independent 128-bit loads (memory-level parallelism instead of 2 serial latencies per bond).
// cp.async (LDGSTS) staging used by the software-pipelined kernels
constexpr int ASYNC_Z = 6;
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile(SCMC_ASYNC_CP " [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

template <class R, int NL>
__device__ __forceinline__ void neighbour_sums(const DevModel<R, NL>& M, int64_t rbase, size_t ps, int p, R (*S)[NG]) {
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
        for (int a = 0; a < NG; a++) S[l][a] = R(0);
    const int32_t* nb = M.nbr + p;
    const int8_t* lb = M.lab + p;
    // opaque 64-bit bases of the replica's four T planes: a neighbour load is then ONE mad.wide (base + 16 * index)
    unsigned long long Tq[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        Tq[q] = (unsigned long long)(M.T + rbase + q * ps);
        asm volatile("" : "+l"(Tq[q]));
    }
    const unsigned stride = (unsigned)M.N;
#pragma unroll 2
    for (int k = 0; k < M.z; k += 3) {
        unsigned j[3];
        int l[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            j[t] = (unsigned)nb[(unsigned)(k + t) * stride];
#if SCMC_ABLATE == 2
            j[t] = (unsigned)p;      // ablation: neighbour = self (no gather traffic)
#endif
            l[t] = NL == 1 ? 0 : lb[(unsigned)(k + t) * stride];
        }
        vec4<R> v[3][4];
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int q = 0; q < 4; q++) v[t][q] = *reinterpret_cast<const vec4<R>*>(Tq[q] + (unsigned long long)j[t] * sizeof(vec4<R>));
#pragma unroll
        for (int t = 0; t < 3; t++) {
#pragma unroll
            for (int ll = 0; ll < NL; ll++)
                if (NL == 1 || l[t] == ll) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        S[ll][4 * q] += v[t][q].x; S[ll][4 * q + 1] += v[t][q].y; S[ll][4 * q + 2] += v[t][q].z; S[ll][4 * q + 3] += v[t][q].w;
                    }
                }
        }
    }
}

// Local field of a several-label model with site-independent slot ranges per label: one label at a time, its neighbour sum in 16
// registers (not in a run-time indexed S[NL][16], which the compiler places in local memory: profiles/README.md), fed straight into that
// label's mat-vec.  Loads, additions and FMAs are those of neighbour_sums + field_from_sums_smem in the same order.
template <class R, int NL>
__device__ __forceinline__ void field_label_major(const DevModel<R, NL>& M, const SweepArgs& A, const R* sJt, int64_t rbase, size_t ps, int p, R* h) {
    const int32_t* nb = M.nbr + p;
    unsigned long long Tq[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        Tq[q] = (unsigned long long)(M.T + rbase + q * ps);
        asm volatile("" : "+l"(Tq[q]));
    }
    const unsigned stride = (unsigned)M.N;
    constexpr bool F32 = sizeof(R) == 4;
    unsigned long long hp[NG / 2];
    if (F32) {
#pragma unroll
        for (int i = 0; i < NG / 2; i++) hp[i] = 0ull;
    } else {
#pragma unroll
        for (int a = 0; a < NG; a++) h[a] = R(0);
    }
#pragma unroll
    for (int l = 0; l < NL; l++) {
        R S[NG];
#pragma unroll
        for (int a = 0; a < NG; a++) S[a] = R(0);
        const int k1 = A.lab_off[l + 1];
        for (int k = A.lab_off[l]; k < k1; k += 3) {
            unsigned j[3];
#pragma unroll
            for (int t = 0; t < 3; t++) j[t] = (k + t < k1) ? (unsigned)nb[(unsigned)(k + t) * stride] : stride;      // past the label's range: zero site
            vec4<R> v[3][4];
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int q = 0; q < 4; q++) v[t][q] = *reinterpret_cast<const vec4<R>*>(Tq[q] + (unsigned long long)j[t] * sizeof(vec4<R>));
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    S[4 * q] += v[t][q].x; S[4 * q + 1] += v[t][q].y; S[4 * q + 2] += v[t][q].z; S[4 * q + 3] += v[t][q].w;
                }
        }
        if constexpr (F32) field_accumulate_smem((const float*)sJt + l * NG * NG, (const float*)S, hp);
        else field_accumulate_smem((const double*)sJt + l * NG * NG, (const double*)S, (double*)h);
    }
    if constexpr (F32) {
#pragma unroll
        for (int i = 0; i < NG / 2; i++) asm("mov.b64 {%0, %1}, %2;" : "=f"(h[2 * i]), "=f"(h[2 * i + 1]) : "l"(hp[i]));
    }
}

// Software prefetch of everything the NEXT site tile of this thread will read (psi planes, neighbour T planes) into L2:
// no registers are held across the hit loop, and the DRAM round trip of the next tile overlaps with the current tile's hits.
template <class R, int NL, int D>
__device__ __forceinline__ void prefetch_site(const DevModel<R, NL>& M, int64_t rbase, size_t ps, int p) {
#pragma unroll
    for (int v = 0; v < D / 2; v++) asm volatile("prefetch.global.L2 [%0];" ::"l"(M.psi + rbase + p + v * ps));
    const int32_t* nb = M.nbr + p;
    const unsigned stride = (unsigned)M.N;
    for (int k = 0; k < M.z; k++) {
        const unsigned j = (unsigned)nb[(unsigned)k * stride];
#pragma unroll
        for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(M.T + rbase + j + q * ps));
    }
}

// One colour pass: every (replica, site of this colour) gets `hits` consecutive Metropolis updates.
// grid = (blocks per replica, replicas of the batch); a block walks over tiles of blockDim.x sites of its replica so that the
// coupling table (global -> shared, coalesced) is staged once per block.  INJECT = deterministic test mode (host-supplied randoms).
template <class R, int GEN, bool ONSITE, int NL, bool INJECT>
__global__ void __launch_bounds__(SCMC_SWEEP_THREADS, sizeof(R) == 8 ? SCMC_SWEEP_BLOCKS_F64 : (NL > 1 ? SCMC_SWEEP_BLOCKS_NL : SCMC_SWEEP_BLOCKS)) k_sweep_colour(const __grid_constant__ DevModel<R, NL> M, const SweepArgs A) {
    constexpr int D = GenOps<GEN>::D;
    const int64_t r = A.r0 + blockIdx.y;
    __shared__ __align__(16) R sJt[NL * NG * NG];
    for (int t = threadIdx.x; t < NL * NG * NG; t += blockDim.x) sJt[t] = ((const R*)A.Jt)[t];
    const bool alive = (A.active == nullptr || A.active[r]);
    const R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
    const R nbeta = neg_beta_scaled(beta);
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    __syncthreads();
    unsigned nacc = 0;
    for (int colour = A.colour; colour <= A.colour_last; colour++) {
    const int nc = M.col_off[colour + 1] - M.col_off[colour];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nc && alive; s += gridDim.x * blockDim.x) {
        const int p = M.col_list[M.col_off[colour] + s];
#if SCMC_SWEEP_PREFETCH
        // indices of this thread's NEXT visit: loaded now, used for L2 prefetches after the mat-vec (no stall, no data registers)
        const int sn = s + gridDim.x * blockDim.x;
        int pn = 0;
        unsigned jn[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        if (sn < nc) {
            pn = M.col_list[M.col_off[colour] + sn];
#pragma unroll
            for (int k = 0; k < 6; k++) jn[k] = k < M.z ? (unsigned)M.nbr[(unsigned)k * (unsigned)M.N + pn] : 0u;
        }
#endif
        R re[D], im[D], T[NG], Tsq[NG], h[NG];
        load_psi<R, D>(M.psi + rbase + p, ps, re, im);
        const int site = M.orig[p];
        if (NL > 1 && A.lab_major) {
            field_label_major<R, NL>(M, A, sJt, rbase, ps, p, h);
        } else {
            R S[NL][NG];
            neighbour_sums<R, NL>(M, rbase, ps, p, S);
#if SCMC_ABLATE == 3
            for (int a = 0; a < NG; a++) h[a] = S[0][a];      // ablation: no mat-vec
#else
            field_from_sums_smem<NL>(sJt, S, h);
#endif
        }
        GenOps<GEN>::expect(re, im, T);
        if (ONSITE) GenOps<GEN>::expect_sq(re, im, Tsq);
        // e_loc = T.h (+ K.Tsq): dE of a proposal is e_loc' - e_loc, which needs one dot product per hit
        R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
#if SCMC_SWEEP_PREFETCH
        if (sn < nc) {
#pragma unroll
            for (int v = 0; v < D / 2; v++) asm volatile("prefetch.global.L2 [%0];" ::"l"(M.psi + rbase + pn + v * ps));
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (k < M.z) {
#pragma unroll
                    for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(M.T + rbase + jn[k] + q * ps));
                }
        }
#endif
        unsigned mine = 0;
#if SCMC_ABLATE == 4
        mine = (h[3] > R(1e30)) ? 1u : 0u;                    // ablation: no hits (keeps h alive)
        for (int hit = 0; hit < 0; hit++) {
#else
        for (int hit = 0; hit < A.hits; hit++) {
#endif
            R xr[D], xi[D], u, dE;
            if (INJECT) {
                const size_t t = ((size_t)hit * M.nrep + r) * M.N + site;
#pragma unroll
                for (int k = 0; k < D; k++) { xr[k] = (R)A.inj_xi[t * 2 * D + k]; xi[k] = (R)A.inj_xi[t * 2 * D + D + k]; }
                u = (R)A.inj_u[t];
            } else {
                stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
            }
            const bool ok = metropolis_hit<R, GEN, ONSITE>(re, im, T, Tsq, h, M.cp.K, xr, xi, u, nbeta, sigma, eloc, dE);
            if (INJECT) {
                const size_t t = ((size_t)hit * M.nrep + r) * M.N + site;
                A.out_acc[t] = ok ? 1 : 0;
                if (A.out_dE) A.out_dE[t] = (double)dE;
            }
            mine += ok ? 1u : 0u;
        }
#if SCMC_ABLATE == 1
        if (mine > 1000u) {                                   // ablation: no write-back
#else
        if (mine) {
#endif
            store_psi<R, D>(M.psi + rbase + p, ps, re, im);
            store_T<R>(M.T + rbase + p, ps, T);
            if (ONSITE) store_T<R>(M.Tsq + rbase + p, ps, Tsq);
        }
        nacc += mine;
    }
    if (colour < A.colour_last) __syncthreads();     // small seam classes: same block, next colour sees this colour's stores
    }
    // acceptance counters: one atomic per warp (a block never spans two replicas)
    const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(A.acc + r, (unsigned long long)tot);
}

// psi-gather variant of the colour pass: neighbours contribute through their STATE (d/2 x 16 bytes) instead of the cached T vector
// (64 bytes).  sum_j T_j = map(sum_j P_j) with P = conj(psi) psi^T, so the gather accumulates density-matrix entries with
// FMAs and maps once per visit.  Nothing but psi is read or written: DRAM traffic per visit drops from ~272 B to ~150 B and
// the T planes become a lazily refreshed cache (ensure_T).  Rounding differs from the T-gather kernels at the 1-ulp level.
// MMA = true (experiment, SCMC_MMA=1, FP32 / 16 density entries / no on-site term only): the 32 sites of a warp multiply the SAME Jq, so
// Hq[16 x 32] = Jq[16 x 16] . P[16 x 32] runs as 24 mma.sync.m16n8k8 in 3xTF32 (operands split into a 19-bit head and an exactly
// representable tail) with the density sums and the result transposed through a per-warp shared-memory tile.  Different rounding
// than the FFMA order of field_from_sums_smem, so it is NOT part of the bit-identity contract of the psi family.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int MMA_LD = 20;           // row stride (floats) of the padded tiles: fragment loads, tile rows and 128-bit row accesses are conflict-free
__device__ __forceinline__ void matvec_mma(const float* sAhi, const float* sAlo, float* tile, const float* P, float* h) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int q = 0; q < 4; q++) *reinterpret_cast<float4*>(tile + lane * MMA_LD + 4 * q) = make_float4(P[4 * q], P[4 * q + 1], P[4 * q + 2], P[4 * q + 3]);
    __syncwarp();
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[j][c] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 2; kh++) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int idx = (g + 8 * (c & 1)) * MMA_LD + 8 * kh + t + 4 * (c >> 1);
            ah[c] = __float_as_uint(sAhi[idx]); al[c] = __float_as_uint(sAlo[idx]);
        }
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const float b = tile[(8 * j + g) * MMA_LD + 8 * kh + t + 4 * c];
                bh[j][c] = __float_as_uint(b) & 0xffffe000u;
                bl[j][c] = __float_as_uint(b - __uint_as_float(bh[j][c]));
            }
        // the three products of one accumulator are issued four MMAs apart (the four site tiles are independent)
#pragma unroll
        for (int j = 0; j < 4; j++) mma_tf32(acc[j], al, bh[j][0], bh[j][1]);
#pragma unroll
        for (int j = 0; j < 4; j++) mma_tf32(acc[j], ah, bl[j][0], bl[j][1]);
#pragma unroll
        for (int j = 0; j < 4; j++) mma_tf32(acc[j], ah, bh[j][0], bh[j][1]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        tile[(8 * j + 2 * t) * MMA_LD + g] = acc[j][0];
        tile[(8 * j + 2 * t + 1) * MMA_LD + g] = acc[j][1];
        tile[(8 * j + 2 * t) * MMA_LD + g + 8] = acc[j][2];
        tile[(8 * j + 2 * t + 1) * MMA_LD + g + 8] = acc[j][3];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 v = *reinterpret_cast<const float4*>(tile + lane * MMA_LD + 4 * q);
        h[4 * q] = v.x; h[4 * q + 1] = v.y; h[4 * q + 2] = v.z; h[4 * q + 3] = v.w;
    }
    __syncwarp();
}

template <class R, int GEN, bool ONSITE, bool MMA = false>
__global__ void __launch_bounds__(SCMC_SWEEP_THREADS, sizeof(R) == 8 ? SCMC_SWEEP_PSI_BLOCKS_F64 : (MMA ? SCMC_SWEEP_BLOCKS_MMA : SCMC_SWEEP_BLOCKS)) k_sweep_colour_psi(const __grid_constant__ DevModel<R, 1> M, const SweepArgs A) {
    constexpr int D = GenOps<GEN>::D, NPV = D / 2, NP = GenOps<GEN>::NDENS;
    static_assert(!MMA || (sizeof(R) == 4 && !ONSITE && NP == NG), "MMA mat-vec: FP32, 16 density entries, no on-site term");
    const int64_t r = A.r0 + blockIdx.y;
    __shared__ __align__(16) R sJt[MMA ? 1 : NG * NG];
    __shared__ __align__(16) float sM[MMA ? 2 * NG * MMA_LD + (SCMC_SWEEP_THREADS / 32) * 32 * MMA_LD : 1];
    const bool fused = !ONSITE && NP == NG && A.Jqt != nullptr;      // density sums -> energy coefficients in one 16x16 mat-vec
    if constexpr (MMA) {
        for (int t = threadIdx.x; t < NG * NG; t += blockDim.x) {      // Jq[m][k] = Jqt[k][m], split into TF32 head and tail
            const float v = ((const float*)A.Jqt)[t];
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const int k = t / NG, m = t % NG;
            sM[m * MMA_LD + k] = hi; sM[NG * MMA_LD + m * MMA_LD + k] = v - hi;
        }
    } else {
        for (int t = threadIdx.x; t < NG * NG; t += blockDim.x) sJt[t] = ((const R*)(fused ? A.Jqt : A.Jt))[t];
    }
    const bool alive = (A.active == nullptr || A.active[r]);
    const R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
    const R nbeta = neg_beta_scaled(beta);
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    unsigned long long Pq[NPV];               // opaque bases of the replica's psi planes (one mad.wide per gather load)
#pragma unroll
    for (int v = 0; v < NPV; v++) {
        Pq[v] = (unsigned long long)(M.psi + rbase + v * ps);
        asm volatile("" : "+l"(Pq[v]));
    }
    const unsigned stride = (unsigned)M.N;
    __syncthreads();
    unsigned nacc = 0;
    for (int colour = A.colour; colour <= A.colour_last; colour++) {
        const int nc = M.col_off[colour + 1] - M.col_off[colour];
        // MMA: the loop condition is warp-uniform (mma.sync needs all 32 lanes); lanes past the end of the class redo its last site and do not store
        for (int s0 = blockIdx.x * blockDim.x + threadIdx.x; (MMA ? s0 - (int)(threadIdx.x & 31) : s0) < nc && alive; s0 += gridDim.x * blockDim.x) {
            const bool valid = !MMA || s0 < nc;
            const int s = valid ? s0 : nc - 1;
            const int cp = M.col_off[colour] + s;          // class position: p, site and neighbour slots are independent loads
            const int p = M.col_list[cp];
            const int site = M.cls_site[cp];
            R re[D], im[D], T[NG], Tsq[NG], h[NG];
            load_psi<R, D>(M.psi + rbase + p, ps, re, im);
            {
                R P[NP];
#pragma unroll
                for (int t = 0; t < NP; t++) P[t] = R(0);
                const int32_t* nb = M.cls_nbr + cp;
                // all six neighbours in flight before the first density product (same summation order => same bits).  Pays only where registers
                // allow it: the MMA variant (64 registers at 8 CTAs per SM, +4.7 %); the FFMA kernel spills at its 96-register cap (-1.4 %, SCMC_GATHER6=1)
                if ((MMA || SCMC_GATHER6) && M.z == 6) {
                    unsigned j[6];
#pragma unroll
                    for (int t = 0; t < 6; t++) j[t] = (unsigned)nb[(unsigned)t * stride];
                    vec4<R> v[6][NPV];
#pragma unroll
                    for (int t = 0; t < 6; t++)
#pragma unroll
                        for (int q = 0; q < NPV; q++) v[t][q] = *reinterpret_cast<const vec4<R>*>(Pq[q] + (unsigned long long)j[t] * sizeof(vec4<R>));
#pragma unroll
                    for (int t = 0; t < 6; t++) {
                        R nr[D], ni[D];
#pragma unroll
                        for (int q = 0; q < NPV; q++) { nr[2 * q] = v[t][q].x; ni[2 * q] = v[t][q].y; nr[2 * q + 1] = v[t][q].z; ni[2 * q + 1] = v[t][q].w; }
                        GenOps<GEN>::density_acc(nr, ni, P);
                    }
                } else
#pragma unroll 2
                for (int k = 0; k < M.z; k += 3) {
                    unsigned j[3];
#pragma unroll
                    for (int t = 0; t < 3; t++) j[t] = (unsigned)nb[(unsigned)(k + t) * stride];
                    vec4<R> v[3][NPV];
#pragma unroll
                    for (int t = 0; t < 3; t++)
#pragma unroll
                        for (int q = 0; q < NPV; q++) v[t][q] = *reinterpret_cast<const vec4<R>*>(Pq[q] + (unsigned long long)j[t] * sizeof(vec4<R>));
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        R nr[D], ni[D];
#pragma unroll
                        for (int q = 0; q < NPV; q++) { nr[2 * q] = v[t][q].x; ni[2 * q] = v[t][q].y; nr[2 * q + 1] = v[t][q].z; ni[2 * q + 1] = v[t][q].w; }
                        GenOps<GEN>::density_acc(nr, ni, P);
                    }
                }
                if constexpr (MMA) {
                    matvec_mma(sM, sM + NG * MMA_LD, sM + 2 * NG * MMA_LD + (threadIdx.x >> 5) * 32 * MMA_LD, P, h);
                } else if (!fused) {
                    R S[1][NG];
                    GenOps<GEN>::expect_from_density(P, S[0]);
                    field_from_sums_smem<1>(sJt, S, h);
                } else if constexpr (NP == NG) {
                    field_from_sums_smem<1>(sJt, (const R(*)[NG])P, h);       // h now holds Hq = Jq . Psum
                }
            }
            unsigned mine = 0;
            if (ONSITE) {
                GenOps<GEN>::expect(re, im, T);
                GenOps<GEN>::expect_sq(re, im, Tsq);
                R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
                for (int hit = 0; hit < A.hits; hit++) {
                    R xr[D], xi[D], u, dE;
                    stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
                    mine += metropolis_hit<R, GEN, ONSITE>(re, im, T, Tsq, h, M.cp.K, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
                }
            } else {
                R Hq[GenOps<GEN>::NDENS];
                if (fused || MMA) {
#pragma unroll
                    for (int t = 0; t < (NP == NG ? NG : 0); t++) Hq[t] = h[t];
                } else {
                    GenOps<GEN>::hq_from_field(h, Hq);
                }
                R eloc = GenOps<GEN>::energy_from_state(re, im, Hq);
                int hit = 0;
#if SCMC_HIT_UNROLL2
                // two hits per iteration: both Philox / Box-Muller streams are independent of psi, so they interleave (ILP)
                for (; hit + 1 < A.hits; hit += 2) {
                    R xr0[D], xi0[D], u0, xr1[D], xi1[D], u1, dE;
                    stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr0, xi0, u0);
                    stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit + 1, xr1, xi1, u1);
                    mine += metropolis_hit_psi<R, GEN>(re, im, Hq, xr0, xi0, u0, nbeta, sigma, eloc, dE) ? 1u : 0u;
                    mine += metropolis_hit_psi<R, GEN>(re, im, Hq, xr1, xi1, u1, nbeta, sigma, eloc, dE) ? 1u : 0u;
                }
#endif
                for (; hit < A.hits; hit++) {
                    R xr[D], xi[D], u, dE;
                    stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
                    mine += metropolis_hit_psi<R, GEN>(re, im, Hq, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
                }
            }
            if (!valid) mine = 0;
            if (mine) store_psi<R, D>(M.psi + rbase + p, ps, re, im);
            nacc += mine;
        }
        if (colour < A.colour_last) __syncthreads();
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(A.acc + r, (unsigned long long)tot);
}

// psi-gather colour pass for SEVERAL coupling labels (d = 4, 16 density entries, no on-site term, label-major neighbour slots: TG/h-BN at n = 1, 3).
// Neighbours enter through their states (32 B instead of a 64-B T vector), one label at a time: the label's density sums in 16 registers, then that
// label's share Jq_l . P_l of the energy coefficients (field_accumulate_smem); the hits are those of the one-label psi family (metropolis_hit_psi:
// e(psi) = sum_n Hq[n] P_n(psi)), nothing but psi is written and the T planes become the lazily refreshed cache (T_stale / ensure_T).
// Reference work unit: localUpdate! with getEnergyDifference!'s per-bond mat-vecs (src/MonteCarlo.jl:263-286, src/Configuration.jl:278-298).
template <class R, int GEN, int NL>
__global__ void __launch_bounds__(SCMC_SWEEP_THREADS, sizeof(R) == 8 ? SCMC_SWEEP_PSI_BLOCKS_F64 : SCMC_SWEEP_BLOCKS) k_sweep_colour_psi_nl(const __grid_constant__ DevModel<R, NL> M, const SweepArgs A) {
    constexpr int D = GenOps<GEN>::D, NPV = D / 2, NP = GenOps<GEN>::NDENS;
    static_assert(NP == NG && D == 4, "16 density entries, d = 4");
    constexpr bool F32 = sizeof(R) == 4;
    const int64_t r = A.r0 + blockIdx.y;
    __shared__ __align__(16) R sJq[NL * NG * NG];
    for (int t = threadIdx.x; t < NL * NG * NG; t += blockDim.x) sJq[t] = ((const R*)A.Jqt)[t];
    const bool alive = (A.active == nullptr || A.active[r]);
    const R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
    const R nbeta = neg_beta_scaled(beta);
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    unsigned long long Pq[NPV];               // opaque bases of the replica's psi planes (one mad.wide per gather load)
#pragma unroll
    for (int v = 0; v < NPV; v++) {
        Pq[v] = (unsigned long long)(M.psi + rbase + v * ps);
        asm volatile("" : "+l"(Pq[v]));
    }
    const unsigned stride = (unsigned)M.N;
    __syncthreads();
    unsigned nacc = 0;
    for (int colour = A.colour; colour <= A.colour_last; colour++) {
        const int nc = M.col_off[colour + 1] - M.col_off[colour];
        for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nc && alive; s += gridDim.x * blockDim.x) {
            const int cp = M.col_off[colour] + s;
            const int p = M.col_list[cp];
            const int site = M.cls_site[cp];
            R re[D], im[D], Hq[NG];
            load_psi<R, D>(M.psi + rbase + p, ps, re, im);
            const int32_t* nb = M.cls_nbr + cp;
            unsigned long long hp[NG / 2];
            if constexpr (F32) {
#pragma unroll
                for (int i = 0; i < NG / 2; i++) hp[i] = 0ull;
            } else {
#pragma unroll
                for (int a = 0; a < NG; a++) Hq[a] = R(0);
            }
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const int k1 = A.lab_off[l + 1];
                if (A.lab_off[l] == k1) continue;              // label without bonds (fewer labels than the build's maximum)
                R P[NP];
#pragma unroll
                for (int t = 0; t < NP; t++) P[t] = R(0);
                for (int k = A.lab_off[l]; k < k1; k += 3) {
                    unsigned j[3];
#pragma unroll
                    for (int t = 0; t < 3; t++) j[t] = (k + t < k1) ? (unsigned)nb[(unsigned)(k + t) * stride] : stride;      // past the label's range: zero site
                    vec4<R> v[3][NPV];
#pragma unroll
                    for (int t = 0; t < 3; t++)
#pragma unroll
                        for (int q = 0; q < NPV; q++) v[t][q] = *reinterpret_cast<const vec4<R>*>(Pq[q] + (unsigned long long)j[t] * sizeof(vec4<R>));
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        R nr[D], ni[D];
#pragma unroll
                        for (int q = 0; q < NPV; q++) { nr[2 * q] = v[t][q].x; ni[2 * q] = v[t][q].y; nr[2 * q + 1] = v[t][q].z; ni[2 * q + 1] = v[t][q].w; }
                        GenOps<GEN>::density_acc(nr, ni, P);
                    }
                }
                if constexpr (F32) field_accumulate_smem((const float*)sJq + l * NG * NG, (const float*)P, hp);
                else field_accumulate_smem((const double*)sJq + l * NG * NG, (const double*)P, (double*)Hq);
            }
            if constexpr (F32) {
#pragma unroll
                for (int i = 0; i < NG / 2; i++) asm("mov.b64 {%0, %1}, %2;" : "=f"(Hq[2 * i]), "=f"(Hq[2 * i + 1]) : "l"(hp[i]));
            }
            R eloc = GenOps<GEN>::energy_from_state(re, im, Hq);
            unsigned mine = 0;
            for (int hit = 0; hit < A.hits; hit++) {
                R xr[D], xi[D], u, dE;
                stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
                mine += metropolis_hit_psi<R, GEN>(re, im, Hq, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
            }
            if (mine) store_psi<R, D>(M.psi + rbase + p, ps, re, im);
            nacc += mine;
        }
        if (colour < A.colour_last) __syncthreads();
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(A.acc + r, (unsigned long long)tot);
}

// cp.async-pipelined form of k_sweep_colour_psi (FP32): the next visit's neighbour states and own state are staged in shared
// memory (14 x 16 B per thread for d = 4, z = 6 -> 28 KB per CTA, 5 CTAs/SM), indices two visits ahead in registers.
template <int GEN, bool ONSITE>
__global__ void __launch_bounds__(SCMC_SWEEP_THREADS, SCMC_SWEEP_BLOCKS) k_sweep_colour_psi_async(const __grid_constant__ DevModel<float, 1> M, const SweepArgs A) {
    using R = float;
    constexpr int D = GenOps<GEN>::D, NPV = D / 2, NP = GenOps<GEN>::NDENS;
    extern __shared__ __align__(16) float4 stage[];          // [(z + 1) * NPV][blockDim.x]
    __shared__ __align__(16) R sJt[NG * NG];
    const int64_t r = A.r0 + blockIdx.y;
    const int bd = blockDim.x, tid = threadIdx.x;
    const bool fused = !ONSITE && NP == NG && A.Jqt != nullptr;
    for (int t = tid; t < NG * NG; t += bd) sJt[t] = ((const R*)(fused ? A.Jqt : A.Jt))[t];
    const bool alive = (A.active == nullptr || A.active[r]);
    const R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
    const R nbeta = neg_beta_scaled(beta);
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    const float4* Pr = M.psi + rbase;
    const unsigned stride = (unsigned)M.N;
    const int z = M.z;
    __syncthreads();
    unsigned nacc = 0;
    for (int colour = A.colour; colour <= A.colour_last; colour++) {
        const int nc = M.col_off[colour + 1] - M.col_off[colour];
        const int32_t* clist = M.col_list + M.col_off[colour];
        const int step = gridDim.x * bd;
        int s = blockIdx.x * bd + tid;
        int p = 0, site = 0, pn = 0, site_n = 0;
        unsigned jn[ASYNC_Z];
#pragma unroll
        for (int k = 0; k < ASYNC_Z; k++) jn[k] = 0u;
        if (s < nc && alive) {
            p = clist[s];
            site = M.orig[p];
#pragma unroll
            for (int k = 0; k < ASYNC_Z; k++)
                if (k < z) {
                    const unsigned j = (unsigned)M.nbr[(unsigned)k * stride + p];
#pragma unroll
                    for (int q = 0; q < NPV; q++) cp_async16(&stage[(k * NPV + q) * bd + tid], Pr + j + q * ps);
                }
#pragma unroll
            for (int v = 0; v < NPV; v++) cp_async16(&stage[(z * NPV + v) * bd + tid], Pr + p + v * ps);
            if (s + step < nc) {
                pn = clist[s + step];
#pragma unroll
                for (int k = 0; k < ASYNC_Z; k++) jn[k] = k < z ? (unsigned)M.nbr[(unsigned)k * stride + pn] : 0u;
                site_n = M.orig[pn];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (; s < nc && alive; s += step) {
            const bool has_next = s + step < nc;
            const bool has_next2 = s + 2 * step < nc;
            int pnn = 0, site_nn = 0;
            unsigned jnn[ASYNC_Z];
#pragma unroll
            for (int k = 0; k < ASYNC_Z; k++) jnn[k] = 0u;
            if (has_next2) {
                pnn = clist[s + 2 * step];
#pragma unroll
                for (int k = 0; k < ASYNC_Z; k++) jnn[k] = k < z ? (unsigned)M.nbr[(unsigned)k * stride + pnn] : 0u;
                site_nn = M.orig[pnn];
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            R re[D], im[D], T[NG], Tsq[NG], h[NG];
#pragma unroll
            for (int v = 0; v < NPV; v++) {
                const float4 q4 = stage[(z * NPV + v) * bd + tid];
                re[2 * v] = q4.x; im[2 * v] = q4.y; re[2 * v + 1] = q4.z; im[2 * v + 1] = q4.w;
            }
            {
                R P[NP];
#pragma unroll
                for (int t = 0; t < NP; t++) P[t] = R(0);
#pragma unroll
                for (int k = 0; k < ASYNC_Z; k++)
                    if (k < z) {
                        R nr[D], ni[D];
#pragma unroll
                        for (int q = 0; q < NPV; q++) {
                            const float4 w4 = stage[(k * NPV + q) * bd + tid];
                            nr[2 * q] = w4.x; ni[2 * q] = w4.y; nr[2 * q + 1] = w4.z; ni[2 * q + 1] = w4.w;
                        }
                        GenOps<GEN>::density_acc(nr, ni, P);
                    }
                if (has_next) {
#pragma unroll
                    for (int k = 0; k < ASYNC_Z; k++)
                        if (k < z) {
#pragma unroll
                            for (int q = 0; q < NPV; q++) cp_async16(&stage[(k * NPV + q) * bd + tid], Pr + jn[k] + q * ps);
                        }
#pragma unroll
                    for (int v = 0; v < NPV; v++) cp_async16(&stage[(z * NPV + v) * bd + tid], Pr + pn + v * ps);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (!fused) {
                    R S[1][NG];
                    GenOps<GEN>::expect_from_density(P, S[0]);
                    field_from_sums_smem<1>(sJt, S, h);
                } else if constexpr (NP == NG) {
                    field_from_sums_smem<1>(sJt, (const R(*)[NG])P, h);       // h now holds Hq = Jq . Psum
                }
            }
            unsigned mine = 0;
            if (ONSITE) {
                GenOps<GEN>::expect(re, im, T);
                GenOps<GEN>::expect_sq(re, im, Tsq);
                R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
                for (int hit = 0; hit < A.hits; hit++) {
                    R xr[D], xi[D], u, dE;
                    stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
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
                    stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
                    mine += metropolis_hit_psi<R, GEN>(re, im, Hq, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
                }
            }
            if (mine) store_psi<R, D>(M.psi + rbase + p, ps, re, im);
            nacc += mine;
            p = pn; site = site_n;
            pn = pnn; site_n = site_nn;
#pragma unroll
            for (int k = 0; k < ASYNC_Z; k++) jn[k] = jnn[k];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (colour < A.colour_last) __syncthreads();
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
    if ((tid & 31) == 0 && tot) atomicAdd(A.acc + r, (unsigned long long)tot);
}

// T (and Tsq) planes from psi for every site of every replica: refreshes the cache after psi-gather sweeps
template <class R, int GEN, bool ONSITE, int NL>
__global__ void __launch_bounds__(256) k_refresh_T(const __grid_constant__ DevModel<R, NL> M) {
    constexpr int D = GenOps<GEN>::D;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M.nrep * M.N) return;
    const int64_t r = idx / M.N;
    const int p = (int)(idx - r * M.N);
    const size_t ps = (size_t)M.nrep * M.Ns;
    R re[D], im[D], T[NG];
    load_psi<R, D>(M.psi + r * M.Ns + p, ps, re, im);
    GenOps<GEN>::expect(re, im, T);
    store_T<R>(M.T + r * M.Ns + p, ps, T);
    if (ONSITE) {
        GenOps<GEN>::expect_sq(re, im, T);
        store_T<R>(M.Tsq + r * M.Ns + p, ps, T);
    }
}

// Software-pipelined variant of k_sweep_colour (FP32, one label, <= 6 neighbour slots): every thread walks over several site
// visits of its replica; while the hits of visit t run, the neighbour T vectors and psi of visit t+1 are already in flight as
// cp.async (LDGSTS) copies into a per-thread shared-memory slot -- no registers are held and the DRAM round trip disappears
// behind ~1000 instructions of arithmetic.  Same arithmetic and order as k_sweep_colour => bit-identical chains.
template <int GEN, bool ONSITE>
__global__ void __launch_bounds__(SCMC_ASYNC_THREADS, SCMC_ASYNC_BLOCKS) k_sweep_colour_async(const __grid_constant__ DevModel<float, 1> M, const SweepArgs A) {
    using R = float;
    constexpr int D = GenOps<GEN>::D, NPV = D / 2;
    extern __shared__ __align__(16) float4 stage[];          // [z * 4 + NPV][blockDim.x]
    __shared__ __align__(16) R sJt[NG * NG];
    const int64_t r = A.r0 + blockIdx.y;
    const int bd = blockDim.x, tid = threadIdx.x;
    for (int t = tid; t < NG * NG; t += bd) sJt[t] = ((const R*)A.Jt)[t];
    const bool alive = (A.active == nullptr || A.active[r]);
    const R beta = ((const R*)A.beta)[r], sigma = ((const R*)A.sigma)[r];
    const R nbeta = neg_beta_scaled(beta);
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    const float4* Tr = M.T + rbase;
    const float4* Pr = M.psi + rbase;
    const unsigned stride = (unsigned)M.N;
    const int z = M.z;
    __syncthreads();
    unsigned nacc = 0;
    for (int colour = A.colour; colour <= A.colour_last; colour++) {
        const int nc = M.col_off[colour + 1] - M.col_off[colour];
        const int32_t* clist = M.col_list + M.col_off[colour];
        const int step = gridDim.x * bd;
        int s = blockIdx.x * bd + tid;
        // Pipeline: indices (colour list -> site -> neighbour slots) are fetched TWO visits ahead into registers, the T / psi
        // data ONE visit ahead into shared memory, so the steady state never waits on a dependent round trip.
        int p = 0, site = 0, pn = 0, site_n = 0;
        unsigned jn[ASYNC_Z];
#pragma unroll
        for (int k = 0; k < ASYNC_Z; k++) jn[k] = 0u;
        if (s < nc && alive) {       // prologue: stage visit 0, fetch the indices of visit 1
            p = clist[s];
            site = M.orig[p];
#pragma unroll
            for (int k = 0; k < ASYNC_Z; k++)
                if (k < z) {
                    const unsigned j = (unsigned)M.nbr[(unsigned)k * stride + p];
#pragma unroll
                    for (int q = 0; q < 4; q++) cp_async16(&stage[(k * 4 + q) * bd + tid], Tr + j + q * ps);
                }
#pragma unroll
            for (int v = 0; v < NPV; v++) cp_async16(&stage[(z * 4 + v) * bd + tid], Pr + p + v * ps);
            if (s + step < nc) {
                pn = clist[s + step];
#pragma unroll
                for (int k = 0; k < ASYNC_Z; k++) jn[k] = k < z ? (unsigned)M.nbr[(unsigned)k * stride + pn] : 0u;
                site_n = M.orig[pn];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (; s < nc && alive; s += step) {
            const bool has_next = s + step < nc;
            const bool has_next2 = s + 2 * step < nc;
            // indices of visit t+2: issued now, consumed at the end of this visit
            int pnn = 0, site_nn = 0;
            unsigned jnn[ASYNC_Z];
            if (has_next2) {
                pnn = clist[s + 2 * step];
#pragma unroll
                for (int k = 0; k < ASYNC_Z; k++) jnn[k] = k < z ? (unsigned)M.nbr[(unsigned)k * stride + pnn] : 0u;
                site_nn = M.orig[pnn];
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            R re[D], im[D], T[NG], Tsq[NG], h[NG];
#pragma unroll
            for (int v = 0; v < NPV; v++) {
                const float4 q4 = stage[(z * 4 + v) * bd + tid];
                re[2 * v] = q4.x; im[2 * v] = q4.y; re[2 * v + 1] = q4.z; im[2 * v + 1] = q4.w;
            }
            {
                R S[1][NG];
#pragma unroll
                for (int a = 0; a < NG; a++) S[0][a] = R(0);
#pragma unroll
                for (int k = 0; k < ASYNC_Z; k++)
                    if (k < z) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float4 v4 = stage[(k * 4 + q) * bd + tid];
                            S[0][4 * q] += v4.x; S[0][4 * q + 1] += v4.y; S[0][4 * q + 2] += v4.z; S[0][4 * q + 3] += v4.w;
                        }
                    }
                // the slot is free again: launch the copies of the next visit (its indices arrived a visit ago)
                if (has_next) {
#pragma unroll
                    for (int k = 0; k < ASYNC_Z; k++)
                        if (k < z) {
#pragma unroll
                            for (int q = 0; q < 4; q++) cp_async16(&stage[(k * 4 + q) * bd + tid], Tr + jn[k] + q * ps);
                        }
#pragma unroll
                    for (int v = 0; v < NPV; v++) cp_async16(&stage[(z * 4 + v) * bd + tid], Pr + pn + v * ps);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                field_from_sums_smem<1>(sJt, S, h);
            }
            GenOps<GEN>::expect(re, im, T);
            if (ONSITE) GenOps<GEN>::expect_sq(re, im, Tsq);
            R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
            unsigned mine = 0;
            for (int hit = 0; hit < A.hits; hit++) {
                R xr[D], xi[D], u, dE;
                stream_any<R, D>(A.stream, A.seed, (uint32_t)site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
                mine += metropolis_hit<R, GEN, ONSITE>(re, im, T, Tsq, h, M.cp.K, xr, xi, u, nbeta, sigma, eloc, dE) ? 1u : 0u;
            }
            if (mine) {
                store_psi<R, D>(M.psi + rbase + p, ps, re, im);
                store_T<R>(M.T + rbase + p, ps, T);
                if (ONSITE) store_T<R>(M.Tsq + rbase + p, ps, Tsq);
            }
            nacc += mine;
            p = pn; site = site_n;
            pn = pnn; site_n = site_nn;
#pragma unroll
            for (int k = 0; k < ASYNC_Z; k++) jn[k] = jnn[k];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (colour < A.colour_last) __syncthreads();
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, nacc);
    if ((tid & 31) == 0 && tot) atomicAdd(A.acc + r, (unsigned long long)tot);
}

// Energy + magnetisation of every replica from the cached T (Tsq) planes (T-gather family).  grid = (CTAs per replica, replicas);
// every CTA leaves NG + 1 FP64 partial sums, k_measure_finish adds them in a fixed order (deterministic) and writes
// row = [E/N, (E/N)^2, |m_sigma|, |m_tau|, |m_sigmatau|, m[16]]  (measure!, ObservablesGeneric.jl:32-46) and Etot[r] = E.
template <class R, int GEN, bool ONSITE, int NL>
__global__ void __launch_bounds__(SCMC_MEASURE_THREADS, sizeof(R) == 8 ? 2 : 4) k_measure(const __grid_constant__ DevModel<R, NL> M, double* partial, const SweepArgs A) {
    const int64_t r = blockIdx.y;
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    // several labels with site-independent slot ranges: the label-major gather of the sweep kernel (couplings staged in shared memory)
    __shared__ __align__(16) R sJt[NL > 1 ? NL * NG * NG : 4];
    __shared__ double red[SCMC_MEASURE_THREADS / 32][NG + 1];
    const bool lab_major = NL > 1 && A.lab_major;
    if (lab_major) {
        for (int t = threadIdx.x; t < NL * NG * NG; t += blockDim.x) sJt[t] = ((const R*)A.Jt)[t];
        __syncthreads();
    }
    // per-thread partial sums: the energy in FP64, the 16 magnetisation sums in the handle's precision (a thread adds at most a few
    // dozen values of magnitude <= d; everything after that is FP64)
    double eacc = 0.0;
    R macc[NG];
#pragma unroll
    for (int a = 0; a < NG; a++) macc[a] = R(0);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < M.N; p += gridDim.x * blockDim.x) {
        asm volatile("" ::: "memory");      // keeps the coupling-table loads inside the loop (hoisted, they spill: see k_measure_psi)
        R T[NG], h[NG];
        load_T<R>(M.T + rbase + p, ps, T);
        if (lab_major) {
            field_label_major<R, NL>(M, A, sJt, rbase, ps, p, h);
        } else {
            R S[NL][NG];
            neighbour_sums<R, NL>(M, rbase, ps, p, S);
            field_from_sums<R, NL>(M.cp, S, h);
        }
        R e = R(0);
#pragma unroll
        for (int a = 0; a < NG; a++) e = fma(T[a], h[a], e);
        e *= R(0.5);
        if (ONSITE) {
            R Tsq[NG];
            load_T<R>(M.Tsq + rbase + p, ps, Tsq);
#pragma unroll
            for (int a = 0; a < NG; a++) e = fma(Tsq[a], M.cp.K[a], e);
        }
        eacc += (double)e;
#pragma unroll
        for (int a = 0; a < NG; a++) macc[a] += T[a];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t <= NG; t++) {
        double v = t < NG ? (double)macc[t < NG ? t : 0] : eacc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x <= NG) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w][threadIdx.x];
        partial[((size_t)r * gridDim.x + blockIdx.x) * (NG + 1) + threadIdx.x] = v;
    }
}

// psi-family measurement: E and sum_i T_i straight from the states, with the gather / density-sum / mat-vec arithmetic of
// k_sweep_colour_psi (no T planes are read, so no refresh after psi-gather sweeps).  grid = (CTAs per replica, replicas); every CTA
// leaves NG + 1 FP64 partial sums, k_measure_finish adds them in a fixed order (deterministic) and writes the measure! row.
// On the fused path (16 density entries, no on-site term) the partial sums are density sums; T = C . P is applied once at the end.
template <class R, int GEN, bool ONSITE>
__global__ void __launch_bounds__(SCMC_MEASURE_THREADS, sizeof(R) == 8 ? 3 : 5) k_measure_psi(const __grid_constant__ DevModel<R, 1> M, const void* Jt_, const void* Jqt_, double* partial, const int n_fwd, const uint64_t fwd_slots) {
    constexpr int D = GenOps<GEN>::D, NPV = D / 2, NP = GenOps<GEN>::NDENS;
    const int64_t r = blockIdx.y;
    __shared__ __align__(16) R sJt[NG * NG];
    __shared__ double red[SCMC_MEASURE_THREADS / 32][NG + 1];
    const bool fused = !ONSITE && NP == NG && Jqt_ != nullptr;
    for (int t = threadIdx.x; t < NG * NG; t += blockDim.x) sJt[t] = ((const R*)(fused ? Jqt_ : Jt_))[t];
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    unsigned long long Pq[NPV];
#pragma unroll
    for (int v = 0; v < NPV; v++) {
        Pq[v] = (unsigned long long)(M.psi + rbase + v * ps);
        asm volatile("" : "+l"(Pq[v]));
    }
    const unsigned stride = (unsigned)M.N;
    __syncthreads();
    // per-thread partial sums: the energy in FP64, the 16 magnetisation sums in the handle's precision (a thread adds at most a few
    // dozen values of magnitude <= 1; everything after that is FP64)
    double eacc = 0.0;
    R macc[NG];
#pragma unroll
    for (int t = 0; t < NG; t++) macc[t] = R(0);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < M.N; p += gridDim.x * blockDim.x) {
        asm volatile("" ::: "memory");      // keeps the 64 coupling-table LDS.128 inside the loop (hoisted, they spill to local memory)
        R re[D], im[D], h[NG], P[NP];
        load_psi<R, D>(M.psi + rbase + p, ps, re, im);
#pragma unroll
        for (int t = 0; t < NP; t++) P[t] = R(0);
        const int32_t* nb = M.nbr + p;
        // forward slots only (n_fwd > 0, fused path): every bond is summed once, from the end that holds it in a forward slot, instead of twice with a factor 1/2
        const bool fwd = fused && n_fwd > 0;
        const int kz = fwd ? n_fwd : M.z;
#pragma unroll 2
        for (int k = 0; k < kz; k += 3) {
            unsigned j[3];
#pragma unroll
            for (int t = 0; t < 3; t++) {
                const unsigned slot = fwd ? (unsigned)(fwd_slots >> (8 * (k + t))) & 0xFFu : (unsigned)(k + t);
                j[t] = (fwd && k + t >= n_fwd) ? (unsigned)M.N : (unsigned)nb[slot * stride];      // beyond the set: the all-zero padding site
            }
            vec4<R> v[3][NPV];
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int q = 0; q < NPV; q++) v[t][q] = *reinterpret_cast<const vec4<R>*>(Pq[q] + (unsigned long long)j[t] * sizeof(vec4<R>));
#pragma unroll
            for (int t = 0; t < 3; t++) {
                R nr[D], ni[D];
#pragma unroll
                for (int q = 0; q < NPV; q++) { nr[2 * q] = v[t][q].x; ni[2 * q] = v[t][q].y; nr[2 * q + 1] = v[t][q].z; ni[2 * q + 1] = v[t][q].w; }
                GenOps<GEN>::density_acc(nr, ni, P);
            }
        }
        if (!fused) {
            R S[1][NG], T[NG];
            GenOps<GEN>::expect_from_density(P, S[0]);
            field_from_sums_smem<1>(sJt, S, h);
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
            for (int t = 0; t < NG; t++) macc[t] += T[t];
        } else if constexpr (NP == NG) {
            field_from_sums_smem<1>(sJt, (const R(*)[NG])P, h);           // h = Hq
            eacc += (fwd ? 1.0 : 0.5) * (double)GenOps<GEN>::energy_from_state(re, im, h);
            R Pown[NP];
#pragma unroll
            for (int t = 0; t < NP; t++) Pown[t] = R(0);
            GenOps<GEN>::density_acc(re, im, Pown);
#pragma unroll
            for (int t = 0; t < NG; t++) macc[t] += Pown[t];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t <= NG; t++) {
        double v = t < NG ? (double)macc[t < NG ? t : 0] : eacc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x <= NG) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w][threadIdx.x];
        partial[((size_t)r * gridDim.x + blockIdx.x) * (NG + 1) + threadIdx.x] = v;
    }
}

template <int GEN>
__global__ void __launch_bounds__(128) k_measure_finish(int64_t nrep, int nparts, int fused, int64_t N, double e_const, const double* partial, double* rows, double* Etot) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrep) return;
    double tot[NG + 1], Tsum[NG];
    for (int t = 0; t <= NG; t++) tot[t] = 0.0;
    for (int b = 0; b < nparts; b++)
        for (int t = 0; t <= NG; t++) tot[t] += partial[((size_t)r * nparts + b) * (NG + 1) + t];
    if (fused) {
        if constexpr (GenOps<GEN>::NDENS == NG) GenOps<GEN>::expect_from_density(tot, Tsum);
    } else {
        for (int t = 0; t < NG; t++) Tsum[t] = tot[t];
    }
    const double E = tot[NG] + e_const, invN = 1.0 / (double)N;
    if (Etot) Etot[r] = E;
    if (rows) {
        double* row = rows + r * SCMC_NOBS;
        double m[NG];
        for (int a = 0; a < NG; a++) m[a] = Tsum[a] * invN;
        row[0] = E * invN;
        row[1] = (E * invN) * (E * invN);
        row[2] = sqrt(m[4] * m[4] + m[8] * m[8] + m[12] * m[12]);
        row[3] = sqrt(m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
        row[4] = sqrt(m[5] * m[5] + m[6] * m[6] + m[7] * m[7] + m[9] * m[9] + m[10] * m[10] + m[11] * m[11] + m[13] * m[13] +
                      m[14] * m[14] + m[15] * m[15]);
        for (int a = 0; a < NG; a++) row[5 + a] = m[a];
    }
}

template <class R, int NL>
__global__ void __launch_bounds__(256) k_local_field(const __grid_constant__ DevModel<R, NL> M, int64_t r, double* out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M.N) return;
    R h[NG], S[NL][NG];
    neighbour_sums<R, NL>(M, r * M.Ns, (size_t)M.nrep * M.Ns, p, S);
    field_from_sums<R, NL>(M.cp, S, h);
    const int site = M.orig[p];
    for (int a = 0; a < NG; a++) out[(size_t)site * NG + a] = (double)h[a];
}

// host layout [count][N][d] doubles (caller's site order)  ->  planes + T caches
template <class R, int GEN, bool ONSITE, int NL>
__global__ void __launch_bounds__(256) k_pack(const __grid_constant__ DevModel<R, NL> M, int64_t r0, int64_t count, const double* re_in,
                                              const double* im_in) {
    constexpr int D = GenOps<GEN>::D;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * M.N) return;
    const int64_t q = idx / M.N;
    const int p = (int)(idx - q * M.N);
    const int site = M.orig[p];
    R re[D], im[D], T[NG];
#pragma unroll
    for (int k = 0; k < D; k++) {
        re[k] = (R)re_in[((size_t)q * M.N + site) * D + k];
        im[k] = (R)im_in[((size_t)q * M.N + site) * D + k];
    }
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t at = (r0 + q) * M.Ns + p;
    store_psi<R, D>(M.psi + at, ps, re, im);
    GenOps<GEN>::expect(re, im, T);
    store_T<R>(M.T + at, ps, T);
    if (ONSITE) {
        GenOps<GEN>::expect_sq(re, im, T);
        store_T<R>(M.Tsq + at, ps, T);
    }
}
template <class R, int GEN, int NL>
__global__ void __launch_bounds__(256) k_unpack(const __grid_constant__ DevModel<R, NL> M, int64_t r0, int64_t count, double* re_out,
                                                double* im_out) {
    constexpr int D = GenOps<GEN>::D;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * M.N) return;
    const int64_t q = idx / M.N;
    const int p = (int)(idx - q * M.N);
    const int site = M.orig[p];
    R re[D], im[D];
    load_psi<R, D>(M.psi + (r0 + q) * M.Ns + p, (size_t)M.nrep * M.Ns, re, im);
#pragma unroll
    for (int k = 0; k < D; k++) {
        re_out[((size_t)q * M.N + site) * D + k] = (double)re[k];
        im_out[((size_t)q * M.N + site) * D + k] = (double)im[k];
    }
}
template <class R, int GEN, bool ONSITE, int NL>
__global__ void __launch_bounds__(256) k_get_T(const __grid_constant__ DevModel<R, NL> M, int64_t r, double* T_out, double* Tsq_out) {
    constexpr int D = GenOps<GEN>::D;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M.N) return;
    const int site = M.orig[p];
    const size_t ps = (size_t)M.nrep * M.Ns;
    R T[NG];
    load_T<R>(M.T + r * M.Ns + p, ps, T);
    for (int a = 0; a < NG; a++) T_out[(size_t)site * NG + a] = (double)T[a];
    if (Tsq_out) {
        if (ONSITE) {
            load_T<R>(M.Tsq + r * M.Ns + p, ps, T);
        } else {
            R re[D], im[D];
            load_psi<R, D>(M.psi + r * M.Ns + p, ps, re, im);
            GenOps<GEN>::expect_sq(re, im, T);
        }
        for (int a = 0; a < NG; a++) Tsq_out[(size_t)site * NG + a] = (double)T[a];
    }
}
// Haar-random product state from stream v1 at the reserved visit number (getRandomState, MonteCarlo.jl:244-246)
template <class R, int GEN, bool ONSITE, int NL>
__global__ void __launch_bounds__(256) k_randomize(const __grid_constant__ DevModel<R, NL> M, uint64_t seed, int64_t roff) {
    constexpr int D = GenOps<GEN>::D;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M.nrep * M.N) return;
    const int64_t r = idx / M.N;
    const int p = (int)(idx - r * M.N);
    R re[D], im[D], T[NG], u;
    stream_v1<R, D>(seed, (uint32_t)M.orig[p], (uint32_t)(roff + r), VISIT_INIT, re, im, u);
    R n2 = R(0);
#pragma unroll
    for (int k = 0; k < D; k++) n2 += re[k] * re[k] + im[k] * im[k];
    const R inv = R(1) / sqrt(n2);
#pragma unroll
    for (int k = 0; k < D; k++) { re[k] *= inv; im[k] *= inv; }
    const size_t ps = (size_t)M.nrep * M.Ns;
    store_psi<R, D>(M.psi + r * M.Ns + p, ps, re, im);
    GenOps<GEN>::expect(re, im, T);
    store_T<R>(M.T + r * M.Ns + p, ps, T);
    if (ONSITE) {
        GenOps<GEN>::expect_sq(re, im, T);
        store_T<R>(M.Tsq + r * M.Ns + p, ps, T);
    }
}

// One update of one site with injected randoms (parity probe for localUpdate!, MonteCarlo.jl:263-286)
template <class R, int GEN, bool ONSITE, int NL>
__global__ void k_single_update(const __grid_constant__ DevModel<R, NL> M, int64_t r, int p, const double* xi_in, double u, double beta,
                                double sigma, double* out /* [0]=dE, [1]=accepted */) {
    constexpr int D = GenOps<GEN>::D;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    R re[D], im[D], T[NG], Tsq[NG], h[NG], S[NL][NG], xr[D], xi[D], dE;
    load_psi<R, D>(M.psi + rbase + p, ps, re, im);
    load_T<R>(M.T + rbase + p, ps, T);
    if (ONSITE) load_T<R>(M.Tsq + rbase + p, ps, Tsq);
    neighbour_sums<R, NL>(M, rbase, ps, p, S);
    field_from_sums<R, NL>(M.cp, S, h);
#pragma unroll
    for (int k = 0; k < D; k++) { xr[k] = (R)xi_in[k]; xi[k] = (R)xi_in[D + k]; }
    R eloc = local_energy<R, ONSITE>(T, Tsq, h, M.cp.K);
    const bool ok = metropolis_hit<R, GEN, ONSITE>(re, im, T, Tsq, h, M.cp.K, xr, xi, (R)u, neg_beta_scaled((R)beta), (R)sigma, eloc, dE);
    if (ok) {
        store_psi<R, D>(M.psi + rbase + p, ps, re, im);
        store_T<R>(M.T + rbase + p, ps, T);
        if (ONSITE) store_T<R>(M.Tsq + rbase + p, ps, Tsq);
    }
    out[0] = (double)dE; out[1] = ok ? 1.0 : 0.0;
}

// ------------------------------------------------------------------------------------------------ dispatch
// ------------------------------------------------------------------------------------------------ self-check of the device tables
// compute-sanitizer / racecheck are closed on the GPU pool this was developed on, so the library carries its own check of what the sweep kernels
// rely on, run ON THE DEVICE COPIES the kernels read: (1) no bond joins two sites of one colour class (sites of a class are updated concurrently:
// this is the race-freedom condition of a colour pass) and every bond has its reverse; (2) the per-class-position tables repeat the neighbour
// table; (3) every thread slot of the bulk-copy tile plan (scmc_tma.cu) resolves, through the tile's runs, to the thread's own site and to exactly
// its neighbours.  scmc_verify_tables returns the number of violations (0 for every handle scmc_create accepts; the host validates the same
// properties before upload, this re-derives them from device memory).
__global__ void k_verify_class_of(int n_col, const int32_t* col_off9, const int32_t* col_list, int32_t* cls_of) {
    const int cp = blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= col_off9[n_col]) return;
    int c = 0;
    while (c + 1 < n_col && cp >= col_off9[c + 1]) c++;
    cls_of[col_list[cp]] = c;
}
__global__ void k_verify_tables(int N, int z, int n_col, const int32_t* col_off9, const int32_t* nbr, const int32_t* col_list, const int32_t* cls_nbr,
                                const int32_t* cls_site, const int32_t* orig, const int32_t* cls_of, unsigned long long* err) {
    const int cp = blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= N) return;
    unsigned bad = 0;
    const int p = col_list[cp];
    if (p < 0 || p >= N) { atomicAdd(err, 1ull); return; }
    if (cls_site[cp] != orig[p]) bad++;
    for (int k = 0; k < z; k++) {
        const int j = nbr[(size_t)k * N + p];
        if (cls_nbr[(size_t)k * N + cp] != j) bad++;
        if (j < 0 || j > N) { bad++; continue; }
        if (j == N) continue;                                   // no bond in this slot (zero site)
        if (cls_of[j] == cls_of[p]) bad++;                      // a colour pass would update both ends of a bond at once
        bool back = false;
        for (int k2 = 0; k2 < z; k2++) back = back || nbr[(size_t)k2 * N + j] == p;
        if (!back) bad++;
    }
    if (bad) atomicAdd(err, (unsigned long long)bad);
}
__global__ void k_verify_tma_plan(int N, int z, int cap, const uint4* meta, const uint2* runs, const int32_t* tile_info, const int32_t* nbr, const int32_t* orig,
                                  const int32_t* col_list, unsigned long long* err) {
    const int tile = blockIdx.x, tid = threadIdx.x;
    const uint4 m0 = meta[((size_t)tile * TMA_TILE + tid) * 2], m1 = meta[((size_t)tile * TMA_TILE + tid) * 2 + 1];
    const int cnt = tile_info[tile * 4 + 0], nruns = tile_info[tile * 4 + 1], slots = tile_info[tile * 4 + 2], cp0 = tile_info[tile * 4 + 3];
    const bool valid = (m0.w >> 16) != 0;
    unsigned bad = 0;
    if (valid != (tid < cnt)) bad++;
    if (nruns < 1 || nruns > TMA_MAX_RUNS || slots + 1 > cap) bad++;
    if (valid) {
        const uint32_t w[4] = {m0.x, m0.y, m0.z, m0.w};
        const int p = (int)m1.y;
        if (p != col_list[cp0 + tid] || (int)m1.x != orig[p]) bad++;
        for (int k = 0; k < 7; k++) {
            const uint32_t l = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
            const int want = k == 0 ? p : (k - 1 < z ? nbr[(size_t)(k - 1) * N + p] : N);
            if (want == N) { if (l != 0xFFFFu) bad++; continue; }       // no bond: the all-zero slot
            int src = -1;
            for (int q = 0; q < nruns; q++) {
                const uint2 rd = runs[(size_t)tile * TMA_MAX_RUNS + q];
                const uint32_t s0 = rd.y & 0xFFFFu, len = rd.y >> 16;
                if (l >= s0 && l < s0 + len) src = (int)rd.x + (int)(l - s0);
            }
            if (src != want) bad++;
        }
    }
    if (bad) atomicAdd(err, (unsigned long long)bad);
}

template <class R, int NL>
static DevModel<R, NL> make_model(const scmc_handle* h) {
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

template <int V> using ic = std::integral_constant<int, V>;
// f(R tag, GEN, ONSITE, NL) for the handle's configuration
template <class F>
static int32_t dispatch(const scmc_handle* h, F&& f) {
    auto with_nl = [&](auto rt, auto gen, auto ons) -> int32_t {
        if (h->n_labels == 1) return f(rt, gen, ons, ic<1>{});
        return f(rt, gen, ons, ic<MAXL>{});
    };
    auto with_gen = [&](auto rt) -> int32_t {
        switch (h->gen_id) {
            case 1: return with_nl(rt, ic<1>{}, ic<0>{});
            case 2: return h->onsite ? with_nl(rt, ic<2>{}, ic<1>{}) : with_nl(rt, ic<2>{}, ic<0>{});
            case 3: return with_nl(rt, ic<3>{}, ic<0>{});
            // caller-supplied tables: dense contraction, always MAXL labels and the on-site term (one instantiation per precision and d)
            case 4: if (set_custom_tables(*h->custom, h, h->device) != cudaSuccess) break; return f(rt, ic<4>{}, ic<1>{}, ic<MAXL>{});
            case 5: if (set_custom_tables(*h->custom, h, h->device) != cudaSuccess) break; return f(rt, ic<5>{}, ic<1>{}, ic<MAXL>{});
        }
        set_error("internal: bad generator id");
        return SCMC_ERR_ARG;
    };
    return h->prec == 64 ? with_gen(double{}) : with_gen(float{});
}

static inline unsigned grid_for(int64_t n, int block = 256) { return (unsigned)((n + block - 1) / block); }

// beta / sigma tables: converted on the host into the handle's pinned staging buffer, one async copy each, no sync
// (the staging buffer is reused only after the stream has consumed it: callers synchronise before returning to the user).
int32_t upload_scalars(scmc_handle* h, const double* beta, const double* sigma) {
    const size_t bytes = h->R * h->esz();
    for (int pass = 0; pass < 2; pass++) {
        const double* src = pass ? sigma : beta;
        if (!src) continue;
        char* stage = (char*)h->pinned + pass * bytes;
        for (int64_t r = 0; r < h->R; r++) {
            if (h->prec == 64) ((double*)stage)[r] = src[r];
            else ((float*)stage)[r] = (float)src[r];
        }
        SCMC_CK(cudaMemcpyAsync(pass ? h->sigma : h->beta, stage, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    return SCMC_OK;
}

// Which arithmetic family the production (stream-driven) sweeps of this handle use.  It depends on the model only -- never on
// the replica count or the number of sweeps per call -- so that a replica's chain is identical however replicas are sharded.
bool uses_psi_gather(const scmc_handle* h) {
    if (h->sweep_mode == 3 || h->sweep_mode == 4 || h->sweep_mode == 5 || h->sweep_mode == 6) return (h->n_labels == 1 && h->gen_id <= 3) || h->psi_labels;
    if (h->sweep_mode == 1 || h->sweep_mode == 2) return false;
    // several labels with d = 4 (TG/h-BN at n = 1, 3): the streaming psi-gather kernel moves half the bytes per neighbour and has the cheap hits (measured on
    // 48x48 x 1024: 2.24e10 updates/s against 1.78e10 for the streaming and 1.39e10 for the resident T-gather kernel); colour passes too small to fill the chip
    // (launch-bound) stay with the replica-resident T-gather kernel where that one is eligible
    if (h->psi_labels) return !(h->res.eligible && resident_preferred(h) && h->R * h->N / std::max(1, h->n_col) < 300000);
    return h->gen_id <= 3 && h->n_labels == 1 && (h->d == 4 || h->onsite) && !h->res_family_T;      // measured: d = 6 without on-site term is on par / slightly faster with T-gather
}

// Where the replica-resident kernel is the better one (measured, profiles/README.md): clusters of up to 2 CTAs, or larger ones when they
// need at most two waves over the chip (few replicas; run! of TG/h-BN 48x48 x 64 in clusters of 4 = 256 CTAs: 50 ms resident against 65 ms
// streaming, x 128 = 512 CTAs: 100 against 81 ms; psi family 128x128 x 64: equal, x 148: 196 against 145 ms).  Many replicas in clusters of 4
// or 8 (C5; TG/h-BN 48x48 x 1024: resident 735 us against streaming 360 us per sweep) belong to the streaming kernels -- for plain sweeps
// and for the device-side run! / annealing drivers alike.
cudaError_t scratch_get(scmc_handle* h, int slot, size_t bytes, void** p) {
    if (bytes == 0) bytes = 16;
    if (h->scratch_bytes[slot] < bytes) {
        if (h->scratch[slot]) {
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) return e;
            cudaFree(h->scratch[slot]);
            h->scratch[slot] = nullptr; h->scratch_bytes[slot] = 0;
        }
        cudaError_t e = cudaMalloc(&h->scratch[slot], bytes);
        if (e != cudaSuccess) { h->scratch[slot] = nullptr; return e; }
        h->scratch_bytes[slot] = bytes;      // (scratch is not part of the model footprint scmc_info reports)
    }
    *p = h->scratch[slot];
    return cudaSuccess;
}

void scratch_release(scmc_handle* h, int slot) {
    if (!h->scratch[slot]) return;
    cudaStreamSynchronize(h->stream);
    cudaFree(h->scratch[slot]);
    h->scratch[slot] = nullptr; h->scratch_bytes[slot] = 0;
}

bool resident_preferred(const scmc_handle* h) {
    if (h->res.cs < 1) return false;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
    return h->res.cs <= 2 || h->R * h->res.cs <= 2 * (int64_t)n_sm;
}

int32_t ensure_T(scmc_handle* h) {
    if (!h->T_stale) return SCMC_OK;
    h->T_stale = false;
    return dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        auto M = make_model<R, NL>(h);
        k_refresh_T<R, GEN, ONS, NL><<<grid_for(h->R * h->N), 256, 0, h->stream>>>(M);
        h->launches++;
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    });
}

static int32_t launch_colour_pass(scmc_handle* h, const SweepArgs& A, cudaStream_t stream, int lanes = 1) {
    const bool psi = !A.inj_xi && uses_psi_gather(h);
    if (!psi && stream == h->stream) SCMC_TRY(ensure_T(h));      // lanes: the caller refreshed T on h->stream before forking
    // TMA-pipelined form of the psi-gather pass (scmc_tma.cu): bit-identical chains, used wherever the class decomposes into bulk copies;
    // mode 3 keeps the direct-gather kernel (the reference point of the bit-identity tests)
    if (psi && h->sweep_mode != 3 && !h->async_pipeline && !(getenv("SCMC_MMA") && atoi(getenv("SCMC_MMA")) != 0) && tma_pass_available(h, A))
        return launch_colour_pass_tma(h, A, stream);
    return dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        int nc = 0;
        for (int c = A.colour; c <= A.colour_last; c++) nc = std::max(nc, h->col_off[c + 1] - h->col_off[c]);
        if (nc == 0 || A.rcount == 0) return SCMC_OK;
        auto M = make_model<R, NL>(h);
        for (int64_t y0 = 0; y0 < A.rcount; y0 += 65535) {      // gridDim.y limit
            SweepArgs B = A;
            B.r0 = A.r0 + y0; B.rcount = std::min<int64_t>(65535, A.rcount - y0);
            B.Jt = h->d_Jt;
            B.Jqt = h->fused_density ? h->d_Jqt : nullptr;
            B.lab_major = h->label_major ? 1 : 0;
            memcpy(B.lab_off, h->lab_off, sizeof B.lab_off);
            const unsigned tiles = grid_for(nc, SCMC_SWEEP_THREADS);
            // CTAs per replica: up to SCMC_SWEEP_TILES site tiles per CTA (coupling table staged once per CTA), but never fewer
            // than a few waves of CTAs over the chip, so small replica counts (8-GPU shards) do not end in a long tail.
            // Measured on B200 (profiles/README.md): 6000 CTAs per launch on one stream, 1500 per lane with two lanes.
            unsigned bpr = (tiles + SCMC_SWEEP_TILES - 1) / SCMC_SWEEP_TILES;
            static const long min_blocks_env = getenv("SCMC_MIN_BLOCKS") ? atol(getenv("SCMC_MIN_BLOCKS")) : 0;
            const long min_blocks = min_blocks_env ? min_blocks_env / lanes : (lanes == 2 ? 1500 : 6000);
            const unsigned min_bpr = (unsigned)((min_blocks + B.rcount - 1) / B.rcount);
            bpr = std::min(tiles, std::max(bpr, min_bpr));
            const dim3 grid(A.colour_last > A.colour ? 1u : bpr, (unsigned)B.rcount);
            if (A.inj_xi) {
                k_sweep_colour<R, GEN, ONS, NL, true><<<grid, SCMC_SWEEP_THREADS, 0, stream>>>(M, B);
            } else if (psi) {
                if constexpr (NL == 1) {
                    bool launched = false;
                    if constexpr (std::is_same<R, float>::value) {
                        if (h->z <= ASYNC_Z && h->async_pipeline) {
                            const size_t smem = (size_t)(h->z + 1) * (GenOps<GEN>::D / 2) * SCMC_SWEEP_THREADS * sizeof(float4);
                            SCMC_CK(cudaFuncSetAttribute(k_sweep_colour_psi_async<GEN, ONS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                            k_sweep_colour_psi_async<GEN, ONS><<<grid, SCMC_SWEEP_THREADS, smem, stream>>>(M, B);
                            launched = true;
                        }
                    }
                    if constexpr (std::is_same<R, float>::value && !ONS && GenOps<GEN>::NDENS == NG) {
                        static const bool use_mma = getenv("SCMC_MMA") && atoi(getenv("SCMC_MMA")) != 0;     // experiment, see matvec_mma
                        if (!launched && use_mma && B.Jqt != nullptr) {
                            k_sweep_colour_psi<R, GEN, ONS, true><<<grid, SCMC_SWEEP_THREADS, 0, stream>>>(M, B);
                            launched = true;
                        }
                    }
                    if (!launched) k_sweep_colour_psi<R, GEN, ONS><<<grid, SCMC_SWEEP_THREADS, 0, stream>>>(M, B);
                } else if constexpr (!ONS && GenOps<GEN>::NDENS == NG) {
                    B.Jqt = h->d_Jqt;                              // one table per label
                    k_sweep_colour_psi_nl<R, GEN, NL><<<grid, SCMC_SWEEP_THREADS, 0, stream>>>(M, B);
                }
                h->T_stale = true;
            } else if constexpr (std::is_same<R, float>::value && NL == 1) {
                if (h->z <= ASYNC_Z && h->async_pipeline) {
                    // pipelined kernel: a block walks over SCMC_ASYNC_TILES visits per thread
                    const size_t smem = (size_t)(h->z * 4 + GenOps<GEN>::D / 2) * SCMC_ASYNC_THREADS * sizeof(float4);
                    SCMC_CK(cudaFuncSetAttribute(k_sweep_colour_async<GEN, ONS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                    const unsigned tiles2 = grid_for(nc, SCMC_ASYNC_THREADS);
                    const dim3 g2(A.colour_last > A.colour ? 1u : (tiles2 + SCMC_ASYNC_TILES - 1) / SCMC_ASYNC_TILES, (unsigned)B.rcount);
                    k_sweep_colour_async<GEN, ONS><<<g2, SCMC_ASYNC_THREADS, smem, stream>>>(M, B);
                } else {
                    k_sweep_colour<R, GEN, ONS, NL, false><<<grid, SCMC_SWEEP_THREADS, 0, stream>>>(M, B);
                }
            } else {
                k_sweep_colour<R, GEN, ONS, NL, false><<<grid, SCMC_SWEEP_THREADS, 0, stream>>>(M, B);
            }
            h->launches++;
        }
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    });
}

// Streaming path: replicas are processed in batches (sized to stay L2-resident across the colour passes of all
// n_sweeps when h->batch > 0); within a batch, sweeps x colours launches.
int32_t launch_sweeps(scmc_handle* h, int64_t n_sweeps, int32_t hits, uint64_t seed, uint64_t sweep_counter, const uint8_t* active) {
    if (h->sweep_mode == 2) {
        if (!h->res.eligible) { set_error("resident T-gather kernel requested but this lattice does not fit in cluster shared memory"); return SCMC_ERR_UNSUPPORTED; }
        SCMC_TRY(ensure_T(h));
        return launch_resident(h, n_sweeps, hits, seed, sweep_counter, active, false);
    }
    if (h->sweep_mode == 4) {
        if (!h->res.eligible_psi) { set_error("resident psi-gather kernel requested but not available for this model / lattice"); return SCMC_ERR_UNSUPPORTED; }
        return launch_resident(h, n_sweeps, hits, seed, sweep_counter, active, true);
    }
    if (h->sweep_mode == 0) {
        // Kernel choice inside one arithmetic family (measured on B200, profiles/README.md): the resident kernel wins when colour
        // passes are small (launch-bound) or many sweeps run per call; the streaming kernel wins for big batches of large lattices.
        const bool psi = uses_psi_gather(h);
        const int64_t work_per_pass = h->R * h->N / std::max(1, h->n_col);
        const bool want_resident = resident_preferred(h) && (n_sweeps >= 4 || work_per_pass < 300000);
        if (want_resident && psi && h->res.eligible_psi) return launch_resident(h, n_sweeps, hits, seed, sweep_counter, active, true);
        if (want_resident && !psi && h->res.eligible) { SCMC_TRY(ensure_T(h)); return launch_resident(h, n_sweeps, hits, seed, sweep_counter, active, false); }
    }
    const int64_t batch = h->batch > 0 ? std::min<int64_t>(h->batch, h->R) : h->R;
    if (!uses_psi_gather(h)) SCMC_TRY(ensure_T(h));      // on h->stream, before any lane forks
    for (int64_t r0 = 0; r0 < h->R; r0 += batch) {
        const int64_t rc = std::min<int64_t>(batch, h->R - r0);
        // trailing colour classes that are small (seam repairs) are visited by one block per replica in a single launch
        int first_small = h->n_col;
        while (first_small > 0 && h->col_off[first_small] - h->col_off[first_small - 1] <= 2 * SCMC_SWEEP_THREADS) first_small--;      // tiny lattices: the whole sweep is ONE launch
        if (first_small == h->n_col - 1) first_small = h->n_col;      // a single small class gains nothing
        // two replica lanes on side streams when a call is a chain of several dependent launches: every launch ends in a tail
        // of partially filled SMs, which the other lane's launches fill (replicas never interact)
        static const int lanes_env = getenv("SCMC_LANES") ? atoi(getenv("SCMC_LANES")) : 0;
        int lanes = h->lanes ? h->lanes : (lanes_env ? lanes_env : 2);
        const int64_t n_launch = n_sweeps * (first_small < h->n_col ? first_small + 1 : h->n_col);
        if (rc < 2 || n_launch < 2 || !h->lane_stream[0]) lanes = 1;
        cudaStream_t st[2] = {h->stream, h->stream};
        int64_t lane_r0[2] = {r0, r0}, lane_rc[2] = {rc, 0};
        if (lanes == 2) {
            lane_rc[0] = rc / 2; lane_r0[1] = r0 + lane_rc[0]; lane_rc[1] = rc - lane_rc[0];
            SCMC_CK(cudaEventRecord(h->ev_fork, h->stream));
            for (int l = 0; l < 2; l++) {
                st[l] = h->lane_stream[l];
                SCMC_CK(cudaStreamWaitEvent(st[l], h->ev_fork, 0));
            }
        }
        int32_t status = SCMC_OK;
        for (int64_t s = 0; s < n_sweeps && status == SCMC_OK; s++)
            for (int c = 0; c < h->n_col && status == SCMC_OK; c++) {
                SweepArgs A{};
                A.colour = c; A.colour_last = c; A.hits = hits;
                if (c >= first_small) { A.colour_last = h->n_col - 1; }
                A.visit0 = (sweep_counter + (uint64_t)s) * (uint64_t)hits; A.seed = seed; A.roff = h->roff; A.stream = h->stream_version;
                A.beta = h->beta; A.sigma = h->sigma; A.active = active; A.acc = h->acc;
                for (int l = 0; l < lanes && status == SCMC_OK; l++) {
                    A.r0 = lane_r0[l]; A.rcount = lane_rc[l];
                    status = launch_colour_pass(h, A, st[l], lanes);
                }
                if (A.colour_last > c) c = A.colour_last;
            }
        if (lanes == 2) {                                 // join even after a failed launch: h->stream must not run ahead of the lanes
            for (int l = 0; l < 2; l++) {
                cudaError_t e = cudaEventRecord(h->ev_join[l], st[l]);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_join[l], 0);
                if (e != cudaSuccess && status == SCMC_OK) status = cuda_fail(e, "lane join", __FILE__, __LINE__);
            }
        }
        SCMC_TRY(status);
    }
    return SCMC_OK;
}

int32_t launch_measure(scmc_handle* h, double* d_rows, double* d_Etot) {
    const bool psi = uses_psi_gather(h) && h->n_labels == 1;      // several labels: the T-family measurement kernel on the (refreshed) T planes
    if (!psi) SCMC_TRY(ensure_T(h));
    // CTAs per replica: SCMC_SWEEP_TILES site tiles each, at most 64, and enough CTAs over all replicas to fill the chip
    const int64_t tiles = (h->N + SCMC_MEASURE_THREADS - 1) / SCMC_MEASURE_THREADS;
    int64_t parts = std::min<int64_t>(64, (tiles + SCMC_SWEEP_TILES - 1) / SCMC_SWEEP_TILES);
    parts = std::min<int64_t>(tiles, std::max<int64_t>(parts, std::min<int64_t>(64, (2000 + h->R - 1) / h->R)));
    const size_t need = (size_t)h->R * parts * (NG + 1) * sizeof(double);
    if (need > h->meas_partial_bytes) {
        if (h->d_meas_partial) { SCMC_CK(cudaStreamSynchronize(h->stream)); SCMC_CK(cudaFree(h->d_meas_partial)); h->d_meas_partial = nullptr; h->meas_partial_bytes = 0; }
        SCMC_CK(cudaMalloc(&h->d_meas_partial, need));
        h->meas_partial_bytes = need;
    }
    return dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        auto M = make_model<R, NL>(h);
        bool fused = false;
        for (int64_t y0 = 0; y0 < h->R; y0 += 65535) {      // gridDim.y limit
            const int64_t yc = std::min<int64_t>(65535, h->R - y0);
            auto My = M;                                     // replica window: blockIdx.y counts from y0
            My.psi = M.psi + y0 * M.Ns;
            My.T = M.T + y0 * M.Ns;
            if (M.Tsq) My.Tsq = M.Tsq + y0 * M.Ns;
            double* part = h->d_meas_partial + (size_t)y0 * parts * (NG + 1);
            const dim3 grid((unsigned)parts, (unsigned)yc);
            bool launched = false;
            if constexpr (NL == 1) {
                if (psi) {
                    fused = h->fused_density && !ONS && GenOps<GEN>::NDENS == NG;
                    uint64_t fs = 0;
                    for (int t = 0; t < h->n_fwd && t < 8; t++) fs |= (uint64_t)h->fwd_slot[t] << (8 * t);
                    const bool fwd_off = getenv("SCMC_MEASURE_FWD") && atoi(getenv("SCMC_MEASURE_FWD")) == 0;      // (read per call: a test switches it)
                    k_measure_psi<R, GEN, ONS><<<grid, SCMC_MEASURE_THREADS, 0, h->stream>>>(My, h->d_Jt, fused ? h->d_Jqt : nullptr, part, fwd_off ? 0 : h->n_fwd, fs);
                    launched = true;
                }
            }
            if (!launched) {
                SweepArgs MA{};
                MA.Jt = h->d_Jt; MA.lab_major = h->label_major ? 1 : 0;
                memcpy(MA.lab_off, h->lab_off, sizeof MA.lab_off);
                k_measure<R, GEN, ONS, NL><<<grid, SCMC_MEASURE_THREADS, 0, h->stream>>>(My, part, MA);
            }
            h->launches++;
        }
        k_measure_finish<GEN><<<grid_for(h->R, 128), 128, 0, h->stream>>>(h->R, (int)parts, fused ? 1 : 0, h->N, h->e_const, h->d_meas_partial, d_rows, d_Etot);
        h->launches++;
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    });
}

}  // namespace scmc

using namespace scmc;

// ================================================================================================ C ABI
extern "C" {

const char* scmc_last_error(void) { return g_err.c_str(); }
int32_t scmc_version(void) { return 100; }
int32_t scmc_device_count(int32_t* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (n) *n = (e == cudaSuccess) ? c : 0;
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    return SCMC_OK;
}

static bool table_matches(const double* re, const double* im, const int8_t* tre, const int8_t* tim, int n) {
    for (int i = 0; i < n; i++)
        if (std::fabs(re[i] - tre[i]) > 1e-12 || std::fabs(im[i] - tim[i]) > 1e-12) return false;
    return true;
}

int32_t scmc_create(scmc_t** out, int32_t device, int32_t precision, int64_t n_sites, int32_t d, int64_t n_replicas,
                    int64_t replica_offset, int32_t z_max, const int32_t* nbr_site, const int32_t* nbr_label, int32_t n_labels,
                    const double* J, const double* K, const double* gen_re, const double* gen_im, const double* gensq_re,
                    const double* gensq_im, const int32_t* colour) {
    SCMC_ARG(out != nullptr, "scmc_create: null handle pointer");
    *out = nullptr;
    SCMC_ARG(precision == 32 || precision == 64, "scmc_create: precision must be 32 or 64");
    SCMC_ARG(n_sites > 0 && n_sites <= (1ll << 27), "scmc_create: n_sites out of range (1 .. 2^27)");
    SCMC_ARG(n_replicas > 0, "scmc_create: n_replicas must be positive");
    SCMC_ARG(z_max > 0 && nbr_site && nbr_label, "scmc_create: neighbour tables required");
    SCMC_ARG(n_labels >= 1 && n_labels <= MAXL, "scmc_create: more coupling labels than this build supports (1..SCMC_MAXL, default 3; rebuild with -DSCMC_MAXL=<n>)");
    SCMC_ARG(J && K && gen_re && gen_im && gensq_re && gensq_im, "scmc_create: null coupling / generator table");
    SCMC_ARG(d == 4 || d == 6, "scmc_create: local dimension d must be 4 (n=1,3) or 6 (n=2)");

    int gen_id = 0;
    if (d == 4 && table_matches(gen_re, gen_im, gen::G1_re, gen::G1_im, 256) && table_matches(gensq_re, gensq_im, gen::Gsq1_re, gen::Gsq1_im, 256)) gen_id = 1;
    else if (d == 4 && table_matches(gen_re, gen_im, gen::G3_re, gen::G3_im, 256) && table_matches(gensq_re, gensq_im, gen::Gsq3_re, gen::Gsq3_im, 256)) gen_id = 3;
    else if (d == 6 && table_matches(gen_re, gen_im, gen::G2_re, gen::G2_im, 576) && table_matches(gensq_re, gensq_im, gen::Gsq2_re, gen::Gsq2_im, 576)) gen_id = 2;
    if (!gen_id) {
        // caller-supplied generators (e.g. the parton operator with its Jordan-Wigner sign): dense-table path, T-gather family
        for (int a = 0; a < 16; a++)
            for (int j = 0; j < d; j++)
                for (int k = 0; k < d; k++) {
                    const size_t jk = ((size_t)a * d + j) * d + k, kj = ((size_t)a * d + k) * d + j;
                    SCMC_ARG(std::fabs(gen_re[jk] - gen_re[kj]) < 1e-12 && std::fabs(gen_im[jk] + gen_im[kj]) < 1e-12 &&
                             std::fabs(gensq_re[jk] - gensq_re[kj]) < 1e-12 && std::fabs(gensq_im[jk] + gensq_im[kj]) < 1e-12,
                             "scmc_create: generator tables must be Hermitian matrices");
                }
        gen_id = d == 4 ? 4 : 5;
    }

    const int64_t N = n_sites;
    // --- table validation
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < z_max; k++) {
            const int32_t j = nbr_site[i * z_max + k];
            SCMC_ARG(j >= -1 && j < N, "scmc_create: neighbour index out of range");
            if (j >= 0) SCMC_ARG(nbr_label[i * z_max + k] >= 0 && nbr_label[i * z_max + k] < n_labels, "scmc_create: bond label out of range");
            SCMC_ARG(j != i, "scmc_create: self-bonds are not supported (use the on-site couplings K)");
        }
    // every bond i->j (label l) needs a reverse bond j->i whose coupling is the transpose
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < z_max; k++) {
            const int32_t j = nbr_site[i * z_max + k];
            if (j < 0) continue;
            const double* Jl = J + (size_t)nbr_label[i * z_max + k] * 256;
            int cnt_fwd = 0, cnt_ok = 0;
            for (int k2 = 0; k2 < z_max; k2++)
                if (nbr_site[i * z_max + k2] == j && nbr_label[i * z_max + k2] == nbr_label[i * z_max + k]) cnt_fwd++;
            for (int k2 = 0; k2 < z_max; k2++) {
                if (nbr_site[(int64_t)j * z_max + k2] != i) continue;
                const double* Jr = J + (size_t)nbr_label[(int64_t)j * z_max + k2] * 256;
                bool tr = true;
                for (int a = 0; a < 16 && tr; a++)
                    for (int b = 0; b < 16; b++)
                        if (std::fabs(Jl[a * 16 + b] - Jr[b * 16 + a]) > 1e-12 * (1.0 + std::fabs(Jl[a * 16 + b]))) { tr = false; break; }
                if (tr) cnt_ok++;
            }
            SCMC_ARG(cnt_ok >= cnt_fwd, "scmc_create: bond i->j has no reverse bond j->i with the transposed coupling matrix "
                                        "(getEnergyDifference! would be inconsistent with getEnergy, Configuration.jl:246-250 vs 292-296)");
        }

    scmc_handle* h = new scmc_handle();
    h->device = device; h->prec = precision; h->d = d; h->z = z_max; h->n_labels = n_labels; h->gen_id = gen_id;
    h->N = N; h->R = n_replicas; h->roff = replica_offset;
    if (const char* e = getenv("SCMC_ASYNC")) h->async_pipeline = atoi(e);      // kernel experiments: 0 = direct-load streaming kernel
    h->Ns = (N + 1 + 7) / 8 * 8;          // site N of every replica is the all-zero padding site
    h->z = (z_max + 2) / 3 * 3;           // neighbour slots padded to a multiple of 3 (chunked loads)
    for (int l = 0; l < n_labels; l++) memcpy(h->J[l], J + (size_t)l * 256, sizeof(double) * 256);
    memcpy(h->K, K, sizeof(double) * 16);
    // One label (J = J^T by the check above): look for z / 2 "forward" slots such that every bond i-j sits in a forward slot at exactly one of its
    // ends (Bravais lattices: three of the six directions).  The energy measurement then gathers half the neighbours (k_measure_psi).
    if (n_labels == 1 && z_max % 2 == 0 && z_max >= 2 && z_max <= 8) {
        std::vector<int8_t> rev((size_t)N * z_max, -1);      // slot of the reverse bond at the neighbour, -2: ambiguous (double bond)
        bool simple = true;
        for (int64_t i = 0; i < N && simple; i++)
            for (int k = 0; k < z_max; k++) {
                const int32_t j = nbr_site[i * z_max + k];
                if (j < 0) { simple = false; break; }      // a missing bond moves the later bonds of that site to other device slots: keep all slots
                int found = -1, cnt = 0;
                for (int k2 = 0; k2 < z_max; k2++)
                    if (nbr_site[(int64_t)j * z_max + k2] == i) { found = k2; cnt++; }
                if (cnt != 1) { simple = false; break; }
                rev[i * z_max + k] = (int8_t)found;
            }
        for (unsigned mask = 0; simple && mask < (1u << z_max) && h->n_fwd == 0; mask++) {
            if (__builtin_popcount(mask) != z_max / 2) continue;
            bool ok = true;
            for (int64_t i = 0; i < N && ok; i++)
                for (int k = 0; k < z_max; k++) {
                    if (nbr_site[i * z_max + k] < 0) continue;
                    if (((mask >> k) & 1u) == ((mask >> rev[i * z_max + k]) & 1u)) { ok = false; break; }
                }
            if (ok) {
                for (int k = 0; k < z_max; k++)
                    if ((mask >> k) & 1u) h->fwd_slot[h->n_fwd++] = (uint8_t)k;
            }
        }
    }
    bool anyK = false;
    double sumK = 0.0;
    for (int a = 0; a < 16; a++) { anyK |= (K[a] != 0.0); sumK += K[a]; }
    h->onsite = (anyK && gen_id == 2) || gen_id >= 4;
    // n = 1: (G^a)^2 = 1; n = 3: (G^a)^2 = c_a 1 with c_1 = 9 (S(0,0) = 3 1): the on-site energy is the constant N sum_a K^a c_a
    h->e_const = 0.0;
    if (gen_id == 1 || gen_id == 3)
        for (int a = 0; a < 16; a++) h->e_const += K[a] * gensq_re[(size_t)a * d * d] * (double)N;
    (void)sumK;
    if (gen_id >= 4) {
        h->custom = new CustomTables();
        memset(h->custom, 0, sizeof(CustomTables));
        const size_t nt = (size_t)16 * d * d;
        memcpy(h->custom->gre, gen_re, nt * 8); memcpy(h->custom->gim, gen_im, nt * 8);
        memcpy(h->custom->qre, gensq_re, nt * 8); memcpy(h->custom->qim, gensq_im, nt * 8);
    }

    // --- colouring
    h->colour.assign(N, -1);
    if (colour) {
        for (int64_t i = 0; i < N; i++) h->colour[i] = colour[i];
    } else {
        std::vector<std::vector<int32_t>> adj(N);
        for (int64_t i = 0; i < N; i++)
            for (int k = 0; k < z_max; k++) {
                const int32_t j = nbr_site[i * z_max + k];
                if (j >= 0) { adj[i].push_back(j); adj[j].push_back((int32_t)i); }
            }
        for (int64_t i = 0; i < N; i++) {
            unsigned used = 0;
            for (int32_t j : adj[i]) if (h->colour[j] >= 0 && h->colour[j] < 32) used |= 1u << h->colour[j];
            int c = 0;
            while (used & (1u << c)) c++;
            h->colour[i] = c;
        }
    }
    int ncol = 0;
    for (int64_t i = 0; i < N; i++) {
        if (h->colour[i] < 0 || h->colour[i] >= MAXC) { delete h; set_error("scmc_create: colour outside 0..15 (too many colours)"); return SCMC_ERR_ARG; }
        ncol = std::max(ncol, h->colour[i] + 1);
    }
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < z_max; k++) {
            const int32_t j = nbr_site[i * z_max + k];
            if (j >= 0 && h->colour[j] == h->colour[i]) { delete h; set_error("scmc_create: invalid colouring, two bonded sites share a colour"); return SCMC_ERR_ARG; }
        }
    h->n_col = ncol;
    h->perm.resize(N); h->orig.resize(N);
    // Site order on the device: (owner CTA of the resident kernel, colour, caller's index).  The owner is a contiguous
    // block of the caller's numbering (a strip of the lattice), so that a CTA's chunk is contiguous in every plane; with one
    // CTA per replica this is plain colour-major order.  col_list groups the permuted indices by colour for the streaming kernel.
    std::vector<int> owner_of_perm(N, 0);
    std::vector<int32_t> col_list(N);
    {
        const size_t w = precision == 64 ? 8 : 4;
        // footprint per site of the arithmetic family this model uses in auto mode (uses_psi_gather): psi only, or psi + T (+ Tsq)
        const bool psi_family = n_labels == 1 && (d == 4 || h->onsite);
        const size_t per_site = (psi_family ? 4 * w * (d / 2) : 4 * w * (4 + d / 2 + (h->onsite ? 4 : 0))) + 2 * (size_t)h->z + (n_labels > 1 ? (size_t)h->z : 0);
        int cs = 0;
        const int cs_min = getenv("SCMC_CLUSTER_MIN") ? atoi(getenv("SCMC_CLUSTER_MIN")) : 1;      // experiments: larger clusters than needed
        for (int c : {1, 2, 4, 8}) {
            if (c < cs_min || gen_id >= 4) continue;      // caller-supplied generators: streaming kernels only
            const int64_t n_own = (N + c - 1) / c;
            const size_t cap = (size_t)(n_own + 1 + 7) / 8 * 8;
            if (cap * per_site + (size_t)MAXL * 256 * w + 1024 <= (size_t)226 * 1024 && n_own + 1 <= 8191 && N >= c) { cs = c; break; }
        }
        // Few replicas (the run! and annealing drivers of both families work inside clusters): grow the cluster while all clusters still fit
        // on the chip in one wave and a CTA keeps >= 512 sites.  Measured (profiles/README.md, cluster size for few replicas): full run! of
        // 48x48 x 64 replicas 47.5 -> 23.9 ms at cluster 2, x 16 replicas 38.2 -> 16.5 ms at cluster 4; a 72-site lattice loses 35 % in clusters
        // and a full chip (148 replicas) loses 12-27 %.  Chains do not depend on the cluster size.
        if (cs >= 1 && !(getenv("SCMC_CLUSTER_FILL") && atoi(getenv("SCMC_CLUSTER_FILL")) == 0)) {
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
            while (cs < 8 && n_replicas * (int64_t)(2 * cs) <= n_sm && N / (2 * cs) >= 512) cs *= 2;
        }
        const int64_t n_nom = cs ? (N + cs - 1) / cs : N;
        int32_t at = 0;
        for (int b = 0; b < (cs ? cs : 1); b++)
            for (int c = 0; c < ncol; c++)
                for (int64_t i = b * n_nom; i < std::min<int64_t>(N, (b + 1) * n_nom); i++)
                    if (h->colour[i] == c) { h->perm[i] = at; h->orig[at] = (int32_t)i; owner_of_perm[at] = b; at++; }
        h->res.cs = cs;
        int32_t cl = 0;
        for (int c = 0; c < ncol; c++) {
            h->col_off[c] = cl;
            for (int32_t q = 0; q < N; q++) if (h->colour[h->orig[q]] == c) col_list[cl++] = q;
        }
        for (int c = ncol; c <= MAXC; c++) h->col_off[c] = cl;
    }

    // --- device
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); }
    std::vector<int32_t> nb((size_t)h->z * N, (int32_t)N);
    std::vector<int8_t> lb((size_t)h->z * N, 0);
    // slots of a site are stably sorted by label (bonds that do not exist last): the order of the bonds inside a label, and with it every
    // per-label neighbour sum, is the caller's; with one label this is the caller's order
    std::vector<int64_t> per_label((size_t)n_labels, -1);
    h->label_major = n_labels > 1 && !(getenv("SCMC_LABEL_MAJOR") && atoi(getenv("SCMC_LABEL_MAJOR")) == 0);
    std::vector<int> order(z_max);
    for (int64_t p = 0; p < N; p++) {
        const int64_t i = h->orig[p];
        auto key = [&](int k) { return nbr_site[i * z_max + k] >= 0 ? (int)nbr_label[i * z_max + k] : n_labels; };
        for (int k = 0; k < z_max; k++) order[k] = k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
        std::vector<int64_t> cnt((size_t)n_labels, 0);
        for (int s = 0; s < z_max; s++) {
            const int k = order[s];
            const int32_t j = nbr_site[i * z_max + k];
            nb[(size_t)s * N + p] = j >= 0 ? h->perm[j] : (int32_t)N;
            lb[(size_t)s * N + p] = j >= 0 ? (int8_t)nbr_label[i * z_max + k] : 0;
            if (j >= 0) cnt[nbr_label[i * z_max + k]]++;
        }
        for (int l = 0; l < n_labels; l++) {
            if (per_label[l] < 0) per_label[l] = cnt[l];
            else if (per_label[l] != cnt[l]) h->label_major = false;      // bond counts per label differ between sites: per-thread labels
        }
    }
    h->lab_off[0] = 0;
    for (int l = 0; l < MAXL; l++) h->lab_off[l + 1] = h->lab_off[l] + (l < n_labels ? (int32_t)std::max<int64_t>(per_label[l], 0) : 0);
    const size_t plane = (size_t)h->R * h->Ns * 4 * h->esz();
    int32_t st = SCMC_OK;
    auto A = [&](void** p, size_t bytes) { if (st == SCMC_OK) st = dev_alloc(h, p, bytes); };
    A(&h->psi, plane * (d / 2));
    A(&h->T, plane * 4);
    if (h->onsite) A(&h->Tsq, plane * 4);
    A((void**)&h->nbr, nb.size() * sizeof(int32_t));
    A((void**)&h->lab, lb.size());
    A((void**)&h->d_orig, N * sizeof(int32_t));
    A((void**)&h->d_col_list, N * sizeof(int32_t));
    A((void**)&h->d_cls_site, N * sizeof(int32_t));
    A((void**)&h->d_cls_nbr, (size_t)h->z * N * sizeof(int32_t));
    A(&h->d_Jt, (size_t)MAXL * 256 * h->esz());
    A(&h->d_Jqt, (size_t)MAXL * 256 * h->esz());
    A(&h->beta, h->R * h->esz());
    A(&h->sigma, h->R * h->esz());
    A((void**)&h->acc, h->R * sizeof(unsigned long long));
    A((void**)&h->active, h->R);
    A((void**)&h->obs, h->R * SCMC_NOBS * sizeof(double));
    if (st == SCMC_OK && cudaMallocHost(&h->pinned, 2 * h->R * 8) != cudaSuccess) { set_error("cudaMallocHost failed"); st = SCMC_ERR_ALLOC; }
    if (st != SCMC_OK) { scmc_destroy(h); return st; }
    cudaError_t ce = cudaSuccess;
    auto C = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
    C(cudaMemcpy(h->nbr, nb.data(), nb.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    C(cudaMemcpy(h->lab, lb.data(), lb.size(), cudaMemcpyHostToDevice));
    C(cudaMemcpy(h->d_orig, h->orig.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice));
    C(cudaMemcpy(h->d_col_list, col_list.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice));
    {
        std::vector<int32_t> cs_site(N), cs_nbr((size_t)h->z * N);
        for (int64_t cp = 0; cp < N; cp++) {
            cs_site[cp] = h->orig[col_list[cp]];
            for (int k = 0; k < h->z; k++) cs_nbr[(size_t)k * N + cp] = nb[(size_t)k * N + col_list[cp]];
        }
        C(cudaMemcpy(h->d_cls_site, cs_site.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice));
        C(cudaMemcpy(h->d_cls_nbr, cs_nbr.data(), cs_nbr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    {
        std::vector<char> jt((size_t)MAXL * 256 * h->esz(), 0);
        for (int l = 0; l < n_labels; l++)
            for (int a = 0; a < 16; a++)
                for (int b = 0; b < 16; b++) {
                    const double v = h->J[l][a * 16 + b];
                    if (precision == 64) ((double*)jt.data())[(l * 16 + b) * 16 + a] = v;
                    else ((float*)jt.data())[(l * 16 + b) * 16 + a] = (float)v;
                }
        C(cudaMemcpy(h->d_Jt, jt.data(), jt.size(), cudaMemcpyHostToDevice));
    }
    if (gen_id == 1 || gen_id == 3) {     // 16 density entries: fuse map, couplings and transpose map into one 16x16 matrix per label
        const int8_t* Cm = gen_id == 1 ? gen::DENSMAP1 : gen::DENSMAP3;
        std::vector<char> jq((size_t)MAXL * 256 * h->esz(), 0);
        for (int l = 0; l < n_labels; l++)
            for (int n = 0; n < 16; n++)
                for (int m = 0; m < 16; m++) {
                    double v = 0.0;
                    for (int a = 0; a < 16; a++)
                        for (int b = 0; b < 16; b++) v += (double)Cm[a * 16 + n] * h->J[l][a * 16 + b] * (double)Cm[b * 16 + m];
                    if (precision == 64) ((double*)jq.data())[(l * 16 + m) * 16 + n] = v;      // transposed: [l][m][n] = Jq_l[n][m]
                    else ((float*)jq.data())[(l * 16 + m) * 16 + n] = (float)v;
                }
        C(cudaMemcpy(h->d_Jqt, jq.data(), jq.size(), cudaMemcpyHostToDevice));
        h->fused_density = n_labels == 1;
        // several labels (TG/h-BN at n = 1, 3): neighbours enter through their states, one density sum and one 16x16 mat-vec per label
        h->psi_labels = n_labels > 1 && h->label_major && !h->onsite && !(getenv("SCMC_PSI_LABELS") && atoi(getenv("SCMC_PSI_LABELS")) == 0);
    }
    C(cudaMemset(h->psi, 0, plane * (d / 2)));
    C(cudaMemset(h->T, 0, plane * 4));
    if (h->onsite) C(cudaMemset(h->Tsq, 0, plane * 4));
    C(cudaMemset(h->acc, 0, h->R * sizeof(unsigned long long)));
    C(cudaMemset(h->active, 1, h->R));
    if (ce != cudaSuccess) { scmc_destroy(h); return cuda_fail(ce, "scmc_create upload", __FILE__, __LINE__); }
    for (int l = 0; l < 2; l++) {
        C(cudaStreamCreateWithFlags(&h->lane_stream[l], cudaStreamNonBlocking));
        C(cudaEventCreateWithFlags(&h->ev_join[l], cudaEventDisableTiming));
    }
    C(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    if (ce != cudaSuccess) { scmc_destroy(h); return cuda_fail(ce, "scmc_create lanes", __FILE__, __LINE__); }
    plan_resident(h, owner_of_perm);     // neighbour codes, launch geometry, eligibility of the replica-resident kernel
    plan_tma(h, col_list, nb);           // tiles, bulk-copy runs and stage slots of the TMA-pipelined streaming kernel
    *out = h;
    return SCMC_OK;
}

int32_t scmc_destroy(scmc_t* h) {
    if (!h) return SCMC_OK;
    cudaSetDevice(h->device);
    void* ptrs[] = {h->sf_acc, h->d_meas_partial, h->psi, h->T, h->Tsq, h->nbr, h->lab, h->d_orig, h->beta, h->sigma, h->acc, h->active, h->obs, h->d_Jt, h->d_Jqt, h->d_col_list, h->d_nbr16, h->d_cls_site, h->d_cls_nbr, h->tma.d_meta, h->tma.d_runs, h->tma.d_tile};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < scmc_handle::N_SCRATCH; i++) if (h->scratch[i]) cudaFree(h->scratch[i]);
    if (h->pinned) cudaFreeHost(h->pinned);
    delete h->custom;
    for (int l = 0; l < 2; l++) {
        if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
        if (h->ev_join[l]) cudaEventDestroy(h->ev_join[l]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    delete h;
    return SCMC_OK;
}

int32_t scmc_set_stream(scmc_t* h, void* s) {
    SCMC_ARG(h, "null handle");
    h->stream = (cudaStream_t)s;
    return SCMC_OK;
}

int32_t scmc_verify_tables(scmc_t* h, int64_t* n_errors) {
    SCMC_ARG(h && n_errors, "null argument");
    SCMC_CK(cudaSetDevice(h->device));
    void *d_cls = nullptr, *d_err = nullptr, *d_off = nullptr;
    SCMC_CK(scratch_get(h, 0, (size_t)(h->N + 1) * sizeof(int32_t), &d_cls));
    SCMC_CK(scratch_get(h, 1, sizeof(unsigned long long), &d_err));
    SCMC_CK(scratch_get(h, 2, sizeof(h->col_off), &d_off));
    SCMC_CK(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), h->stream));
    SCMC_CK(cudaMemcpyAsync(d_off, h->col_off, sizeof(h->col_off), cudaMemcpyHostToDevice, h->stream));
    const int N = (int)h->N;
    // self-test of the checker (tests only): SCMC_VERIFY_PLANT=1 points slot 0 of permuted site 0 at that site itself for the duration of the check
    // (a bond inside a colour class without a reverse bond, which also breaks the class-position table and the tile plan), then restores the entry
    const bool plant = getenv("SCMC_VERIFY_PLANT") && atoi(getenv("SCMC_VERIFY_PLANT")) != 0;
    int32_t saved = 0;
    if (plant) {
        SCMC_CK(cudaStreamSynchronize(h->stream));
        SCMC_CK(cudaMemcpy(&saved, h->nbr, sizeof saved, cudaMemcpyDeviceToHost));
        const int32_t self = 0;
        SCMC_CK(cudaMemcpy(h->nbr, &self, sizeof self, cudaMemcpyHostToDevice));
    }
    k_verify_class_of<<<grid_for(N), 256, 0, h->stream>>>(h->n_col, (const int32_t*)d_off, h->d_col_list, (int32_t*)d_cls);
    k_verify_tables<<<grid_for(N), 256, 0, h->stream>>>(N, h->z, h->n_col, (const int32_t*)d_off, h->nbr, h->d_col_list, h->d_cls_nbr, h->d_cls_site, h->d_orig,
                                                        (const int32_t*)d_cls, (unsigned long long*)d_err);
    if (h->tma.ready) {
        int n_tiles = 0;
        for (int c = 0; c < h->n_col; c++)
            if (h->tma.cls[c].ok) n_tiles = std::max(n_tiles, h->tma.cls[c].first + h->tma.cls[c].n_tiles);
        if (n_tiles > 0)
            k_verify_tma_plan<<<n_tiles, TMA_TILE, 0, h->stream>>>(N, h->z, h->tma.cap, (const uint4*)h->tma.d_meta, (const uint2*)h->tma.d_runs, h->tma.d_tile, h->nbr,
                                                                   h->d_orig, h->d_col_list, (unsigned long long*)d_err);
    }
    h->launches += 3;
    SCMC_CK(cudaGetLastError());
    unsigned long long e = 0;
    SCMC_CK(cudaMemcpyAsync(&e, d_err, sizeof e, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    if (plant) SCMC_CK(cudaMemcpy(h->nbr, &saved, sizeof saved, cudaMemcpyHostToDevice));
    *n_errors = (int64_t)e;
    return SCMC_OK;
}

int32_t scmc_set_lanes(scmc_t* h, int32_t lanes) {
    SCMC_ARG(h, "null handle");
    SCMC_ARG(lanes >= 0 && lanes <= 2, "scmc_set_lanes: lanes must be 0 (auto), 1 or 2");
    h->lanes = lanes;
    return SCMC_OK;
}

int32_t scmc_info(scmc_t* h, int64_t* info) {
    SCMC_ARG(h && info, "null argument");
    info[0] = h->n_col; info[1] = h->gen_id; info[2] = h->bytes; info[3] = h->res.eligible | (h->res.eligible_psi << 1) | (h->tma.ready << 2) | ((h->tma.ready && h->prec == 32 && tma_use_tc(h)) ? 8 : 0); info[4] = h->res.cs; info[5] = h->launches; info[6] = h->N; info[7] = h->R;
    return SCMC_OK;
}

int32_t scmc_set_stream_version(scmc_t* h, int32_t version) {
    SCMC_ARG(h, "null handle");
    SCMC_ARG(version == 1 || version == 2, "scmc_set_stream_version: version must be 1 or 2");
    h->stream_version = version;
    return SCMC_OK;
}

int32_t scmc_set_sweep_mode(scmc_t* h, int32_t mode, int64_t batch) {
    SCMC_ARG(h, "null handle");
    SCMC_ARG(mode >= 0 && mode <= 6 && batch >= 0, "scmc_set_sweep_mode: bad mode / batch");
    h->sweep_mode = mode; h->batch = batch;
    return SCMC_OK;
}

static int32_t state_io(scmc_t* h, int64_t r0, int64_t count, double* re, double* im, bool to_device) {
    SCMC_ARG(h && re && im, "null argument");
    SCMC_ARG(r0 >= 0 && count >= 0 && r0 + count <= h->R, "replica range out of bounds");
    SCMC_CK(cudaSetDevice(h->device));
    const int64_t per = h->N * h->d;                               // doubles per replica per part
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(128ll << 20) / (per * 8));
    double *dre = nullptr, *dim_ = nullptr;
    const size_t cbytes = (size_t)std::min(chunk, std::max<int64_t>(count, 1)) * per * 8;
    SCMC_CK(cudaMalloc(&dre, cbytes));
    if (cudaMalloc(&dim_, cbytes) != cudaSuccess) { cudaFree(dre); return cuda_fail(cudaErrorMemoryAllocation, "staging", __FILE__, __LINE__); }
    int32_t st = SCMC_OK;
    for (int64_t q = 0; q < count && st == SCMC_OK; q += chunk) {
        const int64_t c = std::min(chunk, count - q);
        const size_t bytes = (size_t)c * per * 8;
        if (to_device) {
            cudaMemcpyAsync(dre, re + q * per, bytes, cudaMemcpyHostToDevice, h->stream);
            cudaMemcpyAsync(dim_, im + q * per, bytes, cudaMemcpyHostToDevice, h->stream);
        }
        st = dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
            using R = decltype(rt);
            constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
            constexpr bool ONS = decltype(ons)::value != 0;
            auto M = make_model<R, NL>(h);
            if (to_device) k_pack<R, GEN, ONS, NL><<<grid_for(c * h->N), 256, 0, h->stream>>>(M, r0 + q, c, dre, dim_);
            else k_unpack<R, GEN, NL><<<grid_for(c * h->N), 256, 0, h->stream>>>(M, r0 + q, c, dre, dim_);
            h->launches++;
            SCMC_CK(cudaGetLastError());
            return SCMC_OK;
        });
        if (st == SCMC_OK && !to_device) {
            cudaMemcpyAsync(re + q * per, dre, bytes, cudaMemcpyDeviceToHost, h->stream);
            cudaMemcpyAsync(im + q * per, dim_, bytes, cudaMemcpyDeviceToHost, h->stream);
        }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (st == SCMC_OK && e != cudaSuccess) st = cuda_fail(e, "state_io sync", __FILE__, __LINE__);
    }
    cudaFree(dre); cudaFree(dim_);
    return st;
}
int32_t scmc_set_state(scmc_t* h, int64_t r0, int64_t count, const double* re, const double* im) {
    return state_io(h, r0, count, const_cast<double*>(re), const_cast<double*>(im), true);
}
int32_t scmc_get_state(scmc_t* h, int64_t r0, int64_t count, double* re, double* im) { return state_io(h, r0, count, re, im, false); }

int32_t scmc_randomize(scmc_t* h, uint64_t seed) {
    SCMC_ARG(h, "null handle");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        auto M = make_model<R, NL>(h);
        k_randomize<R, GEN, ONS, NL><<<grid_for(h->R * h->N), 256, 0, h->stream>>>(M, seed, h->roff);
        h->launches++;
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    }));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    h->T_stale = false;          // every replica's T cache was just rebuilt
    return SCMC_OK;
}

int32_t scmc_get_T(scmc_t* h, int64_t r, double* T, double* Tsq) {
    SCMC_ARG(h && T, "null argument");
    SCMC_ARG(r >= 0 && r < h->R, "replica out of range");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    double *dT = nullptr, *dTsq = nullptr;
    const size_t bytes = (size_t)h->N * 16 * 8;
    SCMC_CK(cudaMalloc(&dT, bytes));
    if (Tsq && cudaMalloc(&dTsq, bytes) != cudaSuccess) { cudaFree(dT); return cuda_fail(cudaErrorMemoryAllocation, "staging", __FILE__, __LINE__); }
    int32_t st = dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
        constexpr bool ONS = decltype(ons)::value != 0;
        auto M = make_model<R, NL>(h);
        k_get_T<R, GEN, ONS, NL><<<grid_for(h->N), 256, 0, h->stream>>>(M, r, dT, dTsq);
        h->launches++;
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    });
    if (st == SCMC_OK) {
        cudaMemcpyAsync(T, dT, bytes, cudaMemcpyDeviceToHost, h->stream);
        if (Tsq) cudaMemcpyAsync(Tsq, dTsq, bytes, cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) st = cuda_fail(e, "get_T sync", __FILE__, __LINE__);
    }
    cudaFree(dT); if (dTsq) cudaFree(dTsq);
    return st;
}

int32_t scmc_energy(scmc_t* h, double* E) {
    SCMC_ARG(h && E, "null argument");
    SCMC_CK(cudaSetDevice(h->device));
    double* dE = nullptr;
    SCMC_CK(scratch_get(h, 7, h->R * 8, (void**)&dE));
    int32_t st = launch_measure(h, nullptr, dE);
    if (st == SCMC_OK) {
        cudaMemcpyAsync(E, dE, h->R * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) st = cuda_fail(e, "energy sync", __FILE__, __LINE__);
    }
    return st;
}

int32_t scmc_local_field(scmc_t* h, int64_t r, double* out) {
    SCMC_ARG(h && out, "null argument");
    SCMC_ARG(r >= 0 && r < h->R, "replica out of range");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    double* dH = nullptr;
    const size_t bytes = (size_t)h->N * 16 * 8;
    SCMC_CK(cudaMalloc(&dH, bytes));
    int32_t st = dispatch(h, [&](auto rt, auto, auto, auto nl) -> int32_t {
        using R = decltype(rt);
        constexpr int NL = decltype(nl)::value;
        auto M = make_model<R, NL>(h);
        k_local_field<R, NL><<<grid_for(h->N), 256, 0, h->stream>>>(M, r, dH);
        h->launches++;
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    });
    if (st == SCMC_OK) {
        cudaMemcpyAsync(out, dH, bytes, cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) st = cuda_fail(e, "local_field sync", __FILE__, __LINE__);
    }
    cudaFree(dH);
    return st;
}

int32_t scmc_measure(scmc_t* h, double* obs) {
    SCMC_ARG(h && obs, "null argument");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(launch_measure(h, h->obs));
    SCMC_CK(cudaMemcpyAsync(obs, h->obs, h->R * SCMC_NOBS * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

int32_t scmc_sweep_async(scmc_t* h, int64_t n_sweeps, int32_t hits, uint64_t seed, uint64_t sweep_counter) {
    SCMC_ARG(h, "null handle");
    SCMC_ARG(n_sweeps >= 0 && hits >= 1, "scmc_sweep: n_sweeps >= 0 and hits >= 1 required");
    SCMC_CK(cudaSetDevice(h->device));
    return launch_sweeps(h, n_sweeps, hits, seed, sweep_counter, nullptr);
}

int32_t scmc_sweep(scmc_t* h, int64_t n_sweeps, int32_t hits, const double* beta, const double* sigma, uint64_t seed,
                   uint64_t sweep_counter, int64_t* accepted, int64_t* attempted) {
    SCMC_ARG(h && beta && sigma, "null argument");
    SCMC_ARG(n_sweeps >= 0 && hits >= 1, "scmc_sweep: n_sweeps >= 0 and hits >= 1 required");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(upload_scalars(h, beta, sigma));
    SCMC_CK(cudaMemsetAsync(h->acc, 0, h->R * sizeof(unsigned long long), h->stream));
    SCMC_TRY(launch_sweeps(h, n_sweeps, hits, seed, sweep_counter, nullptr));
    if (accepted) {
        static_assert(sizeof(unsigned long long) == sizeof(int64_t), "");
        SCMC_CK(cudaMemcpyAsync(accepted, h->acc, h->R * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    }
    SCMC_CK(cudaStreamSynchronize(h->stream));
    if (attempted) for (int64_t r = 0; r < h->R; r++) attempted[r] = n_sweeps * hits * h->N;
    return SCMC_OK;
}

int32_t scmc_sweep_deterministic(scmc_t* h, int32_t hits, const double* beta, const double* sigma, const double* xi, const double* u,
                                 uint8_t* accept, double* dE) {
    SCMC_ARG(h && beta && sigma && xi && u && accept, "null argument");
    SCMC_ARG(hits >= 1, "hits >= 1 required");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(upload_scalars(h, beta, sigma));
    const size_t n = (size_t)hits * h->R * h->N;
    double *dxi = nullptr, *du = nullptr, *ddE = nullptr;
    uint8_t* dacc = nullptr;
    int32_t st = SCMC_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && st == SCMC_OK) st = cuda_fail(e, "sweep_deterministic", __FILE__, __LINE__); };
    fail(cudaMalloc(&dxi, n * 2 * h->d * 8)); fail(cudaMalloc(&du, n * 8)); fail(cudaMalloc(&dacc, n)); fail(cudaMalloc(&ddE, n * 8));
    if (st == SCMC_OK) {
        fail(cudaMemcpyAsync(dxi, xi, n * 2 * h->d * 8, cudaMemcpyHostToDevice, h->stream));
        fail(cudaMemcpyAsync(du, u, n * 8, cudaMemcpyHostToDevice, h->stream));
        fail(cudaMemsetAsync(dacc, 0, n, h->stream));
        fail(cudaMemsetAsync(h->acc, 0, h->R * sizeof(unsigned long long), h->stream));
    }
    for (int c = 0; c < h->n_col && st == SCMC_OK; c++) {
        SweepArgs A{};
        A.colour = c; A.colour_last = c; A.hits = hits; A.r0 = 0; A.rcount = h->R; A.visit0 = 0; A.seed = 0; A.roff = h->roff; A.stream = 1;
        A.beta = h->beta; A.sigma = h->sigma; A.active = nullptr; A.acc = h->acc;
        A.inj_xi = dxi; A.inj_u = du; A.out_acc = dacc; A.out_dE = ddE;
        st = launch_colour_pass(h, A, h->stream);
    }
    if (st == SCMC_OK) {
        fail(cudaMemcpyAsync(accept, dacc, n, cudaMemcpyDeviceToHost, h->stream));
        if (dE) fail(cudaMemcpyAsync(dE, ddE, n * 8, cudaMemcpyDeviceToHost, h->stream));
        fail(cudaStreamSynchronize(h->stream));
    }
    cudaFree(dxi); cudaFree(du); cudaFree(dacc); cudaFree(ddE);
    return st;
}

int32_t scmc_update_deterministic(scmc_t* h, int64_t r, int64_t site, const double* xi, double u, double beta, double sigma, double* dE,
                                  int32_t* accepted) {
    SCMC_ARG(h && xi, "null argument");
    SCMC_ARG(r >= 0 && r < h->R && site >= 0 && site < h->N, "replica / site out of range");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    double* dbuf = nullptr;
    SCMC_CK(cudaMalloc(&dbuf, (2 * h->d + 2) * 8));
    int32_t st = SCMC_OK;
    cudaError_t e = cudaMemcpyAsync(dbuf + 2, xi, 2 * h->d * 8, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) st = cuda_fail(e, "update_deterministic", __FILE__, __LINE__);
    if (st == SCMC_OK)
        st = dispatch(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
            using R = decltype(rt);
            constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
            constexpr bool ONS = decltype(ons)::value != 0;
            auto M = make_model<R, NL>(h);
            k_single_update<R, GEN, ONS, NL><<<1, 32, 0, h->stream>>>(M, r, h->perm[site], dbuf + 2, u, beta, sigma, dbuf);
            h->launches++;
            SCMC_CK(cudaGetLastError());
            return SCMC_OK;
        });
    double res[2] = {0, 0};
    if (st == SCMC_OK) {
        e = cudaMemcpyAsync(res, dbuf, 16, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) st = cuda_fail(e, "update_deterministic", __FILE__, __LINE__);
    }
    cudaFree(dbuf);
    if (dE) *dE = res[0];
    if (accepted) *accepted = res[1] != 0.0;
    return st;
}

}  // extern "C"
