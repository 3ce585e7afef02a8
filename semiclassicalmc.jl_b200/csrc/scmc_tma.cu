// scmc_tma.cu -- TMA-pipelined streaming colour pass of the psi-gather family (sm_100a).
//
// Replaces, for the production workload, the per-thread gather of k_sweep_colour_psi (scmc_core.cu): same arithmetic in the same order
// (density sums over the neighbour slots 0..z-1, one 16x16 mat-vec with Jq, hits) => bit-identical chains; what changes is how the states
// reach the SM.  Reference work unit: localUpdate! (src/MonteCarlo.jl:263-286) inside the 2N-update loops (src/MonteCarlo.jl:40-44, 76-80).
//
//   * scmc_create cuts every colour class into tiles of <= 128 class positions (plan_tma).  The states a tile needs -- its own sites and all
//     their neighbours, each ONCE -- are a handful of runs of consecutive permuted indices on translation-invariant lattices (128x128
//     triangular: 547 states in 5.4 runs per tile instead of 7 x 128 gathered states), found from the neighbour table alone.
//   * one CTA owns ONE tile and walks over a chunk of replicas.  Per replica, one elected warp issues a cp.async.bulk (UBLKCP) per run and
//     plane into a 2-stage shared-memory ring, completion on an mbarrier (SYNCS); the stage slots of a thread's own site and neighbours, its
//     site number and its store address are loaded ONCE per CTA and stay in registers for all replicas, so a visit has no index load, no
//     address arithmetic and no DRAM round trip in its dependent chain: the next replica's states land while this one's hits run.
//   * classes whose tiles do not decompose into a few runs (the seam-repair classes of L % 3 != 0, irregular graphs) keep the direct gather.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>
#include "handle.h"

namespace scmc {

#ifndef SCMC_TMA_BLOCKS
#define SCMC_TMA_BLOCKS 5        // CTAs per SM the kernel is compiled for (128 threads each)
#endif
#ifndef SCMC_TMA_ABLATE
#define SCMC_TMA_ABLATE 0        // timing ablations (profiles/README.md, round 2); 0 = production.  1: no mat-vec, 2: one neighbour, 3: no random numbers
#endif
#ifndef SCMC_TMA_PREFETCH
#define SCMC_TMA_PREFETCH 0      // replicas ahead of a stage refill whose runs are prefetched into L2 (0: off)
#endif
#ifndef SCMC_TC_PACKED
#define SCMC_TC_PACKED 1         // tensor-core variant: density sums as packed fma.rn.f32x2 (its own summation order; the variant is not bit-identical anyway)
#endif
#ifndef SCMC_TC_PRE
#define SCMC_TC_PRE 1            // tensor-core variant: generate both hits' random numbers between the MMA issue and the wait for its result
#endif
#ifndef SCMC_TMA_GAP
#define SCMC_TMA_GAP 32
#endif
// measured on C5 (A/B harness, 8192 replicas): 8 -> 5.49e10, 32 -> 5.55e10, 96 -> 5.32e10 updates/s (fewer, longer copies against unused bytes)
constexpr int TMA_GAP = SCMC_TMA_GAP;       // runs closer than this many unused states are merged into one copy

// ------------------------------------------------------------------------------------------------ planning (host)
void plan_tma(scmc_handle* h, const std::vector<int32_t>& col_list, const std::vector<int32_t>& nb) {
    auto& P = h->tma;
    P.ready = 0;
    if (getenv("SCMC_TMA") && atoi(getenv("SCMC_TMA")) == 0) return;
    if (h->n_labels != 1 || h->z > 6) return;
    const int64_t N = h->N;
    const int z = h->z;
    struct Tile { int cnt, nruns, slots, cp0; std::vector<uint32_t> meta; std::vector<uint32_t> runs; };
    std::vector<Tile> tiles;
    int cap = 0, max_runs = 0;
    std::vector<int32_t> loc_of(N + 1, -1);
    for (int c = 0; c < h->n_col; c++) {
        auto& C = P.cls[c];
        C = {};
        const int nc = h->col_off[c + 1] - h->col_off[c];
        if (nc < 4 * TMA_TILE) continue;                              // small classes: the direct-gather kernel visits them in one launch
        const int nt = (nc + TMA_TILE - 1) / TMA_TILE, ts = (nc + nt - 1) / nt;
        std::vector<Tile> mine;
        bool ok = true;
        int cls_cap = 0, cls_runs = 0;
        for (int t = 0; t < nt && ok; t++) {
            const int cp0 = h->col_off[c] + t * ts, cnt = std::min(ts, h->col_off[c + 1] - cp0);
            std::vector<int32_t> need;
            need.reserve((size_t)cnt * (z + 1));
            for (int i = 0; i < cnt; i++) {
                const int32_t p = col_list[cp0 + i];
                need.push_back(p);
                for (int k = 0; k < z; k++) {
                    const int32_t j = nb[(size_t)k * N + p];
                    if (j < N) need.push_back(j);
                }
            }
            std::sort(need.begin(), need.end());
            need.erase(std::unique(need.begin(), need.end()), need.end());
            Tile T;
            T.cnt = cnt; T.cp0 = cp0; T.nruns = 0; T.slots = 0;
            size_t i = 0;
            while (i < need.size()) {
                size_t j = i;
                while (j + 1 < need.size() && need[j + 1] - need[j] <= TMA_GAP + 1) j++;
                const int32_t start = need[i], len = need[j] - need[i] + 1;
                if (len > 65535) { ok = false; break; }
                T.runs.push_back((uint32_t)start);
                T.runs.push_back((uint32_t)T.slots | ((uint32_t)len << 16));
                for (size_t q = i; q <= j; q++) loc_of[need[q]] = T.slots + (need[q] - start);
                T.slots += len; T.nruns++;
                i = j + 1;
            }
            if (!ok || T.nruns > TMA_MAX_RUNS || T.slots + 1 > 4000) { ok = false; break; }
            // average run must be long enough for the bulk copies to pay (irregular graphs: keep the gather)
            if (T.slots < 16 * T.nruns) { ok = false; break; }
            T.meta.assign((size_t)TMA_TILE * 8, 0);
            for (int i2 = 0; i2 < TMA_TILE; i2++) {
                uint32_t* m = &T.meta[(size_t)i2 * 8];
                uint32_t loc[8] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu, 0};      // 0xFFFF = the all-zero slot
                if (i2 < cnt) {
                    const int32_t p = col_list[cp0 + i2];
                    loc[0] = (uint32_t)loc_of[p];
                    for (int k = 0; k < z; k++) {
                        const int32_t j = nb[(size_t)k * N + p];
                        if (j < N) loc[1 + k] = (uint32_t)loc_of[j];
                    }
                    loc[7] = 1;
                    m[4] = (uint32_t)h->orig[p]; m[5] = (uint32_t)p;
                }
                for (int q = 0; q < 4; q++) m[q] = loc[2 * q] | (loc[2 * q + 1] << 16);
            }
            cls_cap = std::max(cls_cap, T.slots + 1);
            cls_runs = std::max(cls_runs, T.nruns);
            mine.push_back(std::move(T));
        }
        if (!ok) continue;
        C.ok = 1; C.n_tiles = nt; C.tile = ts; C.first = (int)tiles.size();
        cap = std::max(cap, cls_cap); max_runs = std::max(max_runs, cls_runs);
        for (auto& T : mine) tiles.push_back(std::move(T));
    }
    if (tiles.empty()) return;
    P.cap = (cap + 7) / 8 * 8;
    P.max_runs = max_runs;
    const size_t nt = tiles.size();
    std::vector<uint32_t> meta(nt * TMA_TILE * 8), runs(nt * TMA_MAX_RUNS * 2, 0);
    std::vector<int32_t> info(nt * 4);
    for (size_t t = 0; t < nt; t++) {
        memcpy(&meta[t * TMA_TILE * 8], tiles[t].meta.data(), sizeof(uint32_t) * TMA_TILE * 8);
        memcpy(&runs[t * TMA_MAX_RUNS * 2], tiles[t].runs.data(), sizeof(uint32_t) * tiles[t].runs.size());
        info[t * 4 + 0] = tiles[t].cnt; info[t * 4 + 1] = tiles[t].nruns; info[t * 4 + 2] = tiles[t].slots; info[t * 4 + 3] = tiles[t].cp0;
    }
    bool okd = cudaMalloc(&P.d_meta, meta.size() * 4) == cudaSuccess && cudaMalloc(&P.d_runs, runs.size() * 4) == cudaSuccess &&
               cudaMalloc((void**)&P.d_tile, info.size() * 4) == cudaSuccess;
    okd = okd && cudaMemcpy(P.d_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
          cudaMemcpy(P.d_runs, runs.data(), runs.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
          cudaMemcpy(P.d_tile, info.data(), info.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!okd) {
        cudaGetLastError();
        if (P.d_meta) cudaFree(P.d_meta);
        if (P.d_runs) cudaFree(P.d_runs);
        if (P.d_tile) cudaFree(P.d_tile);
        P.d_meta = P.d_runs = nullptr; P.d_tile = nullptr;
        return;
    }
    h->bytes += (int64_t)(meta.size() + runs.size() + info.size()) * 4;
    P.ready = 1;
    if (getenv("SCMC_TMA_DEBUG")) {
        size_t slots = 0, nr = 0;
        for (auto& T : tiles) { slots += T.slots; nr += T.nruns; }
        fprintf(stderr, "scmc tile plan: %zu tiles, stage capacity %d slots, %.1f staged states and %.2f runs per tile on average, gap %d\n", nt, P.cap, (double)slots / nt, (double)nr / nt, TMA_GAP);
    }
}

// ------------------------------------------------------------------------------------------------ mbarrier / bulk-copy primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------------ tcgen05 primitives (TC variant)
// The 16x16 mat-vec of a visit for the 128 sites of a tile is ONE tensor-core tile D[128x16] = A[128x16] B[16x16]^T (kind::tf32, 3xTF32 split,
// FP32 accumulation in tensor memory).  A comes from TMEM (tcgen05.st: a thread's TMEM lane is its site), B = Jq from shared memory (K-major,
// no swizzle, 8-row x 16-byte core matrices), D is read back with tcgen05.ld.  Every warp issues its own MMAs over the CTA's A region into
// its own D columns and reads back only its 32 lanes, so the warps of a CTA never wait for each other (rows of other warps that a warp's
// MMA reads may be half-written: they land in D lanes nobody reads).  A shared D region with the other warps' lanes masked out
// (disable-output-lane mask) was measured slower: MMAs into one D region are ordered behind each other.  Checked stand-alone in profiles/microbench/tc_matvec.cu.
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
                   "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    return leader != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: start address, leading (K chunk) and stride (8-row group) byte offsets, sm_100 descriptor version, no swizzle
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
constexpr uint32_t TC_LBO = 256, TC_SBO = 128;                       // Jq in core-matrix order: chunk (b / 4) * 256 + group (a / 8) * 128 + (a % 8) * 16 + (b % 4) * 4
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);      // D = F32, A = B = TF32, K-major, N = 16, M = 128
// Tensor memory of a CTA: A_hi 16 + A_lo 16 + 4 warps x D 16 = 96 columns, taken as THREE allocations of 32 columns (allocations are powers of
// two: one of 128 would cap the SM at 4 CTAs; equal-sized pieces cannot fragment the 512 columns, 5 CTAs x 3 = 15 of 16 pieces).
#ifndef SCMC_TC_PIPE
#define SCMC_TC_PIPE 0           // 1: the MMAs of visit it + 1 run during the hits of visit it (two A / D buffers: three allocations of 32 columns);
#endif                           // 0: MMAs, random numbers and hits of a visit in sequence (one allocation of 64 columns).  Measured on C5 (profiles/README.md):
                                 // pipelined 5.45e10 burst / 5.16e10 sustained against 5.57e10 / 5.24e10 -- the wait on the commit barrier moves, it does not go away
// ONE set of MMAs per tile visit, issued by one warp once all four have stored their rows.  (Every warp its own MMAs into its own D columns -- no
// rendezvous inside the CTA, four times the tensor work -- was 5.5 % slower under the power cap: profiles/README.md.)
constexpr int TC_TMEM_COLS = SCMC_TC_PIPE ? 32 : 64, TC_TMEM_ALLOCS = SCMC_TC_PIPE ? 3 : 1;
#ifndef SCMC_TC_BLOCKS
#define SCMC_TC_BLOCKS 5         // CTAs per SM the tensor-core variant is compiled for
#endif
#ifndef SCMC_TC_REFILL_LATE
#define SCMC_TC_REFILL_LATE 1    // tensor-core variant: warp 0 waits for the stage's release and issues the refill after its tcgen05.st, not before
#endif
#ifndef SCMC_TC_ISSUER
#define SCMC_TC_ISSUER 1         // warp that issues the MMAs of a tile visit (warp 0 carries the stage refills)
#endif
#ifndef SCMC_TC_LAST
#define SCMC_TC_LAST 0           // 1: the warp that stores its rows LAST issues the MMAs of the tile visit (ticket from a shared-memory atomic: nobody waits on an
#endif                           //    "A stored" barrier and no fixed warp has its own work serialised behind the rendezvous); 0: warp 1 waits for all four and issues.
                                 //    Measured on C5 (same box): 6.00-6.03e10 burst / 5.71e10 sustained against 6.15e10 / 5.85e10 for 0 -- the atomic's round trip sits in every warp's path

// Packed density sums of the tensor-core variant (d = 4; basis order of gen::density_acc_g1 / _g3): with the pairs c_m = (re_m, im_m) as they come out
// of the 128-bit loads, c_m * c_m = (re_m^2, im_m^2), c_m * c_n = (re_m re_n, im_m im_n) and c_m * swap(c_n) = (re_m im_n, im_m re_n) are one
// fma.rn.f32x2 each: 16 FFMA2 + 6 moves per neighbour instead of 32 FFMA; the two halves are combined once per visit (sum linear in the neighbours).
__device__ __forceinline__ unsigned long long pk2(float x, float y) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ float2 up2(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
struct PackedDensity {
    unsigned long long dg[4], a[6], b[6];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 4; i++) dg[i] = 0ull;
#pragma unroll
        for (int i = 0; i < 6; i++) { a[i] = 0ull; b[i] = 0ull; }
    }
    __device__ __forceinline__ void add(const float4 q0, const float4 q1) {
        const unsigned long long c[4] = {pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), pk2(q1.z, q1.w)};
        const unsigned long long sw[4] = {0ull, pk2(q0.w, q0.z), pk2(q1.y, q1.x), pk2(q1.w, q1.z)};
#pragma unroll
        for (int m = 0; m < 4; m++) fma2(dg[m], c[m], c[m]);
        int i = 0;
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
            for (int n = m + 1; n < 4; n++, i++) { fma2(a[i], c[m], c[n]); fma2(b[i], c[m], sw[n]); }
    }
    __device__ __forceinline__ void finish(float* P) const {
        constexpr int DIAG[4] = {0, 7, 12, 15}, OFF[6] = {1, 3, 5, 8, 10, 13};      // P index of |psi_m|^2 and of Re(conj(psi_m) psi_n); Im follows Re
#pragma unroll
        for (int m = 0; m < 4; m++) { const float2 v = up2(dg[m]); P[DIAG[m]] = v.x + v.y; }
#pragma unroll
        for (int i = 0; i < 6; i++) { const float2 u = up2(a[i]), w = up2(b[i]); P[OFF[i]] = u.x + u.y; P[OFF[i] + 1] = w.x - w.y; }
    }
};

// L2 prefetch of a run (TMA engine, no shared-memory destination), SCMC_TMA_PREFETCH replicas ahead of the copy that will consume it.
// Measured on C5 (profiles/README.md, round 2): 5.50e10 -> 5.05e10 updates/s with a distance of 2 or 4 -- off by default.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct TmaArgs {
    const uint4* meta;       // [tiles][128][2]
    const uint2* runs;       // [tiles][TMA_MAX_RUNS]
    const int32_t* tile;     // [tiles][4]
    int32_t first, cap, rc;  // first tile of the class, slots per plane and stage, replicas per CTA
    PhiloxKeys7 rk;          // round keys of random stream v2 for A.seed
};

// ------------------------------------------------------------------------------------------------ kernel
// CAPC > 0: the stage has CAPC slots per plane at compile time (the tile plan's cap must not exceed it), so plane and stage strides are immediates of the
// shared-memory loads instead of two integer instructions per load; CAPC = 0: strides from T.cap at run time.
template <class R, int GEN, int Z, int STREAM, int HITS, bool TC, int CAPC = 0>
__global__ void __launch_bounds__(TMA_TILE, TC ? SCMC_TC_BLOCKS : SCMC_TMA_BLOCKS) k_sweep_tma_psi(const __grid_constant__ DevModel<R, 1> M, const SweepArgs A, const TmaArgs T) {
    constexpr int D = GenOps<GEN>::D, NPV = D / 2, NP = GenOps<GEN>::NDENS;
    static_assert(NP == NG, "fused density basis (16 entries) only");
    using V = vec4<R>;
    constexpr uint32_t VB = sizeof(V);
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t cap = CAPC ? (uint32_t)CAPC : (uint32_t)T.cap;
    const uint32_t plane_bytes = cap * VB, stage_bytes = NPV * plane_bytes;
    R* sJt = reinterpret_cast<R*>(smem + TMA_STAGES * stage_bytes);                      // TC: Jq_hi, Jq_lo in core-matrix order (2 x 1 KB)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sJt + (TC ? 2 : 1) * NG * NG);      // full[0..S), empty[0..S), TC: one per warp
    int32_t* sRep = reinterpret_cast<int32_t*>(bars + 2 * TMA_STAGES + TMA_TILE / 32);    // replicas of this CTA that are alive
    R* sBeta = reinterpret_cast<R*>(sRep + TMA_RC);
    R* sSigma = sBeta + TMA_RC;
    __shared__ int s_nit;
    __shared__ uint32_t s_tmem[4];
    __shared__ uint32_t s_ticket;                                  // SCMC_TC_LAST: warps that have stored their rows, over all visits of this CTA
    static_assert(!TC || (std::is_same<R, float>::value && D == 4), "tensor-core mat-vec: FP32, d = 4");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = T.first + blockIdx.x;
    const int64_t rb = A.r0 + (int64_t)blockIdx.y * T.rc;
    const int rcnt = (int)min((int64_t)T.rc, A.r0 + A.rcount - rb);

    // per-thread constants of the tile: stage slots (as byte offsets) of the own site and the neighbours, site number, permuted index
    const uint4 m0 = T.meta[((size_t)tile * TMA_TILE + tid) * 2], m1 = T.meta[((size_t)tile * TMA_TILE + tid) * 2 + 1];
    const uint32_t zero_off = (cap - 1) * VB;
    uint32_t off[7];
    {
        const uint32_t w[4] = {m0.x, m0.y, m0.z, m0.w};
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint32_t l = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
            off[k] = l == 0xFFFFu ? zero_off : l * VB;
        }
    }
    const bool valid = (m0.w >> 16) != 0;
    const uint32_t site = m1.x, p = m1.y;

    if constexpr (TC) {
        if (warp == 0) {                                           // tensor memory of this CTA (freed by the same warp at the end)
#pragma unroll
            for (int q = 0; q < TC_TMEM_ALLOCS; q++)
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem[q])), "n"(TC_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        for (int t = tid; t < NG * NG; t += TMA_TILE) {            // Jqt[b][a] = Jq[a][b] -> B[n = a][k = b], split into TF32 head and tail
            const int b = t / NG, a = t % NG;
            const float j = (float)((const R*)A.Jqt)[t];
            const float hi = __uint_as_float(__float_as_uint(j) & 0xFFFFE000u);
            const uint32_t o = ((b / 4) * TC_LBO + (a / 8) * TC_SBO + (a % 8) * 16 + (b % 4) * 4) / 4;
            reinterpret_cast<float*>(sJt)[o] = hi;
            reinterpret_cast<float*>(sJt)[NG * NG + o] = j - hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // written by the generic proxy, read by the tensor core
    } else {
        for (int t = tid; t < NG * NG; t += TMA_TILE) sJt[t] = ((const R*)A.Jqt)[t];
    }
    if (tid < NPV * TMA_STAGES) {                                  // the all-zero slot of every plane of every stage (never a copy target)
        V zv; zv.x = zv.y = zv.z = zv.w = R(0);
        *reinterpret_cast<V*>(smem + tid * plane_bytes + zero_off) = zv;
    }
    if (warp == 0) {                                               // alive replicas of this chunk, in order (32 at a time)
        int n_alive = 0;
        for (int c0 = 0; c0 < rcnt; c0 += 32) {
            const int j = c0 + lane;
            const bool al = j < rcnt && (A.active == nullptr || A.active[rb + j]);
            const unsigned mask = __ballot_sync(0xffffffffu, al);
            if (al) {
                const int at = n_alive + __popc(mask & ((1u << lane) - 1));
                sRep[at] = j;
                sBeta[at] = ((const R*)A.beta)[rb + j];
                sSigma[at] = ((const R*)A.sigma)[rb + j];
            }
            n_alive += __popc(mask);
        }
        if (lane == 0) {
            s_nit = n_alive;
            s_ticket = 0;
#pragma unroll
            for (int s = 0; s < TMA_STAGES; s++) { mbar_init(smem_u32(bars + s), 1); mbar_init(smem_u32(bars + TMA_STAGES + s), TMA_TILE / 32); }
            if (TC)
                for (int w = 0; w < 4; w++) mbar_init(smem_u32(bars + 2 * TMA_STAGES + w), w >= 2 ? TMA_TILE / 32 : 1);      // D ready 0 / 1, A stored 0 / 1
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    if (TC) tc_fence_before();
    __syncthreads();
    if (TC) tc_fence_after();
    // TC: tensor-memory columns of this CTA.  In-visit MMAs (SCMC_TC_PIPE = 0): one allocation of 64 columns = A_hi 16 | A_lo 16 | D 16.  Pipelined by one
    // visit (SCMC_TC_PIPE = 1): three allocations of 32 columns = A buffer 0 (hi | lo), A buffer 1, D buffers 0 | 1.  Barriers: "D ready" 0 / 1, "A stored
    // by all four warps" 0 / 1 (buffer = visit & 1; each completes once per two visits, so nobody can run a whole phase ahead of a waiter).
    const uint32_t tm_lane = (uint32_t)(warp * 32) << 16;
    auto tm_a = [&](int buf) -> uint32_t { return SCMC_TC_PIPE ? s_tmem[buf] : s_tmem[0]; };
    auto tm_d = [&](int buf) -> uint32_t { return SCMC_TC_PIPE ? s_tmem[2] + 16 * buf : s_tmem[0] + 32; };
    const uint32_t tc_bar0 = smem_u32(bars + 2 * TMA_STAGES);      // + 8 * buf: D ready; + 16 + 8 * buf: A stored
    const uint64_t tc_bhi = tc_smem_desc(smem_u32(sJt), TC_LBO, TC_SBO), tc_blo = tc_smem_desc(smem_u32(sJt + NG * NG), TC_LBO, TC_SBO);
    const int n_it = s_nit;
    const size_t ps = (size_t)M.nrep * M.Ns;

    // copy descriptor of this lane (warp 0 issues): run = lane / NPV, plane = lane % NPV
    const int nruns = T.tile[tile * 4 + 1];
    const uint32_t tile_bytes = (uint32_t)T.tile[tile * 4 + 2] * VB * NPV;
    const char* c_src = nullptr;
    uint32_t c_dst = 0, c_bytes = 0;
    if (warp == 0 && lane < nruns * NPV) {
        const uint2 rd = T.runs[(size_t)tile * TMA_MAX_RUNS + lane / NPV];
        const int v = lane % NPV;
        c_src = reinterpret_cast<const char*>(M.psi + (size_t)v * ps + rb * M.Ns + rd.x);
        c_dst = (uint32_t)v * plane_bytes + (rd.y & 0xFFFFu) * VB;
        c_bytes = (rd.y >> 16) * VB;
    }
    const uint32_t smem0 = smem_u32(smem), bar0 = smem_u32(bars);
    const uint32_t rep_bytes = (uint32_t)M.Ns * VB;
    auto issue = [&](int it, int s) {
        if (lane == 0) mbar_expect_tx(bar0 + 8 * s, tile_bytes);
        __syncwarp();
        if (c_bytes) {
            bulk_g2s(smem0 + s * stage_bytes + c_dst, c_src + (uint32_t)sRep[it] * rep_bytes, c_bytes, bar0 + 8 * s);
            if (SCMC_TMA_PREFETCH && it + SCMC_TMA_PREFETCH < n_it) bulk_prefetch_l2(c_src + (uint32_t)sRep[it + SCMC_TMA_PREFETCH] * rep_bytes, c_bytes);
        }
    };
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; s++)
            if (s < n_it) issue(s, s);
    }

    // HITS = 2: the warp total of accepted hits of replica `it` comes from two ballots (bit 0, bit 1 of a thread's count 0..2) and stays in lane
    // it % 32; one atomic instruction per warp and 32 replicas instead of a reduce -> branch -> atomic chain per visit
    unsigned tot_mine = 0;
    char* const own = reinterpret_cast<char*>(M.psi + rb * M.Ns + p);      // own state of the chunk's first replica (plane 0)
    constexpr bool PRE = TC && HITS == 2 && STREAM == 2 && SCMC_TC_PRE;    // both hits' random numbers are generated while the tensor core works
    constexpr bool PIPE = TC && SCMC_TC_PIPE;                              // the MMAs of visit it + 1 are in flight during the hits of visit it

    // ---- warp 0: once all four warps have released the stage of replica `it`, refill it with the states of replica it + STAGES.  The copies have two
    //      visits of slack, the wait for the slowest warp does not: in the tensor-core variant it runs AFTER this warp's rows are in tensor memory
    //      (SCMC_TC_REFILL_LATE), otherwise the other three warps' MMA waits for warp 0, which waits for them
    auto refill = [&](int it) {
        if (SCMC_TMA_ABLATE != 4 && warp == 0 && it + TMA_STAGES < n_it) {
            mbar_wait(bar0 + 8 * (TMA_STAGES + it % TMA_STAGES), (uint32_t)(it / TMA_STAGES) & 1u);
            issue(it + TMA_STAGES, it % TMA_STAGES);
        }
    };
    // ---- front half of a visit: wait for the stage, own state -> re / im, density sums of the neighbours -> P, release the stage, (warp 0) refill it
    auto gather = [&](int it, R* re, R* im, R* P) {
        const int s = it % TMA_STAGES;
        const uint32_t parity = (it / TMA_STAGES) & 1;
#if SCMC_TMA_ABLATE == 4                                   // timing only: the first two replicas' states are consumed over and over, no refill, no wait
        if (it < TMA_STAGES)
#endif
        mbar_wait(bar0 + 8 * s, parity);
        const unsigned char* st = smem + s * stage_bytes;
        if constexpr (TC && SCMC_TC_PACKED && CAPC > 0) {
            // compile-time strides: one add per slot (32-bit shared address of the slot in stage 0 + stage offset), the plane offset is an immediate of the load
            const uint32_t so = smem0 + (uint32_t)s * (uint32_t)(NPV * CAPC * VB);
            auto lds2 = [&](uint32_t a, float4& q0, float4& q1) {
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(a));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(a), "n"(CAPC * VB));
            };
            float4 q0, q1;
            lds2(so + off[0], q0, q1);
            re[0] = q0.x; im[0] = q0.y; re[1] = q0.z; im[1] = q0.w; re[2] = q1.x; im[2] = q1.y; re[3] = q1.z; im[3] = q1.w;
            PackedDensity pd;
            pd.clear();
#pragma unroll
            for (int k = 0; k < Z; k++) {
                lds2(so + off[1 + k], q0, q1);
                pd.add(q0, q1);
            }
            pd.finish(P);
        } else if constexpr (TC && SCMC_TC_PACKED) {
#pragma unroll
            for (int v = 0; v < NPV; v++) {
                const V q = *reinterpret_cast<const V*>(st + v * plane_bytes + off[0]);
                re[2 * v] = q.x; im[2 * v] = q.y; re[2 * v + 1] = q.z; im[2 * v + 1] = q.w;
            }
            PackedDensity pd;
            pd.clear();
#pragma unroll
            for (int k = 0; k < Z; k++)
                pd.add(*reinterpret_cast<const float4*>(st + off[1 + k]), *reinterpret_cast<const float4*>(st + plane_bytes + off[1 + k]));
            pd.finish(P);
        } else {
#pragma unroll
            for (int v = 0; v < NPV; v++) {
                const V q = *reinterpret_cast<const V*>(st + v * plane_bytes + off[0]);
                re[2 * v] = q.x; im[2 * v] = q.y; re[2 * v + 1] = q.z; im[2 * v + 1] = q.w;
            }
#pragma unroll
            for (int t = 0; t < NP; t++) P[t] = R(0);
#pragma unroll
            for (int k = 0; k < (SCMC_TMA_ABLATE == 2 ? 1 : Z); k++) {
                R nr[D], ni[D];
#pragma unroll
                for (int v = 0; v < NPV; v++) {
                    const V q = *reinterpret_cast<const V*>(st + v * plane_bytes + off[1 + k]);
                    nr[2 * v] = q.x; ni[2 * v] = q.y; nr[2 * v + 1] = q.z; ni[2 * v + 1] = q.w;
                }
                GenOps<GEN>::density_acc(nr, ni, P);
            }
        }
        // this warp is done with the stage: release it; warp 0 refills it with the states of replica it + STAGES once all warps have
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (TMA_STAGES + s));
        if constexpr (!(TC && SCMC_TC_REFILL_LATE)) refill(it);
    };
    // ---- tensor-core mat-vec, part 1: this thread's 16 sums, split into TF32 head and tail, to its lane of A buffer `buf`; arrive on "A stored"
    auto tc_store = [&](const R* P, int buf) -> bool {
        uint32_t hi[NG], lo[NG];
#pragma unroll
        for (int t = 0; t < NG; t++) {
            hi[t] = __float_as_uint(P[t]) & 0xFFFFE000u;
            lo[t] = __float_as_uint(P[t] - __uint_as_float(hi[t]));
        }
        tmem_st16(tm_lane | tm_a(buf), hi);
        tmem_st16(tm_lane | (tm_a(buf) + 16), lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if constexpr (SCMC_TC_LAST && !SCMC_TC_PIPE) {
            // every warp takes one ticket per visit (a warp cannot reach the next visit's store before this visit's result has been read, which needs all
            // four stores): the warp holding ticket 3 (mod 4) arrived last; acq_rel orders the other warps' tensor-memory stores before its MMAs
            uint32_t t = 0;
            if (lane == 0) asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(smem_u32(&s_ticket)) : "memory");
            t = __shfl_sync(0xffffffffu, t, 0);
            return (t & 3u) == 3u;
        } else {
            if (lane == 0) mbar_arrive(tc_bar0 + 16 + 8 * buf);
            return false;
        }
    };
    // ---- part 2 (warp 1; warp 0 carries the stage refills): once all four warps have stored their rows of visit `it`, one elected thread issues
    //      P_lo Jq_hi + P_hi Jq_lo + P_hi Jq_hi into D buffer `buf` and commits to "D ready"
    auto tc_issue = [&](int it, int buf) {
        if constexpr (!(SCMC_TC_LAST && !SCMC_TC_PIPE)) mbar_wait(tc_bar0 + 16 + 8 * buf, (uint32_t)(PIPE ? it >> 1 : it) & 1u);
        if (elect_one()) {                                      // elect.sync: ptxas then emits the MMAs straight, without a loop over the active lanes
            tc_fence_after();
            constexpr uint32_t KS = (2 * TC_LBO) >> 4;          // descriptor step of one K = 8 slice
            const uint32_t ahi = tm_a(buf), alo = ahi + 16, d = tm_d(buf);
            tc_mma_tf32_ts(d, alo, tc_bhi, TC_IDESC, 0);
            tc_mma_tf32_ts(d, alo + 8, tc_bhi + KS, TC_IDESC, 1);
            tc_mma_tf32_ts(d, ahi, tc_blo, TC_IDESC, 1);
            tc_mma_tf32_ts(d, ahi + 8, tc_blo + KS, TC_IDESC, 1);
            tc_mma_tf32_ts(d, ahi, tc_bhi, TC_IDESC, 1);
            tc_mma_tf32_ts(d, ahi + 8, tc_bhi + KS, TC_IDESC, 1);
            tc_commit(tc_bar0 + 8 * buf);
        }
        __syncwarp();
    };
    // ---- back half of a visit: (random numbers,) energy coefficients, the hits, store, acceptance counts
    auto finish_visit = [&](int it, R* re, R* im, const R* P, int next_to_issue) {
        const int rl = sRep[it];
        const int64_t r = rb + rl;
        const R sigma = sSigma[it], nbeta = neg_beta_scaled(sBeta[it]);
        R Hq[NG];
        unsigned mine = 0;
        bool acc_hit[2] = {false, false};
        uint32_t uw[4];                                            // stream v2: acceptance-uniform words of the current group of four visits
        const int n_hits = HITS ? HITS : A.hits;                   // HITS = 2: the two hits of the reference's 2N-update sweep, unrolled
        R pxr[PRE ? 2 : 1][D], pxi[PRE ? 2 : 1][D];
        if constexpr (TC) {
            const int buf = PIPE ? (it & 1) : 0;
            uint32_t tc_dep = 0;
            if constexpr (PRE) {
#pragma unroll
                for (int hit = 0; hit < 2; hit++) stream_v2_normals_d4<R>(T.rk, site, (uint32_t)(A.roff + r), A.visit0 + hit, pxr[hit], pxi[hit]);
                stream_v2_uniform_words(T.rk, site, (uint32_t)(A.roff + r), A.visit0 >> 2, uw);
                // ptxas would hoist the (possibly suspending) try_wait above this arithmetic: make the wait's parity operand depend on the
                // last numbers computed, through a mask that is zero at run time but not at compile time (tile indices are non-negative)
                tc_dep = uw[3];
#pragma unroll
                for (int hit = 0; hit < 2; hit++)
#pragma unroll
                    for (int k = 0; k < D; k++) tc_dep ^= __float_as_uint(pxr[hit][k]);
                tc_dep &= (uint32_t)T.first >> 31;
            }
            // pipelined: the MMAs of the NEXT visit are issued here, after this warp's random numbers (the other warps have stored their rows by then)
            if (PIPE && warp == 1 && next_to_issue >= 0) tc_issue(next_to_issue, next_to_issue & 1);
            mbar_wait(tc_bar0 + 8 * buf, ((uint32_t)(PIPE ? it >> 1 : it) & 1u) ^ tc_dep);
            tc_fence_after();
            tmem_ld16(tm_lane | tm_d(buf), Hq);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#if SCMC_TMA_ABLATE == 1
#pragma unroll
            for (int t = 0; t < NG; t++) Hq[t] = P[t] + sJt[t];
#else
            field_from_sums_smem<1>(sJt, (const R(*)[NG])P, Hq);       // Hq = Jq . Psum
#endif
        }
        R eloc = GenOps<GEN>::energy_from_state(re, im, Hq);
#pragma unroll
        for (int hit = 0; hit < n_hits; hit++) {
            R xr[D], xi[D], u, dE;
#if SCMC_TMA_ABLATE == 3
#pragma unroll
            for (int k = 0; k < D; k++) { xr[k] = Hq[k] * R(0.01) + R(hit); xi[k] = Hq[k + D] * R(0.01); }
            u = R(0.5);
            if (false)
#endif
            if constexpr (PRE) {
#pragma unroll
                for (int k = 0; k < D; k++) { xr[k] = pxr[hit][k]; xi[k] = pxi[hit][k]; }
                // visit0 = 2 x sweep is even: hit 0 takes word 0 or 2, hit 1 word 1 or 3 (one select on a uniform predicate instead of pick_word's branches)
                u = unit_uniform(((uint32_t)A.visit0 & 2u) ? uw[2 + hit] : uw[hit], R(0));
            } else if constexpr (STREAM == 2) {
                const uint64_t visit = A.visit0 + hit;
                stream_v2_normals<R, D>(A.seed, site, (uint32_t)(A.roff + r), visit, xr, xi);
                // HITS = 2: visit0 = 2 x sweep is even, so the second hit never opens a new group of four visits
                if (hit == 0 || (HITS != 2 && ((uint32_t)visit & 3u) == 0)) stream_v2_uniform_words(A.seed, site, (uint32_t)(A.roff + r), visit >> 2, uw);
                u = unit_uniform(pick_word(uw, (uint32_t)visit & 3u), R(0));
            } else {
                stream_v1<R, D>(A.seed, site, (uint32_t)(A.roff + r), A.visit0 + hit, xr, xi, u);
            }
            const bool ok = metropolis_hit_psi<R, GEN>(re, im, Hq, xr, xi, u, nbeta, sigma, eloc, dE) && valid;
            if (HITS == 2) acc_hit[hit & 1] = ok;
            mine += ok ? 1u : 0u;
        }
        // rl < TMA_RC and Ns * 16 B * TMA_RC < 4 GB (checked by tma_pass_available): the replica offset inside the chunk is a 32-bit product
        if (HITS == 2 ? (acc_hit[0] || acc_hit[1]) : mine != 0) store_psi<R, D>(reinterpret_cast<V*>(own + (uint32_t)rl * rep_bytes), ps, re, im);
        if constexpr (HITS == 2) {
            const unsigned b0 = __ballot_sync(0xffffffffu, acc_hit[0]), b1 = __ballot_sync(0xffffffffu, acc_hit[1]);
            if (lane == (it & 31)) tot_mine = __popc(b0) + __popc(b1);
            if ((it & 31) == 31 || it == n_it - 1) {
                const int j = (it & ~31) + lane;
                if (j <= it && tot_mine) atomicAdd(A.acc + rb + sRep[j], (unsigned long long)tot_mine);
            }
        } else {
            const unsigned tot = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0 && tot) atomicAdd(A.acc + r, (unsigned long long)tot);
        }
    };

    if constexpr (PIPE) {
        // software pipeline, one visit deep: gather + tcgen05.st of visit it + 1, then the hits of visit it; the MMAs of a visit have a whole visit
        // of slack before their result is needed, so a warp never waits for the slowest warp of its CTA plus the MMA latency
        R re[D], im[D];
        if (n_it > 0) {
            R P[NP];
            gather(0, re, im, P);
            tc_store(P, 0);
            if (warp == 1) tc_issue(0, 0);
            if constexpr (SCMC_TC_REFILL_LATE) refill(0);
        }
        for (int it = 0; it < n_it; it++) {
            R nre[D], nim[D];
            const bool more = it + 1 < n_it;
            if (more) {
                R P[NP];
                gather(it + 1, nre, nim, P);
                tc_store(P, (it + 1) & 1);
                if constexpr (SCMC_TC_REFILL_LATE) refill(it + 1);
            }
            finish_visit(it, re, im, nullptr, more ? it + 1 : -1);
#pragma unroll
            for (int k = 0; k < D; k++) { re[k] = nre[k]; im[k] = nim[k]; }
        }
    } else {
        for (int it = 0; it < n_it; it++) {
            R re[D], im[D], P[NP];
            gather(it, re, im, P);
            if constexpr (TC) {
                const bool last = tc_store(P, 0);
                if (SCMC_TC_LAST ? last : warp == (SCMC_TC_ISSUER < 0 ? 1 + it % 3 : SCMC_TC_ISSUER)) tc_issue(it, 0);
                if constexpr (SCMC_TC_REFILL_LATE) refill(it);
            }
            finish_visit(it, re, im, P, -1);
        }
    }
    if constexpr (TC) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < TC_TMEM_ALLOCS; q++) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem[q]), "n"(TC_TMEM_COLS) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <class R>
static DevModel<R, 1> make_model_t(const scmc_handle* h) {
    DevModel<R, 1> M;
    memset(&M, 0, sizeof M);
    M.psi = (vec4<R>*)h->psi; M.T = (vec4<R>*)h->T; M.Tsq = (vec4<R>*)h->Tsq;
    M.nbr = h->nbr; M.lab = h->lab; M.orig = h->d_orig; M.col_list = h->d_col_list; M.cls_site = h->d_cls_site; M.cls_nbr = h->d_cls_nbr;
    M.N = (int32_t)h->N; M.Ns = (int32_t)h->Ns; M.z = h->z; M.n_col = h->n_col; M.nrep = h->R;
    for (int c = 0; c <= MAXC; c++) M.col_off[c] = h->col_off[c];
    for (int t = 0; t < 256; t++) M.cp.J[0][t] = (R)h->J[0][t];
    for (int a = 0; a < NG; a++) M.cp.K[a] = (R)h->K[a];
    return M;
}

#ifndef SCMC_TMA_CAPC
#define SCMC_TMA_CAPC 640        // compile-time stage capacity of the production instantiation (C5: 624 slots; 5 CTAs of 45.8 KB per SM); 0: always the run-time form
#endif
static size_t tma_smem_bytes(const scmc_handle* h, bool tc = true, int capc = 0) {
    const size_t VB = 4 * h->esz();
    return (size_t)TMA_STAGES * (h->d / 2) * (capc ? capc : h->tma.cap) * VB + (tc ? 2 : 1) * 256 * h->esz() + (2 * TMA_STAGES + TMA_TILE / 32) * 8 + TMA_RC * (4 + 2 * h->esz()) + 16;
}

// Tensor-core mat-vec variant (mode 6; auto unless SCMC_TC=0): rounding of the 16x16 mat-vec differs from the FFMA order of the other psi-gather
// kernels (3xTF32 products, FP32 accumulation in the tensor core), so its chains agree with theirs to rounding, not bit for bit; mode 5 keeps
// the bit-identical FFMA2 form.
bool tma_use_tc(const scmc_handle* h) {
    static const int env = getenv("SCMC_TC") ? atoi(getenv("SCMC_TC")) : -1;
    if (h->sweep_mode == 6) return true;
    if (h->sweep_mode != 0) return false;      // 5: the FFMA2 form; 1 - 4 never reach the TMA pass
    return env != 0;
}

// The pipelined kernel serves FP32, d = 4 (n = 1, 3), one coupling label, no on-site term, z = 3 or 6 neighbour slots, single-class passes.
bool tma_pass_available(const scmc_handle* h, const SweepArgs& A) {
    if (!h->tma.ready || h->prec != 32 || !h->fused_density || h->onsite || (h->gen_id != 1 && h->gen_id != 3)) return false;
    if (h->z != 3 && h->z != 6) return false;
    if (A.colour_last != A.colour || A.inj_xi) return false;
    if (tma_smem_bytes(h) > 200 * 1024) return false;
    if ((uint64_t)h->Ns * 4 * h->esz() * TMA_RC >= (1ull << 32)) return false;      // 32-bit replica offsets inside a chunk
    return h->tma.cls[A.colour].ok != 0;
}

template <int GEN, int Z, int STREAM, int HITS, bool TC, int CAPC = 0>
static int32_t launch_one(scmc_handle* h, const SweepArgs& A, cudaStream_t stream) {
    if constexpr (CAPC > 0) {
        if (h->tma.cap > CAPC) return launch_one<GEN, Z, STREAM, HITS, TC, 0>(h, A, stream);
    }
    using R = float;
    auto M = make_model_t<R>(h);
    const auto& C = h->tma.cls[A.colour];
    TmaArgs T;
    T.meta = (const uint4*)h->tma.d_meta; T.runs = (const uint2*)h->tma.d_runs; T.tile = h->tma.d_tile;
    T.first = C.first; T.cap = h->tma.cap;
    T.rk = philox_keys7(A.seed);
    // replicas per CTA: up to 64, fewer when that leaves the chip with less than about one wave of CTAs (small shards).  Measured (bench.py, sustained):
    // a threshold of 650 against 3000 CTAs per launch: 256 replicas 4.74 against 4.28e10, 1024 replicas (the 8-GPU shard) 5.21 against 4.94e10, 8192 equal
    const int rc_env = getenv("SCMC_TMA_RC") ? atoi(getenv("SCMC_TMA_RC")) : 0, min_ctas = getenv("SCMC_TMA_MINCTAS") ? atoi(getenv("SCMC_TMA_MINCTAS")) : 650;       // (read per launch: tests switch them)
    int rc = rc_env > 0 && rc_env <= TMA_RC ? rc_env : 64;      // measured on C5: 16 -> 5.18e10, 32 -> 5.38e10, 64 -> 5.40e10, 128 -> 5.39e10 (tensor-core variant)
    while (rc > 2 && (int64_t)C.n_tiles * ((A.rcount + rc - 1) / rc) < min_ctas) rc /= 2;
    T.rc = rc;
    const size_t smem = tma_smem_bytes(h, TC, CAPC);
    static size_t attr_set[64] = {0};      // per device: largest dynamic shared-memory size opted into so far
    if (attr_set[h->device & 63] < smem) {
        SCMC_CK(cudaFuncSetAttribute(k_sweep_tma_psi<R, GEN, Z, STREAM, HITS, TC, CAPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[h->device & 63] = smem;
    }
    for (int64_t y0 = 0; y0 < A.rcount; y0 += (int64_t)65535 * rc) {
        SweepArgs B = A;
        B.r0 = A.r0 + y0; B.rcount = std::min<int64_t>((int64_t)65535 * rc, A.rcount - y0);
        B.Jqt = h->d_Jqt; B.Jt = h->d_Jt;
        const dim3 grid((unsigned)C.n_tiles, (unsigned)((B.rcount + rc - 1) / rc));
        k_sweep_tma_psi<R, GEN, Z, STREAM, HITS, TC, CAPC><<<grid, TMA_TILE, smem, stream>>>(M, B, T);
        h->launches++;
    }
    SCMC_CK(cudaGetLastError());
    h->T_stale = true;
    return SCMC_OK;
}

int32_t launch_colour_pass_tma(scmc_handle* h, const SweepArgs& A, cudaStream_t stream) {
    if (A.rcount == 0) return SCMC_OK;
    static const bool unroll2 = !(getenv("SCMC_TMA_UNROLL2") && atoi(getenv("SCMC_TMA_UNROLL2")) == 0);
    const bool tc = tma_use_tc(h);
    auto go = [&](auto gen, auto zz) -> int32_t {
        constexpr int GEN = decltype(gen)::value, Z = decltype(zz)::value;
        if (tc) {
            // SCMC_TMA_FIXCAP=0 (read per launch: a test switches it): the run-time-capacity instantiation also where the compile-time one applies (same chain)
            const bool fixcap = !(getenv("SCMC_TMA_FIXCAP") && atoi(getenv("SCMC_TMA_FIXCAP")) == 0);
            if (A.hits == 2 && unroll2 && A.stream == 2 && fixcap) return launch_one<GEN, Z, 2, 2, true, SCMC_TMA_CAPC>(h, A, stream);
            if (A.hits == 2 && unroll2) return A.stream == 2 ? launch_one<GEN, Z, 2, 2, true>(h, A, stream) : launch_one<GEN, Z, 1, 2, true>(h, A, stream);
            return A.stream == 2 ? launch_one<GEN, Z, 2, 0, true>(h, A, stream) : launch_one<GEN, Z, 1, 0, true>(h, A, stream);
        }
        if (A.hits == 2 && unroll2) return A.stream == 2 ? launch_one<GEN, Z, 2, 2, false>(h, A, stream) : launch_one<GEN, Z, 1, 2, false>(h, A, stream);
        return A.stream == 2 ? launch_one<GEN, Z, 2, 0, false>(h, A, stream) : launch_one<GEN, Z, 1, 0, false>(h, A, stream);
    };
    using I1 = std::integral_constant<int, 1>; using I3 = std::integral_constant<int, 3>; using I6 = std::integral_constant<int, 6>;
    if (h->gen_id == 1) return h->z == 6 ? go(I1{}, I6{}) : go(I1{}, I3{});
    return h->z == 6 ? go(I3{}, I6{}) : go(I3{}, I3{});
}

}  // namespace scmc
