// tc_matvec.cu -- stand-alone check and timing of the 16x16 mat-vec of a site visit on the 5th-generation tensor cores (sm_100a).
//
//   Hq[site][a] = sum_b Jq[a][b] P[site][b]   for the 128 sites of a tile = one tcgen05.mma tile D[128x16] = A[128x16] B[16x16]^T,
//   3xTF32 (P = P_hi + P_lo, Jq = J_hi + J_lo; P_lo J_hi + P_hi J_lo + P_hi J_hi, FP32 accumulation in TMEM).
//   A comes from TMEM (tcgen05.st: a thread's TMEM lane IS its site), B from shared memory (K-major, no swizzle, 8x16-byte core matrices),
//   D is read back with tcgen05.ld.  Every WARP issues its own MMAs on the CTA's A region into its own D region and reads back only its
//   32 lanes, so the warps of a CTA never synchronise with each other (rows of other warps that a warp's MMA reads are ignored).
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_matvec tc_matvec.cu ; run: ./tc_matvec [lbo] [sbo]
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__host__ __device__ inline uint64_t make_b_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;          // descriptor version of sm_100
    return d;                 // no swizzle, base offset 0
}
// kind::tf32, D = F32, A/B = TF32, both K-major, N = 16, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

constexpr int TMEM_COLS = 128;      // A_hi 16 + A_lo 16 + 4 warps x D 16 = 96 -> 128

// mode 0: tensor cores, 1: FFMA2 from shared memory (what the production kernel does today)
template <int MODE>
__global__ void __launch_bounds__(128, 4) k_matvec(const float* __restrict__ Pin, const float* __restrict__ Jq, float* __restrict__ out, int visits, uint32_t lbo,
                                                   uint32_t sbo, long long* cycles) {
    __shared__ __align__(128) float sB[2][256];      // J_hi, J_lo in core-matrix order
    __shared__ __align__(16) float sJt[256];         // transposed, for the FFMA2 path
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < 256; t += 128) {
        const int a = t / 16, b = t % 16;            // B[n = a][k = b] = Jq[a][b]
        const float j = Jq[t];
        const float hi = __uint_as_float(__float_as_uint(j) & 0xFFFFE000u), lo = j - hi;
        // core matrix (n_blk = a / 8, k_chunk = b / 4): 8 rows x 16 bytes; address = k_chunk * LBO + n_blk * SBO + (a % 8) * 16 + (b % 4) * 4
        const uint32_t off = (b / 4) * lbo + (a / 8) * sbo + (a % 8) * 16 + (b % 4) * 4;
        sB[0][off / 4] = hi;
        sB[1][off / 4] = lo;
        sJt[b * 16 + a] = j;
    }
    if (warp == 0) {
        if (MODE == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (lane == 0) {
            for (int w = 0; w < 4; w++) mbar_init(smem_u32(bars + w), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // sB written by the generic proxy, read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_hi = tm + 0, a_lo = tm + 16, d_col = tm + 32 + 16 * warp;
    const uint32_t bar = smem_u32(bars + warp);
    const uint64_t bhi0 = make_b_desc(smem_u32(&sB[0][0]), lbo, sbo), blo0 = make_b_desc(smem_u32(&sB[1][0]), lbo, sbo);
    const uint32_t kstep = (2 * lbo) >> 4;           // descriptor start-address step of one K = 8 slice (two 16-byte chunks)
    const size_t site = (size_t)blockIdx.x * 128 + tid;
    const long long t0 = clock64();
    float acc[16];
    for (int i = 0; i < 16; i++) acc[i] = 0.f;
    for (int v = 0; v < visits; v++) {
        float P[16];
        const float4* src = reinterpret_cast<const float4*>(Pin + ((size_t)v * gridDim.x * 128 + site) * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) { const float4 x = src[q]; P[4 * q] = x.x + acc[4 * q]; P[4 * q + 1] = x.y; P[4 * q + 2] = x.z; P[4 * q + 3] = x.w; }
        float H[16];
        if (MODE == 0) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                hi[i] = __float_as_uint(P[i]) & 0xFFFFE000u;
                lo[i] = __float_as_uint(P[i] - __uint_as_float(hi[i]));
            }
            tmem_st16(lane_base | a_hi, hi);
            tmem_st16(lane_base | a_lo, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    tc_mma_tf32_ts(d_col, a_lo + 8 * s, bhi0 + (uint64_t)(s * kstep), IDESC, s);
                }
#pragma unroll
                for (int s = 0; s < 2; s++) tc_mma_tf32_ts(d_col, a_hi + 8 * s, blo0 + (uint64_t)(s * kstep), IDESC, 1);
#pragma unroll
                for (int s = 0; s < 2; s++) tc_mma_tf32_ts(d_col, a_hi + 8 * s, bhi0 + (uint64_t)(s * kstep), IDESC, 1);
                tc_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, v & 1);
            tc_fence_after();
            uint32_t d[16];
            tmem_ld16(lane_base | d_col, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; i++) H[i] = __uint_as_float(d[i]);
        } else {
            unsigned long long hp[8];
#pragma unroll
            for (int i = 0; i < 8; i++) hp[i] = 0ull;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                unsigned long long sb;
                asm("mov.b64 %0, {%1, %1};" : "=l"(sb) : "f"(P[b]));
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 j4 = *reinterpret_cast<const float4*>(sJt + b * 16 + 4 * q);
                    unsigned long long j01, j23;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(j01) : "f"(j4.x), "f"(j4.y));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(j23) : "f"(j4.z), "f"(j4.w));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp[2 * q]) : "l"(j01), "l"(sb));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp[2 * q + 1]) : "l"(j23), "l"(sb));
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) asm("mov.b64 {%0, %1}, %2;" : "=f"(H[2 * i]), "=f"(H[2 * i + 1]) : "l"(hp[i]));
        }
        float4* dst = reinterpret_cast<float4*>(out + ((size_t)v * gridDim.x * 128 + site) * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) dst[q] = make_float4(H[4 * q], H[4 * q + 1], H[4 * q + 2], H[4 * q + 3]);
        acc[0] = H[0] * 1e-30f;                        // a (numerically invisible) dependence of the next visit on this one
    }
    const long long t1 = clock64();
    if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
    if (MODE == 0) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(TMEM_COLS) : "memory");
    }
}

// throughput of back-to-back small MMAs: one thread issues `n` MMAs (M = 128, N = NN, K = 8, tf32, A from TMEM, B from shared memory), commits, waits
template <int NN>
__global__ void __launch_bounds__(128, 1) k_mma_rate(int n, long long* cycles) {
    __shared__ __align__(128) float sB[256 * 8 * 2];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < 256 * 8 * 2; t += 128) sB[t] = 1.0f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem;
    uint32_t z[16];
    for (int i = 0; i < 16; i++) z[i] = 0x3f800000u;
    tmem_st16(((uint32_t)(warp * 32) << 16) | tm, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    constexpr uint32_t IDN = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t bd = make_b_desc(smem_u32(sB), NN * 16, 128);      // K chunk stride = NN rows x 16 bytes
    if (warp == 0) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const long long t0 = clock64();
        if (leader) {
            tc_fence_after();
            for (int i = 0; i < n; i++) tc_mma_tf32_ts(tm + 256, tm + 8 * (i & 1), bd, IDN, 1);
            tc_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

template <int NN>
static void mma_rate(long long* dC) {
    for (int n : {64, 1024}) {
        k_mma_rate<NN><<<148, 128>>>(n, dC);
        CK(cudaDeviceSynchronize());
        long long c[148];
        CK(cudaMemcpy(c, dC, sizeof c, cudaMemcpyDeviceToHost));
        printf("MMA rate M=128 N=%d K=8 tf32, A from TMEM: %d back-to-back MMAs on every SM: %.1f cycles per MMA (incl. commit + wait)\n", NN, n, (double)c[0] / n);
    }
}

int main(int argc, char** argv) {
    const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 256, sbo = argc > 2 ? atoi(argv[2]) : 128;
    const int blocks = argc > 3 ? atoi(argv[3]) : 148 * 4, visits = argc > 4 ? atoi(argv[4]) : 16;
    const size_t n = (size_t)visits * blocks * 128 * 16;
    std::vector<float> P(n), J(256), out(n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    for (auto& x : P) x = (float)(6.0 * rnd() - 2.0);
    for (auto& x : J) x = (float)(2.0 * rnd() - 1.0);
    float *dP, *dJ, *dO;
    long long* dC;
    CK(cudaMalloc(&dP, n * 4)); CK(cudaMalloc(&dJ, 1024)); CK(cudaMalloc(&dO, n * 4)); CK(cudaMalloc(&dC, blocks * 8));
    CK(cudaMemcpy(dP, P.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dJ, J.data(), 1024, cudaMemcpyHostToDevice));
    if (argc > 5) { mma_rate<16>(dC); mma_rate<32>(dC); mma_rate<64>(dC); mma_rate<256>(dC); return 0; }
    for (int mode = 0; mode < 2; mode++) {
        CK(cudaMemset(dO, 0, n * 4));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) k_matvec<0><<<blocks, 128>>>(dP, dJ, dO, visits, lbo, sbo, dC);
            else k_matvec<1><<<blocks, 128>>>(dP, dJ, dO, visits, lbo, sbo, dC);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        CK(cudaMemcpy(out.data(), dO, n * 4, cudaMemcpyDeviceToHost));
        std::vector<long long> cyc(blocks);
        CK(cudaMemcpy(cyc.data(), dC, blocks * 8, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0, sum_cyc = 0;
        for (size_t r = 0; r < n / 16; r++) {
            for (int a = 0; a < 16; a++) {
                double ref = 0;
                for (int b = 0; b < 16; b++) ref += (double)J[a * 16 + b] * (double)P[r * 16 + b];
                max_err = std::fmax(max_err, std::fabs(ref - (double)out[r * 16 + a]));
                max_ref = std::fmax(max_ref, std::fabs(ref));
            }
        }
        for (auto c : cyc) sum_cyc += (double)c;
        printf("mode %d (%s) lbo %u sbo %u: max |err| %.3e (max |ref| %.2f), %.3f ms, %.0f cycles per visit and CTA (%d CTAs, %d visits)\n", mode,
               mode == 0 ? "tcgen05 3xTF32" : "FFMA2", lbo, sbo, max_err, max_ref, best, sum_cyc / blocks / visits, blocks, visits);
    }
    return 0;
}
