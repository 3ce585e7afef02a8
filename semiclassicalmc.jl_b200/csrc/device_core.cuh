// device_core.cuh -- per-site arithmetic shared by every kernel of libscmc_b200 (sm_100a).
//
// What it replaces in the reference (paths relative to /root/reference/):
//   stream_v1            <- randn!(local_rng(), ...) / rand()      src/MonteCarlo.jl:255-256, 278   (Philox4x32-10, counter based)
//   propose              <- getRandomState!                         src/MonteCarlo.jl:249-260
//   GenOps<>::expect     <- computeSpinExpectation!                 src/Configuration.jl:166-194     (sparse, unrolled, in registers)
//   metropolis_hit       <- localUpdate! + getEnergyDifference!     src/MonteCarlo.jl:263-286, src/Configuration.jl:278-298
//   field_from_sums      <- the per-neighbour exchangeEnergy mat-vecs src/Configuration.jl:199-218, 292-296, regrouped as
//                           h = sum_l J_l (sum_{j in label l} T_j) so that ONE 16x16 mat-vec per label serves all hits of a visit.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "gen_code.h"

namespace scmc {

constexpr int NG = 16;        // number of generators
#ifndef SCMC_MAXL
#define SCMC_MAXL 3           // build-time: most coupling labels of a model (kernels are unrolled over the labels; -DSCMC_MAXL=6 builds a six-label library)
#endif
constexpr int MAXC = 16;      // max colours (colour-class offsets travel in the kernel parameters)
constexpr int MAXL = SCMC_MAXL;
constexpr uint64_t VISIT_INIT = (1ull << 60) - 1;  // reserved visit number of scmc_randomize

template <class R> struct V4;
template <> struct V4<float> { using type = float4; };
template <> struct V4<double> { using type = double4; };
template <class R> using vec4 = typename V4<R>::type;

template <class R> __device__ __forceinline__ vec4<R> make_v4(R a, R b, R c, R d);
template <> __device__ __forceinline__ float4 make_v4<float>(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
template <> __device__ __forceinline__ double4 make_v4<double>(double a, double b, double c, double d) { return make_double4(a, b, c, d); }

// ------------------------------------------------------------------------------------------------ generator ops
template <int GEN> struct GenOps;
template <> struct GenOps<1> {
    static constexpr int D = 4;
    template <class R> static __device__ __forceinline__ void expect(const R* re, const R* im, R* T) { gen::spin_expect_g1(re, im, T); }
    template <class R> static __device__ __forceinline__ void expect_sq(const R* re, const R* im, R* T) { gen::spin_expect_sq_g1(re, im, T); }
    template <class R, bool O> static __device__ __forceinline__ void hloc(const R* h, const R* K, R* Hr, R* Hi) { gen::hloc_g1<R, O>(h, K, Hr, Hi); }
    static constexpr int NDENS = gen::NDENS_g1;
    template <class R> static __device__ __forceinline__ void density_acc(const R* re, const R* im, R* P) { gen::density_acc_g1(re, im, P); }
    template <class R> static __device__ __forceinline__ void expect_from_density(const R* P, R* T) { gen::expect_from_density_g1(P, T); }
    template <class R> static __device__ __forceinline__ void hq_from_field(const R* h, R* Hq) { gen::hq_from_field_g1(h, Hq); }
    template <class R> static __device__ __forceinline__ R energy_from_state(const R* re, const R* im, const R* Hq) { return gen::energy_from_state_g1(re, im, Hq); }
};
template <> struct GenOps<2> {
    static constexpr int D = 6;
    template <class R> static __device__ __forceinline__ void expect(const R* re, const R* im, R* T) { gen::spin_expect_g2(re, im, T); }
    template <class R> static __device__ __forceinline__ void expect_sq(const R* re, const R* im, R* T) { gen::spin_expect_sq_g2(re, im, T); }
    template <class R, bool O> static __device__ __forceinline__ void hloc(const R* h, const R* K, R* Hr, R* Hi) { gen::hloc_g2<R, O>(h, K, Hr, Hi); }
    static constexpr int NDENS = gen::NDENS_g2;
    template <class R> static __device__ __forceinline__ void density_acc(const R* re, const R* im, R* P) { gen::density_acc_g2(re, im, P); }
    template <class R> static __device__ __forceinline__ void expect_from_density(const R* P, R* T) { gen::expect_from_density_g2(P, T); }
    template <class R> static __device__ __forceinline__ void hq_from_field(const R* h, R* Hq) { gen::hq_from_field_g2(h, Hq); }
    template <class R> static __device__ __forceinline__ R energy_from_state(const R* re, const R* im, const R* Hq) { return gen::energy_from_state_g2(re, im, Hq); }
};
template <> struct GenOps<3> {
    static constexpr int D = 4;
    template <class R> static __device__ __forceinline__ void expect(const R* re, const R* im, R* T) { gen::spin_expect_g3(re, im, T); }
    template <class R> static __device__ __forceinline__ void expect_sq(const R* re, const R* im, R* T) { gen::spin_expect_sq_g3(re, im, T); }
    template <class R, bool O> static __device__ __forceinline__ void hloc(const R* h, const R* K, R* Hr, R* Hi) { gen::hloc_g3<R, O>(h, K, Hr, Hi); }
    static constexpr int NDENS = gen::NDENS_g3;
    template <class R> static __device__ __forceinline__ void density_acc(const R* re, const R* im, R* P) { gen::density_acc_g3(re, im, P); }
    template <class R> static __device__ __forceinline__ void expect_from_density(const R* P, R* T) { gen::expect_from_density_g3(P, T); }
    template <class R> static __device__ __forceinline__ void hq_from_field(const R* h, R* Hq) { gen::hq_from_field_g3(h, Hq); }
    template <class R> static __device__ __forceinline__ R energy_from_state(const R* re, const R* im, const R* Hq) { return gen::energy_from_state_g3(re, im, Hq); }
};

// Caller-supplied generator tables (any 16 Hermitian d x d matrices, d = 4 or 6 -- e.g. a parton operator WITH the Jordan-Wigner sign the
// reference's f omits): no unrolled sparse code exists for them, so T^a = Re psi^dagger G^a psi is the reference's dense
// contraction (computeSpinExpectation!, src/Configuration.jl:166-194) over tables in constant memory.  One table set per translation unit
// and device (set_custom_tables below); the kernels of the T-gather family use it, always with the on-site term and MAXL labels.
struct CustomTables { double gre[NG * 36], gim[NG * 36], qre[NG * 36], qim[NG * 36]; };      // [a][j][k], row stride d
static __constant__ CustomTables c_custom;
template <int D_> struct GenOpsCustom {
    static constexpr int D = D_;
    static constexpr int NDENS = 0;            // no density basis: the psi-gather family is not offered
    template <class R> static __device__ __forceinline__ void contract(const double* gr, const double* gi, const R* re, const R* im, R* T) {
#pragma unroll 1
        for (int a = 0; a < NG; a++) {
            R acc = R(0);
#pragma unroll
            for (int j = 0; j < D; j++)
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const R g_r = (R)gr[(a * D + j) * D + k], g_i = (R)gi[(a * D + j) * D + k];
                    acc += (re[j] * re[k] + im[j] * im[k]) * g_r - (re[j] * im[k] - im[j] * re[k]) * g_i;
                }
            T[a] = acc;
        }
    }
    template <class R> static __device__ __forceinline__ void expect(const R* re, const R* im, R* T) { contract<R>(c_custom.gre, c_custom.gim, re, im, T); }
    template <class R> static __device__ __forceinline__ void expect_sq(const R* re, const R* im, R* T) { contract<R>(c_custom.qre, c_custom.qim, re, im, T); }
    // upper triangle of H_loc = sum_a h^a G^a + K^a (G^a)^2
    template <class R, bool O> static __device__ __forceinline__ void hloc(const R* h, const R* K, R* Hr, R* Hi) {
#pragma unroll
        for (int j = 0; j < D; j++)
#pragma unroll
            for (int k = j; k < D; k++) {
                R hr = R(0), hi = R(0);
#pragma unroll 1
                for (int a = 0; a < NG; a++) {
                    hr += h[a] * (R)c_custom.gre[(a * D + j) * D + k]; hi += h[a] * (R)c_custom.gim[(a * D + j) * D + k];
                    if (O) { hr += K[a] * (R)c_custom.qre[(a * D + j) * D + k]; hi += K[a] * (R)c_custom.qim[(a * D + j) * D + k]; }
                }
                Hr[j * D + k] = hr; Hi[j * D + k] = hi;
            }
    }
};
template <> struct GenOps<4> : GenOpsCustom<4> {};
template <> struct GenOps<5> : GenOpsCustom<6> {};
// host: make `t` the table set of the calling translation unit on the current device (synchronises the device when another handle's set is live)
static inline cudaError_t set_custom_tables(const CustomTables& t, const void* owner, int device) {
    static const void* live[64] = {nullptr};
    if (live[device & 63] == owner) return cudaSuccess;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_custom, &t, sizeof(CustomTables));
    if (e == cudaSuccess) live[device & 63] = owner;
    return e;
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void mulhilo(uint32_t m, uint32_t x, uint32_t& hi, uint32_t& lo) {
    // one IMAD.WIDE.U32; written in PTX so the 64-bit product is split without extra carry/IADD3 instructions
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(m), "r"(x));
}
template <int ROUNDS>
__device__ __forceinline__ void philox4x32(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, hi0, lo0);
        mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ k0; c2 = hi0 ^ c3 ^ k1;
        c1 = lo1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) { philox4x32<10>(c0, c1, c2, c3, k0, k1); }

// math flavours: FP32 uses the SFU approximations (MUFU.LG2/SIN/COS/EX2/RSQ), FP64 the accurate library routines
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& g0, float& g1) {
    const float u1 = (float)((wa >> 8) + 1u) * 5.9604644775390625e-08f;   // (0,1]
    const float ang = (float)(wb >> 8) * (6.283185307179586f * 5.9604644775390625e-08f);
    const float r = sqrtf(-1.3862943611198906f * __log2f(u1));           // sqrt(-2 ln u1)
    float s, c;
    __sincosf(ang, &s, &c);
    g0 = r * c; g1 = r * s;
}
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, double& g0, double& g1) {
    const double u1 = (double)((wa >> 8) + 1u) * (1.0 / 16777216.0);
    const double u2 = (double)(wb >> 8) * (1.0 / 16777216.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    g0 = r * c; g1 = r * s;
}
__device__ __forceinline__ float unit_uniform(uint32_t x, float) { return __uint2float_rz(x) * 2.3283064365386963e-10f; }
__device__ __forceinline__ double unit_uniform(uint32_t x, double) { return (double)x * (1.0 / 4294967296.0); }
__device__ __forceinline__ float inv_sqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double inv_sqrt(double x) { return 1.0 / sqrt(x); }

// "stream v1": the random numbers of update number `visit` of (site, replica).  xr/xi = real / imaginary normals.
template <class R, int D>
__device__ __forceinline__ void stream_v1(uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, R* xr, R* xi, R& u) {
    constexpr int NCALL = (2 * D + 3) / 4;
    uint32_t w[4 * NCALL];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t vlo = (uint32_t)visit, vhi = (uint32_t)(visit >> 32) << 4;
#pragma unroll
    for (int call = 0; call < NCALL; call++) {
        uint32_t c0 = site, c1 = replica, c2 = vlo, c3 = vhi | (uint32_t)call;
        philox4x32_10(c0, c1, c2, c3, k0, k1);
        w[4 * call + 0] = c0; w[4 * call + 1] = c1; w[4 * call + 2] = c2; w[4 * call + 3] = c3;
    }
#pragma unroll
    for (int m = 0; m < D; m++) box_muller(w[2 * m], w[2 * m + 1], xr[m], xi[m]);
    const uint32_t x = (w[0] & 0xFFu) | ((w[2] & 0xFFu) << 8) | ((w[4] & 0xFFu) << 16) | ((w[6] & 0xFFu) << 24);
    u = unit_uniform(x, R(0));
}

// "stream v2" (scmc_set_stream): fewer wide multiplies per update.  Philox4x32-7 (the Crush-resistant round count of
// Salmon et al., SC'11) instead of -10; ONE call gives the four complex normals of an update (d = 4; two calls for d = 6): word m holds a
// 16-bit radius index and a 16-bit angle, u1 = (2 k1 + 1) / 2^17 in (0, 1), u2 = k2 / 2^16; the acceptance uniforms of four consecutive
// visits come from one further call (counter word 3 tagged 8), word visit & 3.  Same counter layout as v1 otherwise.
__device__ __forceinline__ void box_muller16(uint32_t w, float& g0, float& g1) {
    const float a = (float)((w >> 15) | 1u);                                  // 2 k1 + 1
    const float ang = (float)(w & 0xFFFFu) * (6.283185307179586f * 1.52587890625e-05f);
    const float r = sqrtf(fmaf(__log2f(a), -1.3862943611198906f, 17.0f * 1.3862943611198906f));      // sqrt(-2 ln(a / 2^17))
    float s, c;
    __sincosf(ang, &s, &c);
    g0 = r * c; g1 = r * s;
}
__device__ __forceinline__ void box_muller16(uint32_t w, double& g0, double& g1) {
    const double u1 = (double)((w >> 15) | 1u) * (1.0 / 131072.0);
    const double u2 = (double)(w & 0xFFFFu) * (1.0 / 65536.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    g0 = r * c; g1 = r * s;
}
template <class R, int D>
__device__ __forceinline__ void stream_v2_normals(uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, R* xr, R* xi) {
    constexpr int NCALL = (D + 3) / 4;
    uint32_t w[4 * NCALL];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t vlo = (uint32_t)visit, vhi = (uint32_t)(visit >> 32) << 4;
#pragma unroll
    for (int call = 0; call < NCALL; call++) {
        uint32_t c0 = site, c1 = replica, c2 = vlo, c3 = vhi | (uint32_t)call;
        philox4x32<7>(c0, c1, c2, c3, k0, k1);
        w[4 * call + 0] = c0; w[4 * call + 1] = c1; w[4 * call + 2] = c2; w[4 * call + 3] = c3;
    }
#pragma unroll
    for (int m = 0; m < D; m++) box_muller16(w[m], xr[m], xi[m]);
}
// acceptance-uniform words of visits 4 g .. 4 g + 3
__device__ __forceinline__ void stream_v2_uniform_words(uint64_t seed, uint32_t site, uint32_t replica, uint64_t group, uint32_t* w) {
    uint32_t c0 = site, c1 = replica, c2 = (uint32_t)group, c3 = ((uint32_t)(group >> 32) << 4) | 8u;
    philox4x32<7>(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}
// The same two calls with the round keys k0 + r W0, k1 + r W1 (r = 0..6) precomputed on the host: passed in the kernel parameters they are
// constant-bank operands of the xor instead of 12 uniform adds per site visit.  Same words as stream_v2_normals / stream_v2_uniform_words.
struct PhiloxKeys7 { uint32_t k[14]; };
static inline PhiloxKeys7 philox_keys7(uint64_t seed) {
    PhiloxKeys7 K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 7; r++) { K.k[2 * r] = k0; K.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return K;
}
__device__ __forceinline__ void philox4x32_7_rk(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, const PhiloxKeys7& K) {
#pragma unroll
    for (int r = 0; r < 7; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, hi0, lo0);
        mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ K.k[2 * r]; c2 = hi0 ^ c3 ^ K.k[2 * r + 1];
        c1 = lo1; c3 = lo0;
    }
}
template <class R>
__device__ __forceinline__ void stream_v2_normals_d4(const PhiloxKeys7& K, uint32_t site, uint32_t replica, uint64_t visit, R* xr, R* xi) {
    uint32_t c0 = site, c1 = replica, c2 = (uint32_t)visit, c3 = (uint32_t)(visit >> 32) << 4;
    philox4x32_7_rk(c0, c1, c2, c3, K);
    box_muller16(c0, xr[0], xi[0]); box_muller16(c1, xr[1], xi[1]); box_muller16(c2, xr[2], xi[2]); box_muller16(c3, xr[3], xi[3]);
}
__device__ __forceinline__ void stream_v2_uniform_words(const PhiloxKeys7& K, uint32_t site, uint32_t replica, uint64_t group, uint32_t* w) {
    uint32_t c0 = site, c1 = replica, c2 = (uint32_t)group, c3 = ((uint32_t)(group >> 32) << 4) | 8u;
    philox4x32_7_rk(c0, c1, c2, c3, K);
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}
__device__ __forceinline__ uint32_t pick_word(const uint32_t* w, uint32_t i) { return i == 0 ? w[0] : i == 1 ? w[1] : i == 2 ? w[2] : w[3]; }
// one update's numbers as a pure function of (seed, site, replica, visit) -- the form the kernels without a cached uniform call use
template <class R, int D>
__device__ __forceinline__ void stream_v2(uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, R* xr, R* xi, R& u) {
    stream_v2_normals<R, D>(seed, site, replica, visit, xr, xi);
    uint32_t w[4];
    stream_v2_uniform_words(seed, site, replica, visit >> 2, w);
    u = unit_uniform(pick_word(w, (uint32_t)visit & 3u), R(0));
}
template <class R, int D>
__device__ __forceinline__ void stream_any(int version, uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, R* xr, R* xi, R& u) {
    if (version == 2) stream_v2<R, D>(seed, site, replica, visit, xr, xi, u);
    else stream_v1<R, D>(seed, site, replica, visit, xr, xi, u);
}

// ------------------------------------------------------------------------------------------------ local field
// h^a = sum_l sum_b J_l^{ab} S_l^b ; J in kernel-parameter (constant-bank) space so the FMAs take it as an operand.
template <class R, int NL>
struct Couplings {
    R J[NL][NG * NG];
    R K[NG];
};
template <class R, int NL>
__device__ __forceinline__ void field_from_sums(const Couplings<R, NL>& cp, const R (*S)[NG], R* h) {
#pragma unroll
    for (int a = 0; a < NG; a++) {
        R acc = R(0);
#pragma unroll
        for (int l = 0; l < NL; l++)
#pragma unroll
            for (int b = 0; b < NG; b++) acc = fma(cp.J[l][a * NG + b], S[l][b], acc);
        h[a] = acc;
    }
}

// Shared-memory variant used by the sweep kernels: sJt[l][b][a] = J_l^{ab} (transposed so that consecutive a are adjacent).
// FP32: 128-bit broadcast LDS + packed fma.rn.f32x2 (FFMA2: two h components per instruction), 64 LDS + 128 FFMA2 per label
// instead of 256 FFMA fed through spilling uniform registers.
template <int NL>
__device__ __forceinline__ void field_from_sums_smem(const float* sJt, const float (*S)[NG], float* h) {
    unsigned long long hp[NG / 2];
#pragma unroll
    for (int i = 0; i < NG / 2; i++) hp[i] = 0ull;
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
        for (int b = 0; b < NG; b++) {
            unsigned long long sb;
            asm("mov.b64 %0, {%1, %1};" : "=l"(sb) : "f"(S[l][b]));
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 j4 = *reinterpret_cast<const float4*>(sJt + (l * NG + b) * NG + 4 * q);
                unsigned long long j01, j23;
                asm("mov.b64 %0, {%1, %2};" : "=l"(j01) : "f"(j4.x), "f"(j4.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(j23) : "f"(j4.z), "f"(j4.w));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp[2 * q]) : "l"(j01), "l"(sb));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp[2 * q + 1]) : "l"(j23), "l"(sb));
            }
        }
#pragma unroll
    for (int i = 0; i < NG / 2; i++) asm("mov.b64 {%0, %1}, %2;" : "=f"(h[2 * i]), "=f"(h[2 * i + 1]) : "l"(hp[i]));
}
template <int NL>
__device__ __forceinline__ void field_from_sums_smem(const double* sJt, const double (*S)[NG], double* h) {
#pragma unroll
    for (int a = 0; a < NG; a++) h[a] = 0.0;
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
        for (int b = 0; b < NG; b++)
#pragma unroll
            for (int a = 0; a < NG; a++) h[a] = fma(sJt[(l * NG + b) * NG + a], S[l][b], h[a]);
}

// One label's share of field_from_sums_smem: h += J_l . S with the sum S of that label in 16 registers.  Called for l = 0, 1, ... with the
// accumulator carried along, it issues exactly the FMA sequence of field_from_sums_smem<NL> (same order => same bits).
__device__ __forceinline__ void field_accumulate_smem(const float* sJt_l, const float* S, unsigned long long* hp) {
#pragma unroll
    for (int b = 0; b < NG; b++) {
        unsigned long long sb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(sb) : "f"(S[b]));
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 j4 = *reinterpret_cast<const float4*>(sJt_l + b * NG + 4 * q);
            unsigned long long j01, j23;
            asm("mov.b64 %0, {%1, %2};" : "=l"(j01) : "f"(j4.x), "f"(j4.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(j23) : "f"(j4.z), "f"(j4.w));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp[2 * q]) : "l"(j01), "l"(sb));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(hp[2 * q + 1]) : "l"(j23), "l"(sb));
        }
    }
}
__device__ __forceinline__ void field_accumulate_smem(const double* sJt_l, const double* S, double* h) {
#pragma unroll
    for (int b = 0; b < NG; b++)
#pragma unroll
        for (int a = 0; a < NG; a++) h[a] = fma(sJt_l[b * NG + a], S[b], h[a]);
}

// ------------------------------------------------------------------------------------------------ one Metropolis hit
// e_loc = T.h (+ K.Tsq) is the part of the energy that depends on the visited site; dE = e_loc(psi') - e_loc(psi)
// (getEnergyDifference!, Configuration.jl:278-298, with the neighbour mat-vecs hoisted into h).
template <class R, bool ONSITE>
__device__ __forceinline__ R local_energy(const R* T, const R* Tsq, const R* h, const R* K) {
    R e = R(0);
#pragma unroll
    for (int a = 0; a < NG; a++) e = fma(T[a], h[a], e);
    if (ONSITE) {
#pragma unroll
        for (int a = 0; a < NG; a++) e = fma(Tsq[a], K[a], e);
    }
    return e;
}
// acceptance test rand() < exp(-beta dE) (MonteCarlo.jl:277-278): FP32 evaluates 2^(nbeta * dE) with nbeta = -beta * log2(e) on the SFU
__device__ __forceinline__ float neg_beta_scaled(float beta) { return -beta * 1.4426950408889634f; }
__device__ __forceinline__ double neg_beta_scaled(double beta) { return -beta; }
__device__ __forceinline__ float accept_prob(float nbeta, float dE) { return exp2f(nbeta * dE); }
__device__ __forceinline__ double accept_prob(double nbeta, double dE) { return exp(nbeta * dE); }

// State of the visited site lives in registers: psi (re, im), T, (Tsq), e_loc.  Returns true when the proposal was accepted.
template <class R, int GEN, bool ONSITE>
__device__ __forceinline__ bool metropolis_hit(R* re, R* im, R* T, R* Tsq, const R* h, const R* K, const R* xr, const R* xi,
                                               R u, R nbeta, R sigma, R& eloc, R& dE_out) {
    constexpr int D = GenOps<GEN>::D;
    R nre[D], nim[D], nT[NG], nTsq[NG];
    R n2 = R(0);
#pragma unroll
    for (int k = 0; k < D; k++) {
        nre[k] = fma(sigma, xr[k], re[k]);
        nim[k] = fma(sigma, xi[k], im[k]);
        n2 = fma(nre[k], nre[k], n2);
        n2 = fma(nim[k], nim[k], n2);
    }
    const R inv = inv_sqrt(n2);
#pragma unroll
    for (int k = 0; k < D; k++) { nre[k] *= inv; nim[k] *= inv; }
    GenOps<GEN>::expect(nre, nim, nT);
    if (ONSITE) GenOps<GEN>::expect_sq(nre, nim, nTsq);
    const R enew = local_energy<R, ONSITE>(nT, nTsq, h, K);
    const R dE = enew - eloc;
    dE_out = dE;
    const bool acc = u < accept_prob(nbeta, dE);   // no min(1, .): overflow to +inf accepts
    if (acc) {
        eloc = enew;
#pragma unroll
        for (int k = 0; k < D; k++) { re[k] = nre[k]; im[k] = nim[k]; }
#pragma unroll
        for (int a = 0; a < NG; a++) T[a] = nT[a];
        if (ONSITE) {
#pragma unroll
            for (int a = 0; a < NG; a++) Tsq[a] = nTsq[a];
        }
    }
    return acc;
}

// psi-gather family, no on-site term: the local energy is linear in the density matrix, e(psi) = sum_n Hq[n] P_n(psi) with
// Hq = C^T h built once per visit, so a hit needs neither T' nor a 16-term dot product (48 instead of 84 instructions).
// proposal psi' = normalize(psi + sigma xi) (getRandomState!, MonteCarlo.jl:249-260): two partial sums for the norm (re and im parts)
template <int D>
__device__ __forceinline__ void propose_state(const double* re, const double* im, const double* xr, const double* xi, double sigma, double* nre, double* nim) {
    double na = 0.0, nb = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        nre[k] = fma(sigma, xr[k], re[k]);
        nim[k] = fma(sigma, xi[k], im[k]);
        na = fma(nre[k], nre[k], na);
        nb = fma(nim[k], nim[k], nb);
    }
    const double inv = inv_sqrt(na + nb);
#pragma unroll
    for (int k = 0; k < D; k++) { nre[k] *= inv; nim[k] *= inv; }
}
// FP32: the same operations in the same order on (re, im) pairs as packed fma / mul .f32x2 (SASS FFMA2 / FMUL2: one issue slot per complex
// component instead of two; each half is the IEEE operation of the scalar form, so the result is bit for bit the scalar one)
template <int D>
__device__ __forceinline__ void propose_state(const float* re, const float* im, const float* xr, const float* xi, float sigma, float* nre, float* nim) {
#ifdef SCMC_SCALAR_PROPOSAL
    float na = 0.f, nb = 0.f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        nre[k] = fmaf(sigma, xr[k], re[k]);
        nim[k] = fmaf(sigma, xi[k], im[k]);
        na = fmaf(nre[k], nre[k], na);
        nb = fmaf(nim[k], nim[k], nb);
    }
    const float inv = inv_sqrt(na + nb);
#pragma unroll
    for (int k = 0; k < D; k++) { nre[k] *= inv; nim[k] *= inv; }
#else
    unsigned long long n[D], acc = 0ull, ss, ii;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(sigma));
#pragma unroll
    for (int k = 0; k < D; k++) {
        unsigned long long c, x;
        asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(re[k]), "f"(im[k]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(xr[k]), "f"(xi[k]));
        asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(n[k]) : "l"(ss), "l"(x), "l"(c));
        asm("fma.rn.ftz.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(n[k]));
    }
    float na, nb;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(na), "=f"(nb) : "l"(acc));
    const float inv = inv_sqrt(na + nb);
    asm("mov.b64 %0, {%1, %1};" : "=l"(ii) : "f"(inv));
#pragma unroll
    for (int k = 0; k < D; k++) {
        asm("mul.rn.ftz.f32x2 %0, %0, %1;" : "+l"(n[k]) : "l"(ii));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(nre[k]), "=f"(nim[k]) : "l"(n[k]));
    }
#endif
}

template <class R, int GEN>
__device__ __forceinline__ bool metropolis_hit_psi(R* re, R* im, const R* Hq, const R* xr, const R* xi, R u, R nbeta, R sigma, R& eloc, R& dE_out) {
    constexpr int D = GenOps<GEN>::D;
    R nre[D], nim[D];
    propose_state<D>(re, im, xr, xi, sigma, nre, nim);
    const R enew = GenOps<GEN>::energy_from_state(nre, nim, Hq);
    const R dE = enew - eloc;
    dE_out = dE;
    const bool acc = u < accept_prob(nbeta, dE);
    if (acc) {
        eloc = enew;
#pragma unroll
        for (int k = 0; k < D; k++) { re[k] = nre[k]; im[k] = nim[k]; }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------ plane <-> register helpers
// psi planes: NPV = D/2 vec4 per site, plane v holds (re[2v], im[2v], re[2v+1], im[2v+1]).  T planes: 4 vec4 per site.
template <class R, int D>
__device__ __forceinline__ void load_psi(const vec4<R>* base, size_t plane_stride, R* re, R* im) {
#pragma unroll
    for (int v = 0; v < D / 2; v++) {
        const vec4<R> q = base[v * plane_stride];
        re[2 * v] = q.x; im[2 * v] = q.y; re[2 * v + 1] = q.z; im[2 * v + 1] = q.w;
    }
}
template <class R, int D>
__device__ __forceinline__ void store_psi(vec4<R>* base, size_t plane_stride, const R* re, const R* im) {
#pragma unroll
    for (int v = 0; v < D / 2; v++) base[v * plane_stride] = make_v4<R>(re[2 * v], im[2 * v], re[2 * v + 1], im[2 * v + 1]);
}
template <class R>
__device__ __forceinline__ void load_T(const vec4<R>* base, size_t plane_stride, R* T) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const vec4<R> t = base[q * plane_stride];
        T[4 * q] = t.x; T[4 * q + 1] = t.y; T[4 * q + 2] = t.z; T[4 * q + 3] = t.w;
    }
}
template <class R>
__device__ __forceinline__ void add_T(const vec4<R>* base, size_t plane_stride, R* S) {
    vec4<R> t[4];
#pragma unroll
    for (int q = 0; q < 4; q++) t[q] = base[q * plane_stride];
#pragma unroll
    for (int q = 0; q < 4; q++) { S[4 * q] += t[q].x; S[4 * q + 1] += t[q].y; S[4 * q + 2] += t[q].z; S[4 * q + 3] += t[q].w; }
}
template <class R>
__device__ __forceinline__ void store_T(vec4<R>* base, size_t plane_stride, const R* T) {
#pragma unroll
    for (int q = 0; q < 4; q++) base[q * plane_stride] = make_v4<R>(T[4 * q], T[4 * q + 1], T[4 * q + 2], T[4 * q + 3]);
}

}  // namespace scmc
