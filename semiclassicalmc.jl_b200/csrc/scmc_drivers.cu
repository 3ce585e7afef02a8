// scmc_drivers.cu -- device-side schedules of the reference's drivers and the remaining observables.
//
// Reference functions replaced here (paths relative to /root/reference/):
//   scmc_run_metropolis   <- run!                      src/MonteCarlo.jl:3-108   (ramp :24-27, sigma rule :47-54, measure cadence :82-86)
//   scmc_anneal           <- runAnnealing! cooling     src/MonteCarlo.jl:131-204 (per-temperature window :151, cooling/sigma :178-192)
//   scmc_optimize         <- localOptimization! sweeps src/MonteCarlo.jl:214-225, 289-329 (per-site minimisation of the local energy)
//   scmc_structure_factor <- getCorrelations + computeStructureFactorVec  src/Observables/Observables.jl:22-41, 116-128
//   scmc_correlations     <- getCorrelations           src/Observables/Observables.jl:22-41
// Every replica carries its own (beta, sigma, counters, stop flag); the host only issues launches.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>
#include <vector>
#include "handle.h"

namespace scmc {

// beta_r for thermalisation sweep `sweep` (1-based): geometric ramp over the first N_anneal sweeps, then flat (MonteCarlo.jl:24-27)
template <class R>
__global__ void k_set_beta(int64_t nrep, const double* T, const double* Ti, int64_t sweep, int64_t n_anneal, R* beta) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrep) return;
    double t = T[r];
    if (sweep >= 1 && sweep <= n_anneal && n_anneal > 1) t = Ti[r] * pow(T[r] / Ti[r], (double)(sweep - 1) / (double)(n_anneal - 1));
    beta[r] = (R)(1.0 / t);
}
// sigma <- min(sigma * 0.5 / (1 - R_acc), 60) with R_acc = accepted / attempted since the last call (MonteCarlo.jl:50-53)
template <class R>
__global__ void k_adapt_sigma(int64_t nrep, unsigned long long* acc, double attempted, R* sigma) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrep) return;
    const double ratio = (double)acc[r] / attempted;
    const double s = fmin((double)sigma[r] * 0.5 / (1.0 - ratio), 60.0);   // ratio == 1 -> inf -> 60
    sigma[r] = (R)s;
    acc[r] = 0ull;
}
// per-replica annealing controller, run after every sweep (MonteCarlo.jl:151-203)
template <class R>
__global__ void k_anneal_ctrl(int64_t nrep, unsigned long long* acc, int64_t* sweeps_at_T, int64_t* sweeps_done, uint8_t* active,
                              R* beta, R* sigma, double att_per_sweep, double min_accepted, int64_t n_per_T, int last_sweep,
                              double beta_fac, double min_accrate, double sigma_min, int* n_active) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrep || !active[r]) return;
    const int64_t sat = sweeps_at_T[r] + 1;
    sweeps_done[r] += 1;
    const double a = (double)acc[r];
    bool still = true;
    if (!(a < min_accepted && sat < n_per_T && !last_sweep)) {
        const double ratio = a / (att_per_sweep * (double)sat);
        if (ratio > min_accrate) {
            beta[r] = (R)((double)beta[r] * beta_fac);
            sigma[r] = (R)fmax(fmin((double)sigma[r] * 0.5 / (1.0 - ratio), 60.0), sigma_min);
            acc[r] = 0ull;
            sweeps_at_T[r] = 0;
        } else {
            active[r] = 0;
            still = false;
        }
    } else {
        sweeps_at_T[r] = sat;
    }
    if (still && !last_sweep) atomicAdd(n_active, 1);
}

// ------------------------------------------------------------------------------------------------ local optimisation
// Minimise <psi|H_loc|psi>, H_loc = sum_a h^a G^a (+ K^a (G^a)^2), by power iteration on (c - H_loc), c >= lambda_max
// (Gershgorin).  Every iteration lowers the local energy monotonically; the fixed point is the ground state of H_loc, i.e.
// the stationary point the reference's gradient descent (MonteCarlo.jl:303-314) converges to.
template <class R, int NL>
__device__ __forceinline__ void neighbour_sums_d(const DevModel<R, NL>& M, int64_t rbase, size_t ps, int p, R (*S)[NG]) {
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
        for (int a = 0; a < NG; a++) S[l][a] = R(0);
    for (int k = 0; k < M.z; k++) {      // padding slots point at the all-zero site N
        const int j = M.nbr[(size_t)k * M.N + p];
        const int l = NL == 1 ? 0 : M.lab[(size_t)k * M.N + p];
        R t[NG];
        load_T<R>(M.T + rbase + j, ps, t);
#pragma unroll
        for (int ll = 0; ll < NL; ll++)
            if (ll == l) {
#pragma unroll
                for (int a = 0; a < NG; a++) S[ll][a] += t[a];
            }
    }
}

template <class R, int GEN, bool ONSITE, int NL>
__global__ void __launch_bounds__(128) k_optimize_colour(const __grid_constant__ DevModel<R, NL> M, int colour, int iters) {
    constexpr int D = GenOps<GEN>::D;
    const int nc = M.col_off[colour + 1] - M.col_off[colour];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M.nrep * nc) return;
    const int64_t r = idx / nc;
    const int p = M.col_list[M.col_off[colour] + (int)(idx - r * nc)];
    const size_t ps = (size_t)M.nrep * M.Ns;
    const int64_t rbase = r * M.Ns;
    R re[D], im[D], h[NG], Hr[D * D], Hi[D * D];
    load_psi<R, D>(M.psi + rbase + p, ps, re, im);
    {
        R S[NL][NG];
        neighbour_sums_d<R, NL>(M, rbase, ps, p, S);
        field_from_sums<R, NL>(M.cp, S, h);
    }
    GenOps<GEN>::template hloc<R, ONSITE>(h, M.cp.K, Hr, Hi);
    // fill the lower triangle (Hermitian) and find the Gershgorin upper bound
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
        for (int k = 0; k < j; k++) { Hr[j * D + k] = Hr[k * D + j]; Hi[j * D + k] = -Hi[k * D + j]; }
    R c = R(-1e30);
#pragma unroll
    for (int j = 0; j < D; j++) {
        R row = Hr[j * D + j];
#pragma unroll
        for (int k = 0; k < D; k++)
            if (k != j) row += sqrt(Hr[j * D + k] * Hr[j * D + k] + Hi[j * D + k] * Hi[j * D + k]);
        c = row > c ? row : c;
    }
    // A = c - H (positive semi-definite)
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
        for (int k = 0; k < D; k++) {
            Hr[j * D + k] = (j == k ? c : R(0)) - Hr[j * D + k];
            Hi[j * D + k] = -Hi[j * D + k];
        }
    for (int it = 0; it < iters; it++) {
        R nr[D], ni[D], n2 = R(0);
#pragma unroll
        for (int j = 0; j < D; j++) {
            R ar = R(0), ai = R(0);
#pragma unroll
            for (int k = 0; k < D; k++) {
                ar += Hr[j * D + k] * re[k] - Hi[j * D + k] * im[k];
                ai += Hr[j * D + k] * im[k] + Hi[j * D + k] * re[k];
            }
            nr[j] = ar; ni[j] = ai; n2 += ar * ar + ai * ai;
        }
        if (!(n2 > R(0))) break;     // psi orthogonal to everything but the top state of H (or H = c): keep psi
        const R inv = R(1) / sqrt(n2);
#pragma unroll
        for (int j = 0; j < D; j++) { re[j] = nr[j] * inv; im[j] = ni[j] * inv; }
    }
    R T[NG];
    store_psi<R, D>(M.psi + rbase + p, ps, re, im);
    GenOps<GEN>::expect(re, im, T);
    store_T<R>(M.T + rbase + p, ps, T);
    if (ONSITE) {
        GenOps<GEN>::expect_sq(re, im, T);
        store_T<R>(M.Tsq + rbase + p, ps, T);
    }
}

// ------------------------------------------------------------------------------------------------ structure factor
// S^a(k) = |sum_i T_i^a e^{i k.r_i}|^2 / n_rs.  Block = 128 momenta of one replica; FP64 phases and accumulators.
template <class R>
__global__ void __launch_bounds__(128) k_structure_factor(const vec4<R>* __restrict__ Tpl, int N, int Ns, int64_t nrep, const double* __restrict__ ks,
                                                          const double* __restrict__ pos /* permuted order [N][dim] */, int nk, int dim, double inv_nrs,
                                                          double* __restrict__ out /* [R][16][nk] */) {
    const int64_t r = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < nk;
    double kv[3] = {0, 0, 0};
    if (valid) for (int x = 0; x < dim; x++) kv[x] = ks[(size_t)k * dim + x];
    double sr[NG], si[NG];
#pragma unroll
    for (int a = 0; a < NG; a++) { sr[a] = 0.0; si[a] = 0.0; }
    const size_t ps = (size_t)nrep * Ns;
    const vec4<R>* base = Tpl + r * Ns;
    for (int p = 0; p < N; p++) {
        double ph = 0.0;
        for (int x = 0; x < dim; x++) ph += kv[x] * pos[(size_t)p * dim + x];
        double s, c;
        sincos(ph, &s, &c);
        R t[NG];
        load_T<R>(base + p, ps, t);
#pragma unroll
        for (int a = 0; a < NG; a++) { sr[a] = fma((double)t[a], c, sr[a]); si[a] = fma((double)t[a], s, si[a]); }
    }
    if (valid)
#pragma unroll
        for (int a = 0; a < NG; a++) out[((size_t)r * NG + a) * nk + k] = (sr[a] * sr[a] + si[a] * si[a]) * inv_nrs;
}

// ---- structure factor on the allowed momenta of a periodic L1 x L2 cluster as a separable DFT over the cell indices -----------------
// k = (m1 b1 + m2 b2) / L, r = n1 a1 + n2 a2 + delta_b  =>  k.r = 2 pi (m1 n1 / L1 + m2 n2 / L2) + k.delta_b, so for every basis site b
// F_b^a(m1, m2) = sum_{n1, n2} T^a(n1, n2, b) w1^{m1 n1} w2^{m2 n2} gives all L1 L2 momenta of a cell in O(N (L1 + L2)) operations and
// S^a(k) = |sum_b e^{i k.delta_b} F_b^a(m1 mod L1, m2 mod L2)|^2 / n_rs for any allowed k (momenta outside the first cell differ only in the
// basis phases).  FP64 twiddles (host-computed tables) and accumulation, like k_structure_factor.
// pass 1, rows: X[r][b][a][n2][m1] = sum_{n1} T^a(n1, n2, b) w1^{m1 n1};  grid (L2 * nb, replicas)
template <class R>
__global__ void __launch_bounds__(128) k_sf_dft_rows(const vec4<R>* __restrict__ Tpl, int Ns, int64_t nrep, int64_t r0, const int32_t* __restrict__ grid_site /* [nb][L1][L2] */,
                                                     const double2* __restrict__ tw1, int L1, int L2, int nb, double2* __restrict__ X) {
    extern __shared__ __align__(16) unsigned char sf_smem[];
    double* sT = reinterpret_cast<double*>(sf_smem);                  // [L1][16]
    double2* sw = reinterpret_cast<double2*>(sT + (size_t)L1 * NG);   // [L1]
    const int n2 = blockIdx.x % L2, b = blockIdx.x / L2;
    const int64_t r = r0 + blockIdx.y;
    const size_t ps = (size_t)nrep * Ns;
    const vec4<R>* base = Tpl + r * Ns;
    for (int t = threadIdx.x; t < L1 * 4; t += blockDim.x) {
        const int n1 = t >> 2, q = t & 3;
        const vec4<R> v = base[(size_t)q * ps + grid_site[((size_t)b * L1 + n1) * L2 + n2]];
        double* d = sT + (size_t)n1 * NG + 4 * q;
        d[0] = (double)v.x; d[1] = (double)v.y; d[2] = (double)v.z; d[3] = (double)v.w;
    }
    for (int t = threadIdx.x; t < L1; t += blockDim.x) sw[t] = tw1[t];
    __syncthreads();
    for (int m1 = threadIdx.x; m1 < L1; m1 += blockDim.x) {
        double sr[NG], si[NG];
#pragma unroll
        for (int a = 0; a < NG; a++) { sr[a] = 0.0; si[a] = 0.0; }
        int idx = 0;
        for (int n1 = 0; n1 < L1; n1++) {
            const double2 w = sw[idx];
            const double* t = sT + (size_t)n1 * NG;
#pragma unroll
            for (int a = 0; a < NG; a++) { sr[a] = fma(t[a], w.x, sr[a]); si[a] = fma(t[a], w.y, si[a]); }
            idx += m1; if (idx >= L1) idx -= L1;
        }
#pragma unroll
        for (int a = 0; a < NG; a++)
            X[((((size_t)blockIdx.y * nb + b) * NG + a) * L2 + n2) * L1 + m1] = make_double2(sr[a], si[a]);
    }
}
// pass 2, columns: F[r][b][a][m1][m2] = sum_{n2} X[r][b][a][n2][m1] w2^{m2 n2};  grid (ceil(L1 / 8), 16 * nb, replicas), tiles of 8 m1
__global__ void __launch_bounds__(256) k_sf_dft_cols(const double2* __restrict__ X, const double2* __restrict__ tw2, int L1, int L2, int nb, double2* __restrict__ F) {
    extern __shared__ __align__(16) unsigned char sf_smem[];
    double2* sX = reinterpret_cast<double2*>(sf_smem);                // [L2][8]
    double2* sw = sX + (size_t)L2 * 8;                                // [L2]
    const int m1_0 = blockIdx.x * 8;
    const size_t plane = ((size_t)blockIdx.z * nb * NG + blockIdx.y) * (size_t)L1 * L2;      // blockIdx.y = b * 16 + a
    for (int t = threadIdx.x; t < L2 * 8; t += blockDim.x) {
        const int n2 = t >> 3, j = t & 7;
        sX[t] = (m1_0 + j < L1) ? X[plane + (size_t)n2 * L1 + m1_0 + j] : make_double2(0.0, 0.0);
    }
    for (int t = threadIdx.x; t < L2; t += blockDim.x) sw[t] = tw2[t];
    __syncthreads();
    for (int t = threadIdx.x; t < L2 * 8; t += blockDim.x) {
        const int m2 = t >> 3, j = t & 7;
        if (m1_0 + j >= L1) continue;
        double yr = 0.0, yi = 0.0;
        int idx = 0;
        for (int n2 = 0; n2 < L2; n2++) {
            const double2 w = sw[idx], x = sX[n2 * 8 + j];
            yr = fma(x.x, w.x, yr); yr = fma(-x.y, w.y, yr);
            yi = fma(x.x, w.y, yi); yi = fma(x.y, w.x, yi);
            idx += m2; if (idx >= L2) idx -= L2;
        }
        F[plane + (size_t)(m1_0 + j) * L2 + m2] = make_double2(yr, yi);
    }
}
// S^a(k) for the requested momenta: out[r][a][k] = |sum_b phase[k][b] F[r][b][a][m1_k][m2_k]|^2 / n_rs;  grid (ceil(nk / 128), replicas)
// comp_mask != 0: out[r][k] = sum over the components a with bit a set (computeStructureFactor, Observables.jl:101-113)
__global__ void __launch_bounds__(128) k_sf_combine(const double2* __restrict__ F, const int32_t* __restrict__ m, const double2* __restrict__ phase, int nk, int L1, int L2, int nb,
                                                    double inv_nrs, int64_t r0, uint32_t comp_mask, int add, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    const int m1 = m[2 * k], m2 = m[2 * k + 1];
    const size_t at = (size_t)m1 * L2 + m2, plane = (size_t)L1 * L2;
    double total = 0.0;
    for (int a = 0; a < NG; a++) {
        if (comp_mask && !((comp_mask >> a) & 1u)) continue;
        double sr = 0.0, si = 0.0;
        for (int b = 0; b < nb; b++) {
            const double2 f = F[(((size_t)blockIdx.y * nb + b) * NG + a) * plane + at], ph = phase[(size_t)k * nb + b];
            sr += f.x * ph.x - f.y * ph.y;
            si += f.x * ph.y + f.y * ph.x;
        }
        const double v = (sr * sr + si * si) * inv_nrs;
        if (comp_mask) total += v;
        else out[((size_t)(r0 + blockIdx.y) * NG + a) * nk + k] = v;
    }
    if (comp_mask) {
        double* o = out + (size_t)(r0 + blockIdx.y) * nk + k;       // one thread per (replica, momentum): no atomics needed
        *o = add ? *o + total : total;
    }
}

// C[a][project[i][j]] += T_i^a T_j^a  (Observables.jl:22-41); project given in the caller's site order
template <class R>
__global__ void __launch_bounds__(256) k_correlations(const vec4<R>* __restrict__ Tpl, int N, int Ns, int64_t nrep, int64_t r, const int32_t* __restrict__ perm,
                                                      const int32_t* __restrict__ project, int64_t n_rs, double* C) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * N) return;
    const int i = (int)(idx / N), j = (int)(idx - (int64_t)i * N);
    const size_t ps = (size_t)nrep * Ns;
    R ti[NG], tj[NG];
    load_T<R>(Tpl + r * Ns + perm[i], ps, ti);
    load_T<R>(Tpl + r * Ns + perm[j], ps, tj);
    const int32_t m = project[idx];
#pragma unroll
    for (int a = 0; a < NG; a++) atomicAdd(C + (size_t)a * n_rs + m, (double)ti[a] * (double)tj[a]);
}

// mean T over the sites of each sublattice label (ObservablesTgHbn.jl:211-221); one CTA per replica, one pass per label
template <class R>
__global__ void __launch_bounds__(256) k_sublattice_mag(const vec4<R>* __restrict__ Tpl, int N, int Ns, int64_t nrep, const int32_t* __restrict__ label_perm,
                                                        int n_sub, double* __restrict__ out) {
    const int64_t r = blockIdx.x;
    const size_t ps = (size_t)nrep * Ns;
    __shared__ double red[8][NG + 1];
    for (int l = 0; l < n_sub; l++) {
        double acc[NG + 1];
#pragma unroll
        for (int a = 0; a <= NG; a++) acc[a] = 0.0;
        for (int p = threadIdx.x; p < N; p += blockDim.x) {
            if (label_perm[p] != l) continue;
            R t[NG];
            load_T<R>(Tpl + r * Ns + p, ps, t);
#pragma unroll
            for (int a = 0; a < NG; a++) acc[a] += (double)t[a];
            acc[NG] += 1.0;
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int a = 0; a <= NG; a++) {
            double v = acc[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][a] = v;
        }
        __syncthreads();
        if (threadIdx.x <= NG) {
            double v = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w][threadIdx.x];
            red[0][threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x < NG) out[((size_t)r * n_sub + l) * NG + threadIdx.x] = red[0][NG] > 0.0 ? red[0][threadIdx.x] / red[0][NG] : 0.0;
        __syncthreads();
    }
}
// staggered vector chirality (ObservablesTgHbn.jl:224-249); tri in permuted site order
template <class R>
__global__ void __launch_bounds__(256) k_chirality(const vec4<R>* __restrict__ Tpl, int N, int Ns, int64_t nrep, const int32_t* __restrict__ tri,
                                                   double* __restrict__ out) {
    const int64_t r = blockIdx.x;
    const size_t ps = (size_t)nrep * Ns;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
        R t[5][NG];
        load_T<R>(Tpl + r * Ns + p, ps, t[0]);
#pragma unroll
        for (int m = 0; m < 4; m++) load_T<R>(Tpl + r * Ns + tri[4 * p + m], ps, t[m + 1]);
        // (v1 x v2 + v2 x v3 + v3 x v1)_z with the first two entries of each 3-component block starting at `o` (stride `st`)
        auto chi = [&](int i1, int i2, int i3, int o, int st) {
            const double x1 = t[i1][o], y1 = t[i1][o + st], x2 = t[i2][o], y2 = t[i2][o + st], x3 = t[i3][o], y3 = t[i3][o + st];
            return (x1 * y2 - y1 * x2) + (x2 * y3 - y2 * x3) + (x3 * y1 - y3 * x1);
        };
        auto kap = [&](int o, int st) { return chi(0, 1, 2, o, st) - chi(0, 3, 4, o, st); };
        acc[0] += kap(4, 4);                             // components [5, 9, 13] (1-based): sigma
        acc[1] += kap(1, 1);                             // components 2:4: tau
        acc[2] += kap(5, 1) + kap(9, 1) + kap(13, 1);    // 6:8 + 10:12 + 14:16: sigma tau
    }
    __shared__ double red[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w][threadIdx.x];
        out[r * 3 + threadIdx.x] = v * (2.0 / 3.0 / 1.7320508075688772) / (double)N / 2.0;
    }
}

template <int V> using ic = std::integral_constant<int, V>;
template <class R, int NL>
static DevModel<R, NL> make_model_d(const scmc_handle* h) {
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
static int32_t dispatch_d(const scmc_handle* h, F&& f) {
    auto with_nl = [&](auto rt, auto gen, auto ons) -> int32_t {
        if (h->n_labels == 1) return f(rt, gen, ons, ic<1>{});
        return f(rt, gen, ons, ic<MAXL>{});
    };
    auto with_gen = [&](auto rt) -> int32_t {
        switch (h->gen_id) {
            case 1: return with_nl(rt, ic<1>{}, ic<0>{});
            case 2: return h->onsite ? with_nl(rt, ic<2>{}, ic<1>{}) : with_nl(rt, ic<2>{}, ic<0>{});
            case 3: return with_nl(rt, ic<3>{}, ic<0>{});
            case 4: if (set_custom_tables(*h->custom, h, h->device) != cudaSuccess) break; return f(rt, ic<4>{}, ic<1>{}, ic<MAXL>{});
            case 5: if (set_custom_tables(*h->custom, h, h->device) != cudaSuccess) break; return f(rt, ic<5>{}, ic<1>{}, ic<MAXL>{});
        }
        return SCMC_ERR_ARG;
    };
    return h->prec == 64 ? with_gen(double{}) : with_gen(float{});
}
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// Device scratch of one API call: a slot of the handle's grow-only pool (no cudaMalloc / cudaFree per call).  A buffer above 64 MB gives
// ITS OWN slot back when it goes out of scope (never another buffer's: an inner-scope buffer may die while outer ones are still in use).
struct DevBuf {
    scmc_handle* h; int slot; void* p = nullptr;
    DevBuf(scmc_handle* h_, int slot_) : h(h_), slot(slot_) {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p && h->scratch_bytes[slot] > ((size_t)64 << 20)) scratch_release(h, slot); }
    cudaError_t alloc(size_t bytes) { return scratch_get(h, slot, bytes, &p); }
    template <class T> T* as() { return (T*)p; }
};

}  // namespace scmc

using namespace scmc;

extern "C" {

int32_t scmc_run_metropolis(scmc_t* h, const scmc_run_params* p, scmc_run_results* out) {
    SCMC_ARG(h && p && out, "null argument");
    SCMC_ARG(p->T && p->T_i && p->sigma, "scmc_run_metropolis: T, T_i and sigma arrays required");
    SCMC_ARG(p->n_therm >= 0 && p->n_measure >= 0 && p->measurement_rate >= 1 && p->hits >= 1, "scmc_run_metropolis: bad counts");
    SCMC_CK(cudaSetDevice(h->device));
    const int64_t R = h->R, rate = p->measurement_rate;
    const int64_t n_anneal = (int64_t)std::ceil(0.75 * (double)p->n_therm);
    // stop_sweep (0 = run to the end): the call ends after that sweep, so that a caller can write checkpoint files every checkpoint_rate
    // sweeps (MonteCarlo.jl:56-59, 88-91) and continue with sweep0 = stop_sweep; it must not cut a sigma-adaptation / measurement interval
    const int64_t total_all = p->n_therm + p->n_measure;
    SCMC_ARG(p->stop_sweep >= 0 && (p->stop_sweep == 0 || p->stop_sweep >= total_all || p->stop_sweep % p->measurement_rate == 0),
             "scmc_run_metropolis: stop_sweep must be a multiple of measurement_rate");
    const int64_t total = p->stop_sweep > 0 ? std::min<int64_t>(total_all, p->stop_sweep) : total_all;
    const int64_t therm_end = std::min<int64_t>(p->n_therm, total);
    const int64_t meas_start = std::max<int64_t>(p->n_therm + 1, (int64_t)p->sweep0 + 1);
    const int64_t rows_total = total >= meas_start ? total / rate - (meas_start - 1) / rate : 0;
    DevBuf dT(h, 0), dTi(h, 1), dser(h, 2);
    SCMC_CK(dT.alloc(R * 8)); SCMC_CK(dTi.alloc(R * 8));
    SCMC_CK(cudaMemcpyAsync(dT.p, p->T, R * 8, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemcpyAsync(dTi.p, p->T_i, R * 8, cudaMemcpyHostToDevice, h->stream));
    SCMC_TRY(upload_scalars(h, nullptr, p->sigma));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    SCMC_CK(cudaMemsetAsync(h->acc, 0, R * sizeof(unsigned long long), h->stream));
    const size_t row_bytes = (size_t)R * SCMC_NOBS * 8;
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(rows_total, 1), (int64_t)((512ull << 20) / row_bytes)));
    SCMC_CK(dser.alloc(chunk_rows * row_bytes));
    std::vector<double> host_chunk, mean_acc;
    if (out->mean) mean_acc.assign((size_t)R * SCMC_NOBS, 0.0);
    if (!out->series && out->mean) host_chunk.resize((size_t)chunk_rows * R * SCMC_NOBS);
    const double att_per_sweep = (double)p->hits * (double)h->N;
    const unsigned gR = grid_for(R, 128);
    auto set_beta = [&](int64_t sweep) -> int32_t {
        if (h->prec == 64) k_set_beta<double><<<gR, 128, 0, h->stream>>>(R, dT.as<double>(), dTi.as<double>(), sweep, n_anneal, (double*)h->beta);
        else k_set_beta<float><<<gR, 128, 0, h->stream>>>(R, dT.as<double>(), dTi.as<double>(), sweep, n_anneal, (float*)h->beta);
        h->launches++;
        SCMC_CK(cudaGetLastError());
        return SCMC_OK;
    };
    // Device-side driver: when every replica lives in one CTA or one cluster of the psi-family resident kernel, the whole schedule (beta ramp,
    // sigma adaptation, measurement rows) runs inside the persistent kernel, up to 512 sweeps per launch; chains are identical to the
    // host-driven path below, measurement rows agree to rounding (energies re-summed from the resident states).
    const bool psi_fam = uses_psi_gather(h);
    const bool device_side = (psi_fam ? h->res.eligible_psi : h->res.eligible) && (resident_preferred(h) || h->sweep_mode != 0) && (h->sweep_mode == 0 || h->sweep_mode == (psi_fam ? 4 : 2));
    // ---- thermalisation (MonteCarlo.jl:38-66)
    int64_t n_log = 0;
    if (device_side) {
        // energy log entries (one per sigma-adaptation point) are written by the kernel into one device table, copied out at the end
        const int64_t log_base = (int64_t)p->sweep0 / rate + 1;       // g / rate of the first entry
        const int64_t n_log_total = std::max<int64_t>(0, therm_end / rate - (int64_t)p->sweep0 / rate);
        DevBuf dlog(h, 3);
        if (out->therm_energy && n_log_total > 0) SCMC_CK(dlog.alloc((size_t)n_log_total * R * 8));
        for (int64_t sweep = (int64_t)p->sweep0 + 1; sweep <= therm_end;) {
            const int64_t chunk = std::min<int64_t>(512, therm_end - sweep + 1);
            SCMC_TRY(launch_resident_run(h, 1, chunk, p->hits, p->seed, dT.as<double>(), dTi.as<double>(), n_anneal, rate, sweep, log_base,
                                         dlog.p ? dlog.as<double>() : nullptr));
            sweep += chunk;
        }
        if (dlog.p) {
            SCMC_CK(cudaMemcpyAsync(out->therm_energy, dlog.p, (size_t)n_log_total * R * 8, cudaMemcpyDeviceToHost, h->stream));
            SCMC_CK(cudaStreamSynchronize(h->stream));
            n_log = n_log_total;
        }
    }
    for (int64_t sweep = (int64_t)p->sweep0 + 1; !device_side && sweep <= therm_end; sweep++) {
        SCMC_TRY(set_beta(sweep));
        int64_t n_batch = 1;
        if (sweep > n_anneal) {      // ramp finished: beta constant, batch up to the next sigma-adaptation point
            const int64_t next_stop = std::min<int64_t>(therm_end, ((sweep + rate - 1) / rate) * rate);
            n_batch = next_stop - sweep + 1;
        }
        SCMC_TRY(launch_sweeps(h, n_batch, p->hits, p->seed, (uint64_t)(sweep - 1), nullptr));
        sweep += n_batch - 1;
        if (sweep % rate == 0) {
            if (out->therm_energy) {
                SCMC_TRY(launch_measure(h, h->obs));
                SCMC_CK(cudaMemcpy2DAsync(out->therm_energy + n_log * R, 8, h->obs, SCMC_NOBS * 8, 8, R, cudaMemcpyDeviceToHost, h->stream));
                n_log++;
            }
            if (h->prec == 64) k_adapt_sigma<double><<<gR, 128, 0, h->stream>>>(R, h->acc, att_per_sweep * (double)rate, (double*)h->sigma);
            else k_adapt_sigma<float><<<gR, 128, 0, h->stream>>>(R, h->acc, att_per_sweep * (double)rate, (float*)h->sigma);
            h->launches++;
            SCMC_CK(cudaGetLastError());
        }
    }
    // ---- measurement (MonteCarlo.jl:69-99); sigma frozen
    SCMC_TRY(set_beta(0));
    SCMC_CK(cudaMemsetAsync(h->acc, 0, R * sizeof(unsigned long long), h->stream));
    int64_t rows = 0, in_chunk = 0, n_meas_sweeps = 0;
    auto flush = [&]() -> int32_t {
        if (in_chunk == 0) return SCMC_OK;
        double* dst = out->series ? out->series + (size_t)(rows - in_chunk) * R * SCMC_NOBS : (host_chunk.empty() ? nullptr : host_chunk.data());
        if (dst) {
            SCMC_CK(cudaMemcpyAsync(dst, dser.p, (size_t)in_chunk * row_bytes, cudaMemcpyDeviceToHost, h->stream));
            SCMC_CK(cudaStreamSynchronize(h->stream));
            if (out->mean)
                for (int64_t q = 0; q < in_chunk; q++)
                    for (size_t t = 0; t < (size_t)R * SCMC_NOBS; t++) mean_acc[t] += dst[(size_t)q * R * SCMC_NOBS + t];
        }
        in_chunk = 0;
        return SCMC_OK;
    };
    if (device_side) {
        const int64_t row_base0 = (meas_start - 1) / rate + 1;        // g / rate of the first measurement row
        for (int64_t sweep = meas_start; sweep <= total;) {
            // a launch covers whole measurement intervals, at most chunk_rows rows and 512 sweeps
            int64_t last = std::min<int64_t>(total, sweep + 511);
            const int64_t rows_here_max = chunk_rows - in_chunk;
            const int64_t first_row_g = ((sweep + rate - 1) / rate) * rate;
            if (first_row_g <= last) {
                const int64_t n_rows_possible = (last - first_row_g) / rate + 1;
                if (n_rows_possible > rows_here_max) last = first_row_g + (rows_here_max - 1) * rate;
            }
            const int64_t chunk = last - sweep + 1;
            const int64_t n_rows_here = last >= first_row_g ? (last - first_row_g) / rate + 1 : 0;
            // rows of this launch land behind the `in_chunk` rows already waiting in the device buffer
            SCMC_TRY(launch_resident_run(h, 2, chunk, p->hits, p->seed, dT.as<double>(), dTi.as<double>(), n_anneal, rate, sweep,
                                         row_base0 + rows - in_chunk, dser.as<double>()));
            n_meas_sweeps += chunk;
            in_chunk += n_rows_here; rows += n_rows_here;
            sweep = last + 1;
            if (in_chunk == chunk_rows) SCMC_TRY(flush());
        }
    }
    for (int64_t sweep = meas_start; !device_side && sweep <= total; sweep++) {
        // beta and sigma are frozen here: all sweeps up to the next measurement point go out as ONE call (one launch of the
        // replica-resident kernel when it applies)
        const int64_t next_stop = std::min<int64_t>(total, ((sweep + rate - 1) / rate) * rate);
        const int64_t n_batch = next_stop - sweep + 1;
        SCMC_TRY(launch_sweeps(h, n_batch, p->hits, p->seed, (uint64_t)(sweep - 1), nullptr));
        n_meas_sweeps += n_batch;
        sweep = next_stop;
        if (sweep % rate == 0) {
            SCMC_TRY(launch_measure(h, dser.as<double>() + (size_t)in_chunk * R * SCMC_NOBS));
            in_chunk++; rows++;
            if (in_chunk == chunk_rows) SCMC_TRY(flush());
        }
    }
    SCMC_TRY(flush());
    out->n_rows = rows;
    if (out->mean) for (size_t t = 0; t < (size_t)R * SCMC_NOBS; t++) out->mean[t] = rows ? mean_acc[t] / (double)rows : 0.0;
    // results: sigma, acceptance
    std::vector<unsigned long long> acc(R);
    SCMC_CK(cudaMemcpyAsync(acc.data(), h->acc, R * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    std::vector<char> sg(R * h->esz());
    SCMC_CK(cudaMemcpyAsync(sg.data(), h->sigma, sg.size(), cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    for (int64_t r = 0; r < R; r++) {
        p->sigma[r] = h->prec == 64 ? ((double*)sg.data())[r] : (double)((float*)sg.data())[r];
        if (out->acc_rate) out->acc_rate[r] = n_meas_sweeps ? (double)acc[r] / (att_per_sweep * (double)n_meas_sweeps) : 0.0;
    }
    return SCMC_OK;
}

int32_t scmc_anneal(scmc_t* h, const scmc_anneal_params* p, scmc_anneal_results* out) {
    SCMC_ARG(h && p && out, "null argument");
    SCMC_ARG(p->T_i > 0 && p->T_fac > 0 && p->N_max >= 1 && p->N_per_T >= 1 && p->hits >= 1, "scmc_anneal: bad parameters");
    SCMC_CK(cudaSetDevice(h->device));
    const int64_t R = h->R;
    std::vector<double> b0(R, 1.0 / p->T_i), s0(R, p->sigma0);
    SCMC_TRY(upload_scalars(h, b0.data(), s0.data()));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    DevBuf dsat(h, 0), ddone(h, 1), dnact(h, 2);
    SCMC_CK(dsat.alloc(R * 8)); SCMC_CK(ddone.alloc(R * 8)); SCMC_CK(dnact.alloc(sizeof(int)));
    SCMC_CK(cudaMemsetAsync(dsat.p, 0, R * 8, h->stream));
    SCMC_CK(cudaMemsetAsync(ddone.p, 0, R * 8, h->stream));
    SCMC_CK(cudaMemsetAsync(h->acc, 0, R * sizeof(unsigned long long), h->stream));
    SCMC_CK(cudaMemsetAsync(h->active, 1, R, h->stream));
    const double att_per_sweep = (double)p->hits * (double)h->N, min_accepted = (double)p->min_acc_per_site * (double)h->N;
    const unsigned gR = grid_for(R, 128);
    int n_active = (int)std::min<int64_t>(R, 1 << 30);
    const int check_every = 16;
    // energy / beta log (push!(energy, E / N); push!(betas, beta), MonteCarlo.jl:158-159): every replica follows its own schedule on the device,
    // so the series is taken on a fixed grid, every log_rate sweeps, for all replicas at once (stopped replicas repeat their final values)
    const int64_t log_rate = (p->log_rate > 0 && out->energy_log && out->beta_log) ? p->log_rate : 0;
    int64_t n_log = 0;
    std::vector<char> blog(R * h->esz());
    auto log_point = [&]() -> int32_t {
        if (n_log >= out->n_log_cap) return SCMC_OK;
        SCMC_TRY(launch_measure(h, h->obs));
        SCMC_CK(cudaMemcpy2DAsync(out->energy_log + n_log * R, 8, h->obs, SCMC_NOBS * 8, 8, R, cudaMemcpyDeviceToHost, h->stream));
        SCMC_CK(cudaMemcpyAsync(blog.data(), h->beta, blog.size(), cudaMemcpyDeviceToHost, h->stream));
        SCMC_CK(cudaStreamSynchronize(h->stream));
        for (int64_t r = 0; r < R; r++) out->beta_log[n_log * R + r] = h->prec == 64 ? ((double*)blog.data())[r] : (double)((float*)blog.data())[r];
        n_log++;
        return SCMC_OK;
    };
    // Device-side driver: when every replica lives in ONE CTA of the psi-family resident kernel, the whole cooling loop
    // (acceptance window, beta / sigma updates, stop flag) runs inside the persistent kernel, `chunk` sweeps per launch.
    const bool psi_fam = uses_psi_gather(h);
    const bool device_side = (psi_fam ? h->res.eligible_psi : h->res.eligible) && (resident_preferred(h) || h->sweep_mode != 0) && (h->sweep_mode == 0 || h->sweep_mode == (psi_fam ? 4 : 2)) && p->N_max > 1;
    if (device_side) {
        const int64_t n_total = p->N_max - 1;
        for (int64_t launched = 0; launched < n_total && n_active > 0;) {
            int64_t chunk = std::min<int64_t>(512, n_total - launched);
            if (log_rate) chunk = std::min<int64_t>(chunk, log_rate - launched % log_rate);
            SCMC_CK(cudaMemsetAsync(dnact.p, 0, sizeof(int), h->stream));
            SCMC_TRY(launch_resident_anneal(h, chunk, p->hits, p->seed, att_per_sweep, min_accepted, 1.0 / p->T_fac, p->min_accrate, p->sigma_min,
                                            p->N_per_T, n_total, dsat.as<int64_t>(), ddone.as<int64_t>(), dnact.as<int>()));
            SCMC_CK(cudaMemcpyAsync(&n_active, dnact.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            SCMC_CK(cudaStreamSynchronize(h->stream));
            launched += chunk;
            if (log_rate && launched % log_rate == 0) SCMC_TRY(log_point());
        }
    }
    for (int64_t sweep = 1; !device_side && sweep < p->N_max && n_active > 0; sweep++) {
        SCMC_TRY(launch_sweeps(h, 1, p->hits, p->seed, (uint64_t)(sweep - 1), h->active));
        const bool check = (sweep % check_every == 0) || (sweep + 1 >= p->N_max);
        SCMC_CK(cudaMemsetAsync(dnact.p, 0, sizeof(int), h->stream));
        const int last = (sweep + 1 >= p->N_max) ? 1 : 0;
        if (h->prec == 64)
            k_anneal_ctrl<double><<<gR, 128, 0, h->stream>>>(R, h->acc, dsat.as<int64_t>(), ddone.as<int64_t>(), h->active, (double*)h->beta, (double*)h->sigma,
                                                            att_per_sweep, min_accepted, p->N_per_T, last, 1.0 / p->T_fac, p->min_accrate, p->sigma_min,
                                                            dnact.as<int>());
        else
            k_anneal_ctrl<float><<<gR, 128, 0, h->stream>>>(R, h->acc, dsat.as<int64_t>(), ddone.as<int64_t>(), h->active, (float*)h->beta, (float*)h->sigma,
                                                           att_per_sweep, min_accepted, p->N_per_T, last, 1.0 / p->T_fac, p->min_accrate, p->sigma_min,
                                                           dnact.as<int>());
        h->launches++;
        SCMC_CK(cudaGetLastError());
        if (check) {
            SCMC_CK(cudaMemcpyAsync(&n_active, dnact.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            SCMC_CK(cudaStreamSynchronize(h->stream));
        }
        if (log_rate && sweep % log_rate == 0) SCMC_TRY(log_point());
    }
    out->n_log = n_log;
    SCMC_CK(cudaMemsetAsync(h->active, 1, R, h->stream));
    if (out->E) SCMC_TRY(scmc_energy(h, out->E));
    std::vector<char> buf(R * h->esz());
    for (int pass = 0; pass < 2; pass++) {
        double* dst = pass ? out->sigma : out->beta;
        if (!dst) continue;
        SCMC_CK(cudaMemcpyAsync(buf.data(), pass ? h->sigma : h->beta, buf.size(), cudaMemcpyDeviceToHost, h->stream));
        SCMC_CK(cudaStreamSynchronize(h->stream));
        for (int64_t r = 0; r < R; r++) dst[r] = h->prec == 64 ? ((double*)buf.data())[r] : (double)((float*)buf.data())[r];
    }
    if (out->sweeps) {
        SCMC_CK(cudaMemcpyAsync(out->sweeps, ddone.p, R * 8, cudaMemcpyDeviceToHost, h->stream));
        SCMC_CK(cudaStreamSynchronize(h->stream));
    }
    return SCMC_OK;
}

int32_t scmc_optimize(scmc_t* h, int64_t n_sweeps, double* E) {
    SCMC_ARG(h && n_sweeps >= 0, "bad argument");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    for (int64_t s = 0; s < n_sweeps; s++)
        for (int c = 0; c < h->n_col; c++) {
            SCMC_TRY(dispatch_d(h, [&](auto rt, auto gen, auto ons, auto nl) -> int32_t {
                using R = decltype(rt);
                constexpr int GEN = decltype(gen)::value, NL = decltype(nl)::value;
                constexpr bool ONS = decltype(ons)::value != 0;
                const int nc = h->col_off[c + 1] - h->col_off[c];
                if (nc == 0) return SCMC_OK;
                auto M = make_model_d<R, NL>(h);
                k_optimize_colour<R, GEN, ONS, NL><<<grid_for(h->R * nc, 128), 128, 0, h->stream>>>(M, c, 48);
                h->launches++;
                SCMC_CK(cudaGetLastError());
                return SCMC_OK;
            }));
        }
    if (E) return scmc_energy(h, E);
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

int32_t scmc_structure_factor(scmc_t* h, const double* ks, const double* pos, int32_t nk, int32_t dim, double n_rs, double* Sak) {
    SCMC_ARG(h && ks && pos && Sak, "null argument");
    SCMC_ARG(nk >= 1 && dim >= 1 && dim <= 3 && n_rs > 0, "scmc_structure_factor: bad nk / dim / n_rs");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    std::vector<double> ppos((size_t)h->N * dim);
    for (int64_t q = 0; q < h->N; q++)
        for (int x = 0; x < dim; x++) ppos[(size_t)q * dim + x] = pos[(size_t)h->orig[q] * dim + x];
    DevBuf dk(h, 0), dp(h, 1), dout(h, 2);
    SCMC_CK(dk.alloc((size_t)nk * dim * 8)); SCMC_CK(dp.alloc(ppos.size() * 8));
    const size_t out_bytes = (size_t)h->R * 16 * nk * 8;
    SCMC_CK(dout.alloc(out_bytes));
    SCMC_CK(cudaMemcpyAsync(dk.p, ks, (size_t)nk * dim * 8, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemcpyAsync(dp.p, ppos.data(), ppos.size() * 8, cudaMemcpyHostToDevice, h->stream));
    const dim3 grid(grid_for(nk, 128), (unsigned)std::min<int64_t>(h->R, 65535));
    SCMC_ARG(h->R <= 65535, "scmc_structure_factor: more than 65535 replicas per call not supported");
    if (h->prec == 64)
        k_structure_factor<double><<<grid, 128, 0, h->stream>>>((const double4*)h->T, (int)h->N, (int)h->Ns, h->R, dk.as<double>(), dp.as<double>(), nk, dim, 1.0 / n_rs, dout.as<double>());
    else
        k_structure_factor<float><<<grid, 128, 0, h->stream>>>((const float4*)h->T, (int)h->N, (int)h->Ns, h->R, dk.as<double>(), dp.as<double>(), nk, dim, 1.0 / n_rs, dout.as<double>());
    h->launches++;
    SCMC_CK(cudaGetLastError());
    SCMC_CK(cudaMemcpyAsync(Sak, dout.p, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

// shared body of scmc_structure_factor_grid (Sak != nullptr: result copied to the host) and scmc_structure_factor_accumulate (Sak == nullptr:
// the component-summed result is added to the handle's running sum on the device)
static int32_t sf_grid_run(scmc_t* h, int32_t L1, int32_t L2, int32_t n_basis, const int32_t* cell, const int32_t* basis, int32_t nk, const int32_t* m,
                           const double* phase, double n_rs, uint32_t comp_mask, double* Sak) {
    SCMC_ARG(h && cell && basis && m && phase, "null argument");
    SCMC_ARG(L1 >= 1 && L2 >= 1 && n_basis >= 1 && nk >= 1 && n_rs > 0, "scmc_structure_factor_grid: bad sizes");
    SCMC_ARG((int64_t)L1 * L2 * n_basis == h->N, "scmc_structure_factor_grid: L1 * L2 * n_basis must equal the number of sites");
    SCMC_ARG(comp_mask <= 0xFFFFu, "scmc_structure_factor_grid: component mask has bits above 15");
    SCMC_ARG(L1 <= 4096 && L2 <= 4096, "scmc_structure_factor_grid: linear size above 4096 not supported");
    std::vector<int32_t> gs((size_t)h->N, -1);
    for (int64_t i = 0; i < h->N; i++) {
        const int32_t n1 = cell[2 * i], n2 = cell[2 * i + 1], b = basis[i];
        SCMC_ARG(n1 >= 0 && n1 < L1 && n2 >= 0 && n2 < L2 && b >= 0 && b < n_basis, "scmc_structure_factor_grid: cell / basis index out of range");
        int32_t& slot = gs[((size_t)b * L1 + n1) * L2 + n2];
        SCMC_ARG(slot < 0, "scmc_structure_factor_grid: two sites share a (cell, basis) index");
        slot = h->perm[i];
    }
    for (int32_t k = 0; k < nk; k++) SCMC_ARG(m[2 * k] >= 0 && m[2 * k] < L1 && m[2 * k + 1] >= 0 && m[2 * k + 1] < L2, "scmc_structure_factor_grid: momentum index out of range");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    std::vector<double> tw((size_t)2 * (L1 + L2));
    for (int j = 0; j < L1; j++) { tw[2 * j] = std::cos(2.0 * M_PI * j / L1); tw[2 * j + 1] = std::sin(2.0 * M_PI * j / L1); }
    for (int j = 0; j < L2; j++) { tw[2 * (L1 + j)] = std::cos(2.0 * M_PI * j / L2); tw[2 * (L1 + j) + 1] = std::sin(2.0 * M_PI * j / L2); }
    // replicas are processed in batches so that the two FP64 complex work arrays stay below ~1 GiB each
    const size_t per_rep = (size_t)n_basis * 16 * L1 * L2 * sizeof(double2);
    const int64_t rb = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(h->R, 65535), (int64_t)(((size_t)1 << 30) / per_rep)));
    DevBuf dgs(h, 0), dtw(h, 1), dX(h, 2), dF(h, 3), dm(h, 4), dph(h, 5), dout(h, 6);
    SCMC_CK(dgs.alloc(gs.size() * 4)); SCMC_CK(dtw.alloc(tw.size() * 8)); SCMC_CK(dX.alloc(per_rep * rb)); SCMC_CK(dF.alloc(per_rep * rb));
    SCMC_CK(dm.alloc((size_t)nk * 2 * 4)); SCMC_CK(dph.alloc((size_t)nk * n_basis * 16));
    const size_t out_bytes = (size_t)h->R * (comp_mask ? 1 : 16) * nk * 8;
    const bool accumulate = Sak == nullptr;
    if (accumulate) {
        SCMC_ARG(comp_mask != 0, "scmc_structure_factor_accumulate: a component mask is required");
        if (h->sf_acc && (h->sf_acc_nk != nk || h->sf_acc_mask != comp_mask)) {
            set_error("scmc_structure_factor_accumulate: momentum count or component mask differs from the running sum; fetch with reset first");
            return SCMC_ERR_ARG;
        }
        if (!h->sf_acc) {
            SCMC_CK(cudaMalloc(&h->sf_acc, out_bytes));
            SCMC_CK(cudaMemsetAsync(h->sf_acc, 0, out_bytes, h->stream));
            h->sf_acc_nk = nk; h->sf_acc_mask = comp_mask; h->sf_acc_count = 0;
        }
    } else {
        SCMC_CK(dout.alloc(out_bytes));
    }
    double* d_result = accumulate ? h->sf_acc : dout.as<double>();
    SCMC_CK(cudaMemcpyAsync(dgs.p, gs.data(), gs.size() * 4, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemcpyAsync(dtw.p, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemcpyAsync(dm.p, m, (size_t)nk * 2 * 4, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemcpyAsync(dph.p, phase, (size_t)nk * n_basis * 16, cudaMemcpyHostToDevice, h->stream));
    const double2* tw1 = dtw.as<double2>(); const double2* tw2 = tw1 + L1;
    const size_t smem1 = (size_t)L1 * 16 * 8 + (size_t)L1 * 16, smem2 = (size_t)L2 * 8 * 16 + (size_t)L2 * 16;
    SCMC_ARG(smem1 <= 200 * 1024 && smem2 <= 200 * 1024, "scmc_structure_factor_grid: linear size too large for the shared-memory tiles");
    if (smem1 > 48 * 1024) {
        SCMC_CK(cudaFuncSetAttribute(k_sf_dft_rows<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        SCMC_CK(cudaFuncSetAttribute(k_sf_dft_rows<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    }
    if (smem2 > 48 * 1024) SCMC_CK(cudaFuncSetAttribute(k_sf_dft_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    for (int64_t r0 = 0; r0 < h->R; r0 += rb) {
        const int64_t rc = std::min<int64_t>(rb, h->R - r0);
        const dim3 g1((unsigned)(L2 * n_basis), (unsigned)rc);
        if (h->prec == 64) k_sf_dft_rows<double><<<g1, 128, smem1, h->stream>>>((const double4*)h->T, (int)h->Ns, h->R, r0, dgs.as<int32_t>(), tw1, L1, L2, n_basis, dX.as<double2>());
        else k_sf_dft_rows<float><<<g1, 128, smem1, h->stream>>>((const float4*)h->T, (int)h->Ns, h->R, r0, dgs.as<int32_t>(), tw1, L1, L2, n_basis, dX.as<double2>());
        const dim3 g2((unsigned)((L1 + 7) / 8), (unsigned)(16 * n_basis), (unsigned)rc);
        k_sf_dft_cols<<<g2, 256, smem2, h->stream>>>(dX.as<double2>(), tw2, L1, L2, n_basis, dF.as<double2>());
        const dim3 g3(grid_for(nk, 128), (unsigned)rc);
        k_sf_combine<<<g3, 128, 0, h->stream>>>(dF.as<double2>(), dm.as<int32_t>(), dph.as<double2>(), nk, L1, L2, n_basis, 1.0 / n_rs, r0, comp_mask, accumulate ? 1 : 0, d_result);
        h->launches += 3;
        SCMC_CK(cudaGetLastError());
    }
    if (accumulate) h->sf_acc_count++;
    else SCMC_CK(cudaMemcpyAsync(Sak, dout.p, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

int32_t scmc_structure_factor_grid(scmc_t* h, int32_t L1, int32_t L2, int32_t n_basis, const int32_t* cell, const int32_t* basis, int32_t nk, const int32_t* m,
                                   const double* phase, double n_rs, uint32_t comp_mask, double* Sak) {
    SCMC_ARG(Sak != nullptr, "null argument");
    return sf_grid_run(h, L1, L2, n_basis, cell, basis, nk, m, phase, n_rs, comp_mask, Sak);
}

int32_t scmc_structure_factor_accumulate(scmc_t* h, int32_t L1, int32_t L2, int32_t n_basis, const int32_t* cell, const int32_t* basis, int32_t nk, const int32_t* m,
                                         const double* phase, double n_rs, uint32_t comp_mask) {
    return sf_grid_run(h, L1, L2, n_basis, cell, basis, nk, m, phase, n_rs, comp_mask, nullptr);
}

int32_t scmc_structure_factor_fetch(scmc_t* h, double* Sk, int64_t* count, int32_t reset) {
    SCMC_ARG(h && count, "null argument");
    *count = h->sf_acc_count;
    if (!h->sf_acc) return SCMC_OK;
    SCMC_CK(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->R * h->sf_acc_nk * 8;
    if (Sk && h->sf_acc_count > 0) SCMC_CK(cudaMemcpyAsync(Sk, h->sf_acc, bytes, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    if (reset) {
        SCMC_CK(cudaFree(h->sf_acc));
        h->sf_acc = nullptr; h->sf_acc_nk = 0; h->sf_acc_count = 0; h->sf_acc_mask = 0;
    }
    return SCMC_OK;
}

int32_t scmc_correlations(scmc_t* h, int64_t r, const int32_t* project, int64_t n_rs, double* C) {
    SCMC_ARG(h && project && C, "null argument");
    SCMC_ARG(r >= 0 && r < h->R && n_rs >= 1, "replica out of range / bad n_rs");
    SCMC_ARG(h->N <= 46340, "scmc_correlations: N too large for the all-pairs reference algorithm");
    for (int64_t t = 0; t < h->N * h->N; t++) SCMC_ARG(project[t] >= 0 && project[t] < n_rs, "scmc_correlations: project index out of range");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    DevBuf dproj(h, 0), dperm(h, 1), dC(h, 2);
    SCMC_CK(dproj.alloc((size_t)h->N * h->N * 4)); SCMC_CK(dperm.alloc(h->N * 4)); SCMC_CK(dC.alloc((size_t)16 * n_rs * 8));
    SCMC_CK(cudaMemcpyAsync(dproj.p, project, (size_t)h->N * h->N * 4, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemcpyAsync(dperm.p, h->perm.data(), h->N * 4, cudaMemcpyHostToDevice, h->stream));
    SCMC_CK(cudaMemsetAsync(dC.p, 0, (size_t)16 * n_rs * 8, h->stream));
    if (h->prec == 64)
        k_correlations<double><<<grid_for(h->N * h->N, 256), 256, 0, h->stream>>>((const double4*)h->T, (int)h->N, (int)h->Ns, h->R, r, dperm.as<int32_t>(), dproj.as<int32_t>(), n_rs, dC.as<double>());
    else
        k_correlations<float><<<grid_for(h->N * h->N, 256), 256, 0, h->stream>>>((const float4*)h->T, (int)h->N, (int)h->Ns, h->R, r, dperm.as<int32_t>(), dproj.as<int32_t>(), n_rs, dC.as<double>());
    h->launches++;
    SCMC_CK(cudaGetLastError());
    SCMC_CK(cudaMemcpyAsync(C, dC.p, (size_t)16 * n_rs * 8, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

int32_t scmc_sublattice_magnetization(scmc_t* h, const int32_t* site_label, int32_t n_sub, double* out) {
    SCMC_ARG(h && site_label && out, "null argument");
    SCMC_ARG(n_sub >= 1 && n_sub <= 64, "scmc_sublattice_magnetization: 1..64 labels");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    std::vector<int32_t> lp(h->N);
    for (int64_t q = 0; q < h->N; q++) {
        lp[q] = site_label[h->orig[q]];
        SCMC_ARG(lp[q] >= 0 && lp[q] < n_sub, "scmc_sublattice_magnetization: label out of range");
    }
    DevBuf dl(h, 0), dout(h, 1);
    const size_t ob = (size_t)h->R * n_sub * 16 * 8;
    SCMC_CK(dl.alloc(h->N * 4)); SCMC_CK(dout.alloc(ob));
    SCMC_CK(cudaMemcpyAsync(dl.p, lp.data(), h->N * 4, cudaMemcpyHostToDevice, h->stream));
    if (h->prec == 64) k_sublattice_mag<double><<<(unsigned)h->R, 256, 0, h->stream>>>((const double4*)h->T, (int)h->N, (int)h->Ns, h->R, dl.as<int32_t>(), n_sub, dout.as<double>());
    else k_sublattice_mag<float><<<(unsigned)h->R, 256, 0, h->stream>>>((const float4*)h->T, (int)h->N, (int)h->Ns, h->R, dl.as<int32_t>(), n_sub, dout.as<double>());
    h->launches++;
    SCMC_CK(cudaGetLastError());
    SCMC_CK(cudaMemcpyAsync(out, dout.p, ob, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

int32_t scmc_chirality(scmc_t* h, const int32_t* tri, double* out) {
    SCMC_ARG(h && tri && out, "null argument");
    SCMC_CK(cudaSetDevice(h->device));
    SCMC_TRY(ensure_T(h));
    std::vector<int32_t> tp((size_t)h->N * 4);
    for (int64_t q = 0; q < h->N; q++)
        for (int m = 0; m < 4; m++) {
            const int32_t j = tri[(size_t)h->orig[q] * 4 + m];
            SCMC_ARG(j >= 0 && j < h->N, "scmc_chirality: site index out of range");
            tp[(size_t)q * 4 + m] = h->perm[j];
        }
    DevBuf dt(h, 0), dout(h, 1);
    SCMC_CK(dt.alloc(tp.size() * 4)); SCMC_CK(dout.alloc((size_t)h->R * 3 * 8));
    SCMC_CK(cudaMemcpyAsync(dt.p, tp.data(), tp.size() * 4, cudaMemcpyHostToDevice, h->stream));
    if (h->prec == 64) k_chirality<double><<<(unsigned)h->R, 256, 0, h->stream>>>((const double4*)h->T, (int)h->N, (int)h->Ns, h->R, dt.as<int32_t>(), dout.as<double>());
    else k_chirality<float><<<(unsigned)h->R, 256, 0, h->stream>>>((const float4*)h->T, (int)h->N, (int)h->Ns, h->R, dt.as<int32_t>(), dout.as<double>());
    h->launches++;
    SCMC_CK(cudaGetLastError());
    SCMC_CK(cudaMemcpyAsync(out, dout.p, (size_t)h->R * 3 * 8, cudaMemcpyDeviceToHost, h->stream));
    SCMC_CK(cudaStreamSynchronize(h->stream));
    return SCMC_OK;
}

}  // extern "C"
