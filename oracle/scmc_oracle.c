/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * --impl reference legs may use it.
 *
 * Plain-C, FP64, single-threaded restatement of the reference's Metropolis /
 * annealing hot path, keeping the reference's algorithmic shape on purpose
 * (array of small per-site vectors, dense d x d x 16 generator contraction evaluated
 * twice per update, one dense 16x16 mat-vec per neighbour, 2N random-site updates
 * per sweep).  Each function cites the reference lines it follows
 * (paths relative to /root/reference/).
 *
 * Parity status: the reference ships no golden vectors (test/runtests.jl:4-6 is an
 * empty test set) and cannot be run here (no Julia).  The oracle is therefore pinned
 * by (i) the analytic identities of SURVEY.md section 4 and (ii) the recorded
 * notebook outputs (doc/examples.ipynb cells 13, 25), see tests/test_oracle_*.py.
 * Random streams of the reference (Julia Random + VectorizedRNG) are not
 * reproducible outside Julia: "parity unpinned" for chains, statistical only.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NG 16

typedef struct {
    int N, d, z, n_labels;
    int *nbr;        /* [N][z] 0-based neighbour or -1          (cfg.interactionSites,  Configuration.jl:10) */
    int *lab;        /* [N][z] 0-based label                    (cfg.interactionLabels, Configuration.jl:11) */
    double *J;       /* [n_labels][16][16], J[l][a][b] = M[a,b] (cfg.interactions,      Configuration.jl:12) */
    double K[NG];    /* on-site couplings                       (cfg.onsiteInteraction, Configuration.jl:13) */
    double *gre, *gim, *qre, *qim; /* [16][d][d] generators and their squares (Configuration.jl:15-16) */
    double *re, *im; /* [N][d] site states                      (cfg.state,             Configuration.jl:17) */
    double *T, *Tsq; /* [N][16] cached expectations             (Configuration.jl:18-19) */
    double dT[NG], newT[NG], newTsq[NG]; /* scratch (Configuration.jl:20-22) */
    double *nre, *nim;                   /* proposed state      (Configuration.jl:23) */
    uint64_t s[4];   /* xoshiro256++ state for the reference-shaped random-site chain */
    int have_spare; double spare;
    double nscale;   /* standard deviation of the proposal normals (1 = N(0,1) as randn! documents; experiments: orc_set_normal_scale) */
    double *sigma_log; long sigma_log_cap; /* optional: cone width after every adaptation of orc_run_metropolis */
} orc_t;

/* ------------------------------------------------------------------ RNG (oracle's own chain) */
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static uint64_t xnext(orc_t *o) {
    uint64_t *s = o->s, r = rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static double urand(orc_t *o) { return (double)(xnext(o) >> 11) * (1.0 / 9007199254740992.0); }
static double nrand(orc_t *o) {
    if (o->have_spare) { o->have_spare = 0; return o->spare; }
    double u1 = 1.0 - urand(o), u2 = urand(o);
    double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586476925 * u2;
    o->spare = o->nscale * r * sin(a); o->have_spare = 1;
    return o->nscale * r * cos(a);
}
/* experiments on the reference's sigma convention (tests/test_oracle_physics.py): normals of standard deviation s; sigma trace */
void orc_set_normal_scale(orc_t *o, double s) { o->nscale = s; o->have_spare = 0; }
void orc_set_sigma_log(orc_t *o, double *buf, long cap) { o->sigma_log = buf; o->sigma_log_cap = cap; }
void orc_seed(orc_t *o, uint64_t seed) {
    for (int i = 0; i < 4; i++) o->s[i] = splitmix(&seed);
    o->have_spare = 0;
}

/* ------------------------------------------------------------------ expectation values */
/* computeSpinExpectation! (src/Configuration.jl:166-194): T^a = Re sum_jk conj(psi_j) G^a_jk psi_k, dense. */
void orc_spin_expectation(int d, const double *re, const double *im, const double *gre, const double *gim, double *out) {
    for (int a = 0; a < NG; a++) {
        double val = 0.0;
        for (int k = 0; k < d; k++) {
            double rk = re[k], ik = im[k];
            for (int j = 0; j < d; j++) {
                double gr = gre[(a * d + j) * d + k], gi = gim[(a * d + j) * d + k];
                val += (re[j] * gr + im[j] * gi) * rk - (re[j] * gi - im[j] * gr) * ik;
            }
        }
        out[a] = val;
    }
}

/* exchangeEnergy (src/Configuration.jl:199-218): T_i^T M T_j */
static double exchange_energy(const double *Ti, const double *M, const double *Tj) {
    double val = 0.0;
    for (int k = 0; k < NG; k++) {
        double vk = 0.0;
        for (int l = 0; l < NG; l++) vk += M[k * NG + l] * Tj[l];
        val += Ti[k] * vk;
    }
    return val;
}
/* onsiteEnergy (src/Configuration.jl:221-233) */
static double onsite_energy(const double *Tsq, const double *K) {
    double val = 0.0;
    for (int k = 0; k < NG; k++) val += Tsq[k] * K[k];
    return val;
}

static void refresh_site(orc_t *o, int i) {
    orc_spin_expectation(o->d, o->re + i * o->d, o->im + i * o->d, o->gre, o->gim, o->T + i * NG);
    orc_spin_expectation(o->d, o->re + i * o->d, o->im + i * o->d, o->qre, o->qim, o->Tsq + i * NG);
}

/* ------------------------------------------------------------------ construction (Configuration.jl:27-77) */
orc_t *orc_create(int N, int d, int z, int n_labels, const int *nbr, const int *lab, const double *J,
                  const double *K, const double *gre, const double *gim, const double *qre, const double *qim) {
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->N = N; o->d = d; o->z = z; o->n_labels = n_labels;
#define DUP(dst, src, n, T) dst = (T *)malloc(sizeof(T) * (size_t)(n)); memcpy(dst, src, sizeof(T) * (size_t)(n))
    DUP(o->nbr, nbr, N * z, int); DUP(o->lab, lab, N * z, int);
    DUP(o->J, J, n_labels * NG * NG, double);
    DUP(o->gre, gre, NG * d * d, double); DUP(o->gim, gim, NG * d * d, double);
    DUP(o->qre, qre, NG * d * d, double); DUP(o->qim, qim, NG * d * d, double);
    memcpy(o->K, K, sizeof(double) * NG);
    o->re = (double *)calloc((size_t)N * d, sizeof(double)); o->im = (double *)calloc((size_t)N * d, sizeof(double));
    o->T = (double *)calloc((size_t)N * NG, sizeof(double)); o->Tsq = (double *)calloc((size_t)N * NG, sizeof(double));
    o->nre = (double *)calloc(d, sizeof(double)); o->nim = (double *)calloc(d, sizeof(double));
    o->nscale = 1.0; o->sigma_log = NULL; o->sigma_log_cap = 0;
    orc_seed(o, 1);
    return o;
}
void orc_destroy(orc_t *o) {
    if (!o) return;
    free(o->nbr); free(o->lab); free(o->J); free(o->gre); free(o->gim); free(o->qre); free(o->qim);
    free(o->re); free(o->im); free(o->T); free(o->Tsq); free(o->nre); free(o->nim); free(o);
}
void orc_set_state(orc_t *o, const double *re, const double *im) {
    memcpy(o->re, re, sizeof(double) * (size_t)o->N * o->d); memcpy(o->im, im, sizeof(double) * (size_t)o->N * o->d);
    for (int i = 0; i < o->N; i++) refresh_site(o, i);
}
void orc_get_state(const orc_t *o, double *re, double *im) {
    memcpy(re, o->re, sizeof(double) * (size_t)o->N * o->d); memcpy(im, o->im, sizeof(double) * (size_t)o->N * o->d);
}
void orc_get_T(const orc_t *o, double *T, double *Tsq) {
    memcpy(T, o->T, sizeof(double) * (size_t)o->N * NG); if (Tsq) memcpy(Tsq, o->Tsq, sizeof(double) * (size_t)o->N * NG);
}
/* getRandomState (src/MonteCarlo.jl:244-246): normalised complex Gaussian vector, for every site (Configuration.jl:62) */
void orc_randomize(orc_t *o) {
    for (int i = 0; i < o->N; i++) {
        double n2 = 0.0;
        for (int k = 0; k < o->d; k++) {
            double a = nrand(o), b = nrand(o);
            o->re[i * o->d + k] = a; o->im[i * o->d + k] = b; n2 += a * a + b * b;
        }
        double inv = 1.0 / sqrt(n2);
        for (int k = 0; k < o->d; k++) { o->re[i * o->d + k] *= inv; o->im[i * o->d + k] *= inv; }
        refresh_site(o, i);
    }
}

/* ------------------------------------------------------------------ energies */
/* getEnergy (src/Configuration.jl:236-253) */
double orc_energy(const orc_t *o) {
    double E = 0.0;
    for (int i = 0; i < o->N; i++) {
        E += onsite_energy(o->Tsq + i * NG, o->K);
        for (int k = 0; k < o->z; k++) {
            int j = o->nbr[i * o->z + k];
            if (j < 0) continue;
            E += exchange_energy(o->T + i * NG, o->J + o->lab[i * o->z + k] * NG * NG, o->T + j * NG) / 2;
        }
    }
    return E;
}
/* getEnergyDifference (src/Configuration.jl:256-275 / :278-298) for proposed newT, newTsq at site i */
double orc_energy_difference(orc_t *o, int i, const double *newT, const double *newTsq) {
    double dE = 0.0;
    dE += onsite_energy(newTsq, o->K) - onsite_energy(o->Tsq + i * NG, o->K);
    for (int a = 0; a < NG; a++) o->dT[a] = newT[a] - o->T[i * NG + a];
    for (int k = 0; k < o->z; k++) {
        int j = o->nbr[i * o->z + k];
        if (j < 0) continue;
        dE += exchange_energy(o->dT, o->J + o->lab[i * o->z + k] * NG * NG, o->T + j * NG);
    }
    return dE;
}
/* parity probe: h_i^a = sum_j sum_b J_label(i,j)^{ab} T_j^b (the vector dT is contracted with in Configuration.jl:292-296) */
void orc_local_field(const orc_t *o, int i, double *h) {
    for (int a = 0; a < NG; a++) h[a] = 0.0;
    for (int k = 0; k < o->z; k++) {
        int j = o->nbr[i * o->z + k];
        if (j < 0) continue;
        const double *M = o->J + o->lab[i * o->z + k] * NG * NG;
        for (int a = 0; a < NG; a++) {
            double v = 0.0;
            for (int b = 0; b < NG; b++) v += M[a * NG + b] * o->T[j * NG + b];
            h[a] += v;
        }
    }
}

/* ------------------------------------------------------------------ local update with injected randoms */
/* getRandomState! (src/MonteCarlo.jl:249-260) with xi = (d real-part normals, d imaginary-part normals) */
static void propose(orc_t *o, int i, const double *xi, double sigma) {
    int d = o->d; double n2 = 0.0;
    for (int k = 0; k < d; k++) {
        o->nre[k] = xi[k] * sigma + o->re[i * d + k];
        o->nim[k] = xi[d + k] * sigma + o->im[i * d + k];
    }
    for (int k = 0; k < d; k++) n2 += o->nre[k] * o->nre[k] + o->nim[k] * o->nim[k];
    double inv = 1.0 / sqrt(n2);
    for (int k = 0; k < d; k++) { o->nre[k] *= inv; o->nim[k] *= inv; }
}
/* localUpdate! (src/MonteCarlo.jl:263-286).  Returns 1 if accepted; *dE_out gets the energy difference. */
int orc_local_update(orc_t *o, int i, const double *xi, double u, double beta, double sigma, double *dE_out) {
    int d = o->d;
    propose(o, i, xi, sigma);
    orc_spin_expectation(d, o->nre, o->nim, o->gre, o->gim, o->newT);
    orc_spin_expectation(d, o->nre, o->nim, o->qre, o->qim, o->newTsq);
    double dE = orc_energy_difference(o, i, o->newT, o->newTsq);
    if (dE_out) *dE_out = dE;
    double p = exp(-beta * dE);
    if (u < p) {
        memcpy(o->re + i * d, o->nre, sizeof(double) * d); memcpy(o->im + i * d, o->nim, sizeof(double) * d);
        memcpy(o->T + i * NG, o->newT, sizeof(double) * NG); memcpy(o->Tsq + i * NG, o->newTsq, sizeof(double) * NG);
        return 1;
    }
    return 0;
}
/* A prescribed sequence of site updates with injected randoms: xi [n][2d], u [n]; writes accept [n], dE [n]. */
long orc_updates_injected(orc_t *o, long n, const int *sites, const double *xi, const double *u, const double *beta,
                          double sigma, unsigned char *accept, double *dE) {
    long acc = 0;
    for (long t = 0; t < n; t++) {
        double de; int a = orc_local_update(o, sites[t], xi + t * 2 * o->d, u[t], beta[0], sigma, &de);
        if (accept) accept[t] = (unsigned char)a;
        if (dE) dE[t] = de;
        acc += a;
    }
    return acc;
}

/* ------------------------------------------------------------------ reference-shaped sweep: 2N random sites */
/* inner loops of run! / runAnnealing! (src/MonteCarlo.jl:40-44, 76-80, 152-156) */
long orc_sweeps_random(orc_t *o, long n_sweeps, double beta, double sigma, double *E_inout) {
    long acc = 0; double xi[16]; double E = E_inout ? *E_inout : 0.0;
    for (long s = 0; s < n_sweeps; s++)
        for (long t = 0; t < 2L * o->N; t++) {
            int i = (int)(urand(o) * o->N); if (i >= o->N) i = o->N - 1;
            for (int k = 0; k < 2 * o->d; k++) xi[k] = nrand(o);
            double de; int a = orc_local_update(o, i, xi, urand(o), beta, sigma, &de);
            if (a) { E += de; acc++; }
        }
    if (E_inout) *E_inout = E;
    return acc;
}

/* ------------------------------------------------------------------ observables */
/* getMagnetization (src/Observables/Observables.jl:16-18) */
void orc_magnetization(const orc_t *o, double *m) {
    for (int a = 0; a < NG; a++) m[a] = 0.0;
    for (int i = 0; i < o->N; i++) for (int a = 0; a < NG; a++) m[a] += o->T[i * NG + a];
    for (int a = 0; a < NG; a++) m[a] /= o->N;
}
/* measure! scalars (src/Observables/ObservablesGeneric.jl:32-46): out = [E/N, E^2/N^2, |m_sigma|, |m_tau|, |m_sigmatau|, m[16]] */
void orc_measure(const orc_t *o, double E, double *out) {
    double m[NG]; orc_magnetization(o, m);
    out[0] = E / o->N; out[1] = E * E / ((double)o->N * o->N);
    out[2] = sqrt(m[4] * m[4] + m[8] * m[8] + m[12] * m[12]);
    out[3] = sqrt(m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
    double s = 0.0; static const int st[9] = {5, 6, 7, 9, 10, 11, 13, 14, 15};
    for (int t = 0; t < 9; t++) s += m[st[t]] * m[st[t]];
    out[4] = sqrt(s);
    for (int a = 0; a < NG; a++) out[5 + a] = m[a];
}
/* getCorrelations (src/Observables/Observables.jl:22-41): C[a][idx] += T_i^a T_j^a, idx = project[i][j] (0-based), C is [16][n_rs] */
void orc_correlations(const orc_t *o, const int *project, int n_rs, double *C) {
    memset(C, 0, sizeof(double) * (size_t)NG * n_rs);
    for (int i = 0; i < o->N; i++) for (int j = 0; j < o->N; j++) {
        int idx = project[(size_t)i * o->N + j];
        for (int a = 0; a < NG; a++) C[(size_t)a * n_rs + idx] += o->T[i * NG + a] * o->T[j * NG + a];
    }
}
/* computeStructureFactorVec (src/Observables/Observables.jl:116-128): sf[a][k] = (1/n_rs) sum_n cos(k.r_n) C[a][n] */
void orc_structure_factor_vec(const double *C, const double *rs, int n_rs, const double *ks, int nk, int dim, double *sf) {
    for (int k = 0; k < nk; k++) {
        double z[NG]; for (int a = 0; a < NG; a++) z[a] = 0.0;
        for (int n = 0; n < n_rs; n++) {
            double ph = 0.0; for (int x = 0; x < dim; x++) ph += ks[k * dim + x] * rs[n * dim + x];
            double c = cos(ph);
            for (int a = 0; a < NG; a++) z[a] += c * C[(size_t)a * n_rs + n];
        }
        for (int a = 0; a < NG; a++) sf[(size_t)a * nk + k] = z[a] / n_rs;
    }
}

/* ------------------------------------------------------------------ run! (src/MonteCarlo.jl:3-108), no file I/O */
/* series: [n_meas][21] rows as orc_measure; returns number of rows written.  sigma_io in/out.  energy_log/beta_log: optional
 * per-measurement-point logs during thermalisation+measurement (push! at :48-49, :84-85), length >= (N_therm+N_measure)/rate. */
long orc_run_metropolis(orc_t *o, double T, double T_i, long N_therm, long N_measure, long measurement_rate,
                        double *sigma_io, double *series, double *energy_log, double *beta_log, double *acc_rate_meas) {
    long N_anneal = (long)ceil(0.75 * (double)N_therm), rows = 0, nlog = 0;
    double sigma = *sigma_io, E = orc_energy(o);
    long att = 0, acc = 0;
    for (long sweep = 1; sweep <= N_therm; sweep++) {
        double Tn = sweep <= N_anneal ? T_i * pow(T / T_i, (double)(sweep - 1) / (double)(N_anneal - 1)) : T;
        double beta = 1.0 / Tn;
        acc += orc_sweeps_random(o, 1, beta, sigma, &E); att += 2L * o->N;
        if (sweep % measurement_rate == 0) {
            if (energy_log) { energy_log[nlog] = E / o->N; beta_log[nlog] = beta; } nlog++;
            double R = (double)acc / (double)att;
            sigma = fmin(sigma * 0.5 / (1.0 - R), 60.0);
            if (o->sigma_log && nlog - 1 < o->sigma_log_cap) o->sigma_log[nlog - 1] = sigma;
            att = 0; acc = 0;
        }
    }
    att = 0; acc = 0;
    double beta = 1.0 / T;
    for (long sweep = N_therm + 1; sweep <= N_therm + N_measure; sweep++) {
        acc += orc_sweeps_random(o, 1, beta, sigma, &E); att += 2L * o->N;
        if (sweep % measurement_rate == 0) {
            orc_measure(o, E, series + rows * 21); rows++;
            if (energy_log) { energy_log[nlog] = E / o->N; beta_log[nlog] = beta; } nlog++;
        }
    }
    if (acc_rate_meas) *acc_rate_meas = att ? (double)acc / (double)att : 0.0;
    *sigma_io = sigma;
    return rows;
}

/* ------------------------------------------------------------------ localOptimization! (src/MonteCarlo.jl:289-329) */
/* F(x) = dE(i; psi = x) on the sphere S^{2d-1}; gradient descent with Armijo backtracking (initial step 1, contraction 0.99,
 * sufficient decrease 0.5), exponential retraction, stop after 200 iterations or |grad| < 1e-3.  The reference differentiates
 * F by finite differences through Manopt; here the (identical up to O(h^2)) analytic Riemannian gradient is used. */
static double loc_F(orc_t *o, int i, const double *x, double *T, double *Tsq) {
    int d = o->d;
    orc_spin_expectation(d, x, x + d, o->gre, o->gim, T);
    orc_spin_expectation(d, x, x + d, o->qre, o->qim, Tsq);
    return orc_energy_difference(o, i, T, Tsq);
}
double orc_local_optimization(orc_t *o, int i) {
    int d = o->d, n = 2 * d; double x[16], g[16], y[16], h[NG], T[NG], Tsq[NG];
    for (int k = 0; k < d; k++) { x[k] = o->re[i * d + k]; x[d + k] = o->im[i * d + k]; }
    orc_local_field(o, i, h);
    double Fx = loc_F(o, i, x, T, Tsq);
    for (int it = 0; it < 200; it++) {
        /* Euclidean gradient of <psi|H_loc|psi>, H_loc = sum_a h^a G^a + K^a (G^a)^2:  2 H psi ; then project on tangent space */
        double hr[8], hi[8];
        for (int j = 0; j < d; j++) {
            double ar = 0, ai = 0;
            for (int k = 0; k < d; k++) {
                double mr = 0, mi = 0;
                for (int a = 0; a < NG; a++) {
                    mr += h[a] * o->gre[(a * d + j) * d + k] + o->K[a] * o->qre[(a * d + j) * d + k];
                    mi += h[a] * o->gim[(a * d + j) * d + k] + o->K[a] * o->qim[(a * d + j) * d + k];
                }
                ar += mr * x[k] - mi * x[d + k]; ai += mr * x[d + k] + mi * x[k];
            }
            hr[j] = ar; hi[j] = ai;
        }
        double dot = 0.0, gn = 0.0;
        for (int k = 0; k < d; k++) { g[k] = 2 * hr[k]; g[d + k] = 2 * hi[k]; }
        for (int k = 0; k < n; k++) dot += g[k] * x[k];
        for (int k = 0; k < n; k++) { g[k] -= dot * x[k]; gn += g[k] * g[k]; }
        gn = sqrt(gn);
        if (gn < 1e-3) break;
        double step = 1.0; int ok = 0; double Fy = Fx;
        for (int ls = 0; ls < 2000; ls++) {
            double c = cos(step * gn), s = sin(step * gn);
            for (int k = 0; k < n; k++) y[k] = c * x[k] - s * g[k] / gn;   /* exp map along -grad */
            Fy = loc_F(o, i, y, T, Tsq);
            if (Fy <= Fx - 0.5 * step * gn * gn) { ok = 1; break; }
            step *= 0.99;
        }
        if (!ok) break;
        memcpy(x, y, sizeof(double) * n); Fx = Fy;
    }
    double nn = 0.0; for (int k = 0; k < n; k++) nn += x[k] * x[k];
    nn = 1.0 / sqrt(nn);
    for (int k = 0; k < d; k++) { o->re[i * d + k] = x[k] * nn; o->im[i * d + k] = x[d + k] * nn; }
    double Told[NG], Tsqold[NG];
    memcpy(Told, o->T + i * NG, sizeof(Told)); memcpy(Tsqold, o->Tsq + i * NG, sizeof(Tsqold));
    orc_spin_expectation(d, o->re + i * d, o->im + i * d, o->gre, o->gim, T);
    orc_spin_expectation(d, o->re + i * d, o->im + i * d, o->qre, o->qim, Tsq);
    double dE = orc_energy_difference(o, i, T, Tsq);
    memcpy(o->T + i * NG, T, sizeof(T)); memcpy(o->Tsq + i * NG, Tsq, sizeof(Tsq));
    return dE;
}

/* ------------------------------------------------------------------ runAnnealing! (src/MonteCarlo.jl:112-238), no file I/O */
/* Returns the number of annealing sweeps done; *E_out final energy; trace (optional) gets E/N per sweep (capacity trace_cap). */
static long run_annealing_traced(orc_t *o, double T_i, double T_fac, long N_max, long N_o, long N_per_T, long min_acc_per_site,
                                 double min_accrate, double sigma_min, double *sigma_io, double *E_out, double *beta_out,
                                 double *trace, double *beta_trace, double *rate_trace, long trace_cap) {
    double E = orc_energy(o), beta = 1.0 / T_i, beta_fac = 1.0 / T_fac, sigma = *sigma_io;
    long min_accepted = min_acc_per_site * o->N, sweep = 1, nt = 0;
    while (sweep < N_max) {
        long att = 0, acc = 0, sweep_at_T = 0;
        while (acc < min_accepted && sweep_at_T < N_per_T && sweep < N_max) {
            acc += orc_sweeps_random(o, 1, beta, sigma, &E); att += 2L * o->N;
            if (trace && nt < trace_cap) trace[nt] = E / o->N;
            if (beta_trace && nt < trace_cap) beta_trace[nt] = beta;                       /* push!(βs, β), MonteCarlo.jl:159 */
            if (rate_trace && nt < trace_cap) rate_trace[nt] = (double)acc / (double)att;   /* R_cp of the checkpoint line, MonteCarlo.jl:162 */
            nt++;
            sweep_at_T++; sweep++;
        }
        double R = (double)acc / (double)att;
        if (R > min_accrate) {
            beta *= beta_fac;
            sigma = fmax(fmin(sigma * 0.5 / (1.0 - R), 60.0), sigma_min);
        } else break;
    }
    for (long oi = 0; oi < N_o; oi++) {
        for (long t = 0; t < 2L * o->N; t++) {
            int i = (int)(urand(o) * o->N); if (i >= o->N) i = o->N - 1;
            E += orc_local_optimization(o, i);
        }
        if (trace && nt < trace_cap) trace[nt] = E / o->N; nt++;
        sweep++;
    }
    *sigma_io = sigma; if (E_out) *E_out = E; if (beta_out) *beta_out = beta;
    return sweep;
}
long orc_run_annealing(orc_t *o, double T_i, double T_fac, long N_max, long N_o, long N_per_T, long min_acc_per_site,
                       double min_accrate, double sigma_min, double *sigma_io, double *E_out, double *beta_out,
                       double *trace, long trace_cap) {
    return run_annealing_traced(o, T_i, T_fac, N_max, N_o, N_per_T, min_acc_per_site, min_accrate, sigma_min, sigma_io, E_out, beta_out,
                                trace, NULL, NULL, trace_cap);
}
/* the same run with the beta series (MonteCarlo.jl:159) and the running acceptance rate of the current temperature step (the R of the
 * reference's checkpoint line, MonteCarlo.jl:161-167) per annealing sweep: what the notebook's cell 25 prints at sweeps 5000 / 10000 / 15000 */
long orc_run_annealing_traced(orc_t *o, double T_i, double T_fac, long N_max, long N_o, long N_per_T, long min_acc_per_site,
                              double min_accrate, double sigma_min, double *sigma_io, double *E_out, double *beta_out,
                              double *trace, double *beta_trace, double *rate_trace, long trace_cap) {
    return run_annealing_traced(o, T_i, T_fac, N_max, N_o, N_per_T, min_acc_per_site, min_accrate, sigma_min, sigma_io, E_out, beta_out,
                                trace, beta_trace, rate_trace, trace_cap);
}

/* ------------------------------------------------------------------ the framework's counter-based stream, restated */
/* Not from the reference (whose Julia streams are not reproducible): this is the product's own Philox4x32-10
 * contract (DESIGN.md "random stream v1"), restated independently so that FP64 device chains can be followed step by step. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
/* words for (site, replica, visit): xi[2d] (re parts then im parts) and the acceptance uniform */
void orc_stream_v1(uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, int d, double *xi, double *u) {
    uint32_t w[16]; int ncall = (2 * d + 3) / 4;
    for (int call = 0; call < ncall; call++) {
        uint32_t c[4] = {site, replica, (uint32_t)visit, (uint32_t)((visit >> 32) << 4) | (uint32_t)call};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        memcpy(w + 4 * call, c, sizeof(c));
    }
    for (int m = 0; m < d; m++) {
        double u1 = (double)((w[2 * m] >> 8) + 1u) * (1.0 / 16777216.0);
        double u2 = (double)(w[2 * m + 1] >> 8) * (1.0 / 16777216.0);
        double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586476925 * u2;
        xi[m] = r * cos(a); xi[d + m] = r * sin(a);
    }
    uint32_t x = (w[0] & 0xFFu) | ((w[2] & 0xFFu) << 8) | ((w[4] & 0xFFu) << 16) | ((w[6] & 0xFFu) << 24);
    *u = (double)x * (1.0 / 4294967296.0);
}
/* "stream v2" (DESIGN.md section 5), restated independently of the device code: Philox4x32-7; ONE call gives the d complex normals of an
 * update for d <= 4 (two calls for d = 6): word m -> u1 = (2 (w >> 16) + 1) / 2^17, u2 = (w & 0xFFFF) / 2^16, Box-Muller; the acceptance
 * uniform of visit v is word (v & 3) of one further call whose counter holds v >> 2 and the tag 8. */
static void philox4x32_r(uint32_t c[4], uint32_t k0, uint32_t k1, int rounds) {
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
void orc_stream_v2(uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, int d, double *xi, double *u) {
    uint32_t w[8]; int ncall = (d + 3) / 4;
    for (int call = 0; call < ncall; call++) {
        uint32_t c[4] = {site, replica, (uint32_t)visit, (uint32_t)((visit >> 32) << 4) | (uint32_t)call};
        philox4x32_r(c, (uint32_t)seed, (uint32_t)(seed >> 32), 7);
        memcpy(w + 4 * call, c, sizeof(c));
    }
    for (int m = 0; m < d; m++) {
        double u1 = (double)((w[m] >> 15) | 1u) * (1.0 / 131072.0);
        double u2 = (double)(w[m] & 0xFFFFu) * (1.0 / 65536.0);
        double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586476925 * u2;
        xi[m] = r * cos(a); xi[d + m] = r * sin(a);
    }
    uint64_t g = visit >> 2;
    uint32_t c[4] = {site, replica, (uint32_t)g, (uint32_t)((g >> 32) << 4) | 8u};
    philox4x32_r(c, (uint32_t)seed, (uint32_t)(seed >> 32), 7);
    *u = (double)c[visit & 3] * (1.0 / 4294967296.0);
}
void orc_stream(int version, uint64_t seed, uint32_t site, uint32_t replica, uint64_t visit, int d, double *xi, double *u) {
    if (version == 2) orc_stream_v2(seed, site, replica, visit, d, xi, u); else orc_stream_v1(seed, site, replica, visit, d, xi, u);
}
/* Colour-ordered sweeps driven by stream v1 (the product's chain definition): for each sweep, colours in ascending order, every
 * site of the colour receives `hits` consecutive updates, visit = (sweep0 + s) * hits + hit.  site_ids = original site numbers. */
long orc_sweeps_colour(orc_t *o, int version, long n_sweeps, int hits, const int *colour, int n_colours, double beta, double sigma,
                       uint64_t seed, uint32_t replica, uint64_t sweep0, unsigned char *accept /* [n_sweeps][hits][N] or NULL */);
long orc_sweeps_colour_v1(orc_t *o, long n_sweeps, int hits, const int *colour, int n_colours, double beta, double sigma,
                          uint64_t seed, uint32_t replica, uint64_t sweep0, unsigned char *accept /* [n_sweeps][hits][N] or NULL */) {
    return orc_sweeps_colour(o, 1, n_sweeps, hits, colour, n_colours, beta, sigma, seed, replica, sweep0, accept);
}
long orc_sweeps_colour(orc_t *o, int version, long n_sweeps, int hits, const int *colour, int n_colours, double beta, double sigma,
                       uint64_t seed, uint32_t replica, uint64_t sweep0, unsigned char *accept /* [n_sweeps][hits][N] or NULL */) {
    long acc = 0; double xi[16], u;
    for (long s = 0; s < n_sweeps; s++)
        for (int c = 0; c < n_colours; c++)
            for (int i = 0; i < o->N; i++) {
                if (colour[i] != c) continue;
                for (int hit = 0; hit < hits; hit++) {
                    uint64_t visit = (sweep0 + (uint64_t)s) * (uint64_t)hits + (uint64_t)hit;
                    orc_stream(version, seed, (uint32_t)i, replica, visit, o->d, xi, &u);
                    int a = orc_local_update(o, i, xi, u, beta, sigma, NULL);
                    if (accept) accept[((size_t)s * hits + hit) * o->N + i] = (unsigned char)a;
                    acc += a;
                }
            }
    return acc;
}
/* Haar-random initial state from stream v1 with the reserved visit number 2^60-1 (product's scmc_randomize) */
void orc_randomize_v1(orc_t *o, uint64_t seed, uint32_t replica) {
    double xi[16], u;
    for (int i = 0; i < o->N; i++) {
        orc_stream_v1(seed, (uint32_t)i, replica, ((uint64_t)1 << 60) - 1, o->d, xi, &u);
        double n2 = 0.0; for (int k = 0; k < 2 * o->d; k++) n2 += xi[k] * xi[k];
        double inv = 1.0 / sqrt(n2);
        for (int k = 0; k < o->d; k++) { o->re[i * o->d + k] = xi[k] * inv; o->im[i * o->d + k] = xi[o->d + k] * inv; }
        refresh_site(o, i);
    }
}
