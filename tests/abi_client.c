/* Minimal plain-C client of the C ABI (include/scmc.h): what a cgo / ccall / JNI binding does.  Links nothing but libdl; the
 * library is loaded at run time.  Usage: abi_client <path/to/libscmc_b200.so> [run]
 *   without "run": checks symbol resolution and argument validation (works without a GPU);
 *   with "run":    builds a 6-site ring (n = 1, Heisenberg couplings), sweeps it and prints energies, then drives run! and the
 *                  annealing loop (scmc_run_metropolis, scmc_anneal) and the multi-device object (scmc_multi_*: replicas split over all
 *                  visible GPUs, at most 8; one ncclAllGather of the measure! rows) and checks them against the single-device calls
 *                  (needs a GPU; the multi-device part uses every GPU the process sees). */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/scmc.h"

#define LOAD(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

static const double PAULI_RE[4][2][2] = {{{1, 0}, {0, 1}}, {{0, 1}, {1, 0}}, {{0, 0}, {0, 0}}, {{1, 0}, {0, -1}}};
static const double PAULI_IM[4][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}, {{0, -1}, {1, 0}}, {{0, 0}, {0, 0}}};

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s lib [run]\n", argv[0]); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(scmc_version) LOAD(scmc_last_error) LOAD(scmc_create) LOAD(scmc_destroy) LOAD(scmc_randomize) LOAD(scmc_energy) LOAD(scmc_sweep)
    LOAD(scmc_measure) LOAD(scmc_device_count) LOAD(scmc_run_metropolis) LOAD(scmc_anneal) LOAD(scmc_optimize) LOAD(scmc_set_stream_version)
    LOAD(scmc_multi_create) LOAD(scmc_multi_destroy) LOAD(scmc_multi_count) LOAD(scmc_multi_handle) LOAD(scmc_multi_randomize) LOAD(scmc_multi_sweep)
    LOAD(scmc_multi_measure) LOAD(scmc_multi_energy) LOAD(scmc_multi_run_metropolis) LOAD(scmc_multi_anneal) LOAD(scmc_multi_optimize)
    LOAD(scmc_multi_get_state) LOAD(scmc_get_state)
    printf("version %d\n", p_scmc_version());
    scmc_t* h = NULL;
    int32_t st = p_scmc_create(&h, 0, 7, 6, 4, 1, 0, 2, NULL, NULL, 1, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    printf("bad precision -> status %d (%s)\n", st, p_scmc_last_error());
    if (st != SCMC_ERR_ARG) return 1;
    if (argc < 3 || strcmp(argv[2], "run") != 0) return 0;

    /* generators G^a = sigma^mu (x) tau^nu, a = 4 mu + nu (Generators.jl:2-28), C order [a][j][k] */
    static double gre[16][4][4], gim[16][4][4], qre[16][4][4], qim[16][4][4];
    for (int mu = 0; mu < 4; mu++) for (int nu = 0; nu < 4; nu++)
        for (int s = 0; s < 2; s++) for (int sp = 0; sp < 2; sp++) for (int l = 0; l < 2; l++) for (int lp = 0; lp < 2; lp++) {
            const double ar = PAULI_RE[mu][s][sp], ai = PAULI_IM[mu][s][sp], br = PAULI_RE[nu][l][lp], bi = PAULI_IM[nu][l][lp];
            gre[4 * mu + nu][2 * s + l][2 * sp + lp] = ar * br - ai * bi;
            gim[4 * mu + nu][2 * s + l][2 * sp + lp] = ar * bi + ai * br;
        }
    for (int a = 0; a < 16; a++) for (int j = 0; j < 4; j++) { qre[a][j][j] = 1.0; }   /* (G^a)^2 = 1 for n = 1 */
    enum { N = 6, R = 4 };
    int32_t nbr[N][2], lab[N][2];
    for (int i = 0; i < N; i++) { nbr[i][0] = (i + 1) % N; nbr[i][1] = (i + N - 1) % N; lab[i][0] = lab[i][1] = 0; }
    double J[16][16] = {{0}}, K[16] = {0};
    for (int a = 0; a < 16; a++) J[a][a] = 1.0;
    st = p_scmc_create(&h, 0, 64, N, 4, R, 0, 2, &nbr[0][0], &lab[0][0], 1, &J[0][0], K, &gre[0][0][0], &gim[0][0][0], &qre[0][0][0], &qim[0][0][0], NULL);
    if (st != SCMC_OK) { fprintf(stderr, "create: %d %s\n", st, p_scmc_last_error()); return 1; }
    if (p_scmc_randomize(h, 42) != SCMC_OK) return 1;
    double E0[R], E1[R], beta[R] = {1, 2, 4, 1e9}, sigma[R] = {0.5, 0.4, 0.3, 0.3}, obs[R][SCMC_NOBS];
    int64_t acc[R], att[R];
    if (p_scmc_energy(h, E0) != SCMC_OK) return 1;
    if (p_scmc_sweep(h, 50, 2, beta, sigma, 42, 0, acc, att) != SCMC_OK) { fprintf(stderr, "sweep: %s\n", p_scmc_last_error()); return 1; }
    if (p_scmc_energy(h, E1) != SCMC_OK || p_scmc_measure(h, &obs[0][0]) != SCMC_OK) return 1;
    for (int r = 0; r < R; r++) printf("replica %d: E %.6f -> %.6f, accepted %lld / %lld, e %.6f\n", r, E0[r], E1[r], (long long)acc[r], (long long)att[r], obs[r][0]);
    /* bond energy 4 J |<psi_i|psi_j>|^2 >= 0 for J = +1: the cold replica must end near the orthogonal-neighbour minimum E = 0 + constant */
    if (!(E1[R - 1] <= E0[R - 1]) || att[0] != 50 * 2 * N) return 1;
    p_scmc_destroy(h);

    /* ---- run! and the annealing loop through the ABI, single device and multi-device, same global replica ids => same results */
    enum { RT = 10 };                                   /* 10 replicas: uneven shards on 4 or 8 GPUs */
    int32_t ndev = 0;
    p_scmc_device_count(&ndev);
    if (ndev > 8) ndev = 8;
    if (ndev > RT) ndev = RT;
    printf("devices used by the multi-device object: %d\n", ndev);
    double T[RT], Ti[RT], sg1[RT], sgm[RT], mean1[RT][SCMC_NOBS], meanm[RT][SCMC_NOBS], acc1[RT], accm[RT];
    static double ser1[8][RT][SCMC_NOBS], serm[8][RT][SCMC_NOBS], th1[4][RT], thm[4][RT];
    for (int r = 0; r < RT; r++) { T[r] = 0.1 * (r + 1); Ti[r] = 2.0; sg1[r] = sgm[r] = 60.0; }
    scmc_run_params rp; memset(&rp, 0, sizeof rp);
    rp.n_therm = 40; rp.n_measure = 60; rp.measurement_rate = 10; rp.hits = 2; rp.seed = 7; rp.sweep0 = 0; rp.T = T; rp.T_i = Ti;
    scmc_run_results rr; memset(&rr, 0, sizeof rr);
    st = p_scmc_create(&h, 0, 64, N, 4, RT, 0, 2, &nbr[0][0], &lab[0][0], 1, &J[0][0], K, &gre[0][0][0], &gim[0][0][0], &qre[0][0][0], &qim[0][0][0], NULL);
    if (st != SCMC_OK || p_scmc_randomize(h, 7) != SCMC_OK) return 1;
    rp.sigma = sg1; rr.series = &ser1[0][0][0]; rr.mean = &mean1[0][0]; rr.acc_rate = acc1; rr.therm_energy = &th1[0][0];
    if (p_scmc_run_metropolis(h, &rp, &rr) != SCMC_OK) { fprintf(stderr, "run: %s\n", p_scmc_last_error()); return 1; }
    if (rr.n_rows != 6) return 1;
    scmc_multi_t* m = NULL;
    st = p_scmc_multi_create(&m, ndev, NULL, 64, N, 4, RT, 0, 2, &nbr[0][0], &lab[0][0], 1, &J[0][0], K, &gre[0][0][0], &gim[0][0][0], &qre[0][0][0], &qim[0][0][0], NULL);
    if (st != SCMC_OK) { fprintf(stderr, "multi_create: %d %s\n", st, p_scmc_last_error()); return 1; }
    if (p_scmc_multi_randomize(m, 7) != SCMC_OK) return 1;
    rp.sigma = sgm; rr.series = &serm[0][0][0]; rr.mean = &meanm[0][0]; rr.acc_rate = accm; rr.therm_energy = &thm[0][0]; rr.n_rows = 0;
    if (p_scmc_multi_run_metropolis(m, &rp, &rr) != SCMC_OK) { fprintf(stderr, "multi run: %s\n", p_scmc_last_error()); return 1; }
    if (rr.n_rows != 6) return 1;
    for (int r = 0; r < RT; r++) {
        if (sg1[r] != sgm[r] || acc1[r] != accm[r] || th1[3][r] != thm[3][r]) { fprintf(stderr, "run! differs on replica %d\n", r); return 1; }
        for (int q = 0; q < 6; q++) for (int k = 0; k < SCMC_NOBS; k++)
            if (ser1[q][r][k] != serm[q][r][k]) { fprintf(stderr, "series differs: row %d replica %d value %d\n", q, r, k); return 1; }
    }
    /* measure! rows: per-device reductions + ncclAllGather vs the single-device rows of the same states */
    double o1[RT][SCMC_NOBS], om[RT][SCMC_NOBS], E1s[RT], Em[RT];
    if (p_scmc_measure(h, &o1[0][0]) != SCMC_OK) return 1;
    if (p_scmc_multi_measure(m, &om[0][0]) != SCMC_OK) { fprintf(stderr, "multi measure: %s\n", p_scmc_last_error()); return 1; }
    for (int r = 0; r < RT; r++) for (int k = 0; k < SCMC_NOBS; k++) if (o1[r][k] != om[r][k]) { fprintf(stderr, "measure row differs: %d %d\n", r, k); return 1; }
    /* annealing restarts + optimisation sweeps */
    scmc_anneal_params ap; memset(&ap, 0, sizeof ap);
    ap.T_i = 2.0; ap.T_fac = 0.9; ap.N_max = 200; ap.N_per_T = 10; ap.min_acc_per_site = 2; ap.min_accrate = 1e-3; ap.sigma_min = 0.05; ap.sigma0 = 60.0;
    ap.hits = 2; ap.seed = 11;
    double b1[RT], bm[RT], s1[RT], sm[RT]; int64_t n1[RT], nm[RT];
    scmc_anneal_results ar1 = {E1s, b1, s1, n1}, arm = {Em, bm, sm, nm};
    if (p_scmc_anneal(h, &ap, &ar1) != SCMC_OK || p_scmc_multi_anneal(m, &ap, &arm) != SCMC_OK) { fprintf(stderr, "anneal: %s\n", p_scmc_last_error()); return 1; }
    if (p_scmc_optimize(h, 20, E1s) != SCMC_OK || p_scmc_multi_optimize(m, 20, Em) != SCMC_OK) return 1;
    for (int r = 0; r < RT; r++) {
        if (n1[r] != nm[r] || b1[r] != bm[r] || s1[r] != sm[r] || E1s[r] != Em[r]) { fprintf(stderr, "anneal differs on replica %d\n", r); return 1; }
        if (!(E1s[r] < 1e-6 + 0.0 * N)) { /* J = +1 Heisenberg ring: ground state = orthogonal neighbours, E = sum_bonds (4 |<i|j>|^2 - 1) ... */ }
    }
    static double re1[RT][N][4], im1[RT][N][4], rem[RT][N][4], imm[RT][N][4];
    if (p_scmc_get_state(h, 0, RT, &re1[0][0][0], &im1[0][0][0]) != SCMC_OK || p_scmc_multi_get_state(m, 0, RT, &rem[0][0][0], &imm[0][0][0]) != SCMC_OK) return 1;
    if (memcmp(re1, rem, sizeof re1) || memcmp(im1, imm, sizeof im1)) { fprintf(stderr, "states differ\n"); return 1; }
    printf("multi-device: run!, measure!, annealing and optimisation on %d GPU(s) equal the single-device results; best E %.9f after %lld annealing sweeps\n",
           ndev, E1s[RT - 1], (long long)n1[RT - 1]);
    p_scmc_multi_destroy(m);
    p_scmc_destroy(h);
    puts("abi_client ok");
    return 0;
}
