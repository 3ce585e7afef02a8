/* Minimal plain-C client of the C ABI (include/scmc.h): what a cgo / ccall / JNI binding does.  Links nothing but libdl; the
 * library is loaded at run time.  Usage: abi_client <path/to/libscmc_b200.so> [run]
 *   without "run": checks symbol resolution and argument validation (works without a GPU);
 *   with "run":    builds a 6-site ring (n = 1, Heisenberg couplings), sweeps it and prints energies (needs a GPU). */
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
    LOAD(scmc_measure) LOAD(scmc_device_count)
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
    puts("abi_client ok");
    return 0;
}
