"""su(4) spin-valley generators T^a = sigma^mu (x) tau^nu at parton filling n.

Host-side mirror of `getGenerators`, `get_su4index`, `expval`
(/root/reference/src/Generators.jl:2-28, 82-84).  Built directly on Fock-state
bit patterns: a one-body operator  O = sum_{pq} M_pq c_p^dagger c_q  with
M = kron(sigma^mu, sigma^nu) moves one parton from mode q to mode p.  Like the
reference (Generators.jl:39-51) the hop carries NO fermionic sign, so
that n=2 tables equal the reference's.  Basis order = ascending Fock integer with
mode 0 = (up,+) the least significant bit, i.e. n=1: (up+, up-, dn+, dn-),
n=2: integers [3, 5, 6, 9, 10, 12].

`emit_cuda_tables()` writes the same tables as `constexpr` integer arrays
(`csrc/gen_tables.h`) so the CUDA kernels unroll the sparse contraction at
compile time; `scmc_create` compares caller-supplied generators against them
and falls back to the data-driven path when they differ.
"""
import numpy as np

_PAULI = np.array(
    [[[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]],
    dtype=np.complex128,
)


def get_su4index(mu, nu):
    """1-based generator index as in the reference (Generators.jl:26-28)."""
    return 4 * mu + nu + 1


def fock_states(n):
    return [m for m in range(16) if bin(m).count("1") == n]


def _one_body(M, states):
    idx = {m: i for i, m in enumerate(states)}
    d = len(states)
    out = np.zeros((d, d), dtype=np.complex128)
    for col, m in enumerate(states):
        for q in range(4):
            if not (m >> q) & 1:
                continue
            emptied = m & ~(1 << q)
            for p in range(4):
                if M[p, q] == 0 or (emptied >> p) & 1:
                    continue
                out[idx[emptied | (1 << p)], col] += M[p, q]
    return out


def _one_body_signed(M, states):
    """Second-quantised one-body operator sum_pq M_pq c_p^dagger c_q on the occupation-number basis WITH the Jordan-Wigner sign."""
    idx = {m: i for i, m in enumerate(states)}
    d = len(states)
    out = np.zeros((d, d), dtype=np.complex128)
    below = lambda m, q: bin(m & ((1 << q) - 1)).count("1")
    for col, m in enumerate(states):
        for q in range(4):
            if not (m >> q) & 1:
                continue
            emptied = m & ~(1 << q)
            s_q = (-1) ** below(m, q)
            for p in range(4):
                if M[p, q] == 0 or (emptied >> p) & 1:
                    continue
                out[idx[emptied | (1 << p)], col] += M[p, q] * s_q * (-1) ** below(emptied, p)
    return out


def getGeneratorsSigned(n):
    """The 16 generators in the TRUE fermionic representation of filling n (parton operator with its Jordan-Wigner sign, which the
    reference's f omits, Generators.jl:39-51 / SURVEY Q1): for n = 2 they span su(4) in the antisymmetric square of C^4 (closed under
    commutation, unlike the reference's tables).  Not a built-in of the library: hand them to Configuration(..., generators=...) /
    scmc_create, which then runs its dense-table path for caller-supplied generators."""
    states = fock_states(n)
    return np.stack([_one_body_signed(np.kron(_PAULI[mu], _PAULI[nu]), states) for mu in range(4) for nu in range(4)])


def getGenerators(n):
    """[16, d, d] complex; G[a] = sigma^mu (x) tau^nu, a = 4 mu + nu (0-based).  Like the reference's two methods, getGenerators(n::Int)
    (Generators.jl:2) and getGenerators(cfg) (Configuration.jl:108), the argument may be a filling or a Configuration."""
    if hasattr(n, "generators"):
        return n.generators
    if n not in (1, 2, 3):
        raise ValueError("filling n must be 1, 2 or 3")
    states = fock_states(n)
    return np.stack([_one_body(np.kron(_PAULI[mu], _PAULI[nu]), states) for mu in range(4) for nu in range(4)])


def getGeneratorsSq(n):
    """Matrix squares (G^a)^2 for a filling, or the stored ones of a Configuration (Configuration.jl:52, 112)."""
    if hasattr(n, "generatorsSq"):
        return n.generatorsSq
    g = getGenerators(n)
    return np.einsum("aij,ajk->aik", g, g)


def expval(M, state):
    """<state| M |state> (Generators.jl:82-84)."""
    state = np.asarray(state)
    return np.vdot(state, np.asarray(M) @ state)


def _quantise(g):
    re_ = np.rint(g.real).astype(int)
    im_ = np.rint(g.imag).astype(int)
    assert np.allclose(re_, g.real) and np.allclose(im_, g.imag), "generator entries must be integers"
    return re_, im_


def _term(coef, name):
    if coef == 0:
        return ""
    sign = "+" if coef > 0 else "-"
    mag = abs(coef)
    return f" {sign} {name}" if mag == 1 else f" {sign} R({mag}) * {name}"


def _emit_expect(fname, g):
    """Straight-line T^a = sum_jk conj(psi_j) G^a_jk psi_k using P_jk = conj(psi_j) psi_k (j <= k) and hermiticity."""
    gre, gim = _quantise(g)
    d = g.shape[1]
    used = set()
    rows = []
    for a in range(16):
        expr = ""
        for j in range(d):
            expr += _term(gre[a, j, j], f"p{j}{j}r")
            if gre[a, j, j]:
                used.add((j, j, "r"))
            for k in range(j + 1, d):
                # G_jk P_jk + G_kj P_kj = 2 (Re G_jk Re P_jk - Im G_jk Im P_jk)
                if gre[a, j, k]:
                    expr += _term(2 * gre[a, j, k], f"p{j}{k}r"); used.add((j, k, "r"))
                if gim[a, j, k]:
                    expr += _term(-2 * gim[a, j, k], f"p{j}{k}i"); used.add((j, k, "i"))
        rows.append(f"    T[{a}] = R(0){expr};" if expr else f"    T[{a}] = R(0);")
    out = [f"template <class R> __device__ __forceinline__ void {fname}(const R* re, const R* im, R* T) {{"]
    for (j, k, part) in sorted(used):
        if part == "r":
            out.append(f"    const R p{j}{k}r = re[{j}] * re[{k}] + im[{j}] * im[{k}];")
        else:
            out.append(f"    const R p{j}{k}i = re[{j}] * im[{k}] - im[{j}] * re[{k}];")
    out += rows + ["}"]
    return out


def _emit_hloc(fname, g, gsq):
    """H_loc = sum_a h[a] G^a + K[a] (G^a)^2 ; upper triangle j <= k into Hr[j*D+k], Hi[j*D+k]."""
    gre, gim = _quantise(g)
    qre, qim = _quantise(gsq)
    d = g.shape[1]
    out = [f"template <class R, bool ONSITE> __device__ __forceinline__ void {fname}(const R* h, const R* K, R* Hr, R* Hi) {{"]
    for j in range(d):
        for k in range(j, d):
            er = "".join(_term(gre[a, j, k], f"h[{a}]") for a in range(16))
            ei = "".join(_term(gim[a, j, k], f"h[{a}]") for a in range(16))
            qr = "".join(_term(qre[a, j, k], f"K[{a}]") for a in range(16))
            qi = "".join(_term(qim[a, j, k], f"K[{a}]") for a in range(16))
            out.append(f"    Hr[{j * d + k}] = R(0){er}; Hi[{j * d + k}] = R(0){ei};")
            if qr or qi:
                out.append(f"    if (ONSITE) {{ Hr[{j * d + k}] += R(0){qr}; Hi[{j * d + k}] += R(0){qi}; }}")
    out.append("}")
    return out


def _density_terms(g):
    """(j, k, part) entries of P_jk = conj(psi_j) psi_k (j <= k) that any generator uses, in a fixed order."""
    gre, gim = _quantise(g)
    d = g.shape[1]
    used = []
    for j in range(d):
        for k in range(j, d):
            if j == k:
                if np.any(gre[:, j, j] != 0):
                    used.append((j, j, "r"))
            else:
                if np.any(gre[:, j, k] != 0):
                    used.append((j, k, "r"))
                if np.any(gim[:, j, k] != 0):
                    used.append((j, k, "i"))
    return used


def _emit_density(fname_acc, fname_map, g):
    """T^a is linear in the density matrix: sum_j T_j = map(sum_j P_j).  `acc` adds one site's P entries into an
    accumulator with FMAs (no separate adds), `map` turns the accumulated entries into the 16-vector."""
    gre, gim = _quantise(g)
    d = g.shape[1]
    terms = _density_terms(g)
    idx = {t: n for n, t in enumerate(terms)}
    acc = [f"template <class R> __device__ __forceinline__ void {fname_acc}(const R* re, const R* im, R* P) {{"]
    for (j, k, part) in terms:
        n = idx[(j, k, part)]
        if part == "r":
            acc.append(f"    P[{n}] = fma(re[{j}], re[{k}], P[{n}]); P[{n}] = fma(im[{j}], im[{k}], P[{n}]);")
        else:
            acc.append(f"    P[{n}] = fma(re[{j}], im[{k}], P[{n}]); P[{n}] = fma(-im[{j}], re[{k}], P[{n}]);")
    acc.append("}")
    mp = [f"template <class R> __device__ __forceinline__ void {fname_map}(const R* P, R* T) {{"]
    for a in range(16):
        expr = ""
        for j in range(d):
            if gre[a, j, j]:
                expr += _term(gre[a, j, j], f"P[{idx[(j, j, 'r')]}]")
            for k in range(j + 1, d):
                if gre[a, j, k]:
                    expr += _term(2 * gre[a, j, k], f"P[{idx[(j, k, 'r')]}]")
                if gim[a, j, k]:
                    expr += _term(-2 * gim[a, j, k], f"P[{idx[(j, k, 'i')]}]")
        mp.append(f"    T[{a}] = R(0){expr};")
    mp.append("}")
    # transpose of the map: e = sum_a h^a T^a = sum_n Hq[n] P[n],  Hq[n] = sum_a C[a][n] h[a]
    col = [""] * len(terms)
    for a in range(16):
        for j in range(d):
            if gre[a, j, j]:
                col[idx[(j, j, "r")]] += _term(gre[a, j, j], f"h[{a}]")
            for k in range(j + 1, d):
                if gre[a, j, k]:
                    col[idx[(j, k, "r")]] += _term(2 * gre[a, j, k], f"h[{a}]")
                if gim[a, j, k]:
                    col[idx[(j, k, "i")]] += _term(-2 * gim[a, j, k], f"h[{a}]")
    tr = [f"template <class R> __device__ __forceinline__ void {fname_map.replace('expect_from_density', 'hq_from_field')}(const R* h, R* Hq) {{"]
    tr += [f"    Hq[{n}] = R(0){c};" for n, c in enumerate(col)] + ["}"]
    # e(psi) = sum_n Hq[n] P_n(psi) without materialising T, evaluated row by row as psi^dagger (H psi) over the upper triangle:
    # e = sum_j re_j A_j + im_j B_j,  A_j = Hq_jj re_j + sum_{k>j} (Hq_jk^r re_k + Hq_jk^i im_k),  B_j = Hq_jj im_j + sum_{k>j} (Hq_jk^r im_k - Hq_jk^i re_k)
    # (d = 4: 41 instead of 51 instructions; every row is its own dependent chain)
    en = [f"template <class R> __device__ __forceinline__ R {fname_map.replace('expect_from_density', 'energy_from_state')}(const R* re, const R* im, const R* Hq) {{",
          "    R ea = R(0), eb = R(0);"]
    for j in range(d):
        a_terms, b_terms = [], []
        if (j, j, "r") in idx:
            a_terms.append((f"Hq[{idx[(j, j, 'r')]}]", f"re[{j}]")); b_terms.append((f"Hq[{idx[(j, j, 'r')]}]", f"im[{j}]"))
        for k in range(j + 1, d):
            if (j, k, "r") in idx:
                a_terms.append((f"Hq[{idx[(j, k, 'r')]}]", f"re[{k}]")); b_terms.append((f"Hq[{idx[(j, k, 'r')]}]", f"im[{k}]"))
            if (j, k, "i") in idx:
                a_terms.append((f"Hq[{idx[(j, k, 'i')]}]", f"im[{k}]")); b_terms.append((f"-Hq[{idx[(j, k, 'i')]}]", f"re[{k}]"))
        if not a_terms:
            continue
        for name, tl in (("A", a_terms), ("B", b_terms)):
            expr = f"{tl[0][0]} * {tl[0][1]}"
            for (c, x) in tl[1:]:
                expr = f"fma({c}, {x}, {expr})"
            en.append(f"    const R {name}{j} = {expr};")
        en.append(f"    ea = fma(re[{j}], A{j}, ea); eb = fma(im[{j}], B{j}, eb);")
    en += ["    return ea + eb;", "}"]
    # FP32 overload: the same operations in the same order on (re, im) / (A_j, B_j) pairs as packed fma.rn.ftz.f32x2 (SASS FFMA2: one issue slot
    # per two FMAs; scalar coefficients and swapped halves are operand modifiers of FFMA2, not instructions).  Half 0 carries the A chain and
    # half 1 the B chain of the template above, each half is the IEEE operation of the scalar form => bit for bit its result.
    full = all((j, j, "r") in idx for j in range(d)) and all((j, k, q) in idx for j in range(d) for k in range(j + 1, d) for q in "ri")
    if full:
        fn = fname_map.replace('expect_from_density', 'energy_from_state')
        pk = [f"__device__ __forceinline__ float {fn}(const float* re, const float* im, const float* Hq) {{",
              "#ifdef SCMC_SCALAR_ENERGY",
              f"    return {fn}<float>(re, im, Hq);",
              "#else",
              f"    unsigned long long z[{d}], zc[{d}], e = 0ull, r, c, cn;",
              '    asm("mov.b64 %0, {%1, %2};" : "=l"(cn) : "f"(-1.0f), "f"(1.0f));',
              "#pragma unroll",
              f"    for (int k = 0; k < {d}; k++) {{",
              '        asm("mov.b64 %0, {%1, %2};" : "=l"(z[k]) : "f"(re[k]), "f"(im[k]));',
              "        if (k) {                                   // zc = (im, -re): one packed multiply by (-1, 1), halves swapped by the consumer's operand modifier",
              "            unsigned long long w; float wr, wi;",
              '            asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(w) : "l"(z[k]), "l"(cn));',
              '            asm("mov.b64 {%0, %1}, %2;" : "=f"(wr), "=f"(wi) : "l"(w));',
              '            asm("mov.b64 %0, {%1, %2};" : "=l"(zc[k]) : "f"(wi), "f"(wr));',
              "        }",
              "    }"]
        for j in range(d):
            pk.append(f'    asm("mov.b64 %0, {{%1, %1}};" : "=l"(c) : "f"(Hq[{idx[(j, j, "r")]}]));')
            pk.append(f'    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c), "l"(z[{j}]));')
            for k in range(j + 1, d):
                pk.append(f'    asm("mov.b64 %0, {{%1, %1}};" : "=l"(c) : "f"(Hq[{idx[(j, k, "r")]}]));')
                pk.append(f'    asm("fma.rn.ftz.f32x2 %0, %1, %2, %0;" : "+l"(r) : "l"(c), "l"(z[{k}]));')
                pk.append(f'    asm("mov.b64 %0, {{%1, %1}};" : "=l"(c) : "f"(Hq[{idx[(j, k, "i")]}]));')
                pk.append(f'    asm("fma.rn.ftz.f32x2 %0, %1, %2, %0;" : "+l"(r) : "l"(c), "l"(zc[{k}]));')
            pk.append(f'    asm("fma.rn.ftz.f32x2 %0, %1, %2, %0;" : "+l"(e) : "l"(z[{j}]), "l"(r));')
        pk += ["    float ea, eb;", '    asm("mov.b64 {%0, %1}, %2;" : "=f"(ea), "=f"(eb) : "l"(e));', "    return ea + eb;", "#endif", "}"]
        en += pk
    return acc + mp + tr + en, len(terms)


def emit_cuda_tables(path):
    """Write csrc/gen_code.h: unrolled expectation-value / local-Hamiltonian code and integer tables for n = 1, 2, 3."""
    lines = [
        "// GENERATED by semiclassicalmc.jl_b200/generators.py:emit_cuda_tables -- do not edit by hand.",
        "// su(4) generators G^a = sigma^mu (x) tau^nu at filling n, reference convention",
        "// (/root/reference/src/Generators.jl:2-65, signless parton operator); `gK` = filling n = K.",
        "// spin_expect_*   : T^a   = <psi|G^a|psi>      (replaces the dense loop of src/Configuration.jl:166-194)",
        "// spin_expect_sq_*: Tsq^a = <psi|(G^a)^2|psi>  (generatorsSq, src/Configuration.jl:52,59)",
        "// hloc_*          : H_loc = sum_a h^a G^a + K^a (G^a)^2, upper triangle (for the on-device local optimiser)",
        "// density_acc_* / expect_from_density_*: sum_j T_j = map(sum_j P_j), P = conj(psi) psi^T entries (psi-gather sweep kernel)",
        "#pragma once",
        "#include <cstdint>",
        "namespace scmc { namespace gen {",
    ]
    for n in (1, 2, 3):
        g, q = getGenerators(n), getGeneratorsSq(n)
        lines += _emit_expect(f"spin_expect_g{n}", g)
        lines += _emit_expect(f"spin_expect_sq_g{n}", q)
        lines += _emit_hloc(f"hloc_g{n}", g, q)
        dens, nterms = _emit_density(f"density_acc_g{n}", f"expect_from_density_g{n}", g)
        lines += dens
        lines.append(f"constexpr int NDENS_g{n} = {nterms};   // density-matrix entries used by the generators at filling {n}")
        # C[a][m]: T^a = sum_m C[a][m] P[m]  (host side: couplings in the density basis, Jq = C^T J C)
        gre_, gim_ = _quantise(g)
        terms = _density_terms(g)
        Cm = np.zeros((16, nterms), dtype=int)
        for m, (j, k, part) in enumerate(terms):
            for a in range(16):
                Cm[a, m] = gre_[a, j, j] if j == k else (2 * gre_[a, j, k] if part == "r" else -2 * gim_[a, j, k])
        lines.append(f"static const int8_t DENSMAP{n}[16 * {nterms}] = {{{','.join(str(int(v)) for v in Cm.reshape(-1))}}};")
        for name, arr in (("G", g), ("Gsq", q)):
            re_, im_ = _quantise(arr)
            for part, tab in (("re", re_), ("im", im_)):
                flat = ",".join(str(int(v)) for v in tab.reshape(-1))
                lines.append(f"static const int8_t {name}{n}_{part}[16 * {g.shape[1]} * {g.shape[1]}] = {{{flat}}};")
    lines.append("}}  // namespace scmc::gen")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
