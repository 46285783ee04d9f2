/*
 * qmc_oracle.c -- CPU restatement of quMCMC's quantum-proposal hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (qumcmc_b200/) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Parity status: the reference delegates the statevector arithmetic to the
 * third-party simulator `qulacs` (PyPI, un-pinned: /root/reference/setup.py:37),
 * which is not installable in this image, so this file restates qulacs' *published*
 * gate semantics (exp(+i*angle/2*P) rotations, LSB = qubit 0, sequential inverse-CDF
 * sampling) and is pinned against
 *   - every deterministic known-answer value in the reference's notebooks
 *     (tests/golden/known_answers.json), incl. the qulacs rotation
 *     sign vector of rough.ipynb cells 24/26,
 *   - the reference's recorded chains (tests/golden/chain_stats.json, produced by
 *     tests/golden/make_golden.py from the pickles under /root/reference).
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/).
 *
 * Conventions:
 *   basis index idx = int(bitstring, 2); qubit q = bit q of idx (LSB = qubit 0);
 *   spin j (string position j from the left) <-> bit n-1-j; s_j = +1 for '1', -1 for '0'.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11).  Stream contract shared with the GPU  */
/* path (include/qumcmc_b200.h "RNG stream contract").                         */
/* ------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32(uint64_t seed, uint64_t chain, uint32_t hop, uint32_t slot,
                            uint32_t out[4]) {
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)chain, (uint32_t)(chain >> 32), hop, slot};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(uint32_t) * 4);
}

static inline double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

/* Per-hop draws.  slot 0: gamma uniform, evolution time t in {2..11}, mixer-group uniform;
 * slot 1: measurement uniform, acceptance uniform.
 * Replaces np.random.uniform(*gamma) (quantum_mcmc_routines.py:186),
 * np.random.choice(range(2,12)) (:125), np.random.choice(mixers,p) (mixers.py:150),
 * qulacs' internal sampling RNG (:153) and np.random.rand() (classical_mcmc_routines.py:41). */
ORC_API void orc_hop_draws(uint64_t seed, uint64_t chain, uint32_t hop, double *u_gamma, int *t,
                           double *u_group, double *u_meas, double *u_acc) {
    uint32_t w[4];
    orc_philox4x32(seed, chain, hop, 0, w);
    *u_gamma = u53(w[0], w[1]);
    *t = 2 + (int)(((uint64_t)w[2] * 10u) >> 32);
    *u_group = (double)w[3] * (1.0 / 4294967296.0);
    orc_philox4x32(seed, chain, hop, 1, w);
    *u_meas = u53(w[0], w[1]);
    *u_acc = u53(w[2], w[3]);
}

/* Initial state when the caller passes none: uniform index in [0, 2^n)
 * (get_random_state, basic_utils.py:395-404). hop = 0xFFFFFFFF, slot 0. */
ORC_API uint64_t orc_initial_state(uint64_t seed, uint64_t chain, int n) {
    uint32_t w[4];
    orc_philox4x32(seed, chain, 0xFFFFFFFFu, 0, w);
    uint64_t v = ((uint64_t)w[1] << 32) | w[0];
    return n >= 64 ? v : (v & ((1ULL << n) - 1ULL));
}

/* ------------------------------------------------------------------------- */
/* Energies                                                                    */
/* ------------------------------------------------------------------------- */

/* E(s) = 0.5 * s.(J s) + h.s  (energy_models.py:124-144), s_j = +1 for bit n-1-j set.
 * numpy routes J.dot(s) through BLAS whose summation order is unspecified; the oracle fixes
 * row-major sequential fp64 order, no FMA contraction (compile with -ffp-contract=off). */
ORC_API double orc_energy(int n, const double *J, const double *h, uint64_t idx) {
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < n; ++i) {
        double row = 0.0;
        for (int j = 0; j < n; ++j) {
            const double sj = ((idx >> (n - 1 - j)) & 1ULL) ? 1.0 : -1.0;
            row += J[(size_t)i * n + j] * sj;
        }
        const double si = ((idx >> (n - 1 - i)) & 1ULL) ? 1.0 : -1.0;
        t1 += si * row;
        t2 += h[i] * si;
    }
    return 0.5 * t1 + t2;
}

ORC_API void orc_energies(int n, const double *J, const double *h, const uint64_t *idx, int64_t m,
                          double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < m; ++k) out[k] = orc_energy(n, J, h, idx[k]);
}

/* alpha = sqrt(n) / sqrt(sum_{i>j} J_ij^2 + sum h_j^2), 1.0 if J ~ 0
 * (energy_models.py:41-47; np.allclose(J,0) has atol=1e-8). */
ORC_API double orc_alpha(int n, const double *J, const double *h) {
    int all0 = 1;
    for (int i = 0; i < n * n; ++i)
        if (fabs(J[i]) > 1e-8) { all0 = 0; break; }
    if (all0) return 1.0;
    double sJ = 0.0, sh = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) sJ += J[(size_t)i * n + j] * J[(size_t)i * n + j];
    for (int j = 0; j < n; ++j) sh += h[j] * h[j];
    return sqrt((double)n) / sqrt(sJ + sh);
}

/* Phase-layer exponent D(idx) = sum_{j<k} J_jk z_j z_k + sum_j h_j z_j with z_j = -s_j.
 * Follows the gate angles of fn_qckt_problem_half (quantum_mcmc_routines.py:70-89):
 * ZZ angle +2cJ_jk on qubits (n-1-j, n-1-k), RZ angle +2c h_j on qubit n-1-j, qulacs
 * convention exp(+i angle/2 P) (:49-51)  =>  layer = exp(+i c D(idx)).   */
ORC_API double orc_phase_D(int n, const double *J, const double *h, uint64_t idx) {
    double d = 0.0;
    for (int j = 0; j < n - 1; ++j)
        for (int k = j + 1; k < n; ++k) {
            const double zj = ((idx >> (n - 1 - j)) & 1ULL) ? -1.0 : 1.0;
            const double zk = ((idx >> (n - 1 - k)) & 1ULL) ? -1.0 : 1.0;
            d += J[(size_t)j * n + k] * zj * zk;
        }
    for (int j = 0; j < n; ++j) {
        const double zj = ((idx >> (n - 1 - j)) & 1ULL) ? -1.0 : 1.0;
        d += h[j] * zj;
    }
    return d;
}

/* "skip the whole layer if mean|h| <= 1e-3 and mean|J| <= 1e-3" (quantum_mcmc_routines.py:64-67);
 * mean over all n*n entries incl. the zero diagonal. */
ORC_API int orc_problem_layer_active(int n, const double *J, const double *h) {
    double mh = 0.0, mJ = 0.0;
    for (int i = 0; i < n; ++i) mh += fabs(h[i]);
    for (int i = 0; i < n * n; ++i) mJ += fabs(J[i]);
    mh /= (double)n;
    mJ /= (double)n * (double)n;
    return !(mh <= 1e-3 && mJ <= 1e-3);
}

/* r = int(floor(t / delta_time)) (quantum_mcmc_routines.py:127); trotter() with r<=1 still
 * applies one mixer (:95-106), hence r_eff = max(r,1) mixer layers and r_eff-1 phase layers. */
ORC_API int orc_trotter_steps(int t, double delta_time) {
    int r = (int)floor((double)t / delta_time);
    return r < 1 ? 1 : r;
}

/* ------------------------------------------------------------------------- */
/* Gate-at-a-time simulator (the reference's cost structure: one pass per gate) */
/* state: interleaved complex128 re,im, length 2^n                             */
/* ------------------------------------------------------------------------- */

/* exp(+i*angle/2 * Z_a Z_b)  (add_multi_Pauli_rotation_gate([qa,qb],[3,3],angle), :79-81) */
static void gate_zz(double *st, int n, int qa, int qb, double angle) {
    const double c = cos(0.5 * angle), s = sin(0.5 * angle);
    const int64_t N = 1LL << n;
#pragma omp parallel for schedule(static) if (n >= 14)
    for (int64_t i = 0; i < N; ++i) {
        const int par = (int)(((i >> qa) ^ (i >> qb)) & 1);
        const double ss = par ? -s : s; /* eigenvalue +1 when bits equal */
        const double re = st[2 * i], im = st[2 * i + 1];
        st[2 * i] = re * c - im * ss;
        st[2 * i + 1] = re * ss + im * c;
    }
}

/* exp(+i*angle/2 * Z_q)  (add_RZ_gate(index, angle), :89) */
static void gate_rz(double *st, int n, int q, double angle) {
    const double c = cos(0.5 * angle), s = sin(0.5 * angle);
    const int64_t N = 1LL << n;
#pragma omp parallel for schedule(static) if (n >= 14)
    for (int64_t i = 0; i < N; ++i) {
        const double ss = ((i >> q) & 1) ? -s : s;
        const double re = st[2 * i], im = st[2 * i + 1];
        st[2 * i] = re * c - im * ss;
        st[2 * i + 1] = re * ss + im * c;
    }
}

/* exp(+i*angle/2 * X_mask): a'[i] = cos(angle/2) a[i] + i sin(angle/2) a[i ^ mask]
 * (add_multi_Pauli_rotation_gate(S,[1]*|S|,angle), mixers.py:46-50,100-104) */
static void gate_xmask(double *st, int n, uint64_t mask, double angle) {
    const double c = cos(0.5 * angle), s = sin(0.5 * angle);
    const int64_t N = 1LL << n;
    int top = 63;
    while (!((mask >> top) & 1ULL)) --top;
#pragma omp parallel for schedule(static) if (n >= 14)
    for (int64_t i = 0; i < N; ++i) {
        if ((i >> top) & 1) continue; /* visit each pair once */
        const int64_t j = i ^ (int64_t)mask;
        const double ar = st[2 * i], ai = st[2 * i + 1], br = st[2 * j], bi = st[2 * j + 1];
        st[2 * i] = c * ar - s * bi;
        st[2 * i + 1] = c * ai + s * br;
        st[2 * j] = c * br - s * ai;
        st[2 * j + 1] = c * bi + s * ar;
    }
}

/* X on qubit q (initialise_qc, :33-46: one add_X_gate per '1'). */
static void gate_x(double *st, int n, int q) {
    const int64_t N = 1LL << n;
    for (int64_t i = 0; i < N; ++i) {
        if ((i >> q) & 1) continue;
        const int64_t j = i | (1LL << q);
        double t;
        t = st[2 * i]; st[2 * i] = st[2 * j]; st[2 * j] = t;
        t = st[2 * i + 1]; st[2 * i + 1] = st[2 * j + 1]; st[2 * j + 1] = t;
    }
}

/* One proposal circuit, gate by gate, exactly as run_qmcmc_quantum_ckt builds it (:110-152):
 * |0..0> -> X gates -> mixer, then (problem half; mixer) x (r-1).
 * Mixer terms: angle = -2*gamma*weight*delta_time per X-string (mixers.py:42-50; CoherentMixerSum
 * passes gamma*w, :126-129).  circJ/circh are model.circuit_J/circuit_h (:120-121). */
ORC_API void orc_evolve_gatewise(int n, const double *circJ, const double *circh, double alpha,
                                 double gamma, double delta_time, int r, int n_terms,
                                 const uint64_t *masks, const double *weights, uint64_t s_idx,
                                 double *state /* 2*2^n */) {
    const int64_t N = 1LL << n;
    memset(state, 0, sizeof(double) * 2 * (size_t)N);
    state[0] = 1.0;
    /* initialise_qc: string position i <-> qubit n-1-i, i.e. bit q of idx */
    for (int i = 0; i < n; ++i) {
        const int q = n - 1 - i;
        if ((s_idx >> q) & 1ULL) gate_x(state, n, q);
    }
    const int active = orc_problem_layer_active(n, circJ, circh);
    const double pref = 2.0 * -1.0 * (1.0 - gamma) * alpha * delta_time; /* :72 */
    if (r < 1) r = 1;
    for (int layer = 0; layer < r; ++layer) {
        if (layer > 0 && active) {
            for (int j = 0; j < n - 1; ++j)
                for (int k = j + 1; k < n; ++k)
                    gate_zz(state, n, n - 1 - j, n - 1 - k, -1.0 * (pref * circJ[(size_t)j * n + k]));
            for (int j = 0; j < n; ++j) gate_rz(state, n, n - 1 - j, -1.0 * (pref * circh[j]));
        }
        for (int m = 0; m < n_terms; ++m)
            gate_xmask(state, n, masks[m], -1.0 * 2.0 * (gamma * weights[m]) * delta_time);
    }
}

/* Same unitary with each diagonal layer collapsed to one pointwise phase exp(+i c D(idx)).
 * Not the reference's cost structure; used to validate the GPU kernels' algebra. */
ORC_API void orc_evolve_fused(int n, const double *circJ, const double *circh, double alpha,
                              double gamma, double delta_time, int r, int n_terms,
                              const uint64_t *masks, const double *weights, uint64_t s_idx,
                              double *state) {
    const int64_t N = 1LL << n;
    memset(state, 0, sizeof(double) * 2 * (size_t)N);
    state[2 * s_idx] = 1.0;
    const int active = orc_problem_layer_active(n, circJ, circh);
    const double c = (1.0 - gamma) * alpha * delta_time;
    double *D = NULL;
    if (active && r > 1) {
        D = (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) D[i] = orc_phase_D(n, circJ, circh, (uint64_t)i);
    }
    if (r < 1) r = 1;
    for (int layer = 0; layer < r; ++layer) {
        if (layer > 0 && D) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; ++i) {
                const double pc = cos(c * D[i]), ps = sin(c * D[i]);
                const double re = state[2 * i], im = state[2 * i + 1];
                state[2 * i] = re * pc - im * ps;
                state[2 * i + 1] = re * ps + im * pc;
            }
        }
        for (int m = 0; m < n_terms; ++m)
            gate_xmask(state, n, masks[m], -2.0 * gamma * weights[m] * delta_time);
    }
    free(D);
}

/* ------------------------------------------------------------------------- */
/* Measurement, accept, chain                                                  */
/* ------------------------------------------------------------------------- */

/* qulacs QuantumState.sampling(1) (called at quantum_mcmc_routines.py:153): cumulative
 * C_0 = 0, C_{k+1} = C_k + |a_k|^2 in index order (sequential fp64); the sample is the index k
 * with C_k <= u < C_{k+1}.  u is explicit here. The last index absorbs rounding (u >= C_N). */
ORC_API uint64_t orc_sample_index(int n, const double *state, double u) {
    const int64_t N = 1LL << n;
    double c = 0.0;
    for (int64_t k = 0; k < N; ++k) {
        c += state[2 * k] * state[2 * k] + state[2 * k + 1] * state[2 * k + 1];
        if (u < c) return (uint64_t)k;
    }
    /* rounding tail: last index with non-zero probability */
    for (int64_t k = N - 1; k >= 0; --k)
        if (state[2 * k] != 0.0 || state[2 * k + 1] != 0.0) return (uint64_t)k;
    return (uint64_t)(N - 1);
}

ORC_API void orc_probabilities(int n, const double *state, double *probs) {
    const int64_t N = 1LL << n;
    for (int64_t k = 0; k < N; ++k)
        probs[k] = state[2 * k] * state[2 * k] + state[2 * k + 1] * state[2 * k + 1];
}

/* test_accept (classical_mcmc_routines.py:28-41): min(1, exp(-(E'-E)/T)) > u */
ORC_API int orc_test_accept(double e_s, double e_sp, double temperature, double u) {
    const double delta = e_sp - e_s;
    double f = exp(-delta / temperature);
    if (f > 1.0) f = 1.0;
    return f > u;
}

/* Pick a mixer group: np.random.choice(mixers, p) (mixers.py:150) restated as inverse-CDF
 * over the cumulative group probabilities with the explicit uniform u. */
static int pick_group(int n_groups, const double *group_prob, double u) {
    if (n_groups <= 1 || !group_prob) return 0;
    double c = 0.0;
    for (int g = 0; g < n_groups; ++g) {
        c += group_prob[g];
        if (u < c) return g;
    }
    return n_groups - 1;
}

/* quantum_enhanced_mcmc (quantum_mcmc_routines.py:159-227) for one chain with the Philox
 * stream contract.  gatewise != 0 uses the gate-at-a-time simulator (reference cost).
 * Mixer: groups [group_off[g], group_off[g+1]) of (mask, weight) terms; group_prob NULL or
 * n_groups==1 => a single (coherent) mixer; else IncoherentMixerSum (one group per proposal).
 * Outputs: proposals[H], accepted[H] (the MCMCState fields, basic_utils.py:54-57).
 * Returns the number of accepted proposals. */
ORC_API int64_t orc_run_chain(int n, const double *J, const double *h, const double *circJ,
                              const double *circh, double alpha, int n_groups,
                              const int *group_off, const uint64_t *masks, const double *weights,
                              const double *group_prob, uint64_t s0, int64_t H, uint32_t hop0,
                              double temperature, double g_lo, double g_hi, double delta_time,
                              uint64_t seed, uint64_t chain_id, int gatewise, uint64_t *proposals,
                              uint8_t *accepted) {
    const int64_t N = 1LL << n;
    double *state = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    uint64_t cur = s0;
    double e_cur = orc_energy(n, J, h, cur);
    int64_t nacc = 0;
    for (int64_t hop = 0; hop < H; ++hop) {
        double ug, ugrp, um, ua;
        int t;
        orc_hop_draws(seed, chain_id, hop0 + (uint32_t)hop, &ug, &t, &ugrp, &um, &ua);
        const double gamma = g_lo + (g_hi - g_lo) * ug;
        const int r = orc_trotter_steps(t, delta_time);
        const int g = pick_group(n_groups, group_prob, ugrp);
        const int o0 = group_off ? group_off[g] : 0;
        const int o1 = group_off ? group_off[g + 1] : 0;
        if (gatewise)
            orc_evolve_gatewise(n, circJ, circh, alpha, gamma, delta_time, r, o1 - o0, masks + o0,
                                weights + o0, cur, state);
        else
            orc_evolve_fused(n, circJ, circh, alpha, gamma, delta_time, r, o1 - o0, masks + o0,
                             weights + o0, cur, state);
        const uint64_t sp = orc_sample_index(n, state, um);
        const double e_sp = orc_energy(n, J, h, sp);
        const int acc = orc_test_accept(e_cur, e_sp, temperature, ua);
        proposals[hop] = sp;
        accepted[hop] = (uint8_t)acc;
        if (acc) {
            cur = sp;
            e_cur = orc_energy(n, J, h, cur); /* :222 recomputes */
            ++nacc;
        }
    }
    free(state);
    return nacc;
}

/* ------------------------------------------------------------------------- */
/* Exact Boltzmann enumeration (energy_models.py:218-275)                      */
/* ------------------------------------------------------------------------- */

/* Reference arithmetic: w_k = exp(-beta E_k); Z = sum w; p_k = w_k * (1/Z).
 * probs (nullable) receives p in index order; logZ also via a max-shifted sum so it stays
 * finite where the reference overflows. */
ORC_API double orc_exact_boltzmann(int n, const double *J, const double *h, double beta,
                                   double *energies /* nullable */, double *probs /* nullable */) {
    const int64_t N = 1LL << n;
    double *E = energies ? energies : (double *)malloc(sizeof(double) * (size_t)N);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < N; ++k) E[k] = orc_energy(n, J, h, (uint64_t)k);
    double m = -INFINITY;
    for (int64_t k = 0; k < N; ++k)
        if (-beta * E[k] > m) m = -beta * E[k];
    double s = 0.0;
    for (int64_t k = 0; k < N; ++k) s += exp(-beta * E[k] - m);
    const double logZ = m + log(s);
    if (probs) {
        double Z = 0.0;
        for (int64_t k = 0; k < N; ++k) { probs[k] = exp(-1.0 * beta * E[k]); Z += probs[k]; }
        const double inv = 1.0 / Z;
        for (int64_t k = 0; k < N; ++k) probs[k] = probs[k] * inv;
    }
    if (!energies) free(E);
    return logZ;
}

ORC_API void orc_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * classical_mcmc (qumcmc/classical_mcmc_routines.py:44-99) with the proposal rules of
 * qumcmc/classical_mixers.py:27-101, under the classical draw contract of include/qumcmc_b200.h:
 *   slot 2: w0,w1 -> u_method (CombineProposals' np.random.choice(p), :99), w2,w3 -> u_accept (:41)
 *   slot 3..: UniformProposals (:38-40, get_random_state): (w1<<32|w0) & (2^n-1);
 *             FixedWeightProposals (:56-59, random_bstr basic_utils.py:383-388): k distinct string positions by a
 *             partial Fisher-Yates shuffle, step i uses word (i&3) of slot 3+(i>>2); string position p <-> bit n-1-p;
 *             CustomProposals (:72-78): xor with the fixed mask.
 * kind: 0 uniform, 1 fixed weight (param = k), 2 custom (param = mask over index bits).  cum = cumulative
 * probabilities.  Returns the number of accepted proposals. */
static uint64_t classical_propose(int n, int n_methods, const int *kind, const uint64_t *param, const double *cum,
                                  uint64_t cur, uint64_t seed, uint64_t chain, uint32_t hop, double u_method) {
    int m = n_methods - 1;
    for (int i = 0; i < n_methods; ++i)
        if (u_method < cum[i]) { m = i; break; }
    uint32_t w[4];
    if (kind[m] == 0) {
        orc_philox4x32(seed, chain, hop, 3, w);
        const uint64_t v = ((uint64_t)w[1] << 32) | w[0];
        return n >= 64 ? v : (v & ((1ULL << n) - 1ULL));
    }
    if (kind[m] == 2) return cur ^ param[m];
    int pos[64];
    for (int i = 0; i < n; ++i) pos[i] = i;
    uint64_t flips = 0;
    const int k = (int)param[m];
    for (int i = 0; i < k; ++i) {
        if ((i & 3) == 0) orc_philox4x32(seed, chain, hop, 3 + (uint32_t)(i >> 2), w);
        const int j = i + (int)(((uint64_t)w[i & 3] * (uint64_t)(n - i)) >> 32);
        const int t = pos[i]; pos[i] = pos[j]; pos[j] = t;
        flips |= 1ULL << (n - 1 - pos[i]);
    }
    return cur ^ flips;
}

ORC_API int64_t orc_classical_chain(int n, const double *J, const double *h, int n_methods, const int *kind,
                                    const uint64_t *param, const double *cum, uint64_t s0, int64_t H,
                                    uint32_t hop0, double temperature, uint64_t seed, uint64_t chain_id,
                                    uint64_t *proposals, uint8_t *accepted) {
    uint64_t cur = s0;
    double e_cur = orc_energy(n, J, h, cur);
    int64_t nacc = 0;
    for (int64_t hop = 0; hop < H; ++hop) {
        uint32_t w[4];
        orc_philox4x32(seed, chain_id, hop0 + (uint32_t)hop, 2, w);
        const uint64_t sp = classical_propose(n, n_methods, kind, param, cum, cur, seed, chain_id,
                                              hop0 + (uint32_t)hop, u53(w[0], w[1]));
        const double e_sp = orc_energy(n, J, h, sp);
        const int acc = orc_test_accept(e_cur, e_sp, temperature, u53(w[2], w[3]));
        proposals[hop] = sp;
        accepted[hop] = (uint8_t)acc;
        if (acc) { cur = sp; e_cur = e_sp; ++nacc; }
    }
    return nacc;
}
