// common.cuh -- shared host/device helpers for libqumcmc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only (NVTX v3): ranges cost nothing unless a profiler is attached

#include "../../include/qumcmc_b200.h"

// NVTX range over a scope (SURVEY section 5 "tracing"): every C-ABI entry point and every hop of the chain drivers is a
// named range, so an nsys / ncu timeline reads in the vocabulary of the path (proposal, hop, pass, sample).
struct QmcNvtxRange {
    explicit QmcNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~QmcNvtxRange() { nvtxRangePop(); }
    QmcNvtxRange(const QmcNvtxRange &) = delete;
    QmcNvtxRange &operator=(const QmcNvtxRange &) = delete;
};
#define QMC_NVTX_CAT2(a, b) a##b
#define QMC_NVTX_CAT(a, b) QMC_NVTX_CAT2(a, b)
#define QMC_NVTX(name) QmcNvtxRange QMC_NVTX_CAT(qmc_nvtx_, __LINE__)(name)

// ---------------------------------------------------------------------------------------------
// error plumbing: no exceptions cross the ABI; message kept thread-local
// ---------------------------------------------------------------------------------------------
void qmc_set_error(const char *fmt, ...);

#define QMC_CUDA_TRY(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            qmc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return QMC_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define QMC_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            qmc_set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 -- stream contract in include/qumcmc_b200.h, mirrored by the CPU checker under oracle/
// ---------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t w[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t seed, uint64_t chain, uint32_t hop,
                                                          uint32_t slot) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)chain, c1 = (uint32_t)(chain >> 32), c2 = hop, c3 = slot;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

__host__ __device__ __forceinline__ double qmc_u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// Per-hop draws (replaces np.random.uniform(*gamma) :186, np.random.choice(range(2,12)) :125 of
// qumcmc/quantum_mcmc_routines.py, np.random.choice(mixers,p) mixers.py:150, qulacs' sampling
// RNG :153 and np.random.rand() classical_mcmc_routines.py:41).
struct HopDraws {
    double u_gamma, u_group, u_meas, u_acc;
    int t;
};

__host__ __device__ __forceinline__ HopDraws qmc_hop_draws(uint64_t seed, uint64_t chain, uint32_t hop) {
    HopDraws d;
    Philox4 a = philox4x32_10(seed, chain, hop, 0);
    d.u_gamma = qmc_u53(a.w[0], a.w[1]);
    d.t = 2 + (int)(((uint64_t)a.w[2] * 10u) >> 32);
    d.u_group = (double)a.w[3] * (1.0 / 4294967296.0);
    Philox4 b = philox4x32_10(seed, chain, hop, 1);
    d.u_meas = qmc_u53(b.w[0], b.w[1]);
    d.u_acc = qmc_u53(b.w[2], b.w[3]);
    return d;
}

__host__ __device__ __forceinline__ uint64_t qmc_initial_state(uint64_t seed, uint64_t chain, int n) {
    Philox4 a = philox4x32_10(seed, chain, 0xFFFFFFFFu, 0);
    const uint64_t v = ((uint64_t)a.w[1] << 32) | a.w[0];
    return n >= 64 ? v : (v & ((1ULL << n) - 1ULL));
}

// r = int(floor(t/delta_time)) (quantum_mcmc_routines.py:127); trotter() with r <= 1 still applies
// one mixer (:95-106).
__host__ __device__ __forceinline__ int qmc_trotter_steps(int t, double delta_time) {
    const int r = (int)floor((double)t / delta_time);
    return r < 1 ? 1 : r;
}

// inverse-CDF over cumulative group probabilities (IncoherentMixerSum, mixers.py:149-151)
__host__ __device__ __forceinline__ int qmc_pick_group(int n_groups, const double *group_cum, double u) {
    if (n_groups <= 1) return 0;
    for (int g = 0; g < n_groups; ++g)
        if (u < group_cum[g]) return g;
    return n_groups - 1;
}

// ---------------------------------------------------------------------------------------------
// model handle
// ---------------------------------------------------------------------------------------------
struct SweepWorkspace;    // sweep.cu
struct GeneralWorkspace;  // general.cu
struct Sweep128Workspace; // sweep128.cu

struct qmc_model {
    int n = 0;
    int device = 0;
    double alpha = 1.0;
    int phase_active = 1;  // fn_qckt_problem_half skip rule (:64-67)
    // device copies (fp64): J, h (energies) and circuit_J/h (phase layer)
    double *dJ = nullptr, *dh = nullptr;
    // phase exponent D(idx), fp64, index order (shared-memory path only)
    double *dD = nullptr;
    // energy table E(idx) for the shared-memory path (n <= QMC_SMEM_MAX_SPINS_F32)
    double *dE = nullptr;
    // mixer exponent in the Hadamard basis, E_x(x) = sum_S w_S (-1)^popcount(x & S), per group: [n_groups][2^n] fp64
    // (shared-memory path only: a group with many terms is applied as H . exp(-i gamma dt E_x) . H)
    double *dEx = nullptr;
    // mixer
    int n_groups = 0;
    std::vector<int> group_off;       // n_groups+1
    std::vector<uint64_t> masks;      // terms
    std::vector<double> weights;      // terms
    std::vector<double> group_cum;    // cumulative probabilities (n_groups)
    int *d_group_off = nullptr;
    uint64_t *d_masks = nullptr;
    double *d_weights = nullptr;
    double *d_group_cum = nullptr;
    int max_terms_in_group = 0;
    bool all_one_body = false;        // every term is a single-qubit X (a qubit may carry several terms of a group:
                                      // exp(-i a X) exp(-i b X) = exp(-i (a + b) X), their weights add up)
    std::vector<double> qweight;      // all_one_body: summed weight per (group, qubit), [n_groups][n]
    double *d_qweight = nullptr;
    // every term of every group has weight <= 2 (and not all 1): per group E_x(x) = sum_q xh[q] z_q + sum_{q<q'} xJ[q][q'] z_q z_q'
    // is the same kind of quadratic form as the phase layer, so the sweep kernels can host it (sweep.cu, XQ mode)
    bool quad_mixer = false;
    std::vector<double> xJ, xh;       // [n_groups][n][n] symmetric with zero diagonal, [n_groups][n]; index = qubit = index bit
    std::vector<double> host_J, host_h, host_cJ, host_ch;
    // fast exact enumerator: symmetrised couplings / fields in qubit order, constant 0.5 * trace(J)
    double *dJp = nullptr, *dhq_e = nullptr;
    double fast_e0 = 0.0;
    // large-n path
    SweepWorkspace *sweep = nullptr;
    // general path (arbitrary mixers / complex128 in HBM)
    GeneralWorkspace *general = nullptr;
    // complex128 sweep path (one-body mixers, 13 <= n <= 20)
    Sweep128Workspace *sweep128 = nullptr;
    int mixer_epoch = 0;  // bumped by qmc_mixer_set
    // *_host entry points: one device arena (grows, never shrinks) and one non-blocking stream owned by the handle, so
    // that the steady state of a host-buffer call is cudaMemcpyAsync + launches + one stream synchronise
    void *stage_dev = nullptr;
    size_t stage_cap = 0;
    cudaStream_t stage_stream = nullptr;
    // visit histogram accumulator (qmc_set_visit_histogram): int64[2^n], caller-owned, or null
    unsigned long long *visit_hist = nullptr;
    // stats
    int64_t launches = 0;
    int profile = 0;
    double prof_ms = 0.0;
    int64_t prof_launches = 0;
    double prof_bytes = 0.0;
    double prof_moved = 0.0;  // bytes the sweep launches actually read + wrote (first sweep write-only, last read-only)
};

// implemented in the respective .cu files
int qmc_small_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in,
                  uint64_t chain_id0, uint32_t hop0, double temperature, double g_lo, double g_hi,
                  double dt, uint64_t seed, int dtype, const double *gamma_arr, const int *r_arr,
                  const double *u_arr, const int *group_arr, uint64_t *proposals, uint8_t *accepted,
                  uint64_t *final_state, int64_t *acc_count, int64_t *hamming_hist, void *probs_out,
                  void *amps_out, cudaStream_t st);

int qmc_sweep_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in,
                  uint64_t chain_id0, uint32_t hop0, double temperature, double g_lo, double g_hi,
                  double dt, uint64_t seed, int dtype, const double *gamma_arr, const int *r_arr,
                  const double *u_arr, const int *group_arr, uint64_t *proposals, uint8_t *accepted,
                  uint64_t *final_state, int64_t *acc_count, int64_t *hamming_hist, void *probs_out,
                  void *amps_out, cudaStream_t st);
void qmc_sweep_free(qmc_model *m);

int qmc_general_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in,
                    uint64_t chain_id0, uint32_t hop0, double temperature, double g_lo, double g_hi,
                    double dt, uint64_t seed, int dtype, const double *gamma_arr, const int *r_arr,
                    const double *u_arr, const int *group_arr, uint64_t *proposals, uint8_t *accepted,
                    uint64_t *final_state, int64_t *acc_count, int64_t *hamming_hist, void *probs_out,
                    void *amps_out, cudaStream_t st);
void qmc_general_free(qmc_model *m);

int qmc_sweep128_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in, uint64_t chain_id0,
                     uint32_t hop0, double temperature, double g_lo, double g_hi, double dt, uint64_t seed,
                     const double *gamma_arr, const int *r_arr, const double *u_arr, const int *group_arr,
                     uint64_t *proposals, uint8_t *accepted, uint64_t *final_state, int64_t *acc_count,
                     int64_t *hamming_hist, void *probs_out, void *amps_out, cudaStream_t st);
void qmc_sweep128_free(qmc_model *m);
int qmc_general_num_passes(qmc_model *m, int group);

enum { QMC_MODE_CHAIN = 0, QMC_MODE_PROPOSE = 1, QMC_MODE_PROBS = 2 };

// Host-buffer staging (model.cu): carve the six chain-I/O device buffers out of the handle's arena.
struct ChainIoDev {
    uint64_t *s0 = nullptr, *prop = nullptr, *fin = nullptr;
    uint8_t *acc = nullptr;
    int64_t *cnt = nullptr, *hist = nullptr;
};
int qmc_stage_chain_io(qmc_model *m, int64_t n_chains, int64_t n_hops, bool want_s0, bool want_prop, bool want_acc,
                       bool want_cnt, bool want_hist, ChainIoDev *out, cudaStream_t *stream_out);

// One count per visit of the Markov chain (MCMCChain.get_accepted_dict, qumcmc/basic_utils.py:117-131)
__device__ __forceinline__ void qmc_visit(unsigned long long *visit_hist, uint64_t state) {
    if (visit_hist) atomicAdd(visit_hist + state, 1ULL);
}

// energy kernels (model.cu)
int qmc_launch_energies(qmc_model *m, const uint64_t *idx_dev, int64_t count, double *E_dev, cudaStream_t st);

// device-side energy with the oracle's fixed summation order (no FMA contraction)
__device__ __forceinline__ double qmc_energy_dev(int n, const double *__restrict__ J,
                                                 const double *__restrict__ h, uint64_t idx) {
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < n; ++i) {
        double row = 0.0;
        for (int j = 0; j < n; ++j) {
            const double sj = ((idx >> (n - 1 - j)) & 1ULL) ? 1.0 : -1.0;
            row = __dadd_rn(row, __dmul_rn(J[i * n + j], sj));
        }
        const double si = ((idx >> (n - 1 - i)) & 1ULL) ? 1.0 : -1.0;
        t1 = __dadd_rn(t1, __dmul_rn(si, row));
        t2 = __dadd_rn(t2, __dmul_rn(h[i], si));
    }
    return __dadd_rn(__dmul_rn(0.5, t1), t2);
}

// Metropolis rule of test_accept (classical_mcmc_routines.py:28-41)
__device__ __forceinline__ bool qmc_accept_dev(double e_s, double e_sp, double temperature, double u) {
    const double delta = e_sp - e_s;
    double f = exp(-delta / temperature);
    if (f > 1.0) f = 1.0;
    return f > u;
}
