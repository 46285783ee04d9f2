// abi.cu -- extern "C" entry points of the proposal path; dispatch between the shared-memory
// kernel (small n) and the HBM sweep kernels (large n).  See include/qumcmc_b200.h.
#include <stdlib.h>

#include "common.cuh"

static bool use_small_path(const qmc_model *m, int dtype) {
    const int lim = dtype == QMC_DTYPE_F64 ? QMC_SMEM_MAX_SPINS_F64 : QMC_SMEM_MAX_SPINS_F32;
    if (m->n > lim) return false;
    const size_t esz = dtype == QMC_DTYPE_F64 ? 16 : 8;
    const size_t smem = (2 * ((size_t)1 << m->n) + (size_t)m->max_terms_in_group) * esz + sizeof(int) * (size_t)m->max_terms_in_group;
    return smem <= 227 * 1024;
}

static int dispatch(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in,
                    uint64_t chain_id0, uint32_t hop0, double temperature, double g_lo, double g_hi, double dt,
                    uint64_t seed, int dtype, const double *gamma_arr, const int *r_arr, const double *u_arr,
                    const int *group_arr, uint64_t *proposals, uint8_t *accepted, uint64_t *final_state,
                    int64_t *acc_count, int64_t *hamming_hist, void *probs_out, void *amps_out, void *stream) {
    QMC_REQUIRE(m, QMC_ERR_BADARG, "null model handle");
    QMC_REQUIRE(dtype == QMC_DTYPE_F32 || dtype == QMC_DTYPE_F64, QMC_ERR_BADARG, "dtype must be QMC_DTYPE_F32/F64");
    QMC_REQUIRE(n_chains >= 0 && n_hops >= 0, QMC_ERR_BADARG, "negative chain or hop count");
    QMC_REQUIRE(dt > 0.0, QMC_ERR_BADARG, "delta_time must be positive");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (use_small_path(m, dtype))
        return qmc_small_run(m, mode, n_chains, n_hops, s_in, chain_id0, hop0, temperature, g_lo, g_hi, dt, seed,
                             dtype, gamma_arr, r_arr, u_arr, group_arr, proposals, accepted, final_state,
                             acc_count, hamming_hist, probs_out, amps_out, st);
    // specialised sweeps: complex64 + one-body X mixers (every qubit at most once per group)
    if (dtype == QMC_DTYPE_F32 && m->all_one_body && m->n >= 14)
        return qmc_sweep_run(m, mode, n_chains, n_hops, s_in, chain_id0, hop0, temperature, g_lo, g_hi, dt, seed, dtype,
                             gamma_arr, r_arr, u_arr, group_arr, proposals, accepted, final_state, acc_count,
                             hamming_hist, probs_out, amps_out, st);
    // complex64 + one coherent group of one- and two-qubit X strings (TwoBodyMixer, pair CustomMixer, coherent sums of
    // them), 14 <= n <= 20: the mixer is a quadratic diagonal in the Hadamard basis, hosted by the same sweep kernels
    // ... and every other X-string mixer at those sizes with the mixer exponent tabulated (XT mode; QUMCMC_NO_XT=1: general path)
    if (dtype == QMC_DTYPE_F32 && !m->all_one_body && m->n >= 14 && m->n <= 20 && !getenv("QUMCMC_NO_XQ") &&
        (m->quad_mixer || !getenv("QUMCMC_NO_XT")))
        return qmc_sweep_run(m, mode, n_chains, n_hops, s_in, chain_id0, hop0, temperature, g_lo, g_hi, dt, seed, dtype,
                             gamma_arr, r_arr, u_arr, group_arr, proposals, accepted, final_state, acc_count,
                             hamming_hist, probs_out, amps_out, st);
    // complex128 + one-body X mixers, 13 <= n <= 20: fp64 sweeps (sweep128.cu)
    if (dtype == QMC_DTYPE_F64 && m->all_one_body && m->n >= 13 && m->n <= QMC_SWEEP_F64_MAX_SPINS && !getenv("QUMCMC_NO_SWEEP128"))
        return qmc_sweep128_run(m, mode, n_chains, n_hops, s_in, chain_id0, hop0, temperature, g_lo, g_hi, dt, seed,
                                gamma_arr, r_arr, u_arr, group_arr, proposals, accepted, final_state, acc_count,
                                hamming_hist, probs_out, amps_out, st);
    // the same in complex128, 13 <= n <= 20 (sweep128.cu, XQ mode)
    if (dtype == QMC_DTYPE_F64 && !m->all_one_body && m->n >= 13 && m->n <= QMC_SWEEP_F64_MAX_SPINS && !getenv("QUMCMC_NO_XQ") &&
        (m->quad_mixer || !getenv("QUMCMC_NO_XT")) && !getenv("QUMCMC_NO_SWEEP128"))
        return qmc_sweep128_run(m, mode, n_chains, n_hops, s_in, chain_id0, hop0, temperature, g_lo, g_hi, dt, seed,
                                gamma_arr, r_arr, u_arr, group_arr, proposals, accepted, final_state, acc_count,
                                hamming_hist, probs_out, amps_out, st);
    // everything else: tiled general path (k-body / custom / summed mixers, complex128)
    return qmc_general_run(m, mode, n_chains, n_hops, s_in, chain_id0, hop0, temperature, g_lo, g_hi, dt, seed, dtype,
                           gamma_arr, r_arr, u_arr, group_arr, proposals, accepted, final_state, acc_count,
                           hamming_hist, probs_out, amps_out, st);
}

// single-element device staging for the transition_probs hook
struct ProbeArgs {
    uint64_t s;
    double gamma;
    int r;
    int group;
};

extern "C" int qmc_transition_probs(qmc_model *m, uint64_t s, double gamma, int r, double delta_time, int group,
                                    int dtype, void *probs_dev, void *amps_dev, void *stream) {
    QMC_NVTX("qmc_transition_probs");
    QMC_REQUIRE(m && (probs_dev || amps_dev), QMC_ERR_BADARG, "qmc_transition_probs: null argument");
    QMC_REQUIRE(m->n >= 64 || s < (1ULL << m->n), QMC_ERR_BADARG, "qmc_transition_probs: state out of range");
    QMC_REQUIRE(group >= 0 && group < m->n_groups, QMC_ERR_BADARG, "qmc_transition_probs: group %d of %d", group, m->n_groups);
    QMC_REQUIRE(r >= 0, QMC_ERR_BADARG, "qmc_transition_probs: negative r");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    ProbeArgs h = {s, gamma, r, group};
    ProbeArgs *d = nullptr;
    QMC_CUDA_TRY(cudaMallocAsync(&d, sizeof(ProbeArgs), st));
    QMC_CUDA_TRY(cudaMemcpyAsync(d, &h, sizeof(ProbeArgs), cudaMemcpyHostToDevice, st));
    QMC_CUDA_TRY(cudaStreamSynchronize(st));  // h is a stack object
    int rc = dispatch(m, QMC_MODE_PROBS, 1, 1, &d->s, 0, 0, 1.0, gamma, gamma, delta_time, 0, dtype, &d->gamma, &d->r,
                      nullptr, &d->group, nullptr, nullptr, nullptr, nullptr, nullptr, probs_dev, amps_dev, stream);
    cudaFreeAsync(d, st);
    return rc;
}

extern "C" int qmc_propose(qmc_model *m, int64_t n_chains, const uint64_t *s_in_dev, const double *gamma_dev,
                           const int *r_dev, const double *u_dev, const int *group_dev, double delta_time, int dtype,
                           uint64_t *s_out_dev, void *stream) {
    QMC_NVTX("qmc_propose");
    QMC_REQUIRE(m && (n_chains == 0 || (s_in_dev && gamma_dev && r_dev && u_dev && s_out_dev)), QMC_ERR_BADARG,
                "qmc_propose: null argument");
    return dispatch(m, QMC_MODE_PROPOSE, n_chains, 1, s_in_dev, 0, 0, 1.0, 0.0, 0.0, delta_time, 0, dtype, gamma_dev,
                    r_dev, u_dev, group_dev, s_out_dev, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int qmc_run_chains(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_dev,
                              uint64_t chain_id0, uint32_t hop0, double temperature, double gamma_lo, double gamma_hi,
                              double delta_time, uint64_t seed, int dtype, uint64_t *proposals_dev,
                              uint8_t *accepted_dev, uint64_t *final_state_dev, int64_t *acc_count_dev,
                              int64_t *hamming_hist_dev, void *stream) {
    QMC_NVTX("qmc_run_chains");
    QMC_REQUIRE(m && final_state_dev, QMC_ERR_BADARG, "qmc_run_chains: final_state_dev is required");
    QMC_REQUIRE(temperature > 0.0, QMC_ERR_BADARG, "qmc_run_chains: temperature must be positive");
    return dispatch(m, QMC_MODE_CHAIN, n_chains, n_hops, s0_dev, chain_id0, hop0, temperature, gamma_lo, gamma_hi,
                    delta_time, seed, dtype, nullptr, nullptr, nullptr, nullptr, proposals_dev, accepted_dev,
                    final_state_dev, acc_count_dev, hamming_hist_dev, nullptr, nullptr, stream);
}

// Host-buffer variant.  Steady state (same or smaller shapes than an earlier call): no allocation, no device-wide
// synchronisation -- one cudaMemcpyAsync per buffer on the handle's own stream around the device-pointer call, one
// stream synchronise.  Pinned host buffers make the copies truly asynchronous; pageable ones are staged by the runtime.
extern "C" int qmc_run_chains_host(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_host,
                                   uint64_t chain_id0, uint32_t hop0, double temperature, double gamma_lo,
                                   double gamma_hi, double delta_time, uint64_t seed, int dtype,
                                   uint64_t *proposals_host, uint8_t *accepted_host, uint64_t *final_state_host,
                                   int64_t *acc_count_host, int64_t *hamming_hist_host) {
    QMC_NVTX("qmc_run_chains_host");
    QMC_REQUIRE(m && final_state_host, QMC_ERR_BADARG, "qmc_run_chains_host: final_state_host is required");
    QMC_REQUIRE(n_chains >= 0 && n_hops >= 0, QMC_ERR_BADARG, "negative chain or hop count");
    if (n_chains == 0) return QMC_OK;
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    const size_t C = (size_t)n_chains, CH = C * (size_t)n_hops, hb = C * (size_t)(m->n + 1);
    ChainIoDev d;
    cudaStream_t st = nullptr;
    int rc = qmc_stage_chain_io(m, n_chains, n_hops, s0_host != nullptr, proposals_host != nullptr, accepted_host != nullptr,
                                acc_count_host != nullptr, hamming_hist_host != nullptr, &d, &st);
    if (rc != QMC_OK) return rc;
    if (d.s0) QMC_CUDA_TRY(cudaMemcpyAsync(d.s0, s0_host, 8 * C, cudaMemcpyHostToDevice, st));
    rc = qmc_run_chains(m, n_chains, n_hops, d.s0, chain_id0, hop0, temperature, gamma_lo, gamma_hi, delta_time, seed,
                        dtype, d.prop, d.acc, d.fin, d.cnt, d.hist, st);
    if (rc != QMC_OK) {
        cudaStreamSynchronize(st);
        return rc;
    }
    if (d.prop && CH) QMC_CUDA_TRY(cudaMemcpyAsync(proposals_host, d.prop, 8 * CH, cudaMemcpyDeviceToHost, st));
    if (d.acc && CH) QMC_CUDA_TRY(cudaMemcpyAsync(accepted_host, d.acc, CH, cudaMemcpyDeviceToHost, st));
    QMC_CUDA_TRY(cudaMemcpyAsync(final_state_host, d.fin, 8 * C, cudaMemcpyDeviceToHost, st));
    if (d.cnt) QMC_CUDA_TRY(cudaMemcpyAsync(acc_count_host, d.cnt, 8 * C, cudaMemcpyDeviceToHost, st));
    if (d.hist) QMC_CUDA_TRY(cudaMemcpyAsync(hamming_hist_host, d.hist, 8 * hb, cudaMemcpyDeviceToHost, st));
    QMC_CUDA_TRY(cudaStreamSynchronize(st));
    return QMC_OK;
}
