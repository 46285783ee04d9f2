// classical.cu -- classical_mcmc on the device: n_chains independent Metropolis chains with
// the reference's classical proposal rules, sharing K4 (fixed-order fp64 energy, test_accept, Philox streams) with the
// quantum path.
//
// Reference: qumcmc/classical_mcmc_routines.py:44-99 (hop loop), :28-41 (test_accept);
//            qumcmc/classical_mixers.py:27-101 (UniformProposals, FixedWeightProposals, CustomProposals, CombineProposals);
//            qumcmc/basic_utils.py:383-388 (random_bstr: k ones shuffled), :395-404 (get_random_state).
// One thread per chain (the hop loop is a sequential Markov dependency; the energy of one state is n^2 fp64
// operations); J and h are staged in shared memory.  Draws: include/qumcmc_b200.h, "classical draws".
#include <math.h>

#include "common.cuh"

namespace {

struct ClassicalArgs {
    int n, n_methods;
    int kind[QMC_MAX_CLASSICAL_METHODS];
    unsigned long long param[QMC_MAX_CLASSICAL_METHODS];
    double cum[QMC_MAX_CLASSICAL_METHODS];
    const double *J, *h;
    int64_t n_chains, n_hops;
    const uint64_t *s0;
    uint64_t chain_id0, seed;
    uint32_t hop0;
    double temperature;
    uint64_t *proposals;
    uint8_t *accepted;
    uint64_t *final_state;
    int64_t *acc_count, *hist;
    unsigned long long *visit;  // visit histogram accumulator or null
};

// proposal of one hop; string position i <-> index bit n-1-i (bitstrings are MSB first)
__device__ __forceinline__ uint64_t classical_propose(const ClassicalArgs &a, uint64_t cur, uint64_t chain, uint32_t hop,
                                                      double u_method) {
    int m = a.n_methods - 1;
    for (int i = 0; i < a.n_methods; ++i)
        if (u_method < a.cum[i]) {
            m = i;
            break;
        }
    const int n = a.n;
    if (a.kind[m] == QMC_CLASSICAL_UNIFORM) {
        const Philox4 w = philox4x32_10(a.seed, chain, hop, 3);
        const uint64_t v = ((uint64_t)w.w[1] << 32) | w.w[0];
        return n >= 64 ? v : (v & ((1ULL << n) - 1ULL));
    }
    if (a.kind[m] == QMC_CLASSICAL_CUSTOM) return cur ^ a.param[m];
    // FIXED_WEIGHT: k distinct positions by a partial Fisher-Yates shuffle
    const int k = (int)a.param[m];
    unsigned char pos[QMC_MAX_SPINS];
    for (int i = 0; i < n; ++i) pos[i] = (unsigned char)i;
    uint64_t flips = 0ULL;
    Philox4 w;
    for (int i = 0; i < k; ++i) {
        if ((i & 3) == 0) w = philox4x32_10(a.seed, chain, hop, 3 + (uint32_t)(i >> 2));
        const int j = i + (int)(((uint64_t)w.w[i & 3] * (uint64_t)(n - i)) >> 32);
        const unsigned char t = pos[i];
        pos[i] = pos[j];
        pos[j] = t;
        flips |= 1ULL << (n - 1 - pos[i]);
    }
    return cur ^ flips;
}

__global__ void __launch_bounds__(128) classical_chain_kernel(const ClassicalArgs a) {
    extern __shared__ double sJ[];  // J[n*n], h[n]
    const int n = a.n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = a.J[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sJ[n * n + i] = a.h[i];
    __syncthreads();
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= a.n_chains) return;
    const double *J = sJ, *h = sJ + n * n;
    const uint64_t chain = a.chain_id0 + (uint64_t)c;
    uint64_t cur = a.s0 ? a.s0[c] : qmc_initial_state(a.seed, chain, n);
    double e_cur = qmc_energy_dev(n, J, h, cur);
    int64_t nacc = 0;
    if (a.hop0 == 0) qmc_visit(a.visit, cur);  // entry 0 of markov_chain
    for (int64_t k = 0; k < a.n_hops; ++k) {
        const uint32_t hop = a.hop0 + (uint32_t)k;
        const Philox4 d = philox4x32_10(a.seed, chain, hop, 2);
        const uint64_t sp = classical_propose(a, cur, chain, hop, qmc_u53(d.w[0], d.w[1]));
        const double e_sp = qmc_energy_dev(n, J, h, sp);
        const bool acc = qmc_accept_dev(e_cur, e_sp, a.temperature, qmc_u53(d.w[2], d.w[3]));
        if (a.proposals) a.proposals[c * a.n_hops + k] = sp;
        if (a.accepted) a.accepted[c * a.n_hops + k] = (uint8_t)acc;
        if (a.hist) a.hist[c * (n + 1) + __popcll(cur ^ sp)] += 1;
        if (acc) {
            cur = sp;
            e_cur = e_sp;
            ++nacc;
        }
        qmc_visit(a.visit, cur);
    }
    a.final_state[c] = cur;
    if (a.acc_count) a.acc_count[c] = nacc;
}

}  // namespace

extern "C" int qmc_classical_chains(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_dev,
                                    uint64_t chain_id0, uint32_t hop0, double temperature, int n_methods,
                                    const int *kinds, const uint64_t *params, const double *probs, uint64_t seed,
                                    uint64_t *proposals_dev, uint8_t *accepted_dev, uint64_t *final_state_dev,
                                    int64_t *acc_count_dev, int64_t *hamming_hist_dev, void *stream) {
    QMC_NVTX("qmc_classical_chains");
    QMC_REQUIRE(m && final_state_dev && kinds && params, QMC_ERR_BADARG, "qmc_classical_chains: null argument");
    QMC_REQUIRE(n_chains >= 0 && n_hops >= 0, QMC_ERR_BADARG, "negative chain or hop count");
    QMC_REQUIRE(n_methods >= 1 && n_methods <= QMC_MAX_CLASSICAL_METHODS, QMC_ERR_BADARG,
                "qmc_classical_chains: %d proposal methods (1..%d)", n_methods, QMC_MAX_CLASSICAL_METHODS);
    QMC_REQUIRE(temperature > 0.0, QMC_ERR_BADARG, "temperature must be positive");
    ClassicalArgs a;
    memset(&a, 0, sizeof(a));
    a.n = m->n;
    a.n_methods = n_methods;
    double tot = 0.0;
    for (int i = 0; i < n_methods; ++i) {
        const double p = probs ? probs[i] : 1.0 / n_methods;
        QMC_REQUIRE(p >= 0.0, QMC_ERR_BADARG, "negative proposal probability");
        tot += p;
    }
    QMC_REQUIRE(tot > 0.0, QMC_ERR_BADARG, "proposal probabilities sum to zero");
    double run = 0.0;
    for (int i = 0; i < n_methods; ++i) {
        a.kind[i] = kinds[i];
        a.param[i] = params[i];
        run += (probs ? probs[i] : 1.0 / n_methods) / tot;
        a.cum[i] = run;
        QMC_REQUIRE(kinds[i] >= QMC_CLASSICAL_UNIFORM && kinds[i] <= QMC_CLASSICAL_CUSTOM, QMC_ERR_BADARG,
                    "unknown classical proposal kind %d", kinds[i]);
        if (kinds[i] == QMC_CLASSICAL_FIXED_WEIGHT)
            QMC_REQUIRE(params[i] <= (uint64_t)m->n, QMC_ERR_BADARG, "FixedWeightProposals: bodyness %llu > %d spins",
                        (unsigned long long)params[i], m->n);
        if (kinds[i] == QMC_CLASSICAL_CUSTOM && m->n < 64)
            QMC_REQUIRE((params[i] >> m->n) == 0, QMC_ERR_BADARG, "CustomProposals: flip mask outside %d spins", m->n);
    }
    a.cum[n_methods - 1] = 1.0;
    if (n_chains == 0) return QMC_OK;
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    a.J = m->dJ;
    a.h = m->dh;
    a.n_chains = n_chains;
    a.n_hops = n_hops;
    a.s0 = s0_dev;
    a.chain_id0 = chain_id0;
    a.seed = seed;
    a.hop0 = hop0;
    a.temperature = temperature;
    a.proposals = proposals_dev;
    a.accepted = accepted_dev;
    a.final_state = final_state_dev;
    a.acc_count = acc_count_dev;
    a.hist = hamming_hist_dev;
    a.visit = m->visit_hist;
    if (hamming_hist_dev)
        QMC_CUDA_TRY(cudaMemsetAsync(hamming_hist_dev, 0, sizeof(int64_t) * (size_t)n_chains * (size_t)(m->n + 1), st));
    const size_t smem = sizeof(double) * (size_t)(m->n * m->n + m->n);
    // 32 chains per CTA keeps every SM busy already at a few thousand chains
    const int threads = 32;
    classical_chain_kernel<<<(unsigned)((n_chains + threads - 1) / threads), threads, smem, st>>>(a);
    m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

extern "C" int qmc_classical_chains_host(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_host,
                                         uint64_t chain_id0, uint32_t hop0, double temperature, int n_methods,
                                         const int *kinds, const uint64_t *params, const double *probs, uint64_t seed,
                                         uint64_t *proposals_host, uint8_t *accepted_host, uint64_t *final_state_host,
                                         int64_t *acc_count_host, int64_t *hamming_hist_host) {
    QMC_REQUIRE(m && final_state_host, QMC_ERR_BADARG, "qmc_classical_chains_host: final_state_host is required");
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
    rc = qmc_classical_chains(m, n_chains, n_hops, d.s0, chain_id0, hop0, temperature, n_methods, kinds, params, probs, seed,
                              d.prop, d.acc, d.fin, d.cnt, d.hist, st);
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
