// trajectory.cu -- trajectory statistics of recorded chains on the device.
//
// Reference: qumcmc/trajectory_processing.py:114-153 (calculate_running_kl_divergence: vectoried_KL of the empirical
//            distribution of the Markov chain after every step, O(2^n) per step on the host),
//            :183-210 (calculate_runnning_magnetisation), :213-321 (get_trajectory_statistics: energy differences,
//            Hamming distances of accepted / rejected proposals), qumcmc/prob_dist.py:149-156 (vectoried_KL, EPS = 1e-8),
//            qumcmc/basic_utils.py:84-131 (markov_chain, get_accepted_dict).
// One thread per chain walks the hops (the occurrence counts are a sequential dependency).  Running KL without a dense
// 2^n histogram per chain: with c_i = visits of state i among the first k entries of the Markov chain,
//     KL_k = S_p - [ log(EPS) (P_mask - V_k) + W_k - V_k log k ],   V_k = sum_{c_i>0} p_i,  W_k = sum_{c_i>0} p_i log c_i
// (sums over p_i > 1e-10, vectoried_KL's mask; c_i / k >= 1/k > EPS for k < 1e8), so one visit changes V, W by O(1)
// terms; the counts live in a per-chain open-addressing hash table keyed by the state (<= H + 1 distinct keys).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr double KL_EPS = 1e-8;        // prob_dist.py: EPS
constexpr double KL_MASK = 1e-10;      // vectoried_KL: target_vector > 1e-10
constexpr unsigned long long EMPTY_KEY = ~0ULL;

// S_p = sum p log p and P_mask = sum p over p > 1e-10: per-block partials in a fixed order, finished by one thread
__global__ void __launch_bounds__(256) kl_constants_partial(const double *__restrict__ p, int64_t N, double *__restrict__ part) {
    __shared__ double s1[256], s2[256];
    const int64_t per = (N + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = blockIdx.x * per, b1 = b0 + per < N ? b0 + per : N;
    double a = 0.0, b = 0.0;
    for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
        const double x = p[i];
        if (x > KL_MASK) {
            a += x * log(x);
            b += x;
        }
    }
    s1[threadIdx.x] = a;
    s2[threadIdx.x] = b;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            s1[threadIdx.x] += s1[threadIdx.x + off];
            s2[threadIdx.x] += s2[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = s1[0];
        part[2 * blockIdx.x + 1] = s2[0];
    }
}
__global__ void kl_constants_final(const double *__restrict__ part, int nblk, double *__restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblk; ++i) {
        a += part[2 * i];
        b += part[2 * i + 1];
    }
    out[0] = a;
    out[1] = b;
}

struct TrajArgs {
    int n;
    int64_t n_chains, n_hops, chain0, batch_chains;  // this launch covers chains [chain0, chain0 + batch_chains)
    const double *J, *h;
    const uint64_t *s0, *proposals;
    const uint8_t *accepted;
    const double *probs, *klc;  // Boltzmann probabilities (index order), {S_p, P_mask}
    uint64_t *markov;
    double *ediff, *rmag, *rkl;
    int64_t *ham;
    unsigned long long *keys;   // [chains of this batch][cap]
    unsigned *counts;
    unsigned cap_mask;
};

__device__ __forceinline__ unsigned hash64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (unsigned)x;
}

__global__ void __launch_bounds__(64) trajectory_kernel(const TrajArgs a) {
    extern __shared__ double sJ[];  // J[n*n], h[n] (energy differences only)
    const int n = a.n;
    if (a.ediff) {
        for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = a.J[i];
        for (int i = threadIdx.x; i < n; i += blockDim.x) sJ[n * n + i] = a.h[i];
        __syncthreads();
    }
    const int64_t lc = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // chain inside the batch
    const int64_t c = a.chain0 + lc;
    if (lc >= a.batch_chains || c >= a.n_chains) return;  // threads past the batch must not touch the next batch's chains
    const int64_t H = a.n_hops;
    unsigned long long *keys = a.keys ? a.keys + lc * ((size_t)a.cap_mask + 1) : nullptr;
    unsigned *cnts = a.counts ? a.counts + lc * ((size_t)a.cap_mask + 1) : nullptr;
    uint64_t cur = a.s0[c];
    double e_cur = a.ediff ? qmc_energy_dev(n, sJ, sJ + n * n, cur) : 0.0;
    double V = 0.0, W = 0.0;
    const double Sp = a.rkl ? a.klc[0] : 0.0, Pm = a.rkl ? a.klc[1] : 0.0;
    long long mag_sum = 0;
    // entry k of the Markov chain, k = 0 (initial state) .. H
    for (int64_t k = 0; k <= H; ++k) {
        if (k > 0) {
            const uint64_t sp = a.proposals[c * H + k - 1];
            const bool acc = a.accepted[c * H + k - 1] != 0;
            if (a.ham) a.ham[(c * (n + 1) + __popcll(cur ^ sp)) * 2 + (acc ? 0 : 1)] += 1;
            if (a.ediff) {
                const double e_sp = qmc_energy_dev(n, sJ, sJ + n * n, sp);
                a.ediff[c * H + k - 1] = e_sp - e_cur;
                if (acc) e_cur = e_sp;
            }
            if (acc) cur = sp;
        }
        if (a.markov) a.markov[c * (H + 1) + k] = cur;
        if (a.rmag) {  // expectation of the magnetisation over markov_chain[0:k], k = 1..H  (:199-206)
            if (k > 0) a.rmag[c * H + k - 1] = (double)mag_sum / (double)k;
            mag_sum += 2 * __popcll(cur) - n;
        }
        if (a.rkl) {
            unsigned slot = hash64(cur) & a.cap_mask;
            while (keys[slot] != EMPTY_KEY && keys[slot] != cur) slot = (slot + 1) & a.cap_mask;
            const double p = a.probs[cur];
            if (keys[slot] == EMPTY_KEY) {
                keys[slot] = cur;
                cnts[slot] = 1;
                if (p > KL_MASK) V += p;
            } else {
                const unsigned cc = cnts[slot];
                cnts[slot] = cc + 1;
                if (p > KL_MASK) W += p * (log((double)(cc + 1)) - log((double)cc));
            }
            const double T = log(KL_EPS) * (Pm - V) + W - V * log((double)(k + 1));
            a.rkl[c * (H + 1) + k] = Sp - T;
        }
    }
}

}  // namespace

extern "C" int qmc_trajectory_stats(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_dev,
                                    const uint64_t *proposals_dev, const uint8_t *accepted_dev,
                                    const double *boltz_probs_dev, uint64_t *markov_chain_dev, double *energy_diff_dev,
                                    double *running_mag_dev, int64_t *hamming_acc_rej_dev, double *running_kl_dev,
                                    void *stream) {
    QMC_NVTX("qmc_trajectory_stats");
    QMC_REQUIRE(m && s0_dev && (n_hops == 0 || (proposals_dev && accepted_dev)), QMC_ERR_BADARG,
                "qmc_trajectory_stats: null argument");
    QMC_REQUIRE(n_chains >= 0 && n_hops >= 0, QMC_ERR_BADARG, "negative chain or hop count");
    QMC_REQUIRE(!running_kl_dev || boltz_probs_dev, QMC_ERR_BADARG, "running KL needs the Boltzmann probabilities");
    QMC_REQUIRE(!running_kl_dev || n_hops < 99999999, QMC_ERR_UNSUPPORTED, "running KL: n_hops must stay below 1/EPS = 1e8");
    QMC_REQUIRE(!running_kl_dev || m->n <= 40, QMC_ERR_UNSUPPORTED, "running KL needs a 2^n probability table");
    if (n_chains == 0) return QMC_OK;
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = m->n;
    TrajArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.n_chains = n_chains;
    a.n_hops = n_hops;
    a.J = m->dJ;
    a.h = m->dh;
    a.s0 = s0_dev;
    a.proposals = proposals_dev;
    a.accepted = accepted_dev;
    a.probs = boltz_probs_dev;
    a.markov = markov_chain_dev;
    a.ediff = energy_diff_dev;
    a.rmag = running_mag_dev;
    a.rkl = running_kl_dev;
    a.ham = hamming_acc_rej_dev;
    if (hamming_acc_rej_dev)
        QMC_CUDA_TRY(cudaMemsetAsync(hamming_acc_rej_dev, 0, sizeof(int64_t) * 2 * (size_t)n_chains * (size_t)(n + 1), st));
    double *klc = nullptr, *part = nullptr;
    unsigned long long *keys = nullptr;
    unsigned *counts = nullptr;
    int64_t batch = n_chains;
    int rc = QMC_OK;
    cudaError_t e = cudaSuccess;
    if (running_kl_dev) {
        const int nblk = 256;
        e = cudaMalloc(&part, sizeof(double) * 2 * (nblk + 1));
        if (e == cudaSuccess) {
            klc = part + 2 * nblk;
            kl_constants_partial<<<nblk, 256, 0, st>>>(boltz_probs_dev, (int64_t)1 << n, part);
            kl_constants_final<<<1, 32, 0, st>>>(part, nblk, klc);
            m->launches += 2;
            size_t cap = 64;
            while (cap < 2 * (size_t)(n_hops + 1)) cap <<= 1;
            const size_t per_chain = cap * (sizeof(unsigned long long) + sizeof(unsigned));
            size_t budget = (size_t)4 << 30;  // hash tables of one batch of chains
            if (const char *be = getenv("QUMCMC_TRAJ_BUDGET_BYTES")) budget = (size_t)strtoull(be, nullptr, 10);  // tests: force several batches
            batch = (int64_t)std::max<size_t>(1, std::min<size_t>((size_t)n_chains, budget / per_chain));
            e = cudaMalloc(&keys, sizeof(unsigned long long) * cap * (size_t)batch);
            if (e == cudaSuccess) e = cudaMalloc(&counts, sizeof(unsigned) * cap * (size_t)batch);
            a.cap_mask = (unsigned)(cap - 1);
            a.klc = klc;
            a.keys = keys;
            a.counts = counts;
        }
    }
    const size_t smem = energy_diff_dev ? sizeof(double) * (size_t)(n * n + n) : 0;
    for (int64_t c0 = 0; c0 < n_chains && e == cudaSuccess; c0 += batch) {
        const int64_t nc = std::min<int64_t>(batch, n_chains - c0);
        if (keys) e = cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * ((size_t)a.cap_mask + 1) * (size_t)nc, st);
        if (e != cudaSuccess) break;
        a.chain0 = c0;
        a.batch_chains = nc;
        trajectory_kernel<<<(unsigned)((nc + 63) / 64), 64, smem, st>>>(a);
        m->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && (keys || part)) e = cudaStreamSynchronize(st);  // scratch is freed below
    cudaFree(part);
    cudaFree(keys);
    cudaFree(counts);
    if (e != cudaSuccess) {
        qmc_set_error("qmc_trajectory_stats: %s", cudaGetErrorString(e));
        return QMC_ERR_CUDA;
    }
    return rc;
}

// ------------------------------------------------------------------------------------------
// Spin moments of a set of states: the model expectations of the contrastive-divergence
// gradient, cd_J / cd_h of qumcmc/training.py:70-99 (<s_i s_j>, <s_i> over mcmc_chain.accepted_states), as exact
// integer counts:  ones[i] = #states with s_i = +1,  agree[i][j] = #states with s_i == s_j   (spin i <-> index bit
// n-1-i), so that <s_i> = (2 ones_i - N) / N and <s_i s_j> = (2 agree_ij - N) / N.
// grid = (state chunks, n + 1 rows): row i < n fixes spin i, row n counts the ones.
// ------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) state_moments_kernel(int n, int64_t count, const uint64_t *__restrict__ states,
                                                            const uint8_t *__restrict__ mask,
                                                            unsigned long long *__restrict__ used,
                                                            unsigned long long *__restrict__ ones,
                                                            unsigned long long *__restrict__ agree) {
    const int row = blockIdx.y;
    unsigned cnt[QMC_MAX_SPINS];
#pragma unroll
    for (int j = 0; j < QMC_MAX_SPINS; ++j) cnt[j] = 0u;
    unsigned nused = 0u;
    // chunks of at most 2^31 states per thread keep the 32-bit counters exact
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[k]) continue;
        const uint64_t x = states[k];
        // bit b of y is 1 iff the spin at bit b agrees with spin `row` (row n: iff the spin is +1)
        const uint64_t y = (row < n && !((x >> (n - 1 - row)) & 1ULL)) ? ~x : x;
#pragma unroll
        for (int j = 0; j < QMC_MAX_SPINS; ++j) cnt[j] += (unsigned)((y >> j) & 1ULL);
        ++nused;
    }
    __shared__ unsigned long long red[QMC_MAX_SPINS + 1];
    if (threadIdx.x <= QMC_MAX_SPINS) red[threadIdx.x] = 0ULL;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < QMC_MAX_SPINS; ++j) {
        unsigned v = cnt[j];
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && j < n) atomicAdd(&red[j], (unsigned long long)v);
    }
    {
        unsigned v = nused;
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[QMC_MAX_SPINS], (unsigned long long)v);
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {  // bit j <-> spin n-1-j
        const int spin = n - 1 - (int)threadIdx.x;
        if (row < n) atomicAdd(&agree[row * n + spin], red[threadIdx.x]);
        else atomicAdd(&ones[spin], red[threadIdx.x]);
    }
    if (threadIdx.x == 0 && row == n) atomicAdd(used, red[QMC_MAX_SPINS]);
}

}  // namespace

extern "C" int qmc_state_moments(qmc_model *m, int64_t count, const uint64_t *states_dev, const uint8_t *mask_dev,
                                 int64_t *n_used_dev, int64_t *ones_dev, int64_t *agree_dev, void *stream) {
    QMC_REQUIRE(m && n_used_dev && ones_dev && agree_dev && (count == 0 || states_dev), QMC_ERR_BADARG,
                "qmc_state_moments: null argument");
    QMC_REQUIRE(count >= 0, QMC_ERR_BADARG, "negative state count");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = m->n;
    QMC_CUDA_TRY(cudaMemsetAsync(n_used_dev, 0, sizeof(int64_t), st));
    QMC_CUDA_TRY(cudaMemsetAsync(ones_dev, 0, sizeof(int64_t) * n, st));
    QMC_CUDA_TRY(cudaMemsetAsync(agree_dev, 0, sizeof(int64_t) * n * n, st));
    if (count == 0) return QMC_OK;
    int64_t blocks = (count + 256 * 64 - 1) / (256 * 64);  // ~64 states per thread
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    state_moments_kernel<<<dim3((unsigned)blocks, (unsigned)(n + 1)), 256, 0, st>>>(
        n, count, states_dev, mask_dev, reinterpret_cast<unsigned long long *>(n_used_dev),
        reinterpret_cast<unsigned long long *>(ones_dev), reinterpret_cast<unsigned long long *>(agree_dev));
    m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}
