// model.cu -- model handle, phase/energy tables, batched energies (K4 energy part) and the
// exact Boltzmann enumerator with a grid-wide logsumexp (K5).
//
// Reference behaviour restated here:
//   IsingEnergyFunction.get_energy            qumcmc/energy_models.py:124-144
//   fn_qckt_problem_half (phase exponent D)   qumcmc/quantum_mcmc_routines.py:54-91
//   Exact_Sampling.get_boltzmann_distribution qumcmc/energy_models.py:218-275
#include <math.h>
#include <algorithm>
#include <stdarg.h>

#include "common.cuh"

static thread_local std::string g_last_error;

void qmc_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" const char *qmc_last_error(void) { return g_last_error.c_str(); }
extern "C" int qmc_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
#define QMC_MAXN 40

// E(idx[k]) for arbitrary index lists.  J/h staged in shared memory; one thread per state.
__global__ void __launch_bounds__(256) energies_kernel(int n, const double *__restrict__ J,
                                                       const double *__restrict__ h,
                                                       const uint64_t *__restrict__ idx, int64_t count,
                                                       double *__restrict__ out) {
    extern __shared__ double sm[];
    double *sJ = sm, *sh = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = J[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = h[i];
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count;
         k += (int64_t)gridDim.x * blockDim.x)
        out[k] = qmc_energy_dev(n, sJ, sh, idx[k]);
}

// Phase exponent D(idx) = sum_{j<k} J_jk z_j z_k + sum_j h_j z_j, z = -s, same loop order as the
// oracle (orc_phase_D) so the table is bit-identical.
__global__ void __launch_bounds__(256) phase_table_kernel(int n, const double *__restrict__ J,
                                                          const double *__restrict__ h,
                                                          double *__restrict__ D) {
    extern __shared__ double sm[];
    double *sJ = sm, *sh = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = J[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = h[i];
    __syncthreads();
    const int64_t N = 1LL << n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < N;
         idx += (int64_t)gridDim.x * blockDim.x) {
        double d = 0.0;
        for (int j = 0; j < n - 1; ++j) {
            const double zj = ((idx >> (n - 1 - j)) & 1LL) ? -1.0 : 1.0;
            for (int k = j + 1; k < n; ++k) {
                const double zk = ((idx >> (n - 1 - k)) & 1LL) ? -1.0 : 1.0;
                d = __dadd_rn(d, __dmul_rn(__dmul_rn(sJ[j * n + k], zj), zk));
            }
        }
        for (int j = 0; j < n; ++j) {
            const double zj = ((idx >> (n - 1 - j)) & 1LL) ? -1.0 : 1.0;
            d = __dadd_rn(d, __dmul_rn(sh[j], zj));
        }
        D[idx] = d;
    }
}

// K5: enumerate all 2^n states; per-thread online logsumexp of x = -beta*E, block tree-reduce in a
// fixed order, one (max, sumexp) partial per block.
struct LseState {
    double m, s;
};
__device__ __forceinline__ LseState lse_merge(LseState a, LseState b) {
    if (b.m == -INFINITY) return a;
    if (a.m == -INFINITY) return b;
    LseState o;
    if (a.m >= b.m) {
        o.m = a.m;
        o.s = a.s + b.s * exp(b.m - a.m);
    } else {
        o.m = b.m;
        o.s = b.s + a.s * exp(a.m - b.m);
    }
    return o;
}

__global__ void __launch_bounds__(256) exact_enumerate_kernel(int n, const double *__restrict__ J,
                                                              const double *__restrict__ h, double beta,
                                                              int64_t begin, int64_t count,
                                                              double *__restrict__ energies,
                                                              LseState *__restrict__ partials) {
    extern __shared__ double sm[];
    double *sJ = sm, *sh = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = J[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = h[i];
    __syncthreads();
    LseState acc = {-INFINITY, 0.0};
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count;
         k += (int64_t)gridDim.x * blockDim.x) {
        const double e = qmc_energy_dev(n, sJ, sh, (uint64_t)(begin + k));
        if (energies) energies[k] = e;
        const double x = -beta * e;
        if (x > acc.m) {
            acc.s = acc.s * exp(acc.m - x) + 1.0;  // exp(-inf)=0 on the first element
            acc.m = x;
        } else {
            acc.s += exp(x - acc.m);
        }
    }
    // block reduce (fixed pairing order => deterministic)
    __shared__ LseState red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] = lse_merge(red[threadIdx.x], red[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) exact_finalize_kernel(const LseState *__restrict__ partials, int nparts,
                                                             double *__restrict__ logZ, double *__restrict__ pair) {
    __shared__ LseState red[256];
    LseState acc = {-INFINITY, 0.0};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc = lse_merge(acc, partials[i]);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] = lse_merge(red[threadIdx.x], red[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (logZ) logZ[0] = red[0].m + log(red[0].s);
        if (pair) {
            pair[0] = red[0].m;
            pair[1] = red[0].s;
        }
    }
}

// K5 fast path (log Z only / probabilities only -- whenever the energies themselves are not materialised).
// Split the index into hi = idx >> 8 and lo = idx & 255 and write, in qubit order (qubit a = bit a, s_a = +-1),
//     E(idx) = A(hi) + sum_{b<8} s_b(lo) F_b(hi) + R(lo),
//     A  = e0 + sum_{a>=8} h_a s_a + sum_{8<=a<a'} Jp_aa' s_a s_a'      F_b = h_b + sum_{a>=8} Jp_ab s_a
//     R  = sum_{b<b'<8} Jp_bb' s_b s_b'        Jp = (J + J^T)/2 off the diagonal,  e0 = sum_a J_aa / 2.
// A CTA owns 2^16 consecutive indices: thread t first evaluates (A, F_0..F_7) of hi-block t into shared memory
// (O(n^2) once per 256 states), then walks the 256 hi-blocks with lo = t fixed: 8 fused adds + one exp per state,
// coalesced stores.  Same energies as the ordered evaluation up to fp64 rounding of a different summation order
// (~1e-15 relative), which is why the path that returns energies keeps the oracle's order instead.
constexpr int FAST_LO = 8, FAST_BLOCK = 1 << 16;
template <bool PROBS>
__global__ void __launch_bounds__(256) exact_fast_kernel(int n, const double *__restrict__ Jp, const double *__restrict__ hq,
                                                         double e0, double beta, int64_t begin, int64_t n_blocks,
                                                         LseState *__restrict__ partials, const double *__restrict__ logZ,
                                                         double logZ_value, double *__restrict__ probs) {
    extern __shared__ double sm[];
    double *sJ = sm, *sh = sm + n * n, *sA = sh + n, *sF = sA + 256;  // sF[b * 256 + hi_local]
    const int t = threadIdx.x;
    for (int i = t; i < n * n; i += 256) sJ[i] = Jp[i];
    for (int i = t; i < n; i += 256) sh[i] = hq[i];
    __syncthreads();
    // this thread's lo pattern: signs and the intra-lo couplings R(lo)
    double sg[FAST_LO];
#pragma unroll
    for (int b = 0; b < FAST_LO; ++b) sg[b] = ((t >> b) & 1) ? 1.0 : -1.0;
    double R = 0.0;
#pragma unroll
    for (int b = 0; b < FAST_LO; ++b)
#pragma unroll
        for (int b2 = b + 1; b2 < FAST_LO; ++b2) R += sJ[b * n + b2] * sg[b] * sg[b2];
    const double lz = PROBS ? (logZ ? logZ[0] : logZ_value) : 0.0;
    LseState acc = {-INFINITY, 0.0};
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint64_t base = (uint64_t)begin + (uint64_t)blk * FAST_BLOCK;
        {   // (A, F) of hi-block t
            const uint64_t idx = base + ((uint64_t)t << FAST_LO);
            double A = e0, F[FAST_LO];
#pragma unroll
            for (int b = 0; b < FAST_LO; ++b) F[b] = sh[b];
            for (int a = FAST_LO; a < n; ++a) {
                const double sa = ((idx >> a) & 1ULL) ? 1.0 : -1.0;
                double row = sh[a];
                for (int a2 = a + 1; a2 < n; ++a2) row += ((idx >> a2) & 1ULL) ? sJ[a * n + a2] : -sJ[a * n + a2];
                A += sa * row;
#pragma unroll
                for (int b = 0; b < FAST_LO; ++b) F[b] += sJ[b * n + a] * sa;
            }
            __syncthreads();  // previous block's readers are done
            sA[t] = A;
#pragma unroll
            for (int b = 0; b < FAST_LO; ++b) sF[b * 256 + t] = F[b];
            __syncthreads();
        }
        for (int hl = 0; hl < 256; ++hl) {
            double e = sA[hl] + R;
#pragma unroll
            for (int b = 0; b < FAST_LO; ++b) e += sg[b] * sF[b * 256 + hl];
            const double x = -beta * e;
            if (PROBS) {
                probs[(base - (uint64_t)begin) + ((uint64_t)hl << FAST_LO) + t] = exp(x - lz);
            } else if (x > acc.m) {
                acc.s = acc.s * exp(acc.m - x) + 1.0;
                acc.m = x;
            } else {
                acc.s += exp(x - acc.m);
            }
        }
    }
    if (!PROBS) {
        __shared__ LseState red[256];
        red[t] = acc;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (t < off) red[t] = lse_merge(red[t], red[t + off]);
            __syncthreads();
        }
        if (t == 0) partials[blockIdx.x] = red[0];
    }
}

// p = exp(-beta*E - logZ); E read back if materialised, recomputed otherwise.
__global__ void __launch_bounds__(256) exact_probs_kernel(int n, const double *__restrict__ J,
                                                          const double *__restrict__ h, double beta,
                                                          int64_t begin, int64_t count,
                                                          const double *__restrict__ energies,
                                                          const double *__restrict__ logZ, double logZ_value,
                                                          double *__restrict__ probs) {
    extern __shared__ double sm[];
    double *sJ = sm, *sh = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = J[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = h[i];
    __syncthreads();
    const double lz = logZ ? logZ[0] : logZ_value;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count;
         k += (int64_t)gridDim.x * blockDim.x) {
        const double e = energies ? energies[k] : qmc_energy_dev(n, sJ, sh, (uint64_t)(begin + k));
        probs[k] = exp(-beta * e - lz);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int grid_for(int64_t items, int block, int sms = 148, int per_sm = 8) {
    int64_t g = (items + block - 1) / block;
    const int64_t cap = (int64_t)sms * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// Jp / hq / e0 of the fast enumerator (qubit order: qubit a = spin n-1-a), uploaded once per model
static int ensure_fast_tables(qmc_model *m) {
    if (m->dJp) return QMC_OK;
    const int n = m->n;
    std::vector<double> Jp((size_t)n * n, 0.0), hq(n);
    double e0 = 0.0;
    for (int a = 0; a < n; ++a) {
        hq[a] = m->host_h[n - 1 - a];
        e0 += 0.5 * m->host_J[(size_t)(n - 1 - a) * n + (n - 1 - a)];
        for (int b = 0; b < n; ++b)
            if (a != b)
                Jp[(size_t)a * n + b] = 0.5 * (m->host_J[(size_t)(n - 1 - a) * n + (n - 1 - b)] + m->host_J[(size_t)(n - 1 - b) * n + (n - 1 - a)]);
    }
    m->fast_e0 = e0;
    QMC_CUDA_TRY(cudaMalloc(&m->dJp, sizeof(double) * (size_t)n * n));
    QMC_CUDA_TRY(cudaMalloc(&m->dhq_e, sizeof(double) * (size_t)n));
    QMC_CUDA_TRY(cudaMemcpy(m->dJp, Jp.data(), sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
    QMC_CUDA_TRY(cudaMemcpy(m->dhq_e, hq.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    return QMC_OK;
}
static bool fast_range_ok(const qmc_model *m, uint64_t begin, int64_t count) {
    return m->n >= 16 && count > 0 && begin % FAST_BLOCK == 0 && count % FAST_BLOCK == 0;
}
static size_t fast_smem(int n) { return sizeof(double) * (size_t)(n * n + n + 256 + FAST_LO * 256); }

int qmc_launch_energies(qmc_model *m, const uint64_t *idx_dev, int64_t count, double *E_dev, cudaStream_t st) {
    if (count == 0) return QMC_OK;
    const size_t smem = sizeof(double) * (size_t)(m->n * m->n + m->n);
    energies_kernel<<<grid_for(count, 256), 256, smem, st>>>(m->n, m->dJ, m->dh, idx_dev, count, E_dev);
    m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

extern "C" int qmc_model_create(qmc_model **out, int n, const double *J, const double *h, const double *circJ,
                                const double *circh, double alpha, int device) {
    QMC_REQUIRE(out && J && h, QMC_ERR_BADARG, "qmc_model_create: null argument");
    QMC_REQUIRE(n >= 1 && n <= QMC_MAX_SPINS, QMC_ERR_UNSUPPORTED,
                "qmc_model_create: n=%d outside [1,%d] (single-GPU statevector limit)", n, QMC_MAX_SPINS);
    int ndev = 0;
    QMC_CUDA_TRY(cudaGetDeviceCount(&ndev));
    QMC_REQUIRE(device >= 0 && device < ndev, QMC_ERR_BADARG, "qmc_model_create: device %d of %d", device, ndev);
    QMC_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    QMC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    QMC_REQUIRE(prop.major == 10, QMC_ERR_UNSUPPORTED,
                "libqumcmc_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);

    qmc_model *m = new qmc_model();
    m->n = n;
    m->device = device;
    m->alpha = alpha;
    m->host_J.assign(J, J + (size_t)n * n);
    m->host_h.assign(h, h + n);
    m->host_cJ.assign(circJ ? circJ : J, (circJ ? circJ : J) + (size_t)n * n);
    m->host_ch.assign(circh ? circh : h, (circh ? circh : h) + n);
    // skip rule of fn_qckt_problem_half (:64-67): mean|h| <= 1e-3 and mean|J| <= 1e-3
    double mh = 0.0, mJ = 0.0;
    for (int i = 0; i < n; ++i) mh += fabs(m->host_ch[i]);
    for (int i = 0; i < n * n; ++i) mJ += fabs(m->host_cJ[i]);
    mh /= (double)n;
    mJ /= (double)n * (double)n;
    m->phase_active = !(mh <= 1e-3 && mJ <= 1e-3);

    double *dcJ = nullptr, *dch = nullptr;
    auto fail = [&](int code) {
        cudaFree(dcJ);
        cudaFree(dch);
        qmc_model_destroy(m);
        return code;
    };
#define TRY_OR_FAIL(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            qmc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
            return fail(QMC_ERR_CUDA);                                                             \
        }                                                                                          \
    } while (0)
    const size_t bJ = sizeof(double) * (size_t)n * n, bh = sizeof(double) * (size_t)n;
    TRY_OR_FAIL(cudaMalloc(&m->dJ, bJ));
    TRY_OR_FAIL(cudaMalloc(&m->dh, bh));
    TRY_OR_FAIL(cudaMalloc(&dcJ, bJ));
    TRY_OR_FAIL(cudaMalloc(&dch, bh));
    TRY_OR_FAIL(cudaMemcpy(m->dJ, m->host_J.data(), bJ, cudaMemcpyHostToDevice));
    TRY_OR_FAIL(cudaMemcpy(m->dh, m->host_h.data(), bh, cudaMemcpyHostToDevice));
    TRY_OR_FAIL(cudaMemcpy(dcJ, m->host_cJ.data(), bJ, cudaMemcpyHostToDevice));
    TRY_OR_FAIL(cudaMemcpy(dch, m->host_ch.data(), bh, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double) * (size_t)(n * n + n);
    if (n <= QMC_SMEM_MAX_SPINS_F32) {  // only the shared-memory path reads the D table; the sweep paths evaluate D on the fly
        const int64_t N = 1LL << n;
        TRY_OR_FAIL(cudaMalloc(&m->dD, sizeof(double) * (size_t)N));
        phase_table_kernel<<<grid_for(N, 256), 256, smem>>>(n, dcJ, dch, m->dD);
        m->launches++;
        TRY_OR_FAIL(cudaGetLastError());
    }
    if (n <= QMC_SMEM_MAX_SPINS_F32) {
        const int64_t N = 1LL << n;
        TRY_OR_FAIL(cudaMalloc(&m->dE, sizeof(double) * (size_t)N));
        LseState *dummy = nullptr;
        TRY_OR_FAIL(cudaMalloc(&dummy, sizeof(LseState) * 1184));
        exact_enumerate_kernel<<<grid_for(N, 256), 256, smem>>>(n, m->dJ, m->dh, 0.0, 0, N, m->dE, dummy);
        m->launches++;
        TRY_OR_FAIL(cudaGetLastError());
        TRY_OR_FAIL(cudaDeviceSynchronize());
        cudaFree(dummy);
    }
    TRY_OR_FAIL(cudaDeviceSynchronize());
    cudaFree(dcJ);
    cudaFree(dch);
    dcJ = dch = nullptr;
#undef TRY_OR_FAIL
    // default mixer: GenericMixer.OneBodyMixer(n) (README-era default)
    std::vector<int> off = {0, n};
    std::vector<uint64_t> masks(n);
    std::vector<double> w(n, 1.0);
    for (int q = 0; q < n; ++q) masks[q] = 1ULL << q;
    int rc = qmc_mixer_set(m, 1, off.data(), masks.data(), w.data(), nullptr);
    if (rc != QMC_OK) {
        qmc_model_destroy(m);
        return rc;
    }
    *out = m;
    return QMC_OK;
}

extern "C" int qmc_model_destroy(qmc_model *m) {
    if (!m) return QMC_OK;
    cudaSetDevice(m->device);
    qmc_sweep_free(m);
    qmc_general_free(m);
    qmc_sweep128_free(m);
    cudaFree(m->dJ);
    cudaFree(m->dh);
    cudaFree(m->dD);
    cudaFree(m->dJp);
    cudaFree(m->dhq_e);
    cudaFree(m->dE);
    cudaFree(m->dEx);
    cudaFree(m->d_group_off);
    cudaFree(m->d_masks);
    cudaFree(m->d_weights);
    cudaFree(m->d_group_cum);
    cudaFree(m->d_qweight);
    cudaFree(m->stage_dev);
    if (m->stage_stream) cudaStreamDestroy(m->stage_stream);
    delete m;
    return QMC_OK;
}

// Device arena of the *_host entry points: [s0 | proposals | final | acc_count | hamming_hist | accepted], every part
// 256-byte aligned; reallocated only when a call needs more than any call before it.
int qmc_stage_chain_io(qmc_model *m, int64_t n_chains, int64_t n_hops, bool want_s0, bool want_prop, bool want_acc,
                       bool want_cnt, bool want_hist, ChainIoDev *out, cudaStream_t *stream_out) {
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t C = (size_t)n_chains, CH = C * (size_t)n_hops, hb = C * (size_t)(m->n + 1);
    const size_t b_s0 = want_s0 ? up(8 * C) : 0, b_prop = want_prop ? up(8 * CH) : 0, b_fin = up(8 * C);
    const size_t b_cnt = want_cnt ? up(8 * C) : 0, b_hist = want_hist ? up(8 * hb) : 0, b_acc = want_acc ? up(CH) : 0;
    const size_t need = b_s0 + b_prop + b_fin + b_cnt + b_hist + b_acc;
    if (!m->stage_stream) QMC_CUDA_TRY(cudaStreamCreateWithFlags(&m->stage_stream, cudaStreamNonBlocking));
    if (need > m->stage_cap) {
        QMC_CUDA_TRY(cudaStreamSynchronize(m->stage_stream));
        cudaFree(m->stage_dev);
        m->stage_dev = nullptr;
        m->stage_cap = 0;
        QMC_CUDA_TRY(cudaMalloc(&m->stage_dev, need));
        m->stage_cap = need;
    }
    char *p = static_cast<char *>(m->stage_dev);
    auto take = [&](size_t b) { char *q = b ? p : nullptr; p += b; return q; };
    out->s0 = reinterpret_cast<uint64_t *>(take(b_s0));
    out->prop = reinterpret_cast<uint64_t *>(take(b_prop));
    out->fin = reinterpret_cast<uint64_t *>(take(b_fin));
    out->cnt = reinterpret_cast<int64_t *>(take(b_cnt));
    out->hist = reinterpret_cast<int64_t *>(take(b_hist));
    out->acc = reinterpret_cast<uint8_t *>(take(b_acc));
    *stream_out = m->stage_stream;
    return QMC_OK;
}

extern "C" int qmc_set_visit_histogram(qmc_model *m, int64_t *visit_hist_dev) {
    QMC_REQUIRE(m, QMC_ERR_BADARG, "qmc_set_visit_histogram: null model handle");
    QMC_REQUIRE(!visit_hist_dev || m->n <= QMC_VISIT_HIST_MAX_SPINS, QMC_ERR_UNSUPPORTED,
                "visit histogram: 2^%d counters (n <= %d)", m->n, QMC_VISIT_HIST_MAX_SPINS);
    m->visit_hist = reinterpret_cast<unsigned long long *>(visit_hist_dev);
    return QMC_OK;
}

// Static-analysis hook (CPU tests): how the dispatcher classifies a mixer -- 1: every term is a single-qubit X (rotation
// kernels), 2: every term has weight <= 2 (quadratic diagonal in the Hadamard basis: XQ mode), 3: anything else (tabulated
// diagonal: XT mode / on-chip Hadamard-basis mode / general path); -1 on bad arguments.  Same rule as qmc_mixer_set.
extern "C" int qmc_debug_mixer_class(int n, int n_terms, const uint64_t *masks) {
    if (n < 1 || n > 64 || n_terms < 1 || !masks) return -1;
    const uint64_t full = n >= 64 ? ~0ULL : ((1ULL << n) - 1ULL);
    int cls = 1;
    for (int k = 0; k < n_terms; ++k) {
        if (masks[k] == 0 || (masks[k] & ~full)) return -1;
        const int wgt = __builtin_popcountll(masks[k]);
        if (wgt > 2) cls = 3;
        else if (wgt == 2 && cls < 2) cls = 2;
    }
    return cls;
}

extern "C" int qmc_mixer_set(qmc_model *m, int n_groups, const int *group_offsets, const uint64_t *masks,
                             const double *weights, const double *group_prob) {
    QMC_REQUIRE(m && group_offsets && masks && weights, QMC_ERR_BADARG, "qmc_mixer_set: null argument");
    QMC_REQUIRE(n_groups >= 1 && n_groups <= 64, QMC_ERR_BADARG, "qmc_mixer_set: n_groups=%d", n_groups);
    QMC_REQUIRE(group_offsets[0] == 0, QMC_ERR_BADARG, "qmc_mixer_set: group_offsets[0] must be 0");
    const int nt = group_offsets[n_groups];
    int max_terms = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int c = group_offsets[g + 1] - group_offsets[g];
        QMC_REQUIRE(c >= 1, QMC_ERR_BADARG, "qmc_mixer_set: group %d is empty", g);
        if (c > max_terms) max_terms = c;
    }
    const uint64_t full = (m->n >= 64) ? ~0ULL : ((1ULL << m->n) - 1ULL);
    for (int k = 0; k < nt; ++k)
        QMC_REQUIRE(masks[k] != 0 && (masks[k] & ~full) == 0, QMC_ERR_BADARG,
                    "qmc_mixer_set: term %d has an empty mask or a qubit >= n", k);
    const int mixer_class = qmc_debug_mixer_class(m->n, nt, masks);  // 1 one-body, 2 weight <= 2, 3 anything else
    const bool one_body = mixer_class == 1;
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    m->n_groups = n_groups;
    m->group_off.assign(group_offsets, group_offsets + n_groups + 1);
    m->masks.assign(masks, masks + nt);
    m->weights.assign(weights, weights + nt);
    m->group_cum.assign(n_groups, 1.0);
    if (group_prob && n_groups > 1) {
        double tot = 0.0;
        for (int g = 0; g < n_groups; ++g) {
            QMC_REQUIRE(group_prob[g] >= 0.0, QMC_ERR_BADARG, "qmc_mixer_set: negative probability");
            tot += group_prob[g];
        }
        QMC_REQUIRE(tot > 0.0, QMC_ERR_BADARG, "qmc_mixer_set: probabilities sum to zero");
        double c = 0.0;
        for (int g = 0; g < n_groups; ++g) {
            c += group_prob[g] / tot;
            m->group_cum[g] = c;
        }
        m->group_cum[n_groups - 1] = 1.0;
    } else {
        QMC_REQUIRE(n_groups == 1, QMC_ERR_BADARG,
                    "qmc_mixer_set: %d groups need group_prob (coherent sums are one group)", n_groups);
    }
    m->max_terms_in_group = max_terms;
    m->all_one_body = one_body;
    m->quad_mixer = false;
    m->xJ.clear();
    m->xh.clear();
    if (!one_body) {  // every group made of one- and two-qubit X strings: quadratic forms in the x basis, one per group
        if (mixer_class == 2) {
            const int n = m->n;
            m->xJ.assign((size_t)n_groups * n * n, 0.0);
            m->xh.assign((size_t)n_groups * n, 0.0);
            for (int g = 0; g < n_groups; ++g)
                for (int k = group_offsets[g]; k < group_offsets[g + 1]; ++k) {
                    const int q1 = __builtin_ffsll((long long)masks[k]) - 1;
                    const uint64_t rest = masks[k] & (masks[k] - 1);
                    if (!rest) {
                        m->xh[(size_t)g * n + q1] += weights[k];
                    } else {
                        const int q2 = __builtin_ffsll((long long)rest) - 1;
                        m->xJ[((size_t)g * n + q1) * n + q2] += weights[k];
                        m->xJ[((size_t)g * n + q2) * n + q1] += weights[k];
                    }
                }
            m->quad_mixer = true;
        }
    }
    m->qweight.clear();
    cudaFree(m->d_qweight);
    m->d_qweight = nullptr;
    if (one_body) {  // per-qubit rotation weights of the sweep kernels: terms on the same qubit add up
        m->qweight.assign((size_t)n_groups * m->n, 0.0);
        for (int g = 0; g < n_groups; ++g)
            for (int k = group_offsets[g]; k < group_offsets[g + 1]; ++k)
                m->qweight[(size_t)g * m->n + (__builtin_ffsll((long long)masks[k]) - 1)] += weights[k];
        QMC_CUDA_TRY(cudaMalloc(&m->d_qweight, sizeof(double) * m->qweight.size()));
        QMC_CUDA_TRY(cudaMemcpy(m->d_qweight, m->qweight.data(), sizeof(double) * m->qweight.size(), cudaMemcpyHostToDevice));
    }
    // shared-memory path: E_x(x) of every group (host: 2^n x terms sign flips; n <= 13)
    cudaFree(m->dEx);
    m->dEx = nullptr;
    if (m->n <= QMC_SMEM_MAX_SPINS_F32 && !one_body) {
        const size_t N = (size_t)1 << m->n;
        std::vector<double> ex((size_t)n_groups * N, 0.0);
        for (int g = 0; g < n_groups; ++g)
            for (int k = group_offsets[g]; k < group_offsets[g + 1]; ++k) {
                const uint64_t mk = masks[k];
                const double wk = weights[k];
                double *e = &ex[(size_t)g * N];
                for (size_t x = 0; x < N; ++x) e[x] += (__builtin_popcountll((uint64_t)x & mk) & 1) ? -wk : wk;
            }
        QMC_CUDA_TRY(cudaMalloc(&m->dEx, sizeof(double) * ex.size()));
        QMC_CUDA_TRY(cudaMemcpy(m->dEx, ex.data(), sizeof(double) * ex.size(), cudaMemcpyHostToDevice));
    }
    m->mixer_epoch++;
    cudaFree(m->d_group_off);
    cudaFree(m->d_masks);
    cudaFree(m->d_weights);
    cudaFree(m->d_group_cum);
    m->d_group_off = nullptr; m->d_masks = nullptr; m->d_weights = nullptr; m->d_group_cum = nullptr;
    QMC_CUDA_TRY(cudaMalloc(&m->d_group_off, sizeof(int) * (n_groups + 1)));
    QMC_CUDA_TRY(cudaMalloc(&m->d_masks, sizeof(uint64_t) * nt));
    QMC_CUDA_TRY(cudaMalloc(&m->d_weights, sizeof(double) * nt));
    QMC_CUDA_TRY(cudaMalloc(&m->d_group_cum, sizeof(double) * n_groups));
    QMC_CUDA_TRY(cudaMemcpy(m->d_group_off, m->group_off.data(), sizeof(int) * (n_groups + 1), cudaMemcpyHostToDevice));
    QMC_CUDA_TRY(cudaMemcpy(m->d_masks, m->masks.data(), sizeof(uint64_t) * nt, cudaMemcpyHostToDevice));
    QMC_CUDA_TRY(cudaMemcpy(m->d_weights, m->weights.data(), sizeof(double) * nt, cudaMemcpyHostToDevice));
    QMC_CUDA_TRY(cudaMemcpy(m->d_group_cum, m->group_cum.data(), sizeof(double) * n_groups, cudaMemcpyHostToDevice));
    return QMC_OK;
}

extern "C" int qmc_energies(qmc_model *m, const uint64_t *idx_dev, int64_t count, double *E_dev, void *stream) {
    QMC_REQUIRE(m && (count == 0 || (idx_dev && E_dev)) && count >= 0, QMC_ERR_BADARG, "qmc_energies: bad argument");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    return qmc_launch_energies(m, idx_dev, count, E_dev, (cudaStream_t)stream);
}

extern "C" int qmc_energies_host(qmc_model *m, const uint64_t *idx_host, int64_t count, double *E_host) {
    QMC_REQUIRE(m && (count == 0 || (idx_host && E_host)) && count >= 0, QMC_ERR_BADARG, "qmc_energies_host: bad argument");
    if (count == 0) return QMC_OK;
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    uint64_t *di = nullptr;
    double *de = nullptr;
    QMC_CUDA_TRY(cudaMalloc(&di, sizeof(uint64_t) * count));
    cudaError_t e = cudaMalloc(&de, sizeof(double) * count);
    int rc = QMC_OK;
    if (e != cudaSuccess) { cudaFree(di); qmc_set_error("cudaMalloc: %s", cudaGetErrorString(e)); return QMC_ERR_CUDA; }
    e = cudaMemcpy(di, idx_host, sizeof(uint64_t) * count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = qmc_launch_energies(m, di, count, de, 0);
    if (e == cudaSuccess && rc == QMC_OK) e = cudaMemcpy(E_host, de, sizeof(double) * count, cudaMemcpyDeviceToHost);
    cudaFree(di);
    cudaFree(de);
    if (e != cudaSuccess) { qmc_set_error("qmc_energies_host: %s", cudaGetErrorString(e)); return QMC_ERR_CUDA; }
    return rc;
}

extern "C" int qmc_exact_sampling(qmc_model *m, double beta, double *logZ_dev, double *energies_dev,
                                  double *probs_dev, void *stream) {
    QMC_NVTX("qmc_exact_sampling");
    QMC_REQUIRE(m && logZ_dev, QMC_ERR_BADARG, "qmc_exact_sampling: null argument");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = m->n;
    const int64_t N = 1LL << n;
    const size_t smem = sizeof(double) * (size_t)(n * n + n);
    const int grid = grid_for(N, 256);
    LseState *partials = nullptr;
    QMC_CUDA_TRY(cudaMallocAsync(&partials, sizeof(LseState) * grid, st));
    const bool fast = !energies_dev && fast_range_ok(m, 0, N);
    if (fast) {
        int rc = ensure_fast_tables(m);
        if (rc != QMC_OK) return rc;
        const int64_t nb = N / FAST_BLOCK;
        const int fgrid = (int)std::min<int64_t>(nb, grid);
        exact_fast_kernel<false><<<fgrid, 256, fast_smem(n), st>>>(n, m->dJp, m->dhq_e, m->fast_e0, beta, 0, nb, partials, nullptr, 0.0, nullptr);
        exact_finalize_kernel<<<1, 256, 0, st>>>(partials, fgrid, logZ_dev, nullptr);
        m->launches += 2;
        if (probs_dev) {
            exact_fast_kernel<true><<<fgrid, 256, fast_smem(n), st>>>(n, m->dJp, m->dhq_e, m->fast_e0, beta, 0, nb, nullptr, logZ_dev, 0.0, probs_dev);
            m->launches++;
        }
    } else {
        exact_enumerate_kernel<<<grid, 256, smem, st>>>(n, m->dJ, m->dh, beta, 0, N, energies_dev, partials);
        exact_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, logZ_dev, nullptr);
        m->launches += 2;
        if (probs_dev) {
            exact_probs_kernel<<<grid, 256, smem, st>>>(n, m->dJ, m->dh, beta, 0, N, energies_dev, logZ_dev, 0.0, probs_dev);
            m->launches++;
        }
    }
    QMC_CUDA_TRY(cudaGetLastError());
    QMC_CUDA_TRY(cudaFreeAsync(partials, st));
    return QMC_OK;
}

extern "C" int qmc_exact_partial(qmc_model *m, double beta, uint64_t idx_begin, int64_t idx_count, double *lse_pair_dev,
                                 double *energies_dev, void *stream) {
    QMC_REQUIRE(m && lse_pair_dev && idx_count >= 0, QMC_ERR_BADARG, "qmc_exact_partial: bad argument");
    const uint64_t N = 1ULL << m->n;
    QMC_REQUIRE(idx_begin <= N && (uint64_t)idx_count <= N - idx_begin, QMC_ERR_BADARG, "qmc_exact_partial: range outside [0, 2^n)");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = m->n;
    const size_t smem = sizeof(double) * (size_t)(n * n + n);
    const int grid = grid_for(idx_count, 256);
    LseState *partials = nullptr;
    QMC_CUDA_TRY(cudaMallocAsync(&partials, sizeof(LseState) * grid, st));
    if (!energies_dev && fast_range_ok(m, idx_begin, idx_count)) {
        int rc = ensure_fast_tables(m);
        if (rc != QMC_OK) return rc;
        const int64_t nb = idx_count / FAST_BLOCK;
        const int fgrid = (int)std::min<int64_t>(nb, grid);
        exact_fast_kernel<false><<<fgrid, 256, fast_smem(n), st>>>(n, m->dJp, m->dhq_e, m->fast_e0, beta, (int64_t)idx_begin, nb, partials, nullptr, 0.0, nullptr);
        exact_finalize_kernel<<<1, 256, 0, st>>>(partials, fgrid, nullptr, lse_pair_dev);
    } else {
        exact_enumerate_kernel<<<grid, 256, smem, st>>>(n, m->dJ, m->dh, beta, (int64_t)idx_begin, idx_count, energies_dev, partials);
        exact_finalize_kernel<<<1, 256, 0, st>>>(partials, grid, nullptr, lse_pair_dev);
    }
    m->launches += 2;
    QMC_CUDA_TRY(cudaGetLastError());
    QMC_CUDA_TRY(cudaFreeAsync(partials, st));
    return QMC_OK;
}

extern "C" int qmc_exact_probs_range(qmc_model *m, double beta, double logZ, uint64_t idx_begin, int64_t idx_count,
                                     const double *energies_dev, double *probs_dev, void *stream) {
    QMC_REQUIRE(m && probs_dev && idx_count >= 0, QMC_ERR_BADARG, "qmc_exact_probs_range: bad argument");
    const uint64_t N = 1ULL << m->n;
    QMC_REQUIRE(idx_begin <= N && (uint64_t)idx_count <= N - idx_begin, QMC_ERR_BADARG, "qmc_exact_probs_range: range outside [0, 2^n)");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    const int n = m->n;
    const size_t smem = sizeof(double) * (size_t)(n * n + n);
    if (!energies_dev && fast_range_ok(m, idx_begin, idx_count)) {
        int rc = ensure_fast_tables(m);
        if (rc != QMC_OK) return rc;
        const int64_t nb = idx_count / FAST_BLOCK;
        const int fgrid = (int)std::min<int64_t>(nb, grid_for(idx_count, 256));
        exact_fast_kernel<true><<<fgrid, 256, fast_smem(n), (cudaStream_t)stream>>>(n, m->dJp, m->dhq_e, m->fast_e0, beta, (int64_t)idx_begin, nb,
                                                                                    nullptr, nullptr, logZ, probs_dev);
    } else {
        exact_probs_kernel<<<grid_for(idx_count, 256), 256, smem, (cudaStream_t)stream>>>(n, m->dJ, m->dh, beta, (int64_t)idx_begin, idx_count,
                                                                                         energies_dev, nullptr, logZ, probs_dev);
    }
    m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

extern "C" int qmc_exact_sampling_host(qmc_model *m, double beta, double *logZ_host, double *energies_host,
                                       double *probs_host) {
    QMC_REQUIRE(m && logZ_host, QMC_ERR_BADARG, "qmc_exact_sampling_host: null argument");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    const int64_t N = 1LL << m->n;
    double *dl = nullptr, *de = nullptr, *dp = nullptr;
    int rc = QMC_OK;
    cudaError_t e = cudaMalloc(&dl, sizeof(double));
    if (e == cudaSuccess && energies_host) e = cudaMalloc(&de, sizeof(double) * N);
    if (e == cudaSuccess && probs_host) e = cudaMalloc(&dp, sizeof(double) * N);
    if (e == cudaSuccess) rc = qmc_exact_sampling(m, beta, dl, de, dp, nullptr);
    if (e == cudaSuccess && rc == QMC_OK) e = cudaMemcpy(logZ_host, dl, sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == QMC_OK && de) e = cudaMemcpy(energies_host, de, sizeof(double) * N, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == QMC_OK && dp) e = cudaMemcpy(probs_host, dp, sizeof(double) * N, cudaMemcpyDeviceToHost);
    cudaFree(dl);
    cudaFree(de);
    cudaFree(dp);
    if (e != cudaSuccess) { qmc_set_error("qmc_exact_sampling_host: %s", cudaGetErrorString(e)); return QMC_ERR_CUDA; }
    return rc;
}

extern "C" int qmc_model_num_spins(const qmc_model *m) { return m ? m->n : -1; }
extern "C" int64_t qmc_model_launch_count(const qmc_model *m) { return m ? m->launches : -1; }
extern "C" int qmc_profile_enable(qmc_model *m, int on) {
    QMC_REQUIRE(m, QMC_ERR_BADARG, "qmc_profile_enable: null handle");
    m->profile = on;
    m->prof_ms = 0.0;
    m->prof_launches = 0;
    m->prof_bytes = 0.0;
    m->prof_moved = 0.0;
    return QMC_OK;
}
extern "C" int qmc_profile_read(qmc_model *m, double *sweep_ms, int64_t *sweep_launches, double *algorithmic_bytes) {
    QMC_REQUIRE(m, QMC_ERR_BADARG, "qmc_profile_read: null handle");
    if (sweep_ms) *sweep_ms = m->prof_ms;
    if (sweep_launches) *sweep_launches = m->prof_launches;
    if (algorithmic_bytes) *algorithmic_bytes = m->prof_bytes;
    return QMC_OK;
}
extern "C" int qmc_profile_read_moved(qmc_model *m, double *moved_bytes) {
    QMC_REQUIRE(m && moved_bytes, QMC_ERR_BADARG, "qmc_profile_read_moved: null argument");
    *moved_bytes = m->prof_moved;
    return QMC_OK;
}
