// general.cu -- HBM statevector path for everything the specialised kernels do not cover: arbitrary X-string mixers
// (GenericMixer k-body, CustomMixer, CoherentMixerSum, IncoherentMixerSum: qumcmc/mixers.py:21-151) above the
// shared-memory limit, and complex128 statevectors above n = 12.  13 <= n <= QMC_GENERAL_MAX_SPINS.
//
// The reference pays one pass over the statevector per gate (qumcmc/quantum_mcmc_routines.py:54-106).  Here the mixer
// terms of a group are partitioned into TILE passes: a tile is a set of 12 qubit positions (always including qubits
// 0..3, so that HBM accesses come in runs of 16 amplitudes); a CTA gathers the 2^12 amplitudes of one tile instance
// into shared memory, applies every term whose support lies inside the tile, and scatters them back.  All X-string
// rotations commute, so a layer M = prod_tiles M_tile in any order; consecutive layers walk the tiles in opposite
// directions and the turn-around tile applies  M_tile(k) P M_tile(k+1)  in one pass, P being the diagonal phase
// layer exp(+i c D(idx)) read from a 2^n fp64 table.  m tiles => r (m - 1) + 1 passes for r mixer layers (one pass in
// total when a single tile holds every term).  The first pass starts from the basis ket analytically (no read).
// Measurement: 64-amplitude segment sums (fp64) -> block inverse CDF -> the 64 amplitudes of the chosen segment.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int GT = 12;          // tile qubits
constexpr int GN = 1 << GT;     // amplitudes per tile
constexpr int GTH = 256;        // threads per CTA
constexpr int GFORCED = 4;      // qubits 0..3 belong to every tile
constexpr int GSEL = 128;       // threads of the sampler

struct GenPass {
    unsigned char tpos[GT];  // ascending qubit positions of the tile
    int term0, term1;        // range in the assigned-term arrays
};

struct GenChain {  // per chain, rewritten every hop
    unsigned long long cur;
    double gamma, e_cur, u_meas, u_acc;
    long long n_acc;
    int r, grp;
};

struct GenArgs {
    int n, n_groups, phase_active, mode, max_terms;
    double alpha, dt, temperature, g_lo, g_hi;
    const GenPass *passes;
    const int *group_pass_off;       // [n_groups + 1]
    const unsigned short *lmask;     // assigned terms: mask in tile-local bit coordinates
    const double *weight;            // assigned terms: gamma multiplier
    const double *D;                 // [2^n] phase exponent
    const double *J, *h, *group_cum;
    GenChain *chains;
    void *state;                     // [batch][2^n] complex
    double *segsum;                  // [batch][2^n / 64]
    // hop bookkeeping
    uint64_t seed, chain_id0;
    uint32_t hop;
    int64_t n_hops, hop_index, chain0;
    const uint64_t *s_in;
    const double *gamma_arr, *u_arr;
    const int *r_arr, *group_arr;
    uint64_t *proposals;
    uint8_t *accepted;
    uint64_t *final_state;
    int64_t *acc_count, *hist;
    void *probs_out, *amps_out;
    unsigned long long *visit;  // visit histogram accumulator or null
};

template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };

__device__ __forceinline__ float2 gen_phase(double x, float) {
    double t = x * 0.31830988618379067154;  // x / pi
    t -= 2.0 * rint(0.5 * t);
    float s, c;
    sincospif((float)t, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ double2 gen_phase(double x, double) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(c, s);
}

// Phase exponent D(idx) = sum_{j<k} J_jk z_j z_k + sum_j h_j z_j, z = -s (fn_qckt_problem_half, :54-91); same loop
// order as the oracle's orc_phase_D.
__global__ void __launch_bounds__(256) gen_phase_table_kernel(int n, const double *__restrict__ J, const double *__restrict__ h,
                                                              double *__restrict__ D) {
    extern __shared__ double sm[];
    double *sJ = sm, *sh = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sJ[i] = J[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = h[i];
    __syncthreads();
    const int64_t N = 1LL << n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < N; idx += (int64_t)gridDim.x * blockDim.x) {
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

// Per-hop parameters of every chain of the batch (draws: common.cuh; explicit arrays in PROPOSE / PROBS mode).
__global__ void gen_prepare_kernel(const GenArgs a, int64_t nc, int first_hop) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nc) return;
    GenChain &ch = a.chains[c];
    const int64_t gc = a.chain0 + c;                    // chain index inside this call
    const uint64_t gchain = a.chain_id0 + (uint64_t)gc;  // global chain id (Philox key)
    if (a.mode == QMC_MODE_CHAIN) {
        if (first_hop) {
            ch.cur = a.s_in ? a.s_in[gc] : qmc_initial_state(a.seed, gchain, a.n);
            ch.e_cur = qmc_energy_dev(a.n, a.J, a.h, ch.cur);
            ch.n_acc = 0;
            if (a.hop == 0) qmc_visit(a.visit, ch.cur);  // entry 0 of markov_chain
        }
        const HopDraws d = qmc_hop_draws(a.seed, gchain, a.hop);
        ch.gamma = a.g_lo + (a.g_hi - a.g_lo) * d.u_gamma;
        ch.r = qmc_trotter_steps(d.t, a.dt);
        ch.grp = qmc_pick_group(a.n_groups, a.group_cum, d.u_group);
        ch.u_meas = d.u_meas;
        ch.u_acc = d.u_acc;
    } else {
        ch.cur = a.s_in[gc];
        ch.e_cur = 0.0;
        ch.n_acc = 0;
        ch.gamma = a.gamma_arr[gc];
        ch.r = a.r_arr[gc] < 1 ? 1 : a.r_arr[gc];
        ch.grp = a.group_arr ? a.group_arr[gc] : 0;
        ch.u_meas = a.u_arr ? a.u_arr[gc] : 0.0;
        ch.u_acc = 0.0;
    }
}

// One pass (step t of the proposal) for every (tile instance, chain).  Chains whose proposal has fewer steps exit.
template <typename T>
__global__ void __launch_bounds__(GTH) gen_pass_kernel(const GenArgs a, const int t) {
    using T2 = typename V2<T>::type;
    extern __shared__ __align__(16) unsigned char smraw[];
    T2 *st = reinterpret_cast<T2 *>(smraw);           // [GN]
    T2 *cs = st + GN;                                 // [terms of this pass]
    int *tm = reinterpret_cast<int *>(cs + a.max_terms);  // their tile-local masks
    __shared__ unsigned long long hoff[GN >> GFORCED];  // index offset of the local bits 4..11
    __shared__ GenChain ch_s;
    const int tid = threadIdx.x;
    const int n = a.n;
    const int64_t c = blockIdx.y;
    if (tid == 0) ch_s = a.chains[c];
    __syncthreads();
    const int r = ch_s.r, g = ch_s.grp;
    const int p0 = a.group_pass_off[g], m = a.group_pass_off[g + 1] - p0;
    const int T_steps = m == 1 ? 1 : r * (m - 1) + 1;
    if (t >= T_steps) return;
    int tile, reps;
    if (m == 1) {
        tile = 0;
        reps = r;
    } else {
        const int L = t / (m - 1), o = t - L * (m - 1);
        tile = (L & 1) ? m - 1 - o : o;
        reps = (o == 0 && L >= 1 && L < r) ? 2 : 1;
    }
    const GenPass &P = a.passes[p0 + tile];
    // tile instance: the n - 12 bits of blockIdx.x go to the positions outside the tile
    unsigned long long tmask = 0ULL;
#pragma unroll
    for (int k = 0; k < GT; ++k) tmask |= 1ULL << P.tpos[k];
    unsigned long long base = 0ULL;
    {
        unsigned long long b = blockIdx.x;
        for (int q = 0; q < n; ++q)
            if (!((tmask >> q) & 1ULL)) {
                base |= (b & 1ULL) << q;
                b >>= 1;
            }
    }
    {   // local bits 4..11 -> index offset (tid = the 8-bit pattern)
        unsigned long long o = 0ULL;
#pragma unroll
        for (int k = GFORCED; k < GT; ++k)
            if ((tid >> (k - GFORCED)) & 1) o |= 1ULL << P.tpos[k];
        hoff[tid] = o;
    }
    const int nterms = P.term1 - P.term0;
    for (int k = tid; k < nterms; k += GTH) {  // exp(-i gamma w dt X_S) -> (cos, sin)
        double s, cc;
        sincos(ch_s.gamma * a.weight[P.term0 + k] * a.dt, &s, &cc);
        T2 v;
        v.x = (T)cc;
        v.y = (T)s;
        cs[k] = v;
        tm[k] = a.lmask[P.term0 + k];
    }
    __syncthreads();
    T2 *state = reinterpret_cast<T2 *>(a.state) + ((size_t)c << n);
    if (t == 0) {  // |s>
        const unsigned long long cur = ch_s.cur;
#pragma unroll 4
        for (int j = tid; j < GN; j += GTH) {
            const unsigned long long gi = base | hoff[j >> GFORCED] | (unsigned long long)(j & 15);
            T2 z;
            z.x = gi == cur ? T(1) : T(0);
            z.y = T(0);
            st[j] = z;
        }
    } else {
#pragma unroll 4
        for (int j = tid; j < GN; j += GTH) st[j] = state[base | hoff[j >> GFORCED] | (unsigned long long)(j & 15)];
    }
    __syncthreads();
    const double cphase = (1.0 - ch_s.gamma) * a.alpha * a.dt;
    for (int rep = 0; rep < reps; ++rep) {
        if (rep > 0 && a.phase_active) {
#pragma unroll 2
            for (int j = tid; j < GN; j += GTH) {
                const unsigned long long gi = base | hoff[j >> GFORCED] | (unsigned long long)(j & 15);
                const T2 f = gen_phase(cphase * a.D[gi], T(0));
                const T2 x = st[j];
                T2 o;
                o.x = x.x * f.x - x.y * f.y;
                o.y = x.x * f.y + x.y * f.x;
                st[j] = o;
            }
            __syncthreads();
        }
        // Terms two at a time on orbits {i, i^S1, i^S2, i^S1^S2} held in registers (as in small_n.cu): half the
        // shared-memory traffic and half the barriers, same arithmetic per amplitude in the same order.
        int k = 0;
        while (k < nterms) {
            const int m1 = tm[k];
            const int m2 = (k + 1 < nterms) ? tm[k + 1] : m1;
            if (m2 != m1) {
                const int p1 = 31 - __clz(m1);
                const int m2r = ((m2 >> p1) & 1) ? (m2 ^ m1) : m2;
                const int p2 = 31 - __clz(m2r);
                const int lo_b = p1 < p2 ? p1 : p2, hi_b = p1 < p2 ? p2 : p1;
                const int lo_m = (1 << lo_b) - 1, hi_m = (1 << hi_b) - 1;
                const T c1 = cs[k].x, s1 = cs[k].y, c2 = cs[k + 1].x, s2 = cs[k + 1].y;
#pragma unroll 2
                for (int q = tid; q < (GN >> 2); q += GTH) {
                    int i = ((q >> lo_b) << (lo_b + 1)) | (q & lo_m);  // orbit representative: bits p1, p2 clear
                    i = ((i >> hi_b) << (hi_b + 1)) | (i & hi_m);
                    const int i1 = i ^ m1, i2 = i ^ m2, i3 = i1 ^ m2;
                    T2 a0 = st[i], a1 = st[i1], a2 = st[i2], a3 = st[i3];
                    T2 u0, u1;
                    u0.x = c1 * a0.x + s1 * a1.y; u0.y = c1 * a0.y - s1 * a1.x;
                    u1.x = c1 * a1.x + s1 * a0.y; u1.y = c1 * a1.y - s1 * a0.x;
                    a0 = u0; a1 = u1;
                    u0.x = c1 * a2.x + s1 * a3.y; u0.y = c1 * a2.y - s1 * a3.x;
                    u1.x = c1 * a3.x + s1 * a2.y; u1.y = c1 * a3.y - s1 * a2.x;
                    a2 = u0; a3 = u1;
                    u0.x = c2 * a0.x + s2 * a2.y; u0.y = c2 * a0.y - s2 * a2.x;
                    u1.x = c2 * a2.x + s2 * a0.y; u1.y = c2 * a2.y - s2 * a0.x;
                    a0 = u0; a2 = u1;
                    u0.x = c2 * a1.x + s2 * a3.y; u0.y = c2 * a1.y - s2 * a3.x;
                    u1.x = c2 * a3.x + s2 * a1.y; u1.y = c2 * a3.y - s2 * a1.x;
                    a1 = u0; a3 = u1;
                    st[i] = a0; st[i1] = a1; st[i2] = a2; st[i3] = a3;
                }
                __syncthreads();
                k += 2;
                continue;
            }
            const int tb = 31 - __clz(m1);
            const int lowmask = (1 << tb) - 1;
            const T cc = cs[k].x, s = cs[k].y;
#pragma unroll 4
            for (int q = tid; q < (GN >> 1); q += GTH) {
                const int i = ((q >> tb) << (tb + 1)) | (q & lowmask);
                const int j = i ^ m1;
                const T2 x = st[i], y = st[j];
                T2 ox, oy;
                ox.x = cc * x.x + s * y.y;
                ox.y = cc * x.y - s * y.x;
                oy.x = cc * y.x + s * x.y;
                oy.y = cc * y.y - s * x.x;
                st[i] = ox;
                st[j] = oy;
            }
            __syncthreads();
            k += 1;
        }
    }
#pragma unroll 4
    for (int j = tid; j < GN; j += GTH) state[base | hoff[j >> GFORCED] | (unsigned long long)(j & 15)] = st[j];
}

// |a|^2 over 64-amplitude segments, fp64
template <typename T>
__global__ void __launch_bounds__(256) gen_segsum_kernel(const GenArgs a) {
    using T2 = typename V2<T>::type;
    const int n = a.n;
    const int64_t c = blockIdx.y;
    const int64_t nseg = 1LL << (n - 6);
    const T2 *state = reinterpret_cast<const T2 *>(a.state) + ((size_t)c << n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t seg = blockIdx.x * 8 + warp; seg < nseg; seg += (int64_t)gridDim.x * 8) {
        const T2 x = state[seg * 64 + lane], y = state[seg * 64 + 32 + lane];
        double s = (double)x.x * (double)x.x + (double)x.y * (double)x.y + (double)y.x * (double)y.x + (double)y.y * (double)y.y;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) a.segsum[c * nseg + seg] = s;
    }
}

// K3 + K4 for one chain per CTA: inverse CDF over the segment sums, index inside the segment, energy, Metropolis.
template <typename T>
__global__ void __launch_bounds__(GSEL) gen_sample_kernel(const GenArgs a) {
    using T2 = typename V2<T>::type;
    __shared__ double s_incl[GSEL];
    __shared__ double s_warp[GSEL / 32];
    __shared__ long long s_seg;
    __shared__ double s_rem;
    const int tid = threadIdx.x;
    const int n = a.n;
    const int64_t c = blockIdx.x;
    GenChain &ch = a.chains[c];
    const int64_t nseg = 1LL << (n - 6);
    const double *seg = a.segsum + c * nseg;
    const int64_t per = (nseg + GSEL - 1) / GSEL;
    const int64_t i0 = tid * per < nseg ? tid * per : nseg, i1 = i0 + per < nseg ? i0 + per : nseg;
    const int last_owner = (int)((nseg - 1) / per);
    double local = 0.0;
    for (int64_t i = i0; i < i1; ++i) local += seg[i];
    double incl = local;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double x = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    double basev = 0.0;
    for (int w = 0; w < warp; ++w) basev += s_warp[w];
    incl += basev;
    s_incl[tid] = incl;
    __syncthreads();
    const double total = s_incl[GSEL - 1];
    const double target = ch.u_meas * total;
    const double lo = tid == 0 ? 0.0 : s_incl[tid - 1];
    if (tid <= last_owner && target >= lo && (target < incl || tid == last_owner)) {
        double cum = lo;
        int64_t sel = -1;
        double before = 0.0;
        for (int64_t i = i0; i < i1; ++i) {
            const double x = seg[i];
            if (target < cum + x) {
                sel = i;
                before = cum;
                break;
            }
            cum += x;
        }
        if (sel < 0) {
            sel = i1 - 1;
            before = cum - seg[sel];
        }
        s_seg = sel;
        s_rem = target - before;
    }
    __syncthreads();
    if (tid != 0) return;
    const T2 *state = reinterpret_cast<const T2 *>(a.state) + ((size_t)c << n);
    const int64_t sg = s_seg;
    const double rem = s_rem;
    double cum = 0.0;
    int selj = 63;
    for (int j = 0; j < 64; ++j) {
        const T2 x = state[sg * 64 + j];
        cum += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
        if (rem < cum) {
            selj = j;
            break;
        }
    }
    const uint64_t sp = (uint64_t)sg * 64ULL + (uint64_t)selj;
    const int64_t gc = a.chain0 + c;
    if (a.mode == QMC_MODE_PROPOSE) {
        a.proposals[gc] = sp;
        return;
    }
    const double e_sp = qmc_energy_dev(n, a.J, a.h, sp);
    const bool acc = qmc_accept_dev(ch.e_cur, e_sp, a.temperature, ch.u_acc);
    if (a.proposals) a.proposals[gc * a.n_hops + a.hop_index] = sp;
    if (a.accepted) a.accepted[gc * a.n_hops + a.hop_index] = (uint8_t)acc;
    if (a.hist) a.hist[gc * (n + 1) + __popcll(ch.cur ^ sp)] += 1;
    if (acc) {
        ch.cur = sp;
        ch.e_cur = e_sp;
        ch.n_acc += 1;
    }
    qmc_visit(a.visit, ch.cur);
}

__global__ void gen_finish_kernel(const GenArgs a, int64_t nc) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int64_t gc = a.chain0 + c;
    if (a.final_state) a.final_state[gc] = a.chains[c].cur;
    if (a.acc_count) a.acc_count[gc] = a.chains[c].n_acc;
}

// PROBS mode (one chain): normalised probabilities / amplitudes
template <typename T>
__global__ void __launch_bounds__(256) gen_probs_kernel(const GenArgs a) {
    using T2 = typename V2<T>::type;
    __shared__ double red[256];
    const int n = a.n;
    const int64_t N = 1LL << n, nseg = N >> 6;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nseg; i += blockDim.x) s += a.segsum[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    const double total = red[0];
    const T inv = (T)(1.0 / total), inva = (T)(1.0 / sqrt(total));
    const T2 *state = reinterpret_cast<const T2 *>(a.state);
    T *po = reinterpret_cast<T *>(a.probs_out);
    T2 *ao = reinterpret_cast<T2 *>(a.amps_out);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const T2 x = state[i];
        if (po) po[i] = (x.x * x.x + x.y * x.y) * inv;
        if (ao) {
            T2 o;
            o.x = x.x * inva;
            o.y = x.y * inva;
            ao[i] = o;
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct GeneralWorkspace {
    int mixer_epoch = -1;
    std::vector<GenPass> passes;
    std::vector<int> group_pass_off;
    std::vector<unsigned short> lmask;
    std::vector<double> weight;
    int max_terms_per_pass = 0, max_passes = 0;
    GenPass *d_passes = nullptr;
    int *d_group_pass_off = nullptr;
    unsigned short *d_lmask = nullptr;
    double *d_weight = nullptr;
    double *dD = nullptr;
    // per call
    int64_t cap_chains = 0;
    size_t cap_esz = 0;
    void *state = nullptr;
    double *segsum = nullptr;
    GenChain *chains = nullptr;
};

void qmc_general_free(qmc_model *m) {
    GeneralWorkspace *w = m->general;
    if (!w) return;
    cudaFree(w->d_passes); cudaFree(w->d_group_pass_off); cudaFree(w->d_lmask); cudaFree(w->d_weight);
    cudaFree(w->dD); cudaFree(w->state); cudaFree(w->segsum); cudaFree(w->chains);
    delete w;
    m->general = nullptr;
}

// Greedy partition of one group's terms into tiles of GT qubits containing qubits 0..GFORCED-1.
static int plan_group(int n, const uint64_t *masks, const double *weights, int nterms, std::vector<GenPass> &passes,
                      std::vector<unsigned short> &lmask, std::vector<double> &weight) {
    const uint64_t forced = (1ULL << GFORCED) - 1ULL;
    std::vector<char> done(nterms, 0);
    int left = nterms;
    while (left > 0) {
        uint64_t tile = forced;
        for (;;) {  // grow: the undone term that adds the fewest new qubits (ties: lowest new qubit)
            int best = -1, best_cost = 99;
            uint64_t best_new = ~0ULL;
            for (int k = 0; k < nterms; ++k) {
                if (done[k]) continue;
                const uint64_t add = masks[k] & ~tile;
                if (!add) continue;  // already inside
                const int cost = __builtin_popcountll(add);
                if (__builtin_popcountll(tile) + cost > GT) continue;
                const uint64_t low = add & (~add + 1ULL);
                if (cost < best_cost || (cost == best_cost && low < best_new)) {
                    best = k;
                    best_cost = cost;
                    best_new = low;
                }
            }
            if (best < 0) break;
            tile |= masks[best];
        }
        for (int q = 0; q < n && __builtin_popcountll(tile) < GT; ++q) tile |= 1ULL << q;  // fill with low qubits
        GenPass P;
        int k2 = 0;
        for (int q = 0; q < n; ++q)
            if ((tile >> q) & 1ULL) P.tpos[k2++] = (unsigned char)q;
        P.term0 = (int)lmask.size();
        int took = 0;
        for (int k = 0; k < nterms; ++k) {
            if (done[k] || (masks[k] & ~tile)) continue;
            unsigned short lm = 0;
            for (int b = 0; b < GT; ++b)
                if ((masks[k] >> P.tpos[b]) & 1ULL) lm |= (unsigned short)(1u << b);
            lmask.push_back(lm);
            weight.push_back(weights[k]);
            done[k] = 1;
            ++took;
        }
        P.term1 = (int)lmask.size();
        if (!took) {
            qmc_set_error("mixer term with more than %d qubits outside qubits 0..%d cannot be tiled", GT - GFORCED, GFORCED - 1);
            return QMC_ERR_UNSUPPORTED;
        }
        left -= took;
        passes.push_back(P);
    }
    return QMC_OK;
}

static int general_prepare(qmc_model *m, cudaStream_t st) {
    if (!m->general) m->general = new GeneralWorkspace();
    GeneralWorkspace *w = m->general;
    const int n = m->n;
    if (!w->dD) {
        double *dcJ = nullptr, *dch = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&dcJ, sizeof(double) * n * n));
        QMC_CUDA_TRY(cudaMalloc(&dch, sizeof(double) * n));
        QMC_CUDA_TRY(cudaMemcpyAsync(dcJ, m->host_cJ.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice, st));
        QMC_CUDA_TRY(cudaMemcpyAsync(dch, m->host_ch.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
        QMC_CUDA_TRY(cudaMalloc(&w->dD, sizeof(double) << n));
        gen_phase_table_kernel<<<148 * 8, 256, sizeof(double) * (n * n + n), st>>>(n, dcJ, dch, w->dD);
        m->launches++;
        QMC_CUDA_TRY(cudaGetLastError());
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(dcJ);
        cudaFree(dch);
    }
    if (w->mixer_epoch != m->mixer_epoch) {
        w->passes.clear(); w->group_pass_off.assign(1, 0); w->lmask.clear(); w->weight.clear();
        w->max_terms_per_pass = 0; w->max_passes = 0;
        for (int g = 0; g < m->n_groups; ++g) {
            const int t0 = m->group_off[g], t1 = m->group_off[g + 1];
            int rc = plan_group(n, m->masks.data() + t0, m->weights.data() + t0, t1 - t0, w->passes, w->lmask, w->weight);
            if (rc != QMC_OK) return rc;
            w->group_pass_off.push_back((int)w->passes.size());
            w->max_passes = std::max(w->max_passes, w->group_pass_off[g + 1] - w->group_pass_off[g]);
        }
        for (auto &P : w->passes) w->max_terms_per_pass = std::max(w->max_terms_per_pass, P.term1 - P.term0);
        cudaFree(w->d_passes); cudaFree(w->d_group_pass_off); cudaFree(w->d_lmask); cudaFree(w->d_weight);
        w->d_passes = nullptr; w->d_group_pass_off = nullptr; w->d_lmask = nullptr; w->d_weight = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&w->d_passes, sizeof(GenPass) * w->passes.size()));
        QMC_CUDA_TRY(cudaMalloc(&w->d_group_pass_off, sizeof(int) * w->group_pass_off.size()));
        QMC_CUDA_TRY(cudaMalloc(&w->d_lmask, sizeof(unsigned short) * w->lmask.size()));
        QMC_CUDA_TRY(cudaMalloc(&w->d_weight, sizeof(double) * w->weight.size()));
        QMC_CUDA_TRY(cudaMemcpy(w->d_passes, w->passes.data(), sizeof(GenPass) * w->passes.size(), cudaMemcpyHostToDevice));
        QMC_CUDA_TRY(cudaMemcpy(w->d_group_pass_off, w->group_pass_off.data(), sizeof(int) * w->group_pass_off.size(), cudaMemcpyHostToDevice));
        QMC_CUDA_TRY(cudaMemcpy(w->d_lmask, w->lmask.data(), sizeof(unsigned short) * w->lmask.size(), cudaMemcpyHostToDevice));
        QMC_CUDA_TRY(cudaMemcpy(w->d_weight, w->weight.data(), sizeof(double) * w->weight.size(), cudaMemcpyHostToDevice));
        w->mixer_epoch = m->mixer_epoch;
    }
    return QMC_OK;
}

int qmc_general_num_passes(qmc_model *m, int group) {  // test / benchmark hook: tiles of a mixer group
    if (!m || !m->general || group < 0 || group + 1 >= (int)m->general->group_pass_off.size()) return -1;
    return m->general->group_pass_off[group + 1] - m->general->group_pass_off[group];
}

template <typename T>
static int general_run_t(qmc_model *m, GeneralWorkspace *w, GenArgs a, int64_t n_chains, int64_t n_hops, uint32_t hop0,
                         int r_cap, cudaStream_t st) {
    using T2 = typename V2<T>::type;
    const int n = m->n;
    const int64_t N = 1LL << n;
    const size_t smem = sizeof(T2) * (size_t)(GN + w->max_terms_per_pass) + sizeof(int) * (size_t)w->max_terms_per_pass;
    QMC_REQUIRE(smem <= 200 * 1024, QMC_ERR_UNSUPPORTED, "general path: %d mixer terms in one tile", w->max_terms_per_pass);
    QMC_CUDA_TRY(cudaFuncSetAttribute(gen_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    size_t free_b = 0, total_b = 0;
    QMC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t per_chain = sizeof(T2) * (size_t)N + sizeof(double) * (size_t)(N >> 6) + sizeof(GenChain);
    const size_t have = (size_t)w->cap_chains * (w->cap_esz * (size_t)N + sizeof(double) * (size_t)(N >> 6) + sizeof(GenChain));
    int64_t cap = (int64_t)(((double)(free_b + have) * 0.85) / (double)per_chain);
    if (cap > 65535) cap = 65535;
    QMC_REQUIRE(cap >= 1, QMC_ERR_CUDA, "not enough device memory for one 2^%d statevector", n);
    const int64_t batch = std::min<int64_t>(cap, n_chains);
    if (batch > w->cap_chains || sizeof(T2) != w->cap_esz) {
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(w->state); cudaFree(w->segsum); cudaFree(w->chains);
        w->state = nullptr; w->segsum = nullptr; w->chains = nullptr; w->cap_chains = 0;
        QMC_CUDA_TRY(cudaMalloc(&w->state, sizeof(T2) * (size_t)N * batch));
        QMC_CUDA_TRY(cudaMalloc(&w->segsum, sizeof(double) * (size_t)(N >> 6) * batch));
        QMC_CUDA_TRY(cudaMalloc(&w->chains, sizeof(GenChain) * batch));
        w->cap_chains = batch;
        w->cap_esz = sizeof(T2);
    }
    a.state = w->state;
    a.segsum = w->segsum;
    a.chains = w->chains;
    const int max_steps = w->max_passes == 1 ? 1 : r_cap * (w->max_passes - 1) + 1;
    const int64_t tiles = N >> GT;
    const int64_t hops = a.mode == QMC_MODE_CHAIN ? n_hops : 1;
    for (int64_t c0 = 0; c0 < n_chains; c0 += batch) {
        const int64_t nc = std::min<int64_t>(batch, n_chains - c0);
        a.chain0 = c0;
        for (int64_t hop = 0; hop < hops; ++hop) {
            a.hop = hop0 + (uint32_t)hop;
            a.hop_index = hop;
            gen_prepare_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(a, nc, hop == 0);
            for (int t = 0; t < max_steps; ++t)
                gen_pass_kernel<T><<<dim3((unsigned)tiles, (unsigned)nc), GTH, smem, st>>>(a, t);
            const unsigned sb = (unsigned)std::min<int64_t>((N >> 6) / 8, 148 * 16);
            gen_segsum_kernel<T><<<dim3(sb < 1 ? 1 : sb, (unsigned)nc), 256, 0, st>>>(a);
            m->launches += 2 + max_steps;
            if (a.mode == QMC_MODE_PROBS) {
                gen_probs_kernel<T><<<148 * 4, 256, 0, st>>>(a);
            } else {
                gen_sample_kernel<T><<<(unsigned)nc, GSEL, 0, st>>>(a);
            }
            m->launches++;
            QMC_CUDA_TRY(cudaGetLastError());
        }
        if (a.mode == QMC_MODE_CHAIN) {
            gen_finish_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(a, nc);
            m->launches++;
        }
        QMC_CUDA_TRY(cudaGetLastError());
    }
    return QMC_OK;
}

int qmc_general_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in, uint64_t chain_id0,
                    uint32_t hop0, double temperature, double g_lo, double g_hi, double dt, uint64_t seed, int dtype,
                    const double *gamma_arr, const int *r_arr, const double *u_arr, const int *group_arr,
                    uint64_t *proposals, uint8_t *accepted, uint64_t *final_state, int64_t *acc_count,
                    int64_t *hamming_hist, void *probs_out, void *amps_out, cudaStream_t st) {
    const int n = m->n;
    QMC_REQUIRE(n > GT && n <= QMC_GENERAL_MAX_SPINS, QMC_ERR_UNSUPPORTED,
                "general statevector path (k-body / custom mixers, complex128) covers %d <= n <= %d (n=%d)", GT + 1,
                QMC_GENERAL_MAX_SPINS, n);
    QMC_REQUIRE(m->n_groups >= 1, QMC_ERR_BADARG, "no mixer set (qmc_mixer_set)");
    if (n_chains == 0) return QMC_OK;
    int rc = general_prepare(m, st);
    if (rc != QMC_OK) return rc;
    GeneralWorkspace *w = m->general;
    int r_cap = qmc_trotter_steps(11, dt);
    if (mode != QMC_MODE_CHAIN) {
        std::vector<int> r_host(n_chains);
        QMC_CUDA_TRY(cudaMemcpyAsync(r_host.data(), r_arr, sizeof(int) * n_chains, cudaMemcpyDeviceToHost, st));
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        r_cap = 1;
        for (int64_t c = 0; c < n_chains; ++c) r_cap = std::max(r_cap, r_host[c]);
    }
    GenArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.n_groups = m->n_groups; a.phase_active = m->phase_active; a.mode = mode; a.max_terms = w->max_terms_per_pass;
    a.alpha = m->alpha; a.dt = dt; a.temperature = temperature; a.g_lo = g_lo; a.g_hi = g_hi;
    a.passes = w->d_passes; a.group_pass_off = w->d_group_pass_off; a.lmask = w->d_lmask; a.weight = w->d_weight;
    a.D = w->dD; a.J = m->dJ; a.h = m->dh; a.group_cum = m->d_group_cum;
    a.seed = seed; a.chain_id0 = chain_id0; a.n_hops = n_hops;
    a.s_in = s_in; a.gamma_arr = gamma_arr; a.u_arr = u_arr; a.r_arr = r_arr; a.group_arr = group_arr;
    a.proposals = proposals; a.accepted = accepted; a.final_state = final_state; a.acc_count = acc_count;
    a.hist = hamming_hist; a.probs_out = probs_out; a.amps_out = amps_out;
    a.visit = mode == QMC_MODE_CHAIN ? m->visit_hist : nullptr;
    if (mode == QMC_MODE_CHAIN && hamming_hist)
        QMC_CUDA_TRY(cudaMemsetAsync(hamming_hist, 0, sizeof(int64_t) * (size_t)n_chains * (size_t)(n + 1), st));
    if (dtype == QMC_DTYPE_F64) return general_run_t<double>(m, w, a, n_chains, n_hops, hop0, r_cap, st);
    return general_run_t<float>(m, w, a, n_chains, n_hops, hop0, r_cap, st);
}
