// small_n.cu -- shared-memory-resident proposal path: the whole 2^n statevector of one chain lives
// in the shared memory of one CTA, which stays resident over all hops of that chain
// (n <= 13 complex64, n <= 12 complex128).  Thousands of chains = thousands of CTAs.
//
// Per hop (quantum_enhanced_mcmc, qumcmc/quantum_mcmc_routines.py:201-225):
//   draws  -> gamma, t -> r, mixer group, u_measure, u_accept         (Philox, common.cuh)
//   K1     -> phase table exp(+i c D(idx)), c = (1-gamma) alpha dt    (fn_qckt_problem_half :54-91)
//   K2     -> U = M (P M)^{r-1}: X-string butterflies + pointwise phase (trotter :95-106, mixers.py)
//   K3     -> block prefix-sum of |a|^2 + inverse CDF                 (QuantumState.sampling :153)
//   K4     -> energy lookup + Metropolis accept                       (test_accept, classical_mcmc_routines.py:28-41)
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

struct SmallParams {
    int n, mode, phase_active, n_groups, max_terms;
    int64_t n_hops;
    uint64_t chain_id0, seed;
    uint32_t hop0;
    double temperature, g_lo, g_hi, dt, alpha;
    const double *D, *E;
    const double *Ex;   // [n_groups][2^n] mixer exponent in the Hadamard basis, or null
    int hx_min_terms;   // groups with at least this many terms run as H . exp(-i gamma dt E_x) . H (0: never)
    const int *group_off;
    const uint64_t *masks;
    const double *weights, *group_cum;
    const uint64_t *s_in;
    const double *gamma_arr;
    const int *r_arr;
    const double *u_arr;
    const int *group_arr;
    uint64_t *proposals;
    uint8_t *accepted;
    uint64_t *final_state;
    int64_t *acc_count, *hamming_hist;
    void *probs_out, *amps_out;
    unsigned long long *visit;  // visit histogram accumulator or null
};

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// exp(i x) with the argument formed and range-reduced in fp64, evaluated in T
__device__ __forceinline__ float2 phase_factor(double x, float) {
    double t = x * 0.31830988618379067154;  // x / pi
    t -= 2.0 * rint(0.5 * t);               // t in [-1, 1]
    float s, c;
    sincospif((float)t, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ double2 phase_factor(double x, double) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(c, s);
}

constexpr int SMALL_NT = 256;

__device__ __forceinline__ float2 cmul_t(const float2 a, const float2 f) { return make_float2(a.x * f.x - a.y * f.y, a.x * f.y + a.y * f.x); }
__device__ __forceinline__ double2 cmul_t(const double2 a, const double2 f) { return make_double2(a.x * f.x - a.y * f.y, a.x * f.y + a.y * f.x); }

// HXK: compiled with the Hadamard-basis mode (third table, swizzled slots); launched when a mixer group qualifies
template <typename T, bool HXK = false>
__global__ void __launch_bounds__(SMALL_NT) smem_chain_kernel(const SmallParams p) {
    using T2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = p.n;
    const int N = 1 << n;
    T2 *st = reinterpret_cast<T2 *>(smraw);
    T2 *ph = st + N;
    T2 *xph = ph + N;                                     // Hadamard-basis mode: 2^-n exp(-i gamma dt E_x(x)) of the proposal
    T2 *cs = HXK ? xph + N : xph;
    int *tm = reinterpret_cast<int *>(cs + p.max_terms);  // term masks of the current group
    // Shared-memory slot of index i.  The passes over qubits 0..3 (butterfly passes of the Hadamard-basis mode, one-body terms
    // taken two at a time) read four elements per thread at strides of 4 and 16 elements: XOR-ing the four low bits with
    // 5 * (bits 4, 5) makes those accesses conflict-free and leaves runs of 16 consecutive indices inside their aligned
    // 16-slot group (4-way conflicts otherwise: n = 12 three-body 2.08 -> 3.20 M proposals/s).
    auto S = [](int i) { return i ^ (((i >> 4) & 3) * 5); };
    __shared__ double s_incl[SMALL_NT];
    __shared__ double s_warp[SMALL_NT / 32];
    __shared__ unsigned long long s_sel;
    __shared__ int s_hist[QMC_MAX_SPINS + 2];

    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    const int64_t chain = blockIdx.x;
    const uint64_t gchain = p.chain_id0 + (uint64_t)chain;

    uint64_t cur;
    if (p.mode == QMC_MODE_CHAIN)
        cur = p.s_in ? p.s_in[chain] : qmc_initial_state(p.seed, gchain, n);
    else
        cur = p.s_in[chain];
    double e_cur = (p.mode == QMC_MODE_CHAIN) ? p.E[cur] : 0.0;
    long long n_acc = 0;
    for (int i = tid; i < n + 1; i += NT) s_hist[i] = 0;
    if (p.mode == QMC_MODE_CHAIN && p.hop0 == 0 && tid == 0) qmc_visit(p.visit, cur);  // entry 0 of markov_chain

    const int64_t hops = (p.mode == QMC_MODE_CHAIN) ? p.n_hops : 1;
    for (int64_t hop = 0; hop < hops; ++hop) {
        // ---- draws (every thread computes the same values; no broadcast needed)
        double gamma, u_meas, u_acc = 0.0;
        int r, grp;
        if (p.mode == QMC_MODE_CHAIN) {
            const HopDraws d = qmc_hop_draws(p.seed, gchain, p.hop0 + (uint32_t)hop);
            gamma = p.g_lo + (p.g_hi - p.g_lo) * d.u_gamma;
            r = qmc_trotter_steps(d.t, p.dt);
            grp = qmc_pick_group(p.n_groups, p.group_cum, d.u_group);
            u_meas = d.u_meas;
            u_acc = d.u_acc;
        } else {
            gamma = p.gamma_arr[chain];
            r = p.r_arr[chain] < 1 ? 1 : p.r_arr[chain];
            grp = p.group_arr ? p.group_arr[chain] : 0;
            u_meas = p.u_arr ? p.u_arr[chain] : 0.0;
        }
        const int t0 = p.group_off[grp], t1 = p.group_off[grp + 1];
        const int nterms = t1 - t0;
        // Hadamard-basis mode: all X strings commute and are diagonal in the x basis, M = H exp(-i gamma dt E_x) H, so a
        // group with many terms costs 2 ceil(n / 2) butterfly passes per layer instead of one pass per two terms
        const bool hx = HXK && nterms >= p.hx_min_terms;
        if (hx) {
            const double cx = -gamma * p.dt;
            const T norm = (T)ldexp(1.0, -n);  // of the two unnormalised transforms around the diagonal
            const double *ex = p.Ex + (size_t)grp * N;
            for (int i = tid; i < N; i += NT) {
                T2 f = phase_factor(cx * ex[i], T(0));
                f.x *= norm;
                f.y *= norm;
                xph[S(i)] = f;
            }
        }

        // ---- per-term rotation constants: exp(-i gamma w dt X_S) -> (cos, sin); masks staged next to them
        for (int k = tid; k < (hx ? 0 : nterms); k += NT) {
            double s, c;
            sincos(gamma * p.weights[t0 + k] * p.dt, &s, &c);
            T2 v;
            v.x = (T)c;
            v.y = (T)s;
            cs[k] = v;
            tm[k] = (int)p.masks[t0 + k];
        }
        // ---- K1: phase table for this proposal (same c for all r-1 layers)
        const bool phase = p.phase_active && r > 1;
        if (phase) {
            const double c = (1.0 - gamma) * p.alpha * p.dt;
            for (int i = tid; i < N; i += NT) ph[S(i)] = phase_factor(c * p.D[i], T(0));
        }
        // ---- |s>
        for (int i = tid; i < N; i += NT) {
            T2 z;
            z.x = (i == (int)cur) ? T(1) : T(0);
            z.y = T(0);
            st[S(i)] = z;
        }
        __syncthreads();

        // ---- K2: M (P M)^{r-1}
        for (int layer = 0; layer < r; ++layer) {
            if (hx) {
                // [P] H Dx H with the diagonals fused into the butterfly passes: P on the loads of the first pass, Dx on the
                // stores of the last pass of the first transform.  Two qubits per pass (4-amplitude orbits in registers).
                const int passes = (n + 1) >> 1;
                for (int half = 0; half < 2; ++half)
                    for (int ps = 0; ps < passes; ++ps) {
                        const T2 *pre = (half == 0 && ps == 0 && layer > 0 && phase) ? ph : nullptr;
                        const T2 *post = (half == 0 && ps == passes - 1) ? xph : nullptr;
                        const int b1 = 2 * ps, b2 = 2 * ps + 1;
                        if (b2 < n) {
                            const int lo_m = (1 << b1) - 1, hi_m = (1 << b2) - 1;
                            for (int q = tid; q < (N >> 2); q += NT) {
                                int i = ((q >> b1) << (b1 + 1)) | (q & lo_m);
                                i = ((i >> b2) << (b2 + 1)) | (i & hi_m);
                                const int i1 = i | (1 << b1), i2 = i | (1 << b2), i3 = i1 | (1 << b2);
                                T2 a0 = st[S(i)], a1 = st[S(i1)], a2 = st[S(i2)], a3 = st[S(i3)];
                                if (pre) { a0 = cmul_t(a0, pre[S(i)]); a1 = cmul_t(a1, pre[S(i1)]); a2 = cmul_t(a2, pre[S(i2)]); a3 = cmul_t(a3, pre[S(i3)]); }
                                T2 s0, s1, d0, d1;
                                s0.x = a0.x + a1.x; s0.y = a0.y + a1.y; d0.x = a0.x - a1.x; d0.y = a0.y - a1.y;
                                s1.x = a2.x + a3.x; s1.y = a2.y + a3.y; d1.x = a2.x - a3.x; d1.y = a2.y - a3.y;
                                a0.x = s0.x + s1.x; a0.y = s0.y + s1.y; a2.x = s0.x - s1.x; a2.y = s0.y - s1.y;
                                a1.x = d0.x + d1.x; a1.y = d0.y + d1.y; a3.x = d0.x - d1.x; a3.y = d0.y - d1.y;
                                if (post) { a0 = cmul_t(a0, post[S(i)]); a1 = cmul_t(a1, post[S(i1)]); a2 = cmul_t(a2, post[S(i2)]); a3 = cmul_t(a3, post[S(i3)]); }
                                st[S(i)] = a0; st[S(i1)] = a1; st[S(i2)] = a2; st[S(i3)] = a3;
                            }
                        } else {  // odd n: the last qubit alone
                            const int lowmask = (1 << b1) - 1;
                            for (int q = tid; q < (N >> 1); q += NT) {
                                const int i = ((q >> b1) << (b1 + 1)) | (q & lowmask), j = i | (1 << b1);
                                T2 a = st[S(i)], b = st[S(j)];
                                if (pre) { a = cmul_t(a, pre[S(i)]); b = cmul_t(b, pre[S(j)]); }
                                T2 oa, ob;
                                oa.x = a.x + b.x; oa.y = a.y + b.y; ob.x = a.x - b.x; ob.y = a.y - b.y;
                                if (post) { oa = cmul_t(oa, post[S(i)]); ob = cmul_t(ob, post[S(j)]); }
                                st[S(i)] = oa; st[S(j)] = ob;
                            }
                        }
                        __syncthreads();
                    }
                continue;
            }
            if (layer > 0 && phase) {
                for (int i = tid; i < N; i += NT) {
                    const T2 a = st[S(i)], f = ph[S(i)];
                    T2 o;
                    o.x = a.x * f.x - a.y * f.y;
                    o.y = a.x * f.y + a.y * f.x;
                    st[S(i)] = o;
                }
                __syncthreads();
            }
            // Terms two at a time: X_S1, X_S2 (S1 != S2) act on orbits {i, i^S1, i^S2, i^S1^S2}; a thread owns whole
            // orbits, so both rotations run in registers between one load and one store of the four amplitudes --
            // half the shared-memory traffic and half the barriers, the same arithmetic per amplitude in the same order.
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
                    for (int q = tid; q < (N >> 2); q += NT) {
                        int i = ((q >> lo_b) << (lo_b + 1)) | (q & lo_m);   // orbit representative: bits p1, p2 clear
                        i = ((i >> hi_b) << (hi_b + 1)) | (i & hi_m);
                        const int i1 = i ^ m1, i2 = i ^ m2, i3 = i1 ^ m2;
                        T2 a0 = st[S(i)], a1 = st[S(i1)], a2 = st[S(i2)], a3 = st[S(i3)];
                        T2 t0_, t1_;
                        // term k on (a0, a1) and (a2, a3)
                        t0_.x = c1 * a0.x + s1 * a1.y; t0_.y = c1 * a0.y - s1 * a1.x;
                        t1_.x = c1 * a1.x + s1 * a0.y; t1_.y = c1 * a1.y - s1 * a0.x;
                        a0 = t0_; a1 = t1_;
                        t0_.x = c1 * a2.x + s1 * a3.y; t0_.y = c1 * a2.y - s1 * a3.x;
                        t1_.x = c1 * a3.x + s1 * a2.y; t1_.y = c1 * a3.y - s1 * a2.x;
                        a2 = t0_; a3 = t1_;
                        // term k+1 on (a0, a2) and (a1, a3)
                        t0_.x = c2 * a0.x + s2 * a2.y; t0_.y = c2 * a0.y - s2 * a2.x;
                        t1_.x = c2 * a2.x + s2 * a0.y; t1_.y = c2 * a2.y - s2 * a0.x;
                        a0 = t0_; a2 = t1_;
                        t0_.x = c2 * a1.x + s2 * a3.y; t0_.y = c2 * a1.y - s2 * a3.x;
                        t1_.x = c2 * a3.x + s2 * a1.y; t1_.y = c2 * a3.y - s2 * a1.x;
                        a1 = t0_; a3 = t1_;
                        st[S(i)] = a0; st[S(i1)] = a1; st[S(i2)] = a2; st[S(i3)] = a3;
                    }
                    __syncthreads();
                    k += 2;
                    continue;
                }
                const int tb = 31 - __clz(m1);
                const int lowmask = (1 << tb) - 1;
                const T c = cs[k].x, s = cs[k].y;
                for (int q = tid; q < (N >> 1); q += NT) {
                    const int i = ((q >> tb) << (tb + 1)) | (q & lowmask);
                    const int j = i ^ m1;
                    const T2 a = st[S(i)], b = st[S(j)];
                    T2 oa, ob;
                    oa.x = c * a.x + s * b.y;
                    oa.y = c * a.y - s * b.x;
                    ob.x = c * b.x + s * a.y;
                    ob.y = c * b.y - s * a.x;
                    st[S(i)] = oa;
                    st[S(j)] = ob;
                }
                __syncthreads();
                k += 1;
            }
        }

        // ---- K3: measurement. contiguous chunk per thread, fp64 accumulation, block scan
        const int chunk = (N + NT - 1) / NT;
        const int i0 = tid * chunk;
        const int i1 = min(N, i0 + chunk);
        double local = 0.0;
        for (int i = i0; i < i1; ++i) {
            const T2 a = st[S(i)];
            local += (double)(a.x * a.x + a.y * a.y);
        }
        double incl = local;
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        double base = 0.0;
        for (int w = 0; w < warp; ++w) base += s_warp[w];
        incl += base;
        s_incl[tid] = incl;
        __syncthreads();
        const double total = s_incl[NT - 1];
        const double target = u_meas * total;
        if (p.mode == QMC_MODE_PROBS) {
            // normalised output: fp32 rotation constants drift the norm by ~1e-7 per gate
            T *po = reinterpret_cast<T *>(p.probs_out);
            T2 *ao = reinterpret_cast<T2 *>(p.amps_out);
            const T inv = (T)(1.0 / total), inva = (T)(1.0 / sqrt(total));
            for (int i = tid; i < N; i += NT) {
                const T2 a = st[S(i)];
                if (po) po[i] = (a.x * a.x + a.y * a.y) * inv;
                if (ao) {
                    T2 o;
                    o.x = a.x * inva;
                    o.y = a.y * inva;
                    ao[i] = o;
                }
            }
            return;
        }
        const double lo = tid == 0 ? 0.0 : s_incl[tid - 1];
        const double hi = incl;
        const bool last_active = (i1 == N) && (i0 < N);
        if (i0 < N && target >= lo && (target < hi || last_active)) {
            double c = lo;
            int sel = i1 - 1;
            for (int i = i0; i < i1; ++i) {
                const T2 a = st[S(i)];
                c += (double)(a.x * a.x + a.y * a.y);
                if (target < c) {
                    sel = i;
                    break;
                }
            }
            s_sel = (unsigned long long)sel;
        }
        __syncthreads();
        const uint64_t sp = s_sel;

        if (p.mode == QMC_MODE_PROPOSE) {
            if (tid == 0) p.proposals[chain] = sp;
            return;
        }

        // ---- K4: Metropolis accept (all threads evaluate identically)
        const double e_sp = p.E[sp];
        const bool acc = qmc_accept_dev(e_cur, e_sp, p.temperature, u_acc);
        if (tid == 0) {
            if (p.proposals) p.proposals[chain * p.n_hops + hop] = sp;
            if (p.accepted) p.accepted[chain * p.n_hops + hop] = (uint8_t)acc;
            s_hist[__popcll(cur ^ sp)] += 1;
        }
        if (acc) {
            cur = sp;
            e_cur = e_sp;
            ++n_acc;
        }
        if (tid == 0) qmc_visit(p.visit, cur);
        __syncthreads();  // st / s_sel reuse in the next hop
    }

    if (tid == 0) {
        if (p.final_state) p.final_state[chain] = cur;
        if (p.acc_count) p.acc_count[chain] = n_acc;
    }
    __syncthreads();
    if (p.hamming_hist)
        for (int i = tid; i < n + 1; i += NT) p.hamming_hist[chain * (n + 1) + i] = s_hist[i];
}

// ---------------------------------------------------------------------------------------------
// Warp-per-chain kernel: 8 <= n <= 10, one-body X mixers.  One WARP owns a chain for all its hops: the 2^n amplitudes
// sit in registers (R = n - 5 register qubits, 2^R amplitudes per lane; index = lane << R | j in layout 1), so a hop needs
// no __syncthreads at all -- the CTA-per-chain kernel above spends most of a hop in ~60 block barriers.
//   * storage basis S(idx) = (-i)^popcount(idx) a(idx) (sweep.cu): exp(-i theta X_q) is a real rotation, tan form
//     S_lo' = S_lo + t S_hi, S_hi' = S_hi - t S_lo (2 FMA per amplitude per qubit), (prod cos)^r folded into |s>;
//   * register qubits: butterflies inside the lane; then one transposition through a warp-private shared-memory tile
//     (XOR-swizzled, __syncwarp only) makes the top R qubits register qubits (layout 2: index = j << 5 | lane); the 5 - R
//     qubits that are lane bits in both layouts go through __shfl_xor.  Layers alternate the layout they start in;
//   * phase factors exp(+i c D(idx)) of the proposal: computed once per hop into a second warp-private tile, read in
//     either layout;
//   * K3: per-lane sums in index order (layout 1: 2^R consecutive indices per lane) -> warp scan -> owner lane -> index;
//     K4: energy table + Metropolis; the Hamming histogram lives in lane k = distance k.
// Several chains per CTA (one per warp), no interaction between the warps.
// ---------------------------------------------------------------------------------------------
struct WarpParams {
    SmallParams p;
    const double *qweight;  // [n_groups][n] summed one-body weights
    int64_t n_chains;
};

template <int R>
__device__ __forceinline__ int wswz(int idx) { return idx ^ ((idx >> R) & 31); }

// Register form of an amplitude inside warp_layer: complex64 amplitudes are packed into one 64-bit register pair so that a
// butterfly is two packed FMAs (fma.rn.f32x2 -> FFMA2) instead of four scalar ones; complex128 stays double2.
typedef unsigned long long W64;
template <typename T> struct WReg;
template <> struct WReg<float> { using type = W64; };
template <> struct WReg<double> { using type = double2; };
__device__ __forceinline__ W64 w_pack(float2 a) {
    W64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ double2 w_pack(double2 a) { return a; }
__device__ __forceinline__ float2 w_unpack(W64 r) {
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}
__device__ __forceinline__ double2 w_unpack(double2 r) { return r; }
__device__ __forceinline__ W64 w_fma(float t, W64 b, W64 c) {  // t * b + c on both halves
    W64 d, tt;
    asm("mov.b64 %0, {%1, %1};" : "=l"(tt) : "f"(t));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(tt), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ double2 w_fma(double t, double2 b, double2 c) { return make_double2(fma(t, b.x, c.x), fma(t, b.y, c.y)); }
__device__ __forceinline__ W64 w_shfl_xor(W64 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double2 w_shfl_xor(double2 v, int m) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// One mixer layer (preceded by the phase layer when `apply_phase`) that starts in layout 1 (L1 = true: registers = qubits
// 0..R-1, index = lane << R | j) or layout 2 (registers = qubits 5..n-1, index = j << 5 | lane) and ends in the other one.
// The layout is a template parameter, so every shared-memory slot is a compile-time function of j plus one lane term.
template <typename T, int R, bool L1>
__device__ __forceinline__ void warp_layer(typename Vec2<T>::type (&v)[1 << R], const T my_t,
                                           typename Vec2<T>::type *tile, const typename Vec2<T>::type *ph, bool apply_phase,
                                           int lane) {  // my_t: lane q holds tan(theta_q), fetched by shuffle where it is used
    using T2 = typename Vec2<T>::type;
    constexpr int NR = 1 << R;
    constexpr int LEFT = 5 - R;
    const unsigned FULL = 0xffffffffu;
    auto idx_of = [&](bool l1, int j) { return l1 ? ((lane << R) | j) : ((j << 5) | lane); };
    using WR = typename WReg<T>::type;
    WR *wtile = reinterpret_cast<WR *>(tile);
    WR w[NR];
    if (apply_phase) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const T2 f = ph[wswz<R>(idx_of(L1, j))], a = v[j];
            T2 o;
            o.x = a.x * f.x - a.y * f.y;
            o.y = a.x * f.y + a.y * f.x;
            w[j] = w_pack(o);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NR; ++j) w[j] = w_pack(v[j]);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const bool l1 = half == 0 ? L1 : !L1;
#pragma unroll
        for (int b = 0; b < R; ++b) {  // register qubits of the current layout
            const T t = __shfl_sync(FULL, my_t, l1 ? b : 5 + b);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (j & (1 << b)) continue;
                const WR lo = w[j], hi = w[j | (1 << b)];
                w[j] = w_fma(t, hi, lo);
                w[j | (1 << b)] = w_fma(-t, lo, hi);
            }
        }
        if (half == 0) {
#pragma unroll
            for (int k = 0; k < LEFT; ++k) {  // qubits R..4: lane bits in both layouts (layout 1: lane bit q - R; layout 2: q)
                const int q = R + k;
                const int lb = l1 ? k : q;
                const T tq_ = __shfl_sync(FULL, my_t, q);
                const T t = ((lane >> lb) & 1) ? -tq_ : tq_;
#pragma unroll
                for (int j = 0; j < NR; ++j) w[j] = w_fma(t, w_shfl_xor(w[j], 1 << lb), w[j]);
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) wtile[wswz<R>(idx_of(l1, j))] = w[j];  // transposition through the warp's tile
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NR; ++j) w[j] = wtile[wswz<R>(idx_of(!l1, j))];
            __syncwarp();
        }
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = w_unpack(w[j]);
}

template <typename T, int R, bool HX = false>
__global__ void __launch_bounds__(128) warp_chain_kernel(const WarpParams wp) {
    using T2 = typename Vec2<T>::type;
    constexpr int NR = 1 << R;       // amplitudes per lane
    constexpr int n = R + 5;
    constexpr int N = 1 << n;
    constexpr int LEFT = 5 - R;      // qubits R .. 4 are lane bits in both layouts
    extern __shared__ __align__(16) unsigned char smraw[];
    const SmallParams &p = wp.p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (chain >= wp.n_chains) return;  // whole warp
    T2 *tile = reinterpret_cast<T2 *>(smraw) + (size_t)warp * (HX ? 3 : 2) * N;  // transposition tile
    T2 *ph = tile + N;                                                 // phase factors of the current proposal
    T2 *xph = ph + N;                                                  // HX: 2^-n exp(-i gamma dt E_x(x)) of the proposal
    const uint64_t gchain = p.chain_id0 + (uint64_t)chain;
    const unsigned FULL = 0xffffffffu;

    uint64_t cur;
    if (p.mode == QMC_MODE_CHAIN) cur = p.s_in ? p.s_in[chain] : qmc_initial_state(p.seed, gchain, n);
    else cur = p.s_in[chain];
    double e_cur = (p.mode == QMC_MODE_CHAIN) ? p.E[cur] : 0.0;
    long long n_acc = 0;
    int my_hist = 0;  // lane k: proposals at Hamming distance k
    if (p.mode == QMC_MODE_CHAIN && p.hop0 == 0 && lane == 0) qmc_visit(p.visit, cur);

    const int64_t hops = (p.mode == QMC_MODE_CHAIN) ? p.n_hops : 1;
    for (int64_t hop = 0; hop < hops; ++hop) {
        double gamma, u_meas, u_acc = 0.0;
        int r, grp;
        if (p.mode == QMC_MODE_CHAIN) {
            const HopDraws d = qmc_hop_draws(p.seed, gchain, p.hop0 + (uint32_t)hop);
            gamma = p.g_lo + (p.g_hi - p.g_lo) * d.u_gamma;
            r = qmc_trotter_steps(d.t, p.dt);
            grp = qmc_pick_group(p.n_groups, p.group_cum, d.u_group);
            u_meas = d.u_meas;
            u_acc = d.u_acc;
        } else {
            gamma = p.gamma_arr[chain];
            r = p.r_arr[chain] < 1 ? 1 : p.r_arr[chain];
            grp = p.group_arr ? p.group_arr[chain] : 0;
            u_meas = p.u_arr ? p.u_arr[chain] : 0.0;
        }
        // per-qubit rotation: lane q holds tan(theta_q), theta_q = gamma w_q dt; scale = prod cos
        double my_t = 0.0, my_c = 1.0;
        if (HX) {
            my_t = lane < n ? 1.0 : 0.0;
            const double cx = -gamma * p.dt;
            const T norm = (T)ldexp(1.0, -n);
            const double *ex = p.Ex + (size_t)grp * N;
#pragma unroll 4
            for (int j = 0; j < NR; ++j) {
                const int idx = (lane << R) | j;
                T2 f = phase_factor(cx * ex[idx], T(0));
                f.x *= norm;
                f.y *= norm;
                xph[wswz<R>(idx)] = f;
            }
        } else if (lane < n) {
            const double wq = wp.qweight[grp * n + lane];
            if (wq != 0.0) {
                double sn, cs;
                sincos(gamma * wq * p.dt, &sn, &cs);
                my_t = sn / cs;
                my_c = cs;
            }
        }
        double scale = my_c;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) scale *= __shfl_xor_sync(FULL, scale, off);
        scale = HX ? 1.0 : pow(scale, (double)r);
        const T tq = (T)my_t;  // lane q: tan(theta_q)
        // phase factors of this proposal (same c for all r - 1 layers), layout-independent slots
        const bool phase = p.phase_active && r > 1;
        if (phase) {
            const double c = (1.0 - gamma) * p.alpha * p.dt;
#pragma unroll 4
            for (int j = 0; j < NR; ++j) {  // four in flight (a single warp has nothing else to hide the sincos latency), not 2^R copies
                const int idx = (lane << R) | j;
                ph[wswz<R>(idx)] = phase_factor(c * p.D[idx], T(0));
            }
        }
        // |s> in the storage basis: (prod cos)^r (-i)^popcount(s) at idx = s
        T2 v[NR];
        {
            const int pc = HX ? 0 : (__popcll(cur) & 3);  // HX: plain amplitudes
            const T mag = (T)scale;
            T2 s0;
            s0.x = pc == 0 ? mag : (pc == 2 ? -mag : T(0));
            s0.y = pc == 3 ? mag : (pc == 1 ? -mag : T(0));  // (-i)^1 = -i, (-i)^3 = +i
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const bool hit = (int)cur == ((lane << R) | j);
                v[j].x = hit ? s0.x : T(0);
                v[j].y = hit ? s0.y : T(0);
            }
        }
        __syncwarp();
        // layers alternate the layout they start in (two instantiations of the layer body: the code must stay inside the
        // instruction cache when hundreds of warps run different hops)
#pragma unroll 1
        for (int layer = 0; layer < r; ++layer) {
            if (HX) {
                warp_layer<T, R, true>(v, tq, tile, ph, phase && layer > 0, lane);   // [P] then into the x basis
                warp_layer<T, R, false>(v, -tq, tile, xph, true, lane);              // mixer diagonal, then back
            } else if (layer & 1) {
                warp_layer<T, R, false>(v, tq, tile, ph, phase, lane);
            } else {
                warp_layer<T, R, true>(v, tq, tile, ph, phase && layer > 0, lane);
            }
        }
        if (!HX && (r & 1)) {  // odd r ends in layout 2: back to layout 1 (2^R consecutive indices per lane) for the measurement
#pragma unroll
            for (int j = 0; j < NR; ++j) tile[wswz<R>((j << 5) | lane)] = v[j];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NR; ++j) v[j] = tile[wswz<R>((lane << R) | j)];
            __syncwarp();
        }
        // ---- K3
        double local = 0.0;
#pragma unroll
        for (int j = 0; j < NR; ++j) local += (double)(v[j].x * v[j].x + v[j].y * v[j].y);
        double incl = local;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double x = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += x;
        }
        const double total = __shfl_sync(FULL, incl, 31);
        if (p.mode == QMC_MODE_PROBS) {
            T *po = reinterpret_cast<T *>(p.probs_out);
            T2 *ao = reinterpret_cast<T2 *>(p.amps_out);
            const T inv = (T)(1.0 / total), inva = (T)(1.0 / sqrt(total));
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const int idx = (lane << R) | j;
                if (po) po[idx] = (v[j].x * v[j].x + v[j].y * v[j].y) * inv;
                if (ao) {  // a = i^popcount(idx) S
                    const int k = HX ? 0 : (__popc(idx) & 3);
                    T2 o;
                    o.x = (k == 0 ? v[j].x : k == 1 ? -v[j].y : k == 2 ? -v[j].x : v[j].y) * inva;
                    o.y = (k == 0 ? v[j].y : k == 1 ? v[j].x : k == 2 ? -v[j].y : -v[j].x) * inva;
                    ao[idx] = o;
                }
            }
            return;
        }
        const double target = u_meas * total;
        const double lo = incl - local;
        const unsigned owners = __ballot_sync(FULL, target < incl);
        const int owner = owners ? (__ffs(owners) - 1) : 31;
        int sel = NR - 1;
        {
            double c = lo;
            bool found = false;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                c += (double)(v[j].x * v[j].x + v[j].y * v[j].y);
                if (!found && target < c) {
                    sel = j;
                    found = true;
                }
            }
        }
        const uint64_t sp = (uint64_t)((owner << R) | __shfl_sync(FULL, sel, owner));
        if (p.mode == QMC_MODE_PROPOSE) {
            if (lane == 0) p.proposals[chain] = sp;
            return;
        }
        // ---- K4 (every lane evaluates identically)
        const double e_sp = p.E[sp];
        const bool acc = qmc_accept_dev(e_cur, e_sp, p.temperature, u_acc);
        if (lane == 0) {
            if (p.proposals) p.proposals[chain * p.n_hops + hop] = sp;
            if (p.accepted) p.accepted[chain * p.n_hops + hop] = (uint8_t)acc;
        }
        if (lane == __popcll(cur ^ sp)) ++my_hist;
        if (acc) {
            cur = sp;
            e_cur = e_sp;
            ++n_acc;
        }
        if (lane == 0) qmc_visit(p.visit, cur);
    }
    if (lane == 0) {
        if (p.final_state) p.final_state[chain] = cur;
        if (p.acc_count) p.acc_count[chain] = n_acc;
    }
    if (p.hamming_hist && lane <= n) p.hamming_hist[chain * (n + 1) + lane] = my_hist;
}

template <typename T, int R, bool HX = false>
static int launch_warp_kernel(const WarpParams &wp, int64_t n_chains, cudaStream_t st) {
    const size_t per_warp = (HX ? 3 : 2) * ((size_t)1 << (R + 5)) * sizeof(typename Vec2<T>::type);
    int warps = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
    if (n_chains < warps) warps = (int)n_chains;
    const size_t smem = per_warp * warps;
    QMC_CUDA_TRY(cudaFuncSetAttribute(warp_chain_kernel<T, R, HX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * 4)));
    warp_chain_kernel<T, R, HX><<<(unsigned)((n_chains + warps - 1) / warps), warps * 32, smem, st>>>(wp);
    return QMC_OK;
}

int qmc_small_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in,
                  uint64_t chain_id0, uint32_t hop0, double temperature, double g_lo, double g_hi, double dt,
                  uint64_t seed, int dtype, const double *gamma_arr, const int *r_arr, const double *u_arr,
                  const int *group_arr, uint64_t *proposals, uint8_t *accepted, uint64_t *final_state,
                  int64_t *acc_count, int64_t *hamming_hist, void *probs_out, void *amps_out, cudaStream_t st) {
    SmallParams p;
    p.n = m->n;
    p.mode = mode;
    p.phase_active = m->phase_active;
    p.n_groups = m->n_groups;
    p.max_terms = m->max_terms_in_group;
    p.n_hops = n_hops;
    p.chain_id0 = chain_id0;
    p.seed = seed;
    p.hop0 = hop0;
    p.temperature = temperature;
    p.g_lo = g_lo;
    p.g_hi = g_hi;
    p.dt = dt;
    p.alpha = m->alpha;
    p.D = m->dD;
    p.E = m->dE;
    p.Ex = m->dEx;
    p.hx_min_terms = 0;
    p.group_off = m->d_group_off;
    p.masks = m->d_masks;
    p.weights = m->d_weights;
    p.group_cum = m->d_group_cum;
    p.s_in = s_in;
    p.gamma_arr = gamma_arr;
    p.r_arr = r_arr;
    p.u_arr = u_arr;
    p.group_arr = group_arr;
    p.proposals = proposals;
    p.accepted = accepted;
    p.final_state = final_state;
    p.acc_count = acc_count;
    p.hamming_hist = hamming_hist;
    p.probs_out = probs_out;
    p.amps_out = amps_out;
    p.visit = mode == QMC_MODE_CHAIN ? m->visit_hist : nullptr;

    const size_t N = (size_t)1 << m->n;
    const size_t esz = dtype == QMC_DTYPE_F64 ? sizeof(double2) : sizeof(float2);
    size_t smem = (2 * N + (size_t)m->max_terms_in_group) * esz + sizeof(int) * (size_t)m->max_terms_in_group;
    // Hadamard-basis mode for groups with many terms: a layer costs 2 ceil(n / 2) butterfly passes, the term loop one pass
    // per two terms -- worth it from 2 n + 4 terms on (two-body mixers from n = 6, three-body from n = 5); needs a third
    // 2^n table in shared memory.  QUMCMC_SMALL_HX=0 keeps the term loop, =k sets the threshold.
    if (m->dEx && m->max_terms_in_group >= 2 * m->n + 4) {
        int thr = 2 * m->n + 4;
        if (const char *e = getenv("QUMCMC_SMALL_HX")) thr = atoi(e);
        const size_t smem_hx = smem + N * esz;
        if (thr > 0 && smem_hx <= 227 * 1024) {
            p.hx_min_terms = thr;
            smem = smem_hx;
        }
    }
    QMC_REQUIRE(smem <= 227 * 1024, QMC_ERR_UNSUPPORTED,
                "shared-memory path: n=%d with %d mixer terms needs %zu B of shared memory (> 227 KiB)", m->n,
                m->max_terms_in_group, smem);
    QMC_REQUIRE(n_chains <= 0x7fffffffLL, QMC_ERR_BADARG, "too many chains for one launch");
    if (n_chains == 0) return QMC_OK;
    // one-body mixers at 8 <= n <= 10: one warp per chain, state in registers, no block barriers (QUMCMC_NO_WARP_KERNEL=1
    // keeps the CTA-per-chain kernel for A/B runs)
    if (m->all_one_body && m->n >= 8 && m->n <= 10 && m->dD && m->dE && !getenv("QUMCMC_NO_WARP_KERNEL")) {
        WarpParams wp;
        wp.p = p;
        wp.qweight = m->d_qweight;
        wp.n_chains = n_chains;
        int rcw = QMC_OK;
        const bool f64 = dtype == QMC_DTYPE_F64;
        switch (m->n) {
            case 8: rcw = f64 ? launch_warp_kernel<double, 3>(wp, n_chains, st) : launch_warp_kernel<float, 3>(wp, n_chains, st); break;
            case 9: rcw = f64 ? launch_warp_kernel<double, 4>(wp, n_chains, st) : launch_warp_kernel<float, 4>(wp, n_chains, st); break;
            default: rcw = f64 ? launch_warp_kernel<double, 5>(wp, n_chains, st) : launch_warp_kernel<float, 5>(wp, n_chains, st); break;
        }
        if (rcw != QMC_OK) return rcw;
        m->launches++;
        QMC_CUDA_TRY(cudaGetLastError());
        return QMC_OK;
    }
    // any other mixer at 8 <= n <= 10: the same warp-per-chain kernel in Hadamard-basis mode (two warp_layer calls per Trotter
    // layer whatever the number of terms); QUMCMC_NO_WARP_KERNEL=1 or QUMCMC_SMALL_HX=0 keep the CTA-per-chain kernel
    if (!m->all_one_body && m->dEx && m->n >= 8 && m->n <= 10 && m->dD && m->dE && !getenv("QUMCMC_NO_WARP_KERNEL") &&
        !(getenv("QUMCMC_SMALL_HX") && atoi(getenv("QUMCMC_SMALL_HX")) == 0)) {
        WarpParams wp;
        wp.p = p;
        wp.qweight = nullptr;
        wp.n_chains = n_chains;
        int rcw = QMC_OK;
        const bool f64 = dtype == QMC_DTYPE_F64;
        switch (m->n) {
            case 8: rcw = f64 ? launch_warp_kernel<double, 3, true>(wp, n_chains, st) : launch_warp_kernel<float, 3, true>(wp, n_chains, st); break;
            case 9: rcw = f64 ? launch_warp_kernel<double, 4, true>(wp, n_chains, st) : launch_warp_kernel<float, 4, true>(wp, n_chains, st); break;
            default: rcw = f64 ? launch_warp_kernel<double, 5, true>(wp, n_chains, st) : launch_warp_kernel<float, 5, true>(wp, n_chains, st); break;
        }
        if (rcw != QMC_OK) return rcw;
        m->launches++;
        QMC_CUDA_TRY(cudaGetLastError());
        return QMC_OK;
    }
    // threads per chain: one 4-amplitude orbit per thread keeps every lane busy in the term loop (2^n / 4), within
    // [64, 256]; QUMCMC_SMALL_NT overrides for A/B runs
    int threads = (int)std::min<size_t>(SMALL_NT, std::max<size_t>(64, N >> 2));
    if (const char *e = getenv("QUMCMC_SMALL_NT")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) threads = v;
    }
    const bool hxk = p.hx_min_terms > 0;
    if (dtype == QMC_DTYPE_F64) {
        auto kern = hxk ? smem_chain_kernel<double, true> : smem_chain_kernel<double, false>;
        QMC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)n_chains, threads, smem, st>>>(p);
    } else {
        auto kern = hxk ? smem_chain_kernel<float, true> : smem_chain_kernel<float, false>;
        QMC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)n_chains, threads, smem, st>>>(p);
    }
    m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}
