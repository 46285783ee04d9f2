// general.cu -- HBM statevector path for everything the specialised kernels do not cover: arbitrary X-string mixers
// (GenericMixer k-body, CustomMixer, CoherentMixerSum, IncoherentMixerSum: qumcmc/mixers.py:21-151) above the
// shared-memory limit, and complex128 statevectors the fp64 sweep kernels do not take.  13 <= n <= QMC_GENERAL_MAX_SPINS.
//
// The reference pays one pass over the statevector per gate (qumcmc/quantum_mcmc_routines.py:54-106): a two-body mixer
// alone has n(n-1)/2 gates per Trotter layer.  All X-strings commute and are diagonal in the Hadamard basis:
//     M = prod_S exp(-i gamma w_S dt X_S) = H^(x)n  diag( exp(-i gamma dt E_x(x)) )  H^(x)n,   E_x(x) = sum_S w_S chi_S(x),
// chi_S(x) = (-1)^popcount(x & S).  The phase layer P = diag(exp(+i c D(idx))) has the same form in the computational basis
// (D = sum_{j<k} J z_j z_k + sum_j h z_j, a Walsh polynomial of degree 2).  So U = M (P M)^{r-1} is an alternation of
// Walsh-Hadamard transforms and two kinds of diagonals, independent of the NUMBER of mixer terms:
//     U = H Dx H (P H Dx H)^{r-1}.
// H factorises over qubit groups (G0 = qubits 0..11, then groups of 8 from the top); a pass over the tiles of one group
// applies  H_g  diag  H_g  ("host" pass: the diagonal needs every group's H on either side, the host's own ones are fused
// into the pass), the other groups run plain H passes in between.  Hosts alternate: G1 hosts Dx, G0 hosts P, so with two
// groups (n <= 20) a Trotter layer costs TWO passes whatever the mixer is (a one-body layer costs one sweep in sweep.cu);
// with K + 1 groups 2 K passes.  The first pass starts from H|s> analytically (write-only), the last one (always G0)
// leaves the amplitudes in the computational basis and emits the |a|^2 segment sums.
//
// A tile = 12 index positions (always 0..3: HBM runs of 16 amplitudes) handled by a 256-thread CTA; each thread works
// on 16 amplitudes in registers, in three register stages (local bits 8..11 / 4..7 / 0..3) joined by XOR-swizzled
// shared-memory transpositions; the first and last stage read / write HBM directly.
// Diagonals are evaluated without any 2^n table: for a tile with fixed outside bits B,
//     E(l; B) = A(B) + sum_t z_t F_t(B) + R_T(l) [+ residual terms of weight >= 3 through a Walsh transform of their
//     tile-local coefficients], A, F_t from O(n^2) fp64 work per CTA, R_T a 4096-entry table per (tile, diagonal).
// Measurement: 64-amplitude segment sums (fp64) -> block inverse CDF -> the 64 amplitudes of the chosen segment.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "common.cuh"

namespace {

constexpr int XT = 12;           // tile positions
constexpr int XN = 1 << XT;      // amplitudes per tile
constexpr int XTH = 256;         // threads per CTA
constexpr int XR = 16;           // amplitudes per thread (4 register bits per stage)
constexpr int XFORCED = 4;       // index positions 0..3 belong to every tile
constexpr int GSEL = 128;        // threads of the sampler
constexpr int XMAXG = 6;         // qubit groups: G0 (12 qubits) + up to 5 groups of 8

enum { X_FIRST = 0, X_HOST = 1, X_HONLY = 2, X_FINAL = 3 };

struct GenChain {  // per chain, rewritten every hop
    unsigned long long cur;
    double gamma, e_cur, u_meas, u_acc;
    long long n_acc;
    int r, grp;
};

struct XGroup {   // tile geometry of one qubit group
    unsigned char tpos[XT];   // ascending index positions of the tile
    unsigned short active;    // local bits that are the group's qubits
    int n_active;
    unsigned long long tmask; // union of the tile positions
};

// A diagonal exp(i theta E(idx)), E(idx) = sum_S w_S prod_{q in S} z_q, z_q = +1 / -1 for index bit q clear / set, as
// hosted on one tile geometry: terms of weight <= 2 as a linear + symmetric pair form, the rest grouped by their
// tile-local pattern.
struct XDiag {
    const double *lin;             // [n]
    const double *quad;            // [n * n], symmetric, zero diagonal
    const double *R;               // [4096] sum_{t < t'} quad[p_t][p_t'] z_t z_t' over the tile positions
    int n_mu;                      // residual: distinct tile-local patterns
    const int *mu_off;             // [n_mu + 1] term ranges
    const unsigned short *mu;      // [n_mu] local pattern
    const unsigned long long *omask;  // per term: positions outside the tile
    const double *w;               // per term
};

struct GenArgs {
    int n, n_groups, phase_active, mode, K;   // K = number of HI groups (qubit groups = K + 1)
    double alpha, dt, temperature, g_lo, g_hi;
    XGroup groups[XMAXG];
    const XDiag *diag_p;     // P hosted on G0
    const XDiag *diag_x;     // [n_groups] Dx of every mixer group, hosted on G1
    const double *J, *h, *group_cum;
    GenChain *chains;
    void *state;             // [batch][2^n] complex
    double *segsum;          // [batch][2^n / 64]
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

// Register representation of one amplitude.  complex64 amplitudes stay packed in one 64-bit register pair so that the
// Walsh-Hadamard butterflies are the packed FP32 adds of sm_100 (add/sub.rn.f32x2 -> SASS FADD2: one instruction per
// complex add instead of two); complex128 amplitudes are plain double2.
typedef unsigned long long A64;
template <typename T> struct Reg;
template <> struct Reg<float> { using type = A64; };
template <> struct Reg<double> { using type = double2; };
__device__ __forceinline__ A64 r_make(float x, float y) {
    A64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ double2 r_make(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ float2 r_get(A64 r) {
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}
__device__ __forceinline__ double2 r_get(double2 r) { return r; }
__device__ __forceinline__ void r_addsub(A64 &a, A64 &b) {  // (a, b) <- (a + b, a - b)
    A64 s, d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(a), "l"(b));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    a = s;
    b = d;
}
__device__ __forceinline__ void r_addsub(double2 &a, double2 &b) {
    const double2 u = a, w = b;
    a = make_double2(u.x + w.x, u.y + w.y);
    b = make_double2(u.x - w.x, u.y - w.y);
}
__device__ __forceinline__ A64 r_scale(A64 a, float s) {
    A64 d;
    const A64 s2 = r_make(s, s);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(s2));
    return d;
}
__device__ __forceinline__ double2 r_scale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

// exp(2 pi i theta E) for a phase in full turns (theta = angle per unit of E / 2 pi).
// complex64: the product and the range reduction are ONE fp64 FMA -- theta * E + 1.5 * 2^20 has an ulp of 2^-32, so the
// low mantissa word of the sum is frac(theta E) * 2^32 as a two's-complement integer (|theta E| < 2^19 turns); sine and
// cosine of that angle come from the special-function unit (sin/cos.approx: absolute error < 5e-7 on [-pi, pi], the
// unit and error the complex64 sweep kernels use for their phase layer).  complex128: sincospi in fp64.
__device__ __forceinline__ float2 gen_phase_turns(double theta, double E, float) {
    const int fx = __double2loint(fma(theta, E, 1572864.0));
    const float x = (float)fx * 1.4629180792671596e-9f;  // 2 pi / 2^32
    float s, c;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
    return make_float2(c, s);
}
__device__ __forceinline__ double2 gen_phase_turns(double theta, double E, double) {
    double s, c;
    sincospi(2.0 * theta * E, &s, &c);
    return make_double2(c, s);
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

// ---------------------------------------------------------------------------------------------
// register stages: stage s holds local bits [4 s, 4 s + 4) in registers
//   stage 2: l = tid | j << 8          stage 1: l = (tid & 15) | j << 4 | (tid >> 4) << 8          stage 0: l = tid << 4 | j
// shared-memory slot of local index l: XOR the four low bits with bits 4..7 (conflict-free for all three stages, 64- and
// 128-bit elements; checked exhaustively in tests/test_smem_layouts.py through qmc_debug_general_slot)
// ---------------------------------------------------------------------------------------------
template <int S>
__host__ __device__ __forceinline__ int xl(int tid, int j) {
    return S == 2 ? (tid | (j << 8)) : (S == 1 ? ((tid & 15) | (j << 4) | ((tid >> 4) << 8)) : ((tid << 4) | j));
}
__host__ __device__ __forceinline__ int xswz(int l) { return l ^ ((l >> 4) & 15); }

template <typename R>
__device__ __forceinline__ void wht4(R (&v)[XR], int mask) {  // Walsh-Hadamard butterflies on the register bits in `mask`
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (!((mask >> b) & 1)) continue;  // uniform over the CTA
#pragma unroll
        for (int j = 0; j < XR; ++j) {
            if (j & (1 << b)) continue;
            r_addsub(v[j], v[j | (1 << b)]);
        }
    }
}
template <int S, typename T2>
__device__ __forceinline__ void x_to_smem(const T2 (&v)[XR], T2 *sm, int tid) {
#pragma unroll
    for (int j = 0; j < XR; ++j) sm[xswz(xl<S>(tid, j))] = v[j];
}
template <int S, typename T2>
__device__ __forceinline__ void x_from_smem(T2 (&v)[XR], const T2 *sm, int tid) {
#pragma unroll
    for (int j = 0; j < XR; ++j) v[j] = sm[xswz(xl<S>(tid, j))];
}

// Shared scratch of the diagonal evaluation
struct XDiagSmem {
    double F[XT];
    double rows[64];
    double A;
    double SL[68], SH[64];  // SL padded by one entry per 16: stage 0 reads SL[(tid & 3) << 4 | j], a 4-way bank conflict otherwise
};
__host__ __device__ __forceinline__ int sl_slot(int x) { return x + (x >> 4); }

// E(l) tables for the tile with outside bits `full` (index of the tile's element l = 0): A, F_t -> SL / SH.
__device__ __forceinline__ void diag_prepare(XDiagSmem &d, const XDiag &D, const XGroup &G, int n, unsigned long long full,
                                             int tid) {
    if (tid < XT) {  // F_t = lin[p_t] + sum_{b outside} quad[p_t][b] z_b
        const int p = G.tpos[tid];
        double f = D.lin[p];
        for (int b = 0; b < n; ++b)
            if (!((G.tmask >> b) & 1ULL)) f += ((full >> b) & 1ULL) ? -D.quad[p * n + b] : D.quad[p * n + b];
        d.F[tid] = f;
    } else if (tid >= 32 && tid < 32 + n) {  // rows of A over the outside positions (n <= 64 - 12)
        const int b = tid - 32;
        double row = 0.0;
        if (!((G.tmask >> b) & 1ULL)) {
            row = D.lin[b];
            for (int b2 = b + 1; b2 < n; ++b2)
                if (!((G.tmask >> b2) & 1ULL)) row += ((full >> b2) & 1ULL) ? -D.quad[b * n + b2] : D.quad[b * n + b2];
            if ((full >> b) & 1ULL) row = -row;
        }
        d.rows[b] = row;
    }
    __syncthreads();
    if (tid == 0) {
        double A = 0.0;
        for (int b = 0; b < n; ++b) A += d.rows[b];  // fixed order
        d.A = A;
    }
    if (tid >= 64 && tid < 192) {  // SL[x] = sum_{t < 6} z_t(x) F_t,  SH[x] = sum_{t >= 6} z_t(x) F_t
        const int x = (tid - 64) & 63, hi = (tid - 64) >> 6;
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) s += ((x >> t) & 1) ? -d.F[6 * hi + t] : d.F[6 * hi + t];
        if (hi) d.SH[x] = s;
        else d.SL[sl_slot(x)] = s;
    }
    __syncthreads();
}

// Residual terms (weight >= 3): coefficients by tile-local pattern, then a 12-bit real Walsh transform in shared memory:
// c[l] <- sum_mu c_mu (-1)^popcount(l & mu).  Deterministic (one thread per pattern, terms in plan order).
__device__ __forceinline__ void diag_residual(double *c, const XDiag &D, unsigned long long full, int tid) {
    for (int i = tid; i < XN; i += XTH) c[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < D.n_mu; i += XTH) {
        double s = 0.0;
        for (int k = D.mu_off[i]; k < D.mu_off[i + 1]; ++k) s += (__popcll(full & D.omask[k]) & 1) ? -D.w[k] : D.w[k];
        c[xswz(D.mu[i])] = s;
    }
    __syncthreads();
    double v[XR];
#pragma unroll
    for (int j = 0; j < XR; ++j) v[j] = c[xswz(xl<2>(tid, j))];
#pragma unroll
    for (int st = 2; st >= 0; --st) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int j = 0; j < XR; ++j) {
                if (j & (1 << b)) continue;
                const double u = v[j], w = v[j | (1 << b)];
                v[j] = u + w;
                v[j | (1 << b)] = u - w;
            }
        if (st == 2) {
#pragma unroll
            for (int j = 0; j < XR; ++j) c[xswz(xl<2>(tid, j))] = v[j];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < XR; ++j) v[j] = c[xswz(xl<1>(tid, j))];
        } else if (st == 1) {
#pragma unroll
            for (int j = 0; j < XR; ++j) c[xswz(xl<1>(tid, j))] = v[j];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < XR; ++j) v[j] = c[xswz(xl<0>(tid, j))];
        } else {
#pragma unroll
            for (int j = 0; j < XR; ++j) c[xswz(xl<0>(tid, j))] = v[j];
        }
    }
    __syncthreads();
}

// v[j] *= scale * exp(i theta E(l_j)) for the thread's 16 elements in stage S
// theta in turns per unit of E: the phase of element l is 2 pi theta E(l)
template <int S, typename T>
__device__ __forceinline__ void diag_apply(typename Reg<T>::type (&v)[XR], const XDiagSmem &d, const XDiag &D, const double *resid,
                                           double theta, T scale, int tid) {
    using T2 = typename V2<T>::type;
#pragma unroll
    for (int j = 0; j < XR; ++j) {
        const int l = xl<S>(tid, j);
        double E = (d.A + d.SH[l >> 6]) + d.SL[sl_slot(l & 63)] + __ldg(D.R + l);  // first sum: the same for 4 (S = 1) or all j
        if (resid) E += resid[xswz(l)];
        const T2 f = gen_phase_turns(theta, E, T(0));
        const T2 x = r_get(v[j]);
        const T fx = f.x * scale, fy = f.y * scale;
        v[j] = r_make(x.x * fx - x.y * fy, x.x * fy + x.y * fx);
    }
}

// One pass (step t of the proposal) for every (tile instance, chain).  Chains whose proposal has fewer steps exit.
// Steps: 0 = FIRST on G1; then blocks of 2 K steps, block b = 1..r:  [H on G2..GK]  (b < r ? HOST P : FINAL) on G0
// [H on G2..GK]  HOST Dx on G1   -- the second half only for b < r.
template <typename T>
__global__ void __launch_bounds__(XTH, sizeof(T) == 4 ? 3 : 2) gen_pass_kernel(const GenArgs a, const int t) {
    using R = typename Reg<T>::type;  // register / memory representation of one amplitude (same bytes as V2<T>)
    extern __shared__ __align__(16) unsigned char smraw[];
    R *st = reinterpret_cast<R *>(smraw);  // [XN]
    __shared__ unsigned long long hoff[XN >> XFORCED];  // index offset of the local bits 4..11
    __shared__ GenChain ch_s;
    __shared__ XDiagSmem dg;
    __shared__ XGroup G;
    const int tid = threadIdx.x;
    const int n = a.n, K = a.K;
    const int64_t c = blockIdx.y;
    if (tid == 0) ch_s = a.chains[c];
    __syncthreads();
    const int r = ch_s.r;
    int role, gi;
    bool host_x = false;
    if (t == 0) {
        role = X_FIRST;
        gi = 1;
        host_x = true;
    } else {
        const int b = 1 + (t - 1) / (2 * K), pos = (t - 1) % (2 * K);
        if (b > r || (b == r && pos >= K)) return;
        if (pos < K - 1) { role = X_HONLY; gi = 2 + pos; }
        else if (pos == K - 1) { role = b == r ? X_FINAL : X_HOST; gi = 0; }
        else if (pos < 2 * K - 1) { role = X_HONLY; gi = 2 + pos - K; }
        else { role = X_HOST; gi = 1; host_x = true; }
    }
    if (tid < (int)(sizeof(XGroup) / 4)) reinterpret_cast<int *>(&G)[tid] = reinterpret_cast<const int *>(&a.groups[gi])[tid];
    __syncthreads();
    // tile instance: the n - 12 bits of blockIdx.x go to the positions outside the tile
    unsigned long long base = 0ULL;
    {
        unsigned long long b = blockIdx.x;
        for (int q = 0; q < n; ++q)
            if (!((G.tmask >> q) & 1ULL)) {
                base |= (b & 1ULL) << q;
                b >>= 1;
            }
    }
    {   // local bits 4..11 -> index offset (tid = the 8-bit pattern)
        unsigned long long o = 0ULL;
#pragma unroll
        for (int k = XFORCED; k < XT; ++k)
            if ((tid >> (k - XFORCED)) & 1) o |= 1ULL << G.tpos[k];
        hoff[tid] = o;
    }
    __syncthreads();
    R *state = reinterpret_cast<R *>(a.state) + ((size_t)c << n) + base;
    const int m2 = (G.active >> 8) & 15, m1 = (G.active >> 4) & 15, m0 = G.active & 15;
    const bool use1 = (m1 | m0) != 0, use0 = m0 != 0;
    const T hscale = (T)exp2(-0.5 * (double)G.n_active);  // normalisation of one Walsh-Hadamard transform over the group
    R v[XR];
    // stage-2 element (tid, j): local index tid | j << 8 -> offset (tid & 15) + hoff[(tid >> 4) | j << 4]
    if (role == X_FIRST) {  // <x| H^(x)n |s> = 2^(-n/2) (-1)^popcount(s & x)
        const T mag = (T)exp2(-0.5 * (double)n);
#pragma unroll
        for (int j = 0; j < XR; ++j) {
            const unsigned long long x = base | hoff[(tid >> 4) | (j << 4)] | (unsigned long long)(tid & 15);
            v[j] = r_make((__popcll(x & ch_s.cur) & 1) ? -mag : mag, T(0));
        }
    } else {
#pragma unroll
        for (int j = 0; j < XR; ++j) v[j] = state[hoff[(tid >> 4) | (j << 4)] | (unsigned long long)(tid & 15)];
    }
    const bool host = role == X_FIRST || role == X_HOST;
    const bool pre = role != X_FIRST;  // FIRST starts in the basis of its diagonal: no transform before it
    const XDiag &D = host_x ? a.diag_x[ch_s.grp] : *a.diag_p;
    double theta = 0.0;
    double *resid = nullptr;
    if (host) {
        theta = host_x ? -ch_s.gamma * a.dt : (a.phase_active ? (1.0 - ch_s.gamma) * a.alpha * a.dt : 0.0);
        theta *= 0.15915494309189533577;  // in turns: diag_apply forms exp(2 pi i theta E)
        diag_prepare(dg, D, G, n, base, tid);
        if (D.n_mu > 0) {
            resid = reinterpret_cast<double *>(st + XN);
            diag_residual(resid, D, base, tid);
        }
    }
    const T dscale = pre ? hscale : T(1);  // the transform before the diagonal is normalised inside the diagonal factor
    if (pre) wht4(v, m2);
    if (!use1) {
        if (host) {
            diag_apply<2, T>(v, dg, D, resid, theta, dscale, tid);
            wht4(v, m2);
        }
    } else {
        x_to_smem<2>(v, st, tid);
        __syncthreads();
        x_from_smem<1>(v, st, tid);
        if (pre) wht4(v, m1);
        if (!use0) {
            if (host) {
                diag_apply<1, T>(v, dg, D, resid, theta, dscale, tid);
                wht4(v, m1);
            }
        } else {
            x_to_smem<1>(v, st, tid);  // each thread rewrites exactly the slots it read
            __syncthreads();
            x_from_smem<0>(v, st, tid);
            if (pre) wht4(v, m0);
            if (host) {
                diag_apply<0, T>(v, dg, D, resid, theta, dscale, tid);
                wht4(v, m0);
            }
            if (role == X_FINAL) {  // computational basis: 16 consecutive amplitudes per thread, 64 per segment = 4 lanes
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < XR; ++j) {
                    v[j] = r_scale(v[j], hscale);
                    const typename V2<T>::type o = r_get(v[j]);
                    s += (double)o.x * (double)o.x + (double)o.y * (double)o.y;
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if ((tid & 3) == 0) a.segsum[c * ((int64_t)1 << (n - 6)) + (int64_t)(base >> 6) + (tid >> 2)] = s;
                x_to_smem<0>(v, st, tid);
                __syncthreads();
#pragma unroll 4
                for (int j = tid; j < XN; j += XTH) state[j] = st[xswz(j)];  // G0's tile is 4096 consecutive amplitudes
                return;
            }
            x_to_smem<0>(v, st, tid);
            __syncthreads();
            x_from_smem<1>(v, st, tid);
            if (host) wht4(v, m1);
        }
        x_to_smem<1>(v, st, tid);
        __syncthreads();
        x_from_smem<2>(v, st, tid);
        if (host) wht4(v, m2);
    }
    // normalisation of the transform after the diagonal (host pass) or of the only one (H pass)
#pragma unroll
    for (int j = 0; j < XR; ++j) {
        state[hoff[(tid >> 4) | (j << 4)] | (unsigned long long)(tid & 15)] = r_scale(v[j], hscale);
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

// Static-analysis hook (tests/test_smem_layouts.py, CPU): local index and swizzled shared-memory slot of (tid, j) in
// register stage 0 / 1 / 2 of the general path's tile.
extern "C" int qmc_debug_general_slot(int stage, int tid, int j, int *local_out) {
    if (tid < 0 || tid >= XTH || j < 0 || j >= XR || stage < 0 || stage > 2) return -1;
    const int l = stage == 2 ? xl<2>(tid, j) : (stage == 1 ? xl<1>(tid, j) : xl<0>(tid, j));
    if (local_out) *local_out = l;
    return xswz(l);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct DiagDev {  // device allocations behind one XDiag
    double *lin = nullptr, *quad = nullptr, *R = nullptr, *w = nullptr;
    int *mu_off = nullptr;
    unsigned short *mu = nullptr;
    unsigned long long *omask = nullptr;
};

struct GeneralWorkspace {
    int mixer_epoch = -1;
    bool geometry_ready = false;
    int K = 0;
    XGroup groups[XMAXG];
    std::vector<DiagDev> dev;      // [0] = P, [1 + g] = Dx of mixer group g
    std::vector<XDiag> host_diag;  // same order
    XDiag *d_diags = nullptr;      // device copy: [1 + n_groups]
    bool any_residual = false;
    // per call
    int64_t cap_chains = 0;
    size_t cap_esz = 0;
    void *state = nullptr;
    double *segsum = nullptr;
    GenChain *chains = nullptr;
};

static void free_diag(DiagDev &d) {
    cudaFree(d.lin); cudaFree(d.quad); cudaFree(d.R); cudaFree(d.w); cudaFree(d.mu_off); cudaFree(d.mu); cudaFree(d.omask);
    d = DiagDev();
}

void qmc_general_free(qmc_model *m) {
    GeneralWorkspace *w = m->general;
    if (!w) return;
    for (auto &d : w->dev) free_diag(d);
    cudaFree(w->d_diags);
    cudaFree(w->state); cudaFree(w->segsum); cudaFree(w->chains);
    delete w;
    m->general = nullptr;
}

// Qubit groups: G0 = positions 0..11; then groups of 8 from the top, the last one takes what is left (1..8 positions).
// Tile of an HI group = positions 0..3, its qubits, and the lowest free positions to make 12.
static int plan_geometry(int n, GeneralWorkspace *w) {
    std::vector<std::vector<int>> qs;
    std::vector<int> g0;
    for (int q = 0; q < XT; ++q) g0.push_back(q);
    qs.push_back(g0);
    for (int top = n; top > XT;) {
        const int lo = std::max(XT, top - 8);
        std::vector<int> g;
        for (int q = lo; q < top; ++q) g.push_back(q);
        qs.push_back(g);
        top = lo;
    }
    QMC_REQUIRE((int)qs.size() <= XMAXG, QMC_ERR_UNSUPPORTED, "general path: n=%d needs %d qubit groups (max %d)", n, (int)qs.size(), XMAXG);
    w->K = (int)qs.size() - 1;
    for (size_t gi = 0; gi < qs.size(); ++gi) {
        XGroup &G = w->groups[gi];
        memset(&G, 0, sizeof(G));
        unsigned long long tm = 0ULL;
        for (int q : qs[gi]) tm |= 1ULL << q;
        for (int q = 0; q < XFORCED; ++q) tm |= 1ULL << q;
        for (int q = 0; q < n && __builtin_popcountll(tm) < XT; ++q) tm |= 1ULL << q;
        int k = 0;
        unsigned short act = 0;
        for (int q = 0; q < n; ++q)
            if ((tm >> q) & 1ULL) {
                if (std::find(qs[gi].begin(), qs[gi].end(), q) != qs[gi].end()) act |= (unsigned short)(1u << k);
                G.tpos[k++] = (unsigned char)q;
            }
        G.tmask = tm;
        G.active = act;
        G.n_active = (int)qs[gi].size();
    }
    return QMC_OK;
}

// Device form of the diagonal with terms (mask, weight) hosted on tile geometry G.
static int build_diag(int n, const XGroup &G, const std::vector<uint64_t> &masks, const std::vector<double> &weights,
                      DiagDev &dd, XDiag &out, bool &any_residual) {
    std::vector<double> lin(n, 0.0), quad((size_t)n * n, 0.0);
    std::map<unsigned short, std::vector<std::pair<unsigned long long, double>>> resid;
    for (size_t k = 0; k < masks.size(); ++k) {
        const int pc = __builtin_popcountll(masks[k]);
        if (pc == 1) {
            lin[__builtin_ctzll(masks[k])] += weights[k];
        } else if (pc == 2) {
            const int a = __builtin_ctzll(masks[k]), b = 63 - __builtin_clzll(masks[k]);
            quad[(size_t)a * n + b] += weights[k];
            quad[(size_t)b * n + a] += weights[k];
        } else {
            unsigned short lm = 0;
            for (int t = 0; t < XT; ++t)
                if ((masks[k] >> G.tpos[t]) & 1ULL) lm |= (unsigned short)(1u << t);
            resid[lm].push_back({masks[k] & ~G.tmask, weights[k]});
        }
    }
    std::vector<double> R(XN);
    for (int l = 0; l < XN; ++l) {
        double r = 0.0;
        for (int t = 0; t < XT; ++t)
            for (int t2 = t + 1; t2 < XT; ++t2) {
                const double z = (((l >> t) ^ (l >> t2)) & 1) ? -1.0 : 1.0;
                r += quad[(size_t)G.tpos[t] * n + G.tpos[t2]] * z;
            }
        R[l] = r;
    }
    std::vector<int> mu_off(1, 0);
    std::vector<unsigned short> mu;
    std::vector<unsigned long long> om;
    std::vector<double> wv;
    for (auto &kv : resid) {
        mu.push_back(kv.first);
        for (auto &tw : kv.second) {
            om.push_back(tw.first);
            wv.push_back(tw.second);
        }
        mu_off.push_back((int)om.size());
    }
    auto up = [](void **p, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 8));
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*p, src, bytes, cudaMemcpyHostToDevice);
        return e;
    };
    QMC_CUDA_TRY(up((void **)&dd.lin, lin.data(), sizeof(double) * lin.size()));
    QMC_CUDA_TRY(up((void **)&dd.quad, quad.data(), sizeof(double) * quad.size()));
    QMC_CUDA_TRY(up((void **)&dd.R, R.data(), sizeof(double) * R.size()));
    QMC_CUDA_TRY(up((void **)&dd.mu_off, mu_off.data(), sizeof(int) * mu_off.size()));
    QMC_CUDA_TRY(up((void **)&dd.mu, mu.data(), sizeof(unsigned short) * mu.size()));
    QMC_CUDA_TRY(up((void **)&dd.omask, om.data(), sizeof(unsigned long long) * om.size()));
    QMC_CUDA_TRY(up((void **)&dd.w, wv.data(), sizeof(double) * wv.size()));
    out.lin = dd.lin; out.quad = dd.quad; out.R = dd.R; out.n_mu = (int)mu.size(); out.mu_off = dd.mu_off; out.mu = dd.mu;
    out.omask = dd.omask; out.w = dd.w;
    if (!mu.empty()) any_residual = true;
    return QMC_OK;
}

static int general_prepare(qmc_model *m) {
    if (!m->general) m->general = new GeneralWorkspace();
    GeneralWorkspace *w = m->general;
    const int n = m->n;
    if (!w->geometry_ready) {
        int rc = plan_geometry(n, w);
        if (rc != QMC_OK) return rc;
        w->geometry_ready = true;
    }
    if (w->mixer_epoch != m->mixer_epoch) {
        for (auto &d : w->dev) free_diag(d);
        w->dev.assign(1 + m->n_groups, DiagDev());
        w->host_diag.assign(1 + m->n_groups, XDiag());
        w->any_residual = false;
        // P: D(idx) = sum_{j<k} cJ[j][k] z_j z_k + sum_j ch[j] z_j over spins j (fn_qckt_problem_half :54-91); spin j = qubit n-1-j
        std::vector<uint64_t> pm;
        std::vector<double> pw;
        for (int j = 0; j < n; ++j) {
            pm.push_back(1ULL << (n - 1 - j));
            pw.push_back(m->host_ch[j]);
            for (int k = j + 1; k < n; ++k) {
                pm.push_back((1ULL << (n - 1 - j)) | (1ULL << (n - 1 - k)));
                pw.push_back(m->host_cJ[(size_t)j * n + k]);
            }
        }
        int rc = build_diag(n, w->groups[0], pm, pw, w->dev[0], w->host_diag[0], w->any_residual);
        if (rc != QMC_OK) return rc;
        for (int g = 0; g < m->n_groups; ++g) {
            std::vector<uint64_t> xm(m->masks.begin() + m->group_off[g], m->masks.begin() + m->group_off[g + 1]);
            std::vector<double> xw(m->weights.begin() + m->group_off[g], m->weights.begin() + m->group_off[g + 1]);
            rc = build_diag(n, w->groups[1], xm, xw, w->dev[1 + g], w->host_diag[1 + g], w->any_residual);
            if (rc != QMC_OK) return rc;
        }
        cudaFree(w->d_diags);
        w->d_diags = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&w->d_diags, sizeof(XDiag) * w->host_diag.size()));
        QMC_CUDA_TRY(cudaMemcpy(w->d_diags, w->host_diag.data(), sizeof(XDiag) * w->host_diag.size(), cudaMemcpyHostToDevice));
        w->mixer_epoch = m->mixer_epoch;
    }
    return QMC_OK;
}

int qmc_general_num_passes(qmc_model *m, int group) {  // test / benchmark hook: passes per Trotter layer
    (void)group;
    if (!m || !m->general || !m->general->geometry_ready) return -1;
    return 2 * m->general->K;
}

template <typename T>
static int general_run_t(qmc_model *m, GeneralWorkspace *w, GenArgs a, int64_t n_chains, int64_t n_hops, uint32_t hop0,
                         int r_cap, cudaStream_t st) {
    using T2 = typename V2<T>::type;
    const int n = m->n;
    const int64_t N = 1LL << n;
    const size_t smem = sizeof(T2) * (size_t)XN + (w->any_residual ? sizeof(double) * (size_t)XN : 0);
    QMC_CUDA_TRY(cudaFuncSetAttribute(gen_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device: set every call
    int64_t cap;
    if (w->cap_chains >= n_chains && sizeof(T2) == w->cap_esz) {
        cap = w->cap_chains;  // steady state: no allocator / memory query
    } else {
        size_t free_b = 0, total_b = 0;
        QMC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_chain = sizeof(T2) * (size_t)N + sizeof(double) * (size_t)(N >> 6) + sizeof(GenChain);
        const size_t have = (size_t)w->cap_chains * (w->cap_esz * (size_t)N + sizeof(double) * (size_t)(N >> 6) + sizeof(GenChain));
        cap = (int64_t)(((double)(free_b + have) * 0.85) / (double)per_chain);
        if (cap > 65535) cap = 65535;
        QMC_REQUIRE(cap >= 1, QMC_ERR_CUDA, "not enough device memory for one 2^%d statevector", n);
    }
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
    const int K = w->K;
    const int max_steps = 1 + 2 * K * (r_cap - 1) + K;  // FIRST, r - 1 full blocks, the first half of block r
    const int64_t tiles = N >> XT;
    const int64_t hops = a.mode == QMC_MODE_CHAIN ? n_hops : 1;
    for (int64_t c0 = 0; c0 < n_chains; c0 += batch) {
        const int64_t nc = std::min<int64_t>(batch, n_chains - c0);
        a.chain0 = c0;
        for (int64_t hop = 0; hop < hops; ++hop) {
            QMC_NVTX("general hop: prepare, Hadamard-basis passes, sample+accept");
            a.hop = hop0 + (uint32_t)hop;
            a.hop_index = hop;
            gen_prepare_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(a, nc, hop == 0);
            for (int t = 0; t < max_steps; ++t)
                gen_pass_kernel<T><<<dim3((unsigned)tiles, (unsigned)nc), XTH, smem, st>>>(a, t);
            m->launches += 1 + max_steps;
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
    QMC_REQUIRE(n > XT && n <= QMC_GENERAL_MAX_SPINS, QMC_ERR_UNSUPPORTED,
                "general statevector path (k-body / custom mixers, complex128) covers %d <= n <= %d (n=%d)", XT + 1,
                QMC_GENERAL_MAX_SPINS, n);
    QMC_REQUIRE(m->n_groups >= 1, QMC_ERR_BADARG, "no mixer set (qmc_mixer_set)");
    if (n_chains == 0) return QMC_OK;
    int rc = general_prepare(m);
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
    a.n = n; a.n_groups = m->n_groups; a.phase_active = m->phase_active; a.mode = mode; a.K = w->K;
    a.alpha = m->alpha; a.dt = dt; a.temperature = temperature; a.g_lo = g_lo; a.g_hi = g_hi;
    for (int g = 0; g <= w->K; ++g) a.groups[g] = w->groups[g];
    a.diag_p = w->d_diags;
    a.diag_x = w->d_diags + 1;
    a.J = m->dJ; a.h = m->dh; a.group_cum = m->d_group_cum;
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
