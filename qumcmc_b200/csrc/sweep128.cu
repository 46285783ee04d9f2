// sweep128.cu -- complex128 statevectors in HBM, 13 <= n <= 20, one-body X mixers: the reference's own precision
// (qulacs evolves complex128: qumcmc/quantum_mcmc_routines.py:150-152) at sweep-kernel cost instead of the general
// tiled path.  Same algorithm as sweep.cu -- one proposal U = M (P M)^{r-1} is r sweeps, sweep k applies
//     LO: M_lo(k) P M_lo(k+1)      HI: M_hi(k) P M_hi(k+1)
// alternating between the two qubit groups, first sweep write-only from the analytic product state M|s>, last sweep (always
// LO) read-only with |a|^2 segment sums -- re-dimensioned for 16-byte amplitudes:
//   * tile = 2^12 amplitudes (64 KiB), 128 threads x 32 amplitudes (5 register qubits, 128 data registers);
//   * LO group = qubits 0..9 (tile = 4096 contiguous amplitudes), HI group = qubits 10..n-1 (tile = index bits
//     [0, LC) u [10, n), LC = 22 - n): 10 butterfly rounds per half-layer at n = 20 on either side, two register
//     layouts per tile, one shared-memory transposition each way (128-bit accesses, XOR-swizzled, conflict-free);
//   * storage basis S(idx) = (-i)^popcount(idx) a(idx): every exp(-i theta X_q) is a real rotation, tan form
//     S_lo' = S_lo + t S_hi, S_hi' = S_hi - t S_lo (4 DFMA per pair), (prod cos)^r folded into the initial state;
//   * phase layer exp(+i c D(idx)), D = A + sum_b z_b F_b + R(j) per thread: exp(i c A), exp(i c F_b) from six fp64 sincos
//     per thread and tile, exp(i c R(j)) from a 32-entry table per CTA, the 32 factors of a thread as a running product
//     along the Gray-code order (no per-amplitude sincos);
//   * K3/K4 as in sweep.cu with fp64 segment sums (32-amplitude segments).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "group_sched.cuh"

namespace {

constexpr int TB = 12;            // tile bits
constexpr int TILE = 1 << TB;
constexpr int NT = 128;           // threads per CTA
constexpr int RB = 5;             // register qubits per layout
constexpr int APT = 1 << RB;      // amplitudes per thread
constexpr int LOQ = 10;           // qubits of the LO group
constexpr int NMAX = 20;          // largest n of this path
constexpr int SMEM_BYTES = TILE * 16;

struct ChainParams128 {
    unsigned long long cur;
    double e_cur, u_meas, u_acc;
    long long n_acc;
    int r, group, active, pad;
    double cphase;      // (1 - gamma) alpha dt, 0 when the phase layer is skipped
    double cphase_x;    // XQ mode: -gamma dt, the rate of the mixer diagonal exp(-i gamma dt E_x)
    double scale;       // (prod_q cos theta_q)^r, folded into the initial product state
    double tq[NMAX];    // tan(theta_q)
};

// ---------------------------------------------------------------------------------------------
// tile geometries.  A layout assigns the 12 tile-local index bits to 5 register bits, 5 lane bits and 2 warp bits;
// L1 keeps the lanes on local bits 0..4 (HBM accesses in the longest contiguous runs the tile has), L2 is the layout of the
// phase layer.  gpos maps a local bit to its index bit.
// ---------------------------------------------------------------------------------------------
struct LoGeo {  // 4096 contiguous amplitudes; qubits 0..9 mixed here, local bits 10, 11 ride along
    static constexpr bool two = true;
    __host__ __device__ static constexpr int gpos(int i) { return i; }
    __host__ __device__ static constexpr int r1(int b) { return 5 + b; }
    __host__ __device__ static constexpr int l1(int b) { return b; }
    __host__ __device__ static constexpr int w1(int b) { return 10 + b; }
    __host__ __device__ static constexpr int r2(int b) { return b; }
    __host__ __device__ static constexpr int l2(int b) { return 5 + b; }
    __host__ __device__ static constexpr int w2(int b) { return 10 + b; }
    static constexpr int act1 = 31, act2 = 31;
    __host__ __device__ static constexpr int swz(int i) { return i ^ ((i >> 5) & 7); }
    __host__ __device__ static constexpr int64_t tile_base(int64_t tile) { return tile << TB; }
};

template <int LC>
struct HiGeo {  // local [0, LC) <-> index bits [0, LC); local [LC, 12) <-> index bits [10, 22 - LC) = the HI qubits
    static constexpr int W = TB - LC;
    static constexpr bool two = W > RB;
    __host__ __device__ static constexpr int gpos(int i) { return i < LC ? i : LOQ + (i - LC); }
    __host__ __device__ static constexpr int r1(int b) { return 7 + b; }
    __host__ __device__ static constexpr int l1(int b) { return b; }
    __host__ __device__ static constexpr int w1(int b) { return 5 + b; }
    __host__ __device__ static constexpr int r2(int b) { return 2 + b; }
    __host__ __device__ static constexpr int l2(int b) { return b < 2 ? b : 5 + b; }
    __host__ __device__ static constexpr int w2(int b) { return 10 + b; }
    __host__ __device__ static constexpr int mask_from(int first_local) {  // register bits whose local bit is an HI qubit
        int m = 0;
        for (int b = 0; b < RB; ++b)
            if (first_local + b >= LC) m |= 1 << b;
        return m;
    }
    static constexpr int act1 = mask_from(7), act2 = mask_from(2);
    __host__ __device__ static constexpr int swz(int i) { return i ^ (((i >> 7) & 1) << 2); }
    __host__ __device__ static constexpr int64_t tile_base(int64_t tile) { return tile << LC; }
};

// index offset of a tile-local index (sum of its bits moved to their index positions)
template <typename G>
__host__ __device__ constexpr int64_t gmap(int local) {
    int64_t o = 0;
    for (int i = 0; i < TB; ++i)
        if ((local >> i) & 1) o |= (int64_t)1 << G::gpos(i);
    return o;
}
// local index of (thread, register pattern j) in layout 1 / 2
template <typename G>
__host__ __device__ constexpr int local1(int tid, int j) {
    int o = 0;
    for (int b = 0; b < RB; ++b) {
        if ((tid >> b) & 1) o |= 1 << G::l1(b);
        if ((j >> b) & 1) o |= 1 << G::r1(b);
    }
    for (int b = 0; b < 2; ++b)
        if ((tid >> (5 + b)) & 1) o |= 1 << G::w1(b);
    return o;
}
template <typename G>
__host__ __device__ constexpr int local2(int tid, int j) {
    int o = 0;
    for (int b = 0; b < RB; ++b) {
        if ((tid >> b) & 1) o |= 1 << G::l2(b);
        if ((j >> b) & 1) o |= 1 << G::r2(b);
    }
    for (int b = 0; b < 2; ++b)
        if ((tid >> (5 + b)) & 1) o |= 1 << G::w2(b);
    return o;
}

struct Sweep128Args {
    int n;
    double2 *state;            // [chains][2^n]
    ChainParams128 *params;    // [chains]
    const double *af_lo, *af_hi;  // [2^(n-5)][6]: A, F0..F4 of the phase layout of the LO / HI tiles
    double R_lo[APT], R_hi[APT];  // R(j): couplings among the register qubits of the phase layout
    double *segsum;            // [chains][2^(n-5)]
    double gen_scale;          // XQ mode: normalisation of the Walsh-Hadamard transforms of this launch
    const double *xq_R;        // XQ mode, HI sweeps: R(j) of the mixer diagonal, [mixer group][32]
    int64_t xq_af_stride;      // XQ mode: doubles between the af_hi tables of consecutive mixer groups
    const double *xt_ex;       // XT mode: tabulated mixer exponent E_x(x), fp64, [mixer group][2^n] (null: quadratic tables)
    double *probs_out;         // PROBS mode (unnormalised) or null
    double2 *amps_out;
};

// ---------------------------------------------------------------------------------------------
// register-level pieces
// ---------------------------------------------------------------------------------------------
template <int ACTIVE, typename QP>
__device__ __forceinline__ void bfly5(double2 (&v)[APT], const ChainParams128 &cp, QP qpos) {
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        if (!((ACTIVE >> b) & 1)) continue;
        const double t = cp.tq[qpos(b)];
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            if (j & (1 << b)) continue;
            const double2 u = v[j], w = v[j | (1 << b)];
            v[j] = make_double2(fma(t, w.x, u.x), fma(t, w.y, u.y));
            v[j | (1 << b)] = make_double2(fma(-t, u.x, w.x), fma(-t, u.y, w.y));
        }
    }
}
template <int ACTIVE>
__device__ __forceinline__ void bfly5t(double2 (&v)[APT], const double t) {  // one multiplier for every qubit (XQ: +-1)
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        if (!((ACTIVE >> b) & 1)) continue;
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            if (j & (1 << b)) continue;
            const double2 u = v[j], w = v[j | (1 << b)];
            v[j] = make_double2(fma(t, w.x, u.x), fma(t, w.y, u.y));
            v[j | (1 << b)] = make_double2(fma(-t, u.x, w.x), fma(-t, u.y, w.y));
        }
    }
}
__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cis(double x) {
    double s, c;
    sincos(x, &s, &c);
    return make_double2(c, s);
}
__device__ __forceinline__ double2 times_i_pow(double2 z, int k) {
    switch (k & 3) {
        case 0: return z;
        case 1: return make_double2(-z.y, z.x);
        case 2: return make_double2(-z.x, -z.y);
        default: return make_double2(z.y, -z.x);
    }
}

template <typename G> struct Q1 { __device__ __forceinline__ int operator()(int b) const { return G::gpos(G::r1(b)); } };
template <typename G> struct Q2 { __device__ __forceinline__ int operator()(int b) const { return G::gpos(G::r2(b)); } };

// Phase layer on the thread's 32 amplitudes (register layout with qubits qp(0..4)): factor(j) = exp(i c (A + sum_b z_b F_b
// + R(j))), z_b = +1 / -1 for register bit b clear / set.  er[j] = exp(i c R(j)) comes from shared memory; the thread part
// is a running product along the Gray-code order: one complex multiplication per step by exp(-/+ 2 i c F_b).
__device__ __forceinline__ void phase32(double2 (&v)[APT], const double *__restrict__ af, const double2 *er, double c,
                                        double sc = 1.0) {
    double F[RB];
    double a = af[0];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        F[b] = af[1 + b];
        a += F[b];
    }
    double2 f = cis(c * a);  // j = 0: every z_b = +1
    f.x *= sc;               // XQ mode: the running product carries the normalisation of the sweep's transforms
    f.y *= sc;
    double2 g[RB];           // exp(-2 i c F_b): register bit b goes 0 -> 1
#pragma unroll
    for (int b = 0; b < RB; ++b) g[b] = cis(-2.0 * c * F[b]);
#pragma unroll
    for (int i = 0; i < APT; ++i) {
        const int j = i ^ (i >> 1);
        v[j] = cmul(v[j], cmul(f, er[j]));
        if (i < APT - 1) {
            const int i1 = i + 1;
            const int b = (i1 & 1) ? 0 : (i1 & 2) ? 1 : (i1 & 4) ? 2 : (i1 & 8) ? 3 : 4;
            const int jn = i1 ^ (i1 >> 1);
            f = ((jn >> b) & 1) ? cmul(f, g[b]) : cmul(f, make_double2(g[b].x, -g[b].y));
        }
    }
}

// XT mode (any X-string mixer): factor(j) = sc * exp(i c E_x(idx_j)) from the tabulated exponent; ex points at the thread's
// base index, off(j) is the index offset of register j in the phase layout.
template <typename OFF>
__device__ __forceinline__ void phase_lookup32(double2 (&v)[APT], const double *__restrict__ ex, OFF off, double c, double sc) {
    const double ct = c * 0.15915494309189533577;  // turns
#pragma unroll 8
    for (int j = 0; j < APT; ++j) {
        double sn, cs;
        sincospi(2.0 * ct * __ldg(ex + off(j)), &sn, &cs);
        v[j] = cmul(v[j], make_double2(cs * sc, sn * sc));
    }
}

// M|s0> restricted to the thread's 32 amplitudes, in the storage basis (see sweep.cu: init_product):
//   S(idx) = scale^r (-i)^popcount(s0) (-1)^popcount(idx & ~s0) prod_{q in idx ^ s0} t_q
template <typename QP>
__device__ __forceinline__ void init_product32(double2 (&v)[APT], uint64_t base_idx, int n, const ChainParams128 &cp, QP qpos) {
    uint64_t regmask = 0;
#pragma unroll
    for (int b = 0; b < RB; ++b) regmask |= 1ULL << qpos(b);
    const uint64_t diff = (base_idx ^ cp.cur) & ~regmask;
    double mag0 = (__popcll(base_idx & ~cp.cur & ~regmask) & 1) ? -cp.scale : cp.scale;
    for (int q = 0; q < n; ++q)
        if ((diff >> q) & 1ULL) mag0 *= cp.tq[q];
    const double2 g = times_i_pow(make_double2(1.0, 0.0), -(int)__popcll(cp.cur));
    double f[RB];
    int sbit[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        f[b] = cp.tq[qpos(b)];
        sbit[b] = (int)((cp.cur >> qpos(b)) & 1ULL);
    }
#pragma unroll
    for (int j = 0; j < APT; ++j) {
        double mag = mag0;
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int jb = (j >> b) & 1;
            mag *= (jb ^ sbit[b]) ? f[b] : 1.0;
            mag = (jb & (sbit[b] ^ 1)) ? -mag : mag;
        }
        v[j] = make_double2(mag * g.x, mag * g.y);
    }
}

// XQ mode: the x-basis image of |s0> (see sweep.cu: init_hadamard): mag * (-1)^popcount(idx & ~s0), real.
template <typename QP>
__device__ __forceinline__ void init_hadamard32(double2 (&v)[APT], uint64_t base_idx, const ChainParams128 &cp, QP qpos, double mag) {
    uint64_t regmask = 0;
    int sreg = 0;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        regmask |= 1ULL << qpos(b);
        sreg |= (int)((~cp.cur >> qpos(b)) & 1ULL) << b;
    }
    const int p0 = __popcll(base_idx & ~cp.cur & ~regmask) & 1;
#pragma unroll
    for (int j = 0; j < APT; ++j) v[j] = make_double2(((p0 ^ __popc(j & sreg)) & 1) ? -mag : mag, 0.0);
}

template <typename G>
__device__ __forceinline__ void to_smem1(const double2 (&v)[APT], double2 *sm, int tid) {
    const int base = local1<G>(tid, 0);
#pragma unroll
    for (int j = 0; j < APT; ++j) sm[G::swz(base | local1<G>(0, j))] = v[j];
}
template <typename G>
__device__ __forceinline__ void from_smem1(double2 (&v)[APT], const double2 *sm, int tid) {
    const int base = local1<G>(tid, 0);
#pragma unroll
    for (int j = 0; j < APT; ++j) v[j] = sm[G::swz(base | local1<G>(0, j))];
}
template <typename G>
__device__ __forceinline__ void to_smem2(const double2 (&v)[APT], double2 *sm, int tid) {
    const int base = local2<G>(tid, 0);
#pragma unroll
    for (int j = 0; j < APT; ++j) sm[G::swz(base | local2<G>(0, j))] = v[j];
}
template <typename G>
__device__ __forceinline__ void from_smem2(double2 (&v)[APT], const double2 *sm, int tid) {
    const int base = local2<G>(tid, 0);
#pragma unroll
    for (int j = 0; j < APT; ++j) v[j] = sm[G::swz(base | local2<G>(0, j))];
}
template <typename G>
__device__ __forceinline__ void load1(double2 (&v)[APT], const double2 *__restrict__ p) {
#pragma unroll
    for (int j = 0; j < APT; ++j) v[j] = p[gmap<G>(local1<G>(0, j))];
}
template <typename G>
__device__ __forceinline__ void store1(const double2 (&v)[APT], double2 *__restrict__ p) {
#pragma unroll
    for (int j = 0; j < APT; ++j) p[gmap<G>(local1<G>(0, j))] = v[j];
}

__device__ __forceinline__ void stage_params128(ChainParams128 &dst_s, const ChainParams128 *src_g, int tid) {
    const int *src = reinterpret_cast<const int *>(src_g);
    int *dst = reinterpret_cast<int *>(&dst_s);
    for (int i = tid; i < (int)(sizeof(ChainParams128) / 4); i += NT) dst[i] = src[i];
}

enum { R_FIRST = 0, R_MID = 1, R_FINAL = 2, R_FINAL1 = 3 };

// One sweep over (tile, chain).  G = LoGeo: roles FIRST / MID / FINAL / FINAL1; G = HiGeo<LC>: FIRST / MID.
// XQ (one coherent group of one- and two-qubit X strings, see sweep.cu): the butterflies are Walsh-Hadamard transforms
// (multiplier +1 into the x basis, -1 back), HI sweeps host the mixer diagonal (tables a.af_hi / a.R_hi built from the mixer's
// weights, rate cphase_x), LO sweeps the phase layer; plain amplitudes, normalisation a.gen_scale with the phase factor.
// XT (with XQ, HI sweeps): the mixer diagonal from the tabulated exponent -- a separate instantiation, so that the XQ kernels
// do not carry the per-amplitude sincospi code (as a run-time branch it cost the XQ HI sweep a third of its rate).
template <typename G, int ROLE, bool XQ = false, bool XT = false>
__global__ void __launch_bounds__(NT, 3) sweep128_kernel(const Sweep128Args a, const int *__restrict__ order, const int list_off) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double2 *sm = reinterpret_cast<double2 *>(smraw);
    __shared__ ChainParams128 cp_s;
    __shared__ double2 er_s[APT];
    constexpr bool LO = G::gpos(TB - 1) == TB - 1;  // LoGeo maps every local bit to itself
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    const int n = a.n;
    double2 *state = a.state + (chain << n) + G::tile_base(tile);
    const int64_t off1 = gmap<G>(local1<G>(tid, 0));
    double2 v[APT];
    if (ROLE == R_MID || ROLE == R_FINAL) load1<G>(v, state + off1);  // in flight while the parameters are staged
    stage_params128(cp_s, a.params + chain, tid);
    if (ROLE == R_FIRST || ROLE == R_MID) {
        if (tid < APT) {
            const ChainParams128 *cg = a.params + chain;
            if (XQ && !LO) er_s[tid] = XT ? make_double2(1.0, 0.0) : cis(cg->cphase_x * __ldg(a.xq_R + cg->group * APT + tid));  // the chain's mixer group
            else er_s[tid] = cis(cg->cphase * (LO ? a.R_lo[tid] : a.R_hi[tid]));
        }
    }
    __syncthreads();
    const ChainParams128 &cp = cp_s;
    const Q1<G> q1;
    const Q2<G> q2;
    // phase layout = L2 when the tile has two layouts, else L1
    const int64_t base_phase = (int64_t)G::tile_base(tile) + (G::two ? gmap<G>(local2<G>(tid, 0)) : off1);
    const double *af = (LO ? a.af_lo : a.af_hi) + (XQ && !LO ? cp.group * a.xq_af_stride : 0) + (tile * NT + tid) * 6;
    const double t1 = LO ? -1.0 : 1.0, t2 = -t1;  // XQ multipliers before / after the diagonal of the sweep
    const double cdiag = XQ && !LO ? cp.cphase_x : cp.cphase;
    const double sdiag = XQ ? a.gen_scale : 1.0;
    if (ROLE == R_FINAL || ROLE == R_FINAL1) {  // LO only
        if (ROLE == R_FINAL1) {
            init_product32(v, (uint64_t)base_phase, n, cp, q2);
        } else {
            if (XQ) bfly5t<G::act1>(v, t1);
            else bfly5<G::act1>(v, cp, q1);
            to_smem1<G>(v, sm, tid);
            __syncthreads();
            from_smem2<G>(v, sm, tid);
            if (XQ) bfly5t<G::act2>(v, t1);
            else bfly5<G::act2>(v, cp, q2);
        }
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < APT; ++j) s += v[j].x * v[j].x + v[j].y * v[j].y;
        a.segsum[(chain << (n - RB)) + tile * NT + tid] = s;
        if (a.probs_out) {
            double *po = a.probs_out + (chain << n) + base_phase;
#pragma unroll
            for (int j = 0; j < APT; ++j) po[j] = v[j].x * v[j].x + v[j].y * v[j].y;
        }
        if (a.amps_out) {
            double2 *ao = a.amps_out + (chain << n) + base_phase;
            const int kb = __popcll((unsigned long long)base_phase);
#pragma unroll
            for (int j = 0; j < APT; ++j) ao[j] = XQ ? v[j] : times_i_pow(v[j], kb + __builtin_popcount(j));  // a = i^popcount S
        }
        return;
    }
    const double mag0 = XQ ? exp2(-0.5 * (double)n) : 0.0;
    if (G::two) {
        if (ROLE == R_FIRST) {
            if (XQ) init_hadamard32(v, (uint64_t)base_phase, cp, q2, mag0);
            else init_product32(v, (uint64_t)base_phase, n, cp, q2);
        } else {
            if (XQ) bfly5t<G::act1>(v, t1);
            else bfly5<G::act1>(v, cp, q1);
            to_smem1<G>(v, sm, tid);
            __syncthreads();
            from_smem2<G>(v, sm, tid);
            if (XQ) bfly5t<G::act2>(v, t1);
            else bfly5<G::act2>(v, cp, q2);
        }
        if (XT && !LO)
            phase_lookup32(v, a.xt_ex + ((int64_t)cp.group << n) + base_phase, [](int j) { return gmap<G>(local2<G>(0, j)); }, cdiag, sdiag);
        else phase32(v, af, er_s, cdiag, sdiag);
        if (XQ) bfly5t<G::act2>(v, t2);
        else bfly5<G::act2>(v, cp, q2);
        if (ROLE != R_FIRST) __syncthreads();  // every thread has read its layout-2 slots
        to_smem2<G>(v, sm, tid);
        __syncthreads();
        from_smem1<G>(v, sm, tid);
        if (XQ) bfly5t<G::act1>(v, t2);
        else bfly5<G::act1>(v, cp, q1);
    } else {
        if (ROLE == R_FIRST) {
            if (XQ) init_hadamard32(v, (uint64_t)base_phase, cp, q1, mag0);
            else init_product32(v, (uint64_t)base_phase, n, cp, q1);
        } else {
            if (XQ) bfly5t<G::act1>(v, t1);
            else bfly5<G::act1>(v, cp, q1);
        }
        if (XT && !LO)
            phase_lookup32(v, a.xt_ex + ((int64_t)cp.group << n) + base_phase, [](int j) { return gmap<G>(local1<G>(0, j)); }, cdiag, sdiag);
        else phase32(v, af, er_s, cdiag, sdiag);
        if (XQ) bfly5t<G::act1>(v, t2);
        else bfly5<G::act1>(v, cp, q1);
    }
    store1<G>(v, state + off1);
}

// ---------------------------------------------------------------------------------------------
// tables: A, F_0..F_4 of the phase layout, per (tile, thread); fp64, fixed summation order
// ---------------------------------------------------------------------------------------------
template <typename G>
__global__ void af128_kernel(int n, const double *__restrict__ Jq, const double *__restrict__ hq, double *__restrict__ af) {
    const int64_t total = (int64_t)1 << (n - RB);
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = f >> 7;
        const int tid = (int)(f & 127);
        const uint64_t base = (uint64_t)G::tile_base(tile) + (uint64_t)(G::two ? gmap<G>(local2<G>(tid, 0)) : gmap<G>(local1<G>(tid, 0)));
        int p[RB];
        uint64_t regmask = 0;
        for (int b = 0; b < RB; ++b) {
            p[b] = G::gpos(G::two ? G::r2(b) : G::r1(b));
            regmask |= 1ULL << p[b];
        }
        double A = 0.0, F[RB];
        for (int b = 0; b < RB; ++b) F[b] = hq[p[b]];
        for (int q = 0; q < n; ++q) {
            if ((regmask >> q) & 1ULL) continue;
            const double zq = ((base >> q) & 1ULL) ? -1.0 : 1.0;
            A += hq[q] * zq;
            for (int q2 = q + 1; q2 < n; ++q2) {
                if ((regmask >> q2) & 1ULL) continue;
                const double z2 = ((base >> q2) & 1ULL) ? -1.0 : 1.0;
                A += Jq[q * n + q2] * zq * z2;
            }
            for (int b = 0; b < RB; ++b) F[b] += Jq[q * n + p[b]] * zq;
        }
        double *o = af + f * 6;
        o[0] = A;
        for (int b = 0; b < RB; ++b) o[1 + b] = F[b];
    }
}

// ---------------------------------------------------------------------------------------------
// per-hop bookkeeping
// ---------------------------------------------------------------------------------------------
struct Hop128Args {
    int n, mode, n_groups, phase_active;
    int xq;  // XQ mode: cp.r counts half-layers (2 r sweeps per proposal)
    int64_t n_chains, n_hops, hop;
    uint64_t chain_id0, seed;
    uint32_t hop0;
    double temperature, g_lo, g_hi, dt, alpha;
    const double *J, *h, *group_cum, *qweight;
    const uint64_t *s_in;
    const double *gamma_arr;
    const int *r_arr;
    const double *u_arr;
    const int *group_arr;
    uint64_t *proposals;
    uint8_t *accepted;
    uint64_t *final_state;
    int64_t *acc_count, *hamming_hist;
    int *hist;
    unsigned long long *visit;
};

__device__ void fill_params128(ChainParams128 &cp, const Hop128Args &h, double gamma, int r, int grp) {
    cp.r = r < 1 ? 1 : r;
    cp.group = grp;
    cp.cphase_x = 0.0;
    double scale = 1.0;
    for (int q = 0; q < NMAX; ++q) cp.tq[q] = 0.0;
    if (h.xq) {  // see sweep.cu: fill_proposal_params
        const int real_r = cp.r;
        cp.r = 2 * real_r;
        for (int q = 0; q < h.n; ++q) cp.tq[q] = -1.0;  // the sampler repeats the last transform (back on the LO qubits)
        cp.scale = 1.0;
        const double c = (1.0 - gamma) * h.alpha * h.dt;
        cp.active = h.phase_active && real_r > 1 && c != 0.0;
        cp.cphase = cp.active ? c : 0.0;
        cp.cphase_x = -gamma * h.dt;
        return;
    }
    for (int q = 0; q < h.n; ++q) {
        const double wq = h.qweight[grp * h.n + q];
        if (wq == 0.0) continue;
        double s, c;
        sincos(gamma * wq * h.dt, &s, &c);
        cp.tq[q] = s / c;
        scale *= c;
    }
    cp.scale = pow(scale, (double)cp.r);
    const double c = (1.0 - gamma) * h.alpha * h.dt;
    cp.active = h.phase_active && cp.r > 1 && c != 0.0;
    cp.cphase = cp.active ? c : 0.0;
}

__global__ void chains_init128_kernel(const Hop128Args h, ChainParams128 *params) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= h.n_chains) return;
    ChainParams128 &cp = params[c];
    uint64_t cur;
    if (h.mode == QMC_MODE_CHAIN) cur = h.s_in ? h.s_in[c] : qmc_initial_state(h.seed, h.chain_id0 + (uint64_t)c, h.n);
    else cur = h.s_in[c];
    cp.cur = cur;
    cp.e_cur = (h.mode == QMC_MODE_CHAIN) ? qmc_energy_dev(h.n, h.J, h.h, cur) : 0.0;
    cp.n_acc = 0;
    if (h.mode == QMC_MODE_CHAIN && h.hop0 == 0) qmc_visit(h.visit, cur);
    for (int i = 0; i <= h.n; ++i) h.hist[c * (h.n + 1) + i] = 0;
}

__global__ void hop_prepare128_kernel(const Hop128Args h, ChainParams128 *params) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= h.n_chains) return;
    ChainParams128 &cp = params[c];
    if (h.mode == QMC_MODE_CHAIN) {
        const HopDraws d = qmc_hop_draws(h.seed, h.chain_id0 + (uint64_t)c, h.hop0 + (uint32_t)h.hop);
        const double gamma = h.g_lo + (h.g_hi - h.g_lo) * d.u_gamma;
        cp.u_meas = d.u_meas;
        cp.u_acc = d.u_acc;
        fill_params128(cp, h, gamma, qmc_trotter_steps(d.t, h.dt), qmc_pick_group(h.n_groups, h.group_cum, d.u_group));
    } else {
        cp.u_meas = h.u_arr ? h.u_arr[c] : 0.0;
        cp.u_acc = 0.0;
        fill_params128(cp, h, h.gamma_arr[c], h.r_arr[c], h.group_arr ? h.group_arr[c] : 0);
    }
}

// order[] = chains with odd r sorted by r descending, then chains with even r sorted by r descending
constexpr int SCHED_RMAX = 2047;
__global__ void __launch_bounds__(1024) schedule128_kernel(const ChainParams128 *__restrict__ params, int64_t n_chains, int r_cap,
                                                           int *__restrict__ order) {
    __shared__ int hist[SCHED_RMAX + 1];
    __shared__ int start[SCHED_RMAX + 1];
    for (int i = threadIdx.x; i <= r_cap; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n_chains; c += blockDim.x) atomicAdd(&hist[params[c].r], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int pos = 0;
        for (int parity = 1; parity >= 0; --parity)
            for (int r = r_cap; r >= 1; --r)
                if ((r & 1) == parity) {
                    start[r] = pos;
                    pos += hist[r];
                }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= r_cap; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n_chains; c += blockDim.x) {
        const int r = params[c].r;
        order[start[r] + atomicAdd(&hist[r], 1)] = (int)c;
    }
}

// K3 + K4, one CTA per chain: inverse CDF over the fp64 segment sums (sequential semantics, index order), recompute of
// the selected tile, index inside the segment, energy, Metropolis.
__global__ void __launch_bounds__(NT) sample_accept128_kernel(const Sweep128Args a, const Hop128Args h) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double2 *sm = reinterpret_cast<double2 *>(smraw);
    __shared__ ChainParams128 cp_s;
    __shared__ double s_incl[NT];
    __shared__ double s_warp[NT / 32];
    __shared__ long long s_seg;
    __shared__ double s_rem;
    __shared__ unsigned long long s_sel;
    const int tid = threadIdx.x;
    const int64_t chain = blockIdx.x;
    const int n = a.n;
    stage_params128(cp_s, a.params + chain, tid);
    __syncthreads();
    const ChainParams128 &cp = cp_s;
    const int64_t nseg = 1LL << (n - RB);
    const double *seg = a.segsum + (chain << (n - RB));
    {
        const int64_t per = (nseg + NT - 1) / NT;
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
        const double target = cp.u_meas * s_incl[NT - 1];
        const double lo = tid == 0 ? 0.0 : s_incl[tid - 1];
        if (tid <= last_owner && target >= lo && (target < incl || tid == last_owner)) {
            double c = lo;
            int64_t sel = -1;
            double before = 0.0;
            for (int64_t i = i0; i < i1; ++i) {
                const double x = seg[i];
                if (target < c + x) {
                    sel = i;
                    before = c;
                    break;
                }
                c += x;
            }
            if (sel < 0) {
                sel = i1 - 1;
                before = c - seg[sel];
            }
            s_seg = sel;
            s_rem = target - before;
        }
        __syncthreads();
    }
    const int64_t sg = s_seg;
    const int64_t tile = sg / NT;
    const int owner = (int)(sg % NT);
    double2 v[APT];
    const int64_t base_phase = LoGeo::tile_base(tile) + gmap<LoGeo>(local2<LoGeo>(tid, 0));
    if (cp.r == 1) {
        init_product32(v, (uint64_t)base_phase, n, cp, Q2<LoGeo>());
    } else {
        const double2 *state = a.state + (chain << n) + LoGeo::tile_base(tile);
        load1<LoGeo>(v, state + gmap<LoGeo>(local1<LoGeo>(tid, 0)));
        bfly5<LoGeo::act1>(v, cp, Q1<LoGeo>());
        to_smem1<LoGeo>(v, sm, tid);
        __syncthreads();
        from_smem2<LoGeo>(v, sm, tid);
        bfly5<LoGeo::act2>(v, cp, Q2<LoGeo>());
    }
    if (tid == owner) {
        const double rem = s_rem;
        double c = 0.0;
        int selj = APT - 1;
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            c += v[j].x * v[j].x + v[j].y * v[j].y;
            if (selj == APT - 1 && rem < c) selj = j;
        }
        s_sel = (unsigned long long)base_phase | (unsigned long long)selj;  // layout 2 of the LO tile: regs = index bits 0..4
    }
    __syncthreads();
    if (tid == 0) {
        const uint64_t sp = s_sel;
        ChainParams128 &out = a.params[chain];
        if (h.mode == QMC_MODE_PROPOSE) {
            h.proposals[chain] = sp;
        } else {
            const double e_sp = qmc_energy_dev(n, h.J, h.h, sp);
            const bool acc = qmc_accept_dev(cp.e_cur, e_sp, h.temperature, cp.u_acc);
            if (h.proposals) h.proposals[chain * h.n_hops + h.hop] = sp;
            if (h.accepted) h.accepted[chain * h.n_hops + h.hop] = (uint8_t)acc;
            h.hist[chain * (n + 1) + __popcll(cp.cur ^ sp)] += 1;
            if (acc) {
                out.cur = sp;
                out.e_cur = e_sp;
                out.n_acc = cp.n_acc + 1;
            }
            qmc_visit(h.visit, acc ? sp : cp.cur);
        }
    }
}

__global__ void chains_finish128_kernel(const Hop128Args h, const ChainParams128 *params) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= h.n_chains) return;
    if (h.final_state) h.final_state[c] = params[c].cur;
    if (h.acc_count) h.acc_count[c] = params[c].n_acc;
    if (h.hamming_hist)
        for (int i = 0; i <= h.n; ++i) h.hamming_hist[c * (h.n + 1) + i] = h.hist[c * (h.n + 1) + i];
}

__global__ void probs_normalize128_kernel(int n, const double *segsum, double *probs, double2 *amps) {
    __shared__ double red[256];
    const int64_t nseg = 1LL << (n - RB);
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nseg; i += blockDim.x) s += segsum[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    const double total = red[0];
    const double inv = 1.0 / total, inva = 1.0 / sqrt(total);
    const int64_t N = 1LL << n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        if (probs) probs[i] *= inv;
        if (amps) {
            amps[i].x *= inva;
            amps[i].y *= inva;
        }
    }
}

}  // namespace

// Static-analysis hook (tests/test_smem_layouts.py, CPU) for the complex128 tiles: geometry 0 = LO tile, LC (2..9) = HI tile
// of n = 22 - LC; layout 1 / 2.  out[0] = tile-local index, out[1] = index offset inside the statevector (gmap),
// out[2] = shared-memory slot (128-bit words, swizzled), out[3] = active-register mask of that layout.
template <typename G>
static void debug_layout(int layout, int tid, int j, long long *out) {
    const int local = layout == 1 ? local1<G>(tid, j) : local2<G>(tid, j);
    out[0] = local;
    out[1] = (long long)gmap<G>(local);
    out[2] = G::swz(local);
    out[3] = layout == 1 ? G::act1 : G::act2;
}
extern "C" int qmc_debug_layout128(int geometry, int layout, int tid, int j, long long *out) {
    if (!out || (layout != 1 && layout != 2) || tid < 0 || tid >= NT || j < 0 || j >= APT) return -1;
    switch (geometry) {
        case 0: debug_layout<LoGeo>(layout, tid, j, out); return 0;
        case 2: debug_layout<HiGeo<2>>(layout, tid, j, out); return 0;
        case 3: debug_layout<HiGeo<3>>(layout, tid, j, out); return 0;
        case 4: debug_layout<HiGeo<4>>(layout, tid, j, out); return 0;
        case 5: debug_layout<HiGeo<5>>(layout, tid, j, out); return 0;
        case 6: debug_layout<HiGeo<6>>(layout, tid, j, out); return 0;
        case 7: debug_layout<HiGeo<7>>(layout, tid, j, out); return 0;
        case 8: debug_layout<HiGeo<8>>(layout, tid, j, out); return 0;
        case 9: debug_layout<HiGeo<9>>(layout, tid, j, out); return 0;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Sweep128Workspace {
    int64_t cap_chains = 0;
    double2 *state = nullptr;
    ChainParams128 *params = nullptr;
    double *segsum = nullptr;
    int *hist = nullptr;
    int *order = nullptr;
    double *af_lo = nullptr, *af_hi = nullptr;
    double R_lo[APT], R_hi[APT];
    double *dJq = nullptr, *dhq = nullptr;
    bool tables_ready = false;
    QmcGroupSched gsched;  // L2-resident group schedule: side streams + fork / join events
    // XQ mode: tables of the mixer diagonal in the HI phase layout
    int xq_epoch = -1;
    double *afx_hi = nullptr, *dxJ = nullptr, *dxh = nullptr;
    double *d_Rx = nullptr;  // [mixer groups][32]
    int xq_groups = 0;
    // XT mode: tabulated mixer exponent, fp64, [mixer groups][2^n]
    int xt_epoch = -1, xt_groups = 0;
    double *d_xt = nullptr;
};

void qmc_sweep128_free(qmc_model *m) {
    Sweep128Workspace *w = m->sweep128;
    if (!w) return;
    cudaFree(w->state); cudaFree(w->params); cudaFree(w->segsum); cudaFree(w->hist); cudaFree(w->order);
    cudaFree(w->af_lo); cudaFree(w->af_hi); cudaFree(w->dJq); cudaFree(w->dhq);
    w->gsched.destroy();
    cudaFree(w->afx_hi); cudaFree(w->dxJ); cudaFree(w->dxh); cudaFree(w->d_Rx); cudaFree(w->d_xt);
    delete w;
    m->sweep128 = nullptr;
}

typedef unsigned long long DevMask128;
template <typename K>
static int smem_optin(K kernel, DevMask128 &done) {
    int dev = 0;
    QMC_CUDA_TRY(cudaGetDevice(&dev));
    const DevMask128 bit = 1ULL << (dev & 63);
    if (!(done & bit)) {
        QMC_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        done |= bit;
    }
    return QMC_OK;
}

template <typename G, int ROLE, bool XQ = false, bool XT = false>
static int launch128(const Sweep128Args &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    static DevMask128 done = 0;
    int rc = smem_optin(sweep128_kernel<G, ROLE, XQ, XT>, done);
    if (rc != QMC_OK) return rc;
    sweep128_kernel<G, ROLE, XQ, XT><<<dim3((unsigned)tiles, (unsigned)count), NT, SMEM_BYTES, st>>>(a, order, off);
    return QMC_OK;
}
template <int ROLE, bool XQ = false, bool XT = false>
static int launch128_hi(int n, const Sweep128Args &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    switch (22 - n) {  // LC
        case 2: return launch128<HiGeo<2>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 3: return launch128<HiGeo<3>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 4: return launch128<HiGeo<4>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 5: return launch128<HiGeo<5>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 6: return launch128<HiGeo<6>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 7: return launch128<HiGeo<7>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 8: return launch128<HiGeo<8>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
        case 9: return launch128<HiGeo<9>, ROLE, XQ, XT>(a, order, off, tiles, count, st);
    }
    qmc_set_error("complex128 sweep path: n=%d outside [13, 20]", n);
    return QMC_ERR_UNSUPPORTED;
}

template <typename G>
static void host_R(const std::vector<double> &Jq, int n, double (&R)[APT]) {
    int p[RB];
    for (int b = 0; b < RB; ++b) p[b] = G::gpos(G::two ? G::r2(b) : G::r1(b));
    for (int j = 0; j < APT; ++j) {
        double r = 0.0;
        for (int b = 0; b < RB; ++b)
            for (int b2 = b + 1; b2 < RB; ++b2) {
                const double zb = ((j >> b) & 1) ? -1.0 : 1.0, zb2 = ((j >> b2) & 1) ? -1.0 : 1.0;
                r += Jq[(size_t)std::min(p[b], p[b2]) * n + std::max(p[b], p[b2])] * zb * zb2;
            }
        R[j] = r;
    }
}
template <typename G>
static int build_af128(int n, const double *dJq, const double *dhq, double *af, cudaStream_t st) {
    const int grid = (int)std::min<int64_t>(148 * 8, (((int64_t)1 << (n - RB)) + 127) / 128);
    af128_kernel<G><<<grid, 128, 0, st>>>(n, dJq, dhq, af);
    return QMC_OK;
}

static int ensure_workspace128(qmc_model *m, int64_t chains, cudaStream_t st) {
    if (!m->sweep128) m->sweep128 = new Sweep128Workspace();
    Sweep128Workspace *w = m->sweep128;
    const int n = m->n;
    const int64_t N = 1LL << n;
    if (!w->tables_ready) {
        std::vector<double> Jq((size_t)n * n), hq(n);  // circuit couplings in qubit order: qubit a = spin n-1-a
        for (int a = 0; a < n; ++a) {
            hq[a] = m->host_ch[n - 1 - a];
            for (int b = 0; b < n; ++b) Jq[(size_t)a * n + b] = m->host_cJ[(size_t)(n - 1 - a) * n + (n - 1 - b)];
        }
        QMC_CUDA_TRY(cudaMalloc(&w->dJq, sizeof(double) * (size_t)n * n));
        QMC_CUDA_TRY(cudaMalloc(&w->dhq, sizeof(double) * (size_t)n));
        QMC_CUDA_TRY(cudaMemcpy(w->dJq, Jq.data(), sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
        QMC_CUDA_TRY(cudaMemcpy(w->dhq, hq.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
        const size_t af_bytes = sizeof(double) * 6 * (size_t)(N >> RB);
        QMC_CUDA_TRY(cudaMalloc(&w->af_lo, af_bytes));
        QMC_CUDA_TRY(cudaMalloc(&w->af_hi, af_bytes));
        build_af128<LoGeo>(n, w->dJq, w->dhq, w->af_lo, st);
        host_R<LoGeo>(Jq, n, w->R_lo);
        switch (22 - n) {
            case 2: build_af128<HiGeo<2>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<2>>(Jq, n, w->R_hi); break;
            case 3: build_af128<HiGeo<3>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<3>>(Jq, n, w->R_hi); break;
            case 4: build_af128<HiGeo<4>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<4>>(Jq, n, w->R_hi); break;
            case 5: build_af128<HiGeo<5>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<5>>(Jq, n, w->R_hi); break;
            case 6: build_af128<HiGeo<6>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<6>>(Jq, n, w->R_hi); break;
            case 7: build_af128<HiGeo<7>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<7>>(Jq, n, w->R_hi); break;
            case 8: build_af128<HiGeo<8>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<8>>(Jq, n, w->R_hi); break;
            case 9: build_af128<HiGeo<9>>(n, w->dJq, w->dhq, w->af_hi, st); host_R<HiGeo<9>>(Jq, n, w->R_hi); break;
            default: qmc_set_error("complex128 sweep path: n=%d outside [13, 20]", n); return QMC_ERR_UNSUPPORTED;
        }
        m->launches += 2;
        QMC_CUDA_TRY(cudaGetLastError());
        w->tables_ready = true;
    }
    if (chains > w->cap_chains) {
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(w->state); cudaFree(w->params); cudaFree(w->segsum); cudaFree(w->hist); cudaFree(w->order);
        w->state = nullptr; w->params = nullptr; w->segsum = nullptr; w->hist = nullptr; w->order = nullptr;
        w->cap_chains = 0;
        QMC_CUDA_TRY(cudaMalloc(&w->state, sizeof(double2) * (size_t)N * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->params, sizeof(ChainParams128) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->segsum, sizeof(double) * (size_t)(N >> RB) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->hist, sizeof(int) * (size_t)(n + 1) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->order, sizeof(int) * (size_t)chains));
        w->cap_chains = chains;
    }
    return QMC_OK;
}

template <typename G>
static void build_xq128(int n, const double *dJ, const double *dh, double *af, const std::vector<double> &xJg, double *R, cudaStream_t st) {
    build_af128<G>(n, dJ, dh, af, st);
    double Rg[APT];
    host_R<G>(xJg, n, Rg);
    memcpy(R, Rg, sizeof(Rg));
}

// XT mode: E_x = Walsh-Hadamard transform of the coefficient vector (see sweep.cu: ensure_xt_tables), kept in fp64
__global__ void xt128_wht_kernel(double *__restrict__ x, int64_t half, int bit) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < half; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = ((q >> bit) << (bit + 1)) | (q & (((int64_t)1 << bit) - 1)), j = i | ((int64_t)1 << bit);
        const double u = x[i], v = x[j];
        x[i] = u + v;
        x[j] = u - v;
    }
}
static int ensure_xt_tables128(qmc_model *m, cudaStream_t st) {
    Sweep128Workspace *w = m->sweep128;
    if (w->xt_epoch == m->mixer_epoch) return QMC_OK;
    const int n = m->n, ng = m->n_groups;
    const int64_t N = 1LL << n;
    QMC_CUDA_TRY(cudaStreamSynchronize(st));
    if (w->xt_groups != ng) {
        cudaFree(w->d_xt);
        w->d_xt = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&w->d_xt, sizeof(double) * (size_t)ng * N));
        w->xt_groups = ng;
    }
    std::vector<double> coef((size_t)N);
    for (int g = 0; g < ng; ++g) {
        std::fill(coef.begin(), coef.end(), 0.0);
        for (int k = m->group_off[g]; k < m->group_off[g + 1]; ++k) coef[(size_t)m->masks[k]] += m->weights[k];
        double *x = w->d_xt + (size_t)g * N;
        QMC_CUDA_TRY(cudaMemcpyAsync(x, coef.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
        for (int b = 0; b < n; ++b) xt128_wht_kernel<<<148 * 8, 256, 0, st>>>(x, N >> 1, b);
        m->launches += n;
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
    }
    QMC_CUDA_TRY(cudaGetLastError());
    w->xt_epoch = m->mixer_epoch;
    return QMC_OK;
}

static int ensure_xq_tables128(qmc_model *m, cudaStream_t st) {
    Sweep128Workspace *w = m->sweep128;
    if (w->xq_epoch == m->mixer_epoch) return QMC_OK;
    const int n = m->n, ng = m->n_groups;
    const int64_t N = 1LL << n;
    QMC_REQUIRE(m->quad_mixer && (int)m->xh.size() == ng * n, QMC_ERR_UNSUPPORTED,
                "complex128 sweep path: the mixer groups are not made of one- and two-qubit X strings");
    QMC_CUDA_TRY(cudaStreamSynchronize(st));
    const size_t af_doubles = 6 * (size_t)(N >> RB);
    if (w->xq_groups != ng) {
        cudaFree(w->dxJ); cudaFree(w->dxh); cudaFree(w->afx_hi); cudaFree(w->d_Rx);
        w->dxJ = w->dxh = w->afx_hi = w->d_Rx = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&w->dxJ, sizeof(double) * (size_t)ng * n * n));
        QMC_CUDA_TRY(cudaMalloc(&w->dxh, sizeof(double) * (size_t)ng * n));
        QMC_CUDA_TRY(cudaMalloc(&w->afx_hi, sizeof(double) * af_doubles * ng));
        QMC_CUDA_TRY(cudaMalloc(&w->d_Rx, sizeof(double) * (size_t)ng * APT));
        w->xq_groups = ng;
    }
    QMC_CUDA_TRY(cudaMemcpy(w->dxJ, m->xJ.data(), sizeof(double) * (size_t)ng * n * n, cudaMemcpyHostToDevice));
    QMC_CUDA_TRY(cudaMemcpy(w->dxh, m->xh.data(), sizeof(double) * (size_t)ng * n, cudaMemcpyHostToDevice));
    std::vector<double> Rx((size_t)ng * APT, 0.0);
    for (int g = 0; g < ng; ++g) {
        const double *dJ = w->dxJ + (size_t)g * n * n, *dh = w->dxh + (size_t)g * n;
        double *af = w->afx_hi + g * af_doubles, *R = &Rx[(size_t)g * APT];
        const std::vector<double> xJg(m->xJ.begin() + (size_t)g * n * n, m->xJ.begin() + (size_t)(g + 1) * n * n);
        switch (22 - n) {
            case 2: build_xq128<HiGeo<2>>(n, dJ, dh, af, xJg, R, st); break;
            case 3: build_xq128<HiGeo<3>>(n, dJ, dh, af, xJg, R, st); break;
            case 4: build_xq128<HiGeo<4>>(n, dJ, dh, af, xJg, R, st); break;
            case 5: build_xq128<HiGeo<5>>(n, dJ, dh, af, xJg, R, st); break;
            case 6: build_xq128<HiGeo<6>>(n, dJ, dh, af, xJg, R, st); break;
            case 7: build_xq128<HiGeo<7>>(n, dJ, dh, af, xJg, R, st); break;
            case 8: build_xq128<HiGeo<8>>(n, dJ, dh, af, xJg, R, st); break;
            case 9: build_xq128<HiGeo<9>>(n, dJ, dh, af, xJg, R, st); break;
            default: qmc_set_error("complex128 sweep path: n=%d outside [13, 20]", n); return QMC_ERR_UNSUPPORTED;
        }
        m->launches++;
    }
    QMC_CUDA_TRY(cudaGetLastError());
    QMC_CUDA_TRY(cudaMemcpy(w->d_Rx, Rx.data(), sizeof(double) * Rx.size(), cudaMemcpyHostToDevice));
    w->xq_epoch = m->mixer_epoch;
    return QMC_OK;
}

int qmc_sweep128_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in, uint64_t chain_id0,
                     uint32_t hop0, double temperature, double g_lo, double g_hi, double dt, uint64_t seed,
                     const double *gamma_arr, const int *r_arr, const double *u_arr, const int *group_arr,
                     uint64_t *proposals, uint8_t *accepted, uint64_t *final_state, int64_t *acc_count,
                     int64_t *hamming_hist, void *probs_out, void *amps_out, cudaStream_t st) {
    const int n = m->n;
    QMC_REQUIRE(n >= 13 && n <= NMAX, QMC_ERR_UNSUPPORTED, "complex128 sweep path covers 13 <= n <= %d (n=%d)", NMAX, n);
    const bool xt = !m->all_one_body && !m->quad_mixer;  // XT mode: any other mixer, tabulated mixer exponent
    const bool xq = !m->all_one_body;                    // XQ or XT: 2 r sweeps per proposal
    if (n_chains == 0) return QMC_OK;
    {
        static DevMask128 attr = 0;
        int rca = smem_optin(sample_accept128_kernel, attr);
        if (rca != QMC_OK) return rca;
    }
    const int64_t N = 1LL << n;
    const int64_t tiles = N >> TB;
    Sweep128Workspace *w0 = m->sweep128;
    int64_t cap;
    if (w0 && w0->cap_chains >= n_chains) {
        cap = w0->cap_chains;
    } else {
        size_t free_b = 0, total_b = 0;
        QMC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_chain = sizeof(double2) * (size_t)N + sizeof(double) * (size_t)(N >> RB) + sizeof(ChainParams128) + 256;
        const size_t have = w0 ? (size_t)w0->cap_chains * per_chain : 0;
        cap = (int64_t)(((double)(free_b + have) * 0.85) / (double)per_chain);
        if (cap > 65535) cap = 65535;
        QMC_REQUIRE(cap >= 1, QMC_ERR_CUDA, "not enough device memory for one 2^%d complex128 statevector", n);
    }
    const int64_t batch = std::min<int64_t>(cap, n_chains);
    int rc = ensure_workspace128(m, batch, st);
    if (rc != QMC_OK) return rc;
    Sweep128Workspace *w = m->sweep128;
    if (xq) {
        rc = xt ? ensure_xt_tables128(m, st) : ensure_xq_tables128(m, st);
        if (rc != QMC_OK) return rc;
    }
    int r_max = qmc_trotter_steps(11, dt);
    std::vector<int> r_host;
    if (mode != QMC_MODE_CHAIN) {
        r_host.resize(n_chains);
        QMC_CUDA_TRY(cudaMemcpyAsync(r_host.data(), r_arr, sizeof(int) * n_chains, cudaMemcpyDeviceToHost, st));
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        r_max = 1;
        for (int64_t c = 0; c < n_chains; ++c) r_max = std::max(r_max, r_host[c]);
    }
    if (xq) r_max *= 2;  // XQ mode: 2 r sweeps per proposal (HI first, LO last)
    QMC_REQUIRE(r_max <= SCHED_RMAX, QMC_ERR_UNSUPPORTED, "delta_time=%g gives up to %d Trotter layers (> %d)", dt, r_max, SCHED_RMAX);
    std::vector<int> rhist(r_max + 2, 0);
    const QmcGroupKnobs l2k = qmc_group_knobs((int64_t)sizeof(double2) << n, 16, 3);
    for (int64_t c0 = 0; c0 < n_chains; c0 += batch) {
        const int64_t nc = std::min<int64_t>(batch, n_chains - c0);
        Hop128Args h;
        memset(&h, 0, sizeof(h));
        h.n = n; h.mode = mode; h.n_groups = m->n_groups; h.phase_active = m->phase_active;
        h.xq = xq ? 1 : 0;
        h.n_chains = nc; h.n_hops = n_hops; h.hop = 0;
        h.chain_id0 = chain_id0 + (uint64_t)c0; h.seed = seed; h.hop0 = hop0;
        h.temperature = temperature; h.g_lo = g_lo; h.g_hi = g_hi; h.dt = dt; h.alpha = m->alpha;
        h.J = m->dJ; h.h = m->dh; h.group_cum = m->d_group_cum; h.qweight = m->d_qweight;
        h.s_in = s_in ? s_in + c0 : nullptr;
        h.gamma_arr = gamma_arr ? gamma_arr + c0 : nullptr;
        h.r_arr = r_arr ? r_arr + c0 : nullptr;
        h.u_arr = u_arr ? u_arr + c0 : nullptr;
        h.group_arr = group_arr ? group_arr + c0 : nullptr;
        h.proposals = proposals ? proposals + (mode == QMC_MODE_CHAIN ? c0 * n_hops : c0) : nullptr;
        h.accepted = accepted ? accepted + c0 * n_hops : nullptr;
        h.final_state = final_state ? final_state + c0 : nullptr;
        h.acc_count = acc_count ? acc_count + c0 : nullptr;
        h.hamming_hist = hamming_hist ? hamming_hist + c0 * (n + 1) : nullptr;
        h.hist = w->hist;
        h.visit = mode == QMC_MODE_CHAIN ? m->visit_hist : nullptr;
        Sweep128Args a;
        memset(&a, 0, sizeof(a));
        a.n = n; a.state = w->state; a.params = w->params; a.af_lo = w->af_lo; a.af_hi = w->af_hi;
        memcpy(a.R_lo, w->R_lo, sizeof(a.R_lo));
        memcpy(a.R_hi, w->R_hi, sizeof(a.R_hi));
        a.segsum = w->segsum;
        a.probs_out = mode == QMC_MODE_PROBS ? reinterpret_cast<double *>(probs_out) : nullptr;
        a.amps_out = mode == QMC_MODE_PROBS ? reinterpret_cast<double2 *>(amps_out) : nullptr;
        const int gb = (int)((nc + 127) / 128);
        chains_init128_kernel<<<gb, 128, 0, st>>>(h, w->params);
        m->launches++;
        const int64_t hops = mode == QMC_MODE_CHAIN ? n_hops : 1;
        for (int64_t hop = 0; hop < hops; ++hop) {
            QMC_NVTX("sweep128 hop: prepare, schedule, sweeps, sample+accept");
            h.hop = hop;
            hop_prepare128_kernel<<<gb, 128, 0, st>>>(h, w->params);
            schedule128_kernel<<<1, 1024, 0, st>>>(w->params, nc, r_max, w->order);
            m->launches += 2;
            std::fill(rhist.begin(), rhist.end(), 0);
            for (int64_t c = 0; c < nc; ++c) {  // host replica of the draws: class sizes of the device-side schedule
                int r;
                if (mode == QMC_MODE_CHAIN) {
                    const HopDraws d = qmc_hop_draws(seed, h.chain_id0 + (uint64_t)c, hop0 + (uint32_t)hop);
                    r = qmc_trotter_steps(d.t, dt);
                } else {
                    r = std::max(1, r_host[c0 + c]);
                }
                if (xq) r *= 2;
                rhist[r]++;
            }
            int launched = 0;
            int base[2] = {0, 0};  // start of the odd-r / even-r class in order[]
            for (int r = 1; r <= r_max; r += 2) base[1] += rhist[r];
            // XQ mode: LO sweeps host the phase layer, HI sweeps the mixer diagonal (its tables); every sweep's phase factor
            // carries the normalisation of its own transforms, the HI FIRST sweep also that of the FINAL sweep
            Sweep128Args ax_lo = a, ax_hif = a, ax_him = a;
            if (xq) {
                const int nh = n - LOQ;
                ax_lo.gen_scale = ldexp(1.0, -LOQ);
                ax_hif.af_hi = ax_him.af_hi = w->afx_hi;
                ax_hif.xq_R = ax_him.xq_R = w->d_Rx;
                ax_hif.xq_af_stride = ax_him.xq_af_stride = 6 * (N >> RB);
                if (xt) {
                    ax_hif.xt_ex = ax_him.xt_ex = w->d_xt;
                    ax_hif.af_hi = ax_him.af_hi = w->af_hi;  // unused by the lookup; keeps the address arithmetic in bounds
                    ax_hif.xq_af_stride = ax_him.xq_af_stride = 0;
                }
                ax_him.gen_scale = ldexp(1.0, -nh);
                ax_hif.gen_scale = exp2(-0.5 * nh - 0.5 * LOQ);
            }
            auto lo = [&](int kind, int off, int count, cudaStream_t sg) {
                if (xq) {
                    if (kind == QMC_GS_MID) return launch128<LoGeo, R_MID, true>(ax_lo, w->order, off, tiles, count, sg);
                    if (kind == QMC_GS_FINAL) return launch128<LoGeo, R_FINAL, true>(ax_lo, w->order, off, tiles, count, sg);
                    qmc_set_error("XQ schedule: LO role %d cannot occur (every proposal has an even number of sweeps)", kind);
                    return (int)QMC_ERR_CUDA;
                }
                switch (kind) {
                    case QMC_GS_FIRST: return launch128<LoGeo, R_FIRST>(a, w->order, off, tiles, count, sg);
                    case QMC_GS_MID: return launch128<LoGeo, R_MID>(a, w->order, off, tiles, count, sg);
                    case QMC_GS_FINAL: return launch128<LoGeo, R_FINAL>(a, w->order, off, tiles, count, sg);
                    default: return launch128<LoGeo, R_FINAL1>(a, w->order, off, tiles, count, sg);
                }
            };
            auto hi = [&](bool first, int off, int count, cudaStream_t sg) {
                if (xt)
                    return first ? launch128_hi<R_FIRST, true, true>(n, ax_hif, w->order, off, tiles, count, sg)
                                 : launch128_hi<R_MID, true, true>(n, ax_him, w->order, off, tiles, count, sg);
                if (xq)
                    return first ? launch128_hi<R_FIRST, true>(n, ax_hif, w->order, off, tiles, count, sg)
                                 : launch128_hi<R_MID, true>(n, ax_him, w->order, off, tiles, count, sg);
                return first ? launch128_hi<R_FIRST>(n, a, w->order, off, tiles, count, sg)
                             : launch128_hi<R_MID>(n, a, w->order, off, tiles, count, sg);
            };
            if (l2k.group > 0 && nc > l2k.group) {  // L2-resident group schedule, see group_sched.cuh
                rc = qmc_group_schedule(w->gsched, l2k, st, (int)nc, r_max, rhist, lo, hi, launched);
                if (rc != QMC_OK) return rc;
            } else
            for (int step = 1; step <= r_max; ++step) {
                const int p = step & 1;  // chains with r == step (mod 2) sweep LO at this step
                const int lo_base = p ? base[0] : base[1];
                const int hi_base = p ? base[1] : base[0];
                int lo_gt = 0, hi_gt = 0;
                for (int r = step + 2; r <= r_max; r += 2) lo_gt += rhist[r];
                for (int r = step + 1; r <= r_max; r += 2) hi_gt += rhist[r];
                const int lo_eq = rhist[step];
                if (lo_gt > 0) {
                    rc = lo(step == 1 ? QMC_GS_FIRST : QMC_GS_MID, lo_base, lo_gt, st);
                    if (rc != QMC_OK) return rc;
                    ++launched;
                }
                if (lo_eq > 0) {
                    rc = lo(step == 1 ? QMC_GS_FINAL1 : QMC_GS_FINAL, lo_base + lo_gt, lo_eq, st);
                    if (rc != QMC_OK) return rc;
                    ++launched;
                }
                if (hi_gt > 0) {
                    rc = hi(step == 1, hi_base, hi_gt, st);
                    if (rc != QMC_OK) return rc;
                    ++launched;
                }
            }
            m->launches += launched;
            if (mode == QMC_MODE_PROBS) {
                probs_normalize128_kernel<<<148 * 4, 256, 0, st>>>(n, w->segsum, a.probs_out, a.amps_out);
            } else {
                sample_accept128_kernel<<<(unsigned)nc, NT, SMEM_BYTES, st>>>(a, h);
            }
            m->launches++;
            QMC_CUDA_TRY(cudaGetLastError());
        }
        if (mode == QMC_MODE_CHAIN) {
            chains_finish128_kernel<<<gb, 128, 0, st>>>(h, w->params);
            m->launches++;
        }
        QMC_CUDA_TRY(cudaGetLastError());
    }
    return QMC_OK;
}
