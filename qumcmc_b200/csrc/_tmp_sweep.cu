// sweep.cu -- HBM-resident statevector path for 14 <= n <= 20 (complex64), many chains in flight.
//
// The reference pays one memory pass per *gate* (n(n-1)/2 + n diagonal gates + n X rotations per
// Trotter layer: qumcmc/quantum_mcmc_routines.py:54-106).  Here one proposal U = M (P M)^{r-1}
// costs r sweeps over the 2^n amplitudes:
//   * K1: all ZZ/Z gates of a layer collapse into the pointwise phase exp(+i c D(idx)); D comes from
//     a fixed-point table (int32, resolution < 1e-7) and the phase is formed in integer turns
//     (wrap-around = mod 2pi for free) before MUFU sin/cos.
//   * K2: the one-body mixer factorises over qubits, M = M_lo M_hi (bits 0..11 | bits 12..n-1).
//     Sweeps alternate between a LO tile (2^13 contiguous amplitudes) and a HI tile
//     (bits [0,Lc) u [12,n)); sweep k applies the second half of layer k and, after the phase, the
//     first half of layer k+1:   LO: M_lo(k) P M_lo(k+1)    HI: M_hi(k) P M_hi(k+1).
//     The first sweep starts from the analytic product state M|s> (write-only) and the last sweep
//     (always LO) is read-only: it emits sums of |a|^2 over 64-amplitude segments.
//     Butterflies are in "tan form" a' = a - i tan(theta) b (2 FFMA per amplitude per qubit); the
//     prod cos(theta) scale of each layer is folded into the phase factor.
//     Each thread owns 64 amplitudes (6 qubits) in registers; the other qubits of the tile are
//     reached by one register<->shared-memory transposition (XOR-swizzled, conflict-free).
//   * K3: segment sums -> block scan -> inverse CDF; the selected tile is recomputed from the
//     last written state (64 KiB) to resolve the index inside the segment.
//   * K4: energies in the oracle's fixed fp64 order + Metropolis accept with Philox uniforms.
//
// Layout in HBM: state[chain][2^n] interleaved complex64 (8 B/amplitude), index order.
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int TILE_BITS = 13;
constexpr int TILE = 1 << TILE_BITS;
constexpr int NT = 128;       // threads per CTA; 64 amplitudes per thread
constexpr int LO_BITS = 12;   // qubits butterflied by the LO sweep
constexpr int SEG = 64;       // amplitudes per measurement segment (= one thread's registers)

struct ChainParams {
    unsigned long long cur;
    double e_cur;
    double u_meas, u_acc;
    long long n_acc;
    int r;       // mixer layers of this proposal (>= 1)
    int group;
    int cfx;     // round(c/(2 pi) * 2^m), c = (1-gamma) alpha dt
    int shift;   // k_fx + m - 32
    float scale; // prod_q cos(theta_q)
    int active;  // phase layer present
    float tq[QMC_MAX_SPINS];  // tan(theta_q), theta_q = gamma w_q dt
    float2 tpair[QMC_MAX_SPINS][2];  // [q][0] = (t, -t), [q][1] = (-t, t): packed FFMA2 multipliers
};

struct SweepArgs {
    int n;                    // index qubits of the local statevector (addressing)
    int n_total;              // all qubits incl. global (sharded) ones; == n on one GPU
    unsigned long long gbits; // value of the global qubits, already shifted to bit n
    float2 *state;            // [chains][2^n]
    ChainParams *params;      // [chains]
    // on-the-fly phase exponent: D_fx(idx) = A + sum_b z_b F_b + R(j) per register layout;
    // af_*[tile*128 + thread] = {A, F0, F1, F2}, {F3, F4, F5, -}; dr_*[i] = Gray-order increments of R
    const int4 *af_lo, *af_hia, *af_hib;
    int dr_lo[64], dr_hia[64], dr_hib[64];
    float *segsum;            // [chains][2^n / 64]
    float *probs_out;         // PROBS mode: unnormalised |a|^2 (index order) or null
    float2 *amps_out;
    int pf_dist;              // > 0: each CTA L2-prefetches the tile of the CTA launched pf_dist later (in launch order)
};

// ------------------------------------------------------------------------------------------
// register-level pieces
// ------------------------------------------------------------------------------------------
// Amplitudes live in registers as opaque 64-bit pairs so that the packed FP32 FMA of sm_100
// (fma.rn.f32x2 -> SASS FFMA2) consumes them without repacking.
typedef unsigned long long A64;
__device__ __forceinline__ A64 pk(float x, float y) {
    A64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 upk(A64 r) {
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}
__device__ __forceinline__ A64 fma2(A64 a, A64 b, A64 c) {
    A64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float norm2(A64 a) {
    const float2 f = upk(a);
    return f.x * f.x + f.y * f.y;
}

// Storage convention of the sweep path: amplitude idx is held as (re, im) when popcount(idx) is even and
// as (im, re) when it is odd -- in registers, shared memory and HBM alike.  Butterfly partners always
// differ in parity, so a' = a - i t b ; b' = b - i t a becomes two packed FFMA2 (fma.rn.f32x2, sm_100):
//   even' = even + (t, -t) * odd_stored      odd_stored' = odd_stored + (-t, t) * even
// `tpar` = parity of the thread's base index (all index bits outside the 6 register bits).
// reg bit b <-> qubit qpos(b); multiplier pairs are fetched per stage as opaque 64-bit values
// ([q][0] = (t,-t), [q][1] = (-t,t); a thread of odd base parity uses them crosswise).
template <int ACTIVE, typename QP>
__device__ __forceinline__ void bfly6(A64 (&v)[64], const ChainParams &cp, QP qpos, const int tpar) {
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        if (!((ACTIVE >> b) & 1)) continue;
        const A64 *p = reinterpret_cast<const A64 *>(&cp.tpair[qpos(b)][0]);
        const A64 TA = p[tpar], TB = p[tpar ^ 1];
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            if (j & (1 << b)) continue;
            const bool odd = (__builtin_popcount(j) & 1) != 0;  // folds: j is a compile-time constant
            const A64 u = v[j], w = v[j | (1 << b)];
            v[j] = fma2(odd ? TB : TA, w, u);
            v[j | (1 << b)] = fma2(odd ? TA : TB, u, w);
        }
    }
}

__device__ __forceinline__ int parity64(unsigned long long x) { return __popcll(x) & 1; }

// phase factor scale * exp(2 pi i * turns), turns = frac(D * c/(2 pi)) in 32-bit fixed point
__device__ __forceinline__ float2 phase_from_fx(int dfx, int cfx, int shift, float scale) {
    const long long p = (long long)dfx * (long long)cfx;
    const int t32 = (int)(unsigned int)(p >> shift);  // wraps mod 1 turn
    const float ang = (float)t32 * 1.4629180792671596e-09f;  // 2 pi / 2^32
    float s, c;
    __sincosf(ang, &s, &c);
    return make_float2(c * scale, s * scale);
}

__device__ __forceinline__ void apply_phase(A64 &a, const float2 f) {
    const float2 v = upk(a);
    a = pk(v.x * f.x - v.y * f.y, v.x * f.y + v.y * f.x);
}
__device__ __forceinline__ void apply_scale(A64 &a, const float g) {
    const float2 v = upk(a);
    a = pk(v.x * g, v.y * g);
}

// amplitude of M|s0> at index idx (tan form incl. the layer scale):
//   scale * prod_{q in idx^s0} (-i t_q)
__device__ __forceinline__ float2 product_state_amp(float mag, int d) {
    switch (d & 3) {
        case 0: return make_float2(mag, 0.f);
        case 1: return make_float2(0.f, -mag);
        case 2: return make_float2(-mag, 0.f);
        default: return make_float2(0.f, mag);
    }
}

__device__ __forceinline__ int lo_swz(int local) { return local ^ (((local >> 6) & 7) << 1); }

// ------------------------------------------------------------------------------------------
// LO tile: 2^13 contiguous amplitudes, qubits 0..11 butterflied.
//   A layout: regs = local bits 6..11; lane = bits 0..4; warp = bits (5, 12)   (coalesced HBM access)
//   B layout: regs = local bits 0..5 (64 contiguous amplitudes); thread = bits 6..12
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lo_base_a(int tid) {
    const int lane = tid & 31, w = tid >> 5;
    return lane | ((w & 1) << 5) | ((w >> 1) << 12);
}

__device__ __forceinline__ void lo_load_a(A64 (&v)[64], const float2 *__restrict__ tile, int tid) {
    const A64 *p = reinterpret_cast<const A64 *>(tile) + lo_base_a(tid);
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = p[j << 6];
}
__device__ __forceinline__ void lo_store_a(const A64 (&v)[64], float2 *__restrict__ tile, int tid) {
    A64 *p = reinterpret_cast<A64 *>(tile) + lo_base_a(tid);
#pragma unroll
    for (int j = 0; j < 64; ++j) p[j << 6] = v[j];
}
__device__ __forceinline__ void lo_a_to_smem(const A64 (&v)[64], float2 *sm_, int tid) {
    A64 *sm = reinterpret_cast<A64 *>(sm_);
    const int base = lo_base_a(tid);
#pragma unroll
    for (int j = 0; j < 64; ++j) sm[lo_swz(base | (j << 6))] = v[j];
}
__device__ __forceinline__ void lo_smem_to_a(A64 (&v)[64], const float2 *sm_, int tid) {
    const A64 *sm = reinterpret_cast<const A64 *>(sm_);
    const int base = lo_base_a(tid);
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = sm[lo_swz(base | (j << 6))];
}
__device__ __forceinline__ void lo_b_to_smem(const A64 (&v)[64], float2 *sm_, int tid) {
    A64 *sm = reinterpret_cast<A64 *>(sm_);
#pragma unroll
    for (int j = 0; j < 64; j += 2)
        *reinterpret_cast<ulonglong2 *>(&sm[lo_swz((tid << 6) | j)]) = make_ulonglong2(v[j], v[j + 1]);
}
__device__ __forceinline__ void lo_smem_to_b(A64 (&v)[64], const float2 *sm_, int tid) {
    const A64 *sm = reinterpret_cast<const A64 *>(sm_);
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(&sm[lo_swz((tid << 6) | j)]);
        v[j] = x.x;
        v[j + 1] = x.y;
    }
}

// K1, evaluated without a table lookup in the inner loop.  For a thread whose 6 register bits sit
// at qubits p_0..p_5 and whose other index bits are fixed:
//     D(idx) = A + sum_b z_b F_b + R(j),   z_b = +1 / -1 for register bit b clear / set,
// A, F_b (fixed-point int32, exact fp64 values rounded once) come from two 16-byte loads issued at
// the top of the kernel; R(j) couples the register qubits among themselves and enters as a
// compile-time-indexed constant-bank operand.  The 64 amplitudes are visited in Gray-code order so
// that one IADD3 updates D per amplitude.  Odd-parity amplitudes are stored swapped, for which
// multiplying by exp(i phi) is the same formula with phi -> -phi: the rate cfx carries the sign.
struct PhaseAF {
    int4 a0, a1;
};
__device__ __forceinline__ PhaseAF load_af(const int4 *__restrict__ af, int64_t f) {
    PhaseAF r;
    r.a0 = __ldg(af + 2 * f);
    r.a1 = __ldg(af + 2 * f + 1);
    return r;
}
// Per-CTA table of the Gray-order increments of R(j) in 32-bit turns (c/(2 pi) * dR, wraps mod 1 turn),
// hs[0..63] as is and hs[64..127] negated (threads of odd base parity run the whole recurrence negated,
// see phase_gray).  cfx/shift are read straight from the chain's parameter record so that the table is
// complete at the barrier that ends the parameter staging.
__device__ __forceinline__ int turns32(int dfx, int cfx, int shift) {
    return (int)(unsigned int)(((long long)dfx * (long long)cfx) >> shift);
}
template <int NTH>
__device__ __forceinline__ void build_hs(int *hs, const int (&dr)[64], const ChainParams *cpg, int tid) {
    const int cfx = cpg->cfx, shift = cpg->shift;
#pragma unroll
    for (int i = tid; i < 128; i += NTH) {
        const int t = turns32(dr[i & 63], cfx, shift);
        hs[i] = (i & 64) ? -t : t;
    }
}

// K1 for the thread's 64 amplitudes.  The phase angle is accumulated directly in 32-bit integer turns along the
// Gray-code order: t(j') = t(j) +/- G_b + H_i with G_b = turns(2 F_b) (per thread) and H_i = turns(dR_i) (per
// chain, shared memory) -- one IADD3 per amplitude, wrap-around = mod 2 pi for free, rounding <= 64 * 2^-33 turn.
// Amplitudes of odd index parity are stored swapped, for which multiplying by exp(i phi) is the same
// formula with phi -> -phi: compile-time parity of j picks the sign of sin, the base parity tpar negates
// the whole recurrence.  FOLD: the per-layer normalisation prod cos(theta_q) has been folded into the initial
// product state (cp.scale = scale^r), so the factor has unit modulus; otherwise it is scale * exp(i phi).
template <bool FOLD>
__device__ __forceinline__ void phase_gray(A64 (&v)[64], const PhaseAF af, const int *hs_all, const int dr0,
                                           const ChainParams &cp, const int tpar) {
    if (cp.active) {
        const int cfx = tpar ? -cp.cfx : cp.cfx, shift = cp.shift;
        const int F[6] = {af.a0.y, af.a0.z, af.a0.w, af.a1.x, af.a1.y, af.a1.z};
        int G[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) G[b] = turns32(2 * F[b], cfx, shift);
        int t = turns32(af.a0.x + F[0] + F[1] + F[2] + F[3] + F[4] + F[5] + dr0, cfx, shift);
        const int4 *hs = reinterpret_cast<const int4 *>(hs_all + (tpar << 6));
        const float sc = cp.scale;
        int4 hn = hs[0];
#pragma unroll
        for (int i4 = 0; i4 < 16; ++i4) {
            const int4 h4 = hn;  // increments H_{4 i4 .. 4 i4 + 3}; the step i -> i+1 adds H_{i+1}
            if (i4 < 15) hn = hs[i4 + 1];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * i4 + k;
                const int j = i ^ (i >> 1);
                const bool odd = (__builtin_popcount(j) & 1) != 0;
                float sn, cs;
                __sincosf((float)t * 1.4629180792671596e-09f, &sn, &cs);  // 2 pi / 2^32
                if (!FOLD) {
                    sn *= sc;
                    cs *= sc;
                }
                const float2 a = upk(v[j]);
                v[j] = odd ? pk(a.x * cs + a.y * sn, a.y * cs - a.x * sn) : pk(a.x * cs - a.y * sn, a.y * cs + a.x * sn);
                if (i < 63) {
                    const int i1 = i + 1;
                    const int b = (i1 & 1) ? 0 : (i1 & 2) ? 1 : (i1 & 4) ? 2 : (i1 & 8) ? 3 : (i1 & 16) ? 4 : 5;
                    const int jn = i1 ^ (i1 >> 1);
                    const int h = (k == 0) ? h4.y : (k == 1) ? h4.z : (k == 2) ? h4.w : hn.x;
                    t = ((jn >> b) & 1) ? t - G[b] + h : t + G[b] + h;
                }
            }
        }
    } else if (!FOLD) {
#pragma unroll
        for (int j = 0; j < 64; ++j) apply_scale(v[j], cp.scale);
    }
}

// M|s0> restricted to this thread's 64 amplitudes; reg bit b <-> global qubit qpos[b]
template <typename QP>
__device__ __forceinline__ void init_product(A64 (&v)[64], uint64_t base_idx, int n, const ChainParams &cp, QP qpos) {
    const int tpar = parity64(base_idx);
    // fixed part: all qubits outside the register bits
    uint64_t regmask = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) regmask |= 1ULL << qpos(b);
    const uint64_t diff = (base_idx ^ cp.cur) & ~regmask;
    float mag0 = cp.scale;
    for (int q = 0; q < n; ++q)
        if ((diff >> q) & 1ULL) mag0 *= cp.tq[q];
    const int d0 = __popcll(diff);
    float f[6];
    int sbit[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        f[b] = cp.tq[qpos(b)];
        sbit[b] = (int)((cp.cur >> qpos(b)) & 1ULL);
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float mag = mag0;
        int d = d0;
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const int differs = ((j >> b) & 1) ^ sbit[b];
            mag *= differs ? f[b] : 1.0f;
            d += differs;
        }
        const float2 amp = product_state_amp(mag, d);
        const bool odd = ((__builtin_popcount(j) & 1) ^ tpar) != 0;
        v[j] = odd ? pk(amp.y, amp.x) : pk(amp.x, amp.y);
    }
}

struct LoQ0 { __device__ __forceinline__ int operator()(int b) const { return b; } };       // B layout: bits 0..5
struct LoQ6 { __device__ __forceinline__ int operator()(int b) const { return 6 + b; } };   // A layout: bits 6..11

// Final LO computation for (chain, tile): leaves the final amplitudes of the tile in B layout.
// Used by the last sweep (segment sums) and by the sampler (recompute of the selected tile).
// A LO tile is NTH * 64 contiguous amplitudes: 2^13 for 128 threads, 2^12 for 64 threads (bit 12 is a pure
// batch bit of the LO butterflies, so a 128-thread tile is two independent 64-thread tiles).
template <int NTH>
__device__ __forceinline__ void lo_final_tile(A64 (&v)[64], float2 *sm, const float2 *__restrict__ state_chain,
                                              int64_t tile, int tid, int n, const ChainParams &cp,
                                              unsigned long long gb = 0ULL) {
    constexpr int TB = NTH == 128 ? 13 : 12;
    if (cp.r == 1) {
        init_product(v, gb | ((uint64_t)tile << TB) | ((uint64_t)tid << 6), n, cp, LoQ0());
        return;
    }
    lo_load_a(v, state_chain + (tile << TB), tid);
    bfly6<63>(v, cp, LoQ6(), parity64(gb | ((unsigned long long)tile << TB) | (unsigned)lo_base_a(tid)));
    lo_a_to_smem(v, sm, tid);
    __syncthreads();
    lo_smem_to_b(v, sm, tid);
    bfly6<63>(v, cp, LoQ0(), parity64(gb | ((unsigned long long)tile << TB) | ((unsigned)tid << 6)));
}

// ------------------------------------------------------------------------------------------
// HI tile for W = n - 12 X-qubits: local bits [0,Lc) <-> global [0,Lc), local [Lc,13) <-> global [12,n)
//   A' layout: regs = local 7..12; lane = local 0..4; warp = local (5, 6)
//   B' layout (W > 6 only): regs = local 5..10; lane = local 0..4; warp = local (11, 12)
// Both layouts keep lanes on the 32 lowest amplitudes => coalesced HBM access and conflict-free smem.
// ------------------------------------------------------------------------------------------
template <int W>
struct HiGeom {
    static constexpr int Lc = TILE_BITS - W;
    __host__ __device__ static constexpr int gpos(int i) { return i < Lc ? i : 12 + (i - Lc); }
    __host__ __device__ static constexpr int64_t off_a(int j) {  // regs = local 7..12
        int64_t o = 0;
        for (int b = 0; b < 6; ++b)
            if ((j >> b) & 1) o |= (int64_t)1 << gpos(7 + b);
        return o;
    }
    __host__ __device__ static constexpr int64_t off_b(int j) {  // regs = local 5..10
        int64_t o = 0;
        for (int b = 0; b < 6; ++b)
            if ((j >> b) & 1) o |= (int64_t)1 << gpos(5 + b);
        return o;
    }
    static constexpr int active_a = W >= 6 ? 63 : (63 & ~((1 << (6 - W)) - 1));
    static constexpr int active_b = W == 8 ? 3 : (W == 7 ? 2 : 0);
    static constexpr bool two_rounds = W > 6;
};

template <int W> struct HiQA { __device__ __forceinline__ int operator()(int b) const { return HiGeom<W>::gpos(7 + b); } };
template <int W> struct HiQB { __device__ __forceinline__ int operator()(int b) const { return HiGeom<W>::gpos(5 + b); } };

template <int W>
__device__ __forceinline__ int64_t hi_base_a(int tid, int64_t tile) {
    using G = HiGeom<W>;
    const int lane = tid & 31, w = tid >> 5;
    return (int64_t)lane | ((int64_t)(w & 1) << G::gpos(5)) | ((int64_t)(w >> 1) << G::gpos(6)) | (tile << G::Lc);
}
template <int W>
__device__ __forceinline__ int64_t hi_base_b(int tid, int64_t tile) {
    using G = HiGeom<W>;
    const int lane = tid & 31, w = tid >> 5;
    return (int64_t)lane | ((int64_t)(w & 1) << G::gpos(11)) | ((int64_t)(w >> 1) << G::gpos(12)) | (tile << G::Lc);
}
__device__ __forceinline__ int hi_loc_a(int tid, int j) { return (tid & 31) | ((tid >> 5) << 5) | (j << 7); }
__device__ __forceinline__ int hi_loc_b(int tid, int j) { return (tid & 31) | (j << 5) | ((tid >> 5) << 11); }

// L2 prefetch of the tile a later CTA of the same launch will read (CTAs start in linear blockIdx order,
// x fastest): with 3 resident CTAs per SM the CTA `pf_dist` ahead starts roughly when this one retires,
// so its loads hit L2 instead of waiting on HBM with nothing else to overlap.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool pf_target(const SweepArgs &a, const int *__restrict__ order, int list_off, int64_t &chain,
                                          int64_t &tile) {
    if (a.pf_dist <= 0) return false;
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.pf_dist;
    if (lin >= (long long)gridDim.x * gridDim.y) return false;
    const long long y = lin / gridDim.x;
    tile = lin - y * gridDim.x;
    chain = order[list_off + y];
    return true;
}

// ------------------------------------------------------------------------------------------
// sweep kernels: grid = (tiles, chains of one role); chain = order[list_off + blockIdx.y].
// One kernel per role so that every CTA of a launch runs the same straight-line code
// (small instruction footprint) and no CTA exits early:
//   LO_FIRST  r odd >= 3, step 1   : M|s> analytic, P, M_lo(2)                  (write-only)
//   LO_MID    1 < step < r         : M_lo(k) P M_lo(k+1)                         (read + write)
//   LO_FINAL  step == r            : M_lo(r), |a|^2 segment sums                 (read-only)
//   HI_FIRST  r even, step 1       : M|s> analytic, P, M_hi(2)                   (write-only)
//   HI_MID                         : M_hi(k) P M_hi(k+1)                         (read + write)
// ------------------------------------------------------------------------------------------
template <int NTH = NT>
__device__ __forceinline__ void stage_params(ChainParams &dst_s, const ChainParams *src_g, int tid) {
    const int *src = reinterpret_cast<const int *>(src_g);
    int *dst = reinterpret_cast<int *>(&dst_s);
    for (int i = tid; i < (int)(sizeof(ChainParams) / 4); i += NTH) dst[i] = src[i];
    __syncthreads();
}

enum { LO_FIRST = 0, LO_MID = 1, LO_FINAL = 2, LO_SINGLE = 3 };  // SINGLE: M_lo only (3+ qubit groups)

template <int ROLE, int NTH, bool FOLD>
__global__ void __launch_bounds__(NTH, 256 / NTH) lo_sweep_kernel(const SweepArgs a, const int *__restrict__ order, const int list_off) {
    constexpr int TB = NTH == 128 ? 13 : 12;  // tile bits
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[128];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    const int n = a.n;
    float2 *state = a.state + (chain << n);
    A64 v[64];
    if (ROLE == LO_MID) lo_load_a(v, state + (tile << TB), tid);  // in flight while the parameters are staged
    if (ROLE != LO_FIRST && tid == 0) {
        int64_t pc, pt;
        if (pf_target(a, order, list_off, pc, pt)) l2_prefetch_bulk(a.state + (pc << n) + (pt << TB), (1u << TB) * 8);
    }
    if (ROLE == LO_FIRST || ROLE == LO_MID) build_hs<NTH>(hs_s, a.dr_lo, a.params + chain, tid);
    stage_params<NTH>(cp_s, a.params + chain, tid);
    const ChainParams &cp = cp_s;
    if (ROLE == LO_FINAL) {
        lo_final_tile<NTH>(v, sm, state, tile, tid, a.n_total, cp, a.gbits);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) s += norm2(v[j]);
        a.segsum[(chain << (n - 6)) + tile * NTH + tid] = s;
        if (a.probs_out) {
            float *po = a.probs_out + (chain << n) + (tile << TB) + (tid << 6);
#pragma unroll
            for (int j = 0; j < 64; ++j) po[j] = norm2(v[j]);
        }
        if (a.amps_out) {
            float2 *ao = a.amps_out + (chain << n) + (tile << TB) + (tid << 6);
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const bool odd = ((__builtin_popcount(j) & 1) ^ parity64(a.gbits | ((unsigned long long)tile << TB) | ((unsigned)tid << 6))) != 0;
                const float2 f = upk(v[j]);
                ao[j] = odd ? make_float2(f.y, f.x) : f;
            }
        }
        return;
    }
    const int par_a = parity64(a.gbits | ((unsigned long long)tile << TB) | (unsigned)lo_base_a(tid));
    const int par_b = parity64(a.gbits | ((unsigned long long)tile << TB) | ((unsigned)tid << 6));
    if (ROLE == LO_SINGLE) {
        lo_load_a(v, state + (tile << TB), tid);
        bfly6<63>(v, cp, LoQ6(), par_a);
        lo_a_to_smem(v, sm, tid);
        __syncthreads();
        lo_smem_to_b(v, sm, tid);
        bfly6<63>(v, cp, LoQ0(), par_b);
        lo_b_to_smem(v, sm, tid);  // each thread rewrites exactly the slots it read
        __syncthreads();
        lo_smem_to_a(v, sm, tid);
        lo_store_a(v, state + (tile << TB), tid);
        return;
    }
    const PhaseAF af = load_af(a.af_lo, tile * NTH + tid);  // issued before the amplitude loads are consumed
    if (ROLE == LO_FIRST) {
        init_product(v, ((uint64_t)tile << TB) | ((uint64_t)tid << 6), n, cp, LoQ0());
    } else {
        bfly6<63>(v, cp, LoQ6(), par_a);
        lo_a_to_smem(v, sm, tid);
        __syncthreads();
        lo_smem_to_b(v, sm, tid);
        bfly6<63>(v, cp, LoQ0(), par_b);
    }
    phase_gray<FOLD>(v, af, hs_s, a.dr_lo[0], cp, par_b);
    bfly6<63>(v, cp, LoQ0(), par_b);
    lo_b_to_smem(v, sm, tid);  // each thread rewrites exactly the slots it read: no barrier needed before
    __syncthreads();
    lo_smem_to_a(v, sm, tid);
    bfly6<63>(v, cp, LoQ6(), par_a);
    lo_store_a(v, state + (tile << TB), tid);
}

template <int W, bool FIRST, bool FOLD>
__global__ void __launch_bounds__(NT, 3) hi_sweep_kernel(const SweepArgs a, const int *__restrict__ order, const int list_off) {
    using G = HiGeom<W>;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[128];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    const int n = a.n;
    float2 *state = a.state + (chain << n);
    A64 v[64];
    const int64_t base_a = hi_base_a<W>(tid, tile);
    if (!FIRST) {  // in flight while the parameters are staged
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = reinterpret_cast<const A64 *>(state)[base_a + G::off_a(j)];
        int64_t pc, pt;
        if (pf_target(a, order, list_off, pc, pt)) {
            // tile rows: 2^W rows of 2^Lc contiguous amplitudes, row stride 2^12 amplitudes
            constexpr int ROWS = 1 << W, ROW_BYTES = (1 << G::Lc) * 8;
            const float2 *src = a.state + (pc << n) + (pt << G::Lc);
            for (int r = tid; r < ROWS; r += NT) l2_prefetch_bulk(src + ((int64_t)r << 12), ROW_BYTES);
        }
    }
    build_hs<NT>(hs_s, G::two_rounds ? a.dr_hib : a.dr_hia, a.params + chain, tid);
    stage_params(cp_s, a.params + chain, tid);
    const ChainParams &cp = cp_s;
    const PhaseAF af = load_af(G::two_rounds ? a.af_hib : a.af_hia, tile * NT + tid);
    if (G::two_rounds) {
        const int64_t base_b = hi_base_b<W>(tid, tile);
        if (FIRST) {
            init_product(v, (uint64_t)base_b, n, cp, HiQB<W>());
        } else {
            bfly6<G::active_a>(v, cp, HiQA<W>(), parity64((unsigned long long)base_a));
#pragma unroll
            for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(sm)[hi_loc_a(tid, j)] = v[j];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = reinterpret_cast<const A64 *>(sm)[hi_loc_b(tid, j)];
            bfly6<G::active_b>(v, cp, HiQB<W>(), parity64((unsigned long long)base_b));
            __syncthreads();
        }
        phase_gray<FOLD>(v, af, hs_s, a.dr_hib[0], cp, parity64((unsigned long long)base_b));
        bfly6<G::active_b>(v, cp, HiQB<W>(), parity64((unsigned long long)base_b));
#pragma unroll
        for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(sm)[hi_loc_b(tid, j)] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = reinterpret_cast<const A64 *>(sm)[hi_loc_a(tid, j)];
        bfly6<G::active_a>(v, cp, HiQA<W>(), parity64((unsigned long long)base_a));
#pragma unroll
        for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(state)[base_a + G::off_a(j)] = v[j];
    } else {
        if (FIRST) {
            init_product(v, (uint64_t)base_a, n, cp, HiQA<W>());
        } else {
            bfly6<G::active_a>(v, cp, HiQA<W>(), parity64((unsigned long long)base_a));
        }
        phase_gray<FOLD>(v, af, hs_s, a.dr_hia[0], cp, parity64((unsigned long long)base_a));
        bfly6<G::active_a>(v, cp, HiQA<W>(), parity64((unsigned long long)base_a));
#pragma unroll
        for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(state)[base_a + G::off_a(j)] = v[j];
    }
}

// ------------------------------------------------------------------------------------------
// Generic large-n path (n >= 21: three or more qubit groups).  LO group = qubits 0..11 (kernels
// above, roles SINGLE / FINAL); the remaining qubits are covered by register-only HI6 sweeps:
// a thread owns the 64 amplitudes over qubits [p, p+6), lanes sit on qubits 0..4, so every access is
// a coalesced 256 B segment and no shared-memory exchange is needed.  Roles:
//   HI6_FIRST   M|s> analytic, P, M_g(2)       HI6_MID   M_g(k) P M_g(k+1)       HI6_SINGLE   M_g(k)
// The phase exponent fields A, F_b are evaluated per thread in fp64 from the couplings (staged in
// shared memory) -- no table, any n.  ACT = number of active (top) register qubits of the group.
// ------------------------------------------------------------------------------------------
enum { HI6_FIRST = 0, HI6_MID = 1, HI6_SINGLE = 2 };

struct HiP {
    int p;
    __device__ __forceinline__ int operator()(int b) const { return p + b; }
};

struct Hi6Args {
    int p;          // first register qubit
    int dr[64];     // Gray-order increments of R(j) for this group
    const double *Jq, *hq;  // circuit couplings in qubit order
    double fx_scale;        // 2^k_fx
};

template <int ROLE, int ACT>
__global__ void __launch_bounds__(NT, 3) hi6_sweep_kernel(const SweepArgs a, const Hi6Args g, const int *__restrict__ order,
                                                          const int list_off) {
    extern __shared__ __align__(16) double smd[];  // Jq[n*n], hq[n]
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[128];
    const int tid = threadIdx.x;
    const int n = a.n;
    const int64_t tile = blockIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    if (ROLE != HI6_SINGLE) build_hs<NT>(hs_s, g.dr, a.params + chain, tid);
    if (ROLE != HI6_SINGLE) {
        const int nt_ = a.n_total;
        for (int i = tid; i < nt_ * nt_; i += NT) smd[i] = g.Jq[i];
        for (int i = tid; i < nt_; i += NT) smd[nt_ * nt_ + i] = g.hq[i];
    }
    stage_params(cp_s, a.params + chain, tid);  // barrier inside
    const ChainParams &cp = cp_s;
    const int p = g.p;
    constexpr int ACTIVE = 63 & ~((1 << (6 - ACT)) - 1);
    // index bits: lane 0..4, warp bits 5,6, tile-low bits 7..p-1, register bits p..p+5, tile-high bits p+6..n-1
    const int low_bits = p - 7;
    const int64_t base = (int64_t)(tid & 31) | ((int64_t)(tid >> 5) << 5) | ((tile & (((int64_t)1 << low_bits) - 1)) << 7) |
                         ((tile >> low_bits) << (p + 6));
    A64 *state = reinterpret_cast<A64 *>(a.state + (chain << n)) + base;
    const unsigned long long full = (unsigned long long)base | a.gbits;  // incl. the global (sharded) qubits
    const int nt = a.n_total;
    const int tpar = parity64(full);
    const int64_t stride = (int64_t)1 << p;
    const HiP qp{p};
    // phase exponent fields first (fp64, registers still free): A = sum_{a<b fixed} Jq z z + sum_{a fixed} hq z,
    // F_b = hq[p+b] + sum_{a fixed} Jq[a][p+b] z_a
    PhaseAF af;
    af.a0 = make_int4(0, 0, 0, 0);
    af.a1 = af.a0;
    if (ROLE != HI6_SINGLE && cp.active) {
        const double *Jq = smd, *hq = smd + nt * nt;
        double A = 0.0, F[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) F[b] = hq[p + b];
        for (int q = 0; q < nt; ++q) {
            if (q >= p && q < p + 6) continue;
            const double zq = ((full >> q) & 1ULL) ? -1.0 : 1.0;
            double acc = hq[q];
            for (int q2 = q + 1; q2 < nt; ++q2) {
                if (q2 >= p && q2 < p + 6) continue;
                acc += ((full >> q2) & 1ULL) ? -Jq[q * nt + q2] : Jq[q * nt + q2];
            }
            A += zq * acc;
#pragma unroll
            for (int b = 0; b < 6; ++b) F[b] += Jq[q * nt + p + b] * zq;
        }
        af.a0 = make_int4((int)rint(A * g.fx_scale), (int)rint(F[0] * g.fx_scale), (int)rint(F[1] * g.fx_scale),
                          (int)rint(F[2] * g.fx_scale));
        af.a1 = make_int4((int)rint(F[3] * g.fx_scale), (int)rint(F[4] * g.fx_scale), (int)rint(F[5] * g.fx_scale), 0);
    }
    A64 v[64];
    if (ROLE == HI6_FIRST) {
        init_product(v, full, nt, cp, qp);
    } else {
        const A64 *src = state;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            v[j] = *src;
            src += stride;
        }
        bfly6<ACTIVE>(v, cp, qp, tpar);
    }
    if (ROLE != HI6_SINGLE) {
        phase_gray<false>(v, af, hs_s, g.dr[0], cp, tpar);
        bfly6<ACTIVE>(v, cp, qp, tpar);
    }
    A64 *dst = state;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        *dst = v[j];
        dst += stride;
    }
}

// Device-side schedule: order[] = chains with odd r sorted by r descending, then chains with even r
// sorted by r descending.  The host replays the same Philox draws to know the class sizes.
constexpr int SCHED_RMAX = 2047;
__global__ void __launch_bounds__(1024) schedule_kernel(const ChainParams *__restrict__ params, int64_t n_chains,
                                                        int r_cap, int *__restrict__ order, int by_parity) {
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
                if (by_parity ? (r & 1) == parity : parity == 1) {  // by_parity = 0: plain r-descending order
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

// ------------------------------------------------------------------------------------------
// per-hop bookkeeping kernels
// ------------------------------------------------------------------------------------------
struct HopArgs {
    int n, mode, n_groups, phase_active, k_fx;
    int fold;  // 1: cp.scale = (prod cos)^r, applied once in the initial product state
    int64_t n_chains, n_hops, hop;
    uint64_t chain_id0, seed;
    uint32_t hop0;
    double temperature, g_lo, g_hi, dt, alpha;
    const double *J, *h;
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
    int *hist;  // [chains][n+1] running histogram
};

__device__ void fill_proposal_params(ChainParams &cp, const HopArgs &h, double gamma, int r, int grp) {
    cp.r = r < 1 ? 1 : r;
    cp.group = grp;
    for (int q = 0; q < QMC_MAX_SPINS; ++q) cp.tq[q] = 0.f;
    double scale = 1.0;
    for (int k = h.group_off[grp]; k < h.group_off[grp + 1]; ++k) {
        const int q = __ffsll((long long)h.masks[k]) - 1;
        double s, c;
        sincos(gamma * h.weights[k] * h.dt, &s, &c);
        cp.tq[q] = (float)(s / c);
        scale *= c;
    }
    cp.scale = h.fold ? (float)pow(scale, (double)cp.r) : (float)scale;
    for (int q = 0; q < QMC_MAX_SPINS; ++q) {
        cp.tpair[q][0] = make_float2(cp.tq[q], -cp.tq[q]);
        cp.tpair[q][1] = make_float2(-cp.tq[q], cp.tq[q]);
    }
    const double cprime = (1.0 - gamma) * h.alpha * h.dt * 0.15915494309189533577;  // c / (2 pi)
    cp.active = h.phase_active && cp.r > 1 && cprime != 0.0;
    int m = 0;
    if (cp.active) {
        m = (int)floor(log2(2147483647.0 / fabs(cprime)));
        if (m > 40) m = 40;
        if (m < 0) m = 0;
    }
    cp.cfx = cp.active ? (int)rint(ldexp(cprime, m)) : 0;
    cp.shift = h.k_fx + m - 32;
}

__global__ void chains_init_kernel(const HopArgs h, ChainParams *params) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= h.n_chains) return;
    ChainParams &cp = params[c];
    uint64_t cur;
    if (h.mode == QMC_MODE_CHAIN)
        cur = h.s_in ? h.s_in[c] : qmc_initial_state(h.seed, h.chain_id0 + (uint64_t)c, h.n);
    else
        cur = h.s_in[c];
    cp.cur = cur;
    cp.e_cur = (h.mode == QMC_MODE_CHAIN) ? qmc_energy_dev(h.n, h.J, h.h, cur) : 0.0;
    cp.n_acc = 0;
    for (int i = 0; i <= h.n; ++i) h.hist[c * (h.n + 1) + i] = 0;
}

__global__ void hop_prepare_kernel(const HopArgs h, ChainParams *params) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= h.n_chains) return;
    ChainParams &cp = params[c];
    if (h.mode == QMC_MODE_CHAIN) {
        const HopDraws d = qmc_hop_draws(h.seed, h.chain_id0 + (uint64_t)c, h.hop0 + (uint32_t)h.hop);
        const double gamma = h.g_lo + (h.g_hi - h.g_lo) * d.u_gamma;
        cp.u_meas = d.u_meas;
        cp.u_acc = d.u_acc;
        fill_proposal_params(cp, h, gamma, qmc_trotter_steps(d.t, h.dt), qmc_pick_group(h.n_groups, h.group_cum, d.u_group));
    } else {
        cp.u_meas = h.u_arr ? h.u_arr[c] : 0.0;
        cp.u_acc = 0.0;
        fill_proposal_params(cp, h, h.gamma_arr[c], h.r_arr[c], h.group_arr ? h.group_arr[c] : 0);
    }
}

// K3 + K4: one CTA per chain
__global__ void __launch_bounds__(NT) sample_accept_kernel(const SweepArgs a, const HopArgs h) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ double s_incl[NT];
    __shared__ double s_warp[NT / 32];
    __shared__ long long s_seg;
    __shared__ double s_rem;
    __shared__ unsigned long long s_sel;
    const int tid = threadIdx.x;
    const int64_t chain = blockIdx.x;
    const int n = a.n;
    {
        const int *src = reinterpret_cast<const int *>(a.params + chain);
        int *dst = reinterpret_cast<int *>(&cp_s);
        for (int i = tid; i < (int)(sizeof(ChainParams) / 4); i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const ChainParams &cp = cp_s;
    const int64_t nseg = 1LL << (n - 6);
    const int64_t per = nseg / NT;  // >= 2 for n >= 14
    const float *seg = a.segsum + (chain << (n - 6));
    double local = 0.0;
    for (int64_t i = 0; i < per; ++i) local += (double)seg[tid * per + i];
    double incl = local;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double x = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    incl += base;
    s_incl[tid] = incl;
    __syncthreads();
    const double total = s_incl[NT - 1];
    const double target = cp.u_meas * total;
    const double lo = tid == 0 ? 0.0 : s_incl[tid - 1];
    if (target >= lo && (target < incl || tid == NT - 1)) {
        double c = lo;  // cumulative sum before segment i
        int64_t sel = -1;
        double before = 0.0;
        for (int64_t i = 0; i < per; ++i) {
            const double x = (double)seg[tid * per + i];
            if (target < c + x) {
                sel = tid * per + i;
                before = c;
                break;
            }
            c += x;
        }
        if (sel < 0) {  // rounding tail: last segment of this thread
            sel = tid * per + per - 1;
            before = c - (double)seg[sel];
        }
        s_seg = sel;
        s_rem = target - before;
    }
    __syncthreads();
    const int64_t sg = s_seg;
    const int64_t tile = sg / NT;
    const int owner = (int)(sg % NT);
    A64 v[64];
    lo_final_tile<NT>(v, sm, a.state + ((int64_t)chain << n), tile, tid, n, cp);
    if (tid == owner) {
        const double rem = s_rem;
        double c = 0.0;
        int selj = 63;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            c += (double)norm2(v[j]);
            if (selj == 63 && rem < c) selj = j;
        }
        s_sel = ((unsigned long long)tile << TILE_BITS) | ((unsigned long long)owner << 6) | (unsigned long long)selj;
    }
    __syncthreads();
    if (tid == 0) {
        const uint64_t sp = s_sel;
        ChainParams &out = a.params[chain];
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
        }
    }
}

__global__ void chains_finish_kernel(const HopArgs h, const ChainParams *params) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= h.n_chains) return;
    if (h.final_state) h.final_state[c] = params[c].cur;
    if (h.acc_count) h.acc_count[c] = params[c].n_acc;
    if (h.hamming_hist)
        for (int i = 0; i <= h.n; ++i) h.hamming_hist[c * (h.n + 1) + i] = h.hist[c * (h.n + 1) + i];
}

// PROBS mode: divide by the total norm (tan-form butterflies are unnormalised only by rounding)
__global__ void probs_normalize_kernel(int n, const float *segsum, float *probs, float2 *amps) {
    __shared__ double red[256];
    const int64_t nseg = 1LL << (n - 6);
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nseg; i += blockDim.x) s += (double)segsum[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    const double total = red[0];
    const float inv = (float)(1.0 / total), inva = (float)(1.0 / sqrt(total));
    const int64_t N = 1LL << n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        if (probs) probs[i] *= inv;
        if (amps) {
            amps[i].x *= inva;
            amps[i].y *= inva;
        }
    }
}

// A / F_b tables of one register layout: LAYOUT 0 = LO B (regs = qubits 0..5), 1 = HI A', 2 = HI B'.
// Jq/hq are the circuit couplings in qubit order (qubit a = spin n-1-a); values are exact fp64 sums
// rounded once to fixed point (scale 2^k_fx).
template <int W, int LAYOUT>
__global__ void af_table_kernel(int n, const double *__restrict__ Jq, const double *__restrict__ hq, double scale,
                                int4 *__restrict__ af) {
    const int64_t total = (int64_t)1 << (n - 6);
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < total; f += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = f >> 7;
        const int tid = (int)(f & 127);
        uint64_t base;
        int p[6];
        for (int b = 0; b < 6; ++b) p[b] = LAYOUT == 0 ? b : (LAYOUT == 1 ? HiGeom<W>::gpos(7 + b) : HiGeom<W>::gpos(5 + b));
        if (LAYOUT == 0) base = ((uint64_t)tile << TILE_BITS) | ((uint64_t)tid << 6);
        else if (LAYOUT == 1) base = (uint64_t)hi_base_a<W>(tid, tile);
        else base = (uint64_t)hi_base_b<W>(tid, tile);
        uint64_t regmask = 0;
        for (int b = 0; b < 6; ++b) regmask |= 1ULL << p[b];
        double A = 0.0, F[6];
        for (int b = 0; b < 6; ++b) F[b] = hq[p[b]];
        for (int q = 0; q < n; ++q) {
            if ((regmask >> q) & 1ULL) continue;
            const double zq = ((base >> q) & 1ULL) ? -1.0 : 1.0;
            A += hq[q] * zq;
            for (int q2 = q + 1; q2 < n; ++q2) {
                if ((regmask >> q2) & 1ULL) continue;
                const double z2 = ((base >> q2) & 1ULL) ? -1.0 : 1.0;
                A += Jq[q * n + q2] * zq * z2;
            }
            for (int b = 0; b < 6; ++b) F[b] += Jq[q * n + p[b]] * zq;
        }
        af[2 * f] = make_int4((int)rint(A * scale), (int)rint(F[0] * scale), (int)rint(F[1] * scale), (int)rint(F[2] * scale));
        af[2 * f + 1] = make_int4((int)rint(F[3] * scale), (int)rint(F[4] * scale), (int)rint(F[5] * scale), 0);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct SweepWorkspace {
    int64_t cap_chains = 0;
    float2 *state = nullptr;
    ChainParams *params = nullptr;
    float *segsum = nullptr;
    int *hist = nullptr;
    int *order = nullptr;
    int4 *af_lo = nullptr, *af_hia = nullptr, *af_hib = nullptr;
    int dr_lo[64], dr_hia[64], dr_hib[64];
    bool tables_ready = false;
    std::vector<double> Jq, hq;   // circuit couplings in qubit order (host)
    double *dJq = nullptr, *dhq = nullptr;
    double fx_scale = 1.0;
    int k_fx = 0;
    std::vector<cudaEvent_t> ev;
};

void qmc_sweep_free(qmc_model *m) {
    SweepWorkspace *w = m->sweep;
    if (!w) return;
    cudaFree(w->state);
    cudaFree(w->params);
    cudaFree(w->segsum);
    cudaFree(w->hist);
    cudaFree(w->order);
    cudaFree(w->af_lo);
    cudaFree(w->af_hia);
    cudaFree(w->af_hib);
    cudaFree(w->dJq);
    cudaFree(w->dhq);
    for (auto e : w->ev) cudaEventDestroy(e);
    delete w;
    m->sweep = nullptr;
}

// R(j) = sum_{b<b'} Jq[p_b][p_b'] z_b z_b' over the 6 register qubits, as Gray-order increments
static void gray_increments(const std::vector<double> &Jq, int n, double scale, const int (&pos)[6], int (&dr)[64]) {
    int R[64];
    for (int j = 0; j < 64; ++j) {
        double r = 0.0;
        for (int b = 0; b < 6; ++b)
            for (int b2 = b + 1; b2 < 6; ++b2) {
                const double zb = ((j >> b) & 1) ? -1.0 : 1.0, zb2 = ((j >> b2) & 1) ? -1.0 : 1.0;
                r += Jq[(size_t)pos[b] * n + pos[b2]] * zb * zb2;
            }
        R[j] = (int)rint(r * scale);
    }
    dr[0] = R[0];
    for (int i = 1; i < 64; ++i) dr[i] = R[i ^ (i >> 1)] - R[(i - 1) ^ ((i - 1) >> 1)];
}

template <int W>
static int build_af_w(int n, const double *dJq, const double *dhq, double scale, int4 *lo, int4 *hia, int4 *hib,
                      int (&pos)[3][6], cudaStream_t st) {
    const int grid = (int)std::min<int64_t>(148 * 8, (((int64_t)1 << (n - 6)) + 127) / 128);
    af_table_kernel<W, 0><<<grid, 128, 0, st>>>(n, dJq, dhq, scale, lo);
    af_table_kernel<W, 1><<<grid, 128, 0, st>>>(n, dJq, dhq, scale, hia);
    af_table_kernel<W, 2><<<grid, 128, 0, st>>>(n, dJq, dhq, scale, hib);
    for (int b = 0; b < 6; ++b) {
        pos[0][b] = b;
        pos[1][b] = HiGeom<W>::gpos(7 + b);
        pos[2][b] = HiGeom<W>::gpos(5 + b);
    }
    return QMC_OK;
}

static int build_af_tables(int n, const double *dJq, const double *dhq, double scale, int4 *lo, int4 *hia, int4 *hib,
                           int (&pos)[3][6], cudaStream_t st) {
    switch (n - LO_BITS) {
        case 2: return build_af_w<2>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
        case 3: return build_af_w<3>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
        case 4: return build_af_w<4>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
        case 5: return build_af_w<5>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
        case 6: return build_af_w<6>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
        case 7: return build_af_w<7>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
        case 8: return build_af_w<8>(n, dJq, dhq, scale, lo, hia, hib, pos, st);
    }
    qmc_set_error("sweep path: n=%d outside [14, 20]", n);
    return QMC_ERR_UNSUPPORTED;
}

static int ensure_workspace(qmc_model *m, int64_t chains, cudaStream_t st) {
    if (!m->sweep) m->sweep = new SweepWorkspace();
    SweepWorkspace *w = m->sweep;
    const int n = m->n;
    const int64_t N = 1LL << n;
    if (!w->tables_ready) {
        // circuit couplings in qubit order: qubit a = spin n-1-a (quantum_mcmc_routines.py:75,87)
        w->Jq.assign((size_t)n * n, 0.0);
        w->hq.assign(n, 0.0);
        double dmax = 0.0;
        for (int a = 0; a < n; ++a) {
            w->hq[a] = m->host_ch[n - 1 - a];
            dmax += fabs(w->hq[a]);
            for (int b = 0; b < n; ++b) {
                w->Jq[(size_t)a * n + b] = m->host_cJ[(size_t)(n - 1 - a) * n + (n - 1 - b)];
                if (b > a) dmax += fabs(w->Jq[(size_t)a * n + b]);
            }
        }
        if (dmax < 1e-300) dmax = 1.0;
        w->k_fx = (int)floor(log2(2147483647.0 / (dmax * 1.0000001))) - 1;  // one bit of headroom for rounding
        if (w->k_fx > 30) w->k_fx = 30;
        w->fx_scale = ldexp(1.0, w->k_fx);
        QMC_CUDA_TRY(cudaMalloc(&w->dJq, sizeof(double) * (size_t)n * n));
        QMC_CUDA_TRY(cudaMalloc(&w->dhq, sizeof(double) * (size_t)n));
        QMC_CUDA_TRY(cudaMemcpy(w->dJq, w->Jq.data(), sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
        QMC_CUDA_TRY(cudaMemcpy(w->dhq, w->hq.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
        // LO layout (regs = qubits 0..5) is shared by both paths
        const int pos_lo[6] = {0, 1, 2, 3, 4, 5};
        gray_increments(w->Jq, n, w->fx_scale, pos_lo, w->dr_lo);
        if (n <= 20) {
            const size_t af_bytes = sizeof(int4) * 2 * (size_t)(N >> 6);
            QMC_CUDA_TRY(cudaMalloc(&w->af_lo, af_bytes));
            QMC_CUDA_TRY(cudaMalloc(&w->af_hia, af_bytes));
            QMC_CUDA_TRY(cudaMalloc(&w->af_hib, af_bytes));
            int pos[3][6];
            int rc = build_af_tables(n, w->dJq, w->dhq, w->fx_scale, w->af_lo, w->af_hia, w->af_hib, pos, st);
            if (rc != QMC_OK) return rc;
            m->launches += 3;
            QMC_CUDA_TRY(cudaGetLastError());
            gray_increments(w->Jq, n, w->fx_scale, pos[1], w->dr_hia);
            gray_increments(w->Jq, n, w->fx_scale, pos[2], w->dr_hib);
        }
        w->tables_ready = true;
    }
    if (chains > w->cap_chains) {
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(w->state); cudaFree(w->params); cudaFree(w->segsum); cudaFree(w->hist); cudaFree(w->order);
        w->state = nullptr; w->params = nullptr; w->segsum = nullptr; w->hist = nullptr; w->order = nullptr;
        w->cap_chains = 0;
        QMC_CUDA_TRY(cudaMalloc(&w->state, sizeof(float2) * (size_t)N * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->params, sizeof(ChainParams) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->segsum, sizeof(float) * (size_t)(N >> 6) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->hist, sizeof(int) * (size_t)(n + 1) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->order, sizeof(int) * (size_t)chains));
        w->cap_chains = chains;
    }
    return QMC_OK;
}

template <typename K>
static int set_smem_once(K kernel, bool &done, int bytes = TILE * 8) {
    if (!done) {
        QMC_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done = true;
    }
    return QMC_OK;
}

// Launch-time variants: fold = normalisation folded into the initial product state (phase factors of unit
// modulus); lo64 = LO tiles of 2^12 amplitudes on 64-thread CTAs (6 resident per SM instead of 3 of 2^13).
struct SweepTune {
    bool fold = false;
    bool lo64 = false;
};

template <int ROLE, int NTH, bool FOLD>
static int launch_lo_v(const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    static bool done = false;
    constexpr int SMEM = NTH * 64 * 8;
    int rc = set_smem_once(lo_sweep_kernel<ROLE, NTH, FOLD>, done, SMEM);
    if (rc != QMC_OK) return rc;
    const int64_t t = tiles * (128 / NTH);  // `tiles` counts 2^13-amplitude tiles
    lo_sweep_kernel<ROLE, NTH, FOLD><<<dim3((unsigned)t, (unsigned)count), NTH, SMEM, st>>>(a, order, off);
    return QMC_OK;
}
template <int ROLE>
static int launch_lo(const SweepTune &tu, const SweepArgs &a, const int *order, int off, int64_t tiles, int count,
                     cudaStream_t st) {
    constexpr bool PHASED = ROLE == LO_FIRST || ROLE == LO_MID;  // roles without a phase layer ignore `fold`
    if (ROLE == LO_SINGLE) return launch_lo_v<ROLE, 128, false>(a, order, off, tiles, count, st);
    if (tu.lo64) {
        if (PHASED && tu.fold) return launch_lo_v<ROLE, 64, PHASED>(a, order, off, tiles, count, st);
        return launch_lo_v<ROLE, 64, false>(a, order, off, tiles, count, st);
    }
    if (PHASED && tu.fold) return launch_lo_v<ROLE, 128, PHASED>(a, order, off, tiles, count, st);
    return launch_lo_v<ROLE, 128, false>(a, order, off, tiles, count, st);
}

template <int W, bool FIRST, bool FOLD>
static int launch_hi_w(const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    static bool done = false;
    int rc = set_smem_once(hi_sweep_kernel<W, FIRST, FOLD>, done);
    if (rc != QMC_OK) return rc;
    hi_sweep_kernel<W, FIRST, FOLD><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off);
    return QMC_OK;
}

template <bool FIRST, bool FOLD>
static int launch_hi_f(int n, const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    switch (n - LO_BITS) {
        case 2: return launch_hi_w<2, FIRST, FOLD>(a, order, off, tiles, count, st);
        case 3: return launch_hi_w<3, FIRST, FOLD>(a, order, off, tiles, count, st);
        case 4: return launch_hi_w<4, FIRST, FOLD>(a, order, off, tiles, count, st);
        case 5: return launch_hi_w<5, FIRST, FOLD>(a, order, off, tiles, count, st);
        case 6: return launch_hi_w<6, FIRST, FOLD>(a, order, off, tiles, count, st);
        case 7: return launch_hi_w<7, FIRST, FOLD>(a, order, off, tiles, count, st);
        case 8: return launch_hi_w<8, FIRST, FOLD>(a, order, off, tiles, count, st);
    }
    qmc_set_error("sweep path: n=%d outside [14, 20]", n);
    return QMC_ERR_UNSUPPORTED;
}
template <bool FIRST>
static int launch_hi(const SweepTune &tu, int n, const SweepArgs &a, const int *order, int off, int64_t tiles, int count,
                     cudaStream_t st) {
    return tu.fold ? launch_hi_f<FIRST, true>(n, a, order, off, tiles, count, st)
                   : launch_hi_f<FIRST, false>(n, a, order, off, tiles, count, st);
}

template <int ROLE, int ACT>
static int launch_hi6_ra(const SweepArgs &a, const Hi6Args &g, const int *order, int off, int64_t tiles, int count,
                         cudaStream_t st) {
    const size_t smem = ROLE == HI6_SINGLE ? 0 : sizeof(double) * (size_t)(a.n_total * a.n_total + a.n_total);
    hi6_sweep_kernel<ROLE, ACT><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off);
    return QMC_OK;
}
template <int ROLE>
static int launch_hi6(int act, const SweepArgs &a, const Hi6Args &g, const int *order, int off, int64_t tiles, int count,
                      cudaStream_t st) {
    switch (act) {
        case 1: return launch_hi6_ra<ROLE, 1>(a, g, order, off, tiles, count, st);
        case 2: return launch_hi6_ra<ROLE, 2>(a, g, order, off, tiles, count, st);
        case 3: return launch_hi6_ra<ROLE, 3>(a, g, order, off, tiles, count, st);
        case 4: return launch_hi6_ra<ROLE, 4>(a, g, order, off, tiles, count, st);
        case 5: return launch_hi6_ra<ROLE, 5>(a, g, order, off, tiles, count, st);
        case 6: return launch_hi6_ra<ROLE, 6>(a, g, order, off, tiles, count, st);
    }
    qmc_set_error("hi6 sweep: %d active qubits", act);
    return QMC_ERR_BADARG;
}

// Sweep list of one proposal with r mixer layers on the generic path (LO group + K >= 2 HI6 groups):
//   HI6_FIRST(g0); per middle layer: LO_SINGLE, HI6_SINGLE for the other groups, HI6_MID on the next
//   group (which carries its M into the following layer); last layer: HI6_SINGLE..., LO_FINAL.
static int run_generic_schedule(int n, int r, const SweepArgs &a, const std::vector<Hi6Args> &groups,
                                const std::vector<int> &acts, const int *order, int off, int count, int64_t tiles,
                                cudaStream_t st, int &launched) {
    int rc = QMC_OK;
    const int K = (int)groups.size();
    if (r <= 1) {
        ++launched;
        return launch_lo<LO_FINAL>(SweepTune(), a, order, off, tiles, count, st);
    }
    rc = launch_hi6<HI6_FIRST>(acts[0], a, groups[0], order, off, tiles, count, st);
    ++launched;
    int carried = 0;
    for (int layer = 2; layer <= r && rc == QMC_OK; ++layer) {
        const bool last = layer == r;
        const int mid = (carried + 1) % K;
        if (!last) {
            rc = launch_lo<LO_SINGLE>(SweepTune(), a, order, off, tiles, count, st);
            ++launched;
        }
        for (int k = 0; k < K && rc == QMC_OK; ++k) {
            if (k == carried || (!last && k == mid)) continue;
            rc = launch_hi6<HI6_SINGLE>(acts[k], a, groups[k], order, off, tiles, count, st);
            ++launched;
        }
        if (rc != QMC_OK) break;
        if (!last) {
            rc = launch_hi6<HI6_MID>(acts[mid], a, groups[mid], order, off, tiles, count, st);
            carried = mid;
        } else {
            rc = launch_lo<LO_FINAL>(SweepTune(), a, order, off, tiles, count, st);
        }
        ++launched;
    }
    return rc;
}

int qmc_sweep_run(qmc_model *m, int mode, int64_t n_chains, int64_t n_hops, const uint64_t *s_in, uint64_t chain_id0,
                  uint32_t hop0, double temperature, double g_lo, double g_hi, double dt, uint64_t seed, int dtype,
                  const double *gamma_arr, const int *r_arr, const double *u_arr, const int *group_arr,
                  uint64_t *proposals, uint8_t *accepted, uint64_t *final_state, int64_t *acc_count,
                  int64_t *hamming_hist, void *probs_out, void *amps_out, cudaStream_t st) {
    const int n = m->n;
    QMC_REQUIRE(dtype == QMC_DTYPE_F32, QMC_ERR_UNSUPPORTED,
                "complex128 statevectors are supported on the shared-memory path only (n <= %d); n=%d needs dtype fp32",
                QMC_SMEM_MAX_SPINS_F64, n);
    QMC_REQUIRE(n >= 14 && n <= QMC_MAX_SPINS, QMC_ERR_UNSUPPORTED, "HBM sweep path covers 14 <= n <= %d (n=%d)", QMC_MAX_SPINS, n);
    const bool generic = n > 20;  // three or more qubit groups
    QMC_REQUIRE(m->all_one_body, QMC_ERR_UNSUPPORTED,
                "n=%d > %d: only one-body X mixers have an HBM sweep path (k-body/custom mixers: shared-memory path)", n,
                QMC_SMEM_MAX_SPINS_F32);
    if (n_chains == 0) return QMC_OK;
    static bool smem_attr = false;
    if (!smem_attr) {
        QMC_CUDA_TRY(cudaFuncSetAttribute(sample_accept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 8));
        smem_attr = true;
    }
    const int64_t N = 1LL << n;
    const int64_t tiles = N >> TILE_BITS;
    // chains per batch: bounded by free memory (state dominates)
    size_t free_b = 0, total_b = 0;
    QMC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t per_chain = sizeof(float2) * (size_t)N + sizeof(float) * (size_t)(N >> 6) + sizeof(ChainParams) + 256;
    SweepWorkspace *w0 = m->sweep;
    const size_t have = w0 ? (size_t)w0->cap_chains * per_chain : 0;
    int64_t cap = (int64_t)(((double)(free_b + have) * 0.85) / (double)per_chain);
    if (cap > 65535) cap = 65535;  // gridDim.y
    QMC_REQUIRE(cap >= 1, QMC_ERR_CUDA, "not enough device memory for one 2^%d statevector", n);
    const int64_t batch = std::min<int64_t>(cap, n_chains);
    int rc = ensure_workspace(m, batch, st);
    if (rc != QMC_OK) return rc;
    SweepWorkspace *w = m->sweep;
    std::vector<Hi6Args> groups;
    std::vector<int> acts;
    if (generic) {
        const int hi_qubits = n - LO_BITS;
        const int K = (hi_qubits + 5) / 6;
        for (int k = 0; k < K; ++k) {
            Hi6Args g;
            const bool last = (k == K - 1);
            g.p = last ? n - 6 : LO_BITS + 6 * k;
            acts.push_back(last ? hi_qubits - 6 * (K - 1) : 6);
            const int pos[6] = {g.p, g.p + 1, g.p + 2, g.p + 3, g.p + 4, g.p + 5};
            gray_increments(w->Jq, n, w->fx_scale, pos, g.dr);
            g.Jq = w->dJq;
            g.hq = w->dhq;
            g.fx_scale = w->fx_scale;
            groups.push_back(g);
        }
    }

    // tuning knobs (defaults = measured best on B200; the env overrides exist for A/B runs and profiling)
    const char *pf_env = getenv("QUMCMC_SWEEP_PF");
    const int pf_dist = pf_env ? atoi(pf_env) : 222;
    const char *lo64_env = getenv("QUMCMC_SWEEP_LO64");
    const char *fold_env = getenv("QUMCMC_SWEEP_FOLD");
    const int r_cap = qmc_trotter_steps(11, dt);
    int r_max = r_cap;
    std::vector<int> r_host;
    std::vector<double> g_host;
    if (mode != QMC_MODE_CHAIN) {
        r_host.resize(n_chains);
        g_host.resize(n_chains);
        QMC_CUDA_TRY(cudaMemcpyAsync(r_host.data(), r_arr, sizeof(int) * n_chains, cudaMemcpyDeviceToHost, st));
        QMC_CUDA_TRY(cudaMemcpyAsync(g_host.data(), gamma_arr, sizeof(double) * n_chains, cudaMemcpyDeviceToHost, st));
        QMC_CUDA_TRY(cudaStreamSynchronize(st));
        r_max = 1;
        for (int64_t c = 0; c < n_chains; ++c) r_max = std::max(r_max, r_host[c]);
    }
    // Folding the normalisation (prod_q cos theta_q)^r into the initial product state is exact as long as that
    // factor stays far inside the fp32 range: bound it from below over every (gamma, r) this call can draw.
    SweepTune tune;
    tune.lo64 = lo64_env ? lo64_env[0] == '1' : false;
    {
        double log_min = 0.0;  // min over proposals of r * sum_q log|cos(gamma w_q dt)|
        auto log_scale = [&](double gamma, int grp) {
            double ls = 0.0;
            for (int k = m->group_off[grp]; k < m->group_off[grp + 1]; ++k) ls += log(std::max(1e-300, fabs(cos(gamma * m->weights[k] * dt))));
            return ls;
        };
        if (mode == QMC_MODE_CHAIN) {
            for (int grp = 0; grp < m->n_groups; ++grp) {
                bool monotone = true;  // |cos| is monotone on [g_lo w dt, g_hi w dt] when the angles stay below pi/2
                for (int k = m->group_off[grp]; k < m->group_off[grp + 1]; ++k)
                    monotone = monotone && fabs(std::max(fabs(g_lo), fabs(g_hi)) * m->weights[k] * dt) < 1.5;
                if (!monotone) { log_min = -1e300; break; }
                log_min = std::min(log_min, r_max * std::min(log_scale(g_lo, grp), log_scale(g_hi, grp)));
            }
        } else {
            std::vector<int> grp_host(n_chains, 0);
            if (group_arr) {
                QMC_CUDA_TRY(cudaMemcpyAsync(grp_host.data(), group_arr, sizeof(int) * n_chains, cudaMemcpyDeviceToHost, st));
                QMC_CUDA_TRY(cudaStreamSynchronize(st));
            }
            for (int64_t c = 0; c < n_chains; ++c)
                log_min = std::min(log_min, std::max(1, r_host[c]) * log_scale(g_host[c], grp_host[c]));
        }
        tune.fold = !generic && log_min > log(1e-24);
        if (fold_env) tune.fold = tune.fold && fold_env[0] == '1';
    }

    QMC_REQUIRE(r_max <= SCHED_RMAX, QMC_ERR_UNSUPPORTED, "delta_time=%g gives up to %d Trotter layers (> %d)", dt, r_max, SCHED_RMAX);
    std::vector<int> rhist(r_max + 2, 0);
    for (int64_t c0 = 0; c0 < n_chains; c0 += batch) {
        const int64_t nc = std::min<int64_t>(batch, n_chains - c0);
        HopArgs h;
        h.n = n; h.mode = mode; h.n_groups = m->n_groups; h.phase_active = m->phase_active; h.k_fx = w->k_fx; h.fold = tune.fold ? 1 : 0;
        h.n_chains = nc; h.n_hops = n_hops; h.hop = 0;
        h.chain_id0 = chain_id0 + (uint64_t)c0; h.seed = seed; h.hop0 = hop0;
        h.temperature = temperature; h.g_lo = g_lo; h.g_hi = g_hi; h.dt = dt; h.alpha = m->alpha;
        h.J = m->dJ; h.h = m->dh;
        h.group_off = m->d_group_off; h.masks = m->d_masks; h.weights = m->d_weights; h.group_cum = m->d_group_cum;
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
        SweepArgs a;
        a.n = n; a.n_total = n; a.gbits = 0ULL; a.state = w->state; a.params = w->params; a.af_lo = w->af_lo; a.af_hia = w->af_hia; a.af_hib = w->af_hib;
        memcpy(a.dr_lo, w->dr_lo, sizeof(a.dr_lo)); memcpy(a.dr_hia, w->dr_hia, sizeof(a.dr_hia)); memcpy(a.dr_hib, w->dr_hib, sizeof(a.dr_hib));
        a.segsum = w->segsum;
        a.pf_dist = pf_dist;
        a.probs_out = mode == QMC_MODE_PROBS ? reinterpret_cast<float *>(probs_out) : nullptr;
        a.amps_out = mode == QMC_MODE_PROBS ? reinterpret_cast<float2 *>(amps_out) : nullptr;
        // PROBS with amps only: still need a probs-free normalisation; handled by the normalise kernel

        const int gb = (int)((nc + 127) / 128);
        chains_init_kernel<<<gb, 128, 0, st>>>(h, w->params);
        m->launches++;
        const int64_t hops = mode == QMC_MODE_CHAIN ? n_hops : 1;
        for (int64_t hop = 0; hop < hops; ++hop) {
            h.hop = hop;
            hop_prepare_kernel<<<gb, 128, 0, st>>>(h, w->params);
            schedule_kernel<<<1, 1024, 0, st>>>(w->params, nc, r_max, w->order, generic ? 0 : 1);
            m->launches += 2;
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (m->profile) {
                QMC_CUDA_TRY(cudaEventCreate(&e0));
                QMC_CUDA_TRY(cudaEventCreate(&e1));
                QMC_CUDA_TRY(cudaEventRecord(e0, st));
            }
            // host replica of the draws: class sizes of the device-side schedule
            std::fill(rhist.begin(), rhist.end(), 0);
            double sum_r = 0.0;
            for (int64_t c = 0; c < nc; ++c) {
                int r;
                if (mode == QMC_MODE_CHAIN) {
                    const HopDraws d = qmc_hop_draws(seed, h.chain_id0 + (uint64_t)c, hop0 + (uint32_t)hop);
                    r = qmc_trotter_steps(d.t, dt);
                } else {
                    r = std::max(1, r_host[c0 + c]);
                }
                rhist[r]++;
                sum_r += r;
            }
            int launched = 0;
            if (generic) {
                // order[] is sorted by r descending: one schedule per distinct r over its contiguous range
                int off = 0;
                for (int r = r_max; r >= 1; --r) {
                    if (rhist[r] == 0) continue;
                    rc = run_generic_schedule(n, r, a, groups, acts, w->order, off, rhist[r], tiles, st, launched);
                    if (rc != QMC_OK) return rc;
                    off += rhist[r];
                }
            } else {
                int base[2] = {0, 0};  // start of the odd-r / even-r class in order[]
                for (int r = 1; r <= r_max; r += 2) base[1] += rhist[r];  // even class starts after all odd-r chains
                for (int step = 1; step <= r_max; ++step) {
                    const int p = step & 1;                 // chains with r == step (mod 2) sweep LO at this step
                    const int lo_base = p ? base[0] : base[1];
                    const int hi_base = p ? base[1] : base[0];
                    int lo_gt = 0, hi_gt = 0;
                    for (int r = step + 2; r <= r_max; r += 2) lo_gt += rhist[r];
                    for (int r = step + 1; r <= r_max; r += 2) hi_gt += rhist[r];
                    const int lo_eq = rhist[step];
                    if (lo_gt > 0) {
                        rc = step == 1 ? launch_lo<LO_FIRST>(tune, a, w->order, lo_base, tiles, lo_gt, st)
                                        : launch_lo<LO_MID>(tune, a, w->order, lo_base, tiles, lo_gt, st);
                        if (rc != QMC_OK) return rc;
                        ++launched;
                    }
                    if (lo_eq > 0) {
                        rc = launch_lo<LO_FINAL>(tune, a, w->order, lo_base + lo_gt, tiles, lo_eq, st);
                        if (rc != QMC_OK) return rc;
                        ++launched;
                    }
                    if (hi_gt > 0) {
                        rc = step == 1 ? launch_hi<true>(tune, n, a, w->order, hi_base, tiles, hi_gt, st)
                                        : launch_hi<false>(tune, n, a, w->order, hi_base, tiles, hi_gt, st);
                        if (rc != QMC_OK) return rc;
                        ++launched;
                    }
                }
            }
            m->launches += launched;
            if (m->profile) {
                QMC_CUDA_TRY(cudaEventRecord(e1, st));
                w->ev.push_back(e0);
                w->ev.push_back(e1);
                // algorithmic bytes of this hop: 2 * 2^n * 8 B * r per chain
                m->prof_bytes += 2.0 * (double)N * 8.0 * sum_r;
                m->prof_launches += launched;
            }
            if (mode == QMC_MODE_PROBS) {
                probs_normalize_kernel<<<148 * 4, 256, 0, st>>>(n, w->segsum, a.probs_out, a.amps_out);
                m->launches++;
            } else {
                sample_accept_kernel<<<(unsigned)nc, NT, TILE * 8, st>>>(a, h);
                m->launches++;
            }
            QMC_CUDA_TRY(cudaGetLastError());
        }
        if (mode == QMC_MODE_CHAIN) {
            chains_finish_kernel<<<gb, 128, 0, st>>>(h, w->params);
            m->launches++;
        }
        QMC_CUDA_TRY(cudaGetLastError());
        if (m->profile) {
            QMC_CUDA_TRY(cudaStreamSynchronize(st));
            for (size_t i = 0; i + 1 < w->ev.size(); i += 2) {
                float ms = 0.f;
                QMC_CUDA_TRY(cudaEventElapsedTime(&ms, w->ev[i], w->ev[i + 1]));
                m->prof_ms += ms;
                cudaEventDestroy(w->ev[i]);
                cudaEventDestroy(w->ev[i + 1]);
            }
            w->ev.clear();
        }
    }
    return QMC_OK;
}
