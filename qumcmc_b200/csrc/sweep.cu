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
//
// Round 2 additions (DESIGN sections 4.2, 4.2b, 4.3):
//   * L2-resident group schedule (group_sched.cuh): the chains of a group run all their sweeps back to back on side
//     streams, so a sweep finds its statevectors in the L2; bit-identical to the step-major schedule.
//   * XQ mode: mixers made of one- and two-qubit X strings are quadratic diagonals in the Hadamard basis,
//     U = H Dx H (P H Dx H)^(r-1) -- the same kernels with butterfly multipliers +-1 (Walsh-Hadamard transforms), HI
//     sweeps hosting Dx (phase tables built from the mixer's weights), LO sweeps hosting P: 2 r sweeps per proposal.
//   * XT mode: any other X-string mixer, Dx from the tabulated exponent E_x(x) (device Walsh-Hadamard transform of the
//     coefficient vector): one 4-byte lookup per amplitude in the HI sweeps.
//   * n >= 21: generic multi-group sweeps (LO + HI8 / HI6 register groups), sharded statevectors (qmc_shard_*).
//
// Layout in HBM: state[chain][2^n] interleaved complex64 (8 B/amplitude), index order.
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cstdio>

#include "common.cuh"
#include "group_sched.cuh"

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
    int cfx_x;   // XQ mode: rate of the mixer diagonal, round(-gamma dt / (2 pi) * 2^m')
    int shift_x; // k_fx_x + m' - 32
    float tq[QMC_MAX_SPINS];  // tan(theta_q), theta_q = gamma w_q dt
};

// Fast path only: tan(gamma w dt) of the chain at position i of order[] (one value per chain: every qubit carries
// the same mixer weight).  Read at a block-uniform index, so it lands in a uniform register and feeds FFMA2 as a
// scalar-broadcast operand (4.5 cycles/butterfly/sub-partition against 5.2 from a vector register).
constexpr int CTAN_MAX = 12288;
__constant__ float c_tan[CTAN_MAX];

struct SweepArgs {
    int n;                    // index qubits of the local statevector (addressing)
    int n_total;              // all qubits incl. global (sharded) ones; == n on one GPU
    unsigned long long gbits; // value of the global qubits, already shifted to bit n
    float2 *state;            // [chains][2^n]
    ChainParams *params;      // [chains]
    // on-the-fly phase exponent: D_fx(idx) = A + sum_b z_b F_b + R(j) per register layout;
    // af_*[tile*128 + thread] = {A, F0, F1, F2}, {F3, F4, F5, -}; dr_*[i] = Gray-order increments of R
    const int4 *af_lo, *af_hia, *af_hib, *af_hic;
    int dr_lo[64], dr_hia[64], dr_hib[64], dr_hic[64];
    int hiq;                  // n = 20: HI MID sweeps with amplitude pairs in the registers (hiq_sweep_kernel)
    float *segsum;            // [chains][2^n / 64]
    float *probs_out;         // PROBS mode: unnormalised |a|^2 (index order) or null
    float2 *amps_out;
    int pf_dist;              // > 0: each CTA L2-prefetches the tile of the CTA launched pf_dist later (in launch order)
    int pf_lines;             // strided tiles: 1 = per-thread prefetch.global.L2 of 128-byte lines, 0 = bulk prefetch per row
    float gen_scale;          // XQ mode: normalisation of the Walsh-Hadamard transforms of this launch (applied with the phase)
    const int *xq_dr;         // XQ mode: Gray increments of the mixer diagonal, [mixer group][layout: hia, hib, hic][64]
    int64_t xq_af_stride;     // XQ mode: int4 entries between the af_hi* tables of consecutive mixer groups
    const int *xt_ex;         // XT mode: tabulated mixer exponent E_x(x) in fixed point, [mixer group][2^n]
    // Fused exchange (sharded statevector): when xchg_peers != null the pass stores amplitude L of this rank to
    // xchg_peers[L >> xchg_shift][(L & low) | (xchg_rank << xchg_shift)] -- the all-to-all that swaps the top local
    // index bits with the rank bits, done by the sweep's own stores over NVLink peer mappings (no staging copy).
    float2 *const *xchg_peers;
    int xchg_shift, xchg_rank;
    int xchg_g, xchg_tile_bits;  // log2(ranks), log2(tiles of the local shard)
    int64_t xchg_idx0;           // chunked pass: index of the chunk's first amplitude inside the shard (else 0)
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
__device__ __forceinline__ A64 mul2(A64 a, A64 b) {
    A64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float norm2(A64 a) {
    const float2 f = upk(a);
    return f.x * f.x + f.y * f.y;
}

// Storage convention of the sweep path: amplitude a(idx) is held as S(idx) = (-i)^popcount(idx) * a(idx) -- in
// registers, shared memory and HBM alike.  In that basis exp(-i theta X_q) is a real rotation, so the tan-form
// butterfly a' = a - i t b ; b' = b - i t a  becomes, for the partner with qubit q clear (lo) / set (hi),
//     S_lo' = S_lo + t * S_hi        S_hi' = S_hi - t * S_lo
// identically on the real and imaginary halves: two packed FFMA2 whose multiplier is one scalar that is the same
// for every thread of the CTA (SASS: FFMA2 Rd, Rt.F32, Rw, Ru -- scalar broadcast, two 64-bit operand reads;
// measured 5.2 cycles/butterfly/sub-partition against 8.0 for per-thread (t, -t) pairs, benchmarks/micro/).
// Diagonal factors (the phase layer) commute with the basis change, and |S|^2 = |a|^2.
// reg bit b <-> qubit qpos(b).
template <int ACTIVE, bool FAST = false, typename QP>
__device__ __forceinline__ void bfly6(A64 (&v)[64], const ChainParams &cp, QP qpos, const float t_uniform = 0.f) {
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        if (!((ACTIVE >> b) & 1)) continue;
        const float t = FAST ? t_uniform : cp.tq[qpos(b)];
        const A64 TP = pk(t, t), TN = pk(-t, -t);
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            if (j & (1 << b)) continue;
            const A64 u = v[j], w = v[j | (1 << b)];
            v[j] = fma2(TP, w, u);
            v[j | (1 << b)] = fma2(TN, u, w);
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

// z * i^k for k mod 4 (conversion between stored S(idx) and a(idx) = i^popcount(idx) S(idx))
__device__ __forceinline__ float2 times_i_pow(float2 z, int k) {
    switch (k & 3) {
        case 0: return z;
        case 1: return make_float2(-z.y, z.x);
        case 2: return make_float2(-z.x, -z.y);
        default: return make_float2(z.y, -z.x);
    }
}

// Shared-memory swizzle of the LO transposition: XOR the 4 low index bits with bits 6..9.  A-layout accesses
// (lanes on bits 0..4, register bits 6..11 fixed per instruction) stay a permutation inside each aligned group of
// 16 slots; B-layout accesses (lane stride 64 slots) hit 16 distinct 8-byte slots per half-warp: both
// conflict-free with 64-bit accesses.  (128-bit B accesses would pin amplitude pairs (j, j+1) to aligned register
// quads, which costs ~400 MOVs around the qubit-0 butterflies.)
__host__ __device__ __forceinline__ int lo_swz(int local) { return local ^ ((local >> 6) & 15); }

// ------------------------------------------------------------------------------------------
// LO tile: 2^13 contiguous amplitudes, qubits 0..11 butterflied.
//   A layout: regs = local bits 6..11; lane = bits 0..4; warp = bits (5, 12)   (coalesced HBM access)
//   B layout: regs = local bits 0..5 (64 contiguous amplitudes); thread = bits 6..12
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int lo_base_a(int tid) {
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
// The swizzled slot of (thread-fixed bits | register bits) needs one XOR per value of (j & 15); the four
// registers sharing it sit at compile-time offsets, so each transposition costs 16 address computations.
__device__ __forceinline__ void lo_a_to_smem(const A64 (&v)[64], float2 *sm_, int tid) {
    A64 *sm = reinterpret_cast<A64 *>(sm_);
    const int base = lo_base_a(tid);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        A64 *p = sm + (base ^ k);
#pragma unroll
        for (int q = 0; q < 4; ++q) p[(k + 16 * q) << 6] = v[k + 16 * q];
    }
}
__device__ __forceinline__ void lo_smem_to_a(A64 (&v)[64], const float2 *sm_, int tid) {
    const A64 *sm = reinterpret_cast<const A64 *>(sm_);
    const int base = lo_base_a(tid);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const A64 *p = sm + (base ^ k);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[k + 16 * q] = p[(k + 16 * q) << 6];
    }
}
__device__ __forceinline__ void lo_b_to_smem(const A64 (&v)[64], float2 *sm_, int tid) {
    A64 *sm = reinterpret_cast<A64 *>(sm_) + (tid << 6);
    const int t4 = tid & 15;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        A64 *p = sm + (k ^ t4);
#pragma unroll
        for (int q = 0; q < 4; ++q) p[16 * q] = v[k + 16 * q];
    }
}
__device__ __forceinline__ void lo_smem_to_b(A64 (&v)[64], const float2 *sm_, int tid) {
    const A64 *sm = reinterpret_cast<const A64 *>(sm_) + (tid << 6);
    const int t4 = tid & 15;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const A64 *p = sm + (k ^ t4);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[k + 16 * q] = p[16 * q];
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
// Per-CTA table hs[0..63] of the Gray-order increments of R(j) in 32-bit turns (c/(2 pi) * dR, wraps mod 1 turn).
// cfx/shift are read straight from the chain's parameter record so that the table is complete at the barrier that
// ends the parameter staging.
__device__ __forceinline__ int turns32(int dfx, int cfx, int shift) {
    return (int)(unsigned int)(((long long)dfx * (long long)cfx) >> shift);
}
__device__ __forceinline__ void build_hs(int *hs, const int (&dr)[64], const ChainParams *cpg, int tid) {
    if (tid < 64) hs[tid] = turns32(dr[tid], cpg->cfx, cpg->shift);
}
// XQ mode, HI sweeps: increments of the mixer diagonal of the chain's mixer group (global memory), rate cfx_x
__device__ __forceinline__ void build_hs_x(int *hs, const int *__restrict__ drx, const ChainParams *cpg, int tid) {
    if (tid < 64) hs[tid] = turns32(__ldg(drx + tid), cpg->cfx_x, cpg->shift_x);
}

// K1 for the thread's 64 amplitudes.  The phase angle is accumulated directly in 32-bit integer turns along the
// Gray-code order: t(j') = t(j) +/- G_b + H_i with G_b = turns(2 F_b) (per thread) and H_i = turns(dR_i) (per
// chain, shared memory) -- one IADD3 per amplitude, wrap-around = mod 2 pi for free, rounding <= 64 * 2^-33 turn.
// FOLD: the per-layer normalisation prod cos(theta_q) has been folded into the initial product state
// (cp.scale = scale^r), so the factor has unit modulus; otherwise it is scale * exp(i phi).
// No branch on cp.active: an absent phase layer has cfx = 0, i.e. every factor is exactly (1, 0).  (A branch here
// forces the register allocator to re-home all 64 amplitudes at the join: ~450 MOV / XOR-swap instructions.)
template <bool FOLD>
__device__ __forceinline__ void phase_gray(A64 (&v)[64], const PhaseAF af, const int *hs_all, const int dr0, const int cfx,
                                           const int shift, const float sc);
template <bool FOLD>
__device__ __forceinline__ void phase_gray(A64 (&v)[64], const PhaseAF af, const int *hs_all, const int dr0,
                                           const ChainParams &cp) {
    phase_gray<FOLD>(v, af, hs_all, dr0, cp.cfx, cp.shift, cp.scale);
}
template <bool FOLD>
__device__ __forceinline__ void phase_gray(A64 (&v)[64], const PhaseAF af, const int *hs_all, const int dr0, const int cfx,
                                           const int shift, const float sc) {
    const int F[6] = {af.a0.y, af.a0.z, af.a0.w, af.a1.x, af.a1.y, af.a1.z};
    int G[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) G[b] = turns32(2 * F[b], cfx, shift);
    int t = turns32(af.a0.x + F[0] + F[1] + F[2] + F[3] + F[4] + F[5] + dr0, cfx, shift);
    const int4 *hs = reinterpret_cast<const int4 *>(hs_all);
    int4 hn = hs[0];
#pragma unroll
    for (int i4 = 0; i4 < 16; ++i4) {
        const int4 h4 = hn;  // increments H_{4 i4 .. 4 i4 + 3}; the step i -> i+1 adds H_{i+1}
        if (i4 < 15) hn = hs[i4 + 1];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = 4 * i4 + k;
            const int j = i ^ (i >> 1);
            float sn, cs;
            __sincosf((float)t * 1.4629180792671596e-09f, &sn, &cs);  // 2 pi / 2^32
            if (!FOLD) {
                sn *= sc;
                cs *= sc;
            }
            const float2 a = upk(v[j]);
            v[j] = pk(a.x * cs - a.y * sn, a.y * cs + a.x * sn);
            if (i < 63) {
                const int i1 = i + 1;
                const int b = (i1 & 1) ? 0 : (i1 & 2) ? 1 : (i1 & 4) ? 2 : (i1 & 8) ? 3 : (i1 & 16) ? 4 : 5;
                const int jn = i1 ^ (i1 >> 1);
                const int h = (k == 0) ? h4.y : (k == 1) ? h4.z : (k == 2) ? h4.w : hn.x;
                t = ((jn >> b) & 1) ? t - G[b] + h : t + G[b] + h;
            }
        }
    }
}

// XT mode: phase factor of the mixer diagonal from the tabulated exponent (any X-string mixer: E_x is not a quadratic form
// when terms of weight >= 3 are present).  ex points at the thread's base index, off(j) is the index offset of register j.
template <typename OFF>
__device__ __forceinline__ void phase_lookup(A64 (&v)[64], const int *__restrict__ ex, OFF off, const int cfx, const int shift,
                                             const float sc) {
#pragma unroll
    for (int j0 = 0; j0 < 64; j0 += 16) {
        int e[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = __ldg(ex + off(j0 + k));  // 16 loads in flight
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float sn, cs;
            __sincosf((float)turns32(e[k], cfx, shift) * 1.4629180792671596e-09f, &sn, &cs);  // 2 pi / 2^32
            sn *= sc;
            cs *= sc;
            const float2 a = upk(v[j0 + k]);
            v[j0 + k] = pk(a.x * cs - a.y * sn, a.y * cs + a.x * sn);
        }
    }
}

// M|s0> restricted to this thread's 64 amplitudes; reg bit b <-> global qubit qpos[b].
//   a(idx) = scale * prod_{q in idx^s0} (-i t_q)   =>   S(idx) = (-i)^popcount(s0) * (-1)^popcount(idx & ~s0) * |..|
// i.e. real up to the global factor (-i)^popcount(s0).
template <typename QP>
__device__ __forceinline__ void init_product(A64 (&v)[64], uint64_t base_idx, int n, const ChainParams &cp, QP qpos) {
    uint64_t regmask = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) regmask |= 1ULL << qpos(b);
    const uint64_t diff = (base_idx ^ cp.cur) & ~regmask;
    float mag0 = (__popcll(base_idx & ~cp.cur & ~regmask) & 1) ? -cp.scale : cp.scale;
    for (int q = 0; q < n; ++q)
        if ((diff >> q) & 1ULL) mag0 *= cp.tq[q];
    const float2 g = times_i_pow(make_float2(1.f, 0.f), -(int)__popcll(cp.cur));  // (-i)^popcount(s0)
    // Register qubit b contributes e_b(j_b): it differs from s0 -> factor t_q, and a minus sign when the bit is set in idx
    // but not in s0.  The 64 products are built by doubling: level b turns every v[j], j < 2^b, into the pair
    // (v[j] e_b(0), v[j] e_b(1)) -- 126 packed multiplies instead of 6 selects + multiplies per amplitude.
    v[0] = pk(mag0 * g.x, mag0 * g.y);
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const float t = cp.tq[qpos(b)];
        const bool sb = (cp.cur >> qpos(b)) & 1ULL;
        const float e0 = sb ? t : 1.0f, e1 = sb ? 1.0f : -t;
        const A64 p0 = pk(e0, e0), p1 = pk(e1, e1);
#pragma unroll
        for (int j = 0; j < (1 << b); ++j) {
            const A64 x = v[j];
            v[j | (1 << b)] = mul2(x, p1);
            v[j] = mul2(x, p0);
        }
    }
}

// XQ mode: the x-basis image of |s0> restricted to this thread's 64 amplitudes: mag * (-1)^popcount(idx & ~s0), real.
template <typename QP>
__device__ __forceinline__ void init_hadamard(A64 (&v)[64], uint64_t base_idx, const ChainParams &cp, QP qpos, float mag) {
    // The transform into the x basis is the butterfly with multiplier +1, B = [[1, 1], [-1, 1]] per qubit (= Z H sqrt 2), so
    // the start state is B^(x)n |s0>: component x carries (-1)^popcount(x & ~s0) (the convention of init_product).
    uint64_t regmask = 0;
    int sreg = 0;  // bits of ~s0 at the six register positions
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        regmask |= 1ULL << qpos(b);
        sreg |= (int)((~cp.cur >> qpos(b)) & 1ULL) << b;
    }
    const int p0 = __popcll(base_idx & ~cp.cur & ~regmask) & 1;
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = pk(((p0 ^ __popc(j & sreg)) & 1) ? -mag : mag, 0.f);
}

struct LoQ0 { __device__ __forceinline__ int operator()(int b) const { return b; } };       // B layout: bits 0..5
struct LoQ6 { __device__ __forceinline__ int operator()(int b) const { return 6 + b; } };   // A layout: bits 6..11

// Final LO computation for (chain, tile): leaves the final amplitudes of the tile in B layout.
// Used by the last sweep (segment sums) and by the sampler (recompute of the selected tile).
// A LO tile is NTH * 64 contiguous amplitudes: 2^13 for 128 threads, 2^12 for 64 threads (bit 12 is a pure
// batch bit of the LO butterflies, so a 128-thread tile is two independent 64-thread tiles).
// R1: proposal with a single mixer layer -- the final state is the analytic product state (no sweep ran before).
template <int NTH, bool FAST, bool R1>
__device__ __forceinline__ void lo_final_tile(A64 (&v)[64], float2 *sm, const float2 *__restrict__ state_chain,
                                              int64_t tile, int tid, int n, const ChainParams &cp,
                                              unsigned long long gb = 0ULL, const float tu = 0.f) {
    constexpr int TB = NTH == 128 ? 13 : 12;
    if (R1) {
        init_product(v, gb | ((uint64_t)tile << TB) | ((uint64_t)tid << 6), n, cp, LoQ0());
        return;
    }
    lo_load_a(v, state_chain + (tile << TB), tid);
    bfly6<63, FAST>(v, cp, LoQ6(), tu);
    lo_a_to_smem(v, sm, tid);
    __syncthreads();
    lo_smem_to_b(v, sm, tid);
    bfly6<63, FAST>(v, cp, LoQ0(), tu);
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
__host__ __device__ __forceinline__ int hi_loc_a(int tid, int j) { return (tid & 31) | ((tid >> 5) << 5) | (j << 7); }
__host__ __device__ __forceinline__ int hi_loc_b(int tid, int j) { return (tid & 31) | (j << 5) | ((tid >> 5) << 11); }

// Tile of CTA b.  Fused exchange: the top g bits of the tile index select the destination rank of a lower-group
// pass, so the CTAs are renumbered with those bits fastest (and rotated by the rank): at any moment the stores of a
// GPU fan out over all peers instead of every GPU writing to the same peer at the same time.
__device__ __forceinline__ int64_t sweep_tile(const SweepArgs &a, unsigned b) {
    if (!a.xchg_peers) return (int64_t)b;
    if (a.xchg_tile_bits < a.xchg_g) return (int64_t)b;  // a chunk inside one destination: the caller orders the chunks
    const unsigned W1 = (1u << a.xchg_g) - 1u;
    return ((int64_t)((b + (unsigned)a.xchg_rank) & W1) << (a.xchg_tile_bits - a.xchg_g)) | (int64_t)(b >> a.xchg_g);
}

__device__ __forceinline__ A64 *xchg_addr(const SweepArgs &a, int64_t local_idx) {
    local_idx += a.xchg_idx0;
    const int64_t low = local_idx & (((int64_t)1 << a.xchg_shift) - 1);
    return reinterpret_cast<A64 *>(a.xchg_peers[local_idx >> a.xchg_shift]) + (low | ((int64_t)a.xchg_rank << a.xchg_shift));
}

// L2 prefetch of the tile a later CTA of the same launch will read (CTAs start in linear blockIdx order,
// x fastest): with 3 resident CTAs per SM the CTA `pf_dist` ahead starts roughly when this one retires,
// so its loads hit L2 instead of waiting on HBM with nothing else to overlap.
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Per-lane variant for strided tiles: cp.async.bulk.prefetch takes a uniform address, so issuing one per row from
// every lane makes the compiler serialise the warp (a 32-trip R2UR / UBLKPF / BRA.U.ANY loop, ~19 % of the stall
// samples of the HI sweep); prefetch.global.L2 is an ordinary per-lane LSU instruction.
__device__ __forceinline__ void l2_prefetch_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory"); }
// A strided tile = `rows` rows of row_bytes contiguous bytes (64 KiB in total = 512 lines): 4 lines per thread.
__device__ __forceinline__ void l2_prefetch_rows(const char *src, int64_t row_stride_bytes, int row_bytes, int tid) {
    const int lines_per_row = row_bytes >> 7;
#pragma unroll
    for (int i = 0; i < (TILE * 8 / 128) / NT; ++i) {
        const int l = tid + NT * i;
        const int row = l / lines_per_row, within = l - row * lines_per_row;
        l2_prefetch_line(src + row * row_stride_bytes + ((int64_t)within << 7));
    }
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

enum { LO_FIRST = 0, LO_MID = 1, LO_FINAL = 2, LO_SINGLE = 3, LO_FINAL1 = 4 };  // SINGLE: M_lo only (3+ qubit groups); FINAL1: r == 1

// FAST = normalisation folded into the initial product state + one mixer multiplier per chain from c_tan[]
// XQ (general two-body path, roles MID / FINAL only): the butterflies are Walsh-Hadamard transforms -- the same FFMA2 with
// multiplier -1 (back from the x basis: lo - hi, hi + lo) before the phase layer and +1 (into the x basis) after it -- and
// the amplitudes are plain (no (-i)^popcount storage basis); the phase factor carries the normalisation a.gen_scale.
template <int ROLE, int NTH, bool FAST, bool XQ = false>
__global__ void __launch_bounds__(NTH, 384 / NTH) lo_sweep_kernel(const SweepArgs a, const int *__restrict__ order, const int list_off, const int ct_off) {
    constexpr int TB = NTH == 128 ? 13 : 12;  // tile bits
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[64];
    const int tid = threadIdx.x;
    const int64_t tile = sweep_tile(a, blockIdx.x);
    const int64_t chain = order[list_off + blockIdx.y];
    // ct_off == list_off: passed separately so that this index stays pure uniform-datapath arithmetic (LDCU -> UR)
    const float tu = XQ ? -1.f : (FAST ? c_tan[ct_off + blockIdx.y] : 0.f);
    const float tu2 = XQ ? 1.f : tu;  // multiplier of the butterflies after the phase layer
    const int n = a.n;
    float2 *state = a.state + (chain << n);
    A64 v[64];
    if (ROLE == LO_MID) lo_load_a(v, state + (tile << TB), tid);  // in flight while the parameters are staged
    if (ROLE != LO_FIRST && ROLE != LO_FINAL1 && tid == 0) {
        int64_t pc, pt;
        if (pf_target(a, order, list_off, pc, pt)) l2_prefetch_bulk(a.state + (pc << n) + (pt << TB), (1u << TB) * 8);
    }
    if (ROLE == LO_FIRST || ROLE == LO_MID) build_hs(hs_s, a.dr_lo, a.params + chain, tid);
    stage_params<NTH>(cp_s, a.params + chain, tid);
    const ChainParams &cp = cp_s;
    if (ROLE == LO_FINAL || ROLE == LO_FINAL1) {
        lo_final_tile<NTH, FAST, ROLE == LO_FINAL1>(v, sm, state, tile, tid, a.n_total, cp, a.gbits, tu);
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
            const int kb = __popcll(a.gbits | ((unsigned long long)tile << TB) | ((unsigned)tid << 6));
            for (int j = 0; j < 64; ++j) ao[j] = XQ ? upk(v[j]) : times_i_pow(upk(v[j]), kb + __builtin_popcount(j));  // a = i^popcount S
        }
        return;
    }
    if (ROLE == LO_SINGLE) {
        lo_load_a(v, state + (tile << TB), tid);
        bfly6<63, FAST>(v, cp, LoQ6(), tu);
        lo_a_to_smem(v, sm, tid);
        __syncthreads();
        lo_smem_to_b(v, sm, tid);
        bfly6<63, FAST>(v, cp, LoQ0(), tu);
        lo_b_to_smem(v, sm, tid);  // each thread rewrites exactly the slots it read
        __syncthreads();
        lo_smem_to_a(v, sm, tid);
        if (a.xchg_peers) {  // the whole tile shares its top index bits: one destination rank per CTA
            A64 *dst = xchg_addr(a, tile << TB) + lo_base_a(tid);
#pragma unroll
            for (int j = 0; j < 64; ++j) dst[j << 6] = v[j];
        } else {
            lo_store_a(v, state + (tile << TB), tid);
        }
        return;
    }
    const PhaseAF af = load_af(a.af_lo, tile * NTH + tid);  // issued before the amplitude loads are consumed
    if (ROLE == LO_FIRST) {
        init_product(v, ((uint64_t)tile << TB) | ((uint64_t)tid << 6), n, cp, LoQ0());
    } else {
        bfly6<63, FAST>(v, cp, LoQ6(), tu);
        lo_a_to_smem(v, sm, tid);
        __syncthreads();
        lo_smem_to_b(v, sm, tid);
        bfly6<63, FAST>(v, cp, LoQ0(), tu);
    }
    if (XQ) phase_gray<false>(v, af, hs_s, a.dr_lo[0], cp.cfx, cp.shift, a.gen_scale);
    else phase_gray<FAST>(v, af, hs_s, a.dr_lo[0], cp);
    bfly6<63, FAST>(v, cp, LoQ0(), tu2);
    lo_b_to_smem(v, sm, tid);  // each thread rewrites exactly the slots it read: no barrier needed before
    __syncthreads();
    lo_smem_to_a(v, sm, tid);
    bfly6<63, FAST>(v, cp, LoQ6(), tu2);
    lo_store_a(v, state + (tile << TB), tid);
}

#ifdef QMC_PROBES
// Diagnostics only (make PROBES=1 / benchmarks/probe_sweep.cu; never in the shipped library): the LO MID sweep
// with parts removed.  MODE bits: 1 = HBM loads + stores, 2 = butterflies, 4 = phase layer, 8 = transpositions.
// Without bit 1 the store is predicated on a value the compiler cannot prove, so the arithmetic stays.
template <int MODE, int MINB, int TPC = 1>
__global__ void __launch_bounds__(NT, MINB) lo_probe_kernel(const SweepArgs a, const int *__restrict__ order, const int list_off) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[64];
    const int tid = threadIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    const int n = a.n;
    float2 *state = a.state + (chain << n);
    build_hs(hs_s, a.dr_lo, a.params + chain, tid);
    stage_params<NT>(cp_s, a.params + chain, tid);
    const ChainParams &cp = cp_s;
#pragma unroll 1
    for (int it = 0; it < TPC; ++it) {
        const int64_t tile = (int64_t)blockIdx.x * TPC + it;
        A64 v[64];
        if (MODE & 1) {
            lo_load_a(v, state + (tile << 13), tid);
            if (tid == 0) {
                int64_t pc, pt;
                if (pf_target(a, order, list_off, pc, pt)) l2_prefetch_bulk(a.state + (pc << n) + (pt << 13), TILE * 8);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = pk((float)(tid + j), 1.0f);
        }
        const PhaseAF af = load_af(a.af_lo, tile * NT + tid);
        if (MODE & 2) bfly6<63>(v, cp, LoQ6());
        if (MODE & 8) {
            if (TPC > 1) __syncthreads();
            lo_a_to_smem(v, sm, tid);
            __syncthreads();
            lo_smem_to_b(v, sm, tid);
        }
        if (MODE & 2) bfly6<63>(v, cp, LoQ0());
        if (MODE & 4) phase_gray<true>(v, af, hs_s, a.dr_lo[0], cp);
        if (MODE & 2) bfly6<63>(v, cp, LoQ0());
        if (MODE & 8) {
            lo_b_to_smem(v, sm, tid);
            __syncthreads();
            lo_smem_to_a(v, sm, tid);
        }
        if (MODE & 2) bfly6<63>(v, cp, LoQ6());
        if (MODE & 1) {
            lo_store_a(v, state + (tile << 13), tid);
        } else {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) acc += norm2(v[j]);
            if (acc == 1.2345e-30f) lo_store_a(v, state + (tile << 13), tid);
        }
    }
}
#endif

// XQ: the phase layer of an HI sweep is the MIXER diagonal exp(-i gamma dt E_x(x)) (tables a.af_hi* / a.dr_hi* built from
// the mixer's pair weights, rate cfx_x), between a transform into the x basis (+1) and back (-1); FIRST starts from H|s0>.
// XT (with XQ): the same for ANY X-string mixer -- the diagonal comes from the tabulated exponent a.xt_ex (one 4-byte load per
// amplitude, L2-resident: 4 MiB per mixer group at n = 20) instead of the quadratic-form tables.
template <int W, bool FIRST, bool FAST, bool XQ = false, bool XT = false>
__global__ void __launch_bounds__(NT, 3) hi_sweep_kernel(const SweepArgs a, const int *__restrict__ order, const int list_off, const int ct_off) {
    using G = HiGeom<W>;
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[64];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    // ct_off == list_off: passed separately so that this index stays pure uniform-datapath arithmetic (LDCU -> UR)
    const float tu = XQ ? 1.f : (FAST ? c_tan[ct_off + blockIdx.y] : 0.f);
    const float tu2 = XQ ? -1.f : tu;
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
            if (a.pf_lines && ROW_BYTES >= 128) l2_prefetch_rows(reinterpret_cast<const char *>(src), (int64_t)8 << 12, ROW_BYTES, tid);
            else for (int r = tid; r < ROWS; r += NT) l2_prefetch_bulk(src + ((int64_t)r << 12), ROW_BYTES);
        }
    }
    // XQ: tables of the chain's mixer group (IncoherentMixerSum picks one sub-mixer per proposal)
    const int grp = XQ ? a.params[chain].group : 0;
    const int *drx = XQ && !XT ? a.xq_dr + ((int64_t)grp * 3 + (G::two_rounds ? 1 : 0)) * 64 : nullptr;
    const int *ext = XT ? a.xt_ex + ((int64_t)grp << n) : nullptr;
    if (!XT) {
        if (XQ) build_hs_x(hs_s, drx, a.params + chain, tid);
        else build_hs(hs_s, G::two_rounds ? a.dr_hib : a.dr_hia, a.params + chain, tid);
    }
    stage_params(cp_s, a.params + chain, tid);
    const ChainParams &cp = cp_s;
    PhaseAF af;
    if (!XT) af = load_af((G::two_rounds ? a.af_hib : a.af_hia) + (XQ ? grp * a.xq_af_stride : 0), tile * NT + tid);
    const float mag0 = XQ ? exp2f(-0.5f * (float)n) : 0.f;
    if (G::two_rounds) {
        const int64_t base_b = hi_base_b<W>(tid, tile);
        if (FIRST) {
            if (XQ) init_hadamard(v, (uint64_t)base_b, cp, HiQB<W>(), mag0);
            else init_product(v, (uint64_t)base_b, n, cp, HiQB<W>());
        } else {
            bfly6<G::active_a, FAST>(v, cp, HiQA<W>(), tu);
#pragma unroll
            for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(sm)[hi_loc_a(tid, j)] = v[j];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = reinterpret_cast<const A64 *>(sm)[hi_loc_b(tid, j)];
            bfly6<G::active_b, FAST>(v, cp, HiQB<W>(), tu);
            __syncthreads();
        }
        if (XT) phase_lookup(v, ext + base_b, [](int j) { return G::off_b(j); }, cp.cfx_x, cp.shift_x, a.gen_scale);
        else if (XQ) phase_gray<false>(v, af, hs_s, __ldg(drx), cp.cfx_x, cp.shift_x, a.gen_scale);
        else phase_gray<FAST>(v, af, hs_s, a.dr_hib[0], cp);
        bfly6<G::active_b, FAST>(v, cp, HiQB<W>(), tu2);
#pragma unroll
        for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(sm)[hi_loc_b(tid, j)] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = reinterpret_cast<const A64 *>(sm)[hi_loc_a(tid, j)];
        bfly6<G::active_a, FAST>(v, cp, HiQA<W>(), tu2);
#pragma unroll
        for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(state)[base_a + G::off_a(j)] = v[j];
    } else {
        if (FIRST) {
            if (XQ) init_hadamard(v, (uint64_t)base_a, cp, HiQA<W>(), mag0);
            else init_product(v, (uint64_t)base_a, n, cp, HiQA<W>());
        } else {
            bfly6<G::active_a, FAST>(v, cp, HiQA<W>(), tu);
        }
        if (XT) phase_lookup(v, ext + base_a, [](int j) { return G::off_a(j); }, cp.cfx_x, cp.shift_x, a.gen_scale);
        else if (XQ) phase_gray<false>(v, af, hs_s, __ldg(drx), cp.cfx_x, cp.shift_x, a.gen_scale);
        else phase_gray<FAST>(v, af, hs_s, a.dr_hia[0], cp);
        bfly6<G::active_a, FAST>(v, cp, HiQA<W>(), tu2);
#pragma unroll
        for (int j = 0; j < 64; ++j) reinterpret_cast<A64 *>(state)[base_a + G::off_a(j)] = v[j];
    }
}

// ------------------------------------------------------------------------------------------
// HI MID sweep for W = 8 (n = 20) with amplitude PAIRS in the registers: index bit 0 (a LO qubit, never mixed here) is
// register bit 0, so every HBM access is a 128-bit LDG / STG (32 instead of 64 per thread and direction) and the
// transpositions move 128-bit shared-memory words (half the LDS / STS count; the HI sweep is MIO-throttled).
// Tile: local bits 0..4 <-> qubits 0..4, local 5..12 <-> qubits 12..19 (as HiGeom<8>).
//   A'' layout: regs = local {0, 8..12}; lanes = local 1..5; warps = local (6, 7)      active: qubits 15..19
//   B'' layout: regs = local {0, 5, 6, 7, 8, 9}; lanes = local {1..4, 10}; warps = local (11, 12)
//                                                                                      active: qubits 12..14; phase
// ------------------------------------------------------------------------------------------
typedef ulonglong2 P128;
__device__ __forceinline__ int64_t hiq_base_a(int tid, int64_t tile) {
    const int lane = tid & 31, w = tid >> 5;
    return (int64_t)((lane & 15) << 1) | ((int64_t)(lane >> 4) << 12) | ((int64_t)(w & 1) << 13) | ((int64_t)(w >> 1) << 14) |
           (tile << 5);
}
__host__ __device__ __forceinline__ int64_t hiq_base_b(int tid, int64_t tile) {
    const int lane = tid & 31, w = tid >> 5;
    return (int64_t)((lane & 15) << 1) | ((int64_t)(lane >> 4) << 17) | ((int64_t)(w & 1) << 18) | ((int64_t)(w >> 1) << 19) |
           (tile << 5);
}
__host__ __device__ __forceinline__ int hiq_pos_b(int b) { return b == 0 ? 0 : 11 + b; }  // register bit -> qubit, B''
struct HiqQA { __device__ __forceinline__ int operator()(int b) const { return b == 0 ? 0 : 14 + b; } };
struct HiqQB { __device__ __forceinline__ int operator()(int b) const { return hiq_pos_b(b); } };
// shared-memory slot (in 128-bit words) of register pair jp
__host__ __device__ __forceinline__ int hiq_loc_a(int tid, int jp) { return (tid & 31) | ((tid >> 5) << 5) | (jp << 7); }
__host__ __device__ __forceinline__ int hiq_loc_b(int tid, int jp) {
    return (tid & 15) | ((jp & 7) << 4) | ((jp >> 3) << 7) | (((tid >> 4) & 1) << 9) | ((tid >> 5) << 10);
}

template <bool FAST, bool XQ = false>
__global__ void __launch_bounds__(NT, 3) hiq_sweep_kernel(const SweepArgs a, const int *__restrict__ order, const int list_off,
                                                          const int ct_off) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[64];
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t chain = order[list_off + blockIdx.y];
    const float tu = XQ ? 1.f : (FAST ? c_tan[ct_off + blockIdx.y] : 0.f);
    const float tu2 = XQ ? -1.f : tu;
    const int n = a.n;
    A64 *state = reinterpret_cast<A64 *>(a.state + (chain << n)) + hiq_base_a(tid, tile);
    P128 *smq = reinterpret_cast<P128 *>(sm);
    A64 v[64];
#pragma unroll
    for (int jp = 0; jp < 32; ++jp) {  // in flight while the parameters are staged
        const P128 x = *reinterpret_cast<const P128 *>(state + ((int64_t)jp << 15));
        v[2 * jp] = x.x;
        v[2 * jp + 1] = x.y;
    }
    {
        int64_t pc, pt;
        if (pf_target(a, order, list_off, pc, pt))
            l2_prefetch_rows(reinterpret_cast<const char *>(a.state + (pc << n) + (pt << 5)), (int64_t)8 << 12, 256, tid);
    }
    const int grp = XQ ? a.params[chain].group : 0;
    const int *drx = XQ ? a.xq_dr + ((int64_t)grp * 3 + 2) * 64 : nullptr;
    if (XQ) build_hs_x(hs_s, drx, a.params + chain, tid);
    else build_hs(hs_s, a.dr_hic, a.params + chain, tid);
    stage_params(cp_s, a.params + chain, tid);
    const ChainParams &cp = cp_s;
    const PhaseAF af = load_af(a.af_hic + (XQ ? grp * a.xq_af_stride : 0), tile * NT + tid);
    bfly6<62, FAST>(v, cp, HiqQA(), tu);
#pragma unroll
    for (int jp = 0; jp < 32; ++jp) smq[hiq_loc_a(tid, jp)] = make_ulonglong2(v[2 * jp], v[2 * jp + 1]);
    __syncthreads();
#pragma unroll
    for (int jp = 0; jp < 32; ++jp) {
        const P128 x = smq[hiq_loc_b(tid, jp)];
        v[2 * jp] = x.x;
        v[2 * jp + 1] = x.y;
    }
    bfly6<14, FAST>(v, cp, HiqQB(), tu);
    if (XQ) phase_gray<false>(v, af, hs_s, __ldg(drx), cp.cfx_x, cp.shift_x, a.gen_scale);
    else phase_gray<FAST>(v, af, hs_s, a.dr_hic[0], cp);
    bfly6<14, FAST>(v, cp, HiqQB(), tu2);
#pragma unroll
    for (int jp = 0; jp < 32; ++jp) smq[hiq_loc_b(tid, jp)] = make_ulonglong2(v[2 * jp], v[2 * jp + 1]);  // own slots
    __syncthreads();
#pragma unroll
    for (int jp = 0; jp < 32; ++jp) {
        const P128 x = smq[hiq_loc_a(tid, jp)];
        v[2 * jp] = x.x;
        v[2 * jp + 1] = x.y;
    }
    bfly6<62, FAST>(v, cp, HiqQA(), tu2);
#pragma unroll
    for (int jp = 0; jp < 32; ++jp)
        *reinterpret_cast<P128 *>(state + ((int64_t)jp << 15)) = make_ulonglong2(v[2 * jp], v[2 * jp + 1]);
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
    const double *cta_tab;  // [tiles][AF_CTA_STRIDE] tile-uniform phase exponent fields (af_cta_table_kernel)
    const double *thr_tab;  // [NT][AF_THR_STRIDE] thread-level sums
};

// Phase exponent fields of a thread whose six register qubits are p..p+5 (no per-thread table, any n).  The fixed
// qubits split into the seven that differ between the threads of a CTA (positions tpos(0..6): lane and warp bits)
// and the rest, which are uniform over the CTA (= the tile):
//     A   = A_C + sum_a z_a g_a + sum_{a<a'} Jq[a][a'] z_a z_a'       F_b = f_b + sum_a Jq[a][p+b] z_a
// with A_C (pair + field sum over the uniform qubits), g_a = hq[a] + sum_c Jq[a][c] z_c and f_b = hq[p+b] +
// sum_c Jq[c][p+b] z_c.  The 14 per-tile values depend on the couplings and the geometry only, so they are tabulated
// once per (group, layout) by af_cta_table_kernel (16 doubles per tile, 1/512 of the state's size).
constexpr int AF_SCR = 16;
constexpr int AF_CTA_STRIDE = 16;  // doubles per tile: g_a[7], f_b[6], A_C, pad
constexpr int AF_THR_STRIDE = 8;   // doubles per thread: P (pair sum over the thread bits), Q_b[6], pad
// The thread-level sums depend on the thread index only: P(tid) = sum_{a<a'} Jq[a][a'] z_a z_a' and Q_b(tid) =
// sum_a Jq[a][p+b] z_a are a static [128][8] table per group, so a thread adds 7 + 6 fp64 terms:
//     A = A_C + P(tid) + sum_a z_a g_a        F_b = f_b + Q_b(tid)
// af_fetch (before the staging barrier): tile row -> scr, the thread's own table row -> its slot of thr_s.
__device__ __forceinline__ void af_fetch(double *scr, double *thr_s, const double *__restrict__ cta_tab,
                                         const double *__restrict__ thr_tab, int64_t tile, int tid) {
    if (tid < 14) scr[tid] = __ldg(cta_tab + tile * AF_CTA_STRIDE + tid);
    const double2 *src = reinterpret_cast<const double2 *>(thr_tab + tid * AF_THR_STRIDE);
    double2 *dst = reinterpret_cast<double2 *>(thr_s + tid * AF_THR_STRIDE);
#pragma unroll
    for (int i = 0; i < AF_THR_STRIDE / 2; ++i) dst[i] = __ldg(src + i);
}
__device__ __forceinline__ PhaseAF af_thread(const double *scr, const double *thr_s, int tid, double fx_scale, int active) {
    PhaseAF af;
    af.a0 = make_int4(0, 0, 0, 0);
    af.a1 = af.a0;
    if (active) {
        const double2 *row = reinterpret_cast<const double2 *>(thr_s + tid * AF_THR_STRIDE);
        const double2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3];  // P Q0 | Q1 Q2 | Q3 Q4 | Q5 -
        double A = scr[13] + r0.x;
#pragma unroll
        for (int a = 0; a < 7; ++a) A += ((tid >> a) & 1) ? -scr[a] : scr[a];
        af.a0 = make_int4((int)rint(A * fx_scale), (int)rint((scr[7] + r0.y) * fx_scale), (int)rint((scr[8] + r1.x) * fx_scale),
                          (int)rint((scr[9] + r1.y) * fx_scale));
        af.a1 = make_int4((int)rint((scr[10] + r2.x) * fx_scale), (int)rint((scr[11] + r2.y) * fx_scale),
                          (int)rint((scr[12] + r3.x) * fx_scale), 0);
    }
    return af;
}

// Index bits that are uniform over the CTA of `tile`: kind 6 = HI6 geometry (thread bits 0..6, register window
// [p, p+6)), kind 8 = HI8 geometry (thread bits 0..4 and p+6, p+7; window [p, p+8)).
__host__ __device__ __forceinline__ unsigned long long hi_tile_base(int kind, int p, int64_t tile) {
    const int low_bits = kind == 6 ? p - 7 : p - 5;
    const unsigned long long t = (unsigned long long)tile;
    return ((t & ((1ULL << low_bits) - 1ULL)) << (kind == 6 ? 7 : 5)) | ((t >> low_bits) << (kind == 6 ? p + 6 : p + 8));
}
__host__ __device__ __forceinline__ int hi_thread_pos(int kind, int p, int a) { return (kind == 6 || a < 5) ? a : p + 1 + a; }

// One thread per tile: the 14 tile-uniform fields (fp64, fixed summation order).
__global__ void af_cta_table_kernel(int kind, int p, int nt, unsigned long long gbits, const double *__restrict__ Jq,
                                    const double *__restrict__ hq, int64_t tiles, double *__restrict__ out) {
    const int64_t tile = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (tile >= tiles) return;
    const unsigned long long full = hi_tile_base(kind, p, tile) | gbits;
    unsigned long long tmask = 0ULL;
    for (int a = 0; a < 7; ++a) tmask |= 1ULL << hi_thread_pos(kind, p, a);
    const unsigned long long cmask = (nt >= 64 ? ~0ULL : ((1ULL << nt) - 1ULL)) & ~(63ULL << p) & ~tmask;
    double AC = 0.0;
    for (int q = 0; q < nt; ++q) {
        if (!((cmask >> q) & 1ULL)) continue;
        double acc = hq[q];
        for (int q2 = q + 1; q2 < nt; ++q2)
            if ((cmask >> q2) & 1ULL) acc += ((full >> q2) & 1ULL) ? -Jq[q * nt + q2] : Jq[q * nt + q2];
        AC += ((full >> q) & 1ULL) ? -acc : acc;
    }
    double *o = out + tile * AF_CTA_STRIDE;
    for (int k = 0; k < 13; ++k) {
        const int tq = k < 7 ? hi_thread_pos(kind, p, k) : p + (k - 7);
        double val = hq[tq];
        for (int c = 0; c < nt; ++c) {
            if (!((cmask >> c) & 1ULL)) continue;
            const double j = c < tq ? Jq[c * nt + tq] : Jq[tq * nt + c];
            val += ((full >> c) & 1ULL) ? -j : j;
        }
        o[k] = val;
    }
    o[13] = AC;
    o[14] = 0.0;
    o[15] = 0.0;
}

// PRE = active (top) register qubits of the half-layer before the phase, when it differs from ACT: the sharded path
// mixes only the g qubits that the preceding all-to-all brought in, then all six after the phase.
template <int ROLE, int ACT, int PRE = ACT>
__global__ void __launch_bounds__(NT, 3) hi6_sweep_kernel(const SweepArgs a, const Hi6Args g, const int *__restrict__ order,
                                                          const int list_off) {
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[64];
    __shared__ double af_scr[AF_SCR];
    __shared__ __align__(16) double thr_s[ROLE != HI6_SINGLE ? NT * AF_THR_STRIDE : 2];
    const int tid = threadIdx.x;
    const int n = a.n;
    const int64_t tile = sweep_tile(a, blockIdx.x);
    const int64_t chain = order[list_off + blockIdx.y];
    if (ROLE != HI6_SINGLE) build_hs(hs_s, g.dr, a.params + chain, tid);
    if (ROLE != HI6_SINGLE) af_fetch(af_scr, thr_s, g.cta_tab, g.thr_tab, tile, tid);
    stage_params(cp_s, a.params + chain, tid);  // barrier inside
    const ChainParams &cp = cp_s;
    const int p = g.p;
    constexpr int ACTIVE = 63 & ~((1 << (6 - ACT)) - 1);
    constexpr int ACTIVE_PRE = 63 & ~((1 << (6 - PRE)) - 1);
    // index bits: lane 0..4, warp bits 5,6, tile-low bits 7..p-1, register bits p..p+5, tile-high bits p+6..n-1
    const int low_bits = p - 7;
    const int64_t base = (int64_t)(tid & 31) | ((int64_t)(tid >> 5) << 5) | ((tile & (((int64_t)1 << low_bits) - 1)) << 7) |
                         ((tile >> low_bits) << (p + 6));
    A64 *state = reinterpret_cast<A64 *>(a.state + (chain << n)) + base;
    const unsigned long long full = (unsigned long long)base | a.gbits;  // incl. the global (sharded) qubits
    const int nt = a.n_total;
    const int64_t stride = (int64_t)1 << p;
    const HiP qp{p};
    A64 v[64];
    if (ROLE != HI6_FIRST) {  // in flight while the phase exponent fields are evaluated
        const A64 *src = state;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            v[j] = *src;
            src += stride;
        }
    }
    // phase exponent fields (fp64): A = sum_{a<b fixed} Jq z z + sum_{a fixed} hq z,  F_b = hq[p+b] + sum_{a fixed} Jq[a][p+b] z_a
    PhaseAF af;
    af.a0 = make_int4(0, 0, 0, 0);
    af.a1 = af.a0;
    if (ROLE != HI6_SINGLE) af = af_thread(af_scr, thr_s, tid, g.fx_scale, cp.active);
    if (ROLE == HI6_FIRST) {
        init_product(v, full, nt, cp, qp);
    } else {
        bfly6<ACTIVE_PRE>(v, cp, qp);
    }
    if (ROLE != HI6_SINGLE) {
        phase_gray<false>(v, af, hs_s, g.dr[0], cp);
        bfly6<ACTIVE>(v, cp, qp);
    }
    if (a.xchg_peers) {  // register qubits may include the swapped positions: destination resolved per amplitude
#pragma unroll
        for (int j = 0; j < 64; ++j) *xchg_addr(a, base + j * stride) = v[j];
    } else {
        A64 *dst = state;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            *dst = v[j];
            dst += stride;
        }
    }
}

// Exchange pass as a PERSISTENT kernel on a bounded number of CTAs (sharded statevector, overlap mode): the SINGLE pass
// of an HI6 group whose stores go to the peers' buffers over NVLink.  A full-grid launch of this NVLink-bound pass
// occupies every CTA slot of the GPU with CTAs that wait on NVLink back-pressure, so the HBM-bound local passes of the
// next chunk (another stream) cannot run beside it; gridDim.x << tiles keeps most slots free for them while a few dozen
// CTAs keep enough stores in flight to fill the NVLink egress.  CTA b walks tiles b, b + gridDim.x, ...
template <int ACT>
__global__ void __launch_bounds__(NT, 3) hi6_xchg_kernel(const SweepArgs a, const Hi6Args g, const int *__restrict__ order,
                                                         const int list_off, const unsigned n_tiles) {
    __shared__ ChainParams cp_s;
    const int tid = threadIdx.x;
    const int n = a.n;
    const int64_t chain = order[list_off + blockIdx.y];
    stage_params(cp_s, a.params + chain, tid);  // barrier inside
    const ChainParams &cp = cp_s;
    const int p = g.p;
    constexpr int ACTIVE = 63 & ~((1 << (6 - ACT)) - 1);
    const int low_bits = p - 7;
    const int64_t stride = (int64_t)1 << p;
    const HiP qp{p};
    const A64 *state = reinterpret_cast<const A64 *>(a.state + (chain << n));
#pragma unroll 1
    for (unsigned b = blockIdx.x; b < n_tiles; b += gridDim.x) {
        const int64_t tile = sweep_tile(a, b);
        const int64_t base = (int64_t)(tid & 31) | ((int64_t)(tid >> 5) << 5) | ((tile & (((int64_t)1 << low_bits) - 1)) << 7) |
                             ((tile >> low_bits) << (p + 6));
        A64 v[64];
        const A64 *src = state + base;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            v[j] = *src;
            src += stride;
        }
        bfly6<ACTIVE>(v, cp, qp);
#pragma unroll
        for (int j = 0; j < 64; ++j) *xchg_addr(a, base + j * stride) = v[j];
    }
}

// ------------------------------------------------------------------------------------------
// HI8 sweeps: eight qubits [p, p+8) per pass (tile = index bits [0,5) u [p,p+8), 2^13 amplitudes, the geometry of
// hi_sweep_kernel<8> with a run-time window position), so that a large statevector needs ceil((n-12)/8) + 1 qubit
// groups instead of ceil((n-12)/6) + 1.
//   A' layout: regs = qubits p+2..p+7; lane = bits 0..4; warp = qubits (p, p+1)        (HBM loads / stores)
//   B' layout: regs = qubits p..p+5;   lane = bits 0..4; warp = qubits (p+6, p+7)      (phase + qubits p, p+1)
// ACT = 7 / 8 active (top) qubits of the window; PRE = qubits mixed before the phase of a MID pass (ACT, or the 1..6
// top qubits an exchange brought in).  Phase exponent fields as in the HI6 kernels (af_thread, B' layout).
// ------------------------------------------------------------------------------------------
enum { HI8_FIRST = 0, HI8_MID = 1, HI8_SINGLE = 2 };
struct Hi8QA { int p; __device__ __forceinline__ int operator()(int b) const { return p + 2 + b; } };
struct Hi8QB { int p; __device__ __forceinline__ int operator()(int b) const { return p + b; } };

template <int ROLE, int ACT, int PRE>
__global__ void __launch_bounds__(NT, 3) hi8_sweep_kernel(const SweepArgs a, const Hi6Args g, const int *__restrict__ order,
                                                          const int list_off) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    __shared__ __align__(16) int hs_s[64];
    __shared__ double af_scr[AF_SCR];
    __shared__ __align__(16) double thr_s[ROLE != HI8_SINGLE ? NT * AF_THR_STRIDE : 2];
    static_assert(ACT == 7 || ACT == 8, "HI8 windows carry 7 or 8 active qubits");
    constexpr int MASK_B = ACT == 8 ? 3 : 2;
    constexpr int PRE_A = PRE >= 6 ? 63 : (63 & ~((1 << (6 - PRE)) - 1));
    constexpr int PRE_B = PRE == 8 ? 3 : (PRE == 7 ? 2 : 0);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n = a.n, nt = a.n_total, p = g.p;
    const int64_t tile = sweep_tile(a, blockIdx.x);
    const int64_t chain = order[list_off + blockIdx.y];
    const int low_bits = p - 5;
    const int64_t base_tile = ((tile & (((int64_t)1 << low_bits) - 1)) << 5) | ((tile >> low_bits) << (p + 8));
    const int64_t base_a = (int64_t)lane | ((int64_t)(w & 1) << p) | ((int64_t)(w >> 1) << (p + 1)) | base_tile;
    const int64_t base_b = (int64_t)lane | ((int64_t)(w & 1) << (p + 6)) | ((int64_t)(w >> 1) << (p + 7)) | base_tile;
    const int64_t stride_a = (int64_t)1 << (p + 2), stride_b = (int64_t)1 << p;
    A64 *state = reinterpret_cast<A64 *>(a.state + (chain << n));
    A64 *smv = reinterpret_cast<A64 *>(sm);
    A64 v[64];
    if (ROLE != HI8_FIRST) {  // in flight while the parameters and the phase exponent fields are prepared
        const A64 *src = state + base_a;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            v[j] = *src;
            src += stride_a;
        }
    }
    if (ROLE != HI8_FIRST) {  // L2 prefetch of the tile a later CTA will read: 2^8 rows of 32 contiguous amplitudes
        int64_t pc, pt;
        if (pf_target(a, order, list_off, pc, pt)) {
            const float2 *src = a.state + (pc << n) + (int64_t)hi_tile_base(8, p, pt);
            if (a.pf_lines) l2_prefetch_rows(reinterpret_cast<const char *>(src), (int64_t)8 << p, 256, tid);
            else for (int r = tid; r < 256; r += NT) l2_prefetch_bulk(src + ((int64_t)r << p), 256);
        }
    }
    if (ROLE != HI8_SINGLE) build_hs(hs_s, g.dr, a.params + chain, tid);
    if (ROLE != HI8_SINGLE) af_fetch(af_scr, thr_s, g.cta_tab, g.thr_tab, tile, tid);
    stage_params(cp_s, a.params + chain, tid);  // barrier inside
    const ChainParams &cp = cp_s;
    const Hi8QA qa{p};
    const Hi8QB qb{p};
    const unsigned long long full_b = (unsigned long long)base_b | a.gbits;  // incl. the global (sharded) qubits
    PhaseAF af;
    af.a0 = make_int4(0, 0, 0, 0);
    af.a1 = af.a0;
    if (ROLE != HI8_SINGLE) af = af_thread(af_scr, thr_s, tid, g.fx_scale, cp.active);
    if (ROLE == HI8_FIRST) {
        init_product(v, full_b, nt, cp, qb);
    } else {
        bfly6 < ROLE == HI8_SINGLE ? 63 : PRE_A > (v, cp, qa);
#pragma unroll
        for (int j = 0; j < 64; ++j) smv[hi_loc_a(tid, j)] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = smv[hi_loc_b(tid, j)];
        bfly6 < ROLE == HI8_SINGLE ? MASK_B : PRE_B > (v, cp, qb);
    }
    if (ROLE == HI8_SINGLE) {  // B' keeps the lanes on the 32 lowest amplitudes: store from here
        if (a.xchg_peers) {
#pragma unroll
            for (int j = 0; j < 64; ++j) *xchg_addr(a, base_b + j * stride_b) = v[j];
        } else {
            A64 *dst = state + base_b;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                *dst = v[j];
                dst += stride_b;
            }
        }
        return;
    }
    phase_gray<false>(v, af, hs_s, g.dr[0], cp);
    bfly6<MASK_B>(v, cp, qb);
#pragma unroll
    for (int j = 0; j < 64; ++j) smv[hi_loc_b(tid, j)] = v[j];  // each thread rewrites exactly the slots it read
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = smv[hi_loc_a(tid, j)];
    bfly6<63>(v, cp, qa);
    if (a.xchg_peers) {
#pragma unroll
        for (int j = 0; j < 64; ++j) *xchg_addr(a, base_a + j * stride_a) = v[j];
    } else {
        A64 *dst = state + base_a;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            *dst = v[j];
            dst += stride_a;
        }
    }
}

// Device-side schedule: order[] = chains with odd r sorted by r descending, then chains with even r
// sorted by r descending.  The host replays the same Philox draws to know the class sizes.
constexpr int SCHED_RMAX = 2047;
__global__ void __launch_bounds__(1024) schedule_kernel(const ChainParams *__restrict__ params, int64_t n_chains,
                                                        int r_cap, int *__restrict__ order, int by_parity,
                                                        float *__restrict__ tan_sorted) {
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
        const int pos = start[r] + atomicAdd(&hist[r], 1);
        order[pos] = (int)c;
        if (tan_sorted) tan_sorted[pos] = params[c].tq[0];  // fast path: the same multiplier on every qubit
    }
}

// ------------------------------------------------------------------------------------------
// per-hop bookkeeping kernels
// ------------------------------------------------------------------------------------------
struct HopArgs {
    int n, mode, n_groups, phase_active, k_fx;
    int fold;  // 1: cp.scale = (prod cos)^r, applied once in the initial product state
    int xq;    // XQ mode: quadratic X mixer hosted by the sweep kernels; cp.r counts half-layers (2 r)
    int k_fx_x;
    int64_t n_chains, n_hops, hop;
    uint64_t chain_id0, seed;
    uint32_t hop0;
    double temperature, g_lo, g_hi, dt, alpha;
    const double *J, *h;
    const int *group_off;
    const uint64_t *masks;
    const double *weights, *group_cum;
    const double *qweight;  // [n_groups][n] summed mixer weight per qubit
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
    unsigned long long *visit;  // visit histogram accumulator or null
};

__device__ void fill_proposal_params(ChainParams &cp, const HopArgs &h, double gamma, int r, int grp) {
    cp.r = r < 1 ? 1 : r;
    cp.group = grp;
    cp.cfx_x = 0;
    cp.shift_x = 0;
    for (int q = 0; q < QMC_MAX_SPINS; ++q) cp.tq[q] = 0.f;
    if (h.xq) {
        // U = H Dx H (P H Dx H)^(r-1): 2 r passes -- HI FIRST, (LO MID, HI MID) x (r - 1), LO FINAL -- scheduled as a
        // proposal of 2 r sweeps.  tq = -1 is the multiplier of the last transform (x basis -> computational basis on the LO
        // qubits), which the sampler repeats for the selected tile.
        const int real_r = cp.r;
        cp.r = 2 * real_r;
        for (int q = 0; q < h.n; ++q) cp.tq[q] = -1.f;
        cp.scale = 1.f;
        const double cp_ = (1.0 - gamma) * h.alpha * h.dt * 0.15915494309189533577;  // c / (2 pi)
        cp.active = h.phase_active && real_r > 1 && cp_ != 0.0;
        int m = 0;
        if (cp.active) {
            m = (int)floor(log2(2147483647.0 / fabs(cp_)));
            m = m > 40 ? 40 : (m < 0 ? 0 : m);
        }
        cp.cfx = cp.active ? (int)rint(ldexp(cp_, m)) : 0;
        cp.shift = cp.active ? h.k_fx + m - 32 : 0;
        const double cx = -gamma * h.dt * 0.15915494309189533577;  // mixer diagonal: exp(-i gamma dt E_x)
        if (cx != 0.0) {
            int mx = (int)floor(log2(2147483647.0 / fabs(cx)));
            mx = mx > 40 ? 40 : (mx < 0 ? 0 : mx);
            cp.cfx_x = (int)rint(ldexp(cx, mx));
            cp.shift_x = h.k_fx_x + mx - 32;
        }
        return;
    }
    double scale = 1.0;
    for (int q = 0; q < h.n; ++q) {  // theta_q = gamma * (summed weight of the group's terms on qubit q) * dt
        const double wq = h.qweight[grp * h.n + q];
        if (wq == 0.0) continue;
        double s, c;
        sincos(gamma * wq * h.dt, &s, &c);
        cp.tq[q] = (float)(s / c);
        scale *= c;
    }
    cp.scale = h.fold ? (float)pow(scale, (double)cp.r) : (float)scale;
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
    if (h.mode == QMC_MODE_CHAIN && h.hop0 == 0) qmc_visit(h.visit, cur);  // entry 0 of markov_chain
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
// Block-cooperative inverse-CDF step: the entry of arr[0..cnt) at which the running sum (fp64, index order) passes the
// target -- `target` itself, or target * total when `relative` (then target is the uniform in [0,1)).  Thread t owns
// the contiguous chunk [t * per, (t + 1) * per); results in *s_sel (entry) and *s_rem (target minus the sum before
// that entry).  Ends with a barrier.
template <typename T>
__device__ __forceinline__ void block_pick(const T *__restrict__ arr, int64_t cnt, double target, bool relative, int tid,
                                           double *s_incl, double *s_warp, long long *s_sel, double *s_rem) {
    const int64_t per = (cnt + NT - 1) / NT;
    const int64_t i0 = tid * per < cnt ? tid * per : cnt, i1 = i0 + per < cnt ? i0 + per : cnt;
    const int last_owner = (int)((cnt - 1) / per);
    double local = 0.0;
    for (int64_t i = i0; i < i1; ++i) local += (double)arr[i];
    double incl = local;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double x = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += x;
    }
    __syncthreads();  // previous users of the scratch are done
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    incl += base;
    s_incl[tid] = incl;
    __syncthreads();
    if (relative) target *= s_incl[NT - 1];
    const double lo = tid == 0 ? 0.0 : s_incl[tid - 1];
    if (tid <= last_owner && target >= lo && (target < incl || tid == last_owner)) {
        double c = lo;  // cumulative sum before entry i
        int64_t sel = -1;
        double before = 0.0;
        for (int64_t i = i0; i < i1; ++i) {
            const double x = (double)arr[i];
            if (target < c + x) {
                sel = i;
                before = c;
                break;
            }
            c += x;
        }
        if (sel < 0) {  // rounding tail: last entry of this thread
            sel = i1 - 1;
            before = c - (double)arr[sel];
        }
        *s_sel = sel;
        *s_rem = target - before;
    }
    __syncthreads();
}

// Large n: sums of COARSE consecutive segment sums, so that the sampler's scan touches 2^(n-18) + 2^12 entries
// instead of 2^(n-6).  grid = (2^(n-6) / COARSE, chains).
constexpr int COARSE_BITS = 12, COARSE = 1 << COARSE_BITS;
__global__ void __launch_bounds__(256) coarse_sums_kernel(const float *__restrict__ segsum, int n, double *__restrict__ coarse) {
    __shared__ double red[8];
    const int64_t chain = blockIdx.y;
    const float4 *src = reinterpret_cast<const float4 *>(segsum + (chain << (n - 6)) + ((int64_t)blockIdx.x << COARSE_BITS));
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < COARSE / 4 / 256; ++i) {
        const float4 x = src[i * 256 + threadIdx.x];
        s += ((double)x.x + (double)x.y) + ((double)x.z + (double)x.w);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        coarse[(chain << (n - 6 - COARSE_BITS)) + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(NT) sample_accept_kernel(const SweepArgs a, const HopArgs h, const double *__restrict__ coarse) {
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
    const float *seg = a.segsum + (chain << (n - 6));
    if (coarse) {  // two levels: coarse block, then the segment inside it
        block_pick(coarse + (chain << (n - 6 - COARSE_BITS)), nseg >> COARSE_BITS, cp.u_meas, true, tid, s_incl, s_warp,
                   &s_seg, &s_rem);
        const int64_t blk = s_seg;
        const double rem = s_rem;
        block_pick(seg + (blk << COARSE_BITS), (int64_t)COARSE, rem, false, tid, s_incl, s_warp, &s_seg, &s_rem);
        if (tid == 0) s_seg += blk << COARSE_BITS;
    } else {
        block_pick(seg, nseg, cp.u_meas, true, tid, s_incl, s_warp, &s_seg, &s_rem);
    }
    __syncthreads();
    const int64_t sg = s_seg;
    const int64_t tile = sg / NT;
    const int owner = (int)(sg % NT);
    A64 v[64];
    if (cp.r == 1) lo_final_tile<NT, false, true>(v, sm, a.state + ((int64_t)chain << n), tile, tid, n, cp);
    else lo_final_tile<NT, false, false>(v, sm, a.state + ((int64_t)chain << n), tile, tid, n, cp);
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
            qmc_visit(h.visit, acc ? sp : cp.cur);
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
        for (int b = 0; b < 6; ++b)
            p[b] = LAYOUT == 0 ? b : (LAYOUT == 1 ? HiGeom<W>::gpos(7 + b) : (LAYOUT == 2 ? HiGeom<W>::gpos(5 + b) : hiq_pos_b(b)));
        if (LAYOUT == 0) base = ((uint64_t)tile << TILE_BITS) | ((uint64_t)tid << 6);
        else if (LAYOUT == 1) base = (uint64_t)hi_base_a<W>(tid, tile);
        else if (LAYOUT == 2) base = (uint64_t)hi_base_b<W>(tid, tile);
        else base = (uint64_t)hiq_base_b(tid, tile);
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

// Static-analysis hook (tests/test_smem_layouts.py, CPU): the shared-memory slot (in units of the access width) that
// thread `tid` touches for register `j` in one of the transposition layouts of the sweep kernels, evaluated by the very
// functions the kernels call.  layout: 0 LO A, 1 LO B (64-bit slots, j < 64); 2 HI A', 3 HI B' (64-bit, j < 64);
// 4 HIQ A'', 5 HIQ B'' (128-bit slots, j < 32).  Returns -1 for an unknown layout.
// Static-analysis hook (tests/test_group_sched.py, CPU): the launch list of one hop under the L2-resident group schedule.
// out = 4 ints per launch (kind, first position in order[], chains, side stream); returns the number of launches.
extern "C" int qmc_debug_group_plan(int n_chains, int r_max, const int *rhist, int group, int streams, int *out, int cap) {
    if (n_chains < 0 || r_max < 1 || !rhist || group < 1 || streams < 1) return -1;
    std::vector<int> rh(rhist, rhist + r_max + 1), rpos;
    rh.push_back(0);
    int total = 0;
    for (int r = 1; r <= r_max; ++r) total += rh[r];
    if (total != n_chains || rh[0] != 0) return -1;
    std::vector<QmcGroupLaunch> plan;
    qmc_group_plan(n_chains, r_max, rh, group, streams, rpos, plan);
    for (size_t i = 0; i < plan.size() && (int)i < cap; ++i) {
        out[4 * i] = plan[i].kind;
        out[4 * i + 1] = plan[i].off;
        out[4 * i + 2] = plan[i].count;
        out[4 * i + 3] = plan[i].stream;
    }
    return (int)plan.size();
}

extern "C" int qmc_debug_smem_slot(int layout, int tid, int j) {
    switch (layout) {
        case 0: return (lo_base_a(tid) ^ (j & 15)) + (j << 6);                                  // lo_a_to_smem / lo_smem_to_a
        case 1: return (tid << 6) + (((j & 15) ^ (tid & 15)) + (j & ~15));                  // lo_b_to_smem / lo_smem_to_b
        case 2: return hi_loc_a(tid, j);
        case 3: return hi_loc_b(tid, j);
        case 4: return hiq_loc_a(tid, j);
        case 5: return hiq_loc_b(tid, j);
    }
    return -1;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// One register group above the LO qubits: kind 6 = HI6 window [p, p+6) (register-only pass), kind 8 = HI8 window
// [p, p+8) (one shared-memory transposition each way); the top `act` qubits of the window are the group's own.
struct HiGroup {
    int kind = 6, act = 6;
    Hi6Args args;
};

// Geometry for `nl` local qubits, top group first: full HI8 groups stack down from the top; the remainder (1..7
// qubits) is the lowest group -- an HI8 window with 7 active qubits or an HI6 window with 1..6.
// Test hook: QUMCMC_HI_GROUPS="kind:act,kind:act,..." (top group first, kind 6 | 8) replaces the default cover, so that
// schedules with three or four register groups and stacked HI8 windows can be checked against the oracle at sizes it
// still finishes (tests/test_gpu_sweep.py, tests/test_gpu_shard.py).  A window [p, p + kind) owns its top `act` qubits;
// the ones below belong to the next group (or to the LO qubits) and only ride along.  An override that does not
// cover exactly the nl - 12 upper qubits, or that a kernel cannot run, is an error.
static int hi_group_geometry(int nl, std::vector<int> &kind, std::vector<int> &pos, std::vector<int> &act) {
    if (const char *e = getenv("QUMCMC_HI_GROUPS")) {
        if (*e) {
            int top = nl;
            const char *c = e;
            while (*c) {
                char *end = nullptr;
                const int k = (int)strtol(c, &end, 10);
                QMC_REQUIRE(end != c && *end == ':', QMC_ERR_BADARG, "QUMCMC_HI_GROUPS=%s: expected kind:act[,kind:act...]", e);
                c = end + 1;
                const int a = (int)strtol(c, &end, 10);
                QMC_REQUIRE(end != c && (*end == ',' || *end == 0), QMC_ERR_BADARG, "QUMCMC_HI_GROUPS=%s: expected kind:act[,kind:act...]", e);
                c = *end ? end + 1 : end;
                const int p = top - k;
                const bool ok6 = k == 6 && a >= 1 && a <= 6 && p >= 7;
                const bool ok8 = k == 8 && (a == 7 || a == 8) && p >= 5;
                QMC_REQUIRE(ok6 || ok8, QMC_ERR_BADARG, "QUMCMC_HI_GROUPS=%s: group %d:%d at position %d has no kernel", e, k, a, p);
                kind.push_back(k); pos.push_back(p); act.push_back(a);
                top -= a;
            }
            QMC_REQUIRE(top == LO_BITS, QMC_ERR_BADARG, "QUMCMC_HI_GROUPS=%s covers qubits %d..%d of %d local qubits (LO group = 0..%d)", e, top,
                        nl - 1, nl, LO_BITS - 1);
            return QMC_OK;
        }
    }
    const int hi = nl - LO_BITS, full8 = hi / 8, rem = hi % 8;
    for (int k = 0; k < full8; ++k) {
        kind.push_back(8); pos.push_back(nl - 8 * (k + 1)); act.push_back(8);
    }
    if (rem == 7) {
        kind.push_back(8); pos.push_back(LO_BITS + 7 - 8); act.push_back(7);
    } else if (rem > 0) {
        kind.push_back(6); pos.push_back(LO_BITS + rem - 6); act.push_back(rem);
    }
    return QMC_OK;
}

// Tables of one group (see af_thread): tile-uniform fields on the device, thread-level couplings from the host copy.
static int build_group_tables(HiGroup &G, int nl, int nt, unsigned long long gbits, const std::vector<double> &Jq_host,
                              cudaStream_t st) {
    const int64_t tiles = (int64_t)1 << (nl - TILE_BITS);
    double *cta = nullptr, *thr = nullptr;
    QMC_CUDA_TRY(cudaMalloc(&cta, sizeof(double) * AF_CTA_STRIDE * (size_t)tiles));
    QMC_CUDA_TRY(cudaMalloc(&thr, sizeof(double) * NT * AF_THR_STRIDE));
    G.args.cta_tab = cta;
    G.args.thr_tab = thr;
    const int p = G.args.p;
    af_cta_table_kernel<<<(unsigned)((tiles + 127) / 128), 128, 0, st>>>(G.kind, p, nt, gbits, G.args.Jq, G.args.hq, tiles, cta);
    QMC_CUDA_TRY(cudaGetLastError());
    std::vector<double> h((size_t)NT * AF_THR_STRIDE, 0.0);
    auto J = [&](int q1, int q2) { return q1 < q2 ? Jq_host[(size_t)q1 * nt + q2] : Jq_host[(size_t)q2 * nt + q1]; };
    for (int t = 0; t < NT; ++t) {  // thread bit a <-> index position hi_thread_pos(kind, p, a)
        double z[7];
        for (int a = 0; a < 7; ++a) z[a] = ((t >> a) & 1) ? -1.0 : 1.0;
        double P = 0.0;
        for (int a = 1; a < 7; ++a)
            for (int a2 = 0; a2 < a; ++a2) P += J(hi_thread_pos(G.kind, p, a2), hi_thread_pos(G.kind, p, a)) * z[a2] * z[a];
        h[(size_t)t * AF_THR_STRIDE] = P;
        for (int b2 = 0; b2 < 6; ++b2) {
            double Q = 0.0;
            for (int a = 0; a < 7; ++a) Q += J(hi_thread_pos(G.kind, p, a), p + b2) * z[a];
            h[(size_t)t * AF_THR_STRIDE + 1 + b2] = Q;
        }
    }
    QMC_CUDA_TRY(cudaMemcpyAsync(thr, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
    QMC_CUDA_TRY(cudaStreamSynchronize(st));  // h lives on this frame
    return QMC_OK;
}
static void free_group_tables(HiGroup &G) {
    cudaFree(const_cast<double *>(G.args.cta_tab));
    cudaFree(const_cast<double *>(G.args.thr_tab));
    G.args.cta_tab = nullptr;
    G.args.thr_tab = nullptr;
}

struct SweepWorkspace {
    int64_t cap_chains = 0;
    float2 *state = nullptr;
    ChainParams *params = nullptr;
    float *segsum = nullptr;
    double *coarse = nullptr;     // n > 20: sums of COARSE consecutive segment sums, [chains][2^(n-18)]
    int *hist = nullptr;
    int *order = nullptr;
    float *tan_sorted = nullptr;  // fast path: per-position mixer multiplier, copied into c_tan[] every hop
    int4 *af_lo = nullptr, *af_hia = nullptr, *af_hib = nullptr, *af_hic = nullptr;
    int dr_lo[64], dr_hia[64], dr_hib[64], dr_hic[64];
    bool tables_ready = false;
    std::vector<double> Jq, hq;   // circuit couplings in qubit order (host)
    double *dJq = nullptr, *dhq = nullptr;
    double fx_scale = 1.0;
    int k_fx = 0;
    std::vector<HiGroup> groups;  // n > 20: register groups above the LO qubits, with their phase tables
    std::vector<cudaEvent_t> ev;
    QmcGroupSched gsched;  // L2-resident group schedule (n <= 20): side streams + fork / join events
    // XQ mode (one- and two-qubit X strings, 14 <= n <= 20): phase tables of the MIXER diagonal in the HI layouts
    int xq_epoch = -1;            // mixer epoch the tables were built for
    int4 *afx_lo = nullptr, *afx_hia = nullptr, *afx_hib = nullptr, *afx_hic = nullptr;
    int *d_drx = nullptr;         // [mixer groups][hia, hib, hic][64] Gray increments of the mixer diagonals
    int xq_groups = 0;
    double *dxJ = nullptr, *dxh = nullptr;
    double fx_scale_x = 1.0;
    int k_fx_x = 0;
    // XT mode (any other X-string mixer, 14 <= n <= 20): tabulated mixer exponent, int32 fixed point, [mixer groups][2^n]
    int xt_epoch = -1;
    int xt_groups = 0;
    int *d_xt = nullptr;
};

void qmc_sweep_free(qmc_model *m) {
    SweepWorkspace *w = m->sweep;
    if (!w) return;
    cudaFree(w->state);
    cudaFree(w->params);
    cudaFree(w->segsum);
    cudaFree(w->coarse);
    cudaFree(w->hist);
    cudaFree(w->order);
    cudaFree(w->tan_sorted);
    cudaFree(w->af_lo);
    cudaFree(w->af_hia);
    cudaFree(w->af_hib);
    cudaFree(w->af_hic);
    cudaFree(w->dJq);
    cudaFree(w->dhq);
    for (auto &G : w->groups) free_group_tables(G);
    for (auto e : w->ev) cudaEventDestroy(e);
    w->gsched.destroy();
    cudaFree(w->afx_lo); cudaFree(w->afx_hia); cudaFree(w->afx_hib); cudaFree(w->afx_hic); cudaFree(w->dxJ); cudaFree(w->dxh);
    cudaFree(w->d_drx);
    cudaFree(w->d_xt);
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
            if (n == 20) {  // pair-register HI layout (hiq_sweep_kernel)
                QMC_CUDA_TRY(cudaMalloc(&w->af_hic, af_bytes));
                const int grid = (int)std::min<int64_t>(148 * 8, (((int64_t)1 << (n - 6)) + 127) / 128);
                af_table_kernel<8, 3><<<grid, 128, 0, st>>>(n, w->dJq, w->dhq, w->fx_scale, w->af_hic);
                m->launches++;
                const int pc[6] = {hiq_pos_b(0), hiq_pos_b(1), hiq_pos_b(2), hiq_pos_b(3), hiq_pos_b(4), hiq_pos_b(5)};
                gray_increments(w->Jq, n, w->fx_scale, pc, w->dr_hic);
            }
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
        cudaFree(w->state); cudaFree(w->params); cudaFree(w->segsum); cudaFree(w->coarse); cudaFree(w->hist); cudaFree(w->order); cudaFree(w->tan_sorted);
        w->state = nullptr; w->params = nullptr; w->segsum = nullptr; w->coarse = nullptr; w->hist = nullptr; w->order = nullptr; w->tan_sorted = nullptr;
        w->cap_chains = 0;
        QMC_CUDA_TRY(cudaMalloc(&w->state, sizeof(float2) * (size_t)N * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->params, sizeof(ChainParams) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->segsum, sizeof(float) * (size_t)(N >> 6) * chains));
        if (n > 20) QMC_CUDA_TRY(cudaMalloc(&w->coarse, sizeof(double) * (size_t)(N >> (6 + COARSE_BITS)) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->hist, sizeof(int) * (size_t)(n + 1) * chains));
        QMC_CUDA_TRY(cudaMalloc(&w->order, sizeof(int) * (size_t)chains));
        QMC_CUDA_TRY(cudaMalloc(&w->tan_sorted, sizeof(float) * (size_t)chains));
        w->cap_chains = chains;
    }
    return QMC_OK;
}

// XT mode: E_x(x) = sum_S w_S (-1)^popcount(x & S) is the Walsh-Hadamard transform of the coefficient vector w[S] (indexed by
// the term's mask): n in-place butterfly passes in fp64 on the device, then one rounding to int32 fixed point.
__global__ void xt_wht_kernel(double *__restrict__ x, int64_t half, int bit) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < half; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = ((q >> bit) << (bit + 1)) | (q & (((int64_t)1 << bit) - 1)), j = i | ((int64_t)1 << bit);
        const double u = x[i], v = x[j];
        x[i] = u + v;
        x[j] = u - v;
    }
}
__global__ void xt_round_kernel(const double *__restrict__ x, int64_t N, double scale, int *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int)rint(x[i] * scale);
}
static int ensure_xt_tables(qmc_model *m, cudaStream_t st) {
    SweepWorkspace *w = m->sweep;
    if (w->xt_epoch == m->mixer_epoch) return QMC_OK;
    const int n = m->n, ng = m->n_groups;
    const int64_t N = 1LL << n;
    QMC_REQUIRE(n >= 14 && n <= 20, QMC_ERR_UNSUPPORTED, "sweep path: tabulated mixer diagonals cover 14 <= n <= 20 (n=%d)", n);
    QMC_CUDA_TRY(cudaStreamSynchronize(st));
    double dmax = 0.0;  // one fixed-point scale for every group: |E_x| <= sum |w_S|
    for (int g = 0; g < ng; ++g) {
        double d = 0.0;
        for (int k = m->group_off[g]; k < m->group_off[g + 1]; ++k) d += fabs(m->weights[k]);
        dmax = std::max(dmax, d);
    }
    if (dmax < 1e-300) dmax = 1.0;
    w->k_fx_x = (int)floor(log2(2147483647.0 / (dmax * 1.0000001))) - 1;
    if (w->k_fx_x > 30) w->k_fx_x = 30;
    w->fx_scale_x = ldexp(1.0, w->k_fx_x);
    if (w->xt_groups != ng) {
        cudaFree(w->d_xt);
        w->d_xt = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&w->d_xt, sizeof(int) * (size_t)ng * N));
        w->xt_groups = ng;
    }
    double *tmp = nullptr;
    QMC_CUDA_TRY(cudaMalloc(&tmp, sizeof(double) * (size_t)N));
    std::vector<double> coef((size_t)N);
    cudaError_t err = cudaSuccess;
    for (int g = 0; g < ng && err == cudaSuccess; ++g) {
        std::fill(coef.begin(), coef.end(), 0.0);
        for (int k = m->group_off[g]; k < m->group_off[g + 1]; ++k) coef[(size_t)m->masks[k]] += m->weights[k];
        err = cudaMemcpyAsync(tmp, coef.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st);
        if (err != cudaSuccess) break;
        for (int b = 0; b < n; ++b) xt_wht_kernel<<<148 * 8, 256, 0, st>>>(tmp, N >> 1, b);
        xt_round_kernel<<<148 * 8, 256, 0, st>>>(tmp, N, w->fx_scale_x, w->d_xt + (size_t)g * N);
        m->launches += n + 1;
        err = cudaStreamSynchronize(st);  // coef is reused by the next group
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    cudaFree(tmp);  // on every path
    QMC_CUDA_TRY(err);
    w->xt_epoch = m->mixer_epoch;
    return QMC_OK;
}

// XQ mode: A / F_b tables and Gray increments of the mixer diagonal E_x(x) = sum_q xh[q] z_q + sum_{q<q'} xJ[q][q'] z_q z_q'
// for the HI register layouts -- the phase-layer machinery fed with the mixer's weights (rebuilt when the mixer changes).
static int ensure_xq_tables(qmc_model *m, cudaStream_t st) {
    SweepWorkspace *w = m->sweep;
    if (w->xq_epoch == m->mixer_epoch) return QMC_OK;
    const int n = m->n, ng = m->n_groups;
    const int64_t N = 1LL << n;
    QMC_REQUIRE(m->quad_mixer && n >= 14 && n <= 20 && (int)m->xh.size() == ng * n, QMC_ERR_UNSUPPORTED,
                "sweep path: the mixer groups are not made of one- and two-qubit X strings (n=%d)", n);
    QMC_CUDA_TRY(cudaStreamSynchronize(st));  // tables of an earlier mixer may still be in use
    double dmax = 0.0;  // one fixed-point scale for every group
    for (int g = 0; g < ng; ++g) {
        double d = 0.0;
        for (int a = 0; a < n; ++a) {
            d += fabs(m->xh[(size_t)g * n + a]);
            for (int b = a + 1; b < n; ++b) d += fabs(m->xJ[((size_t)g * n + a) * n + b]);
        }
        dmax = std::max(dmax, d);
    }
    if (dmax < 1e-300) dmax = 1.0;
    w->k_fx_x = (int)floor(log2(2147483647.0 / (dmax * 1.0000001))) - 1;
    if (w->k_fx_x > 30) w->k_fx_x = 30;
    w->fx_scale_x = ldexp(1.0, w->k_fx_x);
    const size_t af_entries = 2 * (size_t)(N >> 6);
    const size_t af_bytes = sizeof(int4) * af_entries;
    if (w->xq_groups != ng) {
        cudaFree(w->dxJ); cudaFree(w->dxh); cudaFree(w->afx_lo); cudaFree(w->afx_hia); cudaFree(w->afx_hib); cudaFree(w->afx_hic);
        cudaFree(w->d_drx);
        w->dxJ = w->dxh = nullptr; w->afx_lo = w->afx_hia = w->afx_hib = w->afx_hic = nullptr; w->d_drx = nullptr;
        QMC_CUDA_TRY(cudaMalloc(&w->dxJ, sizeof(double) * (size_t)ng * n * n));
        QMC_CUDA_TRY(cudaMalloc(&w->dxh, sizeof(double) * (size_t)ng * n));
        QMC_CUDA_TRY(cudaMalloc(&w->afx_lo, af_bytes));  // scratch: build_af_tables also writes the LO layout
        QMC_CUDA_TRY(cudaMalloc(&w->afx_hia, af_bytes * ng));
        QMC_CUDA_TRY(cudaMalloc(&w->afx_hib, af_bytes * ng));
        if (n == 20) QMC_CUDA_TRY(cudaMalloc(&w->afx_hic, af_bytes * ng));
        QMC_CUDA_TRY(cudaMalloc(&w->d_drx, sizeof(int) * (size_t)ng * 3 * 64));
        w->xq_groups = ng;
    }
    QMC_CUDA_TRY(cudaMemcpy(w->dxJ, m->xJ.data(), sizeof(double) * (size_t)ng * n * n, cudaMemcpyHostToDevice));
    QMC_CUDA_TRY(cudaMemcpy(w->dxh, m->xh.data(), sizeof(double) * (size_t)ng * n, cudaMemcpyHostToDevice));
    std::vector<int> drx((size_t)ng * 3 * 64, 0);
    for (int g = 0; g < ng; ++g) {
        const double *dJ = w->dxJ + (size_t)g * n * n, *dh = w->dxh + (size_t)g * n;
        const std::vector<double> xJg(m->xJ.begin() + (size_t)g * n * n, m->xJ.begin() + (size_t)(g + 1) * n * n);
        int pos[3][6];
        int tmp[64];
        if (n == 20) {
            const int grid = (int)std::min<int64_t>(148 * 8, (((int64_t)1 << (n - 6)) + 127) / 128);
            af_table_kernel<8, 3><<<grid, 128, 0, st>>>(n, dJ, dh, w->fx_scale_x, w->afx_hic + g * af_entries);
            m->launches++;
            const int pc[6] = {hiq_pos_b(0), hiq_pos_b(1), hiq_pos_b(2), hiq_pos_b(3), hiq_pos_b(4), hiq_pos_b(5)};
            gray_increments(xJg, n, w->fx_scale_x, pc, tmp);
            memcpy(&drx[((size_t)g * 3 + 2) * 64], tmp, sizeof(tmp));
        }
        int rc = build_af_tables(n, dJ, dh, w->fx_scale_x, w->afx_lo, w->afx_hia + g * af_entries, w->afx_hib + g * af_entries, pos, st);
        if (rc != QMC_OK) return rc;
        m->launches += 3;
        gray_increments(xJg, n, w->fx_scale_x, pos[1], tmp);
        memcpy(&drx[((size_t)g * 3 + 0) * 64], tmp, sizeof(tmp));
        gray_increments(xJg, n, w->fx_scale_x, pos[2], tmp);
        memcpy(&drx[((size_t)g * 3 + 1) * 64], tmp, sizeof(tmp));
    }
    QMC_CUDA_TRY(cudaGetLastError());
    QMC_CUDA_TRY(cudaMemcpy(w->d_drx, drx.data(), sizeof(int) * drx.size(), cudaMemcpyHostToDevice));
    w->xq_epoch = m->mixer_epoch;
    return QMC_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: `done` is a bit mask indexed by
// the current device, so a process that drives several GPUs opts in on each of them.
typedef unsigned long long DevMask;
template <typename K>
static int set_smem_once(K kernel, DevMask &done, int bytes = TILE * 8) {
    int dev = 0;
    QMC_CUDA_TRY(cudaGetDevice(&dev));
    const DevMask bit = 1ULL << (dev & 63);
    if (!(done & bit)) {
        QMC_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done |= bit;
    }
    return QMC_OK;
}

// Launch-time variant: fast = normalisation folded into the initial product state (phase factors of unit modulus)
// and one mixer multiplier per chain served from the constant bank (c_tan[], position-indexed).
struct SweepTune {
    bool fast = false;
};

template <int ROLE, bool FAST>
static int launch_lo_v(const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    if (ROLE == LO_MID && FAST && !a.xchg_peers) {
        // 64-thread CTAs on 2^12-amplitude tiles (6 resident CTAs per SM instead of 3 x 128 threads: the same 12 warps, but
        // barriers span two warps and six tiles are in different phases): +1.5 % on C3, measured 3 x A/B on one box
        // (43.78 k vs 43.12 k proposals/s, profiles/r2c_lo_nth_ab.txt).  QUMCMC_LO_NTH=128 restores the 128-thread CTAs.
        static const int nth = getenv("QUMCMC_LO_NTH") ? atoi(getenv("QUMCMC_LO_NTH")) : 64;
        if (nth == 64) {
            static DevMask done64 = 0;
            int rc64 = set_smem_once(lo_sweep_kernel<LO_MID, 64, true>, done64, TILE * 4);
            if (rc64 != QMC_OK) return rc64;
            SweepArgs a64 = a;
            static const int pf64_env = getenv("QUMCMC_LO_PF") ? atoi(getenv("QUMCMC_LO_PF")) : -1;
            a64.pf_dist = pf64_env >= 0 && a.pf_dist > 0 ? pf64_env : 2 * a.pf_dist;
            lo_sweep_kernel<LO_MID, 64, true><<<dim3((unsigned)(2 * tiles), (unsigned)count), 64, TILE * 4, st>>>(a64, order, off, off);
            return QMC_OK;
        }
    }
    static DevMask done = 0;
    int rc = set_smem_once(lo_sweep_kernel<ROLE, NT, FAST>, done);
    if (rc != QMC_OK) return rc;
    lo_sweep_kernel<ROLE, NT, FAST><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
    return QMC_OK;
}
// XQ mode (general two-body path): MID on 64-thread CTAs, FINAL on 128-thread CTAs; multipliers are the constants +-1
template <int ROLE>
static int launch_lo_xq(const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    static_assert(ROLE == LO_MID || ROLE == LO_FINAL, "XQ mode has LO MID and LO FINAL sweeps only");
    if (ROLE == LO_MID) {
        static DevMask done64 = 0;
        int rc64 = set_smem_once(lo_sweep_kernel<LO_MID, 64, true, true>, done64, TILE * 4);
        if (rc64 != QMC_OK) return rc64;
        SweepArgs a64 = a;
        a64.pf_dist = 2 * a.pf_dist;
        lo_sweep_kernel<LO_MID, 64, true, true><<<dim3((unsigned)(2 * tiles), (unsigned)count), 64, TILE * 4, st>>>(a64, order, off, off);
        return QMC_OK;
    }
    static DevMask done = 0;
    int rc = set_smem_once(lo_sweep_kernel<ROLE, NT, true, true>, done);
    if (rc != QMC_OK) return rc;
    lo_sweep_kernel<ROLE, NT, true, true><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
    return QMC_OK;
}
template <int W, bool FIRST>
static int launch_hi_xq_w(const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    if (a.xt_ex) {  // XT mode: tabulated mixer exponent
        static DevMask donet = 0;
        int rct = set_smem_once(hi_sweep_kernel<W, FIRST, true, true, true>, donet);
        if (rct != QMC_OK) return rct;
        hi_sweep_kernel<W, FIRST, true, true, true><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
        return QMC_OK;
    }
    if (W == 8 && !FIRST && a.hiq) {
        static DevMask doneq = 0;
        int rcq = set_smem_once(hiq_sweep_kernel<true, true>, doneq);
        if (rcq != QMC_OK) return rcq;
        hiq_sweep_kernel<true, true><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
        return QMC_OK;
    }
    static DevMask done = 0;
    int rc = set_smem_once(hi_sweep_kernel<W, FIRST, true, true>, done);
    if (rc != QMC_OK) return rc;
    hi_sweep_kernel<W, FIRST, true, true><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
    return QMC_OK;
}
template <bool FIRST>
static int launch_hi_xq(int n, const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    switch (n - LO_BITS) {
        case 2: return launch_hi_xq_w<2, FIRST>(a, order, off, tiles, count, st);
        case 3: return launch_hi_xq_w<3, FIRST>(a, order, off, tiles, count, st);
        case 4: return launch_hi_xq_w<4, FIRST>(a, order, off, tiles, count, st);
        case 5: return launch_hi_xq_w<5, FIRST>(a, order, off, tiles, count, st);
        case 6: return launch_hi_xq_w<6, FIRST>(a, order, off, tiles, count, st);
        case 7: return launch_hi_xq_w<7, FIRST>(a, order, off, tiles, count, st);
        case 8: return launch_hi_xq_w<8, FIRST>(a, order, off, tiles, count, st);
    }
    qmc_set_error("sweep path: n=%d outside [14, 20]", n);
    return QMC_ERR_UNSUPPORTED;
}

template <int ROLE>
static int launch_lo(const SweepTune &tu, const SweepArgs &a, const int *order, int off, int64_t tiles, int count,
                     cudaStream_t st) {
#ifdef QMC_PROBES
    if (ROLE == LO_MID) {
        const char *pe = getenv("QUMCMC_SWEEP_PROBE");
        const int mode = pe ? atoi(pe) : 0;
        static DevMask done1 = 0, done2 = 0;
        if (mode == 1) {
            set_smem_once(lo_probe_kernel<9, 3>, done1);
            lo_probe_kernel<9, 3><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off);
            return QMC_OK;
        }
        if (mode == 2) {
            set_smem_once(lo_probe_kernel<14, 3>, done2);
            lo_probe_kernel<14, 3><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off);
            return QMC_OK;
        }
    }
#endif
    if (ROLE != LO_SINGLE && ROLE != LO_FINAL1 && tu.fast)
        return launch_lo_v<ROLE, ROLE != LO_SINGLE && ROLE != LO_FINAL1>(a, order, off, tiles, count, st);
    return launch_lo_v<ROLE, false>(a, order, off, tiles, count, st);
}

template <int W, bool FIRST, bool FAST>
static int launch_hi_w(const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    if (W == 8 && !FIRST && a.hiq) {
        static DevMask doneq = 0;
        int rcq = set_smem_once(hiq_sweep_kernel<FAST>, doneq);
        if (rcq != QMC_OK) return rcq;
        hiq_sweep_kernel<FAST><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
        return QMC_OK;
    }
    static DevMask done = 0;
    int rc = set_smem_once(hi_sweep_kernel<W, FIRST, FAST>, done);
    if (rc != QMC_OK) return rc;
    hi_sweep_kernel<W, FIRST, FAST><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, order, off, off);
    return QMC_OK;
}

template <bool FIRST, bool FAST>
static int launch_hi_f(int n, const SweepArgs &a, const int *order, int off, int64_t tiles, int count, cudaStream_t st) {
    switch (n - LO_BITS) {
        case 2: return launch_hi_w<2, FIRST, FAST>(a, order, off, tiles, count, st);
        case 3: return launch_hi_w<3, FIRST, FAST>(a, order, off, tiles, count, st);
        case 4: return launch_hi_w<4, FIRST, FAST>(a, order, off, tiles, count, st);
        case 5: return launch_hi_w<5, FIRST, FAST>(a, order, off, tiles, count, st);
        case 6: return launch_hi_w<6, FIRST, FAST>(a, order, off, tiles, count, st);
        case 7: return launch_hi_w<7, FIRST, FAST>(a, order, off, tiles, count, st);
        case 8: return launch_hi_w<8, FIRST, FAST>(a, order, off, tiles, count, st);
    }
    qmc_set_error("sweep path: n=%d outside [14, 20]", n);
    return QMC_ERR_UNSUPPORTED;
}
template <bool FIRST>
static int launch_hi(const SweepTune &tu, int n, const SweepArgs &a, const int *order, int off, int64_t tiles, int count,
                     cudaStream_t st) {
    return tu.fast ? launch_hi_f<FIRST, true>(n, a, order, off, tiles, count, st)
                   : launch_hi_f<FIRST, false>(n, a, order, off, tiles, count, st);
}

template <int ROLE, int ACT>
static int launch_hi6_ra(const SweepArgs &a, const Hi6Args &g, const int *order, int off, int64_t tiles, int count,
                         cudaStream_t st) {
    hi6_sweep_kernel<ROLE, ACT><<<dim3((unsigned)tiles, (unsigned)count), NT, 0, st>>>(a, g, order, off);
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

// MID on the top group of a shard: before the phase only the `pre` swapped-in (top) qubits are mixed
static int launch_hi6_pre(int pre, const SweepArgs &a, const Hi6Args &g, const int *order, int off, int64_t tiles, int count,
                          cudaStream_t st) {
    const size_t smem = 0;
    switch (pre) {
        case 1: hi6_sweep_kernel<HI6_MID, 6, 1><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off); return QMC_OK;
        case 2: hi6_sweep_kernel<HI6_MID, 6, 2><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off); return QMC_OK;
        case 3: hi6_sweep_kernel<HI6_MID, 6, 3><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off); return QMC_OK;
        case 4: hi6_sweep_kernel<HI6_MID, 6, 4><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off); return QMC_OK;
        case 5: hi6_sweep_kernel<HI6_MID, 6, 5><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off); return QMC_OK;
        case 6: hi6_sweep_kernel<HI6_MID, 6, 6><<<dim3((unsigned)tiles, (unsigned)count), NT, smem, st>>>(a, g, order, off); return QMC_OK;
    }
    qmc_set_error("hi6 sweep: %d swapped-in qubits", pre);
    return QMC_ERR_BADARG;
}

template <int ROLE, int ACT, int PRE>
static int launch_hi8_rap(const SweepArgs &a, const Hi6Args &g, const int *order, int off, int64_t tiles, int count,
                          cudaStream_t st) {
    static DevMask done = 0;
    int rc = set_smem_once(hi8_sweep_kernel<ROLE, ACT, PRE>, done);
    if (rc != QMC_OK) return rc;
    hi8_sweep_kernel<ROLE, ACT, PRE><<<dim3((unsigned)tiles, (unsigned)count), NT, TILE * 8, st>>>(a, g, order, off);
    return QMC_OK;
}
template <int ACT>
static int launch_hi8_mid_pre(int pre, const SweepArgs &a, const Hi6Args &g, const int *order, int off, int64_t tiles,
                              int count, cudaStream_t st) {
    switch (pre) {
        case 1: return launch_hi8_rap<HI8_MID, ACT, 1>(a, g, order, off, tiles, count, st);
        case 2: return launch_hi8_rap<HI8_MID, ACT, 2>(a, g, order, off, tiles, count, st);
        case 3: return launch_hi8_rap<HI8_MID, ACT, 3>(a, g, order, off, tiles, count, st);
        case 4: return launch_hi8_rap<HI8_MID, ACT, 4>(a, g, order, off, tiles, count, st);
        case 5: return launch_hi8_rap<HI8_MID, ACT, 5>(a, g, order, off, tiles, count, st);
        case 6: return launch_hi8_rap<HI8_MID, ACT, 6>(a, g, order, off, tiles, count, st);
    }
    qmc_set_error("hi8 sweep: %d swapped-in qubits", pre);
    return QMC_ERR_BADARG;
}

// ROLE: 0 FIRST | 1 MID | 2 SINGLE (HI6_* == HI8_*).  pre > 0: MID mixes only the top `pre` qubits before the phase;
// SINGLE mixes only the top `pre` qubits (the ones an exchange brought in).
template <int ROLE>
static int launch_group(const HiGroup &G, int pre, const SweepArgs &a, const int *order, int off, int64_t tiles, int count,
                        cudaStream_t st) {
    if (G.kind == 6) {
        if (ROLE == HI6_MID && pre > 0 && pre != G.act) {
            if (G.act != 6) {
                qmc_set_error("restricted MID needs a full top group");
                return QMC_ERR_BADARG;
            }
            return launch_hi6_pre(pre, a, G.args, order, off, tiles, count, st);
        }
        return launch_hi6<ROLE>(ROLE == HI6_SINGLE && pre > 0 ? pre : G.act, a, G.args, order, off, tiles, count, st);
    }
    if (ROLE == HI8_SINGLE && pre > 0) {  // the top `pre` (<= 6) qubits of the window: a register-only pass reaches them
        Hi6Args g6 = G.args;
        g6.p = G.args.p + 2;
        return launch_hi6<HI6_SINGLE>(pre, a, g6, order, off, tiles, count, st);
    }
    if (G.act == 8) {
        if (ROLE == HI8_MID && pre > 0 && pre != 8) return launch_hi8_mid_pre<8>(pre, a, G.args, order, off, tiles, count, st);
        return launch_hi8_rap<ROLE, 8, 8>(a, G.args, order, off, tiles, count, st);
    }
    if (ROLE == HI8_MID && pre > 0 && pre != 7) return launch_hi8_mid_pre<7>(pre, a, G.args, order, off, tiles, count, st);
    return launch_hi8_rap<ROLE, 7, 7>(a, G.args, order, off, tiles, count, st);
}

// Sweep list of one proposal with r mixer layers on the generic path (LO group + K >= 2 register groups):
//   FIRST(g0); per middle layer: LO_SINGLE, SINGLE for the other groups, MID on the next group (which carries
//   its M into the following layer); last layer: SINGLE..., LO_FINAL.  K passes per layer for K + 1 qubit groups.
static int run_generic_schedule(int n, int r, const SweepArgs &a, const std::vector<HiGroup> &groups, const int *order,
                                int off, int count, int64_t tiles, cudaStream_t st, int &launched) {
    int rc = QMC_OK;
    const int K = (int)groups.size();
    if (r <= 1) {
        ++launched;
        return launch_lo<LO_FINAL1>(SweepTune(), a, order, off, tiles, count, st);
    }
    rc = launch_group<HI6_FIRST>(groups[0], 0, a, order, off, tiles, count, st);
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
            rc = launch_group<HI6_SINGLE>(groups[k], 0, a, order, off, tiles, count, st);
            ++launched;
        }
        if (rc != QMC_OK) break;
        if (!last) {
            rc = launch_group<HI6_MID>(groups[mid], 0, a, order, off, tiles, count, st);
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
    // XQ mode: one coherent group of one- and two-qubit X strings (abi.cu routes it here for 14 <= n <= 20)
    // XT mode: any other mixer at n <= 20 -- the same schedule and LO sweeps, HI sweeps read the tabulated mixer exponent
    const bool xt = !m->all_one_body && !m->quad_mixer && n <= 20;
    const bool xq = !m->all_one_body && n <= 20;  // XQ or XT: 2 r sweeps per proposal
    QMC_REQUIRE(m->all_one_body || xq, QMC_ERR_UNSUPPORTED,
                "n=%d: the HBM sweep path hosts one-body X mixers (n <= %d) and every other X-string mixer at n <= 20 (XQ / XT modes: "
                "quadratic or tabulated mixer diagonal); n > 20 with other mixers takes the general path", n, QMC_MAX_SPINS);
    if (n_chains == 0) return QMC_OK;
    {
        static DevMask smem_attr = 0;
        int rca = set_smem_once(sample_accept_kernel, smem_attr);
        if (rca != QMC_OK) return rca;
    }
    const int64_t N = 1LL << n;
    const int64_t tiles = N >> TILE_BITS;
    // chains per batch: bounded by free memory (state dominates)
    SweepWorkspace *w0 = m->sweep;
    int64_t cap;
    if (w0 && w0->cap_chains >= n_chains) {
        cap = w0->cap_chains;  // steady state: the workspace of an earlier call holds every chain of this one
    } else {
        size_t free_b = 0, total_b = 0;
        QMC_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t per_chain = sizeof(float2) * (size_t)N + sizeof(float) * (size_t)(N >> 6) + sizeof(ChainParams) + 256;
        const size_t have = w0 ? (size_t)w0->cap_chains * per_chain : 0;
        cap = (int64_t)(((double)(free_b + have) * 0.85) / (double)per_chain);
        if (cap > 65535) cap = 65535;  // gridDim.y
        if (!generic && cap > CTAN_MAX) cap = CTAN_MAX;  // one c_tan[] entry per chain of a batch (fast path)
        QMC_REQUIRE(cap >= 1, QMC_ERR_CUDA, "not enough device memory for one 2^%d statevector", n);
    }
    const int64_t batch = std::min<int64_t>(cap, n_chains);
    int rc = ensure_workspace(m, batch, st);
    if (rc != QMC_OK) return rc;
    SweepWorkspace *w = m->sweep;
    if (xq) {
        rc = xt ? ensure_xt_tables(m, st) : ensure_xq_tables(m, st);
        if (rc != QMC_OK) return rc;
    }
    if (generic && w->groups.empty()) {  // register groups above the LO qubits, built once per model
        std::vector<int> kind, pos, act;
        rc = hi_group_geometry(n, kind, pos, act);
        if (rc != QMC_OK) return rc;
        for (size_t k = 0; k < kind.size(); ++k) {
            HiGroup G;
            G.kind = kind[k];
            G.act = act[k];
            G.args.p = pos[k];
            const int rp[6] = {pos[k], pos[k] + 1, pos[k] + 2, pos[k] + 3, pos[k] + 4, pos[k] + 5};
            gray_increments(w->Jq, n, w->fx_scale, rp, G.args.dr);
            G.args.Jq = w->dJq;
            G.args.hq = w->dhq;
            G.args.fx_scale = w->fx_scale;
            rc = build_group_tables(G, n, n, 0ULL, w->Jq, st);
            w->groups.push_back(G);  // owns the tables from here on
            m->launches++;
            if (rc != QMC_OK) return rc;
        }
    }
    const std::vector<HiGroup> &groups = w->groups;

    // tuning knobs (defaults = measured best on B200; the env overrides exist for A/B runs and profiling)
    const char *pf_env = getenv("QUMCMC_SWEEP_PF");
    const int pf_dist = pf_env ? atoi(pf_env) : 222;
    const char *hiq_env = getenv("QUMCMC_SWEEP_HIQ");
    const bool hiq_on = hiq_env ? atoi(hiq_env) != 0 : true;
    const char *pfl_env = getenv("QUMCMC_SWEEP_PFLINES");
    const int pf_lines = pfl_env ? atoi(pfl_env) : 1;
    const char *fast_env = getenv("QUMCMC_SWEEP_FAST");
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
    if (!xq) {
        double log_min = 0.0;  // min over proposals of r * sum_q log|cos(gamma w_q dt)|
        auto log_scale = [&](double gamma, int grp) {
            double ls = 0.0;
            for (int q = 0; q < n; ++q) ls += log(std::max(1e-300, fabs(cos(gamma * m->qweight[(size_t)grp * n + q] * dt))));
            return ls;
        };
        if (mode == QMC_MODE_CHAIN) {
            for (int grp = 0; grp < m->n_groups; ++grp) {
                bool monotone = true;  // |cos| is monotone on [g_lo w dt, g_hi w dt] when the angles stay below pi/2
                for (int q = 0; q < n; ++q)
                    monotone = monotone && fabs(std::max(fabs(g_lo), fabs(g_hi)) * m->qweight[(size_t)grp * n + q] * dt) < 1.5;
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
        // ... and the constant-bank multiplier needs one tan(gamma w dt) per chain: every group must put the same
        // weight on all n qubits (GenericMixer.OneBodyMixer and sums of it), and the batch must fit c_tan[].
        bool uniform_w = true;
        for (int grp = 0; grp < m->n_groups && uniform_w; ++grp) {
            const double *qw = m->qweight.data() + (size_t)grp * n;
            uniform_w = qw[0] != 0.0;
            for (int q = 0; q < n && uniform_w; ++q) uniform_w = qw[q] == qw[0];
        }
        tune.fast = !generic && log_min > log(1e-24) && uniform_w && batch <= CTAN_MAX;
        if (fast_env) tune.fast = tune.fast && fast_env[0] == '1';
    }

    if (xq) r_max *= 2;  // XQ mode schedules half-layers: 2 r sweeps per proposal (HI first, LO last)
    QMC_REQUIRE(r_max <= SCHED_RMAX, QMC_ERR_UNSUPPORTED, "delta_time=%g gives up to %d Trotter layers (> %d)", dt, r_max, SCHED_RMAX);
    std::vector<int> rhist(r_max + 2, 0);
    QmcGroupKnobs l2k;  // L2-resident group schedule (two-group path only), see group_sched.cuh
    if (!generic) {
        l2k = qmc_group_knobs((int64_t)sizeof(float2) << n, 24, 3);
    } else if (n <= 22) {
        // one chain per group, 64 MiB of statevectors in flight (benchmarks/l2_generic_sweep.py: n = 21 12.0 -> 13.7 k
        // proposals/s, n = 22 6.1 -> 6.9 k; n = 23, a single chain in flight: 2.96 -> 2.86 k, left on the r-major schedule)
        l2k = qmc_group_knobs((int64_t)sizeof(float2) << n, 1, n == 21 ? 4 : 2);
        if (getenv("QUMCMC_L2_GENERIC") && atoi(getenv("QUMCMC_L2_GENERIC")) == 0) l2k.group = 0;
    }
    for (int64_t c0 = 0; c0 < n_chains; c0 += batch) {
        const int64_t nc = std::min<int64_t>(batch, n_chains - c0);
        HopArgs h;
        h.n = n; h.mode = mode; h.n_groups = m->n_groups; h.phase_active = m->phase_active; h.k_fx = w->k_fx; h.fold = tune.fast ? 1 : 0;
        h.xq = xq ? 1 : 0; h.k_fx_x = w->k_fx_x;
        h.n_chains = nc; h.n_hops = n_hops; h.hop = 0;
        h.chain_id0 = chain_id0 + (uint64_t)c0; h.seed = seed; h.hop0 = hop0;
        h.temperature = temperature; h.g_lo = g_lo; h.g_hi = g_hi; h.dt = dt; h.alpha = m->alpha;
        h.J = m->dJ; h.h = m->dh;
        h.group_off = m->d_group_off; h.masks = m->d_masks; h.weights = m->d_weights; h.group_cum = m->d_group_cum; h.qweight = m->d_qweight;
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
        SweepArgs a;
        a.n = n; a.n_total = n; a.gbits = 0ULL; a.state = w->state; a.params = w->params; a.af_lo = w->af_lo; a.af_hia = w->af_hia; a.af_hib = w->af_hib; a.af_hic = w->af_hic;
        memcpy(a.dr_lo, w->dr_lo, sizeof(a.dr_lo)); memcpy(a.dr_hia, w->dr_hia, sizeof(a.dr_hia)); memcpy(a.dr_hib, w->dr_hib, sizeof(a.dr_hib));
        if (w->af_hic) memcpy(a.dr_hic, w->dr_hic, sizeof(a.dr_hic));
        a.hiq = (w->af_hic && hiq_on) ? 1 : 0;
        a.segsum = w->segsum;
        a.pf_dist = pf_dist;
        a.pf_lines = pf_lines;
        a.gen_scale = 1.f;
        a.xq_dr = nullptr;
        a.xq_af_stride = 0;
        a.xt_ex = nullptr;
        a.xchg_peers = nullptr; a.xchg_shift = 0; a.xchg_rank = 0; a.xchg_g = 0; a.xchg_tile_bits = 0; a.xchg_idx0 = 0;
        a.probs_out = mode == QMC_MODE_PROBS ? reinterpret_cast<float *>(probs_out) : nullptr;
        a.amps_out = mode == QMC_MODE_PROBS ? reinterpret_cast<float2 *>(amps_out) : nullptr;
        // PROBS with amps only: still need a probs-free normalisation; handled by the normalise kernel

        const int gb = (int)((nc + 127) / 128);
        chains_init_kernel<<<gb, 128, 0, st>>>(h, w->params);
        m->launches++;
        const int64_t hops = mode == QMC_MODE_CHAIN ? n_hops : 1;
        for (int64_t hop = 0; hop < hops; ++hop) {
            QMC_NVTX("sweep hop: prepare, schedule, sweeps, sample+accept");
            h.hop = hop;
            hop_prepare_kernel<<<gb, 128, 0, st>>>(h, w->params);
            schedule_kernel<<<1, 1024, 0, st>>>(w->params, nc, r_max, w->order, (generic || xq) ? 0 : 1,
                                                tune.fast ? w->tan_sorted : nullptr);
            if (tune.fast)
                QMC_CUDA_TRY(cudaMemcpyToSymbolAsync(c_tan, w->tan_sorted, sizeof(float) * (size_t)nc, 0,
                                                     cudaMemcpyDeviceToDevice, st));
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
                if (xq) r *= 2;
                rhist[r]++;
                sum_r += r;
            }
            int launched = 0;
            const auto host_t0 = std::chrono::steady_clock::now();
            if (generic) {
                // order[] is sorted by r descending: one schedule per distinct r over its contiguous range
                int off = 0;
                if (l2k.group > 0 && nc > l2k.group) {
                    // n = 21, 22: a few statevectors (16 / 32 MiB) still fit the L2 together -- every chain runs all its
                    // passes back to back on one of the side streams (group_sched.cuh), 4 / 2 chains in flight
                    QmcGroupSched &gs = w->gsched;
                    rc = gs.ensure(l2k.streams);
                    if (rc != QMC_OK) return rc;
                    QMC_CUDA_TRY(cudaEventRecord(gs.ev_fork, st));
                    for (int i = 0; i < l2k.streams; ++i) QMC_CUDA_TRY(cudaStreamWaitEvent(gs.aux[i], gs.ev_fork, 0));
                    SweepArgs ag = a;
                    if (!l2k.prefetch) ag.pf_dist = 0;
                    int gi = 0;
                    for (int r = r_max; r >= 1; --r) {
                        for (int g0 = 0; g0 < rhist[r]; g0 += l2k.group, ++gi) {
                            rc = run_generic_schedule(n, r, ag, groups, w->order, off + g0, std::min(l2k.group, rhist[r] - g0), tiles,
                                                      gs.aux[gi % l2k.streams], launched);
                            if (rc != QMC_OK) return rc;
                        }
                        off += rhist[r];
                    }
                    for (int i = 0; i < l2k.streams; ++i) {
                        QMC_CUDA_TRY(cudaEventRecord(gs.ev_join[i], gs.aux[i]));
                        QMC_CUDA_TRY(cudaStreamWaitEvent(st, gs.ev_join[i], 0));
                    }
                } else
                for (int r = r_max; r >= 1; --r) {
                    if (rhist[r] == 0) continue;
                    rc = run_generic_schedule(n, r, a, groups, w->order, off, rhist[r], tiles, st, launched);
                    if (rc != QMC_OK) return rc;
                    off += rhist[r];
                }
            } else {
                int base[2] = {0, 0};  // start of the odd-r / even-r class in order[]
                for (int r = 1; r <= r_max; r += 2) base[1] += rhist[r];  // even class starts after all odd-r chains
                const bool grouped = l2k.group > 0 && nc > l2k.group;
                SweepArgs ag = a;
                if (grouped && !l2k.prefetch) ag.pf_dist = 0;  // the tiles are L2 hits already: no prefetch instructions
                // XQ mode: LO sweeps host the phase layer (model tables), HI sweeps the mixer diagonal (mixer tables); the
                // phase factor of every sweep carries the normalisation of the transforms of that sweep and of the FINAL one
                SweepArgs ax_lo = ag, ax_hif = ag, ax_him = ag;
                if (xq) {
                    const int nh = n - LO_BITS;
                    ax_lo.gen_scale = ldexpf(1.f, -LO_BITS);
                    ax_him.af_hia = ax_hif.af_hia = w->afx_hia;
                    ax_him.af_hib = ax_hif.af_hib = w->afx_hib;
                    ax_him.af_hic = ax_hif.af_hic = w->afx_hic;
                    ax_him.xq_dr = ax_hif.xq_dr = w->d_drx;
                    ax_him.xq_af_stride = ax_hif.xq_af_stride = 2 * (N >> 6);
                    if (xt) ax_him.xt_ex = ax_hif.xt_ex = w->d_xt;
                    ax_him.gen_scale = ldexpf(1.f, -nh);
                    ax_hif.gen_scale = (float)exp2(-0.5 * nh - 0.5 * LO_BITS);  // its own back-transform + the FINAL sweep's
                }
                auto lo = [&](int kind, int off, int count, cudaStream_t sg) {
                    if (xq) {
                        if (kind == QMC_GS_MID) return launch_lo_xq<LO_MID>(ax_lo, w->order, off, tiles, count, sg);
                        if (kind == QMC_GS_FINAL) return launch_lo_xq<LO_FINAL>(ax_lo, w->order, off, tiles, count, sg);
                        qmc_set_error("XQ schedule: LO role %d cannot occur (every proposal has an even number of sweeps)", kind);
                        return (int)QMC_ERR_CUDA;
                    }
                    switch (kind) {
                        case QMC_GS_FIRST: return launch_lo<LO_FIRST>(tune, ag, w->order, off, tiles, count, sg);
                        case QMC_GS_MID: return launch_lo<LO_MID>(tune, ag, w->order, off, tiles, count, sg);
                        case QMC_GS_FINAL: return launch_lo<LO_FINAL>(tune, ag, w->order, off, tiles, count, sg);
                        default: return launch_lo<LO_FINAL1>(tune, ag, w->order, off, tiles, count, sg);
                    }
                };
                auto hi = [&](bool first, int off, int count, cudaStream_t sg) {
                    if (xq)
                        return first ? launch_hi_xq<true>(n, ax_hif, w->order, off, tiles, count, sg)
                                     : launch_hi_xq<false>(n, ax_him, w->order, off, tiles, count, sg);
                    return first ? launch_hi<true>(tune, n, ag, w->order, off, tiles, count, sg)
                                 : launch_hi<false>(tune, n, ag, w->order, off, tiles, count, sg);
                };
                if (grouped) {
                    rc = qmc_group_schedule(w->gsched, l2k, st, (int)nc, r_max, rhist, lo, hi, launched);
                    if (rc != QMC_OK) return rc;
                } else
                for (int step = 1; step <= r_max; ++step) {
                    const int p = step & 1;                 // chains with r == step (mod 2) sweep LO at this step
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
            }
            m->launches += launched;
            if (l2k.debug)
                fprintf(stderr, "[qumcmc] hop %lld: %d sweep launches enqueued in %.3f ms (group %d chains x %d streams)\n",
                        (long long)hop, launched,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(),
                        l2k.group, l2k.streams);
            if (m->profile) {
                QMC_CUDA_TRY(cudaEventRecord(e1, st));
                w->ev.push_back(e0);
                w->ev.push_back(e1);
                // algorithmic bytes of this hop: 2 * 2^n * 8 B * r per chain
                m->prof_bytes += 2.0 * (double)N * 8.0 * sum_r;
                // bytes actually moved: K = groups.size() (1 on the two-group path) passes per layer, the first sweep of a
                // proposal write-only, the last read-only => 16 B * 2^n * K * (r - 1) per proposal
                m->prof_moved += 16.0 * (double)N * (double)(generic ? groups.size() : 1) * (sum_r - (double)nc);
                m->prof_launches += launched;
            }
            if (mode == QMC_MODE_PROBS) {
                probs_normalize_kernel<<<148 * 4, 256, 0, st>>>(n, w->segsum, a.probs_out, a.amps_out);
                m->launches++;
            } else {
                if (generic) {
                    coarse_sums_kernel<<<dim3((unsigned)(N >> (6 + COARSE_BITS)), (unsigned)nc), 256, 0, st>>>(w->segsum, n, w->coarse);
                    ++m->launches;
                }
                sample_accept_kernel<<<(unsigned)nc, NT, TILE * 8, st>>>(a, h, generic ? w->coarse : nullptr);
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


// ==========================================================================================
// Sharded statevector (qmc_shard_*): rank R of 2^g holds the 2^nl amplitudes whose top g index bits equal R.
// Everything below works in PHYSICAL coordinates: position a < nl is bit a of the local index, position nl + j is
// bit j of the rank.  An all-to-all on equal chunks exchanges the top g local positions with the rank positions;
// the library never un-swaps, it only tracks which logical qubit sits at each position (layout 0 = identity,
// layout 1 = after an odd number of exchanges) and feeds the kernels permuted parameters: couplings Jq', hq',
// mixer multipliers tq' and the start state cur'.  (popcount -- the storage convention -- is permutation invariant.)
// One op = one kernel launch over the local shard; the caller (qumcmc_b200/distributed.py) strings ops and
// exchanges together, see sharded_plan().
// ==========================================================================================
struct qmc_shard {
    qmc_model *m = nullptr;
    int rank = 0, g = 0, nl = 0, n = 0;
    ChainParams *params = nullptr;  // one record (device)
    int *order = nullptr;           // {0}
    float *segsum = nullptr;        // [2^(nl-6)]
    double *total = nullptr;        // scratch for the norm
    double *dJq[2] = {nullptr, nullptr}, *dhq[2] = {nullptr, nullptr};
    std::vector<double> Jq[2], hq[2];
    std::vector<int> perm[2];       // perm[layout][position] = logical qubit
    std::vector<HiGroup> groups[2];
    std::vector<int> acts;
    double fx_scale = 1.0;
    int k_fx = 0;
    ChainParams host_cp[2];         // physical-coordinate parameters of the current proposal, per layout
    float2 **d_peers[2] = {nullptr, nullptr};  // device tables: pointer to buffer `which` of every rank
    std::vector<void *> h_peers[2];            // the same addresses on the host (copy-engine exchange)
    int xchg_ctas = 0;              // > 0: exchange passes of HI6 groups run as a persistent kernel on this many CTAs
    int r = 0;
};

namespace {

__global__ void shard_norm_kernel(const float *__restrict__ seg, int64_t nseg, double *__restrict__ total) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nseg; i += (int64_t)gridDim.x * blockDim.x)
        s += (double)seg[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(total, red[0]);
}

// Inverse CDF over the local shard: the amplitude index at which the running sum of |a|^2 (segment sums in
// segsum, segment = 64 amplitudes) passes `target`.  Phase 1: per-block partial sums of contiguous segment ranges.
constexpr int SEL_BLOCKS = 1024;
__global__ void shard_block_sums_kernel(const float *__restrict__ seg, int64_t nseg, double *__restrict__ bsum) {
    __shared__ double red[256];
    const int64_t per = (nseg + SEL_BLOCKS - 1) / SEL_BLOCKS;
    const int64_t b0 = blockIdx.x * per, b1 = b0 + per < nseg ? b0 + per : nseg;
    double s = 0.0;
    for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) s += (double)seg[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[blockIdx.x] = red[0];
}
// Phase 2 (one thread): block, then segment inside the block; writes the segment and the remainder.
__global__ void shard_pick_segment_kernel(const float *__restrict__ seg, int64_t nseg, const double *__restrict__ bsum,
                                          double target, long long *__restrict__ sel_seg, double *__restrict__ rem) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t per = (nseg + SEL_BLOCKS - 1) / SEL_BLOCKS;
    const int nblk = (int)((nseg + per - 1) / per);
    double c = 0.0;
    int blk = nblk - 1;
    for (int b = 0; b < nblk; ++b) {
        if (target < c + bsum[b] || b == nblk - 1) {
            blk = b;
            break;
        }
        c += bsum[b];
    }
    const int64_t b0 = blk * per, b1 = b0 + per < nseg ? b0 + per : nseg;
    int64_t sel = b1 - 1;
    double before = c;
    for (int64_t i = b0; i < b1; ++i) {
        const double x = (double)seg[i];
        if (target < c + x || i == b1 - 1) {  // rounding tail: last segment of the block
            sel = i;
            before = c;
            break;
        }
        c += x;
    }
    *sel_seg = sel;
    *rem = target - before;
}
// Phase 3: recompute the final amplitudes of the tile that holds the segment, resolve the index inside it.
__global__ void __launch_bounds__(NT) shard_resolve_kernel(const SweepArgs a, const long long *__restrict__ sel_seg,
                                                           const double *__restrict__ rem_p, unsigned long long *out) {
    extern __shared__ __align__(16) float2 sm[];
    __shared__ ChainParams cp_s;
    const int tid = threadIdx.x;
    stage_params(cp_s, a.params, tid);
    const ChainParams &cp = cp_s;
    const int64_t sg = *sel_seg;
    const int64_t tile = sg / NT;
    const int owner = (int)(sg % NT);
    A64 v[64];
    if (cp.r == 1) lo_final_tile<NT, false, true>(v, sm, a.state, tile, tid, a.n_total, cp, a.gbits);
    else lo_final_tile<NT, false, false>(v, sm, a.state, tile, tid, a.n_total, cp, a.gbits);
    if (tid == owner) {
        const double rem = *rem_p;
        double c = 0.0;
        int selj = 63;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            c += (double)norm2(v[j]);
            if (selj == 63 && rem < c) selj = j;
        }
        *out = ((unsigned long long)tile << TILE_BITS) | ((unsigned long long)owner << 6) | (unsigned long long)selj;
    }
}

}  // namespace

static void shard_fill_params(qmc_shard *sh, uint64_t s, double gamma, int r, double dt, int grp) {
    const qmc_model *m = sh->m;
    const int n = sh->n;
    double tq[QMC_MAX_SPINS];
    for (int q = 0; q < QMC_MAX_SPINS; ++q) tq[q] = 0.0;
    double scale = 1.0;
    for (int q = 0; q < n; ++q) {  // theta_q = gamma * (summed weight of the group's terms on qubit q) * dt
        const double wq = m->qweight[(size_t)grp * n + q];
        if (wq == 0.0) continue;
        const double th = gamma * wq * dt;
        tq[q] = sin(th) / cos(th);
        scale *= cos(th);
    }
    const double cprime = (1.0 - gamma) * m->alpha * dt * 0.15915494309189533577;
    const int rr = r < 1 ? 1 : r;
    const int active = m->phase_active && rr > 1 && cprime != 0.0;
    int mm = 0;
    if (active) {
        mm = (int)floor(log2(2147483647.0 / fabs(cprime)));
        if (mm > 40) mm = 40;
        if (mm < 0) mm = 0;
    }
    for (int L = 0; L < 2; ++L) {
        ChainParams &cp = sh->host_cp[L];
        memset(&cp, 0, sizeof(cp));
        unsigned long long cur = 0ULL;
        for (int a = 0; a < n; ++a) {
            const int lq = sh->perm[L][a];
            cp.tq[a] = (float)tq[lq];
            if ((s >> lq) & 1ULL) cur |= 1ULL << a;
        }
        cp.cur = cur;
        cp.r = rr;
        cp.group = grp;
        cp.scale = (float)scale;
        cp.active = active;
        cp.cfx = active ? (int)rint(ldexp(cprime, mm)) : 0;
        cp.shift = sh->k_fx + mm - 32;
    }
    sh->r = rr;
}

extern "C" int qmc_shard_create(qmc_shard **out, qmc_model *m, int rank, int log2_world) {
    QMC_REQUIRE(out && m, QMC_ERR_BADARG, "qmc_shard_create: null argument");
    const int n = m->n, g = log2_world, nl = n - g;
    QMC_REQUIRE(g >= 1 && g <= 6 && rank >= 0 && rank < (1 << g), QMC_ERR_BADARG, "qmc_shard_create: rank %d of 2^%d", rank, g);
    QMC_REQUIRE(nl >= 18 && nl <= QMC_MAX_LOCAL_SPINS, QMC_ERR_UNSUPPORTED,
                "sharded statevector: %d local qubits outside [18, %d] (n=%d over 2^%d ranks)", nl, QMC_MAX_LOCAL_SPINS, n, g);
    QMC_REQUIRE(m->all_one_body, QMC_ERR_UNSUPPORTED, "sharded statevector: one-body X mixers only");
    QMC_CUDA_TRY(cudaSetDevice(m->device));
    qmc_shard *sh = new qmc_shard();
    sh->m = m; sh->rank = rank; sh->g = g; sh->nl = nl; sh->n = n;
    // logical couplings in qubit order: qubit a = spin n-1-a (quantum_mcmc_routines.py:75,87)
    std::vector<double> Jl((size_t)n * n), hl(n);
    double dmax = 0.0;
    for (int a = 0; a < n; ++a) {
        hl[a] = m->host_ch[n - 1 - a];
        dmax += fabs(hl[a]);
        for (int b = 0; b < n; ++b) {
            Jl[(size_t)a * n + b] = m->host_cJ[(size_t)(n - 1 - a) * n + (n - 1 - b)];
            if (b > a) dmax += fabs(Jl[(size_t)a * n + b]);
        }
    }
    if (dmax < 1e-300) dmax = 1.0;
    sh->k_fx = (int)floor(log2(2147483647.0 / (dmax * 1.0000001))) - 1;
    if (sh->k_fx > 30) sh->k_fx = 30;
    sh->fx_scale = ldexp(1.0, sh->k_fx);
    // group geometry (positions): full HI8 groups from the top, the partial group lowest; LO = positions 0..11
    std::vector<int> kinds, ps;
    {
        const int rcg = hi_group_geometry(nl, kinds, ps, sh->acts);
        if (rcg != QMC_OK) {
            delete sh;
            return rcg;
        }
    }
    const int K = (int)kinds.size();
    if (g > sh->acts[0] || g > 6) {
        delete sh;
        qmc_set_error("sharded statevector: 2^%d ranks need a top group of >= %d qubits", g, g);
        return QMC_ERR_UNSUPPORTED;
    }
    for (int L = 0; L < 2; ++L) {
        sh->perm[L].resize(n);
        for (int a = 0; a < n; ++a) sh->perm[L][a] = a;
        if (L == 1)
            for (int j = 0; j < g; ++j) std::swap(sh->perm[L][nl - g + j], sh->perm[L][nl + j]);
        sh->Jq[L].assign((size_t)n * n, 0.0);
        sh->hq[L].assign(n, 0.0);
        for (int a = 0; a < n; ++a) {
            sh->hq[L][a] = hl[sh->perm[L][a]];
            for (int b = 0; b < n; ++b) sh->Jq[L][(size_t)a * n + b] = Jl[(size_t)sh->perm[L][a] * n + sh->perm[L][b]];
        }
        QMC_CUDA_TRY(cudaMalloc(&sh->dJq[L], sizeof(double) * (size_t)n * n));
        QMC_CUDA_TRY(cudaMalloc(&sh->dhq[L], sizeof(double) * (size_t)n));
        QMC_CUDA_TRY(cudaMemcpy(sh->dJq[L], sh->Jq[L].data(), sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
        QMC_CUDA_TRY(cudaMemcpy(sh->dhq[L], sh->hq[L].data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
        for (int k = 0; k < K; ++k) {
            HiGroup G;
            G.kind = kinds[k];
            G.act = sh->acts[k];
            Hi6Args &ga = G.args;
            ga.p = ps[k];
            const int pos[6] = {ga.p, ga.p + 1, ga.p + 2, ga.p + 3, ga.p + 4, ga.p + 5};
            gray_increments(sh->Jq[L], n, sh->fx_scale, pos, ga.dr);
            ga.Jq = sh->dJq[L];
            ga.hq = sh->dhq[L];
            ga.fx_scale = sh->fx_scale;
            ga.cta_tab = nullptr;
            ga.thr_tab = nullptr;
            int rc = QMC_OK;
            if (k == 0)  // the plan runs FIRST / MID (the passes with a phase layer) on the top group only
                rc = build_group_tables(G, nl, n, (unsigned long long)rank << nl, sh->Jq[L], 0);
            sh->groups[L].push_back(G);
            if (rc != QMC_OK) {
                qmc_shard_destroy(sh);
                return rc;
            }
        }
    }
    QMC_CUDA_TRY(cudaMalloc(&sh->params, sizeof(ChainParams)));
    QMC_CUDA_TRY(cudaMalloc(&sh->order, sizeof(int)));
    QMC_CUDA_TRY(cudaMemset(sh->order, 0, sizeof(int)));
    QMC_CUDA_TRY(cudaMalloc(&sh->segsum, sizeof(float) * ((size_t)1 << (nl - 6))));
    QMC_CUDA_TRY(cudaMalloc(&sh->total, sizeof(double) * (SEL_BLOCKS + 4)));
    {
        static DevMask attr = 0;
        int rca = set_smem_once(shard_resolve_kernel, attr);
        if (rca != QMC_OK) {
            qmc_shard_destroy(sh);
            return rca;
        }
    }
    *out = sh;
    return QMC_OK;
}

extern "C" int qmc_shard_destroy(qmc_shard *sh) {
    if (!sh) return QMC_OK;
    cudaFree(sh->params); cudaFree(sh->order); cudaFree(sh->segsum); cudaFree(sh->total);
    for (int L = 0; L < 2; ++L) {
        cudaFree(sh->dJq[L]); cudaFree(sh->dhq[L]); cudaFree(sh->d_peers[L]);
        for (auto &G : sh->groups[L]) free_group_tables(G);
    }
    delete sh;
    return QMC_OK;
}

extern "C" int qmc_shard_num_groups(const qmc_shard *sh) { return sh ? (int)sh->acts.size() : 0; }

// Overlap mode: exchange passes (qmc_shard_op_exchange / qmc_shard_op_chunk with dst_which >= 0) on register-only HI6
// groups run as a persistent kernel on `ctas` CTAs (0 = one CTA per tile, the default).  Results are bit-identical.
extern "C" int qmc_shard_set_exchange_ctas(qmc_shard *sh, int ctas) {
    QMC_REQUIRE(sh && ctas >= 0, QMC_ERR_BADARG, "qmc_shard_set_exchange_ctas: bad argument");
    sh->xchg_ctas = ctas;
    return QMC_OK;
}
// 1 when the pass before the exchange (SINGLE on the lowest register group) has a persistent variant
extern "C" int qmc_shard_exchange_can_persist(const qmc_shard *sh) {
    if (!sh || sh->groups[0].size() < 2) return 0;
    return sh->groups[0].back().kind == 6 ? 1 : 0;
}

extern "C" int qmc_shard_begin(qmc_shard *sh, uint64_t s, double gamma, int r, double delta_time, int group) {
    QMC_REQUIRE(sh, QMC_ERR_BADARG, "qmc_shard_begin: null handle");
    QMC_REQUIRE(group >= 0 && group < sh->m->n_groups, QMC_ERR_BADARG, "qmc_shard_begin: group %d of %d", group, sh->m->n_groups);
    QMC_REQUIRE(delta_time > 0.0 && r >= 1, QMC_ERR_BADARG, "qmc_shard_begin: r=%d, delta_time=%g", r, delta_time);
    QMC_REQUIRE(sh->n >= 64 || s < (1ULL << sh->n), QMC_ERR_BADARG, "qmc_shard_begin: state out of range");
    shard_fill_params(sh, s, gamma, r, delta_time, group);
    return QMC_OK;
}

// One local pass.  op: 0 FIRST(top group: analytic M|s>, P, M_top) | 1 MID(top group: M_top restricted to the
// `pre` swapped-in qubits, P, M_top) | 2 SINGLE(group k, its active qubits; k == 0 with pre > 0: only the `pre`
// swapped-in qubits) | 3 LO_SINGLE | 4 LO_FINAL (segment sums; r == 1: of the analytic product state).
static int shard_op_impl(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, int dst_which, void *stream,
                         int chunk_bits = 0, int64_t chunk = 0) {
    QMC_NVTX(dst_which >= 0 ? "shard pass + exchange" : "shard pass");
    QMC_REQUIRE(sh && state_dev, QMC_ERR_BADARG, "qmc_shard_op: null argument");
    QMC_REQUIRE(dst_which < 0 || (dst_which < 2 && sh->d_peers[dst_which]), QMC_ERR_BADARG,
                "qmc_shard_op_exchange: peer table %d not set (qmc_shard_set_peers)", dst_which);
    QMC_REQUIRE(dst_which < 0 || op != 4, QMC_ERR_BADARG, "qmc_shard_op_exchange: LO_FINAL does not write the state");
    QMC_REQUIRE(layout == 0 || layout == 1, QMC_ERR_BADARG, "qmc_shard_op: layout %d", layout);
    const int K = (int)sh->acts.size();
    QMC_REQUIRE(k >= 0 && k < K, QMC_ERR_BADARG, "qmc_shard_op: group %d of %d", k, K);
    QMC_CUDA_TRY(cudaSetDevice(sh->m->device));
    cudaStream_t st = (cudaStream_t)stream;
    QMC_CUDA_TRY(cudaMemcpyAsync(sh->params, &sh->host_cp[layout], sizeof(ChainParams), cudaMemcpyHostToDevice, st));
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.n = sh->nl; a.n_total = sh->n; a.gbits = (unsigned long long)sh->rank << sh->nl;
    a.state = reinterpret_cast<float2 *>(state_dev);
    a.params = sh->params;
    a.segsum = sh->segsum;
    a.pf_dist = 0;
    if (chunk_bits > 0) {
        // A pass that leaves the top chunk_bits local positions alone (LO_SINGLE, SINGLE on a group below them) is the
        // same pass on each of the 2^chunk_bits contiguous sub-blocks: the chunk becomes a smaller "shard" whose fixed
        // top bits join the global bits.  Lets the caller pipeline local passes against the exchange chunk by chunk.
        QMC_REQUIRE(op == 2 || op == 3, QMC_ERR_BADARG, "chunked passes: SINGLE / LO_SINGLE only");
        const int top = op == 3 ? TILE_BITS : sh->groups[layout][k].args.p + (sh->groups[layout][k].kind == 8 ? 8 : 6);
        QMC_REQUIRE(k > 0 || op == 3, QMC_ERR_BADARG, "chunked passes: not on the top group");
        QMC_REQUIRE(sh->nl - chunk_bits >= top && sh->nl - chunk_bits >= TILE_BITS, QMC_ERR_BADARG,
                    "chunked pass: %d chunk bits reach into the pass's qubits", chunk_bits);
        QMC_REQUIRE(chunk >= 0 && chunk < ((int64_t)1 << chunk_bits), QMC_ERR_BADARG, "chunk %lld of 2^%d", (long long)chunk, chunk_bits);
        a.n = sh->nl - chunk_bits;
        a.state += chunk << a.n;
        a.gbits |= (unsigned long long)chunk << a.n;
    }
    if (dst_which >= 0) {
        a.xchg_peers = sh->d_peers[dst_which];
        a.xchg_shift = sh->nl - sh->g;
        a.xchg_rank = sh->rank;
        a.xchg_g = sh->g;
        a.xchg_tile_bits = a.n - TILE_BITS;
        a.xchg_idx0 = chunk_bits > 0 ? chunk << a.n : 0;
        if (chunk_bits > 0 && chunk_bits < sh->g) {
            qmc_set_error("chunked exchange pass: chunk_bits >= log2(world) required");
            return QMC_ERR_BADARG;
        }
        if (chunk_bits > 0) a.xchg_tile_bits = 0;  // one destination per chunk: no renumbering
    }
    const int64_t tiles = (int64_t)1 << (a.n - TILE_BITS);
    const HiGroup &G = sh->groups[layout][k];
    int rc = QMC_OK;
    switch (op) {
        case 0:
            QMC_REQUIRE(k == 0, QMC_ERR_BADARG, "qmc_shard_op: FIRST runs on the top group");
            rc = launch_group<HI6_FIRST>(G, 0, a, sh->order, 0, tiles, 1, st);
            break;
        case 1:
            QMC_REQUIRE(k == 0 && pre == sh->g, QMC_ERR_BADARG, "qmc_shard_op: MID runs on the top group with pre = log2(world)");
            rc = launch_group<HI6_MID>(G, pre, a, sh->order, 0, tiles, 1, st);
            break;
        case 2:
            if (dst_which >= 0 && sh->xchg_ctas > 0 && G.kind == 6 && k > 0 && (int64_t)sh->xchg_ctas < tiles) {
                const unsigned ctas = (unsigned)sh->xchg_ctas;
                switch (G.act) {  // persistent exchange pass (overlap mode)
                    case 1: hi6_xchg_kernel<1><<<dim3(ctas, 1), NT, 0, st>>>(a, G.args, sh->order, 0, (unsigned)tiles); break;
                    case 2: hi6_xchg_kernel<2><<<dim3(ctas, 1), NT, 0, st>>>(a, G.args, sh->order, 0, (unsigned)tiles); break;
                    case 3: hi6_xchg_kernel<3><<<dim3(ctas, 1), NT, 0, st>>>(a, G.args, sh->order, 0, (unsigned)tiles); break;
                    case 4: hi6_xchg_kernel<4><<<dim3(ctas, 1), NT, 0, st>>>(a, G.args, sh->order, 0, (unsigned)tiles); break;
                    case 5: hi6_xchg_kernel<5><<<dim3(ctas, 1), NT, 0, st>>>(a, G.args, sh->order, 0, (unsigned)tiles); break;
                    default: hi6_xchg_kernel<6><<<dim3(ctas, 1), NT, 0, st>>>(a, G.args, sh->order, 0, (unsigned)tiles); break;
                }
                break;
            }
            rc = launch_group<HI6_SINGLE>(G, k == 0 ? pre : 0, a, sh->order, 0, tiles, 1, st);
            break;
        case 3:
            rc = launch_lo<LO_SINGLE>(SweepTune(), a, sh->order, 0, tiles, 1, st);
            break;
        case 4:
            rc = sh->r == 1 ? launch_lo<LO_FINAL1>(SweepTune(), a, sh->order, 0, tiles, 1, st)
                            : launch_lo<LO_FINAL>(SweepTune(), a, sh->order, 0, tiles, 1, st);
            break;
        default:
            qmc_set_error("qmc_shard_op: unknown op %d", op);
            return QMC_ERR_BADARG;
    }
    if (rc != QMC_OK) return rc;
    sh->m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

extern "C" int qmc_shard_op(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, void *stream) {
    return shard_op_impl(sh, op, k, pre, layout, state_dev, -1, stream);
}

// Device-accessible address of buffer `which` (0/1) of every rank, this rank's own included (peer mappings from
// CUDA IPC / torch shared CUDA tensors; on one GPU plain pointers).  Host array of 2^g pointers.
extern "C" int qmc_shard_set_peers(qmc_shard *sh, int which, void *const *ptrs_host) {
    QMC_REQUIRE(sh && ptrs_host && (which == 0 || which == 1), QMC_ERR_BADARG, "qmc_shard_set_peers: bad argument");
    QMC_CUDA_TRY(cudaSetDevice(sh->m->device));
    const size_t bytes = sizeof(void *) * ((size_t)1 << sh->g);
    if (!sh->d_peers[which]) QMC_CUDA_TRY(cudaMalloc(&sh->d_peers[which], bytes));
    QMC_CUDA_TRY(cudaMemcpy(sh->d_peers[which], ptrs_host, bytes, cudaMemcpyHostToDevice));
    sh->h_peers[which].assign(ptrs_host, ptrs_host + ((size_t)1 << sh->g));
    return QMC_OK;
}

// Exchange of one chunk by the COPY ENGINES: the 2^(nl - chunk_bits) contiguous amplitudes of chunk `chunk` of this rank's
// shard go to buffer `dst_which` of the rank that owns them after the qubit swap -- one cudaMemcpyAsync into the peer
// mapping (NVLink P2P), no SM involved, so the HBM-bound local passes of the following chunks run beside it at full
// occupancy.  chunk_bits >= log2(world): one destination rank per chunk (its top log2(world) bits).
extern "C" int qmc_shard_copy_chunk(qmc_shard *sh, const void *state_dev, int dst_which, int chunk_bits, int64_t chunk,
                                    void *stream) {
    QMC_REQUIRE(sh && state_dev && (dst_which == 0 || dst_which == 1), QMC_ERR_BADARG, "qmc_shard_copy_chunk: bad argument");
    QMC_REQUIRE(!sh->h_peers[dst_which].empty(), QMC_ERR_BADARG, "qmc_shard_copy_chunk: peer table %d not set (qmc_shard_set_peers)", dst_which);
    QMC_REQUIRE(chunk_bits >= sh->g && chunk_bits <= sh->nl - TILE_BITS, QMC_ERR_BADARG, "qmc_shard_copy_chunk: chunk_bits %d", chunk_bits);
    QMC_REQUIRE(chunk >= 0 && chunk < ((int64_t)1 << chunk_bits), QMC_ERR_BADARG, "chunk %lld of 2^%d", (long long)chunk, chunk_bits);
    QMC_CUDA_TRY(cudaSetDevice(sh->m->device));
    const int64_t len = (int64_t)1 << (sh->nl - chunk_bits);
    const int64_t src0 = chunk << (sh->nl - chunk_bits);
    const int dest = (int)(chunk >> (chunk_bits - sh->g));
    const int64_t low = src0 & (((int64_t)1 << (sh->nl - sh->g)) - 1);
    const int64_t dst0 = low | ((int64_t)sh->rank << (sh->nl - sh->g));
    const float2 *src = reinterpret_cast<const float2 *>(state_dev) + src0;
    float2 *dst = reinterpret_cast<float2 *>(sh->h_peers[dst_which][dest]) + dst0;
    QMC_CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(float2) * (size_t)len, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return QMC_OK;
}

// The pass `op` with its stores redirected into buffer `dst_which` of the owning ranks: local pass + all-to-all in
// one kernel.  The caller puts a barrier across ranks between this call and the first read of the destination
// buffers (and nothing else: the buffers alternate, so the previous contents of `dst_which` were last read before
// the previous barrier).
extern "C" int qmc_shard_op_exchange(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, int dst_which,
                                     void *stream) {
    QMC_REQUIRE(dst_which == 0 || dst_which == 1, QMC_ERR_BADARG, "qmc_shard_op_exchange: dst_which %d", dst_which);
    return shard_op_impl(sh, op, k, pre, layout, state_dev, dst_which, stream);
}

// One chunk (the `chunk`-th of 2^chunk_bits contiguous sub-blocks of the shard) of a SINGLE / LO_SINGLE pass, local
// (dst_which < 0) or fused with the exchange (dst_which = 0 / 1; needs chunk_bits >= log2(world), i.e. one destination
// rank per chunk).  The caller orders the chunks (destinations rotated by rank) and overlaps streams.
extern "C" int qmc_shard_op_chunk(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, int dst_which,
                                  int chunk_bits, int64_t chunk, void *stream) {
    QMC_REQUIRE(dst_which >= -1 && dst_which <= 1, QMC_ERR_BADARG, "qmc_shard_op_chunk: dst_which %d", dst_which);
    QMC_REQUIRE(chunk_bits >= 1, QMC_ERR_BADARG, "qmc_shard_op_chunk: chunk_bits %d", chunk_bits);
    return shard_op_impl(sh, op, k, pre, layout, state_dev, dst_which, stream, chunk_bits, chunk);
}

// sum of |a|^2 over the local shard (after LO_FINAL) -> total_dev[0]
extern "C" int qmc_shard_norm(qmc_shard *sh, double *total_dev, void *stream) {
    QMC_REQUIRE(sh && total_dev, QMC_ERR_BADARG, "qmc_shard_norm: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    QMC_CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(double), st));
    shard_norm_kernel<<<148 * 4, 256, 0, st>>>(sh->segsum, (int64_t)1 << (sh->nl - 6), total_dev);
    sh->m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

// Owner rank only: local index (physical -> logical) at which the local running sum passes `target`
// (0 <= target < local total).  index_dev[0] = LOGICAL basis index of the measured bitstring.
extern "C" int qmc_shard_select(qmc_shard *sh, void *state_dev, double target, int layout, uint64_t *index_dev, void *stream) {
    QMC_REQUIRE(sh && state_dev && index_dev, QMC_ERR_BADARG, "qmc_shard_select: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nseg = (int64_t)1 << (sh->nl - 6);
    double *bsum = sh->total;  // [SEL_BLOCKS], then sel_seg, rem, out
    long long *sel_seg = reinterpret_cast<long long *>(sh->total + SEL_BLOCKS);
    double *rem = sh->total + SEL_BLOCKS + 1;
    unsigned long long *phys = reinterpret_cast<unsigned long long *>(sh->total + SEL_BLOCKS + 2);
    QMC_CUDA_TRY(cudaMemcpyAsync(sh->params, &sh->host_cp[layout], sizeof(ChainParams), cudaMemcpyHostToDevice, st));
    shard_block_sums_kernel<<<SEL_BLOCKS, 256, 0, st>>>(sh->segsum, nseg, bsum);
    shard_pick_segment_kernel<<<1, 32, 0, st>>>(sh->segsum, nseg, bsum, target, sel_seg, rem);
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.n = sh->nl; a.n_total = sh->n; a.gbits = (unsigned long long)sh->rank << sh->nl;
    a.state = reinterpret_cast<float2 *>(state_dev);
    a.params = sh->params;
    shard_resolve_kernel<<<1, NT, TILE * 8, st>>>(a, sel_seg, rem, phys);
    sh->m->launches += 3;
    QMC_CUDA_TRY(cudaGetLastError());
    unsigned long long hphys = 0ULL;
    QMC_CUDA_TRY(cudaMemcpyAsync(&hphys, phys, sizeof(hphys), cudaMemcpyDeviceToHost, st));
    QMC_CUDA_TRY(cudaStreamSynchronize(st));
    const unsigned long long full = hphys | ((unsigned long long)sh->rank << sh->nl);
    unsigned long long logical = 0ULL;
    for (int a2 = 0; a2 < sh->n; ++a2)
        if ((full >> a2) & 1ULL) logical |= 1ULL << sh->perm[layout][a2];
    QMC_CUDA_TRY(cudaMemcpyAsync(index_dev, &logical, sizeof(logical), cudaMemcpyHostToDevice, st));
    QMC_CUDA_TRY(cudaStreamSynchronize(st));
    return QMC_OK;
}

extern "C" int qmc_shard_final_probs(qmc_shard *sh, int layout, void *state_dev, float *probs_dev, void *stream) {
    QMC_REQUIRE(sh && state_dev && probs_dev, QMC_ERR_BADARG, "qmc_shard_final_probs: null argument");
    QMC_REQUIRE(layout == 0 || layout == 1, QMC_ERR_BADARG, "qmc_shard_final_probs: layout %d", layout);
    QMC_CUDA_TRY(cudaSetDevice(sh->m->device));
    cudaStream_t st = (cudaStream_t)stream;
    QMC_CUDA_TRY(cudaMemcpyAsync(sh->params, &sh->host_cp[layout], sizeof(ChainParams), cudaMemcpyHostToDevice, st));
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.n = sh->nl; a.n_total = sh->n; a.gbits = (unsigned long long)sh->rank << sh->nl;
    a.state = reinterpret_cast<float2 *>(state_dev);
    a.params = sh->params;
    a.segsum = sh->segsum;
    a.probs_out = probs_dev;
    const int64_t tiles = (int64_t)1 << (sh->nl - TILE_BITS);
    int rc = sh->r == 1 ? launch_lo<LO_FINAL1>(SweepTune(), a, sh->order, 0, tiles, 1, st)
                        : launch_lo<LO_FINAL>(SweepTune(), a, sh->order, 0, tiles, 1, st);
    if (rc != QMC_OK) return rc;
    sh->m->launches++;
    QMC_CUDA_TRY(cudaGetLastError());
    return QMC_OK;
}

// ------------------------------------------------------------------------------------------
// CUDA IPC for the fused exchange.  A peer's state buffer has to be mapped into THIS device's context (a mapping
// opened under the owner's device index, which is what torch's shared-CUDA-tensor machinery does, is not reachable
// from kernels running on this device).  Export: handle of the allocation that contains dev_ptr + the offset inside
// it (the caller's allocator may sub-allocate); open: map under `device` with lazy peer access; every distinct
// allocation is opened once per process and cached.
// ------------------------------------------------------------------------------------------
typedef int (*cuMemGetAddressRange_fn)(unsigned long long *pbase, size_t *psize, unsigned long long dptr);

extern "C" int qmc_ipc_export(int device, void *dev_ptr, void *handle_out, uint64_t *offset_out) {
    QMC_REQUIRE(dev_ptr && handle_out && offset_out, QMC_ERR_BADARG, "qmc_ipc_export: null argument");
    QMC_CUDA_TRY(cudaSetDevice(device));
    static cuMemGetAddressRange_fn get_range = nullptr;
    if (!get_range) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        QMC_CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &qr));
        QMC_REQUIRE(fp && qr == cudaDriverEntryPointSuccess, QMC_ERR_CUDA, "cuMemGetAddressRange entry point not found");
        get_range = reinterpret_cast<cuMemGetAddressRange_fn>(fp);
    }
    unsigned long long base = 0;
    size_t size = 0;
    const int drc = get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
    QMC_REQUIRE(drc == 0, QMC_ERR_CUDA, "cuMemGetAddressRange -> %d", drc);
    cudaIpcMemHandle_t h;
    QMC_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t)base)));
    static_assert(sizeof(h) == QMC_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = (uint64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
    return QMC_OK;
}

static std::vector<std::pair<std::string, void *>> g_ipc_open;  // (device:handle bytes) -> mapped base

extern "C" int qmc_ipc_open(int device, const void *handle, uint64_t offset, void **ptr_out) {
    QMC_REQUIRE(handle && ptr_out, QMC_ERR_BADARG, "qmc_ipc_open: null argument");
    QMC_CUDA_TRY(cudaSetDevice(device));
    std::string key(reinterpret_cast<const char *>(handle), QMC_IPC_HANDLE_BYTES);
    key.push_back((char)device);
    for (auto &e : g_ipc_open)
        if (e.first == key) {
            *ptr_out = static_cast<char *>(e.second) + offset;
            return QMC_OK;
        }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *base = nullptr;
    QMC_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    g_ipc_open.emplace_back(key, base);
    *ptr_out = static_cast<char *>(base) + offset;
    return QMC_OK;
}

// Unmap everything qmc_ipc_open mapped in this process (call before the owners free their buffers).
extern "C" int qmc_ipc_close_all(void) {
    for (auto &e : g_ipc_open) {
        cudaSetDevice((int)e.first.back());
        cudaIpcCloseMemHandle(e.second);
    }
    g_ipc_open.clear();
    return QMC_OK;
}

// cudaDeviceEnablePeerAccess(peer) on `device` (idempotent): the fused exchange stores to peer-mapped buffers
extern "C" int qmc_enable_peer_access(int device, int peer) {
    if (device == peer) return QMC_OK;
    QMC_CUDA_TRY(cudaSetDevice(device));
    int can = 0;
    QMC_CUDA_TRY(cudaDeviceCanAccessPeer(&can, device, peer));
    QMC_REQUIRE(can, QMC_ERR_UNSUPPORTED, "device %d cannot access peer %d (no NVLink / P2P path)", device, peer);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return QMC_OK;
    }
    QMC_CUDA_TRY(e);
    return QMC_OK;
}
