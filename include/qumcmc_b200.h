/*
 * qumcmc_b200.h -- C-ABI of libqumcmc_b200.so, the B200-native (sm_100a) replacement for the
 * quantum-proposal hot path of pafloxy/quMCMC.
 *
 * The reference has no FFI of its own: its seam is the Python API that sits on top of the qulacs
 * pybind11 module.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - basis index idx = int(bitstring, 2); qubit q = bit q of idx; spin j (string position j from
 *     the left) is bit n-1-j; s_j = +1 for '1'.
 *   - All pointers named *_dev are device pointers on the model's device; *_host are host pointers.
 *     The caller owns every buffer.  The library owns only the handle and its scratch.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls taking device
 *     buffers are asynchronous on that stream; *_host variants synchronise before returning.
 *   - Return value: 0 ok; QMC_ERR_* (<0) otherwise, message via qmc_last_error() (thread-local).
 *     There is no CPU fallback: unsupported configurations return QMC_ERR_UNSUPPORTED.
 *   - One host thread per handle at a time.
 *
 * RNG stream contract (Philox4x32-10; identical in oracle/qmc_oracle.c)
 *   key = (seed lo32, seed hi32); counter = (chain lo32, chain hi32, hop, slot)
 *   slot 0: w0,w1 -> u_gamma (53 bit), w2 -> t = 2 + floor(w2*10/2^32), w3 -> u_group = w3/2^32
 *   slot 1: w0,w1 -> u_measure, w2,w3 -> u_accept        u53(a,b) = ((a>>5)*2^26 + (b>>6)) / 2^53
 *   initial state (when s0 is NULL): hop = 0xFFFFFFFF, slot 0: (w1<<32 | w0) & (2^n - 1)
 *   classical draws (qmc_classical_chains): slot 2: w0,w1 -> u_method, w2,w3 -> u_accept;
 *     slot 3: UniformProposals state = (w1<<32 | w0) & (2^n - 1); FixedWeightProposals(k): partial Fisher-Yates over the
 *     string positions 0..n-1, step i < k uses word (i & 3) of slot 3 + (i >> 2): j = i + floor(w * (n - i) / 2^32),
 *     swap(pos[i], pos[j]), flip index bit n-1-pos[i].
 *   gamma = g_lo + (g_hi-g_lo)*u_gamma;  r = max(1, floor(t/delta_time)).
 *   `chain` is the GLOBAL chain id, so results do not depend on how chains are sharded over GPUs.
 */
#ifndef QUMCMC_B200_H
#define QUMCMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMC_OK 0
#define QMC_ERR_BADARG (-1)
#define QMC_ERR_CUDA (-2)
#define QMC_ERR_NCCL (-3)
#define QMC_ERR_UNSUPPORTED (-4)

#define QMC_DTYPE_F32 0 /* complex64 statevector  */
#define QMC_DTYPE_F64 1 /* complex128 statevector */

#define QMC_MAX_SPINS 40       /* model limit; one GPU holds a statevector up to n = 34 (2^34 complex64 = 128 GiB) */
#define QMC_MAX_LOCAL_SPINS 34
#define QMC_GENERAL_MAX_SPINS 30 /* general path (any X-string mixer, complex64 / complex128): Hadamard-basis passes, no 2^n table */
#define QMC_SWEEP_F64_MAX_SPINS 20 /* complex128 sweep kernels (one-body mixers): 13 <= n <= 20 */
#define QMC_SMEM_MAX_SPINS_F32 13 /* whole statevector in shared memory, one CTA per chain */
#define QMC_SMEM_MAX_SPINS_F64 12

typedef struct qmc_model qmc_model;

int qmc_version(void);
const char *qmc_last_error(void);

/* IsingEnergyFunction.__init__ (qumcmc/energy_models.py:22-52).  J, circJ: n*n row-major host
 * arrays; h, circh: n.  circJ/circh NULL = J/h.  alpha is model.alpha (:41-47), computed by the
 * caller.  Uploads the model and precomputes the phase table D(idx) of fn_qckt_problem_half
 * (qumcmc/quantum_mcmc_routines.py:54-91) when n > the shared-memory limits. */
int qmc_model_create(qmc_model **out, int n, const double *J_host, const double *h_host,
                     const double *circJ_host, const double *circh_host, double alpha, int device);
int qmc_model_destroy(qmc_model *m);

/* Mixer.get_mixer_circuit (qumcmc/mixers.py:8-18, 34-52, 89-105, 124-130, 149-151) as data:
 * group g owns terms [group_offsets[g], group_offsets[g+1]); term k = X-string over the qubits
 * set in masks[k], evolved by exp(-i * gamma * weights[k] * delta_time * X_mask).
 * group_prob NULL (or n_groups == 1): one coherent mixer (GenericMixer / CustomMixer /
 * CoherentMixerSum, weights = normalised w_m).  Otherwise IncoherentMixerSum: one group per
 * proposal with probability group_prob[g].  All host pointers. */
int qmc_mixer_set(qmc_model *m, int n_groups, const int *group_offsets, const uint64_t *masks,
                  const double *weights, const double *group_prob);

/* IsingEnergyFunction.get_energy (qumcmc/energy_models.py:124-144), batched:
 * E = 0.5 * s.(J s) + h.s in fixed row-major sequential fp64 order (bit-exact vs the oracle). */
int qmc_energies(qmc_model *m, const uint64_t *idx_dev, int64_t count, double *E_dev, void *stream);
int qmc_energies_host(qmc_model *m, const uint64_t *idx_host, int64_t count, double *E_host);

/* Exact_Sampling.get_boltzmann_distribution / run_exact_sampling
 * (qumcmc/energy_models.py:218-275, 277-297): all 2^n energies, log Z by a grid-wide
 * logsumexp, p = exp(-beta*E - logZ).  energies/probs nullable (index order). logZ: 1 double. */
int qmc_exact_sampling(qmc_model *m, double beta, double *logZ_dev, double *energies_dev,
                       double *probs_dev, void *stream);
int qmc_exact_sampling_host(qmc_model *m, double beta, double *logZ_host, double *energies_host,
                            double *probs_host);

/* Sharded enumeration (exact sampler over several GPUs / index ranges): (max, sum exp(x - max)) of
 * x = -beta*E over idx in [idx_begin, idx_begin+idx_count) -> lse_pair_dev[2]; ranks combine the
 * pairs (an all-gather of 16 bytes per rank = the "allreduce for Z").  energies_dev nullable
 * (idx_count doubles, range order).  qmc_exact_probs_range writes exp(-beta*E - logZ) for a range
 * given the global logZ. */
int qmc_exact_partial(qmc_model *m, double beta, uint64_t idx_begin, int64_t idx_count,
                      double *lse_pair_dev, double *energies_dev, void *stream);
int qmc_exact_probs_range(qmc_model *m, double beta, double logZ, uint64_t idx_begin, int64_t idx_count,
                          const double *energies_dev, double *probs_dev, void *stream);

/* Test hook: |<s'|U|s>|^2 for all s' with U = M (P M)^{r-1}, the circuit of
 * run_qmcmc_quantum_ckt (qumcmc/quantum_mcmc_routines.py:110-152) for fixed gamma and r and
 * mixer group `group`.  probs_dev: 2^n float (F32) or double (F64), index order.
 * amps_dev (nullable): 2^n interleaved complex of the same precision. */
int qmc_transition_probs(qmc_model *m, uint64_t s, double gamma, int r, double delta_time, int group,
                         int dtype, void *probs_dev, void *amps_dev, void *stream);

/* Test hook: one proposal per chain with explicit draws; replaces run_qmcmc_quantum_ckt
 * (:110-156) incl. QuantumState.sampling(1) (:153).  u_dev = measurement uniforms in [0,1);
 * group_dev nullable (0). */
int qmc_propose(qmc_model *m, int64_t n_chains, const uint64_t *s_in_dev, const double *gamma_dev,
                const int *r_dev, const double *u_dev, const int *group_dev, double delta_time,
                int dtype, uint64_t *s_out_dev, void *stream);

/* quantum_enhanced_mcmc (qumcmc/quantum_mcmc_routines.py:159-227) for n_chains independent
 * chains x n_hops hops: per hop draw (gamma, t, group), propose, get_energy, test_accept
 * (qumcmc/classical_mcmc_routines.py:28-41), record.  s0_dev NULL = Philox initial states.
 * chain_id0 = global id of chain 0 of this call; hop0 = index of the first hop (resume).
 * Outputs (all nullable except final_state_dev): proposals_dev/accepted_dev [n_chains*n_hops]
 * row-major (chain, hop) = the MCMCState fields (qumcmc/basic_utils.py:54-57);
 * final_state_dev [n_chains] = current state after the last hop; acc_count_dev [n_chains];
 * hamming_hist_dev [n_chains*(n+1)] = proposal Hamming distances w.r.t. the current state. */
int qmc_run_chains(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_dev,
                   uint64_t chain_id0, uint32_t hop0, double temperature, double gamma_lo,
                   double gamma_hi, double delta_time, uint64_t seed, int dtype,
                   uint64_t *proposals_dev, uint8_t *accepted_dev, uint64_t *final_state_dev,
                   int64_t *acc_count_dev, int64_t *hamming_hist_dev, void *stream);
/* Same with host buffers (copies inside): the reference-facing call. */
int qmc_run_chains_host(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_host,
                        uint64_t chain_id0, uint32_t hop0, double temperature, double gamma_lo,
                        double gamma_hi, double delta_time, uint64_t seed, int dtype,
                        uint64_t *proposals_host, uint8_t *accepted_host,
                        uint64_t *final_state_host, int64_t *acc_count_host,
                        int64_t *hamming_hist_host);

/* Visit histogram of the Markov chains on the device: Counter(markov_chain) of MCMCChain.get_accepted_dict
 * (qumcmc/basic_utils.py:117-131), summed over chains, as a dense int64[2^n] in index order (n <= 30).  While a
 * histogram is attached, every qmc_run_chains / qmc_classical_chains call (and their _host variants) on this handle adds
 * one count per chain and hop at the state after the hop, plus one at the initial state when hop0 == 0 (entry 0 of
 * markov_chain), so resumed runs (hop0 > 0) and repeated calls add up to the histogram of the whole run.  The buffer is
 * caller-owned and caller-zeroed; NULL detaches.  Ranks of a chain-sharded run sum their histograms with one
 * all-reduce (qumcmc_b200/distributed.py: gather_statistics). */
#define QMC_VISIT_HIST_MAX_SPINS 30
int qmc_set_visit_histogram(qmc_model *m, int64_t *visit_hist_dev);

/* classical_mcmc (qumcmc/classical_mcmc_routines.py:44-99) for n_chains independent chains x n_hops hops with the
 * proposal rules of qumcmc/classical_mixers.py:27-101.  The proposal method is a mixture (CombineProposals, :76-101)
 * of n_methods rules: kinds[i] UNIFORM (get_random_state), FIXED_WEIGHT (params[i] = bodyness: xor with k random
 * ones, basic_utils.py:383-388) or CUSTOM (params[i] = flip mask over index bits); probs NULL = equal weights.
 * kinds / params / probs are HOST arrays; the remaining arguments as in qmc_run_chains. */
#define QMC_MAX_CLASSICAL_METHODS 16
#define QMC_CLASSICAL_UNIFORM 0
#define QMC_CLASSICAL_FIXED_WEIGHT 1
#define QMC_CLASSICAL_CUSTOM 2
int qmc_classical_chains(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_dev,
                         uint64_t chain_id0, uint32_t hop0, double temperature, int n_methods,
                         const int *kinds, const uint64_t *params, const double *probs, uint64_t seed,
                         uint64_t *proposals_dev, uint8_t *accepted_dev, uint64_t *final_state_dev,
                         int64_t *acc_count_dev, int64_t *hamming_hist_dev, void *stream);
int qmc_classical_chains_host(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_host,
                              uint64_t chain_id0, uint32_t hop0, double temperature, int n_methods,
                              const int *kinds, const uint64_t *params, const double *probs, uint64_t seed,
                              uint64_t *proposals_host, uint8_t *accepted_host, uint64_t *final_state_host,
                              int64_t *acc_count_host, int64_t *hamming_hist_host);

/* Trajectory statistics of recorded chains (qumcmc/trajectory_processing.py:114-153,183-210,213-321), one pass per
 * chain on the device.  Inputs: s0_dev[C], proposals_dev / accepted_dev [C*H] as written by qmc_run_chains /
 * qmc_classical_chains.  Outputs, all nullable:
 *   markov_chain_dev [C*(H+1)]   state after every hop, entry 0 = initial state (basic_utils.py:84-115)
 *   energy_diff_dev  [C*H]       E(proposal) - E(current state) per hop ("energy" statistic, :235-263)
 *   running_mag_dev  [C*H]       mean magnetisation (#1 - #0) of markov_chain[0:k], k = 1..H (:183-210)
 *   hamming_acc_rej_dev [C*(n+1)*2]  proposals per Hamming distance to the current state: accepted, rejected (:279-301)
 *   running_kl_dev   [C*(H+1)]   vectoried_KL(boltzmann, counts of markov_chain[0:k+1] / (k+1)) (:137-151, natural log,
 *                                EPS = 1e-8, mask p > 1e-10); needs boltz_probs_dev[2^n] (index order) and H < 1e8. */
int qmc_trajectory_stats(qmc_model *m, int64_t n_chains, int64_t n_hops, const uint64_t *s0_dev,
                         const uint64_t *proposals_dev, const uint8_t *accepted_dev, const double *boltz_probs_dev,
                         uint64_t *markov_chain_dev, double *energy_diff_dev, double *running_mag_dev,
                         int64_t *hamming_acc_rej_dev, double *running_kl_dev, void *stream);

/* Spin moments of `count` states (the model side of the contrastive-divergence gradient, qumcmc/training.py:70-99):
 * exact counts over the states with mask_dev[k] != 0 (mask_dev NULL = all): n_used_dev[1], ones_dev[n] = #states with
 * s_i = +1, agree_dev[n*n] = #states with s_i == s_j (spin i <-> index bit n-1-i).
 * <s_i> = (2 ones_i - N)/N, <s_i s_j> = (2 agree_ij - N)/N. */
int qmc_state_moments(qmc_model *m, int64_t count, const uint64_t *states_dev, const uint8_t *mask_dev,
                      int64_t *n_used_dev, int64_t *ones_dev, int64_t *agree_dev, void *stream);

/* ---- Sharded statevector: n > 34, or any n with 18 <= n - log2(world) <= 34 ------------------------------
 * Replaces run_qmcmc_quantum_ckt (qumcmc/quantum_mcmc_routines.py:110-156) when 2^n amplitudes do not fit one GPU:
 * rank R of 2^g holds the amplitudes whose top g index bits equal R (2^(n-g) complex64, caller-owned buffer).
 * The proposal U = M (P M)^(r-1) is a sequence of local passes (qmc_shard_op, one kernel each) and all-to-all
 * exchanges on equal chunks, which swap the top g local index bits with the rank bits; the CALLER performs the
 * exchange (torch.distributed / NCCL all_to_all_single between two state buffers) and passes the layout parity
 * (0 = identity, 1 = top g local positions swapped with the rank bits) to every op.  The library tracks which logical
 * qubit sits at which position for each layout and never swaps back.  A proposal with r layers has r - 1 exchanges and
 * STARTS in layout (r - 1) & 1 (the first pass builds the analytic product state directly in that layout), so that it
 * ends in the identity layout and the measurement CDF runs in index order.  Op sequence for r >= 2
 * (K = qmc_shard_num_groups):
 *   FIRST(top) ; then per layer j = 2..r:  [LO_SINGLE unless j == r] SINGLE(k = 1..K-1) <exchange>
 *                                          MID(top, pre = g) if j < r   else   SINGLE(top, pre = g) LO_FINAL
 * and LO_FINAL alone for r == 1.  qumcmc_b200/sharded.py (sharded_plan, initial_layout) generates exactly this list.
 * After LO_FINAL: qmc_shard_norm -> local sum |a|^2; ranks all-gather the W sums, the owner of u * total calls
 * qmc_shard_select and broadcasts the measured index. */
typedef struct qmc_shard qmc_shard;
#define QMC_SHARD_OP_FIRST 0
#define QMC_SHARD_OP_MID 1
#define QMC_SHARD_OP_SINGLE 2
#define QMC_SHARD_OP_LO_SINGLE 3
#define QMC_SHARD_OP_LO_FINAL 4
int qmc_shard_create(qmc_shard **out, qmc_model *m, int rank, int log2_world);
int qmc_shard_destroy(qmc_shard *sh);
int qmc_shard_num_groups(const qmc_shard *sh);
/* Parameters of the next proposal: start state s (logical basis index), gamma, r mixer layers, mixer group. */
int qmc_shard_begin(qmc_shard *sh, uint64_t s, double gamma, int r, double delta_time, int group);
int qmc_shard_op(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, void *stream);
/* Fused pass + exchange: the pass stores every amplitude straight into buffer `dst_which` of the rank that owns it
 * after the qubit swap (peer-mapped pointers over NVLink, set once with qmc_shard_set_peers: 2^g device-accessible
 * addresses of buffer `which` of every rank).  Replaces "pass; all_to_all_single" by "pass_exchange; barrier". */
int qmc_shard_set_peers(qmc_shard *sh, int which, void *const *ptrs_host);
int qmc_enable_peer_access(int device, int peer);
/* CUDA IPC plumbing for the peer tables: export = handle of the allocation containing dev_ptr + offset inside it;
 * open = map it into `device`'s context in another process (cached per allocation); close_all unmaps. */
#define QMC_IPC_HANDLE_BYTES 64
int qmc_ipc_export(int device, void *dev_ptr, void *handle_out /* 64 bytes */, uint64_t *offset_out);
int qmc_ipc_open(int device, const void *handle, uint64_t offset, void **ptr_out);
int qmc_ipc_close_all(void);
int qmc_shard_op_exchange(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, int dst_which, void *stream);
/* One chunk (the chunk-th of 2^chunk_bits contiguous sub-blocks of the shard) of a SINGLE (group k >= 1) or LO_SINGLE
 * pass: local when dst_which < 0, fused with the exchange when dst_which = 0 / 1 (then chunk_bits >= log2(world): one
 * destination rank per chunk).  Lets the caller overlap the local passes of chunk i+1 with the NVLink stores of chunk i. */
int qmc_shard_op_chunk(qmc_shard *sh, int op, int k, int pre, int layout, void *state_dev, int dst_which,
                       int chunk_bits, int64_t chunk, void *stream);
/* Overlap mode of the fused exchange: with ctas > 0 the exchange pass of a register-only (HI6) group runs as a persistent
 * kernel on `ctas` CTAs instead of one CTA per tile, so that the HBM-bound local passes of the next chunk (issued on
 * another stream) run beside the NVLink-bound stores instead of queueing behind them.  Same arithmetic, same results.
 * qmc_shard_exchange_can_persist: 1 when this shard's pre-exchange pass has such a variant. */
int qmc_shard_set_exchange_ctas(qmc_shard *sh, int ctas);
/* Exchange of one chunk by the copy engines (no SM): after the local passes of a chunk (qmc_shard_op_chunk with
 * dst_which = -1) its 2^(nl - chunk_bits) contiguous amplitudes are copied into buffer dst_which of the owning rank
 * (cudaMemcpyAsync into the peer mapping over NVLink).  Needs qmc_shard_set_peers and chunk_bits >= log2(world). */
int qmc_shard_copy_chunk(qmc_shard *sh, const void *state_dev, int dst_which, int chunk_bits, int64_t chunk, void *stream);
int qmc_shard_exchange_can_persist(const qmc_shard *sh);
int qmc_shard_norm(qmc_shard *sh, double *total_dev, void *stream);
int qmc_shard_select(qmc_shard *sh, void *state_dev, double target, int layout, uint64_t *index_dev, void *stream);
/* Test hook: LO_FINAL that also writes |a|^2 of the local shard (2^(n-g) floats, local index order). */
int qmc_shard_final_probs(qmc_shard *sh, int layout, void *state_dev, float *probs_dev, void *stream);

/* Introspection for benchmarks/tests. */
int qmc_model_num_spins(const qmc_model *m);
/* Number of kernels this library has launched on behalf of handle m since creation. */
int64_t qmc_model_launch_count(const qmc_model *m);
/* Events around the dominant sweep kernel: accumulated device ms and launches since reset
 * (timing is recorded only while enabled; off by default). */
int qmc_profile_enable(qmc_model *m, int on);
int qmc_profile_read(qmc_model *m, double *sweep_ms, int64_t *sweep_launches,
                     double *algorithmic_bytes);
/* Bytes the timed sweep launches actually read + wrote: 16 B * 2^n * K * (r - 1) per proposal with K passes per layer
 * (the first sweep of a proposal is write-only, the last read-only), against algorithmic_bytes = 16 B * 2^n * r. */
int qmc_profile_read_moved(qmc_model *m, double *moved_bytes);

/* Static-analysis hooks (CPU, no device needed; tests/test_smem_layouts.py): the shared-memory slot / statevector offset
 * a thread touches in the register layouts of the sweep kernels, evaluated by the functions the kernels themselves use.
 * They back the barrier-free "a thread rewrites exactly the slots it read" steps and the conflict-free swizzles with an
 * exhaustive check (compute-sanitizer is not available on the GPU pool). */
int qmc_debug_smem_slot(int layout, int tid, int j);
/* Launch list of one hop under the L2-resident group schedule (csrc/group_sched.cuh; pure host logic, no GPU needed):
 * rhist[r], r = 0..r_max = chains with r mixer layers (rhist[0] = 0, sum = n_chains); out receives 4 ints per launch --
 * kind (0 LO first, 1 LO mid, 2 LO final, 3 LO final of an r = 1 proposal, 4 HI first, 5 HI mid), first position in the
 * r-sorted chain order, number of chains, side stream.  Returns the number of launches, or -1 on bad arguments. */
int qmc_debug_group_plan(int n_chains, int r_max, const int *rhist, int group, int streams, int *out, int cap);
/* How the dispatcher classifies a mixer given its term masks (qumcmc/mixers.py:21-151 lowered to X-string masks): 1 = one-body
 * (rotation kernels), 2 = terms of weight <= 2 (XQ mode at 14 <= n <= 20), 3 = anything else (XT mode / on-chip Hadamard-basis
 * mode / general path); -1 on bad arguments.  Pure host logic. */
int qmc_debug_mixer_class(int n, int n_terms, const uint64_t *masks);
int qmc_debug_layout128(int geometry, int layout, int tid, int j, long long *out4);
int qmc_debug_general_slot(int stage, int tid, int j, int *local_out);

#ifdef __cplusplus
}
#endif
#endif /* QUMCMC_B200_H */
