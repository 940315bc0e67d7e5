/*
 * uqoc.h -- C ABI of libuqoc.so: B200 (sm_100a) kernels for disorder-sampled
 * unitary propagation + fidelity loss, forward and backward.
 *
 * This is the drop-in boundary for the one hot path of
 * shiminki/universal_quantum_optimal_control (paths relative to that repo):
 *
 *   SCORE.py   = train/unitary_single_qubit_gate/universal_single_qubit_SCORE.py
 *   grape.py   = train/GRAPE/grape_train.py
 *   trainer.py = model/universal_model_trainer.py
 *   util.py    = visualize/util.py
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller (tensor.data_ptr());
 *     the library never allocates or frees device memory;
 *   - arrays are contiguous row-major; "real" is float (dtype = UQOC_F32) or
 *     double (dtype = UQOC_F64); complex arrays are interleaved (re, im) reals;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises except uqoc_fp32_peak_probe;
 *   - return value 0 = success, negative = UQOC_E_* (message via
 *     uqoc_last_error(), thread-local), positive = cudaError_t of a failed launch;
 *   - re-entrant; no global mutable state; cudaSetDevice is the caller's job;
 *   - sample layout follows trainer.py:80-82: sample s = b*M + j for target b and
 *     Monte-Carlo index j; `error` is (E, B*M) with row 0 = delta (ORE), row 1 = eps
 *     (PLE)  [SCORE.py:113-114].
 *   - Hamiltonian is the CODE's form (SCORE.py:117-124):
 *       H = 1/2 (1+eps) (cos(phi) X + sin(phi) Y + delta Z),  U_i = exp(-i H tau_i)
 *   - domain: the table / polynomial sin/cos paths reduce the rotation angle with a magic-number round;
 *     |tau (1+eps) sqrt(1+delta^2)| must stay below ~6000 rad per pulse (the reference's pulse ranges give < 3)
 */
#ifndef UQOC_H_
#define UQOC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UQOC_VERSION 100 /* 0.1.0 */

/* dtype */
#define UQOC_F32 0
#define UQOC_F64 1

/* flags (bit mask) */
#define UQOC_FLAG_FAST_SINCOS 1u /* FP32 only: MUFU sin/cos instead of the polynomial path */
#define UQOC_FLAG_NO_PACKED 2u   /* FP32 only: scalar FFMA kernel instead of the packed f32x2 (FFMA2) one */
#define UQOC_FLAG_RNG_FROM_DEVICE 8u /* `seed` is a DEVICE pointer to uint64 {seed, offset}: lets a captured CUDA
                                       graph draw fresh Philox samples on every replay (offset argument ignored) */
#define UQOC_FLAG_NO_TABLE 4u    /* packed kernel: polynomial sin/cos instead of the shared-memory table */
#define UQOC_FLAG_WPS4 16u       /* packed kernel: force the pulse train to be split over the block's 4 warps */
#define UQOC_FLAG_WPS1 32u       /* packed kernel: force one warp per sample group */
#define UQOC_FLAG_RAW_TARGET 128u /* SU(2) shared-pulse entry points: `target_c` is the RAW complex target array
                                    (B, 2, 2) interleaved (re, im) instead of uqoc_su2_target_coeffs' output: the
                                    8 trace coefficients are formed in the kernel prologue (one launch less) */
#define UQOC_FLAG_SU4_PADE 64u   /* SU(4): per-pulse scaling-and-squaring exponential kernel instead of the default
                                    eigenframe kernel (one real-symmetric eigendecomposition per error sample) */
#define UQOC_FLAG_NO_FIN (1u << 30) /* fused SU(2) step: partials reduction / exchange / loss as separate launches instead
                                       of the in-kernel "last block" epilogue */
#define UQOC_FLAG_NO_FAT (1u << 31) /* packed kernel: no fat (7 x 128-thread, one per SM) blocks for few-target shapes */
/* tuning overrides (0 = let the library choose): samples per thread (1,2,4) and lanes per
 * sample (1,2,4,8,16,32) of the shared-pulse kernels */
#define UQOC_FLAG_ST(n) (((unsigned)(n) & 0xFu) << 8)
#define UQOC_FLAG_LPS(n) (((unsigned)(n) & 0x3Fu) << 12)
#define UQOC_FLAG_SPLITS(n) (((unsigned)(n) & 0xFFFu) << 18)

/* loss kinds for uqoc_loss_finalize (SCORE.py:185-198) */
#define UQOC_LOSS_SHARP 0      /* log(1+exp(-k(Fbar-tau)))*(1-Fbar)  SCORE.py:193-198 */
#define UQOC_LOSS_NLL 1        /* -log(Fbar)                          SCORE.py:185-186 */
#define UQOC_LOSS_INFIDELITY 2 /* 1-Fbar                              SCORE.py:189-190 */
#define UQOC_LOSS_NONE 3       /* Fbar itself                                          */

/* error codes */
#define UQOC_E_BADARG (-1)
#define UQOC_E_UNSUPPORTED (-2)
#define UQOC_E_WORKSPACE (-3)
#define UQOC_E_NODEVICE (-4)

int uqoc_version(void);
const char* uqoc_last_error(void);

/* ------------------------------------------------------------------------
 * Target preparation.  Tr(U_out^dagger T) is linear in the quaternion of U_out;
 * this turns B complex dxd targets (B, d, d, 2 reals) into the coefficient rows
 * the kernels consume: SU(2): (B, 8) reals = Re c_0..3, Im c_0..3 (SCORE.py:172-178).
 * ------------------------------------------------------------------------ */
int uqoc_su2_target_coeffs(const void* U_target, int64_t B, void* target_c, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * Workspace (bytes) needed by the fused / forward kernels for the deterministic
 * two-stage reduction over sample tiles.  Replaces nothing in the reference
 * (autograd keeps every (Bm,L,2,2) intermediate instead).
 * ------------------------------------------------------------------------ */
int64_t uqoc_su2_workspace_bytes(int64_t B, int64_t L, int64_t M, int dtype, unsigned flags);
/* The first 256 bytes of a workspace hold the ticket counter of the in-kernel epilogue: they must be ZERO when the
 * buffer is first handed to the library (which leaves them zero again at the end of every call), and one workspace
 * must not be shared by calls that can run concurrently (one per stream). */

/* ------------------------------------------------------------------------
 * Fused forward+backward for the trainer's Monte-Carlo step
 *   replaces trainer.py:80-88 + the autograd backward of SCORE.py:77-183
 *   (repeat_interleave, error H2D, generator, fidelity, sum over samples, dF/dpulses).
 *
 *   pulses      (B, L, 2) reals [phi, tau]  -- NOT repeated M times
 *   target_c    (B, 8) from uqoc_su2_target_coeffs
 *   err         (2, B*M) explicit errors, or NULL => on-chip Philox4x32-10 + Box-Muller
 *                with delta ~ N(0, sig_d^2), eps ~ N(0, sig_e^2)  (replaces SCORE.py:158-161
 *                + the H2D copy at trainer.py:82); counter = (j0+j, b, offset), key = seed
 *   weight      NULL, or (B*M) per-sample cotangent dLoss/dF_s (general autograd backward)
 *   M           samples per target processed by THIS call (this rank's shard)
 *   j0          global index of this call's first sample (Philox counter base)
 *   F_out       NULL or (B*M): per-sample fidelity  (SCORE.py:168-183)
 *   err_out     NULL or (2, B*M): the errors actually used
 *   Fsum        (B): sum_j F[b,j]
 *   G           (B, L, 2): sum_j weight * dF[b,j]/d pulses[b]   (unscaled by the loss)
 *   workspace   >= uqoc_su2_workspace_bytes(...) bytes
 * ------------------------------------------------------------------------ */
int uqoc_su2_fwdbwd(const void* pulses, const void* target_c, const void* err, const void* weight,
                    int64_t B, int64_t L, int64_t M, int64_t j0,
                    double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                    void* F_out, void* err_out, void* Fsum, void* G,
                    void* workspace, int64_t workspace_bytes,
                    int dtype, unsigned flags, void* stream);

/* Same for a SLICE of the targets: rows [b0, b0 + B) of a larger batch (pulses / target_c / outputs point at the slice).
 * b0 only enters the Philox counter, so that a batch processed in target chunks -- chunk n's all-reduce and device-to-host
 * copy overlapping chunk n+1's kernel -- draws exactly the samples of the un-chunked call. */
int uqoc_su2_fwdbwd_slice(const void* pulses, const void* target_c, const void* err, const void* weight,
                          int64_t B, int64_t L, int64_t M, int64_t j0, int64_t b0,
                          double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                          void* F_out, void* err_out, void* Fsum, void* G,
                          void* workspace, int64_t workspace_bytes,
                          int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Single-GPU training step in one call: uqoc_su2_fwdbwd followed by the loss epilogue of
 * uqoc_loss_finalize (n_total = B*M).  While the partial rows fit one block's bandwidth (few targets, e.g.
 * BASELINE config 3) the partials reduction, the loss and the chain factor run in the LAST block of the fused
 * kernel (one launch for the whole of trainer.py:80-90 between the model forward and the model backward);
 * otherwise in one or two follow-up launches.  loss_out = {loss, Fbar, dloss/dFbar};
 * G comes back already scaled (= d loss / d pulses).
 * ------------------------------------------------------------------------ */
int uqoc_su2_fwdbwd_loss(const void* pulses, const void* target_c, const void* err,
                         int64_t B, int64_t L, int64_t M,
                         double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                         int loss_kind, double tau, double k,
                         void* F_out, void* err_out, void* Fsum, void* G, void* loss_out,
                         void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Multi-GPU step with the exchange fused into the partials reduction, over NVLink / NVSwitch PEER MEMORY
 * instead of ncclAllReduce: uqoc_su2_fwdbwd for this rank's sample shard (j0, M), then ONE kernel that reduces
 * the sample-tile partials, pushes the (B + B*L*2) results into every rank's exchange buffer, raises per-block
 * flags at system scope, waits for the same block of every rank and sums the slots in rank order.  On return
 * (stream order) Fsum / G hold the sums over ALL ranks, bit-identical on every rank; follow with
 * uqoc_loss_finalize(n_total = B * M_global).  There is no reference counterpart (the reference is single
 * process, single device): it replaces the all-reduce this build adds for sample sharding when the
 * exchange vector is small (latency-bound, e.g. BASELINE config 3), NCCL stays the choice for MB-sized vectors.
 *   peer_data[q]  (HOST array, world entries): device address, valid in THIS process, of rank q's exchange
 *                 buffer of >= uqoc_peer_data_bytes(B + B*L*2, world, dtype) bytes (CUDA VMM / IPC peer mapping,
 *                 e.g. torch.distributed._symmetric_memory rendezvous -> buffer_ptrs)
 *                 ZEROED once before first use (FP32: every exchanged real travels as an aligned 8-byte
 *                 {value, epoch} word that the receiver polls - one NVLink latency, no fence / flag round trip)
 *   peer_flags[q] likewise, rank q's flag buffer of uqoc_peer_flag_bytes(world) bytes, zeroed once before first use
 *                 (FP64 and the single-launch small-output path signal arrival through these flags)
 *   epoch         non-zero, the same on every rank, different from the previous call's (e.g. a call counter)
 * All ranks must make the same sequence of calls with the same (B, L, dtype); at most 16 ranks.  A rank whose
 * peers never make the matching call gives up after 10 s and returns NaN in Fsum / G (no hung GPU).
 * ------------------------------------------------------------------------ */
int64_t uqoc_peer_data_bytes(int64_t n, int world, int dtype);
int64_t uqoc_peer_flag_bytes(int world);
int uqoc_su2_fwdbwd_peer(const void* pulses, const void* target_c, const void* err, const void* weight,
                         int64_t B, int64_t L, int64_t M, int64_t j0,
                         double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                         void* F_out, void* err_out, void* Fsum, void* G,
                         void* workspace, int64_t workspace_bytes,
                         int rank, int world, const uint64_t* peer_data, const uint64_t* peer_flags, uint32_t epoch,
                         int dtype, unsigned flags, void* stream);

/* Same step with the loss epilogue of uqoc_loss_finalize folded in (n_total = B * M_global): for small exchange
 * vectors the whole multi-GPU step is ONE launch -- the last block of the fused kernel reduces the partials, pushes
 * them to the peers, waits for theirs, sums in rank order, evaluates the loss and scales G.  loss_out as in
 * uqoc_su2_fwdbwd_loss; every rank ends with bit-identical {loss, Fsum, G}. */
int uqoc_su2_fwdbwd_peer_loss(const void* pulses, const void* target_c, const void* err,
                              int64_t B, int64_t L, int64_t M, int64_t j0, int64_t M_global,
                              double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                              int loss_kind, double tau, double k,
                              void* F_out, void* err_out, void* Fsum, void* G, void* loss_out,
                              void* workspace, int64_t workspace_bytes,
                              int rank, int world, const uint64_t* peer_data, const uint64_t* peer_flags, uint32_t epoch,
                              int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Forward only with pulses shared per target (trainer.py:113-120 evaluate;
 * util.py:214-223 / 244-249 sweeps where one sequence is expand()ed).
 *   U_out  NULL or (B*M, 2, 2, 2): composite unitary U_L...U_1 (SCORE.py:145)
 *   F_out  NULL or (B*M);  Fsum NULL or (B);  err_out NULL or (2,B*M)
 * ------------------------------------------------------------------------ */
int uqoc_su2_forward(const void* pulses, const void* target_c, const void* err,
                     int64_t B, int64_t L, int64_t M, int64_t j0,
                     double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                     void* U_out, void* F_out, void* err_out, void* Fsum,
                     void* workspace, int64_t workspace_bytes,
                     int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Forward-only sweep drivers.
 *  - grid: replaces visualize/util.py:231-249 (fidelity_contour_plot): the (2, Nd*Ne) error tensor
 *    of meshgrid(ORE, PLE, indexing="ij").flatten() is never built; sample j uses
 *    delta = axis_delta[j / n_eps], eps = axis_eps[j % n_eps].  The two axes must be ONE buffer
 *    [axis_delta (n_delta) | axis_eps (n_eps)].  Outputs are (B * n_delta * n_eps).
 *  - sigmas: replaces the 199-iteration Python loop of visualize/util.py:313-326
 *    (plot_fidelity_by_std) / :288-298: target row b draws its M Philox samples with
 *    (sigma_delta, sigma_eps) = sigma_table[b] (B, 2); pass the same pulse row B times.
 * ------------------------------------------------------------------------ */
int uqoc_su2_forward_grid(const void* pulses, const void* target_c, const void* axis_delta, int64_t n_delta,
                          const void* axis_eps, int64_t n_eps, int64_t B, int64_t L,
                          void* U_out, void* F_out, void* Fsum,
                          void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream);
int uqoc_su2_forward_sigmas(const void* pulses, const void* target_c, const void* sigma_table,
                            int64_t B, int64_t L, int64_t M, int64_t j0, uint64_t seed, uint64_t offset,
                            void* F_out, void* Fsum,
                            void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Strict reference signature: one pulse row PER SAMPLE.
 *   replaces batched_unitary_generator (SCORE.py:77-145, grape.py:78-138)
 *   pulses (Bm, L, 2), err (2, Bm)  ->  U_out (Bm, 2, 2) complex
 * and its autograd backward (LinalgMatrixExpBackward0 + BmmBackward0 chain):
 *   grad_U (Bm, 2, 2) complex cotangent (torch convention dL = Re sum conj(g) dU)
 *   -> grad_pulses (Bm, L, 2)
 * ------------------------------------------------------------------------ */
int uqoc_su2_generator_forward(const void* pulses, const void* err, int64_t Bm, int64_t L,
                               void* U_out, int dtype, unsigned flags, void* stream);
int uqoc_su2_generator_backward(const void* pulses, const void* err, const void* grad_U,
                                int64_t Bm, int64_t L, void* grad_pulses,
                                int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Two-qubit SU(4) twins.  NOT IN THE REFERENCE (README.md:86,122 promise train/two_qubit/, absent;
 * fidelity() is already generic in d, SCORE.py:181-183).  Builder-defined contract:
 *   pulses (B, L, 3) = [phi1, phi2, tau];  err (3, B*M) = [delta1; delta2; eps] or NULL => Philox
 *   (delta1, eps from the first Box-Muller pair -- the SU(2) stream -- delta2 from the second);
 *   H = 1/2 [cos phi1 XI + sin phi1 YI + cos phi2 IX + sin phi2 IY + delta1 ZI + delta2 IZ + J ZZ],
 *   U_k = exp(-i H_k tau_k (1+eps));  target (B, 4, 4) complex;  F = (|Tr(U^dagger T)|^2 + 4)/20
 *   G (B, L, 3) = sum_j weight * dF/d[phi1, phi2, tau];  U_out (B*M, 4, 4) complex.
 * UQOC_FLAG_RNG_FROM_DEVICE works as for the SU(2) entry points (`seed` = device pointer to {seed, offset}).
 * uqoc_su4_workspace_bytes fits forward and forward+backward launches of the shape alike.
 * uqoc_su4_generator_backward is the autograd backward of the strict per-sample signature
 * (pulses (Bm, L, 3), err (3, Bm), uqoc_su4_forward with B = Bm, M = 1 is its forward): grad_U (Bm, 4, 4) complex
 * cotangent in torch's convention (dL = Re sum conj(g) dU) -> grad_pulses (Bm, L, 3).
 * ------------------------------------------------------------------------ */
int64_t uqoc_su4_workspace_bytes(int64_t B, int64_t L, int64_t M, int dtype, unsigned flags);
int uqoc_su4_fwdbwd(const void* pulses, const void* target, const void* err, const void* weight,
                    int64_t B, int64_t L, int64_t M, int64_t j0, double J,
                    double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                    void* F_out, void* err_out, void* Fsum, void* G,
                    void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream);
int uqoc_su4_forward(const void* pulses, const void* target, const void* err,
                     int64_t B, int64_t L, int64_t M, int64_t j0, double J,
                     double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                     void* U_out, void* F_out, void* err_out, void* Fsum,
                     void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream);
int uqoc_su4_generator_backward(const void* pulses, const void* err, const void* grad_U,
                                int64_t Bm, int64_t L, double J, void* grad_pulses,
                                int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Pulse heads: the element-wise tail of the two pulse generators, fused to
 * one launch each way.  ranges = {lo_phi, hi_phi, lo_tau, hi_tau} (HOST pointer, 4 doubles).
 *  mode 0, logits (B, L, 2): model/universal_model.py:131-143
 *      p = lo + (hi-lo) sigmoid(x); [p = scale*p + base_pulse (L,2)] (finetune, :135-138);
 *      tau = relu(tau); phi = ((phi + phi_offset[b] + pi) mod 2pi) - pi          (:140-143)
 *  mode 1, logits (B, L, 3): model/GRAPE_model.py:76-89
 *      (ux,uy,ut) = sigmoid(x); phi = lo + (hi-lo) atan2(uy,ux); tau = relu(lo + (hi-lo) ut)
 * ------------------------------------------------------------------------ */
int uqoc_pulse_head_forward(const void* logits, const void* phi_offset, const void* base_pulse, int64_t B, int64_t L,
                            int mode, const double* ranges, double scale, void* pulses, int dtype, void* stream);
int uqoc_pulse_head_backward(const void* logits, const void* base_pulse, const void* grad_pulses, int64_t B, int64_t L,
                             int mode, const double* ranges, double scale, void* grad_logits, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * The same heads FOLDED INTO the fused step (SURVEY.md §8f row f-3 as specified): the head runs while the fused
 * kernel stages the pulse train and its backward while the gradient rows are written, so the whole
 *   logits -> pulses -> Monte-Carlo propagation -> fidelity -> loss -> d loss / d logits
 * of model/universal_model.py:131-143 (or model/GRAPE_model.py:76-89) + trainer.py:80-90 is the fused kernel and its
 * epilogue; no (B, L, 2) pulses tensor is exchanged between model and op.
 *   logits (B, L, 2) head_mode 0 | (B, L, 3) head_mode 1;  ranges / scale / phi_offset / base_pulse as above
 *   G (B, L, 2 | 3) = d loss / d logits (loss_kind >= 0) or d(sum_j F)/d logits (loss_kind = -1); G = NULL: forward only
 *   pulses_out (B, L, 2) nullable: the pulses the head produced (logging, trainer.py:260-266)
 *   loss_out (3) as uqoc_loss_finalize;  errors explicit (2, B*M) or Philox (err = NULL);  workspace as uqoc_su2_fwdbwd
 * ------------------------------------------------------------------------ */
int uqoc_su2_head_step(const void* logits, int head_mode, const double* ranges, double scale, const void* phi_offset,
                       const void* base_pulse, const void* target_c, const void* err,
                       int64_t B, int64_t L, int64_t M, double sig_d, double sig_e, uint64_t seed, uint64_t offset,
                       int loss_kind, double tau, double k,
                       void* pulses_out, void* F_out, void* err_out, void* Fsum, void* G, void* loss_out,
                       void* workspace, int64_t workspace_bytes, int dtype, unsigned flags, void* stream);

/* ------------------------------------------------------------------------
 * Loss epilogue (SCORE.py:185-198 applied to the pooled mean, and the chain rule
 * of loss.backward() at trainer.py:90):
 *   Fbar = sum_b Fsum[b] / n_total;  loss_out[0] = loss(Fbar); loss_out[1] = Fbar;
 *   loss_out[2] = dloss/dFbar;  G[...] *= dloss/dFbar / n_total   (in place, if G != NULL)
 * n_total = B * M_global (all ranks), so after an all-reduce of [Fsum | G] every
 * rank gets identical loss and gradients.
 * ------------------------------------------------------------------------ */
int uqoc_loss_finalize(const void* Fsum, int64_t B, double n_total, int loss_kind,
                       double tau, double k, void* G, int64_t G_numel, void* loss_out,
                       int dtype, void* stream);

/* ------------------------------------------------------------------------
 * Stand-alone fidelity on materialised unitaries, for callers that keep the reference's
 * three-call structure (generator -> fidelity_fn -> loss_fn, trainer.py:84-88):
 *   F[s] = (|Tr(U_out[s]^dagger U_target[s])|^2 + d) / (d (d+1))         SCORE.py:168-183
 *   U_out (Bm, d, d) complex; U_target complex with `target_stride` reals between
 *   consecutive samples (0 = one target broadcast to all samples, 2*d*d = per-sample)
 * backward: grad_U = grad_F * 2/(d(d+1)) * conj(tr) * U_target  (torch complex convention).
 * uqoc_sum: deterministic sum of n reals into out[0] (the mean of SCORE.py:186/190/194 is
 * sum / n); workspace >= 1024 reals when n > 4096.
 * ------------------------------------------------------------------------ */
int uqoc_fidelity_forward(const void* U_out, const void* U_target, int64_t Bm, int d, int64_t target_stride,
                          void* F, int dtype, void* stream);
int uqoc_fidelity_backward(const void* U_out, const void* U_target, const void* grad_F, int64_t Bm, int d,
                           int64_t target_stride, void* grad_U, int dtype, void* stream);
int uqoc_sum(const void* x, int64_t n, void* out, void* workspace, int64_t workspace_bytes, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * Device-side error sampler: replaces get_ore_ple_error_distribution
 * (SCORE.py:158-161) -- err_out (2, B*M), same Philox stream as the fused kernels.
 * ------------------------------------------------------------------------ */
int uqoc_philox_errors(int64_t B, int64_t M, int64_t j0, double sig_d, double sig_e,
                       uint64_t seed, uint64_t offset, void* err_out, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * Roofline denominator: dependency-free FFMA loop on every SM, timed with CUDA
 * events inside the call (synchronises).  Returns achieved TFLOP/s in *tflops:
 * dtype UQOC_F32 = scalar FFMA, UQOC_F64 = DFMA, 2 = packed FFMA2 (f32x2).
 * ------------------------------------------------------------------------ */
int uqoc_fp32_peak_probe(int iters, int dtype, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* UQOC_H_ */
