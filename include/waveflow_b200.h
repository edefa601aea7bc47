/*
 * waveflow_b200 -- C ABI of the B200 (sm_100a) hot path of aspuru-guzik-group/waveflow.
 *
 * The reference is pure Python/JAX and has no FFI of its own; its operator boundary is the closure protocol
 * (SURVEY.md section 8b).  Each entry point below replaces one reference function (cited as file:line relative to
 * /root/reference/waveflow) and is what an XLA-FFI / ctypes binding for that function would call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns and allocates every buffer;
 *   - all arrays are dense row-major float32 unless stated; sizes are int64_t; nothing is allocated or freed here;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant and CUDA-graph capturable;
 *   - return value: 0 = OK, <0 = WF_ERR_* (invalid argument / unsupported configuration), >0 = cudaError_t.
 *   - "table" arguments are the reference's cached basis tables [4][P][T] (isplines_jax.py:112-131) re-laid-out by
 *     wf_table_layout_host() -- see that function.
 */
#ifndef WAVEFLOW_B200_H
#define WAVEFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WF_OK 0
#define WF_ERR_INVALID_ARG (-1)
#define WF_ERR_UNSUPPORTED (-2)
#define WF_ERR_NO_DEVICE (-3)

#define WF_MAX_P 32        /* max number of spline bases per element handled by the fused kernels            */
#define WF_HIDDEN 64       /* conditioner width, fixed by the reference (model_factory.py:38,72)              */
#define WF_WIN 8           /* local-support window of the compact node records                                */
#define WF_MAX_D 8         /* capacity of the per-model arrays (protons, peer ranks = 2 * WF_MAX_D)                    */
#define WF_MAX_FUSED_D 4   /* max flow dimension of the fused live kernels (wf_live_*, wf_local_energy, wf_vqmc_*)      */
#define WF_MAX_LAYERS 16

/* spline kinds */
#define WF_KIND_I 0
#define WF_KIND_M 1
#define WF_KIND_B 2

/* ABI / build info: returns WF_ABI_VERSION; writes the compiled SM arch (e.g. 100) to *sm_arch if non-null. */
#define WF_ABI_VERSION 2
int wf_abi_version(int* sm_arch);
/* Human-readable text for a status code returned by any wf_* call (static storage). */
const char* wf_status_string(int status);

/* Measurement aid (bench.py): runs an FFMA-only loop, 2 * 8 * iters * blocks * 256 FLOPs, to time the FP32 ceiling. */
int wf_probe_fma(int iters, int blocks, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Table layouts (host helper, pure C, no CUDA): converts the reference layout tab[4][P][T] (float32) to
 *   dense_t  [T][4][PP]      PP = P rounded up to a multiple of 4, zero padded   (transposed: one node = one row)
 *   rec      [T][4][WF_WIN]  compact local-support records; rec[m][nd][j] = tab[nd][lo[m]+j][m]
 *   lo       [T] int32       first basis whose value at node m is not the "prefix" value
 * prefix value of basis q < lo[m] is (nd==0 && kind==WF_KIND_I) ? 1 : 0; bases q >= lo[m]+WF_WIN are 0.
 * Returns WF_ERR_UNSUPPORTED (and leaves rec/lo untouched) if some node has more than WF_WIN-1 non-prefix entries or the
 * windows of adjacent nodes shift by more than one basis -- callers then use the dense kernels only.
 * rec/lo may be NULL to request only dense_t.
 * ------------------------------------------------------------------------------------------------------------------ */
int wf_table_layout_host(const float* tab_host, int kind, int P, int T, float* dense_t_host, float* rec_host,
                         int32_t* lo_host);

/* ------------------------------------------------------------------------------------------------------------------
 * Spline operators at the reference's operator boundary.
 * ------------------------------------------------------------------------------------------------------------------ */

/* apply_fun_vec / apply_fun_vec_grad of ISpline_fun, MSpline_fun (isplines_jax.py:139-149, msplines_jax.py:116-126) and,
 * with pre-mixed coefficients, of BSpline_fun:   out_k[m] = sum_q c[m][q] * basis_q^{(nd0+k)}(x[m]),  k = 0..n_out-1,
 * each basis value being the reference's table interpolation (isplines_jax.py:45-56; derivative = next table, :60-66;
 * table index nd > 3 clamps to 3).  Generic path: any x (JAX gather clamp/wrap semantics), any P <= 64.
 *   out[k] may be NULL to skip an order.  out_logd (nullable) = log(out[1] + 1e-7) (made.py:79), requires n_out >= 2. */
int wf_spline_apply_dense(const float* dense_t, int T, int P, const float* c, const float* x, int64_t M, int nd0,
                          int n_out, float* const* out_host /* n_out device pointers in a host array */,
                          float* out_logd, void* stream);

/* Fused value + derivative + log-derivative of a LOCALLY SUPPORTED table spline (I or M tables, nd 0 and 1) -- the
 * HBM-bound "spline + log-det" operator: persistent CTAs, coefficient tiles staged with cp.async.bulk (TMA), node
 * records resident in shared memory.  Same results as wf_spline_apply_dense(nd0=0,n_out=2) up to summation order of
 * exact zeros.  out_val / out_grad / out_logd are each nullable. */
int wf_spline_apply_local(const float* rec, const int32_t* lo, const float* dense_t, int kind, int T, int P,
                          const float* c, const float* x, int64_t M, float* out_val, float* out_grad, float* out_logd,
                          void* stream);

/* BSpline_fun.apply_fun_vec (bsplines_jax.py:127-141): c = w @ ob_to_b; c /= ||c||_2; out = sum_j c_j OB_j^{(nd)}(x).
 * ob_dense_t is the [T][4][PP] layout of the orthonormalised tables, ob_to_b is [P][P] row-major. */
int wf_bspline_apply(const float* ob_dense_t, const float* ob_to_b, int T, int P, const float* w, const float* x,
                     int64_t M, int nd, float* out, void* stream);

/* remove_bias (isplines_jax.py:196-202 for WF_KIND_I, msplines_jax.py:186-192 for WF_KIND_M). out may alias p. */
int wf_remove_bias(int kind, int k, int P, const float* p, int64_t M, float* out, void* stream);

/* enforce_boundary_conditions (isplines_jax.py:158-194, msplines_jax.py:156-184, bsplines_jax.py:173-199).
 * The constraint dictionaries are passed as parallel host arrays in dict order; bv_left_host[i*4 + j] (j <= nd_i) are
 * the boundary basis values basis_j^{(nd_i)}(0) and bv_right_host[i*4 + j] = basis_{P-1-j}^{(nd_i)}(1), i.e. exactly
 * the I_cached(0.0, ...) / I_cached(1.0, ...) constants the reference closes over.  nd_i <= 3.
 * Final normalisation: /sum (I, M) or /||.||_2 (B).  For WF_KIND_I a right constraint {0: 1.0} sets w[P-1] = 0. */
int wf_enforce_bc(int kind, int P, int n_left, const int* nd_left_host, const float* val_left_host,
                  const float* bv_left_host, int n_right, const int* nd_right_host, const float* val_right_host,
                  const float* bv_right_host, const float* w, int64_t M, float* out, void* stream);

/* reverse_fun_vec (isplines_jax.py:153-156) = utils/helpers.py:150-166 bisection on [0,1] of sum_q c_q I_q(x) - y with
 * tolerance tol; returns the LOWER bracket.  n_iter (nullable, int32 [M]) receives the iteration count per element. */
int wf_spline_reverse(const float* dense_t, int T, int P, const float* c, const float* y, int64_t M, float tol,
                      float* out, int32_t* n_iter, void* stream);

/* sample_fun_vec of MSpline_fun (msplines_jax.py:129-154, kind WF_KIND_M) and BSpline_fun (bsplines_jax.py:144-171, kind
 * WF_KIND_B): out[m][s], s < num_samples, are rejection samples of the density proportional to sum_q params[m][q] M_q(x)
 * (M; bound ymax = max_q params[m][q] * n_knots) or (sum_j c_j OB_j(x))^2 with c = params[m] @ ob_to_b / ||.|| (B; bound
 * max((c @ b_to_ob)^2)).  dense_t: the [T][4][PP] layout of the M tables / of the orthonormalised B tables.
 * Philox4x32-10 streams keyed by (seed, row_keys[m] (nullable), m * num_samples + s, attempt): the draws agree with the
 * reference's threefry streams in distribution, not bit by bit. */
int wf_spline_sample(const float* dense_t, int kind, int T, int P, const float* ob_to_b, const float* b_to_ob, int n_knots,
                     const float* params, int64_t M, int num_samples, uint64_t seed, const int64_t* row_keys, float* out,
                     void* stream);

/* unconstrained_RQS (flows/bijections/neural_splines.py:16-71,74-184): rational-quadratic spline with K bins on
 * [-tail_bound, tail_bound], identity tails.  inputs [M], uw/uh [M][K], ud [M][K-1] (unnormalised).
 * flags: WF_RQS_INVERSE runs the inverse branch and returns -logabsdet (the reference's `inverse=True`);
 *        WF_RQS_EXACT_BINS evaluates the knot positions with exactly the float32 operation sequence of the reference's
 *        arithmetic (neural_splines.py:98-107: softmax with a correctly rounded exp, sequential cumsum, no fused
 *        multiply-adds), so bin_idx is bit-identical to it for every input; the default path computes the softmax with
 *        ex2.approx and one reciprocal (knots within ~1e-7 relative: an input closer than that to a knot may be assigned
 *        the neighbouring bin -- both are continuous there, outputs agree to float32 accuracy).
 * bin_idx (nullable, int32 [M]) receives the located bin (-1 in the tails). */
#define WF_RQS_INVERSE 1
#define WF_RQS_EXACT_BINS 2
int wf_rqs_apply(const float* inputs, const float* uw, const float* uh, const float* ud, int64_t M, int K,
                 float tail_bound, int flags, float* outputs, float* logabsdet, int32_t* bin_idx, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused flows.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Static description of a model built by model_factory.get_model / get_waveflow_model (model_factory.py:96-146). */
typedef struct wf_live_model {
  int32_t D;            /* flow dimension (2..WF_MAX_FUSED_D)                                                     */
  int32_t n_layers;     /* number of (IMADE, Reverse) pairs                                                        */
  int32_t T;            /* mesh points of the tables (2000)                                                        */
  int32_t P_I, k_I;     /* I-spline bases / degree                                                                 */
  int32_t prior_kind;   /* WF_KIND_B (Waveflow), WF_KIND_M (MFlow), -1: uniform prior on [0,1] (Flow/IFlow)        */
  int32_t P_P, k_P;     /* prior bases / degree                                                                    */
  int32_t has_box;      /* 1: BoxTransformLayer first (made.py:108-204)                                            */
  int32_t coord_mean;   /* 1: xu_coord_type == 'mean', 0: 'first'                                                  */
  int32_t bc_I;         /* bit0: left {0:0}, bit1: right {0:1}   (other constraint sets: operator-level path only) */
  int32_t bc_P;         /* bit0: left {0:0}, bit1: right {0:0}; bit2: B prior third layer pre-multiplied, see below      */
  float box;            /* box_side L                                                                              */
  float reg;            /* spline_regularization (made.py:68)                                                      */
  float tol;            /* reverse_fun_tol (made.py:44)                                                            */
  int32_t n_knots_P;    /* length of the prior's knot vector (M prior sampling bound, msplines_jax.py:145-148)     */
  int32_t weight_layout; /* WF_WEIGHTS_SIMT: packed weights as described below (CUDA-core kernels; required by the inverse /
                          * sampler); WF_WEIGHTS_TC: the tensor-core image made by wf_live_pack_tc (wf_live_forward,
                          * wf_local_energy, wf_local_energy_exchange: conditioner layers 2 and 3 on tcgen05, 3xTF32)      */
} wf_live_model;
#define WF_WEIGHTS_SIMT 0
#define WF_WEIGHTS_TC 1

/* Device-resident basis tables of a model (a HOST struct of DEVICE pointers), all produced by wf_table_layout_host:
 *   dense_* [T][4][32]  transposed tables zero-padded to WF_MAX_P bases; rec_* [T][4][8] / lo_* [T] the compact
 *   local-support records.  *_I: I-spline tables of the flow layers.  *_P: prior tables -- for the B prior dense_P holds
 *   the ORTHONORMALISED tables (bsplines_jax.py:98-106) and rec_P/lo_P are unused (NULL); for the M prior all three are
 *   the M-spline tables.  ob_to_b [P_P][P_P] row-major (B prior only). */
typedef struct wf_live_tables {
  const float* dense_I;
  const float* rec_I;
  const int32_t* lo_I;
  const float* dense_P;
  const float* rec_P;
  const int32_t* lo_P;
  const float* ob_to_b;
  const float* b_to_ob;   /* [P_P][P_P], B prior, sampler only (bsplines_jax.py:163-165) */
  const float* rec_I_t;   /* rec_I transposed to [T][8][4] (window slot major, the 4 derivative orders contiguous); needed by */
  const float* rec_P_t;   /* the WF_WEIGHTS_TC kernels only (rec_P_t: M prior), 16-byte aligned, else NULL                   */
} wf_live_tables;

/* Packed weights (device, float32), one block per conditioner, IMADE nets first then the prior net; per net
 *   W1m [D][64] | b1 [64] | W2m [64][64] | b2 [64] | W3p [64][D][32] | b3p [D][32]
 * W*m have the MADE masks already applied (model_factory.py:8-19,31-33); W3p/b3p are the third layer re-ordered so
 * that the P coefficients of dimension d are contiguous (p[n,d,q] = o[n, q*D + d], model_factory.py:59-60) and padded
 * with zeros to 32.  Size per net: wf_live_net_floats(D).  The buffer must be 16-byte aligned.
 * B prior with bit 2 of bc_P set (wf_live_forward / wf_local_energy only; P_P <= 31): the prior net's W3p / b3p hold the third
 * layer already multiplied by diag(boundary mask) @ ob_to_b (bsplines_jax.py:132-134,173-198 are linear in the conditioner
 * output), i.e. the kernel receives the un-normalised B-spline coefficients directly, and column 31 of every dimension
 * holds the weights of sum_p o_p (whose sign survives the conditioner's own normalisation, model_factory.py:69-70). */
int64_t wf_live_net_floats(int D);

/* Tensor-core weight image (WF_WEIGHTS_TC) of `n_nets` conditioners packed as above (device -> device, one launch; run it
 * once per parameter set).  Per net, wf_live_net_floats_tc(D) floats:
 *   W2 hi | W2 lo  [2 k-blocks][64 rows = output unit][32 floats]   TF32-exact planes (w = hi + lo) of the masked second
 *   W3 hi | W3 lo  [2 k-blocks][32 D rows = (dim, coefficient)][32]  and third layer, K-major with the 128-byte swizzle of
 *                                                                    the tcgen05 shared-memory descriptors (16-byte chunk
 *                                                                    c of row n stored at c ^ (n % 8))
 *   W1 [D][64] | b1 [64] | b2 [64] | b3 [D][32]                      float32 (first layer and biases stay on CUDA cores)
 * Supported: D in 2..4; a B prior must be packed with the pre-multiplied third layer (bit 2 of bc_P). */
int64_t wf_live_net_floats_tc(int D);
int wf_live_pack_tc(int D, int n_nets, const float* weights, float* weights_tc, void* stream);

/* Outputs selector: any of the output pointers of wf_live_forward may be NULL.
 *   u [N][D]     flow output in the unit cube (Serial.direct_fun, bijections.py:452-460)
 *   logdet [N]   log|det J|
 *   logpdf [N]   MFlow.log_pdf / Waveflow.log_pdf (distributions.py:139-163, wavefunctions.py:33-52)
 *   psi [N]      Waveflow.psi (wavefunctions.py:54-71), B prior only
 * Fused forward pass: BoxTransformLayer -> (IMADE, Reverse) x L -> prior, one thread per sample, no intermediate leaves
 * the SM; conditioner weights stream through shared memory with cp.async.bulk (TMA). */
int wf_live_forward(const wf_live_model* model_host, const wf_live_tables* tables_host, const float* weights,
                    const float* x, int64_t N, float* u, float* logdet, float* logpdf, float* psi, void* stream);

/* Local energy with a fused forward-mode Laplacian (replaces jax.hessian in utils/physics.py:50-52,79-93 and the
 * E_loc of vqmc.py:198-200):  psi, H psi = -1/2 lap psi + V psi, E_loc = H psi / (psi + 1e-8), with V the soft-Coulomb
 * potential of physics.py:60-76 for `n_protons` protons at positions protons_host[n_protons] (1 space dimension).
 * Optional outputs (nullable): psi[N], hpsi[N], eloc[N], grad[N][D] (d psi / dx), lap[N].
 * sums (nullable, double[4], must be zeroed by the caller): += {sum E_loc, sum E_loc^2, count, sum psi^2}. */
int wf_local_energy(const wf_live_model* model_host, const wf_live_tables* tables_host, const float* weights,
                    const float* protons_host, int n_protons, const float* x, int64_t N, float* psi, float* hpsi,
                    float* eloc, float* grad, float* lap, double* sums, void* stream);

/* wf_local_energy followed by the estimator exchange of SURVEY 8e (see wf_p2p_allreduce_sums for the protocol and the
 * buffers): sums_out[0..3] = sum over ranks of this rank's `sums` after the launch.  With WF_WEIGHTS_TC the exchange runs in
 * the TAIL of the local-energy kernel (the last CTA to retire publishes the block sums to the peers and waits for theirs):
 * one launch per VQMC step.  done_counter: device uint32, zero-initialised, private to this stream. */
int wf_local_energy_exchange(const wf_live_model* model_host, const wf_live_tables* tables_host, const float* weights,
                             const float* protons_host, int n_protons, const float* x, int64_t N, float* psi, float* hpsi,
                             float* eloc, float* grad, float* lap, double* sums, const uint64_t* peer_bufs_dev, int rank, int world,
                             uint64_t step, double* sums_out, uint32_t* done_counter, void* stream);

/* Serial.inverse_fun for the live flow (bijections.py:462-463): (Reverse, IMADE.inverse) x L then the box inverse.
 * exact == 0 reproduces the reference (made.py:85-100: coefficients conditioned on the layer INPUT, quirk Q1; bisection
 * of helpers.py:150-166 returning the lower bracket; box inverse of made.py:186-197, correct for D = 2 only).
 * exact != 0 conditions each dimension on the already-inverted prefix and uses the true box inverse. */
int wf_live_inverse(const wf_live_model* model_host, const wf_live_tables* tables_host, const float* weights,
                    const float* u, int64_t N, int exact, float* x, void* stream);

/* Waveflow.sample / MFlow.sample (wavefunctions.py:74-107, distributions.py:165-190): D rounds of conditioner ->
 * per-sample rejection sampling of the prior (bsplines_jax.py:144-171, msplines_jax.py:129-154), then the inverse flow,
 * in one launch.  Counter-based Philox4x32-10 streams keyed by (seed, sample index, column, attempt); NOT
 * bit-compatible with JAX's threefry stream (statistical parity only).  u_out (nullable) receives the prior-space draws. */
int wf_live_sample(const wf_live_model* model_host, const wf_live_tables* tables_host, const float* weights,
                   uint64_t seed, int64_t N, int exact, float* x, float* u_out, void* stream);

/* Serial(NeuralSplineCoupling x L).direct_fun / inverse_fun (flows/bijections/neural_splines.py:244-296), fused over
 * all layers; one thread per sample, the 3K-1 spline parameters of a dimension never leave registers.
 * weights: per layer f1 then f2; per conditioner  head = W1[D/2][Hd] | b1 | W2[Hd][Hd] | b2 (padded to 4 floats), then
 * D/2 blocks  W3_j[Hd][3*KP] | b3_j[3*KP]  with the (W, H, D) columns of target dimension j each padded to KP = 8 (K <= 8)
 * or 32; size wf_rqs_coupling_net_floats(D, K, Hd).  Supported: D in {2, 4, 8}, K <= 32, Hd in {8, 64}.
 * inverse != 0 applies the layers in reverse with the inverse spline and accumulates -logabsdet (:274-292). */
int64_t wf_rqs_coupling_net_floats(int D, int K, int Hd);
int wf_rqs_coupling_flow(const float* weights, int n_layers, int D, int K, int Hd, float tail_bound, int inverse,
                         const float* x, int64_t N, float* y, float* logdet, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 + TMEM + TMA) dense layers with float32-grade accuracy ("3xTF32"), for the wide conditioner MLPs
 * of the coupling flow (neural_splines.py:187-188 at hidden_dim = 512, BASELINE config 5).
 * ------------------------------------------------------------------------------------------------------------------ */

/* x -> TF32-exact planes hi = tf32(x), lo = tf32(x - hi)  (n elements). */
int wf_tf32_split(const float* x, int64_t n, float* hi, float* lo, void* stream);

/* out = act(A W^T + bias):  A = a_hi + a_lo [M][K] (TF32-exact planes), W = w_hi + w_lo [N][K] (row-major, i.e. the
 * stax.Dense kernel transposed), accumulated as a_hi w_hi + a_hi w_lo + a_lo w_hi in fp32 in tensor memory.
 * mode 0: out_hi [M][N] = result (+ bias).   mode 1: tanh, then split into TF32-exact planes out_hi / out_lo (the next
 * layer's A operand).  K % 32 == 0, N % 128 == 0, all pointers 16-byte aligned; bias [N] nullable. */
int wf_tc_dense(const float* a_hi, const float* a_lo, int64_t M, int K, const float* w_hi, const float* w_lo, int N,
                const float* bias, int mode, float* out_hi, float* out_lo, void* stream);

/* Serial(NeuralSplineCoupling x L) for D = 64, K = 64 bins, hidden_dim = 512 (BASELINE config 5) on the tensor cores:
 * three tcgen05 GEMMs per half-update, the rational-quadratic spline evaluated in the epilogue of the third one straight
 * from tensor memory (the [N, 6112] parameter tensor never exists in HBM).
 * weights: per layer f1 then f2, each wf_rqs_coupling_tc_net_floats() floats:
 *   W1t_hi [512][32] | W1t_lo | b1 [512] | W2t_hi [512][512] | W2t_lo | b2 [512] | W3t_hi [32*192][512] | W3t_lo | b3p [32*192]
 * (transposed stax.Dense kernels as TF32-exact hi/lo planes; third layer regrouped per target dimension: 64 widths,
 * 64 heights, 63 derivatives, 1 pad).  workspace: wf_rqs_coupling_tc_workspace_floats(rows) floats for a chunk of `rows`
 * samples (the call loops over chunks); y [N][64] doubles as the state buffer.  All buffers 16-byte aligned. */
int64_t wf_rqs_coupling_tc_net_floats(void);
int64_t wf_rqs_coupling_tc_workspace_floats(int64_t rows);
int wf_rqs_coupling_flow_tc(const float* weights, int n_layers, float tail_bound, int inverse, const float* x, int64_t N,
                            float* y, float* logdet, float* workspace, int64_t workspace_floats, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * VQMC training step: value_and_grad(loss_fn_efficient) + Adam  (vqmc.py:193-221, jax.example_libraries.optimizers.adam)
 * ------------------------------------------------------------------------------------------------------------------ */

/* Parameter layout of the training path ("reference-flat", float32): conditioners in the order IMADE_0 .. IMADE_{L-1},
 * prior; per conditioner the stax.Dense kernels AS STORED by the reference (unmasked, model_factory.py:21-35):
 *   W1 [D][64] | b1 [64] | W2 [64][64] | b2 [64] | W3 [64][D*P] | b3 [D*P] | zero_params [D*P]
 * with P = P_I for the flow and P_P for the prior, i.e. the leaves of the reference's parameter pytree in traversal order
 * (zero_params, model_factory.py:83-84, is unused by this configuration: read by nothing, gradient 0).
 * Gradients use the same layout.  Returns the number of floats, or -1 for an unsupported model.
 * Supported: Waveflow models (B prior, BoxTransformLayer, constraints {0:0}|{0:1} / {0:0}|{0:0}), D in 2..4, D*P_I <= 128 and
 * D*P_P + D <= 128 (the prior's third layer is evaluated pre-multiplied by mask @ ob_to_b plus D sign-sum columns). */
int64_t wf_vqmc_param_floats(const wf_live_model* model_host);

/* Floats of scratch needed to process `walkers` walkers in one chunk (activation jets of every layer). */
int64_t wf_vqmc_grad_workspace_floats(const wf_live_model* model_host, int64_t walkers);

/* Loss and parameter gradient of vqmc.py:193-212 for the walkers x [N][D]:
 *   E_w = H psi_w / (psi_w + 1e-8),   grad += inv_n_total * sum_w [ a_w dpsi_w/dtheta + b_w d(H psi)_w/dtheta ],
 *   a_w = 2 (E_w - running_average) / psi_w - H psi_w / psi_w^2,  b_w = 1 / psi_w     (the custom_jvp of the reference),
 * evaluated by a reverse pass through the forward-mode Laplacian (table derivatives = next table; order 4 clamps to 3).
 * grad [wf_vqmc_param_floats] is ACCUMULATED into (zero it first; NULL = forward only; masked-out weights get 0);
 * inv_n_total = 1 / (number of walkers over all ranks).  psi, hpsi, eloc [N] nullable; sums (nullable, double[4]) +=
 * {sum E, sum E^2, count, sum psi^2}.  The call loops over chunks sized to the workspace; deterministic summation.
 * running_average_dev (nullable, device float[1]) overrides running_average: the call then has no host-side inputs that
 * change from step to step, so a captured CUDA graph of it can be replayed.
 * Execution: chunks of >= 128 x SM-count jet rows (N (D+2)) run the 64-wide conditioner layers, their input adjoints and the
 * weight gradients on tcgen05 (3xTF32, csrc/train_tc.cuh), smaller ones on CUDA-core GEMMs; a row's result does not depend on the
 * chunking.  The weight-gradient kernels are issued on an internal second stream of the device, forked from and joined back into
 * `stream` with events: everything is complete in `stream` order when the call returns, and a stream capture of `stream` records
 * the whole call.  One call at a time per device (the second stream and the workspace are not re-entrant). */
int wf_vqmc_loss_grad(const wf_live_model* model_host, const wf_live_tables* tables_host, const float* params,
                      const float* protons_host, int n_protons, const float* x, int64_t N, float running_average,
                      const float* running_average_dev, float inv_n_total, float* grad, float* psi, float* hpsi, float* eloc,
                      double* sums, float* workspace, int64_t workspace_floats, void* stream);

/* One Adam update (optimizers.adam: m, v moment buffers, bias correction with exponent step + 1), in place.
 * step_dev (nullable, device int64[1]) overrides step (graph replays). */
int wf_adam_step(float* params, float* m, float* v, const float* grad, int64_t n, int64_t step, const int64_t* step_dev, float lr,
                 float b1, float b2, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Estimator exchange over NVLink peer memory (the all-reduce of vqmc.py's energy sums)
 * ------------------------------------------------------------------------------------------------------------------ */

/* One-shot all-reduce (sum) of 4 doubles between `world` ranks of one node.  peer_bufs_dev: device array of `world`
 * pointers, entry r = rank r's symmetric buffer of wf_p2p_allreduce_buffer_bytes(world) bytes (zero-initialised, peer
 * mapped; e.g. torch.distributed._symmetric_memory).  step: sequence number, 1, 2, 3, ... identical on all ranks.  Every
 * rank stores {local[4], step} into its slot of every peer's buffer and sums the slots of its own buffer in rank order
 * (bit-identical results everywhere).  If a peer does not show up within ~2 s the rank returns NaN in out[0..3], sets the
 * sticky error word of its own buffer (the uint64 at byte offset 2 * world * 64: the step that timed out; the host may
 * poll it) and contributes NaN to every later exchange, so the failure reaches all ranks.  One 32-thread kernel. */
int64_t wf_p2p_allreduce_buffer_bytes(int world);
int wf_p2p_allreduce_sums(const uint64_t* peer_bufs_dev, int rank, int world, uint64_t step, const double* local, double* out,
                          void* stream);

/* The same one-shot scheme for a float32 vector (the flat parameter gradient of the training step): data[0..n) is
 * replaced by its sum over the ranks, added in rank order (bit-identical on every rank).  One CTA per 2048 floats stores its
 * chunk into every peer's buffer, raises a per-chunk flag and waits for the peers' flags of the same chunk.
 * peer_bufs_dev: `world` symmetric buffers of wf_p2p_allreduce_vec_buffer_bytes(world, n) bytes, zero-initialised.
 * step: 1, 2, 3, ... identical on all ranks; or step_dev (device uint64, initialised to 1) which the kernel reads and advances
 * itself (the last CTA to finish, counted in done_counter: device uint32, zero-initialised) -- a captured CUDA graph of the
 * call can then be replayed.  Time-outs behave as in wf_p2p_allreduce_sums (NaN, sticky error word = the last 64 bytes).
 * sums / sums_out (nullable, both or neither): four doubles (the loss sums of the step) all-reduced along with the vector. */
int64_t wf_p2p_allreduce_vec_buffer_bytes(int world, int64_t n);
int wf_p2p_allreduce_vec(const uint64_t* peer_bufs_dev, int rank, int world, uint64_t step, uint64_t* step_dev, float* data, int64_t n,
                         const double* sums, double* sums_out, uint32_t* done_counter, void* stream);

/* Single-GPU self-test of the same protocol: ALL `world` ranks are emulated by the warps of one CTA (one launch, so the
 * mutual flag waits cannot dead-lock on a device that serialises kernels).  peer_bufs_dev: `world` buffers on this
 * device; locals / outs: [world][4] doubles.  skip_rank >= 0: that rank never arrives (exercises the time-out path; pass a
 * short timeout_cycles, <= 0 selects the production ~2 s). */
int wf_p2p_allreduce_emulated(const uint64_t* peer_bufs_dev, int world, uint64_t step, const double* locals, double* outs,
                              int skip_rank, int64_t timeout_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVEFLOW_B200_H */
