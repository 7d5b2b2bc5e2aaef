/*
 * b200_fwdsim.h -- C ABI of the B200-native forward-simulation engine (libb200fwdsim.so).
 *
 * This is the drop-in boundary for ONE hot path of pyGSTi: bulk circuit-outcome simulation for the
 * `densitymx` evotype.  Every entry point below names the reference interface it replaces
 * (paths relative to the pyGSTi source tree @ 6822f14).  Plain C: opaque handles, raw pointers and
 * sizes, no C++/torch/numpy types.  Return value: 0 = ok, negative = error (see B200_E_*), message via
 * b200_last_error() (thread-local).  No exception crosses the boundary.  There is NO CPU fallback: with
 * no usable CUDA device every compute entry point fails with B200_E_CUDA.
 *
 * Data model (mirrors what the reference's Cython conversion layer hands its C++ loop,
 * pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:55-143):
 *   - a layout ATOM is a prefix table: rows [iDest, iStart, iCache, (prep), ops...] (pyx:55-77) plus, per
 *     row, the list of (effect index, element index) outcomes (pyx:179-181);
 *   - a MODEL is the dense real superoperators G[n_ops][d][d] (row-major), superkets rho[n_rho][d] and
 *     effect vectors E[n_eff][d], float64, in the Pauli-product basis (what OpCRep_Dense / StateCRep /
 *     EffectCRep_Dense hold: pygsti/evotypes/densitymx/opcreps.cpp:33-54, statecreps.cpp:20-45,
 *     effectcreps.cpp:29-45).  Non-dense reps are densified by the caller through
 *     LinearOperator.to_dense('HilbertSchmidt');
 *   - the parameter DERIVATIVES are a sparse matrix D[w][p] = d(member element w)/d(parameter p) in COO
 *     form; w indexes "W space": op g element (i,j) -> g*d*d+i*d+j ; prep r element i -> n_ops*d*d+r*d+i ;
 *     effect e element i -> n_ops*d*d+n_rho*d+e*d+i   (row-major vec as in matrixforwardsim.py:114-124;
 *     contents from member.deriv_wrt_params()/gpindices as in matrixforwardsim.py:126-170,1111-1136).
 */
#ifndef B200_FWDSIM_H
#define B200_FWDSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK            0
#define B200_E_INVALID    -1   /* bad argument / inconsistent table                         */
#define B200_E_CUDA       -2   /* CUDA runtime error or no device (no CPU fallback exists)  */
#define B200_E_NOMEM      -3   /* host or device allocation failed (-> MemoryError, as
                                  mapforwardsim.py:276-333 / resource_alloc.check_can_allocate_memory) */
#define B200_E_STATE      -4   /* call order violated (e.g. fill before model upload)        */
#define B200_E_UNSUPPORTED -5  /* dimension / size outside what the kernels are built for    */

typedef struct b200_ctx  b200_ctx;    /* one per GPU (per process or several per process)          */
typedef struct b200_atom b200_atom;   /* device-resident layout atom + its current model tensors   */

/* ---- library / device ------------------------------------------------------------------------ */
int         b200_version(void);
const char* b200_last_error(void);
int         b200_device_count(int* n_out);

/* One engine context per GPU.  `stream` = 0 creates a private non-blocking stream; otherwise an
 * existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) is adopted, so that the caller's
 * CUDA events bracket the kernels.  Replaces: the per-rank execution context the reference gets from
 * ResourceAllocation (pygsti/baseobjs/resourceallocation.py:28) -- one rank <-> one ctx. */
int b200_ctx_create(int device, void* stream, b200_ctx** out);
int b200_ctx_destroy(b200_ctx* ctx);
int b200_ctx_sync(b200_ctx* ctx);
/* kernels launched by this ctx since creation (the "gpu_launches" bench counter) */
int b200_ctx_launch_count(b200_ctx* ctx, int64_t* n_out);
/* Measurement aid (bench.py "roofline"): with phase timing on, the d = 16 Jacobian path brackets its three phases --
 * table preparation, trie chains, accumulate (the dominant, HBM-write-bound kernel) -- with CUDA events on the launching
 * stream.  b200_ctx_phase_ms synchronises the stream, returns the summed elapsed milliseconds of each phase over the calls
 * made since timing was switched on (or since the last read) and their number, and clears the record. */
int b200_ctx_phase_timing(b200_ctx* ctx, int on);
int b200_ctx_phase_ms(b200_ctx* ctx, double ms_out[3], int64_t* n_calls_out);
/* How b200_jtj / b200_jtj_dev contract the Jacobian (the reduction fill_jtj performs on the host, distlayout.py:1220-1359):
 *   -1 (default) automatic: the tcgen05 path with 8 digits from 4096 rows up, the FP64 path below;
 *    0  FP64 tensor cores (mma.sync DMMA SYRK, k_atb_dmma);
 *    8 / 7  5th-generation tensor cores through the Ozaki splitting (tcgen05.mma kind::i8, exact int32 accumulation in
 *       TMEM; 8 digits = 62 bits per entry relative to its column's largest magnitude, 7 digits = 55 bits).
 * The environment variable B200_JTJ (dmma | ozaki | ozaki7) sets the initial mode of new contexts. */
int b200_ctx_set_jtj_mode(b200_ctx* ctx, int mode);

/* ---- layout atom (one-time per layout) ---------------------------------------------------------
 * Replaces convert_maplayout / convert_dict_of_intlists / create_rhocache (pyx:55-101), which the
 * reference re-runs on every call; here the tables are uploaded once and live on the device.
 *   row_ptr[n_rows+1], row_ops[row_ptr[n_rows]] : CSR of op indices of each row's remainder
 *   row_istart[n_rows]  : cache slot the row starts from, -1 = starts from prep row_prep[k]
 *   row_icache[n_rows]  : cache slot the row's final state is stored in, -1 = none
 *   out_ptr[n_rows+1], out_eff[], out_el[] : CSR (per row, in row order) of outcomes:
 *                         effect index and atom-local element index (final_indices, pyx:179-181)
 * Rows must be in evaluation order (a row may only start from a cache slot written by an earlier row),
 * as PrefixTable guarantees (pygsti/layouts/prefixtable.py:65-101). */
int b200_atom_upload(b200_ctx* ctx, int dim, int n_ops, int n_rho, int n_eff,
                     int64_t n_rows, const int32_t* row_ptr, const int32_t* row_ops,
                     const int32_t* row_istart, const int32_t* row_prep, const int32_t* row_icache,
                     int32_t cache_size,
                     const int32_t* out_ptr, const int32_t* out_eff, const int32_t* out_el,
                     int64_t n_elements, b200_atom** out);
int b200_atom_free(b200_ctx* ctx, b200_atom* atom);
/* info[0]=n_rows info[1]=n_elements info[2]=table propagations (PrefixTable.num_state_propagations,
 * prefixtable.py:106) info[3]=expanded propagations (no prefix sharing) info[4]=max circuit depth
 * info[5]=W-space width  info[6]=n_params of the uploaded derivative map (or -1)
 * info[7]=1 if the derivative map is a unit partial permutation (fused fast path) */
int b200_atom_info(b200_atom* atom, int64_t info[8]);

/* ---- model tensors (every call of the reference re-reads the reps: pyx:163-167) ----------------
 * Replaces the rep lookups `model._circuit_layer_operator(lbl, typ)._rep` (pyx:164-167) and the
 * OpCRep_Dense / StateCRep / EffectCRep_Dense objects they wrap.  Host pointers, copied immediately. */
int b200_atom_set_model(b200_ctx* ctx, b200_atom* atom,
                        const double* G, const double* rho, const double* E);

/* FACTORED model: every gate (layer operation) is a product of small superoperators embedded on 1-2 qubits -- the device
 * form of the reference's non-dense reps OpCRep_Composed / OpCRep_Embedded (pygsti/evotypes/densitymx/opcreps.cpp:242-276,
 * 93-158; chosen by pyGSTi for dim > 64, pygsti/evotypes/evotype.py:97).  Gate g = factors [op_fptr[g], op_fptr[g+1]) applied
 * in that order; factor f acts with the row-major (4^nq x 4^nq) matrix at mats + f_moff[f] on the f_nq[f] (1 or 2) qubits
 * f_targets[4 f ..] (positions in state-space order, 0 = most significant base-4 digit of the state index).
 * b200_fill_probs then applies the factors directly (4 / 16 multiply-adds per state component instead of d); the dense
 * matrices the derivative paths need are built ON THE DEVICE from the same programs, so the host never calls to_dense on a
 * d x d operation.  Replaces b200_atom_set_model for such models (dim = 4^n, n >= 2).  rho / E as in b200_atom_set_model. */
int b200_atom_set_model_factored(b200_ctx* ctx, b200_atom* atom, int32_t n_factors, const int32_t* op_fptr,
                                 const int32_t* f_nq, const int32_t* f_targets, const int64_t* f_moff,
                                 const double* mats, int64_t n_mats, const double* rho, const double* E);

/* Sparse derivative map D (COO, duplicates summed by the engine), n_params = number of Jacobian
 * columns produced (the caller has already restricted/renumbered to its param_slice,
 * distforwardsim.py:130-144).  Replaces the per-parameter `model.set_parameter_values` loop of
 * mapfill_dprobs_atom (pyx:362-381) and `_doperation`/`_process_wrt_filter` of the Matrix simulator
 * (matrixforwardsim.py:89-170). */
int b200_atom_set_derivs(b200_ctx* ctx, b200_atom* atom, int64_t n_w, int32_t n_params,
                         int64_t nnz, const int32_t* rows, const int32_t* cols, const double* vals);
/* Derivative map in FACTOR space for an atom whose gates were set with b200_atom_set_model_factored: rows index
 * [mats (the factor matrices of that call) | rho (n_rho x d) | E (n_eff x d)], COO like b200_atom_set_derivs.  With it
 * the Jacobian of d = 64 / 256 atoms is evaluated straight from the factor programs -- the device form of the
 * reference's OpCRep_Embedded / OpCRep_Composed (opcreps.cpp:93-158, 242-276) differentiated factor by factor
 * (EmbeddedOp.deriv_wrt_params is the embedded operation's own derivative, embeddedop.py) -- without dense d x d
 * products.  Describes the same parameter block as the dense map, if one is set (call it AFTER b200_atom_set_derivs;
 * that call drops a factor-space map).  Without a dense map only the Jacobian / J^T J calls are available. */
int b200_atom_set_derivs_factored(b200_ctx* ctx, b200_atom* atom, int64_t n_wf, int32_t n_params, int64_t nnz,
                                  const int32_t* rows, const int32_t* cols, const double* vals);

/* ---- on-device model update for members AFFINE in their parameters (first part) ----------
 * FullArbitraryOp / FullTPOp / FullState / TPState / static members and (un)constrained POVM effects are affine in the
 * parameter vector:  M(theta) = M_const + D theta,  D = the derivative map of b200_atom_set_derivs over ALL parameters.
 * b200_atom_bind_params fixes M_const = M - D theta0 from the model tensors currently on the device (b200_atom_set_model)
 * and the parameter vector theta0 they were computed from.  b200_atom_set_params[_dev] then replaces, per optimizer
 * iteration, OpModel.from_vector (pygsti/models/model.py:1163-1196) + the members' to_dense + the upload by one kernel
 * (rows of D summed in a fixed order: deterministic; exact for `full` members, whose rows have a single unit entry).
 * b200_atom_get_model copies M = [G | rho | E] back (n_w doubles) -- used by the parity tests.
 * Errors: B200_E_STATE if no model / derivative map is set, or the map was uploaded for a parameter sub-block. */
int b200_atom_bind_params(b200_ctx* ctx, b200_atom* atom, int32_t n_params, const double* theta0);
int b200_atom_set_params(b200_ctx* ctx, b200_atom* atom, int32_t n_params, const double* theta);
int b200_atom_set_params_dev(b200_ctx* ctx, b200_atom* atom, int32_t n_params, const double* d_theta);
int b200_atom_get_model(b200_ctx* ctx, b200_atom* atom, int64_t n_w, double* M_out);

/* ---- Lindblad-parameterised members on the device (second part) ----------------------------------
 * For every member  exp(L_e) composed with a static part  (kind 0: op G = exp(L) T; 1: state exp(L) rho0; 2: effect exp(L)^T e0),
 * L_e = Re sum_i c_i B_i, returns the dense member and its derivative w.r.t. the generator's parameters:
 * replaces LindbladErrorgen._update_rep / deriv_wrt_params (lindbladerrorgen.py:700-708, 1342-1384), ExpErrorgenOp._update_rep /
 * deriv_wrt_params (experrorgenop.py:114-125, 213-262: scipy expm + commutator series) and the Composed* to_dense / deriv_wrt_params.
 * The coefficients c and their Jacobian dc (complex, real / imaginary parts separately) come from the host
 * (LindbladCoefficientBlock.from_vector / deriv_wrt_params).  All arrays are host arrays, concatenated over generators / members:
 *   B_*  [sum_e n_coeff_e][d*d],  c_* [sum_e n_coeff_e],  dc_* per generator [n_coeff_e][n_par_e],
 *   stat: d*d doubles per op, d per state / effect;  val_out: same sizes;  dval_out: per member [size][n_par of its generator]. */
int b200_lindblad_members(b200_ctx* ctx, int d, int n_eg, const int32_t* eg_ncoeff, const int32_t* eg_npar,
                          const double* B_re, const double* B_im, const double* c_re, const double* c_im,
                          const double* dc_re, const double* dc_im,
                          int n_mem, const int32_t* m_kind, const int32_t* m_eg, const double* stat,
                          double* val_out, double* dval_out);

/* ---- the hot path, HOST buffers (copies inside the call) ---------------------------------------
 * b200_fill_probs   replaces mapfill_probs_atom + dm_mapfill_probs (pyx:149-287):
 *     out[el * out_stride] = E . G_L ... G_1 rho        for every element of the atom.
 * b200_fill_dprobs  replaces mapfill_dprobs_atom (pyx:290-383) with the ANALYTIC Jacobian
 *     (= MatrixForwardSimulator._dprobs_from_rho_e, matrixforwardsim.py:1059-1139):
 *     out[el * row_stride + p] = d p_el / d theta_p ,  p in [0, n_params);  columns outside are untouched.
 *     probs_out may be NULL; otherwise it receives the probabilities as well (pr_array_to_fill,
 *     distforwardsim.py:127-128).  Strides are in doubles. */
int b200_fill_probs(b200_ctx* ctx, b200_atom* atom, double* out, int64_t out_stride);
int b200_fill_dprobs(b200_ctx* ctx, b200_atom* atom, double* out, int64_t row_stride,
                     double* probs_out, int64_t probs_stride);

/* Reference-semantics forward differences (pyx:349-378): base pass + one pass per parameter with the
 * members perturbed to M + eps * dM/dtheta_p (exact reproduction of the reference for members linear in
 * their parameters, first-order for the others), out = (p2 - p) / eps. */
int b200_fill_dprobs_fd(b200_ctx* ctx, b200_atom* atom, double eps, double* out, int64_t row_stride,
                        double* probs_out, int64_t probs_stride);

/* Hessian block for members LINEAR in their parameters (d2M/dtheta2 = 0: full, TP, static members),
 * replaces MapForwardSimulator._mapfill_hprobs_atom (mapforwardsim.py:394-438) analytically:
 *     out[(el * n1 + a) * n2 + b] = d2 p_el / d theta_{p1[a]} d theta_{p2[b]}
 * p1/p2 index columns of the uploaded derivative map. */
int b200_fill_hprobs_linear(b200_ctx* ctx, b200_atom* atom, int32_t n1, const int32_t* p1,
                            int32_t n2, const int32_t* p2, double* out);
/* Hessian block for ARBITRARY members, fully analytic: the linear-member expression above plus the second-derivative
 * term  sum_w (d p_el / d M_w) * d2 M_w / d theta_{p1[a]} d theta_{p2[b]},  with the members' second derivatives given
 * as COO entries (h_rows[t] = w, h_a[t] = a, h_b[t] = b, h_vals[t]) taken from member.hessian_wrt_params() -- the same
 * quantity MatrixForwardSimulator reads in `_hoperation` and `_hprobs_from_rho_e` (matrixforwardsim.py:172-218,
 * 1141-1287).  Replaces the reference Map simulator's finite differences of finite differences
 * (mapforwardsim.py:394-438, (B1+1)(B2+1) table passes per block) for CPTPLND / H+S / ... models.  nnz2 = 0 is
 * b200_fill_hprobs_linear. */
int b200_fill_hprobs(b200_ctx* ctx, b200_atom* atom, int32_t n1, const int32_t* p1, int32_t n2, const int32_t* p2,
                     int64_t nnz2, const int32_t* h_rows, const int32_t* h_a, const int32_t* h_b, const double* h_vals,
                     double* out);

/* "Next" row: one block of the MLE Hessian, reduced on the device.  Replaces
 * `_hessian_from_block` (pygsti/objectivefns/objectivefns.py:4914-4990) applied to the rectangles produced by
 * `_iter_atom_hprobs_by_rectangle` (distforwardsim.py:304-340; loop in `_construct_hessian`, objectivefns.py:1576-1693):
 *     out[a * n2 + b] = sum_el  w_h[el] * d2 p_el / d theta_{p1[a]} d theta_{p2[b]}
 *                             + w_d[el] * d p_el / d theta_{p1[a]} * d p_el / d theta_{p2[b]}
 * with w_h = raw_objfn.dterms(probs, ...) and w_d = raw_objfn.hterms(probs, ...) computed by the caller from the
 * probabilities and the data, exactly as the reference does.  The (n_elements x n1 x n2) Hessian-of-probabilities block
 * and the Jacobian never leave the device: only n1*n2 doubles are copied out (host buffer `out`).  Arguments p1..h_vals
 * as in b200_fill_hprobs. */
int b200_hessian_block(b200_ctx* ctx, b200_atom* atom, int32_t n1, const int32_t* p1, int32_t n2, const int32_t* p2,
                       int64_t nnz2, const int32_t* h_rows, const int32_t* h_a, const int32_t* h_b, const double* h_vals,
                       const double* w_h, const double* w_d, double* out);

/* ---- the hot path, DEVICE buffers (no copies; asynchronous on the ctx stream) ------------------ */
int b200_fill_probs_dev(b200_ctx* ctx, b200_atom* atom, double* d_out);
int b200_fill_dprobs_dev(b200_ctx* ctx, b200_atom* atom, double* d_out, int64_t ld, double* d_probs);

/* ---- multi-GPU, one process per GPU: Jacobian fill FUSED with the exchange over NVLink peer memory ------------------
 * The reference reassembles per-rank shards with gather_local_array -> Allgatherv (pygsti/baseobjs/resourceallocation.py:
 * 323-329) after every rank has filled its own rows.  Here the rows can travel while they are produced: every rank holds the
 * whole sharded array (b200_peer_alloc, a cudaMalloc'd buffer exported through CUDA IPC), maps the arrays of its peers
 * (b200_peer_open) and b200_fill_dprobs_bcast_dev makes the kernel epilogue store each finished Jacobian row / probability
 * into the local array AND, at the same offset, into every peer's array (plain stores over NVLink / NVSwitch).  After a
 * stream synchronisation on every rank plus one barrier, every rank holds every row -- no separate all-gather pass.
 *   d_out / d_probs        : this rank's slot inside ITS OWN array (as for b200_fill_dprobs_dev)
 *   d_out_peers[i] / d_probs_peers[i] : the address of the SAME slot inside peer i's array (n_peers <= 7; d_probs_peers may be NULL)
 * Kernels without a fused peer epilogue (the general W.D path) fill locally and forward the slot with asynchronous
 * peer-to-peer copies on the same stream: same result, same call. */
#define B200_MAX_PEERS 7
int b200_peer_alloc(b200_ctx* ctx, int64_t bytes, void** d_ptr_out, unsigned char handle_out[64]);
int b200_peer_open(b200_ctx* ctx, const unsigned char handle[64], void** d_ptr_out);
int b200_peer_close(b200_ctx* ctx, void* d_ptr);
int b200_peer_free(b200_ctx* ctx, void* d_ptr);
int b200_fill_dprobs_bcast_dev(b200_ctx* ctx, b200_atom* atom, double* d_out, int64_t ld, double* d_probs,
                               int n_peers, double* const* d_out_peers, double* const* d_probs_peers);

/* ---- "next" row: the objective-function Jacobian fill, fused ------------------------------
 * b200_fill_dprobs_scaled  = b200_fill_dprobs followed by the row scaling the objective functions apply to it,
 *     out[el, p] = row_scale[el] * d p_el / d theta_p
 * replacing `dprobs *= dg_probs[:, None]` of TimeIndependentMDCObjectiveFunction.dterms and
 * `jac *= p5over_lsvec[:, None]` of dlsvec (pygsti/objectivefns/objectivefns.py:4609-4616, 4644-4649): the scaling
 * is applied in the kernel epilogue instead of a second pass over a (2.97 GB at BASELINE size) host array.
 * row_scale: host vector [n_elements] (the caller computes it from the probabilities and the data, as the reference
 * does); NULL = no scaling.
 * b200_jtj  keeps the scaled Jacobian J on the device and returns only  JTJ = J^T J  [n_params x n_params, row-major,
 * full symmetric] and, if f != NULL,  JTf = J^T f  [n_params]: what the Levenberg-Marquardt step consumes
 * (DistributableCOPALayout.fill_jtj / fill_jtf, pygsti/layouts/distlayout.py:1220-1359;
 * pygsti/optimize/simplerlm.py:677-678).  jtf_out may be NULL. */
int b200_fill_dprobs_scaled(b200_ctx* ctx, b200_atom* atom, const double* row_scale, double* out, int64_t row_stride,
                            double* probs_out, int64_t probs_stride);
int b200_jtj(b200_ctx* ctx, b200_atom* atom, const double* row_scale, const double* f,
             double* jtj_out, double* jtf_out);
/* Device-buffer variant (asynchronous on the ctx stream; d_row_scale / d_f / d_jtf may be NULL): with one process per
 * GPU each rank reduces its own element shard and the [n_params x n_params] partial sums are added with ONE
 * all-reduce (NCCL) -- the analogue of the reference summing per-rank `fill_jtj` blocks through its shared-memory /
 * MPI reduction (pygsti/layouts/distlayout.py:1220-1359, resourceallocation.py:331-376 allreduce_sum).
 * See pygsti_b200/dist.py: allreduce_jtj. */
int b200_jtj_dev(b200_ctx* ctx, b200_atom* atom, const double* d_row_scale, const double* d_f,
                 double* d_jtj, double* d_jtf);

/* ---- pinned host memory for zero-staging transfers --------------------------------------------- */
int b200_host_alloc(void** out, int64_t bytes);     /* cudaHostAlloc */
int b200_host_free(void* p);
int b200_host_register(void* p, int64_t bytes);     /* cudaHostRegister an existing buffer (cached) */
int b200_host_unregister(void* p);

#ifdef __cplusplus
}
#endif
#endif /* B200_FWDSIM_H */
