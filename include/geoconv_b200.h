/*
 * geoconv_b200 - C ABI of the B200-native ConvIntrinsic (+AngularMaxPooling) hot path.
 *
 * The reference (andreasMazur/geoconv) has NO plugin / FFI seam for this path: the boundary is the
 * Python class API of geoconv.pytorch.layers.  This header is the seam the
 * drop-in layers bind to (ctypes); every entry point names the reference code
 * it replaces.  All paths are relative to /root/reference/src/geoconv/pytorch/layers/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - tensors are dense, row-major, fp32 / int32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued on it and nothing synchronises the host unless stated;
 *   - the library never allocates device memory: callers pass outputs and a workspace whose
 *     size the matching *_workspace_bytes() query returns;
 *   - return value 0 = success; otherwise a GCB_E* code, with text from gcb_last_error().
 *     There is no CPU fallback anywhere.
 *
 * Symbols: V_out rows of the barycentric tensor (output vertices), V_src rows of the signal
 * (gather source; == V_out for a whole mesh, local+halo rows for a vertex-row shard),
 * R radial, A angular, F input features, T templates, n_rot rotations, KRA = R*A*F.
 */
#ifndef GEOCONV_B200_H_
#define GEOCONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCB_ABI_VERSION 7
#define GCB_MAX_ROT 32

/* error codes */
#define GCB_OK 0
#define GCB_EINVAL 1      /* bad argument (shape, null pointer, alignment) */
#define GCB_ECUDA 2       /* CUDA runtime error; text in gcb_last_error() */
#define GCB_EUNSUPPORTED 3 /* shape outside what the selected precision path supports */
#define GCB_EWORKSPACE 4  /* workspace too small */

/* activation ids: the six entries of ACTIVATIONS, conv_intrinsic.py:10-17 */
#define GCB_ACT_RELU 0
#define GCB_ACT_ELU 1
#define GCB_ACT_LEAKY_RELU 2
#define GCB_ACT_SELU 3
#define GCB_ACT_SIGMOID 4
#define GCB_ACT_TANH 5

/* arithmetic of the contraction */
#define GCB_PREC_FP32 0   /* CUDA-core FFMA, plain fp32 */
#define GCB_PREC_TF32 1   /* tcgen05 kind::tf32, one pass (<=2e-3 normalised error) */
#define GCB_PREC_TF32X3 2 /* tcgen05 kind::tf32, 3-term split, fp32-faithful (<=1e-5) */
#define GCB_PREC_F16X3 3  /* tcgen05 kind::f16, operands scaled by powers of two and split into fp16 hi + lo
                             (same 11-bit significands as tf32), 3 products at twice the tf32 rate,
                             fp32-faithful (<=1e-5); shapes: gcb_tc_half_supported */

int gcb_abi_version(void);
const char* gcb_last_error(void);
/* on != 0: a collective is going to run beside the library's persistent GEMMs (data-parallel all-reduce started inside
 * the backward): they size their grids to the tile rounds they need (e.g. 135 CTAs for 270 tiles) instead of one CTA
 * per SM, which costs nothing and leaves whole SMs to the collective.  Process-wide, off by default. */
void gcb_yield_sms(int on);
/* Kernel launches issued by this library in this process so far (bench.py's gpu_launches). */
int64_t gcb_launch_count(void);
/* Measurement hooks (bench.py's roofline leg): while enabled, CUDA events bracket the dominant
 * kernels on their launch stream.  which: 0 = forward contraction, 1 = backward dW contraction,
 * 2 = backward dI contraction, 3 = backward CSR gather-reduce.  gcb_profile_read synchronises
 * on the recorded events and returns their summed duration and count since the last enable. */
void gcb_profile_enable(int on);
int gcb_profile_read(int which, double* ms_total, int64_t* launches);
/* 1 if the tcgen05 kernels can serve this shape, else 0 (then use GCB_PREC_FP32). */
int gcb_tc_supported(int32_t R, int32_t A, int32_t F, int32_t T, int32_t n_rot);
/* 1 if GCB_PREC_F16X3 can serve this shape (needs F % 16 == 0 on top of gcb_tc_supported), else use
 * GCB_PREC_TF32X3. */
int gcb_tc_half_supported(int32_t R, int32_t A, int32_t F, int32_t T, int32_t n_rot);

/* ---- GPC systems -> barycentric coordinates (the preprocessing step before the path) -----
 * Replaces geoconv/preprocessing/barycentric_coordinates.py:127-175 (with :7-69): for every GPC
 * system v and template vertex q the first triangle, in the system's face order, that contains
 * q; float64 with the reference's operation order.
 * tri_xy [n_tri,3,2] f64 cartesian GPC coordinates of the triangle corners, tri_idx [n_tri,3] i32
 * their mesh vertex ids, tri_ptr [V+1] i64 (system v owns triangles tri_ptr[v] .. tri_ptr[v+1]),
 * template_xy [RA,2] f64 cartesian template vertices; bary_out [V,RA,3,2] f64 = (index, weight),
 * zeros where no triangle contains the template vertex.                                     */
int gcb_bary_from_gpc(const double* tri_xy, const int32_t* tri_idx, const int64_t* tri_ptr, int32_t V,
                      const double* template_xy, int32_t RA, double* bary_out, void* stream);

/* ---- mesh -> GPC systems (the wavefront; the step before gcb_bary_from_gpc) ----------------
 * Replaces, for all source vertices in one launch, GPCSystemGroup.compute_gpc_system
 * (geoconv/preprocessing/gpc_system_group.py:42-122) with GPCSystem (gpc_system.py:12-337),
 * compute_distance_and_angle (gpc_system_utils.py:150-212) and the reference's C extension
 * (preprocessing/c_extension/c_extension.c:11-145).  One warp per source vertex, float64.
 * Mesh tables: vertices / normals [n_vertices,3] f64; nbr_ptr [n_vertices+1], nbr [nnz] i32 = the
 * neighbours of every vertex in ASCENDING vertex id (the first one is a system's reference
 * direction); opposite [nnz,2] i32 = for the edge {v, nbr} the third vertex of its first / second
 * incident face in ascending face index, -1 where there is none (boundary edge).
 * sources [n_sources] i32 (NULL = vertices 0 .. n_sources-1).  u_max / eps as in the reference
 * (GPCSystemGroup(eps=1e-6).compute(u_max=0.04)).  max_patch in [8, 1023] = capacity of a system's
 * vertex table; a system owns max_patch rows of out_vid / out_rho / out_theta and 3*max_patch rows
 * of out_tri [.,3] i32 (vertex ids, sorted) / out_xy [.,3,2] f64 (cartesian GPC coordinates), filled
 * in the order the reference inserts them; out_nv / out_nf [n_sources] i32 say how many.  Vertices
 * the wavefront touched but never accepted keep rho = inf, theta = -1 (the reference's fill
 * values).  out_flags [n_sources] i32: GCB_GPC_* bits.                                         */
#define GCB_GPC_OVERFLOW 1    /* a table was too small: raise max_patch (results of that system are partial) */
#define GCB_GPC_NONMANIFOLD 2 /* an edge with more than two faces                                */
#define GCB_GPC_FRAGILE 4     /* the crossing guard judged collinear segments: a verdict that is rounding
                                 noise in the reference as well (gpc_system.py:322-337)           */
int64_t gcb_gpc_smem_per_system(int32_t max_patch);
int gcb_gpc_wavefront(const double* vertices, const double* normals, const int32_t* nbr_ptr,
                      const int32_t* nbr, const int32_t* opposite, int32_t n_vertices,
                      const int32_t* sources, int32_t n_sources, double u_max, double eps,
                      int32_t max_patch, int32_t* out_nv, int32_t* out_vid, double* out_rho,
                      double* out_theta, int32_t* out_nf, int32_t* out_tri, double* out_xy,
                      int32_t* out_flags, void* stream);

/* ---- between two convolutions: BatchNorm1d -> Dropout ------------------------------------------
 * The chain the reference's network runs between the pooled output of one block and the next
 * convolution (geoconv_examples/mpi_faust/pytorch/model.py:93-105, :128-132: do -> isc -> amp -> bn).
 * z, out, d_out, d_z [V,T] fp32 (T % 4 == 0, 16-byte aligned); mean, invstd, gamma, beta, d_gamma, d_beta [T].
 * gcb_bn_stats: training-mode statistics of z (biased variance inside invstd = 1/sqrt(var + eps)), fixed
 *   summation order; running_mean / running_var (both or neither, nullable) are updated in place as
 *   torch does: new = (1 - momentum) * old + momentum * (mean | unbiased variance).
 * gcb_bn_dropout_fwd: out = dropout_p(gamma * (z - mean) * invstd + beta); element e is kept (and scaled
 *   by 1/(1-p)) iff word e%4 of Philox4x32-10(counter = (e/4, stream_id), key = seed) < (1-p) * 2^32.
 *   The mask is never stored: gcb_bn_dropout_bwd regenerates it from (seed, stream_id).  p = 0: no dropout.
 * gcb_bn_dropout_bwd: d_beta = sum dY, d_gamma = sum dY * xhat, and (training != 0)
 *   d_z = gamma * invstd * (dY - d_beta/V - xhat * d_gamma/V), else d_z = gamma * invstd * dY, with
 *   dY = mask * d_out / (1-p).
 * workspace: gcb_bn_workspace_bytes(V, T); its first 256 bytes must be zero before the first use (the
 * kernels leave them zero).                                                                      */
size_t gcb_bn_workspace_bytes(int32_t V, int32_t T);
int gcb_bn_stats(const float* z, int32_t V, int32_t T, float eps, float* mean, float* invstd,
                 float* running_mean, float* running_var, float momentum, void* workspace,
                 size_t workspace_bytes, void* stream);
int gcb_bn_dropout_fwd(const float* z, int32_t V, int32_t T, const float* mean, const float* invstd,
                       const float* gamma, const float* beta, float p, uint64_t seed, uint64_t stream_id,
                       float* out, void* stream);
int gcb_bn_dropout_bwd(const float* d_out, const float* z, int32_t V, int32_t T, const float* mean,
                       const float* invstd, const float* gamma, int training, float p, uint64_t seed,
                       uint64_t stream_id, float* d_z, float* d_gamma, float* d_beta, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- bary -> (idx, w) ------------------------------------------------------------------
 * Replaces conv_intrinsic.py:213-221 (cast to fp32, `[...,0].long()`, `[...,1]`).
 * bary: [n_slots, 2] fp32 or fp64 (n_slots = V_out*R*A*3).  idx = trunc-toward-zero of the fp32
 * cast.  *oob_count (int32, device, caller-zeroed) += number of idx outside [0, v_src).        */
int gcb_prep_bary(const void* bary, int bary_is_f64, int64_t n_slots, int32_t v_src,
                  int32_t* idx, float* w, int32_t* oob_count, void* stream);

/* ---- CSR inversion (neighbour -> vertex), for the deterministic dSignal reduction -------
 * Replaces the atomic scatter_add_ that autograd runs for torch.gather (conv_intrinsic.py:237).
 * row_ptr [v_src+1], entries [n_slots]: slot ids grouped by source vertex, ascending inside a
 * group (stable => bit-reproducible).  Zero-weight slots (the (0,0,0) fallback,
 * barycentric_coordinates.py:69) are dropped; row_ptr[v_src] = number kept.
 * R, A: the template size, or 0, 0.  With the template size the entries of a row are additionally grouped by the
 * radial ring of their slot (ascending slot id inside a ring) - the order the tensor-core backward wants, because a
 * rotation moves a template vertex only inside its ring; pass GCB_CSR_RING_GROUPED to the backward then.
 * stats (nullable, 8 uint32): [1] = max |w| as fp32 bits, [2] = longest row - the properties of the tensor from
 * which the fp16-split backward bounds its operand G; pass them to gcb_conv_amp_bwd.           */
size_t gcb_csr_workspace_bytes(int64_t n_slots, int32_t v_src);
int gcb_csr_build(const int32_t* idx, const float* w, int64_t n_slots, int32_t v_src, int32_t R, int32_t A,
                  int32_t* row_ptr, int32_t* entries, uint32_t* stats, void* workspace,
                  size_t workspace_bytes, void* stream);
/* csr_flags of gcb_conv_bwd / gcb_conv_amp_bwd: how the entries of a row are ordered */
#define GCB_CSR_RING_GROUPED 1 /* built with R, A > 0: grouped by radial ring, ascending slot id inside a ring */

/* ---- prior folding ---------------------------------------------------------------------
 * Replaces the per-call einsum "raxy,kxyf->kraf" (conv_intrinsic.py:194) by folding the
 * rotation-symmetric prior into the weights once per step:
 *   transpose == 0:  out[t,x,y,f] = sum_{r,a} in[t,r,a,f] * prior[r,a,x,y]      (W -> Weff)
 *   transpose == 1:  out[t,r,a,f] = sum_{x,y} in[t,x,y,f] * prior[r,a,x,y]      (dWeff -> dW) */
int gcb_fold_prior(const float* in, const float* prior, float* out, int32_t T, int32_t R,
                   int32_t A, int32_t F, int transpose, void* stream);

/* ---- forward ---------------------------------------------------------------------------
 * Replaces ConvIntrinsic.forward (conv_intrinsic.py:118-171) incl. _signal_retrieval (:198-240)
 * and, when z/argmax are given, AngularMaxPooling.forward (angular_max_pooling.py:27-29):
 *   Y[k,i,t] = act( X[k]·Wself[t] + sum_{x,y,f} Weff[t,x,(y+rot_off[i])%A,f]*I[k,x,y,f] + b[t] )
 *   I[k,x,y,:] = sum_j w[k,x,y,j] * X[idx[k,x,y,j], :]
 *   argmax[k] = first i maximising ||Y[k,i,:]||_2 ; Z[k,:] = Y[k,argmax[k],:]
 * x [V_src,F]; idx,w [V_out,R,A,3]; w_eff [T,R,A,F]; w_self [T,F]; bias [T];
 * rot_off_host: n_rot ints on the HOST (the rotation list arange(0,A,delta) or `orientations`).
 * y [V_out,n_rot,T] (may be NULL if z given); z [V_out,T] and argmax [V_out] (both or neither).
 * w_eff == NULL means an identically-zero prior (ConvZero, conv_zero.py:12-15): the neighbour
 * term is skipped and only the centre term is computed.
 * interp_out (nullable) [V_out,R,A,F]: the tensor-core path stores the interpolations I there
 * (the reference materialises the same tensor, conv_intrinsic.py:240); handing it to
 * gcb_conv_bwd saves the backward one full gather pass.  Ignored by the CUDA-core path.
 * records / tables (nullable): preparation that depends only on the barycentric tensor / on the shape and rotation
 * list (see below); when NULL the tensor-core path redoes it in the workspace on every call.
 * stats_out (nullable, 8 uint32): operand maxima the fp16-split forward computes anyway ([0] max|X|, [2]
 * max|[Weff|Wself]| as fp32 bits; all zero from the other paths) - gcb_conv_amp_bwd of the same step reuses them.
 * row0: output row k is vertex row0 + k, i.e. its centre term reads x[row0 + k] (0 for a whole mesh or a shard whose
 * rows lead the signal; > 0 lets a caller convolve a block of rows in the middle of the signal, e.g. the boundary
 * rows of a vertex-row shard after its interior rows - geoconv_b200/distributed.py).  Forward only.               */
size_t gcb_conv_fwd_workspace_bytes(int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F,
                                    int32_t T, int32_t n_rot, int precision);
int gcb_conv_fwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                 const float* w_self, const float* bias, const int32_t* rot_off_host,
                 int32_t n_rot, int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F,
                 int32_t T, int act, int precision, float* y, float* z, int32_t* argmax,
                 float* interp_out, const void* records, const void* tables, uint32_t* stats_out, int32_t row0,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- forward preparation a caller may cache ----------------------------------------------------
 * The reference rebuilds everything per call (conv_intrinsic.py:198-240).  Two pieces of the tensor-core
 * forward's preparation do not depend on the signal or the weights:
 *   records: the (idx, w) pairs repacked slot-major and padded to 128-vertex tiles (what the gather producers
 *            stage in shared memory) plus max_slots(|w0|+|w1|+|w2|) - a function of the barycentric tensor;
 *   tables:  weight-ring schedules and K-split boundaries - a function of (v_out, R, A, F, T, rotations, precision).
 * Build them once and pass them to gcb_conv_fwd.                                                 */
size_t gcb_fwd_records_bytes(int32_t v_out, int32_t R, int32_t A);
int gcb_fwd_pack_records(const int32_t* idx, const float* w, int32_t v_out, int32_t R, int32_t A, void* records,
                         void* stream);
size_t gcb_fwd_tables_bytes(void);
int gcb_fwd_build_tables(const int32_t* rot_off_host, int32_t n_rot, int32_t v_out, int32_t R, int32_t A, int32_t F,
                         int32_t T, int precision, void* tables, void* stream);

/* ---- AngularMaxPooling on a materialised Y -----------------------------------------------
 * Replaces angular_max_pooling.py:27-29 and its autograd (index_put).                        */
int gcb_amp_fwd(const float* y, int32_t v, int32_t n_rot, int32_t T, float* z, int32_t* argmax,
                void* stream);
int gcb_amp_bwd(const float* grad_z, const int32_t* argmax, int32_t v, int32_t n_rot, int32_t T,
                float* grad_y, void* stream);

/* ---- AngularMinPooling / AngularAvgPooling (and max again) on a materialised Y ---------------
 * Replaces angular_min_pooling.py:27-29 (argmin of the row norms, first minimum wins) and
 * angular_avg_pooling.py:27 (mean over rotations) with their autograd.  `arg` may be NULL for
 * GCB_POOL_AVG.                                                                                */
#define GCB_POOL_MAX 0
#define GCB_POOL_MIN 1
#define GCB_POOL_AVG 2
int gcb_apool_fwd(const float* y, int32_t v, int32_t n_rot, int32_t T, int mode, float* z, int32_t* arg,
                  void* stream);
int gcb_apool_bwd(const float* grad_z, const int32_t* arg, int32_t v, int32_t n_rot, int32_t T, int mode,
                  float* grad_y, void* stream);

/* h = grad * act'(.) with act' written in terms of the activation OUTPUT `out` (n elements). */
int gcb_act_grad(const float* grad, const float* out, int64_t n, int act, float* h, void* stream);

/* rot_sel[k] = rot_off_host[argmax[k]] mod A, in [0, A)  (int32, device; torch.roll accepts any integer shift,
 * conv_intrinsic.py:163, so `orientations` may hold negative values or values >= A) */
int gcb_select_rot(const int32_t* argmax, const int32_t* rot_off_host, int32_t n_rot, int32_t A, int32_t v,
                   int32_t* rot_sel, void* stream);

/* ---- backward --------------------------------------------------------------------------
 * Replaces autograd of conv_intrinsic.py:144-171/:226-240 for ONE rotation per vertex:
 * h [V_out,T] = upstream grad times act', rot_sel[k] the angular shift applied for vertex k.
 *   d_w_eff[t,x,y',f] (+)= sum_k h[k,t] * I[k,x,(y'-rot_sel[k])%A,f]
 *   d_w_self[t,f]     (+)= sum_k h[k,t] * X[k,f]
 *   d_bias[t]         (+)= sum_k h[k,t]
 *   d_x[v,f]          (+)= sum_{slots s->v} w[s]*(sum_t h[k_s,t]*Weff[t,x_s,(y_s+rot_sel[k_s])%A,f])
 *                          + [v<V_out] sum_t h[v,t]*Wself[t,f]
 * accumulate != 0 adds into the outputs (used to back-propagate a dense dY one rotation at a
 * time); == 0 overwrites.  The AMP-sparse backward is a single call with rot_sel from
 * gcb_select_rot.  d_x uses the CSR from gcb_csr_build: fixed summation order, no atomics.
 * d_x == NULL skips the dSignal stages (signal does not require grad); w_eff == NULL as above
 * (then d_w_eff is zero-filled unless accumulating and the CSR pointers may be NULL).
 * precision: GCB_PREC_F16X3 runs the route through the CSR on fp16 hi/lo splits and every other
 * route (no CSR, w_eff == NULL) like GCB_PREC_TF32X3; shapes the tensor kernels cannot tile
 * (F % 32 or T % 32 != 0) take the CUDA-core path whatever the precision.                       */
size_t gcb_conv_bwd_workspace_bytes(int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F,
                                    int32_t T, int precision);
/* Host-only self-check (no device work): 1 if every internal buffer layout gcb_conv_bwd can choose for this shape
 * and precision (tf32 / fp16 operand routes, centre-only route) fits inside gcb_conv_bwd_workspace_bytes. */
int gcb_conv_bwd_workspace_selfcheck(int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F, int32_t T,
                                     int precision);
int gcb_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                 const float* w_self, const float* h, const int32_t* rot_sel,
                 const int32_t* csr_row_ptr, const int32_t* csr_entries, int csr_flags, int32_t v_out,
                 int32_t v_src, int32_t R, int32_t A, int32_t F, int32_t T, int precision,
                 int accumulate, const float* interp /* nullable: interp_out of the forward */, float* d_x,
                 float* d_w_eff, float* d_w_self, float* d_bias, void* workspace, size_t workspace_bytes,
                 void* stream);

/* Dense backward: a gradient for EVERY rotation (the layer used without AngularMaxPooling, or under
 * AngularMinPooling / AngularAvgPooling - autograd of conv_intrinsic.py:144-171 over all n_rot rotations) in ONE pass:
 * h_all [V_out, n_rot, T] = dY * act'(Y) (gcb_act_grad on the whole tensor); the operand G collects, per source vertex
 * and rotated slot, the contributions of all rotations, and the two GEMMs behind it are the size of one rotation.
 * rot_off_host: the n_rot DISTINCT rotation offsets (mod A) on the host.  Needs gcb_conv_bwd_dense_supported() and a
 * CSR built with the template size (GCB_CSR_RING_GROUPED); otherwise call gcb_conv_bwd once per rotation.
 * Workspace: gcb_conv_bwd_workspace_bytes.                                                              */
int gcb_conv_bwd_dense_supported(int32_t R, int32_t A, int32_t F, int32_t T, int precision);
int gcb_conv_bwd_dense(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                       const float* w_self, const float* h_all, const int32_t* rot_off_host, int32_t n_rot,
                       const int32_t* csr_row_ptr, const int32_t* csr_entries, int csr_flags, int32_t v_out,
                       int32_t v_src, int32_t R, int32_t A, int32_t F, int32_t T, int precision,
                       int accumulate, float* d_x, float* d_w_eff, float* d_w_self, float* d_bias,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- backward of conv + AngularMaxPooling in one call -----------------------------------------
 * Replaces autograd of angular_max_pooling.py:27-29 followed by conv_intrinsic.py:144-171/:226-240: starts from
 * grad_z [V_out,T] (gradient of the pooled output), z [V_out,T] (the pooled output itself: the activation
 * derivative is taken from it) and argmax [V_out]; equivalent to gcb_act_grad + gcb_select_rot + gcb_conv_bwd
 * with accumulate = 0, but ONE preparation kernel produces h, the selected shifts, d_bias and - in the fp16-split
 * mode - the operand scales and both operand splits.  csr_stats (gcb_csr_build) and fwd_stats (gcb_conv_fwd's
 * stats_out, same signal and weights) are optional: without them the maxima are recomputed.       */
size_t gcb_conv_amp_bwd_workspace_bytes(int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F, int32_t T,
                                        int precision);
int gcb_conv_amp_bwd(const float* grad_z, const float* z, const int32_t* argmax, const int32_t* rot_off_host,
                     int32_t n_rot, int act, const float* x, const int32_t* idx, const float* w, const float* w_eff,
                     const float* w_self, const int32_t* csr_row_ptr, const int32_t* csr_entries, int csr_flags,
                     const uint32_t* csr_stats, const uint32_t* fwd_stats, int32_t v_out, int32_t v_src, int32_t R,
                     int32_t A, int32_t F, int32_t T, int precision, int stages, float* d_x, float* d_w_eff,
                     float* d_w_self, float* d_bias, void* workspace, size_t workspace_bytes, void* stream);
/* stages: 3 = the whole backward.  1 = everything up to the weight gradients (d_w_eff, d_w_self, d_bias), 2 = the
 * signal gradient from what a stages = 1 call with the same arguments left in the SAME workspace.  Data-parallel
 * callers start the all-reduce of the weight gradients between the two calls, so that it runs beside the dSignal
 * GEMM (geoconv_b200/distributed.py).  Only where gcb_conv_amp_bwd_can_split() returns 1. */
int gcb_conv_amp_bwd_can_split(int32_t R, int32_t A, int32_t F, int32_t T, int precision, int has_neighbor);

/* ---- data-parallel weight-gradient all-reduce over NVLink peer memory ------------
 * The reference has no multi-GPU mode; this is the one exchange step of the batch-of-meshes mode.  Up to three
 * gradient tensors (d_w_eff, d_w_self, d_bias; g1 / g2 may be NULL) are summed over `world` ranks IN PLACE by one
 * kernel working on directly mapped peer buffers (two-shot: every rank reduces its slice of the flattened
 * gradients over all ranks in rank order - fixed order, bit-identical on every rank - and delivers it to all).
 * peer_bufs_dev: device array of `world` addresses, entry q = rank q's symmetric buffer of
 * gcb_peer_allreduce_bytes() bytes mapped into this process (e.g. torch.distributed._symmetric_memory), zeroed
 * once before first use; the same buffers serve every later call (the epoch lives in them, so the launch can be
 * replayed from a CUDA graph).  Every rank must make the same calls in the same order.  n_ctas (default 12): the
 * kernel is meant to run beside the dSignal GEMM on the SMs gcb_yield_sms leaves free.  A peer that never arrives
 * raises an error word after ~2 s instead of hanging: the uint32 at byte offset
 * align256(2 * n_pad * 4) + (2 * world + 4) * 4 of this rank's buffer, n_pad = n_elements rounded up to 4.  */
size_t gcb_peer_allreduce_bytes(int64_t n_elements, int32_t world);
int gcb_peer_allreduce(float* g0, int64_t n0, float* g1, int64_t n1, float* g2, int64_t n2,
                       const uint64_t* peer_bufs_dev, int32_t rank, int32_t world, int32_t n_ctas, void* stream);
/* The same without the two local copies: the caller has written its n_elements (multiple of 4) gradient values to
 * the first n_elements floats of ITS buffer (e.g. by passing pointers into it as d_w_eff / d_w_self / d_bias of
 * gcb_conv_amp_bwd) and finds the sums in the following n_elements floats when the kernel has finished.        */
int gcb_peer_allreduce_staged(int64_t n_elements, const uint64_t* peer_bufs_dev, int32_t rank, int32_t world,
                              int32_t n_ctas, void* stream);

/* ---- vertex-row sharding helper (halo exchange) ---------------------------
 * out[i,:] = src[rows[i],:]  (packs the signal rows a peer asked for into a send buffer). */
int gcb_pack_rows(const float* src, const int32_t* rows, int32_t n_rows, int32_t F, float* out,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOCONV_B200_H_ */
