/*
 * curvb200.h -- C ABI of the B200-native curvature-matvec engine.
 *
 * The reference (f-dangel/curvlinops, pure Python) has no FFI; its seam for this hot path is the
 * Python subclass contract
 *     CurvatureLinearOperator._matmat / _matmat_batch      curvlinops/_torch_base.py:923-989
 *     make_ggn_vector_product (Jv -> H_loss -> J^T)        curvlinops/ggn.py:42-72
 *     make_batch_ggn_mc_vector_product                     curvlinops/ggn.py:100-168
 *     make_batch_hessian_vector_product                    curvlinops/hessian.py:13-69
 *     HooksKFACComputer._compute_kronecker_factors         curvlinops/computers/kfac_hooks.py:176-393
 *     KroneckerProductLinearOperator._matmat / inverse     curvlinops/kronecker.py:141-171,250-373
 *     EighDecomposedLinearOperator._matmat                 curvlinops/eigh.py:84-104
 * Each entry point below replaces the per-mini-batch body of one of those; the Python operators
 * in curvlinops_b200/ keep the reference's class names, constructor arguments and error behaviour
 * and call this library through ctypes.
 *
 * Rules
 *  - plain C types only: device pointers are void* / float*, sizes are int / long long / size_t,
 *    the stream is passed as void* (a cudaStream_t);
 *  - every device buffer (parameters, X, y, V, out, factors, workspace) is allocated and owned by the
 *    caller (PyTorch); the library never allocates device memory after curv_program_create (which
 *    allocates none either -- plans are host-side);
 *  - all work is enqueued on the given stream, no host synchronisation, no hidden streams/threads;
 *  - return value 0 = success; otherwise a CURV_ERR_* code, text via curv_last_error();
 *  - results are run-to-run deterministic (no floating-point atomics; split reductions are summed in
 *    a fixed order);
 *  - there is no CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef CURVB200_H
#define CURVB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CURV_ABI_VERSION 1

enum curv_status {
  CURV_OK = 0,
  CURV_ERR_INVALID = 1,     /* bad argument / unsupported configuration -> ValueError        */
  CURV_ERR_UNSUPPORTED = 2, /* op outside the supported layer set -> NotImplementedError     */
  CURV_ERR_WORKSPACE = 3,   /* workspace too small -> RuntimeError                           */
  CURV_ERR_CUDA = 4         /* CUDA runtime error (no device, launch failure) -> RuntimeError */
};

/* node kinds of the layer program */
enum curv_op {
  CURV_OP_INPUT = 0,    /* network input X, NCHW fp32 in caller memory -> internal NHWC            */
  CURV_OP_CONV = 1,     /* Conv2d / Linear (Linear = 1x1 conv on a 1x1 map, or HxW "valid" conv
                           when it follows a flatten): aten.convolution / aten.addmm / aten.mm      */
  CURV_OP_AFFINE = 2,   /* BatchNorm2d in eval mode: y = (x-mean)/sqrt(var+eps)*gamma+beta          */
  CURV_OP_RELU = 3,
  CURV_OP_ADD = 4,      /* residual join                                                           */
  CURV_OP_MAXPOOL = 5,
  CURV_OP_AVGPOOL = 6,  /* global average pool (AdaptiveAvgPool2d(1) / mean over H,W)               */
  CURV_OP_SIGMOID = 7,
  CURV_OP_TANH = 8,
  CURV_OP_LAYERNORM = 9, /* LayerNorm over the channels of each pixel / token (aten.native_layer_norm over the last
                            dimension): p0 / c0 = weight, p1 / c1 = bias, eps.  GGN / MC / Jacobian kinds; not part
                            of Hessian (R-op) programs                                                          */
  CURV_OP_GELU = 10,     /* exact (erf) GELU                                                                     */
  CURV_OP_ATTENTION = 11, /* multi-head self-attention core softmax(Q K^T / sqrt(d)) V of nn.MultiheadAttention /
                            scaled_dot_product_attention (no mask, no dropout): in0 = packed projections
                            [B, T, 3E] (q | k | v thirds, head h at columns h*d..h*d+d of each), out = [B, T, E],
                            kh = number of heads; E a multiple of 8.  GGN / Jacobian sweeps (not the Hessian R-op) */
  /* Token glue of a vision transformer (values [B, T, C] with C a multiple of 8; not in R-op programs):           */
  CURV_OP_RESHAPE = 12,  /* same elements, other (C, H, W): [B, C, H, W] patch map read as [B, H*W, C] tokens.  The
                            output value ALIASES the input's storage (no copy)                                      */
  CURV_OP_CLSCAT = 13,   /* out[b] = cat(class_token, in[b]) along the tokens; p0 / c0 = class token [C]          */
  CURV_OP_POSADD = 14,   /* out[b] = in[b] + pos, pos [T, C] broadcast over the batch; p0 / c0 = position embedding */
  CURV_OP_TOKSEL = 15    /* out[b] = in[b, kw, :]  (read-out of one token, kw = its index); out is [B, C]           */
};

enum curv_loss { CURV_LOSS_CE = 0, CURV_LOSS_MSE = 1, CURV_LOSS_BCE = 2 };

/* which matrix the matmat entry point applies */
enum curv_kind {
  CURV_KIND_GGN = 0,     /* exact GGN:  J^T H_loss J                 (ggn.py:42-72)                 */
  CURV_KIND_GGN_MC = 1,  /* MC-Fisher:  H_loss ~ sum_m g_m g_m^T     (ggn.py:100-168)               */
  CURV_KIND_HESSIAN = 2, /* full Hessian via R-op (Pearlmutter)      (hessian.py:13-69)             */
  CURV_KIND_JVP = 3,     /* out_pred = J V   (jacobian.py:30-48)                                    */
  CURV_KIND_VJP = 4,     /* out = J^T W      (jacobian.py:77-94)                                    */
  CURV_KIND_FORWARD = 5  /* primal forward only; the prediction is then readable through
                            curv_program_value_layout (used to draw the MC samples, ggn.py:157)     */
};

/* An activation of the network for ONE sample: C channels on an HxW map ([B,C] tensors: H=W=1).
   Internally stored NHWC with C padded to a multiple of 4 (pad lanes are kept at zero). */
typedef struct curv_value_desc {
  int C, H, W;
  int has_tangent; /* 1 if the value depends on a parameter of the operator's `params`          */
} curv_value_desc;

/* One differentiated parameter tensor (an entry of the operator's `params` dict, in dict order). */
typedef struct curv_param_desc {
  long long numel;
  long long offset; /* row offset in the flat [P, K] matrices V / out                              */
} curv_param_desc;

typedef struct curv_node_desc {
  int op;            /* enum curv_op                                                              */
  int in0, in1, out; /* value ids (in1 = -1 unless CURV_OP_ADD)                                   */
  /* parameter slots: index into `params` (differentiated), or -1 (constant / absent).
     CONV: p0 = weight, p1 = bias.  AFFINE: p0 = gamma, p1 = beta.                                */
  int p0, p1;
  /* const slots: index into the per-call `const_ptrs` array for tensors that are NOT
     differentiated (-1 = absent).  CONV: c0 = weight, c1 = bias (when not in params).
     AFFINE: c0 = gamma, c1 = beta (when not in params), c2 = running_mean, c3 = running_var.     */
  int c0, c1, c2, c3;
  int kh, kw, sh, sw, ph, pw; /* CONV / MAXPOOL geometry.  AFFINE / ADD: kh == 2 fuses a following ReLU
                                 (out = relu(...)); not allowed in programs created with hessian=1     */
  float eps;                  /* AFFINE                                                            */
} curv_node_desc;

typedef struct curv_program curv_program;

/* Build the execution plan (host side only) for mini-batches of `batch` samples and up to `kmax`
   simultaneous columns.  `hessian` is a flag word: 1 reserves the extra cotangent storage the R-op needs
   (CURV_KIND_HESSIAN), 2 the scratch of curv_kfac_accumulate_batch, 4 selects bf16 arithmetic for the
   tensor-core contractions (bf16 operators, _torch_base.py:586-589: operands rounded to bf16, ONE
   tcgen05.mma.kind::f16 per product, fp32 accumulation; parameters / X / V / out stay fp32 at this boundary --
   the host mirrors keep exact fp32 copies of bf16 tensors).
   values[0] must be the network input; nodes are in topological order; the last node's `out` is the
   prediction and must have H=W=1. */
int curv_program_create(const curv_value_desc* values, int n_values, const curv_node_desc* nodes,
                        int n_nodes, const curv_param_desc* params, int n_params, int batch, int kmax,
                        int hessian, curv_program** out);
void curv_program_destroy(curv_program* prog);
size_t curv_program_workspace_bytes(const curv_program* prog);

/* Debug / test access: location of a value's activation storage inside the workspace.
   Layout: [slot][batch][H][W][Cp] fp32, slot 0 = primal, slot k = k-th tangent (later cotangent). */
int curv_program_value_layout(const curv_program* prog, int value_id, size_t* byte_offset,
                              size_t* slot_bytes, int* cp);

/* out[:, k0:k0+K] += alpha * (mini-batch matrix) @ V[:, k0:k0+K]        (_torch_base.py:937-944)
 *
 *  param_ptrs[n_params]  device pointers of the differentiated parameters (fp32, contiguous)
 *  const_ptrs[]          device pointers of constant tensors referenced by c0..c3
 *  X                     [batch, C, H, W] fp32 NCHW (or [batch, C]), contiguous
 *  y                     labels: int64 [batch] (CE) or fp32 [batch, C] (MSE/BCE)
 *  mc_grad               CURV_KIND_GGN_MC only: [batch, mc_samples, C] fp32 would-be gradients,
 *                        already divided by sqrt(mc_samples) (ggn_utils.py:369-372)
 *  V, out                [P, ldk] fp32 row-major (K minor, as the reference's [*shape, K] tensors)
 *  loss_scale            reduction constant c of the mini-batch loss Hessian (1/B for CE-mean, ...)
 *  alpha                 mini-batch weight (1 or B/N_data, _empirical_risk.py:340-352)
 *  For CURV_KIND_JVP `out` is [batch, C_out, ldk]; for CURV_KIND_VJP `V` is [batch, C_out, ldk].
 */
int curv_matmat_batch(curv_program* prog, int kind, int loss, const void* const* param_ptrs,
                      const void* const* const_ptrs, const void* X, const void* y,
                      const float* mc_grad, int mc_samples, const float* V, float* out, int K, int ldk,
                      int k0, float loss_scale, float alpha, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Streaming variant for callers whose V / result live in host memory (the reference's SciPy bridge,
 * _torch_base.py:560-592, moves the whole [P, K] matrix to the device, multiplies, and moves the result back).
 * v_ready[n_params] / out_done[n_params] hold cudaEvent_t handles (arrays or single entries may be NULL):
 *   - the columns of V belonging to parameter p are not read before v_ready[p] has completed (the node owning p
 *     is prepared lazily inside the forward sweep), so an upload in parameter order overlaps the first layers;
 *   - out_done[p] is recorded on `stream` as soon as the rows of parameter p in `out` are final (the backward
 *     sweep finishes the LAST parameters first), so their download overlaps the rest of the sweep.
 * Everything else as curv_matmat_batch. */
int curv_matmat_batch_sync(curv_program* prog, int kind, int loss, const void* const* param_ptrs,
                           const void* const* const_ptrs, const void* X, const void* y,
                           const float* mc_grad, int mc_samples, const float* V, float* out, int K, int ldk,
                           int k0, float loss_scale, float alpha, void* workspace, size_t workspace_bytes,
                           void* stream, void* const* v_ready, void* const* out_done);

/* KFAC-expand factor accumulation for one mini-batch (kfac_hooks.py:176-393):
 *   A_l += wA * sum_{n,s} a~ a~^T        a~ = im2col patch of the layer input (+1 if joint bias)
 *   G_l += wG * sum_{v,n,s} g g^T        g  = backpropagated grad_outputs[v]
 *  layer_nodes[n_layers]   node ids of the CONV nodes to collect
 *  A_ptrs / G_ptrs         fp32 device matrices [d_in(+1), d_in(+1)] / [d_out, d_out], patch order
 *                          channel-major (c, kh, kw) like F.unfold (kfac_utils.py:120-121)
 *  joint_bias[n_layers]    1: append the ones column
 *  grad_outputs            [V, batch, C] fp32 seeds (already scaled by the caller), may be NULL if V=0
 */
int curv_kfac_accumulate_batch(curv_program* prog, const void* const* param_ptrs,
                               const void* const* const_ptrs, const void* X, const int* layer_nodes,
                               int n_layers, float* const* A_ptrs, float* const* G_ptrs,
                               const int* joint_bias, const float* grad_outputs, int V, float wA,
                               float wG, void* workspace, size_t workspace_bytes, void* stream);

/* EKFAC eigenvalue correction for one mini-batch (ekfac_hooks.py:25-238), fp32 programs created with flag 2:
 *   lambda_l[i][j] += w * sum_{v, n} ( sum_s (g_{v,n,s} Qg_l)[i] * (a~_{n,s} Qa_l)[j] )^2
 * -- per-example gradients in the Kronecker eigenbasis, squared and summed, on the tensor-core kernels (rotations as
 * K-major GEMMs, the per-example contraction as a split-at-the-example-boundaries run of the weight-gradient kernel).
 *  joint_bias[l]   0: weight group, 1: joint weight + bias group, 2: bias-only group;
 *                  | 4: NO rotation (QA / QG ignored, may be NULL): lambda_l += w * sum_{v,n} (per-example gradient)^2,
 *                  the exact / MC GGN diagonal of the layer in canonical [d_out, d_in (+1)] layout
 *                  (curvlinops/computers/ggn_diagonal.py:49-109) - the contraction reads the patch planes and the
 *                  cotangent planes directly
 *  QA_ptrs[l]      [w_l, w_l] eigenvectors of A_l as columns, rows in F.unfold (c, kh, kw) order (+ joint ones row),
 *                  w_l = d_in (+1 if joint); bias-only groups: the 1 x 1 matrix [[1]], w_l = 1
 *  QG_ptrs[l]      [d_out, d_out] eigenvectors of G_l
 *  lambda_ptrs[l]  [d_out, w_l] fp32, accumulated
 *  grad_outputs    [V, batch, C] seeds (already scaled by the caller)
 * Entries with a NULL lambda pointer are skipped. */
int curv_ekfac_correction_batch(curv_program* prog, const void* const* param_ptrs, const void* const* const_ptrs,
                                const void* X, const int* layer_nodes, int n_layers, const float* const* QA_ptrs,
                                const float* const* QG_ptrs, float* const* lambda_ptrs, const int* joint_bias,
                                const float* grad_outputs, int V, float w, void* workspace, size_t workspace_bytes,
                                void* stream);

/* Y[d_out, d_in, K] = G[d_out,d_out] . X[d_out, d_in, K] . A[d_in,d_in]^T  per column
   (kronecker.py:141-153 'abZ,Aa,Bb->ABZ').  tmp needs d_out*d_in*K floats.  A or G may be NULL
   (identity).  If lambda != NULL this is the eigen-basis apply of eigh.py:98-104:
   Y = Qg ( (Qg^T X Qa) * scale(lambda) ) Qa^T with scale = lambda or 1/(lambda+damping). */
int curv_kron_apply(const float* G, const float* A, int d_out, int d_in, int K, const float* X,
                    float* Y, float* tmp, void* stream);
int curv_eigh_apply(const float* Qg, const float* Qa, const float* lambda, float damping, int inverse,
                    int d_out, int d_in, int K, const float* X, float* Y, float* tmp, float* tmp2,
                    void* stream);

/* The same two-sided product on the tcgen05 contraction kernels (csrc/kron_tc.cuh), for the block sizes of real
   networks (KFAC and its damped inverse, kronecker.py:250-373; with eigenvector matrices as factors the rotations of
   eigh.py:98-104):  Y[d_out, d_in, K] = G . X . A^T for any square G [d_out,d_out], A [d_in,d_in], 1 <= K <= 8, as two
   K-major GEMMs (T_z = G X_z with X_z^T as weight images; Y_z = T_z A^T with A as weight image), fp32-grade (operands
   as fp16 hi/lo planes, three MMAs per product, chunked tensor-memory accumulation).
   factor_ws: caller-owned buffer of curv_kron_apply_tc_factor_bytes(d_out, d_in) bytes holding the operand forms of the
   two factors; the call fills it unless factors_ready != 0 (build once per operator, reuse for every product).
   ws: caller-owned scratch of curv_kron_apply_tc_workspace(d_out, d_in, K) bytes. */
size_t curv_kron_apply_tc_workspace(int d_out, int d_in, int K);
size_t curv_kron_apply_tc_factor_bytes(int d_out, int d_in);
int curv_kron_apply_tc(const float* G, const float* A, int d_out, int d_in, int K, const float* X, float* Y,
                       void* factor_ws, size_t factor_bytes, int factors_ready, void* ws, size_t ws_bytes,
                       void* stream);

/* C[M,N] = alpha * op(A) op(B) + beta * C, fp32 row-major, hand-written kernels (no cuBLAS).  */
int curv_gemm(int transA, int transB, int M, int N, int Kd, float alpha, const float* A, int lda,
              const float* B, int ldb, float beta, float* C, int ldc, void* stream);

/* `batch` independent products with element strides sA / sB / sC between consecutive operands (one launch). */
int curv_gemm_batched(int transA, int transB, int M, int N, int Kd, float alpha, const float* A, int lda,
                      long long sA, const float* B, int ldb, long long sB, float beta, float* C, int ldc,
                      long long sC, int batch, void* stream);

/* Full re-orthogonalisation of a Lanczos vector (the eigensolver consumer of the products, BASELINE.json configs[4];
   the reference leaves this to ARPACK on the host behind PyTorchLinearOperator.to_scipy, _torch_base.py:560-592):
   `rounds` times  c = Q[:m] w ;  w -= Q[:m]^T c  with Q [m, n] fp32 rows of leading dimension ldq, w [n] fp32, all on
   the device.  coeff (may be NULL): [m] floats receiving the coefficients summed over the rounds (coeff[m-1] is the
   Lanczos alpha when w = A q_{m-1}).  Two streaming passes over Q[:m] per round, fixed-order sums (repeatable).
   ws: curv_lanczos_reorth_workspace(m, n) bytes. */
long long curv_lanczos_reorth_workspace(int m, long long n);
int curv_lanczos_reorth(const float* Q, long long ldq, int m, float* w, long long n, int rounds, float* coeff,
                        void* ws, long long ws_bytes, void* stream);

const char* curv_last_error(void);
int curv_abi_version(void);
/* number of kernel launches issued by this library since process start (bench `gpu_launches`) */
long long curv_launch_count(void);
/* replaying a captured CUDA graph that contains n launches of this library: keeps the count truthful */
void curv_add_launch_count(long long n);
/* Per-launch CUDA-event timing of the contraction kernels (class 0: gather GEMM = forward/dgrad,
   class 1: wgrad GEMM).  enable(1) clears and starts recording, read() synchronises on the recorded
   events and returns summed milliseconds, algorithmic FLOPs and launch counts per class. */
int curv_profile_enable(int on);
int curv_profile_read(double* ms, double* flops, long long* count);
/* one class at a time; class 2 = the absmax / fp16 hi-lo split passes feeding the half-split kernels */
int curv_profile_read_class(int cls, double* ms, double* flops, long long* count);
/* 0: SIMT fp32 contraction kernels only, 1 (default): tcgen05 tensor-core kernels (fp16 hi/lo "half-split"
   where the shape allows it, 3xTF32 split for the Hessian R-op and the KFAC Gram matrices) for layers with at least
   256 GEMM rows, 2: tcgen05 for every contraction (tests).  Bits 4.. are A/B switches: 0x10 no tcgen05 gather GEMM,
   0x20 no tcgen05 wgrad GEMM, 0x200 no half-split kernels (3xTF32 instead), 0x400 BatchNorm kernels never write
   operand planes directly (separate split passes), 0x800 no N-stacked kernel for shared-activation layers,
   0x4000 KFAC Gram matrices on the SIMT / 3xTF32 kernels instead of the tensor-core Gram path,
   0x2000 tangent-weight images through a packed fp32 copy of the columns of V.
   Returns the old mode. */
int curv_set_tensor_core_mode(int mode);
/* Everything global that changes which kernels a call launches: the mode word above | (profiling enabled) << 16.
   Host mirrors that replay captured CUDA graphs of curv_matmat_batch key their caches on it. */
int curv_launch_config(void);

#ifdef __cplusplus
}
#endif
#endif /* CURVB200_H */
