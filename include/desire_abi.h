/* desire_abi.h — C-ABI of libdesire_b200.so, the sm_100a kernel library behind the DESIRE hot path.
 *
 * The reference (tdavchev/DESIRE) has no FFI: its hot path is TensorFlow-1 graph code inside
 * model/model.py, entered through sess.run (train.py:181).  Each entry point below replaces the
 * TF/prettytensor ops of one stage of that graph (cited per function) — the binding a maintainer
 * would add is the ctypes stub in INTEGRATION.md (desire_b200/_lib.py is that stub, in use).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to row-major contiguous float32 unless a ld_/stride
 *     argument says otherwise; images are NHWC; nothing is allocated inside the library:
 *     scratch comes from the caller through (ws, ws_bytes) and *_workspace_bytes() queries;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - returns 0 on success, <0 on error (desire_last_error() gives the text, thread-local);
 *   - rows: M = B*N agents (row = b*N+n), R = M*K agent-samples (row = m*K+k);
 *   - trajectories keep the reference's [agents, time, (id,x,y)] layout (model/model.py:91-105).
 *   - GRU weights follow TF-1.x GRUCell: wg [(I+H), 2H] (r|u), bg [2H], wc [(I+H), H], bc [H];
 *     rows 0..I-1 multiply the input, rows I.. multiply the state.
 */
#ifndef DESIRE_ABI_H_
#define DESIRE_ABI_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DESIRE_ABI_VERSION 1

#define DESIRE_OK 0
#define DESIRE_ERR_INVALID (-1)     /* bad argument / unsupported size            */
#define DESIRE_ERR_CUDA (-2)        /* a CUDA runtime call or launch failed       */
#define DESIRE_ERR_WORKSPACE (-3)   /* ws_bytes smaller than *_workspace_bytes()  */

#define DESIRE_ACT_NONE 0
#define DESIRE_ACT_RELU 1
#define DESIRE_ACT_ELU 2
#define DESIRE_ACT_SIGMOID 3

typedef void* desire_stream_t;

/* GRU parameters (TF-1.x GRUCell layout, see header comment). */
typedef struct {
  const float* wg;
  const float* bg;
  const float* wc;
  const float* bc;
} desire_gru_t;

/* One conv / deconv layer of the CVAE: kernel, bias, BN gamma, BN beta. */
typedef struct {
  const float* w;
  const float* b;
  const float* gamma;
  const float* beta;
} desire_convbn_t;

/* vae_encoder, model/model.py:471-492: conv5/2x32, conv5/2x64, conv5 VALIDx128 [kh,kw,in,out], fc 2048->2Z */
typedef struct {
  desire_convbn_t c1, c2, c3;
  const float* fc_w;
  const float* fc_b;
} desire_cvae_enc_t;

/* vae_decoder, model/model.py:453-469: deconv4 VALIDx128, deconv5 VALIDx64, deconv5/2x32, deconv5/2x1
 * filters [kh,kw,out,in] (utils/convolutional_vae_util.py:83) */
typedef struct {
  desire_convbn_t d1, d2, d3, d4;
} desire_cvae_dec_t;

/* stage-2 parameters (absent in the reference, marker model/model.py:312-313) */
typedef struct {
  const float *c1_w, *c1_b, *c2_w, *c2_b, *c3_w, *c3_b; /* scene CNN 5x5: 3->16 s2, 16->32, 32->Cs */
} desire_scene_cnn_t;

typedef struct {
  const float *vel_w, *vel_b;     /* [2,Fv], [Fv]        */
  const float *sp_w, *sp_b;       /* [G*H, H], [H]       */
  desire_gru_t dec2;              /* I = Fv+Cs+2C+H      */
  const float *score_w, *score_b; /* [H], [1]            */
  const float *reg_w, *reg_b;     /* [H, 2*Tf], [2*Tf]   */
  const float* r2_edges;          /* [n_rad+1] squared radial bin edges */
  const float* dirs;              /* [n_ang,2] sector boundary directions (cos,sin) */
} desire_ioc_t;

typedef struct {
  int B, N, K, H, Tf;
  int C;            /* channel multiplier (feature_pooling width is 2C) */
  int Fv, Cs;       /* velocity-fc width, scene channels */
  int n_rad, n_ang; /* log-polar grid */
  int Hm, Wm;       /* scene feature-map size */
  int iters;
} desire_ioc_dims_t;

int desire_version(void);
const char* desire_last_error(void);

/* ---- measurement hooks (bench.py): number of kernels launched by this library so far, and optional
 * per-kernel CUDA-event timing of the tagged launches below. */
long desire_launch_count(void);
/* Hardware self-test: D = A[128,32] @ B[64,32]^T (BF16 inputs) once with the A operand in shared memory (out_ss) and once
 * with A in tensor memory (out_ts; tcgen05.mma with a TMEM A operand).  order = 0: lower half of a 32-bit TMEM column =
 * the smaller k (what the fused social kernel assumes), 1: swapped.  out_* [128,64]. */
int desire_selftest_tsmma(const float* A, const float* B, float* out_ss, float* out_ts, int order, desire_stream_t stream);
/* Timing probe: `grid` CTAs each issue `iters` back-to-back tcgen05.mma (M=128, N, K=16, BF16; mode 0 = A and B in shared
 * memory, 1 = A in tensor memory); out_cycles[0] = SM cycles block 0 needed (device pointer). */
int desire_selftest_mma_rate(int mode, int N, int iters, int grid, long long* out_cycles, desire_stream_t stream);
/* Launches that did NOT take the tensor-core / fused kernel because the shape is outside what it covers (counted
 * per kind since load; DESIRE_LOG_FALLBACK=1 prints the first of each kind to stderr).  Nothing is computed on the
 * host in any case — these are the FP32 CUDA-core / materialising forms of the same entry points. */
#define DESIRE_FALLBACK_GRU_FP32 0      /* GRU recurrence on FP32 CUDA cores (H % 32 != 0, fewer than 64 rows, ...) */
#define DESIRE_FALLBACK_GRU_V2 1        /* tcgen05 recurrence of the second design (H not in {128, 256}, ...) */
#define DESIRE_FALLBACK_SOCIAL_POOL 2   /* materialised [R, G*H] social pooling + GEMM instead of the fused kernel */
#define DESIRE_FALLBACK_GEMM_FP32 3     /* FP32 CUDA-core GEMM (no packing scratch, tiny shapes, gemm mode 0) */
#define DESIRE_FALLBACK_KINDS 4
long desire_fallback_count(int kind);
#define DESIRE_PROF_GRU_DEC1 0      /* Decoder-1 recurrence (a10)                       */
#define DESIRE_PROF_GRU_DEC2 1      /* one Decoder-2 step (a14)                         */
#define DESIRE_PROF_GRU_ENC 2       /* encoder recurrences (a3/a4)                      */
#define DESIRE_PROF_SOCIAL_POOL 3   /* log-polar social pooling, one step               */
#define DESIRE_PROF_SOCIAL_FC 4     /* pooled[R,G*H] @ sp_w GEMM, one step              */
#define DESIRE_PROF_GATHER 5        /* bilinear scene gather, all steps of an iteration */
#define DESIRE_PROF_DECONV2 6       /* CVAE decoder 4x4x128->8x8x64 GEMM, one chunk     */
#define DESIRE_PROF_DECONV3 7       /* CVAE decoder 8x8x64->16x16x32 GEMM, one chunk    */
#define DESIRE_PROF_COL2IM 8        /* col2im+BN+act kernels of the decoder             */
#define DESIRE_PROF_DEC2_XPROJ 9    /* hoisted Decoder-2 input projection GEMMs         */
#define DESIRE_PROF_SCENE_CNN 10    /* scene CNN convs                                  */
#define DESIRE_PROF_READOUT 11      /* read-out + feature pooling (a11)                 */
#define DESIRE_PROF_OTHER 12
#define DESIRE_PROF_SLOTS 16
/* GEMM arithmetic of the dense layers: 3 = tcgen05 3xBF16 split with FP32 accumulation (default; meets the
 * 1e-4 parity bar), 1 = tcgen05 single BF16 pass (fast mode), 0 = FP32 CUDA cores. */
int desire_set_gemm_mode(int mode);
int desire_get_gemm_mode(void);
int desire_prof_enable(int on);                                  /* resets the slots */
int desire_prof_read(int slot, long* launches, double* total_ms); /* synchronises the recorded events */

/* ---- generic dense layer: C = act(A[M,K] @ W[K,N] + bias), replaces tf.nn.xw_plus_b / tf.matmul
 * call sites model/model.py:249-251 (fc_c) and :272-275.  accumulate!=0 adds into C. */
int desire_fc_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                  int M, int N, int K, int act, int accumulate, desire_stream_t stream);

/* same contract on the tcgen05 path (3xBF16 or BF16 per desire_set_gemm_mode); trans_w: W stored [N,K].
 * ws holds the packed BF16 image of W (desire_gemm_tc_workspace_bytes). */
size_t desire_gemm_tc_workspace_bytes(int N, int K);
int desire_gemm_tc_fwd(const float* A, int lda, const float* W, int ldw, int trans_w, const float* bias, float* C,
                       int ldc, int M, int N, int K, int act, int accumulate, void* ws, size_t ws_bytes,
                       desire_stream_t stream);

/* ---- a2  rho_i = relu(depthwise_conv2d(VALID) + b), model/model.py:116-133.
 * obs [M,Tp,3] (id,x,y); w [Tp,2,C]; b [2C]; rho [M,2C] */
int desire_tconv_fwd(const float* obs, int M, int Tp, int C, const float* w, const float* b, float* rho,
                     desire_stream_t stream);

/* ---- a3/a4  static_rnn(GRUCell) from a zero state over the (x,y) columns of traj [M,T,3],
 * model/model.py:233-241.  h_out row stride ld_out (so H_x|H_y can share one [M,2H] buffer). */
int desire_gru_encode_fwd(const float* traj, int M, int T, int H, const desire_gru_t* w, float* h_out,
                          int ld_out, desire_stream_t stream);

/* same result through the tcgen05 recurrence: the width-2 input projection is hoisted for all T steps into ws
 * (desire_gru_encode_workspace_bytes), then the persistent tensor-core GRU of a10 runs them. */
size_t desire_gru_encode_workspace_bytes(int M, int T, int H);
int desire_gru_encode_ws_fwd(const float* traj, int M, int T, int H, const desire_gru_t* w, float* h_out,
                             int ld_out, void* ws, size_t ws_bytes, desire_stream_t stream);

/* ---- a6  vae_encoder, model/model.py:471-492.  v [M,1024] -> mu_logvar [M,2Z] (mean | logvar). */
size_t desire_cvae_encode_workspace_bytes(int M, int Z);
int desire_cvae_encode_fwd(const float* v, int M, int Z, const desire_cvae_enc_t* w, float* mu_logvar,
                           void* ws, size_t ws_bytes, desire_stream_t stream);

/* ---- a7  z = mean + sqrt(exp(logvar))*eps, model/model.py:260-264.  eps [M,K,Z] -> z [M*K,Z]. */
int desire_reparam_fwd(const float* mu_logvar, const float* eps, int M, int K, int Z, float* z,
                       desire_stream_t stream);

/* ---- a8  vae_decoder + deconv2d, model/model.py:453-469, utils/convolutional_vae_util.py:31-135.
 * z [R,Z] -> xr [R,1024].  Rows are processed in chunks so the scratch stays bounded. */
size_t desire_cvae_decode_workspace_bytes(int R, int Z);
int desire_cvae_decode_fwd(const float* z, int R, int Z, const desire_cvae_dec_t* w, float* xr, void* ws,
                           size_t ws_bytes, desire_stream_t stream);

/* ---- a9  x_z = softmax(relu(xr@W2+b2)) * H_x, model/model.py:271-280.
 * xr [R,S2]; w [S2,H]; Hx row m = r/K with row stride ld_hx; x_z [R,H]; ws holds R*H floats. */
size_t desire_mask_softmax_workspace_bytes(int R, int H);
int desire_mask_softmax_fwd(const float* xr, int R, int S2, int H, int K, const float* w, const float* b,
                            const float* Hx, int ld_hx, float* x_z, void* ws, size_t ws_bytes,
                            desire_stream_t stream);

/* ---- a10  Decoder-1: rnn_decoder with a constant input, model/model.py:279-285.
 * x_z [R,H]; h0 row = r/K of Hx (stride ld_hx); hs [R,T,H] (output_states). */
size_t desire_gru_decode_workspace_bytes(int R, int H);
int desire_gru_decode_fwd(const float* x_z, const float* Hx, int ld_hx, int R, int K, int H, int T,
                          const desire_gru_t* w, float* hs, void* ws, size_t ws_bytes,
                          desire_stream_t stream);

/* ---- a11  read-out + feature pooling, model/model.py:286-311.
 * mode 0 (D3): Yhat[r,t,:] = hs[r,t,:]@out_w + out_b + obs[r/K, Tp-1, 1:3]
 * mode 1 (reference split read-out): T2 = n_chunks, Yhat[r,t,c,:] = hs[r,t, c*(H/n_chunks) + {0,1}]
 * fpool [R, T(*n_chunks), 2C] = [y_x*rho[:C], y_y*rho[C:]]  (may be NULL). */
int desire_readout_pool_fwd(const float* hs, int R, int K, int T, int H, int mode, int n_chunks,
                            const float* out_w, const float* out_b, const float* obs, int Tp,
                            const float* rho, int C, float* Yhat, float* fpool, desire_stream_t stream);

/* ---- a12/a13  losses: kld rows (model/model.py:587-589), reconstruction rows (D7) and the masked
 * mean cost = sum_{id!=0}(rows)/count (model/model.py:351-376).  cost[0] = cost, cost[1] = count. */
int desire_kld_rows_fwd(const float* mu_logvar, int M, int Z, float* kld_rows, desire_stream_t stream);
int desire_recon_rows_fwd(const float* Yhat, const float* target, int M, int K, int T, float* recon_rows,
                          desire_stream_t stream);
int desire_masked_cost_fwd(const float* rows_a, const float* rows_b, const float* obs, int M, int Tp,
                           float* cost, desire_stream_t stream);
/* a7 noise: out[0..n) ~ N(0,1), Philox4x32-10 + Box-Muller; `state` = DEVICE pointer to {seed, offset} (two u64), read
 * at execution time so a captured CUDA graph draws fresh noise when the host bumps `offset` between replays.  Replaces
 * the in-graph tf.random_normal of model/model.py:262.  Element e depends on (seed, offset, e) only. */
int desire_randn_fwd(const unsigned long long* state, float* out, size_t n, desire_stream_t stream);
/* D8 existence (model/model.py:206,214,351-366: an object contributes only if neither obj_id nor target_obj_id is
 * the non-existent id 0).  Every entry point that takes `obs` treats "id at observed frame 0 != 0" as "exists";
 * this writes obs_out [M,Tp,3] = obs with that id zeroed for agents that are absent
 *   mode 0: at observed frame 0 only (identity copy);
 *   mode 1: at observed frame 0, at the last observed frame, or at any of the Tf target frames (target [M,Tf,3]).
 * Pass obs_out instead of obs to the rest of the path. */
int desire_existence_fwd(const float* obs, const float* target, int M, int Tp, int Tf, int mode, float* obs_out,
                         desire_stream_t stream);

/* ---- a14  stage 2 pieces (D11) */
size_t desire_scene_cnn_workspace_bytes(int B, int Hi, int Wi);
int desire_scene_cnn_fwd(const float* img, int B, int Hi, int Wi, int Cs, const desire_scene_cnn_t* w,
                         float* fmap, void* ws, size_t ws_bytes, desire_stream_t stream);
/* bilinear gather: fmap [B,Hm,Wm,Cs]; pos = base + row*pos_stride (x,y), rows_per_scene rows per b;
 * out row stride ld_out. */
int desire_scene_gather_fwd(const float* fmap, int B, int Hm, int Wm, int Cs, const float* pos,
                            long pos_stride, int rows_per_scene, float* out, int ld_out,
                            desire_stream_t stream);
/* log-polar social pooling: pos/h rows ordered (b,n,k); mask from obs ids ([B*N,Tp,3], id!=0);
 * pooled [R, G*H]. */
int desire_social_pool_fwd(const float* pos, long pos_stride, const float* h, int ld_h, const float* obs,
                           int Tp, int B, int N, int K, int H, int n_rad, int n_ang,
                           const float* r2_edges, const float* dirs, float* pooled,
                           desire_stream_t stream);
/* One step of the social feature, fused (binning + pooling + fc on tensor cores; the [R, G*H] pooled tensor is never
 * written): fsp [R,H] = relu(pool(h) @ sp_w [G*H,H] + sp_b).  ws: scratch for the packed weights.  Returns
 * DESIRE_ERR_INVALID when the shape is outside the fused kernels (N <= 128 with H in {64, 128}; 129..256 agents with
 * H in {128, 256}) — no fallback here. */
size_t desire_social_fc_workspace_bytes(int H, int n_bins);
int desire_social_fc_fwd(const float* pos, long pos_stride, const float* h, int ld_h, const float* obs, int Tp, int B,
                         int N, int K, int H, int n_rad, int n_ang, const float* r2_edges, const float* dirs,
                         const float* sp_w, const float* sp_b, float* fsp, void* ws, size_t ws_bytes,
                         desire_stream_t stream);
/* full ranking & refinement loop.  Y [R,Tf,2] is refined IN PLACE; scores [iters, R]. */
size_t desire_ioc_workspace_bytes(const desire_ioc_dims_t* d);
int desire_ioc_fwd(const desire_ioc_dims_t* d, const desire_ioc_t* w, const float* fmap, const float* obs,
                   int Tp, const float* Hx, int ld_hx, const float* fpool, float* Y, float* scores,
                   void* ws, size_t ws_bytes, desire_stream_t stream);
/* Same, given rho_i [B*N,2C] (a2) and the stage-1 trajectories Yhat [R,T,2] that feature_pooling was built from
 * (model/model.py:291-311: feature_pooling[r,t] = [yhat_x * rho_i[m,:C] | yhat_y * rho_i[m,C:]]): the feature_pooling
 * columns of the Decoder-2 input projection then collapse to two per-agent vectors scaled by (yhat_x, yhat_y), computed
 * once per call; the [R*T, 2C] tensor is not read at all.  Same results up to FP32 summation order. */
int desire_ioc_factored_fwd(const desire_ioc_dims_t* d, const desire_ioc_t* w, const float* fmap, const float* obs,
                            int Tp, const float* Hx, int ld_hx, const float* fpool, const float* rho_i,
                            const float* Yhat, float* Y, float* scores, void* ws, size_t ws_bytes,
                            desire_stream_t stream);

/* ======================================================================================================
 * Train step (SURVEY 8.0 D9, 8.b "*_bwd twins"): gradients of `cost` (model/model.py:374-376) with respect
 * to every trainable variable — what tf.gradients(self.cost, tvars) at model/model.py:388-391 would return —
 * then clip_by_global_norm (:391) and Adam (:394).  Conventions:
 *   - d<name> arguments mirror the forward tensors; PARAMETER gradients are ACCUMULATED (+=, atomics) so the
 *     caller zeroes its flat gradient buffer once per step; activation gradients are overwritten unless the
 *     comment says "+=";
 *   - every backward recomputes the forward intermediates it needs (pre-BN activations, gates) into the
 *     workspace; nothing is cached between the forward and backward calls;
 *   - conv/deconv biases that sit in front of a batch-norm have an identically zero gradient (BN removes
 *     any per-channel constant), so those bias gradients are left untouched (= 0).
 */
typedef struct { float *wg, *bg, *wc, *bc; } desire_gru_grad_t;
typedef struct { float *w, *b, *gamma, *beta; } desire_convbn_grad_t;
typedef struct { desire_convbn_grad_t c1, c2, c3; float *fc_w, *fc_b; } desire_cvae_enc_grad_t;
typedef struct { desire_convbn_grad_t d1, d2, d3, d4; } desire_cvae_dec_grad_t;

/* a12/a13/D7 backward.  count = device pointer to the (global) number of existing agents (cost[1] of
 * desire_masked_cost_fwd, all-reduced by the caller when scenes are sharded over ranks).
 * dYhat [R,T,2] = d cost / d Yhat;  d_mu_logvar [M,2Z] = KLD part (overwritten). */
int desire_cost_bwd(const float* Yhat, const float* target, const float* mu_logvar, const float* obs,
                    const float* count, int M, int K, int T, int Tp, int Z, float* dYhat, float* d_mu_logvar,
                    desire_stream_t stream);
/* a11 (linear read-out) backward: dhs [R,T,H] = dYhat @ out_w^T (overwritten); d_out_w [H,2], d_out_b [2] += */
int desire_readout_bwd(const float* hs, const float* dYhat, int R, int T, int H, const float* out_w, float* dhs,
                       float* d_out_w, float* d_out_b, desire_stream_t stream);
/* a10 backward through time.  dhs [R,T,H]: in = gradient reaching every state from the read-out, clobbered.
 * dx_z [R,H] overwritten; dHx (row m, stride ld_dhx) += sum_k d h0. */
size_t desire_gru_decode_bwd_workspace_bytes(int R, int H, int T);
int desire_gru_decode_bwd(const float* x_z, const float* Hx, int ld_hx, int R, int K, int H, int T,
                          const desire_gru_t* w, const float* hs, float* dhs, float* dx_z, float* dHx,
                          int ld_dhx, const desire_gru_grad_t* g, void* ws, size_t ws_bytes,
                          desire_stream_t stream);
/* a9 backward: dxr [R,S2] overwritten; dHx += ; dw [S2,H], db [H] += */
size_t desire_mask_softmax_bwd_workspace_bytes(int R, int H);
int desire_mask_softmax_bwd(const float* xr, int R, int S2, int H, int K, const float* w, const float* b,
                            const float* Hx, int ld_hx, const float* dx_z, float* dxr, float* dHx, int ld_dhx,
                            float* dw, float* db, void* ws, size_t ws_bytes, desire_stream_t stream);
/* a8 backward: dz [R,Z] overwritten */
size_t desire_cvae_decode_bwd_workspace_bytes(int R, int Z);
int desire_cvae_decode_bwd(const float* z, int R, int Z, const desire_cvae_dec_t* w, const float* dxr, float* dz,
                           const desire_cvae_dec_grad_t* g, void* ws, size_t ws_bytes, desire_stream_t stream);
/* a7 backward: d_mu_logvar [M,2Z] += (sum_k dz, sum_k dz*eps*0.5*sqrt(exp(logvar))) */
int desire_reparam_bwd(const float* mu_logvar, const float* eps, const float* dz, int M, int K, int Z,
                       float* d_mu_logvar, desire_stream_t stream);
/* a6 backward: dv [M,1024] overwritten */
size_t desire_cvae_encode_bwd_workspace_bytes(int M, int Z);
int desire_cvae_encode_bwd(const float* v, int M, int Z, const desire_cvae_enc_t* w, const float* d_mu_logvar,
                           float* dv, const desire_cvae_enc_grad_t* g, void* ws, size_t ws_bytes,
                           desire_stream_t stream);
/* dense layer backward (a5 fc_c and friends): Cout = act(A@W+b) as produced by desire_fc_fwd; dC is
 * clobbered (becomes the pre-activation gradient).  dA (ldda) is overwritten, or += when accumulate_dA;
 * dA may be NULL.  dW [K,N] (lddw), db [N] +=. */
int desire_fc_bwd(const float* A, int lda, const float* W, int ldw, const float* Cout, int ldc, float* dC,
                  int lddc, int M, int N, int K, int act, float* dA, int ldda, int accumulate_dA, float* dW,
                  int lddw, float* db, desire_stream_t stream);
/* a3/a4 backward through time from the gradient of the final state dh (row stride ld_dh). */
size_t desire_gru_encode_bwd_workspace_bytes(int M, int T, int H);
int desire_gru_encode_bwd(const float* traj, int M, int T, int H, const desire_gru_t* w, const float* dh,
                          int ld_dh, const desire_gru_grad_t* g, void* ws, size_t ws_bytes,
                          desire_stream_t stream);
/* clip_by_global_norm + Adam on flat buffers (model/model.py:391-394; TF-1.x AdamOptimizer update rule):
 *   scale = clip / max(||g||, clip);  m = b1 m + (1-b1) g s;  v = b2 v + (1-b2) (g s)^2;
 *   p -= lr * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps).
 * desire_sumsq_fwd: out[0] (+)= sum g^2 (out is zeroed first unless accumulate). grad_scale multiplies g
 * before everything else (1/world for an averaged all-reduce, 1 otherwise).  clip <= 0 disables clipping. */
int desire_sumsq_fwd(const float* g, long n, float* out, int accumulate, desire_stream_t stream);
int desire_adam_step(float* p, const float* g, float* m, float* v, long n, const float* sumsq, float lr,
                     float beta1, float beta2, float eps, int step, float clip, float grad_scale,
                     desire_stream_t stream);

/* ---- a14 train step (stage 2 is absent in the reference, so is its loss).  Per iteration it:
 *   CE_it = -sum_k q log p, p = softmax_k(score_it), q = softmax_k(-max_t ||Y - Y_it(k)||) (constant);
 *   REG_it = mean_k sum_t ||Y - Y_{it+1}(k)||^2;  ioc_cost = sum_{existing}(sum_it CE+REG) / count.
 * Stage-wise: Yhat, feature_pooling, H_x are constants of this module and features are taken at
 * stop_gradient(Y_it).  desire_ioc_train runs the forward of all iterations (Y [R,Tf,2] = refined output, scores
 * [iters,R] as desire_ioc_fwd), writes ioc_cost[0] (normalised by *count) and ioc_cost[1] (local agent count), and
 * accumulates the gradients of ioc_cost into g (+=) and into dfmap [B,Hm,Wm,Cs] (+=, zeroed by the caller), which
 * desire_scene_cnn_bwd then carries into the scene CNN weights (dfmap is clobbered). */
typedef struct {
  float *vel_w, *vel_b, *sp_w, *sp_b;
  desire_gru_grad_t dec2;
  float *score_w, *score_b, *reg_w, *reg_b;
} desire_ioc_grad_t;
typedef struct { float *c1_w, *c1_b, *c2_w, *c2_b, *c3_w, *c3_b; } desire_scene_cnn_grad_t;
size_t desire_ioc_train_workspace_bytes(const desire_ioc_dims_t* d);
int desire_ioc_train(const desire_ioc_dims_t* d, const desire_ioc_t* w, const float* fmap, const float* obs,
                     int Tp, const float* target, const float* Hx, int ld_hx, const float* fpool,
                     const float* Yhat, const float* count, float* Y, float* scores, float* ioc_cost,
                     const desire_ioc_grad_t* g, float* dfmap, void* ws, size_t ws_bytes, desire_stream_t stream);
size_t desire_scene_cnn_bwd_workspace_bytes(int B, int Hi, int Wi);
int desire_scene_cnn_bwd(const float* img, int B, int Hi, int Wi, int Cs, const desire_scene_cnn_t* w,
                         float* dfmap, const desire_scene_cnn_grad_t* g, void* ws, size_t ws_bytes,
                         desire_stream_t stream);

/* generic weight-gradient product, the backward twin of desire_fc_fwd's W: dW[K,N] (lddw) += A[M,K]^T @ dC[M,N].
 * With ws >= desire_wgrad_workspace_bytes(M,N) (the packed BF16 image of dC) and K >= 64, N >= 16, M >= 2048 it runs
 * on tcgen05 (3xBF16, split over the rows, atomics); otherwise on FP32 CUDA cores (ws may be NULL). */
size_t desire_wgrad_workspace_bytes(int M, int N);
int desire_wgrad_tn(const float* A, int lda, const float* dC, int lddc, float* dW, int lddw, int M, int N, int K,
                    void* ws, size_t ws_bytes, desire_stream_t stream);

/* generic transposed convolution, the operator of utils/convolutional_vae_util.py:31-135 (deconv2d: conv2d_transpose ->
 * +bias -> batch-normalise -> activation) for any square geometry; the CVAE decoder's own four layers run through the
 * fused kernels of desire_cvae_decode_fwd.  x [R,Hin,Hin,Cin] NHWC; w [k,k,Cout,Cin] (:83); same != 0: SAME (output
 * Hin*stride, :165-167) else VALID ((Hin-1)*stride+k, :161-163); bias / gamma+beta may be NULL (no bias / no BN; BN is the
 * per-row statistics of DESIGN.md D5 and needs Cout dividing 256); y [R,Hout,Hout,Cout]. */
size_t desire_deconv2d_workspace_bytes(int R, int Hin, int Cin, int k, int stride, int same, int Cout);
int desire_deconv2d_fwd(const float* x, int R, int Hin, int Cin, const float* w, int k, int stride, int same, int Cout,
                        const float* bias, const float* gamma, const float* beta, int act, float* y, void* ws,
                        size_t ws_bytes, desire_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DESIRE_ABI_H_ */
