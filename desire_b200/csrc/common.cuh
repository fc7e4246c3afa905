// Shared helpers for libdesire_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <functional>

#include "../../include/desire_abi.h"

struct CUtensorMap_st;   // cuda.h (CUtensorMap), forward-declared so this header needs no driver API

namespace desire {

void set_error(const char* fmt, ...);
void count_launch();
void note_fallback(int kind, const char* what, int a = 0, int b = 0, int c = 0);

// Optional per-kernel timing (desire_prof_enable): a ProfScope brackets ONE launch with CUDA events on
// the launching stream; desire_prof_read sums them per slot.  Off by default (zero overhead but a branch).
int prof_set_slot(int slot);
void prof_kernel_begin(cudaStream_t st);
void prof_kernel_end(cudaStream_t st);
struct ProfScope {   // names the slot that DESIRE_LAUNCH-bracketed kernels inside it are charged to
  int old;
  ProfScope(int s, cudaStream_t) : old(prof_set_slot(s)) {}
  ~ProfScope() { prof_set_slot(old); }
};

#define DESIRE_CHECK_ARG(cond, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      ::desire::set_error(__VA_ARGS__);        \
      return DESIRE_ERR_INVALID;               \
    }                                          \
  } while (0)

#define DESIRE_CUDA(call)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      ::desire::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(e_));                                     \
      return DESIRE_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

// every kernel launch of the library goes through this macro, so the counter is the number of OUR kernels
#define DESIRE_LAUNCH_CHECK()                 \
  do {                                        \
    ::desire::count_launch();                 \
    DESIRE_CUDA(cudaGetLastError());          \
  } while (0)

// launch + (optional) tight event bracket + launch counter/error check
#define DESIRE_LAUNCH(st, ...)            \
  do {                                    \
    ::desire::prof_kernel_begin(st);      \
    __VA_ARGS__;                          \
    ::desire::prof_kernel_end(st);        \
    DESIRE_LAUNCH_CHECK();                \
  } while (0)

// raise a kernel's dynamic shared-memory limit once per process (cached per call site)
#define DESIRE_ENSURE_SMEM(kernel, bytes)                                                              \
  do {                                                                                                 \
    static size_t cached_ = 0;                                                                         \
    if ((size_t)(bytes) > cached_) {                                                                   \
      DESIRE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      cached_ = (size_t)(bytes);                                                                       \
    }                                                                                                  \
  } while (0)

#define DESIRE_TRY(call)          \
  do {                            \
    int rc_ = (call);             \
    if (rc_ != DESIRE_OK) return rc_; \
  } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t cap, off;
  Workspace(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <class T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (off + bytes > cap) return nullptr;
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case DESIRE_ACT_RELU: return fmaxf(x, 0.f);
    case DESIRE_ACT_ELU: return x > 0.f ? x : (expf(x) - 1.f);
    case DESIRE_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- log-polar bin of d = pos_j - pos_i; exact arithmetic on the shared tables (oracle: logpolar_bin)
__device__ __forceinline__ int logpolar_bin(float dx, float dy, const float* r2e, int n_rad, const float* dirs,
                                            int n_ang) {
  const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  int rb = -1;
  for (int e = 0; e <= n_rad; ++e) rb += (r2 >= r2e[e]) ? 1 : 0;
  if (rb < 0 || rb >= n_rad) return -1;
  int ab = n_ang - 1;
  bool ge0 = __fsub_rn(__fmul_rn(dirs[0], dy), __fmul_rn(dirs[1], dx)) >= 0.f;
  bool ge = ge0;
  for (int s = 0; s < n_ang; ++s) {
    bool gn = (s + 1 < n_ang) ? (__fsub_rn(__fmul_rn(dirs[2 * (s + 1)], dy), __fmul_rn(dirs[2 * (s + 1) + 1], dx)) >= 0.f)
                              : ge0;
    if (ge && !gn) {
      ab = s;
      break;
    }
    ge = gn;
  }
  return rb * n_ang + ab;
}

// ---- internal cross-file entry points (all asynchronous on `st`)
// C[M,N] (ldc) = act(A @ B + bias) (+C if accumulate).  A via loader (see gemm_f32.cu), B [K,N] ldb,
// or, when trans_b, B stored [N,K] (ldb = K stride).
struct Im2col {
  // A(m,k): m = (img, oy, ox), k = (ky, kx, ci) over an NHWC input, zero outside
  int Hi, Wi, Ci, Ho, Wo, kh, kw, stride, pad_t, pad_l;
};
// pw: scratch for the packed BF16 weight image of the tcgen05 path (gemm_tc.cu); without it (or in
// gemm mode 0) the FP32 CUDA-core kernel runs.
struct PackWs {
  void* p = nullptr;
  size_t bytes = 0;
};
constexpr size_t PACK_WS_BYTES = 8u << 20;   // >= the largest packed weight of the path (venc fc / social fc)
int sgemm(const float* A, int lda, const float* B, int ldb, bool trans_b, const float* bias, float* C,
          int ldc, int M, int N, int K, int act, bool accumulate, cudaStream_t st, PackWs pw = PackWs());
int sgemm_im2col(const float* X, const Im2col& g, const float* B, int ldb, const float* bias, float* C,
                 int ldc, int M, int N, int K, int act, cudaStream_t st, PackWs pw = PackWs());
// conv5_tc.cu: 5x5 / stride 1 / SAME convolution over NHWC as an implicit GEMM on tcgen05 (the input tile is staged once
// as the A operand; a filter tap is a start-address shift).  X [B, H, W, Cin], w [(ky, kx, ci), n] (ldb), Y [B, H, W, ldc].
bool conv5_tc_eligible(const Im2col& g, int Cout, int ldc, int act, const PackWs& pw);
// the same kernel for 5x5 / stride 2 / SAME over a 3-channel image with even sides, staged space-to-depth (conv5_tc()
// takes either geometry)
bool conv5s2_tc_eligible(const Im2col& g, int Cout, int ldc, int act, const PackWs& pw);
int conv5_tc(const float* X, const Im2col& g, int B, const float* w, int ldb, const float* bias, float* Y, int ldc, int Cout,
             int act, cudaStream_t st, PackWs pw);
// train step (gemm_f32.cu): dW[Kd,N] (ldw) += A^T @ B over M rows, A dense or implicit im2col; column sums
// wp: scratch for the packed BF16 image of B (wgrad_tc_pack_bytes(M, N)); with it (and gemm mode != 0, Kd >= 64,
// N >= 16, M >= 2048) the product runs on tcgen05 (gemm_tc.cu, split-K + atomics), otherwise on FP32 CUDA cores.
int wgrad_tn(const float* A, int lda, const float* B, int ldb, float* dW, int ldw, int M, int Kd, int N,
             cudaStream_t st, PackWs wp = PackWs());
int wgrad_tn_im2col(const float* X, const Im2col& g, const float* B, int ldb, float* dW, int ldw, int M, int Kd,
                    int N, cudaStream_t st, PackWs wp = PackWs());
size_t wgrad_tc_pack_bytes(int rows, int N);
bool wgrad_tc_eligible(int rows, int Kd, int N, const void* pack_ws, size_t pack_bytes);
int wgrad_tc(const float* A, int lda, const float* B, int ldb, float* dW, int ldw, int rows, int Kd, int N, void* pack_ws,
             cudaStream_t st);
int wgrad_tc_im2col(const float* X, const Im2col& g, const float* B, int ldb, float* dW, int ldw, int rows, int Kd, int N,
                    void* pack_ws, cudaStream_t st);
int colsum_acc(const float* A, int lda, int M, int N, float* out, cudaStream_t st);
// train.cu: GRU backward through time over recomputed gates (shared by the encoders, Decoder-1 and Decoder-2)
struct GruBptt {
  int R, H, T, I;
  const float *wg, *wc;                 // full TF-layout kernels [(I+H),2H], [(I+H),H]
  const float* xp; long xp_rs, xp_ss;   // hoisted input projection incl. biases, [.., 3H] = (r|u|c)
  const float* hs; long hs_rs, hs_ss;   // forward states h_t
  const float* h0e;                     // [R,H] dense initial state
  float* dhs; long dhs_rs, dhs_ss;      // gradient reaching h_t (in), accumulated in place
  float* dxp; long dxp_rs, dxp_ss;      // += (zeroed by the caller)
  float* dh0;                           // [R,H], zeroed by the caller, receives d h_{-1}
  float *dwg, *dwc;                     // full-layout gradients (+=); only the state rows are touched here
};
size_t gru_bptt_ws_bytes(size_t R, int H, int T);
// after_step(t) runs on the host right after step t's kernels are enqueued (t = T-1 .. 0): Decoder-2 uses it to
// push the social-pooling gradient of step t into dhs[t-1] before step t-1 consumes it.
int gru_bptt(const GruBptt& a, void* ws, size_t ws_bytes, cudaStream_t st,
             const std::function<int(int)>* after_step = nullptr);
int col2im_gather(const float* col, int R, int Hin, int Hout, int k, int stride, int pad, int C, const float* bias,
                  float* out, cudaStream_t st);
// hoisted input projection of an encoder: xp[(m,t), :] = [x,y] @ W_x + b   ([rows, 3H] = r|u|c), traj [rows,3]
int xproj_traj(const float* traj, size_t rows, int H, const desire_gru_t* w, float* xp, cudaStream_t st);
// d <- d * act'(computed from the post-activation output `out`)
int act_bwd_post(const float* out, int ldo, float* d, int ldd, size_t M, int N, int act, cudaStream_t st);
int gemm_mode();
// A weight packed ONCE for many GEMM calls (IOC loop): pack_weight() fills `packed` when the tensor-core
// path will be used; gemm_packed() then skips the per-call packing.
struct PackedW {
  const float* W = nullptr;
  int ldw = 0;
  bool trans = false;
  int K = 0, N = 0;
  const void* packed = nullptr;
};
int pack_weight(PackedW& w, void* ws, size_t ws_bytes, cudaStream_t st);
int gemm_packed(const float* A, int lda, const PackedW& w, const float* bias, float* C, int ldc, int M, int act,
                bool accumulate, cudaStream_t st);
// Extra epilogue term of the tensor-core GEMM: C[m,:] += s[m,0] * P[g,0,:] + s[m,1] * P[g,1,:] with g = m / div.
// It is how the Decoder-2 input projection takes feature_pooling (model/model.py:291-311): row (r,t) of that tensor is
// [yhat_x * rho_i[m,:C] | yhat_y * rho_i[m,C:]], so its product with the weights is yhat_x * (rho_i[m,:C] @ W_x) +
// yhat_y * (rho_i[m,C:] @ W_y) — two per-AGENT vectors (P) scaled by two per-row scalars (s), not a K = 2C GEMM.
struct Rank2 {
  const float* s = nullptr;   // [M,2]
  const float* P = nullptr;   // [groups, 2, N]
  int div = 1;                // rows per group
};
bool gemm_packed_r2(const float* A, int lda, const PackedW& w, const float* bias, float* C, int ldc, int M, int act,
                    const Rank2& r2, cudaStream_t st, int* rc);
// C = [A1 | A2] @ W + bias (K1 columns from A1, the rest from A2) on the tensor-core path; false = not eligible
bool gemm_packed_dual(const float* A1, int lda1, int K1, const float* A2, int lda2, const PackedW& w, const float* bias,
                      float* C, int ldc, int M, int act, cudaStream_t st, int* rc);
// fully fused social pooling + fc on tensor cores (social_tc.cu): fsp = relu(pool(h) @ sp_w + b)
struct SocialFcArgs {
  const float* pos;      // position of row r at pos + r*pos_stride (x,y)
  long pos_stride;
  const float* h;        // [R,H] hidden vectors (row stride ld_h)
  int ld_h;
  const float* obs;      // [B*N,Tp,3] ids for the existence mask
  int Tp, B, N, K, H, n_rad, n_ang;
  const float* r2_edges;
  const float* dirs;
  const void* packed;    // sp_w [G*H, H] packed for N-tile H (tc_pack_b / pack_weight)
  const float* bias;
  float* out;            // fsp [R,H]
};
// social pooling for 129..256 agents per scene as a tcgen05 MMA (social_pm.cu), materialising like social_pool_launch
bool social_pool_mma_eligible(const float* h, int ld_h, int N, int H, int n_rad, int n_ang, const float* pooled);
int social_pool_mma(const float* pos, long pos_stride, const float* h, int ld_h, const float* obs, int Tp, int B, int N,
                    int K, int H, int n_rad, int n_ang, const float* r2_edges, const float* dirs, float* pooled,
                    cudaStream_t st);
// fused pooling + fc for 129..256 agents per scene, H in {128, 256} (social_fm.cu); social_fc_tc() takes it when eligible
bool social_fc_fm_eligible(const SocialFcArgs& a);
int social_fc_fm(const SocialFcArgs& a, cudaStream_t st);
bool social_fc_tc_eligible(const SocialFcArgs& a);
int social_fc_tc(const SocialFcArgs& a, cudaStream_t st);
// second design (social_ts.cu): pooled A operand in tensor memory; social_fc_tc() takes it whenever it is eligible
bool social_fc_ts_eligible(const SocialFcArgs& a);
int social_fc_ts(const SocialFcArgs& a, cudaStream_t st);
size_t gemm_tc_pack_bytes(int N, int K);
bool gemm_tc_eligible(int M, int N, int K, const void* pack_ws, size_t pack_bytes);
int gemm_tc(const float* A, int lda, const float* W, int ldw, bool trans_b, const float* bias, float* C, int ldc,
            int M, int N, int K, int act, bool accumulate, void* pack_ws, cudaStream_t st);
int gemm_tc_im2col(const float* X, const Im2col& g, const float* W, int ldw, const float* bias, float* C, int ldc,
                   int M, int N, int K, int act, void* pack_ws, cudaStream_t st);

// GRU recurrence (gru.cu).  xp: hoisted input projection incl. biases, [.., 3H] = (r|u|c) per row.
struct GruSeqArgs {
  int R, H, T;
  const float* xp;          // may be null when traj != null
  long xp_row_stride;       // floats
  long xp_step_stride;      // 0 => constant input
  const float* traj;        // [R,T,3] (id,x,y): input width 2 projected inline with wx_g/wx_c rows 0,1
  const float *wx_g, *wx_c, *bg, *bc;
  const float* ex;          // extra per-step operand [R,Ka] (only with T==1), or null
  int Ka, ld_ex;
  const float* w_g;         // [(Ka+H), 2H]: rows 0..Ka-1 multiply ex, the rest multiply h
  const float* w_c;         // [(Ka+H), H]
  const float* h0;          // null => zeros; row = r / h0_div, stride ld_h0
  int h0_div, ld_h0;
  float* hs;                // all states: hs + r*hs_row_stride + t*hs_step_stride, or null
  long hs_row_stride, hs_step_stride;
  float* h_final;           // [R] rows, stride ld_hf, or null
  int ld_hf;
  const void* packed = nullptr;   // optional: weights already packed (skips per-call packing) ...
  int packed_fmt = 0;             // ... by gru_tc_pack (0, n-tile 2H) or gru_tc3_pack (3, n-tile H)
};
int gru_tc_pack(const float* w_g, const float* w_c, int H, int Ka, void* ws, size_t ws_bytes, cudaStream_t st);
int gru_seq(const GruSeqArgs& a, cudaStream_t st, PackWs pw = PackWs());
// tcgen05 recurrence (gru_tc.cu): taken when H % 32 == 0, H <= 256, the input projection is hoisted (xp)
// and scratch for the packed weights is available
size_t tc_pack_bytes(int K, int N, int BN);
int tc_pack_b(const float* W, int ldw, bool trans, int K, int N, int BN, void* out, cudaStream_t st);
size_t gru_tc_pack_bytes(int H, int Ka = 0);
bool gru_tc_eligible(const GruSeqArgs& a, const void* pack_ws, size_t pack_bytes);
int gru_seq_tc(const GruSeqArgs& a, void* pack_ws, cudaStream_t st);

// FP32 row-major [rows, row_stride] tensor -> CUtensorMap (cuda.h) with [128 rows x 32 columns] boxes and the 128-byte
// swizzle, through the driver entry point fetched at run time (the library does not link libcuda)
int make_tmap_rows32(::CUtensorMap_st* tm, const float* base, long rows, long row_stride);

// third design of the tcgen05 recurrence (gru_tc3.cu): register-resident state, TMA-staged per-row inputs, phases
// overlapped with the MMA groups; H in {128, 256} without extra operand, H = 128 with it (one Decoder-2 step).
// Weights packed with n-tile = H.
size_t gru_tc3_pack_bytes(int H, int Ka = 0);
int gru_tc3_pack(const float* w_g, const float* w_c, int H, int Ka, void* ws, size_t ws_bytes, cudaStream_t st);
bool gru_tc3_eligible(const GruSeqArgs& a, const void* pack_ws, size_t pack_bytes);
int gru_seq_tc3(const GruSeqArgs& a, void* pack_ws, cudaStream_t st);

// fused transposed conv + bias + per-row BN + activation on tcgen05 (deconv_tc.cu)
size_t deconv_tc_pack_bytes(int Cin, int Cout, int ks);
bool deconv_tc_eligible(int R, int Hin, int Hout, int Cin, int Cout, int ks, int stride, const void* pack_ws,
                        size_t pack_bytes);
// fuse4 (optional, only for the 8x8 -> 16x16 stride-2 layer with Cout == 32): the decoder's last layer (16x16x32 -> 32x32x1,
// k5 s2 SAME, BN + activation) runs on the normalised tile while it is still in shared memory; Y is then not written.
struct DeconvFuse4 {
  const float *W, *bias, *gamma, *beta;   // W [5,5,1,32]
  int act;
  float* Y;                               // [R, 1024]
};
int deconv_tc(const float* X, int R, int Hin, int Hout, int Cin, int Cout, int ks, int stride, int pad, const float* W,
              const float* bias, const float* gamma, const float* beta, int act, float* Y, void* pack_ws,
              cudaStream_t st, const DeconvFuse4* fuse4 = nullptr);

bool deconv1c_tc_eligible();
int deconv1c_tc(const float* X, int R, const float* W, const float* bias, const float* gamma, const float* beta, int act,
                float* Y, void* pack_ws, cudaStream_t st);

// col2im + bias + per-row BN + activation (cvae.cu): col [R*Hin*Hin, k*k*Cout] -> out [R,Hout,Hout,Cout]
// ypre (optional, train step): also stores the pre-BN values (col2im + bias) [R,Hout,Hout,Cout]
int colbn_act(const float* col, int R, int Hin, int Hout, int k, int stride, int pad, int Cout,
              const float* bias, const float* gamma, const float* beta, int act, float* out,
              cudaStream_t st, float* ypre = nullptr);

}  // namespace desire
