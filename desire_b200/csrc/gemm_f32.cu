// FP32 CUDA-core GEMM with fused bias/activation epilogue and pluggable A-operand loaders
// (dense rows, or im2col gather over an NHWC image for the forward convolutions).
// This is the exact-arithmetic baseline path; the tcgen05 path (gemm_tc.cu) takes over the
// large GEMMs, this one stays for odd shapes and as the in-library cross-check.
#include "common.cuh"

namespace desire {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4;

struct DenseA {
  const float* A;
  int lda, M, K;
  __device__ __forceinline__ float operator()(int m, int k) const {
    return (m < M && k < K) ? __ldg(A + (size_t)m * lda + k) : 0.f;
  }
};

struct Im2colA {
  const float* X;
  Im2col g;
  int M, K;
  __device__ __forceinline__ float operator()(int m, int k) const {
    if (m >= M || k >= K) return 0.f;
    int ox = m % g.Wo;
    int t = m / g.Wo;
    int oy = t % g.Ho;
    int img = t / g.Ho;
    int ci = k % g.Ci;
    int t2 = k / g.Ci;
    int kx = t2 % g.kw;
    int ky = t2 / g.kw;
    int iy = oy * g.stride + ky - g.pad_t;
    int ix = ox * g.stride + kx - g.pad_l;
    if (iy < 0 || iy >= g.Hi || ix < 0 || ix >= g.Wi) return 0.f;
    return __ldg(X + (((size_t)img * g.Hi + iy) * g.Wi + ix) * g.Ci + ci);
  }
};

template <class ALoader, bool TRANS_B>
__global__ void __launch_bounds__(NT) sgemm_kernel(ALoader A, const float* __restrict__ B, int ldb,
                                                   const float* __restrict__ bias, float* __restrict__ C,
                                                   int ldc, int M, int N, int K, int act, int accumulate) {
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int m_blk = blockIdx.y * BM, n_blk = blockIdx.x * BN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      const int k = tid % BK, m0 = tid / BK;
#pragma unroll
      for (int i = 0; i < BM / (NT / BK); ++i) {
        int m = m0 + i * (NT / BK);
        As[k][m] = A(m_blk + m, k0 + k);
      }
    }
    if (TRANS_B) {
      const int k = tid % BK, n0 = tid / BK;
#pragma unroll
      for (int i = 0; i < BN / (NT / BK); ++i) {
        int n = n0 + i * (NT / BK);
        int gn = n_blk + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < K) ? __ldg(B + (size_t)gn * ldb + gk) : 0.f;
      }
    } else {
      const int n = tid % BN, kk0 = tid / BN;
#pragma unroll
      for (int i = 0; i < BK / (NT / BN); ++i) {
        int k = kk0 + i * (NT / BN);
        int gn = n_blk + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < K) ? __ldg(B + (size_t)gk * ldb + gn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m_blk + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n_blk + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      float* c = C + (size_t)m * ldc + n;
      if (accumulate) v += *c;
      *c = act_apply(v, act);
    }
  }
}

template <class L>
int launch(const L& a, const float* B, int ldb, bool trans_b, const float* bias, float* C, int ldc, int M,
           int N, int K, int act, bool accumulate, cudaStream_t st) {
  if (M == 0 || N == 0) return DESIRE_OK;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if (grid.y > 65535) {
    set_error("sgemm: M=%d too large for one launch", M);
    return DESIRE_ERR_INVALID;
  }
  if (trans_b)
    DESIRE_LAUNCH(st, (sgemm_kernel<L, true><<<grid, NT, 0, st>>>(a, B, ldb, bias, C, ldc, M, N, K, act, accumulate ? 1 : 0)));
  else
    DESIRE_LAUNCH(st, (sgemm_kernel<L, false><<<grid, NT, 0, st>>>(a, B, ldb, bias, C, ldc, M, N, K, act, accumulate ? 1 : 0)));
  return DESIRE_OK;
}


// ---- weight-gradient GEMM (train step): dW[Kd,N] += sum_m A(m,kd) * B[m,n], FP32 CUDA cores.
// The reduction runs over the ROWS (up to millions), so the grid splits them (blockIdx.z) and the
// partial tiles are added with atomicAdd (the caller zeroes dW once per step).  Loads are coalesced
// along kd / n; A comes through the same loaders as the forward GEMM (dense rows or implicit im2col).
constexpr int WBN = 64, WBK = 16;
template <class ALoader, int WBM>      // WBM = kd rows per tile: 128, or 32 for the narrow filters (Kd <= 32)
__global__ void __launch_bounds__(NT) wgrad_tn_kernel(ALoader A, const float* __restrict__ B, int ldb,
                                                      float* __restrict__ dW, int ldw, int M, int Kd, int N,
                                                      int rows_per_split) {
  __shared__ __align__(16) float As[WBK][WBM + 4];
  __shared__ __align__(16) float Bs[WBK][WBN];
  const int tid = threadIdx.x;
  const int k_blk = blockIdx.y * WBM, n_blk = blockIdx.x * WBN;
  const int ty = tid / 16, tx = tid % 16;
  const long m_begin = (long)blockIdx.z * rows_per_split;
  const long m_end = (m_begin + rows_per_split < M) ? m_begin + rows_per_split : M;
  constexpr int RI = WBM / 16;           // kd rows per thread
  float acc[RI][4];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long m0 = m_begin; m0 < m_end; m0 += WBK) {
    {
      const int kd = tid % WBM, r0 = tid / WBM;          // 2 rows per pass
#pragma unroll
      for (int i = 0; i < WBK / (NT / WBM); ++i) {
        const int r = r0 + i * (NT / WBM);
        const long m = m0 + r;
        As[r][kd] = (m < m_end) ? A((int)m, k_blk + kd) : 0.f;
      }
    }
    {
      const int n = tid % WBN, r0 = tid / WBN;           // 4 rows per pass
#pragma unroll
      for (int i = 0; i < WBK / (NT / WBN); ++i) {
        const int r = r0 + i * (NT / WBN);
        const long m = m0 + r;
        const int gn = n_blk + n;
        Bs[r][n] = (m < m_end && gn < N) ? __ldg(B + (size_t)m * ldb + gn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WBK; ++k) {
      float av[RI];
#pragma unroll
      for (int i = 0; i < RI; ++i) av[i] = As[k][ty * RI + i];
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    const int kd = k_blk + ty * RI + i;
    if (kd >= Kd) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n_blk + tx * 4 + j;
      if (n < N) atomicAdd(dW + (size_t)kd * ldw + n, acc[i][j]);
    }
  }
}

template <class L>
int launch_wgrad(const L& a, const float* B, int ldb, float* dW, int ldw, int M, int Kd, int N, cudaStream_t st) {
  if (M == 0 || N == 0 || Kd == 0) return DESIRE_OK;
  const int WBM = Kd <= 32 ? 32 : 128;
  const int gx = (N + WBN - 1) / WBN, gy = (Kd + WBM - 1) / WBM;
  // ~4 CTAs per SM in flight; every split at least 256 rows
  int splits = (148 * 4 + gx * gy - 1) / (gx * gy);
  const int max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int rows = (M + splits - 1) / splits;
  rows = (rows + WBK - 1) / WBK * WBK;
  splits = (M + rows - 1) / rows;
  dim3 grid(gx, gy, splits);
  if (WBM == 32)
    DESIRE_LAUNCH(st, (wgrad_tn_kernel<L, 32><<<grid, NT, 0, st>>>(a, B, ldb, dW, ldw, M, Kd, N, rows)));
  else
    DESIRE_LAUNCH(st, (wgrad_tn_kernel<L, 128><<<grid, NT, 0, st>>>(a, B, ldb, dW, ldw, M, Kd, N, rows)));
  return DESIRE_OK;
}

// out[n] += sum_m A[m, n]
__global__ void colsum_kernel(const float* __restrict__ A, int lda, int M, int N, int rows_per_block,
                              float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long m0 = (long)blockIdx.y * rows_per_block;
  const long m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  float s = 0.f;
  for (long m = m0; m < m1; ++m) s += __ldg(A + (size_t)m * lda + n);
  atomicAdd(out + n, s);
}

}  // namespace

int sgemm(const float* A, int lda, const float* B, int ldb, bool trans_b, const float* bias, float* C, int ldc,
          int M, int N, int K, int act, bool accumulate, cudaStream_t st, PackWs pw) {
  if (gemm_tc_eligible(M, N, K, pw.p, pw.bytes))
    return gemm_tc(A, lda, B, ldb, trans_b, bias, C, ldc, M, N, K, act, accumulate, pw.p, st);
  note_fallback(DESIRE_FALLBACK_GEMM_FP32, "GEMM on FP32 CUDA cores (M, N, K)", M, N, K);
  // split very tall problems so gridDim.y stays legal
  const int MAXM = 65535 * BM;
  for (long m0 = 0; m0 < M; m0 += MAXM) {
    int mm = (int)((M - m0) < MAXM ? (M - m0) : MAXM);
    DenseA a{A + (size_t)m0 * lda, lda, mm, K};
    DESIRE_TRY(launch(a, B, ldb, trans_b, bias, C + (size_t)m0 * ldc, ldc, mm, N, K, act, accumulate, st));
  }
  return DESIRE_OK;
}

int sgemm_im2col(const float* X, const Im2col& g, const float* B, int ldb, const float* bias, float* C, int ldc,
                 int M, int N, int K, int act, cudaStream_t st, PackWs pw) {
  if (gemm_tc_eligible(M, N, K, pw.p, pw.bytes))
    return gemm_tc_im2col(X, g, B, ldb, bias, C, ldc, M, N, K, act, pw.p, st);
  note_fallback(DESIRE_FALLBACK_GEMM_FP32, "implicit-GEMM convolution on FP32 CUDA cores (M, N, K)", M, N, K);
  DESIRE_CHECK_ARG((M + BM - 1) / BM <= 65535, "sgemm_im2col: M=%d too large", M);
  Im2colA a{X, g, M, K};
  return launch(a, B, ldb, false, bias, C, ldc, M, N, K, act, false, st);
}


int wgrad_tn(const float* A, int lda, const float* B, int ldb, float* dW, int ldw, int M, int Kd, int N,
             cudaStream_t st, PackWs wp) {
  if (wgrad_tc_eligible(M, Kd, N, wp.p, wp.bytes)) return wgrad_tc(A, lda, B, ldb, dW, ldw, M, Kd, N, wp.p, st);
  DenseA a{A, lda, M, Kd};
  return launch_wgrad(a, B, ldb, dW, ldw, M, Kd, N, st);
}

int wgrad_tn_im2col(const float* X, const Im2col& g, const float* B, int ldb, float* dW, int ldw, int M, int Kd,
                    int N, cudaStream_t st, PackWs wp) {
  if (wgrad_tc_eligible(M, Kd, N, wp.p, wp.bytes)) return wgrad_tc_im2col(X, g, B, ldb, dW, ldw, M, Kd, N, wp.p, st);
  Im2colA a{X, g, M, Kd};
  return launch_wgrad(a, B, ldb, dW, ldw, M, Kd, N, st);
}

int colsum_acc(const float* A, int lda, int M, int N, float* out, cudaStream_t st) {
  if (M == 0 || N == 0) return DESIRE_OK;
  const int threads = N >= 128 ? 128 : 32;
  const int gx = (N + threads - 1) / threads;
  int gy = (148 * 8 + gx - 1) / gx;
  const int max_gy = (M + 63) / 64;
  if (gy > max_gy) gy = max_gy;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  const int rows = (M + gy - 1) / gy;
  gy = (M + rows - 1) / rows;
  DESIRE_LAUNCH(st, (colsum_kernel<<<dim3(gx, gy), threads, 0, st>>>(A, lda, M, N, rows, out)));
  return DESIRE_OK;
}

}  // namespace desire

extern "C" int desire_fc_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C,
                             int ldc, int M, int N, int K, int act, int accumulate, desire_stream_t stream) {
  DESIRE_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0, "desire_fc_fwd: bad arguments");
  DESIRE_CHECK_ARG(lda >= K && ldw >= N && ldc >= N, "desire_fc_fwd: leading dimensions too small");
  return desire::sgemm(A, lda, W, ldw, false, bias, C, ldc, M, N, K, act, accumulate != 0, (cudaStream_t)stream);
}
