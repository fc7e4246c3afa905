// tcgen05 GEMM with FP32 interface: C[M,N] = act(A[M,K] @ W[K,N] + bias) (+C).
//
// sm_100a design (one 128 x BN output tile per CTA, up to two CTAs resident per SM so one tile's
// epilogue overlaps the other's main loop):
//   warps 0-3  A producers: FP32 rows (dense, or im2col-gathered from an NHWC image) are loaded with
//              128-bit loads one K stage ahead, split into BF16 hi/lo and written to shared memory in the
//              canonical UMMA K-major layout (tc.cuh); afterwards the same warps run the epilogue
//              (tcgen05.ld of their 32 TMEM lanes -> bias/activation -> 128-byte row segments to HBM).
//   warp 4     MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM
//              accumulator, 3 MMAs per K step in 3xBF16 mode (hi*hi + lo*hi + hi*lo), and frees each
//              shared-memory stage with tcgen05.commit.
//   warp 5     B loader: weights are pre-packed (pack_b_kernel) into per-(n-tile, k-stage) shared-memory
//              images, so a stage is ONE 1-D bulk TMA copy (cp.async.bulk) completing on the stage's
//              mbarrier.
// Full/empty mbarrier ring of `stages` stages; accumulator handed to the epilogue by tcgen05.commit.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc.cuh"

namespace desire {

static int g_gemm_mode = 3;  // 0: FP32 CUDA cores, 1: single BF16 pass, 3: 3xBF16 (parity mode)
int gemm_mode() { return g_gemm_mode; }

namespace {

using namespace tc;

constexpr int TM = 128;       // rows per tile == TMEM lanes
constexpr int BK = 32;        // K elements per stage (4 chunks of 8)
constexpr int KC = BK / 8;
constexpr int NTHR = 192;

// ---- A-operand loaders.  init(m) hoists everything that depends only on the row (pointer arithmetic,
// im2col / neighbour-list decoding) out of the K loop; load_stage(ks, v) fetches one 32-wide K stage
// (4 chunks of 8 FP32) of that row into registers.
struct DenseA8 {
  const float* A;
  int lda, M, K;
  bool vec;
  const float* row;
  __device__ __forceinline__ void init(int m) { row = m < M ? A + (size_t)m * lda : nullptr; }
  __device__ __forceinline__ void load_stage(int ks, float (*v)[8]) const {
    const int k0 = ks * BK;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const int k = k0 + c * 8;
      if (row && vec && k + 8 <= K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(row + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(row + k) + 1);
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
        v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] = (row && k + i < K) ? __ldg(row + k + i) : 0.f;
      }
    }
  }
};

// A = [A1 | A2] column-wise (two row-major sources, K1 a multiple of 8): lets the Decoder-2 input projection read
// feature_pooling in place instead of from a copy inside a concatenated feature matrix
struct DualDenseA8 {
  const float *A1, *A2;
  int lda1, lda2, K1, M, K;
  const float *r1, *r2;
  __device__ __forceinline__ void init(int m) {
    r1 = m < M ? A1 + (size_t)m * lda1 : nullptr;
    r2 = m < M ? A2 + (size_t)m * lda2 - K1 : nullptr;      // indexed with the global k
  }
  __device__ __forceinline__ void load_stage(int ks, float (*v)[8]) const {
    const int k0 = ks * BK;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const int k = k0 + c * 8;
      const float* src = k < K1 ? r1 : r2;                   // a chunk of 8 never straddles K1
      if (r1 && k + 8 <= K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + k) + 1);
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
        v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c][i] = (r1 && k + i < K) ? __ldg(src + k + i) : 0.f;
      }
    }
  }
};

struct Im2colA8 {
  const float* X;
  Im2col g;
  int M, K;
  const float* img;   // start of this row's image, null for rows past M
  int iy0, ix0;       // top-left input coordinate of the receptive field
  __device__ __forceinline__ void init(int m) {
    img = nullptr;
    if (m >= M) return;
    const int ox = m % g.Wo;
    const int t = m / g.Wo;
    const int oy = t % g.Ho;
    img = X + (size_t)(t / g.Ho) * g.Hi * g.Wi * g.Ci;
    iy0 = oy * g.stride - g.pad_t;
    ix0 = ox * g.stride - g.pad_l;
  }
  __device__ __forceinline__ void load_stage(int ks, float (*v)[8]) const {
    const int k0 = ks * BK;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
    }
    if (!img) return;
    if (g.Ci % 8 == 0) {   // 8 consecutive k share (ky,kx): one contiguous 32-byte read per chunk
      int t2 = k0 / g.Ci, ci = k0 - t2 * g.Ci;
      int ky = t2 / g.kw, kx = t2 - ky * g.kw;
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        if (k0 + c * 8 < K) {
          const int iy = iy0 + ky, ix = ix0 + kx;
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) {
            const float4* p = reinterpret_cast<const float4*>(img + ((size_t)iy * g.Wi + ix) * g.Ci + ci);
            const float4 a = __ldg(p), b = __ldg(p + 1);
            v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
            v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
          }
        }
        ci += 8;
        if (ci >= g.Ci) {
          ci = 0;
          if (++kx == g.kw) { kx = 0; ++ky; }
        }
      }
    } else {
      int t2 = k0 / g.Ci, ci = k0 - t2 * g.Ci;
      int ky = t2 / g.kw, kx = t2 - ky * g.kw;
#pragma unroll
      for (int c = 0; c < KC; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (k0 + c * 8 + i < K) {
            const int iy = iy0 + ky, ix = ix0 + kx;
            if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) v[c][i] = __ldg(img + ((size_t)iy * g.Wi + ix) * g.Ci + ci);
          }
          if (++ci == g.Ci) {
            ci = 0;
            if (++kx == g.kw) { kx = 0; ++ky; }
          }
        }
      }
    }
  }
};

// ---- transposed loaders for the weight-gradient GEMM dW[Kd,N] = A^T[Kd,rows] @ B[rows,N]: the GEMM's "row" is kd
// (one thread each) and its K dimension runs over the sample rows, so a warp's 32 threads read 32 consecutive kd of
// one sample row = one coalesced 128-byte request, and the values land directly in the K-major UMMA layout.
struct TransDenseA8 {
  const float* A;      // [rows, Kd] row-major, lda
  int lda, M /*= Kd*/, K /*= rows*/;
  const float* col;    // A + kd, null past Kd
  __device__ __forceinline__ void init(int m) { col = m < M ? A + m : nullptr; }
  __device__ __forceinline__ void load_stage(int ks, float (*v)[8]) const {
    const long r0 = (long)ks * BK;
#pragma unroll
    for (int c = 0; c < KC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long r = r0 + c * 8 + i;
        v[c][i] = (col && r < K) ? __ldg(col + r * lda) : 0.f;
      }
  }
};

struct TransIm2colA8 {
  const float* X;      // NHWC image the patches are cut from
  Im2col g;
  int M /*= Kd = kh*kw*Ci*/, K /*= rows = imgs*Ho*Wo*/;
  int ky, kx, ci;      // this thread's (tap, channel); ci < 0 past Kd
  __device__ __forceinline__ void init(int m) {
    ci = -1;
    if (m >= M) return;
    ci = m % g.Ci;
    const int t2 = m / g.Ci;
    kx = t2 % g.kw;
    ky = t2 / g.kw;
  }
  __device__ __forceinline__ void load_stage(int ks, float (*v)[8]) const {
    const long r0 = (long)ks * BK;
    int ox = (int)(r0 % g.Wo);
    long t = r0 / g.Wo;
    int oy = (int)(t % g.Ho);
    long img = t / g.Ho;
#pragma unroll
    for (int c = 0; c < KC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = 0.f;
        if (ci >= 0 && r0 + c * 8 + i < K) {
          const int iy = oy * g.stride + ky - g.pad_t, ix = ox * g.stride + kx - g.pad_l;
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
            x = __ldg(X + (((size_t)img * g.Hi + iy) * g.Wi + ix) * g.Ci + ci);
        }
        v[c][i] = x;
        if (++ox == g.Wo) {
          ox = 0;
          if (++oy == g.Ho) {
            oy = 0;
            ++img;
          }
        }
      }
  }
};

// W [K,N] (or [N,K] when trans) FP32 -> packed[(jn*nks + ks)] = { hi: [KC][BN][8] bf16, lo: same }
__global__ void pack_b_kernel(const float* __restrict__ W, int ldw, int trans, int K, int N, int BN, int nks,
                              int ntn, uint4* __restrict__ out) {
  const long total = (long)ntn * nks * KC * BN;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % BN);
  long t = idx / BN;
  const int c = (int)(t % KC);
  t /= KC;
  const int ks = (int)(t % nks);
  const int jn = (int)(t / nks);
  const int gn = jn * BN + n, k0 = ks * BK + c * 8;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + i;
    v[i] = (gn < N && k < K) ? (trans ? __ldg(W + (size_t)gn * ldw + k) : __ldg(W + (size_t)k * ldw + gn)) : 0.f;
  }
  Split8 s = split8(v);
  uint4* stage = out + ((size_t)jn * nks + ks) * (2 * KC * BN);
  stage[c * BN + n] = s.hi;
  stage[KC * BN + c * BN + n] = s.lo;
}

// Epilogue of one 32-column chunk held in registers: rank-2 term, bias, activation.  The bias and the rank-2 vectors are
// read as float4 (16-byte aligned: n0 % 32 == 0, N % 4 == 0) before any arithmetic and the activation is selected ONCE per
// chunk: with a per-element `if (bias) x += __ldg(..); x = act_apply(x, act)` the branches of act_apply kept every load
// behind the previous element's — 32 serial L2 round trips, ~7 k cycles per chunk, which is what bounded the tall GEMMs
// (the Decoder-2 input projection spent 21 us per 128 x 192 tile in it).
__device__ __forceinline__ void chunk_epilogue(float (&acc)[32], const float* bias_c, const float* p0, int N, float s0,
                                               float s1, int act) {
  if (p0) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p0 + j));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p0 + N + j));
      acc[j] = fmaf(s0, a.x, fmaf(s1, b.x, acc[j]));
      acc[j + 1] = fmaf(s0, a.y, fmaf(s1, b.y, acc[j + 1]));
      acc[j + 2] = fmaf(s0, a.z, fmaf(s1, b.z, acc[j + 2]));
      acc[j + 3] = fmaf(s0, a.w, fmaf(s1, b.w, acc[j + 3]));
    }
  }
  if (bias_c) {
    float4 bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(bias_c) + j);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[4 * j] += bv[j].x; acc[4 * j + 1] += bv[j].y; acc[4 * j + 2] += bv[j].z; acc[4 * j + 3] += bv[j].w;
    }
  }
  if (act != DESIRE_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = act_apply(acc[j], act);
  }
}

template <class ALoad>
__global__ void __launch_bounds__(NTHR) gemm_tc_kernel(ALoad A_, const uint4* __restrict__ Bp,
                                                       const float* __restrict__ bias, float* __restrict__ C, int ldc,
                                                       int M, int N, int BN, int nks_total, int stages, int act,
                                                       int accumulate, int passes, uint32_t tmem_cols, int ks_split,
                                                       Rank2 r2, const __grid_constant__ CUtensorMap tm_c, int tma_c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_half = KC * TM * 16;              // bytes of A_hi (== A_lo) per stage
  const int b_half = KC * BN * 16;
  const int stage_bytes = 2 * a_half + 2 * b_half;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jn = blockIdx.x, m0 = blockIdx.y * TM;
  ALoad A = A_;
  // split-K (weight gradients): slice blockIdx.z runs K stages [ks_begin, ks_begin + nks) and adds its partial tile
  // with atomics (accumulate == 2)
  const int ks_begin = ks_split > 0 ? (int)blockIdx.z * ks_split : 0;
  const int nks = ks_split > 0 ? min(ks_split, nks_total - ks_begin) : nks_total;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 5);    // 4 producer warps + the B loader's expect_tx arrive
      mbar_init(&empty[s], 1);   // one tcgen05.commit
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc_dyn(tslot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < 4) {
    // ===================== A producers (thread <-> row); the loads of stages ks+1 and ks+2 are in flight
    // while stage ks is converted, so each thread keeps 256 bytes outstanding
    const int m = m0 + tid;
    A.init(m);
    float v0[KC][8], v1[KC][8];
    A.load_stage(ks_begin, v0);
    if (nks > 1) A.load_stage(ks_begin + 1, v1);
    RingPos rp;                                                  // emit() is called for ks = 0, 1, 2, .. in order
    auto emit = [&](int ks, float (*v)[8]) {
      const uint32_t slot = rp.slot, ph = rp.ph;
      rp.next(stages);
      mbar_wait(&empty[slot], ph ^ 1);
      uint8_t* sa = smem + (size_t)slot * stage_bytes;
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        Split8 s = split8(v[c]);
        *reinterpret_cast<uint4*>(sa + c * TM * 16 + tid * 16) = s.hi;
        *reinterpret_cast<uint4*>(sa + a_half + c * TM * 16 + tid * 16) = s.lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
    };
    // manually unrolled by two so each register buffer is addressed statically (loads stay asynchronous)
    for (int ks = 0; ks < nks; ks += 2) {
      emit(ks, v0);
      if (ks + 2 < nks) A.load_stage(ks_begin + ks + 2, v0);
      if (ks + 1 < nks) {
        emit(ks + 1, v1);
        if (ks + 3 < nks) A.load_stage(ks_begin + ks + 3, v1);
      }
    }
    // ===================== epilogue: TMEM lanes 32*warp .. +31 == rows m0 + tid
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const int n0 = jn * BN;
    const bool bias_vec = bias && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && N % 4 == 0;
    if (tma_c) {
      // Output through shared memory + TMA tensor-map stores (SASS UTMASTG): the thread-per-row float4 stores below
      // touch 32 different 128-byte lines per instruction, which throttled the huge-M GEMMs (Decoder-2 input projection:
      // 708 MB of output per launch) on the memory-instruction queue.  Here a thread writes its row of a [128 x 32] box
      // into the (now idle) stage ring with the 128-byte swizzle, one thread stores the box, two boxes alternate.
      const uint32_t obox = smem_u32(smem);
      int ci = 0;
      for (int c0 = 0; c0 < BN && n0 + c0 < N; c0 += 32, ++ci) {
        float acc[32];
        tmem_ld32(trow + c0, acc);
        tmem_ld_wait();
        {
          float s0 = 0.f, s1 = 0.f;
          const float* p0 = nullptr;
          if (r2.P && m < M) {
            s0 = __ldg(r2.s + 2 * (size_t)m);
            s1 = __ldg(r2.s + 2 * (size_t)m + 1);
            p0 = r2.P + (size_t)(m / r2.div) * 2 * N + n0 + c0;
          }
          const bool late_bias = bias && !bias_vec;            // unaligned bias: scalar loads, activation afterwards
          chunk_epilogue(acc, bias_vec ? bias + n0 + c0 : nullptr, p0, N, s0, s1, late_bias ? DESIRE_ACT_NONE : act);
          if (late_bias) {
            float bv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) bv[j] = __ldg(bias + n0 + c0 + j);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += bv[j];
            if (act != DESIRE_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = act_apply(acc[j], act);
            }
          }
        }
        const uint32_t box = obox + (ci & 1) * (TM * 128);
        if (ci >= 2) {                                   // the store that last read this box has finished reading it
          if (tid == 0) tma_store_wait_read1();
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          sts128(swz128_abs(box, tid, c16), make_float4(acc[4 * c16], acc[4 * c16 + 1], acc[4 * c16 + 2], acc[4 * c16 + 3]));
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
          tma_store_2d(&tm_c, smem + (ci & 1) * (TM * 128), n0 + c0, m0);
          tma_store_commit();
        }
      }
      if (tid == 0) tma_store_wait_all();
    } else
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float acc[32];
      const int ncol = min(32, BN - c0);          // BN is a multiple of 16
      if (ncol == 32) tmem_ld32(trow + c0, acc);
      else tmem_ld16(trow + c0, acc);
      tmem_ld_wait();
      if (m < M) {
        float* crow = C + (size_t)m * ldc + n0 + c0;
        if (r2.P) {     // + s[m,0] * P[g,0,:] + s[m,1] * P[g,1,:], g = m / div (a per-group rank-2 term, see Rank2)
          const float s0 = __ldg(r2.s + 2 * (size_t)m), s1 = __ldg(r2.s + 2 * (size_t)m + 1);
          const float* p0 = r2.P + (size_t)(m / r2.div) * 2 * N + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol && n0 + c0 + j < N) acc[j] = fmaf(s0, __ldg(p0 + j), fmaf(s1, __ldg(p0 + N + j), acc[j]));
        }
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0) && (n0 + c0 + ncol <= N);
        if (accumulate == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j >= ncol || n0 + c0 + j >= N) break;
            atomicAdd(crow + j, acc[j]);
          }
        } else {
          // bias first, all loads before any arithmetic or branch (see chunk_epilogue), then (+C), activation, store
          if (bias) {
            float bv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) bv[j] = (j < ncol && n0 + c0 + j < N) ? __ldg(bias + n0 + c0 + j) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += bv[j];
          }
          if (vec_ok) {
            if (accumulate) {
              float4 old[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                old[j] = 4 * j < ncol ? *reinterpret_cast<const float4*>(crow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                acc[4 * j] += old[j].x; acc[4 * j + 1] += old[j].y; acc[4 * j + 2] += old[j].z; acc[4 * j + 3] += old[j].w;
              }
            }
            if (act != DESIRE_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = act_apply(acc[j], act);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j >= ncol) break;
              *reinterpret_cast<float4*>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c0 + j;
              if (j >= ncol || n >= N) break;
              float x = acc[j];
              if (accumulate) x += crow[j];
              crow[j] = act_apply(x, act);
            }
          }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues with warp-uniform operands
    // (tc.cuh: elect_one — under `if (lane == 0)` every UTCHMMA sat inside a lane-broadcast loop, ~93 cycles per MMA)
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = idesc_bf16(TM, BN);
      const uint32_t lbo_a = TM * 16, lbo_b = BN * 16;
      const uint64_t d_a = smem_desc(smem_u32(smem), lbo_a, 128), d_b = smem_desc(smem_u32(smem) + 2 * a_half, lbo_b, 128);
      const bool p3 = passes == 3;
      uint32_t acc_flag = 0;
      RingPos rp;
      for (int ks = 0; ks < nks; ++ks, rp.next(stages)) {
        const uint32_t slot = rp.slot, ph = rp.ph;
        const uint64_t da = desc_adv(d_a, slot * (uint32_t)stage_bytes), db = desc_adv(d_b, slot * (uint32_t)stage_bytes);
        mbar_wait(&full[slot], ph);
        tc_fence_after();
        if (elect_one()) {
          if (p3) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              const uint64_t ahi = desc_adv(da, j * 2 * lbo_a), alo = desc_adv(da, a_half + j * 2 * lbo_a);
              const uint64_t bhi = desc_adv(db, j * 2 * lbo_b), blo = desc_adv(db, b_half + j * 2 * lbo_b);
              mma_bf16(tm, ahi, bhi, idesc, j == 0 ? acc_flag : 1u);
              mma_bf16(tm, alo, bhi, idesc, 1);
              mma_bf16(tm, ahi, blo, idesc, 1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j)
              mma_bf16(tm, desc_adv(da, j * 2 * lbo_a), desc_adv(db, j * 2 * lbo_b), idesc, j == 0 ? acc_flag : 1u);
          }
          mma_commit(&empty[slot]);     // frees the stage when these MMAs have read it
          if (ks == nks - 1) mma_commit(tfull);              // accumulator complete
        }
        acc_flag = 1;
      }
      __syncwarp();
    }
  } else {
    // ===================== B loader: one bulk TMA copy per stage
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(Bp) + ((size_t)jn * nks_total + ks_begin) * (2 * b_half);
      RingPos rp;
      for (int ks = 0; ks < nks; ++ks, rp.next(stages)) {
        const uint32_t slot = rp.slot, ph = rp.ph;
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], 2 * b_half);
        bulk_g2s(smem + (size_t)slot * stage_bytes + 2 * a_half, src + (size_t)ks * (2 * b_half), 2 * b_half, &full[slot]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, tmem_cols);
}

int pick_bn(int N) {
  if (N <= 256) return (N + 15) / 16 * 16;
  int best = 256, best_waste = 1 << 30;
  const int cands[] = {256, 224, 192, 160, 128};
  for (int bn : cands) {
    int waste = (N + bn - 1) / bn * bn - N;
    if (waste < best_waste) {
      best_waste = waste;
      best = bn;
    }
  }
  return best;
}

struct Plan {
  int BN, ntn, nks, stages;
  uint32_t tmem_cols;
  size_t smem, pack_bytes;
};
// M > 0 additionally picks how many CTAs share an SM (1..4, bounded by shared memory and by the 512 TMEM
// columns): the choice that wastes the fewest SM slots in the last wave wins (300 tiles on 2x148 slots would
// run two waves for four tiles), ties go to more CTAs per SM — the A producers keep one stage of loads in
// flight per CTA, so residency is what buys memory-level parallelism.
Plan make_plan(int N, int K, int M = 0) {
  Plan p;
  p.BN = pick_bn(N);
  p.ntn = (N + p.BN - 1) / p.BN;
  p.nks = (K + BK - 1) / BK;
  const size_t stage = 2 * (size_t)KC * TM * 16 + 2 * (size_t)KC * p.BN * 16;
  uint32_t cols = 32;
  while ((int)cols < p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const long tiles = (long)p.ntn * ((M + TM - 1) / TM);
  int best_c = 2, best_st = 2;
  double best_eff = -1.0;
  for (int c = 1; c <= 4; ++c) {
    if ((uint32_t)c * cols > 512) break;
    int st = (int)(((227 * 1024) / c - 2048) / stage);
    if (st < 2) break;
    if (c == 1 && (int)((227 * 1024 / 2 - 2048) / stage) >= 2 && 2 * cols <= 512) continue;   // never run 1 CTA/SM if 2 fit
    if (st > 6) st = 6;
    const long slots = 148L * c;
    const double eff = M > 0 ? (double)tiles / (double)(((tiles + slots - 1) / slots) * slots) : (c == 2 ? 1.0 : 0.0);
    if (eff > best_eff + 0.02) {      // ties keep the smaller c: deeper rings beat residency (measured)
      best_eff = eff;
      best_c = c;
      best_st = st;
    }
  }
  p.stages = best_st;
  if (p.stages > p.nks) p.stages = p.nks < 2 ? 2 : p.nks;
  p.smem = p.stages * stage + (2 * p.stages + 1) * sizeof(uint64_t) + 16;
  p.pack_bytes = (size_t)p.ntn * p.nks * 2 * KC * p.BN * 16;
  return p;
}

int pack_for_plan(const Plan& p, const float* W, int ldw, bool trans_b, int K, int N, void* pack_ws, cudaStream_t st) {
  const long total = (long)p.ntn * p.nks * KC * p.BN;
  pack_b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, ldw, trans_b ? 1 : 0, K, N, p.BN, p.nks, p.ntn,
                                                                 (uint4*)pack_ws);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Persistent, weight-stationary variant for short-K / huge-M products (the Decoder-2 input projection: M = R*T_f =
// 460 800 rows, K = 48, N = 3H = 384 at the bench workload, 708 MB of output per call).  With one output tile per CTA
// (kernel above) 7 200 CTAs each pay barrier set-up, TMEM allocation, the weight fetch and a prologue / main loop /
// epilogue that do not overlap: 0.62 ms per call where the HBM traffic needs 0.12.  Here
//   * a CTA per SM loads the WHOLE packed weight image once (<= 96 KB) and walks over M tiles (stride gridDim.x);
//   * warps 0-3 (thread = row) load + split the A rows of tile i+1 while tile i is multiplied and tile i-1 leaves;
//   * warp 4 issues the MMAs of one (tile, n-tile) into one of TWO accumulators (2 x 256 TMEM columns);
//   * warps 6-9 drain the other accumulator: bias / rank-2 term / activation, 128-byte-swizzled [128 x 32] boxes in
//     shared memory, TMA tensor-map stores (two boxes in flight).
// Requirements (gemm_persist_eligible): dense A, K % 8 == 0, K <= 64, N % 32 == 0, no accumulate, 16-byte aligned rows.
constexpr int PN_A = 2;                 // A tiles in shared memory
constexpr int PNTHR = 320;

__global__ void __launch_bounds__(PNTHR, 1) gemm_tc_persist_kernel(DenseA8 A_, const uint4* __restrict__ Bp,
                                                                   const float* __restrict__ bias, int M, int N, int BN,
                                                                   int ntn, int nks, int act, int passes, Rank2 r2,
                                                                   const __grid_constant__ CUtensorMap tm_c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_half = KC * TM * 16;                      // bytes of A_hi (== A_lo) per K stage
  const int a_tile = nks * 2 * a_half;                  // one A tile: nks stages of { hi, lo }
  const int b_half = KC * BN * 16;
  const int b_blk = 2 * b_half;                         // one packed (n-tile, K stage) block
  const int b_bytes = ntn * nks * b_blk;
  uint8_t* boxes = smem;                                // two [128 x 32] FP32 output boxes (1024-byte aligned)
  uint8_t* a_s = smem + 2 * TM * 128;
  uint8_t* b_s = a_s + PN_A * a_tile;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_s + b_bytes);
  uint64_t* a_empty = a_full + PN_A;
  uint64_t* acc_full = a_empty + PN_A;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtiles = (M + TM - 1) / TM;
  const int my_tiles = ((int)blockIdx.x < mtiles) ? (mtiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < PN_A; ++s) {
      mbar_init(&a_full[s], 4);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<512>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < 4) {
    // ===================== A producers: thread <-> row; the rows of the next tile are in registers while this one is written
    DenseA8 A = A_;
    float v[2][KC][8];
    auto load = [&](int i) {
      A.init(((int)blockIdx.x + i * (int)gridDim.x) * TM + tid);
      A.load_stage(0, v[0]);
      if (nks > 1) A.load_stage(1, v[1]);
    };
    if (my_tiles > 0) load(0);
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i % PN_A;
      mbar_wait(&a_empty[buf], ((i / PN_A) & 1) ^ 1);
      uint8_t* sa = a_s + (size_t)buf * a_tile;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        if (ks < nks) {
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            const Split8 sp = split8(v[ks][c]);
            *reinterpret_cast<uint4*>(sa + ks * 2 * a_half + c * TM * 16 + tid * 16) = sp.hi;
            *reinterpret_cast<uint4*>(sa + ks * 2 * a_half + a_half + c * TM * 16 + tid * 16) = sp.lo;
          }
        }
      }
      if (i + 1 < my_tiles) load(i + 1);                 // in flight while the other roles work on tile i
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (whole warp, elected lane, uniform descriptors)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = idesc_bf16(TM, BN);
    const uint32_t lbo_a = TM * 16, lbo_b = BN * 16;
    const uint64_t d_a = smem_desc(smem_u32(a_s), lbo_a, 128), d_b = smem_desc(smem_u32(b_s), lbo_b, 128);
    const int ksteps = (A_.K + 15) / 16;                 // 16-wide K steps that hold data
    const bool p3 = passes == 3;
    mbar_wait(b_full, 0);
    uint32_t use = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i % PN_A;
      mbar_wait(&a_full[buf], (i / PN_A) & 1);
      tc_fence_after();
      for (int jn = 0; jn < ntn; ++jn, ++use) {
        const int slot = use & 1;
        mbar_wait(&acc_empty[slot], ((use >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tm + slot * 256;
        const uint64_t da = desc_adv(d_a, buf * (uint32_t)a_tile), db = desc_adv(d_b, (uint32_t)(jn * nks) * (uint32_t)b_blk);
        if (elect_one()) {
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint32_t ao = (kk >> 1) * 2 * a_half + (kk & 1) * 2 * lbo_a;
            const uint32_t bo = (kk >> 1) * b_blk + (kk & 1) * 2 * lbo_b;
            mma_bf16(d, desc_adv(da, ao), desc_adv(db, bo), idesc, kk > 0);
            if (p3) {
              mma_bf16(d, desc_adv(da, ao + a_half), desc_adv(db, bo), idesc, 1);
              mma_bf16(d, desc_adv(da, ao), desc_adv(db, bo + b_half), idesc, 1);
            }
          }
          mma_commit(&acc_full[slot]);
          if (jn == ntn - 1) mma_commit(&a_empty[buf]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== weight loader: the whole packed image, once
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, (uint32_t)b_bytes);
      for (int o = 0; o < b_bytes; o += b_blk)
        bulk_g2s(b_s + o, reinterpret_cast<const uint8_t*>(Bp) + o, (uint32_t)b_blk, b_full);
    }
  } else {
    // ===================== epilogue warps 6-9: TMEM lanes 32*(warp%4) .. +31 == rows m0 + et
    const int q = warp & 3, et = q * 32 + lane;          // row of the tile; also this thread's index among the 128
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const bool leader = warp == 6 && lane == 0;
    const uint32_t obox = smem_u32(boxes);
    uint32_t use = 0;
    int ci = 0;                                          // boxes written so far (two alternate)
    for (int i = 0; i < my_tiles; ++i) {
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * TM;
      const int m = m0 + et;
      float s0 = 0.f, s1 = 0.f;
      const float* pr = nullptr;
      if (r2.P && m < M) {
        s0 = __ldg(r2.s + 2 * (size_t)m);
        s1 = __ldg(r2.s + 2 * (size_t)m + 1);
        pr = r2.P + (size_t)(m / r2.div) * 2 * N;
      }
      for (int jn = 0; jn < ntn; ++jn, ++use) {
        const int slot = use & 1;
        const int n0 = jn * BN;
        mbar_wait(&acc_full[slot], (use >> 1) & 1);
        tc_fence_after();
        for (int c0 = 0; c0 < BN && n0 + c0 < N; c0 += 32, ++ci) {
          float acc[32];
          tmem_ld32(trow + slot * 256 + c0, acc);
          tmem_ld_wait();
          if (c0 + 32 >= BN || n0 + c0 + 32 >= N) {      // last chunk of the accumulator: the MMAs may refill it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[slot]);
          }
          chunk_epilogue(acc, bias ? bias + n0 + c0 : nullptr, pr ? pr + n0 + c0 : nullptr, N, s0, s1, act);
          const uint32_t box = obox + (ci & 1) * (TM * 128);
          if (ci >= 2) {                                 // the store that last read this box has finished reading it
            if (leader) tma_store_wait_read1();
            asm volatile("bar.sync 2, 128;" ::: "memory");
          }
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16)
            sts128(swz128_abs(box, et, c16), make_float4(acc[4 * c16], acc[4 * c16 + 1], acc[4 * c16 + 2], acc[4 * c16 + 3]));
          fence_proxy_async();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (leader) {
            tma_store_2d(&tm_c, boxes + (ci & 1) * (TM * 128), n0 + c0, m0);
            tma_store_commit();
          }
        }
      }
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

size_t persist_smem(const Plan& p) {
  return 2 * (size_t)TM * 128 + (size_t)PN_A * p.nks * 2 * KC * TM * 16 + p.pack_bytes + (2 * PN_A + 5) * 8 + 16;
}
// Takes the problem when it is the short-K / huge-M kind this kernel is for; *rc receives the launch status.
bool run_tc_persist(const DenseA8& A, const void* packed, const float* bias, float* C, int ldc, int M, int N, int K,
                    int act, cudaStream_t st, const Rank2& r2, int* rc) {
  static const bool off = [] {
    const char* e = getenv("DESIRE_GEMM_NO_PERSIST");
    return e && e[0] == '1';
  }();
  const Plan p = make_plan(N, K, M);
  if (off || K > 2 * BK || K % 8 != 0 || !A.vec || N % 32 != 0 || p.BN % 32 != 0 || p.BN > 256 || M < 148 * 4 * TM ||
      ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(C) & 15) != 0 || (reinterpret_cast<uintptr_t>(bias) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(r2.P) & 15) != 0 || persist_smem(p) > 227 * 1024)
    return false;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (make_tmap_rows32(&tm, C, M, ldc) != DESIRE_OK) return false;
  const size_t smem = persist_smem(p);
  auto launch = [&]() -> int {
    DESIRE_ENSURE_SMEM(gemm_tc_persist_kernel, smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int mtiles = (M + TM - 1) / TM;
    const unsigned grid = (unsigned)std::min(sms, mtiles);
    DESIRE_LAUNCH(st, (gemm_tc_persist_kernel<<<grid, PNTHR, smem, st>>>(A, (const uint4*)packed, bias, M, N, p.BN, p.ntn, p.nks,
                                                                         act, g_gemm_mode == 1 ? 1 : 3, r2, tm)));
    return DESIRE_OK;
  };
  *rc = launch();
  return true;
}

template <class ALoad>
int run_tc(const ALoad& A, const void* packed, const float* bias, float* C, int ldc, int M, int N, int K, int act,
           bool accumulate, cudaStream_t st, Rank2 r2 = Rank2()) {
  const Plan p = make_plan(N, K, M);
  DESIRE_ENSURE_SMEM(gemm_tc_kernel<ALoad>, p.smem);
  dim3 grid(p.ntn, (M + TM - 1) / TM);
  // tall outputs made of whole 32-column boxes leave through TMA stores (see the kernel's epilogue)
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int tma_c = 0;
  static const bool no_tma = [] {
    const char* e = getenv("DESIRE_GEMM_NO_TMA_STORE");
    return e && e[0] == '1';
  }();
  if (!no_tma && !accumulate && M >= 4096 && N % 32 == 0 && p.BN % 32 == 0 && ldc % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(C) & 15) == 0 && make_tmap_rows32(&tm, C, M, ldc) == DESIRE_OK)
    tma_c = 1;
  DESIRE_LAUNCH(st, (gemm_tc_kernel<ALoad><<<grid, NTHR, p.smem, st>>>(A, (const uint4*)packed, bias, C, ldc, M, N, p.BN,
                                                                       p.nks, p.stages, act, accumulate ? 1 : 0,
                                                                       g_gemm_mode == 1 ? 1 : 3, p.tmem_cols, 0, r2, tm,
                                                                       tma_c)));
  return DESIRE_OK;
}

template <class ALoad>
int launch_tc(const ALoad& A, const float* W, int ldw, bool trans_b, const float* bias, float* C, int ldc, int M, int N,
              int K, int act, bool accumulate, void* pack_ws, cudaStream_t st) {
  DESIRE_TRY(pack_for_plan(make_plan(N, K), W, ldw, trans_b, K, N, pack_ws, st));
  return run_tc(A, pack_ws, bias, C, ldc, M, N, K, act, accumulate, st);
}


template <class ALoad>
int run_wgrad_tc(const ALoad& A, const float* B, int ldb, float* dW, int ldw, int rows, int Kd, int N, void* pack_ws,
                 cudaStream_t st) {
  // GEMM view: M = Kd, K = rows; B [rows,N] is packed like a weight matrix (its K dimension is the sample rows)
  const Plan p = make_plan(N, rows, Kd);
  DESIRE_TRY(pack_for_plan(p, B, ldb, false, rows, N, pack_ws, st));
  const int gy = (Kd + TM - 1) / TM;
  const int tiles = p.ntn * gy;
  int splits = (2 * 148 + tiles - 1) / tiles;
  const int max_splits = p.nks / 8 > 0 ? p.nks / 8 : 1;       // at least 8 K stages (256 rows) per slice
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int ks_split = (p.nks + splits - 1) / splits;
  splits = (p.nks + ks_split - 1) / ks_split;
  int stages = p.stages;
  if (stages > ks_split) stages = ks_split < 2 ? 2 : ks_split;
  const size_t stage = 2 * (size_t)KC * TM * 16 + 2 * (size_t)KC * p.BN * 16;
  const size_t smem = stages * stage + (2 * stages + 1) * sizeof(uint64_t) + 16;
  DESIRE_ENSURE_SMEM(gemm_tc_kernel<ALoad>, smem);
  dim3 grid(p.ntn, gy, splits);
  DESIRE_LAUNCH(st, (gemm_tc_kernel<ALoad><<<grid, NTHR, smem, st>>>(A, (const uint4*)pack_ws, nullptr, dW, ldw, Kd, N, p.BN,
                                                                     p.nks, stages, DESIRE_ACT_NONE, 2,
                                                                     g_gemm_mode == 1 ? 1 : 3, p.tmem_cols, ks_split, Rank2(),
                                                                     CUtensorMap(), 0)));
  return DESIRE_OK;
}

}  // namespace

size_t gemm_tc_pack_bytes(int N, int K) { return align_up(make_plan(N, K).pack_bytes); }

// ---- weight gradients on tcgen05: dW[Kd,N] += A^T @ B over `rows` sample rows (A dense or implicit im2col)
size_t wgrad_tc_pack_bytes(int rows, int N) { return align_up(make_plan(N, rows).pack_bytes); }
bool wgrad_tc_eligible(int rows, int Kd, int N, const void* pack_ws, size_t pack_bytes) {
  return g_gemm_mode != 0 && Kd >= 64 && N >= 16 && rows >= 2048 && pack_ws &&
         pack_bytes >= make_plan(N, rows).pack_bytes && (Kd + TM - 1) / TM <= 65535;
}
int wgrad_tc(const float* A, int lda, const float* B, int ldb, float* dW, int ldw, int rows, int Kd, int N, void* pack_ws,
             cudaStream_t st) {
  TransDenseA8 a{A, lda, Kd, rows, nullptr};
  return run_wgrad_tc(a, B, ldb, dW, ldw, rows, Kd, N, pack_ws, st);
}
int wgrad_tc_im2col(const float* X, const Im2col& g, const float* B, int ldb, float* dW, int ldw, int rows, int Kd, int N,
                    void* pack_ws, cudaStream_t st) {
  TransIm2colA8 a{X, g, Kd, rows, 0, 0, -1};
  return run_wgrad_tc(a, B, ldb, dW, ldw, rows, Kd, N, pack_ws, st);
}


// Packed BF16 hi/lo image of W for an explicit n-tile width BN (used by the GRU kernel): per
// (n-tile, 32-wide k stage) one contiguous block { hi [4][BN][8], lo [4][BN][8] }.
size_t tc_pack_bytes(int K, int N, int BN) {
  return (size_t)((N + BN - 1) / BN) * ((K + BK - 1) / BK) * 2 * KC * BN * 16;
}
int tc_pack_b(const float* W, int ldw, bool trans, int K, int N, int BN, void* out, cudaStream_t st) {
  const int ntn = (N + BN - 1) / BN, nks = (K + BK - 1) / BK;
  const long total = (long)ntn * nks * KC * BN;
  pack_b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, ldw, trans ? 1 : 0, K, N, BN, nks, ntn, (uint4*)out);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

// Returns true when the tensor-core path took the problem.
bool gemm_tc_eligible(int M, int N, int K, const void* pack_ws, size_t pack_bytes) {
  return g_gemm_mode != 0 && M >= 64 && N >= 16 && K >= 8 && pack_ws && pack_bytes >= make_plan(N, K).pack_bytes &&
         (M + TM - 1) / TM <= 65535;
}

int gemm_tc(const float* A, int lda, const float* W, int ldw, bool trans_b, const float* bias, float* C, int ldc, int M,
            int N, int K, int act, bool accumulate, void* pack_ws, cudaStream_t st) {
  DenseA8 a{A, lda, M, K, (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0), nullptr};
  if (!accumulate) {                                             // short-K / huge-M: the persistent, weight-stationary kernel
    const Plan p = make_plan(N, K);
    if (K <= 2 * BK && M >= 148 * 4 * TM) {
      DESIRE_TRY(pack_for_plan(p, W, ldw, trans_b, K, N, pack_ws, st));
      int rc = DESIRE_OK;
      if (run_tc_persist(a, pack_ws, bias, C, ldc, M, N, K, act, st, Rank2(), &rc)) return rc;
      return run_tc(a, pack_ws, bias, C, ldc, M, N, K, act, false, st);
    }
  }
  return launch_tc(a, W, ldw, trans_b, bias, C, ldc, M, N, K, act, accumulate, pack_ws, st);
}

int pack_weight(PackedW& w, void* ws, size_t ws_bytes, cudaStream_t st) {
  w.packed = nullptr;
  if (g_gemm_mode == 0 || w.N < 16 || w.K < 8 || !ws || ws_bytes < make_plan(w.N, w.K).pack_bytes) return DESIRE_OK;
  DESIRE_TRY(pack_for_plan(make_plan(w.N, w.K), w.W, w.ldw, w.trans, w.K, w.N, ws, st));
  w.packed = ws;
  return DESIRE_OK;
}

int gemm_packed(const float* A, int lda, const PackedW& w, const float* bias, float* C, int ldc, int M, int act,
                bool accumulate, cudaStream_t st) {
  if (w.packed && g_gemm_mode != 0 && M >= 64 && (M + TM - 1) / TM <= 65535) {
    DenseA8 a{A, lda, M, w.K, (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0), nullptr};
    int rc = DESIRE_OK;
    if (!accumulate && run_tc_persist(a, w.packed, bias, C, ldc, M, w.N, w.K, act, st, Rank2(), &rc)) return rc;
    return run_tc(a, w.packed, bias, C, ldc, M, w.N, w.K, act, accumulate, st);
  }
  return sgemm(A, lda, w.W, w.ldw, w.trans, bias, C, ldc, M, w.N, w.K, act, accumulate, st);
}

// C = A @ W + bias + s[:,0] * P[g,0,:] + s[:,1] * P[g,1,:] on the tensor-core path (W packed); false = not eligible
bool gemm_packed_r2(const float* A, int lda, const PackedW& w, const float* bias, float* C, int ldc, int M, int act,
                    const Rank2& r2, cudaStream_t st, int* rc) {
  const bool ok = w.packed && g_gemm_mode != 0 && M >= 64 && (M + TM - 1) / TM <= 65535 && r2.P && r2.s && r2.div > 0 &&
                  (reinterpret_cast<uintptr_t>(r2.P) & 15) == 0 && w.N % 4 == 0;
  if (!ok) return false;
  DenseA8 a{A, lda, M, w.K, (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0), nullptr};
  if (run_tc_persist(a, w.packed, bias, C, ldc, M, w.N, w.K, act, st, r2, rc)) return true;
  *rc = run_tc(a, w.packed, bias, C, ldc, M, w.N, w.K, act, false, st, r2);
  return true;
}

// C = [A1 | A2] @ W + bias with W already packed (K = K1 + K2).  Requirements: tensor-core path available, K1 % 8 == 0,
// lda1/lda2 % 4 == 0 and 16-byte aligned sources (returns false otherwise and the caller concatenates).
bool gemm_packed_dual(const float* A1, int lda1, int K1, const float* A2, int lda2, const PackedW& w, const float* bias,
                      float* C, int ldc, int M, int act, cudaStream_t st, int* rc) {
  const bool ok = w.packed && g_gemm_mode != 0 && M >= 64 && (M + TM - 1) / TM <= 65535 && K1 % 8 == 0 && lda1 % 4 == 0 &&
                  lda2 % 4 == 0 && (w.K - K1) % 4 == 0 && ((reinterpret_cast<uintptr_t>(A1) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(A2) & 15) == 0) && K1 > 0 && K1 < w.K;
  if (!ok) return false;
  DualDenseA8 a{A1, A2, lda1, lda2, K1, M, w.K, nullptr, nullptr};
  *rc = run_tc(a, w.packed, bias, C, ldc, M, w.N, w.K, act, false, st);
  return true;
}

int gemm_tc_im2col(const float* X, const Im2col& g, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                   int N, int K, int act, void* pack_ws, cudaStream_t st) {
  Im2colA8 a{X, g, M, K, nullptr, 0, 0};
  return launch_tc(a, W, ldw, false, bias, C, ldc, M, N, K, act, false, pack_ws, st);
}

}  // namespace desire

extern "C" int desire_set_gemm_mode(int mode) {
  DESIRE_CHECK_ARG(mode == 0 || mode == 1 || mode == 3, "desire_set_gemm_mode: mode must be 0 (fp32), 1 (bf16) or 3 (3xbf16)");
  desire::g_gemm_mode = mode;
  return DESIRE_OK;
}
extern "C" int desire_get_gemm_mode(void) { return desire::g_gemm_mode; }

extern "C" size_t desire_gemm_tc_workspace_bytes(int N, int K) { return desire::gemm_tc_pack_bytes(N, K); }

extern "C" int desire_gemm_tc_fwd(const float* A, int lda, const float* W, int ldw, int trans_w, const float* bias,
                                  float* C, int ldc, int M, int N, int K, int act, int accumulate, void* ws,
                                  size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0, "desire_gemm_tc_fwd: bad arguments");
  DESIRE_CHECK_ARG(lda >= K && ldc >= N && ldw >= (trans_w ? K : N), "desire_gemm_tc_fwd: leading dimensions too small");
  if (!ws || ws_bytes < desire::gemm_tc_pack_bytes(N, K)) {
    desire::set_error("desire_gemm_tc_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  DESIRE_CHECK_ARG((M + 127) / 128 <= 65535, "desire_gemm_tc_fwd: M too large");
  return desire::gemm_tc(A, lda, W, ldw, trans_w != 0, bias, C, ldc, M, N, K, act, accumulate != 0, ws,
                         (cudaStream_t)stream);
}
