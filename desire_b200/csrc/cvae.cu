// Conv-CVAE of the sample-generation stage (model/model.py:453-492, utils/convolutional_vae_util.py).
//
// Every conv / deconv layer is a GEMM plus one fused "col2im + bias + per-row BN + activation"
// kernel.  Forward convs gather (im2col loader inside the GEMM, k=1 identity col2im);
// transposed convs run in SCATTER form: col = X[R*Pin, Cin] @ W^T[Cin, k*k*Cout] does exactly the
// algorithmic MACs (no stride-2 zero taps), and the col2im kernel gathers each col element once.
// BN is the reference's batch-of-one statistics: per row and channel over the H*W positions
// (DESIGN.md D5), two-pass variance, eps = 1e-3.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

constexpr float BN_EPS = 1e-3f;

// one CTA per row; out tile [Hout*Hout*Cout] lives in shared memory between the passes
__global__ void __launch_bounds__(256) colbn_act_kernel(const float* __restrict__ col, int Hin, int Hout, int k,
                                                        int stride, int pad, int Cout,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int act,
                                                        float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int Pout = Hout * Hout, n = Pout * Cout;
  float* tile = sm;            // [n]
  float* red = sm + n;         // [256]
  float* stat = red + 256;     // [2*Cout] mean, rstd
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t r = blockIdx.x;
  const int kk = k * k;
  const float* colr = col + r * (size_t)Hin * Hin * kk * Cout;

  for (int e = tid; e < n; e += nthr) {
    int o = e % Cout, p = e / Cout;
    int oy = p / Hout, ox = p % Hout;
    float acc = bias ? __ldg(bias + o) : 0.f;
    for (int ky = 0; ky < k; ++ky) {
      int ty = oy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      int iy = ty / stride;
      if (iy >= Hin) continue;
      for (int kx = 0; kx < k; ++kx) {
        int tx = ox + pad - kx;
        if (tx < 0 || tx % stride) continue;
        int ix = tx / stride;
        if (ix >= Hin) continue;
        acc += __ldg(colr + ((size_t)(iy * Hin + ix) * kk + ky * k + kx) * Cout + o);
      }
    }
    tile[e] = acc;
  }
  __syncthreads();

  // per-channel moments over the Pout positions: thread (o, part) strides positions by `parts`
  const int parts = nthr / Cout;          // Cout in {1,32,64,128} <= nthr
  const int o = tid % Cout, part = tid / Cout;
  const bool on = part < parts;
  float s = 0.f;
  if (on)
    for (int p = part; p < Pout; p += parts) s += tile[p * Cout + o];
  red[tid] = on ? s : 0.f;
  __syncthreads();
  if (tid < Cout) {
    float t = 0.f;
    for (int q = 0; q < parts; ++q) t += red[q * Cout + tid];
    stat[tid] = t / (float)Pout;
  }
  __syncthreads();
  const float mean = stat[o];
  s = 0.f;
  if (on)
    for (int p = part; p < Pout; p += parts) {
      float d = tile[p * Cout + o] - mean;
      s += d * d;
    }
  __syncthreads();
  red[tid] = on ? s : 0.f;
  __syncthreads();
  if (tid < Cout) {
    float t = 0.f;
    for (int q = 0; q < parts; ++q) t += red[q * Cout + tid];
    stat[Cout + tid] = 1.f / sqrtf(t / (float)Pout + BN_EPS);
  }
  __syncthreads();
  float* outr = out + r * (size_t)n;
  for (int e = tid; e < n; e += nthr) {
    int oc = e % Cout;
    float v = __ldg(gamma + oc) * ((tile[e] - stat[oc]) * stat[Cout + oc]) + __ldg(beta + oc);
    outr[e] = act_apply(v, act);
  }
}

// Vectorised variant (Cout % 4 == 0, compile-time kernel size / stride): every output float4 gathers its
// valid taps with independent 128-bit loads (the tap loops unroll, so up to KS*KS loads are in flight per
// thread and each col element is still read exactly once), then the same two-pass BN over the smem tile.
template <int KS, int S>
__global__ void __launch_bounds__(256) colbn_act_v4_kernel(const float* __restrict__ col, int Hin, int Hout, int pad,
                                                           int Cout, const float* __restrict__ bias,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int act,
                                                           float* __restrict__ out, float* __restrict__ ypre) {
  extern __shared__ __align__(16) float sm[];
  const int Pout = Hout * Hout, n = Pout * Cout, C4 = Cout / 4;
  float* tile = sm;
  float* red = sm + n;
  float* stat = red + 256;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t r = blockIdx.x;
  constexpr int KK = KS * KS;
  const float* colr = col + r * (size_t)Hin * Hin * KK * Cout;

  for (int e = tid; e < Pout * C4; e += nthr) {
    const int o4 = e % C4, p = e / C4;
    const int oy = p / Hout, ox = p % Hout;
    float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + o4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const int ty = oy + pad - ky;
      const int iy = ty / S;
      const bool vy = ty >= 0 && (ty % S) == 0 && iy < Hin;
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const int tx = ox + pad - kx;
        const int ix = tx / S;
        if (vy && tx >= 0 && (tx % S) == 0 && ix < Hin) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(colr + ((size_t)(iy * Hin + ix) * KK + ky * KS + kx) * Cout) + o4);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    reinterpret_cast<float4*>(tile)[e] = acc;
    if (ypre) reinterpret_cast<float4*>(ypre + r * (size_t)n)[e] = acc;   // train step: pre-BN values for the backward
  }
  __syncthreads();

  const int parts = nthr / Cout;
  const int o = tid % Cout, part = tid / Cout;
  const bool on = part < parts;
  float s = 0.f;
  if (on)
    for (int p = part; p < Pout; p += parts) s += tile[p * Cout + o];
  red[tid] = on ? s : 0.f;
  __syncthreads();
  if (tid < Cout) {
    float t = 0.f;
    for (int q = 0; q < parts; ++q) t += red[q * Cout + tid];
    stat[tid] = t / (float)Pout;
  }
  __syncthreads();
  const float mean = stat[o];
  s = 0.f;
  if (on)
    for (int p = part; p < Pout; p += parts) {
      const float d = tile[p * Cout + o] - mean;
      s += d * d;
    }
  __syncthreads();
  red[tid] = on ? s : 0.f;
  __syncthreads();
  if (tid < Cout) {
    float t = 0.f;
    for (int q = 0; q < parts; ++q) t += red[q * Cout + tid];
    stat[Cout + tid] = 1.f / sqrtf(t / (float)Pout + BN_EPS);
  }
  __syncthreads();
  float4* outr = reinterpret_cast<float4*>(out + r * (size_t)n);
  for (int e = tid; e < Pout * C4; e += nthr) {
    const int oc = (e % C4) * 4;
    const float4 x = reinterpret_cast<const float4*>(tile)[e];
    // all reads first, branch-free activation (tc.cuh: act_fast): a data-dependent branch per element (act_apply's ELU)
    // keeps the next element's loads behind it
    const float g0 = __ldg(gamma + oc), g1 = __ldg(gamma + oc + 1), g2 = __ldg(gamma + oc + 2), g3 = __ldg(gamma + oc + 3);
    const float b0 = __ldg(beta + oc), b1 = __ldg(beta + oc + 1), b2 = __ldg(beta + oc + 2), b3 = __ldg(beta + oc + 3);
    const float m0 = stat[oc], m1 = stat[oc + 1], m2 = stat[oc + 2], m3 = stat[oc + 3];
    const float r0 = stat[Cout + oc], r1 = stat[Cout + oc + 1], r2 = stat[Cout + oc + 2], r3 = stat[Cout + oc + 3];
    float4 y;
    y.x = tc::act_fast(g0 * ((x.x - m0) * r0) + b0, act);
    y.y = tc::act_fast(g1 * ((x.y - m1) * r1) + b1, act);
    y.z = tc::act_fast(g2 * ((x.z - m2) * r2) + b2, act);
    y.w = tc::act_fast(g3 * ((x.w - m3) * r3) + b3, act);
    outr[e] = y;
  }
}

// Direct single-output-channel transposed conv (the decoder's last layer, model/model.py:468:
// 16x16x32 -> 32x32x1, k5 s2 SAME, BN + sigmoid).  Its GEMM form is degenerate (N = 25 columns, a col matrix
// 8x larger than the input), so: one CTA per sample, input tile [Pin][Cin+4] and the filter in shared memory
// (row stride Cin+4 floats keeps the 128-bit reads of a quarter-warp conflict-free), each thread gathers its
// output pixels over the valid taps with float4 FMAs, then block-wide two-pass BN and the activation.
template <int CIN, int KS, int STRIDE>
__global__ void __launch_bounds__(256) deconv1ch_bn_act_kernel(const float* __restrict__ X, int Hin, int Hout,
                                                               int pad, const float* __restrict__ W,
                                                               const float* __restrict__ bias,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int act,
                                                               float* __restrict__ Y) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LD = CIN + 4;
  const int Pin = Hin * Hin, Pout = Hout * Hout;
  constexpr int kk = KS * KS;
  float* xin = sm;                       // [Pin][LD]
  float* wf = xin + (size_t)Pin * LD;    // [kk][CIN]
  float* red = wf + kk * CIN;            // [32]
  const int tid = threadIdx.x;
  const size_t r = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(X + r * (size_t)Pin * CIN);
  for (int e = tid; e < Pin * (CIN / 4); e += 256) {
    const int p = e / (CIN / 4), c4 = e % (CIN / 4);
    *reinterpret_cast<float4*>(xin + (size_t)p * LD + c4 * 4) = __ldg(src + e);
  }
  for (int e = tid; e < kk * CIN; e += 256) wf[e] = __ldg(W + e);      // [ky,kx,o=0,ci] == [tap][ci]
  __syncthreads();

  float vals[4];                          // Pout <= 1024 => <= 4 pixels per thread (pixel = tid + 256*i, coalesced)
  const float b0 = __ldg(bias);
  float lsum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = tid + 256 * i;
    float acc = 0.f;
    if (q < Pout) {
      const int oy = q / Hout, ox = q - oy * Hout;
      float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
      // only the taps of this pixel's parity class exist: ky = (oy+pad) mod S, +S, ... (no div/mod per tap)
      for (int ky = (oy + pad) % STRIDE; ky < KS; ky += STRIDE) {
        const int iy = (oy + pad - ky) / STRIDE;
        if (iy < 0 || iy >= Hin) continue;
        for (int kx = (ox + pad) % STRIDE; kx < KS; kx += STRIDE) {
          const int ix = (ox + pad - kx) / STRIDE;
          if (ix < 0 || ix >= Hin) continue;
          const float4* xp = reinterpret_cast<const float4*>(xin + (size_t)(iy * Hin + ix) * LD);
          const float4* wp = reinterpret_cast<const float4*>(wf + (ky * KS + kx) * CIN);
#pragma unroll
          for (int c = 0; c < CIN / 4; ++c) {
            const float4 xv = xp[c], wv = wp[c];
            a4.x = fmaf(xv.x, wv.x, a4.x); a4.y = fmaf(xv.y, wv.y, a4.y);
            a4.z = fmaf(xv.z, wv.z, a4.z); a4.w = fmaf(xv.w, wv.w, a4.w);
          }
        }
      }
      acc = (a4.x + a4.y) + (a4.z + a4.w) + b0;
      lsum += acc;
    }
    vals[i] = acc;
  }
  // block reductions (deterministic): warp shuffle, then 8 partials
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
  };
  const float mean = block_sum(lsum) / (float)Pout;
  float lvar = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (tid + 256 * i < Pout) {
      const float d = vals[i] - mean;
      lvar += d * d;
    }
  const float rstd = 1.f / sqrtf(block_sum(lvar) / (float)Pout + BN_EPS);
  const float g = __ldg(gamma), be = __ldg(beta);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = tid + 256 * i;
    if (q < Pout) Y[r * (size_t)Pout + q] = act_apply(g * ((vals[i] - mean) * rstd) + be, act);
  }
}

// The decoder's first layer (1x1xZ -> 4x4xC, k4 VALID: every output position is exactly one tap, col == pre-BN output):
// bias + per-(sample, channel) BN over the 16 positions + activation with NO shared memory and no barrier.  Thread
// (channel quad q = tid / 8, position pair p = tid % 8) holds positions p and p + 8 of its four channels; the statistics
// are three xor-shuffles inside the 8-lane group.  One CTA = 256 threads = 32 quads = 128 channels of one sample; the
// general kernel above spent 0.48 ms on this layer at the bench workload (38 400 CTAs x 4 block-wide barriers).
__global__ void __launch_bounds__(256) bn_act_p16_kernel(const float* __restrict__ col, int Cout, const float* __restrict__ bias,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         int act, float* __restrict__ out, float* __restrict__ ypre) {
  const int tid = threadIdx.x;
  const int q_raw = (int)blockIdx.y * 32 + (tid >> 3), p = tid & 7;  // channel quad, position pair
  const bool on = q_raw * 4 < Cout;                                  // (no early exit: the shuffles below are warp-wide)
  const int q = on ? q_raw : 0;
  const size_t base = (size_t)blockIdx.x * 16 * Cout;
  const float4* c4 = reinterpret_cast<const float4*>(col + base);
  const int C4 = Cout / 4;
  float4 x0 = __ldcs(c4 + (size_t)p * C4 + q), x1 = __ldcs(c4 + (size_t)(p + 8) * C4 + q);
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q);
    x0.x += b.x; x0.y += b.y; x0.z += b.z; x0.w += b.w;
    x1.x += b.x; x1.y += b.y; x1.z += b.z; x1.w += b.w;
  }
  if (ypre && on) {
    reinterpret_cast<float4*>(ypre + base)[(size_t)p * C4 + q] = x0;
    reinterpret_cast<float4*>(ypre + base)[(size_t)(p + 8) * C4 + q] = x1;
  }
  auto sum8 = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
  };
  const float inv = 1.f / 16.f;
  const float4 mean = make_float4(sum8(x0.x + x1.x) * inv, sum8(x0.y + x1.y) * inv, sum8(x0.z + x1.z) * inv, sum8(x0.w + x1.w) * inv);
  const float4 d0 = make_float4(x0.x - mean.x, x0.y - mean.y, x0.z - mean.z, x0.w - mean.w);
  const float4 d1 = make_float4(x1.x - mean.x, x1.y - mean.y, x1.z - mean.z, x1.w - mean.w);
  const float4 rstd = make_float4(1.f / sqrtf(sum8(d0.x * d0.x + d1.x * d1.x) * inv + BN_EPS), 1.f / sqrtf(sum8(d0.y * d0.y + d1.y * d1.y) * inv + BN_EPS),
                                  1.f / sqrtf(sum8(d0.z * d0.z + d1.z * d1.z) * inv + BN_EPS), 1.f / sqrtf(sum8(d0.w * d0.w + d1.w * d1.w) * inv + BN_EPS));
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q), be = __ldg(reinterpret_cast<const float4*>(beta) + q);
  float4 y0, y1;
  y0.x = tc::act_fast(g.x * (d0.x * rstd.x) + be.x, act); y0.y = tc::act_fast(g.y * (d0.y * rstd.y) + be.y, act);
  y0.z = tc::act_fast(g.z * (d0.z * rstd.z) + be.z, act); y0.w = tc::act_fast(g.w * (d0.w * rstd.w) + be.w, act);
  y1.x = tc::act_fast(g.x * (d1.x * rstd.x) + be.x, act); y1.y = tc::act_fast(g.y * (d1.y * rstd.y) + be.y, act);
  y1.z = tc::act_fast(g.z * (d1.z * rstd.z) + be.z, act); y1.w = tc::act_fast(g.w * (d1.w * rstd.w) + be.w, act);
  if (on) {
    float4* o4 = reinterpret_cast<float4*>(out + base);
    o4[(size_t)p * C4 + q] = y0;
    o4[(size_t)(p + 8) * C4 + q] = y1;
  }
}

template <int KS, int S>
int launch_v4(const float* col, int R, int Hin, int Hout, int pad, int Cout, const float* bias, const float* gamma,
              const float* beta, int act, float* out, size_t smem, cudaStream_t st, float* ypre) {
  DESIRE_ENSURE_SMEM((colbn_act_v4_kernel<KS, S>), 227 * 1024);
  DESIRE_LAUNCH(st, (colbn_act_v4_kernel<KS, S><<<R, 256, smem, st>>>(col, Hin, Hout, pad, Cout, bias, gamma, beta, act, out, ypre)));
  return DESIRE_OK;
}

}  // namespace

int colbn_act(const float* col, int R, int Hin, int Hout, int k, int stride, int pad, int Cout, const float* bias,
              const float* gamma, const float* beta, int act, float* out, cudaStream_t st, float* ypre) {
  if (R == 0) return DESIRE_OK;
  if (Hin == 1 && Hout == 4 && k == 4 && stride == 1 && pad == 0 && Cout % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(col) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
        reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(ypre)) & 15) == 0) {
    const dim3 grid((unsigned)R, (unsigned)((Cout / 4 + 31) / 32));
    DESIRE_LAUNCH(st, (bn_act_p16_kernel<<<grid, 256, 0, st>>>(col, Cout, bias, gamma, beta, act, out, ypre)));
    return DESIRE_OK;
  }
  if (Cout % 4 == 0 && Cout <= 256 && 256 % Cout == 0) {
    const size_t smem4 = ((size_t)Hout * Hout * Cout + 256 + 2 * Cout) * sizeof(float);
    if (smem4 <= 227 * 1024) {
      if (k == 1 && stride == 1) return launch_v4<1, 1>(col, R, Hin, Hout, pad, Cout, bias, gamma, beta, act, out, smem4, st, ypre);
      if (k == 4 && stride == 1) return launch_v4<4, 1>(col, R, Hin, Hout, pad, Cout, bias, gamma, beta, act, out, smem4, st, ypre);
      if (k == 5 && stride == 1) return launch_v4<5, 1>(col, R, Hin, Hout, pad, Cout, bias, gamma, beta, act, out, smem4, st, ypre);
      if (k == 5 && stride == 2) return launch_v4<5, 2>(col, R, Hin, Hout, pad, Cout, bias, gamma, beta, act, out, smem4, st, ypre);
    }
  }
  DESIRE_CHECK_ARG(!ypre, "colbn_act: pre-BN output needs the vectorised kernel (Cout %% 4 == 0, known k/stride)");
  DESIRE_CHECK_ARG(Cout >= 1 && Cout <= 256 && 256 % Cout == 0, "colbn_act: Cout=%d unsupported", Cout);
  size_t smem = ((size_t)Hout * Hout * Cout + 256 + 2 * Cout) * sizeof(float);
  DESIRE_CHECK_ARG(smem <= 227 * 1024, "colbn_act: tile too large");
  DESIRE_ENSURE_SMEM(colbn_act_kernel, 227 * 1024);
  DESIRE_LAUNCH(st, (colbn_act_kernel<<<R, 256, smem, st>>>(col, Hin, Hout, k, stride, pad, Cout, bias, gamma, beta, act, out)));
  return DESIRE_OK;
}

}  // namespace desire

using namespace desire;

// ------------------------------------------------------------------------------------------ encoder
static const int ENC_CHUNK = 8192;

extern "C" size_t desire_cvae_encode_workspace_bytes(int M, int Z) {
  (void)Z;
  size_t mc = M < ENC_CHUNK ? M : ENC_CHUNK;
  // col (<= mc*256*32) + a1 (mc*8192) + a2 (mc*4096) + a3 (mc*2048)
  return align_up(mc * 8192 * 4) + align_up(mc * 8192 * 4) + align_up(mc * 4096 * 4) + align_up(mc * 2048 * 4) + PACK_WS_BYTES;
}

extern "C" int desire_cvae_encode_fwd(const float* v, int M, int Z, const desire_cvae_enc_t* w, float* mu_logvar,
                                      void* ws, size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(v && w && mu_logvar && M >= 0 && Z > 0, "desire_cvae_encode_fwd: bad arguments");
  if (!ws || ws_bytes < desire_cvae_encode_workspace_bytes(M, Z)) {
    set_error("desire_cvae_encode_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int m0 = 0; m0 < M; m0 += ENC_CHUNK) {
    const int mc = (M - m0) < ENC_CHUNK ? (M - m0) : ENC_CHUNK;
    Workspace W(ws, ws_bytes);
    float* col = W.take<float>((size_t)mc * 8192);
    float* a1 = W.take<float>((size_t)mc * 8192);
    float* a2 = W.take<float>((size_t)mc * 4096);
    float* a3 = W.take<float>((size_t)mc * 2048);
    PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
    const float* x = v + (size_t)m0 * 1024;
    // conv5/2 SAME 32x32x1 -> 16x16x32 : pad_before = 1 (TF: total 3 -> 1 | 2)
    Im2col g1{32, 32, 1, 16, 16, 5, 5, 2, 1, 1};
    DESIRE_TRY(sgemm_im2col(x, g1, w->c1.w, 32, nullptr, col, 32, mc * 256, 32, 25, DESIRE_ACT_NONE, st, pw));
    DESIRE_TRY(colbn_act(col, mc, 16, 16, 1, 1, 0, 32, w->c1.b, w->c1.gamma, w->c1.beta, DESIRE_ACT_ELU, a1, st));
    // conv5/2 SAME 16x16x32 -> 8x8x64
    Im2col g2{16, 16, 32, 8, 8, 5, 5, 2, 1, 1};
    DESIRE_TRY(sgemm_im2col(a1, g2, w->c2.w, 64, nullptr, col, 64, mc * 64, 64, 800, DESIRE_ACT_NONE, st, pw));
    DESIRE_TRY(colbn_act(col, mc, 8, 8, 1, 1, 0, 64, w->c2.b, w->c2.gamma, w->c2.beta, DESIRE_ACT_ELU, a2, st));
    // conv5 VALID 8x8x64 -> 4x4x128
    Im2col g3{8, 8, 64, 4, 4, 5, 5, 1, 0, 0};
    DESIRE_TRY(sgemm_im2col(a2, g3, w->c3.w, 128, nullptr, col, 128, mc * 16, 128, 1600, DESIRE_ACT_NONE, st, pw));
    DESIRE_TRY(colbn_act(col, mc, 4, 4, 1, 1, 0, 128, w->c3.b, w->c3.gamma, w->c3.beta, DESIRE_ACT_ELU, a3, st));
    // flatten (h,w,c) -> fc 2048 -> 2Z, no BN, no activation
    DESIRE_TRY(sgemm(a3, 2048, w->fc_w, 2 * Z, false, w->fc_b, mu_logvar + (size_t)m0 * 2 * Z, 2 * Z, mc, 2 * Z, 2048,
                     DESIRE_ACT_NONE, false, st, pw));
  }
  return DESIRE_OK;
}

// ------------------------------------------------------------------------------------------ decoder
// Rows per pass.  With the fused deconv kernels the only col matrix left is deconv1's [rows, 2048], so the whole
// batch goes through in one pass (fewer launches, no per-chunk tails); the unfused fallback (gemm mode 0) keeps
// 4096-row chunks because its deconv3 col matrix is 205 KB per row.
static const int DEC_CHUNK_FUSED = 65536, DEC_CHUNK_COL = 4096;
static size_t dec_ws_bytes(size_t rc, size_t col_per_row) {
  return align_up(rc * col_per_row * 4) + align_up(rc * 2048 * 4) + align_up(rc * 4096 * 4) + align_up(rc * 8192 * 4) +
         PACK_WS_BYTES;
}

extern "C" size_t desire_cvae_decode_workspace_bytes(int R, int Z) {
  (void)Z;
  const size_t a = dec_ws_bytes(R < DEC_CHUNK_FUSED ? R : DEC_CHUNK_FUSED, 2048);
  const size_t b = dec_ws_bytes(R < DEC_CHUNK_COL ? R : DEC_CHUNK_COL, 51200);
  return a > b ? a : b;
}

extern "C" int desire_cvae_decode_fwd(const float* z, int R, int Z, const desire_cvae_dec_t* w, float* xr, void* ws,
                                      size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(z && w && xr && R >= 0 && Z > 0, "desire_cvae_decode_fwd: bad arguments");
  if (!ws || ws_bytes < desire_cvae_decode_workspace_bytes(R, Z)) {
    set_error("desire_cvae_decode_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = gemm_mode() != 0;     // deconv2/3 run fused (their eligibility only depends on the mode here)
  const int DEC_CHUNK = fused ? DEC_CHUNK_FUSED : DEC_CHUNK_COL;
  for (int r0 = 0; r0 < R; r0 += DEC_CHUNK) {
    const int rc = (R - r0) < DEC_CHUNK ? (R - r0) : DEC_CHUNK;
    Workspace W(ws, ws_bytes);
    float* col = W.take<float>((size_t)rc * (fused ? 2048 : 51200));
    float* a1 = W.take<float>((size_t)rc * 2048);
    float* a2 = W.take<float>((size_t)rc * 4096);
    float* a3 = W.take<float>((size_t)rc * 8192);
    PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
    const float* zc = z + (size_t)r0 * Z;
    // deconv4 VALID 1x1xZ -> 4x4x128: col[r, (y,x,o)] = z[r,:] . W[y,x,o,:]
    DESIRE_TRY(sgemm(zc, Z, w->d1.w, Z, true, nullptr, col, 2048, rc, 2048, Z, DESIRE_ACT_NONE, false, st, pw));
    {
      ProfScope ps_(DESIRE_PROF_COL2IM, st);
      DESIRE_TRY(colbn_act(col, rc, 1, 4, 4, 1, 0, 128, w->d1.b, w->d1.gamma, w->d1.beta, DESIRE_ACT_ELU, a1, st));
    }
    // deconv5 VALID 4x4x128 -> 8x8x64
    if (deconv_tc_eligible(rc, 4, 8, 128, 64, 5, 1, pw.p, pw.bytes)) {
      ProfScope ps_(DESIRE_PROF_DECONV2, st);
      DESIRE_TRY(deconv_tc(a1, rc, 4, 8, 128, 64, 5, 1, 0, w->d2.w, w->d2.b, w->d2.gamma, w->d2.beta, DESIRE_ACT_ELU, a2,
                           pw.p, st));
    } else {
      {
        ProfScope ps_(DESIRE_PROF_DECONV2, st);
        DESIRE_TRY(sgemm(a1, 128, w->d2.w, 128, true, nullptr, col, 1600, rc * 16, 1600, 128, DESIRE_ACT_NONE, false, st, pw));
      }
      ProfScope ps_(DESIRE_PROF_COL2IM, st);
      DESIRE_TRY(colbn_act(col, rc, 4, 8, 5, 1, 0, 64, w->d2.b, w->d2.gamma, w->d2.beta, DESIRE_ACT_ELU, a2, st));
    }
    // deconv5/2 SAME 8x8x64 -> 16x16x32 (full 19x19, keep [1,17))
    static const bool no_fuse4 = [] {                          // DESIRE_NO_FUSE4=1: last layer as its own kernel (A/B timing)
      const char* e = getenv("DESIRE_NO_FUSE4");
      return e && e[0] == '1';
    }();
    bool fused4 = false;
    if (deconv_tc_eligible(rc, 8, 16, 64, 32, 5, 2, pw.p, pw.bytes)) {
      ProfScope ps_(DESIRE_PROF_DECONV3, st);
      // the last layer (16x16x32 -> 32x32x1 + BN + sigmoid) rides on this kernel's normalised tile: a3 is never written
      DeconvFuse4 f4{w->d4.w, w->d4.b, w->d4.gamma, w->d4.beta, DESIRE_ACT_SIGMOID, xr + (size_t)r0 * 1024};
      fused4 = !no_fuse4 && deconv1c_tc_eligible() && pw.bytes >= align_up(deconv_tc_pack_bytes(64, 32, 5)) + 4096;
      DESIRE_TRY(deconv_tc(a2, rc, 8, 16, 64, 32, 5, 2, 1, w->d3.w, w->d3.b, w->d3.gamma, w->d3.beta, DESIRE_ACT_ELU, a3,
                           pw.p, st, fused4 ? &f4 : nullptr));
    } else {
      {
        ProfScope ps_(DESIRE_PROF_DECONV3, st);
        DESIRE_TRY(sgemm(a2, 64, w->d3.w, 64, true, nullptr, col, 800, rc * 64, 800, 64, DESIRE_ACT_NONE, false, st, pw));
      }
      ProfScope ps_(DESIRE_PROF_COL2IM, st);
      DESIRE_TRY(colbn_act(col, rc, 8, 16, 5, 2, 1, 32, w->d3.b, w->d3.gamma, w->d3.beta, DESIRE_ACT_ELU, a3, st));
    }
    // deconv5/2 SAME 16x16x32 -> 32x32x1, BN + sigmoid: tiny-N tcgen05 kernel with in-kernel col2im + BN
    // (CUDA-core direct kernel when tensor cores are off)
    if (!fused4) {
      ProfScope ps_(DESIRE_PROF_COL2IM, st);
      if (deconv1c_tc_eligible()) {
        DESIRE_TRY(deconv1c_tc(a3, rc, w->d4.w, w->d4.b, w->d4.gamma, w->d4.beta, DESIRE_ACT_SIGMOID,
                               xr + (size_t)r0 * 1024, pw.p, st));
      } else {
        const size_t smem4 = ((size_t)256 * 36 + 25 * 32 + 32) * sizeof(float);
        DESIRE_ENSURE_SMEM((deconv1ch_bn_act_kernel<32, 5, 2>), smem4);
        DESIRE_LAUNCH(st, (deconv1ch_bn_act_kernel<32, 5, 2><<<rc, 256, smem4, st>>>(a3, 16, 32, 1, w->d4.w, w->d4.b,
                                                                                  w->d4.gamma, w->d4.beta,
                                                                                  DESIRE_ACT_SIGMOID,
                                                                                  xr + (size_t)r0 * 1024)));
      }
    }
  }
  return DESIRE_OK;
}
