// HBM-bound pieces of the sample-generation stage: temporal depthwise conv (a2), reparameterisation
// (a7), softmax mask + gating (a9), read-out + feature pooling (a11), losses (a12/a13).
// One warp per output row wherever a row-wise reduction is needed (warp-shuffle reductions).
#include "common.cuh"

using namespace desire;

namespace {

// ---- a2: rho[m, c*C + j] = relu(sum_t obs[m,t,1+c] * w[t,c,j] + b[c*C+j])
__global__ void tconv_kernel(const float* __restrict__ obs, int M, int Tp, int C, const float* __restrict__ w,
                             const float* __restrict__ b, float* __restrict__ rho) {
  const int m = blockIdx.x;
  for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) {
    const int c = j / C, jj = j % C;
    float acc = __ldg(b + j);
    for (int t = 0; t < Tp; ++t)
      acc = fmaf(__ldg(obs + ((size_t)m * Tp + t) * 3 + 1 + c), __ldg(w + (t * 2 + c) * C + jj), acc);
    rho[(size_t)m * 2 * C + j] = fmaxf(acc, 0.f);
  }
}

// ---- a7: z[m,k,:] = mu[m] + sqrt(exp(logvar[m])) * eps[m,k,:]
__global__ void reparam_kernel(const float* __restrict__ ml, const float* __restrict__ eps, size_t total, int K,
                               int Z, float* __restrict__ z) {
  size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i4 * 4;
  if (i >= total) return;
  const size_t m = i / ((size_t)K * Z);
  const int zc = (int)(i % Z);
  float4 mu = *reinterpret_cast<const float4*>(ml + m * 2 * Z + zc);
  float4 lv = *reinterpret_cast<const float4*>(ml + m * 2 * Z + Z + zc);
  float4 e = *reinterpret_cast<const float4*>(eps + i);
  float4 o;
  o.x = mu.x + sqrtf(expf(lv.x)) * e.x;
  o.y = mu.y + sqrtf(expf(lv.y)) * e.y;
  o.z = mu.z + sqrtf(expf(lv.z)) * e.z;
  o.w = mu.w + sqrtf(expf(lv.w)) * e.w;
  *reinterpret_cast<float4*>(z + i) = o;
}

// ---- a9: x_z[r,:] = softmax(logits[r,:]) * Hx[r/K,:]   (logits already relu'd by the GEMM epilogue)
__global__ void softmax_gate_kernel(const float* __restrict__ logits, int R, int H, int K,
                                    const float* __restrict__ Hx, int ld_hx, float* __restrict__ xz) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* l = logits + (size_t)warp * H;
  float mx = -INFINITY;
  for (int c = lane; c < H; c += 32) mx = fmaxf(mx, l[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s += expf(l[c] - mx);
  s = warp_sum(s);
  const float* hx = Hx + (size_t)(warp / K) * ld_hx;
  for (int c = lane; c < H; c += 32) xz[(size_t)warp * H + c] = expf(l[c] - mx) / s * __ldg(hx + c);
}

// ---- a11: one warp per (r,t)
__global__ void readout_pool_kernel(const float* __restrict__ hs, int R, int K, int T, int H, int mode,
                                    int n_chunks, const float* __restrict__ out_w,
                                    const float* __restrict__ out_b, const float* __restrict__ obs, int Tp,
                                    const float* __restrict__ rho, int C, float* __restrict__ Yhat,
                                    float* __restrict__ fpool) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (size_t)R * T) return;
  const size_t r = warp / T;
  const size_t m = r / K;
  const float* h = hs + warp * H;
  const float* rh = rho ? rho + m * 2 * C : nullptr;
  if (mode == 0) {
    float yx = 0.f, yy = 0.f;
    for (int c = lane; c < H; c += 32) {
      float v = h[c];
      yx = fmaf(v, __ldg(out_w + 2 * c), yx);
      yy = fmaf(v, __ldg(out_w + 2 * c + 1), yy);
    }
    yx = warp_sum(yx);
    yy = warp_sum(yy);
    const float* last = obs + (m * Tp + (Tp - 1)) * 3;
    yx += __ldg(out_b) + __ldg(last + 1);
    yy += __ldg(out_b + 1) + __ldg(last + 2);
    if (lane == 0) {
      Yhat[warp * 2] = yx;
      Yhat[warp * 2 + 1] = yy;
    }
    if (fpool) {
      float* f = fpool + warp * 2 * C;
      for (int j = lane; j < 2 * C; j += 32) f[j] = (j < C ? yx : yy) * __ldg(rh + j);
    }
  } else {
    const int cs = H / n_chunks;
    for (int c = 0; c < n_chunks; ++c) {
      const float yx = h[c * cs], yy = h[c * cs + 1];
      const size_t o = warp * n_chunks + c;
      if (lane == 0) {
        Yhat[o * 2] = yx;
        Yhat[o * 2 + 1] = yy;
      }
      if (fpool) {
        float* f = fpool + o * 2 * C;
        for (int j = lane; j < 2 * C; j += 32) f[j] = (j < C ? yx : yy) * __ldg(rh + j);
      }
    }
  }
}

// ---- a12: kld_rows[m] = -0.5 * sum_z (1 + lv - mu^2 - exp(lv))
__global__ void kld_rows_kernel(const float* __restrict__ ml, int M, int Z, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* p = ml + (size_t)warp * 2 * Z;
  float s = 0.f;
  for (int c = lane; c < Z; c += 32) {
    float mu = p[c], lv = p[Z + c];
    s += 1.f + lv - mu * mu - expf(lv);
  }
  s = warp_sum(s);
  if (lane == 0) out[warp] = -0.5f * s;
}

// ---- D7: recon_rows[m] = mean_k sum_t |Y - Yhat_k|^2
__global__ void recon_rows_kernel(const float* __restrict__ Yhat, const float* __restrict__ tgt, int M, int K, int T,
                                  float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float s = 0.f;
  for (int i = lane; i < K * T * 2; i += 32) {
    int k = i / (T * 2), rem = i % (T * 2), t = rem / 2, c = rem % 2;
    float d = Yhat[(((size_t)warp * K + k) * T + t) * 2 + c] - __ldg(tgt + ((size_t)warp * T + t) * 3 + 1 + c);
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  if (lane == 0) out[warp] = s / (float)K;
}

// ---- a13: single-CTA deterministic masked mean
__global__ void masked_cost_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const float* __restrict__ obs, int M, int Tp, float* __restrict__ cost) {
  __shared__ float ssum[32], scnt[32];
  float s = 0.f, n = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    if (__ldg(obs + (size_t)m * Tp * 3) != 0.f) {
      s += a[m] + (b ? b[m] : 0.f);
      n += 1.f;
    }
  }
  s = warp_sum(s);
  n = warp_sum(n);
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = s;
    scnt[threadIdx.x >> 5] = n;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? ssum[threadIdx.x] : 0.f;
    n = threadIdx.x < nw ? scnt[threadIdx.x] : 0.f;
    s = warp_sum(s);
    n = warp_sum(n);
    if (threadIdx.x == 0) {
      cost[0] = s / n;
      cost[1] = n;
    }
  }
}

// ---- D8 existence: a copy of the observations whose id at observed frame 0 — the field every kernel of the path tests
// for "this agent exists" — is zeroed for agents that are absent where the losses and the decoders need them.
//   mode 0: the id at observed frame 0 only (obj_id of model/model.py:214,357);
//   mode 1: also absent at the LAST observed frame (the read-out anchors on that position) or at ANY target frame
//           (`target_obj_id` of model/model.py:358: an object missing from the target must not contribute to the cost).
__global__ void existence_kernel(const float* __restrict__ obs, const float* __restrict__ tgt, int M, int Tp, int Tf,
                                 int mode, float* __restrict__ out) {
  const size_t n = (size_t)M * Tp * 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = obs[i];
    if (i % ((size_t)Tp * 3) == 0 && mode == 1 && v != 0.f) {
      const size_t m = i / ((size_t)Tp * 3);
      bool ok = obs[i + (size_t)(Tp - 1) * 3] != 0.f;
      for (int t = 0; t < Tf && ok; ++t) ok = tgt[(m * Tf + t) * 3] != 0.f;
      if (!ok) v = 0.f;
    }
    out[i] = v;
  }
}

// ---- a7 noise source: eps ~ N(0, I) drawn on the device (the reference draws it inside the graph with
// tf.random_normal, model/model.py:262).  Philox4x32-10 counter-based generator (Salmon et al., SC'11): thread i
// encrypts counter (i, offset) under key `seed` into four 32-bit words -> four uniforms -> two Box-Muller pairs, so
// element e of a draw is a pure function of (seed, offset, e) — any rank or the oracle can reproduce it.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__global__ void randn_kernel(const unsigned long long* __restrict__ state, float* __restrict__ out, size_t n) {
  const unsigned long long seed = state[0], offset = state[1];
  const size_t quads = (n + 3) / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)(c[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-8f;       // (0,1), 24 bits
      const float u2 = ((float)(c[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
      const float rad = sqrtf(-2.f * logf(u1));
      float sn, cs;
      sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * h] = rad * cs;
      z[2 * h + 1] = rad * sn;
    }
    if (4 * i + 3 < n) {
      *reinterpret_cast<float4*>(out + 4 * i) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
      for (size_t e = 4 * i; e < n; ++e) out[e] = z[e - 4 * i];
    }
  }
}

inline unsigned warps_grid(size_t warps, int threads) { return (unsigned)((warps * 32 + threads - 1) / threads); }

}  // namespace

extern "C" int desire_tconv_fwd(const float* obs, int M, int Tp, int C, const float* w, const float* b, float* rho,
                                desire_stream_t stream) {
  DESIRE_CHECK_ARG(obs && w && b && rho && M >= 0 && Tp > 0 && C > 0, "desire_tconv_fwd: bad arguments");
  if (M == 0) return DESIRE_OK;
  tconv_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(obs, M, Tp, C, w, b, rho);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_reparam_fwd(const float* mu_logvar, const float* eps, int M, int K, int Z, float* z,
                                  desire_stream_t stream) {
  DESIRE_CHECK_ARG(mu_logvar && eps && z && M >= 0 && K > 0 && Z > 0 && Z % 4 == 0, "desire_reparam_fwd: bad arguments");
  size_t total = (size_t)M * K * Z;
  if (total == 0) return DESIRE_OK;
  size_t n4 = total / 4;
  reparam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mu_logvar, eps, total, K, Z, z);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" size_t desire_mask_softmax_workspace_bytes(int R, int H) {
  return align_up((size_t)R * H * sizeof(float)) + PACK_WS_BYTES;
}

extern "C" int desire_mask_softmax_fwd(const float* xr, int R, int S2, int H, int K, const float* w, const float* b,
                                       const float* Hx, int ld_hx, float* x_z, void* ws, size_t ws_bytes,
                                       desire_stream_t stream) {
  DESIRE_CHECK_ARG(xr && w && b && Hx && x_z && R >= 0 && K > 0, "desire_mask_softmax_fwd: bad arguments");
  if (!ws || ws_bytes < desire_mask_softmax_workspace_bytes(R, H)) {
    set_error("desire_mask_softmax_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  if (R == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* logits = (float*)ws;
  PackWs pw{(char*)ws + align_up((size_t)R * H * sizeof(float)), PACK_WS_BYTES};
  DESIRE_TRY(sgemm(xr, S2, w, H, false, b, logits, H, R, H, S2, DESIRE_ACT_RELU, false, st, pw));
  softmax_gate_kernel<<<warps_grid(R, 256), 256, 0, st>>>(logits, R, H, K, Hx, ld_hx, x_z);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_readout_pool_fwd(const float* hs, int R, int K, int T, int H, int mode, int n_chunks,
                                       const float* out_w, const float* out_b, const float* obs, int Tp,
                                       const float* rho, int C, float* Yhat, float* fpool, desire_stream_t stream) {
  DESIRE_CHECK_ARG(hs && Yhat && R >= 0 && K > 0 && T > 0, "desire_readout_pool_fwd: bad arguments");
  if (mode == 0)
    DESIRE_CHECK_ARG(out_w && out_b && obs && Tp > 0, "desire_readout_pool_fwd: linear read-out needs out_w/out_b/obs");
  else
    DESIRE_CHECK_ARG(mode == 1 && n_chunks > 0 && H % n_chunks == 0 && H / n_chunks >= 2,
                     "desire_readout_pool_fwd: split read-out needs H divisible by n_chunks with chunks >= 2");
  DESIRE_CHECK_ARG(!fpool || (rho && C > 0), "desire_readout_pool_fwd: feature pooling needs rho");
  if (R == 0) return DESIRE_OK;
  readout_pool_kernel<<<warps_grid((size_t)R * T, 256), 256, 0, (cudaStream_t)stream>>>(
      hs, R, K, T, H, mode, n_chunks, out_w, out_b, obs, Tp, rho, C, Yhat, fpool);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_kld_rows_fwd(const float* mu_logvar, int M, int Z, float* kld_rows, desire_stream_t stream) {
  DESIRE_CHECK_ARG(mu_logvar && kld_rows && M >= 0 && Z > 0, "desire_kld_rows_fwd: bad arguments");
  if (M == 0) return DESIRE_OK;
  kld_rows_kernel<<<warps_grid(M, 256), 256, 0, (cudaStream_t)stream>>>(mu_logvar, M, Z, kld_rows);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_recon_rows_fwd(const float* Yhat, const float* target, int M, int K, int T, float* recon_rows,
                                     desire_stream_t stream) {
  DESIRE_CHECK_ARG(Yhat && target && recon_rows && M >= 0 && K > 0 && T > 0, "desire_recon_rows_fwd: bad arguments");
  if (M == 0) return DESIRE_OK;
  recon_rows_kernel<<<warps_grid(M, 256), 256, 0, (cudaStream_t)stream>>>(Yhat, target, M, K, T, recon_rows);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_randn_fwd(const unsigned long long* state, float* out, size_t n, desire_stream_t stream) {
  DESIRE_CHECK_ARG(state && out && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "desire_randn_fwd: bad arguments");
  if (n == 0) return DESIRE_OK;
  const size_t quads = (n + 3) / 4;
  randn_kernel<<<(unsigned)((quads + 255) / 256 < 148 * 16 ? (quads + 255) / 256 : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      state, out, n);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_existence_fwd(const float* obs, const float* target, int M, int Tp, int Tf, int mode, float* obs_out,
                                    desire_stream_t stream) {
  DESIRE_CHECK_ARG(obs && obs_out && M >= 0 && Tp > 0 && (mode == 0 || (mode == 1 && target && Tf > 0)),
                   "desire_existence_fwd: bad arguments");
  if (M == 0) return DESIRE_OK;
  const size_t n = (size_t)M * Tp * 3;
  existence_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, (cudaStream_t)stream>>>(
      obs, target, M, Tp, Tf, mode, obs_out);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_masked_cost_fwd(const float* rows_a, const float* rows_b, const float* obs, int M, int Tp,
                                      float* cost, desire_stream_t stream) {
  DESIRE_CHECK_ARG(rows_a && obs && cost && M > 0 && Tp > 0, "desire_masked_cost_fwd: bad arguments");
  masked_cost_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rows_a, rows_b, obs, M, Tp, cost);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}
