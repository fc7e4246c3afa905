// Train step of the sample-generation stage: gradients of `cost` (model/model.py:374-376) with respect to
// every variable on the path a3-a13, clip_by_global_norm + Adam (model/model.py:388-394).
//
// Structure: every backward entry point recomputes the forward intermediates it needs (pre-BN activations,
// gates) with the SAME GEMM engine as the forward (tcgen05 3xBF16 through sgemm()/sgemm_im2col(), FP32 CUDA
// cores in mode 0), then runs
//   - input gradients as GEMMs against the transposed weights (deconv dgrad == forward conv through the
//     im2col loader; conv dgrad == dy @ W^T followed by the col2im gather below),
//   - weight gradients with the split-row FP32 wgrad_tn kernels of gemm_f32.cu (A^T @ B over up to millions
//     of rows, implicit im2col for the conv / deconv filters),
//   - the element-wise / per-row pieces (GRU cell, per-row BN + activation, softmax gate, losses) with the
//     kernels in this file.
// Parameter gradients are accumulated with atomicAdd into a caller-zeroed flat buffer.
#include "common.cuh"

using namespace desire;

namespace {

inline unsigned grid1d(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------ losses
__global__ void cost_bwd_y_kernel(const float* __restrict__ Yhat, const float* __restrict__ tgt,
                                  const float* __restrict__ obs, const float* __restrict__ count, size_t total,
                                  int K, int T, int Tp, float* __restrict__ dY) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i & 1);
  const size_t rt = i >> 1;
  const int t = (int)(rt % T);
  const size_t m = (rt / T) / K;
  const float g = (__ldg(obs + m * Tp * 3) != 0.f) ? 1.f / __ldg(count) : 0.f;
  const float d = Yhat[i] - __ldg(tgt + (m * T + t) * 3 + 1 + c);
  dY[i] = 2.f * d / (float)K * g;
}

__global__ void kld_bwd_kernel(const float* __restrict__ ml, const float* __restrict__ obs,
                               const float* __restrict__ count, int M, int Z, int Tp, float* __restrict__ dml) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * Z) return;
  const size_t m = i / Z;
  const int z = (int)(i % Z);
  const float g = (__ldg(obs + m * Tp * 3) != 0.f) ? 1.f / __ldg(count) : 0.f;
  const float mu = ml[m * 2 * Z + z], lv = ml[m * 2 * Z + Z + z];
  dml[m * 2 * Z + z] = mu * g;
  dml[m * 2 * Z + Z + z] = -0.5f * (1.f - expf(lv)) * g;
}

// ------------------------------------------------------------------------------------------ read-out
__global__ void readout_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ w, size_t rows, int H,
                                   float* __restrict__ dhs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const size_t rt = i / H;
  const int h = (int)(i % H);
  dhs[i] = dY[rt * 2] * __ldg(w + 2 * h) + dY[rt * 2 + 1] * __ldg(w + 2 * h + 1);
}

// ------------------------------------------------------------------------------------------ GRU cell
// hoisted input projection of the encoders: xp[m,t,:] = [x,y] @ W_x + b   (r|u from wg/bg, c from wc/bc)
__global__ void xproj_traj_kernel(const float* __restrict__ traj, size_t rows, int H, const float* __restrict__ wg,
                                  const float* __restrict__ bg, const float* __restrict__ wc,
                                  const float* __restrict__ bc, float* __restrict__ xp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 3 * H) return;
  const size_t row = i / (3 * H);
  const int j = (int)(i % (3 * H));
  const float x0 = __ldg(traj + row * 3 + 1), x1 = __ldg(traj + row * 3 + 2);
  float v;
  if (j < 2 * H)
    v = fmaf(x1, __ldg(wg + 2 * H + j), fmaf(x0, __ldg(wg + j), __ldg(bg + j)));
  else
    v = fmaf(x1, __ldg(wc + H + (j - 2 * H)), fmaf(x0, __ldg(wc + (j - 2 * H)), __ldg(bc + (j - 2 * H))));
  xp[i] = v;
}

}  // namespace
namespace desire {
int xproj_traj(const float* traj, size_t rows, int H, const desire_gru_t* w, float* xp, cudaStream_t st) {
  if (rows == 0) return DESIRE_OK;
  DESIRE_LAUNCH(st, (xproj_traj_kernel<<<grid1d(rows * 3 * H), 256, 0, st>>>(traj, rows, H, w->wg, w->bg, w->wc, w->bc, xp)));
  return DESIRE_OK;
}
}  // namespace desire
namespace {
// rows [R,H] <- src[(r / div) * ld + c]   (src may be null: zeros)
__global__ void expand_rows_bwd_kernel(const float* __restrict__ src, int div, int ld, size_t R, int H,
                                       float* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * H) return;
  const size_t r = i / H;
  const int c = (int)(i % H);
  dst[i] = src ? __ldg(src + (r / div) * (size_t)ld + c) : 0.f;
}

// dst[m*ld + c] += sum_{k<K} src[(m*K+k)*H + c]
__global__ void reduce_k_rows_kernel(const float* __restrict__ src, size_t M, int K, int H, float* __restrict__ dst,
                                     int ld) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * H) return;
  const size_t m = i / H;
  const int c = (int)(i % H);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += src[(m * K + k) * H + c];
  dst[m * ld + c] += s;
}

// step-major copy of the previous states: hp_all[t][r][:] = h_{t-1}(r)  (h0e for t == 0)
__global__ void gru_hprev_gather_kernel(const float* __restrict__ hs, long hs_rs, long hs_ss, const float* __restrict__ h0e,
                                        size_t R, int T, int H, float* __restrict__ hp_all) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * T * H) return;
  const int c = (int)(i % H);
  const size_t tr = i / H;
  const size_t r = tr % R;
  const int t = (int)(tr / R);
  hp_all[i] = t > 0 ? hs[r * hs_rs + (size_t)(t - 1) * hs_ss + c] : h0e[r * H + c];
}

// gates of every step: r,u = sigmoid(gh + xp_ru(r,t));  rh = r * h_prev       (all buffers step-major [T][R][.])
__global__ void gru_bwd_gates_all_kernel(const float* __restrict__ gh, const float* __restrict__ xp, long xp_rs, long xp_ss,
                                         const float* __restrict__ hp, size_t R, int T, int H, float* __restrict__ ru,
                                         float* __restrict__ rh) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * T * H) return;
  const int c = (int)(i % H);
  const size_t tr = i / H;
  const size_t r = tr % R;
  const int t = (int)(tr / R);
  const float* x = xp + r * xp_rs + (size_t)t * xp_ss;
  const float rr = sigmoidf_(gh[tr * 2 * H + c] + x[c]);
  const float uu = sigmoidf_(gh[tr * 2 * H + H + c] + x[H + c]);
  ru[tr * 2 * H + c] = rr;
  ru[tr * 2 * H + H + c] = uu;
  rh[i] = rr * hp[i];
}

// c = tanh(ch + xp_c); from dh: dcpre, dupre (-> dg[:,H:]), tmp = dh*u; dxp[c], dxp[u] +=
__global__ void gru_bwd_cand_kernel(const float* __restrict__ ch, const float* __restrict__ xp, long xp_rs,
                                    const float* __restrict__ hp, long hp_rs, const float* __restrict__ ru,
                                    const float* __restrict__ dh, long dh_rs, size_t R, int H,
                                    float* __restrict__ dcpre, float* __restrict__ dg, float* __restrict__ tmp,
                                    float* __restrict__ dxp, long dxp_rs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * H) return;
  const size_t r = i / H;
  const int c = (int)(i % H);
  const float cc = tanhf(ch[i] + xp[r * xp_rs + 2 * H + c]);
  const float u = ru[r * 2 * H + H + c];
  const float h = hp[r * hp_rs + c];
  const float d = dh[r * dh_rs + c];
  const float dcp = d * (1.f - u) * (1.f - cc * cc);
  const float dup = d * (h - cc) * u * (1.f - u);
  dcpre[i] = dcp;
  dg[r * 2 * H + H + c] = dup;
  tmp[i] = d * u;
  dxp[r * dxp_rs + 2 * H + c] += dcp;
  dxp[r * dxp_rs + H + c] += dup;
}

// dr = drh*h_prev; drpre -> dg[:, :H]; target += tmp + drh*r; dxp[r] +=
__global__ void gru_bwd_reset_kernel(const float* __restrict__ drh, const float* __restrict__ hp, long hp_rs,
                                     const float* __restrict__ ru, const float* __restrict__ tmp, size_t R, int H,
                                     float* __restrict__ dg, float* __restrict__ target, long tg_rs,
                                     float* __restrict__ dxp, long dxp_rs) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * H) return;
  const size_t r = i / H;
  const int c = (int)(i % H);
  const float rr = ru[r * 2 * H + c];
  const float d = drh[i];
  const float drp = d * hp[r * hp_rs + c] * rr * (1.f - rr);
  dg[r * 2 * H + c] = drp;
  target[r * tg_rs + c] += tmp[i] + d * rr;
  dxp[r * dxp_rs + c] += drp;
}

}  // namespace

namespace desire {
size_t gru_bptt_ws_bytes(size_t R, int H, int T) {
  // step-major [T][R][.] buffers: gh/dg[2H] ru[2H] hp rh ch dcpre [H each]; per-step drh tmp [H]; GEMM + wgrad packs
  const size_t RT = R * (size_t)T;
  return 2 * align_up(RT * 2 * H * 4) + 4 * align_up(RT * H * 4) + 2 * align_up(R * H * 4) + PACK_WS_BYTES +
         wgrad_tc_pack_bytes((int)RT, 2 * H);
}

// The forward quantities the backward needs (gates r,u and the candidate pre-activation) only depend on the SAVED
// states h_{t-1}, so they are recomputed for ALL T steps at once — two large GEMMs and one element-wise kernel over
// [T*R] rows instead of two small GEMMs per step — and the two recurrent weight gradients become one tcgen05 product
// each over T*R rows after the loop.  The serial part is what is inherently serial: per step two element-wise kernels
// and the two input-gradient GEMMs that carry d h_{t-1}.
int gru_bptt(const GruBptt& a, void* ws, size_t ws_bytes, cudaStream_t st, const std::function<int(int)>* after_step) {
  const size_t R = a.R, RT = R * (size_t)a.T;
  const int H = a.H, T = a.T;
  DESIRE_CHECK_ARG(RT < ((size_t)1 << 31), "gru_bptt: R*T too large");
  Workspace W(ws, ws_bytes);
  float* gh_all = W.take<float>(RT * 2 * H);      // h_prev @ Wg_h for every step; reused as d(gate pre-activations)
  float* ru_all = W.take<float>(RT * 2 * H);
  float* hp_all = W.take<float>(RT * H);
  float* rh_all = W.take<float>(RT * H);
  float* ch_all = W.take<float>(RT * H);
  float* dcpre_all = W.take<float>(RT * H);
  float* drh = W.take<float>(R * H);
  float* tmp = W.take<float>(R * H);
  char* pk = W.take<char>(PACK_WS_BYTES);
  const size_t wpb = wgrad_tc_pack_bytes((int)RT, 2 * H);
  char* wpk = W.take<char>(wpb);
  if (!pk || !wpk) {
    set_error("gru_bptt: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  PackWs pw{pk, PACK_WS_BYTES}, wp{wpk, wpb};
  float* dg_all = gh_all;
  const float* wg_h = a.wg + (size_t)a.I * 2 * H;
  const float* wc_h = a.wc + (size_t)a.I * H;
  float* dwg_h = a.dwg + (size_t)a.I * 2 * H;
  float* dwc_h = a.dwc + (size_t)a.I * H;
  // ---- forward recompute, all steps at once
  DESIRE_LAUNCH(st, (gru_hprev_gather_kernel<<<grid1d(RT * H), 256, 0, st>>>(a.hs, a.hs_rs, a.hs_ss, a.h0e, R, T, H, hp_all)));
  DESIRE_TRY(sgemm(hp_all, H, wg_h, 2 * H, false, nullptr, gh_all, 2 * H, (int)RT, 2 * H, H, DESIRE_ACT_NONE, false, st, pw));
  DESIRE_LAUNCH(st, (gru_bwd_gates_all_kernel<<<grid1d(RT * H), 256, 0, st>>>(gh_all, a.xp, a.xp_rs, a.xp_ss, hp_all, R, T, H,
                                                                             ru_all, rh_all)));
  DESIRE_TRY(sgemm(rh_all, H, wc_h, H, false, nullptr, ch_all, H, (int)RT, H, H, DESIRE_ACT_NONE, false, st, pw));
  // ---- backward through time
  const unsigned g = grid1d(R * H);
  for (int t = T - 1; t >= 0; --t) {
    const float* hp = hp_all + (size_t)t * R * H;
    const float* ru = ru_all + (size_t)t * R * 2 * H;
    float* dg = dg_all + (size_t)t * R * 2 * H;
    float* dcpre = dcpre_all + (size_t)t * R * H;
    const float* xp = a.xp + (size_t)t * a.xp_ss;
    float* dxp = a.dxp + (size_t)t * a.dxp_ss;
    float* target = t > 0 ? a.dhs + (size_t)(t - 1) * a.dhs_ss : a.dh0;
    const long tg_rs = t > 0 ? a.dhs_rs : H;
    DESIRE_LAUNCH(st, (gru_bwd_cand_kernel<<<g, 256, 0, st>>>(ch_all + (size_t)t * R * H, xp, a.xp_rs, hp, H, ru,
                                                              a.dhs + (size_t)t * a.dhs_ss, a.dhs_rs, R, H, dcpre, dg,
                                                              tmp, dxp, a.dxp_rs)));
    // d(r*h) = dcpre @ Wc_h^T
    DESIRE_TRY(sgemm(dcpre, H, wc_h, H, true, nullptr, drh, H, a.R, H, H, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_LAUNCH(st, (gru_bwd_reset_kernel<<<g, 256, 0, st>>>(drh, hp, H, ru, tmp, R, H, dg, target, tg_rs, dxp, a.dxp_rs)));
    // d h_prev += dg @ Wg_h^T
    DESIRE_TRY(sgemm(dg, 2 * H, wg_h, 2 * H, true, nullptr, target, (int)tg_rs, a.R, H, 2 * H, DESIRE_ACT_NONE, true, st, pw));
    if (after_step) DESIRE_TRY((*after_step)(t));
  }
  // ---- recurrent weight gradients over all T*R rows
  DESIRE_TRY(wgrad_tn(hp_all, H, dg_all, 2 * H, dwg_h, 2 * H, (int)RT, H, 2 * H, st, wp));
  DESIRE_TRY(wgrad_tn(rh_all, H, dcpre_all, H, dwc_h, H, (int)RT, H, H, st, wp));
  return DESIRE_OK;
}
}  // namespace desire

namespace {

// ------------------------------------------------------------------------------------------ softmax gate
// one warp per row: beta = softmax(l); dl = relu'(l) * beta * (dbeta - sum beta dbeta), dbeta = dxz*Hx;
// gb = dxz * beta (summed over k into dHx afterwards)
__global__ void softmax_gate_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dxz, int R, int H,
                                        int K, const float* __restrict__ Hx, int ld_hx, float* __restrict__ dl,
                                        float* __restrict__ gb) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* l = logits + (size_t)warp * H;
  const float* d = dxz + (size_t)warp * H;
  const float* hx = Hx + (size_t)(warp / K) * ld_hx;
  float mx = -INFINITY;
  for (int c = lane; c < H; c += 32) mx = fmaxf(mx, l[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s += expf(l[c] - mx);
  s = warp_sum(s);
  float dot = 0.f;
  for (int c = lane; c < H; c += 32) dot += expf(l[c] - mx) / s * d[c] * __ldg(hx + c);
  dot = warp_sum(dot);
  for (int c = lane; c < H; c += 32) {
    const float beta = expf(l[c] - mx) / s;
    const float db = d[c] * __ldg(hx + c);
    gb[(size_t)warp * H + c] = d[c] * beta;
    dl[(size_t)warp * H + c] = l[c] > 0.f ? beta * (db - dot) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------ reparam
__global__ void reparam_bwd_kernel(const float* __restrict__ ml, const float* __restrict__ eps,
                                   const float* __restrict__ dz, int M, int K, int Z, float* __restrict__ dml) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * Z) return;
  const size_t m = i / Z;
  const int z = (int)(i % Z);
  const float sd = sqrtf(expf(ml[m * 2 * Z + Z + z]));
  float s0 = 0.f, s1 = 0.f;
  for (int k = 0; k < K; ++k) {
    const float d = dz[(m * K + k) * Z + z];
    s0 += d;
    s1 = fmaf(d, eps[(m * K + k) * Z + z], s1);
  }
  dml[m * 2 * Z + z] += s0;
  dml[m * 2 * Z + Z + z] += s1 * 0.5f * sd;
}

// ------------------------------------------------------------------------------------------ conv pieces
// col2im, gather form: out[r,oy,ox,c] = bias[c] + sum_{ky,kx} col[(r,iy,ix), (ky,kx,c)],
// iy = (oy + pad - ky)/stride when divisible and in range.  Covers the transposed convs of the decoder
// (forward recompute) and the input gradient of the encoder convs.
__global__ void col2im_gather_kernel(const float* __restrict__ col, size_t total, int Hin, int Hout, int k, int stride,
                                     int pad, int C, const float* __restrict__ bias, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  size_t t = i / C;
  const int ox = (int)(t % Hout);
  t /= Hout;
  const int oy = (int)(t % Hout);
  const size_t r = t / Hout;
  float acc = bias ? __ldg(bias + c) : 0.f;
  const size_t ldc = (size_t)k * k * C;
  for (int ky = 0; ky < k; ++ky) {
    const int ny = oy + pad - ky;
    if (ny < 0 || ny % stride) continue;
    const int iy = ny / stride;
    if (iy >= Hin) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int nx = ox + pad - kx;
      if (nx < 0 || nx % stride) continue;
      const int ix = nx / stride;
      if (ix >= Hin) continue;
      acc += col[((r * Hin + iy) * Hin + ix) * ldc + (size_t)(ky * k + kx) * C + c];
    }
  }
  out[i] = acc;
}

__device__ __forceinline__ float act_grad_from_pre(float yhat, int act) {
  switch (act) {
    case DESIRE_ACT_RELU: return yhat > 0.f ? 1.f : 0.f;
    case DESIRE_ACT_ELU: return yhat > 0.f ? 1.f : expf(yhat);
    case DESIRE_ACT_SIGMOID: {
      const float s = 1.f / (1.f + expf(-yhat));
      return s * (1.f - s);
    }
    default: return 1.f;
  }
}

// per-row BN (+activation), one CTA per row r of y [R,P,C] (C divides 256, C <= 256):
// forward  out = act(gamma*(y-mean)*rstd + beta)
template <bool BWD>
__global__ void __launch_bounds__(256) bn_row_kernel(const float* __restrict__ y, int P, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     int act, float* __restrict__ out /*fwd*/,
                                                     float* __restrict__ g /*bwd: in dout, out dy*/,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s_a[256], s_b[256], s_mean[256], s_rstd[256];
  const int tid = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * P * C;
  const int c = tid % C;
  const int pstep = 256 / C, p0 = tid / C;
  const float invP = 1.f / (float)P;
  // mean
  float s = 0.f;
  for (int p = p0; p < P; p += pstep) s += y[base + (size_t)p * C + c];
  s_a[tid] = s;
  __syncthreads();
  if (tid < C) {
    float a = 0.f;
    for (int j = tid; j < 256; j += C) a += s_a[j];
    s_mean[tid] = a * invP;
  }
  __syncthreads();
  const float mean = s_mean[c];
  // biased variance (two-pass, like the forward kernel)
  s = 0.f;
  for (int p = p0; p < P; p += pstep) {
    const float d = y[base + (size_t)p * C + c] - mean;
    s = fmaf(d, d, s);
  }
  __syncthreads();
  s_a[tid] = s;
  __syncthreads();
  if (tid < C) {
    float a = 0.f;
    for (int j = tid; j < 256; j += C) a += s_a[j];
    s_rstd[tid] = rsqrtf(a * invP + 1e-3f);
  }
  __syncthreads();
  const float rstd = s_rstd[c];
  const float ga = __ldg(gamma + c), be = __ldg(beta + c);
  if (!BWD) {
    for (int p = p0; p < P; p += pstep) {
      const size_t i = base + (size_t)p * C + c;
      out[i] = act_apply(fmaf(ga, (y[i] - mean) * rstd, be), act);
    }
    return;
  }
  // backward: gq = dout * act'(yhat); sums of gq and gq*xhat over P
  float sg = 0.f, sgx = 0.f;
  for (int p = p0; p < P; p += pstep) {
    const size_t i = base + (size_t)p * C + c;
    const float xh = (y[i] - mean) * rstd;
    const float gq = g[i] * act_grad_from_pre(fmaf(ga, xh, be), act);
    sg += gq;
    sgx = fmaf(gq, xh, sgx);
  }
  __syncthreads();
  s_a[tid] = sg;
  s_b[tid] = sgx;
  __syncthreads();
  if (tid < C) {
    float a = 0.f, b = 0.f;
    for (int j = tid; j < 256; j += C) {
      a += s_a[j];
      b += s_b[j];
    }
    s_mean[tid] = a;   // reuse: sum gq
    s_rstd[tid] = b;   //        sum gq*xhat
    atomicAdd(dbeta + tid, a);
    atomicAdd(dgamma + tid, b);
  }
  __syncthreads();
  const float m1 = s_mean[c] * invP, m2 = s_rstd[c] * invP;
  for (int p = p0; p < P; p += pstep) {
    const size_t i = base + (size_t)p * C + c;
    const float xh = (y[i] - mean) * rstd;
    const float gq = g[i] * act_grad_from_pre(fmaf(ga, xh, be), act);
    g[i] = ga * rstd * (gq - m1 - xh * m2);
  }
}

int bn_row_fwd(const float* y, int R, int P, int C, const float* gamma, const float* beta, int act, float* out,
               cudaStream_t st) {
  DESIRE_CHECK_ARG(C > 0 && C <= 256 && 256 % C == 0, "bn_row: C=%d must divide 256", C);
  if (R == 0) return DESIRE_OK;
  DESIRE_LAUNCH(st, (bn_row_kernel<false><<<R, 256, 0, st>>>(y, P, C, gamma, beta, act, out, nullptr, nullptr, nullptr)));
  return DESIRE_OK;
}
int bn_row_bwd(const float* y, int R, int P, int C, const float* gamma, const float* beta, int act, float* g,
               float* dgamma, float* dbeta, cudaStream_t st) {
  DESIRE_CHECK_ARG(C > 0 && C <= 256 && 256 % C == 0, "bn_row: C=%d must divide 256", C);
  if (R == 0) return DESIRE_OK;
  DESIRE_LAUNCH(st, (bn_row_kernel<true><<<R, 256, 0, st>>>(y, P, C, gamma, beta, act, nullptr, g, dgamma, dbeta)));
  return DESIRE_OK;
}

}  // namespace
namespace desire {
int col2im_gather(const float* col, int R, int Hin, int Hout, int k, int stride, int pad, int C, const float* bias,
                  float* out, cudaStream_t st) {
  const size_t total = (size_t)R * Hout * Hout * C;
  if (total == 0) return DESIRE_OK;
  DESIRE_LAUNCH(st, (col2im_gather_kernel<<<grid1d(total), 256, 0, st>>>(col, total, Hin, Hout, k, stride, pad, C, bias, out)));
  return DESIRE_OK;
}
}  // namespace desire
namespace {

// dC <- dC * act'(from the post-activation output)
__global__ void act_bwd_post_kernel(const float* __restrict__ out, int ldo, float* __restrict__ d, int ldd, size_t M,
                                    int N, int act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const size_t m = i / N;
  const int n = (int)(i % N);
  const float o = out[m * ldo + n];
  float f = 1.f;
  if (act == DESIRE_ACT_RELU) f = o > 0.f ? 1.f : 0.f;
  else if (act == DESIRE_ACT_ELU) f = o > 0.f ? 1.f : o + 1.f;
  else if (act == DESIRE_ACT_SIGMOID) f = o * (1.f - o);
  d[m * ldd + n] *= f;
}

}  // namespace
namespace desire {
int act_bwd_post(const float* out, int ldo, float* d, int ldd, size_t M, int N, int act, cudaStream_t st) {
  if (M == 0 || N == 0 || act == DESIRE_ACT_NONE) return DESIRE_OK;
  DESIRE_LAUNCH(st, (act_bwd_post_kernel<<<grid1d(M * N), 256, 0, st>>>(out, ldo, d, ldd, M, N, act)));
  return DESIRE_OK;
}
}  // namespace desire
namespace {
// ------------------------------------------------------------------------------------------ Adam
// Deterministic sum of squares: ONE block, fixed per-thread strides and a fixed reduction tree, so every rank gets the
// bit-identical norm from the bit-identical all-reduced gradient (an atomics-based reduction made the clip factor —
// and then the replicas' weights — differ in the last bit).  ~10 us for the 2.7 M parameters of the path.
__global__ void __launch_bounds__(1024) sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out,
                                                     int accumulate) {
  __shared__ float sm[32];
  float s = 0.f;
  for (size_t i = threadIdx.x; i < n; i += 1024) s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sm[threadIdx.x];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + s;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, const float* __restrict__ sumsq, float lr_t, float b1,
                            float b2, float eps, float clip, float gscale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float scale = gscale;
  if (clip > 0.f && sumsq) {
    const float norm = sqrtf(__ldg(sumsq)) * gscale;
    scale *= clip / fmaxf(norm, clip);
  }
  const float gi = g[i] * scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace

// =========================================================================================== ABI
extern "C" int desire_cost_bwd(const float* Yhat, const float* target, const float* mu_logvar, const float* obs,
                               const float* count, int M, int K, int T, int Tp, int Z, float* dYhat,
                               float* d_mu_logvar, desire_stream_t stream) {
  DESIRE_CHECK_ARG(Yhat && target && mu_logvar && obs && count && dYhat && d_mu_logvar && M >= 0 && K > 0 && T > 0 &&
                       Tp > 0 && Z > 0, "desire_cost_bwd: bad arguments");
  if (M == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)M * K * T * 2;
  DESIRE_LAUNCH(st, (cost_bwd_y_kernel<<<grid1d(total), 256, 0, st>>>(Yhat, target, obs, count, total, K, T, Tp, dYhat)));
  DESIRE_LAUNCH(st, (kld_bwd_kernel<<<grid1d((size_t)M * Z), 256, 0, st>>>(mu_logvar, obs, count, M, Z, Tp, d_mu_logvar)));
  return DESIRE_OK;
}

extern "C" int desire_readout_bwd(const float* hs, const float* dYhat, int R, int T, int H, const float* out_w,
                                  float* dhs, float* d_out_w, float* d_out_b, desire_stream_t stream) {
  DESIRE_CHECK_ARG(hs && dYhat && out_w && dhs && d_out_w && d_out_b && R >= 0 && T > 0 && H > 0,
                   "desire_readout_bwd: bad arguments");
  if (R == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rows = (size_t)R * T;
  DESIRE_CHECK_ARG(rows < (1u << 31), "desire_readout_bwd: R*T too large");
  DESIRE_LAUNCH(st, (readout_bwd_kernel<<<grid1d(rows * H), 256, 0, st>>>(dYhat, out_w, rows, H, dhs)));
  DESIRE_TRY(wgrad_tn(hs, H, dYhat, 2, d_out_w, 2, (int)rows, H, 2, st));
  DESIRE_TRY(colsum_acc(dYhat, 2, (int)rows, 2, d_out_b, st));
  return DESIRE_OK;
}

extern "C" size_t desire_gru_decode_bwd_workspace_bytes(int R, int H, int T) {
  const size_t r = (size_t)R;
  // xp[3H] dxp[3H] h0e[H] dh0[H] + bptt scratch
  return 2 * align_up(r * 3 * H * 4) + 2 * align_up(r * H * 4) + gru_bptt_ws_bytes(r, H, T);
}

extern "C" int desire_gru_decode_bwd(const float* x_z, const float* Hx, int ld_hx, int R, int K, int H, int T,
                                     const desire_gru_t* w, const float* hs, float* dhs, float* dx_z, float* dHx,
                                     int ld_dhx, const desire_gru_grad_t* g, void* ws, size_t ws_bytes,
                                     desire_stream_t stream) {
  DESIRE_CHECK_ARG(x_z && Hx && w && hs && dhs && dx_z && dHx && g && R >= 0 && K > 0 && T > 0 && H > 0 && R % K == 0,
                   "desire_gru_decode_bwd: bad arguments");
  if (!ws || ws_bytes < desire_gru_decode_bwd_workspace_bytes(R, H, T)) {
    set_error("desire_gru_decode_bwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  if (R == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t r = (size_t)R;
  Workspace W(ws, ws_bytes);
  float* xp = W.take<float>(r * 3 * H);
  float* dxp = W.take<float>(r * 3 * H);
  float* h0e = W.take<float>(r * H);
  float* dh0 = W.take<float>(r * H);
  const size_t rest = ws_bytes - W.off;
  void* bws = (char*)ws + W.off;
  PackWs pw{(char*)bws + ((rest - PACK_WS_BYTES) & ~(size_t)255), PACK_WS_BYTES};   // beyond the bptt buffers (its own pack area)
  // hoisted input projection, as in the forward
  DESIRE_TRY(sgemm(x_z, H, w->wg, 2 * H, false, w->bg, xp, 3 * H, R, 2 * H, H, DESIRE_ACT_NONE, false, st, pw));
  DESIRE_TRY(sgemm(x_z, H, w->wc, H, false, w->bc, xp + 2 * H, 3 * H, R, H, H, DESIRE_ACT_NONE, false, st, pw));
  DESIRE_CUDA(cudaMemsetAsync(dxp, 0, r * 3 * H * sizeof(float), st));
  DESIRE_CUDA(cudaMemsetAsync(dh0, 0, r * H * sizeof(float), st));
  DESIRE_LAUNCH(st, (expand_rows_bwd_kernel<<<grid1d(r * H), 256, 0, st>>>(Hx, K, ld_hx, r, H, h0e)));
  GruBptt a{};
  a.R = R; a.H = H; a.T = T; a.I = H;
  a.wg = w->wg; a.wc = w->wc;
  a.xp = xp; a.xp_rs = 3 * H; a.xp_ss = 0;
  a.hs = hs; a.hs_rs = (long)T * H; a.hs_ss = H;
  a.h0e = h0e;
  a.dhs = dhs; a.dhs_rs = (long)T * H; a.dhs_ss = H;
  a.dxp = dxp; a.dxp_rs = 3 * H; a.dxp_ss = 0;
  a.dh0 = dh0;
  a.dwg = g->wg; a.dwc = g->wc;
  DESIRE_TRY(gru_bptt(a, bws, rest, st));
  // input rows: dx_z = dxp_ru @ Wg_x^T + dxp_c @ Wc_x^T; weight / bias gradients of the input rows
  DESIRE_TRY(sgemm(dxp, 3 * H, w->wg, 2 * H, true, nullptr, dx_z, H, R, H, 2 * H, DESIRE_ACT_NONE, false, st, pw));
  DESIRE_TRY(sgemm(dxp + 2 * H, 3 * H, w->wc, H, true, nullptr, dx_z, H, R, H, H, DESIRE_ACT_NONE, true, st, pw));
  DESIRE_TRY(wgrad_tn(x_z, H, dxp, 3 * H, g->wg, 2 * H, R, H, 2 * H, st));
  DESIRE_TRY(wgrad_tn(x_z, H, dxp + 2 * H, 3 * H, g->wc, H, R, H, H, st));
  DESIRE_TRY(colsum_acc(dxp, 3 * H, R, 2 * H, g->bg, st));
  DESIRE_TRY(colsum_acc(dxp + 2 * H, 3 * H, R, H, g->bc, st));
  DESIRE_LAUNCH(st, (reduce_k_rows_kernel<<<grid1d((size_t)(R / K) * H), 256, 0, st>>>(dh0, (size_t)(R / K), K, H, dHx, ld_dhx)));
  return DESIRE_OK;
}

extern "C" size_t desire_gru_encode_bwd_workspace_bytes(int M, int T, int H) {
  const size_t m = (size_t)M;
  // xp, dxp [M,T,3H]; hs, dhs [M,T,H]; h0e, dh0 [M,H]
  return 2 * align_up(m * T * 3 * H * 4) + 2 * align_up(m * T * H * 4) + 2 * align_up(m * H * 4) + gru_bptt_ws_bytes(m, H, T);
}

extern "C" int desire_gru_encode_bwd(const float* traj, int M, int T, int H, const desire_gru_t* w, const float* dh,
                                     int ld_dh, const desire_gru_grad_t* g, void* ws, size_t ws_bytes,
                                     desire_stream_t stream) {
  DESIRE_CHECK_ARG(traj && w && dh && g && M >= 0 && T > 0 && H > 0 && H % 4 == 0, "desire_gru_encode_bwd: bad arguments");
  if (!ws || ws_bytes < desire_gru_encode_bwd_workspace_bytes(M, T, H)) {
    set_error("desire_gru_encode_bwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  if (M == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)M;
  Workspace W(ws, ws_bytes);
  float* xp = W.take<float>(m * T * 3 * H);
  float* dxp = W.take<float>(m * T * 3 * H);
  float* hs = W.take<float>(m * T * H);
  float* dhs = W.take<float>(m * T * H);
  float* h0e = W.take<float>(m * H);
  float* dh0 = W.take<float>(m * H);
  const size_t rest = ws_bytes - W.off;
  void* bws = (char*)ws + W.off;
  // forward recompute: all states with the FP32 recurrence of the forward path
  GruSeqArgs f{};
  f.R = M; f.H = H; f.T = T;
  f.traj = traj;
  f.wx_g = w->wg; f.wx_c = w->wc; f.bg = w->bg; f.bc = w->bc;
  f.w_g = w->wg + 2 * 2 * H;
  f.w_c = w->wc + 2 * H;
  f.h0 = nullptr; f.h0_div = 1;
  f.hs = hs; f.hs_row_stride = (long)T * H; f.hs_step_stride = H;
  DESIRE_TRY(gru_seq(f, st));
  DESIRE_LAUNCH(st, (xproj_traj_kernel<<<grid1d(m * T * 3 * H), 256, 0, st>>>(traj, m * T, H, w->wg, w->bg, w->wc, w->bc, xp)));
  DESIRE_CUDA(cudaMemsetAsync(dxp, 0, m * T * 3 * H * sizeof(float), st));
  DESIRE_CUDA(cudaMemsetAsync(dhs, 0, m * T * H * sizeof(float), st));
  DESIRE_CUDA(cudaMemsetAsync(h0e, 0, m * H * sizeof(float), st));
  DESIRE_CUDA(cudaMemsetAsync(dh0, 0, m * H * sizeof(float), st));
  DESIRE_CUDA(cudaMemcpy2DAsync(dhs + (size_t)(T - 1) * H, (size_t)T * H * sizeof(float), dh, (size_t)ld_dh * sizeof(float),
                                (size_t)H * sizeof(float), m, cudaMemcpyDeviceToDevice, st));
  GruBptt a{};
  a.R = M; a.H = H; a.T = T; a.I = 2;
  a.wg = w->wg; a.wc = w->wc;
  a.xp = xp; a.xp_rs = (long)T * 3 * H; a.xp_ss = 3 * H;
  a.hs = hs; a.hs_rs = (long)T * H; a.hs_ss = H;
  a.h0e = h0e;
  a.dhs = dhs; a.dhs_rs = (long)T * H; a.dhs_ss = H;
  a.dxp = dxp; a.dxp_rs = (long)T * 3 * H; a.dxp_ss = 3 * H;
  a.dh0 = dh0;
  a.dwg = g->wg; a.dwc = g->wc;
  DESIRE_TRY(gru_bptt(a, bws, rest, st));
  // input rows (x,y): dW_x = X^T @ dxp over all (m,t); biases = column sums
  const int rows = M * T;
  DESIRE_TRY(wgrad_tn(traj + 1, 3, dxp, 3 * H, g->wg, 2 * H, rows, 2, 2 * H, st));
  DESIRE_TRY(wgrad_tn(traj + 1, 3, dxp + 2 * H, 3 * H, g->wc, H, rows, 2, H, st));
  DESIRE_TRY(colsum_acc(dxp, 3 * H, rows, 2 * H, g->bg, st));
  DESIRE_TRY(colsum_acc(dxp + 2 * H, 3 * H, rows, H, g->bc, st));
  return DESIRE_OK;
}

extern "C" size_t desire_mask_softmax_bwd_workspace_bytes(int R, int H) {
  return 3 * align_up((size_t)R * H * sizeof(float)) + PACK_WS_BYTES;
}

extern "C" int desire_mask_softmax_bwd(const float* xr, int R, int S2, int H, int K, const float* w, const float* b,
                                       const float* Hx, int ld_hx, const float* dx_z, float* dxr, float* dHx,
                                       int ld_dhx, float* dw, float* db, void* ws, size_t ws_bytes,
                                       desire_stream_t stream) {
  DESIRE_CHECK_ARG(xr && w && b && Hx && dx_z && dxr && dHx && dw && db && R >= 0 && K > 0 && R % K == 0,
                   "desire_mask_softmax_bwd: bad arguments");
  if (!ws || ws_bytes < desire_mask_softmax_bwd_workspace_bytes(R, H)) {
    set_error("desire_mask_softmax_bwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  if (R == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace W(ws, ws_bytes);
  float* logits = W.take<float>((size_t)R * H);
  float* dl = W.take<float>((size_t)R * H);
  float* gb = W.take<float>((size_t)R * H);
  PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
  DESIRE_TRY(sgemm(xr, S2, w, H, false, b, logits, H, R, H, S2, DESIRE_ACT_RELU, false, st, pw));
  DESIRE_LAUNCH(st, (softmax_gate_bwd_kernel<<<grid1d((size_t)R * 32), 256, 0, st>>>(logits, dx_z, R, H, K, Hx, ld_hx, dl, gb)));
  DESIRE_LAUNCH(st, (reduce_k_rows_kernel<<<grid1d((size_t)(R / K) * H), 256, 0, st>>>(gb, (size_t)(R / K), K, H, dHx, ld_dhx)));
  DESIRE_TRY(sgemm(dl, H, w, H, true, nullptr, dxr, S2, R, S2, H, DESIRE_ACT_NONE, false, st, pw));
  DESIRE_TRY(wgrad_tn(xr, S2, dl, H, dw, H, R, S2, H, st));
  DESIRE_TRY(colsum_acc(dl, H, R, H, db, st));
  return DESIRE_OK;
}

extern "C" int desire_reparam_bwd(const float* mu_logvar, const float* eps, const float* dz, int M, int K, int Z,
                                  float* d_mu_logvar, desire_stream_t stream) {
  DESIRE_CHECK_ARG(mu_logvar && eps && dz && d_mu_logvar && M >= 0 && K > 0 && Z > 0, "desire_reparam_bwd: bad arguments");
  if (M == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DESIRE_LAUNCH(st, (reparam_bwd_kernel<<<grid1d((size_t)M * Z), 256, 0, st>>>(mu_logvar, eps, dz, M, K, Z, d_mu_logvar)));
  return DESIRE_OK;
}

// ---- CVAE decoder backward, rows in chunks (the deconv3 col matrix of the recompute is 205 KB per row)
static const int DEC_BWD_CHUNK = 4096;
static size_t dec_bwd_row_floats() {
  // col 51200 | y1 2048 a1 2048 | y2 4096 a2 4096 | y3 8192 a3 8192 | y4 1024 | g3 8192 g2 4096 g1 2048 g4 1024
  return 51200 + 2 * 2048 + 2 * 4096 + 2 * 8192 + 1024 + 8192 + 4096 + 2048 + 1024;
}
static size_t dec_bwd_wpack_bytes(int rc, int Z) {
  // packed B operands of the tcgen05 weight gradients: a2 [rc*64,64], a1 [rc*16,128], z [rc,Z]
  size_t a = wgrad_tc_pack_bytes(rc * 64, 64), b = wgrad_tc_pack_bytes(rc * 16, 128), c = wgrad_tc_pack_bytes(rc, Z);
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}
extern "C" size_t desire_cvae_decode_bwd_workspace_bytes(int R, int Z) {
  const size_t rc = R < DEC_BWD_CHUNK ? R : DEC_BWD_CHUNK;
  return align_up(rc * dec_bwd_row_floats() * 4) + 16 * 256 + PACK_WS_BYTES + dec_bwd_wpack_bytes((int)rc, Z);
}

extern "C" int desire_cvae_decode_bwd(const float* z, int R, int Z, const desire_cvae_dec_t* w, const float* dxr,
                                      float* dz, const desire_cvae_dec_grad_t* g, void* ws, size_t ws_bytes,
                                      desire_stream_t stream) {
  DESIRE_CHECK_ARG(z && w && dxr && dz && g && R >= 0 && Z > 0, "desire_cvae_decode_bwd: bad arguments");
  if (!ws || ws_bytes < desire_cvae_decode_bwd_workspace_bytes(R, Z)) {
    set_error("desire_cvae_decode_bwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int r0 = 0; r0 < R; r0 += DEC_BWD_CHUNK) {
    const int rc = (R - r0) < DEC_BWD_CHUNK ? (R - r0) : DEC_BWD_CHUNK;
    const size_t n = (size_t)rc;
    Workspace W(ws, ws_bytes);
    float* col = W.take<float>(n * 51200);
    float* y1 = W.take<float>(n * 2048);
    float* a1 = W.take<float>(n * 2048);
    float* y2 = W.take<float>(n * 4096);
    float* a2 = W.take<float>(n * 4096);
    float* y3 = W.take<float>(n * 8192);
    float* a3 = W.take<float>(n * 8192);
    float* y4 = W.take<float>(n * 1024);
    float* g3 = W.take<float>(n * 8192);
    float* g2 = W.take<float>(n * 4096);
    float* g1 = W.take<float>(n * 2048);
    float* g4 = W.take<float>(n * 1024);
    PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
    const size_t wpb = dec_bwd_wpack_bytes(rc, Z);
    PackWs wp{W.take<char>(wpb), wpb};
    if (!pw.p || !wp.p) {
      set_error("desire_cvae_decode_bwd: workspace too small");
      return DESIRE_ERR_WORKSPACE;
    }
    const float* zc = z + (size_t)r0 * Z;
    // ---------------- forward recompute (GEMM -> the forward's fused col2im + bias + per-row BN + activation kernel,
    // which here also stores the pre-BN values the backward needs)
    DESIRE_TRY(sgemm(zc, Z, w->d1.w, Z, true, nullptr, col, 2048, rc, 2048, Z, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(colbn_act(col, rc, 1, 4, 4, 1, 0, 128, w->d1.b, w->d1.gamma, w->d1.beta, DESIRE_ACT_ELU, a1, st, y1));
    DESIRE_TRY(sgemm(a1, 128, w->d2.w, 128, true, nullptr, col, 1600, rc * 16, 1600, 128, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(colbn_act(col, rc, 4, 8, 5, 1, 0, 64, w->d2.b, w->d2.gamma, w->d2.beta, DESIRE_ACT_ELU, a2, st, y2));
    DESIRE_TRY(sgemm(a2, 64, w->d3.w, 64, true, nullptr, col, 800, rc * 64, 800, 64, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(colbn_act(col, rc, 8, 16, 5, 2, 1, 32, w->d3.b, w->d3.gamma, w->d3.beta, DESIRE_ACT_ELU, a3, st, y3));
    DESIRE_TRY(sgemm(a3, 32, w->d4.w, 32, true, nullptr, col, 25, rc * 256, 25, 32, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(col2im_gather(col, rc, 16, 32, 5, 2, 1, 1, w->d4.b, y4, st));
    // ---------------- backward
    DESIRE_CUDA(cudaMemcpyAsync(g4, dxr + (size_t)r0 * 1024, n * 1024 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // layer 4: 16x16x32 -> 32x32x1, BN + sigmoid
    DESIRE_TRY(bn_row_bwd(y4, rc, 1024, 1, w->d4.gamma, w->d4.beta, DESIRE_ACT_SIGMOID, g4, g->d4.gamma, g->d4.beta, st));
    Im2col i4{32, 32, 1, 16, 16, 5, 5, 2, 1, 1};
    DESIRE_TRY(wgrad_tn_im2col(g4, i4, a3, 32, g->d4.w, 32, rc * 256, 25, 32, st));
    DESIRE_TRY(sgemm_im2col(g4, i4, w->d4.w, 32, nullptr, g3, 32, rc * 256, 32, 25, DESIRE_ACT_NONE, st, pw));
    // layer 3: 8x8x64 -> 16x16x32
    DESIRE_TRY(bn_row_bwd(y3, rc, 256, 32, w->d3.gamma, w->d3.beta, DESIRE_ACT_ELU, g3, g->d3.gamma, g->d3.beta, st));
    Im2col i3{16, 16, 32, 8, 8, 5, 5, 2, 1, 1};
    DESIRE_TRY(wgrad_tn_im2col(g3, i3, a2, 64, g->d3.w, 64, rc * 64, 800, 64, st, wp));
    DESIRE_TRY(sgemm_im2col(g3, i3, w->d3.w, 64, nullptr, g2, 64, rc * 64, 64, 800, DESIRE_ACT_NONE, st, pw));
    // layer 2: 4x4x128 -> 8x8x64 (VALID)
    DESIRE_TRY(bn_row_bwd(y2, rc, 64, 64, w->d2.gamma, w->d2.beta, DESIRE_ACT_ELU, g2, g->d2.gamma, g->d2.beta, st));
    Im2col i2{8, 8, 64, 4, 4, 5, 5, 1, 0, 0};
    DESIRE_TRY(wgrad_tn_im2col(g2, i2, a1, 128, g->d2.w, 128, rc * 16, 1600, 128, st, wp));
    DESIRE_TRY(sgemm_im2col(g2, i2, w->d2.w, 128, nullptr, g1, 128, rc * 16, 128, 1600, DESIRE_ACT_NONE, st, pw));
    // layer 1: 1x1xZ -> 4x4x128 (a plain GEMM: y1[r,(y,x,o)] = z[r,:] . W[(y,x,o),:])
    DESIRE_TRY(bn_row_bwd(y1, rc, 16, 128, w->d1.gamma, w->d1.beta, DESIRE_ACT_ELU, g1, g->d1.gamma, g->d1.beta, st));
    DESIRE_TRY(wgrad_tn(g1, 2048, zc, Z, g->d1.w, Z, rc, 2048, Z, st, wp));
    DESIRE_TRY(sgemm(g1, 2048, w->d1.w, Z, false, nullptr, dz + (size_t)r0 * Z, Z, rc, Z, 2048, DESIRE_ACT_NONE, false, st, pw));
  }
  return DESIRE_OK;
}

// ---- CVAE encoder backward
static const int ENC_BWD_CHUNK = 4096;
static size_t enc_bwd_row_floats() {
  // col max(64*800, 16*1600, 256*25) = 51200 | y1 a1 g1 8192 each | y2 a2 g2 4096 each | y3 a3 g3 2048 each
  return 51200 + 3 * 8192 + 3 * 4096 + 3 * 2048;
}
extern "C" size_t desire_cvae_encode_bwd_workspace_bytes(int M, int Z) {
  (void)Z;
  const size_t mc = M < ENC_BWD_CHUNK ? M : ENC_BWD_CHUNK;
  return align_up(mc * enc_bwd_row_floats() * 4) + 16 * 256 + PACK_WS_BYTES;
}

extern "C" int desire_cvae_encode_bwd(const float* v, int M, int Z, const desire_cvae_enc_t* w, const float* d_mu_logvar,
                                      float* dv, const desire_cvae_enc_grad_t* g, void* ws, size_t ws_bytes,
                                      desire_stream_t stream) {
  DESIRE_CHECK_ARG(v && w && d_mu_logvar && dv && g && M >= 0 && Z > 0, "desire_cvae_encode_bwd: bad arguments");
  if (!ws || ws_bytes < desire_cvae_encode_bwd_workspace_bytes(M, Z)) {
    set_error("desire_cvae_encode_bwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int m0 = 0; m0 < M; m0 += ENC_BWD_CHUNK) {
    const int mc = (M - m0) < ENC_BWD_CHUNK ? (M - m0) : ENC_BWD_CHUNK;
    const size_t n = (size_t)mc;
    Workspace W(ws, ws_bytes);
    float* col = W.take<float>(n * 51200);
    float* y1 = W.take<float>(n * 8192);
    float* a1 = W.take<float>(n * 8192);
    float* g1 = W.take<float>(n * 8192);
    float* y2 = W.take<float>(n * 4096);
    float* a2 = W.take<float>(n * 4096);
    float* g2 = W.take<float>(n * 4096);
    float* y3 = W.take<float>(n * 2048);
    float* a3 = W.take<float>(n * 2048);
    float* g3 = W.take<float>(n * 2048);
    PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
    if (!pw.p) {
      set_error("desire_cvae_encode_bwd: workspace too small");
      return DESIRE_ERR_WORKSPACE;
    }
    const float* x = v + (size_t)m0 * 1024;
    const float* dp = d_mu_logvar + (size_t)m0 * 2 * Z;
    // ---------------- forward recompute
    Im2col c1{32, 32, 1, 16, 16, 5, 5, 2, 1, 1};
    Im2col c2{16, 16, 32, 8, 8, 5, 5, 2, 1, 1};
    Im2col c3{8, 8, 64, 4, 4, 5, 5, 1, 0, 0};
    DESIRE_TRY(sgemm_im2col(x, c1, w->c1.w, 32, w->c1.b, y1, 32, mc * 256, 32, 25, DESIRE_ACT_NONE, st, pw));
    DESIRE_TRY(bn_row_fwd(y1, mc, 256, 32, w->c1.gamma, w->c1.beta, DESIRE_ACT_ELU, a1, st));
    DESIRE_TRY(sgemm_im2col(a1, c2, w->c2.w, 64, w->c2.b, y2, 64, mc * 64, 64, 800, DESIRE_ACT_NONE, st, pw));
    DESIRE_TRY(bn_row_fwd(y2, mc, 64, 64, w->c2.gamma, w->c2.beta, DESIRE_ACT_ELU, a2, st));
    DESIRE_TRY(sgemm_im2col(a2, c3, w->c3.w, 128, w->c3.b, y3, 128, mc * 16, 128, 1600, DESIRE_ACT_NONE, st, pw));
    DESIRE_TRY(bn_row_fwd(y3, mc, 16, 128, w->c3.gamma, w->c3.beta, DESIRE_ACT_ELU, a3, st));
    // ---------------- backward
    // fc 2048 -> 2Z
    DESIRE_TRY(wgrad_tn(a3, 2048, dp, 2 * Z, g->fc_w, 2 * Z, mc, 2048, 2 * Z, st));
    DESIRE_TRY(colsum_acc(dp, 2 * Z, mc, 2 * Z, g->fc_b, st));
    DESIRE_TRY(sgemm(dp, 2 * Z, w->fc_w, 2 * Z, true, nullptr, g3, 2048, mc, 2048, 2 * Z, DESIRE_ACT_NONE, false, st, pw));
    // conv3 VALID 8x8x64 -> 4x4x128
    DESIRE_TRY(bn_row_bwd(y3, mc, 16, 128, w->c3.gamma, w->c3.beta, DESIRE_ACT_ELU, g3, g->c3.gamma, g->c3.beta, st));
    DESIRE_TRY(wgrad_tn_im2col(a2, c3, g3, 128, g->c3.w, 128, mc * 16, 1600, 128, st));
    DESIRE_TRY(sgemm(g3, 128, w->c3.w, 128, true, nullptr, col, 1600, mc * 16, 1600, 128, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(col2im_gather(col, mc, 4, 8, 5, 1, 0, 64, nullptr, g2, st));
    // conv2 SAME/2 16x16x32 -> 8x8x64
    DESIRE_TRY(bn_row_bwd(y2, mc, 64, 64, w->c2.gamma, w->c2.beta, DESIRE_ACT_ELU, g2, g->c2.gamma, g->c2.beta, st));
    DESIRE_TRY(wgrad_tn_im2col(a1, c2, g2, 64, g->c2.w, 64, mc * 64, 800, 64, st));
    DESIRE_TRY(sgemm(g2, 64, w->c2.w, 64, true, nullptr, col, 800, mc * 64, 800, 64, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(col2im_gather(col, mc, 8, 16, 5, 2, 1, 32, nullptr, g1, st));
    // conv1 SAME/2 32x32x1 -> 16x16x32
    DESIRE_TRY(bn_row_bwd(y1, mc, 256, 32, w->c1.gamma, w->c1.beta, DESIRE_ACT_ELU, g1, g->c1.gamma, g->c1.beta, st));
    DESIRE_TRY(wgrad_tn_im2col(x, c1, g1, 32, g->c1.w, 32, mc * 256, 25, 32, st));
    DESIRE_TRY(sgemm(g1, 32, w->c1.w, 32, true, nullptr, col, 25, mc * 256, 25, 32, DESIRE_ACT_NONE, false, st, pw));
    DESIRE_TRY(col2im_gather(col, mc, 16, 32, 5, 2, 1, 1, nullptr, dv + (size_t)m0 * 1024, st));
  }
  return DESIRE_OK;
}

extern "C" int desire_fc_bwd(const float* A, int lda, const float* W, int ldw, const float* Cout, int ldc, float* dC,
                             int lddc, int M, int N, int K, int act, float* dA, int ldda, int accumulate_dA, float* dW,
                             int lddw, float* db, desire_stream_t stream) {
  DESIRE_CHECK_ARG(A && W && dC && M >= 0 && N > 0 && K > 0, "desire_fc_bwd: bad arguments");
  DESIRE_CHECK_ARG(act == DESIRE_ACT_NONE || Cout, "desire_fc_bwd: an activation needs the forward output");
  if (M == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (act != DESIRE_ACT_NONE)
    DESIRE_LAUNCH(st, (act_bwd_post_kernel<<<grid1d((size_t)M * N), 256, 0, st>>>(Cout, ldc, dC, lddc, (size_t)M, N, act)));
  if (dW) DESIRE_TRY(wgrad_tn(A, lda, dC, lddc, dW, lddw, M, K, N, st));
  if (db) DESIRE_TRY(colsum_acc(dC, lddc, M, N, db, st));
  if (dA) DESIRE_TRY(sgemm(dC, lddc, W, ldw, true, nullptr, dA, ldda, M, K, N, DESIRE_ACT_NONE, accumulate_dA != 0, st));
  return DESIRE_OK;
}

extern "C" int desire_sumsq_fwd(const float* g, long n, float* out, int accumulate, desire_stream_t stream) {
  DESIRE_CHECK_ARG(g && out && n >= 0, "desire_sumsq_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DESIRE_LAUNCH(st, (sumsq_kernel<<<1, 1024, 0, st>>>(g, (size_t)n, out, accumulate ? 1 : 0)));
  return DESIRE_OK;
}

extern "C" int desire_adam_step(float* p, const float* g, float* m, float* v, long n, const float* sumsq, float lr,
                                float beta1, float beta2, float eps, int step, float clip, float grad_scale,
                                desire_stream_t stream) {
  DESIRE_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "desire_adam_step: bad arguments");
  DESIRE_CHECK_ARG(clip <= 0.f || sumsq, "desire_adam_step: clipping needs the sum of squares");
  if (n == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  DESIRE_LAUNCH(st, (adam_kernel<<<grid1d((size_t)n), 256, 0, st>>>(p, g, m, v, (size_t)n, sumsq, (float)lr_t, beta1, beta2,
                                                                   eps, clip, grad_scale)));
  return DESIRE_OK;
}

// generic weight-gradient product (the backward twin of desire_fc_fwd's W): dW[K,N] (lddw) += A[M,K]^T @ dC[M,N]
extern "C" size_t desire_wgrad_workspace_bytes(int M, int N) { return wgrad_tc_pack_bytes(M, N); }
extern "C" int desire_wgrad_tn(const float* A, int lda, const float* dC, int lddc, float* dW, int lddw, int M, int N, int K,
                               void* ws, size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(A && dC && dW && M >= 0 && N > 0 && K > 0 && lda >= K && lddc >= N && lddw >= N,
                   "desire_wgrad_tn: bad arguments");
  return wgrad_tn(A, lda, dC, lddc, dW, lddw, M, K, N, (cudaStream_t)stream, PackWs{ws, ws_bytes});
}

// ---- generic transposed convolution (un-fused): the operator behind utils/convolutional_vae_util.py:deconv2d for
// arbitrary square geometries (the CVAE decoder's four layers have their own fused kernels in cvae.cu/deconv_tc.cu).
//   y = act( [BN_row]( conv2d_transpose(x, w) + bias ) ),  x [R,Hin,Hin,Cin] NHWC, w [k,k,Cout,Cin], y [R,Hout,Hout,Cout]
namespace {
__global__ void act_inplace_kernel(float* __restrict__ y, size_t n, int act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = act_apply(y[i], act);
}
int deconv_out_size(int in, int k, int s, int same) { return same ? in * s : (in - 1) * s + k; }
}  // namespace

extern "C" size_t desire_deconv2d_workspace_bytes(int R, int Hin, int Cin, int k, int stride, int same, int Cout) {
  (void)Cin;
  const size_t Hout = deconv_out_size(Hin, k, stride, same);
  return align_up((size_t)R * Hin * Hin * k * k * Cout * 4) + align_up((size_t)R * Hout * Hout * Cout * 4) + PACK_WS_BYTES;
}

extern "C" int desire_deconv2d_fwd(const float* x, int R, int Hin, int Cin, const float* w, int k, int stride, int same,
                                   int Cout, const float* bias, const float* gamma, const float* beta, int act, float* y,
                                   void* ws, size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(x && w && y && R >= 0 && Hin > 0 && Cin > 0 && k > 0 && stride > 0 && Cout > 0,
                   "desire_deconv2d_fwd: bad arguments");
  DESIRE_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "desire_deconv2d_fwd: batch-norm needs both gamma and beta");
  if (!ws || ws_bytes < desire_deconv2d_workspace_bytes(R, Hin, Cin, k, stride, same, Cout)) {
    set_error("desire_deconv2d_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  if (R == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int Hout = deconv_out_size(Hin, k, stride, same);
  const int full = (Hin - 1) * stride + k;
  const int pad = same ? (full - Hout > 0 ? (full - Hout) / 2 : 0) : 0;      // TF: conv_transpose is the gradient of a conv
  Workspace W(ws, ws_bytes);                                                  // whose SAME padding puts total//2 before
  float* col = W.take<float>((size_t)R * Hin * Hin * k * k * Cout);
  float* ypre = W.take<float>((size_t)R * Hout * Hout * Cout);
  PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
  const long rows = (long)R * Hin * Hin;
  DESIRE_CHECK_ARG(rows < (1L << 31), "desire_deconv2d_fwd: too many input positions");
  DESIRE_TRY(sgemm(x, Cin, w, Cin, true, nullptr, col, k * k * Cout, (int)rows, k * k * Cout, Cin, DESIRE_ACT_NONE, false, st, pw));
  if (gamma) {
    DESIRE_TRY(col2im_gather(col, R, Hin, Hout, k, stride, pad, Cout, bias, ypre, st));
    DESIRE_TRY(bn_row_fwd(ypre, R, Hout * Hout, Cout, gamma, beta, act, y, st));
  } else {
    DESIRE_TRY(col2im_gather(col, R, Hin, Hout, k, stride, pad, Cout, bias, y, st));
    const size_t n = (size_t)R * Hout * Hout * Cout;
    if (act != DESIRE_ACT_NONE) DESIRE_LAUNCH(st, (act_inplace_kernel<<<grid1d(n), 256, 0, st>>>(y, n, act)));
  }
  return DESIRE_OK;
}
