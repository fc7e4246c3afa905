// Stage 2 — IOC ranking & refinement (absent in the reference, marker
// model/model.py:312-313): scene CNN, bilinear scene-feature gather, log-polar social pooling,
// Decoder-2 GRU with per-step scoring, regression refinement.
#include "common.cuh"

using namespace desire;

namespace {

// ---- bilinear gather: one warp per point, lanes over channels (each tap is one coalesced row of Cs floats)
__global__ void scene_gather_kernel(const float* __restrict__ fmap, int Hm, int Wm, int Cs,
                                    const float* __restrict__ pos, long pos_stride, long npts, int rows_per_scene,
                                    float* __restrict__ out, int ld_out) {
  const long pt = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pt >= npts) return;
  const long b = pt / rows_per_scene;
  const float x = __ldg(pos + pt * pos_stride), y = __ldg(pos + pt * pos_stride + 1);
  const float wm1 = (float)(Wm - 1), hm1 = (float)(Hm - 1);
  const float px = fminf(fmaxf(__fmul_rn(x, wm1), 0.f), wm1);
  const float py = fminf(fmaxf(__fmul_rn(y, hm1), 0.f), hm1);
  const int x0 = (int)floorf(px), y0 = (int)floorf(py);
  const int x1 = min(x0 + 1, Wm - 1), y1 = min(y0 + 1, Hm - 1);
  const float fx = px - (float)x0, fy = py - (float)y0;
  const float* base = fmap + (size_t)b * Hm * Wm * Cs;
  const float* p00 = base + ((size_t)y0 * Wm + x0) * Cs;
  const float* p01 = base + ((size_t)y0 * Wm + x1) * Cs;
  const float* p10 = base + ((size_t)y1 * Wm + x0) * Cs;
  const float* p11 = base + ((size_t)y1 * Wm + x1) * Cs;
  float* o = out + pt * (long)ld_out;
  for (int c = lane; c < Cs; c += 32) {
    float v00 = __ldg(p00 + c), v01 = __ldg(p01 + c), v10 = __ldg(p10 + c), v11 = __ldg(p11 + c);
    float top = v00 + fx * (v01 - v00);
    float bot = v10 + fx * (v11 - v10);
    o[c] = top + fy * (bot - top);
  }
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

// ---- social pooling: one CTA per row (b,i,k); accumulate the neighbours' hidden vectors into a
// [G,H] shared-memory tile (sequential over neighbours, threads over H => no atomics), then one
// coalesced store of the averaged tile.  The G*H*4-byte write per row is the algorithmic traffic.
__global__ void __launch_bounds__(128) social_pool_kernel(const float* __restrict__ pos, long pos_stride,
                                                          const float* __restrict__ h, int ld_h,
                                                          const float* __restrict__ obs, int Tp, int N, int K, int H,
                                                          int n_rad, int n_ang, const float* __restrict__ r2_edges,
                                                          const float* __restrict__ dirs,
                                                          float* __restrict__ pooled) {
  extern __shared__ __align__(16) float sm[];
  const int G = n_rad * n_ang;
  float* acc = sm;                         // [G*H]
  float* cnt = acc + G * H;                // [G]
  float* tab = cnt + G;                    // [n_rad+1 + 2*n_ang]
  int* bins = (int*)(tab + n_rad + 1 + 2 * n_ang);   // [N]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long row = blockIdx.x;             // (b*N + i)*K + k
  const int k = (int)(row % K);
  const long bi = row / K;
  const int i = (int)(bi % N);
  const long b = bi / N;
  for (int e = tid; e < n_rad + 1; e += nthr) tab[e] = __ldg(r2_edges + e);
  for (int e = tid; e < 2 * n_ang; e += nthr) tab[n_rad + 1 + e] = __ldg(dirs + e);
  for (int e = tid; e < G * H; e += nthr) acc[e] = 0.f;
  __syncthreads();
  const float xi = __ldg(pos + row * pos_stride), yi = __ldg(pos + row * pos_stride + 1);
  for (int j = tid; j < N; j += nthr) {
    int g = -1;
    if (j != i && __ldg(obs + ((size_t)(b * N + j) * Tp) * 3) != 0.f) {
      const long rj = (b * N + j) * K + k;
      const float dx = __ldg(pos + rj * pos_stride) - xi, dy = __ldg(pos + rj * pos_stride + 1) - yi;
      g = logpolar_bin(dx, dy, tab, n_rad, tab + n_rad + 1, n_ang);
    }
    bins[j] = g;
  }
  __syncthreads();
  if (tid < G) {
    float c = 0.f;
    for (int j = 0; j < N; ++j) c += (bins[j] == tid) ? 1.f : 0.f;
    cnt[tid] = c;
  }
  for (int j = 0; j < N; ++j) {
    const int g = bins[j];
    if (g < 0) continue;
    const float* hj = h + ((b * N + j) * K + k) * (long)ld_h;
    for (int c = tid; c < H; c += nthr) acc[g * H + c] += __ldg(hj + c);
  }
  __syncthreads();
  float* o = pooled + row * (long)G * H;
  for (int e = tid; e < G * H; e += nthr) o[e] = acc[e] / fmaxf(cnt[e / H], 1.f);
}

// ---- velocity fc: Xs[(r,t), 0:Fv] = relu((Y_t - Y_{t-1}) @ vel_w + vel_b), Y_{-1} = last observed position
__global__ void vel_fc_kernel(const float* __restrict__ Y, const float* __restrict__ obs, int Tp, long R, int K, int T,
                              int Fv, const float* __restrict__ w, const float* __restrict__ bias,
                              float* __restrict__ Xs, int ld) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * T * Fv) return;
  const int f = (int)(idx % Fv);
  const long rt = idx / Fv;
  const int t = (int)(rt % T);
  const long r = rt / T;
  float px, py;
  if (t == 0) {
    const float* last = obs + ((r / K) * Tp + (Tp - 1)) * 3;
    px = __ldg(last + 1);
    py = __ldg(last + 2);
  } else {
    px = Y[(rt - 1) * 2];
    py = Y[(rt - 1) * 2 + 1];
  }
  const float vx = Y[rt * 2] - px, vy = Y[rt * 2 + 1] - py;
  float v = fmaf(vy, __ldg(w + Fv + f), fmaf(vx, __ldg(w + f), __ldg(bias + f)));
  Xs[rt * ld + f] = fmaxf(v, 0.f);
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int ncols, long rows, float* __restrict__ dst, int ld) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * ncols) return;
  dst[(idx / ncols) * ld + idx % ncols] = src[idx];
}

__global__ void expand_rows_kernel(const float* __restrict__ src, int ld_src, int K, int H, long R,
                                   float* __restrict__ dst) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * H) return;
  dst[idx] = __ldg(src + (idx / H / K) * ld_src + idx % H);
}

// score[r] (+)= h2[r,:] . w + b     one warp per row
__global__ void score_kernel(const float* __restrict__ h2, long R, int H, const float* __restrict__ w,
                             const float* __restrict__ b, float* __restrict__ score, int first) {
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= R) return;
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s = fmaf(h2[warp * H + c], __ldg(w + c), s);
  s = warp_sum(s);
  if (lane == 0) score[warp] = (first ? 0.f : score[warp]) + s + __ldg(b);
}

inline unsigned blocks(long n, int t) { return (unsigned)((n + t - 1) / t); }

int social_pool_launch(const float* pos, long pos_stride, const float* h, int ld_h, const float* obs, int Tp, int B,
                       int N, int K, int H, int n_rad, int n_ang, const float* r2_edges, const float* dirs,
                       float* pooled, cudaStream_t st) {
  const int G = n_rad * n_ang;
  size_t smem = ((size_t)G * H + G + n_rad + 1 + 2 * n_ang) * sizeof(float) + (size_t)N * sizeof(int);
  DESIRE_CHECK_ARG(G <= 128, "social_pool: at most 128 bins");
  DESIRE_CHECK_ARG(smem <= 227 * 1024, "social_pool: G*H tile does not fit shared memory");
  DESIRE_CUDA(cudaFuncSetAttribute(social_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  long rows = (long)B * N * K;
  if (rows == 0) return DESIRE_OK;
  social_pool_kernel<<<(unsigned)rows, 128, smem, st>>>(pos, pos_stride, h, ld_h, obs, Tp, N, K, H, n_rad, n_ang,
                                                        r2_edges, dirs, pooled);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------ scene CNN
extern "C" size_t desire_scene_cnn_workspace_bytes(int B, int Hi, int Wi) {
  size_t Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  return align_up((size_t)B * Ho * Wo * 16 * 4) + align_up((size_t)B * Ho * Wo * 32 * 4) + PACK_WS_BYTES;
}

extern "C" int desire_scene_cnn_fwd(const float* img, int B, int Hi, int Wi, int Cs, const desire_scene_cnn_t* w,
                                    float* fmap, void* ws, size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(img && w && fmap && B >= 0 && Hi > 0 && Wi > 0 && Cs > 0, "desire_scene_cnn_fwd: bad arguments");
  if (!ws || ws_bytes < desire_scene_cnn_workspace_bytes(B, Hi, Wi)) {
    set_error("desire_scene_cnn_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  Workspace W(ws, ws_bytes);
  float* f1 = W.take<float>((size_t)B * Ho * Wo * 16);
  float* f2 = W.take<float>((size_t)B * Ho * Wo * 32);
  PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
  // TF SAME: total pad = max((out-1)*s + k - in, 0), before = total/2
  const int pt1 = max((Ho - 1) * 2 + 5 - Hi, 0) / 2, pl1 = max((Wo - 1) * 2 + 5 - Wi, 0) / 2;
  // one image at a time keeps gridDim.y legal for any map size
  for (int b = 0; b < B; ++b) {
    const size_t px = (size_t)Ho * Wo;
    Im2col g1{Hi, Wi, 3, Ho, Wo, 5, 5, 2, pt1, pl1};
    DESIRE_TRY(sgemm_im2col(img + (size_t)b * Hi * Wi * 3, g1, w->c1_w, 16, w->c1_b, f1 + b * px * 16, 16, (int)px, 16,
                            75, DESIRE_ACT_RELU, st, pw));
    Im2col g2{Ho, Wo, 16, Ho, Wo, 5, 5, 1, 2, 2};
    DESIRE_TRY(sgemm_im2col(f1 + b * px * 16, g2, w->c2_w, 32, w->c2_b, f2 + b * px * 32, 32, (int)px, 32, 400,
                            DESIRE_ACT_RELU, st, pw));
    Im2col g3{Ho, Wo, 32, Ho, Wo, 5, 5, 1, 2, 2};
    DESIRE_TRY(sgemm_im2col(f2 + b * px * 32, g3, w->c3_w, Cs, w->c3_b, fmap + b * px * Cs, Cs, (int)px, Cs, 800,
                            DESIRE_ACT_RELU, st, pw));
  }
  return DESIRE_OK;
}

extern "C" int desire_scene_gather_fwd(const float* fmap, int B, int Hm, int Wm, int Cs, const float* pos,
                                       long pos_stride, int rows_per_scene, float* out, int ld_out,
                                       desire_stream_t stream) {
  DESIRE_CHECK_ARG(fmap && pos && out && B >= 0 && Hm > 0 && Wm > 0 && Cs > 0 && rows_per_scene >= 0 && ld_out >= Cs,
                   "desire_scene_gather_fwd: bad arguments");
  long npts = (long)B * rows_per_scene;
  if (npts == 0) return DESIRE_OK;
  scene_gather_kernel<<<blocks(npts * 32, 256), 256, 0, (cudaStream_t)stream>>>(fmap, Hm, Wm, Cs, pos, pos_stride, npts,
                                                                                 rows_per_scene, out, ld_out);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_social_pool_fwd(const float* pos, long pos_stride, const float* h, int ld_h, const float* obs,
                                      int Tp, int B, int N, int K, int H, int n_rad, int n_ang, const float* r2_edges,
                                      const float* dirs, float* pooled, desire_stream_t stream) {
  DESIRE_CHECK_ARG(pos && h && obs && r2_edges && dirs && pooled && B >= 0 && N > 0 && K > 0 && H > 0 && n_rad > 0 &&
                       n_ang > 0 && ld_h >= H,
                   "desire_social_pool_fwd: bad arguments");
  return social_pool_launch(pos, pos_stride, h, ld_h, obs, Tp, B, N, K, H, n_rad, n_ang, r2_edges, dirs, pooled,
                            (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------ IOC loop
namespace {
struct IocLayout {
  size_t Xs, XP, pooled, fsp, h2, wsp3, pack, total;
};
IocLayout ioc_layout(const desire_ioc_dims_t* d) {
  const size_t R = (size_t)d->B * d->N * d->K, T = d->Tf, H = d->H;
  const size_t Dst = d->Fv + d->Cs + 2 * d->C, G = (size_t)d->n_rad * d->n_ang;
  IocLayout L;
  size_t off = 0;
  L.Xs = off; off += align_up(R * T * Dst * 4);
  L.XP = off; off += align_up(R * T * 3 * H * 4);
  L.pooled = off; off += align_up(R * G * H * 4);
  L.fsp = off; off += align_up(R * H * 4);
  L.h2 = off; off += align_up(R * H * 4);
  L.wsp3 = off; off += align_up(H * 3 * H * 4);
  L.pack = off; off += PACK_WS_BYTES;
  L.total = off;
  return L;
}
}  // namespace

extern "C" size_t desire_ioc_workspace_bytes(const desire_ioc_dims_t* d) { return d ? ioc_layout(d).total : 0; }

extern "C" int desire_ioc_fwd(const desire_ioc_dims_t* d, const desire_ioc_t* w, const float* fmap, const float* obs,
                              int Tp, const float* Hx, int ld_hx, const float* fpool, float* Y, float* scores, void* ws,
                              size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(d && w && fmap && obs && Hx && fpool && Y && scores, "desire_ioc_fwd: null argument");
  DESIRE_CHECK_ARG(d->B >= 0 && d->N > 0 && d->K > 0 && d->H > 0 && d->Tf > 0 && d->iters >= 0 && d->Fv % 4 == 0 &&
                       d->Cs % 4 == 0 && (2 * d->C) % 4 == 0,
                   "desire_ioc_fwd: bad dimensions");
  const IocLayout L = ioc_layout(d);
  if (!ws || ws_bytes < L.total) {
    set_error("desire_ioc_fwd: workspace too small (%zu < %zu)", ws_bytes, L.total);
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long R = (long)d->B * d->N * d->K;
  if (R == 0) return DESIRE_OK;
  const int T = d->Tf, H = d->H, K = d->K, Fv = d->Fv, Cs = d->Cs, C2 = 2 * d->C;
  const int Dst = Fv + Cs + C2, G = d->n_rad * d->n_ang;
  char* base = (char*)ws;
  float* Xs = (float*)(base + L.Xs);
  float* XP = (float*)(base + L.XP);
  float* pooled = (float*)(base + L.pooled);
  float* fsp = (float*)(base + L.fsp);
  float* h2 = (float*)(base + L.h2);
  float* wsp3 = (float*)(base + L.wsp3);
  PackWs pw{base + L.pack, PACK_WS_BYTES};
  const desire_gru_t& g = w->dec2;
  // [H,3H] = dec2 rows [Dst,Dst+H) of (wg | wc): projection of the social feature fsp, applied per step on
  // top of the hoisted projection of the static features
  DESIRE_CUDA(cudaMemcpy2DAsync(wsp3, 3 * H * sizeof(float), g.wg + (size_t)Dst * 2 * H, 2 * H * sizeof(float),
                                2 * H * sizeof(float), H, cudaMemcpyDeviceToDevice, st));
  DESIRE_CUDA(cudaMemcpy2DAsync(wsp3 + 2 * H, 3 * H * sizeof(float), g.wc + (size_t)Dst * H, H * sizeof(float),
                                H * sizeof(float), H, cudaMemcpyDeviceToDevice, st));

  // feature_pooling columns of the static input are iteration-invariant
  copy_cols_kernel<<<blocks(R * T * C2, 256), 256, 0, st>>>(fpool, C2, R * T, Xs + Fv + Cs, Dst);
  DESIRE_LAUNCH_CHECK();

  for (int it = 0; it < d->iters; ++it) {
    float* score = scores + (size_t)it * R;
    vel_fc_kernel<<<blocks(R * T * Fv, 256), 256, 0, st>>>(Y, obs, Tp, R, K, T, Fv, w->vel_w, w->vel_b, Xs, Dst);
    DESIRE_LAUNCH_CHECK();
    {
      ProfScope ps_(DESIRE_PROF_GATHER, st);
      scene_gather_kernel<<<blocks(R * T * 32, 256), 256, 0, st>>>(fmap, d->Hm, d->Wm, Cs, Y, 2, R * T, d->N * K * T,
                                                                    Xs + Fv, Dst);
    }
    DESIRE_LAUNCH_CHECK();
    // hoisted input projection of the static features for all T steps: XP[(r,t), r|u|c]
    {
      ProfScope ps_(DESIRE_PROF_DEC2_XPROJ, st);
      DESIRE_TRY(sgemm(Xs, Dst, g.wg, 2 * H, false, g.bg, XP, 3 * H, (int)(R * T), 2 * H, Dst, DESIRE_ACT_NONE, false, st, pw));
    }
    {
      ProfScope ps_(DESIRE_PROF_DEC2_XPROJ, st);
      DESIRE_TRY(sgemm(Xs, Dst, g.wc, H, false, g.bc, XP + 2 * H, 3 * H, (int)(R * T), H, Dst, DESIRE_ACT_NONE, false, st, pw));
    }
    expand_rows_kernel<<<blocks(R * H, 256), 256, 0, st>>>(Hx, ld_hx, K, H, R, h2);
    DESIRE_LAUNCH_CHECK();
    for (int t = 0; t < T; ++t) {
      {
        ProfScope ps_(DESIRE_PROF_SOCIAL_POOL, st);
        DESIRE_TRY(social_pool_launch(Y + 2 * t, 2L * T, h2, H, obs, Tp, d->B, d->N, K, H, d->n_rad, d->n_ang,
                                      w->r2_edges, w->dirs, pooled, st));
      }
      {
        ProfScope ps_(DESIRE_PROF_SOCIAL_FC, st);
        DESIRE_TRY(sgemm(pooled, G * H, w->sp_w, H, false, w->sp_b, fsp, H, (int)R, H, G * H, DESIRE_ACT_RELU, false, st, pw));
      }
      {
        // XP[:, t, :] += fsp @ wsp3  (completes the step's input projection)
        ProfScope ps_(DESIRE_PROF_DEC2_XPROJ, st);
        DESIRE_TRY(sgemm(fsp, H, wsp3, 3 * H, false, nullptr, XP + (size_t)t * 3 * H, T * 3 * H, (int)R, 3 * H, H,
                         DESIRE_ACT_NONE, true, st, pw));
      }
      GruSeqArgs a{};
      a.R = (int)R; a.H = H; a.T = 1;
      a.xp = XP + (size_t)t * 3 * H; a.xp_row_stride = (long)T * 3 * H; a.xp_step_stride = 0;
      a.Ka = 0;
      a.w_g = g.wg + (size_t)(Dst + H) * 2 * H;   // state rows
      a.w_c = g.wc + (size_t)(Dst + H) * H;
      a.h0 = h2; a.h0_div = 1; a.ld_h0 = H;
      a.h_final = h2; a.ld_hf = H;
      {
        ProfScope ps_(DESIRE_PROF_GRU_DEC2, st);
        DESIRE_TRY(gru_seq(a, st, pw));
      }
      score_kernel<<<blocks(R * 32, 256), 256, 0, st>>>(h2, R, H, w->score_w, w->score_b, score, t == 0 ? 1 : 0);
      DESIRE_LAUNCH_CHECK();
    }
    // regression refinement: Y[R, 2T] += h2 @ reg_w + reg_b
    DESIRE_TRY(sgemm(h2, H, w->reg_w, 2 * T, false, w->reg_b, Y, 2 * T, (int)R, 2 * T, H, DESIRE_ACT_NONE, true, st, pw));
  }
  return DESIRE_OK;
}
