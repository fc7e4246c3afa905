// TF-1.x GRUCell recurrence, FP32 CUDA-core version (exact-arithmetic path).
//
//   gates = sigmoid(xp_ru + [ex,h] @ w_g)        r | u
//   cand  = tanh   (xp_c  + [ex, r*h] @ w_c)     reset applied BEFORE the matmul (TF semantics)
//   h'    = u*h + (1-u)*cand
//
// Rows are independent, so one CTA owns a tile of rows for ALL T steps: the state tile lives in
// shared memory (transposed, k-major, so a thread reads its 4 rows with one LDS.128), the
// recurrent weights are streamed through shared memory in 16-row chunks, and only h_t leaves the
// SM.  A thread owns 4 rows x the SAME 4 columns of r, u and cand, so u and h_old stay in
// registers between the two dependent matmuls.  Used by the encoders (input width 2 projected
// inline), Decoder-1 (hoisted constant input projection) and, with T==1 plus the extra operand
// `ex`, one Decoder-2 step.
#include "common.cuh"

namespace desire {
namespace {

constexpr int RM = 4;    // rows per thread (the vector width of the k-major state tile)
constexpr int BKW = 16;  // weight rows per smem chunk

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__global__ void __launch_bounds__(256) gru_seq_kernel(GruSeqArgs a, int cgn, int rgn) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, Ka = a.Ka, KA = a.Ka + a.H;
  const int BMt = rgn * RM, LD = BMt + 4;
  float* At = smem;              // [KA][LD]  rows 0..Ka-1: ex^T, rows Ka..: h^T
  float* Rt = At + KA * LD;      // [H][LD]   (r*h)^T
  float* Ws = Rt + H * LD;       // [BKW][2H]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool active = tid < cgn * rgn;
  const int cg = tid % cgn, rg = tid / cgn;
  const int c0 = cg * 4;
  const long row_blk = (long)blockIdx.x * BMt;
  const long row0 = row_blk + rg * RM;

  // ---- initial state (and the extra operand) into shared memory, transposed
  for (int idx = tid; idx < BMt * H; idx += nthr) {
    int r = idx / H, c = idx % H;
    long row = row_blk + r;
    float v = 0.f;
    if (a.h0 && row < a.R) v = __ldg(a.h0 + (row / a.h0_div) * (long)a.ld_h0 + c);
    At[(Ka + c) * LD + r] = v;
  }
  if (a.ex) {
    for (int idx = tid; idx < BMt * Ka; idx += nthr) {
      int r = idx / Ka, c = idx % Ka;
      long row = row_blk + r;
      At[c * LD + r] = row < a.R ? __ldg(a.ex + row * (long)a.ld_ex + c) : 0.f;
    }
  }
  __syncthreads();

  for (int t = 0; t < a.T; ++t) {
    // ---- hoisted / inline input projection for this thread's 4 rows x 4 cols x (r,u,c)
    float xr[RM][4], xu[RM][4], xc[RM][4];
    if (active) {
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        long row = row0 + i;
        bool ok = row < a.R;
        if (a.traj) {
          float x0 = 0.f, x1 = 0.f;
          if (ok) {
            const float* p = a.traj + (row * a.T + t) * 3;
            x0 = __ldg(p + 1);
            x1 = __ldg(p + 2);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int c = c0 + j;
            xr[i][j] = fmaf(x1, __ldg(a.wx_g + 2 * H + c), fmaf(x0, __ldg(a.wx_g + c), __ldg(a.bg + c)));
            xu[i][j] = fmaf(x1, __ldg(a.wx_g + 3 * H + c), fmaf(x0, __ldg(a.wx_g + H + c), __ldg(a.bg + H + c)));
            xc[i][j] = fmaf(x1, __ldg(a.wx_c + H + c), fmaf(x0, __ldg(a.wx_c + c), __ldg(a.bc + c)));
          }
        } else {
          float4 vr = make_float4(0, 0, 0, 0), vu = vr, vc = vr;
          if (ok) {
            const float* p = a.xp + row * a.xp_row_stride + t * a.xp_step_stride + c0;
            vr = ld4(p);
            vu = ld4(p + H);
            vc = ld4(p + 2 * H);
          }
          xr[i][0] = vr.x; xr[i][1] = vr.y; xr[i][2] = vr.z; xr[i][3] = vr.w;
          xu[i][0] = vu.x; xu[i][1] = vu.y; xu[i][2] = vu.z; xu[i][3] = vu.w;
          xc[i][0] = vc.x; xc[i][1] = vc.y; xc[i][2] = vc.z; xc[i][3] = vc.w;
        }
      }
    }

    // ---- gates = [ex,h] @ w_g
    float ar[RM][4], au[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) ar[i][j] = au[i][j] = 0.f;
    for (int k0 = 0; k0 < KA; k0 += BKW) {
      const int kn = min(BKW, KA - k0);
      for (int idx = tid; idx < kn * (2 * H / 4); idx += nthr) {
        int k = idx / (2 * H / 4), c4 = idx % (2 * H / 4);
        *reinterpret_cast<float4*>(&Ws[k * 2 * H + c4 * 4]) = ld4(a.w_g + (size_t)(k0 + k) * 2 * H + c4 * 4);
      }
      __syncthreads();
      if (active) {
#pragma unroll 4
        for (int k = 0; k < kn; ++k) {
          float4 av = ld4(&At[(k0 + k) * LD + rg * RM]);
          float4 wr = ld4(&Ws[k * 2 * H + c0]);
          float4 wu = ld4(&Ws[k * 2 * H + H + c0]);
          float a_[RM] = {av.x, av.y, av.z, av.w};
          float wr_[4] = {wr.x, wr.y, wr.z, wr.w};
          float wu_[4] = {wu.x, wu.y, wu.z, wu.w};
#pragma unroll
          for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ar[i][j] = fmaf(a_[i], wr_[j], ar[i][j]);
              au[i][j] = fmaf(a_[i], wu_[j], au[i][j]);
            }
        }
      }
      __syncthreads();
    }

    // ---- r, u; stage r*h for the candidate matmul
    float hold[RM][4], u[RM][4];
    if (active) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 hv = ld4(&At[(Ka + c0 + j) * LD + rg * RM]);
        hold[0][j] = hv.x; hold[1][j] = hv.y; hold[2][j] = hv.z; hold[3][j] = hv.w;
        float rh[RM];
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          float r = sigmoidf_(ar[i][j] + xr[i][j]);
          u[i][j] = sigmoidf_(au[i][j] + xu[i][j]);
          rh[i] = r * hold[i][j];
        }
        *reinterpret_cast<float4*>(&Rt[(c0 + j) * LD + rg * RM]) = make_float4(rh[0], rh[1], rh[2], rh[3]);
      }
    }
    __syncthreads();

    // ---- cand = [ex, r*h] @ w_c
    float ac[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) ac[i][j] = 0.f;
    for (int k0 = 0; k0 < KA; k0 += BKW) {
      const int kn = min(BKW, KA - k0);
      for (int idx = tid; idx < kn * (H / 4); idx += nthr) {
        int k = idx / (H / 4), c4 = idx % (H / 4);
        *reinterpret_cast<float4*>(&Ws[k * H + c4 * 4]) = ld4(a.w_c + (size_t)(k0 + k) * H + c4 * 4);
      }
      __syncthreads();
      if (active) {
#pragma unroll 4
        for (int k = 0; k < kn; ++k) {
          int kk = k0 + k;
          const float* src = kk < Ka ? &At[kk * LD] : &Rt[(kk - Ka) * LD];
          float4 av = ld4(src + rg * RM);
          float4 wc = ld4(&Ws[k * H + c0]);
          float a_[RM] = {av.x, av.y, av.z, av.w};
          float wc_[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
          for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ac[i][j] = fmaf(a_[i], wc_[j], ac[i][j]);
        }
      }
      __syncthreads();
    }

    // ---- state update; h_t back into shared memory and out to HBM
    if (active) {
      float hn[RM][4];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float c = tanhf(ac[i][j] + xc[i][j]);
          hn[i][j] = u[i][j] * hold[i][j] + (1.f - u[i][j]) * c;
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(&At[(Ka + c0 + j) * LD + rg * RM]) =
            make_float4(hn[0][j], hn[1][j], hn[2][j], hn[3][j]);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        long row = row0 + i;
        if (row >= a.R) continue;
        float4 v = make_float4(hn[i][0], hn[i][1], hn[i][2], hn[i][3]);
        if (a.hs) *reinterpret_cast<float4*>(a.hs + row * a.hs_row_stride + t * a.hs_step_stride + c0) = v;
        if (a.h_final && t == a.T - 1) *reinterpret_cast<float4*>(a.h_final + row * (long)a.ld_hf + c0) = v;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int gru_seq(const GruSeqArgs& a, cudaStream_t st, PackWs pw) {
  if (gru_tc3_eligible(a, pw.p, pw.bytes)) return gru_seq_tc3(a, pw.p, st);
  if (gru_tc_eligible(a, pw.p, pw.bytes)) {
    note_fallback(DESIRE_FALLBACK_GRU_V2, "GRU recurrence on the second tcgen05 design (H, R, T)", a.H, a.R, a.T);
    return gru_seq_tc(a, pw.p, st);
  }
  note_fallback(DESIRE_FALLBACK_GRU_FP32, "GRU recurrence on FP32 CUDA cores (H, R, T)", a.H, a.R, a.T);
  DESIRE_CHECK_ARG(a.H % 4 == 0 && a.H >= 4 && a.H <= 1024, "gru: H=%d must be a multiple of 4 in [4,1024]", a.H);
  DESIRE_CHECK_ARG(a.Ka % 4 == 0, "gru: extra operand width %d must be a multiple of 4", a.Ka);
  DESIRE_CHECK_ARG(!a.ex || a.T == 1, "gru: the extra operand is per-step (T must be 1)");
  DESIRE_CHECK_ARG((a.xp != nullptr) != (a.traj != nullptr), "gru: exactly one of xp / traj");
  if (a.R == 0 || a.T == 0) return DESIRE_OK;
  const int cgn = a.H / 4;
  const int rgn = cgn >= 256 ? 1 : 256 / cgn;   // (smaller row tiles for the few-row encoders measured 3x slower:
                                               //  fewer threads per CTA stream the same weights)
  const int BMt = rgn * RM, LD = BMt + 4;
  const int nthr = ((cgn * rgn + 31) / 32) * 32;
  size_t smem = ((size_t)(a.Ka + 2 * a.H) * LD + (size_t)BKW * 2 * a.H) * sizeof(float);
  DESIRE_CHECK_ARG(smem <= 227 * 1024, "gru: tile does not fit shared memory (H=%d Ka=%d)", a.H, a.Ka);
  DESIRE_ENSURE_SMEM(gru_seq_kernel, smem);
  long grid = ((long)a.R + BMt - 1) / BMt;
  DESIRE_LAUNCH(st, (gru_seq_kernel<<<(unsigned)grid, nthr, smem, st>>>(a, cgn, rgn)));
  return DESIRE_OK;
}

}  // namespace desire

using namespace desire;

extern "C" int desire_gru_encode_fwd(const float* traj, int M, int T, int H, const desire_gru_t* w, float* h_out,
                                     int ld_out, desire_stream_t stream) {
  DESIRE_CHECK_ARG(traj && w && h_out && M >= 0 && T > 0, "desire_gru_encode_fwd: bad arguments");
  DESIRE_CHECK_ARG(ld_out >= H && ld_out % 4 == 0, "desire_gru_encode_fwd: ld_out must be >= H and a multiple of 4");
  GruSeqArgs a{};
  a.R = M; a.H = H; a.T = T;
  a.traj = traj;
  a.wx_g = w->wg; a.wx_c = w->wc; a.bg = w->bg; a.bc = w->bc;   // rows 0,1 = input rows
  a.w_g = w->wg + 2 * 2 * H;                                    // rows 2.. = state rows
  a.w_c = w->wc + 2 * H;
  a.Ka = 0;
  a.h0 = nullptr; a.h0_div = 1;
  a.h_final = h_out; a.ld_hf = ld_out;
  ProfScope ps_(DESIRE_PROF_GRU_ENC, (cudaStream_t)stream);
  return gru_seq(a, (cudaStream_t)stream);
}

// Encoder on the tensor-core recurrence: the width-2 input projection is hoisted for all T steps (one tiny kernel),
// then the same persistent tcgen05 GRU as Decoder-1 runs the T steps (FP32 CUDA-core recurrence when not eligible).
extern "C" size_t desire_gru_encode_workspace_bytes(int M, int T, int H) {
  const size_t m = (size_t)M;
  return align_up(m * T * 3 * H * sizeof(float)) + align_up(m * T * H * sizeof(float)) + PACK_WS_BYTES;
}

extern "C" int desire_gru_encode_ws_fwd(const float* traj, int M, int T, int H, const desire_gru_t* w, float* h_out,
                                        int ld_out, void* ws, size_t ws_bytes, desire_stream_t stream) {
  DESIRE_CHECK_ARG(traj && w && h_out && M >= 0 && T > 0, "desire_gru_encode_ws_fwd: bad arguments");
  DESIRE_CHECK_ARG(ld_out >= H && ld_out % 4 == 0, "desire_gru_encode_ws_fwd: ld_out must be >= H and a multiple of 4");
  if (!ws || ws_bytes < desire_gru_encode_workspace_bytes(M, T, H)) {
    set_error("desire_gru_encode_ws_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  if (M == 0) return DESIRE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)M;
  Workspace W(ws, ws_bytes);
  float* xp = W.take<float>(m * T * 3 * H);
  float* hs = W.take<float>(m * T * H);
  PackWs pw{W.take<char>(PACK_WS_BYTES), PACK_WS_BYTES};
  ProfScope ps_(DESIRE_PROF_GRU_ENC, st);
  DESIRE_TRY(xproj_traj(traj, m * T, H, w, xp, st));
  GruSeqArgs a{};
  a.R = M; a.H = H; a.T = T;
  a.xp = xp; a.xp_row_stride = (long)T * 3 * H; a.xp_step_stride = 3 * H;
  a.w_g = w->wg + 2 * 2 * H;                                    // rows 2.. = state rows
  a.w_c = w->wc + 2 * H;
  a.Ka = 0;
  a.h0 = nullptr; a.h0_div = 1;
  a.hs = hs; a.hs_row_stride = (long)T * H; a.hs_step_stride = H;
  a.h_final = h_out; a.ld_hf = ld_out;
  return gru_seq(a, st, pw);
}

extern "C" size_t desire_gru_decode_workspace_bytes(int R, int H) {
  return align_up((size_t)R * 3 * H * sizeof(float)) + PACK_WS_BYTES;
}

extern "C" int desire_gru_decode_fwd(const float* x_z, const float* Hx, int ld_hx, int R, int K, int H, int T,
                                     const desire_gru_t* w, float* hs, void* ws, size_t ws_bytes,
                                     desire_stream_t stream) {
  DESIRE_CHECK_ARG(x_z && Hx && w && hs && R >= 0 && K > 0 && T > 0, "desire_gru_decode_fwd: bad arguments");
  if (ws_bytes < desire_gru_decode_workspace_bytes(R, H) || !ws) {
    set_error("desire_gru_decode_fwd: workspace too small");
    return DESIRE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* xp = (float*)ws;  // [R,3H] hoisted input projection (the input is the same every step)
  PackWs pw{(char*)ws + align_up((size_t)R * 3 * H * sizeof(float)), PACK_WS_BYTES};
  DESIRE_TRY(sgemm(x_z, H, w->wg, 2 * H, false, w->bg, xp, 3 * H, R, 2 * H, H, DESIRE_ACT_NONE, false, st, pw));
  DESIRE_TRY(sgemm(x_z, H, w->wc, H, false, w->bc, xp + 2 * H, 3 * H, R, H, H, DESIRE_ACT_NONE, false, st, pw));
  GruSeqArgs a{};
  a.R = R; a.H = H; a.T = T;
  a.xp = xp; a.xp_row_stride = 3 * H; a.xp_step_stride = 0;
  a.w_g = w->wg + (size_t)H * 2 * H;
  a.w_c = w->wc + (size_t)H * H;
  a.Ka = 0;
  a.h0 = Hx; a.h0_div = K; a.ld_h0 = ld_hx;
  a.hs = hs; a.hs_row_stride = (long)T * H; a.hs_step_stride = H;
  ProfScope ps_(DESIRE_PROF_GRU_DEC1, st);
  return gru_seq(a, st, pw);
}
