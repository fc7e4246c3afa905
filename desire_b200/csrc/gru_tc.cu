// Persistent tcgen05 GRU recurrence (TF-1.x GRUCell semantics) — the dense contraction of the path.
//
//   gates = sigmoid(xp_ru + h @ Wg_h)      [128 x 2H]  TMEM columns [0,2H) of the tile
//   cand  = tanh   (xp_c  + (r*h) @ Wc_h)  [128 x  H]  accumulated OVER the consumed r columns [0,H)
//   h'    = u*h + (1-u)*cand                           u is read from columns [H,2H) only now
//
// Design (v2, after the ncu capture of v1 showed 12 % tensor-pipe: serial phases, epilogues stalled on
// dependent global loads):
//   * one CTA owns NT independent 128-row tiles for all T steps and ONE MMA warp serves them:
//     gates(0) gates(1) cand(0) cand(1) ...  While a tile's epilogue warps work, the tensor core runs the
//     other tile's MMAs (ping-pong inside the CTA; TMEM = NT * 2H columns).  The kernel supports NT = 2 for H <= 128,
//     but shape_of() ships NT = 1: at the bench workload 150 two-tile CTAs still need two rounds of the 148 SMs and
//     measured no faster than 300 one-tile CTAs ("Measured and not kept").
//   * the hoisted input projection xp is never added in an epilogue: it is PRE-LOADED into the TMEM accumulator
//     with tcgen05.st (xp_r|xp_u by the previous step's epilogue right after it consumed those columns, xp_c by
//     the gate epilogue right after it consumed r) and every MMA accumulates on top of it.  The epilogues'
//     global loads are therefore independent prefetches issued at the top of each 16-column chunk.
//   * epilogue warps: 4 lane quadrants x (H/HC) column splits per tile, HC = 64 columns per thread.  E1 turns r
//     into the next A operand r*h (BF16 hi/lo, UMMA K-major layout, in place of h); E2 finishes the state
//     update, stores h_t (FP32) and writes h_t back as the A operand.  FP32 state never lives in BF16: h_{t-1}
//     is re-read from the FP32 output of the previous step (written by the same thread, L1/L2 resident).
//   * recurrent weights: pre-packed shared-memory images streamed through a 3-stage ring with 1-D bulk TMA
//     copies (12*H^2 bytes per tile-step in 3xBF16 form — they do not fit next to the state tiles).
#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int NS = 3;                 // weight ring stages
constexpr int SLOT_BYTES = 32 * 1024; // 2 (hi,lo) x 4 chunks x 256 rows x 16 B
constexpr int MAXNT = 2;

struct GruTcArgs {
  int R, H, T;
  const float* xp;
  long xp_row_stride, xp_step_stride;
  const float* h0;
  int h0_div, ld_h0;
  float* hs;
  long hs_row_stride, hs_step_stride;
  float* h_final;
  int ld_hf;
  const float* ex;     // optional extra per-step operand [R,Ka] (Decoder-2: the social feature), Ka in {0, H}
  int Ka, ld_ex;
  const uint8_t* wg;   // packed [ntg][(Ka+H)/32] blocks of 128*BNg bytes (rows 0..Ka-1 multiply ex, then h)
  const uint8_t* wc;   // packed [(Ka+H)/32] blocks of 128*H bytes
  int BNg, ntg;
  int passes;
  int NT, HC, EW;      // tiles per CTA, columns per epilogue thread, epilogue warps per tile
  uint32_t tmem_cols;
};

__device__ __forceinline__ void ld16(const float* p, float* v) {   // plain loads (state written by this kernel)
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(p + i);
    v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
  }
}
__device__ __forceinline__ void ld16g(const float* p, float* v) {  // read-only data (xp)
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p + i));
    v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
  }
}
__device__ __forceinline__ void zero16(float* v) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) gru_tc_kernel(GruTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, NT = a.NT;
  const int KA = a.Ka + H;                           // A operand width: [ex | h]
  const int a_half = (KA / 8) * 2048;                // [KA/8 chunks][128 rows][16 B] per tile
  uint8_t* ring = smem + (size_t)NT * 2 * a_half;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NS * SLOT_BYTES);
  uint64_t* empty = full + NS;
  uint64_t* g_done = empty + NS;        // [MAXNT]
  uint64_t* c_done = g_done + MAXNT;
  uint64_t* rh_ready = c_done + MAXNT;
  uint64_t* h_ready = rh_ready + MAXNT;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(h_ready + MAXNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_epi = NT * a.EW;                       // epilogue warps; then MMA warp, loader warp
  const int nks = KA / 32;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int ti = 0; ti < MAXNT; ++ti) {
      mbar_init(&g_done[ti], 1);
      mbar_init(&c_done[ti], 1);
      mbar_init(&rh_ready[ti], a.EW);
      mbar_init(&h_ready[ti], a.EW);
    }
    fence_barrier_init();
  }
  if (warp == n_epi) tmem_alloc_dyn(tslot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < n_epi) {
    // ======================================================================== epilogue warps
    const int ti = warp / a.EW, w8 = warp % a.EW;
    const int q = w8 & 3, cs = w8 >> 2;             // TMEM lane quadrant (== warp % 4), column split
    const int rloc = q * 32 + lane;                 // TMEM lane == row in tile
    const long row_true = ((long)blockIdx.x * NT + ti) * 128 + rloc;
    const bool ok = row_true < a.R;
    const long row = ok ? row_true : (long)a.R - 1;   // rows past the end recompute the last row (loads stay
                                                      // branch-free and in bounds); they never store
    const int HC = a.HC, cbeg = cs * HC;            // this thread's columns [cbeg, cbeg+HC)
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + ti * 2 * H;
    uint8_t* my_hi = smem + (size_t)ti * 2 * a_half + (a.Ka / 8) * 2048 + rloc * 16;   // h / r*h chunks (after ex)
    uint8_t* my_lo = my_hi + a_half;
    const float* xp0 = a.xp + row * a.xp_row_stride;
    const float* h0r = a.h0 ? a.h0 + (row / a.h0_div) * (long)a.ld_h0 : nullptr;

    // ---- extra operand (constant over the launch) -> first Ka/8 chunks of the A operand
    if (a.ex) {
      const float* exr = a.ex + row * (long)a.ld_ex;
      uint8_t* e_hi = smem + (size_t)ti * 2 * a_half + rloc * 16;
      for (int c = cbeg; c < cbeg + HC; c += 16) {
        float ev[16];
        ld16g(exr + c, ev);
        const Split8 s0 = split8(ev), s1 = split8(ev + 8);
        *reinterpret_cast<uint4*>(e_hi + (c / 8) * 2048) = s0.hi;
        *reinterpret_cast<uint4*>(e_hi + a_half + (c / 8) * 2048) = s0.lo;
        *reinterpret_cast<uint4*>(e_hi + (c / 8 + 1) * 2048) = s1.hi;
        *reinterpret_cast<uint4*>(e_hi + a_half + (c / 8 + 1) * 2048) = s1.lo;
      }
    }
    // ---- initial state -> A operand, xp_r|xp_u of step 0 -> TMEM accumulator
    for (int c = cbeg; c < cbeg + HC; c += 16) {
      float hv[16], xr[16], xu[16];
      if (h0r) ld16(h0r + c, hv);
      else zero16(hv);
      ld16g(xp0 + c, xr);
      ld16g(xp0 + H + c, xu);
      const Split8 s0 = split8(hv), s1 = split8(hv + 8);
      *reinterpret_cast<uint4*>(my_hi + (c / 8) * 2048) = s0.hi;
      *reinterpret_cast<uint4*>(my_lo + (c / 8) * 2048) = s0.lo;
      *reinterpret_cast<uint4*>(my_hi + (c / 8 + 1) * 2048) = s1.hi;
      *reinterpret_cast<uint4*>(my_lo + (c / 8 + 1) * 2048) = s1.lo;
      tmem_st16(trow + c, xr);
      tmem_st16(trow + H + c, xu);
    }
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&h_ready[ti]);

    for (int t = 0; t < a.T; ++t) {
      const uint32_t par = t & 1;
      const float* xpt = xp0 + (long)t * a.xp_step_stride;
      const float* xpn = xp0 + (long)(t + 1 < a.T ? t + 1 : t) * a.xp_step_stride;
      const float* hprev = (t == 0) ? h0r : a.hs + row * a.hs_row_stride + (long)(t - 1) * a.hs_step_stride;
      // ---------------- E1: r = sigmoid(.), stage r*h as the candidate's A operand, pre-load xp_c
      mbar_wait(&g_done[ti], par);
      tc_fence_after();
      for (int c = cbeg; c < cbeg + HC; c += 16) {
        float hv[16], xc[16], acc[16];
        if (hprev) ld16(hprev + c, hv);
        else zero16(hv);
        ld16g(xpt + 2 * H + c, xc);
        tmem_ld16(trow + c, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = sigmoid_a(acc[i]) * hv[i];
        const Split8 s0 = split8(acc), s1 = split8(acc + 8);
        *reinterpret_cast<uint4*>(my_hi + (c / 8) * 2048) = s0.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8) * 2048) = s0.lo;
        *reinterpret_cast<uint4*>(my_hi + (c / 8 + 1) * 2048) = s1.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8 + 1) * 2048) = s1.lo;
        tmem_st16(trow + c, xc);                     // the candidate MMA accumulates on top of xp_c
      }
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&rh_ready[ti]);

      // ---------------- E2: u, candidate, state update; pre-load xp_r|xp_u of the next step
      mbar_wait(&c_done[ti], par);
      tc_fence_after();
      float* hout = (ok && a.hs) ? a.hs + row * a.hs_row_stride + (long)t * a.hs_step_stride : nullptr;
      float* hfin = (ok && a.h_final && t == a.T - 1) ? a.h_final + row * (long)a.ld_hf : nullptr;
      for (int c = cbeg; c < cbeg + HC; c += 16) {
        float hv[16], xr[16], xu[16], accc[16], accu[16];
        if (hprev) ld16(hprev + c, hv);
        else zero16(hv);
        ld16g(xpn + c, xr);
        ld16g(xpn + H + c, xu);
        tmem_ld16(trow + c, accc);
        tmem_ld16(trow + H + c, accu);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float u = sigmoid_a(accu[i]);
          const float cd = tanh_a(accc[i]);
          accc[i] = fmaf(u, hv[i] - cd, cd);          // u*h + (1-u)*cd
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 o = make_float4(accc[i], accc[i + 1], accc[i + 2], accc[i + 3]);
          if (hout) *reinterpret_cast<float4*>(hout + c + i) = o;
          if (hfin) *reinterpret_cast<float4*>(hfin + c + i) = o;
        }
        const Split8 s0 = split8(accc), s1 = split8(accc + 8);
        *reinterpret_cast<uint4*>(my_hi + (c / 8) * 2048) = s0.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8) * 2048) = s0.lo;
        *reinterpret_cast<uint4*>(my_hi + (c / 8 + 1) * 2048) = s1.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8 + 1) * 2048) = s1.lo;
        if (t + 1 < a.T) {
          tmem_st16(trow + c, xr);
          tmem_st16(trow + H + c, xu);
        }
      }
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[ti]);
    }
  } else if (warp == n_epi) {
    // ======================================================================== MMA issuer (serves all tiles)
    if (lane == 0) {
      const uint32_t idesc_g = idesc_bf16(128, a.BNg), idesc_c = idesc_bf16(128, H);
      const uint32_t lbo_g = a.BNg * 16, lbo_c = H * 16;
      const uint32_t g_half = 4 * a.BNg * 16, c_half = 4 * H * 16;
      uint32_t it = 0;
      for (int t = 0; t < a.T; ++t) {
        const uint32_t par = t & 1;
        for (int ti = 0; ti < NT; ++ti) {           // gates of every tile
          const uint32_t a_hi = smem_u32(smem + (size_t)ti * 2 * a_half), a_lo = a_hi + a_half;
          mbar_wait(&h_ready[ti], par);
          tc_fence_after();
          for (int ks = 0; ks < nks; ++ks) {
            for (int jn = 0; jn < a.ntg; ++jn, ++it) {
              const int slot = it % NS;
              mbar_wait(&full[slot], (it / NS) & 1);
              tc_fence_after();
              const uint32_t sb = smem_u32(ring + slot * SLOT_BYTES);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t ao = (ks * 4 + j * 2) * 2048;
                const uint64_t ahi = smem_desc(a_hi + ao, 2048, 128), alo = smem_desc(a_lo + ao, 2048, 128);
                const uint64_t bhi = smem_desc(sb + j * 2 * lbo_g, lbo_g, 128);
                const uint64_t blo = smem_desc(sb + g_half + j * 2 * lbo_g, lbo_g, 128);
                const uint32_t d = tmem + ti * 2 * H + jn * a.BNg;
                mma_bf16(d, ahi, bhi, idesc_g, 1);   // accumulates on the pre-loaded xp_r|xp_u
                if (a.passes == 3) {
                  mma_bf16(d, alo, bhi, idesc_g, 1);
                  mma_bf16(d, ahi, blo, idesc_g, 1);
                }
              }
              mma_commit(&empty[slot]);
            }
          }
          mma_commit(&g_done[ti]);
        }
        for (int ti = 0; ti < NT; ++ti) {           // candidates of every tile
          const uint32_t a_hi = smem_u32(smem + (size_t)ti * 2 * a_half), a_lo = a_hi + a_half;
          mbar_wait(&rh_ready[ti], par);
          tc_fence_after();
          for (int ks = 0; ks < nks; ++ks, ++it) {
            const int slot = it % NS;
            mbar_wait(&full[slot], (it / NS) & 1);
            tc_fence_after();
            const uint32_t sb = smem_u32(ring + slot * SLOT_BYTES);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint32_t ao = (ks * 4 + j * 2) * 2048;
              const uint64_t ahi = smem_desc(a_hi + ao, 2048, 128), alo = smem_desc(a_lo + ao, 2048, 128);
              const uint64_t bhi = smem_desc(sb + j * 2 * lbo_c, lbo_c, 128);
              const uint64_t blo = smem_desc(sb + c_half + j * 2 * lbo_c, lbo_c, 128);
              const uint32_t d = tmem + ti * 2 * H;
              mma_bf16(d, ahi, bhi, idesc_c, 1);     // accumulates on the pre-loaded xp_c
              if (a.passes == 3) {
                mma_bf16(d, alo, bhi, idesc_c, 1);
                mma_bf16(d, ahi, blo, idesc_c, 1);
              }
            }
            mma_commit(&empty[slot]);
          }
          mma_commit(&c_done[ti]);
        }
      }
    }
  } else if (warp == n_epi + 1) {
    // ======================================================================== weight streamer
    if (lane == 0) {
      const uint32_t g_bytes = 2 * 4 * a.BNg * 16, c_bytes = 2 * 4 * H * 16;
      uint32_t it = 0;
      for (int t = 0; t < a.T; ++t) {
        for (int ti = 0; ti < NT; ++ti)
          for (int ks = 0; ks < nks; ++ks)
            for (int jn = 0; jn < a.ntg; ++jn, ++it) {
              const int slot = it % NS;
              mbar_wait(&empty[slot], ((it / NS) & 1) ^ 1);
              mbar_arrive_expect_tx(&full[slot], g_bytes);
              bulk_g2s(ring + slot * SLOT_BYTES, a.wg + ((size_t)jn * nks + ks) * g_bytes, g_bytes, &full[slot]);
            }
        for (int ti = 0; ti < NT; ++ti)
          for (int ks = 0; ks < nks; ++ks, ++it) {
            const int slot = it % NS;
            mbar_wait(&empty[slot], ((it / NS) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[slot], c_bytes);
            bulk_g2s(ring + slot * SLOT_BYTES, a.wc + (size_t)ks * c_bytes, c_bytes, &full[slot]);
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == n_epi) tmem_dealloc(tmem, a.tmem_cols);
}

struct Shape {
  int NT, HC, EW, nthr;
  uint32_t cols;
  size_t smem;
};
Shape shape_of(int H, int Ka) {
  Shape s;
  s.HC = (H % 64 == 0) ? 64 : 32;
  s.EW = 4 * (H / s.HC);
  s.NT = 1;   // two tiles per CTA (ping-pong) measured no faster: the epilogue, not the MMA, is the critical path
  auto bytes = [&](int nt) { return (size_t)nt * 2 * ((Ka + H) / 8) * 2048 + NS * SLOT_BYTES + 256; };
  if (s.NT == 2 && bytes(2) > 227 * 1024) s.NT = 1;
  s.smem = bytes(s.NT);
  s.nthr = (s.NT * s.EW + 2) * 32;
  uint32_t c = 32;
  while ((int)c < s.NT * 2 * H) c <<= 1;
  s.cols = c;
  return s;
}

}  // namespace

size_t gru_tc_pack_bytes(int H, int Ka) {
  const int BNg = 2 * H <= 256 ? 2 * H : 256;
  return align_up(tc_pack_bytes(Ka + H, 2 * H, BNg)) + align_up(tc_pack_bytes(Ka + H, H, H));
}

bool gru_tc_eligible(const GruSeqArgs& a, const void* pack_ws, size_t pack_bytes) {
  if (gemm_mode() == 0 || !a.xp || a.traj) return false;
  if (a.H % 32 != 0 || a.H < 32 || a.H > 256) return false;
  if (a.ex ? (a.Ka != a.H || a.T != 1 || a.ld_ex % 4 != 0) : a.Ka != 0) return false;
  if (a.T > 1 && !a.hs) return false;
  if (a.packed ? a.packed_fmt != 0 : (!pack_ws || pack_bytes < gru_tc_pack_bytes(a.H, a.Ka))) return false;
  const Shape s = shape_of(a.H, a.Ka);
  return s.smem <= 227 * 1024 && s.nthr <= 1024 && a.R >= 64;
}

int gru_tc_pack(const float* w_g, const float* w_c, int H, int Ka, void* ws, size_t ws_bytes, cudaStream_t st) {
  DESIRE_CHECK_ARG(ws && ws_bytes >= gru_tc_pack_bytes(H, Ka), "gru_tc_pack: workspace too small");
  const int BNg = 2 * H <= 256 ? 2 * H : 256;
  uint8_t* pg = (uint8_t*)ws;
  uint8_t* pc = pg + align_up(tc_pack_bytes(Ka + H, 2 * H, BNg));
  DESIRE_TRY(tc_pack_b(w_g, 2 * H, false, Ka + H, 2 * H, BNg, pg, st));
  DESIRE_TRY(tc_pack_b(w_c, H, false, Ka + H, H, H, pc, st));
  return DESIRE_OK;
}

int gru_seq_tc(const GruSeqArgs& s, void* pack_ws, cudaStream_t st) {
  const int H = s.H;
  const Shape sh = shape_of(H, s.Ka);
  GruTcArgs a{};
  a.R = s.R; a.H = H; a.T = s.T;
  a.xp = s.xp; a.xp_row_stride = s.xp_row_stride; a.xp_step_stride = s.xp_step_stride;
  a.h0 = s.h0; a.h0_div = s.h0_div > 0 ? s.h0_div : 1; a.ld_h0 = s.ld_h0;
  a.hs = s.hs; a.hs_row_stride = s.hs_row_stride; a.hs_step_stride = s.hs_step_stride;
  a.h_final = s.h_final; a.ld_hf = s.ld_hf;
  a.ex = s.ex; a.Ka = s.Ka; a.ld_ex = s.ld_ex;
  a.BNg = 2 * H <= 256 ? 2 * H : 256;
  a.ntg = (2 * H) / a.BNg;
  a.passes = gemm_mode() == 1 ? 1 : 3;
  a.NT = sh.NT; a.HC = sh.HC; a.EW = sh.EW; a.tmem_cols = sh.cols;
  const uint8_t* pg = (const uint8_t*)(s.packed ? s.packed : pack_ws);
  if (!s.packed) DESIRE_TRY(gru_tc_pack(s.w_g, s.w_c, H, s.Ka, pack_ws, gru_tc_pack_bytes(H, s.Ka), st));
  a.wg = pg;
  a.wc = pg + align_up(tc_pack_bytes(s.Ka + H, 2 * H, a.BNg));
  const long tiles = ((long)s.R + 127) / 128;
  const unsigned grid = (unsigned)((tiles + sh.NT - 1) / sh.NT);
  if (sh.nthr <= 576) {            // H in {32, 64, 128, 192, 256}: up to 113 registers per thread
    DESIRE_ENSURE_SMEM(gru_tc_kernel<576>, sh.smem);
    DESIRE_LAUNCH(st, (gru_tc_kernel<576><<<grid, sh.nthr, sh.smem, st>>>(a)));
  } else {
    DESIRE_ENSURE_SMEM(gru_tc_kernel<1024>, sh.smem);
    DESIRE_LAUNCH(st, (gru_tc_kernel<1024><<<grid, sh.nthr, sh.smem, st>>>(a)));
  }
  return DESIRE_OK;
}

}  // namespace desire
