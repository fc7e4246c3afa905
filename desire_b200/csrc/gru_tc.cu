// Persistent tcgen05 GRU recurrence (TF-1.x GRUCell semantics) — the dense contraction of the path.
//
//   gates = sigmoid(xp_ru + h @ Wg_h)      [128 x 2H]  accumulated in TMEM columns [0,2H)
//   cand  = tanh   (xp_c  + (r*h) @ Wc_h)  [128 x  H]  accumulated OVER the consumed r columns [0,H)
//   h'    = u*h + (1-u)*cand                           u is read from TMEM columns [H,2H) only now
//
// One CTA owns a 128-row tile (rows are independent) for all T steps.  Warp roles:
//   warps 0-7  epilogue: warp w serves TMEM lanes 32*(w%4).. and column half w/4.  After the gates MMA they
//              turn r into the next A operand r*h (BF16 hi/lo, UMMA K-major layout, in place of h), after
//              the candidate MMA they finish the state update, store h_t (FP32) to HBM and write h_t back
//              as the A operand of the next step.  FP32 state never lives in BF16: h_{t-1} is re-read
//              from the FP32 output of the previous step (L2-resident, written by the same thread).
//   warp 8     MMA issuer: tcgen05.mma M=128, N<=256, K=16, 3 MMAs per K step in 3xBF16 mode.
//   warp 9     weight streamer: the recurrent weights are pre-packed into shared-memory images and
//              streamed every step through a 3-stage ring with 1-D bulk TMA copies (they do not fit
//              next to the state tile in 3xBF16 form: 12*H^2 bytes).
// mbarriers: ring full/empty, gates-done, cand-done (tcgen05.commit), rh-ready, h-ready (epilogue).
#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int NS = 3;                 // weight ring stages
constexpr int SLOT_BYTES = 32 * 1024; // 2 (hi,lo) x 4 chunks x 256 rows x 16 B
constexpr int NTHR = 320;

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
  const uint8_t* wg;   // packed [ntg][H/32] blocks of 128*BNg bytes
  const uint8_t* wc;   // packed [H/32] blocks of 128*H bytes
  int BNg, ntg;
  int passes;
  uint32_t tmem_cols;
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2/(1+e^{2x}); saturates correctly for |x| large (e^{2x} -> inf or 0)
  return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
}

__global__ void __launch_bounds__(NTHR, 1) gru_tc_kernel(GruTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H;
  const int a_half = (H / 8) * 2048;                 // [H/8 chunks][128 rows][16 B]
  uint8_t* A_hi = smem;
  uint8_t* A_lo = smem + a_half;
  uint8_t* ring = smem + 2 * a_half;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NS * SLOT_BYTES);
  uint64_t* empty = full + NS;
  uint64_t* g_done = empty + NS;
  uint64_t* c_done = g_done + 1;
  uint64_t* rh_ready = c_done + 1;
  uint64_t* h_ready = rh_ready + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(h_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long row_blk = (long)blockIdx.x * 128;
  const int nks = H / 32;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(g_done, 1);
    mbar_init(c_done, 1);
    mbar_init(rh_ready, 8);
    mbar_init(h_ready, 8);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc_dyn(tslot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < 8) {
    // ======================================================================== epilogue warps
    const int q = warp & 3, hf = warp >> 2;
    const int rloc = q * 32 + lane;                 // TMEM lane == row in tile
    const long row = row_blk + rloc;
    const bool ok = row < a.R;
    const int HC = H / 2, cbeg = hf * HC;           // this thread's columns [cbeg, cbeg+HC)
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* my_hi = A_hi + rloc * 16;
    uint8_t* my_lo = A_lo + rloc * 16;

    // ---- initial state -> A operand
    {
      const float* h0r = (a.h0 && ok) ? a.h0 + (row / a.h0_div) * (long)a.ld_h0 : nullptr;
      for (int c = cbeg; c < cbeg + HC; c += 8) {
        float v[8];
        if (h0r) {
          float4 x = *reinterpret_cast<const float4*>(h0r + c);   // plain loads: h0 may alias h_final
          float4 y = *(reinterpret_cast<const float4*>(h0r + c) + 1);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        Split8 s = split8(v);
        *reinterpret_cast<uint4*>(my_hi + (c / 8) * 2048) = s.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8) * 2048) = s.lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
    }

    for (int t = 0; t < a.T; ++t) {
      const uint32_t par = t & 1;
      const float* xpr = ok ? a.xp + row * a.xp_row_stride + (long)t * a.xp_step_stride : nullptr;
      const float* hprev = nullptr;
      if (ok) {
        if (t == 0) hprev = a.h0 ? a.h0 + (row / a.h0_div) * (long)a.ld_h0 : nullptr;
        else hprev = a.hs + row * a.hs_row_stride + (long)(t - 1) * a.hs_step_stride;
      }
      // ---------------- E1: r = sigmoid(.), stage r*h as the candidate's A operand
      mbar_wait(g_done, par);
      tc_fence_after();
      for (int c = cbeg; c < cbeg + HC; c += 16) {
        float acc[16];
        tmem_ld16(trow + c, acc);
        tmem_ld_wait();
        float xv[16], hv[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 x = xpr ? __ldg(reinterpret_cast<const float4*>(xpr + c + i)) : make_float4(0, 0, 0, 0);
          float4 h = hprev ? *reinterpret_cast<const float4*>(hprev + c + i) : make_float4(0, 0, 0, 0);
          xv[i] = x.x; xv[i + 1] = x.y; xv[i + 2] = x.z; xv[i + 3] = x.w;
          hv[i] = h.x; hv[i + 1] = h.y; hv[i + 2] = h.z; hv[i + 3] = h.w;
        }
        float rh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) rh[i] = fast_sigmoid(acc[i] + xv[i]) * hv[i];
        Split8 s0 = split8(rh), s1 = split8(rh + 8);
        *reinterpret_cast<uint4*>(my_hi + (c / 8) * 2048) = s0.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8) * 2048) = s0.lo;
        *reinterpret_cast<uint4*>(my_hi + (c / 8 + 1) * 2048) = s1.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8 + 1) * 2048) = s1.lo;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(rh_ready);

      // ---------------- E2: u, candidate, state update
      mbar_wait(c_done, par);
      tc_fence_after();
      float* hout = ok ? a.hs ? a.hs + row * a.hs_row_stride + (long)t * a.hs_step_stride : nullptr : nullptr;
      float* hfin = (ok && a.h_final && t == a.T - 1) ? a.h_final + row * (long)a.ld_hf : nullptr;
      for (int c = cbeg; c < cbeg + HC; c += 16) {
        float accc[16], accu[16];
        tmem_ld16(trow + c, accc);
        tmem_ld16(trow + H + c, accu);
        tmem_ld_wait();
        float hn[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 xu = xpr ? __ldg(reinterpret_cast<const float4*>(xpr + H + c + i)) : make_float4(0, 0, 0, 0);
          float4 xc = xpr ? __ldg(reinterpret_cast<const float4*>(xpr + 2 * H + c + i)) : make_float4(0, 0, 0, 0);
          float4 h = hprev ? *reinterpret_cast<const float4*>(hprev + c + i) : make_float4(0, 0, 0, 0);
          const float xu_[4] = {xu.x, xu.y, xu.z, xu.w}, xc_[4] = {xc.x, xc.y, xc.z, xc.w}, h_[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float u = fast_sigmoid(accu[i + e] + xu_[e]);
            const float cd = fast_tanh(accc[i + e] + xc_[e]);
            hn[i + e] = u * h_[e] + (1.f - u) * cd;
          }
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 o = make_float4(hn[i], hn[i + 1], hn[i + 2], hn[i + 3]);
          if (hout) *reinterpret_cast<float4*>(hout + c + i) = o;
          if (hfin) *reinterpret_cast<float4*>(hfin + c + i) = o;
        }
        Split8 s0 = split8(hn), s1 = split8(hn + 8);
        *reinterpret_cast<uint4*>(my_hi + (c / 8) * 2048) = s0.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8) * 2048) = s0.lo;
        *reinterpret_cast<uint4*>(my_hi + (c / 8 + 1) * 2048) = s1.hi;
        *reinterpret_cast<uint4*>(my_lo + (c / 8 + 1) * 2048) = s1.lo;
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
    }
  } else if (warp == 8) {
    // ======================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_g = idesc_bf16(128, a.BNg), idesc_c = idesc_bf16(128, H);
      const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo);
      const uint32_t lbo_g = a.BNg * 16, lbo_c = H * 16;
      const uint32_t g_half = 4 * a.BNg * 16, c_half = 4 * H * 16;
      uint32_t it = 0;
      for (int t = 0; t < a.T; ++t) {
        const uint32_t par = t & 1;
        mbar_wait(h_ready, par);
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
              const uint32_t d = tmem + jn * a.BNg;
              const uint32_t accf = (ks > 0 || j > 0) ? 1u : 0u;
              mma_bf16(d, ahi, bhi, idesc_g, accf);
              if (a.passes == 3) {
                mma_bf16(d, alo, bhi, idesc_g, 1);
                mma_bf16(d, ahi, blo, idesc_g, 1);
              }
            }
            mma_commit(&empty[slot]);
          }
        }
        mma_commit(g_done);
        mbar_wait(rh_ready, par);
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
            const uint32_t accf = (ks > 0 || j > 0) ? 1u : 0u;
            mma_bf16(tmem, ahi, bhi, idesc_c, accf);
            if (a.passes == 3) {
              mma_bf16(tmem, alo, bhi, idesc_c, 1);
              mma_bf16(tmem, ahi, blo, idesc_c, 1);
            }
          }
          mma_commit(&empty[slot]);
        }
        mma_commit(c_done);
      }
    }
  } else {
    // ======================================================================== weight streamer
    if (lane == 0) {
      const uint32_t g_bytes = 2 * 4 * a.BNg * 16, c_bytes = 2 * 4 * H * 16;
      uint32_t it = 0;
      for (int t = 0; t < a.T; ++t) {
        for (int ks = 0; ks < nks; ++ks) {
          for (int jn = 0; jn < a.ntg; ++jn, ++it) {
            const int slot = it % NS;
            mbar_wait(&empty[slot], ((it / NS) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[slot], g_bytes);
            bulk_g2s(ring + slot * SLOT_BYTES, a.wg + ((size_t)jn * nks + ks) * g_bytes, g_bytes, &full[slot]);
          }
        }
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
  if (warp == 8) tmem_dealloc(tmem, a.tmem_cols);
}

}  // namespace

size_t gru_tc_pack_bytes(int H) {
  const int BNg = 2 * H <= 256 ? 2 * H : 256;
  return align_up(tc_pack_bytes(H, 2 * H, BNg)) + align_up(tc_pack_bytes(H, H, H));
}

bool gru_tc_eligible(const GruSeqArgs& a, const void* pack_ws, size_t pack_bytes) {
  if (gemm_mode() == 0 || !a.xp || a.traj || a.ex || a.Ka != 0) return false;
  if (a.H % 32 != 0 || a.H < 32 || a.H > 256) return false;
  if (a.T > 1 && !a.hs) return false;
  if (!a.packed && (!pack_ws || pack_bytes < gru_tc_pack_bytes(a.H))) return false;
  const size_t smem = 2 * (size_t)(a.H / 8) * 2048 + NS * SLOT_BYTES + 128;
  return smem <= 227 * 1024 && a.R >= 64;
}

int gru_seq_tc(const GruSeqArgs& s, void* pack_ws, cudaStream_t st) {
  const int H = s.H;
  GruTcArgs a{};
  a.R = s.R; a.H = H; a.T = s.T;
  a.xp = s.xp; a.xp_row_stride = s.xp_row_stride; a.xp_step_stride = s.xp_step_stride;
  a.h0 = s.h0; a.h0_div = s.h0_div > 0 ? s.h0_div : 1; a.ld_h0 = s.ld_h0;
  a.hs = s.hs; a.hs_row_stride = s.hs_row_stride; a.hs_step_stride = s.hs_step_stride;
  a.h_final = s.h_final; a.ld_hf = s.ld_hf;
  a.BNg = 2 * H <= 256 ? 2 * H : 256;
  a.ntg = (2 * H) / a.BNg;
  a.passes = gemm_mode() == 1 ? 1 : 3;
  uint32_t cols = 32;
  while ((int)cols < 2 * H) cols <<= 1;
  a.tmem_cols = cols;
  const uint8_t* pg = (const uint8_t*)(s.packed ? s.packed : pack_ws);
  if (!s.packed) DESIRE_TRY(gru_tc_pack(s.w_g, s.w_c, H, pack_ws, gru_tc_pack_bytes(H), st));
  a.wg = pg;
  a.wc = pg + align_up(tc_pack_bytes(H, 2 * H, a.BNg));
  const size_t smem = 2 * (size_t)(H / 8) * 2048 + NS * SLOT_BYTES + 128;
  DESIRE_ENSURE_SMEM(gru_tc_kernel, smem);
  const unsigned grid = (unsigned)(((long)s.R + 127) / 128);
  DESIRE_LAUNCH(st, (gru_tc_kernel<<<grid, NTHR, smem, st>>>(a)));
  return DESIRE_OK;
}

int gru_tc_pack(const float* w_g, const float* w_c, int H, void* ws, size_t ws_bytes, cudaStream_t st) {
  DESIRE_CHECK_ARG(ws && ws_bytes >= gru_tc_pack_bytes(H), "gru_tc_pack: workspace too small");
  const int BNg = 2 * H <= 256 ? 2 * H : 256;
  uint8_t* pg = (uint8_t*)ws;
  uint8_t* pc = pg + align_up(tc_pack_bytes(H, 2 * H, BNg));
  DESIRE_TRY(tc_pack_b(w_g, 2 * H, false, H, 2 * H, BNg, pg, st));
  DESIRE_TRY(tc_pack_b(w_c, H, false, H, H, H, pc, st));
  return DESIRE_OK;
}

}  // namespace desire
