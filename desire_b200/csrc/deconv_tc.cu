// Fused transposed convolution for the CVAE decoder (model/model.py:465-468, deconv2d of
// utils/convolutional_vae_util.py:31-135):   Y = act( BN_row( conv2d_transpose(X, W) + b ) )
// in ONE kernel on tcgen05 — the scatter-form GEMM col = X[R*Pin, Cin] @ W^T[Cin, 25*Cout], the col2im
// accumulation, the bias, the per-row batch-norm (batch-of-one statistics) and the activation.
// The col matrix (4*25*Cout bytes per input position; 12 GB written + read per step at the bench workload)
// never leaves the SM.
//
// One CTA = 128 GEMM rows = SPT = 128/Pin whole samples, so every output position of those samples is produced
// inside the CTA.  Channels are processed in groups of 32 (out tile [SPT][Pout][32] FP32 = 64 KB of smem):
//   warps 0-7  (a) convert the X tile once to BF16 hi/lo in the UMMA K-major layout (A stays resident);
//              (b) epilogue: accumulators arrive per n-tile of 8 taps x 32 channels; tap by tap (named barrier
//                  between taps => deterministic, race-free) each thread adds its 16 channels of its input
//                  position into the output position (iy*s+ky-pad, ix*s+kx-pad) of the smem tile (XOR-rotated
//                  channel quads => conflict-free);
//              (c) two-pass BN statistics per (sample, channel) over the tile, normalise, activate, store.
//   warp 8     tcgen05.mma issuer, M=128, N=256 (last n-tile: one tap, N=32), two TMEM accumulators so the MMAs
//              of n-tile t+1 overlap the scatter of n-tile t.
//   warp 9     streams the packed weights (one 32-wide K stage of one n-tile per 1-D bulk TMA copy).
#include <algorithm>
#include <utility>

#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int TM = 128, CG = 32, TPT = 8, NTHR = 320;
constexpr int SLOT = 32 * 1024;   // 2 (hi,lo) x 4 chunks x 256 rows x 16 B

struct DcArgs {
  const float* X;
  int R, Hin, Hout, Cin, Cout, ks, stride, pad;
  const uint8_t* wpack;
  const float *bias, *gamma, *beta;
  int act;
  float* Y;
  int passes, nstg;
  // Fused last decoder layer (model/model.py:468, 16x16x32 -> 32x32x1 k5 s2 SAME + BN + sigmoid) on the normalised tile of
  // THIS layer while it is still in shared memory (only for the <8, 16, 2, 5> geometry with Cout == 32): Y4 != nullptr
  // switches it on; then Y is not written at all.
  const uint8_t* w4pack;   // { hi [4][32][8 bf16], lo } : [N = 25 taps (32), K = 32]
  const float *b4, *g4, *be4;
  int act4;
  float* Y4;               // [R, 1024]
};

struct DcLayout {
  size_t a_hi, a_lo, ring, out, red, stat, bars, w4, total;
};
__host__ __device__ inline DcLayout dc_layout(int Cin, int nstg, int spt) {
  DcLayout L;
  size_t off = 0;
  const size_t a_half = (size_t)(Cin / 8) * TM * 16;
  L.a_hi = off; off += a_half;
  L.a_lo = off; off += a_half;
  L.ring = off; off += (size_t)nstg * SLOT;
  L.out = off; off += 64 * 1024;
  L.red = off; off += 256 * 16;     // float4 partial sums of the vectorised BN phase
  L.stat = off; off += (size_t)2 * spt * CG * 4;
  L.bars = off; off += (2 * 4 + 4 + 1 + 3) * 8 + 16;
  off = (off + 127) / 128 * 128;
  L.w4 = off; off += 4096;                                       // fused last layer: its packed weights
  L.total = off;
  return L;
}

// Order in which the taps are laid out along N and scattered.  Two taps can write the same output cell only if they
// agree in (ky mod stride, kx mod stride), so the taps are taken ROUND by round — round r holds the r-th tap of every
// such class — and a round's taps (up to stride^2 of them) are scattered back to back; the lockstep barrier is only
// needed between rounds: 9 barriers instead of 25 for the 5x5 stride-2 layer (stride 1: every tap is its own round).
template <int KS, int STRIDE>
struct TapOrder {
  int tap[KS * KS];        // slot -> tap (ky * KS + kx)
  bool last[KS * KS];      // slot is the last one of its round
  constexpr TapOrder() : tap{}, last{} {
    int cls_n[STRIDE * STRIDE] = {};
    int cls_tap[STRIDE * STRIDE][KS * KS] = {};
    for (int t = 0; t < KS * KS; ++t) {
      const int c = ((t / KS) % STRIDE) * STRIDE + (t % KS) % STRIDE;
      cls_tap[c][cls_n[c]++] = t;
    }
    int s = 0;
    for (int r = 0; s < KS * KS; ++r) {
      for (int c = 0; c < STRIDE * STRIDE; ++c)
        if (cls_n[c] > r) tap[s++] = cls_tap[c][r];
      last[s - 1] = true;
    }
  }
};
struct TapPerm {
  int tap[32];
};
// everything the scatter loop needs to know about a slot, as constant expressions
template <int KS, int STRIDE, int TPT_>
struct TapPlan {
  static constexpr TapOrder<KS, STRIDE> ORD{};
  static constexpr int NT = KS * KS;
  // the taps loaded so far are scattered after this slot: its round ends, or its n-tile (= the accumulator) ends
  static constexpr bool flush(int s) { return ORD.last[s] || s % TPT_ == TPT_ - 1 || s == NT - 1; }
  // taps of the same n-tile loaded before slot s and not scattered yet
  static constexpr int pending(int s) {
    int n = 0;
    for (int q = s - 1; q >= (s / TPT_) * TPT_ && !flush(q); --q) ++n;
    return n;
  }
};
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// packed weights: for cg, nt, ks: { hi [4][BNt][8 bf16], lo [4][BNt][8 bf16] } with n = slot_local*32 + c (slot -> tap: TapOrder)
__global__ void pack_deconv_kernel(const float* __restrict__ W, int Cin, int Cout, int ntaps, TapPerm perm,
                                   uint4* __restrict__ out) {
  const int nks = Cin / 32, ncg = Cout / CG, nnt = (ntaps + TPT - 1) / TPT;
  // chunk-granular work items: (cg, nt, ks, c, n)
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  // enumerate by walking the tiles (few of them)
  long base_u4 = 0;
  for (int cg = 0; cg < ncg; ++cg)
    for (int nt = 0; nt < nnt; ++nt) {
      const int taps = min(TPT, ntaps - nt * TPT), BNt = taps * CG;
      const long items = (long)nks * 4 * BNt;
      if (idx < items) {
        const int n = (int)(idx % BNt);
        long t = idx / BNt;
        const int c = (int)(t % 4);
        const int ks = (int)(t / 4);
        const int tap = perm.tap[nt * TPT + n / CG], o = cg * CG + n % CG;
        const float* src = W + ((size_t)tap * Cout + o) * Cin + ks * 32 + c * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i);
        const Split8 s = split8(v);
        uint4* blk = out + base_u4 + (size_t)ks * (2 * 4 * BNt);
        blk[c * BNt + n] = s.hi;
        blk[4 * BNt + c * BNt + n] = s.lo;
        return;
      }
      idx -= items;
      base_u4 += (long)nks * 2 * 4 * BNt;
    }
}

__device__ __forceinline__ int swz(int oy, int ox, int sh) { return ((ox >> sh) + 4 * ((oy >> sh) & 1)) & 7; }

template <int HIN, int HOUT, int STRIDE, int KS>
__global__ void __launch_bounds__(NTHR, 1) deconv_tc_kernel(DcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int Pin = HIN * HIN, Pout = HOUT * HOUT, spt = TM / Pin;
  const DcLayout L = dc_layout(a.Cin, a.nstg, spt);
  uint8_t* A_hi = smem + L.a_hi;
  uint8_t* A_lo = smem + L.a_lo;
  uint8_t* ring = smem + L.ring;
  float* outt = reinterpret_cast<float*>(smem + L.out);
  float* red = reinterpret_cast<float*>(smem + L.red);
  float* stat = reinterpret_cast<float*>(smem + L.stat);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + 4;
  uint64_t* acc_full = empty + 4;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* a_ready = acc_empty + 2;
  uint64_t* w4_full = a_ready + 1;      // fused last layer: weights landed / A operand written / accumulators complete
  uint64_t* a4_ready = w4_full + 1;
  uint64_t* acc4_full = a4_ready + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(acc4_full + 1);
  const bool fuse4 = a.Y4 != nullptr;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int ntaps = KS * KS, nnt = (ntaps + TPT - 1) / TPT;
  const int ncg = a.Cout / CG, nks = a.Cin / 32;
  const long samp0 = (long)blockIdx.x * spt;
  constexpr int sh = STRIDE - 1;

  if (tid == 0) {
    for (int s = 0; s < a.nstg; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    mbar_init(a_ready, 8);
    mbar_init(w4_full, 1);
    mbar_init(a4_ready, 8);
    mbar_init(acc4_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc_dyn(tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < 8) {
    const int rl = tid & (TM - 1), half = tid >> 7;           // GEMM row (TMEM lane), column half
    const int s_loc = rl / Pin, p = rl % Pin;
    const long samp = samp0 + s_loc;
    const bool ok = samp < a.R;
    // ---- (a) X tile -> BF16 hi/lo A operand; this thread converts half of the K chunks of its row
    {
      const float* xr = a.X + ((size_t)samp * Pin + p) * a.Cin;
      const int nch = a.Cin / 8;
      for (int c = half * (nch / 2); c < (half + 1) * (nch / 2); ++c) {
        float v[8];
        if (ok) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(xr + c * 8));
          const float4 y = __ldg(reinterpret_cast<const float4*>(xr + c * 8) + 1);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        const Split8 sp = split8(v);
        *reinterpret_cast<uint4*>(A_hi + (size_t)c * TM * 16 + rl * 16) = sp.hi;
        *reinterpret_cast<uint4*>(A_lo + (size_t)c * TM * 16 + rl * 16) = sp.lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
    const int iy = p / HIN, ix = p % HIN;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* osamp = outt + (size_t)s_loc * Pout * CG;
    constexpr int ntile_el = spt * Pout * CG;                  // 16384 floats = 64 KB
    uint32_t use = 0;
    for (int cg = 0; cg < ncg; ++cg) {
      // zero the output tile of this channel group
      for (int e = tid; e < ntile_el / 4; e += 256) reinterpret_cast<float4*>(outt)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- (b) scatter the accumulators, n-tile by n-tile, round by round (TapOrder): the taps of a round cannot meet in
      // an output cell, so their TMEM loads are issued together and their read-modify-writes run back to back; rounds
      // stay in lockstep (named barrier) => deterministic, race-free
      using Plan = TapPlan<KS, STRIDE, TPT>;
      static_for<nnt>([&](auto ntc) {
        constexpr int nt = decltype(ntc)::value;
        const int ab = use & 1;
        mbar_wait(&acc_full[ab], (use >> 1) & 1);
        tc_fence_after();
        float v[STRIDE * STRIDE][16];
        static_for<TPT>([&](auto tlc) {
          constexpr int tl = decltype(tlc)::value, slot = nt * TPT + tl;
          if constexpr (slot < ntaps) {
            constexpr int nvb = Plan::pending(slot);
            tmem_ld16(trow + ab * 256 + tl * CG + half * 16, v[nvb]);
            if constexpr (Plan::flush(slot)) {
              tmem_ld_wait();
              static_for<nvb + 1>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                constexpr int t = Plan::ORD.tap[slot - nvb + k];
                constexpr int ky = t / KS, kx = t - ky * KS;
                const int oy = iy * STRIDE + ky - a.pad, ox = ix * STRIDE + kx - a.pad;
                if (oy >= 0 && oy < HOUT && ox >= 0 && ox < HOUT) {
                  float* o = osamp + (size_t)(oy * HOUT + ox) * CG;
                  const int rot = swz(oy, ox, sh);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    float4* dst = reinterpret_cast<float4*>(o) + ((half * 4 + q + rot) & 7);
                    float4 cur = *dst;
                    cur.x += v[k][4 * q]; cur.y += v[k][4 * q + 1]; cur.z += v[k][4 * q + 2]; cur.w += v[k][4 * q + 3];
                    *dst = cur;
                  }
                }
              });
              if constexpr (Plan::ORD.last[slot]) asm volatile("bar.sync 1, 256;" ::: "memory");
            }
          }
        });
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
        ++use;
      });
      // ---- (c) bias + per-(sample, channel) BN over the Pout positions + activation + store, four channels per
      // thread: a thread owns one channel quad (its bias / gamma / beta live in registers), every tile access is one
      // LDS.128 of the XOR-rotated quad and every store one 16-byte STG (the scalar form of this phase was 47 % of
      // the kernel's instructions).
      constexpr int units = spt * 8, parts = 256 / units;      // (sample, quad) units: 16 or 64; position slices per unit
      float4* red4 = reinterpret_cast<float4*>(red);
      float4* stat4 = reinterpret_cast<float4*>(stat);         // [units] mean, then [units] rstd
      {
        const int unit = tid % units, part = tid / units;
        const int us = unit >> 3, uq = unit & 7;
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + cg * CG) + uq);
        const float* tsamp = outt + (size_t)us * Pout * CG;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = part; q < Pout; q += parts) {
          const int oy = q / HOUT, ox = q - oy * HOUT;
          const float4 x = *reinterpret_cast<const float4*>(tsamp + q * CG + ((uq + swz(oy, ox, sh)) & 7) * 4);
          sum.x += x.x + b4.x; sum.y += x.y + b4.y; sum.z += x.z + b4.z; sum.w += x.w + b4.w;
        }
        red4[tid] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < units) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < parts; ++q) {
            const float4 r = red4[q * units + tid];
            t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
          }
          const float inv = 1.f / (float)Pout;
          stat4[tid] = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float4 mean = stat4[unit];
        sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = part; q < Pout; q += parts) {
          const int oy = q / HOUT, ox = q - oy * HOUT;
          const float4 x = *reinterpret_cast<const float4*>(tsamp + q * CG + ((uq + swz(oy, ox, sh)) & 7) * 4);
          const float dx = x.x + b4.x - mean.x, dy = x.y + b4.y - mean.y, dz = x.z + b4.z - mean.z, dw = x.w + b4.w - mean.w;
          sum.x += dx * dx; sum.y += dy * dy; sum.z += dz * dz; sum.w += dw * dw;
        }
        red4[tid] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < units) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < parts; ++q) {
            const float4 r = red4[q * units + tid];
            t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
          }
          const float inv = 1.f / (float)Pout;
          stat4[units + tid] = make_float4(1.f / sqrtf(t.x * inv + 1e-3f), 1.f / sqrtf(t.y * inv + 1e-3f),
                                           1.f / sqrtf(t.z * inv + 1e-3f), 1.f / sqrtf(t.w * inv + 1e-3f));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      {
        const int qd = tid & 7;                                // 256 % 8 == 0: the quad of a thread never changes
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.gamma + cg * CG) + qd);
        const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + cg * CG) + qd);
        const float4 bb4 = __ldg(reinterpret_cast<const float4*>(a.bias + cg * CG) + qd);
        if (fuse4) {
          // normalised + activated quads go straight into the A operand of the last layer: M-tile = 128 positions of one
          // sample, K = the 32 channels; byte(row, k) = (k/8)*2048 + row*16 + (k%8)*2 (hi image, then lo image)
          uint8_t* a4 = ring;                                    // 4 M-tiles x 16 KB: the weight ring is idle by now
#pragma unroll 1
          for (int it0 = 0; it0 < ntile_el / 4 / 256; it0 += 4) {
            float4 x[4], m4[4], r4[4];
            int pqs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int pq = (tid + (it0 + u) * 256) >> 3;
              pqs[u] = pq;
              const int q = pq % Pout, s2 = pq / Pout;
              const int oy = q / HOUT, ox = q - oy * HOUT;
              x[u] = *reinterpret_cast<const float4*>(outt + ((size_t)s2 * Pout + q) * CG + ((qd + swz(oy, ox, sh)) & 7) * 4);
              m4[u] = stat4[s2 * 8 + qd];
              r4[u] = stat4[units + s2 * 8 + qd];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float y0 = act_fast(g4.x * ((x[u].x + bb4.x - m4[u].x) * r4[u].x) + be4.x, a.act);
              const float y1 = act_fast(g4.y * ((x[u].y + bb4.y - m4[u].y) * r4[u].y) + be4.y, a.act);
              const float y2 = act_fast(g4.z * ((x[u].z + bb4.z - m4[u].z) * r4[u].z) + be4.z, a.act);
              const float y3 = act_fast(g4.w * ((x[u].w + bb4.w - m4[u].w) * r4[u].w) + be4.w, a.act);
              uint2 hi, lo;
              split2(y0, y1, hi.x, lo.x);
              split2(y2, y3, hi.y, lo.y);
              const int mt = pqs[u] >> 7, row = pqs[u] & 127;
              uint8_t* dst = a4 + (size_t)mt * 16384 + (qd >> 1) * 2048 + row * 16 + (qd & 1) * 8;
              *reinterpret_cast<uint2*>(dst) = hi;
              *reinterpret_cast<uint2*>(dst + 8192) = lo;
            }
          }
          // the 32x32 output tiles of the two samples live where this layer's A operand was
          float* out2 = reinterpret_cast<float*>(A_hi);
          for (int e = tid; e < 2 * 1024 / 4; e += 256) reinterpret_cast<float4*>(out2)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          fence_proxy_async();
          tc_fence_before();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (lane == 0) mbar_arrive(a4_ready);
          // ---- scatter: warp w reads TMEM lanes 32*(w%4).. of the two M-tiles of sample w/4
          const int q4 = warp & 3, s2 = warp >> 2;
          mbar_wait(acc4_full, 0);
          tc_fence_after();
          float v0[32], v1[32];
          tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + (2 * s2) * 32, v0);
          tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + (2 * s2 + 1) * 32, v1);
          tmem_ld_wait();
          {
            const int p0 = q4 * 32 + lane;                       // position 0..127 of the sample; the second M-tile: +128
            const int iy0 = p0 >> 4, ix0 = p0 & 15, iy1 = iy0 + 8;
            float* o2 = out2 + s2 * 1024;
            static_for<25>([&](auto sc) {
              constexpr int slot = decltype(sc)::value;
              constexpr int t = TapPlan<5, 2, 8>::ORD.tap[slot];
              constexpr int ky = t / 5, kx = t % 5;
              const int ox = ix0 * 2 + kx - 1;
              const int oya = iy0 * 2 + ky - 1, oyb = iy1 * 2 + ky - 1;
              if (ox >= 0 && ox < 32) {
                if (oya >= 0 && oya < 32) o2[oya * 32 + ox] += v0[t];
                if (oyb >= 0 && oyb < 32) o2[oyb * 32 + ox] += v1[t];
              }
              if constexpr (TapPlan<5, 2, 8>::ORD.last[slot]) asm volatile("bar.sync 1, 256;" ::: "memory");
            });
          }
          // ---- bias + BN over the 1024 pixels of each sample + activation: 128 threads per sample, 8 pixels each
          {
            const int tl = tid & 127;
            const float* o2 = out2 + s2 * 1024;
            const float b0 = __ldg(a.b4);
            float vals[8], ssum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              vals[i] = o2[tl + 128 * i] + b0;
              ssum += vals[i];
            }
            auto half_sum = [&](float xx) {                      // sum over the four warps of this sample
              xx = warp_sum(xx);
              asm volatile("bar.sync 1, 256;" ::: "memory");
              if (lane == 0) red[warp] = xx;
              asm volatile("bar.sync 1, 256;" ::: "memory");
              return red[4 * s2] + red[4 * s2 + 1] + red[4 * s2 + 2] + red[4 * s2 + 3];
            };
            const float mean = half_sum(ssum) * (1.f / 1024.f);
            float qs = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) qs += (vals[i] - mean) * (vals[i] - mean);
            const float rstd = 1.f / sqrtf(half_sum(qs) * (1.f / 1024.f) + 1e-3f);
            const float g = __ldg(a.g4), be = __ldg(a.be4);
            const long smp = samp0 + s2;
            if (smp < a.R) {
#pragma unroll
              for (int i = 0; i < 8; ++i) a.Y4[(size_t)smp * 1024 + tl + 128 * i] = act_fast(g * ((vals[i] - mean) * rstd) + be, a.act4);
            }
          }
        } else {
        // four quads per trip: their tile / statistics reads first, then the branch-free normalise + activation (tc.cuh:
        // act_fast — with act_apply's data-dependent ELU branch and expf every element waited for the previous one: this
        // phase was a third of the kernel), then the 16-byte stores
        constexpr int NIT = ntile_el / 4 / 256;
        static_assert(NIT % 4 == 0, "store loop batches");
        const int act = a.act;
#pragma unroll 1
        for (int it0 = 0; it0 < NIT; it0 += 4) {
          float4 x[4], m4[4], r4[4];
          int qq[4], ss[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int pq = (tid + (it0 + u) * 256) >> 3;
            qq[u] = pq % Pout;
            ss[u] = pq / Pout;
            const int oy = qq[u] / HOUT, ox = qq[u] - oy * HOUT;
            x[u] = *reinterpret_cast<const float4*>(outt + ((size_t)ss[u] * Pout + qq[u]) * CG + ((qd + swz(oy, ox, sh)) & 7) * 4);
            m4[u] = stat4[ss[u] * 8 + qd];
            r4[u] = stat4[units + ss[u] * 8 + qd];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float4 y;
            y.x = act_fast(g4.x * ((x[u].x + bb4.x - m4[u].x) * r4[u].x) + be4.x, act);
            y.y = act_fast(g4.y * ((x[u].y + bb4.y - m4[u].y) * r4[u].y) + be4.y, act);
            y.z = act_fast(g4.z * ((x[u].z + bb4.z - m4[u].z) * r4[u].z) + be4.z, act);
            y.w = act_fast(g4.w * ((x[u].w + bb4.w - m4[u].w) * r4[u].w) + be4.w, act);
            if (samp0 + ss[u] < a.R)
              *reinterpret_cast<float4*>(a.Y + ((size_t)(samp0 + ss[u]) * Pout + qq[u]) * a.Cout + cg * CG + qd * 4) = y;
          }
        }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  } else if (warp == 8) {
    // ===================== MMA issuer: the whole warp runs the loops, one elected lane issues with warp-uniform operands
    // (tc.cuh: elect_one — under `if (lane == 0)` every UTCHMMA sat inside a lane-broadcast loop, ~93 cycles per MMA)
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t d_ahi = smem_desc(smem_u32(A_hi), 2048, 128), d_alo = smem_desc(smem_u32(A_lo), 2048, 128);
      const uint32_t ring_s = smem_u32(ring);
      const bool p3 = a.passes == 3;
      mbar_wait(a_ready, 0);
      tc_fence_after();
      uint32_t use = 0;
      RingPos rp;
      for (int cg = 0; cg < ncg; ++cg) {
        for (int nt = 0; nt < nnt; ++nt, ++use) {
          const int ab = use & 1;
          const int BNt = min(TPT, ntaps - nt * TPT) * CG;
          const uint32_t idesc = idesc_bf16(TM, BNt);
          const uint32_t lbo_b = BNt * 16, b_half = 4 * BNt * 16;
          const uint64_t d_b = smem_desc(ring_s, lbo_b, 128);
          const uint32_t d = tm + ab * 256;
          mbar_wait(&acc_empty[ab], ((use >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int ks = 0; ks < nks; ++ks, rp.next(a.nstg)) {
            const uint32_t slot = rp.slot;
            const uint64_t db = desc_adv(d_b, slot * (uint32_t)SLOT);
            const uint64_t dah = desc_adv(d_ahi, ks * 4 * 2048), dal = desc_adv(d_alo, ks * 4 * 2048);
            mbar_wait(&full[slot], rp.ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint64_t bhi = desc_adv(db, j * 2 * lbo_b), blo = desc_adv(db, b_half + j * 2 * lbo_b);
                const uint32_t accf = (ks > 0 || j > 0) ? 1u : 0u;
                mma_bf16(d, desc_adv(dah, j * 2 * 2048), bhi, idesc, accf);
                if (p3) {
                  mma_bf16(d, desc_adv(dal, j * 2 * 2048), bhi, idesc, 1);
                  mma_bf16(d, desc_adv(dah, j * 2 * 2048), blo, idesc, 1);
                }
              }
              mma_commit(&empty[slot]);
              if (ks == nks - 1) mma_commit(&acc_full[ab]);
            }
          }
        }
      }
      __syncwarp();
      if (fuse4) {
        // last layer: four M-tiles (two samples x two halves of the 256 positions) x [K = 32] x [N = 32 taps]
        mbar_wait(w4_full, 0);
        mbar_wait(a4_ready, 0);
        tc_fence_after();
        const uint64_t d_a4 = smem_desc(ring_s, 2048, 128), d_b4 = smem_desc(smem_u32(smem + L.w4), 512, 128);
        constexpr uint32_t idesc4 = idesc_bf16(128, 32);
        if (elect_one()) {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t ahi = desc_adv(d_a4, mt * 16384 + j * 2 * 2048), alo = desc_adv(d_a4, mt * 16384 + 8192 + j * 2 * 2048);
              const uint64_t bhi = desc_adv(d_b4, j * 2 * 512), blo = desc_adv(d_b4, 4 * 512 + j * 2 * 512);
              mma_bf16(tm + mt * 32, ahi, bhi, idesc4, j > 0 ? 1u : 0u);
              if (p3) {
                mma_bf16(tm + mt * 32, alo, bhi, idesc4, 1);
                mma_bf16(tm + mt * 32, ahi, blo, idesc4, 1);
              }
            }
          }
          mma_commit(acc4_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== weight streamer
    if (lane == 0) {
      if (fuse4) {
        mbar_arrive_expect_tx(w4_full, 4096);
        bulk_g2s(smem + L.w4, a.w4pack, 4096, w4_full);
      }
      uint32_t it = 0;
      const uint8_t* src = a.wpack;
      for (int cg = 0; cg < ncg; ++cg) {
        for (int nt = 0; nt < nnt; ++nt) {
          const int BNt = min(TPT, ntaps - nt * TPT) * CG;
          const uint32_t bytes = 2 * 4 * BNt * 16;
          for (int ks = 0; ks < nks; ++ks, ++it) {
            const int slot = it % a.nstg;
            mbar_wait(&empty[slot], ((it / a.nstg) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * SLOT, src, bytes, &full[slot]);
            src += bytes;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}


// ---------------------------------------------------------------------------------------------------------
// Single-output-channel variant (the decoder's last layer, model/model.py:468: 16x16x32 -> 32x32x1, k5 s2 SAME,
// BN + sigmoid).  One CTA = one sample = 256 input positions = two M=128 tiles against a [K=Cin, N=32] weight
// image whose column t is tap t (25 used); the accumulators (2 x 32 TMEM columns) are scattered tap by tap
// (lockstep) into a 32x32 FP32 tile in shared memory, then block-wide two-pass BN + activation.
// The CUDA-core form of this layer needs 830 M instructions per step; this one ~40 M.
struct Dc1Args {
  const float* X;      // [R, 256, 32]
  int R;
  const uint8_t* wpack;   // one block { hi [4][32][8 bf16], lo }
  const float *bias, *gamma, *beta;
  int act;
  float* Y;            // [R, 1024]
  int passes;
};

__global__ void __launch_bounds__(320) deconv1c_tc_kernel(Dc1Args a) {
  constexpr int HIN = 16, HOUT = 32, CIN = 32, KS = 5, PAD = 1;
  __shared__ __align__(1024) uint8_t A_s[2 * 2 * 4 * 128 * 16];      // [tile][hi|lo][4 chunks][128][16B] = 32 KB
  __shared__ __align__(128) uint8_t B_s[2 * 4 * 32 * 16];             // 4 KB
  __shared__ __align__(16) float out_s[HOUT * HOUT];
  __shared__ float red[8];
  __shared__ __align__(8) uint64_t bars[4];                           // b_full, a_ready, acc_full[2]
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t r = blockIdx.x;
  uint64_t* b_full = &bars[0];
  uint64_t* a_ready = &bars[1];
  uint64_t* acc_full = &bars[2];

  if (tid == 0) {
    mbar_init(b_full, 1);
    mbar_init(a_ready, 8);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc_dyn(&tslot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;

  if (warp < 8) {
    const int p = tid;                                    // input position 0..255
    const int tile = p >> 7, rl = p & 127;
    {
      const float4* xr = reinterpret_cast<const float4*>(a.X + (r * 256 + p) * CIN);
      uint8_t* ah = A_s + (size_t)tile * (2 * 4 * 2048) + rl * 16;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 x = __ldg(xr + 2 * c), y = __ldg(xr + 2 * c + 1);
        const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        const Split8 sp = split8(v);
        *reinterpret_cast<uint4*>(ah + c * 2048) = sp.hi;
        *reinterpret_cast<uint4*>(ah + 4 * 2048 + c * 2048) = sp.lo;
      }
      for (int e = tid; e < HOUT * HOUT / 4; e += 256) reinterpret_cast<float4*>(out_s)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");       // out tile zeroed before any scatter
    const int iy = p / HIN, ix = p % HIN;
    mbar_wait(&acc_full[tile], 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + tile * 32, v);
    tmem_ld_wait();
    // taps in conflict-free rounds (TapOrder: taps of different (ky, kx) parity never meet in a pixel): 9 lockstep
    // barriers instead of 25, still deterministic and race-free
    static_for<KS * KS>([&](auto sc) {
      constexpr int slot = decltype(sc)::value;
      constexpr int t = TapPlan<KS, 2, 8>::ORD.tap[slot];
      constexpr int ky = t / KS, kx = t % KS;
      const int oy = iy * 2 + ky - PAD, ox = ix * 2 + kx - PAD;
      if (oy >= 0 && oy < HOUT && ox >= 0 && ox < HOUT) out_s[oy * HOUT + ox] += v[t];
      if constexpr (TapPlan<KS, 2, 8>::ORD.last[slot]) asm volatile("bar.sync 1, 256;" ::: "memory");
    });
    // ---- bias + BN over the 1024 pixels + activation
    const float b0 = __ldg(a.bias);
    float vals[4], s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      vals[i] = out_s[tid + 256 * i] + b0;
      s += vals[i];
    }
    auto block_sum = [&](float x) {
      x = warp_sum(x);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) red[warp] = x;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float tsum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tsum += red[w];
      return tsum;
    };
    const float mean = block_sum(s) / (float)(HOUT * HOUT);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) q += (vals[i] - mean) * (vals[i] - mean);
    const float rstd = 1.f / sqrtf(block_sum(q) / (float)(HOUT * HOUT) + 1e-3f);
    const float g = __ldg(a.gamma), be = __ldg(a.beta);
#pragma unroll
    for (int i = 0; i < 4; ++i) a.Y[r * (size_t)(HOUT * HOUT) + tid + 256 * i] = act_fast(g * ((vals[i] - mean) * rstd) + be, a.act);
  } else if (warp == 8) {
    mbar_wait(b_full, 0);
    mbar_wait(a_ready, 0);
    tc_fence_after();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    if (elect_one()) {
      const uint32_t tmem = tmem_u;
      const uint32_t idesc = idesc_bf16(128, 32);
      const uint32_t sb = smem_u32(B_s);
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t sa = smem_u32(A_s + (size_t)tile * (2 * 4 * 2048));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t ahi = smem_desc(sa + j * 2 * 2048, 2048, 128), alo = smem_desc(sa + 4 * 2048 + j * 2 * 2048, 2048, 128);
          const uint64_t bhi = smem_desc(sb + j * 2 * 512, 512, 128), blo = smem_desc(sb + 4 * 512 + j * 2 * 512, 512, 128);
          const uint32_t d = tmem + tile * 32;
          mma_bf16(d, ahi, bhi, idesc, j > 0 ? 1u : 0u);
          if (a.passes == 3) {
            mma_bf16(d, alo, bhi, idesc, 1);
            mma_bf16(d, ahi, blo, idesc, 1);
          }
        }
        mma_commit(&acc_full[tile]);
      }
    }
  } else {
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, sizeof(B_s));
      bulk_g2s(B_s, a.wpack, sizeof(B_s), b_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 64);
}

}  // namespace

size_t deconv_tc_pack_bytes(int Cin, int Cout, int ks) { return align_up((size_t)ks * ks * Cout * Cin * 4); }

bool deconv_tc_eligible(int R, int Hin, int Hout, int Cin, int Cout, int ks, int stride, const void* pack_ws,
                        size_t pack_bytes) {
  const int Pin = Hin * Hin, Pout = Hout * Hout;
  if (gemm_mode() == 0 || R < 1) return false;
  if (Pin < 1 || Pin > TM || TM % Pin != 0) return false;
  if (Cin % 32 != 0 || Cin > 128 || Cout % CG != 0) return false;
  if ((size_t)(TM / Pin) * Pout * CG * 4 != 64 * 1024) return false;        // out tile exactly 64 KB
  if (256 % ((TM / Pin) * CG) != 0 && ((TM / Pin) * CG) % 256 != 0) return false;
  if ((TM / Pin) * CG > 256) return false;
  if (!((Hin == 4 && Hout == 8 && stride == 1 && ks == 5) || (Hin == 8 && Hout == 16 && stride == 2 && ks == 5)))
    return false;   // the two instantiated geometries of model/model.py:466-467
  return pack_ws && pack_bytes >= deconv_tc_pack_bytes(Cin, Cout, ks);
}

int deconv_tc(const float* X, int R, int Hin, int Hout, int Cin, int Cout, int ks, int stride, int pad, const float* W,
              const float* bias, const float* gamma, const float* beta, int act, float* Y, void* pack_ws,
              cudaStream_t st, const DeconvFuse4* fuse4) {
  const int nks = Cin / 32, ncg = Cout / CG, ntaps = ks * ks;
  long items = 0;
  for (int nt = 0; nt * TPT < ntaps; ++nt) items += (long)nks * 4 * std::min(TPT, ntaps - nt * TPT) * CG;
  items *= ncg;
  // the pack kernel walks the (cg, nt) tiles per thread; give every tile's items a thread
  TapPerm perm{};
  if (stride == 2 && ks == 5) {
    constexpr TapOrder<5, 2> o{};
    for (int i = 0; i < 25; ++i) perm.tap[i] = o.tap[i];
  } else {
    for (int i = 0; i < 32; ++i) perm.tap[i] = i;               // stride 1: natural order (TapOrder<KS, 1> is the identity)
  }
  pack_deconv_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(W, Cin, Cout, ntaps, perm, (uint4*)pack_ws);
  DESIRE_LAUNCH_CHECK();
  DcArgs a{};
  a.X = X; a.R = R; a.Hin = Hin; a.Hout = Hout; a.Cin = Cin; a.Cout = Cout; a.ks = ks; a.stride = stride; a.pad = pad;
  a.wpack = (const uint8_t*)pack_ws; a.bias = bias; a.gamma = gamma; a.beta = beta; a.act = act; a.Y = Y;
  a.passes = gemm_mode() == 1 ? 1 : 3;
  if (fuse4) {
    DESIRE_CHECK_ARG(Hin == 8 && Hout == 16 && stride == 2 && ks == 5 && Cout == 32 && fuse4->Y && fuse4->W,
                     "deconv_tc: the fused last layer needs the 8x8 -> 16x16 stride-2 geometry with 32 channels");
    // its packed weights ([N = 25 taps, K = 32], one 4 KB block) follow this layer's in the pack scratch
    uint8_t* w4 = (uint8_t*)pack_ws + align_up(deconv_tc_pack_bytes(Cin, Cout, ks));
    DESIRE_TRY(tc_pack_b(fuse4->W, 32, true, 32, 25, 32, w4, st));
    a.w4pack = w4; a.b4 = fuse4->bias; a.g4 = fuse4->gamma; a.be4 = fuse4->beta; a.act4 = fuse4->act; a.Y4 = fuse4->Y;
  }
  const int spt = TM / (Hin * Hin);
  a.nstg = Cin > 64 ? 2 : 3;
  const DcLayout L = dc_layout(Cin, a.nstg, spt);
  DESIRE_CHECK_ARG(L.total <= 227 * 1024, "deconv_tc: shared memory layout too large");
  const unsigned grid = (unsigned)((R + spt - 1) / spt);
  if (Hin == 4 && Hout == 8 && stride == 1 && ks == 5) {
    DESIRE_ENSURE_SMEM((deconv_tc_kernel<4, 8, 1, 5>), L.total);
    DESIRE_LAUNCH(st, (deconv_tc_kernel<4, 8, 1, 5><<<grid, NTHR, L.total, st>>>(a)));
  } else if (Hin == 8 && Hout == 16 && stride == 2 && ks == 5) {
    DESIRE_ENSURE_SMEM((deconv_tc_kernel<8, 16, 2, 5>), L.total);
    DESIRE_LAUNCH(st, (deconv_tc_kernel<8, 16, 2, 5><<<grid, NTHR, L.total, st>>>(a)));
  } else {
    set_error("deconv_tc: geometry %dx%d -> %dx%d k%d s%d has no instantiation", Hin, Hin, Hout, Hout, ks, stride);
    return DESIRE_ERR_INVALID;
  }
  return DESIRE_OK;
}

}  // namespace desire

namespace desire {
// decoder layer 4 on tensor cores: X [R,16,16,32] -> Y [R,32,32] (one channel); pack_ws >= 4 KB
bool deconv1c_tc_eligible() { return gemm_mode() != 0; }
int deconv1c_tc(const float* X, int R, const float* W, const float* bias, const float* gamma, const float* beta, int act,
                float* Y, void* pack_ws, cudaStream_t st) {
  if (R == 0) return DESIRE_OK;
  // W [5,5,1,32] == [N=25 taps, K=32] (K contiguous) -> one packed block with BN = 32
  DESIRE_TRY(tc_pack_b(W, 32, true, 32, 25, 32, pack_ws, st));
  Dc1Args a{X, R, (const uint8_t*)pack_ws, bias, gamma, beta, act, Y, gemm_mode() == 1 ? 1 : 3};
  DESIRE_LAUNCH(st, (deconv1c_tc_kernel<<<R, 320, 0, st>>>(a)));
  return DESIRE_OK;
}
}  // namespace desire
