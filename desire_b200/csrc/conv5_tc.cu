// 5x5, stride-1, SAME convolution (+ bias + ReLU) over NHWC FP32 maps as an IMPLICIT GEMM on tcgen05: the scene CNN's
// second and third layer (DESIGN D13: the reference leaves stage 2 unbuilt, model/model.py:312-313).
//
// Why not im2col (gemm_tc.cu, Im2colA8): every output pixel re-reads its 25 x Cin receptive field, 3.2 KB per pixel at
// Cin = 32 — 2.5 GB of L2 -> SM traffic for the two layers of a 32-image batch, and the A producers (thread = row, 32-byte
// reads) are what the kernel waits for: 0.03 of the tensor peak.
//
// Here the input tile (TH + 4 rows x TW + 4 columns, zero halo) is staged ONCE in shared memory as the K-major A operand
// itself: pixel p = row * WP + col of the padded tile is operand row p, its Cin channels are the K dimension (BF16 hi and
// lo images, 8-channel chunks LBO apart, 16 bytes per row, SBO = 128: with SWIZZLE_NONE the operand's row address is LINEAR
// in the row index).  The receptive-field shift of filter tap (ky, kx) is then nothing but a different START ADDRESS:
//     D[p, :] += A[p + ky*WP + kx, :] @ W[ky, kx]        for all 25 taps, p = 128 consecutive padded positions per MMA
// so the whole convolution is 25 x (Cin/16) x 2 MMAs per 128 positions (3xBF16 with the weights' hi and lo images as ONE
// B operand, see the issuer) with NO data movement between taps.  Positions in the 4 halo columns of each row are computed
// and dropped (128/132 useful).  The accumulators of all MT = 8 position tiles of a CTA live in tensor memory
// (MT x 2*Cout columns), so each tap's weights (a 4 KB bulk-TMA slot) are used by all eight tiles before the next tap is
// needed.  (tools/mma_rate_n.py: a start address anywhere inside a core matrix costs nothing.)
//   warps 0-15  stage the tile (FP32 -> BF16 hi/lo), later the epilogue (tcgen05.ld, bias, ReLU, 128-byte pixel rows)
//   warp 16     issues the MMAs (elected lane, warp-uniform operands), warp 17 streams the packed taps
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int C5_TH = 7, C5_TW = 128, C5_WP = C5_TW + 4, C5_MT = 8;
constexpr int C5_NPIX = (C5_TH + 4) * C5_WP;                    // 1452 staged pixels
constexpr int C5_NPIXA = (C5_NPIX + 7) / 8 * 8;                 // rows per K chunk of the operand image
constexpr int C5_LBO = C5_NPIXA * 16;
constexpr int C5_PW = 16, C5_PT = C5_PW * 32;                   // staging / epilogue warps
constexpr int C5_NTHR = (C5_PW + 2) * 32;
constexpr int C5_NSLOT = 4;
static_assert(C5_TH * C5_WP <= C5_MT * 128, "position tiles cover the tile");

struct C5Args {
  const float* X;
  const uint8_t* packed;                 // [25][Cin/8][hi rows | lo rows][8] BF16
  const float* bias;
  float* Y;
  int B, H, W, Cin, Cout, ldc, act, passes;
  long long* trace;                      // DESIRE_CONV5_TRACE=1: clock64() milestones of block (0, 0, 0)
};

__global__ void c5_pack_kernel(const float* __restrict__ w, int ldb, int Cin, int Cout, uint8_t* __restrict__ packed) {
  // w[(tap*Cin + ci)*ldb + n]  ->  packed[tap][ci/8][n' = n (hi) | Cout + n (lo)][ci%8]: one B operand of 2*Cout rows
  const int total = 25 * Cin * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % Cout, t = i / Cout, ci = t % Cin, tap = t / Cin;
    const float x = __ldg(w + (size_t)(tap * Cin + ci) * ldb + n);
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const size_t slot = (size_t)4 * Cin * Cout;
    const size_t o = (size_t)tap * slot + (size_t)(ci >> 3) * (2 * Cout * 16) + (size_t)n * 16 + (ci & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(packed + o) = hi;
    *reinterpret_cast<__nv_bfloat16*>(packed + o + (size_t)Cout * 16) = lo;
  }
}

// Space-to-depth form of the 5x5 / stride-2 / SAME convolution of a 3-channel image with even sides (TF pads 1 before):
//   out[oy, ox] = sum img[2 oy + ky - 1, 2 ox + kx - 1, c] w[ky, kx, c];   2 oy + ky - 1 = 2 (oy + a) + py,
//   (a, py) = (-1, 1), (0, 0), (0, 1), (1, 0), (1, 1) for ky = 0..4
// = a 3x3 / stride-1 convolution over S[Y, X, (py, px, c)] = img[2Y + py, 2X + px, c] (12 channels, padded to 16) with
// w3[a, b, (py, px, c)] = w[2a + py + 1, 2b + px + 1, c] (zero where that tap does not exist).
__global__ void c5_pack_s2d_kernel(const float* __restrict__ w, int ldb, int Cout, uint8_t* __restrict__ packed) {
  const int total = 9 * 16 * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % Cout, t = i / Cout, ci = t % 16, tap = t / 16;
    const int a = tap / 3 - 1, b = tap % 3 - 1;
    float x = 0.f;
    if (ci < 12) {
      const int py = ci / 6, px = (ci / 3) & 1, c = ci % 3;
      const int ky = 2 * a + py + 1, kx = 2 * b + px + 1;
      if (ky >= 0 && ky < 5 && kx >= 0 && kx < 5) x = __ldg(w + (size_t)((ky * 5 + kx) * 3 + c) * ldb + n);
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const size_t slot = (size_t)4 * 16 * Cout;
    const size_t o = (size_t)tap * slot + (size_t)(ci >> 3) * (2 * Cout * 16) + (size_t)n * 16 + (ci & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(packed + o) = hi;
    *reinterpret_cast<__nv_bfloat16*>(packed + o + (size_t)Cout * 16) = lo;
  }
}

// S2D = false: X is the NHWC map [B, H, W, CIN], 25 taps.  S2D = true (CIN = 16): X is the 3-channel image
// [B, 2H, 2W, 3], staged space-to-depth, 9 taps centred in the same (TH + 4) x (TW + 4) window.
template <int CIN, bool S2D>
__global__ void __launch_bounds__(C5_NTHR, 1) conv5_tc_kernel(C5Args a) {
  constexpr int KS = S2D ? 3 : 5, NT = KS * KS, OFF = S2D ? 1 : 0;
  constexpr int CK = CIN / 8;                                    // 8-channel K chunks
  constexpr int IMG = CK * C5_LBO;                               // bytes of one (hi or lo) operand image
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + IMG;
  const int Cout = a.Cout;
  const uint32_t slot_bytes = (uint32_t)4 * CIN * Cout;
  uint8_t* ring = smem + 2 * IMG;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(ring + C5_NSLOT * slot_bytes);
  uint64_t* wempty = wfull + C5_NSLOT;
  uint64_t* a_ready = wempty + C5_NSLOT;
  uint64_t* tfull = a_ready + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int x0 = blockIdx.x * C5_TW, y0 = blockIdx.y * C5_TH, img = blockIdx.z;
  const uint32_t tcols = (uint32_t)(C5_MT * 2 * Cout);          // 256 or 512: [hi x (hi | lo)] per position tile
  // the last row tile of a map may hold fewer rows: only the position tiles that contain one of its pixels are computed
  const int rows = min(C5_TH, a.H - y0);
  const int mt_used = (rows * C5_WP + 127) / 128;
  const bool tr = a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define TRACE(i) do { if (tr) a.trace[i] = clock64(); } while (0)
  if (tid == 0) TRACE(0);

  if (tid == 0) {
    for (int s = 0; s < C5_NSLOT; ++s) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    mbar_init(a_ready, C5_PW);
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == C5_PW) tmem_alloc_dyn(tslot, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  if (tid == 0) TRACE(1);

  if (warp == C5_PW + 1) {
    // ===================== tap loader
    if (lane == 0) {
      RingPos rp;
      for (int tap = 0; tap < NT; ++tap, rp.next(C5_NSLOT)) {
        mbar_wait(&wempty[rp.slot], rp.ph ^ 1);
        mbar_arrive_expect_tx(&wfull[rp.slot], slot_bytes);
        bulk_g2s_hint(ring + (size_t)rp.slot * slot_bytes, a.packed + (size_t)tap * slot_bytes, slot_bytes, &wfull[rp.slot],
                      L2_EVICT_LAST);
      }
    }
  } else if (warp == C5_PW) {
    // ===================== MMA issuer
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    // 3xBF16 in TWO MMAs per K step: the weights' hi and lo images are one B operand of 2*Cout rows, so
    //   D[:, 0:2C] += A_hi @ [B_hi | B_lo]   (N = 2C)      D[:, 0:C] += A_lo @ B_hi   (N = C)
    // and the epilogue adds the two column halves.  With both operands in shared memory an MMA costs 32 + N/4 cycles
    // (tools/mma_rate_n.py: the 4 KB A read dominates), so 48 + 40 cycles replace 3 x 40.
    const uint32_t idesc2 = idesc_bf16(128, 2 * Cout), idesc1 = idesc_bf16(128, Cout);
    const uint32_t lbo_b = (uint32_t)2 * Cout * 16;
    const uint64_t d_ah = smem_desc(smem_u32(a_hi), C5_LBO, 128), d_al = smem_desc(smem_u32(a_lo), C5_LBO, 128);
    const uint64_t d_ring = smem_desc(smem_u32(ring), lbo_b, 128);
    const bool p3 = a.passes == 3;
    mbar_wait(a_ready, 0);
    tc_fence_after();
    if (lane == 0) TRACE(3);
    RingPos rp;
    int ky = 0, kx = 0;
    for (int tap = 0; tap < NT; ++tap, rp.next(C5_NSLOT)) {
      const uint64_t db = desc_adv(d_ring, rp.slot * slot_bytes);
      const uint32_t shift = (uint32_t)((ky + OFF) * C5_WP + kx + OFF) * 16;   // the tap: a start address, nothing else
      const uint64_t dah = desc_adv(d_ah, shift), dal = desc_adv(d_al, shift);
      mbar_wait(&wfull[rp.slot], rp.ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int mt = 0; mt < C5_MT; ++mt) {
          if (mt >= mt_used) break;
          const uint32_t d = tm + (uint32_t)(mt * 2 * Cout);
#pragma unroll
          for (int ks = 0; ks < CK / 2; ++ks) {
            const uint64_t ah = desc_adv(dah, mt * 128 * 16 + ks * 2 * C5_LBO), al = desc_adv(dal, mt * 128 * 16 + ks * 2 * C5_LBO);
            const uint64_t b = desc_adv(db, ks * 2 * lbo_b);
            const uint32_t acc = (tap > 0 || ks > 0) ? 1u : 0u;
            if (p3) {
              mma_bf16(d, ah, b, idesc2, acc);
              mma_bf16(d, al, b, idesc1, 1);
            } else {
              mma_bf16(d, ah, b, idesc1, acc);
            }
          }
        }
        mma_commit(&wempty[rp.slot]);
        if (tap == NT - 1) mma_commit(tfull);
      }
      if (lane == 0 && tap == 0) TRACE(4);
      if (++kx == KS) {
        kx = 0;
        ++ky;
      }
    }
    if (lane == 0) TRACE(5);
    __syncwarp();
  } else {
    // ===================== stage the tile: item = (pixel, 8-channel chunk); a warp reads 8 pixels x CIN*4 contiguous bytes
    if constexpr (S2D) {
      // item = one space-to-depth pixel: two 24-byte runs of the image (rows 2Y and 2Y + 1, pixels 2X and 2X + 1)
      const int NP = (rows + 4) * C5_WP;
      const int Hi = 2 * a.H, Wi = 2 * a.W;
      const float* xin = a.X + (size_t)img * Hi * Wi * 3;
      constexpr int U = 3;
      for (int i0 = tid; i0 < NP; i0 += U * C5_PT) {
        float v[U][16];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = i0 + u * C5_PT;
          const int r = pix / C5_WP, c = pix - r * C5_WP;
          const int Y = y0 - 2 + r, X = x0 - 2 + c;
          const bool ok = pix < NP && Y >= 0 && Y < a.H && X >= 0 && X < a.W;
          float2 f[6];
#pragma unroll
          for (int e = 0; e < 6; ++e) f[e] = make_float2(0.f, 0.f);
          if (ok) {
            const float2* p0 = reinterpret_cast<const float2*>(xin + ((size_t)(2 * Y) * Wi + 2 * X) * 3);
            const float2* p1 = reinterpret_cast<const float2*>(xin + ((size_t)(2 * Y + 1) * Wi + 2 * X) * 3);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
              f[e] = __ldg(p0 + e);
              f[3 + e] = __ldg(p1 + e);
            }
          }
#pragma unroll
          for (int e = 0; e < 6; ++e) {
            v[u][2 * e] = f[e].x;
            v[u][2 * e + 1] = f[e].y;
          }
          v[u][12] = v[u][13] = v[u][14] = v[u][15] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = i0 + u * C5_PT;
          if (pix < NP) {
            const Split8 s0 = split8(v[u]), s1 = split8(v[u] + 8);
            const size_t o = (size_t)pix * 16;
            *reinterpret_cast<uint4*>(a_hi + o) = s0.hi;
            *reinterpret_cast<uint4*>(a_lo + o) = s0.lo;
            *reinterpret_cast<uint4*>(a_hi + C5_LBO + o) = s1.hi;
            *reinterpret_cast<uint4*>(a_lo + C5_LBO + o) = s1.lo;
          }
        }
      }
    } else {
      const int ITEMS = (rows + 4) * C5_WP * CK;
      const float* xin = a.X + (size_t)img * a.H * a.W * CIN;
      constexpr int U = 6;
      for (int i0 = tid; i0 < ITEMS; i0 += U * C5_PT) {
        float v[U][8];
        int pix[U], q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int item = i0 + u * C5_PT;
          pix[u] = item / CK;
          q[u] = item - pix[u] * CK;
          const int r = pix[u] / C5_WP, c = pix[u] - r * C5_WP;
          const int iy = y0 - 2 + r, ix = x0 - 2 + c;
          const bool ok = item < ITEMS && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
          float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
          if (ok) {
            const float4* p = reinterpret_cast<const float4*>(xin + ((size_t)iy * a.W + ix) * CIN + q[u] * 8);
            f0 = __ldg(p);
            f1 = __ldg(p + 1);
          }
          v[u][0] = f0.x; v[u][1] = f0.y; v[u][2] = f0.z; v[u][3] = f0.w;
          v[u][4] = f1.x; v[u][5] = f1.y; v[u][6] = f1.z; v[u][7] = f1.w;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (i0 + u * C5_PT < ITEMS) {
            const Split8 s = split8(v[u]);
            const size_t o = (size_t)q[u] * C5_LBO + (size_t)pix[u] * 16;
            *reinterpret_cast<uint4*>(a_hi + o) = s.hi;
            *reinterpret_cast<uint4*>(a_lo + o) = s.lo;
          }
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready);
    if (tid == 0) TRACE(2);
    // ===================== epilogue: warp w reads TMEM lanes 32*(w%4).. of position tiles w/4 and w/4 + 4
    mbar_wait(tfull, 0);
    tc_fence_after();
    // (the patches below alias the operand images other warps staged; tfull already orders them — every staging store
    // precedes a_ready, the MMAs and their commit — and this barrier states the same among the 16 warps themselves)
    asm volatile("bar.sync 1, %0;" ::"n"(C5_PT) : "memory");
    if (tid == 0) TRACE(6);
    // thread = position (TMEM lane): add the halves, bias, ReLU, park the 32 x Cout patch in shared memory (the operand
    // images are dead; pitch Cout*4 + 16 bytes keeps the 16-byte accesses conflict-free), then write it out with lanes
    // along the channels: 512 contiguous bytes per store instruction instead of 32 half-used sectors
    const int q4 = warp & 3;
    const bool p3 = a.passes == 3;
    const uint32_t pitch = (uint32_t)Cout * 4 + 16;
    uint8_t* patch = smem + (size_t)warp * 32 * (32 * 4 + 16);   // 16 x 4.5 KB < the smaller operand image pair (91 KB)
    const int cpr = Cout / 4, rpi = 32 / cpr;                    // 16-byte chunks per row, rows per store instruction
#pragma unroll 1
    for (int j = 0; j < C5_MT / 4; ++j) {
      const int mt = (warp >> 2) + 4 * j;
      if (mt >= mt_used) break;                                  // (warp-uniform)
      const uint32_t taddr = tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(mt * 2 * Cout);
#pragma unroll 1
      for (int n0 = 0; n0 < Cout; n0 += 16) {
        float acc[16], acl[16];
        tmem_ld16(taddr + n0, acc);
        if (p3) tmem_ld16(taddr + Cout + n0, acl);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 b4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 y;
          if (p3) y = make_float4(acc[e] + acl[e], acc[e + 1] + acl[e + 1], acc[e + 2] + acl[e + 2], acc[e + 3] + acl[e + 3]);
          else y = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
          y.x += b4.x; y.y += b4.y; y.z += b4.z; y.w += b4.w;
          if (a.act == DESIRE_ACT_RELU) {
            y.x = fmaxf(y.x, 0.f);
            y.y = fmaxf(y.y, 0.f);
            y.z = fmaxf(y.z, 0.f);
            y.w = fmaxf(y.w, 0.f);
          }
          *reinterpret_cast<float4*>(patch + lane * pitch + (n0 + e) * 4) = y;
        }
      }
      __syncwarp();
      const int ch = lane % cpr, rsub = lane / cpr;
#pragma unroll 1
      for (int it = 0; it < 32; it += rpi) {
        const int row = it + rsub;
        const int p = mt * 128 + 32 * q4 + row;
        const int r = p / C5_WP, c = p - r * C5_WP;
        const int oy = y0 + r, ox = x0 + c;
        if (r < C5_TH && c < C5_TW && oy < a.H && ox < a.W) {
          const float4 y = *reinterpret_cast<const float4*>(patch + row * pitch + ch * 16);
          *reinterpret_cast<float4*>(a.Y + (((size_t)img * a.H + oy) * a.W + ox) * a.ldc + ch * 4) = y;
        }
      }
      __syncwarp();
    }
  }
  if (tid == 0) TRACE(7);
  tc_fence_before();
  __syncthreads();
  if (warp == C5_PW) tmem_dealloc(tmem, tcols);
  if (tid == 0) TRACE(8);
#undef TRACE
}

size_t c5_smem(int Cin, int Cout) {
  return (size_t)2 * (Cin / 8) * C5_LBO + (size_t)C5_NSLOT * 4 * Cin * Cout + (2 * C5_NSLOT + 2) * 8 + 16;
}

}  // namespace

size_t conv5_tc_pack_bytes(int Cin, int Cout) { return (size_t)25 * 4 * Cin * Cout; }

namespace {
bool c5_off() {
  static const bool off = [] {
    const char* e = getenv("DESIRE_NO_CONV5");
    return e && e[0] == '1';
  }();
  return off || gemm_mode() == 0;
}

template <int CIN, bool S2D>
int c5_launch(const C5Args& a0, int taps_ci, cudaStream_t st) {
  static long long* trace = nullptr;
  static const bool want_trace = [] {
    const char* e = getenv("DESIRE_CONV5_TRACE");
    return e && e[0] == '1';
  }();
  if (want_trace && !trace) DESIRE_CUDA(cudaMalloc(&trace, 16 * sizeof(long long)));
  C5Args a = a0;
  a.trace = want_trace ? trace : nullptr;
  const dim3 grid((unsigned)((a.W + C5_TW - 1) / C5_TW), (unsigned)((a.H + C5_TH - 1) / C5_TH), (unsigned)a.B);
  const size_t smem = c5_smem(CIN, a.Cout);
  DESIRE_ENSURE_SMEM((conv5_tc_kernel<CIN, S2D>), smem);
  DESIRE_LAUNCH(st, (conv5_tc_kernel<CIN, S2D><<<grid, C5_NTHR, smem, st>>>(a)));
  if (want_trace) {
    static int printed = 0;
    long long h[16];
    DESIRE_CUDA(cudaStreamSynchronize(st));
    DESIRE_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ < 2)
      fprintf(stderr, "conv5 trace Cin=%d%s (cycles from start): setup %lld staged %lld | mma: go %lld first tap issued %lld all issued %lld | "
              "accumulators complete %lld epilogue done %lld end %lld\n", taps_ci, S2D ? " space-to-depth" : "", h[1] - h[0], h[2] - h[0],
              h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0]);
  }
  return DESIRE_OK;
}
}  // namespace

// 5x5 / stride 1 / SAME over NHWC, Cin in {16, 32}, Cout in {16, 32} (8 x 2*Cout accumulator columns), act NONE or RELU
bool conv5_tc_eligible(const Im2col& g, int Cout, int ldc, int act, const PackWs& pw) {
  if (c5_off()) return false;
  return g.kh == 5 && g.kw == 5 && g.stride == 1 && g.pad_t == 2 && g.pad_l == 2 && g.Hi == g.Ho && g.Wi == g.Wo &&
         (g.Ci == 16 || g.Ci == 32) && (Cout == 16 || Cout == 32) && ldc % 4 == 0 &&
         (act == DESIRE_ACT_NONE || act == DESIRE_ACT_RELU) && pw.p && pw.bytes >= conv5_tc_pack_bytes(g.Ci, Cout) &&
         g.Ho <= 65535 * C5_TH;
}

// 5x5 / stride 2 / SAME over a 3-channel image with even sides (TF: one pixel of padding before): the space-to-depth form
bool conv5s2_tc_eligible(const Im2col& g, int Cout, int ldc, int act, const PackWs& pw) {
  if (c5_off()) return false;
  return g.kh == 5 && g.kw == 5 && g.stride == 2 && g.pad_t == 1 && g.pad_l == 1 && g.Ci == 3 && g.Hi % 2 == 0 &&
         g.Wi % 2 == 0 && g.Ho == g.Hi / 2 && g.Wo == g.Wi / 2 && (Cout == 16 || Cout == 32) && ldc % 4 == 0 &&
         (act == DESIRE_ACT_NONE || act == DESIRE_ACT_RELU) && pw.p && pw.bytes >= conv5_tc_pack_bytes(16, Cout) &&
         g.Ho <= 65535 * C5_TH;
}

int conv5_tc(const float* X, const Im2col& g, int B, const float* w, int ldb, const float* bias, float* Y, int ldc, int Cout,
             int act, cudaStream_t st, PackWs pw) {
  if (B <= 0) return DESIRE_OK;
  const bool s2d = g.stride == 2;
  DESIRE_CHECK_ARG((s2d ? conv5s2_tc_eligible(g, Cout, ldc, act, pw) : conv5_tc_eligible(g, Cout, ldc, act, pw)) && B <= 65535,
                   "conv5_tc: not eligible");
  DESIRE_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                       (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
                   "conv5_tc: X, Y and bias must be 16-byte aligned");
  uint8_t* packed = reinterpret_cast<uint8_t*>(pw.p);
  C5Args a{X, packed, bias, Y, B, g.Ho, g.Wo, s2d ? 16 : g.Ci, Cout, ldc, act, gemm_mode() == 1 ? 1 : 3, nullptr};
  if (s2d) {
    DESIRE_LAUNCH(st, (c5_pack_s2d_kernel<<<std::min(148, (9 * 16 * Cout + 255) / 256), 256, 0, st>>>(w, ldb, Cout, packed)));
    return c5_launch<16, true>(a, 3, st);
  }
  DESIRE_LAUNCH(st, (c5_pack_kernel<<<std::min(148, (25 * g.Ci * Cout + 255) / 256), 256, 0, st>>>(w, ldb, g.Ci, Cout, packed)));
  if (g.Ci == 16) return c5_launch<16, false>(a, 16, st);
  return c5_launch<32, false>(a, 32, st);
}

}  // namespace desire
