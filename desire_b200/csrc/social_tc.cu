// Fused log-polar social pooling + fc on tcgen05 (IOC stage):
//     fsp[r, :] = relu( pool(h)[r, :] @ sp_w + sp_b ),   pool(h)[r, g*H + c] = mean_{j in bin g of r} h[j, c]
// The [R, G*H] pooled tensor (4*G*H bytes per row, 708 MB per step at the bench workload) is never written:
// it is the A operand of the GEMM and is assembled on the fly from shared memory.
//
// One CTA = one 128-row tile = 128/Npad complete (scene, sample) groups (Npad = N rounded up to a power of two),
// so every neighbour of every row of the tile is a row of the same tile:
//   prologue   the groups' hidden vectors ([128][H+4] FP32, padded rows) and positions are staged in shared
//              memory once; each producer thread (one row) bins its N-1 neighbours with the oracle's exact
//              arithmetic and counting-sorts them into a private per-bin list (ascending j);
//   warps 0-15 four threads per row, one 8-column chunk each (16 warps so every SM sub-partition has four warps
//              to hide shared-memory latency): per 32-wide K stage (= a 32-column slice of one bin) sum the listed
//              neighbours' slices from shared memory in list order, divide by the count, split to BF16 hi/lo,
//              store as the UMMA A operand; afterwards: epilogue (bias + ReLU) of a 32x32 accumulator block each;
//   warp 16    tcgen05.mma issuer (M=128, N=H, 3xBF16), accumulator in TMEM;
//   warp 17    streams the packed sp_w stages with 1-D bulk TMA copies.
#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int TM = 128, BK = 32, KC = 4, STAGES = 3;
constexpr int NPW = 16;                 // producer/epilogue warps: 4 threads per row, one 8-column chunk each
constexpr int NTHR = (NPW + 2) * 32;
constexpr int MAXG = 64;

struct Layout {
  int hs_ld;            // floats per staged hidden row
  size_t hs, px, py, tab, lists, list_stride, stages, stage_bytes, bars, total;
};
__host__ __device__ inline Layout make_layout(int H, int Npad, int G, int n_rad, int n_ang) {
  Layout L;
  L.hs_ld = H + 4;
  size_t off = 0;
  L.hs = off; off += (size_t)TM * L.hs_ld * 4;
  L.px = off; off += TM * 4;
  L.py = off; off += TM * 4;
  L.tab = off; off += (size_t)((n_rad + 1 + 2 * n_ang + 3) / 4 * 4) * 4;
  L.list_stride = (size_t)((G + 1 + Npad + 3) / 4 * 4);          // bytes per row: off[G+1] then list[Npad]
  L.lists = off; off += TM * L.list_stride;
  off = (off + 1023) / 1024 * 1024;
  L.stage_bytes = 2 * (size_t)KC * TM * 16 + 2 * (size_t)KC * H * 16;
  L.stages = off; off += STAGES * L.stage_bytes;
  L.bars = off; off += (2 * STAGES + 1) * 8 + 16;
  L.total = off;
  return L;
}

__global__ void __launch_bounds__(NTHR, 1) social_fc_tc_kernel(SocialFcArgs a, int Npad, int passes, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int H = a.H, N = a.N, K = a.K, G = a.n_rad * a.n_ang;
  const Layout L = make_layout(H, Npad, G, a.n_rad, a.n_ang);
  float* hs = reinterpret_cast<float*>(smem + L.hs);
  float* px = reinterpret_cast<float*>(smem + L.px);
  float* py = reinterpret_cast<float*>(smem + L.py);
  float* tab = reinterpret_cast<float*>(smem + L.tab);
  uint8_t* lists = smem + L.lists;
  uint8_t* stages = smem + L.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gpt = TM / Npad;                         // groups per tile
  const long ngroups = (long)a.B * K;
  const long grp0 = (long)blockIdx.x * gpt;
  const int nks = (G * H) / BK;
  const int a_half = KC * TM * 16, b_half = KC * H * 16;

  // global row of tile lane l (or -1): group grp0 + l/Npad = (b, k), agent i = l % Npad
  auto row_of = [&](int l) -> long {
    const long grp = grp0 + l / Npad;
    const int i = l % Npad;
    if (grp >= ngroups || i >= N) return -1;
    const long b = grp / K;
    const int k = (int)(grp % K);
    return (b * N + i) * K + k;
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], NPW + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == NPW) tmem_alloc_dyn(tslot, tmem_cols);

  // ---- prologue: stage tables, positions (NaN = non-existent / padding) and hidden vectors
  for (int e = tid; e < a.n_rad + 1; e += NTHR) tab[e] = __ldg(a.r2_edges + e);
  for (int e = tid; e < 2 * a.n_ang; e += NTHR) tab[a.n_rad + 1 + e] = __ldg(a.dirs + e);
  for (int l = tid; l < TM; l += NTHR) {
    const long r = row_of(l);
    float x = __int_as_float(0x7fc00000), y = x;
    if (r >= 0) {
      const long bn = r / K;                         // b*N + i
      if (__ldg(a.obs + (size_t)bn * a.Tp * 3) != 0.f) {
        x = __ldg(a.pos + r * a.pos_stride);
        y = __ldg(a.pos + r * a.pos_stride + 1);
      }
    }
    px[l] = x;
    py[l] = y;
  }
  {
    const int H4 = H / 4;
    for (int e = tid; e < TM * H4; e += NTHR) {
      const int l = e / H4, c4 = e - l * H4;
      const long r = row_of(l);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= 0) v = __ldg(reinterpret_cast<const float4*>(a.h + r * (long)a.ld_h) + c4);
      *reinterpret_cast<float4*>(hs + (size_t)l * L.hs_ld + c4 * 4) = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < NPW) {
    // ===================== 16 producer warps: thread = (row = tid % 128, chunk = tid / 128)
    const int rl = tid & (TM - 1), ch = tid >> 7;              // tile lane (row), 8-column chunk of the stage
    const long myrow = row_of(rl);
    uint8_t* off = lists + (size_t)rl * L.list_stride;         // [G+1]
    uint8_t* lst = off + G + 1;                                 // [Npad]
    const int gbase = (rl / Npad) * Npad;                       // first lane of my group
    if (ch == 0) {
      // ---- binning: counting sort of this row's neighbours into per-bin lists (ascending j)
      const int me = rl % Npad;
      // a masked row still pools its existing neighbours: its own position comes from global memory
      float xi = 0.f, yi = 0.f;
      if (myrow >= 0) {
        xi = __ldg(a.pos + myrow * a.pos_stride);
        yi = __ldg(a.pos + myrow * a.pos_stride + 1);
      }
      for (int g = 0; g <= G; ++g) off[g] = 0;
      if (myrow >= 0) {
        for (int j = 0; j < N; ++j) {                           // pass 1: counts (stored at off[g+1])
          if (j == me) continue;
          const float dx = px[gbase + j] - xi, dy = py[gbase + j] - yi;
          if (dx == dx) {
            const int g = logpolar_bin(dx, dy, tab, a.n_rad, tab + a.n_rad + 1, a.n_ang);
            if (g >= 0) off[g + 1]++;
          }
        }
        for (int g = 0; g < G; ++g) off[g + 1] += off[g];       // prefix: off[g] = start of bin g
        for (int j = 0; j < N; ++j) {                           // pass 2: fill (off[g] is the cursor)
          if (j == me) continue;
          const float dx = px[gbase + j] - xi, dy = py[gbase + j] - yi;
          if (dx == dx) {
            const int g = logpolar_bin(dx, dy, tab, a.n_rad, tab + a.n_rad + 1, a.n_ang);
            if (g >= 0) lst[off[g]++] = (uint8_t)j;
          }
        }
        for (int g = G; g > 0; --g) off[g] = off[g - 1];        // cursors ended at the next bin's start: shift back
        off[0] = 0;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");    // lists visible to the other chunk threads

    // ===================== A producer: this thread's 8 columns of a 32-column slice of one bin per stage
    const float* hgrp = hs + (size_t)gbase * L.hs_ld + ch * 8;
    const uint32_t a_off = ch * TM * 16 + rl * 16;
    for (int ks = 0; ks < nks; ++ks) {
      const int slot = ks % STAGES;
      const uint32_t ph = (ks / STAGES) & 1;
      const int k0 = ks * BK;
      const int g = k0 / H, col = k0 - g * H;
      const int o0 = off[g], o1 = off[g + 1];
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (o1 > o0) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        for (int o = o0; o < o1; ++o) {
          const float4* p = reinterpret_cast<const float4*>(hgrp + (size_t)lst[o] * L.hs_ld + col);
          const float4 x = p[0], y = p[1];
          v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
          v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
        }
        if (o1 - o0 > 1) {
          const float cnt = (float)(o1 - o0);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = v[i] / cnt;
        }
        const Split8 sp = split8(v);
        hi = sp.hi;
        lo = sp.lo;
      }
      mbar_wait(&empty[slot], ph ^ 1);
      uint8_t* sa = stages + (size_t)slot * L.stage_bytes + a_off;
      *reinterpret_cast<uint4*>(sa) = hi;
      *reinterpret_cast<uint4*>(sa + a_half) = lo;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
    }

    // ===================== epilogue: warp w -> TMEM lanes 32*(w%4).., columns 32*(w/4)..
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int c0 = (warp >> 2) * 32;
    if (c0 < H) {
      const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      float acc[32];
      tmem_ld32(trow + c0, acc);
      tmem_ld_wait();
      if (myrow >= 0) {
        float* orow = a.out + myrow * (long)H + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = fmaxf(acc[j] + __ldg(a.bias + c0 + j), 0.f);
          o.y = fmaxf(acc[j + 1] + __ldg(a.bias + c0 + j + 1), 0.f);
          o.z = fmaxf(acc[j + 2] + __ldg(a.bias + c0 + j + 2), 0.f);
          o.w = fmaxf(acc[j + 3] + __ldg(a.bias + c0 + j + 3), 0.f);
          *reinterpret_cast<float4*>(orow + j) = o;
        }
      }
    }
  } else if (warp == NPW) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(TM, H);
      const uint32_t lbo_a = TM * 16, lbo_b = H * 16;
      uint32_t accf = 0;
      for (int ks = 0; ks < nks; ++ks) {
        const int slot = ks % STAGES;
        mbar_wait(&full[slot], (ks / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(stages + (size_t)slot * L.stage_bytes);
        const uint32_t sb = sa + 2 * a_half;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t ahi = smem_desc(sa + j * 2 * lbo_a, lbo_a, 128);
          const uint64_t alo = smem_desc(sa + a_half + j * 2 * lbo_a, lbo_a, 128);
          const uint64_t bhi = smem_desc(sb + j * 2 * lbo_b, lbo_b, 128);
          const uint64_t blo = smem_desc(sb + b_half + j * 2 * lbo_b, lbo_b, 128);
          mma_bf16(tmem, ahi, bhi, idesc, accf);
          accf = 1;
          if (passes == 3) {
            mma_bf16(tmem, alo, bhi, idesc, 1);
            mma_bf16(tmem, ahi, blo, idesc, 1);
          }
        }
        mma_commit(&empty[slot]);
      }
      mma_commit(tfull);
    }
  } else {
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.packed);
      for (int ks = 0; ks < nks; ++ks) {
        const int slot = ks % STAGES;
        mbar_wait(&empty[slot], ((ks / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[slot], 2 * b_half);
        bulk_g2s(stages + (size_t)slot * L.stage_bytes + 2 * a_half, src + (size_t)ks * (2 * b_half), 2 * b_half,
                 &full[slot]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(tmem, tmem_cols);
}

int npad_of(int N) {
  int p = 8;
  while (p < N) p <<= 1;
  return p;
}

}  // namespace

bool social_fc_tc_eligible(const SocialFcArgs& a) {
  const int G = a.n_rad * a.n_ang;
  if (gemm_mode() == 0 || !a.packed) return false;
  if (a.H % 32 != 0 || a.H < 32 || a.H > 128 || a.ld_h % 4 != 0) return false;
  if (a.N > 128 || a.N < 1 || G > MAXG || G < 1) return false;
  const Layout L = make_layout(a.H, npad_of(a.N), G, a.n_rad, a.n_ang);
  return L.total <= 227 * 1024;
}

int social_fc_tc(const SocialFcArgs& a, cudaStream_t st) {
  const int Npad = npad_of(a.N), G = a.n_rad * a.n_ang;
  const Layout L = make_layout(a.H, Npad, G, a.n_rad, a.n_ang);
  const long ngroups = (long)a.B * a.K;
  if (ngroups == 0) return DESIRE_OK;
  const int gpt = TM / Npad;
  uint32_t cols = 32;
  while ((int)cols < a.H) cols <<= 1;
  DESIRE_ENSURE_SMEM(social_fc_tc_kernel, L.total);
  const unsigned grid = (unsigned)((ngroups + gpt - 1) / gpt);
  DESIRE_LAUNCH(st, (social_fc_tc_kernel<<<grid, NTHR, L.total, st>>>(a, Npad, gemm_mode() == 1 ? 1 : 3, cols)));
  return DESIRE_OK;
}

}  // namespace desire
