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
//   warps 0-15 producers (16 warps so every SM sub-partition has four to hide shared-memory latency).  Warp w owns
//              tile rows 8w..8w+7 and 8 lanes cooperate on a row (one 16-byte chunk of a 64-column K stage each):
//              sum the listed neighbours' slices from shared memory in list order, scale by 1/count, split to
//              BF16 hi/lo, store as the UMMA A operand.  Neighbour loops diverge over 4 rows, not 32.
//              Afterwards: epilogue (bias + ReLU) of a 32x32 accumulator block each;
//   warp 16    tcgen05.mma issuer (M=128, N=H, 3xBF16), accumulator in TMEM;
//   warp 17    streams the packed sp_w stages with 1-D bulk TMA copies.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int TM = 128, BK = 64, KC = 8, STAGES = 2;   // a stage = 64 columns of one bin = two packed 32-col blocks
constexpr int NPW = 16;                 // producer/epilogue warps: 4 threads per row, one 8-column chunk each
constexpr int NTHR = (NPW + 2) * 32;
constexpr int MAXG = 64;

struct Layout {
  int hs_ld;            // floats per staged hidden row
  size_t hs, px, py, rowmap, tab, lists, list_stride, bins, stages, stage_bytes, bars, total;
};
__host__ __device__ inline Layout make_layout(int H, int Npad, int G, int n_rad, int n_ang) {
  Layout L;
  L.hs_ld = H + 4;
  size_t off = 0;
  L.hs = off; off += (size_t)TM * L.hs_ld * 4;
  L.px = off; off += TM * 4;
  L.py = off; off += TM * 4;
  L.rowmap = off; off += TM * 8;                                 // global row of every tile lane (-1 = none)
  L.tab = off; off += (size_t)((n_rad + 1 + 2 * n_ang + 3) / 4 * 4) * 4;
  L.list_stride = (size_t)((G + 1 + Npad + 3) / 4 * 4);          // bytes per row: off[G+1] then list[Npad]
  L.lists = off; off += TM * L.list_stride;
  L.bins = off; off += (size_t)TM * (Npad + 4);                 // bin of every (row, neighbour) pair, 255 = none; the row
                                                                // stride Npad+4 bytes keeps the 32 rows a warp writes on 32 banks
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
  long* rowmap = reinterpret_cast<long*>(smem + L.rowmap);
  uint8_t* lists = smem + L.lists;
  uint8_t* bins = smem + L.bins;
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
  const int a_half = KC * TM * 16;          // bytes of A_hi per stage (8 chunk planes)
  const int b_blk = 4 * H * 16;             // bytes of one packed hi (or lo) block of 32 columns
  const int b_half = 2 * b_blk;             // per stage: two blocks, each hi+lo => 2*b_half bytes in total

  // global row of tile lane l (or -1): group grp0 + l/Npad = (b, k), agent i = l % Npad
  auto row_of = [&](int l) -> long {
    const long grp = grp0 + l / Npad;
    const int i = l % Npad;
    if (grp >= ngroups || i >= N) return -1;
    const long b = grp / K;
    const int k = (int)(grp % K);
    return (b * N + i) * K + k;
  };

  if (tid < TM) rowmap[tid] = row_of(tid);     // the 64-bit divisions happen once per lane, not per element
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], NPW + 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == NPW) tmem_alloc_dyn(tslot, tmem_cols);
  __syncthreads();

  // ---- prologue: stage tables, positions (NaN = non-existent / padding) and hidden vectors
  for (int e = tid; e < a.n_rad + 1; e += NTHR) tab[e] = __ldg(a.r2_edges + e);
  for (int e = tid; e < 2 * a.n_ang; e += NTHR) tab[a.n_rad + 1 + e] = __ldg(a.dirs + e);
  for (int l = tid; l < TM; l += NTHR) {
    const long r = rowmap[l];
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
      const long r = rowmap[l];
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
    // ===================== binning, all 512 threads: thread (row rl, quarter q) bins neighbours j = q, q+4, ...
    {
      const int rl = tid & (TM - 1), q = tid >> 7;
      const long r = rowmap[rl];
      const int gbase = (rl / Npad) * Npad, me = rl % Npad;
      uint8_t* brow = bins + (size_t)rl * (Npad + 4);
      // a masked row still pools its existing neighbours: its own position comes from global memory
      float xi = 0.f, yi = 0.f;
      if (r >= 0) {
        xi = __ldg(a.pos + r * a.pos_stride);
        yi = __ldg(a.pos + r * a.pos_stride + 1);
      }
      for (int j = q; j < N; j += 4) {
        int g = -1;
        if (r >= 0 && j != me) {
          const float dx = px[gbase + j] - xi, dy = py[gbase + j] - yi;
          if (dx == dx) g = logpolar_bin(dx, dy, tab, a.n_rad, tab + a.n_rad + 1, a.n_ang);
        }
        brow[j] = (uint8_t)g;                                   // 255 = no bin
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");
    // ===================== per-row counting sort into per-bin lists (ascending j), one thread per row
    if (tid < TM) {
      uint8_t* off = lists + (size_t)tid * L.list_stride;       // [G+1]
      uint8_t* lst = off + G + 1;                                // [Npad]
      const uint8_t* brow = bins + (size_t)tid * (Npad + 4);
      for (int g = 0; g <= G; ++g) off[g] = 0;
      for (int j = 0; j < N; ++j) {                              // counts, stored at off[g+1]
        const int g = brow[j];
        if (g < G) off[g + 1]++;
      }
      for (int g = 0; g < G; ++g) off[g + 1] += off[g];          // prefix: off[g] = start of bin g
      for (int j = 0; j < N; ++j) {                              // fill (off[g] is the cursor)
        const int g = brow[j];
        if (g < G) lst[off[g]++] = (uint8_t)j;
      }
      for (int g = G; g > 0; --g) off[g] = off[g - 1];           // cursors ended at the next bin's start: shift back
      off[0] = 0;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");

    // ===================== A producer.  Warp w owns tile rows 8w..8w+7: lane -> (row 8w + lane%8, chunk pair
    // lane/8 and 4 + lane/8 of the 64-column stage).  A quarter-warp therefore stores 8 consecutive rows of one
    // chunk plane = 128 contiguous bytes (conflict-free), and every lane converts exactly the 16 values it stores.
    const int prow = warp * 8 + (lane & 7);
    const int kq = lane >> 3;
    const uint8_t* off = lists + (size_t)prow * L.list_stride;
    const uint8_t* lst = off + G + 1;
    const float* hrow = hs + (size_t)((prow / Npad) * Npad) * L.hs_ld + kq * 8;
    const uint32_t aoff = kq * TM * 16 + prow * 16;
    // list metadata is prefetched one bin ahead (off[] is a prefix array: the next bin starts where this one ends), so a
    // stage never starts with a chain of dependent shared-memory loads
    const int spb = H / BK;                       // K stages per bin
    int g = 0, half = 0;
    int o0 = 0, o1 = off[1], o1n = G > 1 ? off[2] : 0;
    int jf = lst[0], jfn = lst[o1];              // first member of this / the next bin (unused when the bin is empty)
    for (int ks = 0; ks < nks; ++ks) {
      const int slot = ks % STAGES;
      const uint32_t ph = (ks / STAGES) & 1;
      const int col = half * BK;
      uint4 hi0 = make_uint4(0, 0, 0, 0), lo0 = hi0, hi1 = hi0, lo1 = hi0;
      if (o1 > o0) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        for (int o = o0; o < o1; ++o) {
          const int jm = (o == o0) ? jf : (int)lst[o];
          const float4* p = reinterpret_cast<const float4*>(hrow + (size_t)jm * L.hs_ld + col);
          const float4 x0 = p[0], y0 = p[1], x1 = p[8], y1 = p[9];      // chunks kq and kq+4 (32 floats apart)
          v[0] += x0.x; v[1] += x0.y; v[2] += x0.z; v[3] += x0.w;
          v[4] += y0.x; v[5] += y0.y; v[6] += y0.z; v[7] += y0.w;
          v[8] += x1.x; v[9] += x1.y; v[10] += x1.z; v[11] += x1.w;
          v[12] += y1.x; v[13] += y1.y; v[14] += y1.z; v[15] += y1.w;
        }
        if (o1 - o0 > 1) {
          const float inv = __frcp_rn((float)(o1 - o0));       // mean = sum * (1/count)
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= inv;
        }
        const Split8 s0 = split8(v), s1 = split8(v + 8);
        hi0 = s0.hi; lo0 = s0.lo; hi1 = s1.hi; lo1 = s1.lo;
      }
      mbar_wait(&empty[slot], ph ^ 1);
      uint8_t* sa = stages + (size_t)slot * L.stage_bytes + aoff;
      *reinterpret_cast<uint4*>(sa) = hi0;
      *reinterpret_cast<uint4*>(sa + 4 * TM * 16) = hi1;
      *reinterpret_cast<uint4*>(sa + a_half) = lo0;
      *reinterpret_cast<uint4*>(sa + a_half + 4 * TM * 16) = lo1;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
      if (++half == spb) {                        // next bin: shift the prefetched metadata, fetch the bin after it
        half = 0;
        ++g;
        o0 = o1;
        o1 = o1n;
        jf = jfn;
        o1n = g + 1 < G ? off[g + 2] : 0;
        jfn = lst[o1];
      }
    }
    const long myrow = rowmap[tid & (TM - 1)];

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
          // A: chunk planes 2j, 2j+1 of the stage; B: packed block j/2 = { hi [4][H][16B], lo [4][H][16B] }
          const uint32_t sbb = sb + (j >> 1) * (2 * b_blk) + (j & 1) * 2 * lbo_b;
          const uint64_t ahi = smem_desc(sa + j * 2 * lbo_a, lbo_a, 128);
          const uint64_t alo = smem_desc(sa + a_half + j * 2 * lbo_a, lbo_a, 128);
          const uint64_t bhi = smem_desc(sbb, lbo_b, 128);
          const uint64_t blo = smem_desc(sbb + b_blk, lbo_b, 128);
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

// DESIRE_SOCIAL_V1=1 keeps the first design (A operand in shared memory) for A/B timing
static bool use_ts(const SocialFcArgs& a) {
  static const bool v1 = [] {
    const char* e = getenv("DESIRE_SOCIAL_V1");
    return e && e[0] == '1';
  }();
  return !v1 && social_fc_ts_eligible(a);
}

bool social_fc_tc_eligible(const SocialFcArgs& a) {
  const int G = a.n_rad * a.n_ang;
  if (use_ts(a) || social_fc_fm_eligible(a)) return true;
  if (gemm_mode() == 0 || !a.packed) return false;
  if (a.H % 64 != 0 || a.H < 64 || a.H > 128 || a.ld_h % 4 != 0) return false;   // a 64-column stage never straddles bins
  if (a.N > 128 || a.N < 1 || G > MAXG || G < 1) return false;
  const Layout L = make_layout(a.H, npad_of(a.N), G, a.n_rad, a.n_ang);
  return L.total <= 227 * 1024;
}

int social_fc_tc(const SocialFcArgs& a, cudaStream_t st) {
  if (use_ts(a)) return social_fc_ts(a, st);
  if (social_fc_fm_eligible(a)) return social_fc_fm(a, st);
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
