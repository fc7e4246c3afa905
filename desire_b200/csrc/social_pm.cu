// Log-polar social pooling for scenes of 129..256 agents as a tcgen05 MMA (the materialising entry point
// desire_social_pool_fwd / the IOC path when the fully fused kernel of social_ts.cu does not apply: N > 128 or H = 256):
//     pooled[r, g*H + c] = mean_{j in bin g of r} h[j, c]  =  (1 / count) * (S_g @ h)[r, c]
// S_g is the 0/1 selection matrix of bin g (rows = 128 agents of a (scene, sample) group, columns = the group's up to 256
// agents as neighbours; exact in BF16), h the group's hidden vectors as BF16 hi + lo: two passes, FP32 accumulation in
// tensor memory = the exact sum of the split values (|error| <= 2^-17 per element).  The SIMT row-block kernel in ioc.cu
// moves R*N*H*4 bytes through shared memory for the same sums (cfg3: 2 TB per step) and ran at 0.28 of the HBM peak.
//
// One CTA = 128 rows (agents i0 .. i0+127) of one group.  The hidden dimension is walked in passes of 64 columns: the
// pass's h^T slice (64 columns x 256 neighbours, hi + lo = 64 KB) is the B operand in shared memory; per bin the 16
// builder/finisher warps write S_g straight into tensor memory (tcgen05.st, thread = row: one byte compare per pair
// against the bin ids the prologue stored), warp 16 issues the 32 MMAs (A from tensor memory), the finishers read the
// 64-column accumulator, scale by 1/count, transpose their block through shared memory and write 64-byte row segments.
// TMEM columns: S0 [0,128) | S1 [128,256) | P0 [256,320) | P1 [320,384).
#include <stdlib.h>

#include "common.cuh"
#include "social_common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;
using namespace social;

constexpr int TM = 128;
constexpr int NPW = 16;
constexpr int PT = NPW * 32;
constexpr int NTHR = (NPW + 1) * 32;
constexpr int NC = 64;           // columns of h per pass
constexpr int NJ = 256;          // neighbours (padded)
constexpr int SLD = NC + 4;      // floats per staged output row

struct Layout {
  size_t ht, bins, bin_stride, pc, stage, px, py, exb, tab, bars, total;
};
__host__ __device__ inline Layout make_layout() {
  Layout L;
  size_t off = 0;
  L.ht = off; off += 2 * (size_t)NC * NJ * 2;                    // h^T slice, K-major [NC x NJ] BF16: hi, lo
  L.bin_stride = NJ + 16;
  L.bins = off; off += TM * L.bin_stride;
  L.pc = off; off += 4 * 4 * TM * 4;                             // partial counts [4 buffers][4 parts][128 rows]
  L.stage = off; off += (size_t)NPW * 32 * 20 * 4;               // per warp: a [32 x 16] block being transposed (row pitch 20)
  L.px = off; off += NJ * 4;
  L.py = off; off += NJ * 4;
  L.exb = off; off += NJ;
  L.tab = off; off += 24 * 4;
  L.bars = off; off += 9 * 8 + 32;
  L.total = off;
  return L;
}

struct PmArgs {
  const float* pos;
  long pos_stride;
  const float* h;
  int ld_h;
  const float* obs;
  int Tp, N, K, H, n_rad, n_ang, nrb;
  const float *r2_edges, *dirs;
  float* pooled;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

template <bool P3>
__global__ void __launch_bounds__(NTHR, 1) social_pool_mma_kernel(PmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Layout L = make_layout();
  uint8_t* ht = smem + L.ht;
  uint8_t* bins = smem + L.bins;
  int* pc = reinterpret_cast<int*>(smem + L.pc);
  float* stage = reinterpret_cast<float*>(smem + L.stage);
  float* px = reinterpret_cast<float*>(smem + L.px);
  float* py = reinterpret_cast<float*>(smem + L.py);
  uint8_t* exb = smem + L.exb;
  float* tab = reinterpret_cast<float*>(smem + L.tab);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(smem + L.bars);   // [2] selection matrix written (16 warps)
  uint64_t* sempty = sfull + 2;                                   // [2] ... read by its MMAs (commit)
  uint64_t* pfull = sempty + 2;                                   // [2] accumulator complete (commit)
  uint64_t* pempty = pfull + 2;                                   // [2] ... read into registers (16 warps)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(pempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, K = a.K, H = a.H, G = a.n_rad * a.n_ang;
  const long grp = blockIdx.x / a.nrb;
  const int i0 = (int)(blockIdx.x % a.nrb) * TM;
  const long b = grp / K;
  const int k = (int)(grp % K);
  const int npass = H / NC;

  for (int j = tid; j < NJ; j += NTHR) {                          // positions / existence of the group's agents
    float x = 0.f, y = 0.f;
    bool ex = false;
    if (j < N) {
      const long rj = (b * N + j) * K + k;
      ex = __ldg(a.obs + (size_t)(b * N + j) * a.Tp * 3) != 0.f;
      x = __ldg(a.pos + rj * a.pos_stride);
      y = __ldg(a.pos + rj * a.pos_stride + 1);
    }
    px[j] = x;
    py[j] = y;
    exb[j] = ex ? 1 : 0;
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sfull[s], NPW);
      mbar_init(&sempty[s], 1);
      mbar_init(&pfull[s], 1);
      mbar_init(&pempty[s], NPW);
    }
    fence_barrier_init();
  }
  if (warp == NPW) tmem_alloc<512>(tslot);
  tc_fence_before();
  __syncthreads();

  if (warp == NPW) {
    // ===================== pool MMA issuer (whole warp, elected lane, descriptors = base + constant offsets)
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
    constexpr uint32_t idesc = idesc_bf16(TM, NC);
    constexpr uint32_t lbo_b = NC * 16;
    const uint64_t d_hh = smem_desc(smem_u32(ht), lbo_b, 128), d_hl = desc_adv(d_hh, NC * NJ * 2);
    const int total = G * npass;                                  // every bin is a stage (an empty one writes zeros)
    for (int sg = 0; sg < total; ++sg) {
      const int sb = sg & 1;
      const uint32_t t_s = tmem + sb * (NJ / 2), t_p = tmem + NJ + sb * NC;
      mbar_wait(&sfull[sb], (sg >> 1) & 1);                       // (=> this pass's h^T slice is in place as well)
      if (sg >= 2) mbar_wait(&pempty[sb], ((sg - 2) >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < NJ / 16; ++j) {
          mma_bf16_ts(t_p, t_s + 8 * j, desc_adv(d_hh, j * 2 * lbo_b), idesc, j > 0);
          if (P3) mma_bf16_ts(t_p, t_s + 8 * j, desc_adv(d_hl, j * 2 * lbo_b), idesc, 1);
        }
        mma_commit(&pfull[sb]);
        mma_commit(&sempty[sb]);
      }
    }
    __syncwarp();
  } else {
    // ===================== prologue: tables, bins of the 128 x N pairs
    if (tid < 8) tab[tid] = tid <= a.n_rad ? __ldg(a.r2_edges + tid) : __int_as_float(0x7f800000);
    if (tid >= 32 && tid < 48) tab[8 + tid - 32] = tid - 32 < 2 * a.n_ang ? __ldg(a.dirs + tid - 32) : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");
    {
      const int rl = tid & (TM - 1), q = tid >> 7;
      const int i = i0 + rl;
      const bool valid = i < N;
      const float xi = valid ? px[i] : 0.f, yi = valid ? py[i] : 0.f;
      float re[8], dr[16];
#pragma unroll
      for (int e = 0; e < 8; e += 4) *reinterpret_cast<float4*>(re + e) = *reinterpret_cast<const float4*>(tab + e);
#pragma unroll
      for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(dr + e) = *reinterpret_cast<const float4*>(tab + 8 + e);
#pragma unroll
      for (int e = 0; e < 8; ++e) asm volatile("" : "+f"(re[e]));
#pragma unroll
      for (int e = 0; e < 16; ++e) asm volatile("" : "+f"(dr[e]));
      uint8_t* brow = bins + (size_t)rl * L.bin_stride;
      for (int j = q; j < NJ; j += 4) {
        const bool on = valid && j < N && j != i && exb[j];       // a masked row still pools its existing neighbours
        int gf, gb;
        logpolar_bin_pair(px[j] - xi, py[j] - yi, re, dr, a.n_rad, a.n_ang, gf, gb);
        brow[j] = (uint8_t)(on ? gf : -1);                        // 255 = no bin
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const int nst = G;

    // thread = TMEM lane (row 32*(warp%4) + lane); part = warp/4: 64 neighbours (32 packed columns) of S, 16 columns of P
    const int q4 = warp & 3, part = warp >> 2;
    const int frow = 32 * q4 + lane;
    const uint32_t lane_f = (uint32_t)(32 * q4) << 16;
    const uint8_t* brow = bins + (size_t)frow * L.bin_stride + 64 * part;
    auto build = [&](int sg, int bin) {                           // selection matrix + partial counts of stage sg
      const uint32_t g4 = (uint32_t)bin * 0x01010101u;
      uint32_t sr[32];
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 w = *reinterpret_cast<const uint4*>(brow + 16 * c);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t z = eq_bytes(ww[e], g4);
          cnt += __popc(z);
          sr[8 * c + 2 * e] = ones_lo(z);
          sr[8 * c + 2 * e + 1] = ones_hi(z);
        }
      }
      pc[((sg & 3) * 4 + part) * TM + frow] = cnt;
      if (sg >= 2) mbar_wait(&sempty[sg & 1], ((sg - 2) >> 1) & 1);   // the MMAs of stage sg-2 have read this buffer
      tc_fence_after();
      tmem_st32(tmem + lane_f + (sg & 1) * (NJ / 2) + part * 32, sr);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[sg & 1]);
    };

    int sg = 0;
    for (int p = 0; p < npass; ++p) {
      // ---- h^T slice of this pass (columns 64p .. 64p+63): byte(c, j) = (j/8) * NC*16 + c*16 + (j%8)*2.  Every warp has
      // seen the last accumulator of the previous pass complete, so nothing reads the old slice any more.
      constexpr int ITEMS = NC * (NJ / 8) / PT;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int item = it * PT + tid;
        const int c = item % NC, oct = item / NC;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = oct * 8 + e;
          v[e] = j < N ? __ldg(a.h + ((b * N + j) * K + k) * (long)a.ld_h + p * NC + c) : 0.f;
        }
        const Split8 s8 = split8(v);
        const size_t o = (size_t)oct * NC * 16 + (size_t)c * 16;
        *reinterpret_cast<uint4*>(ht + o) = s8.hi;
        *reinterpret_cast<uint4*>(ht + (size_t)NC * NJ * 2 + o) = s8.lo;
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");

      build(sg, 0);
      for (int s = 0; s < nst; ++s, ++sg) {
        const int bin = s;
        if (s + 1 < nst) build(sg + 1, s + 1);
        mbar_wait(&pfull[sg & 1], (sg >> 1) & 1);
        tc_fence_after();
        const int* pcg = pc + (size_t)(sg & 3) * 4 * TM + frow;
        const int cnt = pcg[0] + pcg[TM] + pcg[2 * TM] + pcg[3 * TM];
        float v[16];
        tmem_ld16(tmem + lane_f + NJ + (sg & 1) * NC + part * 16, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pempty[sg & 1]);
        if (cnt > 1) {
          const float inv = __frcp_rn((float)cnt);               // mean = sum * (1/count)
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= inv;
        }
        // The warp transposes its own [32 rows x 16 columns] block through a private patch of shared memory (no block-wide
        // barrier: the warps stay decoupled) and writes 64-byte row segments: four lanes per row, eight rows per instruction.
        float* st = stage + (size_t)warp * 32 * 20;              // [32][16 + 4] floats
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(st + lane * 20 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = 8 * it + (lane >> 2), c4 = lane & 3;
          const int i = i0 + 32 * q4 + rr;
          if (i < N) {
            const long r = (b * N + i) * K + k;
            __stcs(reinterpret_cast<float4*>(a.pooled + r * (long)G * H + (long)bin * H + p * NC + part * 16) + c4,
                   *reinterpret_cast<const float4*>(st + rr * 20 + c4 * 4));
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(*tslot, 512);
}

}  // namespace

// DESIRE_POOL_ROWS=1 keeps the SIMT row-block kernel (A/B timing)
bool social_pool_mma_eligible(const float* h, int ld_h, int N, int H, int n_rad, int n_ang, const float* pooled) {
  static const bool off = [] {
    const char* e = getenv("DESIRE_POOL_ROWS");
    return e && e[0] == '1';
  }();
  if (off || gemm_mode() == 0) return false;
  if (N <= 128 || N > NJ || H % NC != 0 || n_rad > 7 || n_ang > 8 || n_rad * n_ang > 64) return false;
  return (reinterpret_cast<uintptr_t>(pooled) & 15) == 0 && h != nullptr && ld_h >= H;
}

int social_pool_mma(const float* pos, long pos_stride, const float* h, int ld_h, const float* obs, int Tp, int B, int N,
                    int K, int H, int n_rad, int n_ang, const float* r2_edges, const float* dirs, float* pooled,
                    cudaStream_t st) {
  const long ngroups = (long)B * K;
  if (ngroups == 0) return DESIRE_OK;
  PmArgs a{};
  a.pos = pos; a.pos_stride = pos_stride; a.h = h; a.ld_h = ld_h; a.obs = obs; a.Tp = Tp; a.N = N; a.K = K; a.H = H;
  a.n_rad = n_rad; a.n_ang = n_ang; a.nrb = (N + TM - 1) / TM; a.r2_edges = r2_edges; a.dirs = dirs; a.pooled = pooled;
  const Layout L = make_layout();
  const long grid = ngroups * a.nrb;
  DESIRE_CHECK_ARG(grid < (1L << 31), "social_pool_mma: grid too large");
  if (gemm_mode() == 1) {
    DESIRE_ENSURE_SMEM(social_pool_mma_kernel<false>, L.total);
    DESIRE_LAUNCH(st, (social_pool_mma_kernel<false><<<(unsigned)grid, NTHR, L.total, st>>>(a)));
  } else {
    DESIRE_ENSURE_SMEM(social_pool_mma_kernel<true>, L.total);
    DESIRE_LAUNCH(st, (social_pool_mma_kernel<true><<<(unsigned)grid, NTHR, L.total, st>>>(a)));
  }
  return DESIRE_OK;
}

}  // namespace desire
