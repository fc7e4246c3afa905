// Fused log-polar social pooling + fc, second design: POOLING ITSELF RUNS ON THE TENSOR CORE and the pooled A operand of
// the fc lives in TENSOR MEMORY.
//     fsp[r, :] = relu( pool(h)[r, :] @ sp_w + sp_b ),   pool(h)[r, g*H + c] = mean_{j in bin g of r} h[j, c]
//
// Why (profiles/r2l_social_ts_trace.txt): gathering the neighbours' rows with SIMT loads costs 512 bytes of
// shared-memory traffic per (row, neighbour) pair next to the MMA's own operand fetch, and its loops diverge (lists of
// different rows have different lengths: 32 % of the lanes active) — both the first design (A operand assembled in shared
// memory) and a gather that wrote straight to tensor memory ran at ~120 us per tile where the MMAs need 28.  Here, per bin g:
//   pool MMA   P[128 x H] = S_g[128 x 128] @ h[128 x H]: S_g is the 0/1 selection matrix of the bin (rows = tile rows,
//              columns = tile rows as neighbours; exact in BF16), h the tile's hidden vectors as BF16 hi + lo (two
//              passes), both in shared memory; P accumulates in FP32 in tensor memory: the exact sum of hi + lo;
//   finish     16 warps (thread = row, 32 columns each): tcgen05.ld P, scale by 1/count, split to BF16 hi/lo,
//              tcgen05.st as the A operand of the fc (lane = row, one 32-bit column = two consecutive K values);
//   fc MMA     D[128 x H] += A_g[128 x H] @ sp_w[g*H.., :]  (A from tensor memory, 3xBF16, weights streamed by bulk TMA).
// The tensor pipe alternates pool(g+1) | fc(g); no loop of the kernel depends on how many neighbours a row has.
// Building S_g is one byte compare per (row, neighbour): the prologue stores the bin of every pair as a byte.
//
// One CTA = one 128-row tile = 128/Npad complete (scene, sample) groups.  TMEM columns: D [0,H) | P [H,2H) | A0 | A1.
//   warps 0-15  prologue (bins, transposed hidden vectors), S builder, finisher, epilogue (bias + ReLU)
//   warp 16     issues the fc MMAs, warp 18 the pool MMAs
//   warp 17     streams the packed sp_w blocks (32 K values = one slot) with 1-D bulk TMA copies
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
constexpr int NTHR = (NPW + 3) * 32;
constexpr int PT = NPW * 32;   // producer threads
constexpr int MAXG = 64;
constexpr int MAXNB = 8;       // weight slots in shared memory (at most)

struct Layout {
  size_t ht, sm, bins, bin_stride, pc, px, py, rowmap, tab, bars, ring, slot_bytes, total;
  int nb;
};
__host__ __device__ inline Layout make_layout(int H, int Npad, int n_rad, int n_ang) {
  Layout L;
  size_t off = 0;
  L.ht = off; off += 2 * (size_t)H * TM * 2;                    // h as the pool MMA's B operand [H x 128] K-major: hi, lo
  L.sm = off; off += 2 * (size_t)TM * TM * 2;                   // two selection matrices [128 x 128] BF16, K-major
  L.bin_stride = (size_t)Npad + 16;                             // (+16: eight consecutive rows start on eight bank groups)
  L.bins = off; off += TM * L.bin_stride;                       // bin of every (row, neighbour) pair, 255 = none
  L.pc = off; off += 3 * 4 * TM * 4;                            // partial neighbour counts [3 buffers][4 parts][128 rows]
  L.px = off; off += TM * 4;
  L.py = off; off += TM * 4;
  L.rowmap = off; off += TM * 8;
  off = (off + 15) / 16 * 16;
  L.tab = off; off += 24 * 4;                                   // 8 squared radial edges (+inf padded), 8 directions
  L.bars = off; off += (2 * MAXNB + 13) * 8 + 40;              // barriers, tensor-memory slot, existence bits, bin mask
  off = (off + 1023) / 1024 * 1024;
  L.slot_bytes = 2 * (size_t)4 * H * 16;                        // one packed block of 32 K values: hi + lo
  L.ring = off;
  long room = 227L * 1024 - (long)off;
  int nb = room > 0 ? (int)(room / (long)L.slot_bytes) : 0;
  L.nb = nb > MAXNB ? MAXNB : nb;
  L.total = off + (size_t)L.nb * L.slot_bytes;
  return L;
}

template <int H, bool P3>
__global__ void __launch_bounds__(NTHR, 1) social_fc_ts_kernel(SocialFcArgs a, int Npad, int dbg,
                                                               long long* __restrict__ trace) {
  constexpr uint32_t TCOLS = 4 * H <= 256 ? 256 : 512;
  static_assert(4 * H <= 512, "tensor memory: D + P + two A stages");
  constexpr int CW = H / 4;                          // columns of P a finisher warp converts
  extern __shared__ __align__(1024) uint8_t smem[];
  const int N = a.N, K = a.K, G = a.n_rad * a.n_ang;
  const Layout L = make_layout(H, Npad, a.n_rad, a.n_ang);
  uint8_t* ht = smem + L.ht;
  uint8_t* sm = smem + L.sm;
  uint8_t* bins = smem + L.bins;
  int* pc = reinterpret_cast<int*>(smem + L.pc);
  float* px = reinterpret_cast<float*>(smem + L.px);
  float* py = reinterpret_cast<float*>(smem + L.py);
  float* tab = reinterpret_cast<float*>(smem + L.tab);
  long* rowmap = reinterpret_cast<long*>(smem + L.rowmap);
  uint8_t* ring = smem + L.ring;
  uint64_t* bfull = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* bempty = bfull + MAXNB;
  uint64_t* sfull = bempty + MAXNB;                  // [2] selection matrix written (16 warps)
  uint64_t* pissued = sfull + 2;                     // [3] the pool MMAs of the next stage are in the queue (see the fc issuer)
  uint64_t* afull = pissued + 3;                     // [2] A stage written (16 warps)
  uint64_t* aempty = afull + 2;                      // [2] ... consumed by its fc MMAs (commit)
  uint64_t* pfull = aempty + 2;                      // pool MMAs of a bin complete (commit)
  uint64_t* pempty = pfull + 1;                      // P read into registers (16 warps)
  uint64_t* tfull = pempty + 1;
  uint64_t* lready = tfull + 1;                      // the bin mask of the tile is complete (16 warps)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(lready + 1);
  uint32_t* exist = tslot + 1;                       // bit l of word l/32: tile lane l is an existing agent
  uint32_t* binmask = exist + 4;                     // bit g: some (row, neighbour) pair of the tile falls into bin g
  const int nb = L.nb;
  const bool tr = trace != nullptr && blockIdx.x == 0;
#define TRACE(i) do { if (tr) trace[i] = clock64(); } while (0)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gpt = TM / Npad;                         // groups per tile
  const long ngroups = (long)a.B * K;
  const long grp0 = (long)blockIdx.x * gpt;
  constexpr int CPB = H / 32;                        // weight blocks (slots) per bin
  const int nblk = G * CPB;
  constexpr uint32_t b_blk = 4 * H * 16;                 // bytes of the hi (or lo) half of a weight block
  if (tid == 0) TRACE(0);

  if (tid < TM) {                                    // global row of tile lane l (or -1), its position, its existence
    const long grp = grp0 + tid / Npad;
    const int i = tid % Npad;
    long r = -1;
    float x = 0.f, y = 0.f;
    bool ex = false;
    if (grp < ngroups && i < N) {
      const long b = grp / K;
      const int k = (int)(grp % K);
      r = (b * N + i) * K + k;
      ex = __ldg(a.obs + (size_t)(b * N + i) * a.Tp * 3) != 0.f;
      x = __ldg(a.pos + r * a.pos_stride);
      y = __ldg(a.pos + r * a.pos_stride + 1);
    }
    rowmap[tid] = r;
    px[tid] = x;
    py[tid] = y;
    const uint32_t em = __ballot_sync(0xffffffffu, ex);
    if (lane == 0) exist[warp] = em;
  }
  if (tid == 0) {
    for (int s = 0; s < nb; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sfull[s], NPW);
      mbar_init(&afull[s], NPW);
      mbar_init(&aempty[s], 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&pissued[s], 1);
    mbar_init(pfull, 1);
    mbar_init(pempty, NPW);
    mbar_init(tfull, 1);
    mbar_init(lready, NPW);
    binmask[0] = binmask[1] = 0u;
    fence_barrier_init();
  }
  if (warp == NPW) tmem_alloc<TCOLS>(tslot);
  tc_fence_before();
  __syncthreads();

  if (warp == NPW + 1) {
    // ===================== weight loader
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.packed);
      mbar_wait_idle(lready, 0);                                 // which bins does the tile use at all?
      uint64_t rem = active_bins(binmask);
      for (int kb = 0; rem; rem &= rem - 1) {
        const int g = __ffsll((long long)rem) - 1;
        for (int c = 0; c < CPB; ++c, ++kb) {
          const int slot = kb % nb;
          mbar_wait_idle(&bempty[slot], ((kb / nb) & 1) ^ 1);
          if (kb >= nb && kb - nb < 160) TRACE(400 + kb - nb);    // fc block kb - nb (6 MMAs) is complete
          if ((dbg & 2) && kb >= nb) {                           // timing experiment: no weight traffic after the first ring fill
            mbar_arrive(&bfull[slot]);
            continue;
          }
          mbar_arrive_expect_tx(&bfull[slot], (uint32_t)L.slot_bytes);
          bulk_g2s_hint(ring + (size_t)slot * L.slot_bytes, src + (size_t)(g * CPB + c) * L.slot_bytes,
                        (uint32_t)L.slot_bytes, &bfull[slot], L2_EVICT_LAST);
        }
      }
    }
  } else if (warp == NPW) {
    // ===================== fc MMA issuer.  The whole warp runs the loop with warp-uniform operands and one elected lane
    // issues (tc.cuh: elect_one); every descriptor is a base built once plus a compile-time offset, so nothing but
    // the MMAs sits between two MMAs.  The pool MMAs have their own issuing warp: while one warp waits, commits and
    // prepares, the other one's MMAs keep the tensor pipe busy (the queue behind one issuer is only a few MMAs deep).
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
    constexpr uint32_t idesc = idesc_bf16(TM, H);
    constexpr uint32_t lbo_b = H * 16;
    const uint64_t d_ring = smem_desc(smem_u32(ring), lbo_b, 128);
    mbar_wait(lready, 0);
    const int nst = __popcll(active_bins(binmask));              // stages = bins that hold at least one pair of the tile
    // The issuing warp must stay lean: every instruction between two blocks of MMAs is time in which the queue of the
    // tensor pipe drains (ring slot and barrier phase are counted up, not divided out of a block counter: the division
    // by the run-time ring length cost ~300 cycles per block and made this warp, not the tensor pipe, the bound).
    uint32_t slot = 0, ph = 0, pi = 0, pph = 0;
    uint64_t dsl = d_ring;
    const bool gate = !(dbg & 1024);                             // (1024: timing experiment, fc not held back)
    const uint32_t slot_adv = (uint32_t)L.slot_bytes >> 4;
    for (int g = 0; g < nst; ++g) {                              // (g counts stages here; the weights follow the bin list)
      const int as = g & 1;
      mbar_wait(&afull[as], (g >> 1) & 1);
      // The queue of the tensor pipe is first-in first-out and a dozen MMAs deep.  pool(g+1) is on the critical chain
      // (P is single-buffered: pool -> finishers' tcgen05.ld -> pool), fc(g) is not: fc(g) queues up behind pool(g+1)
      // and runs while P(g+1) is handed over, instead of sitting in front of it.
      // (three barriers in turn: the pool issuer can be two announcements ahead of this wait, not three — the third needs
      // aempty of this very stage)
      if (gate && g + 1 < nst) mbar_wait(&pissued[pi], pph);
      if (++pi == 3) {
        pi = 0;
        pph ^= 1;
      }
      tc_fence_after();
      if (lane == 0) TRACE(16 + 8 * g + 5);
      const uint32_t a_hi = tmem + 2 * H + as * H, a_lo = a_hi + H / 2;
      const uint32_t acc0 = g > 0;
#pragma unroll
      for (int c = 0; c < CPB; ++c) {
        mbar_wait(&bfull[slot], ph);
        tc_fence_after();
        if (elect_one()) {
          if (!(dbg & 4)) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = 2 * c + jj;                          // 16-wide K step inside the bin
              const uint64_t bhi = desc_adv(dsl, jj * 2 * lbo_b), blo = desc_adv(dsl, b_blk + jj * 2 * lbo_b);
              mma_bf16_ts(tmem, a_hi + 8 * j, bhi, idesc, j == 0 ? acc0 : 1u);
              if (P3) {
                mma_bf16_ts(tmem, a_lo + 8 * j, bhi, idesc, 1);
                mma_bf16_ts(tmem, a_hi + 8 * j, blo, idesc, 1);
              }
            }
          }
          mma_commit(&bempty[slot]);
          if (c == CPB - 1) mma_commit(&aempty[as]);
        }
        dsl += slot_adv;
        if (++slot == (uint32_t)nb) {
          slot = 0;
          ph ^= 1;
          dsl = d_ring;
        }
      }
      if (lane == 0) TRACE(16 + 8 * g + 6);
    }
    if (elect_one()) mma_commit(tfull);
    __syncwarp();
  } else if (warp == NPW + 2) {
    // ===================== pool MMA issuer: P = S_g @ h (hi, then lo), both operands in shared memory
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
    constexpr uint32_t idesc = idesc_bf16(TM, H);
    constexpr uint32_t lbo_b = H * 16, lbo_s = TM * 16;
    const uint32_t t_p = tmem + H;
    const uint64_t d_s0 = smem_desc(smem_u32(sm), lbo_s, 128);
    const uint64_t d_hh = smem_desc(smem_u32(ht), lbo_b, 128), d_hl = desc_adv(d_hh, H * TM * 2);
    mbar_wait(lready, 0);
    const int nst = __popcll(active_bins(binmask));
    uint32_t pi = 0;
    for (int g = 0; g < nst; ++g) {                              // stages
      const int sb = g & 1;
      const uint64_t ds = desc_adv(d_s0, sb * (TM * TM * 2));
      mbar_wait(&sfull[sb], (g >> 1) & 1);
      if (g > 0) mbar_wait(pempty, (g - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        if (!(dbg & 12)) {                                         // (timing experiments: 4 = no MMAs, 8 = no pool MMAs,
#pragma unroll                                                   //  16 = pool hi pass only)
          for (int j = 0; j < TM / 16; ++j) {                    // K = the 128 tile rows as neighbours
            mma_bf16(t_p, desc_adv(ds, j * 2 * lbo_s), desc_adv(d_hh, j * 2 * lbo_b), idesc, j > 0);
            if (P3 && !(dbg & 16)) mma_bf16(t_p, desc_adv(ds, j * 2 * lbo_s), desc_adv(d_hl, j * 2 * lbo_b), idesc, 1);
          }
        }
        mma_commit(pfull);                                       // (also frees the selection matrix: see the builders)
        if (g > 0) mbar_arrive(&pissued[pi]);
      }
      if (g > 0 && ++pi == 3) pi = 0;
    }
    __syncwarp();
  } else {
    // ===================== prologue (producer warps): tables, zeroed selection matrices, transposed hidden vectors
    if (tid == 0) TRACE(5);
    if (tid < 8) tab[tid] = tid <= a.n_rad ? __ldg(a.r2_edges + tid) : __int_as_float(0x7f800000);   // +inf: never reached
    if (tid >= 32 && tid < 48) tab[8 + tid - 32] = tid - 32 < 2 * a.n_ang ? __ldg(a.dirs + tid - 32) : 0.f;
    {
      uint4* z = reinterpret_cast<uint4*>(sm);
      for (int e = tid; e < 2 * TM * TM * 2 / 16; e += PT) z[e] = make_uint4(0u, 0u, 0u, 0u);
    }
    // h^T as a K-major operand: byte(c, j) = (j/8) * H*16 + c*16 + (j%8)*2.  A thread owns column c and eight neighbours
    // j (one 16-byte chunk of the hi and of the lo image); a warp reads 32 consecutive columns of a row (coalesced).
    // All loads are issued first and converted after the binning below, which hides their latency.
    if (tid == 0) TRACE(6);
    constexpr int HT_ITEMS = H * (TM / 8) / PT;
    static_assert(H * (TM / 8) % PT == 0, "h^T items per thread");
    float hv[HT_ITEMS][8];
#pragma unroll
    for (int it = 0; it < HT_ITEMS; ++it) {
      const int item = it * PT + tid;
      const int c = item % H, oct = item / H;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const long r = rowmap[oct * 8 + e];
        hv[it][e] = (r >= 0 && !(dbg & 64)) ? __ldg(a.h + r * (long)a.ld_h + c) : 0.f;   // (64: timing experiment)
      }
    }
    {
      uint4* b4 = reinterpret_cast<uint4*>(bins);                // 255 = no bin
      for (int e = tid; e < (int)(TM * L.bin_stride / 16); e += PT) b4[e] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    if (tid == 0) TRACE(7);
    asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");          // the tables and the bin fill are complete
    if (tid == 0) TRACE(8);
    // bins: one evaluation per UNORDERED pair {i, j} of a group gives bins[i][j] and bins[j][i].  Row i takes the pairs
    // {i, (i + k) mod N}, k = 1 .. (N-1)/2 (and k = N/2 for the first half of the rows when N is even): every pair once,
    // the same number of pairs per row; thread (row = tid % 128, quarter = tid / 128) takes k = 1 + quarter, 5 + quarter, ..
    // Entries no pair writes (self, padding, missing agents, rows of no group) keep the 255 of the fill above.
    {
      const int rl = tid & (TM - 1), q = tid >> 7;
      const int gbase = (rl / Npad) * Npad, i = rl % Npad;
      const bool valid = rowmap[rl] >= 0;                        // => every row i' < N of this group is valid
      const float xi = px[rl], yi = py[rl];
      const bool ex_i = (exist[rl >> 5] >> (rl & 31)) & 1u;
      float re[8], dr[16];
#pragma unroll
      for (int e = 0; e < 8; e += 4) *reinterpret_cast<float4*>(re + e) = *reinterpret_cast<const float4*>(tab + e);
#pragma unroll
      for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(dr + e) = *reinterpret_cast<const float4*>(tab + 8 + e);
      // (opaque to the compiler from here on: otherwise it re-reads the tables from shared memory in every iteration)
#pragma unroll
      for (int e = 0; e < 8; ++e) asm volatile("" : "+f"(re[e]));
#pragma unroll
      for (int e = 0; e < 16; ++e) asm volatile("" : "+f"(dr[e]));
      uint8_t* bgrp = bins + (size_t)gbase * L.bin_stride;       // rows of this group
      const int kmax = valid ? (N - 1) / 2 + ((N % 2 == 0 && i < N / 2) ? 1 : 0) : 0;
      uint32_t m0 = 0u, m1 = 0u;                                 // bins seen by this thread
      for (int k = 1 + q; k <= kmax; k += 4) {
        int j = i + k;
        if (j >= N) j -= N;
        const bool ex_j = (exist[(gbase + j) >> 5] >> ((gbase + j) & 31)) & 1u;
        int gf, gb;
        logpolar_bin_pair(px[gbase + j] - xi, py[gbase + j] - yi, re, dr, a.n_rad, a.n_ang, gf, gb);
        // a masked row still pools its existing neighbours: only the NEIGHBOUR has to exist
        if (!ex_j) gf = -1;
        if (!ex_i) gb = -1;
        bgrp[(size_t)i * L.bin_stride + j] = (uint8_t)gf;
        bgrp[(size_t)j * L.bin_stride + i] = (uint8_t)gb;
        if (gf >= 0) {
          if (gf < 32) m0 |= 1u << gf; else m1 |= 1u << (gf - 32);
        }
        if (gb >= 0) {
          if (gb < 32) m0 |= 1u << gb; else m1 |= 1u << (gb - 32);
        }
      }
      m0 = __reduce_or_sync(0xffffffffu, m0);
      m1 = __reduce_or_sync(0xffffffffu, m1);
      if (lane == 0) {
        if (m0) atomicOr(&binmask[0], m0);
        if (m1) atomicOr(&binmask[1], m1);
      }
    }
    if (tid == 0) TRACE(9);
#pragma unroll
    for (int it = 0; it < HT_ITEMS; ++it) {
      const int item = it * PT + tid;
      const int c = item % H, oct = item / H;
      const Split8 s8 = split8(hv[it]);
      const size_t o = (size_t)oct * H * 16 + (size_t)c * 16;
      *reinterpret_cast<uint4*>(ht + o) = s8.hi;
      *reinterpret_cast<uint4*>(ht + (size_t)H * TM * 2 + o) = s8.lo;
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");
    if (lane == 0) mbar_arrive(lready);                          // the other warps may read the bin mask now
    tc_fence_after();
    const uint32_t tmem = *tslot;
    if (tid == 0) TRACE(1);
    // Bins without a single pair in this tile are skipped altogether (their A operand would be zero): the stage list
    // is the set bits of the mask, in ascending bin order, for every role alike.
    const uint64_t act = active_bins(binmask);
    const int nst = __popcll(act);

    // ---- S builder role: thread (row = tid % 128, part = tid / 128) owns JT consecutive neighbours of its row
    const int srow = tid & (TM - 1), part = tid >> 7;
    const int JT = Npad >= 32 ? Npad / 4 : 8;                    // neighbours per thread (8, 16 or 32)
    const bool s_on = part * JT < Npad;
    const uint8_t* sb_src = bins + (size_t)srow * L.bin_stride + part * JT;
    // K index of neighbour j of this row = its tile lane: group base + j
    const uint32_t s_dst = smem_u32(sm) + (uint32_t)(((srow / Npad) * Npad + part * JT) / 8) * (TM * 16) + srow * 16;
    auto build = [&](int g, int bin) {                           // selection matrix + partial counts of stage g = bin `bin`
      const uint32_t g4 = (uint32_t)bin * 0x01010101u;
      int cnt = 0;
      if (s_on && !(dbg & 32)) {                                 // (32: timing experiment, handshakes only)
        const uint32_t dst = s_dst + (uint32_t)(g & 1) * (TM * TM * 2);
        for (int o = 0; o < JT / 8; ++o) {
          const uint2 w = *reinterpret_cast<const uint2*>(sb_src + 8 * o);
          const uint32_t z0 = eq_bytes(w.x, g4), z1 = eq_bytes(w.y, g4);
          cnt += __popc(z0) + __popc(z1);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + o * (TM * 16)), "r"(ones_lo(z0)),
                       "r"(ones_hi(z0)), "r"(ones_lo(z1)), "r"(ones_hi(z1))
                       : "memory");
        }
      }
      pc[((g % 3) * 4 + part) * TM + srow] = cnt;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[g & 1]);
    };

    // ---- finisher role: thread = TMEM lane (row) 32*(warp%4) + lane, columns CW*(warp/4) .. +CW of P
    const int q4 = warp & 3, cg = warp >> 2;
    const int frow = 32 * q4 + lane;
    const uint32_t lane_f = (uint32_t)(32 * q4) << 16;
    const uint32_t t_p = tmem + lane_f + H + cg * CW;
    const uint32_t t_a = tmem + lane_f + 2 * H + cg * (CW / 2);

    // The selection matrix of stage g + 1 is built at the top of stage g: its buffer was read by pool(g-1), which the
    // pfull wait of stage g - 1 has shown to be complete.  (Building it two stages ahead, in the time the finishers wait
    // for the fc MMAs of stage g - 2, measured 2-4 % slower: the stores then coincide with the pool MMAs' operand reads.)
    uint64_t rem = act;
    build(0, __ffsll((long long)rem) - 1);
    rem &= rem - 1;
    for (int g = 0; g < nst; ++g) {                              // stages
      if (tid == 0 && g < 36) TRACE(16 + 8 * g);
      if (g + 1 < nst) {
        build(g + 1, __ffsll((long long)rem) - 1);
        rem &= rem - 1;
      }
      if (tid == 0) TRACE(16 + 8 * g + 1);
      mbar_wait(pfull, g & 1);                                   // pool(g) complete (=> every warp's build(g) is visible)
      tc_fence_after();
      if (tid == 0) TRACE(16 + 8 * g + 2);
      const int* pcg = pc + (size_t)(g % 3) * 4 * TM + frow;     // read before pempty: buffer g%3 is rewritten by build(g+3)
      const int cnt = pcg[0] + pcg[TM] + pcg[2 * TM] + pcg[3 * TM];
      float v[CW];
      if (!(dbg & 32)) {
        if constexpr (CW == 32) tmem_ld32(t_p, v); else tmem_ld16(t_p, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = 0.f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pempty);
      if (cnt > 1) {
        const float inv = __frcp_rn((float)cnt);                 // mean = sum * (1/count)
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] *= inv;
      }
      uint32_t hi[CW / 2], lo[CW / 2];
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
      if (tid == 0) TRACE(16 + 8 * g + 3);
      if (g >= 2) mbar_wait(&aempty[g & 1], ((g - 2) >> 1) & 1);             // fc(g-2) has read this stage
      tc_fence_after();
      const uint32_t ta = t_a + (g & 1) * H;
      if (dbg & 32) {
      } else if constexpr (CW == 32) {
        tmem_st16(ta, reinterpret_cast<const float*>(hi));
        tmem_st16(ta + H / 2, reinterpret_cast<const float*>(lo));
      } else {
        tmem_st8(ta, reinterpret_cast<const float*>(hi));
        tmem_st8(ta + H / 2, reinterpret_cast<const float*>(lo));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[g & 1]);
      if (tid == 0) TRACE(16 + 8 * g + 4);
    }

    // ===================== epilogue: warp w reads TMEM lanes 32*(w%4).., columns 32*(w/4).., adds the bias, applies the
    // ReLU and parks its block in shared memory (the operands of the pool MMAs are dead by now); then every warp
    // writes whole 4*H-byte rows (the thread-per-row stores of the first version took 8.7k cycles: 32 lines per
    // instruction)
    mbar_wait(tfull, 0);
    tc_fence_after();
    if (tid == 0) TRACE(3);
    constexpr int OLD = H + 4;                                   // floats per staged row
    float* ostage = reinterpret_cast<float*>(smem);              // [128][H + 4] <= the h^T + S region
    static_assert((size_t)TM * OLD * 4 <= 2 * (size_t)H * TM * 2 + 2 * (size_t)TM * TM * 2, "epilogue staging");
    {
      const int c0 = cg * 32;
      if (c0 < H) {
        float acc[32];
        tmem_ld32(tmem + lane_f + c0, acc);
        tmem_ld_wait();
        float* orow = ostage + (size_t)frow * OLD + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + j));
          float4 o;
          o.x = fmaxf(acc[j] + bv.x, 0.f);
          o.y = fmaxf(acc[j + 1] + bv.y, 0.f);
          o.z = fmaxf(acc[j + 2] + bv.z, 0.f);
          o.w = fmaxf(acc[j + 3] + bv.w, 0.f);
          *reinterpret_cast<float4*>(orow + j) = o;
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(PT) : "memory");
    if (!(dbg & 4)) {
      constexpr int LPR = H / 4;                                 // lanes per row (one float4 each)
      constexpr int RPI = 32 / LPR;                              // rows per warp instruction
      const int sub = lane / LPR, c4 = lane % LPR;
#pragma unroll
      for (int rr = 0; rr < 8; rr += RPI) {
        const int rl = warp * 8 + rr + sub;
        const long r = rowmap[rl];
        if (r >= 0)
          *reinterpret_cast<float4*>(a.out + r * (long)H + c4 * 4) =
              *reinterpret_cast<const float4*>(ostage + (size_t)rl * OLD + c4 * 4);
      }
    }
    if (tid == 0) TRACE(4);
  }
#undef TRACE
  tc_fence_before();
  __syncthreads();
  if (warp == NPW) tmem_dealloc(*tslot, TCOLS);
}

int npad_of(int N) {
  int p = 8;
  while (p < N) p <<= 1;
  return p;
}

}  // namespace

bool social_fc_ts_eligible(const SocialFcArgs& a) {
  const int G = a.n_rad * a.n_ang;
  if (gemm_mode() == 0 || !a.packed) return false;
  if ((a.H != 64 && a.H != 128) || a.ld_h % 4 != 0) return false;
  if (a.N > 128 || a.N < 1 || G > MAXG || G < 1) return false;
  if ((reinterpret_cast<uintptr_t>(a.h) & 15) || (reinterpret_cast<uintptr_t>(a.bias) & 15)) return false;
  if (a.n_rad > 7 || a.n_ang > 8) return false;
  const Layout L = make_layout(a.H, npad_of(a.N), a.n_rad, a.n_ang);
  return L.nb >= 3;
}

int social_fc_ts(const SocialFcArgs& a, cudaStream_t st) {
  const int Npad = npad_of(a.N), G = a.n_rad * a.n_ang;
  const Layout L = make_layout(a.H, Npad, a.n_rad, a.n_ang);
  const long ngroups = (long)a.B * a.K;
  if (ngroups == 0) return DESIRE_OK;
  const int gpt = TM / Npad;
  const unsigned grid = (unsigned)((ngroups + gpt - 1) / gpt);
  const int passes = gemm_mode() == 1 ? 1 : 3;
  int dbg = 0;                                                   // DESIRE_SOCIAL_DBG: timing experiments (results are wrong)
  if (const char* e = getenv("DESIRE_SOCIAL_DBG")) dbg = atoi(e);
  // DESIRE_SOCIAL_TRACE=1: block 0 records clock64() at its milestones; printed after the launch (timing tool only)
  static long long* trace = nullptr;
  static const bool want_trace = [] {
    const char* e = getenv("DESIRE_SOCIAL_TRACE");
    return e && e[0] == '1';
  }();
  if (want_trace && !trace) DESIRE_CUDA(cudaMalloc(&trace, 1024 * sizeof(long long)));
  if (want_trace) DESIRE_CUDA(cudaMemsetAsync(trace, 0, 1024 * sizeof(long long), st));
#define SOCIAL_TS_LAUNCH(HH, PP)                                                                              \
  do {                                                                                                        \
    DESIRE_ENSURE_SMEM((social_fc_ts_kernel<HH, PP>), L.total);                                               \
    DESIRE_LAUNCH(st, (social_fc_ts_kernel<HH, PP><<<grid, NTHR, L.total, st>>>(a, Npad, dbg, trace)));       \
  } while (0)
  if (a.H == 128) {
    if (passes == 3) SOCIAL_TS_LAUNCH(128, true); else SOCIAL_TS_LAUNCH(128, false);
  } else {
    if (passes == 3) SOCIAL_TS_LAUNCH(64, true); else SOCIAL_TS_LAUNCH(64, false);
  }
#undef SOCIAL_TS_LAUNCH
  if (want_trace) {
    static int printed = 0;
    long long h[1024];
    DESIRE_CUDA(cudaStreamSynchronize(st));
    DESIRE_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ == 4) {
      const long long t0 = h[0];
      fprintf(stderr, "social trace (cycles from start): sync %lld zeroed %lld loads issued %lld tables barrier %lld binned %lld prologue %lld tfull %lld end %lld\n", h[5] - t0, h[6] - t0, h[7] - t0, h[8] - t0, h[9] - t0, h[1] - t0, h[3] - t0, h[4] - t0);
      for (int g = 0; g < G; ++g) {
        const long long* e = h + 16 + 8 * g;
        fprintf(stderr, "  stage %2d: start %7lld build-next %6lld wait-pool %5lld ld+convert %6lld wait+st+arrive %5lld | mma: afull at %7lld fc issue %5lld\n",
                g, e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - t0, e[6] - e[5]);
      }
      fprintf(stderr, "  completion of the fc blocks (6 MMAs each; cycles from start), four per stage:\n");
      for (int g = 0; g < G; ++g)
        fprintf(stderr, "  stage %2d: %7lld %7lld %7lld %7lld | pool complete (seen by warp 0) %7lld\n", g, h[400 + 4 * g] - t0,
                h[401 + 4 * g] - t0, h[402 + 4 * g] - t0, h[403 + 4 * g] - t0, h[16 + 8 * g + 2] - t0);
    }
  }
  return DESIRE_OK;
}

}  // namespace desire
