// Persistent tcgen05 GRU recurrence, third design (TF-1.x GRUCell, reference model/model.py:137-148,279-285):
//
//   r = sigmoid(xp_r + [ex,h] @ Wr)   u = sigmoid(xp_u + [ex,h] @ Wu)   c = tanh(xp_c + [ex, r*h] @ Wc)   h' = u*h + (1-u)*c
//
// What the ncu capture of the second design showed (profiles/r1i_ncu_summaries.txt: tensor pipe 12-19 %, 69 % of the
// warp stalls on the long scoreboard): the epilogue threads own one row each (that is how tcgen05.ld hands out the
// accumulator), so every global load/store of theirs touched 32 different 128-byte lines — 30 sectors per request —
// and with the shared-memory carve-out at its maximum there is no L1 left to catch the re-use; h_{t-1} was re-read
// from global memory twice per step; the four phases of a step ran strictly one after the other.  This design:
//
//   * state in registers: a thread keeps the FP32 h of its (row, HC columns) for the whole kernel;
//   * every per-row input arrives by TMA: 2-D tensor-map copies (cp.async.bulk.tensor, SASS UTMALDG) of
//     [128 rows x 32 columns] FP32 boxes with the 128-byte swizzle into a small ring, issued by a loader thread that
//     runs ahead of the epilogues — the hoisted input projection xp (three boxes per 32 columns and step) and, in the
//     Decoder-2 form (EX), the extra operand and the initial state; a thread reads its row's 128 bytes as eight
//     conflict-free 16-byte loads.  A box is announced on the full-barrier of the COLUMN GROUP that consumes it
//     (each group sees every phase of its own barrier in order; a per-slot barrier shared by groups that alternate on
//     the slot cannot be waited on by parity), and released on the empty-barrier of its ring slot;
//   * a step is three MMA groups — r columns, u columns, candidate — and three epilogue phases that each run in the
//     shadow of the next MMA group:
//         MMA     | r(t)        | u(t)              | cand(t)             | (idle)        | r(t+1) ...
//         SIMT    |  E2b(t-1)   | E1: r, r*h -> A   | E2a: u -> TMEM      | E2b: c, h'    |
//     E1 overwrites the A operand chunk by chunk behind the u-group (a_free[kc] is committed after the u-group's MMAs
//     on K chunk kc), E2a parks sigmoid(u) in the u columns of TMEM, and only E2b (tanh + blend) is exposed;
//   * when the candidate has its own TMEM columns (3H <= 512) the next step's r-group starts on the K chunks of h'
//     as they are produced (h_ready[kc]); with H = 256 the gates fill all 512 columns, the candidate re-uses the r
//     columns and the r-group waits for the whole tile;
//   * h' leaves through TMA as well: E2b overwrites its own row of the xp_c box it has just read with the FP32 h'
//     (same thread, same bytes), and a store thread sends the [128 x 32] box to HBM with cp.async.bulk.tensor
//     (SASS UTMASTG) before it releases the ring slot — 128-byte rows instead of 32 scattered 16-byte sectors per
//     store instruction (the direct stores were 15 % of the kernel's stall samples as MIO throttle, in the one
//     phase that is not hidden behind an MMA group);
//   * recurrent weights: packed BF16 hi/lo images (tc_pack_b, n-tile = H) streamed through a ring of 16 KB slots by
//     1-D bulk copies with an evict-last L2 hint.
//
// 3xBF16 (A_hi B_hi + A_lo B_hi + A_hi B_lo, FP32 accumulation in TMEM) as everywhere else in the library.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {

using namespace tc;

constexpr int TM = 128;
constexpr int XBOX_BYTES = TM * 32 * 4;   // one box: 128 rows x 32 FP32 columns, 128-byte rows
constexpr int EPI_WARPS = 16;             // 4 TMEM lane quadrants x 4 column groups
constexpr int NTHR3 = (EPI_WARPS + 4) * 32;   // + one warpgroup: MMA issuer, weight loader, box loader, store thread

template <int H, bool EX>
struct Cfg3 {
  static constexpr int HC = H / 4;                     // columns per epilogue thread
  static constexpr int NB = HC / 32;                   // 32-column chunks (= boxes per gate) per thread
  static constexpr int NKC = H / 32;                   // K chunks of the state part of the A operand
  static constexpr int NKA = EX ? 2 * NKC : NKC;       // K chunks of the whole A operand: [ex | h]
  static constexpr int KOFF = EX ? NKC : 0;            // first chunk of the state part
  static constexpr int KSLOT = (H >= 256) ? 16 : 32;   // K extent of one weight slot
  static constexpr int SPC = 32 / KSLOT;               // slots per K chunk
  static constexpr int SLOT_HALF = (KSLOT / 8) * H * 16;
  static constexpr int SLOT_BYTES = 2 * SLOT_HALF;     // hi + lo: 16 KB for both sizes
  static constexpr int BLOCK_BYTES = 2 * 4 * H * 16;   // one packed (n-tile, 32-wide K) block of tc_pack_b
  static constexpr int NSW = (H >= 256 || EX) ? 4 : 6; // weight ring slots
  static constexpr int NXB = (H >= 256 || EX) ? 2 : 4; // box ring slots (<= 4: one box in flight per column group)
  static constexpr int A_HALF = NKA * 4 * 2048;        // [K/8 chunks][128 rows][16 B]
  static constexpr bool ALIAS = (3 * H > 512);         // candidate accumulates in the consumed r columns
  static constexpr int CAND_COL = ALIAS ? 0 : 2 * H;
  static constexpr int PRO = EX ? 2 * NKC : 0;         // prologue boxes: ex, then h0 (each in (chunk-in-thread, group) order)
  static constexpr int BPS = 3 * NKC;                  // boxes per step: xp_r | xp_u | xp_c
  static constexpr int WPS = 3 * NKA * SPC;            // weight slots per step
  // columns per tcgen05.ld / wait in the three epilogue phases: as many as the registers allow
  // (16 independent activation chains per wait instead of 8: the phases are latency-bound, 4 warps per scheduler)
  static constexpr int GR1 = 16;                        // E1, E2a (32 does not fit the 112-register budget next to the state)
  static constexpr int GR2 = (HC <= 32) ? 16 : 8;      // E2b holds two accumulator granules next to the state
  static constexpr int NBAR = 2 * NSW + 4 + NXB + 4 + NKA + NKC + 4;
  static constexpr size_t SMEM = 1024 + 2 * (size_t)A_HALF + (size_t)NSW * SLOT_BYTES + (size_t)NXB * XBOX_BYTES + NBAR * 8 + 16;
  static_assert(NXB <= 4, "a column group may have one box in flight");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct Gru3Args {
  int R, T;
  int xp_step;          // column offset of step t inside an xp row = t * xp_step
  const float* h0;      // non-EX: read directly (rows shared by h0_div samples); EX: through tm_h0
  int h0_div, ld_h0;
  int has_hs, hs_step;  // all states through tm_hs: step t at column t * hs_step
  int has_hf;           // final state through tm_hf
  const uint8_t* wg;    // packed gates, n-tile H: [2][K/32] blocks (r columns, then u columns)
  const uint8_t* wc;    // packed candidate: [K/32] blocks
  int passes;
  int xp_const;         // 1: the same xp every step (Decoder-1) -> keep it in L2
  int dbg;              // timing experiments only (DESIRE_GRU3_DBG): 1 = no box copies, 2 = no weight copies (results are wrong)
};

__device__ __forceinline__ float4 lds128(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float* v) {
  if (N == 32) tmem_ld32(taddr, v);
  else if (N == 16) tmem_ld16(taddr, v);
  else tmem_ld8(taddr, v);
}
template <int N>
__device__ __forceinline__ void tmem_stn(uint32_t taddr, const float* v) {
#pragma unroll
  for (int i = 0; i < N; i += 8) tmem_st8(taddr + i, v + i);
}

// WD: watchdog on every mbarrier wait (DESIRE_GRU3_WATCHDOG=1; the tests set it) — a protocol error traps with the
// name of the barrier instead of hanging the device.
template <int H, bool EX, bool WD>
__global__ void __launch_bounds__(NTHR3, 1)
    gru_tc3_kernel(const __grid_constant__ CUtensorMap tm_xp, const __grid_constant__ CUtensorMap tm_ex,
                   const __grid_constant__ CUtensorMap tm_h0, const __grid_constant__ CUtensorMap tm_hs,
                   const __grid_constant__ CUtensorMap tm_hf, Gru3Args a) {
  using C = Cfg3<H, EX>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // the swizzled boxes need 1024-byte alignment
  uint8_t* xring = smem;                                             // [NXB] boxes, each 1024-aligned
  uint8_t* a_hi = xring + (size_t)C::NXB * XBOX_BYTES;
  uint8_t* a_lo = a_hi + C::A_HALF;
  uint8_t* wring = a_lo + C::A_HALF;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(wring + (size_t)C::NSW * C::SLOT_BYTES);
  uint64_t* wempty = wfull + C::NSW;
  uint64_t* xfull = wempty + C::NSW;      // [4]   one per column group: "your next box has landed"
  uint64_t* xempty = xfull + 4;           // [NXB] one per ring slot
  uint64_t* g_r_done = xempty + C::NXB;
  uint64_t* g_u_done = g_r_done + 1;
  uint64_t* c_done = g_u_done + 1;
  uint64_t* rh_ready = c_done + 1;
  uint64_t* h_ready = rh_ready + 1;       // [NKA]  A chunk kc holds its operand for the gate groups
  uint64_t* a_free = h_ready + C::NKA;    // [NKC]  the u-group has read state chunk kc
  uint64_t* hdone = a_free + C::NKC;      // [4]    column group: h' of its current box is in the ring slot
  uint32_t* tslot = reinterpret_cast<uint32_t*>(hdone + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long row0 = (long)blockIdx.x * TM;

  if (tid == 0) {
    for (int s = 0; s < C::NSW; ++s) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&xfull[s], 1);
    for (int s = 0; s < C::NXB; ++s) mbar_init(&xempty[s], 4);   // the four quadrant warps of the group that read the box
    mbar_init(g_r_done, 1);
    mbar_init(g_u_done, 1);
    mbar_init(c_done, 1);
    mbar_init(rh_ready, EPI_WARPS);
    for (int k = 0; k < C::NKA; ++k) mbar_init(&h_ready[k], 4);
    for (int k = 0; k < C::NKC; ++k) mbar_init(&a_free[k], 1);
    for (int k = 0; k < 4; ++k) mbar_init(&hdone[k], 4);
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc<512>(tslot);
  if (warp == EPI_WARPS + 3 && lane == 0) {
    tma_prefetch_desc(&tm_hs);
    tma_prefetch_desc(&tm_hf);
  }
  if (warp == EPI_WARPS + 2 && lane == 0) {
    tma_prefetch_desc(&tm_xp);
    if (EX) {
      tma_prefetch_desc(&tm_ex);
      tma_prefetch_desc(&tm_h0);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  // 640 threads launch with 96 registers each; the last warpgroup (two single-thread roles) gives most of its share
  // back so the epilogue threads can hold 64 FP32 state values next to an 8-column working set.  (setmaxnreg is
  // .aligned: every warp of a warpgroup must execute the SAME instruction, so the release sits before the role split.)
  if (warp >= EPI_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
  if (warp < EPI_WARPS) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ======================================================================== epilogue warps
    const int q = warp & 3, cs = warp >> 2;          // TMEM lane quadrant (must equal warp % 4), column group
    const int rloc = q * 32 + lane;                  // TMEM lane == row of the tile
    const long row = row0 + rloc;
    const bool ok = row < a.R;
    const int cbeg = cs * C::HC;                     // this thread's columns [cbeg, cbeg + HC)
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* my_hi = a_hi + rloc * 16;
    uint8_t* my_lo = a_lo + rloc * 16;
    uint64_t* my_full = &xfull[cs];
    float h[C::HC];
    uint32_t xk = 0;                                 // boxes this column group has consumed (phase of my_full)
    uint32_t xg = cs;                                // ring position of this group's next box: they are 4 apart

    // Wait for this group's next box; returns its ring slot.  The loader issues boxes in ring order and a group's
    // boxes are every fourth one, so the slot sequence is a function of the running position alone.
    auto next_box = [&](int tag) -> int {
      mbar_wait_tag<WD>(my_full, xk & 1, tag);
      const int slot = xg % C::NXB;
      ++xk;
      xg += 4;
      return slot;
    };

    // ---- prologue: extra operand and initial state -> registers + A operand
    if (EX) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {         // 0: ex -> chunks [0,NKC), 1: h0 -> registers + chunks [NKC,2NKC)
#pragma unroll
        for (int j = 0; j < C::NB; ++j) {
          const int kc = part * C::NKC + cs * C::NB + j;
          const int slot = next_box(70 + part);
          const uint8_t* box = xring + (size_t)slot * XBOX_BYTES;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const float4 x0 = lds128(box + swz128(rloc, g8 * 2)), x1 = lds128(box + swz128(rloc, g8 * 2 + 1));
            float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            if (part == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) h[j * 32 + g8 * 8 + i] = v[i];
            }
            const Split8 s = split8(v);
            *reinterpret_cast<uint4*>(my_hi + (kc * 4 + g8) * 2048) = s.hi;
            *reinterpret_cast<uint4*>(my_lo + (kc * 4 + g8) * 2048) = s.lo;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&xempty[slot]);
            mbar_arrive(&h_ready[kc]);
          }
        }
      }
    } else {
      const float* h0r = (a.h0 && ok) ? a.h0 + (row / a.h0_div) * (long)a.ld_h0 + cbeg : nullptr;
#pragma unroll
      for (int c = 0; c < C::HC; c += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h0r) v = __ldg(reinterpret_cast<const float4*>(h0r + c));
        h[c] = v.x; h[c + 1] = v.y; h[c + 2] = v.z; h[c + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < C::NB; ++j) {
        const int kc = cs * C::NB + j;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const Split8 s = split8(&h[j * 32 + g8 * 8]);
          *reinterpret_cast<uint4*>(my_hi + (kc * 4 + g8) * 2048) = s.hi;
          *reinterpret_cast<uint4*>(my_lo + (kc * 4 + g8) * 2048) = s.lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[kc]);
      }
    }

    for (int t = 0; t < a.T; ++t) {
      const uint32_t par = t & 1;
      // ---------------- E1 (behind the u-group): r = sigmoid(.), r*h replaces h in the A operand
      mbar_wait_tag<WD>(g_r_done, par, 10);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < C::NB; ++j) {
        const int kc = cs * C::NB + j;               // chunk of the state (0-based inside the state part)
        const int slot = next_box(60);
        const uint8_t* box = xring + (size_t)slot * XBOX_BYTES;
#pragma unroll
        for (int gi = 0; gi < 32 / C::GR1; ++gi) {
          float acc[C::GR1];
          tmem_ldn<C::GR1>(trow + kc * 32 + gi * C::GR1, acc);
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < C::GR1 / 8; ++c8) {      // 8 columns = one 16-byte chunk of the A operand
            const int g8 = gi * (C::GR1 / 8) + c8;
            const float4 x0 = lds128(box + swz128(rloc, g8 * 2)), x1 = lds128(box + swz128(rloc, g8 * 2 + 1));
            const int hc = j * 32 + g8 * 8;
            float* v = acc + c8 * 8;
            v[0] = sigmoid_a(v[0] + x0.x) * h[hc + 0];
            v[1] = sigmoid_a(v[1] + x0.y) * h[hc + 1];
            v[2] = sigmoid_a(v[2] + x0.z) * h[hc + 2];
            v[3] = sigmoid_a(v[3] + x0.w) * h[hc + 3];
            v[4] = sigmoid_a(v[4] + x1.x) * h[hc + 4];
            v[5] = sigmoid_a(v[5] + x1.y) * h[hc + 5];
            v[6] = sigmoid_a(v[6] + x1.z) * h[hc + 6];
            v[7] = sigmoid_a(v[7] + x1.w) * h[hc + 7];
          }
          if (gi == 0) mbar_wait_tag<WD>(&a_free[kc], par, 20 + kc);     // the u-group's MMAs have read h chunk kc
#pragma unroll
          for (int c8 = 0; c8 < C::GR1 / 8; ++c8) {
            const int g8 = gi * (C::GR1 / 8) + c8;
            const Split8 sp = split8(acc + c8 * 8);
            *reinterpret_cast<uint4*>(my_hi + ((C::KOFF + kc) * 4 + g8) * 2048) = sp.hi;
            *reinterpret_cast<uint4*>(my_lo + ((C::KOFF + kc) * 4 + g8) * 2048) = sp.lo;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&xempty[slot]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(rh_ready);

      // ---------------- E2a (behind the candidate group): u = sigmoid(.) parked in its TMEM columns
      mbar_wait_tag<WD>(g_u_done, par, 11);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < C::NB; ++j) {
        const int kc = cs * C::NB + j;
        const int slot = next_box(62);
        const uint8_t* box = xring + (size_t)slot * XBOX_BYTES;
#pragma unroll
        for (int gi = 0; gi < 32 / C::GR1; ++gi) {
          float acc[C::GR1];
          tmem_ldn<C::GR1>(trow + H + kc * 32 + gi * C::GR1, acc);
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < C::GR1 / 8; ++c8) {
            const int g8 = gi * (C::GR1 / 8) + c8;
            const float4 x0 = lds128(box + swz128(rloc, g8 * 2)), x1 = lds128(box + swz128(rloc, g8 * 2 + 1));
            float* v = acc + c8 * 8;
            v[0] = sigmoid_a(v[0] + x0.x);
            v[1] = sigmoid_a(v[1] + x0.y);
            v[2] = sigmoid_a(v[2] + x0.z);
            v[3] = sigmoid_a(v[3] + x0.w);
            v[4] = sigmoid_a(v[4] + x1.x);
            v[5] = sigmoid_a(v[5] + x1.y);
            v[6] = sigmoid_a(v[6] + x1.z);
            v[7] = sigmoid_a(v[7] + x1.w);
          }
          tmem_stn<C::GR1>(trow + H + kc * 32 + gi * C::GR1, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&xempty[slot]);
      }
      tmem_st_wait();

      // ---------------- E2b (exposed): candidate, state update, h' -> registers, HBM and the A operand
      mbar_wait_tag<WD>(c_done, par, 12);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < C::NB; ++j) {
        const int kc = cs * C::NB + j;
        const int slot = next_box(64);
        uint8_t* box = xring + (size_t)slot * XBOX_BYTES;
#pragma unroll
        for (int gi = 0; gi < 32 / C::GR2; ++gi) {
          float accc[C::GR2], u[C::GR2];
          tmem_ldn<C::GR2>(trow + C::CAND_COL + kc * 32 + gi * C::GR2, accc);
          tmem_ldn<C::GR2>(trow + H + kc * 32 + gi * C::GR2, u);
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < C::GR2 / 8; ++c8) {
            const int g8 = gi * (C::GR2 / 8) + c8;
            const float4 x0 = lds128(box + swz128(rloc, g8 * 2)), x1 = lds128(box + swz128(rloc, g8 * 2 + 1));
            const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const int hc = j * 32 + g8 * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float cd = tanh_a(accc[c8 * 8 + i] + xs[i]);
              h[hc + i] = fmaf(u[c8 * 8 + i], h[hc + i] - cd, cd);     // u*h + (1-u)*cd
            }
            // h' replaces this thread's row of the xp_c box (read above): the store thread sends the box to HBM
            *reinterpret_cast<float4*>(box + swz128(rloc, g8 * 2)) = make_float4(h[hc], h[hc + 1], h[hc + 2], h[hc + 3]);
            *reinterpret_cast<float4*>(box + swz128(rloc, g8 * 2 + 1)) = make_float4(h[hc + 4], h[hc + 5], h[hc + 6], h[hc + 7]);
            const Split8 sp = split8(&h[hc]);
            *reinterpret_cast<uint4*>(my_hi + ((C::KOFF + kc) * 4 + g8) * 2048) = sp.hi;
            *reinterpret_cast<uint4*>(my_lo + ((C::KOFF + kc) * 4 + g8) * 2048) = sp.lo;
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&hdone[cs]);                  // the store thread releases the slot once the box has left
          mbar_arrive(&h_ready[C::KOFF + kc]);
        }
      }
    }
  } else if (warp == EPI_WARPS) {
    // ======================================================================== MMA issuer
    // The whole warp runs the loops with warp-uniform operands and one elected lane issues (tc.cuh: elect_one): under
    // `if (lane == 0)` the compiler wrapped every UTCHMMA in a lane-broadcast loop and an MMA left only every ~93 cycles
    // (tools/mma_rate.py), whatever its shape.  Descriptors are bases built once plus offsets.
    {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      constexpr uint32_t idesc = idesc_bf16(TM, H);
      constexpr uint32_t lbo = H * 16;
      const uint64_t d_ahi = smem_desc(smem_u32(a_hi), 2048, 128), d_alo = smem_desc(smem_u32(a_lo), 2048, 128);
      const uint64_t d_w = smem_desc(smem_u32(wring), lbo, 128);
      const bool p3 = a.passes == 3;
      uint32_t itw = 0;
      for (int t = 0; t < a.T; ++t) {
        const uint32_t par = t & 1;
#pragma unroll 1
        for (int nb = 0; nb < 3; ++nb) {          // r columns, u columns, candidate
          const uint32_t d = tm + (nb == 0 ? 0 : (nb == 1 ? H : C::CAND_COL));
          if (nb == 2) {
            mbar_wait_tag<WD>(rh_ready, par, 30);
            tc_fence_after();
          }
          uint32_t accf = 0;
#pragma unroll 1
          for (int kc = 0; kc < C::NKA; ++kc) {
            if (nb == 0) {
              // the extra-operand chunks are written once (phase 0); state chunk phases advance every step
              if (C::ALIAS) {                      // the r columns still hold the candidate the epilogues are reading
                if (kc == 0)
                  for (int k = 0; k < C::NKA; ++k) mbar_wait_tag<WD>(&h_ready[k], k < C::KOFF ? 0u : par, 40 + k);
              } else {
                mbar_wait_tag<WD>(&h_ready[kc], kc < C::KOFF ? 0u : par, 40 + kc);
              }
              tc_fence_after();
            }
#pragma unroll
            for (int s = 0; s < C::SPC; ++s, ++itw) {
              const int slot = itw % C::NSW;
              const uint64_t dsl = desc_adv(d_w, slot * C::SLOT_BYTES);
              const uint32_t ao = (kc * 4 + s * (C::KSLOT / 16) * 2) * 2048;
              const uint64_t dah = desc_adv(d_ahi, ao), dal = desc_adv(d_alo, ao);
              mbar_wait_tag<WD>(&wfull[slot], (itw / C::NSW) & 1, 50 + slot);
              tc_fence_after();
              if (elect_one()) {
                if (p3) {
#pragma unroll
                  for (int j = 0; j < C::KSLOT / 16; ++j) {
                    const uint64_t bhi = desc_adv(dsl, j * 2 * lbo), blo = desc_adv(dsl, C::SLOT_HALF + j * 2 * lbo);
                    mma_bf16(d, desc_adv(dah, j * 2 * 2048), bhi, idesc, j == 0 ? accf : 1u);
                    mma_bf16(d, desc_adv(dal, j * 2 * 2048), bhi, idesc, 1);
                    mma_bf16(d, desc_adv(dah, j * 2 * 2048), blo, idesc, 1);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < C::KSLOT / 16; ++j)
                    mma_bf16(d, desc_adv(dah, j * 2 * 2048), desc_adv(dsl, j * 2 * lbo), idesc, j == 0 ? accf : 1u);
                }
                mma_commit(&wempty[slot]);
                if (s == C::SPC - 1 && nb == 1 && kc >= C::KOFF) mma_commit(&a_free[kc - C::KOFF]);  // state chunk read by both gate groups
                if (s == C::SPC - 1 && kc == C::NKA - 1) mma_commit(nb == 0 ? g_r_done : (nb == 1 ? g_u_done : c_done));
              }
              accf = 1;
            }
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ======================================================================== loaders: warp 17 weights, warp 18 boxes
    if (warp == EPI_WARPS + 2 && lane == 0) {
      const uint32_t totX = C::PRO + (uint32_t)a.T * C::BPS;
      const uint64_t xpol = a.xp_const ? L2_EVICT_LAST : L2_EVICT_FIRST;
      for (uint32_t g = 0; g < totX; ++g) {
        const int slot = g % C::NXB;
        mbar_wait_tag<WD>(&xempty[slot], ((g / C::NXB) & 1) ^ 1, 80 + slot);
        // box g: prologue [ex x NKC | h0 x NKC] (EX only), then per step [xp_r | xp_u | xp_c] x NKC; inside a
        // block of NKC boxes, box w = j*4 + cs is chunk j of column group cs
        const void* tm = &tm_xp;
        int c0;
        const int w = ((int)g < C::PRO ? g : g - C::PRO) % C::NKC;
        const int col = (w & 3) * C::HC + (w >> 2) * 32;
        if ((int)g < C::PRO) {
          tm = (g < C::NKC) ? (const void*)&tm_ex : (const void*)&tm_h0;
          c0 = col;
        } else {
          const uint32_t t = (g - C::PRO) / C::BPS, ib = (g - C::PRO) % C::BPS;
          c0 = (int)(t * a.xp_step + (ib / C::NKC) * H + col);
        }
        uint64_t* fb = &xfull[w & 3];
        if (a.dbg & 1) {
          mbar_arrive(fb);
        } else {
          mbar_arrive_expect_tx(fb, XBOX_BYTES);
          tma_load_2d(xring + (size_t)slot * XBOX_BYTES, tm, c0, (int)row0, fb, (int)g < C::PRO ? L2_EVICT_FIRST : xpol);
        }
      }
    } else if (warp == EPI_WARPS + 3 && lane == 0) {
      // ---- store thread: the E2b boxes (now holding h') in ring order -> HBM, then the slot is free again
      for (int t = 0; t < a.T; ++t)
        for (int j = 0; j < C::NB; ++j)
          for (int cs = 0; cs < 4; ++cs) {
            const uint32_t g = C::PRO + (uint32_t)t * C::BPS + 2 * C::NKC + j * 4 + cs;
            const int slot = g % C::NXB;
            mbar_wait_tag<WD>(&hdone[cs], (uint32_t)(t * C::NB + j) & 1, 95);
            const uint8_t* box = xring + (size_t)slot * XBOX_BYTES;
            const int col = cs * C::HC + j * 32;
            const bool last = t == a.T - 1;
            if (a.has_hs) tma_store_2d(&tm_hs, box, t * a.hs_step + col, (int)row0);
            if (a.has_hf && last) tma_store_2d(&tm_hf, box, col, (int)row0);
            if (a.has_hs || (a.has_hf && last)) {
              tma_store_commit();
              tma_store_wait_read();
            }
            mbar_arrive_cnt(&xempty[slot], 4);
          }
    } else if (warp == EPI_WARPS + 1 && lane == 0) {
      const uint32_t totW = (uint32_t)a.T * C::WPS;
      for (uint32_t itw = 0; itw < totW; ++itw) {
        const int slot = itw % C::NSW;
        mbar_wait_tag<WD>(&wempty[slot], ((itw / C::NSW) & 1) ^ 1, 90 + slot);
        if (a.dbg & 2) {
          mbar_arrive(&wfull[slot]);
          continue;
        }
        const uint32_t wi = itw % C::WPS;
        const int nb = wi / (C::NKA * C::SPC), r = wi % (C::NKA * C::SPC);
        const int blk = r / C::SPC, part = r % C::SPC;
        const uint8_t* src = (nb < 2 ? a.wg + ((size_t)nb * C::NKA + blk) * C::BLOCK_BYTES
                                     : a.wc + (size_t)blk * C::BLOCK_BYTES);
        uint8_t* dst = wring + (size_t)slot * C::SLOT_BYTES;
        mbar_arrive_expect_tx(&wfull[slot], C::SLOT_BYTES);
        if (C::SPC == 1) {
          bulk_g2s_hint(dst, src, C::SLOT_BYTES, &wfull[slot], L2_EVICT_LAST);
        } else {                               // half a block: chunks {2p, 2p+1} of hi, then of lo
          bulk_g2s_hint(dst, src + (size_t)part * C::SLOT_HALF, C::SLOT_HALF, &wfull[slot], L2_EVICT_LAST);
          bulk_g2s_hint(dst + C::SLOT_HALF, src + 4 * H * 16 + (size_t)part * C::SLOT_HALF, C::SLOT_HALF, &wfull[slot],
                        L2_EVICT_LAST);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem, 512);
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// FP32 row-major [rows, row_stride] tensor -> tiled map with [128 rows x 32 columns] boxes, 128-byte swizzle
int make_tmap_impl(CUtensorMap* tm, const float* base, long rows, long row_stride) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return DESIRE_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)row_stride, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * 4};
  const cuuint32_t box[2] = {32, TM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a [%ld x %ld] FP32 tensor", (int)r, rows, row_stride);
    return DESIRE_ERR_CUDA;
  }
  return DESIRE_OK;
}

struct Maps3 {
  CUtensorMap xp, ex, h0, hs, hf;
};
template <int H, bool EX, bool WD>
int launch3(const Maps3& m, const Gru3Args& a, unsigned grid, cudaStream_t st) {
  DESIRE_ENSURE_SMEM((gru_tc3_kernel<H, EX, WD>), (Cfg3<H, EX>::SMEM));
  DESIRE_LAUNCH(st, (gru_tc3_kernel<H, EX, WD><<<grid, NTHR3, Cfg3<H, EX>::SMEM, st>>>(m.xp, m.ex, m.h0, m.hs, m.hf, a)));
  return DESIRE_OK;
}

bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}

bool tma_ok(const float* p, long ld) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; }

}  // namespace

int make_tmap_rows32(::CUtensorMap_st* tm, const float* base, long rows, long row_stride) {
  return make_tmap_impl(tm, base, rows, row_stride);
}

bool gru_tc3_eligible(const GruSeqArgs& a, const void* pack_ws, size_t pack_bytes) {
  static const bool off = env_flag("DESIRE_GRU_V2");
  if (off || gemm_mode() == 0 || !a.xp || a.traj) return false;
  if (a.R < 64 || a.T < 1) return false;
  if (!tma_ok(a.xp, a.xp_row_stride) || a.xp_step_stride % 4 != 0) return false;
  if (a.xp_row_stride < (long)(a.T - 1) * a.xp_step_stride + 3 * a.H) return false;   // a step is a column window of the row
  if (a.ex) {        // Decoder-2 form: one step, [ex | h] operand, both through TMA
    if (a.H != 128 || a.Ka != a.H || a.T != 1 || !a.h_final) return false;
    if (!tma_ok(a.ex, a.ld_ex) || !tma_ok(a.h0, a.ld_h0) || a.h0_div > 1 || a.ld_ex < a.H || a.ld_h0 < a.H) return false;
  } else {
    if ((a.H != 128 && a.H != 256) || a.Ka != 0) return false;
    if (a.T > 1 && !a.hs && !a.h_final) return false;
  }
  // the states leave through tensor maps as well
  if (a.hs && (!tma_ok(a.hs, a.hs_row_stride) || a.hs_step_stride % 4 != 0 ||
               a.hs_row_stride < (long)(a.T - 1) * a.hs_step_stride + a.H))
    return false;
  if (a.h_final && (!tma_ok(a.h_final, a.ld_hf) || a.ld_hf < a.H)) return false;
  if (!a.hs && !a.h_final) return false;
  if (a.packed ? a.packed_fmt != 3 : (!pack_ws || pack_bytes < gru_tc3_pack_bytes(a.H, a.Ka))) return false;
  return encode_fn() != nullptr;
}

size_t gru_tc3_pack_bytes(int H, int Ka) {
  return align_up(tc_pack_bytes(Ka + H, 2 * H, H)) + align_up(tc_pack_bytes(Ka + H, H, H));
}

int gru_tc3_pack(const float* w_g, const float* w_c, int H, int Ka, void* ws, size_t ws_bytes, cudaStream_t st) {
  DESIRE_CHECK_ARG(ws && ws_bytes >= gru_tc3_pack_bytes(H, Ka), "gru_tc3_pack: workspace too small");
  uint8_t* pg = (uint8_t*)ws;
  DESIRE_TRY(tc_pack_b(w_g, 2 * H, false, Ka + H, 2 * H, H, pg, st));
  DESIRE_TRY(tc_pack_b(w_c, H, false, Ka + H, H, H, pg + align_up(tc_pack_bytes(Ka + H, 2 * H, H)), st));
  return DESIRE_OK;
}

int gru_seq_tc3(const GruSeqArgs& s, void* pack_ws, cudaStream_t st) {
  const int H = s.H;
  if (s.R == 0 || s.T == 0) return DESIRE_OK;
  Gru3Args a{};
  a.R = s.R; a.T = s.T;
  a.xp_step = (int)s.xp_step_stride;
  a.h0 = s.h0; a.h0_div = s.h0_div > 0 ? s.h0_div : 1; a.ld_h0 = s.ld_h0;
  a.has_hs = s.hs ? 1 : 0; a.hs_step = (int)s.hs_step_stride;
  a.has_hf = s.h_final ? 1 : 0;
  a.passes = gemm_mode() == 1 ? 1 : 3;
  a.xp_const = (s.xp_step_stride == 0 && s.T > 1) ? 1 : 0;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("DESIRE_GRU3_DBG");
      dbg = e ? atoi(e) : 0;
    }
    a.dbg = dbg;
  }
  const uint8_t* pg = (const uint8_t*)(s.packed ? s.packed : pack_ws);
  if (!s.packed) DESIRE_TRY(gru_tc3_pack(s.w_g, s.w_c, H, s.Ka, pack_ws, gru_tc3_pack_bytes(H, s.Ka), st));
  a.wg = pg;
  a.wc = pg + align_up(tc_pack_bytes(s.Ka + H, 2 * H, H));
  Maps3 m;
  DESIRE_TRY(make_tmap_rows32(&m.xp, s.xp, s.R, s.xp_row_stride));
  m.ex = m.h0 = m.hs = m.hf = m.xp;
  if (s.ex) {
    DESIRE_TRY(make_tmap_rows32(&m.ex, s.ex, s.R, s.ld_ex));
    DESIRE_TRY(make_tmap_rows32(&m.h0, s.h0, s.R, s.ld_h0));
  }
  if (s.hs) DESIRE_TRY(make_tmap_rows32(&m.hs, s.hs, s.R, s.hs_row_stride));
  if (s.h_final) DESIRE_TRY(make_tmap_rows32(&m.hf, s.h_final, s.R, s.ld_hf));
  const unsigned grid = (unsigned)(((long)s.R + TM - 1) / TM);
  static const bool wd = env_flag("DESIRE_GRU3_WATCHDOG");
  if (s.ex) return wd ? launch3<128, true, true>(m, a, grid, st) : launch3<128, true, false>(m, a, grid, st);
  if (H == 128) return wd ? launch3<128, false, true>(m, a, grid, st) : launch3<128, false, false>(m, a, grid, st);
  return wd ? launch3<256, false, true>(m, a, grid, st) : launch3<256, false, false>(m, a, grid, st);
}

}  // namespace desire
