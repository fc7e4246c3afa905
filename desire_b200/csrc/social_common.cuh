// Device helpers shared by the fused social kernels (social_ts.cu: pooling + fc for N <= 128; social_pm.cu: pooling as an MMA
// for large scenes): tcgen05.mma with the A operand in tensor memory, log-polar bins of a pair from tables in registers,
// byte compares that turn a row of bin ids into BF16 0/1 selection-matrix words.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace social {

using namespace tc;

// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand is read from tensor memory (lane = row, one 32-bit column = two
// consecutive K values, lower half = the smaller k; pinned by tests/test_gpu_selftest.py)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// single-thread roles wait with a suspend-time hint so their polling does not take issue slots from the other warps
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
  } while (!ok);
}

// Log-polar bins of BOTH directions of a pair from tables held in registers (8 squared radial edges padded with +inf,
// 8 sector directions): d = pos_j - pos_i gives the bin of j as seen from i (fwd), -d the bin of i as seen from j (bwd).
// The same arithmetic and the same decisions as logpolar_bin() in common.cuh, branch-free: r2 is the same for both
// directions, and every cross product of -d is exactly the negated cross product of d (negation commutes with the
// roundings), so "cross(-d) >= 0" is "cross(d) <= 0".  n_rad <= 7, n_ang <= 8.
__device__ __forceinline__ void logpolar_bin_pair(float dx, float dy, const float (&re)[8], const float (&dr)[16], int n_rad,
                                                  int n_ang, int& fwd, int& bwd) {
  const float r2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  int rb = -1;
  uint32_t ge = 0, le = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) rb += (r2 >= re[e]) ? 1 : 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const float c = __fsub_rn(__fmul_rn(dr[2 * s], dy), __fmul_rn(dr[2 * s + 1], dx));
    ge |= (c >= 0.f ? 1u : 0u) << s;
    le |= (c <= 0.f ? 1u : 0u) << s;
  }
  const uint32_t all = (1u << n_ang) - 1u;
  ge &= all;
  le &= all;
  // the first sector s with ge[s] and not ge[s+1] (cyclically), else the last one
  const uint32_t hf = ge & ~((ge >> 1) | ((ge & 1u) << (n_ang - 1)));
  const uint32_t hb = le & ~((le >> 1) | ((le & 1u) << (n_ang - 1)));
  const bool out = rb < 0 || rb >= n_rad;
  fwd = out ? -1 : rb * n_ang + (hf ? __ffs(hf) - 1 : n_ang - 1);
  bwd = out ? -1 : rb * n_ang + (hb ? __ffs(hb) - 1 : n_ang - 1);
}

// the tile's stage list: bins with at least one pair; a tile without any pair runs one stage on (empty) bin 0, which
// leaves D = 0 for the epilogue
__device__ __forceinline__ uint64_t active_bins(const uint32_t* binmask) {
  const uint64_t m = (uint64_t)binmask[0] | ((uint64_t)binmask[1] << 32);
  return m ? m : 1ull;
}

// 0x80 in every byte of w that equals the byte replicated in g4 (exact per byte, no cross-byte carries)
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t g4) {
  const uint32_t t = w ^ g4;
  return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
}
// two flag bytes (0x80 / 0) -> two BF16 values 1.0 / 0.0 (0x80 * 0x7F = 0x3F80)
__device__ __forceinline__ uint32_t ones_lo(uint32_t z) { return __byte_perm(z, 0u, 0x4140) * 0x7Fu; }
__device__ __forceinline__ uint32_t ones_hi(uint32_t z) { return __byte_perm(z, 0u, 0x4342) * 0x7Fu; }

}  // namespace social
}  // namespace desire
