// Hardware self-tests of the tcgen05 forms the kernels rely on (run through the C-ABI by tests/test_gpu_selftest.py).
//
// desire_selftest_tsmma: D = A @ B^T with the A operand read from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc —
// CUTLASS: SM100_MMA_F16BF16_TS) against the same product with A in shared memory.  It pins the TMEM layout of a BF16 A
// operand that the fused social kernel writes with tcgen05.st: lane = row, one 32-bit column = two consecutive K
// elements, lower half = the smaller k.
#include "common.cuh"
#include "tc.cuh"

namespace desire {
namespace {
using namespace tc;

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A [128,32], B [64,32] FP32 (rounded to BF16 inside); out_ss / out_ts [128,64]
__global__ void __launch_bounds__(160) tsmma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ out_ss, float* __restrict__ out_ts, int order) {
  __shared__ __align__(1024) uint8_t sa[(K / 8) * M * 16];
  __shared__ __align__(1024) uint8_t sb[(K / 8) * N * 16];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<256>(&tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;               // columns [0,64) D_ss, [64,128) D_ts, [128,144) A
  if (tid < M) {
    // shared-memory A (K-major, no swizzle) and tensor-memory A (packed pairs)
    float v[K];
    for (int k = 0; k < K; ++k) v[k] = A[tid * K + k];
    uint32_t packed[K / 2];
    for (int c = 0; c < K / 8; ++c) {
      uint32_t w[4];
      for (int p = 0; p < 4; ++p) {
        const float lo = v[c * 8 + 2 * p], hi = v[c * 8 + 2 * p + 1];
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[p]) : "f"(hi), "f"(lo));      // upper half = first source
        packed[c * 4 + p] = order == 0 ? w[p] : ((w[p] >> 16) | (w[p] << 16));
      }
      *reinterpret_cast<uint4*>(sa + c * M * 16 + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + 128;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                 "%15, %16};" ::"r"(trow),
                 "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]), "r"(packed[6]),
                 "r"(packed[7]), "r"(packed[8]), "r"(packed[9]), "r"(packed[10]), "r"(packed[11]), "r"(packed[12]),
                 "r"(packed[13]), "r"(packed[14]), "r"(packed[15])
                 : "memory");
    tmem_st_wait();
    if (tid < N) {
      float b[K];
      for (int k = 0; k < K; ++k) b[k] = B[tid * K + k];
      for (int c = 0; c < K / 8; ++c) {
        uint32_t w[4];
        for (int p = 0; p < 4; ++p) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[p]) : "f"(b[c * 8 + 2 * p + 1]), "f"(b[c * 8 + 2 * p]));
        *reinterpret_cast<uint4*>(sb + c * N * 16 + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = idesc_bf16(M, N);
    for (int j = 0; j < K / 16; ++j) {
      const uint64_t da = smem_desc(smem_u32(sa) + j * 2 * M * 16, M * 16, 128);
      const uint64_t db = smem_desc(smem_u32(sb) + j * 2 * N * 16, N * 16, 128);
      mma_bf16(tmem, da, db, idesc, j > 0);
      mma_bf16_ts(tmem + 64, tmem + 128 + j * 8, db, idesc, j > 0);
    }
    mma_commit(&bar[0]);
  }
  if (tid < M) {
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      float acc[32];
      tmem_ld32(trow + c0, acc);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out_ss[tid * N + c0 + j] = acc[j];
      tmem_ld32(trow + 64 + c0, acc);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out_ts[tid * N + c0 + j] = acc[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}


// How fast does one thread's stream of tcgen05.mma (M = 128, K = 16, BF16) retire?  mode 0: both operands in shared
// memory, 1: A in tensor memory.  Every CTA issues `iters` MMAs on the same operands; cycles of block 0 are returned.
__global__ void __launch_bounds__(640) mma_rate_kernel(int mode_, int N, int iters, long long* __restrict__ out) {
  // mode = 100 * s + m: mode m with the A operand of the shared-memory descriptors starting s * 16 bytes into its first
  // 8-row core matrix (conv5_tc.cu addresses filter taps this way)
  const int mode = mode_ % 100, ashift = (mode_ / 100) * 16;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar, dummy, done0;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 64 * 2 + 256 * 64 * 2) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, (mode == 4 || mode == 13 || mode >= 14) ? 2 : 1);
    mbar_init(&dummy, 1 << 20);
    mbar_init(&done0, 1);
    mbar_arrive(&done0);                                         // phase 0 of done0 is complete from the start
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  // mode 4: two threads (warps 0 and 1) issue at the same time, each into its own accumulator (TS)
  if ((tid == 0 && mode < 5) || (mode == 4 && tid == 32)) {
    const int w = tid >> 5;
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t sa = smem_u32(sm), sb = sa + 128 * 64 * 2;
    uint64_t da[4], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                                // four K steps of a 64-wide stage, round robin
      da[j] = smem_desc(sa + j * 2 * 128 * 16, 128 * 16, 128);
      db[j] = smem_desc(sb + j * 2 * N * 16, N * 16, 128);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // modes 2, 3: consecutive MMAs alternate between two accumulators (no back-to-back dependency on D)
        const uint32_t d = (mode == 2 || mode == 3) ? tmem + (j & 1) * 128 : tmem + w * 128;
        if (mode == 0 || mode == 2) mma_bf16(d, da[j], db[j], idesc, 1);
        else mma_bf16_ts(d, tmem + 256 + w * 64 + j * 8, db[j], idesc, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
  }
  // mode 5: the whole warp runs the issue loop with warp-uniform operands and one ELECTED lane issues (TS) — the
  // compiler then feeds UTCHMMA from uniform registers instead of wrapping every MMA in a lane-broadcast loop
  // mode 14: the MMA mix of one stage of the fused social kernel, nothing else running: warp 1 issues 24 MMAs with A in
  // tensor memory (the fc), warp 2 sixteen with both operands in shared memory (the pooling), per round; N = 128
  // modes 18-21: a miniature of the fused social kernel's stage protocol (TMEM D | P | A0 | A1): warp 1 issues the fc MMAs
  // (24 TS) of stage g when A(g) is announced and commits aempty; warp 2 issues pool(g) (16 SS) when P is free and
  // commits pfull; 16 finisher warps wait for pfull, tcgen05.ld P, release it, convert, wait for aempty(g-2),
  // tcgen05.st A(g), announce it.  18: all of it, 19: handshakes only (no ld / st), 20: 18 with the fc MMAs issued in four
  // elect blocks with a commit each (the weight ring), 21: 18 with P released only after the conversion
  if (mode >= 18) {
    __shared__ uint64_t pfull, pempty, afull[2], aempty[2];
    if (tid == 0) {
      mbar_init(&pfull, 1);
      mbar_init(&pempty, 16);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&afull[i], 16);
        mbar_init(&aempty[i], 1);
      }
      fence_barrier_init();
    }
    __syncthreads();
    const int nst = iters / 40;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t sa = smem_u32(sm), sb = sa + 128 * 64 * 2;
    uint64_t da[4], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      da[j] = smem_desc(sa + j * 2 * 128 * 16, 128 * 16, 128);
      db[j] = smem_desc(sb + j * 2 * N * 16, N * 16, 128);
    }
    const long long t0 = clock64();
    if (warp == 1) {
      for (int g = 0; g < nst; ++g) {
        mbar_wait(&afull[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tm + 256 + (g & 1) * 128;
        if (mode == 20) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < 6; ++j) mma_bf16_ts(tm, ta + ((c * 6 + j) & 7) * 8, db[j & 3], idesc, 1);
              mma_commit(&dummy);
              if (c == 3) mma_commit(&aempty[g & 1]);
            }
            __syncwarp();
          }
        } else {
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 24; ++j) mma_bf16_ts(tm, ta + (j & 7) * 8, db[j & 3], idesc, 1);
            mma_commit(&aempty[g & 1]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0 && tid == 32) out[0] = t1 - t0;
    } else if (warp == 2) {
      for (int g = 0; g < nst; ++g) {
        if (g > 0) mbar_wait(&pempty, (g - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mma_bf16(tm + 128, da[j & 3], db[j & 3], idesc, j > 0);
          mma_commit(&pfull);
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
    } else if (warp >= 3) {
      const int w = warp - 3, q4 = warp & 3, cg = w >> 2;
      const uint32_t lane_f = (uint32_t)(32 * q4) << 16;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
      for (int g = 0; g < nst; ++g) {
        mbar_wait(&pfull, g & 1);
        tc_fence_after();
        if (mode != 19) {
          tmem_ld32(tm + lane_f + 128 + cg * 32, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        if (mode != 21) {
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&pempty);
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split2(v[2 * i] * 0.5f, v[2 * i + 1] * 0.5f, hi[i], lo[i]);
        if (mode == 21) {
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&pempty);
        }
        if (mode >= 22) {                                        // 22: fence.proxy.async per thread and stage (what follows
          if (mode == 23) {                                      // the selection-matrix stores), 23: with two 16-byte stores
            uint4* q = reinterpret_cast<uint4*>(sm + 128 * 64 * 2 + 256 * 64 * 2) + (tid - 96);
            q[0] = make_uint4(hi[0], hi[1], lo[0], lo[1]);
            q[512] = make_uint4(hi[2], hi[3], lo[2], lo[3]);
          }
          fence_proxy_async();
        }
        if (g >= 2) mbar_wait(&aempty[g & 1], ((g - 2) >> 1) & 1);
        tc_fence_after();
        if (mode != 19) {
          const uint32_t ta = tm + lane_f + 256 + (g & 1) * 128 + cg * 16;
          tmem_st16(ta, reinterpret_cast<const float*>(hi));
          tmem_st16(ta + 64, reinterpret_cast<const float*>(lo));
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&afull[g & 1]);
      }
    }
  }
  // modes 15-17: the same, with 16 more warps that 15: poll the final barrier (what waiting producer warps do),
  // 16: stream 16-byte shared-memory stores and loads next to the operands, 17: issue tcgen05.ld of idle columns
  if (mode >= 15 && mode <= 17 && warp >= 3) {
    if (mode == 15) {
      mbar_wait(&bar, 0);
    } else if (mode == 16) {
      uint4* q = reinterpret_cast<uint4*>(sm + 128 * 64 * 2 + 256 * 64 * 2) + (tid - 96);
      uint4 v = make_uint4(tid, 0, 0, 0);
      while (!mbar_try_wait(&bar, 0)) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          q[0] = v;
          v.x += q[512 * (r & 1)].y;
        }
      }
      if (v.x == 0x7fffffff) out[1] = 1;
    } else {
      const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 384 + 32 * ((warp - 3) >> 2 & 3);
      float v[32], acc = 0.f;
      while (!mbar_try_wait(&bar, 0)) {
        tmem_ld32(ta, v);
        tmem_ld_wait();
        acc += v[0];
      }
      if (acc == 123.456f) out[1] = 1;
    }
  }
  if (mode >= 14 && mode <= 17 && (warp == 1 || warp == 2)) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t sa = smem_u32(sm), sb = sa + 128 * 64 * 2;
    uint64_t da[4], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      da[j] = smem_desc(sa + j * 2 * 128 * 16, 128 * 16, 128);
      db[j] = smem_desc(sb + j * 2 * N * 16, N * 16, 128);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 40) {
      if (warp == 1) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 24; ++j) mma_bf16_ts(tm, tm + 256 + (j & 3) * 8, db[j & 3], idesc, 1);
        }
      } else {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mma_bf16(tm + 128, da[j & 3], db[j & 3], idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 32) out[0] = t1 - t0;
  }
  // mode 13: TWO warps issue concurrently (elected lanes, uniform operands), each into its own accumulator
  if (mode == 13 && (warp == 1 || warp == 2)) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t sb = smem_u32(sm) + 128 * 64 * 2;
    uint64_t db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) db[j] = smem_desc(sb + j * 2 * N * 16, N * 16, 128);
    const uint32_t d = tm + (warp - 1) * 128, aa = tm + 256 + (warp - 1) * 64;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (elect_one()) mma_bf16_ts(d, aa + j * 8, db[j], idesc, 1);
    }
    if (elect_one()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 32) out[0] = t1 - t0;
  }
  if (mode >= 5 && mode != 13 && mode < 14 && warp == 1) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t sb = smem_u32(sm) + 128 * 64 * 2;
    uint64_t da[4], db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      da[j] = smem_desc(smem_u32(sm) + ashift + j * 2 * 128 * 16, 128 * 16, 128);
      db[j] = smem_desc(sb + j * 2 * N * 16, N * 16, 128);
    }
    const long long t0 = clock64();
    if (mode == 12) {
      // how many MMAs does the queue behind an issuing thread take before the issue itself blocks?  out[1 + k] = cycles
      // after which the (4k + 4)-th MMA had been ISSUED
      for (int k = 0; k < 16; ++k) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_bf16_ts(tm, tm + 256 + j * 8, db[j], idesc, 1);
        }
        __syncwarp();
        if (blockIdx.x == 0 && tid == 32) out[1 + k] = clock64() - t0;
      }
    } else if (mode == 10 || mode == 11) {
      // alternate between two accumulators: 10 = every MMA, 11 = every 6 MMAs
      for (int i = 0; i < iters; i += 12) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const uint32_t d = tm + ((mode == 10 ? j : j / 6) & 1) * 128;
            mma_bf16_ts(d, tm + 256 + (j & 3) * 8, db[j & 3], idesc, 1);
          }
        }
      }
    } else if (mode == 7 || mode == 8 || mode == 9) {
      // 7: a commit (to a barrier nobody waits on) after every 6 MMAs; 8: plus a wait on an already completed barrier and
      // the fence in front of every group of 6; 9: wait + fence only (no commits)
      for (int i = 0; i < iters; i += 6) {
        if (mode >= 8) {
          mbar_wait(&done0, 0);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 6; ++j) mma_bf16_ts(tm, tm + 256 + (j & 3) * 8, db[j & 3], idesc, 1);
          if (mode != 9) mma_commit(&dummy);
        }
      }
    } else if (mode == 5) {
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (elect_one()) mma_bf16_ts(tm, tm + 256 + j * 8, db[j], idesc, 1);
      }
    } else {                                                     // mode 6: both operands in shared memory
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (elect_one()) mma_bf16(tm, da[j], db[j], idesc, 1);
      }
    }
    if (elect_one()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 32) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace desire

extern "C" int desire_selftest_tsmma(const float* A, const float* B, float* out_ss, float* out_ts, int order,
                                     desire_stream_t stream) {
  DESIRE_CHECK_ARG(A && B && out_ss && out_ts && (order == 0 || order == 1), "desire_selftest_tsmma: bad arguments");
  desire::tsmma_kernel<<<1, 160, 0, (cudaStream_t)stream>>>(A, B, out_ss, out_ts, order);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}

extern "C" int desire_selftest_mma_rate(int mode, int N, int iters, int grid, long long* out_cycles, desire_stream_t stream) {
  DESIRE_CHECK_ARG(out_cycles && mode >= 0 && mode % 100 <= 23 && mode / 100 <= 7 && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && grid > 0,
                   "desire_selftest_mma_rate: bad arguments");
  const size_t smem = 128 * 64 * 2 + 256 * 64 * 2 + (mode % 100 == 16 || mode % 100 == 23 ? 1024 * 16 : 0);
  DESIRE_ENSURE_SMEM(desire::mma_rate_kernel, smem);
  desire::mma_rate_kernel<<<grid, mode % 100 >= 15 ? 608 : 128, smem, (cudaStream_t)stream>>>(mode, N, iters, out_cycles);
  DESIRE_LAUNCH_CHECK();
  return DESIRE_OK;
}
