// sm_100a primitives used by the tensor-core kernels: mbarrier, 1-D bulk TMA copies, tcgen05
// (TMEM alloc / MMA / commit / ld), UMMA descriptors, and the BF16 hi/lo operand split.
//
// Operand layout (both A and B, "K-major, no swizzle", canonical UMMA interleave): a tile of
// `rows` x `kc` K-chunks (one chunk = 8 bf16 = 16 B) is stored chunk-major,
//     byte(row, k) = (k/8) * rows*16 + row*16 + (k%8)*2
// i.e. 8 rows x 16 B core matrices, SBO (next 8-row group) = 128 B, LBO (next K chunk) = rows*16 B.
// A warp that writes 32 consecutive rows of one chunk stores 512 contiguous bytes (no conflicts).
//
// 3xBF16: x = hi + lo (+ <=2^-17|x|); A@B ~= Ahi@Bhi + Alo@Bhi + Ahi@Blo with FP32 accumulation in
// TMEM gives ~1e-5 relative error — the mode that meets the 1e-4 parity bar on tensor cores.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace desire {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Position in a ring of n slots with the mbarrier phase of the current lap, counted up: `it % n` and `it / n` with a
// run-time n are ~20 dependent integer instructions each, which an MMA-issuing warp cannot afford between two blocks of
// MMAs (social_ts.cu: that division, not the tensor pipe, bounded the kernel).
struct RingPos {
  uint32_t slot = 0, ph = 0;
  __device__ __forceinline__ void next(uint32_t n) {
    if (++slot == n) {
      slot = 0;
      ph ^= 1;
    }
  }
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// mbar_wait with a watchdog: a protocol error in a warp-specialised kernel shows up as a hang; after ~2 s of waiting
// this names the barrier (tag), the waiter and the phase it wanted, then traps so the launch fails instead of hanging.
static __device__ __noinline__ void mbar_stuck(int tag, uint32_t parity) {
  printf("desire: mbarrier wait stuck: tag %d parity %u block %d thread %d\n", tag, parity, (int)blockIdx.x, (int)threadIdx.x);
  __trap();
}
// WD = false compiles to the plain wait (the watchdog's cold call costs registers around every wait site).
template <bool WD>
__device__ __forceinline__ void mbar_wait_tag(uint64_t* bar, uint32_t parity, int tag) {
  if (!WD) {
    mbar_wait(bar, parity);
    return;
  }
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) mbar_stuck(tag, parity);
  }
}
// non-blocking probe (a thread that serves two rings polls both)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ---------------------------------------------------------------- TMA (1-D bulk copy, global -> shared)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 eviction-priority policies (the constants CUTLASS passes as TMA cache hints)
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------- TMA (2-D tiled tensor map, global -> shared)
// `tmap` is the address of a __grid_constant__ CUtensorMap kernel parameter; (c0, c1) = (innermost element, row) of
// the box origin.  Elements outside the tensor are zero-filled and still count towards the transaction bytes.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// shared -> global through a tensor map (SASS UTMASTG); rows / columns outside the tensor are clipped.  The writer's
// generic-proxy stores must be fenced (fence.proxy.async) before the issuing thread learns that they are done.
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent one have
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all committed bulk stores of this thread are complete (required before the CTA exits)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// swz128 for a box whose base is only 128-byte aligned: the XOR term comes from the row's absolute address
__device__ __forceinline__ uint32_t swz128_abs(uint32_t box_saddr, int row, int c16) {
  const uint32_t ra = box_saddr + row * 128;
  return ra + ((c16 ^ ((ra >> 7) & 7)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// Byte offset of 16-byte chunk `c16` (0..7) of row `row` inside a CU_TENSOR_MAP_SWIZZLE_128B box whose rows are
// 128 bytes and whose base is 1024-byte aligned: the chunk index is XOR-ed with the row's low three bits.
__device__ __forceinline__ uint32_t swz128(int row, int c16) { return (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4)); }

// ---------------------------------------------------------------- tcgen05
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp (elect.sync).  Issue tcgen05.mma / commit / bulk copies as
//     if (elect_one()) mma_bf16(...)
// inside code the WHOLE warp executes with warp-uniform operands: the compiler then keeps descriptors and addresses in
// uniform registers.  Under a per-thread condition such as `if (lane == 0)` it wraps every UTCHMMA in a lane-broadcast
// loop (VOTEU / ELECT / R2UR / BRA.U.ANY) and one thread retires an MMA only every ~93 cycles whatever its shape
// (tools/mma_rate.py: M=128, N=128 needs 64).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T, BF16 inputs, FP32 accumulate, M=128 per CTA
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive FP32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// K-major, SWIZZLE_NONE shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// the same descriptor `bytes` further into shared memory (only the 14-bit start-address field changes: no carry out of
// it as long as the operand stays inside the 256 KB window) — one add instead of rebuilding the descriptor per MMA
__device__ __forceinline__ uint64_t desc_adv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
// cute::UMMA::InstrDescriptor for kind::f16: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- BF16 hi/lo split of 8 consecutive K values
struct Split8 {
  uint4 hi, lo;
};
// 2 elements per cvt (cvt.rn.bf16x2.f32), ~3 instructions per element:
// hi = bf16x2(x1,x0); hi as floats = bits<<16 / bits&0xffff0000; lo = bf16x2(x1-hi1, x0-hi0).
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ Split8 split8(const float* x) {
  Split8 s;
  split2(x[0], x[1], s.hi.x, s.lo.x);
  split2(x[2], x[3], s.hi.y, s.lo.y);
  split2(x[4], x[5], s.hi.z, s.lo.z);
  split2(x[6], x[7], s.hi.w, s.lo.w);
  return s;
}

// fast activations: ex2.approx (2 ulp) + rcp.approx (1 ulp); 4-5 instructions each
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpa(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// branch-free activation for store loops (a data-dependent branch per element keeps the next element's loads behind
// it): ELU and sigmoid through ex2.approx (2 ulp), the others exact
__device__ __forceinline__ float act_fast(float x, int act) {
  const float e = ex2a(1.4426950408889634f * (act == 3 /*DESIRE_ACT_SIGMOID*/ ? -x : fminf(x, 0.f)));
  const float elu = x > 0.f ? x : e - 1.f;
  const float sig = rcpa(1.f + e);
  const float relu = fmaxf(x, 0.f);
  return act == 2 /*ELU*/ ? elu : (act == 3 ? sig : (act == 1 /*RELU*/ ? relu : x));
}
__device__ __forceinline__ float sigmoid_a(float x) { return rcpa(1.f + ex2a(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_a(float x) {     // 1 - 2/(1+e^{2x}); saturates to +-1 for large |x|
  return fmaf(-2.f, rcpa(1.f + ex2a(2.8853900817779268f * x)), 1.f);
}

}  // namespace tc
}  // namespace desire
