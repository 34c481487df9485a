// Inline-PTX helpers shared by the kernels of libqb200 (sm_100a only): mbarriers, TMA bulk copies,
// packed FP32 math, tcgen05 (tensor core / tensor memory) wrappers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace qb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared through the TMA unit (SASS: UBLKCP); bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Packed FP32 pairs (Blackwell FFMA2, PTX fma.rn.f32x2): one instruction = two IEEE round-to-nearest
// FMAs.  A three-register scalar FFMA issues every other cycle per SM sub-partition, so a scalar
// kernel tops out at half the FP32 lanes; FFMA2 is what fills all 128.  ptxas folds a pair built
// from one register twice ({c, c}) into a broadcast operand (SASS "R.F32"), so the codevector element
// needs no duplicate register.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {  // SASS FMNMX3
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}


__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must end the kernel with an error, never hang the GPU.
// RELAXED waiters (roles that run far ahead of the role they wait for) give try_wait a suspend-time hint of
// RELAXED_NS so that their waiting does not take issue slots from the warps doing the work.
template <int RELAXED_NS = 0>
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity, int tag) {
  constexpr bool RELAXED = RELAXED_NS > 0;
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (unsigned int spin = 0; spin < (RELAXED ? (1u << 20) : (1u << 26)); spin++) {
    uint32_t ok;
    if (RELAXED) {
      // try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint
      // expires) instead of the thread polling - no issue slots taken from the working warps, no wake-up skew
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(ok)
          : "r"(addr), "r"(parity), "r"((uint32_t)RELAXED_NS)
          : "memory");
    } else {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(ok)
          : "r"(addr), "r"(parity)
          : "memory");
    }
    if (ok) return;
  }
  printf("libqb200: mbarrier wait timed out (tag %d, block %d, thread %d)\n", tag, (int)blockIdx.x, (int)threadIdx.x);
  __trap();
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05: 5th-generation tensor cores, accumulators in tensor memory (TMEM) ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Whole warp: allocate `cols` TMEM columns (power of two >= 32); the base address lands in *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// Shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; lbo = byte
// distance between the two core matrices along K, sbo = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t umma_idesc_bf16_f32(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// One thread: D[tmem] (+)= A[smem] * B[smem]^T, 128 x N x 16.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// One thread: the mbarrier gets one arrival when every MMA this thread issued so far has finished.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Whole warp: 8 consecutive FP32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
// Ties eight loaded values to the preceding wait: asm volatile statements keep their order, and the
// "+f" operands make every later use of v[] depend on this statement, so the compiler cannot hoist
// arithmetic on the registers above tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld_pin8(float *v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace qb
