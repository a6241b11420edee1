// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA/TMEM).
// Encodings follow the PTX ISA 8.7 tcgen05 chapter; the shared-memory matrix descriptor and
// instruction descriptor bit layouts are restated in gemm_tcgen05.cu next to their use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b2p {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (and fails the launch) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b2p: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// Elected forms for a converged warp (see umma_*_stage below): every lane executes, one issues.
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}\n"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}\n"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_elect(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}\n"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 operands with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- warp-converged issue: the WHOLE warp executes these (operands warp-uniform); elect.sync picks the same lane every
// time (deterministic for a full member mask), which issues the tensor-core instructions.  No C++ divergence around the
// tcgen05.mma, so the compiler keeps the descriptors in uniform registers instead of wrapping every UTCHMMA in a
// per-active-thread serialisation loop (profiles/r2_notes.md: the single MMA-issuing thread was instruction-bound).
// Descriptors are passed as 32-bit halves: lo = (smem address >> 4) | LBO field, hi = constant per launch; advancing K by
// 16 elements inside the swizzle span is +2 on lo.  KSTEPS k-steps of one (A, B) stage per call.
template <int KSTEPS>
__device__ __forceinline__ void umma_f16_stage(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  static_assert(KSTEPS == 2 || KSTEPS == 4, "k-block of 32 or 64 elements");
  if constexpr (KSTEPS == 2) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t.reg .b32 a1, b1;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "add.u32 a1, %1, 2;\n\tadd.u32 b1, %2, 2;\n\t"
        "mov.b64 da, {a1, %3};\n\tmov.b64 db, {b1, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t.reg .b32 a1, b1;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "add.u32 a1, %1, 2;\n\tadd.u32 b1, %2, 2;\n\t"
        "mov.b64 da, {a1, %3};\n\tmov.b64 db, {b1, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t"
        "add.u32 a1, %1, 4;\n\tadd.u32 b1, %2, 4;\n\t"
        "mov.b64 da, {a1, %3};\n\tmov.b64 db, {b1, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t"
        "add.u32 a1, %1, 6;\n\tadd.u32 b1, %2, 6;\n\t"
        "mov.b64 da, {a1, %3};\n\tmov.b64 db, {b1, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// fp16x3 operands: per k-step hi*hi (+)= , hi*lo +=, lo*hi += into the same accumulator
template <int KSTEPS>
__device__ __forceinline__ void umma_f16x3_stage(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                 uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
#pragma unroll
  for (int k = 0; k < KSTEPS; ++k) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 dah, dal, dbh, dbl;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b64 dah, {%1, %5};\n\tmov.b64 dal, {%2, %5};\n\tmov.b64 dbh, {%3, %5};\n\tmov.b64 dbl, {%4, %5};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %6, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %6, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %6, 1;\n\t}\n"
        ::"r"(tmem_d), "r"(a_hi + 2u * k), "r"(a_lo + 2u * k), "r"(b_hi + 2u * k), "r"(b_lo + 2u * k), "r"(desc_hi), "r"(idesc),
          "r"(k == 0 ? accumulate : 1u)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
      ::"r"(bar)
      : "memory");
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, with the destination registers of an in-flight tcgen05.ld tied to it ("+r"), so that no use of them can be
// scheduled above the wait when other work is interleaved between the load and the wait (software-pipelined epilogue).
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :: "memory");
}

}  // namespace b2p
