// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, 1-D bulk async copy
// (UBLKCP), tcgen05 alloc / mma / commit / ld / st, UMMA shared-memory + instruction descriptors.
// Hand-written inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace dvd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU. The bound is wall-clock
// (about 2 s of SM cycles), not a spin count: one mbarrier.try_wait may itself suspend the thread for a system-dependent
// time, so a count of polls says nothing about how long a dead wait lasts.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("dvd_b200: mbarrier wait timeout (block %d thread %d bar %p parity %u)\n", blockIdx.x, threadIdx.x,
             (void*)bar, parity);
      __trap();
    }
  }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (UMMA / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- 1-D bulk async copy global -> shared, completion on an mbarrier ---------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tcgen05: TMEM allocation -----------------------------------------------------------------
// whole warp must execute; ncols power of two in [32, 512]
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand stored as the canonical SWIZZLE_128B image:
// rows of 128 bytes (64 bf16 along K), 8-row atoms of 1024 bytes, 16-byte chunk c of row r stored
// at chunk (c ^ (r & 7)). SBO = 1024 (distance between 8-row groups), LBO unused (one atom along K).
// Advancing K inside the 128-byte row = adding the byte offset (>>4) to the start address.
__device__ __forceinline__ uint64_t make_sdesc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;      // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}

// MN-major operand (the MN index is contiguous in memory, K strided), SWIZZLE_128B image:
// 64-element MN atoms; inside an atom one 128-byte row per K index, 8-row groups of 1024 bytes
// (SBO), 16-byte chunk c of row r stored at chunk (c ^ (r & 7)); next MN atom `lbo_bytes` further
// (LBO = rows_of_K * 128). Advancing K by 16 rows = +2048 bytes on the start address.
__device__ __forceinline__ uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MN-major operand WITHOUT swizzle ("interleave"): 8x8-element core matrices of 128 contiguous bytes
// (8 K-rows of 16 bytes = 8 MN elements each); K-groups of 8 rows LBO = 128 bytes apart, MN-groups of 8 elements
// SBO bytes apart. Element (mn, k) lives at (mn/8)*SBO + (k/8)*128 + (k%8)*16 + (mn%8)*2, so 32 threads that own
// 32 consecutive K indices (pixels) write 512 contiguous bytes per 16-byte store. K advance of 16 rows = +256 bytes.
__device__ __forceinline__ uint64_t make_sdesc_mn_interleave(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;        // LBO: next group of 8 K rows
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // SBO: next group of 8 MN elements
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)0 << 61;                              // SWIZZLE_NONE
  return d;
}
// byte offset of element (mn, k) in a [rows x 64 k] interleaved MN-major block (SBO = 1024)
__host__ __device__ __forceinline__ uint32_t il_offset(uint32_t mn, uint32_t k) {
  return (mn >> 3) * 1024u + k * 16u + (mn & 7u) * 2u;
}

// Instruction descriptor for kind::f16: A,B = bf16, D = fp32. a_mn / b_mn: operand is MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4)                    // c_format = F32
         | (1u << 7)                  // a_format = BF16
         | (1u << 10)                 // b_format = BF16
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// ---- tcgen05.mma ------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T    (both operands K-major)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T    (A: lane = row, packed bf16 pairs along columns)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread complete -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- tcgen05.ld / st : 32 lanes x 32 bit, N consecutive columns per thread ----------------------
// taddr must carry the warp's lane quarter: ((warp_id & 3) * 32) << 16 | column
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r0) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r0) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r0) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- fp32 -> (hi, lo) bf16 split:  x ~= hi + lo with |x - hi - lo| <= 2^-17 |x| -------------------
// pack two consecutive K elements into one 32-bit word (element 2j in the low half)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);  // .x = x0 (low half), .y = x1
  float r0 = x0 - __bfloat162float(h.x);
  float r1 = x1 - __bfloat162float(h.y);
  __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// sin / cos with a two-constant Cody-Waite reduction to [-pi, pi] followed by the MUFU approximations:
// absolute error ~5e-7 for |a| < 2^13 (embedding arguments are f_k * coordinate, f_k <= 17), i.e. below the
// rounding error the fp32 product f_k * x already carries; no slow path, no local memory.
__device__ __forceinline__ void fast_sincos(float a, float& s, float& c) {
  const float n = rintf(a * 0.15915494309189535f);
  float r = fmaf(-n, 6.28318548202514648f, a);        // 2*pi rounded to fp32
  r = fmaf(-n, -1.74845553e-7f, r);                   // 2*pi - fp32(2*pi)
  s = __sinf(r);
  c = __cosf(r);
}

// byte offset of element (row, k) inside one [rows x 64 bf16] SWIZZLE_128B K-major block
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ row) & 7u) << 4) + (k & 7u) * 2u;
}

// byte offset of element (mn, k) inside one MN-major SWIZZLE_128B block with 64 K-rows (k < 64)
__host__ __device__ __forceinline__ uint32_t mn128_offset(uint32_t mn, uint32_t k) {
  return (mn >> 6) * 8192u + (k >> 3) * 1024u + (k & 7u) * 128u + (((((mn & 63u) >> 3) ^ k) & 7u) << 4) + (mn & 7u) * 2u;
}
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

}  // namespace tc
}  // namespace dvd
