// Self-test of the tcgen05 building blocks (descriptors, SS / TS operand modes, bf16x3 split,
// TMEM load/store, commit -> mbarrier). One CTA computes D[128,N] = A[128,K] * B[N,K]^T with the
// same primitives the scene-flow MLP kernels use; the host compares against an fp32 GEMM.
#include "common.cuh"
#include "tc_common.cuh"

namespace dvd {
using namespace tc;

// mode 0: A from shared memory (SS), mode 1: A from tensor memory (TS). passes: 1 (bf16) or 3 (bf16x3)
__global__ void __launch_bounds__(160, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ D, int K, int N, int mode, int passes) {
  DVD_PDL_ENTER();
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms must start on 1024-byte boundaries of the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_holder;
  const int nchunk = K / 64;
  // smem carve-up: per chunk: A_hi (16 KB) A_lo (16 KB) B_hi (N*128) B_lo (N*128)
  // K-major blocks: rows x 128 B; MN-major blocks: ceil(rows/64) atoms of 8 KB
  const uint32_t a_bytes = 128 * 128, b_bytes = (mode == 2) ? (uint32_t)((N + 63) / 64) * 8192u : (uint32_t)N * 128;   // mode 3: N/8 * 1024 = N * 128
  uint8_t* sA_hi = smem;
  uint8_t* sA_lo = sA_hi + nchunk * a_bytes;
  uint8_t* sB_hi = sA_lo + nchunk * a_bytes;
  uint8_t* sB_lo = sB_hi + nchunk * b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4) {
    tmem_alloc(&tmem_base_holder, 512);
    if (lane == 0) {
      mbar_init(&bar_done, 1);
      fence_mbar_init();
    }
  }
  // fill shared-memory operand images (generic proxy writes)
  if (mode == 2 || mode == 3) {
    // MN-major images (mode 2: 128B swizzle, mode 3: no-swizzle interleave): pairs of consecutive rows at the same k
    for (int i = threadIdx.x; i < 128 * K / 2; i += blockDim.x) {
      int row = (i % 64) * 2, k = i / 64;
      uint32_t hi, lo;
      split2(A[row * K + k], A[(row + 1) * K + k], hi, lo);
      uint32_t off = (k / 64) * a_bytes + (mode == 2 ? mn128_offset(row, k % 64) : il_offset(row, k % 64));
      *reinterpret_cast<uint32_t*>(sA_hi + off) = hi;
      *reinterpret_cast<uint32_t*>(sA_lo + off) = lo;
    }
    for (int i = threadIdx.x; i < N * K / 2; i += blockDim.x) {
      int row = (i % (N / 2)) * 2, k = i / (N / 2);
      uint32_t hi, lo;
      split2(B[row * K + k], B[(row + 1) * K + k], hi, lo);
      uint32_t off = (k / 64) * b_bytes + (mode == 2 ? mn128_offset(row, k % 64) : il_offset(row, k % 64));
      *reinterpret_cast<uint32_t*>(sB_hi + off) = hi;
      *reinterpret_cast<uint32_t*>(sB_lo + off) = lo;
    }
  } else {
  for (int i = threadIdx.x; i < 128 * K / 2; i += blockDim.x) {
    int row = i / (K / 2), k = (i % (K / 2)) * 2;
    uint32_t hi, lo;
    split2(A[row * K + k], A[row * K + k + 1], hi, lo);
    uint32_t off = (k / 64) * a_bytes + sw128_offset(row, k % 64);
    *reinterpret_cast<uint32_t*>(sA_hi + off) = hi;
    *reinterpret_cast<uint32_t*>(sA_lo + off) = lo;
  }
  for (int i = threadIdx.x; i < N * K / 2; i += blockDim.x) {
    int row = i / (K / 2), k = (i % (K / 2)) * 2;
    uint32_t hi, lo;
    split2(B[row * K + k], B[row * K + k + 1], hi, lo);
    uint32_t off = (k / 64) * b_bytes + sw128_offset(row, k % 64);
    *reinterpret_cast<uint32_t*>(sB_hi + off) = hi;
    *reinterpret_cast<uint32_t*>(sB_lo + off) = lo;
  }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_holder;
  const uint32_t tD = tmem;             // columns [0, N)
  const uint32_t tA_hi = tmem + 256;    // packed bf16: K/2 columns
  const uint32_t tA_lo = tmem + 384;

  if (mode == 1 && warp < 4) {
    // A operand into tensor memory: thread = row, 32-bit column j holds elements (2j, 2j+1)
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t h[16], l[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split2(A[row * K + 2 * (c0 + j)], A[row * K + 2 * (c0 + j) + 1], h[j], l[j]);
      tmem_st16(tA_hi + lane_base + c0, h);
      tmem_st16(tA_lo + lane_base + c0, l);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, mode >= 2, mode >= 2);
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
      // p = 0: hi*hi, 1: lo(A)*hi(B), 2: hi(A)*lo(B)
      const uint8_t* a_img = (p == 1) ? sA_lo : sA_hi;
      const uint32_t a_tm = (p == 1) ? tA_lo : tA_hi;
      const uint8_t* b_img = (p == 2) ? sB_lo : sB_hi;
      for (int kc = 0; kc < nchunk; ++kc) {
        for (int ks = 0; ks < 4; ++ks) {  // 16 bf16 = 32 bytes per MMA
          if (mode == 3) {
            uint64_t ad = make_sdesc_mn_interleave(smem_u32(a_img + kc * a_bytes) + ks * 256, 1024);
            uint64_t bd3 = make_sdesc_mn_interleave(smem_u32(b_img + kc * b_bytes) + ks * 256, 1024);
            umma_ss(tD, ad, bd3, idesc, acc);
            acc = 1;
            continue;
          }
          if (mode == 2) {
            uint64_t ad = make_sdesc_mn_sw128(smem_u32(a_img + kc * a_bytes) + ks * 2048, 8192);
            uint64_t bd2 = make_sdesc_mn_sw128(smem_u32(b_img + kc * b_bytes) + ks * 2048, 8192);
            umma_ss(tD, ad, bd2, idesc, acc);
            acc = 1;
            continue;
          }
          uint64_t bd = make_sdesc_k_sw128(smem_u32(b_img + kc * b_bytes) + ks * 32);
          if (mode == 0) {
            uint64_t ad = make_sdesc_k_sw128(smem_u32(a_img + kc * a_bytes) + ks * 32);
            umma_ss(tD, ad, bd, idesc, acc);
          } else {
            umma_ts(tD, a_tm + (kc * 4 + ks) * 8, bd, idesc, acc);
          }
          acc = 1;
        }
      }
    }
    umma_commit(&bar_done);
  }
  if (warp < 4) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tD + lane_base + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Probe for halo-resident convolution tiles: A is an image of R pixel rows x 32 fp32 channels in the K-major SWIZZLE_128B layout
// (row r at byte r * 128 from a 1024-aligned base, 16-byte chunk c stored at c ^ (r & 7) - exactly what a TMA box {32 ch, W, H}
// writes). The MMA reads 128 of those rows through a descriptor whose START is shifted by `shift` rows (not a multiple of 8) and
// whose 8-row groups are `sbo_rows` rows apart: D[m][n] = sum_k A[shift + (m / 8) * sbo_rows + m % 8][k] * B[n][k]  (TF32).
// bo_mode 1 additionally writes (start >> 7) & 7 into the descriptor's base-offset field (bits 49-51).
__global__ void __launch_bounds__(160, 1) halo_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ D, int R, int N, int shift, int sbo_rows, int bo_mode) {
  DVD_PDL_ENTER();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_holder;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)((R * 128 + 1023) / 1024) * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4) {
    tmem_alloc(&tmem_base_holder, 256);
    if (lane == 0) {
      mbar_init(&bar_done, 1);
      fence_mbar_init();
    }
  }
  for (int i = threadIdx.x; i < R * 8; i += blockDim.x) {
    const int r = i / 8, c = i % 8;
    *reinterpret_cast<float4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const float4*>(A + r * 32 + c * 4);
  }
  for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
    const int r = i / 8, c = i % 8;
    *reinterpret_cast<float4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const float4*>(B + r * 32 + c * 4);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_holder;
  if (warp == 4 && lane == 0) {
    // kind::tf32, M = 128, both operands K-major
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift * 128u + ks * 32;
      uint64_t ad = 0;
      ad |= (uint64_t)((a_addr >> 4) & 0x3FFF);
      ad |= (uint64_t)((((uint32_t)sbo_rows * 128u) >> 4) & 0x3FFF) << 32;
      ad |= (uint64_t)1 << 46;
      if (bo_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
      ad |= (uint64_t)2 << 61;
      const uint64_t bd = make_sdesc_k_sw128(smem_u32(sB) + ks * 32);
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(idesc), "r"(ks ? 1u : 0u)
          : "memory");
    }
    umma_commit(&bar_done);
  }
  if (warp < 4) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace dvd

// A [128,K] fp32 row-major, B [N,K] fp32 row-major, D [128,N] fp32; K in {64,128}, N in {16..256, %16}
extern "C" int dvd_selftest_umma(const float* A, const float* B, float* D, int K, int N, int mode, int passes,
                                 void* stream) {
  DVD_ARG_CHECK(A && B && D, "null pointer");
  DVD_ARG_CHECK(K == 64 || K == 128, "K must be 64 or 128");
  DVD_ARG_CHECK(N >= 16 && N <= 256 && N % 16 == 0, "N must be a multiple of 16 in [16,256]");
  DVD_ARG_CHECK(mode >= 0 && mode <= 3, "mode 0 (SS), 1 (TS), 2 (SS, MN-major swizzled) or 3 (SS, MN-major interleaved)");
  DVD_ARG_CHECK(passes == 1 || passes == 3, "passes 1 or 3");
  size_t smem = (size_t)(K / 64) * 2 * (128 * 128 + (size_t)((N + 63) / 64) * 8192) + 1024;
  DVD_CUDA_CALL(cudaFuncSetAttribute(dvd::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dvd::launch(dvd::umma_selftest_kernel, 1, 160, smem, (cudaStream_t)stream, A, B, D, K, N, mode, passes);
  DVD_CUDA_LAUNCH_CHECK("umma_selftest");
  return 0;
}

// A [R,32] fp32 (TF32 values), B [N,32], D [128,N]; see halo_selftest_kernel
extern "C" int dvd_selftest_halo(const float* A, const float* B, float* D, int R, int N, int shift, int sbo_rows, int bo_mode,
                                 void* stream) {
  DVD_ARG_CHECK(A && B && D, "null pointer");
  DVD_ARG_CHECK(N >= 16 && N <= 256 && N % 16 == 0 && R >= 128 && R <= 1024, "bad sizes");
  DVD_ARG_CHECK(shift >= 0 && sbo_rows >= 8 && shift + 15 * sbo_rows + 8 <= R, "rows out of range");
  size_t smem = (size_t)((R * 128 + 1023) / 1024) * 1024 + (size_t)N * 128 + 2048;
  DVD_CUDA_CALL(cudaFuncSetAttribute(dvd::halo_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dvd::launch(dvd::halo_selftest_kernel, 1, 160, smem, (cudaStream_t)stream, A, B, D, R, N, shift, sbo_rows, bo_mode);
  DVD_CUDA_LAUNCH_CHECK("halo_selftest");
  return 0;
}
