// Depth-net convolutions on the 5th-generation tensor cores: NHWC fp32 tensors, TF32 operands straight from the fp32
// data (tcgen05.mma.kind::tf32, SS mode), fp32 accumulators in tensor memory, fused per-channel affine (eval-mode
// BatchNorm or bias) + residual + ReLU epilogue.
//
// Replaces, for stride-1 / dense (groups = 1) 1x1 and 3x3 convolutions (SURVEY.md 8(a) rows D1, D1', D2):
//   torch.nn.Conv2d.forward -> cuDNN                       third_party/midas_blocks.py:53-68,121-168 (layerN_rn, RCU convs),
//                                                          third_party/MiDaS.py:188-195 (head), torchvision Bottleneck conv1/conv3
//   + BatchNorm2d (eval) + residual add + ReLU             torchvision Bottleneck.forward, midas_blocks.py:28-39
// and, with the tap-flipped / transposed weight image, the data gradient of the same convolutions.
//
// Implicit GEMM, no im2col buffer:  D[128 pixels, NT channels] += sum over taps t, 32-channel chunks c of
//     A_t,c [128 px x 32 ch]  (TMA box {32 ch, TW, TH, 1} of the NHWC input at spatial offset (dy-1, dx-1): the
//                              hardware zero-fills out-of-image rows/columns = the convolution's zero padding,
//                              and writes the canonical SWIZZLE_128B K-major image tcgen05.mma consumes)
//   x W_t,c [NT out-ch x 32 ch] (TMA box {32, NT, 1} of the packed weights [tap][Cout][Cin]).
// One persistent CTA per SM: warp 0 = TMA producer (4-stage ring of 16 KB + NT x 128 B), warp 1 = MMA issuer and
// tensor-memory owner (two accumulators of NT columns, so the epilogue of tile i overlaps the MMAs of tile i+1),
// warps 2-5 = epilogue (tcgen05.ld 32x32b -> affine / residual / ReLU -> 128-bit stores, one pixel row per thread).
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace dvd {
namespace {

using namespace tc;

constexpr int kConvStages = 4;
constexpr int kConvThreads = 320;          // TMA producer, MMA issuer, 4 epilogue warps, 4 operand-rounding warps
constexpr int kABytes = 128 * 128;          // 128 pixels x 32 fp32 channels
constexpr int kMaxNT = 256;
constexpr int kStageBytesConv = kABytes + kMaxNT * 128;
constexpr size_t kConvSmem = 1024 + (size_t)kConvStages * kStageBytesConv + 512 + 2 * 2 * kMaxNT * 4;

struct ConvParams {
  const float* bias;    // [Cout] conv bias or null
  const float* gamma;   // eval-mode BatchNorm (weight, bias, running_mean, running_var) or all null
  const float* beta;
  const float* mean;
  const float* var;
  float eps;
  const float* res;     // NHWC like y, or null
  float* y;
  int N, H, W, Cin, Cout;
  int taps;             // 1 or 9
  int TW, TH;           // pixel tile = TH x TW = 128
  int tiles_w, tiles_h; // spatial tiles per image
  int NT;               // output channels per tile
  int relu;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs (fp32 words, low 13 mantissa bits ignored), K = 8 per instruction
__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (2u << 7)                    // a_format = TF32
         | (2u << 10)                   // b_format = TF32
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

struct Tile {
  int n0;           // first output channel
  int img, h0, w0;  // image index and top-left pixel of the spatial tile
};
__device__ __forceinline__ Tile decode_tile(const ConvParams& P, int tile, int m_tiles) {
  Tile t;
  const int nt = tile / m_tiles, m = tile - nt * m_tiles;
  t.n0 = nt * P.NT;
  const int per_img = P.tiles_w * P.tiles_h;
  t.img = m / per_img;
  const int r = m - t.img * per_img;
  const int th = r / P.tiles_w;
  t.h0 = th * P.TH;
  t.w0 = (r - th * P.tiles_w) * P.TW;
  return t;
}

// affine / residual / ReLU on NC consecutive output channels of one pixel, 128-bit stores
// `aff` = this tile's folded per-channel scale [256] and shift [256] in shared memory (null: identity)
template <int NC>
__device__ __forceinline__ void epilogue_store(const ConvParams& P, const uint32_t (&r)[NC], const float* aff, float* yrow,
                                               const float* rrow) {
#pragma unroll
  for (int j = 0; j < NC; j += 4) {
    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    if (aff) {
      const float4 sc = *reinterpret_cast<const float4*>(aff + j), sh = *reinterpret_cast<const float4*>(aff + kMaxNT + j);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    if (rrow) {
      const float4 rr = *reinterpret_cast<const float4*>(rrow + j);
      v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
    }
    if (P.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(yrow + j) = v;
  }
}

__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapW,
                                                                  const __grid_constant__ ConvParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage[kConvStages];
  for (int i = 0; i < kConvStages; ++i) stage[i] = base + (size_t)i * kStageBytesConv;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)kConvStages * kStageBytesConv);
  uint64_t* full = bars;                       // [kConvStages]
  uint64_t* empty = bars + kConvStages;        // [kConvStages]
  uint64_t* ready = bars + 2 * kConvStages;    // [kConvStages]  A tile rounded to TF32
  uint64_t* acc_full = bars + 3 * kConvStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* aff_mem = reinterpret_cast<float*>(base + (size_t)kConvStages * kStageBytesConv + 512);   // [2 buffers][scale 256 | shift 256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
  } else if (warp == 0 && lane == 0) {
    for (int i = 0; i < kConvStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&ready[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  const int m_tiles = P.N * P.tiles_w * P.tiles_h;
  const int n_tiles = P.Cout / P.NT;
  const int ntiles = m_tiles * n_tiles;
  const int kchunks = P.Cin / 32;
  const int ksteps = P.taps * kchunks;
  const uint32_t stage_tx = (uint32_t)kABytes + (uint32_t)P.NT * 128u;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const Tile T = decode_tile(P, tile, m_tiles);
        for (int t = 0; t < P.taps; ++t) {
          const int dy = P.taps == 9 ? t / 3 - 1 : 0, dx = P.taps == 9 ? t % 3 - 1 : 0;
          for (int kc = 0; kc < kchunks; ++kc, ++it) {
            const uint32_t s = it % kConvStages, ph = (it / kConvStages) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full[s], stage_tx);
            tma_load_4d(stage[s], &mapA, kc * 32, T.w0 + dx, T.h0 + dy, T.img, &full[s]);
            tma_load_3d(stage[s] + kABytes, &mapW, kc * 32, T.n0, t, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(128, P.NT);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1u, aph = (lt >> 1) & 1u;
        mbar_wait(&acc_empty[buf], aph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem + buf * (uint32_t)kMaxNT;
        for (int k = 0; k < ksteps; ++k, ++it) {
          const uint32_t s = it % kConvStages, ph = (it / kConvStages) & 1u;
          mbar_wait(&ready[s], ph);                // TMA landed AND the activation tile has been rounded
          tc_fence_after();
          const uint32_t sa = smem_u32(stage[s]), sb = sa + kABytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss_tf32(d, make_sdesc_k_sw128(sa + ks * 32), make_sdesc_k_sw128(sb + ks * 32), idesc, (k | ks) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 6) {
    // ===== operand rounding: fp32 -> TF32 with round-to-nearest, in place in shared memory =====
    // tcgen05.mma.kind::tf32 ignores the low 13 mantissa bits of its operands, i.e. truncates: measured slope -7.0e-4 per
    // layer against fp64 (tools/debug_conv_bias.py), where cuDNN's TF32 path (RN conversion) has -3e-6. The weights are
    // rounded once when they are packed; the activation tile is rounded here, after TMA and before the MMAs read it.
    const int t = threadIdx.x - 6 * 32;          // 0..127
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int k = 0; k < ksteps; ++k, ++it) {
        const uint32_t s = it % kConvStages, ph = (it / kConvStages) & 1u;
        mbar_wait(&full[s], ph);
        uint4* a = reinterpret_cast<uint4*>(stage[s]);
#pragma unroll
        for (int i = 0; i < kABytes / 16 / 128; ++i) {
          uint4 v = a[t + i * 128];
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.x) : "f"(__uint_as_float(v.x)));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.y) : "f"(__uint_as_float(v.y)));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.z) : "f"(__uint_as_float(v.z)));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.w) : "f"(__uint_as_float(v.w)));
          a[t + i * 128] = v;
        }
        fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core's async proxy
        mbar_arrive(&ready[s]);
      }
    }
  } else {
    // ===== epilogue: one pixel (accumulator row) per thread =====
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1u, aph = (lt >> 1) & 1u;
      const Tile T = decode_tile(P, tile, m_tiles);
      const int h = T.h0 + row / P.TW, w = T.w0 + row % P.TW;
      const bool valid = h < P.H && w < P.W;
      const size_t pix = ((size_t)T.img * P.H + h) * P.W + w;
      float* yrow = P.y + pix * P.Cout + T.n0;
      const float* rrow = P.res ? P.res + pix * P.Cout + T.n0 : nullptr;
      // fold BatchNorm / bias of this tile's channels once: y = acc * scale + shift
      const bool has_aff = P.gamma != nullptr || P.bias != nullptr;
      float* aff = has_aff ? aff_mem + buf * (2 * kMaxNT) : nullptr;
      if (has_aff) {
        for (int i = row; i < P.NT; i += 128) {
          const int c = T.n0 + i;
          float sc = 1.f, sh = 0.f;
          if (P.gamma) {
            sc = __ldg(P.gamma + c) * rsqrtf(__ldg(P.var + c) + P.eps);
            sh = __ldg(P.beta + c) - __ldg(P.mean + c) * sc;
          }
          if (P.bias) sh = fmaf(__ldg(P.bias + c), sc, sh);
          aff[i] = sc;
          aff[kMaxNT + i] = sh;
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");      // the four epilogue warps
      }
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      const uint32_t d = tmem + buf * (uint32_t)kMaxNT + lane_base;
      if (P.NT % 32 == 0) {
        for (int c0 = 0; c0 < P.NT; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(d + c0, r);
          tmem_ld_wait();
          if (valid) epilogue_store<32>(P, r, aff ? aff + c0 : nullptr, yrow + c0, rrow ? rrow + c0 : nullptr);
        }
      } else {
        for (int c0 = 0; c0 < P.NT; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(d + c0, r);
          tmem_ld_wait();
          if (valid) epilogue_store<16>(P, r, aff ? aff + c0 : nullptr, yrow + c0, rrow ? rrow + c0 : nullptr);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// weight[co, ci, ky, kx] (arbitrary strides, in elements) -> tap-major image [k*k][rows][cols] rounded to TF32 (RN):
// forward image rows = co, cols = ci; data-gradient image rows = ci, cols = co with the taps rotated by 180 degrees
__global__ void __launch_bounds__(256) conv_pack_weight_kernel(const float* __restrict__ w, long s_co, long s_ci, long s_ky,
                                                               long s_kx, float* __restrict__ out, int Cout, int Cin, int k,
                                                               int dgrad) {
  const int rows = dgrad ? Cin : Cout, cols = dgrad ? Cout : Cin;
  const long n = (long)k * k * rows * cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long q = i / cols;
    const int r = (int)(q % rows), t = (int)(q / rows);
    int ky = t / k, kx = t - ky * k;
    int co = r, ci = c;
    if (dgrad) { co = c; ci = r; ky = k - 1 - ky; kx = k - 1 - kx; }
    const float v = w[co * s_co + ci * s_ci + ky * s_ky + kx * s_kx];
    uint32_t o;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(v));
    out[i] = __uint_as_float(o);
  }
}

// =============================================================================================================
// Weight gradient:  dW[t][co][ci] += sum over pixels  dY[px, co] * X[px + offset(t), ci]
// GEMM with K = pixels: both operands are "MN-major" (their M / N index - the channel - is the contiguous one), which
// tcgen05 accepts for TF32. A 64-pixel TMA box {32 ch, TW, TH, 1} lands as 64 rows of 128 bytes, SWIZZLE_128B = one
// MN-atom column of the canonical MN-major image (8-row K groups 1024 B apart = SBO, next 32 channels one box = 8 KB
// further = LBO); one MMA consumes 8 pixels (K = 8), i.e. +1024 B on both descriptors. The spatial shift of the tap is
// again just the TMA coordinate of the X box (zero fill = padding). Split-K over pixel tiles across CTAs, partial sums
// leave through fp32 reductions (vector red.global.add.v4 when the destination's Cin stride is 1).
// Operands are NOT re-rounded here (both are activations): the tensor core truncates them to TF32, a uniform -7e-4
// scale on dW (see the forward kernel) that Adam's normalisation cancels; tolerance of the test is 2e-3.
constexpr int kWgStages = 2;
constexpr int kWgPx = 64;                         // pixels (K) per stage
constexpr int kWgBox = kWgPx * 128;               // bytes of one {32 ch, 64 px} box
constexpr int kWgStageBytes = (4 + 8) * kWgBox;   // A: 128 out-channels, B: up to 256 in-channels
constexpr size_t kWgSmem = 1024 + (size_t)kWgStages * kWgStageBytes + 256;

struct WgradParams {
  float* dw;                       // destination, element strides below
  long s_co, s_ci, s_ky, s_kx;
  int N, H, W, Cin, Cout, taps;
  int TW, TH, tiles_w, tiles_h;    // 64-pixel tiles
  int NT;                          // in-channels per output tile
  int ksplit;                      // CTAs sharing one output tile
};

// MN-major operand of 32-bit elements: SWIZZLE_128B_BASE32B image (UMMA LayoutType 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
// 128-byte rows (32 channels) per K index, 32-byte chunk c of row r stored at chunk (c ^ (r & 3)), K atoms of 4 rows
// 512 B apart (SBO), next 32 channels `lbo_bytes` further (LBO).
__device__ __forceinline__ uint64_t make_sdesc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
  return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);   // A and B MN-major
}

__global__ void __launch_bounds__(192, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapDY,
                                                             const __grid_constant__ CUtensorMap mapX,
                                                             const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)kWgStages * kWgStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    tmem_alloc(tmem_holder, 256);
  } else if (warp == 0 && lane == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  // this CTA: output tile (tap, 128 out-channels, NT in-channels) and a contiguous range of pixel tiles
  const int out_tile = blockIdx.x / P.ksplit, part = blockIdx.x - out_tile * P.ksplit;
  const int n_ci = P.Cin / P.NT, n_co = P.Cout / 128;
  const int t = out_tile / (n_co * n_ci);
  const int rem = out_tile - t * (n_co * n_ci);
  const int m0 = (rem / n_ci) * 128, n0 = (rem % n_ci) * P.NT;
  const int dy = P.taps == 9 ? t / 3 - 1 : 0, dx = P.taps == 9 ? t % 3 - 1 : 0;
  const int px_tiles = P.N * P.tiles_h * P.tiles_w;
  const int per = (px_tiles + P.ksplit - 1) / P.ksplit;
  const int kt0 = part * per, kt1 = min(px_tiles, kt0 + per);
  const int nb = P.NT / 32;
  const uint32_t stage_tx = (uint32_t)(4 + nb) * kWgBox;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int kt = kt0; kt < kt1; ++kt, ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1u;
        const int img = kt / (P.tiles_h * P.tiles_w), r = kt - img * (P.tiles_h * P.tiles_w);
        const int h0 = (r / P.tiles_w) * P.TH, w0 = (r % P.tiles_w) * P.TW;
        uint8_t* st = base + (size_t)s * kWgStageBytes;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], stage_tx);
        for (int j = 0; j < 4; ++j) tma_load_4d(st + j * kWgBox, &mapDY, m0 + 32 * j, w0, h0, img, &full[s]);
        for (int j = 0; j < nb; ++j) tma_load_4d(st + (4 + j) * kWgBox, &mapX, n0 + 32 * j, w0 + dx, h0 + dy, img, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(128, P.NT);
      uint32_t it = 0;
      for (int kt = kt0; kt < kt1; ++kt, ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(base + (size_t)s * kWgStageBytes), sb = sa + 4 * kWgBox;
#pragma unroll
        for (int ks = 0; ks < kWgPx / 8; ++ks)
          umma_ss_tf32(tmem, make_sdesc_mn_sw128_32b(sa + ks * 1024, kWgBox), make_sdesc_mn_sw128_32b(sb + ks * 1024, kWgBox), idesc,
                       (it | ks) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else if (kt1 > kt0) {
    // epilogue: accumulator row = out-channel, columns = in-channels of this tile
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    const uint32_t d = tmem + ((uint32_t)(q * 32) << 16);
    float* dst = P.dw + (long)co * P.s_co + (long)(dy + (P.taps == 9 ? 1 : 0)) * P.s_ky + (long)(dx + (P.taps == 9 ? 1 : 0)) * P.s_kx;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < P.NT; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(d + c0, r);
      tmem_ld_wait();
      if (P.s_ci == 1 && ((reinterpret_cast<uintptr_t>(dst + n0 + c0) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + n0 + c0 + j), "f"(__uint_as_float(r[j])),
                       "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(dst + (long)(n0 + c0 + j) * P.s_ci, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

int make_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  auto fn = encode_fn();
  DVD_ARG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DVD_ARG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

}  // namespace
}  // namespace dvd

using namespace dvd;

extern "C" int dvd_conv_nhwc_fwd(const float* x, const float* w_tkc, const float* bias, const float* bn_gamma, const float* bn_beta,
                                 const float* bn_mean, const float* bn_var, float bn_eps, const float* res, float* y, int N, int H,
                                 int W, int Cin, int Cout, int ksize, int relu, void* stream) {
  DVD_ARG_CHECK(x && w_tkc && y, "null pointer");
  DVD_ARG_CHECK(N >= 1 && H >= 1 && W >= 1, "bad shape N=%d H=%d W=%d", N, H, W);
  DVD_ARG_CHECK(ksize == 1 || ksize == 3, "ksize must be 1 or 3 (stride 1, dense)");
  if (Cin % 32 != 0 || Cout % 16 != 0) { set_error("dvd_conv_nhwc_fwd: needs Cin %% 32 == 0 and Cout %% 16 == 0 (Cin=%d Cout=%d)", Cin, Cout); return -2; }
  DVD_ARG_CHECK(aligned16(x) && aligned16(w_tkc) && aligned16(y) && (!res || aligned16(res)), "tensors must be 16-byte aligned");
  DVD_ARG_CHECK((bn_gamma != nullptr) == (bn_beta != nullptr) && (bn_gamma != nullptr) == (bn_mean != nullptr) &&
                    (bn_gamma != nullptr) == (bn_var != nullptr),
                "BatchNorm needs all of gamma, beta, mean, var (or none)");
  ConvParams P{};
  P.bias = bias; P.gamma = bn_gamma; P.beta = bn_beta; P.mean = bn_mean; P.var = bn_var; P.eps = bn_eps; P.res = res; P.y = y;
  P.N = N; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.taps = ksize * ksize; P.relu = relu;
  P.NT = Cout >= kMaxNT ? kMaxNT : Cout;
  if (Cout % P.NT != 0) { set_error("dvd_conv_nhwc_fwd: Cout=%d is not a multiple of the %d-channel tile", Cout, P.NT); return -2; }
  CUtensorMap mapA, mapW;
  if (ksize == 1) {
    // a 1x1 convolution is a plain GEMM over all N*H*W pixels: one "image" of P x 1
    const long Pn = (long)N * H * W;
    DVD_ARG_CHECK(Pn < (1L << 31), "too many pixels");
    P.N = 1; P.H = 1; P.W = (int)Pn; P.TW = 128; P.TH = 1;
  } else {
    // 128-pixel tile TH x TW with the least padding waste
    long best = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
      const int th = 128 / tw;
      const long padded = (long)((W + tw - 1) / tw * tw) * ((H + th - 1) / th * th);
      if (best < 0 || padded < best) { best = padded; P.TW = tw; P.TH = th; }
    }
  }
  P.tiles_w = (P.W + P.TW - 1) / P.TW;
  P.tiles_h = (P.H + P.TH - 1) / P.TH;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)P.W * Cin * 4, (cuuint64_t)P.H * P.W * Cin * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)P.TW, (cuuint32_t)P.TH, 1};
    if (int e = make_map(&mapA, x, 4, dims, strides, box)) return e;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)P.taps};
    const cuuint64_t strides[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)Cout * Cin * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)P.NT, 1};
    if (int e = make_map(&mapW, w_tkc, 3, dims, strides, box)) return e;
  }
  static const bool attr_ok =
      cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem) == cudaSuccess;
  DVD_ARG_CHECK(attr_ok, "cudaFuncSetAttribute(conv_tc_kernel) failed");
  const long ntiles = (long)P.N * P.tiles_w * P.tiles_h * (Cout / P.NT);
  int grid = num_sms();
  if (ntiles < grid) grid = (int)ntiles;
  conv_tc_kernel<<<grid, kConvThreads, kConvSmem, (cudaStream_t)stream>>>(mapA, mapW, P);
  DVD_CUDA_LAUNCH_CHECK("conv_tc_kernel");
  return 0;
}

extern "C" int dvd_conv_pack_weight(const float* weight, long stride_co, long stride_ci, long stride_ky, long stride_kx,
                                    float* w_tkc, int Cout, int Cin, int ksize, int dgrad, void* stream) {
  DVD_ARG_CHECK(weight && w_tkc, "null pointer");
  DVD_ARG_CHECK(Cout >= 1 && Cin >= 1 && (ksize == 1 || ksize == 3), "bad weight shape");
  const long n = (long)ksize * ksize * Cout * Cin;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  conv_pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(weight, stride_co, stride_ci, stride_ky, stride_kx, w_tkc,
                                                                  Cout, Cin, ksize, dgrad);
  DVD_CUDA_LAUNCH_CHECK("conv_pack_weight_kernel");
  return 0;
}

extern "C" int dvd_conv_nhwc_wgrad(const float* x, const float* gy, float* dweight, long stride_co, long stride_ci, long stride_ky,
                                   long stride_kx, int N, int H, int W, int Cin, int Cout, int ksize, void* stream) {
  DVD_ARG_CHECK(x && gy && dweight, "null pointer");
  DVD_ARG_CHECK(N >= 1 && H >= 1 && W >= 1, "bad shape N=%d H=%d W=%d", N, H, W);
  DVD_ARG_CHECK(ksize == 1 || ksize == 3, "ksize must be 1 or 3 (stride 1, dense)");
  if (Cin % 32 != 0 || Cout % 128 != 0 || (Cin > 256 && Cin % 256 != 0)) {
    set_error("dvd_conv_nhwc_wgrad: needs Cout %% 128 == 0 and Cin %% 32 == 0 (<= 256 or a multiple of 256); Cin=%d Cout=%d", Cin, Cout);
    return -2;
  }
  DVD_ARG_CHECK(aligned16(x) && aligned16(gy), "tensors must be 16-byte aligned");
  WgradParams P{};
  P.dw = dweight; P.s_co = stride_co; P.s_ci = stride_ci; P.s_ky = stride_ky; P.s_kx = stride_kx;
  P.N = N; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.taps = ksize * ksize;
  P.NT = Cin >= 256 ? 256 : Cin;
  if (ksize == 1) {
    const long Pn = (long)N * H * W;
    DVD_ARG_CHECK(Pn < (1L << 31), "too many pixels");
    P.N = 1; P.H = 1; P.W = (int)Pn; P.TW = kWgPx; P.TH = 1;
  } else {
    long best = -1;
    for (int tw = kWgPx; tw >= 8; tw >>= 1) {
      const int th = kWgPx / tw;
      const long padded = (long)((W + tw - 1) / tw * tw) * ((H + th - 1) / th * th);
      if (best < 0 || padded < best) { best = padded; P.TW = tw; P.TH = th; }
    }
  }
  P.tiles_w = (P.W + P.TW - 1) / P.TW;
  P.tiles_h = (P.H + P.TH - 1) / P.TH;
  const int out_tiles = P.taps * (Cout / 128) * (Cin / P.NT);
  const int px_tiles = P.N * P.tiles_h * P.tiles_w;
  int ksplit = (num_sms() + out_tiles - 1) / out_tiles;
  if (ksplit > px_tiles) ksplit = px_tiles;
  if (ksplit < 1) ksplit = 1;
  P.ksplit = ksplit;
  CUtensorMap mapDY, mapX;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)P.W * Cout * 4, (cuuint64_t)P.H * P.W * Cout * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)P.TW, (cuuint32_t)P.TH, 1};
    if (int e = make_map(&mapDY, gy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)P.N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)P.W * Cin * 4, (cuuint64_t)P.H * P.W * Cin * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)P.TW, (cuuint32_t)P.TH, 1};
    if (int e = make_map(&mapX, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  }
  static const bool attr_ok =
      cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem) == cudaSuccess;
  DVD_ARG_CHECK(attr_ok, "cudaFuncSetAttribute(conv_wgrad_tc_kernel) failed");
  conv_wgrad_tc_kernel<<<out_tiles * ksplit, 192, kWgSmem, (cudaStream_t)stream>>>(mapDY, mapX, P);
  DVD_CUDA_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  return 0;
}
