// Depth-net convolutions (every class of the MiDaS / ResNeXt101-32x8d stack) on the 5th-generation tensor cores.
//
// One implicit-GEMM kernel family on NHWC fp32 tensors whose values are already rounded to TF32 (round-to-nearest) by
// the kernel that produced them ("rounded-operand contract"): tcgen05.mma.kind::tf32, SS mode, fp32
// accumulators in tensor memory. Replaces cuDNN behind torch.nn.Conv2d / convolution_backward for
//   dense 1x1 / 3x3 stride 1     third_party/midas_blocks.py:53-68,121-168; third_party/MiDaS.py:188-195; torchvision Bottleneck conv1/conv3
//   1x1 / 3x3 stride 2           torchvision Bottleneck.downsample, conv2 of the first block of layer2-4
//   grouped 3x3 (32 groups)      torchvision Bottleneck.conv2 (ResNeXt), as block-diagonal 64-channel GEMM blocks
//   the data gradient of each    same kernel: transposed (BatchNorm-scaled) weight image, a tap table instead of a fixed
//                                3x3 stencil, stride-2 gradients as 4 sub-pixel phases with a strided store
//   the weight gradient of each  conv_wgrad_kernel below (K = pixels, both operands MN-major)
// with the elementwise neighbours folded into the epilogue:  y = round_tf32(mask * relu(acc * scale + shift + res + res2))
// (eval-mode BatchNorm / bias, residual adds, ReLU forward; residual-gradient add and ReLU mask backward).
//
//   D[128 pixels, NT channels] += sum over taps t, 32-channel chunks c of
//     A_t,c [128 px x 32 ch]  TMA box {32 ch, TW, TH, 1} of the NHWC input at stride * tile origin + (dy_t, dx_t), element
//                             stride = conv stride; out-of-image elements are zero-filled by the hardware (= padding);
//                             lands as the canonical SWIZZLE_128B K-major image
//   x W_t,c [NT x 32]         TMA box {32, NT, 1} of the packed weight image [tap][rows][cols]
// One persistent CTA per SM: warp 0 = TMA producer (ring of kStages stages), warp 1 = MMA issuer and tensor-memory owner
// (two accumulators: the epilogue of tile i overlaps the MMAs of tile i+1), warps 2-5 = epilogue: tcgen05.ld -> registers
// -> swizzled shared-memory staging -> one TMA store per warp and 32-channel block (full 128-byte lines; ragged tiles are
// clipped by the hardware).
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace dvd {
namespace {

using namespace tc;

constexpr int kStages = 4;
constexpr int kThreads = 320;               // TMA producer, MMA issuer, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kABytes = 128 * 128;          // 128 pixels x 32 fp32 channels
constexpr int kMaxNT = 256;
constexpr int kStageBytes = kABytes + kMaxNT * 128;
constexpr int kStgBytes = 32 * 128;         // staging: 32 pixel rows x 32 channels
constexpr int kOffStg = kStages * kStageBytes;              // 8 warps x 1 buffer
constexpr int kOffAff = kOffStg + 8 * kStgBytes;            // scale[256] | shift[256]
constexpr int kOffBar = kOffAff + 2 * kMaxNT * 4;
constexpr size_t kSmem = kOffBar + 256;
constexpr int kWsFlagWords = 256;           // stream-K exchange area: flag words in front of the partial tiles

struct ConvParams {
  dvd_conv_desc d;
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  const float* res;
  const float* res2;
  const float* mask;
  float* y;
  int TW, TH, tiles_w, tiles_h;   // pixel tile = TH x TW = 128 over the (OH, OW) grid
  int NT;                         // output channels per tile
  int kchunks;                    // 32-channel chunks per tap
  int tma_store;                  // epilogue through shared memory + TMA store (identity output mapping, NT % 32 == 0)
  int sbw, sbh;                   // per-warp store box: sbw x sbh = 32 pixels
  int csize;                      // thread-block cluster size (1, 2, 4): the CTAs of a cluster work on consecutive pixel tiles of
                                  // the same channel tile and share the weight stream (each loads NT/csize rows, multicast to all)
  int pair;                       // csize == 2 as ONE tcgen05 CTA pair (cta_group::2): M = 256 pixels per MMA, each CTA keeps its own
                                  // 128-pixel A tile and HALF of the weight tile in shared memory (the tensor core reads both halves),
                                  // so the per-SM shared-memory traffic of the weight operand - the measured bound - halves
  int halo;                       // halo-resident activation tiles (stride-1 stencils): per 32-channel chunk ONE TMA box of the 16 x 8 pixel
                                  // tile plus its stencil halo instead of one box per tap; a tap is a row offset of the MMA's A descriptor
                                  // (8-pixel tile rows = the descriptor's 8-row groups, pitch HW rows apart). The A traffic of a k x k layer
                                  // drops by k*k * 128 / (HW * HH); weights stream through their own ring, one box per (chunk, tap)
  int HW, HH, dy0, dx0;           // halo box in pixels and the offset of its origin from the tile origin (= the smallest tap offsets)
  int a_box, a_stage, nA, nW;     // bytes of a halo box / of its 1024-aligned stage, stages of the two rings
  int sk;                         // stream-K: the (tile, K-step) work list is cut into one contiguous range per cluster instead of
                                  // whole tiles round-robin; a tile cut by a range boundary is finished by the cluster that owns its
                                  // FIRST K-steps (it reaches that tile last), the others leave raw partial accumulators in `ws`
  float* ws;                      // int flags[256] (zero when idle) | [gridDim.x][128][NT] partial tiles
};

// One segment of work: K-steps [k0, k1) of cluster tile `tile`. Data-parallel mode: whole tiles, round-robin over the clusters.
// Stream-K mode: the cluster's contiguous range of the tile-major (tile, K-step) list; range boundaries that fall within 1/8 of a
// tile boundary snap to it (a sliver is not worth a partial-tile exchange).
struct SegIter {
  long u, u1;
  int tile, ntiles, step, ksteps, sk;
  __host__ __device__ static long bound(long c, long n_clusters, long total, int ksteps) {
    if (c >= n_clusters) return total;
    long u = c * total / n_clusters;
    const int r = (int)(u % ksteps);
    if (r * 8 < ksteps) u -= r;
    else if ((ksteps - r) * 8 < ksteps) u += ksteps - r;
    return u;
  }
  __device__ SegIter(int sk_, int cluster_id, int n_clusters, int ntiles_, int ksteps_) : ntiles(ntiles_), step(n_clusters), ksteps(ksteps_), sk(sk_) {
    tile = cluster_id;
    if (sk) {
      const long total = (long)ntiles * ksteps;
      u = bound(cluster_id, n_clusters, total, ksteps);
      u1 = bound(cluster_id + 1, n_clusters, total, ksteps);
    } else {
      u = u1 = 0;
    }
  }
  __device__ bool next(int& t, int& k0, int& k1) {
    if (!sk) {
      if (tile >= ntiles) return false;
      t = tile; k0 = 0; k1 = ksteps;
      tile += step;
      return true;
    }
    if (u >= u1) return false;
    t = (int)(u / ksteps);
    k0 = (int)(u - (long)t * ksteps);
    const long left = u1 - u;
    k1 = (left < (long)(ksteps - k0)) ? k0 + (int)left : ksteps;
    u += k1 - k0;
    return true;
  }
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
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
// all MMAs issued so far by this thread complete -> one arrival on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// cta_group::2 TMA loads: data lands in THIS CTA's shared memory, the bytes are counted on the pair leader's mbarrier
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem of both CTAs, 256 x N] (+)= A[128 rows in each CTA] * B[N/2 rows in each CTA]^T, issued by the pair leader
__device__ __forceinline__ void umma_ss_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs (fp32 words whose low 13 mantissa bits are zero), K = 8 per instruction
__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major SWIZZLE_128B descriptor whose 8-row groups are `sbo_bytes` apart (halo boxes: the pitch of a tile row); the start may
// be any 128-byte row of the image: the swizzle is a function of the absolute shared-memory address (profiles/r3_probe_halo_*)
__device__ __forceinline__ uint64_t make_sdesc_k_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t o;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(o) : "f"(v));
  return __uint_as_float(o);
}

struct Tile {
  int n0, img, h0, w0;
  bool valid;       // false: padding tile of an incomplete cluster (loads are zero-filled, nothing is stored)
};
// cluster tile ct = (channel tile, group of csize consecutive pixel tiles); this CTA takes pixel tile group * csize + rank
__device__ __forceinline__ Tile decode_tile(const ConvParams& P, int ct, int m_tiles, int mgroups, int crank) {
  Tile t;
  const int nt = ct / mgroups, m = (ct - nt * mgroups) * P.csize + crank;
  t.n0 = nt * P.NT;
  t.valid = m < m_tiles;
  const int per_img = P.tiles_w * P.tiles_h;
  t.img = t.valid ? m / per_img : P.d.N;          // image index N is out of bounds for every TMA map: zero fill / no store
  const int r = t.valid ? m - t.img * per_img : 0;
  const int th = r / P.tiles_w;
  t.h0 = th * P.TH;
  t.w0 = (r - th * P.tiles_w) * P.TW;
  return t;
}

// acc -> mask * relu(acc * scale + shift + res + res2), optionally rounded to TF32; NC consecutive channels of one pixel.
// res / mask arrive in registers (the epilogue loads them one block ahead: they are what it would otherwise wait for).
template <int NC>
__device__ __forceinline__ void epilogue_math(const ConvParams& P, uint32_t (&r)[NC], const float* aff, const float4* resv,
                                              const float* r2row, const float4* maskv) {
#pragma unroll
  for (int j = 0; j < NC; j += 4) {
    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    if (aff) {
      const float4 sc = *reinterpret_cast<const float4*>(aff + j), sh = *reinterpret_cast<const float4*>(aff + kMaxNT + j);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    if (resv) {
      const float4 a = resv[j / 4];
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (r2row) {
      const float4 a = *reinterpret_cast<const float4*>(r2row + j);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (P.d.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (maskv) {
      const float4 m = maskv[j / 4];
      v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    if (P.d.round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    r[j] = __float_as_uint(v.x); r[j + 1] = __float_as_uint(v.y); r[j + 2] = __float_as_uint(v.z); r[j + 3] = __float_as_uint(v.w);
  }
}
template <int NV>
__device__ __forceinline__ void load_row(const float* row, float4 (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(row) + j);
}

// kPair: compiled with the cta_group::2 instructions (such a kernel can only be launched in clusters of two: the plain variant
// must not contain them)
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1) conv2d_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                               const __grid_constant__ CUtensorMap mapW,
                                                               const __grid_constant__ CUtensorMap mapY,
                                                               const __grid_constant__ ConvParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* full = bars;                       // [8]  stage ring (halo mode: the weight ring)
  uint64_t* empty = bars + 8;                  // [8]
  uint64_t* fullA = bars + 16;                 // [4]  halo mode: ring of halo boxes
  uint64_t* emptyA = bars + 20;                // [4]
  uint64_t* acc_full = bars + 24;              // [2]
  uint64_t* acc_empty = bars + 26;             // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 28);
  float* aff_mem = reinterpret_cast<float*>(smem + kOffAff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("dvd_b200: conv2d_tc_kernel: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  // halo mode: byte offset of every tap into the halo box, and its weight image index
  // (the last KB of the ring area: the host leaves it free in halo mode)
  uint32_t* tap_off = reinterpret_cast<uint32_t*>(smem + (size_t)kStages * kStageBytes - 1024);
  int* tap_wt = reinterpret_cast<int*>(tap_off + DVD_CONV_MAX_TAPS);
  if (P.halo) {
    for (int t = threadIdx.x; t < P.d.ntaps; t += blockDim.x) {
      tap_off[t] = (uint32_t)((P.d.dy[t] - P.dy0) * P.HW + (P.d.dx[t] - P.dx0)) * 128u;
      tap_wt[t] = P.d.wt[t];
    }
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_2sm(tmem_holder, 512);
    else tmem_alloc(tmem_holder, 512);
  } else if (warp == 0 && lane == 0) {
    // pair mode: one commit of the leader's MMA thread (multicast) frees a stage in both CTAs; the leader's accumulator is
    // drained by the epilogue threads of BOTH CTAs
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kPair ? 1u : (uint32_t)P.csize); }
    for (int i = 0; i < 4; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kPair ? 512u : 256u); }
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    if (P.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapY) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (P.csize > 1) cluster_sync_all();         // every CTA's barriers exist before a peer's multicast can reach them
  const uint32_t tmem = *tmem_holder;
  DVD_PDL_ENTER();                             // set-up done: now wait for the producer of the operands

  const int crank = P.csize > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / P.csize, n_clusters = gridDim.x / P.csize;
  const uint16_t cmask = (uint16_t)((1u << P.csize) - 1u);
  const int m_tiles = P.d.N * P.tiles_w * P.tiles_h;
  const int mgroups = (m_tiles + P.csize - 1) / P.csize;
  const int n_tiles = (P.d.Cout + P.NT - 1) / P.NT;   // the last channel tile may be ragged (whole 32-channel blocks): the weight
                                                      // rows beyond Cout are zero-filled by the TMA unit, the epilogue skips them
  const int ntiles = mgroups * n_tiles;           // cluster tiles
  const int ksteps = P.d.ntaps * P.kchunks;
  const uint32_t stage_tx = (uint32_t)kABytes + (uint32_t)P.NT * 128u;
  const int wslice = P.NT / P.csize;              // weight rows this CTA fetches (and multicasts) per stage

  // halo mode: [nA halo stages | nW weight stages] in the ring area
  uint8_t* const smemW = smem + (size_t)P.nA * P.a_stage;
  if (warp == 0 && P.halo) {
    // ===== TMA producer, halo mode: one halo box per tile and chunk, one weight box per (tile, chunk, tap) =====
    // The halo boxes run AHEAD of the weight stream (up to nA chunks): a box of ~200 rows has a long latency, and the weight ring
    // (which the same thread feeds, blocking) is only a few taps deep. A weight box of chunk c is never issued before the halo box
    // of chunk c (the MMA thread consumes in that order: a blocking wait here can then only be on slots that free by themselves).
    if (lane == 0) {
      uint32_t itA = 0, sa = 0, pha = 0, sw = 0, phw = 0;
      const uint32_t w_tx = (uint32_t)P.NT * 128u;
      int tileA = cluster_id, kcA = 0;
      auto issue_A = [&](bool blocking) -> bool {
        if (tileA >= ntiles) return false;
        if (blocking) mbar_wait(&emptyA[sa], pha ^ 1u);
        else if (!mbar_try_wait(&emptyA[sa], pha ^ 1u)) return false;
        const Tile T = decode_tile(P, tileA, m_tiles, mgroups, crank);
        const int kbase = P.d.kblock ? (T.n0 / P.d.kblock) * P.d.kblock : 0;
        uint8_t* dstA = smem + (size_t)sa * P.a_stage;
        if (kPair) {
          const uint32_t lbar = mapa_u32(smem_u32(&fullA[sa]), 0);
          if (crank == 0) mbar_arrive_expect_tx(&fullA[sa], 2u * (uint32_t)P.a_box);
          tma_load_4d_2sm(dstA, &mapA, kbase + kcA * 32, T.w0 + P.dx0, T.h0 + P.dy0, T.img, lbar);
        } else {
          mbar_arrive_expect_tx(&fullA[sa], (uint32_t)P.a_box);
          tma_load_4d(dstA, &mapA, kbase + kcA * 32, T.w0 + P.dx0, T.h0 + P.dy0, T.img, &fullA[sa]);
        }
        ++itA;
        if (++sa == (uint32_t)P.nA) { sa = 0; pha ^= 1u; }
        if (++kcA == P.kchunks) { kcA = 0; tileA += n_clusters; }
        return true;
      };
      uint32_t chunk = 0;                                   // chunks (tile, kc) whose weight stream has started
      for (int tile = cluster_id; tile < ntiles; tile += n_clusters) {
        const Tile T = decode_tile(P, tile, m_tiles, mgroups, crank);
        for (int kc = 0; kc < P.kchunks; ++kc, ++chunk) {
          while (itA <= chunk) issue_A(true);               // this chunk's halo box first
          for (int t = 0; t < P.d.ntaps; ++t) {
            if ((t & 3) == 0)                                   // then, now and again, further boxes into free slots
              while (itA < chunk + (uint32_t)P.nA && issue_A(false)) {}
            const uint32_t s = sw, ph = phw;
            if (++sw == (uint32_t)P.nW) { sw = 0; phw ^= 1u; }
            uint8_t* st = smemW + (size_t)s * w_tx;
            const int wt = tap_wt[t];
            mbar_wait(&empty[s], ph ^ 1u);
            if (kPair) {
              const uint32_t lbar = mapa_u32(smem_u32(&full[s]), 0);
              if (crank == 0) mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)wslice * 128u);
              tma_load_3d_2sm(st, &mapW, kc * 32, T.n0 + crank * wslice, wt, lbar);
            } else {
              mbar_arrive_expect_tx(&full[s], w_tx);
              if (P.csize > 1) tma_load_3d_mc(st + crank * wslice * 128, &mapW, kc * 32, T.n0 + crank * wslice, wt, &full[s], cmask);
              else tma_load_3d(st, &mapW, kc * 32, T.n0, wt, &full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1 && P.halo) {
    // ===== MMA issuer, halo mode: tap = row offset of the A descriptor into the resident halo box =====
    // (this single thread is the issue path of the whole CTA: ring positions are carried as wrap-around counters, not computed
    // with run-time divisions, and the tap offsets come from a table in shared memory)
    if (lane == 0 && !(kPair && crank != 0)) {
      const uint32_t idesc = make_idesc_tf32(kPair ? 256 : 128, P.NT);
      const uint32_t sbo = (uint32_t)P.HW * 128u;
      const uint32_t w_stage = (uint32_t)P.NT * 128u, w_base = smem_u32(smemW);
      uint32_t sa = 0, pha = 0, sw = 0, phw = 0, lt = 0;
      for (int tile = cluster_id; tile < ntiles; tile += n_clusters, ++lt) {
        const uint32_t buf = lt & 1u, aph = (lt >> 1) & 1u;
        mbar_wait(&acc_empty[buf], aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem + buf * (uint32_t)kMaxNT;
        for (int kc = 0; kc < P.kchunks; ++kc) {
          mbar_wait(&fullA[sa], pha);
          tc_fence_after();
          const uint32_t abase = smem_u32(smem) + sa * (uint32_t)P.a_stage;
          for (int t = 0; t < P.d.ntaps; ++t) {
            mbar_wait(&full[sw], phw);
            tc_fence_after();
            const uint32_t sb = w_base + sw * w_stage;
            const uint32_t sa_t = abase + tap_off[t];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = make_sdesc_k_sw128_sbo(sa_t + ks * 32, sbo), bd = make_sdesc_k_sw128(sb + ks * 32);
              if (kPair) umma_ss_tf32_2sm(d, ad, bd, idesc, (kc | t | ks) ? 1u : 0u);
              else umma_ss_tf32(d, ad, bd, idesc, (kc | t | ks) ? 1u : 0u);
            }
            if (kPair) umma_commit_2sm_mc(&empty[sw], 3);
            else if (P.csize > 1) umma_commit_mc(&empty[sw], cmask);
            else umma_commit(&empty[sw]);
            if (++sw == (uint32_t)P.nW) { sw = 0; phw ^= 1u; }
          }
          if (kPair) umma_commit_2sm_mc(&emptyA[sa], 3);      // the MMAs that read this halo box (in both CTAs) are done
          else umma_commit(&emptyA[sa]);
          if (++sa == (uint32_t)P.nA) { sa = 0; pha ^= 1u; }
        }
        if (kPair) umma_commit_2sm_mc(&acc_full[buf], 3);
        else umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      SegIter seg(P.sk, cluster_id, n_clusters, ntiles, ksteps);
      int tile, k0, k1;
      while (seg.next(tile, k0, k1)) {
        const Tile T = decode_tile(P, tile, m_tiles, mgroups, crank);
        const int kbase = P.d.kblock ? (T.n0 / P.d.kblock) * P.d.kblock : 0;
        const int ws = T.w0 * P.d.stride, hs = T.h0 * P.d.stride;
        int t = k0 / P.kchunks, kc = k0 - t * P.kchunks;
        for (int k = k0; k < k1; ++t, kc = 0) {
          const int dy = P.d.dy[t], dx = P.d.dx[t], wt = P.d.wt[t];
          for (; kc < P.kchunks && k < k1; ++kc, ++k, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            uint8_t* st = smem + (size_t)s * kStageBytes;
            mbar_wait(&empty[s], ph ^ 1u);            // every CTA of the cluster has drained this stage
            if (kPair) {
              // both CTAs' bytes (own A tile + own half of the weight tile each) are counted on the leader's barrier
              const uint32_t lbar = mapa_u32(smem_u32(&full[s]), 0);
              if (crank == 0) mbar_arrive_expect_tx(&full[s], 2u * ((uint32_t)kABytes + (uint32_t)wslice * 128u));
              tma_load_4d_2sm(st, &mapA, kbase + kc * 32, ws + dx, hs + dy, T.img, lbar);
              tma_load_3d_2sm(st + kABytes, &mapW, kc * 32, T.n0 + crank * wslice, wt, lbar);
              continue;
            }
            mbar_arrive_expect_tx(&full[s], stage_tx);
            tma_load_4d(st, &mapA, kbase + kc * 32, ws + dx, hs + dy, T.img, &full[s]);
            if (P.csize > 1)
              tma_load_3d_mc(st + kABytes + crank * wslice * 128, &mapW, kc * 32, T.n0 + crank * wslice, wt, &full[s], cmask);
            else
              tma_load_3d(st + kABytes, &mapW, kc * 32, T.n0, wt, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (pair mode: the leader CTA issues for both) =====
    if (lane == 0 && !(kPair && crank != 0)) {
      const uint32_t idesc = make_idesc_tf32(kPair ? 256 : 128, P.NT);
      uint32_t it = 0, lt = 0;
      SegIter seg(P.sk, cluster_id, n_clusters, ntiles, ksteps);
      int tile, k0, k1;
      for (; seg.next(tile, k0, k1); ++lt) {
        const uint32_t buf = lt & 1u, aph = (lt >> 1) & 1u;
        mbar_wait(&acc_empty[buf], aph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem + buf * (uint32_t)kMaxNT;
        for (int k = 0; k < k1 - k0; ++k, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes), sb = sa + kABytes;
          if (kPair) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss_tf32_2sm(d, make_sdesc_k_sw128(sa + ks * 32), make_sdesc_k_sw128(sb + ks * 32), idesc, (k | ks) ? 1u : 0u);
            umma_commit_2sm_mc(&empty[s], 3);
            continue;
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss_tf32(d, make_sdesc_k_sw128(sa + ks * 32), make_sdesc_k_sw128(sb + ks * 32), idesc, (k | ks) ? 1u : 0u);
          if (P.csize > 1) umma_commit_mc(&empty[s], cmask);   // the stage also holds weight rows written by the peers
          else umma_commit(&empty[s]);
        }
        if (kPair) umma_commit_2sm_mc(&acc_full[buf], 3);
        else umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===== epilogue: one pixel (accumulator row) per thread; two warps per TMEM lane quarter take alternate channel blocks =====
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int hsel = (warp - 2) >> 2;            // 0 / 1: even / odd channel blocks
    const int et = threadIdx.x - 64;             // 0..255 among the epilogue threads
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* sb = smem + kOffStg + (size_t)(warp - 2) * kStgBytes;
    const bool has_aff = P.gamma != nullptr || P.bias != nullptr;
    // "this thread has read its part of accumulator `b`": in pair mode the arrival goes to the leader CTA's barrier
    auto release_acc = [&](uint32_t b) {
      tc_fence_before();
      if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[b]), 0));
      else mbar_arrive(&acc_empty[b]);
    };
    uint32_t lt = 0;
    SegIter seg(P.sk, cluster_id, n_clusters, ntiles, ksteps);
    int tile, k0, k1;
    // stream-K exchange area P.ws: kWsFlagWords flag words (one per CTA, zero when idle - fixed place, so that a launch with another
    // tile shape never finds stale data there), then one partial tile [128][NT] per CTA
    for (; seg.next(tile, k0, k1); ++lt) {
      const uint32_t buf = lt & 1u, aph = (lt >> 1) & 1u;
      if (k0 > 0) {
        // ---- partial tile (its first K-steps belong to an earlier cluster): raw accumulators -> exchange area, raise the flag ----
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after();
        const uint32_t d = tmem + buf * (uint32_t)kMaxNT + lane_base;
        // layout of a partial tile: [16-byte channel chunk][row] - the 32 rows of a warp store 512 contiguous bytes per instruction
        float* wrow = P.ws + kWsFlagWords + (size_t)blockIdx.x * 128 * P.NT + (size_t)row * 4;
        for (int c0 = hsel * 32; c0 < P.NT; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(d + c0, r);
          tmem_ld_wait();
          if (c0 + 64 >= P.NT) release_acc(buf);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + (size_t)(c0 + j) * 128), "r"(r[j]), "r"(r[j + 1]), "r"(r[j + 2]),
                         "r"(r[j + 3])
                         : "memory");
        }
        if (hsel * 32 >= P.NT) release_acc(buf);
        __threadfence();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (et == 0) {
          int* ws_flags = reinterpret_cast<int*>(P.ws);
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ws_flags + blockIdx.x), "r"(1) : "memory");
        }
        continue;
      }
      // a tile this cluster starts but does not finish: the clusters whose ranges begin inside it deliver the rest
      int n_part = 0;
      if (k1 < ksteps) {
        const long total = (long)ntiles * ksteps, tile_end = (long)(tile + 1) * ksteps;
        while (cluster_id + 1 + n_part < n_clusters && SegIter::bound(cluster_id + 1 + n_part, n_clusters, total, ksteps) < tile_end) ++n_part;
      }
      const Tile T = decode_tile(P, tile, m_tiles, mgroups, crank);
      const int h = T.h0 + row / P.TW, w = T.w0 + row % P.TW;
      const int yh = h * P.d.oy_mul + P.d.oy_add, yw = w * P.d.ox_mul + P.d.ox_add;
      const bool valid = T.valid && h < P.d.OH && w < P.d.OW && yh < P.d.YH && yw < P.d.YW;
      const size_t off = valid ? (((size_t)T.img * P.d.YH + yh) * P.d.YW + yw) * P.d.Cout + T.n0 : 0;
      const float* rrow = (P.res && valid) ? P.res + off : nullptr;
      const float* r2row = (P.res2 && valid) ? P.res2 + off : nullptr;
      const float* mrow = (P.mask && valid) ? P.mask + off : nullptr;
      float* yrow = P.y + off;
      if (P.tma_store) {
        const int ntv = min(P.NT, P.d.Cout - T.n0);      // channels of this tile that exist (ragged last tile)
        // residual / mask of this warp's first block: in flight while the MMAs of the tile are still running
        float4 resv[8], maskv[8];
        if (rrow && hsel * 32 < ntv) load_row<8>(rrow + hsel * 32, resv);
        if (mrow && hsel * 32 < ntv) load_row<8>(mrow + hsel * 32, maskv);
        if (has_aff) {
          // fold BatchNorm / bias of this tile's channels once: y = acc * scale + shift
          asm volatile("bar.sync 2, 256;" ::: "memory");      // previous tile's readers are done
          for (int i = et; i < ntv; i += 256) {
            const int c = T.n0 + i;
            float sc = 1.f, sh = 0.f;
            if (P.gamma) {
              sc = __ldg(P.gamma + c) * rsqrtf(__ldg(P.var + c) + P.d.bn_eps);
              sh = __ldg(P.beta + c) - __ldg(P.mean + c) * sc;
            }
            if (P.bias) sh = fmaf(__ldg(P.bias + c), sc, sh);
            aff_mem[i] = sc;
            aff_mem[kMaxNT + i] = sh;
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after();
        if (n_part) {
          // the partial tiles were written long ago (they are the FIRST thing the later clusters do): this wait is short
          if (et < n_part) {
            const int* f = reinterpret_cast<const int*>(P.ws) + (cluster_id + 1 + et) * P.csize + crank;
            int v = 0;
            const long long t0 = clock64();
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              if (!v && clock64() - t0 > 4000000000LL) {
                printf("dvd_b200: conv2d_tc_kernel: stream-K partial tile never arrived (block %d)\n", blockIdx.x);
                __trap();
              }
            } while (!v);
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          // fold the partial tiles into the accumulator (tensor memory, this thread's own row and channel blocks); the regular
          // epilogue below then runs unchanged
          const uint32_t dd = tmem + buf * (uint32_t)kMaxNT + lane_base;
          for (int c0 = hsel * 32; c0 < P.NT; c0 += 64) {
            for (int hh = 0; hh < 32; hh += 16) {
              uint32_t r[16];
              tmem_ld16(dd + c0 + hh, r);
              tmem_ld_wait();
              for (int pi = 0; pi < n_part; ++pi) {
                const float* prow = P.ws + kWsFlagWords + (size_t)((cluster_id + 1 + pi) * P.csize + crank) * 128 * P.NT + (size_t)row * 4 +
                                    (size_t)(c0 + hh) * 128;
                float4 a[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  // strong (gpu-scope) load: the tile was written by another SM; a weak load may be served from a stale L1 line
                  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                               : "=f"(a[j].x), "=f"(a[j].y), "=f"(a[j].z), "=f"(a[j].w)
                               : "l"(prow + (size_t)j * 512)
                               : "memory");
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + a[j].x);
                  r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + a[j].y);
                  r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + a[j].z);
                  r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + a[j].w);
                }
              }
              tmem_st16(dd + c0 + hh, r);
            }
          }
          tmem_st_wait();
        }
        const uint32_t d = tmem + buf * (uint32_t)kMaxNT + lane_base;
        const int bh0 = T.h0 + (q * 32) / P.TW, bw0 = T.w0 + (q * 32) % P.TW;
        for (int c0 = hsel * 32; c0 < ntv; c0 += 64) {
          uint32_t r[32];
          tmem_ld32(d + c0, r);
          tmem_ld_wait();
          if (c0 + 64 >= ntv) release_acc(buf);   // this warp has read its share of the accumulator
          epilogue_math<32>(P, r, has_aff ? aff_mem + c0 : nullptr, rrow ? resv : nullptr, r2row ? r2row + c0 : nullptr,
                            mrow ? maskv : nullptr);
          if (c0 + 64 < ntv) {                    // next block's residual / mask: in flight during the staging / store below
            if (rrow) load_row<8>(rrow + c0 + 64, resv);
            if (mrow) load_row<8>(mrow + c0 + 64, maskv);
          }
          if (lane == 0) bulk_wait_read<0>();      // the store that last read this warp's staging buffer has drained it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(sb + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (T.valid) tma_store_4d(&mapY, sb, T.n0 + c0, bw0, bh0, T.img);
            bulk_commit();
          }
        }
        if (hsel * 32 >= ntv) release_acc(buf);   // 32 channels: the odd warps have no block, but owe their arrival
        if (n_part) {
          // every reader is done with the partial tiles: lower the flags for the next launch (ordered by the kernel boundary)
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (et < n_part) reinterpret_cast<int*>(P.ws)[(cluster_id + 1 + et) * P.csize + crank] = 0;
        }
      } else {
        if (has_aff) {
          asm volatile("bar.sync 2, 256;" ::: "memory");
          for (int i = et; i < P.NT; i += 256) {
            const int c = T.n0 + i;
            float sc = 1.f, sh = 0.f;
            if (P.gamma) {
              sc = __ldg(P.gamma + c) * rsqrtf(__ldg(P.var + c) + P.d.bn_eps);
              sh = __ldg(P.beta + c) - __ldg(P.mean + c) * sc;
            }
            if (P.bias) sh = fmaf(__ldg(P.bias + c), sc, sh);
            aff_mem[i] = sc;
            aff_mem[kMaxNT + i] = sh;
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        mbar_wait(&acc_full[buf], aph);
        tc_fence_after();
        const uint32_t d = tmem + buf * (uint32_t)kMaxNT + lane_base;
        for (int c0 = hsel * 16; c0 < P.NT; c0 += 32) {
          uint32_t r[16];
          tmem_ld16(d + c0, r);
          float4 rcur[4], mcur[4];
          if (rrow) load_row<4>(rrow + c0, rcur);
          if (mrow) load_row<4>(mrow + c0, mcur);
          tmem_ld_wait();
          epilogue_math<16>(P, r, has_aff ? aff_mem + c0 : nullptr, rrow ? rcur : nullptr, r2row ? r2row + c0 : nullptr,
                            mrow ? mcur : nullptr);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(yrow + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          }
        }
        release_acc(buf);
      }
    }
    if (P.tma_store && lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (P.csize > 1) cluster_sync_all();         // no CTA leaves while a peer may still multicast into it / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight images. weight[co, ci_local, ky, kx] (arbitrary element strides), groups of `cpg` in-channels (cpg = Cin: dense).
//   mode 0 (forward)        out[t][co][c]          c over Cin (dense) or over the `kblock` in-channels of co's block
//   mode 1 (data gradient)  out[t][ci][c]          c over Cout (dense) or over the `kblock` out-channels of ci's block,
//                                                  times the eval-BatchNorm scale gamma * rsqrt(var + eps) of that out-channel
// t = ky * k + kx (never flipped: the tap table of the launch carries the offsets). Entries outside a channel's group are
// zero (block-diagonal image of a grouped convolution). Values rounded to TF32 (RN).
__global__ void __launch_bounds__(256) conv_pack_kernel(const float* __restrict__ w, long s_co, long s_ci, long s_ky, long s_kx,
                                                        float* __restrict__ out_fwd, float* __restrict__ out_bwd, int Cout, int Cin,
                                                        int k, int cpg, int kblock, const float* __restrict__ gamma,
                                                        const float* __restrict__ var, float eps) {
  DVD_PDL_ENTER();
  // both images have k*k * C * cols elements when Cin == Cout or dense; they are walked with one index each
  const int opg = Cout / (Cin / cpg);        // out-channels per group
  for (int mode = 0; mode < 2; ++mode) {
    float* out = mode ? out_bwd : out_fwd;
    if (!out) continue;
    const int rows = mode ? Cin : Cout;
    const int cols = kblock ? kblock : (mode ? Cout : Cin);
    const long n = (long)k * k * rows * cols;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
      const int c = (int)(i % cols);
      const long q = i / cols;
      const int r = (int)(q % rows), t = (int)(q / rows);
      const int ky = t / k, kx = t - ky * k;
      const int cabs = kblock ? (r / kblock) * kblock + c : c;     // absolute channel index of the column
      const int co = mode ? cabs : r, ci = mode ? r : cabs;
      float v = 0.f;
      if (co / opg == ci / cpg) {                                  // same group (dense: one group)
        v = w[co * s_co + (ci % cpg) * s_ci + ky * s_ky + kx * s_kx];
        if (mode && gamma) v *= gamma[co] * rsqrtf(var[co] + eps);
      }
      out[i] = round_tf32(v);
    }
  }
}

// All layers of a net in ONE launch (the weights change every optimisation step, so every step re-packs ~250 images: as single
// launches that is 123 grids of a few microseconds each). items[] lives in device memory; item i owns the blocks
// [blk0[i], blk0[i+1]) of the grid.
// One block = one 32 x 32 tile (out-channels x in-channels) of one tap: the forward image is written row by row, the data-gradient
// image is the transposed tile (through shared memory), so that reads of the weights and writes of BOTH images are coalesced.
__global__ void __launch_bounds__(256) conv_pack_batch_kernel(const dvd_pack_item* __restrict__ items, int n_items, int want_bwd) {
  DVD_PDL_ENTER();
  __shared__ float tile[32][33];
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {                       // last item whose first block is <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].blk0 <= (long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const dvd_pack_item it = items[lo];
  const int Cout = it.Cout, Cin = it.Cin, k = it.ksize, kblock = it.kblock;
  const int cpg = Cin / it.groups, opg = Cout / it.groups;
  const int cols = kblock ? kblock : Cin;                 // columns of a forward-image row
  const int nct = cols / 32, nrt = Cout / 32;
  const int tt = (int)((long)blockIdx.x - it.blk0);
  const int t = tt / (nrt * nct), rem = tt - t * (nrt * nct);
  const int co0 = (rem / nct) * 32, c0 = (rem % nct) * 32;
  const int base = kblock ? (co0 / kblock) * kblock : 0;  // first channel of the block-diagonal block (grouped)
  const int ky = t / k, kx = t - ky * k;
  const long tap = (long)ky * it.s_ky + (long)kx * it.s_kx;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  float* bwd = want_bwd ? it.w_bwd : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = wy * 4 + i, co = co0 + row, ci = base + c0 + lane;
    float v = 0.f;
    if (co / opg == ci / cpg) v = it.weight[co * it.s_co + (ci % cpg) * it.s_ci + tap];
    if (it.w_fwd) it.w_fwd[((long)t * Cout + co) * cols + c0 + lane] = round_tf32(v);
    if (bwd) tile[row][lane] = it.bn_gamma ? v * (it.bn_gamma[co] * rsqrtf(it.bn_var[co] + it.bn_eps)) : v;   // same association as conv_pack_kernel
  }
  if (!bwd) return;
  __syncthreads();
  const int cols_b = kblock ? kblock : Cout;              // columns of a data-gradient-image row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = wy * 4 + i, ci = base + c0 + row;     // row of the transposed tile = in-channel
    bwd[((long)t * Cin + ci) * cols_b + (co0 - base) + lane] = round_tf32(tile[lane][row]);
  }
}

// =====================================================================================================================
// Weight gradient:  D[m][n] += sum over pixels  Mop[px (+off), m] * Nop[px (+off), n]
// normally Mop = gy (out-channels), Nop = x shifted by the tap and sampled with the convolution stride; `swap` exchanges
// the roles (needs 128 | M-channels). GEMM with K = pixels: both operands MN-major (channel contiguous). A 64-pixel TMA box
// {32 ch, TW, TH, 1} lands as 64 rows of 128 bytes in the SWIZZLE_128B_BASE32B image (UMMA layout type 1); one MMA consumes
// 8 pixels. Split-K over pixel tiles across CTAs; partial sums leave through fp32 reductions into the caller's gradient
// buffer (any strides). Extras in the epilogue, all on the out-channel rows (swap = 0 only):
//   * eval-BatchNorm scale: dW[co] = sc[co] * sum gm X  (gm = the un-scaled masked gradient the data gradient also consumes)
//   * dgamma[co] += rstd[co] * <W[co], sum gm X>        (d/dgamma of BN(conv(x)) without touching any activation)
//   * grouped convolutions: only the diagonal 128-channel blocks are computed and only in-group entries leave.
constexpr int kWgStages = 2;
constexpr int kWgPx = 64;                         // pixels (K) per stage
constexpr int kWgBox = kWgPx * 128;               // bytes of one {32 ch, 64 px} box
constexpr int kWgStageBytes = (4 + 8) * kWgBox;   // M: 128 channels, N: up to 256 channels
constexpr int kWgOffOnes = kWgStages * kWgStageBytes;         // [64 px x 32] tile of 1.0: N operand of the column-sum MMA
constexpr int kWgOffBar = kWgOffOnes + kWgBox;
constexpr size_t kWgSmem = (size_t)kWgOffBar + 256;

struct WgradParams {
  float* dw;
  const float* w;                  // parameter tensor (same strides as dw), only read for dgamma
  long s_m, s_n, s_ky, s_kx;       // element strides of the M / N channel index and of the kernel taps
  int N, OH, OW;                   // pixel grid of gy
  int Mch, Nch;                    // channel counts of the two operands
  int ntaps, ksize, stride, swap;
  int TW, TH, tiles_w, tiles_h;    // 64-pixel tiles
  int NT;                          // N channels per output tile
  int ksplit;
  int csize;                       // cluster size: the CTAs of a cluster own consecutive 128-row M blocks of the same (tap, N tile,
                                   // pixel range) and share the N-operand stream (each loads every csize-th box, multicast)
  int pair;                        // csize == 2 as one tcgen05 CTA pair: M = 256 rows per MMA, each CTA keeps HALF of the N operand
                                   // (64 KB instead of 96 KB per 64-pixel stage: three stages instead of two)
  int cpg;                         // > 0: grouped (diagonal blocks, NT = 128)
  const float* gamma;              // eval BatchNorm of the out-channels (or null)
  const float* var;
  float eps;
  float* dgamma;
  float* colsum;                   // [Mch] += sum over pixels of the M operand (conv-bias / BatchNorm-beta gradient), or null
  const float* mean;               // with colsum and dgamma: dgamma[m] -= mean[m] * rstd[m] * colsum[m]
  signed char dy[DVD_CONV_MAX_TAPS], dx[DVD_CONV_MAX_TAPS];
  unsigned char wt[DVD_CONV_MAX_TAPS];
};

__device__ __forceinline__ uint64_t make_sdesc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) { return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16); }

template <bool kPair>
__global__ void __launch_bounds__(192, 1) conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapM,
                                                          const __grid_constant__ CUtensorMap mapN,
                                                          const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgOffBar);
  uint64_t* full = bars;            // [<= 4]
  uint64_t* empty = bars + 4;       // [<= 4]
  uint64_t* acc_full = bars + 8;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("dvd_b200: conv_wgrad_kernel: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  const int nb = P.NT / 32;                                   // 32-channel boxes of the N tile
  const int nb_own = kPair ? nb / 2 : nb;                    // ... resident in this CTA's shared memory
  const int nstages = kPair ? 3 : kWgStages;
  const uint32_t stage_bytes = (uint32_t)(4 + nb_own) * kWgBox <= 65536u && kPair ? 65536u : (uint32_t)kWgStageBytes;
  if (P.colsum) {     // every element 1.0: the swizzle of the operand image does not matter
    for (int i = threadIdx.x; i < kWgBox / 4; i += blockDim.x) reinterpret_cast<float*>(smem + kWgOffOnes)[i] = 1.0f;
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_2sm(tmem_holder, 512);
    else tmem_alloc(tmem_holder, 512);
  } else if (warp == 0 && lane == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kPair ? 1u : (uint32_t)P.csize); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (P.csize > 1) cluster_sync_all();
  const uint32_t tmem = *tmem_holder;
  DVD_PDL_ENTER();

  // this CTA: output tile (tap, 128 M-channels, NT N-channels) and a contiguous range of pixel tiles.
  // blockIdx.x = ((tap, M group, N tile) * ksplit + part) * csize + rank, M block = group * csize + rank
  const int crank = P.csize > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << P.csize) - 1u);
  const int cl = blockIdx.x / P.csize;
  const int out_grp = cl / P.ksplit, part = cl - out_grp * P.ksplit;
  const int n_mg = P.Mch / 128 / P.csize, n_n = P.cpg ? 1 : P.Nch / P.NT;
  const int t = out_grp / (n_mg * n_n);
  const int rem = out_grp - t * (n_mg * n_n);
  const int m0 = ((rem / n_n) * P.csize + crank) * 128;
  const int n0 = P.cpg ? m0 : (rem % n_n) * P.NT;
  const int dy = P.dy[t], dx = P.dx[t];
  // the CTAs of the first tap and first N tile also reduce their M operand over the pixels (one extra N = 16 MMA per K step)
  const bool do_colsum = P.colsum != nullptr && t == 0 && (P.cpg || rem % n_n == 0);
  const int px_tiles = P.N * P.tiles_h * P.tiles_w;
  const int per = (px_tiles + P.ksplit - 1) / P.ksplit;
  const int kt0 = part * per, kt1 = min(px_tiles, kt0 + per);
  const uint32_t stage_tx = (uint32_t)(4 + nb) * kWgBox;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int kt = kt0; kt < kt1; ++kt, ++it) {
        const uint32_t s = it % nstages, ph = (it / nstages) & 1u;
        const int img = kt / (P.tiles_h * P.tiles_w), r = kt - img * (P.tiles_h * P.tiles_w);
        const int h0 = (r / P.tiles_w) * P.TH, w0 = (r % P.tiles_w) * P.TW;
        // gy is sampled on the plain pixel grid, x at stride * pixel + tap offset
        const int gw = w0, gh = h0, xw = w0 * P.stride + dx, xh = h0 * P.stride + dy;
        const int mw = P.swap ? xw : gw, mh = P.swap ? xh : gh, nw = P.swap ? gw : xw, nh = P.swap ? gh : xh;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_wait(&empty[s], ph ^ 1u);
        if (kPair) {
          // own 128 M rows + own half of the N tile; both CTAs' bytes are counted on the leader's barrier
          const uint32_t lbar = mapa_u32(smem_u32(&full[s]), 0);
          if (crank == 0) mbar_arrive_expect_tx(&full[s], 2u * (uint32_t)(4 + nb_own) * kWgBox);
          for (int j = 0; j < 4; ++j) tma_load_4d_2sm(st + j * kWgBox, &mapM, m0 + 32 * j, mw, mh, img, lbar);
          for (int j = 0; j < nb_own; ++j)
            tma_load_4d_2sm(st + (4 + j) * kWgBox, &mapN, n0 + 32 * (crank * nb_own + j), nw, nh, img, lbar);
          continue;
        }
        mbar_arrive_expect_tx(&full[s], stage_tx);
        for (int j = 0; j < 4; ++j) tma_load_4d(st + j * kWgBox, &mapM, m0 + 32 * j, mw, mh, img, &full[s]);
        if (P.csize > 1) {
          for (int j = crank; j < nb; j += P.csize)
            tma_load_4d_mc(st + (4 + j) * kWgBox, &mapN, n0 + 32 * j, nw, nh, img, &full[s], cmask);
        } else {
          for (int j = 0; j < nb; ++j) tma_load_4d(st + (4 + j) * kWgBox, &mapN, n0 + 32 * j, nw, nh, img, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && !(kPair && crank != 0)) {
      const int Mi = kPair ? 256 : 128;
      const uint32_t idesc = make_idesc_tf32_mn(Mi, P.NT), idesc1 = make_idesc_tf32_mn(Mi, 16);
      const uint32_t so = smem_u32(smem + kWgOffOnes);
      uint32_t it = 0;
      for (int kt = kt0; kt < kt1; ++kt, ++it) {
        const uint32_t s = it % nstages, ph = (it / nstages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + 4 * kWgBox;
        if (kPair) {
#pragma unroll
          for (int ks = 0; ks < kWgPx / 8; ++ks) {
            umma_ss_tf32_2sm(tmem, make_sdesc_mn_sw128_32b(sa + ks * 1024, kWgBox), make_sdesc_mn_sw128_32b(sb + ks * 1024, kWgBox), idesc,
                             (it | ks) ? 1u : 0u);
            if (do_colsum)
              umma_ss_tf32_2sm(tmem + 256, make_sdesc_mn_sw128_32b(sa + ks * 1024, kWgBox),
                               make_sdesc_mn_sw128_32b(so + ks * 1024, kWgBox), idesc1, (it | ks) ? 1u : 0u);
          }
          umma_commit_2sm_mc(&empty[s], 3);
          continue;
        }
#pragma unroll
        for (int ks = 0; ks < kWgPx / 8; ++ks) {
          umma_ss_tf32(tmem, make_sdesc_mn_sw128_32b(sa + ks * 1024, kWgBox), make_sdesc_mn_sw128_32b(sb + ks * 1024, kWgBox), idesc,
                       (it | ks) ? 1u : 0u);
          if (do_colsum)
            umma_ss_tf32(tmem + 256, make_sdesc_mn_sw128_32b(sa + ks * 1024, kWgBox), make_sdesc_mn_sw128_32b(so + ks * 1024, kWgBox),
                         idesc1, (it | ks) ? 1u : 0u);
        }
        if (P.csize > 1) umma_commit_mc(&empty[s], cmask);
        else umma_commit(&empty[s]);
      }
      if (kPair) umma_commit_2sm_mc(acc_full, 3);
      else umma_commit(acc_full);
    }
  } else if (kt1 > kt0) {
    // epilogue: accumulator row = M channel, columns = N channels of this tile
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const uint32_t d = tmem + ((uint32_t)(q * 32) << 16);
    const int wt = P.wt[t];
    const long tap_off = (long)(wt / P.ksize) * P.s_ky + (long)(wt % P.ksize) * P.s_kx;
    float* dst = P.dw + (long)m * P.s_m + tap_off;
    const float* wrow = P.w ? P.w + (long)m * P.s_m + tap_off : nullptr;
    float sc = 1.f, rstd = 0.f;
    if (P.gamma) {
      rstd = rsqrtf(__ldg(P.var + m) + P.eps);
      sc = __ldg(P.gamma + m) * rstd;
    }
    float dot = 0.f;
    // (first block of the gamma-gradient weights: in flight while the last MMAs run)
    const bool wvec = !P.cpg && wrow != nullptr && P.s_n == 1 && ((reinterpret_cast<uintptr_t>(wrow + n0) & 15) == 0);
    float4 wnext[4];
    if (wvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) wnext[j] = __ldg(reinterpret_cast<const float4*>(wrow + n0) + j);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (P.cpg) {
      // diagonal block: columns of this row's group only; the warp's 32 rows cover max(cpg, 32) columns
      const int win = P.cpg > 32 ? P.cpg : 32;
      const int wstart = ((q * 32) / win) * win;
      const int g_row = (q * 32 + lane) / P.cpg;
      for (int c0 = wstart; c0 < wstart + win; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(d + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          if (col / P.cpg == g_row) {
            const float v = __uint_as_float(r[j]);
            const long o = (long)(col % P.cpg) * P.s_n;
            if (wrow) dot = fmaf(v, __ldg(wrow + o), dot);
            atomicAdd(dst + o, v * sc);
          }
        }
      }
    } else {
      // the weights of the gamma-gradient dot product are loaded one block AHEAD (this loop runs after the last MMA, nothing hides
      // its latencies): contiguous, 16-byte aligned rows (channels-last parameter buffers) as four 128-bit loads per block
      for (int c0 = 0; c0 < P.NT; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(d + c0, r);
        float4 wcur[4];
        if (wvec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) wcur[j] = wnext[j];
          if (c0 + 16 < P.NT) {
#pragma unroll
            for (int j = 0; j < 4; ++j) wnext[j] = __ldg(reinterpret_cast<const float4*>(wrow + n0 + c0 + 16) + j);
          }
        }
        tmem_ld_wait();
        if (wvec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dot = fmaf(__uint_as_float(r[4 * j]), wcur[j].x, dot);
            dot = fmaf(__uint_as_float(r[4 * j + 1]), wcur[j].y, dot);
            dot = fmaf(__uint_as_float(r[4 * j + 2]), wcur[j].z, dot);
            dot = fmaf(__uint_as_float(r[4 * j + 3]), wcur[j].w, dot);
          }
        } else if (wrow) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dot = fmaf(__uint_as_float(r[j]), __ldg(wrow + (long)(n0 + c0 + j) * P.s_n), dot);
        }
        if (P.s_n == 1 && ((reinterpret_cast<uintptr_t>(dst + n0 + c0) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + n0 + c0 + j), "f"(__uint_as_float(r[j]) * sc),
                         "f"(__uint_as_float(r[j + 1]) * sc), "f"(__uint_as_float(r[j + 2]) * sc), "f"(__uint_as_float(r[j + 3]) * sc)
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + (long)(n0 + c0 + j) * P.s_n, __uint_as_float(r[j]) * sc);
        }
      }
    }
    if (do_colsum) {
      uint32_t r1;
      tmem_ld1(d + 256, r1);
      tmem_ld_wait();
      const float cs = __uint_as_float(r1);
      atomicAdd(P.colsum + m, cs);
      if (P.dgamma && P.mean) dot = fmaf(-__ldg(P.mean + m), cs, dot);      // d gamma = rstd * (<W, dW> - mean * sum gm)
    }
    if (P.dgamma) atomicAdd(P.dgamma + m, dot * rstd);
  }
  tc_fence_before();
  __syncthreads();
  if (P.csize > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem, 512);
    else tmem_dealloc(tmem, 512);
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
             const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
  auto fn = encode_fn();
  DVD_ARG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                  elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DVD_ARG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
}

// 4-D map of an NHWC tensor [N][H][W][C]; a box of {32 ch, bw, bh, 1} ELEMENTS sampled every `es` pixels
int make_nhwc_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int bw, int bh, int es, CUtensorMapSwizzle swizzle) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  DVD_ARG_CHECK(bw * es <= 256 && bh * es <= 256, "TMA box too large (%d x %d at element stride %d)", bw, bh, es);
  return make_map(m, ptr, 4, dims, strides, box, estr, swizzle);
}

// TH x TW = npx (power of two) tile over an H x W grid with the least padding
void pick_tile(int H, int W, int npx, int max_tw, int* TW, int* TH) {
  long best = -1;
  for (int tw = npx < max_tw ? npx : max_tw; tw >= 8; tw >>= 1) {
    const int th = npx / tw;
    const long padded = (long)((W + tw - 1) / tw * tw) * ((H + th - 1) / th * th);
    if (best < 0 || padded < best) { best = padded; *TW = tw; *TH = th; }
  }
}

int per_device_attr(const void* func, size_t smem_bytes, bool (&done)[16]) {
  int dev = 0;
  DVD_CUDA_CALL(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) dev = 0;
  if (!done[dev]) {
    DVD_CUDA_CALL(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    done[dev] = true;
  }
  return 0;
}

// launch `grid` CTAs in clusters of `csize` (grid % csize == 0)
template <typename... Args>
int launch_clustered(void (*kernel)(Args...), int grid, int threads, size_t smem_bytes, int csize, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2u : 1u;
  DVD_CUDA_CALL(cudaLaunchKernelEx(&cfg, kernel, args...));
  return 0;
}

// how many clusters of `csize` CTAs of this kernel can be resident at once (cached per device and size)
template <typename K>
int max_clusters(K kernel, int threads, size_t smem_bytes, int csize) {
  static int cache[16][5] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) dev = 0;
  if (cache[dev][csize] > 0) return cache[dev][csize];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(num_sms() / csize * csize));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) n = num_sms() / csize;
  cache[dev][csize] = n;
  return n;
}

}  // namespace
}  // namespace dvd

using namespace dvd;

/* first (tile-major) K-step of every cluster's range under the stream-K schedule: out[0..n_clusters], out[n_clusters] = total
 * (host restatement of the kernel's own rule, for tests of the schedule's invariants) */
extern "C" int dvd_conv2d_streamk_bounds(int ntiles, int ksteps, int n_clusters, long* out) {
  DVD_ARG_CHECK(out && ntiles >= 1 && ksteps >= 1 && n_clusters >= 1 && (long)ntiles * ksteps >= n_clusters, "bad schedule shape");
  for (int c = 0; c <= n_clusters; ++c) out[c] = SegIter::bound(c, n_clusters, (long)ntiles * ksteps, ksteps);
  return 0;
}

extern "C" size_t dvd_conv2d_workspace_bytes(void) {
  // one partial tile [128][256] fp32 and one flag per resident CTA
  return (size_t)kWsFlagWords * sizeof(int) + (size_t)num_sms() * 128 * kMaxNT * sizeof(float);
}

extern "C" int dvd_conv2d_nhwc(const dvd_conv_desc* desc, const float* x, const float* w_img, const float* bias, const float* bn_gamma,
                               const float* bn_beta, const float* bn_mean, const float* bn_var, const float* res, const float* res2,
                               const float* mask, float* y, void* stream) {
  return dvd_conv2d_nhwc_ws(desc, x, w_img, bias, bn_gamma, bn_beta, bn_mean, bn_var, res, res2, mask, y, nullptr, 0, stream);
}

extern "C" int dvd_conv2d_nhwc_ws(const dvd_conv_desc* desc, const float* x, const float* w_img, const float* bias, const float* bn_gamma,
                                  const float* bn_beta, const float* bn_mean, const float* bn_var, const float* res, const float* res2,
                                  const float* mask, float* y, void* workspace, size_t workspace_bytes, void* stream) {
  DVD_ARG_CHECK(desc && x && w_img && y, "null pointer");
  ConvParams P{};
  P.d = *desc;
  dvd_conv_desc& d = P.d;
  DVD_ARG_CHECK(d.N >= 1 && d.H >= 1 && d.W >= 1 && d.OH >= 1 && d.OW >= 1, "bad shape N=%d H=%d W=%d OH=%d OW=%d", d.N, d.H, d.W, d.OH, d.OW);
  DVD_ARG_CHECK(d.stride == 1 || d.stride == 2, "stride must be 1 or 2");
  DVD_ARG_CHECK(d.ntaps >= 1 && d.ntaps <= DVD_CONV_MAX_TAPS, "ntaps out of range");
  DVD_ARG_CHECK(d.kblock == 0 || (d.kblock % 32 == 0 && d.kblock <= kMaxNT && d.Cout % d.kblock == 0 && d.Cin % d.kblock == 0),
                "bad kblock %d (Cin=%d Cout=%d)", d.kblock, d.Cin, d.Cout);
  if (d.Cin % 32 != 0 || d.Cout % 16 != 0) {
    set_error("dvd_conv2d_nhwc: needs Cin %% 32 == 0 and Cout %% 16 == 0 (Cin=%d Cout=%d)", d.Cin, d.Cout);
    return -2;
  }
  DVD_ARG_CHECK(aligned16(x) && aligned16(w_img) && aligned16(y) && (!res || aligned16(res)) && (!res2 || aligned16(res2)) &&
                    (!mask || aligned16(mask)),
                "tensors must be 16-byte aligned");
  DVD_ARG_CHECK((bn_gamma != nullptr) == (bn_beta != nullptr) && (bn_gamma != nullptr) == (bn_mean != nullptr) &&
                    (bn_gamma != nullptr) == (bn_var != nullptr),
                "BatchNorm needs all of gamma, beta, mean, var (or none)");
  P.bias = bias; P.gamma = bn_gamma; P.beta = bn_beta; P.mean = bn_mean; P.var = bn_var;
  P.res = res; P.res2 = res2; P.mask = mask; P.y = y;
  P.NT = d.kblock ? d.kblock : (d.Cout >= kMaxNT ? kMaxNT : d.Cout);
  P.kchunks = (d.kblock ? d.kblock : d.Cin) / 32;
  const bool identity_out = d.oy_mul == 1 && d.ox_mul == 1 && d.oy_add == 0 && d.ox_add == 0 && d.YH == d.OH && d.YW == d.OW;
  const bool pointwise = d.ntaps == 1 && d.dy[0] == 0 && d.dx[0] == 0 && d.stride == 1 && identity_out && d.H == d.OH && d.W == d.OW;
  int inN = d.N, inH = d.H, inW = d.W;
  if (pointwise) {
    // a 1x1 stride-1 convolution is a plain GEMM over all N*H*W pixels: one "image" of P x 1
    const long Pn = (long)d.N * d.H * d.W;
    DVD_ARG_CHECK(Pn < (1L << 31), "too many pixels");
    inN = 1; inH = 1; inW = (int)Pn;
    d.N = 1; d.OH = 1; d.OW = (int)Pn; d.YH = 1; d.YW = (int)Pn;
    P.TW = 128; P.TH = 1;
  } else {
    pick_tile(d.OH, d.OW, 128, 128, &P.TW, &P.TH);   // the TMA box spans TW * stride <= 256 input pixels
  }
  // Halo-resident tiles for the classes that are bound by the activation stream, not by the tensor cores: grouped layers (64-
  // channel blocks) and layers with few output channels. Tile = 16 rows x 8 pixels (a tile row = one 8-row group of the MMA's A
  // operand); per 32-channel chunk one box of (8 + dx range) x (16 + dy range) pixels replaces ntaps boxes of 128 pixels.
  P.halo = 0;
  if (!pointwise && d.stride == 1 && d.ntaps >= 2) {
    int dy0 = 127, dy1 = -128, dx0 = 127, dx1 = -128;
    for (int t = 0; t < d.ntaps; ++t) {
      dy0 = d.dy[t] < dy0 ? d.dy[t] : dy0; dy1 = d.dy[t] > dy1 ? d.dy[t] : dy1;
      dx0 = d.dx[t] < dx0 ? d.dx[t] : dx0; dx1 = d.dx[t] > dx1 ? d.dx[t] : dx1;
    }
    const int HW = 8 + dx1 - dx0, HH = 16 + dy1 - dy0;
    const long a_box = (long)HW * HH * 128, a_stage = (a_box + 1023) / 1024 * 1024;
    const bool legal = HW <= 256 && HH <= 256 && 2 * a_stage + 2 * (long)kMaxNT * 128 + 1024 <= (long)kStages * kStageBytes;
    const int nt0 = d.kblock ? d.kblock : (d.Cout >= kMaxNT ? kMaxNT : d.Cout);
    const long tiles_tap = (long)((d.OW + P.TW - 1) / P.TW) * ((d.OH + P.TH - 1) / P.TH);
    const long tiles_halo = (long)((d.OW + 7) / 8) * ((d.OH + 15) / 16);
    // bytes through the TMA unit per image and 32-channel chunk (weights: half a tile per CTA when two CTAs share them)
    const long wb = (long)nt0 * 128 / 2;
    const bool pays = tiles_halo * (a_box + d.ntaps * wb) * 10 < tiles_tap * d.ntaps * (16384 + wb) * 8;
    // Measured (profiles/r3_halo_vs_taps.txt): correct for every stencil, 3-6x fewer activation bytes through L2 - and still 10-25 %
    // SLOWER than one box per tap on the three classes it was built for (head 128->32, grouped 3x3, dense 3x3). With operands this
    // small the step time is set by the single MMA-issuing thread and the latency of the small weight boxes, not by bytes; the
    // traffic model below is therefore not used to switch the mode on. DVD_CONV_HALO=1 selects it wherever it is legal.
    bool want = false;
    (void)pays;
    (void)nt0;
    if (const char* ev = getenv("DVD_CONV_HALO")) want = legal && atoi(ev) != 0;
    if (want) {
      P.halo = 1; P.HW = HW; P.HH = HH; P.dy0 = dy0; P.dx0 = dx0; P.a_box = (int)a_box; P.a_stage = (int)a_stage;
      P.TW = 8; P.TH = 16;
    }
  }
  P.tiles_w = (d.OW + P.TW - 1) / P.TW;
  P.tiles_h = (d.OH + P.TH - 1) / P.TH;
  P.tma_store = identity_out && (P.NT % 32 == 0);
  P.sbw = P.TW < 32 ? P.TW : 32;
  P.sbh = 32 / P.sbw;
  static bool attr_done[16] = {};
  if (int e = per_device_attr((const void*)conv2d_tc_kernel<false>, kSmem, attr_done)) return e;
  static bool attr_done_p[16] = {};
  if (int e = per_device_attr((const void*)conv2d_tc_kernel<true>, kSmem, attr_done_p)) return e;
  // cluster size: the weight tile is fetched once per cluster (NT / csize rows per CTA, whole 8-row swizzle atoms)
  const long m_tiles_h = (long)d.N * P.tiles_w * P.tiles_h;
  P.csize = 1;
  {
    // clusters of 4 may leave SMs idle (a GPC whose SM count is not a multiple of 4): take 4 only when it keeps the machine full
    const int sm4 = 4 * max_clusters(conv2d_tc_kernel<false>, kThreads, kSmem, 4), sm2 = 2 * max_clusters(conv2d_tc_kernel<false>, kThreads, kSmem, 2);
    if (m_tiles_h >= 8 && P.NT % 32 == 0 && sm4 + 4 >= sm2) P.csize = 4;
    else if (m_tiles_h >= 2 && P.NT % 16 == 0) P.csize = 2;
  }
  if (const char* ev = getenv("DVD_CONV_CLUSTER")) {      // tuning / A-B override: 1, 2 or 4
    const int v = atoi(ev);
    if ((v == 1 || v == 2 || v == 4) && (P.NT / v) % 8 == 0) P.csize = v;
  }
  // dense layers with at least two pixel tiles run as tcgen05 CTA pairs (cta_group::2); DVD_CONV_PAIR=0 falls back to multicast
  P.pair = 0;
  if (P.csize == 2 && d.kblock == 0 && P.NT % 32 == 0) {
    const char* ev = getenv("DVD_CONV_PAIR");
    P.pair = (ev && atoi(ev) == 0) ? 0 : 1;
  }
  const long full = max_clusters(conv2d_tc_kernel<false>, kThreads, kSmem, P.csize);
  const long mgroups_h = (m_tiles_h + P.csize - 1) / P.csize;
  // Channel-tile width. 256 is the most efficient tile, but whole tiles come in waves: 84 tiles on 74 CTA pairs (the 45 1x1
  // layers of ResNeXt's layer3 at 16 images) take two rounds of which the second is 14 % full. A narrower tile that fills the
  // rounds wins when  rounds(NT) * t(NT)  is smaller, t(NT) = 0.47 + 0.53 * NT / 256 the measured relative cost of a K-step
  // (profiles/r2_conv_table: 3x3 256->256 at NT 256 vs 256->128 at NT 128). Dense layers with a TMA-store epilogue only; the
  // last tile may be ragged in whole 32-channel blocks.
  if (!d.kblock && P.tma_store && d.Cout >= 128 && d.Cout % 32 == 0) {
    const char* ev = getenv("DVD_CONV_NT");
    if (ev && atoi(ev) >= 32 && atoi(ev) <= kMaxNT && atoi(ev) % 32 == 0) {
      P.NT = atoi(ev);
    } else if (!(ev && atoi(ev) == 0)) {
      // a layer that cannot even fill one round with 256-wide tiles (small batches, low resolutions) may go down to 64
      const int nt_max = d.Cout >= kMaxNT ? kMaxNT : d.Cout;
      int nt_min = mgroups_h * ((d.Cout + nt_max - 1) / nt_max) < full ? 64 : 128;
      {
        // layers that will take the stream-K schedule (few tiles, long K loop) keep the widest tile: their K-steps are spread
        // over all SMs whatever the tile count, and a wide step is the cheapest per output channel
        const long ksteps = (long)d.ntaps * ((d.kblock ? d.kblock : d.Cin) / 32), tiles0 = mgroups_h * ((d.Cout + nt_max - 1) / nt_max);
        const long waves0 = (tiles0 + full - 1) / full, total0 = tiles0 * ksteps, mu = ksteps / 4 > 8 ? ksteps / 4 : 8;
        long nc0 = total0 / mu;
        if (nc0 > full) nc0 = full;
        const char* evs = getenv("DVD_CONV_STREAMK");
        if (workspace && !P.halo && ksteps >= 8 && !(evs && atoi(evs) == 0) && nc0 >= 1 &&
            waves0 * ksteps - (total0 + nc0 - 1) / nc0 >= 64)
          nt_min = nt_max;
      }
      double best = 1e30;
      for (int nt = nt_max; nt >= nt_min; nt -= 32) {
        const long tiles = mgroups_h * ((d.Cout + nt - 1) / nt);
        const double cost = (double)((tiles + full - 1) / full) * (0.47 + 0.53 * nt / 256.0);
        if (cost < best * 0.97) { best = cost; P.NT = nt; }      // a narrower tile must win by 3 %
      }
    }
  }
  if (d.Cout % P.NT != 0 && !(P.tma_store && d.Cout % 32 == 0 && !d.kblock)) {
    set_error("dvd_conv2d_nhwc: Cout=%d is not a multiple of the %d-channel tile", d.Cout, P.NT);
    return -2;
  }
  if (P.halo) {
    const long ring = (long)kStages * kStageBytes - 1024, wbytes = (long)P.NT * 128;      // 1 KB: the tap tables
    P.nA = 4;
    while (P.nA > 2 && (ring - (long)P.nA * P.a_stage) / wbytes < 4) --P.nA;
    long nW = (ring - (long)P.nA * P.a_stage) / wbytes;
    P.nW = (int)(nW > 8 ? 8 : nW);
    DVD_ARG_CHECK(P.nW >= 2, "halo mode: the rings do not fit (box %d bytes, NT %d)", P.a_box, P.NT);
  }
  CUtensorMap mapA, mapW, mapY;
  if (P.halo) {
    if (int e = make_nhwc_map(&mapA, x, inN, inH, inW, d.Cin, P.HW, P.HH, 1, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  } else if (int e = make_nhwc_map(&mapA, x, inN, inH, inW, d.Cin, P.TW, P.TH, d.stride, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  {
    int max_wt = 0;
    for (int t = 0; t < d.ntaps; ++t) max_wt = d.wt[t] > max_wt ? d.wt[t] : max_wt;
    const int Kw = d.kblock ? d.kblock : d.Cin;
    const cuuint64_t dims[3] = {(cuuint64_t)Kw, (cuuint64_t)d.Cout, (cuuint64_t)(max_wt + 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)Kw * 4, (cuuint64_t)d.Cout * Kw * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)(P.NT / P.csize), 1};
    const cuuint32_t ones[3] = {1, 1, 1};
    if (int e = make_map(&mapW, w_img, 3, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  if (P.tma_store) {
    if (int e = make_nhwc_map(&mapY, y, d.N, d.YH, d.YW, d.Cout, P.sbw, P.sbh, 1, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  } else {
    mapY = mapA;
  }
  const long ctiles = mgroups_h * ((d.Cout + P.NT - 1) / P.NT);      // cluster tiles
  long nclusters = ctiles < full ? ctiles : full;
  // stream-K when whole tiles would leave a large part of the machine idle in the last (or only) wave: cut the (tile, K-step)
  // list into equal contiguous ranges instead. Needs the caller's exchange area (zeroed once; the kernel leaves it zeroed where
  // it matters: the flags). A range is at least max(8, ksteps / 4) K-steps long, so a tile is shared by at most ~4 clusters.
  P.sk = 0;
  P.ws = nullptr;
  {
    const long ksteps = (long)d.ntaps * P.kchunks, total = ctiles * ksteps;
    const long waves = (ctiles + full - 1) / full;
    const double eff = (double)ctiles / (double)(waves * full);
    const char* ev = getenv("DVD_CONV_STREAMK");
    const bool allowed = !(ev && atoi(ev) == 0);
    if (allowed && workspace && P.tma_store && !P.halo && ksteps >= 8 && eff < 0.88) {
      const long min_units = ksteps / 4 > 8 ? ksteps / 4 : 8;
      long nc = total / min_units;
      if (nc > full) nc = full;
      const size_t need = (size_t)kWsFlagWords * sizeof(int) + (size_t)nc * P.csize * 128 * (size_t)P.NT * sizeof(float);
      // the exchange (partial tiles through L2, a short serial tail in the finishing CTA) costs about as much as 50 K-steps
      // (measured: profiles/r3_streamk_table.txt): only layers with few tiles and a long K loop gain - the 3x3 layerN_rn
      // convolutions at 1/16 and 1/32 resolution (2.6x / 1.7x faster), not the 32-step 1x1 layers of the encoder
      const long dp_steps = waves * ksteps, sk_steps = (total + nc - 1) / nc;
      if ((nc > nclusters || (nc == nclusters && waves > 1)) && dp_steps - sk_steps >= 64) {
        DVD_ARG_CHECK(aligned16(workspace), "workspace must be 16-byte aligned");
        if (need <= workspace_bytes && nc * P.csize <= kWsFlagWords) {
          P.sk = 1;
          P.ws = static_cast<float*>(workspace);
          nclusters = nc;
        }
      }
    }
  }
  if (int e = P.pair ? launch_clustered(conv2d_tc_kernel<true>, (int)nclusters * P.csize, kThreads, kSmem, P.csize, (cudaStream_t)stream,
                                        mapA, mapW, mapY, P)
                     : launch_clustered(conv2d_tc_kernel<false>, (int)nclusters * P.csize, kThreads, kSmem, P.csize, (cudaStream_t)stream,
                                        mapA, mapW, mapY, P))
    return e;
  DVD_CUDA_LAUNCH_CHECK("conv2d_tc_kernel");
  return 0;
}

extern "C" int dvd_conv2d_pack(const float* weight, long stride_co, long stride_ci, long stride_ky, long stride_kx, float* w_fwd,
                               float* w_bwd, int Cout, int Cin, int ksize, int groups, int kblock, const float* bn_gamma,
                               const float* bn_var, float bn_eps, void* stream) {
  DVD_ARG_CHECK(weight && (w_fwd || w_bwd), "null pointer");
  DVD_ARG_CHECK(Cout >= 1 && Cin >= 1 && ksize >= 1 && ksize <= 11 && groups >= 1 && Cin % groups == 0 && Cout % groups == 0,
                "bad weight shape");
  DVD_ARG_CHECK((groups == 1) == (kblock == 0), "kblock must be set exactly for grouped convolutions");
  const int cpg = Cin / groups;
  if (kblock) DVD_ARG_CHECK(Cin == Cout && kblock % cpg == 0 && Cin % kblock == 0, "grouped: needs Cin == Cout and cpg | kblock | Cin");
  DVD_ARG_CHECK((bn_gamma != nullptr) == (bn_var != nullptr), "BatchNorm scale needs gamma and var");
  const long n = (long)ksize * ksize * (Cout > Cin ? Cout : Cin) * (kblock ? kblock : (Cout > Cin ? Cin : Cout));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  dvd::launch(conv_pack_kernel, blocks, 256, 0, (cudaStream_t)stream, weight, stride_co, stride_ci, stride_ky, stride_kx, w_fwd, w_bwd, Cout, Cin,
                                                           ksize, cpg, kblock, bn_gamma, bn_var, bn_eps);
  DVD_CUDA_LAUNCH_CHECK("conv_pack_kernel");
  return 0;
}

extern "C" long dvd_conv2d_pack_blocks(int Cout, int Cin, int ksize, int groups, int kblock) {
  if (Cout < 32 || Cin < 32 || ksize < 1 || groups < 1 || Cout % 32 || Cin % 32 || (kblock && kblock % 32)) return -1;
  return (long)ksize * ksize * (Cout / 32) * ((kblock ? kblock : Cin) / 32);      // one block per 32 x 32 tile and tap
}

extern "C" int dvd_conv2d_pack_batch(const dvd_pack_item* items_dev, int n_items, long total_blocks, int want_bwd, void* stream) {
  DVD_ARG_CHECK(items_dev && n_items >= 1 && total_blocks >= 1 && total_blocks < (1L << 31), "bad pack table");
  dvd::launch(conv_pack_batch_kernel, (unsigned)total_blocks, 256, 0, (cudaStream_t)stream, items_dev, n_items, want_bwd);
  DVD_CUDA_LAUNCH_CHECK("conv_pack_batch_kernel");
  return 0;
}

extern "C" int dvd_conv2d_wgrad(const dvd_conv_desc* desc, const float* x, const float* gy, float* dweight, const float* weight,
                                long stride_co, long stride_ci, long stride_ky, long stride_kx, int ksize, int groups,
                                const float* bn_gamma, const float* bn_var, float* dgamma, float* colsum, const float* bn_mean,
                                void* stream) {
  DVD_ARG_CHECK(desc && x && gy && dweight, "null pointer");
  const dvd_conv_desc& d = *desc;
  DVD_ARG_CHECK(d.N >= 1 && d.H >= 1 && d.W >= 1 && d.OH >= 1 && d.OW >= 1, "bad shape");
  DVD_ARG_CHECK(d.stride == 1 || d.stride == 2, "stride must be 1 or 2");
  DVD_ARG_CHECK(d.ntaps >= 1 && d.ntaps <= DVD_CONV_MAX_TAPS, "ntaps out of range");
  DVD_ARG_CHECK(aligned16(x) && aligned16(gy), "tensors must be 16-byte aligned");
  DVD_ARG_CHECK((bn_gamma != nullptr) == (bn_var != nullptr) && (!dgamma || (bn_gamma && weight)), "BatchNorm extras need gamma, var (and weight for dgamma)");
  WgradParams P{};
  P.dw = dweight; P.w = dgamma ? weight : nullptr;
  P.N = d.N; P.OH = d.OH; P.OW = d.OW; P.ntaps = d.ntaps; P.ksize = ksize; P.stride = d.stride;
  P.gamma = bn_gamma; P.var = bn_var; P.eps = d.bn_eps; P.dgamma = dgamma; P.colsum = colsum; P.mean = bn_mean;
  for (int t = 0; t < d.ntaps; ++t) { P.dy[t] = d.dy[t]; P.dx[t] = d.dx[t]; P.wt[t] = d.wt[t]; }
  const int Cin = d.Cin, Cout = d.Cout;
  if (groups > 1) {
    P.cpg = Cin / groups;
    if (Cin != Cout || Cin % 128 != 0 || 128 % P.cpg != 0) {
      set_error("dvd_conv2d_wgrad: grouped needs Cin == Cout, 128 | C and cpg | 128 (C=%d groups=%d)", Cin, groups);
      return -2;
    }
    P.swap = 0; P.Mch = Cout; P.Nch = Cin; P.NT = 128;
    P.s_m = stride_co; P.s_n = stride_ci;
  } else if (Cout % 128 == 0 && Cin % 32 == 0 && (Cin <= 256 || Cin % 256 == 0)) {
    P.swap = 0; P.Mch = Cout; P.Nch = Cin; P.NT = Cin >= 256 ? 256 : Cin;
    P.s_m = stride_co; P.s_n = stride_ci;
  } else if (Cin % 128 == 0 && Cout % 32 == 0 && (Cout <= 256 || Cout % 256 == 0) && !bn_gamma) {
    P.swap = 1; P.Mch = Cin; P.Nch = Cout; P.NT = Cout >= 256 ? 256 : Cout;
    P.s_m = stride_ci; P.s_n = stride_co;
    DVD_ARG_CHECK(colsum == nullptr, "column sums are not available with swapped operands (Cout %% 128 != 0): use dvd_relu_bwd_colsum");
  } else {
    set_error("dvd_conv2d_wgrad: unsupported channel counts Cin=%d Cout=%d", Cin, Cout);
    return -2;
  }
  P.s_ky = stride_ky; P.s_kx = stride_kx;
  int N = d.N, OH = d.OH, OW = d.OW, H = d.H, W = d.W;
  const bool pointwise = d.ntaps == 1 && d.dy[0] == 0 && d.dx[0] == 0 && d.stride == 1 && d.H == d.OH && d.W == d.OW;
  if (pointwise) {
    const long Pn = (long)N * OH * OW;
    DVD_ARG_CHECK(Pn < (1L << 31), "too many pixels");
    N = 1; OH = 1; OW = (int)Pn; H = 1; W = (int)Pn;
    P.N = 1; P.OH = 1; P.OW = (int)Pn;
    P.TW = kWgPx; P.TH = 1;
  } else {
    pick_tile(OH, OW, kWgPx, kWgPx, &P.TW, &P.TH);
  }
  P.tiles_w = (OW + P.TW - 1) / P.TW;
  P.tiles_h = (OH + P.TH - 1) / P.TH;
  const int out_tiles = d.ntaps * (P.Mch / 128) * (P.cpg ? 1 : P.Nch / P.NT);
  const int px_tiles = N * P.tiles_h * P.tiles_w;
  static bool attr_done[16] = {};
  if (int e = per_device_attr((const void*)conv_wgrad_kernel<false>, kWgSmem, attr_done)) return e;
  static bool attr_done_p[16] = {};
  if (int e = per_device_attr((const void*)conv_wgrad_kernel<true>, kWgSmem, attr_done_p)) return e;
  // cluster over consecutive 128-row M blocks: they consume the same N-operand boxes
  P.csize = 1;
  if (!P.cpg) {
    const int n_m = P.Mch / 128, nb = P.NT / 32;
    const int sm4 = 4 * max_clusters(conv_wgrad_kernel<false>, 192, kWgSmem, 4), sm2 = 2 * max_clusters(conv_wgrad_kernel<false>, 192, kWgSmem, 2);
    if (n_m % 4 == 0 && nb % 4 == 0 && sm4 + 4 >= sm2) P.csize = 4;
    else if (n_m % 2 == 0 && nb % 2 == 0) P.csize = 2;
    if (const char* ev = getenv("DVD_WGRAD_CLUSTER")) {
      const int v = atoi(ev);
      if ((v == 1 || v == 2 || v == 4) && n_m % v == 0 && nb % v == 0) P.csize = v;
    }
  }
  P.pair = 0;
  if (P.csize == 2 && (P.NT / 32) % 2 == 0) {
    const char* ev = getenv("DVD_WGRAD_PAIR");
    P.pair = (ev && atoi(ev) == 0) ? 0 : 1;
  }
  // split-K so that all CTAs are resident in ONE wave (a second, nearly empty wave would double the time)
  const int resident = P.csize * max_clusters(conv_wgrad_kernel<false>, 192, kWgSmem, P.csize);
  int ksplit = resident / out_tiles;
  if (ksplit > px_tiles) ksplit = px_tiles;
  if (ksplit < 1) ksplit = 1;
  P.ksplit = ksplit;
  CUtensorMap mapG, mapX;
  if (int e = make_nhwc_map(&mapG, gy, N, OH, OW, Cout, P.TW, P.TH, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  if (int e = make_nhwc_map(&mapX, x, N, H, W, Cin, P.TW, P.TH, d.stride, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  if (int e = P.pair ? launch_clustered(conv_wgrad_kernel<true>, out_tiles * ksplit, 192, kWgSmem, P.csize, (cudaStream_t)stream,
                                        P.swap ? mapX : mapG, P.swap ? mapG : mapX, P)
                     : launch_clustered(conv_wgrad_kernel<false>, out_tiles * ksplit, 192, kWgSmem, P.csize, (cudaStream_t)stream,
                                        P.swap ? mapX : mapG, P.swap ? mapG : mapX, P))
    return e;
  DVD_CUDA_LAUNCH_CHECK("conv_wgrad_kernel");
  return 0;
}

/* resident CTAs of the two tensor-core kernels for cluster sizes 1, 2, 4 (diagnostic): out[0..2] forward / data gradient kernel,
 * out[3..5] weight-gradient kernel */
extern "C" int dvd_conv2d_cluster_info(int* out) {
  static bool a1[16] = {}, a2[16] = {};
  if (int e = per_device_attr((const void*)conv2d_tc_kernel<false>, kSmem, a1)) return e;
  if (int e = per_device_attr((const void*)conv_wgrad_kernel<false>, kWgSmem, a2)) return e;
  const int cs[3] = {1, 2, 4};
  for (int i = 0; i < 3; ++i) {
    out[i] = cs[i] * max_clusters(conv2d_tc_kernel<false>, kThreads, kSmem, cs[i]);
    out[3 + i] = cs[i] * max_clusters(conv_wgrad_kernel<false>, 192, kWgSmem, cs[i]);
  }
  return 0;
}
