// Scene-flow MLP on tcgen05 tensor cores (sm_100a): fused Euler-chain forward, fused dgrad chain,
// split-K wgrad. Replaces (reference paths relative to the reference tree):
//   PeriodicEmbed.forward                     networks/blocks.py:19-34                       (M1)
//   SceneFlowFieldNet.forward / Conv2dBlock   networks/sceneflow_field.py:20-53, blocks.py:50-102 (M2)
//   Model.forward_sf_net (+ /sf_mag_div)      models/scene_flow_motion_field.py:346-358      (M3)
//   Model.forward_sf_net_multi_step           models/scene_flow_motion_field.py:360-367      (M4)
//   autograd backward of the above (the reference materialises a [B,256,H,W] fp32 activation per layer).
//
// Design: one CTA owns a tile of 128 pixels and pushes it through ALL layers and all
// Euler steps without leaving the SM. The activation tile is the A operand and lives in TENSOR MEMORY
// (lane = pixel, packed bf16 pairs along K); the accumulator D [128 x 256] fp32 lives in TMEM as well.
// Weights are the B operand: streamed per layer from L2 into a 6-stage shared-memory ring by 1-D bulk
// async copies of pre-swizzled blocks (see sf_mlp_layout.cuh), consumed by single-thread tcgen05.mma.
// fp32 accuracy comes from the bf16 (hi, lo) split: D += Ahi*Bhi + Alo*Bhi + Ahi*Blo.
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 = epilogue
// (TMEM -> registers: bias, LeakyReLU, hi/lo split -> TMEM; embedding; Euler update); two warps share a TMEM
// lane quarter and split the 256 columns, each keeping its own copy of the per-pixel state.
#include "common.cuh"
#include "sf_mlp_layout.cuh"
#include "tc_common.cuh"

namespace dvd {
using namespace tc;

__host__ __device__ __forceinline__ uint32_t act_offset(uint32_t ch, uint32_t px) {
  return kActInterleave ? il_offset(ch, px) : mn128_offset(ch, px);
}
__device__ __forceinline__ uint64_t act_desc(uint32_t smem_addr, int ks) {
  return kActInterleave ? make_sdesc_mn_interleave(smem_addr + ks * 256, 1024) : make_sdesc_mn_sw128(smem_addr + ks * 2048, 8192);
}

constexpr int kStages = 6;
constexpr uint32_t kStageBytes = 32768;
constexpr int kEpiWarps = 8;                       // two epilogue warps per TMEM lane quarter (column halves)
constexpr int kThreadsMlp = 32 * (2 + kEpiWarps);   // warp 0 producer, warp 1 MMA issuer, warps 2.. epilogue
constexpr int kThreadsWgrad = 192;
constexpr uint32_t kColD = 0, kColAhi = 256, kColAlo = 384;

// =============================================================================================
// weight packing
struct PackParams {
  const float* w[kLayers];
  uint8_t* wf;
  uint8_t* wb;
  MlpLayout L;
};

__global__ void __launch_bounds__(256) pack_weights_kernel(PackParams P) {
  DVD_PDL_ENTER();
  const int img = blockIdx.y / kLayers, l = blockIdx.y % kLayers;
  const int in = layer_in(P.L, l), out = layer_out(l);
  const float* __restrict__ W = P.w[l];
  const int rows = img == 0 ? rows_f(l) : rows_b(P.L, l);
  const int nkc = img == 0 ? nkc_f(P.L, l) : nkc_b(l);
  uint8_t* base = img == 0 ? P.wf + P.L.wf_off[l] : P.wb + P.L.wb_off[l];
  const int total = nkc * rows * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int kp = idx & 31, row = (idx >> 5) % rows, kc = (idx >> 5) / rows;
    int k = kp * 2, gk = kc * 64 + k;
    float v0, v1;
    if (img == 0) {  // B[n = out row][k = in]
      v0 = (row < out && gk < in) ? W[(size_t)row * in + gk] : 0.f;
      v1 = (row < out && gk + 1 < in) ? W[(size_t)row * in + gk + 1] : 0.f;
    } else {         // B[n = in row][k = out]
      v0 = (row < in && gk < out) ? W[(size_t)gk * in + row] : 0.f;
      v1 = (row < in && gk + 1 < out) ? W[(size_t)(gk + 1) * in + row] : 0.f;
    }
    uint32_t hi, lo;
    split2(v0, v1, hi, lo);
    uint8_t* blk = base + (size_t)(kc * 2) * rows * 128;
    uint32_t off = sw128_offset(row, k);
    *reinterpret_cast<uint32_t*>(blk + off) = hi;
    *reinterpret_cast<uint32_t*>(blk + (size_t)rows * 128 + off) = lo;
  }
}

// =============================================================================================
// periodic embedding (M1), feature order of SceneFlowFieldNet.forward: cat([t_emb, xyz_emb])
//   t_emb   = [t, cos(ft_k t) k<FT, sin(ft_k t) k<FT]                       (NT = 1 + 2 FT, if time dependent)
//   xyz_emb = [x, y, z, cos(f_k x), cos(f_k y), cos(f_k z) k<FX, sin(...) k<FX]
// Loops over the frequencies are deliberately NOT unrolled: a fully unrolled epilogue was ~0.5 MB of SASS and
// spent 30 % of its issue slots waiting for the instruction cache (profiles/r1_mlp_fwd_v1_*.txt).
template <int FX, int FT, bool TD>
struct Embed {
  static constexpr int NT = TD ? 1 + 2 * FT : 0;
  static constexpr int NIN = NT + 3 + 6 * FX;
  static constexpr int KPAD = (NIN + 15) / 16 * 16;
  // fast path: cos block starts at an even feature index, so two frequencies give exactly three packed words
  static constexpr bool kPaired = ((NT + 3) % 2 == 0) && (FX % 2 == 0);

  // feature j (runtime index)
  static __device__ __forceinline__ float feature(const dvd_mlp_cfg& c, int j, float t, float x, float y, float z) {
    float s, co;
    if (j < NT) {
      if (j == 0) return t;
      int k = j - 1;
      const bool is_sin = k >= FT;
      if (is_sin) k -= FT;
      fast_sincos(c.freq_t[k] * t, s, co);
      return is_sin ? s : co;
    }
    j -= NT;
    if (j < 3) return j == 0 ? x : (j == 1 ? y : z);
    j -= 3;
    if (j >= 6 * FX) return 0.f;
    const bool is_sin = j >= 3 * FX;
    if (is_sin) j -= 3 * FX;
    const int k = j / 3, d = j - 3 * k;
    fast_sincos(c.freq_xyz[k] * (d == 0 ? x : (d == 1 ? y : z)), s, co);
    return is_sin ? s : co;
  }
};

// shared-memory carve-up common to the chain kernels
struct ChainSmem {
  uint8_t* stage[kStages];
  float* bias;           // 5*256 + 16
  float* xchg;           // [128][4] partial point gradients exchanged between the two warps of a lane quarter
  uint64_t* w_full;      // [kStages]
  uint64_t* w_empty;     // [kStages]
  uint64_t* a_ready;
  uint64_t* d_ready;
  uint32_t* tmem_holder;
};
constexpr size_t kChainSmemBytes = 1024 + (size_t)kStages * kStageBytes + (5 * 256 + 16) * 4 + 128 * 4 * 4 + 256;

__device__ __forceinline__ ChainSmem carve(uint8_t* raw) {
  ChainSmem s;
  uint8_t* p = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  for (int i = 0; i < kStages; ++i) s.stage[i] = p + (size_t)i * kStageBytes;
  p += (size_t)kStages * kStageBytes;
  s.bias = reinterpret_cast<float*>(p);
  p += (5 * 256 + 16) * 4;
  s.xchg = reinterpret_cast<float*>(p);
  p += 128 * 4 * 4;
  s.w_full = reinterpret_cast<uint64_t*>(p);
  s.w_empty = s.w_full + kStages;
  s.a_ready = s.w_empty + kStages;
  s.d_ready = s.a_ready + 1;
  s.tmem_holder = reinterpret_cast<uint32_t*>(s.d_ready + 1);
  return s;
}

__device__ __forceinline__ void chain_setup(ChainSmem& s, int warp, int lane) {
  if (warp == 1) {
    tmem_alloc(s.tmem_holder, 512);
  } else if (warp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&s.w_full[i], 1);
      mbar_init(&s.w_empty[i], 1);
    }
    mbar_init(s.a_ready, 32 * kEpiWarps);
    mbar_init(s.d_ready, 1);
    fence_mbar_init();
  }
}

// =============================================================================================
// forward chain
struct FwdParams {
  const uint8_t* wf;
  const float* bias;
  const float* p0;
  const float* t0;
  float dt;
  int n_eval, n_acc;
  float* acc;
  float* s_steps;
  float* p_steps;
  uint8_t* save;
  long npx, hw;
  MlpLayout L;
  dvd_mlp_cfg cfg;
};

template <int FX, int FT, bool TD, bool SAVE>
__global__ void __launch_bounds__(kThreadsMlp, 1) mlp_chain_fwd_kernel(const __grid_constant__ FwdParams P) {
  using E = Embed<FX, FT, TD>;
  extern __shared__ uint8_t smem_raw[];
  ChainSmem S = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MlpLayout& L = P.L;
  chain_setup(S, warp, lane);
  DVD_PDL_ENTER();               // barriers / tensor memory are set up: now wait for the producer of the operands
  for (int i = threadIdx.x; i < 5 * 256 + 16; i += blockDim.x) S.bias[i] = P.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *S.tmem_holder;
  const long ntiles = L.ntiles;

  if (warp == 0) {
    // ===== weight producer: stream (layer, k-chunk, plane) blocks in MMA order =====
    if (lane == 0) {
      uint32_t it = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int e = 0; e < P.n_eval; ++e) {
          for (int l = 0; l < kLayers; ++l) {
            const uint32_t bytes = (uint32_t)rows_f(l) * 128u;
            const uint8_t* src = P.wf + L.wf_off[l];
            const int n = nkc_f(L, l) * 2;
            for (int j = 0; j < n; ++j, ++it) {
              const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
              mbar_wait(&S.w_empty[s], ph ^ 1u);
              mbar_arrive_expect_tx(&S.w_full[s], bytes);
              bulk_g2s(S.stage[s], src + (size_t)j * bytes, bytes, &S.w_full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, a_phase = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int e = 0; e < P.n_eval; ++e) {
          for (int l = 0; l < kLayers; ++l) {
            mbar_wait(S.a_ready, a_phase);
            a_phase ^= 1u;
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, rows_f(l));
            const int kslices = (l == 0 ? L.kpad0 : kWidth) / 16;
            uint32_t accum = 0;
            for (int kc = 0; kc < nkc_f(L, l); ++kc) {
              const int nks = min(4, kslices - kc * 4);
              // hi plane of the weights: A_hi*W_hi + A_lo*W_hi
              uint32_t s = it % kStages, ph = (it / kStages) & 1u;
              mbar_wait(&S.w_full[s], ph);
              tc_fence_after();
              uint32_t sb = smem_u32(S.stage[s]);
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t bd = make_sdesc_k_sw128(sb + ks * 32);
                const uint32_t acol = (uint32_t)(kc * 4 + ks) * 8u;
                umma_ts(tmem + kColD, tmem + kColAhi + acol, bd, idesc, accum);
                accum = 1;
                umma_ts(tmem + kColD, tmem + kColAlo + acol, bd, idesc, 1);
              }
              umma_commit(&S.w_empty[s]);
              ++it;
              // lo plane: A_hi*W_lo
              s = it % kStages; ph = (it / kStages) & 1u;
              mbar_wait(&S.w_full[s], ph);
              tc_fence_after();
              sb = smem_u32(S.stage[s]);
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t bd = make_sdesc_k_sw128(sb + ks * 32);
                const uint32_t acol = (uint32_t)(kc * 4 + ks) * 8u;
                umma_ts(tmem + kColD, tmem + kColAhi + acol, bd, idesc, 1);
              }
              umma_commit(&S.w_empty[s]);
              ++it;
            }
            umma_commit(S.d_ready);
          }
        }
      }
    }
  } else {
    // ===== epilogue warps: one pixel per thread, two warps (column halves) per lane quarter =====
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tD = tmem + kColD + lane_base, tAhi = tmem + kColAhi + lane_base, tAlo = tmem + kColAlo + lane_base;
    uint32_t d_phase = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long g = tile * kTileM + row;
      const bool valid = g < P.npx;
      const long b = valid ? g / P.hw : 0, i = valid ? g % P.hw : 0;
      const size_t pidx = (size_t)b * 3 * P.hw + i;
      float px = 0.f, py = 0.f, pz = 0.f, t = 0.f;
      if (valid) {
        px = P.p0[pidx]; py = P.p0[pidx + P.hw]; pz = P.p0[pidx + 2 * P.hw];
        if (TD) t = P.t0[(size_t)b * P.hw + i];
      }
      float ax = 0.f, ay = 0.f, az = 0.f;
      const long chunk = tile * 2 + (row >> 6);
      const uint32_t kq = (uint32_t)(row & 63);
      for (int e = 0; e < P.n_eval; ++e) {
        uint8_t* save_e = SAVE ? P.save + (size_t)e * L.save_total : nullptr;
        if (SAVE && valid && hsel == 0) {
          float* ps = P.p_steps + (size_t)e * P.npx * 3 + pidx;
          ps[0] = px; ps[P.hw] = py; ps[2 * P.hw] = pz;
        }
        // ---- embedding -> A operand (layer 0 input)
        {
          uint8_t* x0_hi = SAVE ? save_e + L.xs_off[0] + (size_t)chunk * blk_bytes(E::KPAD) : nullptr;
          uint8_t* x0_lo = SAVE ? x0_hi + (size_t)L.nq * blk_bytes(E::KPAD) : nullptr;
          auto emit = [&](int w, float f0, float f1) {
            uint32_t hi, lo;
            split2(f0, f1, hi, lo);
            tmem_st1(tAhi + w, hi);
            tmem_st1(tAlo + w, lo);
            if (SAVE) {
              const uint32_t o = act_offset(2 * w, kq);
              *reinterpret_cast<uint32_t*>(x0_hi + o) = hi;
              if (kSavePlanes == 2) *reinterpret_cast<uint32_t*>(x0_lo + o) = lo;
            }
          };
          if (E::kPaired) {
            constexpr int W0 = (E::NT + 3) / 2;           // first word of the cos block
            constexpr int KSPLIT = (FX / 2) * 3 / 8;      // warp 0 of the pair: t-part + first frequencies
            if (hsel == 0) {
#pragma unroll 1
              for (int w = 0; w < W0; ++w)
                emit(w, E::feature(P.cfg, 2 * w, t, px, py, pz), E::feature(P.cfg, 2 * w + 1, t, px, py, pz));
            }
#pragma unroll 1
            for (int kp = (hsel == 0 ? 0 : KSPLIT); kp < (hsel == 0 ? KSPLIT : FX / 2); ++kp) {
              const float f0 = P.cfg.freq_xyz[2 * kp], f1 = P.cfg.freq_xyz[2 * kp + 1];
              float s0x, c0x, s0y, c0y, s0z, c0z, s1x, c1x, s1y, c1y, s1z, c1z;
              fast_sincos(f0 * px, s0x, c0x); fast_sincos(f0 * py, s0y, c0y); fast_sincos(f0 * pz, s0z, c0z);
              fast_sincos(f1 * px, s1x, c1x); fast_sincos(f1 * py, s1y, c1y); fast_sincos(f1 * pz, s1z, c1z);
              const int wc = W0 + 3 * kp, ws = W0 + 3 * (FX / 2) + 3 * kp;
              emit(wc, c0x, c0y); emit(wc + 1, c0z, c1x); emit(wc + 2, c1y, c1z);
              emit(ws, s0x, s0y); emit(ws + 1, s0z, s1x); emit(ws + 2, s1y, s1z);
            }
            if (hsel == 1) {
#pragma unroll 1
              for (int w = W0 + 3 * FX; w < E::KPAD / 2; ++w) emit(w, 0.f, 0.f);
            }
          } else {
            constexpr int WH = E::KPAD / 4;
#pragma unroll 1
            for (int w = hsel * WH; w < (hsel == 0 ? WH : E::KPAD / 2); ++w)
              emit(w, E::feature(P.cfg, 2 * w, t, px, py, pz), E::feature(P.cfg, 2 * w + 1, t, px, py, pz));
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(S.a_ready);
        // ---- hidden layers
        for (int l = 0; l < 5; ++l) {
          mbar_wait(S.d_ready, d_phase);
          d_phase ^= 1u;
          tc_fence_after();
          const float* bl = S.bias + l * 256;
          uint8_t* x_hi = SAVE ? save_e + L.xs_off[l + 1] + (size_t)chunk * blk_bytes(kWidth) : nullptr;
          uint8_t* x_lo = SAVE ? x_hi + (size_t)L.nq * blk_bytes(kWidth) : nullptr;
          uint32_t* maskp = SAVE ? reinterpret_cast<uint32_t*>(save_e + L.mask_off +
                                                              (((size_t)l * L.ntiles + tile) * kTileM + row) * 32)
                                 : nullptr;
#pragma unroll 1
          for (int cb = hsel * 4; cb < hsel * 4 + 4; ++cb) {
            const int c0 = cb * 32;
            uint32_t r[32];
            tmem_ld32(tD + c0, r);
            tmem_ld_wait();
            uint32_t hi[16], lo[16], mb = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 bb = *reinterpret_cast<const float2*>(bl + c0 + 2 * j);
              const float y0 = __uint_as_float(r[2 * j]) + bb.x;
              const float y1 = __uint_as_float(r[2 * j + 1]) + bb.y;
              if (SAVE) {
                mb |= (y0 > 0.f ? 1u : 0u) << (2 * j);
                mb |= (y1 > 0.f ? 1u : 0u) << (2 * j + 1);
              }
              // LeakyReLU(0.2): max(y, 0.2 y)
              split2(fmaxf(y0, 0.2f * y0), fmaxf(y1, 0.2f * y1), hi[j], lo[j]);
            }
            tmem_st16(tAhi + c0 / 2, hi);
            tmem_st16(tAlo + c0 / 2, lo);
            if (SAVE) {   // 32 channels = four 16-byte chunks of this pixel's 128-byte rows
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const uint32_t o = act_offset(c0 + 8 * h, kq);
                st_global_v4(x_hi + o, hi[4 * h], hi[4 * h + 1], hi[4 * h + 2], hi[4 * h + 3]);
                if (kSavePlanes == 2) st_global_v4(x_lo + o, lo[4 * h], lo[4 * h + 1], lo[4 * h + 2], lo[4 * h + 3]);
              }
              maskp[cb] = mb;
            }
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(S.a_ready);
        }
        // ---- output layer (N = 16 padded, 3 used): s = (W5 x + b5) / sf_mag_div ; Euler update
        mbar_wait(S.d_ready, d_phase);
        d_phase ^= 1u;
        tc_fence_after();
        {
          uint32_t r[16];
          tmem_ld16(tD, r);
          tmem_ld_wait();
          const float* b5 = S.bias + 5 * 256;
          const float sx = (__uint_as_float(r[0]) + b5[0]) / P.cfg.sf_mag_div;
          const float sy = (__uint_as_float(r[1]) + b5[1]) / P.cfg.sf_mag_div;
          const float sz = (__uint_as_float(r[2]) + b5[2]) / P.cfg.sf_mag_div;
          if (valid && P.s_steps && hsel == 0) {
            float* ss = P.s_steps + (size_t)e * P.npx * 3 + pidx;
            ss[0] = sx; ss[P.hw] = sy; ss[2 * P.hw] = sz;
          }
          if (e < P.n_acc) { ax += sx; ay += sy; az += sz; }
          px += sx; py += sy; pz += sz;
          t += P.dt;
        }
      }
      if (valid && P.acc && hsel == 0) {
        float* ao = P.acc + pidx;
        ao[0] = ax; ao[P.hw] = ay; ao[2 * P.hw] = az;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================
// dgrad chain of one eval
struct DgradParams {
  const uint8_t* wb;
  const float* p_e;
  const float* t0;
  float dt;
  int e, use_g_acc;
  const float* g_acc;
  const float* g_step;
  const float* a_in;
  float* a_out;
  const uint8_t* save_e;
  uint8_t* dy;
  float* g_bias5;
  long npx, hw;
  MlpLayout L;
  dvd_mlp_cfg cfg;
};

template <int FX, int FT, bool TD>
__global__ void __launch_bounds__(kThreadsMlp, 1) mlp_dgrad_kernel(const __grid_constant__ DgradParams P) {
  using E = Embed<FX, FT, TD>;
  extern __shared__ uint8_t smem_raw[];
  ChainSmem S = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MlpLayout& L = P.L;
  chain_setup(S, warp, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *S.tmem_holder;
  const long ntiles = L.ntiles;
  DVD_PDL_ENTER();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = 5; l >= 0; --l) {
          const uint32_t bytes = (uint32_t)rows_b(L, l) * 128u;
          const uint8_t* src = P.wb + L.wb_off[l];
          const int n = nkc_b(l) * 2;
          for (int j = 0; j < n; ++j, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(&S.w_empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&S.w_full[s], bytes);
            bulk_g2s(S.stage[s], src + (size_t)j * bytes, bytes, &S.w_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, a_phase = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = 5; l >= 0; --l) {
          mbar_wait(S.a_ready, a_phase);
          a_phase ^= 1u;
          tc_fence_after();
          const uint32_t idesc = make_idesc_bf16(128, rows_b(L, l));
          const int kslices = (l == 5) ? 1 : kWidth / 16;
          uint32_t accum = 0;
          for (int kc = 0; kc < nkc_b(l); ++kc) {
            const int nks = min(4, kslices - kc * 4);
            uint32_t s = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(&S.w_full[s], ph);
            tc_fence_after();
            uint32_t sb = smem_u32(S.stage[s]);
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t bd = make_sdesc_k_sw128(sb + ks * 32);
              const uint32_t acol = (uint32_t)(kc * 4 + ks) * 8u;
              umma_ts(tmem + kColD, tmem + kColAhi + acol, bd, idesc, accum);
              accum = 1;
              umma_ts(tmem + kColD, tmem + kColAlo + acol, bd, idesc, 1);
            }
            umma_commit(&S.w_empty[s]);
            ++it;
            s = it % kStages; ph = (it / kStages) & 1u;
            mbar_wait(&S.w_full[s], ph);
            tc_fence_after();
            sb = smem_u32(S.stage[s]);
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t bd = make_sdesc_k_sw128(sb + ks * 32);
              const uint32_t acol = (uint32_t)(kc * 4 + ks) * 8u;
              umma_ts(tmem + kColD, tmem + kColAhi + acol, bd, idesc, 1);
            }
            umma_commit(&S.w_empty[s]);
            ++it;
          }
          umma_commit(S.d_ready);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tD = tmem + kColD + lane_base, tAhi = tmem + kColAhi + lane_base, tAlo = tmem + kColAlo + lane_base;
    uint32_t d_phase = 0;
    float b5x = 0.f, b5y = 0.f, b5z = 0.f;  // per-thread partial of the output-bias gradient
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long g = tile * kTileM + row;
      const bool valid = g < P.npx;
      const long b = valid ? g / P.hw : 0, i = valid ? g % P.hw : 0;
      const size_t pidx = (size_t)b * 3 * P.hw + i;
      const long chunk = tile * 2 + (row >> 6);
      const uint32_t kq = (uint32_t)(row & 63);
      float gx = 0.f, gy = 0.f, gz = 0.f, ain_x = 0.f, ain_y = 0.f, ain_z = 0.f;
      if (valid) {
        if (P.a_in) { ain_x = P.a_in[pidx]; ain_y = P.a_in[pidx + P.hw]; ain_z = P.a_in[pidx + 2 * P.hw]; }
        gx = ain_x; gy = ain_y; gz = ain_z;
        if (P.use_g_acc && P.g_acc) { gx += P.g_acc[pidx]; gy += P.g_acc[pidx + P.hw]; gz += P.g_acc[pidx + 2 * P.hw]; }
        if (P.g_step) { gx += P.g_step[pidx]; gy += P.g_step[pidx + P.hw]; gz += P.g_step[pidx + 2 * P.hw]; }
      }
      // dY_5 = gs / sf_mag_div (3 of 16 padded columns)
      const float d5x = gx / P.cfg.sf_mag_div, d5y = gy / P.cfg.sf_mag_div, d5z = gz / P.cfg.sf_mag_div;
      if (hsel == 0) {
        b5x += d5x; b5y += d5y; b5z += d5z;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { hi[j] = 0; lo[j] = 0; }
        split2(d5x, d5y, hi[0], lo[0]);
        split2(d5z, 0.f, hi[1], lo[1]);
        tmem_st8(tAhi, hi);
        tmem_st8(tAlo, lo);
        uint8_t* y_hi = P.dy + L.dy_off[5] + (size_t)chunk * blk_bytes(16);
        uint8_t* y_lo = y_hi + (size_t)L.nq * blk_bytes(16);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t o = act_offset(8 * h, kq);
          st_global_v4(y_hi + o, hi[4 * h], hi[4 * h + 1], hi[4 * h + 2], hi[4 * h + 3]);
          if (kSavePlanes == 2) st_global_v4(y_lo + o, lo[4 * h], lo[4 * h + 1], lo[4 * h + 2], lo[4 * h + 3]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(S.a_ready);
      // MMA(l) produces dX_l (input gradient of layer l); l = 5..1 feed dY_{l-1}
      for (int l = 5; l >= 1; --l) {
        mbar_wait(S.d_ready, d_phase);
        d_phase ^= 1u;
        tc_fence_after();
        const uint32_t* maskp = reinterpret_cast<const uint32_t*>(P.save_e + L.mask_off +
                                                                  (((size_t)(l - 1) * L.ntiles + tile) * kTileM + row) * 32);
        uint8_t* y_hi = P.dy + L.dy_off[l - 1] + (size_t)chunk * blk_bytes(kWidth);
        uint8_t* y_lo = y_hi + (size_t)L.nq * blk_bytes(kWidth);
#pragma unroll 1
        for (int cb = hsel * 4; cb < hsel * 4 + 4; ++cb) {
          const int c0 = cb * 32;
          uint32_t r[32];
          tmem_ld32(tD + c0, r);
          const uint32_t mb = maskp[cb];
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v0 = __uint_as_float(r[2 * j]), v1 = __uint_as_float(r[2 * j + 1]);
            v0 *= ((mb >> (2 * j)) & 1u) ? 1.0f : 0.2f;
            v1 *= ((mb >> (2 * j + 1)) & 1u) ? 1.0f : 0.2f;
            split2(v0, v1, hi[j], lo[j]);
          }
          tmem_st16(tAhi + c0 / 2, hi);
          tmem_st16(tAlo + c0 / 2, lo);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const uint32_t o = act_offset(c0 + 8 * h, kq);
            st_global_v4(y_hi + o, hi[4 * h], hi[4 * h + 1], hi[4 * h + 2], hi[4 * h + 3]);
            if (kSavePlanes == 2) st_global_v4(y_lo + o, lo[4 * h], lo[4 * h + 1], lo[4 * h + 2], lo[4 * h + 3]);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(S.a_ready);
      }
      // MMA(0) -> gradient w.r.t. the embedding; contract with d(embed)/d(xyz)
      mbar_wait(S.d_ready, d_phase);
      d_phase ^= 1u;
      tc_fence_after();
      {
        float px = 0.f, py = 0.f, pz = 0.f, t = 0.f;
        if (valid) {
          px = P.p_e[pidx]; py = P.p_e[pidx + P.hw]; pz = P.p_e[pidx + 2 * P.hw];
          if (TD) {
            t = P.t0[(size_t)b * P.hw + i];
            for (int k = 0; k < P.e; ++k) t += P.dt;   // same fp32 accumulation as the forward chain
          }
        }
        float gpx = 0.f, gpy = 0.f, gpz = 0.f;
        if (hsel == 0) {
          uint32_t gx, gy, gz;
          tmem_ld1(tD + E::NT, gx); tmem_ld1(tD + E::NT + 1, gy); tmem_ld1(tD + E::NT + 2, gz);
          tmem_ld_wait();
          gpx = __uint_as_float(gx); gpy = __uint_as_float(gy); gpz = __uint_as_float(gz);
        }
        constexpr int CB = E::NT + 3, SB = E::NT + 3 + 3 * FX;   // first cos / sin feature
#pragma unroll 1
        for (int k = hsel * (FX / 2); k < (hsel == 0 ? FX / 2 : FX); ++k) {
          uint32_t gc[3], gs[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) { tmem_ld1(tD + CB + 3 * k + d, gc[d]); tmem_ld1(tD + SB + 3 * k + d, gs[d]); }
          const float f = P.cfg.freq_xyz[k];
          float sx, cx, sy, cy, sz, cz;
          fast_sincos(f * px, sx, cx); fast_sincos(f * py, sy, cy); fast_sincos(f * pz, sz, cz);
          tmem_ld_wait();
          // d/dx cos(f x) = -f sin(f x) ; d/dx sin(f x) = f cos(f x)
          gpx += f * (__uint_as_float(gs[0]) * cx - __uint_as_float(gc[0]) * sx);
          gpy += f * (__uint_as_float(gs[1]) * cy - __uint_as_float(gc[1]) * sy);
          gpz += f * (__uint_as_float(gs[2]) * cz - __uint_as_float(gc[2]) * sz);
        }
        // combine the two frequency halves of this pixel through shared memory (pair-wise named barrier)
        if (hsel == 1) {
          S.xchg[row * 4 + 0] = gpx; S.xchg[row * 4 + 1] = gpy; S.xchg[row * 4 + 2] = gpz;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        if (hsel == 0) {
          gpx += S.xchg[row * 4 + 0]; gpy += S.xchg[row * 4 + 1]; gpz += S.xchg[row * 4 + 2];
        }
        const float gp[3] = {gpx, gpy, gpz};
        if (valid && P.a_out && hsel == 0) {
          P.a_out[pidx] = ain_x + gp[0];
          P.a_out[pidx + P.hw] = ain_y + gp[1];
          P.a_out[pidx + 2 * P.hw] = ain_z + gp[2];
        }
      }
    }
    // output-layer bias gradient: warp reduce, one atomic per warp
    b5x = warp_sum(b5x); b5y = warp_sum(b5y); b5z = warp_sum(b5z);
    if (lane == 0 && P.g_bias5 && hsel == 0) {
      atomicAdd(P.g_bias5 + 0, b5x);
      atomicAdd(P.g_bias5 + 1, b5y);
      atomicAdd(P.g_bias5 + 2, b5z);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================
// wgrad: D[128, n] += A_blk[128 ch x 64 px] * B_blk[n ch x 64 px]^T over a range of pixel chunks
// (SS mode, both operands MN-major: channels contiguous, K = pixels)
struct WgradJob {
  const uint8_t* a_hi;
  const uint8_t* a_lo;
  const uint8_t* b_hi;
  const uint8_t* b_lo;
  uint32_t a_blk, a_row_off, b_blk;
  int n, ld, transposed, m_off, m_valid, n_valid;
  float* out;
  float* bias_out;
};
struct WgradParams {
  WgradJob job[12];
  long nq;
};
constexpr int kWgStages = kSavePlanes == 2 ? 2 : 4;
constexpr uint32_t kWgStageBytes = kSavePlanes * (16384 + 32768);      // per plane: A 128 ch x 64 px, B up to 256 ch x 64 px (bf16)
constexpr size_t kWgradSmemBytes = 1024 + (size_t)kWgStages * kWgStageBytes + 8192 + 256;

__global__ void __launch_bounds__(kThreadsWgrad, 1) mlp_wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage[kWgStages];
  for (int i = 0; i < kWgStages; ++i) stage[i] = base + (size_t)i * kWgStageBytes;
  uint8_t* ones = base + (size_t)kWgStages * kWgStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ones + 8192);
  uint64_t* empty = full + kWgStages;
  uint64_t* done = empty + kWgStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WgradJob& J = P.job[blockIdx.x];
  const long per = (P.nq + gridDim.y - 1) / gridDim.y;
  const long q0 = (long)blockIdx.y * per, q1 = min(P.nq, q0 + per);

  if (warp == 1) {
    tmem_alloc(tmem_holder, 512);
  } else if (warp == 0 && lane == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  // "ones" operand [16 ch x 64 px], MN-major: channel 0 = 1.0 (bf16), channels 1..15 = 0 -> bias column sums
  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0u;
  __syncthreads();
  if (threadIdx.x < 64) *reinterpret_cast<uint16_t*>(ones + act_offset(0, threadIdx.x)) = (uint16_t)0x3F80u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const bool have_work = q1 > q0;
  DVD_PDL_ENTER();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t a_bytes = 16384, b_bytes = blk_bytes(J.n);
      for (long qq = q0; qq < q1; ++qq, ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], kSavePlanes * (a_bytes + b_bytes));
        uint8_t* d = stage[s];
        bulk_g2s(d, J.a_hi + (size_t)qq * J.a_blk + J.a_row_off, a_bytes, &full[s]);
        bulk_g2s(d + 16384, J.b_hi + (size_t)qq * J.b_blk, b_bytes, &full[s]);
        if (kSavePlanes == 2) {
          bulk_g2s(d + 49152, J.a_lo + (size_t)qq * J.a_blk + J.a_row_off, a_bytes, &full[s]);
          bulk_g2s(d + 65536, J.b_lo + (size_t)qq * J.b_blk, b_bytes, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && have_work) {
      uint32_t it = 0, accum = 0;
      const uint32_t idesc = make_idesc_bf16(128, J.n, true, true), idesc_b = make_idesc_bf16(128, 16, true, true);
      const uint32_t so = smem_u32(ones);
      for (long qq = q0; qq < q1; ++qq, ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(stage[s]), sb_hi = sa_hi + 16384, sa_lo = sa_hi + 49152, sb_lo = sa_hi + 65536;
        for (int ks = 0; ks < 4; ++ks) {   // 16 pixels (K rows of 128 B) per MMA
          const uint64_t ah = act_desc(sa_hi, ks), bh = act_desc(sb_hi, ks);
          umma_ss(tmem + 0, ah, bh, idesc, accum);
          if (kSavePlanes == 2) {
            umma_ss(tmem + 0, act_desc(sa_lo, ks), bh, idesc, 1);
            umma_ss(tmem + 0, ah, act_desc(sb_lo, ks), idesc, 1);
          }
          if (J.bias_out) {
            const uint64_t od = act_desc(so, ks);
            umma_ss(tmem + 256, ah, od, idesc_b, accum);
            if (kSavePlanes == 2) umma_ss(tmem + 256, act_desc(sa_lo, ks), od, idesc_b, 1);
          }
          accum = 1;
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else if (have_work) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    mbar_wait(done, 0);
    tc_fence_after();
    const int m = J.m_off + row;
    for (int c0 = 0; c0 < J.n; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + c0, r);
      tmem_ld_wait();
      if (m < J.m_valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = c0 + j;
          if (n < J.n_valid) {
            float* dst = J.transposed ? J.out + (size_t)n * J.ld + m : J.out + (size_t)m * J.ld + n;
            atomicAdd(dst, __uint_as_float(r[j]));
          }
        }
      }
    }
    if (J.bias_out) {
      uint32_t r[16];
      tmem_ld16(tmem + lane_base + 256, r);
      tmem_ld_wait();
      if (m < J.m_valid) atomicAdd(J.bias_out + m, __uint_as_float(r[0]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================
// acceleration regulariser on (s0, s1)
__global__ void __launch_bounds__(256) acc_reg_kernel(const float* __restrict__ s0, const float* __restrict__ s1, float c,
                                                      float* __restrict__ g0, float* __restrict__ g1,
                                                      float* __restrict__ partials, long n) {
  DVD_PDL_ENTER();
  __shared__ float red[8];
  float acc = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float d = s1[i] - s0[i];
    acc += fabsf(d);
    float s = (d > 0.f) ? c : ((d < 0.f) ? -c : 0.f);
    if (g0) g0[i] = -s;
    if (g1) g1[i] = s;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) acc_reg_final_kernel(const float* __restrict__ partials, int nb, float scale,
                                                            float* __restrict__ out) {
  DVD_PDL_ENTER();
  __shared__ double red[8];
  double a = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) a += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = (float)(t * scale);
  }
}

// =============================================================================================
static int check_cfg_supported(const dvd_mlp_cfg* cfg, int* variant) {
  DVD_ARG_CHECK(cfg != nullptr, "null mlp cfg");
  if (cfg->time_dependent && cfg->n_freq_xyz == 16 && cfg->n_freq_t == 16) { *variant = 0; return 0; }
  if (!cfg->time_dependent && cfg->n_freq_xyz == 16) { *variant = 1; return 0; }
  set_error("unsupported scene-flow MLP configuration (n_freq_xyz=%d n_freq_t=%d time_dependent=%d): the "
            "tensor-core kernels are instantiated for n_freq_xyz=16 with (time_dependent, n_freq_t=16) or "
            "(not time_dependent)", cfg->n_freq_xyz, cfg->n_freq_t, cfg->time_dependent);
  return -2;
}

static int chain_grid(long ntiles) {
  int sms = num_sms();
  return (int)(ntiles < sms ? ntiles : sms);
}

}  // namespace dvd

using namespace dvd;

extern "C" size_t dvd_mlp_packed_weights_bytes(const dvd_mlp_cfg* cfg) {
  if (!cfg) return 0;
  MlpLayout L = make_layout(*cfg, 128);
  return L.wf_total > L.wb_total ? L.wf_total : L.wb_total;
}
extern "C" size_t dvd_mlp_save_bytes_per_eval(const dvd_mlp_cfg* cfg, long npx) {
  if (!cfg || npx <= 0) return 0;
  return make_layout(*cfg, npx).save_total;
}
extern "C" size_t dvd_mlp_dy_bytes(const dvd_mlp_cfg* cfg, long npx) {
  if (!cfg || npx <= 0) return 0;
  return make_layout(*cfg, npx).dy_total;
}

extern "C" int dvd_mlp_pack_weights(const dvd_mlp_cfg* cfg, const float* const* w, void* packed_fwd, void* packed_bwd,
                                    void* stream) {
  int variant;
  if (int e = check_cfg_supported(cfg, &variant)) return e;
  DVD_ARG_CHECK(w && packed_fwd && packed_bwd, "null pointer");
  PackParams P;
  for (int l = 0; l < kLayers; ++l) {
    DVD_ARG_CHECK(w[l] != nullptr, "null weight pointer for layer %d", l);
    P.w[l] = w[l];
  }
  P.wf = (uint8_t*)packed_fwd;
  P.wb = (uint8_t*)packed_bwd;
  P.L = make_layout(*cfg, 128);
  dvd::launch(pack_weights_kernel, dim3(32, 2 * kLayers), 256, 0, (cudaStream_t)stream, P);
  DVD_CUDA_LAUNCH_CHECK("mlp_pack_weights");
  return 0;
}

template <int FX, int FT, bool TD>
static int launch_fwd(const FwdParams& P, bool save, cudaStream_t st) {
  int grid = chain_grid(P.L.ntiles);
  if (save) {
    DVD_CUDA_CALL(cudaFuncSetAttribute(mlp_chain_fwd_kernel<FX, FT, TD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kChainSmemBytes));
    dvd::launch(mlp_chain_fwd_kernel<FX, FT, TD, true>, grid, kThreadsMlp, kChainSmemBytes, st, P);
  } else {
    DVD_CUDA_CALL(cudaFuncSetAttribute(mlp_chain_fwd_kernel<FX, FT, TD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kChainSmemBytes));
    dvd::launch(mlp_chain_fwd_kernel<FX, FT, TD, false>, grid, kThreadsMlp, kChainSmemBytes, st, P);
  }
  DVD_CUDA_LAUNCH_CHECK("mlp_chain_fwd");
  return 0;
}

extern "C" int dvd_mlp_chain_fwd(const dvd_mlp_cfg* cfg, const void* packed_fwd, const float* bias, const float* p0,
                                 const float* t0, float dt, int n_eval, int n_acc, float* acc, float* s_steps,
                                 float* p_steps, void* save, long npx, long hw, void* stream) {
  int variant;
  if (int e = check_cfg_supported(cfg, &variant)) return e;
  DVD_ARG_CHECK(packed_fwd && bias && p0, "null pointer");
  DVD_ARG_CHECK(!cfg->time_dependent || t0, "t0 required for a time-dependent field");
  DVD_ARG_CHECK(n_eval >= 1 && n_eval <= 64 && n_acc >= 0 && n_acc <= n_eval, "bad n_eval=%d / n_acc=%d", n_eval, n_acc);
  DVD_ARG_CHECK(npx > 0 && hw > 0 && npx % hw == 0, "npx must be a multiple of hw");
  DVD_ARG_CHECK(!save || p_steps, "p_steps required when save is given");
  DVD_ARG_CHECK(cfg->sf_mag_div != 0.f, "sf_mag_div must be non-zero");
  FwdParams P;
  P.wf = (const uint8_t*)packed_fwd; P.bias = bias; P.p0 = p0; P.t0 = t0; P.dt = dt;
  P.n_eval = n_eval; P.n_acc = n_acc; P.acc = acc; P.s_steps = s_steps; P.p_steps = p_steps;
  P.save = (uint8_t*)save; P.npx = npx; P.hw = hw; P.L = make_layout(*cfg, npx); P.cfg = *cfg;
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 0) return launch_fwd<16, 16, true>(P, save != nullptr, st);
  return launch_fwd<16, 0, false>(P, save != nullptr, st);
}

extern "C" int dvd_mlp_dgrad(const dvd_mlp_cfg* cfg, const void* packed_bwd, const float* p_e, const float* t0, float dt,
                             int e, int use_g_acc, const float* g_acc, const float* g_step, const float* a_in,
                             float* a_out, const void* save_e, void* dy_scratch, float* g_bias5, long npx, long hw,
                             void* stream) {
  int variant;
  if (int er = check_cfg_supported(cfg, &variant)) return er;
  DVD_ARG_CHECK(packed_bwd && p_e && save_e && dy_scratch, "null pointer");
  DVD_ARG_CHECK(!cfg->time_dependent || t0, "t0 required for a time-dependent field");
  DVD_ARG_CHECK(e >= 0 && e < 64, "bad eval index");
  DVD_ARG_CHECK(npx > 0 && hw > 0 && npx % hw == 0, "npx must be a multiple of hw");
  DgradParams P;
  P.wb = (const uint8_t*)packed_bwd; P.p_e = p_e; P.t0 = t0; P.dt = dt; P.e = e; P.use_g_acc = use_g_acc;
  P.g_acc = g_acc; P.g_step = g_step; P.a_in = a_in; P.a_out = a_out; P.save_e = (const uint8_t*)save_e;
  P.dy = (uint8_t*)dy_scratch; P.g_bias5 = g_bias5; P.npx = npx; P.hw = hw; P.L = make_layout(*cfg, npx); P.cfg = *cfg;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = chain_grid(P.L.ntiles);
  if (variant == 0) {
    DVD_CUDA_CALL(cudaFuncSetAttribute(mlp_dgrad_kernel<16, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kChainSmemBytes));
    dvd::launch(mlp_dgrad_kernel<16, 16, true>, grid, kThreadsMlp, kChainSmemBytes, st, P);
  } else {
    DVD_CUDA_CALL(cudaFuncSetAttribute(mlp_dgrad_kernel<16, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kChainSmemBytes));
    dvd::launch(mlp_dgrad_kernel<16, 0, false>, grid, kThreadsMlp, kChainSmemBytes, st, P);
  }
  DVD_CUDA_LAUNCH_CHECK("mlp_dgrad");
  return 0;
}

extern "C" int dvd_mlp_wgrad(const dvd_mlp_cfg* cfg, const void* save_e, const void* dy_scratch, float* const* g_w,
                             float* const* g_b, long npx, void* stream) {
  int variant;
  if (int er = check_cfg_supported(cfg, &variant)) return er;
  DVD_ARG_CHECK(save_e && dy_scratch && g_w && g_b, "null pointer");
  DVD_ARG_CHECK(npx > 0, "npx must be positive");
  MlpLayout L = make_layout(*cfg, npx);
  const uint8_t* sv = (const uint8_t*)save_e;
  const uint8_t* dy = (const uint8_t*)dy_scratch;
  WgradParams P;
  P.nq = L.nq;
  int nj = 0;
  for (int l = 0; l < kLayers; ++l) {
    DVD_ARG_CHECK(g_w[l] != nullptr, "null g_w[%d]", l);
    const size_t x_plane = (size_t)L.nq * blk_bytes(rows_x(L, l)), y_plane = (size_t)L.nq * blk_bytes(rows_dy(l));
    for (int mh = 0; mh < 2; ++mh) {
      WgradJob& J = P.job[nj++];
      if (l < 5) {
        DVD_ARG_CHECK(g_b[l] != nullptr, "null g_b[%d]", l);
        J.a_hi = dy + L.dy_off[l]; J.a_lo = J.a_hi + y_plane; J.a_blk = blk_bytes(kWidth); J.a_row_off = mh * 2 * 8192;
        J.b_hi = sv + L.xs_off[l]; J.b_lo = J.b_hi + x_plane; J.b_blk = blk_bytes(rows_x(L, l)); J.n = rows_x(L, l);
        J.out = g_w[l]; J.ld = layer_in(L, l); J.transposed = 0; J.m_off = mh * 128; J.m_valid = kWidth;
        J.n_valid = layer_in(L, l); J.bias_out = g_b[l];
      } else {
        J.a_hi = sv + L.xs_off[5]; J.a_lo = J.a_hi + x_plane; J.a_blk = blk_bytes(kWidth); J.a_row_off = mh * 2 * 8192;
        J.b_hi = dy + L.dy_off[5]; J.b_lo = J.b_hi + y_plane; J.b_blk = blk_bytes(16); J.n = 16;
        J.out = g_w[5]; J.ld = kWidth; J.transposed = 1; J.m_off = mh * 128; J.m_valid = kWidth; J.n_valid = 3;
        J.bias_out = nullptr;
      }
    }
  }
  int ksplit = num_sms() / nj;
  if (ksplit < 1) ksplit = 1;
  if ((long)ksplit > L.nq) ksplit = (int)L.nq;
  DVD_CUDA_CALL(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmemBytes));
  dvd::launch(mlp_wgrad_kernel, dim3(nj, ksplit), kThreadsWgrad, kWgradSmemBytes, (cudaStream_t)stream, P);
  DVD_CUDA_LAUNCH_CHECK("mlp_wgrad");
  return 0;
}

extern "C" int dvd_acc_reg(const float* s0, const float* s1, float acc_mul, float gscale, float* g_s0, float* g_s1,
                           float* partials, float* loss_out, long numel, void* stream) {
  DVD_ARG_CHECK(s0 && s1 && partials && loss_out && numel > 0, "bad arguments");
  const float inv = 1.0f / ((float)numel + 1e-6f);
  int nb = (int)((numel + 255) / 256);
  if (nb > 1024) nb = 1024;
  dvd::launch(acc_reg_kernel, nb, 256, 0, (cudaStream_t)stream, s0, s1, acc_mul * inv * gscale, g_s0, g_s1, partials, numel);
  DVD_CUDA_LAUNCH_CHECK("acc_reg");
  dvd::launch(acc_reg_final_kernel, 1, 256, 0, (cudaStream_t)stream, partials, nb, acc_mul * inv, loss_out);
  DVD_CUDA_LAUNCH_CHECK("acc_reg_final");
  return 0;
}
