// Fused un-project -> scene-flow-advect -> re-project -> bilinear flow-warp -> consistency-loss chain.
//
// Replaces (reference paths relative to the reference tree):
//   unproject_ptcld.forward                 losses/scene_flow_projection.py:48-67      (W3)
//   flow_by_depth.forward                   losses/scene_flow_projection.py:95-153     (W1)
//   scene_flow_projection_slack.forward     losses/scene_flow_projection.py:204-278    (W2)
//   backward_warp + ATen grid_sampler_2d    losses/scene_flow_projection.py:103-112    (bilinear, border, align_corners=True)
//   Model._calc_loss / Model.disp_loss      models/scene_flow_motion_field.py:285-324,140-150 (L1)
//
// The reference runs ~100 ATen launches (16 broadcast 1x3·3x3 batched GEMMs, 3 grid_samples, 3
// nonzero+index_put pairs with host syncs) and materialises ~27 floats/pixel. Here the whole chain is
// one elementwise + gather kernel per direction: it is HBM-bound (no contraction => no tensor cores),
// so the design goals are coalesced 64/128-bit streaming loads, zero intermediate tensors, per-block
// partial sums through warp shuffles, and atomics only for the bilinear scatter of d(depth_2).
//
// Algorithmic HBM bytes per pixel (fp32): fwd 32 (d1 4, d2 4, flow 8, mask 4, sf 12);
// bwd 48 (the same 32 read + g_sf 12 + g_d2 4 written). See DESIGN.md for the full accounting.
#include "common.cuh"
#include <mutex>
#include <utility>
#include <vector>
#include "tc_common.cuh"
#include <initializer_list>
#include <stdlib.h>
#include <atomic>

namespace dvd {

struct __align__(16) Pose {
  float Kinv[9], K[9], R1[9], R2[9], t1[3], t2[3];
  // derived once per block (column-vector algebra, c = (x, y, 1)^T):
  float M1[9];   // R1 * Kinv                 P1  = d1 * (M1 c) + t1
  float A[9];    // R2^T * R1 * Kinv          p12 = d1 * (A c) + cv + R2^T sf
  float cv[3];   // R2^T (t1 - t2)
  float pad;
};
static_assert(sizeof(Pose) == 64 * sizeof(float), "pose layout");

// MUFU.RCP: max relative error 2^-23 (1 ulp), no slow path; inputs here are >= 1e-3 or flagged invalid
__device__ __forceinline__ float rcp_fast(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void mm3(const float* X, const float* Y, float* Z, bool xt) {
  // Z = (xt ? X^T : X) * Y
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float a = 0.f;
      for (int k = 0; k < 3; ++k) a = fmaf(xt ? X[k * 3 + i] : X[i * 3 + k], Y[k * 3 + j], a);
      Z[i * 3 + j] = a;
    }
}

__device__ __forceinline__ void load_pose(Pose& dst, const float* __restrict__ poses, int b) {
  // cooperative copy of one pair's pose block into shared memory + derived matrices
  float* d = reinterpret_cast<float*>(&dst);
  for (int i = threadIdx.x; i < 42; i += blockDim.x) d[i] = __ldg(poses + (size_t)b * DVD_POSE_STRIDE + i);
  __syncthreads();
  if (threadIdx.x == 0) {
    mm3(dst.R1, dst.Kinv, dst.M1, false);
    mm3(dst.R2, dst.M1, dst.A, true);
    float dx = dst.t1[0] - dst.t2[0], dy = dst.t1[1] - dst.t2[1], dz = dst.t1[2] - dst.t2[2];
    for (int i = 0; i < 3; ++i) dst.cv[i] = fmaf(dst.R2[6 + i], dz, fmaf(dst.R2[3 + i], dy, dst.R2[i] * dx));
  }
}

__device__ __forceinline__ void mv(const float* M, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = fmaf(M[2], z, fmaf(M[1], y, M[0] * x));
  oy = fmaf(M[5], z, fmaf(M[4], y, M[3] * x));
  oz = fmaf(M[8], z, fmaf(M[7], y, M[6] * x));
}
__device__ __forceinline__ void mtv(const float* M, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = fmaf(M[6], z, fmaf(M[3], y, M[0] * x));
  oy = fmaf(M[7], z, fmaf(M[4], y, M[1] * x));
  oz = fmaf(M[8], z, fmaf(M[5], y, M[2] * x));
}
// M * (x, y, 1)
__device__ __forceinline__ void ray_of(const float* M, float x, float y, float& rx, float& ry, float& rz) {
  rx = fmaf(M[1], y, M[0] * x) + M[2];
  ry = fmaf(M[4], y, M[3] * x) + M[5];
  rz = fmaf(M[7], y, M[6] * x) + M[8];
}

// Bilinear taps of ATen grid_sampler_2d(bilinear, padding_mode=border, align_corners=True) at the pixel
// coordinate (qx,qy). The reference normalises to [-1,1] (losses/...:107-110) and ATen un-normalises
// again; that round trip is the identity up to ~1e-7 relative (4e-5 px at W=384), far below the 1e-3
// parity bar, so it is skipped (4 IEEE divisions per pixel).
struct Taps {
  int idx[4];    // linear index y*W+x of nw, ne, sw, se (clamped in range)
  float w[4];    // weights, 0 for out-of-range taps
  float ux[2], uy[2];  // tap coordinates as floats: x0,x1 / y0,y1
};
__device__ __forceinline__ Taps make_taps(float qx, float qy, int H, int W) {
  const float hw = (float)(W - 1), hh = (float)(H - 1);
  float ix = fminf(hw, fmaxf(qx, 0.0f));
  float iy = fminf(hh, fmaxf(qy, 0.0f));
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx1 = ix - x0f, wy1 = iy - y0f;
  float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
  int x0 = (int)x0f, y0 = (int)y0f;
  bool vx1 = x0 + 1 <= W - 1, vy1 = y0 + 1 <= H - 1;
  int x1c = vx1 ? x0 + 1 : x0, y1c = vy1 ? y0 + 1 : y0;
  Taps t;
  t.idx[0] = y0 * W + x0;
  t.idx[1] = y0 * W + x1c;
  t.idx[2] = y1c * W + x0;
  t.idx[3] = y1c * W + x1c;
  t.w[0] = wx0 * wy0;
  t.w[1] = vx1 ? wx1 * wy0 : 0.0f;
  t.w[2] = vy1 ? wx0 * wy1 : 0.0f;
  t.w[3] = (vx1 && vy1) ? wx1 * wy1 : 0.0f;
  t.ux[0] = x0f; t.ux[1] = (float)x1c;
  t.uy[0] = y0f; t.uy[1] = (float)y1c;
  return t;
}

// Everything the forward chain produces for one pixel.
struct Px {
  float P1[3];      // global_p1
  float wpc[3];     // warped_p2_camera_2
  float wP2[3];     // warped_global_p2
  float wd;         // depth_warp_1_2
  float p12[3];     // p1_camera_2
  float i12[3];     // K * p12 (z = depth_image_1_2)
  float dflow[2];   // dflow_1_2
  float rz;         // 1 / (i12.z + 1e-8)
  bool zok;         // i12.z >= 1e-3 (projection used; otherwise own coordinate, zero gradient)
};

// kNeedWorld: also produce P1 and warped_global_p2 (only the sf_loss term / materialisation need them)
template <bool kNeedWorld>
__device__ __forceinline__ void forward_px(const Pose& ps, const float* __restrict__ d2img, const Taps& tp,
                                           float x, float y, float d1, float sfx, float sfy, float sfz, Px& o) {
  // bilinear gather: wpc = sum_k w_k d2_k Kinv (u_k, v_k, 1) = Kinv * (sum w d u, sum w d v, sum w d)
  float su = 0.f, sv = 0.f, s1 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float wd = tp.w[k] * __ldg(d2img + tp.idx[k]);
    su = fmaf(wd, tp.ux[k & 1], su);
    sv = fmaf(wd, tp.uy[k >> 1], sv);
    s1 += wd;
  }
  o.wd = s1;
  mv(ps.Kinv, su, sv, s1, o.wpc[0], o.wpc[1], o.wpc[2]);
  if (kNeedWorld) {
    float rx, ry, rz;
    ray_of(ps.M1, x, y, rx, ry, rz);
    o.P1[0] = fmaf(d1, rx, ps.t1[0]);
    o.P1[1] = fmaf(d1, ry, ps.t1[1]);
    o.P1[2] = fmaf(d1, rz, ps.t1[2]);
    // the four in-range bilinear weights sum to 1  =>  sum_k w_k (R2 p_k + t2) = R2 wpc + t2
    mv(ps.R2, o.wpc[0], o.wpc[1], o.wpc[2], o.wP2[0], o.wP2[1], o.wP2[2]);
    o.wP2[0] += ps.t2[0]; o.wP2[1] += ps.t2[1]; o.wP2[2] += ps.t2[2];
  }
  // p12 = R2^T (P1 + sf - t2) = d1 (A c) + cv + R2^T sf ;  i12 = K p12
  float ax, ay, az, bx, by, bz;
  ray_of(ps.A, x, y, ax, ay, az);
  mtv(ps.R2, sfx, sfy, sfz, bx, by, bz);
  o.p12[0] = fmaf(d1, ax, ps.cv[0]) + bx;
  o.p12[1] = fmaf(d1, ay, ps.cv[1]) + by;
  o.p12[2] = fmaf(d1, az, ps.cv[2]) + bz;
  mv(ps.K, o.p12[0], o.p12[1], o.p12[2], o.i12[0], o.i12[1], o.i12[2]);
  o.zok = !(o.i12[2] < 1e-3f);
  o.rz = rcp_fast(o.i12[2] + 1e-8f);
  if (o.zok) {
    o.dflow[0] = o.i12[0] * o.rz - x;
    o.dflow[1] = o.i12[1] * o.rz - y;
  } else {
    o.dflow[0] = 0.0f;
    o.dflow[1] = 0.0f;
  }
}

__device__ __forceinline__ float mask_of(const dvd_loss_cfg& c, float m2, float d1, float wz) {
  float m = m2;
  if (c.midas) {
    m *= (d1 < 100.0f) ? 1.0f : 0.0f;
    m *= (wz < 100.0f) ? 1.0f : 0.0f;
  }
  return m;
}

__device__ __forceinline__ float disp_term(const dvd_loss_cfg& c, float za, float zb) {
  if (c.disp_mode == 0) {
    float a = fmaxf(za, 1e-3f), b = fmaxf(zb, 1e-3f);
    return 100.0f * fabsf(rcp_fast(a) - rcp_fast(b));
  } else if (c.disp_mode == 1) {
    float a = fmaxf(za, 1e-3f), b = fmaxf(zb, 1e-3f);
    return fmaxf(a, b) / fminf(a, b) - 1.0f;
  }
  return fabsf(za - zb);
}

// ---------------------------------------------------------------------------------------------
// vector access helpers: VEC consecutive pixels along x per thread
template <int VEC> struct VecT;
template <> struct VecT<1> { using T = float;  using F = float2; };
template <> struct VecT<2> { using T = float2; using F = float4; };
template <> struct VecT<4> { using T = float4; using F = float4; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[VEC]) {
  if (VEC == 4) {
    float4 t = ldg_stream4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[VEC]) {
  if (VEC == 4) {
    st_stream4(p, make_float4(v[0], v[1], v[2], v[3]));
  } else if (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
    *p = v[0];
  }
}
template <int VEC>
__device__ __forceinline__ void load_flow(const float* __restrict__ p, float (&fx)[VEC], float (&fy)[VEC]) {
  // p points at flow[b,y,x,0]; 2*VEC consecutive floats
  if (VEC == 4) {
    float4 a = ldg_stream4(p), b = ldg_stream4(p + 4);
    fx[0] = a.x; fy[0] = a.y; fx[1] = a.z; fy[1] = a.w;
    fx[2] = b.x; fy[2] = b.y; fx[3] = b.z; fy[3] = b.w;
  } else if (VEC == 2) {
    float4 a = ldg_stream4(p);
    fx[0] = a.x; fy[0] = a.y; fx[1] = a.z; fy[1] = a.w;
  } else {
    float2 a = __ldg(reinterpret_cast<const float2*>(p));
    fx[0] = a.x; fy[0] = a.y;
  }
}

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// un-project forward / adjoint
template <int VEC>
__global__ void __launch_bounds__(kThreads) unproject_fwd_kernel(const float* __restrict__ depth,
                                                                 const float* __restrict__ poses,
                                                                 float* __restrict__ P, int H, int W, int which) {
  DVD_PDL_ENTER();
  __shared__ Pose ps;
  const int b = blockIdx.y;
  load_pose(ps, poses, b);
  __syncthreads();
  const float* R = which == 1 ? ps.R1 : ps.R2;
  const float* t = which == 1 ? ps.t1 : ps.t2;
  const int HW = H * W, items = HW / VEC, Wv = W / VEC;
  // (y, xv) walk of the grid-stride loop without a per-iteration integer division
  const int stride = gridDim.x * blockDim.x, sdy = stride / Wv, sdx = stride - sdy * Wv;
  int it = blockIdx.x * blockDim.x + threadIdx.x;
  int y = it / Wv, xv = it - y * Wv;
  for (; it < items; it += stride, y += sdy, xv += sdx) {
    if (xv >= Wv) { xv -= Wv; ++y; }
    const int x0 = xv * VEC;
    size_t pix = (size_t)y * W + x0;
    float d[VEC], ox[VEC], oy[VEC], oz[VEC];
    load_vec<VEC>(depth + (size_t)b * HW + pix, d);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float rx, ry, rz;
      ray_of(ps.Kinv, (float)(x0 + v), (float)y, rx, ry, rz);
      mv(R, d[v] * rx, d[v] * ry, d[v] * rz, ox[v], oy[v], oz[v]);
      ox[v] += t[0]; oy[v] += t[1]; oz[v] += t[2];
    }
    float* o = P + (size_t)b * 3 * HW + pix;
    store_vec<VEC>(o, ox);
    store_vec<VEC>(o + HW, oy);
    store_vec<VEC>(o + 2 * (size_t)HW, oz);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) unproject_bwd_kernel(const float* __restrict__ gP,
                                                                 const float* __restrict__ poses,
                                                                 float* __restrict__ gd, int H, int W, int which) {
  DVD_PDL_ENTER();
  __shared__ Pose ps;
  const int b = blockIdx.y;
  load_pose(ps, poses, b);
  __syncthreads();
  const float* R = which == 1 ? ps.R1 : ps.R2;
  const int HW = H * W, items = HW / VEC, Wv = W / VEC;
  // (y, xv) walk of the grid-stride loop without a per-iteration integer division
  const int stride = gridDim.x * blockDim.x, sdy = stride / Wv, sdx = stride - sdy * Wv;
  int it = blockIdx.x * blockDim.x + threadIdx.x;
  int y = it / Wv, xv = it - y * Wv;
  for (; it < items; it += stride, y += sdy, xv += sdx) {
    if (xv >= Wv) { xv -= Wv; ++y; }
    const int x0 = xv * VEC;
    size_t pix = (size_t)y * W + x0;
    float gx[VEC], gy[VEC], gz[VEC], o[VEC];
    const float* g = gP + (size_t)b * 3 * HW + pix;
    load_vec<VEC>(g, gx);
    load_vec<VEC>(g + HW, gy);
    load_vec<VEC>(g + 2 * (size_t)HW, gz);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float rx, ry, rz, wx, wy, wz;
      ray_of(ps.Kinv, (float)(x0 + v), (float)y, rx, ry, rz);
      mv(R, rx, ry, rz, wx, wy, wz);  // dP/dd = R * ray
      o[v] = fmaf(gz[v], wz, fmaf(gy[v], wy, gx[v] * wx));
    }
    store_vec<VEC>(gd + (size_t)b * HW + pix, o);
  }
}

// ---------------------------------------------------------------------------------------------
// fused forward: loss partial sums only
template <int VEC, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) reproject_loss_fwd_kernel(
    const float* __restrict__ depth_1, const float* __restrict__ depth_2, const float* __restrict__ flow,
    const float* __restrict__ mask_2, const float* __restrict__ sf, const float* __restrict__ poses,
    dvd_loss_cfg cfg, float* __restrict__ partials, int H, int W) {
  DVD_PDL_ENTER();
  __shared__ Pose ps;
  __shared__ float red[kThreads / 32][4];
  const int b = blockIdx.y;
  load_pose(ps, poses, b);
  __syncthreads();
  const int HW = H * W, items = HW / VEC, Wv = W / VEC;
  const float* d2img = depth_2 + (size_t)b * HW;
  float s_flow = 0.f, s_disp = 0.f, s_sf = 0.f, s_m = 0.f;
  // (y, xv) walk of the grid-stride loop without a per-iteration integer division
  const int stride = gridDim.x * blockDim.x, sdy = stride / Wv, sdx = stride - sdy * Wv;
  int it = blockIdx.x * blockDim.x + threadIdx.x;
  int y = it / Wv, xv = it - y * Wv;
  for (; it < items; it += stride, y += sdy, xv += sdx) {
    if (xv >= Wv) { xv -= Wv; ++y; }
    const int x0 = xv * VEC;
    size_t pix = (size_t)y * W + x0;
    float d1[VEC], m2[VEC], sx[VEC], sy[VEC], sz[VEC], fx[VEC], fy[VEC];
    load_vec<VEC>(depth_1 + (size_t)b * HW + pix, d1);
    load_vec<VEC>(mask_2 + (size_t)b * HW + pix, m2);
    const float* sfp = sf + (size_t)b * 3 * HW + pix;
    load_vec<VEC>(sfp, sx);
    load_vec<VEC>(sfp + HW, sy);
    load_vec<VEC>(sfp + 2 * (size_t)HW, sz);
    load_flow<VEC>(flow + ((size_t)b * HW + pix) * 2, fx, fy);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float x = (float)(x0 + v), yf = (float)y;
      Taps tp = make_taps(x + fx[v], yf + fy[v], H, W);
      Px o;
      forward_px<true>(ps, d2img, tp, x, yf, d1[v], sx[v], sy[v], sz[v], o);
      float m = mask_of(cfg, m2[v], d1[v], o.wpc[2]);
      float ex = o.dflow[0] - fx[v], ey = o.dflow[1] - fy[v];
      float fl = cfg.warm ? (ex * ex + ey * ey) : (fabsf(ex) + fabsf(ey));
      float dl = disp_term(cfg, o.p12[2], o.wpc[2]);
      float sl = fabsf(o.wP2[0] - o.P1[0] - sx[v]) + fabsf(o.wP2[1] - o.P1[1] - sy[v]) +
                 fabsf(o.wP2[2] - o.P1[2] - sz[v]);
      s_flow = fmaf(m, fl, s_flow);
      s_disp = fmaf(m, dl, s_disp);
      s_sf = fmaf(m, sl, s_sf);
      s_m += m;
    }
  }
  s_flow = warp_sum(s_flow); s_disp = warp_sum(s_disp); s_sf = warp_sum(s_sf); s_m = warp_sum(s_m);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[wid][0] = s_flow; red[wid][1] = s_disp; red[wid][2] = s_sf; red[wid][3] = s_m; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) a += red[w][threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = a;
  }
}

// deterministic final reduction (fixed order, double accumulation) + loss assembly
__global__ void __launch_bounds__(256) reproject_finalize_kernel(const float* __restrict__ partials, int n_quads,
                                                                 dvd_loss_cfg cfg, float* __restrict__ scalars) {
  DVD_PDL_ENTER();
  __shared__ double red[8][4];
  double acc[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < n_quads; i += blockDim.x) {
    float4 q = *reinterpret_cast<const float4*>(partials + (size_t)i * 4);
    acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) for (int k = 0; k < 4; ++k) red[wid][k] = acc[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    for (int w = 0; w < 8; ++w) for (int k = 0; k < 4; ++k) t[k] += red[w][k];
    float n = (float)t[3] + 1e-8f;  // torch.sum(occ_mask) + 1e-8 (smf.py:297-306)
    float fl = (float)t[0] / n, dl = (float)t[1] / n, sl = (float)t[2] / n;
    float second = cfg.second_is_disp ? dl : sl;
    scalars[DVD_S_FLOW] = fl;
    scalars[DVD_S_DISP] = dl;
    scalars[DVD_S_SF] = sl;
    scalars[DVD_S_LOSS] = cfg.flow_mul * fl + cfg.disp_mul * second;
    scalars[DVD_S_MASKSUM] = (float)t[3];
    scalars[DVD_S_CF] = cfg.flow_mul / n;
    scalars[DVD_S_CD] = cfg.disp_mul / n;
    scalars[DVD_S_RSVD] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// fused backward: g_sf (== g_global_p1) and scatter-add of g_depth_2
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <int VEC>
__global__ void __launch_bounds__(kThreads) reproject_loss_bwd_kernel(
    const float* __restrict__ depth_1, const float* __restrict__ depth_2, const float* __restrict__ flow,
    const float* __restrict__ mask_2, const float* __restrict__ sf, const float* __restrict__ poses,
    dvd_loss_cfg cfg, const float* __restrict__ scalars, float gscale, const float* __restrict__ gscale_dev,
    float* __restrict__ g_sf,
    float* __restrict__ g_d2, int H, int W) {
  DVD_PDL_ENTER();
  __shared__ Pose ps;
  const int b = blockIdx.y;
  load_pose(ps, poses, b);
  __syncthreads();
  const float gs = gscale * (gscale_dev ? __ldg(gscale_dev) : 1.0f);
  const float cf = __ldg(scalars + DVD_S_CF) * gs;
  const float cd = __ldg(scalars + DVD_S_CD) * gs;
  const int HW = H * W, items = HW / VEC, Wv = W / VEC;
  const float* d2img = depth_2 + (size_t)b * HW;
  float* gd2img = g_d2 ? g_d2 + (size_t)b * HW : nullptr;
  // (y, xv) walk of the grid-stride loop without a per-iteration integer division
  const int stride = gridDim.x * blockDim.x, sdy = stride / Wv, sdx = stride - sdy * Wv;
  int it = blockIdx.x * blockDim.x + threadIdx.x;
  int y = it / Wv, xv = it - y * Wv;
  for (; it < items; it += stride, y += sdy, xv += sdx) {
    if (xv >= Wv) { xv -= Wv; ++y; }
    const int x0 = xv * VEC;
    size_t pix = (size_t)y * W + x0;
    float d1[VEC], m2[VEC], sx[VEC], sy[VEC], sz[VEC], fx[VEC], fy[VEC];
    float ox[VEC], oy[VEC], oz[VEC];
    load_vec<VEC>(depth_1 + (size_t)b * HW + pix, d1);
    load_vec<VEC>(mask_2 + (size_t)b * HW + pix, m2);
    const float* sfp = sf + (size_t)b * 3 * HW + pix;
    load_vec<VEC>(sfp, sx);
    load_vec<VEC>(sfp + HW, sy);
    load_vec<VEC>(sfp + 2 * (size_t)HW, sz);
    load_flow<VEC>(flow + ((size_t)b * HW + pix) * 2, fx, fy);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float x = (float)(x0 + v), yf = (float)y;
      Taps tp = make_taps(x + fx[v], yf + fy[v], H, W);
      Px o;
      if (cfg.second_is_disp) forward_px<false>(ps, d2img, tp, x, yf, d1[v], sx[v], sy[v], sz[v], o);
      else                    forward_px<true>(ps, d2img, tp, x, yf, d1[v], sx[v], sy[v], sz[v], o);
      float m = mask_of(cfg, m2[v], d1[v], o.wpc[2]);
      // --- gradient w.r.t. i12 = K p12 from the flow term
      float gi0 = 0.f, gi1 = 0.f, gi2 = 0.f;
      if (o.zok) {
        float ex = o.dflow[0] - fx[v], ey = o.dflow[1] - fy[v];
        float gux = cfg.warm ? 2.f * ex : sgn(ex);
        float guy = cfg.warm ? 2.f * ey : sgn(ey);
        gux *= m * cf; guy *= m * cf;
        const float rz = o.rz;
        gi0 = gux * rz;
        gi1 = guy * rz;
        gi2 = -(gux * o.i12[0] + guy * o.i12[1]) * rz * rz;
      }
      float gp0, gp1, gp2;  // g_p12 = K^T g_i12
      mtv(ps.K, gi0, gi1, gi2, gp0, gp1, gp2);
      float gwc0 = 0.f, gwc1 = 0.f, gwc2 = 0.f;  // g_warped_p2_camera_2
      float ge0 = 0.f, ge1 = 0.f, ge2 = 0.f;     // g_(sf_by_depth - sf)
      if (cfg.second_is_disp) {
        float za = o.p12[2], zb = o.wpc[2];
        float mc = m * cd;
        if (cfg.disp_mode == 0) {
          float a = fmaxf(za, 1e-3f), bb = fmaxf(zb, 1e-3f);
          const float ra = rcp_fast(a), rb = rcp_fast(bb);
          float s = 100.f * sgn(ra - rb) * mc;
          if (za >= 1e-3f) gp2 -= s * ra * ra;
          if (zb >= 1e-3f) gwc2 += s * rb * rb;
        } else if (cfg.disp_mode == 1) {
          float a = fmaxf(za, 1e-3f), bb = fmaxf(zb, 1e-3f);
          // max(a,b)/min(a,b) - 1
          float ga, gb;
          const float ra = rcp_fast(a), rb = rcp_fast(bb);
          if (a >= bb) { ga = rb; gb = -a * rb * rb; }
          else         { ga = -bb * ra * ra; gb = ra; }
          if (za >= 1e-3f) gp2 += ga * mc;
          if (zb >= 1e-3f) gwc2 += gb * mc;
        } else {
          float s = sgn(za - zb) * mc;
          gp2 += s;
          gwc2 -= s;
        }
      } else {
        float mc = m * cd;
        ge0 = mc * sgn(o.wP2[0] - o.P1[0] - sx[v]);
        ge1 = mc * sgn(o.wP2[1] - o.P1[1] - sy[v]);
        ge2 = mc * sgn(o.wP2[2] - o.P1[2] - sz[v]);
        // wP2 = sum_k w_k (R2 p2c2_k + t2)  =>  g_wpc += R2^T g_e
        float a0, a1, a2;
        mtv(ps.R2, ge0, ge1, ge2, a0, a1, a2);
        gwc0 += a0; gwc1 += a1; gwc2 += a2;
      }
      // g_(P1 + sf) = R2 g_p12 ; sf_loss adds -g_e to both P1 and sf
      float gv0, gv1, gv2;
      mv(ps.R2, gp0, gp1, gp2, gv0, gv1, gv2);
      ox[v] = gv0 - ge0; oy[v] = gv1 - ge1; oz[v] = gv2 - ge2;
      // scatter to depth_2: wpc = Kinv (sum w d u, sum w d v, sum w d)  =>  g_d2_k = w_k (Kinv^T g_wpc).(u_k, v_k, 1)
      if (gd2img) {
        float hu, hv, h1;
        mtv(ps.Kinv, gwc0, gwc1, gwc2, hu, hv, h1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float g = tp.w[k] * fmaf(hu, tp.ux[k & 1], fmaf(hv, tp.uy[k >> 1], h1));
          if (g != 0.f) atomicAdd(gd2img + tp.idx[k], g);
        }
      }
    }
    float* gs = g_sf + (size_t)b * 3 * HW + pix;
    store_vec<VEC>(gs, ox);
    store_vec<VEC>(gs + HW, oy);
    store_vec<VEC>(gs + 2 * (size_t)HW, oz);
  }
}

// ---------------------------------------------------------------------------------------------
// materialise every per-pixel tensor of the two reference modules (not on the training fast path)
__global__ void __launch_bounds__(kThreads) reproject_materialize_kernel(
    const float* __restrict__ depth_1, const float* __restrict__ depth_2, const float* __restrict__ flow,
    const float* __restrict__ sf, const float* __restrict__ poses, float* __restrict__ global_p1,
    float* __restrict__ sf_by_depth, float* __restrict__ warped_global_p2, float* __restrict__ warped_p2_camera_2,
    float* __restrict__ p1_camera_2, float* __restrict__ dflow, float* __restrict__ staticflow,
    float* __restrict__ depth_image, float* __restrict__ depth_warp, int H, int W) {
  DVD_PDL_ENTER();
  __shared__ Pose ps;
  const int b = blockIdx.y;
  load_pose(ps, poses, b);
  __syncthreads();
  const int HW = H * W;
  const float* d2img = depth_2 + (size_t)b * HW;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    int y = pix / W, xi = pix - y * W;
    float x = (float)xi, yf = (float)y;
    float d1 = depth_1[(size_t)b * HW + pix];
    float2 f = *reinterpret_cast<const float2*>(flow + ((size_t)b * HW + pix) * 2);
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (sf) {
      const float* sfp = sf + (size_t)b * 3 * HW + pix;
      sx = sfp[0]; sy = sfp[HW]; sz = sfp[2 * (size_t)HW];
    }
    Taps tp = make_taps(x + f.x, yf + f.y, H, W);
    Px o, os;
    forward_px<true>(ps, d2img, tp, x, yf, d1, sx, sy, sz, o);
    size_t o3 = (size_t)b * 3 * HW + pix, o2 = (size_t)b * 2 * HW + pix, o1 = (size_t)b * HW + pix;
    if (global_p1) { global_p1[o3] = o.P1[0]; global_p1[o3 + HW] = o.P1[1]; global_p1[o3 + 2 * (size_t)HW] = o.P1[2]; }
    if (sf_by_depth) {
      sf_by_depth[o3] = o.wP2[0] - o.P1[0];
      sf_by_depth[o3 + HW] = o.wP2[1] - o.P1[1];
      sf_by_depth[o3 + 2 * (size_t)HW] = o.wP2[2] - o.P1[2];
    }
    if (warped_global_p2) { warped_global_p2[o3] = o.wP2[0]; warped_global_p2[o3 + HW] = o.wP2[1]; warped_global_p2[o3 + 2 * (size_t)HW] = o.wP2[2]; }
    if (warped_p2_camera_2) { warped_p2_camera_2[o3] = o.wpc[0]; warped_p2_camera_2[o3 + HW] = o.wpc[1]; warped_p2_camera_2[o3 + 2 * (size_t)HW] = o.wpc[2]; }
    if (p1_camera_2) { p1_camera_2[o3] = o.p12[0]; p1_camera_2[o3 + HW] = o.p12[1]; p1_camera_2[o3 + 2 * (size_t)HW] = o.p12[2]; }
    if (dflow) { dflow[o2] = o.dflow[0]; dflow[o2 + HW] = o.dflow[1]; }
    if (depth_image) depth_image[o1] = o.i12[2];
    if (depth_warp) depth_warp[o1] = o.wd;
    if (staticflow) {
      forward_px<false>(ps, d2img, tp, x, yf, d1, 0.f, 0.f, 0.f, os);
      staticflow[o2] = os.dflow[0];
      staticflow[o2 + HW] = os.dflow[1];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// adjoint of reproject_materialize_kernel for ARBITRARY cotangents on its nine outputs (operator-level drop-in:
// the mirrors of flow_by_depth / scene_flow_projection_slack stay differentiable for user-defined losses).
struct MatGrads {
  const float* global_p1; const float* sf_by_depth; const float* warped_global_p2; const float* warped_p2_camera_2;
  const float* p1_camera_2; const float* dflow; const float* staticflow; const float* depth_image; const float* depth_warp;
};
__device__ __forceinline__ void ld3(const float* p, size_t o3, int HW, float& a, float& b, float& c) {
  if (p) { a = p[o3]; b = p[o3 + HW]; c = p[o3 + 2 * (size_t)HW]; } else { a = b = c = 0.f; }
}
__global__ void __launch_bounds__(kThreads) reproject_materialize_bwd_kernel(
    const float* __restrict__ depth_1, const float* __restrict__ depth_2, const float* __restrict__ flow,
    const float* __restrict__ sf, const float* __restrict__ poses, MatGrads G, float* __restrict__ g_d1,
    float* __restrict__ g_d2, float* __restrict__ g_sf, int H, int W) {
  DVD_PDL_ENTER();
  __shared__ Pose ps;
  const int b = blockIdx.y;
  load_pose(ps, poses, b);
  __syncthreads();
  const int HW = H * W;
  const float* d2img = depth_2 + (size_t)b * HW;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    const int y = pix / W, xi = pix - y * W;
    const float x = (float)xi, yf = (float)y;
    const float d1 = depth_1[(size_t)b * HW + pix];
    const float2 f = *reinterpret_cast<const float2*>(flow + ((size_t)b * HW + pix) * 2);
    float sx = 0.f, sy = 0.f, sz = 0.f;
    const size_t o3 = (size_t)b * 3 * HW + pix, o2 = (size_t)b * 2 * HW + pix, o1 = (size_t)b * HW + pix;
    if (sf) { sx = sf[o3]; sy = sf[o3 + HW]; sz = sf[o3 + 2 * (size_t)HW]; }
    Taps tp = make_taps(x + f.x, yf + f.y, H, W);
    Px o, os;
    forward_px<false>(ps, d2img, tp, x, yf, d1, sx, sy, sz, o);
    float a0, a1, a2, b0, b1, b2;
    // P1: + global_p1, - sf_by_depth ;  wP2: + sf_by_depth, + warped_global_p2
    float gP0, gP1, gP2, gW0, gW1, gW2;
    ld3(G.global_p1, o3, HW, gP0, gP1, gP2);
    ld3(G.sf_by_depth, o3, HW, a0, a1, a2);
    gP0 -= a0; gP1 -= a1; gP2 -= a2;
    ld3(G.warped_global_p2, o3, HW, gW0, gW1, gW2);
    gW0 += a0; gW1 += a1; gW2 += a2;
    // wpc: + warped_p2_camera_2 + R2^T g_wP2
    float gC0, gC1, gC2;
    ld3(G.warped_p2_camera_2, o3, HW, gC0, gC1, gC2);
    mtv(ps.R2, gW0, gW1, gW2, a0, a1, a2);
    gC0 += a0; gC1 += a1; gC2 += a2;
    // p12: + p1_camera_2 + K^T g_i12
    float gp0, gp1, gp2;
    ld3(G.p1_camera_2, o3, HW, gp0, gp1, gp2);
    float gi0 = 0.f, gi1 = 0.f, gi2 = G.depth_image ? G.depth_image[o1] : 0.f;
    if (G.dflow && o.zok) {
      const float gu = G.dflow[o2], gv = G.dflow[o2 + HW];
      gi0 += gu * o.rz; gi1 += gv * o.rz;
      gi2 -= (gu * o.i12[0] + gv * o.i12[1]) * o.rz * o.rz;
    }
    mtv(ps.K, gi0, gi1, gi2, a0, a1, a2);
    gp0 += a0; gp1 += a1; gp2 += a2;
    mv(ps.R2, gp0, gp1, gp2, b0, b1, b2);        // g_(P1 + sf)
    if (g_sf) { g_sf[o3] = b0; g_sf[o3 + HW] = b1; g_sf[o3 + 2 * (size_t)HW] = b2; }
    gP0 += b0; gP1 += b1; gP2 += b2;
    if (G.staticflow) {
      forward_px<false>(ps, d2img, tp, x, yf, d1, 0.f, 0.f, 0.f, os);
      if (os.zok) {
        const float gu = G.staticflow[o2], gv = G.staticflow[o2 + HW];
        const float s0 = gu * os.rz, s1 = gv * os.rz, s2 = -(gu * os.i12[0] + gv * os.i12[1]) * os.rz * os.rz;
        mtv(ps.K, s0, s1, s2, a0, a1, a2);
        mv(ps.R2, a0, a1, a2, b0, b1, b2);
        gP0 += b0; gP1 += b1; gP2 += b2;
      }
    }
    // P1 = d1 * (M1 c) + t1
    float rx, ry, rz;
    ray_of(ps.M1, x, yf, rx, ry, rz);
    if (g_d1) g_d1[o1] = fmaf(gP2, rz, fmaf(gP1, ry, gP0 * rx));
    if (g_d2) {
      float hu, hv, h1;
      mtv(ps.Kinv, gC0, gC1, gC2, hu, hv, h1);
      if (G.depth_warp) h1 += G.depth_warp[o1];
      float* gd2img = g_d2 + (size_t)b * HW;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float g = tp.w[k] * fmaf(hu, tp.ux[k & 1], fmaf(hv, tp.uy[k >> 1], h1));
        if (g != 0.f) atomicAdd(gd2img + tp.idx[k], g);
      }
    }
  }
}

// =============================================================================================
// Packed-FP32 path of the fused loss kernels (sm_100: FFMA2 / FMUL2 / FADD2 process two fp32 lanes per issue slot).
//
// The scalar kernels above are bound by the FP32 issue rate, not by HBM (profiles/r1_ncu_reproject_*_v3.json:
// ~218 / ~340 warp instructions per pixel forward / backward, issue slots 63 % busy at 27 % of DRAM peak). Here each
// thread owns PAIRS of x-adjacent pixels and every matrix-vector product of the chain runs on float2 operands, the
// pose entries entering as broadcast scalar operands straight from the constant bank (no shared-memory loads, no
// register copies). Pose-derived matrices are prepared once per call by pose_prep_kernel and copied into a
// __constant__ slot with a device-to-device cudaMemcpyToSymbolAsync on the caller's stream (no host round trip).
struct __align__(16) PoseC {
  float Kinv[9];   //  0
  float K[9];      //  9
  float R2[9];     // 18
  float nM1[9];    // 27  -(R1 Kinv)                 -(P1 - t1) = d1 * (nM1 c)
  float A[9];      // 36  R2^T R1 Kinv               p12 = d1 * (A c) + cv + R2^T sf
  float cv[3];     // 45  R2^T (t1 - t2)
  float t21[3];    // 48  t2 - t1                    warped_global_p2 - P1 = R2 wpc + t21 + d1 * (nM1 c)
  float pad[13];
};
static_assert(sizeof(PoseC) == 256, "PoseC layout");
constexpr int kPoseSlots = 3;    // rotating slots: calls in flight on different streams do not share a slot
constexpr int kPosePairs = 64;   // pairs per launch (larger batches are processed in chunks)
__constant__ PoseC c_pose[kPoseSlots][kPosePairs];
__device__ PoseC g_pose_stage[kPoseSlots][kPosePairs];

__global__ void pose_prep_kernel(const float* __restrict__ poses, int B, int slot) {
  DVD_PDL_ENTER();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p = poses + (size_t)b * DVD_POSE_STRIDE;
  float Kinv[9], R1[9], R2[9], M1[9], t1[3], t2[3];
  PoseC& o = g_pose_stage[slot][b];
  for (int i = 0; i < 9; ++i) {
    Kinv[i] = p[i]; R1[i] = p[18 + i]; R2[i] = p[27 + i];
    o.Kinv[i] = Kinv[i]; o.K[i] = p[9 + i]; o.R2[i] = R2[i];
  }
  for (int i = 0; i < 3; ++i) { t1[i] = p[36 + i]; t2[i] = p[39 + i]; }
  mm3(R1, Kinv, M1, false);
  mm3(R2, M1, o.A, true);
  for (int i = 0; i < 9; ++i) o.nM1[i] = -M1[i];
  const float dx = t1[0] - t2[0], dy = t1[1] - t2[1], dz = t1[2] - t2[2];
  for (int i = 0; i < 3; ++i) {
    o.cv[i] = fmaf(R2[6 + i], dz, fmaf(R2[3 + i], dy, R2[i] * dx));
    o.t21[i] = t2[i] - t1[i];
  }
}

__device__ __forceinline__ float2 F2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
// a - b as one FFMA2 (b * -1 + a)
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, F2(-1.0f), a); }
// M * v  /  M^T * v  with broadcast matrix entries
__device__ __forceinline__ void mv2(const float* M, float2 x, float2 y, float2 z, float2& ox, float2& oy, float2& oz) {
  ox = fma2(F2(M[2]), z, fma2(F2(M[1]), y, mul2(F2(M[0]), x)));
  oy = fma2(F2(M[5]), z, fma2(F2(M[4]), y, mul2(F2(M[3]), x)));
  oz = fma2(F2(M[8]), z, fma2(F2(M[7]), y, mul2(F2(M[6]), x)));
}
__device__ __forceinline__ void mtv2(const float* M, float2 x, float2 y, float2 z, float2& ox, float2& oy, float2& oz) {
  ox = fma2(F2(M[6]), z, fma2(F2(M[3]), y, mul2(F2(M[0]), x)));
  oy = fma2(F2(M[7]), z, fma2(F2(M[4]), y, mul2(F2(M[1]), x)));
  oz = fma2(F2(M[8]), z, fma2(F2(M[5]), y, mul2(F2(M[2]), x)));
}

// forward chain of one pixel pair
struct PairOut {
  int i00[2], sx1[2], sy1[2];   // nw tap index; element offsets to the ne / sw taps (0 where clamped by the border)
  float2 w[4];            // bilinear weights; a tap clamped by the border has weight exactly 0
  float2 x0f, y0f;        // tap origin (the ne / sw / se taps sit at +1 wherever their weight is non-zero)
  float2 wpc[3], p12[3], i12[3], rz;
  float2 ex, ey;          // dflow_1_2 - flow_1_2 (zero flow substituted where the projection is rejected)
  float2 e[3];            // warped_global_p2 - P1 - sf            (kWorld only)
  bool zok[2];
};

// Where the four bilinear taps of a pixel come from / where their gradients go.
struct GlobalTaps {
  const float* img;   // depth_2 of this pair
  __device__ __forceinline__ void ld4(int i00, int sx1, int sy1, float (&d)[4]) const {
    const float* p = img + i00;   // one 64-bit address per pixel, the other taps at small element offsets from it
    d[0] = __ldg(p); d[1] = __ldg(p + sx1); d[2] = __ldg(p + sy1); d[3] = __ldg(p + sy1 + sx1);
  }
};
__device__ __forceinline__ void red_global_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_global_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// scatter-add of the four tap gradients of one pixel; the packed kernels require W % 4 == 0, so whenever the nw tap sits
// on an even column the (nw, ne) and (sw, se) taps are two 8-byte aligned pairs: one vector reduction each
__device__ __forceinline__ void scatter4_global(float* q, int i00, int sx1, int sy1, const float (&g)[4]) {
  if (((i00 & 1) == 0) && sx1 == 1) {
    if (g[0] != 0.f || g[1] != 0.f) red_global_v2(q, g[0], g[1]);
    if (g[2] != 0.f || g[3] != 0.f) {
      if (sy1 != 0) red_global_v2(q + sy1, g[2], g[3]);
      else red_global_v2(q, g[2], g[3]);     // clamped bottom row: both weights are exactly 0 here, kept for form
    }
  } else if (((i00 & 3) == 1) && sx1 == 1) {
    // odd column whose pair still lies inside one 16-byte quad: one 4-wide reduction {0, g, g, 0} per row instead of two
    // scalar ones (the LSU's reduction rate is per active lane, not per byte: DESIGN.md 8)
    if (g[0] != 0.f || g[1] != 0.f) red_global_v4(q - 1, 0.f, g[0], g[1], 0.f);
    if (g[2] != 0.f || g[3] != 0.f) red_global_v4(q - 1 + sy1, 0.f, g[2], g[3], 0.f);
  } else {
    if (g[0] != 0.f) atomicAdd(q, g[0]);
    if (g[1] != 0.f) atomicAdd(q + sx1, g[1]);
    if (g[2] != 0.f) atomicAdd(q + sy1, g[2]);
    if (g[3] != 0.f) atomicAdd(q + sy1 + sx1, g[3]);
  }
}
struct GlobalScatter {
  float* gimg;        // g_depth_2 of this pair (nullptr: no scatter)
  __device__ __forceinline__ bool on() const { return gimg != nullptr; }
  __device__ __forceinline__ void add4(int i00, int sx1, int sy1, const float (&g)[4]) const {
    scatter4_global(gimg + i00, i00, sx1, sy1, g);
  }
};
template <bool kWorld, class Taps>
__device__ __forceinline__ void pair_forward(const PoseC& ps, const Taps& taps, int H, int W, float2 nx,
                                             float nyf, float2 fx, float2 fy, float2 d1, float2 sx, float2 sy, float2 sz,
                                             const float (&rn)[3], const float (&ra)[3], PairOut& o) {
  const float hw = (float)(W - 1), hh = (float)(H - 1);
  const float2 nqx = fma2(fx, F2(-1.0f), nx);        // -(x + flow_x)
  const float2 nqy = fma2(fy, F2(-1.0f), F2(nyf));   // -(y + flow_y)
  float ixs[2], iys[2], xfs[2], yfs[2];
  const float nq[2][2] = {{nqx.x, nqy.x}, {nqx.y, nqy.y}};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    ixs[j] = fminf(hw, fmaxf(-nq[j][0], 0.0f));
    iys[j] = fminf(hh, fmaxf(-nq[j][1], 0.0f));
    xfs[j] = floorf(ixs[j]);
    yfs[j] = floorf(iys[j]);
    const int xi = (int)xfs[j], yi = (int)yfs[j];
    o.i00[j] = yi * W + xi;
    o.sx1[j] = xi < W - 1 ? 1 : 0;
    o.sy1[j] = yi < H - 1 ? W : 0;
  }
  float2 dk[4];
  {
    float da[4], db[4];
    taps.ld4(o.i00[0], o.sx1[0], o.sy1[0], da);
    taps.ld4(o.i00[1], o.sx1[1], o.sy1[1], db);
#pragma unroll
    for (int k = 0; k < 4; ++k) dk[k] = make_float2(da[k], db[k]);
  }
  // ---- arithmetic that does not depend on the gathered taps first: it runs while the gather is in flight ----
  // p12 = d1 (A c) + cv + R2^T sf ; c = (x, y, 1): A c = A[:,0] x + (A[:,1] y + A[:,2]) = -A[:,0] nx + ra
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float2 ac = fma2(F2(-ps.A[3 * k]), nx, F2(ra[k]));
    float2 t = fma2(d1, ac, F2(ps.cv[k]));
    t = fma2(F2(ps.R2[k]), sx, t);
    t = fma2(F2(ps.R2[3 + k]), sy, t);
    o.p12[k] = fma2(F2(ps.R2[6 + k]), sz, t);
  }
  mv2(ps.K, o.p12[0], o.p12[1], o.p12[2], o.i12[0], o.i12[1], o.i12[2]);
  const float2 zz = add2(o.i12[2], F2(1e-8f));
  o.rz = make_float2(rcp_fast(zz.x), rcp_fast(zz.y));
  o.zok[0] = !(o.i12[2].x < 1e-3f);
  o.zok[1] = !(o.i12[2].y < 1e-3f);
  // dflow - flow = i12.xy * rz - (c.xy + flow)
  const float2 ex = fma2(o.i12[0], o.rz, nqx), ey = fma2(o.i12[1], o.rz, nqy);
  o.ex = make_float2(o.zok[0] ? ex.x : -fx.x, o.zok[1] ? ex.y : -fx.y);
  o.ey = make_float2(o.zok[0] ? ey.x : -fy.x, o.zok[1] ? ey.y : -fy.y);
  o.x0f = make_float2(xfs[0], xfs[1]);
  o.y0f = make_float2(yfs[0], yfs[1]);
  // clamped coordinate == W-1 (H-1) implies a zero fractional part, so the clamped taps need no explicit masking
  const float2 wx1 = sub2(make_float2(ixs[0], ixs[1]), o.x0f), wy1 = sub2(make_float2(iys[0], iys[1]), o.y0f);
  const float2 wx0 = fma2(wx1, F2(-1.0f), F2(1.0f)), wy0 = fma2(wy1, F2(-1.0f), F2(1.0f));
  o.w[0] = mul2(wx0, wy0); o.w[1] = mul2(wx1, wy0); o.w[2] = mul2(wx0, wy1); o.w[3] = mul2(wx1, wy1);
  float2 et[3];
  if (kWorld) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float2 nr = fma2(F2(-ps.nM1[3 * k]), nx, F2(rn[k]));   // -(M1 c)_k
      et[k] = sub2(fma2(d1, nr, F2(ps.t21[k])), k == 0 ? sx : (k == 1 ? sy : sz));
    }
  }
  // ---- gathered taps ----
  // wpc = sum_k w_k d2_k Kinv (u_k, v_k, 1) = Kinv (su, sv, s1), u_k = x0f (+1), v_k = y0f (+1)
  const float2 wd0 = mul2(o.w[0], dk[0]), wd1 = mul2(o.w[1], dk[1]), wd2 = mul2(o.w[2], dk[2]), wd3 = mul2(o.w[3], dk[3]);
  const float2 eb = add2(wd1, wd3), sb = add2(wd2, wd3);
  const float2 s1 = add2(add2(wd0, wd1), sb);
  const float2 su = fma2(o.x0f, s1, eb), sv = fma2(o.y0f, s1, sb);
  mv2(ps.Kinv, su, sv, s1, o.wpc[0], o.wpc[1], o.wpc[2]);
  if (kWorld) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float2 t = fma2(F2(ps.R2[3 * k]), o.wpc[0], et[k]);
      t = fma2(F2(ps.R2[3 * k + 1]), o.wpc[1], t);
      o.e[k] = fma2(F2(ps.R2[3 * k + 2]), o.wpc[2], t);
    }
  }
}

__device__ __forceinline__ float mask1(const dvd_loss_cfg& c, float m2, float d1, float wz) {
  return (!c.midas || (d1 < 100.0f && wz < 100.0f)) ? m2 : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Staged variant: the five streamed inputs (28 of the 32 bytes per pixel) travel global -> shared memory as 1-D bulk
// async copies issued by a producer warp into a ring of STAGES tiles, completion on mbarriers; the 256 consumer threads
// only see shared-memory latency for them, and the bytes in flight per SM (CTAs x STAGES x 14 KB) no longer depend on
// occupancy or on how the compiler schedules the loads. Only the bilinear gather of depth_2 is a global load.
// Measured dead ends (profiles/r1_ncu_reproject_variants.txt): also staging the depth_2 rows around the tile in shared
// memory (taps as LDS) is not faster, and accumulating g_depth_2 in a shared-memory window is not either - fp32
// red.shared compiles to an ATOMS.CAST.SPIN compare-and-swap loop on sm_100.
template <int NP> struct StagedCfg {
  static constexpr int VEC = 2 * NP;                 // pixels per consumer thread and tile
  static constexpr int TILE = kThreads * VEC;        // pixels per stage
  static constexpr int FLOATS = TILE * 7;            // d1, mask, sf.x, sf.y, sf.z, flow (2 floats per pixel)
};

// stream one tile (pixels [p0, p0 + n) of pair b) into a stage
template <int TILE>
__device__ __forceinline__ void produce_tile(float* dst, uint64_t* bar, const float* depth_1, const float* mask_2,
                                             const float* sf, const float* flow, size_t b, int HW, int p0, int pend) {
  using namespace tc;
  const uint32_t n4 = (uint32_t)min(TILE, pend - p0) * 4u;
  mbar_arrive_expect_tx(bar, n4 * 7u);
  bulk_g2s(dst, depth_1 + b * HW + p0, n4, bar);
  bulk_g2s(dst + TILE, mask_2 + b * HW + p0, n4, bar);
  bulk_g2s(dst + 2 * TILE, sf + (b * 3 + 0) * HW + p0, n4, bar);
  bulk_g2s(dst + 3 * TILE, sf + (b * 3 + 1) * HW + p0, n4, bar);
  bulk_g2s(dst + 4 * TILE, sf + (b * 3 + 2) * HW + p0, n4, bar);
  bulk_g2s(dst + 5 * TILE, flow + (b * HW + p0) * 2, n4 * 2u, bar);
}
// a consumer thread's VEC pixels of a stage -> registers
template <int NP>
__device__ __forceinline__ void consume_tile(const float* src, int tv, float (&d1)[2 * NP], float (&m2)[2 * NP],
                                             float (&sx)[2 * NP], float (&sy)[2 * NP], float (&sz)[2 * NP],
                                             float (&fx)[2 * NP], float (&fy)[2 * NP]) {
  constexpr int VEC = 2 * NP, TILE = StagedCfg<NP>::TILE;
  src += tv;
  if (VEC == 2) {
    const float2 a = *reinterpret_cast<const float2*>(src), bq = *reinterpret_cast<const float2*>(src + TILE);
    const float2 c = *reinterpret_cast<const float2*>(src + 2 * TILE), d = *reinterpret_cast<const float2*>(src + 3 * TILE);
    const float2 e = *reinterpret_cast<const float2*>(src + 4 * TILE);
    const float4 f = *reinterpret_cast<const float4*>(src + 5 * TILE + tv);
    d1[0] = a.x; d1[1] = a.y; m2[0] = bq.x; m2[1] = bq.y; sx[0] = c.x; sx[1] = c.y; sy[0] = d.x; sy[1] = d.y;
    sz[0] = e.x; sz[1] = e.y; fx[0] = f.x; fy[0] = f.y; fx[1] = f.z; fy[1] = f.w;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(src), bq = *reinterpret_cast<const float4*>(src + TILE);
    const float4 c = *reinterpret_cast<const float4*>(src + 2 * TILE), d = *reinterpret_cast<const float4*>(src + 3 * TILE);
    const float4 e = *reinterpret_cast<const float4*>(src + 4 * TILE);
    const float4 f = *reinterpret_cast<const float4*>(src + 5 * TILE + tv), g = *reinterpret_cast<const float4*>(src + 5 * TILE + tv + 4);
    d1[0] = a.x; d1[1] = a.y; d1[2 % VEC] = a.z; d1[3 % VEC] = a.w;
    m2[0] = bq.x; m2[1] = bq.y; m2[2 % VEC] = bq.z; m2[3 % VEC] = bq.w;
    sx[0] = c.x; sx[1] = c.y; sx[2 % VEC] = c.z; sx[3 % VEC] = c.w;
    sy[0] = d.x; sy[1] = d.y; sy[2 % VEC] = d.z; sy[3 % VEC] = d.w;
    sz[0] = e.x; sz[1] = e.y; sz[2 % VEC] = e.z; sz[3 % VEC] = e.w;
    fx[0] = f.x; fy[0] = f.y; fx[1] = f.z; fy[1] = f.w; fx[2 % VEC] = g.x; fy[2 % VEC] = g.y; fx[3 % VEC] = g.z; fy[3 % VEC] = g.w;
  }
}

// NP pixel pairs per consumer thread, STAGES-deep ring, MINB co-resident CTAs per SM
template <int NP, int STAGES, int MINB, bool kDry>
__global__ void __launch_bounds__(kThreads + 32, MINB) reproject_loss_fwd_staged_kernel(
    const float* __restrict__ depth_1, const float* __restrict__ depth_2, const float* __restrict__ flow,
    const float* __restrict__ mask_2, const float* __restrict__ sf, dvd_loss_cfg cfg, float* __restrict__ partials,
    int H, int W, int slot, int b0, int nb, int tiles_per_pair) {
  DVD_PDL_ENTER();
  using namespace tc;
  constexpr int VEC = 2 * NP, TILE = StagedCfg<NP>::TILE, FLOATS = StagedCfg<NP>::FLOATS;
  extern __shared__ __align__(128) float stage_mem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  __shared__ float red[kThreads / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = H * W, ntiles = nb * tiles_per_pair;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kThreads / 32); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == kThreads / 32) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int bl = tile / tiles_per_pair, t = tile - bl * tiles_per_pair;
        produce_tile<TILE>(stage_mem + (size_t)s * FLOATS, &full[s], depth_1, mask_2, sf, flow, (size_t)(b0 + bl), HW, t * TILE, HW);
      }
    }
    return;
  }
  // ===== consumers =====
  float2 a_flow = F2(0.f), a_disp = F2(0.f), a_sf = F2(0.f), a_m = F2(0.f);
  const int tv = threadIdx.x * VEC;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
    mbar_wait(&full[s], ph);
    float d1[VEC], m2[VEC], sx[VEC], sy[VEC], sz[VEC], fx[VEC], fy[VEC];
    consume_tile<NP>(stage_mem + (size_t)s * FLOATS, tv, d1, m2, sx, sy, sz, fx, fy);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);     // values are in registers: hand the stage back
    const int bl = tile / tiles_per_pair, t = tile - bl * tiles_per_pair;
    const int p = t * TILE + tv;
    if (p >= HW) continue;
    if (kDry) {   // memory-system probe (DVD_REPROJECT_DRY=1): stream the inputs, one gather tap, no chain arithmetic
      const float g = __ldg(depth_2 + (size_t)(b0 + bl) * HW + p);
      a_flow = add2(a_flow, make_float2(d1[0] + d1[1] + g, m2[0] + m2[1]));
      a_sf = add2(a_sf, make_float2(sx[0] + sy[1] + sz[0], fx[0] + fy[1]));
      if (VEC == 4) a_m = add2(a_m, make_float2(d1[VEC - 1] + m2[VEC - 1] + sx[VEC - 1] + sy[VEC - 2], sz[VEC - 1] + fx[VEC - 1] + fy[VEC - 2]));
      continue;
    }
    const PoseC& ps = c_pose[slot][bl];
    const float* d2img = depth_2 + (size_t)(b0 + bl) * HW;
    const int y = p / W, x0 = p - y * W;
    const float yf = (float)y;
    float rn[3], ra[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rn[k] = fmaf(ps.nM1[3 * k + 1], yf, ps.nM1[3 * k + 2]);
      ra[k] = fmaf(ps.A[3 * k + 1], yf, ps.A[3 * k + 2]);
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int j = 2 * q;
      const float xf = (float)(x0 + j);
      const float2 nx = make_float2(-xf, -xf - 1.0f);
      const float2 dd = make_float2(d1[j], d1[j + 1]);
      PairOut o;
      pair_forward<true>(ps, GlobalTaps{d2img}, H, W, nx, -yf, make_float2(fx[j], fx[j + 1]), make_float2(fy[j], fy[j + 1]), dd,
                         make_float2(sx[j], sx[j + 1]), make_float2(sy[j], sy[j + 1]), make_float2(sz[j], sz[j + 1]), rn, ra, o);
      const float2 m = make_float2(mask1(cfg, m2[j], dd.x, o.wpc[2].x), mask1(cfg, m2[j + 1], dd.y, o.wpc[2].y));
      float2 fl, dl, sl;
      if (cfg.warm) fl = fma2(o.ex, o.ex, mul2(o.ey, o.ey));
      else fl = make_float2(fabsf(o.ex.x) + fabsf(o.ey.x), fabsf(o.ex.y) + fabsf(o.ey.y));
      dl = make_float2(disp_term(cfg, o.p12[2].x, o.wpc[2].x), disp_term(cfg, o.p12[2].y, o.wpc[2].y));
      sl = make_float2(fabsf(o.e[0].x) + fabsf(o.e[1].x) + fabsf(o.e[2].x), fabsf(o.e[0].y) + fabsf(o.e[1].y) + fabsf(o.e[2].y));
      a_flow = fma2(m, fl, a_flow);
      a_disp = fma2(m, dl, a_disp);
      a_sf = fma2(m, sl, a_sf);
      a_m = add2(a_m, m);
    }
  }
  float s_flow = warp_sum(a_flow.x + a_flow.y), s_disp = warp_sum(a_disp.x + a_disp.y);
  float s_sf = warp_sum(a_sf.x + a_sf.y), s_m = warp_sum(a_m.x + a_m.y);
  if (lane == 0) { red[warp][0] = s_flow; red[warp][1] = s_disp; red[warp][2] = s_sf; red[warp][3] = s_m; }
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");    // consumers only (the producer warp has exited)
  if (threadIdx.x < 4) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) a += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = a;
  }
}

// c * sign(v) (0 at v == 0), lane-wise
__device__ __forceinline__ float2 sgn_scale2(float2 v, float2 c) {
  const unsigned sx = __float_as_uint(c.x) ^ (__float_as_uint(v.x) & 0x80000000u);   // one LOP3 per lane
  const unsigned sy = __float_as_uint(c.y) ^ (__float_as_uint(v.y) & 0x80000000u);
  return make_float2(v.x == 0.f ? 0.f : __uint_as_float(sx), v.y == 0.f ? 0.f : __uint_as_float(sy));
}

// backward of one pixel pair: returns g_(P1 + sf) and scatters g_depth_2
template <class Taps, class Scat>
__device__ __forceinline__ void pair_backward(const PoseC& ps, const dvd_loss_cfg& cfg, const Taps& d2img,
                                              const Scat& scat, int H, int W, float2 nx, float nyf, float2 fx,
                                              float2 fy, float2 dd, float2 mm, float2 psx, float2 psy, float2 psz,
                                              const float (&rn)[3], const float (&ra)[3], float cf, float cd, float2& gv0,
                                              float2& gv1, float2& gv2) {
      PairOut o;
      if (cfg.second_is_disp) pair_forward<false>(ps, d2img, H, W, nx, nyf, fx, fy, dd, psx, psy, psz, rn, ra, o);
      else                    pair_forward<true>(ps, d2img, H, W, nx, nyf, fx, fy, dd, psx, psy, psz, rn, ra, o);
      const float2 m = make_float2(mask1(cfg, mm.x, dd.x, o.wpc[2].x), mask1(cfg, mm.y, dd.y, o.wpc[2].y));
      // --- flow term -> g_i12 (zero where the projection was rejected)
      const float2 mcf = make_float2(o.zok[0] ? m.x * cf : 0.f, o.zok[1] ? m.y * cf : 0.f);
      float2 gux, guy;
      if (cfg.warm) {
        const float2 t2 = add2(mcf, mcf);
        gux = mul2(t2, o.ex); guy = mul2(t2, o.ey);
      } else {
        gux = sgn_scale2(o.ex, mcf); guy = sgn_scale2(o.ey, mcf);
      }
      const float2 gi0 = mul2(gux, o.rz), gi1 = mul2(guy, o.rz);
      const float2 tt = fma2(gux, o.i12[0], mul2(guy, o.i12[1]));
      const float2 gi2 = mul2(mul2(tt, o.rz), mul2(o.rz, F2(-1.0f)));
      float2 gp0, gp1, gp2;   // g_p12 = K^T g_i12
      mtv2(ps.K, gi0, gi1, gi2, gp0, gp1, gp2);
      float2 hu, hv, h1;      // Kinv^T g_warped_p2_camera_2
      float2 ge0 = F2(0.f), ge1 = F2(0.f), ge2 = F2(0.f);
      const float2 mc = mul2(m, F2(cd));
      if (cfg.second_is_disp) {
        const float2 za = o.p12[2], zb = o.wpc[2];
        float2 gwc2;
        if (cfg.disp_mode == 0) {
          const float2 ra2 = make_float2(rcp_fast(fmaxf(za.x, 1e-3f)), rcp_fast(fmaxf(za.y, 1e-3f)));
          const float2 rb2 = make_float2(rcp_fast(fmaxf(zb.x, 1e-3f)), rcp_fast(fmaxf(zb.y, 1e-3f)));
          const float2 s = sgn_scale2(sub2(ra2, rb2), mul2(mc, F2(100.0f)));
          const float2 sa = make_float2(za.x >= 1e-3f ? s.x : 0.f, za.y >= 1e-3f ? s.y : 0.f);
          const float2 sb = make_float2(zb.x >= 1e-3f ? s.x : 0.f, zb.y >= 1e-3f ? s.y : 0.f);
          gp2 = fma2(mul2(sa, ra2), mul2(ra2, F2(-1.0f)), gp2);
          gwc2 = mul2(mul2(sb, rb2), rb2);
        } else if (cfg.disp_mode == 1) {
          float g2a[2], g2b[2];
          const float zas[2] = {za.x, za.y}, zbs[2] = {zb.x, zb.y}, mcs[2] = {mc.x, mc.y};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float a = fmaxf(zas[q], 1e-3f), bb = fmaxf(zbs[q], 1e-3f);
            const float ra1 = rcp_fast(a), rb1 = rcp_fast(bb);
            float ga, gb;
            if (a >= bb) { ga = rb1; gb = -a * rb1 * rb1; }
            else         { ga = -bb * ra1 * ra1; gb = ra1; }
            g2a[q] = zas[q] >= 1e-3f ? ga * mcs[q] : 0.f;
            g2b[q] = zbs[q] >= 1e-3f ? gb * mcs[q] : 0.f;
          }
          gp2 = add2(gp2, make_float2(g2a[0], g2a[1]));
          gwc2 = make_float2(g2b[0], g2b[1]);
        } else {
          const float2 s = sgn_scale2(sub2(za, zb), mc);
          gp2 = add2(gp2, s);
          gwc2 = mul2(s, F2(-1.0f));
        }
        hu = mul2(F2(ps.Kinv[6]), gwc2); hv = mul2(F2(ps.Kinv[7]), gwc2); h1 = mul2(F2(ps.Kinv[8]), gwc2);
      } else {
        ge0 = sgn_scale2(o.e[0], mc); ge1 = sgn_scale2(o.e[1], mc); ge2 = sgn_scale2(o.e[2], mc);
        // warped_global_p2 = R2 wpc + t2  =>  g_wpc = R2^T g_e
        float2 a0, a1, a2;
        mtv2(ps.R2, ge0, ge1, ge2, a0, a1, a2);
        mtv2(ps.Kinv, a0, a1, a2, hu, hv, h1);
      }
      // g_(P1 + sf) = R2 g_p12 ; the sf term adds -g_e to both P1 and sf
      mv2(ps.R2, gp0, gp1, gp2, gv0, gv1, gv2);
      gv0 = sub2(gv0, ge0); gv1 = sub2(gv1, ge1); gv2 = sub2(gv2, ge2);
      // scatter to depth_2: g_d2_k = w_k (hu u_k + hv v_k + h1), (u_k, v_k) = (x0f, y0f) (+1)
      if (scat.on()) {
        const float2 base = fma2(hu, o.x0f, fma2(hv, o.y0f, h1));
        const float2 bx = add2(base, hu);
        const float2 g0 = mul2(o.w[0], base), g1 = mul2(o.w[1], bx);
        const float2 g2 = mul2(o.w[2], add2(base, hv)), g3 = mul2(o.w[3], add2(bx, hv));
        const float ga[4] = {g0.x, g1.x, g2.x, g3.x}, gb[4] = {g0.y, g1.y, g2.y, g3.y};
        scat.add4(o.i00[0], o.sx1[0], o.sy1[0], ga);
        scat.add4(o.i00[1], o.sx1[1], o.sy1[1], gb);
      }
}

// staged backward (same producer / consumer ring as the staged forward), NP pixel pairs per consumer thread
template <int NP, int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads + 32, MINB) reproject_loss_bwd_staged_kernel(
    const float* __restrict__ depth_1, const float* __restrict__ depth_2, const float* __restrict__ flow,
    const float* __restrict__ mask_2, const float* __restrict__ sf, dvd_loss_cfg cfg, const float* __restrict__ scalars,
    float gscale, const float* __restrict__ gscale_dev, float* __restrict__ g_sf, float* __restrict__ g_d2, int H, int W,
    int slot, int b0, int nb, int tiles_per_pair) {
  DVD_PDL_ENTER();
  using namespace tc;
  constexpr int VEC = 2 * NP, TILE = StagedCfg<NP>::TILE, FLOATS = StagedCfg<NP>::FLOATS;
  extern __shared__ __align__(128) float stage_mem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = H * W, ntiles = nb * tiles_per_pair;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kThreads / 32); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == kThreads / 32) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int bl = tile / tiles_per_pair, t = tile - bl * tiles_per_pair;
        produce_tile<TILE>(stage_mem + (size_t)s * FLOATS, &full[s], depth_1, mask_2, sf, flow, (size_t)(b0 + bl), HW, t * TILE, HW);
      }
    }
    return;
  }
  const float gs = gscale * (gscale_dev ? __ldg(gscale_dev) : 1.0f);
  const float cf = __ldg(scalars + DVD_S_CF) * gs;
  const float cd = __ldg(scalars + DVD_S_CD) * gs;
  const int tv = threadIdx.x * VEC;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
    mbar_wait(&full[s], ph);
    float d1[VEC], m2[VEC], sx[VEC], sy[VEC], sz[VEC], fx[VEC], fy[VEC], ox[VEC], oy[VEC], oz[VEC];
    consume_tile<NP>(stage_mem + (size_t)s * FLOATS, tv, d1, m2, sx, sy, sz, fx, fy);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    const int bl = tile / tiles_per_pair, t = tile - bl * tiles_per_pair;
    const int p = t * TILE + tv;
    if (p >= HW) continue;
    const PoseC& ps = c_pose[slot][bl];
    const size_t b = (size_t)(b0 + bl);
    const float* d2img = depth_2 + b * HW;
    float* gd2img = g_d2 ? g_d2 + b * HW : nullptr;
    const int y = p / W, x0 = p - y * W;
    const float yf = (float)y;
    float rn[3], ra[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rn[k] = fmaf(ps.nM1[3 * k + 1], yf, ps.nM1[3 * k + 2]);
      ra[k] = fmaf(ps.A[3 * k + 1], yf, ps.A[3 * k + 2]);
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int j = 2 * q;
      const float xf = (float)(x0 + j);
      const float2 nx = make_float2(-xf, -xf - 1.0f);
      float2 gv0, gv1, gv2;
      pair_backward(ps, cfg, GlobalTaps{d2img}, GlobalScatter{gd2img}, H, W, nx, -yf, make_float2(fx[j], fx[j + 1]), make_float2(fy[j], fy[j + 1]),
                    make_float2(d1[j], d1[j + 1]), make_float2(m2[j], m2[j + 1]), make_float2(sx[j], sx[j + 1]),
                    make_float2(sy[j], sy[j + 1]), make_float2(sz[j], sz[j + 1]), rn, ra, cf, cd, gv0, gv1, gv2);
      ox[j] = gv0.x; ox[j + 1] = gv0.y; oy[j] = gv1.x; oy[j + 1] = gv1.y; oz[j] = gv2.x; oz[j + 1] = gv2.y;
    }
    float* gsp = g_sf + b * 3 * HW + p;
    store_vec<VEC>(gsp, ox);
    store_vec<VEC>(gsp + HW, oy);
    store_vec<VEC>(gsp + 2 * (size_t)HW, oz);
  }
}

// ---------------------------------------------------------------------------------------------
static int pick_vec(int B, int H, int W, std::initializer_list<const void*> ptrs, int max_vec = 4) {
  bool al = true;
  for (const void* p : ptrs) al = al && (p == nullptr || aligned16(p));
  long total = (long)B * H * W;
#ifdef DVD_PROFILING
  if (const char* ev = getenv("DVD_REPROJECT_VEC")) {   // tuning override: 1, 2 or 4
    int v = atoi(ev);
    if ((v == 4 && al && W % 4 == 0) || (v == 2 && al && W % 2 == 0)) return v;
    if (v == 1) return 1;
  }
#endif
  // wide vectors only when enough threads remain to cover HBM latency (>= 512 / 256 threads per SM)
  if (max_vec >= 4 && al && W % 4 == 0 && total / 4 >= (long)num_sms() * 512) return 4;
  if (max_vec >= 2 && al && W % 2 == 0 && total / 2 >= (long)num_sms() * 256) return 2;
  return 1;
}

// ~8 CTAs of 256 threads per SM over the whole grid, split evenly over the pairs. Measured on B200: this mild
// over-subscription (2.7 waves at 3 resident CTAs/SM) beats a single exactly-resident wave (75 vs 89 us forward,
// 134 vs 173 us backward at 64 pairs): short CTAs retire and refill continuously instead of finishing together.
template <typename K>
static dim3 grid_for(K, int B, int items_per_pair) {
  int per_pair = (items_per_pair + kThreads - 1) / kThreads;
  int cap = (num_sms() * 8 + B - 1) / B;
  if (cap < 1) cap = 1;
  if (per_pair > cap) per_pair = cap;
  if (per_pair < 1) per_pair = 1;
  return dim3((unsigned)per_pair, (unsigned)B, 1);
}
// upper bound used to size the partial-sum scratch (8 CTAs per SM is the hardware maximum for 256 threads)
static dim3 grid_bound(int B, int items_per_pair) {
  int per_pair = (items_per_pair + kThreads - 1) / kThreads;
  int cap = (num_sms() * 8 + B - 1) / B;
  if (per_pair > cap) per_pair = cap;
  if (per_pair < 1) per_pair = 1;
  return dim3((unsigned)per_pair, (unsigned)B, 1);
}

// Stage the derived poses of pairs [b0, b0 + nb) into a constant-memory slot (all on `st`, no host synchronisation).
// Slots rotate; a slot is handed out again only after the kernel that last read it has finished: every user records a
// per-slot event behind its kernel (release_pose_slot) and the next owner makes its stream wait on that event before it
// overwrites the slot, so calls in flight on any number of streams (or more than kPoseSlots deep) cannot see each other's
// poses. Symbol address and events are per device.
struct PoseSlots {
  PoseC* stage = nullptr;
  cudaEvent_t ev[kPoseSlots] = {};
  bool used[kPoseSlots] = {};
  unsigned ctr = 0;
  bool ok = false;
};
static std::mutex g_pose_mu;
static PoseSlots* pose_slots() {     // call with g_pose_mu held
  static PoseSlots per_dev[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  PoseSlots& S = per_dev[dev];
  if (!S.ok) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_pose_stage) != cudaSuccess) return nullptr;
    S.stage = static_cast<PoseC*>(p);
    for (int i = 0; i < kPoseSlots; ++i)
      if (cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    S.ok = true;
  }
  return &S;
}
static int stage_poses(const float* poses, int b0, int nb, cudaStream_t st, int* slot_out) {
  std::lock_guard<std::mutex> lk(g_pose_mu);
  PoseSlots* S = pose_slots();
  DVD_ARG_CHECK(S != nullptr, "pose staging: cudaGetSymbolAddress / cudaEventCreate failed");
  const int slot = (int)(S->ctr++ % (unsigned)kPoseSlots);
  // (inside a stream capture the launches of the graph are ordered on the capture stream and replays are ordered on their
  // stream: events from outside the capture must not be waited on there)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  DVD_CUDA_CALL(cudaStreamIsCapturing(st, &cap));
  if (cap == cudaStreamCaptureStatusNone && S->used[slot])
    DVD_CUDA_CALL(cudaStreamWaitEvent(st, S->ev[slot], 0));   // previous reader of this slot is done
  dvd::launch(pose_prep_kernel, 1, kPosePairs, 0, st, poses + (size_t)b0 * DVD_POSE_STRIDE, nb, slot);
  DVD_CUDA_LAUNCH_CHECK("pose_prep");
  DVD_CUDA_CALL(cudaMemcpyToSymbolAsync(c_pose, S->stage + (size_t)slot * kPosePairs, (size_t)nb * sizeof(PoseC),
                                        (size_t)slot * kPosePairs * sizeof(PoseC), cudaMemcpyDeviceToDevice, st));
  *slot_out = slot;
  return 0;
}
// record "the kernel reading `slot` has been enqueued on `st`" (call right after that launch)
static int release_pose_slot(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_pose_mu);
  PoseSlots* S = pose_slots();
  DVD_ARG_CHECK(S != nullptr, "pose staging unavailable");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  DVD_CUDA_CALL(cudaStreamIsCapturing(st, &cap));
  if (cap != cudaStreamCaptureStatusNone) return 0;
  DVD_CUDA_CALL(cudaEventRecord(S->ev[slot], st));
  S->used[slot] = true;
  return 0;
}

static int check_shape(int B, int H, int W) {
  DVD_ARG_CHECK(B >= 1 && H >= 2 && W >= 2, "bad shape B=%d H=%d W=%d (need B>=1, H,W>=2)", B, H, W);
  DVD_ARG_CHECK(B <= 65535, "B=%d exceeds gridDim.y", B);
  DVD_ARG_CHECK((long)H * W < (1L << 30), "image too large");
  return 0;
}

}  // namespace dvd

using namespace dvd;

extern "C" int dvd_reproject_partials_size(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  // upper bound over every VEC choice
  dim3 g = grid_bound(B, H * W);
  long quads = (long)g.x * g.y;
  // staged forward: one quad per persistent CTA, per chunk of kPosePairs pairs
  const long staged = (long)((B + kPosePairs - 1) / kPosePairs) * 3 * num_sms();
  if (staged > quads) quads = staged;
  return (int)(quads * 4);
}

extern "C" int dvd_unproject_fwd(const float* depth, const float* poses, float* P, int B, int H, int W, int which,
                                 void* stream) {
  if (int e = check_shape(B, H, W)) return e;
  DVD_ARG_CHECK(depth && poses && P, "null pointer");
  DVD_ARG_CHECK(which == 1 || which == 2, "which must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  int vec = pick_vec(B, H, W, {depth, P});
  const int ipp = H * W / vec;
  if (vec == 4) dvd::launch(unproject_fwd_kernel<4>, grid_for(unproject_fwd_kernel<4>, B, ipp), kThreads, 0, st, depth, poses, P, H, W, which);
  else if (vec == 2) dvd::launch(unproject_fwd_kernel<2>, grid_for(unproject_fwd_kernel<2>, B, ipp), kThreads, 0, st, depth, poses, P, H, W, which);
  else dvd::launch(unproject_fwd_kernel<1>, grid_for(unproject_fwd_kernel<1>, B, ipp), kThreads, 0, st, depth, poses, P, H, W, which);
  DVD_CUDA_LAUNCH_CHECK("unproject_fwd");
  return 0;
}

extern "C" int dvd_unproject_bwd(const float* gP, const float* poses, float* gdepth, int B, int H, int W, int which,
                                 void* stream) {
  if (int e = check_shape(B, H, W)) return e;
  DVD_ARG_CHECK(gP && poses && gdepth, "null pointer");
  DVD_ARG_CHECK(which == 1 || which == 2, "which must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  int vec = pick_vec(B, H, W, {gP, gdepth});
  const int ipp = H * W / vec;
  if (vec == 4) dvd::launch(unproject_bwd_kernel<4>, grid_for(unproject_bwd_kernel<4>, B, ipp), kThreads, 0, st, gP, poses, gdepth, H, W, which);
  else if (vec == 2) dvd::launch(unproject_bwd_kernel<2>, grid_for(unproject_bwd_kernel<2>, B, ipp), kThreads, 0, st, gP, poses, gdepth, H, W, which);
  else dvd::launch(unproject_bwd_kernel<1>, grid_for(unproject_bwd_kernel<1>, B, ipp), kThreads, 0, st, gP, poses, gdepth, H, W, which);
  DVD_CUDA_LAUNCH_CHECK("unproject_bwd");
  return 0;
}

// cudaFuncSetAttribute once per (function, device)
static int set_smem_once(const void* fn, int smem) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  DVD_CUDA_CALL(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  for (auto& d : done)
    if (d.first == fn && d.second == dev) return 0;
  DVD_CUDA_CALL(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  done.emplace_back(fn, dev);
  return 0;
}

static int check_cfg(const dvd_loss_cfg* cfg) {
  DVD_ARG_CHECK(cfg != nullptr, "null loss cfg");
  DVD_ARG_CHECK(cfg->disp_mode >= 0 && cfg->disp_mode <= 2, "disp_mode must be 0,1,2");
  return 0;
}

extern "C" int dvd_reproject_loss_fwd(const float* depth_1, const float* depth_2, const float* flow_1_2,
                                      const float* mask_2, const float* sf, const float* poses,
                                      const dvd_loss_cfg* cfg, float* partials, float* scalars, int B, int H,
                                      int W, void* stream) {
  if (int e = check_shape(B, H, W)) return e;
  if (int e = check_cfg(cfg)) return e;
  DVD_ARG_CHECK(depth_1 && depth_2 && flow_1_2 && mask_2 && sf && poses && partials && scalars, "null pointer");
  DVD_ARG_CHECK(aligned16(partials), "partials must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int vec = pick_vec(B, H, W, {depth_1, mask_2, sf, flow_1_2});
  const int ipp = H * W / vec;
  dim3 g;
  // DVD_REPROJECT_SCALAR=1 forces the generic scalar kernels (any shape); DVD_REPROJECT_DRY=1 is a memory-system probe
  // DVD_REPROJECT_SCALAR=1 (any build) forces the generic scalar kernels - a cross-check with identical results; the
  // arithmetic-free memory probe exists in -DDVD_PROFILING builds only
  static const bool packed = !(getenv("DVD_REPROJECT_SCALAR") && atoi(getenv("DVD_REPROJECT_SCALAR")));
#ifdef DVD_PROFILING
  static const bool dry = getenv("DVD_REPROJECT_DRY") && atoi(getenv("DVD_REPROJECT_DRY"));
#else
  constexpr bool dry = false;
#endif
  if (vec == 4 && packed) {
    // packed-FP32 kernel, bulk-async staged inputs: 1 pixel pair per consumer thread, 4-deep ring, 3 CTAs per SM
    // (fastest of the measured variants, profiles/r1_ncu_reproject_variants.txt)
    const void* fn = dry ? (const void*)reproject_loss_fwd_staged_kernel<1, 4, 3, true>
                         : (const void*)reproject_loss_fwd_staged_kernel<1, 4, 3, false>;
    constexpr int np = 1, stages = 4, ctas = 3;
    const int tile = kThreads * 2 * np, smem = stages * tile * 7 * 4;
    if (int e = set_smem_once(fn, smem)) return e;
    const int tiles_per_pair = (H * W + tile - 1) / tile;
    unsigned nq = 0;
    for (int b0 = 0; b0 < B; b0 += kPosePairs) {
      int nb = B - b0 < kPosePairs ? B - b0 : kPosePairs;
      int slot = 0;
      if (int e = stage_poses(poses, b0, nb, st, &slot)) return e;
      int gx = nb * tiles_per_pair;
      if (gx > ctas * num_sms()) gx = ctas * num_sms();
      float* part = partials + (size_t)nq * 4;
      int b0v = b0, tpp = tiles_per_pair, Hh = H, Ww = W;
      dvd_loss_cfg cfgv = *cfg;
      void* args[] = {(void*)&depth_1, (void*)&depth_2, (void*)&flow_1_2, (void*)&mask_2, (void*)&sf, (void*)&cfgv, (void*)&part,
                      (void*)&Hh, (void*)&Ww, (void*)&slot, (void*)&b0v, (void*)&nb, (void*)&tpp};
      DVD_CUDA_CALL(cudaLaunchKernel(fn, dim3((unsigned)gx), dim3(kThreads + 32), args, (size_t)smem, st));
      if (int e = release_pose_slot(slot, st)) return e;
      nq += (unsigned)gx;
    }
    g = dim3(nq, 1, 1);
  } else {
    if (vec == 4) g = grid_for(reproject_loss_fwd_kernel<4, 3>, B, ipp);
    else if (vec == 2) g = grid_for(reproject_loss_fwd_kernel<2, 4>, B, ipp);
    else g = grid_for(reproject_loss_fwd_kernel<1, 4>, B, ipp);
    if (vec == 4) dvd::launch(reproject_loss_fwd_kernel<4, 3>, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, mask_2, sf, poses, *cfg, partials, H, W);
    else if (vec == 2) dvd::launch(reproject_loss_fwd_kernel<2, 4>, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, mask_2, sf, poses, *cfg, partials, H, W);
    else dvd::launch(reproject_loss_fwd_kernel<1, 4>, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, mask_2, sf, poses, *cfg, partials, H, W);
  }
  DVD_CUDA_LAUNCH_CHECK("reproject_loss_fwd");
  dvd::launch(reproject_finalize_kernel, 1, 256, 0, st, partials, (int)(g.x * g.y), *cfg, scalars);
  DVD_CUDA_LAUNCH_CHECK("reproject_finalize");
  return 0;
}

extern "C" int dvd_reproject_loss_bwd(const float* depth_1, const float* depth_2, const float* flow_1_2,
                                      const float* mask_2, const float* sf, const float* poses,
                                      const dvd_loss_cfg* cfg, const float* scalars, float gscale, const float* gscale_dev,
                                      float* g_sf, float* g_depth_2, int B, int H, int W, void* stream) {
  if (int e = check_shape(B, H, W)) return e;
  if (int e = check_cfg(cfg)) return e;
  DVD_ARG_CHECK(depth_1 && depth_2 && flow_1_2 && mask_2 && sf && poses && scalars && g_sf, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_depth_2) DVD_CUDA_CALL(cudaMemsetAsync(g_depth_2, 0, (size_t)B * H * W * sizeof(float), st));
#ifdef DVD_PROFILING
  if (getenv("DVD_REPROJECT_BWD_NOSCATTER")) g_depth_2 = nullptr;   // profiling probe only: skip the scatter-add
#endif
  // measured on B200 (profiles/r1_reproject_vec_sweep.txt): the scatter-add backward is fastest with one pixel per
  // thread (1.97 TB/s vs 1.28 / 1.46 for 2 / 4): more warps in flight hide the red.global latency
  static const bool packed = !(getenv("DVD_REPROJECT_SCALAR") && atoi(getenv("DVD_REPROJECT_SCALAR")));
  const int vecp = pick_vec(B, H, W, {depth_1, mask_2, sf, flow_1_2, g_sf}, 4);
  if (packed && vecp == 4) {
    // packed-FP32 kernel, bulk-async staged inputs: 1 pixel pair per consumer thread, 4-deep ring, 2 CTAs per SM
    const void* fn = (const void*)reproject_loss_bwd_staged_kernel<1, 4, 2>;
    constexpr int np = 1, stages = 4, ctas = 2;
    const int tile = kThreads * 2 * np, smem = stages * tile * 7 * 4;
    if (int e = set_smem_once(fn, smem)) return e;
    const int tiles_per_pair = (H * W + tile - 1) / tile;
    for (int b0 = 0; b0 < B; b0 += kPosePairs) {
      int nb = B - b0 < kPosePairs ? B - b0 : kPosePairs;
      int slot = 0;
      if (int e = stage_poses(poses, b0, nb, st, &slot)) return e;
      int gx = nb * tiles_per_pair;
      if (gx > ctas * num_sms()) gx = ctas * num_sms();
      int b0v = b0, tpp = tiles_per_pair, Hh = H, Ww = W;
      dvd_loss_cfg cfgv = *cfg;
      void* args[] = {(void*)&depth_1, (void*)&depth_2, (void*)&flow_1_2, (void*)&mask_2, (void*)&sf, (void*)&cfgv, (void*)&scalars,
                      (void*)&gscale, (void*)&gscale_dev, (void*)&g_sf, (void*)&g_depth_2, (void*)&Hh, (void*)&Ww, (void*)&slot,
                      (void*)&b0v, (void*)&nb, (void*)&tpp};
      DVD_CUDA_CALL(cudaLaunchKernel(fn, dim3((unsigned)gx), dim3(kThreads + 32), args, (size_t)smem, st));
      if (int e = release_pose_slot(slot, st)) return e;
    }
    return 0;
  }
  int vec = pick_vec(B, H, W, {depth_1, mask_2, sf, flow_1_2, g_sf}, 1);
  const int ipp = H * W / vec;
  dim3 g = vec == 4 ? grid_for(reproject_loss_bwd_kernel<4>, B, ipp)
                    : (vec == 2 ? grid_for(reproject_loss_bwd_kernel<2>, B, ipp) : grid_for(reproject_loss_bwd_kernel<1>, B, ipp));
  if (vec == 4) dvd::launch(reproject_loss_bwd_kernel<4>, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, mask_2, sf, poses, *cfg, scalars, gscale, gscale_dev, g_sf, g_depth_2, H, W);
  else if (vec == 2) dvd::launch(reproject_loss_bwd_kernel<2>, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, mask_2, sf, poses, *cfg, scalars, gscale, gscale_dev, g_sf, g_depth_2, H, W);
  else dvd::launch(reproject_loss_bwd_kernel<1>, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, mask_2, sf, poses, *cfg, scalars, gscale, gscale_dev, g_sf, g_depth_2, H, W);
  DVD_CUDA_LAUNCH_CHECK("reproject_loss_bwd");
  return 0;
}

extern "C" int dvd_reproject_materialize(const float* depth_1, const float* depth_2, const float* flow_1_2,
                                         const float* sf, const float* poses, float* global_p1, float* sf_by_depth,
                                         float* warped_global_p2, float* warped_p2_camera_2, float* p1_camera_2,
                                         float* dflow_1_2, float* staticflow_1_2, float* depth_image_1_2,
                                         float* depth_warp_1_2, int B, int H, int W, void* stream) {
  if (int e = check_shape(B, H, W)) return e;
  DVD_ARG_CHECK(depth_1 && depth_2 && flow_1_2 && poses, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g = grid_for(reproject_materialize_kernel, B, H * W);
  dvd::launch(reproject_materialize_kernel, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, sf, poses, global_p1, sf_by_depth,
                                                       warped_global_p2, warped_p2_camera_2, p1_camera_2, dflow_1_2,
                                                       staticflow_1_2, depth_image_1_2, depth_warp_1_2, H, W);
  DVD_CUDA_LAUNCH_CHECK("reproject_materialize");
  return 0;
}

extern "C" int dvd_reproject_materialize_bwd(const float* depth_1, const float* depth_2, const float* flow_1_2,
                                             const float* sf, const float* poses, const float* g_global_p1,
                                             const float* g_sf_by_depth, const float* g_warped_global_p2,
                                             const float* g_warped_p2_camera_2, const float* g_p1_camera_2,
                                             const float* g_dflow_1_2, const float* g_staticflow_1_2,
                                             const float* g_depth_image_1_2, const float* g_depth_warp_1_2,
                                             float* g_depth_1, float* g_depth_2, float* g_sf, int B, int H, int W,
                                             void* stream) {
  if (int e = check_shape(B, H, W)) return e;
  DVD_ARG_CHECK(depth_1 && depth_2 && flow_1_2 && poses, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_depth_2) DVD_CUDA_CALL(cudaMemsetAsync(g_depth_2, 0, (size_t)B * H * W * sizeof(float), st));
  MatGrads G{g_global_p1, g_sf_by_depth, g_warped_global_p2, g_warped_p2_camera_2, g_p1_camera_2,
             g_dflow_1_2, g_staticflow_1_2, g_depth_image_1_2, g_depth_warp_1_2};
  dim3 g = grid_for(reproject_materialize_bwd_kernel, B, H * W);
  dvd::launch(reproject_materialize_bwd_kernel, g, kThreads, 0, st, depth_1, depth_2, flow_1_2, sf, poses, G, g_depth_1, g_depth_2,
                                                         g_sf, H, W);
  DVD_CUDA_LAUNCH_CHECK("reproject_materialize_bwd");
  return 0;
}
