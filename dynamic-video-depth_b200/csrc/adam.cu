// Fused flat Adam (O1): replaces the two torch.optim.Adam instances of the reference
// (models/netinterface.py:96-97,127-129; models/scene_flow_motion_field.py:113-115,212-213), which loop
// over ~430 parameter tensors. One launch over the flat fp32 parameter buffer of a net:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// (torch.optim.Adam, amsgrad=False, weight_decay=0). HBM-bound: 28 B/param (read g,p,m,v; write p,m,v).
// `gscale` multiplies the gradient first (1/world_size after the NCCL sum all-reduce).
#include "common.cuh"

namespace dvd {

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long n, float lr,
                                                        float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                                        float gscale, const float* __restrict__ bc_dev) {
  DVD_PDL_ENTER();
  if (bc_dev) { bc1 = bc_dev[1]; bc2_sqrt = bc_dev[2]; }      // step counter kept on the device (CUDA-graph replays)
  const float step_size = lr / bc1;
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 gg = ldg_stream4(g + 4 * i);
    float4 pp = *reinterpret_cast<const float4*>(p + 4 * i);
    float4 mm = *reinterpret_cast<const float4*>(m + 4 * i);
    float4 vv = *reinterpret_cast<const float4*>(v + 4 * i);
    float* gp = &gg.x; float* ppp = &pp.x; float* mp = &mm.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gp[k] * gscale;
      mp[k] = b1 * mp[k] + (1.0f - b1) * gk;
      vp[k] = b2 * vp[k] + (1.0f - b2) * gk * gk;
      float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      ppp[k] -= step_size * (mp[k] / denom);
    }
    *reinterpret_cast<float4*>(p + 4 * i) = pp;
    *reinterpret_cast<float4*>(m + 4 * i) = mm;
    *reinterpret_cast<float4*>(v + 4 * i) = vv;
  }
  // tail
  for (long i = n4 * 4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gk = g[i] * gscale;
    float mk = b1 * m[i] + (1.0f - b1) * gk;
    float vk = b2 * v[i] + (1.0f - b2) * gk * gk;
    m[i] = mk; v[i] = vk;
    p[i] -= step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
  }
}

// device-side step counter: state = {step (as float bits of an int), 1 - b1^step, sqrt(1 - b2^step)}
__global__ void adam_tick_kernel(float* __restrict__ state, float b1, float b2) {
  DVD_PDL_ENTER();
  const int step = __float_as_int(state[0]) + 1;
  state[0] = __int_as_float(step);
  state[1] = (float)(1.0 - pow((double)b1, (double)step));
  state[2] = (float)sqrt(1.0 - pow((double)b2, (double)step));
}

}  // namespace dvd

extern "C" int dvd_adam_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                             float eps, int step, float gscale, void* stream) {
  DVD_ARG_CHECK(p && g && m && v && n > 0, "bad arguments");
  DVD_ARG_CHECK(step >= 1, "step counts from 1");
  DVD_ARG_CHECK(dvd::aligned16(p) && dvd::aligned16(g) && dvd::aligned16(m) && dvd::aligned16(v), "buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  long blocks = (n / 4 + 255) / 256;
  long cap = (long)dvd::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dvd::launch(dvd::adam_flat_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, (float)bc1,
                                                                        (float)sqrt(bc2), gscale, nullptr);
  DVD_CUDA_LAUNCH_CHECK("adam_flat");
  return 0;
}

extern "C" int dvd_adam_flat_dev(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                                 float eps, float* step_state, float gscale, void* stream) {
  DVD_ARG_CHECK(p && g && m && v && step_state && n > 0, "bad arguments");
  DVD_ARG_CHECK(dvd::aligned16(p) && dvd::aligned16(g) && dvd::aligned16(m) && dvd::aligned16(v), "buffers must be 16-byte aligned");
  long blocks = (n / 4 + 255) / 256;
  long cap = (long)dvd::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dvd::launch(dvd::adam_tick_kernel, 1, 1, 0, (cudaStream_t)stream, step_state, beta1, beta2);
  DVD_CUDA_LAUNCH_CHECK("adam_tick");
  dvd::launch(dvd::adam_flat_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, 1.f, 1.f, gscale,
                                                                        step_state);
  DVD_CUDA_LAUNCH_CHECK("adam_flat");
  return 0;
}
