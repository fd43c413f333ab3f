// Flat-buffer fused AdamW and sum-of-squares (global-norm clipping), HBM-bound.
// Arithmetic follows the reference's optim/adamw.py:80-104: bias-corrected step size computed on
// the host, eps added to sqrt(v) (not inside), decoupled weight decay applied AFTER the Adam
// update on the already-updated parameter with the un-corrected lr.
#include "common.h"
#include "ptx.cuh"

namespace hero {

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, __nv_bfloat16* __restrict__ p_bf16, long long n, float step_size,
             float beta1, float beta2, float eps, float lr_wd, float grad_scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gr * gr;
    float pi = p[i] - step_size * (mi / (sqrtf(vi) + eps));
    if (lr_wd > 0.0f) pi = pi - lr_wd * pi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16(pi);
  }
}

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float t = x[i];
    s = fmaf(t, t, s);
  }
  s = warp_sum(s);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = sm[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// dst[i] = (dst[i] + sum_s slots[s * stride + i]) * scale: the reduction step of the copy-engine
// gradient exchange (peers have written their contributions into `slots`).
__global__ void reduce_slots_kernel(float* __restrict__ dst, const float* __restrict__ slots,
                                    int n_slots, long long stride, long long n4, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(dst)[i];
    for (int s = 0; s < n_slots; ++s) {
      const float4 v = reinterpret_cast<const float4*>(slots + s * stride)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(dst)[i] = acc;
  }
}

}  // namespace hero

using namespace hero;

extern "C" int hero_reduce_slots_f32(float* dst, const float* slots, int32_t n_slots,
                                     int64_t slot_stride, int64_t n, float scale,
                                     int32_t max_ctas, void* stream) {
  HERO_REQUIRE(dst && (slots || n_slots == 0) && n >= 0 && n % 4 == 0 && slot_stride % 4 == 0,
               "reduce_slots: bad args (n and stride must be multiples of 4)");
  if (n == 0) return HERO_OK;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = max_ctas > 0 ? max_ctas : 64;
  if (blocks > cap) blocks = cap;
  reduce_slots_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      dst, slots, n_slots, slot_stride, n / 4, scale);
  HERO_LAUNCH_CHECK();
  return HERO_OK;
}

extern "C" int hero_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16,
                               int64_t n, float step_size, float beta1, float beta2, float eps,
                               float lr_wd, float grad_scale, void* stream) {
  HERO_REQUIRE(p && g && m && v && n >= 0, "adamw: bad args");
  if (n == 0) return HERO_OK;
  const int sms = sm_count();
  if (sms <= 0) return set_error(HERO_ERR_NO_DEVICE, "no CUDA device");
  long long blocks = (n + 255) / 256;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  adamw_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, step_size, beta1, beta2, eps, lr_wd,
      grad_scale);
  HERO_LAUNCH_CHECK();
  return HERO_OK;
}

extern "C" int hero_sumsq_f32(const float* x, int64_t n, float* out, void* stream) {
  HERO_REQUIRE(x && out && n >= 0, "sumsq: bad args");
  if (n == 0) return HERO_OK;
  const int sms = sm_count();
  if (sms <= 0) return set_error(HERO_ERR_NO_DEVICE, "no CUDA device");
  long long blocks = (n + 255) / 256;
  if (blocks > sms * 8LL) blocks = sms * 8LL;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, out);
  HERO_LAUNCH_CHECK();
  return HERO_OK;
}
