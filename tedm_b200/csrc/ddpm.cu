// DDPM arithmetic around the UNet: q_sample, L1/p2 loss, one ancestral sampling update.
// All fp32, memory-bound, vectorised; arithmetic order follows the reference so results are
// bit-exact (q_sample) or within an ulp (the rest).
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// q_sample                                   models/diffusion_model.py:176-203, utils.py:28-29,48-59
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float q_one(float x0, float nz, float a, float b, int normalize) {
  // reference: x0*2-1 (two ops), then a*x0 + b*noise as mul, mul, add -- no fma contraction.
  if (normalize) x0 = __fadd_rn(__fmul_rn(x0, 2.0f), -1.0f);
  return __fadd_rn(__fmul_rn(a, x0), __fmul_rn(b, nz));
}

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       const int64_t* __restrict__ t, const float* __restrict__ sa,
                                                       const float* __restrict__ sb, float* __restrict__ x_t, int chw,
                                                       int chw4, int normalize) {
  const int b = blockIdx.y;
  const int64_t tb = t[b];
  const float a = sa[tb], c = sb[tb];
  const size_t base = (size_t)b * chw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chw4; i += gridDim.x * blockDim.x) {
    float4 x = *reinterpret_cast<const float4*>(x0 + base + 4 * (size_t)i);
    float4 n = *reinterpret_cast<const float4*>(noise + base + 4 * (size_t)i);
    float4 o;
    o.x = q_one(x.x, n.x, a, c, normalize);
    o.y = q_one(x.y, n.y, a, c, normalize);
    o.z = q_one(x.z, n.z, a, c, normalize);
    o.w = q_one(x.w, n.w, a, c, normalize);
    *reinterpret_cast<float4*>(x_t + base + 4 * (size_t)i) = o;
  }
  // tail (chw not a multiple of 4)
  if (blockIdx.x == 0)
    for (int i = chw4 * 4 + threadIdx.x; i < chw; i += blockDim.x)
      x_t[base + i] = q_one(x0[base + i], noise[base + i], a, c, normalize);
}

extern "C" int tedm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                             const float* sqrt_1m_ac, float* x_t, int batch, int chw, int T, int normalize,
                             tedm_stream_t stream) {
  TEDM_CHECK_ARG(x0 && noise && t && sqrt_ac && sqrt_1m_ac && x_t, "tedm_q_sample: null pointer");
  TEDM_CHECK_ARG(batch > 0 && chw > 0 && T > 0, "tedm_q_sample: bad sizes batch=%d chw=%d T=%d", batch, chw, T);
  TEDM_CHECK_ARG(batch <= 65535, "tedm_q_sample: batch %d > 65535", batch);
  // vector path needs 16-byte aligned rows; otherwise everything goes through the scalar tail
  const bool vec = (chw % 4 == 0) && ((((uintptr_t)x0 | (uintptr_t)noise | (uintptr_t)x_t) & 15) == 0);
  const int chw4 = vec ? chw / 4 : 0;
  int gx = ceil_div(chw4 > 0 ? chw4 : 1, 256);
  if (gx > 64) gx = 64;
  q_sample_kernel<<<dim3(gx, batch), 256, 0, (cudaStream_t)stream>>>(x0, noise, t, sqrt_ac, sqrt_1m_ac, x_t, chw, chw4,
                                                                     normalize);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// L1 loss with p2 weighting                                  models/diffusion_model.py:138-143
// ------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < THREADS / 32) ? red[l] : 0.0f;
    v = warp_sum(v);
    if (l == 0) red[0] = v;
  }
  __syncthreads();
  v = red[0];
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(1024) l1_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                       const int64_t* __restrict__ t, const float* __restrict__ p2w,
                                                       float* __restrict__ per_image, float* __restrict__ grad, int chw,
                                                       int batch) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float w = p2w[t[b]];
  const float gscale = w / ((float)chw * (float)batch);
  const size_t base = (size_t)b * chw;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < chw; i += 1024) {
    const float d = pred[base + i] - target[base + i];
    acc += fabsf(d);
    if (grad) grad[base + i] = d > 0.0f ? gscale : (d < 0.0f ? -gscale : 0.0f);
  }
  acc = block_sum<1024>(acc, red);
  if (threadIdx.x == 0) per_image[b] = (acc / (float)chw) * w;
}

__global__ void mean_kernel(const float* __restrict__ v, float* __restrict__ out, int n) {
  float acc = 0.0f;
  for (int i = threadIdx.x; i < n; i += 32) acc += v[i];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) out[0] = acc / (float)n;
}

extern "C" int tedm_l1_loss(const float* pred, const float* target, const int64_t* t, const float* p2_weight,
                            float* loss_per_image, float* loss, float* grad, int batch, int chw, int T,
                            tedm_stream_t stream) {
  TEDM_CHECK_ARG(pred && target && t && p2_weight && loss_per_image && loss, "tedm_l1_loss: null pointer");
  TEDM_CHECK_ARG(batch > 0 && chw > 0 && T > 0, "tedm_l1_loss: bad sizes");
  l1_loss_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(pred, target, t, p2_weight, loss_per_image, grad, chw, batch);
  TEDM_LAUNCH_CHECK();
  mean_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_per_image, loss, batch);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// one reverse-diffusion update                                models/diffusion_model.py:205-286
// One CTA per image.  The dynamic threshold s = quantile(|x0_hat|, q) is EXACT: an 8-bit x 4 pass
// radix select over the fp32 bit patterns (non-negative floats order like their bits) finds the
// order statistic of rank k_lo, the next one is either the same value or the minimum of the larger
// ones; torch.quantile's linear interpolation is then applied with the host-computed weight.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float x0hat_at(const float* x_t, const float* eps, size_t i, float cr, float crm1) {
  return __fsub_rn(__fmul_rn(cr, x_t[i]), __fmul_rn(crm1, eps[i]));
}

__global__ void __launch_bounds__(1024) sampler_step_kernel(const float* __restrict__ x_t, const float* __restrict__ eps,
                                                            const float* __restrict__ z, float* __restrict__ x_prev,
                                                            float* __restrict__ x0_hat, float* __restrict__ s_out,
                                                            float cr, float crm1, float coef1, float coef2, float sigma,
                                                            int k_lo, float q_weight, int chw,
                                                            const float* __restrict__ coefs_dev) {
  __shared__ unsigned int hist[256];
  if (coefs_dev) {   // schedule values of this step from device memory: the launch can be replayed from a CUDA graph
    cr = coefs_dev[0];
    crm1 = coefs_dev[1];
    coef1 = coefs_dev[2];
    coef2 = coefs_dev[3];
    sigma = coefs_dev[4];
  }
  __shared__ unsigned int sh_prefix, sh_krem, sh_less, sh_bin_count, sh_min_gt;
  const size_t base = (size_t)blockIdx.x * chw;
  const int tid = threadIdx.x;

  unsigned int prefix = 0, mask = 0, k_rem = (unsigned int)k_lo, n_less = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < chw; i += 1024) {
      const unsigned int u = __float_as_uint(fabsf(x0hat_at(x_t, eps, base + i, cr, crm1)));
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int cum = 0, bin = 0;
      for (; bin < 256; ++bin) {
        if (cum + hist[bin] > k_rem) break;
        cum += hist[bin];
      }
      sh_prefix = prefix | (bin << shift);
      sh_krem = k_rem - cum;
      sh_less = n_less + cum;
      sh_bin_count = hist[bin];
    }
    __syncthreads();
    prefix = sh_prefix;
    k_rem = sh_krem;
    n_less = sh_less;
    mask |= 255u << shift;
    __syncthreads();
  }
  const unsigned int v_lo_bits = prefix;
  const unsigned int n_le = n_less + sh_bin_count;  // elements <= v_lo
  unsigned int v_hi_bits = v_lo_bits;
  if (n_le <= (unsigned int)k_lo + 1u && (k_lo + 1) < chw) {  // next order statistic is strictly larger
    if (tid == 0) sh_min_gt = 0xffffffffu;
    __syncthreads();
    unsigned int m = 0xffffffffu;
    for (int i = tid; i < chw; i += 1024) {
      const unsigned int u = __float_as_uint(fabsf(x0hat_at(x_t, eps, base + i, cr, crm1)));
      if (u > v_lo_bits) m = min(m, u);
    }
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) atomicMin(&sh_min_gt, m);
    __syncthreads();
    v_hi_bits = sh_min_gt;
  }
  const float v_lo = __uint_as_float(v_lo_bits), v_hi = __uint_as_float(v_hi_bits);
  // at::lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
  float s = (fabsf(q_weight) < 0.5f) ? __fadd_rn(v_lo, __fmul_rn(q_weight, __fsub_rn(v_hi, v_lo)))
                                     : __fsub_rn(v_hi, __fmul_rn(__fsub_rn(v_hi, v_lo), __fsub_rn(1.0f, q_weight)));
  s = fmaxf(s, 1.0f);
  if (tid == 0 && s_out) s_out[blockIdx.x] = s;
  for (int i = tid; i < chw; i += 1024) {
    const float xt = x_t[base + i];
    float x0h = x0hat_at(x_t, eps, base + i, cr, crm1);
    x0h = __fdiv_rn(fminf(fmaxf(x0h, -s), s), s);
    if (x0_hat) x0_hat[base + i] = x0h;
    float mean = __fadd_rn(__fmul_rn(coef1, x0h), __fmul_rn(coef2, xt));
    if (z) mean = __fadd_rn(mean, __fmul_rn(sigma, z[base + i]));
    x_prev[base + i] = mean;
  }
}

extern "C" int tedm_sampler_step(const float* x_t, const float* eps, const float* z, float* x_prev, float* x0_hat,
                                 float* s_out, float sqrt_recip_ac, float sqrt_recipm1_ac, float coef1, float coef2,
                                 float sigma, int k_lo, float q_weight, int batch, int chw, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x_t && eps && x_prev, "tedm_sampler_step: null pointer");
  TEDM_CHECK_ARG(batch > 0 && chw > 0 && k_lo >= 0 && k_lo < chw, "tedm_sampler_step: bad sizes batch=%d chw=%d k_lo=%d",
                 batch, chw, k_lo);
  sampler_step_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(x_t, eps, z, x_prev, x0_hat, s_out, sqrt_recip_ac,
                                                                sqrt_recipm1_ac, coef1, coef2, sigma, k_lo, q_weight, chw,
                                                                nullptr);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_sampler_step_dev(const float* x_t, const float* eps, const float* z, float* x_prev, float* x0_hat,
                                     float* s_out, const float* coefs, int k_lo, float q_weight, int batch, int chw,
                                     tedm_stream_t stream) {
  TEDM_CHECK_ARG(x_t && eps && z && x_prev && coefs, "tedm_sampler_step_dev: null pointer");
  TEDM_CHECK_ARG(batch > 0 && chw > 0 && k_lo >= 0 && k_lo < chw, "tedm_sampler_step_dev: bad sizes batch=%d chw=%d k_lo=%d",
                 batch, chw, k_lo);
  sampler_step_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(x_t, eps, z, x_prev, x0_hat, s_out, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f,
                                                                k_lo, q_weight, chw, coefs);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
