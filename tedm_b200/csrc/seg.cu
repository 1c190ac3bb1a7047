// Supervised-segmentation edges of the path: BCE-with-logits loss (+ gradient), dice / precision / recall counting,
// uint8 -> fp32 input transport.  All HBM-bound single-pass kernels over NCHW fp32 / uint8 tensors.
//   loss     trainers/train_baseline.py:44-45   reduce(bce_with_logits(pred, y, 'none'), 'b c h w -> b c', 'mean').mean()
//   labels   trainers/train_baseline.py:30-31   repeat(y, 'b c h w -> (b step) c h w')  (never materialised: row / repeat)
//   metrics  trainers/train_baseline.py:146-161 dice, precision, recall per (b, c) row
//   inputs   dataloaders/JSRT.py:62-82, dataloaders/CXR14.py:67-70  ToTensor (u8 / 255), (label > .5) summed over lungs
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxChunks = 64;  // partial sums per row

__device__ __forceinline__ float block_sum(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.0f;
  if (w == 0) {
    r = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// max(x,0) - x*y + log1p(exp(-|x|)): the stable form of -(y log s(x) + (1-y) log(1-s(x)))
__device__ __forceinline__ float bce_term(float x, float y) { return fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// grid (chunks, rows).  partial[row][chunk] = sum of the BCE terms of the chunk; grad written in the same pass.
__global__ void __launch_bounds__(kThreads) bce_partial_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                               float* __restrict__ partial, float* __restrict__ grad,
                                                               long long row_len, int target_repeat, float gscale) {
  __shared__ float sh[kThreads / 32];
  const long long row = blockIdx.y;
  const int chunks = gridDim.x;
  const long long per = ((row_len + chunks - 1) / chunks + 3) / 4 * 4;
  const long long lo = (long long)blockIdx.x * per;
  const long long hi = lo + per < row_len ? lo + per : row_len;
  const float* x = logits + row * row_len;
  const float* y = target + (row / target_repeat) * row_len;
  float* g = grad ? grad + row * row_len : nullptr;
  float acc = 0.0f;
  const bool vec = (row_len % 4) == 0;
  if (vec) {
    for (long long i = lo + 4LL * threadIdx.x; i + 3 < hi; i += 4LL * kThreads) {
      const float4 xv = *reinterpret_cast<const float4*>(x + i);
      const float4 yv = *reinterpret_cast<const float4*>(y + i);
      acc += (bce_term(xv.x, yv.x) + bce_term(xv.y, yv.y)) + (bce_term(xv.z, yv.z) + bce_term(xv.w, yv.w));
      if (g) {
        float4 gv;
        gv.x = (sigmoid_f(xv.x) - yv.x) * gscale;
        gv.y = (sigmoid_f(xv.y) - yv.y) * gscale;
        gv.z = (sigmoid_f(xv.z) - yv.z) * gscale;
        gv.w = (sigmoid_f(xv.w) - yv.w) * gscale;
        *reinterpret_cast<float4*>(g + i) = gv;
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
      acc += bce_term(x[i], y[i]);
      if (g) g[i] = (sigmoid_f(x[i]) - y[i]) * gscale;
    }
  }
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[row * chunks + blockIdx.x] = s;
}

// one CTA: row means (fixed summation order -> run-to-run identical), then the mean over rows
__global__ void __launch_bounds__(kThreads) bce_finalize_kernel(const float* __restrict__ partial, float* __restrict__ row_mean,
                                                                float* __restrict__ loss, int n_rows, int chunks,
                                                                float inv_row_len) {
  __shared__ float sh[kThreads / 32];
  float acc = 0.0f;
  for (int r = threadIdx.x; r < n_rows; r += kThreads) {
    float s = 0.0f;
    for (int c = 0; c < chunks; ++c) s += partial[(long long)r * chunks + c];
    s *= inv_row_len;
    if (row_mean) row_mean[r] = s;
    acc += s;
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) loss[0] = tot / (float)n_rows;
}

// One CTA per (b, c) row.  pred is a uint8 mask (nonzero = foreground) or fp32 logits (sigmoid(x) > .5).
// out[row] = {dice, precision, recall, TP, FP, FN, sum(pred), sum(target)}; 0/0 = NaN as in the reference.
template <bool LOGITS>
__global__ void __launch_bounds__(kThreads) seg_metrics_kernel(const void* __restrict__ pred_, const float* __restrict__ target,
                                                               float* __restrict__ out, long long row_len, int target_repeat) {
  __shared__ float sh[kThreads / 32];
  const long long row = blockIdx.x;
  const float* y = target + (row / target_repeat) * row_len;
  float tp = 0.0f, fp = 0.0f, fn = 0.0f, sp = 0.0f, sy = 0.0f;
  for (long long i = threadIdx.x; i < row_len; i += kThreads) {
    bool p;
    if (LOGITS) p = sigmoid_f(reinterpret_cast<const float*>(pred_)[row * row_len + i]) > 0.5f;
    else p = reinterpret_cast<const uint8_t*>(pred_)[row * row_len + i] != 0;
    const float yv = y[i];
    const bool pos = yv != 0.0f;             // logical_and(x, .) : any nonzero label is True
    const bool neg = (1.0f - yv) != 0.0f;    // logical_and(1 - x, .)
    tp += (pos && p) ? 1.0f : 0.0f;
    fp += (neg && p) ? 1.0f : 0.0f;
    fn += (pos && !p) ? 1.0f : 0.0f;
    sp += p ? 1.0f : 0.0f;
    sy += yv;
  }
  tp = block_sum(tp, sh);
  fp = block_sum(fp, sh);
  fn = block_sum(fn, sh);
  sp = block_sum(sp, sh);
  sy = block_sum(sy, sh);
  if (threadIdx.x == 0) {
    float* o = out + row * 8;
    o[0] = __fdiv_rn(2.0f * tp, sp + sy);
    o[1] = __fdiv_rn(tp, tp + fp);
    o[2] = __fdiv_rn(tp, tp + fn);
    o[3] = tp;
    o[4] = fp;
    o[5] = fn;
    o[6] = sp;
    o[7] = sy;
  }
}

// ToTensor: u8 / 255 (true division, bit-exact with torch's .div(255)); 16 pixels per thread-iteration
__global__ void __launch_bounds__(kThreads) u8_to_unit_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * kThreads * 16;
  for (long long i = ((long long)blockIdx.x * kThreads + threadIdx.x) * 16; i < n; i += stride) {
    if (i + 16 <= n) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 f;
        f.x = __fdiv_rn((float)(w[k] & 255u), 255.0f);
        f.y = __fdiv_rn((float)((w[k] >> 8) & 255u), 255.0f);
        f.z = __fdiv_rn((float)((w[k] >> 16) & 255u), 255.0f);
        f.w = __fdiv_rn((float)(w[k] >> 24), 255.0f);
        *reinterpret_cast<float4*>(dst + i + 4 * k) = f;
      }
    } else {
      for (long long j = i; j < n; ++j) dst[j] = __fdiv_rn((float)src[j], 255.0f);
    }
  }
}

// label = min(sum_k [u8_k / 255 > .5], 1): u8/255 > .5  <=>  u8 >= 128
__global__ void __launch_bounds__(kThreads) u8_masks_to_label_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                                     long long n_img, long long hw, int n_masks) {
  const long long total = n_img * hw;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const long long b = i / hw, p = i - b * hw;
    int s = 0;
    for (int k = 0; k < n_masks; ++k) s += src[(b * n_masks + k) * hw + p] >= 128 ? 1 : 0;
    dst[i] = s > 1 ? 1.0f : (float)s;
  }
}

}  // namespace

extern "C" int tedm_bce_logits(const float* logits, const float* target, float* row_mean, float* loss, float* grad,
                               float* workspace, long long n_rows, long long row_len, int target_repeat, float grad_scale,
                               tedm_stream_t stream) {
  TEDM_CHECK_ARG(logits && target && loss && workspace, "tedm_bce_logits: null pointer");
  TEDM_CHECK_ARG(n_rows > 0 && row_len > 0 && target_repeat > 0 && n_rows % target_repeat == 0 && n_rows < (1 << 30),
                 "tedm_bce_logits: bad sizes n_rows=%lld row_len=%lld repeat=%d", n_rows, row_len, target_repeat);
  int chunks = (int)((row_len + 4095) / 4096);
  if (chunks > kMaxChunks) chunks = kMaxChunks;
  TEDM_CHECK_ARG(n_rows <= 65535, "tedm_bce_logits: more than 65535 rows");
  const float gscale = grad_scale / ((float)row_len * (float)n_rows);
  bce_partial_kernel<<<dim3(chunks, (unsigned)n_rows), kThreads, 0, (cudaStream_t)stream>>>(logits, target, workspace, grad,
                                                                                             row_len, target_repeat, gscale);
  TEDM_LAUNCH_CHECK();
  bce_finalize_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(workspace, row_mean, loss, (int)n_rows, chunks,
                                                                1.0f / (float)row_len);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_bce_workspace_floats(long long n_rows) { return (int)(n_rows * kMaxChunks); }

extern "C" int tedm_seg_metrics(const void* pred, int pred_is_logits, const float* target, float* out, long long n_rows,
                                long long row_len, int target_repeat, tedm_stream_t stream) {
  TEDM_CHECK_ARG(pred && target && out, "tedm_seg_metrics: null pointer");
  TEDM_CHECK_ARG(n_rows > 0 && row_len > 0 && row_len < (1 << 24) && target_repeat > 0 && n_rows % target_repeat == 0,
                 "tedm_seg_metrics: bad sizes n_rows=%lld row_len=%lld (counts are exact in fp32 below 2^24) repeat=%d", n_rows,
                 row_len, target_repeat);
  if (pred_is_logits)
    seg_metrics_kernel<true><<<(unsigned)n_rows, kThreads, 0, (cudaStream_t)stream>>>(pred, target, out, row_len, target_repeat);
  else
    seg_metrics_kernel<false><<<(unsigned)n_rows, kThreads, 0, (cudaStream_t)stream>>>(pred, target, out, row_len, target_repeat);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_u8_to_unit(const uint8_t* src, float* dst, long long n, tedm_stream_t stream) {
  TEDM_CHECK_ARG(src && dst && n > 0, "tedm_u8_to_unit: bad arguments");
  TEDM_CHECK_ARG(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "tedm_u8_to_unit: pointers must be 16-byte aligned");
  int grid = ceil_div(n, (long long)kThreads * 16);
  const int cap = resident_ctas(u8_to_unit_kernel, kThreads, 0);
  if (grid > cap) grid = cap;
  u8_to_unit_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(src, dst, n);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_u8_masks_to_label(const uint8_t* src, float* dst, long long n_img, long long hw, int n_masks,
                                      tedm_stream_t stream) {
  TEDM_CHECK_ARG(src && dst && n_img > 0 && hw > 0 && n_masks > 0, "tedm_u8_masks_to_label: bad arguments");
  int grid = ceil_div(n_img * hw, kThreads);
  const int cap = resident_ctas(u8_masks_to_label_kernel, kThreads, 0);
  if (grid > cap) grid = cap;
  u8_masks_to_label_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(src, dst, n_img, hw, n_masks);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
