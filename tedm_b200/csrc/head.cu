// TEDM / LEDM per-pixel MLP head, commuted form.
//
// Reference (models/datasetDM_model.py:57-64,80-88; trainers/train_datasetDM.py:30-42): nearest-
// upsample the four decoder maps to full resolution, concatenate to 960 (x S) channels, then
// Conv1x1(->128) -> ReLU -> BN -> Conv1x1(->32) -> ReLU -> BN -> Conv1x1(->1).  A 1x1 conv commutes
// with nearest upsampling, so layer 1 runs per level at native resolution on the tcgen05 GEMM
// (tedm_conv_igemm_fwd, mode 0) and this kernel finishes the job per output pixel:
//   z1 = b1 + sum_{s < n_sum} sum_l g_l[img*n_sum + s][y >> sh_l][x >> sh_l]      (128 values)
//   h1 = bn1(relu(z1)) ; z2 = W2 h1 + b2 ; h2 = bn2(relu(z2)) ; logit = w3 . h2 + b3
// with eval-mode BatchNorm folded to an affine.  The 960-channel full-resolution tensor of the
// reference (62.9 MB per image and step) is never materialised.
#include "common.cuh"

#define HEAD_C1 128
#define HEAD_C2 32

struct HeadParams {
  const void* g[4];
  int shift[4];
  int n_levels, n_sum, n_img, H, W;
  const float *b1, *a1, *c1, *w2, *b2, *a2, *c2, *w3;
  float b3;
  float* logits;
};

template <bool G_F32>
__global__ void __launch_bounds__(128) head_tail_kernel(const HeadParams p) {
  __shared__ __align__(16) float sw2[HEAD_C1][HEAD_C2];  // transposed: [k][j]
  __shared__ float sb1[HEAD_C1], sa1[HEAD_C1], sc1[HEAD_C1];
  for (int i = threadIdx.x; i < HEAD_C1 * HEAD_C2; i += blockDim.x) {
    const int j = i / HEAD_C1, k = i % HEAD_C1;  // w2 is [c2][c1]
    sw2[k][j] = p.w2[i];
  }
  for (int i = threadIdx.x; i < HEAD_C1; i += blockDim.x) {
    sb1[i] = p.b1[i];
    sa1[i] = p.a1[i];
    sc1[i] = p.c1[i];
  }
  __syncthreads();
  const long long npix = (long long)p.n_img * p.H * p.W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % p.W), y = (int)((pix / p.W) % p.H);
    const long long img = pix / ((long long)p.W * p.H);
    float acc[HEAD_C2];
#pragma unroll
    for (int j = 0; j < HEAD_C2; ++j) acc[j] = p.b2[j];
    for (int k0 = 0; k0 < HEAD_C1; k0 += 8) {
      float z[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = sb1[k0 + e];
      for (int s = 0; s < p.n_sum; ++s) {
        for (int l = 0; l < p.n_levels; ++l) {
          const int sh = p.shift[l];
          const int hl = p.H >> sh, wl = p.W >> sh;
          const size_t off = ((((size_t)img * p.n_sum + s) * hl + (y >> sh)) * wl + (x >> sh)) * HEAD_C1 + k0;
          float f[8];
          if (G_F32) {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.g[l]) + off);
            const float4 u0 = __ldg(src), u1 = __ldg(src + 1);
            f[0] = u0.x; f[1] = u0.y; f[2] = u0.z; f[3] = u0.w; f[4] = u1.x; f[5] = u1.y; f[6] = u1.z; f[7] = u1.w;
          } else {
            unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.g[l]) + off)), f);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) z[e] += f[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float hval = fmaf(fmaxf(z[e], 0.0f), sa1[k0 + e], sc1[k0 + e]);
        const float4* wr = reinterpret_cast<const float4*>(&sw2[k0 + e][0]);
#pragma unroll
        for (int j4 = 0; j4 < HEAD_C2 / 4; ++j4) {
          const float4 w = wr[j4];
          acc[j4 * 4] = fmaf(w.x, hval, acc[j4 * 4]);
          acc[j4 * 4 + 1] = fmaf(w.y, hval, acc[j4 * 4 + 1]);
          acc[j4 * 4 + 2] = fmaf(w.z, hval, acc[j4 * 4 + 2]);
          acc[j4 * 4 + 3] = fmaf(w.w, hval, acc[j4 * 4 + 3]);
        }
      }
    }
    float logit = p.b3;
#pragma unroll
    for (int j = 0; j < HEAD_C2; ++j) logit = fmaf(p.w3[j], fmaf(fmaxf(acc[j], 0.0f), p.a2[j], p.c2[j]), logit);
    p.logits[pix] = logit;
  }
}

extern "C" int tedm_head_infer(const tedm_head_args* a, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && a->logits && a->b1 && a->bn1_a && a->bn1_c && a->w2 && a->b2 && a->bn2_a && a->bn2_c && a->w3,
                 "tedm_head_infer: null pointer");
  TEDM_CHECK_ARG(a->n_levels >= 1 && a->n_levels <= 4 && a->n_sum >= 1 && a->n_img > 0 && a->height > 0 && a->width > 0,
                 "tedm_head_infer: bad sizes");
  TEDM_UNSUPPORTED(a->c1 != HEAD_C1 || a->c2 != HEAD_C2, "tedm_head_infer: head widths %d/%d (only 128/32)", a->c1, a->c2);
  HeadParams p{};
  for (int l = 0; l < a->n_levels; ++l) {
    TEDM_CHECK_ARG(a->g[l] != nullptr && a->shift[l] >= 0 && (a->height >> a->shift[l]) > 0 &&
                       ((a->height >> a->shift[l]) << a->shift[l]) == a->height &&
                       ((a->width >> a->shift[l]) << a->shift[l]) == a->width,
                   "tedm_head_infer: level %d shift %d does not divide %dx%d", l, a->shift[l], a->height, a->width);
    p.g[l] = a->g[l];
    p.shift[l] = a->shift[l];
  }
  p.n_levels = a->n_levels;
  p.n_sum = a->n_sum;
  p.n_img = a->n_img;
  p.H = a->height;
  p.W = a->width;
  p.b1 = a->b1; p.a1 = a->bn1_a; p.c1 = a->bn1_c; p.w2 = a->w2; p.b2 = a->b2; p.a2 = a->bn2_a; p.c2 = a->bn2_c; p.w3 = a->w3;
  p.b3 = a->b3;
  p.logits = a->logits;
  const long long npix = (long long)a->n_img * a->height * a->width;
  long long blocks = (npix + 127) / 128;
  const long long cap = a->g_dtype == 1 ? resident_ctas(head_tail_kernel<true>, 128, 0) : resident_ctas(head_tail_kernel<false>, 128, 0);
  if (blocks > cap) blocks = cap;
  TEDM_CHECK_ARG(a->g_dtype == 0 || a->g_dtype == 1, "tedm_head_infer: g_dtype=%d", a->g_dtype);
  if (a->g_dtype == 1) head_tail_kernel<true><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(p);
  else head_tail_kernel<false><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// prob[b] = mean_s sigmoid(logits[b*S + s]) ; mask = prob > 0.5
// (auxiliary/postprocessing/testing_shared_weights.py:113,120,133-138; app.py:79)
__global__ void __launch_bounds__(256) ensemble_kernel(const float* __restrict__ logits, float* __restrict__ prob,
                                                       uint8_t* __restrict__ mask, int n_steps, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw;
    const int pi = (int)(i % hw);
    float acc = 0.0f;
    for (int s = 0; s < n_steps; ++s) {
      const float v = logits[((size_t)b * n_steps + s) * hw + pi];
      acc += 1.0f / (1.0f + expf(-v));
    }
    const float pr = acc / (float)n_steps;
    if (prob) prob[i] = pr;
    if (mask) mask[i] = pr > 0.5f ? 1 : 0;
  }
}

extern "C" int tedm_ensemble_mask(const float* logits, float* prob, uint8_t* mask, int batch, int n_steps, int hw,
                                  tedm_stream_t stream) {
  TEDM_CHECK_ARG(logits && (prob || mask) && batch > 0 && n_steps > 0 && hw > 0, "tedm_ensemble_mask: bad arguments");
  const long long total = (long long)batch * hw;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)tedm_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  ensemble_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, prob, mask, n_steps, hw, total);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
