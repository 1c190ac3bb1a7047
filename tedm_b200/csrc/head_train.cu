// Training-mode TEDM / LEDM head (models/datasetDM_model.py:57-64; trainers/train_datasetDM.py:30-42):
//   Conv1x1(960[*S] -> 128) -> ReLU -> BatchNorm2d(128) -> Conv1x1(128 -> 32) -> ReLU -> BatchNorm2d(32) -> Conv1x1(32 -> 1)
// with BATCH statistics, and its backward into the head parameters (the UNet is frozen:
// datasetDM_model.py:67 is @torch.no_grad).  All GEMMs run on the tcgen05 kernels of conv_igemm.cu:
//   layer 1 per level at native resolution (commuted with the nearest upsample, as in head.cu),
//   layer 2 as a 1x1 conv on a1 = relu(z1) with BatchNorm-1 folded into its weights (W2 diag(A1), b2 + W2 C1),
//   d h1 = W2^T d z2 as a 1x1 conv, and every weight gradient as tedm_conv_igemm_wgrad.
// This file holds the memory-bound glue around them: gather-upsample-sum + ReLU + channel statistics,
// BatchNorm finalisation / weight folding, the logit, and the three reduction / apply passes of the backward.
// Per-pixel tensors: a1 bf16 [N][H][W][128]; z2 fp32 [N][H][W][64] (channels 32..63 are padding); dz2 bf16 [..][64].
#include "common.cuh"

#define HT_C1 128
#define HT_C2 32
#define HT_C2P 64   // layer-2 width padded to the conv kernel's 64-channel granularity

namespace {

struct GatherParams {
  const float* g[4];
  int shift[4];
  int n_levels, n_sum, H, W;
};

// thread = (pixel lane pl = tid / 16, channel chunk ck = tid % 16 of 8 channels); 128 threads = 8 pixels at a time
__global__ void __launch_bounds__(128) head_z1_kernel(const GatherParams p, const float* __restrict__ b1, bf16* __restrict__ a1,
                                                      float* __restrict__ sums /*[2][128]*/, long long npix) {
  __shared__ float red[2][8][HT_C1];
  const int pl = threadIdx.x >> 4, ck = threadIdx.x & 15;
  float bias[8], s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    bias[e] = b1[ck * 8 + e];
    s[e] = q[e] = 0.0f;
  }
  for (long long pix = (long long)blockIdx.x * 8 + pl; pix < npix; pix += (long long)gridDim.x * 8) {
    const int x = (int)(pix % p.W), y = (int)((pix / p.W) % p.H);
    const long long img = pix / ((long long)p.W * p.H);
    float z[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) z[e] = bias[e];
    for (int st = 0; st < p.n_sum; ++st)
      for (int l = 0; l < p.n_levels; ++l) {
        const int sh = p.shift[l];
        const int hl = p.H >> sh, wl = p.W >> sh;
        const float4* src = reinterpret_cast<const float4*>(
            p.g[l] + ((((size_t)img * p.n_sum + st) * hl + (y >> sh)) * wl + (x >> sh)) * HT_C1 + ck * 8);
        const float4 u0 = __ldg(src), u1 = __ldg(src + 1);
        z[0] += u0.x; z[1] += u0.y; z[2] += u0.z; z[3] += u0.w; z[4] += u1.x; z[5] += u1.y; z[6] += u1.z; z[7] += u1.w;
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      z[e] = fmaxf(z[e], 0.0f);
      // statistics of the value the next layer will actually read (bf16-rounded)
      z[e] = __bfloat162float(__float2bfloat16_rn(z[e]));
      s[e] += z[e];
      q[e] = fmaf(z[e], z[e], q[e]);
    }
    *reinterpret_cast<uint4*>(a1 + (size_t)pix * HT_C1 + ck * 8) = pack8(z);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][pl][ck * 8 + e] = s[e];
    red[1][pl][ck * 8 + e] = q[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * HT_C1; i += 128) {
    const int which = i / HT_C1, c = i % HT_C1;
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[which][r][c];
    atomicAdd(sums + i, t);
  }
}

// BatchNorm (training) finalisation: stats[0][c]=mean, [1]=rstd, [2]=A=gamma*rstd, [3]=Cc=beta-mean*A; running buffers updated
// like torch (momentum, unbiased variance).  One CTA.
__global__ void bn_finalize_kernel(const float* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ stats, int C) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = (double)sums[c] / count;
    double var = (double)sums[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = gamma[c] * rstd;
    stats[c] = (float)mean;
    stats[C + c] = rstd;
    stats[2 * C + c] = a;
    stats[3 * C + c] = beta[c] - (float)mean * a;
    if (running_mean) {
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// W2' = W2 diag(A1) (bf16 [64][128], rows >= 32 zero), b2' = b2 + W2 C1 (fp32 [64]), W2T = W2^T (bf16 [128][64], cols >= 32 zero)
__global__ void head_fold_w2_kernel(const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ stats1,
                                    bf16* __restrict__ w2f, float* __restrict__ b2f, bf16* __restrict__ w2t) {
  const float* A = stats1 + 2 * HT_C1;
  const float* Cc = stats1 + 3 * HT_C1;
  for (int i = threadIdx.x; i < HT_C2P * HT_C1; i += blockDim.x) {
    const int j = i / HT_C1, k = i % HT_C1;
    const float w = j < HT_C2 ? w2[j * HT_C1 + k] : 0.0f;
    w2f[i] = __float2bfloat16_rn(w * A[k]);
    w2t[k * HT_C2P + j] = __float2bfloat16_rn(w);
  }
  for (int j = threadIdx.x; j < HT_C2P; j += blockDim.x) {
    float acc = 0.0f;
    if (j < HT_C2) {
      acc = b2[j];
      for (int k = 0; k < HT_C1; ++k) acc = fmaf(w2[j * HT_C1 + k], Cc[k], acc);
    }
    b2f[j] = acc;
  }
}

// thread = (pixel lane pl = tid / 8, channel quad cq = tid % 8 of 4 channels) over z2's 32 valid channels; 256 threads = 32 pixels
__global__ void __launch_bounds__(256) head_z2_stats_kernel(const float* __restrict__ z2, float* __restrict__ sums /*[2][32]*/,
                                                            long long npix) {
  __shared__ float red[2][32][HT_C2];
  const int pl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (long long pix = (long long)blockIdx.x * 32 + pl; pix < npix; pix += (long long)gridDim.x * 32) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(z2 + (size_t)pix * HT_C2P + cq * 4));
    const float a[4] = {fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f), fmaxf(v.z, 0.0f), fmaxf(v.w, 0.0f)};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[e] += a[e];
      q[e] = fmaf(a[e], a[e], q[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    red[0][pl][cq * 4 + e] = s[e];
    red[1][pl][cq * 4 + e] = q[e];
  }
  __syncthreads();
  if (threadIdx.x < 2 * HT_C2) {
    const int which = threadIdx.x / HT_C2, c = threadIdx.x % HT_C2;
    float t = 0.0f;
    for (int r = 0; r < 32; ++r) t += red[which][r][c];
    atomicAdd(sums + threadIdx.x, t);
  }
}

// MODE 0: logits = w3 . (relu(z2) A2 + C2) + b3
// MODE 1: backward reductions  S[0][j] = sum dh2_j, S[1][j] = sum dh2_j * a2hat_j, S[2][j] = sum dlogit * h2_j, S[3][0] = sum dlogit
// MODE 2: dz2 (bf16, 64 channels, upper 32 zero) and S[4][j] = sum dz2_j
template <int MODE>
__global__ void __launch_bounds__(256) head_tail_train_kernel(const float* __restrict__ z2, const float* __restrict__ stats2,
                                                              const float* __restrict__ w3, const float* __restrict__ b3,
                                                              const float* __restrict__ dlogit, float* __restrict__ logits,
                                                              float* __restrict__ S, bf16* __restrict__ dz2, double count,
                                                              long long npix) {
  __shared__ float red[4][32][HT_C2];
  const int pl = threadIdx.x >> 3, cq = threadIdx.x & 7;
  float mean[4], rstd[4], A[4], Cc[4], w[4], acc[4][4];
  float m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cq * 4 + e;
    mean[e] = stats2[c];
    rstd[e] = stats2[HT_C2 + c];
    A[e] = stats2[2 * HT_C2 + c];
    Cc[e] = stats2[3 * HT_C2 + c];
    w[e] = w3[c];
    acc[0][e] = acc[1][e] = acc[2][e] = acc[3][e] = 0.0f;
    m1[e] = MODE == 2 ? (float)((double)S[c] / count) : 0.0f;
    m2[e] = MODE == 2 ? (float)((double)S[HT_C2 + c] / count) : 0.0f;
  }
  const float bias3 = MODE == 0 ? b3[0] : 0.0f;
  for (long long pix0 = (long long)blockIdx.x * 32; pix0 < npix; pix0 += (long long)gridDim.x * 32) {
    const long long pix = pix0 + pl;
    const bool ok = pix < npix;
    float4 v = make_float4(0, 0, 0, 0);
    if (ok) v = __ldg(reinterpret_cast<const float4*>(z2 + (size_t)pix * HT_C2P + cq * 4));
    const float zz[4] = {v.x, v.y, v.z, v.w};
    if (MODE == 0) {
      float part = 0.0f;
#pragma unroll
      for (int e = 0; e < 4; ++e) part = fmaf(w[e], fmaf(fmaxf(zz[e], 0.0f), A[e], Cc[e]), part);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (ok && cq == 0) logits[pix] = part + bias3;
    } else {
      const float dl = ok ? __ldg(dlogit + pix) : 0.0f;
      float out[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a2 = fmaxf(zz[e], 0.0f);
        const float hat = (a2 - mean[e]) * rstd[e];
        const float dh2 = dl * w[e];
        if (MODE == 1) {
          acc[0][e] += dh2;
          acc[1][e] = fmaf(dh2, hat, acc[1][e]);
          acc[2][e] = fmaf(dl, fmaf(a2, A[e], Cc[e]), acc[2][e]);
          if (cq == 0 && e == 0) acc[3][0] += dl;
        } else {
          const float da2 = A[e] * (dh2 - m1[e] - hat * m2[e]);
          out[e] = zz[e] > 0.0f ? da2 : 0.0f;
          acc[0][e] += out[e];
        }
      }
      if (MODE == 2 && ok) {
        *reinterpret_cast<uint2*>(dz2 + (size_t)pix * HT_C2P + cq * 4) = make_uint2(pack_bf16x2(out[0], out[1]), pack_bf16x2(out[2], out[3]));
        *reinterpret_cast<uint2*>(dz2 + (size_t)pix * HT_C2P + HT_C2 + cq * 4) = make_uint2(0u, 0u);
      }
    }
  }
  if (MODE == 0) return;
  constexpr int NQ = MODE == 1 ? 4 : 1;
#pragma unroll
  for (int qn = 0; qn < NQ; ++qn)
#pragma unroll
    for (int e = 0; e < 4; ++e) red[qn][pl][cq * 4 + e] = acc[qn][e];
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * HT_C2; i += 256) {
    const int qn = i / HT_C2, c = i % HT_C2;
    float t = 0.0f;
    for (int r = 0; r < 32; ++r) t += red[qn][r][c];
    if (MODE == 1) {
      if (qn < 3 || c == 0) atomicAdd(S + qn * HT_C2 + c, t);
    } else {
      atomicAdd(S + 4 * HT_C2 + c, t);
    }
  }
}

// backward through BatchNorm-1 / ReLU-1.  thread = (pixel lane pl = tid / 16, channel chunk ck of 8)
// MODE 0: T[0][k] = sum dh1_k, T[1][k] = sum dh1_k * a1hat_k
// MODE 1: dz1 = relu'(a1) A1 (dh1 - T0/N - a1hat T1/N); db1 += sum dz1; per level the 2^shift block sums of dz1 as bf16 maps
struct PoolParams {
  bf16* d[4];
  int shift[4];
  int n_levels, H, W;
};
template <int MODE>
__global__ void __launch_bounds__(128) head_bn1_bwd_kernel(const float* __restrict__ dh1, const bf16* __restrict__ a1,
                                                           const float* __restrict__ stats1, float* __restrict__ T,
                                                           float* __restrict__ db1, const PoolParams pp, double count,
                                                           long long n_blocks /* 8x8 pixel blocks */) {
  extern __shared__ float tile[];   // MODE 1: [64 px][128] dz1 of one 8x8 block
  __shared__ float red[2][8][HT_C1];
  const int pl = threadIdx.x >> 4, ck = threadIdx.x & 15;
  float mean[8], rstd[8], A[8], m1[8], m2[8], acc0[8], acc1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = ck * 8 + e;
    mean[e] = stats1[c];
    rstd[e] = stats1[HT_C1 + c];
    A[e] = stats1[2 * HT_C1 + c];
    m1[e] = MODE == 1 ? (float)((double)T[c] / count) : 0.0f;
    m2[e] = MODE == 1 ? (float)((double)T[HT_C1 + c] / count) : 0.0f;
    acc0[e] = acc1[e] = 0.0f;
  }
  const int bw = pp.W >> 3, bh = pp.H >> 3;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const int bx = (int)(blk % bw), by = (int)((blk / bw) % bh);
    const long long img = blk / ((long long)bw * bh);
    if (MODE == 1) __syncthreads();   // previous block's pooled reads are done
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
      const int lp = it * 8 + pl;                       // local pixel 0..63 : (ly, lx) = (lp / 8, lp % 8)
      const int y = by * 8 + (lp >> 3), x = bx * 8 + (lp & 7);
      const size_t pix = ((size_t)img * pp.H + y) * pp.W + x;
      float d[8], a[8];
      {
        const float4* dp = reinterpret_cast<const float4*>(dh1 + pix * HT_C1 + ck * 8);
        const float4 u0 = __ldg(dp), u1 = __ldg(dp + 1);
        d[0] = u0.x; d[1] = u0.y; d[2] = u0.z; d[3] = u0.w; d[4] = u1.x; d[5] = u1.y; d[6] = u1.z; d[7] = u1.w;
      }
      unpack8(ldg_stream(a1 + pix * HT_C1 + ck * 8), a);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float hat = (a[e] - mean[e]) * rstd[e];
        if (MODE == 0) {
          acc0[e] += d[e];
          acc1[e] = fmaf(d[e], hat, acc1[e]);
        } else {
          const float da1 = A[e] * (d[e] - m1[e] - hat * m2[e]);
          const float dz = a[e] > 0.0f ? da1 : 0.0f;
          acc0[e] += dz;
          tile[lp * HT_C1 + ck * 8 + e] = dz;
        }
      }
    }
    if (MODE == 1) {
      __syncthreads();
      for (int l = 0; l < pp.n_levels; ++l) {
        const int sh = pp.shift[l];
        const int side = 8 >> sh;                        // pooled pixels per block side
        const int hl = pp.H >> sh, wl = pp.W >> sh;
        for (int op = pl; op < side * side; op += 8) {
          const int oy = op / side, ox = op % side;
          float sum[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) sum[e] = 0.0f;
          for (int yy = 0; yy < (1 << sh); ++yy)
            for (int xx = 0; xx < (1 << sh); ++xx) {
              const float* src = tile + (((oy << sh) + yy) * 8 + (ox << sh) + xx) * HT_C1 + ck * 8;
              const float4 u0 = *reinterpret_cast<const float4*>(src), u1 = *reinterpret_cast<const float4*>(src + 4);
              sum[0] += u0.x; sum[1] += u0.y; sum[2] += u0.z; sum[3] += u0.w;
              sum[4] += u1.x; sum[5] += u1.y; sum[6] += u1.z; sum[7] += u1.w;
            }
          const size_t opix = ((size_t)img * hl + (by * side + oy)) * wl + bx * side + ox;
          *reinterpret_cast<uint4*>(pp.d[l] + opix * HT_C1 + ck * 8) = pack8(sum);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][pl][ck * 8 + e] = acc0[e];
    red[1][pl][ck * 8 + e] = acc1[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (MODE == 0 ? 2 : 1) * HT_C1; i += 128) {
    const int which = i / HT_C1, c = i % HT_C1;
    float t = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[which][r][c];
    atomicAdd((MODE == 0 ? T : db1) + i, t);
  }
}

// parameter gradients that are plain functions of the reduction buffers
//   dW2[j][k] += dW2f[j][k] * A1[k] + db2[j] * C1[k];  db2, dgamma/dbeta of both norms, dw3, db3
__global__ void head_param_grads_kernel(const float* __restrict__ dw2f /*[64][128]*/, const float* __restrict__ stats1,
                                        const float* __restrict__ S /*[5][32]*/, const float* __restrict__ T /*[2][128]*/,
                                        float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ dg1,
                                        float* __restrict__ dbt1, float* __restrict__ dg2, float* __restrict__ dbt2,
                                        float* __restrict__ dw3, float* __restrict__ db3) {
  const float* A1 = stats1 + 2 * HT_C1;
  const float* C1 = stats1 + 3 * HT_C1;
  for (int i = threadIdx.x; i < HT_C2 * HT_C1; i += blockDim.x) {
    const int j = i / HT_C1, k = i % HT_C1;
    dw2[i] += dw2f[j * HT_C1 + k] * A1[k] + S[4 * HT_C2 + j] * C1[k];
  }
  for (int k = threadIdx.x; k < HT_C1; k += blockDim.x) {
    dg1[k] += T[HT_C1 + k];
    dbt1[k] += T[k];
  }
  for (int j = threadIdx.x; j < HT_C2; j += blockDim.x) {
    db2[j] += S[4 * HT_C2 + j];
    dg2[j] += S[HT_C2 + j];
    dbt2[j] += S[j];
    dw3[j] += S[2 * HT_C2 + j];
  }
  if (threadIdx.x == 0) db3[0] += S[3 * HT_C2];
}

int grid_px(long long npix, int px_per_cta, int mult) {
  long long blocks = (npix + px_per_cta - 1) / px_per_cta;
  const long long cap = (long long)tedm_num_sms() * mult;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int tedm_head_train_z1(const tedm_head_args* a, void* a1, float* sums, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && a1 && sums && a->b1, "tedm_head_train_z1: null pointer");
  TEDM_CHECK_ARG(a->n_levels >= 1 && a->n_levels <= 4 && a->n_sum >= 1 && a->n_img > 0 && a->height > 0 && a->width > 0,
                 "tedm_head_train_z1: bad sizes");
  TEDM_UNSUPPORTED(a->c1 != HT_C1 || a->g_dtype != 1, "tedm_head_train_z1: 128-wide fp32 layer-1 maps only");
  GatherParams p{};
  for (int l = 0; l < a->n_levels; ++l) {
    TEDM_CHECK_ARG(a->g[l] != nullptr && a->shift[l] >= 0 && ((a->height >> a->shift[l]) << a->shift[l]) == a->height &&
                       ((a->width >> a->shift[l]) << a->shift[l]) == a->width,
                   "tedm_head_train_z1: level %d shift %d does not divide %dx%d", l, a->shift[l], a->height, a->width);
    p.g[l] = (const float*)a->g[l];
    p.shift[l] = a->shift[l];
  }
  p.n_levels = a->n_levels;
  p.n_sum = a->n_sum;
  p.H = a->height;
  p.W = a->width;
  const long long npix = (long long)a->n_img * a->height * a->width;
  head_z1_kernel<<<grid_px(npix, 8 * 16, 8), 128, 0, (cudaStream_t)stream>>>(p, a->b1, (bf16*)a1, sums, npix);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_bn_finalize(const float* sums, double count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, float* stats, int channels,
                                tedm_stream_t stream) {
  TEDM_CHECK_ARG(sums && gamma && beta && stats && channels > 0 && count >= 1.0, "tedm_bn_finalize: bad arguments");
  TEDM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "tedm_bn_finalize: running buffers come in pairs");
  bn_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(sums, count, gamma, beta, eps, momentum, running_mean, running_var,
                                                          stats, channels);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_head_fold_w2(const float* w2, const float* b2, const float* stats1, void* w2_folded, float* b2_folded,
                                 void* w2_t, tedm_stream_t stream) {
  TEDM_CHECK_ARG(w2 && b2 && stats1 && w2_folded && b2_folded && w2_t, "tedm_head_fold_w2: null pointer");
  head_fold_w2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(w2, b2, stats1, (bf16*)w2_folded, b2_folded, (bf16*)w2_t);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_head_z2_stats(const float* z2, float* sums, int64_t npix, tedm_stream_t stream) {
  TEDM_CHECK_ARG(z2 && sums && npix > 0, "tedm_head_z2_stats: bad arguments");
  head_z2_stats_kernel<<<grid_px(npix, 32 * 16, 8), 256, 0, (cudaStream_t)stream>>>(z2, sums, npix);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_head_train_tail(int mode, const float* z2, const float* stats2, const float* w3, const float* b3,
                                    const float* dlogit, float* logits, float* S, void* dz2, double count, int64_t npix,
                                    tedm_stream_t stream) {
  TEDM_CHECK_ARG(z2 && stats2 && w3 && npix > 0 && mode >= 0 && mode <= 2, "tedm_head_train_tail: bad arguments");
  const int grid = grid_px(npix, 32 * 16, 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) {
    TEDM_CHECK_ARG(b3 && logits, "tedm_head_train_tail: mode 0 needs b3 and logits");
    head_tail_train_kernel<0><<<grid, 256, 0, s>>>(z2, stats2, w3, b3, nullptr, logits, nullptr, nullptr, count, npix);
  } else if (mode == 1) {
    TEDM_CHECK_ARG(dlogit && S, "tedm_head_train_tail: mode 1 needs dlogit and S");
    head_tail_train_kernel<1><<<grid, 256, 0, s>>>(z2, stats2, w3, nullptr, dlogit, nullptr, S, nullptr, count, npix);
  } else {
    TEDM_CHECK_ARG(dlogit && S && dz2 && count >= 1.0, "tedm_head_train_tail: mode 2 needs dlogit, S and dz2");
    head_tail_train_kernel<2><<<grid, 256, 0, s>>>(z2, stats2, w3, nullptr, dlogit, nullptr, S, (bf16*)dz2, count, npix);
  }
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_head_bn1_bwd(int mode, const void* dh1, const void* a1, const float* stats1, float* T, float* db1,
                                 void* const* pooled, const int* shifts, int n_levels, int n_img, int height, int width,
                                 double count, tedm_stream_t stream) {
  TEDM_CHECK_ARG(dh1 && a1 && stats1 && T && (mode == 0 || mode == 1) && n_img > 0, "tedm_head_bn1_bwd: bad arguments");
  TEDM_UNSUPPORTED(height % 8 != 0 || width % 8 != 0, "tedm_head_bn1_bwd: image extent %dx%d must be a multiple of 8", height, width);
  PoolParams pp{};
  pp.H = height;
  pp.W = width;
  if (mode == 1) {
    TEDM_CHECK_ARG(db1 && pooled && shifts && n_levels >= 1 && n_levels <= 4 && count >= 1.0, "tedm_head_bn1_bwd: mode 1 arguments");
    pp.n_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) {
      TEDM_UNSUPPORTED(shifts[l] < 0 || shifts[l] > 3 || pooled[l] == nullptr, "tedm_head_bn1_bwd: level %d shift %d (0..3)", l, shifts[l]);
      pp.d[l] = (bf16*)pooled[l];
      pp.shift[l] = shifts[l];
    }
  }
  const long long n_blocks = (long long)n_img * (height / 8) * (width / 8);
  long long grid = n_blocks;
  const long long cap = (long long)tedm_num_sms() * 6;
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) head_bn1_bwd_kernel<0><<<(int)grid, 128, 0, s>>>((const float*)dh1, (const bf16*)a1, stats1, T, db1, pp, count, n_blocks);
  else head_bn1_bwd_kernel<1><<<(int)grid, 128, 64 * HT_C1 * sizeof(float), s>>>((const float*)dh1, (const bf16*)a1, stats1, T, db1, pp, count, n_blocks);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_head_param_grads(const float* dw2_folded, const float* stats1, const float* S, const float* T, float* dw2,
                                     float* db2, float* dgamma1, float* dbeta1, float* dgamma2, float* dbeta2, float* dw3,
                                     float* db3, tedm_stream_t stream) {
  TEDM_CHECK_ARG(dw2_folded && stats1 && S && T && dw2 && db2 && dgamma1 && dbeta1 && dgamma2 && dbeta2 && dw3 && db3,
                 "tedm_head_param_grads: null pointer");
  head_param_grads_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dw2_folded, stats1, S, T, dw2, db2, dgamma1, dbeta1, dgamma2,
                                                               dbeta2, dw3, db3);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
