// fp32 precision mode (BASELINE north star: "1e-4 in fp32 mode"; the reference's default is fp32,
// config.py:15 `mixed_precision` False).  Inference only.
//
// Activations stay fp32 NHWC between kernels.  The convolutions still run on the tcgen05 implicit-GEMM
// kernel: every operand is split into a bf16 (hi, lo) pair, hi = bf16(x), lo = bf16(x - hi), and the
// product is accumulated in fp32 as hi*hi + lo*hi + hi*lo (the dropped lo*lo term is 2^-18 relative) --
// a convolution over the sources [x_hi, x_lo, x_hi] against the weights [w_hi, w_hi, w_lo]
// (tedm_conv_args.extra_src).  Everything around the convolutions is restated here in plain fp32 CUDA
// with exact expf / division (no approximate intrinsics): these kernels are written for precision and
// clarity, the bf16 path is the fast one.
#include "common.cuh"

namespace {

int grid_cap(long long items, int per_block, int mult) {
  long long blocks = (items + per_block - 1) / per_block;
  const long long cap = (long long)tedm_num_sms() * mult;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ---- fp32 -> (hi, lo) bf16 ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) f32_split_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                        long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const float f[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
      l[j] = f[j] - h[j];
    }
    hi[i] = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
    lo[i] = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
  }
}

// ---- stem: 7x7 pad-3 conv, fp32 NCHW in -> fp32 NHWC out (models/unet_model.py:267,334) -----------
// thread = (pixel, 8 output channels); weights transposed into shared memory as [tap][cout]
__global__ void __launch_bounds__(256) f32_stem_kernel(const float* __restrict__ x, const float* __restrict__ weight,
                                                       const float* __restrict__ bias, float* __restrict__ out, int batch,
                                                       int cin, int H, int W, int cout) {
  extern __shared__ float wsm[];
  const int ntap = cin * 49;
  for (int i = threadIdx.x; i < ntap * cout; i += blockDim.x) {
    const int co = i / ntap, tp = i % ntap;
    wsm[tp * cout + co] = weight[i];
  }
  __syncthreads();
  const int chunks = cout >> 3;
  const long long total = (long long)batch * H * W * chunks;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % chunks);
    long long rest = idx / chunks;
    const int px = (int)(rest % W);
    rest /= W;
    const int py = (int)(rest % H), b = (int)(rest / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int ci = 0; ci < cin; ++ci) {
      const float* xp = x + ((size_t)b * cin + ci) * H * W;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = py + ky - 3;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = px + kx - 3;
          if (xx < 0 || xx >= W) continue;
          const float v = __ldg(xp + (size_t)yy * W + xx);
          const float* wp = wsm + (ci * 49 + ky * 7 + kx) * cout + ch * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
        }
      }
    }
    float* op = out + (((size_t)b * H + py) * W + px) * cout + ch * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) op[j] = acc[j] + (bias ? bias[ch * 8 + j] : 0.0f);
  }
}

// ---- GroupNorm finalise + affine + (scale+1)/shift + SiLU (+ residual), fp32 (unet_model.py:126-135,175) ----
// Statistics come from the conv epilogue's per-(image, part, group) fp32 (sum, sum of squares) of its fp32
// accumulators, folded here in double.
#define F32_GN_MAX_C 1024
__global__ void __launch_bounds__(256) f32_gn_silu_kernel(const float* __restrict__ x, const float* __restrict__ partial, int parts,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ ss, int ss_stride, int ss_offset,
                                                          const float* __restrict__ residual, float* __restrict__ out, int hw,
                                                          int C, int groups, float eps) {
  __shared__ float sA[F32_GN_MAX_C], sB[F32_GN_MAX_C];
  __shared__ double s_mean[32], s_rstd[32];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpg = C / groups;
  for (int g = warp; g < groups; g += 8) {
    double s = 0.0, q = 0.0;
    for (int p = lane; p < parts; p += 32) {
      const float2 v = *reinterpret_cast<const float2*>(partial + (((size_t)b * parts + p) * groups + g) * 2);
      s += (double)v.x;
      q += (double)v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double n = (double)hw * (double)cpg;
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = mean;
      s_rstd[g] = 1.0 / sqrt(var + (double)eps);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    double a = s_rstd[g] * (double)gamma[c];
    double bb = (double)beta[c] - s_mean[g] * a;
    if (ss) {
      const double sc = (double)ss[(size_t)b * ss_stride + ss_offset + c] + 1.0;
      const double sh = (double)ss[(size_t)b * ss_stride + ss_offset + C + c];
      a *= sc;
      bb = bb * sc + sh;
    }
    sA[c] = (float)a;
    sB[c] = (float)bb;
  }
  __syncthreads();
  const long long n = (long long)hw * C;
  const size_t img = (size_t)b * hw * C;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    const int c = (int)(i % C);
    const float4 v = *reinterpret_cast<const float4*>(x + img + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = fmaf(f[j], sA[c + j], sB[c + j]);
      o[j] = z / (1.0f + expf(-z));
    }
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + img + i);
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    *reinterpret_cast<float4*>(out + img + i) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---- channel LayerNorm per pixel (gain only, biased variance) + optional residual (unet_model.py:52-61) ----
// one warp per pixel, two passes over registers (mean, then the centred sum of squares)
template <int NV>   // C = 32 * NV
__global__ void __launch_bounds__(256) f32_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                            const float* __restrict__ residual, float* __restrict__ out,
                                                            long long npix, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int C = 32 * NV;
  for (long long p = warp0; p < npix; p += nwarps) {
    float v[NV];
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j] = x[p * C + j * 32 + lane];
      s += v[j];
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j] -= mean;
      q = fmaf(v[j], v[j], q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float o = v[j] * rstd * g[j * 32 + lane];
      if (residual) o += residual[p * C + j * 32 + lane];
      out[p * C + j * 32 + lane] = o;
    }
  }
}

// ---- LinearAttention core, fp32 (unet_model.py:196-210) ---------------------------------------------
// qkv: [B][n][3*H*32] fp32, channel = part*H*32 + h*32 + d  (part 0 = q, 1 = k, 2 = v).
//   k = softmax over n;  v /= n;  ctx[d][e] = sum_n k[d,n] v[e,n];  q = softmax over d * scale;
//   out[n][h*32+e] = sum_d ctx[d][e] q[d,n]
// pass 1: per (chunk of 128 pixels, head, image): chunk maximum m[d], s[d] = sum exp(k - m), c[d][e] = sum exp(k - m) v
constexpr int LA_CHUNK = 128;
constexpr int LA_PART = 32 + 32 + 32 * 32;   // floats per partial: m, s, c
__global__ void __launch_bounds__(256) f32_linattn_ctx_kernel(const float* __restrict__ qkv, float* __restrict__ part, int n,
                                                              int heads) {
  __shared__ float ks[LA_CHUNK][33], vs[LA_CHUNK][32];
  __shared__ float red[8][32], sm[32];
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n0 = chunk * LA_CHUNK, cnt = min(LA_CHUNK, n - n0);
  const int C3 = 3 * heads * 32;
  const float* base = qkv + ((size_t)b * n + n0) * C3 + h * 32;
  for (int i = threadIdx.x; i < LA_CHUNK * 32; i += 256) {
    const int px = i >> 5, d = i & 31;
    ks[px][d] = px < cnt ? base[(size_t)px * C3 + heads * 32 + d] : -INFINITY;
    vs[px][d] = px < cnt ? base[(size_t)px * C3 + 2 * heads * 32 + d] : 0.0f;
  }
  __syncthreads();
  const int d = threadIdx.x & 31, pt = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int px = pt; px < LA_CHUNK; px += 8) m = fmaxf(m, ks[px][d]);
  red[pt][d] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float mm = red[0][d];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i][d]);
    sm[d] = mm;
  }
  __syncthreads();
  float s = 0.0f;
  for (int px = pt; px < LA_CHUNK; px += 8) {
    const float e = px < cnt ? expf(ks[px][d] - sm[d]) : 0.0f;
    ks[px][d] = e;
    s += e;
  }
  red[pt][d] = s;
  __syncthreads();
  float* dst = part + (((size_t)b * heads + h) * gridDim.x + chunk) * LA_PART;
  if (threadIdx.x < 32) {
    float ss = 0.0f;
    for (int i = 0; i < 8; ++i) ss += red[i][d];
    dst[d] = sm[d];
    dst[32 + d] = ss;
  }
  const int dd = threadIdx.x >> 3, e0 = (threadIdx.x & 7) * 4;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int px = 0; px < cnt; ++px) {
    const float p = ks[px][dd];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fmaf(p, vs[px][e0 + j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) dst[64 + dd * 32 + e0 + j] = acc[j];
}

// pass 2: per (head, image): ctx[d][e] = sum_c c_c[d][e] exp(m_c[d] - M[d]) / (S[d] * n)
__global__ void __launch_bounds__(1024) f32_linattn_combine_kernel(const float* __restrict__ part, float* __restrict__ ctx,
                                                                   int chunks, int n) {
  const int d = threadIdx.x >> 5, e = threadIdx.x & 31;
  const float* src = part + (size_t)blockIdx.x * chunks * LA_PART;
  float M = -INFINITY;
  for (int c = 0; c < chunks; ++c) M = fmaxf(M, src[(size_t)c * LA_PART + d]);
  double S = 0.0, acc = 0.0;
  for (int c = 0; c < chunks; ++c) {
    const float w = expf(src[(size_t)c * LA_PART + d] - M);
    S += (double)src[(size_t)c * LA_PART + 32 + d] * w;
    acc += (double)src[(size_t)c * LA_PART + 64 + d * 32 + e] * w;
  }
  ctx[(size_t)blockIdx.x * 1024 + d * 32 + e] = (float)(acc / (S * (double)n));
}

// pass 3: thread = (pixel, head)
__global__ void __launch_bounds__(128) f32_linattn_out_kernel(const float* __restrict__ qkv, const float* __restrict__ ctx,
                                                              float* __restrict__ out, int n, int heads, float scale) {
  extern __shared__ float cs[];   // [heads][32][32]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < heads * 1024; i += blockDim.x) cs[i] = ctx[(size_t)b * heads * 1024 + i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * heads) return;
  const int h = (int)(idx % heads);
  const long long px = idx / heads;
  const int C3 = 3 * heads * 32;
  const float* qp = qkv + ((size_t)b * n + px) * C3 + h * 32;
  float q[32];
  float m = -INFINITY;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    q[d] = qp[d];
    m = fmaxf(m, q[d]);
  }
  float s = 0.0f;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    q[d] = expf(q[d] - m);
    s += q[d];
  }
  const float inv = scale / s;
  float* op = out + ((size_t)b * n + px) * (heads * 32) + h * 32;
  const float* c = cs + h * 1024;
  for (int e = 0; e < 32; ++e) {
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < 32; ++d) acc = fmaf(c[d * 32 + e], q[d], acc);
    op[e] = acc * inv;
  }
}

// ---- mid Attention core, fp32 (unet_model.py:229-241): q, k L2-normalised over the TOKEN axis, sim * scale, softmax ----
// rnorm[b][c] = 1 / max(||x[b, :, c]||_2, 1e-12) for the 2*H*32 q and k channels
__global__ void __launch_bounds__(256) f32_attn_norm_kernel(const float* __restrict__ qkv, float* __restrict__ rnorm, int n, int heads) {
  __shared__ float red[8];
  const int c = blockIdx.x, b = blockIdx.y, C3 = 3 * heads * 32;
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = qkv[((size_t)b * n + i) * C3 + c];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    rnorm[(size_t)b * 2 * heads * 32 + c] = 1.0f / fmaxf(sqrtf(t), 1e-12f);
  }
}

// thread = one query token; keys / values stream through shared memory in tiles of 32 with an online softmax
__global__ void __launch_bounds__(128) f32_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ rnorm,
                                                       float* __restrict__ out, int n, int heads, float scale) {
  __shared__ float kt[32][32], vt[32][32];
  const int h = blockIdx.y, b = blockIdx.z, C3 = 3 * heads * 32, HD = heads * 32;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool live = i < n;
  const float* rq = rnorm + (size_t)b * 2 * HD + h * 32;
  const float* rk = rq + HD;
  float q[32], acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    q[d] = live ? qkv[((size_t)b * n + i) * C3 + h * 32 + d] * rq[d] : 0.0f;
    acc[d] = 0.0f;
  }
  float m = -INFINITY, l = 0.0f;
  for (int j0 = 0; j0 < n; j0 += 32) {
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * 32; t += 128) {
      const int j = t >> 5, d = t & 31;
      const bool ok = j0 + j < n;
      const float* src = qkv + ((size_t)b * n + j0 + j) * C3 + h * 32 + d;
      kt[j][d] = ok ? src[HD] * rk[d] : 0.0f;
      vt[j][d] = ok ? src[2 * HD] : 0.0f;
    }
    __syncthreads();
    float sc[32];
    float tm = -INFINITY;
    const int cnt = min(32, n - j0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int d = 0; d < 32; ++d) s = fmaf(q[d], kt[j][d], s);
      sc[j] = j < cnt ? s * scale : -INFINITY;
      tm = fmaxf(tm, sc[j]);
    }
    const float mn = fmaxf(m, tm);
    const float corr = expf(m - mn);
    l *= corr;
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float p = expf(sc[j] - mn);
      l += p;
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[d] = fmaf(p, vt[j][d], acc[d]);
    }
    m = mn;
  }
  if (live) {
    float* op = out + ((size_t)b * n + i) * HD + h * 32;
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < 32; ++d) op[d] = acc[d] * inv;
  }
}

// ---- output conv (C -> out_dim, 1x1), fp32 NHWC in -> fp32 NCHW out (unet_model.py:331,368) -------------
__global__ void __launch_bounds__(256) f32_final_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out, int batch,
                                                             int hw, int C, int od) {
  const long long total = (long long)batch * od * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % hw), o = (int)((i / hw) % od), b = (int)(i / ((long long)hw * od));
    const float* xp = x + ((size_t)b * hw + px) * C;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc = fmaf(xp[c], w[o * C + c], acc);
    out[i] = acc + (bias ? bias[o] : 0.0f);
  }
}

__global__ void __launch_bounds__(256) f32_add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o,
                                                      long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = a[i], y = b[i];
    o[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}

}  // namespace

extern "C" int tedm_f32_split(const float* x, void* hi, void* lo, int64_t n, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && hi && lo && n > 0 && n % 4 == 0, "tedm_f32_split: bad arguments (n must be a multiple of 4)");
  f32_split_kernel<<<grid_cap(n / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n / 4);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_f32_stem_conv7x7(const float* x, const float* weight, const float* bias, float* out, int batch, int cin,
                                     int height, int width, int cout, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && weight && out && batch > 0 && cin > 0 && height > 0 && width > 0 && cout > 0, "tedm_f32_stem_conv7x7: bad arguments");
  TEDM_UNSUPPORTED(cout % 8 != 0, "tedm_f32_stem_conv7x7: cout=%d must be a multiple of 8", cout);
  const size_t smem = (size_t)cin * 49 * cout * sizeof(float);
  TEDM_UNSUPPORTED(smem > 96 * 1024, "tedm_f32_stem_conv7x7: cin*49*cout=%d floats do not fit in shared memory", cin * 49 * cout);
  if (smem > 48 * 1024) TEDM_CUDA(cudaFuncSetAttribute(f32_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)batch * height * width * (cout / 8);
  f32_stem_kernel<<<grid_cap(total, 256, 8), 256, smem, (cudaStream_t)stream>>>(x, weight, bias, out, batch, cin, height, width, cout);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_f32_gn_silu(const float* x, const float* gn_partial, int gn_parts, const float* gamma, const float* beta,
                                const float* scale_shift, int ss_stride, int ss_offset, const float* residual, float* out,
                                int batch, int hw, int channels, int groups, float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && gn_partial && gamma && beta && out && batch > 0 && hw > 0 && gn_parts > 0, "tedm_f32_gn_silu: bad arguments");
  TEDM_UNSUPPORTED(channels % 4 != 0 || channels > F32_GN_MAX_C || groups < 1 || groups > 32 || channels % groups != 0,
                   "tedm_f32_gn_silu: channels=%d groups=%d", channels, groups);
  dim3 grid((unsigned)grid_cap((long long)hw * channels / 4, 256, 4), (unsigned)batch);
  if (grid.x * (unsigned)batch > 16u * (unsigned)tedm_num_sms()) grid.x = (16u * (unsigned)tedm_num_sms() + batch - 1) / batch;
  if (grid.x < 1) grid.x = 1;
  f32_gn_silu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, gn_partial, gn_parts, gamma, beta, scale_shift, ss_stride, ss_offset,
                                                             residual, out, hw, channels, groups, eps);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_f32_layernorm(const float* x, const float* g, const float* residual, float* out, int64_t npix, int channels,
                                  float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && g && out && npix > 0, "tedm_f32_layernorm: bad arguments");
  const int grid = grid_cap(npix, 8, 16);
  cudaStream_t s = (cudaStream_t)stream;
  switch (channels) {
    case 64: f32_layernorm_kernel<2><<<grid, 256, 0, s>>>(x, g, residual, out, npix, eps); break;
    case 128: f32_layernorm_kernel<4><<<grid, 256, 0, s>>>(x, g, residual, out, npix, eps); break;
    case 256: f32_layernorm_kernel<8><<<grid, 256, 0, s>>>(x, g, residual, out, npix, eps); break;
    case 512: f32_layernorm_kernel<16><<<grid, 256, 0, s>>>(x, g, residual, out, npix, eps); break;
    case 1024: f32_layernorm_kernel<32><<<grid, 256, 0, s>>>(x, g, residual, out, npix, eps); break;
    default: return tedm_set_error(TEDM_ERR_UNSUPPORTED, "tedm_f32_layernorm: channels=%d (64/128/256/512/1024)", channels);
  }
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int64_t tedm_f32_linear_attention_workspace(int batch, int n, int heads) {
  const long long chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  return (long long)batch * heads * (chunks * LA_PART + 1024);
}

extern "C" int tedm_f32_linear_attention(const float* qkv, float* out, float* workspace, int batch, int n, int heads, int dim_head,
                                         float scale, tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && out && workspace && batch > 0 && n > 0 && heads > 0, "tedm_f32_linear_attention: bad arguments");
  TEDM_UNSUPPORTED(dim_head != 32 || heads > 8 || batch > 65535, "tedm_f32_linear_attention: dim_head=%d heads=%d", dim_head, heads);
  const int chunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  float* part = workspace;
  float* ctx = workspace + (size_t)batch * heads * chunks * LA_PART;
  cudaStream_t s = (cudaStream_t)stream;
  f32_linattn_ctx_kernel<<<dim3((unsigned)chunks, (unsigned)heads, (unsigned)batch), 256, 0, s>>>(qkv, part, n, heads);
  TEDM_LAUNCH_CHECK();
  f32_linattn_combine_kernel<<<batch * heads, 1024, 0, s>>>(part, ctx, chunks, n);
  TEDM_LAUNCH_CHECK();
  const long long items = (long long)n * heads;
  f32_linattn_out_kernel<<<dim3((unsigned)((items + 127) / 128), (unsigned)batch), 128, heads * 1024 * sizeof(float), s>>>(
      qkv, ctx, out, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_f32_attention(const float* qkv, float* out, float* rnorm, int batch, int n, int heads, int dim_head, float scale,
                                  tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && out && rnorm && batch > 0 && n > 0 && heads > 0, "tedm_f32_attention: bad arguments");
  TEDM_UNSUPPORTED(dim_head != 32 || batch > 65535, "tedm_f32_attention: dim_head=%d", dim_head);
  cudaStream_t s = (cudaStream_t)stream;
  f32_attn_norm_kernel<<<dim3((unsigned)(2 * heads * 32), (unsigned)batch), 256, 0, s>>>(qkv, rnorm, n, heads);
  TEDM_LAUNCH_CHECK();
  f32_attn_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)heads, (unsigned)batch), 128, 0, s>>>(qkv, rnorm, out, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_f32_final_conv1x1(const float* x, const float* weight, const float* bias, float* out, int batch, int hw,
                                      int channels, int out_dim, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && weight && out && batch > 0 && hw > 0 && channels > 0 && out_dim > 0, "tedm_f32_final_conv1x1: bad arguments");
  f32_final_conv_kernel<<<grid_cap((long long)batch * hw * out_dim, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, weight, bias, out, batch,
                                                                                                           hw, channels, out_dim);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_f32_add(const float* a, const float* b, float* out, int64_t n, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && b && out && n > 0 && n % 4 == 0, "tedm_f32_add: bad arguments (n must be a multiple of 4)");
  f32_add_kernel<<<grid_cap(n / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>((const float4*)a, (const float4*)b, (float4*)out, n / 4);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
