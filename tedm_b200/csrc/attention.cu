// Attention cores of the UNet (between the to_qkv and to_out 1x1 convs, which run on the tcgen05
// implicit-GEMM kernel).  qkv is NHWC bf16 [B][n][3*heads*32]: q = channels [0,128), k = [128,256),
// v = [256,384), head h owns 32 consecutive channels of each (the reference's
// 'b (h c) x y -> b h c (x y)' after chunk(3, dim=1)).
#include "common.cuh"

#define DH 32  // dim_head is 32 in the reference (models/unet_model.py:180,215)

// ------------------------------------------------------------------------------------------
// LinearAttention                                               models/unet_model.py:197-209
//   q = softmax_d(q) * scale ; k = softmax_n(k) ; v = v / n
//   ctx[d][e] = sum_n k[d][n] v[e][n] ; out[e][n] = sum_d ctx[d][e] q[d][n]
// Phase 1: per (image, head, chunk of LA_NP pixels): local column max m, s = sum exp(k-m),
//          ctx_c = exp(k-m)^T v.                  -> workspace partials
// Phase 2: per (image, head): fold partials with exp(m_c - M), divide by (S * n), fold in `scale`.
// Phase 3: per pixel: softmax over the head's 32 q channels, 32x32 matvec with ctx.
// ------------------------------------------------------------------------------------------
#define LA_NP 128
#define LA_PART (2 * DH + DH * DH)  // m[32], s[32], ctx[32][32]

__global__ void __launch_bounds__(256) linattn_partial_kernel(const bf16* __restrict__ qkv, float* __restrict__ ws, int n,
                                                              int heads, int nchunks) {
  __shared__ __align__(16) float sk[LA_NP][DH];
  __shared__ __align__(16) float sv[LA_NP][DH];
  __shared__ float s_max[DH];
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * DH;
  const int p0 = chunk * LA_NP;
  const int np = min(LA_NP, n - p0);
  // load k and v: (pixel, 8-channel vector) per thread-iteration
  for (int i = tid; i < LA_NP * 8; i += 256) {
    const int pi = i >> 3, part = i & 7;          // part 0..3 -> k, 4..7 -> v
    float f[8];
    if (pi < np) {
      const bf16* src = qkv + ((size_t)b * n + p0 + pi) * C3 + (part < 4 ? heads * DH : 2 * heads * DH) + h * DH + (part & 3) * 8;
      unpack8(ldg_stream(src), f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = part < 4 ? -INFINITY : 0.0f;
    }
    float* dst = part < 4 ? &sk[pi][(part & 3) * 8] : &sv[pi][(part & 3) * 8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = f[j];
  }
  __syncthreads();
  // column max: warp w owns columns 4w..4w+3
  for (int dd = 0; dd < 4; ++dd) {
    const int d = warp * 4 + dd;
    float m = -INFINITY;
    for (int pi = lane; pi < LA_NP; pi += 32) m = fmaxf(m, sk[pi][d]);
    m = warp_max(m);
    if (lane == 0) s_max[d] = m;
  }
  __syncthreads();
  for (int i = tid; i < LA_NP * DH; i += 256) {
    const int pi = i >> 5, d = i & 31;
    sk[pi][d] = __expf(sk[pi][d] - s_max[d]);   // padded rows: exp(-inf) = 0
  }
  __syncthreads();
  float* part_out = ws + (((size_t)b * heads + h) * nchunks + chunk) * LA_PART;
  for (int dd = 0; dd < 4; ++dd) {
    const int d = warp * 4 + dd;
    float s = 0.0f;
    for (int pi = lane; pi < LA_NP; pi += 32) s += sk[pi][d];
    s = warp_sum(s);
    if (lane == 0) {
      part_out[d] = s_max[d];
      part_out[DH + d] = s;
    }
  }
  // ctx: thread = (pixel quarter, 4 d, 4 e) register tile
  const int pq = tid >> 6, d0 = ((tid & 63) >> 3) * 4, e0 = (tid & 7) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int pi = pq * (LA_NP / 4); pi < (pq + 1) * (LA_NP / 4); ++pi) {
    const float4 kk = *reinterpret_cast<const float4*>(&sk[pi][d0]);
    const float4 vv = *reinterpret_cast<const float4*>(&sv[pi][e0]);
    const float ka[4] = {kk.x, kk.y, kk.z, kk.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ka[i], va[j], acc[i][j]);
  }
  __syncthreads();  // everyone is done reading sk/sv: reuse sk as the cross-quarter reduction buffer
  float* redbuf = &sk[0][0];  // [4][32][32]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) redbuf[(pq * DH + d0 + i) * DH + e0 + j] = acc[i][j];
  __syncthreads();
  for (int i = tid; i < DH * DH; i += 256)
    part_out[2 * DH + i] = redbuf[i] + redbuf[DH * DH + i] + redbuf[2 * DH * DH + i] + redbuf[3 * DH * DH + i];
}

__global__ void __launch_bounds__(256) linattn_combine_kernel(float* __restrict__ ws, float* __restrict__ ctx_out, int n,
                                                              int heads, int nchunks, float scale) {
  __shared__ float sM[DH], sS[DH];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* parts = ws + ((size_t)b * heads + h) * nchunks * LA_PART;
  if (tid < DH) {
    float M = -INFINITY;
    for (int c = 0; c < nchunks; ++c) M = fmaxf(M, parts[(size_t)c * LA_PART + tid]);
    float S = 0.0f;
    for (int c = 0; c < nchunks; ++c) S += parts[(size_t)c * LA_PART + DH + tid] * __expf(parts[(size_t)c * LA_PART + tid] - M);
    sM[tid] = M;
    sS[tid] = S;
  }
  __syncthreads();
  for (int i = tid; i < DH * DH; i += 256) {
    const int d = i >> 5;
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c)
      acc += parts[(size_t)c * LA_PART + 2 * DH + i] * __expf(parts[(size_t)c * LA_PART + d] - sM[d]);
    ctx_out[((size_t)b * heads + h) * DH * DH + i] = acc * scale / (sS[d] * (float)n);
  }
}

__global__ void __launch_bounds__(256) linattn_out_kernel(const bf16* __restrict__ qkv, const float* __restrict__ ctx,
                                                          bf16* __restrict__ out, int n, int heads) {
  __shared__ __align__(16) float sc[DH][DH];  // ctx[d][e]
  const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  for (int i = tid; i < DH * DH; i += 256) (&sc[0][0])[i] = ctx[((size_t)b * heads + h) * DH * DH + i];
  __syncthreads();
  const int pi = blockIdx.x * 256 + tid;
  if (pi >= n) return;
  const int C3 = 3 * heads * DH, C = heads * DH;
  float qv[DH];
  const bf16* qp = qkv + ((size_t)b * n + pi) * C3 + h * DH;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float f[8];
    unpack8(ldg_stream(qp + j * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) qv[j * 8 + e] = f[e];
  }
  float m = qv[0];
#pragma unroll
  for (int d = 1; d < DH; ++d) m = fmaxf(m, qv[d]);
  float s = 0.0f;
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    qv[d] = __expf(qv[d] - m);
    s += qv[d];
  }
  const float inv = 1.0f / s;
  float o[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) o[e] = 0.0f;
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float qd = qv[d] * inv;
#pragma unroll
    for (int e4 = 0; e4 < DH / 4; ++e4) {
      const float4 c4 = *reinterpret_cast<const float4*>(&sc[d][e4 * 4]);
      o[e4 * 4] = fmaf(c4.x, qd, o[e4 * 4]);
      o[e4 * 4 + 1] = fmaf(c4.y, qd, o[e4 * 4 + 1]);
      o[e4 * 4 + 2] = fmaf(c4.z, qd, o[e4 * 4 + 2]);
      o[e4 * 4 + 3] = fmaf(c4.w, qd, o[e4 * 4 + 3]);
    }
  }
  uint4* op = reinterpret_cast<uint4*>(out + ((size_t)b * n + pi) * C + h * DH);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    op[j] = make_uint4(pack_bf16x2(o[8 * j], o[8 * j + 1]), pack_bf16x2(o[8 * j + 2], o[8 * j + 3]),
                       pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), pack_bf16x2(o[8 * j + 6], o[8 * j + 7]));
}

extern "C" int64_t tedm_linear_attention_workspace(int batch, int n, int heads, int dim_head) {
  if (batch <= 0 || n <= 0 || heads <= 0 || dim_head != DH) return -1;
  const int64_t nchunks = (n + LA_NP - 1) / LA_NP;
  return (int64_t)batch * heads * (nchunks * LA_PART + DH * DH);
}

extern "C" int tedm_linear_attention_fwd(const void* qkv, void* out, float* workspace, int batch, int n, int heads,
                                         int dim_head, float scale, tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && out && workspace && batch > 0 && n > 0 && heads > 0, "tedm_linear_attention_fwd: bad arguments");
  TEDM_UNSUPPORTED(dim_head != DH, "tedm_linear_attention_fwd: dim_head=%d (only 32)", dim_head);
  TEDM_CHECK_ARG(batch <= 65535 && heads <= 65535, "tedm_linear_attention_fwd: batch/heads too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int nchunks = (n + LA_NP - 1) / LA_NP;
  float* ctx = workspace + (size_t)batch * heads * nchunks * LA_PART;
  linattn_partial_kernel<<<dim3(nchunks, heads, batch), 256, 0, s>>>((const bf16*)qkv, workspace, n, heads, nchunks);
  TEDM_LAUNCH_CHECK();
  linattn_combine_kernel<<<dim3(heads, batch), 256, 0, s>>>(workspace, ctx, n, heads, nchunks, scale);
  TEDM_LAUNCH_CHECK();
  linattn_out_kernel<<<dim3((n + 255) / 256, heads, batch), 256, 0, s>>>((const bf16*)qkv, ctx, (bf16*)out, n, heads);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// full attention of the mid block                               models/unet_model.py:229-240
//   q, k L2-normalised along n (F.normalize(dim=-1), eps 1e-12) ; sim = scale * q^T k ;
//   softmax over keys ; out = attn v.        One CTA per (image, head); thread i owns query i.
// ------------------------------------------------------------------------------------------
#define ATT_MAX_N 256
__global__ void __launch_bounds__(ATT_MAX_N) attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int n,
                                                              int heads, float scale) {
  extern __shared__ __align__(16) float smem_att[];
  float* sk = smem_att;                  // [n][32]
  float* sv = smem_att + ATT_MAX_N * DH; // [n][32]
  __shared__ float red[2][ATT_MAX_N / 32][DH];
  __shared__ float inv_norm[2][DH];
  const int h = blockIdx.x, b = blockIdx.y, i = threadIdx.x, warp = i >> 5, lane = i & 31;
  const int C3 = 3 * heads * DH, C = heads * DH;
  const bool active = i < n;
  float q[DH], k[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) q[d] = k[d] = 0.0f;
  if (active) {
    const bf16* base = qkv + ((size_t)b * n + i) * C3 + h * DH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
      unpack8(ldg_stream(base + j * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) q[j * 8 + e] = f[e];
      unpack8(ldg_stream(base + C + j * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) k[j * 8 + e] = f[e];
      unpack8(ldg_stream(base + 2 * C + j * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) sv[i * DH + j * 8 + e] = f[e];
    }
  }
  // column sums of squares over the n tokens
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float sq = warp_sum(q[d] * q[d]);
    const float sk2 = warp_sum(k[d] * k[d]);
    if (lane == 0) {
      red[0][warp][d] = sq;
      red[1][warp][d] = sk2;
    }
  }
  __syncthreads();
  if (i < 2 * DH) {
    const int which = i / DH, d = i % DH;
    float s = 0.0f;
    for (int w = 0; w < ATT_MAX_N / 32; ++w) s += red[which][w][d];
    inv_norm[which][d] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      q[d] *= inv_norm[0][d] * scale;
      sk[i * DH + d] = k[d] * inv_norm[1][d];
    }
  }
  __syncthreads();
  if (!active) return;
  float m = -INFINITY, l = 0.0f, acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.0f;
  for (int j = 0; j < n; ++j) {
    float s = 0.0f;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      const float4 kk = *reinterpret_cast<const float4*>(sk + j * DH + d4 * 4);
      s = fmaf(q[d4 * 4], kk.x, s);
      s = fmaf(q[d4 * 4 + 1], kk.y, s);
      s = fmaf(q[d4 * 4 + 2], kk.z, s);
      s = fmaf(q[d4 * 4 + 3], kk.w, s);
    }
    const float m_new = fmaxf(m, s);
    const float corr = __expf(m - m_new), pj = __expf(s - m_new);
    l = l * corr + pj;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      const float4 vv = *reinterpret_cast<const float4*>(sv + j * DH + d4 * 4);
      acc[d4 * 4] = fmaf(pj, vv.x, acc[d4 * 4] * corr);
      acc[d4 * 4 + 1] = fmaf(pj, vv.y, acc[d4 * 4 + 1] * corr);
      acc[d4 * 4 + 2] = fmaf(pj, vv.z, acc[d4 * 4 + 2] * corr);
      acc[d4 * 4 + 3] = fmaf(pj, vv.w, acc[d4 * 4 + 3] * corr);
    }
    m = m_new;
  }
  const float inv = 1.0f / l;
  uint4* op = reinterpret_cast<uint4*>(out + ((size_t)b * n + i) * C + h * DH);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    op[j] = make_uint4(pack_bf16x2(acc[8 * j] * inv, acc[8 * j + 1] * inv), pack_bf16x2(acc[8 * j + 2] * inv, acc[8 * j + 3] * inv),
                       pack_bf16x2(acc[8 * j + 4] * inv, acc[8 * j + 5] * inv), pack_bf16x2(acc[8 * j + 6] * inv, acc[8 * j + 7] * inv));
}

extern "C" int tedm_attention_fwd(const void* qkv, void* out, int batch, int n, int heads, int dim_head, float scale,
                                  tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && out && batch > 0 && n > 0 && heads > 0, "tedm_attention_fwd: bad arguments");
  TEDM_UNSUPPORTED(dim_head != DH, "tedm_attention_fwd: dim_head=%d (only 32)", dim_head);
  TEDM_UNSUPPORTED(n > ATT_MAX_N, "tedm_attention_fwd: n=%d tokens > %d", n, ATT_MAX_N);
  TEDM_CHECK_ARG(batch <= 65535, "tedm_attention_fwd: batch too large");
  const int smem = 2 * ATT_MAX_N * DH * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_kernel<<<dim3(heads, batch), ATT_MAX_N, smem, (cudaStream_t)stream>>>((const bf16*)qkv, (bf16*)out, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
