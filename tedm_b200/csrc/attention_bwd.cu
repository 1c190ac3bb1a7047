// Backward of the two attention cores (models/unet_model.py:197-209 and :229-240).  Same tensor
// conventions as attention.cu: qkv / dqkv are NHWC bf16 [B][n][3*heads*32] (q | k | v, head-major),
// out / dout are [B][n][heads*32].  Both backward passes are small next to the convolutions
// (1.4 and 0.12 GFLOP per image) and run on fp32 CUDA cores with shared-memory-resident 32x32
// matrices.
#include "common.cuh"

#define DH 32
#define LA_HEADS 4
#define LA_C (LA_HEADS * DH)
#define LA_CHUNK 1024
#define LA_PART (LA_C + LA_C * DH)   // forward per-chunk partial: s[128], ctx[128][32]  (attention.cu)
#define LAB_TILE 64
#define LAB_PITCH 132                // floats per shared-memory row (128 + 4)
#define LAB_MAT (LA_C * DH)          // 4096 floats: one 32x32 matrix per head

namespace {

__device__ __forceinline__ void load32(const bf16* p, float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t[8];
    unpack8(ldg_stream(p + j * 8), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[j * 8 + e] = t[e];
  }
}
__device__ __forceinline__ void store32(bf16* p, const float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(p + j * 8) =
        make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                   pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
}
__device__ __forceinline__ void softmax32(float (&x)[32]) {
  float m = x[0];
#pragma unroll
  for (int i = 1; i < 32; ++i) m = fmaxf(m, x[i]);
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    x[i] = __expf(x[i] - m);
    s += x[i];
  }
  const float inv = 1.0f / s;
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] *= inv;
}

// ---- LinearAttention backward ---------------------------------------------------------------
// forward (fp32 view):  p = softmax_d(q) ; kh = softmax_n(k) ; C[d][e] = (1/n) sum_n kh[n][d] v[n][e] ;
//                       out[n][e] = scale * sum_d p[n][d] C[d][e]
// backward:  dC[d][e] = scale * sum_n p[n][d] dout[n][e]
//            dp[n][d] = scale * sum_e C[d][e] dout[n][e] ;  dq = p * (dp - <p, dp>)
//            dkh[n][d] = (1/n) sum_e dC[d][e] v[n][e] ;     dk = kh * (dkh - r[d]),  r[d] = sum_e dC[d][e] C[d][e]
//            dv[n][e] = (1/n) sum_d kh[n][d] dC[d][e]
// workspace (fp32, per image): M[128] | S[128] | r[128] | C[4096] | dC[4096] | dC partials [nchunks][4096]
#define LAB_WS_FIXED (3 * LA_C + 2 * LAB_MAT)

// K-prep: fold the forward's per-chunk partials into M, S and the normalised context
__global__ void __launch_bounds__(256) linattn_bwd_prep_kernel(const float* __restrict__ fwd_ws, float* __restrict__ ws, int batch,
                                                               int n, int nchunks, long long ws_stride) {
  __shared__ float sS[LA_C];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pmax = fwd_ws + (size_t)b * nchunks * LA_C;
  const float* part = fwd_ws + (size_t)batch * nchunks * LA_C + (size_t)b * nchunks * LA_PART;
  float* w = ws + (size_t)b * ws_stride;
  if (tid < LA_C) {
    float m = -INFINITY, s = 0.0f;
    for (int c = 0; c < nchunks; ++c) {
      m = fmaxf(m, pmax[(size_t)c * LA_C + tid]);
      s += part[(size_t)c * LA_PART + tid];
    }
    w[tid] = m;
    w[LA_C + tid] = s;
    sS[tid] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < LAB_MAT; idx += 256) {
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c) acc += part[(size_t)c * LA_PART + LA_C + idx];
    w[3 * LA_C + idx] = acc / (sS[idx >> 5] * (float)n);
  }
}

// K-A: per (image, chunk) partial of sum_n p[n][d] dout[n][e]
__global__ void __launch_bounds__(256) linattn_bwd_dctx_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               float* __restrict__ ws, int n, int nchunks, long long ws_stride) {
  extern __shared__ __align__(16) float lab_smem[];
  float* sP = lab_smem;                         // [LAB_TILE][LAB_PITCH]
  float* sD = lab_smem + LAB_TILE * LAB_PITCH;  // [LAB_TILE][LAB_PITCH]
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int p0 = chunk * LA_CHUNK, p1 = min(n, p0 + LA_CHUNK);
  const int lp = tid >> 2, lh = tid & 3;        // loader role: pixel, head
  const int h = tid >> 6, d0 = ((tid & 63) >> 3) * 4, e0 = (tid & 7) * 4;   // accumulator role
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int t0 = p0; t0 < p1; t0 += LAB_TILE) {
    __syncthreads();
    {
      const int px = t0 + lp;
      float q[32], d[32];
      if (px < p1) {
        load32(qkv + ((size_t)b * n + px) * (3 * LA_C) + lh * DH, q);
        load32(dout + ((size_t)b * n + px) * LA_C + lh * DH, d);
        softmax32(q);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) q[i] = d[i] = 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        *reinterpret_cast<float4*>(sP + lp * LAB_PITCH + lh * DH + i * 4) = make_float4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
        *reinterpret_cast<float4*>(sD + lp * LAB_PITCH + lh * DH + i * 4) = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < LAB_TILE; ++px) {
      const float4 pv = *reinterpret_cast<const float4*>(sP + px * LAB_PITCH + h * DH + d0);
      const float4 dv = *reinterpret_cast<const float4*>(sD + px * LAB_PITCH + h * DH + e0);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pa[i], da[j], acc[i][j]);
    }
  }
  float* dst = ws + (size_t)b * ws_stride + LAB_WS_FIXED + (size_t)chunk * LAB_MAT + (size_t)(h * DH + d0) * DH + e0;
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(dst + i * DH) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// K-C: dC = scale * sum of chunk partials ; r[d] = sum_e dC[d][e] C[d][e]
__global__ void __launch_bounds__(256) linattn_bwd_combine_kernel(float* __restrict__ ws, int nchunks, float scale,
                                                                  long long ws_stride) {
  __shared__ float sprod[LAB_MAT];
  float* w = ws + (size_t)blockIdx.x * ws_stride;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < LAB_MAT; idx += 256) {
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c) acc += w[LAB_WS_FIXED + (size_t)c * LAB_MAT + idx];
    acc *= scale;
    w[3 * LA_C + LAB_MAT + idx] = acc;
    sprod[idx] = acc * w[3 * LA_C + idx];
  }
  __syncthreads();
  if (tid < LA_C) {
    float r = 0.0f;
    for (int e = 0; e < DH; ++e) r += sprod[tid * DH + e];
    w[2 * LA_C + tid] = r;
  }
}

// K-B: dqkv for 64 pixels per CTA; warp = 32 pixels of ONE head so that every matrix read is a broadcast
__global__ void __launch_bounds__(256) linattn_bwd_main_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               const float* __restrict__ ws, bf16* __restrict__ dqkv, int n,
                                                               float scale, long long ws_stride) {
  extern __shared__ __align__(16) float lab_smem[];
  float* sC = lab_smem;              // [128][32]
  float* sdC = lab_smem + LAB_MAT;   // [128][32]
  __shared__ float sM[LA_C], sS[LA_C], sr[LA_C];
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* w = ws + (size_t)b * ws_stride;
  for (int i = tid; i < LAB_MAT / 4; i += 256) {
    reinterpret_cast<float4*>(sC)[i] = __ldg(reinterpret_cast<const float4*>(w + 3 * LA_C) + i);
    reinterpret_cast<float4*>(sdC)[i] = __ldg(reinterpret_cast<const float4*>(w + 3 * LA_C + LAB_MAT) + i);
  }
  if (tid < LA_C) {
    sM[tid] = w[tid];
    sS[tid] = w[LA_C + tid];
    sr[tid] = w[2 * LA_C + tid];
  }
  __syncthreads();
  const int h = warp & 3;
  const int px = blockIdx.x * LAB_TILE + (warp >> 2) * 32 + lane;
  if (px >= n) return;
  const bf16* qp = qkv + ((size_t)b * n + px) * (3 * LA_C) + h * DH;
  bf16* gp = dqkv + ((size_t)b * n + px) * (3 * LA_C) + h * DH;
  const float inv_n = 1.0f / (float)n;
  float dO[32], a[32], o[32];
  load32(dout + ((size_t)b * n + px) * LA_C + h * DH, dO);
  // ---- dq
  load32(qp, a);
  softmax32(a);
  float dot = 0.0f;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const float4* row = reinterpret_cast<const float4*>(sC + (h * DH + d) * DH);
    float s = 0.0f;
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
      const float4 c = row[e4];
      s = fmaf(c.x, dO[4 * e4], s);
      s = fmaf(c.y, dO[4 * e4 + 1], s);
      s = fmaf(c.z, dO[4 * e4 + 2], s);
      s = fmaf(c.w, dO[4 * e4 + 3], s);
    }
    o[d] = s * scale;
    dot = fmaf(a[d], o[d], dot);
  }
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = a[d] * (o[d] - dot);
  store32(gp, o);
  // ---- dk (needs v), dv (needs kh)
  load32(qp + 2 * LA_C, dO);   // dO now holds v
  load32(qp + LA_C, a);        // a holds k
#pragma unroll
  for (int e = 0; e < 32; ++e) o[e] = 0.0f;   // dv accumulator
  float dk[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    const float kh = __expf(a[d] - sM[h * DH + d]) / sS[h * DH + d];
    const float4* row = reinterpret_cast<const float4*>(sdC + (h * DH + d) * DH);
    float s = 0.0f;
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
      const float4 c = row[e4];
      s = fmaf(c.x, dO[4 * e4], s);
      s = fmaf(c.y, dO[4 * e4 + 1], s);
      s = fmaf(c.z, dO[4 * e4 + 2], s);
      s = fmaf(c.w, dO[4 * e4 + 3], s);
      o[4 * e4] = fmaf(kh, c.x, o[4 * e4]);
      o[4 * e4 + 1] = fmaf(kh, c.y, o[4 * e4 + 1]);
      o[4 * e4 + 2] = fmaf(kh, c.z, o[4 * e4 + 2]);
      o[4 * e4 + 3] = fmaf(kh, c.w, o[4 * e4 + 3]);
    }
    dk[d] = kh * (s * inv_n - sr[h * DH + d]);
  }
  store32(gp + LA_C, dk);
#pragma unroll
  for (int e = 0; e < 32; ++e) o[e] *= inv_n;
  store32(gp + 2 * LA_C, o);
}

// ---- mid-block attention backward -----------------------------------------------------------
// forward:  qn = q / max(|q|_n, eps) (column norms over the n tokens), kn likewise ;  S = scale * qn kn^T ;
//           A = softmax_j S ;  O = A v.      One CTA per (image, head); thread i owns token i.
#define ATT_MAX_N 256
__global__ void __launch_bounds__(ATT_MAX_N, 1) attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                     bf16* __restrict__ dqkv, int n, int heads, float scale) {
  extern __shared__ __align__(16) float att_smem[];
  float* sq = att_smem;                       // [n][32] normalised q
  float* sk = sq + ATT_MAX_N * DH;            // [n][32] normalised k
  float* sv = sk + ATT_MAX_N * DH;            // [n][32]
  float* sdo = sv + ATT_MAX_N * DH;           // [n][32]
  __shared__ float red[2][ATT_MAX_N / 32][DH];
  __shared__ float inv_norm[2][DH], colsum[2][DH];
  __shared__ float s_m[ATT_MAX_N], s_il[ATT_MAX_N], s_delta[ATT_MAX_N];
  const int h = blockIdx.x, b = blockIdx.y, i = threadIdx.x, warp = i >> 5, lane = i & 31;
  const int C3 = 3 * heads * DH, C = heads * DH;
  const bool active = i < n;
  float x[DH], y[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) x[d] = y[d] = 0.0f;
  if (active) {
    const bf16* base = qkv + ((size_t)b * n + i) * C3 + h * DH;
    float t[32];
    load32(base, x);
    load32(base + C, y);
    load32(base + 2 * C, t);
#pragma unroll
    for (int d = 0; d < DH; ++d) sv[i * DH + d] = t[d];
    load32(dout + ((size_t)b * n + i) * C + h * DH, t);
#pragma unroll
    for (int d = 0; d < DH; ++d) sdo[i * DH + d] = t[d];
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float a = warp_sum(x[d] * x[d]);
    const float c = warp_sum(y[d] * y[d]);
    if (lane == 0) {
      red[0][warp][d] = a;
      red[1][warp][d] = c;
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
      x[d] *= inv_norm[0][d];
      y[d] *= inv_norm[1][d];
      sq[i * DH + d] = x[d];
      sk[i * DH + d] = y[d];
    }
  }
  __syncthreads();
  // ---- phase 1 (thread = query i): softmax statistics, O_i, delta_i = <dO_i, O_i>
  float dO[DH];
  float m = -INFINITY, l = 0.0f;
  if (active) {
    float acc[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      acc[d] = 0.0f;
      dO[d] = sdo[i * DH + d];
    }
    for (int j = 0; j < n; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int d4 = 0; d4 < DH / 4; ++d4) {
        const float4 kk = *reinterpret_cast<const float4*>(sk + j * DH + d4 * 4);
        s = fmaf(x[d4 * 4], kk.x, s);
        s = fmaf(x[d4 * 4 + 1], kk.y, s);
        s = fmaf(x[d4 * 4 + 2], kk.z, s);
        s = fmaf(x[d4 * 4 + 3], kk.w, s);
      }
      s *= scale;
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
    const float il = 1.0f / l;
    float delta = 0.0f;
#pragma unroll
    for (int d = 0; d < DH; ++d) delta = fmaf(dO[d], acc[d] * il, delta);
    s_m[i] = m;
    s_il[i] = il;
    s_delta[i] = delta;
  }
  __syncthreads();
  // ---- phase 2 (thread = query i): dqn_i = scale * sum_j dS_ij kn_j
  float g[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) g[d] = 0.0f;
  if (active) {
    const float il = s_il[i], delta = s_delta[i];
    for (int j = 0; j < n; ++j) {
      float s = 0.0f, da = 0.0f;
      float kk[DH];
#pragma unroll
      for (int d4 = 0; d4 < DH / 4; ++d4) {
        const float4 k4 = *reinterpret_cast<const float4*>(sk + j * DH + d4 * 4);
        const float4 v4 = *reinterpret_cast<const float4*>(sv + j * DH + d4 * 4);
        kk[d4 * 4] = k4.x; kk[d4 * 4 + 1] = k4.y; kk[d4 * 4 + 2] = k4.z; kk[d4 * 4 + 3] = k4.w;
        s = fmaf(x[d4 * 4], k4.x, s);
        s = fmaf(x[d4 * 4 + 1], k4.y, s);
        s = fmaf(x[d4 * 4 + 2], k4.z, s);
        s = fmaf(x[d4 * 4 + 3], k4.w, s);
        da = fmaf(dO[d4 * 4], v4.x, da);
        da = fmaf(dO[d4 * 4 + 1], v4.y, da);
        da = fmaf(dO[d4 * 4 + 2], v4.z, da);
        da = fmaf(dO[d4 * 4 + 3], v4.w, da);
      }
      const float a = __expf(s * scale - m) * il;
      const float ds = a * (da - delta) * scale;
#pragma unroll
      for (int d = 0; d < DH; ++d) g[d] = fmaf(ds, kk[d], g[d]);
    }
  }
  // column sums  sum_i dqn[i][d] * qn[i][d]  for the normalisation backward
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float a = warp_sum(g[d] * x[d]);
    if (lane == 0) red[0][warp][d] = a;
  }
  __syncthreads();
  if (i < DH) {
    float s = 0.0f;
    for (int w = 0; w < ATT_MAX_N / 32; ++w) s += red[0][w][i];
    colsum[0][i] = s;
  }
  __syncthreads();
  if (active) {
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = (g[d] - x[d] * colsum[0][d]) * inv_norm[0][d];
    store32(dqkv + ((size_t)b * n + i) * C3 + h * DH, o);
  }
  // ---- phase 3 (thread = key j = i): dkn_j = scale * sum_i dS_ij qn_i ; dv_j = sum_i A_ij dO_i
  float gv[DH], vj[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    g[d] = 0.0f;
    gv[d] = 0.0f;
    vj[d] = active ? sv[i * DH + d] : 0.0f;
  }
  if (active) {
    for (int r = 0; r < n; ++r) {
      float s = 0.0f, da = 0.0f;
      float qq[DH], dd[DH];
#pragma unroll
      for (int d4 = 0; d4 < DH / 4; ++d4) {
        const float4 q4 = *reinterpret_cast<const float4*>(sq + r * DH + d4 * 4);
        const float4 o4 = *reinterpret_cast<const float4*>(sdo + r * DH + d4 * 4);
        qq[d4 * 4] = q4.x; qq[d4 * 4 + 1] = q4.y; qq[d4 * 4 + 2] = q4.z; qq[d4 * 4 + 3] = q4.w;
        dd[d4 * 4] = o4.x; dd[d4 * 4 + 1] = o4.y; dd[d4 * 4 + 2] = o4.z; dd[d4 * 4 + 3] = o4.w;
        s = fmaf(q4.x, y[d4 * 4], s);
        s = fmaf(q4.y, y[d4 * 4 + 1], s);
        s = fmaf(q4.z, y[d4 * 4 + 2], s);
        s = fmaf(q4.w, y[d4 * 4 + 3], s);
        da = fmaf(o4.x, vj[d4 * 4], da);
        da = fmaf(o4.y, vj[d4 * 4 + 1], da);
        da = fmaf(o4.z, vj[d4 * 4 + 2], da);
        da = fmaf(o4.w, vj[d4 * 4 + 3], da);
      }
      const float a = __expf(s * scale - s_m[r]) * s_il[r];
      const float ds = a * (da - s_delta[r]) * scale;
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        g[d] = fmaf(ds, qq[d], g[d]);
        gv[d] = fmaf(a, dd[d], gv[d]);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float a = warp_sum(g[d] * y[d]);
    if (lane == 0) red[1][warp][d] = a;
  }
  __syncthreads();
  if (i < DH) {
    float s = 0.0f;
    for (int w = 0; w < ATT_MAX_N / 32; ++w) s += red[1][w][i];
    colsum[1][i] = s;
  }
  __syncthreads();
  if (active) {
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = (g[d] - y[d] * colsum[1][d]) * inv_norm[1][d];
    store32(dqkv + ((size_t)b * n + i) * C3 + C + h * DH, o);
    store32(dqkv + ((size_t)b * n + i) * C3 + 2 * C + h * DH, gv);
  }
}

}  // namespace

extern "C" int64_t tedm_linear_attention_bwd_workspace(int batch, int n, int heads, int dim_head) {
  if (batch <= 0 || n <= 0 || heads != LA_HEADS || dim_head != DH) return -1;
  const int64_t nchunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  return (int64_t)batch * (LAB_WS_FIXED + nchunks * LAB_MAT);
}

extern "C" int tedm_linear_attention_bwd(const void* qkv, const void* dout, const float* fwd_workspace, void* dqkv,
                                         float* workspace, int batch, int n, int heads, int dim_head, float scale,
                                         tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && dout && fwd_workspace && dqkv && workspace && batch > 0 && n > 0, "tedm_linear_attention_bwd: bad arguments");
  TEDM_UNSUPPORTED(dim_head != DH || heads != LA_HEADS, "tedm_linear_attention_bwd: heads=%d dim_head=%d (only 4 x 32)", heads, dim_head);
  TEDM_CHECK_ARG(batch <= 65535, "tedm_linear_attention_bwd: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int nchunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  const long long ws_stride = LAB_WS_FIXED + (long long)nchunks * LAB_MAT;
  const int smem_a = 2 * LAB_TILE * LAB_PITCH * (int)sizeof(float);
  const int smem_b = 2 * LAB_MAT * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(linattn_bwd_dctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
    TEDM_CUDA(cudaFuncSetAttribute(linattn_bwd_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    configured = true;
  }
  linattn_bwd_prep_kernel<<<batch, 256, 0, s>>>(fwd_workspace, workspace, batch, n, nchunks, ws_stride);
  TEDM_LAUNCH_CHECK();
  linattn_bwd_dctx_kernel<<<dim3(nchunks, batch), 256, smem_a, s>>>((const bf16*)qkv, (const bf16*)dout, workspace, n, nchunks,
                                                                    ws_stride);
  TEDM_LAUNCH_CHECK();
  linattn_bwd_combine_kernel<<<batch, 256, 0, s>>>(workspace, nchunks, scale, ws_stride);
  TEDM_LAUNCH_CHECK();
  linattn_bwd_main_kernel<<<dim3((n + LAB_TILE - 1) / LAB_TILE, batch), 256, smem_b, s>>>(
      (const bf16*)qkv, (const bf16*)dout, workspace, (bf16*)dqkv, n, scale, ws_stride);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_attention_bwd(const void* qkv, const void* dout, void* dqkv, int batch, int n, int heads, int dim_head,
                                  float scale, tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && dout && dqkv && batch > 0 && n > 0 && heads > 0, "tedm_attention_bwd: bad arguments");
  TEDM_UNSUPPORTED(dim_head != DH, "tedm_attention_bwd: dim_head=%d (only 32)", dim_head);
  TEDM_UNSUPPORTED(n > ATT_MAX_N, "tedm_attention_bwd: n=%d tokens > %d", n, ATT_MAX_N);
  TEDM_CHECK_ARG(batch <= 65535, "tedm_attention_bwd: batch too large");
  const int smem = 4 * ATT_MAX_N * DH * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  attention_bwd_kernel<<<dim3(heads, batch), ATT_MAX_N, smem, (cudaStream_t)stream>>>((const bf16*)qkv, (const bf16*)dout,
                                                                                     (bf16*)dqkv, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
