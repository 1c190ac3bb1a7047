// Backward passes of the memory-bound UNet pieces (training step of models/diffusion_model.py:120-143
// through models/unet_model.py): GroupNorm+scale/shift+SiLU, channel LayerNorm, bias column sums,
// final 1x1 conv, 7x7 stem weight gradient, the time-embedding MLPs, weight re-layouts for the
// dgrad / wgrad tcgen05 GEMMs, and a fused Adam update.  Activations and their gradients are NHWC
// bf16; every parameter gradient is fp32 and is ACCUMULATED (+=) into its destination, like
// torch's .grad.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Per-channel sums held by the threads of a CTA as 8-channel vectors: thread `tid` owns channel vector
// (tid % cvec).  Reduces over the threads that share a vector and atomically adds C results to dst.
template <int NQ>
__device__ __forceinline__ void cta_channel_reduce(float (&acc)[NQ][8], float* red /*[NQ][nthreads][8]*/, int nthreads,
                                                   int cvec, float* dst, int dst_stride /*between quantities*/) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int qn = 0; qn < NQ; ++qn)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[((size_t)qn * nthreads + tid) * 8 + j] = acc[qn][j];
  __syncthreads();
  const int C = cvec * 8;
  for (int item = tid; item < NQ * C; item += nthreads) {
    const int qn = item / C, c = item % C, cv = c >> 3, j = c & 7;
    float s = 0.0f;
    for (int k = cv; k < nthreads; k += cvec) s += red[((size_t)qn * nthreads + k) * 8 + j];
    atomicAdd(dst + (size_t)qn * dst_stride + c, s);
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm + (scale+1)/shift + SiLU backward                  models/unet_model.py:126-135
//   forward:  xh = (x - mean_g) * rstd_g ;  z = (xh*gamma + beta) * sc + sh ;  y = silu(z) [+ residual]
//   pass 1 (reduce): P0[b][c] = sum_p dz, P1[b][c] = sum_p dz*xh, P2[b][c] = sum_p x   (dz = dy * silu'(z))
//   pass 2 (apply):  dx = rstd * (dxh - mean_g(dxh) - xh * mean_g(dxh*xh)),  dxh = dz*gamma*sc
//                    + parameter gradients from P0/P1/P2 (one CTA per image does them).
// ------------------------------------------------------------------------------------------
#define GNB_MAX_C 512

struct GnStats {
  float mean[32], rstd[32];
};

__device__ __forceinline__ void gn_fold_stats(const float* partial, int parts, int b, int groups, int hw, int cpg, float eps,
                                              float* s_mean, float* s_rstd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int g = warp; g < groups; g += nw) {
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
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            const float* __restrict__ partial, int parts,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ ss, int ss_stride, int ss_offset,
                                                            float* __restrict__ red_out /*[B][3][C]*/,
                                                            bf16* __restrict__ dz_out, int hw, int C, int groups, float eps,
                                                            int vec_per_cta) {
  __shared__ float sA[GNB_MAX_C], sB[GNB_MAX_C], sR[GNB_MAX_C], sQ[GNB_MAX_C];
  __shared__ float s_mean[32], s_rstd[32];
  extern __shared__ float red[];  // [3][256][8]
  const int b = blockIdx.y, cpg = C / groups;
  gn_fold_stats(partial, parts, b, groups, hw, cpg, eps, s_mean, s_rstd);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    float a = s_rstd[g] * gamma[c];
    float bb = beta[c] - s_mean[g] * a;
    if (ss) {
      const float sc = ss[(size_t)b * ss_stride + ss_offset + c] + 1.0f;
      const float sh = ss[(size_t)b * ss_stride + ss_offset + C + c];
      a *= sc;
      bb = bb * sc + sh;
    }
    sA[c] = a;
    sB[c] = bb;
    sR[c] = s_rstd[g];
    sQ[c] = -s_mean[g] * s_rstd[g];
  }
  __syncthreads();
  const int cvec = C >> 3;
  const long long nvec = (long long)hw * cvec;
  const long long v0 = (long long)blockIdx.x * vec_per_cta;   // multiple of 256, hence of cvec
  long long v1 = v0 + vec_per_cta;
  if (v1 > nvec) v1 = nvec;
  const size_t img = (size_t)b * hw * C;
  const int c0 = (threadIdx.x % cvec) << 3;
  float acc[3][8], ca[8], cb[8], cr[8], cq[8];   // this thread always meets the same 8 channels
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[0][j] = acc[1][j] = acc[2][j] = 0.0f;
    ca[j] = sA[c0 + j];
    cb[j] = sB[c0 + j];
    cr[j] = sR[c0 + j];
    cq[j] = sQ[c0 + j];
  }
  for (long long vb = v0 + threadIdx.x; vb < v1; vb += 2 * 256) {
   uint4 xv[2], dv[2];    // both iterations' loads are issued before either is consumed
#pragma unroll
   for (int u = 0; u < 2; ++u) {
     const long long v = vb + u * 256;
     if (v < v1) {
       xv[u] = ldg_stream(x + img + v * 8);
       dv[u] = ldg_stream(dy + img + v * 8);
     }
   }
#pragma unroll
   for (int u = 0; u < 2; ++u) {
    const long long v = vb + u * 256;
    if (v >= v1) break;
    float f[8], d[8];
    unpack8(xv[u], f);
    unpack8(dv[u], d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(f[j], ca[j], cb[j]);
      const float sg = fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);     // sigmoid with one MUFU op
      const float dz = d[j] * sg * fmaf(z, 1.0f - sg, 1.0f);
      acc[0][j] += dz;
      acc[1][j] = fmaf(dz, fmaf(f[j], cr[j], cq[j]), acc[1][j]);
      acc[2][j] += f[j];
      d[j] = dz;
    }
    // dz = dy * silu'(z) is parked (bf16) in the dx buffer: the apply pass reads it back instead of recomputing the
    // sigmoid, and overwrites it in place
    *reinterpret_cast<uint4*>(dz_out + img + v * 8) = pack8(d);
   }
  }
  cta_channel_reduce<3>(acc, red, 256, cvec, red_out + (size_t)b * 3 * C, C);
}

__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const bf16* __restrict__ x,
                                                           const float* __restrict__ partial, int parts,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ ss, int ss_stride, int ss_offset,
                                                           const float* __restrict__ red_in /*[B][3][C]*/,
                                                           bf16* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, float* __restrict__ dbias,
                                                           float* __restrict__ dss, int hw, int C, int groups, float eps,
                                                           int vec_per_cta) {
  __shared__ float sK1[GNB_MAX_C], sK2[GNB_MAX_C], sK3[GNB_MAX_C];
  __shared__ float sT1[GNB_MAX_C], sT2[GNB_MAX_C];
  __shared__ float s_mean[32], s_rstd[32], s_m1[32], s_m2[32];
  const int b = blockIdx.y, cpg = C / groups;
  gn_fold_stats(partial, parts, b, groups, hw, cpg, eps, s_mean, s_rstd);
  const float* P0 = red_in + (size_t)b * 3 * C;
  const float* P1 = P0 + C;
  const float* P2 = P1 + C;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float sc = ss ? ss[(size_t)b * ss_stride + ss_offset + c] + 1.0f : 1.0f;
    const float a = gamma[c] * sc;
    sT1[c] = a * P0[c];
    sT2[c] = a * P1[c];
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float t1 = 0.0f, t2 = 0.0f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      t1 += sT1[c];
      t2 += sT2[c];
    }
    const float inv_n = 1.0f / ((float)hw * (float)cpg);
    s_m1[g] = t1 * inv_n;
    s_m2[g] = t2 * inv_n;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    const float rstd = s_rstd[g], mean = s_mean[g];
    const float sc = ss ? ss[(size_t)b * ss_stride + ss_offset + c] + 1.0f : 1.0f;
    const float gm = gamma[c], bt = beta[c];
    const float k1 = gm * sc * rstd;
    const float k2 = -rstd * rstd * s_m2[g];
    sK1[c] = k1;
    sK2[c] = k2;
    sK3[c] = -rstd * s_m1[g] - mean * k2;
    if (blockIdx.x == 0) {
      // parameter gradients of this image (accumulated over images with atomics)
      const float p0 = P0[c], p1 = P1[c], p2 = P2[c];
      atomicAdd(dgamma + c, sc * p1);
      atomicAdd(dbeta + c, sc * p0);
      if (dbias) atomicAdd(dbias + c, k1 * p0 + k2 * (p2 - (float)hw * mean) - rstd * s_m1[g] * (float)hw);
      if (dss) {
        dss[(size_t)b * ss_stride + ss_offset + c] = gm * p1 + bt * p0;
        dss[(size_t)b * ss_stride + ss_offset + C + c] = p0;
      }
    }
  }
  __syncthreads();
  const int cvec = C >> 3;
  const long long nvec = (long long)hw * cvec;
  const long long v0 = (long long)blockIdx.x * vec_per_cta;
  long long v1 = v0 + vec_per_cta;
  if (v1 > nvec) v1 = nvec;
  const size_t img = (size_t)b * hw * C;
  const int c0 = (threadIdx.x % cvec) << 3;
  float k1[8], k2[8], k3[8];       // this thread always meets the same 8 channels
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    k1[j] = sK1[c0 + j];
    k2[j] = sK2[c0 + j];
    k3[j] = sK3[c0 + j];
  }
  for (long long vb = v0 + threadIdx.x; vb < v1; vb += 4 * 256) {
    uint4 xv[4], zv[4];    // four iterations' loads in flight; dz was parked in dx by the reduce pass
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = vb + u * 256;
      if (v < v1) {
        xv[u] = ldg_stream(x + img + v * 8);
        zv[u] = *reinterpret_cast<const uint4*>(dx + img + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = vb + u * 256;
      if (v >= v1) break;
      float f[8], d[8];
      unpack8(xv[u], f);
      unpack8(zv[u], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = fmaf(d[j], k1[j], fmaf(f[j], k2[j], k3[j]));
      *reinterpret_cast<uint4*>(dx + img + v * 8) = pack8(d);
    }
  }
}

// ------------------------------------------------------------------------------------------
// channel LayerNorm backward                                   models/unet_model.py:52-61
// ------------------------------------------------------------------------------------------
template <int L, int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ g,
                                                            const bf16* __restrict__ dy, const bf16* __restrict__ add,
                                                            bf16* __restrict__ dx, float* __restrict__ dg, long long npix,
                                                            float eps) {
  constexpr int C = L * NV * 8;
  constexpr int PIX_PER_WARP = 32 / L;
  __shared__ float sdg[C];
  for (int c = threadIdx.x; c < C; c += 256) sdg[c] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % L;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float gain[NV][8], dgacc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gain[v][j] = g[(v * L + sub) * 8 + j];
      dgacc[v][j] = 0.0f;
    }
  for (long long p0 = warp_global * PIX_PER_WARP; p0 < npix; p0 += nwarps * PIX_PER_WARP) {
    const long long p = p0 + lane / L;
    const bool ok = p < npix;
    float f[NV][8], d[NV][8];
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (ok) {
        unpack8(ldg_stream(x + (size_t)p * C + (v * L + sub) * 8), f[v]);
        unpack8(ldg_stream(dy + (size_t)p * C + (v * L + sub) * 8), d[v]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[v][j] = d[v][j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[v][j];
    }
    s = group_sum<L>(s);
    const float mean = s * (1.0f / C);
    float q = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[v][j] -= mean;
        q = fmaf(f[v][j], f[v][j], q);
      }
    q = group_sum<L>(q);
    const float rstd = rsqrtf(q * (1.0f / C) + eps);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[v][j] *= rstd;                              // xh
        dgacc[v][j] = fmaf(d[v][j], f[v][j], dgacc[v][j]);
        d[v][j] *= gain[v][j];                        // dxh
        s1 += d[v][j];
        s2 = fmaf(d[v][j], f[v][j], s2);
      }
    s1 = group_sum<L>(s1) * (1.0f / C);
    s2 = group_sum<L>(s2) * (1.0f / C);
    if (ok) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (d[v][j] - s1 - f[v][j] * s2);
        if (add) {
          float r[8];
          unpack8(ldg_stream(add + (size_t)p * C + (v * L + sub) * 8), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        *reinterpret_cast<uint4*>(dx + (size_t)p * C + (v * L + sub) * 8) = pack8(o);
      }
    }
  }
  // lanes sub, sub+L, ... of a warp hold the same channels
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = dgacc[v][j];
#pragma unroll
      for (int o = L; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane < L) atomicAdd(&sdg[(v * L + sub) * 8 + j], t);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) atomicAdd(dg + c, sdg[c]);
}

template <int L, int NV>
int launch_layernorm_bwd(const void* x, const float* g, const void* dy, const void* add, void* dx, float* dg, long long npix,
                         float eps, cudaStream_t stream) {
  const long long warps = (npix + (32 / L) - 1) / (32 / L);
  long long blocks = (warps + 7) / 8;
  static int cap = 0;   // one resident wave
  if (cap == 0) cap = resident_ctas(layernorm_bwd_kernel<L, NV>, 256, 0);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  layernorm_bwd_kernel<L, NV><<<(int)blocks, 256, 0, stream>>>((const bf16*)x, g, (const bf16*)dy, (const bf16*)add,
                                                               (bf16*)dx, dg, npix, eps);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// column sums of an NHWC bf16 gradient (conv bias gradients)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out, long long npix,
                                                     int C, int rows_per_cta) {
  extern __shared__ float red[];  // [nthreads][8]
  const int cvec = C >> 3, nthreads = blockDim.x;
  const int cv = threadIdx.x % cvec, r0 = threadIdx.x / cvec, rstep = nthreads / cvec;
  const long long p0 = (long long)blockIdx.x * rows_per_cta;
  long long p1 = p0 + rows_per_cta;
  if (p1 > npix) p1 = npix;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.0f;
  for (long long p = p0 + r0; p < p1; p += rstep) {
    float f[8];
    unpack8(ldg_stream(dy + (size_t)p * C + cv * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += f[j];
  }
  cta_channel_reduce<1>(acc, red, nthreads, cvec, out, 0);
}

// out = a + b (bf16, fp32 add): merges the two gradient streams that meet at a skip connection
__global__ void __launch_bounds__(256) add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                       uint4* __restrict__ out, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    unpack8(ldg_stream(a + i), fa);
    unpack8(ldg_stream(b + i), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    out[i] = pack8(fa);
  }
}

// ------------------------------------------------------------------------------------------
// final 1x1 conv backward                                       models/unet_model.py:331,368
//   dh[p][c] = sum_od dout[b][od][p] * w[od][c]; dw[od][c] += sum_p dout*h; db[od] += sum_p dout
// ------------------------------------------------------------------------------------------
#define FC_MAX_OD 4
__global__ void __launch_bounds__(256) final_conv_bwd_kernel(const bf16* __restrict__ h, const float* __restrict__ w,
                                                             const float* __restrict__ dout, bf16* __restrict__ dh,
                                                             float* __restrict__ dw, float* __restrict__ db, int hw, int C,
                                                             int out_dim, long long npix, int rows_per_cta) {
  extern __shared__ float red[];  // [out_dim][256][8]
  const int cvec = C >> 3;
  const int cv = threadIdx.x % cvec, r0 = threadIdx.x / cvec, rstep = 256 / cvec;
  const long long p0 = (long long)blockIdx.x * rows_per_cta;
  long long p1 = p0 + rows_per_cta;
  if (p1 > npix) p1 = npix;
  float wv[FC_MAX_OD][8], acc[FC_MAX_OD][8], dbacc[FC_MAX_OD];
#pragma unroll
  for (int od = 0; od < FC_MAX_OD; ++od) {
    dbacc[od] = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wv[od][j] = od < out_dim ? w[od * C + cv * 8 + j] : 0.0f;
      acc[od][j] = 0.0f;
    }
  }
  for (long long p = p0 + r0; p < p1; p += rstep) {
    const long long b = p / hw;
    const int pi = (int)(p % hw);
    float f[8], o[8];
    unpack8(ldg_stream(h + (size_t)p * C + cv * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
#pragma unroll
    for (int od = 0; od < FC_MAX_OD; ++od) {
      if (od < out_dim) {
        const float g = __ldg(dout + ((size_t)b * out_dim + od) * hw + pi);
        if (cv == 0) dbacc[od] += g;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = fmaf(g, wv[od][j], o[j]);
          acc[od][j] = fmaf(g, f[j], acc[od][j]);
        }
      }
    }
    *reinterpret_cast<uint4*>(dh + (size_t)p * C + cv * 8) = pack8(o);
  }
  // per-CTA reduction: dw (channel-wise) then db
  for (int od = 0; od < out_dim; ++od) {
    float one[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) one[0][j] = acc[od][j];
    cta_channel_reduce<1>(one, red, 256, cvec, dw + (size_t)od * C, 0);
    __syncthreads();
    float s = cv == 0 ? dbacc[od] : 0.0f;
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.0f) atomicAdd(db + od, s);
  }
}

// ------------------------------------------------------------------------------------------
// 7x7 stem conv weight gradient (input fp32 NCHW, dy NHWC bf16)   models/unet_model.py:267,334
//   dW[co][ci][ky][kx] += sum_{b,y,x} dy[b][y][x][co] * in[b][ci][y+ky-3][x+kx-3];  db[co] += sum dy
// CTA = one 64-pixel row segment at a time: dy tile in smem as fp32, 7 input rows with halo;
// thread = (4 output channels) x (taps tg, tg+16, tg+32, tg+48).
// ------------------------------------------------------------------------------------------
#define STW_PX 64
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ db, int batch, int cin,
                                                         int H, int W, int cout) {
  __shared__ __align__(16) float sdy[STW_PX][64];
  __shared__ float sx[7][STW_PX + 6];
  const int co4 = threadIdx.x & 15, tg = threadIdx.x >> 4;
  const int segs = (W + STW_PX - 1) / STW_PX;
  const long long ntiles = (long long)batch * H * segs;
  for (int cb = 0; cb < cout; cb += 64) {       // output channels in blocks of 64
    for (int ci = 0; ci < cin; ++ci) {
      float acc[4][4];
      float bacc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = 0.0f;
      int tky[4], tkx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tap = tg + 16 * k;
        tky[k] = tap < 49 ? tap / 7 : 0;
        tkx[k] = tap < 49 ? tap % 7 : 0;
      }
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int sx0 = (int)(tile % segs) * STW_PX;
        const int y = (int)((tile / segs) % H);
        const int b = (int)(tile / ((long long)segs * H));
        __syncthreads();
        for (int i = threadIdx.x; i < STW_PX * 8; i += 256) {
          const int px = i >> 3, v = i & 7;
          float f[8];
          if (sx0 + px < W && cb + v * 8 < cout) unpack8(ldg_stream(dy + (((size_t)b * H + y) * W + sx0 + px) * cout + cb + v * 8), f);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) sdy[px][v * 8 + j] = f[j];
        }
        for (int i = threadIdx.x; i < 7 * (STW_PX + 6); i += 256) {
          const int ky = i / (STW_PX + 6), xx = i % (STW_PX + 6);
          const int yy = y + ky - 3, gx = sx0 + xx - 3;
          sx[ky][xx] = (yy >= 0 && yy < H && gx >= 0 && gx < W) ? __ldg(x + (((size_t)b * cin + ci) * H + yy) * W + gx) : 0.0f;
        }
        __syncthreads();
#pragma unroll 4
        for (int px = 0; px < STW_PX; ++px) {
          const float4 d = *reinterpret_cast<const float4*>(&sdy[px][co4 * 4]);
          if (ci == 0 && tg == 0) {
            bacc[0] += d.x; bacc[1] += d.y; bacc[2] += d.z; bacc[3] += d.w;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xin = sx[tky[k]][px + tkx[k]];
            acc[k][0] = fmaf(xin, d.x, acc[k][0]);
            acc[k][1] = fmaf(xin, d.y, acc[k][1]);
            acc[k][2] = fmaf(xin, d.z, acc[k][2]);
            acc[k][3] = fmaf(xin, d.w, acc[k][3]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tap = tg + 16 * k;
        if (tap < 49) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int co = cb + co4 * 4 + j;
            if (co < cout) atomicAdd(dw + ((size_t)co * cin + ci) * 49 + tap, acc[k][j]);
          }
        }
      }
      if (ci == 0 && tg == 0 && db) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cb + co4 * 4 + j < cout) atomicAdd(db + cb + co4 * 4 + j, bacc[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// small fp32 linear layers of the time-embedding path (unet_model.py:150-152, 287-292), backward
//   dY = dYraw * act'(Ypre);  dW[j][k] += sum_b dY[b][j] * act(X[b][k]);  db[j] += sum_b dY[b][j];
//   dXraw[b][k] (+)= sum_j dY[b][j] * W[j][k]
// act: 0 identity, 1 SiLU, 2 GELU(erf)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return v / (1.0f + expf(-v));
  if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
  if (act == 1) {
    const float s = 1.0f / (1.0f + expf(-v));
    return s * (1.0f + v * (1.0f - s));
  }
  if (act == 2) return 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v);
  return 1.0f;
}

#define LB_BCHUNK 32
// grid: (ceil(N/8)); block 256 = 8 warps, warp w owns row j; lanes stride over k
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ dyraw, const float* __restrict__ ypre,
                                                           int act_y, const float* __restrict__ x, int act_x,
                                                           float* __restrict__ dw, float* __restrict__ db, int batch, int N,
                                                           int K) {
  extern __shared__ float sm[];          // [LB_BCHUNK][K] act(x)  +  [LB_BCHUNK][8] dy
  float* sx = sm;
  float* sdy = sm + LB_BCHUNK * K;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  float bsum = 0.0f;
  for (int b0 = 0; b0 < batch; b0 += LB_BCHUNK) {
    const int nb = min(LB_BCHUNK, batch - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * K; i += 256) sx[i] = act_fwd(x[(size_t)b0 * K + i], act_x);
    for (int i = threadIdx.x; i < nb * 8; i += 256) {
      const int bb = i >> 3, jj = blockIdx.x * 8 + (i & 7);
      float v = 0.0f;
      if (jj < N) {
        v = dyraw[(size_t)(b0 + bb) * N + jj];
        if (act_y) v *= act_grad(ypre[(size_t)(b0 + bb) * N + jj], act_y);
      }
      sdy[i] = v;
    }
    __syncthreads();
    if (j < N) {
      for (int k = lane; k < K; k += 32) {
        float acc = 0.0f;
        for (int bb = 0; bb < nb; ++bb) acc = fmaf(sdy[bb * 8 + warp], sx[bb * K + k], acc);
        dw[(size_t)j * K + k] += acc;   // rows are owned by exactly one warp: no atomics needed
      }
      if (lane == 0)
        for (int bb = 0; bb < nb; ++bb) bsum += sdy[bb * 8 + warp];
    }
  }
  if (j < N && lane == 0 && db) db[j] += bsum;
}

// grid: (ceil(K/256), batch, jsplit): dxraw[b][k] += sum_{j in slice} dY[b][j] W[j][k]   (dxraw zeroed by the caller)
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ dyraw, const float* __restrict__ ypre,
                                                           int act_y, const float* __restrict__ w, float* __restrict__ dxraw,
                                                           int N, int K, int jchunk) {
  __shared__ float sdy[512];
  const int b = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.z * jchunk, j1 = min(N, j0 + jchunk);
  float acc = 0.0f;
  for (int jb = j0; jb < j1; jb += 512) {
    const int nj = min(512, j1 - jb);
    __syncthreads();
    for (int i = threadIdx.x; i < nj; i += 256) {
      float v = dyraw[(size_t)b * N + jb + i];
      if (act_y) v *= act_grad(ypre[(size_t)b * N + jb + i], act_y);
      sdy[i] = v;
    }
    __syncthreads();
    if (k < K)
      for (int i = 0; i < nj; ++i) acc = fmaf(sdy[i], __ldg(w + (size_t)(jb + i) * K + k), acc);
  }
  if (k < K) atomicAdd(dxraw + (size_t)b * K + k, acc);
}

// time embedding forward that also keeps what the backward needs (emb, pre-GELU hidden)
__global__ void __launch_bounds__(256) time_embed_train_kernel(const int64_t* __restrict__ t, const float* __restrict__ freq,
                                                               const float* __restrict__ w1, const float* __restrict__ b1,
                                                               const float* __restrict__ w2, const float* __restrict__ b2,
                                                               float* __restrict__ emb_out, float* __restrict__ hid_pre,
                                                               float* __restrict__ temb, int dim, int tdim) {
  extern __shared__ float sm[];
  float* emb = sm;
  float* hid = sm + dim;
  const int b = blockIdx.x, half = dim / 2;
  const float tf = (float)t[b];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = __fmul_rn(tf, freq[i]);
    emb[i] = sinf(a);
    emb[half + i] = cosf(a);
    emb_out[(size_t)b * dim + i] = emb[i];
    emb_out[(size_t)b * dim + half + i] = emb[half + i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < tdim; j += nw) {
    float acc = 0.0f;
    for (int k = lane; k < dim; k += 32) acc = fmaf(w1[(size_t)j * dim + k], emb[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + b1[j];
      hid_pre[(size_t)b * tdim + j] = v;
      hid[j] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    }
  }
  __syncthreads();
  for (int j = warp; j < tdim; j += nw) {
    float acc = 0.0f;
    for (int k = lane; k < tdim; k += 32) acc = fmaf(w2[(size_t)j * tdim + k], hid[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) temb[(size_t)b * tdim + j] = acc + b2[j];
  }
}

// ------------------------------------------------------------------------------------------
// weight re-layouts for the backward GEMMs
// ------------------------------------------------------------------------------------------
// fp32 OIHW parameter -> bf16 operand of the DATA-gradient conv (which runs on the forward kernel):
//   fwd mode 0/1 (k x k, same padding): out[ci][ky][kx][co] = w[co][ci][K-1-ky][K-1-kx]           (run as mode 0/1)
//   fwd mode 2 (4x4 stride 2):          out[par][ci][a][b][co] = w[co][ci][3-2a-py][3-2b-px]      (run as mode 3)
//   fwd mode 3 (nearest x2 + 3x3):      out[ci][ky][kx][co] (4x4) = folded[par][co][a][b][ci],
//                                       py = (ky+1)&1, a = (3-ky-py)/2 (same in x)                (run as mode 2)
__device__ __forceinline__ int fold_src_lo(int par, int a) { return par == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2); }
__device__ __forceinline__ int fold_src_hi(int par, int a) { return par == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2); }

__global__ void weight_to_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ o, int cout, int cin, int mode) {
  const int taps = mode == 0 ? 1 : mode == 1 ? 9 : 16;
  const long long total = (long long)cin * taps * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    float v;
    if (mode == 0) {
      const int ci = (int)(i / cout);
      v = w[(size_t)co * cin + ci];
    } else if (mode == 1) {
      const int tap = (int)((i / cout) % 9), ci = (int)(i / (9LL * cout));
      v = w[((size_t)co * cin + ci) * 9 + (8 - tap)];
    } else if (mode == 2) {
      // i = ((par*cin + ci)*4 + a*2 + b)*cout + co
      const int ab = (int)((i / cout) % 4), ci = (int)((i / (4LL * cout)) % cin), par = (int)(i / (4LL * cout * cin));
      const int a = ab >> 1, b = ab & 1, py = par >> 1, px = par & 1;
      v = w[((size_t)co * cin + ci) * 16 + (3 - 2 * a - py) * 4 + (3 - 2 * b - px)];
    } else {
      // i = ((ci*4 + ky)*4 + kx)*cout + co ; source is the 3x3 kernel
      const int kx = (int)((i / cout) % 4), ky = (int)((i / (4LL * cout)) % 4), ci = (int)(i / (16LL * cout));
      const int py = (ky + 1) & 1, px = (kx + 1) & 1, a = (3 - ky - py) >> 1, b = (3 - kx - px) >> 1;
      v = 0.0f;
      for (int y3 = fold_src_lo(py, a); y3 <= fold_src_hi(py, a); ++y3)
        for (int x3 = fold_src_lo(px, b); x3 <= fold_src_hi(px, b); ++x3) v += w[((size_t)co * cin + ci) * 9 + y3 * 3 + x3];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

// fp32 [cout][taps][cin] (tedm_conv_igemm_wgrad) -> += fp32 OIHW gradient
__global__ void wgrad_to_oihw_kernel(const float* __restrict__ dw, float* __restrict__ grad, int cout, int cin, int mode) {
  const int khw = mode == 0 ? 1 : (mode == 2 ? 16 : 9);
  const long long total = (long long)cout * cin * khw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % khw), ci = (int)((i / khw) % cin), co = (int)(i / ((long long)khw * cin));
    float v = 0.0f;
    if (mode != 3) {
      v = dw[((size_t)co * khw + tap) * cin + ci];
    } else {
      const int ky = tap / 3, kx = tap % 3;
      for (int py = 0; py < 2; ++py) {
        const int a = py == 0 ? (ky == 0 ? 0 : 1) : (ky == 2 ? 1 : 0);
        for (int px = 0; px < 2; ++px) {
          const int b = px == 0 ? (kx == 0 ? 0 : 1) : (kx == 2 ? 1 : 0);
          v += dw[((size_t)co * 16 + (py * 2 + px) * 4 + a * 2 + b) * cin + ci];
        }
      }
    }
    grad[i] += v;
  }
}

// ------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, trainers/train_CXR14.py:139) over a flat fp32 parameter arena
// ------------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, long long n4, float lr, float beta1, float beta2,
                                                   float eps, float weight_decay, const int* __restrict__ step_dev, int step_host,
                                                   float grad_scale) {
  // bias corrections from the 1-based step count (device counter when the step is replayed from a CUDA graph)
  const double step = (double)(step_dev ? *step_dev : step_host);
  const float inv_bc1 = (float)(1.0 / (1.0 - pow((double)beta1, step)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, step)));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = ga[j] * grad_scale;
      if (weight_decay != 0.0f) gr = fmaf(weight_decay, pa[j], gr);
      ma[j] = fmaf(beta1, ma[j], (1.0f - beta1) * gr);
      va[j] = fmaf(beta2, va[j], (1.0f - beta2) * gr * gr);
      const float denom = sqrtf(va[j]) * inv_sqrt_bc2 + eps;
      pa[j] -= lr * inv_bc1 * (ma[j] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

int grid_for(long long items, int per_block, int cap_mult) {
  long long blocks = (items + per_block - 1) / per_block;
  const long long cap = (long long)tedm_num_sms() * cap_mult;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" int tedm_gn_silu_bwd(const void* x, const void* dy, const float* gn_partial, int gn_parts, const float* gamma,
                                const float* beta, const float* scale_shift, int ss_stride, int ss_offset, void* dx,
                                float* workspace, float* dgamma, float* dbeta, float* dbias, float* dscale_shift, int batch,
                                int hw, int channels, int groups, float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && dy && gn_partial && gamma && beta && dx && workspace && dgamma && dbeta, "tedm_gn_silu_bwd: null pointer");
  TEDM_CHECK_ARG(batch > 0 && batch <= 65535 && hw > 0 && gn_parts > 0, "tedm_gn_silu_bwd: bad sizes");
  TEDM_CHECK_ARG(!dscale_shift || scale_shift, "tedm_gn_silu_bwd: dscale_shift without scale_shift");
  const int cvec = channels / 8;
  TEDM_UNSUPPORTED(channels % 8 != 0 || channels > GNB_MAX_C || 256 % cvec != 0 || groups <= 0 || groups > 32 ||
                       channels % groups != 0,
                   "tedm_gn_silu_bwd: channels=%d groups=%d unsupported", channels, groups);
  cudaStream_t s = (cudaStream_t)stream;
  const long long nvec = (long long)hw * cvec;
  // one full wave of CTAs (the same grid serves both passes: size it for the pass with fewer resident CTAs)
  static int capacity = 0;
  if (capacity == 0) {
    const int c1 = resident_ctas(gn_bwd_reduce_kernel, 256, 3 * 256 * 8 * sizeof(float));
    const int c2 = resident_ctas(gn_bwd_apply_kernel, 256, 0);
    capacity = c1 < c2 ? c1 : c2;
  }
  long long per_img = capacity / batch;
  if (per_img < 1) per_img = 1;
  long long vec_per_cta = (nvec + per_img - 1) / per_img;
  if (vec_per_cta < 2048) vec_per_cta = 2048;
  vec_per_cta = (vec_per_cta + 255) / 256 * 256;
  const int gx = (int)((nvec + vec_per_cta - 1) / vec_per_cta);
  TEDM_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * 3 * (size_t)batch * channels, s));
  gn_bwd_reduce_kernel<<<dim3(gx, batch), 256, 3 * 256 * 8 * sizeof(float), s>>>(
      (const bf16*)x, (const bf16*)dy, gn_partial, gn_parts, gamma, beta, scale_shift, ss_stride, ss_offset, workspace,
      (bf16*)dx, hw, channels, groups, eps, (int)vec_per_cta);
  TEDM_LAUNCH_CHECK();
  gn_bwd_apply_kernel<<<dim3(gx, batch), 256, 0, s>>>((const bf16*)x, gn_partial, gn_parts, gamma, beta,
                                                      scale_shift, ss_stride, ss_offset, workspace, (bf16*)dx, dgamma, dbeta,
                                                      dbias, dscale_shift, hw, channels, groups, eps, (int)vec_per_cta);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_layernorm_bwd(const void* x, const float* g, const void* dy, const void* add, void* dx, float* dg,
                                  int64_t npix, int channels, float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && g && dy && dx && dg && npix > 0, "tedm_layernorm_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  switch (channels) {
    case 64: return launch_layernorm_bwd<8, 1>(x, g, dy, add, dx, dg, npix, eps, s);
    case 128: return launch_layernorm_bwd<16, 1>(x, g, dy, add, dx, dg, npix, eps, s);
    case 256: return launch_layernorm_bwd<32, 1>(x, g, dy, add, dx, dg, npix, eps, s);
    case 512: return launch_layernorm_bwd<32, 2>(x, g, dy, add, dx, dg, npix, eps, s);
    case 1024: return launch_layernorm_bwd<32, 4>(x, g, dy, add, dx, dg, npix, eps, s);
    default: return tedm_set_error(TEDM_ERR_UNSUPPORTED, "tedm_layernorm_bwd: channels=%d unsupported", channels);
  }
}

extern "C" int tedm_bias_grad(const void* dy, float* dbias, int64_t npix, int channels, tedm_stream_t stream) {
  TEDM_CHECK_ARG(dy && dbias && npix > 0 && channels > 0, "tedm_bias_grad: bad arguments");
  const int cvec = channels / 8;
  TEDM_UNSUPPORTED(channels % 8 != 0 || cvec > 256, "tedm_bias_grad: channels=%d unsupported", channels);
  const int nthreads = (256 / cvec) * cvec;
  const int rstep = nthreads / cvec;
  long long rows_per_cta = (npix + (long long)tedm_num_sms() * 2 - 1) / ((long long)tedm_num_sms() * 2);
  if (rows_per_cta < 8LL * rstep) rows_per_cta = 8LL * rstep;
  const int grid = (int)((npix + rows_per_cta - 1) / rows_per_cta);
  colsum_kernel<<<grid, nthreads, nthreads * 8 * sizeof(float), (cudaStream_t)stream>>>((const bf16*)dy, dbias, npix, channels,
                                                                                        (int)rows_per_cta);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_add_bf16(const void* a, const void* b, void* out, int64_t n, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && b && out && n > 0 && n % 8 == 0, "tedm_add_bf16: bad arguments (n must be a multiple of 8)");
  add_bf16_kernel<<<grid_for(n / 8, 256, 16), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, (const uint4*)b, (uint4*)out,
                                                                              n / 8);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_final_conv1x1_bwd(const void* h, const float* weight, const float* dout, void* dh, float* dweight,
                                      float* dbias, int batch, int hw, int channels, int out_dim, tedm_stream_t stream) {
  TEDM_CHECK_ARG(h && weight && dout && dh && dweight && dbias && batch > 0 && hw > 0, "tedm_final_conv1x1_bwd: bad arguments");
  const int cvec = channels / 8;
  TEDM_UNSUPPORTED(channels % 8 != 0 || cvec > 256 || 256 % cvec != 0 || out_dim < 1 || out_dim > FC_MAX_OD,
                   "tedm_final_conv1x1_bwd: channels=%d out_dim=%d unsupported", channels, out_dim);
  const long long npix = (long long)batch * hw;
  const int rstep = 256 / cvec;
  long long rows_per_cta = (npix + (long long)tedm_num_sms() * 2 - 1) / ((long long)tedm_num_sms() * 2);
  if (rows_per_cta < 8LL * rstep) rows_per_cta = 8LL * rstep;
  const int grid = (int)((npix + rows_per_cta - 1) / rows_per_cta);
  final_conv_bwd_kernel<<<grid, 256, 256 * 8 * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)h, weight, dout, (bf16*)dh, dweight, dbias, hw, channels, out_dim, npix, (int)rows_per_cta);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// Tensor-core form for the reference's stem (1 -> 64 channels): D[co][tap] += sum_px dy[px][co] * x[px + tap], an mma.sync
// GEMM with M = 64 output channels, N = 56 (49 taps, a column of ones that yields the bias gradient, padding) and
// K = the pixels of one image row.  A = dy^T through ldmatrix.trans from the staged bf16 row; B = the input gathered from
// a 7-row staging that holds x as bf16 (hi, lo) halves (two MMAs per tile: the fp32 input keeps ~16 mantissa bits).
#define SWM_XP 272
__global__ void __launch_bounds__(256) stem_wgrad_mma_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                             float* __restrict__ dw, float* __restrict__ db, int batch, int H,
                                                             int W) {
  __shared__ __align__(16) uint8_t sdy[256 * 144];                 // [px][64 co] bf16, pitch 144 B
  __shared__ __align__(16) bf16 xs[(2 * 7 + 2) * SWM_XP];          // hi rows, lo rows, ones row, zero row
  float (*red)[32][28] = reinterpret_cast<float (*)[32][28]>(sdy);   // [4][32][28], reuses the dy staging after the loop
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const int mt = warp & 3, half = warp >> 2;
  for (int i = tid; i < 2 * SWM_XP; i += 256) xs[14 * SWM_XP + i] = __float2bfloat16_rn(i < SWM_XP ? 1.0f : 0.0f);
  int koff_hi[7], koff_lo[7];
#pragma unroll
  for (int nt = 0; nt < 7; ++nt) {
    const int tap = nt * 8 + g;
    if (tap < 49) {
      koff_hi[nt] = (tap / 7) * SWM_XP + tap % 7;
      koff_lo[nt] = koff_hi[nt] + 7 * SWM_XP;
    } else {
      koff_hi[nt] = (tap == 49 ? 14 : 15) * SWM_XP;
      koff_lo[nt] = 15 * SWM_XP;
    }
  }
  float acc[7][4];
#pragma unroll
  for (int nt = 0; nt < 7; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
  const uint32_t sdy_u = smem_u32(sdy);
  const unsigned short* xu = reinterpret_cast<const unsigned short*>(xs);
  const int ksteps = W / 16, rows = batch * H;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int b = r / H, y = r - b * H;
    __syncthreads();
    const bf16* dsrc = dy + (size_t)r * W * 64;
    for (int i = tid; i < W * 8; i += 256) {
      const int px = i >> 3, v = i & 7;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdy_u + px * 144 + v * 16), "l"(dsrc + px * 64 + v * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < 7 * (W + 6); i += 256) {
      const int ky = i / (W + 6), xi = i - ky * (W + 6);
      const int yy = y + ky - 3, xx = xi - 3;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((size_t)b * H + yy) * W + xx) : 0.0f;
      const bf16 hi = __float2bfloat16_rn(v);
      xs[ky * SWM_XP + xi] = hi;
      xs[(7 + ky) * SWM_XP + xi] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int ks = half; ks < ksteps; ks += 2) {
      uint32_t a[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                   : "r"(sdy_u + (ks * 16 + (j >> 1) * 8 + rr) * 144 + (mt * 16 + (j & 1) * 8) * 2));
      const int p0 = ks * 16 + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        const unsigned short* ph = xu + koff_hi[nt] + p0;
        const unsigned short* pl = xu + koff_lo[nt] + p0;
        const uint32_t h0 = (uint32_t)ph[0] | ((uint32_t)ph[1] << 16), h1 = (uint32_t)ph[8] | ((uint32_t)ph[9] << 16);
        const uint32_t l0 = (uint32_t)pl[0] | ((uint32_t)pl[1] << 16), l1 = (uint32_t)pl[8] | ((uint32_t)pl[9] << 16);
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(h0), "r"(h1));
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(l0), "r"(l1));
      }
    }
  }
  // merge the two pixel halves, then one atomic per (co, tap) and CTA
  __syncthreads();
  if (half == 1) {
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[mt][lane][nt * 4 + e] = acc[nt][e];
  }
  __syncthreads();
  if (half == 0) {
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = acc[nt][e] + red[mt][lane][nt * 4 + e];
        const int co = mt * 16 + g + (e >> 1) * 8, tap = nt * 8 + 2 * t4 + (e & 1);
        if (tap < 49) atomicAdd(dw + co * 49 + tap, v);
        else if (tap == 49 && db) atomicAdd(db + co, v);
      }
  }
}

extern "C" int tedm_stem_conv7x7_wgrad(const float* x, const void* dy, float* dweight, float* dbias, int batch, int cin,
                                       int height, int width, int cout, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && dy && dweight && batch > 0 && cin > 0 && height > 0 && width > 0, "tedm_stem_conv7x7_wgrad: bad arguments");
  TEDM_UNSUPPORTED(cout % 8 != 0, "tedm_stem_conv7x7_wgrad: cout=%d must be a multiple of 8", cout);
  if (cin == 1 && cout == 64 && width % 32 == 0 && width <= 256) {   // the reference's stem: tensor-core path
    int g = tedm_num_sms() * 2;
    if (g > batch * height) g = batch * height;
    stem_wgrad_mma_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x, (const bf16*)dy, dweight, dbias, batch, height, width);
    TEDM_LAUNCH_CHECK();
    return TEDM_OK;
  }
  const long long ntiles = (long long)batch * height * ((width + STW_PX - 1) / STW_PX);
  long long grid = (long long)tedm_num_sms() * 2;
  if (grid > ntiles) grid = ntiles;
  stem_wgrad_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x, (const bf16*)dy, dweight, dbias, batch, cin, height, width,
                                                                 cout);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_time_embed_train(const int64_t* t, const float* freq, const float* w1, const float* b1, const float* w2,
                                     const float* b2, float* emb, float* hidden_pre, float* temb, int batch, int dim,
                                     int tdim, tedm_stream_t stream) {
  TEDM_CHECK_ARG(t && freq && w1 && b1 && w2 && b2 && emb && hidden_pre && temb, "tedm_time_embed_train: null pointer");
  TEDM_CHECK_ARG(batch > 0 && dim > 0 && dim % 2 == 0 && tdim > 0 && (dim + tdim) * 4 <= 48 * 1024,
                 "tedm_time_embed_train: bad sizes batch=%d dim=%d tdim=%d", batch, dim, tdim);
  time_embed_train_kernel<<<batch, 256, (dim + tdim) * sizeof(float), (cudaStream_t)stream>>>(t, freq, w1, b1, w2, b2, emb,
                                                                                               hidden_pre, temb, dim, tdim);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_linear_bwd(const float* dy_raw, const float* y_pre, int act_y, const float* x, int act_x, const float* w,
                               float* dw, float* db, float* dx_raw, int batch, int n_out, int n_in, tedm_stream_t stream) {
  TEDM_CHECK_ARG(dy_raw && x && w && dw && batch > 0 && n_out > 0 && n_in > 0, "tedm_linear_bwd: bad arguments");
  TEDM_CHECK_ARG(act_y >= 0 && act_y <= 2 && act_x >= 0 && act_x <= 2 && (act_y == 0 || y_pre), "tedm_linear_bwd: bad activation");
  TEDM_UNSUPPORTED(n_in > 1024, "tedm_linear_bwd: n_in=%d > 1024", n_in);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)LB_BCHUNK * (n_in + 8) * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    TEDM_CUDA(cudaFuncSetAttribute(linear_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  linear_bwd_w_kernel<<<ceil_div(n_out, 8), 256, smem, s>>>(dy_raw, y_pre, act_y, x, act_x, dw, db, batch, n_out, n_in);
  TEDM_LAUNCH_CHECK();
  if (dx_raw) {
    TEDM_CHECK_ARG(batch <= 65535, "tedm_linear_bwd: batch too large");
    TEDM_CUDA(cudaMemsetAsync(dx_raw, 0, sizeof(float) * (size_t)batch * n_in, s));
    int jsplit = ceil_div(n_out, 512);
    if (jsplit > 64) jsplit = 64;
    const int jchunk = ceil_div(ceil_div(n_out, jsplit), 512) * 512;
    jsplit = ceil_div(n_out, jchunk);
    linear_bwd_x_kernel<<<dim3(ceil_div(n_in, 256), batch, jsplit), 256, 0, s>>>(dy_raw, y_pre, act_y, w, dx_raw, n_out, n_in,
                                                                               jchunk);
    TEDM_LAUNCH_CHECK();
  }
  return TEDM_OK;
}

extern "C" int tedm_weight_to_dgrad(const float* w_oihw, void* w_dgrad, int cout, int cin, int mode, tedm_stream_t stream) {
  TEDM_CHECK_ARG(w_oihw && w_dgrad && cout > 0 && cin > 0 && mode >= 0 && mode <= 3, "tedm_weight_to_dgrad: bad arguments");
  const long long total = (long long)cout * cin * (mode == 0 ? 1 : mode == 1 ? 9 : 16);
  weight_to_dgrad_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(w_oihw, (bf16*)w_dgrad, cout, cin, mode);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_wgrad_to_oihw(const float* dw, float* grad_oihw, int cout, int cin, int mode, tedm_stream_t stream) {
  TEDM_CHECK_ARG(dw && grad_oihw && cout > 0 && cin > 0 && mode >= 0 && mode <= 3, "tedm_wgrad_to_oihw: bad arguments");
  const long long total = (long long)cout * cin * (mode == 0 ? 1 : (mode == 2 ? 16 : 9));
  wgrad_to_oihw_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(dw, grad_oihw, cout, cin, mode);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, int* step_counter,
                              float grad_scale, tedm_stream_t stream) {
  TEDM_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && n % 4 == 0 && (step >= 1 || step_counter),
                 "tedm_adam_step: bad arguments (n must be a multiple of 4)");
  cudaStream_t s = (cudaStream_t)stream;
  if (step_counter) {
    adam_tick_kernel<<<1, 1, 0, s>>>(step_counter);
    TEDM_LAUNCH_CHECK();
  }
  adam_kernel<<<grid_for(n / 4, 256, 16), 256, 0, s>>>((float4*)param, (const float4*)grad, (float4*)exp_avg,
                                                       (float4*)exp_avg_sq, n / 4, lr, beta1, beta2, eps, weight_decay,
                                                       step_counter, step, grad_scale);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
