// Memory-bound UNet pieces: time embedding, stem conv, GroupNorm-apply+SiLU, channel LayerNorm,
// nearest upsample, final 1x1 conv, layout / weight conversions.  Activations are NHWC bf16; every
// thread moves 16-byte vectors (8 channels) so warps read and write whole 128-byte lines.
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// time embedding                                        models/unet_model.py:76-93, 287-292
// one CTA per batch element; freq[] is the reference's exp(arange(half) * -log(1e4)/(half-1))
// table computed once on the host with the same torch ops.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_embed_kernel(const int64_t* __restrict__ t, const float* __restrict__ freq,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ temb, int dim, int tdim) {
  extern __shared__ float sm[];
  float* emb = sm;          // [dim]
  float* hid = sm + dim;    // [tdim]
  const int b = blockIdx.x, half = dim / 2;
  const float tf = (float)t[b];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = __fmul_rn(tf, freq[i]);
    emb[i] = sinf(a);
    emb[half + i] = cosf(a);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < tdim; j += nw) {
    float acc = 0.0f;
    for (int k = lane; k < dim; k += 32) acc = fmaf(w1[(size_t)j * dim + k], emb[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = acc + b1[j];
      hid[j] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // GELU (erf form, nn.GELU default)
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

extern "C" int tedm_time_embed(const int64_t* t, const float* freq, const float* w1, const float* b1, const float* w2,
                               const float* b2, float* temb, int batch, int dim, int tdim, tedm_stream_t stream) {
  TEDM_CHECK_ARG(t && freq && w1 && b1 && w2 && b2 && temb, "tedm_time_embed: null pointer");
  TEDM_CHECK_ARG(batch > 0 && dim > 0 && dim % 2 == 0 && tdim > 0 && (dim + tdim) * 4 <= 48 * 1024,
                 "tedm_time_embed: bad sizes batch=%d dim=%d tdim=%d", batch, dim, tdim);
  time_embed_kernel<<<batch, 256, (dim + tdim) * sizeof(float), (cudaStream_t)stream>>>(t, freq, w1, b1, w2, b2, temb,
                                                                                         dim, tdim);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// out[b][j] = sum_k W[j][k] * silu(temb[b][k]) + bias[j]          models/unet_model.py:150-152,168-171
// CTA = (32 output rows j, 32 batch rows b); lane = batch row, warp = 4 consecutive j.  Weights and activations sit
// transposed in shared memory ([k][j], [k][b]) so the inner loop is one conflict-free load of act[k][lane], one broadcast
// float4 of W[k][4 j] and four FMAs; nothing is reduced across lanes.
#define TP_BCHUNK 32
#define TP_JTILE 32
#define TP_WPITCH 36
__global__ void __launch_bounds__(256) time_proj_kernel(const float* __restrict__ temb, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out, int batch,
                                                        int tdim, int total) {
  extern __shared__ __align__(16) float tp_smem[];
  float* act = tp_smem;                        // [tdim][TP_BCHUNK]  silu(temb)^T
  float* wt = tp_smem + tdim * TP_BCHUNK;      // [tdim][TP_WPITCH]  W^T tile
  const int b0 = blockIdx.y * TP_BCHUNK, j0 = blockIdx.x * TP_JTILE;
  const int nb = min(TP_BCHUNK, batch - b0), nj = min(TP_JTILE, total - j0);
  for (int i = threadIdx.x; i < TP_BCHUNK * tdim; i += blockDim.x) {
    const int bb = i / tdim, k = i - bb * tdim;
    float v = 0.0f;
    if (bb < nb) {
      v = temb[(size_t)(b0 + bb) * tdim + k];
      v = v / (1.0f + expf(-v));
    }
    act[k * TP_BCHUNK + bb] = v;
  }
  for (int i = threadIdx.x; i < TP_JTILE * tdim; i += blockDim.x) {
    const int jj = i / tdim, k = i - jj * tdim;
    wt[k * TP_WPITCH + jj] = jj < nj ? w[(size_t)(j0 + jj) * tdim + k] : 0.0f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 8
  for (int k = 0; k < tdim; ++k) {
    const float a = act[k * TP_BCHUNK + lane];
    const float4 wv = *reinterpret_cast<const float4*>(wt + k * TP_WPITCH + warp * 4);
    acc[0] = fmaf(wv.x, a, acc[0]);
    acc[1] = fmaf(wv.y, a, acc[1]);
    acc[2] = fmaf(wv.z, a, acc[2]);
    acc[3] = fmaf(wv.w, a, acc[3]);
  }
  if (lane < nb) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + warp * 4 + i;
      if (j < total) out[(size_t)(b0 + lane) * total + j] = acc[i] + bias[j];
    }
  }
}

extern "C" int tedm_time_proj(const float* temb, const float* w_cat, const float* b_cat, float* out, int batch, int tdim,
                              int total, tedm_stream_t stream) {
  TEDM_CHECK_ARG(temb && w_cat && b_cat && out, "tedm_time_proj: null pointer");
  TEDM_CHECK_ARG(batch > 0 && tdim > 0 && total > 0 && tdim <= 384, "tedm_time_proj: bad sizes batch=%d tdim=%d total=%d",
                 batch, tdim, total);
  const int smem = tdim * (TP_BCHUNK + TP_WPITCH) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(time_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * (TP_BCHUNK + TP_WPITCH) * 4));
    configured = true;
  }
  dim3 grid(ceil_div(total, TP_JTILE), ceil_div(batch, TP_BCHUNK));
  time_proj_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(temb, w_cat, b_cat, out, batch, tdim, total);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// stem: 7x7 pad-3 conv, fp32 NCHW in -> NHWC bf16 out            models/unet_model.py:267,334
// thread = (pixel, 8 output channels); weights transposed into smem as [tap][cout].
// ------------------------------------------------------------------------------------------
// thread = (strip of 4 consecutive pixels, 8 output channels): every weight vector fetched from shared
// memory feeds 32 FMAs and every input row segment (10 values) is reused by 7 taps x 4 pixels.
#define STEM_PX 4
__global__ void __launch_bounds__(256) stem_conv7x7_kernel(const float* __restrict__ x, const float* __restrict__ weight,
                                                           const float* __restrict__ bias, bf16* __restrict__ out,
                                                           int batch, int cin, int H, int W, int cout) {
  extern __shared__ float wsm[];  // [cin*49][cout]
  const int ntap = cin * 49;
  for (int i = threadIdx.x; i < ntap * cout; i += blockDim.x) {
    const int co = i / ntap, tp = i % ntap;  // weight is [cout][cin][7][7] -> tap index = ci*49 + ky*7 + kx
    wsm[tp * cout + co] = weight[i];
  }
  __syncthreads();
  const int chunks = cout >> 3;
  const int strips_w = (W + STEM_PX - 1) / STEM_PX;
  const long long total = (long long)batch * H * strips_w * chunks;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % chunks);
    long long rest = idx / chunks;
    const int sx = (int)(rest % strips_w);
    rest /= strips_w;
    const int py = (int)(rest % H), b = (int)(rest / H);
    const int px0 = sx * STEM_PX;
    float acc[STEM_PX][8];
#pragma unroll
    for (int p = 0; p < STEM_PX; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = bias ? bias[ch * 8 + j] : 0.0f;
    for (int ci = 0; ci < cin; ++ci) {
      const float* xp = x + ((size_t)b * cin + ci) * H * W;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = py + ky - 3;
        if (yy < 0 || yy >= H) continue;
        float in[STEM_PX + 6];
#pragma unroll
        for (int i = 0; i < STEM_PX + 6; ++i) {
          const int xx = px0 + i - 3;
          in[i] = (xx >= 0 && xx < W) ? __ldg(xp + (size_t)yy * W + xx) : 0.0f;
        }
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(wsm + (ci * 49 + ky * 7 + kx) * cout + ch * 8);
          const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
          for (int p = 0; p < STEM_PX; ++p) {
            const float v = in[p + kx];
            acc[p][0] = fmaf(v, w0.x, acc[p][0]); acc[p][1] = fmaf(v, w0.y, acc[p][1]);
            acc[p][2] = fmaf(v, w0.z, acc[p][2]); acc[p][3] = fmaf(v, w0.w, acc[p][3]);
            acc[p][4] = fmaf(v, w1.x, acc[p][4]); acc[p][5] = fmaf(v, w1.y, acc[p][5]);
            acc[p][6] = fmaf(v, w1.z, acc[p][6]); acc[p][7] = fmaf(v, w1.w, acc[p][7]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < STEM_PX; ++p)
      if (px0 + p < W)
        *reinterpret_cast<uint4*>(out + (((size_t)b * H + py) * W + px0 + p) * cout + ch * 8) = pack8(acc[p]);
  }
}

// Tensor-core stem for the reference's shape (1 -> 64 channels): implicit GEMM with M = 16 pixels of an image row,
// N = 64, K = 2 x 49 -> 112.  The fp32 input is split into bf16 (hi, lo) halves that occupy K slots [0, 49) and [49, 98)
// against the same bf16 weights, so the input keeps ~16 mantissa bits (the weights are bf16 like every other conv's).
// A fragments are gathered straight from a 7-row bf16 staging of the image rows through a K -> (row, tap) offset table.
#define STEM_XP 272          // staged row pitch (elements): W + 6 <= 262
#define STEM_K 112
#define STEM_WPITCH 240      // bytes per weight row (112 bf16 + 16)
__global__ void __launch_bounds__(256) stem_conv7x7_mma_kernel(const float* __restrict__ x, const float* __restrict__ weight,
                                                               const float* __restrict__ bias, bf16* __restrict__ out,
                                                               int batch, int H, int W) {
  __shared__ __align__(16) uint8_t wB[64 * STEM_WPITCH];
  __shared__ __align__(16) bf16 xs[(2 * 7 + 1) * STEM_XP];   // hi rows, lo rows, one zero row
  __shared__ __align__(8) int koff[STEM_K];
  __shared__ __align__(16) uint8_t stage[8][16 * 144];
  __shared__ float sbias[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * STEM_K; i += 256) {
    const int co = i / STEM_K, k = i % STEM_K;
    const float wv = k < 98 ? weight[co * 49 + (k < 49 ? k : k - 49)] : 0.0f;
    *reinterpret_cast<bf16*>(wB + co * STEM_WPITCH + k * 2) = __float2bfloat16_rn(wv);
  }
  for (int k = tid; k < STEM_K; k += 256) {
    const int kk = k < 49 ? k : k - 49;
    koff[k] = k < 98 ? (k < 49 ? 0 : 7 * STEM_XP) + (kk / 7) * STEM_XP + kk % 7 : 14 * STEM_XP;
  }
  for (int i = tid; i < STEM_XP; i += 256) xs[14 * STEM_XP + i] = __float2bfloat16_rn(0.0f);
  if (tid < 64) sbias[tid] = bias ? bias[tid] : 0.0f;
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const uint32_t wB_u = smem_u32(wB);
  const int rows = batch * H;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int b = r / H, y = r - b * H;
    __syncthreads();                       // the previous row's gathers are done (and, first time, the tables are written)
    for (int i = tid; i < 7 * (W + 6); i += 256) {
      const int ky = i / (W + 6), xi = i - ky * (W + 6);
      const int yy = y + ky - 3, xx = xi - 3;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((size_t)b * H + yy) * W + xx) : 0.0f;
      const bf16 hi = __float2bfloat16_rn(v);
      xs[ky * STEM_XP + xi] = hi;
      xs[(7 + ky) * STEM_XP + xi] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    __syncthreads();
    for (int mt = warp; mt < W / 16; mt += 8) {
      const int px0 = mt * 16;
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
      const unsigned short* xu = reinterpret_cast<const unsigned short*>(xs) + px0 + g;
#pragma unroll
      for (int ks = 0; ks < STEM_K / 16; ++ks) {
        uint32_t a[4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int2 o = *reinterpret_cast<const int2*>(koff + ks * 16 + half * 8 + 2 * t4);
          a[half * 2] = (uint32_t)xu[o.x] | ((uint32_t)xu[o.y] << 16);
          a[half * 2 + 1] = (uint32_t)xu[o.x + 8] | ((uint32_t)xu[o.y + 8] << 16);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t bfr[4];
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(bfr[0]), "=r"(bfr[1]), "=r"(bfr[2]), "=r"(bfr[3])
                       : "r"(wB_u + ((2 * np + (j >> 1)) * 8 + rr) * STEM_WPITCH + (ks * 16 + (j & 1) * 8) * 2));
#pragma unroll
          for (int q = 0; q < 2; ++q)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[2 * np + q][0]), "+f"(acc[2 * np + q][1]), "+f"(acc[2 * np + q][2]), "+f"(acc[2 * np + q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bfr[2 * q]), "r"(bfr[2 * q + 1]));
        }
      }
      uint8_t* st = stage[warp];
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t4;
        const float b0 = sbias[c], b1 = sbias[c + 1];
        *reinterpret_cast<uint32_t*>(st + g * 144 + c * 2) = pack_bf16x2(acc[nt][0] + b0, acc[nt][1] + b1);
        *reinterpret_cast<uint32_t*>(st + (g + 8) * 144 + c * 2) = pack_bf16x2(acc[nt][2] + b0, acc[nt][3] + b1);
      }
      __syncwarp();
      bf16* dst = out + (((size_t)b * H + y) * W + px0) * 64;
      for (int i = lane; i < 16 * 8; i += 32) {
        const int row = i >> 3, v = i & 7;
        *reinterpret_cast<uint4*>(dst + (size_t)row * 64 + v * 8) = *reinterpret_cast<const uint4*>(st + row * 144 + v * 16);
      }
    }
  }
}

extern "C" int tedm_stem_conv7x7(const float* x, const float* weight, const float* bias, void* out, int batch, int cin,
                                 int height, int width, int cout, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && weight && out, "tedm_stem_conv7x7: null pointer");
  TEDM_CHECK_ARG(batch > 0 && cin > 0 && height > 0 && width > 0 && cout > 0, "tedm_stem_conv7x7: bad sizes");
  TEDM_UNSUPPORTED(cout % 8 != 0, "tedm_stem_conv7x7: cout=%d must be a multiple of 8", cout);
  if (cin == 1 && cout == 64 && width % 16 == 0 && width <= 256) {   // the reference's stem: tensor-core path
    int grid = batch * height;
    const int cap = resident_ctas(stem_conv7x7_mma_kernel, 256, 0);
    if (grid > cap) grid = cap;
    stem_conv7x7_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, weight, bias, (bf16*)out, batch, height, width);
    TEDM_LAUNCH_CHECK();
    return TEDM_OK;
  }
  const size_t smem = (size_t)cin * 49 * cout * sizeof(float);
  TEDM_UNSUPPORTED(smem > 96 * 1024, "tedm_stem_conv7x7: cin*49*cout=%d floats do not fit in shared memory", cin * 49 * cout);
  if (smem > 48 * 1024)
    TEDM_CUDA(cudaFuncSetAttribute(stem_conv7x7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)batch * height * ((width + STEM_PX - 1) / STEM_PX) * (cout / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = resident_ctas(stem_conv7x7_kernel, 256, smem);   // one resident wave of the grid-stride loop
  if (blocks > cap) blocks = cap;
  stem_conv7x7_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(x, weight, bias, (bf16*)out, batch, cin, height,
                                                                       width, cout);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// weight layout conversions (derived caches; the state_dict stays fp32 OIHW)
// ------------------------------------------------------------------------------------------
__global__ void weight_to_krsc_kernel(const float* __restrict__ w, bf16* __restrict__ o, int cout, int cin, int khw) {
  const long long total = (long long)cout * cin * khw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // destination index i = (co*khw + tap)*cin + ci
    const int ci = (int)(i % cin);
    const int tap = (int)((i / cin) % khw);
    const int co = (int)(i / ((long long)cin * khw));
    o[i] = __float2bfloat16_rn(w[((size_t)co * cin + ci) * khw + tap]);
  }
}

extern "C" int tedm_weight_to_krsc(const float* w_oihw, void* w_krsc, int cout, int cin, int kh, int kw,
                                   tedm_stream_t stream) {
  TEDM_CHECK_ARG(w_oihw && w_krsc && cout > 0 && cin > 0 && kh > 0 && kw > 0, "tedm_weight_to_krsc: bad arguments");
  const long long total = (long long)cout * cin * kh * kw;
  weight_to_krsc_kernel<<<(int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (bf16*)w_krsc, cout, cin, kh * kw);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// nearest-x2 upsample followed by a 3x3 pad-1 conv == for each output parity (py,px) a 2x2 conv on the
// source with summed taps.  folded[par][co][a][b][ci], par = py*2+px; source offset of (a,b) is
// (a - 1 + py, b - 1 + px):  py=0: ky=0 -> a=0, ky=1,2 -> a=1;  py=1: ky=0,1 -> a=0, ky=2 -> a=1.
__global__ void fold_upsample_weight_kernel(const float* __restrict__ w, bf16* __restrict__ o, int cout, int cin) {
  const long long total = 4LL * cout * 4 * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int ab = (int)((i / cin) % 4);
    const int co = (int)((i / (4LL * cin)) % cout);
    const int par = (int)(i / (4LL * cin * cout));
    const int a = ab >> 1, b = ab & 1, py = par >> 1, px = par & 1;
    float acc = 0.0f;
    for (int ky = 0; ky < 3; ++ky) {
      const int aa = py == 0 ? (ky == 0 ? 0 : 1) : (ky == 2 ? 1 : 0);
      if (aa != a) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int bb = px == 0 ? (kx == 0 ? 0 : 1) : (kx == 2 ? 1 : 0);
        if (bb != b) continue;
        acc += w[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx];
      }
    }
    o[i] = __float2bfloat16_rn(acc);
  }
}

extern "C" int tedm_fold_upsample_weight(const float* w_oihw, void* w_folded, int cout, int cin, tedm_stream_t stream) {
  TEDM_CHECK_ARG(w_oihw && w_folded && cout > 0 && cin > 0, "tedm_fold_upsample_weight: bad arguments");
  const long long total = 16LL * cout * cin;
  fold_upsample_weight_kernel<<<(int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0,
                                (cudaStream_t)stream>>>(w_oihw, (bf16*)w_folded, cout, cin);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// GroupNorm finalise + affine + (scale+1)/shift + SiLU (+ residual)   models/unet_model.py:126-135,175
// The conv epilogue left per-(image, part, group) fp32 (sum, sum of squares); each CTA first folds
// its image's partials (double accumulation) into one per-channel affine y = x*A + B, then streams.
// ------------------------------------------------------------------------------------------
#define GN_MAX_C 1024
__global__ void __launch_bounds__(256) gn_silu_kernel(const bf16* __restrict__ x, const float* __restrict__ partial,
                                                      int parts, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ ss,
                                                      int ss_stride, int ss_offset, const bf16* __restrict__ residual,
                                                      bf16* __restrict__ out, int hw, int C, int groups, float eps,
                                                      int vec_per_cta) {
  __shared__ float sA[GN_MAX_C], sB[GN_MAX_C];
  __shared__ float s_mean[32], s_rstd[32];
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
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
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
  }
  __syncthreads();
  const int cvec = C >> 3;
  const long long nvec = (long long)hw * cvec;
  const long long v0 = (long long)blockIdx.x * vec_per_cta;
  long long v1 = v0 + vec_per_cta;
  if (v1 > nvec) v1 = nvec;
  const size_t img = (size_t)b * hw * C;
  if (256 % cvec == 0) {
    // vec_per_cta is a multiple of 256 and 256 of cvec: this thread always meets the same 8 channels,
    // so their affine coefficients live in registers for the whole stream
    const int c0 = (threadIdx.x % cvec) << 3;
    float ca[8], cb[8];     // halved: silu(z) = zh + zh * tanh(zh), zh = z / 2
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ca[j] = 0.5f * sA[c0 + j];
      cb[j] = 0.5f * sB[c0 + j];
    }
    // four independent 16-byte loads (eight with a residual) are in flight per thread before any is consumed
    constexpr int U = 4;
    long long v = v0 + threadIdx.x;
    for (; v + (U - 1) * 256 < v1; v += U * 256) {
      uint4 xv[U], rv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = ldg_stream(x + img + (v + u * 256) * 8);
      if (residual) {
#pragma unroll
        for (int u = 0; u < U; ++u) rv[u] = ldg_stream(residual + img + (v + u * 256) * 8);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[8];
        unpack8(xv[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_half(fmaf(f[j], ca[j], cb[j]));
        if (residual) {
          float r[8];
          unpack8(rv[u], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += r[j];
        }
        *reinterpret_cast<uint4*>(out + img + (v + u * 256) * 8) = pack8(f);
      }
    }
    for (; v < v1; v += 256) {
      float f[8];
      unpack8(ldg_stream(x + img + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = silu_half(fmaf(f[j], ca[j], cb[j]));
      if (residual) {
        float r[8];
        unpack8(ldg_stream(residual + img + v * 8), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      *reinterpret_cast<uint4*>(out + img + v * 8) = pack8(f);
    }
    return;
  }
  for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
    const int c0 = (int)(v % cvec) << 3;
    float f[8];
    unpack8(ldg_stream(x + img + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = silu_f(fmaf(f[j], sA[c0 + j], sB[c0 + j]));
    if (residual) {
      float r[8];
      unpack8(ldg_stream(residual + img + v * 8), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    *reinterpret_cast<uint4*>(out + img + v * 8) = pack8(f);
  }
}

// The prologue of gn_silu_kernel alone: per-(image, channel) halved affine for consumers that normalise on the fly.
__global__ void __launch_bounds__(256) gn_affine_kernel(const float* __restrict__ partial, int parts,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ ss, int ss_stride, int ss_offset,
                                                        float2* __restrict__ affine, int hw, int C, int groups, float eps) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.x;
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
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
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
    affine[(size_t)b * C + c] = make_float2(0.5f * a, 0.5f * bb);
  }
}

extern "C" int tedm_gn_affine(const float* gn_partial, int gn_parts, const float* gamma, const float* beta,
                              const float* scale_shift, int ss_stride, int ss_offset, float* affine, int batch, int hw,
                              int channels, int groups, float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(gn_partial && gamma && beta && affine, "tedm_gn_affine: null pointer");
  TEDM_CHECK_ARG(batch > 0 && hw > 0 && gn_parts > 0, "tedm_gn_affine: bad sizes");
  TEDM_UNSUPPORTED(groups <= 0 || groups > 32 || channels % groups != 0 || channels > GN_MAX_C,
                   "tedm_gn_affine: channels=%d groups=%d unsupported", channels, groups);
  gn_affine_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(gn_partial, gn_parts, gamma, beta, scale_shift, ss_stride, ss_offset,
                                                           reinterpret_cast<float2*>(affine), hw, channels, groups, eps);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_gn_silu_fwd(const void* x, const float* gn_partial, int gn_parts, const float* gamma,
                                const float* beta, const float* scale_shift, int ss_stride, int ss_offset,
                                const void* residual, void* out, int batch, int hw, int channels, int groups, float eps,
                                tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && gn_partial && gamma && beta && out, "tedm_gn_silu_fwd: null pointer");
  TEDM_CHECK_ARG(batch > 0 && batch <= 65535 && hw > 0 && gn_parts > 0, "tedm_gn_silu_fwd: bad sizes");
  TEDM_UNSUPPORTED(channels % 8 != 0 || channels > GN_MAX_C || groups <= 0 || groups > 32 || channels % groups != 0,
                   "tedm_gn_silu_fwd: channels=%d groups=%d unsupported", channels, groups);
  const long long nvec = (long long)hw * (channels / 8);
  // one full wave of CTAs over the whole batch (never a partial second wave), each CTA at least 2048 vectors (32 KB)
  // to amortise the prologue
  static int capacity = 0;
  if (capacity == 0) capacity = resident_ctas(gn_silu_kernel, 256, 0);
  long long per_img = capacity / batch;
  if (per_img < 1) per_img = 1;
  long long vec_per_cta = (nvec + per_img - 1) / per_img;
  if (vec_per_cta < 2048) vec_per_cta = 2048;
  vec_per_cta = (vec_per_cta + 255) / 256 * 256;
  const int gx = (int)((nvec + vec_per_cta - 1) / vec_per_cta);
  gn_silu_kernel<<<dim3(gx, batch), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, gn_partial, gn_parts, gamma, beta, scale_shift, ss_stride, ss_offset, (const bf16*)residual,
      (bf16*)out, hw, channels, groups, eps, (int)vec_per_cta);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// channel LayerNorm (gain only) + optional residual               models/unet_model.py:52-61, 29-36
// L lanes per pixel, each lane NV vectors of 8 channels; two-pass statistics in registers.
// ------------------------------------------------------------------------------------------
template <int L, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ g,
                                                        const bf16* __restrict__ residual, bf16* __restrict__ out,
                                                        long long npix, float eps) {
  constexpr int C = L * NV * 8;
  constexpr int PIX_PER_WARP = 32 / L;
  const int lane = threadIdx.x & 31;
  const int sub = lane % L;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float gain[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) gain[v][j] = g[(v * L + sub) * 8 + j];
  // U pixel groups per iteration: all their loads (input and residual) are issued before the first is consumed
  constexpr int U = NV == 1 ? 4 : (NV == 2 ? 2 : 1);
  const long long stride = nwarps * PIX_PER_WARP;
  for (long long p0 = warp_global * PIX_PER_WARP; p0 < npix; p0 += U * stride) {
    uint4 xv[U][NV], rv[U][NV];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * stride + lane / L;
      ok[u] = p < npix;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        xv[u][v] = ok[u] ? ldg_stream(x + (size_t)p * C + (v * L + sub) * 8) : make_uint4(0, 0, 0, 0);
        if (residual) rv[u][v] = ok[u] ? ldg_stream(residual + (size_t)p * C + (v * L + sub) * 8) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * stride + lane / L;
      float f[NV][8];
      float s = 0.0f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        unpack8(xv[u][v], f[v]);
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
      if (ok[u]) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = f[v][j] * rstd * gain[v][j];
          if (residual) {
            float r[8];
            unpack8(rv[u][v], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += r[j];
          }
          *reinterpret_cast<uint4*>(out + (size_t)p * C + (v * L + sub) * 8) = pack8(o);
        }
      }
    }
  }
}

template <int L, int NV>
static int launch_layernorm(const void* x, const float* g, const void* residual, void* out, long long npix, float eps,
                            cudaStream_t stream) {
  const long long warps = (npix + (32 / L) - 1) / (32 / L);
  long long blocks = (warps + 7) / 8;
  static int cap = 0;   // one resident wave (grid-stride loop: a partial second wave would double the time)
  if (cap == 0) cap = resident_ctas(layernorm_kernel<L, NV>, 256, 0);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  layernorm_kernel<L, NV><<<(int)blocks, 256, 0, stream>>>((const bf16*)x, g, (const bf16*)residual, (bf16*)out, npix, eps);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

extern "C" int tedm_layernorm_fwd(const void* x, const float* g, const void* residual, void* out, int64_t npix,
                                  int channels, float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && g && out && npix > 0, "tedm_layernorm_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  switch (channels) {
    case 64: return launch_layernorm<8, 1>(x, g, residual, out, npix, eps, s);
    case 128: return launch_layernorm<16, 1>(x, g, residual, out, npix, eps, s);
    case 192: return launch_layernorm<8, 3>(x, g, residual, out, npix, eps, s);
    case 256: return launch_layernorm<32, 1>(x, g, residual, out, npix, eps, s);
    case 384: return launch_layernorm<16, 3>(x, g, residual, out, npix, eps, s);
    case 512: return launch_layernorm<32, 2>(x, g, residual, out, npix, eps, s);
    case 768: return launch_layernorm<32, 3>(x, g, residual, out, npix, eps, s);
    case 1024: return launch_layernorm<32, 4>(x, g, residual, out, npix, eps, s);
    default: return tedm_set_error(TEDM_ERR_UNSUPPORTED, "tedm_layernorm_fwd: channels=%d unsupported", channels);
  }
}

// ------------------------------------------------------------------------------------------
// nearest x2 upsample, NHWC bf16                                      models/unet_model.py:42
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H, int W,
                                                         int cvec, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cvec);
    long long p = i / cvec;
    const int ox = (int)(p % (2 * W));
    p /= 2 * W;
    const int oy = (int)(p % (2 * H));
    const long long b = p / (2 * H);
    out[i] = __ldg(x + ((b * H + (oy >> 1)) * W + (ox >> 1)) * cvec + c);
  }
}

extern "C" int tedm_upsample2x(const void* x, void* out, int batch, int height, int width, int channels,
                               tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && out && batch > 0 && height > 0 && width > 0, "tedm_upsample2x: bad arguments");
  TEDM_UNSUPPORTED(channels % 8 != 0, "tedm_upsample2x: channels=%d must be a multiple of 8", channels);
  const long long total = 4LL * batch * height * width * (channels / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)tedm_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  upsample2x_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, height, width, channels / 8,
                                                                  total);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// final 1x1 conv C -> out_dim, NHWC bf16 -> NCHW fp32              models/unet_model.py:331,368
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) final_conv_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int hw,
                                                         int C, int out_dim, long long npix) {
  extern __shared__ float wsm[];  // [out_dim][C]
  for (int i = threadIdx.x; i < out_dim * C; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / hw;
    const int pi = (int)(p % hw);
    for (int od = 0; od < out_dim; ++od) {
      float acc = bias ? bias[od] : 0.0f;
      for (int c = 0; c < C; c += 8) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(x + (size_t)p * C + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(f[j], wsm[od * C + c + j], acc);
      }
      out[((size_t)b * out_dim + od) * hw + pi] = acc;
    }
  }
}

extern "C" int tedm_final_conv1x1(const void* x, const float* weight, const float* bias, float* out, int batch, int hw,
                                  int channels, int out_dim, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && weight && out && batch > 0 && hw > 0 && out_dim > 0, "tedm_final_conv1x1: bad arguments");
  TEDM_UNSUPPORTED(channels % 8 != 0 || (size_t)channels * out_dim * 4 > 48 * 1024,
                   "tedm_final_conv1x1: channels=%d out_dim=%d unsupported", channels, out_dim);
  const long long npix = (long long)batch * hw;
  long long blocks = (npix + 255) / 256;
  const long long cap = (long long)tedm_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  final_conv_kernel<<<(int)blocks, 256, (size_t)channels * out_dim * 4, (cudaStream_t)stream>>>(
      (const bf16*)x, weight, bias, out, hw, channels, out_dim, npix);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// ------------------------------------------------------------------------------------------
// layout conversion at the module boundary (32x32 smem transposes)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ out, int C, int hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && p0 + tx < hw) tile[i][tx] = x[((size_t)b * C + c0 + i) * hw + p0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (p0 + i < hw && c0 + tx < C) out[((size_t)b * hw + p0 + i) * C + c0 + tx] = __float2bfloat16_rn(tile[tx][i]);
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16* __restrict__ x, float* __restrict__ out, int C, int hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (p0 + i < hw && c0 + tx < C) tile[i][tx] = __bfloat162float(x[((size_t)b * hw + p0 + i) * C + c0 + tx]);
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && p0 + tx < hw) out[((size_t)b * C + c0 + i) * hw + p0 + tx] = tile[tx][i];
}

extern "C" int tedm_nchw_f32_to_nhwc_bf16(const float* x, void* out, int batch, int channels, int hw, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && out && batch > 0 && batch <= 65535 && channels > 0 && hw > 0, "tedm_nchw_f32_to_nhwc_bf16: bad arguments");
  nchw_to_nhwc_kernel<<<dim3(ceil_div(hw, 32), ceil_div(channels, 32), batch), 256, 0, (cudaStream_t)stream>>>(
      x, (bf16*)out, channels, hw);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
extern "C" int tedm_nhwc_bf16_to_nchw_f32(const void* x, float* out, int batch, int channels, int hw, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && out && batch > 0 && batch <= 65535 && channels > 0 && hw > 0, "tedm_nhwc_bf16_to_nchw_f32: bad arguments");
  nhwc_to_nchw_kernel<<<dim3(ceil_div(hw, 32), ceil_div(channels, 32), batch), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, out, channels, hw);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
