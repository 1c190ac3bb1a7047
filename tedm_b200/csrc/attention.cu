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
// The two contractions are 32x32xn / nx32x32 per head: far too small in M for tcgen05 (M >= 64), so
// they run on the warp-level bf16 tensor-core path (mma.sync m16n8k16) with fp32 accumulation, and
// the kernels are bound by streaming qkv once from HBM:
//   K1 ctx    : P = exp(k - m) built in registers from ldmatrix fragments with an ONLINE running column maximum m
//               (no max pre-pass), ctx += P^T v on tensor cores, s += sum P; one partial (m, s, ctx) per chunk
//                                                                                        (reads k, v once)
//   K2 combine: merge the partials with exp(m_c - max_c m_c), ctx * scale / (s * n) -> bf16 ctx^T   (tiny)
//   K3 out    : per 64-pixel tile: softmax over each head's 32 q channels in fragment layout,
//               out = softmax(q) ctx on tensor cores, coalesced bf16 store                (reads q)
// ------------------------------------------------------------------------------------------
#define LA_HEADS 4
#define LA_C (LA_HEADS * DH)          // 128 channels per q / k / v
#define LA_SUB 64                      // pixels per shared-memory tile
#define LA_CHUNK 1024                  // pixels per K0/K1 CTA
#define LA_KV_PITCH 528                // bytes per smem row: 512 (k|v) + 16 pad (ldmatrix conflict-free)
#define LA_Q_PITCH 272                 // 256 + 16
#define LA_CT_PITCH 80                 // 64 + 16
#define LA_PART (LA_C + LA_C * DH)     // per-chunk partial: s[128], ctx[128][32]

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// K1: per (image, chunk): s[c] = sum_px exp(k - M), ctx[h][d][e] = sum_px exp(k[px][h,d] - M) v[px][h,e]
__global__ void __launch_bounds__(256) linattn_ctx_kernel(const bf16* __restrict__ qkv, float* __restrict__ pmax,
                                                          float* __restrict__ part, int n, int nchunks) {
  extern __shared__ __align__(16) uint8_t la_smem[];
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = chunk * LA_CHUNK, p1 = min(n, p0 + LA_CHUNK);
  const int nsub = (p1 - p0 + LA_SUB - 1) / LA_SUB;
  const uint32_t tile0 = smem_u32(la_smem);
  constexpr int TILE_BYTES = LA_SUB * LA_KV_PITCH;

  auto load_sub = [&](int sub, int buf) {
    const int base_px = p0 + sub * LA_SUB;
    for (int i = tid; i < LA_SUB * 32; i += 256) {
      const int row = i >> 5, c16 = i & 31;
      const int px = base_px + row;
      const uint32_t dst = tile0 + buf * TILE_BYTES + row * LA_KV_PITCH + c16 * 16;
      if (px < p1) {
        cp_async16(dst, reinterpret_cast<const uint8_t*>(qkv + ((size_t)b * n + px) * (3 * LA_C) + LA_C) + c16 * 16);
      } else {  // padding rows: k = -inf (P = 0), v = 0
        const uint32_t fill = c16 < 16 ? 0xFF80FF80u : 0u;
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(fill) : "memory");
      }
    }
  };

  const int h = warp & 3, slice = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;
  float acc[2][4][4];
  float ssum[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    ssum[a][0] = ssum[a][1] = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][c][e] = 0.0f;
  }

  load_sub(0, 0);
  cp_async_commit();
  float Mrow[2][2];   // running column maxima (online softmax over the pixels): k is read ONCE, there is no max pre-pass
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) Mrow[mt][0] = Mrow[mt][1] = -INFINITY;

  for (int sub = 0; sub < nsub; ++sub) {
    const int buf = sub & 1;
    if (sub + 1 < nsub) {
      load_sub(sub + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t tile = tile0 + buf * TILE_BYTES;
    const int j = lane >> 3, rr = lane & 7;
    // the slice's 32 pixels of k^T as A fragments, their column maxima, the rescale of what was accumulated so far
    uint32_t a[2][2][4];
    float tmax[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int px = slice * 32 + ks * 16 + (j >> 1) * 8 + rr, dcol = h * DH + mt * 16 + (j & 1) * 8;
        ldsm_x4_trans(tile + px * LA_KV_PITCH + dcol * 2, a[ks][mt]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 kv = unpack_bf16x2(a[ks][mt][r]);
          tmax[mt][r & 1] = fmaxf(tmax[mt][r & 1], fmaxf(kv.x, kv.y));
        }
      }
    float corr[2][2];
    bool moved = false;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float t = tmax[mt][r];
        t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
        t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 2));
        const float mn = fmaxf(Mrow[mt][r], t);
        corr[mt][r] = mn == -INFINITY ? 1.0f : __expf(Mrow[mt][r] - mn);   // an all-padding slice leaves -inf in place
        moved = moved || corr[mt][r] != 1.0f;
        Mrow[mt][r] = mn;
        ssum[mt][r] *= corr[mt][r];
      }
    if (__any_sync(0xffffffffu, moved)) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          acc[mt][nt][0] *= corr[mt][0];
          acc[mt][nt][1] *= corr[mt][0];
          acc[mt][nt][2] *= corr[mt][1];
          acc[mt][nt][3] *= corr[mt][1];
        }
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int k0 = slice * 32 + ks * 16;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 kv = unpack_bf16x2(a[ks][mt][r]);
          const float mm = Mrow[mt][r & 1] == -INFINITY ? 0.0f : Mrow[mt][r & 1];   // all-padding slice: exp(-inf) = 0
          const float e0 = __expf(kv.x - mm), e1 = __expf(kv.y - mm);
          ssum[mt][r & 1] += e0 + e1;
          a[ks][mt][r] = pack_bf16x2(e0, e1);
        }
      }
      uint32_t bfr[2][4];
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        const int px = k0 + (j & 1) * 8 + rr, ecol = h * DH + (2 * np + (j >> 1)) * 8;
        ldsm_x4_trans(tile + px * LA_KV_PITCH + 256 + ecol * 2, bfr[np]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_bf16_16816(acc[mt][nt], a[ks][mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
    }
    __syncthreads();  // the buffer is refilled by the next iteration's prefetch
  }
  // quad-reduce the row sums
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      ssum[mt][r] += __shfl_xor_sync(0xffffffffu, ssum[mt][r], 1);
      ssum[mt][r] += __shfl_xor_sync(0xffffffffu, ssum[mt][r], 2);
    }
  // merge the two pixel slices (their running maxima differ) through shared memory, then write the chunk partial:
  // pmax[chunk][c] = the chunk's column maximum, part[chunk] = (s, ctx) relative to THAT maximum
  float* red = reinterpret_cast<float*>(la_smem);  // [4 heads][32 lanes][40]
  if (slice == 1) {
    float* r = red + (h * 32 + lane) * 40;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) r[(mt * 4 + nt) * 4 + e] = acc[mt][nt][e];
      r[32 + mt * 2] = ssum[mt][0];
      r[32 + mt * 2 + 1] = ssum[mt][1];
      r[36 + mt * 2] = Mrow[mt][0];
      r[36 + mt * 2 + 1] = Mrow[mt][1];
    }
  }
  __syncthreads();
  if (slice == 0) {
    const float* r = red + (h * 32 + lane) * 40;
    float* dst = part + ((size_t)b * nchunks + chunk) * LA_PART;
    float* mdst = pmax + ((size_t)b * nchunks + chunk) * LA_C;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float m1 = r[36 + mt * 2 + q];
        const float M = fmaxf(Mrow[mt][q], m1);
        const float f0 = Mrow[mt][q] == -INFINITY ? 0.0f : __expf(Mrow[mt][q] - M);
        const float f1 = m1 == -INFINITY ? 0.0f : __expf(m1 - M);
        const int d = h * DH + mt * 16 + g + q * 8;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float* rr2 = r + (mt * 4 + nt) * 4 + 2 * q;
          *reinterpret_cast<float2*>(dst + LA_C + (size_t)d * DH + nt * 8 + 2 * t4) =
              make_float2(acc[mt][nt][2 * q] * f0 + rr2[0] * f1, acc[mt][nt][2 * q + 1] * f0 + rr2[1] * f1);
        }
        if (t4 == 0) {
          mdst[d] = M;
          dst[d] = ssum[mt][q] * f0 + r[32 + mt * 2 + q] * f1;
        }
      }
  }
}

// K2: ctxT[b][h][e][d] = bf16( scale * sum_c ctx_c[h][d][e] / (n * sum_c s_c[h][d]) )
__global__ void __launch_bounds__(256) linattn_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ part,
                                                              bf16* __restrict__ ctxT, int n, int nchunks, float scale) {
  __shared__ float sS[DH], sMx[DH];
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;   // one CTA per (image, head): 4x the parallelism of per-image
  const float* p = part + (size_t)b * nchunks * LA_PART;
  const float* pm = pmax + (size_t)b * nchunks * LA_C;
  if (tid < DH) {
    const int hd = h * DH + tid;
    float M = -INFINITY;
    for (int c = 0; c < nchunks; ++c) M = fmaxf(M, pm[(size_t)c * LA_C + hd]);
    float s = 0.0f;
    for (int c = 0; c < nchunks; ++c) s += __expf(pm[(size_t)c * LA_C + hd] - M) * p[(size_t)c * LA_PART + hd];
    sMx[tid] = M;
    sS[tid] = s;
  }
  __syncthreads();
  for (int li = tid; li < DH * DH; li += 256) {
    const int d = li >> 5, e = li & 31, hd = h * DH + d, idx = hd * DH + e;
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c) acc += __expf(pm[(size_t)c * LA_C + hd] - sMx[d]) * p[(size_t)c * LA_PART + LA_C + idx];
    ctxT[(((size_t)b * LA_HEADS + h) * DH + e) * DH + d] = __float2bfloat16_rn(acc * scale / (sS[d] * (float)n));
  }
}

// K3: out[px][h*32+e] = sum_d softmax_d(q[px][h*32+:])[d] * ctx[h][d][e]
__global__ void __launch_bounds__(256) linattn_out_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ ctxT,
                                                          bf16* __restrict__ out, int n) {
  __shared__ __align__(16) uint8_t qtile[LA_SUB * LA_Q_PITCH];
  __shared__ __align__(16) uint8_t otile[LA_SUB * LA_Q_PITCH];
  __shared__ __align__(16) uint8_t ctile[LA_HEADS * DH * LA_CT_PITCH];
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = blockIdx.x * LA_SUB;
  for (int i = tid; i < LA_HEADS * DH * 4; i += 256) {  // 128 rows x 4 chunks of 16 B
    const int row = i >> 2, c16 = i & 3;
    *reinterpret_cast<uint4*>(ctile + row * LA_CT_PITCH + c16 * 16) =
        __ldg(reinterpret_cast<const uint4*>(ctxT + ((size_t)b * LA_HEADS * DH + row) * DH) + c16);
  }
  for (int i = tid; i < LA_SUB * 16; i += 256) {
    const int row = i >> 4, c16 = i & 15;
    const int px = p0 + row;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (px < n) v = ldg_stream(qkv + ((size_t)b * n + px) * (3 * LA_C) + c16 * 8);
    *reinterpret_cast<uint4*>(qtile + row * LA_Q_PITCH + c16 * 16) = v;
  }
  __syncthreads();
  const int h = warp & 3, slice = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const uint32_t q_s = smem_u32(qtile), c_s = smem_u32(ctile);
  uint32_t bfr[2][2][4];  // [k-step][n-tile pair][regs]
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int e = (2 * np + (j >> 1)) * 8 + rr, dcol = ks * 16 + (j & 1) * 8;
      ldsm_x4(c_s + (h * DH + e) * LA_CT_PITCH + dcol * 2, bfr[ks][np]);
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int m0 = slice * 32 + mt * 16;
    uint32_t a[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int px = m0 + (j & 1) * 8 + rr, dcol = h * DH + ks * 16 + (j >> 1) * 8;
      ldsm_x4(q_s + px * LA_Q_PITCH + dcol * 2, a[ks]);
    }
    // softmax over the 32 channels of each row: rows g (regs 0,2) and g+8 (regs 1,3)
    float x[2][8];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 f = unpack_bf16x2(a[ks][r]);
        x[r & 1][ks * 4 + (r >> 1) * 2] = f.x;
        x[r & 1][ks * 4 + (r >> 1) * 2 + 1] = f.y;
      }
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float m = x[r][0];
#pragma unroll
      for (int i = 1; i < 8; ++i) m = fmaxf(m, x[r][i]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[r][i] = __expf(x[r][i] - m);
        s += x[r][i];
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      inv[r] = 1.0f / s;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 4; ++r)
        a[ks][r] = pack_bf16x2(x[r & 1][ks * 4 + (r >> 1) * 2], x[r & 1][ks * 4 + (r >> 1) * 2 + 1]);
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) mma_bf16_16816(acc[nt], a[ks], bfr[ks][nt >> 1][(nt & 1) * 2], bfr[ks][nt >> 1][(nt & 1) * 2 + 1]);
      const int col = h * DH + nt * 8 + 2 * t4;
      *reinterpret_cast<uint32_t*>(otile + (m0 + g) * LA_Q_PITCH + col * 2) = pack_bf16x2(acc[nt][0] * inv[0], acc[nt][1] * inv[0]);
      *reinterpret_cast<uint32_t*>(otile + (m0 + g + 8) * LA_Q_PITCH + col * 2) = pack_bf16x2(acc[nt][2] * inv[1], acc[nt][3] * inv[1]);
    }
  }
  __syncthreads();
  for (int i = tid; i < LA_SUB * 16; i += 256) {
    const int row = i >> 4, c16 = i & 15;
    const int px = p0 + row;
    if (px < n)
      *reinterpret_cast<uint4*>(out + ((size_t)b * n + px) * LA_C + c16 * 8) = *reinterpret_cast<const uint4*>(otile + row * LA_Q_PITCH + c16 * 16);
  }
}

extern "C" int64_t tedm_linear_attention_workspace(int batch, int n, int heads, int dim_head) {
  if (batch <= 0 || n <= 0 || heads != LA_HEADS || dim_head != DH) return -1;
  const int64_t nchunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  // fp32 elements: chunk maxima + chunk partials + bf16 ctx^T (LA_C*DH/2 floats)
  return (int64_t)batch * (nchunks * (LA_C + LA_PART) + LA_C * DH / 2);
}

extern "C" int tedm_linear_attention_fwd(const void* qkv, void* out, float* workspace, int batch, int n, int heads,
                                         int dim_head, float scale, tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && out && workspace && batch > 0 && n > 0, "tedm_linear_attention_fwd: bad arguments");
  TEDM_UNSUPPORTED(dim_head != DH || heads != LA_HEADS, "tedm_linear_attention_fwd: heads=%d dim_head=%d (only 4 x 32)", heads, dim_head);
  TEDM_CHECK_ARG(batch <= 65535, "tedm_linear_attention_fwd: batch too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int nchunks = (n + LA_CHUNK - 1) / LA_CHUNK;
  float* pmax = workspace;
  float* part = pmax + (size_t)batch * nchunks * LA_C;
  bf16* ctxT = reinterpret_cast<bf16*>(part + (size_t)batch * nchunks * LA_PART);
  const int smem = 2 * LA_SUB * LA_KV_PITCH;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(linattn_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  linattn_ctx_kernel<<<dim3(nchunks, batch), 256, smem, s>>>((const bf16*)qkv, pmax, part, n, nchunks);
  TEDM_LAUNCH_CHECK();
  linattn_combine_kernel<<<dim3(batch, LA_HEADS), 256, 0, s>>>(pmax, part, ctxT, n, nchunks, scale);
  TEDM_LAUNCH_CHECK();
  linattn_out_kernel<<<dim3((n + LA_SUB - 1) / LA_SUB, batch), 256, 0, s>>>((const bf16*)qkv, ctxT, (bf16*)out, n);
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

// Tensor-core (mma.sync) flash form of the same attention, any n: CTA = (64-query tile, head, image), 4 warps x 16
// queries.  The L2 norms over tokens are recomputed per CTA from L2-resident qkv (n x 64 values), q-hat * scale and k-hat
// are written to shared memory as bf16 (|sim| <= scale, so bf16 operands cost ~1e-3 relative on the logits' scale), the
// softmax over keys is online, P goes register-to-register into the P V product.
#define FA_PITCH 80
__global__ void __launch_bounds__(128) attention_flash_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int n,
                                                              int heads, float scale) {
  __shared__ __align__(16) uint8_t sq[64 * FA_PITCH], sk[64 * FA_PITCH], sv[64 * FA_PITCH];
  __shared__ float s_inv[64], s_red[64];
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * DH, C = heads * DH;
  const bf16* base = qkv + (size_t)b * n * C3;
  {  // column norms of this head's q (channels 0..31) and k (32..63) over all tokens
    const int c = tid & 63, part = tid >> 6;
    const int ch = c < DH ? h * DH + c : C + h * DH + (c - DH);
    float acc = 0.0f;
    for (int px = part; px < n; px += 2) {
      const float v = __bfloat162float(base[(size_t)px * C3 + ch]);
      acc = fmaf(v, v, acc);
    }
    if (part == 1) s_red[c] = acc;
    __syncthreads();
    if (part == 0) {
      const float inv = 1.0f / fmaxf(sqrtf(acc + s_red[c]), 1e-12f);
      s_inv[c] = c < DH ? inv * scale : inv;
    }
    __syncthreads();
  }
  auto load_scaled = [&](uint8_t* dst, int row0, int ch0, const float* inv) {   // 64 rows x 32 channels, optional scaling
    const int row = tid >> 1, half = tid & 1;
    const int px = row0 + row;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      uint4 u = make_uint4(0, 0, 0, 0);
      if (px < n) {
        u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)px * C3 + ch0 + half * 16 + v * 8));
        if (inv) {
          float f[8];
          unpack8(u, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] *= inv[half * 16 + v * 8 + e];
          u = pack8(f);
        }
      }
      *reinterpret_cast<uint4*>(dst + row * FA_PITCH + (half * 16 + v * 8) * 2) = u;
    }
  };
  load_scaled(sq, q0, h * DH, s_inv);
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const uint32_t sq_u = smem_u32(sq), sk_u = smem_u32(sk), sv_u = smem_u32(sv);
  float o[4][4], mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.0f, 0.0f};
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
  for (int k0 = 0; k0 < n; k0 += 64) {
    __syncthreads();                                  // previous tile consumed (first pass: q tile visible after the next sync)
    load_scaled(sk, k0, C + h * DH, s_inv + DH);
    load_scaled(sv, k0, 2 * C + h * DH, nullptr);
    __syncthreads();
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      ldsm_x4(sq_u + (warp * 16 + (j & 1) * 8 + rr) * FA_PITCH + (ks * 16 + (j >> 1) * 8) * 2, a);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        ldsm_x4(sk_u + ((2 * np + (j >> 1)) * 8 + rr) * FA_PITCH + (ks * 16 + (j & 1) * 8) * 2, bfr);
        mma_bf16_16816(sc[2 * np], a, bfr[0], bfr[1]);
        mma_bf16_16816(sc[2 * np + 1], a, bfr[2], bfr[3]);
      }
    }
    if (k0 + 64 > n) {                                // ragged last tile: keys beyond n do not exist
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + nt * 8 + 2 * t4 + (e & 1) >= n) sc[nt][e] = -INFINITY;
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float tm = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) tm = fmaxf(tm, fmaxf(sc[nt][2 * r], sc[nt][2 * r + 1]));
      tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
      tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 2));
      const float mn = fmaxf(mrun[r], tm);
      const float corr = __expf(mrun[r] - mn);
      mrun[r] = mn;
      float ls = 0.0f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float e0 = __expf(sc[nt][2 * r] - mn), e1 = __expf(sc[nt][2 * r + 1] - mn);
        ls += e0 + e1;
        pa[nt >> 1][(nt & 1) * 2 + r] = pack_bf16x2(e0, e1);
      }
      lrun[r] = fmaf(lrun[r], corr, ls);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o[nt][2 * r] *= corr;
        o[nt][2 * r + 1] *= corr;
      }
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bfr[4];
        ldsm_x4_trans(sv_u + (ks * 16 + (j & 1) * 8 + rr) * FA_PITCH + ((2 * np + (j >> 1)) * 8) * 2, bfr);
        mma_bf16_16816(o[2 * np], pa[ks], bfr[0], bfr[1]);
        mma_bf16_16816(o[2 * np + 1], pa[ks], bfr[2], bfr[3]);
      }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = lrun[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.0f / l;
    const int qi = q0 + warp * 16 + g + r * 8;
    if (qi < n) {
      bf16* op = out + ((size_t)b * n + qi) * C + h * DH + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<uint32_t*>(op + nt * 8) = pack_bf16x2(o[nt][2 * r] * inv, o[nt][2 * r + 1] * inv);
    }
  }
}

extern "C" int tedm_attention_fwd(const void* qkv, void* out, int batch, int n, int heads, int dim_head, float scale,
                                  tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && out && batch > 0 && n > 0 && heads > 0, "tedm_attention_fwd: bad arguments");
  TEDM_UNSUPPORTED(dim_head != DH, "tedm_attention_fwd: dim_head=%d (only 32)", dim_head);
  TEDM_CHECK_ARG(batch <= 65535 && heads <= 65535, "tedm_attention_fwd: batch / heads too large");
  if (n >= 64) {   // tensor-core flash form, any n
    attention_flash_kernel<<<dim3((n + 63) / 64, heads, batch), 128, 0, (cudaStream_t)stream>>>((const bf16*)qkv, (bf16*)out, n,
                                                                                            heads, scale);
    TEDM_LAUNCH_CHECK();
    return TEDM_OK;
  }
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
