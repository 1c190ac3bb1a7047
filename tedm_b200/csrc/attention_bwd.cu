// Backward of the two attention cores (models/unet_model.py:197-209 and :229-240).  Same tensor
// conventions as attention.cu: qkv / dqkv are NHWC bf16 [B][n][3*heads*32] (q | k | v, head-major),
// out / dout are [B][n][heads*32].  The LinearAttention backward (1.4 GFLOP per image) runs its
// 32x32 contractions on warp-level bf16 tensor cores (mma.sync) and is bound by streaming qkv / dout;
// the mid-block attention backward (0.12 GFLOP per image, 256 tokens) runs on fp32 CUDA cores.
#include "common.cuh"

#define DH 32
#define LA_HEADS 4
#define LA_C (LA_HEADS * DH)
#define LA_CHUNK 1024
#define LA_PART (LA_C + LA_C * DH)   // forward per-chunk partial: s[128], ctx[128][32]  (attention.cu)
#define LAB_TILE 64
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
// All five contractions are 32x32 per head and run on mma.sync m16n8k16 (bf16 in, fp32 accumulate), like the
// forward (attention.cu); the kernels stream qkv / dout once per pass and are HBM-bound.
// workspace (fp32 units, per image): M[128] | S[128] | r[128] | C[4096] | dC[4096] | bf16 C, dC, dC^T (3 x 2048) |
//                                    dC partials [nchunks][4096]
#define LAB_BF_OFF (3 * LA_C + 2 * LAB_MAT)
#define LAB_WS_FIXED (LAB_BF_OFF + 3 * LAB_MAT / 2)
#define LB_PITCH 784       // q|k|v tile row: 768 B + 16 B pad (ldmatrix conflict-free)
#define LB_DO_PITCH 272    // dout tile row: 256 B + 16
#define LB_QD_PITCH 528    // q|dout tile row of the dC pass: 512 B + 16
#define LB_M_PITCH 80      // 32x32 bf16 matrix row: 64 B + 16

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

// K-prep: fold the forward's per-chunk partials into M, S and the normalised context (fp32 and bf16)
__global__ void __launch_bounds__(256) linattn_bwd_prep_kernel(const float* __restrict__ fwd_ws, float* __restrict__ ws, int batch,
                                                               int n, int nchunks, long long ws_stride) {
  __shared__ float sS[DH], sMx[DH];
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;   // one CTA per (image, head)
  const float* pmax = fwd_ws + (size_t)b * nchunks * LA_C;
  const float* part = fwd_ws + (size_t)batch * nchunks * LA_C + (size_t)b * nchunks * LA_PART;
  float* w = ws + (size_t)b * ws_stride;
  bf16* cbf = reinterpret_cast<bf16*>(w + LAB_BF_OFF);
  if (tid < DH) {     // the forward's partials are relative to their own chunk maxima: merge with exp(m_c - M)
    const int hd = h * DH + tid;
    float m = -INFINITY, s = 0.0f;
    for (int c = 0; c < nchunks; ++c) m = fmaxf(m, pmax[(size_t)c * LA_C + hd]);
    for (int c = 0; c < nchunks; ++c) s += __expf(pmax[(size_t)c * LA_C + hd] - m) * part[(size_t)c * LA_PART + hd];
    w[hd] = m;
    w[LA_C + hd] = s;
    sS[tid] = s;
    sMx[tid] = m;
  }
  __syncthreads();
  for (int li = tid; li < DH * DH; li += 256) {
    const int d = li >> 5, hd = h * DH + d, idx = h * DH * DH + li;
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c) acc += __expf(pmax[(size_t)c * LA_C + hd] - sMx[d]) * part[(size_t)c * LA_PART + LA_C + idx];
    const float v = acc / (sS[d] * (float)n);
    w[3 * LA_C + idx] = v;
    cbf[idx] = __float2bfloat16_rn(v);
  }
}

// K-A: per (image, chunk) partial of sum_n p[n][d] dout[n][e] on tensor cores (same tiling as linattn_ctx_kernel)
__global__ void __launch_bounds__(256) linattn_bwd_dctx_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               float* __restrict__ ws, int n, int nchunks, long long ws_stride) {
  extern __shared__ __align__(16) uint8_t lab_smem[];
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = chunk * LA_CHUNK, p1 = min(n, p0 + LA_CHUNK);
  const int nsub = (p1 - p0 + LAB_TILE - 1) / LAB_TILE;
  const uint32_t tile0 = smem_u32(lab_smem);
  constexpr int TILE_BYTES = LAB_TILE * LB_QD_PITCH;

  auto load_sub = [&](int sub, int buf) {
    const int base_px = p0 + sub * LAB_TILE;
    for (int i = tid; i < LAB_TILE * 32; i += 256) {
      const int row = i >> 5, c16 = i & 31;
      const int px = base_px + row;
      const uint32_t dst = tile0 + buf * TILE_BYTES + row * LB_QD_PITCH + c16 * 16;
      if (px < p1) {
        const bf16* src = c16 < 16 ? qkv + ((size_t)b * n + px) * (3 * LA_C) + c16 * 8
                                   : dout + ((size_t)b * n + px) * LA_C + (c16 - 16) * 8;
        cp_async16(dst, src);
      } else {
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
      }
    }
  };

  const int h = warp & 3, slice = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][c][e] = 0.0f;

  load_sub(0, 0);
  cp_async_commit();
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
    uint8_t* tile_g = lab_smem + buf * TILE_BYTES;
    {  // q -> softmax over each head's 32 channels, in place (thread = pixel x head); padding rows -> 0
      const int row = tid >> 2, hh = tid & 3;
      uint4* qp = reinterpret_cast<uint4*>(tile_g + row * LB_QD_PITCH + hh * 64);
      if (p0 + sub * LAB_TILE + row < p1) {
        float x[32];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float t[8];
          unpack8(qp[j4], t);
#pragma unroll
          for (int e = 0; e < 8; ++e) x[j4 * 8 + e] = t[e];
        }
        softmax32(x);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          qp[j4] = make_uint4(pack_bf16x2(x[8 * j4], x[8 * j4 + 1]), pack_bf16x2(x[8 * j4 + 2], x[8 * j4 + 3]),
                              pack_bf16x2(x[8 * j4 + 4], x[8 * j4 + 5]), pack_bf16x2(x[8 * j4 + 6], x[8 * j4 + 7]));
      }
    }
    __syncthreads();
    const uint32_t tile = tile0 + buf * TILE_BYTES;
    const int j = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int k0 = slice * 32 + ks * 16;
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int px = k0 + (j >> 1) * 8 + rr, dcol = h * DH + mt * 16 + (j & 1) * 8;
        ldsm_x4_trans(tile + px * LB_QD_PITCH + dcol * 2, a[mt]);
      }
      uint32_t bfr[2][4];
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        const int px = k0 + (j & 1) * 8 + rr, ecol = h * DH + (2 * np + (j >> 1)) * 8;
        ldsm_x4_trans(tile + px * LB_QD_PITCH + 256 + ecol * 2, bfr[np]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
    }
    __syncthreads();  // the buffer is refilled by the next iteration's prefetch
  }
  // merge the two pixel slices through shared memory, then write the chunk partial
  float* red = reinterpret_cast<float*>(lab_smem);  // [4 heads][32 lanes][32]
  if (slice == 1) {
    float* r = red + (h * 32 + lane) * 33;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) r[(mt * 4 + nt) * 4 + e] = acc[mt][nt][e];
  }
  __syncthreads();
  if (slice == 0) {
    const float* r = red + (h * 32 + lane) * 33;
    float* dst = ws + (size_t)b * ws_stride + LAB_WS_FIXED + (size_t)chunk * LAB_MAT;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int d_lo = h * DH + mt * 16 + g, d_hi = d_lo + 8;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int e = nt * 8 + 2 * t4;
        const float* rr2 = r + (mt * 4 + nt) * 4;
        *reinterpret_cast<float2*>(dst + (size_t)d_lo * DH + e) = make_float2(acc[mt][nt][0] + rr2[0], acc[mt][nt][1] + rr2[1]);
        *reinterpret_cast<float2*>(dst + (size_t)d_hi * DH + e) = make_float2(acc[mt][nt][2] + rr2[2], acc[mt][nt][3] + rr2[3]);
      }
    }
  }
}

// K-C: dC = scale * sum of chunk partials (fp32, bf16, bf16 transposed) ; r[d] = sum_e dC[d][e] C[d][e]
__global__ void __launch_bounds__(256) linattn_bwd_combine_kernel(float* __restrict__ ws, int nchunks, float scale,
                                                                  long long ws_stride) {
  __shared__ float sprod[DH * DH];
  float* w = ws + (size_t)blockIdx.x * ws_stride;
  bf16* dcbf = reinterpret_cast<bf16*>(w + LAB_BF_OFF) + LAB_MAT;
  bf16* dctbf = dcbf + LAB_MAT;
  const int tid = threadIdx.x, hh = blockIdx.y;                  // one CTA per (image, head)
  for (int li = tid; li < DH * DH; li += 256) {
    const int idx = hh * DH * DH + li;
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c) acc += w[LAB_WS_FIXED + (size_t)c * LAB_MAT + idx];
    acc *= scale;
    w[3 * LA_C + LAB_MAT + idx] = acc;
    sprod[li] = acc * w[3 * LA_C + idx];
    const int d = li >> 5, e = li & 31;
    dcbf[idx] = __float2bfloat16_rn(acc);
    dctbf[(hh * DH + e) * DH + d] = __float2bfloat16_rn(acc);
  }
  __syncthreads();
  if (tid < DH) {
    float r = 0.0f;
    for (int e = 0; e < DH; ++e) r += sprod[tid * DH + e];
    w[2 * LA_C + hh * DH + tid] = r;
  }
}

// K-B: dqkv for 64 pixels per CTA.  warp = (head h, 32-pixel slice); q|k|v and dout tiles in shared memory; results
// overwrite the warp's own (rows, head columns) region of the q|k|v tile, which then leaves as coalesced 16-byte stores.
__global__ void __launch_bounds__(256) linattn_bwd_main_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                               const float* __restrict__ ws, bf16* __restrict__ dqkv, int n,
                                                               float scale, long long ws_stride) {
  extern __shared__ __align__(16) uint8_t lab_smem[];
  uint8_t* tile_g = lab_smem;                                   // [64][LB_PITCH]
  uint8_t* do_g = tile_g + LAB_TILE * LB_PITCH;                 // [64][LB_DO_PITCH]
  uint8_t* mat_g = do_g + LAB_TILE * LB_DO_PITCH;               // 3 x [128][LB_M_PITCH]: C, dC, dC^T
  __shared__ float sM[LA_C], sIS[LA_C], sr[LA_C];
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = blockIdx.x * LAB_TILE;
  const float* w = ws + (size_t)b * ws_stride;
  const bf16* mats = reinterpret_cast<const bf16*>(w + LAB_BF_OFF);
  const uint32_t tile = smem_u32(tile_g), do_s = smem_u32(do_g), mat_s = smem_u32(mat_g);
  for (int i = tid; i < LAB_TILE * 48; i += 256) {
    const int row = i / 48, c16 = i % 48;
    const uint32_t dst = tile + row * LB_PITCH + c16 * 16;
    if (p0 + row < n) cp_async16(dst, qkv + ((size_t)b * n + p0 + row) * (3 * LA_C) + c16 * 8);
    else asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
  }
  for (int i = tid; i < LAB_TILE * 16; i += 256) {
    const int row = i >> 4, c16 = i & 15;
    const uint32_t dst = do_s + row * LB_DO_PITCH + c16 * 16;
    if (p0 + row < n) cp_async16(dst, dout + ((size_t)b * n + p0 + row) * LA_C + c16 * 8);
    else asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
  }
  for (int i = tid; i < 3 * LA_C * 4; i += 256) {   // 3 matrices x 128 rows x 4 chunks of 16 B
    const int m = i / (LA_C * 4), row = (i / 4) % LA_C, c16 = i & 3;
    cp_async16(mat_s + (m * LA_C + row) * LB_M_PITCH + c16 * 16, mats + (size_t)m * LAB_MAT + row * DH + c16 * 8);
  }
  cp_async_commit();
  if (tid < LA_C) {
    sM[tid] = w[tid];
    sIS[tid] = 1.0f / w[LA_C + tid];
    sr[tid] = w[2 * LA_C + tid];
  }
  cp_async_wait<0>();
  __syncthreads();

  const int h = warp & 3, slice = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const float inv_n = 1.0f / (float)n;
  // B fragments of the three matrices (stored [n][k]): C[d][e] (n=d,k=e), dC[d][e] (n=d,k=e), dC^T[e][d] (n=e,k=d)
  uint32_t bC[2][2][4], bdC[2][2][4], bdCT[2][2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int nrow = h * DH + (2 * np + (j >> 1)) * 8 + rr, kcol = ks * 16 + (j & 1) * 8;
      ldsm_x4(mat_s + (0 * LA_C + nrow) * LB_M_PITCH + kcol * 2, bC[ks][np]);
      ldsm_x4(mat_s + (1 * LA_C + nrow) * LB_M_PITCH + kcol * 2, bdC[ks][np]);
      ldsm_x4(mat_s + (2 * LA_C + nrow) * LB_M_PITCH + kcol * 2, bdCT[ks][np]);
    }
  // per-column constants of the k softmax for this thread's 8 columns d = nt*8 + 2*t4 + c
  float cM[8], cIS[8], cr[8];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int d = h * DH + nt * 8 + 2 * t4 + c;
      cM[nt * 2 + c] = sM[d];
      cIS[nt * 2 + c] = sIS[d];
      cr[nt * 2 + c] = sr[d];
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int m0 = slice * 32 + mt * 16;
    const int arow = m0 + (j & 1) * 8 + rr;                    // ldmatrix row address of this lane (A operand)
    const int acol = h * DH + (j >> 1) * 8;                    // + ks*16
    uint32_t aD[2][4], aX[2][4];
    float acc[4][4];
    float x[2][8];
    // ------------------------------------------------ dq
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      ldsm_x4(do_s + arow * LB_DO_PITCH + (acol + ks * 16) * 2, aD[ks]);
      ldsm_x4(tile + arow * LB_PITCH + (acol + ks * 16) * 2, aX[ks]);
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 f = unpack_bf16x2(aX[ks][r]);
        x[r & 1][ks * 4 + (r >> 1) * 2] = f.x;
        x[r & 1][ks * 4 + (r >> 1) * 2 + 1] = f.y;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {   // softmax over the 32 channels of rows g (r=0) and g+8 (r=1): quad reduction
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
      const float inv = 1.0f / s;
#pragma unroll
      for (int i = 0; i < 8; ++i) x[r][i] *= inv;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) mma_bf16_16816(acc[nt], aD[ks], bC[ks][nt >> 1][(nt & 1) * 2], bC[ks][nt >> 1][(nt & 1) * 2 + 1]);
    }
    {
      float dot[2] = {0.0f, 0.0f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[nt][c] *= scale;
          dot[c >> 1] = fmaf(x[c >> 1][nt * 2 + (c & 1)], acc[nt][c], dot[c >> 1]);
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], 1);
        dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], 2);
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = h * DH + nt * 8 + 2 * t4;
        *reinterpret_cast<uint32_t*>(tile_g + (m0 + g) * LB_PITCH + col * 2) =
            pack_bf16x2(x[0][nt * 2] * (acc[nt][0] - dot[0]), x[0][nt * 2 + 1] * (acc[nt][1] - dot[0]));
        *reinterpret_cast<uint32_t*>(tile_g + (m0 + g + 8) * LB_PITCH + col * 2) =
            pack_bf16x2(x[1][nt * 2] * (acc[nt][2] - dot[1]), x[1][nt * 2 + 1] * (acc[nt][3] - dot[1]));
      }
    }
    // ------------------------------------------------ dk, dv
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      ldsm_x4(tile + arow * LB_PITCH + (2 * LA_C + acol + ks * 16) * 2, aD[ks]);   // v
      ldsm_x4(tile + arow * LB_PITCH + (LA_C + acol + ks * 16) * 2, aX[ks]);       // k
    }
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 f = unpack_bf16x2(aX[ks][r]);
        const int i0 = ks * 4 + (r >> 1) * 2;
        x[r & 1][i0] = __expf(f.x - cM[i0]) * cIS[i0];
        x[r & 1][i0 + 1] = __expf(f.y - cM[i0 + 1]) * cIS[i0 + 1];
        aX[ks][r] = pack_bf16x2(x[r & 1][i0], x[r & 1][i0 + 1]);                    // kh as the A operand of dv
      }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) mma_bf16_16816(acc[nt], aD[ks], bdC[ks][nt >> 1][(nt & 1) * 2], bdC[ks][nt >> 1][(nt & 1) * 2 + 1]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = LA_C + h * DH + nt * 8 + 2 * t4;
      const int i0 = nt * 2;
      *reinterpret_cast<uint32_t*>(tile_g + (m0 + g) * LB_PITCH + col * 2) =
          pack_bf16x2(x[0][i0] * (acc[nt][0] * inv_n - cr[i0]), x[0][i0 + 1] * (acc[nt][1] * inv_n - cr[i0 + 1]));
      *reinterpret_cast<uint32_t*>(tile_g + (m0 + g + 8) * LB_PITCH + col * 2) =
          pack_bf16x2(x[1][i0] * (acc[nt][2] * inv_n - cr[i0]), x[1][i0 + 1] * (acc[nt][3] * inv_n - cr[i0 + 1]));
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) mma_bf16_16816(acc[nt], aX[ks], bdCT[ks][nt >> 1][(nt & 1) * 2], bdCT[ks][nt >> 1][(nt & 1) * 2 + 1]);
      const int col = 2 * LA_C + h * DH + nt * 8 + 2 * t4;
      *reinterpret_cast<uint32_t*>(tile_g + (m0 + g) * LB_PITCH + col * 2) = pack_bf16x2(acc[nt][0] * inv_n, acc[nt][1] * inv_n);
      *reinterpret_cast<uint32_t*>(tile_g + (m0 + g + 8) * LB_PITCH + col * 2) = pack_bf16x2(acc[nt][2] * inv_n, acc[nt][3] * inv_n);
    }
  }
  __syncthreads();
  for (int i = tid; i < LAB_TILE * 48; i += 256) {
    const int row = i / 48, c16 = i % 48;
    if (p0 + row < n)
      *reinterpret_cast<uint4*>(dqkv + ((size_t)b * n + p0 + row) * (3 * LA_C) + c16 * 8) =
          *reinterpret_cast<const uint4*>(tile_g + row * LB_PITCH + c16 * 16);
  }
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
  const int smem_a = 2 * LAB_TILE * LB_QD_PITCH;
  const int smem_b = LAB_TILE * (LB_PITCH + LB_DO_PITCH) + 3 * LA_C * LB_M_PITCH;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(linattn_bwd_dctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
    TEDM_CUDA(cudaFuncSetAttribute(linattn_bwd_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    configured = true;
  }
  linattn_bwd_prep_kernel<<<dim3(batch, LA_HEADS), 256, 0, s>>>(fwd_workspace, workspace, batch, n, nchunks, ws_stride);
  TEDM_LAUNCH_CHECK();
  linattn_bwd_dctx_kernel<<<dim3(nchunks, batch), 256, smem_a, s>>>((const bf16*)qkv, (const bf16*)dout, workspace, n, nchunks,
                                                                    ws_stride);
  TEDM_LAUNCH_CHECK();
  linattn_bwd_combine_kernel<<<dim3(batch, LA_HEADS), 256, 0, s>>>(workspace, nchunks, scale, ws_stride);
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
