// Fused inference forward of Residual(PreNorm(LinearAttention))               models/unet_model.py:29-36,64-73,178-210
//
//   out = LayerNorm_out( W_o . linattn( W_qkv . LayerNorm_pre(x) ) + b_o ) + x
//
// for the two high-resolution levels of the UNet (C = 64 at 128^2, C = 128 at 64^2), where the unfused chain
// (LayerNorm, qkv conv, colmax, ctx, out, to_out conv, LayerNorm + residual) moves 3.2 KB per pixel through HBM for
// 0.1 MFLOP of arithmetic.  Here q, k, v, the attention output and the to_out result never leave the SM:
//
//   K-A  ctx : per (image, 1024-pixel chunk): x tile -> LayerNorm -> [k|v]^T = W_kv y^T on mma.sync (accumulator
//              fragments of the transposed product ARE the A / B fragments of the next product, so P = exp(k - m)
//              and v go register-to-register into ctx += P^T v); softmax over n is ONLINE (running per-channel
//              maximum with rescaling), so k is computed once.                      reads x once, writes a partial
//   K-C  combine: merge chunk partials (maxima, sums, contexts) -> bf16 ctx^T * scale / (s n)              (tiny)
//   K-B  out : per 16-pixel group, one warp, no block-level barriers: x -> LayerNorm -> q_h = W_q y -> softmax_d ->
//              q_h ctx_h -> accumulated straight into the to_out product -> + bias -> LayerNorm -> + x -> store.
//                                                                                   reads x once, writes out once
// HBM traffic: 3 x 2C bytes per pixel (384 B at C = 64) instead of ~3200 B.
#include "common.cuh"

namespace {

constexpr int kHeads = 4, kDh = 32, kHid = kHeads * kDh;   // 128 channels each for q, k, v
constexpr int kChunk = 1024;                                // pixels per CTA
constexpr int kSub = 64;                                    // pixels per K-A tile
constexpr int kPart = 2 * kHid + kHid * kDh;                // per-chunk partial: m[128], s[128], ctx[128][32]
constexpr int kCtPitch = 80;                                // bytes per ctx^T row in smem (64 + 16)
constexpr int kWoPitch = kHid * 2 + 16;                     // bytes per W_o row in smem

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// first k-step of a product: C = 0 comes from the zero register instead of 4 MOVs per accumulator tile
__device__ __forceinline__ void mma16816_z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
template <bool ZERO>
__device__ __forceinline__ void mma_acc(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (ZERO) mma16816_z(c, a, b0, b1);
  else mma16816(c, a, b0, b1);
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// In-place channel LayerNorm (gain only) of 8 pixel rows by one warp: lane = (row, quarter of the channels).
// `src` and `dst` may alias.  Rows are PITCH bytes apart; 16-byte vector v of a lane covers channels (4 v + part) * 8.
template <int C, int PITCH>
__device__ __forceinline__ void ln_rows8(const uint8_t* src, uint8_t* dst, const float* gain, float eps, int lane) {
  constexpr int NV = C / 32;
  const int row = lane >> 2, part = lane & 3;
  float f[NV][8];
  float s = 0.0f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    unpack8(*reinterpret_cast<const uint4*>(src + row * PITCH + (v * 4 + part) * 16), f[v]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[v][j];
  }
  const float mean = quad_sum(s) * (1.0f / C);
  float q = 0.0f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[v][j] -= mean;
      q = fmaf(f[v][j], f[v][j], q);
    }
  const float rstd = rsqrtf(quad_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 g0 = *reinterpret_cast<const float4*>(gain + (v * 4 + part) * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(gain + (v * 4 + part) * 8 + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = f[v][j] * rstd * gg[j];
    *reinterpret_cast<uint4*>(dst + row * PITCH + (v * 4 + part) * 16) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K-A.  grid (nchunks, B), 256 threads.  warp = (head h, 32-pixel slice of the 64-pixel tile).
// ------------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256, 2) linattn_fused_ctx_kernel(const bf16* __restrict__ x, const bf16* __restrict__ wqkv,
                                                                const float* __restrict__ g_pre, float* __restrict__ part,
                                                                int n, int nchunks, float eps) {
  constexpr int PITCH = C * 2 + 16;
  constexpr int TILE = kSub * PITCH;
  constexpr int VEC = C / 8;  // 16-byte vectors per row
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* w_s = smem;                          // [256][PITCH]: rows 0..127 = W_k, 128..255 = W_v
  uint8_t* t_s = smem + 2 * kHid * PITCH;       // 2 tiles [64][PITCH]
  float* gain = reinterpret_cast<float*>(t_s + 2 * TILE);
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = chunk * kChunk, p1 = min(n, p0 + kChunk);
  const int nsub = (p1 - p0) / kSub;
  const bf16* xb = x + (size_t)b * n * C;

  auto load_tile = [&](int sub, int buf) {
    const bf16* src = xb + (size_t)(p0 + sub * kSub) * C;
    for (int i = tid; i < kSub * VEC; i += 256) {
      const int row = i / VEC, v = i % VEC;
      cp16(smem_u32(t_s + buf * TILE + row * PITCH + v * 16), src + (size_t)row * C + v * 8);
    }
  };
  for (int i = tid; i < 2 * kHid * VEC; i += 256) {
    const int row = i / VEC, v = i % VEC;
    cp16(smem_u32(w_s + row * PITCH + v * 16), wqkv + (size_t)(kHid + row) * C + v * 8);
  }
  load_tile(0, 0);
  cp_commit();
  for (int i = tid; i < C; i += 256) gain[i] = g_pre[i];

  const int h = warp & 3, slice = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  float ctx[2][4][4];   // [d m-tile][e n-tile][frag]
  float ssum[2][2], mrun[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    ssum[mt][0] = ssum[mt][1] = 0.0f;
    mrun[mt][0] = mrun[mt][1] = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) ctx[mt][nt][e] = 0.0f;
  }
  const uint32_t w_u = smem_u32(w_s);

  for (int sub = 0; sub < nsub; ++sub) {
    const int buf = sub & 1;
    cp_wait<0>();
    __syncthreads();                      // tile `sub` (and, first time, the weights) landed; tile sub-1 fully consumed
    if (sub + 1 < nsub) {
      load_tile(sub + 1, buf ^ 1);
      cp_commit();
    }
    uint8_t* tile = t_s + buf * TILE;
    ln_rows8<C, PITCH>(tile + warp * 8 * PITCH, tile + warp * 8 * PITCH, gain, eps, lane);
    __syncthreads();
    const uint32_t t_u = smem_u32(tile);
    // k_h^T and then v_h^T (32 channels x 32 px each) = W rows x y^T; two passes keep the kernel under 128 registers
    auto gemm_rows = [&](int row0, float (&acc)[2][4][4]) {
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        uint32_t bfr[2][4];
#pragma unroll
        for (int np = 0; np < 2; ++np)
          ldsm_x4(t_u + (slice * 32 + (2 * np + (j >> 1)) * 8 + rr) * PITCH + (ks * 16 + (j & 1) * 8) * 2, bfr[np]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t a[4];
          ldsm_x4(w_u + (row0 + mt * 16 + (j & 1) * 8 + rr) * PITCH + (ks * 16 + (j >> 1) * 8) * 2, a);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (ks == 0) mma16816_z(acc[mt][nt], a, bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
            else mma16816(acc[mt][nt], a, bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
          }
        }
      }
    };
    float acc[2][4][4];
    gemm_rows(h * kDh, acc);
    // online softmax over pixels, per k channel (rows g / g+8 of the two m-tiles)
    uint32_t pa[2][2][4];
    float corrs[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float tm = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) tm = fmaxf(tm, fmaxf(acc[mt][nt][2 * r], acc[mt][nt][2 * r + 1]));
        tm = quad_max(tm);
        const float mn = fmaxf(mrun[mt][r], tm);
        const float corr = ex2f((mrun[mt][r] - mn) * kLog2e);
        mrun[mt][r] = mn;
        const float mn2 = -mn * kLog2e;
        float ls = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float e0 = ex2f(fmaf(acc[mt][nt][2 * r], kLog2e, mn2)), e1 = ex2f(fmaf(acc[mt][nt][2 * r + 1], kLog2e, mn2));
          ls += e0 + e1;
          pa[mt][nt >> 1][(nt & 1) * 2 + r] = pack_bf16x2(e0, e1);
        }
        ssum[mt][r] = fmaf(ssum[mt][r], corr, ls);
        corrs[mt][r] = corr;
      }
    // the running maxima settle after the first tiles: rescale the context only when one of them moved (warp-uniform)
    if (__any_sync(0xffffffffu, corrs[0][0] != 1.0f || corrs[0][1] != 1.0f || corrs[1][0] != 1.0f || corrs[1][1] != 1.0f)) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          ctx[mt][nt][0] *= corrs[mt][0];
          ctx[mt][nt][1] *= corrs[mt][0];
          ctx[mt][nt][2] *= corrs[mt][1];
          ctx[mt][nt][3] *= corrs[mt][1];
        }
    }
    gemm_rows(kHid + h * kDh, acc);
    // ctx[d][e] += sum_px P[d][px] v[e][px]: B fragments straight from the v^T accumulators
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int ent = 0; ent < 4; ++ent) {
        const int mv = ent >> 1, hf = (ent & 1) * 2;
        const uint32_t b0 = pack_bf16x2(acc[mv][2 * ks][hf], acc[mv][2 * ks][hf + 1]);
        const uint32_t b1 = pack_bf16x2(acc[mv][2 * ks + 1][hf], acc[mv][2 * ks + 1][hf + 1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma16816(ctx[mt][ent], pa[mt][ks], b0, b1);
      }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int r = 0; r < 2; ++r) ssum[mt][r] = quad_sum(ssum[mt][r]);
  // merge the two pixel slices (different running maxima) through shared memory, then write the chunk partial
  __syncthreads();
  float* red = reinterpret_cast<float*>(w_s);  // [4 heads][32 lanes][40] (20 KB; the weights are no longer needed)
  if (slice == 1) {
    float* r = red + (h * 32 + lane) * 40;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) r[(mt * 4 + nt) * 4 + e] = ctx[mt][nt][e];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        r[32 + mt * 2 + q] = ssum[mt][q];
        r[36 + mt * 2 + q] = mrun[mt][q];
      }
    }
  }
  __syncthreads();
  if (slice == 0) {
    const float* r = red + (h * 32 + lane) * 40;
    float* dst = part + ((size_t)b * nchunks + chunk) * kPart;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float m1 = r[36 + mt * 2 + q];
        const float M = fmaxf(mrun[mt][q], m1);
        const float f0 = __expf(mrun[mt][q] - M), f1 = __expf(m1 - M);
        const int d = h * kDh + mt * 16 + g + q * 8;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float* rr2 = r + (mt * 4 + nt) * 4 + 2 * q;
          *reinterpret_cast<float2*>(dst + 2 * kHid + (size_t)d * kDh + nt * 8 + 2 * t4) =
              make_float2(ctx[mt][nt][2 * q] * f0 + rr2[0] * f1, ctx[mt][nt][2 * q + 1] * f0 + rr2[1] * f1);
        }
        if (t4 == 0) {
          dst[d] = M;
          dst[kHid + d] = ssum[mt][q] * f0 + r[32 + mt * 2 + q] * f1;
        }
      }
  }
}

// K-C: ctxT[b][h][e][d] = bf16( scale * sum_p w_p ctx_p[h][d][e] / (n * sum_p w_p s_p[h][d]) ),  w_p = exp(m_p - max_p m_p)
__global__ void __launch_bounds__(256) linattn_fused_combine_kernel(const float* __restrict__ part, bf16* __restrict__ ctxT,
                                                                    int n, int nchunks, float scale) {
  __shared__ float sM[kDh], sS[kDh];
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;   // one CTA per (image, head)
  const float* p = part + (size_t)b * nchunks * kPart;
  if (tid < kDh) {
    const int hd = h * kDh + tid;
    float M = -INFINITY;
    for (int c = 0; c < nchunks; ++c) M = fmaxf(M, p[(size_t)c * kPart + hd]);
    float s = 0.0f;
    for (int c = 0; c < nchunks; ++c) s += __expf(p[(size_t)c * kPart + hd] - M) * p[(size_t)c * kPart + kHid + hd];
    sM[tid] = M;
    sS[tid] = s;
  }
  __syncthreads();
  for (int li = tid; li < kDh * kDh; li += 256) {
    const int d = li >> 5, e = li & 31, hd = h * kDh + d;
    float acc = 0.0f;
    for (int c = 0; c < nchunks; ++c) acc += __expf(p[(size_t)c * kPart + hd] - sM[d]) * p[(size_t)c * kPart + 2 * kHid + hd * kDh + e];
    ctxT[(((size_t)b * kHeads + h) * kDh + e) * kDh + d] = __float2bfloat16_rn(acc * scale / (sS[d] * (float)n));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K-B.  grid (n / 2048, B), 512 threads.  Each warp walks groups of 16 MT pixels on its own (private x / y buffers,
// __syncwarp only), so loads, tensor-core work and stores of different warps overlap freely.  MT = 2 at C = 64: every
// weight fragment read from shared memory feeds two m-tiles (the kernel is shared-memory-bandwidth bound at MT = 1).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kChunkB = 2048, kWarpsB = 16;

template <int C, int MT>
__global__ void __launch_bounds__(kWarpsB * 32, 1) linattn_fused_out_kernel(
    const bf16* __restrict__ x, const bf16* __restrict__ wqkv, const bf16* __restrict__ wout, const float* __restrict__ g_pre,
    const float* __restrict__ b_out, const float* __restrict__ g_out, const bf16* __restrict__ ctxT, bf16* __restrict__ out,
    int n, float eps) {
  constexpr int PITCH = C * 2 + 16;
  constexpr int VEC = C / 8;
  constexpr int NT = C / 8;             // n-tiles of the to_out product
  constexpr int ROWS = 16 * MT;         // pixels per group
  constexpr int WBUF = 2 * ROWS * PITCH;  // per warp: x (prefetched one group ahead) + y / output staging
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* wq_s = smem;                               // [128][PITCH]
  uint8_t* wo_s = wq_s + kHid * PITCH;                // [C][kWoPitch]
  uint8_t* ct_s = wo_s + C * kWoPitch;                // [128][kCtPitch]
  uint8_t* wb_s = ct_s + kHid * kCtPitch;             // kWarpsB x WBUF
  float* fpar = reinterpret_cast<float*>(wb_s + kWarpsB * WBUF);   // g_pre[C], b_out[C], g_out[C]
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p0 = chunk * kChunkB, p1 = min(n, p0 + kChunkB);
  const int ngroups = (p1 - p0) / ROWS;
  const bf16* xb = x + (size_t)b * n * C;
  bf16* ob = out + (size_t)b * n * C;
  uint8_t* xs = wb_s + warp * WBUF;
  uint8_t* ys = xs + ROWS * PITCH;

  auto load_group = [&](int grp) {           // ROWS px x C channels, contiguous in global memory
    const bf16* src = xb + (size_t)(p0 + grp * ROWS) * C;
    for (int i = lane; i < ROWS * VEC; i += 32) {
      const int row = i / VEC, v = i % VEC;
      cp16(smem_u32(xs + row * PITCH + v * 16), src + (size_t)row * C + v * 8);
    }
  };
  for (int i = tid; i < kHid * VEC; i += kWarpsB * 32) {
    const int row = i / VEC, v = i % VEC;
    cp16(smem_u32(wq_s + row * PITCH + v * 16), wqkv + (size_t)row * C + v * 8);
  }
  for (int i = tid; i < C * (kHid / 8); i += kWarpsB * 32) {
    const int row = i / (kHid / 8), v = i % (kHid / 8);
    cp16(smem_u32(wo_s + row * kWoPitch + v * 16), wout + (size_t)row * kHid + v * 8);
  }
  for (int i = tid; i < kHid * 4; i += kWarpsB * 32) {
    const int row = i >> 2, v = i & 3;
    cp16(smem_u32(ct_s + row * kCtPitch + v * 16), ctxT + ((size_t)b * kHid + row) * kDh + v * 8);
  }
  if (warp < ngroups) load_group(warp);
  cp_commit();
  for (int i = tid; i < C; i += kWarpsB * 32) {
    fpar[i] = g_pre[i];
    fpar[C + i] = b_out[i];
    fpar[2 * C + i] = g_out[i];
  }
  cp_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const uint32_t wq_u = smem_u32(wq_s), wo_u = smem_u32(wo_s), ct_u = smem_u32(ct_s), y_u = smem_u32(ys);
  for (int grp = warp; grp < ngroups; grp += kWarpsB) {
    cp_wait<0>();
    __syncwarp();
#pragma unroll
    for (int r8 = 0; r8 < 2 * MT; ++r8) ln_rows8<C, PITCH>(xs + r8 * 8 * PITCH, ys + r8 * 8 * PITCH, fpar, eps, lane);
    __syncwarp();
    if (grp + kWarpsB < ngroups) {           // x is consumed: fetch the next group under this one's tensor-core work
      load_group(grp + kWarpsB);
      cp_commit();
    }
    float o2[MT][NT][4];          // starts from the to_out bias (fragment columns nt*8 + 2*t4, +1)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float2 b2 = *reinterpret_cast<const float2*>(fpar + C + nt * 8 + 2 * t4);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        o2[mt][nt][0] = o2[mt][nt][2] = b2.x;
        o2[mt][nt][1] = o2[mt][nt][3] = b2.y;
      }
    }
#pragma unroll 1
    for (int h = 0; h < kHeads; ++h) {
      // q_h (ROWS px x 32 d) = y W_q,h^T
      float q[MT][4][4];
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          ldsm_x4(y_u + (mt * 16 + (j & 1) * 8 + rr) * PITCH + (ks * 16 + (j >> 1) * 8) * 2, a[mt]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t bfr[4];
          ldsm_x4(wq_u + (h * kDh + (2 * np + (j >> 1)) * 8 + rr) * PITCH + (ks * 16 + (j & 1) * 8) * 2, bfr);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (ks == 0) {
              mma16816_z(q[mt][2 * np], a[mt], bfr[0], bfr[1]);
              mma16816_z(q[mt][2 * np + 1], a[mt], bfr[2], bfr[3]);
            } else {
              mma16816(q[mt][2 * np], a[mt], bfr[0], bfr[1]);
              mma16816(q[mt][2 * np + 1], a[mt], bfr[2], bfr[3]);
            }
          }
        }
      }
      // softmax over d for rows g (frag 0, 1) and g + 8 (frag 2, 3); packed straight into A fragments
      float inv[MT][2];
      uint32_t qa[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float m = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) m = fmaxf(m, fmaxf(q[mt][nt][2 * r], q[mt][nt][2 * r + 1]));
          m = -quad_max(m) * kLog2e;
          float s = 0.0f;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float e0 = ex2f(fmaf(q[mt][nt][2 * r], kLog2e, m)), e1 = ex2f(fmaf(q[mt][nt][2 * r + 1], kLog2e, m));
            s += e0 + e1;
            qa[mt][nt >> 1][(nt & 1) * 2 + r] = pack_bf16x2(e0, e1);
          }
          inv[mt][r] = 1.0f / quad_sum(s);
        }
      // o_h (ROWS px x 32 e) = softmax(q_h) ctx_h
      float o[MT][4][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t bfr[4];
          ldsm_x4(ct_u + (h * kDh + (2 * np + (j >> 1)) * 8 + rr) * kCtPitch + (ks * 16 + (j & 1) * 8) * 2, bfr);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (ks == 0) {
              mma16816_z(o[mt][2 * np], qa[mt][ks], bfr[0], bfr[1]);
              mma16816_z(o[mt][2 * np + 1], qa[mt][ks], bfr[2], bfr[3]);
            } else {
              mma16816(o[mt][2 * np], qa[mt][ks], bfr[0], bfr[1]);
              mma16816(o[mt][2 * np + 1], qa[mt][ks], bfr[2], bfr[3]);
            }
          }
        }
      // o2 += o_h W_o[:, h*32 : h*32+32]^T
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          a[mt][0] = pack_bf16x2(o[mt][2 * ks][0] * inv[mt][0], o[mt][2 * ks][1] * inv[mt][0]);
          a[mt][1] = pack_bf16x2(o[mt][2 * ks][2] * inv[mt][1], o[mt][2 * ks][3] * inv[mt][1]);
          a[mt][2] = pack_bf16x2(o[mt][2 * ks + 1][0] * inv[mt][0], o[mt][2 * ks + 1][1] * inv[mt][0]);
          a[mt][3] = pack_bf16x2(o[mt][2 * ks + 1][2] * inv[mt][1], o[mt][2 * ks + 1][3] * inv[mt][1]);
        }
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t bfr[4];
          ldsm_x4(wo_u + ((2 * np + (j >> 1)) * 8 + rr) * kWoPitch + (h * kDh + ks * 16 + (j & 1) * 8) * 2, bfr);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma16816(o2[mt][2 * np], a[mt], bfr[0], bfr[1]);
            mma16816(o2[mt][2 * np + 1], a[mt], bfr[2], bfr[3]);
          }
        }
      }
    }
    // + bias, LayerNorm over the C channels of each row, * g_out, + x (re-read: an L2 hit), -> bf16 staging -> store
    const float* go = fpar + 2 * C;
    const bf16* xg = xb + (size_t)(p0 + grp * ROWS) * C;
    __syncwarp();                            // every lane is done reading y: the buffer becomes the output staging area
    float mean[MT][2], rstd[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float s = 0.0f, qv = 0.0f;                 // one pass: E[x], E[x^2] (values are O(1), fp32 accumulators)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          s += o2[mt][nt][2 * r] + o2[mt][nt][2 * r + 1];
          qv = fmaf(o2[mt][nt][2 * r], o2[mt][nt][2 * r], fmaf(o2[mt][nt][2 * r + 1], o2[mt][nt][2 * r + 1], qv));
        }
        mean[mt][r] = quad_sum(s) * (1.0f / C);
        const float var = fmaxf(quad_sum(qv) * (1.0f / C) - mean[mt][r] * mean[mt][r], 0.0f);
        rstd[mt][r] = rsqrtf(var + eps);
      }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = nt * 8 + 2 * t4;
      const float2 g2 = *reinterpret_cast<const float2*>(go + c);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = mt * 16 + g + r * 8;
          const float2 res = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xg + (size_t)row * C + c)));
          const float sc0 = rstd[mt][r] * g2.x, sc1 = rstd[mt][r] * g2.y;
          const float v0 = fmaf(o2[mt][nt][2 * r] - mean[mt][r], sc0, res.x);
          const float v1 = fmaf(o2[mt][nt][2 * r + 1] - mean[mt][r], sc1, res.y);
          *reinterpret_cast<uint32_t*>(ys + row * PITCH + c * 2) = pack_bf16x2(v0, v1);
        }
    }
    __syncwarp();
    bf16* dst = ob + (size_t)(p0 + grp * ROWS) * C;
    for (int i = lane; i < ROWS * VEC; i += 32) {
      const int row = i / VEC, v = i % VEC;
      *reinterpret_cast<uint4*>(dst + (size_t)row * C + v * 8) = *reinterpret_cast<const uint4*>(ys + row * PITCH + v * 16);
    }
    __syncwarp();   // the staging buffer is rewritten by the next group's LayerNorm
  }
}

template <int C, int MT>
int launch_fused(const bf16* x, const bf16* wqkv, const float* g_pre, const bf16* wout, const float* b_out, const float* g_out,
                 bf16* out, float* workspace, int batch, int n, float scale, float eps, cudaStream_t s) {
  constexpr int PITCH = C * 2 + 16;
  const int nchunks = (n + kChunk - 1) / kChunk;
  float* part = workspace;
  bf16* ctxT = reinterpret_cast<bf16*>(part + (size_t)batch * nchunks * kPart);
  const int smem_a = 2 * kHid * PITCH + 2 * kSub * PITCH + C * 4;
  const int smem_b = kHid * PITCH + C * kWoPitch + kHid * kCtPitch + kWarpsB * 2 * 16 * MT * PITCH + 3 * C * 4;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(linattn_fused_ctx_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
    TEDM_CUDA(cudaFuncSetAttribute(linattn_fused_ctx_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    TEDM_CUDA(cudaFuncSetAttribute(linattn_fused_out_kernel<C, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    TEDM_CUDA(cudaFuncSetAttribute(linattn_fused_out_kernel<C, MT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  linattn_fused_ctx_kernel<C><<<dim3(nchunks, batch), 256, smem_a, s>>>(x, wqkv, g_pre, part, n, nchunks, eps);
  TEDM_LAUNCH_CHECK();
  linattn_fused_combine_kernel<<<dim3(batch, kHeads), 256, 0, s>>>(part, ctxT, n, nchunks, scale);
  TEDM_LAUNCH_CHECK();
  linattn_fused_out_kernel<C, MT><<<dim3((n + kChunkB - 1) / kChunkB, batch), kWarpsB * 32, smem_b, s>>>(
      x, wqkv, wout, g_pre, b_out, g_out, ctxT, out, n, eps);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

}  // namespace

extern "C" int tedm_linear_attention_fused_supported(int n, int channels, int heads, int dim_head) {
  return heads == kHeads && dim_head == kDh && (channels == 64 || channels == 128) && n > 0 && n % kSub == 0;
}

extern "C" int64_t tedm_linear_attention_fused_workspace(int batch, int n) {
  if (batch <= 0 || n <= 0) return -1;
  const int64_t nchunks = (n + kChunk - 1) / kChunk;
  return (int64_t)batch * (nchunks * kPart + kHid * kDh / 2);
}

extern "C" int tedm_linear_attention_fused_fwd(const void* x, const void* wqkv, const float* g_pre, const void* wout,
                                               const float* b_out, const float* g_out, void* out, float* workspace, int batch,
                                               int n, int channels, int heads, int dim_head, float scale, float eps,
                                               tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && wqkv && g_pre && wout && b_out && g_out && out && workspace && batch > 0 && batch <= 65535,
                 "tedm_linear_attention_fused_fwd: bad arguments");
  TEDM_UNSUPPORTED(!tedm_linear_attention_fused_supported(n, channels, heads, dim_head),
                   "tedm_linear_attention_fused_fwd: n=%d channels=%d heads=%d dim_head=%d (needs 4 x 32 heads, 64 or 128 "
                   "channels, n a multiple of 64)", n, channels, heads, dim_head);
  cudaStream_t s = (cudaStream_t)stream;
  if (channels == 64)
    return launch_fused<64, 2>((const bf16*)x, (const bf16*)wqkv, g_pre, (const bf16*)wout, b_out, g_out, (bf16*)out, workspace,
                            batch, n, scale, eps, s);
  return launch_fused<128, 1>((const bf16*)x, (const bf16*)wqkv, g_pre, (const bf16*)wout, b_out, g_out, (bf16*)out, workspace,
                           batch, n, scale, eps, s);
}
