// Backward of the mid-block attention for ANY token count, flash style on mma.sync      models/unet_model.py:213-241
//   forward:  qn = q / |q|_n, kn = k / |k|_n (column norms over the n tokens) ; S = scale * qn kn^T ; A = softmax_j S ; O = A v
//   backward: dV = A^T dO ; dA = dO V^T ; dS = A o (dA - delta), delta_i = <dO_i, O_i> ; dqn = scale dS kn ; dkn = scale dS^T qn ;
//             dq = (dqn - qn colsum(dqn o qn)) / |q|_n, dk likewise.
// Four kernels per (image, head), nothing n x n ever leaves the SM:
//   K0 stats : norms, logsumexp_i (online over key tiles), delta_i                          (tile = 64 queries)
//   K1 dkv   : per 64-key tile, loop over query tiles with the TRANSPOSED products S^T = kn qn^T, dA^T = v dO^T, so the
//              accumulator fragments of P^T / dS^T are the A fragments of dV += P^T dO and dkn += dS^T qn
//   K2 dq    : per 64-query tile, loop over key tiles: dqn += dS kn
//   K3 norm  : the L2-normalisation backward (column sums over tokens, then the rescale), bf16 results into dqkv
// The one-CTA-per-(image, head) fp32 kernel in attention_bwd.cu stays for n < 64.
#include "common.cuh"

namespace {

constexpr int kD = 32, kPitch = 80, kTile = 64;

__device__ __forceinline__ void ldsm(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// workspace per (image, head): inv_q[32] (unscaled), inv_k[32], lse[n], delta[n], dqn[n][32], dkn[n][32]
__host__ __device__ inline long long ws_stride(int n) { return 64 + 2LL * n + 64LL * n; }
struct Ws {
  float *inv, *lse, *delta, *dqn, *dkn;
};
__device__ __forceinline__ Ws ws_of(float* ws, int b, int h, int heads, int n) {
  float* base = ws + ((size_t)b * heads + h) * ws_stride(n);
  return {base, base + 64, base + 64 + n, base + 64 + 2 * (size_t)n, base + 64 + 2 * (size_t)n + 32 * (size_t)n};
}

// 64 rows x 32 channels of `src` (row stride `stride` elements, rows row0.. of n) -> smem tile, optionally scaled per channel
__device__ __forceinline__ void load_tile(uint8_t* dst, const bf16* src, size_t stride, int row0, int n, const float* chan_scale,
                                          float extra, int tid) {
  const int row = tid >> 1, half = tid & 1;
  const int r = row0 + row;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r < n) {
      u = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * stride + half * 16 + v * 8));
      if (chan_scale) {
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] *= chan_scale[half * 16 + v * 8 + e] * extra;
        u = pack8(f);
      }
    }
    *reinterpret_cast<uint4*>(dst + row * kPitch + (half * 16 + v * 8) * 2) = u;
  }
}

// C (16 x 64) = A-fragments (16 rows x 32) x tile^T, tile = [64 cols][32 k] row-major in smem
__device__ __forceinline__ void gemm_16x64(float (&c)[8][4], const uint32_t (&a)[2][4], uint32_t tile_u, int lane) {
  const int j = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm(tile_u + ((2 * np + (j >> 1)) * 8 + rr) * kPitch + (ks * 16 + (j & 1) * 8) * 2, b);
      mma(c[2 * np], a[ks], b[0], b[1]);
      mma(c[2 * np + 1], a[ks], b[2], b[3]);
    }
}
// A fragments (16 rows x 32 k) of rows row0.. of a [64][32] smem tile
__device__ __forceinline__ void load_a(uint32_t (&a)[2][4], uint32_t tile_u, int row0, int lane) {
  const int j = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) ldsm(tile_u + (row0 + (j & 1) * 8 + rr) * kPitch + (ks * 16 + (j >> 1) * 8) * 2, a[ks]);
}
// acc (16 x 32) += P (16 x 64, C-layout fragments packed as A) x tile, tile = [64 k][32 n] row-major in smem (ldmatrix.trans)
__device__ __forceinline__ void gemm_acc_16x32(float (&acc)[4][4], const uint32_t (&pa)[4][4], uint32_t tile_u, int lane) {
  const int j = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b[4];
      ldsm_t(tile_u + (ks * 16 + (j & 1) * 8 + rr) * kPitch + ((2 * np + (j >> 1)) * 8) * 2, b);
      mma(acc[2 * np], pa[ks], b[0], b[1]);
      mma(acc[2 * np + 1], pa[ks], b[2], b[3]);
    }
}

// ---- K0 -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_stats_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                             const bf16* __restrict__ dout, float* __restrict__ ws, int n,
                                                             int heads, float scale) {
  __shared__ __align__(16) uint8_t sq[kTile * kPitch], sk[kTile * kPitch];
  __shared__ float s_inv[64], s_red[64];
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * kD, C = heads * kD;
  const bf16* base = qkv + (size_t)b * n * C3;
  const Ws w = ws_of(ws, b, h, heads, n);
  {
    const int c = tid & 63, part = tid >> 6;
    const int ch = c < kD ? h * kD + c : C + h * kD + (c - kD);
    float acc = 0.0f;
    for (int px = part; px < n; px += 2) {
      const float v = __bfloat162float(base[(size_t)px * C3 + ch]);
      acc = fmaf(v, v, acc);
    }
    if (part == 1) s_red[c] = acc;
    __syncthreads();
    if (part == 0) {
      s_inv[c] = 1.0f / fmaxf(sqrtf(acc + s_red[c]), 1e-12f);
      if (blockIdx.x == 0) w.inv[c] = s_inv[c];
    }
    __syncthreads();
  }
  load_tile(sq, base + h * kD, C3, q0, n, s_inv, scale, tid);
  const int g = lane >> 2, t4 = lane & 3;
  float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.0f, 0.0f};
  uint32_t aq[2][4];
  for (int k0 = 0; k0 < n; k0 += kTile) {
    __syncthreads();
    load_tile(sk, base + C + h * kD, C3, k0, n, s_inv + kD, 1.0f, tid);
    __syncthreads();
    if (k0 == 0) load_a(aq, smem_u32(sq), warp * 16, lane);
    float sc[8][4];
    gemm_16x64(sc, aq, smem_u32(sk), lane);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float tm = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (k0 + nt * 8 + 2 * t4 + e >= n) sc[nt][2 * r + e] = -INFINITY;
          tm = fmaxf(tm, sc[nt][2 * r + e]);
        }
      tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
      tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 2));
      const float mn = fmaxf(mrun[r], tm);
      float ls = 0.0f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) ls += __expf(sc[nt][2 * r] - mn) + __expf(sc[nt][2 * r + 1] - mn);
      lrun[r] = fmaf(lrun[r], __expf(mrun[r] - mn), ls);
      mrun[r] = mn;
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = lrun[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int qi = q0 + warp * 16 + g + r * 8;
    if (t4 == 0 && qi < n) w.lse[qi] = mrun[r] + __logf(l);
  }
  if (tid < kTile && q0 + tid < n) {
    const size_t off = ((size_t)b * n + q0 + tid) * C + h * kD;
    float d = 0.0f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float fo[8], fd[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(o + off) + v), fo);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dout + off) + v), fd);
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(fo[e], fd[e], d);
    }
    w.delta[q0 + tid] = d;
  }
}

// ---- K1 / K2 share the inner step: P and dS fragments of a 16 x 64 block ------------------------------------------------
// s, da: C-layout fragments.  ROWSTAT: lse / delta belong to the rows (K2) or to the columns (K1, transposed products).
template <bool ROWSTAT>
__device__ __forceinline__ void p_and_ds(const float (&s)[8][4], const float (&da)[8][4], const float* lse, const float* delta,
                                         int col0, int ncols_valid, int lane, uint32_t (&pp)[4][4], uint32_t (&pds)[4][4],
                                         const float (&row_lse)[2], const float (&row_delta)[2]) {
  const int t4 = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float p[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = nt * 8 + 2 * t4 + (e & 1);
      const float l = ROWSTAT ? row_lse[e >> 1] : lse[col];
      const float dl = ROWSTAT ? row_delta[e >> 1] : delta[col];
      p[e] = col0 + col < ncols_valid ? __expf(s[nt][e] - l) : 0.0f;
      ds[e] = p[e] * (da[nt][e] - dl);
    }
    pp[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p[0], p[1]);
    pp[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
    pds[nt >> 1][(nt & 1) * 2] = pack_bf16x2(ds[0], ds[1]);
    pds[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
  }
}

// ---- K1: dV and dkn for one 64-key tile -------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                           bf16* __restrict__ dqkv, float* __restrict__ ws, int n, int heads,
                                                           float scale) {
  __shared__ __align__(16) uint8_t sk[kTile * kPitch], sv[kTile * kPitch], sq[kTile * kPitch], sdo[kTile * kPitch];
  __shared__ float s_inv[64], s_lse[kTile], s_delta[kTile];
  const int k0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * kD, C = heads * kD;
  const bf16* base = qkv + (size_t)b * n * C3;
  const Ws w = ws_of(ws, b, h, heads, n);
  if (tid < 64) s_inv[tid] = w.inv[tid];
  __syncthreads();
  load_tile(sk, base + C + h * kD, C3, k0, n, s_inv + kD, 1.0f, tid);
  load_tile(sv, base + 2 * C + h * kD, C3, k0, n, nullptr, 1.0f, tid);
  __syncthreads();
  uint32_t ak[2][4], av[2][4];
  load_a(ak, smem_u32(sk), warp * 16, lane);
  load_a(av, smem_u32(sv), warp * 16, lane);
  float dv[4][4], dk[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.0f;
  const float zero2[2] = {0.0f, 0.0f};
  for (int q0 = 0; q0 < n; q0 += kTile) {
    __syncthreads();
    load_tile(sq, base + h * kD, C3, q0, n, s_inv, scale, tid);                                  // scale * qn
    load_tile(sdo, dout + (size_t)b * n * C + h * kD, C, q0, n, nullptr, 1.0f, tid);
    if (tid < kTile) {
      s_lse[tid] = q0 + tid < n ? w.lse[q0 + tid] : 0.0f;
      s_delta[tid] = q0 + tid < n ? w.delta[q0 + tid] : 0.0f;
    }
    __syncthreads();
    float st[8][4], dat[8][4];
    gemm_16x64(st, ak, smem_u32(sq), lane);       // S^T  (keys x queries)
    gemm_16x64(dat, av, smem_u32(sdo), lane);     // dA^T
    uint32_t pp[4][4], pds[4][4];
    p_and_ds<false>(st, dat, s_lse, s_delta, q0, n, lane, pp, pds, zero2, zero2);
    gemm_acc_16x32(dv, pp, smem_u32(sdo), lane);  // dV  += P^T dO
    gemm_acc_16x32(dk, pds, smem_u32(sq), lane);  // dkn += dS^T (scale qn)
  }
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = k0 + warp * 16 + g + r * 8;
    if (key < n) {
      bf16* dvp = dqkv + ((size_t)b * n + key) * C3 + 2 * C + h * kD + 2 * t4;
      float* dkp = w.dkn + (size_t)key * kD + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<uint32_t*>(dvp + nt * 8) = pack_bf16x2(dv[nt][2 * r], dv[nt][2 * r + 1]);
        *reinterpret_cast<float2*>(dkp + nt * 8) = make_float2(dk[nt][2 * r], dk[nt][2 * r + 1]);
      }
    }
  }
}

// ---- K2: dqn for one 64-query tile ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                          float* __restrict__ ws, int n, int heads, float scale) {
  __shared__ __align__(16) uint8_t sk[kTile * kPitch], sv[kTile * kPitch], sq[kTile * kPitch], sdo[kTile * kPitch];
  __shared__ float s_inv[64];
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * kD, C = heads * kD;
  const bf16* base = qkv + (size_t)b * n * C3;
  const Ws w = ws_of(ws, b, h, heads, n);
  if (tid < 64) s_inv[tid] = w.inv[tid];
  __syncthreads();
  load_tile(sq, base + h * kD, C3, q0, n, s_inv, scale, tid);
  load_tile(sdo, dout + (size_t)b * n * C + h * kD, C, q0, n, nullptr, 1.0f, tid);
  __syncthreads();
  uint32_t aq[2][4], ado[2][4];
  load_a(aq, smem_u32(sq), warp * 16, lane);
  load_a(ado, smem_u32(sdo), warp * 16, lane);
  const int g = lane >> 2, t4 = lane & 3;
  float row_lse[2], row_delta[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = q0 + warp * 16 + g + r * 8;
    row_lse[r] = qi < n ? w.lse[qi] : 0.0f;
    row_delta[r] = qi < n ? w.delta[qi] : 0.0f;
  }
  float dq[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.0f;
  for (int k0 = 0; k0 < n; k0 += kTile) {
    __syncthreads();
    load_tile(sk, base + C + h * kD, C3, k0, n, s_inv + kD, 1.0f, tid);
    load_tile(sv, base + 2 * C + h * kD, C3, k0, n, nullptr, 1.0f, tid);
    __syncthreads();
    float s[8][4], da[8][4];
    gemm_16x64(s, aq, smem_u32(sk), lane);        // S  (queries x keys)
    gemm_16x64(da, ado, smem_u32(sv), lane);      // dA
    uint32_t pp[4][4], pds[4][4];
    p_and_ds<true>(s, da, nullptr, nullptr, k0, n, lane, pp, pds, row_lse, row_delta);
    gemm_acc_16x32(dq, pds, smem_u32(sk), lane);  // dqn/scale += dS kn
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = q0 + warp * 16 + g + r * 8;
    if (qi < n) {
      float* p = w.dqn + (size_t)qi * kD + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<float2*>(p + nt * 8) = make_float2(dq[nt][2 * r] * scale, dq[nt][2 * r + 1] * scale);
    }
  }
}

// ---- K3: normalisation backward ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_bwd_norm_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ dqkv,
                                                            const float* __restrict__ ws, int n, int heads) {
  __shared__ float red[2][8][kD], csum[2][kD];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, d = tid & 31, rg = tid >> 5;
  const int C3 = 3 * heads * kD, C = heads * kD;
  const Ws w = ws_of(const_cast<float*>(ws), b, h, heads, n);
  const float iq = w.inv[d], ik = w.inv[kD + d];
  const bf16* base = qkv + (size_t)b * n * C3 + h * kD + d;
  float aq = 0.0f, ak = 0.0f;
  for (int i = rg; i < n; i += 8) {
    aq = fmaf(w.dqn[(size_t)i * kD + d], __bfloat162float(base[(size_t)i * C3]) * iq, aq);
    ak = fmaf(w.dkn[(size_t)i * kD + d], __bfloat162float(base[(size_t)i * C3 + C]) * ik, ak);
  }
  red[0][rg][d] = aq;
  red[1][rg][d] = ak;
  __syncthreads();
  if (tid < 2 * kD) {
    const int which = tid >> 5;
    float s = 0.0f;
    for (int r = 0; r < 8; ++r) s += red[which][r][d];
    csum[which][d] = s;
  }
  __syncthreads();
  const float cq = csum[0][d], ck = csum[1][d];
  bf16* ob = dqkv + (size_t)b * n * C3 + h * kD + d;
  for (int i = rg; i < n; i += 8) {
    const float qn = __bfloat162float(base[(size_t)i * C3]) * iq, kn = __bfloat162float(base[(size_t)i * C3 + C]) * ik;
    ob[(size_t)i * C3] = __float2bfloat16_rn((w.dqn[(size_t)i * kD + d] - qn * cq) * iq);
    ob[(size_t)i * C3 + C] = __float2bfloat16_rn((w.dkn[(size_t)i * kD + d] - kn * ck) * ik);
  }
}

}  // namespace

extern "C" int64_t tedm_attention_bwd_flash_workspace(int batch, int n, int heads) {
  if (batch <= 0 || n <= 0 || heads <= 0) return -1;
  return (int64_t)batch * heads * ws_stride(n);
}

extern "C" int tedm_attention_bwd_flash(const void* qkv, const void* o, const void* dout, void* dqkv, float* workspace, int batch,
                                        int n, int heads, int dim_head, float scale, tedm_stream_t stream) {
  TEDM_CHECK_ARG(qkv && o && dout && dqkv && workspace && batch > 0 && n > 0 && heads > 0, "tedm_attention_bwd_flash: bad arguments");
  TEDM_UNSUPPORTED(dim_head != kD, "tedm_attention_bwd_flash: dim_head=%d (only 32)", dim_head);
  TEDM_CHECK_ARG(batch <= 65535 && heads <= 65535, "tedm_attention_bwd_flash: batch / heads too large");
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid((n + kTile - 1) / kTile, heads, batch);
  attn_bwd_stats_kernel<<<grid, 128, 0, s>>>((const bf16*)qkv, (const bf16*)o, (const bf16*)dout, workspace, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<<<grid, 128, 0, s>>>((const bf16*)qkv, (const bf16*)dout, (bf16*)dqkv, workspace, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  attn_bwd_dq_kernel<<<grid, 128, 0, s>>>((const bf16*)qkv, (const bf16*)dout, workspace, n, heads, scale);
  TEDM_LAUNCH_CHECK();
  attn_bwd_norm_kernel<<<dim3(heads, batch), 256, 0, s>>>((const bf16*)qkv, (bf16*)dqkv, workspace, n, heads);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
