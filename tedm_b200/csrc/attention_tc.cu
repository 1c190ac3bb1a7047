// Residual(PreNorm(LinearAttention)) inference forward with every GEMM on tcgen05   models/unet_model.py:29-36,64-73,178-210
//
//   out = LayerNorm_out( W_o . linattn( W_qkv . LayerNorm_pre(x) ) + b_o ) + x            (C = 64 at 128^2, C = 128 at 64^2)
//
// The mma.sync version of this block (attention_fused.cu) is bound by instruction issue: 10 non-MMA instructions per
// m16n8k16.  Here a 128-pixel tile is one UMMA row block, one thread owns one pixel end to end (LayerNorm and both
// softmaxes need no cross-thread traffic), and the algebra is rearranged so that nothing but the two softmax operands
// ever makes the TMEM -> registers -> shared memory round trip:
//
//   K-A ctx :  y = LN(x) (bf16, in place in the TMA-landed tile)                                      compute warps
//              k = y Wk^T                        M = 128 px, N = 128, K = C          (tile -> TMEM)   tcgen05
//              P = exp(k - shift)                shift[hd] >= max_n k[hd, n] is a WEIGHT-ONLY bound (Cauchy-Schwarz:
//                                                |k| <= ||w_hd * g||_2 sqrt(C)), so the softmax over n needs no running
//                                                maximum, no rescaling and no second pass
//              G += P^T [y | 1]                  M = 128 (h,d), N = C + 16, K = 128 px, both operands MN-major: the SAME
//                                                shared-memory y tile is the B operand; the ones block yields S = sum_n P
//                 v is never computed per pixel:  ctx[d][e] = sum_n P[n,d] v[n,e] = sum_c G[d][c] Wv[e][c]
//   K-C comb:  per image: ctx = G Wv^T / (S n);  M[c'][hd] = scale * sum_e ctx[h][d][e] W_o[c'][h,e]   (to_out folded in)
//   K-B out :  y = LN(x);  q = y Wq^T (tcgen05);  Q = softmax_d(q) per head (in registers) -> bf16 tile
//              o = Q M^T                         M = 128 px, N = C, K = 128 (tcgen05)  = to_out(attention output)
//              out = LN_out(o + b_o) * g_o + x   -> bf16 tile -> TMA store
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 UMMA issuer, warps 2-5 / 6-9 two compute groups that alternate
// tiles (a thread's TMEM lane quarter is warp % 4, its pixel = quarter * 32 + lane).  HBM traffic: x once per kernel plus
// the output: 3 x 2C bytes per pixel.  Per pixel and kernel the cost is ~128 MUFU ex2 and ~1 k issued instructions.
#include "common.cuh"

namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
// bf16 [rows][cols] row-major -> 2-D map with a (64 cols, box_rows) box, 128B swizzle
int encode_2d(CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_rows) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return tedm_set_error(TEDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2ull};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return tedm_set_error(TEDM_ERR_CUDA, "cuTensorMapEncodeTiled(%lld x %lld) failed: %d", rows, cols, (int)r);
  return TEDM_OK;
}

constexpr int TM = 128;                 // pixels per tile = UMMA M
constexpr int HID = 128;                // heads * dim_head
constexpr int BLK = TM * 128;           // bytes of one [128 rows][64 bf16] swizzled block
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {      // K-major SW128 (rows of 128 B, 8-row atoms 1 KB apart)
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr) {     // MN-major SW128: 64-element MN blocks 16 KB apart, K groups 1 KB
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1024ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(desc), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t sw_addr(uint32_t blk_base, int row, int c16) {   // 16-byte chunk c16 of a block's row
  return blk_base + (uint32_t)row * 128u + (uint32_t)((c16 ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// TPP threads share a pixel (1: the thread owns the whole row; 2: thread h2 owns channels [h2 C/2, (h2+1) C/2)).  A row is
// C/8 16-byte chunks over C/64 swizzled blocks; thread h2 owns chunks h2 * C/(8 TPP) ... of that linear order.
template <int C, int TPP>
__device__ __forceinline__ uint32_t part_addr(uint32_t base, int row, int h2, int i) {
  const int l = h2 * (C / 8 / TPP) + i;
  return sw_addr(base + (l >> 3) * BLK, row, l & 7);
}

// Per-pixel LayerNorm statistics.  TPP = 2: the two half-row partial sums are exchanged through two 8-byte shared-memory
// slots (`mine` written, `theirs` read) around one named barrier of the group's threads.  REUSE: the slots lie inside a
// buffer the caller is about to overwrite (no spare shared memory at C = 128), so a second barrier keeps the partner's
// read ahead of that overwrite.  -> (rstd, -mean * rstd)
template <int C, int TPP, bool REUSE>
__device__ __forceinline__ float2 ln_stats(float s, float q, uint32_t mine, uint32_t theirs, int bar_id, float eps) {
  if constexpr (TPP == 2) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(mine), "f"(s), "f"(q) : "memory");
    named_bar_sync(bar_id, 256);
    float ox, oy;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ox), "=f"(oy) : "r"(theirs) : "memory");
    if (REUSE) named_bar_sync(bar_id, 256);
    s += ox;
    q += oy;
  }
  const float mean = s * (1.0f / C);
  const float var = fmaxf(q * (1.0f / C) - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + eps);
  return make_float2(rstd, -mean * rstd);
}

// LayerNorm (no gain: folded into the projection weights, W' = W diag(g)) of this thread's part of pixel `row`:
// src tile -> bf16 xhat at dst (may equal src).  One pass, four independent accumulators per statistic.
template <int C, int TPP, bool REUSE>
__device__ __forceinline__ void ln_part(uint32_t src, uint32_t dst, int row, int h2, uint32_t mine, uint32_t theirs, int bar_id,
                                        float eps) {
  constexpr int NCH = C / 8 / TPP;                 // 16-byte chunks of this thread
  float v[NCH * 8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8];
    unpack8(lds128(part_addr<C, TPP>(src, row, h2, i)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i * 8 + j] = f[j];
  }
  float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) {
    s4[i & 3] += v[i];
    q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]);
  }
  const float2 st = ln_stats<C, TPP, REUSE>((s4[0] + s4[1]) + (s4[2] + s4[3]), (q4[0] + q4[1]) + (q4[2] + q4[3]), mine, theirs,
                                            bar_id, eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(v[i * 8 + j], st.x, st.y);
    sts128(part_addr<C, TPP>(dst, row, h2, i), pack8(f));
  }
}

__device__ __forceinline__ uint64_t desc_mn_lbo(uint32_t addr, uint32_t lbo_bytes) {   // MN-major SW128 with an explicit block stride
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

struct TcMaps {
  CUtensorMap x, w, m, out;
};

struct TcParams {
  int n;                 // pixels per image
  int tiles_per_image;   // n / 128
  int chunk_tiles;       // K-A: tiles per work item
  int items;             // K-A: work items = batch * (tiles_per_image / chunk_tiles)
  int total_tiles;       // K-B
  float eps;
  float shift_log2;      // log2(e) * an upper bound of every k logit          (K-A)
  float* part;           // [items][128][C + 16] fp32                          (K-A)
  const float* b_out;    // [C]                                                (K-B)
  const float* g_out;    // [C]                                                (K-B)
};

// ==========================================================================================================
// K-A: per (image, chunk of tiles):  G[hd][c] = sum_px P[px][hd] y[px][c],  S[hd] = sum_px P[px][hd]
// ==========================================================================================================
template <int C, int NG, int TPP>
__global__ void __launch_bounds__(64 + NG * 128 * TPP, 1) linattn_tc_ctx_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int GT = 128 * TPP;                    // threads of a compute group: TPP threads per pixel
  constexpr int TC_THREADS = 64 + NG * GT;         // producer warp, issuer warp, NG compute groups
  constexpr int KB = C / 64;                       // 64-channel blocks of a pixel row
  constexpr int N2 = C + 16;                       // G columns + 16 columns of S (the ones block)
  constexpr int NS = C == 64 ? 6 : 3;              // x / y ring: tiles in flight towards this SM (HBM latency is ~2 tile times)
  constexpr int XBUF = KB * BLK;
  constexpr bool MERGE_S = C == 64;                // one 64-column block: [y | ones] is a single B operand (explicit block stride)
  constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  constexpr uint32_t IDESC_G = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)((MERGE_S ? N2 : C) >> 3) << 17) | ((uint32_t)(HID >> 4) << 24);
  constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) |
                               ((uint32_t)(HID >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, x_full[NS], x_empty[NS], y_ready[NG], d1_full[NG], p_ready[NG], d2_full, d2_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 exch[TPP == 2 ? NG : 1][TPP == 2 ? 2 * TM : 1];   // LayerNorm partial sums: [group][half][pixel]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_s = base;                                  // Wk: KB blocks of [128 hd][64 c]
  const uint32_t x_s = w_s + KB * BLK;                        // ring of NS tiles, LayerNormed in place
  const uint32_t one_s = x_s + NS * XBUF;                     // a block of bf16 ones (after the ring: a positive block stride)
  const uint32_t p_s = one_s + BLK;                           // [NG] x P: two MN blocks [128 px][64 hd]

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w);
    mbar_init(smem_u32(&bar_w), 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&x_full[s]), 1);
      mbar_init(smem_u32(&x_empty[s]), 1);
    }
    for (int g = 0; g < NG; ++g) {
      mbar_init(smem_u32(&y_ready[g]), GT);
      mbar_init(smem_u32(&d1_full[g]), 1);
      mbar_init(smem_u32(&p_ready[g]), GT);
    }
    mbar_init(smem_u32(&d2_full), 1);
    mbar_init(smem_u32(&d2_empty), GT);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  for (int i = threadIdx.x; i < BLK / 16; i += TC_THREADS)
    sts128(one_s + i * 16, make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t d2_col = NG * HID;                           // D1[g] at columns g * 128; then G (C columns) and S (16)

  // tiles of this CTA: items blockIdx.x, + gridDim.x, ...; local tile index i counts across items.
  // tile i: ring stage i % NS, compute group i % NG
  const int my_items = p.items > (int)blockIdx.x ? (p.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int my_tiles = my_items * p.chunk_tiles;
  const int chunks_per_image = p.tiles_per_image / p.chunk_tiles;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t bw = smem_u32(&bar_w);
      mbar_expect_tx(bw, KB * BLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * BLK, &maps.w, bw, kb * 64, HID);   // rows 128..255 of wqkv = Wk
      int st = 0;
      uint32_t sph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int it = i / p.chunk_tiles, tl = i - it * p.chunk_tiles;
        const int item = blockIdx.x + it * gridDim.x;
        const int img = item / chunks_per_image, ch = item - img * chunks_per_image;
        const long long row0 = (long long)img * p.n + (long long)(ch * p.chunk_tiles + tl) * TM;
        mbar_wait(smem_u32(&x_empty[st]), sph ^ 1u);
        const uint32_t full = smem_u32(&x_full[st]);
        mbar_expect_tx(full, KB * BLK);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(x_s + st * XBUF + kb * BLK, &maps.x, full, kb * 64, (int)row0);
        if (++st == NS) {
          st = 0;
          sph ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(smem_u32(&bar_w), 0);
      tc_fence_after();
      auto gemm2 = [&](int j) {                     // G (+)= P_j^T y_j ;  S (+)= P_j^T 1
        const int g = j % NG, tl = j % p.chunk_tiles, st = j % NS;
        if (tl == 0 && j > 0) {                     // a new item: the previous item's accumulator must have been drained
          mbar_wait(smem_u32(&d2_empty), (uint32_t)(((j / p.chunk_tiles - 1) & 1)));
          tc_fence_after();
        }
        mbar_wait(smem_u32(&p_ready[g]), (uint32_t)((j / NG) & 1));
        tc_fence_after();
        const uint32_t yb = x_s + st * XBUF;
        const uint64_t adesc = desc_mn(p_s + g * 2 * BLK);
        if constexpr (MERGE_S) {
          const uint64_t bdesc = desc_mn_lbo(yb, one_s - yb);          // block 0 = y, "block 1" = the shared ones block
#pragma unroll
          for (int k = 0; k < TM / 16; ++k) umma_bf16(tmem + d2_col, adesc + 128ull * k, bdesc + 128ull * k, IDESC_G, (tl | k) != 0 ? 1u : 0u);
        } else {
          const uint64_t bdesc = desc_mn(yb), sdesc = desc_mn(one_s);
#pragma unroll
          for (int k = 0; k < TM / 16; ++k) {
            umma_bf16(tmem + d2_col, adesc + 128ull * k, bdesc + 128ull * k, IDESC_G, (tl | k) != 0 ? 1u : 0u);
            umma_bf16(tmem + d2_col + C, adesc + 128ull * k, sdesc + 128ull * k, IDESC_S, (tl | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&x_empty[st]));        // y_j (and P_j) are free once these complete
        if (tl == p.chunk_tiles - 1) umma_commit(smem_u32(&d2_full));
      };
      for (int i = 0; i < my_tiles; ++i) {
        const int g = i % NG, st = i % NS;
        mbar_wait(smem_u32(&y_ready[g]), (uint32_t)((i / NG) & 1));
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t adesc = desc_k(x_s + st * XBUF + kb * BLK), bdesc = desc_k(w_s + kb * BLK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + g * HID, adesc + 2ull * k, bdesc + 2ull * k, IDESC1, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&d1_full[g]));
        if (i > 0) gemm2(i - 1);
      }
      if (my_tiles > 0) gemm2(my_tiles - 1);
    }
    __syncwarp();
  } else {
    const int grp = (warp - 2) / (4 * TPP), h2 = TPP == 2 ? ((warp - 2) >> 2) & 1 : 0, q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float nsh = -p.shift_log2;
    for (int i = grp; i < my_tiles; i += NG) {
      const uint32_t ph = (uint32_t)((i / NG) & 1);
      const int st = i % NS;
      const uint32_t yb = x_s + st * XBUF;
      mbar_wait(smem_u32(&x_full[st]), (uint32_t)((i / NS) & 1));
      ln_part<C, TPP, false>(yb, yb, row, h2, TPP == 2 ? smem_u32(&exch[TPP == 2 ? grp : 0][TPP == 2 ? h2 * TM + row : 0]) : 0u,
                             TPP == 2 ? smem_u32(&exch[TPP == 2 ? grp : 0][TPP == 2 ? (h2 ^ 1) * TM + row : 0]) : 0u, 1 + grp, p.eps);
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&y_ready[grp]));
      mbar_wait(smem_u32(&d1_full[grp]), ph);
      tc_fence_after();
      // this thread's 4 / TPP heads (32 columns each).  One shift for all columns: a per-column constant cancels in G / S.
      constexpr int NH = 4 / TPP;
      const uint32_t pbase = p_s + grp * 2 * BLK;
      uint32_t r[2][32];
      tmem_ld32(lane_addr + (uint32_t)(grp * HID + h2 * NH * 32), r[0]);
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {                    // the next head's load is in flight while this one is processed
        tmem_ld_wait();
        if (hh + 1 < NH) tmem_ld32(lane_addr + (uint32_t)(grp * HID + (h2 * NH + hh + 1) * 32), r[(hh + 1) & 1]);
        const int gh = h2 * NH + hh;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = ex2f(fmaf(__uint_as_float(r[hh & 1][c * 8 + j]), kLog2e, nsh));
          sts128(sw_addr(pbase + (gh >> 1) * BLK, row, (gh & 1) * 4 + c), pack8(f));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&p_ready[grp]));
      if (i % p.chunk_tiles == p.chunk_tiles - 1) {        // last tile of an item: this group drains the accumulator
        const int it = i / p.chunk_tiles;
        mbar_wait(smem_u32(&d2_full), (uint32_t)(it & 1));
        tc_fence_after();
        const int item = blockIdx.x + it * gridDim.x;
        float* dst = p.part + ((size_t)item * HID + row) * N2;
        constexpr int N16 = N2 / 16, SPLIT = TPP == 2 ? (N16 + 1) / 2 : N16;  // 16-column chunks: the first SPLIT to thread 0 of the pixel
#pragma unroll 1
        for (int cc = h2 ? SPLIT : 0; cc < (h2 ? N16 : SPLIT); ++cc) {
          uint32_t rr[16];
          tmem_ld16(lane_addr + d2_col + (uint32_t)(cc * 16), rr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(dst + cc * 16 + 4 * j) = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]),
                                                                            __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&d2_empty));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ==========================================================================================================
// K-C: per image   ctx[h][d][e] = sum_c G[hd][c] Wv[he][c] / (S[hd] n);   M[c'][hd] = scale sum_e ctx[h][d][e] Wo[c'][he]
// ==========================================================================================================
template <int C>
__global__ void __launch_bounds__(256) linattn_tc_combine_kernel(const float* __restrict__ part, const bf16* __restrict__ wqkv,
                                                                 const bf16* __restrict__ wout, bf16* __restrict__ mimg,
                                                                 int chunks, int n, float scale) {
  // CTA = (image, head): G_h [32 d][C], S_h [32] summed over the chunks; ctx_h = G_h Wv_h^T / (S n); M[c'][h*32+d]
  constexpr int N2 = C + 16, GP = C + 1;
  __shared__ float buf[2 * 32 * GP], S[32], ctx[32][33];      // G | Wv_h first, then Wo_h in the same storage
  float* G = buf;
  float* wv = buf + 32 * GP;
  float* wo = buf;                                             // [C][33] <= 2 * 32 * (C + 1) floats
  const int b = blockIdx.x, h = blockIdx.y;
  const float* src = part + ((size_t)b * chunks * HID + h * 32) * N2;
  for (int i = threadIdx.x; i < 32 * GP; i += 256) {               // consecutive threads -> consecutive columns of a row
    const int d = i / GP, c = i % GP;
    float acc = 0.0f;
    for (int k = 0; k < chunks; ++k) acc += src[((size_t)k * HID + d) * N2 + c];
    if (c < C) G[d * GP + c] = acc;
    else S[d] = acc;
  }
  for (int i = threadIdx.x; i < 32 * C; i += 256) {
    const int e = i / C, c = i % C;
    wv[e * GP + c] = __bfloat162float(wqkv[(size_t)(2 * HID + h * 32 + e) * C + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += 256) {
    const int d = i >> 5, e = i & 31;
    float acc = 0.0f;
#pragma unroll 8
    for (int c = 0; c < C; ++c) acc = fmaf(G[d * GP + c], wv[e * GP + c], acc);
    ctx[d][e] = acc / (S[d] * (float)n);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 32; i += 256) {
    const int c = i >> 5, e = i & 31;
    wo[c * 33 + e] = __bfloat162float(wout[(size_t)c * HID + h * 32 + e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 32; i += 256) {
    const int c = i >> 5, d = i & 31;
    float acc = 0.0f;
#pragma unroll
    for (int e = 0; e < 32; ++e) acc = fmaf(ctx[d][e], wo[c * 33 + e], acc);
    mimg[((size_t)b * C + c) * HID + h * 32 + d] = __float2bfloat16_rn(acc * scale);
  }
}

// ==========================================================================================================
// K-B: out = LN_out( softmax_d(LN(x) Wq^T) M^T + b_o ) g_o + x
// ==========================================================================================================
template <int C, int NG, int TPP>
__global__ void __launch_bounds__(64 + NG * 128 * TPP, 1) linattn_tc_out_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  constexpr int GT = 128 * TPP;                    // threads of a compute group: TPP threads per pixel
  constexpr int TC_THREADS = 64 + NG * GT;
  constexpr int KB = C / 64;
  constexpr int NS = C == 64 ? (NG >= 3 ? 4 : 6) : 3;   // x ring (a tile stays until its residual has been added)
  constexpr int XBUF = KB * BLK;
  constexpr int WBUF = 2 * BLK;                    // per group: y (KB blocks) -> Q (2 blocks) -> output staging (KB blocks), in turn
  constexpr int MBLK = C * 128;                    // bytes of one [C rows][64 hd] block of M
  constexpr int HC = C / TPP;                      // channels per thread
  constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  constexpr uint32_t IDESC3 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, x_full[NS], x_empty[NS], y_ready[NG], d1_full[NG], q_ready[NG], d3_full[NG], m_full, m_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ float bo_s[C], go_s[C];
  // LayerNorm partial sums of the two threads of a pixel: at C = 64 a static [group][half][pixel] array; at C = 128 every
  // byte of shared memory is taken, so the slots are the first 8 bytes of each thread's own destination in the work buffer
  constexpr bool EXCH_IN_WB = C == 128;
  constexpr bool EXCH_STATIC = TPP == 2 && !EXCH_IN_WB;
  __shared__ float2 exch[EXCH_STATIC ? NG : 1][EXCH_STATIC ? 2 * TM : 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_s = base;                                  // Wq: KB blocks of [128 hd][64 c]
  const uint32_t m_s = w_s + KB * BLK;                        // M of the current image: 2 blocks of [C][64 hd]
  const uint32_t x_s = m_s + 2 * MBLK;                        // ring of NS x tiles
  const uint32_t b_s = x_s + NS * XBUF;                       // [NG] work buffers

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.m);
    tma_prefetch_desc(&maps.out);
    mbar_init(smem_u32(&bar_w), 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&x_full[s]), 1);
      mbar_init(smem_u32(&x_empty[s]), GT);
    }
    for (int g = 0; g < NG; ++g) {
      mbar_init(smem_u32(&y_ready[g]), GT);
      mbar_init(smem_u32(&d1_full[g]), 1);
      mbar_init(smem_u32(&q_ready[g]), GT);
      mbar_init(smem_u32(&d3_full[g]), 1);
    }
    mbar_init(smem_u32(&m_full), 1);
    mbar_init(smem_u32(&m_empty), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  for (int i = threadIdx.x; i < C; i += TC_THREADS) {
    bo_s[i] = p.b_out[i];
    go_s[i] = p.g_out[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;                            // group g: columns g * 128 hold q, then (q consumed) o

  // a contiguous range of tiles per CTA (consecutive tiles share an image, so M is reloaded rarely).
  // tile i of the range: ring stage i % NS, compute group i % NG
  const long long T = p.total_tiles;
  const int t0 = (int)(T * blockIdx.x / gridDim.x), t1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
  const int my_tiles = t1 - t0;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t bw = smem_u32(&bar_w);
      mbar_expect_tx(bw, KB * BLK);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * BLK, &maps.w, bw, kb * 64, 0);      // rows 0..127 of wqkv = Wq
      int cur_img = -1, n_img = 0, st = 0;
      uint32_t sph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = t0 + i, img = tile / p.tiles_per_image;
        // the x tile FIRST: the issuer runs GEMM 1 of tile i before GEMM 2 of tile i - 1, and only the latter releases M
        mbar_wait(smem_u32(&x_empty[st]), sph ^ 1u);
        const uint32_t full = smem_u32(&x_full[st]);
        mbar_expect_tx(full, KB * BLK);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(x_s + st * XBUF + kb * BLK, &maps.x, full, kb * 64, tile * TM);
        if (++st == NS) {
          st = 0;
          sph ^= 1u;
        }
        if (img != cur_img) {
          mbar_wait(smem_u32(&m_empty), (uint32_t)((n_img & 1) ^ 1));
          const uint32_t mf = smem_u32(&m_full);
          mbar_expect_tx(mf, 2 * MBLK);
          for (int kb = 0; kb < 2; ++kb) tma_load_2d(m_s + kb * MBLK, &maps.m, mf, kb * 64, img * C);
          cur_img = img;
          ++n_img;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(smem_u32(&bar_w), 0);
      tc_fence_after();
      int n_img = 0;
      auto gemm2 = [&](int j) {                     // o_j = Q_j M^T  (into the columns q_j occupied: it has been consumed)
        const int g = j % NG, tile = t0 + j, img = tile / p.tiles_per_image;
        if (j == 0 || (tile - 1) / p.tiles_per_image != img) {        // first tile of an image: its M must have landed
          mbar_wait(smem_u32(&m_full), (uint32_t)(n_img & 1));
          ++n_img;
        }
        mbar_wait(smem_u32(&q_ready[g]), (uint32_t)((j / NG) & 1));
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t adesc = desc_k(b_s + g * WBUF + kb * BLK), bdesc = desc_k(m_s + kb * MBLK);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + g * HID, adesc + 2ull * k, bdesc + 2ull * k, IDESC3, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&d3_full[g]));
        if (j == my_tiles - 1 || (tile + 1) / p.tiles_per_image != img) umma_commit(smem_u32(&m_empty));   // last tile of the image here
      };
      for (int i = 0; i < my_tiles; ++i) {
        const int g = i % NG;
        mbar_wait(smem_u32(&y_ready[g]), (uint32_t)((i / NG) & 1));
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t adesc = desc_k(b_s + g * WBUF + kb * BLK), bdesc = desc_k(w_s + kb * BLK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + g * HID, adesc + 2ull * k, bdesc + 2ull * k, IDESC1, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&d1_full[g]));
        if (i > 0) gemm2(i - 1);
      }
      if (my_tiles > 0) gemm2(my_tiles - 1);
    }
    __syncwarp();
  } else {
    const int grp = (warp - 2) / (4 * TPP), h2 = TPP == 2 ? ((warp - 2) >> 2) & 1 : 0, q = warp & 3, row = q * 32 + lane;
    const int gtid = threadIdx.x - 64 - grp * GT;             // index inside the group
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * HID);
    const uint32_t wb = b_s + grp * WBUF;
    for (int i = grp; i < my_tiles; i += NG) {
      const uint32_t ph = (uint32_t)((i / NG) & 1);
      const int st = i % NS;
      const uint32_t xb = x_s + st * XBUF;
      // the TMA store of this group's previous tile must have drained the work buffer before LayerNorm rewrites it
      if (gtid == 0) tma_store_wait_read<0>();
      named_bar_sync(1 + grp, GT);
      mbar_wait(smem_u32(&x_full[st]), (uint32_t)((i / NS) & 1));
      const uint32_t ex_mine = EXCH_IN_WB ? part_addr<C, TPP>(wb, row, h2, 0) : smem_u32(&exch[EXCH_STATIC ? grp : 0][EXCH_STATIC ? h2 * TM + row : 0]);
      const uint32_t ex_theirs = EXCH_IN_WB ? part_addr<C, TPP>(wb, row, h2 ^ 1, 0)
                                            : smem_u32(&exch[EXCH_STATIC ? grp : 0][EXCH_STATIC ? (h2 ^ 1) * TM + row : 0]);
      // (one static slot pair serves both normalisations of a tile: y_ready / d1_full lie between them, and those complete
      // only after all 256 threads have read their partner's value)
      ln_part<C, TPP, EXCH_IN_WB>(xb, wb, row, h2, ex_mine, ex_theirs, 1 + grp, p.eps);
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&y_ready[grp]));
      // ---- q -> softmax over the 32 channels of each of this thread's two heads -> Q block h2 (K-major)
      mbar_wait(smem_u32(&d1_full[grp]), ph);
      tc_fence_after();
      {
        // |q| is bounded by the same weight-only bound as k (<= 40): exp(q) stays inside fp32 without subtracting a maximum
        constexpr int NH = 4 / TPP;
        uint32_t r[2][32];
        tmem_ld32(lane_addr + (uint32_t)(h2 * NH * 32), r[0]);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          tmem_ld_wait();
          if (hh + 1 < NH) tmem_ld32(lane_addr + (uint32_t)((h2 * NH + hh + 1) * 32), r[(hh + 1) & 1]);
          const int gh = h2 * NH + hh;
          float e[32], s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = ex2f(__uint_as_float(r[hh & 1][j]));          // log2(e) is folded into the q rows of wqkv_g
            s4[j & 3] += e[j];
          }
          const float inv = __fdividef(1.0f, (s4[0] + s4[1]) + (s4[2] + s4[3]));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = e[c * 8 + j] * inv;
            sts128(sw_addr(wb + (gh >> 1) * BLK, row, (gh & 1) * 4 + c), pack8(f));
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&q_ready[grp]));
      // ---- o -> + bias -> LayerNorm over C -> * g_out + x -> bf16 staging tile -> TMA store
      mbar_wait(smem_u32(&d3_full[grp]), ph);
      tc_fence_after();
      float o[HC];
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) {
        uint32_t ro[32];
        tmem_ld32(lane_addr + (uint32_t)(h2 * HC + c0), ro);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o[c0 + j] = __uint_as_float(ro[j]) + bo_s[h2 * HC + c0 + j];
      }
      tc_fence_before();
      float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        s4[c & 3] += o[c];
        q4[c & 3] = fmaf(o[c], o[c], q4[c & 3]);
      }
      const float2 stt = ln_stats<C, TPP, EXCH_IN_WB>((s4[0] + s4[1]) + (s4[2] + s4[3]), (q4[0] + q4[1]) + (q4[2] + q4[3]), ex_mine, ex_theirs,
                                                 1 + grp, p.eps);
#pragma unroll
      for (int c = 0; c < HC / 8; ++c) {
        float res[8], f[8];
        unpack8(lds128(part_addr<C, TPP>(xb, row, h2, c)), res);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(fmaf(o[c * 8 + j], stt.x, stt.y), go_s[h2 * HC + c * 8 + j], res[j]);
        sts128(part_addr<C, TPP>(wb, row, h2, c), pack8(f));
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&x_empty[st]));                   // the residual has been read: the ring stage may be refilled
      named_bar_sync(1 + grp, GT);
      if (gtid == 0) {
        for (int kb = 0; kb < KB; ++kb) tma_store_2d(&maps.out, wb + kb * BLK, kb * 64, (t0 + i) * TM);
        tma_store_commit();
      }
    }
    if (gtid == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int C>
int launch_tc(const bf16* x, const bf16* wqkv, float shift_log2, const bf16* wout, const float* b_out, const float* g_out, bf16* out,
              float* workspace, int batch, int n, float scale, float eps, cudaStream_t s) {
  constexpr int KB = C / 64, N2 = C + 16;
  const int tpi = n / TM, sms = tedm_num_sms();
  // K-A work items: chunks of consecutive tiles of one image; the chunk length that wastes the least of the last wave
  int chunk = tpi, best_waste = 1 << 30;
  for (int c = tpi < 32 ? tpi : 32; c >= 4 && c >= tpi / 64; c >>= 1) {
    if (tpi % c) continue;
    const long long items = (long long)batch * (tpi / c), waves = (items + sms - 1) / sms;
    const int waste = (int)((waves * sms - items) * c + waves * 2);          // idle tile slots + a per-item drain cost
    if (waste < best_waste) {
      best_waste = waste;
      chunk = c;
    }
  }
  const int chunks = tpi / chunk;
  const long long items = (long long)batch * chunks;
  bf16* mimg = reinterpret_cast<bf16*>(workspace);                    // [batch][C][128] bf16 first (tests read it back) ...
  float* part = workspace + (size_t)batch * C * HID / 2;              // ... then the K-A partials [items][128][C + 16] fp32

  alignas(64) TcMaps maps;
  int rc = encode_2d(&maps.x, x, (long long)batch * n, C, TM);
  if (rc) return rc;
  rc = encode_2d(&maps.w, wqkv, 3 * HID, C, HID);
  if (rc) return rc;
  rc = encode_2d(&maps.m, mimg, (long long)batch * C, HID, C);
  if (rc) return rc;
  rc = encode_2d(&maps.out, out, (long long)batch * n, C, TM);
  if (rc) return rc;

  TcParams p{};
  p.n = n;
  p.tiles_per_image = tpi;
  p.chunk_tiles = chunk;
  p.items = (int)items;
  p.total_tiles = batch * tpi;
  p.eps = eps;
  p.shift_log2 = shift_log2;
  p.part = part;
  p.b_out = b_out;
  p.g_out = g_out;

  // measured on B200 (profiles/r02_linattn_tc.txt): at C = 64 three groups with one thread per pixel, at C = 128 (where
  // shared memory leaves room for two groups only) two threads per pixel
  constexpr int NG = C == 64 ? 3 : 2, TPP = C == 64 ? 1 : 2;
  const int smem_a = 1024 + KB * BLK + BLK + (C == 64 ? 6 : 3) * KB * BLK + NG * 2 * BLK;
  const int smem_b = 1024 + KB * BLK + 2 * C * 128 + (C == 64 ? (NG >= 3 ? 4 : 6) : 3) * KB * BLK + NG * 2 * BLK;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(linattn_tc_ctx_kernel<C, NG, TPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
    TEDM_CUDA(cudaFuncSetAttribute(linattn_tc_out_kernel<C, NG, TPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    configured = true;
  }
  const int grid_a = items < sms ? (int)items : sms;
  linattn_tc_ctx_kernel<C, NG, TPP><<<grid_a, 64 + NG * 128 * TPP, smem_a, s>>>(maps, p);
  TEDM_LAUNCH_CHECK();
  linattn_tc_combine_kernel<C><<<dim3((unsigned)batch, 4), 256, 0, s>>>(part, wqkv, wout, mimg, chunks, n, scale);
  TEDM_LAUNCH_CHECK();
  const int grid_b = p.total_tiles < sms ? p.total_tiles : sms;
  linattn_tc_out_kernel<C, NG, TPP><<<grid_b, 64 + NG * 128 * TPP, smem_b, s>>>(maps, p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

}  // namespace

extern "C" int tedm_linear_attention_tc_supported(int n, int channels, int heads, int dim_head) {
  return heads == 4 && dim_head == 32 && (channels == 64 || channels == 128) && n >= 512 && n % 512 == 0;
}

extern "C" int64_t tedm_linear_attention_tc_workspace(int batch, int n, int channels) {
  if (batch <= 0 || n <= 0 || n % TM) return -1;
  // partials: at most one per 4 tiles; + the per-image folded matrices M (bf16, counted in floats)
  const int64_t items = (int64_t)batch * ((n / TM + 3) / 4);
  return items * HID * (channels + 16) + (int64_t)batch * channels * HID / 2 + 64;
}

extern "C" int tedm_linear_attention_tc_fwd(const void* x, const void* wqkv_g, float shift_log2, const void* wout, const float* b_out,
                                            const float* g_out, void* out, float* workspace, int batch, int n, int channels,
                                            int heads, int dim_head, float scale, float eps, tedm_stream_t stream) {
  TEDM_CHECK_ARG(x && wqkv_g && wout && b_out && g_out && out && workspace && batch > 0 && batch <= 65535,
                 "tedm_linear_attention_tc_fwd: bad arguments");
  TEDM_CHECK_ARG(shift_log2 >= 0.0f && shift_log2 <= 60.0f, "tedm_linear_attention_tc_fwd: shift_log2=%f outside [0, 60]", shift_log2);
  TEDM_UNSUPPORTED(!tedm_linear_attention_tc_supported(n, channels, heads, dim_head),
                   "tedm_linear_attention_tc_fwd: n=%d channels=%d heads=%d dim_head=%d (needs 4 x 32 heads, 64 or 128 channels, "
                   "n a multiple of 512)", n, channels, heads, dim_head);
  TEDM_CHECK_ARG((long long)batch * n < 2147483647LL, "tedm_linear_attention_tc_fwd: too many pixels");
  cudaStream_t s = (cudaStream_t)stream;
  if (channels == 64)
    return launch_tc<64>((const bf16*)x, (const bf16*)wqkv_g, shift_log2, (const bf16*)wout, b_out, g_out, (bf16*)out, workspace,
                         batch, n, scale, eps, s);
  return launch_tc<128>((const bf16*)x, (const bf16*)wqkv_g, shift_log2, (const bf16*)wout, b_out, g_out, (bf16*)out, workspace,
                        batch, n, scale, eps, s);
}
