// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixel][cout] = sum_{tap, cin} A[pixel + tap offset][cin] * W[cout][tap][cin]
//
//   M = 128 output pixels per CTA (a tileB x tileH x tileW box of the NHWC output),
//   N = BN output channels (64/128/256), K = taps * (c0 + c1) in blocks of 64 channels.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM owner),
// warps 2..9 = epilogue (TMEM -> registers -> bias/residual/GroupNorm partials -> bf16 staging tile -> TMA store).
// A tiles are fetched by 5-D tiled TMA straight from the NHWC activation tensor: the box is
// (64 ch, tileW, 1, tileH, tileB) at a per-tap pixel offset, out-of-bounds pixels (the conv's zero
// padding, and batch overhang) are zero-filled by the TMA unit, and the 128B swizzle makes the
// landed box exactly the K-major UMMA operand layout.  The skip-connection concat of the UNet
// decoder is a second A tensor map (the K loop walks src0's channels, then src1's); the stride-2
// 4x4 Downsample reads a (2C, W/2, 2, H/2, B) view of the same memory so every tap is again a
// dense box; the nearest-x2 Upsample+3x3 runs as four parity 2x2 convs on the source resolution.
//
// What is in this file, in order:
//   conv_igemm_kernel<BN, AM, CPG, RES, CG>   the general kernel.  AM: 0 = one activation box per (tap, channel block);
//                                             1 = weight-stationary rows (3x3 into 64 channels, resident weights);
//                                             2 = halo tiles (3x3: one 10 x 18-pixel box serves all nine taps);
//                                             3 = halo tiles whose boxes are normalised in shared memory (fused GroupNorm).
//                                             CG = 2: CTA pairs (tcgen05 cta_group::2).  RES: residual / split epilogue,
//                                             incl. the residual tile by TMA and SiLU(GroupNorm(residual)) (ResnetBlock tail).
//   conv_ws4_kernel<CPG, CB, W64, XF>         3x3 into 64 channels on whole rows, four output rows per tile, the three
//                                             vertical taps of an input row in ONE N = 192 UMMA.
//   conv_wgrad_kernel / conv_wgrad3_kernel    weight gradients (generic; 3x3 halo-tile form) and their reduce kernels.
//   tedm_conv_igemm_fwd / _wgrad              host side: shape checks, tile geometry, kernel selection, tensor maps.
// Shared memory serves the tensor core's operand reads AND the TMA's writes (128 B/clk per SM): most of what distinguishes the
// variants is how many bytes of each a FLOP costs (DESIGN.md sections 3 and 8).
#include "common.cuh"

namespace {

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tensorMapEncodeTiled get_encode_fn() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tensorMapEncodeTiled)p;
  }
  return fn;
}

constexpr int BM = 128;          // pixels per tile (UMMA M)
constexpr int BK = 64;           // channels per k-block (128 bytes: one swizzle row)
constexpr int A_BYTES = BM * BK * 2;
constexpr int A_ROW_BYTES = 17 * 1024;   // row mode: 130 pixels x 128 B = 16640 B, padded to the 1 KB swizzle atom
constexpr int ROW_PIX = BM + 2;
constexpr int MAX_STAGES = 8;
constexpr int HALO_W = 8, HALO_H = 16;                         // halo mode: a tile is 8 pixels x 16 rows of one image
constexpr int HALO_PITCH = (HALO_W + 2) * BK * 2;              // bytes between image rows of the 10 x 18-pixel halo box
constexpr int HALO_TX = (HALO_W + 2) * (HALO_H + 2) * BK * 2;  // 23040 B
constexpr int HALO_BYTES = 23 * 1024;                          // padded to the 1 KB swizzle atom
constexpr int MAX_HALO_STAGES = 4;
constexpr int DYN_SMEM_MAX = 221 * 1024;  // + ~3.3 KB static <= 227 KB per CTA

constexpr int MAX_SRC = 6;

struct ConvMaps {                // every tensor map of a launch (one __grid_constant__ parameter: its elements are addressable)
  CUtensorMap a[MAX_SRC];
  CUtensorMap w, out, out2;
  CUtensorMap res;               // res_tma: the residual tensor, same geometry as `out`
};

struct ConvParams {
  int mode, taps, kxc;           // kxc = taps per filter row
  int c0_blocks, c1_blocks;      // weight-stationary row mode: the (at most two) sources it walks
  // A sources, walked in this order inside every tap (K order: tap-major, source, 64-channel block)
  int n_src;
  int src_blocks[MAX_SRC], src_C[MAX_SRC];
  int src_center[MAX_SRC];       // 1: feeds only the centre tap of a 3x3 (a folded 1x1 branch)
  int num_kb;                    // k-blocks per tile
  int cg;                        // host only: 2 = launch as CTA pairs (cta_group::2)
  int tileW, tileH, tileB, tiles_x, tiles_y;
  int B, Ho, Wo;                 // tile space extent (output pixels; source pixels for mode 3)
  int OH, OW, osy, osx;          // output tensor extent and tile->output coordinate scale
  int cout;
  int n_tiles, zdim, num_tiles, stages, out_bufs, a_stages;
  int z_shift, tpg_shift, tx_shift;   // log2 of zdim, tiles_x * tiles_y, tiles_x (all powers of two)
  const float* bias;
  const bf16* residual;
  bf16* out;
  int out_f32;                   // 1: `out` is fp32 (direct stores); 0: bf16 through smem + TMA store
  float* gn_partial;
  int gn_cpg, gn_groups, gn_parts;
  long long out_image_stride;    // elements
  // split output (data gradient of a skip-concat conv): channels [0, split) go to out/residual, channels
  // [split, cout) to out2/residual2 at channel (c - split); split == 0: single output
  int split;
  bf16* out2;
  const bf16* residual2;
  long long out2_image_stride;
  const float2* res_affine;      // [B][cout] (a/2, b/2): the residual enters as SiLU(a * r + b)  (tiles inside one image)
  const float* src_affine;       // AM = 3: [B][c0][2] (a/2, b/2): the conv's input is SiLU(a * src0 + b)
  int res_tma;                   // 1: the residual tile is TMA-loaded into the output staging tile one tile ahead (3 buffers)
};

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  // K-major, 128B swizzle: 8-row atoms of 1024 B (SBO), LBO unused, descriptor version 1 (sm_100).
  // The start address may sit on any 128-byte row (profiles/r01_umma_row_shift_probe.json): the swizzle
  // is a function of the absolute shared-memory address, so row-shifted windows of a TMA-written
  // buffer are valid operands with base_offset = 0.
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// the same with the 8-row groups `sbo_bytes` apart instead of packed (a tile whose rows of 8 pixels sit in a wider buffer)
__device__ __forceinline__ uint64_t make_sw128_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Sum NV per-thread values over the `width` (32 or 16) lanes of a segment with a transposing butterfly:
// every exchange halves the number of values a lane carries, so NV values cost NV-1 + log2(width/NV)
// shuffles instead of NV*log2(width).  Afterwards vals[0] of lane l holds the total of value index
// (l % width) >> (log2(width) - log2(NV)).
template <int NV>
__device__ __forceinline__ void butterfly_sum(float (&vals)[NV], int lane, int width) {
  int nv = NV;
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int off = width >> (step + 1);
    if (off == 0) break;
    if (nv > 1) {
      const int half = nv >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < half) {
          const float send = upper ? vals[i] : vals[i + half];
          const float keep = upper ? vals[i + half] : vals[i];
          vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      nv = half;
    } else {
      vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], off);
    }
  }
}

// SiLU(a * x + b) on the 8 bf16 values of a 16-byte chunk, (a, b) already halved (silu(z) = z/2 + z/2 * tanh(z/2)): the exact
// expression of gn_silu_kernel, so that a fused consumer and the separate pass round to the same bf16 values
__device__ __forceinline__ uint4 affine_silu8(const uint4& raw, const float (&ca)[8], const float (&cb)[8]) {
  float f[8];
  unpack8(raw, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = silu_half(fmaf(f[j], ca[j], cb[j]));
  return pack8(f);
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128_if(bool pred, uint32_t addr, const uint4& v) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %5, 0;\n @p st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n}" ::"r"(addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "r"((int)pred) : "memory");
}

struct TileCoord {
  int b0, y0, x0, n0, par, trem;
};
// `tile` counts the tiles of one CTA, or (CG = 2) of a CTA pair: the pair shares the N tile and the parity and takes two
// consecutive M tiles, one per CTA (`rank`).
template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile, int bn, int rank) {
  TileCoord t;
  // one real division (n_tiles may be 3, 6, 12); everything else is a power of two -- the producer thread decodes a
  // tile per handful of k-blocks on the 1x1 convolutions
  const int rest = tile / p.n_tiles;
  const int nt = tile - rest * p.n_tiles;
  t.par = rest & (p.zdim - 1);
  const int mt = (rest >> p.z_shift) * CG + rank;
  const int bg = mt >> p.tpg_shift;
  t.trem = mt & ((1 << p.tpg_shift) - 1);
  t.b0 = bg * p.tileB;
  t.y0 = (t.trem >> p.tx_shift) * p.tileH;
  t.x0 = (t.trem & (p.tiles_x - 1)) * p.tileW;
  t.n0 = nt * bn;
  return t;
}

// Persistent, warp-specialised: each CTA (one per SM) walks tiles blockIdx.x, +gridDim.x, ... .  The TMA
// producer runs ahead across tile boundaries through a ring of `stages` shared-memory slots; the MMA
// issuer alternates between two TMEM accumulators so the epilogue of tile i overlaps the main loop of
// tile i+1.
//   WS = false: slot = one (tap, 64-channel) k-block: A box 128 pixels + the matching weight tile.
//   WS = true  (3x3, full 128-pixel rows, cout = 64, <= 128 input channels): all 9 x cin weights are
//               loaded ONCE per CTA and stay in shared memory; a slot is one 130-pixel input row
//               (x0-1 .. x0+128) of one 64-channel block, and the three horizontal taps are three
//               UMMAs on row-shifted windows of it -> activations cross L2->SM 3x instead of 9x and
//               weights once instead of once per tile.
constexpr int EPI_THREADS = 256;   // 8 epilogue warps
constexpr int CONV_THREADS = 64 + EPI_THREADS;
constexpr int XF_THREADS = 128;    // AM = 3: four more warps that normalise the halo boxes in place

//   CG = 2: CTA PAIRS (cta_group::2, cluster of two SMs of one TPC).  One tcgen05.mma spans both SMs: M = 256 = the two
//               CTAs' 128-pixel tiles, each CTA feeds its own A rows and HALF of the weight tile (N / 2 rows), the
//               hardware shares the halves.  Per SM and UMMA the shared-memory read drops from A 4 KB + B (BN / 32) KB to
//               A 4 KB + B (BN / 64) KB -- what bounds the N = 64 / 128 tiles.  Only the leader CTA (rank 0) issues MMAs;
//               both issue TMA, all transaction bytes land on the leader's barriers; tcgen05.commit multicasts to both.
//   AM = 2 (HALO; 3x3 on images of at least 16 rows): a tile is 8 pixels x 16 rows, and its 10 x 18-pixel halo box of one
//               64-channel block is loaded ONCE and serves all nine taps: tap (ky, kx) is the window that starts (ky, kx)
//               pixels into the box, its 8-pixel rows 1280 B apart (the descriptor's stride between 8-row groups) instead of
//               packed.  Shared memory carries the tensor core's operand reads AND the TMA's writes; this removes 8/9 of
//               the activation writes (a slot of the main ring then holds only the weight tile of one (tap, channel block)).
//   AM = 3 (halo mode with a fused input transform): the conv's input is SiLU(a * x + b) of the tensor the boxes are loaded
//               from, (a, b) per (image, channel) from tedm_gn_affine -- Block's GroupNorm + scale/shift + SiLU
//               (models/unet_model.py:128-134) applied to the halo box in shared memory by four extra warps, between the
//               TMA's arrival and the tensor core's reads, so the normalised activation never exists in HBM.  A halo box is
//               1.4 tiles of pixels and is loaded once, so the transform costs 1.4 MUFU per element (a per-tap A tile
//               would need 9).  Pixels outside the image stay the TMA's zero fill (the conv pads the ACTIVATION).
template <int BN, int AM, int CPG, bool RES, int CG>
__global__ void __launch_bounds__(AM == 3 ? CONV_THREADS + XF_THREADS : CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p) {
  constexpr bool WS = AM == 1, HALO = AM >= 2, XF = AM == 3;
  const CUtensorMap& mapW = maps.w;
  const CUtensorMap& mapOut = maps.out;
  const CUtensorMap& mapOut2 = maps.out2;
  constexpr int BNC = BN / CG;                     // weight rows (output channels) this CTA stages per tile
  constexpr int B_BYTES = BNC * BK * 2;
  constexpr int STAGE_BYTES = WS ? A_ROW_BYTES : (HALO ? B_BYTES : A_BYTES + B_BYTES);
  constexpr int OUT_BYTES = (BN / 64) * A_BYTES;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ __align__(8) uint64_t bar_afull[MAX_HALO_STAGES];
  __shared__ __align__(8) uint64_t bar_aempty[MAX_HALO_STAGES];
  __shared__ __align__(8) uint64_t bar_aready[MAX_HALO_STAGES];   // AM = 3: box normalised (the pair's leader hears both CTAs)
  __shared__ __align__(8) uint64_t bar_res[3];                    // res_tma: residual tile landed in staging buffer i
  __shared__ uint32_t tmem_slot;
  __shared__ float red[2][4][2][4][2];   // [column half][lane quarter][segment][group][sum, sumsq]
  __shared__ __align__(16) float sbias[256];
  __shared__ __align__(16) float2 saff[RES ? 256 : 1];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb_total = p.c0_blocks + p.c1_blocks;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_bytes = WS ? (uint32_t)(9 * cb_total * B_BYTES) : 0u;
  const uint32_t stage_base = smem_base + w_bytes;
  const uint32_t halo_base = stage_base + (uint32_t)p.stages * STAGE_BYTES;
  const uint32_t out_base = halo_base + (HALO ? (uint32_t)p.a_stages * HALO_BYTES : 0u);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_src; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&mapW);
    if (!p.out_f32) tma_prefetch_desc(&mapOut);
    if (p.split) tma_prefetch_desc(&mapOut2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), CG * (EPI_THREADS / 32));     // the leader's copy hears both CTAs' epilogues
    }
    mbar_init(smem_u32(&bar_w), 1);
    for (int i = 0; i < MAX_HALO_STAGES; ++i) {
      mbar_init(smem_u32(&bar_afull[i]), 1);
      mbar_init(smem_u32(&bar_aempty[i]), 1);
      mbar_init(smem_u32(&bar_aready[i]), CG * (XF_THREADS / 32));
    }
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bar_res[i]), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair<2 * BN>(smem_u32(&tmem_slot));
    else tmem_alloc<2 * BN>(smem_u32(&tmem_slot));
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();        // the peer's barriers exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG, tile_end = p.num_tiles / CG;
  // halo mode on the folded upsample conv (mode 3): the four parity tiles of an M tile read the SAME 10 x 18 box (their 2 x 2
  // taps are the windows (ky + parity_y, kx + parity_x) of it), so a CTA takes them in a row -- with one N tile they are
  // consecutive tile indices -- and the boxes of all channel blocks stay resident for the four of them
  const bool quad = HALO && p.mode == 3;
  auto seq_tile = [&](int k) -> int {        // the k-th tile of this CTA (pair), or -1
    if (quad) {
      const int g4 = (tile0 + (k >> 2) * tile_step) * 4;
      return g4 < tile_end ? g4 + (k & 3) : -1;
    }
    const int tl = tile0 + k * tile_step;
    return tl < tile_end ? tl : -1;
  };

  if (warp == 0) {
    // ===== TMA producer =====================================================================
    if (elect_one()) {
      int stage = 0, astage = 0;
      uint32_t phase = 0, aphase = 0;
      const int cb_all = p.num_kb / p.taps;        // 64-channel blocks over all sources (halo mode: no centre-only sources)
      // CG = 2: both CTAs load (their A rows, their half of the weights); the bytes are counted on the LEADER's barrier,
      // which only the leader arms
      auto load5 = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
        if constexpr (CG == 2) tma_load_5d_pair(dst, m, bar, c0, c1, c2, c3, c4);
        else tma_load_5d(dst, m, bar, c0, c1, c2, c3, c4);
      };
      auto load2 = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
        if constexpr (CG == 2) tma_load_2d_pair(dst, m, bar, c0, c1);
        else tma_load_2d(dst, m, bar, c0, c1);
      };
      int h_tile = tile0, h_cbg = 0;                // AM = 3: the next box to request
      auto request_halo = [&]() {
        if (h_tile >= tile_end) return;
        const TileCoord th = decode_tile<CG>(p, h_tile, BN, rank);
        mbar_wait(smem_u32(&bar_aempty[astage]), aphase ^ 1u);
        const uint32_t afull = smem_u32(&bar_afull[astage]);
        mbar_expect_tx(afull, HALO_TX);              // each CTA's transform warps wait for their OWN box: local barrier
        tma_load_5d(halo_base + astage * HALO_BYTES, &maps.a[0], afull, h_cbg * BK, th.x0 - 1, 0, th.y0 - 1, th.b0);
        if (++astage == p.a_stages) {
          astage = 0;
          aphase ^= 1u;
        }
        if (++h_cbg == cb_all) {
          h_cbg = 0;
          h_tile += tile_step;
        }
      };
      if constexpr (XF) {
        request_halo();
        request_halo();
      }
      if (WS) {
        const uint32_t bw = smem_u32(&bar_w);
        if (rank == 0) mbar_expect_tx(bw, CG * w_bytes);
        for (int i = 0; i < 9 * cb_total; ++i) load2(smem_base + (uint32_t)i * B_BYTES, &mapW, bw, i * BK, rank * BNC);
      }
      for (int k = 0, tile; (tile = seq_tile(k)) >= 0; ++k) {
        const TileCoord t = decode_tile<CG>(p, tile, BN, rank);
        if (WS) {
          for (int dy = 0; dy < 3; ++dy)
            for (int cb = 0; cb < cb_total; ++cb) {
              mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
              const uint32_t full = smem_u32(&bar_full[stage]);
              if (rank == 0) mbar_expect_tx(full, CG * ROW_PIX * BK * 2);
              const bool second = cb >= p.c0_blocks;
              load5(stage_base + stage * STAGE_BYTES, &maps.a[second ? 1 : 0], full,
                    (second ? cb - p.c0_blocks : cb) * BK, t.x0 - 1, 0, t.y0 + dy - 1, t.b0);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
        } else if (XF) {
          // boxes are requested TWO blocks ahead of the weight tiles, across tile boundaries (request_halo): a box has to
          // land and be normalised before the tensor core reaches its block
          for (int cbg = 0; cbg < cb_all; ++cbg)
            for (int tap = 0; tap < 9; ++tap) {
              if (tap == 0) request_halo();
              mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
              const uint32_t full = smem_u32(&bar_full[stage]);
              if (rank == 0) mbar_expect_tx(full, CG * STAGE_BYTES);
              load2(stage_base + stage * STAGE_BYTES, &mapW, full, (tap * cb_all + cbg) * BK, t.n0 + rank * BNC);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
        } else if (HALO) {
          // per 64-channel block: the halo box once, then the weight tiles of its taps; the next block's box is requested
          // after the second weight tile, so that it lands while this block is multiplied
          auto load_halo = [&](int si, int cblk) {
            mbar_wait(smem_u32(&bar_aempty[astage]), aphase ^ 1u);
            const uint32_t afull = smem_u32(&bar_afull[astage]);
            if (rank == 0) mbar_expect_tx(afull, CG * HALO_TX);
            load5(halo_base + astage * HALO_BYTES, &maps.a[si], afull, cblk * BK, t.x0 - 1, 0, t.y0 - 1, t.b0);
            if (++astage == p.a_stages) {
              astage = 0;
              aphase ^= 1u;
            }
          };
          const int ntap = p.taps;                    // 9, or the 2 x 2 of one parity of the folded upsample conv
          if (quad) {
            if ((k & 3) == 0) {                       // the boxes of every channel block, kept for the four parities
              int si = 0, cblk = 0;
              for (int cbg = 0; cbg < cb_all; ++cbg) {
                load_halo(si, cblk);
                if (++cblk == p.src_blocks[si]) {
                  cblk = 0;
                  ++si;
                }
              }
            }
          } else {
            load_halo(0, 0);
          }
          int si = 0, cblk = 0;
          for (int cbg = 0; cbg < cb_all; ++cbg) {
            int nsi = si, ncblk = cblk + 1;            // the block after this one
            if (ncblk == p.src_blocks[si]) {
              ncblk = 0;
              ++nsi;
            }
            for (int tap = 0; tap < ntap; ++tap) {
              if (!quad && tap == 2 && cbg + 1 < cb_all) load_halo(nsi, ncblk);
              mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
              const uint32_t full = smem_u32(&bar_full[stage]);
              if (rank == 0) mbar_expect_tx(full, CG * STAGE_BYTES);
              load2(stage_base + stage * STAGE_BYTES, &mapW, full, (tap * cb_all + cbg) * BK, t.par * p.cout + t.n0 + rank * BNC);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
            si = nsi;
            cblk = ncblk;
          }
        } else {
          const int par_y = t.par >> 1, par_x = t.par & 1;
          // nested counters instead of kb / cb_total etc.: this single thread paces the whole pipeline, and at N = 64 a
          // k-block is only ~150 tensor-core cycles
          int kb = 0;
          for (int ky = 0; ky < p.kxc; ++ky)
            for (int kx = 0; kx < p.kxc; ++kx) {
              int offy = 0, offx = 0, pc = 0, xsel = 0;
              if (p.mode == 1) {
                offy = ky - 1;
                offx = kx - 1;
              } else if (p.mode == 2) {
                offy = ((ky + 1) >> 1) - 1;
                offx = ((kx + 1) >> 1) - 1;
                pc = (ky + 1) & 1;
                xsel = (kx + 1) & 1;
              } else if (p.mode == 3) {
                offy = ky - 1 + par_y;
                offx = kx - 1 + par_x;
              }
              const bool centre = p.mode == 1 && ky == 1 && kx == 1;
              for (int si = 0; si < p.n_src; ++si) {
                if (p.src_center[si] && !centre) continue;
                const CUtensorMap* mapA = &maps.a[si];
                const int chan_off = xsel * p.src_C[si];
                for (int cblk = 0; cblk < p.src_blocks[si]; ++cblk, ++kb) {
                  mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
                  const uint32_t full = smem_u32(&bar_full[stage]);
                  if (rank == 0) mbar_expect_tx(full, CG * STAGE_BYTES);
                  const uint32_t a_dst = stage_base + stage * STAGE_BYTES;
                  load5(a_dst, mapA, full, chan_off + cblk * BK, t.x0 + offx, pc, t.y0 + offy, t.b0);
                  load2(a_dst + A_BYTES, &mapW, full, kb * BK, t.par * p.cout + t.n0 + rank * BNC);
                  if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                  }
                }
              }
            }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (CG = 2: the leader CTA's, for the pair) ================================
    auto umma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
      if constexpr (CG == 2) umma_bf16_pair(d, a, b, IDESC, acc);
      else umma_bf16(d, a, b, IDESC, acc);
    };
    auto commit = [&](uint32_t bar) {
      if constexpr (CG == 2) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    if (rank == 0 && elect_one()) {
      int stage = 0, it = 0, astage = 0;
      uint32_t phase = 0, aphase = 0;
      if (WS) {
        mbar_wait(smem_u32(&bar_w), 0);
        tc_fence_after();
      }
      int g_astage = 0;
      uint32_t g_aphase = 0;
      for (int tile; (tile = seq_tile(it)) >= 0; ++it) {
        const int buf = it & 1;
        mbar_wait(smem_u32(&bar_acc_empty[buf]), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        if (WS) {
          for (int dy = 0; dy < 3; ++dy)
            for (int cb = 0; cb < cb_total; ++cb) {
              mbar_wait(smem_u32(&bar_full[stage]), phase);
              tc_fence_after();
              const uint32_t a_addr = stage_base + stage * STAGE_BYTES;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint64_t adesc = make_sw128_desc(a_addr + (uint32_t)dx * 128u);
                const uint64_t bdesc = make_sw128_desc(smem_base + (uint32_t)(((dy * 3 + dx) * cb_total + cb) * B_BYTES));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma(d_tmem, adesc + 2ull * k, bdesc + 2ull * k, (dy | cb | dx | k) != 0 ? 1u : 0u);
              }
              commit(smem_u32(&bar_empty[stage]));
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
        } else if (HALO) {
          const int cb_all = p.num_kb / p.taps;
          int par_y = 0, par_x = 0;
          if (quad) {
            // the same boxes for the four parities: walk the group's slots again from where its first tile found them
            if ((it & 3) == 0) {
              g_astage = astage;
              g_aphase = aphase;
            } else {
              astage = g_astage;
              aphase = g_aphase;
            }
            const int par = tile & 3;                 // one N tile: the parity is the fastest tile index
            par_y = par >> 1;
            par_x = par & 1;
          }
          for (int cbg = 0; cbg < cb_all; ++cbg) {
            mbar_wait(smem_u32(XF ? &bar_aready[astage] : &bar_afull[astage]), aphase);
            tc_fence_after();
            const uint32_t h_addr = halo_base + astage * HALO_BYTES;
            // one weight slot = the four UMMAs of a tap; this thread paces the tensor core
            auto tap_mma = [&](int tap, int wy, int wx) {
              mbar_wait(smem_u32(&bar_full[stage]), phase);
              tc_fence_after();
              const uint64_t adesc = make_sw128_desc_sbo(h_addr + (uint32_t)(wy * HALO_PITCH + wx * BK * 2), HALO_PITCH);
              const uint64_t bdesc = make_sw128_desc(stage_base + stage * STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma(d_tmem, adesc + 2ull * k, bdesc + 2ull * k, (cbg | tap | k) != 0 ? 1u : 0u);
              commit(smem_u32(&bar_empty[stage]));
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            };
            if (quad) {
              for (int tap = 0; tap < 4; ++tap) tap_mma(tap, (tap >> 1) + par_y, (tap & 1) + par_x);
            } else if constexpr (BN == 256) {
              // measured (same box, alternating libraries): the rolled loop runs 512 -> 512 @16x16 at 1480 TFLOP/s, the
              // unrolled one at 1320; the N <= 128 layers prefer the unrolled loop (1040-1095 against 1011-1022)
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) tap_mma(tap, tap / 3, tap % 3);
            } else {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) tap_mma(tap, tap / 3, tap % 3);
            }
            if (!quad || (it & 3) == 3) commit(smem_u32(&bar_aempty[astage]));
            if (++astage == p.a_stages) {
              astage = 0;
              aphase ^= 1u;
            }
          }
        } else {
          const int num_kb = p.num_kb;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = stage_base + stage * STAGE_BYTES;
            const uint64_t adesc = make_sw128_desc(a_addr), bdesc = make_sw128_desc(a_addr + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma(d_tmem, adesc + 2ull * k, bdesc + 2ull * k, (kb | k) != 0 ? 1u : 0u);
            commit(smem_u32(&bar_empty[stage]));
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        commit(smem_u32(&bar_acc_full[buf]));
      }
    }
    __syncwarp();
  } else if (XF && warp >= CONV_THREADS / 32) {
    // ===== input transform (AM = 3): 128 threads; thread = (16-byte channel chunk, pixel row mod 16) ========
    if constexpr (XF) {
      const int tt = threadIdx.x - CONV_THREADS;
      const int c8 = tt & 7, prow = tt >> 3;
      const int cb_all = p.num_kb / p.taps;
      const int csrc = p.src_C[0];
      int astage = 0;
      uint32_t aphase = 0;
      for (int tile = tile0; tile < tile_end; tile += tile_step) {
        const TileCoord t = decode_tile<CG>(p, tile, BN, rank);
        for (int cbg = 0; cbg < cb_all; ++cbg) {
          // (a / 2, b / 2) of this thread's 8 channels of image b0 (requested before the box is waited for)
          const float4* ap = reinterpret_cast<const float4*>(p.src_affine + 2 * ((size_t)t.b0 * csrc + cbg * BK + c8 * 8));
          float ca[8], cb[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = __ldg(ap + j);
            ca[2 * j] = v.x;
            cb[2 * j] = v.y;
            ca[2 * j + 1] = v.z;
            cb[2 * j + 1] = v.w;
          }
          mbar_wait(smem_u32(&bar_afull[astage]), aphase);
          const uint32_t box = halo_base + astage * HALO_BYTES;
          // pixel px = prow + 16 i of the 10 x 18 box: (px & 7) and with it the swizzled chunk position never change.
          // Straight-line code, six chunks at a time: all loads first (shared memory is busy feeding the tensor core, a
          // dependent load costs hundreds of cycles), predicated stores instead of branches.
          constexpr int NPX = (HALO_W + 2) * (HALO_H + 2), NIT = (NPX + 15) / 16;
          const uint32_t addr0 = box + (uint32_t)prow * 128u + ((uint32_t)(c8 ^ (prow & 7)) << 4);
          int hy = prow / (HALO_W + 2), hx = prow - hy * (HALO_W + 2);
#pragma unroll
          for (int i0 = 0; i0 < NIT; i0 += 6) {
            uint4 raw[6];
            bool inside[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
              const int i = i0 + u;
              const int iy = t.y0 - 1 + hy, ix = t.x0 - 1 + hx;
              inside[u] = i < NIT && prow + 16 * i < NPX && iy >= 0 && iy < p.Ho && ix >= 0 && ix < p.Wo;
              hx += 6;                                 // 16 pixels further = one row and six pixels
              hy += 1;
              if (hx >= HALO_W + 2) {
                hx -= HALO_W + 2;
                hy += 1;
              }
              raw[u] = lds128(addr0 + (uint32_t)(inside[u] ? i : 0) * 2048u);
            }
#pragma unroll
            for (int u = 0; u < 6; ++u)                // pixels outside the image keep the TMA's zero fill
              sts128_if(inside[u], addr0 + (uint32_t)(inside[u] ? i0 + u : 0) * 2048u, affine_silu8(raw[u], ca, cb));
          }
          fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's (async proxy) reads
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_leader(smem_u32(&bar_aready[astage]));
            else mbar_arrive(smem_u32(&bar_aready[astage]));
          }
          if (++astage == p.a_stages) {
            astage = 0;
            aphase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp w reads TMEM lane quarter (w & 3) and half of the tile's columns ==========
    constexpr int CH = BN / 64;                 // 32-column chunks per warp
    constexpr int WCOLS = BN / 2;               // columns per warp
    constexpr int NGL = CPG ? BN / CPG : 1;     // GroupNorm groups covered by this N tile (<= 8)
    constexpr bool SPLITG = CPG > WCOLS;        // one group spans both column halves
    constexpr int NGLW = CPG ? (SPLITG ? 1 : WCOLS / CPG) : 1;   // groups per warp
    constexpr int NV = 2 * NGLW;                // (sum, sum of squares) per group
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int hsel = (warp - 2) >> 2;           // column half
    const int row = q * 32 + lane;
    const int e = threadIdx.x - 64;             // 0..255
    const int hwt = p.tileW * p.tileH;          // pixels of one image inside the tile
    const int tb = row / hwt, rrem = row % hwt;
    const int ty = rrem / p.tileW, tx = rrem % p.tileW;
    const int seg_size = hwt < 32 ? hwt : 32;
    // the residual rows of a thread for one tile.  With a split output, 64-channel sub-tile cp belongs to
    // (residual, `split` channels per pixel) or (residual2, cout - split channels per pixel).
    auto fetch_res = [&](const TileCoord& tt, uint4* dst, bool* has) {
      const int b_ = tt.b0 + tb, y_ = tt.y0 + ty, x_ = tt.x0 + tx;
      const size_t pixoff_ = (size_t)(y_ * p.osy + (tt.par >> 1)) * p.OW + (x_ * p.osx + (tt.par & 1));
      const size_t opix_ = (size_t)b_ * p.out_image_stride + pixoff_ * p.cout + tt.n0;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int chunk = hsel * CH + i, c = tt.n0 + (chunk >> 1) * 64;
        const bf16* rsub;
        if (p.split == 0) rsub = p.residual ? p.residual + opix_ + (chunk >> 1) * 64 : nullptr;
        else if (c < p.split) rsub = p.residual ? p.residual + (size_t)b_ * p.out_image_stride + pixoff_ * p.split + c : nullptr;
        else rsub = p.residual2 ? p.residual2 + (size_t)b_ * p.out2_image_stride + pixoff_ * (p.cout - p.split) + (c - p.split) : nullptr;
        has[i] = rsub != nullptr && b_ < p.B;
        if (has[i]) {
          const uint4* rp = reinterpret_cast<const uint4*>(rsub + (chunk & 1) * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[i * 4 + j] = __ldg(rp + j);
        }
      }
    };
    // N tiles <= 128: the residual rows, bias and residual affine of tile i+1 are requested while tile i is processed -- the
    // 1x1 convolutions have no main loop to hide a global-memory round trip behind
    constexpr bool PIPE = RES && CH <= 2;
    uint4 rnext[PIPE ? CH * 4 : 1];
    bool hnext[PIPE ? CH : 1];
    float bias_next = 0.0f;
    float2 aff_next = make_float2(0.0f, 0.0f);
    // the bias (and residual affine) of the NEXT tile are requested while this one is processed, in every variant: a global
    // round trip at the head of each tile's epilogue is what bounds the tiles with short K loops (1x1, stride-2, upsample)
    auto fetch_consts = [&](const TileCoord& tt) {
      if (e < BN) {
        bias_next = p.bias ? __ldg(p.bias + tt.n0 + e) : 0.0f;
        if constexpr (RES) {
          if (p.res_affine) aff_next = __ldg(p.res_affine + (size_t)tt.b0 * p.cout + tt.n0 + e);
        }
      }
    };
    // res_tma (1x1 convs with a residual): the residual tile of tile i + 1 is TMA-loaded into the staging buffer tile i + 1 will
    // leave through -- coalesced, no registers, a whole tile of lead time -- and added in place; three staging buffers, so the
    // buffer being refilled is the one whose store (tile i - 2) has certainly been read
    const bool res_tma = RES && p.res_tma;
    auto load_res_tile = [&](const TileCoord& tt, int slot) {
      const uint32_t bar = smem_u32(&bar_res[slot]);
      mbar_expect_tx(bar, OUT_BYTES);
      for (int sidx = 0; sidx < BN / 64; ++sidx)
        tma_load_5d(out_base + (uint32_t)(slot * OUT_BYTES + sidx * A_BYTES), &maps.res, bar, tt.n0 + sidx * 64, tt.x0, 0, tt.y0, tt.b0);
    };
    if (seq_tile(0) >= 0) {
      const TileCoord t0 = decode_tile<CG>(p, seq_tile(0), BN, rank);
      if constexpr (PIPE) {
        if (!res_tma) fetch_res(t0, rnext, hnext);
        else if (e == 0) load_res_tile(t0, 0);
      }
      fetch_consts(t0);
    }
    int it = 0, obuf = 0;                       // obuf = it % out_bufs
    for (int tile; (tile = seq_tile(it)) >= 0; ++it, obuf = obuf + 1 == p.out_bufs ? 0 : obuf + 1) {
      const TileCoord t = decode_tile<CG>(p, tile, BN, rank);
      const int par_y = t.par >> 1, par_x = t.par & 1;
      const int b = t.b0 + tb, y = t.y0 + ty, x = t.x0 + tx;
      const bool valid = b < p.B;
      const int oy = y * p.osy + par_y, ox = x * p.osx + par_x;
      const size_t pixoff = (size_t)oy * p.OW + ox;
      const size_t opix = (size_t)b * p.out_image_stride + pixoff * p.cout + t.n0;
      const int buf = it & 1;
      const uint32_t out_buf = out_base + (uint32_t)(obuf * OUT_BYTES);
      const int obuf_next = obuf + 1 == p.out_bufs ? 0 : obuf + 1;
      // the TMA store that last used this staging buffer must have drained it; every thread must be done
      // with the previous tile's `red` / `sbias` before this tile overwrites them (barrier below)
      if (e == 0 && !p.out_f32) {
        if (p.out_bufs >= 2) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
        if constexpr (PIPE) {
          // at most the store of tile i - 1 is still reading: buffer (i + 1) % 3, last used by tile i - 2, is free
          if (res_tma && seq_tile(it + 1) >= 0) load_res_tile(decode_tile<CG>(p, seq_tile(it + 1), BN, rank), obuf_next);
        }
      }
      if (e < BN) {
        sbias[e] = bias_next;
        if constexpr (RES) saff[e] = aff_next;
      }
      named_bar_sync(1, EPI_THREADS);
      // without the one-tile-ahead pipeline the residual values are fetched BEFORE waiting for the accumulator: their
      // latency hides behind the main loop of this tile
      uint4 rpre[RES ? CH * 4 : 1];
      bool has_res[RES ? CH : 1];
      if constexpr (PIPE) {
#pragma unroll
        for (int i = 0; i < CH * 4; ++i) rpre[i] = rnext[i];
#pragma unroll
        for (int i = 0; i < CH; ++i) has_res[i] = hnext[i];
        if (seq_tile(it + 1) >= 0) {
          const TileCoord tn = decode_tile<CG>(p, seq_tile(it + 1), BN, rank);
          if (!res_tma) fetch_res(tn, rnext, hnext);
          fetch_consts(tn);
        }
      } else {
        if constexpr (RES) fetch_res(t, rpre, has_res);
        if (seq_tile(it + 1) >= 0) fetch_consts(decode_tile<CG>(p, seq_tile(it + 1), BN, rank));
      }
      mbar_wait(smem_u32(&bar_acc_full[buf]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if constexpr (PIPE) {
        if (res_tma) mbar_wait(smem_u32(&bar_res[obuf]), (uint32_t)((it / 3) & 1));
      }
      float gv[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) gv[i] = 0.0f;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int chunk = hsel * CH + i, cp = chunk >> 1, half = chunk & 1;
        uint32_t r0[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + chunk * 32), r0);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = *reinterpret_cast<const float4*>(&sbias[chunk * 32 + 4 * j]);
          v[4 * j] = __uint_as_float(r0[4 * j]) + bv.x;
          v[4 * j + 1] = __uint_as_float(r0[4 * j + 1]) + bv.y;
          v[4 * j + 2] = __uint_as_float(r0[4 * j + 2]) + bv.z;
          v[4 * j + 3] = __uint_as_float(r0[4 * j + 3]) + bv.w;
        }
        if (CPG) {
          constexpr int W = CPG < 32 ? (CPG ? CPG : 32) : 32;   // channels per partial sum inside this chunk
#pragma unroll
          for (int sg = 0; sg < 32 / W; ++sg) {
            float s_ = 0.0f, q_ = 0.0f;
#pragma unroll
            for (int j = 0; j < W; ++j) {
              s_ += v[sg * W + j];
              q_ = fmaf(v[sg * W + j], v[sg * W + j], q_);
            }
            const int gl = SPLITG ? 0 : (i * 32 + sg * W) / (CPG ? CPG : 1);   // group index local to this warp
            gv[2 * gl] += s_;
            gv[2 * gl + 1] += q_;
          }
        }
        if constexpr (RES) {
          if (res_tma || has_res[i]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
              if (res_tma) {                       // the residual sits where this thread's result will go
                const uint32_t c16 = (uint32_t)(half * 4 + j);
                unpack8(lds128(out_buf + (uint32_t)cp * A_BYTES + (uint32_t)row * 128u + ((c16 ^ (uint32_t)(row & 7)) << 4)), f);
              } else {
                unpack8(rpre[i * 4 + j], f);
              }
              if (p.res_affine) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float2 ab = saff[chunk * 32 + 8 * j + c];
                  f[c] = silu_half(fmaf(f[c], ab.x, ab.y));
                }
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) v[8 * j + c] += f[c];
            }
          }
        }
        if (p.out_f32) {
          if (valid) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix + chunk * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          // stage the bf16 tile in the 128B-swizzled layout the TMA store expects: sub-tile = 64 channels,
          // row = pixel (128 B), 16-byte chunk index XOR (row & 7)
          const uint32_t sub = out_buf + (uint32_t)cp * A_BYTES + (uint32_t)row * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t c16 = (uint32_t)(half * 4 + j);
            const uint32_t dst = sub + ((c16 ^ (uint32_t)(row & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])),
                         "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                         "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7])) : "memory");
          }
        }
      }
      // this accumulator buffer is free for the MMA issuer again
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_leader(smem_u32(&bar_acc_empty[buf]));
        else mbar_arrive(smem_u32(&bar_acc_empty[buf]));
      }
      if (CPG) {
        butterfly_sum<NV>(gv, lane, seg_size);
        const int ls = lane & (seg_size - 1);
        const int keep_shift = (seg_size == 32 ? 5 : 4) - (NV == 8 ? 3 : NV == 4 ? 2 : 1);
        if ((ls & ((1 << keep_shift) - 1)) == 0) (&red[hsel][q][lane / seg_size][0][0])[ls >> keep_shift] = gv[0];
      }
      if (!p.out_f32) fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(2, EPI_THREADS);
      if (e == 0 && !p.out_f32) {
        const int chan_base = (p.mode == 3 ? par_x * p.cout : 0) + t.n0;
        const int pc_out = p.mode == 3 ? par_y : 0;
        for (int sidx = 0; sidx < BN / 64; ++sidx) {
          const int c = chan_base + sidx * 64;
          if (RES && p.split && c >= p.split)
            tma_store_5d(&mapOut2, out_buf + (uint32_t)sidx * A_BYTES, c - p.split, t.x0, pc_out, t.y0, t.b0);
          else
            tma_store_5d(&mapOut, out_buf + (uint32_t)sidx * A_BYTES, c, t.x0, pc_out, t.y0, t.b0);
        }
        tma_store_commit();
      }
      if (CPG) {
        const int segs_per_img = hwt / seg_size;      // 4, 2 or 1
        const int part = hwt == BM ? t.trem : 0;
        for (int o = e; o < p.tileB * NGL; o += EPI_THREADS) {
          const int img = o / NGL, gl = o % NGL;
          if (t.b0 + img >= p.B) continue;
          float s_ = 0.0f, q_ = 0.0f;
          for (int sgi = 0; sgi < segs_per_img; ++sgi) {
            const int gseg = img * segs_per_img + sgi;       // global segment index in the tile
            const int w = (gseg * seg_size) >> 5, sl = ((gseg * seg_size) & 31) / seg_size;
            if (SPLITG) {
              s_ += red[0][w][sl][0][0] + red[1][w][sl][0][0];
              q_ += red[0][w][sl][0][1] + red[1][w][sl][0][1];
            } else {
              s_ += red[gl / NGLW][w][sl][gl % NGLW][0];
              q_ += red[gl / NGLW][w][sl][gl % NGLW][1];
            }
          }
          float* dst = p.gn_partial + (((size_t)(t.b0 + img) * p.gn_parts + part) * p.gn_groups + t.n0 / (CPG ? CPG : 1) + gl) * 2;
          dst[0] = s_;
          dst[1] = q_;
        }
      }
    }
    if (e == 0 && !p.out_f32) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();        // neither CTA leaves (or frees TMEM) while the pair's MMAs can still touch it
  if (warp == 1) {
    if constexpr (CG == 2) tmem_dealloc_pair<2 * BN>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

int encode_act_map(CUtensorMap* map, const void* ptr, int B, int H, int W, int C, long long image_stride, int mode,
                   int tileW, int tileH, int tileB) {
  PFN_tensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return tedm_set_error(TEDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[5], strides[4];
  if (mode == 2) {
    dims[0] = 2ull * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    strides[0] = 4ull * C; strides[1] = 2ull * W * C; strides[2] = 4ull * W * C; strides[3] = 2ull * image_stride;
  } else {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
    strides[0] = 2ull * C; strides[1] = 2ull * W * C; strides[2] = 2ull * W * C; strides[3] = 2ull * image_stride;
  }
  cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)tileW, 1u, (cuuint32_t)tileH, (cuuint32_t)tileB};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tedm_set_error(TEDM_ERR_CUDA, "cuTensorMapEncodeTiled(activation B=%d H=%d W=%d C=%d mode=%d) failed: %d", B, H,
                          W, C, mode, (int)r);
  return TEDM_OK;
}

int encode_weight_map(CUtensorMap* map, const void* ptr, long long rows, long long ktot, int bn) {
  PFN_tensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return tedm_set_error(TEDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2ull};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return tedm_set_error(TEDM_ERR_CUDA, "cuTensorMapEncodeTiled(weight rows=%lld k=%lld) failed: %d", rows, ktot, (int)r);
  return TEDM_OK;
}

template <int BN, int AM, int CPG, bool RES, int CG>
int launch_conv_cg(const ConvMaps& maps, ConvParams& p, cudaStream_t stream) {
  constexpr bool WS = AM == 1, HALO = AM >= 2;
  constexpr int NTHREADS = AM == 3 ? CONV_THREADS + XF_THREADS : CONV_THREADS;
  const int cb_total = p.c0_blocks + p.c1_blocks;
  const int b_bytes = (BN / CG) * BK * 2;            // per-CTA share of a weight tile
  const int stage_bytes = WS ? A_ROW_BYTES : (HALO ? b_bytes : A_BYTES + b_bytes);
  const int out_bytes = (BN / 64) * A_BYTES;
  // one box per 64-channel block in flight + one being consumed (AM = 3: requested two blocks ahead, one being normalised)
  const int cb_all = p.taps ? p.num_kb / p.taps : 1;
  p.a_stages = HALO ? (AM == 3 ? 4 : (p.num_kb / p.taps > 1 ? 2 : 1) + 1) : 0;
  if (HALO && p.mode == 3) p.a_stages = 2 * cb_all <= MAX_HALO_STAGES ? 2 * cb_all : cb_all;   // resident for four parity tiles
  const int base = 1024 + (WS ? 9 * cb_total * b_bytes : 0) + p.a_stages * HALO_BYTES;
  // double-buffer the output staging tile when that still leaves >= 4 pipeline slots
  p.out_bufs = (DYN_SMEM_MAX - base - 2 * out_bytes) / stage_bytes >= 4 ? 2 : 1;
  if (p.res_tma) {
    if (RES && BN <= 128 && (DYN_SMEM_MAX - base - 3 * out_bytes) / stage_bytes >= 3) p.out_bufs = 3;
    else p.res_tma = 0;
  }
  const int fixed = base + p.out_bufs * out_bytes;
  int stages = (DYN_SMEM_MAX - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  TEDM_UNSUPPORTED(stages < 2, "tedm_conv_igemm_fwd: shared memory too small for this shape");
  p.stages = stages;
  const int smem = fixed + stages * stage_bytes;
  static int configured = 0;
  if (configured < smem) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, AM, CPG, RES, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int work = p.num_tiles / CG / (HALO && p.mode == 3 ? 4 : 1);   // tiles (groups of four parity tiles) per CTA (pair)
  int grid = (work < tedm_num_sms() / CG ? work : tedm_num_sms() / CG) * CG;
  if constexpr (CG == 1) {
    conv_igemm_kernel<BN, AM, CPG, RES, 1><<<grid, NTHREADS, smem, stream>>>(maps, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TEDM_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, AM, CPG, RES, 2>, maps, p));
  }
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

template <int BN, int AM, int CPG, bool RES>
int launch_conv_res(const ConvMaps& maps, ConvParams& p, cudaStream_t stream) {
  if (p.cg == 2) return launch_conv_cg<BN, AM, CPG, RES, 2>(maps, p, stream);
  return launch_conv_cg<BN, AM, CPG, RES, 1>(maps, p, stream);
}

// residual / split-output epilogues exist only without GroupNorm statistics (they never co-occur in the net)
template <int BN, int AM, int CPG>
int launch_conv_cpg(const ConvMaps& maps, ConvParams& p, cudaStream_t stream) {
  if constexpr (CPG == 0 && AM != 3) {
    if (p.residual || p.residual2 || p.split) return launch_conv_res<BN, AM, 0, true>(maps, p, stream);
  }
  return launch_conv_res<BN, AM, CPG, false>(maps, p, stream);
}

template <int BN, int AM>
int launch_conv(const ConvMaps& maps, ConvParams& p, cudaStream_t stream) {
  const int cpg = p.gn_partial ? p.gn_cpg : 0;
  if (cpg == 0) return launch_conv_cpg<BN, AM, 0>(maps, p, stream);
  if constexpr (BN / 8 <= 8) {
    if (cpg == 8) return launch_conv_cpg<BN, AM, 8>(maps, p, stream);
  }
  if constexpr (BN / 16 <= 8) {
    if (cpg == 16) return launch_conv_cpg<BN, AM, 16>(maps, p, stream);
  }
  if (cpg == 32) return launch_conv_cpg<BN, AM, 32>(maps, p, stream);
  if constexpr (BN >= 64) {
    if (cpg == 64) return launch_conv_cpg<BN, AM, 64>(maps, p, stream);
  }
  if constexpr (BN >= 128) {
    if (cpg == 128) return launch_conv_cpg<BN, AM, 128>(maps, p, stream);
  }
  return tedm_set_error(TEDM_ERR_UNSUPPORTED, "tedm_conv_igemm_fwd: %d channels per GroupNorm group with N tile %d", cpg, BN);
}


// ==========================================================================================
// 3x3, 64 -> 64 channels on full 128-pixel rows: FOUR output rows per tile, weights resident.
//
// conv_igemm_kernel<64, WS> on these layers is bound by shared-memory bandwidth, not by the tensor core: a
// 128 x 64 x 16 UMMA reads A 4 KB + B 2 KB for 32 tensor-cycles (192 B/clk against 128 B/clk of shared memory), every
// input row is fetched three times (once per output row it feeds) and read nine times by the tensor core.  Here an input
// row is fetched ONCE per tile and multiplied by the weights of every vertical tap at once:
//     D[pixel][out row r-1 | out row r | out row r+1]  +=  in_row(r)[pixel + dx][ci] * [W(dy=2) | W(dy=1) | W(dy=0)][ci]
// i.e. one N = 192 UMMA whose 64-column thirds are the accumulators of three DIFFERENT output rows, which sit side by
// side in TMEM (a0..a3 = output rows y0..y0+3, 64 columns each).  Six input rows feed a tile:
//     row y0+1 -> a0 a1 a2 (N = 192, starts them)      row y0+4 -> a3 (N = 64, W(dy=2), starts it)
//     row y0+2 -> a1 a2 a3 (N = 192)                   row y0   -> a0 a1 (N = 128, [W1|W0])
//     row y0+3 -> a2 a3    (N = 128, [W2|W1])          row y0-1 -> a0 (N = 64, W0)
// every window of the resident [W2|W1|W0] tile is contiguous, so each is one descriptor.  Per output row the tensor core
// now reads 144 KB of operands instead of 216 KB and TMA writes 25 KB instead of 50 KB; the tensor-cycle floor
// (1152 cycles per row) and the shared-memory floor (1125) meet.
//
// warps: 0 = TMA producer, 1 = MMA issuer, 2..5 / 6..9 = two epilogue groups (one output row each at a time: TMEM ->
// registers -> bias, GroupNorm partial sums -> bf16 -> swizzled staging tile -> TMA store); two TMEM accumulator sets
// (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// CB = 2: 128 input channels (one 128-channel source or two 64-channel ones): a slot is one (row, channel block), the
//         resident weights take 144 KB, which leaves three slots and ONE staging tile, so the eight epilogue warps share a row.
// W64:    64-pixel rows (64 -> 64 at 64 x 64): M = 128 is the same row of TWO images, which makes vertical neighbours whole
//         tiles again; the horizontal taps cannot be row-shifted windows of one box here (the second image would have to
//         start 64 rows after the first, inside its halo), so a slot is one (row, dx) box of 2 x 64 pixels loaded at x = dx - 1.
struct Ws4Params {
  int B, tiles_x, rows4;         // rows4 = H / 4
  int num_tiles, stages, c0_blocks;
  const float* bias;
  float* gn_partial;
  int gn_groups, gn_parts;
  int H, W;
  const float* src_affine;       // XF: [B][64][2] (a/2, b/2): the conv's input is SiLU(a * src + b)
};

constexpr int WS4_THREADS = 64 + 256;

// XF (one channel block, 128-pixel rows): the input rows are normalised in place -- SiLU(a * x + b), the GroupNorm + scale/shift
// + SiLU of the Block that produced them -- by four extra warps between the TMA's arrival and the tensor core's reads (see
// conv_igemm_kernel AM = 3); an input row is fetched once per tile, so this costs 1.5 MUFU per element.
template <int CPG, int CB, bool W64, bool XF>
__global__ void __launch_bounds__(XF ? WS4_THREADS + XF_THREADS : WS4_THREADS, 1)
conv_ws4_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapOut, const Ws4Params p) {
  static_assert(!(W64 && CB != 1), "64-pixel rows: one channel block");
  static_assert(!XF || (CB == 1 && !W64), "input transform: one channel block, 128-pixel rows");
  constexpr int W_TILE = 192 * BK * 2;               // [W(dy=2) | W(dy=1) | W(dy=0)] of one (dx, channel block): 24 KB
  constexpr int W_BYTES = 3 * CB * W_TILE;
  constexpr int SLOT_BYTES = W64 ? A_BYTES : A_ROW_BYTES;
  constexpr int SLOT_TX = W64 ? A_BYTES : ROW_PIX * BK * 2;
  constexpr int G = CB == 1 ? 2 : 1;                 // epilogue groups (each with its own staging tile)
  constexpr int GT = 256 / G;                        // threads per group
  constexpr int WCH = G == 2 ? 2 : 1;                // 32-column chunks per warp
  constexpr int NV = 8 * WCH;                        // (sum, sumsq) of the GroupNorm groups a warp covers
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[2][4][16];     // [group, or column half when one group][lane quarter][values]
  __shared__ __align__(8) uint64_t bar_ready[XF ? MAX_STAGES : 1];   // XF: row normalised
  __shared__ __align__(16) float sbias[XF ? 64 : 4];                 // XF: the bias lives here instead of in registers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base + W_BYTES;
  const uint32_t out_base = stage_base + (uint32_t)p.stages * SLOT_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapOut);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), 8);
    }
    mbar_init(smem_u32(&bar_w), 1);
    if constexpr (XF)
      for (int s = 0; s < p.stages; ++s) mbar_init(smem_u32(&bar_ready[s]), XF_THREADS / 32);
    fence_barrier_init();
  }
  if constexpr (XF) {
    if (threadIdx.x < 64) sbias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.0f;
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // the six input rows of a tile in issue order: row offset, first accumulator, N, first row of the weight window
  constexpr int R_OFF[6] = {1, 4, 2, 0, 3, -1};
  constexpr int A_FIRST[6] = {0, 3, 1, 0, 2, 0};
  constexpr int N_COLS[6] = {192, 64, 192, 128, 128, 64};
  constexpr int W_ROW[6] = {0, 0, 0, 64, 0, 128};
  auto tile_coord = [&](int tile, int& b, int& y0, int& x0) {
    const int tx = tile % p.tiles_x, rest = tile / p.tiles_x;
    y0 = (rest % p.rows4) * 4;
    b = (rest / p.rows4) * (W64 ? 2 : 1);
    x0 = tx * BM;
  };

  if (warp == 0) {
    // ===== TMA producer =====================================================================
    if (elect_one()) {
      const uint32_t bw = smem_u32(&bar_w);
      mbar_expect_tx(bw, W_BYTES);
      for (int dx = 0; dx < 3; ++dx)
        for (int cb = 0; cb < CB; ++cb)
          for (int j = 0; j < 3; ++j)            // j-th 64-row block of the resident tile = vertical tap 2 - j
            tma_load_2d(smem_base + (uint32_t)((dx * CB + cb) * W_TILE + j * 64 * BK * 2), &mapW, bw,
                        (((2 - j) * 3 + dx) * CB + cb) * BK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int b, y0, x0;
        tile_coord(tile, b, y0, x0);
#pragma unroll
        for (int s = 0; s < 6; ++s) {
#pragma unroll
          for (int u = 0; u < (W64 ? 3 : CB); ++u) {     // u = dx (64-pixel rows) or channel block
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            mbar_expect_tx(full, SLOT_TX);
            if (W64) {
              tma_load_5d(stage_base + stage * SLOT_BYTES, &mapA0, full, 0, u - 1, 0, y0 + R_OFF[s], b);
            } else {
              const bool second = u >= p.c0_blocks;
              tma_load_5d(stage_base + stage * SLOT_BYTES, second ? &mapA1 : &mapA0, full, (second ? u - p.c0_blocks : u) * BK,
                          x0 - 1, 0, y0 + R_OFF[s], b);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =========================================================================
    if (elect_one()) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      mbar_wait(smem_u32(&bar_w), 0);
      tc_fence_after();
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int set = it & 1;
        mbar_wait(smem_u32(&bar_acc_empty[set]), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BM >> 4) << 24);
          const uint32_t idesc = IDESC_BASE | ((uint32_t)(N_COLS[s] >> 3) << 17);
          const uint32_t d_tmem = tmem_base + (uint32_t)(set * 256 + A_FIRST[s] * 64);
#pragma unroll
          for (int u = 0; u < (W64 ? 3 : CB); ++u) {
            mbar_wait(smem_u32(XF ? &bar_ready[stage] : &bar_full[stage]), phase);
            tc_fence_after();
            const uint32_t a_addr = stage_base + stage * SLOT_BYTES;
#pragma unroll
            for (int d = 0; d < (W64 ? 1 : 3); ++d) {
              const int dx = W64 ? u : d, cb = W64 ? 0 : u;
              const uint64_t adesc = make_sw128_desc(a_addr + (W64 ? 0u : (uint32_t)dx * 128u));
              const uint64_t bdesc = make_sw128_desc(smem_base + (uint32_t)((dx * CB + cb) * W_TILE + W_ROW[s] * BK * 2));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(d_tmem, adesc + 2ull * k, bdesc + 2ull * k, idesc, (s >= 2 || (u | d | k) != 0) ? 1u : 0u);
            }
            umma_commit(smem_u32(&bar_empty[stage]));
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit(smem_u32(&bar_acc_full[set]));
      }
    }
    __syncwarp();
  } else if (XF && warp >= WS4_THREADS / 32) {
    // ===== input transform (XF): 128 threads; thread = (16-byte channel chunk, pixel mod 16) ================
    if constexpr (XF) {
      const int tt = threadIdx.x - WS4_THREADS;
      const int c8 = tt & 7, prow = tt >> 3;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int b, y0, x0;
        tile_coord(tile, b, y0, x0);
        const float4* ap = reinterpret_cast<const float4*>(p.src_affine + 2 * ((size_t)b * 64 + c8 * 8));
        float ca[8], cb[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = __ldg(ap + j);
          ca[2 * j] = v.x;
          cb[2 * j] = v.y;
          ca[2 * j + 1] = v.z;
          cb[2 * j + 1] = v.w;
        }
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          const int y = y0 + R_OFF[s];
          if (y >= 0 && y < p.H) {                  // rows outside the image stay the TMA's zero fill
            // straight-line: all nine loads first (shared memory is busy feeding the tensor core), predicated stores
            constexpr int NIT = (ROW_PIX + 15) / 16;
            const uint32_t addr0 = stage_base + stage * SLOT_BYTES + (uint32_t)prow * 128u + ((uint32_t)(c8 ^ (prow & 7)) << 4);
            uint4 raw[NIT];
            bool inside[NIT];
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
              const int px = prow + 16 * i, x = x0 - 1 + px;
              inside[i] = px < ROW_PIX && x >= 0 && x < p.W;
              raw[i] = lds128(addr0 + (uint32_t)(inside[i] ? i : 0) * 2048u);
            }
#pragma unroll
            for (int i = 0; i < NIT; ++i)
              sts128_if(inside[i], addr0 + (uint32_t)(inside[i] ? i : 0) * 2048u, affine_silu8(raw[i], ca, cb));
          }
          fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's (async proxy) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_ready[stage]));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===== epilogue ===========================================================================
    // two groups (warps 2..5 / 6..9; 64 columns per thread; group g takes output rows 2g, 2g+1) or one group of eight warps
    // (32 columns per thread, rows 0..3)
    const int g = G == 2 ? (warp - 2) >> 2 : 0;
    const int hsel = G == 2 ? 0 : (warp - 2) >> 2;  // column half (one group)
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                  // M row: pixel of the 128-pixel row, or (image, pixel) of two 64-pixel rows
    const int eg = threadIdx.x - 64 - g * GT;       // index inside the group
    const uint32_t out_buf = out_base + (uint32_t)g * A_BYTES;
    const int bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    float bias_r[XF ? 1 : 32 * WCH];
    if constexpr (!XF) {
#pragma unroll
      for (int j = 0; j < 8 * WCH; ++j) {
        const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + hsel * 8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        bias_r[4 * j] = bv.x;
        bias_r[4 * j + 1] = bv.y;
        bias_r[4 * j + 2] = bv.z;
        bias_r[4 * j + 3] = bv.w;
      }
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      int b, y0, x0;
      tile_coord(tile, b, y0, x0);
      const int set = it & 1;
      float gv[NV];
#pragma unroll
      for (int half = 0; half < 4 / G; ++half) {
        const int a = (4 / G) * g + half;           // accumulator = output row y0 + a
        if (half == 0) {
          mbar_wait(smem_u32(&bar_acc_full[set]), (uint32_t)((it >> 1) & 1));
          tc_fence_after();
        }
        if (!W64 || (half & 1) == 0) {              // 64-pixel rows: a GroupNorm partial covers a PAIR of rows
#pragma unroll
          for (int i = 0; i < NV; ++i) gv[i] = 0.0f;
        }
        // TMEM -> registers -> bias, GroupNorm sums, bf16: all of it BEFORE waiting for the staging tile, so that the TMA
        // store of the previous row drains under this row's arithmetic
        uint32_t pk[16 * WCH];
#pragma unroll
        for (int ch = 0; ch < WCH; ++ch) {
          const int chunk = G == 2 ? ch : hsel;     // 32-column chunk of the 64 output channels
          uint32_t r0[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * 256 + a * 64 + chunk * 32), r0);
          tmem_ld_wait();
          if (half == 4 / G - 1 && ch == WCH - 1) {
            // every row of this warp is in registers: the accumulator set is free for the MMA issuer again
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[set]));
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + (XF ? sbias[chunk * 32 + j] : bias_r[XF ? 0 : ch * 32 + j]);
          if (CPG) {
#pragma unroll
            for (int sg = 0; sg < 4; ++sg) {
              float s_ = 0.0f, q_ = 0.0f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s_ += v[sg * 8 + j];
                q_ = fmaf(v[sg * 8 + j], v[sg * 8 + j], q_);
              }
              gv[2 * (ch * 4 + sg)] += s_;
              gv[2 * (ch * 4 + sg) + 1] += q_;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[ch * 16 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
        const bool emit = CPG != 0 && (!W64 || (half & 1) == 1);
        float t[NV];
        if (emit) {
#pragma unroll
          for (int i = 0; i < NV; ++i) t[i] = gv[i];
          butterfly_sum<NV>(t, lane, 32);           // lane l: value index l >> (NV == 16 ? 1 : 2)
        }
        // the TMA store that last read this group's staging tile must be done with it; `red` of the previous row too
        if (eg == 0) tma_store_wait_read<0>();
        named_bar_sync(bar_a, GT);
        const uint32_t sub = out_buf + (uint32_t)row * 128u;
#pragma unroll
        for (int ch = 0; ch < WCH; ++ch) {
          const int chunk = G == 2 ? ch : hsel;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t c16 = (uint32_t)(chunk * 4 + j);
            const uint32_t dst = sub + ((c16 ^ (uint32_t)(row & 7)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pk[ch * 16 + 4 * j]), "r"(pk[ch * 16 + 4 * j + 1]),
                         "r"(pk[ch * 16 + 4 * j + 2]), "r"(pk[ch * 16 + 4 * j + 3]) : "memory");
          }
        }
        if (emit) {
          constexpr int SH = NV == 16 ? 1 : 2;
          if ((lane & ((1 << SH) - 1)) == 0) red[G == 2 ? g : hsel][q][lane >> SH] = t[0];
        }
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
        named_bar_sync(bar_b, GT);
        if (eg == 0) {
          tma_store_5d(&mapOut, out_buf, 0, W64 ? 0 : x0, 0, y0 + a, b);
          tma_store_commit();
        }
        if (emit) {
          // value index i = 2 * GroupNorm group + (sum, sumsq) = the float offset inside one (image, part) record
          if (W64) {
            if (eg < 32) {
              const int img = eg >> 4, i = eg & 15;
              if (b + img < p.B)
                p.gn_partial[((size_t)(b + img) * p.gn_parts + ((y0 + a) >> 1)) * p.gn_groups * 2 + i] =
                    red[g][2 * img][i] + red[g][2 * img + 1][i];
            }
          } else if (eg < 16) {
            const int rsel = G == 2 ? g : eg >> 3, i = G == 2 ? eg : eg & 7;
            const float tot = red[rsel][0][i] + red[rsel][1][i] + red[rsel][2][i] + red[rsel][3][i];
            p.gn_partial[((size_t)b * p.gn_parts + (y0 + a) * p.tiles_x + x0 / BM) * p.gn_groups * 2 + eg] = tot;
          }
        }
      }
    }
    if (eg == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int CPG, int CB, bool W64, bool XF>
int launch_ws4(const CUtensorMap& mapA0, const CUtensorMap& mapA1, const CUtensorMap& mapW, const CUtensorMap& mapOut,
               Ws4Params& p, cudaStream_t stream) {
  const int slot = W64 ? A_BYTES : A_ROW_BYTES;
  const int fixed = 1024 + 3 * CB * 192 * BK * 2 + (CB == 1 ? 2 : 1) * A_BYTES;
  int stages = (DYN_SMEM_MAX - fixed) / slot;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  TEDM_UNSUPPORTED(stages < 3, "tedm_conv_igemm_fwd: shared memory too small for the 4-row weight-stationary kernel");
  p.stages = stages;
  const int smem = fixed + stages * slot;
  static int configured = 0;
  if (configured < smem) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_ws4_kernel<CPG, CB, W64, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int grid = p.num_tiles < tedm_num_sms() ? p.num_tiles : tedm_num_sms();
  conv_ws4_kernel<CPG, CB, W64, XF><<<grid, XF ? WS4_THREADS + XF_THREADS : WS4_THREADS, smem, stream>>>(mapA0, mapA1, mapW, mapOut, p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

template <int CB, bool W64>
int launch_ws4_cpg(const CUtensorMap& mapA0, const CUtensorMap& mapA1, const CUtensorMap& mapW, const CUtensorMap& mapOut,
                   Ws4Params& p, cudaStream_t stream) {
  if constexpr (CB == 1 && !W64) {
    if (p.src_affine)
      return p.gn_partial ? launch_ws4<8, CB, W64, true>(mapA0, mapA1, mapW, mapOut, p, stream)
                          : launch_ws4<0, CB, W64, true>(mapA0, mapA1, mapW, mapOut, p, stream);
  }
  return p.gn_partial ? launch_ws4<8, CB, W64, false>(mapA0, mapA1, mapW, mapOut, p, stream)
                      : launch_ws4<0, CB, W64, false>(mapA0, mapA1, mapW, mapOut, p, stream);
}

// ==========================================================================================
// weight gradient:  dW[co][tap][ci] = sum_pixels dY[pixel][co] * X[pixel + tap offset][ci]
//
// The contraction runs over pixels, which are the ROWS of both NHWC operands, so both are fed to the
// tensor core as MN-major operands: a TMA box [128 pixels][64 channels] (128B swizzle) is exactly the
// canonical MN-major SW128 layout (8-row K groups 1024 B apart = SBO, 64-channel MN blocks 16 KB
// apart = LBO).  GEMM view:  D[m][n],  m = (tap, ci) in blocks of 2 x 64 (two boxes with their own
// tap shift), n = co tile, k = 128 pixels per pipeline slot (8 UMMAs of K = 16).  Split-K over pixel
// tiles across CTAs; partial tiles are reduced with fp32 red.global.add into a zeroed fp32
// [cout][taps][cin] buffer.
// ==========================================================================================
struct WgradParams {
  int mode, taps, kxc;           // taps counts parities x 2x2 in mode 3 (16)
  int c0_blocks, c1_blocks, C0, C1, ctot;
  int tileW, tileH, tileB, tiles_x, tiles_y;
  int B, cout;
  int items, n_tiles, splitk, num_ptiles, stages;
  int oihw;                      // 1: accumulate into an fp32 OIHW gradient (mode 3: un-folded to 3x3)
  float* dw;
  int* turn;                     // deterministic mode: one counter per (m block, n tile) group -- whose split-K slice adds next
                                 // (self-resetting); NULL: slices add in arrival order
};

__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr) {
  // MN-major, 128B swizzle: LBO = 16 KB between 64-element MN blocks, SBO = 1 KB between 8-row K groups
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1024ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap mapX0, const __grid_constant__ CUtensorMap mapX1,
                  const __grid_constant__ CUtensorMap mapDY, const WgradParams p) {
  constexpr int STAGE_BYTES = 2 * A_BYTES + (BN / 64) * A_BYTES;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int cb_total = p.c0_blocks + p.c1_blocks;
  const int ks = blockIdx.x;
  const int grp = blockIdx.y;
  const int mb = grp / p.n_tiles, nt = grp % p.n_tiles;
  const int n0 = nt * BN;
  const int item0 = 2 * mb, item1 = (2 * mb + 1 < p.items) ? 2 * mb + 1 : 2 * mb;   // odd tail: duplicate, discarded by the reduce
  const int num_kb = (p.num_ptiles - ks + p.splitk - 1) / p.splitk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX0);
    if (p.c1_blocks) tma_prefetch_desc(&mapX1);
    tma_prefetch_desc(&mapDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BN>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // per-CTA constants of the two (tap, channel block) items, hoisted out of the pixel-tile loop
      int i_offy[2], i_offx[2], i_pc[2], i_chan[2], par = 0;
      const CUtensorMap* i_map[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int item = h ? item1 : item0;
        const int tap = item / cb_total, cb = item % cb_total;
        const bool second = cb >= p.c0_blocks;
        const int cblk = second ? cb - p.c0_blocks : cb;
        int offy = 0, offx = 0, pc = 0, chan_off = 0;
        if (p.mode == 1) {
          offy = tap / 3 - 1;
          offx = tap % 3 - 1;
        } else if (p.mode == 2) {
          const int ky = tap >> 2, kx = tap & 3;
          offy = ((ky + 1) >> 1) - 1;
          offx = ((kx + 1) >> 1) - 1;
          pc = (ky + 1) & 1;
          chan_off = ((kx + 1) & 1) * (second ? p.C1 : p.C0);
        } else if (p.mode == 3) {
          par = tap >> 2;
          offy = ((tap >> 1) & 1) - 1 + (par >> 1);
          offx = (tap & 1) - 1 + (par & 1);
        }
        i_offy[h] = offy; i_offx[h] = offx; i_pc[h] = pc; i_chan[h] = chan_off + cblk * BK;
        i_map[h] = second ? &mapX1 : &mapX0;
      }
      const int dy_chan = (p.mode == 3 ? (par & 1) * p.cout : 0) + n0;
      const int dy_pc = p.mode == 3 ? (par >> 1) : 0;
      const int tiles_per_group = p.tiles_x * p.tiles_y;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int pt = ks + kb * p.splitk;
        const int bg = pt / tiles_per_group, trem = pt - bg * tiles_per_group;
        const int ty = trem / p.tiles_x;
        const int b0 = bg * p.tileB, y0 = ty * p.tileH, x0 = (trem - ty * p.tiles_x) * p.tileW;
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        const uint32_t full = smem_u32(&bar_full[stage]);
        mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_5d(a_dst + h * A_BYTES, i_map[h], full, i_chan[h], x0 + i_offx[h], i_pc[h], y0 + i_offy[h], b0);
        for (int j = 0; j < BN / 64; ++j)
          tma_load_5d(a_dst + (2 + j) * A_BYTES, &mapDY, full, dy_chan + j * 64, x0, dy_pc, y0, b0);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
        const uint64_t adesc = make_sw128_mn_desc(a_addr), bdesc = make_sw128_mn_desc(a_addr + 2 * A_BYTES);
#pragma unroll
        for (int k = 0; k < BM / 16; ++k)   // 16 pixel rows (2048 B) per UMMA
          umma_bf16(tmem_base, adesc + 128ull * k, bdesc + 128ull * k, IDESC, (kb | k) != 0 ? 1u : 0u);
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(smem_u32(&bar_acc));
    }
    __syncwarp();
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    const int half = row >> 6;
    const int item = half ? item1 : item0;
    const bool live = num_kb > 0 && (half == 0 || item1 != item0);
    const int tap = item / cb_total, cb = item % cb_total;
    const int ci = cb * BK + (row & 63);
    // destination of output channel co:  base + co * co_stride
    size_t base, co_stride;
    if (!p.oihw) {
      base = (size_t)tap * p.ctot + ci;
      co_stride = (size_t)p.taps * p.ctot;
    } else {
      base = (size_t)ci * p.taps + tap;
      co_stride = (size_t)p.ctot * p.taps;
    }
    if (num_kb > 0) {
      mbar_wait(smem_u32(&bar_acc), 0);
      tc_fence_after();
    }
    // Split-K slices of one (m block, n tile) group add into dW with fire-and-forget fp32 reductions.  Deterministic mode
    // (tedm_conv_set_deterministic) makes them add IN SLICE ORDER: slice ks waits for turn[group] == ks, adds, fences,
    // passes the turn on (the last slice resets it to 0).  Every address then receives its addends in a fixed order and
    // the gradient is bit-reproducible.  Slices of a group have consecutive linear block ids (blockIdx.x = ks), so a
    // waiting CTA's predecessors are resident or finished.
    const bool ordered = p.turn != nullptr && p.splitk > 1;
    int* turn = p.turn + grp;
    if (ordered) {
      if (threadIdx.x == 64) {
        int v;
        const long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(turn) : "memory");
          if (v != ks && clock64() - t0 > 4000000000LL) {      // bounded like mbar_wait: a protocol bug must not hang the box
            printf("tedm_b200: wgrad turn wait timed out (group %d slice %d sees %d)\n", grp, ks, v);
            __trap();
          }
        } while (v != ks);
      }
      named_bar_sync(1, 128);
    }
    if (num_kb > 0) {
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chunk * 32), r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(p.dw + base + (size_t)(n0 + chunk * 32 + j) * co_stride, __uint_as_float(r[j]));
        }
      }
    }
    if (ordered) {
      __threadfence();
      named_bar_sync(1, 128);
      if (threadIdx.x == 64) {
        const int next = ks == p.splitk - 1 ? 0 : ks + 1;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(turn), "r"(next) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

template <int BN>
int launch_wgrad(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& dy, WgradParams& p, int m_blocks,
                 cudaStream_t stream) {
  const int stage_bytes = (2 + BN / 64) * A_BYTES;
  int stages = (DYN_SMEM_MAX - 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  const int smem = 1024 + stages * stage_bytes;
  static int configured = 0;
  if (configured < smem) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid((unsigned)p.splitk, (unsigned)(m_blocks * p.n_tiles));
  conv_wgrad_kernel<BN><<<grid, 192, smem, stream>>>(x0, x1, dy, p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

constexpr long long WGRAD_PART_FLOATS_PER_SM = 3LL * 5 * BM * 64;   // halo-tile kernel: 3 waves of CTAs x 5 x 128 x 64 fp32 partials
constexpr long long WGRAD_RAW_FLOATS = 4LL << 20;                   // raw [co][16][ci] intermediate of the folded upsample conv
constexpr long long WGRAD_TURN_INTS = 1LL << 16;                    // split-K turn counters of the generic kernel

// ==========================================================================================
// weight gradient of the 3x3 convolution, halo-tile form
//
// The generic kernel above fetches one 128-pixel activation box per (tap, 64-channel block) and is bound by L2 -> SM
// traffic on the thin layers (48 KB per 1 MFLOP-pair).  Here one CTA owns a (64 input channels) x (64 output channels)
// block of dW for ALL nine taps: per pixel tile it loads the activation tile ONCE with a one-pixel halo
// ((tileH+2) x (tileW+2) pixels, out-of-bounds = the conv's zero padding) plus the 128-pixel dy tile, and the nine taps
// are nine row-shifted windows of the same shared-memory box (MN-major UMMA operands: a window is just a start
// address; two taps are stacked into one M = 128 operand through the descriptor's leading-dimension byte offset).
// Accumulators: 5 blocks of [128 = 2 taps x 64 ci][64 co] fp32 in TMEM (taps (0,1) (2,3) (4,5) (6,7) (7,8); the second
// copy of tap 7 is discarded).  Split-K partial tiles go to a workspace with plain stores and a second kernel sums
// them in slice order and lays dW out (OIHW accumulate, or [co][tap][ci]): deterministic, no floating-point atomics.
// ==========================================================================================
struct Wgrad3Params {
  int c0_blocks, c1_blocks, ctot;
  int tileW, tileH, tileB, tiles_x, tiles_y;
  int B, cout, n_tiles, splitk, num_ptiles, stages;
  int x_bytes;       // shared-memory bytes reserved for one halo box (multiple of 1024)
  int x_tx_bytes;    // bytes the TMA writes for one halo box
  int oihw;
  float* dw;
  float* part;       // split-K partial tiles [group][ks][5][128][64] fp32
};

__device__ __forceinline__ uint64_t make_sw128_mn_desc2(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(192, 1)
conv_wgrad3_kernel(const __grid_constant__ CUtensorMap mapX0, const __grid_constant__ CUtensorMap mapX1,
                   const __grid_constant__ CUtensorMap mapDY, const Wgrad3Params p) {
  constexpr int BN = 64;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = p.x_bytes + A_BYTES;
  const int ks = blockIdx.x;
  const int cb = blockIdx.y / p.n_tiles, nt = blockIdx.y % p.n_tiles;
  const int n0 = nt * BN;
  const int num_kb = (p.num_ptiles - ks + p.splitk - 1) / p.splitk;
  const int hw_pitch = p.tileW + 2;   // halo box row pitch in pixels

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX0);
    if (p.c1_blocks) tma_prefetch_desc(&mapX1);
    tma_prefetch_desc(&mapDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const bool second = cb >= p.c0_blocks;
      const int cblk = second ? cb - p.c0_blocks : cb;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int pt = ks + kb * p.splitk;
        const int tiles_per_group = p.tiles_x * p.tiles_y;
        const int bg = pt / tiles_per_group, trem = pt % tiles_per_group;
        const int b0 = bg * p.tileB, y0 = (trem / p.tiles_x) * p.tileH, x0 = (trem % p.tiles_x) * p.tileW;
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        const uint32_t full = smem_u32(&bar_full[stage]);
        mbar_expect_tx(full, (uint32_t)(p.x_tx_bytes + A_BYTES));
        const uint32_t dst = smem_base + stage * stage_bytes;
        tma_load_5d(dst, second ? &mapX1 : &mapX0, full, cblk * BK, x0 - 1, 0, y0 - 1, b0);
        tma_load_5d(dst + p.x_bytes, &mapDY, full, n0, x0, 0, y0, b0);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // one UMMA covers 16 consecutive tile pixels = two 8-pixel K groups: contiguous inside an image row when the
      // tile is at least 16 pixels wide, one image row apart for 8-pixel-wide tiles
      const uint32_t sbo = p.tileW >= 16 ? 1024u : (uint32_t)hw_pitch * 128u;
      // everything address-like is hoisted out of the issue loop (one thread issues 40 UMMAs per pixel tile): byte offset
      // of each 16-pixel K step inside the halo box, and the descriptor template (start offset, leading-dimension byte
      // offset between the two stacked taps) of each tap pair
      uint32_t koff[BM / 16];
#pragma unroll
      for (int k = 0; k < BM / 16; ++k) {
        const int m = 16 * k;
        const int tb = m / (p.tileW * p.tileH), rem = m % (p.tileW * p.tileH);
        koff[k] = (uint32_t)((tb * (p.tileH + 2) + rem / p.tileW) * hw_pitch + rem % p.tileW) * 128u;
      }
      uint64_t ablk[5];
#pragma unroll
      for (int blk = 0; blk < 5; ++blk) {
        const int t0 = blk < 4 ? 2 * blk : 7, t1 = t0 + 1;
        const int off0 = (t0 / 3) * hw_pitch + t0 % 3, off1 = (t1 / 3) * hw_pitch + t1 % 3;
        ablk[blk] = make_sw128_mn_desc2((uint32_t)off0 * 128u, (uint32_t)(off1 - off0) * 128u, sbo);   // + tile address per use
      }
      const uint64_t btmpl = make_sw128_mn_desc2(0u, 16384u, 1024u);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t xbase = smem_base + stage * stage_bytes, dybase = xbase + p.x_bytes;
#pragma unroll
        for (int blk = 0; blk < 5; ++blk) {
#pragma unroll
          for (int k = 0; k < BM / 16; ++k) {
            const uint64_t adesc = ablk[blk] + (uint64_t)(((xbase + koff[k]) & 0x3FFFFu) >> 4);
            const uint64_t bdesc = btmpl + (uint64_t)(((dybase + (uint32_t)k * 2048u) & 0x3FFFFu) >> 4);
            umma_bf16(tmem_base + (uint32_t)(blk * BN), adesc, bdesc, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(smem_u32(&bar_acc));
    }
    __syncwarp();
  } else {
    const int q = warp & 3, row = q * 32 + lane;
    if (num_kb > 0) {
      mbar_wait(smem_u32(&bar_acc), 0);
      tc_fence_after();
    }
    float* ptile = p.part + ((size_t)blockIdx.y * p.splitk + ks) * (5 * BM * BN) + (size_t)row * BN;
#pragma unroll 1
    for (int blk = 0; blk < 5; ++blk) {
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t r[32];
        if (num_kb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * BN + chunk * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        float4* dst = reinterpret_cast<float4*>(ptile + (size_t)blk * (BM * BN) + chunk * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// Sums the split-K partial tiles of conv_wgrad3_kernel and writes dW.  CTA = (input channel, 64 x 64 block, tap):
// 64 output channels x 4 split-K slices; every thread keeps 8 independent loads in flight (the sum over up to 148
// partials is latency-bound otherwise), partial reads are coalesced along the output channel.
//   oihw: dw[co][ci][3][3] += sum      else: dw[co][tap][ci] = sum
template <int SLICES>   // 4: CTA = 1 input channel x 64 co x 4 split-K slices (deep splits); 1: CTA = 4 input channels x 64 co
__global__ void __launch_bounds__(256) wgrad3_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int splitk,
                                                            int n_tiles, int ctot, int oihw) {
  __shared__ float red[4][64];
  const int g = blockIdx.y, cb = g / n_tiles, n0 = (g % n_tiles) * 64;
  const int t = blockIdx.z;                                    // tap
  const int co_l = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int ci_l = SLICES == 4 ? blockIdx.x : blockIdx.x * 4 + sub;
  const int slice = SLICES == 4 ? sub : 0;
  const int blk = t < 8 ? t >> 1 : 4, half = t < 8 ? t & 1 : 1;
  constexpr size_t TILE = 5 * BM * 64;
  const float* src = part + (size_t)g * splitk * TILE + ((size_t)blk * BM + half * 64 + ci_l) * 64 + co_l;
  float a[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) a[u] = 0.0f;
  int ks = slice;
  for (; ks + 7 * SLICES < splitk; ks += 8 * SLICES) {
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] += __ldg(src + (size_t)(ks + SLICES * u) * TILE);
  }
  for (; ks < splitk; ks += SLICES) a[0] += __ldg(src + (size_t)ks * TILE);
  float acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  if (SLICES == 4) {
    red[slice][co_l] = acc;
    __syncthreads();
    if (slice != 0) return;
    acc = (red[0][co_l] + red[1][co_l]) + (red[2][co_l] + red[3][co_l]);
  }
  const int co = n0 + co_l, ci = cb * 64 + ci_l;
  if (oihw) dw[((size_t)co * ctot + ci) * 9 + t] += acc;
  else dw[((size_t)co * 9 + t) * ctot + ci] = acc;
}

// Shallow splits (many 64 x 64 blocks, few partials each), OIHW accumulate: CTA = 4 output channels x 64 input channels x
// all nine taps, so that every thread adds 9 consecutive floats and a warp covers 288-byte runs of the OIHW gradient
// (the per-tap variant above scatters 4-byte read-modify-writes 9 * ctot floats apart).
__global__ void __launch_bounds__(256) wgrad3_reduce_oihw_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                                                 int splitk, int n_tiles, int ctot) {
  const int g = blockIdx.y, cb = g / n_tiles, n0 = (g % n_tiles) * 64;
  const int co_l = blockIdx.x * 4 + (threadIdx.x & 3), ci_l = threadIdx.x >> 2;
  constexpr size_t TILE = 5 * BM * 64;
  const float* src = part + (size_t)g * splitk * TILE + (size_t)ci_l * 64 + co_l;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.0f;
  for (int ks = 0; ks < splitk; ++ks) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int blk = t < 8 ? t >> 1 : 4, half = t < 8 ? t & 1 : 1;
      acc[t] += __ldg(src + (size_t)ks * TILE + ((size_t)blk * BM + half * 64) * 64);
    }
  }
  float* d = dw + ((size_t)(n0 + co_l) * ctot + cb * 64 + ci_l) * 9;
#pragma unroll
  for (int t = 0; t < 9; ++t) d[t] += acc[t];
}

int g_enable_wgrad3 = 1;  // tedm_conv_set_wgrad_halo: 0 off, 1 automatic, 2 wherever the geometry allows
int g_deterministic = 0;  // tedm_conv_set_deterministic

int g_enable_ws = 1;  // tedm_conv_set_ws
int g_enable_halo = 1;  // tedm_conv_set_halo
int g_enable_res_tma = 1;   // tedm_conv_set_halo(2) turns it off (A/B)
int g_enable_halo3 = 0;     // tedm_conv_set_halo(3): halo tiles for the folded upsample conv too (measured slower)
int g_enable_pairs = 1;  // tedm_conv_set_cta_pairs
int g_force_bn = 0;  // debug/tuning override (tedm_conv_set_tile_n)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" int tedm_conv_set_tile_n(int bn) {
  TEDM_CHECK_ARG(bn == 0 || bn == 64 || bn == 128 || bn == 256, "tedm_conv_set_tile_n: bn=%d", bn);
  g_force_bn = bn;
  return TEDM_OK;
}

extern "C" int tedm_conv_set_wgrad_halo(int enable) {
  g_enable_wgrad3 = enable;
  return TEDM_OK;
}

extern "C" int tedm_conv_set_deterministic(int enable) {
  g_deterministic = enable != 0;
  return TEDM_OK;
}

extern "C" int tedm_conv_set_cta_pairs(int enable) {
  g_enable_pairs = enable;                // 0 off, 1 automatic, 2 wherever the geometry allows
  return TEDM_OK;
}

extern "C" int tedm_conv_set_ws(int enable) {
  g_enable_ws = enable;                   // 0 off, 1 on (four-row tiles where they apply), 2 single-row tiles only
  return TEDM_OK;
}

extern "C" int tedm_conv_set_halo(int enable) {
  g_enable_res_tma = enable != 2;         // 2 = halo tiles on, residual tiles of the 1x1 convs through registers (A/B)
  g_enable_halo3 = enable == 3;           // 3 = halo tiles for the folded upsample conv as well (A/B, tests)
  g_enable_halo = enable != 0;
  return TEDM_OK;
}

extern "C" int tedm_conv_src_affine_supported(int height, int width, int c0, int cout) {
  // mirrors the halo-mode selection of tedm_conv_igemm_fwd for a single-source 3x3 conv with GroupNorm groups of cout / 8
  if (!is_pow2(height) || !is_pow2(width) || c0 % BK != 0 || cout % 64 != 0) return 0;
  const bool ws_geom = g_enable_ws && width >= BM && cout == 64 && c0 <= 128;
  const bool w64_geom = g_enable_ws == 1 && width == 64 && height % 4 == 0 && c0 == 64 && cout == 64;
  if (ws_geom) return (g_enable_ws == 1 && c0 == 64 && height % 4 == 0) ? 1 : 0;     // conv_ws4_kernel<.., XF>
  return (g_enable_halo && width % HALO_W == 0 && height % HALO_H == 0 && !w64_geom) ? 1 : 0;
}

extern "C" int tedm_conv_gn_parts(int out_height, int out_width) {
  const long long hw = (long long)out_height * out_width;
  return hw >= BM ? (int)(hw / BM) : 1;
}

extern "C" int tedm_conv_igemm_fwd(const tedm_conv_args* a, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && a->src0 && a->weight && a->out, "tedm_conv_igemm_fwd: null pointer");
  TEDM_CHECK_ARG(a->mode >= 0 && a->mode <= 3, "tedm_conv_igemm_fwd: mode=%d", a->mode);
  TEDM_CHECK_ARG(a->batch > 0 && a->height > 0 && a->width > 0 && a->c0 > 0 && a->c1 >= 0 && a->cout > 0,
                 "tedm_conv_igemm_fwd: bad sizes");
  TEDM_CHECK_ARG((a->c1 > 0) == (a->src1 != nullptr), "tedm_conv_igemm_fwd: src1/c1 mismatch");
  TEDM_UNSUPPORTED(a->c0 % BK != 0 || a->c1 % BK != 0, "tedm_conv_igemm_fwd: channel counts (%d, %d) must be multiples of %d",
                   a->c0, a->c1, BK);
  TEDM_UNSUPPORTED(a->cout % 64 != 0, "tedm_conv_igemm_fwd: cout=%d must be a multiple of 64", a->cout);
  TEDM_CHECK_ARG(a->n_extra >= 0 && a->n_extra <= MAX_SRC - 2, "tedm_conv_igemm_fwd: n_extra=%d", a->n_extra);

  ConvParams p{};
  p.mode = a->mode;
  p.kxc = a->mode == 0 ? 1 : a->mode == 1 ? 3 : a->mode == 2 ? 4 : 2;
  p.taps = p.kxc * p.kxc;
  p.c0_blocks = a->c0 / BK;
  p.c1_blocks = a->c1 / BK;
  // the A sources in K order: src0, src1, then the extras
  const void* src_ptr[MAX_SRC];
  long long src_stride[MAX_SRC];
  long long ktot = 0;
  auto add_src = [&](const void* ptr, int c, long long stride, int centre) {
    src_ptr[p.n_src] = ptr;
    src_stride[p.n_src] = stride ? stride : (long long)a->height * a->width * c;
    p.src_blocks[p.n_src] = c / BK;
    p.src_C[p.n_src] = c;
    p.src_center[p.n_src] = centre;
    p.num_kb += (centre ? 1 : p.taps) * (c / BK);
    ktot += (long long)(centre ? 1 : p.taps) * c;
    ++p.n_src;
  };
  add_src(a->src0, a->c0, a->src0_image_stride, 0);
  if (a->src1) add_src(a->src1, a->c1, a->src1_image_stride, 0);
  for (int i = 0; i < a->n_extra; ++i) {
    TEDM_CHECK_ARG(a->extra_src[i] && a->extra_c[i] > 0, "tedm_conv_igemm_fwd: extra source %d is empty", i);
    TEDM_UNSUPPORTED(a->extra_c[i] % BK != 0, "tedm_conv_igemm_fwd: extra source %d has %d channels (multiple of %d needed)", i,
                     a->extra_c[i], BK);
    TEDM_UNSUPPORTED(a->extra_center[i] && a->mode != 1, "tedm_conv_igemm_fwd: centre-tap sources belong to a 3x3 conv");
    add_src(a->extra_src[i], a->extra_c[i], a->extra_image_stride[i], a->extra_center[i] != 0);
  }
  p.B = a->batch;
  p.Ho = a->mode == 2 ? a->height / 2 : a->height;
  p.Wo = a->mode == 2 ? a->width / 2 : a->width;
  p.osy = p.osx = a->mode == 3 ? 2 : 1;
  p.OH = p.Ho * p.osy;
  p.OW = p.Wo * p.osx;
  TEDM_UNSUPPORTED(!is_pow2(p.Ho) || !is_pow2(p.Wo) || (a->mode == 2 && ((a->height | a->width) & 1)),
                   "tedm_conv_igemm_fwd: spatial extent %dx%d must be powers of two", a->height, a->width);
  TEDM_UNSUPPORTED((long long)p.Ho * p.Wo < 16, "tedm_conv_igemm_fwd: output extent %dx%d has fewer than 16 pixels", p.Ho, p.Wo);
  p.tileW = p.Wo < BM ? p.Wo : BM;
  p.tileH = p.Ho < BM / p.tileW ? p.Ho : BM / p.tileW;
  p.tileB = BM / (p.tileW * p.tileH);
  p.tiles_x = p.Wo / p.tileW;
  p.tiles_y = p.Ho / p.tileH;
  // halo mode (3x3 whose nine taps read one 10 x 18-pixel box per 64-channel block): tiles of 8 pixels x 16 rows.  The layers
  // the weight-stationary kernels take (whole 128-pixel rows into 64 channels; 64 -> 64 on 64-pixel rows) keep those.
  const bool ws_geom = g_enable_ws && a->mode == 1 && p.Wo >= BM && a->cout == 64 && (a->c0 + a->c1) <= 128 && a->n_extra == 0;
  const bool w64_geom = g_enable_ws == 1 && a->mode == 1 && p.Wo == 64 && p.Ho % 4 == 0 && a->c0 == 64 && a->c1 == 0 &&
                        a->n_extra == 0 && a->cout == 64 && a->out_dtype == 0 && !a->residual && !a->split &&
                        (!a->gn_partial || a->cout / a->gn_groups == 8);
  // ... and, only with tedm_conv_set_halo(3), the folded upsample conv (mode 3) with one N tile and at most four channel blocks
  // (their boxes stay resident for the four parity tiles of an M tile).  Measured on B200: 128 -> 64 @64x64 0.322 ms against
  // 0.300 ms with one box per (parity, tap), 256 -> 128 @32x32 0.178 against 0.135: these tiles have K = 4 taps x 2-4 blocks
  // and are bound by their epilogues, not by operand traffic, so the default keeps the per-tap tiles.
  const bool halo3 = g_enable_halo3 && a->mode == 3 && a->n_extra == 0 && a->c1 == 0 && p.Wo % HALO_W == 0 && p.Ho % HALO_H == 0 &&
                     a->c0 <= 4 * BK && (a->cout == 64 || a->cout == 128 || a->cout == 256) && g_force_bn == 0;
  const bool halo = (g_enable_halo && a->mode == 1 && a->n_extra == 0 && p.Wo % HALO_W == 0 && p.Ho % HALO_H == 0 && !ws_geom &&
                     !w64_geom) || halo3;
  if (a->src0_affine) {
    // the halo-tile path, or the four-row weight-stationary kernel with one channel block on 128-pixel rows
    const bool ws4_xf = ws_geom && g_enable_ws == 1 && a->c0 == 64 && p.Ho % 4 == 0 && (!a->gn_partial || a->cout / a->gn_groups == 8);
    TEDM_UNSUPPORTED((!halo && !ws4_xf) || a->src1 || a->residual || a->split || a->out_dtype != 0,
                     "tedm_conv_igemm_fwd: src0_affine needs a single-source 3x3 conv on the halo-tile or four-row path, plain "
                     "bf16 output (see tedm_conv_src_affine_supported)");
    p.src_affine = a->src0_affine;
  }
  if (halo) {
    p.tileW = HALO_W;
    p.tileH = HALO_H;
    p.tileB = 1;
    p.tiles_x = p.Wo / HALO_W;
    p.tiles_y = p.Ho / HALO_H;
  }
  p.cout = a->cout;
  p.bias = a->bias;
  p.residual = (const bf16*)a->residual;
  p.out = (bf16*)a->out;
  p.out_f32 = a->out_dtype == 1;
  TEDM_CHECK_ARG(a->out_dtype == 0 || a->out_dtype == 1, "tedm_conv_igemm_fwd: out_dtype=%d", a->out_dtype);
  p.split = a->split;
  if (a->split) {
    TEDM_CHECK_ARG(a->out2 && a->split > 0 && a->split < a->cout, "tedm_conv_igemm_fwd: split=%d needs out2 and 0 < split < cout", a->split);
    TEDM_UNSUPPORTED(a->split % 64 != 0 || p.out_f32 || a->gn_partial || a->mode > 1 || a->out_image_stride != 0,
                     "tedm_conv_igemm_fwd: split output needs split %% 64 == 0, bf16 dense outputs, no GroupNorm, mode 0/1");
    p.out2 = (bf16*)a->out2;
    p.residual2 = (const bf16*)a->residual2;
    p.out2_image_stride = (long long)p.OH * p.OW * (a->cout - a->split);
  }
  p.out_image_stride = a->out_image_stride ? a->out_image_stride : (long long)p.OH * p.OW * (a->split ? a->split : a->cout);
  if (a->residual_affine) {
    TEDM_CHECK_ARG(a->residual != nullptr, "tedm_conv_igemm_fwd: residual_affine without a residual");
    TEDM_UNSUPPORTED(a->split != 0 || p.tileB != 1 || a->cout > 256 * 1024,
                     "tedm_conv_igemm_fwd: residual_affine needs a single output and tiles inside one image (Ho * Wo >= 128)");
    p.res_affine = reinterpret_cast<const float2*>(a->residual_affine);
  }
  p.gn_partial = a->gn_partial;
  if (a->gn_partial) {
    TEDM_CHECK_ARG(a->gn_groups > 0 && a->cout % a->gn_groups == 0, "tedm_conv_igemm_fwd: gn_groups=%d", a->gn_groups);
    TEDM_UNSUPPORTED(a->mode == 3, "tedm_conv_igemm_fwd: GroupNorm partials are not produced in upsample mode");
    TEDM_UNSUPPORTED(a->residual != nullptr || a->split != 0,
                     "tedm_conv_igemm_fwd: GroupNorm partials together with a residual / split output are not supported");
    p.gn_groups = a->gn_groups;
    p.gn_cpg = a->cout / a->gn_groups;
    p.gn_parts = tedm_conv_gn_parts(p.Ho, p.Wo);
    TEDM_UNSUPPORTED(p.gn_cpg != 8 && p.gn_cpg != 16 && p.gn_cpg != 32 && p.gn_cpg != 64 && p.gn_cpg != 128,
                     "tedm_conv_igemm_fwd: %d channels per GroupNorm group unsupported (8/16/32/64/128)", p.gn_cpg);
  }

  // ---- N tile: the widest that divides cout, covers whole GroupNorm groups and still gives every SM a tile
  const long long m_tiles = (long long)ceil_div(p.B, p.tileB) * p.tiles_x * p.tiles_y;
  const int zdim = a->mode == 3 ? 4 : 1;
  auto legal = [&](int c) {
    if (a->residual_affine && c == 256) return false;    // the one-tile-ahead residual pipeline exists for N <= 128
    return a->cout % c == 0 && (!p.gn_partial || (c % p.gn_cpg == 0 && c / p.gn_cpg <= 8));
  };
  int bn = 0;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3 && bn == 0; ++i)
    if (legal(cands[i]) && m_tiles * zdim * (a->cout / cands[i]) >= tedm_num_sms()) bn = cands[i];
  for (int i = 2; i >= 0 && bn == 0; --i)  // nothing fills the machine: take the narrowest legal tile
    if (legal(cands[i])) bn = cands[i];
  if (g_force_bn && legal(g_force_bn)) bn = g_force_bn;
  if (halo3) bn = a->cout;                   // one N tile: the four parities of an M tile are consecutive tile indices
  TEDM_UNSUPPORTED(bn == 0, "tedm_conv_igemm_fwd: no N tile for cout=%d with %d-channel GroupNorm groups", a->cout, p.gn_cpg);
  const long long num_tiles = m_tiles * zdim * (a->cout / bn);
  TEDM_CHECK_ARG(num_tiles <= 2147483647LL, "tedm_conv_igemm_fwd: too many tiles");
  p.n_tiles = a->cout / bn;
  p.zdim = zdim;
  p.num_tiles = (int)num_tiles;
  auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  p.z_shift = ilog2(zdim);
  p.tx_shift = ilog2(p.tiles_x);
  p.tpg_shift = p.tx_shift + ilog2(p.tiles_y);

  // weight-stationary row mode: 3x3, whole 128-pixel rows per tile, one N tile of 64, weights fit in smem
  const bool ws = g_enable_ws && a->mode == 1 && p.tileH == 1 && p.tileB == 1 && p.tileW == BM && a->cout == 64 && bn == 64 &&
                  (a->c0 + a->c1) <= 128 && a->n_extra == 0;

  // CTA pairs (cta_group::2): a real K loop (3x3 / 4x4 / folded upsample over >= 128 input channels).  Shared memory serves
  // the tensor core's operand reads AND the TMA's writes; a pair halves the weight half of both.  Measured on B200
  // (profiles/r02_conv_pairs_ab.txt): +4-8 % on N <= 128 tiles, +8-10 % on N = 256 tiles; the HBM-bound 1x1 convolutions
  // lose 20-70 % in lock-step pairs and single-channel-block layers 30 %, so those stay one CTA per tile.
  // four output rows per tile (conv_ws4_kernel): 64 or 128 input channels on 128-pixel rows, or 64 on 64-pixel rows (two
  // images per tile); plain bf16 output, GroupNorm groups of 8 or none
  const bool ws4_epilogue = g_enable_ws == 1 && p.Ho % 4 == 0 && !p.out_f32 && !p.residual && !p.split && (!p.gn_partial || p.gn_cpg == 8);
  const bool w64 = ws4_epilogue && a->mode == 1 && p.Wo == 64 && a->c0 == 64 && a->c1 == 0 && a->n_extra == 0 && a->cout == 64;
  const bool ws4 = (ws && ws4_epilogue) || w64;
  p.cg = (g_enable_pairs && m_tiles % 2 == 0 && tedm_num_sms() >= 2 && a->mode != 0 && ktot / p.taps >= 128) ? 2 : 1;
  if (g_enable_pairs == 3 && bn == 256) p.cg = 1;                           // A/B: pairs on N <= 128 tiles only
  if (g_enable_pairs == 2 && m_tiles % 2 == 0) p.cg = 2;                    // forced (tests)
  if (ws4) p.cg = 1;

  alignas(64) ConvMaps maps;
  CUtensorMap& mapW = maps.w;
  CUtensorMap& mapOut = maps.out;
  CUtensorMap& mapOut2 = maps.out2;
  const int boxW = ws ? ROW_PIX : (halo ? HALO_W + 2 : p.tileW);
  if (w64) {                                 // M = one 64-pixel row of two images
    p.tileH = 1;
    p.tileB = 2;
  }
  int rc = TEDM_OK;
  for (int i = 0; i < MAX_SRC; ++i) {
    if (i < p.n_src) {
      rc = encode_act_map(&maps.a[i], src_ptr[i], a->batch, a->height, a->width, p.src_C[i], src_stride[i], a->mode, boxW,
                          halo ? HALO_H + 2 : p.tileH, p.tileB);
      if (rc) return rc;
    } else {
      maps.a[i] = maps.a[0];
    }
  }
  rc = encode_weight_map(&mapW, a->weight, (long long)zdim * a->cout, ktot, bn / p.cg);
  if (rc) return rc;
  if (!p.out_f32) {
    // bf16 outputs leave through a TMA store; the upsample mode scatters each parity through the same
    // (2C, W, 2, H, B) view the stride-2 mode uses for its input
    rc = encode_act_map(&mapOut, a->out, a->batch, p.OH, p.OW, a->split ? a->split : a->cout, p.out_image_stride,
                        a->mode == 3 ? 2 : 0, p.tileW, p.tileH, p.tileB);
    if (rc) return rc;
  } else {
    mapOut = maps.a[0];
  }
  if (a->split) {
    rc = encode_act_map(&mapOut2, a->out2, a->batch, p.OH, p.OW, a->cout - a->split, p.out2_image_stride, 0, p.tileW, p.tileH,
                        p.tileB);
    if (rc) return rc;
  } else {
    mapOut2 = mapOut;
  }
  // 1x1 convs with a residual: the residual tile travels by TMA into the output staging tile (launch_conv_cg decides whether
  // three staging buffers fit)
  maps.res = mapOut;
  if (g_enable_res_tma && a->mode == 0 && a->residual && !a->split && !p.out_f32 && bn <= 128) {
    rc = encode_act_map(&maps.res, a->residual, a->batch, p.OH, p.OW, a->cout, p.out_image_stride, 0, p.tileW, p.tileH, p.tileB);
    if (rc) return rc;
    p.res_tma = 1;
  }

  cudaStream_t s = (cudaStream_t)stream;
  if (ws4) {
    Ws4Params q{};
    q.B = p.B;
    q.tiles_x = w64 ? 1 : p.tiles_x;
    q.rows4 = p.Ho / 4;
    q.num_tiles = (w64 ? (p.B + 1) / 2 : p.B) * q.rows4 * q.tiles_x;
    q.c0_blocks = p.c0_blocks;
    q.bias = p.bias;
    q.gn_partial = p.gn_partial;
    q.gn_groups = p.gn_groups;
    q.gn_parts = p.gn_parts;
    q.H = p.Ho;
    q.W = p.Wo;
    q.src_affine = p.src_affine;
    const CUtensorMap& a1 = maps.a[p.n_src > 1 ? 1 : 0];
    if (w64) return launch_ws4_cpg<1, true>(maps.a[0], a1, mapW, mapOut, q, s);
    if (a->c0 + a->c1 == 64) return launch_ws4_cpg<1, false>(maps.a[0], a1, mapW, mapOut, q, s);
    return launch_ws4_cpg<2, false>(maps.a[0], a1, mapW, mapOut, q, s);
  }
  if (ws) return launch_conv<64, 1>(maps, p, s);
  if (halo && p.src_affine) {
    switch (bn) {
      case 64: return launch_conv<64, 3>(maps, p, s);
      case 128: return launch_conv<128, 3>(maps, p, s);
      default: return launch_conv<256, 3>(maps, p, s);
    }
  }
  if (halo) {
    switch (bn) {
      case 64: return launch_conv<64, 2>(maps, p, s);
      case 128: return launch_conv<128, 2>(maps, p, s);
      default: return launch_conv<256, 2>(maps, p, s);
    }
  }
  switch (bn) {
    case 64: return launch_conv<64, 0>(maps, p, s);
    case 128: return launch_conv<128, 0>(maps, p, s);
    default: return launch_conv<256, 0>(maps, p, s);
  }
}

// dw: fp32 [cout][taps][c0+c1] (taps = 1 / 9 / 16; mode 3: 16 = parity*4 + a*2 + b of the folded kernel), overwritten;
// or, with oihw_accumulate, the fp32 OIHW parameter gradient itself (+=; the folded taps of mode 3 are scattered to 3x3).
extern "C" int64_t tedm_conv_igemm_wgrad_workspace(void) {
  // split-K partial tiles of the halo-tile 3x3 kernel, the raw intermediate of the folded upsample conv's gradient, and
  // the generic kernel's turn counters (which must be ZERO when the workspace is first used; they reset themselves)
  return WGRAD_PART_FLOATS_PER_SM * tedm_num_sms() + WGRAD_RAW_FLOATS + WGRAD_TURN_INTS;
}

extern "C" int tedm_conv_igemm_wgrad(const tedm_conv_args* a, const void* dy, float* dw, int oihw_accumulate,
                                     float* workspace, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && a->src0 && dy && dw, "tedm_conv_igemm_wgrad: null pointer");
  TEDM_CHECK_ARG(a->mode >= 0 && a->mode <= 3, "tedm_conv_igemm_wgrad: mode=%d", a->mode);
  TEDM_CHECK_ARG(a->batch > 0 && a->height > 0 && a->width > 0 && a->c0 > 0 && a->c1 >= 0 && a->cout > 0,
                 "tedm_conv_igemm_wgrad: bad sizes");
  TEDM_CHECK_ARG((a->c1 > 0) == (a->src1 != nullptr), "tedm_conv_igemm_wgrad: src1/c1 mismatch");
  TEDM_UNSUPPORTED(a->c0 % BK != 0 || a->c1 % BK != 0 || a->cout % 64 != 0,
                   "tedm_conv_igemm_wgrad: channel counts (%d, %d -> %d) must be multiples of 64", a->c0, a->c1, a->cout);
  TEDM_UNSUPPORTED(a->mode == 2 && a->c1 != 0, "tedm_conv_igemm_wgrad: stride-2 mode takes one source");
  WgradParams p{};
  p.mode = a->mode;
  p.kxc = a->mode == 0 ? 1 : a->mode == 1 ? 3 : 4;
  p.taps = a->mode == 0 ? 1 : a->mode == 1 ? 9 : 16;
  p.C0 = a->c0;
  p.C1 = a->c1;
  p.ctot = a->c0 + a->c1;
  p.c0_blocks = a->c0 / BK;
  p.c1_blocks = a->c1 / BK;
  p.B = a->batch;
  p.cout = a->cout;
  const int Ho = a->mode == 2 ? a->height / 2 : a->height, Wo = a->mode == 2 ? a->width / 2 : a->width;  // tile space
  TEDM_UNSUPPORTED(!is_pow2(Ho) || !is_pow2(Wo) || (a->mode == 2 && ((a->height | a->width) & 1)) || (long long)Ho * Wo < 16,
                   "tedm_conv_igemm_wgrad: spatial extent %dx%d unsupported", a->height, a->width);
  p.tileW = Wo < BM ? Wo : BM;
  p.tileH = Ho < BM / p.tileW ? Ho : BM / p.tileW;
  p.tileB = BM / (p.tileW * p.tileH);
  p.tiles_x = Wo / p.tileW;
  p.tiles_y = Ho / p.tileH;
  p.num_ptiles = ceil_div(p.B, p.tileB) * p.tiles_x * p.tiles_y;
  p.items = p.taps * (p.c0_blocks + p.c1_blocks);
  p.dw = dw;
  p.oihw = oihw_accumulate != 0;
  cudaStream_t s = (cudaStream_t)stream;

  // ---- 3x3: halo-tile kernel (every tap from one shared-memory box); needs 16-pixel K runs inside image rows, or
  //      8-pixel-wide tiles with an even number of rows.  Measured on B200 (profiles/r01_wgrad_kernels.txt): 790-960
  //      TFLOP/s on the 64-channel layers against 400-560 for the generic kernel; only the widest layer (768 -> 512),
  //      where the generic kernel runs N = 256 tiles, stays on the generic kernel.
  const bool halo_geometry = a->mode == 1 && (p.tileW >= 16 || (p.tileW == 8 && p.tileH % 2 == 0));
  const bool halo_pays = (long long)p.ctot * a->cout < 768LL * 512;
  if (halo_geometry && (g_enable_wgrad3 == 2 || (g_enable_wgrad3 == 1 && halo_pays))) {
    Wgrad3Params w{};
    w.c0_blocks = p.c0_blocks;
    w.c1_blocks = p.c1_blocks;
    w.ctot = p.ctot;
    w.tileW = p.tileW; w.tileH = p.tileH; w.tileB = p.tileB; w.tiles_x = p.tiles_x; w.tiles_y = p.tiles_y;
    w.B = p.B;
    w.cout = a->cout;
    w.n_tiles = a->cout / 64;
    w.num_ptiles = p.num_ptiles;
    w.oihw = p.oihw;
    w.dw = dw;
    w.x_tx_bytes = p.tileB * (p.tileH + 2) * (p.tileW + 2) * BK * 2;
    w.x_bytes = (w.x_tx_bytes + 1023) / 1024 * 1024;
    const int stage_bytes = w.x_bytes + A_BYTES;
    int stages = (DYN_SMEM_MAX - 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const long long groups3 = (long long)(p.c0_blocks + p.c1_blocks) * (a->cout / 64);
    // layers too wide for the partial-tile workspace even unsplit go to the generic kernel (which batches its groups)
    if (stages >= 2 && groups3 * 5 * BM * 64 <= WGRAD_PART_FLOATS_PER_SM * tedm_num_sms()) {
      w.stages = stages;
      const long long groups = (long long)(p.c0_blocks + p.c1_blocks) * w.n_tiles, sms = tedm_num_sms();
      const long long cap = p.num_ptiles / 4 > 0 ? p.num_ptiles / 4 : 1;
      long long best_sk = 1;
      double best_eff = -1.0;
      for (int wv = 1; wv <= 3; ++wv) {        // fewest waves that fill the machine (every CTA emits a 160 KB partial)
        long long sk = (wv * sms) / groups;
        if (sk < 1) sk = 1;
        if (sk > cap) sk = cap;
        while (sk > 1 && groups * sk * 5 * BM * 64 > WGRAD_PART_FLOATS_PER_SM * sms) --sk;
        const long long ctas = groups * sk, waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best_eff + 0.02) {
          best_eff = eff;
          best_sk = sk;
        }
      }
      w.splitk = (int)best_sk;
      alignas(64) CUtensorMap mX0, mX1, mDY;
      int rc3 = encode_act_map(&mX0, a->src0, a->batch, a->height, a->width, a->c0,
                               a->src0_image_stride ? a->src0_image_stride : (long long)a->height * a->width * a->c0, 0,
                               p.tileW + 2, p.tileH + 2, p.tileB);
      if (rc3) return rc3;
      if (a->src1) {
        rc3 = encode_act_map(&mX1, a->src1, a->batch, a->height, a->width, a->c1,
                             a->src1_image_stride ? a->src1_image_stride : (long long)a->height * a->width * a->c1, 0,
                             p.tileW + 2, p.tileH + 2, p.tileB);
        if (rc3) return rc3;
      } else {
        mX1 = mX0;
      }
      rc3 = encode_act_map(&mDY, dy, a->batch, Ho, Wo, a->cout,
                           a->out_image_stride ? a->out_image_stride : (long long)Ho * Wo * a->cout, 0, p.tileW, p.tileH, p.tileB);
      if (rc3) return rc3;
      w.part = workspace;     // sized by tedm_conv_igemm_wgrad_workspace(): groups * splitk <= 3 waves of CTAs
      const int smem = 1024 + stages * stage_bytes;
      static int configured3 = 0;
      if (configured3 < smem) {
        TEDM_CUDA(cudaFuncSetAttribute(conv_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured3 = smem;
      }
      dim3 grid((unsigned)w.splitk, (unsigned)groups);
      conv_wgrad3_kernel<<<grid, 192, smem, s>>>(mX0, mX1, mDY, w);
      TEDM_LAUNCH_CHECK();
      {
        if (w.splitk <= 16 && p.oihw)
          wgrad3_reduce_oihw_kernel<<<dim3(16, (unsigned)groups), 256, 0, s>>>(w.part, dw, w.splitk, w.n_tiles, p.ctot);
        else if (w.splitk > 16)
          wgrad3_reduce_kernel<4><<<dim3(64, (unsigned)groups, 9), 256, 0, s>>>(w.part, dw, w.splitk, w.n_tiles, p.ctot, p.oihw);
        else
          wgrad3_reduce_kernel<1><<<dim3(16, (unsigned)groups, 9), 256, 0, s>>>(w.part, dw, w.splitk, w.n_tiles, p.ctot, p.oihw);
        TEDM_LAUNCH_CHECK();
      }
      return TEDM_OK;
    }
  }

  int bn = a->cout % 256 == 0 ? 256 : (a->cout % 128 == 0 ? 128 : 64);
  if (g_force_bn && a->cout % g_force_bn == 0) bn = g_force_bn;
  p.n_tiles = a->cout / bn;
  const int m_blocks = (p.items + 1) / 2;
  const long long groups = (long long)m_blocks * p.n_tiles, sms = tedm_num_sms();
  const long long part_floats = WGRAD_PART_FLOATS_PER_SM * sms;
  // split K over pixel tiles: one CTA per SM is resident (192 KB of pipeline smem), so pick the split whose CTA count
  // fills whole waves of the machine (up to 4 waves, each CTA at least 4 pixel tiles)
  {
    const long long cap = p.num_ptiles / 4 > 0 ? p.num_ptiles / 4 : 1;
    long long best_sk = 1;
    double best_eff = -1.0;
    for (int w = 1; w <= 4; ++w) {
      long long sk = (w * sms) / groups;
      if (sk < 1) sk = 1;
      if (sk > cap) sk = cap;
      if (g_deterministic && sk > 4) sk = 4;      // ordered slices add one after the other: keep the chains short
      const long long ctas = groups * sk, waves = (ctas + sms - 1) / sms;
      const double eff = (double)ctas / (double)(waves * sms);
      if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && sk > best_sk)) {
        best_eff = eff;
        best_sk = sk;
      }
    }
    p.splitk = (int)best_sk;
  }
  TEDM_CHECK_ARG(workspace != nullptr, "tedm_conv_igemm_wgrad: workspace (tedm_conv_igemm_wgrad_workspace() floats) is required");
  const bool unfold = p.oihw && a->mode == 3;     // folded taps -> raw [co][16][ci] intermediate -> 3x3 OIHW accumulate
  TEDM_UNSUPPORTED(unfold && (long long)a->cout * 16 * p.ctot > WGRAD_RAW_FLOATS,
                   "tedm_conv_igemm_wgrad: upsample conv %d -> %d too wide for the raw-gradient workspace", p.ctot, a->cout);
  TEDM_UNSUPPORTED(groups > WGRAD_TURN_INTS, "tedm_conv_igemm_wgrad: %lld output tiles exceed the turn-counter workspace", groups);
  float* raw = workspace + part_floats;
  p.turn = g_deterministic ? reinterpret_cast<int*>(workspace + part_floats + WGRAD_RAW_FLOATS) : nullptr;   // zero at allocation, self-resetting
  if (unfold) {
    p.dw = raw;
    p.oihw = 0;
  }

  const int dy_h = a->mode == 3 ? 2 * a->height : Ho, dy_w = a->mode == 3 ? 2 * a->width : Wo;
  alignas(64) CUtensorMap mapX0, mapX1, mapDY;
  int rc = encode_act_map(&mapX0, a->src0, a->batch, a->height, a->width, a->c0,
                          a->src0_image_stride ? a->src0_image_stride : (long long)a->height * a->width * a->c0, a->mode,
                          p.tileW, p.tileH, p.tileB);
  if (rc) return rc;
  if (a->src1) {
    rc = encode_act_map(&mapX1, a->src1, a->batch, a->height, a->width, a->c1,
                        a->src1_image_stride ? a->src1_image_stride : (long long)a->height * a->width * a->c1, a->mode,
                        p.tileW, p.tileH, p.tileB);
    if (rc) return rc;
  } else {
    mapX1 = mapX0;
  }
  rc = encode_act_map(&mapDY, dy, a->batch, dy_h, dy_w, a->cout,
                      a->out_image_stride ? a->out_image_stride : (long long)dy_h * dy_w * a->cout, a->mode == 3 ? 2 : 0,
                      p.tileW, p.tileH, p.tileB);
  if (rc) return rc;
  if (!p.oihw) TEDM_CUDA(cudaMemsetAsync(p.dw, 0, sizeof(float) * (size_t)a->cout * p.taps * p.ctot, s));
  switch (bn) {
    case 64: rc = launch_wgrad<64>(mapX0, mapX1, mapDY, p, m_blocks, s); break;
    case 128: rc = launch_wgrad<128>(mapX0, mapX1, mapDY, p, m_blocks, s); break;
    default: rc = launch_wgrad<256>(mapX0, mapX1, mapDY, p, m_blocks, s); break;
  }
  if (rc) return rc;
  if (unfold) return tedm_wgrad_to_oihw(raw, dw, a->cout, p.ctot, 3, stream);
  return TEDM_OK;
}
