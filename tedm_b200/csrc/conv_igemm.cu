// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixel][cout] = sum_{tap, cin} A[pixel + tap offset][cin] * W[cout][tap][cin]
//
//   M = 128 output pixels per CTA (a tileB x tileH x tileW box of the NHWC output),
//   N = BN output channels (64/128/256), K = taps * (c0 + c1) in blocks of 64 channels.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> bias/residual/GroupNorm partials -> bf16 global).
// A tiles are fetched by 5-D tiled TMA straight from the NHWC activation tensor: the box is
// (64 ch, tileW, 1, tileH, tileB) at a per-tap pixel offset, out-of-bounds pixels (the conv's zero
// padding, and batch overhang) are zero-filled by the TMA unit, and the 128B swizzle makes the
// landed box exactly the K-major UMMA operand layout.  The skip-connection concat of the UNet
// decoder is a second A tensor map (the K loop walks src0's channels, then src1's); the stride-2
// 4x4 Downsample reads a (2C, W/2, 2, H/2, B) view of the same memory so every tap is again a
// dense box; the nearest-x2 Upsample+3x3 runs as four parity 2x2 convs on the source resolution.
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
constexpr int NGMAX = 32;        // max GroupNorm groups touched by one N tile

struct ConvParams {
  int mode, taps, kxc;           // kxc = taps per filter row
  int c0_blocks, c1_blocks, C0, C1;
  int tileW, tileH, tileB, tiles_x, tiles_y;
  int B, Ho, Wo;                 // tile space extent (output pixels; source pixels for mode 3)
  int OH, OW, osy, osx;          // output tensor extent and tile->output coordinate scale
  int cout;
  const float* bias;
  const bf16* residual;
  bf16* out;
  int out_f32;                   // 1: `out` is fp32 (direct stores); 0: bf16 through smem + TMA store
  float* gn_partial;
  int gn_cpg, gn_groups, gn_parts;
  long long out_image_stride;    // elements
};

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  // K-major, 128B swizzle: 8-row atoms of 1024 B (SBO), LBO unused, descriptor version 1 (sm_100).
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int W>
__device__ __forceinline__ void gn_accumulate(const float (&v)[32], int lane, int seg_size, float* red_row, int gl0) {
  // W = channels per partial sum inside this 32-column chunk (8, 16 or 32)
#pragma unroll
  for (int sg = 0; sg < 32 / W; ++sg) {
    float s = 0.0f, q = 0.0f;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      s += v[sg * W + j];
      q = fmaf(v[sg * W + j], v[sg * W + j], q);
    }
    for (int o = seg_size >> 1; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((lane & (seg_size - 1)) == 0) {
      float* r = red_row + (size_t)(gl0 + sg) * 2;
      r[0] += s;
      r[1] += q;
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapOut,
                  const ConvParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[4][2][NGMAX][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // ---- tile coordinates -----------------------------------------------------------------
  const int tiles_per_group = p.tiles_x * p.tiles_y;
  const int bg = blockIdx.x / tiles_per_group;
  const int trem = blockIdx.x % tiles_per_group;
  const int b0 = bg * p.tileB, y0 = (trem / p.tiles_x) * p.tileH, x0 = (trem % p.tiles_x) * p.tileW;
  const int n0 = blockIdx.y * BN;
  const int par = blockIdx.z, par_y = par >> 1, par_x = par & 1;  // mode 3 only
  const int num_kb = p.taps * (p.c0_blocks + p.c1_blocks);

  for (int i = threadIdx.x; i < 4 * 2 * NGMAX * 2; i += blockDim.x) (&red[0][0][0][0])[i] = 0.0f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.c1_blocks) tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapW);
    if (!p.out_f32) tma_prefetch_desc(&mapOut);
    for (int s = 0; s < STAGES; ++s) {
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
    // ===== TMA producer =====================================================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int cb_total = p.c0_blocks + p.c1_blocks;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        const uint32_t full = smem_u32(&bar_full[stage]);
        mbar_expect_tx(full, STAGE_BYTES);
        const int tap = kb / cb_total, cb = kb % cb_total;
        const int ky = tap / p.kxc, kx = tap % p.kxc;
        int offy = 0, offx = 0, pc = 0, chan_off = 0;
        const bool second = cb >= p.c0_blocks;
        const int cblk = second ? cb - p.c0_blocks : cb;
        if (p.mode == 1) {
          offy = ky - 1;
          offx = kx - 1;
        } else if (p.mode == 2) {
          offy = ((ky + 1) >> 1) - 1;
          offx = ((kx + 1) >> 1) - 1;
          pc = (ky + 1) & 1;
          chan_off = ((kx + 1) & 1) * (second ? p.C1 : p.C0);
        } else if (p.mode == 3) {
          offy = ky - 1 + par_y;
          offx = kx - 1 + par_x;
        }
        const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
        tma_load_5d(a_dst, second ? &mapA1 : &mapA0, full, chan_off + cblk * BK, x0 + offx, pc, y0 + offy, b0);
        tma_load_2d(a_dst + A_BYTES, &mapW, full, kb * BK, par * p.cout + n0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =========================================================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
        const uint64_t adesc = make_sw128_desc(a_addr), bdesc = make_sw128_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, adesc + 2ull * k, bdesc + 2ull * k, IDESC, (kb | k) != 0 ? 1u : 0u);
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(smem_u32(&bar_acc));
    }
    __syncwarp();
  } else {
    // ===== epilogue ===========================================================================
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int hwt = p.tileW * p.tileH;    // pixels of one image inside the tile
    const int tb = row / hwt, rrem = row % hwt;
    const int ty = rrem / p.tileW, tx = rrem % p.tileW;
    const int b = b0 + tb, y = y0 + ty, x = x0 + tx;
    const bool valid = b < p.B;
    const int oy = y * p.osy + par_y, ox = x * p.osx + par_x;
    const size_t opix = (size_t)b * p.out_image_stride + ((size_t)oy * p.OW + ox) * p.cout + n0;
    const int seg_size = hwt < 32 ? hwt : 32;
    float* red_row = &red[q][lane / seg_size][0][0];

    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chunk * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.bias) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + chunk * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = __ldg(bp + j);
          v[4 * j] += bv.x; v[4 * j + 1] += bv.y; v[4 * j + 2] += bv.z; v[4 * j + 3] += bv.w;
        }
      }
      if (p.gn_partial) {
        const int gl0 = (chunk * 32) / p.gn_cpg;
        if (p.gn_cpg == 8) gn_accumulate<8>(v, lane, seg_size, red_row, gl0);
        else if (p.gn_cpg == 16) gn_accumulate<16>(v, lane, seg_size, red_row, gl0);
        else gn_accumulate<32>(v, lane, seg_size, red_row, gl0);
      }
      if (p.residual && valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + opix + chunk * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[8];
          unpack8(__ldg(rp + j), f);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[8 * j + e] += f[e];
        }
      }
      if (p.out_f32) {
        if (valid) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix + chunk * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else {
        // stage the bf16 tile in the (now idle) pipeline buffers in the 128B-swizzled layout TMA expects:
        // sub-tile = 64 channels, row = pixel (128 B), 16-byte chunk index XOR (row & 7)
        const uint32_t sub = smem_base + (uint32_t)(chunk >> 1) * A_BYTES + (uint32_t)row * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c16 = (uint32_t)((chunk & 1) * 4 + j);
          const uint32_t dst = sub + ((c16 ^ (uint32_t)(row & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])),
                       "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                       "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7])) : "memory");
        }
      }
    }
    if (!p.out_f32) {
      fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(2, 128);
      if (threadIdx.x == 64) {
        const int chan_base = (p.mode == 3 ? par_x * p.cout : 0) + n0;
        const int pc_out = p.mode == 3 ? par_y : 0;
        for (int sidx = 0; sidx < BN / 64; ++sidx)
          tma_store_5d(&mapOut, smem_base + (uint32_t)sidx * A_BYTES, chan_base + sidx * 64, x0, pc_out, y0, b0);
        tma_store_commit();
        tma_store_wait_read<0>();         // smem must outlive the bulk read
      }
    }
    if (p.gn_partial) {
      named_bar_sync(1, 128);
      const int e = threadIdx.x - 64;               // 0..127
      const int ngl = BN / p.gn_cpg;                // groups covered by this N tile
      const int segs_per_img = hwt / seg_size;      // 4, 2 or 1
      const int nimg = p.tileB;
      const int part = hwt == BM ? trem : 0;
      for (int o = e; o < nimg * ngl; o += 128) {
        const int img = o / ngl, gl = o % ngl;
        if (b0 + img >= p.B) continue;
        float s = 0.0f, qq = 0.0f;
        for (int sgi = 0; sgi < segs_per_img; ++sgi) {
          const int gseg = img * segs_per_img + sgi;       // global segment index in the tile
          const int w = (gseg * seg_size) >> 5, sl = ((gseg * seg_size) & 31) / seg_size;
          s += red[w][sl][gl][0];
          qq += red[w][sl][gl][1];
        }
        float* dst = p.gn_partial + (((size_t)(b0 + img) * p.gn_parts + part) * p.gn_groups + n0 / p.gn_cpg + gl) * 2;
        dst[0] = s;
        dst[1] = qq;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tmem_base);
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

template <int BN, int STAGES>
int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o,
                const ConvParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (A_BYTES + BN * BK * 2) + 1024;
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  conv_igemm_kernel<BN, STAGES><<<grid, 192, smem, stream>>>(a0, a1, w, o, p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

int g_force_bn = 0;  // debug/tuning override (tedm_conv_set_tile_n)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" int tedm_conv_set_tile_n(int bn) {
  TEDM_CHECK_ARG(bn == 0 || bn == 64 || bn == 128 || bn == 256, "tedm_conv_set_tile_n: bn=%d", bn);
  g_force_bn = bn;
  return TEDM_OK;
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
  TEDM_UNSUPPORTED(a->mode == 2 && a->c1 != 0, "tedm_conv_igemm_fwd: stride-2 mode takes one source");

  ConvParams p{};
  p.mode = a->mode;
  p.kxc = a->mode == 0 ? 1 : a->mode == 1 ? 3 : a->mode == 2 ? 4 : 2;
  p.taps = p.kxc * p.kxc;
  p.C0 = a->c0;
  p.C1 = a->c1;
  p.c0_blocks = a->c0 / BK;
  p.c1_blocks = a->c1 / BK;
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
  p.cout = a->cout;
  p.bias = a->bias;
  p.residual = (const bf16*)a->residual;
  p.out = (bf16*)a->out;
  p.out_f32 = a->out_dtype == 1;
  TEDM_CHECK_ARG(a->out_dtype == 0 || a->out_dtype == 1, "tedm_conv_igemm_fwd: out_dtype=%d", a->out_dtype);
  p.out_image_stride = a->out_image_stride ? a->out_image_stride : (long long)p.OH * p.OW * a->cout;
  p.gn_partial = a->gn_partial;
  if (a->gn_partial) {
    TEDM_CHECK_ARG(a->gn_groups > 0 && a->cout % a->gn_groups == 0, "tedm_conv_igemm_fwd: gn_groups=%d", a->gn_groups);
    TEDM_UNSUPPORTED(a->mode == 3, "tedm_conv_igemm_fwd: GroupNorm partials are not produced in upsample mode");
    p.gn_groups = a->gn_groups;
    p.gn_cpg = a->cout / a->gn_groups;
    p.gn_parts = tedm_conv_gn_parts(p.Ho, p.Wo);
    TEDM_UNSUPPORTED(p.gn_cpg % 8 != 0 || (p.gn_cpg < 32 && 32 % p.gn_cpg != 0) || (p.gn_cpg > 32 && p.gn_cpg % 32 != 0),
                     "tedm_conv_igemm_fwd: %d channels per GroupNorm group unsupported", p.gn_cpg);
  }

  // ---- N tile: the widest that divides cout, covers whole GroupNorm groups and still fills the SMs
  const long long m_tiles = (long long)ceil_div(p.B, p.tileB) * p.tiles_x * p.tiles_y;
  const int zdim = a->mode == 3 ? 4 : 1;
  int bn = 0;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3 && bn == 0; ++i) {
    const int c = cands[i];
    if (a->cout % c) continue;
    if (p.gn_partial && (c % p.gn_cpg != 0 || c / p.gn_cpg > NGMAX)) continue;
    if (m_tiles * zdim * (a->cout / c) >= tedm_num_sms() || c == 64) bn = c;
  }
  if (bn == 0) {  // nothing fills the machine: take the narrowest legal tile
    for (int i = 2; i >= 0 && bn == 0; --i)
      if (a->cout % cands[i] == 0 && (!p.gn_partial || (cands[i] % p.gn_cpg == 0 && cands[i] / p.gn_cpg <= NGMAX))) bn = cands[i];
  }
  if (g_force_bn && a->cout % g_force_bn == 0 && (!p.gn_partial || (g_force_bn % p.gn_cpg == 0 && g_force_bn / p.gn_cpg <= NGMAX)))
    bn = g_force_bn;
  TEDM_UNSUPPORTED(bn == 0, "tedm_conv_igemm_fwd: no N tile for cout=%d with %d-channel GroupNorm groups", a->cout, p.gn_cpg);
  TEDM_CHECK_ARG(m_tiles <= 2147483647LL, "tedm_conv_igemm_fwd: too many tiles");

  alignas(64) CUtensorMap mapA0, mapA1, mapW, mapOut;
  int rc = encode_act_map(&mapA0, a->src0, a->batch, a->height, a->width, a->c0,
                          a->src0_image_stride ? a->src0_image_stride : (long long)a->height * a->width * a->c0, a->mode,
                          p.tileW, p.tileH, p.tileB);
  if (rc) return rc;
  if (a->src1) {
    rc = encode_act_map(&mapA1, a->src1, a->batch, a->height, a->width, a->c1,
                        a->src1_image_stride ? a->src1_image_stride : (long long)a->height * a->width * a->c1, a->mode,
                        p.tileW, p.tileH, p.tileB);
    if (rc) return rc;
  } else {
    mapA1 = mapA0;
  }
  const long long ktot = (long long)p.taps * (a->c0 + a->c1);
  rc = encode_weight_map(&mapW, a->weight, (long long)zdim * a->cout, ktot, bn);
  if (rc) return rc;

  if (!p.out_f32) {
    // bf16 outputs leave through a TMA store; the upsample mode scatters each parity through the same
    // (2C, W, 2, H, B) view the stride-2 mode uses for its input
    rc = encode_act_map(&mapOut, a->out, a->batch, p.OH, p.OW, a->cout, p.out_image_stride, a->mode == 3 ? 2 : 0, p.tileW,
                        p.tileH, p.tileB);
    if (rc) return rc;
  } else {
    mapOut = mapA0;
  }

  dim3 grid((unsigned)m_tiles, (unsigned)(a->cout / bn), (unsigned)zdim);
  cudaStream_t s = (cudaStream_t)stream;
  switch (bn) {
    case 64: return launch_conv<64, 4>(mapA0, mapA1, mapW, mapOut, p, grid, s);
    case 128: return launch_conv<128, 3>(mapA0, mapA1, mapW, mapOut, p, grid, s);
    default: return launch_conv<256, 4>(mapA0, mapA1, mapW, mapOut, p, grid, s);
  }
}
