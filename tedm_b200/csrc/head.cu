// TEDM / LEDM per-pixel MLP head, commuted form.
//
// Reference (models/datasetDM_model.py:57-64,80-88; trainers/train_datasetDM.py:30-42): nearest-
// upsample the four decoder maps to full resolution, concatenate to 960 (x S) channels, then
// Conv1x1(->128) -> ReLU -> BN -> Conv1x1(->32) -> ReLU -> BN -> Conv1x1(->1).  A 1x1 conv commutes
// with nearest upsampling, so layer 1 runs per level at native resolution on the tcgen05 GEMM
// (tedm_conv_igemm_fwd, mode 0) and this kernel finishes the job per output pixel:
//   z1 = b1 + sum_{s < n_sum} sum_l g_l[img*n_sum + s][y >> sh_l][x >> sh_l]      (128 values)
//   h1 = bn1(relu(z1)) ; z2 = W2 h1 + b2 ; h2 = bn2(relu(z2)) ; logit = w3 . h2 + b3
// with eval-mode BatchNorm folded to an affine.  The 960-channel full-resolution tensor of the
// reference (62.9 MB per image and step) is never materialised.
#include "common.cuh"

#define HEAD_C1 128
#define HEAD_C2 32

struct HeadParams {
  const void* g[4];
  int shift[4];
  int n_levels, n_sum, n_img, H, W;
  const float *b1, *a1, *c1, *w2, *b2, *a2, *c2, *w3;
  float b3;
  float* logits;
};

template <bool G_F32>
__global__ void __launch_bounds__(128) head_tail_kernel(const HeadParams p) {
  __shared__ __align__(16) float sw2[HEAD_C1][HEAD_C2];  // transposed: [k][j]
  __shared__ float sb1[HEAD_C1], sa1[HEAD_C1], sc1[HEAD_C1];
  for (int i = threadIdx.x; i < HEAD_C1 * HEAD_C2; i += blockDim.x) {
    const int j = i / HEAD_C1, k = i % HEAD_C1;  // w2 is [c2][c1]
    sw2[k][j] = p.w2[i];
  }
  for (int i = threadIdx.x; i < HEAD_C1; i += blockDim.x) {
    sb1[i] = p.b1[i];
    sa1[i] = p.a1[i];
    sc1[i] = p.c1[i];
  }
  __syncthreads();
  const long long npix = (long long)p.n_img * p.H * p.W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % p.W), y = (int)((pix / p.W) % p.H);
    const long long img = pix / ((long long)p.W * p.H);
    float acc[HEAD_C2];
#pragma unroll
    for (int j = 0; j < HEAD_C2; ++j) acc[j] = p.b2[j];
    for (int k0 = 0; k0 < HEAD_C1; k0 += 8) {
      float z[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = sb1[k0 + e];
      for (int s = 0; s < p.n_sum; ++s) {
        for (int l = 0; l < p.n_levels; ++l) {
          const int sh = p.shift[l];
          const int hl = p.H >> sh, wl = p.W >> sh;
          const size_t off = ((((size_t)img * p.n_sum + s) * hl + (y >> sh)) * wl + (x >> sh)) * HEAD_C1 + k0;
          float f[8];
          if (G_F32) {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.g[l]) + off);
            const float4 u0 = __ldg(src), u1 = __ldg(src + 1);
            f[0] = u0.x; f[1] = u0.y; f[2] = u0.z; f[3] = u0.w; f[4] = u1.x; f[5] = u1.y; f[6] = u1.z; f[7] = u1.w;
          } else {
            unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.g[l]) + off)), f);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) z[e] += f[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float hval = fmaf(fmaxf(z[e], 0.0f), sa1[k0 + e], sc1[k0 + e]);
        const float4* wr = reinterpret_cast<const float4*>(&sw2[k0 + e][0]);
#pragma unroll
        for (int j4 = 0; j4 < HEAD_C2 / 4; ++j4) {
          const float4 w = wr[j4];
          acc[j4 * 4] = fmaf(w.x, hval, acc[j4 * 4]);
          acc[j4 * 4 + 1] = fmaf(w.y, hval, acc[j4 * 4 + 1]);
          acc[j4 * 4 + 2] = fmaf(w.z, hval, acc[j4 * 4 + 2]);
          acc[j4 * 4 + 3] = fmaf(w.w, hval, acc[j4 * 4 + 3]);
        }
      }
    }
    float logit = p.b3;
#pragma unroll
    for (int j = 0; j < HEAD_C2; ++j) logit = fmaf(p.w3[j], fmaf(fmaxf(acc[j], 0.0f), p.a2[j], p.c2[j]), logit);
    p.logits[pix] = logit;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Tensor-core form of the tail (fp32 g maps).  One warp owns 16 consecutive output pixels end to end:
//   [FUSE] z1 = f_full W1_full^T            the full-resolution level's layer 1 (K = 64) on mma.sync, so that level's
//                                           fp32 map (512 B per pixel written and read back) never exists
//   z1 += b1 + gathered g_l                 fragment-layout float2 loads of the coarser levels (L1-resident)
//   a1 = bn1(relu(z1)) -> (hi, lo) bf16     accumulator fragments ARE the next product's A fragments
//   z2 = a1 W2^T                            3 MMAs per tile (hi*hi + lo*hi + hi*lo): fp32-grade, W2 split at load
//   logit = w3 . bn2(relu(z2 + b2)) + b3    quad reduction
// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int kW2Pitch = HEAD_C1 * 2 + 16;   // bytes per W2 row
constexpr int kFPitch = 64 * 2 + 16;         // bytes per row of the 64-channel full-resolution tile / W1_full row
constexpr int kHeadWarps = 8;

__device__ __forceinline__ void h_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void h_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pack(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 xh = __float2bfloat16_rn(x), yh = __float2bfloat16_rn(y);
  hi = pack_bf16x2(x, y);
  lo = pack_bf16x2(x - __bfloat162float(xh), y - __bfloat162float(yh));
}

struct HeadFull {
  const bf16* f;    // [n_img][H][W][64] bf16 (n_sum must be 1)
  const bf16* w1;   // [128][64] bf16
};

// STAGED (n_sum == 1, every g level has 1 <= shift <= 3, W % 16 == 0): the 14 source rows (8 + 4 + 2 pixels x 512 B) a
// 16-pixel group gathers from are copied into shared memory with cp.async while the layer-1 product runs, instead of
// being fetched lane by lane afterwards (the kernel was waiting on those loads for two thirds of its time).
constexpr int kGPitch = HEAD_C1 * 4 + 16;    // bytes per staged g row
constexpr int kGRows = 14;
template <bool FUSE, bool STAGED>
__global__ void __launch_bounds__(kHeadWarps * 32) head_tail_mma_kernel(const HeadParams p, const HeadFull full) {
  extern __shared__ __align__(16) uint8_t hsm[];
  uint8_t* w2h_s = hsm;                                   // [32][kW2Pitch]
  uint8_t* w2l_s = w2h_s + HEAD_C2 * kW2Pitch;
  float* par = reinterpret_cast<float*>(w2l_s + HEAD_C2 * kW2Pitch);   // b1, a1, c1 [128]; b2, a2, c2, w3 [32]
  uint8_t* w1_s = reinterpret_cast<uint8_t*>(par + 3 * HEAD_C1 + 4 * HEAD_C2);  // [128][kFPitch]        (FUSE)
  uint8_t* f_s = w1_s + HEAD_C1 * kFPitch;                                       // 8 warps x [16][kFPitch] (FUSE)
  uint8_t* g_s = FUSE ? f_s + kHeadWarps * 16 * kFPitch : w1_s;                  // 8 warps x [14][kGPitch] (STAGED)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < HEAD_C2 * HEAD_C1; i += kHeadWarps * 32) {
    const int j = i / HEAD_C1, k = i % HEAD_C1;
    const float w = p.w2[i];
    const __nv_bfloat16 wh = __float2bfloat16_rn(w);
    *reinterpret_cast<__nv_bfloat16*>(w2h_s + j * kW2Pitch + k * 2) = wh;
    *reinterpret_cast<__nv_bfloat16*>(w2l_s + j * kW2Pitch + k * 2) = __float2bfloat16_rn(w - __bfloat162float(wh));
  }
  for (int i = tid; i < HEAD_C1; i += kHeadWarps * 32) {
    par[i] = p.b1[i];
    par[HEAD_C1 + i] = p.a1[i];
    par[2 * HEAD_C1 + i] = p.c1[i];
  }
  if (tid < HEAD_C2) {
    par[3 * HEAD_C1 + tid] = p.b2[tid];
    par[3 * HEAD_C1 + HEAD_C2 + tid] = p.a2[tid];
    par[3 * HEAD_C1 + 2 * HEAD_C2 + tid] = p.c2[tid];
    par[3 * HEAD_C1 + 3 * HEAD_C2 + tid] = p.w3[tid];
  }
  if (FUSE) {
    for (int i = tid; i < HEAD_C1 * 8; i += kHeadWarps * 32) {
      const int row = i >> 3, v = i & 7;
      *reinterpret_cast<uint4*>(w1_s + row * kFPitch + v * 16) = __ldg(reinterpret_cast<const uint4*>(full.w1 + row * 64) + v);
    }
  }
  __syncthreads();
  const float *sb1 = par, *sa1 = par + HEAD_C1, *sc1 = par + 2 * HEAD_C1;
  const float *sb2 = par + 3 * HEAD_C1, *sa2 = sb2 + HEAD_C2, *sc2 = sa2 + HEAD_C2, *sw3 = sc2 + HEAD_C2;
  const int g = lane >> 2, t4 = lane & 3, j = lane >> 3, rr = lane & 7;
  const uint32_t w2h_u = smem_u32(w2h_s), w2l_u = smem_u32(w2l_s), w1_u = smem_u32(w1_s);
  uint8_t* my_f = f_s + warp * 16 * kFPitch;
  const uint32_t f_u = smem_u32(my_f);
  const long long npix = (long long)p.n_img * p.H * p.W;
  const long long ngroups = npix / 16;
  const long long hw = (long long)p.H * p.W;

  for (long long grp = (long long)blockIdx.x * kHeadWarps + warp; grp < ngroups; grp += (long long)gridDim.x * kHeadWarps) {
    const long long pix0 = grp * 16;
    float z1[16][4];
    uint8_t* my_g = g_s + warp * kGRows * kGPitch;
    auto stage_g = [&]() {   // the group lies in one image row; level l contributes 16 >> shift consecutive source pixels
      const long long img = pix0 / hw;
      const int rem = (int)(pix0 - img * hw);
      const int y = rem / p.W, x0 = rem - y * p.W;
      int row_off = 0;
      for (int l = 0; l < p.n_levels; ++l) {
        const int sh = p.shift[l], npx = 16 >> sh;
        const float* src = reinterpret_cast<const float*>(p.g[l]) +
                           (((size_t)img * (p.H >> sh) + (y >> sh)) * (p.W >> sh) + (x0 >> sh)) * HEAD_C1;
        for (int i = lane; i < npx * 32; i += 32) {
          const int px = i >> 5, c16 = i & 31;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_g + (row_off + px) * kGPitch + c16 * 16)),
                       "l"(src + px * HEAD_C1 + c16 * 4) : "memory");
        }
        row_off += npx;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (FUSE) {
      const bf16* src = full.f + pix0 * 64;                // 16 px x 128 B, contiguous
      __syncwarp();
      for (int i = lane; i < 16 * 8; i += 32) {
        const int row = i >> 3, v = i & 7;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(f_u + row * kFPitch + v * 16), "l"(src + row * 64 + v * 8) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (STAGED) {
        stage_g();                                       // lands while the layer-1 product below runs
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) z1[nt][0] = z1[nt][1] = z1[nt][2] = z1[nt][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        h_ldsm_x4(f_u + ((j & 1) * 8 + rr) * kFPitch + (ks * 16 + (j >> 1) * 8) * 2, a);
#pragma unroll
        for (int np = 0; np < 8; ++np) {
          uint32_t bfr[4];
          h_ldsm_x4(w1_u + ((2 * np + (j >> 1)) * 8 + rr) * kFPitch + (ks * 16 + (j & 1) * 8) * 2, bfr);
          h_mma(z1[2 * np], a, bfr[0], bfr[1]);
          h_mma(z1[2 * np + 1], a, bfr[2], bfr[3]);
        }
      }
    } else {
      if (STAGED) {
        __syncwarp();
        stage_g();
      }
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) z1[nt][0] = z1[nt][1] = z1[nt][2] = z1[nt][3] = 0.0f;
    }
    // gathered maps of the other levels, at the fragment positions (rows g, g+8; columns nt*8 + 2*t4, +1)
    if (STAGED) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      int row_off = 0;
      for (int l = 0; l < p.n_levels; ++l) {
        const int sh = p.shift[l];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float* src = reinterpret_cast<const float*>(my_g + (row_off + ((g + 8 * r) >> sh)) * kGPitch) + 2 * t4;
#pragma unroll
          for (int nt = 0; nt < 16; ++nt) {
            const float2 v = *reinterpret_cast<const float2*>(src + nt * 8);
            z1[nt][2 * r] += v.x;
            z1[nt][2 * r + 1] += v.y;
          }
        }
        row_off += 16 >> sh;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long pix = pix0 + g + 8 * r;
        const long long img = pix / hw;
        const int rem = (int)(pix - img * hw);
        const int y = rem / p.W, x = rem - y * p.W;
        for (int s = 0; s < p.n_sum; ++s)
          for (int l = 0; l < p.n_levels; ++l) {
            const int sh = p.shift[l];
            const int hl = p.H >> sh, wl = p.W >> sh;
            const float* src = reinterpret_cast<const float*>(p.g[l]) +
                               ((((size_t)img * p.n_sum + s) * hl + (y >> sh)) * wl + (x >> sh)) * HEAD_C1 + 2 * t4;
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
              const float2 v = __ldg(reinterpret_cast<const float2*>(src + nt * 8));
              z1[nt][2 * r] += v.x;
              z1[nt][2 * r + 1] += v.y;
            }
          }
      }
    }
    // layer 2 on the fly: a1 fragments of k-step ks come from z1 tiles 2ks, 2ks+1
    float z2[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) z2[nt][0] = z2[nt][1] = z2[nt][2] = z2[nt][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nt = 2 * ks + half, c = nt * 8 + 2 * t4;
        const float2 b = *reinterpret_cast<const float2*>(sb1 + c);
        const float2 sa = *reinterpret_cast<const float2*>(sa1 + c);
        const float2 sc = *reinterpret_cast<const float2*>(sc1 + c);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float v0 = fmaf(fmaxf(z1[nt][2 * r] + b.x, 0.0f), sa.x, sc.x);
          const float v1 = fmaf(fmaxf(z1[nt][2 * r + 1] + b.y, 0.0f), sa.y, sc.y);
          split_pack(v0, v1, ah[half * 2 + r], al[half * 2 + r]);
        }
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bh[4], bl[4];
        const uint32_t off = ((2 * np + (j >> 1)) * 8 + rr) * kW2Pitch + (ks * 16 + (j & 1) * 8) * 2;
        h_ldsm_x4(w2h_u + off, bh);
        h_ldsm_x4(w2l_u + off, bl);
        h_mma(z2[2 * np], ah, bh[0], bh[1]);
        h_mma(z2[2 * np], al, bh[0], bh[1]);
        h_mma(z2[2 * np], ah, bl[0], bl[1]);
        h_mma(z2[2 * np + 1], ah, bh[2], bh[3]);
        h_mma(z2[2 * np + 1], al, bh[2], bh[3]);
        h_mma(z2[2 * np + 1], ah, bl[2], bl[3]);
      }
    }
    float part[2] = {0.0f, 0.0f};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = nt * 8 + 2 * t4;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float b2v = sb2[c + e], a2v = sa2[c + e], c2v = sc2[c + e], w3v = sw3[c + e];
#pragma unroll
        for (int r = 0; r < 2; ++r) part[r] = fmaf(w3v, fmaf(fmaxf(z2[nt][2 * r + e] + b2v, 0.0f), a2v, c2v), part[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float v = part[r];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t4 == 0) p.logits[pix0 + g + 8 * r] = v + p.b3;
    }
  }
}
}  // namespace

extern "C" int tedm_head_infer(const tedm_head_args* a, tedm_stream_t stream) {
  TEDM_CHECK_ARG(a && a->logits && a->b1 && a->bn1_a && a->bn1_c && a->w2 && a->b2 && a->bn2_a && a->bn2_c && a->w3,
                 "tedm_head_infer: null pointer");
  TEDM_CHECK_ARG(a->n_levels >= (a->f_full ? 0 : 1) && a->n_levels <= 4 && a->n_sum >= 1 && a->n_img > 0 && a->height > 0 && a->width > 0,
                 "tedm_head_infer: bad sizes");
  TEDM_UNSUPPORTED(a->c1 != HEAD_C1 || a->c2 != HEAD_C2, "tedm_head_infer: head widths %d/%d (only 128/32)", a->c1, a->c2);
  HeadParams p{};
  for (int l = 0; l < a->n_levels; ++l) {
    TEDM_CHECK_ARG(a->g[l] != nullptr && a->shift[l] >= 0 && (a->height >> a->shift[l]) > 0 &&
                       ((a->height >> a->shift[l]) << a->shift[l]) == a->height &&
                       ((a->width >> a->shift[l]) << a->shift[l]) == a->width,
                   "tedm_head_infer: level %d shift %d does not divide %dx%d", l, a->shift[l], a->height, a->width);
    p.g[l] = a->g[l];
    p.shift[l] = a->shift[l];
  }
  p.n_levels = a->n_levels;
  p.n_sum = a->n_sum;
  p.n_img = a->n_img;
  p.H = a->height;
  p.W = a->width;
  p.b1 = a->b1; p.a1 = a->bn1_a; p.c1 = a->bn1_c; p.w2 = a->w2; p.b2 = a->b2; p.a2 = a->bn2_a; p.c2 = a->bn2_c; p.w3 = a->w3;
  p.b3 = a->b3;
  p.logits = a->logits;
  const long long npix = (long long)a->n_img * a->height * a->width;
  TEDM_CHECK_ARG(a->g_dtype == 0 || a->g_dtype == 1, "tedm_head_infer: g_dtype=%d", a->g_dtype);
  TEDM_CHECK_ARG(!a->exact || (a->g_dtype == 1 && a->f_full == nullptr), "tedm_head_infer: exact mode takes fp32 g maps for every level");
  if (a->g_dtype == 1 && npix % 16 == 0 && !a->exact) {   // tensor-core tail
    const bool fuse = a->f_full != nullptr;
    TEDM_CHECK_ARG(!fuse || (a->w1_full && a->c_full == 64 && a->n_sum == 1),
                   "tedm_head_infer: the fused full-resolution level needs its weight slice, 64 channels and n_sum == 1");
    HeadFull full{(const bf16*)a->f_full, (const bf16*)a->w1_full};
    bool staged = a->n_sum == 1 && a->width % 16 == 0 && a->n_levels >= 1;
    int rows = 0;
    for (int l = 0; l < a->n_levels; ++l) {
      staged = staged && a->shift[l] >= 1 && a->shift[l] <= 3;
      rows += 16 >> a->shift[l];
    }
    staged = staged && rows <= kGRows;
    const int smem = 2 * HEAD_C2 * kW2Pitch + (3 * HEAD_C1 + 4 * HEAD_C2) * 4 +
                     (fuse ? HEAD_C1 * kFPitch + kHeadWarps * 16 * kFPitch : 0) + (staged ? kHeadWarps * kGRows * kGPitch : 0);
    static bool configured = false;
    if (!configured) {
      TEDM_CUDA(cudaFuncSetAttribute(head_tail_mma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      TEDM_CUDA(cudaFuncSetAttribute(head_tail_mma_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      TEDM_CUDA(cudaFuncSetAttribute(head_tail_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      TEDM_CUDA(cudaFuncSetAttribute(head_tail_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      TEDM_CUDA(cudaFuncSetAttribute(head_tail_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
      configured = true;
    }
    long long nb = (npix / 16 + kHeadWarps - 1) / kHeadWarps;
    const void* kfn = fuse ? (staged ? (const void*)head_tail_mma_kernel<true, true> : (const void*)head_tail_mma_kernel<true, false>)
                           : (staged ? (const void*)head_tail_mma_kernel<false, true> : (const void*)head_tail_mma_kernel<false, false>);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kHeadWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long capm = (long long)per_sm * tedm_num_sms();
    if (nb > capm) nb = capm;
    void* kargs[] = {(void*)&p, (void*)&full};
    TEDM_CUDA(cudaLaunchKernel(kfn, dim3((unsigned)nb), dim3(kHeadWarps * 32), kargs, smem, (cudaStream_t)stream));
    TEDM_LAUNCH_CHECK();
    return TEDM_OK;
  }
  TEDM_CHECK_ARG(a->f_full == nullptr, "tedm_head_infer: the fused full-resolution level needs fp32 g maps and n_img*H*W %% 16 == 0");
  long long blocks = (npix + 127) / 128;
  const long long cap = a->g_dtype == 1 ? resident_ctas(head_tail_kernel<true>, 128, 0) : resident_ctas(head_tail_kernel<false>, 128, 0);
  if (blocks > cap) blocks = cap;
  if (a->g_dtype == 1) head_tail_kernel<true><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(p);
  else head_tail_kernel<false><<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}

// prob[b] = mean_s sigmoid(logits[b*S + s]) ; mask = prob > 0.5
// (auxiliary/postprocessing/testing_shared_weights.py:113,120,133-138; app.py:79)
__global__ void __launch_bounds__(256) ensemble_kernel(const float* __restrict__ logits, float* __restrict__ prob,
                                                       uint8_t* __restrict__ mask, int n_steps, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw;
    const int pi = (int)(i % hw);
    float acc = 0.0f;
    for (int s = 0; s < n_steps; ++s) {
      const float v = logits[((size_t)b * n_steps + s) * hw + pi];
      acc += 1.0f / (1.0f + expf(-v));
    }
    const float pr = acc / (float)n_steps;
    if (prob) prob[i] = pr;
    if (mask) mask[i] = pr > 0.5f ? 1 : 0;
  }
}

extern "C" int tedm_ensemble_mask(const float* logits, float* prob, uint8_t* mask, int batch, int n_steps, int hw,
                                  tedm_stream_t stream) {
  TEDM_CHECK_ARG(logits && (prob || mask) && batch > 0 && n_steps > 0 && hw > 0, "tedm_ensemble_mask: bad arguments");
  const long long total = (long long)batch * hw;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)tedm_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  ensemble_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, prob, mask, n_steps, hw, total);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
