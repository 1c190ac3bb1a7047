// One-launch re-layout of EVERY implicit-GEMM conv weight of the net: fp32 OIHW parameters (the
// state_dict layout of models/unet_model.py, untouched) -> the bf16 operand layouts the tcgen05
// kernels read, for the forward conv and (training) for the data-gradient conv.  A training step
// changes all 36 M weights, so this runs once per step; tiles go through shared memory so that
// both the fp32 reads and the bf16 writes are contiguous runs.
#include "common.cuh"

namespace {

constexpr int WT = 32;   // tile: 32 output channels x 32 input channels x all taps

__device__ __forceinline__ int fold_lo(int par, int a) { return par == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2); }
__device__ __forceinline__ int fold_hi(int par, int a) { return par == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2); }

// One 32 x 32 x taps tile.  Index arithmetic is hoisted: a warp owns whole (row, tap) runs and its lanes walk the
// contiguous inner index, so no thread divides per element.
template <int MODE>
__device__ __forceinline__ void relayout_tile(const tedm_weight_entry& e, int co0, int ci0, float* sw) {
  constexpr int khw = MODE == 0 ? 1 : (MODE == 2 ? 16 : 9);   // taps of the stored parameter
  constexpr int pitch = WT * khw + 1;
  const int cout = e.cout, cin = e.cin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* w = (const float*)e.w;
  for (int co_l = warp; co_l < WT; co_l += 8) {
    const float* src = w + ((size_t)(co0 + co_l) * cin + ci0) * khw;
    for (int col = lane; col < WT * khw; col += 32) sw[co_l * pitch + col] = src[col];
  }
  __syncthreads();
  auto at = [&](int co_l, int ci_l, int tap) { return sw[co_l * pitch + ci_l * khw + tap]; };
  bf16* fwd = (bf16*)e.fwd;
  bf16* dg = (bf16*)e.dgrad;
  // ---- forward operand (lane = input channel)
  if (fwd) {
    if (MODE != 3) {            // KRSC [co][tap][ci]
      for (int r = warp; r < WT * khw; r += 8) {
        const int co_l = r / khw, tap = r - co_l * khw;
        fwd[((size_t)(co0 + co_l) * khw + tap) * cin + ci0 + lane] = __float2bfloat16_rn(at(co_l, lane, tap));
      }
    } else {                    // folded [par][co][a][b][ci]
      for (int r = warp; r < 4 * WT * 4; r += 8) {
        const int ab = r & 3, co_l = (r >> 2) & (WT - 1), par = r >> 7;
        const int a = ab >> 1, b = ab & 1, py = par >> 1, px = par & 1;
        float v = 0.0f;
        for (int y3 = fold_lo(py, a); y3 <= fold_hi(py, a); ++y3)
          for (int x3 = fold_lo(px, b); x3 <= fold_hi(px, b); ++x3) v += at(co_l, lane, y3 * 3 + x3);
        fwd[(((size_t)par * cout + co0 + co_l) * 4 + ab) * cin + ci0 + lane] = __float2bfloat16_rn(v);
      }
    }
  }
  // ---- data-gradient operand (see tedm_weight_to_dgrad; lane = output channel)
  if (dg) {
    if (MODE == 0 || MODE == 1) {       // [ci][tap'][co] = w[co][ci][khw-1-tap']
      for (int r = warp; r < WT * khw; r += 8) {
        const int ci_l = r / khw, tap = r - ci_l * khw;
        dg[((size_t)(ci0 + ci_l) * khw + tap) * cout + co0 + lane] = __float2bfloat16_rn(at(lane, ci_l, khw - 1 - tap));
      }
    } else if (MODE == 2) {             // [par][ci][a][b][co] = w[co][ci][3-2a-py][3-2b-px]
      for (int r = warp; r < 4 * WT * 4; r += 8) {
        const int ab = r & 3, ci_l = (r >> 2) & (WT - 1), par = r >> 7;
        const int a = ab >> 1, b = ab & 1, py = par >> 1, px = par & 1;
        dg[(((size_t)par * cin + ci0 + ci_l) * 4 + ab) * cout + co0 + lane] =
            __float2bfloat16_rn(at(lane, ci_l, (3 - 2 * a - py) * 4 + (3 - 2 * b - px)));
      }
    } else {                            // [ci][ky][kx][co] (4x4) from the 3x3 parameter
      for (int r = warp; r < WT * 16; r += 8) {
        const int kx = r & 3, ky = (r >> 2) & 3, ci_l = r >> 4;
        const int py = (ky + 1) & 1, px = (kx + 1) & 1, a = (3 - ky - py) >> 1, b = (3 - kx - px) >> 1;
        float v = 0.0f;
        for (int y3 = fold_lo(py, a); y3 <= fold_hi(py, a); ++y3)
          for (int x3 = fold_lo(px, b); x3 <= fold_hi(px, b); ++x3) v += at(lane, ci_l, y3 * 3 + x3);
        dg[(((size_t)(ci0 + ci_l) * 4 + ky) * 4 + kx) * cout + co0 + lane] = __float2bfloat16_rn(v);
      }
    }
  }
}

__global__ void __launch_bounds__(256) prepare_weights_kernel(const tedm_weight_entry* __restrict__ table, int n_entries) {
  extern __shared__ float sw[];   // [WT co][WT*khw + 1]
  // entry lookup: last entry whose cta_begin <= blockIdx.x
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].cta_begin <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const tedm_weight_entry e = table[lo];
  const int local = blockIdx.x - e.cta_begin;
  const int ci_tiles = e.cin / WT;
  const int co0 = (local / ci_tiles) * WT, ci0 = (local % ci_tiles) * WT;
  switch (e.mode) {
    case 0: relayout_tile<0>(e, co0, ci0, sw); break;
    case 1: relayout_tile<1>(e, co0, ci0, sw); break;
    case 2: relayout_tile<2>(e, co0, ci0, sw); break;
    default: relayout_tile<3>(e, co0, ci0, sw); break;
  }
}

}  // namespace

extern "C" int tedm_prepare_weights(const tedm_weight_entry* table_dev, int n_entries, int total_ctas, tedm_stream_t stream) {
  TEDM_CHECK_ARG(table_dev && n_entries > 0 && total_ctas > 0, "tedm_prepare_weights: bad arguments");
  const int smem = WT * (WT * 16 + 1) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    TEDM_CUDA(cudaFuncSetAttribute(prepare_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  prepare_weights_kernel<<<total_ctas, 256, smem, (cudaStream_t)stream>>>(table_dev, n_entries);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
