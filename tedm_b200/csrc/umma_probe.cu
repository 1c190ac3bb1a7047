// Hardware probe (test-only entry point): does a K-major SWIZZLE_128B UMMA descriptor whose start
// address is shifted by a whole number of 128-byte rows (not a multiple of the 1024-byte swizzle
// atom) read rows [shift, shift+128) of a TMA-written buffer correctly?  The answer decides whether
// the 3x3 conv can reuse ONE halo tile in shared memory for all nine taps instead of re-fetching
// the A operand from L2 per tap.  One CTA; variants run back to back.
#include "common.cuh"

namespace {
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int PROBE_ROWS = 384;
constexpr int PROBE_MAXVAR = 32;
struct ProbeParams {
  int nvar;
  int shift[PROBE_MAXVAR];
  int base_off[PROBE_MAXVAR];
  float* out;
};

__global__ void __launch_bounds__(192, 1) umma_probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB, const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_acc;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + PROBE_ROWS * 128;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_load), 1);
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar_load), PROBE_ROWS * 128 + 64 * 128);
    for (int i = 0; i < PROBE_ROWS / 128; ++i) tma_load_2d(a_addr + i * 128 * 128, &mapA, smem_u32(&bar_load), 0, i * 128);
    tma_load_2d(b_addr, &mapB, smem_u32(&bar_load), 0, 0);
  }
  mbar_wait(smem_u32(&bar_load), 0);
  tc_fence_after();
  for (int v = 0; v < p.nvar; ++v) {
    if (warp == 1) {
      if (elect_one()) {
        const uint32_t sa = a_addr + (uint32_t)p.shift[v] * 128u;
        const uint64_t adesc = (uint64_t)((sa & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
                               ((uint64_t)(p.base_off[v] & 7) << 49) | (2ull << 61);
        const uint64_t bdesc = (uint64_t)((b_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + 2ull * k, bdesc + 2ull * k, IDESC, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&bar_acc));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar_acc), (uint32_t)(v & 1));
    tc_fence_after();
    if (warp >= 2) {
      const int q = warp & 3, row = q * 32 + lane;
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chunk * 32), r);
        tmem_ld_wait();
        float* dst = p.out + ((size_t)v * 128 + row) * 64 + chunk * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 1) tmem_dealloc<64>(tmem_base);
}
}  // namespace

// A: device [384][64] bf16, Bm: device [64][64] bf16, out: device [nvar][128][64] fp32 with
// out[v][m][n] = sum_k A[shift_v + m][k] * Bm[n][k] when the descriptor variant works.
extern "C" int tedm_debug_umma_probe(const void* A, const void* Bm, const int* shifts, const int* base_offsets, int nvar,
                                     float* out, tedm_stream_t stream) {
  TEDM_CHECK_ARG(A && Bm && shifts && base_offsets && out && nvar > 0 && nvar <= PROBE_MAXVAR, "tedm_debug_umma_probe: bad arguments");
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  TEDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  TEDM_CHECK_ARG(fp && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled entry point not found");
  PFN_tensorMapEncodeTiled enc = (PFN_tensorMapEncodeTiled)fp;
  alignas(64) CUtensorMap mapA, mapB;
  cuuint64_t dimsA[2] = {64, PROBE_ROWS}, dimsB[2] = {64, 64}, strides[1] = {128};
  cuuint32_t boxA[2] = {64, 128}, boxB[2] = {64, 64}, estr[2] = {1, 1};
  CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dimsA, strides, boxA, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TEDM_CHECK_ARG(r == CUDA_SUCCESS, "probe: encode A failed: %d", (int)r);
  r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Bm), dimsB, strides, boxB, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TEDM_CHECK_ARG(r == CUDA_SUCCESS, "probe: encode B failed: %d", (int)r);
  ProbeParams p{};
  p.nvar = nvar;
  for (int i = 0; i < nvar; ++i) {
    TEDM_CHECK_ARG(shifts[i] >= 0 && shifts[i] + 128 <= PROBE_ROWS, "probe: shift %d out of range", shifts[i]);
    p.shift[i] = shifts[i];
    p.base_off[i] = base_offsets[i];
  }
  p.out = out;
  const int smem = PROBE_ROWS * 128 + 64 * 128 + 1024;
  TEDM_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 192, smem, (cudaStream_t)stream>>>(mapA, mapB, p);
  TEDM_LAUNCH_CHECK();
  return TEDM_OK;
}
