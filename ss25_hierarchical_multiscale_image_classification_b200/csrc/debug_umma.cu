// Test hook: does a UMMA shared-memory descriptor whose start address is shifted by whole rows
// (not a multiple of the 8-row swizzle atom) address a TMA-written swizzled tile correctly?
// The "row-tile" convolution kernels rely on it: all nine 3x3 taps read ONE staged activation region
// through descriptors that differ only in their start address.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace hipac {

// A: [256][kcols] bf16 (kcols = 64 for 128B swizzle, 16 for 32B swizzle), B: [64][kcols] bf16,
// D[m][n] = sum_k A[m + shift][k] * B[n][k], m < 128, n < 64.
template <int SWZ>
__global__ void __launch_bounds__(128, 1)
k_debug_umma_shift(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int shift) {
  constexpr int ROWB = SWZ;                    // bytes per row
  constexpr int KSTEPS = SWZ == 128 ? 4 : 1;   // UMMA K = 16 bf16 = 32 bytes
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                  // 256 rows
  uint8_t* sB = base + 256 * ROWB;     // 64 rows (1024-aligned for both swizzles)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * ROWB);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, (256 + 64) * ROWB);
    ptx::tma_load_2d(sA, &tmA, bar, 0, 0);
    ptx::tma_load_2d(sB, &tmB, bar, 0, 0);
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    const uint32_t a_addr = ptx::smem_u32(sA) + shift * ROWB, b_addr = ptx::smem_u32(sB);
    for (int k = 0; k < KSTEPS; k++)
      ptx::umma_bf16(tmem, ptx::make_smem_desc(a_addr + k * 32, SWZ), ptx::make_smem_desc(b_addr + k * 32, SWZ),
                     ptx::make_idesc_bf16(128, 64), k != 0);
    ptx::umma_commit(mma_bar);
  }
  ptx::mbar_wait(mma_bar, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; i++) D[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(v[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace hipac

using namespace hipac;

extern "C" int hipac_debug_umma_shift(const void* d_A, const void* d_B, float* d_D, int shift_rows, int swizzle_bytes,
                                      void* stream_) {
  HIPAC_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 32, "swizzle must be 128 or 32");
  HIPAC_REQUIRE(shift_rows >= 0 && shift_rows <= 128, "shift out of range");
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  HIPAC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  HIPAC_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "driver lacks cuTensorMapEncodeTiled");
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int kcols = swizzle_bytes / 2;
  CUtensorMap tmA, tmB;
  auto mk = [&](CUtensorMap* m, const void* p, int rows) {
    cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kcols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kcols, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  HIPAC_REQUIRE(mk(&tmA, d_A, 256) == CUDA_SUCCESS && mk(&tmB, d_B, 64) == CUDA_SUCCESS, "tensor map encode failed");
  const int smem = (256 + 64) * swizzle_bytes + 1024 + 64;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (swizzle_bytes == 128) {
    HIPAC_CHECK_CUDA(cudaFuncSetAttribute(k_debug_umma_shift<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_debug_umma_shift<128><<<1, 128, smem, stream>>>(tmA, tmB, d_D, shift_rows);
  } else {
    HIPAC_CHECK_CUDA(cudaFuncSetAttribute(k_debug_umma_shift<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_debug_umma_shift<32><<<1, 128, smem, stream>>>(tmA, tmB, d_D, shift_rows);
  }
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
