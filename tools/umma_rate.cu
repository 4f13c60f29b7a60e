// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of the operand swizzle mode
// (128-byte rows / 32-byte rows) and of a row-shifted descriptor start, one CTA per SM, operands in (zeroed) shared memory.
// Prints clocks per UMMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ss25_hierarchical_multiscale_image_classification_b200/csrc/umma.cuh"
using namespace hipac;

// mode: 0 = SW128 aligned, k-steps inside the 128-byte row; 1 = SW128 start shifted by `shift` rows per tap;
//       2 = SW32 aligned; 3 = SW32 shifted by `shift` 32-byte rows per tap
template <int N>
__global__ void __launch_bounds__(64, 1) k_rate(int mode, int shift, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (threadIdx.x == 0) ptx::mbar_init(&bar, 1), ptx::fence_barrier_init();
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512), ptx::tmem_relinquish();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t a0 = ptx::smem_u32(base), b0 = ptx::smem_u32(base + 96 * 1024);
    const bool sw32 = mode >= 2;
    const uint64_t adesc = ptx::make_smem_desc(a0, sw32 ? 32 : 128), bdesc = ptx::make_smem_desc(b0, 128);
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const int row_units = sw32 ? 2 : 8;   // descriptor units (16 B) per row
    long long t0 = 0, t1 = 0;
    uint64_t ad[16], bd[16];   // descriptors are loop invariants: the timed loop is UMMA issue only
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (mode == 0) ad[k] = adesc + 2 * (k & 3);
      else if (mode == 1) ad[k] = adesc + (uint64_t)((k >> 2) * shift * row_units) + 2 * (k & 3);
      else if (mode == 2) ad[k] = adesc;
      else ad[k] = adesc + (uint64_t)(((k >> 2) * shift + (k & 3)) * row_units);
      bd[k] = bdesc + 2 * (k & 3);
    }
    for (int rep = 0; rep < 2; rep++) {
      t0 = clock64();
      if (ptx::elect_one()) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
          for (int k = 0; k < 16; k++) ptx::umma_bf16(tm + (i & 1) * N, ad[k], bd[k], idesc, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&bar);
      }
      __syncwarp();
      ptx::mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

template <int N>
static void run(int mode, int shift, long long* d_out) {
  const int iters = 2000;
  cudaFuncSetAttribute(k_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_rate<N><<<148, 64, 200 * 1024>>>(mode, shift, iters, d_out);
  long long clk = 0;
  cudaError_t e = cudaMemcpy(&clk, d_out, sizeof(clk), cudaMemcpyDeviceToHost);
  static const char* names[] = {"SW128 aligned", "SW128 row-shifted", "SW32 aligned", "SW32 row-shifted"};
  printf("{\"N\": %d, \"mode\": \"%s\", \"shift_rows\": %d, \"clk_per_umma\": %.1f, \"ideal_clk\": %d, \"err\": \"%s\"}\n", N, names[mode], shift,
         (double)clk / (iters * 16.0), N / 2, cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  for (int mode = 0; mode < 4; mode++) {
    const int shift = mode == 1 ? 58 : (mode == 3 ? 115 : 0);
    run<64>(mode, shift, d_out);
    run<128>(mode, shift, d_out);
    run<256>(mode, shift, d_out);
  }
  // cost of the start offset modulo the 8-row swizzle atom (every UMMA of the loop uses a start of k/4 * shift rows)
  for (int shift : {1, 2, 3, 4, 6, 8, 16, 64}) run<64>(1, shift, d_out);
  for (int shift : {1, 2, 4, 8, 120}) run<64>(3, shift, d_out);
  return 0;
}
