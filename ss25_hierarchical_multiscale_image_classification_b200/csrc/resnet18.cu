// Stage 2 of the HiPAC hot path: ResNet18 forward (reference src/models/resnet.py:22-77, i.e. the
// torchvision resnet18 trunk in eval mode + global average pool [+ Linear(512,k)]) as implicit-GEMM
// tcgen05/TMEM kernels.  Activations are bf16 NHWC; eval-mode BatchNorm is folded into the weights
// (fp32 fold, bf16 round) and a per-channel fp32 bias added in the epilogue together with the
// residual and ReLU.  The A operand (im2col of the activations) is produced by TMA im2col loads,
// the B operand (weights, K-major) by tiled TMA loads; accumulators live in TMEM, double buffered
// so the epilogue of tile i overlaps the MMAs of tile i+1.  See DESIGN.md.
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "resnet18_layers.h"
#include "umma.cuh"

namespace hipac {

// ==========================================================================================
// implicit-GEMM convolution kernel
// ==========================================================================================
struct ConvParams {
  int M_total;       // images * hout * wout (GEMM M)
  int hw_out, wout;  // hout*wout, wout
  int stride, pad_w, pad_h;
  int kw;            // filter width (taps per filter row)
  int kc_blocks;     // 64-channel blocks per tap (cin / 64)
  int num_kb;        // k-blocks per tile from the main operand
  int num_kb2;       // extra k-blocks from the second operand (fused 1x1 projection shortcut), 0 if none
  int stride2;       // traversal stride of the second operand (pad 0, 1x1)
  int num_m_tiles, num_n_tiles;
  int cout;
  int relu;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  const int* n_dev;  // device-side patch count (null: the host-side sizes above are exact)
  int n_base;        // first patch of this chunk: the kernel works on clamp(*n_dev - n_base, 0, host n) patches
  int reverse;       // walk the tiles from the last to the first (see g_reverse)
};

// Device-count mode: the host sizes grids, tensor maps and buffers for a CAPACITY; the number of patches that actually
// exist is read from device memory when the kernel starts, so a whole step can be enqueued without a host round trip.
__device__ __forceinline__ int effective_patches(const int* n_dev, int n_base, int n_cap) {
  if (!n_dev) return n_cap;
  const int n = __ldg(n_dev) - n_base;
  return n < 0 ? 0 : (n < n_cap ? n : n_cap);
}
static thread_local const int* g_n_dev = nullptr;   // set by hipac_resnet18_forward_dcount around its launches
static thread_local int g_n_base = 0;
// Boustrophedon tile order across the layers of one forward pass: a layer whose producer wrote images 0..n-1 walks
// them n-1..0, so its first ~100 MB of reads are the producer's LAST writes and still sit in the 126 MB L2 (with the
// same order on both sides they would be the oldest, long evicted).  Flipped after every launch of a forward pass.
static thread_local int g_reverse = 0;

constexpr int kBM = 128;
constexpr int kS2dW = HIPAC_S2D16_WIDTH;  // 112 + 3 explicit zero columns (2 left, 1 right)
constexpr int kABytes = kBM * 128;  // 128 rows x 64 bf16
constexpr int kConvThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
// 64-wide tiles have so little MMA work per tile that one epilogue warpgroup cannot keep up: they get two, which
// take alternate tiles (= alternate TMEM accumulator buffers).  Wider tiles keep one (register budget).
__host__ __device__ constexpr int epi_groups(int bn) { return bn == 64 ? 2 : 1; }
__host__ __device__ constexpr int conv_threads(int bn) { return 64 + 128 * epi_groups(bn); }

// 32 consecutive output channels of one output pixel: + folded-BN bias (+ residual, already in registers)
// (+ ReLU) -> bf16, 64-byte store.
__device__ __forceinline__ void epilogue_math32(const uint32_t (&v)[32], const float* __restrict__ bias, const uint4* res,
                                                __nv_bfloat16* __restrict__ out, int relu) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  uint4 o[4];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float4 b = __ldg(b4 + i);
    float x0 = __uint_as_float(v[4 * i + 0]) + b.x, x1 = __uint_as_float(v[4 * i + 1]) + b.y;
    float x2 = __uint_as_float(v[4 * i + 2]) + b.z, x3 = __uint_as_float(v[4 * i + 3]) + b.w;
    if (res) {
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[i >> 1]) + (i & 1) * 2;
      const __nv_bfloat162 ra = *reinterpret_cast<const __nv_bfloat162*>(&rw[0]);
      const __nv_bfloat162 rb = *reinterpret_cast<const __nv_bfloat162*>(&rw[1]);
      x0 += __bfloat162float(ra.x), x1 += __bfloat162float(ra.y);
      x2 += __bfloat162float(rb.x), x3 += __bfloat162float(rb.y);
    }
    if (relu) x0 = fmaxf(x0, 0.f), x1 = fmaxf(x1, 0.f), x2 = fmaxf(x2, 0.f), x3 = fmaxf(x3, 0.f);
    __nv_bfloat162 lo = __floats2bfloat162_rn(x0, x1), hi = __floats2bfloat162_rn(x2, x3);
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o[i >> 1]) + (i & 1) * 2;
    ow[0] = *reinterpret_cast<uint32_t*>(&lo);
    ow[1] = *reinterpret_cast<uint32_t*>(&hi);
  }
  uint4* dst = reinterpret_cast<uint4*>(out);
#pragma unroll
  for (int i = 0; i < 4; i++) dst[i] = o[i];
}

// Epilogue of one accumulator row (BN channels of one output pixel).  The residual row is fetched into
// registers BEFORE waiting for the accumulator (its address does not depend on the MMAs), so its global-load
// latency hides behind the tile's main loop; TMEM is drained 64 columns per wait.
template <int BN>
__device__ __forceinline__ void epilogue_row(uint32_t tmem_row, const float* __restrict__ bias,
                                             const __nv_bfloat16* __restrict__ residual_row, __nv_bfloat16* __restrict__ out_row,
                                             int relu, bool valid, uint64_t* tfull_bar, uint32_t phase) {
  constexpr bool kPrefetch = BN <= 128;  // 16 B per 8 channels: 32 (BN=64) / 64 (BN=128) registers
  uint4 res[kPrefetch ? BN / 8 : 8];
  const bool has_res = residual_row != nullptr && valid;
  if (kPrefetch && has_res) {
#pragma unroll
    for (int i = 0; i < BN / 8; i++) res[i] = __ldg(reinterpret_cast<const uint4*>(residual_row) + i);
  }
  ptx::mbar_wait(tfull_bar, phase);
  ptx::tc_fence_after();
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 64) {
    uint32_t v0[32], v1[32];
    ptx::tmem_ld_32x32b_x32(tmem_row + c0, v0);
    ptx::tmem_ld_32x32b_x32(tmem_row + c0 + 32, v1);
    if (!kPrefetch && has_res) {
#pragma unroll
      for (int i = 0; i < 8; i++) res[i] = __ldg(reinterpret_cast<const uint4*>(residual_row + c0) + i);
    }
    ptx::tmem_ld_wait();
    if (valid) {
      const uint4* r0 = has_res ? &res[kPrefetch ? c0 / 8 : 0] : nullptr;
      epilogue_math32(v0, bias + c0, r0, out_row + c0, relu);
      epilogue_math32(v1, bias + c0 + 32, has_res ? r0 + 4 : nullptr, out_row + c0 + 32, relu);
    }
  }
}

// RESB: the whole weight matrix (<= kResBlocks k-blocks of [BN x 64]) stays resident in shared memory for the lifetime of the
// persistent CTA and the ring carries the im2col A tiles only.  Used where it fits (layer2.0.conv1: 9 x 16 KB): the
// streamed form re-reads 16 KB of weights from L2 per 16 KB of activations, which makes that layer L2-bandwidth bound.
template <int BN, bool RESB = false>
struct ConvCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kResBlocks = RESB ? 9 : 0;
  static constexpr int kStage = RESB ? kABytes : kABytes + kBBytes;
  static constexpr int kStages = RESB ? 5 : (BN == 256 ? 4 : 6);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStage + kResBlocks * kBBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BN, bool RESB = false>
__global__ void __launch_bounds__(conv_threads(BN), 1)
k_conv_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmA2, const ConvParams p) {
  using Cfg = ConvCfg<BN, RESB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* resB = base + Cfg::kStages * Cfg::kStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(resB + Cfg::kResBlocks * Cfg::kBBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* b_full = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (p.num_kb2) ptx::prefetch_tensormap(&tmA2);
    for (int s = 0; s < Cfg::kStages; s++) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    for (int a = 0; a < 2; a++) ptx::mbar_init(&tfull[a], 1), ptx::mbar_init(&tempty[a], 4);
    ptx::mbar_init(b_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (RESB && warp == 0) {   // resident weights: constant data, loaded before the dependency wait
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(b_full, (uint32_t)(p.num_kb * Cfg::kBBytes));
      for (int kb = 0; kb < p.num_kb; kb++) ptx::tma_load_2d(resB + kb * Cfg::kBBytes, &tmB, b_full, kb * 64, 0);
    }
    __syncwarp();
  }
  ptx::pdl_wait();           // the producer kernel's activations (and the device-side patch count) are visible from here on
  const int M_total = effective_patches(p.n_dev, p.n_base, p.M_total / p.hw_out) * p.hw_out;
  const int num_tiles = ((M_total + kBM - 1) / kBM) * p.num_n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop with warp-uniform values (descriptors and coordinates then live in uniform
    // registers); one elected lane issues the TMA instructions.
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int vt = p.reverse ? num_tiles - 1 - tile : tile;
        const int m_tile = vt / p.num_n_tiles, n_tile = vt - m_tile * p.num_n_tiles;
        const int m0 = m_tile * kBM;
        const int img = m0 / p.hw_out, rem = m0 - img * p.hw_out;
        const int p0 = rem / p.wout, q0 = rem - p0 * p.wout;
        const int cw = q0 * p.stride - p.pad_w, ch = p0 * p.stride - p.pad_h;
        for (int kb = 0; kb < p.num_kb + p.num_kb2; kb++) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_dst = base + stage * Cfg::kStage;
          uint8_t* b_dst = a_dst + kABytes;
          const int tap = kb / p.kc_blocks, kc = kb - tap * p.kc_blocks;
          const int r = tap / p.kw, s = tap - r * p.kw;
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&full[stage], Cfg::kStage);
            if (kb < p.num_kb)
              ptx::tma_load_im2col_4d(a_dst, &tmA, &full[stage], kc * 64, cw, ch, img, (uint16_t)s, (uint16_t)r);
            else  // fused projection shortcut: 1x1 / stride2 / pad 0 over the block input, channels (kb - num_kb) * 64
              ptx::tma_load_im2col_4d(a_dst, &tmA2, &full[stage], (kb - p.num_kb) * 64, q0 * p.stride2, p0 * p.stride2, img, 0, 0);
            if (!RESB) ptx::tma_load_2d(b_dst, &tmB, &full[stage], kb * 64, n_tile * BN);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-uniform loop; one elected lane (always the same one) issues the UMMAs and their commits.
    {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      if (RESB) ptx::mbar_wait(b_full, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb + p.num_kb2; kb++) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(base + stage * Cfg::kStage);
          const uint64_t adesc = ptx::make_smem_desc(a_addr, 128);
          const uint64_t bdesc = ptx::make_smem_desc(RESB ? ptx::smem_u32(resB + kb * Cfg::kBBytes) : a_addr + kABytes, 128);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; k++)  // 4 x (K = 16) per 64-wide k-block: +32 B = +2 in the descriptor's address field
              ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::umma_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == Cfg::kStages) stage = 0, phase ^= 1;
        }
        if (ptx::elect_one()) ptx::umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: TMEM -> +bias (+residual) -> ReLU -> bf16 NHWC =====================
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;
    const int grp = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (epi_groups(BN) == 2 && (it & 1) != grp) continue;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int vt = p.reverse ? num_tiles - 1 - tile : tile;
      const int m_tile = vt / p.num_n_tiles, n_tile = vt - m_tile * p.num_n_tiles;
      const int m = m_tile * kBM + row;
      const bool valid = m < M_total;
      {
        const size_t off = (size_t)m * p.cout + (size_t)n_tile * BN;
        epilogue_row<BN>(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BN, p.bias + n_tile * BN,
                         p.residual ? p.residual + off : nullptr, p.out + off, p.relu, valid, &tfull[acc], acc_phase);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

#include "conv_umma2.cuh"
#include "conv_rows.cuh"
#include "conv_rows_tma.cuh"
#include "conv_rows2.cuh"
#include "conv_rows2_s2.cuh"
#include "conv_stem.cuh"

// ==========================================================================================
// small memory-bound kernels
// ==========================================================================================
// MaxPool 3x3 / stride 2 / pad 1 on bf16 NHWC [n,112,112,64] -> [n,56,56,64]; 8 channels per thread.
__global__ void __launch_bounds__(256) k_maxpool(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                 int n_img) {
  const int64_t total = (int64_t)n_img * 56 * 56 * 8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t & 7);
    int64_t pix = t >> 3;
    const int ox = (int)(pix % 56);
    pix /= 56;
    const int oy = (int)(pix % 56);
    const int img = (int)(pix / 56);
    __nv_bfloat162 m[4];
    bool first = true;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const int iy = 2 * oy + dy;
      if (iy < 0 || iy >= 112) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        const int ix = 2 * ox + dx;
        if (ix < 0 || ix >= 112) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((int64_t)img * 112 + iy) * 112 + ix) * 64 + cg * 8));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
        if (first) {
          m[0] = h[0], m[1] = h[1], m[2] = h[2], m[3] = h[3];
          first = false;
        } else {
          m[0] = __hmax2(m[0], h[0]), m[1] = __hmax2(m[1], h[1]), m[2] = __hmax2(m[2], h[2]), m[3] = __hmax2(m[3], h[3]);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + (((int64_t)img * 56 + oy) * 56 + ox) * 64 + cg * 8) = *reinterpret_cast<uint4*>(m);
  }
}

// Global average pool over 7x7 (fp32 accumulate) + optional Linear(512,k) in fp32. One CTA per patch.
__global__ void __launch_bounds__(256) k_avgpool_fc(const __nv_bfloat16* __restrict__ in, float* __restrict__ feats,
                                                    float* __restrict__ logits, const float* __restrict__ fc_w,
                                                    const float* __restrict__ fc_b, int num_classes,
                                                    const int* __restrict__ n_dev, int n_base) {
  const int img = blockIdx.x;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  if (n_dev && img >= __ldg(n_dev) - n_base) return;
  __shared__ float f[512];
  const __nv_bfloat16* src = in + (int64_t)img * 49 * 512;
  for (int c2 = threadIdx.x; c2 < 256; c2 += 256) {
    float s0 = 0.f, s1 = 0.f;
    for (int px = 0; px < 49; px++) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + px * 512 + 2 * c2);
      s0 += __bfloat162float(v.x), s1 += __bfloat162float(v.y);
    }
    s0 *= (1.0f / 49.0f), s1 *= (1.0f / 49.0f);
    f[2 * c2] = s0, f[2 * c2 + 1] = s1;
    feats[(int64_t)img * 512 + 2 * c2] = s0;
    feats[(int64_t)img * 512 + 2 * c2 + 1] = s1;
  }
  if (!logits) return;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < num_classes; k += 8) {
    float s = 0.f;
    for (int c = lane; c < 512; c += 32) s += f[c] * __ldg(fc_w + k * 512 + c);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logits[(int64_t)img * num_classes + k] = s + __ldg(fc_b + k);
  }
}

// bf16 NHWC3 [n,224,224,3] -> S2D16 [n,112,115,16] (used when the caller hands the plain layout).
__global__ void __launch_bounds__(256) k_pack_s2d16(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int n_img) {
  const int64_t total = (int64_t)n_img * 112 * kS2dW;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(t % kS2dW) - 2;
    const int Y = (int)((t / kS2dW) % 112);
    const int64_t img = t / (112 * kS2dW);
    uint16_t v[16];
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < 2; dx++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          v[(dy * 2 + dx) * 3 + c] = (X >= 0 && X < 112) ? in[((img * 224 + 2 * Y + dy) * 224 + 2 * X + dx) * 3 + c] : (uint16_t)0;
    v[12] = v[13] = v[14] = v[15] = 0;
    uint4* dst = reinterpret_cast<uint4*>(out + t * 16);
    dst[0] = *reinterpret_cast<uint4*>(&v[0]);
    dst[1] = *reinterpret_cast<uint4*>(&v[8]);
  }
}

// ==========================================================================================
// host: tensor maps, launches
// ==========================================================================================
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn g_encode_im2col = nullptr;
static EncodeTiledFn g_encode_tiled = nullptr;
static thread_local int g_num_sms = 0;   // SM count of the CURRENT device, refreshed by init_driver_api() on every entry
static bool g_use_fused_stem = true;   // HIPAC_FUSED_STEM=0 runs conv1 and the max pool as two kernels
static bool g_fuse_downsample = true; // HIPAC_FUSE_DS=0 runs the 1x1 projection shortcuts as separate kernels
static bool g_use_row_kernels = true;  // HIPAC_CONV_ROWS=0 forces the im2col kernel everywhere (A/B comparison)
static bool g_use_cta_pairs = true;    // HIPAC_CTA_PAIRS=0: single-CTA row kernels instead of the cta_group::2 ones
static bool g_use_cta_pairs_stem = true;   // HIPAC_CTA_PAIRS_STEM=0: the fused stem on single CTAs
static bool g_use_cta_pairs_c128_im2col = false;  // HIPAC_CTA_PAIRS_C128_IM2COL=1: layer2.0.conv1 on CTA pairs instead of a single CTA with resident weights (measured 2 % slower)
static bool g_use_s2_rows = true;          // HIPAC_S2_ROWS=0: layer2.0.conv1 through im2col (k_conv_umma<128, RESB>)
static bool g_use_cta_pairs_c256 = true;   // HIPAC_CTA_PAIRS_C256=0: layer3 / layer4 on single CTAs (k_conv_umma<256>)
static bool g_resident_weights = true; // HIPAC_RESIDENT_B=0: the im2col kernel streams the weights of layer2.0.conv1 like everywhere else
static bool g_tma_epilogue_c128 = true;   // HIPAC_TMA_EPILOGUE_C128=0: the 128-channel residual layer keeps per-thread stores / residual loads
static bool g_tma_epilogue = true;     // HIPAC_TMA_EPILOGUE=0: 64-channel layers store / fetch the residual per thread (k_conv3x3_rows)
static bool g_use_cta_pairs_c64 = true;   // HIPAC_CTA_PAIRS_C64=0: layer1 on single CTAs (k_conv3x3_rows_tma)

// A/B switches for measurements; read once (the workspace size depends on them).
static bool g_boustrophedon = true;   // alternate the tile order from layer to layer (A/B switch: HIPAC_BOUSTROPHEDON=0)
static void read_env_flags() {
  static std::once_flag once;
  std::call_once(once, [] {
    if (const char* e = getenv("HIPAC_CONV_ROWS")) g_use_row_kernels = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_CTA_PAIRS")) g_use_cta_pairs = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_CTA_PAIRS_C64")) g_use_cta_pairs_c64 = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_TMA_EPILOGUE")) g_tma_epilogue = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_RESIDENT_B")) g_resident_weights = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_CTA_PAIRS_C256")) g_use_cta_pairs_c256 = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_CTA_PAIRS_STEM")) g_use_cta_pairs_stem = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_S2_ROWS")) g_use_s2_rows = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_CTA_PAIRS_C128_IM2COL")) g_use_cta_pairs_c128_im2col = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_TMA_EPILOGUE_C128")) g_tma_epilogue_c128 = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_FUSED_STEM")) g_use_fused_stem = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_FUSE_DS")) g_fuse_downsample = atoi(e) != 0;
    if (const char* e = getenv("HIPAC_BOUSTROPHEDON")) g_boustrophedon = atoi(e) != 0;
  });
}

// Called at the top of every entry point: driver entry points once per process, the SM count for the CURRENT device.
static int init_driver_api() {
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!(g_encode_im2col && g_encode_tiled)) {
      cudaDriverEntryPointQueryResult q;
      void* fn = nullptr;
      HIPAC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
      HIPAC_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "driver lacks cuTensorMapEncodeIm2col");
      g_encode_im2col = (EncodeIm2colFn)fn;
      HIPAC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
      HIPAC_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "driver lacks cuTensorMapEncodeTiled");
      g_encode_tiled = (EncodeTiledFn)fn;
    }
  }
  if (int e = device_sm_count(&g_num_sms)) return e;
  read_env_flags();
  return 0;
}

// NHWC activation tensor [n][h][w][c] (bf16) as an im2col map: box = 128 output pixels x `chan_box` channels.
struct Im2colDesc {
  int n, h, w, c;                 // logical NHWC extents seen by TMA
  int64_t pix_stride, row_stride, img_stride;  // bytes
  int kh, kw, stride;             // filter extent and traversal stride
  int pad_w_lo, pad_w_hi, pad_h_lo, pad_h_hi;
};
static int make_im2col_map(CUtensorMap* map, const void* ptr, const Im2colDesc& d) {
  cuuint64_t dims[4] = {(cuuint64_t)d.c, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
  cuuint64_t strides[3] = {(cuuint64_t)d.pix_stride, (cuuint64_t)d.row_stride, (cuuint64_t)d.img_stride};
  int lower[2] = {-d.pad_w_lo, -d.pad_h_lo};
  int upper[2] = {d.pad_w_hi - (d.kw - 1), d.pad_h_hi - (d.kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
  const int n = d.n, h = d.h, w = d.w, c = d.c;
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper,
                               64u, (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed with CUresult " + std::to_string((int)r));
    return -5;
  }
  // Same driver quirk CUTLASS works around (copy_traits_sm90_im2col.hpp): for tensors smaller than 128 KiB
  // drivers <= 13.1 set a descriptor bit that breaks im2col loads.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (size_t)n * h * w * c * 2 < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return 0;
}

// Weights [cout][K] bf16, K-major: box = 64 (K) x bn rows, 128-byte swizzle.
static int make_weight_map(CUtensorMap* map, const void* ptr, int cout, int K, int bn) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return -5;
  }
  return 0;
}

// NHWC activation tensor as a tiled 4-D map whose box is one row-tile region: 64 channels x (w+2) x (r+2) x 1.
static int make_region_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int r) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(w + 2), (cuuint32_t)(r + 2), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult res = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (region map) failed with CUresult " + std::to_string((int)res));
    return -5;
  }
  return 0;
}

template <int BN, int KC, int W, int R, bool RESIDENT>
static int launch_rows_t(const uint8_t* d_packed, const PackedLayout& L, int layer, const void* in, const void* residual, void* out,
                         int n, bool relu, cudaStream_t stream, const char* name) {
  using Cfg = RowCfg<BN, KC, W, R, RESIDENT>;
  if (int e = ensure_dyn_smem(k_conv3x3_rows<BN, KC, W, R, RESIDENT>, Cfg::kSmemBytes)) return e;
  CUtensorMap tmA, tmB;
  if (int e = make_region_map(&tmA, in, n, W, W, KC * 64, R)) return e;
  if (int e = make_weight_map(&tmB, d_packed + L.w_off[layer], BN, 9 * KC * 64, BN)) return e;
  RowConvParams p;
  p.n_img = n, p.num_tiles = n * (W / R), p.relu = relu ? 1 : 0;
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  p.bias = reinterpret_cast<const float*>(d_packed + L.b_off[layer]);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  {
    ProfileScope ps(name, stream, 2.0 * n * W * W * BN * 9 * KC * 64);
    k_conv3x3_rows<BN, KC, W, R, RESIDENT><<<grid, conv_threads(BN), Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// NHWC activation tensor as a strided tiled map for the fused 1x1 / stride-2 projection: the box spans 2*wp x 2*r input
// pixels with element strides 2, i.e. delivers wp x r pixels x 64 channels (out-of-range pixels zero filled).
static int make_strided_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int wp, int r) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(2 * wp), (cuuint32_t)(2 * r), 1};
  cuuint32_t estr[4] = {1, 2, 2, 1};
  CUresult res = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (strided projection map) failed with CUresult " + std::to_string((int)res));
    return -5;
  }
  return 0;
}

// NHWC activation tensor as a tiled map whose box is one OUTPUT row tile: 64 channels x w x r x 1 (TMA store of the
// epilogue's staging tile / TMA load of the residual tile).
static int make_tile_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int r) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)w, (cuuint32_t)r, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult res = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (tile map) failed with CUresult " + std::to_string((int)res));
    return -5;
  }
  return 0;
}

// 64-channel row kernel with the TMA epilogue (conv_rows_tma.cuh).
template <int KC, int W, int R>
static int launch_rows_tma_t(const uint8_t* d_packed, const PackedLayout& L, int layer, const void* in, const void* residual, void* out,
                             int n, bool relu, cudaStream_t stream, const char* name) {
  using Cfg = RowTmaCfg<KC, W, R>;
  if (int e = ensure_dyn_smem(k_conv3x3_rows_tma<KC, W, R>, Cfg::kSmemBytes)) return e;
  CUtensorMap tmA, tmB, tmO, tmR;
  if (int e = make_region_map(&tmA, in, n, W, W, KC * 64, R)) return e;
  if (int e = make_weight_map(&tmB, d_packed + L.w_off[layer], 64, 9 * KC * 64, 64)) return e;
  if (int e = make_tile_map(&tmO, out, n, W, W, 64, R)) return e;
  if (residual) {
    if (int e = make_tile_map(&tmR, residual, n, W, W, 64, R)) return e;
  } else {
    tmR = tmO;
  }
  RowConvParams p;
  p.n_img = n, p.num_tiles = n * (W / R), p.relu = relu ? 1 : 0;
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  p.bias = reinterpret_cast<const float*>(d_packed + L.b_off[layer]);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  {
    ProfileScope ps(name, stream, 2.0 * n * W * W * 64 * 9 * KC * 64);
    HIPAC_CHECK_CUDA(launch_ex(k_conv3x3_rows_tma<KC, W, R>, dim3(grid), dim3(conv_threads(64)), Cfg::kSmemBytes, stream, 1, true, tmA, tmB, tmO,
                               tmR, p));
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// CTA-pair row kernel (conv_rows2.cuh).  weights = [BN][9*KC*64 (+ KDS*64)] K-major at w_ptr; ds_in = block input of the
// fused projection shortcut (KDS = 1) or null.
template <int BN, int KC, int W, int R, int KDS, bool TEPI = (BN == 64)>
static int launch_rows2_t(const void* w_ptr, const float* bias, const void* in, const void* ds_in, const void* residual, void* out,
                          int n, bool relu, cudaStream_t stream, const char* name, double flops) {
  using Cfg = Row2Cfg<BN, KC, W, R, KDS, TEPI>;
  if (int e = ensure_dyn_smem(k_conv3x3_rows2<BN, KC, W, R, KDS, TEPI>, Cfg::kSmemBytes)) return e;
  CUtensorMap tmA, tmB, tmA2;
  if (int e = make_region_map(&tmA, in, n, W, W, KC * 64, R)) return e;
  if (int e = make_weight_map(&tmB, w_ptr, BN, (9 * KC + KDS) * 64, BN / 2)) return e;
  if (KDS) {
    if (int e = make_strided_map(&tmA2, ds_in, n, 2 * W, 2 * W, 64, W + 2, R)) return e;
  } else {
    tmA2 = tmA;
  }
  CUtensorMap tmO = tmA, tmR = tmA;
  if (Cfg::kTmaEpi) {
    if (int e = make_tile_map(&tmO, out, n, W, W, BN, R)) return e;
    tmR = tmO;
    if (residual)
      if (int e = make_tile_map(&tmR, residual, n, W, W, BN, R)) return e;
  }
  RowConvParams p;
  p.n_img = n, p.num_tiles = n * (W / R), p.relu = relu ? 1 : 0;
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int pairs = (p.num_tiles + 1) / 2;
  const int max_pairs = g_num_sms / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  {
    ProfileScope ps(name, stream, flops);
    HIPAC_CHECK_CUDA(launch_ex(k_conv3x3_rows2<BN, KC, W, R, KDS, TEPI>, dim3((unsigned)grid), dim3((unsigned)conv_threads(BN)), Cfg::kSmemBytes,
                               stream, 2, true, tmA, tmB, tmA2, tmO, tmR, p));
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// layer2.0.conv1 (3x3 / stride 2, 64 -> 128, 56x56 -> 28x28) as a parity-split row kernel on CTA pairs (conv_rows2_s2.cuh).
static int launch_rows2_s2(const void* w_ptr, const float* bias, const void* in, void* out, int n, bool relu, cudaStream_t stream,
                           const char* name, double flops) {
  using Cfg = S2Cfg;
  if (int e = ensure_dyn_smem(k_conv3x3s2_rows2, Cfg::kSmemBytes)) return e;
  CUtensorMap tmE, tmO, tmB;
  if (int e = make_strided_map(&tmE, in, n, 56, 56, 64, Cfg::Wp, Cfg::R)) return e;
  if (int e = make_strided_map(&tmO, in, n, 56, 56, 64, Cfg::Wp, Cfg::R + 1)) return e;
  if (int e = make_weight_map(&tmB, w_ptr, Cfg::BN, 9 * 64, Cfg::BN / 2)) return e;
  RowConvParams p;
  p.n_img = n, p.num_tiles = n * (Cfg::W / Cfg::R), p.relu = relu ? 1 : 0;
  p.bias = bias, p.residual = nullptr, p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  const int pairs = (p.num_tiles + 1) / 2, max_pairs = g_num_sms / 2;
  const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
  {
    ProfileScope ps(name, stream, flops);
    HIPAC_CHECK_CUDA(launch_ex(k_conv3x3s2_rows2, dim3((unsigned)grid), dim3((unsigned)conv_threads(128)), Cfg::kSmemBytes, stream, 2, true, tmE,
                               tmO, tmB, p));
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Fused conv1 + BN + ReLU + maxpool on the S2D16 batch -> [n][56][56][64].
static int run_stem(const uint8_t* d_packed, const PackedLayout& L, const void* in, void* out, int n, cudaStream_t stream) {
  const bool pair = g_use_cta_pairs && g_use_cta_pairs_stem;
  if (int e = pair ? ensure_dyn_smem(k_conv1_pool<true>, kStemSmemPair) : ensure_dyn_smem(k_conv1_pool<false>, kStemSmem)) return e;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {16, (cuuint64_t)kS2dW, 112, (cuuint64_t)n};
    cuuint64_t strides[3] = {32, (cuuint64_t)kS2dW * 32, (cuuint64_t)112 * kS2dW * 32};
    cuuint32_t box[4] = {16, (cuuint32_t)kS2dW, (cuuint32_t)kStemRows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult res = g_encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (stem region map) failed with CUresult " + std::to_string((int)res));
      return -5;
    }
  }
  if (int e = make_weight_map(&tmB, d_packed + L.w_off[0], 64, 256, pair ? 32 : 64)) return e;
  StemParams p;
  p.num_blocks = n * (56 / kStemPB);
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  p.bias = reinterpret_cast<const float*>(d_packed + L.b_off[0]);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int items = pair ? (p.num_blocks + 1) / 2 : p.num_blocks, slots = pair ? g_num_sms / 2 : g_num_sms;
  const int grid = (pair ? 2 : 1) * (items < slots ? items : slots);
  {
    ProfileScope ps("conv1_pool_fused", stream, 2.0 * n * 112 * 112 * 64 * 147);
    if (pair)
      HIPAC_CHECK_CUDA(launch_ex(k_conv1_pool<true>, dim3(grid), dim3(kStemThreads), kStemSmemPair, stream, 2, true, tmA, tmB, p));
    else
      HIPAC_CHECK_CUDA(launch_ex(k_conv1_pool<false>, dim3(grid), dim3(kStemThreads), kStemSmem, stream, 1, true, tmA, tmB, p));
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, bool RESB = false>
static int launch_conv_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream,
                         const char* name, double flops, const CUtensorMap* tmA2 = nullptr) {
  using Cfg = ConvCfg<BN, RESB>;
  if (int e = ensure_dyn_smem(k_conv_umma<BN, RESB>, Cfg::kSmemBytes)) return e;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  {
    ProfileScope ps(name, stream, flops);
    HIPAC_CHECK_CUDA(launch_ex(k_conv_umma<BN, RESB>, dim3(grid), dim3(conv_threads(BN)), Cfg::kSmemBytes, stream, 1, true, tmA, tmB,
                               tmA2 ? *tmA2 : tmA, p));
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// im2col layers on CTA pairs (conv_umma2.cuh); tmB must have been built with a box of BN/2 weight rows.
template <int BN>
static int launch_conv2(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream, const char* name,
                        double flops, const CUtensorMap* tmA2 = nullptr) {
  using Cfg = Conv2Cfg<BN>;
  if (int e = ensure_dyn_smem(k_conv_umma2<BN>, Cfg::kSmemBytes)) return e;
  const int ptiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  const int max_pairs = g_num_sms / 2;
  const int grid = 2 * (ptiles < max_pairs ? ptiles : max_pairs);
  {
    ProfileScope ps(name, stream, flops);
    HIPAC_CHECK_CUDA(launch_ex(k_conv_umma2<BN>, dim3((unsigned)grid), dim3((unsigned)conv_threads(BN)), Cfg::kSmemBytes, stream, 2, true, tmA,
                               tmB, tmA2 ? *tmA2 : tmA, p));
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// One conv layer of the network on `n` patches. in/out are bf16 NHWC (layer 0: S2D16 input).
static int run_conv(const uint8_t* d_packed, const PackedLayout& L, int layer, const void* in, const void* residual, void* out,
                    int n, bool relu, cudaStream_t stream) {
  const ConvSpec& cs = kConvs[layer];
  const int K = conv_gemm_k(layer);
  if (g_use_row_kernels && g_use_cta_pairs && g_use_s2_rows && cs.k == 3 && cs.stride == 2 && cs.cin == 64 && cs.cout == 128 && cs.hin == 56 &&
      residual == nullptr)
    return launch_rows2_s2(d_packed + L.w_off[layer], reinterpret_cast<const float*>(d_packed + L.b_off[layer]), in, out, n, relu, stream,
                           "conv3x3_c128", 2.0 * n * 28 * 28 * 128 * 576);
  if (g_use_row_kernels && cs.k == 3 && cs.stride == 1) {
    const float* bias = reinterpret_cast<const float*>(d_packed + L.b_off[layer]);
    if (cs.cin == 64 && cs.hin == 56) {
      if (g_use_cta_pairs && g_use_cta_pairs_c64)
        return launch_rows2_t<64, 1, 56, 2, 0>(d_packed + L.w_off[layer], bias, in, nullptr, residual, out, n, relu, stream, "conv3x3_c64",
                                               2.0 * n * 56 * 56 * 64 * 576);
      if (g_tma_epilogue) return launch_rows_tma_t<1, 56, 2>(d_packed, L, layer, in, residual, out, n, relu, stream, "conv3x3_c64");
      return launch_rows_t<64, 1, 56, 2, true>(d_packed, L, layer, in, residual, out, n, relu, stream, "conv3x3_c64");
    }
    if (cs.cin == 128 && cs.hin == 28) {
      if (g_use_cta_pairs && residual && g_tma_epilogue_c128)   // the residual variant: staged epilogue (see conv_rows2.cuh)
        return launch_rows2_t<128, 2, 28, 4, 0, true>(d_packed + L.w_off[layer], bias, in, nullptr, residual, out, n, relu, stream,
                                                      "conv3x3_c128", 2.0 * n * 28 * 28 * 128 * 1152);
      if (g_use_cta_pairs)
        return launch_rows2_t<128, 2, 28, 4, 0>(d_packed + L.w_off[layer], bias, in, nullptr, residual, out, n, relu, stream, "conv3x3_c128",
                                                2.0 * n * 28 * 28 * 128 * 1152);
      return launch_rows_t<128, 2, 28, 4, false>(d_packed, L, layer, in, residual, out, n, relu, stream, "conv3x3_c128");
    }
  }
  ConvParams p;
  p.M_total = n * cs.hout * cs.hout;
  p.hw_out = cs.hout * cs.hout, p.wout = cs.hout;
  p.cout = cs.cout;
  p.relu = relu ? 1 : 0;
  p.bias = reinterpret_cast<const float*>(d_packed + L.b_off[layer]);
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.num_m_tiles = (p.M_total + kBM - 1) / kBM;
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  p.num_kb2 = 0, p.stride2 = 1;
  const int bn = cs.cout >= 256 ? 256 : (cs.cout >= 128 ? 128 : 64);
  p.num_n_tiles = cs.cout / bn;
  const bool pairs256 = bn >= 128 && layer != 0 && g_use_cta_pairs && (bn == 256 ? g_use_cta_pairs_c256 : g_use_cta_pairs_c128_im2col);
  CUtensorMap tmA, tmB;
  if (int e = make_weight_map(&tmB, d_packed + L.w_off[layer], cs.cout, K, pairs256 ? bn / 2 : bn)) return e;
  if (layer == 0) {
    // 7x7/s2/p3 over 3 channels == 4x4/s1 over the 2x2 space-to-depth image (16 ch), pad 2 before / 1 after.
    // The S2D16 batch stores the W padding explicitly ([n][112][115][16], zero columns 0,1,114), so the four
    // W-taps of one filter row are 64 CONTIGUOUS bf16: the conv is a 4x1 filter over a 64-"channel" view whose
    // pixel stride (32 B) is smaller than its channel extent (128 B).  H padding stays TMA zero fill.
    p.stride = 1, p.pad_w = 0, p.pad_h = 2, p.kw = 1, p.kc_blocks = 1, p.num_kb = 4;
    Im2colDesc d{n, 112, 112, 64, 32, (int64_t)kS2dW * 32, (int64_t)112 * kS2dW * 32, 4, 1, 1, 0, 0, 2, 1};
    if (int e = make_im2col_map(&tmA, in, d)) return e;
    return launch_conv_t<64>(tmA, tmB, p, stream, "conv1_7x7s2_c64", 2.0 * p.M_total * 64 * 147);
  }
  p.stride = cs.stride, p.pad_w = p.pad_h = cs.pad, p.kw = cs.k, p.kc_blocks = cs.cin / 64, p.num_kb = cs.k * cs.k * p.kc_blocks;
  Im2colDesc d{n, cs.hin, cs.hin, cs.cin, (int64_t)cs.cin * 2, (int64_t)cs.hin * cs.cin * 2, (int64_t)cs.hin * cs.hin * cs.cin * 2,
               cs.k, cs.k, cs.stride, cs.pad, cs.pad, cs.pad, cs.pad};
  if (int e = make_im2col_map(&tmA, in, d)) return e;
  static const char* kNames[4][2] = {{"conv3x3_c64", "conv1x1_c64"}, {"conv3x3_c128", "conv1x1_c128"},
                                     {"conv3x3_c256", "conv1x1_c256"}, {"conv3x3_c512", "conv1x1_c512"}};
  const int gi = cs.cout == 64 ? 0 : cs.cout == 128 ? 1 : cs.cout == 256 ? 2 : 3;
  const char* name = kNames[gi][cs.k == 1 ? 1 : 0];
  const double flops = 2.0 * p.M_total * cs.cout * K;
  if (pairs256) return bn == 256 ? launch_conv2<256>(tmA, tmB, p, stream, name, flops) : launch_conv2<128>(tmA, tmB, p, stream, name, flops);
  if (bn == 256) return launch_conv_t<256>(tmA, tmB, p, stream, name, flops);
  if (bn == 128 && g_resident_weights && p.num_n_tiles == 1 && p.num_kb <= ConvCfg<128, true>::kResBlocks)
    return launch_conv_t<128, true>(tmA, tmB, p, stream, name, flops);      // layer2.0.conv1: the 144 KB of weights stay in shared memory
  return bn == 128 ? launch_conv_t<128>(tmA, tmB, p, stream, name, flops) : launch_conv_t<64>(tmA, tmB, p, stream, name, flops);
}

// conv2 of a stage's first block with its projection shortcut fused: out = relu(conv3x3(in) + conv1x1_s2(block_in) + bias).
// Both GEMMs accumulate into the same TMEM tile (the shortcut is extra K-blocks), so the shortcut tensor never exists.
static int run_conv_ds_fused(const uint8_t* d_packed, const PackedLayout& L, int s, const void* in, const void* block_in, void* out,
                             int n, cudaStream_t stream) {
  const int lc = fused_conv_layer(s), ld = fused_ds_layer(s);
  const ConvSpec &cs = kConvs[lc], &ds = kConvs[ld];
  if (s == 0 && g_use_row_kernels && g_use_cta_pairs) {
    // layer2.0: 3x3 128 -> 128 at 28x28 + 1x1/s2 64 -> 128 from the 56x56 block input, on CTA pairs with resident weights
    return launch_rows2_t<128, 2, 28, 4, 1>(d_packed + L.wf_off[0], reinterpret_cast<const float*>(d_packed + L.bf_off[0]), in, block_in,
                                            nullptr, out, n, true, stream, "conv3x3+ds_c128", 2.0 * n * 28 * 28 * 128 * fused_gemm_k(0));
  }
  ConvParams p;
  p.M_total = n * cs.hout * cs.hout;
  p.hw_out = cs.hout * cs.hout, p.wout = cs.hout;
  p.cout = cs.cout, p.relu = 1;
  p.bias = reinterpret_cast<const float*>(d_packed + L.bf_off[s]);
  p.residual = nullptr;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.num_m_tiles = (p.M_total + kBM - 1) / kBM;
  p.n_dev = g_n_dev, p.n_base = g_n_base, p.reverse = g_reverse;
  const int bn = cs.cout >= 256 ? 256 : 128;
  p.num_n_tiles = cs.cout / bn;
  p.stride = 1, p.pad_w = p.pad_h = 1, p.kw = 3, p.kc_blocks = cs.cin / 64, p.num_kb = 9 * p.kc_blocks;
  p.num_kb2 = ds.cin / 64, p.stride2 = 2;
  const bool pairs256 = bn == 256 && g_use_cta_pairs && g_use_cta_pairs_c256;
  CUtensorMap tmA, tmA2, tmB;
  if (int e = make_weight_map(&tmB, d_packed + L.wf_off[s], cs.cout, fused_gemm_k(s), pairs256 ? bn / 2 : bn)) return e;
  Im2colDesc d{n, cs.hin, cs.hin, cs.cin, (int64_t)cs.cin * 2, (int64_t)cs.hin * cs.cin * 2, (int64_t)cs.hin * cs.hin * cs.cin * 2,
               3, 3, 1, 1, 1, 1, 1};
  if (int e = make_im2col_map(&tmA, in, d)) return e;
  Im2colDesc d2{n, ds.hin, ds.hin, ds.cin, (int64_t)ds.cin * 2, (int64_t)ds.hin * ds.cin * 2, (int64_t)ds.hin * ds.hin * ds.cin * 2,
                1, 1, 2, 0, 0, 0, 0};
  if (int e = make_im2col_map(&tmA2, block_in, d2)) return e;
  static const char* kNames[3] = {"conv3x3+ds_c128", "conv3x3+ds_c256", "conv3x3+ds_c512"};
  const double flops = 2.0 * p.M_total * cs.cout * fused_gemm_k(s);
  if (pairs256) return launch_conv2<256>(tmA, tmB, p, stream, kNames[s], flops, &tmA2);
  return bn == 256 ? launch_conv_t<256>(tmA, tmB, p, stream, kNames[s], flops, &tmA2)
                   : launch_conv_t<128>(tmA, tmB, p, stream, kNames[s], flops, &tmA2);
}

static uint16_t host_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);  // inf / nan passthrough
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// per-chunk activation buffers (bf16 NHWC), sizes per patch
constexpr size_t kC1Bytes = (size_t)112 * 112 * 64 * 2;  // conv1 output
constexpr size_t kActBytes = (size_t)56 * 56 * 64 * 2;   // largest post-pool activation
constexpr size_t kS2dBytes = (size_t)112 * kS2dW * 16 * 2;

static int clamp_chunk(int chunk, int n) {
  if (chunk <= 0) chunk = 4096;
  if (chunk > n) chunk = n;
  return chunk < 1 ? 1 : chunk;
}

}  // namespace hipac

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace hipac;

extern "C" size_t hipac_resnet18_packed_bytes(int num_classes) { return packed_layout(num_classes).total; }

extern "C" int hipac_resnet18_pack(const float* const* t, int num_tensors, int num_classes, float bn_eps, void* h_packed,
                                   size_t packed_bytes) {
  HIPAC_REQUIRE(num_tensors == HIPAC_RESNET18_NUM_TENSORS, "expected 102 tensors (20 x [conv, bn.w, bn.b, bn.mean, bn.var] + fc.w + fc.b)");
  HIPAC_REQUIRE(num_classes >= 0 && num_classes <= 1024, "bad num_classes");
  const PackedLayout L = packed_layout(num_classes);
  HIPAC_REQUIRE(h_packed && packed_bytes >= L.total, "packed buffer too small");
  memset(h_packed, 0, L.total);
  uint8_t* dst = reinterpret_cast<uint8_t*>(h_packed);
  for (int l = 0; l < HIPAC_RESNET18_NUM_CONVS; l++) {
    const ConvSpec& cs = kConvs[l];
    const float *w = t[5 * l], *g = t[5 * l + 1], *b = t[5 * l + 2], *mu = t[5 * l + 3], *var = t[5 * l + 4];
    HIPAC_REQUIRE(w && g && b && mu && var, "null tensor");
    uint16_t* wp = reinterpret_cast<uint16_t*>(dst + L.w_off[l]);
    float* bp = reinterpret_cast<float*>(dst + L.b_off[l]);
    const int K = conv_gemm_k(l);
    for (int o = 0; o < cs.cout; o++) {
      const float scale = g[o] / sqrtf(var[o] + bn_eps);  // eval-mode BN folded in fp32
      bp[o] = b[o] - mu[o] * scale;
      for (int c = 0; c < cs.cin; c++)
        for (int r = 0; r < cs.k; r++)
          for (int s = 0; s < cs.k; s++) {
            const float v = w[(((size_t)o * cs.cin + c) * cs.k + r) * cs.k + s] * scale;
            size_t kidx;
            if (l == 0) {
              // kh = 2a+dy-1, kw = 2b+dx-1 (filter padded to 8x8 with a zero first row/column)
              const int a = (r + 1) >> 1, dy = (r + 1) & 1, bb = (s + 1) >> 1, dx = (s + 1) & 1;
              kidx = (size_t)(a * 4 + bb) * 16 + (dy * 2 + dx) * 3 + c;
            } else {
              kidx = (size_t)(r * cs.k + s) * cs.cin + c;
            }
            wp[(size_t)o * K + kidx] = host_bf16(v);
          }
    }
  }
  for (int s = 0; s < 3; s++) {
    const int lc = fused_conv_layer(s), ld = fused_ds_layer(s);
    const int cout = kConvs[lc].cout, k1 = conv_gemm_k(lc), k2 = conv_gemm_k(ld), kf = k1 + k2;
    const uint16_t* w1 = reinterpret_cast<const uint16_t*>(dst + L.w_off[lc]);
    const uint16_t* w2 = reinterpret_cast<const uint16_t*>(dst + L.w_off[ld]);
    const float *b1 = reinterpret_cast<const float*>(dst + L.b_off[lc]), *b2 = reinterpret_cast<const float*>(dst + L.b_off[ld]);
    uint16_t* wf = reinterpret_cast<uint16_t*>(dst + L.wf_off[s]);
    float* bf = reinterpret_cast<float*>(dst + L.bf_off[s]);
    for (int o = 0; o < cout; o++) {
      memcpy(wf + (size_t)o * kf, w1 + (size_t)o * k1, (size_t)k1 * 2);
      memcpy(wf + (size_t)o * kf + k1, w2 + (size_t)o * k2, (size_t)k2 * 2);
      bf[o] = b1[o] + b2[o];
    }
  }
  if (num_classes > 0) {
    const float *fw = t[100], *fb = t[101];
    HIPAC_REQUIRE(fw && fb, "num_classes > 0 needs fc weight and bias");
    memcpy(dst + L.fc_w_off, fw, (size_t)num_classes * 512 * 4);
    memcpy(dst + L.fc_b_off, fb, (size_t)num_classes * 4);
  }
  return 0;
}

extern "C" size_t hipac_resnet18_workspace_bytes(int n_patches, int chunk) {
  if (n_patches <= 0) return 256;
  const size_t c = (size_t)clamp_chunk(chunk, n_patches);
  read_env_flags();
  // conv1's 112x112x64 output and the projection-shortcut tensor only exist in the unfused A/B modes; the S2D16
  // repack buffer only when the caller hands the plain NHWC3 layout (sized for it unconditionally: 0.4 MB / patch)
  return c * ((g_use_fused_stem ? 0 : kC1Bytes) + (g_fuse_downsample ? 3 : 4) * kActBytes + kS2dBytes) + 6 * 1024;
}

extern "C" int hipac_resnet18_conv_layer(const void* d_packed, int num_classes, int layer, const void* d_in,
                                         const void* d_residual, void* d_out, int n_patches, int relu, void* stream_) {
  HIPAC_REQUIRE(d_packed && d_in && d_out, "null pointer");
  HIPAC_REQUIRE(layer >= 0 && layer < HIPAC_RESNET18_NUM_CONVS, "layer index out of range");
  HIPAC_REQUIRE(n_patches > 0, "n_patches must be positive");
  if (int e = init_driver_api()) return e;
  const PackedLayout L = packed_layout(num_classes);
  return run_conv(reinterpret_cast<const uint8_t*>(d_packed), L, layer, d_in, d_residual, d_out, n_patches, relu != 0,
                  (cudaStream_t)stream_);
}

extern "C" int hipac_resnet18_conv_ds_fused(const void* d_packed, int num_classes, int stage, const void* d_in, const void* d_block_in,
                                            void* d_out, int n_patches, void* stream_) {
  HIPAC_REQUIRE(d_packed && d_in && d_block_in && d_out && n_patches > 0, "bad arguments");
  HIPAC_REQUIRE(stage >= 0 && stage < 3, "stage must be 0 (layer2), 1 (layer3) or 2 (layer4)");
  if (int e = init_driver_api()) return e;
  const PackedLayout L = packed_layout(num_classes);
  return run_conv_ds_fused(reinterpret_cast<const uint8_t*>(d_packed), L, stage, d_in, d_block_in, d_out, n_patches,
                           (cudaStream_t)stream_);
}

extern "C" int hipac_resnet18_stem(const void* d_packed, int num_classes, const void* d_in, void* d_out, int n_patches,
                                   void* stream_) {
  HIPAC_REQUIRE(d_packed && d_in && d_out && n_patches > 0, "bad arguments");
  if (int e = init_driver_api()) return e;
  const PackedLayout L = packed_layout(num_classes);
  return run_stem(reinterpret_cast<const uint8_t*>(d_packed), L, d_in, d_out, n_patches, (cudaStream_t)stream_);
}

static int forward_impl(const void* d_packed, int num_classes, const void* d_batch, int layout, int n_patches,
                        float* d_feats, float* d_logits, void* d_workspace, size_t workspace_bytes, int chunk,
                        void* stream_, const int* d_count) {
  cudaStream_t stream = (cudaStream_t)stream_;
  struct CountScope {   // every launch below picks the device count up from these thread-locals
    explicit CountScope(const int* c) { g_n_dev = c, g_n_base = 0, g_reverse = 0; }
    ~CountScope() { g_n_dev = nullptr, g_n_base = 0, g_reverse = 0; }
  } count_scope(d_count);
  HIPAC_REQUIRE(n_patches >= 0, "negative n_patches");
  if (n_patches == 0) return 0;
  HIPAC_REQUIRE(d_packed && d_batch && d_feats && d_workspace, "null pointer");
  HIPAC_REQUIRE(layout == HIPAC_LAYOUT_S2D16_BF16 || layout == HIPAC_LAYOUT_NHWC3_BF16, "unknown batch layout");
  HIPAC_REQUIRE(!d_logits || num_classes > 0, "logits requested but the packed weights have no classifier head");
  HIPAC_REQUIRE(workspace_bytes >= hipac_resnet18_workspace_bytes(n_patches, chunk), "workspace too small");
  HIPAC_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
  if (int e = init_driver_api()) return e;
  const PackedLayout L = packed_layout(num_classes);
  const uint8_t* pk = reinterpret_cast<const uint8_t*>(d_packed);
  const int cmax = clamp_chunk(chunk, n_patches);
  uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
  auto carve = [&](size_t bytes) {
    uint8_t* r = ws;
    ws += align_up(bytes, 1024);
    return r;
  };
  uint8_t* c1 = g_use_fused_stem ? nullptr : carve(cmax * kC1Bytes);
  uint8_t* A = carve(cmax * kActBytes);
  uint8_t* B = carve(cmax * kActBytes);
  uint8_t* C = carve(cmax * kActBytes);
  uint8_t* D = g_fuse_downsample ? nullptr : carve(cmax * kActBytes);
  uint8_t* s2d = carve(cmax * kS2dBytes);

  for (int i0 = 0; i0 < n_patches; i0 += cmax) {
    const int n = n_patches - i0 < cmax ? n_patches - i0 : cmax;
    g_n_base = i0;
    const void* x0;
    if (layout == HIPAC_LAYOUT_S2D16_BF16) {
      x0 = reinterpret_cast<const uint8_t*>(d_batch) + (size_t)i0 * kS2dBytes;
    } else {
      const uint16_t* src = reinterpret_cast<const uint16_t*>(d_batch) + (size_t)i0 * 224 * 224 * 3;
      {
        ProfileScope ps("pack_s2d16", stream, (double)n * (224 * 224 * 3 * 2 + kS2dBytes));
        k_pack_s2d16<<<g_num_sms * 8, 256, 0, stream>>>(src, reinterpret_cast<uint16_t*>(s2d), n);
      }
      count_launch(1);
      x0 = s2d;
    }
    // the batch was written in ascending patch order, so the first layer walks it descending, the next ascending, ...
    int order = 1;
    auto next_order = [&]() {
      g_reverse = g_boustrophedon ? order : 0;
      order ^= 1;
    };
    auto conv = [&](int layer, const void* in, const void* res, void* out, bool relu) {
      next_order();
      return run_conv(pk, L, layer, in, res, out, n, relu, stream);
    };
    int e = 0;
    if (g_use_fused_stem) {
      next_order();
      if ((e = run_stem(pk, L, x0, A, n, stream))) return e;
    } else {
      if ((e = conv(0, x0, nullptr, c1, true))) return e;
      {
        ProfileScope ps("maxpool3x3s2", stream, (double)n * (kC1Bytes + kActBytes));
        k_maxpool<<<g_num_sms * 8, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(c1), reinterpret_cast<__nv_bfloat16*>(A), n);
      }
      count_launch(1);
    }
    // layer1 (two basic blocks, identity shortcuts)
    if ((e = conv(1, A, nullptr, B, true)) || (e = conv(2, B, A, C, true))) return e;
    if ((e = conv(3, C, nullptr, B, true)) || (e = conv(4, B, C, A, true))) return e;
    // layer2..4: first block has a strided 1x1 projection shortcut
    for (int s = 0; s < 3; s++) {
      const int l0 = 5 + 5 * s;
      if ((e = conv(l0, A, nullptr, B, true))) return e;           // 3x3 / stride 2
      if (g_fuse_downsample) {
        next_order();
        if ((e = run_conv_ds_fused(pk, L, s, B, A, C, n, stream))) return e;   // 3x3 + 1x1/s2 projection + ReLU, one accumulator
      } else {
        if ((e = conv(l0 + 2, A, nullptr, D, false))) return e;    // downsample 1x1 / stride 2 (+BN), no ReLU
        if ((e = conv(l0 + 1, B, D, C, true))) return e;           // 3x3 + shortcut + ReLU
      }
      if ((e = conv(l0 + 3, C, nullptr, B, true)) || (e = conv(l0 + 4, B, C, A, true))) return e;
    }
    ProfileScope ps_pool("avgpool_fc", stream, (double)n * (49 * 512 * 2 + 512 * 4));
    HIPAC_CHECK_CUDA(launch_ex(k_avgpool_fc, dim3(n), dim3(256), 0, stream, 1, true, reinterpret_cast<const __nv_bfloat16*>(A),
                               d_feats + (size_t)i0 * 512, d_logits ? d_logits + (size_t)i0 * num_classes : nullptr,
                               reinterpret_cast<const float*>(pk + L.fc_w_off), reinterpret_cast<const float*>(pk + L.fc_b_off),
                               num_classes, g_n_dev, g_n_base));
    count_launch(1);
  }
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hipac_resnet18_forward(const void* d_packed, int num_classes, const void* d_batch, int layout, int n_patches,
                                      float* d_feats, float* d_logits, void* d_workspace, size_t workspace_bytes, int chunk,
                                      void* stream) {
  return forward_impl(d_packed, num_classes, d_batch, layout, n_patches, d_feats, d_logits, d_workspace, workspace_bytes, chunk,
                      stream, nullptr);
}

extern "C" int hipac_resnet18_forward_dcount(const void* d_packed, int num_classes, const void* d_batch, int layout, int capacity,
                                             const int32_t* d_count, float* d_feats, float* d_logits, void* d_workspace,
                                             size_t workspace_bytes, int chunk, void* stream) {
  HIPAC_REQUIRE(d_count != nullptr, "null device count");
  return forward_impl(d_packed, num_classes, d_batch, layout, capacity, d_feats, d_logits, d_workspace, workspace_bytes, chunk,
                      stream, d_count);
}
