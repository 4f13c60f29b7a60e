// Row-tile 3x3 / stride-1 / pad-1 convolution (included by resnet18.cu inside namespace hipac).
//
// The im2col kernel re-reads every activation byte nine times through L2 (once per filter tap), which
// makes the 64- and 128-channel layers L2-bandwidth bound.  Here one M-tile is R whole image rows in
// PADDED-width coordinates (Wp = W + 2 positions per row, R * Wp <= 128): a single tiled TMA load brings
// the (R + 2) x Wp input pixels of one 64-channel slice into shared memory -- TMA's out-of-bounds zero
// fill materialises the conv padding -- and all nine taps read that ONE staged region through UMMA
// descriptors that differ only in their start address (tap (r, s) starts r * Wp + s pixels = rows of
// 128 B further; tests/test_umma_shift_gpu.py pins that hardware behaviour).  Positions x >= W of each
// padded row produce junk accumulator rows that are simply not stored.
//
//   RESIDENT = true : the whole weight matrix (9 * KC * BN * 128 B) stays in shared memory for the
//                     lifetime of the persistent CTA (64-channel layers: 72 KB)
//   RESIDENT = false: weights stream through a ring, one [BN x 64] block per (tap, channel slice)
#pragma once

struct RowConvParams {
  int n_img;
  int num_tiles;  // n_img * (H / R)
  int relu;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  const int* n_dev;  // device-count mode, see effective_patches()
  int n_base;
  int reverse;       // tile order, see g_reverse
};

template <int BN, int KC, int W, int R, bool RESIDENT>
struct RowCfg {
  static constexpr int Wp = W + 2;
  static constexpr int kRegionRows = 128 + 2 * Wp + 2;                       // rows any tap's 128-row window can touch
  static constexpr int kRegionBytes = (kRegionRows * 128 + 1023) / 1024 * 1024;
  static constexpr int kLoadBytes = (R + 2) * Wp * 128;                      // bytes one region TMA delivers
  static constexpr int kAStages = RESIDENT ? 3 : 2;
  static constexpr int kBBlock = BN * 128;                                   // one (tap, slice) weight block
  static constexpr int kBStages = RESIDENT ? 9 * KC : 6;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kAStages * kRegionBytes + kBStages * kBBlock + 1024 + 256;
  static_assert(R * Wp <= 128, "tile does not fit one UMMA M = 128");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BN, int KC, int W, int R, bool RESIDENT>
__global__ void __launch_bounds__(conv_threads(BN), 1)
k_conv3x3_rows(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const RowConvParams p) {
  using Cfg = RowCfg<BN, KC, W, R, RESIDENT>;
  constexpr int Wp = Cfg::Wp, H = W, TILES_PER_IMG = H / R;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;
  uint8_t* sB = base + Cfg::kAStages * Cfg::kRegionBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + Cfg::kBStages * Cfg::kBBlock);
  uint64_t* a_empty = a_full + Cfg::kAStages;
  uint64_t* b_full = a_empty + Cfg::kAStages;   // RESIDENT: only b_full[0] is used (all weights landed)
  uint64_t* b_empty = b_full + 8;
  uint64_t* tfull = b_empty + 8;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < Cfg::kAStages; s++) ptx::mbar_init(&a_full[s], 1), ptx::mbar_init(&a_empty[s], 1);
    for (int s = 0; s < 8; s++) ptx::mbar_init(&b_full[s], 1), ptx::mbar_init(&b_empty[s], 1);
    for (int a = 0; a < 2; a++) ptx::mbar_init(&tfull[a], 1), ptx::mbar_init(&tempty[a], 4);
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
  const int num_tiles = effective_patches(p.n_dev, p.n_base, p.n_img) * (p.num_tiles / p.n_img);

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      if (RESIDENT && ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&b_full[0], 9 * KC * Cfg::kBBlock);
        for (int kb = 0; kb < 9 * KC; kb++) ptx::tma_load_2d(sB + kb * Cfg::kBBlock, &tmB, &b_full[0], kb * 64, 0);
      }
      __syncwarp();
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int vt = p.reverse ? num_tiles - 1 - tile : tile;
        const int img = vt / TILES_PER_IMG, p0 = (vt - img * TILES_PER_IMG) * R;
        for (int kc = 0; kc < KC; kc++) {
          ptx::mbar_wait(&a_empty[sa], pa ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(&a_full[sa], Cfg::kLoadBytes);
            ptx::tma_load_4d(sA + sa * Cfg::kRegionBytes, &tmA, &a_full[sa], kc * 64, -1, p0 - 1, img);
          }
          __syncwarp();
          if (++sa == Cfg::kAStages) sa = 0, pa ^= 1;
          if (!RESIDENT) {
            for (int tap = 0; tap < 9; tap++) {
              ptx::mbar_wait(&b_empty[sb], pb ^ 1);
              if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&b_full[sb], Cfg::kBBlock);
                ptx::tma_load_2d(sB + sb * Cfg::kBBlock, &tmB, &b_full[sb], (tap * KC + kc) * 64, 0);
              }
              __syncwarp();
              if (++sb == Cfg::kBStages) sb = 0, pb ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0, acc = 0, acc_phase = 0;
      if (RESIDENT) ptx::mbar_wait(&b_full[0], 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < KC; kc++) {
          ptx::mbar_wait(&a_full[sa], pa);
          ptx::tc_fence_after();
          const uint64_t a_region = ptx::make_smem_desc(ptx::smem_u32(sA + sa * Cfg::kRegionBytes), 128);
          if (RESIDENT) {
            const uint64_t b_all = ptx::make_smem_desc(ptx::smem_u32(sB + kc * Cfg::kBBlock), 128);
            if (ptx::elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; tap++) {
                // tap (r, s) = the staged region shifted by r*Wp + s pixels (128-byte rows): +8 per row in the address field
                const uint64_t adesc = a_region + (uint64_t)(((tap / 3) * Wp + tap % 3) * 8);
                const uint64_t bdesc = b_all + (uint64_t)(tap * KC * (Cfg::kBBlock >> 4));
#pragma unroll
                for (int k = 0; k < 4; k++) ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tap = 0; tap < 9; tap++) {
              ptx::mbar_wait(&b_full[sb], pb);
              ptx::tc_fence_after();
              const uint64_t bdesc = ptx::make_smem_desc(ptx::smem_u32(sB + sb * Cfg::kBBlock), 128);
              const uint64_t adesc = a_region + (uint64_t)(((tap / 3) * Wp + tap % 3) * 8);
              if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; k++) ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&b_empty[sb]);
              }
              __syncwarp();
              if (++sb == Cfg::kBStages) sb = 0, pb ^= 1;
            }
          }
          if (ptx::elect_one()) ptx::umma_commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == Cfg::kAStages) sa = 0, pa ^= 1;
        }
        if (ptx::elect_one()) ptx::umma_commit(&tfull[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int wq = warp & 3;
    const int pos = wq * 32 + lane;          // position in the padded-width tile
    const int rr = pos / Wp, x = pos - rr * Wp;
    const bool valid = rr < R && x < W;
    const int grp = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (epi_groups(BN) == 2 && (it & 1) != grp) continue;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int vt = p.reverse ? num_tiles - 1 - tile : tile;
      const int img = vt / TILES_PER_IMG, p0 = (vt - img * TILES_PER_IMG) * R;
      const size_t pix = ((size_t)img * H + p0 + rr) * W + x;
      epilogue_row<BN>(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BN, p.bias, p.residual ? p.residual + pix * BN : nullptr,
                       p.out + pix * BN, p.relu, valid, &tfull[acc], acc_phase);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}
