// 3x3 / STRIDE-2 / pad-1 convolution, 64 -> 128 channels, 56x56 -> 28x28 (layer2.0.conv1) as a row-tile kernel on CTA pairs.
// Included by resnet18.cu inside namespace hipac after conv_rows2.cuh (same pair protocol, RowConvParams, epilogue_row).
//
// Why.  The im2col form of this layer is bound by the TMA unit: a stride-2 im2col load is one 128-byte request per output
// pixel and tap, 2.25 x the input, and ncu shows the producer warp blocked on TMA issue and the MMA warp waiting for data a
// third of the time (tensor pipe 37 %).  Here the input of a tile is loaded ONCE, split by the parity of its row and
// column into four dense sub-images (tiled TMA boxes with element strides 2): output (p, q), tap (r, s) reads input
// (2p + r - 1, 2q + s - 1), i.e. sub-image (parity of r - 1, parity of s - 1) at (p + dr, q + ds) with
// (dr, ds) in {-1, 0} -- a row / pixel SHIFT of the UMMA descriptor inside that sub-image, exactly like the stride-1 row
// kernels.  One M-tile = R = 4 output rows x Wp = 30 columns (28 + the left halo + 1); half of the 144 KB weight matrix stays
// resident per CTA.
#pragma once

struct S2Cfg {
  static constexpr int W = 28, R = 4, Wp = 30, BN = 128;
  static constexpr int kEE = 0;                          // even rows, even cols: tap (1,1)                 128 rows read
  static constexpr int kEO = 16384;                      // even rows, odd cols:  taps (1,0) (1,2)          129
  static constexpr int kOE = kEO + 17408;                // odd rows, even cols:  taps (0,1) (2,1)          128 + Wp
  static constexpr int kOO = kOE + 20480;                // odd rows, odd cols:   taps (0,0) (0,2) (2,0) (2,2)  129 + Wp
  static constexpr int kStageBytes = kOO + 20480;        // 73 KB
  static constexpr int kLoadBytes = (2 * R * Wp + 2 * (R + 1) * Wp) * 128;     // what TMA writes per stage and CTA
  static constexpr int kStages = 2;
  static constexpr int kBBlock = (BN / 2) * 128;         // one tap's [64 rows x 64 channels] block of this CTA's half
  static constexpr int kBBytes = 9 * kBBlock;            // 72 KB
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBBytes + 1024 + 512;
  static_assert(R * Wp <= 128 && kSmemBytes <= 232448, "tile / shared memory budget");
};

__global__ void __launch_bounds__(conv_threads(128), 1)
k_conv3x3s2_rows2(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmB,
                  const RowConvParams p) {
  using Cfg = S2Cfg;
  constexpr int W = Cfg::W, R = Cfg::R, Wp = Cfg::Wp, BN = Cfg::BN, TILES_PER_IMG = W / R, NS = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;
  uint8_t* sB = base + NS * Cfg::kStageBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + Cfg::kBBytes);    // leader
  uint64_t* a_empty = a_full + NS;                                       // local
  uint64_t* b_full = a_empty + NS;                                       // leader
  uint64_t* tfull = b_full + 1;                                          // local
  uint64_t* tempty = tfull + 2;                                          // leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  ptx::pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmE);
    ptx::prefetch_tensormap(&tmO);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < NS; s++) ptx::mbar_init(&a_full[s], 1), ptx::mbar_init(&a_empty[s], 1);
    ptx::mbar_init(b_full, 1);
    for (int a = 0; a < 2; a++) ptx::mbar_init(&tfull[a], 1), ptx::mbar_init(&tempty[a], 8);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {   // this CTA's half of the resident weights: constant data, loaded before the dependency wait
    if (ptx::elect_one()) {
      if (leader) ptx::mbar_arrive_expect_tx(b_full, 2 * Cfg::kBBytes);
      for (int tap = 0; tap < 9; tap++) ptx::tma2_load_2d(sB + tap * Cfg::kBBlock, &tmB, b_full, tap * 64, (int)rank * (BN / 2));
    }
    __syncwarp();
  }
  ptx::pdl_wait();
  const int num_tiles = effective_patches(p.n_dev, p.n_base, p.n_img) * (p.num_tiles / p.n_img);
  const int num_pairs = (num_tiles + 1) >> 1;
  const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
  auto tile_of = [&](int tp, bool& active) {
    int t = 2 * tp + (int)rank;
    active = t < num_tiles;
    if (!active) t = num_tiles - 1;
    return p.reverse ? num_tiles - 1 - t : t;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): the four parity sub-images of this CTA's tile =====================
    int sa = 0;
    uint32_t pa = 0;
    for (int tp = pair0; tp < num_pairs; tp += pair_step) {
      bool active;
      const int vt = tile_of(tp, active);
      const int img = vt / TILES_PER_IMG, p0 = (vt - img * TILES_PER_IMG) * R;
      ptx::mbar_wait(&a_empty[sa], pa ^ 1);
      if (ptx::elect_one()) {
        uint8_t* st = sA + sa * Cfg::kStageBytes;
        if (leader) ptx::mbar_arrive_expect_tx(&a_full[sa], 2 * Cfg::kLoadBytes);
        ptx::tma2_load_4d(st + Cfg::kEE, &tmE, &a_full[sa], 0, 0, 2 * p0, img);
        ptx::tma2_load_4d(st + Cfg::kEO, &tmE, &a_full[sa], 0, -1, 2 * p0, img);
        ptx::tma2_load_4d(st + Cfg::kOE, &tmO, &a_full[sa], 0, 0, 2 * p0 - 1, img);
        ptx::tma2_load_4d(st + Cfg::kOO, &tmO, &a_full[sa], 0, -1, 2 * p0 - 1, img);
      }
      __syncwarp();
      if (++sa == NS) sa = 0, pa ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(256, BN);
      int sa = 0;
      uint32_t pa = 0, acc = 0, acc_phase = 0;
      ptx::mbar_wait(b_full, 0);
      ptx::tc_fence_after();
      const uint64_t b_all = ptx::make_smem_desc(ptx::smem_u32(sB), 128);
      for (int tp = pair0; tp < num_pairs; tp += pair_step) {
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        ptx::mbar_wait(&a_full[sa], pa);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(sA + sa * Cfg::kStageBytes);
        if (ptx::elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            constexpr int kSub[2][2] = {{Cfg::kOO, Cfg::kOE}, {Cfg::kEO, Cfg::kEE}};   // [row even][col even]
            const int r = tap / 3, s = tap % 3;
            const int shift = (r == 2 ? Wp : 0) + (s == 2 ? 1 : 0);                    // rows of 128 bytes inside the sub-image
            const uint64_t adesc = ptx::make_smem_desc(st + kSub[r == 1][s == 1], 128) + (uint64_t)(shift * 8);
            const uint64_t bdesc = b_all + (uint64_t)(tap * (Cfg::kBBlock >> 4));
#pragma unroll
            for (int k = 0; k < 4; k++) ptx::umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
          }
          ptx::umma2_commit_mc(&a_empty[sa]);
          ptx::umma2_commit_mc(&tfull[acc]);
        }
        __syncwarp();
        if (++sa == NS) sa = 0, pa ^= 1;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 accumulator rows) =====================
    const int wq = warp & 3;
    const int pos = wq * 32 + lane;
    const int rr = pos / Wp, x = pos - rr * Wp;
    const bool in_tile = rr < R && x < W;
    int it = 0;
    for (int tp = pair0; tp < num_pairs; tp += pair_step, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      bool active;
      const int vt = tile_of(tp, active);
      const int img = vt / TILES_PER_IMG, p0 = (vt - img * TILES_PER_IMG) * R;
      const size_t pix = ((size_t)img * W + p0 + rr) * W + x;
      const bool valid = in_tile && active;
      epilogue_row<BN>(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BN, p.bias, nullptr, p.out + pix * BN, p.relu, valid, &tfull[acc],
                       acc_phase);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&tempty[acc]);
        else ptx::mbar_arrive_cluster(&tempty[acc], 0);
      }
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}
