// 64-channel row-tile 3x3 convolution whose EPILOGUE moves through shared memory and TMA in both directions (included by
// resnet18.cu inside namespace hipac after conv_rows.cuh; same producer / MMA warps, same tile geometry).
//
// In k_conv3x3_rows every epilogue thread owns one output pixel and stores its 64 channels (128 bytes) straight to global
// memory, and fetches the 128 residual bytes the same way: a warp instruction touches 32 different 128-byte lines, i.e. 32
// LSU wavefronts, 512 wavefronts per warp and tile with a residual -- more cycles on the SM's load/store path than the
// tile's 36 UMMAs take.  Here the residual tile (R x W pixels x 64 channels = 14 KB) arrives by ONE TMA load into a
// 128-byte-swizzled staging buffer, every thread reads / rewrites only its own 128-byte row of that buffer (16-byte
// accesses, conflict free: 4 wavefronts per instruction) and the finished tile leaves by ONE TMA store.  Two staging
// buffers per epilogue warpgroup: the store of tile j and the residual load of tile j + 1 overlap the math of tile j.
#pragma once

template <int KC, int W, int R>
struct RowTmaCfg {
  static constexpr int BN = 64;
  static constexpr int Wp = W + 2;
  static constexpr int kRegionRows = 128 + 2 * Wp + 2;
  static constexpr int kRegionBytes = (kRegionRows * 128 + 1023) / 1024 * 1024;
  static constexpr int kLoadBytes = (R + 2) * Wp * 128;
  static constexpr int kAStages = 3;
  static constexpr int kBBlock = BN * 128;
  static constexpr int kBBytes = 9 * KC * kBBlock;
  static constexpr int kTileBytes = R * W * 128;                                  // one output / residual tile
  static constexpr int kStageBytes = (kTileBytes + 1023) / 1024 * 1024;
  static constexpr int kSmemBytes = kAStages * kRegionBytes + kBBytes + 4 * kStageBytes + 1024 + 512;
  static_assert(R * Wp <= 128, "tile does not fit one UMMA M = 128");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// bias + residual (16-byte chunk of 8 bf16, already loaded) + ReLU on 8 accumulator columns -> 8 bf16
__device__ __forceinline__ uint4 epilogue_chunk8(const uint32_t* v, const float* __restrict__ bias, uint4 res, bool has_res, int relu) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias)), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
  float x[8] = {__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y, __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w,
                __uint_as_float(v[4]) + b1.x, __uint_as_float(v[5]) + b1.y, __uint_as_float(v[6]) + b1.z, __uint_as_float(v[7]) + b1.w};
  if (has_res) {
    const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&res);
#pragma unroll
    for (int i = 0; i < 4; i++) x[2 * i] += __bfloat162float(r2[i].x), x[2 * i + 1] += __bfloat162float(r2[i].y);
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fmaxf(x[i], 0.f);
  }
  uint4 o;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(x[0], x[1]), o.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(x[2], x[3]), o.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(x[4], x[5]), o.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(x[6], x[7]), o.w = *reinterpret_cast<uint32_t*>(&t);
  return o;
}

template <int KC, int W, int R>
__global__ void __launch_bounds__(conv_threads(64), 1)
k_conv3x3_rows_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const RowConvParams p) {
  using Cfg = RowTmaCfg<KC, W, R>;
  constexpr int BN = 64, Wp = Cfg::Wp, H = W, TILES_PER_IMG = H / R;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;
  uint8_t* sB = base + Cfg::kAStages * Cfg::kRegionBytes;
  uint8_t* sO = sB + Cfg::kBBytes;                                   // [group][buffer] staging tiles
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sO + 4 * Cfg::kStageBytes);
  uint64_t* a_empty = a_full + Cfg::kAStages;
  uint64_t* b_full = a_empty + Cfg::kAStages;
  uint64_t* tfull = b_full + 1;
  uint64_t* tempty = tfull + 2;
  uint64_t* st_ready = tempty + 2;                                   // [group][buffer]: staging buffer free (+ residual landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_ready + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    ptx::prefetch_tensormap(&tmO);
    if (p.residual) ptx::prefetch_tensormap(&tmR);
    for (int s = 0; s < Cfg::kAStages; s++) ptx::mbar_init(&a_full[s], 1), ptx::mbar_init(&a_empty[s], 1);
    ptx::mbar_init(b_full, 1);
    for (int a = 0; a < 2; a++) ptx::mbar_init(&tfull[a], 1), ptx::mbar_init(&tempty[a], 4);
    for (int s = 0; s < 4; s++) ptx::mbar_init(&st_ready[s], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {   // resident weights: constant data, loaded before the dependency wait
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(b_full, Cfg::kBBytes);
      for (int kb = 0; kb < 9 * KC; kb++) ptx::tma_load_2d(sB + kb * Cfg::kBBlock, &tmB, b_full, kb * 64, 0);
    }
    __syncwarp();
  }
  ptx::pdl_wait();
  const int num_tiles = effective_patches(p.n_dev, p.n_base, p.n_img) * (p.num_tiles / p.n_img);

  if (warp == 0) {
    // ===================== TMA producer =====================
    int sa = 0;
    uint32_t pa = 0;
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
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN);
    int sa = 0;
    uint32_t pa = 0, acc = 0, acc_phase = 0;
    ptx::mbar_wait(b_full, 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kc = 0; kc < KC; kc++) {
        ptx::mbar_wait(&a_full[sa], pa);
        ptx::tc_fence_after();
        const uint64_t a_region = ptx::make_smem_desc(ptx::smem_u32(sA + sa * Cfg::kRegionBytes), 128);
        const uint64_t b_all = ptx::make_smem_desc(ptx::smem_u32(sB + kc * Cfg::kBBlock), 128);
        if (ptx::elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const uint64_t adesc = a_region + (uint64_t)(((tap / 3) * Wp + tap % 3) * 8);
            const uint64_t bdesc = b_all + (uint64_t)(tap * KC * (Cfg::kBBlock >> 4));
#pragma unroll
            for (int k = 0; k < 4; k++) ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&a_empty[sa]);
        }
        __syncwarp();
        if (++sa == Cfg::kAStages) sa = 0, pa ^= 1;
      }
      if (ptx::elect_one()) ptx::umma_commit(&tfull[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue: TMEM -> staging tile (residual in place) -> TMA store =====================
    const int wq = warp & 3;
    const int grp = (warp - 2) >> 2;                     // warpgroup g takes this CTA's tiles g, g + 2, g + 4, ...
    const int gt = threadIdx.x - 64 - 128 * grp;         // 0..127 inside the warpgroup
    const int pos = wq * 32 + lane;
    const int rr = pos / Wp, x = pos - rr * Wp;
    const bool valid = rr < R && x < W;
    const int srow = rr * W + x;                         // row of the staging tile ([R][W] pixels of 128 bytes)
    const bool has_res = p.residual != nullptr;
    uint8_t* stage0 = sO + grp * 2 * Cfg::kStageBytes;
    uint64_t* ready = st_ready + grp * 2;
    const int first = blockIdx.x + grp * gridDim.x, step = 2 * gridDim.x;
    // thread 0 of the warpgroup owns the bulk groups (stores) and prepares staging buffers: buffer b is free once the
    // store that last used it has finished reading it; with a residual the free buffer is then filled by a TMA load
    auto prepare = [&](int tile, int b) {
      if (has_res) {
        const int vt = p.reverse ? num_tiles - 1 - tile : tile;
        const int img = vt / TILES_PER_IMG, p0 = (vt - img * TILES_PER_IMG) * R;
        ptx::mbar_arrive_expect_tx(&ready[b], Cfg::kTileBytes);
        ptx::tma_load_4d(stage0 + b * Cfg::kStageBytes, &tmR, &ready[b], 0, 0, p0, img);
      } else {
        ptx::mbar_arrive(&ready[b]);
      }
    };
    if (gt == 0 && first < num_tiles) prepare(first, 0);
    int j = 0;
    for (int tile = first; tile < num_tiles; tile += step, ++j) {
      const int b = j & 1;
      const uint32_t acc = grp, acc_phase = j & 1;       // CTA-local tile index 2j + grp: accumulator grp, phase toggles per use
      uint8_t* buf = stage0 + b * Cfg::kStageBytes;
      ptx::mbar_wait(&ready[b], (j >> 1) & 1);
      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BN;
      ptx::tmem_ld_32x32b_x32(trow, v0);
      ptx::tmem_ld_32x32b_x32(trow + 32, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);     // accumulator drained: the MMA warp may start the tile after next
      if (valid) {
        uint4* row = reinterpret_cast<uint4*>(buf + srow * 128);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int ph = i ^ (srow & 7);                 // 128-byte swizzle: 16-byte chunk i of row r sits at chunk i ^ (r & 7)
          uint4 res = make_uint4(0u, 0u, 0u, 0u);
          if (has_res) res = row[ph];
          row[ph] = epilogue_chunk8(i < 4 ? &v0[8 * i] : &v1[8 * (i - 4)], p.bias + 8 * i, res, has_res, p.relu);
        }
      }
      ptx::fence_proxy_async();                          // generic-proxy writes of the tile -> visible to the TMA store
      ptx::named_bar_sync(1 + grp, 128);
      if (gt == 0) {
        const int vt = p.reverse ? num_tiles - 1 - tile : tile;
        const int img = vt / TILES_PER_IMG, p0 = (vt - img * TILES_PER_IMG) * R;
        ptx::tma_store_4d(&tmO, buf, 0, 0, p0, img);
        ptx::bulk_commit_group();
        if (tile + step < num_tiles) {
          ptx::bulk_wait_group_read<1>();                // the store of tile j - 1 (other buffer) has read its source
          prepare(tile + step, b ^ 1);
        }
      }
    }
    if (gt == 0) ptx::bulk_wait_group<0>();              // all stores complete before the CTA's shared memory goes away
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 2 * BN);
}
