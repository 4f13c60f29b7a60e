// Streaming stage-1 pass (sm_100a): ONE read of the level image produces both the per-cell byte sums of the tissue
// test and the globally Pillow-resampled level D with its ring-variant planes (see tile_scan_fused.cuh for why the
// planes are exact).  Included by tile_scan.cu inside namespace hipac, after tile_scan_fused.cuh.
//
// Work decomposition: a WARP is the unit of work, nothing is shared between warps.
//   * the image is cut into column strips of 31 "units" (unit = 8 pixels = 24 bytes) and row chunks; one warp owns one
//     (strip, chunk) item, lane l owns unit l of the strip and walks DOWN the rows, so both Pillow passes and the
//     vertical accumulators live in that lane's registers -- no shared-memory intermediate, no block barrier;
//   * rows arrive through a per-warp ring of TMA bulk copies (cp.async.bulk, one 768..800-byte copy per image row,
//     completion on a per-stage mbarrier), 8 rows per stage;
//   * a lane de-interleaves its 24 bytes with PRMT into per-channel words and evaluates the triangle filter with dp4a.
//     Pillow's interior window for scale F is 2F taps with weights (2m+1)/2F^2 ascending then descending, i.e. two
//     adjacent F-pixel blocks: out[I] = U(I) + V(I+1), U = sum (2m+1) x_m over block I, V = 2F*sum(x) - U over block
//     I+1.  Every block is therefore touched once (U and sum), the neighbour's V comes from lane+1 by shuffle (the 32nd
//     lane of a strip only supplies V), and the same decomposition runs vertically on the uint8 intermediate;
//   * the block sums are exactly the byte sums the tissue test needs, so the cell statistics cost six dp4a per row;
//   * ring-variant COLUMNS (clamped 3F/2-tap windows at patch edges, <= 10 per strip) are computed by up to 30 lanes as
//     a fourth "channel" from the same staged rows; ring-variant ROWS are extra accumulators that are only live in the
//     row blocks around the patch-edge lattice.
//
// Applicability (else the cp.async kernels of tile_scan_fused.cuh run): pitch % 16 == 0, 16-byte aligned image, and at
// most kStreamMaxVar ring columns per strip.
#pragma once
#include <type_traits>

constexpr int kStripUnits = 31;        // units per strip that produce outputs
constexpr int kStreamRows = 8;         // image rows per pipeline stage
constexpr int kStreamRowBytes = 800;   // staged row: 768 bytes + alignment phase + 64-bit over-read
constexpr int kStreamMaxVar = 10;      // ring-variant columns per strip (3 channels each -> 30 lanes)
#ifndef HIPAC_STREAM_STAGES
#define HIPAC_STREAM_STAGES 2
#endif
#ifndef HIPAC_STREAM_WARPS
#define HIPAC_STREAM_WARPS 8
#endif
constexpr int kStreamStages = HIPAC_STREAM_STAGES;
constexpr int kStreamWarps = HIPAC_STREAM_WARPS;   // warps (independent items) per CTA

struct StreamGeom {
  int n_strips, n_chunks;   // items = n_strips * n_chunks, item id = chunk * n_strips + strip
  int srow_lo, srow_hi;     // image rows whose bytes are counted in the cell sums
};

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

// ---- compile-time byte plumbing ---------------------------------------------------------------------------------
// The lane's 8 loaded words hold its 24 unit bytes at byte positions OFF .. OFF+23.  gather_stride3<G0> collects the
// bytes at positions G0, G0+3, G0+6, G0+9 (four consecutive pixels of one channel) into one register with 2-3 PRMTs.
__host__ __device__ constexpr uint32_t prmt_nib(int g, int base, int keep_below, int i) {
  // selector nibble for output byte i when merging (acc, w[base]): bytes from words < base were placed earlier (keep),
  // bytes of word `base` come from the second operand
  return (g >> 2) == base ? (uint32_t)(4 + (g & 3)) : (((g >> 2) < base && keep_below) ? (uint32_t)i : 0u);
}
template <int G0>
__device__ __forceinline__ uint32_t gather_stride3(const uint32_t (&w)[8]) {
  constexpr int g0 = G0, g1 = G0 + 3, g2 = G0 + 6, g3 = G0 + 9;
  constexpr int wa = g0 >> 2, wl = g3 >> 2;
  static_assert(wl < 8 && wl - wa >= 2 && wl - wa <= 3, "unexpected span");
  // step 1: (w[wa], w[wa+1]) -- bytes of w[wa] use selectors 0..3
  constexpr uint32_t n0 = (g0 >> 2) == wa ? (g0 & 3) : ((g0 >> 2) == wa + 1 ? 4 + (g0 & 3) : 0);
  constexpr uint32_t n1 = (g1 >> 2) == wa ? (g1 & 3) : ((g1 >> 2) == wa + 1 ? 4 + (g1 & 3) : 0);
  constexpr uint32_t n2 = (g2 >> 2) == wa ? (g2 & 3) : ((g2 >> 2) == wa + 1 ? 4 + (g2 & 3) : 0);
  constexpr uint32_t n3 = (g3 >> 2) == wa ? (g3 & 3) : ((g3 >> 2) == wa + 1 ? 4 + (g3 & 3) : 0);
  uint32_t acc = __byte_perm(w[wa], w[wa + 1], n0 | (n1 << 4) | (n2 << 8) | (n3 << 12));
  constexpr uint32_t s2 = prmt_nib(g0, wa + 2, 1, 0) | (prmt_nib(g1, wa + 2, 1, 1) << 4) | (prmt_nib(g2, wa + 2, 1, 2) << 8) |
                          (prmt_nib(g3, wa + 2, 1, 3) << 12);
  acc = __byte_perm(acc, w[wa + 2], s2);
  if constexpr (wl == wa + 3) {
    constexpr uint32_t s3 = prmt_nib(g0, wa + 3, 1, 0) | (prmt_nib(g1, wa + 3, 1, 1) << 4) | (prmt_nib(g2, wa + 3, 1, 2) << 8) |
                            (prmt_nib(g3, wa + 3, 1, 3) << 12);
    acc = __byte_perm(acc, w[wa + 3], s3);
  }
  return acc;
}

// dp4a weights of block b (pixels bF .. bF+F-1 of the unit) inside channel word j (pixels 4j .. 4j+3)
template <int F>
__host__ __device__ constexpr uint32_t wU(int b, int j) {
  uint32_t w = 0;
  for (int i = 0; i < 4; i++) {
    const int p = 4 * j + i;
    if (p >= b * F && p < b * F + F) w |= (uint32_t)(2 * (p - b * F) + 1) << (8 * i);
  }
  return w;
}
template <int F>
__host__ __device__ constexpr uint32_t wS(int b, int j) {
  uint32_t w = 0;
  for (int i = 0; i < 4; i++) {
    const int p = 4 * j + i;
    if (p >= b * F && p < b * F + F) w |= 1u << (8 * i);
  }
  return w;
}
// dp4a masks of raw loaded word j for the cell sums: bytes of the first F/2 pixels of the unit / of the rest
template <int F, int OFF>
__host__ __device__ constexpr uint32_t wFirst(int j) {
  uint32_t w = 0;
  for (int i = 0; i < 4; i++) {
    const int g = 4 * j + i;
    if (g >= OFF && g < OFF + 3 * (F / 2)) w |= 1u << (8 * i);
  }
  return w;
}
template <int F, int OFF>
__host__ __device__ constexpr uint32_t wRest(int j) {
  uint32_t w = 0;
  for (int i = 0; i < 4; i++) {
    const int g = 4 * j + i;
    if (g >= OFF + 3 * (F / 2) && g < OFF + 24) w |= 1u << (8 * i);
  }
  return w;
}

template <int F>
__global__ void __launch_bounds__(kStreamWarps * 32) k_scan_planes(ScanParams p, FusedGeom G, StreamGeom Z) {
  constexpr int NB = 8 / F, HALF = F / 2, NE = 3 * F / 2, NH = 3 * NB + 1, NS = kStreamStages;
  constexpr int SHIFT = F == 8 ? 7 : (F == 4 ? 5 : 3);   // interior weights are (2m+1) / 2^SHIFT exactly
  constexpr int OFF = (64 - 3 * F / 2) % 8;               // (byte offset of a unit) mod 8: 4, 2, 5 for F = 8, 4, 2
  constexpr int RND = 1 << (SHIFT - 1);
  constexpr int VRND = 1 << (kPrecisionBits - 1);
  extern __shared__ __align__(128) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int item = blockIdx.x * kStreamWarps + warp;
  if (item >= Z.n_strips * Z.n_chunks) return;
  const int chunk = item / Z.n_strips, strip = item - chunk * Z.n_strips;
  uint8_t* ring = sm + (size_t)warp * NS * kStreamRows * kStreamRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)kStreamWarps * NS * kStreamRows * kStreamRowBytes) + warp * NS;
  const CoeffSet& cs = c_coef[G.lf];

  // ---- rows of this item: D rows [Ja, Jb) need row blocks Ja .. Jb (block J = image rows F*J - F/2 .. F*J + F/2 - 1) ----
  const int Ja = G.Jbase + (int)((int64_t)G.Dh * chunk / Z.n_chunks);
  const int Jb = G.Jbase + (int)((int64_t)G.Dh * (chunk + 1) / Z.n_chunks);
  if (Jb <= Ja) return;
  const int nblocks = Jb - Ja + 1;
  const int r_item = F * Ja - HALF;
  const int nrows = nblocks * F;
  const int nstages = (nrows + kStreamRows - 1) / kStreamRows;
  const bool last_chunk = chunk == Z.n_chunks - 1;
  const int s_lo = max(Z.srow_lo, r_item);
  const int s_hi = max(s_lo, last_chunk ? Z.srow_hi : min(Z.srow_hi, F * Jb - HALF));

  // ---- columns of this strip ----
  const int u0 = strip * kStripUnits;               // first unit; unit u = pixels 8u - F/2 .. 8u - F/2 + 7
  const int pix0 = 8 * u0 - HALF;
  const int a0 = 3 * pix0;                          // byte offset of the strip inside a row (may be < 0 for strip 0)
  const int A0 = a0 & ~15;                          // 16-byte aligned start of the staged window (floor also for a0 < 0)
  const int ph = a0 - A0;
  const int cbeg = max(A0, 0);
  const int row_lim = (int)min((int64_t)((p.W * 3 + 15) & ~15), p.pitch);
  const int cend = min((a0 + 768 + 15) & ~15, row_lim);
  const int copy_bytes = max(cend - cbeg, 0);       // 0: the whole strip lies right of the image (all white)
  const bool edge_cols = a0 + 768 > p.W * 3;
  const int fill_from = max(p.W * 3 - A0, 0);       // staged-row offset of the first byte right of the image
  const int lane_off = ph - OFF + 24 * lane;        // multiple of 8
  const int I_first = NB * (u0 + lane);             // first D column of this lane

  // ---- ring-variant column task of this lane: the (lane/3)-th variant column of the strip, channel lane%3 ----
  int v_kind = 0, v_I = 0;
  {
    int n = 0;
    const int want = lane / 3;
    const int Iend = min(NB * (u0 + kStripUnits), G.Dw);
    for (int I = NB * u0; I < Iend; I++) {
      if (I % G.Sf == 0 && I / G.Sf < G.nx) {
        if (n == want) v_kind = 1, v_I = I;
        n++;
      }
      if (I >= OUT - 1 && (I - (OUT - 1)) % G.Sf == 0 && (I - (OUT - 1)) / G.Sf < G.nx) {
        if (n == want) v_kind = 2, v_I = I;
        n++;
      }
    }
  }
  const bool has_task = v_kind != 0;
  const int v_c = lane % 3;
  const int v_ix = v_kind == 1 ? v_I / G.Sf : (v_I - (OUT - 1)) / G.Sf;
  const int v_off = ph + 3 * ((v_kind == 1 ? F * v_I : F * v_I - HALF) - pix0) + v_c;   // byte offset inside a staged row
  int kh[NE];
#pragma unroll
  for (int t = 0; t < NE; t++) kh[t] = v_kind == 1 ? cs.left[t] : cs.right[t];

  // ---- output pointers (row index added per emit) ----
  uint8_t* o_int = G.plane[0][0] + (size_t)I_first * 3;
  uint8_t* o_top = G.plane[1][0] + (size_t)I_first * 3;
  uint8_t* o_bot = G.plane[2][0] + (size_t)I_first * 3;
  uint8_t* ov_int = G.plane[0][has_task ? v_kind : 1] + (size_t)v_ix * 3 + v_c;
  uint8_t* ov_top = G.plane[1][has_task ? v_kind : 1] + (size_t)v_ix * 3 + v_c;
  uint8_t* ov_bot = G.plane[2][has_task ? v_kind : 1] + (size_t)v_ix * 3 + v_c;
  const size_t rs_int = (size_t)G.Dw * 3, rs_var = (size_t)G.nx * 3;
  bool st_ok[NH];
#pragma unroll
  for (int e = 0; e < 3 * NB; e++) st_ok[e] = lane < kStripUnits && I_first + e / 3 < G.Dw;
  st_ok[3 * NB] = has_task;

  // ---- cell-sum bookkeeping ----
  const int u = u0 + lane;
  int cx_f = -1, cx_r = -1;
  uint32_t corr_f = 0, corr_r = 0;
  if (lane < kStripUnits) {
    if (u > 0 && 8 * u - HALF < p.W) cx_f = (8 * u - 1) / G.g;
    if (8 * u < p.W) cx_r = 8 * u / G.g;
    const int nf = min(max(8 * u - p.W, 0), HALF);                    // pixels of [8u - F/2, 8u) at or beyond W
    const int nr = min(max(8 * u + 8 - HALF - p.W, 0), 8 - HALF);     // pixels of [8u, 8u + 8 - F/2) at or beyond W
    corr_f = 255u * 3u * (uint32_t)(u > 0 ? nf : 0);
    corr_r = 255u * 3u * (uint32_t)nr;
  }
  uint32_t acc_f = 0, acc_r = 0;
  auto flush_cells = [&](int cy_global) {
    const int cyl = cy_global - G.cy0;
    if (cyl >= 0 && cyl < G.ncy) {
      const int cxa = (8 * u0 - (u0 > 0 ? 1 : 0)) / G.g;
      const int cxb = min((8 * (u0 + kStripUnits - 1)) / G.g, G.ncx - 1);
      for (int cx = cxa; cx <= cxb; cx++) {
        uint32_t v = (cx_f == cx ? acc_f : 0u) + (cx_r == cx ? acc_r : 0u);
        v = __reduce_add_sync(0xffffffffu, v);
        if (lane == 0 && v) atomicAdd(&G.cell_sum[(size_t)cyl * G.ncx + cx], v);
      }
    }
    acc_f = acc_r = 0;
  };

  // ---- pipeline ----
  if (lane == 0) {
    for (int s = 0; s < NS; s++) ptx::mbar_init(&bars[s], 1);
    ptx::fence_barrier_init();
  }
  if (a0 < 0) {   // strip 0: the bytes left of pixel 0 are never written by the copies
    for (int k = lane; k < NS * kStreamRows; k += 32) *reinterpret_cast<uint4*>(ring + k * kStreamRowBytes) = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  auto issue = [&](int st) {
    uint8_t* buf = ring + (size_t)(st % NS) * kStreamRows * kStreamRowBytes;
    const int r = r_item + st * kStreamRows + lane;
    const bool valid = lane < kStreamRows && st * kStreamRows + lane < nrows && r >= 0 && r < p.H && copy_bytes > 0;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    ptx::fence_proxy_async();   // order this lane's generic-proxy fills of the buffer before the async-proxy writes
    if (lane == 0) {
      if (m) ptx::mbar_arrive_expect_tx(&bars[st % NS], (uint32_t)(__popc(m) * copy_bytes));
      else ptx::mbar_arrive(&bars[st % NS]);
    }
    __syncwarp();
    if (valid) {
      HIPAC_DEV_ASSERT(cbeg >= A0 && (cbeg - A0) + copy_bytes <= kStreamRowBytes && (copy_bytes & 15) == 0);
      HIPAC_DEV_ASSERT((int64_t)r * p.pitch + cbeg + copy_bytes <= (int64_t)p.H * p.pitch);
      bulk_copy_g2s(buf + lane * kStreamRowBytes + (cbeg - A0), p.rgb + (int64_t)r * p.pitch + cbeg, (uint32_t)copy_bytes, &bars[st % NS]);
    }
  };

  HIPAC_DEV_ASSERT(lane_off >= 0 && lane_off + 32 <= kStreamRowBytes && (lane_off & 7) == 0);
  HIPAC_DEV_ASSERT(!has_task || (v_off >= 0 && v_off + 3 * NE <= kStreamRowBytes));
  int Uv[NH], Sv[NH], Up[NH], Tt[NH], Bt[NH];
#pragma unroll
  for (int e = 0; e < NH; e++) Uv[e] = Sv[e] = Up[e] = Tt[e] = Bt[e] = 0;

  for (int st = 0; st < NS - 1 && st < nstages; st++) issue(st);
  for (int st = 0; st < nstages; st++) {
    __syncwarp();
    if (st + NS - 1 < nstages) issue(st + NS - 1);
    ptx::mbar_wait(&bars[st % NS], (uint32_t)((st / NS) & 1));
    uint8_t* buf = ring + (size_t)(st % NS) * kStreamRows * kStreamRowBytes;
    const int r_stage = r_item + st * kStreamRows;
    const int rows_here = min(kStreamRows, nrows - st * kStreamRows);
    // white fill: rows outside the image, and the bytes right of the image edge (warp-uniform tests)
    if (r_stage < 0 || r_stage + rows_here > p.H || edge_cols) {
      for (int rr = 0; rr < rows_here; rr++) {
        const int r = r_stage + rr;
        uint8_t* rowp = buf + rr * kStreamRowBytes;
        if (r < 0 || r >= p.H) {
          const uint32_t fill = r < 0 ? 0u : 0xFFFFFFFFu;
          for (int k = lane; k < kStreamRowBytes / 16; k += 32) *reinterpret_cast<uint4*>(rowp + k * 16) = make_uint4(fill, fill, fill, fill);
        } else if (edge_cols) {
          for (int k = fill_from + lane; k < kStreamRowBytes; k += 32) rowp[k] = 255;
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int b = 0; b < kStreamRows / F; b++) {
      const int q = st * (kStreamRows / F) + b;
      if (q >= nblocks) break;
      const int J = Ja + q;
      // ring-variant rows: top variant of lattice row Jt uses the second half of block Jt and all of block Jt+1;
      // bottom variant of Jb' uses all of block Jb' and the first half of block Jb'+1
      auto is_top = [&](int j) { return j >= Ja && j % G.Sf == 0 && (unsigned)(j / G.Sf - G.iy_begin) < (unsigned)G.ny; };
      auto is_bot = [&](int j) {
        return j >= Ja && j >= OUT - 1 && (j - (OUT - 1)) % G.Sf == 0 && (unsigned)((j - (OUT - 1)) / G.Sf - G.iy_begin) < (unsigned)G.ny;
      };
      const bool tA = is_top(J), tB = is_top(J - 1), bA = is_bot(J), bB = is_bot(J - 1);
      const bool var = tA || tB || bA || bB;
      // the eight (F) rows of the block, instantiated with and without the ring-variant row accumulators
      auto rows_of_block = [&](auto var_tag) {
        constexpr bool VAR = decltype(var_tag)::value;
#pragma unroll
      for (int m = 0; m < F; m++) {
        const int rr = b * F + m;
        const int r = r_stage + rr;
        const uint8_t* rowp = buf + rr * kStreamRowBytes;
        // ---- horizontal pass of this lane's unit ----
        uint32_t w[8];
        {
          const uint2 q0 = *reinterpret_cast<const uint2*>(rowp + lane_off);
          const uint2 q1 = *reinterpret_cast<const uint2*>(rowp + lane_off + 8);
          const uint2 q2 = *reinterpret_cast<const uint2*>(rowp + lane_off + 16);
          const uint2 q3 = *reinterpret_cast<const uint2*>(rowp + lane_off + 24);
          w[0] = q0.x, w[1] = q0.y, w[2] = q1.x, w[3] = q1.y, w[4] = q2.x, w[5] = q2.y, w[6] = q3.x, w[7] = q3.y;
        }
        int h[NH];
        {
          int Ub[3][NB], Vb[3][NB];
          const uint32_t c00 = gather_stride3<OFF + 0>(w), c01 = gather_stride3<OFF + 12>(w);
          const uint32_t c10 = gather_stride3<OFF + 1>(w), c11 = gather_stride3<OFF + 13>(w);
          const uint32_t c20 = gather_stride3<OFF + 2>(w), c21 = gather_stride3<OFF + 14>(w);
          const uint32_t cw[3][2] = {{c00, c01}, {c10, c11}, {c20, c21}};
#pragma unroll
          for (int c = 0; c < 3; c++)
#pragma unroll
            for (int bb = 0; bb < NB; bb++) {
              uint32_t U = 0, S = 0;
#pragma unroll
              for (int j = 0; j < 2; j++) {
                if (wU<F>(bb, j)) U = __dp4a(cw[c][j], wU<F>(bb, j), U);
                if (wS<F>(bb, j)) S = __dp4a(cw[c][j], wS<F>(bb, j), S);
              }
              Ub[c][bb] = (int)U;
              Vb[c][bb] = (int)(2 * F * S - U);
            }
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const int vnext = __shfl_down_sync(0xffffffffu, Vb[c][0], 1);
#pragma unroll
            for (int bb = 0; bb < NB; bb++) {
              const int T = Ub[c][bb] + (bb + 1 < NB ? Vb[c][bb + 1] : vnext);
              h[bb * 3 + c] = (T + RND) >> SHIFT;
            }
          }
        }
        // ---- cell sums (rows of the image that this item owns) ----
        {
          uint32_t f = 0, g2 = 0;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (wFirst<F, OFF>(j)) f = __dp4a(w[j], wFirst<F, OFF>(j), f);
            if (wRest<F, OFF>(j)) g2 = __dp4a(w[j], wRest<F, OFF>(j), g2);
          }
          const bool counted = (unsigned)(r - s_lo) < (unsigned)(s_hi - s_lo);   // branch-free: s_hi >= s_lo
          acc_f += counted ? f - corr_f : 0u;
          acc_r += counted ? g2 - corr_r : 0u;
        }
        // ---- ring-variant column (clamped 3F/2-tap window, 22-bit weights) ----
        h[3 * NB] = 0;
        if (has_task) {
          const uint8_t* pix = rowp + v_off;
          int a = VRND;
#pragma unroll
          for (int t = 0; t < NE; t++) a += kh[t] * (int)pix[3 * t];
          h[3 * NB] = min(max(a >> kPrecisionBits, 0), 255);
        }
        // ---- vertical pass: this is row m of block J ----
#pragma unroll
        for (int e = 0; e < NH; e++) {
          Uv[e] += (2 * m + 1) * h[e];
          Sv[e] += h[e];
        }
        if constexpr (VAR) {
#pragma unroll
          for (int e = 0; e < NH; e++) {
            if (tA && m >= HALF) Tt[e] += cs.left[m - HALF] * h[e];
            if (tB) Tt[e] += cs.left[m + HALF] * h[e];
            if (bA) Bt[e] += cs.right[m] * h[e];
            if (bB && m < HALF) Bt[e] += cs.right[m + F] * h[e];
          }
          if (bB && m == HALF - 1) {   // bottom-variant row J-1 is complete
            const size_t iy = (size_t)((J - 1 - (OUT - 1)) / G.Sf - G.iy_begin);
#pragma unroll
            for (int e = 0; e < NH; e++) {
              const uint8_t v = (uint8_t)min(max((Bt[e] + VRND) >> kPrecisionBits, 0), 255);
              if (e < 3 * NB) {
                if (st_ok[e]) o_bot[iy * rs_int + e] = v;
              } else if (st_ok[e]) {
                ov_bot[iy * rs_var] = v;
              }
              Bt[e] = 0;
            }
          }
        }
        if (m == HALF - 1 && (F * J) % G.g == 0 && F * J > s_lo) flush_cells(F * J / G.g - 1);
      }
      };
      if (var) rows_of_block(std::true_type{});
      else rows_of_block(std::false_type{});
      // ---- end of block J: D row J-1 = U(J-1) + V(J) ----
      if (q >= 1) {
        const size_t jr = (size_t)(J - 1 - G.Jbase);
        HIPAC_DEV_ASSERT(J - 1 >= G.Jbase && jr < (size_t)G.Dh);
#pragma unroll
        for (int e = 0; e < NH; e++) {
          const uint8_t v = (uint8_t)((Up[e] + 2 * F * Sv[e] - Uv[e] + RND) >> SHIFT);
          if (e < 3 * NB) {
            if (st_ok[e]) o_int[jr * rs_int + e] = v;
          } else if (st_ok[e]) {
            ov_int[jr * rs_var] = v;
          }
        }
        if (tB) {   // top-variant row J-1 is complete
          const size_t iy = (size_t)((J - 1) / G.Sf - G.iy_begin);
#pragma unroll
          for (int e = 0; e < NH; e++) {
            const uint8_t v = (uint8_t)min(max((Tt[e] + VRND) >> kPrecisionBits, 0), 255);
            if (e < 3 * NB) {
              if (st_ok[e]) o_top[iy * rs_int + e] = v;
            } else if (st_ok[e]) {
              ov_top[iy * rs_var] = v;
            }
          }
        }
      }
      if (tB) {
#pragma unroll
        for (int e = 0; e < NH; e++) Tt[e] = 0;
      }
#pragma unroll
      for (int e = 0; e < NH; e++) Up[e] = Uv[e], Uv[e] = 0, Sv[e] = 0;
    }
  }
  if (s_hi > s_lo) flush_cells((s_hi - 1) / G.g);
}
