// The one exchange step of the HiPAC hot path: per-rank survivors (coordinates, labels, 512-d
// features, logits) -> one canonically ordered result on every rank.  The reference has no counterpart (its only
// multi-GPU code is nn.DataParallel around a model that never runs, src/main.py:839-842); the contract is that the
// N-rank result is array-equal to a single-rank run, i.e. rows in the reference's emission order: x outer, y inner
// (src/main.py:682-683).
//
// Data model: a SEGMENT is the output of one (tile scan -> ResNet18) pass over a contiguous range of candidate grid rows:
// its rows are already in emission order and successive segments cover ascending, disjoint y ranges (rank after rank,
// and row group after row group inside a rank).  Nothing here needs the host to know a survivor count:
//
//   hipac_exchange_pack    segment outputs (capacity-sized, count on the device) -> one packed byte matrix
//                          [1 + capacity][row_bytes]: a header row holding the count, then one row per survivor
//                          (x, y + y_offset | label | features | logits).  This matrix is the send buffer of ONE
//                          fixed-size all-gather (NCCL over NVLink; the header rides along, no count exchange).
//   hipac_exchange_merge   all segments of all ranks -> final arrays.  Because every segment is emission-ordered, the rows
//                          of grid column ix are a contiguous run inside each segment: k_exchange_index finds the run
//                          boundaries by binary search and takes one exclusive prefix sum over (column, segment);
//                          k_exchange_scatter then moves every row straight to its final position.  No sort, no host
//                          round trip; the total is left in device memory.
#include "common.cuh"

namespace hipac {

constexpr int kXchgHeadBytes = 16;   // x int32, y int32, label u8, 7 bytes padding

__host__ __device__ inline int xchg_row_bytes(int feat_dim, int num_classes) { return (kXchgHeadBytes + feat_dim * 4 + num_classes * 4 + 15) / 16 * 16; }

// one warp per row
__global__ void __launch_bounds__(256) k_exchange_pack(const int32_t* __restrict__ coords, const uint8_t* __restrict__ labels,
                                                       const float* __restrict__ feats, const float* __restrict__ logits,
                                                       int feat_dim, int num_classes, const int32_t* __restrict__ count, int capacity,
                                                       int y_offset, uint8_t* __restrict__ send, int row_bytes) {
  const int n = min(max(__ldg(count), 0), capacity);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    int4 h = make_int4(n, 0, 0, 0);
    *reinterpret_cast<int4*>(send) = h;
  }
  if (warp >= n) return;
  uint8_t* row = send + (size_t)(1 + warp) * row_bytes;
  if (lane == 0) {
    int4 h;
    h.x = coords[2 * warp], h.y = coords[2 * warp + 1] + y_offset, h.z = (int)labels[warp], h.w = 0;
    *reinterpret_cast<int4*>(row) = h;
  }
  if (feat_dim) {
    const uint4* f = reinterpret_cast<const uint4*>(feats + (size_t)warp * feat_dim);
    uint4* d = reinterpret_cast<uint4*>(row + kXchgHeadBytes);
    for (int k = lane; k < feat_dim / 4; k += 32) d[k] = __ldg(f + k);
  }
  if (logits) {
    float* dl = reinterpret_cast<float*>(row + kXchgHeadBytes + feat_dim * 4);
    for (int k = lane; k < num_classes; k += 32) dl[k] = logits[(size_t)warp * num_classes + k];
  }
}

// Single CTA.  colstart[s][ix] = first row of stored segment s whose x >= ix * stride (ix = nx: the segment's count);
// base[ix * nseg + order(s)] = final position of the first row of (column ix, segment s).
// y-order rank of stored segment s.  Natural layout (cyc_world = 0): segments are stored in ascending y order.  Block-cyclic
// sharding (cyc_world = W ranks, spr segments each, stored rank-major as an all-gather delivers them): local segment g of
// rank r is grid-row block g * W + r of the level.
__device__ __forceinline__ int seg_order(int s, int nseg, int cyc_world) {
  if (cyc_world <= 0) return s;
  const int spr = nseg / cyc_world;
  return (s % spr) * cyc_world + s / spr;
}

__global__ void __launch_bounds__(1024) k_exchange_index(const uint8_t* __restrict__ recv, int nseg, int seg_rows, int row_bytes,
                                                         int stride, int nx, int cyc_world, int32_t* __restrict__ colstart,
                                                         int32_t* __restrict__ base, int32_t* __restrict__ total, int out_capacity) {
  const size_t seg_bytes = (size_t)(1 + seg_rows) * row_bytes;
  for (int t = threadIdx.x; t < nseg * (nx + 1); t += blockDim.x) {
    const int s = t / (nx + 1), ix = t - s * (nx + 1);
    const uint8_t* seg = recv + (size_t)s * seg_bytes;
    const int cnt = min(max(*reinterpret_cast<const int32_t*>(seg), 0), seg_rows);
    int lo = 0, hi = cnt;
    if (ix < nx) {
      const int want = ix * stride;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (*reinterpret_cast<const int32_t*>(seg + (size_t)(1 + mid) * row_bytes) < want) lo = mid + 1;
        else hi = mid;
      }
    } else {
      lo = cnt;
    }
    colstart[t] = lo;
  }
  __syncthreads();
  // exclusive scan over k = ix * nseg + s of len[k] = colstart[s][ix + 1] - colstart[s][ix], 1024 elements per round
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = nseg * nx;
  for (int k0 = 0; k0 < n; k0 += 1024) {
    const int k = k0 + threadIdx.x;
    int len = 0;
    if (k < n) {   // k = ix * nseg + (y-order rank); the segment holding that rank is found by inverting seg_order
      const int ix = k / nseg, o = k - ix * nseg;
      int s = o;
      if (cyc_world > 0) {
        const int spr = nseg / cyc_world;
        s = (o % cyc_world) * spr + o / cyc_world;
      }
      len = colstart[s * (nx + 1) + ix + 1] - colstart[s * (nx + 1) + ix];
    }
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int woff = 0, round_tot = 0;
    for (int w = 0; w < 32; w++) {
      woff += w < warp ? warp_tot[w] : 0;
      round_tot += warp_tot[w];
    }
    if (k < n) base[k] = carry + woff + inc - len;
    __syncthreads();
    if (threadIdx.x == 0) carry += round_tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    total[0] = carry;                       // may exceed out_capacity: the scatter drops rows beyond it
    total[1] = out_capacity;
  }
}

// one warp per gathered row
__global__ void __launch_bounds__(256) k_exchange_scatter(const uint8_t* __restrict__ recv, int nseg, int seg_rows, int row_bytes,
                                                          int stride, int nx, int cyc_world, const int32_t* __restrict__ colstart,
                                                          const int32_t* __restrict__ base, int feat_dim, int num_classes,
                                                          int32_t* __restrict__ coords,
                                                          uint8_t* __restrict__ labels, float* __restrict__ feats,
                                                          float* __restrict__ logits, int out_capacity) {
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (int64_t)nseg * seg_rows) return;
  const int s = (int)(gw / seg_rows), i = (int)(gw - (int64_t)s * seg_rows);
  const uint8_t* seg = recv + (size_t)s * (size_t)(1 + seg_rows) * row_bytes;
  const int cnt = min(max(*reinterpret_cast<const int32_t*>(seg), 0), seg_rows);
  if (i >= cnt) return;
  const uint8_t* row = seg + (size_t)(1 + i) * row_bytes;
  const int4 h = *reinterpret_cast<const int4*>(row);
  const int ix = min(h.x / stride, nx - 1);
  const int dst = base[ix * nseg + seg_order(s, nseg, cyc_world)] + (i - colstart[s * (nx + 1) + ix]);
  if (dst < 0 || dst >= out_capacity) return;
  if (lane == 0) {
    coords[2 * dst] = h.x, coords[2 * dst + 1] = h.y;
    labels[dst] = (uint8_t)h.z;
  }
  if (feat_dim) {
    const uint4* f = reinterpret_cast<const uint4*>(row + kXchgHeadBytes);
    uint4* d = reinterpret_cast<uint4*>(feats + (size_t)dst * feat_dim);
    for (int k = lane; k < feat_dim / 4; k += 32) d[k] = f[k];
  }
  if (logits) {
    const float* sl = reinterpret_cast<const float*>(row + kXchgHeadBytes + feat_dim * 4);
    for (int k = lane; k < num_classes; k += 32) logits[(size_t)dst * num_classes + k] = sl[k];
  }
}

}  // namespace hipac

using namespace hipac;

static bool xchg_dims_ok(int feat_dim, int num_classes) {
  return (feat_dim == 0 || feat_dim == HIPAC_FEATURE_DIM) && num_classes >= 0 && num_classes <= 1024;
}

extern "C" size_t hipac_exchange_row_bytes(int feat_dim, int num_classes) {
  return xchg_dims_ok(feat_dim, num_classes) ? (size_t)xchg_row_bytes(feat_dim, num_classes) : 0;
}

extern "C" size_t hipac_exchange_segment_bytes(int capacity, int feat_dim, int num_classes) {
  return capacity < 0 || !xchg_dims_ok(feat_dim, num_classes) ? 0 : (size_t)(1 + capacity) * xchg_row_bytes(feat_dim, num_classes);
}

extern "C" size_t hipac_exchange_workspace_bytes(int num_segments, int nx) {
  if (num_segments <= 0 || nx <= 0) return 0;
  return align_up((size_t)num_segments * (nx + 1) * 4, 256) + align_up((size_t)num_segments * nx * 4, 256);
}

extern "C" int hipac_exchange_pack(const int32_t* d_coords, const uint8_t* d_labels, const float* d_feats, const float* d_logits,
                                   int feat_dim, int num_classes, const int32_t* d_count, int capacity, int y_offset, void* d_segment,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HIPAC_REQUIRE(d_coords && d_labels && d_count && d_segment, "null pointer");
  HIPAC_REQUIRE(xchg_dims_ok(feat_dim, num_classes), "feat_dim must be 0 or 512, num_classes in [0, 1024]");
  HIPAC_REQUIRE(capacity >= 0 && (num_classes == 0 || d_logits) && (feat_dim == 0 || d_feats), "bad capacity or missing feature / logit buffer");
  HIPAC_REQUIRE(((uintptr_t)d_segment & 15) == 0 && ((uintptr_t)d_feats & 15) == 0, "segment and feature buffers must be 16-byte aligned");
  const int row_bytes = xchg_row_bytes(feat_dim, num_classes);
  const int warps = capacity > 0 ? capacity : 1;
  {
    ProfileScope ps("exchange_pack", stream, (double)capacity * row_bytes);
    k_exchange_pack<<<(warps + 7) / 8, 256, 0, stream>>>(d_coords, d_labels, d_feats, num_classes ? d_logits : nullptr, feat_dim, num_classes,
                                                         d_count, capacity, y_offset, reinterpret_cast<uint8_t*>(d_segment), row_bytes);
  }
  count_launch(1);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hipac_exchange_merge(const void* d_segments, int num_segments, int capacity, int feat_dim, int num_classes, int stride, int nx,
                                    int cyclic_world, int32_t* d_coords, uint8_t* d_labels, float* d_feats, float* d_logits, int32_t* d_total,
                                    int out_capacity, void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HIPAC_REQUIRE(d_segments && d_coords && d_labels && d_total && d_workspace, "null pointer");
  HIPAC_REQUIRE(num_segments > 0 && capacity >= 0 && stride > 0 && nx > 0 && out_capacity >= 0, "bad geometry");
  HIPAC_REQUIRE(cyclic_world >= 0 && (cyclic_world == 0 || num_segments % cyclic_world == 0), "cyclic_world must divide num_segments");
  HIPAC_REQUIRE(xchg_dims_ok(feat_dim, num_classes), "feat_dim must be 0 or 512, num_classes in [0, 1024]");
  HIPAC_REQUIRE((num_classes == 0 || d_logits) && (feat_dim == 0 || d_feats), "feature / logit buffer missing");
  HIPAC_REQUIRE(workspace_bytes >= hipac_exchange_workspace_bytes(num_segments, nx), "workspace too small");
  HIPAC_REQUIRE(((uintptr_t)d_workspace & 255) == 0 && ((uintptr_t)d_segments & 15) == 0 && ((uintptr_t)d_feats & 15) == 0,
                "workspace must be 256-byte, segments and features 16-byte aligned");
  const int row_bytes = xchg_row_bytes(feat_dim, num_classes);
  int32_t* colstart = reinterpret_cast<int32_t*>(d_workspace);
  int32_t* base = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(d_workspace) + align_up((size_t)num_segments * (nx + 1) * 4, 256));
  const uint8_t* recv = reinterpret_cast<const uint8_t*>(d_segments);
  {
    ProfileScope ps("exchange_index", stream, 0.0);
    k_exchange_index<<<1, 1024, 0, stream>>>(recv, num_segments, capacity, row_bytes, stride, nx, cyclic_world, colstart, base, d_total, out_capacity);
  }
  count_launch(1);
  const int64_t warps = (int64_t)num_segments * capacity;
  if (warps > 0) {
    ProfileScope ps("exchange_scatter", stream, (double)warps * row_bytes);
    k_exchange_scatter<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(recv, num_segments, capacity, row_bytes, stride, nx, cyclic_world, colstart, base,
                                                                      feat_dim, num_classes, d_coords, d_labels, d_feats,
                                                                      num_classes ? d_logits : nullptr, out_capacity);
    count_launch(1);
  }
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
