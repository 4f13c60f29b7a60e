// Lesion-annotation rasterisation on the device, bit-exact against Pillow's ImageDraw.polygon(outline=255, fill=255) on an
// "L" image -- the call the reference's parse_xml_mask makes for every CAMELYON16 annotation (src/main.py:388-409) and
// whose result is the lesion mask the tile scan votes on (src/main.py:705-716).  With it the mask is born in HBM: the
// host ships a few kilobytes of integer vertices instead of scanning and uploading an H x W byte image.
//
// Pillow (third-party, not vendored by the reference; 12.2.0 in the container) is restated, not linked:
// oracle/pil_polygon.py documents the algorithm as read back from the installed binary and pins it against the installed
// Pillow on thousands of random polygons; tests/test_polygon_gpu.py compares this kernel with both.  Summary:
//   * outline == fill  ->  only the fill is drawn (ImageDraw.polygon), by polygon_generic;
//   * edges join consecutive vertices (+ closing edge unless last == first); horizontal edges are drawn as
//     hline(xmin..xmax) (merging consecutive collinear horizontal edges, as ImagingDrawPolygon does, yields the same
//     pixels); every other edge has dx = (float)(x1 - x0) / (y1 - y0);
//   * for every row y of the polygon's (clamped) y range, the edges active on it (ymin <= y <= ymax), IN EDGE ORDER,
//     contribute x = (float)(y - y0) * dx + (float)x0 (float32 multiply, then add); an edge ending on the row
//     contributes it twice unless the row is the polygon's last; an edge with a vertex on the row may be pulled towards
//     the adjacent row's span by the "discontiguous corner" rule (see corner_rule below); the sorted values are paired
//     and each pair fills ROUND_UP(x0) .. ROUND_DOWN(x1), clipped to the image.
// All polygons are filled with the same ink, so the result does not depend on the order in which rows or polygons finish.
//
// One CTA per (polygon, row).  Edge data is recomputed from the vertex list (two int2 loads per edge), the row's active
// edges are compacted in order into shared memory, every active edge evaluates its own intersection (the corner rule only
// reads edge data, never other intersections), a rank sort orders the handful of values, and the spans are written with
// coalesced byte stores.
#include "common.cuh"

namespace hipac {

constexpr int kPolyThreads = 128;
constexpr int kPolyMaxActive = 1024;   // intersections per row (a row crossing more edges than this is reported as an error)
constexpr int kPolyMaxVertexEdges = 256;

struct PolyEdge {   // Pillow's Edge (Draw.c), rebuilt on the fly
  int x0, y0, ymin, ymax, xmin, xmax;
  float dx;
};

__device__ __forceinline__ PolyEdge make_edge(int x0, int y0, int x1, int y1) {
  PolyEdge e;
  e.x0 = x0, e.y0 = y0;
  e.xmin = min(x0, x1), e.xmax = max(x0, x1);
  e.ymin = min(y0, y1), e.ymax = max(y0, y1);
  e.dx = y0 == y1 ? 0.0f : __fdiv_rn((float)(x1 - x0), (float)(y1 - y0));
  return e;
}
// edge i of a polygon with n vertices: i -> i + 1, the closing edge (n - 1 -> 0) has index n - 1
__device__ __forceinline__ PolyEdge load_edge(const int2* __restrict__ v, int n, int i) {
  const int2 a = v[i], b = v[i + 1 == n ? 0 : i + 1];
  return make_edge(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float x_at(const PolyEdge& e, int y) { return __fadd_rn(__fmul_rn((float)(y - e.y0), e.dx), (float)e.x0); }
__device__ __forceinline__ float c_roundf(float x) {   // C roundf: half away from zero (exact: x - trunc(x) is exact)
  const float t = truncf(x);
  return fabsf(__fsub_rn(x, t)) >= 0.5f ? __fadd_rn(t, copysignf(1.0f, x)) : t;
}
// Pillow's ROUND_UP / ROUND_DOWN macros: float arithmetic for f >= 0, double for f < 0 (fabs() promotes)
__device__ __forceinline__ int round_up(float f) {
  return f >= 0.0f ? (int)floorf(__fadd_rn(f, 0.5f)) : -(int)floor(__dadd_rn(fabs((double)f), 0.5));
}
__device__ __forceinline__ int round_down(float f) {
  return f >= 0.0f ? (int)ceilf(__fsub_rn(f, 0.5f)) : -(int)ceil(__dsub_rn(fabs((double)f), 0.5));
}

struct PolyJob {        // one per polygon, built by the host wrapper from the vertex lists
  int v_off, n_vtx;     // vertices [v_off, v_off + n_vtx)
  int n_edges;          // n_vtx - 1 (+ 1 if the polygon is not explicitly closed)
  int ymin, ymax;       // Pillow's clamped scan range: max(min y, 0) .. min(max y, H)
  int y_first;          // first row of this polygon that this call rasterises (row window of the level image)
  int row_off;          // first work item (row) of this polygon
};

constexpr int kPolyJobsPerLaunch = 128;
struct PolyJobs {       // travels as a kernel parameter (3 KB): no host -> device copy, nothing to wait for
  int n;
  PolyJob j[kPolyJobsPerLaunch];
};

__global__ void __launch_bounds__(kPolyThreads) k_polygon_fill(const int2* __restrict__ vtx, const __grid_constant__ PolyJobs jobs,
                                                              uint8_t* __restrict__ mask, int y_begin, int y_end, int W, int64_t pitch,
                                                              int* __restrict__ overflow) {
  // ---- which (polygon, row) is this CTA? (binary search over the per-polygon row offsets) ----
  int lo = 0, hi = jobs.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs.j[mid].row_off <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const PolyJob job = jobs.j[lo];
  const int y = job.y_first + ((int)blockIdx.x - job.row_off);
  if (y > job.ymax || y < y_begin || y >= y_end) return;
  const int2* v = vtx + job.v_off;
  const int n = job.n_vtx, ne = job.n_edges;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* row = mask + (int64_t)(y - y_begin) * pitch;

  __shared__ int s_act[kPolyMaxActive];          // edge indices active on this row, in edge order
  __shared__ int s_vtx[kPolyMaxVertexEdges];     // non-horizontal edges with a vertex on this row, in edge order
  __shared__ float s_x[2 * kPolyMaxActive];
  __shared__ float s_sorted[2 * kPolyMaxActive];
  __shared__ int s_warp[2][kPolyThreads / 32];
  __shared__ int s_nact, s_nvtx, s_nx;
  if (tid == 0) s_nact = 0, s_nvtx = 0, s_nx = 0;
  __syncthreads();

  // ---- pass 1: ordered compaction of the active / vertex edges; horizontal edges of this row are drawn right away ----
  for (int i0 = 0; i0 < ne; i0 += kPolyThreads) {
    const int i = i0 + tid;
    bool act = false, vt = false;
    if (i < ne) {
      const PolyEdge e = load_edge(v, n, i);
      if (e.ymin == e.ymax) {
        if (e.ymin == y) {                       // hline(xmin, y, xmax)
          const int a = max(e.xmin, 0), b = min(e.xmax, W - 1);
          if (e.xmin < W && e.xmax >= 0)
            for (int x = a; x <= b; x++) row[x] = 255;
        }
      } else {
        act = y >= e.ymin && y <= e.ymax;
        vt = act && (y == e.ymin || y == e.ymax) && e.dx != 0.0f;
      }
    }
    const unsigned ba = __ballot_sync(0xffffffffu, act), bv = __ballot_sync(0xffffffffu, vt);
    if (lane == 0) s_warp[0][warp] = __popc(ba), s_warp[1][warp] = __popc(bv);
    __syncthreads();
    int offa = s_nact, offv = s_nvtx;
    for (int w = 0; w < warp; w++) offa += s_warp[0][w], offv += s_warp[1][w];
    const int pa = offa + __popc(ba & ((1u << lane) - 1)), pv = offv + __popc(bv & ((1u << lane) - 1));
    if (act && pa < kPolyMaxActive) s_act[pa] = i;
    if (vt && pv < kPolyMaxVertexEdges) s_vtx[pv] = i;
    __syncthreads();
    if (tid == 0) {
      int ta = 0, tv = 0;
      for (int w = 0; w < kPolyThreads / 32; w++) ta += s_warp[0][w], tv += s_warp[1][w];
      s_nact += ta, s_nvtx += tv;
    }
    __syncthreads();
  }
  const int nact = s_nact, nvtx = s_nvtx;
  if (nact > kPolyMaxActive || nvtx > kPolyMaxVertexEdges) {       // (s_x holds two values per active edge)
    if (tid == 0) atomicExch(overflow, 1);
    return;
  }

  // ---- pass 2: every active edge evaluates its intersection(s) ----
  for (int a = tid; a < nact; a += kPolyThreads) {
    const int i = s_act[a];
    const PolyEdge cur = load_edge(v, n, i);
    float x = x_at(cur, y);
    int reps = 1;
    if (y == cur.ymax && y < job.ymax) {
      reps = 2;                                  // "needed to draw consistent polygons"
    } else if ((y == cur.ymin || y == cur.ymax) && cur.dx != 0.0f) {
      // discontiguous-corner rule: an EARLIER edge sharing this vertex (same rounded x on this row) that is active on the
      // adjacent row; if the corner lies more than a pixel beyond both edges there, pull it to one pixel beyond the farther
      const int adj = y == cur.ymax ? y - 1 : y + 1;
      const float rx = c_roundf(x);
      for (int q = 0; q < nvtx; q++) {
        const int k = s_vtx[q];
        if (k >= i) break;
        const PolyEdge other = load_edge(v, n, k);
        if (c_roundf(x_at(other, y)) != rx) continue;
        if (adj < other.ymin || adj > other.ymax) continue;
        const float A = x_at(cur, adj), B = x_at(other, adj);
        if (x > __fadd_rn(A, 1.0f) && x > __fadd_rn(B, 1.0f)) x = __fadd_rn(c_roundf(fmaxf(A, B)), 1.0f);
        else if (__fsub_rn(A, 1.0f) > x && __fsub_rn(B, 1.0f) > x) x = __fsub_rn(c_roundf(fminf(A, B)), 1.0f);
        break;
      }
    }
    const int slot = atomicAdd(&s_nx, reps);     // order is irrelevant: the values are sorted next
    s_x[slot] = x;
    if (reps == 2) s_x[slot + 1] = x;
  }
  __syncthreads();
  const int nx = s_nx;

  // ---- rank sort (ties broken by slot: any order of equal values gives the same pairs) ----
  for (int a = tid; a < nx; a += kPolyThreads) {
    const float xa = s_x[a];
    int rank = 0;
    for (int b = 0; b < nx; b++) {
      const float xb = s_x[b];
      rank += (xb < xa) || (xb == xa && b < a);
    }
    s_sorted[rank] = xa;
  }
  __syncthreads();

  // ---- spans ----
  for (int pr = 0; pr + 1 < nx; pr += 2) {
    int xs = round_up(s_sorted[pr]), xe = round_down(s_sorted[pr + 1]);
    if (xs < 0) xs = 0;
    else if (xs >= W) continue;
    if (xe < 0) continue;
    else if (xe >= W) xe = W - 1;
    for (int x = xs + tid; x <= xe; x += kPolyThreads) row[x] = 255;
  }
}

}  // namespace hipac

using namespace hipac;

extern "C" size_t hipac_polygon_workspace_bytes(int num_polygons) { return num_polygons < 0 ? 0 : 256; }

// h_xy: all polygons' integer vertices back to back; h_offsets[p] .. h_offsets[p + 1] = vertices of polygon p (HOST arrays:
// the scan range of every polygon is derived from them here).  d_xy: the same vertices in device memory.
extern "C" int hipac_polygon_fill(const int32_t* h_xy, const int32_t* h_offsets, int num_polygons, const int32_t* d_xy, uint8_t* d_mask,
                                  int H, int W, int64_t pitch, int y_begin, int n_rows, int clear, void* d_workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HIPAC_REQUIRE(d_mask && H > 0 && W > 0 && pitch >= W, "bad mask geometry");
  HIPAC_REQUIRE(y_begin >= 0 && n_rows >= 0 && y_begin + n_rows <= H, "row window outside the level image");
  const int y_end = y_begin + n_rows;
  HIPAC_REQUIRE(num_polygons >= 0 && d_workspace && (num_polygons == 0 || (h_xy && h_offsets && d_xy)), "null pointer");
  HIPAC_REQUIRE(workspace_bytes >= hipac_polygon_workspace_bytes(num_polygons), "workspace too small");
  HIPAC_REQUIRE(((uintptr_t)d_xy & 7) == 0, "vertex buffer must be 8-byte aligned");
  if (clear && n_rows > 0) HIPAC_CHECK_CUDA(cudaMemset2DAsync(d_mask, (size_t)pitch, 0, (size_t)W, (size_t)n_rows, stream));
  int* d_overflow = reinterpret_cast<int*>(d_workspace);
  HIPAC_CHECK_CUDA(cudaMemsetAsync(d_overflow, 0, 4, stream));
  PolyJobs jobs;
  jobs.n = 0;
  long long rows = 0;
  auto flush = [&]() -> int {
    if (jobs.n == 0 || rows == 0) return 0;
    {
      ProfileScope ps("polygon_fill", stream, (double)rows);
      k_polygon_fill<<<(unsigned)rows, kPolyThreads, 0, stream>>>(reinterpret_cast<const int2*>(d_xy), jobs, d_mask, y_begin, y_end, W, pitch,
                                                                  d_overflow);
    }
    count_launch(1);
    HIPAC_CHECK_CUDA(cudaGetLastError());
    jobs.n = 0, rows = 0;
    return 0;
  };
  for (int p = 0; p < num_polygons; p++) {
    const int off = h_offsets[p], n = h_offsets[p + 1] - off;
    HIPAC_REQUIRE(n >= 0, "polygon offsets must be non-decreasing");
    if (n == 0) continue;
    const int32_t* xy = h_xy + 2 * (size_t)off;
    const bool closed = xy[2 * (n - 1)] == xy[0] && xy[2 * (n - 1) + 1] == xy[1];
    const int ne = (n - 1) + (closed ? 0 : 1);          // ImagingDrawPolygon: closing edge unless last == first
    if (ne <= 0) continue;                              // a single vertex draws nothing
    int ymin = H - 1, ymax = 0;                         // polygon_generic's initial values and clamps
    for (int i = 0; i < n; i++) {
      ymin = xy[2 * i + 1] < ymin ? xy[2 * i + 1] : ymin;
      ymax = xy[2 * i + 1] > ymax ? xy[2 * i + 1] : ymax;
    }
    if (ymin < 0) ymin = 0;
    if (ymax > H) ymax = H;
    const int y_first = ymin > y_begin ? ymin : y_begin;            // rows of this polygon inside the window
    const int y_last = ymax < y_end - 1 ? ymax : y_end - 1;
    if (y_last < y_first) continue;
    if (jobs.n == kPolyJobsPerLaunch || rows + (y_last - y_first + 1) >= (1ll << 30))
      if (int e = flush()) return e;
    PolyJob& j = jobs.j[jobs.n++];
    j.v_off = off, j.n_vtx = n, j.n_edges = ne, j.ymin = ymin, j.ymax = ymax, j.y_first = y_first, j.row_off = (int)rows;
    rows += (long long)(y_last - y_first + 1);
  }
  return flush();
}

// 1 if a scan line of the last hipac_polygon_fill on this workspace crossed more edges than the kernel's row buffer holds
extern "C" int hipac_polygon_overflowed(const void* d_workspace, void* stream_) {
  int flag = 0;
  HIPAC_CHECK_CUDA(cudaMemcpyAsync(&flag, d_workspace, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  HIPAC_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  return flag;
}
