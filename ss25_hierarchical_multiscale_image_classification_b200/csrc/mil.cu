// ABMIL head over a bag of patch features: the consumer of the [N, 512] feature matrix the hot
// path produces.  Reference: MILAttentionPooling + MILClassifier, src/models/mil_classifier.py:5-45
//     A = tanh(x V^T + b_V)  [N,128];  s = A u + b_u  [N];  a = softmax(s over the N instances);  M = sum_i a_i x_i  [512]
//     logits = W2 relu(W1 M + b1) + b2
// (pooling = 'mean' / 'max' replace the attention by x.mean(0) / x.max(0)).  Everything is fp32, as in the reference.
//
// Two launches, no host round trip (the instance count may live in device memory, e.g. hipac_tile_scan's survivor counter):
//   k_mil_partial  every CTA walks tiles of 32 instances: scores (the 512 x 128 projection is read from L2 in a [512][128]
//                  transposed layout so that a warp reads consecutive floats; the tile of x sits in shared memory), then
//                  an ONLINE softmax: running max m, running sum l = sum exp(s_i - m) and running pooled vector
//                  sum exp(s_i - m) x_i, rescaled whenever the maximum moves.  One partial (m, l, vector) per CTA.
//   k_mil_finish   one CTA merges the partials (the usual log-sum-exp merge), normalises the attention weights in place,
//                  and runs the two small Linear layers.
#include "common.cuh"

namespace hipac {

constexpr int kMilF = HIPAC_FEATURE_DIM;   // 512
constexpr int kMilA = 128;                 // attention width (reference default attn_dim = 128) and hidden width of the MLP
constexpr int kMilTile = 32;               // instances per tile
constexpr int kMilThreads = 256;

// packed parameter layout (floats): Vt[512][128] | bV[128] | u[128] | bu[1 (+3 pad)] | W1t[512][128] | b1[128] | W2[k][128] | b2[k]
__host__ __device__ inline size_t mil_off_bv() { return (size_t)kMilF * kMilA; }
__host__ __device__ inline size_t mil_off_u() { return mil_off_bv() + kMilA; }
__host__ __device__ inline size_t mil_off_bu() { return mil_off_u() + kMilA; }
__host__ __device__ inline size_t mil_off_w1() { return mil_off_bu() + 4; }
__host__ __device__ inline size_t mil_off_b1() { return mil_off_w1() + (size_t)kMilF * kMilA; }
__host__ __device__ inline size_t mil_off_w2() { return mil_off_b1() + kMilA; }
__host__ __device__ inline size_t mil_off_b2(int k) { return mil_off_w2() + (size_t)k * kMilA; }
__host__ __device__ inline size_t mil_total(int k) { return mil_off_b2(k) + (size_t)((k + 3) / 4 * 4); }

struct MilPartial {   // per CTA: [0] = running max, [1] = running sum, [2 .. 2 + 512) = running pooled vector
  float v[2 + kMilF];
};

__device__ __forceinline__ int mil_count(const int* n_dev, int n_cap) {
  if (!n_dev) return n_cap;
  const int n = __ldg(n_dev);
  return n < 0 ? 0 : (n < n_cap ? n : n_cap);
}

// mode 0 = attention, 1 = mean (all scores 0), 2 = max (column-wise maximum; "l" counts instances)
__global__ void __launch_bounds__(kMilThreads) k_mil_partial(const float* __restrict__ x, int n_cap, const int* __restrict__ n_dev,
                                                             const float* __restrict__ prm, int mode, float* __restrict__ scores,
                                                             MilPartial* __restrict__ part) {
  extern __shared__ float sm[];
  float* sx = sm;                         // [32][512] tile of instances
  float* ssc = sx + kMilTile * kMilF;     // [32] scores of the tile
  float* sred = ssc + kMilTile;           // [8 warps][32] partial dot products
  const int n = mil_count(n_dev, n_cap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this thread owns feature columns tid and tid + 256 of the running pooled vector
  float m_run = -INFINITY, l_run = 0.f, acc0 = mode == 2 ? -INFINITY : 0.f, acc1 = acc0;
  const float* Vt = prm;
  const float bv = __ldg(prm + mil_off_bv() + (tid & (kMilA - 1)));
  const float uj = __ldg(prm + mil_off_u() + (tid & (kMilA - 1)));
  const float bu = __ldg(prm + mil_off_bu());
  for (int t0 = blockIdx.x * kMilTile; t0 < n; t0 += gridDim.x * kMilTile) {
    const int cnt = min(kMilTile, n - t0);
    __syncthreads();   // previous tile fully consumed
    for (int e = tid; e < kMilTile * kMilF / 4; e += kMilThreads) {
      const int i = e / (kMilF / 4), c4 = e - i * (kMilF / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < cnt) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)(t0 + i) * kMilF) + c4);
      reinterpret_cast<float4*>(sx)[e] = v;
    }
    __syncthreads();
    if (mode == 0) {
      // thread = (attention unit j, instance half hf): 16 instances x one unit; Vt[k][j] is coalesced over j
      const int j = tid & (kMilA - 1), hf = tid >> 7;
      float a[16];
#pragma unroll
      for (int q = 0; q < 16; q++) a[q] = 0.f;
      for (int k = 0; k < kMilF; k += 4) {
        const float w0 = __ldg(Vt + (size_t)k * kMilA + j), w1 = __ldg(Vt + (size_t)(k + 1) * kMilA + j);
        const float w2 = __ldg(Vt + (size_t)(k + 2) * kMilA + j), w3 = __ldg(Vt + (size_t)(k + 3) * kMilA + j);
#pragma unroll
        for (int q = 0; q < 16; q++) {
          const float4 xv = *reinterpret_cast<const float4*>(&sx[(hf * 16 + q) * kMilF + k]);   // warp-wide broadcast
          a[q] = fmaf(w3, xv.w, fmaf(w2, xv.z, fmaf(w1, xv.y, fmaf(w0, xv.x, a[q]))));
        }
      }
      // s_i = sum_j u_j tanh(a_ij + bV_j): reduce over the 128 units = 4 warps of this half
#pragma unroll
      for (int q = 0; q < 16; q++) {
        float v = uj * tanhf(a[q] + bv);
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sred[warp * 32 + q] = v;
      }
      __syncthreads();
      if (tid < kMilTile) {
        const int h2 = tid >> 4, q = tid & 15;
        const float s = sred[(h2 * 4 + 0) * 32 + q] + sred[(h2 * 4 + 1) * 32 + q] + sred[(h2 * 4 + 2) * 32 + q] +
                        sred[(h2 * 4 + 3) * 32 + q] + bu;
        ssc[tid] = s;
        if (tid < cnt && scores) scores[t0 + tid] = s;
      }
    } else if (tid < kMilTile) {
      ssc[tid] = 0.f;
      if (tid < cnt && scores) scores[t0 + tid] = 0.f;
    }
    __syncthreads();
    if (mode == 2) {
      for (int i = 0; i < cnt; i++) acc0 = fmaxf(acc0, sx[i * kMilF + tid]), acc1 = fmaxf(acc1, sx[i * kMilF + tid + 256]);
      l_run += (float)cnt;
      m_run = 0.f;
    } else {
      float m_new = m_run;
      for (int i = 0; i < cnt; i++) m_new = fmaxf(m_new, ssc[i]);
      const float scale = m_run == -INFINITY ? 0.f : expf(m_run - m_new);
      acc0 *= scale, acc1 *= scale, l_run *= scale;
      for (int i = 0; i < cnt; i++) {
        const float w = expf(ssc[i] - m_new);
        l_run += w;
        acc0 = fmaf(w, sx[i * kMilF + tid], acc0), acc1 = fmaf(w, sx[i * kMilF + tid + 256], acc1);
      }
      m_run = m_new;
    }
  }
  MilPartial& p = part[blockIdx.x];
  if (tid == 0) p.v[0] = m_run, p.v[1] = l_run;
  p.v[2 + tid] = acc0, p.v[2 + tid + 256] = acc1;
}

__global__ void __launch_bounds__(512) k_mil_finish(const MilPartial* __restrict__ part, int n_part, int n_cap, const int* __restrict__ n_dev,
                                                    const float* __restrict__ prm, int num_classes, int mode, float* __restrict__ scores,
                                                    float* __restrict__ pooled_out, float* __restrict__ logits) {
  __shared__ float s_m, s_l;
  __shared__ float s_pool[kMilF];
  __shared__ float s_hid[kMilA];
  const int tid = threadIdx.x;
  const int n = mil_count(n_dev, n_cap);
  if (tid == 0) {
    float m = -INFINITY, l = 0.f;
    for (int b = 0; b < n_part; b++) m = fmaxf(m, part[b].v[0]);
    for (int b = 0; b < n_part; b++)
      if (part[b].v[1] > 0.f) l += mode == 2 ? part[b].v[1] : part[b].v[1] * expf(part[b].v[0] - m);
    s_m = m, s_l = l;
  }
  __syncthreads();
  const float m = s_m, l = s_l;
  {
    float acc = mode == 2 ? -INFINITY : 0.f;
    for (int b = 0; b < n_part; b++) {
      if (!(part[b].v[1] > 0.f)) continue;                       // CTA saw no instance
      if (mode == 2) acc = fmaxf(acc, part[b].v[2 + tid]);
      else acc = fmaf(part[b].v[2 + tid], expf(part[b].v[0] - m), acc);
    }
    const float pooled = n == 0 ? 0.f : (mode == 2 ? acc : acc / l);
    s_pool[tid] = pooled;
    if (pooled_out) pooled_out[tid] = pooled;
  }
  __syncthreads();
  // attention weights a_i = exp(s_i - m) / l, in place (mean: 1 / n; max pooling has none)
  if (scores && mode != 2)
    for (int i = tid; i < n; i += blockDim.x) scores[i] = expf(scores[i] - m) / l;
  // classifier: Linear(512,128) + ReLU + Linear(128,k)
  if (tid < kMilA) {
    const float* W1t = prm + mil_off_w1();
    float h = __ldg(prm + mil_off_b1() + tid);
    for (int k = 0; k < kMilF; k++) h = fmaf(__ldg(W1t + (size_t)k * kMilA + tid), s_pool[k], h);
    s_hid[tid] = fmaxf(h, 0.f);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = warp; c < num_classes; c += blockDim.x / 32) {
    const float* w2 = prm + mil_off_w2() + (size_t)c * kMilA;
    float s = 0.f;
    for (int j = lane; j < kMilA; j += 32) s = fmaf(__ldg(w2 + j), s_hid[j], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logits[c] = s + __ldg(prm + mil_off_b2(num_classes) + c);
  }
}

static int mil_grid(int n_cap) {
  int sms = 148;
  device_sm_count(&sms);
  const int tiles = (n_cap + kMilTile - 1) / kMilTile;
  return tiles < 1 ? 1 : (tiles < 2 * sms ? tiles : 2 * sms);
}

}  // namespace hipac

using namespace hipac;

extern "C" size_t hipac_mil_packed_floats(int num_classes) { return num_classes <= 0 || num_classes > 1024 ? 0 : mil_total(num_classes); }

// Host-side repack of the reference module's tensors (row-major, torch layout):
//   V [128][512], bV [128], u [128] (attn_U.weight [1][128]), bu [1], W1 [128][512], b1 [128], W2 [k][128], b2 [k]
// V / u / bV / bu may be NULL for 'mean' / 'max' pooling (that part of the blob stays zero).
extern "C" int hipac_mil_pack(const float* V, const float* bV, const float* u, const float* bu, const float* W1, const float* b1,
                              const float* W2, const float* b2, int num_classes, float* h_packed) {
  HIPAC_REQUIRE(W1 && b1 && W2 && b2 && h_packed && num_classes > 0 && num_classes <= 1024, "bad classifier tensors");
  memset(h_packed, 0, mil_total(num_classes) * sizeof(float));
  if (V) {
    HIPAC_REQUIRE(bV && u && bu, "attention pooling needs V, bV, u and bu");
    for (int j = 0; j < kMilA; j++)
      for (int k = 0; k < kMilF; k++) h_packed[(size_t)k * kMilA + j] = V[(size_t)j * kMilF + k];
    memcpy(h_packed + mil_off_bv(), bV, kMilA * sizeof(float));
    memcpy(h_packed + mil_off_u(), u, kMilA * sizeof(float));
    h_packed[mil_off_bu()] = bu[0];
  }
  for (int j = 0; j < kMilA; j++)
    for (int k = 0; k < kMilF; k++) h_packed[mil_off_w1() + (size_t)k * kMilA + j] = W1[(size_t)j * kMilF + k];
  memcpy(h_packed + mil_off_b1(), b1, kMilA * sizeof(float));
  memcpy(h_packed + mil_off_w2(), W2, (size_t)num_classes * kMilA * sizeof(float));
  memcpy(h_packed + mil_off_b2(num_classes), b2, (size_t)num_classes * sizeof(float));
  return 0;
}

extern "C" size_t hipac_mil_workspace_bytes(int n_instances) {
  if (n_instances < 0) return 0;
  return align_up((size_t)2 * 148 * 2 * sizeof(MilPartial), 256);   // partials of at most 2 CTAs per SM (generous for any B200)
}

extern "C" int hipac_mil_forward(const float* d_x, int n_instances, const int32_t* d_count, const float* d_packed, int num_classes,
                                 int pooling, float* d_logits, float* d_attention, float* d_pooled, void* d_workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  HIPAC_REQUIRE(d_packed && d_logits && d_workspace && n_instances >= 0 && (n_instances == 0 || d_x), "null pointer");
  HIPAC_REQUIRE(num_classes > 0 && num_classes <= 1024, "bad num_classes");
  HIPAC_REQUIRE(pooling >= 0 && pooling <= 2, "pooling: 0 attention, 1 mean, 2 max");
  HIPAC_REQUIRE(workspace_bytes >= hipac_mil_workspace_bytes(n_instances), "workspace too small");
  HIPAC_REQUIRE(((uintptr_t)d_x & 15) == 0 && ((uintptr_t)d_workspace & 15) == 0, "features and workspace must be 16-byte aligned");
  MilPartial* part = reinterpret_cast<MilPartial*>(d_workspace);
  const int grid = mil_grid(n_instances);
  HIPAC_REQUIRE((size_t)grid * sizeof(MilPartial) <= workspace_bytes, "workspace too small for this device");
  const size_t smem = (size_t)(kMilTile * kMilF + kMilTile + 8 * 32) * sizeof(float);
  if (int e = ensure_dyn_smem(k_mil_partial, (int)smem)) return e;
  {
    ProfileScope ps("mil_partial", stream, 2.0 * n_instances * kMilF * kMilA);
    k_mil_partial<<<grid, kMilThreads, smem, stream>>>(d_x, n_instances, d_count, d_packed, pooling, d_attention, part);
  }
  {
    ProfileScope ps("mil_finish", stream, 0.0);
    k_mil_finish<<<1, 512, 0, stream>>>(part, grid, n_instances, d_count, d_packed, num_classes, pooling, d_attention, d_pooled, d_logits);
  }
  count_launch(2);
  HIPAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
