/*
 * hipac_b200.h -- C ABI of the B200-native HiPAC hot path (libhipac_b200.so).
 *
 * The reference (anacarsi/ss25_Hierarchical_Multiscale_Image_Classification) has no FFI
 * or plugin layer: its boundary is plain Python functions.  Each entry
 * point below names the reference code it replaces; the Python mirror of the
 * reference interface (package ss25_hierarchical_multiscale_image_classification_b200)
 * binds these symbols with ctypes.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (PyTorch),
 *     h_* is host memory; the library allocates no persistent device memory
 *     except small constant tables.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return 0 on success, negative on error; hipac_last_error() gives the text.
 *     Nothing throws across the ABI.
 */
#ifndef HIPAC_B200_H
#define HIPAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIPAC_ABI_VERSION 1

/* batch layouts written by hipac_tile_scan / consumed by hipac_resnet18_forward */
#define HIPAC_LAYOUT_NHWC3_BF16 1 /* bf16 [N,224,224,3], (u8/255-mean)/std, the reference's tensor (src/main.py:815-816) in NHWC */
#define HIPAC_LAYOUT_S2D16_BF16 2 /* bf16 [N,112,115,16]: 2x2 space-to-depth of the above, ch=(dy*2+dx)*3+c, 12..15 zero; column X of the
                                    image sits at index X+2, columns 0,1,114 are zero (conv1's W padding made explicit so that the four
                                    W-taps of a filter row are one contiguous 128-byte TMA row) */
#define HIPAC_S2D16_WIDTH 115

/* stage-1 algorithm selector */
#define HIPAC_SCAN_AUTO   0 /* fused read-once path when stride %% (P/224) == 0, else direct */
#define HIPAC_SCAN_DIRECT 1 /* one window reduction + one resample per patch (any stride) */
#define HIPAC_SCAN_FUSED  2 /* single streaming pass over the level image (error if not applicable) */
#define HIPAC_SCAN_NO_STREAM 0x200 /* OR-able flag: run the fused path with its cp.async kernels (two reads of the image) instead of the
                                      TMA streaming pass; the two are bit-identical, the flag exists for the cross-check tests */
#define HIPAC_SCAN_KEEP_ALL 0x100 /* OR-able flag: skip the tissue rejection (every candidate survives); used when the input
                                     is a stack of already-extracted patches, as in the reference's extract_features */

const char* hipac_last_error(void);
int hipac_abi_version(void);

/* ---------------------------------------------------------------------------------------
 * Stage 1: multiscale patch extraction of ONE level image resident in HBM.
 * Replaces the hot loop of extract_patches (reference src/main.py:682-727): candidate grid
 * x in range(0,W,S), y in range(0,H,S) (x outer), white padding of border patches (688-703),
 * tissue test mean>240 -> reject (718-720, integer form sum <= 240*3*P*P), lesion label
 * any(mask[y:y+P,x:x+P]>0) (705-716), and -- instead of the PNG round trip -- the
 * Resize((224,224)) + ToTensor + Normalize of src/main.py:812-818 (Pillow-exact antialiased
 * bilinear in 22-bit fixed point, two uint8 passes).
 *
 *   d_rgb        uint8 [H][pitch_bytes], 3 bytes per pixel (RGB), level-L image
 *   d_mask       uint8 [H][mask_pitch] rasterised lesion mask (>0 = lesion) or NULL (all "normal")
 *   P, S         patch size (1792>>level) and stride (reference CLI: 224)
 *   iy_begin/end candidate grid rows y/S in [iy_begin, iy_end)  (the multi-GPU shard unit)
 *   outputs      survivors in the reference's emission order (x outer, y inner):
 *     d_coords   int32 [capacity][2] (x, y) level-L pixel coordinates
 *     d_labels   uint8 [capacity]    1 = tumor, 0 = normal
 *     d_batch_u8 uint8 [capacity][224][224][3] Pillow-exact resized patch, or NULL
 *     d_batch    bf16 batch in `layout`, or NULL
 *     d_count    int32 [2]: {survivors, candidates}; survivors may exceed capacity, in which
 *                case only the first `capacity` are written (caller re-runs with a smaller row range)
 *   d_workspace  >= hipac_tile_scan_workspace_bytes(...) bytes, 256-byte aligned
 * ------------------------------------------------------------------------------------- */
size_t hipac_tile_scan_workspace_bytes(int H, int W, int P, int S, int iy_begin, int iy_end, int mode);

int hipac_tile_scan(const uint8_t* d_rgb, int H, int W, int64_t pitch_bytes,
                    const uint8_t* d_mask, int64_t mask_pitch,
                    int P, int S, int iy_begin, int iy_end,
                    int32_t* d_coords, uint8_t* d_labels,
                    uint8_t* d_batch_u8, void* d_batch, int layout,
                    int32_t* d_count, int capacity,
                    void* d_workspace, size_t workspace_bytes,
                    int mode, void* stream);

/* Early survivor count.  If a pinned host buffer int32[2] is registered (calling thread), every hipac_tile_scan copies
 * {survivors, candidates} into it right after the compaction kernel and records an event; hipac_tile_scan_wait_count
 * blocks only until that copy has landed -- the resample / gather kernels of the same call keep running, so the caller
 * can enqueue the next stage without a pipeline bubble.  Pass NULL to unregister. */
int hipac_tile_scan_set_count_buffer(int32_t* h_count_pinned);
int hipac_tile_scan_wait_count(void);

/* Host -> device copy of `rows` rows of `row_bytes` bytes between pitched buffers (one cudaMemcpy2DAsync on `stream`).
 * The read-once streaming pass of hipac_tile_scan needs a level image whose row pitch is a multiple of 16 bytes and whose
 * base is 16-byte aligned (TMA bulk copies); 3*W generally is not, so callers allocate the device image with
 * pitch = (3*W + 15) & ~15 and land the reader's [rows][3*W] host rows in it with this call (no extra pass over HBM).
 * Replaces the `slide.read_region(...)` -> numpy hand-over of the reference's loop (src/main.py:693-697) as the point
 * where pixels enter the path. */
int hipac_upload_rows(void* d_dst, int64_t dst_pitch, const void* h_src, int64_t src_pitch, int64_t row_bytes, int64_t rows,
                      void* stream);

/* ---------------------------------------------------------------------------------------
 * Lesion-mask rasterisation on the device: replaces the Pillow call inside parse_xml_mask
 * (reference src/main.py:388-409: ImageDraw.polygon(coords, outline=255, fill=255) on an "L" image of the level size, one
 * call per annotation), bit-exact against Pillow 12.2.0's polygon fill.  The vertices are the reference's
 * (int(X * level_w / w0), int(Y * level_h / h0)) pairs, computed by the caller.
 *   h_xy / h_offsets  HOST arrays: int32 [total][2] vertices of all polygons back to back, int32 [num_polygons + 1] offsets
 *   d_xy              the same vertices in device memory (8-byte aligned)
 *   H, W              size of the level image the vertices refer to (Pillow clamps the scan range to it)
 *   y_begin, n_rows   the row window [y_begin, y_begin + n_rows) of that image which is rasterised (a slab of the level)
 *   d_mask            uint8 [n_rows][pitch], row 0 = level row y_begin; every covered pixel becomes 255; `clear` zeroes the
 *                     n_rows x W window first
 *   d_workspace       >= hipac_polygon_workspace_bytes() bytes; holds the overflow flag read by hipac_polygon_overflowed
 *                     (1 if some scan line crossed more than 1024 edges: the mask is then incomplete)
 * Asynchronous on `stream` (hipac_polygon_overflowed synchronises it). */
size_t hipac_polygon_workspace_bytes(int num_polygons);
int hipac_polygon_fill(const int32_t* h_xy, const int32_t* h_offsets, int num_polygons, const int32_t* d_xy, uint8_t* d_mask,
                       int H, int W, int64_t pitch, int y_begin, int n_rows, int clear, void* d_workspace, size_t workspace_bytes,
                       void* stream);
int hipac_polygon_overflowed(const void* d_workspace, void* stream);

/* Pillow coefficient tables used by the kernels (known-answer hook for the CPU tests; no GPU needed).
 * scale in {2,4,8}; writes interior[2*scale], left_edge[3*scale/2], right_edge[3*scale/2] (22-bit fixed point). */
int hipac_pillow_coeffs(int scale, int32_t* h_interior, int32_t* h_left, int32_t* h_right);

/* float32 -> bf16 bits of the ToTensor+Normalize LUT, [256][3] (known-answer hook, host only). */
int hipac_normalize_lut_bf16(uint16_t* h_lut);

/* ---------------------------------------------------------------------------------------
 * Stage 2: ResNet18 forward (reference src/models/resnet.py:22-77 = torchvision resnet18
 * trunk, eval-mode BN, global average pool [+ Linear(512,k)]) on a bf16 batch.
 * ------------------------------------------------------------------------------------- */

/* Number of fp32 tensors hipac_resnet18_pack expects and their order: see INTEGRATION.md
 * (conv weight, bn weight, bn bias, bn running_mean, bn running_var per conv, network order;
 * then fc weight [k][512] and fc bias [k], or NULL,NULL for the headless feature extractor). */
#define HIPAC_RESNET18_NUM_CONVS 20
#define HIPAC_RESNET18_NUM_TENSORS (HIPAC_RESNET18_NUM_CONVS * 5 + 2)

size_t hipac_resnet18_packed_bytes(int num_classes);

/* Fold eval-mode BatchNorm into the conv weights (fp32), convert to the bf16 K-major layouts the
 * implicit-GEMM kernels read, write the blob to HOST memory `h_packed` (caller uploads it). */
int hipac_resnet18_pack(const float* const* h_tensors, int num_tensors, int num_classes, float bn_eps,
                        void* h_packed, size_t packed_bytes);

size_t hipac_resnet18_workspace_bytes(int n_patches, int chunk);

/* d_batch: bf16 batch of n_patches in `layout` (HIPAC_LAYOUT_S2D16_BF16 is the native one).
 * d_feats: float32 [n][512]; d_logits: float32 [n][num_classes] or NULL. */
int hipac_resnet18_forward(const void* d_packed, int num_classes,
                           const void* d_batch, int layout, int n_patches,
                           float* d_feats, float* d_logits,
                           void* d_workspace, size_t workspace_bytes, int chunk, void* stream);

/* Same forward with the patch count on the DEVICE: buffers, grids and tensor maps are sized for `capacity` patches and
 * every kernel reads d_count[0] (e.g. the survivor counter written by hipac_tile_scan) when it starts, working on
 * min(d_count[0], capacity) patches.  Nothing waits on the host, so tile scan + forward of one level image are enqueued
 * back to back (and the sequence is CUDA-graph capturable); rows >= d_count[0] of d_feats / d_logits are left untouched.
 * Workspace as for hipac_resnet18_workspace_bytes(capacity, chunk). */
int hipac_resnet18_forward_dcount(const void* d_packed, int num_classes,
                                  const void* d_batch, int layout, int capacity, const int32_t* d_count,
                                  float* d_feats, float* d_logits,
                                  void* d_workspace, size_t workspace_bytes, int chunk, void* stream);

/* Test hook: one conv layer of the network on its own (layer index 0..19, network order), bf16 NHWC in/out
 * (layer 0 takes the S2D16 batch), optional residual.  Used by the per-layer parity tests. */
int hipac_resnet18_conv_layer(const void* d_packed, int num_classes, int layer,
                              const void* d_in, const void* d_residual, void* d_out,
                              int n_patches, int relu, void* stream);

/* Test hook: conv2 of layer{2,3,4}.0 with its 1x1/stride-2 projection shortcut fused into the same accumulator:
 * out = relu(conv3x3(d_in) + bn + conv1x1_s2(d_block_in) + bn); stage 0/1/2 = layer2/3/4. */
int hipac_resnet18_conv_ds_fused(const void* d_packed, int num_classes, int stage, const void* d_in, const void* d_block_in,
                                 void* d_out, int n_patches, void* stream);

/* Test hook: the fused stem (conv1 + folded BN + ReLU + 3x3/s2 max pool): S2D16 batch -> bf16 [n][56][56][64]. */
int hipac_resnet18_stem(const void* d_packed, int num_classes, const void* d_in, void* d_out, int n_patches, void* stream);

/* ---------------------------------------------------------------------------------------
 * The exchange step: per-segment survivors -> ONE canonically ordered result.
 * A segment = the outputs of one tile-scan + forward pass over a contiguous range of candidate grid rows (a rank's
 * shard, or one row group of it): rows in emission order (x outer, y inner), count on the device.  The reference has no
 * counterpart (nn.DataParallel only, src/main.py:839-842); the contract is array equality with a single-rank run, whose
 * row order is the reference's loop order (src/main.py:682-683).
 *
 *   hipac_exchange_pack   writes segment = [1 + capacity][row_bytes] bytes: header row {count}, then per survivor
 *                         {x, y + y_offset, label, features[512], logits[num_classes]}.  Segments of equal capacity laid
 *                         side by side are the send / receive buffers of one fixed-size all-gather (the counts ride in
 *                         the headers: no separate count exchange, no host round trip).
 *   hipac_exchange_merge  num_segments segments, ordered by ascending y range (rank after rank, row group after row
 *                         group) -> final arrays in (x, y) order: a binary search per (segment, grid column), one prefix
 *                         sum, one scatter.  d_total[0] = number of rows (device memory; rows beyond out_capacity are
 *                         dropped), d_total[1] = out_capacity.  stride / nx = candidate grid stride and column count.
 * ------------------------------------------------------------------------------------- */
#define HIPAC_FEATURE_DIM 512
/* feat_dim is 512 (rows carry the features) or 0 (coords / labels / logits only, e.g. the heatmap of configs[3]).
 * cyclic_world: 0 = the segments are stored in ascending y order (contiguous row shards); W > 0 = block-cyclic sharding over W
 * ranks (stored rank-major, as an all-gather delivers them): local segment g of rank r is grid-row block g * W + r, which balances
 * spatially clumped tissue across the ranks.
 * hipac_exchange_pack: `capacity` = rows available in the INPUT tensors (<= the segment's row capacity); min(*d_count,
 * capacity) rows are packed.  hipac_exchange_merge: `capacity` = row capacity of every segment (segment stride =
 * hipac_exchange_segment_bytes(capacity, ...)). */
size_t hipac_exchange_row_bytes(int feat_dim, int num_classes);
size_t hipac_exchange_segment_bytes(int capacity, int feat_dim, int num_classes);
size_t hipac_exchange_workspace_bytes(int num_segments, int nx);
int hipac_exchange_pack(const int32_t* d_coords, const uint8_t* d_labels, const float* d_feats, const float* d_logits,
                        int feat_dim, int num_classes, const int32_t* d_count, int capacity, int y_offset, void* d_segment, void* stream);
int hipac_exchange_merge(const void* d_segments, int num_segments, int capacity, int feat_dim, int num_classes, int stride, int nx,
                         int cyclic_world, int32_t* d_coords, uint8_t* d_labels, float* d_feats, float* d_logits, int32_t* d_total,
                         int out_capacity, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * ABMIL head over a bag of patch features: MILAttentionPooling + MILClassifier of the reference
 * (src/models/mil_classifier.py:5-45), fp32:  a = softmax_i(u . tanh(V x_i + bV) + bu);  M = sum_i a_i x_i;
 * logits = W2 relu(W1 M + b1) + b2.  pooling: 0 attention, 1 mean (x.mean(0)), 2 max (x.max(0)).
 *   hipac_mil_pack     host repack of the module's tensors (torch layouts: V [128][512], bV [128], u = attn_U.weight [1][128],
 *                      bu [1], W1 [128][512], b1 [128], W2 [k][128], b2 [k]; V/bV/u/bu may be NULL for mean / max pooling)
 *                      into hipac_mil_packed_floats(k) floats; the caller uploads the blob.
 *   hipac_mil_forward  d_x float32 [n_instances][512] (e.g. hipac_resnet18_forward's features); d_count: optional int32 device
 *                      counter (min(*d_count, n_instances) instances are pooled, nothing waits on the host);
 *                      d_logits [k]; d_attention [n_instances] or NULL (attention weights, the module's second output);
 *                      d_pooled [512] or NULL.
 * ------------------------------------------------------------------------------------- */
size_t hipac_mil_packed_floats(int num_classes);
int hipac_mil_pack(const float* V, const float* bV, const float* u, const float* bu, const float* W1, const float* b1,
                   const float* W2, const float* b2, int num_classes, float* h_packed);
size_t hipac_mil_workspace_bytes(int n_instances);
int hipac_mil_forward(const float* d_x, int n_instances, const int32_t* d_count, const float* d_packed, int num_classes, int pooling,
                      float* d_logits, float* d_attention, float* d_pooled, void* d_workspace, size_t workspace_bytes, void* stream);

/* Number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py's "gpu_launches"). */
long long hipac_launch_count(int reset);

/* Test hook for the shifted-descriptor mechanism of the row-tile conv kernels:
 * D[m][n] = sum_k A[m + shift_rows][k] * B[n][k] (m < 128, n < 64) with A [256][K], B [64][K] bf16,
 * K = 64 (swizzle_bytes = 128) or 16 (swizzle_bytes = 32); A and B are TMA-loaded whole, the UMMA A
 * descriptor starts shift_rows rows into the tile. */
int hipac_debug_umma_shift(const void* d_A, const void* d_B, float* d_D, int shift_rows, int swizzle_bytes, void* stream);

/* Optional per-kernel profiler (calling thread): when enabled every kernel launch of the library is
 * bracketed by CUDA events on its stream.  hipac_profile_report synchronises them, writes one line per
 * kernel "<name> <launches> <total_ms> <total_work>" (work = algorithmic bytes for the stage-1 kernels,
 * flops for the conv kernels), clears the records and returns the report length. */
int hipac_profile_enable(int on);
long long hipac_profile_report(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* HIPAC_B200_H */
