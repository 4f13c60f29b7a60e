// ResNet18 layer table and packed-weight layout (host side).
// Network = torchvision resnet18 as wrapped by the reference (src/models/resnet.py:22-77).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace hipac {

struct ConvSpec {
  int cin, cout, k, stride, pad;  // square kernels
  int hin, hout;                  // square feature maps (per 224x224 patch)
};

// State-dict order: conv1; layerL.0.conv1, layerL.0.conv2, [layerL.0.downsample.0], layerL.1.conv1, layerL.1.conv2
static const ConvSpec kConvs[HIPAC_RESNET18_NUM_CONVS] = {
    {3, 64, 7, 2, 3, 224, 112},                                                         //  0 conv1
    {64, 64, 3, 1, 1, 56, 56},    {64, 64, 3, 1, 1, 56, 56},                            //  1, 2  layer1.0
    {64, 64, 3, 1, 1, 56, 56},    {64, 64, 3, 1, 1, 56, 56},                            //  3, 4  layer1.1
    {64, 128, 3, 2, 1, 56, 28},   {128, 128, 3, 1, 1, 28, 28}, {64, 128, 1, 2, 0, 56, 28},    //  5, 6, 7(ds)  layer2.0
    {128, 128, 3, 1, 1, 28, 28},  {128, 128, 3, 1, 1, 28, 28},                          //  8, 9  layer2.1
    {128, 256, 3, 2, 1, 28, 14},  {256, 256, 3, 1, 1, 14, 14}, {128, 256, 1, 2, 0, 28, 14},   // 10,11,12(ds) layer3.0
    {256, 256, 3, 1, 1, 14, 14},  {256, 256, 3, 1, 1, 14, 14},                          // 13,14  layer3.1
    {256, 512, 3, 2, 1, 14, 7},   {512, 512, 3, 1, 1, 7, 7},   {256, 512, 1, 2, 0, 14, 7},    // 15,16,17(ds) layer4.0
    {512, 512, 3, 1, 1, 7, 7},    {512, 512, 3, 1, 1, 7, 7},                            // 18,19  layer4.1
};

// GEMM K of the packed weights: conv1 is re-expressed as a 4x4 stride-1 conv over the 2x2
// space-to-depth input with 16 (12 used) channels -> K = 4*4*16 = 256; others K = k*k*cin.
static inline int conv_gemm_k(int l) { return l == 0 ? 256 : kConvs[l].k * kConvs[l].k * kConvs[l].cin; }

struct PackedLayout {
  size_t w_off[HIPAC_RESNET18_NUM_CONVS];  // bf16 [cout][K]
  size_t b_off[HIPAC_RESNET18_NUM_CONVS];  // f32 [cout]
  size_t fc_w_off, fc_b_off;               // f32 [k][512], f32 [k]
  // projection-shortcut fusion: for stage s (layer2..4) conv2 of block 0 and its 1x1/stride-2 downsample share one
  // accumulator: weights [cout][9*cout + cin_ds] = conv2's K followed by the downsample's K, bias = sum of both
  size_t wf_off[3], bf_off[3];
  size_t total;
};

static inline int fused_conv_layer(int s) { return 6 + 5 * s; }   // layerL.0.conv2
static inline int fused_ds_layer(int s) { return 7 + 5 * s; }     // layerL.0.downsample.0
static inline int fused_gemm_k(int s) { return conv_gemm_k(fused_conv_layer(s)) + conv_gemm_k(fused_ds_layer(s)); }

static inline PackedLayout packed_layout(int num_classes) {
  PackedLayout L;
  size_t off = 0;
  auto bump = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 255) / 256 * 256;
    return o;
  };
  for (int l = 0; l < HIPAC_RESNET18_NUM_CONVS; l++) {
    L.w_off[l] = bump((size_t)kConvs[l].cout * conv_gemm_k(l) * 2);
    L.b_off[l] = bump((size_t)kConvs[l].cout * 4);
  }
  L.fc_w_off = bump((size_t)(num_classes > 0 ? num_classes : 0) * 512 * 4);
  L.fc_b_off = bump((size_t)(num_classes > 0 ? num_classes : 0) * 4);
  for (int s = 0; s < 3; s++) {
    L.wf_off[s] = bump((size_t)kConvs[fused_conv_layer(s)].cout * fused_gemm_k(s) * 2);
    L.bf_off[s] = bump((size_t)kConvs[fused_conv_layer(s)].cout * 4);
  }
  L.total = off;
  return L;
}

}  // namespace hipac
