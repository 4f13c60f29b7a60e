// Types shared by the direct and fused stage-1 kernels.
#pragma once
#include "common.cuh"

namespace hipac {

struct CoeffSet {       // Pillow 22-bit fixed-point triangle weights for one integer scale
  int32_t interior[16]; // 2*scale taps
  int32_t left[12];     // output 0   (3*scale/2 taps)
  int32_t right[12];    // output 223 (3*scale/2 taps)
};

struct ScanParams {
  const uint8_t* rgb;   // [H][pitch] RGB bytes
  int H, W;
  int64_t pitch;
  const uint8_t* mask;  // [H][mask_pitch] or null
  int64_t mask_pitch;
  int P, S;             // patch size, stride
  int nx, ny;           // candidate grid of this call: nx columns, ny rows starting at iy_begin
  int iy_begin;
};

struct OutParams {
  uint8_t* batch_u8;    // [cap][224][224][3] or null
  uint16_t* batch;      // bf16 bits, layout below, or null
  int layout;
};

}  // namespace hipac
