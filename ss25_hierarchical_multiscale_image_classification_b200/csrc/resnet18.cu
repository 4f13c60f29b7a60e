// Stage 2 (ResNet18 forward) -- placeholder until the tcgen05 kernels land.
#include "common.cuh"
using namespace hipac;
extern "C" size_t hipac_resnet18_packed_bytes(int) { return 0; }
extern "C" int hipac_resnet18_pack(const float* const*, int, int, float, void*, size_t) { set_error("stage 2 not built"); return -4; }
extern "C" size_t hipac_resnet18_workspace_bytes(int, int) { return 0; }
extern "C" int hipac_resnet18_forward(const void*, int, const void*, int, int, float*, float*, void*, size_t, int, void*) { set_error("stage 2 not built"); return -4; }
extern "C" int hipac_resnet18_conv_layer(const void*, int, int, const void*, const void*, void*, int, int, void*) { set_error("stage 2 not built"); return -4; }
