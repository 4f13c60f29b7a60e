// Shared host/device helpers for libhipac_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/hipac_b200.h"

namespace hipac {

void set_error(const std::string& msg);
void count_launch(int n = 1);

// RAII CUDA-event bracket around one kernel launch; a no-op unless hipac_profile_enable(1).
class ProfileScope {
 public:
  ProfileScope(const char* name, cudaStream_t stream, double work = 0.0);
  ~ProfileScope();
 private:
  int idx_;
  cudaStream_t stream_;
};

#define HIPAC_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::hipac::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
      return -2;                                                                            \
    }                                                                                       \
  } while (0)

#define HIPAC_REQUIRE(cond, msg)                                     \
  do {                                                               \
    if (!(cond)) {                                                   \
      ::hipac::set_error(std::string("invalid argument: ") + (msg)); \
      return -1;                                                     \
    }                                                                \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int OUT = 224;  // network input size (reference src/main.py:814)

}  // namespace hipac
