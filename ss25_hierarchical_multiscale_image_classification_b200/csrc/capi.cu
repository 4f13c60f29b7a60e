// Error reporting and bookkeeping for the C ABI.
#include "common.cuh"

namespace hipac {
static thread_local std::string g_last_error;
static thread_local long long g_launches = 0;
void set_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches += n; }
}  // namespace hipac

extern "C" const char* hipac_last_error(void) { return hipac::g_last_error.c_str(); }
extern "C" int hipac_abi_version(void) { return HIPAC_ABI_VERSION; }
extern "C" long long hipac_launch_count(int reset) {
  long long v = hipac::g_launches;
  if (reset) hipac::g_launches = 0;
  return v;
}
