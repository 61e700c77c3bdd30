// Error plumbing and bookkeeping shared by every entry point of libga_b200.
#include "ga_common.cuh"

namespace ga {
static thread_local char g_err[1024] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { ++g_launches; }
}  // namespace ga

extern "C" const char* ga_last_error(void) { return ga::g_err; }
extern "C" int ga_abi_version(void) { return GA_ABI_VERSION; }
extern "C" int64_t ga_launch_count(int reset) {
  int64_t v = ga::g_launches;
  if (reset) ga::g_launches = 0;
  return v;
}
