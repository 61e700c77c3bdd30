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
static const uint64_t* g_seed_salt = nullptr;
const uint64_t* seed_salt_ptr() { return g_seed_salt; }
__global__ void seed_salt_bump_kernel(uint64_t* salt) { *salt += 0x9E3779B97F4A7C15ull; }
static int g_round_tf32 = 0;
int round_tf32_enabled() { return g_round_tf32; }
}  // namespace ga

extern "C" const char* ga_last_error(void) { return ga::g_err; }
extern "C" int ga_abi_version(void) { return GA_ABI_VERSION; }
extern "C" int64_t ga_launch_count(int reset) {
  int64_t v = ga::g_launches;
  if (reset) ga::g_launches = 0;
  return v;
}

extern "C" int ga_seed_salt_set(uint64_t* dev_salt) {
  ga::g_seed_salt = dev_salt;
  return 0;
}
extern "C" int ga_seed_salt_bump(void* stream) {
  GA_CHECK(ga::g_seed_salt != nullptr, "ga_seed_salt_bump: no salt buffer registered (ga_seed_salt_set)");
  ga::seed_salt_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(const_cast<uint64_t*>(ga::g_seed_salt));
  GA_LAUNCH_OK();
  return 0;
}

extern "C" int ga_f32_round_tf32(int on) {
  ga::g_round_tf32 = on ? 1 : 0;
  return 0;
}
