// Error string, ABI version and launch counter of the dgdm_b200 C ABI (include/dgdm_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace dgdm {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace dgdm

extern "C" const char* dgdm_last_error(void) { return dgdm::g_err; }
extern "C" int dgdm_abi_version(void) { return DGDM_ABI_VERSION; }
extern "C" uint64_t dgdm_launch_count(void) { return dgdm::g_launches.load(); }
