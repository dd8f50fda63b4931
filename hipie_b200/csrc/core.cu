// Error string, ABI version and launch counter shared by every entry point.
#include "common.cuh"
#include <stdarg.h>

namespace hipie {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace hipie

extern "C" const char* hipie_last_error(void) { return hipie::g_err; }
extern "C" int hipie_abi_version(void) { return 1; }
extern "C" int64_t hipie_launch_count(void) { return hipie::g_launch_count.load(); }
