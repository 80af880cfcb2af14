// Error reporting and version for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fmhr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace fmhr

extern "C" int fmhr_version(void) { return 100; }
extern "C" const char* fmhr_last_error_string(void) { return fmhr::g_err; }
