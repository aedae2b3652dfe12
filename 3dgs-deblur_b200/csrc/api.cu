// ABI version + thread-local error string of libb200splat.
#include <stdarg.h>

#include "common.cuh"

namespace b200 {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b200

extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" const char *b200_last_error(void) { return b200::g_err; }
