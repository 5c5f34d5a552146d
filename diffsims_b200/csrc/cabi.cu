// C-ABI housekeeping: version and the per-thread error string.
#include <stdarg.h>

#include "common.cuh"

namespace ds {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace ds

extern "C" int ds_abi_version(void) { return DS_ABI_VERSION; }
extern "C" const char *ds_last_error(void) { return ds::g_err; }
