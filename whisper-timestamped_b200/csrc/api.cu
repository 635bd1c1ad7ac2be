// libwts: error reporting and version entry points of the C-ABI (include/wts.h).
#include <stdarg.h>

#include "common.cuh"

namespace wts {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace wts

extern "C" int wts_version(void) { return WTS_VERSION; }

extern "C" const char* wts_last_error(void) { return wts::g_err; }
