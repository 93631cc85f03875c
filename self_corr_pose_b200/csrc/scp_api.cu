// Library-level entry points of the C ABI (include/scp_b200.h): version + error text.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/scp_b200.h"
#include "scp_common.cuh"

namespace scp {
static thread_local char g_err[512] = "";

void set_last_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace scp

extern "C" int scp_abi_version(void) { return 7; }
extern "C" const char *scp_last_error(void) { return scp::g_err; }
