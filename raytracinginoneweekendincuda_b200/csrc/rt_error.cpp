// rt_error.cpp -- thread-local last-error string behind rt_last_error().
// Replaces the reference's print + cudaDeviceReset() + exit(99) convention
// (reference kernel.cu:29-40): entry points return a status and never exit.
#include <cstdarg>
#include <cstdio>

#include "../../include/rt_abi.h"

static thread_local char g_err[1024] = "";

void rt_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char* rt_last_error(void) { return g_err; }
