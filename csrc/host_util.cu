#include <stdio.h>
#include <string.h>

#include "launchers.h"

static thread_local char g_err[1024] = "";

void ztb_set_error(const char* expr, const char* msg, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s:%d)", expr, msg, file, line);
}
const char* ztb_last_error() { return g_err; }
