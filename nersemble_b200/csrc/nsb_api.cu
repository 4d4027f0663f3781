// libnsb: error handling and version (C ABI glue).
#include <stdarg.h>
#include <stdio.h>

#include "nsb_common.cuh"

namespace nsb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

}  // namespace nsb

extern "C" int nsb_version(void) { return NSB_VERSION; }
extern "C" const char *nsb_last_error(void) { return nsb::g_err; }
