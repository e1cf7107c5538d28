// Error plumbing + version for the davi C ABI (include/davi.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "davi_common.cuh"

namespace davi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the (non-sticky) launch error
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return DAVI_ECUDA;
    }
    return DAVI_OK;
}

}  // namespace davi

extern "C" int davi_version(void) { return DAVI_VERSION; }

extern "C" int davi_last_error(char* buf, size_t buf_bytes) {
    if (!buf || buf_bytes == 0) return DAVI_EINVAL;
    strncpy(buf, davi::g_err, buf_bytes - 1);
    buf[buf_bytes - 1] = '\0';
    return DAVI_OK;
}
