// Error plumbing + version for the davi C ABI (include/davi.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

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

// process-wide implementation switches (davi_set_option): relaxed atomics, read on every call
static std::atomic<int> g_options[DAVI_OPT_COUNT_];

int get_option(int option) {
    return (option >= 0 && option < DAVI_OPT_COUNT_) ? g_options[option].load(std::memory_order_relaxed) : 0;
}

}  // namespace davi

extern "C" int davi_set_option(int option, int value) {
    if (option < 0 || option >= DAVI_OPT_COUNT_) {
        davi::set_error("davi_set_option: unknown option %d", option);
        return DAVI_EINVAL;
    }
    davi::g_options[option].store(value, std::memory_order_relaxed);
    return DAVI_OK;
}

extern "C" int davi_get_option(int option) { return davi::get_option(option); }

extern "C" int davi_version(void) { return DAVI_VERSION; }

extern "C" int davi_last_error(char* buf, size_t buf_bytes) {
    if (!buf || buf_bytes == 0) return DAVI_EINVAL;
    strncpy(buf, davi::g_err, buf_bytes - 1);
    buf[buf_bytes - 1] = '\0';
    return DAVI_OK;
}
