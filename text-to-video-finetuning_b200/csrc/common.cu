#include "common.h"

#include <cstdlib>

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace t2v {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code < 0 ? code : -code - 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int launch_checked(int cuda_err, const char* what) {
    if (cuda_err != 0) {
        return fail(-100 - cuda_err, "%s: CUDA launch failed: %s", what, cudaGetErrorString(static_cast<cudaError_t>(cuda_err)));
    }
    count_launch();
    return 0;
}

const char* last_error() { return g_err; }
int64_t launches() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = std::getenv("T2V_NO_PDL") == nullptr;
    return on;
}
}  // namespace t2v

namespace t2v {
const char* last_error();
int64_t launches();
}

extern "C" {
int t2v_version(void) { return 2; }
const char* t2v_last_error(void) { return t2v::last_error(); }
int64_t t2v_launch_count(void) { return t2v::launches(); }
int64_t t2v_stream_capture_id(void* stream) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo(static_cast<cudaStream_t>(stream), &status, &id) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return status == cudaStreamCaptureStatusActive ? static_cast<int64_t>(id) : 0;
}
}
