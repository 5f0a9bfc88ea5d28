// Shared host-side helpers for the C ABI: error reporting and launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/t2v_b200.h"

namespace t2v {
// Records a formatted thread-local error message and returns `code` (negative).
int fail(int code, const char* fmt, ...);
// Maps a cudaError_t (as int) from a launch to the ABI convention; counts the launch on success.
int launch_checked(int cuda_err, const char* what);
void count_launch(int n = 1);
}  // namespace t2v
