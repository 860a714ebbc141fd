// Host-side plumbing shared by the translation units of libflowmap_b200.so: the thread-local
// error string behind fm_last_error() and the launch counter behind fm_launch_count().
#pragma once
#include <cuda_runtime.h>

namespace fm_host {
int fail(const char* what, cudaError_t e);  // records "<what>: <cuda error>", returns 1
int fail_msg(const char* what);             // records <what>, returns 2
void count_launch();
const char* last_error();
unsigned long long launches();
}  // namespace fm_host

#define FM_CHECK_LAUNCH(name)                                   \
  do {                                                          \
    cudaError_t e_ = cudaGetLastError();                        \
    if (e_ != cudaSuccess) return fm_host::fail(name, e_);      \
    fm_host::count_launch();                                    \
  } while (0)
