// nav_common.h — error plumbing shared by the translation units of libnavbot_b200.so.
#ifndef NAV_COMMON_H_
#define NAV_COMMON_H_

#include <cuda_runtime.h>

#include <string>

// Records `msg` as the calling thread's nav_last_error() and returns `code`.
int nav_fail(int code, const std::string& msg);

#define NAV_CUDA_TRY(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return nav_fail(NAVSIM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
  } while (0)

#endif  // NAV_COMMON_H_
