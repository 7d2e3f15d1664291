// Shared host-side helpers: error codes, thread-local error detail, launch counting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/prl_b200.h"

namespace prl {

struct ErrState {
  char detail[512];
  int64_t launches;
};
ErrState& err_state();

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(err_state().detail, sizeof(err_state().detail), fmt, a, b, c);
  return code;
}
inline int fail_cuda(cudaError_t e, const char* where) {
  snprintf(err_state().detail, sizeof(err_state().detail), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return PRL_ERR_CUDA;
}
inline void count_launch(int n = 1) { err_state().launches += n; }

// check the launch that just happened; returns PRL_OK or PRL_ERR_CUDA
inline int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, where);
  count_launch();
  return PRL_OK;
}

int sm_count();  // cached multiprocessor count of the current device (148 on B200)

#define PRL_REQUIRE(cond, code, msg) \
  do {                                \
    if (!(cond)) return ::prl::fail((code), "%s", (msg)); \
  } while (0)

}  // namespace prl
