// Shared host-side helpers for the aeroapps_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define AB_CUDA_OK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ab200_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
      return AB200_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

enum { AB200_ERR_CUDA = -1, AB200_ERR_ARG = -2 };

void ab200_set_error(const std::string &msg);

// grow-only device workspace
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T *as() { return static_cast<T *>(p); }
};
