// Shared host/device helpers for libcellseg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cellseg_b200.h"

namespace cs {

// Records the text returned by cs_last_error() (thread local).
void set_error(const char* fmt, ...);
const char* last_error();

// SM count of the current device, cached per device (B200: 2 dies x 74 SMs = 148).
int num_sms();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Number of grid positions along one axis; dataset/dataset.py:728-740.
__host__ __device__ inline int grid_count(int dim, int tile, int interval) {
  if (dim < tile || tile <= 0 || interval <= 0) return 0;
  int span = dim - tile;
  return span / interval + 1 + ((span % interval) != 0 ? 1 : 0);
}
// Coordinate of grid position g: 0, I, 2I, ..., then dim - S.
__host__ __device__ inline int grid_coord(int g, int dim, int tile, int interval) {
  int c = g * interval;
  int last = dim - tile;
  return c < last ? c : last;
}

// Grid positions whose tile covers coordinate c along one axis: the contiguous range [*lo, *hi]
// (position coordinates are monotonic); empty (*lo > *hi) for a pixel in the gap between tiles
// when the stride exceeds the tile.  gcnt = grid_count(dim, S, I) > 0, 0 <= c < dim.
__host__ __device__ inline void grid_cover(int c, int dim, int S, int I, int gcnt, int* lo, int* hi) {
  const int last = dim - S;
  int h = c / I;
  if (h > gcnt - 1 || c >= last) h = gcnt - 1;
  int l = c - S + 1 <= 0 ? 0 : (c - S + I) / I;            // ceil((c - S + 1) / I)
  if (l > gcnt - 1) l = gcnt - 1;
  *lo = l;
  *hi = h;
}

// float -> uint32 key whose unsigned order is numpy's sort order for float32:
// ascending value, -0.0 == +0.0, every NaN equal and last.
__host__ __device__ inline uint32_t float_sort_key(float f) {
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cvt; cvt.f = f; b = cvt.u;
#endif
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN
  if (b == 0x80000000u) b = 0u;                            // -0.0 -> +0.0
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

}  // namespace cs

#define CS_CUDA(expr)                                                               \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      cs::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,              \
                    cudaGetErrorString(e__));                                       \
      return CS_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define CS_REQUIRE(cond, ...)                                                       \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      cs::set_error(__VA_ARGS__);                                                   \
      return CS_ERR_INVALID_ARG;                                                    \
    }                                                                               \
  } while (0)

#define CS_LAUNCH_CHECK() CS_CUDA(cudaGetLastError())

namespace cs {
// Kernel launch with the programmatic-dependent-launch attribute (and an optional 1-D cluster).
// Kernels launched this way call pdl_wait() before reading upstream data.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace cs
