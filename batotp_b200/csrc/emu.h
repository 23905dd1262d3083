// emu.h — lets the kernel sources compile as plain C++ for CPU-only CI.
//
// This is NOT a product path and is never part of libbatotp_cuda.so: the shipped
// library is built by nvcc without BATOTP_HOST_EMU and fails loudly when no CUDA
// device is present.  With -DBATOTP_HOST_EMU (tests/emu build only, g++) every
// __global__ kernel body is run sequentially, one "thread" at a time, so the exact
// per-thread logic that runs on the B200 can be checked against the oracle inside
// the CPU-only container before any GPU minute is spent.
#pragma once

#ifdef BATOTP_HOST_EMU
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __constant__

struct emu_dim3 {
  unsigned x, y, z;
  emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef emu_dim3 dim3;
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
typedef int cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

template <class F>
static inline void emu_launch(dim3 grid, dim3 block, F f) {
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned ty = 0; ty < block.y; ++ty)
          for (unsigned tx = 0; tx < block.x; ++tx) {
            blockIdx = emu_dim3(bx, by, bz);
            threadIdx = emu_dim3(tx, ty, 0);
            f();
          }
}
#define BATOTP_LAUNCH(kern, grid, block, stream, ...) emu_launch((grid), (block), [&] { kern(__VA_ARGS__); })
#define BATOTP_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) emu_launch((grid), (block), [&] { kern(__VA_ARGS__); })

static inline int atomicAdd(int *p, int v) {
  int o = *p;
  *p = o + v;
  return o;
}
static inline int atomicOr(int *p, int v) {
  int o = *p;
  *p = o | v;
  return o;
}
static inline int atomicMax(int *p, int v) {
  int o = *p;
  if (v > o) *p = v;
  return o;
}
#else
#define BATOTP_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define BATOTP_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
