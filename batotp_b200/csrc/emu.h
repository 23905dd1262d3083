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
#include <mutex>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
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

// One emulated launch at a time: kernels keep their "shared memory" in function-local statics and the fibre stacks
// come from one pool, so host threads that drive different contexts (the tail chunk of a batch) take turns.
static inline std::recursive_mutex &emu_launch_mutex() {
  static std::recursive_mutex m;
  return m;
}
template <class F>
static inline void emu_launch(dim3 grid, dim3 block, F f) {
  std::lock_guard<std::recursive_mutex> emu_lk(emu_launch_mutex());
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

// ---- warp-cooperative kernels -------------------------------------------------------------------
// Kernels that use warp collectives (ballot / shuffle) cannot be run one thread after another.  For
// them every thread of a CTA is a fibre (ucontext); a fibre that reaches a collective parks until all
// lanes named in the mask have arrived, then every participant reads the deposited values.  The
// round-robin scheduler below only decides the interleaving, which collectives make immaterial.
#include <vector>
#define EMU_WARP 32
#if defined(__x86_64__)
// minimal System-V context switch (callee-saved registers + stack pointer); glibc's swapcontext
// costs a signal-mask system call per switch, far too slow for millions of collectives
extern "C" void emu_ctx_switch(void **save_sp, void *next_sp);
__asm__(
    ".text\n.globl emu_ctx_switch\n.type emu_ctx_switch,@function\nemu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_ctx_switch,.-emu_ctx_switch\n");
struct EmuCtx {
  void *sp = nullptr;
};
static inline void emu_ctx_make(EmuCtx &c, char *stack, size_t size, void (*entry)()) {
  uintptr_t top = ((uintptr_t)stack + size) & ~(uintptr_t)15;
  void **p = (void **)top;
  *--p = nullptr;         // keeps the entry frame 16-byte aligned as after a call
  *--p = (void *)entry;   // popped by `ret`
  for (int i = 0; i < 6; ++i) *--p = nullptr;
  c.sp = p;
}
static inline void emu_ctx_swap(EmuCtx &from, EmuCtx &to) { emu_ctx_switch(&from.sp, to.sp); }
#else
#include <ucontext.h>
struct EmuCtx {
  ucontext_t u;
};
static inline void emu_ctx_make(EmuCtx &c, char *stack, size_t size, void (*entry)()) {
  getcontext(&c.u);
  c.u.uc_stack.ss_sp = stack;
  c.u.uc_stack.ss_size = size;
  c.u.uc_link = nullptr;
  makecontext(&c.u, entry, 0);
}
static inline void emu_ctx_swap(EmuCtx &from, EmuCtx &to) { swapcontext(&from.u, &to.u); }
#endif
struct EmuWarpState {
  unsigned arrived = 0, done = 0;
  unsigned long long slot[EMU_WARP];
};
struct EmuFiber {
  EmuCtx ctx;
  char *stack = nullptr;  // from the per-thread pool below: allocated once, reused by every CTA
  emu_dim3 tid;
  bool finished = false;
};
#define EMU_STACK (1 << 19)
static inline char *emu_stack(int t) {
  static thread_local std::vector<char *> pool;
  while ((int)pool.size() <= t) pool.push_back((char *)malloc(EMU_STACK));
  return pool[t];
}
struct EmuCta {
  EmuCtx sched;
  std::vector<EmuFiber> fib;
  std::vector<EmuWarpState> warps;
  int cur = 0;
  unsigned ctaArrived = 0, ctaGen = 0;
  void (*body)(void *) = nullptr;
  void *arg = nullptr;
};
extern thread_local EmuCta *emu_cta;
static inline void emu_yield() { emu_ctx_swap(emu_cta->fib[emu_cta->cur].ctx, emu_cta->sched); }
static void emu_fiber_entry() {
  EmuCta *c = emu_cta;
  c->body(c->arg);
  c->fib[c->cur].finished = true;
  for (;;) emu_yield();
}
template <class F>
static inline void emu_launch_fibers(dim3 grid, dim3 block, F f) {
  std::lock_guard<std::recursive_mutex> emu_lk(emu_launch_mutex());
  gridDim = grid;
  blockDim = block;
  const int nt = (int)(block.x * block.y);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        EmuCta cta;
        cta.fib.resize(nt);
        cta.warps.resize((nt + EMU_WARP - 1) / EMU_WARP);
        cta.body = [](void *a) { (*static_cast<F *>(a))(); };
        cta.arg = &f;
        emu_cta = &cta;
        for (int t = 0; t < nt; ++t) {
          EmuFiber &fb = cta.fib[t];
          fb.stack = emu_stack(t);
          fb.tid = emu_dim3(t % block.x, t / block.x, 0);  // linear thread id = ty * blockDim.x + tx, warps of 32
          emu_ctx_make(fb.ctx, fb.stack, EMU_STACK, emu_fiber_entry);
        }
        int live = nt;
        while (live > 0) {
          live = 0;
          for (int t = 0; t < nt; ++t) {
            if (cta.fib[t].finished) continue;
            ++live;
            cta.cur = t;
            blockIdx = emu_dim3(bx, by, bz);
            threadIdx = cta.fib[t].tid;
            emu_ctx_swap(cta.sched, cta.fib[t].ctx);
          }
        }
        emu_cta = nullptr;
      }
}
// deposit `v`, wait for every lane in `mask`, return a pointer to the 32 deposited values
static inline const unsigned long long *emu_rendezvous(unsigned mask, unsigned long long v) {
  EmuCta *c = emu_cta;
  const int t = c->cur, lane = t % EMU_WARP;
  EmuWarpState &w = c->warps[t / EMU_WARP];
  const unsigned bit = 1u << lane;
  while (w.arrived & bit) emu_yield();  // the previous collective of this lane has not drained yet
  w.slot[lane] = v;
  w.arrived |= bit;
  while ((w.arrived & mask) != mask) emu_yield();
  return w.slot;
}
static inline void emu_release(unsigned mask) {
  EmuCta *c = emu_cta;
  const int t = c->cur, lane = t % EMU_WARP;
  EmuWarpState &w = c->warps[t / EMU_WARP];
  w.done |= 1u << lane;
  if ((w.done & mask) == mask) {  // last reader frees the slots of this collective
    w.done &= ~mask;
    w.arrived &= ~mask;
  }
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  const unsigned long long *s = emu_rendezvous(mask, pred ? 1ull : 0ull);
  unsigned r = 0;
  for (int l = 0; l < EMU_WARP; ++l)
    if ((mask >> l) & 1u) r |= (unsigned)(s[l] & 1ull) << l;
  emu_release(mask);
  return r;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)__ballot_sync(mask, 0); }
template <class T>
static inline T emu_shfl_idx(unsigned mask, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  const unsigned long long *s = emu_rendezvous(mask, bits);
  const unsigned long long got = s[src & (EMU_WARP - 1)];
  emu_release(mask);
  T o;
  memcpy(&o, &got, sizeof(T));
  return o;
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src) { return emu_shfl_idx(mask, v, src); }
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
  const int lane = emu_cta->cur % EMU_WARP;
  return emu_shfl_idx(mask, v, lane >= (int)delta ? lane - (int)delta : lane);
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask) {
  return emu_shfl_idx(mask, v, (emu_cta->cur % EMU_WARP) ^ lanemask);
}
static inline void __syncthreads() {
  EmuCta *c = emu_cta;
  if (!c) return;  // sequential emulation: nothing to wait for
  const unsigned gen = c->ctaGen;
  int live = 0;
  for (auto &f : c->fib) live += f.finished ? 0 : 1;
  if (++c->ctaArrived >= (unsigned)live) {
    c->ctaArrived = 0;
    c->ctaGen++;
    return;
  }
  while (c->ctaGen == gen) emu_yield();
}
#define BATOTP_LAUNCH_WARP(kern, grid, block, smem, stream, ...) \
  emu_launch_fibers((grid), (block), [&] { kern(__VA_ARGS__); })
#define EMU_SHARED static
#else
#define BATOTP_LAUNCH(kern, grid, block, stream, ...) kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define BATOTP_LAUNCH_SMEM(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BATOTP_LAUNCH_WARP(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define EMU_SHARED __shared__
#endif
