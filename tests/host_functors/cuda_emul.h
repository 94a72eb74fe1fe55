// TEST INFRASTRUCTURE.  A minimal single-process emulation of the CUDA execution model, enough to run the sweep kernels'
// SOURCE on the CPU for small grids: every thread of a block is a user-level fiber (ucontext); the fibers of a block are
// resumed round-robin and each runs until its next synchronisation point (__syncthreads or a warp shuffle), so that in
// round k every fiber has executed up to its k-th synchronisation point -- which is all that warp-synchronous / barrier
// code relies on.  Shuffles exchange through a double-buffered per-block array indexed by a per-fiber phase counter.
// Blocks run one after the other.  Only what the kernels under test use is provided.  Nothing in the product includes this.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace emul {
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
constexpr int MAX_THREADS = 256;
inline Dim3 g_tid[MAX_THREADS], g_bid, g_bdim, g_gdim;
inline int g_cur = 0;                       // fiber (= linear thread id) that is running
inline int g_n = 0;
inline ucontext_t g_sched, g_ctx[MAX_THREADS];
inline bool g_done[MAX_THREADS];
inline unsigned g_phase[MAX_THREADS];
inline double g_xbuf[2][MAX_THREADS];
inline std::vector<std::vector<char>> g_stacks;
inline std::function<void()> g_body;
inline double g_dyn_smem[1 << 17];          // dynamic shared memory of the running block (1 MiB)

inline void yield_sync() { swapcontext(&g_ctx[g_cur], &g_sched); }
inline void trampoline() { g_body(); g_done[g_cur] = true; swapcontext(&g_ctx[g_cur], &g_sched); }

// run `body` once per thread of every block of the grid
inline void launch(Dim3 grid, Dim3 block, const std::function<void()>& body) {
  g_gdim = grid; g_bdim = block; g_body = body;
  g_n = (int)(block.x * block.y * block.z);
  if ((int)g_stacks.size() < g_n) g_stacks.resize(g_n, std::vector<char>(1 << 19));
  for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
    g_bid.x = bx; g_bid.y = by; g_bid.z = bz;
    for (int t = 0; t < g_n; ++t) {
      g_tid[t].x = t % block.x; g_tid[t].y = (t / block.x) % block.y; g_tid[t].z = t / (block.x * block.y);
      g_done[t] = false; g_phase[t] = 0;
      getcontext(&g_ctx[t]);
      g_ctx[t].uc_stack.ss_sp = g_stacks[t].data();
      g_ctx[t].uc_stack.ss_size = g_stacks[t].size();
      g_ctx[t].uc_link = &g_sched;
      makecontext(&g_ctx[t], (void (*)())trampoline, 0);
    }
    for (bool any = true; any;) {
      any = false;
      for (int t = 0; t < g_n; ++t)
        if (!g_done[t]) { any = true; g_cur = t; swapcontext(&g_sched, &g_ctx[t]); }
    }
  }
}

inline double shfl(double v, int src_of_lane(int lane, int arg), int arg) {
  const int me = g_cur, base = me & ~31, lane = me & 31;
  const unsigned ph = g_phase[me]++ & 1u;
  g_xbuf[ph][me] = v;
  yield_sync();
  return g_xbuf[ph][base + src_of_lane(lane, arg)];
}
// warp vote: true iff the predicate holds on all 32 lanes of the calling fiber's warp (all of them reach the call)
inline bool vote_all(bool pred) {
  const int me = g_cur, base = me & ~31;
  const unsigned ph = g_phase[me]++ & 1u;
  g_xbuf[ph][me] = pred ? 1.0 : 0.0;
  yield_sync();
  bool all = true;
  for (int l = 0; l < 32 && base + l < g_n; ++l) all = all && (g_xbuf[ph][base + l] != 0.0);
  return all;
}
}  // namespace emul

#define __global__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define threadIdx (emul::g_tid[emul::g_cur])
#define blockIdx emul::g_bid
#define blockDim emul::g_bdim
#define gridDim emul::g_gdim
using std::min; using std::max;
typedef void* cudaStream_t;

inline void __syncthreads() { emul::yield_sync(); }
inline double __shfl_up_sync(unsigned, double v, int d) {
  return emul::shfl(v, [](int lane, int dd) { return lane >= dd ? lane - dd : lane; }, d);
}
inline double __shfl_down_sync(unsigned, double v, int d) {
  return emul::shfl(v, [](int lane, int dd) { return lane + dd <= 31 ? lane + dd : lane; }, d);
}
inline double __shfl_xor_sync(unsigned, double v, int o) {
  return emul::shfl(v, [](int lane, int oo) { return lane ^ oo; }, o);
}
inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, 8); return r; }
inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
  const unsigned long long old = *a; if (v > old) *a = v; return old;
}
inline unsigned atomicAdd(unsigned* a, unsigned v) { const unsigned old = *a; *a = old + v; return old; }
inline void __threadfence() {}
