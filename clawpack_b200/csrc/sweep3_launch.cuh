// Launch heuristics of the 3-D sweep kernels (grid shape, march-segment length) as templates over the Riemann functor:
// shared by the built-in instantiations (sweep3_inst.cu) and by user-supplied solvers compiled as plug-ins
// (rp_plugin.cu with -DCLAW_PLUGIN_NDIM=3).  Included inside a translation unit that defines CLAW_NS.
#pragma once
#include <cstdlib>
#include "launch.h"
#include "sweep3.cuh"

namespace clawcuda {
namespace CLAW_NS {

// Cells per march segment.  A launch runs tiles * nseg CTAs on `slots` resident ones: tiles * nseg / slots rounds of
// (len / nseg + warm) march steps each, plus on average half a round at the ragged end (warm = the cost of a segment's
// redundant warm-up steps in full steps: 2 planes for the x-sweep, ~1 for the marches, whose warm-up steps skip the
// transverse solves).  The nseg in 1..32 with the fewest steps wins; segments are not shorter than 16 cells.  Measured
// at 512^3 (profiles/r02_ze): one segment per tile 33.1 ms per step, 256-cell segments 31.3, 128 (this model's choice)
// 30.5, 64 30.7.
static int pick_seg(int len, long long tiles, int slots, int warm) {
  static const int forced = getenv("CLAW_SEG3") ? atoi(getenv("CLAW_SEG3")) : 0;      // A/B: cells per segment
  if (forced > 0) return forced < len ? forced : len;
  if (slots < 1) slots = 1;
  double best_cost = -1.0;
  int best_seg = len;
  for (int nseg = 1; nseg <= 32; ++nseg) {
    const int seg = (len + nseg - 1) / nseg;
    if (nseg > 1 && seg < 16) break;
    const long long ns = (len + seg - 1) / seg;
    const double rounds = (double)(tiles * ns) / slots;
    const double cost = ((rounds < 1.0 ? 1.0 : rounds) + 0.5) * (seg + warm);
    if (best_cost < 0.0 || cost < best_cost) { best_cost = cost; best_seg = seg; }
  }
  return best_seg < 1 ? 1 : best_seg;
}

template <class RP, int BY, int MINB, int SPEC = 0>
static cudaError_t launch_x(Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info) {
  constexpr int NT = 32 * BY;
  auto kern = sweep3_x<RP, BY, MINB, SPEC>;
  const int smem = (int)Smem3<RP, 0, NT>::BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const bool split = P.m3 < 0;
  const int nx = P.n[0] + 4, ny = P.n[1] + 4, nz = P.n[2] + 4;
  const int gx = (nx - 2 + 27) / 28;
  const int gy = split ? (ny + BY - 1) / BY : (P.n[1] + BY - 3) / (BY - 2);
  const int len = split ? nz : P.n[2];
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
  if (occ < 1) occ = 1;
  P.seg = pick_seg(len, (long long)gx * gy, num_sms * occ, 2);
  dim3 grid(gx, gy, (len + P.seg - 1) / P.seg), block(32, BY);
  kern<<<grid, block, smem, st>>>(P);
  if (info) { info->grid_x = grid.x; info->grid_y = grid.y * grid.z; info->block = NT; info->smem_bytes = smem; info->ctas_per_sm = occ; }
  return cudaGetLastError();
}

template <class RP, int IXYZ, int BX, int BY, int MINB, int SPEC = 0>
static cudaError_t launch_march(Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info) {
  constexpr int NT = BX * BY;
  auto kern = sweep3_march<RP, IXYZ, BX, BY, MINB, SPEC>;
  const int smem = (int)Smem3<RP, IXYZ, NT>::BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const bool split = P.m3 < 0;
  const int TD = (IXYZ == 1) ? 2 : 1;
  const int nx = P.n[0] + 4, nT = P.n[TD] + 4, nlen = P.n[IXYZ] + 4;
  const int gx = split ? (nx + BX - 1) / BX : (P.n[0] + BX - 3) / (BX - 2);
  const int gy = split ? (nT + BY - 1) / BY : (P.n[TD] + BY - 3) / (BY - 2);
  const int len = split ? nlen : P.n[IXYZ];
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
  if (occ < 1) occ = 1;
  P.seg = pick_seg(len, (long long)gx * gy, num_sms * occ, 1);
  dim3 grid(gx, gy, (len + P.seg - 1) / P.seg), block(BX, BY);
  kern<<<grid, block, smem, st>>>(P);
  if (info) { info->grid_x = grid.x; info->grid_y = grid.y * grid.z; info->block = NT; info->smem_bytes = smem; info->ctas_per_sm = occ; }
  return cudaGetLastError();
}

template <class RP, int BY, int MINB, int SPEC = 0>
static cudaError_t launch_dir(int dir, Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info) {
  switch (dir) {
    case 0: return launch_x<RP, BY, MINB, SPEC>(P, num_sms, st, info);
    case 1: return launch_march<RP, 1, 32, BY, MINB, SPEC>(P, num_sms, st, info);
    case 2: return launch_march<RP, 2, 32, BY, MINB, SPEC>(P, num_sms, st, info);
  }
  return cudaErrorInvalidValue;
}

}  // namespace CLAW_NS
}  // namespace clawcuda
