// Launch heuristics of the 2-D sweep kernels (grid shape, march-segment length, tile ordering) as templates over the
// kernel / the Riemann functor: shared by the built-in instantiations (sweep_inst.cu) and by user-supplied Riemann
// solvers compiled as plug-ins (rp_plugin.cu).  Included inside a translation unit that defines CLAW_NS.
#pragma once
#include <cstdlib>
#include "launch.h"
#include "sweep2_ws.cuh"
#include "sweep2_aux.cuh"

namespace clawcuda {
namespace CLAW_NS {

template <class KERN>
static cudaError_t launch_kernel(KERN kern, int NT, int smem, bool x_only, Sweep2Params& P, int seg_hint, int num_sms,
                                 cudaStream_t st, LaunchInfo* info, int cols_per_cta = 0, int max_seg = 256,
                                 bool hot_ok = false) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int occ = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int TX = cols_per_cta > 0 ? cols_per_cta : NT - 4;
  const int strips = (P.mx + TX - 1) / TX;
  if (P.j1 <= P.j0) {                         // default: the whole grid
    P.j0 = x_only ? 0 : 2;
    P.j1 = x_only ? P.my + 4 : P.my + 2;
  }
  const int rows = P.j1 - P.j0;
  int seg = seg_hint;
  if (seg <= 0) {
    // enough CTAs for >= 6 waves of resident CTAs, but segments not shorter than 32 rows
    // (each segment pays 3-4 warm-up rows) nor longer than 256.  Small grids that cannot even fill the GPU once with
    // 32-row segments are latency-bound on the march length instead: there the segments are as short as one wave of CTAs
    // allows (the redundant warm-up rows cost nothing on an otherwise idle GPU), down to a single row.
    const long long want = 6LL * num_sms * occ;
    long long nseg = (want + strips - 1) / strips;
    if (nseg < 1) nseg = 1;
    seg = (int)((rows + nseg - 1) / nseg);
    int min_seg = 32;
    if ((long long)strips * ((rows + 31) / 32) < (long long)num_sms * occ) {
      const long long one_wave = ((long long)rows * strips + (long long)num_sms * occ - 1) / ((long long)num_sms * occ);
      min_seg = one_wave < 1 ? 1 : (int)one_wave;
    }
    if (seg < min_seg) seg = min_seg;
    if (seg > max_seg) seg = max_seg;
  }
  if (seg > rows) seg = rows;
  P.seg_rows = seg;
  static const int tune = getenv("CLAW_TUNE") ? atoi(getenv("CLAW_TUNE")) : 0;    // A/B switches (Sweep2Params::tune)
  P.tune = tune;
  dim3 grid(strips, (rows + seg - 1) / seg);
  // tile ordering (sweep2_ws.cuh, HotTiles): extra grid rows in front of the regular grid run last step's heavy tiles
  // first; worth it only when the launch is several waves of CTAs long
  P.hot_rows = 0;
  if (hot_ok && P.hot != nullptr && !(tune & 2) && !x_only) {
    const long long ntiles = (long long)grid.x * grid.y;
    int hy = (int)grid.y / 16;
    if (hy < 1) hy = 1;
    while (hy > 0 && (long long)hy * grid.x > P.hot_cap) --hy;
    // (CLAW_TUNE bit 5: also on launches shorter than four waves -- lets the parity tests exercise the ordering on grids
    // the oracle finishes quickly)
    if (hy > 0 && ntiles <= P.hot_tiles && (ntiles >= 4LL * num_sms * occ || (tune & 32)) && grid.x < 65536 &&
        grid.y + hy < 65536) {
      P.hot_rows = hy;
      grid.y += hy;
    }
  }
  kern<<<grid, NT, smem, st>>>(P);
  if (info) { info->grid_x = grid.x; info->grid_y = grid.y; info->block = NT; info->smem_bytes = smem; info->ctas_per_sm = occ; }
  return cudaGetLastError();
}

template <class RP, int NT, int MINB = 1>
static cudaError_t launch_one(int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st,
                              LaunchInfo* info) {
  using S = Smem2<RP, NT>;
  auto kern = (kind == 0) ? sweep2_unsplit<RP, NT, MINB> : sweep2_split<RP, NT, MINB>;
  return launch_kernel(kern, NT, (int)S::BYTES, (kind == 1) && !P.do_y, P, seg_hint, num_sms, st, info);
}

// warp-synchronous unsplit kernel (sweep2_ws.cuh): NW independent 28-column strips per CTA
template <class RP, int NW, int MINB, int SPEC, int UNR = 1, int QD = 0, int TMA = 0>
static cudaError_t launch_ws(Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st, LaunchInfo* info) {
  using S = Smem2ws<RP, NW, (QD > 0 ? QD : 2)>;
  // with the uniform-region fast path the cost of a CTA depends on how much of a front it holds: 64-row segments keep
  // the slowest CTAs short enough for the tail of the launch (measured at 8192^2, profiles/r02_i_segrows.jsonl: 1.38 ms
  // with 256-row segments, 1.20 / 1.16 / 1.29 ms with 128 / 64 / 32)
  return launch_kernel(sweep2_unsplit_ws<RP, NW, MINB, SPEC, UNR, QD, TMA>, NW * 32, (int)S::BYTES, false, P, seg_hint, num_sms,
                       st, info, NW * 28, QD > 0 ? 64 : 256, UnsplitQuiet<RP, SPEC, QD>::value && NW == 1);
}

// warp-synchronous dimension-split kernel (sweep2_ws.cuh)
template <class RP, int NW, int MINB, int PF = 1, int UNR = 1, int LIM = 0, int QD = 0>
static cudaError_t launch_split_ws(Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st, LaunchInfo* info) {
  using S = Smem2ws<RP, NW, (QD > 0 ? QD : 2)>;
  return launch_kernel(sweep2_split_ws<RP, NW, MINB, PF, UNR, LIM, QD>, NW * 32, (int)S::BYTES, !P.do_y, P, seg_hint, num_sms,
                       st, info, NW * 28, QD > 0 ? 32 : 256, SplitQuiet<RP, PF, LIM, QD>::value && NW == 1 && P.do_x && P.do_y);
}

// aux-reading solvers (sweep2_aux.cuh): warp-synchronous, NW strips per CTA
template <class RP, int NW>
static cudaError_t launch_aux(int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st, LaunchInfo* info) {
  using S = Smem2aux<RP, NW>;
  if (kind == 0)
    return launch_kernel(sweep2_unsplit_aux<RP, NW>, NW * 32, (int)S::BYTES, false, P, seg_hint, num_sms, st, info, NW * 28);
  return launch_kernel(sweep2_split_aux<RP, NW>, NW * 32, (int)S::BYTES, !P.do_y, P, seg_hint, num_sms, st, info, NW * 28);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
