// Kernel instantiations + launch heuristics.  Compiled twice (see csrc/build.py):
//   -DCLAW_NS=strict -DCLAW_STRICT=1 -fmad=false     op-for-op the reference (bit-exact vs oracle)
//   -DCLAW_NS=fast   -fmad=true                      FMA contraction allowed
#include <cstdlib>
#include "launch.h"
#include "sweep2_ws.cuh"
#include "sweep2_aux.cuh"
#include "rp2_aux.cuh"

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

// order 2, transverse mode 2, every wave family limited with limiter id `lim`
template <int W> static bool headline_method(const Sweep2Params& P, int lim) {
  if (P.order == 1 || P.mtrans != 2 || !P.limit) return false;
  for (int mw = 0; mw < W; ++mw) if (P.mthlim[mw] != lim) return false;
  return true;
}

// order 2, every wave family limited with limiter id `lim` (dimension-split step: no transverse mode involved)
template <int W> static bool split_method(const Sweep2Params& P, int lim) {
  if (P.order == 1 || !P.limit) return false;
  for (int mw = 0; mw < W; ++mw) if (P.mthlim[mw] != lim) return false;
  return true;
}

cudaError_t launch_sweep2(int rp_id, int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st,
                          LaunchInfo* info) {
  switch (rp_id) {
    case 1: return launch_one<RpAdvection2D, 128>(kind, P, seg_hint, num_sms, st, info);
    case 2:
      if (kind == 1) {
        static const int variant = getenv("CLAW_VARIANT") ? atoi(getenv("CLAW_VARIANT")) : 0;   // 1: generic kernels only (A/B timing)
        if (variant == 2 && split_method<2>(P, 4))             // without the uniform-region fast path (A/B)
          return launch_split_ws<RpAcoustics2D, 1, 16, 1, 1, 4>(P, seg_hint, num_sms, st, info);
        if (variant != 1 && split_method<2>(P, 4) && P.do_x && P.do_y)   // BASELINE config 1: MC limiter as a template argument
          return launch_split_ws<RpAcoustics2D, 1, 16, 1, 1, 4, 4>(P, seg_hint, num_sms, st, info);
        if (variant != 1 && split_method<2>(P, 4))
          return launch_split_ws<RpAcoustics2D, 1, 16, 1, 1, 4>(P, seg_hint, num_sms, st, info);
        return launch_split_ws<RpAcoustics2D, 1, 16>(P, seg_hint, num_sms, st, info);
      }
      return launch_ws<RpAcoustics2D, 1, 16, 6>(P, seg_hint, num_sms, st, info);          // generic methods + cp.async ring
    case 3: {
      static const int variant = getenv("CLAW_VARIANT") ? atoi(getenv("CLAW_VARIANT")) : 0;   // 1: generic kernels only (A/B timing)
      if (kind == 0 && headline_method<4>(P, 4)) {
        constexpr int MC = 4 << 4;
        // fastest measured: spill-free 8 warps/SM, per-cell Roe data, cp.async row ring (profiles/r01_h, r01_i).
        // Measured and dropped: 12/16 warps with spills, L1/L2 prefetch hints, branch-free limiter / entropy fix, a warp
        // pair per strip (5.21 ms, profiles/r01_j), the march loop unrolled x2 (5.15 ms, profiles/r02_a_variants.jsonl).
        // uniform-region fast path (fast mode only; see sweep2_ws.cuh) with a 2 / 4 / 8-stage row ring
        // default: with the uniform-region fast path (a no-op in strict mode) and an 8-stage row ring; measured at
        // 8192^2 on the shock-bubble data (profiles/r02_g_variants.jsonl): 4.62 ms without it, 1.75 / 1.49 / 1.44 ms
        // with a 2 / 4 / 8-stage ring.  CLAW_VARIANT=2: the same kernel without the fast path (dense-data timing, A/B).
        if (variant == 2) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6>(P, seg_hint, num_sms, st, info);
        if (variant == 71) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6, 1, 4>(P, seg_hint, num_sms, st, info);
        // the row ring fed by TMA bulk copies (cp.async.bulk + mbarrier, one 256-byte copy per component and row issued
        // by lane 0) instead of 32 lanes x 8-byte cp.async: A/B
        if (variant == 91) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6, 1, 8, 1>(P, seg_hint, num_sms, st, info);
        if (variant == 92) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6, 1, 16, 1>(P, seg_hint, num_sms, st, info);
        if (variant == 73) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6, 1, 16>(P, seg_hint, num_sms, st, info);
        if (variant != 1) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6, 1, 8>(P, seg_hint, num_sms, st, info);
      }
      if (kind == 0) return launch_ws<RpEuler4Wave2D, 1, 8, 6>(P, seg_hint, num_sms, st, info);
      if (variant != 1 && split_method<4>(P, 4))              // MC limiter as a template argument
        return launch_split_ws<RpEuler4Wave2D, 1, 8, 1, 1, 4>(P, seg_hint, num_sms, st, info);
      return launch_split_ws<RpEuler4Wave2D, 1, 8>(P, seg_hint, num_sms, st, info);
    }
    case 4: {
      static const int variant = getenv("CLAW_VARIANT") ? atoi(getenv("CLAW_VARIANT")) : 0;   // 1: generic kernels only (A/B timing)
      if (kind == 1) {
        // Limiter known at compile time (MC on all waves, order 2: BASELINE config 3).  Measured on B200 at 16384^2
        // (profiles/r02_a_c3_variants.jsonl): generic 9.48 ms, specialised 8.97 ms; the march loop unrolled x2 was
        // slower in every combination (11.4 / 11.3 / 9.65 ms) and is gone.
        if (variant == 2 && split_method<3>(P, 4))             // without the uniform-region fast path (A/B)
          return launch_split_ws<RpShallowRoeEfix2D, 1, 16, 1, 1, 4>(P, seg_hint, num_sms, st, info);
        // Ring depth: about 50 rows in flight per SM is the optimum of the uniform-region path (16384^2, uniform data,
        // profiles/r02_s: 16 warps x 8 stages 3.01 ms, 16 x 4 2.19 ms, 16 x 2 2.72 ms, 12 x 4 2.37 ms, 8 x 8 2.14 ms,
        // 8 x 4 2.37 ms); 16 warps per SM keep the FP64-bound dense path fastest (12.2 ms against 14.1 at 8, 13.6 at 12).
        if (variant == 82 && split_method<3>(P, 4) && P.do_x && P.do_y)  // A/B: 8 warps per SM, 8-stage ring
          return launch_split_ws<RpShallowRoeEfix2D, 1, 8, 1, 1, 4, 8>(P, seg_hint, num_sms, st, info);
        if (variant != 1 && split_method<3>(P, 4) && P.do_x && P.do_y)   // fused step: with the uniform-region fast path
          return launch_split_ws<RpShallowRoeEfix2D, 1, 16, 1, 1, 4, 4>(P, seg_hint, num_sms, st, info);
        if (variant != 1 && split_method<3>(P, 4))
          return launch_split_ws<RpShallowRoeEfix2D, 1, 16, 1, 1, 4>(P, seg_hint, num_sms, st, info);
        return launch_split_ws<RpShallowRoeEfix2D, 1, 16>(P, seg_hint, num_sms, st, info);
      }
      return launch_ws<RpShallowRoeEfix2D, 1, 12, 6>(P, seg_hint, num_sms, st, info);
    }
    case 9:
      if (kind == 1) return launch_split_ws<RpEuler5Wave2D, 1, 8>(P, seg_hint, num_sms, st, info);
      return launch_ws<RpEuler5Wave2D, 1, 8, 6>(P, seg_hint, num_sms, st, info);
    case 7: return launch_aux<RpVcAcoustics2D, 2>(kind, P, seg_hint, num_sms, st, info);
    case 8: return launch_aux<RpVcAdvection2D, 2>(kind, P, seg_hint, num_sms, st, info);
    case 10: return launch_aux<RpPsystem2D, 2>(kind, P, seg_hint, num_sms, st, info);
    case 16:
      if (kind == 1) return launch_split_ws<RpBurgers2D, 1, 16>(P, seg_hint, num_sms, st, info);
      return launch_ws<RpBurgers2D, 1, 16, 6>(P, seg_hint, num_sms, st, info);
    case 18:                                                 // kpp_2D: scalar, like burgers_2D
      if (kind == 1) return launch_split_ws<RpKpp2D, 1, 16>(P, seg_hint, num_sms, st, info);
      return launch_ws<RpKpp2D, 1, 16, 6>(P, seg_hint, num_sms, st, info);
    case 20: return launch_aux<RpVcAdvectionVector2D, 2>(kind, P, seg_hint, num_sms, st, info);
    case 19: return launch_aux<RpVcElasticity2D, 2>(kind, P, seg_hint, num_sms, st, info);
    case 13: return launch_aux<RpAcousticsMapped2D, 1>(kind, P, seg_hint, num_sms, st, info);
    case 14: return launch_aux<RpShallowBathymetryFwave2D, 2>(kind, P, seg_hint, num_sms, st, info);
    case 11:
      if (kind == 1) return launch_split_ws<RpEulerHlle2D, 1, 12>(P, seg_hint, num_sms, st, info);
      return launch_ws<RpEulerHlle2D, 1, 12, 6>(P, seg_hint, num_sms, st, info);
    case 12:
      if (kind == 1) return launch_split_ws<RpShallowHlle2D, 1, 12>(P, seg_hint, num_sms, st, info);
      return launch_ws<RpShallowHlle2D, 1, 12, 6>(P, seg_hint, num_sms, st, info);
  }
  return cudaErrorInvalidValue;
}

}  // namespace CLAW_NS
}  // namespace clawcuda
