// Kernel instantiations + launch heuristics.  Compiled twice (see csrc/build.py):
//   -DCLAW_NS=strict -DCLAW_STRICT=1 -fmad=false     op-for-op the reference (bit-exact vs oracle)
//   -DCLAW_NS=fast   -fmad=true                      FMA contraction allowed
#include "sweep2_launch.cuh"
#include "rp2_aux.cuh"

namespace clawcuda {
namespace CLAW_NS {

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
        if ((variant == 2 || P.dense_pref) && split_method<2>(P, 4))   // without the uniform-region fast path (A/B; dense data)
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
        if (variant == 2 || P.dense_pref) return launch_ws<RpEuler4Wave2DT<3>, 1, 8, MC | 1 | 6>(P, seg_hint, num_sms, st, info);
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
        if ((variant == 2 || P.dense_pref) && split_method<3>(P, 4))   // without the uniform-region fast path (A/B; dense data)
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
