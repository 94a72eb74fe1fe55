// 3-D kernel instantiations + launch heuristics.  Compiled twice (see csrc/build.py), like
// sweep_inst.cu:  -DCLAW_NS=strict -fmad=false  /  -DCLAW_NS=fast -fmad=true.
#include "sweep3_launch.cuh"

namespace clawcuda {
namespace CLAW_NS {

// unsplit, order 2, method(3) = 22, every wave family limited with limiter id `lim`
template <int W> static bool headline_method3(const Sweep3Params& P, int lim) {
  if (P.order == 1 || P.m3 != 2 || P.m4 != 2 || !P.limit) return false;
  for (int mw = 0; mw < W; ++mw) if (P.mthlim[mw] != lim) return false;
  return true;
}

// dir: 0 = x-sweep, 1 = y-sweep, 2 = z-sweep
cudaError_t launch_sweep3(int rp_id, int dir, Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info) {
  static const int variant = getenv("CLAW_VARIANT3") ? atoi(getenv("CLAW_VARIANT3")) : 0;   // 1: generic kernels only (A/B timing, tests)
  switch (rp_id) {
    case 5:
      // Method set as template arguments when it is the BASELINE config-4 set (unsplit, order 2, method(3) = 22, MC on
      // every wave), x-sweep as two 32x6 CTAs per SM (168 registers, no spills).  Measured on B200 at 512^3
      // (profiles/r02_b_c4_variants.jsonl): 49.6 ms generic -> 35.4 ms specialised -> 32.4 ms with the 32x6 x-sweep.
      if (variant != 1 && headline_method3<2>(P, 4) && P.capa == 0) {       // (the specialised kernels have no capacity branch)
        // (two copies of the gadd / hadd exchange area alternating by step parity, so that one of the two barriers per
        // march step goes away, were measured at 512^3: marches 12.3 against 11.8 ms -- the 213 KB of shared memory leave
        // the 3x3 aux neighbourhood fetches hardly any L1 --, x-sweep within 1 %: dropped, profiles/r02_zd, r02_ze)
        if (dir == 0) return launch_x<RpVcAcoustics3D, 6, 2, 41>(P, num_sms, st, info);
        return launch_dir<RpVcAcoustics3D, 8, 1, 41>(dir, P, num_sms, st, info);
      }
      return launch_dir<RpVcAcoustics3D, 8, 1>(dir, P, num_sms, st, info);
    case 6: return launch_dir<RpEuler3D, 8, 1>(dir, P, num_sms, st, info);
    case 15: return launch_dir<RpBurgers3D, 8, 1>(dir, P, num_sms, st, info);
    case 17: return launch_dir<RpVcAdvection3D, 8, 1>(dir, P, num_sms, st, info);
  }
  return cudaErrorInvalidValue;
}

}  // namespace CLAW_NS
}  // namespace clawcuda
