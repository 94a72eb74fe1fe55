// 1-D kernel instantiations.  Compiled twice (see csrc/build.py), like sweep_inst.cu:
//   -DCLAW_NS=strict -DCLAW_STRICT=1 -fmad=false   /   -DCLAW_NS=fast -fmad=true
#include "launch.h"
#include "rp1.cuh"
#include "sweep1.cuh"

namespace clawcuda {
namespace CLAW_NS {

template <class RP>
static cudaError_t launch1(const Sweep2Params& P, cudaStream_t st, LaunchInfo* info) {
  constexpr int NT = 128;
  const int ncell = P.mx + 4;
  const int grid = (ncell + (NT - 4) - 1) / (NT - 4);
  sweep1<RP, NT><<<grid, NT, 0, st>>>(P);
  if (info) { info->grid_x = grid; info->grid_y = 1; info->block = NT; info->smem_bytes = 0; info->ctas_per_sm = 0; }
  return cudaGetLastError();
}

cudaError_t launch_sweep1(int rp_id, Sweep2Params& P, cudaStream_t st, LaunchInfo* info) {
  switch (rp_id) {
    case 21: return launch1<RpAdvection1D>(P, st, info);
    case 22: return launch1<RpAcoustics1D>(P, st, info);
    case 23: return launch1<RpBurgers1D>(P, st, info);
    case 24: return launch1<RpShallowRoeEfix1D>(P, st, info);
    case 25: return launch1<RpShallowHlle1D>(P, st, info);
    case 26: return launch1<RpTraffic1D>(P, st, info);
    case 27: return launch1<RpAdvectionColor1D>(P, st, info);
    case 28: return launch1<RpAcousticsVariable1D>(P, st, info);
    case 29: return launch1<RpShallowBathymetryFwave1D>(P, st, info);
    case 30: return launch1<RpEulerEfix1D>(P, st, info);
    case 31: return launch1<RpEulerHlle1D>(P, st, info);
    case 32: return launch1<RpNonlinearElasticityFwave1D>(P, st, info);
    case 33: return launch1<RpCubic1D>(P, st, info);
    case 34: return launch1<RpTrafficVc1D>(P, st, info);
    case 35: return launch1<RpTrafficVcFwave1D>(P, st, info);
    case 36: return launch1<RpTrafficVcTracer1D>(P, st, info);
  }
  return cudaErrorInvalidValue;
}

}  // namespace CLAW_NS
}  // namespace clawcuda
