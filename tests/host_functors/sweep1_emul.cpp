// TEST INFRASTRUCTURE.  The 1-D sweep KERNEL source (clawpack_b200/csrc/sweep1.cuh: one thread per cell, shared-memory
// exchange of neighbouring interfaces, block-wide CFL reduction) executed on the CPU by the fiber emulation of
// cuda_emul.h, behind an entry point with the argument list of claw_step1_host (csrc/clawcuda.cu) -- so that the kernel
// plumbing around the functors (aux columns, f-wave branch, tile edges) can be compared with the oracle's step1 without a
// GPU (tests/test_kernel_emulation.py).  Strict arithmetic (-ffp-contract=off), or namespace fast with -DCLAW_HOST_FAST.  Nothing in the product links this file.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST          // namespace fast: the kernels' fast-mode code paths (folded limiter, FMA contraction)
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include "cuda_emul.h"
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign; using std::fmax;
#include "../../clawpack_b200/csrc/rp1.cuh"
#include "../../clawpack_b200/csrc/sweep1.cuh"

using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

template <class RP> static void run1(const Sweep2Params& P) {
  constexpr int NT = 128;                                  // as launch1() in csrc/sweep1_inst.cu
  const int ncell = P.mx + 4;
  emul::Dim3 grid, block;
  grid.x = (ncell + (NT - 4) - 1) / (NT - 4);
  block.x = NT;
  emul::launch(grid, block, [&]() { sweep1<RP, NT>(P); });
}

extern "C" int emul_step1_host(int meqn, int mwaves, int maux, int num_ghost, int mx, double* q, const double* aux,
                               double dx, double dt, const int* method7, const int* mthlim, int use_fwave, int rp_id,
                               const double* cparam, int ncparam, double* cfl) {
  if (num_ghost != 2) return 3;
  const size_t ncell = (size_t)mx + 4;
  std::vector<double> qout(q, q + ncell * meqn);           // the device path copies q into the output buffer first
  unsigned long long cfl_bits = 0;
  Sweep2Params P{};                                        // filled exactly as claw_step1_host does
  P.qin = q; P.qout = qout.data(); P.aux = aux;
  P.cfl_bits = &cfl_bits;
  P.mx = mx; P.my = 1; P.pitch = meqn; P.plane = 1;
  P.dtdx = dt; P.dtdy = dx;
  P.order = method7[1]; P.mtrans = 0; P.limit = 0;
  for (int mw = 0; mw < mwaves; ++mw) { P.mthlim[mw] = mthlim[mw]; if (mthlim[mw] > 0) P.limit = 1; }
  P.maux = maux; P.capa = method7[5]; P.fwave = use_fwave;
  for (int k = 0; k < 4; ++k) P.rp.p[k] = (cparam && k < ncparam) ? cparam[k] : 0.0;
  switch (rp_id) {                                         // as launch_sweep1() in csrc/sweep1_inst.cu
    case 21: run1<RpAdvection1D>(P); break;
    case 22: run1<RpAcoustics1D>(P); break;
    case 23: run1<RpBurgers1D>(P); break;
    case 24: run1<RpShallowRoeEfix1D>(P); break;
    case 25: run1<RpShallowHlle1D>(P); break;
    case 26: run1<RpTraffic1D>(P); break;
    case 27: run1<RpAdvectionColor1D>(P); break;
    case 28: run1<RpAcousticsVariable1D>(P); break;
    case 29: run1<RpShallowBathymetryFwave1D>(P); break;
    case 30: run1<RpEulerEfix1D>(P); break;
    case 31: run1<RpEulerHlle1D>(P); break;
    case 32: run1<RpNonlinearElasticityFwave1D>(P); break;
    case 33: run1<RpCubic1D>(P); break;
    case 34: run1<RpTrafficVc1D>(P); break;
    case 35: run1<RpTrafficVcFwave1D>(P); break;
    case 36: run1<RpTrafficVcTracer1D>(P); break;
    default: return 5;
  }
  std::memcpy(q, qout.data(), sizeof(double) * ncell * meqn);
  std::memcpy(cfl, &cfl_bits, sizeof(double));
  return 0;
}
