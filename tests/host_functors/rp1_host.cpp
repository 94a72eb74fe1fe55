// TEST INFRASTRUCTURE.  The 1-D DEVICE functor source (clawpack_b200/csrc/rp1.cuh) compiled by the host compiler with
// -DCLAW_HOST_CHECK -ffp-contract=off, i.e. with the strict mode's unfused IEEE arithmetic, so that the arithmetic of
// every functor can be compared bit for bit with the oracle on a machine without a GPU (tests/test_host_functors.py).
// Nothing in the product links this file.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST          // namespace fast: FMA contraction allowed, restructured divisions / square roots
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include <cmath>
#include <type_traits>
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign;
#include "../../clawpack_b200/csrc/rp1.cuh"

namespace {
using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

template <class RP, class = void> struct AuxN { static constexpr int N = 0; };
template <class RP> struct AuxN<RP, std::void_t<decltype(RP::MAUX)>> { static constexpr int N = RP::MAUX; };

// n interfaces: interface k lies between cells k and k+1 of q(meqn, n+1) / aux(maux, n+1) (PyClaw's (m, i) order)
template <class RP>
int run(const double* cparam, int ncparam, int n, int maux, const double* q, const double* aux, double* wave, double* s,
        double* amdq, double* apdq) {
  constexpr int M = RP::MEQN, W = RP::MWAVES;
  RpParams P{};
  for (int k = 0; k < 4 && k < ncparam; ++k) P.p[k] = cparam[k];
  for (int i = 0; i < n; ++i) {
    double w[W][M], sp[W], am[M], ap[M];
    const double* ql = q + (size_t)i * M;
    const double* qr = q + (size_t)(i + 1) * M;
    if constexpr (AuxN<RP>::N > 0) {
      if (maux < AuxN<RP>::N || !aux) return 2;
      RP::normal(P, ql, qr, aux + (size_t)i * maux, aux + (size_t)(i + 1) * maux, w, sp, am, ap);
    } else {
      RP::normal(P, ql, qr, w, sp, am, ap);
    }
    for (int mw = 0; mw < W; ++mw) {
      s[(size_t)i * W + mw] = sp[mw];
      for (int m = 0; m < M; ++m) wave[((size_t)i * W + mw) * M + m] = w[mw][m];
    }
    for (int m = 0; m < M; ++m) { amdq[(size_t)i * M + m] = am[m]; apdq[(size_t)i * M + m] = ap[m]; }
  }
  return 0;
}
}  // namespace

extern "C" int hostfun_rp1(int rp_id, const double* cparam, int ncparam, int n, int maux, const double* q,
                           const double* aux, double* wave, double* s, double* amdq, double* apdq) {
#define CASE(ID, RP) case ID: return run<RP>(cparam, ncparam, n, maux, q, aux, wave, s, amdq, apdq);
  switch (rp_id) {
    CASE(21, RpAdvection1D) CASE(22, RpAcoustics1D) CASE(23, RpBurgers1D) CASE(24, RpShallowRoeEfix1D)
    CASE(25, RpShallowHlle1D) CASE(26, RpTraffic1D) CASE(27, RpAdvectionColor1D) CASE(28, RpAcousticsVariable1D)
    CASE(29, RpShallowBathymetryFwave1D) CASE(30, RpEulerEfix1D) CASE(31, RpEulerHlle1D)
    CASE(32, RpNonlinearElasticityFwave1D) CASE(33, RpCubic1D) CASE(34, RpTrafficVc1D)
    CASE(35, RpTrafficVcFwave1D) CASE(36, RpTrafficVcTracer1D)
  }
#undef CASE
  return 1;
}
