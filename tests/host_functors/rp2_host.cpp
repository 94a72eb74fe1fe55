// TEST INFRASTRUCTURE.  The 2-D DEVICE functor source of the solvers without aux (clawpack_b200/csrc/rp2.cuh) compiled by
// the host compiler with -DCLAW_HOST_CHECK -ffp-contract=off (strict mode's unfused IEEE arithmetic): normal() and the
// two transverse() solves per interface, in the rotated frame the sweep kernels use, mapped back to the natural
// component order for comparison with the oracle's rpn2 / rpt2 (tests/test_host_functors.py).  Nothing in the product
// links this file.  euler_4wave_2D is checked as RpEuler4Wave2DT<3> (per-cell quantities shared): the functor both the strict
// kernels and the default fast kernel of bench.py instantiate.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST          // namespace fast: FMA contraction allowed, restructured divisions / square roots
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include <cmath>
#include <cstddef>
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign; using std::sin; using std::cos; using std::fmod;
#include "../../clawpack_b200/csrc/rp2.cuh"

namespace {
using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

template <int M> constexpr int rot(int m) { return (M >= 3 && m == 1) ? 2 : ((M >= 3 && m == 2) ? 1 : m); }   // sweep2.cuh:89

// q(meqn, n+1) natural order; interface k between cells k and k+1
template <class RP>
int run(int dir, const double* cparam, int ncparam, int n, const double* q, double* wave, double* s, double* amdq,
        double* apdq, double* bm_a, double* bp_a, double* bm_b, double* bp_b) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, NC = RP::NCELL > 0 ? RP::NCELL : 1;
  RpParams P{};
  for (int k = 0; k < 4 && k < ncparam; ++k) P.p[k] = cparam[k];
  for (int i = 0; i < n; ++i) {
    double ql[M], qr[M], cl[NC], cr[NC], w[W][M], sp[W], am[M], ap[M], roe[NR], bm[M], bp[M];
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;                   // rotated slot m holds natural component mn
      ql[m] = q[(size_t)i * M + mn];
      qr[m] = q[(size_t)(i + 1) * M + mn];
    }
    if (RP::NCELL > 0) { RP::cell(P, ql, cl); RP::cell(P, qr, cr); }
    RP::normal(P, dir, ql, qr, cl, cr, w, sp, am, ap, roe);
    for (int mw = 0; mw < W; ++mw) {
      s[(size_t)i * W + mw] = sp[mw];
      for (int m = 0; m < M; ++m) wave[((size_t)i * W + mw) * M + (dir ? rot<M>(m) : m)] = w[mw][m];
    }
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      amdq[(size_t)i * M + mn] = am[m];
      apdq[(size_t)i * M + mn] = ap[m];
    }
    RP::transverse(P, dir, roe, am, bm, bp);
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      bm_a[(size_t)i * M + mn] = bm[m]; bp_a[(size_t)i * M + mn] = bp[m];
    }
    RP::transverse(P, dir, roe, ap, bm, bp);
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      bm_b[(size_t)i * M + mn] = bm[m]; bp_b[(size_t)i * M + mn] = bp[m];
    }
  }
  return 0;
}
}  // namespace

extern "C" int hostfun_rp2(int rp_id, int dir, const double* cparam, int ncparam, int n, const double* q, double* wave,
                           double* s, double* amdq, double* apdq, double* bm_a, double* bp_a, double* bm_b,
                           double* bp_b) {
#define CASE(ID, RP) case ID: return run<RP>(dir, cparam, ncparam, n, q, wave, s, amdq, apdq, bm_a, bp_a, bm_b, bp_b);
  switch (rp_id) {
    CASE(1, RpAdvection2D) CASE(2, RpAcoustics2D) CASE(3, RpEuler4Wave2DT<3>) CASE(4, RpShallowRoeEfix2D)
    CASE(9, RpEuler5Wave2D) CASE(11, RpEulerHlle2D) CASE(12, RpShallowHlle2D) CASE(16, RpBurgers2D) CASE(18, RpKpp2D)
  }
#undef CASE
  return 1;
}
