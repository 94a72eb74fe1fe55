// TEST INFRASTRUCTURE.  The 2-D DEVICE functor source of the aux-reading solvers (clawpack_b200/csrc/rp2_aux.cuh) compiled
// by the host compiler with -DCLAW_HOST_CHECK -ffp-contract=off, for comparison with the oracle's rpn2 / rpt2
// (tests/test_host_functors.py).  See rp2_host.cpp.  Nothing in the product links this file.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST          // namespace fast: FMA contraction allowed, restructured divisions / square roots
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include <cmath>
#include <cstddef>
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign; using std::llround;
#include "../../clawpack_b200/csrc/rp2_aux.cuh"

namespace {
using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

template <int M> constexpr int rot(int m) { return (M >= 3 && m == 1) ? 2 : ((M >= 3 && m == 2) ? 1 : m); }   // sweep2.cuh:89

// q(meqn, n+1), aux1/2/3(maux, n+1) natural order; interface k between cells k and k+1
template <class RP>
int run(int dir, const double* cparam, int ncparam, int n, int maux, const double* q, const double* a1,
        const double* a2, const double* a3, double* wave, double* s, double* amdq, double* apdq, double* bm_a,
        double* bp_a, double* bm_b, double* bp_b, int with_rpt) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE;
  if (maux < RP::MAUX) return 2;
  RpParams P{};
  for (int k = 0; k < 4 && k < ncparam; ++k) P.p[k] = cparam[k];
  for (int i = 0; i < n; ++i) {
    double ql[M], qr[M], w[W][M], sp[W], am[M], ap[M], roe[NR], bm[M], bp[M];
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      ql[m] = q[(size_t)i * M + mn];
      qr[m] = q[(size_t)(i + 1) * M + mn];
    }
    const double *l1 = a1 + (size_t)i * maux, *l2 = a2 + (size_t)i * maux, *l3 = a3 + (size_t)i * maux;
    const double *r1 = a1 + (size_t)(i + 1) * maux, *r2 = a2 + (size_t)(i + 1) * maux, *r3 = a3 + (size_t)(i + 1) * maux;
    RP::normal(P, dir, ql, qr, l2, r2, w, sp, am, ap, roe);
    for (int mw = 0; mw < W; ++mw) {
      s[(size_t)i * W + mw] = sp[mw];
      for (int m = 0; m < M; ++m) wave[((size_t)i * W + mw) * M + (dir ? rot<M>(m) : m)] = w[mw][m];
    }
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      amdq[(size_t)i * M + mn] = am[m];
      apdq[(size_t)i * M + mn] = ap[m];
    }
    if (!with_rpt) continue;
    RP::transverse(P, dir, roe, l1, l2, l3, r1, r2, r3, am, bm, bp);          // imp = 1: enters the left cell
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      bm_a[(size_t)i * M + mn] = bm[m]; bp_a[(size_t)i * M + mn] = bp[m];
    }
    RP::transverse(P, dir, roe, r1, r2, r3, r1, r2, r3, ap, bm, bp);          // imp = 2: enters the right cell
    for (int m = 0; m < M; ++m) {
      const int mn = dir ? rot<M>(m) : m;
      bm_b[(size_t)i * M + mn] = bm[m]; bp_b[(size_t)i * M + mn] = bp[m];
    }
  }
  return 0;
}
}  // namespace

extern "C" int hostfun_rp2_aux(int rp_id, int dir, const double* cparam, int ncparam, int n, int maux, const double* q,
                               const double* a1, const double* a2, const double* a3, double* wave, double* s,
                               double* amdq, double* apdq, double* bm_a, double* bp_a, double* bm_b, double* bp_b,
                               int with_rpt) {
#define CASE(ID, RP) \
  case ID: return run<RP>(dir, cparam, ncparam, n, maux, q, a1, a2, a3, wave, s, amdq, apdq, bm_a, bp_a, bm_b, bp_b, with_rpt);
  switch (rp_id) {
    CASE(7, RpVcAcoustics2D) CASE(8, RpVcAdvection2D) CASE(10, RpPsystem2D) CASE(13, RpAcousticsMapped2D)
    CASE(14, RpShallowBathymetryFwave2D) CASE(19, RpVcElasticity2D) CASE(20, RpVcAdvectionVector2D)
  }
#undef CASE
  return 1;
}
