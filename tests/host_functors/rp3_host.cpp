// TEST INFRASTRUCTURE.  The 3-D DEVICE functor source (clawpack_b200/csrc/rp3.cuh) compiled by the host compiler with
// -DCLAW_HOST_CHECK -ffp-contract=off: normal<IXYZ>() for every solver and, for the solvers whose transverse routines read
// no aux (euler_3D, burgers_3D), transverse<IXYZ,ICOOR>() as rpt3 on both fluctuations and as rptt3 on one of the results,
// for comparison with the oracle (tests/test_host_functors.py).  Which aux cells a transverse solve of the aux-reading
// solvers sees is decided by the sweep kernel (flux3_side, sweep3.cuh) and is covered by the -m gpu tests.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST          // namespace fast: FMA contraction allowed, restructured divisions / square roots
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include <cmath>
#include <cstddef>
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign;
#include "../../clawpack_b200/csrc/rp3.cuh"

namespace {
using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

template <class RP, int IXYZ, int ICOOR>
int run(const double* cparam, int ncparam, int n, int maux, const double* q, const double* aux, double* wave, double* s,
        double* amdq, double* apdq, double* bm_a, double* bp_a, double* bm_b, double* bp_b, double* cm_tt, double* cp_tt,
        int with_rpt) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NC = RP::NCTX > 0 ? RP::NCTX : 1;
  if (maux < RP::MAUX) return 2;
  RpParams P{};
  for (int k = 0; k < 4 && k < ncparam; ++k) P.p[k] = cparam[k];
  const double zero[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    double w[W][M], sp[W], am[M], ap[M], ctx[NC], bm[M], bp[M], cm[M], cp[M];
    const double* al = maux > 0 ? aux + (size_t)i * maux : zero;
    const double* ar = maux > 0 ? aux + (size_t)(i + 1) * maux : zero;
    RP::template normal<IXYZ>(P, q + (size_t)i * M, q + (size_t)(i + 1) * M, al, ar, w, sp, am, ap, ctx);
    for (int mw = 0; mw < W; ++mw) {
      s[(size_t)i * W + mw] = sp[mw];
      for (int m = 0; m < M; ++m) wave[((size_t)i * W + mw) * M + m] = w[mw][m];
    }
    for (int m = 0; m < M; ++m) { amdq[(size_t)i * M + m] = am[m]; apdq[(size_t)i * M + m] = ap[m]; }
    if (!with_rpt) continue;
    RP::template transverse<IXYZ, ICOOR>(P, ctx, zero, zero, zero, am, bm, bp);
    for (int m = 0; m < M; ++m) { bm_a[(size_t)i * M + m] = bm[m]; bp_a[(size_t)i * M + m] = bp[m]; }
    RP::template transverse<IXYZ, 5 - ICOOR, true>(P, ctx, zero, zero, zero, bp, cm, cp);      // rptt3 on bpasdq
    for (int m = 0; m < M; ++m) { cm_tt[(size_t)i * M + m] = cm[m]; cp_tt[(size_t)i * M + m] = cp[m]; }
    RP::template transverse<IXYZ, ICOOR>(P, ctx, zero, zero, zero, ap, bm, bp);
    for (int m = 0; m < M; ++m) { bm_b[(size_t)i * M + m] = bm[m]; bp_b[(size_t)i * M + m] = bp[m]; }
  }
  return 0;
}

template <class RP>
int dispatch(int ixyz, int icoor, const double* cparam, int ncparam, int n, int maux, const double* q, const double* aux,
             double* wave, double* s, double* amdq, double* apdq, double* a, double* b, double* c, double* d, double* e,
             double* f, int with_rpt) {
#define GO(I, C) return run<RP, I, C>(cparam, ncparam, n, maux, q, aux, wave, s, amdq, apdq, a, b, c, d, e, f, with_rpt)
  if (icoor == 2) { if (ixyz == 0) GO(0, 2); if (ixyz == 1) GO(1, 2); if (ixyz == 2) GO(2, 2); }
  if (icoor == 3) { if (ixyz == 0) GO(0, 3); if (ixyz == 1) GO(1, 3); if (ixyz == 2) GO(2, 3); }
#undef GO
  return 3;
}
}  // namespace

extern "C" int hostfun_rp3(int rp_id, int ixyz, int icoor, const double* cparam, int ncparam, int n, int maux,
                           const double* q, const double* aux, double* wave, double* s, double* amdq, double* apdq,
                           double* bm_a, double* bp_a, double* bm_b, double* bp_b, double* cm_tt, double* cp_tt,
                           int with_rpt) {
#define CASE(ID, RP) \
  case ID: return dispatch<RP>(ixyz, icoor, cparam, ncparam, n, maux, q, aux, wave, s, amdq, apdq, bm_a, bp_a, bm_b, bp_b, cm_tt, cp_tt, with_rpt);
  switch (rp_id) {
    CASE(5, RpVcAcoustics3D) CASE(6, RpEuler3D) CASE(15, RpBurgers3D) CASE(17, RpVcAdvection3D)
  }
#undef CASE
  return 1;
}
