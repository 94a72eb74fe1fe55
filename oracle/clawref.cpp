// ORACLE -- TEST INFRASTRUCTURE ONLY (see clawref2.hpp header).  C entry points for ctypes.
//
// build:  make -C oracle        (g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC)
#include "clawref2.hpp"
#include "clawref1.hpp"
#ifdef CLAWREF_WITH_3D
#include "clawref3.hpp"
#endif

#include <cstdint>
#include <cstring>

namespace clawref {

// ---------------------------------------------------------------------
// Counting scalar: every arithmetic operator bumps a thread-local tally so that one
// oracle step yields the exact operation mix of the reference algorithm.
// ---------------------------------------------------------------------
struct OpTally { uint64_t add, mul, div, sqrt, cmp, abs; };
static thread_local OpTally g_tally_alg, g_tally_rec;
#define g_tally (*(g_recompute_scope ? &g_tally_rec : &g_tally_alg))

struct Counted {
  double v;
  Counted() : v(0.0) {}
  explicit Counted(double x) : v(x) {}
};
inline Counted operator+(Counted a, Counted b) { ++g_tally.add; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { ++g_tally.add; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { ++g_tally.mul; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { ++g_tally.div; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline bool operator>(Counted a, Counted b) { ++g_tally.cmp; return a.v > b.v; }
inline bool operator<(Counted a, Counted b) { ++g_tally.cmp; return a.v < b.v; }
inline bool operator>=(Counted a, Counted b) { ++g_tally.cmp; return a.v >= b.v; }
inline bool operator<=(Counted a, Counted b) { ++g_tally.cmp; return a.v <= b.v; }
inline bool operator==(Counted a, Counted b) { ++g_tally.cmp; return a.v == b.v; }
inline bool operator!=(Counted a, Counted b) { ++g_tally.cmp; return a.v != b.v; }
inline Counted r_sqrt(Counted x) { ++g_tally.sqrt; return Counted(std::sqrt(x.v)); }
inline Counted r_abs(Counted x) { ++g_tally.abs; return Counted(std::fabs(x.v)); }
inline Counted r_max(Counted a, Counted b) { ++g_tally.cmp; return (b.v > a.v) ? b : a; }
inline Counted r_min(Counted a, Counted b) { ++g_tally.cmp; return (b.v < a.v) ? b : a; }
inline Counted r_sign1(Counted s) { return Counted(std::copysign(1.0, s.v)); }
inline Counted r_exp(Counted x) { ++g_tally.sqrt; return Counted(std::exp(x.v)); }     // tallied with the square roots
inline int r_nint(Counted x) { return (int)std::lround(x.v); }
inline Counted r_sin(Counted x) { ++g_tally.sqrt; return Counted(std::sin(x.v)); }     // library calls: tallied with the square roots
inline Counted r_cos(Counted x) { ++g_tally.sqrt; return Counted(std::cos(x.v)); }
inline Counted r_atan(Counted x) { return Counted(std::atan(x.v)); }
inline Counted r_mod(Counted a, Counted p) { ++g_tally.div; return Counted(std::fmod(a.v, p.v)); }

}  // namespace clawref

using namespace clawref;

extern "C" {

// qnew,cfl = step2(maxm,num_ghost,mx,my,qold,qnew,aux,dx,dy,dt,method,mthlim,...,use_fwave,rpn2,rpt2)
// (f2py signature of pyclaw/src/pyclaw/classic/step2.f90:2-5).  The (rpn2,rpt2) C-pointer pair is
// replaced by rp_id + cparam (the COMMON /cparam/ contents in declaration order).
int clawref_step2(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my,
                  double* qold, double* qnew, double* aux, double dx, double dy, double dt,
                  const int* method, const int* mthlim, double* cfl, int use_fwave,
                  int rp_id, const double* cparam, int nthreads) {
  return step2<double>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, qold, qnew, aux, dx, dy, dt,
                       method, mthlim, cfl, use_fwave != 0, rp_id, cparam, nthreads);
}

int clawref_step2ds(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my,
                    double* qold, double* qnew, double* aux, double dx, double dy, double dt,
                    const int* method, const int* mthlim, double* cfl, int ids, int use_fwave,
                    int rp_id, const double* cparam, int nthreads) {
  return step2ds<double>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, qold, qnew, aux, dx, dy, dt,
                         method, mthlim, cfl, ids, use_fwave != 0, rp_id, cparam, nthreads);
}

#ifdef CLAWREF_WITH_3D
int clawref_step3(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my, int mz,
                  double* qold, double* qnew, double* aux, double dx, double dy, double dz, double dt,
                  const int* method, const int* mthlim, double* cfl, int use_fwave,
                  int rp_id, const double* cparam, int nthreads) {
  return step3<double>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, mz, qold, qnew, aux, dx, dy, dz, dt,
                       method, mthlim, cfl, use_fwave != 0, rp_id, cparam, nthreads);
}

int clawref_step3ds(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my, int mz,
                    double* qold, double* qnew, double* aux, double dx, double dy, double dz, double dt,
                    const int* method, const int* mthlim, double* cfl, int idir, int use_fwave,
                    int rp_id, const double* cparam, int nthreads) {
  return step3ds<double>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, mz, qold, qnew, aux, dx, dy, dz, dt,
                         method, mthlim, cfl, idir, use_fwave != 0, rp_id, cparam, nthreads);
}

#endif

// Operation counts of ONE hyperbolic step on the given data (serial, counting scalar).
// split != 0: step2ds(ids=1) then step2ds(ids=2) aliased, as classic/solver.py:527-535.
// counts[12] = add(+,-), mul, div, sqrt, compare/min/max, abs of the algorithm, then the same six for the
// work the reference merely repeats (Roe averages rebuilt inside rpt2)
int clawref_count_ops2(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my,
                       const double* qbc, const double* aux, double dx, double dy, double dt,
                       const int* method, const int* mthlim, int split, int use_fwave,
                       int rp_id, const double* cparam, uint64_t* counts) {
  const size_t ncell = (size_t)(mx + 2 * num_ghost) * (my + 2 * num_ghost);
  std::vector<Counted> qold(ncell * num_eqn), qnew(ncell * num_eqn), ax(ncell * (num_aux > 0 ? num_aux : 1));
  for (size_t k = 0; k < ncell * num_eqn; ++k) { qold[k] = Counted(qbc[k]); qnew[k] = Counted(qbc[k]); }
  if (num_aux > 0) for (size_t k = 0; k < ncell * num_aux; ++k) ax[k] = Counted(aux[k]);
  std::memset(&g_tally_alg, 0, sizeof(OpTally)); std::memset(&g_tally_rec, 0, sizeof(OpTally));
  Counted cfl; int rc;
  if (split) {
    rc = step2ds<Counted>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, qold.data(), qnew.data(), ax.data(),
                          Counted(dx), Counted(dy), Counted(dt), method, mthlim, &cfl, 1, use_fwave != 0, rp_id, cparam, 1);
    if (rc == 0)
      rc = step2ds<Counted>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, qnew.data(), qnew.data(), ax.data(),
                            Counted(dx), Counted(dy), Counted(dt), method, mthlim, &cfl, 2, use_fwave != 0, rp_id, cparam, 1);
  } else {
    rc = step2<Counted>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, qold.data(), qnew.data(), ax.data(),
                        Counted(dx), Counted(dy), Counted(dt), method, mthlim, &cfl, use_fwave != 0, rp_id, cparam, 1);
  }
  counts[0] = g_tally_alg.add; counts[1] = g_tally_alg.mul; counts[2] = g_tally_alg.div;
  counts[3] = g_tally_alg.sqrt; counts[4] = g_tally_alg.cmp; counts[5] = g_tally_alg.abs;
  counts[6] = g_tally_rec.add; counts[7] = g_tally_rec.mul; counts[8] = g_tally_rec.div;
  counts[9] = g_tally_rec.sqrt; counts[10] = g_tally_rec.cmp; counts[11] = g_tally_rec.abs;
  return rc;
}

#ifdef CLAWREF_WITH_3D
int clawref_count_ops3(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my, int mz,
                       const double* qbc, const double* aux, double dx, double dy, double dz, double dt,
                       const int* method, const int* mthlim, int split, int use_fwave,
                       int rp_id, const double* cparam, uint64_t* counts) {
  const size_t ncell = (size_t)(mx + 2 * num_ghost) * (my + 2 * num_ghost) * (mz + 2 * num_ghost);
  std::vector<Counted> qold(ncell * num_eqn), qnew(ncell * num_eqn), ax(ncell * (num_aux > 0 ? num_aux : 1));
  for (size_t k = 0; k < ncell * num_eqn; ++k) { qold[k] = Counted(qbc[k]); qnew[k] = Counted(qbc[k]); }
  if (num_aux > 0) for (size_t k = 0; k < ncell * num_aux; ++k) ax[k] = Counted(aux[k]);
  std::memset(&g_tally_alg, 0, sizeof(OpTally)); std::memset(&g_tally_rec, 0, sizeof(OpTally));
  Counted cfl; int rc = 0;
  if (split) {
    for (int idir = 1; idir <= 3 && rc == 0; ++idir)
      rc = step3ds<Counted>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, mz,
                            idir == 1 ? qold.data() : qnew.data(), qnew.data(), ax.data(),
                            Counted(dx), Counted(dy), Counted(dz), Counted(dt), method, mthlim, &cfl, idir,
                            use_fwave != 0, rp_id, cparam, 1);
  } else {
    rc = step3<Counted>(maxm, num_eqn, num_waves, num_aux, num_ghost, mx, my, mz, qold.data(), qnew.data(), ax.data(),
                        Counted(dx), Counted(dy), Counted(dz), Counted(dt), method, mthlim, &cfl, use_fwave != 0,
                        rp_id, cparam, 1);
  }
  counts[0] = g_tally_alg.add; counts[1] = g_tally_alg.mul; counts[2] = g_tally_alg.div;
  counts[3] = g_tally_alg.sqrt; counts[4] = g_tally_alg.cmp; counts[5] = g_tally_alg.abs;
  counts[6] = g_tally_rec.add; counts[7] = g_tally_rec.mul; counts[8] = g_tally_rec.div;
  counts[9] = g_tally_rec.sqrt; counts[10] = g_tally_rec.cmp; counts[11] = g_tally_rec.abs;
  return rc;
}

#endif

// q,cfl = step1(num_ghost,mx,q,aux,dx,dt,method,mthlim,use_fwave,rp1)  (f2py signature of
// pyclaw/src/pyclaw/classic/step1.f90:2-3 as called at classic/solver.py:302); q is updated in place.
int clawref_step1(int num_eqn, int num_waves, int num_ghost, int num_aux, int mx, double* q, double* aux, double dx,
                  double dt, const int* method, const int* mthlim, double* cfl, int use_fwave, int rp_id,
                  const double* cparam) {
  return step1<double>(num_eqn, num_waves, num_ghost, num_aux, mx, q, aux, dx, dt, method, mthlim, cfl,
                       use_fwave != 0, rp_id, cparam);
}

// One rp1 call on a pencil of n = mx + 2*num_ghost cells: q(num_eqn, n), aux(num_aux, n); interface slot k (array
// index, k = 1 .. n-1) lies between cells k-1 and k: wave(num_eqn, num_waves, n), s(num_waves, n), amdq/apdq(num_eqn, n).
// Used by tests/test_host_functors.py to compare the device functor source (host-compiled) with the oracle.
int clawref_rp1(int num_eqn, int num_waves, int num_ghost, int num_aux, int mx, double* q, double* aux, double* wave,
                double* s, double* amdq, double* apdq, int rp_id, const double* cparam) {
  Rp1<double> rp;
  if (!select_rp1<double>(rp_id, rp)) return 1;
  const int lo = 1 - num_ghost, na = num_aux > 0 ? num_aux : 1;
  V2<double> Q{q, num_eqn, lo}, A{aux, na, lo}, S{s, num_waves, lo}, AM{amdq, num_eqn, lo}, AP{apdq, num_eqn, lo};
  V3<double> W{wave, num_eqn, num_waves, lo};
  rp(mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, A, A, W, S, AM, AP, cparam);
  return 0;
}

// One rpn2 call followed by rpt2 on its two fluctuations (imp = 1 on amdq, imp = 2 on apdq) on a pencil of
// n = mx + 2*num_ghost cells of a solver WITHOUT aux: q(num_eqn, n) in the natural component order; outputs in the
// Fortran layouts with interface slot k between cells k-1 and k.  rpt2 is skipped for solvers with rpt2_dummy.
// Used by tests/test_host_functors.py.
int clawref_rp2(int ixy, int num_eqn, int num_waves, int num_ghost, int mx, double* q, double* wave, double* s,
                double* amdq, double* apdq, double* bm_a, double* bp_a, double* bm_b, double* bp_b, int rp_id,
                const double* cparam) {
  Rpn2<double> rpn; Rpt2<double> rpt;
  if (!select_rp2<double>(rp_id, rpn, rpt)) return 1;
  const int lo = 1 - num_ghost;
  std::vector<double> dummy((size_t)(mx + 2 * num_ghost), 0.0);
  V2<double> Q{q, num_eqn, lo}, A{dummy.data(), 1, lo}, S{s, num_waves, lo}, AM{amdq, num_eqn, lo}, AP{apdq, num_eqn, lo};
  V3<double> W{wave, num_eqn, num_waves, lo};
  rpn(ixy, mx, num_eqn, num_waves, 0, num_ghost, mx, Q, Q, A, A, W, S, AM, AP, cparam);
  if (rpt == rpt2_dummy<double>) return 0;
  rpt(ixy, 1, mx, num_eqn, num_waves, 0, num_ghost, mx, Q, Q, A, A, A, AM, V2<double>{bm_a, num_eqn, lo},
      V2<double>{bp_a, num_eqn, lo}, cparam);
  rpt(ixy, 2, mx, num_eqn, num_waves, 0, num_ghost, mx, Q, Q, A, A, A, AP, V2<double>{bm_b, num_eqn, lo},
      V2<double>{bp_b, num_eqn, lo}, cparam);
  return 0;
}

// The same for solvers that read aux: aux1 / aux2 / aux3 (num_aux, n) are the three pencils rpt2 sees (rows j-1, j, j+1
// of an x-sweep); rpn2 gets aux2.  with_rpt = 0 skips the transverse solves.
int clawref_rp2_aux(int ixy, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, double* q, double* aux1,
                    double* aux2, double* aux3, double* wave, double* s, double* amdq, double* apdq, double* bm_a,
                    double* bp_a, double* bm_b, double* bp_b, int rp_id, const double* cparam, int with_rpt) {
  Rpn2<double> rpn; Rpt2<double> rpt;
  if (!select_rp2<double>(rp_id, rpn, rpt)) return 1;
  const int lo = 1 - num_ghost;
  V2<double> Q{q, num_eqn, lo}, A1{aux1, num_aux, lo}, A2{aux2, num_aux, lo}, A3{aux3, num_aux, lo}, S{s, num_waves, lo},
      AM{amdq, num_eqn, lo}, AP{apdq, num_eqn, lo};
  V3<double> W{wave, num_eqn, num_waves, lo};
  rpn(ixy, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, A2, A2, W, S, AM, AP, cparam);
  if (!with_rpt || rpt == rpt2_dummy<double>) return 0;
  rpt(ixy, 1, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, A1, A2, A3, AM, V2<double>{bm_a, num_eqn, lo},
      V2<double>{bp_a, num_eqn, lo}, cparam);
  rpt(ixy, 2, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, A1, A2, A3, AP, V2<double>{bm_b, num_eqn, lo},
      V2<double>{bp_b, num_eqn, lo}, cparam);
  return 0;
}

#ifdef CLAWREF_WITH_3D
// 3-D: one rpn3 call on a pencil (aux(num_aux, n) is the pencil's own aux row) and, for solvers whose transverse
// routines read no aux (with_rpt != 0: euler_3D, burgers_3D), rpt3 with `icoor` on amdq (imp = 1) and apdq (imp = 2)
// and rptt3 (imp = 1, impt = 1) on the first of those results.  Used by tests/test_host_functors.py.
int clawref_rp3(int ixyz, int icoor, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, double* q,
                double* aux, double* wave, double* s, double* amdq, double* apdq, double* bm_a, double* bp_a,
                double* bm_b, double* bp_b, double* cm_tt, double* cp_tt, int rp_id, const double* cparam,
                int with_rpt) {
  Rpn3<double> rpn; Rpt3<double> rpt; Rptt3<double> rptt;
  if (!select_rp3<double>(rp_id, rpn, rpt, rptt)) return 1;
  const int lo = 1 - num_ghost, n = mx + 2 * num_ghost, na = num_aux > 0 ? num_aux : 1;
  std::vector<double> dummy((size_t)3 * n * na, 0.0);
  V2<double> Q{q, num_eqn, lo}, A{num_aux > 0 ? aux : dummy.data(), na, lo}, S{s, num_waves, lo}, AM{amdq, num_eqn, lo},
      AP{apdq, num_eqn, lo};
  V3<double> W{wave, num_eqn, num_waves, lo};
  rpn(ixyz, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, A, A, W, S, AM, AP, cparam);
  if (!with_rpt) return 0;
  A3<double> D{dummy.data(), na, lo, n};
  V2<double> BMA{bm_a, num_eqn, lo}, BPA{bp_a, num_eqn, lo};
  rpt(ixyz, icoor, 1, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, D, D, D, AM, BMA, BPA, cparam);
  rpt(ixyz, icoor, 2, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, D, D, D, AP, V2<double>{bm_b, num_eqn, lo},
      V2<double>{bp_b, num_eqn, lo}, cparam);
  // double-transverse solve in the OTHER transverse direction on the up-going part of the first split
  rptt(ixyz, 5 - icoor, 1, 1, mx, num_eqn, num_waves, num_aux, num_ghost, mx, Q, Q, D, D, D, BPA,
       V2<double>{cm_tt, num_eqn, lo}, V2<double>{cp_tt, num_eqn, lo}, cparam);
  return 0;
}
#endif

int clawref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
