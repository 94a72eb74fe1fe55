// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the reference's 2-D classic wave-propagation step
// (PyClaw f2py module `classic2` + the in-scope `clawpack.riemann` solvers).
// Each routine follows the reference Fortran loop for loop and cites the file:line
// it restates.  Compile with  g++ -O2 -ffp-contract=off  (the reference build is
// gfortran on baseline x86-64: no FMA contraction, IEEE div/sqrt).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
//
// Deviations from the reference (documented, deliberate):
//  * the `maxm2 = 1800` pencil-length cap of rpn2_euler_4wave.f90:52,77-80 /
//    rpt2_euler.f90:26,34-37 / rpn2_shallow_roe_with_efix.f90:52 is dropped
//    (scratch arrays are sized from maxm), so BASELINE grids can be run;
//  * Fortran `stop` paths become error returns.
//
// The scalar type is a template parameter so that a counting scalar (clawref.cpp,
// `Counted`) yields exact add/mul/div/sqrt counts per cell.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace clawref {

// Counting hook: when the scalar is the counting type, operations inside a "recompute" scope are
// work the reference repeats verbatim (rpt2 rebuilding the Roe averages rpn2 already produced,
// rpt2_euler.f90:53-70) and are tallied separately from the algorithmic count.
inline thread_local int g_recompute_scope = 0;
struct RecomputeScope { RecomputeScope() { ++g_recompute_scope; } ~RecomputeScope() { --g_recompute_scope; } };

// ---- scalar helpers (overloaded for the counting scalar in clawref.cpp) ----
inline double r_sqrt(double x) { return std::sqrt(x); }
inline double r_abs(double x) { return std::fabs(x); }
// Fortran dmax1/dmin1 (gfortran MAX/MIN on reals: plain compare-and-select)
// gfortran lowers MAX(a,b) to "mvar = a; if (b > mvar) mvar = b" (NaN handling aside).
inline double r_max(double a, double b) { return (b > a) ? b : a; }
inline double r_min(double a, double b) { return (b < a) ? b : a; }
inline double r_sign1(double s) { return std::copysign(1.0, s); }  // dsign(1.d0,s)
inline double r_exp(double x) { return std::exp(x); }               // dexp
inline double r_sin(double x) { return std::sin(x); }               // dsin
inline double r_cos(double x) { return std::cos(x); }               // dcos
inline double r_atan(double x) { return std::atan(x); }             // datan
inline double r_mod(double a, double p) { return std::fmod(a, p); } // dmod: a - int(a/p)*p, sign of a
inline int r_nint(double x) { return (int)std::lround(x); }         // nint: nearest, halves away from zero

template <class R> inline R r_max3(R a, R b, R c) { return r_max(r_max(a, b), c); }
template <class R> inline R r_min3(R a, R b, R c) { return r_min(r_min(a, b), c); }

// Fortran-style views: first index 1-based component(s), last index i with lower bound `lo`.
template <class R> struct V1 {  // a(i), i = lo..
  R* p; int lo;
  R& operator()(int i) const { return p[i - lo]; }
};
template <class R> struct V2 {  // a(m,i)  m=1..n0
  R* p; int n0; int lo;
  R& operator()(int m, int i) const { return p[(size_t)(i - lo) * n0 + (m - 1)]; }
};
template <class R> struct V3 {  // a(m,mw,i)  m=1..n0, mw=1..n1
  R* p; int n0; int n1; int lo;
  R& operator()(int m, int mw, int i) const {
    return p[((size_t)(i - lo) * n1 + (mw - 1)) * n0 + (m - 1)];
  }
};

enum RpId {
  RP_ADVECTION_2D = 1,
  RP_ACOUSTICS_2D = 2,
  RP_EULER_4WAVE_2D = 3,
  RP_SHALLOW_ROE_EFIX_2D = 4,
  RP_VC_ACOUSTICS_3D = 5,
  RP_VC_ACOUSTICS_2D = 7,       // (6 = euler_3D, clawref3.hpp)
  RP_VC_ADVECTION_2D = 8,
  RP_EULER_5WAVE_2D = 9,
  RP_PSYSTEM_2D = 10,
  RP_EULER_HLLE_2D = 11,
  RP_SHALLOW_HLLE_2D = 12,
  RP_KPP_2D = 18,
  RP_VC_ELASTICITY_2D = 19,
  RP_VC_ADVECTION_VECTOR_2D = 20,
  RP_ACOUSTICS_MAPPED_2D = 13,
  RP_SHALLOW_BATHYMETRY_FWAVE_2D = 14,
  RP_BURGERS_2D = 16,                // (15 = burgers_3D, clawref3.hpp)
};

// =====================================================================
// philim  -- pyclaw/src/pyclaw/classic/philim.f90:2-37
// =====================================================================
template <class R> inline R philim(R a, R b, int meth) {
  R r = b / a;                                        // philim.f90:14
  switch (meth) {
    case 1: return r_max(R(0.0), r_min(R(1.0), r));   // :18 minmod
    case 2: return r_max3(R(0.0), r_min(R(1.0), R(2.0) * r), r_min(R(2.0), r));  // :21 superbee
    case 3: return (r + r_abs(r)) / (R(1.0) + r_abs(r));                        // :24 van Leer
    case 4: { R c = (R(1.0) + r) / R(2.0);                                       // :27 MC
              return r_max(R(0.0), r_min3(c, R(2.0), R(2.0) * r)); }             // :28
    case 5: return r;                                                            // :31 Beam-Warming
  }
  return R(0.0);  // undefined in the reference (function result unset)
}

// =====================================================================
// limiter -- pyclaw/src/pyclaw/classic/limiter.f90:2-57
// =====================================================================
template <class R>
void limiter(int maxm, int num_eqn, int num_waves, int num_ghost, int mx,
             V3<R> wave, V2<R> s, const int* mthlim) {
  for (int mw = 1; mw <= num_waves; ++mw) {           // limiter.f90:28
    if (mthlim[mw - 1] == 0) continue;                 // :29
    R dotr = R(0.0);                                   // :30
    for (int i = 0; i <= mx + 1; ++i) {                // :31
      R wnorm2 = R(0.0);
      R dotl = dotr;
      dotr = R(0.0);
      for (int m = 1; m <= num_eqn; ++m) {             // :35-38
        wnorm2 = wnorm2 + wave(m, mw, i) * wave(m, mw, i);
        dotr = dotr + wave(m, mw, i) * wave(m, mw, i + 1);
      }
      if (i == 0) continue;                            // :39
      if (wnorm2 == R(0.0)) continue;                  // :40
      R wlimitr;
      if (s(mw, i) > R(0.0)) wlimitr = philim(wnorm2, dotl, mthlim[mw - 1]);  // :42-46
      else                   wlimitr = philim(wnorm2, dotr, mthlim[mw - 1]);
      for (int m = 1; m <= num_eqn; ++m) wave(m, mw, i) = wlimitr * wave(m, mw, i);  // :48-50
    }
  }
}

// Riemann-solver call signatures (f2py C pointers in the reference, classic/solver.py:516-521)
template <class R>
using Rpn2 = void (*)(int ixy, int maxm, int meqn, int mwaves, int maux, int mbc, int mx,
                      V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                      V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam);
template <class R>
using Rpt2 = void (*)(int ixy, int imp, int maxm, int meqn, int mwaves, int maux, int mbc, int mx,
                      V2<R> ql, V2<R> qr, V2<R> aux1, V2<R> aux2, V2<R> aux3,
                      V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam);

// =====================================================================
// advection -- riemann/src/rpn2_advection.f90, rpt2_advection.f90   cparam = (u, v)
// =====================================================================
template <class R>
void rpn2_advection(int ixy, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                    V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R u = R(cparam[0]), v = R(cparam[1]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // rpn2_advection.f90 loop 30
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = (ixy == 1) ? u : v;
    amdq(1, i) = r_min(s(1, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(s(1, i), R(0.0)) * wave(1, 1, i);
  }
}
template <class R>
void rpt2_advection(int ixy, int, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R>, V2<R>, V2<R>,
                    V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  const R u = R(cparam[0]), v = R(cparam[1]);
  R stran = (ixy == 1) ? v : u;                        // rpt2_advection.f90: stran
  R stranm = r_min(stran, R(0.0));
  R stranp = r_max(stran, R(0.0));
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 10
    bmasdq(1, i) = stranm * asdq(1, i);
    bpasdq(1, i) = stranp * asdq(1, i);
  }
}

// =====================================================================
// acoustics -- riemann/src/rpn2_acoustics.f90:79-106, rpt2_acoustics.f90:38-54
// cparam = (rho, bulk, cc, zz)
// =====================================================================
template <class R>
void rpn2_acoustics(int ixy, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                    V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R cc = R(cparam[2]), zz = R(cparam[3]);
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 20, rpn2_acoustics.f90:79-97
    R delta1 = ql(1, i) - qr(1, i - 1);
    R delta2 = ql(mu, i) - qr(mu, i - 1);
    R a1 = (-delta1 + zz * delta2) / (R(2.0) * zz);
    R a2 = (delta1 + zz * delta2) / (R(2.0) * zz);
    wave(1, 1, i) = -a1 * zz;
    wave(mu, 1, i) = a1;
    wave(mv, 1, i) = R(0.0);
    s(1, i) = -cc;
    wave(1, 2, i) = a2 * zz;
    wave(mu, 2, i) = a2;
    wave(mv, 2, i) = R(0.0);
    s(2, i) = cc;
  }
  for (int i = 2 - mbc; i <= mx + mbc; ++i)            // forall, :102-105
    for (int m = 1; m <= meqn; ++m) {
      amdq(m, i) = s(1, i) * wave(m, 1, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i);
    }
}
template <class R>
void rpt2_acoustics(int ixy, int, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R>, V2<R>, V2<R>,
                    V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  const R cc = R(cparam[2]), zz = R(cparam[3]);
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 20, rpt2_acoustics.f90:38-52
    R a1 = (-asdq(1, i) + zz * asdq(mv, i)) / (R(2.0) * zz);
    R a2 = (asdq(1, i) + zz * asdq(mv, i)) / (R(2.0) * zz);
    bmasdq(1, i) = cc * a1 * zz;
    bmasdq(mu, i) = R(0.0);
    bmasdq(mv, i) = -cc * a1;
    bpasdq(1, i) = cc * a2 * zz;
    bpasdq(mu, i) = R(0.0);
    bpasdq(mv, i) = cc * a2;
  }
}

// =====================================================================
// Euler 4-wave Roe solver with entropy fix
//   riemann/src/rpn2_euler_4wave.f90:108-309, rpt2_euler.f90:53-111      cparam = (gamma)
// =====================================================================
template <class R> struct EulerRoe {  // the u2v2,u,v,enth,a,g1a2,euv pencil arrays
  std::vector<R> buf; int lo, n;
  V1<R> u2v2, u, v, enth, a, g1a2, euv;
  EulerRoe(int maxm, int mbc) : buf((size_t)7 * (maxm + 2 * mbc + 2)), lo(1 - mbc), n(maxm + 2 * mbc + 2) {
    R* p = buf.data();
    u2v2 = {p, lo}; u = {p + n, lo}; v = {p + 2 * n, lo}; enth = {p + 3 * n, lo};
    a = {p + 4 * n, lo}; g1a2 = {p + 5 * n, lo}; euv = {p + 6 * n, lo};
  }
};

template <class R>
inline void euler_roe_averages(int mu, int mv, int mbc, int mx, V2<R> ql, V2<R> qr, R gamma1, EulerRoe<R>& w) {
  // loop 10 of rpn2_euler_4wave.f90:108-125 (identical in rpt2_euler.f90:53-70)
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R rhsqrtl = r_sqrt(qr(1, i - 1));
    R rhsqrtr = r_sqrt(ql(1, i));
    R pl = gamma1 * (qr(4, i - 1) - R(0.5) * (qr(2, i - 1) * qr(2, i - 1) + qr(3, i - 1) * qr(3, i - 1)) / qr(1, i - 1));
    R pr = gamma1 * (ql(4, i) - R(0.5) * (ql(2, i) * ql(2, i) + ql(3, i) * ql(3, i)) / ql(1, i));
    R rhsq2 = rhsqrtl + rhsqrtr;
    w.u(i) = (qr(mu, i - 1) / rhsqrtl + ql(mu, i) / rhsqrtr) / rhsq2;
    w.v(i) = (qr(mv, i - 1) / rhsqrtl + ql(mv, i) / rhsqrtr) / rhsq2;
    w.enth(i) = (((qr(4, i - 1) + pl) / rhsqrtl + (ql(4, i) + pr) / rhsqrtr)) / rhsq2;
    w.u2v2(i) = w.u(i) * w.u(i) + w.v(i) * w.v(i);
    R a2 = gamma1 * (w.enth(i) - R(.5) * w.u2v2(i));
    w.a(i) = r_sqrt(a2);
    w.g1a2(i) = gamma1 / a2;
    w.euv(i) = w.enth(i) - w.u2v2(i);
  }
}

template <class R>
void rpn2_euler_4wave(int ixy, int maxm, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                      V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);                     // rpn2_euler_4wave.f90:82
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }   // :88-94
  EulerRoe<R> w(maxm, mbc);
  euler_roe_averages(mu, mv, mbc, mx, ql, qr, gamma1, w);        // loop 10

  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 20, :131-175
    R delta1 = ql(1, i) - qr(1, i - 1);
    R delta2 = ql(mu, i) - qr(mu, i - 1);
    R delta3 = ql(mv, i) - qr(mv, i - 1);
    R delta4 = ql(4, i) - qr(4, i - 1);
    R a3 = w.g1a2(i) * (w.euv(i) * delta1 + w.u(i) * delta2 + w.v(i) * delta3 - delta4);
    R a2 = delta3 - w.v(i) * delta1;
    R a4 = (delta2 + (w.a(i) - w.u(i)) * delta1 - w.a(i) * a3) / (R(2.0) * w.a(i));
    R a1 = delta1 - a3 - a4;
    // acoustic
    wave(1, 1, i) = a1;
    wave(mu, 1, i) = a1 * (w.u(i) - w.a(i));
    wave(mv, 1, i) = a1 * w.v(i);
    wave(4, 1, i) = a1 * (w.enth(i) - w.u(i) * w.a(i));
    s(1, i) = w.u(i) - w.a(i);
    // shear
    wave(1, 2, i) = R(0.0);
    wave(mu, 2, i) = R(0.0);
    wave(mv, 2, i) = a2;
    wave(4, 2, i) = a2 * w.v(i);
    s(2, i) = w.u(i);
    // entropy
    wave(1, 3, i) = a3;
    wave(mu, 3, i) = a3 * w.u(i);
    wave(mv, 3, i) = a3 * w.v(i);
    wave(4, 3, i) = a3 * R(0.5) * w.u2v2(i);
    s(3, i) = w.u(i);
    // acoustic
    wave(1, 4, i) = a4;
    wave(mu, 4, i) = a4 * (w.u(i) + w.a(i));
    wave(mv, 4, i) = a4 * w.v(i);
    wave(4, 4, i) = a4 * (w.enth(i) + w.u(i) * w.a(i));
    s(4, i) = w.u(i) + w.a(i);
  }

  // efix = .true. (data statement :75) -> label 110, :216-296
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 200
    // check 1-wave
    R rhoim1 = qr(1, i - 1);
    R pim1 = gamma1 * (qr(4, i - 1) - R(0.5) * (qr(mu, i - 1) * qr(mu, i - 1) + qr(mv, i - 1) * qr(mv, i - 1)) / rhoim1);
    R cim1 = r_sqrt(gamma * pim1 / rhoim1);
    R s0 = qr(mu, i - 1) / rhoim1 - cim1;              // u-c in left state
    if (s0 >= R(0.0) && s(1, i) > R(0.0)) {            // :228 fully supersonic
      for (int m = 1; m <= meqn; ++m) amdq(m, i) = R(0.0);
      continue;
    }
    R rho1 = qr(1, i - 1) + wave(1, 1, i);
    R rhou1 = qr(mu, i - 1) + wave(mu, 1, i);
    R rhov1 = qr(mv, i - 1) + wave(mv, 1, i);
    R en1 = qr(4, i - 1) + wave(4, 1, i);
    R p1 = gamma1 * (en1 - R(0.5) * (rhou1 * rhou1 + rhov1 * rhov1) / rho1);
    R c1 = r_sqrt(gamma * p1 / rho1);
    R s1 = rhou1 / rho1 - c1;                          // u-c to right of 1-wave
    R sfract;
    if (s0 < R(0.0) && s1 > R(0.0)) {
      sfract = s0 * (s1 - s(1, i)) / (s1 - s0);        // :245 transonic rarefaction
    } else if (s(1, i) < R(0.0)) {
      sfract = s(1, i);
    } else {
      sfract = R(0.0);
    }
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = sfract * wave(m, 1, i);   // loop 120
    // check contact discontinuity
    if (s(2, i) >= R(0.0)) continue;                   // :260
    for (int m = 1; m <= meqn; ++m) {                  // loop 140
      amdq(m, i) = amdq(m, i) + s(2, i) * wave(m, 2, i);
      amdq(m, i) = amdq(m, i) + s(3, i) * wave(m, 3, i);
    }
    // check 4-wave
    R rhoi = ql(1, i);
    R pi = gamma1 * (ql(4, i) - R(0.5) * (ql(mu, i) * ql(mu, i) + ql(mv, i) * ql(mv, i)) / rhoi);
    R ci = r_sqrt(gamma * pi / rhoi);
    R s3 = ql(mu, i) / rhoi + ci;                      // u+c in right state
    R rho2 = ql(1, i) - wave(1, 4, i);
    R rhou2 = ql(mu, i) - wave(mu, 4, i);
    R rhov2 = ql(mv, i) - wave(mv, 4, i);
    R en2 = ql(4, i) - wave(4, 4, i);
    R p2 = gamma1 * (en2 - R(0.5) * (rhou2 * rhou2 + rhov2 * rhov2) / rho2);
    R c2 = r_sqrt(gamma * p2 / rho2);
    R s2 = rhou2 / rho2 + c2;                          // u+c to left of 4-wave
    if (s2 < R(0.0) && s3 > R(0.0)) {
      sfract = s2 * (s3 - s(4, i)) / (s3 - s2);        // :285
    } else if (s(4, i) < R(0.0)) {
      sfract = s(4, i);
    } else {
      continue;                                        // 4-wave is rightgoing
    }
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + sfract * wave(m, 4, i);  // loop 160
  }
  // apdq = df - amdq, :301-309
  for (int m = 1; m <= meqn; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      R df = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) df = df + s(mw, i) * wave(m, mw, i);
      apdq(m, i) = df - amdq(m, i);
    }
}

template <class R>
void rpt2_euler(int ixy, int, int maxm, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V2<R>,
                V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  EulerRoe<R> w(maxm, mbc);
  { RecomputeScope scope; euler_roe_averages(mu, mv, mbc, mx, ql, qr, gamma1, w); }   // rpt2_euler.f90:53-70
  R waveb[5][4]; R sb[4];                                         // waveb(4,3), sb(3): 1-based
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {                    // loop 20, :72-111
    R a3 = w.g1a2(i) * (w.euv(i) * asdq(1, i) + w.u(i) * asdq(mu, i) + w.v(i) * asdq(mv, i) - asdq(4, i));
    R a2 = asdq(mu, i) - w.u(i) * asdq(1, i);
    R a4 = (asdq(mv, i) + (w.a(i) - w.v(i)) * asdq(1, i) - w.a(i) * a3) / (R(2.0) * w.a(i));
    R a1 = asdq(1, i) - a3 - a4;
    waveb[1][1] = a1;
    waveb[mu][1] = a1 * w.u(i);
    waveb[mv][1] = a1 * (w.v(i) - w.a(i));
    waveb[4][1] = a1 * (w.enth(i) - w.v(i) * w.a(i));
    sb[1] = w.v(i) - w.a(i);
    waveb[1][2] = a3;
    waveb[mu][2] = a3 * w.u(i) + a2;
    waveb[mv][2] = a3 * w.v(i);
    waveb[4][2] = a3 * R(0.5) * w.u2v2(i) + a2 * w.u(i);
    sb[2] = w.v(i);
    waveb[1][3] = a4;
    waveb[mu][3] = a4 * w.u(i);
    waveb[mv][3] = a4 * (w.v(i) + w.a(i));
    waveb[4][3] = a4 * (w.enth(i) + w.v(i) * w.a(i));
    sb[3] = w.v(i) + w.a(i);
    for (int m = 1; m <= meqn; ++m) {
      bmasdq(m, i) = R(0.0);
      bpasdq(m, i) = R(0.0);
      for (int mw = 1; mw <= 3; ++mw) {
        bmasdq(m, i) = bmasdq(m, i) + r_min(sb[mw], R(0.0)) * waveb[m][mw];
        bpasdq(m, i) = bpasdq(m, i) + r_max(sb[mw], R(0.0)) * waveb[m][mw];
      }
    }
  }
}

// =====================================================================
// Shallow water Roe solver with entropy fix
//   riemann/src/rpn2_shallow_roe_with_efix.f90:91-241, rpt2_shallow_roe_with_efix.f90:42-86
//   cparam = (grav)
// =====================================================================
template <class R>
void rpn2_shallow_roe_efix(int ixy, int maxm, int, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                           V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R grav = R(cparam[0]);
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  const int n = maxm + 2 * mbc + 2, lo = 1 - mbc;
  std::vector<R> buf((size_t)4 * n);
  V1<R> u{buf.data(), lo}, v{buf.data() + n, lo}, a{buf.data() + 2 * n, lo}, h{buf.data() + 3 * n, lo};
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 10, :91-99
    h(i) = (qr(1, i - 1) + ql(1, i)) * R(0.50);
    R hsqrtl = r_sqrt(qr(1, i - 1));
    R hsqrtr = r_sqrt(ql(1, i));
    R hsq2 = hsqrtl + hsqrtr;
    u(i) = (qr(mu, i - 1) / hsqrtl + ql(mu, i) / hsqrtr) / hsq2;
    v(i) = (qr(mv, i - 1) / hsqrtl + ql(mv, i) / hsqrtr) / hsq2;
    a(i) = r_sqrt(grav * h(i));
  }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 20, :106-130
    R delta1 = ql(1, i) - qr(1, i - 1);
    R delta2 = ql(mu, i) - qr(mu, i - 1);
    R delta3 = ql(mv, i) - qr(mv, i - 1);
    R a1 = ((u(i) + a(i)) * delta1 - delta2) * (R(0.50) / a(i));
    R a2 = -v(i) * delta1 + delta3;
    R a3 = (-(u(i) - a(i)) * delta1 + delta2) * (R(0.50) / a(i));
    wave(1, 1, i) = a1;
    wave(mu, 1, i) = a1 * (u(i) - a(i));
    wave(mv, 1, i) = a1 * v(i);
    s(1, i) = u(i) - a(i);
    wave(1, 2, i) = R(0.0);
    wave(mu, 2, i) = R(0.0);
    wave(mv, 2, i) = a2;
    s(2, i) = u(i);
    wave(1, 3, i) = a3;
    wave(mu, 3, i) = a3 * (u(i) + a(i));
    wave(mv, 3, i) = a3 * v(i);
    s(3, i) = u(i) + a(i);
  }
  // efix = .true. -> label 110, :170-229
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 200
    R him1 = qr(1, i - 1);
    R s0 = qr(mu, i - 1) / him1 - r_sqrt(grav * him1);
    if (s0 > R(0.0) && s(1, i) > R(0.0)) {             // :175 (strict, unlike Euler)
      for (int m = 1; m <= 3; ++m) amdq(m, i) = R(0.0);
      continue;
    }
    R h1 = qr(1, i - 1) + wave(1, 1, i);
    R hu1 = qr(mu, i - 1) + wave(mu, 1, i);
    R s1 = hu1 / h1 - r_sqrt(grav * h1);
    R sfract;
    if (s0 < R(0.0) && s1 > R(0.0)) {
      sfract = s0 * ((s1 - s(1, i)) / (s1 - s0));      // :187 (parenthesised differently from Euler)
    } else if (s(1, i) < R(0.0)) {
      sfract = s(1, i);
    } else {
      sfract = R(0.0);
    }
    for (int m = 1; m <= 3; ++m) amdq(m, i) = sfract * wave(m, 1, i);
    if (s(2, i) > R(0.0)) continue;                    // :200 (strict)
    for (int m = 1; m <= 3; ++m) amdq(m, i) = amdq(m, i) + s(2, i) * wave(m, 2, i);
    R hi = ql(1, i);
    R s03 = ql(mu, i) / hi + r_sqrt(grav * hi);
    R h3 = ql(1, i) - wave(1, 3, i);
    R hu3 = ql(mu, i) - wave(mu, 3, i);
    R s3 = hu3 / h3 + r_sqrt(grav * h3);
    if (s3 < R(0.0) && s03 > R(0.0)) {
      sfract = s3 * ((s03 - s(3, i)) / (s03 - s3));
    } else if (s(3, i) < R(0.0)) {
      sfract = s(3, i);
    } else {
      continue;
    }
    for (int m = 1; m <= 3; ++m) amdq(m, i) = amdq(m, i) + sfract * wave(m, 3, i);
  }
  for (int m = 1; m <= 3; ++m)                          // :232-240
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      R df = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) df = df + s(mw, i) * wave(m, mw, i);
      apdq(m, i) = df - amdq(m, i);
    }
}

template <class R>
void rpt2_shallow_roe_efix(int ixy, int, int maxm, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr,
                           V2<R>, V2<R>, V2<R>, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  const R grav = R(cparam[0]);
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  const int n = maxm + 2 * mbc + 2, lo = 1 - mbc;
  std::vector<R> buf((size_t)4 * n);
  V1<R> u{buf.data(), lo}, v{buf.data() + n, lo}, a{buf.data() + 2 * n, lo}, h{buf.data() + 3 * n, lo};
  {
  RecomputeScope scope;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 10, rpt2_shallow...:42-50
    h(i) = (qr(1, i - 1) + ql(1, i)) * R(0.50);
    R hsqrtl = r_sqrt(qr(1, i - 1));
    R hsqrtr = r_sqrt(ql(1, i));
    R hsq2 = hsqrtl + hsqrtr;
    u(i) = (qr(mu, i - 1) / hsqrtl + ql(mu, i) / hsqrtr) / hsq2;
    v(i) = (qr(mv, i - 1) / hsqrtl + ql(mv, i) / hsqrtr) / hsq2;
    a(i) = r_sqrt(grav * h(i));
  }
  }
  R waveb[4][4]; R sb[4];
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 20, :52-84
    R a1 = (R(0.50) / a(i)) * ((v(i) + a(i)) * asdq(1, i) - asdq(mv, i));
    R a2 = asdq(mu, i) - u(i) * asdq(1, i);
    R a3 = (R(0.50) / a(i)) * (-(v(i) - a(i)) * asdq(1, i) + asdq(mv, i));
    waveb[1][1] = a1;
    waveb[mu][1] = a1 * u(i);
    waveb[mv][1] = a1 * (v(i) - a(i));
    sb[1] = v(i) - a(i);
    waveb[1][2] = R(0.0);
    waveb[mu][2] = a2;
    waveb[mv][2] = R(0.0);
    sb[2] = v(i);
    waveb[1][3] = a3;
    waveb[mu][3] = a3 * u(i);
    waveb[mv][3] = a3 * (v(i) + a(i));
    sb[3] = v(i) + a(i);
    for (int m = 1; m <= meqn; ++m) {
      bmasdq(m, i) = R(0.0);
      bpasdq(m, i) = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) {
        bmasdq(m, i) = bmasdq(m, i) + r_min(sb[mw], R(0.0)) * waveb[m][mw];
        bpasdq(m, i) = bpasdq(m, i) + r_max(sb[mw], R(0.0)) * waveb[m][mw];
      }
    }
  }
}

// =====================================================================
// vc_acoustics_2D -- riemann/src/rpn2_vc_acoustics.f90:40-97, rpt2_vc_acoustics.f90:27-80
// aux(1) = density, aux(2) = sound speed; no common block
// =====================================================================
template <class R>
void rpn2_vc_acoustics(int ixy, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                       V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }   // rpn2_vc_acoustics.f90:62-68
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {                   // loop 20, :73-91
    R delta1 = ql(1, i) - qr(1, i - 1);
    R delta2 = ql(mu, i) - qr(mu, i - 1);
    R zi = auxl(1, i) * auxl(2, i);
    R zim = auxr(1, i - 1) * auxr(2, i - 1);
    R a1 = (-delta1 + zi * delta2) / (zim + zi);
    R a2 = (delta1 + zim * delta2) / (zim + zi);
    wave(1, 1, i) = -a1 * zim;
    wave(mu, 1, i) = a1;
    wave(mv, 1, i) = R(0.0);
    s(1, i) = -auxr(2, i - 1);
    wave(1, 2, i) = a2 * zi;
    wave(mu, 2, i) = a2;
    wave(mv, 2, i) = R(0.0);
    s(2, i) = auxl(2, i);
  }
  for (int m = 1; m <= meqn; ++m)                               // :96-101
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = s(1, i) * wave(m, 1, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i);
    }
}
template <class R>
void rpt2_vc_acoustics(int ixy, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R> aux1, V2<R> aux2,
                       V2<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }   // rpt2_vc_acoustics.f90:27-33
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {                   // loop 20, :36-77
    const int i1 = (imp == 1) ? i - 1 : i;                      // :41-45
    R cm = aux1(2, i1), c = aux2(2, i1), cp = aux3(2, i1);
    (void)c;
    R zm = aux1(1, i1) * aux1(2, i1);
    R zz = aux2(1, i1) * aux2(2, i1);
    R zp = aux3(1, i1) * aux3(2, i1);
    R a1 = (-asdq(1, i) + asdq(mv, i) * zz) / (zm + zz);        // :58-59
    R a2 = (asdq(1, i) + asdq(mv, i) * zz) / (zz + zp);         // :61-62
    bmasdq(1, i) = cm * a1 * zm;                                // :66-68
    bmasdq(mu, i) = R(0.0);
    bmasdq(mv, i) = -cm * a1;
    bpasdq(1, i) = cp * a2 * zp;                                // :72-74
    bpasdq(mu, i) = R(0.0);
    bpasdq(mv, i) = cp * a2;
  }
}

// =====================================================================
// euler_5wave_2D -- riemann/src/rpn2_euler_5wave.f90:66-232, rpt2_euler_5wave.f90:52-118   cparam = (gamma)
// 4-equation Roe solver with entropy fix + a passively advected tracer (component 5, wave 5).
// =====================================================================
template <class R>
void rpn2_euler_5wave(int ixy, int, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                      V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);                      // :66
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }   // :69-75
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {           // :82-147 (wave = 0, s = 0 first, :78-79)
    for (int mw = 1; mw <= mwaves; ++mw) { s(mw, i) = R(0.0); for (int m = 1; m <= meqn; ++m) wave(m, mw, i) = R(0.0); }
    R rho_sqrtl = r_sqrt(qr(1, i - 1));
    R rho_sqrtr = r_sqrt(ql(1, i));
    R rho_sq2 = rho_sqrtl + rho_sqrtr;
    R u = (qr(mu, i - 1) / rho_sqrtl + ql(mu, i) / rho_sqrtr) / rho_sq2;
    R v = (qr(mv, i - 1) / rho_sqrtl + ql(mv, i) / rho_sqrtr) / rho_sq2;
    R u2v2 = u * u + v * v;
    R pl = gamma1 * (qr(4, i - 1) - R(0.5) * (qr(2, i - 1) * qr(2, i - 1) + qr(3, i - 1) * qr(3, i - 1)) / qr(1, i - 1));
    R pr = gamma1 * (ql(4, i) - R(0.5) * (ql(2, i) * ql(2, i) + ql(3, i) * ql(3, i)) / ql(1, i));
    R enth = (((qr(4, i - 1) + pl) / rho_sqrtl + (ql(4, i) + pr) / rho_sqrtr)) / rho_sq2;
    R c = r_sqrt(gamma1 * (enth - R(0.5) * u2v2));
    R g1c2 = gamma1 / (gamma1 * (enth - R(0.5) * u2v2));
    R euv = enth - u2v2;
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = ql(mu, i) - qr(mu, i - 1);
    R d3 = ql(mv, i) - qr(mv, i - 1);
    R d4 = ql(4, i) - qr(4, i - 1);
    R a3 = g1c2 * (euv * d1 + u * d2 + v * d3 - d4);
    R a2 = d3 - v * d1;
    R a4 = (d2 + (c - u) * d1 - c * a3) / (R(2.0) * c);
    R a1 = d1 - a3 - a4;
    wave(1, 1, i) = a1; wave(mu, 1, i) = a1 * (u - c); wave(mv, 1, i) = a1 * v; wave(4, 1, i) = a1 * (enth - u * c);
    wave(5, 1, i) = R(0.0);
    s(1, i) = u - c;
    wave(1, 2, i) = R(0.0); wave(mu, 2, i) = R(0.0); wave(mv, 2, i) = a2; wave(4, 2, i) = a2 * v; wave(5, 2, i) = R(0.0);
    s(2, i) = u;
    wave(1, 3, i) = a3; wave(mu, 3, i) = a3 * u; wave(mv, 3, i) = a3 * v; wave(4, 3, i) = a3 * R(0.5) * u2v2;
    wave(5, 3, i) = R(0.0);
    s(3, i) = u;
    wave(1, 4, i) = a4; wave(mu, 4, i) = a4 * (u + c); wave(mv, 4, i) = a4 * v; wave(4, 4, i) = a4 * (enth + u * c);
    wave(5, 4, i) = R(0.0);
    s(4, i) = u + c;
    wave(5, 5, i) = ql(5, i) - qr(5, i - 1);
    s(5, i) = u;
  }
  // efix = .true. (:60), :163-232
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    for (int m = 1; m <= meqn; ++m) { amdq(m, i) = R(0.0); apdq(m, i) = R(0.0); }
    R rhoim1 = qr(1, i - 1);
    R pim1 = gamma1 * (qr(4, i - 1) - R(0.5) * (qr(mu, i - 1) * qr(mu, i - 1) + qr(mv, i - 1) * qr(mv, i - 1)) / rhoim1);
    R cim1 = r_sqrt(gamma * pim1 / rhoim1);
    R s0 = qr(mu, i - 1) / rhoim1 - cim1;
    if (s0 >= R(0.0) && s(1, i) > R(0.0)) continue;     // :172-174
    R rho1 = qr(1, i - 1) + wave(1, 1, i);
    R rhou1 = qr(mu, i - 1) + wave(mu, 1, i);
    R rhov1 = qr(mv, i - 1) + wave(mv, 1, i);
    R en1 = qr(4, i - 1) + wave(4, 1, i);
    R p1 = gamma1 * (en1 - R(0.5) * (rhou1 * rhou1 + rhov1 * rhov1) / rho1);
    R c1 = r_sqrt(gamma * p1 / rho1);
    R s1 = rhou1 / rho1 - c1;
    R sfract;
    if (s0 < R(0.0) && s1 > R(0.0)) sfract = s0 * (s1 - s(1, i)) / (s1 - s0);
    else if (s(1, i) < R(0.0)) sfract = s(1, i);
    else sfract = R(0.0);
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = sfract * wave(m, 1, i);
    if (s(2, i) >= R(0.0)) continue;                    // :197-199
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + s(2, i) * wave(m, 2, i);
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + s(3, i) * wave(m, 3, i);
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + s(5, i) * wave(m, 5, i);
    R rhoi = ql(1, i);
    R pi = gamma1 * (ql(4, i) - R(0.5) * (ql(mu, i) * ql(mu, i) + ql(mv, i) * ql(mv, i)) / rhoi);
    R ci = r_sqrt(gamma * pi / rhoi);
    R s3 = ql(mu, i) / rhoi + ci;
    R rho2 = ql(1, i) - wave(1, 4, i);
    R rhou2 = ql(mu, i) - wave(mu, 4, i);
    R rhov2 = ql(mv, i) - wave(mv, 4, i);
    R en2 = ql(4, i) - wave(4, 4, i);
    R p2 = gamma1 * (en2 - R(0.5) * (rhou2 * rhou2 + rhov2 * rhov2) / rho2);
    R c2 = r_sqrt(gamma * p2 / rho2);
    R s2 = rhou2 / rho2 + c2;
    if (s2 < R(0.0) && s3 > R(0.0)) sfract = s2 * (s3 - s(4, i)) / (s3 - s2);
    else if (s(4, i) < R(0.0)) sfract = s(4, i);
    else continue;
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + sfract * wave(m, 4, i);
  }
  for (int m = 1; m <= meqn; ++m)                       // :223-231
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      R df = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) df = df + s(mw, i) * wave(m, mw, i);
      apdq(m, i) = df - amdq(m, i);
    }
}

template <class R>
void rpt2_euler_5wave(int ixy, int, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V2<R>,
                      V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  R waveb[6][5]; R sb[5];
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {           // rpt2_euler_5wave.f90:52-117
    for (int m = 1; m <= meqn; ++m) { bmasdq(m, i) = R(0.0); bpasdq(m, i) = R(0.0); }
    R u, v, u2v2, enth, c;
    {
      RecomputeScope scope;                             // the Roe averages rpn2 already produced
      R rho_sqrtl = r_sqrt(qr(1, i - 1));
      R rho_sqrtr = r_sqrt(ql(1, i));
      R rho_sq2 = rho_sqrtl + rho_sqrtr;
      u = (qr(mu, i - 1) / rho_sqrtl + ql(mu, i) / rho_sqrtr) / rho_sq2;
      v = (qr(mv, i - 1) / rho_sqrtl + ql(mv, i) / rho_sqrtr) / rho_sq2;
      u2v2 = u * u + v * v;
      R pl = gamma1 * (qr(4, i - 1) - R(0.5) * (qr(2, i - 1) * qr(2, i - 1) + qr(3, i - 1) * qr(3, i - 1)) / qr(1, i - 1));
      R pr = gamma1 * (ql(4, i) - R(0.5) * (ql(2, i) * ql(2, i) + ql(3, i) * ql(3, i)) / ql(1, i));
      enth = (((qr(4, i - 1) + pl) / rho_sqrtl + (ql(4, i) + pr) / rho_sqrtr)) / rho_sq2;
      c = r_sqrt(gamma1 * (enth - R(0.5) * u2v2));
    }
    R g1c2 = gamma1 / (c * c);                          // :71  (NOT the rpn2 expression)
    R euv = enth - u2v2;
    R a3 = g1c2 * (euv * asdq(1, i) + u * asdq(mu, i) + v * asdq(mv, i) - asdq(4, i));
    R a2 = asdq(mu, i) - u * asdq(1, i);
    R a4 = (asdq(mv, i) + (c - v) * asdq(1, i) - c * a3) / (R(2.0) * c);
    R a1 = asdq(1, i) - a3 - a4;
    waveb[1][1] = a1; waveb[mu][1] = a1 * u; waveb[mv][1] = a1 * (v - c); waveb[4][1] = a1 * (enth - v * c);
    waveb[5][1] = R(0.0);
    sb[1] = v - c;
    waveb[1][2] = a3; waveb[mu][2] = a3 * u + a2; waveb[mv][2] = a3 * v; waveb[4][2] = a3 * R(0.5) * u2v2 + a2 * u;
    waveb[5][2] = R(0.0);
    sb[2] = v;
    waveb[1][3] = a4; waveb[mu][3] = a4 * u; waveb[mv][3] = a4 * (v + c); waveb[4][3] = a4 * (enth + v * c);
    waveb[5][3] = R(0.0);
    sb[3] = v + c;
    waveb[1][4] = R(0.0); waveb[mu][4] = R(0.0); waveb[mv][4] = R(0.0); waveb[4][4] = R(0.0); waveb[5][4] = asdq(5, i);
    sb[4] = v;
    for (int mw = 1; mw <= 4; ++mw)                     // :112-115
      for (int m = 1; m <= meqn; ++m) {
        bmasdq(m, i) = bmasdq(m, i) + r_min(sb[mw], R(0.0)) * waveb[m][mw];
        bpasdq(m, i) = bpasdq(m, i) + r_max(sb[mw], R(0.0)) * waveb[m][mw];
      }
  }
}

// =====================================================================
// vc_advection_2D -- riemann/src/rpn2_vc_advection.f90:47-53, rpt2_vc_advection.f90:33-41
// colour equation; aux(1) = u at the left edge, aux(2) = v at the bottom edge of each cell
// =====================================================================
template <class R>
void rpn2_vc_advection(int ixy, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R>,
                       V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 30
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = auxl(ixy, i);
    amdq(1, i) = r_min(auxl(ixy, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(auxl(ixy, i), R(0.0)) * wave(1, 1, i);
  }
}
template <class R>
void rpt2_vc_advection(int ixy, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R>, V2<R> aux2,
                       V2<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  const int kv = 3 - ixy;                               // rpt2_vc_advection.f90:33
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {           // loop 10
    const int i1 = i - 2 + imp;
    bmasdq(1, i) = r_min(aux2(kv, i1), R(0.0)) * asdq(1, i);
    bpasdq(1, i) = r_max(aux3(kv, i1), R(0.0)) * asdq(1, i);
  }
}

// =====================================================================
// psystem_2D -- riemann/src/rpn2_psystem.f90:83-192, rpt2_psystem.f90:46-119   (no common block)
// nonlinear p-system (elasticity) with variable coefficients, f-wave form: 3 equations, 2 f-waves.
// aux(1) = density, aux(2) = elastic modulus, aux(3) = stress-strain relation id, aux(4) = strain (set by the
// example's before_step for the transverse solver).
// =====================================================================
template <class R> inline R psystem_sigma(R eps, R E, int linearity_mat) {       // rpn2_psystem.f90:156-173
  const R beta = R(5.0);
  switch (linearity_mat) {
    case 1: return E * eps;
    case 2: return r_exp(E * eps) - R(1.0);
    // `0.1` is a default-kind (single precision) literal in the Fortran source: its value is float(0.1)
    case 3: return R((double)0.1f) * E * eps + beta * (eps * eps * eps) * (E * E * E);
  }
  return R(0.0);      // select case without a default leaves the result undefined in the reference
}
template <class R> inline R psystem_sigmap(R eps, R E, int linearity_mat) {      // rpn2_psystem.f90:175-192
  switch (linearity_mat) {
    case 1: return E;
    case 2: return E * r_exp(E * eps);
    case 3: return R((double)0.1f) * E + R(15.0) * (eps * eps) * (E * E * E);     // 3*beta folds to 15
  }
  return R(0.0);
}
template <class R>
void rpn2_psystem(int ixy, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                  V3<R> fwave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {                   // loop 10, :83-150
    R pim = auxr(1, i - 1), pi = auxl(1, i), Eim = auxr(2, i - 1), Ei = auxl(2, i);
    R epsi = ql(1, i), urhoi = ql(2, i), vrhoi = ql(3, i);
    R epsim = qr(1, i - 1), urhoim = qr(2, i - 1), vrhoim = qr(3, i - 1);
    const int lin_i = r_nint(auxl(3, i)), lin_im = r_nint(auxr(3, i - 1));
    R sigmai = psystem_sigma(epsi, Ei, lin_i);
    R sigmaim = psystem_sigma(epsim, Eim, lin_im);
    R sigmapi = psystem_sigmap(epsi, Ei, lin_i);
    R sigmapim = psystem_sigmap(epsim, Eim, lin_im);
    R r11 = R(1.0) / r_sqrt(sigmapim * pim);                    // :106-107
    R r13 = -(R(1.0) / r_sqrt(sigmapi * pi));
    s(1, i) = -r_sqrt(sigmapim / pim);                          // :109-110
    s(2, i) = r_sqrt(sigmapi / pi);
    R dF1, dF2;
    if (ixy == 1) dF1 = -(urhoi / pi - urhoim / pim);           // :114 / :129
    else dF1 = -(vrhoi / pi - vrhoim / pim);
    dF2 = -(sigmai - sigmaim);
    R beta1 = (dF1 - r13 * dF2) / (r11 - r13);
    R beta3 = (-dF1 + r11 * dF2) / (r11 - r13);
    const int mu = (ixy == 1) ? 2 : 3, mv = (ixy == 1) ? 3 : 2;
    fwave(1, 1, i) = beta1 * r11;
    fwave(mu, 1, i) = beta1 * R(1.0);
    fwave(mv, 1, i) = beta1 * R(0.0);
    fwave(1, 2, i) = beta3 * r13;
    fwave(mu, 2, i) = beta3 * R(1.0);
    fwave(mv, 2, i) = beta3 * R(0.0);
    for (int m = 1; m <= 3; ++m) { amdq(m, i) = fwave(m, 1, i); apdq(m, i) = fwave(m, 2, i); }   // :142-147
  }
}
template <class R>
void rpt2_psystem(int ixy, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R> aux1, V2<R> aux2,
                  V2<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {                   // loop 10, :46-114
    const int ix = (imp == 1) ? i - 1 : i;
    R pjm = aux1(1, ix), pjp = aux3(1, ix);
    R Ejm = aux1(2, ix), Ejp = aux3(2, ix);
    const int lin_jm = r_nint(aux1(3, ix)), lin_jp = r_nint(aux3(3, ix));
    R epsjm = aux1(4, ix), epsjp = aux3(4, ix);
    R sigmapjm = psystem_sigmap(epsjm, Ejm, lin_jm);            // (sigmapj of the middle row is computed but unused)
    R sigmapjp = psystem_sigmap(epsjp, Ejp, lin_jp);
    (void)aux2;
    R r11 = R(1.0) / r_sqrt(sigmapjm * pjm);
    R r13 = -(R(1.0) / r_sqrt(sigmapjp * pjp));
    R s1 = -r_sqrt(sigmapjm / pjm);
    R s2 = r_sqrt(sigmapjp / pjp);
    // the transverse momentum is component 3 for an x-sweep and 2 for a y-sweep (:86-87 / :100-101)
    const int mt = (ixy == 1) ? 3 : 2, mn = (ixy == 1) ? 2 : 3;
    R gamma1 = (asdq(1, i) - r13 * asdq(mt, i)) / (r11 - r13);
    R gamma3 = (-asdq(1, i) + r11 * asdq(mt, i)) / (r11 - r13);
    bmasdq(1, i) = s1 * gamma1 * r11;
    bmasdq(mn, i) = s1 * gamma1 * R(0.0);
    bmasdq(mt, i) = s1 * gamma1 * R(1.0);
    bpasdq(1, i) = s2 * gamma3 * r13;
    bpasdq(mn, i) = s2 * gamma3 * R(0.0);
    bpasdq(mt, i) = s2 * gamma3 * R(1.0);
  }
}

// =====================================================================
// euler_hlle_2D -- riemann/src/rpn2_euler_hlle.f90:40-114 (+ rpt2_dummy.f90)        cparam = (gamma)
// two-wave HLLE solver.  Kept as written: H_r uses p_l (:60), c_l/c_r use gamma1*H - u**2/2 (:62-63).
// =====================================================================
template <class R>
void rpn2_euler_hlle(int ixy, int, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                     V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  const int rho = 1, E = 4;
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R u_l = qr(mu, i - 1) / qr(rho, i - 1);
    R u_r = ql(mu, i) / ql(rho, i);
    R v_l = qr(mv, i - 1) / qr(rho, i - 1);
    R v_r = ql(mv, i) / ql(rho, i);
    R p_l = gamma1 * (qr(E, i - 1) - qr(rho, i - 1) * (u_l * u_l + v_l * v_l) / R(2.0));
    R p_r = gamma1 * (ql(E, i) - ql(rho, i) * (u_r * u_r + v_r * v_r) / R(2.0));
    R H_l = (qr(E, i - 1) + p_l) / qr(rho, i - 1);
    R H_r = (ql(E, i) + p_l) / ql(rho, i);                               // p_l, as in the reference (:60)
    R c_l = r_sqrt(gamma1 * H_l - u_l * u_l / R(2.0));
    R c_r = r_sqrt(gamma1 * H_r - u_r * u_r / R(2.0));
    R E_l = qr(E, i - 1), E_r = ql(E, i);
    R rhsqrt_l = r_sqrt(qr(rho, i - 1));
    R rhsqrt_r = r_sqrt(ql(rho, i));
    R rhsq2 = rhsqrt_l + rhsqrt_r;
    R u_hat = (qr(mu, i - 1) / rhsqrt_l + ql(mu, i) / rhsqrt_r) / rhsq2;
    R v_hat = (qr(mv, i - 1) / rhsqrt_l + ql(mv, i) / rhsqrt_r) / rhsq2;
    R H_hat = (((qr(E, i - 1) + p_l) / rhsqrt_l + (ql(E, i) + p_r) / rhsqrt_r)) / rhsq2;
    R c_hat = r_sqrt(gamma1 * (H_hat - R(0.5) * (u_hat * u_hat + v_hat * v_hat)));
    R s1 = r_min(u_l - c_l, u_hat - c_hat);
    R s2 = r_max(u_r + c_r, u_hat + c_hat);
    R rho_m = (ql(mu, i) - qr(mu, i - 1) - s2 * ql(rho, i) + s1 * qr(rho, i - 1)) / (s1 - s2);
    R rhou_m = (ql(rho, i) * (u_r * u_r) - qr(rho, i - 1) * (u_l * u_l) + p_r - p_l - s2 * ql(mu, i)
                + s1 * qr(mu, i - 1)) / (s1 - s2);
    R rhov_m = (ql(mv, i) * u_r - qr(mv, i - 1) * u_l - s2 * ql(mv, i) + s1 * qr(mv, i - 1)) / (s1 - s2);
    R E_m = (u_r * (E_r + p_r) - u_l * (E_l + p_l) - s2 * E_r + s1 * E_l) / (s1 - s2);
    wave(rho, 1, i) = rho_m - qr(rho, i - 1);
    wave(mu, 1, i) = rhou_m - qr(mu, i - 1);
    wave(mv, 1, i) = rhov_m - qr(mv, i - 1);
    wave(E, 1, i) = E_m - qr(E, i - 1);
    s(1, i) = s1;
    wave(rho, 2, i) = ql(rho, i) - rho_m;
    wave(mu, 2, i) = ql(mu, i) - rhou_m;
    wave(mv, 2, i) = ql(mv, i) - rhov_m;
    wave(E, 2, i) = ql(E, i) - E_m;
    s(2, i) = s2;
  }
  for (int m = 1; m <= meqn; ++m)                                        // :101-112
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = R(0.0); apdq(m, i) = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) {
        if (s(mw, i) < R(0.0)) amdq(m, i) = amdq(m, i) + s(mw, i) * wave(m, mw, i);
        else apdq(m, i) = apdq(m, i) + s(mw, i) * wave(m, mw, i);
      }
    }
}

// =====================================================================
// shallow_hlle_2D -- riemann/src/rpn2_shallow_hlle.f90:35-100 (+ rpt2_dummy.f90)    cparam = (grav)
// =====================================================================
template <class R>
void rpn2_shallow_hlle(int ixy, int, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                       V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R grav = R(cparam[0]);
  const int depth = 1;
  int mu, mv;
  if (ixy == 1) { mu = 2; mv = 3; } else { mu = 3; mv = 2; }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R h_l = qr(depth, i - 1), h_r = ql(depth, i);
    R u_l = qr(mu, i - 1) / qr(depth, i - 1);
    R u_r = ql(mu, i) / ql(depth, i);
    R v_l = qr(mv, i - 1) / qr(depth, i - 1);
    R v_r = ql(mv, i) / ql(depth, i);
    R c_l = r_sqrt(grav * h_l);
    R c_r = r_sqrt(grav * h_r);
    R hsqrt_l = r_sqrt(qr(depth, i - 1));
    R hsqrt_r = r_sqrt(ql(depth, i));
    R hsq2 = hsqrt_l + hsqrt_r;
    R h_hat = R(0.5) * (h_l + h_r);
    R u_hat = (u_l * hsqrt_l + u_r * hsqrt_r) / hsq2;
    R v_hat = (v_l * hsqrt_l + v_r * hsqrt_r) / hsq2;
    (void)v_hat;
    R c_hat = r_sqrt(grav * h_hat);
    R s1 = r_min(u_l - c_l, u_hat - c_hat);
    R s2 = r_max(u_r + c_r, u_hat + c_hat);
    R h_m = (ql(mu, i) - qr(mu, i - 1) - s2 * ql(depth, i) + s1 * qr(depth, i - 1)) / (s1 - s2);
    R hu_m = (ql(mu, i) * (u_r - s2) - qr(mu, i - 1) * (u_l - s1) + R(0.5) * grav * (h_r * h_r - h_l * h_l)) / (s1 - s2);
    R hv_m = (ql(mv, i) * u_r - qr(mv, i - 1) * u_l - s2 * ql(mv, i) + s1 * qr(mv, i - 1)) / (s1 - s2);
    wave(depth, 1, i) = h_m - h_l;
    wave(mu, 1, i) = hu_m - qr(mu, i - 1);
    wave(mv, 1, i) = hv_m - qr(mv, i - 1);
    s(1, i) = s1;
    wave(depth, 2, i) = h_r - h_m;
    wave(mu, 2, i) = ql(mu, i) - hu_m;
    wave(mv, 2, i) = ql(mv, i) - hv_m;
    s(2, i) = s2;
  }
  for (int m = 1; m <= meqn; ++m)                                        // :87-98
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = R(0.0); apdq(m, i) = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) {
        if (s(mw, i) < R(0.0)) amdq(m, i) = amdq(m, i) + s(mw, i) * wave(m, mw, i);
        else apdq(m, i) = apdq(m, i) + s(mw, i) * wave(m, mw, i);
      }
    }
}

// rpt2_dummy.f90: prints an error and leaves its outputs unset; only legal with method(3) <= 0.  The oracle returns
// zeros and step2/step2ds refuse transverse modes for these solvers (select_rp2 sets `dummy_rpt`).
template <class R>
void rpt2_dummy(int, int, int, int meqn, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R>, V2<R>, V2<R>, V2<R>,
                V2<R> bmasdq, V2<R> bpasdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i)
    for (int m = 1; m <= meqn; ++m) { bmasdq(m, i) = R(0.0); bpasdq(m, i) = R(0.0); }
}

// =====================================================================
// acoustics_mapped_2D -- riemann/src/rpn2_acoustics_mapped.f90:60-131, rpt2_acoustics_mapped.f90:41-88
// acoustics on a mapped quadrilateral grid.  aux(1:3) = unit normal + length ratio of the left edge, aux(4:6) = the
// same of the bottom edge, aux(7) = capacity (cell area ratio), aux(8) = impedance, aux(9) = sound speed.
// =====================================================================
template <class R>
void rpn2_acoustics_mapped(int ixy, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R>,
                           V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  int inx, iny, ilenrat;
  if (ixy == 1) { inx = 1; iny = 2; ilenrat = 3; } else { inx = 4; iny = 5; ilenrat = 6; }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R alpha = auxl(inx, i), beta = auxl(iny, i);
    R unorl = alpha * ql(2, i) + beta * ql(3, i);
    R unorr = alpha * qr(2, i - 1) + beta * qr(3, i - 1);
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = unorl - unorr;
    R zi = auxl(8, i), zim = auxl(8, i - 1), ci = auxl(9, i), cim = auxl(9, i - 1);
    R a1 = (-d1 + zi * d2) / (zim + zi);
    R a2 = (d1 + zim * d2) / (zim + zi);
    wave(1, 1, i) = -a1 * zim; wave(2, 1, i) = a1 * alpha; wave(3, 1, i) = a1 * beta;
    s(1, i) = -cim * auxl(ilenrat, i);
    wave(1, 2, i) = a2 * zi; wave(2, 2, i) = a2 * alpha; wave(3, 2, i) = a2 * beta;
    s(2, i) = ci * auxl(ilenrat, i);
  }
  for (int m = 1; m <= meqn; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = s(1, i) * wave(m, 1, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i);
    }
}
template <class R>
void rpt2_acoustics_mapped(int ixy, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R> aux1, V2<R> aux2,
                           V2<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  int ixtran, iytran, ilenrat;
  if (ixy == 1) { ixtran = 4; iytran = 5; ilenrat = 6; } else { ixtran = 1; iytran = 2; ilenrat = 3; }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    const int i1 = (imp == 1) ? i - 1 : i;
    R asdqp = asdq(1, i);
    R cm = aux1(9, i1), cp = aux3(9, i1);
    R zm = aux1(8, i1), zz = aux2(8, i1), zp = aux3(8, i1);
    R xtran = aux3(ixtran, i), ytran = aux3(iytran, i);        // geometry of cell i, whatever imp is (:56-57)
    R asdqt = xtran * asdq(2, i) + ytran * asdq(3, i);
    R a2 = (asdqp + zz * asdqt) / (zz + zp);
    R bpasdqp = a2 * zp, bpasdqt = a2;
    bpasdq(1, i) = cp * bpasdqp;
    bpasdq(2, i) = cp * xtran * bpasdqt;
    bpasdq(3, i) = cp * ytran * bpasdqt;
    xtran = aux2(ixtran, i); ytran = aux2(iytran, i);
    asdqt = xtran * asdq(2, i) + ytran * asdq(3, i);
    R a1 = (-asdqp + zz * asdqt) / (zm + zz);
    R bmasdqp = -a1 * zm, bmasdqt = a1;
    bmasdq(1, i) = -cm * bmasdqp;
    bmasdq(2, i) = -cm * xtran * bmasdqt;
    bmasdq(3, i) = -cm * ytran * bmasdqt;
    for (int m = 1; m <= 3; ++m) {
      bmasdq(m, i) = bmasdq(m, i) * aux2(ilenrat, i);
      bpasdq(m, i) = bpasdq(m, i) * aux3(ilenrat, i);
    }
  }
}

// =====================================================================
// shallow_bathymetry_fwave_2D -- riemann/src/rpn2_shallow_bathymetry_fwave.f90:60-140
// cparam = (grav, dry_tolerance, sea_level); aux(1) = bathymetry; 3 f-waves.
// Its rpt2 (rpt2_shallow_bathymetry_fwave.f90:80) reads s(2) before assigning it (`s(2) = 0.5d0*(s(1)+s(2))`), so the
// transverse solver has no defined result: like rpt2_dummy it is refused (the reference example, shallow_2d/sill.py:48,
// runs dimensionally split: "No transverse solver available").
// =====================================================================
template <class R>
void rpn2_shallow_bathymetry_fwave(int ixy, int, int num_eqn, int num_waves, int, int mbc, int mx, V2<R> ql, V2<R> qr,
                                   V2<R> auxl, V2<R> auxr, V3<R> fwave, V2<R> s, V2<R> amdq, V2<R> apdq,
                                   const double* cparam) {
  const R grav = R(cparam[0]), dry_tolerance = R(cparam[1]);
  const int ni = (ixy == 1) ? 2 : 3, ti = (ixy == 1) ? 3 : 2;
  // `-1.0e-14` / `1.0e-14` in the fluctuation split are single-precision literals
  const R eps = R((double)1.0e-14f);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    for (int m = 1; m <= num_eqn; ++m) { amdq(m, i) = R(0.0); apdq(m, i) = R(0.0); }     // amdq = 0.d0; apdq = 0.d0
    R dry_l = (qr(1, i - 1) < dry_tolerance) ? R(0.0) : R(1.0);
    R dry_r = (ql(1, i) < dry_tolerance) ? R(0.0) : R(1.0);
    R hl = qr(1, i - 1) * dry_l;
    R ul = qr(ni, i - 1) / qr(1, i - 1) * dry_l;
    R vl = qr(ti, i - 1) / qr(1, i - 1) * dry_l;
    R phil = (R(0.5) * grav * (hl * hl) + hl * (ul * ul)) * dry_l;
    R bl = auxr(1, i - 1);
    R hr = ql(1, i) * dry_r;
    R ur = ql(ni, i) / ql(1, i) * dry_r;
    R vr = ql(ti, i) / ql(1, i) * dry_r;
    R phir = (R(0.5) * grav * (hr * hr) + hr * (ur * ur)) * dry_r;
    R br = auxl(1, i);
    R hbar = R(0.5) * (hr + hl);
    R uhat = (r_sqrt(hr) * ur + r_sqrt(hl) * ul) / (r_sqrt(hr) + r_sqrt(hl));
    R chat = r_sqrt(grav * hbar);
    R fd1 = hr * ur - hl * ul;
    R fd2 = phir - phil + grav * hbar * (br - bl);
    R fd3 = hr * ur * vr - hl * ul * vl;
    s(1, i) = r_min(uhat - chat, ul - r_sqrt(grav * hl));
    s(3, i) = r_max(uhat + chat, ur + r_sqrt(grav * hr));
    s(2, i) = R(0.5) * (s(1, i) + s(3, i));
    R beta1 = (s(3, i) * fd1 - fd2) / (s(3, i) - s(1, i));
    R beta3 = (fd2 - s(1, i) * fd1) / (s(3, i) - s(1, i));
    R beta2 = fd3 - beta1 * vl - beta3 * vr;
    fwave(1, 1, i) = beta1 * R(1.0); fwave(ni, 1, i) = beta1 * s(1, i); fwave(ti, 1, i) = beta1 * vl;
    fwave(1, 2, i) = beta2 * R(0.0); fwave(ni, 2, i) = beta2 * R(0.0);  fwave(ti, 2, i) = beta2 * R(1.0);
    fwave(1, 3, i) = beta3 * R(1.0); fwave(ni, 3, i) = beta3 * s(3, i); fwave(ti, 3, i) = beta3 * vr;
    for (int k = 1; k <= num_waves; ++k) {
      if (s(k, i) < -eps) {
        for (int m = 1; m <= num_eqn; ++m) amdq(m, i) = amdq(m, i) + fwave(m, k, i);
      } else if (s(k, i) > eps) {
        for (int m = 1; m <= num_eqn; ++m) apdq(m, i) = apdq(m, i) + fwave(m, k, i);
      } else {
        for (int m = 1; m <= num_eqn; ++m) {
          amdq(m, i) = amdq(m, i) + R(0.5) * fwave(m, k, i);
          apdq(m, i) = apdq(m, i) + R(0.5) * fwave(m, k, i);
        }
      }
    }
  }
}

// =====================================================================
// burgers_2D -- riemann/src/rpn2_burgers.f90:40-56, rpt2_burgers.f90:38-42   q_t + (q^2/2)_x + (q^2/2)_y = 0
// Parity unpinned: the reference ships no 2-D Burgers example or test.
// =====================================================================
template <class R>
void rpn2_burgers(int, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                  V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = R(0.5) * (qr(1, i - 1) + ql(1, i));
    amdq(1, i) = r_min(s(1, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(s(1, i), R(0.0)) * wave(1, 1, i);
    if (qr(1, i - 1) < R(0.0) && ql(1, i) > R(0.0)) {          // efix = .true.
      amdq(1, i) = -(R(0.5) * (qr(1, i - 1) * qr(1, i - 1)));
      apdq(1, i) = R(0.5) * (ql(1, i) * ql(1, i));
    }
  }
}
template <class R>
void rpt2_burgers(int, int, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V2<R>,
                  V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R sb = R(0.5) * (qr(1, i - 1) + ql(1, i));
    bmasdq(1, i) = r_min(sb, R(0.0)) * asdq(1, i);
    bpasdq(1, i) = r_max(sb, R(0.0)) * asdq(1, i);
  }
}

// ---------------------------------------------------------------------
// kpp_2D -- riemann/src/rpn2_kpp.f90:25-139 (+ rpt2_dummy.f90): q_t + sin(q)_x + cos(q)_y = 0, non-convex flux; the
// Godunov flux `fedge` is the minimum (ul < ur) or maximum of the flux over [ul, ur], found from where the interval sits
// relative to the extrema of sin / cos.
// ---------------------------------------------------------------------
template <class R>
void rpn2_kpp(int ixy, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave, V2<R> s,
              V2<R> amdq, V2<R> apdq, const double*) {
  const R pi = R(4.0) * r_atan(R(1.0));
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R ul = qr(1, i - 1), ur = ql(1, i);
    R fl, fr, fedge;
    if (ixy == 1) {
      fl = r_sin(ul); fr = r_sin(ur);
      if (ul < ur) {
        fedge = r_min(fl, fr);
        if (ur - ul > R(2.0) * pi) fedge = R(-1.0);
        else {
          R reml = r_mod(ul, R(2.0) * pi), remr = r_mod(ur, R(2.0) * pi);
          if (remr > R(1.5) * pi && reml < R(1.5) * pi) fedge = R(-1.0);
          else if (ur - ul > remr + R(0.5) * pi) fedge = R(-1.0);
        }
      } else {
        fedge = r_max(fl, fr);
        if (ul - ur > R(2.0) * pi) fedge = R(1.0);
        else {
          R reml = r_mod(ul, R(2.0) * pi), remr = r_mod(ur, R(2.0) * pi);
          if (reml > R(0.5) * pi && remr < R(0.5) * pi) fedge = R(1.0);
          else if (ul - ur > reml + R(1.5) * pi) fedge = R(1.0);
        }
      }
    } else {
      fl = r_cos(ul); fr = r_cos(ur);
      if (ul < ur) {
        fedge = r_min(fl, fr);
        if (ur - ul > R(2.0) * pi) fedge = R(-1.0);
        else {
          R reml = r_mod(ul, R(2.0) * pi), remr = r_mod(ur, R(2.0) * pi);
          if (remr > R(1.0) * pi && reml < R(1.0) * pi) fedge = R(-1.0);
          else if (ur - ul > remr + R(1.0) * pi) fedge = R(-1.0);
        }
      } else {
        fedge = r_max(fl, fr);
        if (ul - ur > R(2.0) * pi) fedge = R(1.0);
        else {
          R reml = r_mod(ul, R(2.0) * pi), remr = r_mod(ur, R(2.0) * pi);
          if (reml < remr) fedge = R(1.0);
        }
      }
    }
    if (ul != ur) s(1, i) = (fr - fl) / (ur - ul);
    else s(1, i) = R(0.0);
    wave(1, 1, i) = ur - ul;
    amdq(1, i) = fedge - fl;
    apdq(1, i) = fr - fedge;
  }
}

// ---------------------------------------------------------------------
// vc_elasticity_2D -- riemann/src/rpn2_vc_elasticity.f90:60-183, rpt2_vc_elasticity.f90:50-176
// q = (sigma11, sigma22, sigma12, u, v); aux = (density, lambda, mu, cp, cs); 4 waves: the two P-waves, then the two
// S-waves (not ordered by speed; the zero-speed fifth wave is not computed).  Where the Fortran STOPs on a zero
// determinant ("det=0 in rpn2") the division simply yields non-finite values here.
// ---------------------------------------------------------------------
template <class R>
void rpn2_vc_elasticity(int ixy, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                        V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  const int ksig11 = ixy == 1 ? 1 : 2, ksig22 = ixy == 1 ? 2 : 1, ku = ixy == 1 ? 4 : 5, kv = ixy == 1 ? 5 : 4;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R dsig11 = ql(ksig11, i) - qr(ksig11, i - 1);
    R dsig12 = ql(3, i) - qr(3, i - 1);
    R du = ql(ku, i) - qr(ku, i - 1);
    R dv = ql(kv, i) - qr(kv, i - 1);
    R alamr = auxl(2, i), amur = auxl(3, i);
    R bulkr = alamr + R(2.0) * amur;
    R cpr = auxl(4, i), csr = auxl(5, i);
    R alaml = auxr(2, i - 1), amul = auxr(3, i - 1);
    R bulkl = alaml + R(2.0) * amul;
    R cpl = auxr(4, i - 1), csl = auxr(5, i - 1);
    R det = bulkl * cpr + bulkr * cpl;
    R a1 = (cpr * dsig11 + bulkr * du) / det;
    R a2 = (cpl * dsig11 - bulkl * du) / det;
    det = amul * csr + amur * csl;
    R a3, a4;
    if (det == R(0.0)) { a3 = R(0.0); a4 = R(0.0); }          // no shear wave in a fluid
    else {
      a3 = (csr * dsig12 + amur * dv) / det;
      a4 = (csl * dsig12 - amul * dv) / det;
    }
    wave(ksig11, 1, i) = a1 * bulkl; wave(ksig22, 1, i) = a1 * alaml; wave(3, 1, i) = R(0.0);
    wave(ku, 1, i) = a1 * cpl;       wave(kv, 1, i) = R(0.0);         s(1, i) = -cpl;
    wave(ksig11, 2, i) = a2 * bulkr; wave(ksig22, 2, i) = a2 * alamr; wave(3, 2, i) = R(0.0);
    wave(ku, 2, i) = -a2 * cpr;      wave(kv, 2, i) = R(0.0);         s(2, i) = cpr;
    wave(ksig11, 3, i) = R(0.0);     wave(ksig22, 3, i) = R(0.0);     wave(3, 3, i) = a3 * amul;
    wave(ku, 3, i) = R(0.0);         wave(kv, 3, i) = a3 * csl;       s(3, i) = -csl;
    wave(ksig11, 4, i) = R(0.0);     wave(ksig22, 4, i) = R(0.0);     wave(3, 4, i) = a4 * amur;
    wave(ku, 4, i) = R(0.0);         wave(kv, 4, i) = -a4 * csr;      s(4, i) = csr;
    for (int m = 1; m <= meqn; ++m) {
      amdq(m, i) = s(1, i) * wave(m, 1, i) + s(3, i) * wave(m, 3, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i) + s(4, i) * wave(m, 4, i);
    }
  }
}

template <class R>
void rpt2_vc_elasticity(int ixy, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>, V2<R> aux1, V2<R> aux2,
                        V2<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  const int ksig11 = ixy == 1 ? 1 : 2, ksig22 = ixy == 1 ? 2 : 1, ku = ixy == 1 ? 4 : 5, kv = ixy == 1 ? 5 : 4;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    const int i1 = (imp == 1) ? i - 1 : i;
    R dsig22 = asdq(ksig22, i), dsig12 = asdq(3, i), du = asdq(ku, i), dv = asdq(kv, i);
    R alamm = aux1(2, i1), alam = aux2(2, i1), alamp = aux3(2, i1);
    R amum = aux1(3, i1), amu = aux2(3, i1), amup = aux3(3, i1);
    R bulkm = alamm + R(2.0) * amum, bulk = alam + R(2.0) * amu, bulkp = alamp + R(2.0) * amup;
    R cpm = aux1(4, i1), cp = aux2(4, i1), cpp = aux3(4, i1);
    R csm = aux1(5, i1), cs = aux2(5, i1), csp = aux3(5, i1);
    R det = bulkm * cp + bulk * cpm;
    R a1 = (cp * dsig22 + bulk * dv) / det;
    det = bulk * cpp + bulkp * cp;
    R a2 = (cp * dsig22 - bulk * dv) / det;
    R a3, a4;
    det = -(amum * cs + amu * csm);
    if (det == R(0.0)) a3 = R(0.0); else a3 = (cs * dsig12 + amu * du) / det;
    det = -(amu * csp + amup * cs);
    if (det == R(0.0)) a4 = R(0.0); else a4 = (cs * dsig12 - amu * du) / det;
    bmasdq(ksig11, i) = -cpm * a1 * alamm;
    bmasdq(ksig22, i) = -cpm * a1 * bulkm;
    bmasdq(3, i) = -csm * a3 * (-amum);
    bmasdq(ku, i) = -csm * a3 * (-csm);
    bmasdq(kv, i) = -cpm * a1 * (cpm);
    bpasdq(ksig11, i) = cpp * a2 * alamp;
    bpasdq(ksig22, i) = cpp * a2 * bulkp;
    bpasdq(3, i) = csp * a4 * (-amup);
    bpasdq(ku, i) = csp * a4 * (csp);
    bpasdq(kv, i) = cpp * a2 * (-cpp);
  }
}

// ---------------------------------------------------------------------
// vc_advection_vector_2D -- pyclaw/examples/advection_reaction_2d/rpn2_vc_advection_vector.f90:40-47 (+ rpt2_dummy.f90):
// the colour equation for a VECTOR of unknowns sharing one velocity field; aux(ixy) = edge velocity.  (The example
// builds it as the extension module `advection_2d`; it is not part of clawpack.riemann.)
// ---------------------------------------------------------------------
template <class R>
void rpn2_vc_advection_vector(int ixy, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R>,
                              V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    for (int m = 1; m <= meqn; ++m) wave(m, 1, i) = ql(m, i) - qr(m, i - 1);
    s(1, i) = auxl(ixy, i);
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = r_min(auxl(ixy, i), R(0.0)) * wave(m, 1, i);
    for (int m = 1; m <= meqn; ++m) apdq(m, i) = r_max(auxl(ixy, i), R(0.0)) * wave(m, 1, i);
  }
}

template <class R> inline bool select_rp2(int rp_id, Rpn2<R>& rpn, Rpt2<R>& rpt) {
  switch (rp_id) {
    case RP_ADVECTION_2D: rpn = rpn2_advection<R>; rpt = rpt2_advection<R>; return true;
    case RP_ACOUSTICS_2D: rpn = rpn2_acoustics<R>; rpt = rpt2_acoustics<R>; return true;
    case RP_EULER_4WAVE_2D: rpn = rpn2_euler_4wave<R>; rpt = rpt2_euler<R>; return true;
    case RP_SHALLOW_ROE_EFIX_2D: rpn = rpn2_shallow_roe_efix<R>; rpt = rpt2_shallow_roe_efix<R>; return true;
    case RP_VC_ACOUSTICS_2D: rpn = rpn2_vc_acoustics<R>; rpt = rpt2_vc_acoustics<R>; return true;
    case RP_VC_ADVECTION_2D: rpn = rpn2_vc_advection<R>; rpt = rpt2_vc_advection<R>; return true;
    case RP_EULER_5WAVE_2D: rpn = rpn2_euler_5wave<R>; rpt = rpt2_euler_5wave<R>; return true;
    case RP_PSYSTEM_2D: rpn = rpn2_psystem<R>; rpt = rpt2_psystem<R>; return true;
    case RP_EULER_HLLE_2D: rpn = rpn2_euler_hlle<R>; rpt = rpt2_dummy<R>; return true;
    case RP_SHALLOW_HLLE_2D: rpn = rpn2_shallow_hlle<R>; rpt = rpt2_dummy<R>; return true;
    case RP_ACOUSTICS_MAPPED_2D: rpn = rpn2_acoustics_mapped<R>; rpt = rpt2_acoustics_mapped<R>; return true;
    case RP_SHALLOW_BATHYMETRY_FWAVE_2D: rpn = rpn2_shallow_bathymetry_fwave<R>; rpt = rpt2_dummy<R>; return true;
    case RP_BURGERS_2D: rpn = rpn2_burgers<R>; rpt = rpt2_burgers<R>; return true;
    case RP_KPP_2D: rpn = rpn2_kpp<R>; rpt = rpt2_dummy<R>; return true;
    case RP_VC_ELASTICITY_2D: rpn = rpn2_vc_elasticity<R>; rpt = rpt2_vc_elasticity<R>; return true;
    case RP_VC_ADVECTION_VECTOR_2D: rpn = rpn2_vc_advection_vector<R>; rpt = rpt2_dummy<R>; return true;
  }
  return false;
}

// =====================================================================
// flux2 -- pyclaw/src/pyclaw/classic/flux2.f90:2-207
// =====================================================================
template <class R> struct Flux2Work {
  std::vector<R> buf;
  V2<R> q1d, qadd, fadd, amdq, apdq, cqxx, bmasdq, bpasdq, s, aux1, aux2, aux3;
  V3<R> wave, gadd;  // gadd(m, 1:2, i)
  V1<R> dtdx1d, dtdy1d;
  Flux2Work(int maxm, int meqn, int mwaves, int maux, int mbc) {
    const int n = maxm + 2 * mbc, lo = 1 - mbc;
    const int ma = maux > 0 ? maux : 1;
    size_t tot = (size_t)n * (8 * meqn + mwaves + meqn * mwaves + 2 * meqn + 3 * ma + 2);
    buf.assign(tot, R(0.0));
    R* p = buf.data();
    auto take = [&](size_t k) { R* r = p; p += k; return r; };
    q1d = {take((size_t)n * meqn), meqn, lo};   qadd = {take((size_t)n * meqn), meqn, lo};
    fadd = {take((size_t)n * meqn), meqn, lo};  amdq = {take((size_t)n * meqn), meqn, lo};
    apdq = {take((size_t)n * meqn), meqn, lo};  cqxx = {take((size_t)n * meqn), meqn, lo};
    bmasdq = {take((size_t)n * meqn), meqn, lo}; bpasdq = {take((size_t)n * meqn), meqn, lo};
    s = {take((size_t)n * mwaves), mwaves, lo};
    wave = {take((size_t)n * meqn * mwaves), meqn, mwaves, lo};
    gadd = {take((size_t)n * meqn * 2), meqn, 2, lo};
    aux1 = {take((size_t)n * ma), ma, lo}; aux2 = {take((size_t)n * ma), ma, lo}; aux3 = {take((size_t)n * ma), ma, lo};
    dtdx1d = {take(n), lo}; dtdy1d = {take(n), lo};
  }
};

template <class R>
void flux2(int ixy, int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx,
           V2<R> q1d, V1<R> dtdx1d, V2<R> aux1, V2<R> aux2, V2<R> aux3, const int* method, const int* mthlim,
           V2<R> qadd, V2<R> fadd, V3<R> gadd, R& cfl1d, V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, V2<R> cqxx,
           V2<R> bmasdq, V2<R> bpasdq, Rpn2<R> rpn2, Rpt2<R> rpt2, bool use_fwave, const double* cparam) {
  bool limit = false;                                  // flux2.f90:82-85
  for (int mw = 0; mw < num_waves; ++mw) if (mthlim[mw] > 0) limit = true;

  for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)  // :90-95
    for (int m = 1; m <= num_eqn; ++m) {
      qadd(m, i) = R(0.0); fadd(m, i) = R(0.0); gadd(m, 1, i) = R(0.0); gadd(m, 2, i) = R(0.0);
    }

  rpn2(ixy, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux2, aux2, wave, s, amdq, apdq, cparam);  // :101

  // forall with two statements: statement 1 for all (m,i), then statement 2 (:105-108)
  for (int i = 1; i <= mx + 1; ++i)
    for (int m = 1; m <= num_eqn; ++m) qadd(m, i) = qadd(m, i) - dtdx1d(i) * apdq(m, i);
  for (int i = 1; i <= mx + 1; ++i)
    for (int m = 1; m <= num_eqn; ++m) qadd(m, i - 1) = qadd(m, i - 1) - dtdx1d(i - 1) * amdq(m, i);

  cfl1d = R(0.0);                                      // :111-119
  for (int mw = 1; mw <= num_waves; ++mw)
    for (int i = 1; i <= mx + 1; ++i)
      cfl1d = r_max3(cfl1d, dtdx1d(i) * s(mw, i), -dtdx1d(i - 1) * s(mw, i));

  if (method[1] != 1) {                                // :121
    if (limit) limiter(maxm, num_eqn, num_waves, num_ghost, mx, wave, s, mthlim);  // :127
    for (int i = 2 - num_ghost; i <= mx + num_ghost; ++i) {                         // :134-160
      R dtdxave = R(0.5) * (dtdx1d(i - 1) + dtdx1d(i));
      for (int m = 1; m <= num_eqn; ++m) {
        cqxx(m, i) = R(0.0);
        for (int mw = 1; mw <= num_waves; ++mw) {
          if (!use_fwave)
            cqxx(m, i) = cqxx(m, i) + r_abs(s(mw, i)) * (R(1.0) - r_abs(s(mw, i)) * dtdxave) * wave(m, mw, i);
          else
            cqxx(m, i) = cqxx(m, i) + r_sign1(s(mw, i)) * (R(1.0) - r_abs(s(mw, i)) * dtdxave) * wave(m, mw, i);
        }
        fadd(m, i) = fadd(m, i) + R(0.5) * cqxx(m, i);
      }
    }
  }

  if (method[2] <= 0) return;                          // :165 no transverse propagation

  if (method[1] > 1 && method[2] == 2) {               // :167-173
    for (int i = 1; i <= mx + 1; ++i)
      for (int m = 1; m <= num_eqn; ++m) {
        amdq(m, i) = amdq(m, i) + cqxx(m, i);
        apdq(m, i) = apdq(m, i) - cqxx(m, i);
      }
  }

  rpt2(ixy, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, amdq, bmasdq, bpasdq, cparam);  // :181
  for (int i = 1; i <= mx + 1; ++i)                    // :185-190
    for (int m = 1; m <= num_eqn; ++m) {
      gadd(m, 1, i - 1) = gadd(m, 1, i - 1) - R(0.5) * dtdx1d(i - 1) * bmasdq(m, i);
      gadd(m, 2, i - 1) = gadd(m, 2, i - 1) - R(0.5) * dtdx1d(i - 1) * bpasdq(m, i);
    }
  rpt2(ixy, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, apdq, bmasdq, bpasdq, cparam);  // :194
  for (int i = 1; i <= mx + 1; ++i)                    // :198-203
    for (int m = 1; m <= num_eqn; ++m) {
      gadd(m, 1, i) = gadd(m, 1, i) - R(0.5) * dtdx1d(i) * bmasdq(m, i);
      gadd(m, 2, i) = gadd(m, 2, i) - R(0.5) * dtdx1d(i) * bpasdq(m, i);
    }
}

// 3-index view of q(m,i,j) / aux(ma,i,j) with ghost cells, Fortran order.
template <class R> struct Q2 {
  R* p; int n0; int nx; int mbc;  // nx = mx + 2*mbc
  R& operator()(int m, int i, int j) const {
    return p[((size_t)(j + mbc - 1) * nx + (i + mbc - 1)) * n0 + (m - 1)];
  }
};

// One x-pencil of step2 (row j).  apply: bit0 own row, bit1 row j-1, bit2 row j+1  (step2.f90:92-165)
template <class R>
inline void step2_xrow(int j, int apply, int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx,
                       Q2<R> qold, Q2<R> qnew, Q2<R> aux, R dtdx, R dtdy, const int* method, const int* mthlim,
                       Flux2Work<R>& w, Rpn2<R> rpn2, Rpt2<R> rpt2, bool use_fwave, const double* cparam, R& cfl1d) {
  const int index_capa = method[5];
  for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)        // :95-99
    for (int m = 1; m <= num_eqn; ++m) w.q1d(m, i) = qold(m, i, j);
  if (index_capa > 0)                                           // :101-105
    for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i) w.dtdx1d(i) = dtdx / aux(index_capa, i, j);
  if (num_aux > 0)                                              // :107-115
    for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)
      for (int ma = 1; ma <= num_aux; ++ma) {
        w.aux1(ma, i) = aux(ma, i, j - 1); w.aux2(ma, i) = aux(ma, i, j); w.aux3(ma, i) = aux(ma, i, j + 1);
      }
  flux2(1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, w.q1d, w.dtdx1d, w.aux1, w.aux2, w.aux3, method, mthlim,
        w.qadd, w.fadd, w.gadd, cfl1d, w.wave, w.s, w.amdq, w.apdq, w.cqxx, w.bmasdq, w.bpasdq, rpn2, rpt2, use_fwave, cparam);
  if (index_capa == 0) {                                        // :137-149
    for (int i = 1; i <= mx; ++i)
      for (int m = 1; m <= num_eqn; ++m) {
        if (apply & 1)
          qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, i) - dtdx * (w.fadd(m, i + 1) - w.fadd(m, i))
                          - dtdy * (w.gadd(m, 2, i) - w.gadd(m, 1, i));
        if (apply & 2) qnew(m, i, j - 1) = qnew(m, i, j - 1) - dtdy * w.gadd(m, 1, i);
        if (apply & 4) qnew(m, i, j + 1) = qnew(m, i, j + 1) + dtdy * w.gadd(m, 2, i);
      }
  } else {                                                      // :151-164
    for (int i = 1; i <= mx; ++i)
      for (int m = 1; m <= num_eqn; ++m) {
        if (apply & 1)
          qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, i)
                          - (dtdx * (w.fadd(m, i + 1) - w.fadd(m, i)) + dtdy * (w.gadd(m, 2, i) - w.gadd(m, 1, i)))
                            / aux(index_capa, i, j);
        if (apply & 2) qnew(m, i, j - 1) = qnew(m, i, j - 1) - dtdy * w.gadd(m, 1, i) / aux(index_capa, i, j - 1);
        if (apply & 4) qnew(m, i, j + 1) = qnew(m, i, j + 1) + dtdy * w.gadd(m, 2, i) / aux(index_capa, i, j + 1);
      }
  }
}

// One y-pencil of step2 (column i).  (step2.f90:172-247)
template <class R>
inline void step2_ycol(int i, int apply, int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int my,
                       Q2<R> qold, Q2<R> qnew, Q2<R> aux, R dtdx, R dtdy, const int* method, const int* mthlim,
                       Flux2Work<R>& w, Rpn2<R> rpn2, Rpt2<R> rpt2, bool use_fwave, const double* cparam, R& cfl1d) {
  const int index_capa = method[5];
  for (int j = 1 - num_ghost; j <= my + num_ghost; ++j)        // :174-178
    for (int m = 1; m <= num_eqn; ++m) w.q1d(m, j) = qold(m, i, j);
  if (index_capa > 0)                                           // :180-184
    for (int j = 1 - num_ghost; j <= my + num_ghost; ++j) w.dtdy1d(j) = dtdy / aux(index_capa, i, j);
  if (num_aux > 0)                                              // :186-194
    for (int j = 1 - num_ghost; j <= my + num_ghost; ++j)
      for (int ma = 1; ma <= num_aux; ++ma) {
        w.aux1(ma, j) = aux(ma, i - 1, j); w.aux2(ma, j) = aux(ma, i, j); w.aux3(ma, j) = aux(ma, i + 1, j);
      }
  flux2(2, maxm, num_eqn, num_waves, num_aux, num_ghost, my, w.q1d, w.dtdy1d, w.aux1, w.aux2, w.aux3, method, mthlim,
        w.qadd, w.fadd, w.gadd, cfl1d, w.wave, w.s, w.amdq, w.apdq, w.cqxx, w.bmasdq, w.bpasdq, rpn2, rpt2, use_fwave, cparam);
  if (index_capa == 0) {                                        // :216-230
    for (int j = 1; j <= my; ++j)
      for (int m = 1; m <= num_eqn; ++m) {
        if (apply & 1)
          qnew(m, i, j) = qnew(m, i, j) + (w.qadd(m, j) - dtdy * (w.fadd(m, j + 1) - w.fadd(m, j))
                                           - dtdx * (w.gadd(m, 2, j) - w.gadd(m, 1, j)));
        if (apply & 2) qnew(m, i - 1, j) = qnew(m, i - 1, j) - dtdx * w.gadd(m, 1, j);
        if (apply & 4) qnew(m, i + 1, j) = qnew(m, i + 1, j) + dtdx * w.gadd(m, 2, j);
      }
  } else {                                                      // :232-246
    for (int j = 1; j <= my; ++j)
      for (int m = 1; m <= num_eqn; ++m) {
        if (apply & 1)
          qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, j)
                          - (dtdy * (w.fadd(m, j + 1) - w.fadd(m, j)) + dtdx * (w.gadd(m, 2, j) - w.gadd(m, 1, j)))
                            / aux(index_capa, i, j);
        if (apply & 2) qnew(m, i - 1, j) = qnew(m, i - 1, j) - dtdx * w.gadd(m, 1, j) / aux(index_capa, i - 1, j);
        if (apply & 4) qnew(m, i + 1, j) = qnew(m, i + 1, j) + dtdx * w.gadd(m, 2, j) / aux(index_capa, i + 1, j);
      }
  }
}

// =====================================================================
// step2 -- pyclaw/src/pyclaw/classic/step2.f90:2-250
//
// nthreads == 1 is the reference loop order verbatim.  nthreads > 1 splits the
// pencil range into contiguous blocks, one per thread; each thread additionally
// evaluates the pencil just before / just after its block and applies only that
// pencil's contribution to the thread's own first / last row.  Every qnew entry
// therefore receives the same addends in the same order as in the serial loop,
// so the threaded result is bit-identical (checked in tests/test_oracle.py).
// =====================================================================
template <class R>
int step2(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my,
          R* qold_, R* qnew_, R* aux_, R dx, R dy, R dt, const int* method, const int* mthlim, R* cfl_out,
          bool use_fwave, int rp_id, const double* cparam, int nthreads) {
  Rpn2<R> rpn2; Rpt2<R> rpt2;
  if (!select_rp2<R>(rp_id, rpn2, rpt2)) return 1;
  // rpt2_dummy.f90 ("Error: Dummy transverse Riemann solver called!") leaves its outputs unset: refuse instead
  if ((rp_id == RP_EULER_HLLE_2D || rp_id == RP_SHALLOW_HLLE_2D || rp_id == RP_SHALLOW_BATHYMETRY_FWAVE_2D || rp_id == RP_KPP_2D ||
       rp_id == RP_VC_ADVECTION_VECTOR_2D) && method[2] > 0)
    return 2;
  if (maxm < mx || maxm < my) return 2;
  const int nx = mx + 2 * num_ghost;
  const int maux_eff = method[6];                       // step2.f90:75  num_aux = method(7)
  (void)num_aux;
  Q2<R> qold{qold_, num_eqn, nx, num_ghost}, qnew{qnew_, num_eqn, nx, num_ghost};
  Q2<R> aux{aux_, maux_eff > 0 ? maux_eff : 1, nx, num_ghost};
  const int index_capa = method[5];
  const R dtdx = dt / dx, dtdy = dt / dy;               // :77-78
  if (nthreads < 1) nthreads = 1;
  std::vector<R> cfls(nthreads, R(0.0));

  for (int pass = 0; pass < 2; ++pass) {                // pass 0: x-pencils j=0..my+1; pass 1: y-pencils i=0..mx+1
    const int npen = (pass == 0 ? my : mx) + 2;         // pencil ids 0..m+1
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
    {
      int tid = 0, nt = 1;
#ifdef _OPENMP
      if (nthreads > 1) { tid = omp_get_thread_num(); nt = omp_get_num_threads(); }
#endif
      Flux2Work<R> w(maxm, num_eqn, num_waves, maux_eff, num_ghost);
      if (index_capa == 0)                              // :80-85
        for (int i = 1 - num_ghost; i <= maxm + num_ghost; ++i) { w.dtdx1d(i) = dtdx; w.dtdy1d(i) = dtdy; }
      const int pa = (int)((long long)npen * tid / nt), pb = (int)((long long)npen * (tid + 1) / nt) - 1;
      R cfl = R(0.0), c1;
      auto run = [&](int p, int apply, bool count_cfl) {
        if (pass == 0)
          step2_xrow(p, apply, maxm, num_eqn, num_waves, maux_eff, num_ghost, mx, qold, qnew, aux, dtdx, dtdy, method,
                     mthlim, w, rpn2, rpt2, use_fwave, cparam, c1);
        else
          step2_ycol(p, apply, maxm, num_eqn, num_waves, maux_eff, num_ghost, my, qold, qnew, aux, dtdx, dtdy, method,
                     mthlim, w, rpn2, rpt2, use_fwave, cparam, c1);
        if (count_cfl) cfl = r_max(cfl, c1);            // :129, :211
      };
      if (pb >= pa) {
        if (pa > 0) run(pa - 1, 4, false);              // predecessor pencil: only its "+ gadd(2)" into my first row
        for (int p = pa; p <= pb; ++p) {
          int apply = 1;
          if (p > pa || pa == 0) apply |= 2;            // row p-1 is mine (or the outer ghost row of the serial loop)
          if (p < pb || pb == npen - 1) apply |= 4;
          run(p, apply, true);
        }
        if (pb < npen - 1) run(pb + 1, 2, false);       // successor pencil: only its "- gadd(1)" into my last row
      }
      cfls[tid] = r_max(cfls[tid], cfl);
    }
  }
  R cfl = R(0.0);
  for (int t = 0; t < nthreads; ++t) cfl = r_max(cfl, cfls[t]);
  *cfl_out = cfl;
  return 0;
}

// =====================================================================
// step2ds -- pyclaw/src/pyclaw/classic/step2ds.f90:2-264   (one split sweep, ids = 1 | 2)
// qold may alias qnew (classic/solver.py:531-533): a split pencil only updates itself.
// =====================================================================
template <class R>
int step2ds(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my,
            R* qold_, R* qnew_, R* aux_, R dx, R dy, R dt, const int* method, const int* mthlim, R* cfl_out,
            int ids, bool use_fwave, int rp_id, const double* cparam, int nthreads) {
  Rpn2<R> rpn2; Rpt2<R> rpt2;
  if (!select_rp2<R>(rp_id, rpn2, rpt2)) return 1;
  if (maxm < mx || maxm < my) return 2;
  if (ids != 1 && ids != 2) return 3;
  const int nx = mx + 2 * num_ghost;
  const int maux_eff = method[6];
  (void)num_aux;
  Q2<R> qold{qold_, num_eqn, nx, num_ghost}, qnew{qnew_, num_eqn, nx, num_ghost};
  Q2<R> aux{aux_, maux_eff > 0 ? maux_eff : 1, nx, num_ghost};
  const int index_capa = method[5];
  const R dtdx = dt / dx, dtdy = dt / dy;
  if (nthreads < 1) nthreads = 1;
  std::vector<R> cfls(nthreads, R(0.0));
  const int plo = 1 - num_ghost, phi = (ids == 1 ? my : mx) + num_ghost;   // :96, :180
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
  {
    int tid = 0;
#ifdef _OPENMP
    if (nthreads > 1) tid = omp_get_thread_num();
#endif
    Flux2Work<R> w(maxm, num_eqn, num_waves, maux_eff, num_ghost);
    if (index_capa == 0)
      for (int i = 1 - num_ghost; i <= maxm + num_ghost; ++i) { w.dtdx1d(i) = dtdx; w.dtdy1d(i) = dtdy; }
    R cfl = R(0.0), cfl1d;
#pragma omp for schedule(static)
    for (int p = plo; p <= phi; ++p) {
      if (ids == 1) {
        const int j = p;
        for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)
          for (int m = 1; m <= num_eqn; ++m) w.q1d(m, i) = qold(m, i, j);            // :98-100
        if (index_capa > 0)
          for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i) w.dtdx1d(i) = dtdx / aux(index_capa, i, j);
        if (maux_eff > 0)                                                             // :108-132
          for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)
            for (int ma = 1; ma <= maux_eff; ++ma) {
              w.aux2(ma, i) = aux(ma, i, j);
              if (j != 1 - num_ghost) w.aux1(ma, i) = aux(ma, i, j - 1);
              if (j != my + num_ghost) w.aux3(ma, i) = aux(ma, i, j + 1);
            }
        flux2(1, maxm, num_eqn, num_waves, maux_eff, num_ghost, mx, w.q1d, w.dtdx1d, w.aux1, w.aux2, w.aux3, method, mthlim,
              w.qadd, w.fadd, w.gadd, cfl1d, w.wave, w.s, w.amdq, w.apdq, w.cqxx, w.bmasdq, w.bpasdq, rpn2, rpt2,
              use_fwave, cparam);
        cfl = r_max(cfl, cfl1d);                                                      // :144
        for (int i = 1; i <= mx; ++i)
          for (int m = 1; m <= num_eqn; ++m) {
            if (index_capa == 0)                                                      // :149-152
              qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, i) - dtdx * (w.fadd(m, i + 1) - w.fadd(m, i));
            else                                                                      // :157-161
              qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, i) - dtdx * (w.fadd(m, i + 1) - w.fadd(m, i)) / aux(index_capa, i, j);
          }
      } else {
        const int i = p;
        for (int j = 1 - num_ghost; j <= my + num_ghost; ++j)
          for (int m = 1; m <= num_eqn; ++m) w.q1d(m, j) = qold(m, i, j);            // :182-184
        if (index_capa > 0)
          for (int j = 1 - num_ghost; j <= my + num_ghost; ++j) w.dtdy1d(j) = dtdy / aux(index_capa, i, j);
        if (maux_eff > 0)
          for (int j = 1 - num_ghost; j <= my + num_ghost; ++j)
            for (int ma = 1; ma <= maux_eff; ++ma) {
              w.aux2(ma, j) = aux(ma, i, j);
              if (i != 1 - num_ghost) w.aux1(ma, j) = aux(ma, i - 1, j);
              if (i != mx + num_ghost) w.aux3(ma, j) = aux(ma, i + 1, j);
            }
        flux2(2, maxm, num_eqn, num_waves, maux_eff, num_ghost, my, w.q1d, w.dtdy1d, w.aux1, w.aux2, w.aux3, method, mthlim,
              w.qadd, w.fadd, w.gadd, cfl1d, w.wave, w.s, w.amdq, w.apdq, w.cqxx, w.bmasdq, w.bpasdq, rpn2, rpt2,
              use_fwave, cparam);
        cfl = r_max(cfl, cfl1d);                                                      // :232
        for (int j = 1; j <= my; ++j)
          for (int m = 1; m <= num_eqn; ++m) {
            if (index_capa == 0)                                                      // :237-240
              qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, j) - dtdy * (w.fadd(m, j + 1) - w.fadd(m, j));
            else                                                                      // :246-250
              qnew(m, i, j) = qnew(m, i, j) + w.qadd(m, j) - dtdy * (w.fadd(m, j + 1) - w.fadd(m, j)) / aux(index_capa, i, j);
          }
      }
    }
    cfls[tid] = r_max(cfls[tid], cfl);
  }
  R cfl = R(0.0);
  for (int t = 0; t < nthreads; ++t) cfl = r_max(cfl, cfls[t]);
  *cfl_out = cfl;
  return 0;
}

}  // namespace clawref
