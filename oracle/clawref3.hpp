// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the reference's 3-D classic wave-propagation step
// (PyClaw f2py module `classic3`: step3.f90, step3ds.f90, flux3.f90 + limiter.f90/philim.f90)
// and of the 3-D Riemann solvers vc_acoustics_3D (rpn3/rpt3/rptt3_vc_acoustics.f90) and
// euler_3D (rpn3/rpt3/rptt3_euler.f90; used to pin flux3/step3/step3ds against the
// reference's stored euler_3d golden vectors).  Loop-for-loop, each routine cites file:line.
// Compile with g++ -O2 -ffp-contract=off (see clawref2.hpp).
//
// Deviations (documented): the `maxmrp = 1002` pencil cap of rpn3_euler.f90 and the Fortran
// `stop` paths (rho <= 0, a**2 <= 0, work-array size) are dropped / become error returns.
#pragma once
#include "clawref2.hpp"

namespace clawref {

enum RpId3 { RP_EULER_3D = 6, RP_BURGERS_3D = 15, RP_VC_ADVECTION_3D = 17 };

// gadd(m, side 1:2, slice -1:1, i)    (flux3.f90:121-122)
template <class R> struct V4 {
  R* p; int n0; int lo;
  R& operator()(int m, int side, int slice, int i) const {
    return p[(((size_t)(i - lo) * 3 + (slice + 1)) * 2 + (side - 1)) * n0 + (m - 1)];
  }
};
// auxN(ma, i, 1:3)                     (flux3.f90:154-156)
template <class R> struct A3 {
  R* p; int n0; int lo; int n;   // n = pencil length maxm + 2*mbc
  R& operator()(int ma, int i, int pl) const { return p[((size_t)(pl - 1) * n + (i - lo)) * n0 + (ma - 1)]; }
  V2<R> plane(int pl) const { return V2<R>{p + (size_t)(pl - 1) * n * n0, n0, lo}; }
};

template <class R>
using Rpn3 = void (*)(int ixyz, int maxm, int meqn, int mwaves, int maux, int mbc, int mx,
                      V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                      V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam);
template <class R>
using Rpt3 = void (*)(int ixyz, int icoor, int imp, int maxm, int meqn, int mwaves, int maux, int mbc, int mx,
                      V2<R> ql, V2<R> qr, A3<R> aux1, A3<R> aux2, A3<R> aux3,
                      V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam);
template <class R>
using Rptt3 = void (*)(int ixyz, int icoor, int imp, int impt, int maxm, int meqn, int mwaves, int maux, int mbc,
                       int mx, V2<R> ql, V2<R> qr, A3<R> aux1, A3<R> aux2, A3<R> aux3,
                       V2<R> bsasdq, V2<R> cmbsasdq, V2<R> cpbsasdq, const double* cparam);

inline void rp3_dirs(int ixyz, int& mu, int& mv, int& mw) {
  if (ixyz == 1) { mu = 2; mv = 3; mw = 4; }
  else if (ixyz == 2) { mu = 3; mv = 4; mw = 2; }
  else { mu = 4; mv = 2; mw = 3; }
}

// =====================================================================
// vc_acoustics_3D -- riemann/src/rpn3_vc_acoustics.f90:58-117
// aux(1) = impedance, aux(2) = sound speed; no COMMON block
// =====================================================================
template <class R>
void rpn3_vc_acoustics(int ixyz, int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                       V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  int mu, mv, mw;
  rp3_dirs(ixyz, mu, mv, mw);                          // :58-70
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 20, :81-105
    R delta1 = ql(1, i) - qr(1, i - 1);
    R delta2 = ql(mu, i) - qr(mu, i - 1);
    R zi = auxl(1, i);
    R zim = auxr(1, i - 1);
    R a1 = (-delta1 + zi * delta2) / (zim + zi);
    R a2 = (delta1 + zim * delta2) / (zim + zi);
    wave(1, 1, i) = -a1 * zim;
    wave(mu, 1, i) = a1;
    wave(mv, 1, i) = R(0.0);
    wave(mw, 1, i) = R(0.0);
    s(1, i) = -auxr(2, i - 1);
    wave(1, 2, i) = a2 * zi;
    wave(mu, 2, i) = a2;
    wave(mv, 2, i) = R(0.0);
    wave(mw, 2, i) = R(0.0);
    s(2, i) = auxl(2, i);
  }
  for (int m = 1; m <= meqn; ++m)                       // :111-116
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = s(1, i) * wave(m, 1, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i);
    }
}

// the six (zm,zz,zp,cm,c,cp) of one transverse solve
template <class R> struct ZC { R zm, zz, zp, cm, c, cp; };

template <class R>
inline void vc_acoustics_split(int iuvw, const ZC<R>& z, V2<R> asdq, V2<R> bm, V2<R> bp, int i) {
  // rpt3_vc_acoustics.f90:126-147 == rptt3_vc_acoustics.f90:157-177
  R a1 = (-asdq(1, i) + asdq(iuvw, i) * z.zz) / (z.zm + z.zz);
  R a2 = (asdq(1, i) + asdq(iuvw, i) * z.zz) / (z.zz + z.zp);
  bm(1, i) = z.cm * a1 * z.zm;
  bm(2, i) = R(0.0); bm(3, i) = R(0.0); bm(4, i) = R(0.0);
  bm(iuvw, i) = -z.cm * a1;
  bp(1, i) = z.cp * a2 * z.zp;
  bp(2, i) = R(0.0); bp(3, i) = R(0.0); bp(4, i) = R(0.0);
  bp(iuvw, i) = z.cp * a2;
}

// riemann/src/rpt3_vc_acoustics.f90:80-147
template <class R>
void rpt3_vc_acoustics(int ixyz, int icoor, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>,
                       A3<R> aux1, A3<R> aux2, A3<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  int iuvw = ixyz + icoor;                              // :80-81
  if (iuvw > 4) iuvw = iuvw - 3;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // loop 10
    const int i1 = i - 2 + imp;                         // :100
    ZC<R> z;
    if (icoor == 2) {                                   // :102-111
      z.zm = aux1(1, i1, 2); z.zz = aux2(1, i1, 2); z.zp = aux3(1, i1, 2);
      z.cm = aux1(2, i1, 2); z.c = aux2(2, i1, 2); z.cp = aux3(2, i1, 2);
    } else {                                            // :112-122
      z.zm = aux2(1, i1, 1); z.zz = aux2(1, i1, 2); z.zp = aux2(1, i1, 3);
      z.cm = aux2(2, i1, 1); z.c = aux2(2, i1, 2); z.cp = aux2(2, i1, 3);
    }
    vc_acoustics_split(iuvw, z, asdq, bmasdq, bpasdq, i);
  }
}

// riemann/src/rptt3_vc_acoustics.f90:86-177
template <class R>
void rptt3_vc_acoustics(int ixyz, int icoor, int imp, int impt, int, int, int, int, int mbc, int mx, V2<R>, V2<R>,
                        A3<R> aux1, A3<R> aux2, A3<R> aux3, V2<R> bsasdq, V2<R> cmbsasdq, V2<R> cpbsasdq,
                        const double*) {
  int iuvw = ixyz + icoor;
  if (iuvw > 4) iuvw = iuvw - 3;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    const int i1 = i - 2 + imp;                         // :106
    ZC<R> z;
    if (impt == 1) {                                    // :108-129
      if (icoor == 2) {
        z.zm = aux1(1, i1, 1); z.zz = aux2(1, i1, 1); z.zp = aux3(1, i1, 1);
        z.cm = aux1(2, i1, 1); z.c = aux2(2, i1, 1); z.cp = aux3(2, i1, 1);
      } else {
        z.zm = aux1(1, i1, 1); z.zz = aux1(1, i1, 2); z.zp = aux1(1, i1, 3);
        z.cm = aux1(2, i1, 1); z.c = aux1(2, i1, 2); z.cp = aux1(2, i1, 3);
      }
    } else {                                            // :130-152
      if (icoor == 2) {
        z.zm = aux1(1, i1, 3); z.zz = aux2(1, i1, 3); z.zp = aux3(1, i1, 3);
        z.cm = aux1(2, i1, 3); z.c = aux2(2, i1, 3); z.cp = aux3(2, i1, 3);
      } else {
        z.zm = aux3(1, i1, 1); z.zz = aux3(1, i1, 2); z.zp = aux3(1, i1, 3);
        z.cm = aux3(2, i1, 1); z.c = aux3(2, i1, 2); z.cp = aux3(2, i1, 3);
      }
    }
    vc_acoustics_split(iuvw, z, bsasdq, cmbsasdq, cpbsasdq, i);
  }
}

// =====================================================================
// euler_3D -- riemann/src/rpn3_euler.f90, rpt3_euler.f90, rptt3_euler.f90     cparam = (gamma)
// 5 equations, 3 waves, Roe solver with entropy fix
// =====================================================================
template <class R> struct Euler3Roe {
  std::vector<R> buf; int lo, n;
  V1<R> u2v2w2, u, v, w, enth, a, g1a2, euv;
  Euler3Roe(int maxm, int mbc) : buf((size_t)8 * (maxm + 2 * mbc + 2)), lo(1 - mbc), n(maxm + 2 * mbc + 2) {
    R* p = buf.data();
    u2v2w2 = {p, lo}; u = {p + n, lo}; v = {p + 2 * n, lo}; w = {p + 3 * n, lo}; enth = {p + 4 * n, lo};
    a = {p + 5 * n, lo}; g1a2 = {p + 6 * n, lo}; euv = {p + 7 * n, lo};
  }
};

template <class R>
inline void euler3_roe_averages(int mu, int mv, int mw, int mbc, int mx, V2<R> ql, V2<R> qr, R gamma1,
                                Euler3Roe<R>& r) {
  // rpn3_euler.f90:138-180 (identical loop in rpt3_euler.f90 / rptt3_euler.f90)
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R rhsqrtl = r_sqrt(qr(1, i - 1));
    R rhsqrtr = r_sqrt(ql(1, i));
    R pl = gamma1 * (qr(5, i - 1) - R(0.5) * (qr(mu, i - 1) * qr(mu, i - 1) + qr(mv, i - 1) * qr(mv, i - 1)
                                              + qr(mw, i - 1) * qr(mw, i - 1)) / qr(1, i - 1));
    R pr = gamma1 * (ql(5, i) - R(0.5) * (ql(mu, i) * ql(mu, i) + ql(mv, i) * ql(mv, i) + ql(mw, i) * ql(mw, i))
                                     / ql(1, i));
    R rhsq2 = rhsqrtl + rhsqrtr;
    r.u(i) = (qr(mu, i - 1) / rhsqrtl + ql(mu, i) / rhsqrtr) / rhsq2;
    r.v(i) = (qr(mv, i - 1) / rhsqrtl + ql(mv, i) / rhsqrtr) / rhsq2;
    r.w(i) = (qr(mw, i - 1) / rhsqrtl + ql(mw, i) / rhsqrtr) / rhsq2;
    r.enth(i) = (((qr(5, i - 1) + pl) / rhsqrtl + (ql(5, i) + pr) / rhsqrtr)) / rhsq2;
    r.u2v2w2(i) = r.u(i) * r.u(i) + r.v(i) * r.v(i) + r.w(i) * r.w(i);
    R asqrd = gamma1 * (r.enth(i) - R(.5) * r.u2v2w2(i));
    r.a(i) = r_sqrt(asqrd);
    r.g1a2(i) = gamma1 / asqrd;
    r.euv(i) = r.enth(i) - r.u2v2w2(i);
  }
}

template <class R>
void rpn3_euler(int ixyz, int maxm, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  int mu, mv, mw;
  rp3_dirs(ixyz, mu, mv, mw);
  // `wave = 0.d0; s = 0.d0` (whole arrays) -- every entry that is read is overwritten below
  Euler3Roe<R> r(maxm, mbc);
  euler3_roe_averages(mu, mv, mw, mbc, mx, ql, qr, gamma1, r);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // rpn3_euler.f90:184-221
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = ql(mu, i) - qr(mu, i - 1);
    R d3 = ql(mv, i) - qr(mv, i - 1);
    R d4 = ql(mw, i) - qr(mw, i - 1);
    R d5 = ql(5, i) - qr(5, i - 1);
    R al4 = r.g1a2(i) * (r.euv(i) * d1 + r.u(i) * d2 + r.v(i) * d3 + r.w(i) * d4 - d5);
    R al2 = d3 - r.v(i) * d1;
    R al3 = d4 - r.w(i) * d1;
    R al5 = (d2 + (r.a(i) - r.u(i)) * d1 - r.a(i) * al4) / (R(2.0) * r.a(i));
    R al1 = d1 - al4 - al5;
    wave(1, 1, i) = al1;
    wave(mu, 1, i) = al1 * (r.u(i) - r.a(i));
    wave(mv, 1, i) = al1 * r.v(i);
    wave(mw, 1, i) = al1 * r.w(i);
    wave(5, 1, i) = al1 * (r.enth(i) - r.u(i) * r.a(i));
    s(1, i) = r.u(i) - r.a(i);
    wave(1, 2, i) = al4;
    wave(mu, 2, i) = al4 * r.u(i);
    wave(mv, 2, i) = al4 * r.v(i) + al2;
    wave(mw, 2, i) = al4 * r.w(i) + al3;
    wave(5, 2, i) = al4 * R(0.5) * r.u2v2w2(i) + al2 * r.v(i) + al3 * r.w(i);
    s(2, i) = r.u(i);
    wave(1, 3, i) = al5;
    wave(mu, 3, i) = al5 * (r.u(i) + r.a(i));
    wave(mv, 3, i) = al5 * r.v(i);
    wave(mw, 3, i) = al5 * r.w(i);
    wave(5, 3, i) = al5 * (r.enth(i) + r.u(i) * r.a(i));
    s(3, i) = r.u(i) + r.a(i);
  }
  for (int i = 1 - mbc; i <= maxm + mbc; ++i)          // amdq = 0.d0; apdq = 0.d0 (whole arrays)
    for (int m = 1; m <= meqn; ++m) { amdq(m, i) = R(0.0); apdq(m, i) = R(0.0); }
  // efix = .TRUE.
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {          // :244-303
    R rhoim1 = qr(1, i - 1);
    R pim1 = gamma1 * (qr(5, i - 1) - R(0.5) * (qr(mu, i - 1) * qr(mu, i - 1) + qr(mv, i - 1) * qr(mv, i - 1)
                                                + qr(mw, i - 1) * qr(mw, i - 1)) / rhoim1);
    R cim1 = r_sqrt(gamma * pim1 / rhoim1);
    R s0 = qr(mu, i - 1) / rhoim1 - cim1;
    if (s0 >= R(0.0) && s(1, i) > R(0.0)) continue;    // everything is right-going
    R rho1 = qr(1, i - 1) + wave(1, 1, i);
    R rhou1 = qr(mu, i - 1) + wave(mu, 1, i);
    R rhov1 = qr(mv, i - 1) + wave(mv, 1, i);
    R rhow1 = qr(mw, i - 1) + wave(mw, 1, i);
    R en1 = qr(5, i - 1) + wave(5, 1, i);
    R p1 = gamma1 * (en1 - R(0.5) * (rhou1 * rhou1 + rhov1 * rhov1 + rhow1 * rhow1) / rho1);
    R c1 = r_sqrt(gamma * p1 / rho1);
    R s1 = rhou1 / rho1 - c1;
    R sfract;
    if (s0 < R(0.0) && s1 > R(0.0)) sfract = s0 * (s1 - s(1, i)) / (s1 - s0);
    else if (s(1, i) < R(0.0)) sfract = s(1, i);
    else sfract = R(0.0);
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = sfract * wave(m, 1, i);
    if (s(2, i) >= R(0.0)) continue;                   // 2- and 3-waves are right-going
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + s(2, i) * wave(m, 2, i);
    R rhoi = ql(1, i);
    R pi = gamma1 * (ql(5, i) - R(0.5) * (ql(mu, i) * ql(mu, i) + ql(mv, i) * ql(mv, i) + ql(mw, i) * ql(mw, i)) / rhoi);
    R ci = r_sqrt(gamma * pi / rhoi);
    R s3 = ql(mu, i) / rhoi + ci;
    R rho2 = ql(1, i) - wave(1, 3, i);
    R rhou2 = ql(mu, i) - wave(mu, 3, i);
    R rhov2 = ql(mv, i) - wave(mv, 3, i);
    R rhow2 = ql(mw, i) - wave(mw, 3, i);
    R en2 = ql(5, i) - wave(5, 3, i);
    R p2 = gamma1 * (en2 - R(0.5) * (rhou2 * rhou2 + rhov2 * rhov2 + rhow2 * rhow2) / rho2);
    R c2 = r_sqrt(gamma * p2 / rho2);
    R s2 = rhou2 / rho2 + c2;
    if (s2 < R(0.0) && s3 > R(0.0)) sfract = s2 * (s3 - s(3, i)) / (s3 - s2);
    else if (s(3, i) < R(0.0)) sfract = s(3, i);
    else continue;
    for (int m = 1; m <= meqn; ++m) amdq(m, i) = amdq(m, i) + sfract * wave(m, 3, i);
  }
  for (int m = 1; m <= meqn; ++m)                       // :306-314
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      R df = R(0.0);
      for (int mws = 1; mws <= mwaves; ++mws) df = df + s(mws, i) * wave(m, mws, i);
      apdq(m, i) = df - amdq(m, i);
    }
}

// body shared by rpt3_euler.f90:118-187 and rptt3_euler.f90 (identical arithmetic)
template <class R>
inline void euler3_transverse(int ixyz, int icoor, int maxm, int meqn, int mwaves, int mbc, int mx, V2<R> ql, V2<R> qr,
                              V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  int mu, mv, mw;
  rp3_dirs(ixyz, mu, mv, mw);
  Euler3Roe<R> r(maxm, mbc);
  { RecomputeScope scope; euler3_roe_averages(mu, mv, mw, mbc, mx, ql, qr, gamma1, r); }
  R waveb[6][4]; R sb[4];                               // waveb(5,3), sb(3), 1-based
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R al4 = r.g1a2(i) * (r.euv(i) * asdq(1, i) + r.u(i) * asdq(mu, i) + r.v(i) * asdq(mv, i)
                         + r.w(i) * asdq(mw, i) - asdq(5, i));
    R al2 = asdq(mu, i) - r.u(i) * asdq(1, i);
    if (icoor == 2) {
      R al3 = asdq(mw, i) - r.w(i) * asdq(1, i);
      R al5 = (asdq(mv, i) + (r.a(i) - r.v(i)) * asdq(1, i) - r.a(i) * al4) / (R(2.0) * r.a(i));
      R al1 = asdq(1, i) - al4 - al5;
      waveb[1][1] = al1;
      waveb[mu][1] = al1 * r.u(i);
      waveb[mv][1] = al1 * (r.v(i) - r.a(i));
      waveb[mw][1] = al1 * r.w(i);
      waveb[5][1] = al1 * (r.enth(i) - r.v(i) * r.a(i));
      sb[1] = r.v(i) - r.a(i);
      waveb[1][2] = al4;
      waveb[mu][2] = al2 + r.u(i) * al4;
      waveb[mv][2] = al4 * r.v(i);
      waveb[mw][2] = al3 + al4 * r.w(i);
      waveb[5][2] = al4 * R(0.5) * r.u2v2w2(i) + al2 * r.u(i) + al3 * r.w(i);
      sb[2] = r.v(i);
      waveb[1][3] = al5;
      waveb[mu][3] = al5 * r.u(i);
      waveb[mv][3] = al5 * (r.v(i) + r.a(i));
      waveb[mw][3] = al5 * r.w(i);
      waveb[5][3] = al5 * (r.enth(i) + r.v(i) * r.a(i));
      sb[3] = r.v(i) + r.a(i);
    } else {
      R al3 = asdq(mv, i) - r.v(i) * asdq(1, i);
      R al5 = (asdq(mw, i) + (r.a(i) - r.w(i)) * asdq(1, i) - r.a(i) * al4) / (R(2.0) * r.a(i));
      R al1 = asdq(1, i) - al4 - al5;
      waveb[1][1] = al1;
      waveb[mu][1] = al1 * r.u(i);
      waveb[mv][1] = al1 * r.v(i);
      waveb[mw][1] = al1 * (r.w(i) - r.a(i));
      waveb[5][1] = al1 * (r.enth(i) - r.w(i) * r.a(i));
      sb[1] = r.w(i) - r.a(i);
      waveb[1][2] = al4;
      waveb[mu][2] = al2 + al4 * r.u(i);
      waveb[mv][2] = al3 + al4 * r.v(i);
      waveb[mw][2] = al4 * r.w(i);
      waveb[5][2] = al4 * R(0.5) * r.u2v2w2(i) + al2 * r.u(i) + al3 * r.v(i);
      sb[2] = r.w(i);
      waveb[1][3] = al5;
      waveb[mu][3] = al5 * r.u(i);
      waveb[mv][3] = al5 * r.v(i);
      waveb[mw][3] = al5 * (r.w(i) + r.a(i));
      waveb[5][3] = al5 * (r.enth(i) + r.w(i) * r.a(i));
      sb[3] = r.w(i) + r.a(i);
    }
    for (int m = 1; m <= meqn; ++m) { bmasdq(m, i) = R(0.0); bpasdq(m, i) = R(0.0); }
    for (int mws = 1; mws <= mwaves; ++mws)
      for (int m = 1; m <= meqn; ++m) {
        bmasdq(m, i) = bmasdq(m, i) + r_min(sb[mws], R(0.0)) * waveb[m][mws];
        bpasdq(m, i) = bpasdq(m, i) + r_max(sb[mws], R(0.0)) * waveb[m][mws];
      }
  }
}

template <class R>
void rpt3_euler(int ixyz, int icoor, int, int maxm, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr,
                A3<R>, A3<R>, A3<R>, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  euler3_transverse(ixyz, icoor, maxm, meqn, mwaves, mbc, mx, ql, qr, asdq, bmasdq, bpasdq, cparam);
}
template <class R>
void rptt3_euler(int ixyz, int icoor, int, int, int maxm, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql,
                 V2<R> qr, A3<R>, A3<R>, A3<R>, V2<R> bsasdq, V2<R> cmbsasdq, V2<R> cpbsasdq, const double* cparam) {
  euler3_transverse(ixyz, icoor, maxm, meqn, mwaves, mbc, mx, ql, qr, bsasdq, cmbsasdq, cpbsasdq, cparam);
}

// =====================================================================
// burgers_3D -- riemann/src/rpn3_burgers.f90:50-68, rpt3_burgers.f90:55-61, rptt3_burgers.f90:60-66
// q_t + (a u^2/2)_x + (b u^2/2)_y + (c u^2/2)_z = 0;  cparam = coeff(3) = (a, b, c).  Parity unpinned: no reference
// example or test runs this solver.
// =====================================================================
template <class R>
void rpn3_burgers(int ixyz, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                  V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R coeff = R(cparam[ixyz - 1]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = coeff * R(0.5) * (qr(1, i - 1) + ql(1, i));
    amdq(1, i) = r_min(s(1, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(s(1, i), R(0.0)) * wave(1, 1, i);
    if (qr(1, i - 1) < R(0.0) && ql(1, i) > R(0.0)) {              // efix = .true. (:43, :61-66)
      amdq(1, i) = -(coeff * R(0.5) * (qr(1, i - 1) * qr(1, i - 1)));
      apdq(1, i) = coeff * R(0.5) * (ql(1, i) * ql(1, i));
    }
  }
}
template <class R>
void rpt3_burgers(int ixyz, int icoor, int, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr,
                  A3<R>, A3<R>, A3<R>, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double* cparam) {
  int iuvw = ixyz + icoor - 1;
  if (iuvw > 3) iuvw -= 3;
  const R coeff = R(cparam[iuvw - 1]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R sb = coeff * R(0.5) * (qr(1, i - 1) + ql(1, i));
    bmasdq(1, i) = r_min(sb, R(0.0)) * asdq(1, i);
    bpasdq(1, i) = r_max(sb, R(0.0)) * asdq(1, i);
  }
}
template <class R>
void rptt3_burgers(int ixyz, int icoor, int, int, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr,
                   A3<R>, A3<R>, A3<R>, V2<R> bsasdq, V2<R> cmbsasdq, V2<R> cpbsasdq, const double* cparam) {
  int iuvw = ixyz + icoor - 1;
  if (iuvw > 3) iuvw -= 3;
  const R coeff = R(cparam[iuvw - 1]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R sc = coeff * R(0.5) * (qr(1, i - 1) + ql(1, i));
    cmbsasdq(1, i) = r_min(sc, R(0.0)) * bsasdq(1, i);
    cpbsasdq(1, i) = r_max(sc, R(0.0)) * bsasdq(1, i);
  }
}

// =====================================================================
// vc_advection_3D -- riemann/src/rpn3_vc_advection.f90:52-57, rpt3_vc_advection.f90:60-72, rptt3_vc_advection.f90:62-84
// colour equation q_t + u q_x + v q_y + w q_z = 0 with edge velocities aux(1:3) = (u at the x-lower face, v at the
// y-lower face, w at the z-lower face of each cell).  Parity unpinned: no reference example or test runs it.
// =====================================================================
template <class R>
void rpn3_vc_advection(int ixyz, int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R>,
                       V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = auxl(ixyz, i);
    amdq(1, i) = r_min(auxl(ixyz, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(auxl(ixyz, i), R(0.0)) * wave(1, 1, i);
  }
}
template <class R>
void rpt3_vc_advection(int ixyz, int icoor, int imp, int, int, int, int, int mbc, int mx, V2<R>, V2<R>,
                       A3<R>, A3<R> aux2, A3<R> aux3, V2<R> asdq, V2<R> bmasdq, V2<R> bpasdq, const double*) {
  int iuvw = ixyz + icoor - 1;
  if (iuvw > 3) iuvw -= 3;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    const int i1 = i - 2 + imp;
    if (icoor == 2) {
      bmasdq(1, i) = r_min(aux2(iuvw, i1, 2), R(0.0)) * asdq(1, i);
      bpasdq(1, i) = r_max(aux3(iuvw, i1, 2), R(0.0)) * asdq(1, i);
    } else {
      bmasdq(1, i) = r_min(aux2(iuvw, i1, 2), R(0.0)) * asdq(1, i);
      bpasdq(1, i) = r_max(aux2(iuvw, i1, 3), R(0.0)) * asdq(1, i);
    }
  }
}
template <class R>
void rptt3_vc_advection(int ixyz, int icoor, int imp, int impt, int, int, int, int, int mbc, int mx, V2<R>, V2<R>,
                        A3<R> aux1, A3<R> aux2, A3<R> aux3, V2<R> bsasdq, V2<R> cmbsasdq, V2<R> cpbsasdq,
                        const double*) {
  int iuvw = ixyz + icoor - 1;
  if (iuvw > 3) iuvw -= 3;
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    const int i1 = i - 2 + imp;
    if (icoor == 2) {
      const int pl = (impt == 1) ? 1 : 3;
      cmbsasdq(1, i) = r_min(aux2(iuvw, i1, pl), R(0.0)) * bsasdq(1, i);
      cpbsasdq(1, i) = r_max(aux3(iuvw, i1, pl), R(0.0)) * bsasdq(1, i);
    } else {
      A3<R> a = (impt == 1) ? aux1 : aux3;
      cmbsasdq(1, i) = r_min(a(iuvw, i1, 2), R(0.0)) * bsasdq(1, i);
      cpbsasdq(1, i) = r_max(a(iuvw, i1, 3), R(0.0)) * bsasdq(1, i);
    }
  }
}

template <class R> inline bool select_rp3(int rp_id, Rpn3<R>& rpn, Rpt3<R>& rpt, Rptt3<R>& rptt) {
  switch (rp_id) {
    case RP_VC_ACOUSTICS_3D: rpn = rpn3_vc_acoustics<R>; rpt = rpt3_vc_acoustics<R>; rptt = rptt3_vc_acoustics<R>; return true;
    case RP_EULER_3D: rpn = rpn3_euler<R>; rpt = rpt3_euler<R>; rptt = rptt3_euler<R>; return true;
    case RP_BURGERS_3D: rpn = rpn3_burgers<R>; rpt = rpt3_burgers<R>; rptt = rptt3_burgers<R>; return true;
    case RP_VC_ADVECTION_3D: rpn = rpn3_vc_advection<R>; rpt = rpt3_vc_advection<R>; rptt = rptt3_vc_advection<R>; return true;
  }
  return false;
}

// =====================================================================
// flux3 work arrays -- step3.f90:59-92 (the 31*num_eqn + ... partition of `work`)
// =====================================================================
template <class R> struct Flux3Work {
  std::vector<R> buf;
  int n, lo;
  V2<R> q1d, qadd, fadd, s, amdq, apdq, cqxx;
  V2<R> bmamdq, bmapdq, bpamdq, bpapdq, cmamdq, cmapdq, cpamdq, cpapdq;
  V2<R> cmamdq2, cmapdq2, cpamdq2, cpapdq2;
  V2<R> bmcqxxp, bpcqxxp, bmcqxxm, bpcqxxm, cmcqxxp, cpcqxxp, cmcqxxm, cpcqxxm;
  V2<R> bmcmamdq, bmcmapdq, bpcmamdq, bpcmapdq, bmcpamdq, bmcpapdq, bpcpamdq, bpcpapdq;
  V3<R> wave;
  V4<R> gadd, hadd;
  A3<R> aux1, aux2, aux3;
  V1<R> dtdx1d, dtdy1d, dtdz1d;
  Flux3Work(int maxm, int meqn, int mwaves, int maux, int mbc) {
    n = maxm + 2 * mbc; lo = 1 - mbc;
    const int ma = maux > 0 ? maux : 1;
    size_t tot = (size_t)n * (34 * meqn + mwaves + meqn * mwaves + 12 * meqn + 9 * ma + 3);
    buf.assign(tot, R(0.0));
    R* p = buf.data();
    auto take = [&](size_t k) { R* r = p; p += k; return r; };
    auto v2 = [&]() { return V2<R>{take((size_t)n * meqn), meqn, lo}; };
    q1d = v2(); qadd = v2(); fadd = v2(); amdq = v2(); apdq = v2(); cqxx = v2();
    bmamdq = v2(); bmapdq = v2(); bpamdq = v2(); bpapdq = v2();
    cmamdq = v2(); cmapdq = v2(); cpamdq = v2(); cpapdq = v2();
    cmamdq2 = v2(); cmapdq2 = v2(); cpamdq2 = v2(); cpapdq2 = v2();
    bmcqxxp = v2(); bpcqxxp = v2(); bmcqxxm = v2(); bpcqxxm = v2();
    cmcqxxp = v2(); cpcqxxp = v2(); cmcqxxm = v2(); cpcqxxm = v2();
    bmcmamdq = v2(); bmcmapdq = v2(); bpcmamdq = v2(); bpcmapdq = v2();
    bmcpamdq = v2(); bmcpapdq = v2(); bpcpamdq = v2(); bpcpapdq = v2();
    s = {take((size_t)n * mwaves), mwaves, lo};
    wave = {take((size_t)n * meqn * mwaves), meqn, mwaves, lo};
    gadd = {take((size_t)n * meqn * 6), meqn, lo};
    hadd = {take((size_t)n * meqn * 6), meqn, lo};
    aux1 = {take((size_t)n * ma * 3), ma, lo, n};
    aux2 = {take((size_t)n * ma * 3), ma, lo, n};
    aux3 = {take((size_t)n * ma * 3), ma, lo, n};
    dtdx1d = {take(n), lo}; dtdy1d = {take(n), lo}; dtdz1d = {take(n), lo};
  }
};

// =====================================================================
// flux3 -- pyclaw/src/pyclaw/classic/flux3.f90:2-597
// =====================================================================
template <class R>
void flux3(int ixyz, int maxm, int num_eqn, int num_waves, int num_ghost, int mx, V1<R> dtdx1d, R dtdy, R dtdz,
           int num_aux, const int* method, const int* mthlim, Flux3Work<R>& w, R& cfl1d, bool use_fwave,
           Rpn3<R> rpn3, Rpt3<R> rpt3, Rptt3<R> rptt3, const double* cparam) {
  V2<R> q1d = w.q1d, qadd = w.qadd, fadd = w.fadd, amdq = w.amdq, apdq = w.apdq, cqxx = w.cqxx, s = w.s;
  V3<R> wave = w.wave;
  V4<R> gadd = w.gadd, hadd = w.hadd;
  A3<R> aux1 = w.aux1, aux2 = w.aux2, aux3 = w.aux3;

  bool limit = false;                                   // flux3.f90:168-171
  for (int mw = 0; mw < num_waves; ++mw) if (mthlim[mw] > 0) limit = true;

  for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)  // :176-183
    for (int m = 1; m <= num_eqn; ++m) {
      qadd(m, i) = R(0.0); fadd(m, i) = R(0.0);
      for (int sl = -1; sl <= 1; ++sl)
        for (int k = 1; k <= 2; ++k) { gadd(m, k, sl, i) = R(0.0); hadd(m, k, sl, i) = R(0.0); }
    }

  int m3, m4;                                            // :186-194
  if (method[2] < 0) { m3 = -1; m4 = 0; }
  else { m3 = method[2] / 10; m4 = method[2] - 10 * m3; }

  // :201-202  rpn3 sees the centre plane of aux2
  rpn3(ixyz, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux2.plane(2), aux2.plane(2), wave, s,
       amdq, apdq, cparam);

  // forall: statement 1 for all (m,i), then statement 2 (:205-208)
  for (int i = 1; i <= mx + 1; ++i)
    for (int m = 1; m <= num_eqn; ++m) qadd(m, i) = qadd(m, i) - dtdx1d(i) * apdq(m, i);
  for (int i = 1; i <= mx + 1; ++i)
    for (int m = 1; m <= num_eqn; ++m) qadd(m, i - 1) = qadd(m, i - 1) - dtdx1d(i - 1) * amdq(m, i);

  cfl1d = R(0.0);                                        // :211-218
  for (int i = 1; i <= mx + 1; ++i)
    for (int mw = 1; mw <= num_waves; ++mw)
      cfl1d = r_max3(cfl1d, dtdx1d(i) * s(mw, i), -dtdx1d(i - 1) * s(mw, i));

  if (method[1] != 1) {                                  // :220
    if (limit) limiter(maxm, num_eqn, num_waves, num_ghost, mx, wave, s, mthlim);   // :228
    for (int i = 2 - num_ghost; i <= mx + num_ghost; ++i) {                          // :231-259
      R dtdxave = R(0.5) * (dtdx1d(i - 1) + dtdx1d(i));
      for (int m = 1; m <= num_eqn; ++m) cqxx(m, i) = R(0.0);
      for (int mw = 1; mw <= num_waves; ++mw)
        for (int m = 1; m <= num_eqn; ++m) {
          if (!use_fwave)
            cqxx(m, i) = cqxx(m, i) + R(0.5) * r_abs(s(mw, i)) * (R(1.0) - r_abs(s(mw, i)) * dtdxave) * wave(m, mw, i);
          else
            cqxx(m, i) = cqxx(m, i) + R(0.5) * r_sign1(s(mw, i)) * (R(1.0) - r_abs(s(mw, i)) * dtdxave) * wave(m, mw, i);
        }
      for (int m = 1; m <= num_eqn; ++m) fadd(m, i) = fadd(m, i) + cqxx(m, i);
    }
  }

  if (m3 <= 0) return;                                   // :263 no transverse propagation

  // :272-291  split amdq / apdq in the y-like and z-like directions
  rpt3(ixyz, 2, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, amdq, w.bmamdq, w.bpamdq, cparam);
  rpt3(ixyz, 2, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, apdq, w.bmapdq, w.bpapdq, cparam);
  rpt3(ixyz, 3, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, amdq, w.cmamdq, w.cpamdq, cparam);
  rpt3(ixyz, 3, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, apdq, w.cmapdq, w.cpapdq, cparam);

  if (m3 == 2) {                                         // :296-344
    if (num_aux > 0) {
      rpt3(ixyz, 2, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, cqxx, w.bmcqxxm, w.bpcqxxm, cparam);
      rpt3(ixyz, 2, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, cqxx, w.bmcqxxp, w.bpcqxxp, cparam);
      rpt3(ixyz, 3, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, cqxx, w.cmcqxxm, w.cpcqxxm, cparam);
      rpt3(ixyz, 3, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, cqxx, w.cmcqxxp, w.cpcqxxp, cparam);
    } else {
      rpt3(ixyz, 2, 0, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, cqxx, w.bmcqxxm, w.bpcqxxm, cparam);
      rpt3(ixyz, 3, 0, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, cqxx, w.cmcqxxm, w.cpcqxxm, cparam);
      for (int i = 0; i <= mx + 2; ++i)
        for (int m = 1; m <= num_eqn; ++m) {
          w.bmcqxxp(m, i) = w.bmcqxxm(m, i); w.bpcqxxp(m, i) = w.bpcqxxm(m, i);
          w.cmcqxxp(m, i) = w.cmcqxxm(m, i); w.cpcqxxp(m, i) = w.cpcqxxm(m, i);
        }
    }
  }

  // ---- modify G fluxes in the y-like direction -------------------------------------------
  if (m4 == 1) {                                         // :353-367
    for (int i = 0; i <= mx + 2; ++i)
      for (int m = 1; m <= num_eqn; ++m) {
        w.cpapdq2(m, i) = w.cpapdq(m, i); w.cpamdq2(m, i) = w.cpamdq(m, i);
        w.cmapdq2(m, i) = w.cmapdq(m, i); w.cmamdq2(m, i) = w.cmamdq(m, i);
      }
  } else if (m4 == 2) {
    for (int i = 0; i <= mx + 2; ++i)
      for (int m = 1; m <= num_eqn; ++m) {
        w.cpapdq2(m, i) = w.cpapdq(m, i) - R(3.0) * w.cpcqxxp(m, i);
        w.cpamdq2(m, i) = w.cpamdq(m, i) + R(3.0) * w.cpcqxxm(m, i);
        w.cmapdq2(m, i) = w.cmapdq(m, i) - R(3.0) * w.cmcqxxp(m, i);
        w.cmamdq2(m, i) = w.cmamdq(m, i) + R(3.0) * w.cmcqxxm(m, i);
      }
  }
  if (m4 > 0) {                                          // :375-384
    rptt3(ixyz, 2, 2, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.cpapdq2, w.bmcpapdq, w.bpcpapdq, cparam);
    rptt3(ixyz, 2, 1, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.cpamdq2, w.bmcpamdq, w.bpcpamdq, cparam);
    rptt3(ixyz, 2, 2, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.cmapdq2, w.bmcmapdq, w.bpcmapdq, cparam);
    rptt3(ixyz, 2, 1, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.cmamdq2, w.bmcmamdq, w.bpcmamdq, cparam);
  }
  const R sixth = R(1.0) / R(6.0);
  for (int i = 1; i <= mx + 1; ++i)                      // :390-469
    for (int m = 1; m <= num_eqn; ++m) {
      gadd(m, 1, 0, i - 1) = gadd(m, 1, 0, i - 1) - R(0.5) * dtdx1d(i - 1) * w.bmamdq(m, i);
      gadd(m, 2, 0, i - 1) = gadd(m, 2, 0, i - 1) - R(0.5) * dtdx1d(i - 1) * w.bpamdq(m, i);
      gadd(m, 1, 0, i) = gadd(m, 1, 0, i) - R(0.5) * dtdx1d(i) * w.bmapdq(m, i);
      gadd(m, 2, 0, i) = gadd(m, 2, 0, i) - R(0.5) * dtdx1d(i) * w.bpapdq(m, i);
      if (m4 > 0) {
        gadd(m, 2, 0, i) = gadd(m, 2, 0, i) + sixth * dtdx1d(i) * dtdz * (w.bpcpapdq(m, i) - w.bpcmapdq(m, i));
        gadd(m, 1, 0, i) = gadd(m, 1, 0, i) + sixth * dtdx1d(i) * dtdz * (w.bmcpapdq(m, i) - w.bmcmapdq(m, i));
        gadd(m, 2, 1, i) = gadd(m, 2, 1, i) - sixth * dtdx1d(i) * dtdz * w.bpcpapdq(m, i);
        gadd(m, 1, 1, i) = gadd(m, 1, 1, i) - sixth * dtdx1d(i) * dtdz * w.bmcpapdq(m, i);
        gadd(m, 2, -1, i) = gadd(m, 2, -1, i) + sixth * dtdx1d(i) * dtdz * w.bpcmapdq(m, i);
        gadd(m, 1, -1, i) = gadd(m, 1, -1, i) + sixth * dtdx1d(i) * dtdz * w.bmcmapdq(m, i);
        gadd(m, 2, 0, i - 1) = gadd(m, 2, 0, i - 1) + sixth * dtdx1d(i - 1) * dtdz * (w.bpcpamdq(m, i) - w.bpcmamdq(m, i));
        gadd(m, 1, 0, i - 1) = gadd(m, 1, 0, i - 1) + sixth * dtdx1d(i - 1) * dtdz * (w.bmcpamdq(m, i) - w.bmcmamdq(m, i));
        gadd(m, 2, 1, i - 1) = gadd(m, 2, 1, i - 1) - sixth * dtdx1d(i - 1) * dtdz * w.bpcpamdq(m, i);
        gadd(m, 1, 1, i - 1) = gadd(m, 1, 1, i - 1) - sixth * dtdx1d(i - 1) * dtdz * w.bmcpamdq(m, i);
        gadd(m, 2, -1, i - 1) = gadd(m, 2, -1, i - 1) + sixth * dtdx1d(i - 1) * dtdz * w.bpcmamdq(m, i);
        gadd(m, 1, -1, i - 1) = gadd(m, 1, -1, i - 1) + sixth * dtdx1d(i - 1) * dtdz * w.bmcmamdq(m, i);
      }
      if (m3 >= 2) {
        gadd(m, 2, 0, i) = gadd(m, 2, 0, i) + dtdx1d(i) * w.bpcqxxp(m, i);
        gadd(m, 1, 0, i) = gadd(m, 1, 0, i) + dtdx1d(i) * w.bmcqxxp(m, i);
        gadd(m, 2, 0, i - 1) = gadd(m, 2, 0, i - 1) - dtdx1d(i - 1) * w.bpcqxxm(m, i);
        gadd(m, 1, 0, i - 1) = gadd(m, 1, 0, i - 1) - dtdx1d(i - 1) * w.bmcqxxm(m, i);
      }
    }

  // ---- modify H fluxes in the z-like direction -------------------------------------------
  if (m4 == 2) {                                         // :479-486
    for (int i = 0; i <= mx + 2; ++i)
      for (int m = 1; m <= num_eqn; ++m) {
        w.bpapdq(m, i) = w.bpapdq(m, i) - R(3.0) * w.bpcqxxp(m, i);
        w.bpamdq(m, i) = w.bpamdq(m, i) + R(3.0) * w.bpcqxxm(m, i);
        w.bmapdq(m, i) = w.bmapdq(m, i) - R(3.0) * w.bmcqxxp(m, i);
        w.bmamdq(m, i) = w.bmamdq(m, i) + R(3.0) * w.bmcqxxm(m, i);
      }
  }
  if (m4 > 0) {                                          // :498-507
    rptt3(ixyz, 3, 2, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.bpapdq, w.bmcpapdq, w.bpcpapdq, cparam);
    rptt3(ixyz, 3, 1, 2, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.bpamdq, w.bmcpamdq, w.bpcpamdq, cparam);
    rptt3(ixyz, 3, 2, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.bmapdq, w.bmcmapdq, w.bpcmapdq, cparam);
    rptt3(ixyz, 3, 1, 1, maxm, num_eqn, num_waves, num_aux, num_ghost, mx, q1d, q1d, aux1, aux2, aux3, w.bmamdq, w.bmcmamdq, w.bpcmamdq, cparam);
  }
  for (int i = 1; i <= mx + 1; ++i)                      // :513-594
    for (int m = 1; m <= num_eqn; ++m) {
      hadd(m, 1, 0, i - 1) = hadd(m, 1, 0, i - 1) - R(0.5) * dtdx1d(i - 1) * w.cmamdq(m, i);
      hadd(m, 2, 0, i - 1) = hadd(m, 2, 0, i - 1) - R(0.5) * dtdx1d(i - 1) * w.cpamdq(m, i);
      hadd(m, 1, 0, i) = hadd(m, 1, 0, i) - R(0.5) * dtdx1d(i) * w.cmapdq(m, i);
      hadd(m, 2, 0, i) = hadd(m, 2, 0, i) - R(0.5) * dtdx1d(i) * w.cpapdq(m, i);
      if (m4 > 0) {
        hadd(m, 2, 0, i) = hadd(m, 2, 0, i) + sixth * dtdx1d(i) * dtdy * (w.bpcpapdq(m, i) - w.bpcmapdq(m, i));
        hadd(m, 1, 0, i) = hadd(m, 1, 0, i) + sixth * dtdx1d(i) * dtdy * (w.bmcpapdq(m, i) - w.bmcmapdq(m, i));
        hadd(m, 2, 1, i) = hadd(m, 2, 1, i) - sixth * dtdx1d(i) * dtdy * w.bpcpapdq(m, i);
        hadd(m, 1, 1, i) = hadd(m, 1, 1, i) - sixth * dtdx1d(i) * dtdy * w.bmcpapdq(m, i);
        hadd(m, 2, -1, i) = hadd(m, 2, -1, i) + sixth * dtdx1d(i) * dtdy * w.bpcmapdq(m, i);
        hadd(m, 1, -1, i) = hadd(m, 1, -1, i) + sixth * dtdx1d(i) * dtdy * w.bmcmapdq(m, i);
        hadd(m, 2, 0, i - 1) = hadd(m, 2, 0, i - 1) + sixth * dtdx1d(i - 1) * dtdy * (w.bpcpamdq(m, i) - w.bpcmamdq(m, i));
        hadd(m, 1, 0, i - 1) = hadd(m, 1, 0, i - 1) + sixth * dtdx1d(i - 1) * dtdy * (w.bmcpamdq(m, i) - w.bmcmamdq(m, i));
        hadd(m, 2, 1, i - 1) = hadd(m, 2, 1, i - 1) - sixth * dtdx1d(i - 1) * dtdy * w.bpcpamdq(m, i);
        hadd(m, 1, 1, i - 1) = hadd(m, 1, 1, i - 1) - sixth * dtdx1d(i - 1) * dtdy * w.bmcpamdq(m, i);
        hadd(m, 2, -1, i - 1) = hadd(m, 2, -1, i - 1) + sixth * dtdx1d(i - 1) * dtdy * w.bpcmamdq(m, i);
        hadd(m, 1, -1, i - 1) = hadd(m, 1, -1, i - 1) + sixth * dtdx1d(i - 1) * dtdy * w.bmcmamdq(m, i);
      }
      if (m3 >= 2) {
        hadd(m, 2, 0, i) = hadd(m, 2, 0, i) + dtdx1d(i) * w.cpcqxxp(m, i);
        hadd(m, 1, 0, i) = hadd(m, 1, 0, i) + dtdx1d(i) * w.cmcqxxp(m, i);
        hadd(m, 2, 0, i - 1) = hadd(m, 2, 0, i - 1) - dtdx1d(i - 1) * w.cpcqxxm(m, i);
        hadd(m, 1, 0, i - 1) = hadd(m, 1, 0, i - 1) - dtdx1d(i - 1) * w.cmcqxxm(m, i);
      }
    }
}

// q(m,i,j,k) / aux(ma,i,j,k) with ghost cells, Fortran order
template <class R> struct Q3 {
  R* p; int n0; int nx; int ny; int mbc;
  R& operator()(int m, int i, int j, int k) const {
    return p[(((size_t)(k + mbc - 1) * ny + (j + mbc - 1)) * nx + (i + mbc - 1)) * n0 + (m - 1)];
  }
};

// qnew update of one pencil of an unsplit sweep (step3.f90:191-226 / 346-381 / 506-543, no capa).
// (ci,cj,ck) is the cell index along the pencil as a function of the pencil coordinate p:
// cell(p) = base + p*e_n;  ey / ez are the unit offsets of the y-like / z-like directions.
template <class R>
inline void step3_scatter(Q3<R> qnew, int num_eqn, int mlen, const int* base, const int* en, const int* ey,
                          const int* ez, R dtdn, R dtdy, R dtdz, Flux3Work<R>& w) {
  V2<R> qadd = w.qadd, fadd = w.fadd;
  V4<R> gadd = w.gadd, hadd = w.hadd;
  auto Q = [&](int m, int p, int a, int b) -> R& {
    return qnew(m, base[0] + p * en[0] + a * ey[0] + b * ez[0], base[1] + p * en[1] + a * ey[1] + b * ez[1],
                base[2] + p * en[2] + a * ey[2] + b * ez[2]);
  };
  // The nine forall statements act on nine different pencils of qnew, so their relative order
  // is immaterial; each statement is evaluated left to right as written.
  for (int i = 1; i <= mlen; ++i)
    for (int m = 1; m <= num_eqn; ++m) {
      Q(m, i, 0, 0) = Q(m, i, 0, 0) + qadd(m, i) - dtdn * (fadd(m, i + 1) - fadd(m, i))
                      - dtdy * (gadd(m, 2, 0, i) - gadd(m, 1, 0, i)) - dtdz * (hadd(m, 2, 0, i) - hadd(m, 1, 0, i));
      Q(m, i, -1, 0) = Q(m, i, -1, 0) - dtdy * gadd(m, 1, 0, i) - dtdz * (hadd(m, 2, -1, i) - hadd(m, 1, -1, i));
      Q(m, i, -1, -1) = Q(m, i, -1, -1) - dtdy * gadd(m, 1, -1, i) - dtdz * hadd(m, 1, -1, i);
      Q(m, i, 0, -1) = Q(m, i, 0, -1) - dtdy * (gadd(m, 2, -1, i) - gadd(m, 1, -1, i)) - dtdz * hadd(m, 1, 0, i);
      Q(m, i, 1, -1) = Q(m, i, 1, -1) + dtdy * gadd(m, 2, -1, i) - dtdz * hadd(m, 1, 1, i);
      Q(m, i, 1, 0) = Q(m, i, 1, 0) + dtdy * gadd(m, 2, 0, i) - dtdz * (hadd(m, 2, 1, i) - hadd(m, 1, 1, i));
      Q(m, i, 1, 1) = Q(m, i, 1, 1) + dtdy * gadd(m, 2, 1, i) + dtdz * hadd(m, 2, 1, i);
      Q(m, i, 0, 1) = Q(m, i, 0, 1) - dtdy * (gadd(m, 2, 1, i) - gadd(m, 1, 1, i)) + dtdz * hadd(m, 2, 0, i);
      Q(m, i, -1, 1) = Q(m, i, -1, 1) - dtdy * gadd(m, 1, 1, i) + dtdz * hadd(m, 2, -1, i);
    }
}

// The same with a capacity function (step3.f90:229-273 / 384-431 / 546-593): every bracketed flux sum is divided by
// the capacity of the cell it goes into.  C(p, a, b) = aux(index_capa, .) of the cell at pencil coordinate p, offsets
// (a, b) in the y-like / z-like directions.  dtdn is the SCALAR dt/dx of the sweep direction (qadd already carries the
// per-cell dtdx1d).  The three sweeps' statements are the same nine expressions in the rotated (a, b) frame.
template <class R>
inline void step3_scatter_capa(Q3<R> qnew, Q3<R> aux, int index_capa, int num_eqn, int mlen, const int* base,
                               const int* en, const int* ey, const int* ez, R dtdn, R dtdy, R dtdz, Flux3Work<R>& w) {
  V2<R> qadd = w.qadd, fadd = w.fadd;
  V4<R> gadd = w.gadd, hadd = w.hadd;
  auto Q = [&](int m, int p, int a, int b) -> R& {
    return qnew(m, base[0] + p * en[0] + a * ey[0] + b * ez[0], base[1] + p * en[1] + a * ey[1] + b * ez[1],
                base[2] + p * en[2] + a * ey[2] + b * ez[2]);
  };
  auto C = [&](int p, int a, int b) -> R {
    return aux(index_capa, base[0] + p * en[0] + a * ey[0] + b * ez[0], base[1] + p * en[1] + a * ey[1] + b * ez[1],
               base[2] + p * en[2] + a * ey[2] + b * ez[2]);
  };
  for (int i = 1; i <= mlen; ++i)
    for (int m = 1; m <= num_eqn; ++m) {
      Q(m, i, 0, 0) = Q(m, i, 0, 0) + qadd(m, i)
                      - (dtdn * (fadd(m, i + 1) - fadd(m, i)) + dtdy * (gadd(m, 2, 0, i) - gadd(m, 1, 0, i))
                         + dtdz * (hadd(m, 2, 0, i) - hadd(m, 1, 0, i))) / C(i, 0, 0);
      Q(m, i, -1, 0) = Q(m, i, -1, 0) - (dtdy * gadd(m, 1, 0, i) + dtdz * (hadd(m, 2, -1, i) - hadd(m, 1, -1, i))) / C(i, -1, 0);
      Q(m, i, -1, -1) = Q(m, i, -1, -1) - (dtdy * gadd(m, 1, -1, i) + dtdz * hadd(m, 1, -1, i)) / C(i, -1, -1);
      Q(m, i, 0, -1) = Q(m, i, 0, -1) - (dtdy * (gadd(m, 2, -1, i) - gadd(m, 1, -1, i)) + dtdz * hadd(m, 1, 0, i)) / C(i, 0, -1);
      Q(m, i, 1, -1) = Q(m, i, 1, -1) + (dtdy * gadd(m, 2, -1, i) - dtdz * hadd(m, 1, 1, i)) / C(i, 1, -1);
      Q(m, i, 1, 0) = Q(m, i, 1, 0) + (dtdy * gadd(m, 2, 0, i) - dtdz * (hadd(m, 2, 1, i) - hadd(m, 1, 1, i))) / C(i, 1, 0);
      Q(m, i, 1, 1) = Q(m, i, 1, 1) + (dtdy * gadd(m, 2, 1, i) + dtdz * hadd(m, 2, 1, i)) / C(i, 1, 1);
      Q(m, i, 0, 1) = Q(m, i, 0, 1) - (dtdy * (gadd(m, 2, 1, i) - gadd(m, 1, 1, i)) - dtdz * hadd(m, 2, 0, i)) / C(i, 0, 1);
      Q(m, i, -1, 1) = Q(m, i, -1, 1) - (dtdy * gadd(m, 1, 1, i) - dtdz * hadd(m, 2, -1, i)) / C(i, -1, 1);
    }
}

// =====================================================================
// step3 -- pyclaw/src/pyclaw/classic/step3.f90:2-597 (with and without a capacity function)
// =====================================================================
template <class R>
int step3(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my, int mz,
          R* qold_, R* qnew_, R* aux_, R dx, R dy, R dz, R dt, const int* method, const int* mthlim, R* cfl_out,
          bool use_fwave, int rp_id, const double* cparam, int nthreads) {
  (void)nthreads;
  Rpn3<R> rpn3; Rpt3<R> rpt3; Rptt3<R> rptt3;
  if (!select_rp3<R>(rp_id, rpn3, rpt3, rptt3)) return 1;
  if (maxm < mx || maxm < my || maxm < mz) return 2;
  const int index_capa = method[5];                      // :103
  if (index_capa < 0 || index_capa > num_aux) return 4;
  const int nx = mx + 2 * num_ghost, ny = my + 2 * num_ghost;
  Q3<R> qold{qold_, num_eqn, nx, ny, num_ghost}, qnew{qnew_, num_eqn, nx, ny, num_ghost};
  Q3<R> aux{aux_, num_aux > 0 ? num_aux : 1, nx, ny, num_ghost};
  const R dtdx = dt / dx, dtdy = dt / dy, dtdz = dt / dz;   // :106-108
  Flux3Work<R> w(maxm, num_eqn, num_waves, num_aux, num_ghost);
  for (int i = 1 - num_ghost; i <= maxm + num_ghost; ++i) { w.dtdx1d(i) = dtdx; w.dtdy1d(i) = dtdy; w.dtdz1d(i) = dtdz; }
  R cfl = R(0.0), cfl1d;

  // ---- x sweeps (:124-276) ------------------------------------------------------------------
  for (int k = 0; k <= mz + 1; ++k)
    for (int j = 0; j <= my + 1; ++j) {
      for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)
        for (int m = 1; m <= num_eqn; ++m) w.q1d(m, i) = qold(m, i, j, k);
      if (num_aux > 0)                                    // :144-150
        for (int ka = -1; ka <= 1; ++ka)
          for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i)
            for (int ma = 1; ma <= num_aux; ++ma) {
              w.aux1(ma, i, 2 + ka) = aux(ma, i, j - 1, k + ka);
              w.aux2(ma, i, 2 + ka) = aux(ma, i, j, k + ka);
              w.aux3(ma, i, 2 + ka) = aux(ma, i, j + 1, k + ka);
            }
      if (index_capa > 0)                                 // :132-136
        for (int i = 1 - num_ghost; i <= mx + num_ghost; ++i) w.dtdx1d(i) = dtdx / aux(index_capa, i, j, k);
      flux3(1, maxm, num_eqn, num_waves, num_ghost, mx, w.dtdx1d, dtdy, dtdz, num_aux, method, mthlim, w, cfl1d,
            use_fwave, rpn3, rpt3, rptt3, cparam);         // :163
      cfl = r_max(cfl, cfl1d);
      const int base[3] = {0, j, k}, en[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
      if (index_capa == 0) step3_scatter(qnew, num_eqn, mx, base, en, ey, ez, dtdx, dtdy, dtdz, w);   // :193-226
      else step3_scatter_capa(qnew, aux, index_capa, num_eqn, mx, base, en, ey, ez, dtdx, dtdy, dtdz, w);   // :229-273
    }
  // ---- y sweeps (:283-431): y-like = z, z-like = x ---------------------------------------------
  for (int k = 0; k <= mz + 1; ++k)
    for (int i = 0; i <= mx + 1; ++i) {
      for (int j = 1 - num_ghost; j <= my + num_ghost; ++j)
        for (int m = 1; m <= num_eqn; ++m) w.q1d(m, j) = qold(m, i, j, k);
      if (num_aux > 0)                                    // :300-304
        for (int ia = -1; ia <= 1; ++ia)
          for (int j = 1 - num_ghost; j <= my + num_ghost; ++j)
            for (int ma = 1; ma <= num_aux; ++ma) {
              w.aux1(ma, j, 2 + ia) = aux(ma, i + ia, j, k - 1);
              w.aux2(ma, j, 2 + ia) = aux(ma, i + ia, j, k);
              w.aux3(ma, j, 2 + ia) = aux(ma, i + ia, j, k + 1);
            }
      if (index_capa > 0)                                 // :291-295
        for (int j = 1 - num_ghost; j <= my + num_ghost; ++j) w.dtdy1d(j) = dtdy / aux(index_capa, i, j, k);
      flux3(2, maxm, num_eqn, num_waves, num_ghost, my, w.dtdy1d, dtdz, dtdx, num_aux, method, mthlim, w, cfl1d,
            use_fwave, rpn3, rpt3, rptt3, cparam);         // :317
      cfl = r_max(cfl, cfl1d);
      const int base[3] = {i, 0, k}, en[3] = {0, 1, 0}, ey[3] = {0, 0, 1}, ez[3] = {1, 0, 0};
      if (index_capa == 0) step3_scatter(qnew, num_eqn, my, base, en, ey, ez, dtdy, dtdz, dtdx, w);   // :348-381
      else step3_scatter_capa(qnew, aux, index_capa, num_eqn, my, base, en, ey, ez, dtdy, dtdz, dtdx, w);   // :384-431
    }
  // ---- z sweeps (:439-593): y-like = x, z-like = y ---------------------------------------------
  for (int j = 0; j <= my + 1; ++j)
    for (int i = 0; i <= mx + 1; ++i) {
      for (int k = 1 - num_ghost; k <= mz + num_ghost; ++k)
        for (int m = 1; m <= num_eqn; ++m) w.q1d(m, k) = qold(m, i, j, k);
      if (num_aux > 0)                                    // :458-464
        for (int ja = -1; ja <= 1; ++ja)
          for (int k = 1 - num_ghost; k <= mz + num_ghost; ++k)
            for (int ma = 1; ma <= num_aux; ++ma) {
              w.aux1(ma, k, 2 + ja) = aux(ma, i - 1, j + ja, k);
              w.aux2(ma, k, 2 + ja) = aux(ma, i, j + ja, k);
              w.aux3(ma, k, 2 + ja) = aux(ma, i + 1, j + ja, k);
            }
      if (index_capa > 0)                                 // :447-451
        for (int k = 1 - num_ghost; k <= mz + num_ghost; ++k) w.dtdz1d(k) = dtdz / aux(index_capa, i, j, k);
      flux3(3, maxm, num_eqn, num_waves, num_ghost, mz, w.dtdz1d, dtdx, dtdy, num_aux, method, mthlim, w, cfl1d,
            use_fwave, rpn3, rpt3, rptt3, cparam);         // :477
      cfl = r_max(cfl, cfl1d);
      const int base[3] = {i, j, 0}, en[3] = {0, 0, 1}, ey[3] = {1, 0, 0}, ez[3] = {0, 1, 0};
      if (index_capa == 0) step3_scatter(qnew, num_eqn, mz, base, en, ey, ez, dtdz, dtdx, dtdy, w);   // :510-543
      else step3_scatter_capa(qnew, aux, index_capa, num_eqn, mz, base, en, ey, ez, dtdz, dtdx, dtdy, w);   // :546-593
    }
  *cfl_out = cfl;
  return 0;
}

// =====================================================================
// step3ds -- pyclaw/src/pyclaw/classic/step3ds.f90 (one split sweep, idir = 1|2|3; with and without a capacity function)
// Only ONE ghost layer is swept (:134-135, 215-216, 298-299).  qold may alias qnew.
// =====================================================================
template <class R>
int step3ds(int maxm, int num_eqn, int num_waves, int num_aux, int num_ghost, int mx, int my, int mz,
            R* qold_, R* qnew_, R* aux_, R dx, R dy, R dz, R dt, const int* method, const int* mthlim, R* cfl_out,
            int idir, bool use_fwave, int rp_id, const double* cparam, int nthreads) {
  Rpn3<R> rpn3; Rpt3<R> rpt3; Rptt3<R> rptt3;
  if (!select_rp3<R>(rp_id, rpn3, rpt3, rptt3)) return 1;
  if (maxm < mx || maxm < my || maxm < mz) return 2;
  if (idir < 1 || idir > 3) return 3;
  const int index_capa = method[5];                      // :112
  if (index_capa < 0 || index_capa > num_aux) return 4;
  const int nx = mx + 2 * num_ghost, ny = my + 2 * num_ghost;
  Q3<R> qold{qold_, num_eqn, nx, ny, num_ghost}, qnew{qnew_, num_eqn, nx, ny, num_ghost};
  Q3<R> aux{aux_, num_aux > 0 ? num_aux : 1, nx, ny, num_ghost};
  const R dtdx = dt / dx, dtdy = dt / dy, dtdz = dt / dz;
  if (nthreads < 1) nthreads = 1;
  std::vector<R> cfls(nthreads, R(0.0));
  const int mm[3] = {mx, my, mz};
  // pencil direction n, outer loop coordinate o, inner loop coordinate p (step3ds.f90:134-135 etc.)
  const int n_ax = idir - 1;
  const int o_ax = (idir == 3) ? 1 : 2;                  // x,y sweeps: k outer; z sweep: j outer
  const int p_ax = (idir == 1) ? 1 : 0;                  // x sweep: j inner; y,z sweeps: i inner
  const int mlen = mm[n_ax];
  const R dtdn = (idir == 1) ? dtdx : (idir == 2 ? dtdy : dtdz);
  const R dty = (idir == 1) ? dtdy : (idir == 2 ? dtdz : dtdx);
  const R dtz = (idir == 1) ? dtdz : (idir == 2 ? dtdx : dtdy);
  // aux2(:,:,2+a) = aux at offset a in the z-like direction of the sweep (:149-151, 229-231, 312-314)
  const int z_ax = (idir == 1) ? 2 : (idir == 2 ? 0 : 1);
#pragma omp parallel num_threads(nthreads) if (nthreads > 1)
  {
    int tid = 0;
#ifdef _OPENMP
    if (nthreads > 1) tid = omp_get_thread_num();
#endif
    Flux3Work<R> w(maxm, num_eqn, num_waves, num_aux, num_ghost);
    V1<R> dt1d = w.dtdx1d;
    for (int i = 1 - num_ghost; i <= maxm + num_ghost; ++i) dt1d(i) = dtdn;
    R cfl = R(0.0), cfl1d;
#pragma omp for schedule(static) collapse(2)
    for (int o = 0; o <= mm[o_ax] + 1; ++o)
      for (int p = 0; p <= mm[p_ax] + 1; ++p) {
        int c[3];
        c[o_ax] = o; c[p_ax] = p;
        auto at = [&](Q3<R> a, int m, int t, int zoff) -> R& {
          int d[3] = {c[0], c[1], c[2]};
          d[n_ax] = t;
          d[z_ax] += zoff;
          return a(m, d[0], d[1], d[2]);
        };
        for (int t = 1 - num_ghost; t <= mlen + num_ghost; ++t)
          for (int m = 1; m <= num_eqn; ++m) w.q1d(m, t) = at(qold, m, t, 0);
        if (num_aux > 0)
          for (int a = -1; a <= 1; ++a)
            for (int t = 1 - num_ghost; t <= mlen + num_ghost; ++t)
              for (int ma = 1; ma <= num_aux; ++ma) w.aux2(ma, t, 2 + a) = at(aux, ma, t, a);
        if (index_capa > 0)                               // :142-146, 223-227, 306-310
          for (int t = 1 - num_ghost; t <= mlen + num_ghost; ++t) dt1d(t) = dtdn / at(aux, index_capa, t, 0);
        flux3(idir, maxm, num_eqn, num_waves, num_ghost, mlen, dt1d, dty, dtz, num_aux, method, mthlim, w, cfl1d,
              use_fwave, rpn3, rpt3, rptt3, cparam);
        cfl = r_max(cfl, cfl1d);
        if (index_capa == 0) {
          for (int t = 1; t <= mlen; ++t)                  // :195-198, 278-281, 361-364
            for (int m = 1; m <= num_eqn; ++m)
              at(qnew, m, t, 0) = at(qnew, m, t, 0) + w.qadd(m, t) - dtdn * (w.fadd(m, t + 1) - w.fadd(m, t));
        } else {
          for (int t = 1; t <= mlen; ++t)                  // :200-205, 282-287, 365-370: dtdx * (df) / capa, left to right
            for (int m = 1; m <= num_eqn; ++m)
              at(qnew, m, t, 0) = at(qnew, m, t, 0) + w.qadd(m, t)
                                  - dtdn * (w.fadd(m, t + 1) - w.fadd(m, t)) / at(aux, index_capa, t, 0);
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
