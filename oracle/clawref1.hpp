// ORACLE -- TEST INFRASTRUCTURE ONLY (see clawref2.hpp header).  1-D classic step.
//
// Restatement of pyclaw/src/pyclaw/classic/step1.f90:2-148 and of the 1-D Riemann solvers
//   riemann/src/rp1_advection.f90:40-50, rp1_acoustics.f90:45-80, rp1_burgers.f90:34-48,
//   rp1_shallow_roe_with_efix.f90:60-193, rp1_shallow_hlle.f90:35-87.
// Pinned by the reference's stored vectors burgers_1d `fortran_classic`, shallow_1d
// `dambreak_fortran_classic_roe / _hlle` and by the error values asserted in advection_1d/test_advection.py:30-31
// and acoustics_1d_homogeneous/test_acoustics.py:28-29 (tests/test_oracle.py).
#pragma once
#include "clawref2.hpp"

namespace clawref {

enum RpId1 {
  RP_ADVECTION_1D = 21,
  RP_ACOUSTICS_1D = 22,
  RP_BURGERS_1D = 23,
  RP_SHALLOW_ROE_EFIX_1D = 24,
  RP_SHALLOW_HLLE_1D = 25,
  RP_TRAFFIC_1D = 26,
  RP_ADVECTION_COLOR_1D = 27,
  RP_ACOUSTICS_VARIABLE_1D = 28,
  RP_SHALLOW_BATHYMETRY_FWAVE_1D = 29,
  RP_EULER_EFIX_1D = 30,
  RP_EULER_HLLE_1D = 31,
  RP_NONLINEAR_ELASTICITY_FWAVE_1D = 32,
  RP_CUBIC_1D = 33,
  RP_TRAFFIC_VC_1D = 34,
  RP_TRAFFIC_VC_FWAVE_1D = 35,
  RP_TRAFFIC_VC_TRACER_1D = 36,
};

template <class R>
using Rp1 = void (*)(int maxmx, int meqn, int mwaves, int maux, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl,
                     V2<R> auxr, V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam);

template <class R>
void rp1_advection(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave, V2<R> s,
                   V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R u = R(cparam[0]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = u;
    amdq(1, i) = r_min(u, R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(u, R(0.0)) * wave(1, 1, i);
  }
}

template <class R>
void rp1_acoustics(int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave, V2<R> s,
                   V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R cc = R(cparam[2]), zz = R(cparam[3]);           // common /cparam/ rho,bulk,cc,zz
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = ql(2, i) - qr(2, i - 1);
    R a1 = (-d1 + zz * d2) / (R(2.0) * zz);
    R a2 = (d1 + zz * d2) / (R(2.0) * zz);
    wave(1, 1, i) = -a1 * zz; wave(2, 1, i) = a1; s(1, i) = -cc;
    wave(1, 2, i) = a2 * zz;  wave(2, 2, i) = a2; s(2, i) = cc;
  }
  for (int m = 1; m <= meqn; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = s(1, i) * wave(m, 1, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i);
    }
}

template <class R>
void rp1_burgers(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave, V2<R> s,
                 V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = R(0.5) * (qr(1, i - 1) + ql(1, i));
    amdq(1, i) = r_min(s(1, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(s(1, i), R(0.0)) * wave(1, 1, i);
    if (ql(1, i) > R(0.0) && qr(1, i - 1) < R(0.0)) {      // efix = .true. (:31, :41-46)
      amdq(1, i) = -(R(0.5) * (qr(1, i - 1) * qr(1, i - 1)));   // - 1.d0/2.d0 * q**2 : the constant folds to 0.5
      apdq(1, i) = R(0.5) * (ql(1, i) * ql(1, i));
    }
  }
}

template <class R>
void rp1_shallow_roe_efix(int, int, int num_waves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                          V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R grav = R(cparam[0]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {              // loop 30
    R ubar = (qr(2, i - 1) / r_sqrt(qr(1, i - 1)) + ql(2, i) / r_sqrt(ql(1, i))) /
             (r_sqrt(qr(1, i - 1)) + r_sqrt(ql(1, i)));
    R cbar = r_sqrt(R(0.5) * grav * (qr(1, i - 1) + ql(1, i)));
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = ql(2, i) - qr(2, i - 1);
    R a1 = R(0.5) * (-d2 + (ubar + cbar) * d1) / cbar;
    R a2 = R(0.5) * (d2 - (ubar - cbar) * d1) / cbar;
    wave(1, 1, i) = a1; wave(2, 1, i) = a1 * (ubar - cbar); s(1, i) = ubar - cbar;
    wave(1, 2, i) = a2; wave(2, 2, i) = a2 * (ubar + cbar); s(2, i) = ubar + cbar;
  }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {              // efix = .true.: loop 200
    // (amdq is NOT zeroed on entry by the Fortran; every path below assigns it before use)
    R s0 = qr(2, i - 1) / qr(1, i - 1) - r_sqrt(grav * qr(1, i - 1));
    if (s0 >= R(0.0) && s(1, i) > R(0.0)) {
      for (int m = 1; m <= 2; ++m) amdq(m, i) = R(0.0);
      continue;
    }
    R hr1 = qr(1, i - 1) + wave(1, 1, i);
    R uhr1 = qr(2, i - 1) + wave(2, 1, i);
    R s1 = uhr1 / hr1 - r_sqrt(grav * hr1);
    R sfract;
    if (s0 < R(0.0) && s1 > R(0.0)) sfract = s0 * (s1 - s(1, i)) / (s1 - s0);
    else if (s(1, i) < R(0.0)) sfract = s(1, i);
    else sfract = R(0.0);
    for (int m = 1; m <= 2; ++m) amdq(m, i) = sfract * wave(m, 1, i);
    R s3 = ql(2, i) / ql(1, i) + r_sqrt(grav * ql(1, i));
    R hl2 = ql(1, i) - wave(1, 2, i);
    R uhl2 = ql(2, i) - wave(2, 2, i);
    R s2 = uhl2 / hl2 + r_sqrt(grav * hl2);
    if (s2 < R(0.0) && s3 > R(0.0)) sfract = s2 * (s3 - s(2, i)) / (s3 - s2);
    else if (s(2, i) < R(0.0)) sfract = s(2, i);
    else continue;
    for (int m = 1; m <= 2; ++m) amdq(m, i) = amdq(m, i) + sfract * wave(m, 2, i);
  }
  for (int m = 1; m <= 2; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      R df = R(0.0);
      for (int mw = 1; mw <= num_waves; ++mw) df = df + s(mw, i) * wave(m, mw, i);
      apdq(m, i) = df - amdq(m, i);
    }
}

template <class R>
void rp1_shallow_hlle(int, int meqn, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>,
                      V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R grav = R(cparam[0]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R h_l = qr(1, i - 1), h_r = ql(1, i);
    R u_l = qr(2, i - 1) / qr(1, i - 1);
    R u_r = ql(2, i) / ql(1, i);
    R c_l = r_sqrt(grav * h_l);
    R c_r = r_sqrt(grav * h_r);
    // `1.e-14` is a single-precision literal in the Fortran source
    if (h_l < R((double)1.e-14f) && h_r < R((double)1.e-14f)) {
      wave(1, 1, i) = R(0.0); wave(2, 1, i) = R(0.0); s(1, i) = R(0.0);
      wave(1, 2, i) = R(0.0); wave(2, 2, i) = R(0.0); s(2, i) = R(0.0);
    } else {
      R hsqrt_l = r_sqrt(qr(1, i - 1));
      R hsqrt_r = r_sqrt(ql(1, i));
      R h_hat = R(0.5) * (h_l + h_r);
      R u_hat = (hsqrt_l * u_l + hsqrt_r * u_r) / (hsqrt_l + hsqrt_r);
      R c_hat = r_sqrt(grav * h_hat);
      R s1 = r_min(u_l - c_l, u_hat - c_hat);
      R s2 = r_max(u_r + c_r, u_hat + c_hat);
      R h_m = (ql(2, i) - qr(2, i - 1) - s2 * h_r + s1 * h_l) / (s1 - s2);
      R hu_m = (ql(2, i) * (u_r - s2) - qr(2, i - 1) * (u_l - s1) + R(0.5) * grav * (h_r * h_r - h_l * h_l)) / (s1 - s2);
      wave(1, 1, i) = h_m - h_l; wave(2, 1, i) = hu_m - qr(2, i - 1); s(1, i) = s1;
      wave(1, 2, i) = h_r - h_m; wave(2, 2, i) = ql(2, i) - hu_m;     s(2, i) = s2;
    }
  }
  for (int m = 1; m <= meqn; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = R(0.0); apdq(m, i) = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) {
        if (s(mw, i) < R(0.0)) amdq(m, i) = amdq(m, i) + s(mw, i) * wave(m, mw, i);
        else apdq(m, i) = apdq(m, i) + s(mw, i) * wave(m, mw, i);
      }
    }
}

// traffic_1D -- riemann/src/rp1_traffic.f90:40-64     cparam = (umax); entropy fix for transonic rarefactions
template <class R>
void rp1_traffic(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave, V2<R> s,
                 V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R umax = R(cparam[0]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = umax * (R(1.0) - (qr(1, i - 1) + ql(1, i)));
    amdq(1, i) = r_min(s(1, i), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(s(1, i), R(0.0)) * wave(1, 1, i);
    R sim1 = umax * (R(1.0) - R(2.0) * qr(1, i - 1));
    R si = umax * (R(1.0) - R(2.0) * ql(1, i));
    if (sim1 < R(0.0) && si > R(0.0)) {
      R flux0 = R(0.25) * umax;
      R fluxim1 = qr(1, i - 1) * umax * (R(1.0) - qr(1, i - 1));
      R fluxi = ql(1, i) * umax * (R(1.0) - ql(1, i));
      amdq(1, i) = flux0 - fluxim1;
      apdq(1, i) = fluxi - flux0;
    }
  }
}

// advection_color_1D -- riemann/src/rp1_advection_color.f90:44-56     aux(1) = cell-centred velocity
template <class R>
void rp1_advection_color(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                         V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R u = r_min(auxr(1, i - 1), R(0.0)) + r_max(auxl(1, i), R(0.0));
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = u;
    amdq(1, i) = r_min(auxr(1, i - 1), R(0.0)) * wave(1, 1, i);
    apdq(1, i) = r_max(auxl(1, i), R(0.0)) * wave(1, 1, i);
  }
}

// acoustics_variable_1D -- riemann/src/rp1_acoustics_variable.f90:49-82     aux = (impedance Z, sound speed c)
template <class R>
void rp1_acoustics_variable(int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                            V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = ql(2, i) - qr(2, i - 1);
    R zi = auxl(1, i), zim = auxr(1, i - 1);
    R a1 = (-d1 + zi * d2) / (zim + zi);
    R a2 = (d1 + zim * d2) / (zim + zi);
    wave(1, 1, i) = -a1 * zim; wave(2, 1, i) = a1; s(1, i) = -auxr(2, i - 1);
    wave(1, 2, i) = a2 * zi;   wave(2, 2, i) = a2; s(2, i) = auxl(2, i);
  }
  for (int m = 1; m <= meqn; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = s(1, i) * wave(m, 1, i);
      apdq(m, i) = s(2, i) * wave(m, 2, i);
    }
}

// shallow_bathymetry_fwave_1D -- riemann/src/rp1_shallow_bathymetry_fwave.f90:35-118
// cparam = (grav, dry_tolerance, sea_level); aux(1) = bathymetry; f-waves
template <class R>
void rp1_shallow_bathymetry_fwave(int, int, int num_waves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl,
                                  V2<R> auxr, V3<R> fwave, V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R grav = R(cparam[0]), dry_tolerance = R(cparam[1]);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    for (int m = 1; m <= 2; ++m) { amdq(m, i) = R(0.0); apdq(m, i) = R(0.0); }    // amdq = 0.d0; apdq = 0.d0 (:52-53)
    R hl = qr(1, i - 1), ul, phil;
    if (hl > dry_tolerance) {
      ul = qr(2, i - 1) / hl;
      phil = qr(1, i - 1) * (ul * ul) + R(0.5) * grav * (qr(1, i - 1) * qr(1, i - 1));
    } else { ul = R(0.0); phil = R(0.0); }
    R bl = auxr(1, i - 1);
    R hr = ql(1, i), ur, phir;
    if (hr > dry_tolerance) {
      ur = ql(2, i) / hr;
      phir = ql(1, i) * (ur * ur) + R(0.5) * grav * (ql(1, i) * ql(1, i));
    } else { ur = R(0.0); phir = R(0.0); }
    R br = auxl(1, i);
    R hbar = R(0.5) * (hr + hl);
    R uhat = (r_sqrt(hl) * ul + r_sqrt(hr) * ur) / (r_sqrt(hl) + r_sqrt(hr));
    R chat = r_sqrt(grav * hbar);
    R fd1 = (hr * ur) - (hl * ul);
    R fd2 = phir - phil + grav * hbar * (br - bl);
    s(1, i) = r_min(ul - r_sqrt(grav * hl), uhat - chat);
    s(2, i) = r_max(ur + r_sqrt(grav * hr), uhat + chat);
    R beta1 = (s(2, i) * fd1 - fd2) / (s(2, i) - s(1, i));
    R beta2 = (fd2 - s(1, i) * fd1) / (s(2, i) - s(1, i));
    fwave(1, 1, i) = beta1 * R(1.0); fwave(2, 1, i) = beta1 * s(1, i);
    fwave(1, 2, i) = beta2 * R(1.0); fwave(2, 2, i) = beta2 * s(2, i);
    for (int k = 1; k <= num_waves; ++k) {
      if (s(k, i) < R(0.0)) {
        for (int m = 1; m <= 2; ++m) amdq(m, i) = amdq(m, i) + fwave(m, k, i);
      } else if (s(k, i) > R(0.0)) {
        for (int m = 1; m <= 2; ++m) apdq(m, i) = apdq(m, i) + fwave(m, k, i);
      } else {
        for (int m = 1; m <= 2; ++m) {
          amdq(m, i) = amdq(m, i) + R(0.5) * fwave(m, k, i);
          apdq(m, i) = apdq(m, i) + R(0.5) * fwave(m, k, i);
        }
      }
    }
  }
}

// euler_with_efix_1D -- riemann/src/rp1_euler_with_efix.f90:47-226     cparam = (gamma); Roe solver + entropy fix
template <class R>
void rp1_euler_efix(int, int, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave,
                    V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    // loop 20: Roe averages
    R rhsqrtl = r_sqrt(qr(1, i - 1));
    R rhsqrtr = r_sqrt(ql(1, i));
    R pl = gamma1 * (qr(3, i - 1) - R(0.5) * (qr(2, i - 1) * qr(2, i - 1)) / qr(1, i - 1));
    R pr = gamma1 * (ql(3, i) - R(0.5) * (ql(2, i) * ql(2, i)) / ql(1, i));
    R rhsq2 = rhsqrtl + rhsqrtr;
    R u = (qr(2, i - 1) / rhsqrtl + ql(2, i) / rhsqrtr) / rhsq2;
    R enth = ((qr(3, i - 1) + pl) / rhsqrtl + (ql(3, i) + pr) / rhsqrtr) / rhsq2;
    R a2 = gamma1 * (enth - R(0.5) * (u * u));
    R a = r_sqrt(a2);
    // loop 30: wave strengths
    R d1 = ql(1, i) - qr(1, i - 1);
    R d2 = ql(2, i) - qr(2, i - 1);
    R d3 = ql(3, i) - qr(3, i - 1);
    a2 = gamma1 / (a * a) * ((enth - u * u) * d1 + u * d2 - d3);
    R a3 = (d2 + (a - u) * d1 - a * a2) / (R(2.0) * a);
    R a1 = d1 - a2 - a3;
    wave(1, 1, i) = a1; wave(2, 1, i) = a1 * (u - a); wave(3, 1, i) = a1 * (enth - u * a); s(1, i) = u - a;
    wave(1, 2, i) = a2; wave(2, 2, i) = a2 * u;       wave(3, 2, i) = a2 * R(0.5) * (u * u); s(2, i) = u;
    wave(1, 3, i) = a3; wave(2, 3, i) = a3 * (u + a); wave(3, 3, i) = a3 * (enth + u * a); s(3, i) = u + a;
  }
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {              // efix = .true.: loop 200 (:140-213)
    R rhoim1 = qr(1, i - 1);
    R pim1 = gamma1 * (qr(3, i - 1) - R(0.5) * (qr(2, i - 1) * qr(2, i - 1)) / rhoim1);
    R cim1 = r_sqrt(gamma * pim1 / rhoim1);
    R s0 = qr(2, i - 1) / rhoim1 - cim1;
    if (s0 >= R(0.0) && s(1, i) > R(0.0)) {
      for (int m = 1; m <= 3; ++m) amdq(m, i) = R(0.0);
      continue;
    }
    R rho1 = qr(1, i - 1) + wave(1, 1, i);
    R rhou1 = qr(2, i - 1) + wave(2, 1, i);
    R en1 = qr(3, i - 1) + wave(3, 1, i);
    R p1 = gamma1 * (en1 - R(0.5) * (rhou1 * rhou1) / rho1);
    R c1 = r_sqrt(gamma * p1 / rho1);
    R s1 = rhou1 / rho1 - c1;
    R sfract;
    if (s0 < R(0.0) && s1 > R(0.0)) sfract = s0 * (s1 - s(1, i)) / (s1 - s0);
    else if (s(1, i) < R(0.0)) sfract = s(1, i);
    else sfract = R(0.0);
    for (int m = 1; m <= 3; ++m) amdq(m, i) = sfract * wave(m, 1, i);
    if (s(2, i) >= R(0.0)) continue;                       // 2-wave is rightgoing
    for (int m = 1; m <= 3; ++m) amdq(m, i) = amdq(m, i) + s(2, i) * wave(m, 2, i);
    R rhoi = ql(1, i);
    R pi = gamma1 * (ql(3, i) - R(0.5) * (ql(2, i) * ql(2, i)) / rhoi);
    R ci = r_sqrt(gamma * pi / rhoi);
    R s3 = ql(2, i) / rhoi + ci;
    R rho2 = ql(1, i) - wave(1, 3, i);
    R rhou2 = ql(2, i) - wave(2, 3, i);
    R en2 = ql(3, i) - wave(3, 3, i);
    R p2 = gamma1 * (en2 - R(0.5) * (rhou2 * rhou2) / rho2);
    R c2 = r_sqrt(gamma * p2 / rho2);
    R s2 = rhou2 / rho2 + c2;
    if (s2 < R(0.0) && s3 > R(0.0)) sfract = s2 * (s3 - s(3, i)) / (s3 - s2);
    else if (s(3, i) < R(0.0)) sfract = s(3, i);
    else continue;
    for (int m = 1; m <= 3; ++m) amdq(m, i) = amdq(m, i) + sfract * wave(m, 3, i);
  }
  for (int m = 1; m <= 3; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      R df = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) df = df + s(mw, i) * wave(m, mw, i);
      apdq(m, i) = df - amdq(m, i);
    }
}

// euler_hlle_1D -- riemann/src/rp1_euler_hlle.f90:36-112     cparam = (gamma).  `H_r` is formed with p_l, as the
// reference has it (:62).
template <class R>
void rp1_euler_hlle(int, int, int mwaves, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave,
                    V2<R> s, V2<R> amdq, V2<R> apdq, const double* cparam) {
  const R gamma = R(cparam[0]);
  const R gamma1 = gamma - R(1.0);
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R u_l = qr(2, i - 1) / qr(1, i - 1);
    R u_r = ql(2, i) / ql(1, i);
    R p_l = gamma1 * (qr(3, i - 1) - qr(1, i - 1) * (u_l * u_l) / R(2.0));
    R p_r = gamma1 * (ql(3, i) - ql(1, i) * (u_r * u_r) / R(2.0));
    R H_l = (qr(3, i - 1) + p_l) / qr(1, i - 1);
    R H_r = (ql(3, i) + p_l) / ql(1, i);
    R c_l = r_sqrt(gamma1 * H_l - (u_l * u_l) / R(2.0));
    R c_r = r_sqrt(gamma1 * H_r - (u_r * u_r) / R(2.0));
    R E_l = qr(3, i - 1), E_r = ql(3, i);
    R rhsqrt_l = r_sqrt(qr(1, i - 1));
    R rhsqrt_r = r_sqrt(ql(1, i));
    R rhsq2 = rhsqrt_l + rhsqrt_r;
    R u_hat = (qr(2, i - 1) / rhsqrt_l + ql(2, i) / rhsqrt_r) / rhsq2;
    R H_hat = ((qr(3, i - 1) + p_l) / rhsqrt_l + (ql(3, i) + p_r) / rhsqrt_r) / rhsq2;
    R c_hat = r_sqrt(gamma1 * (H_hat - R(0.5) * (u_hat * u_hat)));
    R s1 = r_min(u_l - c_l, u_hat - c_hat);
    R s2 = r_max(u_r + c_r, u_hat + c_hat);
    R rho_m = (ql(2, i) - qr(2, i - 1) - s2 * ql(1, i) + s1 * qr(1, i - 1)) / (s1 - s2);
    R rhou_m = (ql(1, i) * (u_r * u_r) - qr(1, i - 1) * (u_l * u_l) + p_r - p_l - s2 * ql(2, i) + s1 * qr(2, i - 1)) /
               (s1 - s2);
    R E_m = (u_r * (E_r + p_r) - u_l * (E_l + p_l) - s2 * E_r + s1 * E_l) / (s1 - s2);
    wave(1, 1, i) = rho_m - qr(1, i - 1); wave(2, 1, i) = rhou_m - qr(2, i - 1); wave(3, 1, i) = E_m - qr(3, i - 1);
    s(1, i) = s1;
    wave(1, 2, i) = ql(1, i) - rho_m; wave(2, 2, i) = ql(2, i) - rhou_m; wave(3, 2, i) = ql(3, i) - E_m;
    s(2, i) = s2;
  }
  for (int m = 1; m <= 3; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = R(0.0); apdq(m, i) = R(0.0);
      for (int mw = 1; mw <= mwaves; ++mw) {
        if (s(mw, i) < R(0.0)) amdq(m, i) = amdq(m, i) + s(mw, i) * wave(m, mw, i);
        else apdq(m, i) = apdq(m, i) + s(mw, i) * wave(m, mw, i);
      }
    }
}

// nonlinear_elasticity_fwave_1D -- riemann/src/rp1_nonlinear_elasticity_fwave.f90:47-128
// aux = (density, coefA, coefB): sigma = exp(coefA*eps) - 1 if coefA > 0 else coefB*eps; f-waves
template <class R> inline R ne_sigma(R eps, R coefA, R coefB) {
  if (coefA > R(0.0)) return r_exp(coefA * eps) - R(1.0);
  return coefB * eps;
}
template <class R> inline R ne_sigmap(R eps, R coefA, R coefB) {
  if (coefA > R(0.0)) return coefA * r_exp(coefA * eps);
  return coefB;
}
template <class R>
void rp1_nonlinear_elasticity_fwave(int, int meqn, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl,
                                    V2<R> auxr, V3<R> fwave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R rhoi = auxl(1, i), rhoim = auxr(1, i - 1);
    R epsi = ql(1, i), epsim = qr(1, i - 1);
    R urhoi = ql(2, i), urhoim = qr(2, i - 1);
    R bulki = ne_sigmap(epsi, auxl(2, i), auxl(3, i));
    R bulkim = ne_sigmap(epsim, auxr(2, i - 1), auxr(3, i - 1));
    R ci = r_sqrt(bulki / rhoi);
    R cim = r_sqrt(bulkim / rhoim);
    R zi = ci * rhoi, zim = cim * rhoim;
    R du = urhoi / rhoi - urhoim / rhoim;
    R dsig = ne_sigma(epsi, auxl(2, i), auxl(3, i)) - ne_sigma(epsim, auxr(2, i - 1), auxr(3, i - 1));
    R b1 = -(zi * du + dsig) / (zim + zi);
    R b2 = -(zim * du - dsig) / (zim + zi);
    fwave(1, 1, i) = b1; fwave(2, 1, i) = b1 * zim;   s(1, i) = -cim;
    fwave(1, 2, i) = b2; fwave(2, 2, i) = b2 * (-zi); s(2, i) = ci;
  }
  for (int m = 1; m <= meqn; ++m)
    for (int i = 2 - mbc; i <= mx + mbc; ++i) {
      amdq(m, i) = fwave(m, 1, i);
      apdq(m, i) = fwave(m, 2, i);
    }
}

// cubic_1D -- riemann/src/rp1_cubic.f90:25-38 (q_t + (q^3)_x = 0; all fluctuation to the right)
template <class R>
void rp1_cubic(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R>, V2<R>, V3<R> wave, V2<R> s,
               V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    wave(1, 1, i) = ql(1, i) - qr(1, i - 1);
    s(1, i) = ql(1, i) * ql(1, i) + ql(1, i) * qr(1, i - 1) + qr(1, i - 1) * qr(1, i - 1);
    amdq(1, i) = R(0.0);
    apdq(1, i) = ql(1, i) * ql(1, i) * ql(1, i) - qr(1, i - 1) * qr(1, i - 1) * qr(1, i - 1);
  }
}

// traffic_vc_1D / traffic_vc_tracer_1D -- riemann/src/rp1_traffic_vc.f90:38-84, rp1_traffic_vc_tracer.f90:40-97
// LWR traffic with a spatially varying speed limit aux(1); Godunov flux f0 of the concave flux v q (1 - q) with a
// discontinuous coefficient.  The tracer variant carries a second component advected at the car speed (wave 2).
template <class R> inline R traffic_vc_f0(R v_l, R v_r, R s_l, R s_r, R f_l, R f_r) {
  if (f_l >= R(0.25) * v_r && s_r > R(0.0)) return R(0.25) * v_r;
  if (f_r >= R(0.25) * v_l && s_l < R(0.0)) return R(0.25) * v_l;
  if (s_r <= R(0.0) && f_l > f_r) return f_r;
  if (s_l >= R(0.0) && f_r > f_l) return f_l;
  if (s_l <= R(0.0) && s_r >= R(0.0)) return (v_r <= v_l) ? R(0.25) * v_r : R(0.25) * v_l;
  if (f_l <= f_r) return f_r;
  return f_l;
}
template <class R>
void rp1_traffic_vc(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr, V3<R> wave,
                    V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R q_r = ql(1, i), q_l = qr(1, i - 1), v_r = auxl(1, i), v_l = auxr(1, i - 1);
    R s_l = v_l * (R(1.0) - R(2.0) * q_l), s_r = v_r * (R(1.0) - R(2.0) * q_r);
    R f_l = v_l * q_l * (R(1.0) - q_l), f_r = v_r * q_r * (R(1.0) - q_r);
    wave(1, 1, i) = q_r - q_l;
    s(1, i) = R(0.5) * (s_l + s_r);
    R f0 = traffic_vc_f0(v_l, v_r, s_l, s_r, f_l, f_r);
    amdq(1, i) = f0 - f_l;
    apdq(1, i) = f_r - f0;
  }
}
template <class R>
void rp1_traffic_vc_tracer(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                           V3<R> wave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R q_r = ql(1, i), q_l = qr(1, i - 1), v_r = auxl(1, i), v_l = auxr(1, i - 1);
    R s_l = v_l * (R(1.0) - R(2.0) * q_l), s_r = v_r * (R(1.0) - R(2.0) * q_r);
    s(2, i) = v_r * (R(1.0) - q_r);
    R f_l = v_l * q_l * (R(1.0) - q_l), f_r = v_r * q_r * (R(1.0) - q_r);
    wave(1, 1, i) = q_r - q_l; wave(2, 1, i) = R(0.0);
    wave(1, 2, i) = R(0.0);    wave(2, 2, i) = s(2, i) * (ql(2, i) - qr(2, i - 1));     // as the reference has it (:56)
    apdq(2, i) = wave(2, 2, i);
    amdq(2, i) = R(0.0);
    s(1, i) = R(0.5) * (s_l + s_r);
    R f0 = traffic_vc_f0(v_l, v_r, s_l, s_r, f_l, f_r);
    amdq(1, i) = f0 - f_l;
    apdq(1, i) = f_r - f0;
  }
}

// traffic_vc_fwave_1D -- riemann/src/rp1_traffic_vc_fwave.f90:38-90 (the same model as one f-wave)
template <class R>
void rp1_traffic_vc_fwave(int, int, int, int, int mbc, int mx, V2<R> ql, V2<R> qr, V2<R> auxl, V2<R> auxr,
                          V3<R> fwave, V2<R> s, V2<R> amdq, V2<R> apdq, const double*) {
  for (int i = 2 - mbc; i <= mx + mbc; ++i) {
    R sim1 = auxr(1, i - 1) * (R(1.0) - R(2.0) * qr(1, i - 1));
    R si = auxl(1, i) * (R(1.0) - R(2.0) * ql(1, i));
    R fim1 = auxr(1, i - 1) * qr(1, i - 1) * (R(1.0) - qr(1, i - 1));
    R fi = auxl(1, i) * ql(1, i) * (R(1.0) - ql(1, i));
    fwave(1, 1, i) = fi - fim1;
    if (sim1 < R(0.0) && si <= R(0.0)) {
      s(1, i) = sim1; amdq(1, i) = fwave(1, 1, i); apdq(1, i) = R(0.0);
    } else if (sim1 >= R(0.0) && si > R(0.0)) {
      s(1, i) = si; amdq(1, i) = R(0.0); apdq(1, i) = fwave(1, 1, i);
    } else if (sim1 < R(0.0) && si > R(0.0)) {
      s(1, i) = R(0.5) * (sim1 + si);
      R f0 = r_min(auxr(1, i - 1), auxl(1, i)) * R(0.25);
      amdq(1, i) = f0 - fim1; apdq(1, i) = fi - f0;
    } else {
      s(1, i) = R(0.5) * (sim1 + si);
      R d = fi - fim1;
      if (d < R(0.0)) { amdq(1, i) = fwave(1, 1, i); apdq(1, i) = R(0.0); }
      else if (d > R(0.0)) { amdq(1, i) = R(0.0); apdq(1, i) = fwave(1, 1, i); }
      else { amdq(1, i) = R(0.5) * fwave(1, 1, i); apdq(1, i) = R(0.5) * fwave(1, 1, i); }
    }
  }
}

template <class R> inline bool select_rp1(int rp_id, Rp1<R>& rp) {
  switch (rp_id) {
    case RP_ADVECTION_1D: rp = rp1_advection<R>; return true;
    case RP_ACOUSTICS_1D: rp = rp1_acoustics<R>; return true;
    case RP_BURGERS_1D: rp = rp1_burgers<R>; return true;
    case RP_SHALLOW_ROE_EFIX_1D: rp = rp1_shallow_roe_efix<R>; return true;
    case RP_SHALLOW_HLLE_1D: rp = rp1_shallow_hlle<R>; return true;
    case RP_TRAFFIC_1D: rp = rp1_traffic<R>; return true;
    case RP_ADVECTION_COLOR_1D: rp = rp1_advection_color<R>; return true;
    case RP_ACOUSTICS_VARIABLE_1D: rp = rp1_acoustics_variable<R>; return true;
    case RP_SHALLOW_BATHYMETRY_FWAVE_1D: rp = rp1_shallow_bathymetry_fwave<R>; return true;
    case RP_EULER_EFIX_1D: rp = rp1_euler_efix<R>; return true;
    case RP_EULER_HLLE_1D: rp = rp1_euler_hlle<R>; return true;
    case RP_NONLINEAR_ELASTICITY_FWAVE_1D: rp = rp1_nonlinear_elasticity_fwave<R>; return true;
    case RP_CUBIC_1D: rp = rp1_cubic<R>; return true;
    case RP_TRAFFIC_VC_1D: rp = rp1_traffic_vc<R>; return true;
    case RP_TRAFFIC_VC_FWAVE_1D: rp = rp1_traffic_vc_fwave<R>; return true;
    case RP_TRAFFIC_VC_TRACER_1D: rp = rp1_traffic_vc_tracer<R>; return true;
  }
  return false;
}

// =====================================================================
// step1 -- pyclaw/src/pyclaw/classic/step1.f90:2-148.  q(num_eqn, 1-num_ghost:mx+num_ghost) is updated IN PLACE.
// =====================================================================
template <class R>
int step1(int num_eqn, int num_waves, int num_ghost, int num_aux, int mx, R* q_, R* aux_, R dx, R dt,
          const int* method, const int* mthlim, R* cfl_out, bool use_fwave, int rp_id, const double* cparam) {
  Rp1<R> rp1;
  if (!select_rp1<R>(rp_id, rp1)) return 1;
  const int lo = 1 - num_ghost, n = mx + 2 * num_ghost;
  const int na = num_aux > 0 ? num_aux : 1;
  V2<R> q{q_, num_eqn, lo}, aux{aux_, na, lo};
  std::vector<R> f_(n * num_eqn), s_(n * num_waves), wave_((size_t)n * num_eqn * num_waves), amdq_(n * num_eqn),
      apdq_(n * num_eqn), dtdx_(n);
  V2<R> f{f_.data(), num_eqn, lo}, s{s_.data(), num_waves, lo}, amdq{amdq_.data(), num_eqn, lo},
      apdq{apdq_.data(), num_eqn, lo};
  V3<R> wave{wave_.data(), num_eqn, num_waves, lo};
  V1<R> dtdx{dtdx_.data(), lo};
  bool limit = false;                                                   // :64-67
  for (int mw = 0; mw < num_waves; ++mw) if (mthlim[mw] > 0) limit = true;
  const int index_capa = method[5];                                     // :69-80
  for (int i = lo; i <= mx + num_ghost; ++i) {
    if (index_capa > 0) {
      if (aux(index_capa, i) <= R(0.0)) return 3;                       // 'Error -- capa must be positive'
      dtdx(i) = dt / (dx * aux(index_capa, i));
    } else {
      dtdx(i) = dt / dx;
    }
  }
  rp1(mx, num_eqn, num_waves, num_aux, num_ghost, mx, q, q, aux, aux, wave, s, amdq, apdq, cparam);   // :86-87
  // forall with two statements: each statement completes for ALL (i, m) before the next starts (:93-96)
  for (int i = 1; i <= mx + 1; ++i)
    for (int m = 1; m <= num_eqn; ++m) q(m, i) = q(m, i) - dtdx(i) * apdq(m, i);
  for (int i = 1; i <= mx + 1; ++i)
    for (int m = 1; m <= num_eqn; ++m) q(m, i - 1) = q(m, i - 1) - dtdx(i - 1) * amdq(m, i);
  R cfl = R(0.0);                                                       // :100-105
  for (int mw = 1; mw <= num_waves; ++mw)
    for (int i = 1; i <= mx + 1; ++i) cfl = r_max3(cfl, dtdx(i) * s(mw, i), -dtdx(i - 1) * s(mw, i));
  *cfl_out = cfl;
  if (method[1] == 1) return 0;                                         // :107
  for (int i = lo; i <= mx + num_ghost; ++i)
    for (int m = 1; m <= num_eqn; ++m) f(m, i) = R(0.0);                // :111-113
  if (limit) limiter<R>(mx, num_eqn, num_waves, num_ghost, mx, wave, s, mthlim);   // :116-117
  for (int i = 1; i <= mx + 1; ++i)                                     // :119-141
    for (int m = 1; m <= num_eqn; ++m)
      for (int mw = 1; mw <= num_waves; ++mw) {
        R dtdxave = R(0.5) * (dtdx(i - 1) + dtdx(i));
        R lead = use_fwave ? r_sign1(s(mw, i)) : r_abs(s(mw, i));
        f(m, i) = f(m, i) + R(0.5) * lead * (R(1.0) - r_abs(s(mw, i)) * dtdxave) * wave(m, mw, i);
      }
  for (int i = 1; i <= mx + 1; ++i)                                     // :145-147
    for (int m = 1; m <= num_eqn; ++m) q(m, i) = q(m, i) - dtdx(i) * (f(m, i + 1) - f(m, i));
  return 0;
}

}  // namespace clawref
