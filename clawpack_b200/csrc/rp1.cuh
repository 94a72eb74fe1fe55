// Point-wise 1-D Riemann solvers (device functors for sweep1.cuh), restating riemann/src/rp1_*.f90.
//   RP::normal(P, ql, qr, wave, s, amdq, apdq)     ql = left cell (Fortran qr(:,i-1)), qr = right cell (ql(:,i))
#pragma once
#include "rp2.cuh"

namespace clawcuda {
namespace CLAW_NS {

// advection_1D -- rp1_advection.f90:40-50     cparam = (u)
struct RpAdvection1D {
  static constexpr int MEQN = 1, MWAVES = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double u = P.p[0];
    wave[0][0] = qr[0] - ql[0];
    s[0] = u;
    amdq[0] = fmin2(u, 0.0) * wave[0][0];
    apdq[0] = fmax2(u, 0.0) * wave[0][0];
  }
};

// acoustics_1D -- rp1_acoustics.f90:45-80     cparam = (rho, bulk, cc, zz)
struct RpAcoustics1D {
  static constexpr int MEQN = 2, MWAVES = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double cc = P.p[2], zz = P.p[3];
    const double d1 = qr[0] - ql[0], d2 = qr[1] - ql[1];
    const double a1 = ddiv(-d1 + zz * d2, 2.0 * zz);
    const double a2 = ddiv(d1 + zz * d2, 2.0 * zz);
    wave[0][0] = -a1 * zz; wave[0][1] = a1; s[0] = -cc;
    wave[1][0] = a2 * zz;  wave[1][1] = a2; s[1] = cc;
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
  }
};

// burgers_1D -- rp1_burgers.f90:34-48 (efix = .true.)
struct RpBurgers1D {
  static constexpr int MEQN = 1, MWAVES = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    wave[0][0] = qr[0] - ql[0];
    s[0] = 0.5 * (ql[0] + qr[0]);
    amdq[0] = fmin2(s[0], 0.0) * wave[0][0];
    apdq[0] = fmax2(s[0], 0.0) * wave[0][0];
    if (qr[0] > 0.0 && ql[0] < 0.0) {                      // transonic rarefaction (:41-46)
      amdq[0] = -(0.5 * (ql[0] * ql[0]));
      apdq[0] = 0.5 * (qr[0] * qr[0]);
    }
  }
};

// shallow_roe_with_efix_1D -- rp1_shallow_roe_with_efix.f90:60-193     cparam = (grav)
struct RpShallowRoeEfix1D {
  static constexpr int MEQN = 2, MWAVES = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double grav = P.p[0];
    const double sl = sqrt(ql[0]), sr = sqrt(qr[0]);
    const double ubar = ddiv(ddiv(ql[1], sl) + ddiv(qr[1], sr), sl + sr);
    const double cbar = sqrt(0.5 * grav * (ql[0] + qr[0]));
    const double d1 = qr[0] - ql[0], d2 = qr[1] - ql[1];
    const double a1 = ddiv(0.5 * (-d2 + (ubar + cbar) * d1), cbar);
    const double a2 = ddiv(0.5 * (d2 - (ubar - cbar) * d1), cbar);
    wave[0][0] = a1; wave[0][1] = a1 * (ubar - cbar); s[0] = ubar - cbar;
    wave[1][0] = a2; wave[1][1] = a2 * (ubar + cbar); s[1] = ubar + cbar;
    // entropy fix, loop 200
    const double s0 = ddiv(ql[1], ql[0]) - sqrt(grav * ql[0]);
    if (s0 >= 0.0 && s[0] > 0.0) {
      amdq[0] = 0.0; amdq[1] = 0.0;
    } else {
      const double hr1 = ql[0] + wave[0][0], uhr1 = ql[1] + wave[0][1];
      const double s1 = ddiv(uhr1, hr1) - sqrt(grav * hr1);
      double sfract;
      if (s0 < 0.0 && s1 > 0.0) sfract = ddiv(s0 * (s1 - s[0]), s1 - s0);
      else if (s[0] < 0.0) sfract = s[0];
      else sfract = 0.0;
      amdq[0] = sfract * wave[0][0]; amdq[1] = sfract * wave[0][1];
      const double s3 = ddiv(qr[1], qr[0]) + sqrt(grav * qr[0]);
      const double hl2 = qr[0] - wave[1][0], uhl2 = qr[1] - wave[1][1];
      const double s2 = ddiv(uhl2, hl2) + sqrt(grav * hl2);
      bool add2 = true;
      if (s2 < 0.0 && s3 > 0.0) sfract = ddiv(s2 * (s3 - s[1]), s3 - s2);
      else if (s[1] < 0.0) sfract = s[1];
      else add2 = false;
      if (add2) { amdq[0] = amdq[0] + sfract * wave[1][0]; amdq[1] = amdq[1] + sfract * wave[1][1]; }
    }
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      double df = from0(s[0] * wave[0][m]);
      df = df + s[1] * wave[1][m];
      apdq[m] = df - amdq[m];
    }
  }
};

// shallow_hlle_1D -- rp1_shallow_hlle.f90:35-87     cparam = (grav)
struct RpShallowHlle1D {
  static constexpr int MEQN = 2, MWAVES = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double grav = P.p[0];
    const double h_l = ql[0], h_r = qr[0];
    const double tiny = (double)1.e-14f;                     // single-precision literal in the Fortran source
    if (h_l < tiny && h_r < tiny) {
      wave[0][0] = 0.0; wave[0][1] = 0.0; s[0] = 0.0;
      wave[1][0] = 0.0; wave[1][1] = 0.0; s[1] = 0.0;
    } else {
      const double u_l = ddiv(ql[1], ql[0]), u_r = ddiv(qr[1], qr[0]);
      const double c_l = sqrt(grav * h_l), c_r = sqrt(grav * h_r);
      const double hsqrt_l = sqrt(h_l), hsqrt_r = sqrt(h_r);
      const double h_hat = 0.5 * (h_l + h_r);
      const double u_hat = ddiv(hsqrt_l * u_l + hsqrt_r * u_r, hsqrt_l + hsqrt_r);
      const double c_hat = sqrt(grav * h_hat);
      const double s1 = fmin2(u_l - c_l, u_hat - c_hat);
      const double s2 = fmax2(u_r + c_r, u_hat + c_hat);
      const double h_m = ddiv(qr[1] - ql[1] - s2 * h_r + s1 * h_l, s1 - s2);
      const double hu_m = ddiv(qr[1] * (u_r - s2) - ql[1] * (u_l - s1) + 0.5 * grav * (h_r * h_r - h_l * h_l), s1 - s2);
      wave[0][0] = h_m - h_l; wave[0][1] = hu_m - ql[1]; s[0] = s1;
      wave[1][0] = h_r - h_m; wave[1][1] = qr[1] - hu_m; s[1] = s2;
    }
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      double am = 0.0, ap = 0.0;
#pragma unroll
      for (int mw = 0; mw < MWAVES; ++mw) {
        if (s[mw] < 0.0) am = am + s[mw] * wave[mw][m];
        else ap = ap + s[mw] * wave[mw][m];
      }
      amdq[m] = am; apdq[m] = ap;
    }
  }
};

// =============================================================================================================
// Solvers below may read aux: functors that declare MAUX > 0 get the aux columns of the left and right cell,
//   RP::normal(P, ql, qr, al, ar, wave, s, amdq, apdq)      al = Fortran auxr(:,i-1), ar = auxl(:,i)
// For f-wave solvers (`fwave` in riemann/static.py) `wave` holds f-waves, as in the reference.
// =============================================================================================================

// traffic_1D -- rp1_traffic.f90:40-64     cparam = (umax)
struct RpTraffic1D {
  static constexpr int MEQN = 1, MWAVES = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double umax = P.p[0];
    wave[0][0] = qr[0] - ql[0];
    s[0] = umax * (1.0 - (ql[0] + qr[0]));
    amdq[0] = fmin2(s[0], 0.0) * wave[0][0];
    apdq[0] = fmax2(s[0], 0.0) * wave[0][0];
    const double sim1 = umax * (1.0 - 2.0 * ql[0]);
    const double si = umax * (1.0 - 2.0 * qr[0]);
    if (sim1 < 0.0 && si > 0.0) {                          // transonic rarefaction (:54-61)
      const double flux0 = 0.25 * umax;
      const double fluxim1 = ql[0] * umax * (1.0 - ql[0]);
      const double fluxi = qr[0] * umax * (1.0 - qr[0]);
      amdq[0] = flux0 - fluxim1;
      apdq[0] = fluxi - flux0;
    }
  }
};

// advection_color_1D -- rp1_advection_color.f90:44-56     aux(1) = cell-centred velocity
struct RpAdvectionColor1D {
  static constexpr int MEQN = 1, MWAVES = 1, MAUX = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double um = fmin2(al[0], 0.0), up = fmax2(ar[0], 0.0);
    wave[0][0] = qr[0] - ql[0];
    s[0] = um + up;
    amdq[0] = um * wave[0][0];
    apdq[0] = up * wave[0][0];
  }
};

// acoustics_variable_1D -- rp1_acoustics_variable.f90:49-82     aux = (impedance, sound speed)
struct RpAcousticsVariable1D {
  static constexpr int MEQN = 2, MWAVES = 2, MAUX = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double d1 = qr[0] - ql[0], d2 = qr[1] - ql[1];
    const double zi = ar[0], zim = al[0];
    const double a1 = ddiv(-d1 + zi * d2, zim + zi);
    const double a2 = ddiv(d1 + zim * d2, zim + zi);
    wave[0][0] = -a1 * zim; wave[0][1] = a1; s[0] = -al[1];
    wave[1][0] = a2 * zi;   wave[1][1] = a2; s[1] = ar[1];
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
  }
};

// shallow_bathymetry_fwave_1D -- rp1_shallow_bathymetry_fwave.f90:35-118
// cparam = (grav, dry_tolerance, sea_level); aux(1) = bathymetry; f-waves
struct RpShallowBathymetryFwave1D {
  static constexpr int MEQN = 2, MWAVES = 2, MAUX = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*fwave)[MEQN], double* s, double* amdq, double* apdq) {
    const double grav = P.p[0], dry_tolerance = P.p[1];
    const double hl = ql[0], hr = qr[0];
    double ul = 0.0, phil = 0.0, ur = 0.0, phir = 0.0;
    if (hl > dry_tolerance) {
      ul = ddiv(ql[1], hl);
      phil = ql[0] * (ul * ul) + 0.5 * grav * (ql[0] * ql[0]);
    }
    if (hr > dry_tolerance) {
      ur = ddiv(qr[1], hr);
      phir = qr[0] * (ur * ur) + 0.5 * grav * (qr[0] * qr[0]);
    }
    const double bl = al[0], br = ar[0];
    const double hbar = 0.5 * (hr + hl);
    const double sql = sqrt(hl), sqr = sqrt(hr);
    const double uhat = ddiv(sql * ul + sqr * ur, sql + sqr);
    const double chat = sqrt(grav * hbar);
    const double fd1 = (hr * ur) - (hl * ul);
    const double fd2 = phir - phil + grav * hbar * (br - bl);
    s[0] = fmin2(ul - sqrt(grav * hl), uhat - chat);
    s[1] = fmax2(ur + sqrt(grav * hr), uhat + chat);
    const double beta1 = ddiv(s[1] * fd1 - fd2, s[1] - s[0]);
    const double beta2 = ddiv(fd2 - s[0] * fd1, s[1] - s[0]);
    fwave[0][0] = beta1; fwave[0][1] = beta1 * s[0];
    fwave[1][0] = beta2; fwave[1][1] = beta2 * s[1];
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = 0.0; apdq[m] = 0.0; }
#pragma unroll
    for (int k = 0; k < MWAVES; ++k) {
#pragma unroll
      for (int m = 0; m < MEQN; ++m) {
        if (s[k] < 0.0) amdq[m] = amdq[m] + fwave[k][m];
        else if (s[k] > 0.0) apdq[m] = apdq[m] + fwave[k][m];
        else {
          amdq[m] = amdq[m] + 0.5 * fwave[k][m];
          apdq[m] = apdq[m] + 0.5 * fwave[k][m];
        }
      }
    }
  }
};

// euler_with_efix_1D -- rp1_euler_with_efix.f90:47-226     cparam = (gamma); Roe solver + entropy fix
struct RpEulerEfix1D {
  static constexpr int MEQN = 3, MWAVES = 3;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double gamma = P.p[0];
    const double gamma1 = gamma - 1.0;
    const double rhsqrtl = sqrt(ql[0]), rhsqrtr = sqrt(qr[0]);
    const double pl = gamma1 * (ql[2] - ddiv(0.5 * (ql[1] * ql[1]), ql[0]));
    const double pr = gamma1 * (qr[2] - ddiv(0.5 * (qr[1] * qr[1]), qr[0]));
    const double rhsq2 = rhsqrtl + rhsqrtr;
    const double u = ddiv(ddiv(ql[1], rhsqrtl) + ddiv(qr[1], rhsqrtr), rhsq2);
    const double enth = ddiv(ddiv(ql[2] + pl, rhsqrtl) + ddiv(qr[2] + pr, rhsqrtr), rhsq2);
    const double a = sqrt(gamma1 * (enth - 0.5 * (u * u)));
    const double d1 = qr[0] - ql[0], d2 = qr[1] - ql[1], d3 = qr[2] - ql[2];
    const double a2 = ddiv(gamma1, a * a) * ((enth - u * u) * d1 + u * d2 - d3);
    const double a3 = ddiv(d2 + (a - u) * d1 - a * a2, 2.0 * a);
    const double a1 = d1 - a2 - a3;
    wave[0][0] = a1; wave[0][1] = a1 * (u - a); wave[0][2] = a1 * (enth - u * a); s[0] = u - a;
    wave[1][0] = a2; wave[1][1] = a2 * u;       wave[1][2] = a2 * 0.5 * (u * u); s[1] = u;
    wave[2][0] = a3; wave[2][1] = a3 * (u + a); wave[2][2] = a3 * (enth + u * a); s[2] = u + a;
    // entropy fix, loop 200 (:140-213).  pim1 is the same expression as pl (and pi as pr): same bits.
    const double cim1 = sqrt(ddiv(gamma * pl, ql[0]));
    const double s0 = ddiv(ql[1], ql[0]) - cim1;
    if (s0 >= 0.0 && s[0] > 0.0) {
#pragma unroll
      for (int m = 0; m < MEQN; ++m) amdq[m] = 0.0;
    } else {
      const double rho1 = ql[0] + wave[0][0], rhou1 = ql[1] + wave[0][1], en1 = ql[2] + wave[0][2];
      const double p1 = gamma1 * (en1 - ddiv(0.5 * (rhou1 * rhou1), rho1));
      const double c1 = sqrt(ddiv(gamma * p1, rho1));
      const double s1 = ddiv(rhou1, rho1) - c1;
      double sfract;
      if (s0 < 0.0 && s1 > 0.0) sfract = ddiv(s0 * (s1 - s[0]), s1 - s0);
      else if (s[0] < 0.0) sfract = s[0];
      else sfract = 0.0;
#pragma unroll
      for (int m = 0; m < MEQN; ++m) amdq[m] = sfract * wave[0][m];
      if (!(s[1] >= 0.0)) {                                // 2-wave is leftgoing
#pragma unroll
        for (int m = 0; m < MEQN; ++m) amdq[m] = amdq[m] + s[1] * wave[1][m];
        const double ci = sqrt(ddiv(gamma * pr, qr[0]));
        const double s3 = ddiv(qr[1], qr[0]) + ci;
        const double rho2 = qr[0] - wave[2][0], rhou2 = qr[1] - wave[2][1], en2 = qr[2] - wave[2][2];
        const double p2 = gamma1 * (en2 - ddiv(0.5 * (rhou2 * rhou2), rho2));
        const double c2 = sqrt(ddiv(gamma * p2, rho2));
        const double s2 = ddiv(rhou2, rho2) + c2;
        bool add3 = true;
        if (s2 < 0.0 && s3 > 0.0) sfract = ddiv(s2 * (s3 - s[2]), s3 - s2);
        else if (s[2] < 0.0) sfract = s[2];
        else add3 = false;
        if (add3) {
#pragma unroll
          for (int m = 0; m < MEQN; ++m) amdq[m] = amdq[m] + sfract * wave[2][m];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      double df = from0(s[0] * wave[0][m]);
      df = df + s[1] * wave[1][m];
      df = df + s[2] * wave[2][m];
      apdq[m] = df - amdq[m];
    }
  }
};

// euler_hlle_1D -- rp1_euler_hlle.f90:36-112     cparam = (gamma).  H_r is formed with p_l, as the reference has it (:62)
struct RpEulerHlle1D {
  static constexpr int MEQN = 3, MWAVES = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double gamma1 = P.p[0] - 1.0;
    const double u_l = ddiv(ql[1], ql[0]), u_r = ddiv(qr[1], qr[0]);
    const double p_l = gamma1 * (ql[2] - ql[0] * (u_l * u_l) * 0.5);       // x/2.d0 == x*0.5 bit for bit
    const double p_r = gamma1 * (qr[2] - qr[0] * (u_r * u_r) * 0.5);
    const double H_l = ddiv(ql[2] + p_l, ql[0]);
    const double H_r = ddiv(qr[2] + p_l, qr[0]);
    const double c_l = sqrt(gamma1 * H_l - (u_l * u_l) * 0.5);
    const double c_r = sqrt(gamma1 * H_r - (u_r * u_r) * 0.5);
    const double E_l = ql[2], E_r = qr[2];
    const double rhsqrt_l = sqrt(ql[0]), rhsqrt_r = sqrt(qr[0]);
    const double rhsq2 = rhsqrt_l + rhsqrt_r;
    const double u_hat = ddiv(ddiv(ql[1], rhsqrt_l) + ddiv(qr[1], rhsqrt_r), rhsq2);
    const double H_hat = ddiv(ddiv(ql[2] + p_l, rhsqrt_l) + ddiv(qr[2] + p_r, rhsqrt_r), rhsq2);
    const double c_hat = sqrt(gamma1 * (H_hat - 0.5 * (u_hat * u_hat)));
    const double s1 = fmin2(u_l - c_l, u_hat - c_hat);
    const double s2 = fmax2(u_r + c_r, u_hat + c_hat);
    const double rho_m = ddiv(qr[1] - ql[1] - s2 * qr[0] + s1 * ql[0], s1 - s2);
    const double rhou_m = ddiv(qr[0] * (u_r * u_r) - ql[0] * (u_l * u_l) + p_r - p_l - s2 * qr[1] + s1 * ql[1], s1 - s2);
    const double E_m = ddiv(u_r * (E_r + p_r) - u_l * (E_l + p_l) - s2 * E_r + s1 * E_l, s1 - s2);
    wave[0][0] = rho_m - ql[0]; wave[0][1] = rhou_m - ql[1]; wave[0][2] = E_m - ql[2]; s[0] = s1;
    wave[1][0] = qr[0] - rho_m; wave[1][1] = qr[1] - rhou_m; wave[1][2] = qr[2] - E_m; s[1] = s2;
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      double am = 0.0, ap = 0.0;
#pragma unroll
      for (int mw = 0; mw < MWAVES; ++mw) {
        if (s[mw] < 0.0) am = am + s[mw] * wave[mw][m];
        else ap = ap + s[mw] * wave[mw][m];
      }
      amdq[m] = am; apdq[m] = ap;
    }
  }
};

// nonlinear_elasticity_fwave_1D -- rp1_nonlinear_elasticity_fwave.f90:47-128     aux = (density, coefA, coefB);
// sigma = exp(coefA*eps) - 1 if coefA > 0 else coefB*eps; f-waves.  exp() is the one library call on the 1-D path:
// its last bit may differ from the host library's, so only the linear relation is bit-reproducible.
struct RpNonlinearElasticityFwave1D {
  static constexpr int MEQN = 2, MWAVES = 2, MAUX = 3;
  __device__ __forceinline__ static void stress(double eps, double coefA, double coefB, double& sigma, double& sigmap) {
    if (coefA > 0.0) {
      const double e = exp(coefA * eps);
      sigma = e - 1.0;
      sigmap = coefA * e;
    } else {
      sigma = coefB * eps;
      sigmap = coefB;
    }
  }
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*fwave)[MEQN], double* s, double* amdq, double* apdq) {
    const double rhoi = ar[0], rhoim = al[0];
    double sigi, bulki, sigim, bulkim;
    stress(qr[0], ar[1], ar[2], sigi, bulki);
    stress(ql[0], al[1], al[2], sigim, bulkim);
    const double ci = sqrt(ddiv(bulki, rhoi)), cim = sqrt(ddiv(bulkim, rhoim));
    const double zi = ci * rhoi, zim = cim * rhoim;
    const double du = ddiv(qr[1], rhoi) - ddiv(ql[1], rhoim);
    const double dsig = sigi - sigim;
    const double b1 = ddiv(-(zi * du + dsig), zim + zi);
    const double b2 = ddiv(-(zim * du - dsig), zim + zi);
    fwave[0][0] = b1; fwave[0][1] = b1 * zim;   s[0] = -cim;
    fwave[1][0] = b2; fwave[1][1] = b2 * (-zi); s[1] = ci;
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = fwave[0][m]; apdq[m] = fwave[1][m]; }
  }
};

// cubic_1D -- rp1_cubic.f90:25-38 (q_t + (q^3)_x = 0; the whole flux difference goes right)
struct RpCubic1D {
  static constexpr int MEQN = 1, MWAVES = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    wave[0][0] = qr[0] - ql[0];
    s[0] = qr[0] * qr[0] + qr[0] * ql[0] + ql[0] * ql[0];
    amdq[0] = 0.0;
    apdq[0] = qr[0] * qr[0] * qr[0] - ql[0] * ql[0] * ql[0];
  }
};

// traffic_vc_1D / traffic_vc_tracer_1D -- rp1_traffic_vc.f90:38-84, rp1_traffic_vc_tracer.f90:40-97     aux(1) = speed limit
__device__ __forceinline__ double traffic_vc_f0(double v_l, double v_r, double s_l, double s_r, double f_l, double f_r) {
  if (f_l >= 0.25 * v_r && s_r > 0.0) return 0.25 * v_r;
  if (f_r >= 0.25 * v_l && s_l < 0.0) return 0.25 * v_l;
  if (s_r <= 0.0 && f_l > f_r) return f_r;
  if (s_l >= 0.0 && f_r > f_l) return f_l;
  if (s_l <= 0.0 && s_r >= 0.0) return (v_r <= v_l) ? 0.25 * v_r : 0.25 * v_l;
  if (f_l <= f_r) return f_r;
  return f_l;
}
struct RpTrafficVc1D {
  static constexpr int MEQN = 1, MWAVES = 1, MAUX = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double q_l = ql[0], q_r = qr[0], v_l = al[0], v_r = ar[0];
    const double s_l = v_l * (1.0 - 2.0 * q_l), s_r = v_r * (1.0 - 2.0 * q_r);
    const double f_l = v_l * q_l * (1.0 - q_l), f_r = v_r * q_r * (1.0 - q_r);
    wave[0][0] = q_r - q_l;
    s[0] = 0.5 * (s_l + s_r);
    const double f0 = traffic_vc_f0(v_l, v_r, s_l, s_r, f_l, f_r);
    amdq[0] = f0 - f_l;
    apdq[0] = f_r - f0;
  }
};
struct RpTrafficVcTracer1D {
  static constexpr int MEQN = 2, MWAVES = 2, MAUX = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*wave)[MEQN], double* s, double* amdq, double* apdq) {
    const double q_l = ql[0], q_r = qr[0], v_l = al[0], v_r = ar[0];
    const double s_l = v_l * (1.0 - 2.0 * q_l), s_r = v_r * (1.0 - 2.0 * q_r);
    s[1] = v_r * (1.0 - q_r);
    const double f_l = v_l * q_l * (1.0 - q_l), f_r = v_r * q_r * (1.0 - q_r);
    wave[0][0] = q_r - q_l; wave[0][1] = 0.0;
    wave[1][0] = 0.0;       wave[1][1] = s[1] * (qr[1] - ql[1]);
    apdq[1] = wave[1][1];
    amdq[1] = 0.0;
    s[0] = 0.5 * (s_l + s_r);
    const double f0 = traffic_vc_f0(v_l, v_r, s_l, s_r, f_l, f_r);
    amdq[0] = f0 - f_l;
    apdq[0] = f_r - f0;
  }
};

// traffic_vc_fwave_1D -- rp1_traffic_vc_fwave.f90:38-90 (one f-wave)
struct RpTrafficVcFwave1D {
  static constexpr int MEQN = 1, MWAVES = 1, MAUX = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double* al, const double* ar,
                                                double (*fwave)[MEQN], double* s, double* amdq, double* apdq) {
    const double sim1 = al[0] * (1.0 - 2.0 * ql[0]);
    const double si = ar[0] * (1.0 - 2.0 * qr[0]);
    const double fim1 = al[0] * ql[0] * (1.0 - ql[0]);
    const double fi = ar[0] * qr[0] * (1.0 - qr[0]);
    const double fw = fi - fim1;
    fwave[0][0] = fw;
    if (sim1 < 0.0 && si <= 0.0) {
      s[0] = sim1; amdq[0] = fw; apdq[0] = 0.0;
    } else if (sim1 >= 0.0 && si > 0.0) {
      s[0] = si; amdq[0] = 0.0; apdq[0] = fw;
    } else if (sim1 < 0.0 && si > 0.0) {
      s[0] = 0.5 * (sim1 + si);
      const double f0 = fmin2(al[0], ar[0]) * 0.25;
      amdq[0] = f0 - fim1; apdq[0] = fi - f0;
    } else {
      s[0] = 0.5 * (sim1 + si);
      if (fw < 0.0) { amdq[0] = fw; apdq[0] = 0.0; }
      else if (fw > 0.0) { amdq[0] = 0.0; apdq[0] = fw; }
      else { amdq[0] = 0.5 * fw; apdq[0] = 0.5 * fw; }
    }
  }
};

}  // namespace CLAW_NS
}  // namespace clawcuda
