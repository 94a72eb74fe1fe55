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

}  // namespace CLAW_NS
}  // namespace clawcuda
