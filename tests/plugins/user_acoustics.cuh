// A "user-written" Riemann solver for the plug-in mechanism (clawpack_b200/csrc/rp_plugin.cu): constant-coefficient
// acoustics, written the way a user of the reference would port riemann/src/rpn2_acoustics.f90:79-106 and
// rpt2_acoustics.f90:38-54 to a device functor.  The parity tests run it against the oracle's acoustics_2D.
//   cparam = (rho, bulk, cc, zz) as in COMMON /cparam/ of the Fortran file
//   rotated frame: component 1 is the velocity normal to the interface, component 2 the transverse one
#pragma once

namespace clawcuda {
namespace CLAW_NS {

struct UserAcoustics2D {
  static constexpr int MEQN = 3, MWAVES = 2;
  static constexpr int NROE = 1;        // doubles handed from normal() to transverse() per interface (unused here)
  static constexpr int NCELL = 0;       // no per-cell precomputation
  static constexpr int S_HI = 1;        // index of the fastest wave
  static constexpr int WARPS_PER_SM = 16;

  __device__ __forceinline__ static void cell(const RpParams&, const double*, double*) {}

  // rpn2: waves, speeds and fluctuations of the interface between ql (left cell) and qr (right cell)
  __device__ __forceinline__ static void normal(const RpParams& P, int /*dir*/, const double* ql, const double* qr,
                                                const double*, const double*, double (*wave)[MEQN], double* s,
                                                double* amdq, double* apdq, double* roe) {
    const double cc = P.p[2], zz = P.p[3];
    const double dp = qr[0] - ql[0], du = qr[1] - ql[1];
    const double a1 = ddiv(-dp + zz * du, 2.0 * zz);
    const double a2 = ddiv(dp + zz * du, 2.0 * zz);
    wave[0][0] = -a1 * zz; wave[0][1] = a1; wave[0][2] = 0.0;
    wave[1][0] = a2 * zz;  wave[1][1] = a2; wave[1][2] = 0.0;
    s[0] = -cc; s[1] = cc;
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
    roe[0] = 0.0;
  }

  // wave speeds of an interface from what normal() left in roe[] (used for the CFL number of re-visited interfaces)
  __device__ __forceinline__ static void speeds(const RpParams& P, int, const double*, double* s) {
    s[0] = -P.p[2]; s[1] = P.p[2];
  }

  // rpt2: split the fluctuation asdq into down-going (bm) and up-going (bp) parts
  __device__ __forceinline__ static void transverse(const RpParams& P, int, const double*, const double* asdq, double* bm,
                                                    double* bp) {
    const double cc = P.p[2], zz = P.p[3];
    const double a1 = ddiv(-asdq[0] + zz * asdq[2], 2.0 * zz);
    const double a2 = ddiv(asdq[0] + zz * asdq[2], 2.0 * zz);
    bm[0] = cc * a1 * zz; bm[1] = 0.0; bm[2] = -cc * a1;
    bp[0] = cc * a2 * zz; bp[1] = 0.0; bp[2] = cc * a2;
  }
};

}  // namespace CLAW_NS
}  // namespace clawcuda
