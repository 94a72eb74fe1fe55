// A "user-written" aux-reading Riemann solver for the plug-in mechanism: acoustics in a heterogeneous medium
// (riemann/src/rpn2_vc_acoustics.f90:62-101, rpt2_vc_acoustics.f90:27-77), aux = (density, sound speed).
// MAUX > 0 selects the aux-reading interface (clawpack_b200/csrc/rp2_aux.cuh): normal() receives the aux vectors of
// the two cells, transverse() those of the three cells the split fluctuation can enter.
#pragma once

namespace clawcuda {
namespace CLAW_NS {

struct UserVcAcoustics2D {
  static constexpr int MEQN = 3, MWAVES = 2, NROE = 2, NCELL = 0, MAUX = 2;

  __device__ __forceinline__ static void normal(const RpParams&, int, const double* ql, const double* qr,
                                                const double* auxl, const double* auxr, double (*wave)[MEQN], double* s,
                                                double* amdq, double* apdq, double* roe) {
    const double dp = qr[0] - ql[0], du = qr[1] - ql[1];
    const double zr = auxr[0] * auxr[1], zl = auxl[0] * auxl[1];      // impedances right / left of the interface
    const double a1 = ddiv(-dp + zr * du, zl + zr);
    const double a2 = ddiv(dp + zl * du, zl + zr);
    wave[0][0] = -a1 * zl; wave[0][1] = a1; wave[0][2] = 0.0;
    wave[1][0] = a2 * zr;  wave[1][1] = a2; wave[1][2] = 0.0;
    s[0] = -auxl[1]; s[1] = auxr[1];
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
    roe[0] = auxl[1]; roe[1] = auxr[1];
  }

  __device__ __forceinline__ static void speeds(const RpParams&, int, const double* roe, double* s) {
    s[0] = -roe[0]; s[1] = roe[1];
  }

  // a1/a2/a3: aux of the cell the fluctuation enters, in the pencil below / the same / above (b*: aux of cell i, unused)
  __device__ __forceinline__ static void transverse(const RpParams&, int, const double*, const double* a1x,
                                                    const double* a2x, const double* a3x, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    const double cm = a1x[1], cp = a3x[1];
    const double zm = a1x[0] * a1x[1], zz = a2x[0] * a2x[1], zp = a3x[0] * a3x[1];
    const double b1 = ddiv(-asdq[0] + asdq[2] * zz, zm + zz);
    const double b2 = ddiv(asdq[0] + asdq[2] * zz, zz + zp);
    bm[0] = cm * b1 * zm; bm[1] = 0.0; bm[2] = -cm * b1;
    bp[0] = cp * b2 * zp; bp[1] = 0.0; bp[2] = cp * b2;
  }
};

}  // namespace CLAW_NS
}  // namespace clawcuda
