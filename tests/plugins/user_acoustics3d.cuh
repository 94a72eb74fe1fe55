// A "user-written" 3-D Riemann solver for the plug-in mechanism (compile with -DCLAW_PLUGIN_NDIM=3): acoustics in a
// heterogeneous medium, ported from riemann/src/rpn3_vc_acoustics.f90:58-117, rpt3_vc_acoustics.f90:80-147 and
// rptt3_vc_acoustics.f90:86-177 to the functor interface of clawpack_b200/csrc/rp3.cuh.
//   q = (pressure, u, v, w);  aux = (impedance, sound speed)
//   normal<IXYZ>:                 rpn3 at the interface between the cells (ql, auxl) and (qr, auxr), sweep direction IXYZ
//   transverse<IXYZ, ICOOR, TT>:  one rpt3 (TT = false) or rptt3 (TT = true) split of `asdq` in the transverse direction
//                                 ICOOR (2: "y-like", 3: "z-like"); auxm / auxz / auxp = the cells below / at / above
#pragma once

namespace clawcuda {
namespace CLAW_NS {

struct UserAcoustics3D {
  static constexpr int MEQN = 4, MWAVES = 2, MAUX = 2;
  static constexpr int NCTX = 0;                       // nothing carried from normal() to transverse()
  // which components of a transverse split can be non-zero (all of them, in this plain port)
  template <int IXYZ, int ICOOR> struct TV {
    static constexpr int N = MEQN;
    __host__ __device__ static constexpr int comp(int k) { return k; }
  };

  template <int IXYZ>
  __device__ __forceinline__ static void normal(const RpParams&, const double* ql, const double* qr, const double* auxl,
                                                const double* auxr, double (*wave)[MEQN], double* s, double* amdq,
                                                double* apdq, double*) {
    constexpr int MU = 1 + IXYZ, MV = 1 + (IXYZ + 1) % 3, MW = 1 + (IXYZ + 2) % 3;
    const double dp = qr[0] - ql[0], du = qr[MU] - ql[MU];
    const double zr = auxr[0], zl = auxl[0];
    const double a1 = ddiv(-dp + zr * du, zl + zr);
    const double a2 = ddiv(dp + zl * du, zl + zr);
    wave[0][0] = -a1 * zl; wave[0][MU] = a1; wave[0][MV] = 0.0; wave[0][MW] = 0.0;
    wave[1][0] = a2 * zr;  wave[1][MU] = a2; wave[1][MV] = 0.0; wave[1][MW] = 0.0;
    s[0] = -auxl[1];
    s[1] = auxr[1];
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
  }

  template <int IXYZ, int ICOOR, bool TT = false>
  __device__ __forceinline__ static void transverse(const RpParams&, const double*, const double* auxm, const double* auxz,
                                                    const double* auxp, const double* asdq, double* bm, double* bp) {
    // velocity component along the split direction: iuvw = ixyz + icoor wrapped into 2..4 (1-based), here 0-based
    constexpr int IU = (IXYZ + 1 + ICOOR > 4) ? (IXYZ + ICOOR - 3) : (IXYZ + ICOOR);
    const double zm = auxm[0], zz = auxz[0], zp = auxp[0];
    const double a1 = ddiv(-asdq[0] + asdq[IU] * zz, zm + zz);
    const double a2 = ddiv(asdq[0] + asdq[IU] * zz, zz + zp);
    for (int m = 0; m < MEQN; ++m) { bm[m] = 0.0; bp[m] = 0.0; }
    bm[0] = auxm[1] * a1 * zm; bm[IU] = -auxm[1] * a1;
    bp[0] = auxp[1] * a2 * zp; bp[IU] = auxp[1] * a2;
  }
};

}  // namespace CLAW_NS
}  // namespace clawcuda
