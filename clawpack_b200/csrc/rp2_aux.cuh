// Point-wise 2-D Riemann solvers that read aux arrays (device functors for sweep2_aux.cuh).
// Rotated frame as in rp2.cuh: component 1 = momentum/velocity normal to the interface (`mu`), 2 = transverse (`mv`).
// normal():      ql/axl = state/aux of the LEFT cell (Fortran qr(:,i-1), auxr(:,i-1)), qr/axr = RIGHT cell (ql(:,i), auxl(:,i))
// transverse():  ax1/ax2/ax3 = aux of the cell the fluctuation enters (i1 = i-2+imp) in pencils 1, 2, 3
//                (rows j-1, j, j+1 for an x-sweep; columns i-1, i, i+1 for a y-sweep);
//                bx1/bx2/bx3 = aux of cell i (the cell on the right of the interface) in the same pencils -- equal to
//                ax* for imp = 2; only rpt2_acoustics_mapped reads geometry at i regardless of imp
#pragma once
#include "rp2.cuh"

namespace clawcuda {
namespace CLAW_NS {

// ------------------------------------------------------------------------------------
// vc_acoustics_2D -- riemann/src/rpn2_vc_acoustics.f90:62-101, rpt2_vc_acoustics.f90:27-77
// aux(1) = density, aux(2) = sound speed; no common block.   roe = (c_left, c_right)
// ------------------------------------------------------------------------------------
struct RpVcAcoustics2D {
  static constexpr int MEQN = 3, MWAVES = 2, NROE = 2, NCELL = 0, MAUX = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    const double delta1 = qr[0] - ql[0];
    const double delta2 = qr[1] - ql[1];
    const double zi = axr[0] * axr[1];                       // :76  auxl(1,i)*auxl(2,i)
    const double zim = axl[0] * axl[1];                      // :77  auxr(1,i-1)*auxr(2,i-1)
    const double a1 = ddiv(-delta1 + zi * delta2, zim + zi);
    const double a2 = ddiv(delta1 + zim * delta2, zim + zi);
    wave[0][0] = -a1 * zim; wave[0][1] = a1; wave[0][2] = 0.0;
    s[0] = -axl[1];
    wave[1][0] = a2 * zi;   wave[1][1] = a2; wave[1][2] = 0.0;
    s[1] = axr[1];
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
    roe[0] = axl[1]; roe[1] = axr[1];
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = -roe[0]; s[1] = roe[1];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double* ax1,
                                                    const double* ax2, const double* ax3, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    const double cm = ax1[1], cp = ax3[1];
    const double zm = ax1[0] * ax1[1];
    const double zz = ax2[0] * ax2[1];
    const double zp = ax3[0] * ax3[1];
    const double a1 = ddiv(-asdq[0] + asdq[2] * zz, zm + zz);     // :58-59
    const double a2 = ddiv(asdq[0] + asdq[2] * zz, zz + zp);      // :61-62
    bm[0] = cm * a1 * zm; bm[1] = 0.0; bm[2] = -cm * a1;
    bp[0] = cp * a2 * zp; bp[1] = 0.0; bp[2] = cp * a2;
  }
};

// ------------------------------------------------------------------------------------
// vc_advection_2D -- riemann/src/rpn2_vc_advection.f90, rpt2_vc_advection.f90 (colour equation, edge velocities)
// aux(1) = u at the left edge, aux(2) = v at the bottom edge of the cell
// ------------------------------------------------------------------------------------
struct RpVcAdvection2D {
  static constexpr int MEQN = 1, MWAVES = 1, NROE = 1, NCELL = 0, MAUX = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    wave[0][0] = qr[0] - ql[0];
    s[0] = axr[dir];                                          // auxl(ixy,i)
    amdq[0] = fmin2(axr[dir], 0.0) * wave[0][0];
    apdq[0] = fmax2(axr[dir], 0.0) * wave[0][0];
    roe[0] = axr[dir];
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = roe[0];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double* ax1,
                                                    const double* ax2, const double* ax3, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    const int kv = 1 - dir;                                   // kv = 3-ixy
    bm[0] = fmin2(ax2[kv], 0.0) * asdq[0];
    bp[0] = fmax2(ax3[kv], 0.0) * asdq[0];
  }
};

// ------------------------------------------------------------------------------------
// psystem_2D -- riemann/src/rpn2_psystem.f90:83-192, rpt2_psystem.f90:46-114 (f-wave solver, no common block)
// aux(1) = density, aux(2) = elastic modulus, aux(3) = stress-strain relation id (1 linear, 2 exponential, 3 cubic),
// aux(4) = strain of the cell (filled by the user's before_step for rpt2).   roe = (s1, s2)
// The exponential relation calls exp(): CUDA's and the host libm's results may differ in the last bit, so this
// material is compared with a tolerance; relations 1 and 3 are bit-exact in strict mode.
// ------------------------------------------------------------------------------------
struct RpPsystem2D {
  static constexpr int MEQN = 3, MWAVES = 2, NROE = 2, NCELL = 0, MAUX = 4;
  __device__ __forceinline__ static int nint_(double x) { return (int)llround(x); }
  __device__ __forceinline__ static double sigma(double eps, double E, int lin) {        // rpn2_psystem.f90:156-173
    switch (lin) {
      case 1: return E * eps;
      case 2: return exp(E * eps) - 1.0;
      case 3: return (double)0.1f * E * eps + 5.0 * (eps * eps * eps) * (E * E * E);       // `0.1` is single precision there
    }
    return 0.0;
  }
  __device__ __forceinline__ static double sigmap(double eps, double E, int lin) {       // rpn2_psystem.f90:175-192
    switch (lin) {
      case 1: return E;
      case 2: return E * exp(E * eps);
      case 3: return (double)0.1f * E + 15.0 * (eps * eps) * (E * E * E);
    }
    return 0.0;
  }
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    const double pim = axl[0], pi = axr[0], Eim = axl[1], Ei = axr[1];
    const int lin_i = nint_(axr[2]), lin_im = nint_(axl[2]);
    const double sigmai = sigma(qr[0], Ei, lin_i), sigmaim = sigma(ql[0], Eim, lin_im);
    const double sigmapi = sigmap(qr[0], Ei, lin_i), sigmapim = sigmap(ql[0], Eim, lin_im);
    const double r11 = ddiv(1.0, sqrt(sigmapim * pim));                           // :106-107
    const double r13 = -ddiv(1.0, sqrt(sigmapi * pi));
    s[0] = -sqrt(ddiv(sigmapim, pim));                                            // :109-110
    s[1] = sqrt(ddiv(sigmapi, pi));
    const double dF1 = -(ddiv(qr[1], pi) - ddiv(ql[1], pim));                     // :114 / :129 (normal momentum)
    const double dF2 = -(sigmai - sigmaim);
    const double beta1 = ddiv(dF1 - r13 * dF2, r11 - r13);
    const double beta3 = ddiv(-dF1 + r11 * dF2, r11 - r13);
    wave[0][0] = beta1 * r11; wave[0][1] = beta1 * 1.0; wave[0][2] = beta1 * 0.0;
    wave[1][0] = beta3 * r13; wave[1][1] = beta3 * 1.0; wave[1][2] = beta3 * 0.0;
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = wave[0][m]; apdq[m] = wave[1][m]; }
    roe[0] = s[0]; roe[1] = s[1];
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = roe[0]; s[1] = roe[1];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double* ax1,
                                                    const double* ax2, const double* ax3, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    const double sigmapjm = sigmap(ax1[3], ax1[1], nint_(ax1[2]));
    const double sigmapjp = sigmap(ax3[3], ax3[1], nint_(ax3[2]));
    const double r11 = ddiv(1.0, sqrt(sigmapjm * ax1[0]));
    const double r13 = -ddiv(1.0, sqrt(sigmapjp * ax3[0]));
    const double s1 = -sqrt(ddiv(sigmapjm, ax1[0]));
    const double s2 = sqrt(ddiv(sigmapjp, ax3[0]));
    const double gamma1 = ddiv(asdq[0] - r13 * asdq[2], r11 - r13);               // :86-87 / :100-101
    const double gamma3 = ddiv(-asdq[0] + r11 * asdq[2], r11 - r13);
    bm[0] = s1 * gamma1 * r11; bm[1] = s1 * gamma1 * 0.0; bm[2] = s1 * gamma1 * 1.0;
    bp[0] = s2 * gamma3 * r13; bp[1] = s2 * gamma3 * 0.0; bp[2] = s2 * gamma3 * 1.0;
  }
};

// ------------------------------------------------------------------------------------
// acoustics_mapped_2D -- riemann/src/rpn2_acoustics_mapped.f90:60-131, rpt2_acoustics_mapped.f90:41-88
// aux planes (0-based): 0-2 normal + length ratio of the left edge, 3-5 of the bottom edge, 6 capacity, 7 impedance,
// 8 sound speed.  The solver works with the PHYSICAL velocity components (no mu/mv swap), so the rotated frame of the
// sweep kernels is undone here: natural u = q[1 + dir], v = q[2 - dir].   roe = (s1, s2)
// ------------------------------------------------------------------------------------
struct RpAcousticsMapped2D {
  static constexpr int MEQN = 3, MWAVES = 2, NROE = 2, NCELL = 0, MAUX = 9;
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    const int iu = 1 + dir, iv = 2 - dir;                   // rotated slots of the natural components 2 and 3
    const int inx = 3 * dir, iny = 3 * dir + 1, ilen = 3 * dir + 2;
    const double alpha = axr[inx], beta = axr[iny];
    const double unorl = alpha * qr[iu] + beta * qr[iv];    // Fortran ql(:,i)   = right cell
    const double unorr = alpha * ql[iu] + beta * ql[iv];    // Fortran qr(:,i-1) = left cell
    const double d1 = qr[0] - ql[0];
    const double d2 = unorl - unorr;
    const double zi = axr[7], zim = axl[7], ci = axr[8], cim = axl[8];
    const double a1 = ddiv(-d1 + zi * d2, zim + zi);
    const double a2 = ddiv(d1 + zim * d2, zim + zi);
    wave[0][0] = -a1 * zim; wave[0][iu] = a1 * alpha; wave[0][iv] = a1 * beta;
    s[0] = -cim * axr[ilen];
    wave[1][0] = a2 * zi;   wave[1][iu] = a2 * alpha; wave[1][iv] = a2 * beta;
    s[1] = ci * axr[ilen];
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
    roe[0] = s[0]; roe[1] = s[1];
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = roe[0]; s[1] = roe[1];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double* ax1,
                                                    const double* ax2, const double* ax3, const double* bx1,
                                                    const double* bx2, const double* bx3, const double* asdq,
                                                    double* bm, double* bp) {
    const int iu = 1 + dir, iv = 2 - dir;
    const int ixt = 3 * (1 - dir), iyt = ixt + 1, ilen = ixt + 2;   // geometry of the TRANSVERSE edges
    const double asdqp = asdq[0];
    const double cm = ax1[8], cp = ax3[8];
    const double zm = ax1[7], zz = ax2[7], zp = ax3[7];
    double xtran = bx3[ixt], ytran = bx3[iyt];
    double asdqt = xtran * asdq[iu] + ytran * asdq[iv];
    const double a2 = ddiv(asdqp + zz * asdqt, zz + zp);
    const double bpasdqp = a2 * zp, bpasdqt = a2;
    bp[0] = cp * bpasdqp; bp[iu] = cp * xtran * bpasdqt; bp[iv] = cp * ytran * bpasdqt;
    xtran = bx2[ixt]; ytran = bx2[iyt];
    asdqt = xtran * asdq[iu] + ytran * asdq[iv];
    const double a1 = ddiv(-asdqp + zz * asdqt, zm + zz);
    const double bmasdqp = -a1 * zm, bmasdqt = a1;
    bm[0] = -cm * bmasdqp; bm[iu] = -cm * xtran * bmasdqt; bm[iv] = -cm * ytran * bmasdqt;
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { bm[m] = bm[m] * bx2[ilen]; bp[m] = bp[m] * bx3[ilen]; }
  }
};

// ------------------------------------------------------------------------------------
// shallow_bathymetry_fwave_2D -- riemann/src/rpn2_shallow_bathymetry_fwave.f90:60-140   cparam = (grav, dry_tolerance,
// sea_level), aux(1) = bathymetry, 3 f-waves.   roe = (s1, s2, s3).  No usable transverse solver (its rpt2 reads an
// unassigned s(2), rpt2_shallow_bathymetry_fwave.f90:80): validate_config refuses method(3) > 0.
// ------------------------------------------------------------------------------------
struct RpShallowBathymetryFwave2D {
  static constexpr int MEQN = 3, MWAVES = 3, NROE = 3, NCELL = 0, MAUX = 1;
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    const double grav = P.p[0], dry_tolerance = P.p[1];
    const double dry_l = (ql[0] < dry_tolerance) ? 0.0 : 1.0;
    const double dry_r = (qr[0] < dry_tolerance) ? 0.0 : 1.0;
    const double hl = ql[0] * dry_l;
    const double ul = ddiv(ql[1], ql[0]) * dry_l;
    const double vl = ddiv(ql[2], ql[0]) * dry_l;
    const double phil = (0.5 * grav * (hl * hl) + hl * (ul * ul)) * dry_l;
    const double bl = axl[0];
    const double hr = qr[0] * dry_r;
    const double ur = ddiv(qr[1], qr[0]) * dry_r;
    const double vr = ddiv(qr[2], qr[0]) * dry_r;
    const double phir = (0.5 * grav * (hr * hr) + hr * (ur * ur)) * dry_r;
    const double br = axr[0];
    const double hbar = 0.5 * (hr + hl);
    const double shr = sqrt(hr), shl = sqrt(hl);
    const double uhat = ddiv(shr * ur + shl * ul, shr + shl);
    const double chat = sqrt(grav * hbar);
    const double fd1 = hr * ur - hl * ul;
    const double fd2 = phir - phil + grav * hbar * (br - bl);
    const double fd3 = hr * ur * vr - hl * ul * vl;
    s[0] = fmin2(uhat - chat, ul - sqrt(grav * hl));
    s[2] = fmax2(uhat + chat, ur + sqrt(grav * hr));
    s[1] = 0.5 * (s[0] + s[2]);
    const double beta1 = ddiv(s[2] * fd1 - fd2, s[2] - s[0]);
    const double beta3 = ddiv(fd2 - s[0] * fd1, s[2] - s[0]);
    const double beta2 = fd3 - beta1 * vl - beta3 * vr;
    wave[0][0] = beta1 * 1.0; wave[0][1] = beta1 * s[0]; wave[0][2] = beta1 * vl;
    wave[1][0] = beta2 * 0.0; wave[1][1] = beta2 * 0.0;  wave[1][2] = beta2 * 1.0;
    wave[2][0] = beta3 * 1.0; wave[2][1] = beta3 * s[2]; wave[2][2] = beta3 * vr;
    const double eps = (double)1.0e-14f;                      // single-precision literals in the Fortran source
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = 0.0; apdq[m] = 0.0; }
#pragma unroll
    for (int k = 0; k < MWAVES; ++k) {
      if (s[k] < -eps) {
#pragma unroll
        for (int m = 0; m < MEQN; ++m) amdq[m] = amdq[m] + wave[k][m];
      } else if (s[k] > eps) {
#pragma unroll
        for (int m = 0; m < MEQN; ++m) apdq[m] = apdq[m] + wave[k][m];
      } else {
#pragma unroll
        for (int m = 0; m < MEQN; ++m) { amdq[m] = amdq[m] + 0.5 * wave[k][m]; apdq[m] = apdq[m] + 0.5 * wave[k][m]; }
      }
    }
    roe[0] = s[0]; roe[1] = s[1]; roe[2] = s[2];
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = roe[0]; s[1] = roe[1]; s[2] = roe[2];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double*,
                                                    const double*, const double*, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { bm[m] = 0.0; bp[m] = 0.0; }
  }
};

// ------------------------------------------------------------------------------------
// vc_elasticity_2D -- riemann/src/rpn2_vc_elasticity.f90:60-183, rpt2_vc_elasticity.f90:50-176
// q = (sigma11, sigma22, sigma12, u, v); aux = (density, lambda, mu, cp, cs).  4 waves: P-, P+, S-, S+ (not ordered
// by speed).  roe = (cp_left, cp_right, cs_left, cs_right).
// The reference swaps sigma11 <-> sigma22 and u <-> v for a y-sweep (ksig11/ksig22/ku/kv), which is not the swap of
// components 2 and 3 the sweep kernels apply (rot<5>): the functor therefore addresses the NATURAL components through
// their rotated slots -- natural component n sits in slot rot<5>(n) when dir = 1.
// ------------------------------------------------------------------------------------
struct RpVcElasticity2D {
  static constexpr int MEQN = 5, MWAVES = 4, NROE = 4, NCELL = 0, MAUX = 5;
  // slots of (sigma_nn, sigma_tt, sigma_12, u_n, u_t): n = sweep direction, t = transverse
  __device__ __forceinline__ static void slots(int dir, int& knn, int& ktt, int& k12, int& kun, int& kut) {
    if (dir == 0) { knn = 0; ktt = 1; k12 = 2; kun = 3; kut = 4; }
    else          { knn = 2; ktt = 0; k12 = 1; kun = 4; kut = 3; }      // natural 1, 0, 2, 4, 3 through rot<5>
  }
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    int knn, ktt, k12, kun, kut;
    slots(dir, knn, ktt, k12, kun, kut);
    const double dsig11 = qr[knn] - ql[knn];
    const double dsig12 = qr[k12] - ql[k12];
    const double du = qr[kun] - ql[kun];
    const double dv = qr[kut] - ql[kut];
    const double alamr = axr[1], amur = axr[2];
    const double bulkr = alamr + 2.0 * amur;
    const double cpr = axr[3], csr = axr[4];
    const double alaml = axl[1], amul = axl[2];
    const double bulkl = alaml + 2.0 * amul;
    const double cpl = axl[3], csl = axl[4];
    const double det = bulkl * cpr + bulkr * cpl;
    const double a1 = ddiv(cpr * dsig11 + bulkr * du, det);
    const double a2 = ddiv(cpl * dsig11 - bulkl * du, det);
    const double det2 = amul * csr + amur * csl;
    double a3 = 0.0, a4 = 0.0;
    if (det2 != 0.0) {
      a3 = ddiv(csr * dsig12 + amur * dv, det2);
      a4 = ddiv(csl * dsig12 - amul * dv, det2);
    }
#pragma unroll
    for (int mw = 0; mw < MWAVES; ++mw)
#pragma unroll
      for (int m = 0; m < MEQN; ++m) wave[mw][m] = 0.0;
    wave[0][knn] = a1 * bulkl; wave[0][ktt] = a1 * alaml; wave[0][kun] = a1 * cpl;  s[0] = -cpl;
    wave[1][knn] = a2 * bulkr; wave[1][ktt] = a2 * alamr; wave[1][kun] = -a2 * cpr; s[1] = cpr;
    wave[2][k12] = a3 * amul;  wave[2][kut] = a3 * csl;                             s[2] = -csl;
    wave[3][k12] = a4 * amur;  wave[3][kut] = -a4 * csr;                            s[3] = csr;
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      amdq[m] = s[0] * wave[0][m] + s[2] * wave[2][m];
      apdq[m] = s[1] * wave[1][m] + s[3] * wave[3][m];
    }
    roe[0] = cpl; roe[1] = cpr; roe[2] = csl; roe[3] = csr;
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = -roe[0]; s[1] = roe[1]; s[2] = -roe[2]; s[3] = roe[3];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double* ax1,
                                                    const double* ax2, const double* ax3, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    int knn, ktt, k12, kun, kut;
    slots(dir, knn, ktt, k12, kun, kut);
    const double dsig22 = asdq[ktt], dsig12 = asdq[k12], du = asdq[kun], dv = asdq[kut];
    const double alamm = ax1[1], alam = ax2[1], alamp = ax3[1];
    const double amum = ax1[2], amu = ax2[2], amup = ax3[2];
    const double bulkm = alamm + 2.0 * amum, bulk = alam + 2.0 * amu, bulkp = alamp + 2.0 * amup;
    const double cpm = ax1[3], cp = ax2[3], cpp = ax3[3];
    const double csm = ax1[4], cs = ax2[4], csp = ax3[4];
    const double a1 = ddiv(cp * dsig22 + bulk * dv, bulkm * cp + bulk * cpm);
    const double a2 = ddiv(cp * dsig22 - bulk * dv, bulk * cpp + bulkp * cp);
    const double det3 = -(amum * cs + amu * csm), det4 = -(amu * csp + amup * cs);
    const double a3 = (det3 == 0.0) ? 0.0 : ddiv(cs * dsig12 + amu * du, det3);
    const double a4 = (det4 == 0.0) ? 0.0 : ddiv(cs * dsig12 - amu * du, det4);
    bm[knn] = -cpm * a1 * alamm;
    bm[ktt] = -cpm * a1 * bulkm;
    bm[k12] = -csm * a3 * (-amum);
    bm[kun] = -csm * a3 * (-csm);
    bm[kut] = -cpm * a1 * cpm;
    bp[knn] = cpp * a2 * alamp;
    bp[ktt] = cpp * a2 * bulkp;
    bp[k12] = csp * a4 * (-amup);
    bp[kun] = csp * a4 * csp;
    bp[kut] = cpp * a2 * (-cpp);
  }
};

// ------------------------------------------------------------------------------------
// vc_advection_vector_2D -- pyclaw/examples/advection_reaction_2d/rpn2_vc_advection_vector.f90:40-47 (+ rpt2_dummy.f90)
// Colour equation for two unknowns sharing one velocity field; aux(1) = u at the left edge, aux(2) = v at the bottom
// edge of the cell (read from the cell on the RIGHT of the interface).   roe = (s).  No transverse solver.
// ------------------------------------------------------------------------------------
struct RpVcAdvectionVector2D {
  static constexpr int MEQN = 2, MWAVES = 1, NROE = 1, NCELL = 0, MAUX = 2;
  __device__ __forceinline__ static void normal(const RpParams& P, int dir, const double* ql, const double* qr,
                                                const double* axl, const double* axr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double* roe) {
    const double vel = axr[dir];
    s[0] = vel;
    const double vm = fmin2(vel, 0.0), vp = fmax2(vel, 0.0);
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      wave[0][m] = qr[m] - ql[m];
      amdq[m] = vm * wave[0][m];
      apdq[m] = vp * wave[0][m];
    }
    roe[0] = vel;
  }
  __device__ __forceinline__ static void speeds(const RpParams& P, int dir, const double* roe, double* s) {
    s[0] = roe[0];
  }
  __device__ __forceinline__ static void transverse(const RpParams& P, int dir, const double* roe, const double*,
                                                    const double*, const double*, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { bm[m] = 0.0; bp[m] = 0.0; }
  }
};

}  // namespace CLAW_NS
}  // namespace clawcuda
