// Point-wise 3-D Riemann solvers as device functors (normal + transverse / double-transverse),
// sm_100a fp64.  Each functor restates one clawpack.riemann module for ONE interface.
//
// Components stay in their natural order; the sweep direction IXYZ (0,1,2 = x,y,z) and the
// transverse direction ICOOR (2 = "y-like", 3 = "z-like", rpt3_vc_acoustics.f90:23-43) are
// template parameters, so every component index is a compile-time constant (arrays stay in
// registers).
//
//   normal<IXYZ>(P, ql, qr, auxl, auxr, wave, s, amdq, apdq, ctx)
//       ql/auxl = cell i-1, qr/auxr = cell i.  ctx[NCTX] = whatever transverse() needs from the
//       interface (Roe averages); the reference recomputes it inside rpt3/rptt3.
//   transverse<IXYZ,ICOOR>(P, ctx, auxm, auxz, auxp, asdq, bm, bp)
//       one rpt3 OR rptt3 solve: the two differ only in WHICH three aux cells (below / at /
//       above in the split direction) they read (rpt3_vc_acoustics.f90:100-122 vs
//       rptt3_vc_acoustics.f90:106-152); the caller (flux3_side, sweep3.cuh) selects them.
//
// Compiled twice like rp2.cuh: namespace strict (-fmad=false, op-for-op) and fast.
#pragma once
#include "rp2.cuh"

namespace clawcuda {
namespace CLAW_NS {

template <int IXYZ> struct Dirs3 {
  // 0-based component indices of the normal / y-like / z-like velocity (rpn3_vc_acoustics.f90:58-70)
  static constexpr int MU = 1 + IXYZ;
  static constexpr int MV = 1 + (IXYZ + 1) % 3;
  static constexpr int MW = 1 + (IXYZ + 2) % 3;
};

// ------------------------------------------------------------------------------------
// vc_acoustics_3D -- riemann/src/rpn3_vc_acoustics.f90:58-117, rpt3_vc_acoustics.f90:80-147,
//                    rptt3_vc_acoustics.f90:86-177.   aux = (impedance, sound speed)
// ------------------------------------------------------------------------------------
struct RpVcAcoustics3D {
  static constexpr int MEQN = 4, MWAVES = 2, MAUX = 2, NCTX = 0;
  // Structurally non-zero components of a transverse solve's outputs: pressure and the velocity along the split
  // direction (rpt3_vc_acoustics.f90:134-145 sets the other two to 0.d0).  Fast mode stores / adds only those
  // (adding the reference's exact zeros cannot change a value); strict mode keeps all four, op for op.
  template <int IXYZ, int ICOOR> struct TV {
    static constexpr int IU = (IXYZ + 1 + ICOOR > 4) ? (IXYZ + 1 + ICOOR - 3 - 1) : (IXYZ + 1 + ICOOR - 1);
#ifdef CLAW_STRICT
    static constexpr int N = MEQN;
    __host__ __device__ static constexpr int comp(int k) { return k; }
#else
    static constexpr int N = 2;
    __host__ __device__ static constexpr int comp(int k) { return k == 0 ? 0 : IU; }
#endif
  };
  template <int IXYZ>
  __device__ __forceinline__ static void normal(const RpParams&, const double* ql, const double* qr,
                                                const double* auxl, const double* auxr, double (*wave)[MEQN],
                                                double* s, double* amdq, double* apdq, double*) {
    constexpr int MU = Dirs3<IXYZ>::MU, MV = Dirs3<IXYZ>::MV, MW = Dirs3<IXYZ>::MW;
    const double delta1 = qr[0] - ql[0];
    const double delta2 = qr[MU] - ql[MU];
    const double zi = auxr[0];
    const double zim = auxl[0];
#ifdef CLAW_STRICT
    const double a1 = ddiv(-delta1 + zi * delta2, zim + zi);
    const double a2 = ddiv(delta1 + zim * delta2, zim + zi);
#else
    const double rz = rcp_fast(zim + zi);
    const double a1 = (-delta1 + zi * delta2) * rz;
    const double a2 = (delta1 + zim * delta2) * rz;
#endif
    wave[0][0] = -a1 * zim; wave[0][MU] = a1; wave[0][MV] = 0.0; wave[0][MW] = 0.0;
    s[0] = -auxl[1];
    wave[1][0] = a2 * zi; wave[1][MU] = a2; wave[1][MV] = 0.0; wave[1][MW] = 0.0;
    s[1] = auxr[1];
#pragma unroll
    for (int m = 0; m < MEQN; ++m) { amdq[m] = s[0] * wave[0][m]; apdq[m] = s[1] * wave[1][m]; }
  }
  // TT: a double-transverse (rptt3) solve -- its input is the output of a split in the OTHER transverse
  // direction, whose velocity component along this direction is an exact zero
  template <int IXYZ, int ICOOR, bool TT = false>
  __device__ __forceinline__ static void transverse(const RpParams&, const double*, const double* auxm,
                                                    const double* auxz, const double* auxp, const double* asdq,
                                                    double* bm, double* bp) {
    // iuvw = ixyz + icoor, wrapped into 2..4 (rpt3_vc_acoustics.f90:80-81); 0-based here
    constexpr int IU = (IXYZ + 1 + ICOOR > 4) ? (IXYZ + 1 + ICOOR - 3 - 1) : (IXYZ + 1 + ICOOR - 1);
    const double zm = auxm[0], zz = auxz[0], zp = auxp[0], cm = auxm[1], cp = auxp[1];
#ifdef CLAW_STRICT
    const double a1 = ddiv(-asdq[0] + asdq[IU] * zz, zm + zz);
    const double a2 = ddiv(asdq[0] + asdq[IU] * zz, zz + zp);
#else
    // the reciprocals depend on the aux cells only: the 4 solves of a cell that share an aux triple (amdq, apdq
    // and the two cqxx splits) reuse them after inlining (common-subexpression elimination): 12 reciprocals per
    // cell and sweep instead of 32 divisions
    const double a1 = (TT ? -asdq[0] : (-asdq[0] + asdq[IU] * zz)) * rcp_fast(zm + zz);
    const double a2 = (TT ? asdq[0] : (asdq[0] + asdq[IU] * zz)) * rcp_fast(zz + zp);
#endif
    bm[0] = cm * a1 * zm; bm[1] = 0.0; bm[2] = 0.0; bm[3] = 0.0; bm[IU] = -cm * a1;
    bp[0] = cp * a2 * zp; bp[1] = 0.0; bp[2] = 0.0; bp[3] = 0.0; bp[IU] = cp * a2;
  }
};

// ------------------------------------------------------------------------------------
// euler_3D -- riemann/src/rpn3_euler.f90, rpt3_euler.f90, rptt3_euler.f90    cparam = (gamma)
// ctx = (u, v, w, enth, a, g1a2, euv, u2v2w2)  (the Roe averages; v, w = y-like, z-like)
// ------------------------------------------------------------------------------------
struct RpEuler3D {
  static constexpr int MEQN = 5, MWAVES = 3, MAUX = 0, NCTX = 8;
  template <int IXYZ, int ICOOR> struct TV {          // every component of a transverse output can be non-zero
    static constexpr int N = MEQN;
    __host__ __device__ static constexpr int comp(int k) { return k; }
  };
  template <int IXYZ>
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double*, const double*, double (*wave)[MEQN], double* s,
                                                double* amdq, double* apdq, double* ctx) {
    constexpr int MU = Dirs3<IXYZ>::MU, MV = Dirs3<IXYZ>::MV, MW = Dirs3<IXYZ>::MW;
    const double gamma = P.p[0];
    const double gamma1 = gamma - 1.0;
    // Roe averages, rpn3_euler.f90:138-180 (left state qr(:,i-1) = ql here)
#ifdef CLAW_STRICT
    const double rhsqrtl = sqrt(ql[0]);
    const double rhsqrtr = sqrt(qr[0]);
    const double pl = gamma1 * (ql[4] - ddiv(0.5 * (ql[MU] * ql[MU] + ql[MV] * ql[MV] + ql[MW] * ql[MW]), ql[0]));
    const double pr = gamma1 * (qr[4] - ddiv(0.5 * (qr[MU] * qr[MU] + qr[MV] * qr[MV] + qr[MW] * qr[MW]), qr[0]));
    const double rhsq2 = rhsqrtl + rhsqrtr;
    const double u = ddiv(ddiv(ql[MU], rhsqrtl) + ddiv(qr[MU], rhsqrtr), rhsq2);
    const double v = ddiv(ddiv(ql[MV], rhsqrtl) + ddiv(qr[MV], rhsqrtr), rhsq2);
    const double w = ddiv(ddiv(ql[MW], rhsqrtl) + ddiv(qr[MW], rhsqrtr), rhsq2);
    const double enth = ddiv(ddiv(ql[4] + pl, rhsqrtl) + ddiv(qr[4] + pr, rhsqrtr), rhsq2);
    const double u2v2w2 = u * u + v * v + w * w;
    const double asqrd = gamma1 * (enth - .5 * u2v2w2);
    const double a = sqrt(asqrd);
    const double g1a2 = ddiv(gamma1, asqrd);
#else
    // fast mode: rsqrt / reciprocal sharing as in RpEuler4Wave2D (rp2.cuh)
    const double irl = rsqrt_fast(ql[0]), irr = rsqrt_fast(qr[0]);
    const double iql0 = irl * irl, iqr0 = irr * irr;
    const double pl = gamma1 * (ql[4] - 0.5 * (ql[MU] * ql[MU] + ql[MV] * ql[MV] + ql[MW] * ql[MW]) * iql0);
    const double pr = gamma1 * (qr[4] - 0.5 * (qr[MU] * qr[MU] + qr[MV] * qr[MV] + qr[MW] * qr[MW]) * iqr0);
    const double irs = rcp_fast(ql[0] * irl + qr[0] * irr);
    const double u = (ql[MU] * irl + qr[MU] * irr) * irs;
    const double v = (ql[MV] * irl + qr[MV] * irr) * irs;
    const double w = (ql[MW] * irl + qr[MW] * irr) * irs;
    const double enth = ((ql[4] + pl) * irl + (qr[4] + pr) * irr) * irs;
    const double u2v2w2 = u * u + v * v + w * w;
    const double asqrd = gamma1 * (enth - .5 * u2v2w2);
    const double ia = rsqrt_fast(asqrd);
    const double a = asqrd * ia;
    const double g1a2 = gamma1 * (ia * ia);
    const double r2a = 0.5 * ia;
#endif
    const double euv = enth - u2v2w2;
    ctx[0] = u; ctx[1] = v; ctx[2] = w; ctx[3] = enth; ctx[4] = a; ctx[5] = g1a2; ctx[6] = euv; ctx[7] = u2v2w2;
#ifndef CLAW_STRICT
    ctx[6] = r2a;      // fast mode carries 1/(2a) instead of euv (= enth - u2v2w2, one subtraction to rebuild)
#endif
    // waves, :184-221
    const double d1 = qr[0] - ql[0];
    const double d2 = qr[MU] - ql[MU];
    const double d3 = qr[MV] - ql[MV];
    const double d4 = qr[MW] - ql[MW];
    const double d5 = qr[4] - ql[4];
    const double al4 = g1a2 * (euv * d1 + u * d2 + v * d3 + w * d4 - d5);
    const double al2 = d3 - v * d1;
    const double al3 = d4 - w * d1;
#ifdef CLAW_STRICT
    const double al5 = ddiv(d2 + (a - u) * d1 - a * al4, 2.0 * a);
#else
    const double al5 = (d2 + (a - u) * d1 - a * al4) * r2a;
#endif
    const double al1 = d1 - al4 - al5;
    wave[0][0] = al1; wave[0][MU] = al1 * (u - a); wave[0][MV] = al1 * v; wave[0][MW] = al1 * w;
    wave[0][4] = al1 * (enth - u * a);
    s[0] = u - a;
    wave[1][0] = al4; wave[1][MU] = al4 * u; wave[1][MV] = al4 * v + al2; wave[1][MW] = al4 * w + al3;
    wave[1][4] = al4 * 0.5 * u2v2w2 + al2 * v + al3 * w;
    s[1] = u;
    wave[2][0] = al5; wave[2][MU] = al5 * (u + a); wave[2][MV] = al5 * v; wave[2][MW] = al5 * w;
    wave[2][4] = al5 * (enth + u * a);
    s[2] = u + a;
    // entropy fix, :244-303 (amdq pre-zeroed, :224)
#pragma unroll
    for (int m = 0; m < MEQN; ++m) amdq[m] = 0.0;
#ifdef CLAW_STRICT
    const double cim1 = sqrt(ddiv(gamma * pl, ql[0]));
    const double s0 = ddiv(ql[MU], ql[0]) - cim1;
#else
    const double cim1 = sqrt_fast(gamma * pl * iql0);
    const double s0 = ql[MU] * iql0 - cim1;
#endif
    if (!(s0 >= 0.0 && s[0] > 0.0)) {
      const double rho1 = ql[0] + wave[0][0];
      const double rhou1 = ql[MU] + wave[0][MU];
      const double rhov1 = ql[MV] + wave[0][MV];
      const double rhow1 = ql[MW] + wave[0][MW];
      const double en1 = ql[4] + wave[0][4];
#ifdef CLAW_STRICT
      const double p1 = gamma1 * (en1 - ddiv(0.5 * (rhou1 * rhou1 + rhov1 * rhov1 + rhow1 * rhow1), rho1));
      const double c1 = sqrt(ddiv(gamma * p1, rho1));
      const double s1 = ddiv(rhou1, rho1) - c1;
#else
      const double irho1 = rcp_fast(rho1);
      const double p1 = gamma1 * (en1 - 0.5 * (rhou1 * rhou1 + rhov1 * rhov1 + rhow1 * rhow1) * irho1);
      const double c1 = sqrt_fast(gamma * p1 * irho1);
      const double s1 = rhou1 * irho1 - c1;
#endif
      double sfract;
      if (s0 < 0.0 && s1 > 0.0) sfract = ddiv(s0 * (s1 - s[0]), s1 - s0);
      else if (s[0] < 0.0) sfract = s[0];
      else sfract = 0.0;
#pragma unroll
      for (int m = 0; m < MEQN; ++m) amdq[m] = sfract * wave[0][m];
      if (!(s[1] >= 0.0)) {
#pragma unroll
        for (int m = 0; m < MEQN; ++m) amdq[m] = amdq[m] + s[1] * wave[1][m];
        const double rho2 = qr[0] - wave[2][0];
        const double rhou2 = qr[MU] - wave[2][MU];
        const double rhov2 = qr[MV] - wave[2][MV];
        const double rhow2 = qr[MW] - wave[2][MW];
        const double en2 = qr[4] - wave[2][4];
#ifdef CLAW_STRICT
        const double ci = sqrt(ddiv(gamma * pr, qr[0]));
        const double s3 = ddiv(qr[MU], qr[0]) + ci;
        const double p2 = gamma1 * (en2 - ddiv(0.5 * (rhou2 * rhou2 + rhov2 * rhov2 + rhow2 * rhow2), rho2));
        const double c2 = sqrt(ddiv(gamma * p2, rho2));
        const double s2 = ddiv(rhou2, rho2) + c2;
#else
        const double ci = sqrt_fast(gamma * pr * iqr0);
        const double s3 = qr[MU] * iqr0 + ci;
        const double irho2 = rcp_fast(rho2);
        const double p2 = gamma1 * (en2 - 0.5 * (rhou2 * rhou2 + rhov2 * rhov2 + rhow2 * rhow2) * irho2);
        const double c2 = sqrt_fast(gamma * p2 * irho2);
        const double s2 = rhou2 * irho2 + c2;
#endif
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
    for (int m = 0; m < MEQN; ++m) {                                        // :306-314
      double df = from0(s[0] * wave[0][m]);
      df = df + s[1] * wave[1][m];
      df = df + s[2] * wave[2][m];
      apdq[m] = df - amdq[m];
    }
  }
  template <int IXYZ, int ICOOR, bool TT = false>
  __device__ __forceinline__ static void transverse(const RpParams&, const double* ctx, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    constexpr int MU = Dirs3<IXYZ>::MU, MV = Dirs3<IXYZ>::MV, MW = Dirs3<IXYZ>::MW;
#ifdef CLAW_STRICT
    const double u = ctx[0], v = ctx[1], w = ctx[2], enth = ctx[3], a = ctx[4], g1a2 = ctx[5], euv = ctx[6],
                 u2v2w2 = ctx[7];
#else
    const double u = ctx[0], v = ctx[1], w = ctx[2], enth = ctx[3], a = ctx[4], g1a2 = ctx[5], r2a = ctx[6],
                 u2v2w2 = ctx[7];
    const double euv = enth - u2v2w2;
#endif
    // rpt3_euler.f90:118-187 (== rptt3_euler.f90)
    const double al4 = g1a2 * (euv * asdq[0] + u * asdq[MU] + v * asdq[MV] + w * asdq[MW] - asdq[4]);
    const double al2 = asdq[MU] - u * asdq[0];
    double wb[3][MEQN];
    double sb0, sb1, sb2;
    if (ICOOR == 2) {
      const double al3 = asdq[MW] - w * asdq[0];
#ifdef CLAW_STRICT
      const double al5 = ddiv(asdq[MV] + (a - v) * asdq[0] - a * al4, 2.0 * a);
#else
      const double al5 = (asdq[MV] + (a - v) * asdq[0] - a * al4) * r2a;
#endif
      const double al1 = asdq[0] - al4 - al5;
      wb[0][0] = al1; wb[0][MU] = al1 * u; wb[0][MV] = al1 * (v - a); wb[0][MW] = al1 * w;
      wb[0][4] = al1 * (enth - v * a);
      sb0 = v - a;
      wb[1][0] = al4; wb[1][MU] = al2 + u * al4; wb[1][MV] = al4 * v; wb[1][MW] = al3 + al4 * w;
      wb[1][4] = al4 * 0.5 * u2v2w2 + al2 * u + al3 * w;
      sb1 = v;
      wb[2][0] = al5; wb[2][MU] = al5 * u; wb[2][MV] = al5 * (v + a); wb[2][MW] = al5 * w;
      wb[2][4] = al5 * (enth + v * a);
      sb2 = v + a;
    } else {
      const double al3 = asdq[MV] - v * asdq[0];
#ifdef CLAW_STRICT
      const double al5 = ddiv(asdq[MW] + (a - w) * asdq[0] - a * al4, 2.0 * a);
#else
      const double al5 = (asdq[MW] + (a - w) * asdq[0] - a * al4) * r2a;
#endif
      const double al1 = asdq[0] - al4 - al5;
      wb[0][0] = al1; wb[0][MU] = al1 * u; wb[0][MV] = al1 * v; wb[0][MW] = al1 * (w - a);
      wb[0][4] = al1 * (enth - w * a);
      sb0 = w - a;
      wb[1][0] = al4; wb[1][MU] = al2 + al4 * u; wb[1][MV] = al3 + al4 * v; wb[1][MW] = al4 * w;
      wb[1][4] = al4 * 0.5 * u2v2w2 + al2 * u + al3 * v;
      sb1 = w;
      wb[2][0] = al5; wb[2][MU] = al5 * u; wb[2][MV] = al5 * v; wb[2][MW] = al5 * (w + a);
      wb[2][4] = al5 * (enth + w * a);
      sb2 = w + a;
    }
    const double m0 = fmin2(sb0, 0.0), m1 = fmin2(sb1, 0.0), m2 = fmin2(sb2, 0.0);
    const double p0 = fmax2(sb0, 0.0), p1 = fmax2(sb1, 0.0), p2 = fmax2(sb2, 0.0);
#pragma unroll
    for (int m = 0; m < MEQN; ++m) {
      double b = from0(m0 * wb[0][m]); b = b + m1 * wb[1][m]; b = b + m2 * wb[2][m];
      bm[m] = b;
      double c = from0(p0 * wb[0][m]); c = c + p1 * wb[1][m]; c = c + p2 * wb[2][m];
      bp[m] = c;
    }
  }
};

// ------------------------------------------------------------------------------------
// burgers_3D -- riemann/src/rpn3_burgers.f90:50-68, rpt3_burgers.f90:55-61, rptt3_burgers.f90:60-66
// cparam = coeff(3); ctx = (q_left + q_right) of the interface (the transverse speeds are built from it)
// ------------------------------------------------------------------------------------
struct RpBurgers3D {
  static constexpr int MEQN = 1, MWAVES = 1, MAUX = 0, NCTX = 1;
  template <int IXYZ, int ICOOR> struct TV {
    static constexpr int N = MEQN;
    __host__ __device__ static constexpr int comp(int k) { return k; }
  };
  template <int IXYZ>
  __device__ __forceinline__ static void normal(const RpParams& P, const double* ql, const double* qr,
                                                const double*, const double*, double (*wave)[MEQN], double* s,
                                                double* amdq, double* apdq, double* ctx) {
    const double coeff = P.p[IXYZ];
    wave[0][0] = qr[0] - ql[0];
    ctx[0] = ql[0] + qr[0];
    s[0] = coeff * 0.5 * ctx[0];
    amdq[0] = fmin2(s[0], 0.0) * wave[0][0];
    apdq[0] = fmax2(s[0], 0.0) * wave[0][0];
    if (ql[0] < 0.0 && qr[0] > 0.0) {                        // transonic rarefaction (:61-66)
      amdq[0] = -(coeff * 0.5 * (ql[0] * ql[0]));
      apdq[0] = coeff * 0.5 * (qr[0] * qr[0]);
    }
  }
  template <int IXYZ, int ICOOR, bool TT = false>
  __device__ __forceinline__ static void transverse(const RpParams& P, const double* ctx, const double*, const double*,
                                                    const double*, const double* asdq, double* bm, double* bp) {
    constexpr int IUVW = (IXYZ + ICOOR - 1) % 3;             // 0-based form of iuvw = ixyz + icoor - 1 (wrapped)
    const double sb = P.p[IUVW] * 0.5 * ctx[0];
    bm[0] = fmin2(sb, 0.0) * asdq[0];
    bp[0] = fmax2(sb, 0.0) * asdq[0];
  }
};

// ------------------------------------------------------------------------------------
// vc_advection_3D -- riemann/src/rpn3_vc_advection.f90:52-57, rpt3_vc_advection.f90:60-72, rptt3_vc_advection.f90:62-84
// aux = edge velocities (u, v, w at the lower x-, y-, z-face of the cell).  In every rpt3 / rptt3 variant the
// down-going part uses the velocity of the cell "at" and the up-going part that of the cell "above" in the split
// direction -- exactly the (auxz, auxp) pair sweep3.cuh hands over.
// ------------------------------------------------------------------------------------
struct RpVcAdvection3D {
  static constexpr int MEQN = 1, MWAVES = 1, MAUX = 3, NCTX = 0;
  template <int IXYZ, int ICOOR> struct TV {
    static constexpr int N = MEQN;
    __host__ __device__ static constexpr int comp(int k) { return k; }
  };
  template <int IXYZ>
  __device__ __forceinline__ static void normal(const RpParams&, const double* ql, const double* qr,
                                                const double*, const double* auxr, double (*wave)[MEQN], double* s,
                                                double* amdq, double* apdq, double*) {
    wave[0][0] = qr[0] - ql[0];
    s[0] = auxr[IXYZ];                                       // auxl(ixyz, i)
    amdq[0] = fmin2(auxr[IXYZ], 0.0) * wave[0][0];
    apdq[0] = fmax2(auxr[IXYZ], 0.0) * wave[0][0];
  }
  template <int IXYZ, int ICOOR, bool TT = false>
  __device__ __forceinline__ static void transverse(const RpParams&, const double*, const double*, const double* auxz,
                                                    const double* auxp, const double* asdq, double* bm, double* bp) {
    constexpr int IUVW = (IXYZ + ICOOR - 1) % 3;
    bm[0] = fmin2(auxz[IUVW], 0.0) * asdq[0];
    bp[0] = fmax2(auxp[IUVW], 0.0) * asdq[0];
  }
};

}  // namespace CLAW_NS
}  // namespace clawcuda
