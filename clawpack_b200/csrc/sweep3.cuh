// Fused 3-D classic wave-propagation sweeps for sm_100a (fp64 stencil, FP64-pipe bound; no
// tensor cores).  One kernel per sweep direction does what the reference does in
// step3.f90 / step3ds.f90 + flux3.f90 + limiter.f90 + philim.f90 + rpn3/rpt3/rptt3 for that
// direction: normal solve, CFL max, wave limiter, second-order correction, the 4+4 transverse
// and 8 double-transverse solves, and the update of the 3x3 neighbourhood of pencils.
//
// Design ("2.5-D blocking"): a CTA owns a 2-D tile of columns (thread-x <-> i, thread-y <-> j
// or k) and marches along the third direction.
//   * The transverse solves are organised per CELL, not per interface: cell c owns the
//     "p-side" of interface c (apdq, imp = 2) and the "m-side" of interface c+1 (amdq,
//     imp = 1) -- exactly the solves that read the aux column of cell c (i1 = i-2+imp,
//     rpt3_vc_acoustics.f90:100) and that add into gadd/hadd(.,.,.,c) (flux3.f90:390-469,
//     513-594).  A cell therefore builds its own 12 gadd/hadd vectors with the reference's
//     statement order and nothing but amdq/cqxx of the next interface crosses threads.
//   * The reference's scatter (each pencil adds into 9 pencils of qnew, step3.f90:193-226)
//     becomes a gather: the 12 vectors go to shared memory once, and each cell adds the nine
//     contributions it receives in the reference's pencil order, each as the same two
//     sequential addends (g term, then h term).  No atomics, bit-reproducible.
//   * x-sweep (sweep3_x): tile (i, j), lanes along i (neighbouring interfaces by warp shuffle),
//     march in k with three running accumulators (planes k-1, k, k+1).
//   * y-/z-sweep (sweep3_march): tile (i, k) / (i, j), march ALONG the pencil: consecutive
//     interfaces are produced by the same thread on consecutive steps (limiter pipeline in
//     registers); the nine contributions of a cell meet in the same march step.
// Unsplit: the x-sweep writes qout = qin + (its 9 terms), the y- and z-sweeps read-modify-
// write qout, so every cell receives its 27 addends in the reference order.
// Dimension-split (method(3) < 0): same kernels without the exchange; each pass writes the
// whole array (cells outside step3ds's range are copied through), ping-ponging two buffers.
//
// Array coordinates (I,J,K) = Fortran (i+1, j+1, k+1); interior cells [2, n+2); interface I
// separates cells I-1 and I.
#pragma once
#include "rp3.cuh"
#include "sweep2.cuh"

namespace clawcuda {

#ifndef CLAWCUDA_SWEEP3_PARAMS
#define CLAWCUDA_SWEEP3_PARAMS
struct Sweep3Params {
  const double* qin;        // SoA planes: component m at + m*plane, (I,J,K) at K*kstride + J*pitch + I
  double* qout;
  const double* aux;
  unsigned long long* cfl_bits;
  int n[3];                 // interior cells
  int pitch;
  long long kstride;        // pitch * (my+4)
  long long plane;          // kstride * (mz+4)
  double dtd[3];            // dt/dx, dt/dy, dt/dz
  int order;                // method(2)
  int m3, m4;               // flux3.f90:186-194 (m3 = -1: dimensional splitting)
  int limit;
  int mthlim[8];
  int first_pass;           // unsplit: 1 -> qout = qin + terms (x-sweep), 0 -> qout += terms
  int seg;                  // cells per march segment (blockIdx.z)
  int zlo, zhi;             // pencil planes K swept by the x- and y-sweeps: [1, mz+2], widened by one at
                            // rank-to-rank faces of a slab decomposition when dimension-split
  RpParams rp;
  // device-resident step control (null: qin / qout / dtd above are final); see StepCtl in sweep2.cuh
  const StepCtl* ctl;
  double* qbuf[3];          // ping-pong q buffers + the third buffer of the dimension-split step
  int src_role, dst_role;   // 0 = current, 1 = alternate, 2 = third buffer
  double delta[3];          // dx, dy, dz: dtd = ctl->dt / delta as step3.f90:106-108 computes it
  int capa;                 // method(6) = index_capa + 1: 1-based aux plane of the capacity function, 0 = none
                            // (last member: the offsets of the others are those of the profiled builds)
};
#endif

#define CLAW_RESOLVE_STEP3(P, qin_, qout_, d0_, d1_, d2_)                                  \
  const double* qin_ = (P).qin;                                                           \
  double* qout_ = (P).qout;                                                               \
  double d0_ = (P).dtd[0], d1_ = (P).dtd[1], d2_ = (P).dtd[2];                            \
  if ((P).ctl != nullptr) {                                                               \
    if ((P).ctl->done) return;                                                            \
    qin_ = ctl_buffer((P).ctl, (P).qbuf, (P).src_role);                                   \
    qout_ = ctl_buffer((P).ctl, (P).qbuf, (P).dst_role);                                  \
    const double dt_ = (P).ctl->dt;                                                       \
    d0_ = dt_ / (P).delta[0]; d1_ = dt_ / (P).delta[1]; d2_ = dt_ / (P).delta[2];         \
  }

namespace CLAW_NS {

template <int M, int W>
__device__ __forceinline__ void wave_dots3(const double (*wave)[M], const double (*waveR)[M], double* wnorm2,
                                           double* dotr) {
  // limiter.f90:33-38
#pragma unroll
  for (int mw = 0; mw < W; ++mw) {
    double wn = 0.0, dr = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      wn = (m == 0) ? from0(wave[mw][m] * wave[mw][m]) : wn + wave[mw][m] * wave[mw][m];
      dr = (m == 0) ? from0(wave[mw][m] * waveR[mw][m]) : dr + wave[mw][m] * waveR[mw][m];
    }
    wnorm2[mw] = wn; dotr[mw] = dr;
  }
}

template <int M, int W, int LIM = 0>      // LIM > 0: every wave family uses limiter id LIM (known at compile time)
__device__ __forceinline__ void limit_waves3(const int* mthlim, double (*wave)[M], const double* s,
                                             const double* wnorm2, const double* dotl, const double* dotr) {
  // limiter.f90:39-50
#pragma unroll
  for (int mw = 0; mw < W; ++mw) {
    const int meth = LIM ? LIM : mthlim[mw];
    if (meth == 0) continue;
    if (wave_is_zero(wnorm2[mw])) continue;
    const double wl = philim(wnorm2[mw], (s[mw] > 0.0) ? dotl[mw] : dotr[mw], meth);
#pragma unroll
    for (int m = 0; m < M; ++m) wave[mw][m] = wl * wave[mw][m];
  }
}

// flux3.f90:231-259 (use_fwave = .false.): cqxx (= the fadd increment) of one interface
// dtd_l / dtd_r = dtdx1d of the cells left / right of the interface (equal without a capacity function)
template <int M, int W>
__device__ __forceinline__ void second_order3(double dtd_l, double dtd_r, const double (*wave)[M], const double* s,
                                              double* cqxx) {
  const double dtdxave = 0.5 * (dtd_l + dtd_r);                                         // flux3.f90:236
#pragma unroll
  for (int m = 0; m < M; ++m) {
    double c = 0.0;
#pragma unroll
    for (int mw = 0; mw < W; ++mw) {
      const double t = 0.5 * fabs(s[mw]) * (1.0 - fabs(s[mw]) * dtdxave) * wave[mw][m];
      c = (mw == 0) ? from0(t) : c + t;
    }
    cqxx[m] = c;
  }
}

// -------------------------------------------------------------------------------------------
// One side of cell c's transverse work: SIDE 0 = the p-side of interface c (asdq = apdq, imp = 2),
// SIDE 1 = the m-side of interface c+1 (asdq = amdq, imp = 1).  Accumulates into
// g[e][m], h[e][m] with e = (slice+1)*2 + (side-1)  <->  gadd/hadd(m, side, slice, c).
// A[dy][dz][ma] = aux of the cell column at offsets (dy, dz) in the (y-like, z-like) directions
// = aux{1,2,3}(ma, i1, 1..3) of flux3 (step3.f90:144-150).
// Statement order follows flux3.f90:272-291 (rpt3), 296-344 (cqxx splits), 353-384 (c*2, rptt3),
// 390-469 (gadd), 479-507 (b*, rptt3), 513-594 (hadd).
// -------------------------------------------------------------------------------------------
// M3C / M4C >= 0: method(3) = 10*M3C + M4C known at compile time (the run-time arguments are then ignored)
template <class RP, int IXYZ, int SIDE, int M3C = -1, int M4C = -1>
__device__ __forceinline__ void flux3_side(const RpParams& P, int m3_rt, int m4_rt, double dtdx, double dty, double dtz,
                                           const double* asdq, const double* cqxx, const double* ctx,
                                           const double (*A)[3][RP::MAUX > 0 ? RP::MAUX : 1],
                                           double (*g)[RP::template TV<IXYZ, 2>::N],
                                           double (*h)[RP::template TV<IXYZ, 3>::N]) {
  constexpr int M = RP::MEQN;
  const int m3 = (M3C >= 0) ? M3C : m3_rt, m4 = (M4C >= 0) ? M4C : m4_rt;
  using TY = typename RP::template TV<IXYZ, 2>;            // non-zero components of y-like splits (G terms)
  using TZ = typename RP::template TV<IXYZ, 3>;            // ... of z-like splits (H terms)
  constexpr double SGN = (SIDE == 0) ? -1.0 : 1.0;        // cqxx enters the p-side with -, the m-side with +
  const double hd = 0.5 * dtdx;
  const double sixth = 1.0 / 6.0;
  const double sdz = sixth * dtdx * dtz;                   // (1/6)*dtdx1d*dtdz   (flux3.f90:412)
  const double sdy = sixth * dtdx * dty;                   // (1/6)*dtdx1d*dtdy   (flux3.f90:537)
  // ---- z-like splits (rpt3 icoor = 3) ---------------------------------------------------
  double cm[M], cp[M], cmq[M], cpq[M];
  RP::template transverse<IXYZ, 3>(P, ctx, A[1][0], A[1][1], A[1][2], asdq, cm, cp);
#pragma unroll
  for (int m = 0; m < M; ++m) { cmq[m] = 0.0; cpq[m] = 0.0; }
  if (m3 == 2) RP::template transverse<IXYZ, 3>(P, ctx, A[1][0], A[1][1], A[1][2], cqxx, cmq, cpq);
  // ---- y-like splits (rpt3 icoor = 2) ---------------------------------------------------
  double bm[M], bp[M], bmq[M], bpq[M];
  RP::template transverse<IXYZ, 2>(P, ctx, A[0][1], A[1][1], A[2][1], asdq, bm, bp);
#pragma unroll
  for (int m = 0; m < M; ++m) { bmq[m] = 0.0; bpq[m] = 0.0; }
  if (m3 == 2) RP::template transverse<IXYZ, 2>(P, ctx, A[0][1], A[1][1], A[2][1], cqxx, bmq, bpq);
  // ---- G: y-like fluxes ---------------------------------------------------------------------
  {
    double bmcp[M], bpcp[M], bmcm[M], bpcm[M];
    if (m4 > 0) {
      double cp2[M], cm2[M];
#pragma unroll
      for (int m = 0; m < M; ++m) { cp2[m] = 0.0; cm2[m] = 0.0; }
#pragma unroll
      for (int k = 0; k < TZ::N; ++k) {
        const int m = TZ::comp(k);
        if (m4 == 2) { cp2[m] = cp[m] + SGN * (3.0 * cpq[m]); cm2[m] = cm[m] + SGN * (3.0 * cmq[m]); }
        else { cp2[m] = cp[m]; cm2[m] = cm[m]; }
      }
      RP::template transverse<IXYZ, 2, true>(P, ctx, A[0][2], A[1][2], A[2][2], cp2, bmcp, bpcp);   // impt = 2
      RP::template transverse<IXYZ, 2, true>(P, ctx, A[0][0], A[1][0], A[2][0], cm2, bmcm, bpcm);   // impt = 1
    }
#pragma unroll
    for (int k = 0; k < TY::N; ++k) {
      const int m = TY::comp(k);
      g[2][k] = g[2][k] - hd * bm[m];                                  // gadd(m,1,0,.)
      g[3][k] = g[3][k] - hd * bp[m];                                  // gadd(m,2,0,.)
      if (m4 > 0) {
        g[3][k] = g[3][k] + sdz * (bpcp[m] - bpcm[m]);
        g[2][k] = g[2][k] + sdz * (bmcp[m] - bmcm[m]);
        g[5][k] = g[5][k] - sdz * bpcp[m];                             // gadd(m,2,1,.)
        g[4][k] = g[4][k] - sdz * bmcp[m];                             // gadd(m,1,1,.)
        g[1][k] = g[1][k] + sdz * bpcm[m];                             // gadd(m,2,-1,.)
        g[0][k] = g[0][k] + sdz * bmcm[m];                             // gadd(m,1,-1,.)
      }
      if (m3 >= 2) {
        if (SIDE == 0) { g[3][k] = g[3][k] + dtdx * bpq[m]; g[2][k] = g[2][k] + dtdx * bmq[m]; }
        else           { g[3][k] = g[3][k] - dtdx * bpq[m]; g[2][k] = g[2][k] - dtdx * bmq[m]; }
      }
    }
  }
  // ---- H: z-like fluxes ---------------------------------------------------------------------
  {
    double bmcp[M], bpcp[M], bmcm[M], bpcm[M];
    if (m4 > 0) {
      double bp2[M], bm2[M];
#pragma unroll
      for (int m = 0; m < M; ++m) { bp2[m] = 0.0; bm2[m] = 0.0; }
#pragma unroll
      for (int k = 0; k < TY::N; ++k) {
        const int m = TY::comp(k);
        if (m4 == 2) { bp2[m] = bp[m] + SGN * (3.0 * bpq[m]); bm2[m] = bm[m] + SGN * (3.0 * bmq[m]); }
        else { bp2[m] = bp[m]; bm2[m] = bm[m]; }
      }
      RP::template transverse<IXYZ, 3, true>(P, ctx, A[2][0], A[2][1], A[2][2], bp2, bmcp, bpcp);   // impt = 2
      RP::template transverse<IXYZ, 3, true>(P, ctx, A[0][0], A[0][1], A[0][2], bm2, bmcm, bpcm);   // impt = 1
    }
#pragma unroll
    for (int k = 0; k < TZ::N; ++k) {
      const int m = TZ::comp(k);
      h[2][k] = h[2][k] - hd * cm[m];
      h[3][k] = h[3][k] - hd * cp[m];
      if (m4 > 0) {
        h[3][k] = h[3][k] + sdy * (bpcp[m] - bpcm[m]);
        h[2][k] = h[2][k] + sdy * (bmcp[m] - bmcm[m]);
        h[5][k] = h[5][k] - sdy * bpcp[m];
        h[4][k] = h[4][k] - sdy * bmcp[m];
        h[1][k] = h[1][k] + sdy * bpcm[m];
        h[0][k] = h[0][k] + sdy * bmcm[m];
      }
      if (m3 >= 2) {
        if (SIDE == 0) { h[3][k] = h[3][k] + dtdx * cpq[m]; h[2][k] = h[2][k] + dtdx * cmq[m]; }
        else           { h[3][k] = h[3][k] - dtdx * cpq[m]; h[2][k] = h[2][k] - dtdx * cmq[m]; }
      }
    }
  }
}

// shared-memory image of a cell's gadd/hadd: g entry e, k-th non-zero component at sm[(e*NG + k)*NT],
// h at sm[(6*NG + e*NH + k)*NT]
template <class RP, int IXYZ> struct GH3 {
  static constexpr int NG = RP::template TV<IXYZ, 2>::N, NH = RP::template TV<IXYZ, 3>::N;
  static constexpr int PER_THREAD = 6 * (NG + NH);
};

// Both sides of cell c, results to shared memory
template <class RP, int IXYZ, int NT, int M3C = -1, int M4C = -1>
__device__ __forceinline__ void flux3_cell(const Sweep3Params& P, double dtdx, double dty, double dtz,
                                           const double* apdqC, const double* cqxxC, const double* ctxC,
                                           const double* amdqR, const double* cqxxR, const double* ctxR,
                                           const double (*A)[3][RP::MAUX > 0 ? RP::MAUX : 1], double* sm) {
  constexpr int NG = GH3<RP, IXYZ>::NG, NH = GH3<RP, IXYZ>::NH;
  double g[6][NG], h[6][NH];
#pragma unroll
  for (int e = 0; e < 6; ++e) {
#pragma unroll
    for (int k = 0; k < NG; ++k) g[e][k] = 0.0;
#pragma unroll
    for (int k = 0; k < NH; ++k) h[e][k] = 0.0;
  }
  flux3_side<RP, IXYZ, 0, M3C, M4C>(P.rp, P.m3, P.m4, dtdx, dty, dtz, apdqC, cqxxC, ctxC, A, g, h);   // iteration i = c
  flux3_side<RP, IXYZ, 1, M3C, M4C>(P.rp, P.m3, P.m4, dtdx, dty, dtz, amdqR, cqxxR, ctxR, A, g, h);   // iteration i = c+1
#pragma unroll
  for (int e = 0; e < 6; ++e) {
#pragma unroll
    for (int k = 0; k < NG; ++k) sm[(e * NG + k) * NT] = g[e][k];
#pragma unroll
    for (int k = 0; k < NH; ++k) sm[(6 * NG + e * NH + k) * NT] = h[e][k];
  }
}

// The (a, b) target term of a source pencil (step3.f90:198-225): a / b = offset of the target
// pencil in the y-like / z-like direction.  sm = shared-memory slot of the SOURCE thread.
// G terms touch the components TY::comp(k), H terms TZ::comp(k); per component the g term is
// added before the h term, as the reference's expression does.
template <class RP, int IXYZ, int A, int B, int NT>
__device__ __forceinline__ void apply_target(double* acc, const double* sm, double dty, double dtz) {
  using TY = typename RP::template TV<IXYZ, 2>;
  using TZ = typename RP::template TV<IXYZ, 3>;
  constexpr int NG = TY::N, NH = TZ::N;
#define G_(side, slice) sm[((((slice) + 1) * 2 + (side)-1) * NG + k) * NT]
#define H_(side, slice) sm[(6 * NG + (((slice) + 1) * 2 + (side)-1) * NH + k) * NT]
  if (NG == RP::MEQN && NH == RP::MEQN) {
    // dense: interleave per component exactly like the reference statement
#pragma unroll
    for (int k = 0; k < RP::MEQN; ++k) {
      double r = acc[k];
      if (A == -1 && B == 0) { r = r - dty * G_(1, 0); r = r - dtz * (H_(2, -1) - H_(1, -1)); }
      if (A == -1 && B == -1) { r = r - dty * G_(1, -1); r = r - dtz * H_(1, -1); }
      if (A == 0 && B == -1) { r = r - dty * (G_(2, -1) - G_(1, -1)); r = r - dtz * H_(1, 0); }
      if (A == 1 && B == -1) { r = r + dty * G_(2, -1); r = r - dtz * H_(1, 1); }
      if (A == 1 && B == 0) { r = r + dty * G_(2, 0); r = r - dtz * (H_(2, 1) - H_(1, 1)); }
      if (A == 1 && B == 1) { r = r + dty * G_(2, 1); r = r + dtz * H_(2, 1); }
      if (A == 0 && B == 1) { r = r - dty * (G_(2, 1) - G_(1, 1)); r = r + dtz * H_(2, 0); }
      if (A == -1 && B == 1) { r = r - dty * G_(1, 1); r = r + dtz * H_(2, -1); }
      acc[k] = r;
    }
  } else {
    // sparse: the g term of every touched component, then the h term (for a component touched by both,
    // e.g. the pressure, this is still g before h)
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const int m = TY::comp(k);
      double r = acc[m];
      if (A == -1 && B == 0) r = r - dty * G_(1, 0);
      if (A == -1 && B == -1) r = r - dty * G_(1, -1);
      if (A == 0 && B == -1) r = r - dty * (G_(2, -1) - G_(1, -1));
      if (A == 1 && B == -1) r = r + dty * G_(2, -1);
      if (A == 1 && B == 0) r = r + dty * G_(2, 0);
      if (A == 1 && B == 1) r = r + dty * G_(2, 1);
      if (A == 0 && B == 1) r = r - dty * (G_(2, 1) - G_(1, 1));
      if (A == -1 && B == 1) r = r - dty * G_(1, 1);
      acc[m] = r;
    }
#pragma unroll
    for (int k = 0; k < NH; ++k) {
      const int m = TZ::comp(k);
      double r = acc[m];
      if (A == -1 && B == 0) r = r - dtz * (H_(2, -1) - H_(1, -1));
      if (A == -1 && B == -1) r = r - dtz * H_(1, -1);
      if (A == 0 && B == -1) r = r - dtz * H_(1, 0);
      if (A == 1 && B == -1) r = r - dtz * H_(1, 1);
      if (A == 1 && B == 0) r = r - dtz * (H_(2, 1) - H_(1, 1));
      if (A == 1 && B == 1) r = r + dtz * H_(2, 1);
      if (A == 0 && B == 1) r = r + dtz * H_(2, 0);
      if (A == -1 && B == 1) r = r + dtz * H_(2, -1);
      acc[m] = r;
    }
  }
}
// the pencil's own cell (step3.f90:194-197)
template <class RP, int IXYZ, int NT>
__device__ __forceinline__ void apply_center(double* acc, const double* sm, const double* qadd, const double* fdiff,
                                             double dtdn, double dty, double dtz, bool transverse) {
  using TY = typename RP::template TV<IXYZ, 2>;
  using TZ = typename RP::template TV<IXYZ, 3>;
  constexpr int NG = TY::N, NH = TZ::N;
#pragma unroll
  for (int m = 0; m < RP::MEQN; ++m) acc[m] = (acc[m] + qadd[m]) - dtdn * fdiff[m];
  if (transverse) {
    if (NG == RP::MEQN && NH == RP::MEQN) {
#pragma unroll
      for (int k = 0; k < RP::MEQN; ++k) {
        double r = acc[k];
        r = r - dty * (G_(2, 0) - G_(1, 0));
        r = r - dtz * (H_(2, 0) - H_(1, 0));
        acc[k] = r;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NG; ++k) { const int m = TY::comp(k); acc[m] = acc[m] - dty * (G_(2, 0) - G_(1, 0)); }
#pragma unroll
      for (int k = 0; k < NH; ++k) { const int m = TZ::comp(k); acc[m] = acc[m] - dtz * (H_(2, 0) - H_(1, 0)); }
    }
  }
}
// ---- capacity function (step3.f90:229-273 / 384-431 / 546-593) -------------------------------------------------------
// Every bracketed flux sum is divided by the capacity of the cell it goes into: one addend per source pencil,
//   acc = acc + sgn * (dty * Gterm + s2 * dtz * Hterm) / capa      (the reference's nine expressions in the (a, b) frame)
// instead of the two sequential addends of the plain form.  Components a transverse solve leaves structurally zero
// (fast mode stores only TV::N of them) enter as exact zeros, which is what the reference adds.
template <class RP, int IXYZ, int A, int B, int NT>
__device__ __forceinline__ void apply_target_capa(double* acc, const double* sm, double dty, double dtz, double capa) {
  using TY = typename RP::template TV<IXYZ, 2>;
  using TZ = typename RP::template TV<IXYZ, 3>;
  constexpr int NG = TY::N, NH = TZ::N, M = RP::MEQN;
  double gt[M], ht[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { gt[m] = 0.0; ht[m] = 0.0; }
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const int m = TY::comp(k);
    if (A == -1 && B == 0) gt[m] = G_(1, 0);
    if (A == -1 && B == -1) gt[m] = G_(1, -1);
    if (A == 0 && B == -1) gt[m] = G_(2, -1) - G_(1, -1);
    if (A == 1 && B == -1) gt[m] = G_(2, -1);
    if (A == 1 && B == 0) gt[m] = G_(2, 0);
    if (A == 1 && B == 1) gt[m] = G_(2, 1);
    if (A == 0 && B == 1) gt[m] = G_(2, 1) - G_(1, 1);
    if (A == -1 && B == 1) gt[m] = G_(1, 1);
  }
#pragma unroll
  for (int k = 0; k < NH; ++k) {
    const int m = TZ::comp(k);
    if (A == -1 && B == 0) ht[m] = H_(2, -1) - H_(1, -1);
    if (A == -1 && B == -1) ht[m] = H_(1, -1);
    if (A == 0 && B == -1) ht[m] = H_(1, 0);
    if (A == 1 && B == -1) ht[m] = H_(1, 1);
    if (A == 1 && B == 0) ht[m] = H_(2, 1) - H_(1, 1);
    if (A == 1 && B == 1) ht[m] = H_(2, 1);
    if (A == 0 && B == 1) ht[m] = H_(2, 0);
    if (A == -1 && B == 1) ht[m] = H_(2, -1);
  }
  constexpr bool plus_outer = (A == 1);                                        // targets (1, b): "+ (...)", else "- (...)"
  constexpr bool minus_inner = (A == 1 && B <= 0) || (A <= 0 && B == 1);       // "g - h" inside the bracket, else "g + h"
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const double g = dty * gt[m], h = dtz * ht[m];
    const double t = ddiv(minus_inner ? g - h : g + h, capa);
    acc[m] = plus_outer ? acc[m] + t : acc[m] - t;
  }
}
// the pencil's own cell (step3.f90:231-235): dtdn is the SCALAR dt/dx of the sweep direction
template <class RP, int IXYZ, int NT>
__device__ __forceinline__ void apply_center_capa(double* acc, const double* sm, const double* qadd, const double* fdiff,
                                                  double dtdn, double dty, double dtz, bool transverse, double capa) {
  using TY = typename RP::template TV<IXYZ, 2>;
  using TZ = typename RP::template TV<IXYZ, 3>;
  constexpr int NG = TY::N, NH = TZ::N, M = RP::MEQN;
  double gt[M], ht[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { gt[m] = 0.0; ht[m] = 0.0; }
  if (transverse) {
#pragma unroll
    for (int k = 0; k < NG; ++k) gt[TY::comp(k)] = G_(2, 0) - G_(1, 0);
#pragma unroll
    for (int k = 0; k < NH; ++k) ht[TZ::comp(k)] = H_(2, 0) - H_(1, 0);
  }
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = (acc[m] + qadd[m]) - ddiv((dtdn * fdiff[m] + dty * gt[m]) + dtz * ht[m], capa);
}
// flux3.f90:211-218 with dtdx1d of the two cells of the interface
template <int W>
__device__ __forceinline__ double cfl_update_capa(double cfl, double dtd_l, double dtd_r, const double* s) {
#pragma unroll
  for (int mw = 0; mw < W; ++mw) cfl = fmax2(fmax2(cfl, dtd_r * s[mw]), -dtd_l * s[mw]);
  return cfl;
}
#undef G_
#undef H_

template <int NT> __device__ __forceinline__ void block_max_cfl3(double cfl, unsigned long long* addr, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cfl = fmax(cfl, __shfl_xor_sync(0xffffffffu, cfl, o));
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  __syncthreads();
  if (lane == 0) s_red[warp] = cfl;
  __syncthreads();
  if (warp == 0) {
    double c = (lane < NT / 32) ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
    if (lane == 0 && c > 0.0) atomic_max_cfl(addr, c);
  }
}

// 8-byte asynchronous global -> shared copy (LDGSTS): prefetches the next march step's operands without
// holding registers; the issuing thread reads its own slots after cp_async_wait_all()
#ifdef CLAW_HOST_CHECK                      // kernel emulation (tests/host_functors/cuda_emul.h): the copy completes at once
inline void cp_async8(double* smem_dst, const double* gsrc) { *smem_dst = *gsrc; }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#else
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <class RP, int IXYZ, int NT> struct Smem3 {
  static constexpr int PER_THREAD = GH3<RP, IXYZ>::PER_THREAD;
  // march kernels: two stages of per-thread prefetch slots: q(n) [M], aux(n) [MA], the 3x3 aux neighbourhood of
  // cell n-2 [9 MA] and the old qout(n-2) [M]
  static constexpr int MA = RP::MAUX > 0 ? RP::MAUX : 0;
  // x-sweep: ring of 4 K-plane slots x 3 rows x MA aux values per thread (slot = plane & 3)
  static constexpr int PF = (IXYZ == 0) ? (6 * MA) : (2 * RP::MEQN + 10 * MA);
  static constexpr int OFF_RED = PER_THREAD * NT;
  static constexpr int OFF_PF = OFF_RED + 64;
  static constexpr size_t BYTES = (size_t)(OFF_PF + 2 * PF * NT) * sizeof(double);
};

// =====================================================================================
// x-sweep: block (32, BY); lanes <-> I (28 owner lanes + 2 halo lanes per side), thread-y <-> J,
// march in K.   step3.f90:124-276 (unsplit) / step3ds.f90 idir = 1.
// =====================================================================================
// SPEC = 10*LIM + 1: the headline method set of BASELINE config 4 known at compile time -- unsplit, order 2,
// method(3) = 22, limiter id LIM on every wave family -- so that the run-time method tests, the zero-initialised
// dummies of the unused branches and the limiter switch disappear (what SPEC does for sweep2_unsplit_ws).
template <class RP, int BY, int MINB = 1, int SPEC = 0>
__global__ void __launch_bounds__(32 * BY, MINB) sweep3_x(const __grid_constant__ Sweep3Params P) {
  constexpr bool SP = (SPEC & 1) != 0;
  constexpr int LIM = SP ? SPEC / 10 : 0, M3C = SP ? 2 : -1, M4C = SP ? 2 : -1;
  constexpr int M = RP::MEQN, W = RP::MWAVES, MA = RP::MAUX > 0 ? RP::MAUX : 1, NC = RP::NCTX > 0 ? RP::NCTX : 1;
  constexpr int NT = 32 * BY;
  constexpr int TX = 28;
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + t;
  double* const sm = smem + tid;
  double* const sred = smem + Smem3<RP, 0, NT>::OFF_RED;

  const bool split = SP ? false : (P.m3 < 0);
  const int nx = P.n[0] + 4, ny = P.n[1] + 4, nz = P.n[2] + 4;
  const int I = blockIdx.x * TX + t;
  const int Ic = I < nx ? I : nx - 1;
  const int HT = split ? 0 : 1;                              // transverse halo of the tile
  const int J = (split ? 0 : 1) + blockIdx.y * (BY - 2 * HT) + ty;
  const int Jc = J < ny ? J : ny - 1;
  const int Jm = Jc > 0 ? Jc - 1 : 0, Jp = Jc < ny - 1 ? Jc + 1 : ny - 1;
  // output planes [kb, ke)
  const int k_lo = split ? 0 : 2, k_hi = split ? nz : P.n[2] + 2;
  const int kb = k_lo + blockIdx.z * P.seg;
  const int ke = min(kb + P.seg, k_hi);
  const int k0 = split ? kb : kb - 1, k1 = split ? ke - 1 : ke;     // pencil planes k0..k1

  CLAW_RESOLVE_STEP3(P, qin, qout, dtdx, dtdy, dtdz);
  const bool second = SP ? true : (P.order != 1);
  const bool limit = SP ? true : (second && (P.limit != 0));
  const bool transverse = SP ? true : (P.m3 > 0);
  // capacity function (method(6)): generic kernels of aux-carrying solvers only -- the specialised BASELINE kernels
  // compile the branch out
  const bool capa = !SP && P.capa > 0;
  const int icapa = capa ? P.capa - 1 : 0;

  const bool owner_lane = (t >= 2 && t <= 29 && I < nx) || (blockIdx.x == 0 && t < 2);
  const bool interior_i = (I >= 2 && I <= P.n[0] + 1);
  const bool row_pencil = (J >= 1 && J <= P.n[1] + 2);       // Fortran j in 0..my+1
  const bool row_out = split ? (J < ny) : (ty >= 1 && ty <= BY - 2 && J >= 2 && J <= P.n[1] + 1);
  const bool cfl_lane = (t >= 1) && (I >= 2) && (I <= P.n[0] + 2) && row_pencil;

  const double* qbase = qin + (long long)Jc * P.pitch + Ic;
  const double* abase = P.aux + Ic;
  double q[M], qnext[M];
  double acc0[M], acc1[M], acc2[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    qnext[m] = qbase[m * P.plane + (long long)k0 * P.kstride];
    acc0[m] = 0.0; acc1[m] = qnext[m]; acc2[m] = 0.0;
  }
  // aux(I, J-1..J+1, plane p) lives in my shared-memory ring slot p & 3, fetched asynchronously two planes ahead
  using SM3 = Smem3<RP, 0, NT>;
  double* const ring = smem + SM3::OFF_PF + tid;
  auto fetch_plane = [&](int p) {
    if (RP::MAUX > 0) {
      const int pc = p < 0 ? 0 : (p < nz ? p : nz - 1);
      const int jr[3] = {Jm, Jc, Jp};
      double* slot = ring + (size_t)((p & 3) * 3 * SM3::MA) * NT;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int ma = 0; ma < SM3::MA; ++ma)
          cp_async8(slot + (dy * SM3::MA + ma) * NT,
                    abase + ma * P.plane + (long long)pc * P.kstride + (long long)jr[dy] * P.pitch);
    }
    cp_async_commit();
  };
  fetch_plane(k0 - 1);
  fetch_plane(k0);
  fetch_plane(k0 + 1);
  double cfl = 0.0;

  for (int k = k0; k <= k1; ++k) {
    // ---- rotate: pencil plane k; prefetch plane k+1 -------------------------------------
    const int kn = k + 1 < nz ? k + 1 : nz - 1;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      q[m] = qnext[m];
      qnext[m] = qbase[m * P.plane + (long long)kn * P.kstride];
      acc2[m] = qnext[m];
    }
    cp_async_wait_all();                                     // planes k-1, k, k+1 are in my ring
    fetch_plane(k + 2);
    double A[3][3][MA];
    if (RP::MAUX > 0) {
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
        const double* slot = ring + (size_t)(((k - 1 + dz) & 3) * 3 * SM3::MA) * NT;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int ma = 0; ma < MA; ++ma) A[dy][dz][ma] = slot[(dy * SM3::MA + ma) * NT];
      }
    }
    const bool plane_pencil = (k >= P.zlo && k <= P.zhi);    // Fortran k in 0..mz+1
    // ---- interface I (my left face) ------------------------------------------------------
    double ql[M], auxl[MA];
#pragma unroll
    for (int m = 0; m < M; ++m) ql[m] = shfl_up1(q[m]);
#pragma unroll
    for (int ma = 0; ma < MA; ++ma) auxl[ma] = (RP::MAUX > 0) ? shfl_up1(A[1][1][ma]) : 0.0;
    double wave[W][M], s[W], amdq[M], apdq[M], ctx[NC], cqxx[M];
    RP::template normal<0>(P.rp, ql, q, auxl, A[1][1], wave, s, amdq, apdq, ctx);      // flux3.f90:201
    // capacity of my cell in the planes k-1, k, k+1 (the capacity plane need not be one the Riemann solver reads, so it
    // is fetched on its own; this path is not a throughput target)
    double c0 = 1.0, c1 = 1.0, c2 = 1.0;
    if (!SP && capa) {
      const double* cp = P.aux + (long long)icapa * P.plane + (long long)Jc * P.pitch + Ic;
      c0 = cp[(long long)(k > 0 ? k - 1 : 0) * P.kstride];
      c1 = cp[(long long)k * P.kstride];
      c2 = cp[(long long)kn * P.kstride];
    }
    // dtdx1d of my cell and of the cell on my left (step3.f90:132-136); without a capacity function both are dt/dx
    const double dI = (!SP && capa) ? ddiv(dtdx, c1) : dtdx;
    const double dIm1 = (!SP && capa) ? shfl_up1(dI) : dtdx;
    if (cfl_lane && plane_pencil) {                                                     // flux3.f90:211-218
        cfl = (!SP && capa) ? cfl_update_capa<W>(cfl, dIm1, dI, s) : cfl_update<W>(cfl, dtdx, s);
    }
#pragma unroll
    for (int m = 0; m < M; ++m) cqxx[m] = 0.0;
    if (second) {
      if (limit) {                                                                      // flux3.f90:228
        double wnorm2[W], dotr[W], dotl[W];
        {
          double waveR[W][M];
#pragma unroll
          for (int mw = 0; mw < W; ++mw)
#pragma unroll
            for (int m = 0; m < M; ++m) waveR[mw][m] = shfl_dn1(wave[mw][m]);
          wave_dots3<M, W>(wave, waveR, wnorm2, dotr);
        }
#pragma unroll
        for (int mw = 0; mw < W; ++mw) dotl[mw] = shfl_up1(dotr[mw]);
        limit_waves3<M, W, LIM>(P.mthlim, wave, s, wnorm2, dotl, dotr);
      }
      second_order3<M, W>(dIm1, dI, wave, s, cqxx);                                     // flux3.f90:231-259
    }
    // ---- the next interface's left-going pieces (cell I owns the m-side of interface I+1) --
    double amdqR[M], cqxxR[M], ctxR[NC];
#pragma unroll
    for (int m = 0; m < M; ++m) { amdqR[m] = shfl_dn1(amdq[m]); cqxxR[m] = shfl_dn1(cqxx[m]); }
#pragma unroll
    for (int c = 0; c < NC; ++c) ctxR[c] = (RP::NCTX > 0) ? shfl_dn1(ctx[c]) : 0.0;
    double qadd[M], fdiff[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      qadd[m] = (0.0 - dI * apdq[m]) - dI * amdqR[m];                                  // flux3.f90:205-208
      fdiff[m] = (0.0 + cqxxR[m]) - (0.0 + cqxx[m]);                                    // fadd(i+1) - fadd(i)
    }
    if (split) {
      // step3ds.f90:195-198: only the pencil itself is updated; everything else is copied through
      const bool upd = interior_i && row_pencil && plane_pencil;
      if (owner_lane && row_out && k >= kb && k < ke) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const double r = (!SP && capa) ? (q[m] + qadd[m]) - ddiv(dtdx * fdiff[m], c1)         // step3ds.f90:200-205
                                : (q[m] + qadd[m]) - dtdx * fdiff[m];
          qout[m * P.plane + (long long)k * P.kstride + (long long)J * P.pitch + I] = upd ? r : q[m];
        }
      }
      continue;
    }
    if (transverse) flux3_cell<RP, 0, NT, M3C, M4C>(P, dI, dtdy, dtdz, apdq, cqxx, ctx, amdqR, cqxxR, ctxR, A, sm);
    __syncthreads();
    if (!SP && row_out && capa) {
      if (transverse) {
        apply_target_capa<RP, 0, 1, -1, NT>(acc0, sm - 32, dtdy, dtdz, c0);
        apply_target_capa<RP, 0, 0, -1, NT>(acc0, sm, dtdy, dtdz, c0);
        apply_target_capa<RP, 0, -1, -1, NT>(acc0, sm + 32, dtdy, dtdz, c0);
        apply_target_capa<RP, 0, 1, 0, NT>(acc1, sm - 32, dtdy, dtdz, c1);
        apply_center_capa<RP, 0, NT>(acc1, sm, qadd, fdiff, dtdx, dtdy, dtdz, true, c1);
        apply_target_capa<RP, 0, -1, 0, NT>(acc1, sm + 32, dtdy, dtdz, c1);
        apply_target_capa<RP, 0, 1, 1, NT>(acc2, sm - 32, dtdy, dtdz, c2);
        apply_target_capa<RP, 0, 0, 1, NT>(acc2, sm, dtdy, dtdz, c2);
        apply_target_capa<RP, 0, -1, 1, NT>(acc2, sm + 32, dtdy, dtdz, c2);
      } else {
        apply_center_capa<RP, 0, NT>(acc1, sm, qadd, fdiff, dtdx, dtdy, dtdz, false, c1);
      }
      if (owner_lane && interior_i && k - 1 >= kb && k - 1 < ke) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          qout[m * P.plane + (long long)(k - 1) * P.kstride + (long long)J * P.pitch + I] = acc0[m];
      }
    } else if (row_out) {
      if (transverse) {
        // pencils (J-1, k), (J, k), (J+1, k) in this order (step3.f90:124-125: k outer, j inner);
        // a = J_target - J_source, b = plane_target - k
        apply_target<RP, 0, 1, -1, NT>(acc0, sm - 32, dtdy, dtdz);
        apply_target<RP, 0, 0, -1, NT>(acc0, sm, dtdy, dtdz);
        apply_target<RP, 0, -1, -1, NT>(acc0, sm + 32, dtdy, dtdz);
        apply_target<RP, 0, 1, 0, NT>(acc1, sm - 32, dtdy, dtdz);
        apply_center<RP, 0, NT>(acc1, sm, qadd, fdiff, dtdx, dtdy, dtdz, true);
        apply_target<RP, 0, -1, 0, NT>(acc1, sm + 32, dtdy, dtdz);
        apply_target<RP, 0, 1, 1, NT>(acc2, sm - 32, dtdy, dtdz);
        apply_target<RP, 0, 0, 1, NT>(acc2, sm, dtdy, dtdz);
        apply_target<RP, 0, -1, 1, NT>(acc2, sm + 32, dtdy, dtdz);
      } else {
        apply_center<RP, 0, NT>(acc1, sm, qadd, fdiff, dtdx, dtdy, dtdz, false);
      }
      if (owner_lane && interior_i && k - 1 >= kb && k - 1 < ke) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          qout[m * P.plane + (long long)(k - 1) * P.kstride + (long long)J * P.pitch + I] = acc0[m];
      }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) { acc0[m] = acc1[m]; acc1[m] = acc2[m]; }
    __syncthreads();
  }
  block_max_cfl3<NT>(cfl, P.cfl_bits, sred);
}

// =====================================================================================
// y-sweep (IXYZ = 1: march J, thread-y <-> K) and z-sweep (IXYZ = 2: march K, thread-y <-> J);
// thread-x <-> I.   step3.f90:283-431 / 439-593, step3ds.f90 idir = 2 / 3.
// =====================================================================================
template <class RP, int IXYZ, int BX, int BY, int MINB = 1, int SPEC = 0>      // SPEC: see sweep3_x
__global__ void __launch_bounds__(BX * BY, MINB) sweep3_march(const __grid_constant__ Sweep3Params P) {
  constexpr bool SP = (SPEC & 1) != 0;
  constexpr int LIM = SP ? SPEC / 10 : 0, M3C = SP ? 2 : -1, M4C = SP ? 2 : -1;
  constexpr int M = RP::MEQN, W = RP::MWAVES, MA = RP::MAUX > 0 ? RP::MAUX : 1, NC = RP::NCTX > 0 ? RP::NCTX : 1;
  constexpr int NT = BX * BY;
  constexpr int TD = (IXYZ == 1) ? 2 : 1;                    // array dimension carried by thread-y
  CLAW_DYN_SMEM(smem);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * BX + tx;
  double* const sm = smem + tid;
  double* const sred = smem + Smem3<RP, IXYZ, NT>::OFF_RED;
  using SM3 = Smem3<RP, IXYZ, NT>;
  double* const spf = smem + SM3::OFF_PF + tid;               // my prefetch slots: stage s, slot j at spf[(s*PF + j)*NT]

  const bool split = SP ? false : (P.m3 < 0);
  const int nx = P.n[0] + 4, nT = P.n[TD] + 4, nlen = P.n[IXYZ] + 4, mlen = P.n[IXYZ];
  const int HT = split ? 0 : 1;
  const int I = (split ? 0 : 1) + blockIdx.x * (BX - 2 * HT) + tx;
  const int T = (split ? 0 : 1) + blockIdx.y * (BY - 2 * HT) + ty;
  const int Ic = I < nx ? I : nx - 1, Tc = T < nT ? T : nT - 1;
  const int Im = Ic > 0 ? Ic - 1 : 0, Ip = Ic < nx - 1 ? Ic + 1 : nx - 1;
  const int Tm = Tc > 0 ? Tc - 1 : 0, Tp = Tc < nT - 1 ? Tc + 1 : nT - 1;
  const long long strideT = (IXYZ == 1) ? P.kstride : (long long)P.pitch;     // thread-y stride
  const long long strideN = (IXYZ == 1) ? (long long)P.pitch : P.kstride;     // march stride
  // y-sweep: y-like = z (thread-y), z-like = x (thread-x).  z-sweep: y-like = x, z-like = y.
  CLAW_RESOLVE_STEP3(P, qin, qout, dtd0, dtd1, dtd2);
  const double dtdn = (IXYZ == 1) ? dtd1 : dtd2;
  const double dty = (IXYZ == 1) ? dtd2 : dtd0;
  const double dtz = (IXYZ == 1) ? dtd0 : dtd1;

  const int c_lo = split ? 0 : 2, c_hi = split ? nlen : mlen + 2;
  const int cb = c_lo + blockIdx.z * P.seg;
  const int ce = min(cb + P.seg, c_hi);

  const bool second = SP ? true : (P.order != 1);
  const bool limit = SP ? true : (second && (P.limit != 0));
  const bool transverse = SP ? true : (P.m3 > 0);
  const bool capa = !SP && P.capa > 0;                      // capacity function: see sweep3_x
  const int icapa = capa ? P.capa - 1 : 0;
  const bool pencil = (I >= 1 && I <= P.n[0] + 2) &&
                      (IXYZ == 1 ? (T >= P.zlo && T <= P.zhi) : (T >= 1 && T <= P.n[TD] + 2));
  const bool out_thread = split ? (I < nx && T < nT)
                                : (tx >= 1 && tx <= BX - 2 && ty >= 1 && ty <= BY - 2 && I >= 2 && I <= P.n[0] + 1 &&
                                   T >= 2 && T <= P.n[TD] + 1);

  const long long col = (long long)Tc * strideT + Ic;
  const double* qcol = qin + col;
  double* ocol = qout + (long long)T * strideT + I;
  double q1[M], q2[M];                                       // q(n-1), q(n-2)
  double auxP[MA];                                           // aux(n-1)
  double waveP[W][M], sP[W], amdqP[M], apdqP[M], ctxP[NC], dotlP[W];   // interface n-1 (pending, unlimited)
  double apdqC[M], cqxxC[M], ctxC[NC];                       // interface n-2 (finished): the p-side of cell n-2
#pragma unroll
  for (int m = 0; m < M; ++m) { q1[m] = 0.0; q2[m] = 0.0; amdqP[m] = 0.0; apdqP[m] = 0.0; apdqC[m] = 0.0; cqxxC[m] = 0.0; }
#pragma unroll
  for (int mw = 0; mw < W; ++mw) {
    sP[mw] = 0.0; dotlP[mw] = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) waveP[mw][m] = 0.0;
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) { ctxP[c] = 0.0; ctxC[c] = 0.0; }
#pragma unroll
  for (int ma = 0; ma < MA; ++ma) auxP[ma] = 1.0;
  double cfl = 0.0;
  // dtdx1d (step3.f90:291-295 / 447-451) and the capacity of the cells n-1 (P) and n-2 (C) of my column
  double dP = dtdn, dC = dtdn, capaP = 1.0, capaC = 1.0;

  // operands of march step n, fetched asynchronously one step ahead into stage (n & 1):
  //   slots [0, M): q(n);  [M, M+MA): aux(n);  [M+MA, M+10MA): aux 3x3 of cell n-2;  [M+10MA, 2M+10MA): old qout(n-2)
  constexpr int PFQ = 0, PFA = M, PFN = M + SM3::MA, PFO = M + 10 * SM3::MA;
  const bool rmw = !split && !P.first_pass;
  auto prefetch = [&](int n) {
    double* st = spf + (size_t)((n & 1) * SM3::PF) * NT;
    const int nc = n < 0 ? 0 : (n < nlen ? n : nlen - 1);
#pragma unroll
    for (int m = 0; m < M; ++m) cp_async8(st + (PFQ + m) * NT, qcol + m * P.plane + (long long)nc * strideN);
    if (RP::MAUX > 0) {
#pragma unroll
      for (int ma = 0; ma < SM3::MA; ++ma)
        cp_async8(st + (PFA + ma) * NT, P.aux + ma * P.plane + col + (long long)nc * strideN);
      const int c = n - 2;
      if (c >= cb && c < ce && (SP || P.m3 > 0)) {
        // A[dy][dz]: y-sweep: dy <-> thread-y (K), dz <-> thread-x (I); z-sweep: dy <-> I, dz <-> J
        const int ti[3] = {Im, Ic, Ip};
        const int tt[3] = {Tm, Tc, Tp};
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            const int ii = (IXYZ == 1) ? ti[dz] : ti[dy];
            const int tj = (IXYZ == 1) ? tt[dy] : tt[dz];
#pragma unroll
            for (int ma = 0; ma < SM3::MA; ++ma)
              cp_async8(st + (PFN + (dy * 3 + dz) * SM3::MA + ma) * NT,
                        P.aux + ma * P.plane + (long long)tj * strideT + ii + (long long)c * strideN);
          }
      }
    }
    if (rmw && out_thread && n - 2 >= cb && n - 2 < ce) {
#pragma unroll
      for (int m = 0; m < M; ++m) cp_async8(st + (PFO + m) * NT, ocol + m * P.plane + (long long)(n - 2) * strideN);
    }
    cp_async_commit();
  };
  prefetch(cb - 2);

  for (int n = cb - 2; n <= ce + 1; ++n) {
    cp_async_wait_all();                                     // operands of step n have landed (my own slots)
    if (n < ce + 1) prefetch(n + 1);
    const double* st = spf + (size_t)((n & 1) * SM3::PF) * NT;
    double qN[M], auxN[MA];
#pragma unroll
    for (int m = 0; m < M; ++m) qN[m] = st[(PFQ + m) * NT];
#pragma unroll
    for (int ma = 0; ma < MA; ++ma) auxN[ma] = (RP::MAUX > 0) ? st[(PFA + ma) * NT] : 0.0;
    // ---- interface n (raw) ---------------------------------------------------------------
    double waveN[W][M], sN[W], amdqN[M], apdqN[M], ctxN[NC];
    RP::template normal<IXYZ>(P.rp, q1, qN, auxP, auxN, waveN, sN, amdqN, apdqN, ctxN);
    double capaN = 1.0;                                      // (fetched on its own: see sweep3_x)
    if (!SP && capa) {
      const int nc = n < 0 ? 0 : (n < nlen ? n : nlen - 1);
      capaN = P.aux[(long long)icapa * P.plane + col + (long long)nc * strideN];
    }
    const double dN = (!SP && capa) ? ddiv(dtdn, capaN) : dtdn;
    // (the specialised kernels see the constant: no carried copies in their loop)
    const double dPe = SP ? dtdn : dP, dCe = SP ? dtdn : dC;
    if (pencil && n >= 2 && n <= mlen + 2 && n > cb - 2) {                              // flux3.f90:211-218
        cfl = (!SP && capa) ? cfl_update_capa<W>(cfl, dPe, dN, sN) : cfl_update<W>(cfl, dtdn, sN);
    }
    // ---- finish interface n-1: limiter (needs waves n-2 via dotl, n-1, n) + cqxx -----------
    double cqxxP[M];
#pragma unroll
    for (int m = 0; m < M; ++m) cqxxP[m] = 0.0;
    double dotn[W];
#pragma unroll
    for (int mw = 0; mw < W; ++mw) dotn[mw] = 0.0;
    if (second) {
      if (limit) {
        double wnorm2[W];
        wave_dots3<M, W>(waveP, waveN, wnorm2, dotn);
        limit_waves3<M, W, LIM>(P.mthlim, waveP, sP, wnorm2, dotlP, dotn);
      }
      second_order3<M, W>(dCe, dPe, waveP, sP, cqxxP);           // interface n-1: cells n-2 | n-1
    }
    // ---- cell c = n-2: p-side = interface n-2 (carried), m-side = interface n-1 ---------------
    const int c = n - 2;
    const bool c_out = (c >= cb && c < ce);
    double qadd[M], fdiff[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      qadd[m] = (0.0 - dCe * apdqC[m]) - dCe * amdqP[m];
      fdiff[m] = (0.0 + cqxxP[m]) - (0.0 + cqxxC[m]);
    }
    if (split) {
      const bool upd = pencil && c >= 2 && c <= mlen + 1;
      if (out_thread && c_out) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const double r = (!SP && capa) ? (q2[m] + qadd[m]) - ddiv(dtdn * fdiff[m], capaC)    // step3ds.f90:282-287, 365-370
                                : (q2[m] + qadd[m]) - dtdn * fdiff[m];
          ocol[m * P.plane + (long long)c * strideN] = upd ? r : q2[m];
        }
      }
    } else if (c >= cb - 0 && c < ce) {                      // uniform across the CTA
      if (transverse) {
        double A[3][3][MA];
        if (RP::MAUX > 0) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)
#pragma unroll
              for (int ma = 0; ma < MA; ++ma) A[dy][dz][ma] = st[(PFN + (dy * 3 + dz) * SM3::MA + ma) * NT];
        }
        flux3_cell<RP, IXYZ, NT, M3C, M4C>(P, dCe, dty, dtz, apdqC, cqxxC, ctxC, amdqP, cqxxP, ctxP, A, sm);
      }
      __syncthreads();
      if (out_thread) {
        double acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = P.first_pass ? q2[m] : st[(PFO + m) * NT];
        if (!SP && capa) {                                   // every addend divided by my cell's capacity
          if (transverse) {
            if (IXYZ == 1) {
              apply_target_capa<RP, IXYZ, 1, 1, NT>(acc, sm - BX - 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 1, 0, NT>(acc, sm - BX, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 1, -1, NT>(acc, sm - BX + 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 0, 1, NT>(acc, sm - 1, dty, dtz, capaC);
              apply_center_capa<RP, IXYZ, NT>(acc, sm, qadd, fdiff, dtdn, dty, dtz, true, capaC);
              apply_target_capa<RP, IXYZ, 0, -1, NT>(acc, sm + 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, -1, 1, NT>(acc, sm + BX - 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, -1, 0, NT>(acc, sm + BX, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, -1, -1, NT>(acc, sm + BX + 1, dty, dtz, capaC);
            } else {
              apply_target_capa<RP, IXYZ, 1, 1, NT>(acc, sm - BX - 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 0, 1, NT>(acc, sm - BX, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, -1, 1, NT>(acc, sm - BX + 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 1, 0, NT>(acc, sm - 1, dty, dtz, capaC);
              apply_center_capa<RP, IXYZ, NT>(acc, sm, qadd, fdiff, dtdn, dty, dtz, true, capaC);
              apply_target_capa<RP, IXYZ, -1, 0, NT>(acc, sm + 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 1, -1, NT>(acc, sm + BX - 1, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, 0, -1, NT>(acc, sm + BX, dty, dtz, capaC);
              apply_target_capa<RP, IXYZ, -1, -1, NT>(acc, sm + BX + 1, dty, dtz, capaC);
            }
          } else {
            apply_center_capa<RP, IXYZ, NT>(acc, sm, qadd, fdiff, dtdn, dty, dtz, false, capaC);
          }
        } else
        if (transverse) {
          // source pencils in the reference's loop order: thread-y outer, thread-x inner
          // (step3.f90:283-284 "do k / do i", :439-440 "do j / do i")
          if (IXYZ == 1) {   // a <-> thread-y, b <-> thread-x;   (a, b) = -(source offset)
            apply_target<RP, IXYZ, 1, 1, NT>(acc, sm - BX - 1, dty, dtz);
            apply_target<RP, IXYZ, 1, 0, NT>(acc, sm - BX, dty, dtz);
            apply_target<RP, IXYZ, 1, -1, NT>(acc, sm - BX + 1, dty, dtz);
            apply_target<RP, IXYZ, 0, 1, NT>(acc, sm - 1, dty, dtz);
            apply_center<RP, IXYZ, NT>(acc, sm, qadd, fdiff, dtdn, dty, dtz, true);
            apply_target<RP, IXYZ, 0, -1, NT>(acc, sm + 1, dty, dtz);
            apply_target<RP, IXYZ, -1, 1, NT>(acc, sm + BX - 1, dty, dtz);
            apply_target<RP, IXYZ, -1, 0, NT>(acc, sm + BX, dty, dtz);
            apply_target<RP, IXYZ, -1, -1, NT>(acc, sm + BX + 1, dty, dtz);
          } else {           // a <-> thread-x, b <-> thread-y
            apply_target<RP, IXYZ, 1, 1, NT>(acc, sm - BX - 1, dty, dtz);
            apply_target<RP, IXYZ, 0, 1, NT>(acc, sm - BX, dty, dtz);
            apply_target<RP, IXYZ, -1, 1, NT>(acc, sm - BX + 1, dty, dtz);
            apply_target<RP, IXYZ, 1, 0, NT>(acc, sm - 1, dty, dtz);
            apply_center<RP, IXYZ, NT>(acc, sm, qadd, fdiff, dtdn, dty, dtz, true);
            apply_target<RP, IXYZ, -1, 0, NT>(acc, sm + 1, dty, dtz);
            apply_target<RP, IXYZ, 1, -1, NT>(acc, sm + BX - 1, dty, dtz);
            apply_target<RP, IXYZ, 0, -1, NT>(acc, sm + BX, dty, dtz);
            apply_target<RP, IXYZ, -1, -1, NT>(acc, sm + BX + 1, dty, dtz);
          }
        } else {
          apply_center<RP, IXYZ, NT>(acc, sm, qadd, fdiff, dtdn, dty, dtz, false);
        }
#pragma unroll
        for (int m = 0; m < M; ++m) ocol[m * P.plane + (long long)c * strideN] = acc[m];
      }
      __syncthreads();
    }
    // ---- carry ------------------------------------------------------------------------------------
#pragma unroll
    for (int m = 0; m < M; ++m) {
      q2[m] = q1[m]; q1[m] = qN[m];
      apdqC[m] = apdqP[m]; cqxxC[m] = cqxxP[m];
      amdqP[m] = amdqN[m]; apdqP[m] = apdqN[m];
    }
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) { ctxC[cc] = ctxP[cc]; ctxP[cc] = ctxN[cc]; }
#pragma unroll
    for (int ma = 0; ma < MA; ++ma) auxP[ma] = auxN[ma];
    if (!SP) { dC = dP; dP = dN; capaC = capaP; capaP = capaN; }
#pragma unroll
    for (int mw = 0; mw < W; ++mw) {
      sP[mw] = sN[mw]; dotlP[mw] = dotn[mw];
#pragma unroll
      for (int m = 0; m < M; ++m) waveP[mw][m] = waveN[mw][m];
    }
  }
  block_max_cfl3<NT>(cfl, P.cfl_bits, sred);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
