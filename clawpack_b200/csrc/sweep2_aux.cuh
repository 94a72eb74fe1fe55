// 2-D classic sweeps for Riemann solvers that read aux arrays, with capacity functions (index_capa) and
// f-wave solvers (use_fwave) -- the remaining branches of step2.f90 / step2ds.f90 / flux2.f90:
//   aux1/aux2/aux3 pencils      step2.f90:107-115, 186-194   (aux of rows j-1, j, j+1 / columns i-1, i, i+1)
//   dtdx1d(i) = dtdx/aux(capa)  step2.f90:101-105, 180-184;  capa form of the update :151-164, 232-246
//   f-wave second-order term    flux2.f90:148-160
// Same strip-marching, gather-instead-of-scatter algorithm and accumulation order as sweep2_unsplit /
// sweep2_split (sweep2.cuh), in the warp-synchronous form of sweep2_ws.cuh: every warp owns 28 columns (+2 halo
// columns per side), x-neighbour data -- including the aux values of the neighbouring columns that the
// transverse solves read -- move by warp shuffle, the pending y-interface lives in thread-private shared memory.
// The aux planes use the same SoA layout as q.
//
// Functor interface (rp2_aux.cuh), components in the rotated frame like rp2.cuh:
//   RP::MAUX                                  aux planes the solver reads (the first MAUX planes)
//   RP::normal(P, dir, ql, qr, axl, axr, wave, s, amdq, apdq, roe)   ql/axl = left cell, qr/axr = right cell
//   RP::speeds(P, dir, roe, s)
//   RP::transverse(P, dir, roe, ax1, ax2, ax3, bx1, bx2, bx3, asdq, bm, bp)   ax* = aux of the cell the fluctuation enters
//                                                                    (i1 = i-2+imp) in pencils 1, 2, 3
#pragma once
#include "sweep2_ws.cuh"

namespace clawcuda {
namespace CLAW_NS {

// flux2.f90:134-160, both forms
template <class RP>
__device__ __forceinline__ void second_order_g(double dtdxave, const double (*wave)[RP::MEQN], const double* s,
                                               bool fwave, double* cqxx) {
#pragma unroll
  for (int m = 0; m < RP::MEQN; ++m) {
    double c = 0.0;
#pragma unroll
    for (int mw = 0; mw < RP::MWAVES; ++mw) {
      const double lead = fwave ? copysign(1.0, s[mw]) : fabs(s[mw]);
      const double t = lead * (1.0 - fabs(s[mw]) * dtdxave) * wave[mw][m];
      c = (mw == 0) ? from0(t) : c + t;
    }
    cqxx[m] = c;
  }
}

// flux2.f90:111-119 with per-cell dtdx1d: max(cfl, dtdx1d(i)*s, -dtdx1d(i-1)*s)
template <int W>
__device__ __forceinline__ double cfl_update2(double cfl, double dtd_r, double dtd_l, const double* s) {
#pragma unroll
  for (int mw = 0; mw < W; ++mw) cfl = fmax2(fmax2(cfl, dtd_r * s[mw]), -dtd_l * s[mw]);
  return cfl;
}

template <class RP, int NW> struct Smem2aux {
  static constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE;
  static constexpr int OFF_YW = 0;                  // [W*M]   waves of the pending y-interface
  static constexpr int OFF_YA = OFF_YW + W * M;     // [2M+NR+W] amdq, apdq, roe, dotl
  static constexpr int PER_THREAD = OFF_YA + 2 * M + NR + W;
  static constexpr size_t BYTES = (size_t)PER_THREAD * NW * 32 * sizeof(double) + 64 * sizeof(double);
};

// =====================================================================================
// Unsplit step (step2.f90), method(3) in {0,1,2}
// =====================================================================================
template <class RP, int NW>
__global__ void __launch_bounds__(NW * 32) sweep2_unsplit_aux(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, MA = RP::MAUX, NT = NW * 32;
  using S = Smem2aux<RP, NW>;
  constexpr int TX = 28;
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x, lane = t & 31;
  double* const syw = smem + S::OFF_YW * NT + t;
  double* const sya = smem + S::OFF_YA * NT + t;
  double* const sred = smem + S::PER_THREAD * NT;

  const int nx = P.mx + 4;
  const int I = (blockIdx.x * NW + (t >> 5)) * TX + lane;
  const int Ic = I < nx ? I : nx - 1;
  const int Jb = P.j0 + blockIdx.y * P.seg_rows;
  const int Je = min(Jb + P.seg_rows, P.j1);
  const int Jtop = P.my + 3;                               // last row of the array

  CLAW_RESOLVE_STEP2(P, qin, qout, dtdx, dtdy);
  const bool second = (P.order != 1);
  const bool limit = second && (P.limit != 0);
  const int mt = P.mtrans;
  const bool fw = P.fwave != 0;
  const bool cp = P.capa > 0;

  const double* qcol = qin + Ic;
  const double* acol = P.aux + Ic;
  const double* kcol = cp ? P.aux + (long long)(P.capa - 1) * P.plane + Ic : nullptr;
  auto row = [&](int J) { return (long long)(J < 0 ? 0 : (J > Jtop ? Jtop : J)) * P.pitch; };

  double q1[M], qn0[M], qn1[M];
  double a0[MA], a1[MA];                                   // aux of rows J, J+1
  double k0 = 1.0, k1 = 1.0;                               // capacity of rows J, J+1
  double yI_apdq[M], yI_fadd[M], yI_bm[M], yI_bp[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    q1[m] = qcol[m * P.plane + row(Jb - 2)];
    qn0[m] = 0.0; qn1[m] = 0.0; yI_apdq[m] = 0.0; yI_fadd[m] = 0.0; yI_bm[m] = 0.0; yI_bp[m] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < MA; ++k) { a0[k] = acol[k * P.plane + row(Jb - 3)]; a1[k] = acol[k * P.plane + row(Jb - 2)]; }
  if (cp) { k0 = kcol[row(Jb - 3)]; k1 = kcol[row(Jb - 2)]; }
  double cfl = 0.0;
  const bool x_cfl_col = (lane >= 1) && (I >= 2) && (I <= P.mx + 2);
  const bool y_cfl_col = (I >= 1) && (I <= P.mx + 2);
  const bool store_col = (lane >= 2) && (lane <= 29) && (I <= P.mx + 1);

  for (int J = Jb - 3; J < Je; ++J) {
    double qN[M], aN[MA], kN = 1.0;
#pragma unroll
    for (int m = 0; m < M; ++m) qN[m] = qcol[m * P.plane + row(J + 2)];
#pragma unroll
    for (int k = 0; k < MA; ++k) aN[k] = acol[k * P.plane + row(J + 2)];
    if (cp) kN = kcol[row(J + 2)];
    // dtdy1d of the y-pencil cells J, J+1, J+2 (step2.f90:180-184); dtdx1d of the x-pencil cell (I, J+1) (:101-105)
    const double dy0 = cp ? ddiv(dtdy, k0) : dtdy, dy1 = cp ? ddiv(dtdy, k1) : dtdy, dyN = cp ? ddiv(dtdy, kN) : dtdy;
    const double dxC = cp ? ddiv(dtdx, k1) : dtdx;

    // =========================== y-pipeline ==========================================
    double yq[M], yb[M], gy1[M], gy2[M];                   // natural frame: qadd, flux bracket, gadd(1), gadd(2)
#pragma unroll
    for (int m = 0; m < M; ++m) { yq[m] = 0.0; yb[m] = 0.0; gy1[m] = 0.0; gy2[m] = 0.0; }
    {
      double waveN[W][M], sN[W], amdqN[M], apdqN[M], roeN[NR];
      {
        double ql[M], qr[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { ql[m] = q1[rot<M>(m)]; qr[m] = qN[rot<M>(m)]; }
        RP::normal(P.rp, 1, ql, qr, a1, aN, waveN, sN, amdqN, apdqN, roeN);
      }
      if (y_cfl_col && (J + 2 >= 2) && (J + 2 <= P.my + 2)) cfl = cfl_update2<W>(cfl, dyN, dy1, sN);
      double dotn[W];
#pragma unroll
      for (int mw = 0; mw < W; ++mw) dotn[mw] = 0.0;
      // aux of the neighbouring columns in rows J and J+1 (whole warp, every step)
      double aL0[MA], aR0[MA], aL1[MA], aR1[MA];
#pragma unroll
      for (int k = 0; k < MA; ++k) {
        aL0[k] = shfl_left(a0[k]); aR0[k] = shfl_right(a0[k]);
        aL1[k] = shfl_left(a1[k]); aR1[k] = shfl_right(a1[k]);
      }
      if (J >= Jb - 2) {
        double waveC[W][M], wnorm2[W];
#pragma unroll
        for (int mw = 0; mw < W; ++mw) {
          wnorm2[mw] = 0.0;
#pragma unroll
          for (int m = 0; m < M; ++m) waveC[mw][m] = syw[(mw * M + m) * NT];
        }
        if (limit) wave_dots<RP, true>(waveC, waveN, wnorm2, dotn);
        if (J >= Jb - 1) {
          double sC[W], amdqC[M], apdqC[M], roeC[NR], dotl[W];
#pragma unroll
          for (int m = 0; m < M; ++m) { amdqC[m] = sya[m * NT]; apdqC[m] = sya[(M + m) * NT]; }
#pragma unroll
          for (int k = 0; k < NR; ++k) roeC[k] = sya[(2 * M + k) * NT];
#pragma unroll
          for (int mw = 0; mw < W; ++mw) dotl[mw] = sya[(2 * M + NR + mw) * NT];
          RP::speeds(P.rp, 1, roeC, sC);
          double qadd[M];
#pragma unroll
          for (int m = 0; m < M; ++m) {                                    // flux2.f90:105-108, cell J
            qadd[m] = (0.0 - dy0 * yI_apdq[m]) - dy0 * amdqC[m];
            yI_apdq[m] = apdqC[m];
          }
          double fadd[M];
#pragma unroll
          for (int m = 0; m < M; ++m) fadd[m] = 0.0;
          if (second) {
            if (limit) limit_waves<RP, 0>(P, waveC, sC, wnorm2, dotl, dotn);
            double cqxx[M];
            second_order_g<RP>(0.5 * (dy0 + dy1), waveC, sC, fw, cqxx);    // flux2.f90:134-160
#pragma unroll
            for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
            if (mt == 2) {                                                 // flux2.f90:167-173
#pragma unroll
              for (int m = 0; m < M; ++m) { amdqC[m] = amdqC[m] + cqxx[m]; apdqC[m] = apdqC[m] - cqxx[m]; }
            }
          }
          double bmA[M], bpA[M], bmP[M], bpP[M];
#pragma unroll
          for (int m = 0; m < M; ++m) { bmA[m] = 0.0; bpA[m] = 0.0; bmP[m] = 0.0; bpP[m] = 0.0; }
          if (mt > 0) {
            RP::transverse(P.rp, 1, roeC, aL0, a0, aR0, aL1, a1, aR1, amdqC, bmA, bpA);  // imp = 1: cell J  (flux2.f90:181)
            RP::transverse(P.rp, 1, roeC, aL1, a1, aR1, aL1, a1, aR1, apdqC, bmP, bpP);  // imp = 2: cell J+1 (:194)
          }
          const double hd0 = 0.5 * dy0;
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const double g1 = (0.0 - hd0 * bmA[m]) - hd0 * yI_bm[m];       // flux2.f90:185-203, cell J
            const double g2 = (0.0 - hd0 * bpA[m]) - hd0 * yI_bp[m];
            yq[rot<M>(m)] = qadd[m];
            // no capa (step2.f90:221-225): qadd - dtdy*dfadd - dtdx*dgadd;  capa (:234-238): the bracket / capa
            yb[rot<M>(m)] = cp ? ddiv(dtdy * (fadd[m] - yI_fadd[m]) + dtdx * (g2 - g1), k0)
                               : (qadd[m] - dtdy * (fadd[m] - yI_fadd[m])) - dtdx * (g2 - g1);
            gy1[rot<M>(m)] = g1;
            gy2[rot<M>(m)] = g2;
            yI_fadd[m] = fadd[m];
          }
          // fluctuations entering cell J+1 are scaled by ITS dtdy1d when that cell is finished next step
#pragma unroll
          for (int m = 0; m < M; ++m) { yI_bm[m] = bmP[m]; yI_bp[m] = bpP[m]; }
        }
      }
#pragma unroll
      for (int mw = 0; mw < W; ++mw)
#pragma unroll
        for (int m = 0; m < M; ++m) syw[(mw * M + m) * NT] = waveN[mw][m];
#pragma unroll
      for (int m = 0; m < M; ++m) { sya[m * NT] = amdqN[m]; sya[(M + m) * NT] = apdqN[m]; }
#pragma unroll
      for (int k = 0; k < NR; ++k) sya[(2 * M + k) * NT] = roeN[k];
#pragma unroll
      for (int mw = 0; mw < W; ++mw) sya[(2 * M + NR + mw) * NT] = dotn[mw];
    }

    // =========================== x-pipeline on row J+1 ===============================
    double qn2[M];
#pragma unroll
    for (int m = 0; m < M; ++m) qn2[m] = qN[m];
    {
      // aux of the left neighbour column in rows J, J+1, J+2 (for imp = 1) -- whole warp, every step
      double aLl0[MA], aLl1[MA], aLlN[MA];
#pragma unroll
      for (int k = 0; k < MA; ++k) { aLl0[k] = shfl_left(a0[k]); aLl1[k] = shfl_left(a1[k]); aLlN[k] = shfl_left(aN[k]); }
      const double dxL = shfl_left(dxC);
      if (J >= Jb - 2) {
        double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
        {
          double ql[M];
#pragma unroll
          for (int m = 0; m < M; ++m) ql[m] = shfl_left(q1[m]);
          RP::normal(P.rp, 0, ql, q1, aLl1, a1, wave, s, amdq, apdq, roe);  // flux2.f90:101
        }
        if (x_cfl_col && (J + 1 >= 1) && (J + 1 <= P.my + 2)) cfl = cfl_update2<W>(cfl, dxC, dxL, s);
#pragma unroll
        for (int m = 0; m < M; ++m) { cqxx[m] = 0.0; fadd[m] = 0.0; }
        if (second) {
          if (limit) {
            double wnorm2[W], dotr[W], dotl[W];
            {
              double waveR[W][M];
#pragma unroll
              for (int mw = 0; mw < W; ++mw)
#pragma unroll
                for (int m = 0; m < M; ++m) waveR[mw][m] = shfl_right(wave[mw][m]);
              wave_dots<RP, false>(wave, waveR, wnorm2, dotr);
            }
#pragma unroll
            for (int mw = 0; mw < W; ++mw) dotl[mw] = shfl_left(dotr[mw]);
            limit_waves<RP, 0>(P, wave, s, wnorm2, dotl, dotr);
          }
          second_order_g<RP>(0.5 * (dxL + dxC), wave, s, fw, cqxx);
#pragma unroll
          for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
        }
        double bmA[M], bpA[M], bmP[M], bpP[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { bmA[m] = 0.0; bpA[m] = 0.0; bmP[m] = 0.0; bpP[m] = 0.0; }
        if (mt > 0) {
          double tm[M], tp[M];
#pragma unroll
          for (int m = 0; m < M; ++m) { tm[m] = amdq[m]; tp[m] = apdq[m]; }
          if (second && mt == 2) {
#pragma unroll
            for (int m = 0; m < M; ++m) { tm[m] = tm[m] + cqxx[m]; tp[m] = tp[m] - cqxx[m]; }
          }
          RP::transverse(P.rp, 0, roe, aLl0, aLl1, aLlN, a0, a1, aN, tm, bmA, bpA);     // imp = 1: cell I-1
          RP::transverse(P.rp, 0, roe, a0, a1, aN, a0, a1, aN, tp, bmP, bpP);           // imp = 2: cell I
        }
        const double hdC = 0.5 * dxC;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const double amdqR = shfl_right(amdq[m]), faddR = shfl_right(fadd[m]);
          const double bmAR = shfl_right(bmA[m]), bpAR = shfl_right(bpA[m]);
          const double qadd = (0.0 - dxC * apdq[m]) - dxC * amdqR;         // flux2.f90:105-108
          const double g1 = (0.0 - hdC * bmAR) - hdC * bmP[m];            // flux2.f90:185-203
          const double g2 = (0.0 - hdC * bpAR) - hdC * bpP[m];
          if (!cp) {                                                       // step2.f90:139-147
            qn0[m] = qn0[m] - dtdy * g1;
            qn1[m] = ((qn1[m] + qadd) - dtdx * (faddR - fadd[m])) - dtdy * (g2 - g1);
            qn2[m] = qN[m] + dtdy * g2;
          } else {                                                         // step2.f90:151-164
            qn0[m] = qn0[m] - ddiv(dtdy * g1, k0);
            qn1[m] = (qn1[m] + qadd) - ddiv(dtdx * (faddR - fadd[m]) + dtdy * (g2 - g1), k1);
            qn2[m] = qN[m] + ddiv(dtdy * g2, kN);
          }
        }
      } else {
        // keep the warp's shuffle count uniform across the warm-up steps
#pragma unroll
        for (int m = 0; m < M; ++m) (void)shfl_left(q1[m]);
      }
    }

    // =========================== cell (I, J) is complete =============================
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const double g2L = shfl_left(gy2[m]), g1R = shfl_right(gy1[m]);
      double r;
      if (!cp) {                                                           // step2.f90:221-229
        r = qn0[m] + dtdx * g2L;
        r = r + yb[m];
        r = r - dtdx * g1R;
      } else {                                                             // step2.f90:232-246
        r = qn0[m] + ddiv(dtdx * g2L, k0);
        r = (r + yq[m]) - yb[m];
        r = r - ddiv(dtdx * g1R, k0);
      }
      if (store_col && J >= Jb) qout[m * P.plane + (long long)J * P.pitch + I] = r;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) { qn0[m] = qn1[m]; qn1[m] = qn2[m]; q1[m] = qN[m]; }
#pragma unroll
    for (int k = 0; k < MA; ++k) { a0[k] = a1[k]; a1[k] = aN[k]; }
    k0 = k1; k1 = kN;
  }
  block_max_cfl<NT>(cfl, P.cfl_bits, sred);
}

// =====================================================================================
// Dimension-split step (two step2ds.f90 calls fused; do_x / do_y select a single sweep).
//   x-sweep: all rows, cells I in [2, mx+1]      (step2ds.f90:96, 149-161)
//   y-sweep: all columns, cells J in [2, my+1]   (step2ds.f90:180, 237-250)
// =====================================================================================
template <class RP, int NW>
__global__ void __launch_bounds__(NW * 32) sweep2_split_aux(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, MA = RP::MAUX, NT = NW * 32;
  using S = Smem2aux<RP, NW>;
  constexpr int TX = 28;
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x, lane = t & 31;
  double* const syw = smem + S::OFF_YW * NT + t;
  double* const sya = smem + S::OFF_YA * NT + t;
  double* const sred = smem + S::PER_THREAD * NT;

  const int nx = P.mx + 4;
  const int I = (blockIdx.x * NW + (t >> 5)) * TX + lane;
  const int Ic = I < nx ? I : nx - 1;
  const int Jtop = P.my + 3;
  CLAW_RESOLVE_STEP2(P, qin, qout, dtdx, dtdy);
  const bool second = (P.order != 1);
  const bool limit = second && (P.limit != 0);
  const bool fw = P.fwave != 0;
  const bool cp = P.capa > 0;
  const bool do_x = P.do_x != 0, do_y = P.do_y != 0;
  const double* qcol = qin + Ic;
  const double* acol = P.aux + Ic;
  const double* kcol = cp ? P.aux + (long long)(P.capa - 1) * P.plane + Ic : nullptr;
  auto row = [&](int J) { return (long long)(J < 0 ? 0 : (J > Jtop ? Jtop : J)) * P.pitch; };
  double cfl = 0.0;
  const bool x_cfl_col = (lane >= 1) && (I >= 2) && (I <= P.mx + 2);
  const bool interior_col = (I >= 2) && (I <= P.mx + 1);
  const bool x_ok = (lane >= 2) && (lane <= 29);
  const bool col_ok = (I <= P.mx + 3) && (!do_x || !interior_col || x_ok);

  // x-sweep of one row: q (in) -> x-updated value (out, interior columns of valid lanes only)
  auto x_sweep = [&](const double* q, const double* a, double kap, double* out) {
    const double dxC = cp ? ddiv(dtdx, kap) : dtdx;
    const double dxL = shfl_left(dxC);
    double aL[MA], ql[M];
#pragma unroll
    for (int k = 0; k < MA; ++k) aL[k] = shfl_left(a[k]);
#pragma unroll
    for (int m = 0; m < M; ++m) ql[m] = shfl_left(q[m]);
    double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
    RP::normal(P.rp, 0, ql, q, aL, a, wave, s, amdq, apdq, roe);
    if (x_cfl_col) cfl = cfl_update2<W>(cfl, dxC, dxL, s);
#pragma unroll
    for (int m = 0; m < M; ++m) { cqxx[m] = 0.0; fadd[m] = 0.0; }
    if (second) {
      if (limit) {
        double wnorm2[W], dotr[W], dotl[W];
        {
          double waveR[W][M];
#pragma unroll
          for (int mw = 0; mw < W; ++mw)
#pragma unroll
            for (int m = 0; m < M; ++m) waveR[mw][m] = shfl_right(wave[mw][m]);
          wave_dots<RP, false>(wave, waveR, wnorm2, dotr);
        }
#pragma unroll
        for (int mw = 0; mw < W; ++mw) dotl[mw] = shfl_left(dotr[mw]);
        limit_waves<RP, 0>(P, wave, s, wnorm2, dotl, dotr);
      }
      second_order_g<RP>(0.5 * (dxL + dxC), wave, s, fw, cqxx);
#pragma unroll
      for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const double amdqR = shfl_right(amdq[m]), faddR = shfl_right(fadd[m]);
      const double qadd = (0.0 - dxC * apdq[m]) - dxC * amdqR;
      // step2ds.f90:149-152 / 157-161
      out[m] = cp ? (q[m] + qadd) - ddiv(dtdx * (faddR - fadd[m]), kap) : (q[m] + qadd) - dtdx * (faddR - fadd[m]);
    }
  };

  if (!do_y) {
    const int Jb = P.j0 + blockIdx.y * P.seg_rows;
    const int Je = min(Jb + P.seg_rows, P.j1);
    for (int J = Jb; J < Je; ++J) {
      double q[M], a[MA], kap = 1.0, r[M];
#pragma unroll
      for (int m = 0; m < M; ++m) q[m] = qcol[m * P.plane + row(J)];
#pragma unroll
      for (int k = 0; k < MA; ++k) a[k] = acol[k * P.plane + row(J)];
      if (cp) kap = kcol[row(J)];
      x_sweep(q, a, kap, r);
      if (interior_col && x_ok) {
#pragma unroll
        for (int m = 0; m < M; ++m) qout[m * P.plane + (long long)J * P.pitch + I] = r[m];
      }
    }
    block_max_cfl<NT>(cfl, P.cfl_bits, sred);
    return;
  }

  const int Jb = P.j0 + blockIdx.y * P.seg_rows;
  const int Je = min(Jb + P.seg_rows, P.j1);
  double qx0[M], qx1[M], a1[MA], k0 = 1.0, k1 = 1.0;       // x-updated rows J, J+1; aux / capa of row J+1
  double yI_apdq[M], yI_fadd[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { qx0[m] = 0.0; qx1[m] = 0.0; yI_apdq[m] = 0.0; yI_fadd[m] = 0.0; }
#pragma unroll
  for (int k = 0; k < MA; ++k) a1[k] = acol[k * P.plane + row(Jb - 3)];

  for (int J = Jb - 4; J < Je; ++J) {
    double qx2[M], aN[MA], kN = 1.0;
#pragma unroll
    for (int m = 0; m < M; ++m) qx2[m] = qcol[m * P.plane + row(J + 2)];
#pragma unroll
    for (int k = 0; k < MA; ++k) aN[k] = acol[k * P.plane + row(J + 2)];
    if (cp) kN = kcol[row(J + 2)];
    if (do_x) {
      double r[M];
      x_sweep(qx2, aN, kN, r);
      if (interior_col) {
#pragma unroll
        for (int m = 0; m < M; ++m) qx2[m] = r[m];
      }
    }
    const double dy0 = cp ? ddiv(dtdy, k0) : dtdy, dy1 = cp ? ddiv(dtdy, k1) : dtdy, dyN = cp ? ddiv(dtdy, kN) : dtdy;
    if (J >= Jb - 3) {
      double waveN[W][M], sN[W], amdqN[M], apdqN[M], roeN[NR];
      {
        double ql[M], qr[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { ql[m] = qx1[rot<M>(m)]; qr[m] = qx2[rot<M>(m)]; }
        RP::normal(P.rp, 1, ql, qr, a1, aN, waveN, sN, amdqN, apdqN, roeN);
      }
      if (col_ok && (J + 2 >= 2) && (J + 2 <= P.my + 2)) cfl = cfl_update2<W>(cfl, dyN, dy1, sN);
      double dotn[W];
#pragma unroll
      for (int mw = 0; mw < W; ++mw) dotn[mw] = 0.0;
      if (J >= Jb - 2) {
        double waveC[W][M], wnorm2[W];
#pragma unroll
        for (int mw = 0; mw < W; ++mw) {
          wnorm2[mw] = 0.0;
#pragma unroll
          for (int m = 0; m < M; ++m) waveC[mw][m] = syw[(mw * M + m) * NT];
        }
        if (limit) wave_dots<RP, true>(waveC, waveN, wnorm2, dotn);
        if (J >= Jb - 1) {
          double sC[W], amdqC[M], apdqC[M], roeC[NR], dotl[W];
#pragma unroll
          for (int m = 0; m < M; ++m) { amdqC[m] = sya[m * NT]; apdqC[m] = sya[(M + m) * NT]; }
#pragma unroll
          for (int k = 0; k < NR; ++k) roeC[k] = sya[(2 * M + k) * NT];
#pragma unroll
          for (int mw = 0; mw < W; ++mw) dotl[mw] = sya[(2 * M + NR + mw) * NT];
          RP::speeds(P.rp, 1, roeC, sC);
          double fadd[M];
#pragma unroll
          for (int m = 0; m < M; ++m) fadd[m] = 0.0;
          if (second) {
            if (limit) limit_waves<RP, 0>(P, waveC, sC, wnorm2, dotl, dotn);
            double cqxx[M];
            second_order_g<RP>(0.5 * (dy0 + dy1), waveC, sC, fw, cqxx);
#pragma unroll
            for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
          }
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const double qadd = (0.0 - dy0 * yI_apdq[m]) - dy0 * amdqC[m];
            // step2ds.f90:237-240 / 246-250
            const double r = cp ? (qx0[rot<M>(m)] + qadd) - ddiv(dtdy * (fadd[m] - yI_fadd[m]), k0)
                                : (qx0[rot<M>(m)] + qadd) - dtdy * (fadd[m] - yI_fadd[m]);
            if (J >= Jb && col_ok && (do_x ? interior_col : true))
              qout[rot<M>(m) * P.plane + (long long)J * P.pitch + I] = r;
            yI_apdq[m] = apdqC[m]; yI_fadd[m] = fadd[m];
          }
        }
      }
#pragma unroll
      for (int mw = 0; mw < W; ++mw)
#pragma unroll
        for (int m = 0; m < M; ++m) syw[(mw * M + m) * NT] = waveN[mw][m];
#pragma unroll
      for (int m = 0; m < M; ++m) { sya[m * NT] = amdqN[m]; sya[(M + m) * NT] = apdqN[m]; }
#pragma unroll
      for (int k = 0; k < NR; ++k) sya[(2 * M + k) * NT] = roeN[k];
#pragma unroll
      for (int mw = 0; mw < W; ++mw) sya[(2 * M + NR + mw) * NT] = dotn[mw];
    }
#pragma unroll
    for (int m = 0; m < M; ++m) { qx0[m] = qx1[m]; qx1[m] = qx2[m]; }
#pragma unroll
    for (int k = 0; k < MA; ++k) a1[k] = aN[k];
    k0 = k1; k1 = kN;
  }
  block_max_cfl<NT>(cfl, P.cfl_bits, sred);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
