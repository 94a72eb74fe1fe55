// Warp-SPECIALISED form of the fused unsplit 2-D sweep for the headline method set (order 2, one limiter id on all
// wave families, method(3) = 2); same algorithm, accumulation order and bits as sweep2_unsplit_ws (sweep2_ws.cuh).
//
// Why: sweep2_unsplit_ws needs ~250 registers per thread for the x- and the y-pipeline together, which caps the SM at
// 2 warps per scheduler; ncu (profiles/r01_i) shows the FP64 pipe 61 % busy with `wait` (dependent-issue latency) as
// the top stall -- not enough independent instruction streams.  Here a CTA is a PAIR of warps working on the same
// 28-column strip:
//   warp 0 ("Y"): loads the rows (cp.async ring), runs the y-pipeline (normal solve of y-interface J+2, limiter +
//                 second-order + two transverse solves of interface J+1) -- purely per column, no shuffles -- and
//                 publishes for cell (I, J) the three vectors ycomb, gadd1, gadd2, plus every row it loaded (q and the
//                 per-cell Roe data) in a 4-row ring;
//   warp 1 ("X"): one march step behind, runs the x-pipeline of row J+1 (neighbouring interfaces by warp shuffle),
//                 accumulates the three row sums and finishes cell (I, J) with warp Y's vectors.
// One __syncthreads (64 threads) per march step hands the double-buffered exchange over.  Each role keeps about half
// of the live state, so the kernel fits 168 registers and 3 warps per scheduler run instead of 2.
#pragma once
#include "sweep2_ws.cuh"

namespace clawcuda {
namespace CLAW_NS {

template <class RP> struct Smem2wsp {
  static constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, NC = RP::NCELL;
  static constexpr int OFF_YW = 0;                        // [W*M]      Y: waves of the pending y-interface
  static constexpr int OFF_YA = OFF_YW + W * M;           // [2M+NR+W]  Y: amdq, apdq, roe, dotl
  static constexpr int OFF_PF = OFF_YA + 2 * M + NR + W;  // [2*M]      Y: cp.async ring of rows
  static constexpr int OFF_ROW = OFF_PF + 2 * M;          // [4*(M+NC)] Y -> X: rows (q, cell data), ring of 4
  static constexpr int OFF_EX = OFF_ROW + 4 * (M + NC);   // [2*3M]     Y -> X: ycomb, gadd1, gadd2, two stages
  static constexpr int PER_LANE = OFF_EX + 6 * M;
  static constexpr size_t BYTES = (size_t)PER_LANE * 32 * sizeof(double) + 64 * sizeof(double);
};

template <class RP, int MINB, int LIM>
__global__ void __launch_bounds__(64, MINB) sweep2_unsplit_wsp(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, NT = 32;
  constexpr int NC = RP::NCELL, NC1 = NC > 0 ? NC : 1;
  using S = Smem2wsp<RP>;
  constexpr int TX = 28;
  CLAW_DYN_SMEM(smem);
  const int lane = threadIdx.x & 31;
  const bool roleY = threadIdx.x < 32;
  double* const syw = smem + S::OFF_YW * NT + lane;
  double* const sya = smem + S::OFF_YA * NT + lane;
  double* const spf = smem + S::OFF_PF * NT + lane;
  double* const srow = smem + S::OFF_ROW * NT + lane;
  double* const sex = smem + S::OFF_EX * NT + lane;
  double* const sred = smem + S::PER_LANE * NT;

  const int nx = P.mx + 4;
  const int I = blockIdx.x * TX + lane;                    // my column (array coords)
  const int Ic = I < nx ? I : nx - 1;
  const int Jb = P.j0 + blockIdx.y * P.seg_rows;
  const int Je = min(Jb + P.seg_rows, P.j1);
  const double dtdx = P.dtdx, dtdy = P.dtdy;
  const double hdx = 0.5 * dtdx, hdy = 0.5 * dtdy;
  double cfl = 0.0;

  if (roleY) {
    // =============================== warp Y ===============================================
    const double* qcol = P.qin + Ic;
    const bool y_cfl_col = (I >= 1) && (I <= P.mx + 2);
    double q1[M], c1[NC1];
    double yI_apdq[M], yI_fadd[M], yI_bm[M], yI_bp[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      q1[m] = qcol[m * P.plane + (long long)(Jb - 2) * P.pitch];
      yI_apdq[m] = 0.0; yI_fadd[m] = 0.0; yI_bm[m] = 0.0; yI_bp[m] = 0.0;
    }
    if constexpr (NC > 0) RP::cell(P.rp, q1, c1);
    {                                                      // publish row Jb-2 for warp X
      double* r = srow + (((Jb - 2) & 3) * (M + NC)) * NT;
#pragma unroll
      for (int m = 0; m < M; ++m) r[m * NT] = q1[m];
#pragma unroll
      for (int k = 0; k < NC; ++k) r[(M + k) * NT] = c1[k];
    }
#pragma unroll
    for (int m = 0; m < M; ++m)
      cp_async8(spf + (((Jb - 3) & 1) * M + m) * NT, qcol + m * P.plane + (long long)(Jb - 1) * P.pitch);
    cp_async_commit();

    for (int J = Jb - 3; J <= Je; ++J) {                   // one extra iteration: warp X runs one step behind
      if (J < Je) {
        double qN[M], cN[NC1];
        cp_async_wait_all();
#pragma unroll
        for (int m = 0; m < M; ++m) qN[m] = spf[((J & 1) * M + m) * NT];
        if (J + 1 < Je) {
#pragma unroll
          for (int m = 0; m < M; ++m)
            cp_async8(spf + (((J + 1) & 1) * M + m) * NT, qcol + m * P.plane + (long long)(J + 3) * P.pitch);
        }
        cp_async_commit();
        if constexpr (NC > 0) RP::cell(P.rp, qN, cN);
        {                                                  // publish row J+2
          double* r = srow + (((J + 2) & 3) * (M + NC)) * NT;
#pragma unroll
          for (int m = 0; m < M; ++m) r[m * NT] = qN[m];
#pragma unroll
          for (int k = 0; k < NC; ++k) r[(M + k) * NT] = cN[k];
        }
        double ycomb[M], gy1[M], gy2[M];                   // natural frame
        double waveN[W][M], sN[W], amdqN[M], apdqN[M], roeN[NR];
        {
          double ql[M], qr[M];
#pragma unroll
          for (int m = 0; m < M; ++m) { ql[m] = q1[rot<M>(m)]; qr[m] = qN[rot<M>(m)]; }
          RP::normal(P.rp, 1, ql, qr, c1, cN, waveN, sN, amdqN, apdqN, roeN);
        }
        if (y_cfl_col && (J + 2 >= 2) && (J + 2 <= P.my + 2)) cfl = cfl_update<RP>(cfl, dtdy, sN);   // flux2.f90:111-119
        double dotn[W];
        {
          double waveC[W][M], wnorm2[W];
#pragma unroll
          for (int mw = 0; mw < W; ++mw) {
            wnorm2[mw] = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) waveC[mw][m] = syw[(mw * M + m) * NT];
          }
          wave_dots<RP, true>(waveC, waveN, wnorm2, dotn);
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
          for (int m = 0; m < M; ++m) {
            qadd[m] = (0.0 - dtdy * yI_apdq[m]) - dtdy * amdqC[m];          // flux2.f90:105-108 (original fluctuations)
            yI_apdq[m] = apdqC[m];
          }
          double cqxx[M], fadd[M];
#ifndef CLAW_STRICT
          limited_second_order<RP, LIM>(P, dtdy, waveC, sC, wnorm2, dotl, dotn, cqxx);
#else
          limit_waves<RP, LIM, false>(P, waveC, sC, wnorm2, dotl, dotn);
          second_order<RP>(dtdy, waveC, sC, cqxx);
#endif
#pragma unroll
          for (int m = 0; m < M; ++m) {
            fadd[m] = 0.0 + 0.5 * cqxx[m];
            amdqC[m] = amdqC[m] + cqxx[m]; apdqC[m] = apdqC[m] - cqxx[m];   // flux2.f90:167-173
          }
          double bmA[M], bpA[M], bmP[M], bpP[M];
          RP::transverse(P.rp, 1, roeC, amdqC, bmA, bpA);                   // flux2.f90:181
          RP::transverse(P.rp, 1, roeC, apdqC, bmP, bpP);                   // flux2.f90:194
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const double g1 = (0.0 - hdy * bmA[m]) - hdy * yI_bm[m];        // flux2.f90:185-203
            const double g2 = (0.0 - hdy * bpA[m]) - hdy * yI_bp[m];
            ycomb[rot<M>(m)] = (qadd[m] - dtdy * (fadd[m] - yI_fadd[m])) - dtdx * (g2 - g1);   // step2.f90:223-225
            gy1[rot<M>(m)] = g1;
            gy2[rot<M>(m)] = g2;
            yI_fadd[m] = fadd[m]; yI_bm[m] = bmP[m]; yI_bp[m] = bpP[m];
          }
        }
        // the new interface becomes the pending one
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
        // hand cell (I, J) over to warp X
        double* e = sex + ((J & 1) * 3 * M) * NT;
#pragma unroll
        for (int m = 0; m < M; ++m) { e[m * NT] = ycomb[m]; e[(M + m) * NT] = gy1[m]; e[(2 * M + m) * NT] = gy2[m]; }
#pragma unroll
        for (int m = 0; m < M; ++m) q1[m] = qN[m];
#pragma unroll
        for (int k = 0; k < NC; ++k) c1[k] = cN[k];
      }
      __syncthreads();
    }
  } else {
    // =============================== warp X ===============================================
    const bool x_cfl_col = (lane >= 1) && (I >= 2) && (I <= P.mx + 2);
    const bool store_col = (lane >= 2) && (lane <= 29) && (I <= P.mx + 1);
    const int ll = lane > 0 ? -1 : 0, lr = lane < 31 ? 1 : 0;     // neighbour offsets in the exchange buffer
    double qn0[M], qn1[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { qn0[m] = 0.0; qn1[m] = 0.0; }

    for (int J = Jb - 4; J < Je; ++J) {                    // iteration k of warp Y is step J+1 there
      if (J >= Jb - 3) {
        double q1[M], c1[NC1], qN[M];
        {
          const double* r1 = srow + (((J + 1) & 3) * (M + NC)) * NT;
          const double* r2 = srow + (((J + 2) & 3) * (M + NC)) * NT;
#pragma unroll
          for (int m = 0; m < M; ++m) { q1[m] = r1[m * NT]; qN[m] = r2[m * NT]; }
#pragma unroll
          for (int k = 0; k < NC; ++k) c1[k] = r1[(M + k) * NT];
        }
        double qn2[M];
        double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
        {
          double ql[M], cl[NC1];
#pragma unroll
          for (int m = 0; m < M; ++m) ql[m] = shfl_left(q1[m]);
#pragma unroll
          for (int k = 0; k < NC; ++k) cl[k] = shfl_left(c1[k]);
          RP::normal(P.rp, 0, ql, q1, cl, c1, wave, s, amdq, apdq, roe);    // flux2.f90:101
        }
        {                                                                    // flux2.f90:127
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
#ifndef CLAW_STRICT
          limited_second_order<RP, LIM>(P, dtdx, wave, s, wnorm2, dotl, dotr, cqxx);
#else
          limit_waves<RP, LIM, false>(P, wave, s, wnorm2, dotl, dotr);
          second_order<RP>(dtdx, wave, s, cqxx);                             // flux2.f90:134-146
#endif
        }
#pragma unroll
        for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
        if (x_cfl_col && (J + 1 >= 1) && (J + 1 <= P.my + 2)) cfl = cfl_update<RP>(cfl, dtdx, s);   // flux2.f90:111-119
        double bmA[M], bpA[M], bmP[M], bpP[M];
        {
          double tm[M], tp[M];
#pragma unroll
          for (int m = 0; m < M; ++m) { tm[m] = amdq[m] + cqxx[m]; tp[m] = apdq[m] - cqxx[m]; }   // flux2.f90:167-173
          RP::transverse(P.rp, 0, roe, tm, bmA, bpA);                        // flux2.f90:181
          RP::transverse(P.rp, 0, roe, tp, bmP, bpP);                        // flux2.f90:194
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const double amdqR = shfl_right(amdq[m]), faddR = shfl_right(fadd[m]);
          const double bmAR = shfl_right(bmA[m]), bpAR = shfl_right(bpA[m]);
          const double qadd = (0.0 - dtdx * apdq[m]) - dtdx * amdqR;         // flux2.f90:105-108
          const double g1 = (0.0 - hdx * bmAR) - hdx * bmP[m];              // flux2.f90:185-203
          const double g2 = (0.0 - hdx * bpAR) - hdx * bpP[m];
          qn0[m] = qn0[m] - dtdy * g1;                                       // step2.f90:139-147, rows J, J+1, J+2
          qn1[m] = ((qn1[m] + qadd) - dtdx * (faddR - fadd[m])) - dtdy * (g2 - g1);
          qn2[m] = qN[m] + dtdy * g2;
        }
        // cell (I, J) is complete: step2.f90:221-229, columns I-1, I, I+1 (warp Y's vectors of step J)
        const double* e = sex + ((J & 1) * 3 * M) * NT;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          double r = qn0[m] + dtdx * e[(2 * M + m) * NT + ll];
          r = r + e[m * NT];
          r = r - dtdx * e[(M + m) * NT + lr];
          if (store_col && J >= Jb) P.qout[m * P.plane + (long long)J * P.pitch + I] = r;
        }
#pragma unroll
        for (int m = 0; m < M; ++m) { qn0[m] = qn1[m]; qn1[m] = qn2[m]; }
      }
      __syncthreads();
    }
  }
  block_max_cfl<64>(cfl, P.cfl_bits, sred);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
