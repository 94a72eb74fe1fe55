// 1-D classic step: one kernel does what pyclaw/src/pyclaw/classic/step1.f90 + limiter.f90 + philim.f90 + rp1 do.
//
// One thread per cell (array index I = 0 .. mx+3, two ghost cells per side); a CTA of NT threads owns NT-4 cells
// plus 2 halo cells per side; neighbouring interfaces meet through shared memory.  Thread t solves the Riemann problem
// at the LEFT face of its cell (Fortran interface i = I-1), limits it with the waves of both neighbours, forms the
// second-order flux f(i) and then updates its cell with the same association as step1.f90:93-96, 145-147:
//     q1 = (q - dtdx(i)*apdq(i)) - dtdx(i)*amdq(i+1);   qnew = q1 - dtdx(i)*(f(i+1) - f(i)).
// The reference updates q in place; here input and output are separate buffers (the halo reads of neighbouring CTAs
// must see the old values).  1-D problems are small: this kernel is about completeness and parity, not throughput.
#pragma once
#include <type_traits>
#include "sweep2.cuh"

namespace clawcuda {
namespace CLAW_NS {

// number of aux components a 1-D functor reads (0 unless it declares MAUX; see rp1.cuh)
template <class RP, class = void> struct Rp1Aux { static constexpr int N = 0; };
template <class RP> struct Rp1Aux<RP, std::void_t<decltype(RP::MAUX)>> { static constexpr int N = RP::MAUX; };

template <class RP, int NT>
__global__ void __launch_bounds__(NT) sweep1(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES;
  __shared__ double sq[M][NT], sdt[NT], sw[W][M][NT], sdot[W][NT], sam[M][NT], sf[M][NT], sred[NT / 32];
  const int t = threadIdx.x;
  const int I = blockIdx.x * (NT - 4) + t - 2;             // my cell (array index); may be outside [0, mx+3]
  const int ncell = P.mx + 4;
  const int Ic = I < 0 ? 0 : (I >= ncell ? ncell - 1 : I);
  const bool in_range = (I >= 0) && (I < ncell);
  const bool owned = (t >= 2) && (t <= NT - 3) && in_range;
  const int tl = t > 0 ? t - 1 : 0, tr = t < NT - 1 ? t + 1 : NT - 1;

  double q[M];
#pragma unroll
  for (int m = 0; m < M; ++m) q[m] = P.qin[m * P.plane + (long long)Ic * P.pitch];
  // dtdx(i) = dt/(dx*aux(capa,i)) or dt/dx (step1.f90:69-80); P.dtdx carries dt, P.dtdy carries dx here
  const double dtdx = (P.capa > 0) ? ddiv(P.dtdx, P.dtdy * P.aux[(P.capa - 1) + (long long)Ic * P.maux])
                                   : ddiv(P.dtdx, P.dtdy);
#pragma unroll
  for (int m = 0; m < M; ++m) sq[m][t] = q[m];
  sdt[t] = dtdx;
  __syncthreads();
  double ql[M];
#pragma unroll
  for (int m = 0; m < M; ++m) ql[m] = sq[m][tl];
  const double dtdx_l = sdt[tl];

  // interface at my left face: Fortran i = I-1; rp1 range i = 2-mbc .. mx+mbc  <=>  I = 1 .. mx+3
  double wave[W][M], s[W], amdq[M], apdq[M];
  if constexpr (Rp1Aux<RP>::N > 0) {                               // aux columns of the left cell and of mine
    constexpr int NA = Rp1Aux<RP>::N;
    const int Il = Ic > 0 ? Ic - 1 : 0;
    double al[NA], ar[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      al[k] = P.aux[k + (long long)Il * P.maux];
      ar[k] = P.aux[k + (long long)Ic * P.maux];
    }
    RP::normal(P.rp, ql, q, al, ar, wave, s, amdq, apdq);
  } else {
    RP::normal(P.rp, ql, q, wave, s, amdq, apdq);
  }
  const bool active = (I >= 2) && (I <= P.mx + 2) && (t >= 1);   // i = 1 .. mx+1: update / CFL / second-order range

  double cfl = 0.0;
  if (owned && active) {                                           // step1.f90:100-105
#pragma unroll
    for (int mw = 0; mw < W; ++mw) cfl = fmax2(fmax2(cfl, dtdx * s[mw]), -dtdx_l * s[mw]);
  }

  double f[M];
#pragma unroll
  for (int m = 0; m < M; ++m) f[m] = 0.0;
  if (P.order != 1) {
    if (P.limit) {                                                 // limiter.f90 (unlimited neighbours)
#pragma unroll
      for (int mw = 0; mw < W; ++mw)
#pragma unroll
        for (int m = 0; m < M; ++m) sw[mw][m][t] = wave[mw][m];
      __syncthreads();
      double wnorm2[W], dotr[W], dotl[W];
      {
        double waveR[W][M];
#pragma unroll
        for (int mw = 0; mw < W; ++mw)
#pragma unroll
          for (int m = 0; m < M; ++m) waveR[mw][m] = sw[mw][m][tr];
        wave_dots<RP, false>(wave, waveR, wnorm2, dotr);
      }
#pragma unroll
      for (int mw = 0; mw < W; ++mw) sdot[mw][t] = dotr[mw];
      __syncthreads();
#pragma unroll
      for (int mw = 0; mw < W; ++mw) dotl[mw] = sdot[mw][tl];
      if (active) limit_waves<RP, 0>(P, wave, s, wnorm2, dotl, dotr);
    }
    if (active) {                                                  // step1.f90:119-141
      const double dtdxave = 0.5 * (dtdx_l + dtdx);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        double acc = 0.0;
#pragma unroll
        for (int mw = 0; mw < W; ++mw) {
          const double lead = P.fwave ? copysign(1.0, s[mw]) : fabs(s[mw]);
          acc = acc + 0.5 * lead * (1.0 - fabs(s[mw]) * dtdxave) * wave[mw][m];
        }
        f[m] = acc;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) { sam[m][t] = amdq[m]; sf[m][t] = f[m]; }
  __syncthreads();

  // cell I (Fortran i = I-1): first statement of the forall for i = 1..mx+1, second for cells 0..mx, third 1..mx+1
  const bool right_active = (I + 1 >= 2) && (I + 1 <= P.mx + 2);
#pragma unroll
  for (int m = 0; m < M; ++m) {
    double r = q[m];
    if (active) r = r - dtdx * apdq[m];
    if (right_active) r = r - dtdx * sam[m][tr];
    if (P.order != 1 && active) r = r - dtdx * (sf[m][tr] - f[m]);
    if (owned) P.qout[m * P.plane + (long long)I * P.pitch] = r;
  }
  block_max_cfl<NT>(cfl, P.cfl_bits, sred);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
