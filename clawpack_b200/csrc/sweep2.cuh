// Fused 2-D classic wave-propagation sweep for sm_100a (fp64, no tensor cores: this is a
// stencil bound by the FP64 pipe).
//
// One kernel performs what the reference does in step2.f90 + flux2.f90 + limiter.f90 +
// philim.f90 + rpn2/rpt2 (unsplit), or in two step2ds.f90 calls (dimension-split):
// normal Riemann solve, CFL max, wave limiting, second-order correction, transverse
// propagation and the conservative update, reading q once and writing q once.
//
// Decomposition ("strip marching"): a CTA owns a strip of NT-4 interior columns plus a 2-cell
// halo on either side (one thread per column) and marches through a segment of rows.  At
// march step J each thread
//   * loads q(I, J+2) (coalesced, SoA planes),
//   * solves the y-interface J+2 and finishes y-interface J+1 (its limiter needs the waves of
//     interfaces J, J+1, J+2: the same thread produced them on consecutive steps, they stay
//     in registers / thread-private shared-memory slots),
//   * runs the x-pipeline of row J+1 (neighbouring interfaces live in neighbouring threads:
//     exchanged through shared memory),
//   * gathers the six contributions of cell (I, J) in the reference's accumulation order
//     (step2.f90:139-147 rows j-1, j, j+1; then :221-229 columns i-1, i, i+1) and stores it.
// The scatter of the reference (each pencil adds into three rows of qnew) becomes a gather,
// so no atomics and no second pass over q are needed; redundant work is limited to the
// 4 halo columns per strip and 3-4 warm-up rows per segment.
//
// Index convention: array coordinates (I, J) = Fortran (i + 1, j + 1) with num_ghost = 2;
// interior cells are I in [2, mx+1]; interface I separates cells I-1 and I.
#pragma once
#include "rp2.cuh"

namespace clawcuda {

#ifndef CLAWCUDA_SWEEP2_PARAMS
#define CLAWCUDA_SWEEP2_PARAMS
// Device-resident step control (clawcuda.cu: ctl_begin_kernel / ctl_end_kernel): the state of Solver.evolve_to_time
// (pyclaw/src/pyclaw/solver.py:524-666) -- t, dt, the CFL number of the running attempt, which of the two q buffers is
// current, whether the loop has ended -- lives at the start of the work buffer, so that the host can queue step
// attempts without reading the CFL number back: the kernels of attempt n+1 pick up dt and the current buffer that the
// control kernel of attempt n decided on, and turn into no-ops once `done` is set.  Every field is 8 bytes (the host
// mirrors the block as an array of 64-bit words).
struct StepCtl {
  unsigned long long cfl_bits;     // [0]  CFL slot of the running attempt (atomicMax target; all-reduced across ranks)
  double t, dt, dt_max, cfl_desired, cfl_max, cfl_cached;      // [1..6]
  double tstart, tend, told;                                   // [7..9]
  double dtmin, dtmax, cflmax;                                 // [10..12] solver.status
  long long numsteps;              // [13] status['numsteps']
  long long cur;                   // [14] index of the current q buffer
  long long done;                  // [15] 1: the attempts still queued are no-ops
  long long error;                 // [16] 1: CFL too large with a fixed dt, 2: max_steps reached before tend
  long long take_one, dt_variable, have_dt_old;                // [17..19]
  long long n, num_steps, max_steps;                           // [20..22] loop counter, accepted steps of this call
  long long attempts;              // [23] attempts logged so far (log ring index)
  long long src_done;              // [24] 1: the attempt's trailing source step must be skipped (cfl >= cfl_max)
  long long pad[7];
};
// layout version of StepCtl / Sweep2Params / LaunchInfo: a Riemann-solver plug-in (rp_plugin.cu) compiled against another
// version is refused by claw_rp_plugin_load
#define CLAW_PLUGIN_ABI 5
static_assert(sizeof(StepCtl) == 32 * 8, "StepCtl is mirrored by the host as 32 words");

// buffer roles of a launch that runs under a StepCtl: 0 = current q, 1 = alternate q, 2 = third buffer (3-D split)
__device__ __forceinline__ double* ctl_buffer(const StepCtl* ctl, double* const* qbuf, int role) {
  const int c = (int)ctl->cur;
  return role == 2 ? qbuf[2] : qbuf[role == 0 ? c : (c ^ 1)];
}

struct Sweep2Params {
  const double* qin;        // SoA: plane m at qin + m*plane, row J at + J*pitch
  double* qout;
  unsigned long long* cfl_bits;   // running max of the CFL number (bit pattern of a double >= 0)
  int mx, my;               // interior cells
  int pitch;                // row pitch in doubles
  long long plane;          // plane stride in doubles
  double dtdx, dtdy;
  int order;                // method(2)
  int mtrans;               // method(3): 0, 1, 2 (unsplit kernel)
  int limit;                // any(mthlim > 0)
  int mthlim[8];
  int seg_rows;             // rows per march segment
  int j0, j1;               // output rows [j0, j1) of this launch (whole grid: [2, my+2); x-only sweep: [0, my+4));
                            // a sub-range lets the host pipeline transfers against compute chunk by chunk
  int do_x, do_y;           // split kernel: which sweeps to perform (fused step: both)
  RpParams rp;
  // aux-reading solvers / capacity functions / f-waves (sweep2_aux.cuh)
  const double* aux;        // SoA planes like q (same pitch / plane stride), ghost cells filled
  int maux;                 // planes present
  int capa;                 // method(6) = index_capa + 1 (1-based aux plane of the capacity function), 0 = none
  int fwave;                // use_fwave
  // device-resident step control (null: qin / qout / dtdx / dtdy above are final)
  const StepCtl* ctl;
  double* qbuf[3];          // the two ping-pong q buffers (+ unused third): qin = qbuf[cur], qout = qbuf[cur ^ 1]
  double delta[2];          // dx, dy: dtdx = ctl->dt / dx as step2.f90:77-78 computes it
  int tune;                 // A/B switches of the launch code (env CLAW_TUNE; 0 = defaults): bit 0 = warm-up rows of a
                            // march segment never take the uniform-region fast path; bit 1 = no tile ordering;
                            // bit 5 = tile ordering on launches of any length
  // tile ordering of the kernels with the uniform-region fast path (sweep2_ws.cuh, HotTiles): device buffer, capacity of
  // one tile list, capacity of one flag array, grid rows of "hot" CTAs in front of the regular ones (0: off)
  unsigned* hot;
  int hot_cap, hot_tiles, hot_rows;
  unsigned hot_epoch;       // launch number when the launch does not run under a StepCtl
  int dense_pref;           // 1: the launch code takes the kernel WITHOUT the uniform-region fast path (the previous launches
                            // reported most tiles heavy; clawcuda.cu: dense_update).  Last member: the offsets of the others
                            // are those of the profiled builds.
};
#endif

// Prologue of every 2-D sweep kernel: the launch parameters, or -- under a StepCtl -- what the control kernel of the
// previous attempt left behind.  Returns false when the kernel has nothing to do (the step loop has ended).
#define CLAW_RESOLVE_STEP2(P, qin_, qout_, dtdx_, dtdy_)                                  \
  const double* qin_ = (P).qin;                                                           \
  double* qout_ = (P).qout;                                                               \
  double dtdx_ = (P).dtdx, dtdy_ = (P).dtdy;                                              \
  if ((P).ctl != nullptr) {                                                               \
    if ((P).ctl->done) return;                                                            \
    qin_ = ctl_buffer((P).ctl, (P).qbuf, 0);                                              \
    qout_ = ctl_buffer((P).ctl, (P).qbuf, 1);                                             \
    const double dt_ = (P).ctl->dt;                                                       \
    dtdx_ = dt_ / (P).delta[0];                                                           \
    dtdy_ = dt_ / (P).delta[1];                                                           \
  }

namespace CLAW_NS {

// limiter.f90:40: a wave of zero norm is left alone.  Fast mode widens "zero" to wnorm2 <= 2^-900 (waves below
// ~1e-135, e.g. the exponentially decaying numerical tail ahead of a front): the approximate-reciprocal seed flushes
// denormal inputs to zero and would turn the ratio into NaN, and a wave that small cannot move q by 1e-12.
__device__ __forceinline__ bool wave_is_zero(double wnorm2) {
#ifdef CLAW_STRICT
  return wnorm2 == 0.0;
#else
  return !(wnorm2 > 0x1p-900);
#endif
}

__device__ __forceinline__ double philim(double a, double b, int meth) {
  // philim.f90:14-35
#ifdef CLAW_STRICT
  const double r = ddiv(b, a);
#else
  const double r = b * rcp_fast(a);              // fast mode: ~2 ulp instead of the correctly rounded quotient
#endif
  switch (meth) {
    case 1: return fmax2(0.0, fmin2(1.0, r));
    case 2: return fmax2(fmax2(0.0, fmin2(1.0, 2.0 * r)), fmin2(2.0, r));
    case 3: return ddiv(r + fabs(r), 1.0 + fabs(r));
    case 4: { const double c = (1.0 + r) / 2.0; return fmax2(0.0, fmin2(fmin2(c, 2.0), 2.0 * r)); }
    case 5: return r;
  }
  return 0.0;
}

// natural <-> rotated component index (swap of the two momenta for a y-pencil)
template <int M> __host__ __device__ constexpr int rot(int m) {
  return (M >= 3 && m == 1) ? 2 : ((M >= 3 && m == 2) ? 1 : m);
}

// limiter.f90:33-38: wnorm2 and the dot product with the right/upper neighbour.  The sums run
// over the components in the reference's NATURAL order (m = 1..meqn) -- for a y-pencil the waves
// are held in the rotated frame, so the index is mapped back (floating-point sums do not commute
// across three terms).
template <class RP, bool ROTATED>
__device__ __forceinline__ void wave_dots(const double (*wave)[RP::MEQN], const double (*waveR)[RP::MEQN],
                                          double* wnorm2, double* dotr) {
#pragma unroll
  for (int mw = 0; mw < RP::MWAVES; ++mw) {
    double wn = 0.0, dr = 0.0;
#pragma unroll
    for (int mn = 0; mn < RP::MEQN; ++mn) {
      const int m = ROTATED ? rot<RP::MEQN>(mn) : mn;
      wn = (mn == 0) ? from0(wave[mw][m] * wave[mw][m]) : wn + wave[mw][m] * wave[mw][m];
      dr = (mn == 0) ? from0(wave[mw][m] * waveR[mw][m]) : dr + wave[mw][m] * waveR[mw][m];
    }
    wnorm2[mw] = wn; dotr[mw] = dr;
  }
}

// limiter.f90:39-50 for one interface, given the dot products with both neighbours
template <class RP, int LIM = 0, bool BF = false>
__device__ __forceinline__ void limit_waves(const Sweep2Params& P, double (*wave)[RP::MEQN], const double* s,
                                            const double* wnorm2, const double* dotl, const double* dotr) {
#pragma unroll
  for (int mw = 0; mw < RP::MWAVES; ++mw) {
    const int meth = LIM ? LIM : P.mthlim[mw];
    if (meth == 0) continue;
    double wl;
#ifndef CLAW_STRICT
    if constexpr (BF) {
      // branch-free (fast mode only): the divisions of the wave families become independent chains the scheduler
      // can overlap (a zero wave yields NaN here and is left untouched by the select)
      wl = philim(wnorm2[mw], (s[mw] > 0.0) ? dotl[mw] : dotr[mw], meth);
      wl = wave_is_zero(wnorm2[mw]) ? 1.0 : wl;
    } else
#endif
    {
      if (wave_is_zero(wnorm2[mw])) continue;
      wl = philim(wnorm2[mw], (s[mw] > 0.0) ? dotl[mw] : dotr[mw], meth);
    }
#pragma unroll
    for (int m = 0; m < RP::MEQN; ++m) wave[mw][m] = wl * wave[mw][m];
  }
}

// flux2.f90:134-146 (use_fwave = .false.): cqxx of one interface
template <class RP>
__device__ __forceinline__ void second_order(double dtd, const double (*wave)[RP::MEQN], const double* s,
                                             double* cqxx) {
  const double dtdxave = 0.5 * (dtd + dtd);
#pragma unroll
  for (int m = 0; m < RP::MEQN; ++m) {
    double c = 0.0;
#pragma unroll
    for (int mw = 0; mw < RP::MWAVES; ++mw) {
      const double t = fabs(s[mw]) * (1.0 - fabs(s[mw]) * dtdxave) * wave[mw][m];
      c = (mw == 0) ? from0(t) : c + t;
    }
    cqxx[m] = c;
  }
}

// Fast mode only: limiter.f90:39-50 + flux2.f90:134-146 fused.  After the limiter the waves are used by nothing but
// cqxx, so phi is folded into the per-family coefficient |s|(1 - |s| dtdx) instead of rescaling MEQN components of
// every wave (W multiplies instead of W*MEQN; the products are associated differently than in the reference).
template <class RP, int LIM>
__device__ __forceinline__ void limited_second_order(const Sweep2Params& P, double dtd, const double (*wave)[RP::MEQN],
                                                     const double* s, const double* wnorm2, const double* dotl,
                                                     const double* dotr, double* cqxx) {
  double coef[RP::MWAVES];
#pragma unroll
  for (int mw = 0; mw < RP::MWAVES; ++mw) {
    const double as = fabs(s[mw]);
    double c = as * (1.0 - as * dtd);
    const int meth = LIM ? LIM : P.mthlim[mw];
    if (meth != 0 && !wave_is_zero(wnorm2[mw])) c *= philim(wnorm2[mw], (s[mw] > 0.0) ? dotl[mw] : dotr[mw], meth);
    coef[mw] = c;
  }
#pragma unroll
  for (int m = 0; m < RP::MEQN; ++m) {
    double c = coef[0] * wave[0][m];
#pragma unroll
    for (int mw = 1; mw < RP::MWAVES; ++mw) c = fma(coef[mw], wave[mw][m], c);
    cqxx[m] = c;
  }
}

// flux2.f90:111-119 for one interface: cfl = max(cfl, dtd*s, -dtd*s) over the wave families.  dtd > 0, so
// max(dtd*s, -dtd*s) = dtd*|s| bit for bit; fast mode uses that and, for solvers whose speeds are ordered
// (s[0] = smallest, s[RP::S_HI] = largest: all of rp2.cuh), only the two extreme speeds.
template <class RP>
__device__ __forceinline__ double cfl_update(double cfl, double dtd, const double* s) {
#ifdef CLAW_STRICT
#pragma unroll
  for (int mw = 0; mw < RP::MWAVES; ++mw) cfl = fmax2(fmax2(cfl, dtd * s[mw]), -dtd * s[mw]);
  return cfl;
#else
  const double a = fabs(s[0]), b = fabs(s[RP::S_HI]);
  return fmax2(cfl, dtd * ((b > a) ? b : a));
#endif
}

// ordered speeds s[0] <= ... <= s[W-1] (the 3-D solvers of rp3.cuh)
template <int W>
__device__ __forceinline__ double cfl_update(double cfl, double dtd, const double* s) {
#ifdef CLAW_STRICT
#pragma unroll
  for (int mw = 0; mw < W; ++mw) cfl = fmax2(fmax2(cfl, dtd * s[mw]), -dtd * s[mw]);
  return cfl;
#else
  const double a = fabs(s[0]), b = fabs(s[W - 1]);
  return fmax2(cfl, dtd * ((b > a) ? b : a));
#endif
}

__device__ __forceinline__ void atomic_max_cfl(unsigned long long* addr, double v) {
  // v >= 0: the IEEE bit patterns of non-negative doubles are ordered like unsigned integers
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

template <int NT> __device__ __forceinline__ void block_max_cfl(double cfl, unsigned long long* addr, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cfl = fmax(cfl, __shfl_xor_sync(0xffffffffu, cfl, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

template <class RP, int NT> struct Smem2 {
  static constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, NC = RP::NCELL;
  static constexpr int XW = (W * M > 4 * M) ? W * M : 4 * M;
  static constexpr int QC = (M + NC > W) ? M + NC : W;
  // doubles per thread column
  static constexpr int OFF_Q = 0;               // [M+NC]  q and per-cell data of the x-row (neighbour access);
  static constexpr int OFF_DOT = OFF_Q;         // [W]     ... then the dotr exchange (q slot is free after barrier B)
  static constexpr int OFF_XW = OFF_Q + QC;     // [XW]    x waves, then x outputs (aliased)
  static constexpr int OFF_GY = OFF_XW + XW;    // [2][2M] y gadd exchange, double-buffered by step parity
  static constexpr int OFF_YW = OFF_GY + 4 * M; // [W*M]   private: waves of the pending y-interface
  static constexpr int OFF_YA = OFF_YW + W * M; // [2M+NR+W] private: amdq, apdq, roe, dotl
  static constexpr int PER_THREAD = OFF_YA + 2 * M + NR + W;
  static constexpr size_t BYTES = (size_t)PER_THREAD * NT * sizeof(double) + 64 * sizeof(double);
};

// ---- x-pipeline pieces shared by the unsplit and split kernels ------------------------

// normal solve + (limiter via neighbour exchange) + second-order correction at my left face.
// On return: wave limited, s, amdq/apdq ORIGINAL fluctuations, cqxx, fadd (= 0.5 cqxx or 0).
// Contains __syncthreads(): must be called by all threads of the CTA.
template <class RP, int NT, int LIM = 0>
__device__ __forceinline__ void x_face(const Sweep2Params& P, double dtdx, const double* q, const double* c, int t, int tl, int tr,
                                       bool second, bool limit, double* sq, double* sxw, double* sdot,
                                       double (*wave)[RP::MEQN], double* s, double* amdq, double* apdq,
                                       double* roe, double* cqxx, double* fadd) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NC = RP::NCELL;
#pragma unroll
  for (int m = 0; m < M; ++m) sq[m * NT] = q[m];
#pragma unroll
  for (int k = 0; k < NC; ++k) sq[(M + k) * NT] = c[k];
  __syncthreads();                                                        // (A)
  double ql[M], cl[NC > 0 ? NC : 1];
#pragma unroll
  for (int m = 0; m < M; ++m) ql[m] = sq[m * NT + tl];
#pragma unroll
  for (int k = 0; k < NC; ++k) cl[k] = sq[(M + k) * NT + tl];
  RP::normal(P.rp, 0, ql, q, cl, c, wave, s, amdq, apdq, roe);           // flux2.f90:101
#pragma unroll
  for (int m = 0; m < M; ++m) { cqxx[m] = 0.0; fadd[m] = 0.0; }
  if (second) {
    if (limit) {                                                          // flux2.f90:127
#pragma unroll
      for (int mw = 0; mw < W; ++mw)
#pragma unroll
        for (int m = 0; m < M; ++m) sxw[(mw * M + m) * NT] = wave[mw][m];
      __syncthreads();                                                    // (B)
      double wnorm2[W], dotr[W], dotl[W];
      {
        double waveR[W][M];
#pragma unroll
        for (int mw = 0; mw < W; ++mw)
#pragma unroll
          for (int m = 0; m < M; ++m) waveR[mw][m] = sxw[(mw * M + m) * NT + tr];
        wave_dots<RP, false>(wave, waveR, wnorm2, dotr);
      }
#pragma unroll
      for (int mw = 0; mw < W; ++mw) sdot[mw * NT] = dotr[mw];
      __syncthreads();                                                    // (C)
#pragma unroll
      for (int mw = 0; mw < W; ++mw) dotl[mw] = sdot[mw * NT + tl];
      limit_waves<RP, LIM>(P, wave, s, wnorm2, dotl, dotr);
    }
    second_order<RP>(dtdx, wave, s, cqxx);                                // flux2.f90:134-146
#pragma unroll
    for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
  }
}

// =====================================================================================
// Unsplit step (step2.f90 semantics), method(3) in {0,1,2}
// =====================================================================================
// SPEC bit 0: compile-time specialisation for the headline method set -- order 2, every wave family limited with
//   the same limiter id SPEC >> 4 (4 = MC), method(3) = 2 -- which removes the mode branches (and the register
//   moves / zero fills at their merge points) from the march loop and lets the warm-up rows run the full body
//   (their results never reach an output row, see below); SPEC = 0 is the generic kernel.
// SPEC bit 1: fetch row J+3 one march step ahead.
template <class RP, int NT, int MINB = 1, int SPEC = 0>
__global__ void __launch_bounds__(NT, MINB) sweep2_unsplit(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE;
  constexpr bool SP = (SPEC & 1) != 0, PF = (SPEC & 2) != 0;
  constexpr int NC = RP::NCELL, NC1 = NC > 0 ? NC : 1;
  constexpr int LIM = SP ? (SPEC >> 4) : 0;
  using S = Smem2<RP, NT>;
  constexpr int TX = NT - 4;
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x;
  double* const sq = smem + S::OFF_Q * NT + t;
  double* const sxw = smem + S::OFF_XW * NT + t;
  double* const sdot = smem + S::OFF_DOT * NT + t;
  double* const sgy0 = smem + S::OFF_GY * NT + t;
  double* const syw = smem + S::OFF_YW * NT + t;
  double* const sya = smem + S::OFF_YA * NT + t;
  double* const sred = smem + S::PER_THREAD * NT;

  const int nx = P.mx + 4;
  const int I = blockIdx.x * TX + t;                       // my column (array coords): 2 + bx*TX - 2 + t
  const int Ic = I < nx ? I : nx - 1;                      // clamped for loads
  const int Jb = P.j0 + blockIdx.y * P.seg_rows;           // first output row
  const int Je = min(Jb + P.seg_rows, P.j1);               // one past the last output row
  const int tl = t > 0 ? -1 : 0;                           // neighbour offsets (clamped at the strip edge)
  const int tr = t < NT - 1 ? 1 : 0;

  CLAW_RESOLVE_STEP2(P, qin, qout, dtdx, dtdy);
  const double hdx = 0.5 * dtdx, hdy = 0.5 * dtdy;
  const bool second = SP ? true : (P.order != 1);
  const bool limit = SP ? true : (second && (P.limit != 0));
  const int mt = SP ? 2 : P.mtrans;

  const double* qcol = qin + Ic;
  double q1[M], qn0[M], qn1[M], c1[NC1];
  double yI_apdq[M], yI_fadd[M], yI_bm[M], yI_bp[M];      // finished y-interface J (rotated frame)
#pragma unroll
  for (int m = 0; m < M; ++m) {
    q1[m] = qcol[m * P.plane + (long long)(Jb - 2) * P.pitch];
    qn0[m] = 0.0; qn1[m] = 0.0; yI_apdq[m] = 0.0; yI_fadd[m] = 0.0; yI_bm[m] = 0.0; yI_bp[m] = 0.0;
  }
  if constexpr (NC > 0) RP::cell(P.rp, q1, c1);
  double cfl = 0.0;
  const bool x_cfl_col = (t >= 1) && (I >= 2) && (I <= P.mx + 2);
  const bool y_cfl_col = (I >= 1) && (I <= P.mx + 2);
  const bool store_col = (t >= 2) && (t <= NT - 3) && (I <= P.mx + 1);

  // Warm-up rows (J < Jb).  The generic kernel skips the stages whose inputs do not exist yet.  The specialised
  // kernel runs the full body from the first step: what the skipped stages would have produced only feeds cells
  // J < Jb (never stored) -- row Jb's accumulator is created fresh at step Jb-2 (qn2 = qN + ...), the carried
  // y-interface data are overwritten at step Jb-1 -- so the stored values are bit-identical.
  double qP[M];
  if (PF) {
#pragma unroll
    for (int m = 0; m < M; ++m) qP[m] = qcol[m * P.plane + (long long)(Jb - 1) * P.pitch];
  }
  const int Jlast = P.my + 3;
  for (int J = Jb - 3; J < Je; ++J) {
    double* const sgy = sgy0 + (J & 1) * (2 * M * NT);
    // ---- load row J+2 ---------------------------------------------------------------
    double qN[M];
    if (PF) {
      const int Jn = (J + 3 < Jlast) ? J + 3 : Jlast;
#pragma unroll
      for (int m = 0; m < M; ++m) { qN[m] = qP[m]; qP[m] = qcol[m * P.plane + (long long)Jn * P.pitch]; }
    } else {
#pragma unroll
      for (int m = 0; m < M; ++m) qN[m] = qcol[m * P.plane + (long long)(J + 2) * P.pitch];
    }
    double cN[NC1];
    if constexpr (NC > 0) RP::cell(P.rp, qN, cN);

    // =========================== y-pipeline ==========================================
    double ycomb[M];                                       // (qadd - dtdy*df) - dtdx*(g2 - g1), natural frame
#pragma unroll
    for (int m = 0; m < M; ++m) ycomb[m] = 0.0;
    {
      double waveN[W][M], sN[W], amdqN[M], apdqN[M], roeN[NR];
      {
        double ql[M], qr[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { ql[m] = q1[rot<M>(m)]; qr[m] = qN[rot<M>(m)]; }
        RP::normal(P.rp, 1, ql, qr, c1, cN, waveN, sN, amdqN, apdqN, roeN);
      }
      if (y_cfl_col && (J + 2 >= 2) && (J + 2 <= P.my + 2)) {             // flux2.f90:111-119
        cfl = cfl_update<RP>(cfl, dtdy, sN);
      }
      double dotn[W];
#pragma unroll
      for (int mw = 0; mw < W; ++mw) dotn[mw] = 0.0;
      if (SP || J >= Jb - 2) {
        double waveC[W][M], wnorm2[W];
#pragma unroll
        for (int mw = 0; mw < W; ++mw) {
          wnorm2[mw] = 0.0;
#pragma unroll
          for (int m = 0; m < M; ++m) waveC[mw][m] = syw[(mw * M + m) * NT];
        }
        if (limit) wave_dots<RP, true>(waveC, waveN, wnorm2, dotn);
        if (SP || J >= Jb - 1) {
          double sC[W], amdqC[M], apdqC[M], roeC[NR], dotl[W];
#pragma unroll
          for (int m = 0; m < M; ++m) { amdqC[m] = sya[m * NT]; apdqC[m] = sya[(M + m) * NT]; }
#pragma unroll
          for (int k = 0; k < NR; ++k) roeC[k] = sya[(2 * M + k) * NT];
#pragma unroll
          for (int mw = 0; mw < W; ++mw) dotl[mw] = sya[(2 * M + NR + mw) * NT];
          RP::speeds(P.rp, 1, roeC, sC);
          // first-order increment of cell J uses the ORIGINAL fluctuations (flux2.f90:105-108)
          double qadd[M];
#pragma unroll
          for (int m = 0; m < M; ++m) {
            qadd[m] = (0.0 - dtdy * yI_apdq[m]) - dtdy * amdqC[m];
            yI_apdq[m] = apdqC[m];                                         // carry: original apdq of interface J+1
          }
          double fadd[M];
#pragma unroll
          for (int m = 0; m < M; ++m) fadd[m] = 0.0;
          if (second) {
            if (limit) limit_waves<RP, LIM>(P, waveC, sC, wnorm2, dotl, dotn);
            double cqxx[M];
            second_order<RP>(dtdy, waveC, sC, cqxx);
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
            RP::transverse(P.rp, 1, roeC, amdqC, bmA, bpA);               // flux2.f90:181
            RP::transverse(P.rp, 1, roeC, apdqC, bmP, bpP);               // flux2.f90:194
          }
          // cell J of this y-pencil: interface J (carried) and interface J+1 (just finished)
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const double g1 = (0.0 - hdy * bmA[m]) - hdy * yI_bm[m];      // flux2.f90:185-203
            const double g2 = (0.0 - hdy * bpA[m]) - hdy * yI_bp[m];
            // step2.f90:223-225: qadd - dtdy*(fadd(j+1)-fadd(j)) - dtdx*(gadd(2)-gadd(1))
            ycomb[rot<M>(m)] = (qadd[m] - dtdy * (fadd[m] - yI_fadd[m])) - dtdx * (g2 - g1);
            sgy[rot<M>(m) * NT] = g1;
            sgy[(M + rot<M>(m)) * NT] = g2;
            yI_fadd[m] = fadd[m]; yI_bm[m] = bmP[m]; yI_bp[m] = bpP[m];
          }
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
    }

    // =========================== x-pipeline on row J+1 ===============================
    double qn2[M];
#pragma unroll
    for (int m = 0; m < M; ++m) qn2[m] = qN[m];
    if (SP || J >= Jb - 2) {
      double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
      x_face<RP, NT, LIM>(P, dtdx, q1, c1, t, tl, tr, second, limit, sq, sxw, sdot, wave, s, amdq, apdq, roe, cqxx, fadd);
      if (x_cfl_col && (J + 1 >= 1) && (J + 1 <= P.my + 2)) {             // flux2.f90:111-119
        cfl = cfl_update<RP>(cfl, dtdx, s);
      }
      double bmA[M], bpA[M], bmP[M], bpP[M];
#pragma unroll
      for (int m = 0; m < M; ++m) { bmA[m] = 0.0; bpA[m] = 0.0; bmP[m] = 0.0; bpP[m] = 0.0; }
      if (mt > 0) {
        double tm[M], tp[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { tm[m] = amdq[m]; tp[m] = apdq[m]; }
        if (second && mt == 2) {                                           // flux2.f90:167-173
#pragma unroll
          for (int m = 0; m < M; ++m) { tm[m] = tm[m] + cqxx[m]; tp[m] = tp[m] - cqxx[m]; }
        }
        RP::transverse(P.rp, 0, roe, tm, bmA, bpA);                        // flux2.f90:181
        RP::transverse(P.rp, 0, roe, tp, bmP, bpP);                        // flux2.f90:194
      }
      // hand my left-face results to the cell on my left (sxw is free again after barrier C / A)
#pragma unroll
      for (int m = 0; m < M; ++m) {
        sxw[m * NT] = amdq[m]; sxw[(M + m) * NT] = fadd[m];
        sxw[(2 * M + m) * NT] = bmA[m]; sxw[(3 * M + m) * NT] = bpA[m];
      }
      __syncthreads();                                                     // (D) also publishes sgy
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double amdqR = sxw[m * NT + tr], faddR = sxw[(M + m) * NT + tr];
        const double bmAR = sxw[(2 * M + m) * NT + tr], bpAR = sxw[(3 * M + m) * NT + tr];
        const double qadd = (0.0 - dtdx * apdq[m]) - dtdx * amdqR;         // flux2.f90:105-108
        const double g1 = (0.0 - hdx * bmAR) - hdx * bmP[m];              // flux2.f90:185-203
        const double g2 = (0.0 - hdx * bpAR) - hdx * bpP[m];
        // step2.f90:139-147, pencil J+1: rows J, J+1, J+2
        qn0[m] = qn0[m] - dtdy * g1;
        qn1[m] = ((qn1[m] + qadd) - dtdx * (faddR - fadd[m])) - dtdy * (g2 - g1);
        qn2[m] = qN[m] + dtdy * g2;
      }
    }

    // =========================== cell (I, J) is complete =============================
    if (J >= Jb) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        // step2.f90:221-229, columns I-1, I, I+1
        double r = qn0[m] + dtdx * sgy[(M + m) * NT + tl];
        r = r + ycomb[m];
        r = r - dtdx * sgy[m * NT + tr];
        if (store_col) qout[m * P.plane + (long long)J * P.pitch + I] = r;
      }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) { qn0[m] = qn1[m]; qn1[m] = qn2[m]; q1[m] = qN[m]; }
#pragma unroll
    for (int k = 0; k < NC; ++k) c1[k] = cN[k];
  }
  block_max_cfl<NT>(cfl, P.cfl_bits, sred);
}

// =====================================================================================
// Dimension-split step: x-sweeps (step2ds ids=1) then y-sweeps (ids=2) of the x-updated data,
// fused.  do_x / do_y select a single sweep (the f2py-shaped plugin entry points).
//   x-sweep: all rows J in [0, my+3], cells I in [2, mx+1]          (step2ds.f90:96, 149-152)
//   y-sweep: all columns I in [0, mx+3], cells J in [2, my+1]       (step2ds.f90:180, 237-240)
// Segments: blockIdx.y covers output rows [Jb, Je) of [2, my+2) (fused / y-only) or of
// [0, my+4) (x-only).
// =====================================================================================
template <class RP, int NT, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) sweep2_split(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE;
  using S = Smem2<RP, NT>;
  constexpr int TX = NT - 4;
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x;
  double* const sq = smem + S::OFF_Q * NT + t;
  double* const sxw = smem + S::OFF_XW * NT + t;
  double* const sdot = smem + S::OFF_DOT * NT + t;
  double* const syw = smem + S::OFF_YW * NT + t;
  double* const sya = smem + S::OFF_YA * NT + t;
  double* const sred = smem + S::PER_THREAD * NT;

  const int nx = P.mx + 4;
  const int I = blockIdx.x * TX + t;
  const int Ic = I < nx ? I : nx - 1;
  const int tl = t > 0 ? -1 : 0;
  const int tr = t < NT - 1 ? 1 : 0;
  CLAW_RESOLVE_STEP2(P, qin, qout, dtdx, dtdy);
  const bool second = (P.order != 1);
  const bool limit = second && (P.limit != 0);
  const bool do_x = P.do_x != 0, do_y = P.do_y != 0;
  constexpr int NC = RP::NCELL, NC1 = NC > 0 ? NC : 1;
  const double* qcol = qin + Ic;
  double cfl = 0.0;
  const bool x_cfl_col = (t >= 1) && (I >= 2) && (I <= P.mx + 2);
  const bool interior_col = (I >= 2) && (I <= P.mx + 1);
  const bool x_ok = (t >= 2) && (t <= NT - 3);            // my x-updated value is complete
  // column whose x-updated data are valid in this strip: x-updated interior, or untouched ghost
  const bool col_ok = (I <= P.mx + 3) && (!do_x || !interior_col || x_ok);

  if (!do_y) {
    // ---- x-sweep only: rows [Jb, Je) of the whole array -------------------------------
    const int Jb = P.j0 + blockIdx.y * P.seg_rows;
    const int Je = min(Jb + P.seg_rows, P.j1);
    for (int J = Jb; J < Je; ++J) {
      double q[M];
#pragma unroll
      for (int m = 0; m < M; ++m) q[m] = qcol[m * P.plane + (long long)J * P.pitch];
      double cq[NC1];
      if constexpr (NC > 0) RP::cell(P.rp, q, cq);
      double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
      x_face<RP, NT>(P, dtdx, q, cq, t, tl, tr, second, limit, sq, sxw, sdot, wave, s, amdq, apdq, roe, cqxx, fadd);
      if (x_cfl_col) {
        cfl = cfl_update<RP>(cfl, dtdx, s);
      }
#pragma unroll
      for (int m = 0; m < M; ++m) { sxw[m * NT] = amdq[m]; sxw[(M + m) * NT] = fadd[m]; }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double qadd = (0.0 - dtdx * apdq[m]) - dtdx * sxw[m * NT + tr];
        const double r = (q[m] + qadd) - dtdx * (sxw[(M + m) * NT + tr] - fadd[m]);   // step2ds.f90:149-152
        if (interior_col && x_ok) qout[m * P.plane + (long long)J * P.pitch + I] = r;
      }
    }
    block_max_cfl<NT>(cfl, P.cfl_bits, sred);
    return;
  }

  const int Jb = P.j0 + blockIdx.y * P.seg_rows;
  const int Je = min(Jb + P.seg_rows, P.j1);
  double qx0[M], qx1[M], cx1[NC1];                         // x-updated rows J, J+1 (natural frame)
  double yI_apdq[M], yI_fadd[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { qx0[m] = 0.0; qx1[m] = 0.0; yI_apdq[m] = 0.0; yI_fadd[m] = 0.0; }
#pragma unroll
  for (int k = 0; k < NC1; ++k) cx1[k] = 0.0;

  for (int J = Jb - 4; J < Je; ++J) {
    // ---- x-sweep of row J+2 ----------------------------------------------------------
    double qx2[M];
#pragma unroll
    for (int m = 0; m < M; ++m) qx2[m] = qcol[m * P.plane + (long long)(J + 2) * P.pitch];
    if (do_x) {
      double cq[NC1];
      if constexpr (NC > 0) RP::cell(P.rp, qx2, cq);
      double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
      x_face<RP, NT>(P, dtdx, qx2, cq, t, tl, tr, second, limit, sq, sxw, sdot, wave, s, amdq, apdq, roe, cqxx, fadd);
      if (x_cfl_col) {
        cfl = cfl_update<RP>(cfl, dtdx, s);
      }
#pragma unroll
      for (int m = 0; m < M; ++m) { sxw[m * NT] = amdq[m]; sxw[(M + m) * NT] = fadd[m]; }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double qadd = (0.0 - dtdx * apdq[m]) - dtdx * sxw[m * NT + tr];
        const double r = (qx2[m] + qadd) - dtdx * (sxw[(M + m) * NT + tr] - fadd[m]);
        if (interior_col) qx2[m] = r;                      // ghost columns are not updated by the x-sweep
      }
    }
    // ---- y-pipeline --------------------------------------------------------------------
    double cx2[NC1];
    if constexpr (NC > 0) RP::cell(P.rp, qx2, cx2);                                // of the x-updated cell
    if (J >= Jb - 3) {
      double waveN[W][M], sN[W], amdqN[M], apdqN[M], roeN[NR];
      {
        double ql[M], qr[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { ql[m] = qx1[rot<M>(m)]; qr[m] = qx2[rot<M>(m)]; }
        RP::normal(P.rp, 1, ql, qr, cx1, cx2, waveN, sN, amdqN, apdqN, roeN);
      }
      if (col_ok && (J + 2 >= 2) && (J + 2 <= P.my + 2)) {
        cfl = cfl_update<RP>(cfl, dtdy, sN);
      }
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
            if (limit) limit_waves<RP>(P, waveC, sC, wnorm2, dotl, dotn);
            double cqxx[M];
            second_order<RP>(dtdy, waveC, sC, cqxx);
#pragma unroll
            for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
          }
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const double qadd = (0.0 - dtdy * yI_apdq[m]) - dtdy * amdqC[m];
            // step2ds.f90:237-240
            const double r = (qx0[rot<M>(m)] + qadd) - dtdy * (fadd[m] - yI_fadd[m]);
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
    for (int k = 0; k < NC; ++k) cx1[k] = cx2[k];
  }
  block_max_cfl<NT>(cfl, P.cfl_bits, sred);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
