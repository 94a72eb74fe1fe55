// Warp-synchronous variant of the fused unsplit 2-D sweep (same algorithm, same accumulation order and therefore
// the same bits as sweep2_unsplit in sweep2.cuh; see the header comment there for the march).
//
// Why: ncu on sweep2_unsplit<RpEuler4Wave2D,64,6> shows the LSU data pipe 65 % busy (shared-memory exchange of
// neighbouring interfaces: 177 eight-byte accesses per cell-step = 2 wavefronts each, plus register spills) next
// to an FP64 pipe that is 53 % busy: the kernel is co-limited by shared-memory wavefronts, not by arithmetic
// alone.  Here every warp owns its own strip (32 columns, 28 owned + 2 halo columns per side) and exchanges the
// x-neighbour data with warp shuffles -- one SHFL moves what a STS + LDS pair moved -- so the march loop has no
// __syncthreads and shared memory only holds the thread-private pending y-interface (waves, fluctuations, Roe
// data, limiter dot products) that must survive one march step.
#pragma once
#include "sweep2.cuh"

namespace clawcuda {
namespace CLAW_NS {

__device__ __forceinline__ double shfl_left(double v) {          // value of lane-1 (lane 0: its own)
  return __shfl_up_sync(0xffffffffu, v, 1);
}
__device__ __forceinline__ double shfl_right(double v) {         // value of lane+1 (lane 31: its own)
  return __shfl_down_sync(0xffffffffu, v, 1);
}

#ifdef CLAW_HOST_CHECK                      // kernel emulation (tests/host_functors/cuda_emul.h): the copy completes at once
inline void cp_async8(double* smem_dst, const double* gsrc) { *smem_dst = *gsrc; }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
template <int LEVEL> inline void prefetch_row(const double*) {}
#else
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int LEVEL> __device__ __forceinline__ void prefetch_row(const double* g) {
  if constexpr (LEVEL == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(g));
  else asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
}
#endif

template <class RP, int NW, int RD = 2> struct Smem2ws {
  static constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE;
  static constexpr int OFF_YW = 0;                  // [W*M]   waves of the pending y-interface
  static constexpr int OFF_YA = OFF_YW + W * M;     // [2M+NR+W] amdq, apdq, roe, dotl
  static constexpr int OFF_PF = OFF_YA + 2 * M + NR + W;   // [RD*M] ring of prefetched q rows (cp.async), RD stages
  static constexpr int PER_THREAD = OFF_PF + RD * M;
  static constexpr size_t BYTES = (size_t)PER_THREAD * NW * 32 * sizeof(double) + 64 * sizeof(double);
};

template <int N> __device__ __forceinline__ void cp_async_wait_keep() {   // wait until at most N groups are pending
#ifdef CLAW_HOST_CHECK
#else
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ bool same_bits(double a, double b) {
  return __double_as_longlong(a) == __double_as_longlong(b);
}
__device__ __forceinline__ unsigned long long diff_bits(double a, double b) {
  return (unsigned long long)(__double_as_longlong(a) ^ __double_as_longlong(b));
}
// the value is needed in a register HERE: keeps the compiler from sinking the assignment that produced it into the
// exits of a loop that follows (where it would be repeated on every iteration)
__device__ __forceinline__ void pin(double& v) {
#ifndef CLAW_HOST_CHECK
  asm volatile("" : "+d"(v));
#endif
}
__device__ __forceinline__ bool warp_all(bool pred) {
#ifdef CLAW_HOST_CHECK
  return emul::vote_all(pred);
#else
  return __all_sync(0xffffffffu, pred) != 0;
#endif
}

// Row ring in shared memory, thread-private: component m of ring slot s lives at spf[(m * RD + s) * NT] (component-major,
// so that the slot offset wraps with one AND for any number of components).  The tight loop below addresses it through
// the 32-bit shared window (one conversion per kernel, constant byte offsets folded into the instructions) instead of
// generic pointers (a window conversion per access).
#ifdef CLAW_HOST_CHECK
using ring_t = double*;
inline ring_t ring_of(double* p) { return p; }
template <int OFF> inline void ring_cp8(ring_t r, int off, const double* g) { r[off + OFF] = *g; }
template <int OFF> inline double ring_ld(ring_t r, int off) { return r[off + OFF]; }
#else
using ring_t = unsigned;
__device__ __forceinline__ ring_t ring_of(double* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int OFF> __device__ __forceinline__ void ring_cp8(ring_t r, int off, const double* g) {
  asm volatile("cp.async.ca.shared.global [%0+%2], [%1], 8;" ::"r"(r + 8u * (unsigned)off), "l"(g), "n"(8 * OFF) : "memory");
}
template <int OFF> __device__ __forceinline__ double ring_ld(ring_t r, int off) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(r + 8u * (unsigned)off), "n"(8 * OFF) : "memory");
  return v;
}
#endif
template <int M, int RD, int NT, int m = 0> __device__ __forceinline__ void ring_cp_row(ring_t r, int off, const double* (&g)[M]) {
  if constexpr (m < M) { ring_cp8<m * RD * NT>(r, off, g[m]); ring_cp_row<M, RD, NT, m + 1>(r, off, g); }
}
template <int M, int RD, int NT, int m = 0> __device__ __forceinline__ void ring_ld_row(ring_t r, int off, double (&q)[M]) {
  if constexpr (m < M) { q[m] = ring_ld<m * RD * NT>(r, off); ring_ld_row<M, RD, NT, m + 1>(r, off, q); }
}

// TMA form of the row ring (template flag TMA of sweep2_unsplit_ws): one bulk copy per component and row --
// cp.async.bulk global -> shared, 256 bytes = the 32 columns of the strip, issued by lane 0, completion counted in bytes
// on an mbarrier per ring slot -- instead of 32 lanes x 8-byte cp.async.  `mbar` = shared-window address of the slot's
// mbarrier; the k-th use of a slot completes phase parity k & 1.
#ifdef CLAW_HOST_CHECK
inline void tma_mbar_init(ring_t, int) {}
inline void tma_row_issue(ring_t dst, int off, const double* src, int bytes, ring_t, bool first, int total) {
  (void)first; (void)total;
  std::memcpy(dst + off, src, (size_t)bytes);
}
inline void tma_wait(ring_t, unsigned) {}
inline void warp_sync() { emul::yield_sync(); }
#else
__device__ __forceinline__ void tma_mbar_init(ring_t mbar, int n) {
  for (int k = 0; k < n; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 8u * k) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one component of a row; `first` arms the slot's mbarrier with the byte count of the whole row (`total`)
__device__ __forceinline__ void tma_row_issue(ring_t dst, int off, const double* src, int bytes, ring_t mbar, bool first,
                                              int total) {
  if (first) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst + 8u * (unsigned)off),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void tma_wait(ring_t mbar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void warp_sync() { __syncwarp(); }
#endif

// Tight loop of the uniform-region fast path: a run of rows that carry the bits of `ref` in all 32 columns of the warp.
// Per row: queue the row RD steps ahead, wait for the next row of the ring, compare it with `ref`, vote, store `sto`
// (which carries the bits of `ref`) as the output row.  J is the march step whose row has been stored last; the loop
// ends at the segment end or at the first row that differs, which is handed over in qT (return value true) for the
// general classification of step J + 1.  Ring bookkeeping (wslot, rslot: slot offsets in doubles; gfetch: row to queue
// next) is the caller's and is left as ring_fetch would have left it.
template <int M, int RD, int NT, bool TMA = false>
__device__ __forceinline__ bool quiet_run(int& J, const int Je, const double (&ref)[M], const double (&sto)[M], double (&qT)[M],
                                          double* spf, int& wslot, int& rslot, const double*& gfetch, const long long plane,
                                          const int pitch, double* orow, const bool st, [[maybe_unused]] ring_t mbar = ring_t(),
                                          [[maybe_unused]] unsigned* rphase = nullptr, [[maybe_unused]] bool live = true) {
  constexpr int WRAP = RD * NT - 1;
  const ring_t ring = ring_of(spf);
  [[maybe_unused]] const bool lane0 = (threadIdx.x & 31) == 0;
  [[maybe_unused]] const ring_t ring0 = ring_of(spf - (threadIdx.x & 31));     // lane 0's slot: start of the 32-lane row
  // running pointers per component (an invariant base + 32-bit row offset was measured to compile to more instructions:
  // at 255 registers the compiler rematerialises the plane pointers on every row)
  const double* gf[M];
  double* os[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { gf[m] = gfetch + m * plane; os[m] = orow + m * plane; }
  int nf = Je - J - RD;                                    // rows still to be queued (step Js queues row Js+RD+1 <= Je+1)
  int wo = wslot, ro = rslot;
  bool differs = false;
  const int J0 = J;
  while (J + 1 < Je) {
    if constexpr (TMA) {
      warp_sync();                                         // every lane has read the slot that is queued again
      if (nf > 0 && lane0) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          tma_row_issue(ring0, wo + m * RD * NT, gf[m], NT * 8, mbar + 8u * (unsigned)(wo / NT), m == 0, M * NT * 8);
      }
    } else {
      if (nf > 0) ring_cp_row<M, RD, NT>(ring, wo, gf);
      cp_async_commit();
    }
    --nf;
#pragma unroll
    for (int m = 0; m < M; ++m) gf[m] += pitch;
    wo = (wo + NT) & WRAP;
    if constexpr (TMA) {
      const unsigned sl = (unsigned)(ro / NT);
      tma_wait(mbar + 8u * sl, (*rphase >> sl) & 1u);
      *rphase ^= 1u << sl;
    } else {
      cp_async_wait_keep<RD - 1>();
    }
    ring_ld_row<M, RD, NT>(ring, ro, qT);
    ro = (ro + NT) & WRAP;
    unsigned long long d = 0ull;
#pragma unroll
    for (int m = 0; m < M; ++m) d |= diff_bits(ref[m], qT[m]);
    if (!warp_all(d == 0ull || !live)) { differs = true; break; }
    ++J;
    if (st) {
#pragma unroll
      for (int m = 0; m < M; ++m) *os[m] = sto[m];
    }
#pragma unroll
    for (int m = 0; m < M; ++m) os[m] += pitch;
  }
  gfetch += (long long)(J - J0 + (differs ? 1 : 0)) * pitch;
  wslot = wo; rslot = ro;
  return differs;
}

// Measured and dropped (profiles/r02_q, r02_s): a WIDE form of this loop -- 16-byte loads straight into registers (two
// LDG.128 per row for up to four components, U rows in flight, no shared memory), 16-byte stores of the 28 owned columns,
// the loop proper as a __noinline__ function with its own register allocation (39 instead of 65-75 instructions per row).
// It was slower on both kernels (8192^2 Euler 0.79 against 0.75 ms, 16384^2 shallow water 2.38 against 2.19 ms).  What
// the ncu capture of the shallow-water kernel (mio_throttle 5.5, short_scoreboard 4.1 warps per issue cycle at 4.3 TB/s)
// really showed was too many cp.async rows in flight per SM: 16 warps x 7 rows ran at 4.3 TB/s, 16 x 3 at 5.9 TB/s,
// 8 x 7 at 6.0 TB/s, 8 x 3 at 5.4 TB/s -- about 50 rows per SM is the optimum, which fixes the ring depth per kernel
// (sweep_inst.cu).

// ---------------------------------------------------------------------------------------------------------------------
// Tile ordering for the kernels with the uniform-region fast path.  There the cost of a tile (one strip x one march
// segment) depends on the data: a tile inside a uniform region streams at HBM speed, a tile full of a front costs the
// full arithmetic of every row -- 50 to 100 times more -- and the heavy tiles that happen to be dispatched last are the
// tail of the launch (measured at 8192^2, profiles/r02_n: planar shock alone +0.10 to +0.15 ms on 0.75 ms).  Fronts move
// by at most one cell per step, so last step's heavy tiles are this step's: every CTA reports whether its tile was heavy,
// and the next launch runs the reported tiles FIRST, from `hot_rows` extra grid rows in front of the regular grid (CTAs
// are dispatched in linear order); the regular CTA of such a tile exits at once.  Purely a scheduling hint: each tile is
// still processed exactly once per launch, by the same code.
//
// Device buffer (unsigned words): three tile lists used in rotation -- launch number e (the attempt counter of the
// StepCtl, or a host counter) READS list e % 3, WRITES list (e+1) % 3 and CLEARS list (e+2) % 3, so that no word is read
// and written in the same launch and no CTA has to wait for another.  [0..2] entries in list 0 / 1 / 2, [3..5] stamp of
// the grid each list was written for, [8 .. 8+3*cap) the lists, then three flag arrays (hot_tiles words each):
// flag[tile] = 1 + position in the list, 0 = not listed.  List and flag are checked against each other on both sides, so
// stale or foreign contents (a skipped launch number, another segmentation) can only cost the ordering, never a tile.
// ---------------------------------------------------------------------------------------------------------------------
// which instantiations carry the uniform-region fast path (and with it the tile ordering): shared by the kernels and the
// launch code, which adds the hot grid rows only for these
template <class RP, int SPEC, int QD> struct UnsplitQuiet {
#ifdef CLAW_STRICT
  static constexpr bool value = false;
#else
  static constexpr bool value = (QD > 0) && ((SPEC & 1) != 0) && (((SPEC >> 1) & 3) == 3) && (RP::NCELL > 0);
#endif
};
template <class RP, int PF, int LIM, int QD> struct SplitQuiet {
#ifdef CLAW_STRICT
  static constexpr bool value = false;
#else
  static constexpr bool value = (QD > 0) && (LIM > 0) && (PF != 0) && (RP::NCELL == 0);
#endif
};

struct HotTile {
  int bx, by;               // the tile this CTA works on (strip, segment)
  unsigned tile, epoch;
  bool run;                 // false: nothing to do for this CTA (empty hot slot, or the tile is run by a hot CTA)
};

__device__ __forceinline__ unsigned hot_stamp(const Sweep2Params& P, unsigned strips, unsigned nseg) {
  return ((strips << 16) ^ nseg ^ ((unsigned)P.seg_rows << 8)) | 0x80000000u;
}

__device__ __forceinline__ HotTile hot_begin(const Sweep2Params& P) {
  HotTile h;
  h.bx = (int)blockIdx.x; h.by = (int)blockIdx.y; h.tile = 0; h.epoch = 0; h.run = true;
  if (P.hot_rows <= 0) return h;
  const unsigned strips = gridDim.x, nseg = gridDim.y - (unsigned)P.hot_rows, ntiles = strips * nseg;
  const unsigned cap = (unsigned)P.hot_rows * strips;
  const unsigned* hot = P.hot;
  h.epoch = (P.ctl != nullptr) ? (unsigned)P.ctl->attempts : (unsigned)P.hot_epoch;
  const unsigned p = h.epoch % 3u;
  const unsigned* list = hot + 8 + p * (unsigned)P.hot_cap;
  const unsigned* flag = hot + 8 + 3u * (unsigned)P.hot_cap + p * (unsigned)P.hot_tiles;
  unsigned cnt = hot[p];
  const unsigned stamp = hot[3 + p];
  if (blockIdx.y < (unsigned)P.hot_rows) {
    const unsigned idx = blockIdx.y * strips + blockIdx.x;
    h.run = false;
    if (stamp != hot_stamp(P, strips, nseg)) cnt = 0u;
    if (cnt > cap) cnt = cap;
    if (idx < cnt) {
      const unsigned tl = list[idx];
      if (tl < ntiles && flag[tl] == idx + 1u) { h.run = true; h.tile = tl; h.bx = (int)(tl % strips); h.by = (int)(tl / strips); }
    }
  } else {
    h.by = (int)blockIdx.y - P.hot_rows;
    h.tile = (unsigned)h.by * strips + (unsigned)h.bx;
    const unsigned f = flag[h.tile];                        // (independent of the two header loads above)
    if (stamp != hot_stamp(P, strips, nseg)) cnt = 0u;
    if (cnt > cap) cnt = cap;
    if (f != 0u && f - 1u < cnt && list[f - 1u] == h.tile) h.run = false;       // a hot CTA runs this tile
  }
  return h;
}

// called by every thread of every CTA of the launch at its end; `heavy`: this CTA's tile goes first next time
__device__ __forceinline__ void hot_end(const Sweep2Params& P, const HotTile& h, bool heavy) {
  if (P.hot_rows <= 0 || threadIdx.x != 0) return;
  unsigned* hot = P.hot;
  const unsigned strips = gridDim.x, nseg = gridDim.y - (unsigned)P.hot_rows;
  const unsigned cap = (unsigned)P.hot_rows * strips;
  const unsigned pn = (h.epoch + 1u) % 3u, pz = (h.epoch + 2u) % 3u;
  if (h.run) {
    unsigned f = 0u;
    if (heavy) {
      const unsigned idx = atomicAdd(hot + pn, 1u);
      if (idx < cap) { hot[8 + pn * (unsigned)P.hot_cap + idx] = h.tile; f = idx + 1u; }
    }
    hot[8 + 3u * (unsigned)P.hot_cap + pn * (unsigned)P.hot_tiles + h.tile] = f;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {                // one CTA per launch: stamp the list being written, clear the
    hot[3 + pn] = hot_stamp(P, strips, nseg);              // one the NEXT launch will write
    hot[pz] = 0u;
  }
}

// UNR > 1 unrolls the march loop so that the row-to-row hand-over of the carried state (q1 <- qN, qn0 <- qn1 <- qn2,
// c1 <- cN) becomes register renaming instead of moves: statically (scripts/sass_mix.py) 1476 -> 1430 instructions per
// march step for the headline Euler kernel at UNR = 2 with no spills (UNR = 3 spills); unmeasured on the GPU so far.
// (Measured and dropped in round 2, profiles/r02_f_variants.jsonl: 2 / 4 / 8 warps per CTA are no faster than 1 -- 4.77,
// 4.82, 5.29 ms against 4.78 -- and a __syncthreads per march step, tried to keep the warps of a CTA at the same place of
// the 23 KB loop body for the instruction cache, costs more than it gains: 4.80 / 4.97 / 5.70 ms.)
//
// QD > 0 (fast mode, specialised kernel, cp.async ring): UNIFORM-REGION FAST PATH with a QD-stage row ring.  Where the
// state is exactly uniform -- ahead of a front, behind a planar shock, inside an undisturbed bubble: most cells of the
// BASELINE shock-bubble data -- every jump is +-0, so every wave, fluctuation, correction and transverse term of the
// step is exactly zero and the update leaves q as it is; the reference spends its full operation count there all the
// same.  Per march step the warp votes on "the x-interfaces of row J+1 and the y-interface J+2 of all 32 columns join
// bit-identical states"; when that holds for the new interface and held for the two before it (the pending one that is
// finished in this step and the one whose results the cell still carries), the step reduces to: CFL speed |u_n| + c of
// the two interfaces from the per-cell data, store, rotate the carried rows.  The result is the same as the full path's
// (which would add exact zeros); the CFL speeds come from the cell data instead of the Roe average of two identical
// states (equal to rounding).  The kernel is then bound by HBM in such regions and by the FP64 pipe only near fronts;
// the deeper ring keeps enough rows in flight for the former.
template <class RP, int NW, int MINB = 1, int SPEC = 0, int UNR = 1, int QD = 0, int TMA = 0>
__global__ void __launch_bounds__(NW * 32, MINB) sweep2_unsplit_ws(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, NT = NW * 32;
  constexpr bool SP = (SPEC & 1) != 0;
  constexpr int LIM = SP ? (SPEC >> 4) : 0;
  constexpr bool BFL = (SPEC & 8) != 0;                    // branch-free limiter (fast mode)
  constexpr int PF = (SPEC >> 1) & 3;                      // q-row prefetch: 0 none, 1 L1 hint, 2 L2 hint, 3 cp.async ring
  constexpr int NC = RP::NCELL, NC1 = NC > 0 ? NC : 1;
  constexpr int RD = QD > 0 ? QD : 2;                      // stages of the row ring
  constexpr bool QUIET = UnsplitQuiet<RP, SPEC, QD>::value;
  static_assert((RD & (RD - 1)) == 0, "ring depth must be a power of two");
  using S = Smem2ws<RP, NW, RD>;
  constexpr int TX = 28;                                   // owned columns per warp
  // Fast mode, specialised kernel: the transverse corrections are carried as the unscaled sums G1 = bm(i+1) + bm'(i),
  // G2 = bp(i+1) + bp'(i) and scaled once by kxy = (dt/dx)(dt/dy)/2 where they enter q (gadd = -(dt/2dx) G, then
  // -dt/dy gadd, flux2.f90:185-203 / step2.f90:139-147): two FP64 instructions less per component and sweep.
#ifdef CLAW_STRICT
  constexpr bool FG = false;
#else
  constexpr bool FG = SP;
#endif
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x, lane = t & 31;
  double* const syw = smem + S::OFF_YW * NT + t;
  double* const sya = smem + S::OFF_YA * NT + t;
  double* const spf = smem + S::OFF_PF * NT + t;
  double* const sred = smem + S::PER_THREAD * NT;

  CLAW_RESOLVE_STEP2(P, qin, qout, dtdx, dtdy);
  [[maybe_unused]] HotTile hot;
  int bx = (int)blockIdx.x, by = (int)blockIdx.y;
  [[maybe_unused]] int nfull = 0;                          // march steps of this tile that took the full path
  if constexpr (QUIET && NW == 1) {
    hot = hot_begin(P);
    if (!hot.run) { hot_end(P, hot, false); return; }
    bx = hot.bx; by = hot.by;
  }
  const int nx = P.mx + 4;
  const int I = (bx * NW + (t >> 5)) * TX + lane;          // my column (array coords)
  const int Ic = I < nx ? I : nx - 1;                      // clamped for loads
  const int Jb = P.j0 + by * P.seg_rows;                   // first output row
  const int Je = min(Jb + P.seg_rows, P.j1);               // one past the last output row

  const double hdx = 0.5 * dtdx, hdy = 0.5 * dtdy;
  [[maybe_unused]] const double kxy = hdx * dtdy;
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
  const bool x_cfl_col = (lane >= 1) && (I >= 2) && (I <= P.mx + 2);
  const bool y_cfl_col = (I >= 1) && (I <= P.mx + 2);
  const bool store_col = (lane >= 2) && (lane <= 29) && (I <= P.mx + 1);

  // The first consumer of a freshly loaded row is the Roe average of the y-interface: without a prefetch the warp
  // sits on the DRAM latency once per march step (ncu: 13 % of all stall samples on that one instruction).  PF=3
  // keeps row J+3 in flight (cp.async into a thread-private two-stage ring) while row J+2 is being worked on.
  // TMA ring (one warp per CTA): lanes beyond the end of the row read the row's padding instead of a clamped copy of the
  // last column -- they never reach a stored cell or the CFL number, and the votes of the fast path leave them out; a
  // strip that would run over the padded row keeps the per-lane cp.async ring
  constexpr bool TM = (TMA != 0) && (PF == 3) && (NW == 1);
  [[maybe_unused]] const bool use_tma = TM && (bx * TX + 32 <= P.pitch);
  [[maybe_unused]] const bool live = !use_tma || (I < nx);
  [[maybe_unused]] const ring_t ring_s = ring_of(spf - lane), mbar_s = ring_of(sred + 32);
  [[maybe_unused]] unsigned rphase = 0u;
  if constexpr (TM) {
    if (use_tma && lane == 0) tma_mbar_init(mbar_s, RD);
    warp_sync();
  }
  // ring slot of array row r: r & (RD - 1); rows Jb-1 .. Jb+RD-3 are in flight when the loop starts
  if constexpr (PF == 3) {
#pragma unroll
    for (int k = 0; k < RD - 1; ++k) {
      const int r = Jb - 1 + k;
      if (r <= Je + 1) {
        if (TM && use_tma) {
          if (lane == 0) {
#pragma unroll
            for (int m = 0; m < M; ++m)
              tma_row_issue(ring_s, (m * RD + (r & (RD - 1))) * NT, qcol + m * P.plane + (long long)r * P.pitch, NT * 8,
                            mbar_s + 8u * (unsigned)(r & (RD - 1)), m == 0, M * NT * 8);
          }
        } else {
#pragma unroll
          for (int m = 0; m < M; ++m)
            cp_async8(spf + (m * RD + (r & (RD - 1))) * NT, qcol + m * P.plane + (long long)r * P.pitch);
        }
      }
      cp_async_commit();
    }
  }
  // running ring positions (slot offsets in doubles, component-major ring: see ring_t above): the slot written at the
  // first step (row Jb+RD-2) and the slot read at the first step (row Jb-1); running global pointer of the row to fetch
  [[maybe_unused]] int wslot = ((Jb + RD - 2) & (RD - 1)) * NT, rslot = ((Jb - 1) & (RD - 1)) * NT;
  [[maybe_unused]] const double* gfetch = qcol + (long long)(Jb + RD - 2) * P.pitch;
  // uniform-region fast path: were the y-interfaces J+1 and J zero on all lanes?  was the last step a quiet one?  has the
  // CFL speed of the current run of identical rows been taken for the x- / the y-interface?
  // A march segment starts three rows below its first output row.  The y-interfaces Jb-2 and Jb-3 that the vote of the
  // first two steps would have to know about only reach cells below Jb, which this segment does not store: they count as
  // quiet, so that a segment inside a uniform region does not pay two full march steps before it may go quiet (the
  // interfaces from Jb-1 on, which do reach stored cells, are classified from the data as everywhere else).
  [[maybe_unused]] const bool wq = (P.tune & 1) == 0;
  [[maybe_unused]] bool yq1 = wq, yq2 = wq, prev_quiet = false, qcfl_x = false, qcfl_y = false;

  // warm-up rows: see sweep2_unsplit (the specialised kernel runs the full body from the first step)
  // step Js of the ring: queue row Js+RD+1 into the slot row Js+1 was read from, wait for row Js+2, read it
  auto ring_fetch = [&](int Js, double* q) {
    if (TM && use_tma) {
      warp_sync();                                         // every lane has read the slot that is queued again
      if (Js + RD + 1 <= Je + 1 && lane == 0) {
        const double* g = gfetch;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          tma_row_issue(ring_s, wslot + m * RD * NT, g, NT * 8, mbar_s + 8u * (unsigned)(wslot / NT), m == 0, M * NT * 8);
          g += P.plane;
        }
      }
      gfetch += P.pitch;
      wslot = (wslot + NT) & (RD * NT - 1);
      const unsigned sl = (unsigned)(rslot / NT);
      tma_wait(mbar_s + 8u * sl, (rphase >> sl) & 1u);
      rphase ^= 1u << sl;
    } else {
      if (Js + RD + 1 <= Je + 1) {
        const double* g = gfetch;
#pragma unroll
        for (int m = 0; m < M; ++m) { cp_async8(spf + wslot + m * RD * NT, g); g += P.plane; }
      }
      cp_async_commit();
      gfetch += P.pitch;
      wslot = (wslot + NT) & (RD * NT - 1);
      cp_async_wait_keep<RD - 1>();
    }
#pragma unroll
    for (int m = 0; m < M; ++m) q[m] = spf[rslot + m * RD * NT];
    rslot = (rslot + NT) & (RD * NT - 1);
  };
  double qN[M];
  [[maybe_unused]] bool have_row = false;                  // qN already holds row J+2 (handed over by the tight quiet loop)
  [[maybe_unused]] int nquiet = 0;                         // consecutive quiet steps
#pragma unroll UNR
  for (int J = Jb - 3; J < Je; ++J) {
    if constexpr (PF == 3) {
      if (!QUIET || !have_row) ring_fetch(J, qN);
      have_row = false;
    } else {
#pragma unroll
      for (int m = 0; m < M; ++m) qN[m] = qcol[m * P.plane + (long long)(J + 2) * P.pitch];
      if constexpr (PF == 1 || PF == 2) {
        if (J + 1 < Je) {
#pragma unroll
          for (int m = 0; m < M; ++m) prefetch_row<PF>(qcol + m * P.plane + (long long)(J + 3) * P.pitch);
        }
      }
    }
    double cN[NC1];
    [[maybe_unused]] double qlx[M];                        // row J+1 of the column on my left (x-interface of this step)
    if constexpr (QUIET) {
      bool yz = true, xz = true;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        qlx[m] = shfl_left(q1[m]);
        yz = yz && same_bits(q1[m], qN[m]);
        xz = xz && same_bits(qlx[m], q1[m]);
      }
      const bool yq0 = warp_all(yz || !live);              // y-interface J+2 joins identical states in all 32 columns
      const bool xq0 = warp_all(xz || !live);
      const bool quiet = yq0 && xq0 && yq1 && yq2;
      yq2 = yq1; yq1 = yq0;
      if (quiet) {
        // Rows J+1 and J+2 carry the same bits in every column, so q1 and the cell data c1 stay as they are, and the CFL
        // speed of the two zero-jump interfaces is the same for every row of this run of quiet steps: it is taken once
        // per run and interface direction, at the first step whose row is inside the range the reference visits
        // (flux2.f90:111-119 over the pencils of step2.f90:92,172).
        const bool ry = (J + 2 >= 2) && (J + 2 <= P.my + 2), rx = (J + 1 >= 1) && (J + 1 <= P.my + 2);
        if (ry && !qcfl_y) { if (y_cfl_col) cfl = fmax2(cfl, dtdy * RP::quiet_smax(P.rp, 1, q1, c1)); qcfl_y = true; }
        if (rx && !qcfl_x) { if (x_cfl_col) cfl = fmax2(cfl, dtdx * RP::quiet_smax(P.rp, 0, q1, c1)); qcfl_x = true; }
        if (store_col && J >= Jb) {
          double* o = qout + (long long)J * P.pitch + I;
#pragma unroll
          for (int m = 0; m < M; ++m) { *o = qn0[m]; o += P.plane; }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) { qn0[m] = qn1[m]; qn1[m] = q1[m]; }
        prev_quiet = true;
        ++nquiet;
        // From the third quiet step of a run on, the carried rows qn0, qn1 hold the bits of q1 too, and a further row
        // with those bits also passes the x-test (it is the same 32 values again): while rows keep coming in unchanged
        // the step is "fetch, compare with q1, vote, store q1" -- a tight loop of its own, without the rotation of the
        // carried state and without the bookkeeping of the general step.
        if (nquiet >= 2 && qcfl_x && qcfl_y && J >= Jb) {
#pragma unroll
          for (int m = 0; m < M; ++m) { pin(qn0[m]); pin(qn1[m]); pin(q1[m]); }
          // (qn1 and qn0 carry the bits of q1 here; comparing with the one and storing the other keeps both alive in
          // the loop, so that the compiler does not sink their last assignments into every exit of it)
          if (TM && use_tma)
            have_row = quiet_run<M, RD, NT, TM>(J, Je, qn1, qn0, qN, spf, wslot, rslot, gfetch, P.plane, P.pitch,
                                                qout + (long long)(J + 1) * P.pitch + I, store_col, mbar_s, &rphase, live);
          else
            have_row = quiet_run<M, RD, NT>(J, Je, qn1, qn0, qN, spf, wslot, rslot, gfetch, P.plane, P.pitch,
                                            qout + (long long)(J + 1) * P.pitch + I, store_col);
        }
        continue;
      }
      qcfl_x = false; qcfl_y = false;
      nquiet = 0;
      ++nfull;
      if (prev_quiet) {
        // the pending y-interface (J+1) was skipped: it is a zero interface -- zero waves, fluctuations and limiter dot
        // products, finite stand-in Roe data (they only ever multiply those zeros); the carried results of interface J
        // (yI_*) are zero already, which is what made the skipped steps quiet
        double roeQ[NR];
        RP::quiet_roe(roeQ);
#pragma unroll
        for (int k = 0; k < W * M; ++k) syw[k * NT] = 0.0;
#pragma unroll
        for (int k = 0; k < 2 * M; ++k) sya[k * NT] = 0.0;
#pragma unroll
        for (int k = 0; k < NR; ++k) sya[(2 * M + k) * NT] = roeQ[k];
#pragma unroll
        for (int mw = 0; mw < W; ++mw) sya[(2 * M + NR + mw) * NT] = 0.0;
        prev_quiet = false;
      }
    }
    if constexpr (NC > 0) RP::cell(P.rp, qN, cN);

    // =========================== y-pipeline ==========================================
    double ycomb[M], gy1[M], gy2[M];                       // natural frame
#pragma unroll
    for (int m = 0; m < M; ++m) { ycomb[m] = 0.0; gy1[m] = 0.0; gy2[m] = 0.0; }
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
            double cqxx[M];
#ifndef CLAW_STRICT
            if constexpr (SP && !BFL) {
              limited_second_order<RP, LIM>(P, dtdy, waveC, sC, wnorm2, dotl, dotn, cqxx);
            } else
#endif
            {
              if (limit) limit_waves<RP, LIM, BFL>(P, waveC, sC, wnorm2, dotl, dotn);
              second_order<RP>(dtdy, waveC, sC, cqxx);
            }
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
            if constexpr (FG) {
              const double G1 = bmA[m] + yI_bm[m], G2 = bpA[m] + yI_bp[m];
              ycomb[rot<M>(m)] = fma(kxy, G2 - G1, qadd[m] - dtdy * (fadd[m] - yI_fadd[m]));
              gy1[rot<M>(m)] = G1;
              gy2[rot<M>(m)] = G2;
            } else {
              const double g1 = (0.0 - hdy * bmA[m]) - hdy * yI_bm[m];      // flux2.f90:185-203
              const double g2 = (0.0 - hdy * bpA[m]) - hdy * yI_bp[m];
              // step2.f90:223-225: qadd - dtdy*(fadd(j+1)-fadd(j)) - dtdx*(gadd(2)-gadd(1))
              ycomb[rot<M>(m)] = (qadd[m] - dtdy * (fadd[m] - yI_fadd[m])) - dtdx * (g2 - g1);
              gy1[rot<M>(m)] = g1;
              gy2[rot<M>(m)] = g2;
            }
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
      {
        double ql[M], cl[NC1];
#pragma unroll
        for (int m = 0; m < M; ++m) {
          if constexpr (QUIET) ql[m] = qlx[m];
          else ql[m] = shfl_left(q1[m]);
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) cl[k] = shfl_left(c1[k]);
        RP::normal(P.rp, 0, ql, q1, cl, c1, wave, s, amdq, apdq, roe);    // flux2.f90:101
      }
#pragma unroll
      for (int m = 0; m < M; ++m) { cqxx[m] = 0.0; fadd[m] = 0.0; }
      if (second) {
        bool done = false;
        if (limit) {                                                       // flux2.f90:127
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
          if constexpr (SP && !BFL) {
            limited_second_order<RP, LIM>(P, dtdx, wave, s, wnorm2, dotl, dotr, cqxx);
            done = true;
          } else
#endif
          limit_waves<RP, LIM, BFL>(P, wave, s, wnorm2, dotl, dotr);
        }
        if (!done) second_order<RP>(dtdx, wave, s, cqxx);                  // flux2.f90:134-146
#pragma unroll
        for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
      }
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
      // my left-face results belong to the cell on my left: fetch the right neighbour's
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double amdqR = shfl_right(amdq[m]), faddR = shfl_right(fadd[m]);
        const double bmAR = shfl_right(bmA[m]), bpAR = shfl_right(bpA[m]);
        const double qadd = (0.0 - dtdx * apdq[m]) - dtdx * amdqR;         // flux2.f90:105-108
        if constexpr (FG) {
          const double G1 = bmAR + bmP[m], G2 = bpAR + bpP[m];
          qn0[m] = fma(kxy, G1, qn0[m]);
          qn1[m] = fma(kxy, G2 - G1, (qn1[m] + qadd) - dtdx * (faddR - fadd[m]));
          qn2[m] = fma(-kxy, G2, qN[m]);
        } else {
          const double g1 = (0.0 - hdx * bmAR) - hdx * bmP[m];              // flux2.f90:185-203
          const double g2 = (0.0 - hdx * bpAR) - hdx * bpP[m];
          // step2.f90:139-147, pencil J+1: rows J, J+1, J+2
          qn0[m] = qn0[m] - dtdy * g1;
          qn1[m] = ((qn1[m] + qadd) - dtdx * (faddR - fadd[m])) - dtdy * (g2 - g1);
          qn2[m] = qN[m] + dtdy * g2;
        }
      }
    }

    // =========================== cell (I, J) is complete =============================
    // (the shuffles are executed by the whole warp on every step; only the stores are guarded)
#pragma unroll
    for (int m = 0; m < M; ++m) {
      // step2.f90:221-229, columns I-1, I, I+1
      double r;
      if constexpr (FG) {
        r = fma(-kxy, shfl_left(gy2[m]), qn0[m]);
        r = r + ycomb[m];
        r = fma(kxy, shfl_right(gy1[m]), r);
      } else {
        r = qn0[m] + dtdx * shfl_left(gy2[m]);
        r = r + ycomb[m];
        r = r - dtdx * shfl_right(gy1[m]);
      }
      if (store_col && J >= Jb) qout[m * P.plane + (long long)J * P.pitch + I] = r;
    }
#pragma unroll
    for (int m = 0; m < M; ++m) { qn0[m] = qn1[m]; qn1[m] = qn2[m]; q1[m] = qN[m]; }
#pragma unroll
    for (int k = 0; k < NC; ++k) c1[k] = cN[k];
  }
  block_max_cfl<NT>(cfl, P.cfl_bits, sred);
  if constexpr (QUIET && NW == 1) hot_end(P, hot, 4 * nfull >= P.seg_rows);
}

// =====================================================================================
// Dimension-split step, warp-synchronous (same algorithm, order of operations and bits as sweep2_split in
// sweep2.cuh): x-sweep of row J+2 with the neighbouring interfaces exchanged by warp shuffle, then the y-pipeline on
// the x-updated rows in thread-private registers / shared memory; no __syncthreads in the march loop.  Rows are
// prefetched one march step ahead with cp.async (PF != 0).  do_y = 0: x-sweep only, over rows [j0, j1).
// =====================================================================================
// LIM > 0: order 2 with limiter id LIM on every wave family, known at compile time (the generic kernel selects the
// limiter per wave at run time: a five-way switch per wave and interface in both sweeps).  UNR: see sweep2_unsplit_ws.
// QD > 0 (fast mode, LIM > 0, fused x+y step): uniform-region fast path with a QD-stage row ring, as in
// sweep2_unsplit_ws: when the x-interfaces of row J+2 and the y-interfaces J+2, J+1, J of all 32 columns join
// bit-identical states, the x-sweep leaves the row as it is and the y-update adds exact zeros; the step then only takes
// the CFL speed of the uniform state (once per run of such steps), stores row J and rotates the carried rows.
template <class RP, int NW, int MINB = 1, int PF = 1, int UNR = 1, int LIM = 0, int QD = 0>
__global__ void __launch_bounds__(NW * 32, MINB) sweep2_split_ws(const __grid_constant__ Sweep2Params P) {
  constexpr int M = RP::MEQN, W = RP::MWAVES, NR = RP::NROE, NT = NW * 32;
  constexpr int NC = RP::NCELL, NC1 = NC > 0 ? NC : 1;
  constexpr int RD = QD > 0 ? QD : 2;
  constexpr bool QUIET = SplitQuiet<RP, PF, LIM, QD>::value;
  static_assert((RD & (RD - 1)) == 0, "ring depth must be a power of two");
  using S = Smem2ws<RP, NW, RD>;
  constexpr int TX = 28;
  CLAW_DYN_SMEM(smem);
  const int t = threadIdx.x, lane = t & 31;
  double* const syw = smem + S::OFF_YW * NT + t;
  double* const sya = smem + S::OFF_YA * NT + t;
  double* const spf = smem + S::OFF_PF * NT + t;
  double* const sred = smem + S::PER_THREAD * NT;

  CLAW_RESOLVE_STEP2(P, qin, qout, dtdx, dtdy);
  [[maybe_unused]] HotTile hot;
  int bx = (int)blockIdx.x, by = (int)blockIdx.y;
  [[maybe_unused]] int nfull = 0;
  if constexpr (QUIET && NW == 1) {
    hot = hot_begin(P);
    if (!hot.run) { hot_end(P, hot, false); return; }
    bx = hot.bx; by = hot.by;
  }
  const int nx = P.mx + 4;
  const int I = (bx * NW + (t >> 5)) * TX + lane;
  const int Ic = I < nx ? I : nx - 1;
  const bool second = LIM ? true : (P.order != 1);
  const bool limit = LIM ? true : (second && (P.limit != 0));
  const bool do_x = P.do_x != 0, do_y = P.do_y != 0;
  const double* qcol = qin + Ic;
  double cfl = 0.0;
  const bool x_cfl_col = (lane >= 1) && (I >= 2) && (I <= P.mx + 2);
  const bool interior_col = (I >= 2) && (I <= P.mx + 1);
  const bool x_ok = (lane >= 2) && (lane <= 29);            // my x-updated value is complete
  const bool col_ok = (I <= P.mx + 3) && (!do_x || !interior_col || x_ok);
  const int Jb = P.j0 + by * P.seg_rows;
  const int Je = min(Jb + P.seg_rows, P.j1);

  // x-sweep of one row (step2ds.f90:96-160 with ids = 1): q -> x-updated value of my cell
  auto x_sweep = [&](const double* q, double* out, const double* ql_in) {
    double cq[NC1], cl[NC1], ql[M];
    if constexpr (NC > 0) RP::cell(P.rp, q, cq);
#pragma unroll
    for (int m = 0; m < M; ++m) ql[m] = (ql_in != nullptr) ? ql_in[m] : shfl_left(q[m]);
#pragma unroll
    for (int k = 0; k < NC; ++k) cl[k] = shfl_left(cq[k]);
    double wave[W][M], s[W], amdq[M], apdq[M], roe[NR], cqxx[M], fadd[M];
    RP::normal(P.rp, 0, ql, q, cl, cq, wave, s, amdq, apdq, roe);          // flux2.f90:101
    if (x_cfl_col) cfl = cfl_update<RP>(cfl, dtdx, s);                     // flux2.f90:111-119
#pragma unroll
    for (int m = 0; m < M; ++m) { cqxx[m] = 0.0; fadd[m] = 0.0; }
    if (second) {
      bool done = false;
      if (limit) {                                                         // flux2.f90:127
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
        done = true;
#else
        limit_waves<RP, LIM>(P, wave, s, wnorm2, dotl, dotr);
#endif
      }
      if (!done) second_order<RP>(dtdx, wave, s, cqxx);                    // flux2.f90:134-146
#pragma unroll
      for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const double amdqR = shfl_right(amdq[m]), faddR = shfl_right(fadd[m]);
      const double qadd = (0.0 - dtdx * apdq[m]) - dtdx * amdqR;           // flux2.f90:105-108
      out[m] = (q[m] + qadd) - dtdx * (faddR - fadd[m]);                   // step2ds.f90:149-152
    }
  };

  // rows through the two-stage cp.async ring: row(J) is consumed while row(J+1) is in flight
  auto prefetch = [&](int slot, int Jrow) {
#pragma unroll
    for (int m = 0; m < M; ++m) cp_async8(spf + ((slot & 1) * M + m) * NT, qcol + m * P.plane + (long long)Jrow * P.pitch);
  };
  auto fetch = [&](int slot, int Jrow, int Jnext, bool more, double* q) {
    if constexpr (PF != 0) {
      cp_async_wait_all();
#pragma unroll
      for (int m = 0; m < M; ++m) q[m] = spf[((slot & 1) * M + m) * NT];
      if (more) prefetch(slot + 1, Jnext);
      cp_async_commit();
    } else {
#pragma unroll
      for (int m = 0; m < M; ++m) q[m] = qcol[m * P.plane + (long long)Jrow * P.pitch];
    }
  };

  if (!do_y) {
    if constexpr (PF != 0) { prefetch(Jb, Jb); cp_async_commit(); }
    for (int J = Jb; J < Je; ++J) {
      double q[M], r[M];
      fetch(J, J, J + 1, J + 1 < Je, q);
      x_sweep(q, r, nullptr);
      if (interior_col && x_ok) {
#pragma unroll
        for (int m = 0; m < M; ++m) qout[m * P.plane + (long long)J * P.pitch + I] = r[m];
      }
    }
    block_max_cfl<NT>(cfl, P.cfl_bits, sred);
    return;
  }

  double qx0[M], qx1[M], cx1[NC1];                         // x-updated rows J, J+1 (natural frame)
  double yI_apdq[M], yI_fadd[M];
#pragma unroll
  for (int m = 0; m < M; ++m) { qx0[m] = 0.0; qx1[m] = 0.0; yI_apdq[m] = 0.0; yI_fadd[m] = 0.0; }
#pragma unroll
  for (int k = 0; k < NC1; ++k) cx1[k] = 0.0;
  // QD-stage ring (QUIET builds): rows Jb-2 .. Jb+RD-4 are in flight when the loop starts; running slot offsets
  [[maybe_unused]] int wslot = ((Jb + RD - 3) & (RD - 1)) * NT, rslot = ((Jb - 2) & (RD - 1)) * NT;
  [[maybe_unused]] const double* gfetch = qcol + (long long)(Jb + RD - 3) * P.pitch;
  // warm-up steps of a segment (J < Jb): interfaces that only reach cells below Jb count as quiet (see sweep2_unsplit_ws)
  [[maybe_unused]] const bool wq = (P.tune & 1) == 0;
  [[maybe_unused]] bool yq1 = wq, yq2 = wq, prev_quiet = false, qcfl_x = false, qcfl_y = false;
  if constexpr (QD > 0) {
#pragma unroll
    for (int k = 0; k < RD - 1; ++k) {
      const int r = Jb - 2 + k;
      if (r <= Je + 1) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          cp_async8(spf + (m * RD + (r & (RD - 1))) * NT, qcol + m * P.plane + (long long)r * P.pitch);
      }
      cp_async_commit();
    }
  } else if constexpr (PF != 0) { prefetch(Jb, Jb - 2); cp_async_commit(); }

  // step Js of the QD-stage ring: queue row Js+RD+1 into the slot row Js+1 was read from, wait for row Js+2, read it
  auto ring_fetch = [&](int Js, double* q) {
    if (Js + RD + 1 <= Je + 1) {
      const double* g = gfetch;
#pragma unroll
      for (int m = 0; m < M; ++m) { cp_async8(spf + wslot + m * RD * NT, g); g += P.plane; }
    }
    cp_async_commit();
    gfetch += P.pitch;
    wslot = (wslot + NT) & (RD * NT - 1);
    cp_async_wait_keep<RD - 1>();
#pragma unroll
    for (int m = 0; m < M; ++m) q[m] = spf[rslot + m * RD * NT];
    rslot = (rslot + NT) & (RD * NT - 1);
  };
  double qx2[M];
  [[maybe_unused]] bool have_row = false;                  // qx2 already holds row J+2 (handed over by the tight quiet loop)
  [[maybe_unused]] int nquiet = 0;
#pragma unroll UNR
  for (int J = Jb - 4; J < Je; ++J) {
    // ---- x-sweep of row J+2 ----------------------------------------------------------
    if constexpr (QD > 0) {
      if (!QUIET || !have_row) ring_fetch(J, qx2);
      have_row = false;
    } else {
      fetch(J - (Jb - 4) + Jb, J + 2, J + 3, J + 1 < Je, qx2);
    }
    [[maybe_unused]] double qlx[M];
    if constexpr (QUIET) {
      bool xz = true, yz = true;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        qlx[m] = shfl_left(qx2[m]);
        xz = xz && same_bits(qlx[m], qx2[m]);
        yz = yz && same_bits(qx1[m], qx2[m]);
      }
      const bool xq0 = warp_all(xz);                       // the x-sweep leaves row J+2 as it is in all 32 columns
      if (wq && J == Jb - 4) {
        // first warm-up step: row Jb-2 is only x-swept (no y-interface yet); a uniform row stays as it is
        if (do_x && xq0) {
          if (!qcfl_x) { if (x_cfl_col) cfl = fmax2(cfl, dtdx * RP::quiet_smax(P.rp, 0, qx2)); qcfl_x = true; }
#pragma unroll
          for (int m = 0; m < M; ++m) { qx0[m] = qx1[m]; qx1[m] = qx2[m]; }
          continue;
        }
      }
      const bool yq0 = warp_all(yz) && xq0 && (J >= Jb - 3);   // the x-updated rows J+1 and J+2 carry the same bits
      const bool quiet = do_x && yq0 && yq1 && yq2 && (J >= (wq ? Jb - 3 : Jb - 1));
      if (!wq || J >= Jb - 3) { yq2 = yq1; yq1 = yq0; }
      if (quiet) {
        const bool ry = (J + 2 >= 2) && (J + 2 <= P.my + 2);
        if (!qcfl_x) { if (x_cfl_col) cfl = fmax2(cfl, dtdx * RP::quiet_smax(P.rp, 0, qx2)); qcfl_x = true; }
        if (ry && !qcfl_y) { if (col_ok) cfl = fmax2(cfl, dtdy * RP::quiet_smax(P.rp, 1, qx2)); qcfl_y = true; }
        if (J >= Jb && col_ok && interior_col) {
          double* o = qout + (long long)J * P.pitch + I;
#pragma unroll
          for (int m = 0; m < M; ++m) { *o = qx0[m]; o += P.plane; }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) qx0[m] = qx1[m];       // qx1 = qx2: the same bits already
        prev_quiet = true;
        ++nquiet;
        // from the second quiet step of a run on, qx0 carries the bits of qx1 too: while rows keep coming in with those
        // bits (which also means they pass the x-test again) the step is "fetch, compare, vote, store"
        if (nquiet >= 2 && qcfl_x && qcfl_y && J >= Jb) {
#pragma unroll
          for (int m = 0; m < M; ++m) { pin(qx0[m]); pin(qx1[m]); }
          have_row = quiet_run<M, RD, NT>(J, Je, qx1, qx0, qx2, spf, wslot, rslot, gfetch, P.plane, P.pitch,
                                          qout + (long long)(J + 1) * P.pitch + I, col_ok && interior_col);
        }
        continue;
      }
      qcfl_x = false; qcfl_y = false;
      nquiet = 0;
      ++nfull;
      if (prev_quiet) {                                    // the pending y-interface was skipped: a zero interface
        double roeQ[NR];
        RP::quiet_roe(roeQ);
#pragma unroll
        for (int k = 0; k < W * M; ++k) syw[k * NT] = 0.0;
#pragma unroll
        for (int k = 0; k < 2 * M; ++k) sya[k * NT] = 0.0;
#pragma unroll
        for (int k = 0; k < NR; ++k) sya[(2 * M + k) * NT] = roeQ[k];
#pragma unroll
        for (int mw = 0; mw < W; ++mw) sya[(2 * M + NR + mw) * NT] = 0.0;
        prev_quiet = false;
      }
    }
    if (do_x) {
      double r[M];
      if constexpr (QUIET) x_sweep(qx2, r, qlx);
      else x_sweep(qx2, r, nullptr);
      if (interior_col) {                                  // ghost columns are not updated by the x-sweep
#pragma unroll
        for (int m = 0; m < M; ++m) qx2[m] = r[m];
      }
    }
    // ---- y-pipeline --------------------------------------------------------------------
    double cx2[NC1];
    if constexpr (NC > 0) RP::cell(P.rp, qx2, cx2);        // of the x-updated cell
    if (J >= Jb - 3) {
      double waveN[W][M], sN[W], amdqN[M], apdqN[M], roeN[NR];
      {
        double ql[M], qr[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { ql[m] = qx1[rot<M>(m)]; qr[m] = qx2[rot<M>(m)]; }
        RP::normal(P.rp, 1, ql, qr, cx1, cx2, waveN, sN, amdqN, apdqN, roeN);
      }
      if (col_ok && (J + 2 >= 2) && (J + 2 <= P.my + 2)) cfl = cfl_update<RP>(cfl, dtdy, sN);
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
            double cqxx[M];
#ifndef CLAW_STRICT
            if (limit) limited_second_order<RP, LIM>(P, dtdy, waveC, sC, wnorm2, dotl, dotn, cqxx);
            else second_order<RP>(dtdy, waveC, sC, cqxx);
#else
            if (limit) limit_waves<RP, LIM>(P, waveC, sC, wnorm2, dotl, dotn);
            second_order<RP>(dtdy, waveC, sC, cqxx);
#endif
#pragma unroll
            for (int m = 0; m < M; ++m) fadd[m] = 0.0 + 0.5 * cqxx[m];
          }
#pragma unroll
          for (int m = 0; m < M; ++m) {
            const double qadd = (0.0 - dtdy * yI_apdq[m]) - dtdy * amdqC[m];
            const double r = (qx0[rot<M>(m)] + qadd) - dtdy * (fadd[m] - yI_fadd[m]);     // step2ds.f90:237-240
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
  if constexpr (QUIET && NW == 1) hot_end(P, hot, 4 * nfull >= P.seg_rows);
}

}  // namespace CLAW_NS
}  // namespace clawcuda
