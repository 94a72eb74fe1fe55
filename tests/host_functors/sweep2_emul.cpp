// TEST INFRASTRUCTURE.  The warp-synchronous 2-D sweep KERNEL sources (clawpack_b200/csrc/sweep2_ws.cuh: unsplit and
// dimension-split march kernels with shuffle exchange and the row-prefetch ring; csrc/sweep2_aux.cuh: their aux-reading
// counterparts) executed on the CPU by the fiber emulation of cuda_emul.h, behind entry points with the semantics of
// claw_step2_host / claw_step2ds_host (csrc/clawcuda.cu: AoS -> SoA planes, one launch over the whole grid in march
// segments, region read-back).  `variant` selects the instantiation like CLAW_VARIANT does in csrc/sweep_inst.cu.
// Strict arithmetic (-ffp-contract=off), or namespace fast with -DCLAW_HOST_FAST.  Nothing in the product links this file.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST          // namespace fast: the kernels' fast-mode code paths (folded limiter, FMA contraction)
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include "cuda_emul.h"
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign; using std::fmax; using std::sin;
using std::cos; using std::fmod; using std::llround;
#include "../../clawpack_b200/csrc/sweep2.cuh"
#include "../../clawpack_b200/csrc/sweep2_ws.cuh"
#include "../../clawpack_b200/csrc/rp2_aux.cuh"
#include "../../clawpack_b200/csrc/sweep2_aux.cuh"

using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

namespace {
// the grid / segment logic of launch_kernel() in csrc/sweep_inst.cu (seg_hint > 0 as set by claw_set_tuning)
template <class F> void launch(F kern, int nthreads, bool x_only, Sweep2Params& P, int seg) {
  const int TX = 28 * (nthreads / 32);
  const int strips = (P.mx + TX - 1) / TX;
  P.j0 = x_only ? 0 : 2;
  P.j1 = x_only ? P.my + 4 : P.my + 2;
  const int rows = P.j1 - P.j0;
  if (seg <= 0 || seg > rows) seg = rows;
  P.seg_rows = seg;
  emul::Dim3 grid, block;
  grid.x = strips; grid.y = (rows + seg - 1) / seg; block.x = nthreads;
  P.hot_rows = 0;
  if (P.hot != nullptr && !x_only) {                    // tile ordering as in launch_kernel(): hot grid rows in front
    int hy = (int)grid.y / 16;
    if (hy < 1) hy = 1;
    while (hy > 0 && (long long)hy * grid.x > P.hot_cap) --hy;
    if (hy > 0 && (long long)grid.x * grid.y <= P.hot_tiles) { P.hot_rows = hy; grid.y += hy; }
  }
  const Sweep2Params Pc = P;
  emul::launch(grid, block, [&]() { kern(Pc); });
}

template <int W> bool all_limiters(const Sweep2Params& P, int lim) {
  if (P.order == 1 || !P.limit) return false;
  for (int mw = 0; mw < W; ++mw) if (P.mthlim[mw] != lim) return false;
  return true;
}

template <class RP> void split_ws(Sweep2Params& P, int seg) { launch(sweep2_split_ws<RP, 1, 1, 1, 1, 0>, 32, !P.do_y, P, seg); }
template <class RP> void unsplit_ws(Sweep2Params& P, int seg) { launch(sweep2_unsplit_ws<RP, 1, 1, 6, 1>, 32, false, P, seg); }
template <class RP> void aux_k(int kind, Sweep2Params& P, int seg) {
  if (kind == 0) launch(sweep2_unsplit_aux<RP, 2>, 64, false, P, seg);
  else launch(sweep2_split_aux<RP, 2>, 64, !P.do_y, P, seg);
}

int run(int rp_id, int kind, int variant, Sweep2Params& P, int seg) {
  switch (rp_id) {
    case 2:
      if (kind == 1 && variant == 53 && all_limiters<2>(P, 4)) {
        launch(sweep2_split_ws<RpAcoustics2D, 1, 1, 1, 1, 4>, 32, !P.do_y, P, seg);
        return 0;
      }
      if (kind == 1 && variant == 73 && all_limiters<2>(P, 4) && P.do_x && P.do_y) {   // + uniform-region fast path, 8-stage ring
        launch(sweep2_split_ws<RpAcoustics2D, 1, 1, 1, 1, 4, 8>, 32, false, P, seg);
        return 0;
      }
      kind ? split_ws<RpAcoustics2D>(P, seg) : unsplit_ws<RpAcoustics2D>(P, seg); return 0;
    case 3:
      if (kind == 0 && P.mtrans == 2 && all_limiters<4>(P, 4) && variant != 9) {      // the headline kernel of bench.py
        // variants 72 / 71: with the uniform-region fast path and an 8- / 4-stage row ring (fast mode; in strict mode the
        // template argument only changes the ring depth)
        if (variant == 91) launch(sweep2_unsplit_ws<RpEuler4Wave2DT<3>, 1, 1, (4 << 4) | 1 | 6, 1, 8, 1>, 32, false, P, seg);   // TMA ring
        else if (variant == 72) launch(sweep2_unsplit_ws<RpEuler4Wave2DT<3>, 1, 1, (4 << 4) | 1 | 6, 1, 8>, 32, false, P, seg);
        else if (variant == 71) launch(sweep2_unsplit_ws<RpEuler4Wave2DT<3>, 1, 1, (4 << 4) | 1 | 6, 1, 4>, 32, false, P, seg);
        else launch(sweep2_unsplit_ws<RpEuler4Wave2DT<3>, 1, 1, (4 << 4) | 1 | 6, 1>, 32, false, P, seg);
        return 0;
      }
      if (kind == 1 && variant == 53 && all_limiters<4>(P, 4)) {
        launch(sweep2_split_ws<RpEuler4Wave2D, 1, 1, 1, 1, 4>, 32, !P.do_y, P, seg);
        return 0;
      }
      kind ? split_ws<RpEuler4Wave2D>(P, seg) : unsplit_ws<RpEuler4Wave2D>(P, seg); return 0;
    case 4:
      if (kind == 1 && (variant == 53 || variant == 54) && all_limiters<3>(P, 4)) {     // limiter as a template argument
        launch(sweep2_split_ws<RpShallowRoeEfix2D, 1, 1, 1, 1, 4>, 32, !P.do_y, P, seg);
        return 0;
      }
      if (kind == 1 && variant == 73 && all_limiters<3>(P, 4) && P.do_x && P.do_y) {   // + uniform-region fast path, 8-stage ring
        launch(sweep2_split_ws<RpShallowRoeEfix2D, 1, 1, 1, 1, 4, 8>, 32, false, P, seg);
        return 0;
      }
      if (kind == 1 && variant == 74 && all_limiters<3>(P, 4) && P.do_x && P.do_y) {   // the same with a 4-stage ring
        launch(sweep2_split_ws<RpShallowRoeEfix2D, 1, 1, 1, 1, 4, 4>, 32, false, P, seg);
        return 0;
      }
      kind ? split_ws<RpShallowRoeEfix2D>(P, seg) : unsplit_ws<RpShallowRoeEfix2D>(P, seg); return 0;
    case 9: kind ? split_ws<RpEuler5Wave2D>(P, seg) : unsplit_ws<RpEuler5Wave2D>(P, seg); return 0;
    case 11: kind ? split_ws<RpEulerHlle2D>(P, seg) : unsplit_ws<RpEulerHlle2D>(P, seg); return 0;
    case 12: kind ? split_ws<RpShallowHlle2D>(P, seg) : unsplit_ws<RpShallowHlle2D>(P, seg); return 0;
    case 16: kind ? split_ws<RpBurgers2D>(P, seg) : unsplit_ws<RpBurgers2D>(P, seg); return 0;
    case 18: kind ? split_ws<RpKpp2D>(P, seg) : unsplit_ws<RpKpp2D>(P, seg); return 0;
    case 7: aux_k<RpVcAcoustics2D>(kind, P, seg); return 0;
    case 8: aux_k<RpVcAdvection2D>(kind, P, seg); return 0;
    case 10: aux_k<RpPsystem2D>(kind, P, seg); return 0;
    case 14: aux_k<RpShallowBathymetryFwave2D>(kind, P, seg); return 0;
    case 19: aux_k<RpVcElasticity2D>(kind, P, seg); return 0;
    case 20: aux_k<RpVcAdvectionVector2D>(kind, P, seg); return 0;
  }
  return 5;
}

std::vector<unsigned> g_hot(8 + 3 * 4096 + 3 * 65536, 0u);
unsigned g_hot_epoch = 0;      // HotTiles buffer of the 1xx variants

struct Soa {
  int nx, ny, pitch; long long plane;
  std::vector<double> in, out, aux;
};

// ids: 0 = unsplit step2, 1 / 2 = step2ds sweep, 3 = fused dimension-split step (x-sweep of all rows, then y-sweep)
int step(int ids, int meqn, int mwaves, int maux, int mx, int my, const double* qold, double* qnew, const double* auxbc,
         double dx, double dy, double dt, const int* method7, const int* mthlim, int use_fwave, int rp_id,
         const double* cparam, int ncparam, int variant, int seg, double* cfl) {
  Soa s;
  s.nx = mx + 4; s.ny = my + 4; s.pitch = s.nx + 3;                 // an odd pitch: rows are not aligned
  if (variant >= 100) s.pitch = (s.nx + 15) / 16 * 16;              // the product's pitch
  s.plane = (long long)s.pitch * s.ny;
  s.in.assign((size_t)s.plane * meqn, 0.0);
  s.out.assign((size_t)s.plane * meqn, -7.0e300);                  // stale buffer: only the region read back may be used
  for (int j = 0; j < s.ny; ++j) for (int i = 0; i < s.nx; ++i) for (int m = 0; m < meqn; ++m)
    s.in[m * s.plane + (long long)j * s.pitch + i] = qold[m + (size_t)meqn * (i + (size_t)s.nx * j)];
  if (maux > 0) {
    s.aux.assign((size_t)s.plane * maux, 0.0);
    for (int j = 0; j < s.ny; ++j) for (int i = 0; i < s.nx; ++i) for (int m = 0; m < maux; ++m)
      s.aux[m * s.plane + (long long)j * s.pitch + i] = auxbc[m + (size_t)maux * (i + (size_t)s.nx * j)];
  }
  unsigned long long cfl_bits = 0;
  Sweep2Params P{};                                                 // as launch_step2() in csrc/clawcuda.cu
  P.qin = s.in.data(); P.qout = s.out.data(); P.cfl_bits = &cfl_bits;
  P.mx = mx; P.my = my; P.pitch = s.pitch; P.plane = s.plane;
  P.dtdx = dt / dx; P.dtdy = dt / dy;
  P.order = method7[1]; P.mtrans = method7[2]; P.limit = 0;
  for (int mw = 0; mw < mwaves; ++mw) { P.mthlim[mw] = mthlim[mw]; if (mthlim[mw] > 0) P.limit = 1; }
  P.do_x = (ids != 2); P.do_y = (ids != 1);
  P.aux = maux > 0 ? s.aux.data() : nullptr; P.maux = maux; P.capa = method7[5]; P.fwave = use_fwave;
  for (int k = 0; k < 4; ++k) P.rp.p[k] = (cparam && k < ncparam) ? cparam[k] : 0.0;
  if (variant >= 100) {
    // variants 1xx: variant xx with the tile ordering of the uniform-region kernels (HotTiles, sweep2_ws.cuh).  The
    // buffer lives across calls like the handle's does across steps -- and across test cases with other grids, which
    // exercises the stale-list checks.
    P.hot = g_hot.data(); P.hot_cap = 4096; P.hot_tiles = 65536; P.hot_epoch = g_hot_epoch++;
    variant -= 100;
  }
  if (int rc = run(rp_id, ids == 0 ? 0 : 1, variant, P, seg)) return rc;
  int i0 = 2, i1 = 2 + mx, j0 = 2, j1 = 2 + my;                    // get_interior / get_region of the plug-in calls
  if (ids == 1) { j0 = 0; j1 = s.ny; }
  if (ids == 2) { i0 = 0; i1 = s.nx; }
  for (int j = j0; j < j1; ++j) for (int i = i0; i < i1; ++i) for (int m = 0; m < meqn; ++m)
    qnew[m + (size_t)meqn * (i + (size_t)s.nx * j)] = s.out[m * s.plane + (long long)j * s.pitch + i];
  std::memcpy(cfl, &cfl_bits, sizeof(double));
  return 0;
}
}  // namespace

extern "C" int emul_step2_host(int meqn, int mwaves, int maux, int mx, int my, const double* qold, double* qnew,
                               const double* aux, double dx, double dy, double dt, const int* method7, const int* mthlim,
                               int ids, int use_fwave, int rp_id, const double* cparam, int ncparam, int variant,
                               int seg, double* cfl) {
  return step(ids, meqn, mwaves, maux, mx, my, qold, qnew, aux, dx, dy, dt, method7, mthlim, use_fwave, rp_id, cparam,
              ncparam, variant, seg, cfl);
}

// header words of the HotTiles buffer (entries of the three lists, their grid stamps) and the launch counter -- the tests
// check that the tile ordering was really exercised
extern "C" void emul_hot_header(unsigned* out7) { for (int k = 0; k < 6; ++k) out7[k] = g_hot[k]; out7[6] = g_hot_epoch; }
