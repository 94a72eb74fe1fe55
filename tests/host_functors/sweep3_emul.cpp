// TEST INFRASTRUCTURE.  The 3-D sweep KERNEL sources (clawpack_b200/csrc/sweep3.cuh: sweep3_x and the y- / z-march
// kernels, 32 x 8 threads per CTA, shared-memory exchange, operand prefetch ring) executed on the CPU by the fiber
// emulation of cuda_emul.h, behind entry points with the semantics of claw_step3_host / claw_step3ds_host
// (csrc/clawcuda.cu: launch_pass3 / launch_step3 on one device, no slab decomposition).  Strict arithmetic
// (-ffp-contract=off), or namespace fast with -DCLAW_HOST_FAST.  Nothing in the product links this file.
#define CLAW_HOST_CHECK 1
#ifdef CLAW_HOST_FAST
#define CLAW_NS fast
#else
#define CLAW_NS strict
#define CLAW_STRICT 1
#endif
#include "cuda_emul.h"
using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::copysign; using std::fmax;
#include "../../clawpack_b200/csrc/sweep2.cuh"
#include "../../clawpack_b200/csrc/rp3.cuh"
#include "../../clawpack_b200/csrc/sweep3.cuh"

using namespace clawcuda;
using namespace clawcuda::CLAW_NS;

namespace {
int pick_seg(int len, int seg_hint) { int s = seg_hint > 0 ? seg_hint : len; if (s > len) s = len; return s < 1 ? 1 : s; }

// grids of launch_x / launch_march in csrc/sweep3_inst.cu (the segment length is a test parameter here)
template <class RP, int SPEC = 0, int BY = 8> void pass_x(Sweep3Params& P, int seg_hint) {
  const bool split = P.m3 < 0;
  const int nx = P.n[0] + 4, ny = P.n[1] + 4, nz = P.n[2] + 4;
  const int len = split ? nz : P.n[2];
  P.seg = pick_seg(len, seg_hint);
  emul::Dim3 grid, block;
  grid.x = (nx - 2 + 27) / 28;
  grid.y = split ? (ny + BY - 1) / BY : (P.n[1] + BY - 3) / (BY - 2);
  grid.z = (len + P.seg - 1) / P.seg;
  block.x = 32; block.y = BY;
  const Sweep3Params Pc = P;
  emul::launch(grid, block, [&]() { sweep3_x<RP, BY, 1, SPEC>(Pc); });
}
template <class RP, int IXYZ, int SPEC = 0> void pass_march(Sweep3Params& P, int seg_hint) {
  constexpr int BX = 32, BY = 8, TD = (IXYZ == 1) ? 2 : 1;
  const bool split = P.m3 < 0;
  const int nx = P.n[0] + 4, nT = P.n[TD] + 4, nlen = P.n[IXYZ] + 4;
  const int len = split ? nlen : P.n[IXYZ];
  P.seg = pick_seg(len, seg_hint);
  emul::Dim3 grid, block;
  grid.x = split ? (nx + BX - 1) / BX : (P.n[0] + BX - 3) / (BX - 2);
  grid.y = split ? (nT + BY - 1) / BY : (P.n[TD] + BY - 3) / (BY - 2);
  grid.z = (len + P.seg - 1) / P.seg;
  block.x = BX; block.y = BY;
  const Sweep3Params Pc = P;
  emul::launch(grid, block, [&]() { sweep3_march<RP, IXYZ, BX, BY, 1, SPEC>(Pc); });
}
template <class RP, int SPEC = 0> void pass(int dir, Sweep3Params& P, int seg_hint) {
  if (dir == 0) pass_x<RP, SPEC>(P, seg_hint);
  else if (dir == 1) pass_march<RP, 1, SPEC>(P, seg_hint);
  else pass_march<RP, 2, SPEC>(P, seg_hint);
}
bool headline3(const Sweep3Params& P, int W, int lim) {              // headline_method3() of csrc/sweep3_inst.cu
  if (P.order == 1 || P.m3 != 2 || P.m4 != 2 || !P.limit) return false;
  for (int mw = 0; mw < W; ++mw) if (P.mthlim[mw] != lim) return false;
  return true;
}
int run_pass(int rp_id, int dir, Sweep3Params& P, int seg_hint, int variant) {
  switch (rp_id) {
    case 5:
      if (variant == 5 && headline3(P, 2, 4)) { pass<RpVcAcoustics3D, 41>(dir, P, seg_hint); return 0; }
      if (variant == 6 && headline3(P, 2, 4)) {                      // x-sweep with 32 x 6 tiles
        if (dir == 0) pass_x<RpVcAcoustics3D, 41, 6>(P, seg_hint); else pass<RpVcAcoustics3D, 41>(dir, P, seg_hint);
        return 0;
      }
      pass<RpVcAcoustics3D>(dir, P, seg_hint); return 0;
    case 6: pass<RpEuler3D>(dir, P, seg_hint); return 0;
    case 15: pass<RpBurgers3D>(dir, P, seg_hint); return 0;
    case 17: pass<RpVcAdvection3D>(dir, P, seg_hint); return 0;
  }
  return 5;
}
}  // namespace

// idir = 0: unsplit step3 (three passes into one output, interior read back); idir = 1..3: one step3ds pass (every cell)
extern "C" int emul_step3_host(int meqn, int mwaves, int maux, int mx, int my, int mz, const double* qold, double* qnew,
                               const double* auxbc, double dx, double dy, double dz, double dt, const int* method7,
                               const int* mthlim, int idir, int rp_id, const double* cparam, int ncparam, int seg_hint,
                               int variant, double* cfl) {
  const int nx = mx + 4, ny = my + 4, nz = mz + 4, pitch = nx + 1;
  const long long kstride = (long long)pitch * ny, plane = kstride * nz;
  auto to_soa = [&](const double* aos, int nc, std::vector<double>& soa) {
    soa.assign((size_t)plane * nc, 0.0);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) for (int m = 0; m < nc; ++m)
      soa[m * plane + k * kstride + (long long)j * pitch + i] = aos[m + (size_t)nc * (i + (size_t)nx * (j + (size_t)ny * k))];
  };
  std::vector<double> src, aux, dst((size_t)plane * meqn, -7.0e300), tmp;
  to_soa(qold, meqn, src);
  if (maux > 0) to_soa(auxbc, maux, aux);
  unsigned long long cfl_bits = 0;
  const double delta[3] = {dx, dy, dz};
  auto params = [&](int first_pass, const double* in, double* out) {     // as launch_pass3() in csrc/clawcuda.cu
    Sweep3Params P{};
    P.qin = in; P.qout = out; P.aux = maux > 0 ? aux.data() : nullptr; P.cfl_bits = &cfl_bits;
    P.n[0] = mx; P.n[1] = my; P.n[2] = mz;
    for (int d = 0; d < 3; ++d) P.dtd[d] = dt / delta[d];
    P.pitch = pitch; P.kstride = kstride; P.plane = plane;
    P.order = method7[1];
    if (method7[2] < 0) { P.m3 = -1; P.m4 = 0; } else { P.m3 = method7[2] / 10; P.m4 = method7[2] - 10 * P.m3; }
    P.limit = 0;
    for (int mw = 0; mw < mwaves; ++mw) { P.mthlim[mw] = mthlim[mw]; if (mthlim[mw] > 0) P.limit = 1; }
    P.first_pass = first_pass;
    P.capa = method7[5];
    P.zlo = 1; P.zhi = mz + 2;
    for (int k = 0; k < 4; ++k) P.rp.p[k] = (cparam && k < ncparam) ? cparam[k] : 0.0;
    return P;
  };
  int i0 = 0, i1 = nx, j0 = 0, j1 = ny, k0 = 0, k1 = nz;
  if (idir == 0) {                                                 // launch_step3, unsplit branch
    if (method7[2] < 0) return 41;
    for (int dir = 0; dir < 3; ++dir) {
      Sweep3Params P = params(dir == 0, src.data(), dst.data());
      if (int rc = run_pass(rp_id, dir, P, seg_hint, variant)) return rc;
    }
    i0 = j0 = k0 = 2; i1 = 2 + mx; j1 = 2 + my; k1 = 2 + mz;       // get_interior
  } else {                                                         // claw_step3ds_host: one pass, get_all
    if (method7[2] >= 0) return 41;
    Sweep3Params P = params(1, src.data(), dst.data());
    if (int rc = run_pass(rp_id, idir - 1, P, seg_hint, variant)) return rc;
  }
  for (int k = k0; k < k1; ++k) for (int j = j0; j < j1; ++j) for (int i = i0; i < i1; ++i) for (int m = 0; m < meqn; ++m)
    qnew[m + (size_t)meqn * (i + (size_t)nx * (j + (size_t)ny * k))] = dst[m * plane + k * kstride + (long long)j * pitch + i];
  std::memcpy(cfl, &cfl_bits, sizeof(double));
  return 0;
}
