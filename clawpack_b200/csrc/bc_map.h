// Ghost cells of the reference's built-in boundary conditions as an index map.
//
// Solver._bc_lower / _bc_upper (pyclaw/src/pyclaw/solver.py:439-519) fill the two ghost cells of a face by slicing:
//   extrap    qbc[..., k] = qbc[..., num_ghost]            (lower)   /  qbc[..., -num_ghost - 1]   (upper)
//   periodic  qbc[..., k] = qbc[..., n + k]                (lower)   /  qbc[..., k - n]            (upper)
//   wall      qbc[..., k] = qbc[..., 2*num_ghost - 1 - k]  (lower), mirror image (upper); the normal velocity negated
// dimension by dimension (x first, then y: `_apply_bcs`, solver.py:380-437), so that a corner cell holds the y-rule
// applied to the x-filled column.  Every rule reads ONE interior cell, hence the two passes compose into one map
//   value(i, j) = sign_y * sign_x * q(src_x(i), src_y(j))      (or the fixed state of a CLAW_BC_FIXED_STATE face)
// and a single launch can fill all ghost cells of a 2-D array without an ordering between its threads: no thread reads a
// cell that another one writes.  Shared by the device kernel (clawcuda.cu: bc_xy2) and its host-compiled check
// (tests/host_functors/bc_map_host.cpp).
#pragma once

#ifndef CLAW_HD
#ifdef __CUDACC__
#define CLAW_HD __host__ __device__ __forceinline__
#else
#define CLAW_HD inline
#endif
#endif

namespace clawcuda {

// BC type codes of include/clawcuda.h (kept numerically identical; the header is not included here so that the host check
// compiles without it)
enum { BCM_CUSTOM = 0, BCM_EXTRAP = 1, BCM_PERIODIC = 2, BCM_WALL = 3, BCM_FIXED = 4 };

struct BcSrc {
  int active;   // 1: index k is a ghost cell this launch fills
  int src;      // cell along the same dimension whose value it takes (an interior cell)
  int negate;   // 1: BC.wall -- the component normal to the face changes sign
  int fixed;    // 0: copy; 1 / 2: the fixed state of the lower / upper face
};

// k in [0, n + 4): index along one dimension of an array with two ghost cells per side and n interior cells
CLAW_HD BcSrc bc_src(int k, int n, int lo_type, int up_type, bool lo_on, bool up_on) {
  BcSrc r;
  r.active = 0; r.src = k; r.negate = 0; r.fixed = 0;
  if (k < 2) {
    if (!lo_on) return r;
    if (lo_type == BCM_EXTRAP) { r.active = 1; r.src = 2; }
    else if (lo_type == BCM_PERIODIC) { if (up_on) { r.active = 1; r.src = n + k; } }
    else if (lo_type == BCM_WALL) { r.active = 1; r.src = 3 - k; r.negate = 1; }
    else if (lo_type == BCM_FIXED) { r.active = 1; r.fixed = 1; }
  } else if (k >= n + 2) {
    if (!up_on) return r;
    if (up_type == BCM_EXTRAP) { r.active = 1; r.src = n + 1; }
    else if (up_type == BCM_PERIODIC) { if (lo_on) { r.active = 1; r.src = k - n; } }
    else if (up_type == BCM_WALL) { r.active = 1; r.src = 2 * n + 3 - k; r.negate = 1; }
    else if (up_type == BCM_FIXED) { r.active = 1; r.fixed = 2; }
  }
  return r;
}

// The c-th ghost-frame cell of an (mx + 4) x (my + 4) array: first the four ghost columns over all rows, then the four
// ghost rows over the interior columns.  Returns false past the end.
CLAW_HD bool bc_frame_cell(long long c, int mx, int my, int& i, int& j) {
  const int ny = my + 4;
  if (c < 4LL * ny) {
    const int gi = (int)(c / ny);
    i = gi < 2 ? gi : mx + gi;
    j = (int)(c % ny);
    return true;
  }
  c -= 4LL * ny;
  if (c >= 4LL * mx) return false;
  const int gj = (int)(c / mx);
  i = 2 + (int)(c % mx);
  j = gj < 2 ? gj : my + gj;
  return true;
}

// Value of ghost-frame cell (i, j) of component plane `qm` (row pitch `pitch`) after the x-then-y fill; `wrote` = false
// when neither pass touches the cell.  reflect_x / reflect_y: whether this component is the one BC.wall negates in x / y.
CLAW_HD double bc_xy_value(const double* qm, int pitch, int i, int j, int mx, int my, const int* lo, const int* up,
                           const int* lo_on, const int* up_on, bool reflect_x, bool reflect_y, double xlo_val,
                           double xup_val, double ylo_val, double yup_val, bool& wrote) {
  const BcSrc ax = bc_src(i, mx, lo[0], up[0], lo_on[0] != 0, up_on[0] != 0);
  const BcSrc ay = bc_src(j, my, lo[1], up[1], lo_on[1] != 0, up_on[1] != 0);
  wrote = ax.active || ay.active;
  if (!wrote) return 0.0;
  if (ay.active && ay.fixed) return ay.fixed == 1 ? ylo_val : yup_val;
  const int jj = ay.active ? ay.src : j;
  double v;
  if (ax.active) {
    if (ax.fixed) v = ax.fixed == 1 ? xlo_val : xup_val;
    else {
      v = qm[(long long)jj * pitch + ax.src];
      if (ax.negate && reflect_x) v = -v;
    }
  } else {
    v = qm[(long long)jj * pitch + i];
  }
  if (ay.active && ay.negate && reflect_y) v = -v;
  return v;
}

}  // namespace clawcuda
