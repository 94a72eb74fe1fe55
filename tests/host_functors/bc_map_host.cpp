// Host-compiled check of the one-launch ghost fill (clawpack_b200/csrc/bc_map.h): runs the index map over every
// ghost-frame cell of a 2-D SoA array exactly as the device kernel bc_xy2 does (gather into a scratch copy first, so that
// an accidental read of a written cell would show), tests/test_host.py compares with the reference's x-then-y slicing.
#include <cstring>
#include <vector>
#include "../../clawpack_b200/csrc/bc_map.h"

extern "C" int hostfun_bc_xy2(double* q, int ncomp, int mx, int my, const int* lo, const int* up, const int* lo_on,
                              const int* up_on, const int* reflect, const double* xlo, const double* xup,
                              const double* ylo, const double* yup) {
  using namespace clawcuda;
  const int pitch = mx + 4;
  const long long plane = (long long)pitch * (my + 4);
  const long long ncell = 4LL * (my + 4) + 4LL * mx;
  std::vector<double> val((size_t)(ncell * ncomp));
  std::vector<char> wr((size_t)(ncell * ncomp));
  long long seen = 0;
  for (long long idx = 0; idx < ncell * ncomp; ++idx) {
    int i, j;
    if (!bc_frame_cell(idx % ncell, mx, my, i, j)) return 1;
    const int m = (int)(idx / ncell);
    bool wrote;
    val[(size_t)idx] = bc_xy_value(q + m * plane, pitch, i, j, mx, my, lo, up, lo_on, up_on, m == reflect[0],
                                   m == reflect[1], xlo[m], xup[m], ylo[m], yup[m], wrote);
    wr[(size_t)idx] = wrote ? 1 : 0;
    ++seen;
  }
  int dummy_i, dummy_j;
  if (bc_frame_cell(ncell, mx, my, dummy_i, dummy_j)) return 2;        // the enumeration ends where the kernel's grid does
  for (long long idx = 0; idx < ncell * ncomp; ++idx) {
    int i, j;
    bc_frame_cell(idx % ncell, mx, my, i, j);
    if (wr[(size_t)idx]) q[(idx / ncell) * plane + (long long)j * pitch + i] = val[(size_t)idx];
  }
  return seen == ncell * ncomp ? 0 : 3;
}
