// Kernel instantiations + launch heuristics.  Compiled twice (see csrc/build.py):
//   -DCLAW_NS=strict -DCLAW_STRICT=1 -fmad=false     op-for-op the reference (bit-exact vs oracle)
//   -DCLAW_NS=fast   -fmad=true                      FMA contraction allowed
#include <cstdlib>
#include "launch.h"

namespace clawcuda {
namespace CLAW_NS {

template <class RP, int NT, int MINB = 1>
static cudaError_t launch_one(int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st,
                              LaunchInfo* info) {
  using S = Smem2<RP, NT>;
  auto kern = (kind == 0) ? sweep2_unsplit<RP, NT, MINB> : sweep2_split<RP, NT, MINB>;
  const int smem = (int)S::BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int occ = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const int TX = NT - 4;
  const int strips = (P.mx + TX - 1) / TX;
  const bool x_only = (kind == 1) && !P.do_y;
  if (P.j1 <= P.j0) {                         // default: the whole grid
    P.j0 = x_only ? 0 : 2;
    P.j1 = x_only ? P.my + 4 : P.my + 2;
  }
  const int rows = P.j1 - P.j0;
  int seg = seg_hint;
  if (seg <= 0) {
    // enough CTAs for >= 6 waves of resident CTAs, but segments not shorter than 32 rows
    // (each segment pays 3-4 warm-up rows) nor longer than 256
    const long long want = 6LL * num_sms * occ;
    long long nseg = (want + strips - 1) / strips;
    if (nseg < 1) nseg = 1;
    seg = (int)((rows + nseg - 1) / nseg);
    if (seg < 32) seg = 32;
    if (seg > 256) seg = 256;
  }
  if (seg > rows) seg = rows;
  P.seg_rows = seg;
  dim3 grid(strips, (rows + seg - 1) / seg);
  kern<<<grid, NT, smem, st>>>(P);
  if (info) { info->grid_x = grid.x; info->grid_y = grid.y; info->block = NT; info->smem_bytes = smem; info->ctas_per_sm = occ; }
  return cudaGetLastError();
}

cudaError_t launch_sweep2(int rp_id, int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st,
                          LaunchInfo* info) {
  switch (rp_id) {
    case 1: return launch_one<RpAdvection2D, 128>(kind, P, seg_hint, num_sms, st, info);
    case 2: return launch_one<RpAcoustics2D, 128>(kind, P, seg_hint, num_sms, st, info);
    case 3: {
      static const int variant = getenv("CLAW_VARIANT") ? atoi(getenv("CLAW_VARIANT")) : 0;   // tuning experiments
      if (kind == 0 && variant == 3) return launch_one<RpEuler4Wave2D, 128, 3>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 4) return launch_one<RpEuler4Wave2D, 128, 4>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 5) return launch_one<RpEuler4Wave2D, 256, 1>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 6) return launch_one<RpEuler4Wave2D, 256, 2>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 7) return launch_one<RpEuler4Wave2D, 64, 4>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 8) return launch_one<RpEuler4Wave2D, 192, 2>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 2) return launch_one<RpEuler4Wave2D, 128, 2>(kind, P, seg_hint, num_sms, st, info);
      if (kind == 0 && variant == 0) return launch_one<RpEuler4Wave2D, 64, 6>(kind, P, seg_hint, num_sms, st, info);
      return launch_one<RpEuler4Wave2D, 128>(kind, P, seg_hint, num_sms, st, info);
    }
    case 4: {
      static const int variant = getenv("CLAW_VARIANT") ? atoi(getenv("CLAW_VARIANT")) : 0;   // tuning experiments
      if (variant == 2) return launch_one<RpShallowRoeEfix2D, 128, 2>(kind, P, seg_hint, num_sms, st, info);
      if (variant == 3) return launch_one<RpShallowRoeEfix2D, 128, 3>(kind, P, seg_hint, num_sms, st, info);
      if (variant == 4) return launch_one<RpShallowRoeEfix2D, 128, 4>(kind, P, seg_hint, num_sms, st, info);
      if (variant == 5) return launch_one<RpShallowRoeEfix2D, 64, 6>(kind, P, seg_hint, num_sms, st, info);
      if (variant == 6) return launch_one<RpShallowRoeEfix2D, 64, 8>(kind, P, seg_hint, num_sms, st, info);
      return launch_one<RpShallowRoeEfix2D, 128>(kind, P, seg_hint, num_sms, st, info);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace CLAW_NS
}  // namespace clawcuda
