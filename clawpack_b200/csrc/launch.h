// Declarations of the per-arithmetic-mode kernel launchers (defined in sweep_inst.cu, which is
// compiled twice: -fmad=false -> namespace strict, -fmad=true -> namespace fast).
#pragma once
#include <cuda_runtime.h>
#include "sweep2.cuh"

namespace clawcuda {
struct Sweep3Params;
struct LaunchInfo { int grid_x, grid_y, block, smem_bytes, ctas_per_sm; };
namespace strict {
// kind: 0 unsplit, 1 split (fused or single sweep per P.do_x / P.do_y)
cudaError_t launch_sweep2(int rp_id, int kind, Sweep2Params& P, int seg_rows_hint, int num_sms,
                          cudaStream_t st, LaunchInfo* info);
}
namespace fast {
cudaError_t launch_sweep2(int rp_id, int kind, Sweep2Params& P, int seg_rows_hint, int num_sms,
                          cudaStream_t st, LaunchInfo* info);
}
// 1-D (sweep1_inst.cu): step1; P.dtdx = dt, P.dtdy = dx, arrays in PyClaw's (m, i) order (plane = 1, pitch = meqn)
namespace strict {
cudaError_t launch_sweep1(int rp_id, Sweep2Params& P, cudaStream_t st, LaunchInfo* info);
}
namespace fast {
cudaError_t launch_sweep1(int rp_id, Sweep2Params& P, cudaStream_t st, LaunchInfo* info);
}
// 3-D (sweep3_inst.cu): dir 0/1/2 = x/y/z-sweep of step3 / step3ds
namespace strict {
cudaError_t launch_sweep3(int rp_id, int dir, Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info);
}
namespace fast {
cudaError_t launch_sweep3(int rp_id, int dir, Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info);
}
}  // namespace clawcuda
