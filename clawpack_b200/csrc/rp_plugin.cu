// User-supplied 2-D Riemann solvers as plug-ins.
//
// In the reference a Riemann solver reaches the stepper as two C pointers, `rpn2` / `rpt2`, taken from an f2py
// module the user compiled (`self.rp.rpn2._cpointer`, pyclaw/src/pyclaw/classic/solver.py:516-521;
// `solver.rp = my_module`).  A function pointer called per pencil is not a form a GPU kernel can use; the B200
// equivalent of "the user compiles a solver and hands it to the stepper" is: the user writes a DEVICE FUNCTOR with the
// interface of rp2.cuh / rp2_aux.cuh (point-wise `normal()` = rpn2 at one interface, `transverse()` = rpt2 of one
// fluctuation, in the rotated frame), and this file is compiled against it -- once per arithmetic mode, like the built-in
// solvers -- into a small shared object that instantiates the same warp-synchronous sweep kernels for that functor
// (sweep2_launch.cuh) and exports three C symbols.  `claw_rp_plugin_load()` (include/clawcuda.h) dlopen's it and
// returns an rp_id (>= CLAW_RP_PLUGIN_BASE) that claw_create / the f2py-shaped host calls accept like a built-in id.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -DCLAW_NS=strict -DCLAW_STRICT=1 -fmad=false \
//        -DCLAW_PLUGIN_HEADER='"my_solver.cuh"' -DCLAW_PLUGIN_FUNCTOR=MySolver -c rp_plugin.cu -o my_strict.o
//   nvcc ...                       -DCLAW_NS=fast -fmad=true  (same defines)           -c rp_plugin.cu -o my_fast.o
//   nvcc -shared -o libclawrp_my.so my_strict.o my_fast.o
// A 3-D solver: add -DCLAW_PLUGIN_NDIM=3; the functor then has the interface of rp3.cuh (normal<IXYZ>, transverse<IXYZ, ICOOR,
// TT>, the TV component traits, NCTX, MAUX) and the object exports claw_rp_plugin_launch3 instead of _launch2.
// (clawpack_b200.riemann.compile_cuda_solver does exactly this.)
//
// The functor lives in namespace clawcuda::CLAW_NS (so that ddiv / dsqrt / rcp_fast / fmax2 of rp2.cuh are in reach and
// the strict build is op-for-op IEEE) and defines
//   static constexpr int MEQN, MWAVES, NROE, NCELL;   [+ MAUX > 0: the aux-reading interface of rp2_aux.cuh]
//   normal(), speeds(), transverse()                  [+ cell() when NCELL > 0]
// Optional:  static constexpr int WARPS_PER_SM  (launch bound; default 8),
//            static constexpr bool NO_TRANSVERSE = true  (an rpt2_dummy solver: transverse_waves must be 0 or the step split).
#ifndef CLAW_PLUGIN_NDIM
#define CLAW_PLUGIN_NDIM 2
#endif
#if CLAW_PLUGIN_NDIM == 3
#include "sweep3_launch.cuh"          // 3-D functor interface: rp3.cuh (normal<IXYZ>, transverse<IXYZ, ICOOR, TT>, TV, NCTX, MAUX)
#else
#include "sweep2_launch.cuh"
#endif
#include <type_traits>

#ifndef CLAW_PLUGIN_HEADER
#error "compile with -DCLAW_PLUGIN_HEADER='\"your_functor.cuh\"' -DCLAW_PLUGIN_FUNCTOR=YourFunctor"
#endif
#include CLAW_PLUGIN_HEADER

namespace clawcuda {
namespace CLAW_NS {

using PluginRP = CLAW_PLUGIN_FUNCTOR;

template <class T, class = void> struct plugin_maux { static constexpr int value = 0; };
template <class T> struct plugin_maux<T, std::void_t<decltype(T::MAUX)>> { static constexpr int value = T::MAUX; };
template <class T, class = void> struct plugin_warps { static constexpr int value = 8; };
template <class T> struct plugin_warps<T, std::void_t<decltype(T::WARPS_PER_SM)>> { static constexpr int value = T::WARPS_PER_SM; };
template <class T, class = void> struct plugin_no_transverse { static constexpr bool value = false; };
template <class T> struct plugin_no_transverse<T, std::void_t<decltype(T::NO_TRANSVERSE)>> { static constexpr bool value = T::NO_TRANSVERSE; };

#if CLAW_PLUGIN_NDIM == 3
// dir 0 / 1 / 2 = the x- / y- / z-sweep of step3 (unsplit) or step3ds (P.m3 < 0)
cudaError_t plugin_launch3(int dir, Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info) {
  return launch_dir<PluginRP, 8, 1>(dir, P, num_sms, st, info);
}
#else
cudaError_t plugin_launch2(int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st, LaunchInfo* info) {
  if constexpr (plugin_maux<PluginRP>::value > 0) {
    return launch_aux<PluginRP, 2>(kind, P, seg_hint, num_sms, st, info);
  } else {
    constexpr int NWSM = plugin_warps<PluginRP>::value;
    if (kind == 1) return launch_split_ws<PluginRP, 1, NWSM>(P, seg_hint, num_sms, st, info);
    return launch_ws<PluginRP, 1, NWSM, 6>(P, seg_hint, num_sms, st, info);
  }
}
#endif

}  // namespace CLAW_NS
}  // namespace clawcuda

// The exported face of the plug-in: once per shared object (from the fast-mode translation unit).
#ifndef CLAW_STRICT
namespace clawcuda { namespace strict {
#if CLAW_PLUGIN_NDIM == 3
cudaError_t plugin_launch3(int dir, Sweep3Params& P, int num_sms, cudaStream_t st, LaunchInfo* info);
#else
cudaError_t plugin_launch2(int kind, Sweep2Params& P, int seg_hint, int num_sms, cudaStream_t st, LaunchInfo* info);
#endif
} }

extern "C" {
// layout version of Sweep2Params / Sweep3Params / LaunchInfo / StepCtl this object was compiled against (sweep2.cuh)
int claw_rp_plugin_abi(void) { return CLAW_PLUGIN_ABI; }
// info[0..5] = num_dim, MEQN, MWAVES, MAUX read by the solver, 1 if it has no transverse solver, sizeof(Sweep{2,3}Params)
void claw_rp_plugin_info(int* info) {
  using RP = clawcuda::fast::PluginRP;
  info[0] = CLAW_PLUGIN_NDIM; info[1] = RP::MEQN; info[2] = RP::MWAVES; info[3] = clawcuda::fast::plugin_maux<RP>::value;
  info[4] = clawcuda::fast::plugin_no_transverse<RP>::value ? 1 : 0;
#if CLAW_PLUGIN_NDIM == 3
  info[5] = (int)sizeof(clawcuda::Sweep3Params);
#else
  info[5] = (int)sizeof(clawcuda::Sweep2Params);
#endif
}
#if CLAW_PLUGIN_NDIM == 3
// mode: 0 strict, 1 fast (CLAW_MODE_*); dir: 0 / 1 / 2 = x- / y- / z-sweep
int claw_rp_plugin_launch3(int mode, int dir, void* params, int num_sms, void* stream, void* launch_info) {
  clawcuda::Sweep3Params& P = *static_cast<clawcuda::Sweep3Params*>(params);
  clawcuda::LaunchInfo* li = static_cast<clawcuda::LaunchInfo*>(launch_info);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return (int)(mode == 1 ? clawcuda::fast::plugin_launch3(dir, P, num_sms, st, li)
                         : clawcuda::strict::plugin_launch3(dir, P, num_sms, st, li));
}
#else
// mode: 0 strict, 1 fast (CLAW_MODE_*); kind: 0 unsplit (step2), 1 dimension-split (step2ds; P->do_x / do_y)
int claw_rp_plugin_launch2(int mode, int kind, void* params, int seg_hint, int num_sms, void* stream, void* launch_info) {
  clawcuda::Sweep2Params& P = *static_cast<clawcuda::Sweep2Params*>(params);
  clawcuda::LaunchInfo* li = static_cast<clawcuda::LaunchInfo*>(launch_info);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return (int)(mode == 1 ? clawcuda::fast::plugin_launch2(kind, P, seg_hint, num_sms, st, li)
                         : clawcuda::strict::plugin_launch2(kind, P, seg_hint, num_sms, st, li));
}
#endif
}
#endif
