/* clawcuda.h -- C ABI of libclawcuda.so: the B200 (sm_100a) classic wave-propagation step.
 *
 * This is the drop-in boundary for PyClaw's kernel plug-in point `solver.fmod`
 * (reference: pyclaw/src/pyclaw/classic/solver.py:208-211, 498-549, 648-706).  Everything is
 * `extern "C"`, plain pointers and sizes; no torch / CUDA types appear in a signature
 * (a CUDA stream is passed as an opaque `void*`, NULL = default stream).
 *
 * Two families of entry points:
 *
 *  (1) f2py-shaped, HOST buffers -- bind these where the reference binds the f2py module
 *      `clawpack.pyclaw.classic.classic2` / `classic3`:
 *        claw_step2_host    <->  step2    (pyclaw/src/pyclaw/classic/step2.f90:2-5)
 *        claw_step2ds_host  <->  step2ds  (pyclaw/src/pyclaw/classic/step2ds.f90:2-5)
 *        claw_step3_host    <->  step3    (pyclaw/src/pyclaw/classic/step3.f90:2-5)
 *        claw_step3ds_host  <->  step3ds  (pyclaw/src/pyclaw/classic/step3ds.f90:2-5)
 *        claw_step1_host    <->  step1    (pyclaw/src/pyclaw/classic/step1.f90:2-3; module classic1)
 *      Arrays are float64, Fortran-contiguous, WITH ghost cells, exactly as PyClaw passes
 *      them (qbc / auxbc).  The (rpn, rpt[, rptt]) C-function-pointer arguments of the
 *      reference (classic/solver.py:516-521) become `rp_id` + `cparam` (the COMMON /cparam/
 *      contents, reference pyclaw/src/pyclaw/state.py:192-217).  Host<->device copies happen
 *      inside the call.
 *
 *  (2) device-resident solver handle -- what ClawSolver2D/3D(kernel_language='CUDA') drives:
 *      q lives in HBM across steps; per step one ghost fill + one fused sweep kernel + an
 *      8-byte CFL read-back; accept/reject is a pointer swap (replaces the q_backup copy of
 *      pyclaw/src/pyclaw/solver.py:611,645).
 *
 * All functions return 0 on success, non-zero on error (never abort; the reference's Fortran
 * `stop` paths, step2.f90:66-71, become error codes).  claw_last_error() gives the message.
 * A handle is not thread-safe; distinct handles are independent.
 */
#ifndef CLAWCUDA_H
#define CLAWCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Riemann solver ids (clawpack.riemann module names, riemann/riemann/static.py) */
enum {
  CLAW_RP_ADVECTION_2D = 1,            /* riemann/src/rpn2_advection.f90, rpt2_advection.f90 */
  CLAW_RP_ACOUSTICS_2D = 2,            /* rpn2_acoustics.f90, rpt2_acoustics.f90 */
  CLAW_RP_EULER_4WAVE_2D = 3,          /* rpn2_euler_4wave.f90, rpt2_euler.f90 */
  CLAW_RP_SHALLOW_ROE_EFIX_2D = 4,     /* rpn2_shallow_roe_with_efix.f90, rpt2_shallow_roe_with_efix.f90 */
  CLAW_RP_VC_ACOUSTICS_3D = 5,         /* rpn3_vc_acoustics.f90, rpt3_vc_acoustics.f90, rptt3_vc_acoustics.f90 */
  CLAW_RP_EULER_3D = 6,                /* rpn3_euler.f90, rpt3_euler.f90, rptt3_euler.f90 */
  CLAW_RP_VC_ACOUSTICS_2D = 7,         /* rpn2_vc_acoustics.f90, rpt2_vc_acoustics.f90 (aux: density, sound speed) */
  CLAW_RP_VC_ADVECTION_2D = 8,         /* rpn2_vc_advection.f90, rpt2_vc_advection.f90 (aux: edge velocities) */
  CLAW_RP_EULER_5WAVE_2D = 9,          /* rpn2_euler_5wave.f90, rpt2_euler_5wave.f90 (Euler + tracer) */
  CLAW_RP_PSYSTEM_2D = 10,             /* rpn2_psystem.f90, rpt2_psystem.f90 (f-wave; aux: density, modulus, relation id, strain) */
  CLAW_RP_EULER_HLLE_2D = 11,          /* rpn2_euler_hlle.f90 + rpt2_dummy.f90 (no transverse solver) */
  CLAW_RP_SHALLOW_HLLE_2D = 12,        /* rpn2_shallow_hlle.f90 + rpt2_dummy.f90 (no transverse solver) */
  CLAW_RP_ACOUSTICS_MAPPED_2D = 13,    /* rpn2_acoustics_mapped.f90, rpt2_acoustics_mapped.f90 (9 aux planes, mapped grid) */
  CLAW_RP_SHALLOW_BATHYMETRY_FWAVE_2D = 14, /* rpn2_shallow_bathymetry_fwave.f90 (f-wave, aux: bathymetry; no transverse solver) */
  CLAW_RP_BURGERS_3D = 15,             /* rpn3_burgers.f90, rpt3_burgers.f90, rptt3_burgers.f90 (cparam = coeff(3)) */
  CLAW_RP_BURGERS_2D = 16,             /* rpn2_burgers.f90, rpt2_burgers.f90 */
  CLAW_RP_VC_ADVECTION_3D = 17,        /* rpn3_vc_advection.f90, rpt3_vc_advection.f90, rptt3_vc_advection.f90 (aux: edge velocities) */
  CLAW_RP_KPP_2D = 18,                 /* rpn2_kpp.f90 + rpt2_dummy.f90 (non-convex scalar flux; no transverse solver) */
  CLAW_RP_VC_ELASTICITY_2D = 19,       /* rpn2_vc_elasticity.f90, rpt2_vc_elasticity.f90 (aux: density, lambda, mu, cp, cs) */
  CLAW_RP_VC_ADVECTION_VECTOR_2D = 20, /* pyclaw/examples/advection_reaction_2d/rpn2_vc_advection_vector.f90 + rpt2_dummy.f90 (aux: edge velocities) */
  /* 1-D (claw_step1_host only) */
  CLAW_RP_ADVECTION_1D = 21,           /* riemann/src/rp1_advection.f90 */
  CLAW_RP_ACOUSTICS_1D = 22,           /* rp1_acoustics.f90 */
  CLAW_RP_BURGERS_1D = 23,             /* rp1_burgers.f90 */
  CLAW_RP_SHALLOW_ROE_EFIX_1D = 24,    /* rp1_shallow_roe_with_efix.f90 */
  CLAW_RP_SHALLOW_HLLE_1D = 25,        /* rp1_shallow_hlle.f90 */
  CLAW_RP_TRAFFIC_1D = 26,             /* rp1_traffic.f90 (cparam = umax) */
  CLAW_RP_ADVECTION_COLOR_1D = 27,     /* rp1_advection_color.f90 (aux: velocity) */
  CLAW_RP_ACOUSTICS_VARIABLE_1D = 28,  /* rp1_acoustics_variable.f90 (aux: impedance, sound speed) */
  CLAW_RP_SHALLOW_BATHYMETRY_FWAVE_1D = 29, /* rp1_shallow_bathymetry_fwave.f90 (f-wave; aux: bathymetry) */
  CLAW_RP_EULER_EFIX_1D = 30,          /* rp1_euler_with_efix.f90 (cparam = gamma) */
  CLAW_RP_EULER_HLLE_1D = 31,          /* rp1_euler_hlle.f90 (cparam = gamma) */
  CLAW_RP_NONLINEAR_ELASTICITY_FWAVE_1D = 32, /* rp1_nonlinear_elasticity_fwave.f90 (f-wave; aux: density, coefA, coefB) */
  CLAW_RP_CUBIC_1D = 33,               /* rp1_cubic.f90 */
  CLAW_RP_TRAFFIC_VC_1D = 34,          /* rp1_traffic_vc.f90 (aux: speed limit) */
  CLAW_RP_TRAFFIC_VC_FWAVE_1D = 35,    /* rp1_traffic_vc_fwave.f90 (f-wave; aux: speed limit) */
  CLAW_RP_TRAFFIC_VC_TRACER_1D = 36    /* rp1_traffic_vc_tracer.f90 (2 equations; aux: speed limit) */
};

/* User-supplied Riemann solvers (2-D and 3-D).  The reference takes them as two C pointers from a module the user compiled
 * (`self.rp.rpn2._cpointer`, `self.rp.rpt2._cpointer`: pyclaw/src/pyclaw/classic/solver.py:516-521, after
 * `solver.rp = my_module`).  Here the user compiles a device functor against clawpack_b200/csrc/rp_plugin.cu into a
 * shared object (see that file for the recipe and the functor interface) and registers it:
 *   claw_rp_plugin_load(path, &rp_id)   dlopen + layout check; rp_id >= CLAW_RP_PLUGIN_BASE is then accepted wherever
 *                                       a built-in rp_id is (claw_create, claw_step2_host / claw_step2ds_host, or
 *                                       claw_step3_host / claw_step3ds_host for a plug-in compiled with -DCLAW_PLUGIN_NDIM=3)
 *   claw_rp_plugin_info(rp_id, info)    info[0..4] = num_dim, meqn, mwaves, aux planes read, 1 if no transverse solver
 * Loading the same path twice returns the same id. */
#define CLAW_RP_PLUGIN_BASE 1000
int claw_rp_plugin_load(const char* path, int* rp_id);
int claw_rp_plugin_info(int rp_id, int* info5);

/* Boundary-condition ids = pyclaw.BC (pyclaw/src/pyclaw/solver.py:13-19) */
enum { CLAW_BC_CUSTOM = 0, CLAW_BC_EXTRAP = 1, CLAW_BC_PERIODIC = 2, CLAW_BC_WALL = 3,
       /* device-side stand-in for a BC.custom callback that writes one fixed state into the ghost cells of a face
          (e.g. incoming_shock, pyclaw/examples/euler_2d/shock_bubble_interaction.py:93-109); state via claw_set_bc_state */
       CLAW_BC_FIXED_STATE = 4 };

/* device-side source-term functors (claw_set_source): the `step_source` callbacks of reference examples, restated as
   kernels so that a run with a source term stays device-resident (Godunov splitting, classic/solver.py:126-144) */
enum { CLAW_SRC_NONE = 0,
       CLAW_SRC_EULER_RADIAL_2D = 1   /* step_Euler_radial, examples/euler_2d/shock_bubble_interaction.py:112-143;
                                         aux(1) = radius, params = (gamma) */ };

/* arithmetic mode: STRICT = no FMA contraction, op-for-op the reference (bit-exact vs the oracle);
   FAST = FMA contraction allowed (within the 1e-12/step tolerance) */
enum { CLAW_MODE_STRICT = 0, CLAW_MODE_FAST = 1 };

typedef struct claw_solver claw_solver;

const char* claw_version(void);
/* message of the last error on this handle (h may be NULL: last error of a failed create / host call) */
const char* claw_last_error(const claw_solver* h);
/* number of CUDA devices visible (0 when there is none; never fails) */
int claw_device_count(void);

/* ---- (2) device-resident solver ------------------------------------------------------ */

/* ndim 2|3; n[ndim] interior cells; num_ghost must be 2 (classic/solver.py:78); stream: cudaStream_t or NULL */
int claw_create(claw_solver** out, int ndim, int rp_id, int meqn, int mwaves, int maux,
                const int* n, int num_ghost, int device, void* stream);
int claw_destroy(claw_solver* h);

/* element counts (doubles) of the device buffers the caller must provide via claw_bind */
int claw_buffer_sizes(const claw_solver* h, size_t* q_elems, size_t* aux_elems, size_t* work_elems);
/* borrow caller-owned device buffers (e.g. torch tensors): two q buffers (ping-pong), aux, staging */
int claw_bind(claw_solver* h, double* d_q0, double* d_q1, double* d_aux, double* d_work);

/* COMMON /cparam/ values in the solver's declaration order (state.set_cparam, state.py:192-217) */
int claw_set_cparam(claw_solver* h, const double* cparam, int n);
/* method(7), mthlim(mwaves), use_fwave as packed by ClawSolver._set_method (classic/solver.py:173-193) */
int claw_set_method(claw_solver* h, const int* method7, const int* mthlim, int use_fwave);
/* BC ids per dimension; reflect_index = component negated by BC.wall (classic/solver.py:445,608) */
int claw_set_bc(claw_solver* h, const int* lower, const int* upper, const int* aux_lower,
                const int* aux_upper, const int* reflect_index);
/* which faces of this patch are physical boundaries (petclaw/geometry.py:21-41); default all */
int claw_set_boundary_ownership(claw_solver* h, const int* on_lower, const int* on_upper);
int claw_set_mode(claw_solver* h, int mode);
/* source functor run after the sweeps of every attempt under claw_ctl_* (skipped when that attempt's CFL >= cfl_max, as
   ClawSolver.step does, classic/solver.py:134); src_id 0 switches it off */
int claw_set_source(claw_solver* h, int src_id, const double* params, int nparams);
/* ghost-cell state (meqn values) of a CLAW_BC_FIXED_STATE face; side 0 = lower, 1 = upper */
int claw_set_bc_state(claw_solver* h, int dim, int side, const double* values);
/* tuning knobs: rows per march segment (0 = auto) */
int claw_set_tuning(claw_solver* h, int seg_rows);

/* host <-> device state transfer; host arrays are Fortran-ordered (meqn, mx, my[, mz]) WITHOUT ghosts
   (state.q / state.aux, pyclaw/src/pyclaw/state.py:316-321) */
int claw_put_q(claw_solver* h, const double* host_q);
int claw_get_q(claw_solver* h, double* host_q);
/* q of a few interior cells (0-based interior indices, ndim ints per cell) of the current state ->
 * host_out[c*meqn + m]: the gauge read of Solver.write_gauge_values (pyclaw/src/pyclaw/solver.py:680-715) without
 * a full-array download. */
int claw_get_cells(claw_solver* h, int ncells, const int* idx, double* host_out);
int claw_put_aux(claw_solver* h, const double* host_aux);
/* same WITH ghost cells (qbc / auxbc, solver.py:341-347) */
int claw_put_qbc(claw_solver* h, const double* host_qbc);
int claw_get_qbc(claw_solver* h, double* host_qbc);
int claw_put_auxbc(claw_solver* h, const double* host_auxbc);

/* ghost fill of the current q (and aux) -- Solver._apply_bcs (solver.py:351-437) */
int claw_apply_bcs(claw_solver* h);

/* one hyperbolic step attempt: [ghost fill] + fused sweep + CFL max.  The new state goes to the
   alternate buffer; call claw_accept to make it current or claw_reject to drop it
   (ClawSolver2D/3D.step_hyperbolic + evolve_to_time accept/reject, solver.py:605-657).
   delta[ndim] = cell sizes.  fill_ghosts != 0 runs claw_apply_bcs first. */
int claw_step_launch(claw_solver* h, const double* delta, double dt, int fill_ghosts);
/* device address of the 8-byte CFL result of the last launch (a double; for an NCCL all-reduce) */
int claw_cfl_device_ptr(claw_solver* h, void** d_cfl);
/* wait for the launch and return its CFL number */
int claw_cfl_wait(claw_solver* h, double* cfl);
/* the CFL slot as it is on the device after everything queued so far (e.g. a reduction across ranks behind
   claw_step_launch); waits for the stream */
int claw_cfl_read(claw_solver* h, double* cfl);
/* claw_step_launch + claw_cfl_wait */
int claw_step(claw_solver* h, const double* delta, double dt, int fill_ghosts, double* cfl);
int claw_accept(claw_solver* h);
int claw_reject(claw_solver* h);
int claw_sync(claw_solver* h);

/* ---- device-resident step control -----------------------------------------------------------------------------------
   Solver.evolve_to_time / get_dt / accept_reject_step (pyclaw/src/pyclaw/solver.py:524-666) driven from the device: the
   loop state -- t, dt, the CFL number of the running attempt, which q buffer is current, solver.status, whether the loop
   has ended -- lives in a 32-word control block at the start of the work buffer (word 0 is the CFL slot, see
   claw_cfl_device_ptr).  The host queues step attempts WITHOUT reading the CFL number back: the kernels of attempt n+1
   take dt and the current buffer from what the control kernel of attempt n decided, a rejected attempt is simply
   retried with the new dt from the unchanged buffer (no q_backup copy), and attempts queued past the end of the loop are
   no-ops.  The arithmetic of dt, t and accept/reject is the reference's operation by operation (same bits).

   Control block as 64-bit words (doubles unless noted):
     0 CFL slot   1 t   2 dt   3 dt_max   4 cfl_desired   5 cfl_max   6 cfl of the last attempt   7 tstart   8 tend
     9 told   10 status dtmin   11 status dtmax   12 status cflmax   13 status numsteps (int64)   14 current buffer (int64)
     15 done (int64)   16 error (int64: 1 = CFL too large with fixed dt, 2 = max_steps reached)   17 take_one_step
     18 dt_variable   19 dt_old is set   20 loop counter n   21 accepted steps of this call   22 max_steps
     23 attempts logged so far   24..31 reserved                                  (17..23 int64)

   claw_ctl_begin   start of one evolve_to_time call: tstart = t, done = 0, then the get_dt of the first attempt.
                    host_words != NULL uploads the host's values of words 1..6, 10..13, 18, 19 first (required on the
                    first call; later calls pass NULL to carry on from the device's state without a read-back).
   claw_ctl_attempts  queue n attempts: each = [ghost fill] + sweep(s) [+ source functor] [+ peer-memory copies] + control
                    kernel, through a CUDA graph of one attempt captured once per configuration (single-rank runs, and
                    multi-rank runs under the peer-memory transport, claw_peer_*).
   claw_ctl_sweep / claw_ctl_end  the two halves of ONE attempt for multi-rank runs: the caller all-reduces (max) the
                    CFL slot across ranks in between (ncclAllReduce on the same stream; petclaw/cfl.py:36).
   claw_ctl_read    wait for everything queued; returns the control block and the log entries (dt, cfl, accepted) of
                    attempts [from_attempt, word 23) -- at most max_entries, the newest ones if more are pending. */
int claw_ctl_begin(claw_solver* h, const double* host_words, double tend, int take_one_step, long long max_steps);
int claw_ctl_attempts(claw_solver* h, const double* delta, int n, int fill_ghosts);
int claw_ctl_sweep(claw_solver* h, const double* delta, int fill_ghosts);
int claw_ctl_end(claw_solver* h);
int claw_ctl_read(claw_solver* h, double* words32, long long from_attempt, double* log3, int max_entries,
                  int* n_entries);
/* CUDA graph for claw_ctl_attempts on (default) / off */
int claw_ctl_use_graph(claw_solver* h, int on);

/* slab-decomposition ghost exchange along the slowest dimension (replaces DMDA.globalToLocal,
   petclaw/state.py:223-231): pack the 2 interior rows/planes next to face `side` (0 lower, 1 upper)
   of the current q into a contiguous device buffer / unpack a received buffer into that face's ghosts */
int claw_halo_elems(const claw_solver* h, size_t* elems);
int claw_halo_pack(claw_solver* h, int side, double* d_buf);
int claw_halo_unpack(claw_solver* h, int side, const double* d_buf);
/* the same on the buffer the last claw_step_launch wrote (current after claw_accept): the next step's exchange can be
   queued right behind the sweep, before the host has read the CFL number; a rejected step simply discards it */
int claw_halo_pack_next(claw_solver* h, int side, double* d_buf);
int claw_halo_unpack_next(claw_solver* h, int side, const double* d_buf);
/* the same for the (static) aux array: exchanged once at setup (petclaw/state.py:233-241) */
int claw_halo_elems_aux(const claw_solver* h, size_t* elems);
int claw_halo_pack_aux(claw_solver* h, int side, double* d_buf);
int claw_halo_unpack_aux(claw_solver* h, int side, const double* d_buf);

/* ---- ghost exchange and CFL reduction over NCCL, for hosts that are not Python (an MPI / C driver) ----------------
   What PetClaw does with DMDA globalToLocal (petclaw/state.py:223-231) and an MPI max-reduce (petclaw/cfl.py:22,36),
   queued on the handle's stream: pack kernels -> one grouped ncclSend/ncclRecv with each slab neighbour -> unpack
   kernels; ncclAllReduce(max) in place on the 8-byte CFL slot.  `nccl_comm` is the caller's ncclComm_t (one rank per
   GPU).  NCCL is resolved at run time (dlopen of libnccl.so.2: the library itself has no link dependency on it).
   Neighbours as ranks of that communicator, -1 = none: channel A carries my upper interior rows / planes to the upper
   neighbour's lower ghosts, channel B my lower interior rows / planes to the lower neighbour's upper ghosts (a periodic
   slab dimension wraps by naming the rank at the other end; petclaw/util.py: slab_neighbours).  `aux` != 0 exchanges the
   static aux array instead of q (once, after claw_put_aux). */
int claw_halo_exchange(claw_solver* h, void* nccl_comm, int send_up_to, int recv_lo_from, int send_lo_to,
                       int recv_up_from, int aux);
int claw_allreduce_cfl(claw_solver* h, void* nccl_comm);
/* building a communicator without an MPI launcher: rank 0 creates the 128-byte id, the host broadcasts it by its own
   means, every rank calls claw_nccl_comm_init (ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy) */
int claw_nccl_unique_id(void* id128);
int claw_nccl_comm_init(void** nccl_comm, int nranks, int rank, const void* id128, int device);
int claw_nccl_comm_destroy(void* nccl_comm);

/* ---- peer-memory transport: ghost cells and CFL numbers over NVLink without a collective library -----------------
   One process per GPU on ONE node.  Every rank exports IPC handles of its (library-owned) q / aux buffers and of a small
   mailbox (claw_peer_export -> CLAW_PEER_INFO_BYTES bytes), the host gathers the blocks of all ranks by its own means,
   and claw_peer_connect maps the slab neighbours' buffers and everybody's mailbox (cudaIpcOpenMemHandle).  From then on
   an attempt under claw_ctl_* needs no exchange call at all: after the sweep(s) the rows / planes next to a rank face are
   written straight into the neighbour's ghost cells of the buffer that becomes current on accept, and the control kernel
   all-reduces (max) the CFL number through the mailboxes before it decides -- identically on every rank.  So
   claw_ctl_attempts (and its CUDA graph) serve N ranks like one.  claw_peer_sync refreshes the neighbours' ghost cells of
   the CURRENT buffer (aux != 0: of the aux array) and ends with a barrier across ranks: call it on EVERY rank after the
   initial data (or any later host upload) are in place.  claw_peer_allreduce_cfl is the reduction alone, for host-driven
   attempts (claw_step_launch).  Replaces DMDA globalToLocal / Vec.max of PetClaw (petclaw/state.py:223-231, cfl.py:36). */
#define CLAW_PEER_INFO_BYTES 320
int claw_own_buffers(claw_solver* h);          /* allocate q0 / q1 / aux / work inside the library (instead of claw_bind) */
int claw_peer_export(claw_solver* h, void* info);
int claw_peer_connect(claw_solver* h, int nranks, int rank, const void* infos_of_all_ranks, int send_up_to, int send_lo_to);
int claw_peer_sync(claw_solver* h, int aux);
int claw_peer_allreduce_cfl(claw_solver* h, double* cfl);

/* time `reps` back-to-back launches of the sweep kernel alone (no ghost fill, no swap) with CUDA
   events on the handle's stream; returns the average duration in milliseconds (bench.py roofline) */
int claw_time_sweep(claw_solver* h, const double* delta, double dt, int reps, double* avg_ms);

/* geometry of the last sweep launch: grid.x, grid.y(*z), threads per CTA, dynamic shared memory bytes, resident CTAs
   per SM (occupancy calculator) -- diagnostics for bench.py / profiles */
int claw_last_launch_info(const claw_solver* h, int* info5);
/* Dense-data switch of the 2-D kernels with the uniform-region fast path: state3 = { 1 if the stepping
 * currently uses the kernels WITHOUT the fast path (most tiles were heavy), 1 if the last claw_time_sweep timed those,
 * 1 if the switch is enabled (CLAW_DENSE_SWITCH != 0) }. */
int claw_dense_state(const claw_solver* h, int* state3);
/* kernels launched by this handle so far (bench.py `gpu_launches`) */
int claw_launch_count(const claw_solver* h, uint64_t* count);

/* ---- (1) f2py-shaped plug-in calls on HOST buffers ------------------------------------ */

/* requires qnew == qold on the interior on entry (as PyClaw always calls it, classic/solver.py:514);
   on return the interior of qnew holds the updated cells; ghost cells of qnew are left as passed
   (the reference leaves partial sums there which PyClaw never reads). */
int claw_step2_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my,
                    const double* qold, double* qnew, const double* aux, double dx, double dy, double dt,
                    const int* method7, const int* mthlim, int use_fwave,
                    int rp_id, const double* cparam, int ncparam, int mode, double* cfl);
/* ids = 1: x-sweeps of all rows; ids = 2: y-sweeps of all columns (qold may alias qnew) */
int claw_step2ds_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my,
                      const double* qold, double* qnew, const double* aux, double dx, double dy, double dt,
                      const int* method7, const int* mthlim, int ids, int use_fwave,
                      int rp_id, const double* cparam, int ncparam, int mode, double* cfl);
int claw_step3_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my, int mz,
                    const double* qold, double* qnew, const double* aux, double dx, double dy, double dz, double dt,
                    const int* method7, const int* mthlim, int use_fwave,
                    int rp_id, const double* cparam, int ncparam, int mode, double* cfl);
int claw_step3ds_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my, int mz,
                      const double* qold, double* qnew, const double* aux, double dx, double dy, double dz, double dt,
                      const int* method7, const int* mthlim, int idir, int use_fwave,
                      int rp_id, const double* cparam, int ncparam, int mode, double* cfl);
/* 1-D: claw_step1_host <-> step1 (pyclaw/src/pyclaw/classic/step1.f90:2-3 as called at classic/solver.py:302):
   q(meqn, mx+2*num_ghost) with ghost cells is updated IN PLACE (interior and the ghost cells the Fortran touches),
   aux(maux, mx+2*num_ghost) may be NULL when maux == 0; the f2py scratch arguments (f, wave, s, amdq, apdq, dtdx) are
   internal.  1-D Riemann solver ids: CLAW_RP_*_1D above, 21..36 (riemann/src/rp1_*.f90); solvers that read aux
   (27, 28, 29, 32, 34, 35, 36) fail with code 7 when aux is NULL or has too few components. */
int claw_step1_host(int meqn, int mwaves, int maux, int num_ghost, int mx, double* q, const double* aux,
                    double dx, double dt, const int* method7, const int* mthlim, int use_fwave,
                    int rp_id, const double* cparam, int ncparam, int mode, double* cfl);
/* drop the cached device state of the host entry points */
int claw_host_release(void);

/* pinned host memory for fast transfers of the (1)-family buffers */
int claw_host_alloc(void** p, size_t bytes);
int claw_host_free(void* p);

/* FP64 issue-rate micro-benchmark (DFMA chain on all SMs): returns measured G DFMA/s; used by
   bench.py as the roofline denominator of the FP64-bound kernels */
int claw_measure_fp64_peak(int device, double* gdfma_per_s);

#ifdef __cplusplus
}
#endif
#endif /* CLAWCUDA_H */
