// libclawcuda.so -- C ABI (include/clawcuda.h) over the fused sm_100a sweep kernels.
//
// Device data layout (2-D): structure-of-arrays planes, q[m][J][I] with I fastest, row pitch
// padded to a multiple of 16 doubles, ghost cells included (array coords I in [0, mx+4)).
// Two q buffers are ping-ponged: a step reads `cur`, writes the interior of `alt`; accept
// swaps them, reject keeps `cur` (this replaces PyClaw's q_backup copy, solver.py:611,645, and
// the qold = qbc.copy of classic/solver.py:514).
#include "../../include/clawcuda.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

#include "launch.h"
#include "bc_map.h"
#include "sweep3.cuh"

using namespace clawcuda;

namespace {

thread_local std::string g_last_error;

struct Dev2 {  // geometry of the device arrays
  int ndim = 2;
  int n[3] = {1, 1, 1};      // interior cells
  int ng = 2;
  int nt[3] = {1, 1, 1};     // with ghosts
  int pitch = 0;             // row pitch (doubles)
  long long kstride = 0;     // 3-D: doubles per K plane (pitch * nt[1]); 0 in 2-D
  long long plane = 0;       // doubles per component plane
};

}  // namespace

struct claw_solver {
  Dev2 g;
  int rp_id = 0, meqn = 0, mwaves = 0, maux = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  double* q[2] = {nullptr, nullptr};
  double* aux = nullptr;
  double* work = nullptr;    // [0..8): CFL slot + scratch; [8..): AoS staging
  bool owns_buffers = false;
  int cur = 0;
  size_t q_elems = 0, aux_elems = 0, work_elems = 0;
  double cparam[4] = {0, 0, 0, 0};
  int method[7] = {1, 2, -1, 0, 0, 0, 0};
  int mthlim[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int use_fwave = 0;
  int bc_lower[3] = {1, 1, 1}, bc_upper[3] = {1, 1, 1};
  int aux_bc_lower[3] = {1, 1, 1}, aux_bc_upper[3] = {1, 1, 1};
  int reflect_index[3] = {1, 2, 3};
  int on_lower[3] = {1, 1, 1}, on_upper[3] = {1, 1, 1};
  int mode = CLAW_MODE_STRICT;
  int seg_rows = 0;
  // device-side callbacks (claw_set_source / claw_set_bc_state): id of the source-term functor run after the sweeps of
  // an attempt (Godunov splitting, classic/solver.py:137-144) and its parameters; ghost-cell states of BC type
  // CLAW_BC_FIXED_STATE per dimension and side
  int src_id = 0;
  double src_params[4] = {0, 0, 0, 0};
  // tile ordering of the 2-D kernels with the uniform-region fast path (sweep2_ws.cuh, HotTiles): small device buffer
  // owned by the handle; capacity of one tile list / one flag array in words
  unsigned* hot = nullptr;
  int hot_cap = 0, hot_tiles = 0;
  unsigned hot_epoch = 0;
  // Dense-data switch (dense_update): the fast-path kernels report how many tiles took the full path; when that is most
  // of the grid the launch code takes the kernels without the fast path (14-20 % faster there), and probes again later
  unsigned* h_hot = nullptr;         // pinned copy of the 8 header words of `hot`
  cudaEvent_t ev_hot = nullptr;
  bool hot_copy_pending = false, hot_copy_from_quiet = false, dense_probe = false;
  int dense_pref = 0;
  long long last_ntiles = 0, dense_batches = 0;
  int last_hot_rows = 0;
  int time_sweep_dense = 0;          // which kernel the last claw_time_sweep timed (reported by claw_last_launch_info)
  // send / receive buffers of claw_halo_exchange (send up, receive lower, send lower, receive upper), q and aux
  double* hx[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t hx_elems = 0;
  // peer-memory transport (claw_peer_*): the slab neighbours' q / aux buffers and every rank's mailbox, mapped into
  // this process through CUDA IPC; rows / planes next to a rank face are written straight into the neighbour's ghost
  // cells over NVLink, the CFL numbers are exchanged through the mailboxes
  struct Peer {
    bool active = false;
    int nranks = 1, rank = 0;
    int up = -1, lo = -1;                      // ranks whose LOWER / UPPER ghost cells I fill (channels A / B)
    double* up_q[2] = {nullptr, nullptr};      // the upper neighbour's two q buffers
    double* lo_q[2] = {nullptr, nullptr};
    double* up_aux = nullptr; double* lo_aux = nullptr;
    long long up_plane = 0, lo_plane = 0;      // their plane strides (slabs may differ in height)
    int lo_nlast = 0;                          // interior cells of the lower neighbour along the slab dimension
    unsigned long long* mail = nullptr;        // my mailbox (device memory owned by the handle)
    unsigned long long* peer_mail[64] = {};    // every rank's mailbox ([rank] = mine)
    void* opened[64][4] = {};                  // base pointers returned by cudaIpcOpenMemHandle (q0, q1, aux, mailbox)
  } peer;
  double bc_state[3][2][8] = {};
  double* h_cfl = nullptr;   // pinned
  cudaEvent_t ev_cfl = nullptr;
  // transfer/compute pipeline of the f2py-shaped host entry points
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[64] = {}, ev_cmp[64] = {};
  uint64_t launches = 0;
  LaunchInfo last_launch{};
  std::string err;
  // device-resident step control (claw_ctl_*): attempts queued since the last claw_ctl_read; the device's `cur` is
  // authoritative while ctl_mode is set and h->cur mirrors it at every read
  bool ctl_mode = false;
  bool ctl_pending = false;
  double* h_ctl = nullptr;   // pinned mirror of the control block + log ring
  // one step attempt captured as a CUDA graph (single-rank runs): ghost fill + sweep(s) + control kernel
  cudaGraphExec_t graph = nullptr;
  cudaStream_t s_cap = nullptr;
  uint64_t graph_nodes = 0;
  uint64_t cfg_epoch = 0, graph_epoch = 0;
  double graph_delta[3] = {0, 0, 0};
  int graph_fill = -1;
  bool aux_ghosts_valid = false;     // the ghost cells of the (static) aux array have been filled since it last changed
  int use_graph = 1;
};

namespace {

int fail(claw_solver* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (h) h->err = buf;
  return code;
}

#define CK(h, call)                                                                          \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(h, 100 + (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                       \
  } while (0)

// Header of the `work` buffer: the step-control block (StepCtl, 32 words; word 0 = the CFL slot) followed by the
// log ring of the attempted steps, three doubles (dt, cfl, accepted) per attempt.
constexpr int CTL_WORDS = 32;
constexpr int LOG_CAP = 4096;
constexpr int WORK_HDR = CTL_WORDS + 3 * LOG_CAP;

// ---------------------------------------------------------------------------------------
// layout conversion: host/staging AoS (m fastest, Fortran order) <-> device SoA planes
// region [i0,i1) x [j0,j1) of the ghosted array; staging holds exactly that region.
// ---------------------------------------------------------------------------------------
__global__ void aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int ncomp, int i0, int i1,
                           int j0, int j1, int k0, int k1, int pitch, long long kstride, long long plane) {
  const int w = i1 - i0, hgt = j1 - j0;
  const long long ncell = (long long)w * hgt * (k1 - k0);
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
       c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % w), j = (int)((c / w) % hgt), k = (int)(c / ((long long)w * hgt));
    const double* src = aos + c * ncomp;
    double* dst = soa + (long long)(k0 + k) * kstride + (long long)(j0 + j) * pitch + (i0 + i);
    for (int m = 0; m < ncomp; ++m) dst[m * plane] = src[m];
  }
}

__global__ void soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int ncomp, int i0, int i1,
                           int j0, int j1, int k0, int k1, int pitch, long long kstride, long long plane) {
  const int w = i1 - i0, hgt = j1 - j0;
  const long long ncell = (long long)w * hgt * (k1 - k0);
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
       c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % w), j = (int)((c / w) % hgt), k = (int)(c / ((long long)w * hgt));
    const double* src = soa + (long long)(k0 + k) * kstride + (long long)(j0 + j) * pitch + (i0 + i);
    double* dst = aos + c * ncomp;
    for (int m = 0; m < ncomp; ++m) dst[m] = src[m * plane];
  }
}

// ---------------------------------------------------------------------------------------
// ghost fill -- Solver._bc_lower/_bc_upper (pyclaw/src/pyclaw/solver.py:439-519), 2-D.
// x strips first (all rows, ghost rows included), then y strips (all columns, x-ghosts
// included), so corners are the composition x-then-y exactly as the numpy slicing does.
// reflect < 0: no negation (aux arrays / non-wall).
// ---------------------------------------------------------------------------------------
struct BcState { double lo[8], up[8]; };    // ghost-cell state per component of a CLAW_BC_FIXED_STATE face

__device__ __forceinline__ void bc_fill_pair(double* base, long long stride, int n, int lo_type, int up_type,
                                             bool lo_on, bool up_on, bool negate, double lo_val = 0.0,
                                             double up_val = 0.0) {
  // base[k*stride], k in [0, n+4): cells along the BC dimension
  if (lo_on) {
    if (lo_type == CLAW_BC_EXTRAP) {
      const double v = base[2 * stride];
      base[0] = v; base[stride] = v;
    } else if (lo_type == CLAW_BC_PERIODIC) {
      if (up_on) { base[0] = base[(long long)n * stride]; base[stride] = base[(long long)(n + 1) * stride]; }
    } else if (lo_type == CLAW_BC_WALL) {
      const double v3 = base[3 * stride], v2 = base[2 * stride];
      base[0] = negate ? -v3 : v3; base[stride] = negate ? -v2 : v2;
    } else if (lo_type == CLAW_BC_FIXED_STATE) {
      base[0] = lo_val; base[stride] = lo_val;
    }
  }
  if (up_on) {
    if (up_type == CLAW_BC_EXTRAP) {
      const double v = base[(long long)(n + 1) * stride];
      base[(long long)(n + 2) * stride] = v; base[(long long)(n + 3) * stride] = v;
    } else if (up_type == CLAW_BC_PERIODIC) {
      if (lo_on) { base[(long long)(n + 2) * stride] = base[2 * stride]; base[(long long)(n + 3) * stride] = base[3 * stride]; }
    } else if (up_type == CLAW_BC_WALL) {
      const double vn1 = base[(long long)(n + 1) * stride], vn = base[(long long)n * stride];
      base[(long long)(n + 2) * stride] = negate ? -vn1 : vn1; base[(long long)(n + 3) * stride] = negate ? -vn : vn;
    } else if (up_type == CLAW_BC_FIXED_STATE) {
      base[(long long)(n + 2) * stride] = up_val; base[(long long)(n + 3) * stride] = up_val;
    }
  }
}

// Which array a ghost-fill / halo kernel works on: a fixed pointer, or -- under a StepCtl -- the buffer that is current
// (role 0) or alternate (role 1) when the kernel RUNS; null once the step loop has ended (the kernel is a no-op).
struct ArrSel { double* q[3]; const StepCtl* ctl; int role; };
__device__ __forceinline__ double* sel_array(const ArrSel& a) {
  if (a.ctl == nullptr) return a.q[0];
  if (a.ctl->done) return nullptr;
  return ctl_buffer(a.ctl, a.q, a.role);
}

__global__ void bc_x2(ArrSel sel, int ncomp, int mx, int ny, int pitch, long long plane, int lo_type, int up_type,
                      int lo_on, int up_on, int reflect, BcState bs) {
  double* q = sel_array(sel);
  if (q == nullptr) return;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ny * ncomp) return;
  const int j = idx % ny, m = idx / ny;
  bc_fill_pair(q + m * plane + (long long)j * pitch, 1, mx, lo_type, up_type, lo_on != 0, up_on != 0, m == reflect,
               bs.lo[m & 7], bs.up[m & 7]);
}

__global__ void bc_y2(ArrSel sel, int ncomp, int nx, int my, int pitch, long long plane, int lo_type, int up_type,
                      int lo_on, int up_on, int reflect, BcState bs) {
  double* q = sel_array(sel);
  if (q == nullptr) return;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nx * ncomp) return;
  const int i = idx % nx, m = idx / nx;
  bc_fill_pair(q + m * plane + i, pitch, my, lo_type, up_type, lo_on != 0, up_on != 0, m == reflect, bs.lo[m & 7],
               bs.up[m & 7]);
}

// halo rows [j0, j0+2) x all columns x all components <-> contiguous buffer
__global__ void halo_pack2(ArrSel sel, double* buf, int ncomp, int nx, int j0, int pitch, long long plane) {
  const double* q = sel_array(sel);
  if (q == nullptr) return;
  const long long n = (long long)ncomp * 2 * nx;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(k % nx);
    const int r = (int)((k / nx) % 2);
    const int m = (int)(k / (2LL * nx));
    buf[k] = q[m * plane + (long long)(j0 + r) * pitch + i];
  }
}
__global__ void halo_unpack2(ArrSel sel, const double* buf, int ncomp, int nx, int j0, int pitch, long long plane) {
  double* q = sel_array(sel);
  if (q == nullptr) return;
  const long long n = (long long)ncomp * 2 * nx;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(k % nx);
    const int r = (int)((k / nx) % 2);
    const int m = (int)(k / (2LL * nx));
    q[m * plane + (long long)(j0 + r) * pitch + i] = buf[k];
  }
}

// 3-D ghost fill of one dimension: one thread per (component, a, b) where a, b run over the FULL
// extent (ghosts included) of the other two dimensions, so edges/corners are the composition
// x-then-y-then-z of the numpy slicing (solver.py:377-437).
__global__ void bc_dim3(ArrSel sel, int ncomp, int n_d, long long stride_d, int ext_a, long long stride_a, int ext_b,
                        long long stride_b, long long plane, int lo_type, int up_type, int lo_on, int up_on,
                        int reflect, BcState bs) {
  double* q = sel_array(sel);
  if (q == nullptr) return;
  // grid: x over the faster transverse dimension a (adjacent threads = adjacent a), y = b, z = component
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, m = blockIdx.z;
  if (a >= ext_a) return;
  bc_fill_pair(q + m * plane + a * stride_a + b * stride_b, stride_d, n_d, lo_type, up_type, lo_on != 0, up_on != 0,
               m == reflect, bs.lo[m & 7], bs.up[m & 7]);
}

// ncomp runs of `len` contiguous doubles <-> one contiguous buffer (3-D halo planes)
__global__ void gather_runs(ArrSel sel, long long off, long long src_stride, double* dst, long long len, int ncomp) {
  const double* src = sel_array(sel);
  if (src == nullptr) return;
  src += off;
  const long long n = len * ncomp;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    dst[k] = src[(k / len) * src_stride + (k % len)];
}
__global__ void scatter_runs(ArrSel sel, long long off, long long dst_stride, const double* src, long long len, int ncomp) {
  double* dst = sel_array(sel);
  if (dst == nullptr) return;
  dst += off;
  const long long n = len * ncomp;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    dst[(k / len) * dst_stride + (k % len)] = src[k];
}

// gauges: values of up to 32 cells (array coordinates incl. ghosts) -> out[c*ncomp + m]
struct CellList { int n; long long off[32]; };
__global__ void gather_cells(const double* q, int ncomp, long long plane, CellList cells, double* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cells.n * ncomp) return;
  const int c = k / ncomp, m = k % ncomp;
  out[k] = q[m * plane + cells.off[c]];
}

// FP64 issue-rate probe: 8 independent DFMA chains per thread
__global__ void dfma_probe(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}


// ---------------------------------------------------------------------------------------
// Device-resident step control: Solver.evolve_to_time / get_dt / accept_reject_step
// (pyclaw/src/pyclaw/solver.py:524-666) as two single-thread kernels around the sweeps of an attempt.  The
// arithmetic is the reference's, operation by operation, in round-to-nearest fp64 without contraction, so that dt, t
// and the accept/reject decisions carry the same bits as the host loop's.
// ---------------------------------------------------------------------------------------
__device__ void ctl_pre_step(StepCtl& c) {
  // get_dt (solver.py:531-552) with the CFL number of the previous attempt ...
  const double cfl = c.cfl_cached;
  if (c.dt_variable && c.have_dt_old) {
    if (cfl > 0.0) {
      const double x = __ddiv_rn(__dmul_rn(c.dt, c.cfl_desired), cfl);     // get_dt_new (:526-529)
      c.dt = (x < c.dt_max) ? x : c.dt_max;
      if (!(c.dtmin < c.dt)) c.dtmin = c.dt;
      if (!(c.dtmax > c.dt)) c.dtmax = c.dt;
    } else {
      c.dt = c.dt_max;
    }
  } else {
    c.have_dt_old = 1;                                                     // self.dt_old = self.dt
  }
  if (!c.take_one) {
    if (__dadd_rn(c.t, c.dt) > c.tend && c.tstart < c.tend) c.dt = __dsub_rn(c.tend, c.t);
    if (__dsub_rn(__dsub_rn(c.tend, c.t), c.dt) < __dmul_rn(1.e-14, c.t)) c.dt = __dsub_rn(c.tend, c.t);
  }
  c.told = c.t;
  c.cfl_bits = 0ull;                                                       // ... then cfl.set_global_max(0.)
}

struct CtlBeginArgs {
  double tend;
  long long take_one, max_steps, use_host, cur;
  StepCtl host;              // the host's values of t, dt, dt_max, cfl_*, status, flags (when use_host)
};

__global__ void ctl_begin_kernel(StepCtl* ctl, CtlBeginArgs a) {
  StepCtl c = *ctl;
  if (a.use_host) {
    const long long attempts = c.attempts;
    c = a.host;
    c.cur = a.cur;
    c.attempts = attempts;   // the log ring keeps running
  }
  c.tstart = c.t;
  c.tend = a.tend;
  c.take_one = a.take_one;
  c.max_steps = a.max_steps;
  c.n = 0;
  c.num_steps = 0;
  c.error = 0;
  c.done = (a.max_steps <= 0 || (!a.take_one && a.tend <= c.tstart)) ? 1 : 0;
  if (!c.done) ctl_pre_step(c);
  *ctl = c;
}

__device__ void ctl_end_body(StepCtl* ctl, double* log) {
  StepCtl c = *ctl;
  if (c.done) return;
  const double cfl = __longlong_as_double((long long)c.cfl_bits);
  c.cfl_cached = cfl;
  const bool accept = !(cfl > c.cfl_max);                                  // accept_reject_step (:524-526)
  double* e = log + 3 * (c.attempts % LOG_CAP);
  e[0] = c.dt; e[1] = cfl; e[2] = accept ? 1.0 : 0.0;
  c.attempts += 1;
  if (accept) {
    c.cur ^= 1;                                                            // the buffer the sweep wrote becomes current
    if (!(c.cflmax > cfl)) c.cflmax = cfl;
    if (c.dt_variable) c.t = __dadd_rn(c.t, c.dt);                         // :630-633
    else c.t = __dadd_rn(c.tstart, __dmul_rn((double)(c.n + 1), c.dt));
    c.num_steps += 1;
    c.numsteps += 1;
  } else if (!c.dt_variable) {                                             // :647-651 'CFL too large, giving up!'
    if (!(c.cflmax > cfl)) c.cflmax = cfl;
    c.error = 1;
    c.done = 1;
    *ctl = c;
    return;
  }                                                                        // else: t = told, q = q_backup (cur unchanged)
  c.n += 1;
  if (c.take_one || c.t >= c.tend || c.n >= c.max_steps) {                 // :654-655 and the end of the loop
    c.done = 1;
    if (!c.take_one && c.dt_variable && c.t < c.tend && c.num_steps == c.max_steps) c.error = 2;   // :660-664
  } else {
    ctl_pre_step(c);
  }
  *ctl = c;
}

__global__ void ctl_end_kernel(StepCtl* ctl, double* log) { ctl_end_body(ctl, log); }

// Both passes in ONE launch (bc_map.h: the x-then-y fill as an index map; one thread per ghost-frame cell and component;
// no thread reads a cell that another one writes).  Config 1 (200^2 acoustics), per step attempt: 18.8 us with the two
// launches of bc_x2 / bc_y2, 16.3 us with this one (profiles/r02_zb).  Folding the attempt's control step into this
// launch as well (every CTA evaluates the decision of the previous attempt on a private copy, the last CTA to finish
// publishes it: two launches per attempt) was measured SLOWER, 18.0 us -- the chain "read the control block -> fill ->
// fence -> ticket -> publish" is four dependent round trips to L2 inside one small kernel -- and was dropped.
struct BcXY { int lo[2], up[2], lo_on[2], up_on[2], reflect[2]; BcState bx, by; };
__global__ void bc_xy2(ArrSel sel, int ncomp, int mx, int my, int pitch, long long plane, BcXY b, long long ncell) {
  double* q = sel_array(sel);
  if (q == nullptr) return;
  // ncell: 4 ghost columns x all rows [+ 4 ghost rows x the interior columns when this rank fills a y face]
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int i, j;
  if (idx >= ncell * ncomp || !bc_frame_cell(idx % ncell, mx, my, i, j)) return;
  const int m = (int)(idx / ncell);
  double* qm = q + m * plane;
  bool wrote;
  const double v = bc_xy_value(qm, pitch, i, j, mx, my, b.lo, b.up, b.lo_on, b.up_on, m == b.reflect[0], m == b.reflect[1],
                               b.bx.lo[m & 7], b.bx.up[m & 7], b.by.lo[m & 7], b.by.up[m & 7], wrote);
  if (wrote) qm[(long long)j * pitch + i] = v;
}

// ---------------------------------------------------------------------------------------
// Peer-memory transport.  Mailbox of a rank (unsigned 64-bit words, device memory, mapped by every other rank):
//   [0 .. 256)    CFL slots  [parity][writer rank] = (bits of the writer's CFL number, tag)
//   [256 .. 320)  barrier slots [writer rank] = tag
//   [320], [321]  the owner's own running tags (CFL exchanges / barriers done so far; never reset)
// A writer stores the value, fences at system scope, then stores the tag; the owner polls its own memory.  Tags only grow;
// the parity double-buffers the CFL slots (a rank can be at most one exchange ahead of another: it cannot finish exchange
// n+1 without everybody's value of n+1).
// ---------------------------------------------------------------------------------------
constexpr int MAIL_CFL = 0, MAIL_BAR = 256, MAIL_TAGS = 320, MAIL_WORDS = 512, MAIL_MAXR = 64;
struct PeerMail { unsigned long long* local; unsigned long long* remote[MAIL_MAXR]; int nranks, rank; };
constexpr long long PEER_SPIN_LIMIT = 1LL << 26;       // ~ tens of seconds of polling: a dead peer must not hang the GPU

// max over ranks of `mine` (bit pattern of a non-negative double); called by 64 threads of one CTA; false on timeout
__device__ bool peer_max_bits(const PeerMail& pm, unsigned long long mine, unsigned long long* out) {
  __shared__ unsigned long long s_max;
  __shared__ int s_ok;
  const int r = threadIdx.x;
  const unsigned long long tag = pm.local[MAIL_TAGS] + 1ull;
  const unsigned par = (unsigned)(tag & 1ull);
  if (r == 0) { s_max = 0ull; s_ok = 1; }
  __syncthreads();
  if (r < pm.nranks) {
    volatile unsigned long long* dst = pm.remote[r] + MAIL_CFL + (par * MAIL_MAXR + pm.rank) * 2;
    dst[0] = mine;
    __threadfence_system();
    dst[1] = tag;
    volatile unsigned long long* src = pm.local + MAIL_CFL + (par * MAIL_MAXR + r) * 2;
    long long spins = 0;
    while (src[1] < tag) { if (++spins > PEER_SPIN_LIMIT) { s_ok = 0; break; } }
    __threadfence_system();
    atomicMax(&s_max, src[0]);
  }
  __syncthreads();
  if (r == 0) { pm.local[MAIL_TAGS] = tag; *out = s_max; }
  return s_ok != 0;
}

// control kernel of a multi-rank attempt: all-reduce (max) of the CFL slot through the mailboxes, then the same decision
// on every rank
__global__ void ctl_end_peer_kernel(StepCtl* ctl, double* log, PeerMail pm) {
  if (ctl->done) return;                                   // identical on every rank: nobody exchanges
  unsigned long long mx = 0ull;
  const bool ok = peer_max_bits(pm, ctl->cfl_bits, &mx);
  if (threadIdx.x == 0) {
    if (!ok) { ctl->error = 3; ctl->done = 1; return; }    // a peer never answered
    ctl->cfl_bits = mx;
    ctl_end_body(ctl, log);
  }
}
// the same reduction on the bare CFL slot (host-driven attempts: claw_step_launch)
__global__ void peer_allreduce_kernel(unsigned long long* cfl_bits, PeerMail pm, int* status) {
  unsigned long long mx = 0ull;
  const bool ok = peer_max_bits(pm, *cfl_bits, &mx);
  if (threadIdx.x == 0) { if (ok) *cfl_bits = mx; else *status = 1; }
}
// barrier across ranks (after a rank has refreshed its neighbours' ghost cells by hand: initial data, host uploads)
__global__ void peer_barrier_kernel(PeerMail pm, int* status) {
  const int r = threadIdx.x;
  const unsigned long long tag = pm.local[MAIL_TAGS + 1] + 1ull;
  __shared__ int s_ok;
  if (r == 0) s_ok = 1;
  __syncthreads();
  if (r < pm.nranks) {
    __threadfence_system();
    volatile unsigned long long* dst = pm.remote[r] + MAIL_BAR + pm.rank;
    *dst = tag;
    volatile unsigned long long* src = pm.local + MAIL_BAR + r;
    long long spins = 0;
    while (*src < tag) { if (++spins > PEER_SPIN_LIMIT) { s_ok = 0; break; } }
    __threadfence_system();
  }
  __syncthreads();
  if (r == 0) { pm.local[MAIL_TAGS + 1] = tag; if (!s_ok) *status = 1; }
}
// The same barrier inside a step attempt (3-D dimension-split steps under the peer-memory transport, see ctl_enqueue_end)
__global__ void peer_barrier_ctl_kernel(StepCtl* ctl, PeerMail pm) {
  if (ctl->done) return;                                   // identical on every rank: nobody waits
  const int r = threadIdx.x;
  const unsigned long long tag = pm.local[MAIL_TAGS + 1] + 1ull;
  __shared__ int s_ok;
  if (r == 0) s_ok = 1;
  __syncthreads();
  if (r < pm.nranks) {
    __threadfence_system();
    volatile unsigned long long* dst = pm.remote[r] + MAIL_BAR + pm.rank;
    *dst = tag;
    volatile unsigned long long* src = pm.local + MAIL_BAR + r;
    long long spins = 0;
    while (*src < tag) { if (++spins > PEER_SPIN_LIMIT) { s_ok = 0; break; } }
    __threadfence_system();
  }
  __syncthreads();
  if (r == 0) {
    pm.local[MAIL_TAGS + 1] = tag;
    if (!s_ok) { ctl->error = 3; ctl->done = 1; }          // a peer never answered
  }
}
// ncomp runs of `len` doubles from my array into a neighbour's (both selected by the same role under the same StepCtl:
// the current-buffer index is identical on every rank)
struct PeerChannel { long long src_off; ArrSel dst; long long dst_off, dst_stride; };
__global__ void peer_copy_runs(ArrSel src_sel, long long src_stride, PeerChannel ch0, PeerChannel ch1, long long len, int ncomp) {
  // blockIdx.z = channel (up / down neighbour): both directions in ONE launch; blockIdx.y = component
  const PeerChannel& ch = blockIdx.z == 0 ? ch0 : ch1;
  const double* src = sel_array(src_sel);
  double* dst = sel_array(ch.dst);
  if (src == nullptr || dst == nullptr) return;
  src += ch.src_off; dst += ch.dst_off;
  // 16-byte accesses when the runs allow it (row pitches are multiples of 16 doubles, so they do for every array this
  // library allocates): the stores cross NVLink
  const int m = blockIdx.y;
  const double* s = src + m * src_stride;
  double* d = dst + m * ch.dst_stride;
  const long long stride = (long long)gridDim.x * blockDim.x, k0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if ((len & 1) == 0 && ((reinterpret_cast<unsigned long long>(s) | reinterpret_cast<unsigned long long>(d)) & 15ull) == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(s);
    double2* d2 = reinterpret_cast<double2*>(d);
    for (long long k = k0; k < len / 2; k += stride) d2[k] = s2[k];
  } else {
    for (long long k = k0; k < len; k += stride) d[k] = s[k];
  }
}

// ---------------------------------------------------------------------------------------
// Device-side source terms (the `step_source` callback of classic/solver.py:126-144 for users who hand over a device
// functor instead of a Python function).  Godunov splitting: the source step follows the sweeps of an attempt, on the
// buffer they wrote, and -- as in ClawSolver.step, :134 -- only when that attempt's CFL number is below cfl_max.
//
// id 1: geometric source of the 2-D Euler equations with cylindrical symmetry, two-stage Runge-Kutta
//       (pyclaw/examples/euler_2d/shock_bubble_interaction.py:112-143, step_Euler_radial): components 0..3 =
//       (rho, rho u, rho v, E), aux plane 0 = radial coordinate of the cell centre, params = (gamma).  The arithmetic is
//       numpy's elementwise sequence operation by operation (round-to-nearest, no contraction), so the result carries
//       the same bits as the host callback's.
// ---------------------------------------------------------------------------------------
__global__ void src_euler_radial2(ArrSel sel, const double* __restrict__ aux, int mx, int my, int pitch, long long plane,
                                  double gamma, double dt_host) {
  double* q = sel_array(sel);
  if (q == nullptr) return;
  double dt = dt_host;
  if (sel.ctl != nullptr) {
    const double cfl = __longlong_as_double((long long)sel.ctl->cfl_bits);
    if (cfl >= sel.ctl->cfl_max) return;                                   // classic/solver.py:134
    dt = sel.ctl->dt;
  }
  const long long n = (long long)mx * my;
  const double g1 = __dsub_rn(gamma, 1.0);
  const double dt2 = __ddiv_rn(dt, 2.0);
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % mx) + 2, j = (int)(c / mx) + 2;
    double* p = q + (long long)j * pitch + i;
    const double rad = aux[(long long)j * pitch + i];
    const double q0 = p[0], q1 = p[plane], q2 = p[2 * plane], q3 = p[3 * plane];
    // first stage
    double rho = q0;
    double u = __ddiv_rn(q1, rho), v = __ddiv_rn(q2, rho);
    double press = __dmul_rn(g1, __dsub_rn(q3, __dmul_rn(__dmul_rn(0.5, rho), __dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v)))));
    const double a2 = __ddiv_rn(dt2, rad);
    const double s0 = __dsub_rn(q0, __dmul_rn(a2, q2));
    const double s1 = __dsub_rn(q1, __dmul_rn(__dmul_rn(__dmul_rn(a2, rho), u), v));
    const double s2 = __dsub_rn(q2, __dmul_rn(__dmul_rn(__dmul_rn(a2, rho), v), v));
    const double s3 = __dsub_rn(q3, __dmul_rn(__dmul_rn(a2, v), __dadd_rn(q3, press)));
    // second stage
    rho = s0;
    u = __ddiv_rn(s1, rho); v = __ddiv_rn(s2, rho);
    press = __dmul_rn(g1, __dsub_rn(s3, __dmul_rn(__dmul_rn(0.5, rho), __dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v)))));
    const double a = __ddiv_rn(dt, rad);
    p[0] = __dsub_rn(q0, __dmul_rn(a, s2));
    p[plane] = __dsub_rn(q1, __dmul_rn(__dmul_rn(__dmul_rn(a, rho), u), v));
    p[2 * plane] = __dsub_rn(q2, __dmul_rn(__dmul_rn(__dmul_rn(a, rho), v), v));
    p[3 * plane] = __dsub_rn(q3, __dmul_rn(__dmul_rn(a, v), __dadd_rn(s3, press)));
  }
}

int check_handle(const claw_solver* h) {
  if (!h) { g_last_error = "null solver handle"; return 1; }
  return 0;
}

int ensure_buffers(claw_solver* h) {
  if (h->q[0]) return 0;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMalloc(&h->q[0], h->q_elems * sizeof(double)));
  CK(h, cudaMalloc(&h->q[1], h->q_elems * sizeof(double)));
  if (h->aux_elems) CK(h, cudaMalloc(&h->aux, h->aux_elems * sizeof(double)));
  CK(h, cudaMalloc(&h->work, h->work_elems * sizeof(double)));
  CK(h, cudaMemsetAsync(h->q[0], 0, h->q_elems * sizeof(double), h->stream));
  CK(h, cudaMemsetAsync(h->q[1], 0, h->q_elems * sizeof(double), h->stream));
  h->owns_buffers = true;
  return 0;
}

// Attempts queued under device-side step control: wait for them and bring the host's `cur` up to date.
int ctl_settle(claw_solver* h) {
  if (!h->ctl_pending) return 0;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemcpyAsync(h->h_ctl, h->work, CTL_WORDS * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->cur = (int)reinterpret_cast<const StepCtl*>(h->h_ctl)->cur;
  h->ctl_pending = false;
  return 0;
}
// host-driven entry points (claw_step_launch, claw_accept, ...) take the handle out of device-side step control
int ctl_leave(claw_solver* h) {
  if (int rc = ctl_settle(h)) return rc;
  h->ctl_mode = false;
  return 0;
}

int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  if (g > 148LL * 32) g = 148LL * 32;
  if (g < 1) g = 1;
  return (int)g;
}

// copy a region of the ghosted host array (AoS) to / from the device SoA array
int put_region(claw_solver* h, double* dsoa, int ncomp, const double* host, bool host_has_ghosts) {
  const Dev2& g = h->g;
  if (int rc = ensure_buffers(h)) return rc;
  double* stage = h->work + WORK_HDR;
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int d = 0; d < g.ndim; ++d) {
    lo[d] = host_has_ghosts ? 0 : g.ng;
    hi[d] = host_has_ghosts ? g.nt[d] : g.ng + g.n[d];
  }
  const long long ncell = (long long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  CK(h, cudaMemcpyAsync(stage, host, (size_t)ncomp * ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  aos_to_soa<<<grid_for(ncell, 256), 256, 0, h->stream>>>(stage, dsoa, ncomp, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2],
                                                          g.pitch, g.kstride, g.plane);
  ++h->launches;
  CK(h, cudaGetLastError());
  return 0;
}

// region [lo, hi) of the device array -> the same cells of a host array whose ghosted/unghosted extents are
// host_n[] and whose first stored cell is host_lo[] (array coordinates)
int get_region(claw_solver* h, const double* dsoa, int ncomp, double* host, const int* lo, const int* hi,
               const int* host_lo, const int* host_n) {
  const Dev2& g = h->g;
  double* stage = h->work + WORK_HDR;
  const int w = hi[0] - lo[0], hgt = hi[1] - lo[1], dep = hi[2] - lo[2];
  const long long ncell = (long long)w * hgt * dep;
  soa_to_aos<<<grid_for(ncell, 256), 256, 0, h->stream>>>(dsoa, stage, ncomp, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2],
                                                          g.pitch, g.kstride, g.plane);
  ++h->launches;
  CK(h, cudaGetLastError());
  const size_t row_bytes = (size_t)ncomp * w * sizeof(double);
  const size_t host_row = (size_t)ncomp * host_n[0] * sizeof(double);
  char* dst = (char*)host + (((size_t)(lo[2] - host_lo[2]) * host_n[1] + (lo[1] - host_lo[1])) * host_n[0] +
                             (lo[0] - host_lo[0])) * ncomp * sizeof(double);
  if (dep == 1) {
    CK(h, cudaMemcpy2DAsync(dst, host_row, stage, row_bytes, row_bytes, (size_t)hgt, cudaMemcpyDeviceToHost, h->stream));
  } else {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(stage, row_bytes, row_bytes, (size_t)hgt);
    p.dstPtr = make_cudaPitchedPtr(dst, host_row, host_row, (size_t)host_n[1]);
    p.extent = make_cudaExtent(row_bytes, (size_t)hgt, (size_t)dep);
    p.kind = cudaMemcpyDeviceToHost;
    CK(h, cudaMemcpy3DAsync(&p, h->stream));
  }
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// interior / whole-array downloads
int get_interior(claw_solver* h, const double* dsoa, int ncomp, double* host, bool host_has_ghosts) {
  const Dev2& g = h->g;
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, hlo[3] = {0, 0, 0}, hn[3] = {1, 1, 1};
  for (int d = 0; d < g.ndim; ++d) {
    lo[d] = g.ng; hi[d] = g.ng + g.n[d];
    hlo[d] = host_has_ghosts ? 0 : g.ng;
    hn[d] = host_has_ghosts ? g.nt[d] : g.n[d];
  }
  return get_region(h, dsoa, ncomp, host, lo, hi, hlo, hn);
}
int get_all(claw_solver* h, const double* dsoa, int ncomp, double* host) {
  const Dev2& g = h->g;
  int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, hn[3] = {1, 1, 1};
  for (int d = 0; d < g.ndim; ++d) { hi[d] = g.nt[d]; hn[d] = g.nt[d]; }
  return get_region(h, dsoa, ncomp, host, lo, hi, lo, hn);
}

ArrSel sel_fixed(double* p) { ArrSel a{}; a.q[0] = p; a.ctl = nullptr; a.role = 0; return a; }
// the q buffer that is current (role 0) / alternate (role 1): by the host's `cur`, or by the device's under ctl_mode
ArrSel sel_q(claw_solver* h, int role) {
  ArrSel a{};
  if (!h->ctl_mode) { a.q[0] = h->q[role == 0 ? h->cur : (h->cur ^ 1)]; return a; }
  a.q[0] = h->q[0]; a.q[1] = h->q[1]; a.q[2] = nullptr;
  a.ctl = reinterpret_cast<const StepCtl*>(h->work);
  a.role = role;
  return a;
}

BcState bc_state_of(const claw_solver* h, int d) {
  BcState b{};
  for (int m = 0; m < 8; ++m) { b.lo[m] = h->bc_state[d][0][m]; b.up[m] = h->bc_state[d][1][m]; }
  return b;
}

int fill_ghosts_array(claw_solver* h, ArrSel arr, int ncomp, const int* lo, const int* up, bool is_q) {
  const Dev2& g = h->g;
  for (int d = 0; d < g.ndim; ++d) {
    // BC.custom (0) on an AUX face is allowed: the host filled those ghost cells once (claw_put_auxbc) and the
    // kernels below leave type 0 untouched; on q it needs a host round trip every step
    const int tmin = is_q ? 1 : 0, tmax = is_q ? CLAW_BC_FIXED_STATE : 3;
    if ((h->on_lower[d] && (lo[d] < tmin || lo[d] > tmax)) || (h->on_upper[d] && (up[d] < tmin || up[d] > tmax)))
      return fail(h, 21, "boundary condition %d/%d in dimension %d cannot run on the device "
                         "(BC.custom needs a host round trip)", lo[d], up[d], d);
  }
  if (g.ndim == 3) {
    const long long st[3] = {1, (long long)g.pitch, g.kstride};
    for (int d = 0; d < 3; ++d) {
      // a = the transverse dimension with the smaller stride (adjacent threads touch adjacent cells where the face
      // allows it: y- and z-faces are contiguous in x; an x-face thread owns the 2 x 2 ghost cells of its pencil)
      const int a = (d == 0) ? 1 : 0, b = (d == 2) ? 1 : 2;
      const int refl = is_q ? h->reflect_index[d] : -1;
      const dim3 grid((unsigned)((g.nt[a] + 127) / 128), (unsigned)g.nt[b], (unsigned)ncomp);
      if (g.nt[b] > 65535) return fail(h, 23, "3-D ghost fill: more than 65535 cells across");
      bc_dim3<<<grid, 128, 0, h->stream>>>(arr, ncomp, g.n[d], st[d], g.nt[a], st[a], g.nt[b],
                                                                  st[b], g.plane, lo[d], up[d], h->on_lower[d],
                                                                  h->on_upper[d], refl, bc_state_of(h, d));
      ++h->launches;
    }
    CK(h, cudaGetLastError());
    return 0;
  }
  // CLAW_BC_TWO_PASS=1: the ghost fill as two launches, x strips then y strips (A/B and a second implementation for
  // the parity tests); default: one launch through the composed index map
  static const bool two_pass = getenv("CLAW_BC_TWO_PASS") && atoi(getenv("CLAW_BC_TWO_PASS")) != 0;
  // (a face reads interior cells only when the grid has at least two cells across: below that, keep the two passes)
  if (!two_pass && g.n[0] >= 2 && g.n[1] >= 2) {
    BcXY b{};
    for (int d = 0; d < 2; ++d) {
      b.lo[d] = lo[d]; b.up[d] = up[d]; b.lo_on[d] = h->on_lower[d]; b.up_on[d] = h->on_upper[d];
      b.reflect[d] = is_q ? h->reflect_index[d] : -1;     // only consulted by the BC.wall branches (solver.py:464-473)
    }
    b.bx = bc_state_of(h, 0); b.by = bc_state_of(h, 1);
    // an inner rank of a slab decomposition fills no y face: its launch covers the ghost columns only (the x-fill of the
    // received ghost rows included)
    const bool y_active = (b.lo_on[1] && b.lo[1] != CLAW_BC_CUSTOM) || (b.up_on[1] && b.up[1] != CLAW_BC_CUSTOM);
    const long long ncell = 4LL * g.nt[1] + (y_active ? 4LL * g.n[0] : 0LL);
    const long long n = ncell * ncomp;
    bc_xy2<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(arr, ncomp, g.n[0], g.n[1], g.pitch, g.plane, b, ncell);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
  }
  {
    // the negation flag is only consulted by the BC.wall branches (solver.py:464-473)
    const int refl = is_q ? h->reflect_index[0] : -1;
    const int n = g.nt[1] * ncomp;
    bc_x2<<<(n + 127) / 128, 128, 0, h->stream>>>(arr, ncomp, g.n[0], g.nt[1], g.pitch, g.plane, lo[0], up[0],
                                                  h->on_lower[0], h->on_upper[0], refl, bc_state_of(h, 0));
    ++h->launches;
  }
  {
    const int refl = is_q ? h->reflect_index[1] : -1;
    const int n = g.nt[0] * ncomp;
    bc_y2<<<(n + 127) / 128, 128, 0, h->stream>>>(arr, ncomp, g.nt[0], g.n[1], g.pitch, g.plane, lo[1], up[1],
                                                  h->on_lower[1], h->on_upper[1], refl, bc_state_of(h, 1));
    ++h->launches;
  }
  CK(h, cudaGetLastError());
  return 0;
}

// Riemann-solver plug-ins (rp_plugin.cu): shared objects registered by claw_rp_plugin_load
struct RpPlugin {
  std::string path;
  void* dl = nullptr;
  int (*launch2)(int, int, void*, int, int, void*, void*) = nullptr;      // 2-D solvers
  int (*launch3)(int, int, void*, int, void*, void*) = nullptr;           // 3-D solvers
  int ndim = 2, meqn = 0, mwaves = 0, maux = 0, no_transverse = 0;
};
std::mutex g_plugin_mutex;
// rp_id = CLAW_RP_PLUGIN_BASE + index; entries are never removed and never move (pointers handed out by plugin_of stay
// valid while other threads register more plug-ins)
constexpr int MAX_PLUGINS = 256;
RpPlugin g_plugins[MAX_PLUGINS];
int g_num_plugins = 0;
const RpPlugin* plugin_of(int rp_id) {
  std::lock_guard<std::mutex> lock(g_plugin_mutex);
  const int k = rp_id - CLAW_RP_PLUGIN_BASE;
  return (k >= 0 && k < g_num_plugins) ? &g_plugins[k] : nullptr;
}

int launch_step2(claw_solver* h, double dx, double dy, double dt, int kind, int do_x, int do_y,
                 const double* src, double* dst, int j0 = 0, int j1 = 0, cudaStream_t stream_override = nullptr,
                 bool under_ctl = false) {
  Sweep2Params P{};
  if (under_ctl) {                                     // src / dst / dt are decided on the device (StepCtl)
    P.ctl = reinterpret_cast<const StepCtl*>(h->work);
    P.qbuf[0] = h->q[0]; P.qbuf[1] = h->q[1]; P.qbuf[2] = nullptr;
    P.delta[0] = dx; P.delta[1] = dy;
  }
  P.j0 = j0; P.j1 = j1;                                // 0,0 = whole grid
  cudaStream_t st = stream_override ? stream_override : h->stream;
  P.qin = src;
  P.qout = dst;
  P.cfl_bits = reinterpret_cast<unsigned long long*>(h->work);
  P.mx = h->g.n[0]; P.my = h->g.n[1];
  P.pitch = h->g.pitch; P.plane = h->g.plane;
  P.dtdx = dt / dx;                                   // step2.f90:77-78
  P.dtdy = dt / dy;
  P.order = h->method[1];
  P.mtrans = h->method[2];
  P.limit = 0;
  for (int mw = 0; mw < h->mwaves; ++mw) { P.mthlim[mw] = h->mthlim[mw]; if (h->mthlim[mw] > 0) P.limit = 1; }
  P.do_x = do_x; P.do_y = do_y;
  P.aux = h->aux; P.maux = h->maux; P.capa = h->method[5]; P.fwave = h->use_fwave;
  for (int k = 0; k < 4; ++k) P.rp.p[k] = h->cparam[k];
  // tile ordering: full-grid launches of the device-resident path only (the chunks of the pipelined host entry points
  // are different grids on other streams)
  const bool full_grid = h->hot && j0 == 0 && j1 == 0 && stream_override == nullptr;
  if (full_grid) {
    P.hot = h->hot; P.hot_cap = h->hot_cap; P.hot_tiles = h->hot_tiles;
    P.hot_epoch = h->hot_epoch++;
    P.dense_pref = h->dense_pref;
  }
  cudaError_t e;
  if (h->rp_id >= CLAW_RP_PLUGIN_BASE) {               // user-supplied Riemann solver: its own instantiation of the kernels
    const RpPlugin* p = plugin_of(h->rp_id);
    if (!p || !p->launch2) return fail(h, 30, "rp_id %d is not a registered 2-D Riemann-solver plug-in", h->rp_id);
    e = (cudaError_t)p->launch2(h->mode == CLAW_MODE_FAST ? 1 : 0, kind, &P, h->seg_rows, h->num_sms, (void*)st,
                                &h->last_launch);
  } else {
    e = (h->mode == CLAW_MODE_FAST)
            ? fast::launch_sweep2(h->rp_id, kind, P, h->seg_rows, h->num_sms, st, &h->last_launch)
            : strict::launch_sweep2(h->rp_id, kind, P, h->seg_rows, h->num_sms, st, &h->last_launch);
  }
  ++h->launches;
  if (e != cudaSuccess) return fail(h, 100 + (int)e, "sweep kernel launch failed: %s", cudaGetErrorString(e));
  if (full_grid) {                                     // what dense_update needs to read the heavy-tile counts
    h->last_hot_rows = P.hot_rows;
    h->last_ntiles = (long long)h->last_launch.grid_x * (h->last_launch.grid_y - P.hot_rows);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// Dense-data switch.  On data without uniform regions the kernels with the uniform-region fast path are 14-20 % slower
// than the same kernels without it (shorter march segments, the per-step classification, the deeper
// ring).  The fast-path kernels count their heavy tiles for the tile ordering anyway (HotTiles, sweep2_ws.cuh), so the
// host copies the three counters after a batch of attempts -- asynchronously, consumed at the next batch without a
// wait -- and
//   * switches to the plain kernels when more than 60 % of the tiles were heavy,
//   * probes with the fast-path kernels every 64th batch while in that state, and switches back below 40 %.
// Only launches long enough for the tile ordering (several waves of CTAs) report counts; smaller grids never switch.
// CLAW_DENSE_SWITCH=0 turns it off.  A switch re-captures the attempt graph (cfg_epoch).
// ---------------------------------------------------------------------------------------
bool dense_switch_enabled() {
  static const bool on = !(getenv("CLAW_DENSE_SWITCH") && atoi(getenv("CLAW_DENSE_SWITCH")) == 0);
  return on;
}
// fraction of heavy tiles in the counters copied last, or -1 when they say nothing
double dense_fraction(const claw_solver* h) {
  if (!h->h_hot || h->last_ntiles <= 0 || h->last_hot_rows <= 0) return -1.0;
  unsigned mx = h->h_hot[0];
  if (h->h_hot[1] > mx) mx = h->h_hot[1];
  if (h->h_hot[2] > mx) mx = h->h_hot[2];
  return (double)mx / (double)h->last_ntiles;
}
void dense_decide(claw_solver* h, double frac) {
  if (frac < 0.0) return;
  const int before = h->dense_pref;
  if (h->dense_probe) {                                // the batch that just ran was a probe out of the dense state
    h->dense_probe = false;
    h->dense_pref = frac > 0.4 ? 1 : 0;
    h->dense_batches = 0;
    if (h->dense_pref != 0) ++h->cfg_epoch;            // (the probe ran the fast-path kernels: back to the plain ones)
    return;
  }
  if (!h->dense_pref && frac > 0.6) h->dense_pref = 1;
  if (h->dense_pref != before) { ++h->cfg_epoch; h->dense_batches = 0; }
}
// called before a batch of attempts is queued
void dense_update(claw_solver* h) {
  if (!dense_switch_enabled() || !h->hot || h->g.ndim != 2) return;
  // (the verdict of a probe is waited for -- once per probe -- so that a host that queues far ahead of the GPU does not
  // keep running the probe's kernels on dense data for as many batches as it is ahead)
  const bool wait = h->hot_copy_pending && h->dense_probe && h->hot_copy_from_quiet;
  if (h->hot_copy_pending && (wait ? cudaEventSynchronize(h->ev_hot) : cudaEventQuery(h->ev_hot)) == cudaSuccess) {
    h->hot_copy_pending = false;
    if (h->hot_copy_from_quiet) dense_decide(h, dense_fraction(h));
  }
  static const long long probe_every = (getenv("CLAW_DENSE_PROBE") && atoi(getenv("CLAW_DENSE_PROBE")) > 0)
                                           ? atoi(getenv("CLAW_DENSE_PROBE")) : 64;      // (tests shorten it)
  if (h->dense_pref && !h->dense_probe && ++h->dense_batches >= probe_every) {       // look again: has the data become sparse?
    h->dense_probe = true;
    h->dense_pref = 0;
    ++h->cfg_epoch;
    // the plain kernels never cleared the counters: the probe must not add to what an earlier launch left behind
    cudaMemsetAsync(h->hot, 0, 3 * sizeof(unsigned), h->stream);
  }
}
// called after a batch of attempts has been queued
int dense_collect(claw_solver* h) {
  if (!dense_switch_enabled() || !h->hot || h->g.ndim != 2 || h->hot_copy_pending) return 0;
  if (!h->h_hot) {
    CK(h, cudaHostAlloc(&h->h_hot, 8 * sizeof(unsigned), cudaHostAllocDefault));
    CK(h, cudaEventCreateWithFlags(&h->ev_hot, cudaEventDisableTiming));
  }
  CK(h, cudaMemcpyAsync(h->h_hot, h->hot, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaEventRecord(h->ev_hot, h->stream));
  h->hot_copy_pending = true;
  h->hot_copy_from_quiet = (h->dense_pref == 0);
  return 0;
}

// one directional pass of the 3-D step (dir 0/1/2 = x/y/z)
int launch_pass3(claw_solver* h, const double* delta, double dt, int dir, int first_pass, const double* src,
                 double* dst, int src_role = -1, int dst_role = -1) {
  Sweep3Params P{};
  if (src_role >= 0) {                                 // buffers and dt are decided on the device (StepCtl)
    P.ctl = reinterpret_cast<const StepCtl*>(h->work);
    P.qbuf[0] = h->q[0]; P.qbuf[1] = h->q[1]; P.qbuf[2] = h->work + WORK_HDR;
    P.src_role = src_role; P.dst_role = dst_role;
    for (int d = 0; d < 3; ++d) P.delta[d] = delta[d];
  }
  P.qin = src;
  P.qout = dst;
  P.aux = h->aux;
  P.capa = h->method[5];                                                             // step3.f90:103
  P.cfl_bits = reinterpret_cast<unsigned long long*>(h->work);
  for (int d = 0; d < 3; ++d) { P.n[d] = h->g.n[d]; P.dtd[d] = dt / delta[d]; }      // step3.f90:106-108
  P.pitch = h->g.pitch; P.kstride = h->g.kstride; P.plane = h->g.plane;
  P.order = h->method[1];
  if (h->method[2] < 0) { P.m3 = -1; P.m4 = 0; }                                     // flux3.f90:186-194
  else { P.m3 = h->method[2] / 10; P.m4 = h->method[2] - 10 * P.m3; }
  P.limit = 0;
  for (int mw = 0; mw < h->mwaves; ++mw) { P.mthlim[mw] = h->mthlim[mw]; if (h->mthlim[mw] > 0) P.limit = 1; }
  P.first_pass = first_pass;
  // dimension-split sweeps cover one ghost layer (step3ds.f90:134-135); at a rank-to-rank face of the slab
  // dimension both received ghost planes are swept so that the sharded run equals the SERIAL reference
  // (a PetClaw run differs there)
  P.zlo = (h->method[2] < 0 && !h->on_lower[2]) ? 0 : 1;
  P.zhi = h->g.n[2] + ((h->method[2] < 0 && !h->on_upper[2]) ? 3 : 2);
  for (int k = 0; k < 4; ++k) P.rp.p[k] = h->cparam[k];
  cudaError_t e;
  if (h->rp_id >= CLAW_RP_PLUGIN_BASE) {               // user-supplied Riemann solver
    const RpPlugin* p = plugin_of(h->rp_id);
    if (!p || !p->launch3) return fail(h, 30, "rp_id %d is not a registered 3-D Riemann-solver plug-in", h->rp_id);
    e = (cudaError_t)p->launch3(h->mode == CLAW_MODE_FAST ? 1 : 0, dir, &P, h->num_sms, (void*)h->stream, &h->last_launch);
  } else {
    e = (h->mode == CLAW_MODE_FAST)
            ? fast::launch_sweep3(h->rp_id, dir, P, h->num_sms, h->stream, &h->last_launch)
            : strict::launch_sweep3(h->rp_id, dir, P, h->num_sms, h->stream, &h->last_launch);
  }
  ++h->launches;
  if (e != cudaSuccess) return fail(h, 100 + (int)e, "3-D sweep kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// whole 3-D step: src (ghosts filled) -> dst.  Unsplit (step3.f90): x-sweep writes dst = src + terms, the y- and
// z-sweeps add into dst.  Split (3 x step3ds.f90, classic/solver.py:680-690): src -> dst -> tmp -> dst.
int launch_step3(claw_solver* h, const double* delta, double dt, const double* src, double* dst, bool under_ctl) {
  if (under_ctl) {                                     // roles: 0 = current, 1 = alternate, 2 = third buffer
    if (h->method[2] < 0) {
      if (int rc = launch_pass3(h, delta, dt, 0, 1, nullptr, nullptr, 0, 1)) return rc;
      if (int rc = launch_pass3(h, delta, dt, 1, 1, nullptr, nullptr, 1, 2)) return rc;
      return launch_pass3(h, delta, dt, 2, 1, nullptr, nullptr, 2, 1);
    }
    if (int rc = launch_pass3(h, delta, dt, 0, 1, nullptr, nullptr, 0, 1)) return rc;
    if (int rc = launch_pass3(h, delta, dt, 1, 0, nullptr, nullptr, 0, 1)) return rc;
    return launch_pass3(h, delta, dt, 2, 0, nullptr, nullptr, 0, 1);
  }
  if (h->method[2] < 0) {
    double* tmp = h->work + WORK_HDR;
    if (int rc = launch_pass3(h, delta, dt, 0, 1, src, dst)) return rc;
    if (int rc = launch_pass3(h, delta, dt, 1, 1, dst, tmp)) return rc;
    return launch_pass3(h, delta, dt, 2, 1, tmp, dst);
  }
  if (int rc = launch_pass3(h, delta, dt, 0, 1, src, dst)) return rc;
  if (int rc = launch_pass3(h, delta, dt, 1, 0, src, dst)) return rc;
  return launch_pass3(h, delta, dt, 2, 0, src, dst);
}

int launch_step(claw_solver* h, const double* delta, double dt, const double* src, double* dst, bool under_ctl = false) {
  if (h->g.ndim == 3) return launch_step3(h, delta, dt, src, dst, under_ctl);
  const int kind = (h->method[2] < 0) ? 1 : 0;         // method(3) = -1: dimensional splitting (classic/solver.py:186)
  return launch_step2(h, delta[0], delta[1], dt, kind, 1, 1, src, dst, 0, 0, nullptr, under_ctl);
}

// aux planes read by the 2-D Riemann solvers (rp2.cuh: none; rp2_aux.cuh: RP::MAUX); -1 = unknown id
int rp2_aux_planes(int rp_id) {
  if (rp_id >= CLAW_RP_PLUGIN_BASE) { const RpPlugin* p = plugin_of(rp_id); return (p && p->ndim == 2) ? p->maux : -1; }
  switch (rp_id) {
    case 1: case 2: case 3: case 4: case 9: case 11: case 12: case 16: case 18: return 0;
    case 7: return 2;      // vc_acoustics_2D: density, sound speed
    case 8: return 2;      // vc_advection_2D: edge velocities
    case 10: return 4;     // psystem_2D: density, modulus, stress-strain relation id, strain
    case 13: return 9;     // acoustics_mapped_2D: edge normals + length ratios, capacity, impedance, sound speed
    case 14: return 1;     // shallow_bathymetry_fwave_2D: bathymetry
    case 19: return 5;     // vc_elasticity_2D: density, lambda, mu, cp, cs
    case 20: return 2;     // vc_advection_vector_2D: edge velocities
  }
  return -1;
}

int validate_config(claw_solver* h) {
  if (h->g.ndim == 2) {
    const int need_aux = rp2_aux_planes(h->rp_id);          // aux planes the solver reads (-1: not a 2-D solver)
    if (need_aux < 0) return fail(h, 30, "rp_id %d is not a 2-D Riemann solver of this build", h->rp_id);
    if (h->maux < need_aux)
      return fail(h, 31, "Riemann solver %d reads %d aux components, the state has num_aux=%d", h->rp_id, need_aux, h->maux);
    if (h->method[5] < 0 || h->method[5] > h->maux)
      return fail(h, 31, "index_capa=%d is outside the aux array (num_aux=%d)", h->method[5] - 1, h->maux);
    if (need_aux == 0 && h->method[5] != 0)
      return fail(h, 31, "capacity functions are only implemented for the aux-reading 2-D solvers (rp_id %d reads none)", h->rp_id);
    if (need_aux == 0 && h->use_fwave)
      return fail(h, 32, "rp_id %d is not an f-wave solver", h->rp_id);
    if (h->rp_id >= CLAW_RP_PLUGIN_BASE) {
      const RpPlugin* p = plugin_of(h->rp_id);
      if (p->meqn != h->meqn || p->mwaves != h->mwaves)
        return fail(h, 37, "Riemann-solver plug-in %s solves %d equations with %d waves, the solver was created with "
                           "num_eqn=%d, num_waves=%d", p->path.c_str(), p->meqn, p->mwaves, h->meqn, h->mwaves);
      if (p->no_transverse && h->method[2] > 0)
        return fail(h, 36, "Riemann-solver plug-in %s has no transverse solver: transverse_waves must be 0 or the step "
                           "dimensionally split", p->path.c_str());
    }
    if ((h->rp_id == 11 || h->rp_id == 12 || h->rp_id == 14 || h->rp_id == 18 || h->rp_id == 20) && h->method[2] > 0)
      return fail(h, 36, "rp_id %d has no usable transverse Riemann solver (rpt2_dummy.f90; rpt2_shallow_bathymetry_fwave "
                         "reads an unassigned value): transverse_waves must be 0 or the step dimensionally split", h->rp_id);
    if (h->method[1] != 1 && h->method[1] != 2) return fail(h, 33, "order must be 1 or 2 (got %d)", h->method[1]);
    if (h->method[2] < -1 || h->method[2] > 2) return fail(h, 34, "transverse_waves must be 0, 1 or 2 in 2-D (got %d)", h->method[2]);
    for (int mw = 0; mw < h->mwaves; ++mw)
      if (h->mthlim[mw] < 0 || h->mthlim[mw] > 5)
        return fail(h, 35, "limiter id %d of wave family %d is not defined by philim.f90 (0..5)", h->mthlim[mw], mw);
    return 0;
  }
  const RpPlugin* plug3 = (h->rp_id >= CLAW_RP_PLUGIN_BASE) ? plugin_of(h->rp_id) : nullptr;
  if (plug3 && plug3->ndim != 3) plug3 = nullptr;
  if (!plug3 && h->rp_id != 5 && h->rp_id != 6 && h->rp_id != 15 && h->rp_id != 17)
    return fail(h, 30, "rp_id %d is not a 3-D Riemann solver of this build", h->rp_id);
  if (plug3 && (plug3->meqn != h->meqn || plug3->mwaves != h->mwaves))
    return fail(h, 37, "Riemann-solver plug-in %s solves %d equations with %d waves, the solver was created with num_eqn=%d, "
                       "num_waves=%d", plug3->path.c_str(), plug3->meqn, plug3->mwaves, h->meqn, h->mwaves);
  {
    const int need_aux = plug3 ? plug3->maux : (h->rp_id == 5) ? 2 : (h->rp_id == 17 ? 3 : 0);      // RP::MAUX of rp3.cuh
    if (h->maux < need_aux)
      return fail(h, 31, "Riemann solver %d reads %d aux components, the state has num_aux=%d", h->rp_id, need_aux, h->maux);
  }
  if (h->method[5] < 0 || h->method[5] > h->maux)
    return fail(h, 31, "index_capa=%d is outside the aux array (num_aux=%d)", h->method[5] - 1, h->maux);

  if (h->use_fwave) return fail(h, 32, "f-wave solvers are not supported by the CUDA kernels yet");
  if (h->method[1] != 1 && h->method[1] != 2) return fail(h, 33, "order must be 1 or 2 (got %d)", h->method[1]);
  {
    const int mt = h->method[2];
    if (!(mt == -1 || mt == 0 || mt == 10 || mt == 11 || mt == 20 || mt == 21 || mt == 22))
      return fail(h, 34, "transverse_waves must be one of 0, 10, 11, 20, 21, 22 in 3-D (got %d)", mt);
    if (mt >= 20 && h->method[1] != 2) return fail(h, 34, "transverse_waves=%d needs order=2 (flux3.f90:66-74)", mt);
  }
  for (int mw = 0; mw < h->mwaves; ++mw)
    if (h->mthlim[mw] < 0 || h->mthlim[mw] > 5)
      return fail(h, 35, "limiter id %d of wave family %d is not defined by philim.f90 (0..5)", h->mthlim[mw], mw);
  return 0;
}

}  // namespace

// =========================================================================================
extern "C" {

const char* claw_version(void) { return "clawcuda 0.1 (sm_100a)"; }

const char* claw_last_error(const claw_solver* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int claw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int claw_rp_plugin_load(const char* path, int* rp_id) {
  if (!path || !rp_id) return fail(nullptr, 1, "claw_rp_plugin_load: NULL argument");
  std::lock_guard<std::mutex> lock(g_plugin_mutex);
  for (int k = 0; k < g_num_plugins; ++k)
    if (g_plugins[k].path == path) { *rp_id = CLAW_RP_PLUGIN_BASE + k; return 0; }
  if (g_num_plugins == MAX_PLUGINS) return fail(nullptr, 64, "claw_rp_plugin_load: %d plug-ins are registered already", MAX_PLUGINS);
  void* dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) return fail(nullptr, 60, "claw_rp_plugin_load: %s", dlerror());
  auto abi = reinterpret_cast<int (*)(void)>(dlsym(dl, "claw_rp_plugin_abi"));
  auto info = reinterpret_cast<void (*)(int*)>(dlsym(dl, "claw_rp_plugin_info"));
  auto launch2 = reinterpret_cast<int (*)(int, int, void*, int, int, void*, void*)>(dlsym(dl, "claw_rp_plugin_launch2"));
  auto launch3 = reinterpret_cast<int (*)(int, int, void*, int, void*, void*)>(dlsym(dl, "claw_rp_plugin_launch3"));
  if (!abi || !info || (!launch2 && !launch3)) {
    dlclose(dl);
    return fail(nullptr, 61, "claw_rp_plugin_load: %s does not export claw_rp_plugin_abi / _info / _launch2 (or _launch3) "
                             "(compile clawpack_b200/csrc/rp_plugin.cu against your functor)", path);
  }
  int v[6] = {0, 0, 0, 0, 0, 0};
  info(v);
  const int want_bytes = (v[0] == 3) ? (int)sizeof(Sweep3Params) : (int)sizeof(Sweep2Params);
  if (abi() != CLAW_PLUGIN_ABI || v[5] != want_bytes) {
    const int got = abi();
    dlclose(dl);
    return fail(nullptr, 62, "claw_rp_plugin_load: %s was compiled against layout version %d (launch parameters %d bytes), "
                             "this library has version %d (%d bytes): recompile the plug-in", path, got, v[5], CLAW_PLUGIN_ABI,
                want_bytes);
  }
  if ((v[0] != 2 && v[0] != 3) || (v[0] == 2 && !launch2) || (v[0] == 3 && !launch3) || v[1] < 1 || v[1] > 8 || v[2] < 1 ||
      v[2] > 8 || v[3] < 0) {
    dlclose(dl);
    return fail(nullptr, 63, "claw_rp_plugin_load: %s reports num_dim=%d, meqn=%d, mwaves=%d, maux=%d (2-D and 3-D solvers "
                             "with 1..8 equations and waves are supported)", path, v[0], v[1], v[2], v[3]);
  }
  RpPlugin p;
  p.path = path; p.dl = dl; p.launch2 = launch2; p.launch3 = launch3;
  p.ndim = v[0]; p.meqn = v[1]; p.mwaves = v[2]; p.maux = v[3]; p.no_transverse = v[4];
  g_plugins[g_num_plugins] = p;
  *rp_id = CLAW_RP_PLUGIN_BASE + g_num_plugins;
  ++g_num_plugins;
  return 0;
}

int claw_rp_plugin_info(int rp_id, int* info5) {
  if (!info5) return fail(nullptr, 1, "claw_rp_plugin_info: NULL argument");
  const RpPlugin* p = plugin_of(rp_id);
  if (!p) return fail(nullptr, 30, "rp_id %d is not a registered Riemann-solver plug-in", rp_id);
  info5[0] = p->ndim; info5[1] = p->meqn; info5[2] = p->mwaves; info5[3] = p->maux; info5[4] = p->no_transverse;
  return 0;
}

int claw_create(claw_solver** out, int ndim, int rp_id, int meqn, int mwaves, int maux, const int* n, int num_ghost,
                int device, void* stream) {
  if (!out) return fail(nullptr, 1, "claw_create: out is NULL");
  *out = nullptr;
  if (ndim != 2 && ndim != 3) return fail(nullptr, 2, "claw_create: ndim must be 2 or 3 (got %d)", ndim);
  if (num_ghost != 2) return fail(nullptr, 3, "claw_create: num_ghost must be 2 (classic solvers), got %d", num_ghost);
  if (meqn < 1 || meqn > 8 || mwaves < 1 || mwaves > 8 || maux < 0)
    return fail(nullptr, 4, "claw_create: bad sizes meqn=%d mwaves=%d maux=%d", meqn, mwaves, maux);
  static const int k_meqn[] = {0, 1, 3, 4, 3, 4, 5, 3, 1, 5, 3, 4, 3, 3, 3, 1, 1, 1, 1, 5, 2},
                   k_mwaves[] = {0, 1, 2, 4, 3, 2, 3, 2, 1, 5, 2, 2, 2, 2, 3, 1, 1, 1, 1, 4, 1},
                   k_ndim[] = {0, 2, 2, 2, 2, 3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 3, 2, 2, 2};
  if (rp_id >= CLAW_RP_PLUGIN_BASE) {                    // user-supplied solver (claw_rp_plugin_load)
    const RpPlugin* p = plugin_of(rp_id);
    if (!p) return fail(nullptr, 5, "claw_create: rp_id %d is not a registered Riemann-solver plug-in", rp_id);
    if (p->meqn != meqn || p->mwaves != mwaves || ndim != p->ndim)
      return fail(nullptr, 6, "claw_create: plug-in %s expects ndim=%d meqn=%d mwaves=%d (got %d, %d, %d)", p->path.c_str(),
                  p->ndim, p->meqn, p->mwaves, ndim, meqn, mwaves);
  } else {
  if (rp_id < 1 || rp_id > 20) return fail(nullptr, 5, "claw_create: unknown rp_id %d", rp_id);
  if (k_meqn[rp_id] != meqn || k_mwaves[rp_id] != mwaves || k_ndim[rp_id] != ndim)
    return fail(nullptr, 6, "claw_create: rp_id %d expects ndim=%d meqn=%d mwaves=%d (got %d, %d, %d)", rp_id,
                k_ndim[rp_id], k_meqn[rp_id], k_mwaves[rp_id], ndim, meqn, mwaves);
  }
  for (int d = 0; d < ndim; ++d)
    if (n[d] < 1) return fail(nullptr, 7, "claw_create: n[%d] = %d", d, n[d]);
  int ndev = claw_device_count();
  if (ndev == 0) return fail(nullptr, 8, "claw_create: no CUDA device available (the CUDA path has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(nullptr, 9, "claw_create: device %d out of range (%d devices)", device, ndev);
  claw_solver* h = new claw_solver();
  h->g.ndim = ndim;
  for (int d = 0; d < ndim; ++d) { h->g.n[d] = n[d]; h->g.nt[d] = n[d] + 2 * num_ghost; }
  h->g.ng = num_ghost;
  h->g.pitch = (h->g.nt[0] + 15) / 16 * 16;
  long long rows = 1;
  for (int d = 1; d < ndim; ++d) rows *= h->g.nt[d];
  h->g.plane = (long long)h->g.pitch * rows;
  h->g.kstride = (ndim == 3) ? (long long)h->g.pitch * h->g.nt[1] : 0;
  h->rp_id = rp_id; h->meqn = meqn; h->mwaves = mwaves; h->maux = maux;
  h->device = device;
  h->stream = (cudaStream_t)stream;
  h->q_elems = (size_t)h->g.plane * meqn;
  h->aux_elems = (size_t)h->g.plane * maux;
  long long cells = 1;
  for (int d = 0; d < ndim; ++d) cells *= h->g.nt[d];
  h->work_elems = WORK_HDR + (size_t)cells * (meqn > maux ? meqn : maux);
  // 3-D dimension-split steps use the staging area as a third q buffer (x: cur->alt, y: alt->tmp, z: tmp->alt)
  if (ndim == 3 && h->work_elems < WORK_HDR + h->q_elems) h->work_elems = WORK_HDR + h->q_elems;
  cudaError_t e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  if (e == cudaSuccess) e = cudaHostAlloc(&h->h_cfl, 64, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&h->h_ctl, WORK_HDR * sizeof(double), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_cfl, cudaEventDisableTiming);
  if (e == cudaSuccess && ndim == 2) {
    // tiles of the finest march segmentation the launch code uses (8 rows) on 28-column strips
    const long long strips = (n[0] + 27) / 28, segs = (n[1] + 4 + 7) / 8 + 1;
    if (strips * segs < (1LL << 26)) {
      h->hot_tiles = (int)(strips * segs);
      h->hot_cap = (int)(strips * (segs / 16 + 1));
      const size_t bytes = (8 + 3 * (size_t)h->hot_cap + 3 * (size_t)h->hot_tiles) * sizeof(unsigned);
      e = cudaMalloc(&h->hot, bytes);
      if (e == cudaSuccess) e = cudaMemset(h->hot, 0, bytes);
    }
  }
  if (e != cudaSuccess) {
    int rc = fail(nullptr, 100 + (int)e, "claw_create: %s", cudaGetErrorString(e));
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int claw_destroy(claw_solver* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->owns_buffers) {
    cudaFree(h->q[0]); cudaFree(h->q[1]); cudaFree(h->aux); cudaFree(h->work);
  }
  if (h->h_cfl) cudaFreeHost(h->h_cfl);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->hot) cudaFree(h->hot);
  if (h->h_hot) cudaFreeHost(h->h_hot);
  if (h->ev_hot) cudaEventDestroy(h->ev_hot);
  for (int k = 0; k < 4; ++k) if (h->hx[k]) cudaFree(h->hx[k]);
  for (int r = 0; r < 64; ++r)
    for (int k = 0; k < 4; ++k) if (h->peer.opened[r][k]) cudaIpcCloseMemHandle(h->peer.opened[r][k]);
  if (h->peer.mail) cudaFree(h->peer.mail);
  if (h->graph) cudaGraphExecDestroy(h->graph);
  if (h->s_cap) cudaStreamDestroy(h->s_cap);
  if (h->ev_cfl) cudaEventDestroy(h->ev_cfl);
  if (h->s_in) {
    cudaStreamDestroy(h->s_in); cudaStreamDestroy(h->s_cmp); cudaStreamDestroy(h->s_out);
    for (int k = 0; k < 64; ++k) { cudaEventDestroy(h->ev_in[k]); cudaEventDestroy(h->ev_cmp[k]); }
  }
  delete h;
  return 0;
}

int claw_buffer_sizes(const claw_solver* h, size_t* q_elems, size_t* aux_elems, size_t* work_elems) {
  if (check_handle(h)) return 1;
  if (q_elems) *q_elems = h->q_elems;
  if (aux_elems) *aux_elems = h->aux_elems;
  if (work_elems) *work_elems = h->work_elems;
  return 0;
}

int claw_bind(claw_solver* h, double* d_q0, double* d_q1, double* d_aux, double* d_work) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  if (h->owns_buffers) return fail(h, 10, "claw_bind: buffers were already allocated internally");
  if (!d_q0 || !d_q1 || !d_work || (h->maux > 0 && !d_aux)) return fail(h, 11, "claw_bind: NULL buffer");
  h->q[0] = d_q0; h->q[1] = d_q1; h->aux = d_aux; h->work = d_work;
  h->cur = 0;
  ++h->cfg_epoch;
  return 0;
}

int claw_set_cparam(claw_solver* h, const double* cparam, int n) {
  if (check_handle(h)) return 1;
  ++h->cfg_epoch;
  if (n < 0 || n > 4) return fail(h, 12, "claw_set_cparam: n=%d (max 4)", n);
  for (int k = 0; k < n; ++k) h->cparam[k] = cparam[k];
  return 0;
}

int claw_set_method(claw_solver* h, const int* method7, const int* mthlim, int use_fwave) {
  if (check_handle(h)) return 1;
  ++h->cfg_epoch;
  for (int k = 0; k < 7; ++k) h->method[k] = method7[k];
  for (int k = 0; k < h->mwaves; ++k) h->mthlim[k] = mthlim[k];
  h->use_fwave = use_fwave;
  return validate_config(h);
}

int claw_set_bc(claw_solver* h, const int* lower, const int* upper, const int* aux_lower, const int* aux_upper,
                const int* reflect_index) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  ++h->cfg_epoch;
  for (int d = 0; d < h->g.ndim; ++d) {
    h->bc_lower[d] = lower[d]; h->bc_upper[d] = upper[d];
    if (aux_lower) h->aux_bc_lower[d] = aux_lower[d];
    if (aux_upper) h->aux_bc_upper[d] = aux_upper[d];
    if (reflect_index) h->reflect_index[d] = reflect_index[d];
  }
  return 0;
}

int claw_set_boundary_ownership(claw_solver* h, const int* on_lower, const int* on_upper) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  ++h->cfg_epoch;
  for (int d = 0; d < h->g.ndim; ++d) { h->on_lower[d] = on_lower[d]; h->on_upper[d] = on_upper[d]; }
  return 0;
}

int claw_set_source(claw_solver* h, int src_id, const double* params, int nparams) {
  if (check_handle(h)) return 1;
  ++h->cfg_epoch;
  if (src_id == 0) { h->src_id = 0; return 0; }
  if (src_id != CLAW_SRC_EULER_RADIAL_2D) return fail(h, 60, "claw_set_source: unknown source functor %d", src_id);
  if (h->g.ndim != 2 || h->meqn < 4 || h->maux < 1)
    return fail(h, 61, "claw_set_source: the radial Euler source needs a 2-D state with >= 4 components and aux(1) = radius "
                       "(ndim=%d meqn=%d maux=%d)", h->g.ndim, h->meqn, h->maux);
  if (nparams < 1 || !params) return fail(h, 62, "claw_set_source: the radial Euler source takes one parameter (gamma)");
  h->src_id = src_id;
  for (int k = 0; k < 4; ++k) h->src_params[k] = (k < nparams) ? params[k] : 0.0;
  return 0;
}

int claw_set_bc_state(claw_solver* h, int dim, int side, const double* values) {
  if (check_handle(h)) return 1;
  if (dim < 0 || dim >= h->g.ndim || side < 0 || side > 1 || !values) return fail(h, 63, "claw_set_bc_state: bad arguments");
  ++h->cfg_epoch;
  for (int m = 0; m < 8; ++m) h->bc_state[dim][side][m] = (m < h->meqn) ? values[m] : 0.0;
  return 0;
}

int claw_set_mode(claw_solver* h, int mode) {
  if (check_handle(h)) return 1;
  ++h->cfg_epoch;
  if (mode != CLAW_MODE_STRICT && mode != CLAW_MODE_FAST) return fail(h, 13, "claw_set_mode: unknown mode %d", mode);
  h->mode = mode;
  return 0;
}

int claw_set_tuning(claw_solver* h, int seg_rows) {
  if (check_handle(h)) return 1;
  ++h->cfg_epoch;
  h->seg_rows = seg_rows < 0 ? 0 : seg_rows;
  return 0;
}

int claw_put_q(claw_solver* h, const double* host_q) {
  if (check_handle(h)) return 1;
  if (int rc = ensure_buffers(h)) return rc;
  if (int rc = ctl_settle(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  if (int rc = put_region(h, h->q[h->cur], h->meqn, host_q, false)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int claw_get_q(claw_solver* h, double* host_q) {
  if (check_handle(h)) return 1;
  if (int rc = ensure_buffers(h)) return rc;
  if (int rc = ctl_settle(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  return get_interior(h, h->q[h->cur], h->meqn, host_q, false);
}

// q of `ncells` interior cells (0-based interior indices, ndim ints per cell) of the CURRENT state -> host_out[c*meqn+m]
// (gauges, solver.py:680-715: a few cells per step instead of a full-array download)
int claw_get_cells(claw_solver* h, int ncells, const int* idx, double* host_out) {
  if (check_handle(h)) return 1;
  if (ncells < 0 || (ncells > 0 && (!idx || !host_out))) return fail(h, 12, "claw_get_cells: bad arguments");
  if (int rc = ensure_buffers(h)) return rc;
  if (int rc = ctl_settle(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  const Dev2& g = h->g;
  double* stage = h->work + WORK_HDR;
  for (int c0 = 0; c0 < ncells; c0 += 32) {
    CellList cl{};
    cl.n = ncells - c0 < 32 ? ncells - c0 : 32;
    for (int c = 0; c < cl.n; ++c) {
      const int* ix = idx + (size_t)(c0 + c) * g.ndim;
      for (int d = 0; d < g.ndim; ++d)
        if (ix[d] < 0 || ix[d] >= g.n[d]) return fail(h, 12, "claw_get_cells: cell index %d out of range in dimension %d", ix[d], d);
      cl.off[c] = (long long)(ix[0] + 2) + (long long)(ix[1] + 2) * g.pitch + (g.ndim == 3 ? (long long)(ix[2] + 2) * g.kstride : 0);
    }
    gather_cells<<<(cl.n * h->meqn + 63) / 64, 64, 0, h->stream>>>(h->q[h->cur], h->meqn, g.plane, cl, stage);
    ++h->launches;
    CK(h, cudaMemcpyAsync(host_out + (size_t)c0 * h->meqn, stage, sizeof(double) * cl.n * h->meqn, cudaMemcpyDeviceToHost,
                          h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int claw_put_aux(claw_solver* h, const double* host_aux) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  if (h->maux == 0) return 0;
  if (int rc = ensure_buffers(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  if (int rc = put_region(h, h->aux, h->maux, host_aux, false)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int claw_put_qbc(claw_solver* h, const double* host_qbc) {
  if (check_handle(h)) return 1;
  if (int rc = ensure_buffers(h)) return rc;
  if (int rc = ctl_settle(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  if (int rc = put_region(h, h->q[h->cur], h->meqn, host_qbc, true)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int claw_get_qbc(claw_solver* h, double* host_qbc) {
  if (check_handle(h)) return 1;
  if (int rc = ensure_buffers(h)) return rc;
  if (int rc = ctl_settle(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  return get_all(h, h->q[h->cur], h->meqn, host_qbc);
}

int claw_put_auxbc(claw_solver* h, const double* host_auxbc) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  if (h->maux == 0) return 0;
  if (int rc = ensure_buffers(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  if (int rc = put_region(h, h->aux, h->maux, host_auxbc, true)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int claw_apply_bcs(claw_solver* h) {
  if (check_handle(h)) return 1;
  if (int rc = ensure_buffers(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  // aux is static between uploads: its ghost cells are filled once after every change (upload, halo exchange, BC change),
  // not once per step attempt (3 of the 6 ghost-fill launches of a 3-D attempt)
  if (h->maux > 0 && !h->aux_ghosts_valid) {
    if (int rc = fill_ghosts_array(h, sel_fixed(h->aux), h->maux, h->aux_bc_lower, h->aux_bc_upper, false)) return rc;
    if (!(h->s_cap != nullptr && h->stream == h->s_cap)) h->aux_ghosts_valid = true;    // (not while a graph is captured)
  }
  return fill_ghosts_array(h, sel_q(h, 0), h->meqn, h->bc_lower, h->bc_upper, true);
}

int claw_step_launch(claw_solver* h, const double* delta, double dt, int fill_ghosts) {
  if (check_handle(h)) return 1;
  if (int rc = validate_config(h)) return rc;
  if (int rc = ensure_buffers(h)) return rc;
  if (int rc = ctl_leave(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  if (fill_ghosts)
    if (int rc = claw_apply_bcs(h)) return rc;
  CK(h, cudaMemsetAsync(h->work, 0, sizeof(double), h->stream));
  if (int rc = launch_step(h, delta, dt, h->q[h->cur], h->q[h->cur ^ 1])) return rc;
  CK(h, cudaMemcpyAsync(h->h_cfl, h->work, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaEventRecord(h->ev_cfl, h->stream));
  return 0;
}

int claw_cfl_device_ptr(claw_solver* h, void** d_cfl) {
  if (check_handle(h)) return 1;
  if (int rc = ensure_buffers(h)) return rc;
  *d_cfl = h->work;
  return 0;
}

int claw_cfl_wait(claw_solver* h, double* cfl) {
  if (check_handle(h)) return 1;
  CK(h, cudaEventSynchronize(h->ev_cfl));
  if (cfl) *cfl = *h->h_cfl;
  return 0;
}

// the CFL slot as it is NOW on the device (after a reduction across ranks queued behind claw_step_launch); waits
int claw_cfl_read(claw_solver* h, double* cfl) {
  if (check_handle(h)) return 1;
  if (!h->work || !cfl) return fail(h, 15, "claw_cfl_read: no buffers / NULL argument");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemcpyAsync(h->h_cfl, h->work, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  *cfl = *h->h_cfl;
  return 0;
}

int claw_step(claw_solver* h, const double* delta, double dt, int fill_ghosts, double* cfl) {
  if (int rc = claw_step_launch(h, delta, dt, fill_ghosts)) return rc;
  return claw_cfl_wait(h, cfl);
}

int claw_accept(claw_solver* h) {
  if (check_handle(h)) return 1;
  if (int rc = ctl_leave(h)) return rc;
  h->cur ^= 1;
  return 0;
}

int claw_reject(claw_solver* h) {
  if (check_handle(h)) return 1;
  return 0;  // `cur` still holds the pre-step state (its ghost cells were refilled, its interior is untouched)
}

int claw_sync(claw_solver* h) {
  if (check_handle(h)) return 1;
  CK(h, cudaStreamSynchronize(h->stream));
  return 0;
}


// ---- device-resident step control ----------------------------------------------------------
namespace {
// one attempt under the control block: [ghost fill] + sweep(s); the control kernel follows (claw_ctl_end)
PeerMail peer_mail_of(const claw_solver* h) {
  PeerMail pm{};
  pm.local = h->peer.mail;
  for (int r = 0; r < MAIL_MAXR; ++r) pm.remote[r] = h->peer.peer_mail[r];
  pm.nranks = h->peer.nranks; pm.rank = h->peer.rank;
  return pm;
}
// my interior rows / planes next to the rank faces -> the neighbours' ghost rows / planes.  role 0 / 1 = the current /
// the alternate q buffer (picked on the device under a StepCtl); aux = the static aux array instead
int peer_push(claw_solver* h, int role, bool aux) {
  const Dev2& g = h->g;
  const int last = g.ndim - 1, ng = g.ng;
  const int ncomp = aux ? h->maux : h->meqn;
  if (ncomp == 0) return 0;
  // one "layer" of the slab dimension: a row (2-D, pitch doubles) or a plane (3-D, kstride doubles)
  const long long layer = (g.ndim == 2) ? (long long)g.pitch : g.kstride;
  const long long len = ng * layer;
  auto sel_remote = [&](double* const* q2, double* a) {
    ArrSel d{};
    if (aux) { d.q[0] = a; return d; }
    if (!h->ctl_mode) { d.q[0] = q2[role == 0 ? h->cur : (h->cur ^ 1)]; return d; }
    d.q[0] = q2[0]; d.q[1] = q2[1]; d.q[2] = nullptr;
    d.ctl = reinterpret_cast<const StepCtl*>(h->work);
    d.role = role;
    return d;
  };
  const ArrSel src = aux ? sel_fixed(h->aux) : sel_q(h, role);
  // channel A: my upper interior layers [n, n+ng) -> the upper neighbour's lower ghosts [0, ng);
  // channel B: my lower interior layers [ng, 2ng) -> the lower neighbour's upper ghosts.  One launch for both.
  PeerChannel ch[2];
  int nch = 0;
  if (h->peer.up >= 0)
    ch[nch++] = PeerChannel{(long long)g.n[last] * layer, sel_remote(h->peer.up_q, h->peer.up_aux), 0, h->peer.up_plane};
  if (h->peer.lo >= 0)
    ch[nch++] = PeerChannel{(long long)ng * layer, sel_remote(h->peer.lo_q, h->peer.lo_aux),
                            (long long)(h->peer.lo_nlast + ng) * layer, h->peer.lo_plane};
  if (nch > 0) {
    if (nch == 1) ch[1] = ch[0];
    const dim3 blocks(grid_for((len + 1) / 2, 256), ncomp, nch);
    peer_copy_runs<<<blocks, 256, 0, h->stream>>>(src, g.plane, ch[0], ch[1], len, ncomp);
    ++h->launches;
  }
  CK(h, cudaGetLastError());
  return 0;
}

int ctl_enqueue_sweep(claw_solver* h, const double* delta, int fill_ghosts) {
  if (fill_ghosts)
    if (int rc = claw_apply_bcs(h)) return rc;
  return launch_step(h, delta, 0.0, nullptr, nullptr, true);
}
int ctl_enqueue_end(claw_solver* h) {
  if (h->src_id == 1) {                                // source step on the buffer the sweeps wrote, before accept/reject
    const Dev2& g = h->g;
    const long long n = (long long)g.n[0] * g.n[1];
    src_euler_radial2<<<grid_for(n, 256), 256, 0, h->stream>>>(sel_q(h, 1), h->aux, g.n[0], g.n[1], g.pitch, g.plane,
                                                               h->src_params[0], 0.0);
    ++h->launches;
  }
  if (h->peer.active) {
    // the rows / planes next to the rank faces of the buffer this attempt wrote go straight into the neighbours' ghost
    // cells of THEIR alternate buffer (it becomes current on every rank together, or is discarded on every rank)
    // A dimension-split 3-D step (step3ds x 3) passes the WHOLE array through its three buffers, ghost planes included: a
    // neighbour that is one pass ahead must not write my ghost planes of the alternate buffer before my last pass has
    // gone over them, nor while my y-sweep still reads them.  All ranks finish their sweeps, then everybody pushes.
    // (The unsplit kernels and the 2-D ones only ever write interior cells: no barrier there.)
    if (h->g.ndim == 3 && h->method[2] < 0) {
      peer_barrier_ctl_kernel<<<1, MAIL_MAXR, 0, h->stream>>>(reinterpret_cast<StepCtl*>(h->work), peer_mail_of(h));
      ++h->launches;
    }
    if (int rc = peer_push(h, 1, false)) return rc;
    ctl_end_peer_kernel<<<1, MAIL_MAXR, 0, h->stream>>>(reinterpret_cast<StepCtl*>(h->work), h->work + CTL_WORDS,
                                                       peer_mail_of(h));
  } else {
    ctl_end_kernel<<<1, 1, 0, h->stream>>>(reinterpret_cast<StepCtl*>(h->work), h->work + CTL_WORDS);
  }
  ++h->launches;
  CK(h, cudaGetLastError());
  return 0;
}
}  // namespace

int claw_ctl_begin(claw_solver* h, const double* host_words, double tend, int take_one_step, long long max_steps) {
  if (check_handle(h)) return 1;
  if (int rc = validate_config(h)) return rc;
  if (int rc = ensure_buffers(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  CtlBeginArgs a{};
  a.tend = tend; a.take_one = take_one_step ? 1 : 0; a.max_steps = max_steps;
  if (host_words) {
    if (int rc = ctl_settle(h)) return rc;
    if (!h->ctl_mode) CK(h, cudaMemsetAsync(h->work, 0, CTL_WORDS * sizeof(double), h->stream));   // attempts = 0
    memcpy(&a.host, host_words, sizeof(StepCtl));
    a.use_host = 1;
    a.cur = h->cur;
    h->ctl_mode = true;
  } else if (!h->ctl_mode) {
    return fail(h, 50, "claw_ctl_begin: the first call on a handle must pass the host's control values");
  }
  ctl_begin_kernel<<<1, 1, 0, h->stream>>>(reinterpret_cast<StepCtl*>(h->work), a);
  ++h->launches;
  CK(h, cudaGetLastError());
  h->ctl_pending = true;
  return 0;
}

int claw_ctl_sweep(claw_solver* h, const double* delta, int fill_ghosts) {
  if (check_handle(h)) return 1;
  if (!h->ctl_mode) return fail(h, 51, "claw_ctl_sweep: call claw_ctl_begin first");
  CK(h, cudaSetDevice(h->device));
  h->ctl_pending = true;
  return ctl_enqueue_sweep(h, delta, fill_ghosts);
}

int claw_ctl_end(claw_solver* h) {
  if (check_handle(h)) return 1;
  if (!h->ctl_mode) return fail(h, 51, "claw_ctl_end: call claw_ctl_begin first");
  CK(h, cudaSetDevice(h->device));
  h->ctl_pending = true;
  return ctl_enqueue_end(h);
}

int claw_ctl_attempts(claw_solver* h, const double* delta, int n, int fill_ghosts) {
  if (check_handle(h)) return 1;
  if (!h->ctl_mode) return fail(h, 51, "claw_ctl_attempts: call claw_ctl_begin first");
  if (n < 0 || n > LOG_CAP) return fail(h, 52, "claw_ctl_attempts: n = %d (0..%d)", n, LOG_CAP);
  CK(h, cudaSetDevice(h->device));
  h->ctl_pending = true;
  const int ndim = h->g.ndim;
  if (fill_ghosts && h->maux > 0 && !h->aux_ghosts_valid) {      // before a graph is captured: it then holds no aux fill
    if (int rc = fill_ghosts_array(h, sel_fixed(h->aux), h->maux, h->aux_bc_lower, h->aux_bc_upper, false)) return rc;
    h->aux_ghosts_valid = true;
  }
  dense_update(h);                                   // may switch kernels (and with it re-capture the graph below)
  bool graph_ok = h->use_graph && n >= 1;          // (a single attempt too: the graph is captured once per configuration)
  if (graph_ok) {
    bool stale = (h->graph == nullptr) || h->graph_epoch != h->cfg_epoch || h->graph_fill != (fill_ghosts ? 1 : 0);
    for (int d = 0; d < ndim; ++d) stale = stale || (h->graph_delta[d] != delta[d]);
    if (stale) {
      // capture ONE attempt; every buffer and dt is picked up from the control block when the node runs
      if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
      const uint64_t launches0 = h->launches;
      cudaGraph_t g = nullptr;
      // captured on a private stream (the handle's stream may be the legacy default stream, which cannot capture);
      // the instantiated graph is launched into the handle's stream
      if (!h->s_cap && cudaStreamCreateWithFlags(&h->s_cap, cudaStreamNonBlocking) != cudaSuccess) h->s_cap = nullptr;
      int rc = 1;
      cudaError_t e = cudaErrorUnknown;
      if (h->s_cap && cudaStreamBeginCapture(h->s_cap, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        cudaStream_t user = h->stream;
        h->stream = h->s_cap;
        rc = ctl_enqueue_sweep(h, delta, fill_ghosts);
        if (!rc) rc = ctl_enqueue_end(h);
        h->stream = user;
        e = cudaStreamEndCapture(h->s_cap, &g);
      }
      h->graph_nodes = h->launches - launches0;      // kernels per attempt
      h->launches = launches0;                       // nothing ran
      if (rc || e != cudaSuccess || g == nullptr) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        h->use_graph = 0;                            // fall back to plain launches for good
        graph_ok = false;
      } else {
        e = cudaGraphInstantiate(&h->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { cudaGetLastError(); h->graph = nullptr; h->use_graph = 0; graph_ok = false; }
        else {
          h->graph_epoch = h->cfg_epoch; h->graph_fill = fill_ghosts ? 1 : 0;
          for (int d = 0; d < ndim; ++d) h->graph_delta[d] = delta[d];
        }
      }
    }
  }
  if (graph_ok) {
    for (int k = 0; k < n; ++k) CK(h, cudaGraphLaunch(h->graph, h->stream));
    h->launches += h->graph_nodes * (uint64_t)n;
    return n > 0 ? dense_collect(h) : 0;
  }
  for (int k = 0; k < n; ++k) {
    if (int rc = ctl_enqueue_sweep(h, delta, fill_ghosts)) return rc;
    if (int rc = ctl_enqueue_end(h)) return rc;
  }
  return n > 0 ? dense_collect(h) : 0;
}

int claw_ctl_read(claw_solver* h, double* words, long long from_attempt, double* log3, int max_entries, int* n_entries) {
  if (check_handle(h)) return 1;
  if (!h->ctl_mode) return fail(h, 51, "claw_ctl_read: call claw_ctl_begin first");
  h->ctl_pending = true;                             // force the read-back
  if (int rc = ctl_settle(h)) return rc;
  const StepCtl* c = reinterpret_cast<const StepCtl*>(h->h_ctl);
  if (words) memcpy(words, h->h_ctl, CTL_WORDS * sizeof(double));
  int n = 0;
  if (log3 && max_entries > 0) {
    long long first = from_attempt;
    if (c->attempts - first > LOG_CAP) first = c->attempts - LOG_CAP;     // older entries were overwritten
    if (c->attempts - first > max_entries) first = c->attempts - max_entries;
    n = (int)(c->attempts - first);
    if (n > 0) {
      const long long i0 = first % LOG_CAP, i1 = (first + n - 1) % LOG_CAP;
      double* dlog = h->work + CTL_WORDS;
      if (i1 >= i0) {
        CK(h, cudaMemcpyAsync(log3, dlog + 3 * i0, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
      } else {
        const long long n0 = LOG_CAP - i0;
        CK(h, cudaMemcpyAsync(log3, dlog + 3 * i0, sizeof(double) * 3 * n0, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaMemcpyAsync(log3 + 3 * n0, dlog, sizeof(double) * 3 * (n - n0), cudaMemcpyDeviceToHost, h->stream));
      }
      CK(h, cudaStreamSynchronize(h->stream));
    }
  }
  if (n_entries) *n_entries = n;
  return 0;
}

int claw_ctl_use_graph(claw_solver* h, int on) {
  if (check_handle(h)) return 1;
  h->use_graph = on ? 1 : 0;
  return 0;
}

namespace {
size_t halo_elems_of(const claw_solver* h, int ncomp) {
  return (h->g.ndim == 2) ? (size_t)ncomp * 2 * h->g.nt[0] : (size_t)ncomp * 2 * h->g.kstride;
}
// pack (interior rows/planes next to `side`) or unpack (ghost rows/planes of `side`) of one array
int halo_move(claw_solver* h, ArrSel arr, int ncomp, int side, double* d_buf, bool pack) {
  if (int rc = ensure_buffers(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  const Dev2& g = h->g;
  const int last = g.ndim - 1;
  const int s0 = pack ? (side == 0 ? g.ng : g.n[last])            // interior [2,4) or [m, m+2)
                      : (side == 0 ? 0 : g.n[last] + g.ng);        // ghost [0,2) or [m+2, m+4)
  if (g.ndim == 2) {
    const long long n = (long long)ncomp * 2 * g.nt[0];
    if (pack) halo_pack2<<<grid_for(n, 256), 256, 0, h->stream>>>(arr, d_buf, ncomp, g.nt[0], s0, g.pitch, g.plane);
    else halo_unpack2<<<grid_for(n, 256), 256, 0, h->stream>>>(arr, d_buf, ncomp, g.nt[0], s0, g.pitch, g.plane);
  } else {
    const long long len = 2 * g.kstride;
    const long long off = (long long)s0 * g.kstride;
    if (pack) gather_runs<<<grid_for(len * ncomp, 256), 256, 0, h->stream>>>(arr, off, g.plane, d_buf, len, ncomp);
    else scatter_runs<<<grid_for(len * ncomp, 256), 256, 0, h->stream>>>(arr, off, g.plane, d_buf, len, ncomp);
  }
  ++h->launches;
  CK(h, cudaGetLastError());
  return 0;
}
}  // namespace

int claw_halo_elems(const claw_solver* h, size_t* elems) {
  if (check_handle(h)) return 1;
  *elems = halo_elems_of(h, h->meqn);
  return 0;
}
int claw_halo_pack(claw_solver* h, int side, double* d_buf) {
  if (check_handle(h)) return 1;
  return halo_move(h, sel_q(h, 0), h->meqn, side, d_buf, true);
}
int claw_halo_unpack(claw_solver* h, int side, const double* d_buf) {
  if (check_handle(h)) return 1;
  return halo_move(h, sel_q(h, 0), h->meqn, side, const_cast<double*>(d_buf), false);
}
// the same on the buffer the last claw_step_launch WROTE (the state that becomes current on claw_accept): lets the
// host queue the next step's exchange right behind the sweep, before it has read the CFL number back
int claw_halo_pack_next(claw_solver* h, int side, double* d_buf) {
  if (check_handle(h)) return 1;
  return halo_move(h, sel_q(h, 1), h->meqn, side, d_buf, true);
}
int claw_halo_unpack_next(claw_solver* h, int side, const double* d_buf) {
  if (check_handle(h)) return 1;
  return halo_move(h, sel_q(h, 1), h->meqn, side, const_cast<double*>(d_buf), false);
}
int claw_halo_elems_aux(const claw_solver* h, size_t* elems) {
  if (check_handle(h)) return 1;
  *elems = halo_elems_of(h, h->maux);
  return 0;
}
int claw_halo_pack_aux(claw_solver* h, int side, double* d_buf) {
  if (check_handle(h)) return 1;
  if (h->maux == 0) return 0;
  return halo_move(h, sel_fixed(h->aux), h->maux, side, d_buf, true);
}
int claw_halo_unpack_aux(claw_solver* h, int side, const double* d_buf) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  if (h->maux == 0) return 0;
  return halo_move(h, sel_fixed(h->aux), h->maux, side, const_cast<double*>(d_buf), false);
}

// ---- peer-memory transport (CUDA IPC over NVLink) ------------------------------------------------------------------
namespace {
struct PeerInfo {                                   // what a rank tells the others (CLAW_PEER_INFO_BYTES = 320)
  cudaIpcMemHandle_t q0, q1, aux, mail;             // 4 x 64 bytes
  long long plane;
  int nlast, has_aux, ndim, pitch;
  char pad[320 - 4 * 64 - 8 - 16];
};
static_assert(sizeof(PeerInfo) == 320, "PeerInfo is exchanged as 320 bytes");
}  // namespace

int claw_own_buffers(claw_solver* h) {
  if (check_handle(h)) return 1;
  h->aux_ghosts_valid = false;
  return ensure_buffers(h);
}

int claw_peer_export(claw_solver* h, void* info320) {
  if (check_handle(h)) return 1;
  if (!info320) return fail(h, 70, "claw_peer_export: NULL buffer");
  if (int rc = ensure_buffers(h)) return rc;
  if (!h->owns_buffers)
    return fail(h, 71, "claw_peer_export: the buffers must be the library's own allocations (claw_own_buffers instead of "
                       "claw_bind): an IPC handle names a whole allocation");
  CK(h, cudaSetDevice(h->device));
  if (!h->peer.mail) {
    CK(h, cudaMalloc(&h->peer.mail, MAIL_WORDS * sizeof(unsigned long long)));
    CK(h, cudaMemset(h->peer.mail, 0, MAIL_WORDS * sizeof(unsigned long long)));
  }
  PeerInfo inf{};
  CK(h, cudaIpcGetMemHandle(&inf.q0, h->q[0]));
  CK(h, cudaIpcGetMemHandle(&inf.q1, h->q[1]));
  if (h->aux) CK(h, cudaIpcGetMemHandle(&inf.aux, h->aux));
  CK(h, cudaIpcGetMemHandle(&inf.mail, h->peer.mail));
  inf.plane = h->g.plane; inf.nlast = h->g.n[h->g.ndim - 1]; inf.has_aux = h->aux ? 1 : 0;
  inf.ndim = h->g.ndim; inf.pitch = h->g.pitch;
  std::memcpy(info320, &inf, sizeof inf);
  return 0;
}

int claw_peer_connect(claw_solver* h, int nranks, int rank, const void* infos, int send_up_to, int send_lo_to) {
  if (check_handle(h)) return 1;
  if (!infos || nranks < 1 || nranks > MAIL_MAXR || rank < 0 || rank >= nranks)
    return fail(h, 72, "claw_peer_connect: bad arguments (nranks = %d, at most %d)", nranks, MAIL_MAXR);
  if (!h->peer.mail) return fail(h, 73, "claw_peer_connect: call claw_peer_export first");
  if (send_up_to >= nranks || send_lo_to >= nranks) return fail(h, 72, "claw_peer_connect: neighbour out of range");
  CK(h, cudaSetDevice(h->device));
  const PeerInfo* all = static_cast<const PeerInfo*>(infos);
  auto& P = h->peer;
  P.nranks = nranks; P.rank = rank; P.up = send_up_to; P.lo = send_lo_to;
  auto open = [&](int r, int which, const cudaIpcMemHandle_t& hd, void** out) -> int {
    if (!P.opened[r][which]) {
      cudaError_t e = cudaIpcOpenMemHandle(&P.opened[r][which], hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        P.opened[r][which] = nullptr;
        cudaGetLastError();
        return fail(h, 74, "claw_peer_connect: cudaIpcOpenMemHandle of rank %d failed: %s", r, cudaGetErrorString(e));
      }
    }
    *out = P.opened[r][which];
    return 0;
  };
  for (int r = 0; r < nranks; ++r) {
    if (all[r].ndim != h->g.ndim || all[r].pitch != h->g.pitch)
      return fail(h, 75, "claw_peer_connect: rank %d has another grid layout", r);
    if (r == rank) { P.peer_mail[r] = P.mail; continue; }
    void* p = nullptr;
    if (int rc = open(r, 3, all[r].mail, &p)) return rc;
    P.peer_mail[r] = static_cast<unsigned long long*>(p);
  }
  auto neighbour = [&](int r, double** q2, double** a, long long* plane) -> int {
    if (r < 0) return 0;
    if (r == rank) {                                 // a periodic slab dimension on one rank never gets here
      return fail(h, 76, "claw_peer_connect: a rank cannot be its own neighbour");
    }
    void* p = nullptr;
    if (int rc = open(r, 0, all[r].q0, &p)) return rc;
    q2[0] = static_cast<double*>(p);
    if (int rc = open(r, 1, all[r].q1, &p)) return rc;
    q2[1] = static_cast<double*>(p);
    if (all[r].has_aux) {
      if (int rc = open(r, 2, all[r].aux, &p)) return rc;
      *a = static_cast<double*>(p);
    }
    *plane = all[r].plane;
    return 0;
  };
  if (int rc = neighbour(send_up_to, P.up_q, &P.up_aux, &P.up_plane)) return rc;
  if (int rc = neighbour(send_lo_to, P.lo_q, &P.lo_aux, &P.lo_plane)) return rc;
  if (send_lo_to >= 0) P.lo_nlast = all[send_lo_to].nlast;
  P.active = true;
  ++h->cfg_epoch;
  return 0;
}

namespace {
int peer_status_check(claw_solver* h, int* d_status, const char* what) {
  int st = 0;
  CK(h, cudaMemcpyAsync(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  if (st) return fail(h, 77, "%s: a peer rank did not answer (timeout)", what);
  return 0;
}
int* peer_status_word(claw_solver* h) {               // a word of the mailbox nobody else writes
  return reinterpret_cast<int*>(h->peer.mail + MAIL_TAGS + 8);
}
}  // namespace

int claw_peer_sync(claw_solver* h, int aux) {
  if (check_handle(h)) return 1;
  if (aux) h->aux_ghosts_valid = false;
  if (!h->peer.active) return fail(h, 73, "claw_peer_sync: no peers connected");
  CK(h, cudaSetDevice(h->device));
  if (int rc = peer_push(h, 0, aux != 0)) return rc;
  CK(h, cudaMemsetAsync(peer_status_word(h), 0, sizeof(int), h->stream));
  peer_barrier_kernel<<<1, MAIL_MAXR, 0, h->stream>>>(peer_mail_of(h), peer_status_word(h));
  ++h->launches;
  CK(h, cudaGetLastError());
  return peer_status_check(h, peer_status_word(h), "claw_peer_sync");
}

int claw_peer_allreduce_cfl(claw_solver* h, double* cfl) {
  if (check_handle(h)) return 1;
  if (!h->peer.active) return fail(h, 73, "claw_peer_allreduce_cfl: no peers connected");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaMemsetAsync(peer_status_word(h), 0, sizeof(int), h->stream));
  peer_allreduce_kernel<<<1, MAIL_MAXR, 0, h->stream>>>(reinterpret_cast<unsigned long long*>(h->work), peer_mail_of(h),
                                                       peer_status_word(h));
  ++h->launches;
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(h->h_cfl, h->work, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (int rc = peer_status_check(h, peer_status_word(h), "claw_peer_allreduce_cfl")) return rc;
  if (cfl) *cfl = *h->h_cfl;
  return 0;
}

// ---- NCCL through the C ABI -------------------------------------------------------------------------------------
namespace {
// the handful of NCCL entry points used, resolved from the NCCL already in the process (PyTorch's) or from the system's
struct NcclApi {
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, const void*, int) = nullptr;      // ncclUniqueId is passed BY VALUE: see init below
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* CommInitRankRaw = nullptr;
  bool ok = false;
};
struct NcclId { char bytes[128]; };                                  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
constexpr int kNcclFloat64 = 8, kNcclMax = 2;                        // ncclDataType_t / ncclRedOp_t values of NCCL 2.x

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // already loaded by the host (e.g. PyTorch)?
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    api.GetUniqueId = reinterpret_cast<int (*)(void*)>(sym("ncclGetUniqueId"));
    api.CommInitRankRaw = sym("ncclCommInitRank");
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<int (*)(const void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<int (*)(void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclRecv"));
    api.AllReduce =
        reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRankRaw && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send &&
             api.Recv && api.AllReduce;
  });
  return api;
}
int nccl_fail(claw_solver* h, int rc, const char* what) {
  const char* msg = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
  return fail(h, 300 + rc, "%s: NCCL error %d (%s)", what, rc, msg);
}
int nccl_need(claw_solver* h) {
  if (!nccl().ok) return fail(h, 299, "NCCL is not available in this process (dlopen of libnccl.so.2 failed)");
  return 0;
}
}  // namespace

int claw_nccl_unique_id(void* id128) {
  if (int rc = nccl_need(nullptr)) return rc;
  if (!id128) return fail(nullptr, 298, "claw_nccl_unique_id: NULL buffer");
  const int rc = nccl().GetUniqueId(id128);
  return rc ? nccl_fail(nullptr, rc, "ncclGetUniqueId") : 0;
}
int claw_nccl_comm_init(void** comm, int nranks, int rank, const void* id128, int device) {
  if (int rc = nccl_need(nullptr)) return rc;
  if (!comm || !id128) return fail(nullptr, 298, "claw_nccl_comm_init: NULL argument");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, 100 + (int)e, "claw_nccl_comm_init: %s", cudaGetErrorString(e));
  NcclId id;
  std::memcpy(id.bytes, id128, sizeof(id.bytes));
  // ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId commId, int rank): the id travels by value
  auto init = reinterpret_cast<int (*)(void**, int, NcclId, int)>(nccl().CommInitRankRaw);
  const int rc = init(comm, nranks, id, rank);
  return rc ? nccl_fail(nullptr, rc, "ncclCommInitRank") : 0;
}
int claw_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  if (int rc = nccl_need(nullptr)) return rc;
  const int rc = nccl().CommDestroy(comm);
  return rc ? nccl_fail(nullptr, rc, "ncclCommDestroy") : 0;
}

int claw_halo_exchange(claw_solver* h, void* comm, int send_up_to, int recv_lo_from, int send_lo_to, int recv_up_from,
                       int aux) {
  if (check_handle(h)) return 1;
  if (aux) h->aux_ghosts_valid = false;
  if (int rc = nccl_need(h)) return rc;
  if (!comm) return fail(h, 298, "claw_halo_exchange: NULL communicator");
  const int ncomp = aux ? h->maux : h->meqn;
  if (ncomp == 0) return 0;
  CK(h, cudaSetDevice(h->device));
  const size_t need = halo_elems_of(h, h->meqn > h->maux ? h->meqn : h->maux);
  if (h->hx_elems < need) {
    for (int k = 0; k < 4; ++k) {
      if (h->hx[k]) CK(h, cudaFree(h->hx[k]));
      h->hx[k] = nullptr;
      CK(h, cudaMalloc(&h->hx[k], need * sizeof(double)));
    }
    h->hx_elems = need;
  }
  const size_t n = halo_elems_of(h, ncomp);
  const ArrSel arr = aux ? sel_fixed(h->aux) : sel_q(h, 0);
  if (send_up_to >= 0) if (int rc = halo_move(h, arr, ncomp, 1, h->hx[0], true)) return rc;
  if (send_lo_to >= 0) if (int rc = halo_move(h, arr, ncomp, 0, h->hx[2], true)) return rc;
  NcclApi& N = nccl();
  int rc = N.GroupStart();
  if (rc) return nccl_fail(h, rc, "ncclGroupStart");
  // every rank posts (send A, recv A, send B, recv B) in this order: two messages between the same pair match in order
  if (!rc && send_up_to >= 0) rc = N.Send(h->hx[0], n, kNcclFloat64, send_up_to, comm, h->stream);
  if (!rc && recv_lo_from >= 0) rc = N.Recv(h->hx[1], n, kNcclFloat64, recv_lo_from, comm, h->stream);
  if (!rc && send_lo_to >= 0) rc = N.Send(h->hx[2], n, kNcclFloat64, send_lo_to, comm, h->stream);
  if (!rc && recv_up_from >= 0) rc = N.Recv(h->hx[3], n, kNcclFloat64, recv_up_from, comm, h->stream);
  const int rc2 = N.GroupEnd();
  if (rc) return nccl_fail(h, rc, "ncclSend / ncclRecv");
  if (rc2) return nccl_fail(h, rc2, "ncclGroupEnd");
  if (recv_lo_from >= 0) if (int r = halo_move(h, arr, ncomp, 0, h->hx[1], false)) return r;
  if (recv_up_from >= 0) if (int r = halo_move(h, arr, ncomp, 1, h->hx[3], false)) return r;
  return 0;
}

int claw_allreduce_cfl(claw_solver* h, void* comm) {
  if (check_handle(h)) return 1;
  if (int rc = nccl_need(h)) return rc;
  if (!comm) return fail(h, 298, "claw_allreduce_cfl: NULL communicator");
  if (!h->work) return fail(h, 297, "claw_allreduce_cfl: no buffers bound");
  CK(h, cudaSetDevice(h->device));
  // the CFL slot holds a non-negative double; max over ranks in place (petclaw/cfl.py:36)
  const int rc = nccl().AllReduce(h->work, h->work, 1, kNcclFloat64, kNcclMax, comm, h->stream);
  return rc ? nccl_fail(h, rc, "ncclAllReduce") : 0;
}

int claw_time_sweep(claw_solver* h, const double* delta, double dt, int reps, double* avg_ms) {
  if (check_handle(h)) return 1;
  if (int rc = validate_config(h)) return rc;
  if (int rc = ensure_buffers(h)) return rc;
  if (reps < 1) return fail(h, 14, "claw_time_sweep: reps must be >= 1");
  if (int rc = ctl_settle(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  cudaEvent_t a, b;
  CK(h, cudaEventCreate(&a));
  CK(h, cudaEventCreate(&b));
  // the kernel a run on this state would settle on: one untimed launch of the fast-path kernels reports the heavy tiles,
  // the same rule as dense_decide picks (restored afterwards: the attempt graph was captured for the saved choice)
  const int saved_pref = h->dense_pref;
  if (dense_switch_enabled() && h->hot && h->g.ndim == 2) {
    h->dense_pref = 0;
    if (saved_pref) CK(h, cudaMemsetAsync(h->hot, 0, 3 * sizeof(unsigned), h->stream));   // (see dense_update)
    if (int rc = launch_step(h, delta, dt, h->q[h->cur], h->q[h->cur ^ 1])) return rc;
    if (!h->hot_copy_pending) {
      if (int rc = dense_collect(h)) return rc;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    h->hot_copy_pending = false;
    const double frac = dense_fraction(h);
    h->dense_pref = (frac > 0.6) ? 1 : 0;
  }
  CK(h, cudaMemsetAsync(h->work, 0, sizeof(double), h->stream));
  CK(h, cudaEventRecord(a, h->stream));
  for (int r = 0; r < reps; ++r)
    if (int rc = launch_step(h, delta, dt, h->q[h->cur], h->q[h->cur ^ 1])) { h->dense_pref = saved_pref; return rc; }
  CK(h, cudaEventRecord(b, h->stream));
  CK(h, cudaEventSynchronize(b));
  h->time_sweep_dense = h->dense_pref;
  h->dense_pref = saved_pref;
  float ms = 0.f;
  CK(h, cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *avg_ms = (double)ms / reps;
  return 0;
}

int claw_last_launch_info(const claw_solver* h, int* info5) {
  if (check_handle(h)) return 1;
  info5[0] = h->last_launch.grid_x; info5[1] = h->last_launch.grid_y; info5[2] = h->last_launch.block;
  info5[3] = h->last_launch.smem_bytes; info5[4] = h->last_launch.ctas_per_sm;
  return 0;
}

int claw_dense_state(const claw_solver* h, int* state3) {
  if (check_handle(h)) return 1;
  state3[0] = h->dense_pref; state3[1] = h->time_sweep_dense; state3[2] = dense_switch_enabled() ? 1 : 0;
  if (getenv("CLAW_DENSE_DEBUG")) {                    // development aid: the caller passes room for 12 ints
    state3[3] = h->dense_probe; state3[4] = h->hot_copy_pending; state3[5] = h->hot_copy_from_quiet;
    state3[6] = h->last_hot_rows; state3[7] = (int)h->last_ntiles; state3[8] = (int)h->dense_batches;
    state3[9] = h->h_hot ? (int)h->h_hot[0] : -1; state3[10] = h->h_hot ? (int)h->h_hot[1] : -1;
    state3[11] = h->h_hot ? (int)h->h_hot[2] : -1;
  }
  return 0;
}

int claw_launch_count(const claw_solver* h, uint64_t* count) {
  if (check_handle(h)) return 1;
  *count = h->launches;
  return 0;
}

// ---- f2py-shaped host entry points ---------------------------------------------------------
namespace {
std::mutex g_host_mutex;
claw_solver* g_host = nullptr;

int host_handle(int ndim, int rp_id, int meqn, int mwaves, int maux, const int* n, claw_solver** out) {
  if (g_host && (g_host->g.ndim != ndim || g_host->rp_id != rp_id || g_host->meqn != meqn || g_host->maux != maux ||
                 g_host->g.n[0] != n[0] || g_host->g.n[1] != n[1] || (ndim == 3 && g_host->g.n[2] != n[2]))) {
    claw_destroy(g_host);
    g_host = nullptr;
  }
  if (!g_host) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (int rc = claw_create(&g_host, ndim, rp_id, meqn, mwaves, maux, n, 2, dev, nullptr)) return rc;
  }
  *out = g_host;
  return 0;
}

int host_fail(claw_solver* h, int rc) {
  if (h) g_last_error = h->err;
  return rc;
}
}  // namespace

namespace {
// claw_step2_host for large grids: the rows are cut into chunks and three streams overlap
//   H2D + AoS->SoA of chunk c+1  |  sweep of chunk c  |  SoA->AoS + D2H of chunk c-1
// (PCIe is full duplex, so the step costs ~max(H2D, D2H) instead of H2D + sweep + D2H).  Chunk c needs the input
// rows [j0-2, j1+2) of its output rows [j0, j1); uploads are disjoint row ranges issued in order.
int step2_host_pipelined(claw_solver* h, const double* qold, double* qnew, double dx, double dy, double dt,
                         double* cfl) {
  const Dev2& g = h->g;
  const int meqn = h->meqn, mx = g.n[0], my = g.n[1], nx = g.nt[0], ny = g.nt[1];
  if (int rc = ensure_buffers(h)) return rc;
  CK(h, cudaSetDevice(h->device));
  if (!h->s_in) {
    CK(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CK(h, cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
    CK(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 64; ++k) {
      CK(h, cudaEventCreateWithFlags(&h->ev_in[k], cudaEventDisableTiming));
      CK(h, cudaEventCreateWithFlags(&h->ev_cmp[k], cudaEventDisableTiming));
    }
  }
  CK(h, cudaStreamSynchronize(h->stream));
  int nchunk = (my + 511) / 512;
  if (nchunk > 64) nchunk = 64;
  if (nchunk < 2) nchunk = 2;
  const int rows = (my + nchunk - 1) / nchunk;
  nchunk = (my + rows - 1) / rows;
  double* stage = h->work + WORK_HDR;                  // AoS image of the whole ghosted array
  const size_t row_elems = (size_t)meqn * nx;
  const double* src = h->q[h->cur];
  double* dst = h->q[h->cur ^ 1];
  CK(h, cudaMemsetAsync(h->work, 0, sizeof(double), h->s_cmp));
  int uploaded = 0;                                    // rows [0, uploaded) are on the device
  for (int c = 0; c < nchunk; ++c) {
    const int j0 = g.ng + c * rows, j1 = (c == nchunk - 1) ? g.ng + my : j0 + rows;
    const int need = (c == nchunk - 1) ? ny : j1 + 2;
    if (need > uploaded) {
      CK(h, cudaMemcpyAsync(stage + (size_t)uploaded * row_elems, qold + (size_t)uploaded * row_elems,
                            (size_t)(need - uploaded) * row_elems * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
      const long long ncell = (long long)nx * (need - uploaded);
      aos_to_soa<<<grid_for(ncell, 256), 256, 0, h->s_in>>>(stage + (size_t)uploaded * row_elems, h->q[h->cur], meqn, 0,
                                                            nx, uploaded, need, 0, 1, g.pitch, g.kstride, g.plane);
      ++h->launches;
      uploaded = need;
    }
    CK(h, cudaEventRecord(h->ev_in[c], h->s_in));
    CK(h, cudaStreamWaitEvent(h->s_cmp, h->ev_in[c], 0));
    if (int rc = launch_step2(h, dx, dy, dt, 0, 1, 1, src, dst, j0, j1, h->s_cmp)) return rc;
    CK(h, cudaEventRecord(h->ev_cmp[c], h->s_cmp));
    CK(h, cudaStreamWaitEvent(h->s_out, h->ev_cmp[c], 0));
    // interior cells of rows [j0, j1) -> staging (same rows of the AoS image, already consumed) -> qnew
    double* ostage = stage + (size_t)j0 * row_elems;
    const long long ncell = (long long)mx * (j1 - j0);
    soa_to_aos<<<grid_for(ncell, 256), 256, 0, h->s_out>>>(dst, ostage, meqn, g.ng, g.ng + mx, j0, j1, 0, 1, g.pitch,
                                                           g.kstride, g.plane);
    ++h->launches;
    const size_t wbytes = (size_t)meqn * mx * sizeof(double);
    CK(h, cudaMemcpy2DAsync(qnew + ((size_t)j0 * nx + g.ng) * meqn, row_elems * sizeof(double), ostage, wbytes, wbytes,
                            (size_t)(j1 - j0), cudaMemcpyDeviceToHost, h->s_out));
  }
  CK(h, cudaGetLastError());
  CK(h, cudaMemcpyAsync(h->h_cfl, h->work, sizeof(double), cudaMemcpyDeviceToHost, h->s_cmp));
  CK(h, cudaStreamSynchronize(h->s_cmp));
  CK(h, cudaStreamSynchronize(h->s_out));
  if (cfl) *cfl = *h->h_cfl;
  return 0;
}
}  // namespace

int claw_step2_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my, const double* qold,
                    double* qnew, const double* aux, double dx, double dy, double dt, const int* method7,
                    const int* mthlim, int use_fwave, int rp_id, const double* cparam, int ncparam, int mode,
                    double* cfl) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (num_ghost != 2) return fail(nullptr, 3, "claw_step2_host: num_ghost must be 2");
  if (maxm < mx || maxm < my) return fail(nullptr, 40, "claw_step2_host: maxm=%d < max(mx,my)", maxm);
  const int n[2] = {mx, my};
  claw_solver* h = nullptr;
  if (int rc = host_handle(2, rp_id, meqn, mwaves, maux, n, &h)) return rc;
  int rc;
  if ((rc = claw_set_mode(h, mode))) return host_fail(h, rc);
  if ((rc = claw_set_cparam(h, cparam, ncparam))) return host_fail(h, rc);
  if ((rc = claw_set_method(h, method7, mthlim, use_fwave))) return host_fail(h, rc);
  if (h->method[2] < 0) return host_fail(h, fail(h, 41, "claw_step2_host: method(3) = -1 (dimensional splitting) belongs to step2ds"));
  if (maux == 0 && (long long)mx * my >= (1LL << 20) && my >= 64) {
    rc = step2_host_pipelined(h, qold, qnew, dx, dy, dt, cfl);
    return rc ? host_fail(h, rc) : 0;
  }
  if ((rc = claw_put_qbc(h, qold))) return host_fail(h, rc);
  if (maux > 0 && (rc = claw_put_auxbc(h, aux))) return host_fail(h, rc);
  const double delta[2] = {dx, dy};
  if ((rc = claw_step(h, delta, dt, 0, cfl))) return host_fail(h, rc);
  rc = get_interior(h, h->q[h->cur ^ 1], meqn, qnew, true);
  return rc ? host_fail(h, rc) : 0;
}

int claw_step2ds_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my, const double* qold,
                      double* qnew, const double* aux, double dx, double dy, double dt, const int* method7,
                      const int* mthlim, int ids, int use_fwave, int rp_id, const double* cparam, int ncparam,
                      int mode, double* cfl) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (num_ghost != 2) return fail(nullptr, 3, "claw_step2ds_host: num_ghost must be 2");
  if (maxm < mx || maxm < my) return fail(nullptr, 40, "claw_step2ds_host: maxm=%d < max(mx,my)", maxm);
  if (ids != 1 && ids != 2) return fail(nullptr, 42, "claw_step2ds_host: ids must be 1 or 2 (got %d)", ids);
  const int n[2] = {mx, my};
  claw_solver* h = nullptr;
  if (int rc = host_handle(2, rp_id, meqn, mwaves, maux, n, &h)) return rc;
  int rc;
  if ((rc = claw_set_mode(h, mode))) return host_fail(h, rc);
  if ((rc = claw_set_cparam(h, cparam, ncparam))) return host_fail(h, rc);
  if ((rc = claw_set_method(h, method7, mthlim, use_fwave))) return host_fail(h, rc);
  if ((rc = claw_put_qbc(h, qold))) return host_fail(h, rc);
  if (maux > 0 && (rc = claw_put_auxbc(h, aux))) return host_fail(h, rc);
  if (cudaMemsetAsync(h->work, 0, sizeof(double), h->stream) != cudaSuccess) return fail(nullptr, 43, "memset failed");
  if ((rc = launch_step2(h, dx, dy, dt, 1, ids == 1, ids == 2, h->q[h->cur], h->q[h->cur ^ 1]))) return host_fail(h, rc);
  if (cudaMemcpyAsync(h->h_cfl, h->work, sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
    return fail(nullptr, 43, "cfl copy failed");
  const Dev2& g = h->g;
  {
    const int zero[3] = {0, 0, 0}, hn[3] = {g.nt[0], g.nt[1], 1};
    if (ids == 1) {   // rows: all; columns: interior   (step2ds.f90:96,149)
      const int lo[3] = {g.ng, 0, 0}, hi[3] = {g.ng + mx, g.nt[1], 1};
      rc = get_region(h, h->q[h->cur ^ 1], meqn, qnew, lo, hi, zero, hn);
    } else {          // columns: all; rows: interior   (step2ds.f90:180,237)
      const int lo[3] = {0, g.ng, 0}, hi[3] = {g.nt[0], g.ng + my, 1};
      rc = get_region(h, h->q[h->cur ^ 1], meqn, qnew, lo, hi, zero, hn);
    }
  }
  if (rc) return host_fail(h, rc);
  if (cfl) *cfl = *h->h_cfl;    // get_region synchronised the stream
  return 0;
}

int claw_step3_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my, int mz,
                    const double* qold, double* qnew, const double* aux, double dx, double dy, double dz, double dt,
                    const int* method7, const int* mthlim, int use_fwave, int rp_id, const double* cparam,
                    int ncparam, int mode, double* cfl) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (num_ghost != 2) return fail(nullptr, 3, "claw_step3_host: num_ghost must be 2");
  if (maxm < mx || maxm < my || maxm < mz) return fail(nullptr, 40, "claw_step3_host: maxm=%d < max(mx,my,mz)", maxm);
  const int n[3] = {mx, my, mz};
  claw_solver* h = nullptr;
  if (int rc = host_handle(3, rp_id, meqn, mwaves, maux, n, &h)) return rc;
  int rc;
  if ((rc = claw_set_mode(h, mode))) return host_fail(h, rc);
  if ((rc = claw_set_cparam(h, cparam, ncparam))) return host_fail(h, rc);
  if ((rc = claw_set_method(h, method7, mthlim, use_fwave))) return host_fail(h, rc);
  if (h->method[2] < 0) return host_fail(h, fail(h, 41, "claw_step3_host: method(3) = -1 (dimensional splitting) belongs to step3ds"));
  if ((rc = claw_put_qbc(h, qold))) return host_fail(h, rc);
  if (maux > 0 && (rc = claw_put_auxbc(h, aux))) return host_fail(h, rc);
  const double delta[3] = {dx, dy, dz};
  if ((rc = claw_step(h, delta, dt, 0, cfl))) return host_fail(h, rc);
  rc = get_interior(h, h->q[h->cur ^ 1], meqn, qnew, true);
  return rc ? host_fail(h, rc) : 0;
}

int claw_step3ds_host(int maxm, int meqn, int mwaves, int maux, int num_ghost, int mx, int my, int mz,
                      const double* qold, double* qnew, const double* aux, double dx, double dy, double dz, double dt,
                      const int* method7, const int* mthlim, int idir, int use_fwave, int rp_id,
                      const double* cparam, int ncparam, int mode, double* cfl) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (num_ghost != 2) return fail(nullptr, 3, "claw_step3ds_host: num_ghost must be 2");
  if (maxm < mx || maxm < my || maxm < mz) return fail(nullptr, 40, "claw_step3ds_host: maxm=%d < max(mx,my,mz)", maxm);
  if (idir < 1 || idir > 3) return fail(nullptr, 42, "claw_step3ds_host: idir must be 1, 2 or 3 (got %d)", idir);
  const int n[3] = {mx, my, mz};
  claw_solver* h = nullptr;
  if (int rc = host_handle(3, rp_id, meqn, mwaves, maux, n, &h)) return rc;
  int rc;
  if ((rc = claw_set_mode(h, mode))) return host_fail(h, rc);
  if ((rc = claw_set_cparam(h, cparam, ncparam))) return host_fail(h, rc);
  if ((rc = claw_set_method(h, method7, mthlim, use_fwave))) return host_fail(h, rc);
  if (h->method[2] >= 0) return host_fail(h, fail(h, 41, "claw_step3ds_host: needs method(3) = -1 (dimensional splitting)"));
  if ((rc = claw_put_qbc(h, qold))) return host_fail(h, rc);
  if (maux > 0 && (rc = claw_put_auxbc(h, aux))) return host_fail(h, rc);
  if (cudaMemsetAsync(h->work, 0, sizeof(double), h->stream) != cudaSuccess) return fail(nullptr, 43, "memset failed");
  const double delta[3] = {dx, dy, dz};
  if ((rc = launch_pass3(h, delta, dt, idir - 1, 1, h->q[h->cur], h->q[h->cur ^ 1]))) return host_fail(h, rc);
  if (cudaMemcpyAsync(h->h_cfl, h->work, sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
    return fail(nullptr, 43, "cfl copy failed");
  // the pass wrote every cell (cells outside step3ds's range are copied through from qold)
  if ((rc = get_all(h, h->q[h->cur ^ 1], meqn, qnew))) return host_fail(h, rc);
  if (cfl) *cfl = *h->h_cfl;    // get_region synchronised the stream
  return 0;
}

// ---- 1-D: q,cfl = step1(num_ghost,mx,q,aux,dx,dt,method,mthlim,use_fwave,rp1) (step1.f90:2-3, classic/solver.py:302) ----
namespace {
struct Host1 { double* d[4] = {nullptr, nullptr, nullptr, nullptr}; size_t cap[4] = {0, 0, 0, 0}; };
Host1 g_host1;                        // q in, q out, aux, cfl slot
int host1_reserve(int k, size_t bytes) {
  if (g_host1.cap[k] >= bytes) return 0;
  if (g_host1.d[k]) cudaFree(g_host1.d[k]);
  g_host1.d[k] = nullptr; g_host1.cap[k] = 0;
  if (cudaMalloc(&g_host1.d[k], bytes) != cudaSuccess) return 1;
  g_host1.cap[k] = bytes;
  return 0;
}
}  // namespace

int claw_step1_host(int meqn, int mwaves, int maux, int num_ghost, int mx, double* q, const double* aux, double dx,
                    double dt, const int* method7, const int* mthlim, int use_fwave, int rp_id, const double* cparam,
                    int ncparam, int mode, double* cfl) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (num_ghost != 2) return fail(nullptr, 3, "claw_step1_host: num_ghost must be 2");
  if (!q || !method7 || !mthlim || !cfl || mx < 1) return fail(nullptr, 12, "claw_step1_host: bad arguments");
  // rp_id 21..36 (include/clawcuda.h): num_eqn, num_waves (riemann/riemann/static.py) and the aux components read
  static const int k_meqn[] = {1, 2, 1, 2, 2, 1, 1, 2, 2, 3, 3, 2, 1, 1, 1, 2},
                   k_mwaves[] = {1, 2, 1, 2, 2, 1, 1, 2, 2, 3, 2, 2, 1, 1, 1, 2},
                   k_maux[] = {0, 0, 0, 0, 0, 0, 1, 2, 1, 0, 0, 3, 0, 1, 1, 1};
  if (rp_id < 21 || rp_id > 36) return fail(nullptr, 5, "claw_step1_host: unknown 1-D rp_id %d", rp_id);
  if (k_meqn[rp_id - 21] != meqn || k_mwaves[rp_id - 21] != mwaves)
    return fail(nullptr, 6, "claw_step1_host: rp_id %d expects meqn=%d mwaves=%d (got %d, %d)", rp_id,
                k_meqn[rp_id - 21], k_mwaves[rp_id - 21], meqn, mwaves);
  if (k_maux[rp_id - 21] > 0 && (!aux || maux < k_maux[rp_id - 21]))
    return fail(nullptr, 7, "claw_step1_host: rp_id %d reads %d aux component(s) (got num_aux=%d%s)", rp_id,
                k_maux[rp_id - 21], maux, aux ? "" : ", aux=NULL");
  if (method7[1] != 1 && method7[1] != 2) return fail(nullptr, 33, "order must be 1 or 2 (got %d)", method7[1]);
  if (method7[5] < 0 || method7[5] > maux) return fail(nullptr, 31, "index_capa=%d is outside the aux array (num_aux=%d)", method7[5] - 1, maux);
  if (method7[5] > 0 && !aux) return fail(nullptr, 31, "claw_step1_host: capacity function requested but aux is NULL");
  for (int mw = 0; mw < mwaves; ++mw)
    if (mthlim[mw] < 0 || mthlim[mw] > 5)
      return fail(nullptr, 35, "limiter id %d of wave family %d is not defined by philim.f90 (0..5)", mthlim[mw], mw);
  const size_t ncell = (size_t)mx + 4, qb = sizeof(double) * ncell * meqn, ab = sizeof(double) * ncell * (maux > 0 ? maux : 1);
  if (host1_reserve(0, qb) || host1_reserve(1, qb) || host1_reserve(2, ab) || host1_reserve(3, 64))
    return fail(nullptr, 20, "claw_step1_host: device allocation failed");
  cudaStream_t st = nullptr;
  if (cudaMemcpyAsync(g_host1.d[0], q, qb, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(g_host1.d[1], g_host1.d[0], qb, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
      cudaMemsetAsync(g_host1.d[3], 0, 64, st) != cudaSuccess)
    return fail(nullptr, 21, "claw_step1_host: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (maux > 0 && aux && cudaMemcpyAsync(g_host1.d[2], aux, ab, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return fail(nullptr, 21, "claw_step1_host: aux upload failed");
  Sweep2Params P{};
  P.qin = g_host1.d[0]; P.qout = g_host1.d[1]; P.aux = g_host1.d[2];
  P.cfl_bits = reinterpret_cast<unsigned long long*>(g_host1.d[3]);
  P.mx = mx; P.my = 1; P.pitch = meqn; P.plane = 1;          // PyClaw's (m, i) order
  P.dtdx = dt; P.dtdy = dx;                                   // the kernel forms dt/dx or dt/(dx*capa) itself
  P.order = method7[1]; P.mtrans = 0; P.limit = 0;
  for (int mw = 0; mw < mwaves; ++mw) { P.mthlim[mw] = mthlim[mw]; if (mthlim[mw] > 0) P.limit = 1; }
  P.maux = maux; P.capa = method7[5]; P.fwave = use_fwave;
  for (int k = 0; k < 4; ++k) P.rp.p[k] = (cparam && k < ncparam) ? cparam[k] : 0.0;
  cudaError_t e = (mode == CLAW_MODE_FAST) ? fast::launch_sweep1(rp_id, P, st, nullptr)
                                           : strict::launch_sweep1(rp_id, P, st, nullptr);
  if (e != cudaSuccess) return fail(nullptr, 100 + (int)e, "1-D sweep kernel launch failed: %s", cudaGetErrorString(e));
  double c = 0.0;
  if (cudaMemcpyAsync(q, g_host1.d[1], qb, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaMemcpyAsync(&c, g_host1.d[3], sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return fail(nullptr, 22, "claw_step1_host: download failed: %s", cudaGetErrorString(cudaGetLastError()));
  *cfl = c;
  return 0;
}

int claw_host_release(void) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (g_host) { claw_destroy(g_host); g_host = nullptr; }
  return 0;
}

int claw_host_alloc(void** p, size_t bytes) {
  cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(nullptr, 100 + (int)e, "claw_host_alloc: %s", cudaGetErrorString(e));
  return 0;
}

int claw_host_free(void* p) {
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) return fail(nullptr, 100 + (int)e, "claw_host_free: %s", cudaGetErrorString(e));
  return 0;
}

int claw_measure_fp64_peak(int device, double* gdfma_per_s) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, 100 + (int)e, "claw_measure_fp64_peak: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* d = nullptr;
  if ((e = cudaMalloc(&d, (size_t)blocks * threads * sizeof(double))) != cudaSuccess)
    return fail(nullptr, 100 + (int)e, "claw_measure_fp64_peak: %s", cudaGetErrorString(e));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    dfma_probe<<<blocks, threads>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double rate = (double)blocks * threads * iters * 8 / (ms * 1e-3) / 1e9;
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(d);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail(nullptr, 100 + (int)e, "dfma probe: %s", cudaGetErrorString(e));
  *gdfma_per_s = best;
  return 0;
}

}  // extern "C"
