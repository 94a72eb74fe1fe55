"""CUDA back end of the classic solvers: the bridge between the PyClaw-shaped host classes and
libclawcuda.so (include/clawcuda.h), via ctypes.  PyTorch only provides the device buffers.

* ``DeviceStepper``  -- device-resident state for ClawSolver2D/3D(kernel_language='CUDA').
* ``fmod(num_dim)``  -- object with the f2py signatures of the reference's ``classic2`` /
  ``classic3`` extension modules (step2, step2ds, step3, step3ds; step2.f90:2-5 etc.), running
  the same kernels on host arrays.  Assign it to ``solver.fmod`` wherever the reference would
  use its Fortran module.
"""
import ctypes
import math
import os
import weakref

import numpy as np

from .. import _lib

_MODES = {'strict': 0, 'fast': 1}
LOG_CAP = 4096        # attempts the device log ring holds (csrc/clawcuda.cu)
BATCH = 64            # attempts queued between two reads of the control block (evolve_to_time with an end time)
# control block words (include/clawcuda.h)
W_T, W_DT, W_DTMAX, W_CFLDES, W_CFLMAX, W_CFL = 1, 2, 3, 4, 5, 6
W_DTMIN, W_DTMAXS, W_CFLMAXS, W_NUMSTEPS, W_DONE, W_ERROR = 10, 11, 12, 13, 15, 16
W_DTVAR, W_HAVEOLD, W_ATTEMPTS = 18, 19, 23


class DeviceMirrored(object):
    """Solver attribute whose authoritative value may live in the device's control block while step attempts are
    queued: reading it waits for the queue and refreshes the host copy; writing it makes the host copy authoritative
    again (it is uploaded with the next claw_ctl_begin)."""

    def __init__(self, name):
        self.slot = '_m_' + name

    def __get__(self, obj, cls):
        if obj is None:
            return self
        d = obj.__dict__
        dev = d.get('_dev')
        if dev is not None and dev.pending:
            dev.drain()
        try:
            return d[self.slot]
        except KeyError:
            raise AttributeError(self.slot[3:])

    def __set__(self, obj, value):
        d = obj.__dict__
        dev = d.get('_dev')
        if dev is not None:
            if dev.pending:
                dev.drain()
            dev.host_ctl_dirty = True
        d[self.slot] = value


def _iarr(values, n=None):
    a = np.zeros(n if n is not None else len(values), dtype=np.intc)
    for k, v in enumerate(values):
        a[k] = -1 if v is None else int(v)
    return a


def _cparam_of(rp):
    cp = getattr(rp, 'cparam', None)
    vals = cp.values() if cp is not None else []
    return np.array(list(vals) + [0.0], dtype=np.float64), len(vals)


class DeviceStepper(object):
    """Owns the libclawcuda handle and the torch device buffers of one State."""

    def __init__(self, solver, state):
        import torch
        L = _lib.lib()
        if L.claw_device_count() == 0 or not torch.cuda.is_available():
            raise _lib.ClawCudaError("kernel_language='CUDA' needs a CUDA device; none is visible "
                                     "(there is no CPU fallback in this package)")
        self.L = L
        self.torch = torch
        dev = solver.cuda_device if solver.cuda_device is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', int(dev))
        self.num_dim = solver.num_dim
        n = _iarr(state.grid.num_cells)
        self.handle = ctypes.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(L.claw_create(ctypes.byref(self.handle), solver.num_dim, int(solver.rp.rp_id), state.num_eqn,
                                 solver.num_waves, state.num_aux, _lib.ip(n), solver.num_ghost, int(dev),
                                 ctypes.c_void_p(stream)))
        h = self.handle
        nq, na, nw = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(L.claw_buffer_sizes(h, ctypes.byref(nq), ctypes.byref(na), ctypes.byref(nw)), h)
        f64 = torch.float64
        # The peer-memory transport of multi-rank runs (petclaw.device_halo; the default, CLAW_HALO=peer) maps the neighbours' buffers
        # through CUDA IPC, which names whole allocations: the buffers then are the library's own (cudaMalloc each)
        # instead of tensors cut from PyTorch's caching allocator.
        self.own_buffers = (os.environ.get('CLAW_HALO', 'peer') == 'peer' and getattr(state.patch, 'num_ranks', 1) > 1)
        if self.own_buffers:
            _lib.check(L.claw_own_buffers(h), h)
            self.q0 = self.q1 = self.aux = self.work = self.cfl_tensor = None
        else:
            self.q0 = torch.zeros(nq.value, dtype=f64, device=self.device)
            self.q1 = torch.zeros(nq.value, dtype=f64, device=self.device)
            self.aux = torch.zeros(max(na.value, 1), dtype=f64, device=self.device)
            self.work = torch.zeros(nw.value, dtype=f64, device=self.device)
            _lib.check(L.claw_bind(h, ctypes.c_void_p(self.q0.data_ptr()), ctypes.c_void_p(self.q1.data_ptr()),
                                   ctypes.c_void_p(self.aux.data_ptr() if na.value else 0),
                                   ctypes.c_void_p(self.work.data_ptr())), h)
            self.cfl_tensor = self.work[0:1]      # the 8-byte CFL slot (for the multi-GPU max all-reduce)
        cp, ncp = _cparam_of(solver.rp)
        _lib.check(L.claw_set_cparam(h, _lib.dp(cp), ncp), h)
        mth = _iarr(solver._mthlim)
        _lib.check(L.claw_set_method(h, _lib.ip(np.ascontiguousarray(solver._method, dtype=np.intc)), _lib.ip(mth),
                                     int(bool(solver.fwave))), h)
        nd = solver.num_dim
        # BC.custom faces whose fill is a device functor (device_callbacks.FixedStateBC) become device BC types
        bc_lo, bc_up = list(solver.bc_lower), list(solver.bc_upper)
        for side, (types, fun) in enumerate(((bc_lo, solver.user_bc_lower), (bc_up, solver.user_bc_upper))):
            for d in range(nd):
                if types[d] == 0 and hasattr(fun, 'device_bc'):
                    bc_type, values = fun.device_bc
                    types[d] = bc_type
                    vals = np.zeros(8, dtype=np.float64)
                    vals[:len(values)] = values
                    _lib.check(L.claw_set_bc_state(h, d, side, _lib.dp(vals)), h)
        if solver.step_source is not None:
            src_id, params = solver.step_source.device_source
            pa = np.array(list(params) + [0.0], dtype=np.float64)
            _lib.check(L.claw_set_source(h, int(src_id), _lib.dp(pa), len(params)), h)
        _lib.check(L.claw_set_bc(h, _lib.ip(_iarr(bc_lo, 3)), _lib.ip(_iarr(bc_up, 3)),
                                 _lib.ip(_iarr(solver.aux_bc_lower, 3)), _lib.ip(_iarr(solver.aux_bc_upper, 3)),
                                 _lib.ip(_iarr(solver.reflect_index, 3))), h)
        grid = state.grid
        lo = [1 if b in (True, None) else 0 for b in grid.on_lower_boundary] + [1] * (3 - nd)
        up = [1 if b in (True, None) else 0 for b in grid.on_upper_boundary] + [1] * (3 - nd)
        _lib.check(L.claw_set_boundary_ownership(h, _lib.ip(_iarr(lo)), _lib.ip(_iarr(up))), h)
        if solver.cuda_precision not in _MODES:
            raise ValueError("cuda_precision must be 'strict' or 'fast'")
        _lib.check(L.claw_set_mode(h, _MODES[solver.cuda_precision]), h)
        _lib.check(L.claw_set_tuning(h, int(solver.cuda_seg_rows)), h)
        self.delta = np.array(list(grid.delta) + [1.0] * (3 - nd), dtype=np.float64)
        # host <-> device coherence of state.q
        self.host_valid = True      # state._q equals the device's current buffer
        self.host_dirty = True      # state._q may have been changed on the host since the last upload
        self.aux_dirty = False      # state.aux was replaced on the host since the upload at setup
        self.peer_dirty = True      # (peer transport) the neighbours' ghost cells of the current buffer need a refresh
        self.halo = None            # multi-rank runs: petclaw.device_halo.DeviceHalo (ghost exchange + CFL all-reduce)
        self.halo_aux = None        # the same for the aux ghost planes (exchanged when aux is uploaded)
        # device-resident step control (claw_ctl_*)
        self.pending = False        # attempts are queued whose outcome the host has not read yet
        self.host_ctl_dirty = True  # the host's t / dt / cfl limits / status are newer than the device's
        self.ctl_started = False
        self.read_attempts = 0      # attempts whose log entries the host has read
        self.queued = 0             # attempts queued since the last read
        self._done = False
        self._words = np.zeros(32, dtype=np.float64)
        self._log = np.zeros(3 * LOG_CAP, dtype=np.float64)
        self.solver_ref = weakref.ref(solver)
        self.state_ref = weakref.ref(state)
        if state.num_aux > 0:
            if any(b == 0 for b in list(solver.aux_bc_lower) + list(solver.aux_bc_upper)):
                # BC.custom aux faces (e.g. edge velocities on the extended grid, examples/advection_2d_annulus):
                # the host callbacks run ONCE here, the filled array goes to the device, and the device-side ghost
                # fill leaves custom faces untouched (aux is static in the device-resident mode)
                solver._allocate_bc_arrays(state)
                auxbc = np.asfortranarray(solver.auxbc)
                _lib.check(L.claw_put_auxbc(h, auxbc.ctypes.data_as(ctypes.c_void_p)), h)
            else:
                aux = np.asfortranarray(state.aux)
                _lib.check(L.claw_put_aux(h, aux.ctypes.data_as(ctypes.c_void_p)), h)
        state._device = self

    # ---- coherence protocol used by State.q --------------------------------------------------
    def sync_host(self, state):
        if self.pending:
            self.drain()
        if not self.host_valid:
            q = state._q
            if not (q.flags['F_CONTIGUOUS'] and q.flags['WRITEABLE'] and q.flags['OWNDATA']):
                q = np.empty(q.shape, order='F')
                state._q = q
            _lib.check(self.L.claw_get_q(self.handle, q.ctypes.data_as(ctypes.c_void_p)), self.handle)
            self.host_valid = True
        self.host_dirty = True          # the caller may write through the returned array

    def host_modified(self):
        if self.pending:
            self.drain()
        self.host_valid = True
        self.host_dirty = True

    def _upload_if_needed(self, state):
        if self.host_dirty:
            if self.pending:
                self.drain()
            q = np.asfortranarray(state._q, dtype=np.float64)
            _lib.check(self.L.claw_put_q(self.handle, q.ctypes.data_as(ctypes.c_void_p)), self.handle)
            self.host_dirty = False
            self.peer_dirty = True          # peer transport: the neighbours' copies of my boundary rows are stale
        if self.aux_dirty:                                   # state.aux = new_array after setup
            if state.num_aux > 0:
                aux = np.asfortranarray(state.aux, dtype=np.float64)
                _lib.check(self.L.claw_put_aux(self.handle, aux.ctypes.data_as(ctypes.c_void_p)), self.handle)
                if self.halo_aux is not None:
                    self.halo_aux.exchange_aux(self)
            self.aux_dirty = False

    # ---- device-resident evolve_to_time (solver.py:554-666 without a host round trip per step) -------------------
    def _pack_host_words(self, solver, state):
        d = solver.__dict__
        w = np.zeros(32, dtype=np.float64)
        iw = w.view(np.int64)
        w[W_T] = state._t
        w[W_DT] = d['_m_dt']
        w[W_DTMAX] = d['_m_dt_max']
        w[W_CFLDES] = d['_m_cfl_desired']
        w[W_CFLMAX] = d['_m_cfl_max']
        w[W_CFL] = d['_m_cfl'].get_cached_max()
        st = d['_m_status']
        w[W_DTMIN], w[W_DTMAXS], w[W_CFLMAXS] = st['dtmin'], st['dtmax'], st['cflmax']
        iw[W_NUMSTEPS] = int(st['numsteps'])
        iw[W_DTVAR] = 1 if d['_m_dt_variable'] else 0
        iw[W_HAVEOLD] = 0 if d['_m_dt_old'] is None else 1
        return w

    def _attempt(self, n):
        """Queue n step attempts: [ghost exchange] + ghost fill + sweep + [CFL all-reduce] + control kernel."""
        L, h = self.L, self.handle
        if self.halo is None:
            _lib.check(L.claw_ctl_attempts(h, _lib.dp(self.delta), int(n), 1), h)
        elif self.halo.transport == 'peer':
            # ghost cells travel inside the attempt (peer-memory writes after the sweep, CFL through the mailboxes): the
            # N-rank attempt is queued like the 1-rank one.  Only after a host upload -- which every rank must do alike
            # -- are the neighbours' ghost cells refreshed by hand.
            if self.peer_dirty:
                self.halo.exchange(self)
                self.peer_dirty = False
            _lib.check(L.claw_ctl_attempts(h, _lib.dp(self.delta), int(n), 1), h)
        else:
            for _ in range(n):
                self.halo.exchange(self)
                _lib.check(L.claw_ctl_sweep(h, _lib.dp(self.delta), 1), h)
                self.halo.all_reduce_cfl(self)
                _lib.check(L.claw_ctl_end(h), h)
        self.queued += n
        self.pending = True

    def evolve(self, solver, solution, tend):
        state = solution.state
        take_one = tend is None
        self._upload_if_needed(state)
        d = solver.__dict__
        if not d['_m_dt_variable']:
            if self.pending:
                self.drain()
            if take_one:
                max_steps = 1
            else:
                tstart, dt = state._t, d['_m_dt']
                max_steps = int((tend - tstart + 1e-10) / dt)
                if abs(max_steps * dt - (tend - tstart)) > 1e-5 * (tend - tstart):
                    raise Exception('dt does not divide (tend-tstart) and dt is fixed!')
        else:
            max_steps = int(solver.max_steps)
            if d['_m_dt_variable'] == 1 and d['_m_cfl_desired'] > d['_m_cfl_max']:
                raise Exception('Variable time-stepping and desired CFL > maximum CFL')
        words = None
        if self.host_ctl_dirty or not self.ctl_started:
            if self.pending:
                self.drain()
            words = self._pack_host_words(solver, state)
        _lib.check(self.L.claw_ctl_begin(self.handle, None if words is None else words.ctypes.data_as(ctypes.c_void_p),
                                         float('inf') if take_one else float(tend), 1 if take_one else 0,
                                         max_steps), self.handle)
        self.host_ctl_dirty = False
        self.ctl_started = True
        self.pending = True
        gauges = len(state.grid.gauges) > 0
        if take_one:
            self._attempt(1)
            if gauges or self.queued >= LOG_CAP - BATCH:
                self._drain_with_gauges(solver, solution)
            return solver.__dict__['_m_status']
        while True:
            self.drain()
            if self._done:
                break
            if gauges:
                n = 1
            else:
                t, dt = state._t, d['_m_dt']
                left = (tend - t) / dt if dt > 0 else 1.0
                n = int(min(BATCH, max(1.0, math.ceil(left)))) if math.isfinite(left) else 1
            self._attempt(n)
            if gauges:
                self._drain_with_gauges(solver, solution)
        return solver.__dict__['_m_status']

    def _drain_with_gauges(self, solver, solution):
        n0 = solver.__dict__['_m_status']['numsteps']
        self.drain()
        if solver.__dict__['_m_status']['numsteps'] > n0:
            solver.write_gauge_values(solution)

    def drain(self):
        """Wait for the queued attempts and refresh the host's copies of t, dt, the CFL number, solver.status and the
        step log from the device's control block."""
        solver, state = self.solver_ref(), self.state_ref()
        n = ctypes.c_int(0)
        _lib.check(self.L.claw_ctl_read(self.handle, self._words.ctypes.data_as(ctypes.c_void_p), self.read_attempts,
                                        self._log.ctypes.data_as(ctypes.c_void_p), LOG_CAP, ctypes.byref(n)),
                   self.handle)
        self.pending = False
        self.queued = 0
        w, iw = self._words, self._words.view(np.int64)
        self.read_attempts = int(iw[W_ATTEMPTS])
        self._done = bool(iw[W_DONE])
        entries = [(float(self._log[3 * k]), float(self._log[3 * k + 1]), bool(self._log[3 * k + 2] != 0.0))
                   for k in range(n.value)]
        if any(e[2] for e in entries):
            self.host_valid = False
        if state is not None:
            state._t = float(w[W_T])
        if solver is not None:
            d = solver.__dict__
            d['_m_dt'] = float(w[W_DT])
            d['_m_cfl'].set_global_max(float(w[W_CFL]))
            st = d['_m_status']
            st['dtmin'], st['dtmax'], st['cflmax'] = float(w[W_DTMIN]), float(w[W_DTMAXS]), float(w[W_CFLMAXS])
            st['numsteps'] = int(iw[W_NUMSTEPS])
            if iw[W_HAVEOLD] and d['_m_dt_old'] is None:
                d['_m_dt_old'] = entries[0][0] if entries else float(w[W_DT])
            d['_m_step_log'].extend(entries)
            if entries:
                d['accept_step'] = entries[-1][2]
        err = int(iw[W_ERROR])
        if err:
            iw[W_ERROR] = 0
            self.host_ctl_dirty = True          # the next begin starts from the host's (consistent) values
            if err == 1:
                raise Exception('CFL too large, giving up!')
            if err == 3:
                raise _lib.ClawCudaError('a peer rank did not deliver its CFL number (peer-memory transport timed out)')
            raise Exception('Maximum number of timesteps have been taken')

    # ---- stepping ------------------------------------------------------------------------------
    def step(self, state, dt):
        """HOST-driven attempt (``solver.step`` / ``step_hyperbolic`` called directly): ghost fill + fused sweep, then
        the CFL number is read back; returns the global CFL number.  ``evolve_to_time`` does not come through here."""
        self._upload_if_needed(state)
        if self.pending:
            self.drain()
        self.host_ctl_dirty = True
        self.ctl_started = False
        if self.halo is not None:
            self.halo.exchange(self)
        _lib.check(self.L.claw_step_launch(self.handle, _lib.dp(self.delta), float(dt), 1), self.handle)
        if self.halo is not None:
            if self.halo.transport == 'peer':
                self.peer_dirty = True       # the next attempt starts from ghost cells refreshed by hand
                return self.halo.all_reduce_cfl(self)
            self.halo.all_reduce_cfl(self)
            cfl = ctypes.c_double(0.0)
            _lib.check(self.L.claw_cfl_read(self.handle, ctypes.byref(cfl)), self.handle)
            return cfl.value
        cfl = ctypes.c_double(0.0)
        _lib.check(self.L.claw_cfl_wait(self.handle, ctypes.byref(cfl)), self.handle)
        return cfl.value

    def get_cells(self, state, cells):
        """q[:, cell] for a list of interior cell index tuples (gauges) without downloading the array."""
        self._upload_if_needed(state)
        if self.pending:
            self.drain()
        idx = np.ascontiguousarray(cells, dtype=np.intc).reshape(len(cells), self.num_dim)
        out = np.empty((len(cells), state.num_eqn), dtype=np.float64)
        _lib.check(self.L.claw_get_cells(self.handle, len(cells), _lib.ip(idx), _lib.dp(out)), self.handle)
        return out

    def accept(self):
        _lib.check(self.L.claw_accept(self.handle), self.handle)
        self.host_valid = False
        self.host_ctl_dirty = True

    def reject(self):
        _lib.check(self.L.claw_reject(self.handle), self.handle)

    def launch_count(self):
        c = ctypes.c_uint64()
        _lib.check(self.L.claw_launch_count(self.handle, ctypes.byref(c)), self.handle)
        return int(c.value)

    def synchronize(self):
        _lib.check(self.L.claw_sync(self.handle), self.handle)

    def close(self, sync=True):
        """Bring the final state back to the host, detach from the State, destroy the handle."""
        if self.handle:
            state = self.state_ref()
            if state is not None and state.__dict__.get('_device') is self:
                try:
                    if sync:
                        self.sync_host(state)
                finally:
                    state._device = None
            self.L.claw_destroy(self.handle)
            self.handle = ctypes.c_void_p()
            self.pending = False

    def __del__(self):
        try:
            self.close(sync=False)
        except Exception:
            pass


class _CudaFmod(object):
    """f2py-shaped plug-in object.  Same positional arguments as the reference modules; the
    work/aux1..3 scratch arguments are accepted and ignored; ``rpn*`` must be a
    clawpack_b200.riemann solver object (it carries rp_id and cparam)."""

    def __init__(self, num_dim, precision='strict'):
        self.num_dim = num_dim
        self.mode = _MODES[precision]
        self.L = _lib.lib()
        if self.L.claw_device_count() == 0:
            raise _lib.ClawCudaError('the CUDA plug-in needs a CUDA device; none is visible (no CPU fallback)')

    @staticmethod
    def _prep(qold, qnew, aux, method, mthlim):
        qold = np.asfortranarray(qold, dtype=np.float64)
        if not (isinstance(qnew, np.ndarray) and qnew.flags['F_CONTIGUOUS'] and qnew.dtype == np.float64):
            qnew = np.asfortranarray(qnew, dtype=np.float64)
        aux = None if aux is None or aux.size == 0 else np.asfortranarray(aux, dtype=np.float64)
        return qold, qnew, aux, np.ascontiguousarray(method, dtype=np.intc), _iarr(mthlim)

    @staticmethod
    def _vp(a):
        return ctypes.c_void_p(0) if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def step1(self, num_ghost, mx, q, aux, dx, dt, method, mthlim, use_fwave, rp1):
        """q, cfl = step1(...) of the f2py module classic1 as called at classic/solver.py:302: q (with ghost cells)
        is updated in place and returned."""
        if not (isinstance(q, np.ndarray) and q.flags['F_CONTIGUOUS'] and q.dtype == np.float64):
            q = np.asfortranarray(q, dtype=np.float64)
        aux = None if aux is None or aux.size == 0 else np.asfortranarray(aux, dtype=np.float64)
        method = np.ascontiguousarray(method, dtype=np.intc)
        mth = _iarr(mthlim)
        cp, ncp = _cparam_of(rp1)
        cfl = ctypes.c_double(0.0)
        _lib.check(self.L.claw_step1_host(q.shape[0], len(mth), 0 if aux is None else aux.shape[0], num_ghost, mx,
                                          self._vp(q), self._vp(aux), dx, dt, _lib.ip(method), _lib.ip(mth),
                                          int(bool(use_fwave)), int(rp1.rp_id), _lib.dp(cp), ncp, self.mode,
                                          ctypes.byref(cfl)))
        return q, cfl.value

    def step2(self, maxm, num_ghost, mx, my, qold, qnew, aux, dx, dy, dt, method, mthlim, aux1, aux2, aux3, work,
              use_fwave, rpn2, rpt2):
        qold, qnew, aux, method, mth = self._prep(qold, qnew, aux, method, mthlim)
        cp, ncp = _cparam_of(rpn2)
        cfl = ctypes.c_double(0.0)
        _lib.check(self.L.claw_step2_host(maxm, qold.shape[0], len(mth), int(method[6]), num_ghost, mx, my,
                                          self._vp(qold), self._vp(qnew), self._vp(aux), dx, dy, dt, _lib.ip(method),
                                          _lib.ip(mth), int(bool(use_fwave)), int(rpn2.rp_id), _lib.dp(cp), ncp,
                                          self.mode, ctypes.byref(cfl)))
        return qnew, cfl.value

    def step2ds(self, maxm, num_ghost, mx, my, qold, qnew, aux, dx, dy, dt, method, mthlim, aux1, aux2, aux3, work,
                ids, use_fwave, rpn2, rpt2):
        qold, qnew, aux, method, mth = self._prep(qold, qnew, aux, method, mthlim)
        cp, ncp = _cparam_of(rpn2)
        cfl = ctypes.c_double(0.0)
        _lib.check(self.L.claw_step2ds_host(maxm, qold.shape[0], len(mth), int(method[6]), num_ghost, mx, my,
                                            self._vp(qold), self._vp(qnew), self._vp(aux), dx, dy, dt,
                                            _lib.ip(method), _lib.ip(mth), int(ids), int(bool(use_fwave)),
                                            int(rpn2.rp_id), _lib.dp(cp), ncp, self.mode, ctypes.byref(cfl)))
        return qnew, cfl.value

    def step3(self, maxm, num_ghost, mx, my, mz, qold, qnew, aux, dx, dy, dz, dt, method, mthlim, aux1, aux2, aux3,
              work, use_fwave, rpn3, rpt3, rptt3):
        qold, qnew, aux, method, mth = self._prep(qold, qnew, aux, method, mthlim)
        cp, ncp = _cparam_of(rpn3)
        cfl = ctypes.c_double(0.0)
        _lib.check(self.L.claw_step3_host(maxm, qold.shape[0], len(mth), int(method[6]), num_ghost, mx, my, mz,
                                          self._vp(qold), self._vp(qnew), self._vp(aux), dx, dy, dz, dt,
                                          _lib.ip(method), _lib.ip(mth), int(bool(use_fwave)), int(rpn3.rp_id),
                                          _lib.dp(cp), ncp, self.mode, ctypes.byref(cfl)))
        return qnew, cfl.value

    def step3ds(self, maxm, num_ghost, mx, my, mz, qold, qnew, aux, dx, dy, dz, dt, method, mthlim, aux1, aux2,
                aux3, work, idir, use_fwave, rpn3, rpt3, rptt3):
        qold, qnew, aux, method, mth = self._prep(qold, qnew, aux, method, mthlim)
        cp, ncp = _cparam_of(rpn3)
        cfl = ctypes.c_double(0.0)
        _lib.check(self.L.claw_step3ds_host(maxm, qold.shape[0], len(mth), int(method[6]), num_ghost, mx, my, mz,
                                            self._vp(qold), self._vp(qnew), self._vp(aux), dx, dy, dz, dt,
                                            _lib.ip(method), _lib.ip(mth), int(idir), int(bool(use_fwave)),
                                            int(rpn3.rp_id), _lib.dp(cp), ncp, self.mode, ctypes.byref(cfl)))
        return qnew, cfl.value


def fmod(num_dim, precision='strict'):
    return _CudaFmod(num_dim, precision)
