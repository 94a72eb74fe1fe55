"""CPU tests of the PyClaw-shaped host layer and of the C ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from clawpack_b200 import pyclaw, riemann, scenarios as S
from oracle_fmod import OracleFmod
from refutil import goldens, ref_solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_loop_with_plugin_matches_oracle_loop(oracle_lib):
    """kernel_language='Fortran' + an injected fmod (here the test-only oracle plug-in) drives exactly the
    reference's step sequence: identical q and identical (dt, cfl, accepted) logs."""
    for sc in (S.advection_2d(dimensional_split=False, transverse_waves=2),
               S.quadrants(mx=40, my=40, tfinal=0.2), S.radial_dam_break(mx=40, my=40, tfinal=0.5),
               S.kpp_2d(mx=30, my=30, tfinal=0.3, num_output_times=3),
               S.elasticity_2d(mx=24, my=20, tfinal=0.1, num_output_times=2),
               S.advection_reaction_2d(mx=24, my=20, tfinal=1.0, num_output_times=2),
               S.elasticity_2d(mx=20, my=24, tfinal=0.1, num_output_times=2, dimensional_split=True)):
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert np.array_equal(claw.frames[-1].q, r.q)
        assert claw.solver.step_log == r.step_log
        assert any(not s[2] for s in r.step_log) or sc['rp'] in ('shallow_roe_with_efix_2D', 'kpp_2D', 'vc_elasticity_2D')    # (advection_reaction: 1 reject)


def test_host_loop_reference_golden(oracle_lib):
    sc = S.quadrants()
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw.run()
    rho = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    dx, dy = claw.solution.domain.grid.delta
    assert dx * dy * np.sum(np.abs(rho - goldens()['euler_2d__quadrants_roe'])) < 1e-6


def test_host_loop_3d_with_plugin_matches_oracle_loop(oracle_lib):
    """ClawSolver3D (classic/solver.py:648-706) with the oracle plug-in: unsplit (step3) with aux arrays and
    wall/periodic BCs, dimension-split (3 aliased step3ds calls), and BC.custom + wall on euler_3D."""
    G = goldens()
    cases = [S.acoustics_3d(mx=12, my=10, mz=9, tfinal=0.3, num_output_times=2),
             S.acoustics_3d(mx=16, my=4, mz=4, heterogeneous=False, tfinal=0.3, num_output_times=2),
             S.euler_3d_shocktube(mx=4, my=5, mz=24, tfinal=0.05, num_output_times=1),
             S.euler_3d_shock_bubble(G['euler_3d__shock_bubble_q0'][:, :20, :6, :6].copy(order='F'),
                                     G['euler_3d__shock_bubble_consts'], tfinal=0.02, num_output_times=1)]
    for sc in cases:
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert np.array_equal(claw.frames[-1].q, r.q)
        assert claw.solver.step_log == r.step_log


def test_host_loop_3d_reference_golden(oracle_lib):
    sc = S.acoustics_3d(mx=15, my=15, mz=15)
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw.run()
    p = claw.frames[-1].q[0].reshape(-1)
    assert np.max(np.abs(p - goldens()['acoustics_3d__verify_classic_heterogeneous'])) < 1e-13


def test_api_surface():
    solver = pyclaw.ClawSolver2D(riemann.euler_4wave_2D)
    assert (solver.num_eqn, solver.num_waves, solver.num_ghost) == (4, 4, 2)
    assert solver.kernel_language == 'CUDA' and solver.dimensional_split is True
    assert solver.transverse_waves == pyclaw.ClawSolver2D.trans_inc == 1
    assert solver.limiters == pyclaw.limiters.tvd.minmod == 1 and pyclaw.limiters.tvd.MC == 4
    assert (pyclaw.BC.custom, pyclaw.BC.extrap, pyclaw.BC.periodic, pyclaw.BC.wall) == (0, 1, 2, 3)
    with pytest.raises(TypeError):          # attribute set is frozen (solver.py:139-144)
        solver.no_such_option = 1
    solver.all_bcs = pyclaw.BC.extrap
    assert solver.bc_lower == [1, 1] and solver.bc_upper == [1, 1]
    s3 = pyclaw.ClawSolver3D(riemann.vc_acoustics_3D)
    assert s3.transverse_waves == 22 and s3.num_eqn == 4 and s3.num_waves == 2
    dom = pyclaw.Domain([0., 0.], [1., 1.], [4, 3])
    sol = pyclaw.Solution(4, dom)
    assert sol.q.shape == (4, 4, 3) and sol.q.flags['F_CONTIGUOUS']
    assert np.allclose(dom.grid.p_centers[0][:, 0], [0.125, 0.375, 0.625, 0.875])
    st = pyclaw.State(dom, 3)
    st.problem_data['grav'] = 1.0
    st.set_cparam(riemann.shallow_roe_with_efix_2D)
    assert riemann.shallow_roe_with_efix_2D.cparam.grav == 1.0
    with pytest.raises(Exception):          # missing cparam member raises (state.py:212-215)
        pyclaw.State(dom, 4).set_cparam(riemann.euler_4wave_2D)


def test_cuda_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from clawpack_b200._lib import ClawCudaError
    claw = S.make_claw(S.advection_2d(), kernel_language='CUDA')
    with pytest.raises(ClawCudaError):
        claw.run()


def test_fortran_path_needs_reference_module():
    claw = S.make_claw(S.advection_2d(), kernel_language='Fortran')
    with pytest.raises(ImportError):
        claw.run()


def test_abi_exports_every_declared_symbol():
    """libclawcuda.so loads on a CPU-only box and exports every function include/clawcuda.h declares."""
    from clawpack_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, 'include', 'clawcuda.h')).read()
    declared = set(re.findall(r'\b(claw_[a-z0-9_]+)\s*\(', header))
    declared.discard('claw_solver')
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(L, name), 'missing export ' + name
    assert set(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ declared
    assert b'clawcuda' in L.claw_version()
    assert L.claw_device_count() >= 0
    # error paths never abort: bad arguments give a code + message
    h = ctypes.c_void_p()
    n = (ctypes.c_int * 3)(8, 8, 1)
    rc = L.claw_create(ctypes.byref(h), 4, 3, 4, 4, 0, n, 2, 0, None)
    assert rc != 0 and b'ndim' in L.claw_last_error(None)
    rc = L.claw_create(ctypes.byref(h), 2, 3, 3, 4, 0, n, 2, 0, None)
    assert rc != 0 and b'expects' in L.claw_last_error(None)


def test_slab_partition():
    from clawpack_b200.petclaw.util import slab_range
    for n, p in ((8192, 8), (100, 3), (17, 4)):
        parts = [slab_range(n, p, r) for r in range(p)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_gauges_host_loop_matches_oracle(oracle_lib, tmp_path):
    """Solver.write_gauge_values / Grid.add_gauges / setup_gauge_files (solver.py:680-715, geometry.py:392-423) in the
    PyClaw-shaped host loop: same records in memory and in the gauge files as the oracle loop, initial record included."""
    sc = S.psystem_2d(cells_per_layer=3, tfinal=0.4, num_output_times=2, size=8.0)
    sc['gauges'] = [[0.25, 0.25], [3.3, 1.25], [100.0, 100.0]]            # the last one is outside the grid: ignored
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw.outdir = str(tmp_path)
    claw.output_format = None
    claw.run()
    r = ref_solver(sc)
    r.run(sc['tfinal'], sc['num_output_times'])
    state = claw.solution.state
    assert state.grid.gauges == [[0, 0], [9, 3]] == r.gauges
    assert len(state.gauge_data) == 2
    for i in range(2):
        assert np.array_equal(state.gauge_data[i], r.gauge_data[i])
        assert state.gauge_data[i].shape == (len([s_ for s_ in r.step_log if s_[2]]) + 1, 3)
    assert claw.solver.step_log == r.step_log


def _golden_small_solutions():
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden3', os.path.join(ROOT, 'tests', 'golden', 'make_golden3.py'))
    # only small_solutions() is needed, and the module body imports the reference: re-state the loader guard here
    src = open(spec.origin).read().split('def small_solutions')[1].split("if __name__")[0]
    ns = {'pyclaw': pyclaw, 'np': np}
    exec('def small_solutions' + src, ns)
    return ns['small_solutions']()


def test_ascii_frames_match_reference_writer(tmp_path):
    """Solution.write(file_format='ascii') produces byte-for-byte the files the reference's own writer
    (pyclaw/src/pyclaw/fileio/ascii.py:17-126) produced for the same solutions (tests/golden/ascii_frames, written by
    tests/golden/make_golden3.py), and Solution.read gives back the stored (8-digit) values."""
    gdir = os.path.join(ROOT, 'tests', 'golden', 'ascii_frames')
    for name, sol, frame in _golden_small_solutions():
        out = tmp_path / name
        sol.write(frame, str(out), 'ascii', write_aux=True)
        for fname in sorted(os.listdir(os.path.join(gdir, name))):
            assert open(out / fname, 'rb').read() == open(os.path.join(gdir, name, fname), 'rb').read(), fname
        back = pyclaw.Solution(frame, path=os.path.join(gdir, name), file_format='ascii', read_aux=True)
        assert back.t == float('%18.8e' % sol.t)
        assert back.state.q.shape == sol.state.q.shape and back.state.q.flags['F_CONTIGUOUS']
        rounded = np.vectorize(lambda v: float('%18.8e' % v))(sol.state.q)
        assert np.array_equal(back.state.q, rounded)
        assert [d.num_cells for d in back.domain.patch.dimensions] == [d.num_cells for d in sol.domain.patch.dimensions]
        assert np.allclose(back.domain.patch.lower_global, sol.domain.patch.lower_global)
        if sol.num_aux:
            assert np.array_equal(back.state.aux, np.vectorize(lambda v: float('%18.8e' % v))(sol.state.aux))
        # format auto-detection from the .t header (solution.py:336-345)
        assert np.array_equal(pyclaw.Solution(frame, path=str(out)).state.q, back.state.q)


def test_controller_writes_frames(oracle_lib, tmp_path):
    """Controller.run with the reference's default output_format='ascii' (controller.py:141, 366-381): one set of frame
    files per output time, aux only for frame 0 with write_aux_init; the last frame reads back to the final state."""
    sc = S.acoustics_2d(mx=12, my=10, tfinal=0.04, num_output_times=2)
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    assert pyclaw.Controller().output_format == 'ascii'
    claw.output_format = 'ascii'
    claw.outdir = str(tmp_path / 'out')
    claw.run()
    names = sorted(os.listdir(claw.outdir))
    assert [n for n in names if n.startswith('fort.q')] == ['fort.q0000', 'fort.q0001', 'fort.q0002']
    assert [n for n in names if n.startswith('fort.t')] == ['fort.t0000', 'fort.t0001', 'fort.t0002']
    back = pyclaw.Solution(2, path=claw.outdir)
    assert abs(back.t - sc['tfinal']) < 1e-9
    assert np.max(np.abs(back.state.q - claw.solution.state.q)) <= 5e-9 * max(1.0, np.max(np.abs(claw.solution.state.q)))
    with pytest.raises(ValueError):
        claw.solution.write(0, claw.outdir, 'hdf5')


def test_host_loop_1d_with_plugin_matches_oracle_loop(oracle_lib):
    """ClawSolver1D (classic/solver.py:249-303, Fortran branch) with the oracle plug-in: step1 updates qbc in place."""
    for sc in (S.burgers_1d(mx=60, tfinal=0.2), S.dam_break_1d(mx=50, tfinal=0.5), S.acoustics_1d(num_cells=40, tfinal=0.3),
               S.advection_1d(nx=30, tfinal=0.2), S.dam_break_1d(mx=50, tfinal=0.5, riemann_solver='hlle')):
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert np.array_equal(claw.solution.state.q, r.q), sc['rp']
        assert claw.solver.step_log == r.step_log


def test_host_loop_1d_aux_fwave_custom_bc_with_plugin(oracle_lib):
    """ClawSolver1D with solvers that read aux (periodic / extrapolated aux ghost cells), f-wave solvers, and the
    stegoton example's moving-wall ``BC.custom`` (reads t and aux, leaves one ghost cell stale): the host mirror with
    the oracle plug-in against the oracle's own loop, and -- end to end through Controller.run() -- against the
    reference's stored vectors (traffic/test_traffic.py:21-22, advection_1d_variable/...:35-36,
    shallow_1d/test_shallow1d.py:41-44; L1 < 1e-6)."""
    from refutil import goldens4
    g4 = goldens4()
    for sc, key in ((S.traffic_1d(), 'traffic__classic'),
                    (S.advection_1d_variable(), 'advection_1d_variable__fortran_classic'),
                    (S.sill_1d(), 'shallow_1d__sill_Fortran'),
                    (S.stegoton_1d(xupper=40.0, tfinal=25.0, num_output_times=5), None),
                    (S.stegoton_1d(xupper=30.0, tfinal=22.0, num_output_times=4, exponential=False), None),
                    (S.shocktube_1d(mx=100), None), (S.blast_1d(mx=120, tfinal=0.01), None),
                    (S.cubic_1d(mx=80, tfinal=0.05), None), (S.acoustics_1d_variable(mx=90, tfinal=0.4), None),
                    (S.traffic_vc_1d(mx=60, tfinal=0.3), None), (S.traffic_vc_1d(mx=60, tfinal=0.3, kind='fwave'), None),
                    (S.traffic_vc_1d(mx=60, tfinal=0.3, kind='tracer'), None)):
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert np.array_equal(claw.solution.state.q, r.q), sc['rp']
        assert claw.solver.step_log == r.step_log
        if key is not None:
            q = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
            assert claw.solution.domain.grid.delta[0] * np.sum(np.abs(q - g4[key])) < 1e-6, key


def _golden_binary_solutions():
    src = open(os.path.join(ROOT, 'tests', 'golden', 'make_golden5.py')).read()
    src = src.split('def small_solutions')[1].split("if __name__")[0]
    ns = {'pyclaw': pyclaw, 'np': np}
    exec('def small_solutions' + src, ns)
    return ns['small_solutions']()


def test_binary_frames_are_what_the_reference_reader_parses(tmp_path):
    """Solution.write(file_format='binary' | 'binary32') reproduces the committed files (tests/golden/binary_frames,
    written by this package and read back there by the REFERENCE's own reader, pyclaw/src/pyclaw/fileio/binary.py:19-208,
    see tests/golden/make_golden5.py), and Solution.read returns exactly what the reference reader parsed."""
    gdir = os.path.join(ROOT, 'tests', 'golden', 'binary_frames')
    ref = np.load(os.path.join(gdir, 'reference_parsed.npz'))
    for name, sol, frame, fmt in _golden_binary_solutions():
        out = tmp_path / name
        sol.write(frame, str(out), fmt, write_aux=True)
        for fname in sorted(os.listdir(os.path.join(gdir, name))):
            assert open(out / fname, 'rb').read() == open(os.path.join(gdir, name, fname), 'rb').read(), (name, fname)
        back = pyclaw.Solution(frame, path=os.path.join(gdir, name), file_format=fmt, read_aux=True)
        assert back.state.q.flags['F_CONTIGUOUS'] and back.state.q.dtype == np.float64
        assert np.array_equal(back.state.q, ref[name + '__q'])
        assert back.t == float(ref[name + '__t'])
        assert [d.num_cells for d in back.domain.patch.dimensions] == list(ref[name + '__num_cells'])
        assert np.array_equal([d.lower for d in back.domain.patch.dimensions], ref[name + '__lower'])
        assert np.array_equal([d.delta for d in back.domain.patch.dimensions], ref[name + '__delta'])
        if sol.num_aux:
            # the reference reader slices aux with [mbc:-mbc] even when mbc == 0 (fileio/binary.py:195-199; its q branch
            # has the mbc == 0 case, :104-107) and so hands back an EMPTY aux for ghost-free frames; this reader
            # returns the stored values
            assert ref[name + '__aux'].size == 0
            want = sol.state.aux if fmt == 'binary64' else sol.state.aux.astype(np.float32).astype(np.float64)
            assert np.array_equal(back.state.aux, want)
        if fmt == 'binary64':
            assert np.array_equal(back.state.q, sol.state.q)          # lossless, unlike the 8-digit ascii format
        # format (and precision) auto-detected from the .t header
        assert np.array_equal(pyclaw.Solution(frame, path=str(out)).state.q, back.state.q)


def test_binary_reader_strips_ghost_cells(tmp_path):
    """AMRClaw writes its binary frames WITH ghost cells and says so in the .t header; the reader drops them
    (fileio/binary.py:104-121)."""
    dom = pyclaw.Domain([pyclaw.Dimension(0.0, 1.0, 5, name='x'), pyclaw.Dimension(0.0, 1.0, 4, name='y')])
    st = pyclaw.State(dom, 2, 0)
    st.q[...] = np.arange(40.0).reshape((2, 5, 4), order='F')
    sol = pyclaw.Solution(st, dom)
    sol.write(1, str(tmp_path), 'binary')
    g = 2
    big = np.full((2, 5 + 2 * g, 4 + 2 * g), -7.0, order='F')
    big[:, g:-g, g:-g] = st.q
    big.ravel(order='F').tofile(str(tmp_path / 'fort.b0001'))
    t = open(tmp_path / 'fort.t0001').read().replace('    0                  num_ghost', '    2                  num_ghost')
    open(tmp_path / 'fort.t0001', 'w').write(t)
    back = pyclaw.Solution(1, path=str(tmp_path), file_format='binary')
    assert np.array_equal(back.state.q, st.q)


def test_controller_binary_output_restart_functionals_and_overwrite_guard(oracle_lib, tmp_path):
    """Controller.run with output_format='binary': lossless frames, a run restarted from a frame file continues
    bit-identically; compute_p / compute_F products (controller.py:334-342, 399-409); overwrite=False refuses an
    existing outdir (controller.py:331-333)."""
    sc = S.acoustics_2d(mx=14, my=10, tfinal=0.06, num_output_times=3)
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw.output_format = 'binary'
    claw.outdir = str(tmp_path / 'out')

    def compute_p(state):
        state.p[0] = state.q[0] ** 2

    def compute_F(state):
        state.F[0] = state.q[0] * state.grid.delta[0] * state.grid.delta[1]

    claw.solution.state.mp = 1
    claw.solution.state.mF = 1
    claw.compute_p = compute_p
    claw.compute_F = compute_F
    claw.run()
    names = sorted(os.listdir(claw.outdir))
    assert [n for n in names if n.startswith('fort.b')] == ['fort.b000%d' % k for k in range(4)]
    back = pyclaw.Solution(3, path=claw.outdir)
    assert back.t == float('%18.8e' % claw.solution.t)
    assert np.array_equal(back.state.q, claw.solution.state.q)
    pback = pyclaw.Solution(3, path=claw.outdir_p, file_prefix='claw_p')
    assert pback.state.q.shape[0] == 1 and np.array_equal(pback.state.q[0], claw.solution.state.q[0] ** 2)
    rows = [list(map(float, ln.split())) for ln in open(claw.F_path)]
    assert len(rows) == 4 and rows[0][0] == 0.0
    assert abs(rows[-1][1] - np.sum(np.abs(claw.solution.state.q[0])) * np.prod(claw.solution.state.grid.delta)) < 1e-14
    # restart from frame 1 and run to the same tfinal: same final state as the uninterrupted run
    mid = pyclaw.Solution(1, path=claw.outdir)
    claw2 = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw2.solution.state.q[...] = mid.state.q
    claw2.solution.state.t = claw.frames[1].t                      # the .t header keeps 8 digits; use the exact time
    claw2.num_output_times = 2
    claw2.solver.dt_initial = [s for s in claw.solver.step_log][0][0]
    claw2.run()
    assert abs(claw2.solution.t - sc['tfinal']) < 1e-12
    assert np.max(np.abs(claw2.solution.state.q - claw.solution.state.q)) < 1e-12
    claw3 = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw3.output_format = 'binary'
    claw3.outdir = claw.outdir
    claw3.overwrite = False
    with pytest.raises(Exception, match='Refusing to overwrite'):
        claw3.run()


def test_strang_source_splitting_fixed_dt_and_max_steps(oracle_lib):
    """ClawSolver.step options of classic/solver.py:96-146 and Solver.evolve_to_time (solver.py:554-666) that no
    reference example on the classic path exercises together: Strang splitting of a source term (source_split = 2:
    half step before and after the hyperbolic step), fixed time steps (dt_variable = False: t = tstart + n dt, 'dt does
    not divide' and 'CFL too large, giving up!' errors) and the max_steps guard -- host mirror with the oracle plug-in
    against the oracle's own loop, step for step."""
    def damping(q, aux, dt):                       # q_t = -2 q: exact update, so Godunov and Strang splitting differ
        q *= np.exp(-2.0 * dt)

    base = S.acoustics_2d(mx=20, my=16, tfinal=0.06, num_output_times=2)
    for split in (1, 2):
        sc = dict(base, step_source=damping, source_split=split)
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert claw.solver.source_split == split
        assert np.array_equal(claw.solution.state.q, r.q), split
        assert claw.solver.step_log == r.step_log
    # fixed dt: 12 steps of 0.005
    claw = S.make_claw(base, kernel_language='Fortran', fmod=OracleFmod(), dt_variable=False, dt_initial=0.005)
    claw.run()
    r = ref_solver(dict(base, dt_initial=0.005))
    r.dt_variable = False
    r._method[0] = 0
    r.run(base['tfinal'], base['num_output_times'])
    # (get_dt still trims the last step of an output interval to land on tend exactly, solver.py:547-552)
    assert len(claw.solver.step_log) == 12 and np.allclose([s[0] for s in claw.solver.step_log], 0.005, rtol=1e-12)
    assert claw.solver.step_log == r.step_log
    assert np.array_equal(claw.solution.state.q, r.q)
    assert abs(claw.solution.t - 0.06) < 1e-15
    claw = S.make_claw(base, kernel_language='Fortran', fmod=OracleFmod(), dt_variable=False, dt_initial=0.0047)
    with pytest.raises(Exception, match='dt does not divide'):
        claw.run()
    claw = S.make_claw(base, kernel_language='Fortran', fmod=OracleFmod(), dt_variable=False, dt_initial=0.03)
    with pytest.raises(Exception, match='CFL too large'):
        claw.run()
    claw = S.make_claw(base, kernel_language='Fortran', fmod=OracleFmod(), max_steps=1)
    with pytest.raises(Exception, match='Maximum number of timesteps'):
        claw.run()


def test_reads_the_reference_amrclaw_binary_frame(tmp_path):
    """pyclaw/src/pyclaw/tests/test_io.py:24-27 (test_io_from_binary): the reference's own regression frame -- written by
    AMRClaw, 17 patches on three levels, two ghost cells per side in the .b file, old-style .t header without a format
    line -- is read to exactly what the reference reader parses (tests/golden/make_golden6.py), then written and read
    back as ascii (8 digits) and as binary (lossless), patch by patch."""
    gdir = os.path.join(ROOT, 'tests', 'golden', 'amrclaw_binary_frame')
    ref = np.load(os.path.join(gdir, 'reference_parsed.npz'))
    sol = pyclaw.Solution(0, path=gdir, file_format='binary')
    assert len(sol.states) == int(ref['nstates']) == 17 and sol.t == float(ref['t'])
    for k, st in enumerate(sol.states):
        meta = ref['meta%02d' % k]
        dims = st.patch.dimensions
        assert [st.patch.patch_index, st.patch.level] == [int(meta[0]), int(meta[1])]
        assert [d.num_cells for d in dims] == [int(meta[2]), int(meta[3])]
        assert np.array_equal([d.lower for d in dims], meta[4:6]) and np.array_equal([d.delta for d in dims], meta[6:8])
        assert np.array_equal(st.q, ref['q%02d' % k])
    for fmt in ('ascii', 'binary'):
        out = str(tmp_path / fmt)
        sol.write(0, out, fmt)
        back = pyclaw.Solution(0, path=out, file_format=fmt)
        assert len(back.states) == 17
        for a, b in zip(back.states, sol.states):
            assert a.patch.level == b.patch.level and a.q.shape == b.q.shape
            if fmt == 'binary':
                assert np.array_equal(a.q, b.q)
            else:
                assert np.max(np.abs(a.q - b.q)) <= 5e-9 * max(1.0, np.max(np.abs(b.q)))


def test_device_callback_objects_are_reference_shaped_host_callbacks(oracle_lib):
    """pyclaw.device_callbacks: EulerRadialSource / FixedStateBC called as ordinary host callbacks (step_source(solver,
    state, dt); user_bc_lower(state, dim, t, qbc, auxbc, num_ghost)) give the bits of the example's own Python functions
    (shock_bubble_interaction.py:93-143) in the host loop -- and are recognised as device functors by the eligibility
    test of the device-resident path."""
    g2 = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens2.npz'))
    q0 = g2['euler_2d__shock_bubble_interaction_q0'][:, :64, :]
    sc = S.shock_bubble_interaction(q0, tfinal=0.02, device_callbacks=True)
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    assert hasattr(claw.solver.step_source, 'device_source') and hasattr(claw.solver.user_bc_lower, 'device_bc')
    claw.solver.kernel_language = 'CUDA'
    try:
        claw.solver.fmod = None
        assert claw.solver._device_path_eligible()
        claw.solver.source_split = 2                      # Strang splitting: host path
        assert not claw.solver._device_path_eligible()
        claw.solver.source_split = 1
    finally:
        claw.solver.kernel_language = 'Fortran'
        claw.solver.fmod = OracleFmod()
    claw.run()
    r = ref_solver(S.shock_bubble_interaction(q0, tfinal=0.02))
    r.run(sc['tfinal'], sc['num_output_times'])
    assert claw.solver.step_log == r.step_log
    assert np.array_equal(claw.frames[-1].q, r.q)
