"""The DEVICE functor source of the 1-D Riemann solvers (clawpack_b200/csrc/rp1.cuh), compiled by the host compiler
with the strict mode's unfused IEEE arithmetic (tests/host_functors/rp1_host.cpp), against the oracle's rp1 routines
on the same pencils: bit for bit.  This checks the arithmetic the GPU will execute in strict mode on a machine without
a GPU; the kernels around the functors (sweep1.cuh) are covered by the `-m gpu` parity tests.

Every check also runs against the FAST-mode source (namespace `fast`, -DCLAW_HOST_FAST, -ffp-contract=fast -mfma): the
restructured divisions / square roots / limiter algebra of the benchmarked mode must stay within 5e-13 of the oracle
relative to the largest entry of each output array (the device's approximate reciprocal seeds + one correction step are
replaced by exact 1/x on the host, so this checks the algebra, not the seeds)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from refutil import pyclaw_ref

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_functors', 'rp1_host.cpp')
CXX = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'


def _build(tmp_path_factory, src, tag, mode):
    out = str(tmp_path_factory.mktemp(tag + mode) / ('lib%s_%s.so' % (tag, mode)))
    flags = ['-ffp-contract=off'] if mode == 'strict' else ['-ffp-contract=fast', '-DCLAW_HOST_FAST'] + (
        ['-mfma'] if 'fma' in open('/proc/cpuinfo').read() else [])
    subprocess.run([CXX, '-O2'] + flags + ['-std=c++17', '-fPIC', '-shared', '-Wno-unknown-pragmas', '-o', out, src],
                   check=True)
    return ctypes.CDLL(out)


def _same(a, b, mode):
    """strict: equal bits; fast: within 5e-13 of the largest entry."""
    if mode == 'strict':
        return np.array_equal(a, b)
    return float(np.max(np.abs(a - b))) <= 5e-13 * max(float(np.max(np.abs(b))), 1e-300)


@pytest.fixture(scope='module', params=['strict', 'fast'])
def hostlib(tmp_path_factory, request):
    L = _build(tmp_path_factory, SRC, 'rp1', request.param)
    L.mode = request.param
    dp = ctypes.POINTER(ctypes.c_double)
    L.hostfun_rp1.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [dp] * 6
    L.hostfun_rp1.restype = ctypes.c_int
    R = pyclaw_ref.lib()
    R.clawref_rp1.argtypes = [ctypes.c_int] * 5 + [dp] * 6 + [ctypes.c_int, dp]
    R.clawref_rp1.restype = ctypes.c_int
    return L, R


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _pencils(name, rng, n):
    """Random states of a physically admissible kind for each solver, mixing smooth data, jumps, sign changes and
    (where the solver has one) states that trigger the entropy fix / dry-state branches."""
    aux = None
    cp = []
    if name in ('advection_1D',):
        q = rng.standard_normal((1, n)); cp = [rng.choice([-0.7, 1.3])]
    elif name == 'acoustics_1D':
        q = rng.standard_normal((2, n)); cp = [1.3, 2.9, np.sqrt(2.9 / 1.3), 1.3 * np.sqrt(2.9 / 1.3)]
    elif name in ('burgers_1D', 'cubic_1D'):
        q = rng.standard_normal((1, n)) * 2
    elif name == 'traffic_1D':
        q = rng.uniform(0, 1, (1, n)); cp = [rng.uniform(0.5, 2.0)]
    elif name in ('traffic_vc_1D', 'traffic_vc_fwave_1D', 'traffic_vc_tracer_1D'):
        rows = [rng.uniform(0, 1, n)] + ([rng.standard_normal(n)] if name == 'traffic_vc_tracer_1D' else [])
        q = np.stack(rows)
        q[0, ::9] = 0.5; q[0, 1::9] = 0.5                     # equal states at the flux maximum: the tie branches
        aux = rng.choice([0.5, 1.0, 1.0, 2.0], (1, n))
    elif name in ('shallow_roe_with_efix_1D', 'shallow_hlle_1D'):
        h = rng.uniform(0.05, 3.0, n); u = rng.standard_normal(n) * 2
        q = np.stack([h, h * u]); cp = [9.81]
    elif name == 'advection_color_1D':
        q = rng.standard_normal((1, n)); aux = rng.standard_normal((1, n))
    elif name == 'acoustics_variable_1D':
        q = rng.standard_normal((2, n)); aux = rng.uniform(0.5, 4.0, (2, n))
    elif name == 'shallow_bathymetry_fwave_1D':
        h = rng.uniform(0.0, 2.0, n); h[rng.random(n) < 0.2] = 5e-4          # some cells below dry_tolerance
        u = rng.standard_normal(n)
        q = np.stack([h, h * u]); aux = rng.uniform(-1.0, 0.0, (1, n)); cp = [9.8, 1e-3, 0.0]
    elif name in ('euler_with_efix_1D', 'euler_hlle_1D'):
        rho = rng.uniform(0.1, 3.0, n); u = rng.standard_normal(n) * 2; p = rng.uniform(0.1, 5.0, n)
        if name == 'euler_hlle_1D':
            # rp1_euler_hlle.f90:62-64 forms H_r with p_l and c^2 as gamma1*H - u^2/2 (= 1.4 p/rho - 0.3 u^2): keep the
            # square roots real with mild pressure jumps and |u| < 1.5
            rho = rng.uniform(0.8, 1.2, n); p = rng.uniform(1.0, 1.3, n); u = rng.uniform(-1.5, 1.5, n)
        q = np.stack([rho, rho * u, p / 0.4 + 0.5 * rho * u * u]); cp = [1.4]
    elif name == 'nonlinear_elasticity_fwave_1D':
        q = np.stack([rng.uniform(-0.2, 0.5, n), rng.standard_normal(n)])
        aux = np.stack([rng.uniform(1.0, 4.0, n), np.zeros(n), rng.uniform(1.0, 4.0, n)])     # linear relation: no exp()
    else:
        raise KeyError(name)
    # smooth stretches and exact repeats (zero-strength waves) as well
    q[:, n // 2: n // 2 + 8] = q[:, n // 2: n // 2 + 1]
    return np.asfortranarray(q), (None if aux is None else np.asfortranarray(aux)), np.array(cp + [0.0])


NAMES_1D = sorted(k for k in pyclaw_ref.RP_IDS if k.endswith('_1D'))


@pytest.mark.parametrize('name', NAMES_1D)
def test_device_functor_source_matches_oracle_rp1(hostlib, name):
    L, R = hostlib
    rp_id = pyclaw_ref.RP_IDS[name]
    M, W = pyclaw_ref.NUM_EQN[name], pyclaw_ref.NUM_WAVES[name]
    rng = np.random.default_rng(1234 + rp_id)
    mx, g = 4093, 2
    n = mx + 2 * g
    q, aux, cp = _pencils(name, rng, n)
    maux = 0 if aux is None else aux.shape[0]
    auxo = np.zeros((1, n), order='F') if aux is None else aux
    wo, so = np.zeros((M, W, n), order='F'), np.zeros((W, n), order='F')
    amo, apo = np.zeros((M, n), order='F'), np.zeros((M, n), order='F')
    assert R.clawref_rp1(M, W, g, maux, mx, _dp(q), _dp(auxo), _dp(wo), _dp(so), _dp(amo), _dp(apo), rp_id, _dp(cp)) == 0
    # host-compiled device functor: interface k between cells k and k+1  <->  oracle slot k+1
    wd, sd = np.zeros((M, W, n - 1), order='F'), np.zeros((W, n - 1), order='F')
    amd, apd = np.zeros((M, n - 1), order='F'), np.zeros((M, n - 1), order='F')
    assert L.hostfun_rp1(rp_id, _dp(cp), len(cp) - 1, n - 1, maux, _dp(q), _dp(auxo), _dp(wd), _dp(sd), _dp(amd),
                         _dp(apd)) == 0
    for a, b, what in ((wd, wo[..., 1:], 'wave'), (sd, so[:, 1:], 's'), (amd, amo[:, 1:], 'amdq'), (apd, apo[:, 1:], 'apdq')):
        assert np.isfinite(b).all(), (name, what)
        assert _same(a, b, L.mode), (name, L.mode, what, float(np.max(np.abs(a - b))))


def test_nonlinear_elasticity_exponential_relation_within_roundoff(hostlib):
    """coefA > 0: sigma = exp(coefA*eps) - 1.  Same library on both sides here (so equal bits); on the GPU exp() may
    differ in the last bit, which is why the GPU parity tests hold this material to the 1e-12 tolerance."""
    L, R = hostlib
    name = 'nonlinear_elasticity_fwave_1D'
    rp_id = pyclaw_ref.RP_IDS[name]
    rng = np.random.default_rng(7)
    mx, g = 509, 2
    n = mx + 2 * g
    q = np.asfortranarray(np.stack([rng.uniform(-0.2, 0.5, n), rng.standard_normal(n)]))
    aux = np.asfortranarray(np.stack([rng.uniform(1.0, 4.0, n), rng.uniform(1.0, 4.0, n), np.zeros(n)]))
    cp = np.zeros(1)
    wo, so = np.zeros((2, 2, n), order='F'), np.zeros((2, n), order='F')
    amo, apo = np.zeros((2, n), order='F'), np.zeros((2, n), order='F')
    assert R.clawref_rp1(2, 2, g, 3, mx, _dp(q), _dp(aux), _dp(wo), _dp(so), _dp(amo), _dp(apo), rp_id, _dp(cp)) == 0
    wd, sd = np.zeros((2, 2, n - 1), order='F'), np.zeros((2, n - 1), order='F')
    amd, apd = np.zeros((2, n - 1), order='F'), np.zeros((2, n - 1), order='F')
    assert L.hostfun_rp1(rp_id, _dp(cp), 0, n - 1, 3, _dp(q), _dp(aux), _dp(wd), _dp(sd), _dp(amd), _dp(apd)) == 0
    assert _same(wd, wo[..., 1:], L.mode) and _same(sd, so[:, 1:], L.mode)


# ---- 2-D solvers without aux (csrc/rp2.cuh): normal() + both transverse() solves per interface ---------------------
SRC2 = os.path.join(HERE, 'host_functors', 'rp2_host.cpp')
NO_RPT2 = ('euler_hlle_2D', 'shallow_hlle_2D', 'kpp_2D')           # rpt2_dummy.f90


@pytest.fixture(scope='module', params=['strict', 'fast'])
def hostlib2(tmp_path_factory, request):
    L = _build(tmp_path_factory, SRC2, 'rp2', request.param)
    L.mode = request.param
    dp = ctypes.POINTER(ctypes.c_double)
    L.hostfun_rp2.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int] + [dp] * 9
    L.hostfun_rp2.restype = ctypes.c_int
    R = pyclaw_ref.lib()
    R.clawref_rp2.argtypes = [ctypes.c_int] * 5 + [dp] * 9 + [ctypes.c_int, dp]
    R.clawref_rp2.restype = ctypes.c_int
    return L, R


def _pencils2(name, rng, n):
    cp = []
    if name == 'advection_2D':
        q = rng.standard_normal((1, n)); cp = [0.5, -1.0]
    elif name == 'acoustics_2D':
        q = rng.standard_normal((3, n)); cp = [1.3, 2.9, np.sqrt(2.9 / 1.3), 1.3 * np.sqrt(2.9 / 1.3)]
    elif name == 'burgers_2D':
        q = rng.standard_normal((1, n)) * 2
    elif name == 'kpp_2D':
        q = rng.uniform(-3 * np.pi, 5 * np.pi, (1, n))
        q[0, ::7] = 0.25 * np.pi; q[0, 1::7] = 3.5 * np.pi       # the states of the reference example
    elif name in ('shallow_roe_with_efix_2D', 'shallow_hlle_2D'):
        h = rng.uniform(0.05, 3.0, n); u = rng.standard_normal(n) * 2; v = rng.standard_normal(n) * 2
        q = np.stack([h, h * u, h * v]); cp = [9.81]
    elif name in ('euler_4wave_2D', 'euler_5wave_2D', 'euler_hlle_2D'):
        rho = rng.uniform(0.1, 3.0, n); u = rng.standard_normal(n) * 2; v = rng.standard_normal(n) * 2
        p = rng.uniform(0.1, 5.0, n)
        if name == 'euler_hlle_2D':      # rpn2_euler_hlle.f90:60-63 (H_r with p_l, c^2 = gamma1*H - u^2/2): keep the roots real
            rho = rng.uniform(0.8, 1.2, n); p = rng.uniform(1.0, 1.3, n)
            u = rng.uniform(-1.0, 1.0, n); v = rng.uniform(-1.0, 1.0, n)
        rows = [rho, rho * u, rho * v, p / 0.4 + 0.5 * rho * (u * u + v * v)]
        if name == 'euler_5wave_2D':
            rows.append(rho * rng.uniform(0, 1, n))
        q = np.stack(rows); cp = [1.4]
    else:
        raise KeyError(name)
    q[:, n // 2: n // 2 + 8] = q[:, n // 2: n // 2 + 1]
    return np.asfortranarray(q), np.array(cp + [0.0])


NAMES_2D = ['advection_2D', 'acoustics_2D', 'burgers_2D', 'kpp_2D', 'shallow_roe_with_efix_2D', 'shallow_hlle_2D',
            'euler_4wave_2D', 'euler_5wave_2D', 'euler_hlle_2D']


@pytest.mark.parametrize('ixy', [1, 2])
@pytest.mark.parametrize('name', NAMES_2D)
def test_device_functor_source_matches_oracle_rp2(hostlib2, name, ixy):
    L, R = hostlib2
    rp_id = pyclaw_ref.RP_IDS[name]
    M, W = pyclaw_ref.NUM_EQN[name], pyclaw_ref.NUM_WAVES[name]
    rng = np.random.default_rng(4321 + 10 * rp_id + ixy)
    mx, g = 4093, 2
    n = mx + 2 * g
    q, cp = _pencils2(name, rng, n)
    z = lambda *shape: np.zeros(shape, order='F')
    wo, so, amo, apo = z(M, W, n), z(W, n), z(M, n), z(M, n)
    bo = [z(M, n) for _ in range(4)]
    assert R.clawref_rp2(ixy, M, W, g, mx, _dp(q), _dp(wo), _dp(so), _dp(amo), _dp(apo), *[_dp(b) for b in bo],
                         rp_id, _dp(cp)) == 0
    wd, sd, amd, apd = z(M, W, n - 1), z(W, n - 1), z(M, n - 1), z(M, n - 1)
    bd = [z(M, n - 1) for _ in range(4)]
    assert L.hostfun_rp2(rp_id, ixy - 1, _dp(cp), len(cp) - 1, n - 1, _dp(q), _dp(wd), _dp(sd), _dp(amd), _dp(apd),
                         *[_dp(b) for b in bd]) == 0
    pairs = [(wd, wo[..., 1:], 'wave'), (sd, so[:, 1:], 's'), (amd, amo[:, 1:], 'amdq'), (apd, apo[:, 1:], 'apdq')]
    if name not in NO_RPT2:
        pairs += [(bd[k], bo[k][:, 1:], what) for k, what in enumerate(('bmasdq(amdq)', 'bpasdq(amdq)', 'bmasdq(apdq)',
                                                                        'bpasdq(apdq)'))]
    for a, b, what in pairs:
        assert np.isfinite(b).all(), (name, what)
        assert _same(a, b, L.mode), (name, L.mode, ixy, what, float(np.max(np.abs(a - b))))


# ---- 2-D solvers that read aux (csrc/rp2_aux.cuh) ------------------------------------------------------------------
SRC2A = os.path.join(HERE, 'host_functors', 'rp2_aux_host.cpp')


@pytest.fixture(scope='module', params=['strict', 'fast'])
def hostlib2a(tmp_path_factory, request):
    L = _build(tmp_path_factory, SRC2A, 'rp2a', request.param)
    L.mode = request.param
    dp = ctypes.POINTER(ctypes.c_double)
    L.hostfun_rp2_aux.argtypes = [ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [dp] * 12 + [ctypes.c_int]
    L.hostfun_rp2_aux.restype = ctypes.c_int
    R = pyclaw_ref.lib()
    R.clawref_rp2_aux.argtypes = [ctypes.c_int] * 6 + [dp] * 12 + [ctypes.c_int, dp, ctypes.c_int]
    R.clawref_rp2_aux.restype = ctypes.c_int
    return L, R


def _pencils2a(name, rng, n):
    """q(meqn, n) and three aux pencils (rows j-1, j, j+1) with admissible coefficients."""
    cp = []

    def three(make):
        return [np.asfortranarray(make()) for _ in range(3)]

    if name == 'vc_acoustics_2D':
        q = rng.standard_normal((3, n))
        aux = three(lambda: np.stack([rng.uniform(0.5, 4.0, n), rng.uniform(0.5, 3.0, n)]))
    elif name == 'vc_advection_2D':
        q = rng.standard_normal((1, n))
        aux = three(lambda: rng.standard_normal((2, n)))
    elif name == 'psystem_2D':
        q = np.stack([rng.uniform(-0.2, 0.5, n), rng.standard_normal(n), rng.standard_normal(n)])
        aux = three(lambda: np.stack([rng.uniform(1.0, 4.0, n), rng.uniform(1.0, 4.0, n),
                                      rng.choice([1.0, 2.0, 3.0], n), rng.uniform(-0.2, 0.5, n)]))
    elif name == 'acoustics_mapped_2D':
        q = rng.standard_normal((3, n))

        def geom():
            th1, th2 = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
            return np.stack([np.cos(th1), np.sin(th1), rng.uniform(0.5, 1.5, n), np.cos(th2), np.sin(th2),
                             rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 1.5, n), rng.uniform(0.5, 4.0, n),
                             rng.uniform(0.5, 3.0, n)])
        aux = three(geom)
    elif name == 'vc_advection_vector_2D':
        q = rng.standard_normal((2, n))
        aux = three(lambda: rng.standard_normal((2, n)))
    elif name == 'vc_elasticity_2D':
        q = rng.standard_normal((5, n))

        def material():
            rho, lam, mu = rng.uniform(1.0, 4.0, n), rng.uniform(1.0, 4.0, n), rng.uniform(0.5, 2.0, n)
            mu[::6] = 0.0                                # fluid cells: no shear waves (the det == 0 branches)
            return np.stack([rho, lam, mu, np.sqrt((lam + 2 * mu) / rho), np.sqrt(mu / rho)])
        aux = three(material)
    elif name == 'shallow_bathymetry_fwave_2D':
        h = rng.uniform(0.01, 2.0, n); h[::5] = 5e-4      # isolated dry cells (two adjacent ones give 0/0 in the reference, :77)
        q = np.stack([h, h * rng.standard_normal(n), h * rng.standard_normal(n)])
        aux = three(lambda: rng.uniform(-1.0, 0.0, (1, n)))
        cp = [9.8, 1e-3, 0.0]
    else:
        raise KeyError(name)
    return np.asfortranarray(q), aux, np.array(cp + [0.0])


NAMES_2D_AUX = ['vc_acoustics_2D', 'vc_advection_2D', 'psystem_2D', 'acoustics_mapped_2D', 'shallow_bathymetry_fwave_2D',
                'vc_elasticity_2D', 'vc_advection_vector_2D']


@pytest.mark.parametrize('ixy', [1, 2])
@pytest.mark.parametrize('name', NAMES_2D_AUX)
def test_device_aux_functor_source_matches_oracle_rp2(hostlib2a, name, ixy):
    L, R = hostlib2a
    rp_id = pyclaw_ref.RP_IDS[name]
    M, W = pyclaw_ref.NUM_EQN[name], pyclaw_ref.NUM_WAVES[name]
    rng = np.random.default_rng(999 + 10 * rp_id + ixy)
    mx, g = 2045, 2
    n = mx + 2 * g
    q, aux, cp = _pencils2a(name, rng, n)
    maux = aux[0].shape[0]
    with_rpt = 0 if name in ('shallow_bathymetry_fwave_2D', 'vc_advection_vector_2D') else 1       # its rpt2 reads an unassigned value: refused
    z = lambda *shape: np.zeros(shape, order='F')
    wo, so, amo, apo = z(M, W, n), z(W, n), z(M, n), z(M, n)
    bo = [z(M, n) for _ in range(4)]
    assert R.clawref_rp2_aux(ixy, M, W, maux, g, mx, _dp(q), _dp(aux[0]), _dp(aux[1]), _dp(aux[2]), _dp(wo), _dp(so),
                             _dp(amo), _dp(apo), *[_dp(b) for b in bo], rp_id, _dp(cp), with_rpt) == 0
    wd, sd, amd, apd = z(M, W, n - 1), z(W, n - 1), z(M, n - 1), z(M, n - 1)
    bd = [z(M, n - 1) for _ in range(4)]
    assert L.hostfun_rp2_aux(rp_id, ixy - 1, _dp(cp), len(cp) - 1, n - 1, maux, _dp(q), _dp(aux[0]), _dp(aux[1]),
                             _dp(aux[2]), _dp(wd), _dp(sd), _dp(amd), _dp(apd), *[_dp(b) for b in bd], with_rpt) == 0
    pairs = [(wd, wo[..., 1:], 'wave'), (sd, so[:, 1:], 's'), (amd, amo[:, 1:], 'amdq'), (apd, apo[:, 1:], 'apdq')]
    if with_rpt:
        pairs += [(bd[k], bo[k][:, 1:], what) for k, what in enumerate(('bmasdq(amdq)', 'bpasdq(amdq)', 'bmasdq(apdq)',
                                                                        'bpasdq(apdq)'))]
    for a, b, what in pairs:
        assert np.isfinite(b).all(), (name, what)
        assert _same(a, b, L.mode), (name, L.mode, ixy, what, float(np.max(np.abs(a - b))))


# ---- 3-D solvers (csrc/rp3.cuh) ------------------------------------------------------------------------------------
SRC3 = os.path.join(HERE, 'host_functors', 'rp3_host.cpp')


@pytest.fixture(scope='module', params=['strict', 'fast'])
def hostlib3(tmp_path_factory, request):
    L = _build(tmp_path_factory, SRC3, 'rp3', request.param)
    L.mode = request.param
    dp = ctypes.POINTER(ctypes.c_double)
    L.hostfun_rp3.argtypes = [ctypes.c_int] * 3 + [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [dp] * 12 + [ctypes.c_int]
    L.hostfun_rp3.restype = ctypes.c_int
    R = pyclaw_ref.lib()
    R.clawref_rp3.argtypes = [ctypes.c_int] * 7 + [dp] * 12 + [ctypes.c_int, dp, ctypes.c_int]
    R.clawref_rp3.restype = ctypes.c_int
    return L, R


def _pencils3(name, rng, n):
    aux, cp = None, []
    if name == 'vc_acoustics_3D':
        q = rng.standard_normal((4, n)); aux = rng.uniform(0.5, 4.0, (2, n))
    elif name == 'euler_3D':
        rho = rng.uniform(0.1, 3.0, n); vel = rng.standard_normal((3, n)) * 2; p = rng.uniform(0.1, 5.0, n)
        q = np.vstack([rho, rho * vel, p / 0.4 + 0.5 * rho * np.sum(vel * vel, axis=0)]); cp = [1.4]
    elif name == 'burgers_3D':
        q = rng.standard_normal((1, n)) * 2; cp = [1.0, 0.5, -0.75]
    elif name == 'vc_advection_3D':
        q = rng.standard_normal((1, n)); aux = rng.standard_normal((3, n))
    else:
        raise KeyError(name)
    q[:, n // 2: n // 2 + 8] = q[:, n // 2: n // 2 + 1]
    return np.asfortranarray(q), (None if aux is None else np.asfortranarray(aux)), np.array(cp + [0.0])


@pytest.mark.parametrize('icoor', [2, 3])
@pytest.mark.parametrize('ixyz', [1, 2, 3])
@pytest.mark.parametrize('name', ['vc_acoustics_3D', 'euler_3D', 'burgers_3D', 'vc_advection_3D'])
def test_device_functor_source_matches_oracle_rp3(hostlib3, name, ixyz, icoor):
    """normal<IXYZ>() of every 3-D functor against rpn3, and -- for the solvers whose transverse routines read no aux --
    transverse<IXYZ,ICOOR>() against rpt3 (both fluctuations) and rptt3 (on one of the results)."""
    L, R = hostlib3
    rp_id = pyclaw_ref.RP_IDS[name]
    M, W = pyclaw_ref.NUM_EQN[name], pyclaw_ref.NUM_WAVES[name]
    rng = np.random.default_rng(77 + 100 * rp_id + 10 * ixyz + icoor)
    mx, g = 1021, 2
    n = mx + 2 * g
    q, aux, cp = _pencils3(name, rng, n)
    maux = 0 if aux is None else aux.shape[0]
    auxo = np.zeros((1, n), order='F') if aux is None else aux
    with_rpt = 1 if name in ('euler_3D', 'burgers_3D') else 0
    z = lambda *shape: np.zeros(shape, order='F')
    wo, so, amo, apo = z(M, W, n), z(W, n), z(M, n), z(M, n)
    bo = [z(M, n) for _ in range(6)]
    assert R.clawref_rp3(ixyz, icoor, M, W, maux, g, mx, _dp(q), _dp(auxo), _dp(wo), _dp(so), _dp(amo), _dp(apo),
                         *[_dp(b) for b in bo], rp_id, _dp(cp), with_rpt) == 0
    wd, sd, amd, apd = z(M, W, n - 1), z(W, n - 1), z(M, n - 1), z(M, n - 1)
    bd = [z(M, n - 1) for _ in range(6)]
    assert L.hostfun_rp3(rp_id, ixyz - 1, icoor, _dp(cp), len(cp) - 1, n - 1, maux, _dp(q), _dp(auxo), _dp(wd), _dp(sd),
                         _dp(amd), _dp(apd), *[_dp(b) for b in bd], with_rpt) == 0
    pairs = [(wd, wo[..., 1:], 'wave'), (sd, so[:, 1:], 's'), (amd, amo[:, 1:], 'amdq'), (apd, apo[:, 1:], 'apdq')]
    if with_rpt:
        names = ('bmasdq(amdq)', 'bpasdq(amdq)', 'bmasdq(apdq)', 'bpasdq(apdq)', 'cmbsasdq', 'cpbsasdq')
        pairs += [(bd[k], bo[k][:, 1:], what) for k, what in enumerate(names)]
    for a, b, what in pairs:
        assert np.isfinite(b).all(), (name, what)
        assert _same(a, b, L.mode), (name, L.mode, ixyz, icoor, what, float(np.max(np.abs(a - b))))


# ---------------------------------------------------------------------------------------------------------------
# One-launch ghost fill (clawpack_b200/csrc/bc_map.h, kernel bc_xy2): the index map against the reference's slicing order
# ---------------------------------------------------------------------------------------------------------------
def _bc_two_pass(q, lo, up, lo_on, up_on, reflect, fixed):
    """Solver._bc_lower / _bc_upper (pyclaw/src/pyclaw/solver.py:439-519) dimension by dimension, x first; q is
    (ncomp, my+4, mx+4) [row-major image of the SoA planes]; type 4 = fixed state (the device form of BC.custom)."""
    q = q.copy()
    for d in (0, 1):
        a = np.moveaxis(q, 2 - d, 0)            # a[k, ...]: k along dimension d
        n = a.shape[0] - 4
        for side in (0, 1):
            t = (lo, up)[side][d]
            on = (lo_on, up_on)[side][d]
            other_on = (up_on, lo_on)[side][d]
            if not on:
                continue
            for g in range(2):
                k = g if side == 0 else n + 2 + g
                if t == 1:
                    a[k] = a[2] if side == 0 else a[n + 1]
                elif t == 2:
                    if other_on:
                        a[k] = a[n + k] if side == 0 else a[k - n]
                elif t == 3:
                    src = 3 - k if side == 0 else 2 * n + 3 - k
                    a[k] = a[src]
                    if reflect[d] >= 0:
                        a[k, reflect[d]] = -a[src, reflect[d]]          # a[k, m, ...]: the component axis stays second
                elif t == 4:
                    for m in range(q.shape[0]):
                        a[k, m] = fixed[d][side][m]
    return q


def test_one_launch_ghost_fill_map_matches_two_passes(tmp_path_factory):
    src = os.path.join(HERE, 'host_functors', 'bc_map_host.cpp')
    L = _build(tmp_path_factory, src, 'bcmap', 'strict')
    ip = ctypes.POINTER(ctypes.c_int)
    dp = ctypes.POINTER(ctypes.c_double)
    L.hostfun_bc_xy2.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ip] * 5 + [dp] * 4
    L.hostfun_bc_xy2.restype = ctypes.c_int
    rng = np.random.default_rng(5)
    ncase = 0
    for mx, my in ((7, 5), (4, 9), (2, 2), (13, 2)):       # (a face reads interior cells only from 2 cells on)
        for _ in range(60):
            lo = rng.integers(0, 5, 2).astype(np.int32)
            up = rng.integers(0, 5, 2).astype(np.int32)
            lo_on = (rng.random(2) < 0.8).astype(np.int32)
            up_on = (rng.random(2) < 0.8).astype(np.int32)
            ncomp = int(rng.integers(1, 5))
            reflect = np.array([rng.integers(-1, ncomp), rng.integers(-1, ncomp)], dtype=np.int32)
            fixed = rng.standard_normal((2, 2, 8))
            q0 = rng.standard_normal((ncomp, my + 4, mx + 4))
            want = _bc_two_pass(q0, lo, up, lo_on, up_on, reflect, fixed)
            got = np.ascontiguousarray(q0.copy())
            ia = lambda a: a.ctypes.data_as(ip)
            da = lambda a: np.ascontiguousarray(a).ctypes.data_as(dp)
            xlo, xup, ylo, yup = (np.ascontiguousarray(fixed[0, 0]), np.ascontiguousarray(fixed[0, 1]),
                                  np.ascontiguousarray(fixed[1, 0]), np.ascontiguousarray(fixed[1, 1]))
            rc = L.hostfun_bc_xy2(got.ctypes.data_as(dp), ncomp, mx, my, ia(lo), ia(up), ia(lo_on), ia(up_on), ia(reflect),
                                  da(xlo), da(xup), da(ylo), da(yup))
            assert rc == 0
            assert np.array_equal(got, want), (mx, my, lo, up, lo_on, up_on, reflect)
            ncase += 1
    assert ncase == 240
