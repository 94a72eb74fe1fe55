"""User-supplied Riemann solvers for ``kernel_language='CUDA'``.

The reference lets a user compile a Fortran Riemann solver with f2py and hand the module to the solver
(``solver.rp = my_module``; its ``rpn2`` / ``rpt2`` C pointers reach the stepper in
pyclaw/src/pyclaw/classic/solver.py:516-521).  The CUDA path needs the solver as a device functor instead
(interface: clawpack_b200/csrc/rp2.cuh, rp2_aux.cuh; recipe: clawpack_b200/csrc/rp_plugin.cu):

    rp = riemann.compile_cuda_solver('my_solver.cuh', 'MySolver', cparam=('gamma',))
    solver = pyclaw.ClawSolver2D(rp)            # num_eqn / num_waves come from the functor
    solver.kernel_language = 'CUDA'

``compile_cuda_solver`` runs nvcc for sm_100a (two arithmetic modes, like the built-in solvers) and registers the
shared object with libclawcuda; ``load_cuda_solver`` registers one that is already built.
"""
import ctypes
import hashlib
import os
import subprocess

from . import static
from ._make import make as _make

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'csrc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
HOSTCXX = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda_solver(header, functor, out_dir=None, name=None, force=False, num_dim=2):
    """Compile clawpack_b200/csrc/rp_plugin.cu against the functor `functor` defined in `header`; returns the path of
    the shared object (lib<name>.so in `out_dir`, default: next to the header).  Works without a GPU.  num_dim=3: a 3-D
    solver (functor interface of clawpack_b200/csrc/rp3.cuh)."""
    header = os.path.abspath(header)
    out_dir = os.path.abspath(out_dir or os.path.join(os.path.dirname(header), '_build'))
    os.makedirs(out_dir, exist_ok=True)
    name = name or 'clawrp_' + functor.lower()
    out = os.path.join(out_dir, 'lib%s.so' % name)
    deps = [header] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    if not force and not _stale(out, deps):
        return out
    common = ['nvcc', '-ccbin', HOSTCXX] + ARCH + ['-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-I', CSRC,
                                                    '-DCLAW_PLUGIN_HEADER="%s"' % header, '-DCLAW_PLUGIN_FUNCTOR=%s' % functor,
                                                    '-DCLAW_PLUGIN_NDIM=%d' % num_dim]
    tag = hashlib.sha1((header + functor).encode()).hexdigest()[:10]
    objs, jobs = [], []
    for ns, flags in (('strict', ['-DCLAW_NS=strict', '-DCLAW_STRICT=1', '-fmad=false']),
                      ('fast', ['-DCLAW_NS=fast', '-fmad=true'])):
        o = os.path.join(out_dir, '%s_%s_%s.o' % (name, tag, ns))
        jobs.append(subprocess.Popen(common + flags + ['-c', os.path.join(CSRC, 'rp_plugin.cu'), '-o', o]))
        objs.append(o)
    for j in jobs:
        if j.wait() != 0:
            raise RuntimeError('nvcc failed on the Riemann-solver plug-in %s (%s)' % (functor, header))
    subprocess.check_call(['nvcc', '-ccbin', HOSTCXX] + ARCH + ['-shared', '-o', out] + objs + ['-lcudart'])
    for o in objs:
        os.remove(o)
    return out


def load_cuda_solver(so_path, name=None, cparam=()):
    """Register a built plug-in with libclawcuda and return the module-like solver object that ``ClawSolver2D`` takes
    (like the objects of clawpack_b200.riemann: rp_id, cparam, num_dim).  `cparam` = the ordered names of the values the
    functor reads from ``RpParams::p`` (they are filled from ``state.problem_data`` like COMMON /cparam/)."""
    from .. import _lib
    L = _lib.lib()
    so_path = os.path.abspath(so_path)
    rp_id = ctypes.c_int(0)
    _lib.check(L.claw_rp_plugin_load(so_path.encode(), ctypes.byref(rp_id)))
    info = (ctypes.c_int * 5)()
    _lib.check(L.claw_rp_plugin_info(rp_id.value, info))
    if len(cparam) > 4:
        raise ValueError('a plug-in functor reads at most 4 cparam values (RpParams::p)')
    name = name or os.path.splitext(os.path.basename(so_path))[0].replace('lib', '', 1)
    mod = _make(name, rp_id.value, tuple(cparam), int(info[0]))
    mod.plugin_path = so_path
    mod.num_eqn, mod.num_waves, mod.num_aux_read = int(info[1]), int(info[2]), int(info[3])
    mod.no_transverse = bool(info[4])
    static.num_eqn[name] = mod.num_eqn          # Solver.__init__ resolves them by name (pyclaw/solver.py:230-238)
    static.num_waves[name] = mod.num_waves
    static.fwave[name] = False
    return mod


def compile_cuda_solver(header, functor, cparam=(), out_dir=None, name=None, force=False, num_dim=2):
    """build_cuda_solver + load_cuda_solver."""
    so = build_cuda_solver(header, functor, out_dir=out_dir, name=name, force=force, num_dim=num_dim)
    return load_cuda_solver(so, name=name or functor, cparam=cparam)
