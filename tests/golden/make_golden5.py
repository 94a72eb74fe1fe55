"""Binary frame files written by THIS package's writer (clawpack_b200/pyclaw/fileio/binary.py) together with what the
REFERENCE's own reader (pyclaw/src/pyclaw/fileio/binary.py) parses from them -> tests/golden/binary_frames/.

Run in the build container only (reads /root/reference):   python tests/golden/make_golden5.py
The reference modules (fileio/binary.py, fileio/ascii.py:read_t, util.py:read_data_line) are loaded from their source
files with `clawpack.pyclaw` aliased to this package (the reader only constructs Dimension / Patch / State / Domain
objects, which both share); only the parsed ARRAYS are stored (reference_parsed.npz)."""
import importlib.util
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import clawpack_b200                                       # noqa: E402,F401
from clawpack_b200 import pyclaw                           # noqa: E402

REF = '/root/reference/pyclaw/src/pyclaw'


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_reader():
    fake = types.ModuleType('clawpack')
    fake.pyclaw = pyclaw
    sys.modules['clawpack'] = fake
    sys.modules['clawpack.pyclaw'] = pyclaw
    sys.modules['clawpack.pyclaw.geometry'] = pyclaw.geometry
    sys.modules['clawpack.pyclaw.state'] = pyclaw.state
    util = _load('clawpack.pyclaw.util', os.path.join(REF, 'util.py'))
    fio = types.ModuleType('clawpack.pyclaw.fileio')
    fio.__path__ = [os.path.join(REF, 'fileio')]
    sys.modules['clawpack.pyclaw.fileio'] = fio
    _load('clawpack.pyclaw.fileio.ascii', os.path.join(REF, 'fileio', 'ascii.py'))
    return _load('clawpack.pyclaw.fileio.binary', os.path.join(REF, 'fileio', 'binary.py')), util


def small_solutions():
    """Deterministic 1-D / 2-D / 3-D solutions (shared with tests/test_host.py::test_binary_frames_*)."""
    out = []
    dom = pyclaw.Domain([pyclaw.Dimension(-1.0, 2.0, 7, name='x')])
    st = pyclaw.State(dom, 2, 0)
    k = np.arange(14, dtype=np.float64).reshape((2, 7), order='F')
    st.q[...] = np.sin(k) * 10.0 ** ((k % 5) - 2)
    st.t = 0.75
    out.append(('1d', pyclaw.Solution(st, dom), 2, 'binary64'))
    dom = pyclaw.Domain([pyclaw.Dimension(-1.0, 2.0, 4, name='x'), pyclaw.Dimension(0.5, 1.25, 3, name='y')])
    st = pyclaw.State(dom, 2, 1)
    k = np.arange(24, dtype=np.float64).reshape((2, 4, 3), order='F')
    st.q[...] = np.sin(k) * 10.0 ** ((k % 7) - 3) * np.where(k % 3 == 0, -1.0, 1.0)
    st.aux[...] = (np.cos(np.arange(12.0)) * 1e-3).reshape((1, 4, 3), order='F')
    st.t = 0.1234567891
    st.problem_data['gamma'] = 1.4
    out.append(('2d', pyclaw.Solution(st, dom), 3, 'binary64'))
    out.append(('2d_f32', pyclaw.Solution(st, dom), 4, 'binary32'))
    dims = [pyclaw.Dimension(0.0, 1.0, 3, name='x'), pyclaw.Dimension(0.0, 2.0, 2, name='y'),
            pyclaw.Dimension(-1.0, 1.0, 2, name='z')]
    dom = pyclaw.Domain(dims)
    st = pyclaw.State(dom, 3, 0)
    k = np.arange(36, dtype=np.float64).reshape((3, 3, 2, 2), order='F')
    st.q[...] = np.exp(-k / 5.0) * np.where(k % 2 == 0, 1.0, -1e5)
    st.t = 12.5
    out.append(('3d', pyclaw.Solution(st, dom), 12, 'binary64'))
    return out


if __name__ == '__main__':
    ref_binary, _ = reference_reader()
    dst = os.path.join(HERE, 'binary_frames')
    shutil.rmtree(dst, ignore_errors=True)
    parsed = {}
    for name, sol, frame, fmt in small_solutions():
        d = os.path.join(dst, name)
        sol.write(frame, d, fmt, write_aux=True)
        os.remove(os.path.join(d, 'fort.pkl' + str(frame).zfill(4)))       # pickles are not portable fixtures
        back = pyclaw.Solution()
        ref_binary.read(back, frame, d, 'fort', read_aux=True, options={'format': fmt})
        parsed[name + '__q'] = np.array(back.states[0].q, dtype=np.float64)
        parsed[name + '__t'] = np.array(back.states[0].t)
        parsed[name + '__lower'] = np.array([dm.lower for dm in back.states[0].patch.dimensions])
        parsed[name + '__delta'] = np.array([dm.delta for dm in back.states[0].patch.dimensions])
        parsed[name + '__num_cells'] = np.array([dm.num_cells for dm in back.states[0].patch.dimensions])
        if sol.num_aux > 0:
            parsed[name + '__aux'] = np.array(back.states[0].aux, dtype=np.float64)
        tol = 0 if fmt == 'binary64' else 1e-7 * np.max(np.abs(sol.state.q))
        assert np.max(np.abs(parsed[name + '__q'] - sol.state.q)) <= tol, name
        print(name, 'reference reader parsed', parsed[name + '__q'].shape, 'max diff vs written',
              np.max(np.abs(parsed[name + '__q'] - sol.state.q)))
    np.savez_compressed(os.path.join(dst, 'reference_parsed.npz'), **parsed)
