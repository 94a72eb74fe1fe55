"""Golden frame FILES written by the reference's own pure-Python ascii writer
(pyclaw/src/pyclaw/fileio/ascii.py) for two small deterministic solutions -> tests/golden/ascii_frames/.

Run in the build container only (reads /root/reference):   python tests/golden/make_golden3.py
The reference module is loaded from its source file with `clawpack.pyclaw` aliased to this package (the writer only
touches Solution / State / Patch attributes that both share); only its OUTPUT is stored."""
import importlib.util
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import clawpack_b200                                       # noqa: E402
from clawpack_b200 import pyclaw                           # noqa: E402

fake = types.ModuleType('clawpack')
fake.pyclaw = pyclaw
util = types.ModuleType('clawpack.pyclaw.util')
util.read_data_line = lambda *a, **k: None
sys.modules['clawpack'] = fake
sys.modules['clawpack.pyclaw'] = pyclaw
sys.modules['clawpack.pyclaw.util'] = util
spec = importlib.util.spec_from_file_location('ref_ascii', '/root/reference/pyclaw/src/pyclaw/fileio/ascii.py')
ref_ascii = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_ascii)


def small_solutions():
    """The two solutions of tests/test_host.py::test_ascii_frames_match_reference_writer."""
    out = []
    x = pyclaw.Dimension(-1.0, 2.0, 4, name='x')
    y = pyclaw.Dimension(0.5, 1.25, 3, name='y')
    dom = pyclaw.Domain([x, y])
    st = pyclaw.State(dom, 2, 1)
    k = np.arange(2 * 4 * 3, dtype=np.float64).reshape((2, 4, 3), order='F')
    st.q[...] = np.sin(k) * 10.0 ** ((k % 7) - 3) * np.where(k % 3 == 0, -1.0, 1.0)
    st.aux[...] = (np.cos(np.arange(12.0)) * 1e-3).reshape((1, 4, 3), order='F')
    st.t = 0.1234567891
    st.problem_data['gamma'] = 1.4
    out.append(('2d', pyclaw.Solution(st, dom), 3))
    dims = [pyclaw.Dimension(0.0, 1.0, 3, name='x'), pyclaw.Dimension(0.0, 2.0, 2, name='y'),
            pyclaw.Dimension(-1.0, 1.0, 2, name='z')]
    dom = pyclaw.Domain(dims)
    st = pyclaw.State(dom, 3, 0)
    k = np.arange(3 * 3 * 2 * 2, dtype=np.float64).reshape((3, 3, 2, 2), order='F')
    st.q[...] = np.exp(-k / 5.0) * np.where(k % 2 == 0, 1.0, -1e5)
    st.t = 12.5
    out.append(('3d', pyclaw.Solution(st, dom), 12))
    return out


if __name__ == '__main__':
    dst = os.path.join(HERE, 'ascii_frames')
    shutil.rmtree(dst, ignore_errors=True)
    for name, sol, frame in small_solutions():
        d = os.path.join(dst, name)
        os.makedirs(d)
        ref_ascii.write(sol, frame, d, write_aux=True)
        for f in os.listdir(d):
            if '.pkl' in f:                                # the pickle names module paths: not a byte-level fixture
                os.remove(os.path.join(d, f))
        print(name, sorted(os.listdir(d)))
