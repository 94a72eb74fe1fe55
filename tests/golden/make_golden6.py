"""The reference's own binary regression frame (pyclaw/src/pyclaw/tests/test_data/advection_2d_binary: an AMRClaw
frame of 17 patches on two levels, written WITH two ghost cells per side, old-style .t header) -> tests/golden/
amrclaw_binary_frame/ : the three DATA files as they are, plus what the reference's own reader
(pyclaw/src/pyclaw/fileio/binary.py) parses from them (reference_parsed.npz).

Run in the build container only (reads /root/reference):   python tests/golden/make_golden6.py"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden5 import reference_reader, pyclaw          # noqa: E402

SRC = '/root/reference/pyclaw/src/pyclaw/tests/test_data/advection_2d_binary'

if __name__ == '__main__':
    ref_binary, _ = reference_reader()
    dst = os.path.join(HERE, 'amrclaw_binary_frame')
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    for f in ('fort.t0000', 'fort.q0000', 'fort.b0000'):
        shutil.copyfile(os.path.join(SRC, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    sol = pyclaw.Solution()
    ref_binary.read(sol, 0, SRC, 'fort', read_aux=False, options={})
    parsed = {'nstates': np.array(len(sol.states)), 't': np.array(sol.states[0].t)}
    for k, st in enumerate(sol.states):
        parsed['q%02d' % k] = np.array(st.q, dtype=np.float64)
        parsed['meta%02d' % k] = np.array([st.patch.patch_index, st.patch.level] + [d.num_cells for d in st.patch.dimensions]
                                          + [d.lower for d in st.patch.dimensions] + [d.delta for d in st.patch.dimensions], dtype=np.float64)
    np.savez_compressed(os.path.join(dst, 'reference_parsed.npz'), **parsed)
    print('patches', len(sol.states), 'levels', sorted(set(st.patch.level for st in sol.states)),
          'cells', sum(st.q[0].size for st in sol.states))
