import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from test_cuda_parity import CASES, _run_cuda, _run_ref
names = sys.argv[1:] or sorted(CASES)
for name in names:
    sc = CASES[name]()
    for nst in (2,):
        q, log, _ = _run_cuda(sc, nst, 'strict')
        qr, logr = _run_ref(sc, nst)
        d = np.abs(q - qr)
        bad = np.argwhere(d > 0)
        print(name, 'steps', nst, 'log_equal', log == logr, 'maxdiff', d.max(), 'nbad', len(bad), 'of', d.size,
              'first', bad[:6].tolist())
