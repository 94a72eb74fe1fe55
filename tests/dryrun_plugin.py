"""DEVELOPMENT TOOL (test infrastructure, never imported by the product or by a normal test run).

A pytest plug-in that reroutes ``kernel_language='CUDA'`` in ``scenarios.make_claw`` to the ORACLE plug-in
(tests/oracle_fmod.py), so that the Python side of the ``-m gpu`` tests -- scenario construction, callbacks, golden
lookups, assertions -- can be exercised on a machine without a GPU before GPU time is spent:

    PYTHONPATH=tests python -m pytest tests/test_zz_cuda_parity_more.py -q -m gpu -p dryrun_plugin

Tests that need the real library (device-resident handles, refusal codes, pipelined transfers) fail in this mode by
design.  It proves nothing about the kernels; tests/test_kernel_emulation.py and the hardware run do that."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

from clawpack_b200 import scenarios as S        # noqa: E402
from oracle_fmod import OracleFmod              # noqa: E402

_orig = S.make_claw


def make_claw(sc, kernel_language='CUDA', pyclaw=None, **kw):
    kw.pop('cuda_precision', None)
    if kernel_language == 'CUDA':
        return _orig(sc, kernel_language='Fortran', pyclaw=pyclaw, fmod=OracleFmod(), **kw)
    return _orig(sc, kernel_language=kernel_language, pyclaw=pyclaw, **kw)


S.make_claw = make_claw
