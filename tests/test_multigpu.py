"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): NCCL ghost exchange + CFL all-reduce
of the device-resident path against the serial oracle, bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize('world', [2, 4])
def test_slab_nccl_matches_serial_oracle(world):
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29700 + world), os.path.join(HERE, 'dist_gpu_worker.py')]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-4000:]
    assert p.stdout.count('q bit-identical to serial oracle: True, step log identical: True') == 18, p.stdout[-6000:]      # six cases x three transports (torch process group, C ABI, peer memory)
