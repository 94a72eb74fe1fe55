#!/bin/bash
# 8 GPUs, final build: peer transport parity at N = 4 (six cases vs the serial oracle), bench at N = 8
out=gpurun_out; tag=r02_zy
CLAW_TEST_TRANSPORTS=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29720 tests/dist_gpu_worker.py > $out/${tag}_worker_peer_n4.log 2>&1; echo "worker exit $?" >> $out/${tag}_worker_peer_n4.log
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_worker_peer_n4.log | grep "case\|worker exit" | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
print('n8', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'), d['gpu_launches'])
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_n8.err | tail -3
