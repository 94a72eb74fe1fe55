#!/bin/bash
# 2 GPUs: peer-memory transport: parity (six cases) + bench A/B
out=gpurun_out; tag=r02_x
CLAW_TEST_TRANSPORTS=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 tests/dist_gpu_worker.py > $out/${tag}_worker_peer.log 2>&1; echo "worker exit $?" >> $out/${tag}_worker_peer.log
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_worker_peer.log | tail -14 | cut -c1-250
for tr in torch peer; do
CLAW_HALO=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n2_$tr.json 2> $out/${tag}_bench_n2_$tr.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2_$tr.json').read().splitlines() if l.startswith('{')][-1])
print('$tr', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'), d['gpu_launches'])
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_n2_$tr.err | tail -3
done
