#!/bin/bash
# 2 GPUs, committed build (dense-data switch, 3-D capacity branch): six-case parity worker through the three transports, bench
out=gpurun_out; tag=r02_zzh
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 tests/dist_gpu_worker.py > $out/${tag}_worker.log 2>&1; echo "worker exit $?" >> $out/${tag}_worker.log
grep -c "True, step log identical: True" $out/${tag}_worker.log; grep "False\|worker exit" $out/${tag}_worker.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print('n2', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'), d['gpu_launches'])
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_n2.err | tail -3
