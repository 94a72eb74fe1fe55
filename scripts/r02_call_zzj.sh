#!/bin/bash
# 8 GPUs, committed build: bench (its parity check runs a slab case on the 8 ranks against the serial oracle first)
out=gpurun_out; tag=r02_zzj
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
print('n8', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'), d['gpu_launches'])
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_n8.err | tail -3
