#!/bin/bash
# 2 GPUs after the one-launch ghost fill: parity of the transports (six cases vs the serial oracle), the multi-GPU pytest
# file, bench at N=2; plus (one GPU) the per-slab cost of the strong-scaling configs
out=gpurun_out; tag=r02_zf
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 tests/dist_gpu_worker.py > $out/${tag}_worker.log 2>&1; echo "worker exit $?" >> $out/${tag}_worker.log
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_worker.log | tail -12 | cut -c1-250
timeout 900 python -m pytest tests/test_multigpu.py tests/test_rp_plugin.py -x -q -m gpu > $out/${tag}_pytest_mg.log 2>&1; tail -3 $out/${tag}_pytest_mg.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print('n2', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'), d['gpu_launches'])
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_n2.err | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/slab_probe.py --config c3 --world 8 --ranks 0 3 --steps 20 > $out/${tag}_slab_c3.jsonl 2>>$out/${tag}.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/slab_probe.py --config c4 --world 8 --ranks 0 3 --steps 6 > $out/${tag}_slab_c4.jsonl 2>>$out/${tag}.err
cat $out/${tag}_slab_c*.jsonl; tail -3 $out/${tag}.err
