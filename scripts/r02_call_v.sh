#!/bin/bash
# 2 GPUs: multi-GPU parity (both transports) + bench N=2 with the exchange through torch.distributed / through the C ABI
out=gpurun_out; tag=r02_v
python -m pytest tests/test_multigpu.py -x -q -m gpu > $out/${tag}_pytest_mg.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_mg.log
tail -30 $out/${tag}_pytest_mg.log | cut -c1-300
for tr in torch cabi; do
CLAW_HALO=$tr python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n2_$tr.json 2> $out/${tag}_bench_n2_$tr.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2_$tr.json').read().splitlines() if l.startswith('{')][-1])
print('$tr', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'))
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
tail -2 $out/${tag}_bench_n2_$tr.err
done
