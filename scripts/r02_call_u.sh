#!/bin/bash
# 2 GPUs: multi-GPU parity tests + bench N=2 (per-step overhead of the exchange now that the step is < 1 ms) + c1 after the segment change
out=gpurun_out; tag=r02_u
python bench.py --config c1 --steps 200 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print('c1', r['kernel_ms'], d['ms_per_step'], d['gpu_launches'], r['launch'])"
python -m pytest tests/test_multigpu.py -x -q -m gpu > $out/${tag}_pytest_mg.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_mg.log
tail -4 $out/${tag}_pytest_mg.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-dense > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
for f in n1 n2; do python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_$f.json').read().splitlines() if l.startswith('{')][-1])
print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity'))
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"; done
tail -3 $out/${tag}_bench_n2.err
