#!/bin/bash
# one-launch ghost fill (bc_xy2) + control step folded into it (two launches per attempt on small grids):
# whole GPU suite, config 1 with and without the deferred control step / the one-launch fill, launch list of config 1
out=gpurun_out; tag=r02_zb
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
B="--config c1 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs"
python bench.py $B > $out/${tag}_c1.json 2> $out/${tag}_c1.err
CLAW_CTL_DEFER=0 python bench.py $B > $out/${tag}_c1_nodefer.json 2>> $out/${tag}_c1.err
CLAW_BC_TWO_PASS=1 python bench.py $B > $out/${tag}_c1_twopass.json 2>> $out/${tag}_c1.err
python - <<PY
import json
for f in ('c1','c1_nodefer','c1_twopass'):
    try:
        d=json.loads([l for l in open('$out/${tag}_%s.json'%f).read().splitlines() if l.startswith('{')][-1])
        print(f, d['ms_per_step'], d['gpu_launches'], d['steps'], d['roofline']['kernel_ms'], d.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $out/${tag}_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_c1.csv python bench.py $B --no-parity > /dev/null 2>>$out/${tag}.err
CLAW_BC_TWO_PASS=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_parity_aux.py -x -q -m gpu > $out/${tag}_pytest_twopass.log 2>&1; tail -2 $out/${tag}_pytest_twopass.log
