#!/bin/bash
# 3-D kernels: one barrier per march step (double-buffered gadd/hadd exchange) + cost-model segment choice; A/B of the segments
out=gpurun_out; tag=r02_zd
python -m pytest tests/test_cuda_parity3d.py tests/test_cuda_parity_scale.py -x -q -m gpu > $out/${tag}_pytest3d.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest3d.log
tail -3 $out/${tag}_pytest3d.log
B="--config c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity"
python bench.py $B > $out/${tag}_c4.json 2> $out/${tag}_c4.err
CLAW_SEG3=512 python bench.py $B > $out/${tag}_c4_seg512.json 2>> $out/${tag}_c4.err
CLAW_SEG3=128 python bench.py $B > $out/${tag}_c4_seg128.json 2>> $out/${tag}_c4.err
python - <<PY
import json
for f in ('c4','c4_seg512','c4_seg128'):
    try:
        d=json.loads([l for l in open('$out/${tag}_%s.json'%f).read().splitlines() if l.startswith('{')][-1])
        print(f, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $out/${tag}_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_c4.csv python bench.py $B > /dev/null 2>>$out/${tag}.err
grep -c sweep3 $out/${tag}_launches_c4.csv
