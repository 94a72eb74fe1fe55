#!/bin/bash
# final single-GPU verification of the round: the whole GPU suite, smoke(), both bench arms, the launch list of the bench
# command, ncu --set full captures of the dominant kernels (bench state and dense data)
out=gpurun_out; tag=r02_zz
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> $out/${tag}_smoke.log; tail -3 $out/${tag}_smoke.log | cut -c1-200
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 2 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench.json').read().splitlines() if l.startswith('{')][-1])
r=d['roofline']; print('c2', d['value'], d['ms_per_step'], r['bound'], r['frac'], r['kernel_ms'], r.get('dense',{}).get('kernel_ms'), r.get('dense',{}).get('frac'), d['e2e']['value'], d['cpu_baseline']['value'])
for o in d['other_configs']: print(o['config'], o['value'], o['ms_per_step'], o['roofline']['bound'], o['roofline']['frac'], o['roofline'].get('dense',{}).get('kernel_ms'))
d=json.loads([l for l in open('$out/${tag}_bench_reference_arm.json').read().splitlines() if l.startswith('{')][-1])
print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > /dev/null 2>>$out/${tag}.err
ncu --set full --clock-control none --import-source on -k regex:sweep2_unsplit_ws -s 10 -c 1 -o $out/${tag}_c2_bench -f \
   python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > $out/${tag}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep2_unsplit_ws -s 3 -c 1 -o $out/${tag}_c2_dense -f \
   python scripts/quiet_probe.py --config c2 --seg 0 --data dense > $out/${tag}_ncu_c2_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep2_split_ws -s 6 -c 1 -o $out/${tag}_c3_bench -f \
   python bench.py --config c3 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > $out/${tag}_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep3_march -s 4 -c 1 -o $out/${tag}_c4_march -f \
   python bench.py --config c4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > $out/${tag}_ncu_c4.log 2>&1
ls -la $out/${tag}_*.ncu-rep
