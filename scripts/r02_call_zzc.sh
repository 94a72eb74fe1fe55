#!/bin/bash
# ncu --set full captures that were still missing: config 1's kernel and the x-sweep of config 4
out=gpurun_out; tag=r02_zzc
ncu --set full --clock-control none --import-source on -k regex:sweep2_split_ws -s 30 -c 1 -o $out/${tag}_c1_bench -f \
   python bench.py --config c1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > $out/${tag}_ncu_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep3_x -s 3 -c 1 -o $out/${tag}_c4_x -f \
   python bench.py --config c4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > $out/${tag}_ncu_c4x.log 2>&1
ls -la $out/${tag}_*.ncu-rep; tail -2 $out/${tag}_ncu_c1.log
