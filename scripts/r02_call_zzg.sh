#!/bin/bash
# ncu --set full of the kernel roofline.dense measures since the dense-data switch: the Euler kernel WITHOUT the fast path on
# dense data (launch #4 of quiet_probe: 0 = untimed fast-path launch, 1-2 warm-up, 3 = untimed fast-path launch, 4-6 timed)
out=gpurun_out; tag=r02_zzg
ncu --set full --clock-control none --import-source on -k regex:sweep2_unsplit_ws -s 4 -c 1 -o $out/${tag}_c2_dense_plain -f \
   python scripts/quiet_probe.py --config c2 --seg 0 --data dense --reps 3 > $out/${tag}_ncu.log 2>&1
ls -la $out/${tag}_*.ncu-rep; grep "kernel_ms" $out/${tag}_ncu.log | cut -c1-200
