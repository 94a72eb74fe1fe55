#!/bin/bash
# tile ordering with three rotating lists (no tickets): probe A/B, ncu of the shallow-water split kernel
out=gpurun_out; tag=r02_p
CLAW_TUNE=0 python scripts/quiet_probe.py --config c2 --seg 32 64 128 --data actual uniform shock > $out/${tag}_tune0.jsonl 2>$out/${tag}.err
CLAW_TUNE=2 python scripts/quiet_probe.py --config c2 --seg 64 --data actual uniform > $out/${tag}_tune2.jsonl 2>>$out/${tag}.err
CLAW_TUNE=0 python scripts/quiet_probe.py --config c3 --seg 32 64 --data actual uniform > $out/${tag}_c3_tune0.jsonl 2>>$out/${tag}.err
cat $out/${tag}_*.jsonl
tail -5 $out/${tag}.err
ncu --set full --clock-control none --import-source on -k regex:sweep2_split_ws -s 3 -c 1 -o $out/${tag}_c3_uniform -f \
   python scripts/quiet_probe.py --config c3 --seg 64 --data uniform > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log
