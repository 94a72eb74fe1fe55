#!/bin/bash
# quiet-path probe: data variants x segment lengths, with and without the warm-up classification
out=gpurun_out; tag=r02_n
CLAW_TUNE=1 python scripts/quiet_probe.py --config c2 --seg 64 --data actual uniform shock > $out/${tag}_tune1.jsonl 2>$out/${tag}.err
CLAW_TUNE=0 python scripts/quiet_probe.py --config c2 --seg 32 64 128 --data actual uniform shock > $out/${tag}_tune0.jsonl 2>>$out/${tag}.err
CLAW_TUNE=1 python scripts/quiet_probe.py --config c3 --seg 64 --data actual uniform > $out/${tag}_c3_tune1.jsonl 2>>$out/${tag}.err
CLAW_TUNE=0 python scripts/quiet_probe.py --config c3 --seg 32 64 128 --data actual uniform > $out/${tag}_c3_tune0.jsonl 2>>$out/${tag}.err
cat $out/${tag}_*.jsonl
tail -5 $out/${tag}.err
