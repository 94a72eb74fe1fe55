#!/bin/bash
# lean quiet loop + tile ordering: probe (A/B through CLAW_TUNE), the whole GPU suite, bench
out=gpurun_out; tag=r02_o
CLAW_TUNE=0 python scripts/quiet_probe.py --config c2 --seg 32 64 128 --data actual uniform shock > $out/${tag}_tune0.jsonl 2>$out/${tag}.err
CLAW_TUNE=2 python scripts/quiet_probe.py --config c2 --seg 32 64 --data actual uniform shock > $out/${tag}_tune2.jsonl 2>>$out/${tag}.err
CLAW_TUNE=0 python scripts/quiet_probe.py --config c3 --seg 32 64 --data actual uniform > $out/${tag}_c3_tune0.jsonl 2>>$out/${tag}.err
CLAW_TUNE=2 python scripts/quiet_probe.py --config c3 --seg 32 64 --data actual > $out/${tag}_c3_tune2.jsonl 2>>$out/${tag}.err
cat $out/${tag}_*.jsonl
tail -5 $out/${tag}.err
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 3000 $out/${tag}_bench.json
