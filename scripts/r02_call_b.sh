#!/bin/bash
# round 2, call b: FP64 opcode-mix probe, kpp tests, 3-D unsplit variants
out=gpurun_out; tag=r02_b
scripts/probes/fp64_issue r2 > $out/${tag}_fp64_mix_probe.txt 2>&1; cat $out/${tag}_fp64_mix_probe.txt
python -m pytest tests -q -m gpu -k "kpp" > $out/${tag}_pytest_kpp.log 2>&1; tail -3 $out/${tag}_pytest_kpp.log
for v in 0 5 6; do CLAW_VARIANT3=$v python scripts/bench_configs.py --configs c4 --steps 5 >> $out/${tag}_c4_variants.jsonl 2>>$out/${tag}.err; done
cut -c1-330 $out/${tag}_c4_variants.jsonl
CLAW_VARIANT3=6 python -m pytest tests/test_cuda_parity3d.py -q -m gpu > $out/${tag}_pytest_gpu_variant3_6.log 2>&1; tail -2 $out/${tag}_pytest_gpu_variant3_6.log
python scripts/bench_configs.py --configs c1 --steps 200 >> $out/${tag}_c1.jsonl 2>>$out/${tag}.err; cut -c1-330 $out/${tag}_c1.jsonl
