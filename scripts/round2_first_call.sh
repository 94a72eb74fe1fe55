#!/bin/bash
# First GPU call of a round, in ONE gpurun invocation (everything lands in gpurun_out/):
#   1. the GPU parity suite (incl. tests/test_zz_cuda_parity_more.py, which round 1 could only check on the CPU)
#   2. smoke()
#   3. the bench line (default workload) and the reference arm
#   4. kernel-only timings of the tuning variants worth a look (0 = default, 50 = march loop unrolled x2)
#   5. the ncu launch list of a short bench run
# usage:  gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh r02_a'
tag=${1:-r02_a}
out=gpurun_out
mkdir -p $out
python -m pytest tests -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -5 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 3 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
bash scripts/variant_sweep.sh "0 50 7" $out/${tag}_variants.jsonl
# shallow-water split kernel variants (config 3) are timed by scripts/bench_configs.py with CLAW_VARIANT=0, 51..54 (53/54: compile-time limiter)
for v in 0 51 52 53 54; do CLAW_VARIANT=$v python scripts/bench_configs.py --configs c3 >> $out/${tag}_c3_variants.jsonl 2>>$out/${tag}_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
# parity of the specialised shallow-water split kernel (variant 53) before it may become the default
CLAW_VARIANT=53 python -m pytest tests -q -m gpu -k "shallow or dam" > $out/${tag}_pytest_gpu_variant53.log 2>&1; tail -2 $out/${tag}_pytest_gpu_variant53.log
# 3-D: method set as template arguments (CLAW_VARIANT3=5) -- parity on the hardware, then timing against the generic kernels
CLAW_VARIANT3=5 python -m pytest tests/test_cuda_parity3d.py -q -m gpu > $out/${tag}_pytest_gpu_variant3_5.log 2>&1; tail -2 $out/${tag}_pytest_gpu_variant3_5.log
for v in 0 5 6; do CLAW_VARIANT3=$v python scripts/bench_configs.py --configs c4s >> $out/${tag}_c4_variants.jsonl 2>>$out/${tag}_bench.err; done
