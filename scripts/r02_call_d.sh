#!/bin/bash
# round 2, call d: device-resident step control -- whole GPU suite, smoke, bench, other configs
out=gpurun_out; tag=r02_d
python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -25 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -3 $out/${tag}_smoke.log
python bench.py --no-cpu-baseline --steps 20 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 1500 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
python scripts/bench_configs.py --configs c1,c3,c4 --steps 10 > $out/${tag}_configs.jsonl 2>>$out/${tag}_bench.err; cut -c1-300 $out/${tag}_configs.jsonl
