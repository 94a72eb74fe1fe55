#!/bin/bash
# round 2, call e (2 GPUs): new bench.py at N=1 and N=2, reference arm, multi-GPU parity
out=gpurun_out; tag=r02_e
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; tail -c 3000 $out/${tag}_bench_n1.json; tail -5 $out/${tag}_bench_n1.err
python -m pytest tests/test_multigpu.py -q -m gpu -x > $out/${tag}_pytest_multigpu.log 2>&1; tail -15 $out/${tag}_pytest_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; tail -c 2500 $out/${tag}_bench_n2.json; tail -5 $out/${tag}_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 2 > $out/${tag}_bench_ref_n2.json 2> $out/${tag}_bench_ref_n2.err; tail -c 1200 $out/${tag}_bench_ref_n2.json
