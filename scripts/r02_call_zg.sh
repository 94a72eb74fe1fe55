#!/bin/bash
# 2 GPUs: peer transport with the barrier before the push of 3-D dimension-split steps (race found in r02_zf:
# euler3d_split_extrap through peer memory), vectorised peer copies; 3-D Riemann-solver plug-in tests
out=gpurun_out; tag=r02_zg
for rep in 1 2 3; do
CLAW_TEST_TRANSPORTS=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$rep tests/dist_gpu_worker.py > $out/${tag}_worker_peer_$rep.log 2>&1; echo "worker exit $?" >> $out/${tag}_worker_peer_$rep.log
grep -c "True, step log identical: True" $out/${tag}_worker_peer_$rep.log; grep "False\|worker exit" $out/${tag}_worker_peer_$rep.log | cut -c1-200
done
timeout 900 python -m pytest tests/test_rp_plugin.py tests/test_multigpu.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-dense > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench_n2.json').read().splitlines() if l.startswith('{')][-1])
print('n2', d['n_gpus'], d['value'], d['ms_per_step'], d.get('parity',{}).get('bit_identical'), d['gpu_launches'])
for o in d.get('other_configs',[]): print('  ', o['config'], o['n_gpus'], o['value'], o['ms_per_step'])
"
grep -v "^\*\|OMP_NUM\|^$" $out/${tag}_bench_n2.err | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/slab_probe.py --config c4 --world 2 --ranks 0 --steps 4 > $out/${tag}_slab_c4_w2.jsonl 2>>$out/${tag}.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/slab_probe.py --config c3 --world 2 --ranks 0 --steps 10 > $out/${tag}_slab_c3_w2.jsonl 2>>$out/${tag}.err
cat $out/${tag}_slab_c*.jsonl; tail -3 $out/${tag}.err
