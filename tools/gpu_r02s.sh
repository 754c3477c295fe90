#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 420 --csv --log-file $OUT/launches_forward_r02s.csv python tools/forward_once.py > /dev/null 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --train-batch 0 --highres-scenes 0 > $OUT/bench_n2_r02s.json 2> $OUT/bench_n2_r02s.err
tail -3 $OUT/bench_n2_r02s.err; python tools/show_bench.py $OUT/bench_n2_r02s.json | head -8; python -c "
import json; d=json.loads(open('$OUT/bench_n2_r02s.json').read().strip().splitlines()[-1]); print(d['config'])"
