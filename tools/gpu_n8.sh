#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 300 --warmup 5 --train-batch 8 --train-steps 2 --highres-scenes 2 --full-steps 100 > $OUT/bench_n8_r02z.json 2> $OUT/bench_n8_r02z.err
tail -3 $OUT/bench_n8_r02z.err; python tools/show_bench.py $OUT/bench_n8_r02z.json 2>/dev/null | cut -c1-400 | head -12
python -c "
import json; d=json.loads(open('$OUT/bench_n8_r02z.json').read().strip().splitlines()[-1]); print(d['config']['parallelism'])"
nvidia-smi topo -m 2>/dev/null | head -14 > $OUT/topo_n8.txt; cat $OUT/topo_n8.txt | cut -c1-150
