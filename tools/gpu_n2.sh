#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 300 --warmup 5 > $OUT/bench_n2_r02z.json 2> $OUT/bench_n2_r02z.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > $OUT/bench_ref_n2_r02z.json 2> $OUT/bench_ref_n2_r02z.err
tail -3 $OUT/bench_n2_r02z.err; python tools/show_bench.py $OUT/bench_n2_r02z.json; head -c 400 $OUT/bench_ref_n2_r02z.json
