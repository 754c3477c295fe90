#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python tools/try_train_graph.py 32 2>&1 | tail -5 ) > $OUT/try_train_graph.txt
( timeout 900 python bench.py --no-cpu --highres-scenes 0 --skip-full --steps 50 --e2e-steps 50 --train-batch 16 ) > $OUT/bench_dev16.json 2> $OUT/bench_dev.err
( timeout 900 python bench.py --no-cpu --highres-scenes 0 --skip-full --steps 50 --e2e-steps 50 --train-batch 32 --train-steps 6 ) > $OUT/bench_dev32.json 2> $OUT/bench_dev.err
cat $OUT/try_train_graph.txt; python tools/show_bench.py $OUT/bench_dev16.json | grep train_step; python tools/show_bench.py $OUT/bench_dev32.json | grep train_step
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv
