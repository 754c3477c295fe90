#!/bin/bash
TAG=r02b; OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 ) > $OUT/pytest_gpu_$TAG.txt
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
cat $OUT/pytest_gpu_$TAG.txt; tail -5 $OUT/bench_$TAG.err; python tools/show_bench.py $OUT/bench_$TAG.json 2>/dev/null | head -60; head -c 1500 $OUT/bench_ref_$TAG.json
