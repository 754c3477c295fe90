#!/bin/bash
TAG=${1:-r02p}; OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|batched vs|own-chain" | tail -30 ) > $OUT/pytest_gpu_$TAG.txt
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
( timeout 600 python tools/profile_forward.py 2>&1 | grep -E "tc_conv=True two_stream=True|vgn|depth-mean" ) > $OUT/profile_forward_$TAG.txt
cat $OUT/pytest_gpu_$TAG.txt; tail -3 $OUT/bench_$TAG.err; python tools/show_bench.py $OUT/bench_$TAG.json; cat $OUT/profile_forward_$TAG.txt
