#!/bin/bash
# short pass: GPU tests, smoke(), one default bench run   usage: tools/gpu_dev.sh <tag>
TAG=${1:-dev}
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $OUT/pytest_gpu_$TAG.txt
( timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ) > $OUT/smoke_$TAG.txt
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
cat $OUT/pytest_gpu_$TAG.txt $OUT/smoke_$TAG.txt; python tools/show_bench.py $OUT/bench_$TAG.json | cut -c1-330; tail -2 $OUT/bench_$TAG.err
