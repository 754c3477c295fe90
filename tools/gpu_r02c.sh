#!/bin/bash
TAG=r02c; OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > $OUT/pytest_gpu_$TAG.txt
( timeout 600 python tools/profile_forward.py 2>&1 | tail -70 ) > $OUT/profile_forward_$TAG.txt
cat $OUT/pytest_gpu_$TAG.txt; cat $OUT/profile_forward_$TAG.txt
