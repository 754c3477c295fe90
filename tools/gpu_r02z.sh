#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_render_backward.py -m gpu -q -x 2>&1 | tail -25 ) > $OUT/pytest_gpu_dev.txt
cat $OUT/pytest_gpu_dev.txt
