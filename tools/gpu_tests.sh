#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -${1:-60} ) > $OUT/pytest_gpu_dev.txt
cat $OUT/pytest_gpu_dev.txt
