#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/pytest_gpu_dev.txt
( timeout 300 python tools/time_k2a.py; timeout 300 python tools/time_k1.py 40 GN_X=0 ) > $OUT/k2a_tma_r02j.txt 2>&1
cat $OUT/pytest_gpu_dev.txt | tail -25; cat $OUT/k2a_tma_r02j.txt
