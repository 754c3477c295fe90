#!/bin/bash
# round-2 experiment A: K1 staged bulk stores, K2a FHFMA split + packed pooling
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest_gpu_r02a.txt
( timeout 300 python tools/time_k1.py 40 ) > $OUT/k1_ab_r02a.txt 2>&1
( timeout 300 python tools/time_k2a.py; GN_LIB_TAG=_cvt timeout 300 python tools/time_k2a.py; timeout 300 python tools/time_k2a.py ) > $OUT/k2a_ab_r02a.txt 2>&1
cat $OUT/pytest_gpu_r02a.txt $OUT/k1_ab_r02a.txt $OUT/k2a_ab_r02a.txt
