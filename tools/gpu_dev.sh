#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $OUT/pytest_gpu_dev.txt
( timeout 600 python tools/time_forward_engine.py 2>&1 | grep -E "slots=(3|4) " ) > $OUT/forward_engine_dev.txt
cat $OUT/pytest_gpu_dev.txt $OUT/forward_engine_dev.txt
