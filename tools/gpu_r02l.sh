#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^conv |passed|failed|FAILED|Error|assert" | tail -40 ) > $OUT/pytest_gpu_dev.txt
( timeout 300 python tools/time_k2a.py ) > $OUT/k2a_time_dev.txt 2>&1
( timeout 600 python tools/profile_forward.py 2>&1 | grep -E "encoders|vgn|sample_volume|depth-mean|Error|error" ) > $OUT/profile_forward_dev.txt
cat $OUT/pytest_gpu_dev.txt $OUT/k2a_time_dev.txt $OUT/profile_forward_dev.txt
