#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $OUT/pytest_gpu_dev.txt
( timeout 600 python tools/time_forward_engine.py 2>&1 | tail -12 ) > $OUT/forward_engine_dev.txt
( timeout 600 python tools/profile_forward.py 2>&1 | grep -E "encoders GRAPH|vgn|sample_volume|depth-mean|Error|error" ) > $OUT/profile_forward_dev.txt
cat $OUT/pytest_gpu_dev.txt $OUT/forward_engine_dev.txt $OUT/profile_forward_dev.txt
