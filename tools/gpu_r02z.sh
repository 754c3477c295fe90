#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $OUT/pytest_gpu_dev.txt
for t in _base "" _base ""; do
  echo "tag=[$t]"
  GN_LIB_TAG=$t timeout 600 python tools/profile_forward.py 2>&1 | grep -E "encoders GRAPH fp32 fused=True tc_conv=True|Error|error"
done > $OUT/ab_forward.txt
( timeout 600 python tools/time_forward_engine.py 2>&1 | grep -E "slots=(3|4) " ) >> $OUT/ab_forward.txt
cat $OUT/pytest_gpu_dev.txt $OUT/ab_forward.txt
