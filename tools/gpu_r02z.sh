#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for t in "" _st2 _st4 ""; do
  echo "tag=[$t]"
  GN_LIB_TAG=$t timeout 600 python tools/profile_forward.py 2>&1 | grep -E "encoders GRAPH fp32 fused=True tc_conv=True|Error|error"
done > $OUT/ab_forward.txt
cat $OUT/ab_forward.txt
