#!/bin/bash
# compute-sanitizer memcheck over __graft_entry__.smoke(): forward + backward of the volume path and one planner call (encoders K6 / K7,
# layout glue, K1 / K2a / K2b, depth-mean head, VGN K5, grasp post-processing K4) on the small scene.   usage: tools/sanitize.sh <tag> [seconds]
# (a fresh box needs ~1 min before python makes its first CUDA call: hence --launch-timeout; the round-2 attempt ran out of GPU budget
# before a result - the last sanitizer evidence is profiles/sanitizer_r01e.txt)
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
{
echo "# compute-sanitizer --tool memcheck : __graft_entry__.smoke()"
timeout ${2:-300} compute-sanitizer --tool memcheck --target-processes all --launch-timeout 600 --print-limit 20 python __graft_entry__.py --smoke 2>&1 | grep -vE "^=========     (at|in|by) |Host Frame|Device Frame" | tail -25
} > $OUT/sanitizer_$TAG.txt 2>&1
cat $OUT/sanitizer_$TAG.txt
