#!/bin/bash
TAG=r02h; OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > $OUT/pytest_gpu_$TAG.txt
( for t in "" _l2 _pf _l2pf ""; do GN_LIB_TAG=$t timeout 300 python tools/time_volume.py 1 40 tc 2>&1 | tail -1 | sed "s/^/[$t] /"; done ) > $OUT/l2_ab_$TAG.txt
( timeout 600 python tools/profile_forward.py 2>&1 | grep -E "encoders|vgn|sample_volume|depth-mean" ) > $OUT/profile_forward_$TAG.txt
cat $OUT/pytest_gpu_$TAG.txt | tail -15; cat $OUT/l2_ab_$TAG.txt $OUT/profile_forward_$TAG.txt
