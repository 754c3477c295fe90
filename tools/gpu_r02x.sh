#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 900 --csv --log-file $OUT/launches_forward_dev.csv \
    python tools/forward_once.py > /dev/null 2>&1
( timeout 600 python tools/profile_train.py 2>&1 | cut -c1-200 | head -120 ) > $OUT/profile_train_dev.txt
ls -la $OUT | tail -4
