#!/bin/bash
# One GPU-box pass that produces every artefact profiles/ cites.  usage: tools/gpu_measure.sh <tag>   (run under gpurun)
# Numbers printed by the commands running under ncu are never bench values; bench.py is run separately, un-profiled.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $OUT/pytest_gpu_$TAG.txt
( timeout 300 python tools/time_k2a.py ) > $OUT/k2a_time_$TAG.txt 2>&1
( timeout 600 python tools/time_forward_engine.py 2>&1 | tail -8 ) > $OUT/forward_engine_$TAG.txt
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --train-batch 0 --highres-scenes 0 --skip-full > /dev/null 2>&1
# launch list of ONE eager full forward (encoders K6/K7 + hot path + VGN K5 + post K4): kernel shares of the planner's call
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 900 --csv --log-file $OUT/launches_forward_$TAG.csv \
    python tools/forward_once.py > /dev/null 2>&1
for k in gn_k1_kernel gn_k2a_tc3_kernel gn_k2b_attn_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${k}_full_$TAG \
      python tools/time_volume.py 1 2 tc > /dev/null 2>&1
done
# encoder / VGN kernels: one launch each out of an eager full forward (the 40th K7 launch is a 32-channel 144x256 layer)
for k in gn_k7_conv_kernel gn_k6_norm_act_pad_kernel gn_k5_conv_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o $OUT/${k}_full_$TAG \
      python tools/forward_once.py > /dev/null 2>&1
done
( timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > $OUT/smoke_$TAG.txt; cat $OUT/smoke_$TAG.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.csv
cat $OUT/pytest_gpu_$TAG.txt $OUT/k2a_time_$TAG.txt; python tools/show_bench.py $OUT/bench_$TAG.json; head -c 600 $OUT/bench_ref_$TAG.json
ls -la $OUT | tail -8
