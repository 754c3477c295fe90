#!/bin/bash
# compute-sanitizer passes over the small smoke scene and the K1 cross-check test (run under gpurun).  usage: tools/sanitize.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
export GN_SANITIZE=1
{
echo "# compute-sanitizer --tool memcheck : __graft_entry__.smoke() (forward + backward of the volume path, small scene)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke 2>&1 | grep -E "ERROR SUMMARY|smoke ok|Invalid|out of bounds|Error" | head -20
echo "# compute-sanitizer --tool memcheck : K1 walking kernel vs tile kernel (volume, rays, ragged, 12 views), staged + direct stores"
GN_K1_STAGE=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_volume.py -q -k walk 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" | head -20
GN_K1_STAGE=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_volume.py -q -k walk 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" | head -20
echo "# compute-sanitizer --tool racecheck : K1 kernels (shared-memory hazards between phase A / phase B / bulk store)"
GN_K1_STAGE=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_volume.py -q -k "walk and (rays_ragged or views12)" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard" | head -20
echo "# compute-sanitizer --tool memcheck : RGB head training (ray-mode K1 / K2a forward + reverse kernels)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_render_backward.py -q -k gradients 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" | head -20
} > $OUT/sanitizer_$TAG.txt 2>&1
cat $OUT/sanitizer_$TAG.txt
