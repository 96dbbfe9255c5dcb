#!/bin/bash
# Runs on the GPU box (under gpurun): default bench, then the ncu launch list and one full capture of the top kernels.
# usage: bash tools_gpu_profile.sh <tag>
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
$SMALL > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_tc2_kernel' -s 108 -c 2 -o gpurun_out/${TAG}_prof_conv $SMALL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu conv exit $?"
$SMALL > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'decode_kernel|nms_kernel' -s 6 -c 2 -o gpurun_out/${TAG}_prof_post $SMALL > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu post exit $?"
ls -la gpurun_out | tail -20
