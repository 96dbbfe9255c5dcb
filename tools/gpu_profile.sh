#!/bin/bash
# Runs on the GPU box (under gpurun): default bench, then the ncu launch list, the DRAM traffic of one forward pass and
# one full capture of the top kernels.  Every ncu pass follows a plain run of the same command that exited 0.
# usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-rX}
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu"
ONE="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
# DRAM bytes of every conv launch (the summariser keeps the last forward pass of the capture)
$ONE > gpurun_out/${TAG}_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'conv_' -c 200 --csv --log-file gpurun_out/${TAG}_traffic.csv $ONE > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu traffic exit $?"
# full captures: the 50-layer persistent launch (the dominant kernel), the stem, the pixel-pair Cin = 32 layer
ncu --set full --clock-control none --import-source on -k regex:'conv_chain_kernel' -s 9 -c 1 -o gpurun_out/${TAG}_prof_chain $ONE > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu chain exit $?"
ncu --set full --clock-control none --import-source on -k regex:"conv_stem_band_kernel|conv_tc_kernel" -s 20 -c 4 -o gpurun_out/${TAG}_prof_first $ONE > gpurun_out/${TAG}_ncu5.log 2>&1
echo "ncu first exit $?"
ncu --set full --clock-control none --import-source on -k regex:'decode_kernel|nms_kernel' -s 6 -c 2 -o gpurun_out/${TAG}_prof_post $ONE > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu post exit $?"
ls -la gpurun_out | tail -14
