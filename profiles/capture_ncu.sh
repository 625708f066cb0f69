#!/bin/bash
# Runs on the GPU box (gpurun): ncu --set full of one frame of every kernel of a workload (tests/gpu_profile_r2.py brackets
# the frame with cudaProfilerStart/Stop), after the same command has exited 0 without ncu.  The report stays on the box;
# what comes back is its raw page (all metrics per launch) and its source page (per-instruction samples and stall
# reasons, SASS + line info) as CSV.   usage: profiles/capture_ncu.sh <tag> <synth1m|blub4k|bob1080|build>
tag=$1; what=${2:-synth1m}
out=gpurun_out
python tests/gpu_profile_r2.py $what > $out/${tag}_prof_plain_${what}.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_prof_plain_${what}.log; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/${tag}_${what} \
    python tests/gpu_profile_r2.py $what > $out/${tag}_ncu_${what}.log 2>&1
ncu -i /tmp/${tag}_${what}.ncu-rep --page raw --csv > $out/${tag}_prof_${what}.raw.csv
ncu -i /tmp/${tag}_${what}.ncu-rep --page source --csv 2> /dev/null | gzip > $out/${tag}_prof_${what}.source.csv.gz
ls -la /tmp/${tag}_${what}.ncu-rep $out/${tag}_prof_${what}.*
tail -2 $out/${tag}_prof_plain_${what}.log
