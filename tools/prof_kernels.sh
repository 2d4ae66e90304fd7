#!/bin/bash
# usage: tools/prof_kernels.sh <tag> <kernel-regex> <skip> <count>  — one `ncu --set full` capture of the matching launches
# of an 8-spp wavefront render (C3) -> gpurun_out/<tag>.ncu-rep
tag=$1; k=$2; skip=$3; cnt=$4
env SPP=8 REPS=1 python tools/prof_wf.py > gpurun_out/${tag}_plain.log 2>&1 && \
env SPP=8 REPS=1 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c $cnt -f -o gpurun_out/$tag python tools/prof_wf.py > gpurun_out/${tag}_ncu.log 2>&1
echo "$tag rc $?"
