#!/bin/bash
# usage: tools/launch_list.sh <tag> [lib]  — per-launch time / DRAM bytes / lanes of one 8-spp wavefront render (C3) -> gpurun_out/<tag>_ll.csv
tag=$1; lib=${2:-cpu-raytracing-rt_b200/librt_b200.so}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
env RT_B200_LIB=$lib SPP=8 REPS=1 python tools/prof_wf.py > gpurun_out/${tag}_plain.log 2>&1 && \
env RT_B200_LIB=$lib SPP=8 REPS=1 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_ll.csv python tools/prof_wf.py > gpurun_out/${tag}_ncu.log 2>&1
echo "$tag rc $?"
