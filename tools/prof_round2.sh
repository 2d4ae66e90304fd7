set -x
B="python bench.py --steps 2 --warmup 3 --spp-per-step 8 --no-cpu-baseline --no-strong"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_fp64.sum
$B > gpurun_out/r02_plain.log 2>&1 && ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc $?"
P="env SPP=8 REPS=1 python tools/prof_wf.py"
$P > gpurun_out/r02_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:k_wf_(confirm|shade|trace)' -s 18 -c 3 -f -o gpurun_out/r02_wf_all $P > gpurun_out/r02_ncu_full.log 2>&1
echo "full rc $?"
tail -3 gpurun_out/r02_plain2.log
