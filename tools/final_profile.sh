#!/bin/bash
# Round-2 evidence on one B200: the bench lines, the launch list and one full-set capture of a steady-state iteration.
set -x
python bench.py > gpurun_out/r02_bench_c3.log 2>&1; tail -c 600 gpurun_out/r02_bench_c3.log
python bench.py --impl reference > gpurun_out/r02_bench_ref.log 2>&1
tools/launch_list.sh r02
tools/prof_kernels.sh r02_wf_all "k_wf_(confirm|shade_hit|regen|trace)" 50 5
python bench.py --config c4 --steps 2 --warmup 1 --strong-steps 1 --no-cpu-baseline > gpurun_out/r02_bench_c4_1gpu.log 2>&1; tail -c 400 gpurun_out/r02_bench_c4_1gpu.log
python bench.py --config c5 --steps 3 --warmup 2 --strong-steps 2 --no-cpu-baseline > gpurun_out/r02_bench_c5_1gpu.log 2>&1; tail -c 400 gpurun_out/r02_bench_c5_1gpu.log
