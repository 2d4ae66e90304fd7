#!/bin/bash
# Round-2 evidence on an 8-GPU box: in-library NCCL reduce at 8 ranks (C3 weak + strong, C4, C5) and the CLI on all devices.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_multi_${N}gpu.log 2>&1; tail -2 gpurun_out/r02_multi_${N}gpu.log
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r02_bench_c3_${N}gpu.log 2>&1; tail -c 300 gpurun_out/r02_bench_c3_${N}gpu.log; echo
timeout 400 $TR --master-port 29522 bench.py --gpus $N --config c4 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_${N}gpu.log 2>&1; tail -c 300 gpurun_out/r02_bench_c4_${N}gpu.log; echo
B=/tmp/rtb200_bench_cache/hf_500x1000.gltf
( time env RT_DEVICES=all RT_SEED=7 cpu-raytracing-rt_b200/cpu-raytracing-rt glTF $B 1920 1080 256 /tmp/cli_all.ppm ) > gpurun_out/r02_cli_${N}gpu.log 2>&1
( time env RT_SEED=7 cpu-raytracing-rt_b200/cpu-raytracing-rt glTF $B 1920 1080 256 /tmp/cli_one.ppm ) >> gpurun_out/r02_cli_${N}gpu.log 2>&1
python - >> gpurun_out/r02_cli_${N}gpu.log 2>&1 <<'PY'
import numpy as np
def ppm(p):
    d = open(p, "rb").read(); k = 0
    for _ in range(3): k = d.index(b"\n", k) + 1
    return np.frombuffer(d[k:], np.uint8)
a, b = ppm("/tmp/cli_all.ppm"), ppm("/tmp/cli_one.ppm")
print("CLI all-devices vs one device: bytes differing", int((a != b).sum()), "of", a.size, "max abs diff", int(np.abs(a.astype(int) - b.astype(int)).max()))
PY
tail -12 gpurun_out/r02_cli_${N}gpu.log
timeout 600 $TR --master-port 29523 bench.py --gpus $N --config c5 --steps 3 --warmup 3 --strong-steps 2 > gpurun_out/r02_bench_c5_${N}gpu.log 2>&1; tail -c 300 gpurun_out/r02_bench_c5_${N}gpu.log; echo
