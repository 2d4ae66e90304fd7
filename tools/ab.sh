#!/bin/bash
# usage: tools/ab.sh <out.log> <spp> name[:ENV=VAL,...] ...   — renders the C3 scene with each variant build
# (build/variants/<name>.so, or the in-tree library for "tree") and appends "name: Mseg/s" lines to <out.log>.
out=$1; spp=$2; shift 2
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  lib=build/variants/$name.so; [ "$name" = tree ] && lib=cpu-raytracing-rt_b200/librt_b200.so
  echo "== $spec" >> $out
  env $envs RT_B200_LIB=$lib SPP=$spp REPS=3 timeout 300 python tools/prof_wf.py 2>&1 | tail -2 >> $out
done
