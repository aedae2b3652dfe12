#!/bin/bash
# A/B of library builds / env switches on one box:  bash tools/gpu_ab.sh <tag> <spec> [<spec> ...]
# spec = <variant>[:ENV=VAL[,ENV=VAL...]]; variant "default" = gsplat/lib/libb200splat.so, else libb200splat_<variant>.so
# (built here with B200_BUILD_VARIANT=<v> B200_NVCC_FLAGS=... python 3dgs-deblur_b200/build.py)
TAG=$1; shift
CFGS=${AB_CONFIGS:-"c2 c4"}
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  [ "$spec" != "$v" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  lib=3dgs-deblur_b200/gsplat/lib/libb200splat_$v.so
  [ "$v" = "default" ] && lib=3dgs-deblur_b200/gsplat/lib/libb200splat.so
  for cfg in $CFGS; do
    echo "== $spec $cfg" | tee -a gpurun_out/${TAG}_ab.txt
    env $envs B200SPLAT_LIB=$PWD/$lib python tools/blend_probe.py --config $cfg --reps 30 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_ab.txt
  done
done
