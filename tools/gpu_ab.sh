#!/bin/bash
# A/B of library builds on one box: bash tools/gpu_ab.sh <tag> <variant> [<variant> ...]
# (variants are built here with B200_BUILD_VARIANT=<v> B200_NVCC_FLAGS=... python 3dgs-deblur_b200/build.py; "" = default)
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=3dgs-deblur_b200/gsplat/lib/libb200splat${v:+_$v}.so
  [ "$v" = "default" ] && lib=3dgs-deblur_b200/gsplat/lib/libb200splat.so
  for cfg in c2 c4; do
    echo "== $v $cfg" | tee -a gpurun_out/${TAG}_ab.txt
    B200SPLAT_LIB=$PWD/$lib python tools/blend_probe.py --config $cfg --reps 30 2>&1 | tail -3 | tee -a gpurun_out/${TAG}_ab.txt
  done
done
