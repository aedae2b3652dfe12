for v in default cc8; do
  lib=3dgs-deblur_b200/gsplat/lib/libb200splat_$v.so; [ "$v" = default ] && lib=3dgs-deblur_b200/gsplat/lib/libb200splat.so
  for rep in 1 2; do
  B200SPLAT_LIB=$PWD/$lib python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-fused-path > gpurun_out/r2p_$v.json 2> gpurun_out/r2p_$v.err || tail -c 1000 gpurun_out/r2p_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],1), round(d["ms_per_step"],4), d["step_ms"]["p50"], "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
PY
  done
done
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-fused-path --timeline gpurun_out/r2p_timeline_n1.tsv > /dev/null 2>&1
