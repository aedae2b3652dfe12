TAG=r2z
mkdir -p gpurun_out
for v in default latesort default latesort; do
  lib=3dgs-deblur_b200/gsplat/lib/libb200splat.so; [ "$v" != "default" ] && lib=3dgs-deblur_b200/gsplat/lib/libb200splat_$v.so
  B200SPLAT_LIB=$PWD/$lib timeout 200 python bench.py --steps 160 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-fused-path --timeline gpurun_out/${TAG}_timeline_$v.tsv > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],1), round(d["ms_per_step"],4), d.get("step_ms",{}).get("p50"), "launches", d["gpu_launches"])
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.txt
