#!/bin/bash
# One gpurun call that produces a round's evidence (run from the repo root on a 1-GPU box):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh r2a'
# -> full GPU test suite, the bench lines (default / photometric loss / fused operator), the ncu launch list of a train
#    step and `ncu --set full` captures of the blend kernels and the small kernels, all under gpurun_out/ with the given
#    tag; summarise them into profiles/ with tools/ncu_summary.py, tools/ncu_regions.py, tools/launch_summary.py.
TAG=${1:-r1k}
python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -c 1500 gpurun_out/${TAG}_bench.err
python bench.py --no-cpu-baseline --loss photometric --steps 200 > gpurun_out/${TAG}_bench_photometric.json 2>/dev/null
python bench.py --no-cpu-baseline --fused --steps 200 > gpurun_out/${TAG}_bench_fused.json 2>/dev/null
python - <<PY
import json
for f in ("${TAG}_bench", "${TAG}_bench_photometric", "${TAG}_bench_fused"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "fused_path", (d.get("fused_path") or {}).get("value"), "launches", d["gpu_launches"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k: v["ms"] for k, v in d["kernels"].items()})
print(d["roofline"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fused-path > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_backward_kernel -c 1 -o gpurun_out/prof_bwd_${TAG} -f python tools/blend_probe.py --reps 1 --what bwd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_forward_kernel -c 1 -o gpurun_out/prof_fwd_${TAG} -f python tools/blend_probe.py --reps 1 --what fwd > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"ssim_|adam_kernel|l1_loss|cull_chunks" -c 12 -o gpurun_out/prof_small_${TAG} -f python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-fused-path --loss photometric > /dev/null 2>&1
ls -la gpurun_out/*${TAG}* | head -20
