#!/bin/bash
# One gpurun call that produces a round's evidence (run from the repo root on a 1-GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh r2a'
# -> the driver's own entry point (__graft_entry__.smoke), the full GPU test suite, the default bench line, the ncu launch
#    list of a train step and `ncu --set full` captures of the blend kernels, all under gpurun_out/ with the given tag;
#    summarise them into profiles/ with tools/ncu_summary.py, tools/ncu_regions.py, tools/launch_summary.py.
# Optional second argument: "quick" skips the ncu captures.
TAG=${1:-r2a}
MODE=${2:-full}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.txt
python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.txt
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -c 1500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), round(d["ms_per_step"],4), d.get("step_ms"), "e2e", round(d["e2e"]["value"],1), "other_ops", (d.get("other_operators") or {}).get("value"), "launches", d["gpu_launches"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
print({k: v["ms"] for k, v in d["kernels"].items()})
print(d["roofline"])
PY
if [ "$MODE" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fused-path --no-ref-gpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_backward_kernel -c 1 -o gpurun_out/prof_bwd_${TAG} -f python tools/blend_probe.py --reps 1 --what bwd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_forward_kernel -c 1 -o gpurun_out/prof_fwd_${TAG} -f python tools/blend_probe.py --reps 1 --what fwd > /dev/null 2>&1
fi
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_refarm.json 2> gpurun_out/${TAG}_refarm.err; tail -c 400 gpurun_out/${TAG}_refarm.json
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-fused-path --timeline gpurun_out/${TAG}_timeline_n1.tsv > gpurun_out/${TAG}_bench_tl.json 2>/dev/null
ls -la gpurun_out/*${TAG}* | head -20
