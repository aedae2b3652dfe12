python -m pytest tests -m gpu -q 2>&1 | tail -12
python bench.py --no-cpu-baseline --loss photometric --steps 200 --profile > gpurun_out/r1l_bench_photometric.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r1l_bench_photometric.json").read().strip().splitlines()[-1])
print("photometric", round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "fused_path", (d.get("fused_path") or {}).get("value"), d["gpu_busy"]["kernel_ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"ssim_" -c 4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fused-path --loss photometric 2>&1 | grep -E "ssim_|gpu__time|inst_executed|sm__throughput" | head -20
