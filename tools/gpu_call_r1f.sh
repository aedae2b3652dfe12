python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --no-cpu-baseline --profile > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err || tail -c 1500 gpurun_out/r1f_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r1f_bench.json").read().strip().splitlines()[-1])
print("dropin", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "fused_path", d.get("fused_path", {}).get("value"), "launches", d["gpu_launches"])
gb=d.get("gpu_busy") or {}
print("kernel_ms", gb.get("kernel_ms_per_step"), "idle_total", gb.get("idle_total_us_per_step"))
for k,v in (gb.get("idle_before_us_per_step") or {}).items(): print("   ", v, k[:70])
print({k: v["ms"] for k, v in d["kernels"].items()})
print(d["roofline"])
PY
python bench.py --no-cpu-baseline --fused --profile > gpurun_out/r1f_bench_fused.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r1f_bench_fused.json").read().strip().splitlines()[-1])
print("fused", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
gb=d.get("gpu_busy") or {}
print("kernel_ms", gb.get("kernel_ms_per_step"), "idle_total", gb.get("idle_total_us_per_step"))
for k,v in (gb.get("idle_before_us_per_step") or {}).items(): print("   ", v, k[:70])
PY
