python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --profile > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err || tail -c 1500 gpurun_out/r1e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r1e_bench.json").read().strip().splitlines()[-1])
print("dropin", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "fused_path", d.get("fused_path"))
print(json.dumps(d.get("gpu_busy"), indent=0))
print({k: v["ms"] for k, v in d["kernels"].items()})
PY
for c in c3_rs c3_rs10 c4; do
  python bench.py --config $c --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/r1e_bench_$c.json 2> gpurun_out/r1e_bench_$c.err || tail -c 800 gpurun_out/r1e_bench_$c.err
  python bench.py --impl refgpu --config $c --steps 8 --warmup 3 > gpurun_out/r1e_refgpu_$c.json 2> gpurun_out/r1e_refgpu_$c.err || tail -c 800 gpurun_out/r1e_refgpu_$c.err
  python - <<PY
import json
for f in ("gpurun_out/r1e_bench_$c.json", "gpurun_out/r1e_refgpu_$c.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d.get("fused_path", {}).get("value") if isinstance(d.get("fused_path"), dict) else None)
    except Exception as e: print(f, "ERR", e)
PY
done
python bench.py --impl refgpu --steps 30 --warmup 5 > gpurun_out/r1e_refgpu_c2.json 2>/dev/null; tail -c 400 gpurun_out/r1e_refgpu_c2.json
