python -m pytest tests/test_gpu_pipeline.py -q -x 2>&1 | tail -3
for c in 2 0 1 4; do
  B200_SH_UPDATE_CTAS=$c python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-fused-path > gpurun_out/r2q_$c.json 2> gpurun_out/r2q_$c.err || tail -c 1000 gpurun_out/r2q_$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_$c.json").read().strip().splitlines()[-1])
print("ctas $c", round(d["value"],1), round(d["ms_per_step"],4), d["step_ms"]["p50"], "e2e", round(d["e2e"]["value"],1))
PY
done
B200_SH_UPDATE_CTAS=2 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-fused-path --timeline gpurun_out/r2q_timeline_n1.tsv > /dev/null 2>&1
