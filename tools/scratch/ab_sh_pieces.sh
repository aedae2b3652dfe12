for c in 8 1 16 4; do
  B200_SH_UPDATE_PIECES=$c python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-fused-path > gpurun_out/r2r_$c.json 2> gpurun_out/r2r_$c.err || tail -c 1000 gpurun_out/r2r_$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_$c.json").read().strip().splitlines()[-1])
print("pieces $c", round(d["value"],1), round(d["ms_per_step"],4), d["step_ms"]["p50"], "e2e", round(d["e2e"]["value"],1))
PY
done
B200_SH_UPDATE_PIECES=8 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-ref-gpu --no-fused-path --timeline gpurun_out/r2r_timeline_n1.tsv > /dev/null 2>&1
