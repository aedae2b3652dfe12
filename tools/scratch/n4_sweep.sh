N=${1:-4}
for cfg in "1 8" "2 16" "1 16"; do
  set -- $cfg; ch=$1; im=$2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --sh-chunks $ch --images $im > gpurun_out/r2t_n${N}_c${ch}_i${im}.json 2> gpurun_out/r2t_n${N}_c${ch}_i${im}.err || tail -c 800 gpurun_out/r2t_n${N}_c${ch}_i${im}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2t_n${N}_c${ch}_i${im}.json").read().strip().splitlines()[-1])
print("N$N chunks $ch images $im", round(d["value"],1), round(d["ms_per_step"],4), d["step_ms"]["p50"], "e2e", round(d["e2e"]["value"],1), d["details"]["balance"])
PY
done
