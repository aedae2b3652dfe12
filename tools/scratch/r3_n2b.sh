TAG=r3n2b
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"
tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("N=2", round(d["value"],1), round(d["ms_per_step"],4), d.get("step_ms"), "e2e", round(d["e2e"]["value"],1), d.get("warmup_run"), d["details"]["trainer_status"])
PY
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-ref-gpu > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_n1.json").read().strip().splitlines()[-1])
print("N=1", round(d["value"],1), round(d["ms_per_step"],4), d.get("step_ms"), "e2e", round(d["e2e"]["value"],1), d.get("warmup_run"), d["details"]["trainer_status"], d["cpu_baseline"])
PY
