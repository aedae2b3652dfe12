TAG=r3n8
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("N=8", round(d["value"],1), round(d["ms_per_step"],4), d.get("step_ms"), "e2e", round(d["e2e"]["value"],1), d.get("balance"))
PY
