python -m pytest tests -m gpu -q 2>&1 | tail -6
for extra in "" "--fused"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 20 --no-fused-path $extra > gpurun_out/r1i_bench_n2$extra.json 2> gpurun_out/r1i_bench_n2$extra.err || tail -c 1500 gpurun_out/r1i_bench_n2$extra.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1i_bench_n2$extra.json").read().strip().splitlines()[-1])
print("N=2 $extra", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
python bench.py --no-cpu-baseline --steps 300 --profile > gpurun_out/r1i_bench_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r1i_bench_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "fused", d["fused_path"]["value"], "launches", d["gpu_launches"], d["gpu_busy"]["kernel_ms_per_step"], d["roofline"]["issue"])
PY
