python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for extra in "" "--fused"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 20 --no-fused-path $extra > gpurun_out/r1h_bench_n2$extra.json 2> gpurun_out/r1h_bench_n2$extra.err || tail -c 1500 gpurun_out/r1h_bench_n2$extra.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1h_bench_n2$extra.json").read().strip().splitlines()[-1])
print("N=2 $extra", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
python bench.py --no-cpu-baseline --steps 300 > gpurun_out/r1h_bench_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r1h_bench_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "fused", d["fused_path"]["value"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 | tail -c 600
