for extra in "" "--fused"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 300 --warmup 20 --no-fused-path $extra > gpurun_out/r1m_bench_n4$extra.json 2> gpurun_out/r1m_bench_n4$extra.err || tail -c 1500 gpurun_out/r1m_bench_n4$extra.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1m_bench_n4$extra.json").read().strip().splitlines()[-1])
print("N=4 $extra", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r1m_bench_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r1m_bench_n1.json").read().strip().splitlines()[-1])
print("N=1", round(d["value"],1), d["ms_per_step"], "fused", d["fused_path"]["value"])
PY
