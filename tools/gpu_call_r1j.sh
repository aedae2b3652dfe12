python -m pytest tests/test_gpu_losses.py tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -15
run() { tag=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 20 --no-fused-path "$@" > gpurun_out/r1j_$tag.json 2> gpurun_out/r1j_$tag.err || tail -c 1500 gpurun_out/r1j_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r1j_$tag.json").read().strip().splitlines()[-1])
print("N=2 $tag", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
}
run flat1 --sh-chunks 1 --optimizer b200
run torch1 --sh-chunks 1 --optimizer torch
run flat1_again --sh-chunks 1 --optimizer b200
run flat1_fused --sh-chunks 1 --optimizer b200 --fused
python bench.py --no-cpu-baseline --steps 300 > gpurun_out/r1j_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r1j_n1.json").read().strip().splitlines()[-1])
print("N=1", round(d["value"],1), d["ms_per_step"], "e2e", d["e2e"]["value"], "fused", d["fused_path"]["value"])
PY
