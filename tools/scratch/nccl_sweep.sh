N=${1:-4}
run() { tag=$1; op=$2; shift; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/micro/nccl_ar_probe.py --tag $tag --op $op 2>gpurun_out/nccl_$tag.err | grep '^{' >> gpurun_out/nccl_sweep2_n$N.jsonl; }
rm -f gpurun_out/nccl_sweep2_n$N.jsonl
run default_avg avg FOO=1
run default_sum sum FOO=1
run nvls_sum sum NCCL_ALGO=NVLS
python - <<PY
import json
for l in open("gpurun_out/nccl_sweep2_n$N.jsonl"):
    d=json.loads(l); print(d["tag"], {k:(d[k]["us"], d[k]["busbw_GBs"]) for k in d if isinstance(d[k], dict) and "us" in d[k]})
PY
tail -3 gpurun_out/nccl_nvls_sum.err
