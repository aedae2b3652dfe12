TAG=r2y
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_ref_cuda_pin.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.txt
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "c2 or c3_rs10" > gpurun_out/${TAG}_pytest_full.txt 2>&1; echo "pytest fullsize rc=$?"; tail -3 gpurun_out/${TAG}_pytest_full.txt
AB_CONFIGS="c2 c3_rs10" timeout 300 bash tools/gpu_ab.sh ${TAG} default vote > /dev/null 2>&1; cat gpurun_out/${TAG}_ab.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -c 1500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), round(d["ms_per_step"],4), d.get("step_ms"), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "other", (d.get("other_operators") or {}).get("value"))
print({k: v["ms"] for k, v in d["kernels"].items()})
print(d.get("ref_gpu",{}).get("ratio_vs_cuda_projection"), d.get("cpu_baseline"))
PY
