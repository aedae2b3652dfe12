TAG=r2x
mkdir -p gpurun_out
tools/micro/rcp_check; echo "rcp_check rc=$?"
AB_CONFIGS="c2" timeout 300 bash tools/gpu_ab.sh ${TAG} default vote nosel fvisit default > /dev/null 2>&1; cat gpurun_out/${TAG}_ab.txt
AB_CONFIGS="c4" timeout 300 bash tools/gpu_ab.sh ${TAG}c4 default vote > /dev/null 2>&1; cat gpurun_out/${TAG}c4_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_ref_cuda_pin.py -m gpu -q -x > gpurun_out/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.txt
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "c2 or c3_rs10" > gpurun_out/${TAG}_pytest_full.txt 2>&1; echo "pytest fullsize rc=$?"; tail -3 gpurun_out/${TAG}_pytest_full.txt
B200SPLAT_LIB=$PWD/3dgs-deblur_b200/gsplat/lib/libb200splat_nosel.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "not c4 and not c3" > gpurun_out/${TAG}_pytest_nosel.txt 2>&1; echo "pytest nosel rc=$?"; tail -3 gpurun_out/${TAG}_pytest_nosel.txt
