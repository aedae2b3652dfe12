TAG=r2w
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_backward_kernel2 -c 1 -o gpurun_out/prof_bwd_${TAG} -f python tools/blend_probe.py --reps 1 --what bwd > gpurun_out/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_forward_kernel2 -c 1 -o gpurun_out/prof_fwd_${TAG} -f python tools/blend_probe.py --reps 1 --what fwd > gpurun_out/${TAG}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
ls -la gpurun_out/prof_*_${TAG}.ncu-rep
