cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sampler" > gpurun_out/t44_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/t44_gpu.log
SWEEP_TIMEOUT=200 timeout 600 python tools/config_sweep.py c5 c3 > gpurun_out/sweep44.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep44.txt
