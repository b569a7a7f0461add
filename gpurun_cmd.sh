cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "predictors_and_correctors" > gpurun_out/t61.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/t61.log
