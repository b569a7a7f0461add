cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ST_PDL=1 timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 --durations=8 > gpurun_out/t1_pdl1.log 2>&1
tail -n 14 gpurun_out/t1_pdl1.log
ST_PDL=0 timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 --durations=4 > gpurun_out/t1_pdl0.log 2>&1
tail -n 8 gpurun_out/t1_pdl0.log
