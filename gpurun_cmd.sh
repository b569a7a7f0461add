cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loss_branch or full_size" > gpurun_out/t50.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/t50.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm_fused or prepare_batch or colsum_batched" > gpurun_out/t50_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 6 gpurun_out/t50_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm_fused" > gpurun_out/t50_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 6 gpurun_out/t50_racecheck.log
