cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 -k "checkpoint" > gpurun_out/t6_ckpt.log 2>&1
tail -n 3 gpurun_out/t6_ckpt.log
timeout -s KILL 300 python tools/swap_probe.py > gpurun_out/swap_probe.txt 2>&1
cat gpurun_out/swap_probe.txt
