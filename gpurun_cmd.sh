cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 > gpurun_out/t1_kernels.log 2>&1
tail -n 15 gpurun_out/t1_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 > gpurun_out/t4_parity.log 2>&1
tail -n 8 gpurun_out/t4_parity.log
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err
cat gpurun_out/bench24.json | cut -c1-200; tail -n 3 gpurun_out/bench24.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_train24.csv python tools/profile_step.py --batch 512 > gpurun_out/prof24.log 2>&1
tail -n 2 gpurun_out/prof24.log
