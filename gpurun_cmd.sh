cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 > gpurun_out/t1_kernels.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train13.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu1.log 2>&1
for f in gpurun_out/t1_kernels.log gpurun_out/t4_parity.log; do tail -n 5 $f; done; cat gpurun_out/bench13.json | cut -c1-300; tail -n 3 gpurun_out/bench13.err
