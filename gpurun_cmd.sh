cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 600 > gpurun_out/t1_kernels.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity_auto.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train2.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu1.log 2>&1
for f in gpurun_out/t1_kernels.log gpurun_out/t4_parity_auto.log; do tail -n 3 $f; done; cat gpurun_out/bench2.json; tail -n 3 gpurun_out/bench2.err
