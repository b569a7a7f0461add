cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 600 > gpurun_out/t1_kernels.log 2>&1
ST_GEMM_BACKEND=simt timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t3_parity_simt.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity_auto.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/t5_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 4 -o gpurun_out/prof_gemm_tc python tools/profile_step.py --batch 512 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/t*.log; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
