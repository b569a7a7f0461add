cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "not tcgen05" --timeout 300 > gpurun_out/t1_kernels.log 2>&1
for cs in 2 1; do
ST_TC_CLUSTER=$cs timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "tcgen05" --timeout 300 > gpurun_out/t2_tc_cs$cs.log 2>&1
done
timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity_auto.log 2>&1
for cs in 2 1; do
ST_TC_CLUSTER=$cs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_cs$cs.json 2> gpurun_out/bench4_cs$cs.err
done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train4.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu1.log 2>&1
for f in gpurun_out/t1_kernels.log gpurun_out/t2_tc_cs*.log gpurun_out/t4_parity_auto.log; do tail -n 3 $f; done; cat gpurun_out/gemm_bench.txt; cat gpurun_out/bench4_*.json | cut -c1-200; tail -n 3 gpurun_out/bench4_*.err
