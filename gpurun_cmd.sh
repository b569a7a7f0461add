cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ST_TC_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k tcgen05 --timeout 300 > gpurun_out/t2_tc_v2.log 2>&1
timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench5.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err
ST_TC_VARIANT=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench10_v2.json 2> gpurun_out/bench10_v2.err
for f in gpurun_out/t2_tc_v2.log gpurun_out/t4_parity.log; do tail -n 3 $f; done; cat gpurun_out/gemm_bench5.txt; cat gpurun_out/bench10*.json | cut -c1-200; tail -n 3 gpurun_out/bench10*.err
