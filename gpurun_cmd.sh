cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 > gpurun_out/t1_kernels.log 2>&1
tail -n 5 gpurun_out/t1_kernels.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity.log 2>&1
tail -n 5 gpurun_out/t4_parity.log
timeout 600 python tools/gn_bench.py > gpurun_out/gn_bench19.txt 2>&1
cat gpurun_out/gn_bench19.txt
timeout 600 python tools/gemm_shapes.py > gpurun_out/gemm_shapes19.txt 2>&1
head -n 45 gpurun_out/gemm_shapes19.txt
ST_GN_CSUM_OCC=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench19_occ3.json 2> gpurun_out/bench19_occ3.err
ST_GN_CSUM_OCC=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench19_occ2.json 2> gpurun_out/bench19_occ2.err
cat gpurun_out/bench19_*.json | cut -c1-200; tail -n 3 gpurun_out/bench19_*.err
