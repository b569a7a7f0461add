cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 > gpurun_out/t1_kernels.log 2>&1
tail -n 12 gpurun_out/t1_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 > gpurun_out/t4_parity.log 2>&1
tail -n 6 gpurun_out/t4_parity.log
ST_TC_WGRAD_NT=1 timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes33_nt1.txt 2>&1
ST_TC_WGRAD_NT=0 timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes33_nt0.txt 2>&1
head -n 1 gpurun_out/gemm_shapes33_nt1.txt gpurun_out/gemm_shapes33_nt0.txt
grep -h "^wgrad . 32 128 " gpurun_out/gemm_shapes33_nt1.txt; echo; grep -h "^wgrad . 32 128 " gpurun_out/gemm_shapes33_nt0.txt
ST_TC_WGRAD_NT=1 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench33_nt1.json 2> gpurun_out/bench33_nt1.err
ST_TC_WGRAD_NT=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench33_nt0.json 2> gpurun_out/bench33_nt0.err
for f in nt0 nt1; do cut -c1-180 gpurun_out/bench33_$f.json; done; tail -n 3 gpurun_out/bench33_*.err
