cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 -x > gpurun_out/t1_kernels.log 2>&1
tail -n 25 gpurun_out/t1_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 -x > gpurun_out/t4_parity.log 2>&1
tail -n 8 gpurun_out/t4_parity.log
ST_TC_EPI=1 timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes23_epi1.txt 2>&1
ST_TC_EPI=0 timeout -s KILL 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes23_epi0.txt 2>&1
head -n 1 gpurun_out/gemm_shapes23_epi1.txt gpurun_out/gemm_shapes23_epi0.txt
ST_TC_EPI=1 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench23_epi1.json 2> gpurun_out/bench23_epi1.err
ST_TC_EPI=0 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench23_epi0.json 2> gpurun_out/bench23_epi0.err
cat gpurun_out/bench23_*.json | cut -c1-200; tail -n 3 gpurun_out/bench23_*.err
for k in 16; do ST_GN_STATS_DEPTH=$k timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench23_depth$k.txt 2>&1; done
ST_GN_STATS_SWAP=1 timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench23_swap.txt 2>&1
timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench23_default.txt 2>&1
cut -c1-30 gpurun_out/gn_bench23_default.txt | head -n 8; cut -c1-30 gpurun_out/gn_bench23_depth16.txt | head -n 8; cut -c1-30 gpurun_out/gn_bench23_swap.txt | head -n 8
