cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 > gpurun_out/t1_kernels.log 2>&1
tail -n 5 gpurun_out/t1_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 > gpurun_out/t4_parity.log 2>&1
tail -n 4 gpurun_out/t4_parity.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench31.json 2> gpurun_out/bench31.err
cut -c1-200 gpurun_out/bench31.json; tail -n 3 gpurun_out/bench31.err
ST_FUSE_CSUM=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench31_nofuse.json 2> gpurun_out/bench31_nofuse.err
cut -c1-200 gpurun_out/bench31_nofuse.json
