cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 > gpurun_out/t1_kernels.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity.log 2>&1
ST_FUSE_CSUM=0 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 -k "unet or train" > gpurun_out/t4_parity_nofuse.log 2>&1
ST_FUSE_CSUM=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_fuse.json 2> gpurun_out/bench14_fuse.err
ST_FUSE_CSUM=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_nofuse.json 2> gpurun_out/bench14_nofuse.err
for f in gpurun_out/t1_kernels.log gpurun_out/t4_parity.log gpurun_out/t4_parity_nofuse.log; do tail -n 2 $f; done; cat gpurun_out/bench14_*.json | cut -c1-220; tail -n 3 gpurun_out/bench14_*.err
