cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cs in 2 4 1; do
ST_TC_CLUSTER=$cs timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "tcgen05" --timeout 300 > gpurun_out/t2_tc_cs$cs.log 2>&1
done
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity_auto.log 2>&1
for cs in 2 4; do
ST_TC_CLUSTER=$cs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_cs$cs.json 2> gpurun_out/bench3_cs$cs.err
done
ST_TC_VARIANT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_v1.json 2> gpurun_out/bench3_v1.err
for f in gpurun_out/t2_tc_cs*.log gpurun_out/t4_parity_auto.log; do tail -n 3 $f; done; cat gpurun_out/bench3_*.json | cut -c1-1600; tail -n 3 gpurun_out/bench3_*.err
