cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 > gpurun_out/t1_kernels.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 > gpurun_out/t4_parity_auto.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train5.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu1.log 2>&1
GP_CASE=fwd128 ST_TC_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_v2_fwd128 python tools/gemm_probe.py > gpurun_out/ncu_a.log 2>&1
GP_CASE=dgrad128 ST_TC_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_v2_dgrad128 python tools/gemm_probe.py > gpurun_out/ncu_b.log 2>&1
GP_CASE=fwd128 ST_TC_VARIANT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_v1_fwd128 python tools/gemm_probe.py > gpurun_out/ncu_c.log 2>&1
GP_CASE=fwd256 ST_TC_VARIANT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_v2_fwd256 python tools/gemm_probe.py > gpurun_out/ncu_d.log 2>&1
for f in gpurun_out/t1_kernels.log gpurun_out/t4_parity_auto.log; do tail -n 3 $f; done; cat gpurun_out/bench5.json | cut -c1-300; tail -n 3 gpurun_out/bench5.err
