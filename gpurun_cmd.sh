cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 120 > gpurun_out/t1_kernels.log 2>&1
tail -n 3 gpurun_out/t1_kernels.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 > gpurun_out/t4_parity.log 2>&1
tail -n 4 gpurun_out/t4_parity.log
timeout -s KILL 600 python bench.py > gpurun_out/bench27.json 2> gpurun_out/bench27.err
cut -c1-300 gpurun_out/bench27.json; tail -n 3 gpurun_out/bench27.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench27_ref.json 2> gpurun_out/bench27_ref.err
cut -c1-300 gpurun_out/bench27_ref.json; tail -n 3 gpurun_out/bench27_ref.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_train27.csv python tools/profile_step.py --batch 512 > gpurun_out/prof27.log 2>&1
tail -n 2 gpurun_out/prof27.log
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench27.txt 2>&1
timeout 300 python tools/gemm_shapes.py > gpurun_out/gemm_shapes27.txt 2>&1
timeout 300 python tools/gn_bench.py > gpurun_out/gn_bench27.txt 2>&1
head -n 2 gpurun_out/gemm_shapes27.txt; head -n 4 gpurun_out/gn_bench27.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_ -s 40 -c 6 -f -o gpurun_out/r27_gn python tools/profile_step.py --batch 512 > gpurun_out/ncu27c.log 2>&1
GP_CASE=fwd256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r27_gemm_fwd256 python tools/gemm_probe.py > gpurun_out/ncu27a.log 2>&1
GP_CASE=nin768 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r27_gemm_nin768 python tools/gemm_probe.py > gpurun_out/ncu27b.log 2>&1
GP_CASE=attn_qk timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -f -o gpurun_out/r27_gemm_attn_qk python tools/gemm_probe.py > gpurun_out/ncu27d.log 2>&1
tail -n 1 gpurun_out/ncu27a.log gpurun_out/ncu27b.log gpurun_out/ncu27c.log gpurun_out/ncu27d.log
