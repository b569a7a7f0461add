# GPU run 23 (one B200): final validation + evidence of round 2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/test_stats.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/t_gpu_final.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/t_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_1gpu.json
for c in c4 c3 c5 deepest; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; echo "$c rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_bench_$c.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], (d.get('roofline') or {}).get('whole_step_frac'))")"; grep -c "capture of the training step failed" gpurun_out/r02_bench_$c.err
done
timeout 600 python bench.py --config c2 --mode sampler --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_sampler_c2_n1000.json 2> gpurun_out/r02_bench_sampler_c2_n1000.err; echo "c2 sampler rc=$?"; cut -c1-230 gpurun_out/r02_bench_sampler_c2_n1000.json
timeout 900 python bench.py --config c5 --mode sampler --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_sampler_c5_n2000.json 2> gpurun_out/r02_bench_sampler_c5_n2000.err; echo "c5 sampler rc=$?"; cut -c1-230 gpurun_out/r02_bench_sampler_c5_n2000.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_t.csv python tools/profile_step.py --batch 512 > gpurun_out/ncu_t.log 2>&1; echo "ncu train rc=$?"
python tools/summarize_launches.py gpurun_out/launches_t.csv > gpurun_out/r02_launches_train_step.md; head -8 gpurun_out/r02_launches_train_step.md; rm -f gpurun_out/launches_t.csv
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_s.csv python tools/profile_step.py --batch 1024 --mode forward > gpurun_out/ncu_s.log 2>&1; echo "ncu sampler rc=$?"
python tools/summarize_launches.py gpurun_out/launches_s.csv > gpurun_out/r02_launches_sampler_step.md; head -8 gpurun_out/r02_launches_sampler_step.md; rm -f gpurun_out/launches_s.csv
timeout 300 python tools/gemm_shapes.py > gpurun_out/r02_gemm_shapes.txt 2>&1; echo "gemm_shapes rc=$?"; head -2 gpurun_out/r02_gemm_shapes.txt
