# GPU validation run (one B200): CTA-pair GEMM check first, then tests, GEMM shape bench, bench with reference legs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/test_stats.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "bench_shapes" --timeout=300 > gpurun_out/t_cg2.log 2>&1; CG2=$?; echo "cg2 check rc=$CG2"; tail -n 5 gpurun_out/t_cg2.log
if [ $CG2 -ne 0 ]; then export ST_TC_CG=1; echo "falling back to ST_TC_CG=1 for the rest"; fi
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/t_gpu.log
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench_r2.txt 2>&1; echo "gemm_bench rc=$?"; cat gpurun_out/gemm_bench_r2.txt
timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r2b.json
ST_TC_CG=1 timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_r2b_cg1.json 2> gpurun_out/bench_r2b_cg1.err; echo "cg1 rc=$?"; cut -c1-300 gpurun_out/bench_r2b_cg1.json
ST_PDL_CAPTURE=0 timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_r2b_nopdlcap.json 2> gpurun_out/bench_r2b_nopdlcap.err; echo "nopdlcap rc=$?"; cut -c1-300 gpurun_out/bench_r2b_nopdlcap.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2b_ref.json 2> gpurun_out/bench_r2b_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_r2b_ref.json
