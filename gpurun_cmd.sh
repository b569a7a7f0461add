# GPU run 28 (one B200): FIR defaults after the occupancy sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -k "upfirdn or fir or full_width" --timeout=200 2>&1 | tail -1
timeout 200 python tools/upfirdn_bench.py > gpurun_out/r02_upfirdn_bench.txt 2>&1; cat gpurun_out/r02_upfirdn_bench.txt | cut -c1-130
timeout 300 python bench.py --config c5 --no-cpu-baseline --no-gpu-reference --steps 10 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "c5 rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_bench_c5.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])")"
timeout 300 python bench.py --config c3 --no-cpu-baseline --no-gpu-reference --steps 10 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "c3 rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r02_bench_c3.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])")"
