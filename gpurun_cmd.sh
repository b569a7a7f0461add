# GPU run 25 (one B200): sampler-mode lines with gpu_launches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python bench.py --config c2 --mode sampler --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_sampler_c2_n1000.json 2> gpurun_out/r02_bench_sampler_c2_n1000.err; echo "c2 sampler rc=$?"; python -c "import json;d=json.loads(open('gpurun_out/r02_bench_sampler_c2_n1000.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks']['sm_mhz'])"
timeout 900 python bench.py --config c5 --mode sampler --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_sampler_c5_n2000.json 2> gpurun_out/r02_bench_sampler_c5_n2000.err; echo "c5 sampler rc=$?"; python -c "import json;d=json.loads(open('gpurun_out/r02_bench_sampler_c5_n2000.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks']['sm_mhz'])"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "sampler" --timeout=200 > gpurun_out/t_samp.log 2>&1; echo "sampler tests rc=$?"; tail -n 2 gpurun_out/t_samp.log
