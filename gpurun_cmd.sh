# GPU run 24 (one B200): reconstruction-term parity + sanity of the default path
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "reconstruction or likelihood_weighted or step_fn" --timeout=300 > gpurun_out/t_recon.log 2>&1; echo "recon tests rc=$?"; grep -E "^E  |passed|failed|Error" gpurun_out/t_recon.log | head -20
timeout 400 python bench.py --no-cpu-baseline --no-gpu-reference --steps 10 --warmup 3 > gpurun_out/bench_sanity.json 2> gpurun_out/bench_sanity.err; echo "bench rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench_sanity.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])")"
