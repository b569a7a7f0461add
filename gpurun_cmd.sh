# GPU run 16 (one B200): parallel GroupNorm prologues, cluster dsm_loss: C5 / C3 / C2 lines
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm or phase1 or dsm" --timeout=300 > gpurun_out/t_gn.log 2>&1; echo "gn tests rc=$?"; grep -E "^E  |passed|failed|Error" gpurun_out/t_gn.log | head -20
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout=600 -k "not reference_drivers and not trajectory_100" > gpurun_out/t_par.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/t_par.log
for c in c5 c3 c2; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-gpu-reference --steps 10 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['clocks'])")"
done
timeout 900 python bench.py --config c5 --mode sampler --no-cpu-baseline --no-gpu-reference --sample-steps 200 > gpurun_out/bench_c5_samp.json 2> gpurun_out/bench_c5_samp.err; echo "c5 sampler rc=$?"; cut -c1-240 gpurun_out/bench_c5_samp.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/profile_step.py --batch 16 --config ve/celebahq/uncsnpp_st > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
python tools/summarize_launches.py gpurun_out/launches_c5.csv > gpurun_out/r02_launches_train_step_c5b.md; head -24 gpurun_out/r02_launches_train_step_c5b.md
rm -f gpurun_out/launches_c5.csv
