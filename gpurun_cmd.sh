cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --sample-steps 10 > gpurun_out/bench60.json 2> gpurun_out/bench60.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench60.json'));print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e_u8']['ms_per_step'], d['clocks'], d['cpu_baseline'])"; tail -n 3 gpurun_out/bench60.err
