cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sample-steps 1000 > gpurun_out/bench56_n1000.json 2> gpurun_out/bench56_n1000.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench56_n1000.json'));print(d['ms_per_step'], d['value']); print(d['sampler'])"
