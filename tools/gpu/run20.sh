cd $GRAFT_REPO_ROOT
timeout 100 python bench.py --steps 2 --warmup 1 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_s5_text.json 2> gpurun_out/bench_s5_text.err; tail -3 gpurun_out/bench_s5_text.err
grep "^{" gpurun_out/bench_s5_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e'], d['parity']['device_verify']['ok'])"
