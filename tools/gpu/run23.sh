cd $GRAFT_REPO_ROOT
timeout 55 python bench.py --steps 3 --warmup 3 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_s7_text.json 2> gpurun_out/bench_s7_text.err; tail -3 gpurun_out/bench_s7_text.err
grep "^{" gpurun_out/bench_s7_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e'], d['roofline']['frac'], d['roofline']['kernel'][:40], d['roofline']['sort_stage']['frac'], d['parity']['timed_output_equals_oracle_golden'], d['parity']['device_verify']['ok'])"
