cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_verify.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 2 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2k_text.json 2> gpurun_out/bench_r2k_text.err; tail -3 gpurun_out/bench_r2k_text.err; grep "^{" gpurun_out/bench_r2k_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['parity']['timed_output_equals_oracle_golden'], d['parity']['device_verify'])"
