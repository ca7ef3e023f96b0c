cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hosts.py tests/test_gpu_zip.py -q -m gpu -x 2>&1 | tail -6
timeout 600 python bench.py --steps 3 --warmup 2 --stage-times --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2l_text.json 2> gpurun_out/bench_r2l_text.err; tail -3 gpurun_out/bench_r2l_text.err; grep "^{" gpurun_out/bench_r2l_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['parity']['timed_output_equals_oracle_golden'], d['parity']['device_verify']['ok'], d['stage_ms'])"
timeout 600 python bench.py --config entries --entries 20000 --steps 2 --warmup 1 > gpurun_out/bench_r2l_entries.json 2> gpurun_out/bench_r2l_entries.err; tail -3 gpurun_out/bench_r2l_entries.err; grep "^{" gpurun_out/bench_r2l_entries.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['entries_per_s'], d['parity'])"
