cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_verify.py -q -m gpu -x 2>&1 | tail -25
timeout 600 python bench.py --steps 3 --warmup 2 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2i_text.json 2> gpurun_out/bench_r2i_text.err; tail -3 gpurun_out/bench_r2i_text.err; grep "^{" gpurun_out/bench_r2i_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['parity'])"
