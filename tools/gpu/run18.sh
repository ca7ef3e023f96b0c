# round 2, session 2, call 3: the defaults as built (k_scatter3, adaptive single batch) timed against their neighbours, GPU tests, bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/variants.jsonl
timeout 150 python tools/gpu_variants.py 1024 3 third > gpurun_out/variants3.log 2>&1; echo "variants rc=$?"; tail -2 gpurun_out/variants3.log | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/variants.jsonl"):
    r = json.loads(l)
    if r.get("status") != "started":
        print(r["env"], r.get("ok"), r.get("ms_per_step"), r.get("scatter_ms_per_step"), r.get("sort_ms_per_step"), r.get("scatter_GBps"), r.get("error", ""))
PY
timeout 170 python bench.py --steps 3 --warmup 3 --stage-times --no-decode > gpurun_out/bench_s4_text.json 2> gpurun_out/bench_s4_text.err; tail -3 gpurun_out/bench_s4_text.err
grep "^{" gpurun_out/bench_s4_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['sort_stage']['frac'], d['parity']['timed_output_equals_oracle_golden'], d['parity']['device_verify'], d.get('stage_ms'))"
timeout 330 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -8
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
