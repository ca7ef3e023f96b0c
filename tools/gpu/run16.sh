# round 2, session 2, call 1: variants (parity + time), GPU tests and bench under the best variant, ncu evidence
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/variants.jsonl gpurun_out/best_env.sh
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
for i in 1 2 3 4 5; do
  timeout 420 python tools/gpu_variants.py 1024 2 > gpurun_out/variants_run$i.log 2>&1
  rc=$?
  echo "variants run $i rc=$rc"; tail -2 gpurun_out/variants_run$i.log | cut -c1-400
  [ -f gpurun_out/best_env.sh ] && break
done
python - <<'PY'
import json
for l in open("gpurun_out/variants.jsonl"):
    r = json.loads(l)
    if r.get("status") != "started":
        print(r["env"], r.get("ok"), r.get("ms_per_step"), r.get("scatter_ms_per_step"), r.get("sort_ms_per_step"), r.get("scatter_GBps"), r.get("error", ""))
PY
[ -f gpurun_out/best_env.sh ] && . gpurun_out/best_env.sh && cat gpurun_out/best_env.sh
env | grep B2GPU
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -12
timeout 400 python bench.py --steps 3 --warmup 3 --stage-times --no-decode > gpurun_out/bench_s2_text.json 2> gpurun_out/bench_s2_text.err; tail -3 gpurun_out/bench_s2_text.err
grep "^{" gpurun_out/bench_s2_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['sort_stage']['frac'], d['parity']['timed_output_equals_oracle_golden'], d['parity']['device_verify']['ok'], d.get('stage_ms'))"
timeout 300 python bench.py --config entries --entries 20000 --steps 2 --warmup 1 > gpurun_out/bench_s2_entries20k.json 2> gpurun_out/bench_s2_entries20k.err; tail -3 gpurun_out/bench_s2_entries20k.err
grep "^{" gpurun_out/bench_s2_entries20k.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d.get('entries_per_s'), d.get('parity'))"
B2GPU_PM=0 timeout 300 python bench.py --config entries --entries 20000 --steps 2 --warmup 1 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('PM=0', d['value'], d.get('entries_per_s'))"
timeout 300 python bench.py --config mixed --size-mb 256 --steps 3 --warmup 2 --no-decode > gpurun_out/bench_s2_mixed256.json 2> gpurun_out/bench_s2_mixed256.err; tail -3 gpurun_out/bench_s2_mixed256.err
grep "^{" gpurun_out/bench_s2_mixed256.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['timed_output_equals_oracle_golden'])"
# ncu: launch list of one bench step, full capture of one scatter launch
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_text1g.csv python bench.py --steps 1 --warmup 0 --no-decode --cpu-sample-mb 0.25 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
python tools/kernel_times.py gpurun_out/s2_launches_text1g.csv | head -30
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_scatter --launch-skip 2 --launch-count 1 -o gpurun_out/s2_scatter_full python tools/stage_times.py --kind markov --size-mb 256 --no-stage --cpu-gen > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out | head -40
