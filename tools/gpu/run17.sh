# round 2, session 2, call 2 (the last one): k_scatter3 against k_scatter2, then tests / bench / ncu evidence under the winner
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/variants.jsonl gpurun_out/best_env.sh
for i in 1 2 3; do
  timeout 200 python tools/gpu_variants.py 1024 3 second > gpurun_out/variants2_run$i.log 2>&1
  echo "variants run $i rc=$?"; tail -1 gpurun_out/variants2_run$i.log | cut -c1-400
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
timeout 360 python -m pytest tests -m gpu -q --maxfail=6 2>&1 | tail -8
timeout 200 python bench.py --steps 3 --warmup 3 --stage-times --no-decode > gpurun_out/bench_s3_text.json 2> gpurun_out/bench_s3_text.err; tail -3 gpurun_out/bench_s3_text.err
grep "^{" gpurun_out/bench_s3_text.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['sort_stage']['frac'], d['parity']['timed_output_equals_oracle_golden'], d['parity']['device_verify']['ok'], d.get('stage_ms'))"
timeout 170 ncu --set full --import-source on --clock-control none -k regex:k_scatter --launch-skip 2 --launch-count 1 -o gpurun_out/s3_scatter_full python tools/one_encode.py 256 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
timeout 260 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/s3_launches_text256.csv python tools/one_encode.py 256 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log | cut -c1-300
python tools/kernel_times.py gpurun_out/s3_launches_text256.csv | head -16
timeout 150 python bench.py --config mixed --size-mb 256 --steps 3 --warmup 2 --no-decode > gpurun_out/bench_s3_mixed256.json 2> gpurun_out/bench_s3_mixed256.err; tail -3 gpurun_out/bench_s3_mixed256.err
grep "^{" gpurun_out/bench_s3_mixed256.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['timed_output_equals_oracle_golden'])"
timeout 150 python bench.py --config entries --entries 20000 --steps 2 --warmup 1 > gpurun_out/bench_s3_entries20k.json 2> gpurun_out/bench_s3_entries20k.err
grep "^{" gpurun_out/bench_s3_entries20k.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d.get('entries_per_s'), d.get('parity'))"
ls -la gpurun_out | head -30
