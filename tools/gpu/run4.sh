cd $GRAFT_REPO_ROOT
for cfg in "2 805306368" "3 536870912" "4 402653184" "2 1610612736"; do set -- $cfg
  B2GPU_PIPELINE=$1 B2GPU_BATCH_POSITIONS=$2 timeout 300 python bench.py --steps 2 --warmup 1 --no-decode --cpu-sample-mb 0.25 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('PIPELINE=$1 BATCH=$2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity']['timed_output_equals_oracle_golden'])
" | tee -a gpurun_out/pipe_r2d.log
done
