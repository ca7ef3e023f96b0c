set -x
cd $GRAFT_REPO_ROOT
for k in "markov 512" "mixed 768" "random 128" "sparse 256" "zipf 256"; do set -- $k; timeout 600 python tools/stage_times.py --kind $1 --size-mb $2 >> gpurun_out/stage_r2a.jsonl 2>> gpurun_out/stage_r2a.err; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a_markov.csv python tools/stage_times.py --kind markov --size-mb 256 --no-stage > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a_mixed.csv python tools/stage_times.py --kind mixed --size-mb 192 --no-stage > /dev/null 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc
