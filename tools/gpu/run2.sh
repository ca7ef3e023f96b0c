set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_cfg_r2b.log
cat gpurun_out/pytest_cfg_r2b.log
timeout 600 python bench.py --steps 2 --warmup 1 --stage-times > gpurun_out/bench_r2b_text.json 2> gpurun_out/bench_r2b_text.err; tail -3 gpurun_out/bench_r2b_text.err; cat gpurun_out/bench_r2b_text.json
timeout 500 ncu -k regex:^k_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b_markov.csv python tools/stage_times.py --kind markov --size-mb 128 --no-stage --cpu-gen > /dev/null 2>&1
timeout 300 ncu -k regex:^k_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b_random.csv python tools/stage_times.py --kind random --size-mb 48 --no-stage --cpu-gen > /dev/null 2>&1
timeout 300 ncu -k regex:^k_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b_sparse.csv python tools/stage_times.py --kind sparse --size-mb 128 --no-stage --cpu-gen > /dev/null 2>&1
