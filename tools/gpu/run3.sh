set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/pytest_gpu_r2c.log
cat gpurun_out/pytest_gpu_r2c.log
timeout 600 python bench.py --config mixed --size-mb 1024 --steps 2 --warmup 1 --stage-times --no-decode > gpurun_out/bench_r2c_mixed1g.json 2> gpurun_out/bench_r2c_mixed1g.err; tail -3 gpurun_out/bench_r2c_mixed1g.err; cat gpurun_out/bench_r2c_mixed1g.json
timeout 600 python bench.py --config entries --entries 20000 --steps 2 --warmup 1 > gpurun_out/bench_r2c_entries.json 2> gpurun_out/bench_r2c_entries.err; tail -3 gpurun_out/bench_r2c_entries.err; cat gpurun_out/bench_r2c_entries.json
timeout 600 ncu -k regex:^k_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2c_text1g.csv python bench.py --steps 1 --warmup 0 --no-decode > /dev/null 2>&1
