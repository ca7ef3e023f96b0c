cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 2 --stage-times --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2g_text.json 2> gpurun_out/bench_r2g_text.err; tail -3 gpurun_out/bench_r2g_text.err; cat gpurun_out/bench_r2g_text.json
timeout 600 ncu -k regex:^k_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2g_text1g.csv python bench.py --steps 1 --warmup 0 --no-decode --cpu-sample-mb 0.25 > /dev/null 2>&1
