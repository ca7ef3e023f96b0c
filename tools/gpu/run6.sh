cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2f.log
cat gpurun_out/pytest_gpu_r2f.log
timeout 600 python bench.py --steps 3 --warmup 2 --stage-times --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2f_text.json 2> gpurun_out/bench_r2f_text.err; tail -3 gpurun_out/bench_r2f_text.err; cat gpurun_out/bench_r2f_text.json
