cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2e_text_n$N.json 2> gpurun_out/bench_r2e_text_n$N.err
tail -5 gpurun_out/bench_r2e_text_n$N.err; cat gpurun_out/bench_r2e_text_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config mixed --steps 2 --warmup 1 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2e_mixed_n$N.json 2> gpurun_out/bench_r2e_mixed_n$N.err
tail -5 gpurun_out/bench_r2e_mixed_n$N.err; cat gpurun_out/bench_r2e_mixed_n$N.json
