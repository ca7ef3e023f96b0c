cd $GRAFT_REPO_ROOT
N=${1:-4}
timeout 300 python bench.py --steps 3 --warmup 2 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2h_text_n1.json 2> gpurun_out/bench_r2h_text_n1.err; tail -2 gpurun_out/bench_r2h_text_n1.err; cut -c1-400 gpurun_out/bench_r2h_text_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2h_text_n$N.json 2> gpurun_out/bench_r2h_text_n$N.err
grep -v ProcessGroupNCCL gpurun_out/bench_r2h_text_n$N.err | tail -5; grep "^{" gpurun_out/bench_r2h_text_n$N.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config mixed --steps 2 --warmup 1 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2h_mixed_n$N.json 2> gpurun_out/bench_r2h_mixed_n$N.err
grep -v ProcessGroupNCCL gpurun_out/bench_r2h_mixed_n$N.err | tail -5; grep "^{" gpurun_out/bench_r2h_mixed_n$N.json | cut -c1-400
