cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2m_text_n$N.json 2> gpurun_out/bench_r2m_text_n$N.err
grep -v ProcessGroupNCCL gpurun_out/bench_r2m_text_n$N.err | tail -5; grep "^{" gpurun_out/bench_r2m_text_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']); print(json.dumps(d['shard_phases_ms_per_rank']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config mixed --steps 3 --warmup 2 --no-decode --cpu-sample-mb 0.25 > gpurun_out/bench_r2m_mixed_n$N.json 2> gpurun_out/bench_r2m_mixed_n$N.err
grep -v ProcessGroupNCCL gpurun_out/bench_r2m_mixed_n$N.err | tail -5; grep "^{" gpurun_out/bench_r2m_mixed_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']); print(json.dumps(d['shard_phases_ms_per_rank']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config entries --steps 2 --warmup 1 > gpurun_out/bench_r2m_entries_n$N.json 2> gpurun_out/bench_r2m_entries_n$N.err
grep -v ProcessGroupNCCL gpurun_out/bench_r2m_entries_n$N.err | tail -5; grep "^{" gpurun_out/bench_r2m_entries_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['entries_per_s'], d['stored_entries'], d['parity'])"
