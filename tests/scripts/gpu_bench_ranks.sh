# per-rank times of the N-rank bench (members: C2 and line-shuffled C2)
N=${1:-2}
X3_BENCH_RANKS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json
grep "^rank" gpurun_out/bench_n$N.err; cut -c1-220 gpurun_out/bench_n$N.json
