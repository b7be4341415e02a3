mkdir -p gpurun_out/r3c; o=gpurun_out/r3c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $o/bench_n8.json 2> $o/bench_n8.err; tail -2 $o/bench_n8.err; cut -c1-400 $o/bench_n8.json
