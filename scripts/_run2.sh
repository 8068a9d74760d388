mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-2} --steps 10 --warmup 3 > gpurun_out/bench_${NG:-2}gpu.log 2> gpurun_out/bench_${NG:-2}gpu.err; echo "bench2 rc=$?"; tail -c 600 gpurun_out/bench_${NG:-2}gpu.log; tail -3 gpurun_out/bench_${NG:-2}gpu.err
