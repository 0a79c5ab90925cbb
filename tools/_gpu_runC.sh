mkdir -p gpurun_out
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo rc=$? >> gpurun_out/r02_bench_8gpu.err)
tail -3 gpurun_out/r02_bench_8gpu.err; head -c 3500 gpurun_out/r02_bench_8gpu.json
