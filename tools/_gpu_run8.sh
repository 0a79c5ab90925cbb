mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_t_sharded_2gpu.log 2>&1
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo rc=$? >> gpurun_out/r2_bench_2gpu.err)
cat gpurun_out/r2_t_sharded_2gpu.log; tail -3 gpurun_out/r2_bench_2gpu.err; head -c 2500 gpurun_out/r2_bench_2gpu.json
