mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_sharded_gpu.py tests/test_fitmodel_gpu.py tests/test_c1_fitmodel_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_t_dense.log 2>&1
(BGP_LIB_NAME=libbgp_b200_timing.so timeout 300 python tools/wide_timers.py 1000 2>&1 | tail -60) > gpurun_out/r2_wide_timers.log 2>&1
(timeout 300 python tools/time_wide.py 1000 1 3 1 2>&1 | tail -5) > gpurun_out/r2_wide_1000.log 2>&1
(timeout 300 python tools/time_wide.py 100 1 5 1 2>&1 | tail -5) > gpurun_out/r2_wide_100.log 2>&1
(timeout 300 python tools/time_wide.py 2000 1 2 1 2>&1 | tail -5) > gpurun_out/r2_wide_2000.log 2>&1
(timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_try.json 2> gpurun_out/r2_bench_try.err; echo rc=$? >> gpurun_out/r2_bench_try.err)
cat gpurun_out/r2_t_dense.log gpurun_out/r2_wide_timers.log gpurun_out/r2_wide_1000.log gpurun_out/r2_wide_100.log gpurun_out/r2_wide_2000.log; tail -5 gpurun_out/r2_bench_try.err; head -c 3000 gpurun_out/r2_bench_try.json
