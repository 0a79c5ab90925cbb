mkdir -p gpurun_out
(timeout 20 ./tools/microbench/block_kernels_test 0) > gpurun_out/r2_block_kernels.log 2>&1 || { echo "block kernel harness failed"; cat gpurun_out/r2_block_kernels.log; exit 1; }
(timeout 120 python tools/time_wide.py 1000 1 3 1 2>&1 | tail -5) > gpurun_out/r2_wide_1000.log 2>&1
cat gpurun_out/r2_block_kernels.log gpurun_out/r2_wide_1000.log
grep -q '"mode"' gpurun_out/r2_wide_1000.log || { echo "wide run failed"; exit 1; }
(timeout 600 python -m pytest tests/test_dense_gpu.py tests/test_predict_fit_gpu.py tests/test_sparse_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_t_dense.log 2>&1
(BGP_LIB_NAME=libbgp_b200_timing.so timeout 120 python tools/wide_timers.py 1000 2>&1 | tail -60) > gpurun_out/r2_wide_timers.log 2>&1
(timeout 120 python tools/time_wide.py 100 1 5 1 2>&1 | tail -5) > gpurun_out/r2_wide_100.log 2>&1
(timeout 120 python tools/time_wide.py 2000 1 2 1 2>&1 | tail -5) > gpurun_out/r2_wide_2000.log 2>&1
(timeout 120 python tools/time_wide.py 1000 148 2 0 2>&1 | tail -5) > gpurun_out/r2_narrow_1000x148.log 2>&1
(timeout 200 python tools/time_sparse.py 8 c4 2>&1 | tail -3) > gpurun_out/r2_c4.log 2>&1
cat gpurun_out/r2_t_dense.log; head -14 gpurun_out/r2_wide_timers.log; cat gpurun_out/r2_wide_100.log gpurun_out/r2_wide_2000.log gpurun_out/r2_narrow_1000x148.log gpurun_out/r2_c4.log
