mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_predict_fit_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_t_dense.log 2>&1
(BGP_LIB_NAME=libbgp_b200_timing.so timeout 300 python tools/wide_timers.py 1000 2>&1 | tail -60) > gpurun_out/r2_wide_timers.log 2>&1
(timeout 300 python tools/time_wide.py 1000 1 3 1 2>&1 | tail -5) > gpurun_out/r2_wide_1000.log 2>&1
(timeout 300 python tools/time_wide.py 100 1 5 1 2>&1 | tail -5) > gpurun_out/r2_wide_100.log 2>&1
(timeout 300 python tools/time_wide.py 2000 1 2 1 2>&1 | tail -5) > gpurun_out/r2_wide_2000.log 2>&1
(timeout 300 python tools/time_wide.py 1000 148 2 0 2>&1 | tail -5) > gpurun_out/r2_narrow_1000x148.log 2>&1
cat gpurun_out/r2_t_dense.log; head -14 gpurun_out/r2_wide_timers.log; cat gpurun_out/r2_wide_1000.log gpurun_out/r2_wide_100.log gpurun_out/r2_wide_2000.log gpurun_out/r2_narrow_1000x148.log
