mkdir -p gpurun_out
(BGP_LIB_NAME=libbgp_b200_timing.so timeout 300 python tools/wide_timers.py 1000 2>&1 | tail -30) > gpurun_out/r2_wide_timers.log 2>&1
(timeout 1500 python -m pytest tests/test_predict_fit_gpu.py tests/test_fit_batch_gpu.py tests/test_fitmodel_gpu.py tests/test_reference_run_gpu.py tests/test_c1_fitmodel_gpu.py -q -m gpu 2>&1 | tail -25) > gpurun_out/r2_t_rest.log 2>&1
tail -n 30 gpurun_out/r2_wide_timers.log gpurun_out/r2_t_rest.log
