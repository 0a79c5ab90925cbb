mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -25) > gpurun_out/r2_t_dense.log 2>&1
(timeout 600 python -m pytest tests/test_sampling_gpu.py tests/test_predict_fit_gpu.py -q -m gpu 2>&1 | tail -40) > gpurun_out/r2_t_sampling.log 2>&1
(timeout 300 python tools/time_wide.py 1000 1 3 1 2>&1 | tail -5) > gpurun_out/r2_wide_1000.log 2>&1
(timeout 300 python tools/time_wide.py 100 1 5 1 2>&1 | tail -5) > gpurun_out/r2_wide_100.log 2>&1
(timeout 300 python tools/time_wide.py 2000 1 2 1 2>&1 | tail -5) > gpurun_out/r2_wide_2000.log 2>&1
(timeout 300 python tools/time_wide.py 1000 8 2 1 2>&1 | tail -5) > gpurun_out/r2_wide_1000x8.log 2>&1
tail -n 12 gpurun_out/r2_t_dense.log gpurun_out/r2_t_sampling.log gpurun_out/r2_wide_*.log
