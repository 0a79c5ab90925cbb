mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dense_gpu.py -x -q -m gpu 2>&1 | tail -25) > gpurun_out/r2_t_dense.log 2>&1
(timeout 300 python tools/time_wide.py 1000 1 3 2>&1 | tail -5) > gpurun_out/r2_wide_1000.log 2>&1
(timeout 300 python tools/time_wide.py 100 1 5 2>&1 | tail -5) > gpurun_out/r2_wide_100.log 2>&1
(timeout 300 python tools/time_wide.py 2000 1 2 2>&1 | tail -5) > gpurun_out/r2_wide_2000.log 2>&1
(timeout 300 python tools/time_wide.py 1000 20 2 2>&1 | tail -5) > gpurun_out/r2_wide_1000x20.log 2>&1
(timeout 900 python -m pytest tests/test_sparse_gpu.py tests/test_c1_fitmodel_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -25) > gpurun_out/r2_t_misc.log 2>&1
tail -n 8 gpurun_out/r2_t_dense.log gpurun_out/r2_wide_*.log gpurun_out/r2_t_misc.log
