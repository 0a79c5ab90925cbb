mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_sharded_gpu.py tests/test_fit_batch_gpu.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/r2_t_dense.log 2>&1
cat gpurun_out/r2_t_dense.log
grep -q "passed" gpurun_out/r2_t_dense.log || exit 1
(timeout 120 python tools/time_wide.py 1000 1 3 1 2>&1 | tail -5) > gpurun_out/r2_wide_1000.log 2>&1
(timeout 120 python tools/time_wide.py 1000 148 2 0 2>&1 | tail -5) > gpurun_out/r2_narrow_1000x148.log 2>&1
(BGP_LIB_NAME=libbgp_b200_timing.so timeout 120 python tools/wide_timers.py 1000 2>&1 | tail -60) > gpurun_out/r2_wide_timers.log 2>&1
cat gpurun_out/r2_wide_1000.log gpurun_out/r2_narrow_1000x148.log; grep -A5 "ADIAG" gpurun_out/r2_wide_timers.log
CMD2="python tools/time_wide.py 1000 1 1 1"
$CMD2 > gpurun_out/r02_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_wide_launches.csv $CMD2 > gpurun_out/r02_ncu2.log 2>&1
echo "wide launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_wadj|k_wchol" -c 3 -o gpurun_out/r02_wide_full $CMD2 > gpurun_out/r02_ncu3.log 2>&1
echo "wide full rc=$?"
