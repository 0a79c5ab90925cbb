mkdir -p gpurun_out
CMD1="python bench.py --steps 2 --warmup 1 --no-e2e --no-fit --no-sparse --no-c5 --no-cpu --no-peak"
CMD2="python tools/time_wide.py 1000 1 1 1"
$CMD1 > gpurun_out/r02_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD1 > gpurun_out/r02_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD2 > gpurun_out/r02_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_wide_launches.csv $CMD2 > gpurun_out/r02_ncu2.log 2>&1
echo "wide launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_wadj|k_wchol" -c 3 -o gpurun_out/r02_wide_full $CMD2 > gpurun_out/r02_ncu3.log 2>&1
echo "wide full rc=$?"
ls -la gpurun_out/ | tail -12
