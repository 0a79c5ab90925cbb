mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12) > gpurun_out/r2_t_all.log 2>&1
cat gpurun_out/r2_t_all.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r2_smoke.log 2>&1
cat gpurun_out/r2_smoke.log
(timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo rc=$? >> gpurun_out/r02_bench.err)
tail -3 gpurun_out/r02_bench.err; head -c 1500 gpurun_out/r02_bench.json
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo rc=$? >> gpurun_out/r02_bench_reference.err)
cat gpurun_out/r02_bench_reference.json
