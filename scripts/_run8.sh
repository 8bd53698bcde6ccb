cd /root/repo
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log
# ncu: full capture of the dominant kernel per workload (after the plain run above exited 0)
for w in L L4 P5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 8 -c 1 -f -o gpurun_out/prof_r2_bench_$w python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_bench_$w.log 2>&1; echo "ncu $w rc=$?"
done
# the Jet-free (cost-only) variant through kbench
timeout 300 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_r2_costonly ./build/kbench/kb_cur 13682 4456117 28987644 0 0 costonly > gpurun_out/ncu_r2_costonly.log 2>&1; echo "ncu cost rc=$?"
# launch list of the headline bench
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_L.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_launches.log 2>&1; echo "launch list rc=$?"
# sanitizers on the parity tests (small problems)
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bal or golden or subset or pose" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.log
python scripts/d2h_ceiling.py --mb 6000 > gpurun_out/r2_d2h_ceiling_1gpu.json 2> gpurun_out/r2_d2h_ceiling_1gpu.err; cat gpurun_out/r2_d2h_ceiling_1gpu.json
