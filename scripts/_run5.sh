cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
python bench.py --workload L --steps 10 --warmup 3 > gpurun_out/r2_bench_L.json 2> gpurun_out/r2_bench_L.err; echo "bench L rc=$?"
python bench.py --workload L4 --steps 10 --warmup 3 > gpurun_out/r2_bench_L4.json 2> gpurun_out/r2_bench_L4.err; echo "bench L4 rc=$?"
python bench.py --workload P5 --steps 10 --warmup 3 > gpurun_out/r2_bench_P5.json 2> gpurun_out/r2_bench_P5.err; echo "bench P5 rc=$?"
python bench.py --workload M --steps 10 --warmup 3 > gpurun_out/r2_bench_M.json 2> gpurun_out/r2_bench_M.err; echo "bench M rc=$?"
cat gpurun_out/r2_bench_*.json | cut -c1-600
