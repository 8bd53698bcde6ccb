cd /root/repo
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu.log
python bench.py --workload L4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_L4.json 2> gpurun_out/r2_bench_L4.err; echo "bench L4 rc=$?"
python bench.py --workload P5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_P5.json 2> gpurun_out/r2_bench_P5.err; echo "bench P5 rc=$?"
python bench.py --workload L --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_L.json 2> gpurun_out/r2_bench_L.err; echo "bench L rc=$?"
for w in L L4 P5; do python -c "
import json
d=json.load(open('gpurun_out/r2_bench_$w.json')); r=d['roofline']
print('$w kernel %.3f device %.3f frac %.3f cost_only %.3f e2e %.1f'%(d['kernel_ms_per_step'], d['device_ms_per_step'], r['frac'], r['cost_only_kernel_ms'], d['e2e']['ms_per_step']))"; done
