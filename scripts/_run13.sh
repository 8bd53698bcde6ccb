cd /root/repo
for v in p128 p64; do ./build/kbench/kb_$v 13682 4456117 28987644 1 1 $v 2>&1 | grep "KBENCH\|checksums"; done
for v in nbase ncur; do KBENCH_NORMAL=1 ./build/kbench/kb_$v 13682 4456117 28987644 1 1 $v 2>&1 | grep "KNORMAL"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:NormalProduct --launch-skip 5 -c 1 -f -o gpurun_out/prof_r2d_kb_normal env KBENCH_NORMAL=1 ./build/kbench/kb_ncur > gpurun_out/ncu_r2d_kb_normal.log 2>&1; echo "ncu normal rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest_gpu.log
python scripts/bench_solve.py --shape L > gpurun_out/r2d_bench_solve_L.json 2> gpurun_out/r2d_bench_solve_L.err; echo "bench_solve rc=$?"; cat gpurun_out/r2d_bench_solve_L.json
for w in L P5 L4; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err; echo "bench $w rc=$?"; done
for w in L L4 P5; do python -c "
import json
d=json.load(open('gpurun_out/r2d_bench_$w.json')); r=d['roofline']
print('$w kernel %.3f device %.3f frac %.3f cost_only %.3f e2e %.1f'%(d['kernel_ms_per_step'], d['device_ms_per_step'], r['frac'], r['cost_only_kernel_ms'], d['e2e']['ms_per_step']))"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 8 -c 1 -f -o gpurun_out/prof_r2d_bench_P5 python bench.py --workload P5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2d_bench_P5.log 2>&1; echo "ncu P5 rc=$?"
