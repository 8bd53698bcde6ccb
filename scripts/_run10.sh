cd /root/repo
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
python scripts/bench_solve.py --shape L > gpurun_out/r2_bench_solve_L.json 2> gpurun_out/r2_bench_solve_L.err; echo "bench_solve rc=$?"; cat gpurun_out/r2_bench_solve_L.json
for w in P5 L4 L; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"; done
for w in L L4 P5; do python -c "
import json
d=json.load(open('gpurun_out/r2_bench_$w.json')); r=d['roofline']
print('$w kernel %.3f device %.3f frac %.3f cost_only %.3f e2e %.1f'%(d['kernel_ms_per_step'], d['device_ms_per_step'], r['frac'], r['cost_only_kernel_ms'], d['e2e']['ms_per_step']))"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:NormalProduct --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_normal python scripts/bench_solve.py --shape L --iterations 3 > gpurun_out/ncu_r2_normal.log 2>&1; echo "ncu normal rc=$?"
CB200_SOLVER_TIMING=1 timeout 900 ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=6 > gpurun_out/r2_solve_L_device_resident.txt 2>&1; echo "solve rc=$?"; grep -E "iteration [0-9]|Linear solver|Minimizer  |Jacobian &|Residual only" gpurun_out/r2_solve_L_device_resident.txt
