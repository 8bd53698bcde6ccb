#!/bin/bash
# One round's measurements on a single B200 (run under gpurun): GPU tests, smoke, benches of every
# workload (both arms), device-resident solve, ncu captures of the headline and product kernels and
# the launch list.  Outputs go to gpurun_out/<tag>_*; scripts/ncu_traffic.py turns the captures into
# profiles/*_ncu_summary.txt and profiles/r2_traffic.json.   usage: measure_round.sh <tag>
cd /root/repo
T=${1:-r2z}
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_L.json 2> gpurun_out/${T}_bench_L.err; echo "bench L rc=$?"
for w in M L4 P5; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "bench $w rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; echo "reference arm rc=$?"
for w in L M L4 P5; do python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_$w.json')); r=d['roofline']; c=d.get('cpu_baseline') or {}
print('$w kernel %.3f device %.3f wall %.3f frac %.3f cost_only %.3f e2e %.1f cpu %s'%(d['kernel_ms_per_step'], d['device_ms_per_step'], d['ms_per_step'], r['frac'], r['cost_only_kernel_ms'], d['e2e']['ms_per_step'], c.get('value')))"; done
python scripts/bench_solve.py --shape L > gpurun_out/${T}_bench_solve_L.json 2> gpurun_out/${T}_bench_solve_L.err; echo "bench_solve rc=$?"; cat gpurun_out/${T}_bench_solve_L.json
CB200_SOLVER_TIMING=1 timeout 600 ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=6 > gpurun_out/${T}_solve_L.txt 2>&1; echo "solve rc=$?"; grep -E "Linear solver|Minimizer  |Preprocessor" gpurun_out/${T}_solve_L.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 8 -c 1 -f -o gpurun_out/prof_${T}_bench_L python bench.py --workload L --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${T}_bench_L.log 2>&1; echo "ncu L rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:NormalProduct --launch-skip 3 -c 1 -f -o gpurun_out/prof_${T}_normal python scripts/bench_solve.py --shape L --iterations 3 > gpurun_out/ncu_${T}_normal.log 2>&1; echo "ncu normal rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_L.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${T}_launches.log 2>&1; echo "launch list rc=$?"
