cd /root/repo
for v in old prev cur gd0 p64; do KBENCH_GENERIC=1 ./build/kbench/kb_$v 13682 4456117 28987644 1 1 generic_$v 2>&1 | grep "KBENCH\|checksums"; done
for v in prev cur; do ./build/kbench/kb_$v 13682 4456117 28987644 1 1 plain_$v 2>&1 | grep "KBENCH"; done
for v in prev cur; do ./build/kbench/kb_$v 13682 4456117 28987644 0 0 costonly_$v 2>&1 | grep "KBENCH"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 5 -c 1 -f -o gpurun_out/prof_r2e_kb_generic env KBENCH_GENERIC=1 ./build/kbench/kb_cur > gpurun_out/ncu_r2e_kb_generic.log 2>&1; echo "ncu generic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 5 -c 1 -f -o gpurun_out/prof_r2e_kb_costonly ./build/kbench/kb_cur 13682 4456117 28987644 0 0 > gpurun_out/ncu_r2e_kb_costonly.log 2>&1; echo "ncu costonly rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest_gpu.log
for w in P5 L4; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_$w.json 2> gpurun_out/r2e_bench_$w.err; echo "bench $w rc=$?"; done
for w in L4 P5; do python -c "
import json
d=json.load(open('gpurun_out/r2e_bench_$w.json')); r=d['roofline']
print('$w kernel %.3f device %.3f frac %.3f cost_only %.3f e2e %.1f'%(d['kernel_ms_per_step'], d['device_ms_per_step'], r['frac'], r['cost_only_kernel_ms'], d['e2e']['ms_per_step']))"; done
