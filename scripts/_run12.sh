cd /root/repo
for v in cur nosq A4 C4 C3 pass2_3 pass2_A4 A4noseg; do ./build/kbench/kb_$v 13682 4456117 28987644 1 1 $v 2>&1 | grep "KBENCH\|smem plan\|checksums"; done
./build/kbench/kb_cur 13682 4456117 28987644 0 0 cur_costonly | grep KBENCH
./build/kbench/kb_A4 13682 4456117 28987644 0 0 A4_costonly | grep KBENCH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 5 -c 1 -f -o gpurun_out/prof_r2c_kb_cur ./build/kbench/kb_cur > gpurun_out/ncu_r2c_kb_cur.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 5 -c 1 -f -o gpurun_out/prof_r2c_kb_A4 ./build/kbench/kb_A4 > gpurun_out/ncu_r2c_kb_A4.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c_pytest_gpu.log
python bench.py --workload L --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_L.json 2> gpurun_out/r2c_bench_L.err; echo "bench L rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2c_bench_L.json')); r=d['roofline']
print('L kernel %.3f device %.3f frac %.3f cost_only %.3f e2e %.1f'%(d['kernel_ms_per_step'], d['device_ms_per_step'], r['frac'], r['cost_only_kernel_ms'], d['e2e']['ms_per_step']))"
