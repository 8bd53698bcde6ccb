cd /root/repo
./build/kbench/kb_cur 13682 4456117 28987644 1 1 plain | grep "KBENCH\|checksums"
KBENCH_UNIFORM=1 KBENCH_CHUNK=992 ./build/kbench/kb_cur 13682 4456117 28987644 1 1 uniform992 | grep "KBENCH\|checksums"
timeout 200 python -m pytest tests/test_full_size_properties.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2y_bench_L.json; python -c "
import json
d=json.load(open('gpurun_out/r2y_bench_L.json')); r=d['roofline']
print('L kernel %.3f device %.3f wall %.3f frac %.3f e2e %.1f cost %s'%(d['kernel_ms_per_step'], d['device_ms_per_step'], d['ms_per_step'], r['frac'], d['e2e']['ms_per_step'], d['cost']))"
