#!/bin/bash
# Multi-GPU check and bench on N GPUs of one box (gpurun --gpus N): sharded vs unsharded
# evaluation at 5 % scale, then bench.py; `all` as second argument also times the separate-push
# exchange.   usage: measure_multigpu.sh <N> [all]
cd /root/repo
export MASTER_ADDR=127.0.0.1
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR scripts/check_multigpu.py 0.05 > gpurun_out/r2k_check_multigpu_${N}gpu.log 2>&1; echo "check rc=$?"; grep -E "rank 0|MISMATCH|Error|error" gpurun_out/r2k_check_multigpu_${N}gpu.log | head -12
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2k_bench_L_${N}gpu.json 2> gpurun_out/r2k_bench_L_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2k_bench_L_${N}gpu.err
if [ "$2" = "all" ]; then
CB200_NO_CHUNKED_KERNEL=1 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-multi-gpu-check > gpurun_out/r2k_bench_L_${N}gpu_unfused.json 2> gpurun_out/r2k_bench_L_${N}gpu_unfused.err; echo "bench unfused rc=$?"
fi
for f in L_${N}gpu L_${N}gpu_unfused; do [ -s gpurun_out/r2k_bench_$f.json ] && python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_$f.json'))
print('$f', 'kernel %.3f device %.3f ms/step %.3f e2e %.1f check %s'%(d['kernel_ms_per_step'], d['device_ms_per_step'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('multi_gpu_check')))"; done
