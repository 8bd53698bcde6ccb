cd /root/repo
# one GPU: the copy path with local stand-in peers (checks that nothing breaks and what it costs)
for pe in 1 7; do KBENCH_PEERS=$pe KBENCH_CHUNK=992 ./build/kbench/kb_cur 13682 4456117 28987644 1 1 chunk992_peers$pe | grep "KBENCH\|checksums"; done
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
CB200_CHUNKED_KERNEL=1 timeout 400 $TR scripts/check_multigpu.py 0.05 > gpurun_out/r2l_check_multigpu_2gpu.log 2>&1; echo "check rc=$?"; grep -E "rank 0/2 fmt 0: exchange|MISMATCH|Error|error" gpurun_out/r2l_check_multigpu_2gpu.log | head -6
CB200_CHUNKED_KERNEL=1 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2l_bench_L_2gpu_fused.json 2> gpurun_out/r2l_bench_L_2gpu_fused.err; echo "bench fused rc=$?"; tail -2 gpurun_out/r2l_bench_L_2gpu_fused.err
python -c "
import json
d=json.load(open('gpurun_out/r2l_bench_L_2gpu_fused.json'))
print('fused', 'kernel %.3f device %.3f ms/step %.3f check %s'%(d['kernel_ms_per_step'], d['device_ms_per_step'], d['ms_per_step'], d.get('multi_gpu_check')))"
