set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
for v in base c3 c4 noseg nostg; do ./build/kbench/kb_$v 13682 4456117 28987644 1 1 $v 2>&1 | tail -2; done > gpurun_out/r2_kbench1.log 2>&1
./build/kbench/kb_c4 13682 4456117 28987644 1 1 c4_table 0 2>&1 | tail -2 >> gpurun_out/r2_kbench1.log
./build/kbench/kb_c3 13682 4456117 28987644 1 1 c3_table 0 2>&1 | tail -2 >> gpurun_out/r2_kbench1.log
./build/kbench/kb_c4 13682 4456117 28987644 0 0 c4_costonly 2>&1 | tail -1 >> gpurun_out/r2_kbench1.log
./build/kbench/kb_base 13682 4456117 28987644 0 0 base_costonly 2>&1 | tail -1 >> gpurun_out/r2_kbench1.log
cat gpurun_out/r2_kbench1.log
