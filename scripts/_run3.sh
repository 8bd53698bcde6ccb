cd /root/repo
for v in A A3 B B3 C D3; do ./build/kbench/kb_$v 13682 4456117 28987644 1 1 $v 2>&1 | tail -2; done > gpurun_out/r2_kbench2.log 2>&1
./build/kbench/kb_B 13682 4456117 28987644 0 0 B_costonly 2>&1 | tail -1 >> gpurun_out/r2_kbench2.log
./build/kbench/kb_A 13682 4456117 28987644 0 0 A_costonly 2>&1 | tail -1 >> gpurun_out/r2_kbench2.log
./build/kbench/kb_B 13682 4456117 28987644 1 1 B_table 0 2>&1 | tail -1 >> gpurun_out/r2_kbench2.log
grep KBENCH gpurun_out/r2_kbench2.log
