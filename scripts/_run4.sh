cd /root/repo
for v in T128x3 T160x3 T96x5 T128x2 T64x7 T96x4; do ./build/kbench/kb_$v 13682 4456117 28987644 1 1 $v 2>&1 | tail -1; done > gpurun_out/r2_kbench3.log 2>&1
grep KBENCH gpurun_out/r2_kbench3.log
