cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:EvaluateKernel --launch-skip 4 -c 1 -o gpurun_out/prof_r2_c4 -f ./build/kbench/kb_c4 13682 4456117 28987644 1 1 c4 > gpurun_out/ncu_r2_c4.log 2>&1
tail -3 gpurun_out/ncu_r2_c4.log
