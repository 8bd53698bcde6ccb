cd /root/repo
for op in 0 1 2 3 4; do KBENCH_NORMAL=1 KBENCH_NORMAL_OP=$op ./build/kbench/kb_cur 13682 4456117 28987644 1 1 op$op 2>&1 | grep "KNORMAL"; done
CB200_VERBOSE=1 timeout 900 python -m pytest tests/test_device_linear_algebra.py tests/test_lm.py tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|cb200:|Error" | sort | uniq -c | head -20
CB200_VERBOSE=1 CB200_SOLVER_TIMING=1 timeout 600 ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=6 > gpurun_out/r2j_solve_L.txt 2>&1; echo "solve rc=$?"; grep -E "cb200|iteration [0-9]|Linear solver|Minimizer  |Preprocessor|^ +[0-9] " gpurun_out/r2j_solve_L.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2j_solve_launches.csv ./build/examples/bundle_adjuster --synthetic=13682,4456117,28987644 --robustify --bulk --linear_solver=cgnr_cuda --num_iterations=1 > gpurun_out/r2j_solve_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_solve_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:45]:
    print(r[ki][:80].ljust(80), r[vi], r[ui])
PY
